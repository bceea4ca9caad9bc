// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), fed by TMA.
//
//   out = epilogue( A[M,K] * W[N,K]^T )        A, W 16-bit row-major (both K-major; each bf16 or fp16), fp32 accumulate
//
// This one kernel carries every dense contraction of the tower (reference path: the nn.Linear / Conv2d calls inside
// transformers' SiglipEncoderLayer / CLIPEncoderLayer / Blip2EncoderLayer):
//   patch-embedding conv-as-GEMM (+bias +position embedding, row remap past the class token),
//   fused QKV projection (+bias), attention out-proj (+bias +fp32 residual),
//   MLP fc1 (+bias +GELU variant) and fc2 (+bias +fp32 residual).
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      : TMA producer   — 128B-swizzled [128 x 64] A tiles and [BN x 64] W tiles into a STAGES-deep ring
//   warp 1      : MMA issuer     — one lane issues tcgen05.mma 128 x BN x 16, commits free ring slots / signal epilogue
//   warps 2..9  : epilogue       — one thread = one output row (= one TMEM lane): tcgen05.ld a 32-column chunk, fused
//                                  bias / activation / residual in registers, the chunk goes out through a swizzled
//                                  smem box and a TMA store; the fp32 residual chunk comes IN through a 3-deep ring
//                                  of TMA loads issued two chunks ahead. No global address arithmetic, no bounds
//                                  predicates (TMA clips ragged M / N) and no accumulator transpose.
//                                  (GEN = true keeps the older transpose-and-store epilogue for the row-remapped /
//                                  broadcast-residual case the patch embedding needs.)
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// CG = 2 runs the same roles on a CTA PAIR (cluster of two CTAs on one TPC, tcgen05 cta_group::2): the pair owns a
// 256 x BN tile, each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows), the leader CTA's MMA lane
// issues 256 x BN x 16 instructions that read both halves, and each CTA's epilogue drains its own 128 TMEM lanes.
// Per SM that halves the W bytes moved L2 -> smem and smem -> tensor core, which is what bounds the single-CTA form.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int STAGING_PER_WARP = 32 * 128;  // generic epilogue: 32 rows x 32 fp32
// residual epilogue, per warp: ring of RES_NBUF [32 rows x RES_COLS fp32] TMA boxes (64B-swizzled) + one
// [32 rows x 32 cols] 16-bit box for the fused cast of the updated residual stream (next GEMM's A operand)
#ifndef VLMFW_RES_NBUF
#define VLMFW_RES_NBUF 3
#endif
constexpr int RES_NBUF = VLMFW_RES_NBUF;
constexpr int RES_COLS = 16;
constexpr int RES_BOX_BYTES = 32 * RES_COLS * 4;
constexpr int H16_BOX_BYTES = 32 * 32 * 2;
constexpr int SMEM_LIMIT = 232448;          // 227 KB opt-in maximum per CTA

template <int BN, int CG, int EPI, bool GEN>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_ROWS = BN / CG;  // W rows staged by one CTA
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // per-warp epilogue staging: generic 4 KB transpose buffer; 16-bit TMA-store path 2 x [32 x 64 B] boxes;
    // residual path RES_NBUF x [32 x RES_COLS fp32] boxes + the 16-bit cast box
    static constexpr int STG_PER_WARP =
        GEN ? STAGING_PER_WARP
            : (EPI == VLMFW_EPI_BIAS_RESID_F32 ? RES_NBUF * RES_BOX_BYTES + H16_BOX_BYTES
                                               : (EPI == VLMFW_EPI_BIAS_RESID_16 ? RES_NBUF * H16_BOX_BYTES : 2 * 2048));
    static constexpr int STAGING_BYTES = EPI_WARPS * STG_PER_WARP;
    static constexpr int BAR_BYTES = 512;
    static_assert((2 * 6 + 4 + EPI_WARPS * RES_NBUF) * 8 + 4 <= BAR_BYTES, "barrier block too small for the residual ring");
    static constexpr int RING_MAX = (SMEM_LIMIT - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = RING_MAX > 6 ? 6 : RING_MAX;
    static_assert(STAGES >= 2, "operand ring needs at least two stages");
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
};

struct KParams {
    int m, n, k;
    uint32_t idesc;  // tcgen05 instruction descriptor (the operand format is a run-time choice)
    int out_f16;     // 16-bit epilogue output format
    // ragged last N tile of a CTA-pair GEMM: its real width rounded up to 16 (0 = the last tile is full / feature off) and
    // the instruction descriptor of that width. SigLIP's N = 1152 = 4.5 tiles: a full-width MMA on a half-empty tile is
    // 10 % of the kernel's tensor work (and power) spent on zero padding.
    int tail_n;
    uint32_t idesc_tail;
    const float* bias;
    void* out;
    int ldo;
    const float* resid;
    int ldr;
    int resid_mod;
    int grp, grp_out, grp_off;
    // LayerNorm folded into the consumer (16-bit epilogue): out = act(rstd[m] acc + bias[n]); the weights carry gamma
    // and the mean subtraction (rows centred: x (I - 11^T/K) W'^T == x (W' - rowmean(W'))^T), rstd comes from the
    // (sum, sum of squares) partials the producer of A's fp32 source left in ln_stats
    const float2* ln_stats;
    int ln_np, ln_row_mul;
    float ln_inv_dim, ln_eps;
    // ... and its producer side (residual epilogue): 16-bit copy of the updated rows through tmap_h16 and per-row
    // (sum, sum of squares) over this warp's 128 / 96 / 64 columns at stats_out[row * stats_np + 2 * n_tile + half]
    float2* stats_out;
    int stats_np;
    int h16_on;
    // scale of the accumulator: the 16-bit W operand holds W / acc_scale when a weight exceeded the fp16 range
    // (power of two, so the product is exact); 1 otherwise
    float acc_scale;
    // fp16 outputs saturate to +-65504 (cvt.rn.satfinite) instead of overflowing; *sat counts the (row, chunk) stores in
    // which at least one value was clamped, so the host can tell that a forward clipped (NULL: not counted)
    unsigned long long* sat;
    int dbg;  // VLMFW_GEMM_DBG bit mask (bottleneck experiments only): 1 no TMA after the ring is primed, 2 no epilogue
              // global traffic, 4 no MMA issue
};

template <int ACT>
__device__ __forceinline__ float activation(float x) {
    if constexpr (ACT == VLMFW_ACT_GELU_TANH) {
        // 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3)   (activations.py gelu_pytorch_tanh)
        // One MUFU (tanh.approx, |rel err| <= 2^-11: half the rounding step of the 16-bit output it feeds) and five
        // FP32 ops; the sigmoid form needs two MUFUs and the SFU pipe is what paces this epilogue (ncu: 'mio' stalls).
        constexpr float C0 = 0.7978845608028654f, C1 = 0.7978845608028654f * 0.044715f;
        const float u = x * fmaf(x * x, C1, C0);
        const float hx = 0.5f * x;
        return fmaf(hx, tanh_approx(u), hx);
    } else if constexpr (ACT == VLMFW_ACT_QUICK_GELU) {
        // x * sigmoid(1.702 x)  (activations.py QuickGELU) = 0.5 x (1 + tanh(0.851 x)): one MUFU (tanh.approx, the accuracy
        // class of the gelu_pytorch_tanh branch above) and three FP32 ops. The ex2 + rcp form needs two MUFUs per element,
        // and with K = 1024 the tile's MMAs last only 8192 cycles: ncu had CLIP's fc1 at 87.5 % tensor-active against 97 %
        // for SigLIP's fc1 with its single-MUFU epilogue.
#ifdef VLMFW_QUICK_GELU_EX2  // A/B baseline (tools/ variant builds only)
        return __fdividef(x, 1.0f + ex2_approx(x * (-1.702f * 1.4426950408889634f)));
#endif
        const float hx = 0.5f * x;
        return fmaf(hx, tanh_approx(0.851f * x), hx);
    } else if constexpr (ACT == VLMFW_ACT_GELU_ERF) {
        // 0.5 x (1 + erf(x / sqrt 2))  (activations.py GELUActivation, EVA / BLIP-2). erf from Abramowitz & Stegun 7.1.26,
        // erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z >= 0 (|error| <= 1.5e-7; the whole expression
        // stays within 5e-7 of the fp64 GELU, a tenth of the smallest fp16 step it is rounded to): two MUFUs (rcp, ex2) and
        // twelve FP32 ops without a branch. libdevice's erff cost ~27 instructions per element and made this epilogue, not
        // the tensor pipe, pace EVA's fc1 (ncu: 63 % tensor-active against 79-83 % for the other two towers' fc1).
#ifdef VLMFW_ERF_LIBDEVICE  // A/B baseline (tools/ variant builds only)
        return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#endif
        const float az = fabsf(x) * 0.70710678118654752f;
        const float t = rcp_approx(fmaf(az, 0.3275911f, 1.0f));
        float pl = fmaf(t, 1.061405429f, -1.453152027f);
        pl = fmaf(pl, t, 1.421413741f);
        pl = fmaf(pl, t, -0.284496736f);
        pl = fmaf(pl, t, 0.254829592f);
        pl *= t;
        const float e = ex2_approx(az * az * -1.4426950408889634f);
        const float erf_abs = fmaf(-pl, e, 1.0f);
        const float hx = 0.5f * x;
        return fmaf(fabsf(hx), erf_abs, hx);  // x >= 0: hx (1 + erf|z|); x < 0: hx (1 - erf|z|)
    } else {
        return x;
    }
}

template <int BN, int EPI, int ACT, int CG, bool GEN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_wt, const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_h16, const KParams p) {
    using C = Cfg<BN, CG, EPI, GEN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;  // 1024-byte aligned: every stage is a multiple of 1 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;  // [EPI_WARPS][RES_NBUF], residual epilogue only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS * RES_NBUF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // a "worker" is one CTA (CG = 1) or one CTA pair (CG = 2); it owns TILE_M x BN output tiles
    constexpr int TILE_M = BM * CG;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int worker = static_cast<int>(blockIdx.x) / CG;
    const int num_workers = static_cast<int>(gridDim.x) / CG;
    const int m_tiles = (p.m + TILE_M - 1) / TILE_M;
    const int n_tiles = (p.n + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.k + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_wt);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], EPI_WARPS * CG);  // pair: the leader's barrier collects both CTAs' epilogue warps
        }
        for (int s = 0; s < EPI_WARPS * RES_NBUF; ++s) mbar_init(&res_bar[s], 1);
        if constexpr (!GEN) {
            tma_prefetch_desc(&tmap_out);
            if constexpr (EPI != VLMFW_EPI_BIAS_BF16) tma_prefetch_desc(&tmap_res);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, C::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers must exist before anything is signalled on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Everything above (barriers, TMEM, descriptor prefetch) touched no global data and may overlap the tail of the
    // previous kernel in the stream; from here on its outputs are read and its inputs overwritten.
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // The whole warp runs the loop (uniform control flow, every lane waits on the ring barrier); one elected lane
        // issues. Tile / stage arithmetic and the tensor-map operands then live in uniform registers.
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM;
                // narrow last tile: each CTA of the pair stages tail_n / 2 rows of W through the narrower box
                const bool tail = CG == 2 && p.tail_n > 0 && (tile % n_tiles) == n_tiles - 1;
                const int b_rows = tail ? p.tail_n / 2 : C::B_ROWS;
                const int n0 = (tile % n_tiles) * BN + static_cast<int>(cta_rank) * b_rows;
                const uint32_t stage_tx = C::A_BYTES + static_cast<uint32_t>(b_rows) * BK * 2;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    if (elect_one()) {
                        if (VLMFW_DBG(p, 1) && (tile != worker || kb >= C::STAGES)) {
                            if (cta_rank == 0) mbar_arrive(&full_bar[stage]);
                        } else if constexpr (CG == 2) {
                            // both CTAs load into their own smem; all bytes are counted on the LEADER's full barrier
                            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
                            tma_load_2d_pair(sa, &tmap_a, lead_bar, kb * BK, m0);
                            tma_load_2d_pair(sb, tail ? &tmap_wt : &tmap_w, lead_bar, kb * BK, n0);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
                            tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (same form: uniform loop, elected issue)
        if (cta_rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const uint32_t idesc =
                    (CG == 2 && p.tail_n > 0 && (tile % n_tiles) == n_tiles - 1) ? p.idesc_tail : p.idesc;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + C::A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            if (VLMFW_DBG(p, 4)) break;
                            // advancing K inside the 128B swizzle row: +32 bytes -> +2 in the (addr >> 4) field
                            if constexpr (CG == 2)
                                umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(ks * 2),
                                                  bdesc + static_cast<uint64_t>(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
                            else
                                umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(ks * 2),
                                             bdesc + static_cast<uint64_t>(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
                        }
                        // ring slot reusable (in both CTAs of a pair) once these MMAs have read it
                        if constexpr (CG == 2) {
                            umma_commit_pair(&empty_bar[stage], 3);
                            if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc], 3);
                        } else {
                            umma_commit(&empty_bar[stage]);
                            if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int ew = warp - 2;
        const int quarter = warp & 3;        // TMEM lanes [32*quarter, +32) are the only ones this warp may read
        const int half = ew >> 2;            // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        constexpr int CHUNKS = COLS_PER_WARP / 32;
        uint8_t* stg = staging + ew * C::STG_PER_WARP;
        const uint32_t stg_u32 = smem_u32(stg);
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t lead_tmem_empty = (CG == 2) ? mapa_shared(smem_u32(tmem_empty), 0) : 0u;
        auto release_acc = [&]() {
            // accumulator drained -> hand the TMEM buffer back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(lead_tmem_empty + static_cast<uint32_t>(acc) * 8u);
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        };
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                    static_cast<uint32_t>(half * COLS_PER_WARP);

        if constexpr (!GEN && EPI == VLMFW_EPI_BIAS_BF16) {
            // ---- 16-bit output: act(acc + bias) -> fp16/bf16 -> [32 rows x 32 cols] 64B-swizzled box -> TMA store.
            // Two boxes per warp alternate, so a store may still be reading one while the next chunk is written.
            const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;  // SWIZZLE_64B: 16B chunk ^= (row / 2) % 4
            const uint32_t row_u32 = stg_u32 + static_cast<uint32_t>(lane) * 64u;
            uint32_t q = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                const int n0 = (tile % n_tiles) * BN + half * COLS_PER_WARP;
                // folded LayerNorm: this thread's row statistics from the producer's partial sums (fixed order)
                float mu = 0.f, rstd = 1.f;
                if (p.ln_stats != nullptr && m0 + lane < p.m) {
                    const float2* sp = p.ln_stats + static_cast<size_t>(m0 + lane) * p.ln_row_mul * p.ln_np;
                    float s1 = 0.f, s2 = 0.f;
                    for (int i = 0; i < p.ln_np; ++i) {
                        const float2 t = __ldg(sp + i);
                        s1 += t.x;
                        s2 += t.y;
                    }
                    mu = s1 * p.ln_inv_dim;
                    rstd = rsqrtf(fmaxf(s2 * p.ln_inv_dim - mu * mu, 0.f) + p.ln_eps);
                }
                rstd *= p.acc_scale;
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(acc * BN), v);
#pragma unroll 1
                for (int c = 0; c < CHUNKS; ++c, ++q) {
                    const int nc = n0 + c * 32;
                    float4 bq[8];  // bias (an L2 round trip each: no L1 left beside 227 KB of shared memory) under the TMEM wait
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr && nc + 4 * j < p.n) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 4 * j));
                    }
                    tmem_ld_wait();
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bv = bq[j];
                        y[4 * j + 0] = fmaf(rstd, __uint_as_float(v[4 * j + 0]), bv.x);
                        y[4 * j + 1] = fmaf(rstd, __uint_as_float(v[4 * j + 1]), bv.y);
                        y[4 * j + 2] = fmaf(rstd, __uint_as_float(v[4 * j + 2]), bv.z);
                        y[4 * j + 3] = fmaf(rstd, __uint_as_float(v[4 * j + 3]), bv.w);
                    }
                    // the next chunk's TMEM read flies while this one is activated, packed and stored
                    if (c + 1 < CHUNKS) tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(acc * BN + (c + 1) * 32), v);
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 32; ++j) y[j] = activation<ACT>(y[j]);
                    if (p.out_f16) {
                        if (p.sat != nullptr) {  // warp-uniform
                            float amax = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; j += 2) amax = fmaxf(amax, fmaxf(fabsf(y[j]), fabsf(y[j + 1])));
                            if (amax > 65504.f) atomicAdd(p.sat, 1ull);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = pack_f16x2(y[2 * j], y[2 * j + 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(y[2 * j], y[2 * j + 1]);
                    }
                    if (VLMFW_DBG(p, 2)) continue;
                    const uint32_t box = row_u32 + (q & 1u) * 2048u;
                    if (lane == 0) tma_store_wait_read<1>();  // the store issued two chunks ago has left this box
                    __syncwarp();
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + ((j ^ sw) << 4)), "r"(o[4 * j]),
                                     "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                                     : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, stg_u32 + (q & 1u) * 2048u, nc, m0);
                        tma_store_commit();
                    }
                }
                release_acc();
            }
            if (lane == 0) tma_store_wait_all();
        } else if constexpr (!GEN && EPI == VLMFW_EPI_BIAS_RESID_16) {
            // ---- 16-bit residual stream, updated in place: out = round16(acc + bias + resid). ONE box type, [32 rows x 32
            // cols] of 16-bit (64-byte rows, 64B-swizzled), ring of RES_NBUF per warp: TMA load, every thread updates its
            // row in place, TMA store of the same box. The stored values are at once the residual stream and the next
            // GEMM's A operand, and the row statistics are taken from the ROUNDED values, so the LayerNorm folded into
            // that GEMM normalises exactly what it multiplies. Half the HBM and shared-memory bytes of the fp32 stream.
            constexpr int RC = 32;
            constexpr int RCHUNKS = COLS_PER_WARP / RC;
            constexpr int BOX = H16_BOX_BYTES;
            const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;  // SWIZZLE_64B
            uint64_t* my_bar = res_bar + ew * RES_NBUF;
            const int my_tiles = worker < num_tiles ? (num_tiles - worker + num_workers - 1) / num_workers : 0;
            const uint32_t total_q = static_cast<uint32_t>(my_tiles) * RCHUNKS;
            auto issue_load = [&](uint32_t qq) {  // lane 0 only
                const int t = worker + static_cast<int>(qq / RCHUNKS) * num_workers;
                const int c = static_cast<int>(qq % RCHUNKS);
                const int cx = (t % n_tiles) * BN + half * COLS_PER_WARP + c * RC;
                const int cy = (t / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                const uint32_t b = qq % RES_NBUF;
                mbar_arrive_expect_tx(&my_bar[b], BOX);
                tma_load_2d(stg + b * BOX, &tmap_res, &my_bar[b], cx, cy);
            };
            if (lane == 0 && !VLMFW_DBG(p, 2)) {
                for (uint32_t qq = 0; qq < RES_NBUF - 1 && qq < total_q; ++qq) issue_load(qq);
            }
            auto unpack2 = [&](uint32_t u, float& lo, float& hi) {
                if (p.out_f16) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
                    lo = f.x;
                    hi = f.y;
                } else {
                    lo = __uint_as_float(u << 16);
                    hi = __uint_as_float(u & 0xFFFF0000u);
                }
            };
            uint32_t q = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                const int n0 = (tile % n_tiles) * BN + half * COLS_PER_WARP;
                float s1 = 0.f, s2 = 0.f;
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < RCHUNKS; ++c, ++q) {
                    if (VLMFW_DBG(p, 2)) continue;
                    const int nc = n0 + c * RC;
                    // The chunk's 32 bias values first: with the whole L1 carved out as shared memory every __ldg is an L2
                    // round trip, and inside the update loop below each of the four loads sat behind the previous group's
                    // st.shared (ncu source view of out-proj: 28 % of all samples on the four FFMAs that consume them —
                    // 3700 cycles per chunk, so that the K = 1152 GEMM was bound by its epilogue). Issued here they fly
                    // under the TMEM load and the barrier waits.
                    float4 bq[RC / 4];
#pragma unroll
                    for (int j = 0; j < RC / 4; ++j) {
                        bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr && nc + 4 * j < p.n) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 4 * j));
                    }
                    uint32_t v[RC];
                    tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(acc * BN + c * RC), v);
                    const uint32_t b = q % RES_NBUF;
                    const uint32_t box = stg_u32 + b * BOX + static_cast<uint32_t>(lane) * 64u;
                    mbar_wait(&my_bar[b], (q / RES_NBUF) & 1u);
                    tmem_ld_wait();
                    // warp-uniform: columns a narrow last tile never wrote (its MMA is as wide as N rounded up to 16) hold
                    // stale TMEM and must not reach the statistics
                    const int n16 = (p.n + 15) & ~15;
                    float amax = 0.f;  // largest |value| this thread stores in this chunk (fp16 saturation report)
#pragma unroll
                    for (int g16 = 0; g16 < RC / 16; ++g16) {
                        if (nc + 16 * g16 >= n16) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[16 * g16 + j] = 0u;
                        }
                    }
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {  // 16 bytes = 8 columns
                        const float4 b0 = bq[2 * j], b1 = bq[2 * j + 1];
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const uint32_t a = box + ((j ^ sw) << 4);
                        uint32_t u[4];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
                                     : "r"(a));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float lo, hi;
                            unpack2(u[i], lo, hi);
                            lo += fmaf(__uint_as_float(v[8 * j + 2 * i]), p.acc_scale, bv[2 * i]);
                            hi += fmaf(__uint_as_float(v[8 * j + 2 * i + 1]), p.acc_scale, bv[2 * i + 1]);
                            amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
                            u[i] = p.out_f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
                            unpack2(u[i], lo, hi);  // statistics of what is stored
                            s1 += lo + hi;
                            s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(u[0]), "r"(u[1]), "r"(u[2]),
                                     "r"(u[3])
                                     : "memory");
                    }
                    if (p.out_f16 && p.sat != nullptr && amax > 65504.f) atomicAdd(p.sat, 1ull);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, stg_u32 + b * BOX, nc, m0);
                        tma_store_commit();
                        // every bulk group but the newest has read its box: the box of chunk q - 1 takes the next load
                        tma_store_wait_read<1>();
                        if (q + RES_NBUF - 1 < total_q) issue_load(q + RES_NBUF - 1);
                    }
                    __syncwarp();
                }
                if (p.stats_out != nullptr && m0 + lane < p.m && !VLMFW_DBG(p, 2)) {
                    p.stats_out[static_cast<size_t>(m0 + lane) * p.stats_np + (tile % n_tiles) * 2 + half] =
                        make_float2(s1, s2);
                }
                release_acc();
            }
            if (lane == 0) tma_store_wait_all();
        } else if constexpr (!GEN && EPI == VLMFW_EPI_BIAS_RESID_F32) {
            // ---- fp32 residual: out = acc + bias + resid, usually in place on the residual stream. The residual chunk
            // [32 rows x 16 fp32] arrives by TMA in a 64B-swizzled box (ring of RES_NBUF per warp, the next chunk's
            // load issued as soon as its box is free, across tile boundaries), every thread updates its own row in
            // place, and the same box is TMA-stored. Fused for the LayerNorm that follows in the tower: a 16-bit copy
            // of the updated rows (TMA-stored from a second box every two chunks) and the row's (sum, sum of squares)
            // over this warp's columns, so the next GEMM normalises in its epilogue and no LayerNorm kernel runs.
            constexpr int RCHUNKS = COLS_PER_WARP / RES_COLS;
            static_assert(RCHUNKS % 2 == 0, "the 16-bit box spans two residual chunks");
            const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;  // SWIZZLE_64B
            const uint32_t h16_u32 = stg_u32 + RES_NBUF * RES_BOX_BYTES;
            uint64_t* my_bar = res_bar + ew * RES_NBUF;
            const int my_tiles = worker < num_tiles ? (num_tiles - worker + num_workers - 1) / num_workers : 0;
            const uint32_t total_q = static_cast<uint32_t>(my_tiles) * RCHUNKS;
            auto issue_load = [&](uint32_t qq) {  // lane 0 only
                const int t = worker + static_cast<int>(qq / RCHUNKS) * num_workers;
                const int c = static_cast<int>(qq % RCHUNKS);
                const int cx = (t % n_tiles) * BN + half * COLS_PER_WARP + c * RES_COLS;
                int cy = (t / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                if (p.resid_mod > 0) cy %= p.resid_mod;  // broadcast residual (position embedding): resid_mod % 32 == 0
                const uint32_t b = qq % RES_NBUF;
                mbar_arrive_expect_tx(&my_bar[b], RES_BOX_BYTES);
                tma_load_2d(stg + b * RES_BOX_BYTES, &tmap_res, &my_bar[b], cx, cy);
            };
            if (lane == 0 && !VLMFW_DBG(p, 2)) {
                for (uint32_t qq = 0; qq < RES_NBUF - 1 && qq < total_q; ++qq) issue_load(qq);
            }
            static_assert(RES_NBUF >= 3, "ring: chunk q's store may still read its box while q + 1 is processed");
            uint32_t q = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                const int n0 = (tile % n_tiles) * BN + half * COLS_PER_WARP;
                float s1 = 0.f, s2 = 0.f;
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < RCHUNKS; ++c, ++q) {
                    if (VLMFW_DBG(p, 2)) continue;
                    const int nc = n0 + c * RES_COLS;
                    float4 bq[RES_COLS / 4];  // bias first (see the 16-bit residual epilogue)
#pragma unroll
                    for (int j = 0; j < RES_COLS / 4; ++j) {
                        bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr && nc + 4 * j < p.n) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 4 * j));
                    }
                    uint32_t v[RES_COLS];
                    tmem_ld_row(lane_taddr + static_cast<uint32_t>(acc * BN + c * RES_COLS), v);
                    const uint32_t b = q % RES_NBUF;
                    const uint32_t box = stg_u32 + b * RES_BOX_BYTES + static_cast<uint32_t>(lane) * (RES_COLS * 4);
                    mbar_wait(&my_bar[b], (q / RES_NBUF) & 1u);
                    tmem_ld_wait();
                    if (nc >= p.n) {  // warp-uniform: columns a narrow last tile never wrote (stale TMEM) must not reach the row statistics
#pragma unroll
                        for (int j = 0; j < RES_COLS; ++j) v[j] = 0u;
                    }
                    uint32_t hp[RES_COLS / 2];
#pragma unroll
                    for (uint32_t j = 0; j < RES_COLS / 4; ++j) {
                        const float4 bv = bq[j];
                        const uint32_t a = box + ((j ^ sw) << 4);
                        float4 r;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                                     : "r"(a));
                        r.x += fmaf(__uint_as_float(v[4 * j + 0]), p.acc_scale, bv.x);
                        r.y += fmaf(__uint_as_float(v[4 * j + 1]), p.acc_scale, bv.y);
                        r.z += fmaf(__uint_as_float(v[4 * j + 2]), p.acc_scale, bv.z);
                        r.w += fmaf(__uint_as_float(v[4 * j + 3]), p.acc_scale, bv.w);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(r.x), "f"(r.y), "f"(r.z),
                                     "f"(r.w)
                                     : "memory");
                        s1 += (r.x + r.y) + (r.z + r.w);
                        s2 = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, s2))));
                        hp[2 * j] = p.out_f16 ? pack_f16x2(r.x, r.y) : pack_bf16x2(r.x, r.y);
                        hp[2 * j + 1] = p.out_f16 ? pack_f16x2(r.z, r.w) : pack_bf16x2(r.z, r.w);
                    }
                    if (p.h16_on) {
                        // this chunk is 32 bytes of the 64-byte row of the 16-bit box: 16B slots 2 (c & 1) + {0, 1}
                        const uint32_t hrow = h16_u32 + static_cast<uint32_t>(lane) * 64u;
                        const uint32_t slot = static_cast<uint32_t>(c & 1) * 2u;
#pragma unroll
                        for (uint32_t j = 0; j < 2; ++j) {
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + (((slot + j) ^ sw) << 4)),
                                         "r"(hp[4 * j]), "r"(hp[4 * j + 1]), "r"(hp[4 * j + 2]), "r"(hp[4 * j + 3])
                                         : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.h16_on && (c & 1)) {  // its own, OLDER bulk group: the wait below then covers it
                            tma_store_2d(&tmap_h16, h16_u32, nc - RES_COLS, m0);
                            tma_store_commit();
                        }
                        tma_store_2d(&tmap_out, stg_u32 + b * RES_BOX_BYTES, nc, m0);
                        tma_store_commit();
                        // every bulk group but the newest (this chunk's fp32 store) has read its box: the box chunk
                        // q - 1 was stored from (= box (q + NBUF - 1) % NBUF) takes the load issued NBUF - 1 chunks
                        // ahead, and the 16-bit box may be rewritten by the next chunk
                        tma_store_wait_read<1>();
                        if (q + RES_NBUF - 1 < total_q) issue_load(q + RES_NBUF - 1);
                    }
                    __syncwarp();
                }
                if (p.stats_out != nullptr && m0 + lane < p.m && !VLMFW_DBG(p, 2)) {
                    p.stats_out[static_cast<size_t>(m0 + lane) * p.stats_np + (tile % n_tiles) * 2 + half] =
                        make_float2(s1, s2);
                }
                release_acc();
            }
            if (lane == 0) tma_store_wait_all();
        } else {
        const int rsub = lane >> 3;  // read-back: 4 rows per pass
        const int cj = lane & 7;     // read-back: 16-byte column group (4 fp32)

        // Residual epilogue: the fp32 residual tile is the only global READ of the epilogue and it does not depend on
        // the accumulator, so it is software-pipelined one 32-column chunk ahead (across tile boundaries too): the
        // loads of chunk i+1 are in flight while chunk i waits for TMEM / transposes / stores.
        float4 rnext[8];
        auto resid_issue = [&](int tile, int c, float4 (&r)[8]) {
            if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
                if (VLMFW_DBG(p, 2)) return;
                const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
                const int col = (tile % n_tiles) * BN + half * COLS_PER_WARP + c * 32 + cj * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = m0 + it * 4 + rsub;
                    r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < p.m && col < p.n) {
                        const int rrow = (p.resid_mod > 0) ? (row % p.resid_mod) : row;
                        r[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(rrow) * p.ldr + col);
                    }
                }
            }
        };
        if (worker < num_tiles) resid_issue(worker, 0, rnext);

        for (int tile = worker; tile < num_tiles; tile += num_workers) {
            const int m0 = (tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
            const int n0 = (tile % n_tiles) * BN + half * COLS_PER_WARP;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                const int nc = n0 + c * 32;
                float4 rcur[8];
                if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) rcur[it] = rnext[it];
                    // next chunk of this tile, or chunk 0 of this CTA's next tile
                    if (c + 1 < CHUNKS) resid_issue(tile, c + 1, rnext);
                    else if (tile + num_workers < num_tiles) resid_issue(tile + num_workers, 0, rnext);
                }
                if (nc >= p.n || VLMFW_DBG(p, 2)) continue;  // warp-uniform: whole chunk is past the N edge (prefetch above still ran)
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                       static_cast<uint32_t>(acc * BN + half * COLS_PER_WARP + c * 32);
                tmem_ld_32x32(taddr, v);
                tmem_ld_wait();
                // transpose through smem: thread `lane` owns tile row `lane`; 16B groups XOR-swizzled by row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t dst = stg_u32 + lane * 128 + ((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                                 "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                 : "memory");
                }
                __syncwarp();
                const int col = nc + cj * 4;
                const bool col_ok = col < p.n;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr && col_ok) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rsub;
                    const int row = m0 + rr;
                    float4 a;
                    const uint32_t src = stg_u32 + rr * 128 + ((cj ^ (rr & 7)) << 4);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                                 : "r"(src));
                    if (row < p.m && col_ok) {
                        a.x = fmaf(a.x, p.acc_scale, bv.x);
                        a.y = fmaf(a.y, p.acc_scale, bv.y);
                        a.z = fmaf(a.z, p.acc_scale, bv.z);
                        a.w = fmaf(a.w, p.acc_scale, bv.w);
                        int orow = row;
                        if (p.grp > 0) orow = (row / p.grp) * p.grp_out + (row % p.grp) + p.grp_off;
                        if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
                            a.x += rcur[it].x;
                            a.y += rcur[it].y;
                            a.z += rcur[it].z;
                            a.w += rcur[it].w;
                            *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(orow) * p.ldo +
                                                       col) = a;
                        } else {
                            uint2 o;
                            const float y0 = activation<ACT>(a.x), y1 = activation<ACT>(a.y);
                            const float y2 = activation<ACT>(a.z), y3 = activation<ACT>(a.w);
                            o.x = p.out_f16 ? pack_f16x2(y0, y1) : pack_bf16x2(y0, y1);
                            o.y = p.out_f16 ? pack_f16x2(y2, y3) : pack_bf16x2(y2, y3);
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) +
                                                      static_cast<size_t>(orow) * p.ldo + col) = o;
                        }
                    }
                }
                __syncwarp();
            }
            release_acc();
        }
        }  // generic epilogue
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its smem
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
        else tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

struct MapKey {
    const void* ptr;
    int rows, cols, ld, box_rows, box_cols, dt, swz;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols && dt == o.dt && swz == o.swz;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
        h ^= (static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
        h ^= (static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull) + (h << 6) + (h >> 2);
        h ^= static_cast<size_t>(((k.box_rows * 512 + k.box_cols) * 4 + k.dt) * 256 + k.swz) + (h << 6) + (h >> 2);
        return h;
    }
};

enum MapDt { MAP_BF16 = 0, MAP_F16 = 1, MAP_F32 = 2 };

// [rows, cols] row-major matrix (row stride ld elements) -> tiled map with box {box_cols, box_rows} and the given
// swizzle; out-of-bounds elements read as zero and are not written (this is what makes ragged M / N / K edges free).
int get_tensor_map(const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols, MapDt dt, int swizzle_bytes,
                   CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows, box_cols, static_cast<int>(dt), swizzle_bytes};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return VLMFW_OK;
        }
    }
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return VLMFW_E_CUDA;
    }
    const int esize = dt == MAP_F32 ? 4 : 2;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esize};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estride[2] = {1, 1};
    const CUtensorMapDataType cdt = dt == MAP_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                    : dt == MAP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tm;
    CUresult r = fn(&tm, cdt, 2, const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) ptr=%p rows=%d cols=%d ld=%d box=%dx%d dt=%d",
                  static_cast<int>(r), ptr, rows, cols, ld, box_rows, box_cols, static_cast<int>(dt));
        return VLMFW_E_CUDA;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 4096) cache.clear();
        cache.emplace(key, tm);
    }
    *out = tm;
    return VLMFW_OK;
}

template <int BN, int EPI, int ACT, int CG, bool GEN>
int launch_one(const GemmArgs& g, cudaStream_t stream) {
    using C = Cfg<BN, CG, EPI, GEN>;
    const MapDt dt16 = g.f16 ? MAP_F16 : MAP_BF16;
    CUtensorMap ta, tw, twt, tout, tres, th16;
    int rc = get_tensor_map(g.a, g.m, g.k, g.lda, BM, BK, dt16, 128, &ta);
    if (rc != VLMFW_OK) return rc;
    rc = get_tensor_map(g.w, g.n, g.k, g.ldw, C::B_ROWS, BK, dt16, 128, &tw);
    if (rc != VLMFW_OK) return rc;
    twt = tw;
    int tail_n = 0;
    if constexpr (CG == 2 && !GEN) {
        static const bool tail_on = [] { const char* e = getenv("VLMFW_GEMM_TAIL"); return e == nullptr || atoi(e) != 0; }();
        const int rem = g.n - ((g.n + BN - 1) / BN - 1) * BN;  // real width of the last N tile
        const int want = (rem + 15) / 16 * 16;
        if (tail_on && want < BN) {
            tail_n = want;
            rc = get_tensor_map(g.w, g.n, g.k, g.ldw, tail_n / 2, BK, dt16, 128, &twt);
            if (rc != VLMFW_OK) return rc;
        }
    }
    tout = ta;  // unused by the generic epilogue, but a kernel parameter must be a valid object
    tres = ta;
    th16 = ta;
    if constexpr (!GEN) {
        if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
            rc = get_tensor_map(g.out, g.m, g.n, g.ldo, 32, RES_COLS, MAP_F32, RES_COLS * 4, &tout);
            if (rc != VLMFW_OK) return rc;
            rc = get_tensor_map(g.resid, g.resid_mod > 0 ? g.resid_mod : g.m, g.n, g.ldr, 32, RES_COLS, MAP_F32, RES_COLS * 4, &tres);
            if (rc != VLMFW_OK) return rc;
            if (g.h16_out != nullptr) {
                rc = get_tensor_map(g.h16_out, g.m, g.n, g.ld_h16, 32, 32, dt16, 64, &th16);
                if (rc != VLMFW_OK) return rc;
            }
        } else {
            rc = get_tensor_map(g.out, g.m, g.n, g.ldo, 32, 32, dt16, 64, &tout);
            if (rc != VLMFW_OK) return rc;
            if constexpr (EPI == VLMFW_EPI_BIAS_RESID_16) {
                rc = get_tensor_map(g.resid, g.m, g.n, g.ldr, 32, 32, dt16, 64, &tres);
                if (rc != VLMFW_OK) return rc;
            }
        }
    }
    auto kern = gemm_bf16_tn_kernel<BN, EPI, ACT, CG, GEN>;
    static std::atomic<uint64_t> smem_opt_in{0};  // per instantiation, one bit per device
    VLMFW_CUDA_OK(ensure_dynamic_smem(kern, C::SMEM_BYTES, smem_opt_in));
    KParams p;
    p.m = g.m;
    p.n = g.n;
    p.k = g.k;
    p.idesc = umma_idesc_16(BM * CG, BN, g.f16, g.f16);
    p.tail_n = tail_n;
    p.idesc_tail = tail_n > 0 ? umma_idesc_16(BM * CG, tail_n, g.f16, g.f16) : p.idesc;
    p.out_f16 = g.f16 ? 1 : 0;
    p.bias = g.bias;
    p.out = g.out;
    p.ldo = g.ldo;
    p.resid = g.resid;
    p.ldr = g.ldr;
    p.resid_mod = g.resid_mod;
    p.grp = g.grp;
    p.grp_out = g.grp_out;
    p.grp_off = g.grp_off;
    p.ln_stats = reinterpret_cast<const float2*>(g.ln_stats);
    p.ln_np = g.ln_np;
    p.ln_row_mul = g.ln_row_mul > 0 ? g.ln_row_mul : 1;
    p.ln_inv_dim = g.ln_dim > 0 ? 1.0f / static_cast<float>(g.ln_dim) : 0.f;
    p.ln_eps = g.ln_eps;
    p.stats_out = reinterpret_cast<float2*>(g.stats_out);
    p.stats_np = 2 * ((g.n + BN - 1) / BN);
    p.h16_on = g.h16_out != nullptr ? 1 : 0;
    p.acc_scale = g.acc_scale > 0.f ? g.acc_scale : 1.f;
    p.sat = g.sat_count;
#ifdef VLMFW_DEBUG_MASKS
    static const int dbg = []() {
        const char* e = getenv("VLMFW_GEMM_DBG");
        return e ? atoi(e) : 0;
    }();
    p.dbg = dbg;
#else
    p.dbg = 0;
#endif
    constexpr int TILE_M = BM * CG;
    const int m_tiles = (g.m + TILE_M - 1) / TILE_M;
    const int n_tiles = (g.n + BN - 1) / BN;
    const int tiles = m_tiles * n_tiles;
    const int workers = num_sms() / CG;
    const int grid = (tiles < workers ? tiles : workers) * CG;
    if constexpr (CG == 2) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        VLMFW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tw, twt, tout, tres, th16, p));
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        VLMFW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tw, twt, tout, tres, th16, p));
    }
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

template <int BN, int CG>
int dispatch_epi(const GemmArgs& g, cudaStream_t stream) {
    // row remap (patch embedding around the class-token slot) and a broadcast residual (position embedding) go through
    // the generic epilogue; everything else moves its output (and residual) with TMA
    const bool generic = g.grp > 0 || (g.resid_mod > 0 && (g.resid_mod % 32 != 0 || g.epi != VLMFW_EPI_BIAS_RESID_F32)) ||
                         (g.block_n & VLMFW_GEMM_FORCE_GENERIC_EPI);
    if (g.epi == VLMFW_EPI_BIAS_RESID_16) {
        VLMFW_REQUIRE(!generic, "gemm: the 16-bit residual epilogue has no row remap / broadcast residual");
        return launch_one<BN, VLMFW_EPI_BIAS_RESID_16, VLMFW_ACT_NONE, CG, false>(g, stream);
    }
    if (g.epi == VLMFW_EPI_BIAS_RESID_F32) {
        if (generic) return launch_one<BN, VLMFW_EPI_BIAS_RESID_F32, VLMFW_ACT_NONE, CG, true>(g, stream);
        return launch_one<BN, VLMFW_EPI_BIAS_RESID_F32, VLMFW_ACT_NONE, CG, false>(g, stream);
    }
    if (generic) {
        VLMFW_REQUIRE(g.act == VLMFW_ACT_NONE, "gemm: the row-remapped 16-bit epilogue has no fused activation");
        return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_NONE, CG, true>(g, stream);
    }
    switch (g.act) {
        case VLMFW_ACT_GELU_TANH:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_GELU_TANH, CG, false>(g, stream);
        case VLMFW_ACT_QUICK_GELU:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_QUICK_GELU, CG, false>(g, stream);
        case VLMFW_ACT_GELU_ERF:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_GELU_ERF, CG, false>(g, stream);
        default:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_NONE, CG, false>(g, stream);
    }
}

template <int CG>
int dispatch_bn(int bn, const GemmArgs& g, cudaStream_t stream) {
    switch (bn) {
        case 256:
            return dispatch_epi<256, CG>(g, stream);
        case 192:
            return dispatch_epi<192, CG>(g, stream);
        case 128:
            return dispatch_epi<128, CG>(g, stream);
        default:
            set_error("gemm: unsupported block_n %d", bn);
            return VLMFW_E_INVALID;
    }
}

// Pick the N tile by a measured cost model: MMA-only throughput on B200 at M = 256 (CTA pair) is 1.00 / 0.85 / 0.59 of
// the BN = 256 rate for BN = 256 / 192 / 128 (the A-operand read does not shrink with N), so a tile costs
// 1.00 / 0.88 / 0.85 time units; 256 wins whenever N > 192 even with a ragged last tile (TMA zero-fills it for free).
int pick_block_n(int n) {
    const int cands[3] = {256, 192, 128};
    const float unit[3] = {1.00f, 0.88f, 0.85f};
    int best = 256;
    float best_cost = -1.f;
    for (int i = 0; i < 3; ++i) {
        const float cost = static_cast<float>((n + cands[i] - 1) / cands[i]) * unit[i];
        if (best_cost < 0.f || cost < best_cost) {
            best_cost = cost;
            best = cands[i];
        }
    }
    return best;
}

}  // namespace

int encode_tensor_map(CUtensorMap* out, bool f16, int rank, const void* ptr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return VLMFW_E_CUDA;
    }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estride[i] = 1;
        if (i > 0) gstride[i - 1] = strides_bytes[i - 1];
    }
    const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                    const_cast<void*>(ptr), gdim, gstride, bx, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(rank %d) failed (CUresult %d)", rank, static_cast<int>(r));
        return VLMFW_E_CUDA;
    }
    return VLMFW_OK;
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("VLMFW_PDL"); return e == nullptr || atoi(e) != 0; }();
    return on;
}

int gemm_stats_parts(int n, int block_n) {
    const int bn_req = block_n & 0xFFF;
    const int bn = bn_req > 0 ? bn_req : pick_block_n(n);
    return 2 * ((n + bn - 1) / bn);
}

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
    VLMFW_REQUIRE(g.a && g.w && g.out, "gemm: null operand");
    VLMFW_REQUIRE(g.m > 0 && g.n > 0 && g.k > 0, "gemm: bad shape m=%d n=%d k=%d", g.m, g.n, g.k);
    VLMFW_REQUIRE(g.k % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm: k/lda/ldw must be multiples of 8 (16-byte TMA rows)");
    VLMFW_REQUIRE(g.n % 8 == 0 && g.ldo % 8 == 0, "gemm: n/ldo must be multiples of 8");
    VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
                  "gemm: operands must be 16-byte aligned");
    VLMFW_REQUIRE(g.lda >= g.k && g.ldw >= g.k && g.ldo >= g.n, "gemm: leading dimension smaller than row");
    if (g.epi == VLMFW_EPI_BIAS_RESID_F32) {
        VLMFW_REQUIRE(g.resid != nullptr && g.ldr % 4 == 0 && g.ldr >= g.n, "gemm: residual epilogue needs resid/ldr");
        VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.resid) & 15) == 0, "gemm: resid must be 16-byte aligned");
    } else if (g.epi == VLMFW_EPI_BIAS_RESID_16) {
        VLMFW_REQUIRE(g.resid != nullptr && g.ldr % 8 == 0 && g.ldr >= g.n, "gemm: 16-bit residual epilogue needs resid/ldr");
        VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.resid) & 15) == 0, "gemm: resid must be 16-byte aligned");
        VLMFW_REQUIRE(g.h16_out == nullptr && g.act == VLMFW_ACT_NONE, "gemm: 16-bit residual epilogue has no h16_out / activation");
    } else {
        VLMFW_REQUIRE(g.epi == VLMFW_EPI_BIAS_BF16, "gemm: unknown epilogue %d", g.epi);
    }
    if (g.bias) VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
    const bool generic_epi = g.grp > 0 || (g.resid_mod > 0 && (g.resid_mod % 32 != 0 || g.epi != VLMFW_EPI_BIAS_RESID_F32)) ||
                             (g.block_n & VLMFW_GEMM_FORCE_GENERIC_EPI);
    if (g.ln_stats != nullptr) {
        VLMFW_REQUIRE(g.epi == VLMFW_EPI_BIAS_BF16 && !generic_epi, "gemm: folded LayerNorm needs the 16-bit TMA epilogue");
        VLMFW_REQUIRE(g.ln_np > 0 && g.ln_dim > 0, "gemm: folded LayerNorm needs ln_np / ln_dim");
        VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.ln_stats) & 7) == 0, "gemm: ln_stats must be 8-byte aligned");
    }
    if (g.stats_out != nullptr || g.h16_out != nullptr) {
        VLMFW_REQUIRE((g.epi == VLMFW_EPI_BIAS_RESID_F32 || g.epi == VLMFW_EPI_BIAS_RESID_16) && !generic_epi,
                      "gemm: row statistics / 16-bit copy are produced by the TMA residual epilogues only");
        if (g.h16_out != nullptr)
            VLMFW_REQUIRE(g.ld_h16 >= g.n && g.ld_h16 % 8 == 0 && (reinterpret_cast<uintptr_t>(g.h16_out) & 15) == 0,
                          "gemm: h16_out needs ld_h16 >= n, a multiple of 8, and 16-byte alignment");
    }
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int bn_req = g.block_n & 0xFFF;
    const int bn = bn_req > 0 ? bn_req : pick_block_n(g.n);
    // CTA pairs (256-row tiles) unless the caller forces a form or the problem has a single 128-row tile
    int cg = g.m > BM ? 2 : 1;
    if (g.block_n & VLMFW_GEMM_FORCE_1CTA) cg = 1;
    if (g.block_n & VLMFW_GEMM_FORCE_2CTA) cg = 2;
    return cg == 2 ? dispatch_bn<2>(bn, g, stream) : dispatch_bn<1>(bn, g, stream);
}

}  // namespace vlmfw
