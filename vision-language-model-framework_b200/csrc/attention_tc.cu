// Fused multi-head attention on the 5th-gen tensor cores: softmax(Q K^T / sqrt(dh)) V, no mask, eval mode.
// Restates hf integrations/sdpa_attention.py:92 as called from modeling_siglip.py:285-309, modeling_clip.py:310-334 and
// modeling_blip_2.py:328-351 — same contract and layouts as attention.cu (qkv [B*S, 3D] in, ctx [B*S, D] out).
//
// Persistent, warp-specialised, one CTA per SM (384 threads); a work item = (image, head, PAIR of 128-query blocks):
//   warp 0       TMA loader  — 5-D tensor map over qkv (head_dim, head, q|k|v, token, image): out-of-range head_dim
//                              columns and tokens past S arrive as zeros, so head_dim 72 / 88 need no padding in HBM
//   warp 1       MMA issuer  — S_qb = Q_qb K_j^T (tcgen05.mma 128x128xdh, accumulator in TMEM), O_qb += P_qb V_j
//                              (128 x dh x 128, V read MN-major straight from the token-major tile, no transpose)
//   warps 4..7   softmax of query block 0, warps 8..11 of query block 1 (one thread = one query row = one TMEM lane):
//                tcgen05.ld the 128 scores, row max / exp2 / row sum in registers with no shuffles, P -> fp16/bf16 ->
//                128B-swizzled smem as the A operand of the P·V MMA. The two blocks ping-pong so the tensor pipe works
//                on one while the SFUs exponentiate the other; both share every K/V tile.
// O accumulates in TMEM across key blocks. The running max is only raised when it grows by more than 2^8 (exact: the
// same reference max divides numerator and denominator); then O is rescaled in TMEM by the softmax warps.
// head_dim = 64 + tail (0 / 8->16 / 24->32 columns): the tail lives in its own narrow 32B / 64B-swizzled tile and is
// one extra K-step of Q·K^T and one extra narrow-N MMA of P·V.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int TC_THREADS = 384;  // 3 warpgroups: {loader, MMA, 2 idle} + softmax(q block 0) + softmax(q block 1)
constexpr int BLK = 128;  // query rows per block (= TMEM lanes) and keys per block

template <int DH>
struct TcCfg {
    static constexpr int TAIL = DH == 64 ? 0 : (DH == 72 ? 16 : 32);  // columns of the second head_dim group (zero padded)
    static constexpr int DHP = 64 + TAIL;
    static constexpr int TAIL_ROW = TAIL * 2;                 // bytes per row of the tail tile (= its swizzle span)
    static constexpr uint32_t TAIL_LAYOUT = TAIL == 16 ? 6u : 4u;  // SWIZZLE_32B : SWIZZLE_64B
    static constexpr int MAIN_BYTES = BLK * 128;
    static constexpr int TAIL_BYTES = BLK * TAIL_ROW;
    static constexpr int TILE_BYTES = MAIN_BYTES + TAIL_BYTES;
    static constexpr int P_BYTES = BLK * BLK * 2;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = 6 * TILE_BYTES + 2 * P_BYTES + BAR_BYTES + 1024;
};

struct TcParams {
    void* out;
    int S, H, B, D;
    int nqp;        // query-block pairs per (image, head)
    int J;          // key blocks per (image, head)
    int num_items;  // B * H * nqp
    float scale_log2;
    int f16;
    int dbg;  // VLMFW_ATTN_DBG bit mask (bottleneck experiments only): 1 skip exp+P store, 2 skip P·V MMAs, 4 skip S MMAs, 8 skip max
};

// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512)
constexpr uint32_t TM_S = 0, TM_O = 256, TM_STRIDE = 128;

template <int DH>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_main, const __grid_constant__ CUtensorMap map_tail,
                    const TcParams p) {
    using C = TcCfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // [2][TILE]
    uint8_t* sK = sQ + 2 * C::TILE_BYTES;     // [2][TILE]
    uint8_t* sV = sK + 2 * C::TILE_BYTES;     // [2][TILE]
    uint8_t* sP = sV + 2 * C::TILE_BYTES;     // [2][P]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * C::P_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [2]
    uint64_t* k_empty = bars + 4;   // [2]
    uint64_t* v_full = bars + 6;    // [2]
    uint64_t* v_empty = bars + 8;   // [2]
    uint64_t* s_full = bars + 10;   // [2] per query block
    uint64_t* s_empty = bars + 12;  // [2]
    uint64_t* p_full = bars + 14;   // [2]
    uint64_t* pv_done = bars + 16;  // [2]
    uint64_t* o_empty = bars + 18;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_main);
        if (C::TAIL) tma_prefetch_desc(&map_tail);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
            mbar_init(&o_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int J = p.J;

    // Register rebalancing (setmaxnreg works per warpgroup, inside the 384 x 168 registers the CTA was launched with:
    // 128 x 56 + 256 x 224 = 64512): the producer warpgroup needs a handful of registers, a
    // softmax thread keeps one full 128-score row (+ the O row in the epilogue) in registers.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA loader
        if (lane == 0) {
            int it = 0, g = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
                const int qp = item % p.nqp;
                const int h = (item / p.nqp) % p.H;
                const int b = item / (p.nqp * p.H);
                const int nqb = ((qp * 2 + 1) * BLK < p.S) ? 2 : 1;  // the odd last query block has no partner
                mbar_wait(q_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, nqb * C::TILE_BYTES);
                for (int qb = 0; qb < nqb; ++qb) {
                    const int tok0 = (qp * 2 + qb) * BLK;
                    tma_load_5d(sQ + qb * C::TILE_BYTES, &map_main, q_full, 0, h, 0, tok0, b);
                    if (C::TAIL) tma_load_5d(sQ + qb * C::TILE_BYTES + C::MAIN_BYTES, &map_tail, q_full, 64, h, 0, tok0, b);
                }
                for (int j = 0; j < J; ++j, ++g) {
                    const int st = g & 1;
                    const uint32_t ph = (g >> 1) & 1;
                    mbar_wait(&k_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&k_full[st], C::TILE_BYTES);
                    tma_load_5d(sK + st * C::TILE_BYTES, &map_main, &k_full[st], 0, h, 1, j * BLK, b);
                    if (C::TAIL) tma_load_5d(sK + st * C::TILE_BYTES + C::MAIN_BYTES, &map_tail, &k_full[st], 64, h, 1, j * BLK, b);
                    mbar_wait(&v_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&v_full[st], C::TILE_BYTES);
                    tma_load_5d(sV + st * C::TILE_BYTES, &map_main, &v_full[st], 0, h, 2, j * BLK, b);
                    if (C::TAIL) tma_load_5d(sV + st * C::TILE_BYTES + C::MAIN_BYTES, &map_tail, &v_full[st], 64, h, 2, j * BLK, b);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const bool f16 = p.f16 != 0;
            const uint32_t idesc_pv = umma_idesc_16_bmn(BLK, 64, f16);
            const uint32_t idesc_pvt = umma_idesc_16_bmn(BLK, C::TAIL ? C::TAIL : 16, f16);
            const uint32_t q_u32 = smem_u32(sQ), k_u32 = smem_u32(sK), v_u32 = smem_u32(sV), p_u32 = smem_u32(sP);

            // kcols = keys of this block rounded up to 32: the ragged last block (S = 257 -> 1 key, 577 -> 65) runs a
            // narrower S MMA, fewer exponentials and fewer P·V K-steps instead of a full 128-key block
            auto issue_s = [&](int qb, int st, int kcols) {
                const uint32_t d = tmem_base + TM_S + qb * TM_STRIDE;
                const uint32_t a0 = q_u32 + qb * C::TILE_BYTES, b0 = k_u32 + st * C::TILE_BYTES;
                const uint64_t ad = umma_desc_sw128(a0), bd = umma_desc_sw128(b0);
                const uint32_t idesc_s = umma_idesc_16(BLK, kcols, f16, f16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(d, ad + 2 * ks, bd + 2 * ks, idesc_s, ks > 0 ? 1u : 0u);
                if (C::TAIL) {
                    const uint64_t at = umma_desc(a0 + C::MAIN_BYTES, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
                    const uint64_t bt = umma_desc(b0 + C::MAIN_BYTES, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
#pragma unroll
                    for (int ks = 0; ks < C::TAIL / 16; ++ks) umma_bf16_ss(d, at + 2 * ks, bt + 2 * ks, idesc_s, 1u);
                }
            };
            auto issue_pv = [&](int qb, int st, bool acc, int kcols) {
                const uint32_t d = tmem_base + TM_O + qb * TM_STRIDE;
                const uint32_t a0 = p_u32 + qb * C::P_BYTES, b0 = v_u32 + st * C::TILE_BYTES;
                const int ksteps = kcols >> 4;  // 16 keys per step
#pragma unroll 1
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t ad = umma_desc_sw128(a0 + (ks >> 2) * 16384) + 2 * (ks & 3);
                    // V main tile: token-major rows of 128 B (64 head_dim columns) = MN-major SWIZZLE_128B operand;
                    // 8-key groups are 1024 B apart, one K-step (16 keys) = 2048 B
                    const uint64_t bd = umma_desc(b0 + ks * 2048, 16, 1024, 2);
                    umma_bf16_ss(d, ad, bd, idesc_pv, (acc || ks > 0) ? 1u : 0u);
                    if (C::TAIL) {
                        const uint64_t bt = umma_desc(b0 + C::MAIN_BYTES + ks * 16 * C::TAIL_ROW, 16, 8 * C::TAIL_ROW,
                                                      C::TAIL_LAYOUT);
                        umma_bf16_ss(d + 64, ad, bt, idesc_pvt, (acc || ks > 0) ? 1u : 0u);
                    }
                }
            };
            auto kcols_of = [&](int j) {
                const int valid = p.S - j * BLK;
                return valid >= BLK ? BLK : ((valid + 31) >> 5) << 5;
            };
            auto nqb_of = [&](int k) {
                const int qp = (static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x)) % p.nqp;
                return ((qp * 2 + 1) * BLK < p.S) ? 2 : 1;
            };

            // The key blocks of all of this CTA's items form one stream t = 0 .. T-1 (item t / J, block t % J): the
            // scores of block t+1 are issued before P·V of block t — across item boundaries too — so the softmax warps
            // always find their next S tile ready.
            const int n_local = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                static_cast<int>(gridDim.x);
            const int T = n_local * J;
            uint32_t n[2] = {0, 0};    // S tiles issued per query block
            uint32_t npv[2] = {0, 0};  // P·V blocks issued per query block
            uint32_t itq[2] = {0, 0};  // items started per query block (o_empty phase)
            auto issue_s_stage = [&](int t) {
                const int k = t / J, j = t - k * J, st = t & 1;
                if (j == 0) mbar_wait(q_full, k & 1);
                mbar_wait(&k_full[st], (t >> 1) & 1);
                const int nq = nqb_of(k), kcols = kcols_of(j);
                for (int qb = 0; qb < nq; ++qb) {
                    mbar_wait(&s_empty[qb], (n[qb] & 1) ^ 1);
                    tc_fence_after();
                    if (!(p.dbg & 4)) issue_s(qb, st, kcols);
                    umma_commit(&s_full[qb]);
                    ++n[qb];
                }
                umma_commit(&k_empty[st]);
                if (j == J - 1) umma_commit(q_empty);
            };
            if (T > 0) issue_s_stage(0);
            for (int t = 0; t < T; ++t) {
                const int k = t / J, j = t - k * J, st = t & 1;
                // next block's scores go first, except across an item boundary: there they would wait for the next
                // item's Q tile and hold up this item's last P·V (and with it the epilogue)
                const bool s_first = j + 1 < J;
                if (s_first && t + 1 < T) issue_s_stage(t + 1);
                const int nq = nqb_of(k), kcols = kcols_of(j);
                mbar_wait(&v_full[st], (t >> 1) & 1);
                for (int qb = 0; qb < nq; ++qb) {
                    mbar_wait(&p_full[qb], npv[qb] & 1);
                    if (j == 0) {
                        mbar_wait(&o_empty[qb], (itq[qb] & 1) ^ 1);
                        ++itq[qb];
                    }
                    tc_fence_after();
                    if (!(p.dbg & 2)) issue_pv(qb, st, j > 0, kcols);
                    umma_commit(&pv_done[qb]);
                    ++npv[qb];
                }
                umma_commit(&v_empty[st]);
                if (!s_first && t + 1 < T) issue_s_stage(t + 1);
            }
        }
    }
    } else {
        // ------------------------------------------------------------------------------------------ softmax warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int qb = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t s_addr = lane_base + TM_S + qb * TM_STRIDE;
        const uint32_t o_addr = lane_base + TM_O + qb * TM_STRIDE;
        uint8_t* prow = sP + qb * C::P_BYTES + row * 128;
        const uint32_t prow_u32 = smem_u32(prow);
        const bool f16 = p.f16 != 0;
        uint32_t n = 0;  // key blocks processed by this query block (global count)
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            const int qp = item % p.nqp;
            const int h = (item / p.nqp) % p.H;
            const int b = item / (p.nqp * p.H);
            if ((qp * 2 + qb) * BLK >= p.S) continue;  // no such query block (loader and MMA warp skip it too)
            const int tok = (qp * 2 + qb) * BLK + row;
            // warps whose 32 rows all lie past the last token keep the barrier protocol but skip the arithmetic
            const bool warp_live = (qp * 2 + qb) * BLK + quarter * 32 < p.S;
            float m = 0.f, l = 0.f;
            for (int j = 0; j < J; ++j, ++n) {
                const int valid = p.S - j * BLK;
                const int nch = valid >= BLK ? 4 : (valid + 31) >> 5;  // 32-key chunks of this block
                mbar_wait(&s_full[qb], n & 1);
                tc_fence_after();
                uint32_t v[BLK];
                if (warp_live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) tmem_ld_32x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[qb]);
                float factor = 1.f;
                bool need = false;
                if (warp_live && !(p.dbg & 8)) {
                    if (valid < nch * 32) {
#pragma unroll
                        for (int i = 0; i < BLK; ++i)
                            if (i >= valid) v[i] = 0xFF800000u;  // -inf: key past the end of the image
                    }
                    // four independent max chains (one per 32-key chunk) instead of one 128-long dependent chain
                    float mxc[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        mxc[c] = __uint_as_float(v[0]);
                        if (c < nch) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) mxc[c] = fmaxf(mxc[c], __uint_as_float(v[c * 32 + i]));
                        }
                    }
                    const float mx = fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3]));
                    const float tmax = mx * p.scale_log2;
                    if (j == 0) {
                        m = tmax;
                    } else if (tmax > m + 8.f) {
                        factor = ex2_approx(m - tmax);
                        m = tmax;
                        need = true;
                    }
                }
                if (j > 0) {
                    // P·V of the previous block has finished: P (smem) and O (TMEM) may be touched again
                    mbar_wait(&pv_done[qb], (n - 1) & 1);
                    if (__any_sync(0xffffffffu, need)) {
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < C::DHP / 16; ++c) {
                            uint32_t o[16];
                            tmem_ld_32x16(o_addr + c * 16, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                            tmem_st_32x16(o_addr + c * 16, o);
                        }
                        tmem_st_wait();
                        l *= factor;
                    }
                }
                if (warp_live && !(p.dbg & 1)) {
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
                    const float neg_m = -m, sc = p.scale_log2;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {  // 16 chunks of 8 keys = 16 bytes of P
                        if ((c >> 2) < nch) {
                            // in place, 8 independent FFMA then 8 independent MUFU.EX2 (no shared temporary: the SFU
                            // pipe is the bound of this kernel, the chain FFMA -> EX2 -> FFMA must not serialise)
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                v[c * 8 + i] = __float_as_uint(fmaf(__uint_as_float(v[c * 8 + i]), sc, neg_m));
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                v[c * 8 + i] = __float_as_uint(ex2_approx(__uint_as_float(v[c * 8 + i])));
#pragma unroll
                            for (int i = 0; i < 8; ++i) sum[i & 3] += __uint_as_float(v[c * 8 + i]);
                            uint32_t w[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float e0 = __uint_as_float(v[c * 8 + 2 * i]), e1 = __uint_as_float(v[c * 8 + 2 * i + 1]);
                                w[i] = f16 ? pack_f16x2(e0, e1) : pack_bf16x2(e0, e1);
                            }
                            // K-major SWIZZLE_128B A tile: key group (c >> 3) is a [128 x 128 B] slab, 16-byte chunk
                            // (c & 7) of this row sits at chunk index (c & 7) ^ (row & 7)
                            const uint32_t dst = prow_u32 + (c >> 3) * 16384 + ((((c & 7) ^ (row & 7))) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]),
                                         "r"(w[2]), "r"(w[3])
                                         : "memory");
                        }
                    }
                    l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
                }
                fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[qb]);
            }
            // ---- epilogue: O / l -> 16-bit -> ctx[b, tok, h*dh : (h+1)*dh]
            mbar_wait(&pv_done[qb], (n - 1) & 1);
            tc_fence_after();
            const float inv = 1.f / l;
            uint16_t* orow = static_cast<uint16_t*>(p.out) + (static_cast<size_t>(b) * p.S + tok) * p.D + h * DH;
            const bool tok_ok = tok < p.S;
            uint32_t o[C::DHP];
#pragma unroll
            for (int c = 0; c < C::DHP / 16; ++c) tmem_ld_32x16(o_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&o[c * 16]));
            tmem_ld_wait();
            tc_fence_before();  // O is in registers: hand the accumulator back to the MMA warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[qb]);
            if (tok_ok) {
#pragma unroll
                for (int c = 0; c < C::DHP / 8; ++c) {
                    if (c * 8 < DH) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x0 = __uint_as_float(o[c * 8 + 2 * i]) * inv, x1 = __uint_as_float(o[c * 8 + 2 * i + 1]) * inv;
                            w[i] = f16 ? pack_f16x2(x0, x1) : pack_bf16x2(x0, x1);
                        }
                        *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct MapKey {
    const void* ptr;
    int B, S, H, dh, f16;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && B == o.B && S == o.S && H == o.H && dh == o.dh && f16 == o.f16;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        for (int v : {k.B, k.S, k.H, k.dh, k.f16}) h = h * 1000003u ^ static_cast<size_t>(v);
        return h;
    }
};
struct MapPair {
    CUtensorMap main, tail;
};

template <int DH>
int get_maps(const void* qkv, bool f16, int B, int S, int H, MapPair* out) {
    using C = TcCfg<DH>;
    static std::mutex mu;
    static std::unordered_map<MapKey, MapPair, MapKeyHash> cache;
    MapKey key{qkv, B, S, H, DH, f16 ? 1 : 0};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return VLMFW_OK;
        }
    }
    const uint64_t D = static_cast<uint64_t>(H) * DH;
    // dims: head_dim, head, q|k|v, token, image
    const uint64_t dims[5] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(H), 3, static_cast<uint64_t>(S),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[4] = {static_cast<uint64_t>(DH) * 2, D * 2, 3 * D * 2, static_cast<uint64_t>(S) * 3 * D * 2};
    const uint32_t box_main[5] = {64, 1, 1, BLK, 1};
    MapPair mp;
    int rc = encode_tensor_map(&mp.main, f16, 5, qkv, dims, strides, box_main, 128);
    if (rc != VLMFW_OK) return rc;
    if (C::TAIL) {
        const uint32_t box_tail[5] = {static_cast<uint32_t>(C::TAIL), 1, 1, BLK, 1};
        rc = encode_tensor_map(&mp.tail, f16, 5, qkv, dims, strides, box_tail, C::TAIL_ROW);
        if (rc != VLMFW_OK) return rc;
    } else {
        mp.tail = mp.main;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 1024) cache.clear();
        cache.emplace(key, mp);
    }
    *out = mp;
    return VLMFW_OK;
}

template <int DH>
int launch_tc(const void* qkv, bool f16, int batch, int tokens, int heads, void* out, cudaStream_t stream) {
    using C = TcCfg<DH>;
    MapPair mp;
    int rc = get_maps<DH>(qkv, f16, batch, tokens, heads, &mp);
    if (rc != VLMFW_OK) return rc;
    auto kern = attention_tc_kernel<DH>;
    static bool attr_set = false;
    if (!attr_set) {
        VLMFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    TcParams p;
    p.out = out;
    p.S = tokens;
    p.H = heads;
    p.B = batch;
    p.D = heads * DH;
    const int nq = (tokens + BLK - 1) / BLK;
    p.nqp = (nq + 1) / 2;
    p.J = nq;
    p.num_items = batch * heads * p.nqp;
    p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    p.f16 = f16 ? 1 : 0;
    {
        const char* e = getenv("VLMFW_ATTN_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
    kern<<<grid, TC_THREADS, C::SMEM_BYTES, stream>>>(mp.main, mp.tail, p);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

}  // namespace

// head_dim 64 / 72 / 88 (the three towers) run on tcgen05; returns 1 when the shape is not covered (caller falls back to
// the generic mma.sync kernel of attention.cu, which takes any head_dim % 8 == 0 up to 96).
int launch_attention_tc(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                        cudaStream_t stream) {
    const size_t row_bytes = static_cast<size_t>(3) * heads * head_dim * 2;
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0 || row_bytes % 16 != 0) return 1;
    switch (head_dim) {
        case 64:
            return launch_tc<64>(qkv, f16, batch, tokens, heads, out, stream);
        case 72:
            return launch_tc<72>(qkv, f16, batch, tokens, heads, out, stream);
        case 88:
            return launch_tc<88>(qkv, f16, batch, tokens, heads, out, stream);
        default:
            return 1;
    }
}

}  // namespace vlmfw
