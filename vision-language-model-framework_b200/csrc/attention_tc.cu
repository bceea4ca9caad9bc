// Fused multi-head attention on the 5th-gen tensor cores: softmax(Q K^T / sqrt(dh)) V, no mask, eval mode.
// Restates hf integrations/sdpa_attention.py:92 as called from modeling_siglip.py:285-309, modeling_clip.py:310-334 and
// modeling_blip_2.py:328-351 — same contract and layouts as attention.cu (qkv [B*S, 3D] in, ctx [B*S, D] out).
//
// Persistent, warp-specialised, TWO CTAs per SM (192 threads, 256 TMEM columns and ~100 KB of shared memory each); a
// work item = (image, head, one block of 128 queries), keys in blocks of 64:
//   warp 0       TMA loader  — 5-D tensor maps over qkv (head_dim, head, q|k|v, token, image): out-of-range head_dim
//                              columns and tokens past S arrive as zeros, so head_dim 72 / 88 need no padding in HBM;
//                              Q tile + 2-deep rings of K and V tiles, running ahead across items
//   warp 1       MMA issuer  — S_j = Q K_j^T (tcgen05.mma 128 x 64 x dh into one of TWO score buffers in TMEM) and
//                              O += P_j V_j (128 x dh x 64, V read MN-major straight from the token-major tile); the
//                              scores of block j+1 are issued before P·V of block j, across item boundaries too
//   warps 2..5   softmax     — one thread = one query row = one TMEM lane: tcgen05.ld the 64 scores, row max / exp2 /
//                              row sum in registers with no shuffles, P -> fp16/bf16 -> one of TWO 128B-swizzled smem
//                              tiles (the A operand of P·V); then the item's epilogue O / l -> ctx
// The first version paired two query blocks in one 384-thread CTA per SM with 128-key blocks: its ncu capture showed
// 13.5 % tensor-pipe activity and memory, SFU and MMA phases adding up instead of overlapping (0.31 ms where each of
// the three alone is ~0.1 ms at SigLIP B=256) — one short dependent chain per SM. Two independent CTAs per SM with
// double-buffered S and P give the SM four things to interleave.
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

constexpr int TC_THREADS = 256;  // warpgroup 0: loader, S issuer, P·V issuer, (idle); warpgroup 1: 4 softmax warps (TMEM lane quarter = warp % 4)
constexpr int QB = 128;          // query rows per item (= TMEM lanes)
constexpr int KB = 64;           // keys per block
#ifndef VLMFW_ATT_NKV
#define VLMFW_ATT_NKV 2
#endif
#ifndef VLMFW_ATT_NP
#define VLMFW_ATT_NP 2
#endif
constexpr int NKV_WANT = VLMFW_ATT_NKV;  // K / V ring depth asked for (<= 4); each head dim takes what fits
constexpr int NP = VLMFW_ATT_NP;         // P tiles (1 or 2)
constexpr int SMEM_PER_CTA = 233472 / 2 - 1024;  // two CTAs per SM, 1 KB reserved per CTA

template <int DH>
struct TcCfg {
    static constexpr int TAIL = DH == 64 ? 0 : (DH == 72 ? 16 : 32);  // columns of the second head_dim group (zero padded)
    static constexpr int DHP = 64 + TAIL;
    static constexpr int TAIL_ROW = TAIL * 2;                      // bytes per row of the tail tile (= its swizzle span)
    static constexpr uint32_t TAIL_LAYOUT = TAIL == 16 ? 6u : 4u;  // SWIZZLE_32B : SWIZZLE_64B
    static constexpr int Q_MAIN = QB * 128;
    static constexpr int Q_BYTES = Q_MAIN + QB * TAIL_ROW;
    static constexpr int KV_MAIN = KB * 128;
    static constexpr int KV_BYTES = KV_MAIN + KB * TAIL_ROW;
    static constexpr int P_BYTES = QB * KB * 2;
    static constexpr int BAR_BYTES = 256;
#ifdef VLMFW_ATT_NOSLACK
    static constexpr int ALIGN_SLACK = 0;  // dynamic shared memory declared 1024-byte aligned (checked at run time)
#else
    static constexpr int ALIGN_SLACK = 1024;
#endif
    static constexpr int FIXED = Q_BYTES + NP * P_BYTES + BAR_BYTES + ALIGN_SLACK;
    static constexpr int NKV_FIT = (SMEM_PER_CTA - FIXED) / (2 * KV_BYTES);
    static constexpr int NKV = NKV_FIT < NKV_WANT ? NKV_FIT : NKV_WANT;
    static_assert(NKV >= 2, "K / V rings need two stages");
    static constexpr int SMEM_BYTES = FIXED + 2 * NKV * KV_BYTES;
};

struct TcParams {
    void* out;
    int S, H, B, D;
    int nq;         // query blocks per (image, head)
    int J;          // key blocks per (image, head)
    int num_items;  // B * H * nq
    float scale_log2;
    int f16;
    int dbg;  // VLMFW_ATTN_DBG bit mask (bottleneck experiments only): 1 skip exp+P store, 2 skip P·V MMAs, 4 skip S MMAs,
              // 8 skip tail-tile loads, 16 / 64 / 128 skip V / K / Q loads, 32 skip the ctx stores
};

// TMEM columns (256 allocated per CTA): S0 [0,64) S1 [64,128) O [128, 128 + DHP)
constexpr uint32_t TM_S = 0, TM_O = 128;

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

#ifdef VLMFW_ATT_TRACE
// timeline experiment (variant builds only): CTA 0 logs (clock << 8 | event) per role
__device__ long long g_att_trace[4][8192];
#define TR_DECL int tr_n = 0
#define TR(role, code)                                                            \
    do {                                                                          \
        if (blockIdx.x == 0 && tr_n < 8192) g_att_trace[role][tr_n++] = (clock64() << 8) | (code); \
    } while (0)
#else
#define TR_DECL
#define TR(role, code)
#endif

template <int DH, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                    const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_kvt,
                    const __grid_constant__ CUtensorMap map_o, const TcParams p) {
    using C = TcCfg<DH>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    if (C::ALIGN_SLACK == 0 && smem != smem_raw) __trap();  // swizzled tiles need the 1024-byte alignment
    uint8_t* sQ = smem;                         // [Q_BYTES]
    uint8_t* sK = sQ + C::Q_BYTES;              // [C::NKV][KV_BYTES]
    uint8_t* sV = sK + C::NKV * C::KV_BYTES;       // [C::NKV][KV_BYTES]
    uint8_t* sP = sV + C::NKV * C::KV_BYTES;       // [NP][P_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NP * C::P_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* o_full = bars + 2;
    uint64_t* o_empty = bars + 3;
    uint64_t* s_full = bars + 4;    // [2] score buffers
    uint64_t* s_empty = bars + 6;   // [2]
    uint64_t* p_full = bars + 8;    // [2] probability tiles
    uint64_t* p_empty = bars + 10;  // [2] (= P·V of that block complete: P free, O updated)
    uint64_t* k_full = bars + 12;   // [4]
    uint64_t* k_empty = bars + 16;  // [4]
    uint64_t* v_full = bars + 20;   // [4]
    uint64_t* v_empty = bars + 24;  // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_kv);
        tma_prefetch_desc(&map_o);
        if (C::TAIL) {
            tma_prefetch_desc(&map_qt);
            tma_prefetch_desc(&map_kvt);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        mbar_init(o_empty, 4);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // the prologue above may overlap the QKV GEMM's tail; qkv is read (and ctx written) only from here on
    pdl_launch_dependents();
    const int J = p.J;
    const int n_local = static_cast<int>(blockIdx.x) < p.num_items
                            ? (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                  static_cast<int>(gridDim.x)
                            : 0;

    // Register split (setmaxnreg is per warpgroup): the three single-thread roles keep 48 registers, the softmax
    // warpgroup takes 208 (2 CTAs x (128 x 48 + 128 x 208) = the whole register file).
    // (the instruction sits at the top of each role's branch so that ptxas allocates per region)
    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA loader
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        // One in-order stream in CONSUMPTION order: the scores of block s are issued before P·V of block s - 1, so K
        // runs one block ahead of V — a V slot that frees late (after its P·V) never holds up a K load.
        if (lane == 0) {
            const int T = n_local * J;
            auto coords = [&](int s, int& b, int& h, int& j, int& qb) {
                const int k = s / J;
                j = s - k * J;
                const int item = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
                qb = item % p.nq;
                h = (item / p.nq) % p.H;
                b = item / (p.nq * p.H);
            };
            TR_DECL;
            for (int s = 0; s <= T; ++s) {
                int b, h, j, qb;
                if (s < T) {
                    coords(s, b, h, j, qb);
                    if (j == 0) {
                        const int k = s / J;
                        TR(0, 1);
                        mbar_wait(q_empty, (k & 1) ^ 1);
                        TR(0, 2);
                        const bool ld = !VLMFW_DBG(p, 128), ldt = C::TAIL && ld && !VLMFW_DBG(p, 8);
                        mbar_arrive_expect_tx(q_full, (ld ? C::Q_MAIN : 0) + (ldt ? C::Q_BYTES - C::Q_MAIN : 0));
                        if (ld) tma_load_5d(sQ, &map_q, q_full, 0, h, 0, qb * QB, b);
                        if (ldt) tma_load_5d(sQ + C::Q_MAIN, &map_qt, q_full, 64, h, 0, qb * QB, b);
                    }
                    const int st = s % C::NKV;
                    TR(0, 3);
                    mbar_wait(&k_empty[st], ((s / C::NKV) & 1) ^ 1);
                    TR(0, 4);
                    const bool ld = !VLMFW_DBG(p, 64), ldt = C::TAIL && ld && !VLMFW_DBG(p, 8);
                    mbar_arrive_expect_tx(&k_full[st], (ld ? C::KV_MAIN : 0) + (ldt ? C::KV_BYTES - C::KV_MAIN : 0));
                    if (ld) tma_load_5d(sK + st * C::KV_BYTES, &map_kv, &k_full[st], 0, h, 1, j * KB, b);
                    if (ldt) tma_load_5d(sK + st * C::KV_BYTES + C::KV_MAIN, &map_kvt, &k_full[st], 64, h, 1, j * KB, b);
                }
                if (s >= 1) {
                    const int sv = s - 1;
                    coords(sv, b, h, j, qb);
                    const int st = sv % C::NKV;
                    TR(0, 5);
                    mbar_wait(&v_empty[st], ((sv / C::NKV) & 1) ^ 1);
                    TR(0, 6);
                    const bool ld = !VLMFW_DBG(p, 16), ldt = C::TAIL && ld && !VLMFW_DBG(p, 8);
                    mbar_arrive_expect_tx(&v_full[st], (ld ? C::KV_MAIN : 0) + (ldt ? C::KV_BYTES - C::KV_MAIN : 0));
                    if (ld) tma_load_5d(sV + st * C::KV_BYTES, &map_kv, &v_full[st], 0, h, 2, j * KB, b);
                    if (ldt) tma_load_5d(sV + st * C::KV_BYTES + C::KV_MAIN, &map_kvt, &v_full[st], 64, h, 2, j * KB, b);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        // TWO single-thread issuers: warp 1 the scores S_t = Q K_t^T, warp 2 the products O += P_t V_t. One thread doing
        // both spent ~4200 cycles per key block in issue + commit instructions alone (timeline trace: a tcgen05.mma of
        // this size costs ~80 cycles to issue, a commit ~150) and was the critical path; the two streams only meet
        // through the softmax warps' barriers, so they need no ordering between them.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (lane == 0) {
            constexpr bool f16 = F16;
            const uint32_t idesc_pv = umma_idesc_16_bmn(QB, 64, f16);
            const uint32_t idesc_pvt = umma_idesc_16_bmn(QB, C::TAIL ? C::TAIL : 16, f16);
            const uint32_t q_u32 = smem_u32(sQ), k_u32 = smem_u32(sK), v_u32 = smem_u32(sV), p_u32 = smem_u32(sP);
            // keys of block j rounded up to the MMA's N granularity: the ragged last block (S = 257 -> 1 key,
            // 577 -> 1 key) runs a 16-wide S MMA, 16 exponentials per row and one P·V K-step
            auto kcols_of = [&](int j) {
                const int valid = p.S - j * KB;
                return valid >= KB ? KB : ((valid + 15) >> 4) << 4;
            };
            // The key blocks of all of this CTA's items form one stream t = 0 .. T-1 (item t / J, block t % J)
            const int T = n_local * J;
            TR_DECL;
            auto issue_s = [&](int t) {
                const int k = t / J, j = t - k * J, st = t % C::NKV, sb = t & 1;
                TR(1, 1);
                if (j == 0) mbar_wait(q_full, k & 1);
                TR(1, 2);
                mbar_wait(&k_full[st], (t / C::NKV) & 1);
                TR(1, 3);
                mbar_wait(&s_empty[sb], ((t >> 1) & 1) ^ 1);
                TR(1, 4);
                tc_fence_after();
                if (!VLMFW_DBG(p, 4)) {
                    const uint32_t d = tmem_base + TM_S + sb * KB;
                    const uint32_t b0 = k_u32 + st * C::KV_BYTES;
                    const uint64_t ad = umma_desc_sw128(q_u32), bd = umma_desc_sw128(b0);
                    const uint32_t idesc_s = umma_idesc_16(QB, kcols_of(j), f16, f16);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(d, ad + 2 * ks, bd + 2 * ks, idesc_s, ks > 0 ? 1u : 0u);
                    if (C::TAIL) {
                        const uint64_t at = umma_desc(q_u32 + C::Q_MAIN, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
                        const uint64_t bt = umma_desc(b0 + C::KV_MAIN, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
#pragma unroll
                        for (int ks = 0; ks < C::TAIL / 16; ++ks) umma_bf16_ss(d, at + 2 * ks, bt + 2 * ks, idesc_s, 1u);
                    }
                }
                umma_commit(&s_full[sb]);
                umma_commit(&k_empty[st]);
                if (j == J - 1) umma_commit(q_empty);
            };
            auto issue_pv = [&](int t) {
                const int k = t / J, j = t - k * J, st = t % C::NKV, pb = t % NP;
                TR(3, 5);
                mbar_wait(&v_full[st], (t / C::NKV) & 1);
                TR(3, 6);
                mbar_wait(&p_full[pb], (t / NP) & 1);
                TR(3, 7);
                if (j == 0) mbar_wait(o_empty, (k & 1) ^ 1);
                TR(3, 8);
                tc_fence_after();
                if (!VLMFW_DBG(p, 2)) {
                    const uint32_t d = tmem_base + TM_O;
                    const uint32_t a0 = p_u32 + pb * C::P_BYTES, b0 = v_u32 + st * C::KV_BYTES;
                    const int ksteps = kcols_of(j) >> 4;  // 16 keys per step
                    // main-tile MMAs, then tail-tile MMAs (two instruction descriptors). A full block is unrolled with
                    // compile-time offsets — the descriptor arithmetic then stays in uniform registers; in the rolled
                    // loop every MMA cost ~140 cycles to issue against ~70 for the (unrolled) score MMAs.
                    auto pv_step = [&](int ks, bool tail) {
                        const uint64_t ad = umma_desc_sw128(a0) + 2 * ks;
                        const uint32_t accum = (j > 0 || ks > 0) ? 1u : 0u;
                        if (!tail) {
                            // V main tile: token-major rows of 128 B (64 head_dim columns) = MN-major SWIZZLE_128B operand;
                            // 8-key groups are 1024 B apart, one K-step (16 keys) = 2048 B
                            umma_bf16_ss(d, ad, umma_desc(b0 + ks * 2048, 16, 1024, 2), idesc_pv, accum);
                        } else {
                            umma_bf16_ss(d + 64, ad,
                                         umma_desc(b0 + C::KV_MAIN + ks * 16 * C::TAIL_ROW, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT),
                                         idesc_pvt, accum);
                        }
                    };
                    if (ksteps == KB / 16) {
#pragma unroll
                        for (int ks = 0; ks < KB / 16; ++ks) pv_step(ks, false);
                        if (C::TAIL) {
#pragma unroll
                            for (int ks = 0; ks < KB / 16; ++ks) pv_step(ks, true);
                        }
                    } else {
#pragma unroll 1
                        for (int ks = 0; ks < ksteps; ++ks) pv_step(ks, false);
                        if (C::TAIL) {
#pragma unroll 1
                            for (int ks = 0; ks < ksteps; ++ks) pv_step(ks, true);
                        }
                    }
                }
                TR(3, 9);
                umma_commit(&p_empty[pb]);
                umma_commit(&v_empty[st]);
                if (j == J - 1) umma_commit(o_full);
            };
            if (warp == 1) {
                for (int t = 0; t < T; ++t) issue_s(t);
            } else {
                for (int t = 0; t < T; ++t) issue_pv(t);
            }
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    } else {
        // ------------------------------------------------------------------------------------------ softmax warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t o_addr = lane_base + TM_O;
        const uint32_t prow_u32 = smem_u32(sP) + row * 128;
        const uint32_t sw = static_cast<uint32_t>(row) & 7u;
        const float sc = p.scale_log2;
        int t = 0;  // key blocks processed by this CTA
        bool store_pending = false;
        static_assert(NP == 2 && 16 * DH * 2 <= 4096, "ctx staging uses this warp's rows of both P tiles");
#ifdef VLMFW_ATT_TRACE
        int tr_n = (warp == 4 && lane == 0) ? 0 : 8192;
#endif
        for (int it = 0; it < n_local; ++it) {
            const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int qb = item % p.nq;
            const int h = (item / p.nq) % p.H;
            const int b = item / (p.nq * p.H);
            const int tok = qb * QB + row;
            // warps whose 32 rows all lie past the last token (ragged last query block: EVA 257, CLIP 577) keep the
            // barrier protocol but skip the TMEM reads and the arithmetic
            const bool warp_live = qb * QB + quarter * 32 < p.S;
            const bool warp_live_q = warp_live;
            float m = 0.f, l = 0.f;
            for (int j = 0; j < J; ++j, ++t) {
                const int sb = t & 1;
                const int valid = p.S - j * KB;
                const int nch = valid >= KB ? 4 : (valid + 15) >> 4;  // 16-key chunks of this block
                TR(2, 1);
                mbar_wait(&s_full[sb], (t >> 1) & 1);
                TR(2, 2);
                tc_fence_after();
                uint32_t v[KB];
                const uint32_t s_addr = lane_base + TM_S + sb * KB;
                if (warp_live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) tmem_ld_32x16(s_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c * 16]));
                    tmem_ld_wait();
                }
                TR(2, 3);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sb]);
                if (!warp_live) {  // nothing to compute: release the P tile in step with the other warps
                    const int pbd = t % NP;
                    mbar_wait(&p_empty[pbd], ((t / NP) & 1) ^ 1);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[pbd]);
                    continue;
                }
                if (valid < nch * 16) {
#pragma unroll
                    for (int i = 0; i < KB; ++i)
                        if (i >= valid) v[i] = 0xFF800000u;  // -inf: key past the end of the image
                }
                // independent max chains (one per 16-key chunk) instead of one 64-long dependent chain; a ragged last
                // block has -inf in every column past `valid` (set above, also for the chunks that were not loaded)
                const bool full = nch == 4;
                if (!full) {
#pragma unroll
                    for (int c = 1; c < 4; ++c)
                        if (c >= nch) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[c * 16 + i] = 0xFF800000u;
                        }
                }
                float mxc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mxc[c] = __uint_as_float(v[c * 16]);
#pragma unroll
                    for (int i = 1; i < 16; ++i) mxc[c] = fmaxf(mxc[c], __uint_as_float(v[c * 16 + i]));
                }
                const float tmax = fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])) * sc;
                float factor = 1.f;
                bool need = false;
                if (j == 0) {
                    m = tmax;
                } else if (tmax > m + 8.f) {
                    factor = ex2_approx(m - tmax);
                    m = tmax;
                    need = true;
                }
                if (j > 0 && __any_sync(0xffffffffu, need)) {
                    // P·V of the previous block must have landed before O is rescaled
                    mbar_wait(&p_empty[(t - 1) % NP], ((t - 1) / NP) & 1);
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
                if (store_pending) {  // the previous item's ctx store still reads this warp's slabs of the P tiles
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    store_pending = false;
                }
                // this P tile was last read by P·V NP blocks ago
                const int pb = t % NP;
                TR(2, 4);
                mbar_wait(&p_empty[pb], ((t / NP) & 1) ^ 1);
                TR(2, 5);
                if (!VLMFW_DBG(p, 1)) {
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
                    const float neg_m = -m;
                    const uint32_t pdst = prow_u32 + pb * C::P_BYTES;
                    // straight-line over the whole 64-key block: the columns of a ragged last block hold -inf (exp2 -> 0)
                    // and its P chunks past kcols are written but never read by the narrower P·V
#pragma unroll
                    for (int c = 0; c < 8; ++c) {  // 8 chunks of 8 keys = 16 bytes of P
                        // in place, 8 independent FFMA then 8 independent MUFU.EX2 (the chain FFMA -> EX2 -> FADD must
                        // not serialise)
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
                            w[i] = pack_16x2<F16>(e0, e1);
                        }
                        // K-major SWIZZLE_128B A tile [128 rows x 128 B]: 16-byte chunk c of this row at c ^ (row & 7)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pdst + ((static_cast<uint32_t>(c) ^ sw) << 4)),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                                     : "memory");
                    }
                    l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
                }
                TR(2, 6);
                fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[pb]);
                TR(2, 7);
            }
            // ---- epilogue: O / l -> 16-bit -> ctx[b, tok, h*dh : (h+1)*dh]
            TR(2, 8);
            mbar_wait(o_full, it & 1);
            TR(2, 9);
            tc_fence_after();
            uint32_t o[C::DHP];
#pragma unroll
            for (int c = 0; c < C::DHP / 16; ++c) tmem_ld_32x16(o_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&o[c * 16]));
            tmem_ld_wait();
            tc_fence_before();  // O is in registers: hand the accumulator back to the MMA warp
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            if (warp_live_q && !VLMFW_DBG(p, 32)) {
                // O / l -> 16-bit -> this warp's staging slabs (its own 32 rows of the two P tiles, free once the item's
                // last P·V has completed) -> two 16-row TMA stores into ctx[b, tok, h, :] (rows past S are clipped).
                // Row-strided 16-byte global stores from registers cost ~2300 cycles per item here (timeline trace).
                const float inv = 1.f / l;
                const uint32_t slab = smem_u32(sP) + (lane >> 4) * C::P_BYTES + quarter * 4096;
                const uint32_t dst = slab + (lane & 15) * (DH * 2);
                // head_dim 64: rows are exactly 128 B, every lane would hit the same banks — that map is 128B-swizzled
                // (chunk c of row r at c ^ (r & 7)); 144- / 176-byte rows spread over the banks by themselves
                const uint32_t osw = DH == 64 ? (static_cast<uint32_t>(lane) & 7u) : 0u;
#pragma unroll
                for (int c = 0; c < DH / 8; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x0 = __uint_as_float(o[c * 8 + 2 * i]) * inv, x1 = __uint_as_float(o[c * 8 + 2 * i + 1]) * inv;
                        w[i] = pack_16x2<F16>(x0, x1);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((static_cast<uint32_t>(c) ^ osw) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                                 "r"(w[3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const int tok0 = qb * QB + quarter * 32;
                    tma_store_4d(&map_o, smem_u32(sP) + quarter * 4096, 0, h, tok0, b);
                    tma_store_4d(&map_o, smem_u32(sP) + C::P_BYTES + quarter * 4096, 0, h, tok0 + 16, b);
                    tma_store_commit();
                }
                store_pending = true;
            }
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct MapKey {
    const void* ptr;
    const void* out;
    int B, S, H, dh, f16;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && out == o.out && B == o.B && S == o.S && H == o.H && dh == o.dh && f16 == o.f16;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr) * 31u ^ reinterpret_cast<size_t>(k.out);
        for (int v : {k.B, k.S, k.H, k.dh, k.f16}) h = h * 1000003u ^ static_cast<size_t>(v);
        return h;
    }
};
struct MapSet {
    CUtensorMap q, qt, kv, kvt, o;
};

template <int DH>
int get_maps(const void* qkv, const void* ctx, bool f16, int B, int S, int H, MapSet* out) {
    using C = TcCfg<DH>;
    static std::mutex mu;
    static std::unordered_map<MapKey, MapSet, MapKeyHash> cache;
    MapKey key{qkv, ctx, B, S, H, DH, f16 ? 1 : 0};
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
    const uint32_t box_q[5] = {64, 1, 1, QB, 1}, box_kv[5] = {64, 1, 1, KB, 1};
    MapSet ms;
    int rc = encode_tensor_map(&ms.q, f16, 5, qkv, dims, strides, box_q, 128);
    if (rc != VLMFW_OK) return rc;
    {
        // ctx [B, S, H, dh]: 16-row x dh boxes; dense rows in shared memory, 128B-swizzled when a row is exactly 128 B
        const uint64_t odims[4] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(H), static_cast<uint64_t>(S),
                                   static_cast<uint64_t>(B)};
        const uint64_t ostrides[3] = {static_cast<uint64_t>(DH) * 2, D * 2, static_cast<uint64_t>(S) * D * 2};
        const uint32_t obox[4] = {static_cast<uint32_t>(DH), 1, 16, 1};
        rc = encode_tensor_map(&ms.o, f16, 4, ctx, odims, ostrides, obox, DH == 64 ? 128 : 0);
        if (rc != VLMFW_OK) return rc;
    }
    rc = encode_tensor_map(&ms.kv, f16, 5, qkv, dims, strides, box_kv, 128);
    if (rc != VLMFW_OK) return rc;
    if (C::TAIL) {
        const uint32_t box_qt[5] = {static_cast<uint32_t>(C::TAIL), 1, 1, QB, 1};
        const uint32_t box_kvt[5] = {static_cast<uint32_t>(C::TAIL), 1, 1, KB, 1};
        rc = encode_tensor_map(&ms.qt, f16, 5, qkv, dims, strides, box_qt, C::TAIL_ROW);
        if (rc != VLMFW_OK) return rc;
        rc = encode_tensor_map(&ms.kvt, f16, 5, qkv, dims, strides, box_kvt, C::TAIL_ROW);
        if (rc != VLMFW_OK) return rc;
    } else {
        ms.qt = ms.q;
        ms.kvt = ms.kv;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 1024) cache.clear();
        cache.emplace(key, ms);
    }
    *out = ms;
    return VLMFW_OK;
}

template <int DH, bool F16>
int launch_tc(const void* qkv, int batch, int tokens, int heads, void* out, int q_rows, cudaStream_t stream) {
    using C = TcCfg<DH>;
    constexpr bool f16 = F16;
    MapSet ms;
    int rc = get_maps<DH>(qkv, out, f16, batch, tokens, heads, &ms);
    if (rc != VLMFW_OK) return rc;
    auto kern = attention_tc_kernel<DH, F16>;
    static std::atomic<uint64_t> smem_opt_in{0};  // per instantiation, one bit per device
    VLMFW_CUDA_OK(ensure_dynamic_smem(kern, C::SMEM_BYTES, smem_opt_in));
    TcParams p;
    p.out = out;
    p.S = tokens;
    p.H = heads;
    p.B = batch;
    p.D = heads * DH;
    p.nq = (q_rows + QB - 1) / QB;
    p.J = (tokens + KB - 1) / KB;
    p.num_items = batch * heads * p.nq;
    p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    p.f16 = f16 ? 1 : 0;
#ifdef VLMFW_DEBUG_MASKS
    {
        const char* e = getenv("VLMFW_ATTN_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
#else
    p.dbg = 0;
#endif
    const int slots = 2 * num_sms();
    const int grid = p.num_items < slots ? p.num_items : slots;
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        VLMFW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ms.q, ms.qt, ms.kv, ms.kvt, ms.o, p));
    }
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

}  // namespace

#ifdef VLMFW_ATT_TRACE
extern "C" int vlmfw_debug_att_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_att_trace, sizeof(g_att_trace)) == cudaSuccess ? 0 : -1;
}
#endif

// head_dim 64 / 72 / 88 (the three towers) run on tcgen05; returns 1 when the shape is not covered (caller falls back to
// the generic mma.sync kernel of attention.cu, which takes any head_dim % 8 == 0 up to 96).
int launch_attention_tc(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                        int q_rows, cudaStream_t stream) {
    const size_t row_bytes = static_cast<size_t>(3) * heads * head_dim * 2;
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0 || row_bytes % 16 != 0) return 1;
    switch (head_dim) {
        case 64:
            return f16 ? launch_tc<64, true>(qkv, batch, tokens, heads, out, q_rows, stream)
                       : launch_tc<64, false>(qkv, batch, tokens, heads, out, q_rows, stream);
        case 72:
            return f16 ? launch_tc<72, true>(qkv, batch, tokens, heads, out, q_rows, stream)
                       : launch_tc<72, false>(qkv, batch, tokens, heads, out, q_rows, stream);
        case 88:
            return f16 ? launch_tc<88, true>(qkv, batch, tokens, heads, out, q_rows, stream)
                       : launch_tc<88, false>(qkv, batch, tokens, heads, out, q_rows, stream);
        default:
            return 1;
    }
}

}  // namespace vlmfw
