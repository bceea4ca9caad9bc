// Fused multi-head attention, second generation ("wide block"): softmax(Q K^T / sqrt(dh)) V, no mask, eval mode.
// Restates hf integrations/sdpa_attention.py:92 as called from modeling_siglip.py:285-309, modeling_clip.py:310-334 and
// modeling_blip_2.py:328-351 — same contract and layouts as attention_tc.cu (qkv [B*S, 3D] in, ctx [B*S, D] out).
//
// What the round-1 kernel (attention_tc.cu, P through shared memory) measured: 188 us per SigLIP layer at B = 256, tensor
// pipe 23.7 % active — every 64-key block paid a fixed chain of TMEM load, P tile to shared memory + proxy fence, three
// mbarrier round trips and eight small P.V MMAs issued by one thread at ~75 cycles each. This kernel:
//   * keeps the probabilities out of shared memory: each softmax thread holds its row's 64 scores of a key block in
//     registers, writes P as packed 16-bit pairs into tensor memory (tcgen05.st) and P.V takes its A operand from there
//     (tcgen05.mma with A in TMEM) — no st.shared, no fence.proxy.async, no P tiles (64 KB per CTA freed);
//   * gives P its own two TMEM buffers, so the ONE score buffer goes back to the score issuer as soon as the softmax warps
//     hold the block in registers: the score MMA of block t + 1 runs under the exponentials of block t, and the softmax
//     warps request block t + 1 from TMEM right after handing P of block t to the store (software pipeline across items);
//   * issues P.V per 64-key block by TWO single-thread issuers for the head dims with a tail tile (72 = 64 + 8 -> 16,
//     88 = 64 + 24 -> 32): main-tile and tail-tile MMAs write disjoint TMEM columns, so they need no order between them;
//   * folds the lone last key of the class-token towers (CLIP 577 = 9 x 64 + 1, EVA 257 = 4 x 64 + 1) in on CUDA cores
//     (one dot product by the epilogue warpgroup, one exponential and one axpy per query row) instead of a block step;
//   * feeds everything from ONE loader thread that blocks in order on Q, K(t), V(t - 1) — two hardware-sleep waits and
//     four TMA issues per block. The first loader polled three streams with test_wait and rebuilt coordinates by
//     division: 2250 cycles per block, as long as the softmax chain itself (found with the exponentials compiled out:
//     the kernel got only 8 % faster and the score issuer then waited for K).
// Persistent, TWO CTAs per SM (256 TMEM columns each: S [0,64), P0 / P1 [64,128), O [128, 128 + dh_padded)), one thread
// = one query row = one TMEM lane, fp32 softmax with exp2 and a lazily raised running max, O / l -> 16-bit -> shared-
// memory staging -> TMA store into ctx[b, tok, h, :].
//
// Roles (384 threads): warpgroup 0 = TMA loader, score issuer, P.V issuer (main tile), P.V issuer (tail tile; idle for
// head_dim 64); warpgroup 1 = softmax; warpgroup 2 = lone-key scores + epilogue.
//
// Where the time goes now (clock64 timeline of one CTA, SigLIP, 2550 cycles per 64-key block and CTA): reading the 32 KB
// score block out of tensor memory ~430, max + 64 exponentials + pack ~1150 (917 with the SM to itself: the MUFU pipe
// takes 8 cycles per warp instruction), P store + hand-offs ~600, barrier waits ~350. Compiling the exponentials AND the
// 16-bit pack out leaves 2390: the block period is the serial chain of one warp per row quarter — TMEM load, arithmetic,
// TMEM store, three mbarrier hand-offs — of which two CTAs per SM overlap at most two. Tried on top and dropped (no
// change): K and V of a block through one TMA (a box two deep along q|k|v), a half-period start offset between the two
// CTAs of an SM, polynomial exp2 on the FMA pipe for part of the elements.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int WB_THREADS = 384;  // warpgroup 0: loader + three MMA issuers; 1: softmax; 2: lone-key scores + epilogue
constexpr int QB = 128;   // query rows per item (= TMEM lanes)
constexpr int KBW = 64;   // keys per block (= score columns of one of the TWO score buffers in TMEM)
constexpr int WB_SMEM_PER_CTA = 233472 / 2 - 1024;  // two CTAs per SM, 1 KB reserved per CTA

template <int DH>
struct WbCfg {
    static constexpr int TAIL = DH == 64 ? 0 : (DH == 72 ? 16 : 32);  // columns of the second head_dim group (zero padded)
    static constexpr int DHP = 64 + TAIL;
    static constexpr int NPV = TAIL ? 2 : 1;                           // P.V issuer threads
    static constexpr int TAIL_ROW = TAIL * 2;                          // bytes per row of the tail tile (= its swizzle span)
    static constexpr uint32_t TAIL_LAYOUT = TAIL == 16 ? 6u : 4u;      // SWIZZLE_32B : SWIZZLE_64B
    static constexpr int Q_MAIN = QB * 128;
    static constexpr int Q_BYTES = Q_MAIN + QB * TAIL_ROW;             // one Q slot: [main tile | tail tile]
    static constexpr int KV_MAIN = KBW * 128;
    static constexpr int KV_BYTES = KV_MAIN + KBW * TAIL_ROW;          // one K or V slot: [main tile | tail tile]
    // ctx staging lives in the Q slot the item has consumed: one slab of 32 rows x DH 16-bit values per softmax warp
    // (1024-byte aligned slabs for head_dim 64, whose ctx map is 128B-swizzled; 128-byte aligned otherwise)
    static constexpr int STG_SLAB = 32 * DH * 2;
    static_assert(4 * STG_SLAB <= Q_BYTES, "ctx staging must fit in a Q slot");
    // [2 Q slots][k_L row | v_L row] (256 B each) + per-row hand-offs between the warpgroups: lone-key score (float) and
    // the softmax result (l, p_L) (float2), both [2 Q slots][128 rows]
    static constexpr int LONE_BYTES = 1024 + 2 * QB * 4 + 2 * QB * 8;
    static constexpr int BAR_BYTES = 1024;  // keeps SMEM_BYTES a multiple of 1 KB (the allocation unit): both CTAs of an SM start 1 KB aligned
    static constexpr int FIXED = 2 * Q_BYTES + LONE_BYTES + BAR_BYTES;
    // K and V rings of 64-key tiles: as deep as fits (K deeper than V when the count is odd: K is needed a block earlier)
    static constexpr int FIT = (WB_SMEM_PER_CTA - FIXED) / KV_BYTES;
    static_assert(FIT >= 4, "two K and two V tiles must fit");
    static constexpr int NK = FIT >= 8 ? 4 : (FIT + 1) / 2;
    static constexpr int NV = FIT >= 8 ? 4 : FIT / 2;
    static constexpr int SMEM_BYTES = FIXED + (NK + NV) * KV_BYTES;
    static_assert(SMEM_BYTES <= WB_SMEM_PER_CTA, "two CTAs per SM must fit");
};

struct WbParams {
    const void* qkv;
    int S, H, B, D;
    int s_eff;      // keys that go through the tensor core (S, or S - 1 when the last key is handled on CUDA cores)
    int lone;       // 1: key S - 1 is folded in by the softmax threads
    int nq;         // query blocks per (image, head)
    int J;          // key blocks per (image, head)
    int num_items;  // B * H * nq
    float scale_log2;
};

// TMEM columns (256 allocated per CTA): ONE score buffer S [0,64) (fp32 scores of the current 64-key block), two
// probability buffers P0 [64,96), P1 [96,128) (packed 16-bit pairs; block t uses buffer t & 1), O [128, 128 + DHP).
// S belongs to the score issuer again as soon as the softmax warps hold the block in registers (s_free), so the score MMA
// of block t + 1 runs under the exponentials of block t instead of behind the P.V MMAs of block t - 1 (the first layout
// of this kernel kept P in the columns its scores came from: two S buffers, each held until its P.V had retired, and the
// softmax warps could not fetch block t + 1 before they had finished block t).
constexpr uint32_t TM_S = 0;
constexpr uint32_t TM_O = 128;
__device__ __forceinline__ constexpr uint32_t tm_p(int buf) { return 64u + static_cast<uint32_t>(buf) * (KBW / 2); }

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns of packed 16-bit pairs per K-step of 16
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 64 consecutive columns as two 32-column loads: ptxas refuses a 64-register instruction under the 80-register launch cap
// of a 2 x 384-thread kernel, even inside the setmaxnreg.inc region of the softmax warpgroup
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t* v) {
    tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(v));
    tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(v + 32));
}

// registers -> TMEM, 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// 1-D bulk copy global -> shared (bytes a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

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

template <bool F16>
__device__ __forceinline__ float2 unpack16x2(uint32_t u) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    else return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}

// bottleneck experiments (variant builds only): -DWB_EXP_NOP replaces the exponential by a multiply, -DWB_PACK_NOP the
// 16-bit pack by a bit operation — results are wrong, only the timeline is looked at
#ifdef WB_EXP_NOP
#define WB_EXP2(x) ((x) * 0.5f)
#else
#define WB_EXP2(x) ex2_approx(x)
#endif
#ifdef WB_PACK_NOP
#define WB_PACK(a, b) (__float_as_uint(a) ^ (__float_as_uint(b) >> 16))
#else
#define WB_PACK(a, b) pack_16x2<F16>(a, b)
#endif

#ifdef VLMFW_ATT_TRACE
// timeline experiment (variant builds only): CTA 0 logs (clock << 8 | event) per role
__device__ long long g_wb_trace[4][8192];
#define WTR_DECL int tr_n = 0
#define WTR(role, code)                                                           \
    do {                                                                          \
        if (blockIdx.x == 0 && tr_n < 8192) g_wb_trace[role][tr_n++] = (clock64() << 8) | (code); \
    } while (0)
#else
#define WTR_DECL
#define WTR(role, code)
#endif

template <int DH, bool F16>
__global__ void __launch_bounds__(WB_THREADS, 2)
attention_wb_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                    const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_kvt,
                    const __grid_constant__ CUtensorMap map_o, const WbParams p) {
    using C = WbCfg<DH>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled tiles need the 1024-byte alignment (no slack is budgeted)
    // every slot is a multiple of 1 KB: [main tile (128B swizzle) | tail tile (32B / 64B swizzle)]
    uint8_t* sQ = smem;                              // [2][Q_BYTES]
    uint8_t* sK = sQ + 2 * C::Q_BYTES;               // [NK][KV_BYTES]
    uint8_t* sV = sK + C::NK * C::KV_BYTES;          // [NV][KV_BYTES]
    uint8_t* sLone = sV + C::NV * C::KV_BYTES;       // [2][k_L row 256 B | v_L row 256 B] + hand-off arrays
    uint64_t* bars = reinterpret_cast<uint64_t*>(sLone + C::LONE_BYTES);
    uint64_t* q_full = bars + 0;    // [2]
    uint64_t* q_empty = bars + 2;   // [2] last score MMA of the item retired AND the four softmax warps are done with the slot
    uint64_t* s_full = bars + 4;    // score buffer written by the score MMA (completion t = block t)
    uint64_t* s_free = bars + 5;    // the four softmax warps have block t's scores in registers
    uint64_t* p_empty = bars + 6;   // [2] all P.V MMAs of the block retired: its P buffer may be overwritten
    uint64_t* p_full = bars + 8;    // [2] P of the block written to TMEM by the four softmax warps
    uint64_t* o_full = bars + 10;
    uint64_t* o_empty = bars + 11;
    uint64_t* k_full = bars + 12;   // [4]
    uint64_t* k_empty = bars + 16;  // [4]
    uint64_t* v_full = bars + 20;   // [4]
    uint64_t* v_empty = bars + 24;  // [4]
    uint64_t* l_full = bars + 28;   // [2][4] (l, p_L) of the item's rows written by softmax warp `quarter`
    uint64_t* sl_full = bars + 36;  // [2][4] lone-key scores of the item's rows written by epilogue warp `quarter`
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
    float* sScore = reinterpret_cast<float*>(sLone + 1024);             // [2][QB]
    float2* sStat = reinterpret_cast<float2*>(sLone + 1024 + 2 * QB * 4);  // [2][QB]

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
        mbar_init(o_full, C::NPV);
        mbar_init(o_empty, 4);
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&p_empty[i], C::NPV);
            mbar_init(&p_full[i], 4);
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1 + 4);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], C::NPV);
        }
        for (int i = 0; i < 8; ++i) {
            mbar_init(&l_full[i], 1);
            mbar_init(&sl_full[i], 1);
        }
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
    // keys of block j that go through the tensor core, rounded up to the MMA's N / K granularity of 16
    auto kcols_of = [&](int j) {
        const int valid = p.s_eff - j * KBW;
        return valid >= KBW ? KBW : ((valid + 15) >> 4) << 4;
    };
    auto item_coords = [&](int it, int& b, int& h, int& qb) {
        const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        qb = item % p.nq;
        h = (item / p.nq) % p.H;
        b = item / (p.nq * p.H);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA loader
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // warpgroup 0
        // Like the GEMM's producer: the whole warp runs the loop and waits, one elected lane issues (elect.sync), so the
        // tensor-map operands and coordinates stay in uniform registers instead of an ELECT + R2UR loop per instruction.
        {
            const uint8_t* qkv8 = static_cast<const uint8_t*>(p.qkv);
            const size_t row_bytes = static_cast<size_t>(3) * p.D * 2;
            WTR_DECL;
            // ONE in-order request stream with blocking waits: Q of item `it` (slot it & 1, free once item it - 2 has been
            // stored), then per 64-key block the K tile of block t and the V tile of block t - 1 — K runs one block ahead
            // because the score MMA of block t is issued before the P.V MMAs of block t - 1. The first form of this loader
            // polled three streams (Q, K, V) without blocking and rebuilt every block's coordinates by division: ~2250
            // cycles of this one thread per block (clock64 timeline: three mbarrier.test_wait at ~150 cycles each, four TMA
            // issues at ~250, the divisions) — as long as the softmax chain it was feeding, so that nothing done to the
            // softmax warps could show (with the exponentials compiled out the kernel ran only 8 % faster and the score
            // issuer waited for K). Now: two try_waits (hardware sleep), four TMA issues, no division per block. Being
            // parked on a Q slot is harmless: the request comes NK blocks ahead of the consumers, the slot frees within ~2.
            int t = 0;
            int pb = 0, ph = 0, pj = 0;  // image / head / block of the V tile still to request (block t - 1)
            auto load_v = [&](int tv, int b, int h, int j) {
                const int vs = tv % C::NV;
                mbar_wait(&v_empty[vs], ((tv / C::NV) & 1) ^ 1);
                WTR(0, 5);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&v_full[vs], C::KV_BYTES);
                    tma_load_5d(sV + vs * C::KV_BYTES, &map_kv, &v_full[vs], 0, h, 2, j * KBW, b);
                    if (C::TAIL) tma_load_5d(sV + vs * C::KV_BYTES + C::KV_MAIN, &map_kvt, &v_full[vs], 64, h, 2, j * KBW, b);
                }
                __syncwarp();
            };
            for (int it = 0; it < n_local; ++it) {
                int b, h, qb;
                item_coords(it, b, h, qb);
                const int qs = it & 1;
                mbar_wait(&q_empty[qs], ((it >> 1) & 1) ^ 1);
                WTR(0, 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[qs], C::Q_BYTES + (p.lone ? 2 * DH * 2 : 0));
                    tma_load_5d(sQ + qs * C::Q_BYTES, &map_q, &q_full[qs], 0, h, 0, qb * QB, b);
                    if (C::TAIL) tma_load_5d(sQ + qs * C::Q_BYTES + C::Q_MAIN, &map_qt, &q_full[qs], 64, h, 0, qb * QB, b);
                    if (p.lone) {  // the last token's k and v rows of this head, for the CUDA-core fold-in
                        const uint8_t* rowp = qkv8 + (static_cast<size_t>(b) * p.S + p.S - 1) * row_bytes + static_cast<size_t>(h) * DH * 2;
                        bulk_load_1d(sLone + qs * 512, rowp + static_cast<size_t>(p.D) * 2, DH * 2, &q_full[qs]);
                        bulk_load_1d(sLone + qs * 512 + 256, rowp + static_cast<size_t>(2 * p.D) * 2, DH * 2, &q_full[qs]);
                    }
                }
                __syncwarp();
                for (int j = 0; j < J; ++j, ++t) {
                    const int ks = t % C::NK;
                    mbar_wait(&k_empty[ks], ((t / C::NK) & 1) ^ 1);
                    WTR(0, 3);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&k_full[ks], C::KV_BYTES);
                        tma_load_5d(sK + ks * C::KV_BYTES, &map_kv, &k_full[ks], 0, h, 1, j * KBW, b);
                        if (C::TAIL) tma_load_5d(sK + ks * C::KV_BYTES + C::KV_MAIN, &map_kvt, &k_full[ks], 64, h, 1, j * KBW, b);
                    }
                    __syncwarp();
                    if (t > 0) load_v(t - 1, pb, ph, pj);
                    pb = b, ph = h, pj = j;
                }
            }
            if (t > 0) load_v(t - 1, pb, ph, pj);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ score issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // warpgroup 0
        {
            int t = 0;
            WTR_DECL;
            for (int it = 0; it < n_local; ++it) {
                const int qs = it & 1;
                const uint32_t q_u32 = smem_u32(sQ) + qs * C::Q_BYTES;
                WTR(1, 1);
                mbar_wait(&q_full[qs], (it >> 1) & 1);
                for (int j = 0; j < J; ++j, ++t) {
                    const int ks = t % C::NK;
                    WTR(1, 2);
                    mbar_wait(&k_full[ks], (t / C::NK) & 1);
                    WTR(1, 3);
                    mbar_wait(s_free, (t & 1) ^ 1);  // the softmax warps have taken block t - 1 out of the score buffer
                    WTR(1, 4);
                    tc_fence_after();
                    const uint32_t d = tmem_base + TM_S;
                    const uint32_t k_u32 = smem_u32(sK) + ks * C::KV_BYTES;
                    const uint64_t ad = umma_desc_sw128(q_u32);
                    const uint64_t bd = umma_desc_sw128(k_u32);
                    const uint32_t idesc = umma_idesc_16(QB, kcols_of(j), F16, F16);
                    const uint64_t at = umma_desc(q_u32 + C::Q_MAIN, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
                    const uint64_t bt = umma_desc(k_u32 + C::KV_MAIN, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT);
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) umma_bf16_ss(d, ad + 2 * s, bd + 2 * s, idesc, s > 0 ? 1u : 0u);
                        if (C::TAIL) {
#pragma unroll
                            for (int s = 0; s < C::TAIL / 16; ++s) umma_bf16_ss(d, at + 2 * s, bt + 2 * s, idesc, 1u);
                        }
                        umma_commit(s_full);
                        umma_commit(&k_empty[ks]);
                        if (j == J - 1) umma_commit(&q_empty[qs]);
                    }
                    __syncwarp();
                    WTR(1, 5);
                }
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ------------------------------------------------------------------------------------------ P.V issuers
        // warp 2: O[:, 0:64) += P V_main; warp 3 (head dims with a tail tile): O[:, 64:64+TAIL) += P V_tail. Disjoint
        // accumulator columns: no order between the two streams; every barrier they release counts NPV arrivals.
        // Chunks are consumed in the order the softmax warps produce them: the upper 64 keys first (see there).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // warpgroup 0
        const bool tail_role = warp == 3;
        if (!tail_role || C::TAIL) {
            const uint32_t idesc = tail_role ? umma_idesc_16_bmn(QB, C::TAIL ? C::TAIL : 16, F16) : umma_idesc_16_bmn(QB, 64, F16);
            const uint32_t d = tmem_base + TM_O + (tail_role ? 64u : 0u);
            int t = 0;
#ifdef VLMFW_ATT_TRACE
            int tr_n = tail_role ? 8192 : 0;
#endif
            for (int it = 0; it < n_local; ++it) {
                for (int j = 0; j < J; ++j, ++t) {
                    const int kcols = kcols_of(j);
                    const int vs = t % C::NV, sb = t & 1;
                    if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);  // the previous item's O has been read out
                    WTR(3, 1);
                    mbar_wait(&v_full[vs], (t / C::NV) & 1);
                    WTR(3, 2);
                    mbar_wait(&p_full[sb], (t >> 1) & 1);
                    WTR(3, 4);
                    tc_fence_after();
                    const uint32_t b0 = smem_u32(sV) + vs * C::KV_BYTES + (tail_role ? C::KV_MAIN : 0);
                    const int steps = kcols >> 4;  // 16 keys per K-step
                    const bool issuer = elect_one();
                    auto step = [&](int s) {
                        const uint32_t a = tmem_base + tm_p(sb) + static_cast<uint32_t>(s * 8);
                        // V main tile: token-major rows of 128 B = MN-major SWIZZLE_128B operand, 8-key groups 1024 B apart,
                        // one K-step (16 keys) = 2048 B; tail tile: rows of TAIL_ROW bytes in its own swizzle
                        const uint64_t bd = tail_role ? umma_desc(b0 + s * 16 * C::TAIL_ROW, 16, 8 * C::TAIL_ROW, C::TAIL_LAYOUT)
                                                      : umma_desc(b0 + s * 2048, 16, 1024, 2);
                        umma_f16_ts(d, a, bd, idesc, (j > 0 || s > 0) ? 1u : 0u);
                    };
                    if (issuer) {
                        if (steps == KBW / 16) {
#pragma unroll
                            for (int s = 0; s < KBW / 16; ++s) step(s);
                        } else {
#pragma unroll 1
                            for (int s = 0; s < steps; ++s) step(s);
                        }
                        umma_commit(&v_empty[vs]);
                        umma_commit(&p_empty[sb]);
                        if (j == J - 1) umma_commit(o_full);
                    }
                    __syncwarp();
                    WTR(3, 7);
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------------------------------ softmax warps
        // Register budget: the CTA is launched with 80 registers per thread (2 CTAs x 384 threads); warpgroup 0 gives back
        // down to 48, the epilogue warpgroup down to 56, which leaves 136 for these four warps.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t o_addr = lane_base + TM_O;
        const float sc = p.scale_log2;
        int t = 0;
#ifdef VLMFW_ATT_TRACE
        int tr_n = (warp == 4 && lane == 0) ? 0 : 8192;
#endif
        // Software pipeline over the CTA's key blocks (across item boundaries): the scores of block t + 1 are requested from
        // TMEM as soon as block t's P has been handed to the tensor memory store, so the load's latency runs under the
        // store's completion, the p_full hand-off and the loop turn instead of in front of the next block's arithmetic.
        const int T = n_local * J;
        uint32_t v[KBW];
        bool have = false;  // v holds (an in-flight load of) block t's scores
        for (int it = 0; it < n_local; ++it) {
            int b, h, qb;
            item_coords(it, b, h, qb);
            const int qs = it & 1;
            // warps whose 32 rows all lie past the last token (ragged last query block: EVA 257, CLIP 577) keep the
            // barrier protocol but skip the TMEM reads and the arithmetic
            const bool warp_live = qb * QB + quarter * 32 < p.S;
            bool next_item_live = false;
            if (it + 1 < n_local) {
                int b2, h2, qb2;
                item_coords(it + 1, b2, h2, qb2);
                next_item_live = qb2 * QB + quarter * 32 < p.S;
            }
            float sL = -INFINITY, pL = 0.f;  // lone last key: its score comes from the epilogue warpgroup
            float m = 0.f, l = 0.f;
            for (int j = 0; j < J; ++j, ++t) {
                const int valid = p.s_eff - j * KBW;
                const int pb = t & 1;
                const uint32_t s_addr = lane_base + TM_S;
                WTR(2, 1);
                if (!have) {
                    mbar_wait(s_full, t & 1);
                    tc_fence_after();
                    if (warp_live) tmem_ld_32x64(s_addr, v);
                }
                WTR(2, 2);
                tmem_ld_wait();
                // the block is in registers: the score buffer goes back to the issuer (block t + 1's MMA may overwrite it)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free);
                uint32_t w[KBW / 2];
                if (warp_live) {
                    if (valid < KBW) {
#pragma unroll
                        for (int i = 0; i < KBW; ++i)
                            if (i >= valid) v[i] = 0xFF800000u;  // -inf: key past the end / column the narrow MMA never wrote
                    }
                    float mx[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        mx[g] = __uint_as_float(v[g * 16]);
#pragma unroll
                        for (int i = 1; i < 16; ++i) mx[g] = fmaxf(mx[g], __uint_as_float(v[g * 16 + i]));
                    }
                    const float tmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;
                    WTR(2, 3);
                    float factor = 1.f;
                    bool need = false;
                    if (j == 0) {
                        if (p.lone) {
                            mbar_wait(&sl_full[qs * 4 + quarter], (it >> 1) & 1);
                            sL = sScore[qs * QB + row];
                        }
                        m = fmaxf(tmax, sL);
                        if (p.lone) {
                            pL = ex2_approx(sL - m);
                            l = pL;
                        }
                    } else if (tmax > m + 8.f) {
                        factor = ex2_approx(m - tmax);
                        m = tmax;
                        need = true;
                    }
                    if (j > 0 && __any_sync(0xffffffffu, need)) {
                        // O must hold every P.V up to block t - 1 before it is rescaled: that block's MMAs signal its P
                        // buffer's "empty" barrier when they retire (and the MMAs of one issuer retire in order)
                        mbar_wait(&p_empty[pb ^ 1], ((t - 1) >> 1) & 1);
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
                        pL *= factor;
                    }
                    // exponentials, row sum, P as packed 16-bit pairs
                    const float neg_m = -m;
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int g = 0; g < KBW / 8; ++g) {
                        // 8 independent FFMA, then 8 independent MUFU.EX2 (the chain FFMA -> EX2 -> FADD must not serialise)
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[g * 8 + i] = __float_as_uint(fmaf(__uint_as_float(v[g * 8 + i]), sc, neg_m));
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[g * 8 + i] = __float_as_uint(WB_EXP2(__uint_as_float(v[g * 8 + i])));
#pragma unroll
                        for (int i = 0; i < 8; ++i) sum[i & 3] += __uint_as_float(v[g * 8 + i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            w[g * 4 + i] = WB_PACK(__uint_as_float(v[g * 8 + 2 * i]), __uint_as_float(v[g * 8 + 2 * i + 1]));
                    }
                    l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
                } else {
                    // dead warp: its lanes of P must still be defined (they feed MMA rows whose output is never stored)
#pragma unroll
                    for (int i = 0; i < KBW / 2; ++i) w[i] = 0u;
                }
                WTR(2, 4);
                // P.V of block t - 2 has read this P buffer
                mbar_wait(&p_empty[pb], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                tmem_st_32x32(lane_base + tm_p(pb), w);
                // block t + 1: if its scores are already there (the usual case), start the load now
                have = false;
                if (t + 1 < T) {
                    have = __all_sync(0xffffffffu, mbar_test(s_full, (t + 1) & 1));
                    if (have) {
                        tc_fence_after();
                        if (j + 1 < J ? warp_live : next_item_live) tmem_ld_32x64(s_addr, v);
                        WTR(2, 6);
                    } else {
                        WTR(2, 7);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[pb]);
                WTR(2, 5);
            }
            // hand (l, p_L) to the epilogue warpgroup and move on to the next item
            sStat[qs * QB + row] = make_float2(l, pL);
            __syncwarp();
            if (lane == 0) mbar_arrive(&l_full[qs * 4 + quarter]);
            WTR(2, 8);
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue warps
        // (a) the lone last key's score s_L = q . k_L for the NEXT item (CUDA cores, from the shared-memory Q tile and the
        //     k_L row the loader fetched: fp32 accumulation of the exact 16-bit products, like the MMA), one item ahead so
        //     the softmax warps never wait for it; (b) the item's epilogue (O + p_L v_L) / l -> 16-bit -> staging -> TMA
        //     store into ctx[b, tok, h*dh : (h+1)*dh], after which the Q slot (whose bytes the staging reuses) goes back
        //     to the loader. The softmax warps used to do both: 3600 of their 11200 cycles per SigLIP item.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t o_addr = lane_base + TM_O;
        const float sc = p.scale_log2;
        auto lone_score = [&](int it) {
            int b, h, qb;
            item_coords(it, b, h, qb);
            const int qs = it & 1;
            mbar_wait(&q_full[qs], (it >> 1) & 1);
            const uint32_t qrow = smem_u32(sQ) + qs * C::Q_BYTES + row * 128;
            const uint32_t qtail = smem_u32(sQ) + qs * C::Q_BYTES + C::Q_MAIN + row * C::TAIL_ROW;
            const uint32_t krow = smem_u32(sLone) + qs * 512;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
                // 16-byte chunk c of this row: main tile 128B swizzle (chunk ^ row % 8); tail tile 32B swizzle
                // (chunk ^ (row / 4) % 2) or 64B swizzle (chunk ^ (row / 2) % 4)
                uint32_t qa_addr;
                if (c < 8) qa_addr = qrow + ((static_cast<uint32_t>(c) ^ (row & 7)) << 4);
                else if (C::TAIL == 16) qa_addr = qtail + ((static_cast<uint32_t>(c - 8) ^ ((row >> 2) & 1)) << 4);
                else qa_addr = qtail + ((static_cast<uint32_t>(c - 8) ^ ((row >> 1) & 3)) << 4);
                const uint4 qa = lds128(qa_addr);
                const uint4 ka = lds128(krow + c * 16);
                const uint32_t qw[4] = {qa.x, qa.y, qa.z, qa.w}, kw[4] = {ka.x, ka.y, ka.z, ka.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = unpack16x2<F16>(qw[i]), y = unpack16x2<F16>(kw[i]);
                    a0 = fmaf(x.x, y.x, a0);
                    a1 = fmaf(x.y, y.y, a1);
                }
            }
            sScore[qs * QB + row] = (a0 + a1) * sc;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sl_full[qs * 4 + quarter]);
        };
#ifdef VLMFW_ATT_TRACE
        int tr_n = (warp == 8 && lane == 0) ? 4096 : 8192;  // upper half of the softmax role's buffer
#endif
        if (p.lone && n_local > 0) lone_score(0);
        for (int it = 0; it < n_local; ++it) {
            int b, h, qb;
            item_coords(it, b, h, qb);
            const int qs = it & 1;
            const bool warp_live = qb * QB + quarter * 32 < p.S;
            if (p.lone && it + 1 < n_local) lone_score(it + 1);
            WTR(2, 20);
            mbar_wait(&l_full[qs * 4 + quarter], (it >> 1) & 1);
            const float2 st = sStat[qs * QB + row];
            mbar_wait(o_full, it & 1);
            WTR(2, 21);
            tc_fence_after();
            // staging slab: this warp's 32 rows inside the Q slot the item has consumed (o_full implies that every score
            // MMA that read the slot has retired; this warpgroup's own reads of it are behind it in program order)
            const uint32_t slab = smem_u32(sQ) + qs * C::Q_BYTES + quarter * C::STG_SLAB;
            const uint32_t dst = slab + lane * (DH * 2);
            const uint32_t vrow = smem_u32(sLone) + qs * 512 + 256;
            // head_dim 64: rows are exactly 128 B, every lane would hit the same banks — that map is 128B-swizzled
            // (chunk c of row r at c ^ (r & 7)); 144- / 176-byte rows spread over the banks by themselves
            const uint32_t osw = DH == 64 ? (static_cast<uint32_t>(lane) & 7u) : 0u;
            const float inv = __fdividef(1.f, st.x), pL = st.y;
            if (warp_live) {
#pragma unroll
                for (int g = 0; g < (DH + 15) / 16; ++g) {  // 16 accumulator columns at a time (56 registers in this warpgroup)
                    uint32_t o[16];
                    tmem_ld_32x16(o_addr + g * 16, o);
                    tmem_ld_wait();
                    if (g == (DH + 15) / 16 - 1) {  // O is out of TMEM: hand the accumulator back to the P.V issuers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(o_empty);
                    }
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg) {
                        const int c = g * 2 + cg;  // 16-byte chunk of the output row
                        if (c * 8 >= DH) continue;
                        float x[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o[cg * 8 + i]);
                        if (p.lone) {
                            const uint4 va = lds128(vrow + c * 16);
                            const uint32_t vw[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 y = unpack16x2<F16>(vw[i]);
                                x[2 * i] = fmaf(pL, y.x, x[2 * i]);
                                x[2 * i + 1] = fmaf(pL, y.y, x[2 * i + 1]);
                            }
                        }
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) w[i] = pack_16x2<F16>(x[2 * i] * inv, x[2 * i + 1] * inv);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((static_cast<uint32_t>(c) ^ osw) << 4)),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                                     : "memory");
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&map_o, slab, 0, h, qb * QB + quarter * 32, b);  // rows past S are clipped
                    tma_store_commit();
                    tma_store_wait_read<0>();  // the slab has been read: the Q slot may be refilled
                    mbar_arrive(&q_empty[qs]);
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(o_empty);
                    mbar_arrive(&q_empty[qs]);
                }
            }
            WTR(2, 22);
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
struct WbKey {
    const void* ptr;
    const void* out;
    int B, S, H, dh, f16;
    bool operator==(const WbKey& o) const {
        return ptr == o.ptr && out == o.out && B == o.B && S == o.S && H == o.H && dh == o.dh && f16 == o.f16;
    }
};
struct WbKeyHash {
    size_t operator()(const WbKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr) * 31u ^ reinterpret_cast<size_t>(k.out);
        for (int v : {k.B, k.S, k.H, k.dh, k.f16}) h = h * 1000003u ^ static_cast<size_t>(v);
        return h;
    }
};
struct WbMaps {
    CUtensorMap q, qt, kv, kvt, o;
};

template <int DH>
int wb_maps(const void* qkv, const void* ctx, bool f16, int B, int S, int H, WbMaps* out) {
    using C = WbCfg<DH>;
    static std::mutex mu;
    static std::unordered_map<WbKey, WbMaps, WbKeyHash> cache;
    WbKey key{qkv, ctx, B, S, H, DH, f16 ? 1 : 0};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return VLMFW_OK;
        }
    }
    const uint64_t D = static_cast<uint64_t>(H) * DH;
    // dims: head_dim, head, q|k|v, token, image — columns past head_dim and tokens past S arrive as zeros
    const uint64_t dims[5] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(H), 3, static_cast<uint64_t>(S),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[4] = {static_cast<uint64_t>(DH) * 2, D * 2, 3 * D * 2, static_cast<uint64_t>(S) * 3 * D * 2};
    const uint32_t box_q[5] = {64, 1, 1, QB, 1}, box_kv[5] = {64, 1, 1, KBW, 1};
    WbMaps ms;
    int rc = encode_tensor_map(&ms.q, f16, 5, qkv, dims, strides, box_q, 128);
    if (rc != VLMFW_OK) return rc;
    rc = encode_tensor_map(&ms.kv, f16, 5, qkv, dims, strides, box_kv, 128);
    if (rc != VLMFW_OK) return rc;
    {
        // ctx [B, S, H, dh]: one 32-row x dh box per softmax warp; dense rows in shared memory, 128B-swizzled when a row is
        // exactly 128 B
        const uint64_t odims[4] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(H), static_cast<uint64_t>(S),
                                   static_cast<uint64_t>(B)};
        const uint64_t ostrides[3] = {static_cast<uint64_t>(DH) * 2, D * 2, static_cast<uint64_t>(S) * D * 2};
        const uint32_t obox[4] = {static_cast<uint32_t>(DH), 1, 32, 1};
        rc = encode_tensor_map(&ms.o, f16, 4, ctx, odims, ostrides, obox, DH == 64 ? 128 : 0);
        if (rc != VLMFW_OK) return rc;
    }
    if (C::TAIL) {
        const uint32_t box_qt[5] = {static_cast<uint32_t>(C::TAIL), 1, 1, QB, 1};
        const uint32_t box_kvt[5] = {static_cast<uint32_t>(C::TAIL), 1, 1, KBW, 1};
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
int launch_wb(const void* qkv, int batch, int tokens, int heads, void* out, cudaStream_t stream) {
    using C = WbCfg<DH>;
    WbMaps ms;
    int rc = wb_maps<DH>(qkv, out, F16, batch, tokens, heads, &ms);
    if (rc != VLMFW_OK) return rc;
    auto kern = attention_wb_kernel<DH, F16>;
    static std::atomic<uint64_t> smem_opt_in{0};  // per instantiation, one bit per device
    VLMFW_CUDA_OK(ensure_dynamic_smem(kern, C::SMEM_BYTES, smem_opt_in));
    WbParams p;
    p.qkv = qkv;
    p.S = tokens;
    p.H = heads;
    p.B = batch;
    p.D = heads * DH;
    // exactly one key past a multiple of 16 (the class token of CLIP 577 / EVA 257): CUDA cores take it
    p.lone = (tokens > 16 && tokens % 16 == 1) ? 1 : 0;
    p.s_eff = tokens - p.lone;
    p.nq = (tokens + QB - 1) / QB;
    p.J = (p.s_eff + KBW - 1) / KBW;
    p.num_items = batch * heads * p.nq;
    p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    int slots = 2 * num_sms();
    int smem_bytes = C::SMEM_BYTES;
#ifdef VLMFW_ATT_TRACE
    if (getenv("VLMFW_WB_ONE_CTA")) {  // timeline experiment: one CTA per SM (no MUFU / issue-slot sharing between CTAs)
        slots = num_sms();
        smem_bytes = 120 * 1024;
        VLMFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    }
#endif
    const int grid = p.num_items < slots ? p.num_items : slots;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(WB_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    VLMFW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ms.q, ms.qt, ms.kv, ms.kvt, ms.o, p));
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

}  // namespace

#ifdef VLMFW_ATT_TRACE
extern "C" int vlmfw_debug_wb_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_wb_trace, sizeof(g_wb_trace)) == cudaSuccess ? 0 : -1;
}
#endif

// head_dim 64 / 72 / 88 (the three towers); returns 1 when the shape is not covered (caller falls back to attention_tc.cu /
// the generic mma.sync kernel of attention.cu).
int launch_attention_wb(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                        cudaStream_t stream) {
    const size_t row_bytes = static_cast<size_t>(3) * heads * head_dim * 2;
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0 || row_bytes % 16 != 0) return 1;
    switch (head_dim) {
        case 64:
            return f16 ? launch_wb<64, true>(qkv, batch, tokens, heads, out, stream)
                       : launch_wb<64, false>(qkv, batch, tokens, heads, out, stream);
        case 72:
            return f16 ? launch_wb<72, true>(qkv, batch, tokens, heads, out, stream)
                       : launch_wb<72, false>(qkv, batch, tokens, heads, out, stream);
        case 88:
            return f16 ? launch_wb<88, true>(qkv, batch, tokens, heads, out, stream)
                       : launch_wb<88, false>(qkv, batch, tokens, heads, out, stream);
        default:
            return 1;
    }
}

}  // namespace vlmfw
