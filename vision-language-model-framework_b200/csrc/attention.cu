// Fused multi-head attention for the vision towers: softmax(Q K^T / sqrt(dh)) V, no mask, no dropout (eval).
// Restates hf integrations/sdpa_attention.py:92 (F.scaled_dot_product_attention) as called from
// modeling_siglip.py:285-309, modeling_clip.py:310-334 and modeling_blip_2.py:328-351.
//
// Layout: qkv bf16 [B*S, 3*D] straight out of the fused QKV GEMM — thirds are q | k | v, head h occupies columns
// [h*dh, (h+1)*dh) of each third (BLIP-2's packed (which, head, dh) order is the same thing). Output bf16 [B*S, D],
// head-major columns, which is exactly the A operand of the out-proj GEMM: no permute kernels anywhere.
//
// One CTA = 64 query rows of one (image, head); 4 warps x 16 rows. K/V stream through a 2-deep cp.async ring in
// 64-key tiles; scores live in registers (mma.sync m16n8k16, fp32 accumulate), online softmax in fp32 with
// exp2 and warp-shuffle quad reductions; P is re-used from registers as the A operand of P·V.
// head_dim 72 / 88 are zero-padded to 80 / 96 in shared memory only (K of QK^T must be a multiple of 16).
// S in {256, 257, 577}: the ragged last key tile is zero-filled and masked to -inf.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int BQ = 64;
constexpr int BKV = 64;
constexpr int ATT_THREADS = 128;

template <int DHP>
struct AttCfg {
    static constexpr int PITCH = DHP + 8;  // elements; +16 B keeps ldmatrix rows on distinct banks
    static constexpr int TILE_ELEMS = 64 * PITCH;
    static constexpr int SMEM_BYTES = 5 * TILE_ELEMS * 2;  // Q + 2 x (K, V)
    static constexpr int CHUNKS_PER_ROW = DHP / 8;         // 16-byte chunks
};

// Copy 64 rows x DHP columns (rows >= rows_valid and columns >= dh zero-filled) of one head slice into smem.
template <int DHP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* smem_tile, const __nv_bfloat16* gbase, size_t row_stride,
                                          int row0, int rows_total, int dh) {
    using C = AttCfg<DHP>;
    constexpr int TOTAL = 64 * C::CHUNKS_PER_ROW;
#pragma unroll
    for (int i = 0; i < TOTAL / ATT_THREADS; ++i) {
        const int idx = i * ATT_THREADS + threadIdx.x;
        const int r = idx / C::CHUNKS_PER_ROW;
        const int c = idx - r * C::CHUNKS_PER_ROW;
        const bool ok = (row0 + r < rows_total) && (c * 8 < dh);
        const __nv_bfloat16* src = ok ? gbase + static_cast<size_t>(row0 + r) * row_stride + c * 8 : gbase;
        cp_async_16(smem_tile + r * C::PITCH + c * 8, src, ok ? 16u : 0u);
    }
}

template <int DHP, bool F16>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                __nv_bfloat16* __restrict__ out, int tokens,
                                                                int heads, int dh, float scale_log2, int q_begin) {
    using C = AttCfg<DHP>;
    constexpr int KSTEPS = DHP / 16;  // k-steps of Q K^T
    constexpr int NT_O = DHP / 8;     // n-tiles of the output accumulator
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sK = sQ + C::TILE_ELEMS;      // [2][64][PITCH]
    __nv_bfloat16* sV = sK + 2 * C::TILE_ELEMS;  // [2][64][PITCH]

    const int q0 = q_begin + blockIdx.x * BQ;
    const int head = blockIdx.y;
    const int img = blockIdx.z;
    const int dmodel = heads * dh;
    const size_t row_stride = static_cast<size_t>(3) * dmodel;
    const __nv_bfloat16* qg = qkv + static_cast<size_t>(img) * tokens * row_stride + head * dh;
    const __nv_bfloat16* kg = qg + dmodel;
    const __nv_bfloat16* vg = qg + 2 * dmodel;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int nkv = (tokens + BKV - 1) / BKV;

    load_tile<DHP>(sQ, qg, row_stride, q0, tokens, dh);
    load_tile<DHP>(sK, kg, row_stride, 0, tokens, dh);
    load_tile<DHP>(sV, vg, row_stride, 0, tokens, dh);
    cp_async_commit();

    uint32_t qf[KSTEPS][4];
    float o[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
            load_tile<DHP>(sK + (buf ^ 1) * C::TILE_ELEMS, kg, row_stride, (j + 1) * BKV, tokens, dh);
            load_tile<DHP>(sV + (buf ^ 1) * C::TILE_ELEMS, vg, row_stride, (j + 1) * BKV, tokens, dh);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(smem_u32(sQ + r * C::PITCH + c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
        }
        const __nv_bfloat16* kb = sK + buf * C::TILE_ELEMS;
        const __nv_bfloat16* vb = sV + buf * C::TILE_ELEMS;

        // ---- scores: 16 query rows x 64 keys per warp
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
                uint32_t b0, b1, b2, b3;
                const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(smem_u32(kb + key * C::PITCH + c), b0, b1, b2, b3);
                mma_16816<F16>(s[2 * np], qf[kk], b0, b1);
                mma_16816<F16>(s[2 * np + 1], qf[kk], b2, b3);
            }
        }
        // ---- mask the ragged tail, online softmax (rows g and g+8 of this warp's 16)
        const int key_base = j * BKV;
        if (key_base + BKV > tokens) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k0 = key_base + i * 8 + 2 * t;
                if (k0 >= tokens) s[i][0] = s[i][2] = -INFINITY;
                if (k0 + 1 >= tokens) s[i][1] = s[i][3] = -INFINITY;
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], mscaled[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = exp2f((m_run[r] - m_new) * scale_log2);  // first tile: exp2(-inf) = 0
            m_run[r] = m_new;
            mscaled[r] = m_new * scale_log2;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = exp2f(s[i][0] * scale_log2 - mscaled[0]);
            s[i][1] = exp2f(s[i][1] * scale_log2 - mscaled[0]);
            s[i][2] = exp2f(s[i][2] * scale_log2 - mscaled[1]);
            s[i][3] = exp2f(s[i][3] * scale_log2 - mscaled[1]);
            rs[0] += s[i][0] + s[i][1];
            rs[1] += s[i][2] + s[i][3];
        }
        l_run[0] = l_run[0] * corr[0] + rs[0];
        l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
        for (int i = 0; i < NT_O; ++i) {
            o[i][0] *= corr[0];
            o[i][1] *= corr[0];
            o[i][2] *= corr[1];
            o[i][3] *= corr[1];
        }
        // ---- O += P V  (P straight from the score registers)
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
            uint32_t pa[4];
            pa[0] = pack_16x2<F16>(s[2 * ks][0], s[2 * ks][1]);
            pa[1] = pack_16x2<F16>(s[2 * ks][2], s[2 * ks][3]);
            pa[2] = pack_16x2<F16>(s[2 * ks + 1][0], s[2 * ks + 1][1]);
            pa[3] = pack_16x2<F16>(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
            for (int np = 0; np < NT_O / 2; ++np) {
                uint32_t b0, b1, b2, b3;
                const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = np * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(smem_u32(vb + key * C::PITCH + c), b0, b1, b2, b3);
                mma_16816<F16>(o[2 * np], pa, b0, b1);
                mma_16816<F16>(o[2 * np + 1], pa, b2, b3);
            }
        }
        __syncthreads();  // everyone is done with `buf` before the next prefetch overwrites buf^1's partner
    }

    // ---- normalise, stage through smem (Q tile is dead), coalesced 16-byte stores
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    {
        const int r0 = warp * 16 + g;
#pragma unroll
        for (int i = 0; i < NT_O; ++i) {
            const int c = i * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(sQ + r0 * C::PITCH + c) = pack_16x2<F16>(o[i][0] * inv0, o[i][1] * inv0);
            *reinterpret_cast<uint32_t*>(sQ + (r0 + 8) * C::PITCH + c) = pack_16x2<F16>(o[i][2] * inv1, o[i][3] * inv1);
        }
    }
    __syncthreads();
    __nv_bfloat16* og = out + static_cast<size_t>(img) * tokens * dmodel + head * dh;
    const int chunks = dh >> 3;
    for (int idx = threadIdx.x; idx < BQ * chunks; idx += ATT_THREADS) {
        const int r = idx / chunks, c = idx - r * chunks;
        if (q0 + r < tokens) {
            const uint4 v = *reinterpret_cast<const uint4*>(sQ + r * C::PITCH + c * 8);
            *reinterpret_cast<uint4*>(og + static_cast<size_t>(q0 + r) * dmodel + c * 8) = v;
        }
    }
}

template <int DHP, bool F16>
int launch_att(const void* qkv, int batch, int tokens, int heads, int dh, void* out, cudaStream_t stream,
               int q_begin = 0) {
    using C = AttCfg<DHP>;
    auto kern = attention_kernel<DHP, F16>;
    static std::atomic<uint64_t> smem_opt_in{0};  // per instantiation, one bit per device
    VLMFW_CUDA_OK(ensure_dynamic_smem(kern, C::SMEM_BYTES, smem_opt_in));
    dim3 grid((tokens - q_begin + BQ - 1) / BQ, heads, batch);
    const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
    kern<<<grid, ATT_THREADS, C::SMEM_BYTES, stream>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                       static_cast<__nv_bfloat16*>(out), tokens, heads, dh, scale_log2,
                                                       q_begin);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}


// ---------------------------------------------------------------------------------------------- class-token query only
// Last layer of a tower whose adapter keeps only the class token (blip2.py:94-95 pooler_output, llava.py:114-115 "cls"):
// every other query row of that layer is dead work. One CTA = one (head, image): q = row 0, all `tokens` keys. A half-warp
// owns a key (lane = 16-byte chunk of the head slice: 11 of 16 lanes busy at dh = 88, loads coalesced per key); scores, the
// softmax and P.V stay fp32 (P is never rounded to 16 bits here). Fixed reduction order, deterministic. HBM-bound: the K
// and V slices are read once (2 * tokens * dh * 2 bytes per CTA).
constexpr int CLS_THREADS = 256;
constexpr int CLS_HW = CLS_THREADS / 16;  // half-warps per CTA

template <bool F16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (F16) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        } else {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
}

template <bool F16>
__global__ void __launch_bounds__(CLS_THREADS) attention_cls_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out,
                                                                    int tokens, int heads, int dh, int ld_out,
                                                                    float scale_log2) {
    extern __shared__ __align__(16) float cls_smem[];
    const int dmodel = heads * dh;
    const int chunks = dh >> 3;
    float* sq = cls_smem;                 // [128] q * scale * log2(e)
    float* sred = sq + 128;               // [16]
    float* sacc = sred + 16;              // [CLS_HW][128] partial outputs
    float* sp = sacc + CLS_HW * 128;      // [tokens] scores -> probabilities
    const int head = blockIdx.x, img = blockIdx.y;
    const size_t row_stride = static_cast<size_t>(3) * dmodel;
    const uint16_t* base = qkv + static_cast<size_t>(img) * tokens * row_stride + static_cast<size_t>(head) * dh;
    const int lane16 = threadIdx.x & 15, hw = threadIdx.x >> 4;
    if (threadIdx.x < chunks) {
        float f[8];
        unpack8<F16>(*reinterpret_cast<const uint4*>(base + threadIdx.x * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) sq[threadIdx.x * 8 + i] = f[i] * scale_log2;
    }
    __syncthreads();
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = lane16 < chunks ? sq[lane16 * 8 + i] : 0.f;
    const uint16_t* kb = base + dmodel + lane16 * 8;
    const uint16_t* vb = base + 2 * dmodel + lane16 * 8;
    // ---- scores (log2 domain)
    // (the trip count is the same for every thread: both half-warps of a warp take part in the full-mask shuffles below)
    for (int base = 0; base < tokens; base += 4 * CLS_HW) {
        const int j0 = base + hw;
        uint4 u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j0 + i * CLS_HW;
            u[i] = (j < tokens && lane16 < chunks) ? *reinterpret_cast<const uint4*>(kb + static_cast<size_t>(j) * row_stride)
                                                   : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float f[8], d = 0.f;
            unpack8<F16>(u[i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(q[e], f[e], d);
            d += __shfl_xor_sync(0xffffffffu, d, 8);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            const int j = j0 + i * CLS_HW;
            if (lane16 == 0 && j < tokens) sp[j] = d;
        }
    }
    __syncthreads();
    // ---- softmax: max and sum over the block (warp shuffles, then the 8 warp results in order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < tokens; j += CLS_THREADS) mx = fmaxf(mx, sp[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sred[warp] = mx;
    __syncthreads();
    mx = sred[0];
#pragma unroll
    for (int w = 1; w < CLS_THREADS / 32; ++w) mx = fmaxf(mx, sred[w]);
    float sum = 0.f;
    for (int j = threadIdx.x; j < tokens; j += CLS_THREADS) {
        const float e = exp2f(sp[j] - mx);
        sp[j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();  // every thread has read sred (max) before it is reused
    if (lane == 0) sred[8 + warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < CLS_THREADS / 32; ++w) sum += sred[8 + w];
    // ---- P.V: half-warp hw takes keys hw, hw + 16, ...; lane = chunk of 8 output columns
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int base = 0; base < tokens; base += 4 * CLS_HW) {
        const int j0 = base + hw;
        uint4 u[4];
        float pj[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j0 + i * CLS_HW;
            const bool ok = j < tokens && lane16 < chunks;
            u[i] = ok ? *reinterpret_cast<const uint4*>(vb + static_cast<size_t>(j) * row_stride) : make_uint4(0, 0, 0, 0);
            pj[i] = j < tokens ? sp[j] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float f[8];
            unpack8<F16>(u[i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj[i], f[e], acc[e]);
        }
    }
    if (lane16 < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sacc[hw * 128 + lane16 * 8 + e] = acc[e];
    }
    __syncthreads();
    if (threadIdx.x < dh / 2) {
        const int d = threadIdx.x * 2;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int g = 0; g < CLS_HW; ++g) {
            o0 += sacc[g * 128 + d];
            o1 += sacc[g * 128 + d + 1];
        }
        const float inv = 1.f / sum;
        *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(img) * ld_out + head * dh + d) = pack_16x2<F16>(o0 * inv, o1 * inv);
    }
}

}  // namespace

int launch_attention(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                     cudaStream_t stream) {
    VLMFW_REQUIRE(qkv && out, "attention: null pointer");
    VLMFW_REQUIRE(batch > 0 && tokens > 0 && heads > 0, "attention: bad shape");
    VLMFW_REQUIRE(head_dim % 8 == 0 && head_dim >= 8 && head_dim <= 96, "attention: head_dim %d unsupported", head_dim);
    VLMFW_REQUIRE(batch <= 65535 && heads <= 65535, "attention: grid too large");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    // head_dim 64 / 72 / 88 run on the tcgen05 kernel; anything else (and VLMFW_ATTN_LEGACY=1, for A/B timing) on the
    // generic mma.sync kernel below
    static const bool legacy = []() {
        const char* e = getenv("VLMFW_ATTN_LEGACY");
        return e != nullptr && e[0] == '1';
    }();
    // The tcgen05 kernel can be limited to the first q_tail query rows with the generic kernel taking the rest (disjoint
    // output rows, same stream). Measured on EVA (257 = 2 x 128 + the class token) the generic kernel's 64-query tile for
    // that one row costs more than the tcgen05 kernel's third, mostly idle, work item (0.445 vs 0.412 ms at B = 255), so
    // the split stays off and the tcgen05 kernel skips the arithmetic of softmax warps whose rows lie past the sequence.
    const int q_tail = tokens;
    // Second-generation tcgen05 kernel (attention_wb.cu: 128-key blocks, P kept in tensor memory) first; VLMFW_ATTN=tc
    // pins the round-1 kernel (attention_tc.cu) for A/B timing, and it stays the fallback for shapes wb does not cover.
    static const bool wb_on = []() {
        const char* e = getenv("VLMFW_ATTN");
        return !(e != nullptr && e[0] == 't');
    }();
    if (!legacy && wb_on) {
        rc = launch_attention_wb(qkv, f16, batch, tokens, heads, head_dim, out, stream);
        if (rc <= 0) return rc;
    }
    if (!legacy) {
        rc = launch_attention_tc(qkv, f16, batch, tokens, heads, head_dim, out, q_tail, stream);
        if (rc < 0) return rc;
        if (rc == 0 && q_tail == tokens) return rc;
    }
    const int q_begin = (!legacy && rc == 0) ? q_tail : 0;
    if (f16) {
        if (head_dim <= 64) return launch_att<64, true>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
        if (head_dim <= 80) return launch_att<80, true>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
        return launch_att<96, true>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
    }
    if (head_dim <= 64) return launch_att<64, false>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
    if (head_dim <= 80) return launch_att<80, false>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
    return launch_att<96, false>(qkv, batch, tokens, heads, head_dim, out, stream, q_begin);
}

int launch_attention_cls(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out, int ld_out,
                         cudaStream_t stream) {
    VLMFW_REQUIRE(qkv && out, "attention_cls: null pointer");
    VLMFW_REQUIRE(batch > 0 && tokens > 0 && heads > 0, "attention_cls: bad shape");
    VLMFW_REQUIRE(head_dim % 8 == 0 && head_dim >= 8 && head_dim <= 128, "attention_cls: head_dim %d unsupported", head_dim);
    VLMFW_REQUIRE(batch <= 65535 && tokens <= 8192, "attention_cls: grid / sequence too large");
    VLMFW_REQUIRE(ld_out >= heads * head_dim && ld_out % 2 == 0, "attention_cls: ld_out");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const size_t smem = (128 + 16 + CLS_HW * 128 + static_cast<size_t>(tokens)) * sizeof(float);
    const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
    dim3 grid(heads, batch);
    if (f16)
        attention_cls_kernel<true><<<grid, CLS_THREADS, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out),
                                                                       tokens, heads, head_dim, ld_out, scale_log2);
    else
        attention_cls_kernel<false><<<grid, CLS_THREADS, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out),
                                                                        tokens, heads, head_dim, ld_out, scale_log2);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

}  // namespace vlmfw
