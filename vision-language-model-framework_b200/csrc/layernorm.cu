// LayerNorm and the pooling epilogues of the three adapters. All HBM-bound: one pass, 128-bit coalesced accesses,
// fp32 statistics (two-pass in registers, like torch's native_layer_norm), warp-shuffle reductions.
//
//   layernorm_kernel    : hf modeling_siglip.py:348,357 / modeling_clip.py:371,380 (+pre_layrnorm :677) /
//                         modeling_blip_2.py:389,398 — reads the fp32 residual stream, writes the bf16 GEMM operand
//   pool_ln_mean_kernel : SigLIP post_layernorm (modeling_siglip.py:618) fused with the adapter's mean(dim=1)
//                         (reference models/paligemma.py:111-112)
//   pool_mean_kernel    : LLaVA adapter mean / CLS over the raw last_hidden_state (reference models/llava.py:114-117)
//   pool_cls_ln2_kernel : BLIP-2 pooler_output = LN(LN(h)[:,0]) (modeling_blip_2.py:523-526, reference blip2.py:94-95)
#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int LN_MAX_CHUNKS = 12;  // float4 chunks per lane: dim <= 12*32*4 = 1536

// One warp per row. dim % 4 == 0.
template <int OUT_DT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows, int dim,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, void* out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nchunks = dim >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * dim);
    float4 v[LN_MAX_CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) {
            v[i] = xr[c];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / static_cast<float>(dim);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) {
            const float4 g = __ldg(g4 + c);
            const float4 b = __ldg(b4 + c);
            float4 y;
            y.x = (v[i].x - mean) * rstd * g.x + b.x;
            y.y = (v[i].y - mean) * rstd * g.y + b.y;
            y.z = (v[i].z - mean) * rstd * g.z + b.z;
            y.w = (v[i].w - mean) * rstd * g.w + b.w;
            if constexpr (OUT_DT == VLMFW_DT_F32) {
                reinterpret_cast<float4*>(static_cast<float*>(out) + static_cast<size_t>(warp) * dim)[c] = y;
            } else {
                uint2 o;
                o.x = pack_16x2<OUT_DT == VLMFW_DT_F16>(y.x, y.y);
                o.y = pack_16x2<OUT_DT == VLMFW_DT_F16>(y.z, y.w);
                reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + static_cast<size_t>(warp) * dim)[c] = o;
            }
        }
    }
}

// Input element type of the pooling kernels: the fp32 residual stream, or the 16-bit one (fp16 / bf16 towers that keep the
// stream in the tensor-core format). ld4: elements [4c, 4c + 4) of a row; ld1: one element.
template <int DT>
__device__ __forceinline__ float4 ld4(const void* row, int c) {
    if constexpr (DT == VLMFW_DT_F32) {
        return reinterpret_cast<const float4*>(row)[c];
    } else {
        const uint2 u = reinterpret_cast<const uint2*>(row)[c];
        float4 f;
        if constexpr (DT == VLMFW_DT_F16) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            f = make_float4(a.x, a.y, b.x, b.y);
        } else {
            f = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                            __uint_as_float(u.y & 0xFFFF0000u));
        }
        return f;
    }
}
template <int DT>
__device__ __forceinline__ float ld1(const void* base, size_t i) {
    if constexpr (DT == VLMFW_DT_F32) {
        return static_cast<const float*>(base)[i];
    } else if constexpr (DT == VLMFW_DT_F16) {
        return __half2float(static_cast<const __half*>(base)[i]);
    } else {
        return __uint_as_float(static_cast<uint32_t>(static_cast<const uint16_t*>(base)[i]) << 16);
    }
}
template <int DT>
__device__ __forceinline__ const void* row_ptr(const void* base, size_t elems) {
    return static_cast<const uint8_t*>(base) + elems * (DT == VLMFW_DT_F32 ? 4 : 2);
}

// SigLIP: pooled[b, d] = mean_s ( (h[b,s,d] - mu_s) * rstd_s ) * gamma[d] + beta[d].
// One CTA (POOL_WARPS warps) per image, ONE pass: a warp owns a token at a time — row into registers, two-pass statistics in
// registers, the normalised row added to the warp's own per-lane column accumulators — and the warps' accumulators meet in
// shared memory, added in warp order (fixed summation order). Every row is read once, by one warp, straight from HBM. (The
// previous version wrote per-token (mu, rstd) to shared memory and re-read the image column by column: served by L2 for the
// 16-bit stream, but 2x the algorithmic DRAM bytes for the fp32 one in the round-1 ncu capture, and 94 us for 151 MB.)
constexpr int POOL_WARPS = 16;
template <int DT>
__global__ void __launch_bounds__(32 * POOL_WARPS) pool_ln_mean_kernel(const void* __restrict__ h, int tokens, int dim,
                                                                       const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, float eps,
                                                                       float* __restrict__ pooled) {
    extern __shared__ float4 pool_part[];  // [POOL_WARPS][dim / 4]
    const int b = blockIdx.x;
    const void* hb = row_ptr<DT>(h, static_cast<size_t>(b) * tokens * dim);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = dim >> 2;
    float4 acc[LN_MAX_CHUNKS];
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = warp; s < tokens; s += POOL_WARPS) {
        const void* xr = row_ptr<DT>(hb, static_cast<size_t>(s) * dim);
        float4 v[LN_MAX_CHUNKS];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) {
                v[i] = ld4<DT>(xr, c);
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(sum) / static_cast<float>(dim);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) {
                const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + bb * bb) + (cc * cc + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + eps);
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) {
                acc[i].x += (v[i].x - mean) * rstd;
                acc[i].y += (v[i].y - mean) * rstd;
                acc[i].z += (v[i].z - mean) * rstd;
                acc[i].w += (v[i].w - mean) * rstd;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) pool_part[warp * nchunks + c] = acc[i];
    }
    __syncthreads();
    const float inv_tokens = 1.f / static_cast<float>(tokens);
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
        float4 t = pool_part[c];
#pragma unroll
        for (int w = 1; w < POOL_WARPS; ++w) {
            const float4 u = pool_part[w * nchunks + c];
            t.x += u.x;
            t.y += u.y;
            t.z += u.z;
            t.w += u.w;
        }
        const float4 g = reinterpret_cast<const float4*>(gamma)[c];
        const float4 bt = reinterpret_cast<const float4*>(beta)[c];
        float4 o;
        o.x = t.x * inv_tokens * g.x + bt.x;
        o.y = t.y * inv_tokens * g.y + bt.y;
        o.z = t.z * inv_tokens * g.z + bt.z;
        o.w = t.w * inv_tokens * g.w + bt.w;
        reinterpret_cast<float4*>(pooled + static_cast<size_t>(b) * dim)[c] = o;
    }
}

// CLIP adapter: mean over all tokens (first = 0, count = tokens) or the CLS row (count = 1) of the raw hidden state.
template <int DT>
__global__ void __launch_bounds__(256) pool_mean_kernel(const void* __restrict__ h, int tokens, int dim, int count,
                                                        float* __restrict__ pooled) {
    const int b = blockIdx.x;
    const int d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    const void* hb = row_ptr<DT>(h, static_cast<size_t>(b) * tokens * dim);
    float acc = 0.f;
    for (int s = 0; s < count; ++s) acc += ld1<DT>(hb, static_cast<size_t>(s) * dim + d);
    pooled[static_cast<size_t>(b) * dim + d] = acc / static_cast<float>(count);
}

// BLIP-2: y = LN(h[b,0,:]); pooled = LN(y) with the same gamma/beta. One warp per image.
template <int DT>
__global__ void __launch_bounds__(128) pool_cls_ln2_kernel(const void* __restrict__ h, int batch, int tokens, int dim,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           float* __restrict__ pooled) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= batch) return;
    const int nchunks = dim >> 2;
    const void* xr = row_ptr<DT>(h, static_cast<size_t>(b) * tokens * dim);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    float4 v[LN_MAX_CHUNKS];
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) v[i] = ld4<DT>(xr, c);
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) / static_cast<float>(dim);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) {
                const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + bb * bb) + (cc * cc + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + eps);
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = i * 32 + lane;
            if (c < nchunks) {
                const float4 g = __ldg(g4 + c);
                const float4 bt = __ldg(b4 + c);
                v[i].x = (v[i].x - mean) * rstd * g.x + bt.x;
                v[i].y = (v[i].y - mean) * rstd * g.y + bt.y;
                v[i].z = (v[i].z - mean) * rstd * g.z + bt.z;
                v[i].w = (v[i].w - mean) * rstd * g.w + bt.w;
            }
        }
    }
    float4* o = reinterpret_cast<float4*>(pooled + static_cast<size_t>(b) * dim);
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) o[c] = v[i];
    }
}

// Entry of the folded-LayerNorm chain: the first LayerNorm of the tower has no GEMM epilogue in front
// of it, so this kernel does what the residual epilogue does for every later one — a plain 16-bit cast of the fp32
// rows (the next GEMM's A operand) and the row's (sum, sum of squares) in partial slot 0, the other slots zeroed.
template <bool F16>
__global__ void __launch_bounds__(256) row_stats_cast_kernel(const float* __restrict__ x, int rows, int dim,
                                                             void* __restrict__ out16, float2* __restrict__ stats,
                                                             int np, unsigned long long* __restrict__ sat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nchunks = dim >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * dim);
    uint2* orow = reinterpret_cast<uint2*>(static_cast<uint16_t*>(out16) + static_cast<size_t>(warp) * dim);
    float s1 = 0.f, s2 = 0.f, amax = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) {
            const float4 v = xr[c];
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
            uint2 o;
            o.x = pack_16x2<F16>(v.x, v.y);
            o.y = pack_16x2<F16>(v.z, v.w);
            orow[c] = o;
        }
    }
    if (F16 && sat != nullptr && amax > 65504.f) atomicAdd(sat, 1ull);  // the fp16 cast above clamped (satfinite)
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane < np) stats[static_cast<size_t>(warp) * np + lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
}

// LayerNorm folded into the weights of the GEMM that consumes it (one block per output row n). With P = I - 11^T/K
// (the mean subtraction, a projection): LN(x) W^T + b = rstd (x P) (gamma . W)^T + (beta W^T + b)
//                                                     = rstd x (P (gamma . W)^T) + c2, so
//   W'[n,k] = round16(gamma[k] W[n,k] - mean_j(gamma[j] W[n,j]))   (rows centred: the tensor core subtracts the mean)
//   c2[n]   = sum_k beta[k] W[n,k] + b[n]
// and the consuming epilogue is a single FMA per element: act(rstd[m] acc + c2[n]).
template <bool F16>
__global__ void __launch_bounds__(256) ln_fold_weight_kernel(const float* __restrict__ w, int k, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ bias,
                                                             uint16_t* __restrict__ w16, int ld, float* __restrict__ c2,
                                                             float inv_scale, unsigned int* __restrict__ amax_bits) {
    const int n = blockIdx.x;
    const float* wr = w + static_cast<size_t>(n) * k;
    uint16_t* orow = w16 + static_cast<size_t>(n) * ld;
    double a1 = 0.0, a2 = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double wv = static_cast<double>(wr[i]);
        a1 += wv * static_cast<double>(gamma[i]);
        a2 += wv * static_cast<double>(beta[i]);
    }
    __shared__ double r1[8], r2[8];
    __shared__ float row_mean;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        r1[threadIdx.x >> 5] = a1;
        r2[threadIdx.x >> 5] = a2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < 8; ++i) {
            t1 += r1[i];
            t2 += r2[i];
        }
        row_mean = static_cast<float>(t1 / static_cast<double>(k));
        c2[n] = static_cast<float>(t2 + (bias != nullptr ? static_cast<double>(bias[n]) : 0.0));
    }
    __syncthreads();
    const float mean = row_mean;
    unsigned int am = 0;  // max |folded value| of this thread as a bit pattern (orders like the float; NaN / Inf on top)
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        uint16_t bits = 0;
        if (i < k) {
            const float wf = fmaf(wr[i], gamma[i], -mean);
            am = max(am, __float_as_uint(wf) & 0x7FFFFFFFu);
            const float ws = wf * inv_scale;  // power of two (exact) or 1
            if constexpr (F16) bits = __half_as_ushort(__float2half_rn(ws));
            else bits = __bfloat16_as_ushort(__float2bfloat16_rn(ws));
        }
        orow[i] = bits;
    }
    if (amax_bits != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) am = max(am, __shfl_xor_sync(0xffffffffu, am, o));
        if ((threadIdx.x & 31) == 0 && am != 0) atomicMax(amax_bits, am);
    }
}

}  // namespace

int launch_row_stats_cast(const float* x, int rows, int dim, void* out16, bool f16, void* stats, int np,
                          cudaStream_t stream, unsigned long long* sat) {
    VLMFW_REQUIRE(x && out16 && stats, "row_stats_cast: null pointer");
    VLMFW_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_CHUNKS * 128, "row_stats_cast: bad dim %d", dim);
    VLMFW_REQUIRE(np >= 1 && np <= 32, "row_stats_cast: %d partial slots (1..32)", np);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int grid = (rows + 7) / 8;
    if (f16) row_stats_cast_kernel<true><<<grid, 256, 0, stream>>>(x, rows, dim, out16, static_cast<float2*>(stats), np, sat);
    else row_stats_cast_kernel<false><<<grid, 256, 0, stream>>>(x, rows, dim, out16, static_cast<float2*>(stats), np, sat);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

int launch_ln_fold_weight(const float* w, int n, int k, const float* gamma, const float* beta, const float* bias,
                          void* w16, int ld, bool f16, float* c2, cudaStream_t stream, float inv_scale,
                          unsigned int* amax_bits) {
    VLMFW_REQUIRE(w && gamma && beta && w16 && c2, "ln_fold_weight: null pointer");
    VLMFW_REQUIRE(n > 0 && k > 0 && ld >= k, "ln_fold_weight: bad shape n=%d k=%d ld=%d", n, k, ld);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    if (f16)
        ln_fold_weight_kernel<true><<<n, 256, 0, stream>>>(w, k, gamma, beta, bias, static_cast<uint16_t*>(w16), ld, c2,
                                                           inv_scale, amax_bits);
    else
        ln_fold_weight_kernel<false><<<n, 256, 0, stream>>>(w, k, gamma, beta, bias, static_cast<uint16_t*>(w16), ld, c2,
                                                            inv_scale, amax_bits);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

int launch_layernorm(const float* x, int rows, int dim, const float* gamma, const float* beta, float eps, void* out,
                     int out_dtype, cudaStream_t stream) {
    VLMFW_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
    VLMFW_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_CHUNKS * 128, "layernorm: bad dim %d", dim);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int warps_per_block = 8;
    const int grid = (rows + warps_per_block - 1) / warps_per_block;
    if (out_dtype == VLMFW_DT_F32)
        layernorm_kernel<VLMFW_DT_F32><<<grid, warps_per_block * 32, 0, stream>>>(x, rows, dim, gamma, beta, eps, out);
    else if (out_dtype == VLMFW_DT_F16)
        layernorm_kernel<VLMFW_DT_F16><<<grid, warps_per_block * 32, 0, stream>>>(x, rows, dim, gamma, beta, eps, out);
    else if (out_dtype == VLMFW_DT_BF16)
        layernorm_kernel<VLMFW_DT_BF16><<<grid, warps_per_block * 32, 0, stream>>>(x, rows, dim, gamma, beta, eps, out);
    else {
        set_error("layernorm: unknown out_dtype %d", out_dtype);
        return VLMFW_E_INVALID;
    }
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

template <int DT>
static int launch_pool_t(const void* h, int batch, int tokens, int dim, int pool, const float* gamma, const float* beta,
                         float eps, float* pooled, cudaStream_t stream) {
    const int threads = 256;
    dim3 grid(batch, (dim + threads - 1) / threads);
    switch (pool) {
        case VLMFW_POOL_LN_MEAN:
            VLMFW_REQUIRE(gamma && beta, "pool: LN_MEAN needs gamma/beta");
        {
            VLMFW_REQUIRE(dim % 4 == 0 && dim <= LN_MAX_CHUNKS * 128, "pool: LN_MEAN needs dim %% 4 == 0 and dim <= %d", LN_MAX_CHUNKS * 128);
            const size_t smem = static_cast<size_t>(POOL_WARPS) * dim * sizeof(float);
            static std::atomic<uint64_t> opt_in{0};  // per instantiation, one bit per device
            VLMFW_CUDA_OK(ensure_dynamic_smem(pool_ln_mean_kernel<DT>, POOL_WARPS * LN_MAX_CHUNKS * 128 * 4, opt_in));
            pool_ln_mean_kernel<DT><<<batch, 32 * POOL_WARPS, smem, stream>>>(h, tokens, dim, gamma, beta, eps, pooled);
            break;
        }
        case VLMFW_POOL_MEAN:
            pool_mean_kernel<DT><<<grid, threads, 0, stream>>>(h, tokens, dim, tokens, pooled);
            break;
        case VLMFW_POOL_CLS:
            pool_mean_kernel<DT><<<grid, threads, 0, stream>>>(h, tokens, dim, 1, pooled);
            break;
        case VLMFW_POOL_LN_CLS_LN:
            VLMFW_REQUIRE(gamma && beta, "pool: LN_CLS_LN needs gamma/beta");
            pool_cls_ln2_kernel<DT><<<(batch + 3) / 4, 128, 0, stream>>>(h, batch, tokens, dim, gamma, beta, eps, pooled);
            break;
        default:
            set_error("pool: unknown mode %d", pool);
            return VLMFW_E_INVALID;
    }
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

int launch_pool_dt(const void* h, int in_dt, int batch, int tokens, int dim, int pool, const float* gamma,
                   const float* beta, float eps, float* pooled, cudaStream_t stream) {
    VLMFW_REQUIRE(h && pooled, "pool: null pointer");
    VLMFW_REQUIRE(batch > 0 && tokens > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_CHUNKS * 128,
                  "pool: bad shape batch=%d tokens=%d dim=%d", batch, tokens, dim);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    if (in_dt == VLMFW_DT_F16) return launch_pool_t<VLMFW_DT_F16>(h, batch, tokens, dim, pool, gamma, beta, eps, pooled, stream);
    if (in_dt == VLMFW_DT_BF16) return launch_pool_t<VLMFW_DT_BF16>(h, batch, tokens, dim, pool, gamma, beta, eps, pooled, stream);
    return launch_pool_t<VLMFW_DT_F32>(h, batch, tokens, dim, pool, gamma, beta, eps, pooled, stream);
}

int launch_pool(const float* h, int batch, int tokens, int dim, int pool, const float* gamma, const float* beta,
                float eps, float* pooled, cudaStream_t stream) {
    return launch_pool_dt(h, VLMFW_DT_F32, batch, tokens, dim, pool, gamma, beta, eps, pooled, stream);
}

}  // namespace vlmfw
