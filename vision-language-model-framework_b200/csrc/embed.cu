// Patch-embedding front end: pixel gather (conv-as-GEMM operand), class-token rows, weight repacking.
//
//   im2col_kernel : hf modeling_siglip.py:124-130,178 / modeling_clip.py:146-158 / modeling_blip_2.py:192-201 —
//                   Conv2d(3->D, k=p, s=p) is a GEMM over non-overlapping patches: A[b*P + py*G + px, c*p*p + ky*p + kx].
//                   The image-processor arithmetic (x/255 - mean)/std (hf image_processing_* rescale + normalize) is
//                   fused here when the input is raw uint8.
//   fill_cls      : class_embedding + position_embedding[0] written to token 0 of every image
//                   (modeling_clip.py:208-217, modeling_blip_2.py:247-255)
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

struct PixNorm {
    float scale[3];  // 1 / (255 * std)
    float shift[3];  // -mean / std
};

// One thread produces 8 consecutive k of one patch row (16-byte store).
template <int PIX_FMT, bool OUT_F16>
__global__ void __launch_bounds__(256) im2col_kernel(const void* __restrict__ pixels, int batch, int img, int patch,
                                                     PixNorm nrm, uint16_t* __restrict__ out, int kpad) {
    const int grid_side = img / patch;
    const int patches = grid_side * grid_side;
    const int kreal = 3 * patch * patch;
    const int groups = kpad >> 3;
    const size_t total = static_cast<size_t>(batch) * patches * groups;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups);
        const size_t row = idx / groups;
        const int pidx = static_cast<int>(row % patches);
        const int b = static_cast<int>(row / patches);
        const int py = pidx / grid_side, px = pidx % grid_side;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;
            float val = 0.f;
            if (k < kreal) {
                const int c = k / (patch * patch);
                const int rem = k - c * patch * patch;
                const int ky = rem / patch, kx = rem - ky * patch;
                const int y = py * patch + ky, x = px * patch + kx;
                if constexpr (PIX_FMT == VLMFW_PIX_U8_NHWC) {
                    const uint8_t u =
                        static_cast<const uint8_t*>(pixels)[((static_cast<size_t>(b) * img + y) * img + x) * 3 + c];
                    // same op order as the HF processor: rescale (x * 1/255) then normalize ((x - mean) / std)
                    val = static_cast<float>(u) * nrm.scale[c] + nrm.shift[c];
                } else {
                    val = static_cast<const float*>(pixels)[((static_cast<size_t>(b) * 3 + c) * img + y) * img + x];
                }
            }
            f[j] = val;
        }
        uint4 o;
        o.x = pack_16x2<OUT_F16>(f[0], f[1]);
        o.y = pack_16x2<OUT_F16>(f[2], f[3]);
        o.z = pack_16x2<OUT_F16>(f[4], f[5]);
        o.w = pack_16x2<OUT_F16>(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + row * kpad + g * 8) = o;
    }
}

// uint8 NHWC input, one CTA per (image, patch row): the strip of `patch` image rows (patch * img * 3 contiguous bytes) is
// staged in shared memory with 16-byte loads, a table maps every k = c p^2 + ky p + kx to its byte offset inside the strip
// (built once per CTA: the per-element divisions of the generic kernel are gone), and a thread then produces 8 consecutive
// k of one patch from shared-memory bytes and stores 16 bytes; consecutive threads write consecutive chunks of an output
// row. The generic kernel above reads single bytes from global memory with three integer divisions per element:
// 149 us per 256 SigLIP images (0.5 TB/s) against the ~25 us the 116 MB moved here cost at the HBM rate.
template <bool OUT_F16>
__global__ void __launch_bounds__(256) im2col_u8_strip_kernel(const uint8_t* __restrict__ pixels, int img, int patch,
                                                              PixNorm nrm, uint16_t* __restrict__ out, int kpad) {
    extern __shared__ __align__(16) uint8_t strip_smem[];
    const int grid_side = img / patch;
    const int kreal = 3 * patch * patch;
    const int row_bytes = img * 3;
    const int strip_bytes = patch * row_bytes;            // a multiple of 16 (checked by the launcher)
    // [kpad]: offset of k inside the strip for px = 0 (low 14 bits) | channel << 14 (0xFFFF: padding column)
    uint16_t* tab = reinterpret_cast<uint16_t*>(strip_smem + strip_bytes);
    const int b = blockIdx.x / grid_side, py = blockIdx.x % grid_side;
    const uint4* src = reinterpret_cast<const uint4*>(pixels + (static_cast<size_t>(b) * img + static_cast<size_t>(py) * patch) * row_bytes);
    uint4* dst4 = reinterpret_cast<uint4*>(strip_smem);
    for (int i = threadIdx.x; i < strip_bytes / 16; i += blockDim.x) dst4[i] = __ldg(src + i);
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        int off = 0xFFFF;  // padding column: zero
        if (k < kreal) {
            const int c = k / (patch * patch);
            const int rem = k - c * patch * patch;
            const int ky = rem / patch, kx = rem - ky * patch;
            off = (ky * row_bytes + kx * 3 + c) | (c << 14);
        }
        tab[k] = static_cast<uint16_t>(off);
    }
    __syncthreads();
    const int groups = kpad >> 3;
    const size_t row0 = (static_cast<size_t>(b) * grid_side + py) * grid_side;  // first patch of this strip
    // per-channel constants in registers, picked by selects: indexing the kernel-parameter arrays with a per-lane channel
    // is a divergent constant load, and those saturated the address-divergence unit (ncu: ADU pipe 97 % busy, 120 us)
    const float s0 = nrm.scale[0], s1 = nrm.scale[1], s2 = nrm.scale[2];
    const float h0 = nrm.shift[0], h1 = nrm.shift[1], h2 = nrm.shift[2];
    for (int idx = threadIdx.x; idx < grid_side * groups; idx += blockDim.x) {
        const int px = idx / groups, g = idx - px * groups;
        const int base = px * patch * 3;
        const uint4 t4 = *reinterpret_cast<const uint4*>(tab + g * 8);
        const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t e = (tw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            float val = 0.f;
            if (e != 0xFFFFu) {
                const uint32_t c = e >> 14, off = e & 0x3FFFu;
                const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), sh = c == 0 ? h0 : (c == 1 ? h1 : h2);
                // same op order as the HF processor: rescale (x * 1/255) then normalize ((x - mean) / std)
                val = static_cast<float>(strip_smem[base + off]) * sc + sh;
            }
            f[j] = val;
        }
        uint4 o;
        o.x = pack_16x2<OUT_F16>(f[0], f[1]);
        o.y = pack_16x2<OUT_F16>(f[2], f[3]);
        o.z = pack_16x2<OUT_F16>(f[4], f[5]);
        o.w = pack_16x2<OUT_F16>(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + (row0 + px) * kpad + g * 8) = o;
    }
}

__global__ void fill_cls_kernel(const float* __restrict__ cls_pos0, float* __restrict__ h, int batch, int tokens,
                                int dim) {
    const int b = blockIdx.x;
    for (int d = threadIdx.x; d < dim; d += blockDim.x)
        h[static_cast<size_t>(b) * tokens * dim + d] = cls_pos0[d];
}

template <bool F16>
__global__ void pack_16_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, uint16_t* dst,
                               int64_t dst_ld, float inv_scale) {
    const int64_t total = rows * dst_ld;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / dst_ld, c = i - r * dst_ld;
        const float v = c < cols ? src[r * cols + c] * inv_scale : 0.f;  // inv_scale: power of two (exact) or 1
        if constexpr (F16) dst[i] = __half_as_ushort(__float2half_rn(v));
        else dst[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}

// max |x| via a bit-pattern atomicMax (non-negative floats order like unsigned ints); NaN/Inf map above any finite
__global__ void absmax_kernel(const float* __restrict__ src, int64_t numel, unsigned int* out) {
    unsigned int m = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = max(m, __float_as_uint(src[i]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// out[r, c] = a[r + a_row0, c] + (b ? b[c] : 0)
__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                int64_t rows, int64_t cols, int64_t a_row0) {
    const int64_t total = rows * cols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        out[i] = a[(r + a_row0) * cols + c] + (b ? b[c] : 0.f);
    }
}

int grid_for(size_t total, int threads) {
    size_t g = (total + threads - 1) / threads;
    const size_t cap = static_cast<size_t>(num_sms()) * 16;
    return static_cast<int>(g < cap ? (g == 0 ? 1 : g) : cap);
}

}  // namespace

int launch_im2col(const void* pixels, int pix_fmt, int batch, int image_size, int patch, const float* mean3,
                  const float* std3, void* out, int kpad, bool out_f16, cudaStream_t stream) {
    VLMFW_REQUIRE(pixels && out, "im2col: null pointer");
    VLMFW_REQUIRE(batch > 0 && patch > 0 && image_size % patch == 0, "im2col: image %d not a multiple of patch %d",
                  image_size, patch);
    VLMFW_REQUIRE(kpad % 8 == 0 && kpad >= 3 * patch * patch, "im2col: kpad %d too small / unaligned", kpad);
    VLMFW_REQUIRE(pix_fmt == VLMFW_PIX_U8_NHWC || pix_fmt == VLMFW_PIX_F32_NCHW, "im2col: unknown pixel format %d",
                  pix_fmt);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    PixNorm nrm;
    for (int c = 0; c < 3; ++c) {
        const float sd = std3 ? std3[c] : 1.f;
        const float mu = mean3 ? mean3[c] : 0.f;
        nrm.scale[c] = 1.0f / (255.0f * sd);
        nrm.shift[c] = -mu / sd;
    }
    const int g = image_size / patch;
    const size_t total = static_cast<size_t>(batch) * g * g * (kpad / 8);
    const int threads = 256;
    const int grid = grid_for(total, threads);
    uint16_t* o = static_cast<uint16_t*>(out);
    // uint8 input whose patch-row strip (+ the k table) fits the default 48 KB of shared memory and keeps 16-byte alignment
    // (all three towers: 14 x 672 B, 14 x 1008 B): strip kernel; anything else: the generic gather
    const size_t strip_bytes = static_cast<size_t>(patch) * image_size * 3;
    const size_t strip_smem = strip_bytes + static_cast<size_t>(kpad) * 2;
    if (pix_fmt == VLMFW_PIX_U8_NHWC && strip_bytes % 16 == 0 && strip_bytes < 0x4000 && strip_smem <= 48 * 1024 &&
        (reinterpret_cast<uintptr_t>(pixels) & 15) == 0 && (static_cast<size_t>(image_size) * image_size * 3) % 16 == 0) {
        const int sgrid = batch * g;
        if (out_f16) im2col_u8_strip_kernel<true><<<sgrid, threads, strip_smem, stream>>>(static_cast<const uint8_t*>(pixels), image_size, patch, nrm, o, kpad);
        else im2col_u8_strip_kernel<false><<<sgrid, threads, strip_smem, stream>>>(static_cast<const uint8_t*>(pixels), image_size, patch, nrm, o, kpad);
        VLMFW_LAUNCH_CHECK();
        return VLMFW_OK;
    }
    if (pix_fmt == VLMFW_PIX_U8_NHWC) {
        if (out_f16) im2col_kernel<VLMFW_PIX_U8_NHWC, true><<<grid, threads, 0, stream>>>(pixels, batch, image_size, patch, nrm, o, kpad);
        else im2col_kernel<VLMFW_PIX_U8_NHWC, false><<<grid, threads, 0, stream>>>(pixels, batch, image_size, patch, nrm, o, kpad);
    } else {
        if (out_f16) im2col_kernel<VLMFW_PIX_F32_NCHW, true><<<grid, threads, 0, stream>>>(pixels, batch, image_size, patch, nrm, o, kpad);
        else im2col_kernel<VLMFW_PIX_F32_NCHW, false><<<grid, threads, 0, stream>>>(pixels, batch, image_size, patch, nrm, o, kpad);
    }
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

int launch_fill_cls(const float* cls_pos0, float* h, int batch, int tokens, int dim, cudaStream_t stream) {
    fill_cls_kernel<<<batch, 256, 0, stream>>>(cls_pos0, h, batch, tokens, dim);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

int launch_pack_16(const float* src, int64_t rows, int64_t cols, void* dst, int64_t dst_ld, bool f16, cudaStream_t stream,
                   float inv_scale) {
    VLMFW_REQUIRE(src && dst && rows > 0 && cols > 0 && dst_ld >= cols, "pack_16: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int grid = grid_for(static_cast<size_t>(rows * dst_ld), 256);
    if (f16) pack_16_kernel<true><<<grid, 256, 0, stream>>>(src, rows, cols, static_cast<uint16_t*>(dst), dst_ld, inv_scale);
    else pack_16_kernel<false><<<grid, 256, 0, stream>>>(src, rows, cols, static_cast<uint16_t*>(dst), dst_ld, inv_scale);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

// Power-of-two scale s >= 1 with amax / s <= 32768 (half the fp16 range: headroom for the rounding of the LayerNorm
// fold). The 16-bit operand then holds W / s exactly-scaled and the GEMM epilogue multiplies the accumulator by s.
float fp16_weight_scale(float amax) {
    float s = 1.f;
    while (amax / s > 32768.f && s < 1.0e30f) s *= 2.f;
    return s;
}

int device_absmax(const float* src, int64_t numel, float* host_out, cudaStream_t stream) {
    VLMFW_REQUIRE(src && host_out && numel > 0, "absmax: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    unsigned int* d = nullptr;
    VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(unsigned int)));
    VLMFW_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(unsigned int), stream));
    absmax_kernel<<<grid_for(static_cast<size_t>(numel), 256), 256, 0, stream>>>(src, numel, d);
    ++launch_counter();
    unsigned int bits = 0;
    cudaError_t e = cudaMemcpyAsync(&bits, d, sizeof(bits), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    VLMFW_CUDA_OK(e);
    memcpy(host_out, &bits, sizeof(float));
    return VLMFW_OK;
}

int launch_add_rows(const float* a, const float* b, float* out, int64_t rows, int64_t cols, int64_t a_row0,
                    cudaStream_t stream) {
    add_rows_kernel<<<grid_for(static_cast<size_t>(rows * cols), 256), 256, 0, stream>>>(a, b, out, rows, cols, a_row0);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

}  // namespace vlmfw
