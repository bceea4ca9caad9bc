// Shared host-side declarations of the kernel launchers (one translation unit per kernel family).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/vlmfw_b200.h"

namespace vlmfw {

// thread-local error message behind vlmfw_last_error()
void set_error(const char* fmt, ...);
const char* get_error();
int64_t& launch_counter();  // counts kernels launched by this library on the calling thread

#define VLMFW_CUDA_OK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::vlmfw::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                        \
            return VLMFW_E_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define VLMFW_REQUIRE(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            ::vlmfw::set_error(__VA_ARGS__); \
            return VLMFW_E_INVALID;       \
        }                                 \
    } while (0)

#define VLMFW_LAUNCH_CHECK()                   \
    do {                                       \
        ++::vlmfw::launch_counter();           \
        VLMFW_CUDA_OK(cudaGetLastError());     \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: one process may drive several
// GPUs (TowerEngine(device=...)), so the opt-in is remembered per (kernel instantiation, device), not per process.
// `done` is a bit mask over device ordinals (one static per kernel instantiation at the call site).
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kern, int bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (static_cast<unsigned>(dev) & 63u);
    if (dev < 64 && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// Bottleneck-experiment switches inside the hot kernels (VLMFW_GEMM_DBG / VLMFW_ATTN_DBG bit masks) exist only in
// builds compiled with -DVLMFW_DEBUG_MASKS (tools/ variants); the shipped library folds them to `false`.
#ifdef VLMFW_DEBUG_MASKS
#define VLMFW_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define VLMFW_DBG(p, bit) false
#endif

struct GemmArgs {
    const void* a = nullptr;  // [M,K] bf16
    int lda = 0;
    const void* w = nullptr;  // [N,K] bf16
    int ldw = 0;
    int m = 0, n = 0, k = 0;
    const float* bias = nullptr;  // [N] or null
    int epi = VLMFW_EPI_BIAS_BF16;
    int act = VLMFW_ACT_NONE;
    void* out = nullptr;
    int ldo = 0;
    const float* resid = nullptr;
    int ldr = 0;
    int resid_mod = 0;
    int grp = 0, grp_out = 0, grp_off = 0;
    int block_n = 0;  // 0 = pick automatically
    // LayerNorm folded into this GEMM (16-bit epilogue): A is the plain 16-bit cast of the fp32 rows, W carries gamma
    // and the mean subtraction (centred rows), bias carries beta W^T + b; ln_stats = [M][ln_np] float2 (sum, sum of
    // squares) partials the epilogue turns into rstd
    const void* ln_stats = nullptr;
    int ln_np = 0, ln_dim = 0;
    float ln_eps = 0.f;
    int ln_row_mul = 1;  // GEMM row m reads ln_stats row m * ln_row_mul (class-token rows of a [B*S] statistics table: S)
    // residual epilogue extras: 16-bit copy of the updated rows and their (sum, sum of squares) partials
    // ([M][gemm_stats_parts(n, block_n)] float2) for the LayerNorm folded into the NEXT GEMM
    void* h16_out = nullptr;
    int ld_h16 = 0;
    void* stats_out = nullptr;
    bool f16 = false;  // 16-bit format of A, W and the 16-bit output: false = bf16, true = fp16 (A and W must agree)
    // the accumulator is multiplied by acc_scale before bias / residual / LayerNorm scaling: W holds W_true / acc_scale
    // when W_true does not fit the fp16 range (tower.cu picks a power of two)
    float acc_scale = 1.f;
    // fp16 outputs only: device counter of (row, chunk) stores that clamped at +-65504 (NULL = not counted)
    unsigned long long* sat_count = nullptr;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda). dims[0] / box[0] = innermost.
// strides_bytes has rank-1 entries (dims 1..rank-1). swizzle_bytes in {0, 32, 64, 128}.
int encode_tensor_map(CUtensorMap* out, bool f16, int rank, const void* ptr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int device_check();
// VLMFW_PDL=0 switches programmatic dependent launch off (A/B timing)
bool pdl_enabled();
int num_sms();
int launch_gemm(const GemmArgs& g, cudaStream_t stream);
// number of (sum, sum of squares) partials per row the residual epilogue writes for an N-column output
int gemm_stats_parts(int n, int block_n);
int launch_layernorm(const float* x, int rows, int dim, const float* gamma, const float* beta, float eps, void* out,
                     int out_dtype, cudaStream_t stream);
// fp32 rows -> plain 16-bit cast + (sum, sum of squares) in partial slot 0 of stats [rows][np] float2 (others zeroed)
int launch_row_stats_cast(const float* x, int rows, int dim, void* out16, bool f16, void* stats, int np,
                          cudaStream_t stream, unsigned long long* sat = nullptr);
// W'[n,k] = round16(gamma[k] W[n,k] - mean_j(gamma[j] W[n,j])) (row stride ld), c2[n] = sum_k beta[k] W[n,k] + bias[n]
// inv_scale multiplies the folded values before rounding (1 / acc_scale of the consuming GEMM); amax_bits (device, may be
// NULL) receives max |unscaled folded value| as a float bit pattern through atomicMax (zero it first)
int launch_ln_fold_weight(const float* w, int n, int k, const float* gamma, const float* beta, const float* bias,
                          void* w16, int ld, bool f16, float* c2, cudaStream_t stream, float inv_scale = 1.f,
                          unsigned int* amax_bits = nullptr);
int launch_attention(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                     cudaStream_t stream);
// tcgen05 attention (attention_tc.cu): 0 = launched, 1 = shape not covered (use launch_attention's generic kernel), <0 = error
// Only query rows [0, q_rows) are computed (q_rows a multiple of 128, or == tokens); all `tokens` keys are attended.
int launch_attention_tc(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                        int q_rows, cudaStream_t stream);
// second-generation tcgen05 attention (attention_wb.cu): 0 = launched, 1 = shape not covered, <0 = error
int launch_attention_wb(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out,
                        cudaStream_t stream);
// attention of ONE query row per image (the class token, row 0) over all keys: out[b, head*dh ...] of a compact [B, D]
// matrix with row stride ld_out elements (last layer of a class-token-pooled tower, tower.cu cls tail)
int launch_attention_cls(const void* qkv, bool f16, int batch, int tokens, int heads, int head_dim, void* out, int ld_out,
                         cudaStream_t stream);
int launch_im2col(const void* pixels, int pix_fmt, int batch, int image_size, int patch, const float* mean3,
                  const float* std3, void* out, int kpad, bool out_f16, cudaStream_t stream);
int launch_fill_cls(const float* cls_pos0, float* h, int batch, int tokens, int dim, cudaStream_t stream);
int launch_pool(const float* h, int batch, int tokens, int dim, int pool, const float* gamma, const float* beta,
                float eps, float* pooled, cudaStream_t stream);
// same, reading a 16-bit (VLMFW_DT_F16 / VLMFW_DT_BF16) or fp32 hidden state
int launch_pool_dt(const void* h, int in_dt, int batch, int tokens, int dim, int pool, const float* gamma,
                   const float* beta, float eps, float* pooled, cudaStream_t stream);
int launch_pack_16(const float* src, int64_t rows, int64_t cols, void* dst, int64_t dst_ld, bool f16, cudaStream_t stream,
                   float inv_scale = 1.f);
// power-of-two scale s >= 1 with amax / s <= 32768: what a weight matrix is divided by so that it fits fp16
float fp16_weight_scale(float amax);
// max |x| over a device fp32 array, returned to the host (synchronises `stream`; load-time only)
int device_absmax(const float* src, int64_t numel, float* host_out, cudaStream_t stream);
int launch_add_rows(const float* a, const float* b, float* out, int64_t rows, int64_t cols, int64_t a_row0,
                    cudaStream_t stream);

}  // namespace vlmfw
