// Probe train step for the deeper head: reference probing/models/base_probe.py:14-22
//   make_head_deeper = BatchNorm1d(D) -> Dropout -> Linear(D, Hd) -> GELU(erf) -> Dropout -> Linear(Hd, C)
// trained by probing/train/singletask_trainer.py:183-191 / multitask_trainer.py:188-207 (weighted CE over the rows whose
// label is present) + base_trainer.py:166-181 (backward, AdamW). Same phase contract as the linear head in probe.cu:
// BatchNorm uses the GLOBAL batch statistics handed in, every output is an un-normalised sum over the local rows in a
// fixed order (no atomics), so shards add up to the single-process step.
//   bucket = [dW1 (Hd*D) | db1 (Hd) | dW2 (C*Hd) | db2 (C) | dgamma (D) | dbeta (D) | loss_sum | weight_sum]
//   params = [ W1        |  b1      |  W2        |  b2     |  gamma     |  beta ]
// All fp32 on CUDA cores: B x D x Hd = 256 x 1152 x 512 is 0.15 GFLOP per contraction — microseconds, latency-bound.
#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int MAX_CLASSES = 32;

// C[m,n] = sum_k A(m,k) B(k,n), fixed summation order over k. A(m,k) = TA ? a[k*lda + m] : a[m*lda + k];
// B(k,n) = TB ? b[n*ldb + k] : b[k*ldb + n]. 64 x 64 tile, 16-deep k slices, 256 threads, 4 x 4 outputs per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                    int ldb, float* __restrict__ c, int ldc, int m, int n, int k) {
    __shared__ float sa[16][64 + 4];
    __shared__ float sb[16][64 + 4];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < k; k0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            int kk, mm;
            if (TA) {
                kk = i >> 6;
                mm = i & 63;
            } else {
                mm = i >> 4;
                kk = i & 15;
            }
            const int gm = m0 + mm, gk = k0 + kk;
            sa[kk][mm] = (gm < m && gk < k) ? (TA ? a[static_cast<size_t>(gk) * lda + gm] : a[static_cast<size_t>(gm) * lda + gk]) : 0.f;
            int kb, nn;
            if (TB) {
                nn = i >> 4;
                kb = i & 15;
            } else {
                kb = i >> 6;
                nn = i & 63;
            }
            const int gn = n0 + nn, gkb = k0 + kb;
            sb[kb][nn] = (gn < n && gkb < k) ? (TB ? b[static_cast<size_t>(gn) * ldb + gkb] : b[static_cast<size_t>(gkb) * ldb + gn]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = sa[kk][ty * 4 + i];
                bv[i] = sb[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < m && gn < n) c[static_cast<size_t>(gm) * ldc + gn] = acc[i][j];
        }
}

template <bool TA, bool TB>
int sgemm(const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k, cudaStream_t s) {
    dim3 grid((n + 63) / 64, (m + 63) / 64);
    sgemm_kernel<TA, TB><<<grid, 256, 0, s>>>(a, lda, b, ldb, c, ldc, m, n, k);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

__device__ __forceinline__ void bn_mean_rstd(const double* stats, int dim, int d, int n, float eps, float& mean,
                                             float& rstd) {
    const double mu = stats[d] / n;
    double var = stats[dim + d] / n - mu * mu;  // biased variance, as BatchNorm normalises with in train mode
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(mu);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// z1 = dropout1(BN_train(x)), xh = (x - mean) * rstd kept for the BatchNorm backward
__global__ void __launch_bounds__(256) bn_dropout_kernel(const float* __restrict__ x, int batch, int dim,
                                                         const double* __restrict__ stats, int n_global, float eps,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const uint8_t* __restrict__ keep, float keep_scale,
                                                         float* __restrict__ z1, float* __restrict__ xh) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    float mean, rstd;
    bn_mean_rstd(stats, dim, d, n_global, eps, mean, rstd);
    const float g = gamma[d], bt = beta[d];
    for (int b = blockIdx.y; b < batch; b += gridDim.y) {
        const size_t i = static_cast<size_t>(b) * dim + d;
        const float h = (x[i] - mean) * rstd;
        float y = h * g + bt;
        if (keep != nullptr) y = keep[i] ? y * keep_scale : 0.f;
        xh[i] = h;
        z1[i] = y;
    }
}

__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
// d/dv [0.5 v (1 + erf(v / sqrt 2))] = 0.5 (1 + erf(v / sqrt 2)) + v exp(-v^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float v) {
    return 0.5f * (1.0f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
}

// a += b1 (kept: the GELU backward needs the pre-activation); z2 = dropout2(gelu(a))
__global__ void __launch_bounds__(256) gelu_dropout_kernel(float* __restrict__ a, const float* __restrict__ b1,
                                                           int batch, int hidden, const uint8_t* __restrict__ keep,
                                                           float keep_scale, float* __restrict__ z2) {
    const size_t total = static_cast<size_t>(batch) * hidden;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = a[i] + b1[i % hidden];
        a[i] = v;
        float g = gelu_erf_f(v);
        if (keep != nullptr) g = keep[i] ? g * keep_scale : 0.f;
        z2[i] = g;
    }
}

// one thread per row: logits += b2, weighted softmax cross entropy with ignore (target outside [0, C)), dlogits
__global__ void __launch_bounds__(128) ce_kernel(float* __restrict__ logits, const float* __restrict__ b2,
                                                 const int64_t* __restrict__ targets, int batch, int classes,
                                                 const float* __restrict__ class_weight, float* __restrict__ dlog,
                                                 float* __restrict__ loss_b, float* __restrict__ w_b,
                                                 float* __restrict__ logits_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float lg[MAX_CLASSES];
    float mx = -INFINITY;
    for (int c = 0; c < classes; ++c) {
        lg[c] = logits[static_cast<size_t>(b) * classes + c] + b2[c];
        logits_out[static_cast<size_t>(b) * classes + c] = lg[c];
        mx = fmaxf(mx, lg[c]);
    }
    const int64_t y = targets[b];
    if (y < 0 || y >= classes) {
        for (int c = 0; c < classes; ++c) dlog[static_cast<size_t>(b) * classes + c] = 0.f;
        loss_b[b] = 0.f;
        w_b[b] = 0.f;
        return;
    }
    float se = 0.f;
    for (int c = 0; c < classes; ++c) se += expf(lg[c] - mx);
    const float lse = mx + logf(se);
    const float w = class_weight ? class_weight[y] : 1.f;
    for (int c = 0; c < classes; ++c)
        dlog[static_cast<size_t>(b) * classes + c] = w * (expf(lg[c] - lse) - (c == y ? 1.f : 0.f));
    loss_b[b] = w * (lse - lg[y]);
    w_b[b] = w;
}

// Column sums over the batch rows. A block owns 32 columns (lane = column: a warp load is one 128-byte line of a row), its
// RED_WARPS warps deal the rows round-robin and their partials meet in shared memory, added in warp order — a fixed
// summation order, so the result does not depend on scheduling. (The first version ran one THREAD per column over all rows:
// a 256-deep chain of dependent L2 round trips per launch, the pattern that cost the linear head 2x before probe.cu's
// bn_stats / backward kernels were dealt the same way.)
constexpr int RED_WARPS = 8;
__global__ void __launch_bounds__(32 * RED_WARPS) colsum_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                                                float* __restrict__ out) {
    __shared__ float part[RED_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (j < cols)
        for (int b = warp; b < rows; b += RED_WARPS) s += src[static_cast<size_t>(b) * ld + j];
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < RED_WARPS; ++w) t += part[w][lane];
        out[j] = t;
    }
}

// the three batch sums behind the cross entropy in one launch: db2[c] = sum_b dlog[b, c], tail = [sum_b loss_b, sum_b w_b];
// column j < C reads dlog, column C the losses, column C + 1 the weights (grid = ceil((C + 2) / 32) blocks)
__global__ void __launch_bounds__(32 * RED_WARPS) ce_sums_kernel(const float* __restrict__ dlog, const float* __restrict__ loss_b,
                                                                 const float* __restrict__ w_b, int rows, int classes,
                                                                 float* __restrict__ db2, float* __restrict__ tail) {
    __shared__ float part[RED_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (j < classes + 2)
        for (int b = warp; b < rows; b += RED_WARPS)
            s += j < classes ? dlog[static_cast<size_t>(b) * classes + j] : (j == classes ? loss_b[b] : w_b[b]);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < classes + 2) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < RED_WARPS; ++w) t += part[w][lane];
        if (j < classes)
            db2[j] = t;
        else
            tail[j - classes] = t;
    }
}

// da = dz2 * keep2 * gelu'(a), in place over dz2
__global__ void __launch_bounds__(256) gelu_bwd_kernel(float* __restrict__ dz2, const float* __restrict__ a, int batch,
                                                       int hidden, const uint8_t* __restrict__ keep, float keep_scale) {
    const size_t total = static_cast<size_t>(batch) * hidden;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float g = dz2[i];
        if (keep != nullptr) g = keep[i] ? g * keep_scale : 0.f;
        dz2[i] = g * gelu_erf_grad(a[i]);
    }
}

// dgamma[d] = sum_b dz1 keep1 xh, dbeta[d] = sum_b dz1 keep1 — rows dealt over the block's warps like colsum_kernel
__global__ void __launch_bounds__(32 * RED_WARPS) bn_bwd_kernel(const float* __restrict__ dz1, const float* __restrict__ xh,
                                                                int batch, int dim, const uint8_t* __restrict__ keep,
                                                                float keep_scale, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta) {
    __shared__ float pg[RED_WARPS][32];
    __shared__ float pb[RED_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + lane;
    float sg = 0.f, sb = 0.f;
    if (d < dim)
        for (int b = warp; b < batch; b += RED_WARPS) {
            const size_t i = static_cast<size_t>(b) * dim + d;
            float dy = dz1[i];
            if (keep != nullptr) dy = keep[i] ? dy * keep_scale : 0.f;
            sg += dy * xh[i];
            sb += dy;
        }
    pg[warp][lane] = sg;
    pb[warp][lane] = sb;
    __syncthreads();
    if (warp == 0 && d < dim) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < RED_WARPS; ++w) {
            tg += pg[w][lane];
            tb += pb[w][lane];
        }
        dgamma[d] = tg;
        dbeta[d] = tb;
    }
}

// torch.optim.AdamW with grad = grad_scale * bucket[i] / bucket[wsum_index]; a head with no labelled row in the global
// batch (weight sum 0) is left untouched, like a parameter whose .grad is None
__global__ void __launch_bounds__(256) adamw_ex_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                       const float* __restrict__ bucket, int64_t numel,
                                                       int64_t wsum_index, float grad_scale, float lr, float beta1,
                                                       float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    const float wsum = bucket[wsum_index];
    if (!(wsum > 0.f)) return;
    const float inv = grad_scale / wsum;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float g = bucket[i] * inv;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        p[i] = pi;
    }
}

// the same update with everything that changes between optimisation steps read from device memory — the step number
// (steps completed so far; this is step *steps_done + 1) and the task weight — so a multitask step replays from a CUDA graph
__global__ void __launch_bounds__(256) adamw_ex_dev_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                           const float* __restrict__ bucket, int64_t numel,
                                                           int64_t wsum_index, const float* __restrict__ grad_scale,
                                                           const int64_t* __restrict__ steps_done, float lr, float beta1,
                                                           float beta2, float eps, float wd) {
    const float wsum = bucket[wsum_index];
    if (!(wsum > 0.f)) return;
    const float step = static_cast<float>(*steps_done + 1);
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
    const float inv = *grad_scale / wsum;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float g = bucket[i] * inv;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        p[i] = pi;
    }
}

}  // namespace
}  // namespace vlmfw

using namespace vlmfw;

extern "C" size_t vlmfw_probe_deep_scratch_floats(int32_t batch, int32_t dim, int32_t hidden, int32_t classes) {
    if (batch <= 0 || dim <= 0 || hidden <= 0 || classes <= 0) return 0;
    const size_t b = batch;
    // z1 | xh | dz1 (B*D each), a | z2 | dz2 (B*Hd each), raw logits | dlog (B*C each), loss_b | w_b (B each)
    return 3 * b * dim + 3 * b * hidden + 2 * b * classes + 2 * b;
}

extern "C" int vlmfw_probe_deep_grads(const float* x, const int64_t* targets, int32_t batch, int32_t dim, int32_t hidden,
                                      int32_t classes, const double* stats, int32_t n_global, float bn_eps,
                                      const float* params, const float* class_weight, const uint8_t* keep1,
                                      const uint8_t* keep2, float dropout_p, float* bucket, float* logits_out,
                                      float* scratch, void* stream) {
    VLMFW_REQUIRE(x && targets && stats && params && bucket && logits_out && scratch, "probe_deep_grads: null pointer");
    VLMFW_REQUIRE(batch > 0 && dim > 0 && hidden > 0 && classes > 0 && classes <= MAX_CLASSES && n_global >= batch,
                  "probe_deep_grads: bad shape B=%d D=%d Hd=%d C=%d n_global=%d", batch, dim, hidden, classes, n_global);
    VLMFW_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "probe_deep_grads: dropout_p out of range");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t B = batch, D = dim, H = hidden, C = classes;
    const float* W1 = params;
    const float* b1 = W1 + H * D;
    const float* W2 = b1 + H;
    const float* b2 = W2 + C * H;
    const float* gamma = b2 + C;
    const float* beta = gamma + D;
    float* dW1 = bucket;
    float* db1 = dW1 + H * D;
    float* dW2 = db1 + H;
    float* db2 = dW2 + C * H;
    float* dgamma = db2 + C;
    float* dbeta = dgamma + D;
    float* tail = dbeta + D;  // loss_sum, weight_sum
    float* z1 = scratch;
    float* xh = z1 + B * D;
    float* dz1 = xh + B * D;
    float* a = dz1 + B * D;
    float* z2 = a + B * H;
    float* dz2 = z2 + B * H;
    float* raw = dz2 + B * H;
    float* dlog = raw + B * C;
    float* loss_b = dlog + B * C;
    float* w_b = loss_b + B;
    const float ks = 1.f / (1.f - dropout_p);
    const int eb = static_cast<int>((B * H + 255) / 256) < 1184 ? static_cast<int>((B * H + 255) / 256) : 1184;

    // forward
    bn_dropout_kernel<<<dim3((dim + 255) / 256, batch < 64 ? batch : 64), 256, 0, s>>>(x, batch, dim, stats, n_global, bn_eps,
                                                                                     gamma, beta, keep1, ks, z1, xh);
    VLMFW_LAUNCH_CHECK();
    if ((rc = sgemm<false, true>(z1, dim, W1, dim, a, hidden, batch, hidden, dim, s)) != VLMFW_OK) return rc;
    gelu_dropout_kernel<<<eb, 256, 0, s>>>(a, b1, batch, hidden, keep2, ks, z2);
    VLMFW_LAUNCH_CHECK();
    if ((rc = sgemm<false, true>(z2, hidden, W2, hidden, raw, classes, batch, classes, hidden, s)) != VLMFW_OK) return rc;
    ce_kernel<<<(batch + 127) / 128, 128, 0, s>>>(raw, b2, targets, batch, classes, class_weight, dlog, loss_b, w_b, logits_out);
    VLMFW_LAUNCH_CHECK();
    // backward
    if ((rc = sgemm<true, false>(dlog, classes, z2, hidden, dW2, hidden, classes, hidden, batch, s)) != VLMFW_OK) return rc;
    ce_sums_kernel<<<(classes + 2 + 31) / 32, 32 * RED_WARPS, 0, s>>>(dlog, loss_b, w_b, batch, classes, db2, tail);
    VLMFW_LAUNCH_CHECK();
    if ((rc = sgemm<false, false>(dlog, classes, W2, hidden, dz2, hidden, batch, hidden, classes, s)) != VLMFW_OK) return rc;
    gelu_bwd_kernel<<<eb, 256, 0, s>>>(dz2, a, batch, hidden, keep2, ks);
    VLMFW_LAUNCH_CHECK();
    if ((rc = sgemm<true, false>(dz2, hidden, z1, dim, dW1, dim, hidden, dim, batch, s)) != VLMFW_OK) return rc;
    colsum_kernel<<<(hidden + 31) / 32, 32 * RED_WARPS, 0, s>>>(dz2, batch, hidden, hidden, db1);
    VLMFW_LAUNCH_CHECK();
    if ((rc = sgemm<false, false>(dz2, hidden, W1, dim, dz1, dim, batch, dim, hidden, s)) != VLMFW_OK) return rc;
    bn_bwd_kernel<<<(dim + 31) / 32, 32 * RED_WARPS, 0, s>>>(dz1, xh, batch, dim, keep1, ks, dgamma, dbeta);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_adamw_ex(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                                    int64_t weight_sum_index, float grad_scale, int32_t step, float lr, float beta1,
                                    float beta2, float eps, float weight_decay, void* stream) {
    VLMFW_REQUIRE(params && exp_avg && exp_avg_sq && bucket && numel > 0 && step >= 1, "probe_adamw_ex: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    int grid = static_cast<int>((numel + 255) / 256);
    if (grid > 1184) grid = 1184;
    adamw_ex_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, exp_avg, exp_avg_sq, bucket, numel,
                                                                         weight_sum_index, grad_scale, lr, beta1, beta2,
                                                                         eps, weight_decay, bc1, bc2_sqrt);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_adamw_ex_dev(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                                        int64_t weight_sum_index, const float* grad_scale, const int64_t* steps_done, float lr,
                                        float beta1, float beta2, float eps, float weight_decay, void* stream) {
    VLMFW_REQUIRE(params && exp_avg && exp_avg_sq && bucket && grad_scale && steps_done && numel > 0,
                  "probe_adamw_ex_dev: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    int grid = static_cast<int>((numel + 255) / 256);
    if (grid > 1184) grid = 1184;
    adamw_ex_dev_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, exp_avg, exp_avg_sq, bucket, numel,
                                                                             weight_sum_index, grad_scale, steps_done, lr,
                                                                             beta1, beta2, eps, weight_decay);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}
