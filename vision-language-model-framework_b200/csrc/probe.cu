// Probe heads on pooled features: inference epilogue and the data-parallel train step, all fp32.
//
//   probe_forward_kernel : reference probing/models/base_probe.py:7-22 (make_head / make_head_deeper) in eval mode,
//                          every head of a LinearProbe / MultiTaskProbe in ONE launch, plus the argmax that
//                          LinearProbe.predict / MultiTaskProbe.predict take (linear_probe.py:41-45,
//                          multitask_probe.py:41-45) — the "parsed attribute" integers of base_tester.py:53-57.
//   train step kernels   : reference probing/train/singletask_trainer.py:183-191 (head_only(x) -> weighted CE with
//                          ignore_index=-1, :105-108) + base_trainer.py:166-181 (backward, AdamW step), split into
//                          phases with plain-sum outputs so that a data-parallel caller can all-reduce them and land
//                          on exactly the single-process result for the concatenated batch.
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {
namespace {

constexpr int MAX_HEADS = 16;
constexpr int MAX_CLASSES = 32;

struct HeadPack {
    vlmfw_head h[MAX_HEADS];
    int n;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// grid = batch, block = 256. smem: xhat[dim] | hid[max_hidden] | lg[MAX_CLASSES]
__global__ void __launch_bounds__(256) probe_forward_kernel(const float* __restrict__ pooled, int dim, HeadPack hp,
                                                            int max_hidden, float* __restrict__ logits, int logits_ld,
                                                            int* __restrict__ argmax_out) {
    extern __shared__ float psm[];
    float* xhat = psm;
    float* hid = psm + dim;
    float* lg = hid + max_hidden;
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float* x = pooled + static_cast<size_t>(b) * dim;
    for (int hi = 0; hi < hp.n; ++hi) {
        const vlmfw_head& H = hp.h[hi];
        for (int d = threadIdx.x; d < dim; d += blockDim.x) {
            // BatchNorm1d eval: (x - running_mean) / sqrt(running_var + eps) * weight + bias
            xhat[d] = (x[d] - H.bn_mean[d]) * rsqrtf(H.bn_var[d] + H.bn_eps) * H.bn_weight[d] + H.bn_bias[d];
        }
        __syncthreads();
        const float* vec = xhat;
        int vec_len = dim;
        const float* wout = H.w1;
        const float* bout = H.b1;
        if (H.hidden > 0) {
            for (int j = warp; j < H.hidden; j += nwarps) {
                const float* wr = H.w1 + static_cast<size_t>(j) * dim;
                float acc = 0.f;
                for (int d = lane; d < dim; d += 32) acc += wr[d] * xhat[d];
                acc = warp_sum(acc);
                if (lane == 0) hid[j] = gelu_erf(acc + H.b1[j]);
            }
            __syncthreads();
            vec = hid;
            vec_len = H.hidden;
            wout = H.w2;
            bout = H.b2;
        }
        for (int c = warp; c < H.classes; c += nwarps) {
            const float* wr = wout + static_cast<size_t>(c) * vec_len;
            float acc = 0.f;
            for (int d = lane; d < vec_len; d += 32) acc += wr[d] * vec[d];
            acc = warp_sum(acc);
            if (lane == 0) {
                const float v = acc + bout[c];
                lg[c] = v;
                logits[static_cast<size_t>(b) * logits_ld + H.logit_offset + c] = v;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && argmax_out != nullptr) {
            int best = 0;
            float bv = lg[0];
            for (int c = 1; c < H.classes; ++c) {
                if (lg[c] > bv) {  // strict: first maximum wins, like torch.argmax
                    bv = lg[c];
                    best = c;
                }
            }
            argmax_out[static_cast<size_t>(b) * hp.n + hi] = best;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ train step
// Data-parallel exchange over NVLink peer memory, fused into the step's own kernels (vlmfw_comm below). Every rank owns one
// exchange buffer [stats slots [world][2D] x 16 B | bucket slots [world][NB] x 8 B] that all peers have mapped (cudaIpc).
// Every value travels WITH its flag in one store (the scheme NCCL calls LL): a bucket float goes as the 8-byte pair
// {bits, epoch + 1}, a BatchNorm sum as the 16-byte quad {lo, epoch + 1, hi, epoch + 1} — 8-byte aligned stores are
// single-copy atomic, over NVLink too, so a reader that sees the flag of this exchange sees the value that came with it. A
// producer kernel writes its rank's contribution into slot `rank` of EVERY rank's buffer (plain stores to peer addresses);
// the consumer thread of an element polls its LOCAL slots until each carries this exchange's flag and adds them in rank
// order — the same order on every rank, so all replicas stay bit-identical. No fence, no separate flag, no collective
// kernel: the cost of an exchange is one NVLink store latency (the first version — plain stores, fence.sys, one flag per
// CTA per peer — paid two system-scope fences per producer CTA: 16 us per step on ONE rank; the two dependent
// ncclAllReduce launches this replaces cost 21 us of a 96 us step at two ranks). epoch = exchanges completed through this
// communicator, a device-side counter the step's last kernel advances, so flags never repeat and nothing is reset.
// Slot reuse is safe without further handshakes: a rank reaches bn_stats of step t + 1 only after its AdamW of step t has
// seen every rank's backward of step t, which those ranks launched after their own consumers of the step-t statistics.
constexpr int DP_MAX_WORLD = 8;
struct DpLink {
    int world = 1, rank = 0;
    uint4* stats[DP_MAX_WORLD];   // stats[r] = rank r's stats slots [world][2D] as mapped here
    uint2* bucket[DP_MAX_WORLD];  // bucket[r] = rank r's bucket slots [world][NB]
    const int64_t* epoch = nullptr;  // exchanges completed so far through this communicator (NOT the optimiser's step
                                     // count, which the host may reset between graph replays)
    int* error = nullptr;            // set to 1 when a poll ran out of patience (a peer died): the step's result is invalid
};
constexpr long long DP_SPIN_LIMIT = 1ll << 24;  // ~ seconds of polling

__device__ __forceinline__ uint32_t dp_flag(const DpLink& lk) { return static_cast<uint32_t>(*lk.epoch + 1); }
__device__ __forceinline__ void dp_put_f32(const DpLink& lk, size_t at, int nb, float t, uint32_t flag) {
    const uint2 v = make_uint2(__float_as_uint(t), flag);
    for (int r = 0; r < lk.world; ++r) {
        uint2* dst = lk.bucket[r] + static_cast<size_t>(lk.rank) * nb + at;
        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(v.x), "r"(v.y) : "memory");
    }
}
__device__ __forceinline__ void dp_put_f64(const DpLink& lk, size_t at, int len, double t, uint32_t flag) {
    const uint64_t bits = static_cast<uint64_t>(__double_as_longlong(t));
    const uint32_t lo = static_cast<uint32_t>(bits), hi = static_cast<uint32_t>(bits >> 32);
    for (int r = 0; r < lk.world; ++r) {
        uint4* dst = lk.stats[r] + static_cast<size_t>(lk.rank) * len + at;
        asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(flag), "r"(hi), "r"(flag)
                     : "memory");
    }
}
// sum over ranks (rank order) of element `at` of this rank's bucket / stats slots, each polled until it carries `flag`.
// The loads of all ranks' slots are issued together and only the missing ones are polled again: one L2 round trip when
// everything has arrived, instead of `world` dependent ones (8 ranks: ~4 us per consumer kernel with a slot-by-slot loop).
__device__ __forceinline__ float dp_get_f32(const DpLink& lk, size_t at, int nb, uint32_t flag) {
    float vals[DP_MAX_WORLD];
    unsigned pending = (1u << lk.world) - 1u;
    long long spins = 0;
    const uint2* base = lk.bucket[lk.rank] + at;
    while (pending) {
        uint32_t x[DP_MAX_WORLD], f[DP_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
            if (pending >> r & 1u)
                asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];"
                             : "=r"(x[r]), "=r"(f[r])
                             : "l"(base + static_cast<size_t>(r) * nb)
                             : "memory");
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
            if ((pending >> r & 1u) && f[r] == flag) {
                vals[r] = __uint_as_float(x[r]);
                pending &= ~(1u << r);
            }
        if (pending && ++spins > DP_SPIN_LIMIT) {
            *lk.error = 1;
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (pending >> r & 1u) vals[r] = 0.f;
            break;
        }
    }
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
        if (r < lk.world) t += vals[r];
    return t;
}
__device__ __forceinline__ double dp_get_f64(const DpLink& lk, size_t at, int len, uint32_t flag) {
    double vals[DP_MAX_WORLD];
    unsigned pending = (1u << lk.world) - 1u;
    long long spins = 0;
    const uint4* base = lk.stats[lk.rank] + at;
    while (pending) {
        uint32_t lo[DP_MAX_WORLD], f0[DP_MAX_WORLD], hi[DP_MAX_WORLD], f1[DP_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
            if (pending >> r & 1u)
                asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lo[r]), "=r"(f0[r]), "=r"(hi[r]), "=r"(f1[r])
                             : "l"(base + static_cast<size_t>(r) * len)
                             : "memory");
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
            if ((pending >> r & 1u) && f0[r] == flag && f1[r] == flag) {
                vals[r] = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(hi[r]) << 32) | lo[r]));
                pending &= ~(1u << r);
            }
        if (pending && ++spins > DP_SPIN_LIMIT) {
            *lk.error = 1;
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (pending >> r & 1u) vals[r] = 0.0;
            break;
        }
    }
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
        if (r < lk.world) t += vals[r];
    return t;
}

// stats[d] = sum_b x[b,d], stats[dim+d] = sum_b x[b,d]^2 in fp64 (the variance is a difference of these).
// A block owns 32 feature columns (lane = column: every warp load is one 128-byte line); its STAT_WARPS warps deal the
// rows round-robin and their partials are combined in warp order through shared memory — a fixed summation order, so
// the result does not depend on scheduling. (The first version ran ONE thread per column over all rows on 9 blocks:
// a 256-deep chain of dependent L2 round trips, ~20x the time of everything else in the step.)
constexpr int STAT_WARPS = 16;
template <bool DP>
__global__ void __launch_bounds__(32 * STAT_WARPS) bn_stats_kernel(const float* __restrict__ x, int batch, int dim,
                                                                   double* __restrict__ stats, const DpLink lk) {
    __shared__ double ss[STAT_WARPS][32];
    __shared__ double sq[STAT_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    if (d < dim) {
#pragma unroll 4
        for (int b = warp; b < batch; b += STAT_WARPS) {
            const double v = static_cast<double>(x[static_cast<size_t>(b) * dim + d]);
            s += v;
            q += v * v;
        }
    }
    ss[warp][lane] = s;
    sq[warp][lane] = q;
    __syncthreads();
    if (warp == 0 && d < dim) {
        double st = 0.0, qt = 0.0;
#pragma unroll
        for (int w = 0; w < STAT_WARPS; ++w) {
            st += ss[w][lane];
            qt += sq[w][lane];
        }
        if constexpr (DP) {
            // this rank's sums into slot `rank` of every rank's buffer (its own included: the consumer adds all slots)
            const uint32_t flag = dp_flag(lk);
            dp_put_f64(lk, d, 2 * dim, st, flag);
            dp_put_f64(lk, dim + d, 2 * dim, qt, flag);
        } else {
            stats[d] = st;
            stats[dim + d] = qt;
        }
    }
}

// DP: global statistics = sum over ranks of the slots, in rank order, once per step (fused into the head of the dropout-mask
// kernel's grid)
__device__ __forceinline__ void dp_combine_stats(const DpLink& lk, int dim, double* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * dim) stats[i] = dp_get_f64(lk, i, 2 * dim, dp_flag(lk));
}

struct TrainArgs {
    const float* x;
    const int64_t* targets;
    int batch, dim, classes;
    const double* stats;
    int n_global;
    float bn_eps;
    const float* gamma;
    const float* beta;
    const float* weight;
    const float* bias;
    const float* class_weight;
    const uint8_t* keep;
    float keep_scale;  // 1 / (1 - p)
    float* bucket;
    float* logits;
    float* scratch;  // dlogits [B,C] | loss_b [B] | w_b [B]
};

__device__ __forceinline__ void bn_mean_rstd(const double* stats, int dim, int d, int n, float eps, float& mean,
                                             float& rstd) {
    const double mu = stats[d] / n;
    double var = stats[dim + d] / n - mu * mu;  // biased variance, as BatchNorm normalises with in train mode
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(mu);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// Phase A: one block per sample. z = dropout(BN_train(x)); logits = W z + b; softmax-CE with class weights.
__global__ void __launch_bounds__(256) probe_train_fwd_kernel(TrainArgs a) {
    extern __shared__ float tsm[];
    float* z = tsm;           // [dim]
    float* lg = tsm + a.dim;  // [classes]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int d = threadIdx.x; d < a.dim; d += blockDim.x) {
        float mean, rstd;
        bn_mean_rstd(a.stats, a.dim, d, a.n_global, a.bn_eps, mean, rstd);
        float y = (a.x[static_cast<size_t>(b) * a.dim + d] - mean) * rstd * a.gamma[d] + a.beta[d];
        if (a.keep != nullptr) y = a.keep[static_cast<size_t>(b) * a.dim + d] ? y * a.keep_scale : 0.f;
        z[d] = y;
    }
    __syncthreads();
    for (int c = warp; c < a.classes; c += nwarps) {
        const float* wr = a.weight + static_cast<size_t>(c) * a.dim;
        float acc = 0.f;
        for (int d = lane; d < a.dim; d += 32) acc += wr[d] * z[d];
        acc = warp_sum(acc);
        if (lane == 0) {
            lg[c] = acc + a.bias[c];
            a.logits[static_cast<size_t>(b) * a.classes + c] = lg[c];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float* dl = a.scratch + static_cast<size_t>(b) * a.classes;
        float* loss_b = a.scratch + static_cast<size_t>(a.batch) * a.classes;
        float* w_b = loss_b + a.batch;
        const int64_t y = a.targets[b];
        if (y < 0 || y >= a.classes) {  // ignore_index = -1: no loss, no gradient
            for (int c = 0; c < a.classes; ++c) dl[c] = 0.f;
            loss_b[b] = 0.f;
            w_b[b] = 0.f;
        } else {
            float mx = lg[0];
            for (int c = 1; c < a.classes; ++c) mx = fmaxf(mx, lg[c]);
            float se = 0.f;
            for (int c = 0; c < a.classes; ++c) se += expf(lg[c] - mx);
            const float lse = mx + logf(se);
            const float w = a.class_weight ? a.class_weight[y] : 1.f;
            for (int c = 0; c < a.classes; ++c) dl[c] = w * (expf(lg[c] - lse) - (c == y ? 1.f : 0.f));
            loss_b[b] = w * (lse - lg[y]);
            w_b[b] = w;
        }
    }
}

// Phase B: a block owns 32 feature columns (lane = column), its BWD_WARPS warps deal the batch rows round-robin; the
// per-warp partial sums of (dW[c][d], dgamma[d], dbeta[d]) meet in shared memory and are added in warp order — fixed
// summation order, deterministic. bucket = [dW (C*D) | db (C) | dgamma (D) | dbeta (D) | loss_sum | weight_sum]
constexpr int BWD_WARPS = 8;
template <bool DP>
__global__ void __launch_bounds__(32 * BWD_WARPS) probe_train_bwd_kernel(TrainArgs a, const DpLink lk, int nb) {
    // DP: every bucket element goes to slot `rank` of every rank's exchange buffer instead of a.bucket
    uint32_t flag = 0;
    if constexpr (DP) flag = dp_flag(lk);
    auto put = [&](size_t at, float t) {
        if constexpr (DP) {
            dp_put_f32(lk, at, nb, t, flag);
        } else {
            a.bucket[at] = t;
        }
    };
    extern __shared__ float bsm[];  // dlogits [B*C] | partials [BWD_WARPS][C + 2][32]
    const int nd = a.batch * a.classes;
    for (int i = threadIdx.x; i < nd; i += blockDim.x) bsm[i] = a.scratch[i];
    __syncthreads();
    const int C = a.classes, D = a.dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + lane;
    float* part = bsm + ((nd + 31) & ~31);
    {
        float wcol[MAX_CLASSES];
        float dw[MAX_CLASSES];
        float dgamma = 0.f, dbeta = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_CLASSES; ++c) {
            wcol[c] = (c < C && d < D) ? a.weight[static_cast<size_t>(c) * D + d] : 0.f;
            dw[c] = 0.f;
        }
        if (d < D) {
            float mean, rstd;
            bn_mean_rstd(a.stats, D, d, a.n_global, a.bn_eps, mean, rstd);
            const float gm = a.gamma[d], bt = a.beta[d];
#pragma unroll 2
            for (int b = warp; b < a.batch; b += BWD_WARPS) {
                const float xh = (a.x[static_cast<size_t>(b) * D + d] - mean) * rstd;
                float keep = 1.f;
                if (a.keep != nullptr) keep = a.keep[static_cast<size_t>(b) * D + d] ? a.keep_scale : 0.f;
                const float z = (xh * gm + bt) * keep;
                float dz = 0.f;
#pragma unroll
                for (int c = 0; c < MAX_CLASSES; ++c) {
                    if (c < C) {
                        const float g = bsm[b * C + c];
                        dw[c] += g * z;
                        dz += g * wcol[c];
                    }
                }
                const float dy = dz * keep;
                dgamma += dy * xh;
                dbeta += dy;
            }
        }
        float* mine = part + static_cast<size_t>(warp) * (C + 2) * 32 + lane;
#pragma unroll
        for (int c = 0; c < MAX_CLASSES; ++c)
            if (c < C) mine[c * 32] = dw[c];
        mine[C * 32] = dgamma;
        mine[(C + 1) * 32] = dbeta;
    }
    __syncthreads();
    if (d < D) {
        for (int k = warp; k < C + 2; k += BWD_WARPS) {  // quantity k of column d: add the warps' partials in order
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < BWD_WARPS; ++w) t += part[(static_cast<size_t>(w) * (C + 2) + k) * 32 + lane];
            const size_t at = k < C ? static_cast<size_t>(k) * D + d
                                    : static_cast<size_t>(C) * D + C + static_cast<size_t>(k - C) * D + d;
            put(at, t);
        }
    }
    if (blockIdx.x == 0) {
        // db[c] and the two loss sums: one warp per quantity, lanes stride the batch, xor-shuffle tree (fixed order)
        const float* loss_b = a.scratch + static_cast<size_t>(a.batch) * C;
        const float* w_b = loss_b + a.batch;
        for (int k = warp; k < C + 2; k += BWD_WARPS) {
            float t = 0.f;
            for (int b = lane; b < a.batch; b += 32) t += k < C ? bsm[b * C + k] : (k == C ? loss_b[b] : w_b[b]);
            t = warp_sum(t);
            if (lane == 0) put(static_cast<size_t>(C) * D + (k < C ? k : C + 2 * D + (k - C)), t);
        }
    }
}

// torch.optim.AdamW (decoupled weight decay), grad = bucket[i] / bucket[weight_sum_index]
__device__ __forceinline__ void adamw_body(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                           const float* __restrict__ bucket, int64_t numel, int64_t wsum_index, float lr,
                                           float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    const float wsum = bucket[wsum_index];
    // every target of the (global) batch ignored: the reference's mean loss is 0/0 and its step poisons the head with NaN;
    // here the step is skipped (no decay, no moment update), like the multitask path (probe_deep.cu adamw_ex_kernel)
    if (!(wsum > 0.f)) return;
    const float inv = 1.f / wsum;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float g = bucket[i] * inv;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ bucket, int64_t numel,
                                                    int64_t wsum_index, float lr, float beta1, float beta2, float eps,
                                                    float wd, float bc1, float bc2_sqrt) {
    adamw_body(p, m, v, bucket, numel, wsum_index, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
}

// The same update with the step number read from device memory (steps completed so far; this is step *steps_done + 1):
// nothing in the launch changes from one optimisation step to the next, so the whole train step replays from a CUDA graph.
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, float* __restrict__ m,
                                                        float* __restrict__ v, const float* __restrict__ bucket,
                                                        int64_t numel, int64_t wsum_index,
                                                        const int64_t* __restrict__ steps_done, float lr, float beta1,
                                                        float beta2, float eps, float wd) {
    const float step = static_cast<float>(*steps_done + 1);
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
    adamw_body(p, m, v, bucket, numel, wsum_index, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
}

// DP form: the gradient bucket is the sum over ranks of this rank's slots (rank order: bit-identical on every rank); block 0
// also leaves the all-reduced [loss_sum, weight_sum] in `bucket_out` for the host's loss read-back. `per_step` = CTAs of
// probe_train_bwd_kernel.
__global__ void __launch_bounds__(256) adamw_dp_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                       const DpLink lk, int nb, float* __restrict__ bucket_out,
                                                       int64_t numel, int64_t wsum_index,
                                                       const int64_t* __restrict__ steps_done, float lr, float beta1,
                                                       float beta2, float eps, float wd) {
    const uint32_t flag = dp_flag(lk);
    auto total = [&](int64_t i) { return dp_get_f32(lk, static_cast<size_t>(i), nb, flag); };
    const float step = static_cast<float>(*steps_done + 1);
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
    const float wsum = total(wsum_index);
    if (blockIdx.x == 0 && threadIdx.x < 2) bucket_out[numel + threadIdx.x] = total(numel + threadIdx.x);
    if (!(wsum > 0.f)) return;  // all targets ignored: skip the update (see adamw_body)
    const float inv = 1.f / wsum;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float g = total(i) * inv;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

__global__ void counter_add_kernel(int64_t* ctr, int64_t inc) { *ctr += inc; }
__global__ void counter_add2_kernel(int64_t* a, int64_t* b) {
    *a += 1;
    *b += 1;
}

// running_mean/var <- (1-momentum) * old + momentum * batch stat (unbiased variance), nn.BatchNorm1d train mode
__global__ void bn_running_kernel(float* __restrict__ rm, float* __restrict__ rv, const double* __restrict__ stats,
                                  int dim, int n, float momentum) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    const double mu = stats[d] / n;
    double var = stats[dim + d] / n - mu * mu;
    if (var < 0.0) var = 0.0;
    const double unbiased = n > 1 ? var * n / (n - 1) : var;
    rm[d] = (1.f - momentum) * rm[d] + momentum * static_cast<float>(mu);
    rv[d] = (1.f - momentum) * rv[d] + momentum * static_cast<float>(unbiased);
}

}  // namespace
}  // namespace vlmfw

using namespace vlmfw;

extern "C" int vlmfw_probe_forward(const float* pooled, int32_t batch, int32_t dim, const vlmfw_head* heads,
                                   int32_t n_heads, float* logits_out, int32_t logits_ld, int32_t* argmax_out,
                                   void* stream) {
    VLMFW_REQUIRE(pooled && heads && logits_out, "probe_forward: null pointer");
    VLMFW_REQUIRE(batch > 0 && dim > 0, "probe_forward: bad shape");
    VLMFW_REQUIRE(n_heads > 0 && n_heads <= MAX_HEADS, "probe_forward: 1..%d heads supported", MAX_HEADS);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    HeadPack hp;
    hp.n = n_heads;
    int max_hidden = 0, total = 0;
    for (int i = 0; i < n_heads; ++i) {
        hp.h[i] = heads[i];
        VLMFW_REQUIRE(heads[i].classes > 0 && heads[i].classes <= MAX_CLASSES, "probe_forward: head %d has %d classes",
                      i, heads[i].classes);
        VLMFW_REQUIRE(heads[i].bn_weight && heads[i].bn_bias && heads[i].bn_mean && heads[i].bn_var && heads[i].w1 &&
                          heads[i].b1,
                      "probe_forward: head %d has null parameters", i);
        if (heads[i].hidden > 0) VLMFW_REQUIRE(heads[i].w2 && heads[i].b2, "probe_forward: deeper head %d lacks w2/b2", i);
        max_hidden = heads[i].hidden > max_hidden ? heads[i].hidden : max_hidden;
        total = heads[i].logit_offset + heads[i].classes > total ? heads[i].logit_offset + heads[i].classes : total;
    }
    VLMFW_REQUIRE(logits_ld >= total, "probe_forward: logits_ld %d < %d", logits_ld, total);
    const size_t smem = static_cast<size_t>(dim + max_hidden + MAX_CLASSES) * sizeof(float);
    VLMFW_REQUIRE(smem <= 48 * 1024, "probe_forward: dim+hidden too large");
    probe_forward_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(pooled, dim, hp, max_hidden,
                                                                                  logits_out, logits_ld, argmax_out);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_bn_stats(const float* x, int32_t batch, int32_t dim, double* stats, void* stream) {
    VLMFW_REQUIRE(x && stats && batch > 0 && dim > 0, "probe_bn_stats: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    bn_stats_kernel<false><<<(dim + 31) / 32, 32 * STAT_WARPS, 0, static_cast<cudaStream_t>(stream)>>>(x, batch, dim, stats, DpLink());
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_linear_grads(const float* x, const int64_t* targets, int32_t batch, int32_t dim,
                                        int32_t classes, const double* stats, int32_t n_global, float bn_eps,
                                        const float* gamma, const float* beta, const float* weight, const float* bias,
                                        const float* class_weight, const uint8_t* keep_mask, float dropout_p,
                                        float* bucket, float* logits_out, float* scratch, void* stream) {
    VLMFW_REQUIRE(x && targets && stats && gamma && beta && weight && bias && bucket && logits_out && scratch,
                  "probe_linear_grads: null pointer");
    VLMFW_REQUIRE(batch > 0 && dim > 0 && classes > 0 && classes <= MAX_CLASSES && n_global >= batch,
                  "probe_linear_grads: bad shape");
    VLMFW_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "probe_linear_grads: dropout_p out of range");
    // the backward kernel stages dlogits [B*C] plus its per-warp partials in shared memory (opt-in 96 KB): B*C <= ~21 k
    VLMFW_REQUIRE((((static_cast<size_t>(batch) * classes + 31) & ~static_cast<size_t>(31)) +
                   static_cast<size_t>(BWD_WARPS) * (classes + 2) * 32) * sizeof(float) <= 96 * 1024,
                  "probe_linear_grads: batch*classes = %d*%d exceeds the backward kernel's 96 KB staging (split the batch)", batch,
                  classes);
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    TrainArgs a;
    a.x = x;
    a.targets = targets;
    a.batch = batch;
    a.dim = dim;
    a.classes = classes;
    a.stats = stats;
    a.n_global = n_global;
    a.bn_eps = bn_eps;
    a.gamma = gamma;
    a.beta = beta;
    a.weight = weight;
    a.bias = bias;
    a.class_weight = class_weight;
    a.keep = keep_mask;
    a.keep_scale = 1.f / (1.f - dropout_p);
    a.bucket = bucket;
    a.logits = logits_out;
    a.scratch = scratch;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    probe_train_fwd_kernel<<<batch, 256, static_cast<size_t>(dim + classes) * sizeof(float), s>>>(a);
    VLMFW_LAUNCH_CHECK();
    const size_t bwd_smem = (((static_cast<size_t>(batch) * classes + 31) & ~static_cast<size_t>(31)) +
                             static_cast<size_t>(BWD_WARPS) * (classes + 2) * 32) * sizeof(float);
    static std::atomic<uint64_t> bwd_opt_in{0};  // one bit per device
    VLMFW_CUDA_OK(ensure_dynamic_smem(probe_train_bwd_kernel<false>, 96 * 1024, bwd_opt_in));
    probe_train_bwd_kernel<false><<<(dim + 31) / 32, 32 * BWD_WARPS, bwd_smem, s>>>(a, DpLink(), 0);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_adamw(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                                 int64_t weight_sum_index, int32_t step, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, void* stream) {
    VLMFW_REQUIRE(params && exp_avg && exp_avg_sq && bucket && numel > 0 && step >= 1, "probe_adamw: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    int grid = static_cast<int>((numel + 255) / 256);
    if (grid > 1024) grid = 1024;
    adamw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, exp_avg, exp_avg_sq, bucket, numel,
                                                                      weight_sum_index, lr, beta1, beta2, eps,
                                                                      weight_decay, bc1, bc2_sqrt);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_adamw_dev(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                                     int64_t weight_sum_index, const int64_t* steps_done, float lr, float beta1,
                                     float beta2, float eps, float weight_decay, void* stream) {
    VLMFW_REQUIRE(params && exp_avg && exp_avg_sq && bucket && steps_done && numel > 0, "probe_adamw_dev: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    int grid = static_cast<int>((numel + 255) / 256);
    if (grid > 1024) grid = 1024;
    adamw_dev_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, exp_avg, exp_avg_sq, bucket, numel,
                                                                          weight_sum_index, steps_done, lr, beta1, beta2,
                                                                          eps, weight_decay);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_counter_add(int64_t* counter, int64_t inc, void* stream) {
    VLMFW_REQUIRE(counter, "counter_add: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    counter_add_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(counter, inc);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_probe_bn_running_update(float* running_mean, float* running_var, const double* stats, int32_t dim,
                                             int32_t n_global, float momentum, void* stream) {
    VLMFW_REQUIRE(running_mean && running_var && stats && dim > 0 && n_global > 0, "probe_bn_running_update: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    bn_running_kernel<<<(dim + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(running_mean, running_var, stats,
                                                                                        dim, n_global, momentum);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

// Dropout keep mask, generated where it is consumed. keep[r, c] depends only on (key, GLOBAL row row0 + r, c): a
// counter-based generator (splitmix64 finaliser over the element's global index), so a rank that holds rows
// [row0, row0 + rows) of the batch draws exactly the slice of the single-process mask — the property the sharded train
// step rests on — without anybody materialising the global mask (the host generator it replaces drew n_global x cols
// numbers on every rank and copied its slice over PCIe each step). One thread = 16 consecutive bytes = one 128-bit store.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// DP = true: the grid's first ceil(2 dim / 256) blocks also combine the ranks' BatchNorm sums (dp_combine_stats) — the mask
// does not depend on them, the kernels behind it do, so the exchange's consumer side costs no launch of its own.
template <bool DP>
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ out, int64_t total, int32_t cols, int64_t row0, uint64_t key,
                                    uint32_t threshold, const int64_t* __restrict__ steps_done, uint64_t step_stride,
                                    const DpLink lk, int dim, double* __restrict__ stats) {
    if constexpr (DP) {
        if (static_cast<int>(blockIdx.x) * 256 < 2 * dim) dp_combine_stats(lk, dim, stats);
    }
    const int64_t first = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
    if (first >= total) return;
    const uint64_t base = static_cast<uint64_t>(row0) * static_cast<uint64_t>(cols) + static_cast<uint64_t>(first);
    // graph replay: the per-step part of the key comes from the device-side step counter (key + steps_done * stride)
    if (steps_done != nullptr) key += static_cast<uint64_t>(*steps_done) * step_stride;
    const uint64_t k = splitmix64(key);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 16; ++j) {  // one hash per element, keyed by the element's GLOBAL index
        const uint32_t u = static_cast<uint32_t>(splitmix64(k ^ (base + j)) >> 32);
        w[j >> 2] |= (u >= threshold ? 1u : 0u) << (8 * (j & 3));
    }
    if (first + 16 <= total) {
        *reinterpret_cast<uint4*>(out + first) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int j = 0; first + j < total; ++j) out[first + j] = static_cast<uint8_t>((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
    }
}

static int dropout_keep_mask_launch(uint8_t* keep, int32_t rows, int32_t cols, int64_t row0, uint64_t key,
                                    const int64_t* steps_done, uint64_t step_stride, float dropout_p, void* stream) {
    VLMFW_REQUIRE(rows >= 0 && cols > 0 && row0 >= 0 && dropout_p >= 0.f && dropout_p < 1.f,
                  "dropout_keep_mask: bad arguments");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    if (rows == 0) return VLMFW_OK;  // a rank that owns no row of this batch
    VLMFW_REQUIRE(keep && (reinterpret_cast<uintptr_t>(keep) & 15) == 0, "dropout_keep_mask: keep must be a 16-byte aligned device pointer");
    const int64_t total = static_cast<int64_t>(rows) * cols;
    // keep <=> u >= p * 2^32 with u uniform on [0, 2^32): P(keep) = 1 - p to within 2^-32
    const double t = static_cast<double>(dropout_p) * 4294967296.0;
    const uint32_t threshold = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
    const int64_t threads = (total + 15) / 16;
    dropout_mask_kernel<false><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keep, total, cols, row0, key, threshold, steps_done, step_stride, DpLink(), 0, nullptr);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

extern "C" int vlmfw_dropout_keep_mask(uint8_t* keep, int32_t rows, int32_t cols, int64_t row0, uint64_t key,
                                       float dropout_p, void* stream) {
    return dropout_keep_mask_launch(keep, rows, cols, row0, key, nullptr, 0, dropout_p, stream);
}

extern "C" int vlmfw_dropout_keep_mask_dev(uint8_t* keep, int32_t rows, int32_t cols, int64_t row0, uint64_t key,
                                           const int64_t* steps_done, uint64_t step_stride, float dropout_p,
                                           void* stream) {
    VLMFW_REQUIRE(steps_done, "dropout_keep_mask_dev: steps_done is null");
    return dropout_keep_mask_launch(keep, rows, cols, row0, key, steps_done, step_stride, dropout_p, stream);
}


// ------------------------------------------------------------------------------------------------ NVLink peer exchange
// One exchange buffer per rank (cudaMalloc, exported with cudaIpc and mapped by every peer of the node):
//   [256-byte header (unused) | stats slots [world][2D] x {lo, flag, hi, flag} | bucket slots [world][NB] x {bits, flag}]
struct vlmfw_comm {
    int rank = 0, world = 1;
    int64_t stats_len = 0, bucket_len = 0;  // 2D doubles / NB floats per slot
    size_t stats_off = 0, bucket_off = 0, bytes = 0;
    uint8_t* local = nullptr;
    uint8_t* peer[DP_MAX_WORLD] = {};
    int* error_dev = nullptr;
    int64_t* epoch_dev = nullptr;  // exchanges completed (see DpLink::epoch)
    bool connected = false;
};

extern "C" int vlmfw_comm_create(int32_t rank, int32_t world, int64_t stats_doubles, int64_t bucket_floats, vlmfw_comm** out) {
    VLMFW_REQUIRE(out && world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "comm_create: rank %d / world %d (max %d ranks)",
                  rank, world, DP_MAX_WORLD);
    VLMFW_REQUIRE(stats_doubles > 0 && bucket_floats > 0, "comm_create: empty slots");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    vlmfw_comm* c = new vlmfw_comm();
    c->rank = rank;
    c->world = world;
    c->stats_len = stats_doubles;
    c->bucket_len = bucket_floats;
    c->stats_off = 256;
    c->bucket_off = c->stats_off + static_cast<size_t>(world) * stats_doubles * 16;  // {lo, flag, hi, flag} per double
    c->bucket_off = (c->bucket_off + 255) / 256 * 256;
    c->bytes = c->bucket_off + static_cast<size_t>(world) * bucket_floats * 8;  // {bits, flag} per float
    if (cudaMalloc(reinterpret_cast<void**>(&c->local), c->bytes) != cudaSuccess ||
        cudaMemset(c->local, 0, c->bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&c->error_dev), 16) != cudaSuccess ||  // [int error | pad | int64 epoch]
        cudaMemset(c->error_dev, 0, 16) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        set_error("comm_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (c->local) cudaFree(c->local);
        if (c->error_dev) cudaFree(c->error_dev);
        delete c;
        return VLMFW_E_CUDA;
    }
    c->epoch_dev = reinterpret_cast<int64_t*>(c->error_dev + 2);
    c->peer[rank] = c->local;
    c->connected = world == 1;
    *out = c;
    return VLMFW_OK;
}

extern "C" int32_t vlmfw_comm_handle_bytes(void) { return static_cast<int32_t>(sizeof(cudaIpcMemHandle_t)); }

extern "C" int vlmfw_comm_handle(vlmfw_comm* c, void* handle_out) {
    VLMFW_REQUIRE(c && handle_out, "comm_handle: null argument");
    cudaIpcMemHandle_t h;
    VLMFW_CUDA_OK(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle_out, &h, sizeof(h));
    return VLMFW_OK;
}

// handles: world x vlmfw_comm_handle_bytes() bytes, entry r = rank r's handle (this rank's own entry is ignored)
extern "C" int vlmfw_comm_connect(vlmfw_comm* c, const void* handles) {
    VLMFW_REQUIRE(c && handles, "comm_connect: null argument");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->peer[r] != nullptr) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t*>(handles) + static_cast<size_t>(r) * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        VLMFW_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<uint8_t*>(ptr);
    }
    c->connected = true;
    return VLMFW_OK;
}

extern "C" void vlmfw_comm_destroy(vlmfw_comm* c) {
    if (!c) return;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    if (c->error_dev) cudaFree(c->error_dev);
    delete c;
}

// 1 when a wait of some step gave up (a peer stopped reporting): every result since is invalid. Synchronises the device.
extern "C" int vlmfw_comm_error(vlmfw_comm* c, int32_t* host_flag) {
    VLMFW_REQUIRE(c && host_flag, "comm_error: null argument");
    int v = 0;
    VLMFW_CUDA_OK(cudaMemcpy(&v, c->error_dev, sizeof(int), cudaMemcpyDeviceToHost));
    *host_flag = v;
    return VLMFW_OK;
}

// One whole optimisation step of the linear head (bn_stats -> dropout mask -> forward -> backward -> AdamW -> BatchNorm
// running statistics -> step counter) with the data-parallel exchange fused into its kernels. Seven launches, no collective
// kernel; graph-capturable (everything that changes per step is read from *steps_done). With comm->world == 1 the arithmetic is
// that of the separate entry points above.
extern "C" int vlmfw_probe_linear_step_dp(const vlmfw_probe_step* a, vlmfw_comm* c, void* stream) {
    VLMFW_REQUIRE(a && c, "probe_linear_step_dp: null argument");
    VLMFW_REQUIRE(c->connected, "probe_linear_step_dp: vlmfw_comm_connect has not been called");
    VLMFW_REQUIRE(a->x && a->targets && a->stats && a->params && a->exp_avg && a->exp_avg_sq && a->bucket && a->logits &&
                      a->scratch && a->running_mean && a->running_var && a->steps_done,
                  "probe_linear_step_dp: null pointer");
    const int B = a->batch, D = a->dim, C = a->classes;
    VLMFW_REQUIRE(B > 0 && D > 0 && C > 0 && C <= MAX_CLASSES && a->n_global >= B, "probe_linear_step_dp: bad shape");
    VLMFW_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "probe_linear_step_dp: dropout_p out of range");
    VLMFW_REQUIRE(a->dropout_p == 0.f || a->keep != nullptr, "probe_linear_step_dp: dropout needs a keep-mask buffer");
    VLMFW_REQUIRE((((static_cast<size_t>(B) * C + 31) & ~static_cast<size_t>(31)) + static_cast<size_t>(BWD_WARPS) * (C + 2) * 32) *
                          sizeof(float) <= 96 * 1024,
                  "probe_linear_step_dp: batch*classes exceeds the backward kernel's 96 KB staging (split the batch)");
    const int64_t n_params = static_cast<int64_t>(C) * D + C + 2 * D;
    VLMFW_REQUIRE(c->stats_len == 2 * D && c->bucket_len == n_params + 2, "probe_linear_step_dp: the communicator was created for other sizes");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DpLink lk;
    lk.world = c->world;
    lk.rank = c->rank;
    for (int r = 0; r < c->world; ++r) {
        lk.stats[r] = reinterpret_cast<uint4*>(c->peer[r] + c->stats_off);
        lk.bucket[r] = reinterpret_cast<uint2*>(c->peer[r] + c->bucket_off);
    }
    lk.epoch = c->epoch_dev;
    lk.error = c->error_dev;
    const int nb = static_cast<int>(n_params + 2);
    const unsigned stat_ctas = static_cast<unsigned>((D + 31) / 32), bwd_ctas = stat_ctas;

    bn_stats_kernel<true><<<stat_ctas, 32 * STAT_WARPS, 0, s>>>(a->x, B, D, a->stats, lk);
    VLMFW_LAUNCH_CHECK();
    {
        const int64_t total = a->dropout_p > 0.f ? static_cast<int64_t>(B) * D : 0;
        const double t = static_cast<double>(a->dropout_p) * 4294967296.0;
        const uint32_t threshold = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
        int64_t blocks = ((total + 15) / 16 + 255) / 256;
        const int64_t comb = (2 * D + 255) / 256;
        if (blocks < comb) blocks = comb;
        dropout_mask_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, s>>>(a->keep, total, D, a->row0, a->mask_key, threshold,
                                                                              a->steps_done, a->mask_step_stride, lk, D, a->stats);
        VLMFW_LAUNCH_CHECK();
    }
    TrainArgs ta;
    ta.x = a->x;
    ta.targets = a->targets;
    ta.batch = B;
    ta.dim = D;
    ta.classes = C;
    ta.stats = a->stats;
    ta.n_global = a->n_global;
    ta.bn_eps = a->bn_eps;
    ta.weight = a->params;
    ta.bias = a->params + static_cast<size_t>(C) * D;
    ta.gamma = ta.bias + C;
    ta.beta = ta.gamma + D;
    ta.class_weight = a->class_weight;
    ta.keep = a->dropout_p > 0.f ? a->keep : nullptr;
    ta.keep_scale = 1.f / (1.f - a->dropout_p);
    ta.bucket = a->bucket;
    ta.logits = a->logits;
    ta.scratch = a->scratch;
    probe_train_fwd_kernel<<<B, 256, static_cast<size_t>(D + C) * sizeof(float), s>>>(ta);
    VLMFW_LAUNCH_CHECK();
    const size_t bwd_smem = (((static_cast<size_t>(B) * C + 31) & ~static_cast<size_t>(31)) + static_cast<size_t>(BWD_WARPS) * (C + 2) * 32) * sizeof(float);
    static std::atomic<uint64_t> bwd_opt_in{0};
    VLMFW_CUDA_OK(ensure_dynamic_smem(probe_train_bwd_kernel<true>, 96 * 1024, bwd_opt_in));
    probe_train_bwd_kernel<true><<<bwd_ctas, 32 * BWD_WARPS, bwd_smem, s>>>(ta, lk, nb);
    VLMFW_LAUNCH_CHECK();
    int grid = static_cast<int>((n_params + 255) / 256);
    if (grid > 1024) grid = 1024;
    adamw_dp_kernel<<<grid, 256, 0, s>>>(a->params, a->exp_avg, a->exp_avg_sq, lk, nb, a->bucket, n_params, n_params + 1, a->steps_done, a->lr,
                                         a->beta1, a->beta2, a->eps, a->weight_decay);
    VLMFW_LAUNCH_CHECK();
    bn_running_kernel<<<(D + 127) / 128, 128, 0, s>>>(a->running_mean, a->running_var, a->stats, D, a->n_global, a->bn_momentum);
    VLMFW_LAUNCH_CHECK();
    counter_add2_kernel<<<1, 1, 0, s>>>(a->steps_done, c->epoch_dev);  // optimiser step count and exchange epoch
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}
