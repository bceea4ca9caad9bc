/*
 * vlmfw_b200 — C ABI of the B200-native vision-tower + probe hot path.
 *
 * This is the drop-in boundary for ONE path of andrealb022/Vision-Language-Model-Framework:
 *   VisionBackbone.forward(images) -> [B, D]          (reference models/vision_backbone.py:28-38)
 *     PaliGemmaBackbone.forward  (SigLIP-So400m)      (reference models/paligemma.py:100-112)
 *     LLaVABackbone.forward      (CLIP ViT-L/14-336)  (reference models/llava.py:101-119)
 *     BLIP2Backbone.forward      (EVA ViT-g/14)       (reference models/blip2.py:83-95)
 *   probe heads make_head / make_head_deeper          (reference probing/models/base_probe.py:7-22)
 *   LinearProbe / MultiTaskProbe forward + predict    (reference probing/models/linear_probe.py:35-45,
 *                                                      probing/models/multitask_probe.py:35-45)
 *   probe train step on cached features               (reference probing/train/singletask_trainer.py:183-191,
 *                                                      probing/train/base_trainer.py:166-181)
 * The tower arithmetic the reference delegates to Hugging Face transformers (modeling_siglip.py,
 * modeling_clip.py, modeling_blip_2.py) is implemented here as hand-written sm_100a kernels.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VLMFW_E_* code; nothing throws or aborts across the ABI;
 *     vlmfw_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers owned by the caller unless the name says `host`.
 *   - `stream` is a cudaStream_t passed as void*; work is asynchronous on it.
 *   - a tower handle owns its packed weights; it is not thread-safe.
 *   - there is no CPU fallback: without a sm_100 device every compute entry point fails with VLMFW_E_NODEVICE.
 */
#ifndef VLMFW_B200_H
#define VLMFW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLMFW_ABI_VERSION 2

enum {
    VLMFW_OK = 0,
    VLMFW_E_INVALID = -1,   /* bad argument / shape */
    VLMFW_E_NODEVICE = -2,  /* no CUDA device or not sm_100 */
    VLMFW_E_CUDA = -3,      /* a CUDA runtime / driver call failed */
    VLMFW_E_STATE = -4,     /* weights missing, workspace too small, ... */
    VLMFW_E_NOTFOUND = -5,  /* unknown weight name */
    VLMFW_E_RANGE = -6      /* a weight tensor holds a non-finite value (NaN / Inf) */
};

/* tower family: which HF vision transformer the adapter wrapped */
enum { VLMFW_ARCH_SIGLIP = 0, VLMFW_ARCH_CLIP = 1, VLMFW_ARCH_EVA = 2 };
/* MLP activation (transformers/activations.py): gelu_pytorch_tanh, quick_gelu, gelu (erf) */
enum { VLMFW_ACT_GELU_TANH = 0, VLMFW_ACT_QUICK_GELU = 1, VLMFW_ACT_GELU_ERF = 2, VLMFW_ACT_NONE = 3 };
/* pooling done by the reference adapter after the tower */
enum {
    VLMFW_POOL_LN_MEAN = 0, /* post_layernorm on all tokens then mean(dim=1)   (paligemma.py:111-112) */
    VLMFW_POOL_MEAN = 1,    /* mean(dim=1) over raw last_hidden_state incl CLS (llava.py:116-117)     */
    VLMFW_POOL_CLS = 2,     /* raw last_hidden_state[:,0]                      (llava.py:114-115)     */
    VLMFW_POOL_LN_CLS_LN = 3 /* pooler_output = LN(LN(h)[:,0])                 (blip2.py:94-95)        */
};
/* pixel input formats */
enum {
    VLMFW_PIX_U8_NHWC = 0, /* raw RGB bytes [B,H,W,3]; rescale 1/255 and (x-mean)/std are fused on the GPU */
    VLMFW_PIX_F32_NCHW = 1 /* HF processor output `pixel_values` [B,3,H,W], already normalised */
};
/* element types of kernel operands / outputs */
enum { VLMFW_DT_BF16 = 0, VLMFW_DT_F16 = 1, VLMFW_DT_F32 = 2 };
/* 16-bit format of the tensor-core operands of a tower (weights AND activations: one tcgen05.mma kind::f16 takes A and
 * B in the same format — mixing raises an illegal-instruction fault on B200). Same tensor pipe, same peak.
 *   F16  (default): 11-bit mantissa. Weight rounding is the dominant error term of this path because it is coherent
 *         across tokens and survives mean pooling (DESIGN.md "numerics": bf16 weights alone put the probe-logit error at
 *         1.5e-2, above the 1e-2 tolerance; fp16 gives 1.2e-3). It is also what the reference itself runs on a GPU
 *         (fp16 autocast, probing/test/base_tester.py:50). LayerNorm statistics, softmax and all accumulation stay fp32;
 *         the residual stream itself is kept in fp16 (it is at once the next GEMM's A operand; measured cost 3e-4 of
 *         logit error, gain 4.6 % of the step — VLMFW_RESIDUAL=32 in the environment keeps an fp32 stream);
 *         16-bit stores saturate to +-65504 instead of overflowing to inf, and every clamped store is counted
 *         (vlmfw_tower_saturation). A weight matrix whose largest entry exceeds the fp16 range is stored divided by a
 *         power of two and its GEMM's accumulator multiplied back (exact), so the mode never depends on weight range.
 *   BF16: 8-bit mantissa, fp32 range, fp32 residual stream. Weight rounding alone costs 1.5e-2 of probe-logit error: this
 *         mode does NOT meet the 1e-2 logit tolerance (cosine >= 0.999 and the labels hold); it exists for A/B measurements
 *         and for callers that ask for bf16 explicitly. */
enum { VLMFW_CDT_F16 = 0, VLMFW_CDT_BF16 = 1 };
/* GEMM epilogues (kernel-level entry point, used by the parity tests) */
enum {
    VLMFW_EPI_BIAS_BF16 = 0,     /* out_bf16 = act(acc + bias)                         */
    VLMFW_EPI_BIAS_RESID_F32 = 1, /* out_f32  = acc + bias + resid_f32[row % resid_mod] */
    VLMFW_EPI_BIAS_RESID_16 = 2   /* out_16   = round16(acc + bias + resid_16): the 16-bit residual stream, `resid` points to
                                     16-bit data of the GEMM's dtype (ldr in elements), out may alias it; stats_out allowed
                                     (sums of the ROUNDED values), h16_out not (the output is the 16-bit copy) */
};

typedef struct vlmfw_tower_cfg {
    int32_t arch;       /* VLMFW_ARCH_*                                   */
    int32_t image_size; /* 224 / 336                                      */
    int32_t patch_size; /* 14                                             */
    int32_t hidden;     /* D                                              */
    int32_t heads;      /* H (head_dim = D / H, must be 64, 72 or 88)     */
    int32_t mlp;        /* I                                              */
    int32_t layers;     /* L                                              */
    int32_t has_cls;    /* class token prepended (CLIP, EVA)              */
    int32_t has_pre_ln; /* CLIP pre_layrnorm                              */
    int32_t conv_bias;  /* patch conv has a bias (not CLIP)               */
    int32_t act;        /* VLMFW_ACT_*                                    */
    int32_t pool;       /* VLMFW_POOL_*                                   */
    float ln_eps;
    float pix_mean[3]; /* used by VLMFW_PIX_U8_NHWC */
    float pix_std[3];
    int32_t compute_dtype; /* VLMFW_CDT_* */
} vlmfw_tower_cfg;

typedef struct vlmfw_tower vlmfw_tower;

/* one probe head: BatchNorm1d(D) -> Dropout -> Linear(D,C)              (base_probe.py:7-12)
 * deeper:         BatchNorm1d(D) -> Dropout -> Linear(D,Hd) -> GELU(erf) -> Dropout -> Linear(Hd,C)  (:14-22)
 * all pointers fp32 device memory; hidden == 0 selects the linear head (w2/b2 ignored). */
typedef struct vlmfw_head {
    const float* bn_weight; /* [D] */
    const float* bn_bias;   /* [D] */
    const float* bn_mean;   /* [D] running_mean */
    const float* bn_var;    /* [D] running_var  */
    const float* w1;        /* [C,D] or [Hd,D]  */
    const float* b1;        /* [C]   or [Hd]    */
    const float* w2;        /* [C,Hd] (deeper)  */
    const float* b2;        /* [C]    (deeper)  */
    int32_t classes;        /* C      */
    int32_t hidden;         /* Hd or 0 */
    float bn_eps;           /* 1e-5   */
    int32_t logit_offset;   /* column of this head's first logit in the packed [B, sum C] output */
} vlmfw_head;

int vlmfw_abi_version(void);
const char* vlmfw_last_error(void);
/* 0 if a sm_100 device is usable, else VLMFW_E_NODEVICE */
int vlmfw_device_ok(void);

/* ------------------------------------------------------------------ tower */
int vlmfw_tower_create(const vlmfw_tower_cfg* cfg, vlmfw_tower** out);
void vlmfw_tower_destroy(vlmfw_tower* t);
/* Canonical weight names (fp32 device source, any row-major contiguous shape with the stated element count):
 *   patch.w [D,3,p,p]  patch.b [D]  cls [D]  pos [S,D]  pre_ln.w/.b [D]  post_ln.w/.b [D]
 *   L<i>.ln1.w/.b  L<i>.qkv.w [3D,D]  L<i>.qkv.b [3D]  L<i>.o.w [D,D]  L<i>.o.b
 *   L<i>.ln2.w/.b  L<i>.fc1.w [I,D]   L<i>.fc1.b [I]   L<i>.fc2.w [D,I] L<i>.fc2.b [D]
 * Matrices are repacked to cfg.compute_dtype (patch K padded 588->640: whole 128-byte rows); vectors stay fp32. */
int vlmfw_tower_set_weight(vlmfw_tower* t, const char* name, const float* src, int64_t numel, void* stream);
/* 0 when every tensor the configuration needs has been set, else VLMFW_E_STATE (message names the first missing) */
int vlmfw_tower_ready(vlmfw_tower* t);
/* Load-time step after the last set_weight: builds the derived operands (LayerNorm folded into the QKV / fc1 matrices,
 * position-embedding tables). Synchronises `stream` (it reads back one scalar per folded matrix to pick the fp16 operand
 * scale), so call it outside CUDA-graph capture; vlmfw_tower_forward does it lazily when it has not been called. */
int vlmfw_tower_prepare(vlmfw_tower* t, void* stream);
/* Power-of-two scale the 16-bit operand of weight matrix `name` is stored divided by (1 unless max |W| > 32768). */
int vlmfw_tower_weight_scale(const vlmfw_tower* t, const char* name, float* host_scale);
/* Class-token-pooled towers (VLMFW_POOL_CLS: llava.py:114-115; VLMFW_POOL_LN_CLS_LN: blip2.py:94-95) run their LAST encoder
 * layer for the class-token rows only by default: K / V of every row, then Q, one-query attention, out-proj and the MLP as
 * M = B GEMMs over rows b*S. The pooled features are those of the full layer (same kernels' arithmetic on the rows that
 * matter; attention probabilities stay fp32); rows 1..S-1 of the final residual stream keep the previous layer's values.
 * enable = 0 computes the whole last layer (callers that read vlmfw_tower_hidden_ptr); VLMFW_CLS_TAIL=0 in the environment
 * sets that default at creation. Enabling it on a tower whose pooling reads other rows fails with VLMFW_E_INVALID. */
int vlmfw_tower_set_cls_tail(vlmfw_tower* t, int32_t enable);
int32_t vlmfw_tower_cls_tail(const vlmfw_tower* t);

/* fp16 towers: number of 16-bit activation stores (one thread's 32-column chunk of one row of the QKV / fc1 outputs or of
 * the residual stream) in which at least one value was clamped to +-65504 since creation / the last reset. 0 means no
 * forward clipped. Synchronises `stream`. bf16 towers always report 0 (fp32 exponent range). */
int vlmfw_tower_saturation(vlmfw_tower* t, int64_t* host_count, int32_t reset, void* stream);
size_t vlmfw_tower_workspace_bytes(const vlmfw_tower* t, int32_t batch);
int32_t vlmfw_tower_tokens(const vlmfw_tower* t); /* S */
/* pixels -> pooled features [B, D] fp32 (what VisionBackbone.forward returns) */
int vlmfw_tower_forward(vlmfw_tower* t, const void* pixels, int32_t pix_fmt, int32_t batch, void* workspace,
                        size_t workspace_bytes, float* pooled_out, void* stream);
/* same, but also exposes the final hidden state [B*S, D] fp32 inside the workspace (for parity tests) */
const float* vlmfw_tower_hidden_ptr(const vlmfw_tower* t, void* workspace);
/* fp16 towers keep the residual stream in fp16 (the final hidden state then lives in the 16-bit buffer this returns, [B*S, D]
 * of the tower's compute dtype); NULL when the tower keeps the fp32 stream (bf16 towers, VLMFW_RESIDUAL=32), in which case
 * vlmfw_tower_hidden_ptr is the final hidden state. */
const void* vlmfw_tower_hidden16_ptr(const vlmfw_tower* t, void* workspace, int32_t batch);
/* number of kernels the last forward launched (bench.py's gpu_launches claim) */
int64_t vlmfw_tower_last_launches(const vlmfw_tower* t);
/* Host-buffer convenience for non-torch hosts: H2D copy, forward, D2H copy, synchronous.
 * host_pixels: VLMFW_PIX_U8_NHWC bytes; host_pooled: [B, D] floats. Buffers are owned by the handle. */
int vlmfw_tower_forward_host(vlmfw_tower* t, const uint8_t* host_pixels, int32_t batch, float* host_pooled);

/* ------------------------------------------------------------------ image ingest
 * Replaces the CPU image-processor call in front of every tower forward (reference models/paligemma.py:108,
 * models/llava.py:109, models/blip2.py:91 -> transformers TorchvisionBackend.resize / center_crop on uint8 tensors):
 * raw RGB images of any size -> uint8 [B, S, S, 3], bit-exact with that processor's resized image (bicubic, antialias,
 * ATen's fixed-point separable kernel). The rescale + normalize that follows is fused into the tower's patch gather
 * (VLMFW_PIX_U8_NHWC). */
enum {
    VLMFW_RESIZE_SQUARE = 0,       /* resize to S x S ignoring aspect (SigLIP, BLIP image processors)              */
    VLMFW_RESIZE_SHORTEST_CROP = 1 /* shortest edge -> S (long = int(S * long / short)), centre crop S x S (CLIP) */
};
typedef struct vlmfw_preproc vlmfw_preproc;
int vlmfw_preproc_create(int32_t mode, int32_t out_size, vlmfw_preproc** out);
void vlmfw_preproc_destroy(vlmfw_preproc* p);
/* device workspace a batch with these image sizes needs (descriptor + weight tables + horizontal-pass rows) */
size_t vlmfw_preproc_workspace_bytes(vlmfw_preproc* p, const int32_t* heights_host, const int32_t* widths_host,
                                     int32_t batch);
/* packed: DEVICE bytes (16-byte aligned), image i = RGB HWC at packed + offsets_host[i], heights/widths/offsets HOST
 * arrays of `batch` entries. out: DEVICE uint8 [batch, S, S, 3]. Two launches per call, asynchronous on `stream`. */
int vlmfw_preproc_run(vlmfw_preproc* p, const uint8_t* packed, int64_t packed_bytes, const int64_t* offsets_host,
                      const int32_t* heights_host, const int32_t* widths_host, int32_t batch, uint8_t* out,
                      void* workspace, size_t workspace_bytes, void* stream);
int64_t vlmfw_preproc_last_launches(const vlmfw_preproc* p);
/* HOST-only (no device needed): tap windows and int16 weights of output positions [offset, offset + count) of the
 * in_size -> out_full antialiased bicubic resize; weights [count, weights_ld], *taps <= weights_ld. */
int vlmfw_resize_weights(int32_t in_size, int32_t out_full, int32_t offset, int32_t count, int32_t* xmin,
                         int32_t* xsize, int16_t* weights, int32_t weights_ld, int32_t* taps, int32_t* precision);

/* ------------------------------------------------------------------ probe */
/* pooled [B,D] -> packed logits [B, sum C] (+ per-head argmax [B, n_heads], first-max tie rule of torch.argmax).
 * `heads` is a HOST array of n_heads descriptors. */
int vlmfw_probe_forward(const float* pooled, int32_t batch, int32_t dim, const vlmfw_head* heads, int32_t n_heads,
                        float* logits_out, int32_t logits_ld, int32_t* argmax_out, void* stream);

/* Probe train step on cached features, linear head, split so a data-parallel caller can all-reduce between phases:
 *  1. bn_stats     : stats[0:D] = sum_b x, stats[D:2D] = sum_b x^2 over the local batch, fp64      (all-reduce SUM)
 *  2. linear_grads : forward BN(train)+dropout+Linear+weighted CE(ignore_index=-1) and backward.
 *                    bucket = [dW (C*D) | db (C) | dgamma (D) | dbeta (D) | loss_sum | weight_sum], UNNORMALISED
 *                    sums over the local batch                                                     (all-reduce SUM)
 *  3. adamw        : params -= lr * (m_hat / (sqrt(v_hat)+eps) + wd * params) with grad = bucket / weight_sum;
 *                    also updates BN running stats (momentum 0.1, unbiased var) from the global stats.
 * `n_global` is the number of rows behind `stats` (global batch). keep_mask: uint8 [B,D] or NULL (no dropout). */
int vlmfw_probe_bn_stats(const float* x, int32_t batch, int32_t dim, double* stats, void* stream);
/* scratch: fp32 [batch * (classes + 2)]; logits_out: fp32 [batch, classes] */
int vlmfw_probe_linear_grads(const float* x, const int64_t* targets, int32_t batch, int32_t dim, int32_t classes,
                             const double* stats, int32_t n_global, float bn_eps, const float* gamma,
                             const float* beta, const float* weight, const float* bias, const float* class_weight,
                             const uint8_t* keep_mask, float dropout_p, float* bucket, float* logits_out,
                             float* scratch, void* stream);
int vlmfw_probe_adamw(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                      int64_t weight_sum_index, int32_t step, float lr, float beta1, float beta2, float eps,
                      float weight_decay, void* stream);
/* Deeper head (base_probe.py:14-22: BatchNorm1d -> Dropout -> Linear(D,Hd) -> GELU(erf) -> Dropout -> Linear(Hd,C)) and the
 * multitask step (multitask_trainer.py:188-263): same phase contract as the linear head. `params` is the head's flat
 * parameter vector [W1 (Hd*D) | b1 (Hd) | W2 (C*Hd) | b2 (C) | gamma (D) | beta (D)], `bucket` has that layout followed by
 * [loss_sum, weight_sum], all un-normalised sums over the local rows. keep1 uint8 [B,D], keep2 uint8 [B,Hd] (or NULL).
 * scratch: vlmfw_probe_deep_scratch_floats(...) floats. A multitask step calls this (or vlmfw_probe_linear_grads) once per
 * head on the shared statistics, all-reduces the concatenated buckets once, then updates every head with
 * vlmfw_probe_adamw_ex, whose gradient is grad_scale * bucket / weight_sum (grad_scale = the task's loss weight) and which
 * leaves a head untouched when no row of the global batch carried its label. */
size_t vlmfw_probe_deep_scratch_floats(int32_t batch, int32_t dim, int32_t hidden, int32_t classes);
int vlmfw_probe_deep_grads(const float* x, const int64_t* targets, int32_t batch, int32_t dim, int32_t hidden,
                           int32_t classes, const double* stats, int32_t n_global, float bn_eps, const float* params,
                           const float* class_weight, const uint8_t* keep1, const uint8_t* keep2, float dropout_p,
                           float* bucket, float* logits_out, float* scratch, void* stream);
int vlmfw_probe_adamw_ex(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                         int64_t weight_sum_index, float grad_scale, int32_t step, float lr, float beta1, float beta2,
                         float eps, float weight_decay, void* stream);
/* the same update with the step number (*steps_done = optimisation steps completed so far) and the task weight (*grad_scale)
 * read from device memory: nothing in the launch changes between steps, so a multitask step replays from one CUDA graph */
int vlmfw_probe_adamw_ex_dev(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                             int64_t weight_sum_index, const float* grad_scale, const int64_t* steps_done, float lr,
                             float beta1, float beta2, float eps, float weight_decay, void* stream);
int vlmfw_probe_bn_running_update(float* running_mean, float* running_var, const double* stats, int32_t dim,
                                  int32_t n_global, float momentum, void* stream);
/* Dropout keep mask of the train-mode heads (nn.Dropout(p) in make_head / make_head_deeper,
 * probing/models/base_probe.py:7-22), drawn on the device: keep [rows, cols] uint8 (16-byte aligned), 1 = kept.
 * keep[r, c] is a function of (key, row0 + r, c) only — a counter-based generator — so the rank that owns global rows
 * [row0, row0 + rows) draws exactly its slice of the single-process mask; the caller folds seed, step and task into key.
 * P(keep) = 1 - dropout_p. The result feeds keep_mask / keep1 / keep2 of the *_grads entry points above. */
int vlmfw_dropout_keep_mask(uint8_t* keep, int32_t rows, int32_t cols, int64_t row0, uint64_t key, float dropout_p,
                            void* stream);
/* Graph-replayable forms: the only things that change from one optimisation step to the next — AdamW's step number
 * (bias corrections) and the per-step part of the dropout key — are read from a device-side counter, *steps_done =
 * optimisation steps completed so far, which vlmfw_counter_add advances at the end of the step. With them one train step
 * (probing/train/base_trainer.py:166-181) is a fixed launch sequence and replays from ONE CUDA graph, all-reduces and
 * the loss read-back included. vlmfw_probe_adamw_dev performs step number *steps_done + 1; vlmfw_dropout_keep_mask_dev
 * uses the key  key + *steps_done * step_stride. */
int vlmfw_probe_adamw_dev(float* params, float* exp_avg, float* exp_avg_sq, const float* bucket, int64_t numel,
                          int64_t weight_sum_index, const int64_t* steps_done, float lr, float beta1, float beta2,
                          float eps, float weight_decay, void* stream);
int vlmfw_dropout_keep_mask_dev(uint8_t* keep, int32_t rows, int32_t cols, int64_t row0, uint64_t key,
                                const int64_t* steps_done, uint64_t step_stride, float dropout_p, void* stream);
int vlmfw_counter_add(int64_t* counter, int64_t inc, void* stream);

/* Data-parallel exchange of the train step over NVLink peer memory, fused into the step's own kernels (BASELINE configs[4]:
 * "only the tiny probe-weight gradients are allreduced"; the reference itself has no multi-GPU training —
 * probing/train/base_trainer.py:30). One process per GPU of ONE node. Each rank creates a communicator (one exchange
 * buffer: per-rank slots for the BatchNorm sums [2D] fp64 and the gradient bucket [C*D + C + 2D + 2] fp32), the host
 * exchanges the handles (vlmfw_comm_handle -> all-gather -> vlmfw_comm_connect maps every peer's buffer with cudaIpc), and
 * vlmfw_probe_linear_step_dp then runs whole optimisation steps: the producing kernels store their rank's sums straight into
 * every peer's slots, every value together with the flag of this exchange in ONE 8-byte (fp32) / 16-byte (fp64) store; the
 * consuming kernels poll their LOCAL slots until each carries the flag and add them in rank order (bit-identical parameters
 * on all ranks). No fence, no separate flag, no collective kernel. world == 1 needs no connect. */
typedef struct vlmfw_comm vlmfw_comm;
int vlmfw_comm_create(int32_t rank, int32_t world, int64_t stats_doubles, int64_t bucket_floats, vlmfw_comm** out);
int32_t vlmfw_comm_handle_bytes(void);
int vlmfw_comm_handle(vlmfw_comm* c, void* handle_out);
int vlmfw_comm_connect(vlmfw_comm* c, const void* handles);
void vlmfw_comm_destroy(vlmfw_comm* c);
/* *host_flag = 1 when a wait gave up because a peer stopped reporting (results since then are invalid) */
int vlmfw_comm_error(vlmfw_comm* c, int32_t* host_flag);
/* one optimisation step of the linear head (probing/train/singletask_trainer.py:183-191 + base_trainer.py:166-181) as seven
 * launches: bn_stats, dropout mask (+ combine of the ranks' BatchNorm sums), forward, backward, AdamW (+ combine of the ranks'
 * gradient buckets), BatchNorm running statistics, step counter. params / moments are the flat vector
 * [W (C*D) | b (C) | gamma (D) | beta (D)]; bucket [n_params + 2] receives the all-reduced [loss_sum, weight_sum] in its
 * last two entries; keep uint8 [batch, dim] (16-byte aligned; unused when dropout_p == 0); scratch fp32 [batch * (classes + 2)];
 * *steps_done = optimisation steps completed so far (advanced by this call's last kernel). */
typedef struct vlmfw_probe_step {
    const float* x; const int64_t* targets; int32_t batch, dim, classes, n_global; int64_t row0;
    double* stats; float* params; float* exp_avg; float* exp_avg_sq; float* bucket; float* logits; float* scratch;
    uint8_t* keep; const float* class_weight; float* running_mean; float* running_var; int64_t* steps_done;
    uint64_t mask_key, mask_step_stride;
    float dropout_p, bn_eps, bn_momentum, lr, beta1, beta2, eps, weight_decay;
} vlmfw_probe_step;
int vlmfw_probe_linear_step_dp(const vlmfw_probe_step* a, vlmfw_comm* c, void* stream);

/* ------------------------------------------------------------------ tester sink
 * Per-task confusion counts of a batch of predictions, accumulated on the device: counts [tasks, classes, classes] int32
 * (row = ground truth, column = prediction — sklearn.metrics.confusion_matrix as the reference's Evaluator calls it,
 * datasets_vlm/evaluate_dataset.py:74-97). pred int32 [rows, tasks] (vlmfw_probe_forward's argmax_out), gt int64
 * [rows, tasks] with -1 = label missing (skipped, evaluate_dataset.py:78); *dropped counts labels outside [0, classes). */
int vlmfw_confusion_update(const int32_t* pred, const int64_t* gt, int32_t rows, int32_t tasks, int32_t classes,
                           int32_t* counts, int32_t* dropped, void* stream);

/* ----------------------------------------------------- kernel-level entry points (parity tests, profiling) */
/* out = epilogue(A[M,K] x W[N,K]^T), fp32 accumulation in TMEM. A (row stride lda) and W (row stride ldw) are both
 * `dtype` (VLMFW_DT_BF16 or VLMFW_DT_F16). VLMFW_EPI_BIAS_BF16 writes act(acc + bias) in that same 16-bit dtype;
 * VLMFW_EPI_BIAS_RESID_F32 writes fp32 acc + bias + resid.
 * Row mapping of the output: out_row = (m / grp) * grp_out + (m % grp) + grp_off  (grp <= 0: out_row = m).
 * resid row = m % resid_mod (resid_mod <= 0: m). K, lda, ldw multiples of 8; N, ldo multiples of 8.
 * block_n: N tile of the kernel (128 / 192 / 256), 0 = choose the one that wastes the fewest MMA columns; OR-ed with
 * VLMFW_GEMM_FORCE_1CTA / VLMFW_GEMM_FORCE_2CTA to pin the single-CTA (128-row tiles) or CTA-pair (tcgen05
 * cta_group::2, 256-row tiles) form of the kernel — by default pairs are used whenever M > 128. */
#define VLMFW_GEMM_FORCE_1CTA 0x1000
#define VLMFW_GEMM_FORCE_2CTA 0x2000
/* test hook: use the transpose-and-store epilogue (normally only taken for grp / resid_mod) instead of the TMA one */
#define VLMFW_GEMM_FORCE_GENERIC_EPI 0x4000
int vlmfw_gemm_tn(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t dtype, int32_t m, int32_t n,
                  int32_t k, const float* bias, int32_t epi, int32_t act, void* out, int32_t ldo, const float* resid,
                  int32_t ldr, int32_t resid_mod, int32_t grp, int32_t grp_out, int32_t grp_off, int32_t block_n,
                  void* stream);
/* The same kernel with LayerNorm folded around it, as the tower chains it: no LayerNorm kernel runs
 * between GEMMs.
 *  producer side (epi = VLMFW_EPI_BIAS_RESID_F32): besides out = acc + bias + resid, the epilogue stores the 16-bit cast
 *    of the updated rows to h16_out [M, ld_h16] and, per row, `vlmfw_gemm_stats_parts(n, block_n)` float2 partials
 *    (sum, sum of squares) to stats_out [M, parts];
 *  consumer side (epi = VLMFW_EPI_BIAS_BF16): A is that plain 16-bit cast, W = round16(gamma . W) with every row
 *    centred (the mean subtraction of LayerNorm is the projection I - 11^T/K, moved onto the weights) and
 *    out = act(rstd[m] acc + bias[n]) with rstd from ln_stats [M, ln_np] over ln_dim columns;
 *    W and bias = beta W^T + b come from vlmfw_ln_fold_weight.
 * Fields left 0 / NULL switch the corresponding part off. */
typedef struct vlmfw_gemm_desc {
    const void* a; int32_t lda; const void* w; int32_t ldw; int32_t dtype; int32_t m, n, k;
    const float* bias; int32_t epi, act; void* out; int32_t ldo; const float* resid; int32_t ldr; int32_t block_n;
    const void* ln_stats; int32_t ln_np, ln_dim; float ln_eps;
    void* h16_out; int32_t ld_h16; void* stats_out;
    float acc_scale;  /* accumulator scale (W holds W_true / acc_scale); 0 is read as 1 */
    void* sat_count;  /* device uint64 counter of clamped fp16 stores, or NULL */
} vlmfw_gemm_desc;
int vlmfw_gemm_tn_ex(const vlmfw_gemm_desc* desc, void* stream);
int32_t vlmfw_gemm_stats_parts(int32_t n, int32_t block_n);
/* x fp32 [rows, dim] -> out16 = 16-bit cast (dtype), stats [rows, parts] float2 with (sum, sum of squares) in slot 0 */
int vlmfw_row_stats_cast(const float* x, int32_t rows, int32_t dim, void* out16, int32_t dtype, void* stats,
                         int32_t parts, void* stream);
/* w fp32 [n, k] -> w16 [n, ld] = round16(gamma[k] w[n,k] - mean_j(gamma[j] w[n,j])); c2[n] = sum_k beta[k] w[n,k] + bias[n] */
int vlmfw_ln_fold_weight(const float* w, int32_t n, int32_t k, const float* gamma, const float* beta, const float* bias,
                         void* w16, int32_t ld, int32_t dtype, float* c2, void* stream);
/* y = LayerNorm(x) * gamma + beta over rows of x [rows, dim] fp32; out_dtype VLMFW_DT_* (fp32 output may alias x) */
int vlmfw_layernorm(const float* x, int32_t rows, int32_t dim, const float* gamma, const float* beta, float eps,
                    void* out, int32_t out_dtype, void* stream);
/* softmax(Q K^T / sqrt(dh)) V for every (image, head); qkv 16-bit (`dtype`) [B*S, 3*D] = [q | k | v], heads contiguous
 * inside each third; out same dtype [B*S, D]. head_dim in {64, 72, 88} (or any multiple of 8 <= 96). */
int vlmfw_attention(const void* qkv, int32_t dtype, int32_t batch, int32_t tokens, int32_t heads, int32_t head_dim,
                    void* out, void* stream);
/* attention of query row 0 of every image over all `tokens` keys (the last layer of a class-token-pooled tower):
 * out 16-bit [batch, ld_out] compact, head-major columns like vlmfw_attention */
int vlmfw_attention_cls(const void* qkv, int32_t dtype, int32_t batch, int32_t tokens, int32_t heads, int32_t head_dim,
                        void* out, int32_t ld_out, void* stream);
/* patch gather: pixels -> A [B*P, kpad] (out_dtype VLMFW_DT_*) with k = c*p*p + ky*p + kx (conv weight flattening),
 * zero padded */
int vlmfw_im2col(const void* pixels, int32_t pix_fmt, int32_t batch, int32_t image_size, int32_t patch_size,
                 const float* mean, const float* std, void* out, int32_t kpad, int32_t out_dtype, void* stream);
/* pooling epilogue over the final hidden state h [B*S, D] fp32 */
int vlmfw_pool(const float* h, int32_t batch, int32_t tokens, int32_t dim, int32_t pool, const float* gamma,
               const float* beta, float eps, float* pooled_out, void* stream);
/* dst_16[rows, dst_ld] = round_to_nearest_even(src_f32[rows, cols]) as VLMFW_DT_* (zero pad cols..dst_ld) */
int vlmfw_pack_16(const float* src, int64_t rows, int64_t cols, void* dst, int64_t dst_ld, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLMFW_B200_H */
