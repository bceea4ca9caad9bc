// Tower runtime: owns the packed weights of one vision tower and runs its forward as a fixed kernel sequence on the
// caller's stream. This is the native replacement for `self.vision_model(pixel_values)` + the adapter pooling in
//   reference models/paligemma.py:108-112, models/llava.py:109-117, models/blip2.py:91-95
// i.e. hf SiglipVisionTransformer.forward (modeling_siglip.py:604-625), CLIPVisionTransformer.forward
// (modeling_clip.py:665-692) and Blip2VisionModel.forward (modeling_blip_2.py:499-533).
//
// HBM layout per forward (M = B*S tokens), all inside one caller-provided workspace:
//   h    fp32 [M, D]    residual stream — kept in fp32 end to end (bf16 residuals miss the
//                       1e-2 logit tolerance); read by LayerNorm, read-modify-written by the out-proj / fc2 epilogues
//   xn   16b  [M, D]    plain 16-bit cast of h = A operand of the QKV / fc1 GEMMs (16b = fp16 or bf16,
//                       cfg.compute_dtype); written by the out-proj / fc2 epilogues, never by a LayerNorm kernel
//   stats f32 [M, NP, 2] per-row (sum, sum of squares) partials of h from the same epilogues: LayerNorm itself is
//                       folded into the QKV / fc1 GEMMs (gamma in the weights, mean / rstd applied in their epilogue)
//   qkv  16b  [M, 3D]   fused projection, consumed in place by the attention kernel (no head permute)
//   ctx  16b  [M, D]    attention output = A operand of out-proj
//   mid  16b  [M, I]    GELU(fc1) = A operand of fc2; the patch-gather matrix [B*P, 592] aliases it before layer 0
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace vlmfw {

namespace {
thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
int64_t& launch_counter() { return g_launches; }

namespace {
struct DevInfo {
    int rc = VLMFW_E_NODEVICE;
    int sms = 0;
    char msg[256] = "";
};
// One record per device ordinal: a process may drive several GPUs (one TowerEngine per device), and both the
// architecture check and the SM count belong to the CURRENT device of the calling thread.
constexpr int kMaxDevices = 64;
DevInfo probe_device(int dev) {
    DevInfo d;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        snprintf(d.msg, sizeof(d.msg), "cudaGetDeviceProperties(%d) failed", dev);
        cudaGetLastError();
        return d;
    }
    if (p.major != 10) {
        snprintf(d.msg, sizeof(d.msg), "device %d (%.160s) is sm_%d%d; vlmfw_b200 kernels are sm_100a only", dev, p.name,
                 p.major, p.minor);
        return d;
    }
    d.rc = VLMFW_OK;
    d.sms = p.multiProcessorCount;
    return d;
}
const DevInfo& dev_info() {
    static DevInfo none = []() {
        DevInfo d;
        snprintf(d.msg, sizeof(d.msg), "no CUDA device; vlmfw_b200 has no CPU fallback");
        return d;
    }();
    static DevInfo table[kMaxDevices];
    static std::once_flag once[kMaxDevices];
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(none.msg, sizeof(none.msg), "no CUDA device (%s); vlmfw_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return none;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        cudaGetLastError();
        return none;
    }
    std::call_once(once[dev], [dev]() { table[dev] = probe_device(dev); });
    return table[dev];
}
}  // namespace

int device_check() {
    const DevInfo& d = dev_info();
    if (d.rc != VLMFW_OK) set_error("%s", d.msg);
    return d.rc;
}
int num_sms() { return dev_info().sms > 0 ? dev_info().sms : 148; }

}  // namespace vlmfw

using namespace vlmfw;

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Slot {
    void** dst = nullptr;
    float** master = nullptr;  // matrices a LayerNorm is folded into keep their fp32 values: the 16-bit operand is
                               // rebuilt from (W, gamma, beta, bias) whenever any of them changes (build_derived)
    bool matrix = false;  // repack to fp16 / bf16 [rows, ld]
    float* scale = nullptr;  // matrices: where the power-of-two operand scale of this matrix lives (acc_scale of its GEMM)
    int64_t rows = 0, cols = 0, ld = 0;
    bool required = true;
    bool set = false;
};

struct LayerW {
    float *ln1w = nullptr, *ln1b = nullptr, *ln2w = nullptr, *ln2b = nullptr;
    float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
    void *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // wqkv / w1 hold gamma-folded weights
    float *wqkv32 = nullptr, *w132 = nullptr;                           // fp32 masters of the folded matrices
    float *c2qkv = nullptr, *c2fc1 = nullptr;  // derived: beta W^T + b (ln_fold_weight_kernel)
    // fp16 towers: the 16-bit operand holds W / s (s a power of two, 1 unless max |W| exceeds the fp16 range) and the GEMM
    // epilogue multiplies its accumulator by s — a weight beyond +-65504 no longer forces the whole tower into bf16
    float s_qkv = 1.f, s_o = 1.f, s_fc1 = 1.f, s_fc2 = 1.f;
};

}  // namespace

struct vlmfw_tower {
    vlmfw_tower_cfg cfg{};
    int P = 0, S = 0, D = 0, H = 0, dh = 0, I = 0, L = 0, kreal = 0, kpad = 0;
    std::vector<LayerW> layers;
    void* patch_w = nullptr;  // 16-bit [D, kpad]
    float s_patch = 1.f;
    unsigned long long* sat_dev = nullptr;  // fp16 towers: device counter of clamped 16-bit stores (see vlmfw_tower_saturation)
    unsigned int* amax_dev = nullptr;       // load-time scratch of build_derived
    bool f16 = true;  // cfg.compute_dtype == VLMFW_CDT_F16
    // fp16 towers keep the residual stream itself in fp16 (xn is the stream): half the HBM traffic of the
    // out-proj / fc2 epilogues, which put out-proj on the HBM roofline with an fp32 stream. bf16 towers (8-bit mantissa)
    // and VLMFW_RESIDUAL=32 keep the fp32 stream.
    bool resid16 = false;
    // Class-token-pooled towers (VLMFW_POOL_CLS, VLMFW_POOL_LN_CLS_LN): the adapter keeps row 0 of every image only, so the
    // LAST encoder layer needs K and V of all rows but Q, attention, out-proj and the MLP of the class-token rows alone:
    // GEMMs with M = B over rows b*S (row stride S*D), one-query attention. Same result for the pooled features; the other
    // rows of the final residual stream are left as the last layer found them (vlmfw_tower_set_cls_tail(t, 0) computes them).
    bool cls_tail = false;
    float* patch_b = nullptr;
    float* cls = nullptr;
    float* pos = nullptr;  // [S, D]
    float *pre_w = nullptr, *pre_b = nullptr, *post_w = nullptr, *post_b = nullptr;
    float* posb = nullptr;      // derived: pos[has_cls:, :] + patch_b   [P, D]
    float* cls_pos0 = nullptr;  // derived: cls + pos[0]                 [D]
    bool derived_ok = false;
    std::unordered_map<std::string, Slot> slots;
    int64_t last_launches = 0;
    // host-buffer path
    cudaStream_t own_stream = nullptr;
    void* host_pix_dev = nullptr;
    size_t host_pix_bytes = 0;
    void* host_ws = nullptr;
    size_t host_ws_bytes = 0;
    float* host_pooled_dev = nullptr;
    size_t host_pooled_bytes = 0;

    void add_slot(const std::string& name, void** dst, bool matrix, int64_t rows, int64_t cols, int64_t ld,
                  bool required = true, float** master = nullptr, float* scale = nullptr) {
        Slot s;
        s.dst = dst;
        s.master = master;
        s.scale = scale;
        s.matrix = matrix;
        s.rows = rows;
        s.cols = cols;
        s.ld = ld;
        s.required = required;
        slots.emplace(name, s);
    }
    size_t off_h() const { return 0; }
    size_t M(int B) const { return static_cast<size_t>(B) * S; }
    size_t off_xn(int B) const { return align_up(M(B) * D * 4, 1024); }
    size_t off_qkv(int B) const { return off_xn(B) + align_up(M(B) * D * 2, 1024); }
    size_t off_ctx(int B) const { return off_qkv(B) + align_up(M(B) * 3 * D * 2, 1024); }
    // Row pitch of `mid` and of the fc2 weight in elements: I rounded up to whole 128-byte lines. SigLIP's I = 4304 gives
    // 8608-byte rows, so every 128-byte TMA box row of the fc2 operands straddled two L2 lines: the operand loads ALONE
    // took 0.394 ms (as long as the MMAs) against 0.204 ms for CLIP's line-aligned K = 4096 of similar size.
    int Ild() const { return static_cast<int>(align_up(static_cast<size_t>(I), 64)); }
    size_t off_mid(int B) const { return off_ctx(B) + align_up(M(B) * D * 2, 1024); }
    size_t off_stats(int B) const {
        size_t mid = M(B) * static_cast<size_t>(Ild()) * 2;
        size_t pa = static_cast<size_t>(B) * P * kpad * 2;
        return off_mid(B) + align_up(mid > pa ? mid : pa, 1024);
    }
    int stats_parts() const { return gemm_stats_parts(D, 0); }
    size_t ws_bytes(int B) const { return off_stats(B) + align_up(M(B) * stats_parts() * 8, 1024); }
};

extern "C" int vlmfw_abi_version(void) { return VLMFW_ABI_VERSION; }
extern "C" const char* vlmfw_last_error(void) { return get_error(); }
extern "C" int vlmfw_device_ok(void) { return device_check(); }

extern "C" int vlmfw_tower_create(const vlmfw_tower_cfg* cfg, vlmfw_tower** out) {
    VLMFW_REQUIRE(cfg && out, "tower_create: null argument");
    VLMFW_REQUIRE(cfg->arch >= VLMFW_ARCH_SIGLIP && cfg->arch <= VLMFW_ARCH_EVA, "tower_create: unknown arch %d", cfg->arch);
    VLMFW_REQUIRE(cfg->patch_size > 0 && cfg->image_size > 0 && cfg->image_size % cfg->patch_size == 0,
                  "tower_create: image_size %d must be a multiple of patch_size %d", cfg->image_size, cfg->patch_size);
    VLMFW_REQUIRE(cfg->hidden > 0 && cfg->heads > 0 && cfg->hidden % cfg->heads == 0, "tower_create: hidden %d / heads %d",
                  cfg->hidden, cfg->heads);
    const int dh = cfg->hidden / cfg->heads;
    VLMFW_REQUIRE(dh % 8 == 0 && dh <= 96, "tower_create: head_dim %d unsupported (multiple of 8, <= 96)", dh);
    VLMFW_REQUIRE(cfg->hidden % 8 == 0 && cfg->hidden <= 1536, "tower_create: hidden %d unsupported", cfg->hidden);
    VLMFW_REQUIRE(cfg->mlp > 0 && cfg->mlp % 8 == 0, "tower_create: mlp %d must be a multiple of 8", cfg->mlp);
    VLMFW_REQUIRE(cfg->layers > 0, "tower_create: layers %d", cfg->layers);
    VLMFW_REQUIRE(cfg->act >= VLMFW_ACT_GELU_TANH && cfg->act <= VLMFW_ACT_GELU_ERF, "tower_create: act %d", cfg->act);
    VLMFW_REQUIRE(cfg->pool >= VLMFW_POOL_LN_MEAN && cfg->pool <= VLMFW_POOL_LN_CLS_LN, "tower_create: pool %d", cfg->pool);
    VLMFW_REQUIRE(cfg->compute_dtype == VLMFW_CDT_F16 || cfg->compute_dtype == VLMFW_CDT_BF16,
                  "tower_create: compute_dtype %d", cfg->compute_dtype);
    if (cfg->pool == VLMFW_POOL_CLS || cfg->pool == VLMFW_POOL_LN_CLS_LN)
        VLMFW_REQUIRE(cfg->has_cls, "tower_create: CLS pooling needs has_cls");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;

    vlmfw_tower* t = new vlmfw_tower();
    t->cfg = *cfg;
    t->f16 = cfg->compute_dtype == VLMFW_CDT_F16;
    {
        const char* e = getenv("VLMFW_RESIDUAL");
        t->resid16 = t->f16 && !(e != nullptr && atoi(e) == 32);
        const char* c = getenv("VLMFW_CLS_TAIL");
        t->cls_tail = cfg->has_cls && (cfg->pool == VLMFW_POOL_CLS || cfg->pool == VLMFW_POOL_LN_CLS_LN) &&
                      !(c != nullptr && atoi(c) == 0);
    }
    const int g = cfg->image_size / cfg->patch_size;
    t->P = g * g;
    t->S = t->P + (cfg->has_cls ? 1 : 0);
    t->D = cfg->hidden;
    t->H = cfg->heads;
    t->dh = dh;
    t->I = cfg->mlp;
    t->L = cfg->layers;
    t->kreal = 3 * cfg->patch_size * cfg->patch_size;
    // K of the patch GEMM padded to whole 128-byte rows (588 -> 640): operand rows never straddle L2 lines
    t->kpad = static_cast<int>(align_up(t->kreal, 64));
    t->layers.resize(t->L);
    const int64_t D = t->D, I = t->I;
    t->add_slot("patch.w", &t->patch_w, true, D, t->kreal, t->kpad, true, nullptr, &t->s_patch);
    t->add_slot("patch.b", reinterpret_cast<void**>(&t->patch_b), false, 1, D, D, cfg->conv_bias != 0);
    t->add_slot("cls", reinterpret_cast<void**>(&t->cls), false, 1, D, D, cfg->has_cls != 0);
    t->add_slot("pos", reinterpret_cast<void**>(&t->pos), false, t->S, D, D);
    t->add_slot("pre_ln.w", reinterpret_cast<void**>(&t->pre_w), false, 1, D, D, cfg->has_pre_ln != 0);
    t->add_slot("pre_ln.b", reinterpret_cast<void**>(&t->pre_b), false, 1, D, D, cfg->has_pre_ln != 0);
    const bool need_post = cfg->pool == VLMFW_POOL_LN_MEAN || cfg->pool == VLMFW_POOL_LN_CLS_LN;
    t->add_slot("post_ln.w", reinterpret_cast<void**>(&t->post_w), false, 1, D, D, need_post);
    t->add_slot("post_ln.b", reinterpret_cast<void**>(&t->post_b), false, 1, D, D, need_post);
    for (int i = 0; i < t->L; ++i) {
        LayerW& w = t->layers[i];
        const std::string p = "L" + std::to_string(i) + ".";
        t->add_slot(p + "ln1.w", reinterpret_cast<void**>(&w.ln1w), false, 1, D, D);
        t->add_slot(p + "ln1.b", reinterpret_cast<void**>(&w.ln1b), false, 1, D, D);
        t->add_slot(p + "qkv.w", &w.wqkv, true, 3 * D, D, D, true, &w.wqkv32, &w.s_qkv);
        t->add_slot(p + "qkv.b", reinterpret_cast<void**>(&w.bqkv), false, 1, 3 * D, 3 * D);
        t->add_slot(p + "o.w", &w.wo, true, D, D, D, true, nullptr, &w.s_o);
        t->add_slot(p + "o.b", reinterpret_cast<void**>(&w.bo), false, 1, D, D);
        t->add_slot(p + "ln2.w", reinterpret_cast<void**>(&w.ln2w), false, 1, D, D);
        t->add_slot(p + "ln2.b", reinterpret_cast<void**>(&w.ln2b), false, 1, D, D);
        t->add_slot(p + "fc1.w", &w.w1, true, I, D, D, true, &w.w132, &w.s_fc1);
        t->add_slot(p + "fc1.b", reinterpret_cast<void**>(&w.b1), false, 1, I, I);
        t->add_slot(p + "fc2.w", &w.w2, true, D, I, t->Ild(), true, nullptr, &w.s_fc2);
        t->add_slot(p + "fc2.b", reinterpret_cast<void**>(&w.b2), false, 1, D, D);
    }
    if (cudaMalloc(reinterpret_cast<void**>(&t->sat_dev), sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&t->amax_dev), sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(t->sat_dev, 0, sizeof(unsigned long long)) != cudaSuccess) {
        set_error("tower_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        vlmfw_tower_destroy(t);
        return VLMFW_E_CUDA;
    }
    *out = t;
    return VLMFW_OK;
}

extern "C" void vlmfw_tower_destroy(vlmfw_tower* t) {
    if (!t) return;
    for (auto& kv : t->slots) {
        if (*kv.second.dst) cudaFree(*kv.second.dst);
        if (kv.second.master && *kv.second.master) cudaFree(*kv.second.master);
    }
    for (auto& w : t->layers)
        for (float* ptr : {w.c2qkv, w.c2fc1})
            if (ptr) cudaFree(ptr);
    if (t->sat_dev) cudaFree(t->sat_dev);
    if (t->amax_dev) cudaFree(t->amax_dev);
    if (t->posb) cudaFree(t->posb);
    if (t->cls_pos0) cudaFree(t->cls_pos0);
    if (t->host_pix_dev) cudaFree(t->host_pix_dev);
    if (t->host_ws) cudaFree(t->host_ws);
    if (t->host_pooled_dev) cudaFree(t->host_pooled_dev);
    if (t->own_stream) cudaStreamDestroy(t->own_stream);
    delete t;
}

extern "C" int vlmfw_tower_set_weight(vlmfw_tower* t, const char* name, const float* src, int64_t numel, void* stream) {
    VLMFW_REQUIRE(t && name && src, "tower_set_weight: null argument");
    auto it = t->slots.find(name);
    if (it == t->slots.end()) {
        set_error("tower_set_weight: unknown weight '%s'", name);
        return VLMFW_E_NOTFOUND;
    }
    Slot& s = it->second;
    VLMFW_REQUIRE(numel == s.rows * s.cols, "tower_set_weight: '%s' expects %lld elements, got %lld", name,
                  static_cast<long long>(s.rows * s.cols), static_cast<long long>(numel));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (*s.dst == nullptr) {
        const size_t bytes = s.matrix ? static_cast<size_t>(s.rows * s.ld) * 2 : static_cast<size_t>(numel) * 4;
        VLMFW_CUDA_OK(cudaMalloc(s.dst, bytes));
    }
    if (s.matrix) {
        float amax = 0.f;
        int rc = device_absmax(src, numel, &amax, st);
        if (rc != VLMFW_OK) return rc;
        if (!(amax < 3.0e38f)) {  // NaN / Inf: bit pattern above any finite value
            set_error("tower_set_weight: '%s' holds a non-finite value", name);
            return VLMFW_E_RANGE;
        }
        if (s.master != nullptr) {
            // LayerNorm-folded matrix: keep the fp32 values; build_derived() writes the 16-bit operand (and picks its scale
            // from the FOLDED values gamma . W - rowmean, which are what is rounded to 16 bits)
            if (*s.master == nullptr) VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(s.master), static_cast<size_t>(numel) * 4));
            VLMFW_CUDA_OK(cudaMemcpyAsync(*s.master, src, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
        } else {
            const float scale = t->f16 ? fp16_weight_scale(amax) : 1.f;
            if (s.scale != nullptr) *s.scale = scale;
            rc = launch_pack_16(src, s.rows, s.cols, *s.dst, s.ld, t->f16, st, 1.f / scale);
            if (rc != VLMFW_OK) return rc;
        }
    } else {
        VLMFW_CUDA_OK(cudaMemcpyAsync(*s.dst, src, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
    }
    s.set = true;
    t->derived_ok = false;
    return VLMFW_OK;
}

extern "C" int vlmfw_tower_ready(vlmfw_tower* t) {
    VLMFW_REQUIRE(t, "tower_ready: null handle");
    for (auto& kv : t->slots) {
        if (kv.second.required && !kv.second.set) {
            set_error("tower weights incomplete: '%s' has not been set", kv.first.c_str());
            return VLMFW_E_STATE;
        }
    }
    return VLMFW_OK;
}

extern "C" size_t vlmfw_tower_workspace_bytes(const vlmfw_tower* t, int32_t batch) {
    if (!t || batch <= 0) return 0;
    return t->ws_bytes(batch);
}
extern "C" int32_t vlmfw_tower_tokens(const vlmfw_tower* t) { return t ? t->S : 0; }
extern "C" const float* vlmfw_tower_hidden_ptr(const vlmfw_tower* t, void* workspace) {
    (void)t;
    return static_cast<const float*>(workspace);
}
extern "C" int64_t vlmfw_tower_last_launches(const vlmfw_tower* t) { return t ? t->last_launches : 0; }
extern "C" const void* vlmfw_tower_hidden16_ptr(const vlmfw_tower* t, void* workspace, int32_t batch) {
    if (!t || !workspace || batch <= 0 || !t->resid16) return nullptr;
    return static_cast<const uint8_t*>(workspace) + t->off_xn(batch);
}

extern "C" int vlmfw_tower_set_cls_tail(vlmfw_tower* t, int32_t enable) {
    VLMFW_REQUIRE(t, "tower_set_cls_tail: null tower");
    const bool can = t->cfg.has_cls && (t->cfg.pool == VLMFW_POOL_CLS || t->cfg.pool == VLMFW_POOL_LN_CLS_LN);
    VLMFW_REQUIRE(!enable || can, "tower_set_cls_tail: this tower's pooling reads more than the class token");
    t->cls_tail = enable != 0;
    return VLMFW_OK;
}
extern "C" int32_t vlmfw_tower_cls_tail(const vlmfw_tower* t) { return t && t->cls_tail ? 1 : 0; }

extern "C" int vlmfw_tower_saturation(vlmfw_tower* t, int64_t* host_count, int32_t reset, void* stream) {
    VLMFW_REQUIRE(t && host_count, "tower_saturation: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long v = 0;
    VLMFW_CUDA_OK(cudaMemcpyAsync(&v, t->sat_dev, sizeof(v), cudaMemcpyDeviceToHost, st));
    if (reset) VLMFW_CUDA_OK(cudaMemsetAsync(t->sat_dev, 0, sizeof(v), st));
    VLMFW_CUDA_OK(cudaStreamSynchronize(st));
    *host_count = static_cast<int64_t>(v);
    return VLMFW_OK;
}

extern "C" int vlmfw_tower_weight_scale(const vlmfw_tower* t, const char* name, float* host_scale) {
    VLMFW_REQUIRE(t && name && host_scale, "tower_weight_scale: null argument");
    auto it = t->slots.find(name);
    if (it == t->slots.end() || !it->second.matrix || it->second.scale == nullptr) {
        set_error("tower_weight_scale: '%s' is not a weight matrix of this tower", name);
        return VLMFW_E_NOTFOUND;
    }
    if (it->second.master != nullptr && !t->derived_ok) {
        set_error("tower_weight_scale: '%s' is LayerNorm-folded; call vlmfw_tower_prepare first", name);
        return VLMFW_E_STATE;
    }
    *host_scale = *it->second.scale;
    return VLMFW_OK;
}

static int build_derived(vlmfw_tower* t, cudaStream_t st);
extern "C" int vlmfw_tower_prepare(vlmfw_tower* t, void* stream) {
    int rc = vlmfw_tower_ready(t);
    if (rc != VLMFW_OK) return rc;
    rc = device_check();
    if (rc != VLMFW_OK) return rc;
    return build_derived(t, static_cast<cudaStream_t>(stream));
}

static int build_derived(vlmfw_tower* t, cudaStream_t st) {
    if (t->derived_ok) return VLMFW_OK;
    if (!t->posb) VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&t->posb), static_cast<size_t>(t->P) * t->D * 4));
    int rc = launch_add_rows(t->pos, t->cfg.conv_bias ? t->patch_b : nullptr, t->posb, t->P, t->D, t->cfg.has_cls ? 1 : 0, st);
    if (rc != VLMFW_OK) return rc;
    if (t->cfg.has_cls) {
        if (!t->cls_pos0) VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&t->cls_pos0), static_cast<size_t>(t->D) * 4));
        rc = launch_add_rows(t->pos, t->cls, t->cls_pos0, 1, t->D, 0, st);
        if (rc != VLMFW_OK) return rc;
    }
    // fold ln1 into the QKV projection and ln2 into fc1 (layernorm.cu: ln_fold_weight_kernel). fp16 towers: the folded
    // values are what gets rounded, so their max decides the operand scale — one extra pass only when it is needed.
    auto fold = [&](const float* w32, int n, const float* g, const float* b, const float* bias, void* w16, float* c2,
                    float* scale) -> int {
        *scale = 1.f;
        if (!t->f16) return launch_ln_fold_weight(w32, n, t->D, g, b, bias, w16, t->D, false, c2, st);
        VLMFW_CUDA_OK(cudaMemsetAsync(t->amax_dev, 0, sizeof(unsigned int), st));
        int r = launch_ln_fold_weight(w32, n, t->D, g, b, bias, w16, t->D, true, c2, st, 1.f, t->amax_dev);
        if (r != VLMFW_OK) return r;
        unsigned int bits = 0;
        VLMFW_CUDA_OK(cudaMemcpyAsync(&bits, t->amax_dev, sizeof(bits), cudaMemcpyDeviceToHost, st));
        VLMFW_CUDA_OK(cudaStreamSynchronize(st));
        float amax;
        memcpy(&amax, &bits, sizeof(amax));
        if (!(amax < 3.0e38f)) {
            set_error("tower weights: a LayerNorm-folded matrix holds a non-finite value");
            return VLMFW_E_RANGE;
        }
        *scale = fp16_weight_scale(amax);
        if (*scale != 1.f) r = launch_ln_fold_weight(w32, n, t->D, g, b, bias, w16, t->D, true, c2, st, 1.f / *scale);
        return r;
    };
    for (auto& w : t->layers) {
        if (!w.c2qkv) VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&w.c2qkv), static_cast<size_t>(3 * t->D) * 4));
        if (!w.c2fc1) VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&w.c2fc1), static_cast<size_t>(t->I) * 4));
        rc = fold(w.wqkv32, 3 * t->D, w.ln1w, w.ln1b, w.bqkv, w.wqkv, w.c2qkv, &w.s_qkv);
        if (rc != VLMFW_OK) return rc;
        rc = fold(w.w132, t->I, w.ln2w, w.ln2b, w.b1, w.w1, w.c2fc1, &w.s_fc1);
        if (rc != VLMFW_OK) return rc;
    }
    t->derived_ok = true;
    return VLMFW_OK;
}

extern "C" int vlmfw_tower_forward(vlmfw_tower* t, const void* pixels, int32_t pix_fmt, int32_t batch, void* workspace,
                                   size_t workspace_bytes, float* pooled_out, void* stream) {
    VLMFW_REQUIRE(t && pixels && workspace && pooled_out, "tower_forward: null argument");
    VLMFW_REQUIRE(batch > 0, "tower_forward: batch %d", batch);
    int rc = vlmfw_tower_ready(t);
    if (rc != VLMFW_OK) return rc;
    if (workspace_bytes < t->ws_bytes(batch)) {
        set_error("tower_forward: workspace %zu bytes < required %zu", workspace_bytes, t->ws_bytes(batch));
        return VLMFW_E_STATE;
    }
    VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "tower_forward: workspace must be 1024-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t launches0 = launch_counter();
    rc = build_derived(t, st);
    if (rc != VLMFW_OK) return rc;

    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* h = reinterpret_cast<float*>(ws + t->off_h());
    void* xn = ws + t->off_xn(batch);
    void* qkv = ws + t->off_qkv(batch);
    void* ctx = ws + t->off_ctx(batch);
    void* mid = ws + t->off_mid(batch);
    void* stats = ws + t->off_stats(batch);
    const int np = t->stats_parts();
    void* patch_a = mid;
    const int D = t->D, I = t->I, S = t->S, P = t->P;
    const int M = batch * S;

    // ---- embeddings: conv-as-GEMM (+bias +position embedding), class token, optional pre-LN
    bool head_done = false;  // xn / stats of the embeddings already produced by the patch GEMM's epilogue
    rc = launch_im2col(pixels, pix_fmt, batch, t->cfg.image_size, t->cfg.patch_size, t->cfg.pix_mean, t->cfg.pix_std,
                       patch_a, t->kpad, t->f16, st);
    if (rc != VLMFW_OK) return rc;
    {
        GemmArgs g;
        g.a = patch_a;
        g.lda = t->kpad;
        g.f16 = t->f16;
        g.w = t->patch_w;
        g.acc_scale = t->s_patch;
        g.sat_count = t->f16 ? t->sat_dev : nullptr;
        g.ldw = t->kpad;
        g.m = batch * P;
        g.n = D;
        g.k = t->kpad;
        g.epi = VLMFW_EPI_BIAS_RESID_F32;
        g.out = h;
        g.ldo = D;
        g.resid = t->posb;
        g.ldr = D;
        g.resid_mod = P;
        if (t->cfg.has_cls) {  // rows move over by the class-token slot of every image: generic epilogue
            g.grp = P;
            g.grp_out = S;
            g.grp_off = 1;
        } else if (!t->cfg.has_pre_ln && P % 32 == 0) {
            // no row remap: the TMA residual epilogue (broadcast position embedding) also leaves the 16-bit copy and the
            // row statistics the first QKV GEMM needs — no row_stats_cast launch
            g.h16_out = xn;
            g.ld_h16 = D;
            g.stats_out = stats;
            head_done = true;
        }
        rc = launch_gemm(g, st);
        if (rc != VLMFW_OK) return rc;
    }
    if (t->cfg.has_cls) {
        rc = launch_fill_cls(t->cls_pos0, h, batch, S, D, st);
        if (rc != VLMFW_OK) return rc;
    }
    if (t->cfg.has_pre_ln) {
        rc = launch_layernorm(h, M, D, t->pre_w, t->pre_b, t->cfg.ln_eps, h, VLMFW_DT_F32, st);
        if (rc != VLMFW_OK) return rc;
    }
    // ---- encoder layers (pre-LN transformer blocks). No LayerNorm kernel runs in here: ln1 / ln2 are folded into the
    // QKV / fc1 GEMMs; the out-proj / fc2 epilogues leave the 16-bit copy of h (xn) and its row statistics behind.
    if (!head_done) {
        rc = launch_row_stats_cast(h, M, D, xn, t->f16, stats, np, st, t->f16 ? t->sat_dev : nullptr);
        if (rc != VLMFW_OK) return rc;
    }
    for (int li = 0; li < t->L; ++li) {
        const LayerW& w = t->layers[li];
        // last layer of a class-token-pooled tower: only rows b*S go past the K / V projection (see cls_tail)
        const bool tail = t->cls_tail && li + 1 == t->L;
        const int rows = tail ? batch : M;          // GEMM rows after attention
        const int ld_rows = tail ? S * D : D;       // row stride of those rows in xn / h
        GemmArgs g;
        g.a = xn;
        g.lda = D;
        g.acc_scale = w.s_qkv;
        g.sat_count = t->f16 ? t->sat_dev : nullptr;
        g.f16 = t->f16;
        g.ldw = D;
        g.k = D;
        g.ln_stats = stats;
        g.ln_np = np;
        g.ln_dim = D;
        g.ln_eps = t->cfg.ln_eps;
        g.epi = VLMFW_EPI_BIAS_BF16;
        g.act = VLMFW_ACT_NONE;
        g.ldo = 3 * D;
        if (!tail) {
            g.w = w.wqkv;
            g.m = M;
            g.n = 3 * D;
            g.bias = w.c2qkv;
            g.out = qkv;
            rc = launch_gemm(g, st);
            if (rc != VLMFW_OK) return rc;
            rc = launch_attention(qkv, t->f16, batch, S, t->H, t->dh, ctx, st);
            if (rc != VLMFW_OK) return rc;
        } else {
            // K | V of every row: weight rows [D, 3D), output columns [D, 3D)
            GemmArgs kv = g;
            kv.w = static_cast<const uint16_t*>(w.wqkv) + static_cast<size_t>(D) * D;
            kv.m = M;
            kv.n = 2 * D;
            kv.bias = w.c2qkv + D;
            kv.out = static_cast<uint16_t*>(qkv) + D;
            rc = launch_gemm(kv, st);
            if (rc != VLMFW_OK) return rc;
            // Q of the class-token rows (row b of this GEMM = row b*S of xn / stats / qkv)
            GemmArgs q = g;
            q.lda = S * D;
            q.ln_row_mul = S;
            q.w = w.wqkv;
            q.m = batch;
            q.n = D;
            q.bias = w.c2qkv;
            q.out = qkv;
            q.ldo = S * 3 * D;
            rc = launch_gemm(q, st);
            if (rc != VLMFW_OK) return rc;
            rc = launch_attention_cls(qkv, t->f16, batch, S, t->H, t->dh, ctx, D, st);  // ctx rows 0..B-1, compact
            if (rc != VLMFW_OK) return rc;
        }
        g = GemmArgs();
        g.a = ctx;
        g.lda = D;
        g.w = w.wo;
        g.acc_scale = w.s_o;
        g.sat_count = t->f16 ? t->sat_dev : nullptr;
        g.f16 = t->f16;
        g.ldw = D;
        g.m = rows;
        g.n = D;
        g.k = D;
        g.bias = w.bo;
        if (t->resid16) {  // xn IS the residual stream: updated in place, statistics of the rounded rows
            g.epi = VLMFW_EPI_BIAS_RESID_16;
            g.out = xn;
            g.resid = static_cast<const float*>(xn);
        } else {
            g.epi = VLMFW_EPI_BIAS_RESID_F32;
            g.out = h;
            g.resid = h;
            g.h16_out = xn;
            g.ld_h16 = ld_rows;
        }
        g.ldo = ld_rows;
        g.ldr = ld_rows;
        g.stats_out = stats;  // tail: statistics rows 0..B-1 (compact), read back by fc1 below with the same indexing
        rc = launch_gemm(g, st);
        if (rc != VLMFW_OK) return rc;
        g = GemmArgs();
        g.a = xn;
        g.lda = ld_rows;
        g.w = w.w1;
        g.acc_scale = w.s_fc1;
        g.sat_count = t->f16 ? t->sat_dev : nullptr;
        g.f16 = t->f16;
        g.ldw = D;
        g.m = rows;
        g.n = I;
        g.k = D;
        g.bias = w.c2fc1;
        g.ln_stats = stats;
        g.ln_np = np;
        g.ln_dim = D;
        g.ln_eps = t->cfg.ln_eps;
        g.epi = VLMFW_EPI_BIAS_BF16;
        g.act = t->cfg.act;
        g.out = mid;
        g.ldo = t->Ild();
        rc = launch_gemm(g, st);
        if (rc != VLMFW_OK) return rc;
        g = GemmArgs();
        g.a = mid;
        g.lda = t->Ild();
        g.w = w.w2;
        g.acc_scale = w.s_fc2;
        g.sat_count = t->f16 ? t->sat_dev : nullptr;
        g.f16 = t->f16;
        g.ldw = t->Ild();
        g.m = rows;
        g.n = D;
        g.k = I;
        g.bias = w.b2;
        g.ldo = ld_rows;
        g.ldr = ld_rows;
        if (t->resid16) {
            g.epi = VLMFW_EPI_BIAS_RESID_16;
            g.out = xn;
            g.resid = static_cast<const float*>(xn);
            if (li + 1 < t->L) g.stats_out = stats;
        } else {
            g.epi = VLMFW_EPI_BIAS_RESID_F32;
            g.out = h;
            g.resid = h;
            if (li + 1 < t->L) {  // the last layer feeds the pooling kernel, which reads the fp32 rows
                g.h16_out = xn;
                g.ld_h16 = D;
                g.stats_out = stats;
            }
        }
        rc = launch_gemm(g, st);
        if (rc != VLMFW_OK) return rc;
    }
    // ---- adapter pooling -> [B, D] fp32
    if (t->resid16)
        rc = launch_pool_dt(xn, t->f16 ? VLMFW_DT_F16 : VLMFW_DT_BF16, batch, S, D, t->cfg.pool, t->post_w, t->post_b,
                            t->cfg.ln_eps, pooled_out, st);
    else
        rc = launch_pool(h, batch, S, D, t->cfg.pool, t->post_w, t->post_b, t->cfg.ln_eps, pooled_out, st);
    if (rc != VLMFW_OK) return rc;
    t->last_launches = launch_counter() - launches0;
    return VLMFW_OK;
}

extern "C" int vlmfw_tower_forward_host(vlmfw_tower* t, const uint8_t* host_pixels, int32_t batch, float* host_pooled) {
    VLMFW_REQUIRE(t && host_pixels && host_pooled && batch > 0, "tower_forward_host: bad argument");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    if (!t->own_stream) VLMFW_CUDA_OK(cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking));
    const size_t pix_bytes = static_cast<size_t>(batch) * t->cfg.image_size * t->cfg.image_size * 3;
    const size_t ws_bytes = t->ws_bytes(batch);
    const size_t pooled_bytes = static_cast<size_t>(batch) * t->D * 4;
    if (t->host_pix_bytes < pix_bytes) {
        if (t->host_pix_dev) cudaFree(t->host_pix_dev);
        t->host_pix_dev = nullptr;
        t->host_pix_bytes = 0;
        VLMFW_CUDA_OK(cudaMalloc(&t->host_pix_dev, pix_bytes));
        t->host_pix_bytes = pix_bytes;
    }
    if (t->host_ws_bytes < ws_bytes) {
        if (t->host_ws) cudaFree(t->host_ws);
        t->host_ws = nullptr;
        t->host_ws_bytes = 0;
        VLMFW_CUDA_OK(cudaMalloc(&t->host_ws, ws_bytes + 1024));  // forward() wants 1024-byte alignment (TMA tiles)
        t->host_ws_bytes = ws_bytes;
    }
    if (t->host_pooled_bytes < pooled_bytes) {
        if (t->host_pooled_dev) cudaFree(t->host_pooled_dev);
        t->host_pooled_dev = nullptr;
        t->host_pooled_bytes = 0;
        VLMFW_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&t->host_pooled_dev), pooled_bytes));
        t->host_pooled_bytes = pooled_bytes;
    }
    VLMFW_CUDA_OK(cudaMemcpyAsync(t->host_pix_dev, host_pixels, pix_bytes, cudaMemcpyHostToDevice, t->own_stream));
    void* ws_aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(t->host_ws) + 1023) & ~uintptr_t(1023));
    rc = vlmfw_tower_forward(t, t->host_pix_dev, VLMFW_PIX_U8_NHWC, batch, ws_aligned, t->host_ws_bytes,
                             t->host_pooled_dev, t->own_stream);
    if (rc != VLMFW_OK) return rc;
    VLMFW_CUDA_OK(cudaMemcpyAsync(host_pooled, t->host_pooled_dev, pooled_bytes, cudaMemcpyDeviceToHost, t->own_stream));
    VLMFW_CUDA_OK(cudaStreamSynchronize(t->own_stream));
    return VLMFW_OK;
}

// ------------------------------------------------------------------------- kernel-level entry points (tests)
extern "C" int vlmfw_gemm_tn(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t dtype, int32_t m, int32_t n,
                             int32_t k, const float* bias, int32_t epi, int32_t act, void* out, int32_t ldo,
                             const float* resid, int32_t ldr, int32_t resid_mod, int32_t grp, int32_t grp_out,
                             int32_t grp_off, int32_t block_n, void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "gemm: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    GemmArgs g;
    g.a = a;
    g.lda = lda;
    g.w = w;
    g.ldw = ldw;
    g.f16 = dtype == VLMFW_DT_F16;
    g.m = m;
    g.n = n;
    g.k = k;
    g.bias = bias;
    g.epi = epi;
    g.act = act;
    g.out = out;
    g.ldo = ldo;
    g.resid = resid;
    g.ldr = ldr;
    g.resid_mod = resid_mod;
    g.grp = grp;
    g.grp_out = grp_out;
    g.grp_off = grp_off;
    g.block_n = block_n;
    return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_gemm_tn_ex(const vlmfw_gemm_desc* d, void* stream) {
    VLMFW_REQUIRE(d != nullptr, "gemm_ex: null descriptor");
    VLMFW_REQUIRE(d->dtype == VLMFW_DT_BF16 || d->dtype == VLMFW_DT_F16, "gemm: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    GemmArgs g;
    g.a = d->a;
    g.lda = d->lda;
    g.w = d->w;
    g.ldw = d->ldw;
    g.f16 = d->dtype == VLMFW_DT_F16;
    g.m = d->m;
    g.n = d->n;
    g.k = d->k;
    g.bias = d->bias;
    g.epi = d->epi;
    g.act = d->act;
    g.out = d->out;
    g.ldo = d->ldo;
    g.resid = d->resid;
    g.ldr = d->ldr;
    g.block_n = d->block_n;
    g.ln_stats = d->ln_stats;
    g.ln_np = d->ln_np;
    g.ln_dim = d->ln_dim;
    g.ln_eps = d->ln_eps;
    g.h16_out = d->h16_out;
    g.ld_h16 = d->ld_h16;
    g.stats_out = d->stats_out;
    g.acc_scale = d->acc_scale > 0.f ? d->acc_scale : 1.f;
    g.sat_count = static_cast<unsigned long long*>(d->sat_count);
    return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

extern "C" int32_t vlmfw_gemm_stats_parts(int32_t n, int32_t block_n) { return gemm_stats_parts(n, block_n); }

extern "C" int vlmfw_row_stats_cast(const float* x, int32_t rows, int32_t dim, void* out16, int32_t dtype, void* stats,
                                    int32_t parts, void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "row_stats_cast: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    return launch_row_stats_cast(x, rows, dim, out16, dtype == VLMFW_DT_F16, stats, parts, static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_ln_fold_weight(const float* w, int32_t n, int32_t k, const float* gamma, const float* beta,
                                    const float* bias, void* w16, int32_t ld, int32_t dtype, float* c2, void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "ln_fold_weight: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    return launch_ln_fold_weight(w, n, k, gamma, beta, bias, w16, ld, dtype == VLMFW_DT_F16, c2,
                                 static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_layernorm(const float* x, int32_t rows, int32_t dim, const float* gamma, const float* beta,
                               float eps, void* out, int32_t out_dtype, void* stream) {
    return launch_layernorm(x, rows, dim, gamma, beta, eps, out, out_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_attention(const void* qkv, int32_t dtype, int32_t batch, int32_t tokens, int32_t heads,
                               int32_t head_dim, void* out, void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "attention: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    return launch_attention(qkv, dtype == VLMFW_DT_F16, batch, tokens, heads, head_dim, out,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_attention_cls(const void* qkv, int32_t dtype, int32_t batch, int32_t tokens, int32_t heads,
                                   int32_t head_dim, void* out, int32_t ld_out, void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "attention_cls: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    return launch_attention_cls(qkv, dtype == VLMFW_DT_F16, batch, tokens, heads, head_dim, out, ld_out,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_im2col(const void* pixels, int32_t pix_fmt, int32_t batch, int32_t image_size, int32_t patch_size,
                            const float* mean, const float* std, void* out, int32_t kpad, int32_t out_dtype,
                            void* stream) {
    return launch_im2col(pixels, pix_fmt, batch, image_size, patch_size, mean, std, out, kpad, out_dtype == VLMFW_DT_F16,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_pool(const float* h, int32_t batch, int32_t tokens, int32_t dim, int32_t pool, const float* gamma,
                          const float* beta, float eps, float* pooled_out, void* stream) {
    return launch_pool(h, batch, tokens, dim, pool, gamma, beta, eps, pooled_out, static_cast<cudaStream_t>(stream));
}

extern "C" int vlmfw_pack_16(const float* src, int64_t rows, int64_t cols, void* dst, int64_t dst_ld, int32_t dtype,
                             void* stream) {
    VLMFW_REQUIRE(dtype == VLMFW_DT_BF16 || dtype == VLMFW_DT_F16, "pack_16: dtype must be VLMFW_DT_BF16 or VLMFW_DT_F16");
    return launch_pack_16(src, rows, cols, dst, dst_ld, dtype == VLMFW_DT_F16, static_cast<cudaStream_t>(stream));
}
