// Image ingest on the GPU: the resize / crop of the Hugging Face image processor the reference calls in front of every
// tower (reference models/paligemma.py:108, models/llava.py:109, models/blip2.py:91 -> hf image_processing_backends.py
// TorchvisionBackend.resize / center_crop). On a uint8 tensor torchvision's bicubic antialiased resize is ATen's
// separable FIXED-POINT kernel (PyTorch aten/src/ATen/native/cpu/UpSampleKernel.cpp: int16 weights with a shared
// precision, horizontal pass, then vertical pass, each rounded back to uint8) — integer work, so the result here is
// bit-exact, not "close".
//
//   host   : AxisTable  — per (input size, output size) the tap window [xmin, xmin + xsize) and int16 weights of every
//                         output position, in double arithmetic with ATen's operation order; cached per handle.
//   device : resize_rows_kernel — horizontal pass, ragged batch: one CTA = 4 input rows of one image staged in shared
//                         memory with 16-byte loads, one thread = one output pixel (3 channels share the weights),
//                         output row staged in shared memory and written with 16-byte stores.
//            resize_cols_kernel — vertical pass: one thread = 4 consecutive output bytes, taps read as coalesced words.
//   Only what the crop keeps is computed (CLIP: the 336 centre columns / rows of the shortest-edge resize).
//   Both passes are HBM/LSU-bound byte kernels; the raw image is read once, the intermediate [rows, S, 3] is L2-sized.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace vlmfw {
namespace {

// ------------------------------------------------------------------------------------------------ host: weight tables
struct AxisTable {
    int n = 0;     // output positions kept (the crop window)
    int taps = 0;  // row length of w
    int prec = 0;
    std::vector<int32_t> xmin, xsize;
    std::vector<int16_t> w;  // [n][taps]
};

inline double cubic_aa(double x) {
    const double a = -0.5;
    x = std::fabs(x);
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a;
    return 0.0;
}

// Output positions [off, off + n) of the `in` -> `out_full` antialiased bicubic resize. in == out_full is the identity
// (ATen skips that pass): one tap of weight 2^prec reproduces the pixel exactly through (2^(prec-1) + v 2^prec) >> prec.
AxisTable build_axis_table(int in, int out_full, int off, int n) {
    AxisTable t;
    t.n = n;
    t.xmin.resize(n);
    t.xsize.resize(n);
    if (in == out_full) {
        t.taps = 1;
        t.prec = 1;
        t.w.assign(n, 2);
        for (int i = 0; i < n; ++i) {
            t.xmin[i] = off + i;
            t.xsize[i] = 1;
        }
        return t;
    }
    const double scale = static_cast<double>(in) / static_cast<double>(out_full);
    const double support = scale >= 1.0 ? 2.0 * scale : 2.0;
    const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
    const int max_taps = static_cast<int>(std::ceil(support)) * 2 + 1;
    t.taps = max_taps;
    std::vector<double> wd(static_cast<size_t>(n) * max_taps, 0.0);
    // ATen takes the maximum weight over the WHOLE axis (all out_full positions), not just the crop window
    double wmax = 0.0;
    std::vector<double> row(max_taps);
    for (int i = 0; i < out_full; ++i) {
        const double center = scale * (i + 0.5);
        const int64_t lo = std::max<int64_t>(static_cast<int64_t>(center - support + 0.5), 0);
        int64_t cnt = std::min<int64_t>(static_cast<int64_t>(center + support + 0.5), in) - lo;
        cnt = std::min<int64_t>(std::max<int64_t>(cnt, 0), max_taps);
        double total = 0.0;
        for (int64_t j = 0; j < cnt; ++j) {
            const double wv = cubic_aa((j + lo - center + 0.5) * invscale);
            row[j] = wv;
            total += wv;
        }
        for (int64_t j = 0; j < cnt; ++j) {
            if (total != 0.0) row[j] /= total;
            wmax = std::max(wmax, row[j]);
        }
        if (i >= off && i < off + n) {
            t.xmin[i - off] = static_cast<int32_t>(lo);
            t.xsize[i - off] = static_cast<int32_t>(cnt);
            for (int64_t j = 0; j < cnt; ++j) wd[static_cast<size_t>(i - off) * max_taps + j] = row[j];
        }
    }
    int prec = 0;
    for (; prec < 22; ++prec) {
        const int next = static_cast<int>(0.5 + wmax * static_cast<double>(1 << (prec + 1)));
        if (next >= (1 << 15)) break;
    }
    t.prec = prec;
    t.w.resize(wd.size());
    const double mul = static_cast<double>(1 << prec);
    for (size_t i = 0; i < wd.size(); ++i) {
        const double v = wd[i];
        t.w[i] = static_cast<int16_t>(v < 0.0 ? -0.5 + v * mul : 0.5 + v * mul);
    }
    return t;
}

void shortest_edge(int h, int w, int s, int* nh, int* nw) {
    // hf image_transforms.get_resize_output_image_size(default_to_square=False): int(size * long / short)
    if (h <= w) {
        *nh = s;
        *nw = static_cast<int>(static_cast<double>(s) * w / h);
    } else {
        *nh = static_cast<int>(static_cast<double>(s) * h / w);
        *nw = s;
    }
}

// ------------------------------------------------------------------------------------------------ device side
struct ImageDesc {
    int64_t src_off;  // byte offset of the image inside the packed input
    int64_t mid_off;  // byte offset of its intermediate rows inside the workspace
    int32_t h, w;
    int32_t row0, rows;  // input rows the vertical pass will read: [row0, row0 + rows)
    int32_t tab_x, tab_y;  // byte offsets of the two tables inside the blob
};

// table blob layout: int32 {taps, prec, n, 0} | int32 xmin[n] | int32 xsize[n] | int16 w[n * taps] (padded to 16 bytes)
struct TableView {
    int taps, prec;
    const int32_t* xmin;
    const int32_t* xsize;
    const int16_t* w;
};
__device__ __forceinline__ TableView view_table(const uint8_t* blob, int off) {
    const int32_t* hdr = reinterpret_cast<const int32_t*>(blob + off);
    TableView v;
    v.taps = hdr[0];
    v.prec = hdr[1];
    const int n = hdr[2];
    v.xmin = hdr + 4;
    v.xsize = hdr + 4 + n;
    v.w = reinterpret_cast<const int16_t*>(hdr + 4 + 2 * n);
    return v;
}

__device__ __forceinline__ int clamp_u8(int v) { return min(max(v, 0), 255); }

constexpr int kRowsPerCta = 4;

// Horizontal pass. grid = (ceil(max rows / 4), batch); dynamic smem = 4 * (in_row_bytes_max + 32) + 4 * align16(S * 3).
__global__ void __launch_bounds__(256) resize_rows_kernel(const uint8_t* __restrict__ packed, int64_t packed_bytes,
                                                          const uint8_t* __restrict__ blob, uint8_t* __restrict__ work,
                                                          int out_n, int in_stride_smem) {
    extern __shared__ __align__(16) uint8_t smem[];
    const ImageDesc d = reinterpret_cast<const ImageDesc*>(blob)[blockIdx.y];
    const int r_first = blockIdx.x * kRowsPerCta;
    if (r_first >= d.rows) return;
    const int nrows = min(kRowsPerCta, d.rows - r_first);
    const TableView tx = view_table(blob, d.tab_x);
    const int row_bytes = d.w * 3;
    uint8_t* s_out = smem + kRowsPerCta * in_stride_smem;
    const int out_bytes = (out_n * 3 + 15) & ~15;  // intermediate row stride: 16-byte multiple (pad bytes are never read back)

    // stage the input rows: 16-byte chunks from the aligned-down address, byte loads only where a chunk would leave the
    // packed buffer (first / last bytes of the batch)
    for (int r = 0; r < nrows; ++r) {
        const int64_t g0 = d.src_off + static_cast<int64_t>(d.row0 + r_first + r) * row_bytes;
        const int shift = static_cast<int>(g0 & 15);
        const int64_t a0 = g0 - shift;
        const int chunks = (shift + row_bytes + 15) >> 4;
        uint8_t* dst = smem + r * in_stride_smem;
        for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
            const int64_t ga = a0 + static_cast<int64_t>(c) * 16;
            if (ga + 16 <= packed_bytes) {
                *reinterpret_cast<uint4*>(dst + c * 16) = __ldg(reinterpret_cast<const uint4*>(packed + ga));
            } else {
                for (int b = 0; b < 16; ++b) dst[c * 16 + b] = ga + b < packed_bytes ? packed[ga + b] : 0;
            }
        }
    }
    __syncthreads();
    const int init = 1 << (tx.prec - 1);
    for (int idx = threadIdx.x; idx < nrows * out_n; idx += blockDim.x) {
        const int r = idx / out_n, x = idx - r * out_n;
        const int64_t g0 = d.src_off + static_cast<int64_t>(d.row0 + r_first + r) * row_bytes;
        const uint8_t* src = smem + r * in_stride_smem + static_cast<int>(g0 & 15) + tx.xmin[x] * 3;
        const int16_t* wr = tx.w + static_cast<size_t>(x) * tx.taps;
        const int cnt = tx.xsize[x];
        int a0 = init, a1 = init, a2 = init;
        for (int j = 0; j < cnt; ++j) {
            const int wv = wr[j];
            a0 += wv * src[3 * j];
            a1 += wv * src[3 * j + 1];
            a2 += wv * src[3 * j + 2];
        }
        uint8_t* o = s_out + r * out_bytes + x * 3;
        o[0] = static_cast<uint8_t>(clamp_u8(a0 >> tx.prec));
        o[1] = static_cast<uint8_t>(clamp_u8(a1 >> tx.prec));
        o[2] = static_cast<uint8_t>(clamp_u8(a2 >> tx.prec));
    }
    __syncthreads();
    // intermediate rows sit at 16-byte aligned offsets with a 16-byte multiple stride
    uint8_t* mid = work + d.mid_off + static_cast<int64_t>(r_first) * out_bytes;
    const int vec = nrows * out_bytes / 16;
    for (int c = threadIdx.x; c < vec; c += blockDim.x)
        reinterpret_cast<uint4*>(mid)[c] = reinterpret_cast<const uint4*>(s_out)[c];
}

// Vertical pass. grid = (ceil(align16(S*3)/4 / 256), S, batch): one thread = 4 consecutive bytes of one output row.
__global__ void __launch_bounds__(256) resize_cols_kernel(const uint8_t* __restrict__ blob,
                                                          const uint8_t* __restrict__ work, uint8_t* __restrict__ out,
                                                          int out_n) {
    const ImageDesc d = reinterpret_cast<const ImageDesc*>(blob)[blockIdx.z];
    const TableView ty = view_table(blob, d.tab_y);
    const int y = blockIdx.y;
    const int row_bytes = out_n * 3;
    const int words = ((row_bytes + 15) & ~15) / 4;  // intermediate row stride in words
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi * 4 >= row_bytes) return;
    const int cnt = ty.xsize[y];
    const int16_t* wr = ty.w + static_cast<size_t>(y) * ty.taps;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(work + d.mid_off) +
                          static_cast<int64_t>(ty.xmin[y] - d.row0) * words + wi;
    const int init = 1 << (ty.prec - 1);
    int a0 = init, a1 = init, a2 = init, a3 = init;
    for (int j = 0; j < cnt; ++j) {
        const int wv = wr[j];
        const uint32_t p = __ldg(src + static_cast<int64_t>(j) * words);
        a0 += wv * static_cast<int>(p & 0xffu);
        a1 += wv * static_cast<int>((p >> 8) & 0xffu);
        a2 += wv * static_cast<int>((p >> 16) & 0xffu);
        a3 += wv * static_cast<int>(p >> 24);
    }
    const uint32_t o = static_cast<uint32_t>(clamp_u8(a0 >> ty.prec)) | (static_cast<uint32_t>(clamp_u8(a1 >> ty.prec)) << 8) |
                       (static_cast<uint32_t>(clamp_u8(a2 >> ty.prec)) << 16) | (static_cast<uint32_t>(clamp_u8(a3 >> ty.prec)) << 24);
    uint8_t* dst = out + (static_cast<int64_t>(blockIdx.z) * out_n + y) * row_bytes + wi * 4;
    if ((row_bytes & 3) == 0) {
        *reinterpret_cast<uint32_t*>(dst) = o;
    } else {  // sizes whose rows are not whole words (test-sized towers): byte stores
        for (int b = 0; b < 4 && wi * 4 + b < row_bytes; ++b) dst[b] = static_cast<uint8_t>(o >> (8 * b));
    }
}

inline size_t align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

}  // namespace
}  // namespace vlmfw

using namespace vlmfw;

struct vlmfw_preproc {
    int mode = 0, size = 0;
    std::map<std::tuple<int, int, int>, AxisTable> cache;  // (in, out_full, off) -> table
    // pinned staging ring for the descriptor + table blob of each call
    static constexpr int kSlots = 4;
    uint8_t* slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t slot_bytes[kSlots] = {0, 0, 0, 0};
    cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    int next = 0;
    int64_t last_launches = 0;

    const AxisTable& table(int in, int out_full, int off) {
        auto key = std::make_tuple(in, out_full, off);
        auto it = cache.find(key);
        if (it == cache.end()) {
            if (cache.size() > 4096) cache.clear();
            it = cache.emplace(key, build_axis_table(in, out_full, off, size)).first;
        }
        return it->second;
    }
};

namespace {

struct Plan {
    std::vector<ImageDesc> descs;
    std::vector<uint8_t> blob;  // descs + tables
    size_t mid_bytes = 0;
    int max_rows = 0, max_w = 0;
};

size_t table_bytes(const AxisTable& t) { return align16(16 + 8 * static_cast<size_t>(t.n) + 2 * t.w.size()); }

int make_plan(vlmfw_preproc* p, const int64_t* offsets, const int32_t* hs, const int32_t* ws, int batch, Plan* plan) {
    const int S = p->size;
    plan->descs.resize(batch);
    std::map<const AxisTable*, int> placed;
    std::vector<const AxisTable*> order;
    size_t tab_off = align16(sizeof(ImageDesc) * static_cast<size_t>(batch));
    size_t mid = 0;
    for (int i = 0; i < batch; ++i) {
        const int h = hs[i], w = ws[i];
        VLMFW_REQUIRE(h > 0 && w > 0 && h <= 32768 && w <= 16384, "preproc: image %d has unsupported size %d x %d", i, h, w);
        int nh = S, nw = S, top = 0, left = 0;
        if (p->mode == VLMFW_RESIZE_SHORTEST_CROP) {
            shortest_edge(h, w, S, &nh, &nw);
            // hf image_processing_backends.py:357-358: int((size - crop) / 2), truncating
            top = static_cast<int>((nh - S) / 2.0);
            left = static_cast<int>((nw - S) / 2.0);
        }
        const AxisTable& tx = p->table(w, nw, left);
        const AxisTable& ty = p->table(h, nh, top);
        ImageDesc& d = plan->descs[i];
        d.src_off = offsets ? offsets[i] : 0;
        d.h = h;
        d.w = w;
        d.row0 = ty.xmin[0];
        d.rows = ty.xmin[S - 1] + ty.xsize[S - 1] - d.row0;
        d.mid_off = static_cast<int64_t>(mid);
        mid += static_cast<size_t>(d.rows) * align16(static_cast<size_t>(S) * 3);
        for (const AxisTable* t : {&tx, &ty}) {
            if (!placed.count(t)) {
                placed[t] = static_cast<int>(tab_off);
                order.push_back(t);
                tab_off += table_bytes(*t);
            }
        }
        d.tab_x = placed[&tx];
        d.tab_y = placed[&ty];
        plan->max_rows = std::max(plan->max_rows, d.rows);
        plan->max_w = std::max(plan->max_w, w);
    }
    plan->mid_bytes = mid;
    plan->blob.assign(tab_off, 0);
    std::memcpy(plan->blob.data(), plan->descs.data(), sizeof(ImageDesc) * static_cast<size_t>(batch));
    for (const AxisTable* t : order) {
        uint8_t* dst = plan->blob.data() + placed[t];
        const int32_t hdr[4] = {t->taps, t->prec, t->n, 0};
        std::memcpy(dst, hdr, 16);
        std::memcpy(dst + 16, t->xmin.data(), 4 * static_cast<size_t>(t->n));
        std::memcpy(dst + 16 + 4 * static_cast<size_t>(t->n), t->xsize.data(), 4 * static_cast<size_t>(t->n));
        std::memcpy(dst + 16 + 8 * static_cast<size_t>(t->n), t->w.data(), 2 * t->w.size());
    }
    return VLMFW_OK;
}

}  // namespace

extern "C" int vlmfw_preproc_create(int32_t mode, int32_t out_size, vlmfw_preproc** out) {
    VLMFW_REQUIRE(out, "preproc_create: null argument");
    VLMFW_REQUIRE(mode == VLMFW_RESIZE_SQUARE || mode == VLMFW_RESIZE_SHORTEST_CROP, "preproc_create: unknown mode %d", mode);
    VLMFW_REQUIRE(out_size > 0 && out_size <= 4096, "preproc_create: out_size %d out of range (1..4096)", out_size);
    auto* p = new vlmfw_preproc();
    p->mode = mode;
    p->size = out_size;
    *out = p;
    return VLMFW_OK;
}

extern "C" void vlmfw_preproc_destroy(vlmfw_preproc* p) {
    if (!p) return;
    for (int i = 0; i < vlmfw_preproc::kSlots; ++i) {
        if (p->ev[i]) {
            cudaEventSynchronize(p->ev[i]);
            cudaEventDestroy(p->ev[i]);
        }
        if (p->slot[i]) cudaFreeHost(p->slot[i]);
    }
    delete p;
}

extern "C" int vlmfw_resize_weights(int32_t in_size, int32_t out_full, int32_t offset, int32_t count, int32_t* xmin,
                                    int32_t* xsize, int16_t* weights, int32_t weights_ld, int32_t* taps,
                                    int32_t* precision) {
    VLMFW_REQUIRE(in_size > 0 && out_full > 0 && offset >= 0 && count > 0 && offset + count <= out_full,
                  "resize_weights: bad sizes in=%d out=%d offset=%d count=%d", in_size, out_full, offset, count);
    const AxisTable t = build_axis_table(in_size, out_full, offset, count);
    if (taps) *taps = t.taps;
    if (precision) *precision = t.prec;
    if (xmin) std::memcpy(xmin, t.xmin.data(), 4 * static_cast<size_t>(count));
    if (xsize) std::memcpy(xsize, t.xsize.data(), 4 * static_cast<size_t>(count));
    if (weights) {
        VLMFW_REQUIRE(weights_ld >= t.taps, "resize_weights: weights_ld %d < taps %d", weights_ld, t.taps);
        for (int i = 0; i < count; ++i) {
            std::memset(weights + static_cast<size_t>(i) * weights_ld, 0, 2 * static_cast<size_t>(weights_ld));
            std::memcpy(weights + static_cast<size_t>(i) * weights_ld, t.w.data() + static_cast<size_t>(i) * t.taps,
                        2 * static_cast<size_t>(t.taps));
        }
    }
    return VLMFW_OK;
}

extern "C" size_t vlmfw_preproc_workspace_bytes(vlmfw_preproc* p, const int32_t* heights_host, const int32_t* widths_host,
                                                int32_t batch) {
    if (!p || !heights_host || !widths_host || batch <= 0) return 0;
    Plan plan;
    if (make_plan(p, nullptr, heights_host, widths_host, batch, &plan) != VLMFW_OK) return 0;
    return align16(plan.blob.size()) + plan.mid_bytes + 256;
}

extern "C" int64_t vlmfw_preproc_last_launches(const vlmfw_preproc* p) { return p ? p->last_launches : 0; }

extern "C" int vlmfw_preproc_run(vlmfw_preproc* p, const uint8_t* packed, int64_t packed_bytes, const int64_t* offsets_host,
                                 const int32_t* heights_host, const int32_t* widths_host, int32_t batch, uint8_t* out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    VLMFW_REQUIRE(p && heights_host && widths_host && offsets_host, "preproc_run: null argument");
    if (batch == 0) return VLMFW_OK;
    VLMFW_REQUIRE(batch > 0 && batch <= 65535, "preproc_run: batch %d out of range", batch);
    VLMFW_REQUIRE(packed && out && workspace, "preproc_run: null device pointer");
    VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                  "preproc_run: packed / workspace must be 16-byte aligned, out 4-byte aligned");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    Plan plan;
    rc = make_plan(p, offsets_host, heights_host, widths_host, batch, &plan);
    if (rc != VLMFW_OK) return rc;
    for (int i = 0; i < batch; ++i) {
        const int64_t end = offsets_host[i] + static_cast<int64_t>(heights_host[i]) * widths_host[i] * 3;
        VLMFW_REQUIRE(offsets_host[i] >= 0 && end <= packed_bytes, "preproc_run: image %d ([%lld, %lld)) leaves the packed buffer (%lld bytes)",
                      i, static_cast<long long>(offsets_host[i]), static_cast<long long>(end), static_cast<long long>(packed_bytes));
    }
    const size_t blob_bytes = align16(plan.blob.size());
    if (blob_bytes + plan.mid_bytes > workspace_bytes) {
        set_error("preproc_run: workspace too small (%zu bytes, need %zu)", workspace_bytes, blob_bytes + plan.mid_bytes);
        return VLMFW_E_STATE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // blob -> pinned slot -> device (async; the slot is reused only after its copy has completed)
    const int s = p->next;
    p->next = (p->next + 1) % vlmfw_preproc::kSlots;
    if (!p->ev[s]) VLMFW_CUDA_OK(cudaEventCreateWithFlags(&p->ev[s], cudaEventDisableTiming));
    else VLMFW_CUDA_OK(cudaEventSynchronize(p->ev[s]));
    if (p->slot_bytes[s] < blob_bytes) {
        if (p->slot[s]) VLMFW_CUDA_OK(cudaFreeHost(p->slot[s]));
        p->slot[s] = nullptr;
        p->slot_bytes[s] = 0;
        VLMFW_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&p->slot[s]), blob_bytes * 2));
        p->slot_bytes[s] = blob_bytes * 2;
    }
    std::memcpy(p->slot[s], plan.blob.data(), plan.blob.size());
    uint8_t* blob_dev = static_cast<uint8_t*>(workspace);
    uint8_t* work = blob_dev + blob_bytes;
    VLMFW_CUDA_OK(cudaMemcpyAsync(blob_dev, p->slot[s], plan.blob.size(), cudaMemcpyHostToDevice, st));
    VLMFW_CUDA_OK(cudaEventRecord(p->ev[s], st));

    const int S = p->size;
    const int in_stride = static_cast<int>(align16(static_cast<size_t>(plan.max_w) * 3 + 32));
    const size_t smem = static_cast<size_t>(kRowsPerCta) * in_stride + static_cast<size_t>(kRowsPerCta) * align16(static_cast<size_t>(S) * 3);
    VLMFW_REQUIRE(smem <= 220 * 1024, "preproc_run: image width %d needs %zu bytes of shared memory", plan.max_w, smem);
    if (smem > 48 * 1024) {
        static std::atomic<uint64_t> rows_opt_in{0};  // one bit per device
        VLMFW_CUDA_OK(ensure_dynamic_smem(resize_rows_kernel, 220 * 1024, rows_opt_in));
    }
    const int64_t before = launch_counter();
    dim3 g1((plan.max_rows + kRowsPerCta - 1) / kRowsPerCta, batch);
    resize_rows_kernel<<<g1, 256, smem, st>>>(packed, packed_bytes, blob_dev, work, S, in_stride);
    VLMFW_LAUNCH_CHECK();
    const int words = static_cast<int>(align16(static_cast<size_t>(S) * 3) / 4);
    dim3 g2((words + 255) / 256, S, batch);
    resize_cols_kernel<<<g2, 256, 0, st>>>(blob_dev, work, out, S);
    VLMFW_LAUNCH_CHECK();
    p->last_launches = launch_counter() - before;
    return VLMFW_OK;
}
