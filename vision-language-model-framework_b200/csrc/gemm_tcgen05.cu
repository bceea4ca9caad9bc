// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), fed by TMA.
//
//   out = epilogue( A[M,K] * W[N,K]^T )        A, W 16-bit row-major (both K-major; each bf16 or fp16), fp32 accumulate
//
// This one kernel carries every dense contraction of the tower (reference path: the nn.Linear / Conv2d calls inside
// transformers' SiglipEncoderLayer / CLIPEncoderLayer / Blip2EncoderLayer):
//   patch-embedding conv-as-GEMM (+bias +position embedding, row remap past the class token),
//   fused QKV projection (+bias), attention out-proj (+bias +fp32 residual),
//   MLP fc1 (+bias +GELU variant) and fc2 (+bias +fp32 residual).
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      : TMA producer   — 128B-swizzled [128 x 64] A tiles and [BN x 64] W tiles into a STAGES-deep ring
//   warp 1      : MMA issuer     — one lane issues tcgen05.mma 128 x BN x 16, commits free ring slots / signal epilogue
//   warps 2..9  : epilogue       — tcgen05.ld 32x32 fp32 chunks, transpose through swizzled smem, fused bias /
//                                  activation / residual, fully coalesced global stores
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace vlmfw {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int STAGING_PER_WARP = 32 * 128;  // 32 rows x 32 fp32

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 128) ? 6 : 4;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STAGING_PER_WARP + BAR_BYTES + 1024;
};

struct KParams {
    int m, n, k;
    uint32_t idesc;  // tcgen05 instruction descriptor (the operand format is a run-time choice)
    int out_f16;     // 16-bit epilogue output format
    const float* bias;
    void* out;
    int ldo;
    const float* resid;
    int ldr;
    int resid_mod;
    int grp, grp_out, grp_off;
};

template <int ACT>
__device__ __forceinline__ float activation(float x) {
    if constexpr (ACT == VLMFW_ACT_GELU_TANH) {
        // 0.5 x (1 + tanh(u)) == x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3)   (activations.py gelu_pytorch_tanh)
        float u = 1.5957691216057308f * (x + 0.044715f * x * x * x);
        return __fdividef(x, 1.0f + __expf(-u));
    } else if constexpr (ACT == VLMFW_ACT_QUICK_GELU) {
        return __fdividef(x, 1.0f + __expf(-1.702f * x));  // x * sigmoid(1.702 x)  (activations.py QuickGELU)
    } else if constexpr (ACT == VLMFW_ACT_GELU_ERF) {
        return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    } else {
        return x;
    }
}

template <int BN, int EPI, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const KParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_WARPS * STAGING_PER_WARP);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (p.m + BM - 1) / BM;
    const int n_tiles = (p.n + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.k + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM;
                const int n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
                    tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, n0);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = p.idesc;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + C::A_BYTES);
#pragma unroll
                    for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                        // advancing K inside the 128B swizzle row: +32 bytes -> +2 in the (addr >> 4) field
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(ks * 2), bdesc + static_cast<uint64_t>(ks * 2),
                                     idesc, (kb | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // ring slot reusable once these MMAs have read it
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int ew = warp - 2;
        const int quarter = warp & 3;        // TMEM lanes [32*quarter, +32) are the only ones this warp may read
        const int half = ew >> 2;            // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        constexpr int CHUNKS = COLS_PER_WARP / 32;
        uint8_t* stg = staging + ew * STAGING_PER_WARP;
        const uint32_t stg_u32 = smem_u32(stg);
        const int rsub = lane >> 3;  // read-back: 4 rows per pass
        const int cj = lane & 7;     // read-back: 16-byte column group (4 fp32)
        int acc = 0;
        uint32_t acc_phase = 0;

        // Residual epilogue: the fp32 residual tile is the only global READ of the epilogue and it does not depend on
        // the accumulator, so it is software-pipelined one 32-column chunk ahead (across tile boundaries too): the
        // loads of chunk i+1 are in flight while chunk i waits for TMEM / transposes / stores.
        float4 rnext[8];
        auto resid_issue = [&](int tile, int c, float4 (&r)[8]) {
            if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
                const int m0 = (tile / n_tiles) * BM + quarter * 32;
                const int col = (tile % n_tiles) * BN + half * COLS_PER_WARP + c * 32 + cj * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = m0 + it * 4 + rsub;
                    r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < p.m && col < p.n) {
                        const int rrow = (p.resid_mod > 0) ? (row % p.resid_mod) : row;
                        r[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(rrow) * p.ldr + col);
                    }
                }
            }
        };
        if (blockIdx.x < num_tiles) resid_issue(blockIdx.x, 0, rnext);

        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BM + quarter * 32;
            const int n0 = (tile % n_tiles) * BN + half * COLS_PER_WARP;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                const int nc = n0 + c * 32;
                float4 rcur[8];
                if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) rcur[it] = rnext[it];
                    // next chunk of this tile, or chunk 0 of this CTA's next tile
                    if (c + 1 < CHUNKS) resid_issue(tile, c + 1, rnext);
                    else if (tile + static_cast<int>(gridDim.x) < num_tiles) resid_issue(tile + gridDim.x, 0, rnext);
                }
                if (nc >= p.n) continue;  // warp-uniform: whole chunk is past the N edge (prefetch above still ran)
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                       static_cast<uint32_t>(acc * BN + half * COLS_PER_WARP + c * 32);
                tmem_ld_32x32(taddr, v);
                tmem_ld_wait();
                // transpose through smem: thread `lane` owns tile row `lane`; 16B groups XOR-swizzled by row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t dst = stg_u32 + lane * 128 + ((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                                 "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                 : "memory");
                }
                __syncwarp();
                const int col = nc + cj * 4;
                const bool col_ok = col < p.n;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr && col_ok) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rsub;
                    const int row = m0 + rr;
                    float4 a;
                    const uint32_t src = stg_u32 + rr * 128 + ((cj ^ (rr & 7)) << 4);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                                 : "r"(src));
                    if (row < p.m && col_ok) {
                        a.x += bv.x;
                        a.y += bv.y;
                        a.z += bv.z;
                        a.w += bv.w;
                        int orow = row;
                        if (p.grp > 0) orow = (row / p.grp) * p.grp_out + (row % p.grp) + p.grp_off;
                        if constexpr (EPI == VLMFW_EPI_BIAS_RESID_F32) {
                            a.x += rcur[it].x;
                            a.y += rcur[it].y;
                            a.z += rcur[it].z;
                            a.w += rcur[it].w;
                            *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(orow) * p.ldo +
                                                       col) = a;
                        } else {
                            uint2 o;
                            const float y0 = activation<ACT>(a.x), y1 = activation<ACT>(a.y);
                            const float y2 = activation<ACT>(a.z), y3 = activation<ACT>(a.w);
                            o.x = p.out_f16 ? pack_f16x2(y0, y1) : pack_bf16x2(y0, y1);
                            o.y = p.out_f16 ? pack_f16x2(y2, y3) : pack_bf16x2(y2, y3);
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) +
                                                      static_cast<size_t>(orow) * p.ldo + col) = o;
                        }
                    }
                }
                __syncwarp();
            }
            // accumulator drained -> hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

struct MapKey {
    const void* ptr;
    int rows, cols, ld, box_rows, f16;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && f16 == o.f16;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
        h ^= (static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
        h ^= (static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull) + (h << 6) + (h >> 2);
        h ^= static_cast<size_t>(k.box_rows * 2 + k.f16) + (h << 6) + (h >> 2);
        return h;
    }
};

// [rows, cols] bf16 row-major (row stride ld elements) -> tiled map with box {64 cols, box_rows}, 128B swizzle,
// out-of-bounds elements read as zero (this is what makes ragged M / N / K edges free).
int get_tensor_map(const void* ptr, int rows, int cols, int ld, int box_rows, bool f16, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows, f16 ? 1 : 0};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return VLMFW_OK;
        }
    }
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return VLMFW_E_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estride[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = fn(&tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d",
                  static_cast<int>(r), ptr, rows, cols, ld, box_rows);
        return VLMFW_E_CUDA;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 4096) cache.clear();
        cache.emplace(key, tm);
    }
    *out = tm;
    return VLMFW_OK;
}

template <int BN, int EPI, int ACT>
int launch_one(const GemmArgs& g, cudaStream_t stream) {
    using C = Cfg<BN>;
    CUtensorMap ta, tw;
    int rc = get_tensor_map(g.a, g.m, g.k, g.lda, BM, g.f16, &ta);
    if (rc != VLMFW_OK) return rc;
    rc = get_tensor_map(g.w, g.n, g.k, g.ldw, BN, g.f16, &tw);
    if (rc != VLMFW_OK) return rc;
    auto kern = gemm_bf16_tn_kernel<BN, EPI, ACT>;
    static bool attr_set = false;
    if (!attr_set) {
        VLMFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    KParams p;
    p.m = g.m;
    p.n = g.n;
    p.k = g.k;
    p.idesc = umma_idesc_16(BM, BN, g.f16, g.f16);
    p.out_f16 = g.f16 ? 1 : 0;
    p.bias = g.bias;
    p.out = g.out;
    p.ldo = g.ldo;
    p.resid = g.resid;
    p.ldr = g.ldr;
    p.resid_mod = g.resid_mod;
    p.grp = g.grp;
    p.grp_out = g.grp_out;
    p.grp_off = g.grp_off;
    const int m_tiles = (g.m + BM - 1) / BM;
    const int n_tiles = (g.n + BN - 1) / BN;
    const int tiles = m_tiles * n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(ta, tw, p);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}

template <int BN>
int dispatch_epi(const GemmArgs& g, cudaStream_t stream) {
    if (g.epi == VLMFW_EPI_BIAS_RESID_F32) return launch_one<BN, VLMFW_EPI_BIAS_RESID_F32, VLMFW_ACT_NONE>(g, stream);
    switch (g.act) {
        case VLMFW_ACT_GELU_TANH:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_GELU_TANH>(g, stream);
        case VLMFW_ACT_QUICK_GELU:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_QUICK_GELU>(g, stream);
        case VLMFW_ACT_GELU_ERF:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_GELU_ERF>(g, stream);
        default:
            return launch_one<BN, VLMFW_EPI_BIAS_BF16, VLMFW_ACT_NONE>(g, stream);
    }
}

// Pick the N tile that wastes the fewest MMA columns; ties go to the wider tile (less A re-streaming).
int pick_block_n(int n) {
    const int cands[3] = {256, 192, 128};
    int best = 256;
    long best_cost = -1;
    for (int bn : cands) {
        long tiles = (n + bn - 1) / bn;
        long cost = tiles * bn;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace

int encode_tensor_map(CUtensorMap* out, bool f16, int rank, const void* ptr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return VLMFW_E_CUDA;
    }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estride[i] = 1;
        if (i > 0) gstride[i - 1] = strides_bytes[i - 1];
    }
    const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                    const_cast<void*>(ptr), gdim, gstride, bx, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(rank %d) failed (CUresult %d)", rank, static_cast<int>(r));
        return VLMFW_E_CUDA;
    }
    return VLMFW_OK;
}

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
    VLMFW_REQUIRE(g.a && g.w && g.out, "gemm: null operand");
    VLMFW_REQUIRE(g.m > 0 && g.n > 0 && g.k > 0, "gemm: bad shape m=%d n=%d k=%d", g.m, g.n, g.k);
    VLMFW_REQUIRE(g.k % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm: k/lda/ldw must be multiples of 8 (16-byte TMA rows)");
    VLMFW_REQUIRE(g.n % 8 == 0 && g.ldo % 8 == 0, "gemm: n/ldo must be multiples of 8");
    VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
                  "gemm: operands must be 16-byte aligned");
    VLMFW_REQUIRE(g.lda >= g.k && g.ldw >= g.k && g.ldo >= g.n, "gemm: leading dimension smaller than row");
    if (g.epi == VLMFW_EPI_BIAS_RESID_F32) {
        VLMFW_REQUIRE(g.resid != nullptr && g.ldr % 4 == 0 && g.ldr >= g.n, "gemm: residual epilogue needs resid/ldr");
        VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.resid) & 15) == 0, "gemm: resid must be 16-byte aligned");
    } else {
        VLMFW_REQUIRE(g.epi == VLMFW_EPI_BIAS_BF16, "gemm: unknown epilogue %d", g.epi);
    }
    if (g.bias) VLMFW_REQUIRE((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int bn = g.block_n > 0 ? g.block_n : pick_block_n(g.n);
    switch (bn) {
        case 256:
            return dispatch_epi<256>(g, stream);
        case 192:
            return dispatch_epi<192>(g, stream);
        case 128:
            return dispatch_epi<128>(g, stream);
        default:
            set_error("gemm: unsupported block_n %d", bn);
            return VLMFW_E_INVALID;
    }
}

}  // namespace vlmfw
