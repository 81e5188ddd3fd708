// K3: tensor-core GEMM core for sm_100a -- tcgen05.mma (kind::f16, fp32 accumulators in TMEM) fed by TMA.
//
//   D[b,h][m][n] = epilogue( alpha * sum_k A[b,h][m][k] * B[b,h][n][k] )
//
// Replaces the cuBLAS calls under whisper.model.Linear / Conv1d / the q@k^T and w@v contractions of
// MultiHeadAttention.qkv_attention (reached from stable_whisper/timing.py:60-61, decode.py:29,40).
//
// Operands are K-major "split fp16" views: hi plane (+ optional lo plane, hi+lo == fp32 value to 2^-22).  With both
// lo planes present the kernel issues three MMA passes per k-block (hi*hi + hi*lo + lo*hi) into the same TMEM
// accumulator, which reproduces an fp32 GEMM to ~1e-6 relative on the fp16 tensor pipe (the reference's CPU/align
// path is fp32; SURVEY.md section 7 "precision vs the 1e-3 logits gate").  One pass = STB_PREC_FP16.
//
// Structure (PERSISTENT: one CTA per SM loops over 128 x BN output tiles; 192 threads; the TMEM accumulator is
// double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1):
//   warp 0      : TMA producer  (cp.async.bulk.tensor.4d, SWIZZLE_128B tiles, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2..5  : epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / GELU / residual / scale -> global
//                 (fp32 row-major, split-fp16 row-major, or either transposed).
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace stb {

struct GemmArgs {
    int M, N, K, H, Z;        // Z = n_batch * H independent (batch, head) problems
    int permA[3], permB[3];   // which of (row=0, head=1, batch=2) feeds TMA coordinate 1,2,3
    float* out_f32;
    __half* out_hi;
    __half* out_lo;
    long long ld_out, out_h, out_b;
    int transposed;
    const float* bias;
    int bias_per_row;
    const float* res;
    long long ld_res, res_h, res_b;
    float alpha;
    int act;
};

template <int BN, int PASSES>
struct GemmCfg {
    static constexpr int NPL = PASSES == 3 ? 2 : 1;                 // planes per operand
    static constexpr uint32_t A_TILE = 128 * 128;                   // 128 rows x 128 B
    static constexpr uint32_t B_TILE = BN * 128;
    static constexpr uint32_t STAGE = NPL * (A_TILE + B_TILE);
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr uint32_t SMEM = STAGES * STAGE + 1024;         // + slack for 1024 B alignment
    static constexpr uint32_t TMEM_COLS = 2 * (BN < 32 ? 32 : BN);   // double-buffered accumulator
};

__device__ __forceinline__ void pick_coords(const int (&perm)[3], int row, int h, int b, int& c1, int& c2, int& c3) {
    int v[4] = {row, h, b, 0};      // perm id 3 = broadcast dim (stride 0): coordinate pinned to 0
    c1 = v[perm[0]];
    c2 = v[perm[1]];
    c3 = v[perm[2]];
}

template <int BN, int PASSES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const GemmArgs g) {
    using Cfg = GemmCfg<BN, PASSES>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t acc_full[2];             // MMA -> epilogue: accumulator stage complete
    __shared__ __align__(8) uint64_t acc_empty[2];            // epilogue -> MMA: accumulator stage drained
    __shared__ uint32_t tmem_slot;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int num_kb = (g.K + 63) >> 6;
    // persistent tile loop: tile id -> (z, m block, n block), n fastest so concurrently running CTAs share the A tile in L2
    const int tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + 127) >> 7;
    const long long num_tiles = (long long)tiles_n * tiles_m * g.Z;
    constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAh);
        tma_prefetch_desc(&tmBh);
        if (NPL == 2) {
            tma_prefetch_desc(&tmAl);
            tma_prefetch_desc(&tmBl);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&acc_full[s], 1);
                mbar_init(&acc_empty[s], 4);                    // one arrival per epilogue warp
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, Cfg::TMEM_COLS);                 // 2 accumulator stages
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped
    // the tail of the previous kernel in the stream; its outputs (our operands / residual) are visible after the wait.
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nb = (int)(tile % tiles_n);
                const int mb = (int)((tile / tiles_n) % tiles_m);
                const int z = (int)(tile / ((long long)tiles_n * tiles_m));
                int a1, a2, a3, b1, b2, b3;
                pick_coords(g.permA, mb * 128, z % g.H, z / g.H, a1, a2, a3);
                pick_coords(g.permB, nb * BN, z % g.H, z / g.H, b1, b2, b3);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
                    uint8_t* s = tiles + (size_t)stage * Cfg::STAGE;
                    const int k0 = kb * 64;
                    tma_load_4d(s, &tmAh, &full_bar[stage], k0, a1, a2, a3);
                    if (NPL == 2) tma_load_4d(s + Cfg::A_TILE, &tmAl, &full_bar[stage], k0, a1, a2, a3);
                    tma_load_4d(s + NPL * Cfg::A_TILE, &tmBh, &full_bar[stage], k0, b1, b2, b3);
                    if (NPL == 2) tma_load_4d(s + NPL * Cfg::A_TILE + Cfg::B_TILE, &tmBl, &full_bar[stage], k0, b1, b2, b3);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;                          // accumulator stage (double-buffered TMEM)
                mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1, 4);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                uint32_t accum = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 2);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)stage * Cfg::STAGE);
                    const uint32_t a_hi = sa, a_lo = sa + Cfg::A_TILE;
                    const uint32_t b_hi = sa + NPL * Cfg::A_TILE, b_lo = b_hi + Cfg::B_TILE;
#pragma unroll
                    for (int pass = 0; pass < PASSES; ++pass) {
                        // pass 0: hi*hi, pass 1: hi*lo, pass 2: lo*hi
                        const uint32_t pa = (pass == 2) ? a_lo : a_hi;
                        const uint32_t pb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {       // 4 x (K = 16 fp16 = 32 B) per 128 B swizzle row
                            umma_f16(tmem_d, umma_desc_k128(pa + k4 * 32), umma_desc_k128(pb + k4 * 32), idesc, accum);
                            accum = 1;
                        }
                    }
                    umma_commit(&empty_bar[stage]);            // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&acc_full[as]);                     // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue (4 warps; TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        int it = 0;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int nbk = (int)(tile % tiles_n);
        const int mbk = (int)((tile / tiles_n) % tiles_m);
        const int z = (int)(tile / ((long long)tiles_n * tiles_m));
        const int n0 = nbk * BN, m0 = mbk * 128, h = z % g.H, b = z / g.H;
        const int as = it & 1;
        const uint32_t tmem_acc = tmem_base + as * ACC_COLS;
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < g.M;
        mbar_wait(&acc_full[as], (it >> 1) & 1, 3);
        tc_fence_after();
        const long long zo = (long long)b * g.out_b + (long long)h * g.out_h;
        const long long zr = (long long)b * g.res_b + (long long)h * g.res_h;
        float bias_row = 0.f;
        if (g.bias != nullptr && g.bias_per_row && row_ok) bias_row = __ldg(g.bias + m);
        constexpr int CW = BN < 32 ? 16 : 32;                  // columns per tcgen05.ld
#pragma unroll 1
        for (int c = 0; c < BN / CW; ++c) {
            float v[CW];
            {
                uint32_t r[CW];
                const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * CW);
                if constexpr (CW == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]) * g.alpha;
            }
            const int nb = n0 + c * CW;
            if (nb >= g.N) break;                              // warp-uniform
            const bool full = nb + CW <= g.N;
            if (g.bias != nullptr) {
                if (g.bias_per_row) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] += bias_row;
                } else if (full) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + nb + j));
                        v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        if (nb + j < g.N) v[j] += __ldg(g.bias + nb + j);
                }
            }
            if (g.act == STB_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = gelu_erf(v[j]);
            }
            if (row_ok) {                                      // lanes past M only take part in the collective ld
            if (g.res != nullptr) {
                const float* rp = g.res + zr + (long long)m * g.ld_res + nb;
                if (full && ((g.ld_res & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4) {
                        const float4 rv = *reinterpret_cast<const float4*>(rp + j);
                        v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        if (nb + j < g.N) v[j] += rp[j];
                }
            }
            if (!g.transposed) {
                const long long off = zo + (long long)m * g.ld_out + nb;
                if (g.out_f32 != nullptr) {
                    float* op = g.out_f32 + off;
                    if (full && ((g.ld_out & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < CW; j += 4)
                            *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j)
                            if (nb + j < g.N) op[j] = v[j];
                    }
                }
                if (g.out_hi != nullptr) {
                    __align__(16) __half hi[CW];
                    __align__(16) __half lo[CW];
#pragma unroll
                    for (int j = 0; j < CW; ++j) split_f16(v[j], hi[j], lo[j]);
                    if (full && ((g.ld_out & 7) == 0)) {
#pragma unroll
                        for (int j = 0; j < CW; j += 8) {
                            *reinterpret_cast<uint4*>(g.out_hi + off + j) = *reinterpret_cast<const uint4*>(hi + j);
                            if (g.out_lo != nullptr)
                                *reinterpret_cast<uint4*>(g.out_lo + off + j) = *reinterpret_cast<const uint4*>(lo + j);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j)
                            if (nb + j < g.N) {
                                g.out_hi[off + j] = hi[j];
                                if (g.out_lo != nullptr) g.out_lo[off + j] = lo[j];
                            }
                    }
                }
            } else {
                // transposed store: consecutive lanes (rows m) hit consecutive addresses -> coalesced
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    if (nb + j < g.N) {
                        const long long off = zo + (long long)(nb + j) * g.ld_out + m;
                        if (g.out_f32 != nullptr) g.out_f32[off] = v[j];
                        if (g.out_hi != nullptr) {
                            __half hi, lo;
                            split_f16(v[j], hi, lo);
                            g.out_hi[off] = hi;
                            if (g.out_lo != nullptr) g.out_lo[off] = lo;
                        }
                    }
                }
            }
            }   // row_ok
            __syncwarp();                                      // reconverge before the next .sync.aligned tcgen05.ld
        }
        tc_fence_before();                                     // this warp's TMEM reads of the stage are complete
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        }   // tile loop
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps (cached) + launch
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct TmapKey {
    const void* base;
    int rows, k, H, B, box_rows;
    long long rs, hs, bs;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ p[i]) * 1099511628211ull;
        return (size_t)h;
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey must be a multiple of 8 bytes");

static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, TmapVal, TmapKeyHash> g_tmap_cache;

// 4-D fp16 tensor map over one plane: inner dim = k (contiguous), then (row, head, batch) ordered by increasing stride.
int make_tmap(const void* base, int rows, int k, int H, int B, long long rs, long long hs, long long bs, int box_rows,
              TmapVal* out) {
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base; key.rows = rows; key.k = k; key.H = H; key.B = B; key.box_rows = box_rows;
    key.rs = rs; key.hs = hs; key.bs = bs;
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) {
            *out = it->second;
            return STB_OK;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    STB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    STB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    struct D { unsigned long long size, stride; int id; };
    D d[3] = {{(unsigned long long)rows, (unsigned long long)rs * 2, 0},
              {(unsigned long long)H, (unsigned long long)hs * 2, 1},
              {(unsigned long long)B, (unsigned long long)bs * 2, 2}};
    for (int i = 1; i < 3; ++i)
        if (d[i].stride == 0 || d[i].size == 1) {   // broadcast (e.g. weights shared by every batch item) or degenerate
            d[i].size = 1;
            d[i].id = 3;
        }
    // sort by stride ascending; size-1 dims go last (their stride is irrelevant and is rewritten below)
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) {
            const bool i_last = d[i].size == 1, j_last = d[j].size == 1;
            const bool swap = (i_last && !j_last) || (i_last == j_last && d[j].stride < d[i].stride);
            if (swap) { D t = d[i]; d[i] = d[j]; d[j] = t; }
        }
    unsigned long long prev = 128;   // harmless stride for degenerate dims
    for (int i = 0; i < 3; ++i) {
        if (d[i].size == 1) d[i].stride = prev;
        STB_REQUIRE(d[i].stride % 16 == 0 && d[i].stride > 0, "TMA stride %llu (dim %d) must be a positive multiple of 16 B",
                    d[i].stride, d[i].id);
        prev = d[i].stride * d[i].size;
        if (prev % 16) prev = (prev + 15) / 16 * 16;
    }
    cuuint64_t gdim[4] = {(cuuint64_t)k, d[0].size, d[1].size, d[2].size};
    cuuint64_t gstr[3] = {d[0].stride, d[1].stride, d[2].stride};
    cuuint32_t box[4] = {64, 1, 1, 1};
    for (int i = 0; i < 3; ++i)
        if (d[i].id == 0) box[1 + i] = (cuuint32_t)box_rows;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    TmapVal v;
    CUresult r = enc(&v.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    STB_REQUIRE(r == CUDA_SUCCESS,
                "cuTensorMapEncodeTiled failed (%d): dims {%llu,%llu,%llu,%llu} strides {%llu,%llu,%llu} box rows %d",
                (int)r, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                (unsigned long long)gdim[3], (unsigned long long)gstr[0], (unsigned long long)gstr[1],
                (unsigned long long)gstr[2], box_rows);
    for (int i = 0; i < 3; ++i) v.perm[i] = d[i].id;
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        // keyed by raw pointers: a long-lived process that keeps allocating new buffers would grow it without bound
        if (g_tmap_cache.size() >= 16384) g_tmap_cache.clear();
        g_tmap_cache.emplace(key, v);
    }
    *out = v;
    return STB_OK;
}

template <int BN, int PASSES>
static int launch_gemm(const TmapVal& ah, const TmapVal& al, const TmapVal& bh, const TmapVal& bl, GemmArgs& g,
                       int n_batch, cudaStream_t st) {
    using Cfg = GemmCfg<BN, PASSES>;
    static bool attr_set[64] = {};    // per instantiation and device ordinal (function attributes are per device)
    int dev = 0;
    STB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        STB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    for (int i = 0; i < 3; ++i) {
        g.permA[i] = ah.perm[i];
        g.permB[i] = bh.perm[i];
    }
    g.Z = n_batch * g.H;
    const long long num_tiles = (long long)cdiv(g.N, BN) * cdiv(g.M, 128) * g.Z;
    dim3 grid((unsigned)(num_tiles < sm_count() ? num_tiles : sm_count()));     // persistent: one CTA per SM
    {
        // algorithmic FLOPs of ONE fp32-grade GEMM (not x PASSES); bytes: operands once + output once
        const double zz = (double)n_batch * g.H;
        ProfScope ps("gemm_tc", st, zz * ((double)g.M * g.K + (double)g.N * g.K) * 2.0 * Cfg::NPL + zz * (double)g.M * g.N * 4.0,
                     2.0 * g.M * (double)g.N * g.K * zz);
        STB_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, PASSES>, grid, dim3(192), (size_t)Cfg::SMEM, st, ah.map, al.map, bh.map,
                               bl.map, g));
    }
    STB_LAUNCH_OK();
    return STB_OK;
}

int gemm(const stb_operand& A, const stb_operand& B, int n_batch, int n_head, const stb_epilogue& ep, cudaStream_t st) {
    STB_REQUIRE(A.k == B.k, "gemm: K mismatch %d vs %d", A.k, B.k);
    STB_REQUIRE(A.hi && B.hi, "gemm: hi planes are required");
    STB_REQUIRE((A.lo != nullptr) == (B.lo != nullptr), "gemm: lo planes must be given for both operands or neither");
    STB_REQUIRE(ep.out_f32 || ep.out_hi, "gemm: no output");
    const int passes = A.lo ? 3 : 1;
    const int N = B.rows;
    int BN = N <= 16 ? 16 : N <= 32 ? 32 : N <= 64 ? 64 : 128;
    // 128 x 256 tiles: A 4 KB + B 8 KB of shared-memory reads per 128-cycle MMA (96 B/clk) instead of 128 B/clk -- the
    // shared-memory port limit of a single-CTA 128 x 128 MMA.  Needs enough tiles to fill the SMs (env STB_GEMM_BN256=0 off).
    static const bool bn256 = []() { const char* e = getenv("STB_GEMM_BN256"); return !(e && e[0] == '0'); }();
    if (bn256 && N >= 256 && N % 256 == 0 && (long long)cdiv(N, 256) * cdiv(A.rows, 128) * n_batch * n_head >= 2LL * sm_count()) BN = 256;
    if (A.rows <= 128 && n_batch * n_head == 1) {
        // decode-step shape (M = batch of sequences): HBM-bound on the weights; use narrow N tiles so that enough CTAs
        // (>= ~120 of the 148 SMs) stream them concurrently
        while (BN > 16 && cdiv(N, BN) < 120) BN >>= 1;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = A.rows; g.N = N; g.K = A.k; g.H = n_head;
    g.out_f32 = ep.out_f32; g.out_hi = (__half*)ep.out_hi; g.out_lo = (__half*)ep.out_lo;
    g.ld_out = ep.ld_out; g.out_h = ep.out_h_stride; g.out_b = ep.out_b_stride; g.transposed = ep.transposed;
    g.bias = ep.bias; g.bias_per_row = ep.bias_per_row;
    g.res = ep.residual; g.ld_res = ep.ld_res; g.res_h = ep.res_h_stride; g.res_b = ep.res_b_stride;
    g.alpha = ep.alpha; g.act = ep.act;
    TmapVal ah, al, bh, bl;
    STB_TRY(make_tmap(A.hi, A.rows, A.k, n_head, n_batch, A.row_stride, A.h_stride, A.b_stride, 128, &ah));
    STB_TRY(make_tmap(B.hi, B.rows, B.k, n_head, n_batch, B.row_stride, B.h_stride, B.b_stride, BN, &bh));
    if (passes == 3) {
        STB_TRY(make_tmap(A.lo, A.rows, A.k, n_head, n_batch, A.row_stride, A.h_stride, A.b_stride, 128, &al));
        STB_TRY(make_tmap(B.lo, B.rows, B.k, n_head, n_batch, B.row_stride, B.h_stride, B.b_stride, BN, &bl));
    } else {
        al = ah;
        bl = bh;
    }
#define STB_GEMM_CASE(bn)                                                                   \
    case bn:                                                                                \
        return passes == 3 ? launch_gemm<bn, 3>(ah, al, bh, bl, g, n_batch, st)             \
                           : launch_gemm<bn, 1>(ah, al, bh, bl, g, n_batch, st);
    switch (BN) {
        STB_GEMM_CASE(16)
        STB_GEMM_CASE(32)
        STB_GEMM_CASE(64)
        STB_GEMM_CASE(128)
        STB_GEMM_CASE(256)
    }
#undef STB_GEMM_CASE
    return STB_ERR_UNSUPPORTED;
}

}  // namespace stb

extern "C" int stb_gemm(const stb_operand* A, const stb_operand* B, int n_batch, int n_head, const stb_epilogue* ep,
                        void* stream) {
    STB_REQUIRE(A && B && ep, "stb_gemm: null argument");
    return stb::gemm(*A, *B, n_batch, n_head, *ep, (cudaStream_t)stream);
}
