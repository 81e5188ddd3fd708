// EXPERIMENTAL (STB_DECODE_CHAIN=1; written after round 1's GPU budget was spent, not yet run on hardware):
// a chain of decode-step linear layers in ONE persistent kernel.
//
// The decode step of 17..128 sequences is a chain of single-wave kernels -- per decoder layer 6 swapped split-K tcgen05
// GEMMs and 6 finish kernels at 8..22 us each (profiles/r1_summary_c.md) against ~1 us of HBM time for the weights.  This
// kernel runs a list of such ops back to back on one resident grid (one CTA per SM, cooperative launch):
//
//     op GEMM    swapped split-K GEMM (features on the 128-row M side, sequences on N = BN, the K range cut into `split`
//                slices): same TMA -> smem ring -> tcgen05.mma -> TMEM -> tcgen05.ld pipeline as gemm_tc.cu, partial tiles
//                [split][B][n_feat] to L2; or DIRECT (split = 1, no bias): straight to the output (vocabulary projection)
//     op FINISH  out[b][:] = act(sum_z P[z][b][:] + bias) + res, optionally followed by LayerNorm -> split planes
//                (one sequence per CTA, as splitk_finish_kernel)
//
// with a grid-wide barrier (~2 us: one atomic per CTA + a generation word) instead of a kernel boundary between ops.  The
// smem ring and the mbarrier phases run continuously through all GEMM ops; the TMEM accumulator is allocated once.
//
// Cross-proxy ordering: activations are written with generic stores (finish / epilogue) and read by TMA (async proxy) in the
// next op, so writers issue fence.proxy.async before arriving at the grid barrier and the TMA thread issues it again after.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace stb {

constexpr int CH_MAX_OPS = 10;
constexpr int CH_MAX_GEMMS = 5;
constexpr int CH_THREADS = 192;
enum { CH_OP_GEMM = 0, CH_OP_FINISH = 1 };

struct ChainOp {
    int type;
    // ---- GEMM
    int gemm_slot;                // index into the tensor-map table
    int n_feat, mt, split, kb_per_tile;
    float* partial;               // [split][B][n_feat]  (nullptr: direct)
    float* direct_out;            // direct mode: out[b][ld_direct], no bias
    long long ld_direct;
    // ---- FINISH
    const float* P;
    int f_split, N;
    const float* bias;
    int act;
    const float* res;
    float* out_f32;
    __half* out_hi;
    __half* out_lo;
    long long ld;
    const float* ln_g;
    const float* ln_b;
    __half* ln_hi;
    __half* ln_lo;
};

struct ChainParams {
    int n_ops, B;
    unsigned int* bar;            // [0] arrival count, [1] generation (both persist across launches)
    int permA[CH_MAX_GEMMS][3], permB[CH_MAX_GEMMS][3];
    ChainOp op[CH_MAX_OPS];
};

struct ChainMaps {
    CUtensorMap a_hi[CH_MAX_GEMMS], a_lo[CH_MAX_GEMMS], b_hi[CH_MAX_GEMMS], b_lo[CH_MAX_GEMMS];
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of every CTA call this.  Sense by generation: the last CTA to arrive resets the count and bumps the generation.
__device__ __forceinline__ void grid_barrier(unsigned int* bar) {
    fence_proxy_async_all();                                  // generic writes of this op -> visible to later TMA reads
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int gen = ld_acquire_u32(bar + 1);
        __threadfence();
        const unsigned int prev = atomicAdd(bar, 1u);
        if (prev == gridDim.x - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (ld_acquire_u32(bar + 1) == gen) {
                if (clock64() - t0 > 4000000000LL) {
                    printf("stb: grid barrier timeout block=%d gen=%u count=%u\n", blockIdx.x, gen, bar[0]);
                    __trap();
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void pick3(const int (&perm)[3], int row, int z, int& c1, int& c2, int& c3) {
    int v[4] = {row, 0, z, 0};                                // (row, head = 0, batch = K slice, broadcast)
    c1 = v[perm[0]];
    c2 = v[perm[1]];
    c3 = v[perm[2]];
}

__device__ __forceinline__ float chain_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < CH_THREADS / 32; ++i) t += red[i];
    return t;
}

// One sequence (row) of a FINISH op, all CH_THREADS threads of the CTA (same arithmetic as splitk_finish_kernel).
__device__ void chain_finish_row(const ChainOp& o, int B, int row, float* red) {
    constexpr int LNV = 2;                                    // float4 per thread kept for the LayerNorm (N <= 1536)
    const int nv = o.N >> 2;
    const long long zs = (long long)B * o.N;
    const float* p0 = o.P + (long long)row * o.N;
    const bool do_ln = o.ln_g != nullptr;
    float4 keep[LNV];
    float s = 0.f;
#pragma unroll 1
    for (int it = 0, c4 = threadIdx.x; c4 < nv; c4 += CH_THREADS, ++it) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int z = 0;
        for (; z + 4 <= o.f_split; z += 4) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(p0 + (long long)z * zs + c4 * 4));
            const float4 b = __ldcg(reinterpret_cast<const float4*>(p0 + (long long)(z + 1) * zs + c4 * 4));
            const float4 c = __ldcg(reinterpret_cast<const float4*>(p0 + (long long)(z + 2) * zs + c4 * 4));
            const float4 e = __ldcg(reinterpret_cast<const float4*>(p0 + (long long)(z + 3) * zs + c4 * 4));
            v.x += (a.x + b.x) + (c.x + e.x);
            v.y += (a.y + b.y) + (c.y + e.y);
            v.z += (a.z + b.z) + (c.z + e.z);
            v.w += (a.w + b.w) + (c.w + e.w);
        }
        for (; z < o.f_split; ++z) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(p0 + (long long)z * zs + c4 * 4));
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        if (o.bias != nullptr) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(o.bias) + c4);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        if (o.act == STB_ACT_GELU) {
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        }
        if (o.res != nullptr) {
            const float4 rv = __ldcg(reinterpret_cast<const float4*>(o.res + (long long)row * o.ld + c4 * 4));
            v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
        }
        const long long off = (long long)row * o.ld + c4 * 4;
        if (o.out_f32 != nullptr) *reinterpret_cast<float4*>(o.out_f32 + off) = v;
        if (o.out_hi != nullptr) {
            __half h[4], l[4];
            split_f16(v.x, h[0], l[0]); split_f16(v.y, h[1], l[1]);
            split_f16(v.z, h[2], l[2]); split_f16(v.w, h[3], l[3]);
            *reinterpret_cast<uint2*>(o.out_hi + off) = *reinterpret_cast<uint2*>(h);
            if (o.out_lo != nullptr) *reinterpret_cast<uint2*>(o.out_lo + off) = *reinterpret_cast<uint2*>(l);
        }
        if (do_ln && it < LNV) {
            keep[it] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    }
    if (!do_ln) return;                                       // uniform over the CTA
    const float mean = chain_block_sum(s, red) / (float)o.N;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < LNV; ++it)
        if (threadIdx.x + it * CH_THREADS < nv) {
            const float a = keep[it].x - mean, b = keep[it].y - mean, c = keep[it].z - mean, e = keep[it].w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
    const float rstd = 1.0f / sqrtf(chain_block_sum(q, red) / (float)o.N + 1e-5f);
#pragma unroll
    for (int it = 0; it < LNV; ++it) {
        const int c4 = threadIdx.x + it * CH_THREADS;
        if (c4 < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(o.ln_g) + c4);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(o.ln_b) + c4);
            float4 y;
            y.x = (keep[it].x - mean) * rstd * g.x + bb.x;
            y.y = (keep[it].y - mean) * rstd * g.y + bb.y;
            y.z = (keep[it].z - mean) * rstd * g.z + bb.z;
            y.w = (keep[it].w - mean) * rstd * g.w + bb.w;
            __half h[4], l[4];
            split_f16(y.x, h[0], l[0]); split_f16(y.y, h[1], l[1]);
            split_f16(y.z, h[2], l[2]); split_f16(y.w, h[3], l[3]);
            const long long off = (long long)row * o.N + c4 * 4;
            *reinterpret_cast<uint2*>(o.ln_hi + off) = *reinterpret_cast<uint2*>(h);
            if (o.ln_lo != nullptr) *reinterpret_cast<uint2*>(o.ln_lo + off) = *reinterpret_cast<uint2*>(l);
        }
    }
}

template <int BN, int PASSES>
struct ChainCfg {
    static constexpr int NPL = PASSES == 3 ? 2 : 1;
    static constexpr uint32_t A_TILE = 128 * 128;
    static constexpr uint32_t B_TILE = BN * 128;
    static constexpr uint32_t STAGE = NPL * (A_TILE + B_TILE);
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr uint32_t SMEM = STAGES * STAGE + 1024;
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(CH_THREADS, 1)
gemm_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams P) {
    using Cfg = ChainCfg<BN, PASSES>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t acc_full;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[CH_THREADS / 32];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (warp == 1) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(&acc_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    pdl_trigger();
    pdl_wait();

    // ring / accumulator state of the role threads, continuous over all GEMM ops
    int stage = 0;
    uint32_t phase = 0;          // producer: waits empty[stage] with phase ^ 1; MMA: waits full[stage] with phase
    uint32_t acc_phase = 0;      // MMA commits, epilogue waits

    for (int oi = 0; oi < P.n_ops; ++oi) {
        const ChainOp& o = P.op[oi];
        if (o.type == CH_OP_GEMM) {
            const int g = o.gemm_slot;
            const int tiles_total = o.mt * o.split;
            if (warp == 0) {
                if (lane == 0) {
                    fence_proxy_async_all();                  // operands written by generic stores of the previous op
                    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                        const int mb = tile % o.mt, z = tile / o.mt;
                        int a1, a2, a3, b1, b2, b3;
                        pick3(P.permA[g], mb * 128, z, a1, a2, a3);
                        pick3(P.permB[g], 0, z, b1, b2, b3);
                        for (int kb = 0; kb < o.kb_per_tile; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1, 11);
                            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
                            uint8_t* s = tiles + (size_t)stage * Cfg::STAGE;
                            const int k0 = kb * 64;
                            tma_load_4d(s, &maps.a_hi[g], &full_bar[stage], k0, a1, a2, a3);
                            if (NPL == 2) tma_load_4d(s + Cfg::A_TILE, &maps.a_lo[g], &full_bar[stage], k0, a1, a2, a3);
                            tma_load_4d(s + NPL * Cfg::A_TILE, &maps.b_hi[g], &full_bar[stage], k0, b1, b2, b3);
                            if (NPL == 2)
                                tma_load_4d(s + NPL * Cfg::A_TILE + Cfg::B_TILE, &maps.b_lo[g], &full_bar[stage], k0, b1, b2, b3);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
                __syncwarp();
            } else if (warp == 1) {
                if (lane == 0) {
                    constexpr uint32_t idesc = umma_idesc_f16(128, BN);
                    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                        uint32_t accum = 0;
                        for (int kb = 0; kb < o.kb_per_tile; ++kb) {
                            mbar_wait(&full_bar[stage], phase, 12);
                            tc_fence_after();
                            const uint32_t sa = smem_u32(tiles + (size_t)stage * Cfg::STAGE);
                            const uint32_t a_hi = sa, a_lo = sa + Cfg::A_TILE;
                            const uint32_t b_hi = sa + NPL * Cfg::A_TILE, b_lo = b_hi + Cfg::B_TILE;
#pragma unroll
                            for (int pass = 0; pass < PASSES; ++pass) {
                                const uint32_t pa = (pass == 2) ? a_lo : a_hi;
                                const uint32_t pb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) {
                                    umma_f16(tmem_d, umma_desc_k128(pa + k4 * 32), umma_desc_k128(pb + k4 * 32), idesc, accum);
                                    accum = 1;
                                }
                            }
                            umma_commit(&empty_bar[stage]);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                        umma_commit(&acc_full);
                    }
                }
                __syncwarp();
            } else {
                // epilogue warps 2..5: TMEM lane quarter = warp % 4; transposed store [sequence][feature]
                const int q = warp & 3;
                for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                    const int mb = tile % o.mt, z = tile / o.mt;
                    const int m = mb * 128 + q * 32 + lane;
                    const bool row_ok = m < o.n_feat;
                    mbar_wait(&acc_full, acc_phase, 13);
                    acc_phase ^= 1;
                    tc_fence_after();
                    float* dst = o.partial != nullptr ? o.partial + (long long)z * P.B * o.n_feat : o.direct_out;
                    const long long ldd = o.partial != nullptr ? (long long)o.n_feat : o.ld_direct;
                    constexpr int CW = BN < 32 ? 16 : 32;
#pragma unroll 1
                    for (int c = 0; c < BN / CW; ++c) {
                        uint32_t r[CW];
                        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * CW);
                        if constexpr (CW == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
                        tmem_ld_wait();
                        const int nb = c * CW;
                        if (nb >= P.B) break;                  // warp-uniform
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < CW; ++j)
                                if (nb + j < P.B) dst[(long long)(nb + j) * ldd + m] = __uint_as_float(r[j]);
                        }
                        __syncwarp();
                    }
                    tc_fence_before();
                }
            }
            // NOTE: ops with more than one tile per CTA (the vocabulary projection) rely on the single accumulator being
            // drained before the next tile's first MMA: the MMA thread would have to wait for the epilogue.  Until that
            // handshake exists the host only builds chains whose GEMM ops have tiles <= gridDim.x (checked in decode_chain).
        } else {
            for (int row = blockIdx.x; row < P.B; row += gridDim.x) chain_finish_row(o, P.B, row, red);
        }
        if (oi + 1 < P.n_ops) grid_barrier(P.bar);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool decode_chain_enabled() {
    return option(OPT_DECODE_CHAIN) != 0;
}

struct ChainBuilder {
    ChainMaps maps;
    ChainParams p;
    int n_gemm;
    bool lo;
};

ChainBuilder* chain_new() { return new ChainBuilder(); }
void chain_free(ChainBuilder* cb) { delete cb; }

void chain_begin(ChainBuilder& cb, int B, unsigned int* bar, bool lo) {
    memset(&cb, 0, sizeof(cb));
    cb.p.B = B;
    cb.p.bar = bar;
    cb.lo = lo;
}

static int chain_bn(int B) { return B <= 32 ? 32 : B <= 64 ? 64 : 128; }

// swapped split-K GEMM of W [n][k] (split planes) with x [B][k] -> partial [split][B][n]  (or direct_out when split == 1)
int chain_add_gemm(ChainBuilder& cb, const void* w_hi, const void* w_lo, int n, int k, const void* x_hi, const void* x_lo,
                   int split, float* partial, float* direct_out, long long ld_direct) {
    STB_REQUIRE(cb.p.n_ops < CH_MAX_OPS && cb.n_gemm < CH_MAX_GEMMS, "decode chain: too many ops");
    STB_REQUIRE(k % 64 == 0 && (k / 64) % split == 0, "decode chain: K = %d cannot be cut into %d slices of 64-blocks", k, split);
    const int ks = k / split, g = cb.n_gemm++;
    const int BN = chain_bn(cb.p.B);
    TmapVal ah, al, bh, bl;
    STB_TRY(make_tmap(w_hi, n, ks, 1, split, k, 0, ks, 128, &ah));
    STB_TRY(make_tmap(x_hi, cb.p.B, ks, 1, split, k, 0, ks, BN, &bh));
    if (cb.lo) {
        STB_TRY(make_tmap(w_lo, n, ks, 1, split, k, 0, ks, 128, &al));
        STB_TRY(make_tmap(x_lo, cb.p.B, ks, 1, split, k, 0, ks, BN, &bl));
    } else {
        al = ah;
        bl = bh;
    }
    cb.maps.a_hi[g] = ah.map; cb.maps.a_lo[g] = al.map; cb.maps.b_hi[g] = bh.map; cb.maps.b_lo[g] = bl.map;
    for (int i = 0; i < 3; ++i) {
        cb.p.permA[g][i] = ah.perm[i];
        cb.p.permB[g][i] = bh.perm[i];
    }
    ChainOp& o = cb.p.op[cb.p.n_ops++];
    memset(&o, 0, sizeof(o));
    o.type = CH_OP_GEMM;
    o.gemm_slot = g;
    o.n_feat = n;
    o.mt = cdiv(n, 128);
    o.split = split;
    o.kb_per_tile = ks / 64;
    o.partial = partial;
    o.direct_out = direct_out;
    o.ld_direct = ld_direct;
    return STB_OK;
}

int chain_add_finish(ChainBuilder& cb, const float* P, int split, int N, const float* bias, int act, const float* res,
                     float* out_f32, void* out_hi, void* out_lo, long long ld, const float* ln_g, const float* ln_b, void* ln_hi,
                     void* ln_lo) {
    STB_REQUIRE(cb.p.n_ops < CH_MAX_OPS, "decode chain: too many ops");
    STB_REQUIRE(N % 4 == 0 && ld % 4 == 0, "decode chain: N, ld must be multiples of 4");
    STB_REQUIRE(ln_g == nullptr || N <= 4 * 2 * CH_THREADS, "decode chain: fused LayerNorm needs N <= %d", 4 * 2 * CH_THREADS);
    ChainOp& o = cb.p.op[cb.p.n_ops++];
    memset(&o, 0, sizeof(o));
    o.type = CH_OP_FINISH;
    o.P = P; o.f_split = split; o.N = N; o.bias = bias; o.act = act; o.res = res; o.out_f32 = out_f32;
    o.out_hi = (__half*)out_hi; o.out_lo = (__half*)out_lo; o.ld = ld;
    o.ln_g = ln_g; o.ln_b = ln_b; o.ln_hi = (__half*)ln_hi; o.ln_lo = (__half*)ln_lo;
    return STB_OK;
}

template <int BN, int PASSES>
static int chain_launch_t(const ChainBuilder& cb, cudaStream_t st) {
    using Cfg = ChainCfg<BN, PASSES>;
    static bool attr_set = false;
    if (!attr_set) {
        STB_CUDA_OK(cudaFuncSetAttribute(gemm_chain_kernel<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(sm_count());
    cfg.blockDim = dim3(CH_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;              // all CTAs resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;                                         // PDL and cooperative launch are not combined
    ProfScope ps("gemm_chain", st);
    STB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_chain_kernel<BN, PASSES>, cb.maps, cb.p));
    STB_LAUNCH_OK();
    return STB_OK;
}

int chain_launch(const ChainBuilder& cb, cudaStream_t st) {
    for (int i = 0; i < cb.p.n_ops; ++i)
        if (cb.p.op[i].type == CH_OP_GEMM)
            STB_REQUIRE(cb.p.op[i].mt * cb.p.op[i].split <= sm_count(), "decode chain: GEMM op %d has more tiles than SMs", i);
    const int BN = chain_bn(cb.p.B);
    if (cb.lo) {
        if (BN == 32) return chain_launch_t<32, 3>(cb, st);
        if (BN == 64) return chain_launch_t<64, 3>(cb, st);
        return chain_launch_t<128, 3>(cb, st);
    }
    if (BN == 32) return chain_launch_t<32, 1>(cb, st);
    if (BN == 64) return chain_launch_t<64, 1>(cb, st);
    return chain_launch_t<128, 1>(cb, st);
}

}  // namespace stb
