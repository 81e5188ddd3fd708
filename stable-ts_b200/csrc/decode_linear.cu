// Linear layers of the KV-cached decode step for 17..128 sequences (stable_whisper/decode.py:40 -> whisper's
// TextDecoder on the newest token): out[s][n] = epilogue( sum_k x[s][k] * W[n][k] ), a few hundred MB of weights per
// step streamed once, ~120 sequences: HBM-bound on the weights, and -- with ~200 such GEMMs per step -- latency-bound
// on everything between them.  Round 1 ran each as a swapped split-K tcgen05 GEMM writing fp32 partials to L2 plus a
// second "finish" kernel (bias / GELU / residual): 12 dependent launches per decoder layer, ~0.14 of the HBM floor.
//
// This kernel does one Linear in ONE launch with no partials in global memory:
//   * SWAPPED: the 128 output features of a tile sit on the MMA's M side (every weight byte is read by exactly one CTA),
//     the sequences on the N side (BN = 32 / 64 / 128 >= B, rows past B are zero-filled by TMA);
//   * split-K across a THREAD-BLOCK CLUSTER: the CTAs of a cluster own the same feature tile and consecutive K ranges, so
//     tiles x split ~ 80-120 CTAs stream the weights concurrently;
//   * each CTA: TMA (SWIZZLE_128B, mbarrier ring) -> single-thread tcgen05.mma kind::f16 (hi*hi + hi*lo + lo*hi passes in
//     parity mode) -> fp32 accumulator in TMEM -> registers -> its OWN shared memory as an fp32 [sequence][feature] tile;
//   * cluster barrier; CTA r then sums column slice r of all peers' tiles over distributed shared memory
//     (ld.shared::cluster, fixed summation order: bit-reproducible) and applies bias / GELU / residual, storing fp32 and/or
//     split-fp16 rows [sequence][feature] coalesced along the features;
//   * programmatic dependent launch: barrier init, TMEM allocation and the WEIGHT tiles of the first ring stages are issued
//     before griddepcontrol.wait (weights do not depend on the previous kernel); only the activation tiles wait.
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace stb {

struct DLArgs {
    int n, B, kb_total;       // output features, sequences, 64-wide k blocks
    const float* bias;
    int act;
    const float* res;
    long long ld_res;
    float* out_f32;
    __half* out_hi;
    __half* out_lo;
    long long ld_out;
    // LayerNorm folded into this Linear (DLFuse): out = rstd_s (acc - mean_s wsum[n]) + bias[n], the row statistics of
    // sequence s summed from ln_tiles partial (sum, sum of squares) pairs [ln_tiles][B][2] left by the producer of x
    const float* ln_stats;
    int ln_tiles;
    float ln_dinv;            // 1 / (features of the normalised row)
    const float* ln_wsum;     // [n]
    // this Linear produces a row that the NEXT Linear normalises: per 128-feature tile the (sum, sum of squares) of the
    // final values (after bias / activation / residual) of every sequence, [tiles][B][2]
    float* stats_out;
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// exit barrier: nothing is published, the peers only need to know that this CTA has finished READING their tiles
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// 16-byte load from a peer CTA's shared memory (volatile + memory clobber keep it behind the cluster barrier; the loads of
// one reduction step go to distinct registers, so they still issue back to back and their latencies overlap)
__device__ __forceinline__ float4 dsmem_ld4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int BN, int PASSES>
struct DLCfg {
    static constexpr int NPL = PASSES == 3 ? 2 : 1;
    static constexpr uint32_t A_TILE = 128 * 128;             // 128 features x 64 fp16
    static constexpr uint32_t B_TILE = BN * 128;               // BN sequences x 64 fp16
    static constexpr uint32_t STAGE = NPL * (A_TILE + B_TILE);
    static constexpr uint32_t PART = BN * 128 * 4;             // fp32 [BN sequences][128 features], aliases the ring
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr uint32_t RING = STAGES * STAGE;
    static constexpr uint32_t SMEM = (RING > PART ? RING : PART) + 1024;
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(192, 1)
decode_linear_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                     const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl, const DLArgs g) {
    using Cfg = DLCfg<BN, PASSES>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t acc_full;
    __shared__ uint32_t tmem_slot;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int rank = (int)cluster_rank(), split = (int)cluster_size();
    const int m0 = (blockIdx.x / split) * 128;                 // first output feature of this cluster's tile
    const int kb0 = (int)((long long)rank * g.kb_total / split), kb1 = (int)((long long)(rank + 1) * g.kb_total / split);
    const int nk = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmWh);
        tma_prefetch_desc(&tmXh);
        if (NPL == 2) {
            tma_prefetch_desc(&tmWl);
            tma_prefetch_desc(&tmXl);
        }
    }
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
    const uint32_t tmem_base = tmem_slot;
    pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            const int pre = nk < STAGES ? nk : STAGES;
            // weights of the first ring stages: independent of the previous kernel, issued before the dependency wait
            for (int i = 0; i < pre; ++i) {
                uint8_t* s = tiles + (size_t)i * Cfg::STAGE;
                mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE);
                const int k0 = (kb0 + i) * 64;
                tma_load_4d(s, &tmWh, &full_bar[i], k0, m0, 0, 0);
                if (NPL == 2) tma_load_4d(s + Cfg::A_TILE, &tmWl, &full_bar[i], k0, m0, 0, 0);
            }
            pdl_wait();                                         // the activations are the previous kernel's output
            for (int i = 0; i < pre; ++i) {
                uint8_t* s = tiles + (size_t)i * Cfg::STAGE;
                const int k0 = (kb0 + i) * 64;
                tma_load_4d(s + NPL * Cfg::A_TILE, &tmXh, &full_bar[i], k0, 0, 0, 0);
                if (NPL == 2) tma_load_4d(s + NPL * Cfg::A_TILE + Cfg::B_TILE, &tmXl, &full_bar[i], k0, 0, 0, 0);
            }
            int stage = pre == STAGES ? 0 : pre;
            uint32_t phase = pre == STAGES ? 1 : 0;
            for (int i = pre; i < nk; ++i) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 11);
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
                uint8_t* s = tiles + (size_t)stage * Cfg::STAGE;
                const int k0 = (kb0 + i) * 64;
                tma_load_4d(s, &tmWh, &full_bar[stage], k0, m0, 0, 0);
                if (NPL == 2) tma_load_4d(s + Cfg::A_TILE, &tmWl, &full_bar[stage], k0, m0, 0, 0);
                tma_load_4d(s + NPL * Cfg::A_TILE, &tmXh, &full_bar[stage], k0, 0, 0, 0);
                if (NPL == 2) tma_load_4d(s + NPL * Cfg::A_TILE + Cfg::B_TILE, &tmXl, &full_bar[stage], k0, 0, 0, 0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            int stage = 0;
            uint32_t phase = 0, accum = 0;
            for (int i = 0; i < nk; ++i) {
                mbar_wait(&full_bar[stage], phase, 12);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + (size_t)stage * Cfg::STAGE);
                const uint32_t a_hi = sa, a_lo = sa + Cfg::A_TILE;
                const uint32_t b_hi = sa + NPL * Cfg::A_TILE, b_lo = b_hi + Cfg::B_TILE;
#pragma unroll
                for (int pass = 0; pass < PASSES; ++pass) {    // pass 0: hi*hi, 1: hi*lo, 2: lo*hi
                    const uint32_t pa = (pass == 2) ? a_lo : a_hi;
                    const uint32_t pb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        umma_f16(tmem_base, umma_desc_k128(pa + k4 * 32), umma_desc_k128(pb + k4 * 32), idesc, accum);
                        accum = 1;
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(&acc_full);                             // arrives at once when nk == 0 (nothing outstanding)
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ accumulator -> own shared memory, fp32 [sequence][feature]
        const int q = warp & 3;
        float* part = reinterpret_cast<float*>(tiles);
        const int t = q * 32 + lane;                           // feature row inside the tile == TMEM lane
        mbar_wait(&acc_full, 0, 13);
        tc_fence_after();
        constexpr int CW = BN < 32 ? 16 : 32;
#pragma unroll 1
        for (int c = 0; c < BN / CW; ++c) {
            uint32_t r[CW];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * CW);
            if (nk > 0) {
                if constexpr (CW == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < CW; ++j) r[j] = 0u;       // empty K range: contributes zeros
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) part[(c * CW + j) * 128 + t] = __uint_as_float(r[j]);
        }
        tc_fence_before();
    }
    __syncwarp();
    cluster_sync_all();                                         // every peer's tile is complete and visible
    if (warp >= 2) {
        // Reduce-scatter over distributed shared memory: rank r owns sequences [r BN / split, (r + 1) BN / split).  A warp
        // handles one sequence per iteration, a lane 4 consecutive features of it: it issues the 16-byte loads of ALL peers
        // first (independent, so their DSMEM latencies overlap), adds them in rank order (bit-reproducible), applies the
        // epilogue and stores 16 bytes.  Every lane walks the loop (warp-uniform bounds): the folded LayerNorm and the row
        // statistics below are warp reductions.
        const int e = (warp - 2) * 32 + lane;                  // 0 .. 127
        const int quad = e & 31;                               // features m0 + 4 quad .. + 3
        const int m = m0 + quad * 4;
        const int c0 = rank * BN / split, c1 = (rank + 1) * BN / split;
        const uint32_t local = smem_u32(tiles);
        uint32_t peer[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) peer[p] = p < split ? dsmem_addr(local, (uint32_t)p) : 0u;
        pdl_wait();                                             // residual and row statistics are earlier kernels' outputs
        const bool full = m + 3 < g.n;                          // all 4 features exist
        const bool vec = full && (g.ld_out & 3) == 0 && (g.res == nullptr || (g.ld_res & 3) == 0);
        auto load4 = [&](const float* p) -> float4 {            // 4 per-feature constants, zero past n
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p == nullptr) return r;
            if (full) return __ldg(reinterpret_cast<const float4*>(p + m));
            if (m < g.n) r.x = __ldg(p + m);
            if (m + 1 < g.n) r.y = __ldg(p + m + 1);
            if (m + 2 < g.n) r.z = __ldg(p + m + 2);
            return r;
        };
        const float4 bias = load4(g.bias);
        const float4 wsum = load4(g.ln_wsum);
        const int tile = blockIdx.x / split;
        for (int c = c0 + (e >> 5); c < c1 && c < g.B; c += 4) {
            const uint32_t off = (uint32_t)((c * 128 + quad * 4) * 4);
            float4 v[8];
#pragma unroll
            for (int p = 0; p < 8; ++p)
                if (p < split) v[p] = dsmem_ld4(peer[p] + off);
            float4 a = v[0];
#pragma unroll
            for (int p = 1; p < 8; ++p)
                if (p < split) { a.x += v[p].x; a.y += v[p].y; a.z += v[p].z; a.w += v[p].w; }
            if (g.ln_stats != nullptr) {                        // folded LayerNorm of the input row of sequence c
                float s1 = 0.f, s2 = 0.f;
                if (quad < g.ln_tiles) {
                    const float2 pr = *reinterpret_cast<const float2*>(g.ln_stats + ((long long)quad * g.B + c) * 2);
                    s1 = pr.x;
                    s2 = pr.y;
                }
                s1 = warp_sum(s1);                              // fixed shuffle tree: reproducible
                s2 = warp_sum(s2);
                const double mu = (double)s1 * (double)g.ln_dinv;
                const double var = (double)s2 * (double)g.ln_dinv - mu * mu;
                const float rstd = (float)(1.0 / sqrt(var + 1e-5));
                const float muf = (float)mu;
                a.x = (a.x - muf * wsum.x) * rstd; a.y = (a.y - muf * wsum.y) * rstd;
                a.z = (a.z - muf * wsum.z) * rstd; a.w = (a.w - muf * wsum.w) * rstd;
            }
            a.x += bias.x; a.y += bias.y; a.z += bias.z; a.w += bias.w;
            if (g.act == STB_ACT_GELU) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
            if (g.res != nullptr) {
                const float* rp = g.res + (long long)c * g.ld_res + m;
                if (vec) {
                    const float4 r = *reinterpret_cast<const float4*>(rp);
                    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                } else {
                    if (m < g.n) a.x += rp[0];
                    if (m + 1 < g.n) a.y += rp[1];
                    if (m + 2 < g.n) a.z += rp[2];
                    if (m + 3 < g.n) a.w += rp[3];
                }
            }
            if (g.stats_out != nullptr) {                       // partial row statistics of this 128-feature tile
                float s1 = 0.f, s2 = 0.f;
                if (m < g.n) { s1 += a.x; s2 += a.x * a.x; }
                if (m + 1 < g.n) { s1 += a.y; s2 += a.y * a.y; }
                if (m + 2 < g.n) { s1 += a.z; s2 += a.z * a.z; }
                if (m + 3 < g.n) { s1 += a.w; s2 += a.w * a.w; }
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (quad == 0) *reinterpret_cast<float2*>(g.stats_out + ((long long)tile * g.B + c) * 2) = make_float2(s1, s2);
            }
            const long long off_o = (long long)c * g.ld_out + m;
            if (vec) {
                if (g.out_f32 != nullptr) *reinterpret_cast<float4*>(g.out_f32 + off_o) = a;
                if (g.out_hi != nullptr) {
                    __half h[4], l[4];
                    split_f16(a.x, h[0], l[0]); split_f16(a.y, h[1], l[1]);
                    split_f16(a.z, h[2], l[2]); split_f16(a.w, h[3], l[3]);
                    *reinterpret_cast<uint2*>(g.out_hi + off_o) = *reinterpret_cast<const uint2*>(h);
                    if (g.out_lo != nullptr) *reinterpret_cast<uint2*>(g.out_lo + off_o) = *reinterpret_cast<const uint2*>(l);
                }
            } else {
                const float vals[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (m + j >= g.n) break;
                    if (g.out_f32 != nullptr) g.out_f32[off_o + j] = vals[j];
                    if (g.out_hi != nullptr) {
                        __half hi, lo;
                        split_f16(vals[j], hi, lo);
                        g.out_hi[off_o + j] = hi;
                        if (g.out_lo != nullptr) g.out_lo[off_o + j] = lo;
                    }
                }
            }
        }
    }
    __syncwarp();
    cluster_sync_relaxed();                                     // nobody exits while a peer may still read its tile
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// cluster size in 1..8 (every size up to 8 is portable) that puts the most CTAs on the machine: every rank gets at least
// one k block, and all clusters are co-resident with margin for GPC granularity (a cluster lives inside one GPC: with
// 16 / 18 / 20 SMs per GPC, 128 CTAs fit for every cluster size)
int decode_linear_split(int n, int k) {
    const int mt = cdiv(n, 128), nkb = k / 64;
    const int budget = sm_count() >= 140 ? 128 : (sm_count() * 7) / 8;
    int split = 1;
    for (int s = 2; s <= 8; ++s)
        if (s <= nkb && (long long)mt * s <= budget) split = s;
    return split;
}

template <int BN, int PASSES>
static int launch_dl(const TmapVal& wh, const TmapVal& wl, const TmapVal& xh, const TmapVal& xl, const DLArgs& g, int split,
                     cudaStream_t st) {
    using Cfg = DLCfg<BN, PASSES>;
    static bool attr_set[64] = {};
    int dev = 0;
    STB_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        STB_CUDA_OK(cudaFuncSetAttribute(decode_linear_kernel<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set[dev] = true;
    }
    const int mt = cdiv(g.n, 128);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(mt * split));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[3];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)split;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled(1)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (launch_priority() != 0) {
        attr[na].id = cudaLaunchAttributePriority;
        attr[na].val.priority = launch_priority();
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    {
        const double wbytes = (double)g.n * g.kb_total * 64 * 2.0 * Cfg::NPL;
        ProfScope ps("decode_linear", st, wbytes + (double)g.B * g.kb_total * 64 * 2.0 * Cfg::NPL + (double)g.B * g.n * 4.0,
                     2.0 * g.n * (double)g.B * g.kb_total * 64);
        STB_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_linear_kernel<BN, PASSES>, wh.map, wl.map, xh.map, xl.map, g));
    }
    STB_LAUNCH_OK();
    return STB_OK;
}

// out[B][n] = epilogue(x[B][k] . W[n][k]^T); x / W split-fp16 planes, K-major; B <= 128; k a multiple of 64.
int decode_linear(const void* x_hi, const void* x_lo, int B, int k, const void* w_hi, const void* w_lo, int n, const float* bias,
                  int act, const float* res, long long ld_res, float* out_f32, void* out_hi, void* out_lo, long long ld_out,
                  const DLFuse* fuse, cudaStream_t st) {
    STB_REQUIRE(x_hi && w_hi && (out_f32 || out_hi), "decode_linear: null argument");
    STB_REQUIRE(B >= 1 && B <= 128 && k % 64 == 0 && n >= 1, "decode_linear: unsupported shape B=%d k=%d n=%d", B, k, n);
    STB_REQUIRE((x_lo != nullptr) == (w_lo != nullptr), "decode_linear: lo planes must be given for both operands or neither");
    const int BN = B <= 32 ? 32 : B <= 64 ? 64 : 128;
    DLArgs g;
    memset(&g, 0, sizeof(g));
    g.n = n; g.B = B; g.kb_total = k / 64;
    g.bias = bias; g.act = act; g.res = res; g.ld_res = ld_res;
    g.out_f32 = out_f32; g.out_hi = (__half*)out_hi; g.out_lo = (__half*)out_lo; g.ld_out = ld_out;
    if (fuse != nullptr) {
        if (fuse->stats_in != nullptr) {
            STB_REQUIRE(fuse->wsum != nullptr && fuse->tiles_in >= 1 && fuse->tiles_in <= 32 && fuse->row_features >= 1,
                        "decode_linear: folded LayerNorm needs wsum and 1..32 statistic tiles");
            g.ln_stats = fuse->stats_in; g.ln_tiles = fuse->tiles_in; g.ln_dinv = 1.0f / (float)fuse->row_features;
            g.ln_wsum = fuse->wsum;
        }
        g.stats_out = fuse->stats_out;
    }
    const int split = decode_linear_split(n, k);
    TmapVal wh, wl, xh, xl;
    STB_TRY(make_tmap(w_hi, n, k, 1, 1, k, 0, 0, 128, &wh));
    STB_TRY(make_tmap(x_hi, B, k, 1, 1, k, 0, 0, BN, &xh));
    if (w_lo) {
        STB_TRY(make_tmap(w_lo, n, k, 1, 1, k, 0, 0, 128, &wl));
        STB_TRY(make_tmap(x_lo, B, k, 1, 1, k, 0, 0, BN, &xl));
    } else {
        wl = wh;
        xl = xh;
    }
#define STB_DL_CASE(bn)                                                                                   \
    case bn:                                                                                              \
        return w_lo ? launch_dl<bn, 3>(wh, wl, xh, xl, g, split, st) : launch_dl<bn, 1>(wh, wl, xh, xl, g, split, st);
    switch (BN) {
        STB_DL_CASE(32)
        STB_DL_CASE(64)
        STB_DL_CASE(128)
    }
#undef STB_DL_CASE
    return STB_ERR_UNSUPPORTED;
}

}  // namespace stb
