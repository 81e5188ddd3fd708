// K4: fused multi-head attention for sm_100a -- softmax(q k^T / 8) v with the scores kept in TMEM and the
// probabilities handed to the second MMA through shared memory; nothing but q, k, v^T and the output touches HBM.
//
// Replaces MultiHeadAttention.qkv_attention of whisper.model for the encoder's 1500 x 1500 self-attention
// (reached from stable_whisper/timing.py:60 / decode.py:29).  The unfused path (scores fp32 -> HBM -> softmax ->
// split-fp16 probabilities -> HBM -> P.V) moved ~2.2 GB per layer for 4 windows (profiles/r1_summary_a.md).
//
// One CTA = one (batch, head, 128-query tile); 10 warps:
//   warp 0      TMA producer: Q tile once, K tiles through a 2-stage ring (twice: pass 1 and pass 2), V^T tiles (pass 2)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..9  softmax / epilogue: warp w and w+4 share a TMEM lane quarter and split the 128 key columns of a tile
//
// Exact two-pass softmax (same formula as the reference: exp(s - max) / sum), no accumulator rescaling:
//   pass 1   S_t = Q K_t^T  (TMEM, double buffered)  ->  row max m and sum l (online only for the two scalars)
//   pass 2   S_t again      ->  p = exp(s - m) / l  ->  split fp16 (hi, lo) -> shared memory in the K-major SWIZZLE_128B
//            layout the UMMA descriptor expects  ->  O += P_t V_t  (TMEM accumulator, 64 columns)
// Split precision as in the GEMM core: every product is hi*hi + hi*lo + lo*hi (3 MMA passes), fp32 accumulation.
#include "common.cuh"
#include "kernels.h"

namespace stb {

constexpr int AT_BQ = 128;       // queries per CTA
constexpr int AT_BK = 128;       // keys per tile
constexpr int AT_THREADS = 320;

struct AttnArgs {
    int Mq, Mk, H;
    int permQ[3], permK[3], permV[3];
    __half* out_hi;
    __half* out_lo;
    long long ld_out, out_h, out_b;
};

template <int PASSES>
struct AttnCfg {
    static constexpr int NPL = PASSES == 3 ? 2 : 1;
    static constexpr uint32_t Q_TILE = AT_BQ * 128;                 // 16 KB: 128 rows x 128 B
    static constexpr uint32_t K_TILE = AT_BK * 128;                 // 16 KB
    static constexpr uint32_t V_CHUNK = 64 * 128;                   // 8 KB: 64 head dims x 64 keys
    static constexpr uint32_t P_CHUNK = AT_BQ * 128;                // 16 KB: 128 rows x 64 keys
    static constexpr uint32_t OFF_Q = 0;
    static constexpr uint32_t OFF_K = OFF_Q + NPL * Q_TILE;         // 2 stages x NPL x K_TILE
    static constexpr uint32_t OFF_V = OFF_K + 2 * NPL * K_TILE;     // 2 stages x NPL x 2 chunks
    static constexpr uint32_t OFF_P = OFF_V + 2 * NPL * 2 * V_CHUNK;  // NPL x 2 chunks
    static constexpr uint32_t SMEM = OFF_P + NPL * 2 * P_CHUNK + 1024;
};

__device__ __forceinline__ void pick3(const int (&perm)[3], int row, int h, int b, int& c1, int& c2, int& c3) {
    int v[4] = {row, h, b, 0};
    c1 = v[perm[0]];
    c2 = v[perm[1]];
    c3 = v[perm[2]];
}

template <int PASSES>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                    const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                    const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl, const AttnArgs g) {
    using Cfg = AttnCfg<PASSES>;
    constexpr int NPL = Cfg::NPL;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], p_full,
        p_empty, o_full;
    __shared__ uint32_t tmem_slot;
    // (m, l) exchange between the two column halves of a row: lives in the P buffer, which is idle during pass 1
    float* s_xm = reinterpret_cast<float*>(sm + Cfg::OFF_P);      // [2][AT_BQ]
    float* s_xl = s_xm + 2 * AT_BQ;                               // [2][AT_BQ]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int NT = (g.Mk + AT_BK - 1) / AT_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmKh); tma_prefetch_desc(&tmVh);
        if (NPL == 2) { tma_prefetch_desc(&tmQl); tma_prefetch_desc(&tmKl); tma_prefetch_desc(&tmVl); }
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&q_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
                mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
            }
            mbar_init(&p_full, 8); mbar_init(&p_empty, 1); mbar_init(&o_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t TM_S0 = tmem, TM_O = tmem + 256;               // S buffers at columns 0 / 128, O at 256..319

    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            int c1, c2, c3;
            pick3(g.permQ, q0, h, b, c1, c2, c3);
            mbar_arrive_expect_tx(&q_full, NPL * Cfg::Q_TILE);
            tma_load_4d(sm + Cfg::OFF_Q, &tmQh, &q_full, 0, c1, c2, c3);
            if (NPL == 2) tma_load_4d(sm + Cfg::OFF_Q + Cfg::Q_TILE, &tmQl, &q_full, 0, c1, c2, c3);
            for (int i = 0; i < 2 * NT; ++i) {
                const int t = i % NT, st = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&k_empty[st], ph ^ 1, 11);
                mbar_arrive_expect_tx(&k_full[st], NPL * Cfg::K_TILE);
                pick3(g.permK, t * AT_BK, h, b, c1, c2, c3);
                uint8_t* kd = sm + Cfg::OFF_K + st * NPL * Cfg::K_TILE;
                tma_load_4d(kd, &tmKh, &k_full[st], 0, c1, c2, c3);
                if (NPL == 2) tma_load_4d(kd + Cfg::K_TILE, &tmKl, &k_full[st], 0, c1, c2, c3);
                if (i >= NT) {                                   // pass 2: the V^T tile of the same keys
                    const int vs = t & 1;
                    const uint32_t vph = (t >> 1) & 1;
                    mbar_wait(&v_empty[vs], vph ^ 1, 12);
                    mbar_arrive_expect_tx(&v_full[vs], NPL * 2 * Cfg::V_CHUNK);
                    pick3(g.permV, 0, h, b, c1, c2, c3);
                    uint8_t* vd = sm + Cfg::OFF_V + vs * NPL * 2 * Cfg::V_CHUNK;
                    tma_load_4d(vd, &tmVh, &v_full[vs], t * AT_BK, c1, c2, c3);
                    tma_load_4d(vd + Cfg::V_CHUNK, &tmVh, &v_full[vs], t * AT_BK + 64, c1, c2, c3);
                    if (NPL == 2) {
                        tma_load_4d(vd + 2 * Cfg::V_CHUNK, &tmVl, &v_full[vs], t * AT_BK, c1, c2, c3);
                        tma_load_4d(vd + 3 * Cfg::V_CHUNK, &tmVl, &v_full[vs], t * AT_BK + 64, c1, c2, c3);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128);
            constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64);
            const uint32_t q_hi = smem_u32(sm + Cfg::OFF_Q), q_lo = q_hi + Cfg::Q_TILE;
            const uint32_t p_base = smem_u32(sm + Cfg::OFF_P);
            mbar_wait(&q_full, 0, 20);
            auto issue_qk = [&](int i) {
                const int st = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&k_full[st], ph, 21);
                mbar_wait(&s_empty[st], ph ^ 1, 22);
                tc_fence_after();
                const uint32_t k_hi = smem_u32(sm + Cfg::OFF_K + st * NPL * Cfg::K_TILE), k_lo = k_hi + Cfg::K_TILE;
                const uint32_t d = TM_S0 + st * 128;
                uint32_t accum = 0;
#pragma unroll
                for (int pass = 0; pass < PASSES; ++pass) {
                    const uint32_t pa = (pass == 2) ? q_lo : q_hi;
                    const uint32_t pb = (pass == 1) ? k_lo : k_hi;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        umma_f16(d, umma_desc_k128(pa + k4 * 32), umma_desc_k128(pb + k4 * 32), idesc_qk, accum);
                        accum = 1;
                    }
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[st]);
            };
            for (int i = 0; i < NT; ++i) issue_qk(i);           // pass 1
            issue_qk(NT);                                        // pass 2 prologue
            uint32_t o_accum = 0;
            for (int t = 0; t < NT; ++t) {
                if (t + 1 < NT) issue_qk(NT + t + 1);            // overlaps the softmax of tile t
                const int vs = t & 1;
                mbar_wait(&v_full[vs], (t >> 1) & 1, 23);
                mbar_wait(&p_full, t & 1, 24);
                tc_fence_after();
                const uint32_t v_base = smem_u32(sm + Cfg::OFF_V + vs * NPL * 2 * Cfg::V_CHUNK);
#pragma unroll
                for (int pass = 0; pass < PASSES; ++pass) {
                    const uint32_t pa = p_base + ((pass == 2) ? 2 * Cfg::P_CHUNK : 0);
                    const uint32_t pb = v_base + ((pass == 1) ? 2 * Cfg::V_CHUNK : 0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {               // 128 keys = 2 chunks x 4 k-steps of 16
                        const uint32_t a = pa + (j >> 2) * Cfg::P_CHUNK + (j & 3) * 32;
                        const uint32_t bb = pb + (j >> 2) * Cfg::V_CHUNK + (j & 3) * 32;
                        umma_f16(TM_O, umma_desc_k128(a), umma_desc_k128(bb), idesc_pv, o_accum);
                        o_accum = 1;
                    }
                }
                umma_commit(&v_empty[vs]);
                umma_commit(&p_empty);
            }
            umma_commit(&o_full);
        }
        __syncwarp();
    } else {
        // =========================================================== softmax / epilogue (8 warps)
        const int sw = warp - 2;                     // 0..7
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int half = sw >> 2;                    // which 64 key columns of a tile
        const int row = quarter * 32 + lane;         // query row inside the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        constexpr float LOG2E = 1.4426950408889634f;
        const float sc = 0.125f * LOG2E;             // scores in log2 units: exp(x) = exp2(x * log2e)
        float m = -INFINITY, l = 0.f;
        // ---------------- pass 1: row max only (the sum is accumulated in pass 2 and applied to O in the epilogue:
        // softmax(s) v == (sum_j exp(s_j - m) v_j) / (sum_j exp(s_j - m)))
        for (int t = 0; t < NT; ++t) {
            const int st = t & 1;
            mbar_wait(&s_full[st], (t >> 1) & 1, 31);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(TM_S0 + st * 128 + lane_off + half * 64 + c * 32, r);
                tmem_ld_wait();
                const int kbase = t * AT_BK + half * 64 + c * 32;
                if (kbase + 32 <= g.Mk) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (kbase + j < g.Mk) m = fmaxf(m, __uint_as_float(r[j]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
        }
        // combine the two column halves of every row (named barrier over the 256 softmax threads)
        s_xm[half * AT_BQ + row] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(m, s_xm[(half ^ 1) * AT_BQ + row]) * sc;         // row max in log2 units
        asm volatile("bar.sync 1, 256;" ::: "memory");           // exchange area is about to be reused as the P tile
        // ---------------- pass 2: probabilities -> shared memory (A operand of P.V)
        uint8_t* p_hi = sm + Cfg::OFF_P + half * Cfg::P_CHUNK;           // this warp's 64-key chunk
        uint8_t* p_lo = p_hi + 2 * Cfg::P_CHUNK;
        for (int t = 0; t < NT; ++t) {
            const int i = NT + t, st = i & 1;
            mbar_wait(&s_full[st], (i >> 1) & 1, 32);
            tc_fence_after();
            // exponentials + hi/lo split of this warp's 64 keys are computed into registers BEFORE waiting for the P
            // buffer, so they overlap the P.V MMAs of the previous tile; the score buffer is released right after the loads
            uint4 hi4[8], lo4[8];
            const int kt0 = t * AT_BK + half * 64;
            const bool full_tile = kt0 + 64 <= g.Mk;             // warp-uniform: no per-element key masking needed
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(TM_S0 + st * 128 + lane_off + half * 64 + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {                 // 8 keys = one 16-byte swizzle chunk
                    __align__(16) __half hi8[8];
                    __align__(16) __half lo8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = j8 * 8 + e;
                        float p = exp2f(fmaf(__uint_as_float(r[j]), sc, -m));
                        if (!full_tile && kt0 + c * 32 + j >= g.Mk) p = 0.f;
                        l += p;
                        split_f16(p, hi8[e], lo8[e]);
                    }
                    hi4[c * 4 + j8] = *reinterpret_cast<const uint4*>(hi8);
                    lo4[c * 4 + j8] = *reinterpret_cast<const uint4*>(lo8);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);            // scores consumed: Q.K^T of tile t+2 may overwrite them
            mbar_wait(&p_empty, (t & 1) ^ 1, 33);                // P.V of tile t-1 has consumed the P buffer
            // K-major SWIZZLE_128B: row r -> 128 B at r*128; 16-byte chunk index XOR (r & 7)
#pragma unroll
            for (int chunk = 0; chunk < 8; ++chunk) {
                const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(p_hi + off) = hi4[chunk];
                if (NPL == 2) *reinterpret_cast<uint4*>(p_lo + off) = lo4[chunk];
            }
            fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full);
        }
        // ---------------- epilogue: O (TMEM cols 256..319) -> split fp16 -> attention output
        mbar_wait(&o_full, 0, 34);
        tc_fence_after();
        s_xl[half * AT_BQ + row] = l;                              // P buffer is idle again: all P.V MMAs have completed
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv_l = 1.0f / (l + s_xl[(half ^ 1) * AT_BQ + row]);
        {
            uint32_t r[32];
            tmem_ld_32x32(TM_O + lane_off + half * 32, r);
            tmem_ld_wait();
            const int qrow = q0 + row;
            if (qrow < g.Mq) {
                __align__(16) __half hi[32];
                __align__(16) __half lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) split_f16(__uint_as_float(r[j]) * inv_l, hi[j], lo[j]);
                const long long off = (long long)b * g.out_b + (long long)h * g.out_h + (long long)qrow * g.ld_out + half * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    *reinterpret_cast<uint4*>(g.out_hi + off + j) = *reinterpret_cast<const uint4*>(hi + j);
                    if (g.out_lo != nullptr) *reinterpret_cast<uint4*>(g.out_lo + off + j) = *reinterpret_cast<const uint4*>(lo + j);
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int PASSES>
static int launch_attention(const TmapVal (&tm)[6], AttnArgs& g, int n_batch, cudaStream_t st) {
    using Cfg = AttnCfg<PASSES>;
    static bool attr_set[64] = {};    // per device ordinal
    int dev = 0;
    STB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        STB_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    dim3 grid(cdiv(g.Mq, AT_BQ), g.H, n_batch);
    const double zz = (double)n_batch * g.H;
    ProfScope ps("attention_tc", st, zz * ((double)g.Mq * 64 + 2.0 * g.Mk * 64) * 2.0 * Cfg::NPL + zz * g.Mq * 64.0 * 2.0 * Cfg::NPL,
                 zz * 4.0 * g.Mq * (double)g.Mk * 64);
    attention_tc_kernel<PASSES><<<grid, AT_THREADS, Cfg::SMEM, st>>>(tm[0].map, tm[1].map, tm[2].map, tm[3].map, tm[4].map,
                                                                      tm[5].map, g);
    STB_LAUNCH_OK();
    return STB_OK;
}

// q: rows Mq, k = 64; k: rows Mk, k = 64; vT: rows 64, k = keys (>= Mk, zero padded); out split [..][ld_out].
int fused_attention(const stb_operand& q, const stb_operand& k, const stb_operand& vT, int n_batch, int n_head, int Mq, int Mk,
                    void* out_hi, void* out_lo, long long ld_out, long long out_h, long long out_b, cudaStream_t st) {
    STB_REQUIRE(q.k == 64 && k.k == 64 && vT.rows == 64, "fused_attention: head_dim must be 64");
    STB_REQUIRE((q.lo != nullptr) == (k.lo != nullptr) && (q.lo != nullptr) == (vT.lo != nullptr),
                "fused_attention: lo planes must be given for all operands or none");
    STB_REQUIRE((ld_out & 7) == 0, "fused_attention: ld_out must be a multiple of 8");
    const bool lo = q.lo != nullptr;
    TmapVal tm[6];
    STB_TRY(make_tmap(q.hi, Mq, 64, n_head, n_batch, q.row_stride, q.h_stride, q.b_stride, AT_BQ, &tm[0]));
    STB_TRY(make_tmap(k.hi, Mk, 64, n_head, n_batch, k.row_stride, k.h_stride, k.b_stride, AT_BK, &tm[2]));
    STB_TRY(make_tmap(vT.hi, 64, vT.k, n_head, n_batch, vT.row_stride, vT.h_stride, vT.b_stride, 64, &tm[4]));
    if (lo) {
        STB_TRY(make_tmap(q.lo, Mq, 64, n_head, n_batch, q.row_stride, q.h_stride, q.b_stride, AT_BQ, &tm[1]));
        STB_TRY(make_tmap(k.lo, Mk, 64, n_head, n_batch, k.row_stride, k.h_stride, k.b_stride, AT_BK, &tm[3]));
        STB_TRY(make_tmap(vT.lo, 64, vT.k, n_head, n_batch, vT.row_stride, vT.h_stride, vT.b_stride, 64, &tm[5]));
    } else {
        tm[1] = tm[0]; tm[3] = tm[2]; tm[5] = tm[4];
    }
    AttnArgs g;
    g.Mq = Mq; g.Mk = Mk; g.H = n_head;
    for (int i = 0; i < 3; ++i) { g.permQ[i] = tm[0].perm[i]; g.permK[i] = tm[2].perm[i]; g.permV[i] = tm[4].perm[i]; }
    g.out_hi = (__half*)out_hi; g.out_lo = (__half*)out_lo; g.ld_out = ld_out; g.out_h = out_h; g.out_b = out_b;
    return lo ? launch_attention<3>(tm, g, n_batch, st) : launch_attention<1>(tm, g, n_batch, st);
}

}  // namespace stb

extern "C" int stb_attention(const stb_operand* q, const stb_operand* k, const stb_operand* vT, int n_batch, int n_head, int Mq,
                             int Mk, void* out_hi, void* out_lo, long long ld_out, long long out_h_stride,
                             long long out_b_stride, void* stream) {
    STB_REQUIRE(q && k && vT && out_hi && n_batch >= 1 && n_head >= 1 && Mq >= 1 && Mk >= 1, "stb_attention: bad arguments");
    return stb::fused_attention(*q, *k, *vT, n_batch, n_head, Mq, Mk, out_hi, out_lo, ld_out, out_h_stride, out_b_stride,
                                (cudaStream_t)stream);
}
