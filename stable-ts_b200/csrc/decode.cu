// a9 / K10: KV-cached autoregressive decode step (stable_whisper/decode.py:33-65 -> whisper PyTorchInference.logits with
// kv-cache hooks, the logit filters, GreedyDecoder.update).
//
//   stb_decode_step     one decoder forward for the newest token of B sequences: GEMMs on the tcgen05 core (M = B rows,
//                       narrow N tiles so >=120 CTAs stream the weights), attention over the caches on CUDA cores
//                       (one query row per (sequence, head): no tensor-core shape), HBM-bound by weights + cross K/V.
//   stb_sample_greedy   SuppressBlank / SuppressTokens / ApplyTimestampRules / silent-timestamp mask / argmax /
//                       log-prob accumulation / EOT latching fused into one kernel per step, state kept on the device.
//
// Everything position-dependent is read from a DEVICE counter (`pos`), so one captured CUDA graph replays every step.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace stb {

// ---------------------------------------------------------------------------------------------------------
// self-attention over the fp32 K/V cache.  grid (H, B), 128 threads.
//   qkv [B][3d] fp32 (q | k | v of the newest token); caches Kc, Vc [B][ctx][d] fp32; out split [B][d].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
decode_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ Kc, float* __restrict__ Vc, int d, int ctx,
                        const int32_t* __restrict__ pos_ptr, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                        float* __restrict__ out_f32) {
    __shared__ float s_q[64];
    __shared__ float s_p[448 + 32];
    __shared__ float s_red[4];
    __shared__ float s_o[2][64];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const int pos = *pos_ptr;
    const int n = pos + 1;
    const float* row = qkv + (long long)b * 3 * d;
    float* kc = Kc + (long long)b * ctx * d + h * 64;
    float* vc = Vc + (long long)b * ctx * d + h * 64;
    if (tid < 64) {
        s_q[tid] = row[h * 64 + tid];
        kc[(long long)pos * d + tid] = row[d + h * 64 + tid];
    } else {
        vc[(long long)pos * d + (tid - 64)] = row[2 * d + h * 64 + (tid - 64)];
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < n; j += 128) {
        const float4* kr = reinterpret_cast<const float4*>(kc + (long long)j * d);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 kv = kr[c];
            acc = fmaf(s_q[4 * c], kv.x, acc); acc = fmaf(s_q[4 * c + 1], kv.y, acc);
            acc = fmaf(s_q[4 * c + 2], kv.z, acc); acc = fmaf(s_q[4 * c + 3], kv.w, acc);
        }
        acc *= 0.125f;
        s_p[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < n; j += 128) {
        const float e = expf(s_p[j] - mx);
        s_p[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) s_red[tid >> 5] = sum;
    __syncthreads();
    const float inv = 1.0f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
    const int c = tid & 63, half = tid >> 6;
    float acc = 0.f;
    for (int j = half; j < n; j += 2) acc = fmaf(s_p[j], vc[(long long)j * d + c], acc);
    s_o[half][c] = acc;
    __syncthreads();
    if (tid < 64) {
        const float o = (s_o[0][tid] + s_o[1][tid]) * inv;
        if (out_f32) out_f32[(long long)b * d + h * 64 + tid] = o;
        if (out_hi) {
            __half hi, lo;
            split_f16(o, hi, lo);
            out_hi[(long long)b * d + h * 64 + tid] = hi;
            if (out_lo) out_lo[(long long)b * d + h * 64 + tid] = lo;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// cross-attention of one new token per sequence over the per-window K and V, both split fp16 and head-major
// [B][H][T][64]: every (sequence, head) is two contiguous 192 KB streams per plane.  HBM-bound:
// 2 x 1500 x 64 x (2+2) B = 768 KB per (sequence, head) per step.
//
// Flash-decoding layout: the keys of one (b,h) are cut into XS splits; one CTA (4 warps) streams its split ONCE,
// reading K and V rows of the same key together (8 lanes per 128-byte row, 4 keys per warp load, 16 independent
// 16-byte loads in flight per lane), with an online softmax per lane group.  Each CTA writes (m, l, acc[64]); the
// last CTA of a (b,h) to finish (atomic ticket) merges the XS partials and writes the output.
// ---------------------------------------------------------------------------------------------------------
constexpr int XS = 4;                       // key splits per (sequence, head)
constexpr int XS_KEYS = 376;                // keys per split (multiple of 8; 4 x 376 >= 1500)

__device__ __forceinline__ void unpack8(const uint4& a, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 t = __half22float2(h[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}

__global__ void __launch_bounds__(128, 3)
decode_cross_attn_kernel(const float* __restrict__ q, const __half* __restrict__ k_hi, const __half* __restrict__ k_lo,
                         const __half* __restrict__ v_hi, const __half* __restrict__ v_lo, int d, int T,
                         float* __restrict__ partial, int* __restrict__ tickets, __half* __restrict__ out_hi,
                         __half* __restrict__ out_lo, float* __restrict__ out_f32) {
    __shared__ float s_m[4][4], s_l[4][4];
    __shared__ float s_acc[4][4][64];
    __shared__ int s_last;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & 7, grp = lane >> 3;
    const int key0 = split * XS_KEYS, key1 = min(T, key0 + XS_KEYS);
    pdl_trigger();
    pdl_wait();                                              // q comes from the preceding GEMV
    float qr[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[e] = q[(long long)b * d + h * 64 + sub * 8 + e] * 0.125f;
    const long long base = ((long long)b * H + h) * T * 64 + sub * 8;
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // this lane group's keys: key0 + w*4 + grp + 16*i
    for (int j0 = key0 + w * 4 + grp; j0 < key1 + 48; j0 += 64) {          // warp-uniform trip count
        uint4 kh[4], kl[4], vh[4], vl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 16;
            kh[u] = kl[u] = vh[u] = vl[u] = make_uint4(0, 0, 0, 0);
            if (j < key1) {
                const long long off = base + (long long)j * 64;
                kh[u] = __ldg(reinterpret_cast<const uint4*>(k_hi + off));
                vh[u] = __ldg(reinterpret_cast<const uint4*>(v_hi + off));
                if (k_lo) {
                    kl[u] = __ldg(reinterpret_cast<const uint4*>(k_lo + off));
                    vl[u] = __ldg(reinterpret_cast<const uint4*>(v_lo + off));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 16;
            float kf[8], t[8];
            unpack8(kh[u], kf);
            if (k_lo) {
                unpack8(kl[u], t);
#pragma unroll
                for (int e = 0; e < 8; ++e) kf[e] += t[e];
            }
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(qr[e], kf[e], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (j < key1) {                                   // uniform within the 8-lane group
                const float mn = fmaxf(m, s);
                const float corr = expf(m - mn);              // exp(-inf) = 0 on the first key
                const float p = expf(s - mn);
                float vf[8];
                unpack8(vh[u], vf);
                if (v_lo) {
                    unpack8(vl[u], t);
#pragma unroll
                    for (int e = 0; e < 8; ++e) vf[e] += t[e];
                }
                l = l * corr + p;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vf[e], acc[e] * corr);
                m = mn;
            }
        }
    }
    // ---- combine the 16 lane groups of the CTA ----
    if (sub == 0) { s_m[w][grp] = m; s_l[w][grp] = l; }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[w][grp][sub * 8 + e] = acc[e];
    __syncthreads();
    float* part = partial + (((long long)b * H + h) * XS + split) * 66;
    if (threadIdx.x < 64) {
        float M = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) M = fmaxf(M, s_m[i >> 2][i & 3]);
        float Lsum = 0.f, o = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float mi = s_m[i >> 2][i & 3];
            const float sc = (mi == -INFINITY) ? 0.f : expf(mi - M);
            Lsum += s_l[i >> 2][i & 3] * sc;
            o += s_acc[i >> 2][i & 3][threadIdx.x] * sc;
        }
        part[2 + threadIdx.x] = o;
        if (threadIdx.x == 0) { part[0] = M; part[1] = Lsum; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(tickets + b * H + h, 1);
        s_last = (t == XS - 1);
        if (s_last) tickets[b * H + h] = 0;                  // re-arm for the next launch
    }
    __syncthreads();
    if (s_last && threadIdx.x < 64) {
        __threadfence();
        const float* p0 = partial + ((long long)b * H + h) * XS * 66;
        float M = -INFINITY;
#pragma unroll
        for (int i = 0; i < XS; ++i) M = fmaxf(M, __ldcg(p0 + i * 66));
        float Lsum = 0.f, o = 0.f;
#pragma unroll
        for (int i = 0; i < XS; ++i) {
            const float mi = __ldcg(p0 + i * 66);
            const float sc = (mi == -INFINITY) ? 0.f : expf(mi - M);
            Lsum += __ldcg(p0 + i * 66 + 1) * sc;
            o += __ldcg(p0 + i * 66 + 2 + threadIdx.x) * sc;
        }
        o /= Lsum;
        const long long oo = (long long)b * d + h * 64 + threadIdx.x;
        if (out_f32) out_f32[oo] = o;
        if (out_hi) {
            __half hi, lo;
            split_f16(o, hi, lo);
            out_hi[oo] = hi;
            if (out_lo) out_lo[oo] = lo;
        }
    }
}

// V^T split [B][H][64][Tp] -> V split head-major [B][H][T][64] (decode-step layout).  32x32 smem tile transpose.
__global__ void __launch_bounds__(256) v_headmajor_kernel(const __half* __restrict__ vT, int T, int Tp, __half* __restrict__ v) {
    __shared__ __half tile[64][72];
    const long long bh = blockIdx.y;
    const int t0 = blockIdx.x * 64;
    const __half* src = vT + bh * 64 * Tp;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int c = i >> 6, t = i & 63;
        tile[c][t] = (t0 + t < T) ? src[(long long)c * Tp + t0 + t] : __float2half(0.f);
    }
    __syncthreads();
    __half* dst = v + bh * T * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int t = i >> 6, c = i & 63;
        if (t0 + t < T) dst[(long long)(t0 + t) * 64 + c] = tile[c][t];
    }
}

__global__ void embed_step_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ pos_ptr, int d,
                                  const float* __restrict__ emb, const float* __restrict__ posemb, float* __restrict__ x) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const int pos = *pos_ptr;
    const float4* e = reinterpret_cast<const float4*>(emb + (long long)tokens[b] * d);
    const float4* p = reinterpret_cast<const float4*>(posemb + (long long)pos * d);
    float4* o = reinterpret_cast<float4*>(x + (long long)b * d);
    for (int i = threadIdx.x; i < (d >> 2); i += blockDim.x) {
        const float4 a = __ldg(e + i), c = __ldg(p + i);
        o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
}

// ---------------------------------------------------------------------------------------------------------
// fused logit filters + greedy pick.  One CTA (1024 threads) per sequence.
// ---------------------------------------------------------------------------------------------------------
struct BlockRed {
    float f[32];
    int i[32];
};

__device__ __forceinline__ float block_max(float v, BlockRed& r) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) r.f[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = r.f[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, r.f[k]);
    return m;
}
__device__ __forceinline__ float block_sum(float v, BlockRed& r) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) r.f[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += r.f[k];
    return s;
}

__global__ void __launch_bounds__(1024)
sample_greedy_kernel(float* __restrict__ logits, long long ld, int V, int eot, int ts_begin, int no_timestamps,
                     const uint8_t* __restrict__ suppress, const uint8_t* __restrict__ first_mask,
                     const uint8_t* __restrict__ ts_mask, int max_initial_ts, int apply_ts_rules,
                     const int32_t* __restrict__ forced_table, stb_seq_state* __restrict__ states,
                     int32_t* __restrict__ next_out, int32_t* __restrict__ token_table, int32_t* __restrict__ argmax_table,
                     int table_rows) {
    __shared__ BlockRed red;
    __shared__ int s_arg;
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    float* l = logits + (long long)b * ld;
    stb_seq_state st = states[b];
    const bool first = st.n_sampled == 0;
    const bool last_ts = st.n_sampled >= 1 && st.last_tok >= ts_begin;
    const bool penult_ts = st.n_sampled < 2 || st.prev_tok >= ts_begin;
    int ts_floor = -1;                                      // timestamps below this id are forbidden
    if (apply_ts_rules && st.last_ts >= 0) ts_floor = (last_ts && !penult_ts) ? st.last_ts : st.last_ts + 1;
    // ---- pass 1: SuppressBlank / SuppressTokens / ApplyTimestampRules masks; max over text / timestamp ranges
    float mx_text = -INFINITY, mx_ts = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        float v = l[i];
        bool kill = (suppress && suppress[i]) || (first && first_mask && first_mask[i]);
        if (apply_ts_rules) {
            if (i == no_timestamps) kill = true;
            if (last_ts) {
                if (penult_ts) { if (i >= ts_begin) kill = true; }      // pair complete -> text next
                else { if (i < eot) kill = true; }                      // open pair -> timestamp / EOT next
            }
            if (i >= ts_begin && i < ts_floor) kill = true;
            if (first) {
                if (i < ts_begin) kill = true;
                if (max_initial_ts >= 0 && i > ts_begin + max_initial_ts) kill = true;
            }
        }
        if (kill) v = -INFINITY;
        l[i] = v;
        if (i < ts_begin) mx_text = fmaxf(mx_text, v); else mx_ts = fmaxf(mx_ts, v);
    }
    mx_text = block_max(mx_text, red);
    mx_ts = block_max(mx_ts, red);
    bool kill_text = false;
    if (apply_ts_rules && mx_ts > -INFINITY) {              // logsumexp over timestamps vs best text logit
        float s = 0.f;                                      // (evaluated BEFORE the silent-timestamp mask, as the
        for (int i = ts_begin + threadIdx.x; i < V; i += blockDim.x) s += expf(l[i] - mx_ts);   //  reference does)
        s = block_sum(s, red);
        kill_text = (mx_ts + logf(s)) > mx_text;
    }
    __syncthreads();
    // ---- pass 2: timestamp-vs-text rule, silent-timestamp mask (decode.py:14-16,53), nan_to_num_(-inf); final max
    float gmax = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        float v = l[i];
        if (kill_text && i < ts_begin) v = -INFINITY;
        if (ts_mask && i >= ts_begin && ts_mask[i - ts_begin]) v = -INFINITY;
        // logits.nan_to_num_(-inf) (decode.py:55): NaN -> -inf, while -inf / +inf become the lowest / greatest finite
        // float -- so a fully masked row degenerates to a uniform distribution (argmax = index 0), not to NaN
        if (v != v) v = -INFINITY;
        else v = fminf(fmaxf(v, -FLT_MAX), FLT_MAX);
        l[i] = v;
        gmax = fmaxf(gmax, v);
    }
    gmax = block_max(gmax, red);
    __syncthreads();
    // ---- pass 3: argmax (first index on ties) + logsumexp of the final logits
    float s = 0.f;
    int arg = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float v = l[i];
        s += expf(v - gmax);
        if (v == gmax && i < arg) arg = i;
    }
    s = block_sum(s, red);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) arg = min(arg, __shfl_xor_sync(0xffffffffu, arg, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red.i[threadIdx.x >> 5] = arg;
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = red.i[0];
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) a = min(a, red.i[k]);
        s_arg = a;
        const float log_s = logf(s);                          // lse = gmax + log_s; kept apart: gmax may be -FLT_MAX
        const bool was_done = st.n_sampled >= 1 && st.last_tok == eot;
        int next = a;
        const int B = gridDim.x;
        const bool in_table = st.n_sampled < table_rows;
        if (argmax_table && in_table) argmax_table[(long long)st.n_sampled * B + b] = a;
        if (forced_table && in_table) next = forced_table[(long long)st.n_sampled * B + b];
        const float lp = (l[next] - gmax) - log_s;
        if (!was_done) st.sum_logprob += lp;
        if (was_done) next = eot;                           // finished rows keep emitting EOT
        if (token_table && in_table) token_table[(long long)st.n_sampled * B + b] = next;
        st.prev_tok = st.last_tok;
        st.last_tok = next;
        if (next >= ts_begin) st.last_ts = next;
        st.n_sampled += 1;
        st.done = next == eot;
        states[b] = st;
        next_out[b] = next;
    }
}

__global__ void bump_pos_kernel(int32_t* pos) {
    pdl_trigger();
    pdl_wait();
    *pos += 1;
}

}  // namespace stb

extern "C" int stb_sample_greedy(float* logits, long long ld, int B, int V, int eot, int ts_begin, int no_timestamps,
                                 const uint8_t* suppress_mask, const uint8_t* first_step_mask, const uint8_t* ts_mask,
                                 int max_initial_ts, int apply_ts_rules, const int32_t* forced_table, stb_seq_state* states,
                                 int32_t* next_out, int32_t* token_table, int32_t* argmax_table, int table_rows, void* stream) {
    STB_REQUIRE(logits && states && next_out && B >= 1 && V >= 1 && ld >= V, "stb_sample_greedy: bad arguments");
    stb::ProfScope ps("sample_greedy", (cudaStream_t)stream, (double)B * V * 4.0 * 3);
    STB_CUDA_OK(stb::launch_pdl(stb::sample_greedy_kernel, dim3(B), dim3(1024), 0, (cudaStream_t)stream, logits, ld, V, eot, ts_begin,
                                no_timestamps, suppress_mask, first_step_mask, ts_mask, max_initial_ts, apply_ts_rules, forced_table,
                                states, next_out, token_table, argmax_table, table_rows));
    STB_LAUNCH_OK();
    return STB_OK;
}

namespace stb {
int decode_attn_self(const float* qkv, float* Kc, float* Vc, int B, int H, int d, int ctx, const int32_t* pos, __half* oh,
                     __half* ol, float* of, cudaStream_t st) {
    ProfScope ps("decode_self_attn", st);
    STB_CUDA_OK(launch_pdl(decode_self_attn_kernel, dim3(H, B), dim3(128), 0, st, qkv, Kc, Vc, d, ctx, pos, oh, ol, of));
    STB_LAUNCH_OK();
    return STB_OK;
}
int decode_attn_cross(const float* q, const __half* kh, const __half* kl, const __half* vh, const __half* vl, int B, int H,
                      int d, float* partial, int* tickets, __half* oh, __half* ol, float* of, cudaStream_t st) {
    ProfScope ps("decode_cross_attn", st, (double)B * H * 2.0 * STB_N_AUDIO_CTX * 64 * 2.0 * (kl ? 2 : 1));
    STB_CUDA_OK(launch_pdl(decode_cross_attn_kernel, dim3(XS, H, B), dim3(128), 0, st, q, kh, kl, vh, vl, d, (int)STB_N_AUDIO_CTX,
                           partial, tickets, oh, ol, of));
    STB_LAUNCH_OK();
    return STB_OK;
}
size_t decode_cross_scratch_bytes(int B, int H) { return (size_t)B * H * XS * 66 * sizeof(float) + (size_t)B * H * sizeof(int) + 256; }
int v_headmajor(const __half* vT, int BH, int T, int Tp, __half* v, cudaStream_t st) {
    ProfScope ps("v_headmajor", st, (double)BH * T * 64 * 4.0);
    v_headmajor_kernel<<<dim3(cdiv(T, 64), BH), 256, 0, st>>>(vT, T, Tp, v);
    STB_LAUNCH_OK();
    return STB_OK;
}
int embed_step(const int32_t* tokens, const int32_t* pos, int B, int d, const float* emb, const float* posemb, float* x,
               cudaStream_t st) {
    STB_CUDA_OK(launch_pdl(embed_step_kernel, dim3(B), dim3(128), 0, st, tokens, pos, d, emb, posemb, x));
    STB_LAUNCH_OK();
    return STB_OK;
}
int bump_pos(int32_t* pos, cudaStream_t st) {
    STB_CUDA_OK(launch_pdl(bump_pos_kernel, dim3(1), dim3(1), 0, st, pos));
    STB_LAUNCH_OK();
    return STB_OK;
}
}  // namespace stb
