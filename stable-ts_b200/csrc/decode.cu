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
// cross-attention over the per-window K (split, head-major [B][H][T][64]) and V^T (split [B][H][64][Tp]).
// grid (H, B), 256 threads.  HBM-bound (2 x 1500 x 64 x (2+2) B = 768 KB per (sequence, head) per step):
//   scores: 8 lanes cover one 128-byte key row (hi and lo planes), 4 keys per warp load, 4 loads in flight per lane;
//   output: warp per head-dim row of V^T, lanes read 8 consecutive keys (16 B) per load, probabilities from smem.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot8(const uint4& a, const float* q) {
    const __half2* h = reinterpret_cast<const __half2*>(&a);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc = fmaf(q[2 * e], f.x, acc);
        acc = fmaf(q[2 * e + 1], f.y, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(256)
decode_cross_attn_kernel(const float* __restrict__ q, const __half* __restrict__ k_hi, const __half* __restrict__ k_lo,
                         const __half* __restrict__ v_hi, const __half* __restrict__ v_lo, int d, int T, int Tp,
                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ out_f32) {
    __shared__ __align__(16) float s_p[STB_KPAD + 32];
    __shared__ float s_red[8];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, H = gridDim.x;
    const int w = tid >> 5, lane = tid & 31;
    const int sub = lane & 7, grp = lane >> 3;              // 8 lanes per key row, 4 keys per warp load
    float qr[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[e] = q[(long long)b * d + h * 64 + sub * 8 + e];
    const long long kbase = ((long long)b * H + h) * T * 64 + sub * 8;      // K is head-major [B][H][T][64]
    float mx = -INFINITY;
    // keys handled by this warp: j = it*32 + w*4 + grp
    for (int base = 0; base < T; base += 128) {             // warp-uniform trip count (shuffles below); 4 loads in flight
        const int j0 = base + w * 4 + grp;
        float part[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 32;
            part[u] = 0.f;
            if (j < T) {
                const long long off = kbase + (long long)j * 64;
                part[u] = dot8(__ldg(reinterpret_cast<const uint4*>(k_hi + off)), qr);
                if (k_lo) part[u] += dot8(__ldg(reinterpret_cast<const uint4*>(k_lo + off)), qr);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v = part[u];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            const int j = j0 + u * 32;
            if (j < T) {
                v *= 0.125f;
                if (sub == 0) s_p[j] = v;
                mx = fmaxf(mx, v);
            }
        }
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[w] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < Tp; j += 256) {
        const float e = (j < T) ? expf(s_p[j] - mx) : 0.f;   // pad keys get probability 0
        s_p[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[w] = sum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_red[i];
    const float inv = 1.0f / tot;
    for (int c = w; c < 64; c += 8) {
        const long long off = (((long long)b * H + h) * 64 + c) * Tp;
        // Tp = 1504 = 188 x 8: lane l owns 16-byte groups l, l+32, ... (6 trips); all 12 loads are issued before any use
        uint4 vh[6], vl[6];
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            const int j = (it * 32 + lane) * 8;
            vh[it] = make_uint4(0, 0, 0, 0);
            vl[it] = make_uint4(0, 0, 0, 0);
            if (j < Tp) {
                vh[it] = __ldg(reinterpret_cast<const uint4*>(v_hi + off + j));
                if (v_lo) vl[it] = __ldg(reinterpret_cast<const uint4*>(v_lo + off + j));
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            const int j = (it * 32 + lane) * 8;
            if (j < Tp) {                                   // V^T pad columns are zero and p[pad] is zero
                acc += dot8(vh[it], s_p + j);
                if (v_lo) acc += dot8(vl[it], s_p + j);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float o = acc * inv;
            if (out_f32) out_f32[(long long)b * d + h * 64 + c] = o;
            if (out_hi) {
                __half hi, lo;
                split_f16(o, hi, lo);
                out_hi[(long long)b * d + h * 64 + c] = hi;
                if (out_lo) out_lo[(long long)b * d + h * 64 + c] = lo;
            }
        }
    }
}

__global__ void embed_step_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ pos_ptr, int d,
                                  const float* __restrict__ emb, const float* __restrict__ posemb, float* __restrict__ x) {
    const int b = blockIdx.x;
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

__global__ void bump_pos_kernel(int32_t* pos) { *pos += 1; }

}  // namespace stb

extern "C" int stb_sample_greedy(float* logits, long long ld, int B, int V, int eot, int ts_begin, int no_timestamps,
                                 const uint8_t* suppress_mask, const uint8_t* first_step_mask, const uint8_t* ts_mask,
                                 int max_initial_ts, int apply_ts_rules, const int32_t* forced_table, stb_seq_state* states,
                                 int32_t* next_out, int32_t* token_table, int32_t* argmax_table, int table_rows, void* stream) {
    STB_REQUIRE(logits && states && next_out && B >= 1 && V >= 1 && ld >= V, "stb_sample_greedy: bad arguments");
    stb::ProfScope ps("sample_greedy", (cudaStream_t)stream, (double)B * V * 4.0 * 3);
    stb::sample_greedy_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(logits, ld, V, eot, ts_begin, no_timestamps, suppress_mask,
                                                                     first_step_mask, ts_mask, max_initial_ts, apply_ts_rules,
                                                                     forced_table, states, next_out, token_table, argmax_table,
                                                                     table_rows);
    STB_LAUNCH_OK();
    return STB_OK;
}

namespace stb {
int decode_attn_self(const float* qkv, float* Kc, float* Vc, int B, int H, int d, int ctx, const int32_t* pos, __half* oh,
                     __half* ol, float* of, cudaStream_t st) {
    ProfScope ps("decode_self_attn", st);
    decode_self_attn_kernel<<<dim3(H, B), 128, 0, st>>>(qkv, Kc, Vc, d, ctx, pos, oh, ol, of);
    STB_LAUNCH_OK();
    return STB_OK;
}
int decode_attn_cross(const float* q, const __half* kh, const __half* kl, const __half* vh, const __half* vl, int B, int H,
                      int d, __half* oh, __half* ol, float* of, cudaStream_t st) {
    ProfScope ps("decode_cross_attn", st, (double)B * H * 2.0 * STB_N_AUDIO_CTX * 64 * 2.0 * (kl ? 2 : 1));
    decode_cross_attn_kernel<<<dim3(H, B), 256, 0, st>>>(q, kh, kl, vh, vl, d, STB_N_AUDIO_CTX, STB_KPAD, oh, ol, of);
    STB_LAUNCH_OK();
    return STB_OK;
}
int embed_step(const int32_t* tokens, const int32_t* pos, int B, int d, const float* emb, const float* posemb, float* x,
               cudaStream_t st) {
    embed_step_kernel<<<B, 128, 0, st>>>(tokens, pos, d, emb, posemb, x);
    STB_LAUNCH_OK();
    return STB_OK;
}
int bump_pos(int32_t* pos, cudaStream_t st) {
    bump_pos_kernel<<<1, 1, 0, st>>>(pos);
    STB_LAUNCH_OK();
    return STB_OK;
}
}  // namespace stb
