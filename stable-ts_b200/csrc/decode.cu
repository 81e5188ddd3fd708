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
                        const int32_t* __restrict__ pos_ptr, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
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
        __half hi, lo;
        split_f16(o, hi, lo);
        out_hi[(long long)b * d + h * 64 + tid] = hi;
        if (out_lo) out_lo[(long long)b * d + h * 64 + tid] = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------
// cross-attention over the per-window K (split [B*T][d]) and V^T (split [B][H][64][Tp]).  grid (H, B), 256 threads.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_cross_attn_kernel(const float* __restrict__ q, const __half* __restrict__ k_hi, const __half* __restrict__ k_lo,
                         const __half* __restrict__ v_hi, const __half* __restrict__ v_lo, int d, int T, int Tp,
                         __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    __shared__ float s_q[64];
    __shared__ float s_p[STB_KPAD];
    __shared__ float s_red[8];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, H = gridDim.x;
    if (tid < 64) s_q[tid] = q[(long long)b * d + h * 64 + tid];
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < T; j += 256) {
        const long long off = ((long long)b * T + j) * d + h * 64;
        const uint4* rh = reinterpret_cast<const uint4*>(k_hi + off);
        const uint4* rl = k_lo ? reinterpret_cast<const uint4*>(k_lo + off) : nullptr;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 a = rh[c];
            const __half2* ah = reinterpret_cast<const __half2*>(&a);
            float2 f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = __half22float2(ah[e]);
            if (rl) {
                const uint4 l = rl[c];
                const __half2* lh = reinterpret_cast<const __half2*>(&l);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 g = __half22float2(lh[e]);
                    f[e].x += g.x;
                    f[e].y += g.y;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc = fmaf(s_q[8 * c + 2 * e], f[e].x, acc);
                acc = fmaf(s_q[8 * c + 2 * e + 1], f[e].y, acc);
            }
        }
        acc *= 0.125f;
        s_p[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < T; j += 256) {
        const float e = expf(s_p[j] - mx);
        s_p[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) s_red[tid >> 5] = sum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_red[i];
    const float inv = 1.0f / tot;
    const int w = tid >> 5, lane = tid & 31;
    for (int c = w; c < 64; c += 8) {                          // warp per output dim, lanes across keys (coalesced)
        const long long off = (((long long)b * H + h) * 64 + c) * Tp;
        float acc = 0.f;
        for (int j = lane; j < T; j += 32) {
            float v = __half2float(v_hi[off + j]);
            if (v_lo) v += __half2float(v_lo[off + j]);
            acc = fmaf(s_p[j], v, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            __half hi, lo;
            split_f16(acc * inv, hi, lo);
            out_hi[(long long)b * d + h * 64 + c] = hi;
            if (out_lo) out_lo[(long long)b * d + h * 64 + c] = lo;
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
    stb::sample_greedy_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(logits, ld, V, eot, ts_begin, no_timestamps, suppress_mask,
                                                                     first_step_mask, ts_mask, max_initial_ts, apply_ts_rules,
                                                                     forced_table, states, next_out, token_table, argmax_table,
                                                                     table_rows);
    STB_LAUNCH_OK();
    return STB_OK;
}

namespace stb {
int decode_attn_self(const float* qkv, float* Kc, float* Vc, int B, int H, int d, int ctx, const int32_t* pos, __half* oh,
                     __half* ol, cudaStream_t st) {
    decode_self_attn_kernel<<<dim3(H, B), 128, 0, st>>>(qkv, Kc, Vc, d, ctx, pos, oh, ol);
    STB_LAUNCH_OK();
    return STB_OK;
}
int decode_attn_cross(const float* q, const __half* kh, const __half* kl, const __half* vh, const __half* vl, int B, int H,
                      int d, __half* oh, __half* ol, cudaStream_t st) {
    decode_cross_attn_kernel<<<dim3(H, B), 256, 0, st>>>(q, kh, kl, vh, vl, d, STB_N_AUDIO_CTX, STB_KPAD, oh, ol);
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
