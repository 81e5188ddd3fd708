// a5 / K6+K8: cross-attention post-processing of the legacy alignment-head path
// (stable_whisper/timing.py:105-110 + the head mean of :194; median = whisper.timing.median_filter):
//
//   qk[b][a][S..M-2][0..F)  --softmax(x*qk_scale) over frames-->  --z-norm over the R token rows of every frame
//   (biased std)-->  --median filter (width 7, reflect) along frames-->  --mean over the A heads--> matrix[b][R][F]
//
// Three small kernels over an L2-resident working set (A*R*F*4 B, ~6 MB at A=10,R=101,F=1500):
//   1. row softmax           one warp per (b,a,r) row, values kept in registers, written once to W
//   2. column z-norm         one thread per (b,a,column), coalesced across columns, two-pass mean/var, in place on W
//   3. median-7 + head mean  one thread per (b,r,column): 7-tap sorting network per head, accumulate over heads
#include "common.cuh"

namespace stb {

constexpr int QK_MAXV = 47;

__global__ void __launch_bounds__(256) qk_softmax_kernel(const float* __restrict__ qk, long long n_rows, int M,
                                                         long long ldq, int S, int R, int F, float scale,
                                                         float* __restrict__ W, int Fp) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // over B*A*R
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const long long ba = row / R;
    const int r = (int)(row - ba * R);
    const float* src = qk + (ba * M + (S + r)) * ldq;
    float v[QK_MAXV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < QK_MAXV; ++i) {
        const int c = i * 32 + lane;
        v[i] = (c < F) ? src[c] * scale : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < QK_MAXV; ++i) {
        const int c = i * 32 + lane;
        v[i] = (c < F) ? expf(v[i] - mx) : 0.f;
        sum += v[i];
    }
    sum = warp_sum(sum);
    float* dst = W + row * Fp;
#pragma unroll
    for (int i = 0; i < QK_MAXV; ++i) {
        const int c = i * 32 + lane;
        if (c < F) dst[c] = v[i] / sum;
    }
}

__global__ void __launch_bounds__(128) qk_znorm_kernel(float* __restrict__ W, int R, int F, int Fp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    float* base = W + (long long)blockIdx.y * R * Fp + c;       // blockIdx.y over B*A
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += base[(long long)r * Fp];
    const float mean = s / (float)R;
    float q = 0.f;
    for (int r = 0; r < R; ++r) {
        const float d = base[(long long)r * Fp] - mean;
        q += d * d;
    }
    const float sd = sqrtf(q / (float)R);
    for (int r = 0; r < R; ++r) base[(long long)r * Fp] = (base[(long long)r * Fp] - mean) / sd;
}

__device__ __forceinline__ void cswap(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}
// median of 7 via a 16-comparator sorting network (only element 3 is needed)
__device__ __forceinline__ float median7(float a0, float a1, float a2, float a3, float a4, float a5, float a6) {
    cswap(a0, a5); cswap(a0, a3); cswap(a1, a6); cswap(a2, a4); cswap(a0, a1); cswap(a3, a5); cswap(a2, a6);
    cswap(a2, a3); cswap(a3, a6); cswap(a4, a5); cswap(a1, a4); cswap(a1, a3); cswap(a3, a4);
    return a3;
}

__device__ __forceinline__ int reflect(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(128) qk_median_mean_kernel(const float* __restrict__ W, int A, int R, int F, int Fp,
                                                             int width, float* __restrict__ out, long long ldm) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    const int r = blockIdx.y, b = blockIdx.z;
    float acc = 0.f;
    const bool filt = (width == 7) && (F > 3);
    for (int a = 0; a < A; ++a) {
        const float* row = W + (((long long)b * A + a) * R + r) * Fp;
        float v;
        if (filt) {
            v = median7(row[reflect(c - 3, F)], row[reflect(c - 2, F)], row[reflect(c - 1, F)], row[c],
                        row[reflect(c + 1, F)], row[reflect(c + 2, F)], row[reflect(c + 3, F)]);
        } else {
            v = row[c];
        }
        acc += v;
    }
    out[((long long)b * R + r) * ldm + c] = acc / (float)A;
}

}  // namespace stb

extern "C" size_t stb_qkpost_ws_bytes(int B, int A, int R, int F) {
    const int Fp = (F + 3) & ~3;
    return (size_t)B * A * R * Fp * sizeof(float);
}

extern "C" int stb_qk_postprocess(const float* qk, int B, int A, int M, long long ldq, int S, int R, int F, float qk_scale,
                                  int medfilt_width, float* matrix, long long ldm, void* ws, size_t ws_bytes, void* stream) {
    STB_REQUIRE(S >= 0 && S + R <= M, "stb_qk_postprocess: rows S=%d R=%d exceed M=%d", S, R, M);
    STB_REQUIRE(qk && matrix && ws, "stb_qk_postprocess: null pointer");
    STB_REQUIRE(R >= 1 && F >= 1 && F <= 1504 && A >= 1 && B >= 1, "stb_qk_postprocess: bad shape R=%d F=%d A=%d B=%d", R, F, A, B);
    STB_REQUIRE(medfilt_width == 7 || medfilt_width == 1, "stb_qk_postprocess: medfilt_width %d unsupported (7 or 1)", medfilt_width);
    STB_REQUIRE(ws_bytes >= stb_qkpost_ws_bytes(B, A, R, F), "stb_qk_postprocess: workspace too small");
    const int Fp = (F + 3) & ~3;
    float* W = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_rows = (long long)B * A * R;
    stb::ProfScope ps("qk_postprocess(3 kernels)", st, (double)B * A * R * F * 4.0 + (double)B * R * F * 4.0);
    stb::qk_softmax_kernel<<<stb::cdiv(n_rows, 8), 256, 0, st>>>(qk, n_rows, M, ldq, S, R, F, qk_scale, W, Fp);
    STB_LAUNCH_OK();
    stb::qk_znorm_kernel<<<dim3(stb::cdiv(F, 128), B * A), 128, 0, st>>>(W, R, F, Fp);
    STB_LAUNCH_OK();
    stb::qk_median_mean_kernel<<<dim3(stb::cdiv(F, 128), R, B), 128, 0, st>>>(W, A, R, F, Fp, medfilt_width, matrix, ldm);
    STB_LAUNCH_OK();
    return STB_OK;
}
