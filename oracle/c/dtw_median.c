/* CPU oracle, plain C: DTW and width-w median filter.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restates the CPU semantics of openai-whisper 20250625 `whisper/timing.py` (third-party; call sites in the
 * reference: stable_whisper/timing.py:110,138,195):
 *   dtw_cpu:   cost float32 (N+1)x(M+1), +inf border, cost[0][0]=0; sweep j outer / i inner;
 *              diag iff c0<c1 && c0<c2; else up iff c1<c0 && c1<c2; else left;  cost = x + c  (stored as float32)
 *   backtrace: trace[0][:]=2, trace[:][0]=1; from (N,M) emit (i-1,j-1) until (0,0); reversed.
 *   median_filter: reflect pad w/2 on the last dim, sort each window, take element w/2; rows with
 *              len <= w/2 are returned untouched.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x: [N][M] float32 (the reference casts to float64 first; x + c in float64 then rounded to float32 equals the
 * float32 sum because 53 >= 2*24+2, so float32 arithmetic is bit-identical).
 * out_text/out_time: capacity N+M; returns path length. jumps (optional, [N]): first time index of each row. */
int oracle_dtw(const float *x, int N, int M, int32_t *out_text, int32_t *out_time, int32_t *jumps)
{
    size_t W = (size_t)M + 1;
    float *cost = (float *)malloc(sizeof(float) * (size_t)(N + 1) * W);
    int8_t *trace = (int8_t *)malloc((size_t)(N + 1) * W);
    if (!cost || !trace) { free(cost); free(trace); return -1; }
    for (size_t k = 0; k < (size_t)(N + 1) * W; ++k) { cost[k] = INFINITY; trace[k] = -1; }
    cost[0] = 0.0f;
    for (int j = 1; j <= M; ++j)
        for (int i = 1; i <= N; ++i) {
            float c0 = cost[(size_t)(i - 1) * W + (j - 1)];
            float c1 = cost[(size_t)(i - 1) * W + j];
            float c2 = cost[(size_t)i * W + (j - 1)];
            float c; int8_t t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            cost[(size_t)i * W + j] = x[(size_t)(i - 1) * M + (j - 1)] + c;
            trace[(size_t)i * W + j] = t;
        }
    for (int j = 0; j <= M; ++j) trace[j] = 2;
    for (int i = 0; i <= N; ++i) trace[(size_t)i * W] = 1;
    int i = N, j = M, n = 0;
    while (i > 0 || j > 0) {
        out_text[n] = i - 1; out_time[n] = j - 1; ++n;
        int8_t t = trace[(size_t)i * W + j];
        if (t == 0) { --i; --j; } else if (t == 1) { --i; } else { --j; }
    }
    for (int a = 0, b = n - 1; a < b; ++a, --b) {
        int32_t s = out_text[a]; out_text[a] = out_text[b]; out_text[b] = s;
        s = out_time[a]; out_time[a] = out_time[b]; out_time[b] = s;
    }
    if (jumps) {                              /* stable_whisper/timing.py:197-198 */
        int r = 0;
        for (int k = 0; k < n; ++k)
            if (k == 0 || out_text[k] != out_text[k - 1]) {
                if (r < N) jumps[r] = out_time[k] < 0 ? 0 : out_time[k];
                ++r;
            }
    }
    free(cost); free(trace);
    return n;
}

static int cmp_f32(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* x,out: [rows][len] float32, w odd <= 63 */
void oracle_median_filter(const float *x, float *out, long rows, int len, int w)
{
    int p = w / 2;
    if (len <= p) { memcpy(out, x, sizeof(float) * (size_t)rows * len); return; }
    float win[64];
    for (long r = 0; r < rows; ++r) {
        const float *xr = x + (size_t)r * len;
        float *o = out + (size_t)r * len;
        for (int c = 0; c < len; ++c) {
            for (int k = -p; k <= p; ++k) {
                int q = c + k;
                if (q < 0) q = -q;                       /* reflect (no edge repeat) */
                if (q >= len) q = 2 * (len - 1) - q;
                win[k + p] = xr[q];
            }
            qsort(win, (size_t)w, sizeof(float), cmp_f32);
            o[c] = win[p];
        }
    }
}
