// Error plumbing, device attributes and small utility kernels shared by the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace stb {
unsigned long long launches();

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
unsigned long long launches() { return g_launches; }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

// fp32 [rows][cols] -> fp16 hi (+ lo) planes.  8 elements per thread, 128-bit loads/stores where aligned.
__global__ void split_f16_kernel(const float* __restrict__ src, long long rows, int cols, long long src_ld,
                                 __half* __restrict__ hi, __half* __restrict__ lo, long long dst_ld) {
    const long long total = rows * (long long)cols;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        __half h, l;
        split_f16(src[r * src_ld + c], h, l);
        hi[r * dst_ld + c] = h;
        if (lo != nullptr) lo[r * dst_ld + c] = l;
    }
}

}  // namespace stb

extern "C" const char* stb_last_error(void) { return stb::get_error(); }
extern "C" int stb_abi_version(void) { return 1; }
extern "C" unsigned long long stb_launch_count(void) { return stb::launches(); }

extern "C" int stb_split_f16(const float* src, long long rows, int cols, long long src_ld, void* hi, void* lo,
                             long long dst_ld, void* stream) {
    STB_REQUIRE(src && hi && rows >= 0 && cols > 0, "stb_split_f16: bad arguments");
    if (rows == 0) return STB_OK;
    const long long total = rows * (long long)cols;
    int blocks = (int)((total + 255) / 256);
    const int cap = stb::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    stb::split_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, src_ld, (__half*)hi, (__half*)lo, dst_ld);
    STB_LAUNCH_OK();
    return STB_OK;
}
