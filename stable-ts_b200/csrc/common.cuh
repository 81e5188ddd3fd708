// Shared device/host helpers for the B200 (sm_100a) kernels of the stable-ts word-timestamp hot path.
// Everything here is hand-written for sm_100a: mbarrier / TMA / tcgen05 inline PTX, warp-shuffle reductions,
// split-fp16 ("hi + lo") helpers used to run fp32-accurate GEMMs on the fp16 tensor pipe.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stablets_b200.h"

// NO kernel parameter of this library is restrict-qualified.  Under programmatic dependent launch a kernel reads its
// predecessor's output after griddepcontrol.wait; with `const T* __restrict__` parameters nvcc turns such reads into
// ld.global.nc, and -- invariant loads carry no memory dependence -- schedules them ABOVE the wait (round 2, seen in SASS:
// LDG.E.CONSTANT of the token id and of the position counter before ACQBULK in embed_step_kernel; under CUDA-graph replay
// the step then embedded the PREVIOUS token).  The qualifier stays in the sources as documentation of intent and is erased
// here; data that really is constant for a kernel's lifetime (weights, biases, tables) is read with explicit __ldg().
// tools/check_pdl_sass.py lists every non-coherent load that precedes a kernel's wait.
#define __restrict__

namespace stb {

// ---- host-side error plumbing (thread-local message returned by stb_last_error) ----
void set_error(const char* fmt, ...);
const char* get_error();

#define STB_CUDA_OK(expr)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            stb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));      \
            (void)cudaGetLastError(); /* do not leave the error latched for an unrelated later launch check */ \
            return STB_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

#define STB_REQUIRE(cond, ...)                                                                         \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            stb::set_error(__VA_ARGS__);                                                               \
            return STB_ERR_ARG;                                                                        \
        }                                                                                              \
    } while (0)

#define STB_LAUNCH_OK()                                                                                \
    do {                                                                                               \
        stb::count_launch();                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess) {                                                                      \
            stb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));  \
            return STB_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

#define STB_TRY(expr)                                                                                  \
    do {                                                                                               \
        int r__ = (expr);                                                                              \
        if (r__ != STB_OK) return r__;                                                                 \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Optional per-launch CUDA-event profiling (stb_prof_enable): events are recorded on the launching stream around one
// kernel; stb_prof_report() aggregates time / algorithmic bytes / algorithmic FLOPs per kernel name.
struct ProfScope {
    ProfScope(const char* name, cudaStream_t st, double bytes = 0.0, double flops = 0.0);
    ~ProfScope();
    int slot;
    cudaStream_t st;
};
void count_launch();   // every kernel launch of this library bumps the counter read by stb_launch_count()
int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with `launch_pdl` may start while its predecessor in the
// stream is still running; it must execute pdl_wait() before touching anything the predecessor writes.  pdl_trigger()
// lets the NEXT kernel start early.  Both are no-ops for ordinary launches.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled(int kind = 0);   // STB_PDL bit mask (misc.cu): kind 0 = ordinary kernels, 1 = decode_linear
// run-time switches (stb_set_option / stb_get_option, misc.cu); defaults come from the environment variable of the same
// meaning so a whole process can be flipped without code (STB_DECODE_SPLITK_LEGACY)
enum Option { OPT_DECODE_SPLITK_LEGACY = 0, OPT_DECODE_LIN_PRIORITY, OPT_XATTN_TC, OPT_DECODE_FUSED_LN, OPT_COUNT };
int option(Option o);
// launch priority of the kernels launched next by this host thread (cudaLaunchAttributePriority; 0 = the stream's own).
// The decode step raises it for its latency-bound linear layers, so that -- when two half-batches are stepped on two streams --
// the block scheduler serves a pending linear before the other half's remaining cross-attention CTAs.
int& launch_priority();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (launch_priority() != 0) {
        attr[n].id = cudaLaunchAttributePriority;
        attr[n].val.priority = launch_priority();
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fp32 -> (hi, lo) fp16 pair with hi + lo == x to ~2^-22 relative.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX (sm_100a)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait (~2 s of SM clock): a protocol bug must trap (launch error) rather than hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("stb: mbarrier timeout tag=%d block=(%d,%d,%d) thread=%d parity=%u\n", tag, blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16/bf16 inputs, fp32 accumulate). One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor for a K-major tile whose rows are exactly 128 bytes (64 x fp16) written by TMA
// with CU_TENSOR_MAP_SWIZZLE_128B: 8-row x 128 B swizzle atoms stacked every 1024 B (stride byte offset), version 1
// (Blackwell), layout type 2 (SWIZZLE_128B).  Bit layout follows the sm_100 descriptor: start>>4 [0,14),
// LBO>>4 [16,30) (unused for swizzled K-major; 1), SBO>>4 [32,46), version [46,48), layout [61,64).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bits[4,6)=1), A=B=f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace stb
