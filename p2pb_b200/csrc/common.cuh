// common.cuh -- shared helpers for the p2pb_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define P2PB_API extern "C" __attribute__((visibility("default")))

// error codes returned through the C ABI (include/p2pb_b200.h)
#define P2PB_OK 0
#define P2PB_ERR_INVALID (-1)
#define P2PB_ERR_CUDA (-2)
#define P2PB_ERR_UNSUPPORTED (-3)

// thread-local last-error text, see p2pb_last_error()
void p2pb_set_error(const char* fmt, ...);

#define P2PB_CHECK_ARG(cond, ...)             \
    do {                                      \
        if (!(cond)) {                        \
            p2pb_set_error(__VA_ARGS__);      \
            return P2PB_ERR_INVALID;          \
        }                                     \
    } while (0)

#define P2PB_CUDA_OK(expr)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            p2pb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return P2PB_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// launch check: never exit() (the reference does, cuda_utils.cuh:30-40); report through the return code
extern unsigned long long g_p2pb_launches;  // kernels launched through this library (p2pb_launch_count)
#define P2PB_LAUNCH_OK()                                                                      \
    do {                                                                                      \
        ++g_p2pb_launches;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            p2pb_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return P2PB_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// Small kernels must be able to share an SM with the persistent tensor-core kernels (which hold ~220 KB of dynamic shared
// memory per SM): an SM's shared-memory carve-out can only change while the SM is idle, so a kernel that prefers the
// default (L1-heavy) carve-out waits for the big kernel to drain instead of running next to it.  Host-side and
// idempotent; the hot path replays a CUDA graph, so it costs nothing there.
static inline void p2pb_prefer_max_smem(const void* kernel)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

// Programmatic dependent launch: every hot-path kernel starts with P2PB_PDL_SYNC() -- wait until the grids this launch depends
// on have completed and flushed (a no-op for a plain launch) -- and is launched through p2pb_launch(), which sets programmatic
// stream serialisation.  No kernel triggers early (griddepcontrol.launch_dependents measured slower, see abi_common.cu): the
// implicit trigger at CTA exit lets the successor's launch overlap the predecessor's drain.
#define P2PB_PDL_SYNC()                                               \
    do {                                                              \
        asm volatile("griddepcontrol.wait;" ::: "memory");            \
    } while (0)

extern int g_p2pb_pdl;      // 1: launches carry the programmatic-serialisation attribute (p2pb_set_pdl)

template <typename... KArgs, typename... Args>
static inline cudaError_t p2pb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_p2pb_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int p2pb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

extern int g_p2pb_smem_budget_kb;  // see p2pb_set_smem_budget_kb (abi_common.cu)

// number of SMs of the current device (148 on B200), cached
int p2pb_num_sms();

// squared distance with the reference's nvcc contraction order (see oracle/p2pb_oracle.c sqdist3):
//   t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t)
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz)
{
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    t = __fmaf_rn(dz, dz, t);
    return t;
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m)
{
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = shfl_xor_u64(v, m);
        v = o > v ? o : v;
    }
    return v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// x * sigmoid(x); fast reciprocal (2 ulp) instead of the IEEE division: the activation passes are instruction-bound otherwise
__device__ __forceinline__ float swishf(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
