// abi_common.cu -- error reporting and device queries shared by every entry point of libp2pb_b200.so
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void p2pb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

// text of the last error raised on the calling thread ("" if none)
P2PB_API const char* p2pb_last_error() { return g_last_error; }

P2PB_API int p2pb_abi_version() { return 1; }

unsigned long long g_p2pb_launches = 0;
// number of kernels this library has launched (or recorded into a CUDA graph under stream capture) so far
P2PB_API unsigned long long p2pb_launch_count() { return g_p2pb_launches; }

int p2pb_num_sms()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

P2PB_API int p2pb_device_sm_count() { return p2pb_num_sms(); }

// Dynamic shared memory the persistent tensor-core kernels may take per CTA (KiB).  Anything left of the SM's 228 KiB
// is what lets small kernels of the other half-batch chain (DualEngine) become resident next to them.
int g_p2pb_smem_budget_kb = 227;
P2PB_API int p2pb_set_smem_budget_kb(int kb)
{
    if (kb < 128 || kb > 227) {
        p2pb_set_error("smem budget %d KiB out of range [128, 227]", kb);
        return P2PB_ERR_INVALID;
    }
    g_p2pb_smem_budget_kb = kb;
    return P2PB_OK;
}

// programmatic dependent launch on the hot path (see P2PB_PDL_SYNC in common.cuh); 0 = plain stream-ordered launches.
// Measured on B200 inside the CUDA graph: with an early griddepcontrol.launch_dependents the evaluation got 6 % SLOWER (311 vs
// 331 patches/s: the successors' CTAs become resident and wait while the persistent kernels still need the SMs); with the
// implicit trigger at CTA exit (wait only, what P2PB_PDL_SYNC does now) it is 0.5 % faster (338.0 vs 336.3).
int g_p2pb_pdl = 1;
P2PB_API int p2pb_set_pdl(int on)
{
    g_p2pb_pdl = on ? 1 : 0;
    return P2PB_OK;
}

