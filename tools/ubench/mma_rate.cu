// Microbenchmark: tcgen05.mma issue/execute rate on B200 for operand layouts / N / dtype (operands resident in smem).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <int KIND>  // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// amode: 0 SW128 (advance 32B per k-step inside atom), 1 NONE aligned, 2 NONE misaligned by 1 row (16 B), 3 NONE misaligned by 3 rows
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int amode, int bmode, int iters, long long* out_cycles)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t afmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (afmt << 7) | (afmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 64 * 1024);
        const uint32_t W = 199;
        // descriptors precomputed so that the timed loop is (almost) nothing but tcgen05.mma issues
        uint64_t adv[36], bdv[4];
        for (int t9 = 0; t9 < 9; ++t9)
            for (int j = 0; j < 4; ++j) {
                uint64_t ad;
                if (amode == 0) ad = desc_sw128(abase) + (uint64_t)(j * 2);
                else if (amode == 4) ad = desc_sw128(abase + 128u * 3u) + (uint64_t)(j * 2);
                else if (amode == 5) {
                    const uint32_t roff = (uint32_t)(35 + (t9 / 3 - 1) * 34 + (t9 % 3 - 1));
                    ad = desc_sw128(abase + 128u * roff) + (uint64_t)(j * 2);
                } else {
                    const uint32_t roff = amode == 1 ? 0u : (amode == 2 ? 1u : 3u);
                    ad = desc_none(abase + (2 * j * W + roff) * 16, W * 16, 128);
                }
                adv[t9 * 4 + j] = ad;
            }
        for (int j = 0; j < 4; ++j) bdv[j] = bmode == 0 ? desc_sw128(bbase) + (uint64_t)(j * 2) : desc_none(bbase + (2 * j * 256) * 16, 256 * 16, 128);
        long long t0 = clock64();
        for (int it = 0; it < iters / 9; ++it) {
#pragma unroll
            for (int q = 0; q < 36; ++q) mma<KIND>(tmem + (uint32_t)((it & 1) * 256), adv[q], bdv[q & 3], idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        long long t1 = clock64();
        if (blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main()
{
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    const char* an[6] = {"SW128", "NONE-aligned", "NONE+1row", "NONE+3rows", "SW128+3rows", "SW128-9taps"};
    const char* bn[2] = {"SW128", "NONE"};
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {32, 64, 128, 256})
            for (int am : {0, 1, 4, 5})
                for (int bm = 0; bm < 1; ++bm) {
                    for (int rep = 0; rep < 2; ++rep) {
                        if (kind == 0) bench<0><<<148, 128, 170 * 1024>>>(N, am, bm, iters, d);
                        else bench<1><<<148, 128, 170 * 1024>>>(N, am, bm, iters, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    long long c;
                    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    const double per = (double)c / ((iters / 9) * 36);
                    const double k = kind == 0 ? 8 : 16;
                    const double flops_per_cycle = 2.0 * 128 * N * k / per;
                    printf("%s N=%3d A=%-13s B=%-6s : %7.1f cycles/MMA  -> %7.0f flop/cycle/SM = %6.0f TFLOP/s @1.9GHz x148\n",
                           kind == 0 ? "tf32" : "bf16", N, an[am], bn[bm], per, flops_per_cycle, flops_per_cycle * 1.9e9 * 148 / 1e12);
                }
    return 0;
}
