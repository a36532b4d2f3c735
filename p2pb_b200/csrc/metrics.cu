// metrics.cu -- parity metrics of the hot path: Chamfer nearest-neighbour distances.  Replaces
//   metrics/chamfer3D/chamfer3D.cu:12-134 NmDistanceKernel (launched <<<(32,16),512>>> twice on the legacy stream).
// Same result definition: for every point of xyz1 [B,n,3] the SQUARED distance to its nearest neighbour in
// xyz2 [B,m,3] and that neighbour's index, ties -> lowest index; same fp32 contraction order as the reference
// build (dy*dy first, then fma dx, fma dz).  Candidates are split across `gridDim.z` segments and merged with a
// packed 64-bit atomicMin (distance bits high, index low: lowest index wins ties), queries are tiled over CTAs,
// candidate tiles are staged in shared memory as SoA.
#include "common.cuh"

#define NM_TILE 1024

__global__ void __launch_bounds__(128) nm_distance_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                          int n, int m, int seg_len, unsigned long long* __restrict__ packed)
{
    __shared__ float sx[NM_TILE], sy[NM_TILE], sz[NM_TILE];
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k_begin = blockIdx.z * seg_len;
    const int k_end = min(m, k_begin + seg_len);
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (j < n) {
        const float* p = xyz1 + ((size_t)b * n + j) * 3;
        x1 = p[0]; y1 = p[1]; z1 = p[2];
    }
    float best = INFINITY;
    int bi = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += NM_TILE) {
        const int len = min(NM_TILE, k_end - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const float* q = xyz2 + ((size_t)b * m + k0 + i) * 3;
            sx[i] = q[0]; sy[i] = q[1]; sz[i] = q[2];
        }
        __syncthreads();
        if (j < n) {
#pragma unroll 4
            for (int i = 0; i < len; ++i) {
                const float d = sqdist3(sx[i] - x1, sy[i] - y1, sz[i] - z1);
                if (d < best) { best = d; bi = k0 + i; }
            }
        }
    }
    if (j < n && k_end > k_begin) {
        const unsigned long long v = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)bi;
        atomicMin(packed + (size_t)b * n + j, v);
    }
}

__global__ void nm_unpack_kernel(const unsigned long long* __restrict__ packed, float* __restrict__ dist,
                                 int* __restrict__ idx, long long total)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned long long v = packed[i];
    dist[i] = __uint_as_float((unsigned)(v >> 32));
    idx[i] = (int)(unsigned)v;
}

// xyz1 [B,n,3], xyz2 [B,m,3] -> dist [B,n], idx [B,n]; scratch: B*n u64
P2PB_API int p2pb_nm_distance(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist, int* idx,
                              unsigned long long* scratch, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(B >= 0 && n > 0 && m > 0, "nm_distance: bad sizes");
    if (B == 0) return P2PB_OK;
    P2PB_CUDA_OK(cudaMemsetAsync(scratch, 0xff, sizeof(unsigned long long) * (size_t)B * n, s));
    const int gx = p2pb_cdiv(n, 128);
    int segs = p2pb_cdiv(2 * p2pb_num_sms(), (long long)gx * B);
    const int max_segs = p2pb_cdiv(m, NM_TILE);
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    int seg_len = p2pb_cdiv(m, segs);
    seg_len = p2pb_cdiv(seg_len, NM_TILE) * NM_TILE;
    segs = p2pb_cdiv(m, seg_len);
    nm_distance_kernel<<<dim3(gx, B, segs), 128, 0, s>>>(xyz1, xyz2, n, m, seg_len, scratch);
    P2PB_LAUNCH_OK();
    const long long total = (long long)B * n;
    nm_unpack_kernel<<<p2pb_cdiv(total, 256), 256, 0, s>>>(scratch, dist, idx, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
