// ops_voxel.cu -- voxelization (scatter-mean) and trilinear devoxelization of the PVConv block.  Replaces
//   third_party/openpoints/cpp/pointnet2_batch/src/vox_gpu.cu:18-78        (grid_stats + avg_voxelize, fp32 atomics)
//   .../trilinear_devox_gpu.cu:21-109                                       (8-corner gather)
//   models/pvcnn.py:215-231 Voxelization.forward                            (~10 ATen launches of coordinate prep)
//
// Design (B200): the scatter is turned into a GATHER.  One CTA per patch sorts the (voxel id, point id) keys in
// shared memory (bitonic, N <= 16384) and writes a CSR (start/cnt per voxel + point order).  The dense grid is
// then produced by a streaming kernel in which every output element is written exactly once (zero for empty
// voxels): no memset, no fp32 atomics, coalesced 128-bit stores, deterministic index-ordered sums (bit-exact to
// oracle/p2pb_oracle.c ora_avg_voxelize_forward).  The CSR depends only on the level's coordinates, so the
// engine builds it once per (level, resolution) per network evaluation and reuses it for every PVConv there.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// voxel_prep: per patch -> ind[N], CSR(order[N], start[r3], cnt[r3]) and (float mode) norm_coords[3][N]
//   mode 0: int voxel coords given (drop-in op avg_voxelize_forward)        ind = x*r2 + y*r + z  (vox_gpu.cu:33)
//   mode 1: raw float coords (fused models/pvcnn.py:215-231): centre, /(2*max||.||+eps), +0.5, *r, clamp, rint
// ---------------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, 1) voxel_prep_kernel(const int* __restrict__ icoords, const float* __restrict__ fcoords,
                                                          int N, int Np2, int r, int normalize, float eps,
                                                          float* __restrict__ norm_coords, int* __restrict__ ind,
                                                          int* __restrict__ order, int* __restrict__ start,
                                                          int* __restrict__ cnt)
{
    P2PB_PDL_SYNC();
    extern __shared__ unsigned s_key[];  // [Np2]
    __shared__ double s_red[3][T / 32];
    __shared__ float s_redf[T / 32];
    __shared__ float s_mean[3], s_denom;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int r2 = r * r, r3 = r2 * r;
    ind += (size_t)b * N;
    order += (size_t)b * N;
    start += (size_t)b * r3;
    cnt += (size_t)b * r3;
    for (int v = t; v < r3; v += T) cnt[v] = 0;

    if (fcoords != nullptr) {
        const float* c = fcoords + (size_t)b * 3 * N;
        float* nc = norm_coords + (size_t)b * 3 * N;
        // mean over N in double (order-independent after rounding to fp32)
        double sx = 0, sy = 0, sz = 0;
        for (int k = t; k < N; k += T) {
            sx += (double)c[k];
            sy += (double)c[k + N];
            sz += (double)c[k + 2 * N];
        }
        sx = warp_sum_d(sx); sy = warp_sum_d(sy); sz = warp_sum_d(sz);
        if (lane == 0) { s_red[0][warp] = sx; s_red[1][warp] = sy; s_red[2][warp] = sz; }
        __syncthreads();
        if (t < 3) {
            double s = 0;
            for (int w = 0; w < T / 32; ++w) s += s_red[t][w];
            s_mean[t] = (float)(s / (double)N);
        }
        __syncthreads();
        const float mx_ = s_mean[0], my_ = s_mean[1], mz_ = s_mean[2];
        float mx = 0.f;
        for (int k = t; k < N; k += T) {
            const float x = __fsub_rn(c[k], mx_), y = __fsub_rn(c[k + N], my_), z = __fsub_rn(c[k + 2 * N], mz_);
            float s = __fmul_rn(x, x);
            s = __fadd_rn(s, __fmul_rn(y, y));
            s = __fadd_rn(s, __fmul_rn(z, z));
            mx = fmaxf(mx, __fsqrt_rn(s));
        }
        mx = warp_max(mx);
        if (lane == 0) s_redf[warp] = mx;
        __syncthreads();
        if (t == 0) {
            float m = 0.f;
            for (int w = 0; w < T / 32; ++w) m = fmaxf(m, s_redf[w]);
            s_denom = __fadd_rn(__fmul_rn(m, 2.0f), eps);
        }
        __syncthreads();
        const float denom = s_denom, rf = (float)r, hi = (float)(r - 1);
        for (int k = t; k < Np2; k += T) {
            unsigned key = 0xffffffffu;
            if (k < N) {
                int vi[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float v = __fsub_rn(c[k + a * N], s_mean[a]);
                    if (normalize) v = __fadd_rn(__fdiv_rn(v, denom), 0.5f);
                    else v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
                    v = __fmul_rn(v, rf);
                    v = fminf(fmaxf(v, 0.0f), hi);
                    nc[k + a * N] = v;
                    vi[a] = __float2int_rn(v);  // round half to even == torch.round (pvcnn.py:228)
                }
                const int id = vi[0] * r2 + vi[1] * r + vi[2];
                ind[k] = id;
                key = (unsigned)id * (unsigned)Np2 + (unsigned)k;
            }
            s_key[k] = key;
        }
    } else {
        const int* c = icoords + (size_t)b * 3 * N;
        for (int k = t; k < Np2; k += T) {
            unsigned key = 0xffffffffu;
            if (k < N) {
                const int id = c[k] * r2 + c[k + N] * r + c[k + 2 * N];
                ind[k] = id;
                key = (unsigned)id * (unsigned)Np2 + (unsigned)k;
            }
            s_key[k] = key;
        }
    }
    __syncthreads();
    // bitonic sort of Np2 keys in shared memory
    for (int k = 2; k <= Np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < Np2; i += T) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a = s_key[i], c2 = s_key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > c2) == up) { s_key[i] = c2; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const unsigned mask = (unsigned)Np2 - 1u;
    int sh = 0;
    while ((1 << sh) < Np2) ++sh;
    for (int i = t; i < N; i += T) {
        const unsigned key = s_key[i];
        order[i] = (int)(key & mask);
        const unsigned v = key >> sh;
        if (i == 0 || (s_key[i - 1] >> sh) != v) start[v] = i;
    }
    __syncthreads();
    for (int i = t; i < N; i += T) {
        const unsigned v = s_key[i] >> sh;
        if (i == N - 1 || (s_key[i + 1] >> sh) != v) cnt[v] = i + 1 - start[v];
    }
}

static int launch_voxel_prep(const int* icoords, const float* fcoords, int B, int N, int r, int normalize, float eps,
                             float* norm_coords, int* ind, int* order, int* start, int* cnt, cudaStream_t s)
{
    P2PB_CHECK_ARG(B >= 0 && N > 0 && r > 0, "voxel_prep: bad sizes");
    if (B == 0) return P2PB_OK;
    int Np2 = 1;
    while (Np2 < N) Np2 <<= 1;
    P2PB_CHECK_ARG(Np2 <= 32768, "voxel_prep: N=%d > 32768 unsupported", N);
    P2PB_CHECK_ARG((double)r * r * r * Np2 < 4294967295.0, "voxel_prep: r^3*N too large for 32-bit sort keys");
    const size_t smem = (size_t)Np2 * sizeof(unsigned);
    if (smem > 48 * 1024)
        P2PB_CUDA_OK(cudaFuncSetAttribute(voxel_prep_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p2pb_prefer_max_smem((const void*)voxel_prep_kernel<1024>);
    (void)p2pb_launch(voxel_prep_kernel<1024>, dim3(B), dim3(1024), (size_t)(smem), s, icoords, fcoords, N, Np2, r, normalize, eps, norm_coords, ind, order,
                                                  start, cnt);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// float coords [B,3,N] -> norm_coords [B,3,N] (un-rounded, clamped), ind [B,N], CSR order [B,N], start/cnt [B,r^3]
P2PB_API int p2pb_voxel_prep(const float* coords, int B, int N, int r, int normalize, float eps, float* norm_coords,
                             int* ind, int* order, int* start, int* cnt, void* stream)
{
    return launch_voxel_prep(nullptr, coords, B, N, r, normalize, eps, norm_coords, ind, order, start, cnt,
                             (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// dense scatter-mean from the CSR, channel-first output [B, C, r3] (reference layout): thread per (c, voxel),
// voxel fastest.  out = sum_{p in voxel, ascending p} feat[c,p] * (1/cnt)     (vox_gpu.cu:70-75)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxelize_cf_kernel(const float* __restrict__ feat, const int* __restrict__ order,
                                                          const int* __restrict__ start, const int* __restrict__ cnt,
                                                          float* __restrict__ out, int C, int N, int r3)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.z;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= r3) return;
    const int n = cnt[(size_t)b * r3 + v];
    float* o = out + (size_t)b * C * r3 + v;
    if (n == 0) {
        for (int c = blockIdx.y; c < C; c += gridDim.y) o[(size_t)c * r3] = 0.f;
        return;
    }
    const int s0 = start[(size_t)b * r3 + v];
    const int* ord = order + (size_t)b * N + s0;
    const float inv = (float)(1.0 / (double)(float)n);
    const float* f = feat + (size_t)b * C * N;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, __fmul_rn(__ldg(f + (size_t)c * N + ord[i]), inv));
        o[(size_t)c * r3] = acc;
    }
}

// drop-in for avg_voxelize_forward (vox.cpp:17-44): feat [B,C,N], int coords [B,3,N] -> out [B,C,r3], ind [B,N], cnt [B,r3]
// scratch_order [B,N] int, scratch_start [B,r3] int
P2PB_API int p2pb_avg_voxelize(const float* feat, const int* coords, int B, int C, int N, int r, float* out, int* ind,
                               int* cnt, int* scratch_order, int* scratch_start, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(C > 0, "avg_voxelize: bad C");
    int rc = launch_voxel_prep(coords, nullptr, B, N, r, 0, 0.f, nullptr, ind, scratch_order, scratch_start, cnt, s);
    if (rc != P2PB_OK || B == 0) return rc;
    const int r3 = r * r * r;
    dim3 grid(p2pb_cdiv(r3, 256), C < 32 ? C : 32, B);
    p2pb_prefer_max_smem((const void*)voxelize_cf_kernel);
    (void)p2pb_launch(voxelize_cf_kernel, dim3(grid), dim3(256), (size_t)(0), s, feat, scratch_order, scratch_start, cnt, out, C, N, r3);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// trilinear devoxelize, channel-first (trilinear_devox_gpu.cu:21-109, inference branch)
// ---------------------------------------------------------------------------------------------------------
struct TriCorners {
    int i[8];
    float w[8];
};

__device__ __forceinline__ TriCorners tri_corners(float x, float y, float z, int r)
{
    const int r2 = r * r;
    const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
    const float xd1 = __fsub_rn(x, xl), yd1 = __fsub_rn(y, yl), zd1 = __fsub_rn(z, zl);
    const float xd0 = __fsub_rn(1.0f, xd1), yd0 = __fsub_rn(1.0f, yd1), zd0 = __fsub_rn(1.0f, zd1);
    TriCorners t;
    t.w[0] = __fmul_rn(__fmul_rn(xd0, yd0), zd0);
    t.w[1] = __fmul_rn(__fmul_rn(xd0, yd0), zd1);
    t.w[2] = __fmul_rn(__fmul_rn(xd0, yd1), zd0);
    t.w[3] = __fmul_rn(__fmul_rn(xd0, yd1), zd1);
    t.w[4] = __fmul_rn(__fmul_rn(xd1, yd0), zd0);
    t.w[5] = __fmul_rn(__fmul_rn(xd1, yd0), zd1);
    t.w[6] = __fmul_rn(__fmul_rn(xd1, yd1), zd0);
    t.w[7] = __fmul_rn(__fmul_rn(xd1, yd1), zd1);
    const int xlo = (int)xl, ylo = (int)yl, zlo = (int)zl;
    const int xh = (xd1 > 0) ? r2 : 0, yh = (yd1 > 0) ? r : 0, zh = (zd1 > 0) ? 1 : 0;
    t.i[0] = xlo * r2 + ylo * r + zlo;
    t.i[1] = t.i[0] + zh;
    t.i[2] = t.i[0] + yh;
    t.i[3] = t.i[2] + zh;
    t.i[4] = t.i[0] + xh;
    t.i[5] = t.i[4] + zh;
    t.i[6] = t.i[4] + yh;
    t.i[7] = t.i[6] + zh;
    return t;
}

__global__ void __launch_bounds__(256) devox_cf_kernel(const float* __restrict__ coords, const float* __restrict__ grid,
                                                       float* __restrict__ out, int C, int N, int r)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* co = coords + (size_t)b * 3 * N;
    const TriCorners t = tri_corners(co[i], co[i + N], co[i + 2 * N], r);
    const int r3 = r * r * r;
    const float* g = grid + (size_t)b * C * r3;
    float* o = out + (size_t)b * C * N;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const float* gc = g + (size_t)c * r3;
        // FMA chain order of the reference build: w001*f001 first, then w000, w010, w011, w100, w101, w110, w111
        float acc = __fmul_rn(t.w[1], __ldg(gc + t.i[1]));
        acc = __fmaf_rn(t.w[0], __ldg(gc + t.i[0]), acc);
#pragma unroll
        for (int q = 2; q < 8; ++q) acc = __fmaf_rn(t.w[q], __ldg(gc + t.i[q]), acc);
        o[(size_t)c * N + i] = acc;
    }
}

// drop-in for trilinear_devoxelize_forward (trilinear_devox.cpp:18-59), is_training=false
P2PB_API int p2pb_trilinear_devoxelize(const float* coords, const float* grid, int B, int C, int N, int r, float* out,
                                       void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && C > 0 && N > 0 && r > 0, "devoxelize: bad sizes");
    if (B == 0) return P2PB_OK;
    dim3 g(p2pb_cdiv(N, 256), C < 32 ? C : 32, B);
    p2pb_prefer_max_smem((const void*)devox_cf_kernel);
    (void)p2pb_launch(devox_cf_kernel, dim3(g), dim3(256), (size_t)(0), (cudaStream_t)stream, coords, grid, out, C, N, r);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
