// room.cu -- device-side patch creation and reassembly for the room sweep.
// Replaces the host loop of denoise_room.py: create_patches (:352-421: numpy fancy indexing, np.random padding, one
// fpsample call per over-full patch), the per-patch normalisation of denoise_patch_batch (:141-146) and the numba running
// mean update_prediction_noisy_batches (:262-289).  Everything works on the radius-query CSR (metrics.cu) of the room:
//   room [N,3] fp32 row-major, off int64 [P+1], csr int32 [off[P]] (ascending room indices per patch).
//
//   room_pad_patches    under-full patches (n < M): rows 0..n-1 = the patch, rows n..M-1 = random duplicates + N(0, s^2)
//                       jitter, s = 1e-2 * |bbox diagonal|; cut = n.  One launch for all of them; counter-based RNG
//                       (seed, patch, slot) so the result does not depend on launch shape, rank count or patch order --
//                       or pre-drawn host randoms (the reference's np.random sequence, --strict_ref).
//   room_fps_patches    over-full patches (n >= M): exact furthest point sampling of M points from a given start index, one
//                       8-CTA thread-block cluster per (patch, replica) job, points + running distances in distributed
//                       shared memory (the scheme of fps_cluster_kernel, ops_points.cu), ragged n, gather through the CSR.
//   patch_normalize     centre (mean) and max-norm scale per patch in fp64 -> x_start [P,3,M] fp32 (+ centre, scale as f64)
//   room_accumulate     de-normalise in fp64 and add into per-point FIXED-POINT sums (int64, 2^-40 m) + counts with integer
//                       atomics: integer addition is associative, so the result is bit-identical for any patch order, any
//                       batch split and any number of ranks (the ranks all_reduce the int64 sums); the reference's
//                       sequential running mean equals sum / count up to f64 rounding.
#include "common.cuh"

namespace {

// ---- counter-based RNG: splitmix64 finaliser over (seed, patch, slot, draw) -- restated bit-for-bit in oracle/ops.py ----
__host__ __device__ __forceinline__ unsigned long long rm_mix(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ unsigned long long rm_draw(unsigned long long seed, unsigned patch, unsigned slot, unsigned draw)
{
    return rm_mix(rm_mix(seed ^ ((unsigned long long)patch << 32 | slot)) + draw);
}
__device__ __forceinline__ float rm_uniform(unsigned long long h) { return ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

__global__ void __launch_bounds__(256) room_pad_kernel(const float* __restrict__ room, const long long* __restrict__ off,
                                                       const int* __restrict__ csr, const int* __restrict__ job_patch,
                                                       const int* __restrict__ job_key, int M,
                                                       unsigned long long seed, const long long* __restrict__ pre_off,
                                                       const int* __restrict__ pre_idx, const float* __restrict__ pre_noise,
                                                       float* __restrict__ xyz_out, int* __restrict__ idx_out, int* __restrict__ cut_out)
{
    __shared__ float s_mn[3][8], s_mx[3][8];
    __shared__ float s_sigma;
    const int j = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int p = job_patch[j];
    const unsigned key = (unsigned)(job_key != nullptr ? job_key[j] : p);     // RNG key: the GLOBAL patch number
    const long long o = off[p];
    const int n = (int)(off[p + 1] - o);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = t; i < n; i += 256) {
        const float* q = room + (size_t)csr[o + i] * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], q[a]);
            mx[a] = fmaxf(mx[a], q[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int m = 16; m > 0; m >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], m));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], m));
        }
        if (lane == 0) {
            s_mn[a][warp] = mn[a];
            s_mx[a][warp] = mx[a];
        }
    }
    __syncthreads();
    if (t == 0) {
        double d2 = 0.0;
        for (int a = 0; a < 3; ++a) {
            float lo = s_mn[a][0], hi = s_mx[a][0];
            for (int w = 1; w < 8; ++w) {
                lo = fminf(lo, s_mn[a][w]);
                hi = fmaxf(hi, s_mx[a][w]);
            }
            const double d = (double)hi - (double)lo;
            d2 += d * d;
        }
        s_sigma = (float)(sqrt(d2) * 1e-2);       // denoise_room.py:377
        cut_out[j] = n;
    }
    __syncthreads();
    const float sigma = s_sigma;
    for (int s = t; s < M; s += 256) {
        int src_local;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (s < n) {
            src_local = s;
        } else if (pre_idx != nullptr) {
            const long long q = pre_off[j] + (s - n);
            src_local = pre_idx[q];
            nx = pre_noise[q * 3];
            ny = pre_noise[q * 3 + 1];
            nz = pre_noise[q * 3 + 2];
        } else {
            src_local = (int)(rm_draw(seed, key, (unsigned)s, 0) % (unsigned long long)n);
            // Box-Muller: two uniforms -> two normals, a third from a second pair
            const float u1 = rm_uniform(rm_draw(seed, key, (unsigned)s, 1)), u2 = rm_uniform(rm_draw(seed, key, (unsigned)s, 2));
            const float u3 = rm_uniform(rm_draw(seed, key, (unsigned)s, 3)), u4 = rm_uniform(rm_draw(seed, key, (unsigned)s, 4));
            const float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u3));
            nx = sigma * r1 * cospif(2.0f * u2);
            ny = sigma * r1 * sinpif(2.0f * u2);
            nz = sigma * r2 * cospif(2.0f * u4);
        }
        const int src = csr[o + src_local];
        const float* q = room + (size_t)src * 3;
        float* w = xyz_out + ((size_t)j * M + s) * 3;
        w[0] = q[0] + nx;
        w[1] = q[1] + ny;
        w[2] = q[2] + nz;
        idx_out[(size_t)j * M + s] = src;
    }
}

__device__ __forceinline__ unsigned rm_tie_key(int k)      // same order as ops_points.cu fps_tie_key: (k mod 512, k), smaller wins
{
    return 0xffffffffu - ((((unsigned)k & 511u) << 23) | ((unsigned)k >> 9));
}
__device__ __forceinline__ int rm_key_to_index(unsigned key)
{
    const unsigned t = 0xffffffffu - key;
    return (int)(((t & 0x7fffffu) << 9) | (t >> 23));
}

constexpr int RF_CL = 8;        // CTAs per (patch, replica) job
constexpr int RF_T = 512;

__global__ void __launch_bounds__(RF_T, 1) room_fps_kernel(const float* __restrict__ room, const long long* __restrict__ off,
                                                           const int* __restrict__ csr, const int* __restrict__ job_patch,
                                                           const int* __restrict__ job_start, int M, int chunk_max,
                                                           float* __restrict__ xyz_out, int* __restrict__ idx_out)
{
    extern __shared__ float s_pts[];                 // [4][chunk_max]: x, y, z, running min distance
    __shared__ unsigned long long s_warp[32];
    __shared__ int s_warp_loc[32];
    __shared__ __align__(16) unsigned long long s_slot[2][RF_CL][4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = RF_T / 32;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int j = blockIdx.x / RF_CL;
    const int p = job_patch[j];
    const long long o = off[p];
    const int n = (int)(off[p + 1] - o);
    const int chunk = ((n + RF_CL - 1) / RF_CL + 31) & ~31;       // <= chunk_max by construction of the launch
    const int k0 = (int)rank * chunk;
    const int n_loc = max(0, min(chunk, n - k0));
    float* sx = s_pts, *sy = s_pts + chunk_max, *sz = s_pts + 2 * chunk_max, *sd = s_pts + 3 * chunk_max;
    for (int i = t; i < n_loc; i += RF_T) {
        const float* q = room + (size_t)csr[o + k0 + i] * 3;
        sx[i] = q[0];
        sy[i] = q[1];
        sz[i] = q[2];
        sd[i] = 1e38f;
    }
    const int start = job_start[j];
    const int src0 = csr[o + start];
    float x1 = room[(size_t)src0 * 3], y1 = room[(size_t)src0 * 3 + 1], z1 = room[(size_t)src0 * 3 + 2];
    float* xo = xyz_out + (size_t)j * M * 3;
    int* io = idx_out + (size_t)j * M;
    if (rank == 0 && t == 0) {
        io[0] = src0;
        xo[0] = x1;
        xo[1] = y1;
        xo[2] = z1;
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int it = 1; it < M; ++it) {
        unsigned long long best = 0ull;
        int best_loc = 0;
        for (int i = t; i < n_loc; i += RF_T) {
            const float d = sqdist3(sx[i] - x1, sy[i] - y1, sz[i] - z1);
            const float d2 = fminf(d, sd[i]);
            sd[i] = d2;
            const unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | rm_tie_key(k0 + i);
            if (cand > best) {
                best = cand;
                best_loc = i;
            }
        }
        const unsigned long long wbest = warp_max_u64(best);
        const unsigned owner = __ballot_sync(0xffffffffu, best == wbest && best != 0ull);
        const int wloc = __shfl_sync(0xffffffffu, best_loc, owner ? (__ffs(owner) - 1) : 0);
        if (lane == 0) {
            s_warp[warp] = wbest;
            s_warp_loc[warp] = wloc;
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned long long v = lane < NW ? s_warp[lane] : 0ull;
            const unsigned long long cbest = warp_max_u64(v);
            const unsigned own = __ballot_sync(0xffffffffu, v == cbest && lane < NW);
            const int loc = s_warp_loc[own ? (__ffs(own) - 1) : 0];
            if (lane < RF_CL) {
                float bx = 0.f, by = 0.f, bz = 0.f;
                if (cbest != 0ull) {
                    bx = sx[loc];
                    by = sy[loc];
                    bz = sz[loc];
                }
                const unsigned local = (unsigned)__cvta_generic_to_shared(&s_slot[it & 1][rank][0]);
                unsigned remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((unsigned)lane));
                const unsigned long long xy = ((unsigned long long)__float_as_uint(by) << 32) | __float_as_uint(bx);
                asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(remote), "l"(cbest), "l"(xy) : "memory");
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote + 16), "l"((unsigned long long)__float_as_uint(bz)) : "memory");
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        unsigned long long win = 0ull;
        int wr = 0;
#pragma unroll
        for (int r = 0; r < RF_CL; ++r) {
            const unsigned long long v = s_slot[it & 1][r][0];
            if (v > win) {
                win = v;
                wr = r;
            }
        }
        const unsigned long long xy = s_slot[it & 1][wr][1];
        x1 = __uint_as_float((unsigned)xy);
        y1 = __uint_as_float((unsigned)(xy >> 32));
        z1 = __uint_as_float((unsigned)s_slot[it & 1][wr][2]);
        if (rank == 0 && t == 0) {
            io[it] = csr[o + rm_key_to_index((unsigned)win)];
            xo[it * 3] = x1;
            xo[it * 3 + 1] = y1;
            xo[it * 3 + 2] = z1;
        }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// one CTA per patch: mean and max-norm in fp64 (denoise_room.py:141-146 works on float64 numpy arrays)
__global__ void __launch_bounds__(256) patch_normalize_kernel(const float* __restrict__ xyz, int M, float* __restrict__ x_start,
                                                              double* __restrict__ center, double* __restrict__ scale)
{
    __shared__ double s_red[3][8];
    __shared__ double s_c[3];
    __shared__ double s_scale;
    const int j = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* q = xyz + (size_t)j * M * 3;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = t; i < M; i += 256) {
        acc[0] += (double)q[i * 3];
        acc[1] += (double)q[i * 3 + 1];
        acc[2] += (double)q[i * 3 + 2];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        acc[a] = warp_sum_d(acc[a]);
        if (lane == 0) s_red[a][warp] = acc[a];
    }
    __syncthreads();
    if (t < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_red[t][w];       // fixed order
        s_c[t] = s / (double)M;
    }
    __syncthreads();
    const double cx = s_c[0], cy = s_c[1], cz = s_c[2];
    double mx = 0.0;
    for (int i = t; i < M; i += 256) {
        const double dx = (double)q[i * 3] - cx, dy = (double)q[i * 3 + 1] - cy, dz = (double)q[i * 3 + 2] - cz;
        mx = fmax(mx, dx * dx + dy * dy + dz * dz);
    }
    for (int m = 16; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    __syncthreads();
    if (lane == 0) s_red[0][warp] = mx;
    __syncthreads();
    if (t == 0) {
        double m2 = 0.0;
        for (int w = 0; w < 8; ++w) m2 = fmax(m2, s_red[0][w]);
        s_scale = sqrt(m2);
        scale[j] = s_scale;
        center[j * 3] = cx;
        center[j * 3 + 1] = cy;
        center[j * 3 + 2] = cz;
    }
    __syncthreads();
    const double inv = 1.0 / s_scale;          // the reference divides; x / s and x * (1/s) agree to 1 ulp of fp64, far below fp32
    float* xs = x_start + (size_t)j * 3 * M;
    for (int i = t; i < M; i += 256) {
        xs[i] = (float)(((double)q[i * 3] - cx) * inv);
        xs[i + M] = (float)(((double)q[i * 3 + 1] - cy) * inv);
        xs[i + 2 * M] = (float)(((double)q[i * 3 + 2] - cz) * inv);
    }
}

constexpr double RM_FIXED = 1099511627776.0;      // 2^40

__global__ void __launch_bounds__(256) room_accumulate_kernel(const float* __restrict__ x_pred, const double* __restrict__ center,
                                                              const double* __restrict__ scale, const int* __restrict__ idx,
                                                              const int* __restrict__ cut, int M, long long total,
                                                              unsigned long long* __restrict__ sum_fixed, int* __restrict__ count)
{
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= total) return;
    const int j = (int)(e / M), s = (int)(e - (long long)j * M);
    if (s >= cut[j]) return;                     // padded duplicates never update the room (denoise_room.py:271-274)
    const int i = idx[e];
    const float* xp = x_pred + (size_t)j * 3 * M + s;
    const double sc = scale[j];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double v = (double)xp[(size_t)a * M] * sc + center[j * 3 + a];      // denoise_room.py:176
        atomicAdd(sum_fixed + (size_t)i * 3 + a, (unsigned long long)__double2ll_rn(v * RM_FIXED));
    }
    atomicAdd(count + i, 1);
}

}  // namespace

// job_patch int32 [n_jobs]: CSR patch of every UNDER-full job (n < M); job_key int32 [n_jobs] (optional): the number the RNG is
// keyed by (the patch's GLOBAL number when the CSR only holds one rank's shard; default = job_patch).  pre_off / pre_idx / pre_noise (optional, all or none):
// host-drawn randoms of the reference's np.random sequence, pre_off int64 [n_jobs+1] (M - n draws per job).
// -> xyz_out [n_jobs, M, 3], idx_out int32 [n_jobs, M] (room indices), cut_out int32 [n_jobs]
P2PB_API int p2pb_room_pad_patches(const float* room, const long long* off, const int* csr, const int* job_patch, const int* job_key,
                                   int n_jobs, int M, unsigned long long seed, const long long* pre_off, const int* pre_idx, const float* pre_noise,
                                   float* xyz_out, int* idx_out, int* cut_out, void* stream)
{
    P2PB_CHECK_ARG(n_jobs >= 0 && M > 0, "room_pad_patches: bad sizes");
    P2PB_CHECK_ARG((pre_off == nullptr) == (pre_idx == nullptr) && (pre_idx == nullptr) == (pre_noise == nullptr),
                   "room_pad_patches: pre-drawn randoms need pre_off, pre_idx and pre_noise together");
    if (n_jobs == 0) return P2PB_OK;
    room_pad_kernel<<<n_jobs, 256, 0, (cudaStream_t)stream>>>(room, off, csr, job_patch, job_key, M, seed, pre_off, pre_idx, pre_noise, xyz_out,
                                                              idx_out, cut_out);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// job_patch / job_start int32 [n_jobs]: CSR patch and LOCAL start index (0 <= start < n) of every over-full job (n >= M);
// n_max = the largest n among the jobs.  -> xyz_out [n_jobs, M, 3] (FPS order), idx_out int32 [n_jobs, M]
P2PB_API int p2pb_room_fps_patches(const float* room, const long long* off, const int* csr, const int* job_patch, const int* job_start,
                                   int n_jobs, int n_max, int M, float* xyz_out, int* idx_out, void* stream)
{
    P2PB_CHECK_ARG(n_jobs >= 0 && M > 0 && n_max >= M, "room_fps_patches: bad sizes (n_jobs=%d M=%d n_max=%d)", n_jobs, M, n_max);
    if (n_jobs == 0) return P2PB_OK;
    const int chunk_max = ((n_max + RF_CL - 1) / RF_CL + 31) & ~31;
    const size_t smem = (size_t)4 * chunk_max * sizeof(float);
    P2PB_CHECK_ARG(smem <= 200 * 1024, "room_fps_patches: a radius patch of %d points exceeds the %d-CTA cluster's shared memory (%d max)",
                   n_max, RF_CL, RF_CL * 12800);
    P2PB_CUDA_OK(cudaFuncSetAttribute(room_fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_jobs * RF_CL), 1, 1);
    cfg.blockDim = dim3(RF_T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RF_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, room_fps_kernel, room, off, csr, job_patch, job_start, M, chunk_max, xyz_out, idx_out));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// xyz [P, M, 3] (world coordinates) -> x_start [P, 3, M] fp32 (centred, max-norm scaled), center f64 [P,3], scale f64 [P]
P2PB_API int p2pb_patch_normalize(const float* xyz, int P, int M, float* x_start, double* center, double* scale, void* stream)
{
    P2PB_CHECK_ARG(P >= 0 && M > 0, "patch_normalize: bad sizes");
    if (P == 0) return P2PB_OK;
    patch_normalize_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(xyz, M, x_start, center, scale);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// x_pred [P, 3, M] (network output, normalised) + center/scale/idx/cut of the same P patches -> sum_fixed int64 [N,3] (2^-40
// units, two's complement), count int32 [N]: both ACCUMULATE (caller zeroes them once per room)
P2PB_API int p2pb_room_accumulate(const float* x_pred, const double* center, const double* scale, const int* idx, const int* cut, int P,
                                  int M, long long* sum_fixed, int* count, void* stream)
{
    P2PB_CHECK_ARG(P >= 0 && M > 0, "room_accumulate: bad sizes");
    if (P == 0) return P2PB_OK;
    const long long total = (long long)P * M;
    room_accumulate_kernel<<<p2pb_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x_pred, center, scale, idx, cut, M, total,
                                                                                   reinterpret_cast<unsigned long long*>(sum_fixed), count);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
