// ops_points.cu -- point-set ops of the PVCNN hot path: furthest point sampling, gather, ball query, grouping,
// 3-NN inverse-distance interpolation.  Replaces the reference kernels
//   third_party/openpoints/cpp/pointnet2_batch/src/pvcnn_sampling_gpu.cu:17-33,92-184
//   .../pvcnn_ball_query_gpu.cu:19-56, .../pvcnn_grouping_gpu.cu:18-39, .../pvcnn_neighbor_interpolate_gpu.cu:20-124
// which all launch <<<B, <=512>>> (one block per patch => at most B of the 148 SMs busy).  Here every kernel is
// gridded over (patch x tile) so the whole chip is used, coordinates are staged in shared memory, neighbour
// search uses warp ballots / shuffles, and gathers are coalesced.  Index results are bit-exact to the reference
// (same fp32 contraction order, same tie-breaks); see oracle/p2pb_oracle.c.
//
// Feature tensors come in two layouts selected by `cl`:
//   cl = 0  channel-first  [B, C, N]  (the reference's layout; used by the drop-in op API)
//   cl = 1  channels-last  [B, N, ld] (rows; used by the fused engine, ld >= C)
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// Furthest point sampling.  One CTA per patch, the patch's points live in REGISTERS (P per thread), the running
// min-distance too; per iteration: update + local argmax, one shuffle reduction, one barrier (double-buffered
// exchange slots), second shuffle reduction.  The reference needs a strided smem/global scan plus a 9-level
// __syncthreads tree per iteration (pvcnn_sampling_gpu.cu:122-182).
// Tie-break of the reference restated as a key: winner = max d2, then min (k mod 512, k)  (see oracle).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fps_tie_key(int k)
{
    // smaller is better; (k mod 512) major, (k div 512) minor; inverted so that a u64 max picks the minimum
    return 0xffffffffu - ((((unsigned)k & 511u) << 23) | ((unsigned)k >> 9));
}
__device__ __forceinline__ int fps_key_to_index(unsigned key)
{
    unsigned t = 0xffffffffu - key;
    return (int)(((t & 0x7fffffu) << 9) | (t >> 23));
}

template <int P, int T>
__global__ void __launch_bounds__(T, (P <= 8 && T <= 256) ? 1024 / T : 1) fps_reg_kernel(const float* __restrict__ coords, int N, int M,
                                                       int* __restrict__ idx, float* __restrict__ centers)
{
    P2PB_PDL_SYNC();
    extern __shared__ float s_xyz[];  // [3][N]
    __shared__ unsigned long long s_slot[2][32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = T / 32;
    coords += (size_t)b * 3 * N;
    idx += (size_t)b * M;
    float px[P], py[P], pz[P], pd[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = t + i * T;
        if (k < N) {
            px[i] = coords[k];
            py[i] = coords[k + N];
            pz[i] = coords[k + 2 * N];
            s_xyz[k] = px[i];
            s_xyz[k + N] = py[i];
            s_xyz[k + 2 * N] = pz[i];
        } else {
            px[i] = py[i] = pz[i] = 0.f;
        }
        pd[i] = 1e38f;  // pvcnn_sampling.cpp:56
    }
    __syncthreads();
    int old = 0;
    if (t == 0) {
        idx[0] = 0;
        if (centers) {
            centers[((size_t)b * 3 + 0) * M] = s_xyz[0];
            centers[((size_t)b * 3 + 1) * M] = s_xyz[N];
            centers[((size_t)b * 3 + 2) * M] = s_xyz[2 * N];
        }
    }
    for (int j = 1; j < M; ++j) {
        const float x1 = s_xyz[old], y1 = s_xyz[old + N], z1 = s_xyz[old + 2 * N];
        unsigned long long best = 0ull;  // below any real candidate (d2 >= 0 has key bits > 0)
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int k = t + i * T;
            const float d = sqdist3(px[i] - x1, py[i] - y1, pz[i] - z1);
            const float d2 = fminf(d, pd[i]);
            pd[i] = d2;
            const unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | fps_tie_key(k);
            if (k < N && cand > best) best = cand;
        }
        best = warp_max_u64(best);
        if (lane == 0) s_slot[j & 1][warp] = best;
        __syncthreads();
        unsigned long long v = lane < NW ? s_slot[j & 1][lane] : 0ull;
        v = warp_max_u64(v);
        old = fps_key_to_index((unsigned)v);
        if (t == 0) {
            idx[j] = old;
            if (centers) {
                centers[((size_t)b * 3 + 0) * M + j] = s_xyz[old];
                centers[((size_t)b * 3 + 1) * M + j] = s_xyz[old + N];
                centers[((size_t)b * 3 + 2) * M + j] = s_xyz[old + 2 * N];
            }
        }
    }
}

// Large clouds (object-level seed / merge FPS, N up to millions): distances in global memory, coordinates read
// through L1/L2; same selection rule.  One CTA of 1024 threads per cloud.
__global__ void __launch_bounds__(1024, 1) fps_global_kernel(const float* __restrict__ coords, int N, int M,
                                                             float* __restrict__ dist, int* __restrict__ idx)
{
    P2PB_PDL_SYNC();
    __shared__ unsigned long long s_slot[2][32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    coords += (size_t)b * 3 * N;
    dist += (size_t)b * N;
    idx += (size_t)b * M;
    for (int k = t; k < N; k += 1024) dist[k] = 1e38f;
    __syncthreads();
    int old = 0;
    if (t == 0) idx[0] = 0;
    for (int j = 1; j < M; ++j) {
        const float x1 = coords[old], y1 = coords[old + N], z1 = coords[old + 2 * N];
        unsigned long long best = 0ull;
        for (int k = t; k < N; k += 1024) {
            const float d = sqdist3(coords[k] - x1, coords[k + N] - y1, coords[k + 2 * N] - z1);
            const float d2 = fminf(d, dist[k]);
            dist[k] = d2;
            const unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | fps_tie_key(k);
            if (cand > best) best = cand;
        }
        best = warp_max_u64(best);
        if (lane == 0) s_slot[j & 1][warp] = best;
        __syncthreads();
        unsigned long long v = s_slot[j & 1][lane];
        v = warp_max_u64(v);
        old = fps_key_to_index((unsigned)v);
        if (t == 0) idx[j] = old;
    }
}

// Cluster FPS: the points of ONE cloud are split over the CL CTAs of a thread-block cluster (each on its own SM); per
// iteration every CTA finds its local farthest point, publishes {key, x, y, z} into a slot of EVERY CTA of the cluster
// through distributed shared memory, one cluster barrier, and every CTA picks the same winner locally.  Same selection
// key as fps_reg_kernel (bit-identical indices).  Used where one CTA per cloud is ALU-bound: N = 8192 patches (PVDL level 0:
// 2.3 -> ~1 ms for 2048 iterations) and whole clouds of 10^4..2*10^5 points (seed / merge FPS of denoise_object, where
// the one-CTA global-memory kernel spent ~10 us per iteration).  Points and running distances live in shared memory (SoA).
template <int CL>
__global__ void __launch_bounds__(1024, 1) fps_cluster_kernel(const float* __restrict__ coords, int N, int M, int chunk,
                                                              int* __restrict__ idx, float* __restrict__ centers)
{
    P2PB_PDL_SYNC();
    extern __shared__ float s_pts[];                 // [4][chunk]: x, y, z, running min distance
    __shared__ unsigned long long s_warp[32];
    __shared__ int s_warp_loc[32];
    __shared__ __align__(16) unsigned long long s_slot[2][CL][4];   // {key, (x,y), z, pad} per source CTA, double-buffered
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, NW = T >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / CL;
    coords += (size_t)b * 3 * N;
    idx += (size_t)b * M;
    const int k0 = (int)rank * chunk;
    const int n_loc = max(0, min(chunk, N - k0));
    float* sx = s_pts, *sy = s_pts + chunk, *sz = s_pts + 2 * chunk, *sd = s_pts + 3 * chunk;
    for (int i = t; i < n_loc; i += T) {
        sx[i] = coords[k0 + i];
        sy[i] = coords[k0 + i + N];
        sz[i] = coords[k0 + i + 2 * N];
        sd[i] = 1e38f;
    }
    float x1 = coords[0], y1 = coords[N], z1 = coords[2 * N];
    if (rank == 0 && t == 0) {
        idx[0] = 0;
        if (centers) {
            centers[((size_t)b * 3 + 0) * M] = x1;
            centers[((size_t)b * 3 + 1) * M] = y1;
            centers[((size_t)b * 3 + 2) * M] = z1;
        }
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int j = 1; j < M; ++j) {
        unsigned long long best = 0ull;
        int best_loc = 0;
        for (int i = t; i < n_loc; i += T) {
            const float d = sqdist3(sx[i] - x1, sy[i] - y1, sz[i] - z1);
            const float d2 = fminf(d, sd[i]);
            sd[i] = d2;
            const unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | fps_tie_key(k0 + i);
            if (cand > best) {
                best = cand;
                best_loc = i;
            }
        }
        // warp argmax (key is unique per point, so the owner lane is the one whose key equals the maximum)
        const unsigned long long wbest = warp_max_u64(best);
        const unsigned owner = __ballot_sync(0xffffffffu, best == wbest && best != 0ull);
        const int wloc = __shfl_sync(0xffffffffu, best_loc, owner ? (__ffs(owner) - 1) : 0);
        if (lane == 0) {
            s_warp[warp] = wbest;
            s_warp_loc[warp] = wloc;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = lane < NW ? s_warp[lane] : 0ull;
            const unsigned long long cbest = warp_max_u64(v);
            const unsigned own = __ballot_sync(0xffffffffu, v == cbest && lane < NW);
            const int src = own ? (__ffs(own) - 1) : 0;
            const int loc = s_warp_loc[src];
            // lanes 0..CL-1 each publish this CTA's candidate into one CTA of the cluster
            if (lane < CL) {
                float bx = 0.f, by = 0.f, bz = 0.f;
                if (cbest != 0ull) {
                    bx = sx[loc];
                    by = sy[loc];
                    bz = sz[loc];
                }
                const unsigned local = (unsigned)__cvta_generic_to_shared(&s_slot[j & 1][rank][0]);
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
        for (int r = 0; r < CL; ++r) {
            const unsigned long long v = s_slot[j & 1][r][0];
            if (v > win) {
                win = v;
                wr = r;
            }
        }
        const unsigned long long xy = s_slot[j & 1][wr][1];
        x1 = __uint_as_float((unsigned)xy);
        y1 = __uint_as_float((unsigned)(xy >> 32));
        z1 = __uint_as_float((unsigned)s_slot[j & 1][wr][2]);
        if (rank == 0 && t == 0) {
            idx[j] = fps_key_to_index((unsigned)win);
            if (centers) {
                centers[((size_t)b * 3 + 0) * M + j] = x1;
                centers[((size_t)b * 3 + 1) * M + j] = y1;
                centers[((size_t)b * 3 + 2) * M + j] = z1;
            }
        }
    }
    // nobody leaves while a peer may still write into this CTA's slots
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL>
static int launch_fps_cluster(const float* coords, int B, int N, int M, int* idx, float* centers, int threads, cudaStream_t s)
{
    const int chunk = ((N + CL - 1) / CL + 31) & ~31;
    const size_t smem = (size_t)4 * chunk * sizeof(float);
    P2PB_CHECK_ARG(smem <= 200 * 1024, "fps: N=%d too large for a %d-CTA cluster (shared-memory resident points)", N, CL);
    P2PB_CUDA_OK(cudaFuncSetAttribute(fps_cluster_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    if (CL > 8) P2PB_CUDA_OK(cudaFuncSetAttribute(fps_cluster_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    p2pb_prefer_max_smem((const void*)fps_cluster_kernel<CL>);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, fps_cluster_kernel<CL>, &cfg) != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        return P2PB_ERR_UNSUPPORTED;      // this cluster shape cannot be scheduled here: the caller falls back
    }
    P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CL>, coords, N, M, chunk, idx, centers));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// Whole-GPU FPS for clouds beyond the cluster kernel's capacity (room-sized: 2*10^5 .. 2*10^6 points; the patch-centre planner of
// denoise_room draws ~10^3 centres from 2 M points): a cooperative grid of one CTA per SM, every CTA keeps its contiguous chunk of
// points AND running distances in shared memory (16 B per point, <= 13 800 points per CTA), per iteration: local argmax ->
// {key, x, y, z} into this CTA's global slot -> one grid barrier (arrive counter + generation flag) -> every CTA reduces the <= 148
// slots and continues with the winner's coordinates.  ~2.5 us per iteration instead of ~300 us for the one-CTA global-memory scan
// (0.3 s -> 3 ms for the planner of a 2 M-point room).  Same selection key as every other FPS kernel here: bit-identical indices.
struct FpsGridSlot {
    unsigned long long key;
    float x, y, z, pad;
};

__device__ __forceinline__ void fps_grid_barrier(unsigned* count, volatile unsigned* gen, unsigned nblocks, unsigned& local_gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(count, 1u);
        if (prev == nblocks - 1) {
            *count = 0;
            __threadfence();
            atomicAdd(const_cast<unsigned*>(gen), 1u);
        } else {
            while (*gen == local_gen) {
            }
        }
        __threadfence();
    }
    ++local_gen;
    __syncthreads();
}

__global__ void __launch_bounds__(1024, 1) fps_grid_kernel(const float* __restrict__ coords, int N, int M, int chunk,
                                                           FpsGridSlot* __restrict__ slots, unsigned* __restrict__ bar,
                                                           int* __restrict__ idx)
{
    extern __shared__ float s_pts[];                 // [4][chunk]: x, y, z, running min distance
    __shared__ unsigned long long s_warp[32];
    __shared__ int s_warp_loc[32];
    __shared__ FpsGridSlot s_win;
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, NW = T >> 5;
    const int nb = gridDim.x, cta = blockIdx.x;
    const int k0 = cta * chunk;
    const int n_loc = max(0, min(chunk, N - k0));
    float* sx = s_pts, *sy = s_pts + chunk, *sz = s_pts + 2 * chunk, *sd = s_pts + 3 * chunk;
    for (int i = t; i < n_loc; i += T) {
        sx[i] = coords[k0 + i];
        sy[i] = coords[k0 + i + N];
        sz[i] = coords[k0 + i + 2 * N];
        sd[i] = 1e38f;
    }
    float x1 = coords[0], y1 = coords[N], z1 = coords[2 * N];
    if (cta == 0 && t == 0) idx[0] = 0;
    unsigned local_gen = 0;
    __syncthreads();
    for (int j = 1; j < M; ++j) {
        unsigned long long best = 0ull;
        int best_loc = 0;
        for (int i = t; i < n_loc; i += T) {
            const float d = sqdist3(sx[i] - x1, sy[i] - y1, sz[i] - z1);
            const float d2 = fminf(d, sd[i]);
            sd[i] = d2;
            const unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | fps_tie_key(k0 + i);
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
            if (lane == 0) {
                FpsGridSlot o;
                o.key = cbest;
                o.x = cbest != 0ull ? sx[loc] : 0.f;
                o.y = cbest != 0ull ? sy[loc] : 0.f;
                o.z = cbest != 0ull ? sz[loc] : 0.f;
                o.pad = 0.f;
                slots[(size_t)(j & 1) * nb + cta] = o;
            }
        }
        fps_grid_barrier(bar, bar + 1, (unsigned)nb, local_gen);
        // every CTA picks the same winner among the nb candidates (warp 0: up to 5 slots per lane for 148 CTAs)
        if (warp == 0) {
            unsigned long long win = 0ull;
            int wr = 0;
            for (int r = lane; r < nb; r += 32) {
                const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(&slots[(size_t)(j & 1) * nb + r].key);
                if (v > win) {
                    win = v;
                    wr = r;
                }
            }
            const unsigned long long gwin = warp_max_u64(win);
            const unsigned own = __ballot_sync(0xffffffffu, win == gwin);
            const int src = __ffs(own) - 1;
            const int wcta = __shfl_sync(0xffffffffu, wr, src);
            if (lane == 0) {
                const volatile FpsGridSlot* w = &slots[(size_t)(j & 1) * nb + wcta];
                s_win.key = gwin;
                s_win.x = w->x;
                s_win.y = w->y;
                s_win.z = w->z;
                if (cta == 0) idx[j] = fps_key_to_index((unsigned)gwin);
            }
        }
        __syncthreads();
        x1 = s_win.x;
        y1 = s_win.y;
        z1 = s_win.z;
    }
}

// one cloud; scratch: >= 2 * grid * 32 B slots + 8 B barrier (the B*N-float scratch of the ABI is ample)
static int launch_fps_grid(const float* coords, int N, int M, int* idx, float* scratch, cudaStream_t s)
{
    int grid = p2pb_num_sms();
    int chunk = ((N + grid - 1) / grid + 31) & ~31;
    const size_t smem = (size_t)4 * chunk * sizeof(float);
    if (smem > 216 * 1024) return P2PB_ERR_UNSUPPORTED;
    grid = (N + chunk - 1) / chunk;
    P2PB_CUDA_OK(cudaFuncSetAttribute(fps_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(216 * 1024)));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fps_grid_kernel, 1024, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        return P2PB_ERR_UNSUPPORTED;
    }
    FpsGridSlot* slots = reinterpret_cast<FpsGridSlot*>(scratch);
    unsigned* bar = reinterpret_cast<unsigned*>(slots + 2 * (size_t)grid);
    P2PB_CUDA_OK(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), s));
    void* args[] = {(void*)&coords, (void*)&N, (void*)&M, (void*)&chunk, (void*)&slots, (void*)&bar, (void*)&idx};
    // cooperative launch: the runtime guarantees that all CTAs are co-resident (the grid barrier spins)
    P2PB_CUDA_OK(cudaLaunchCooperativeKernel((const void*)fps_grid_kernel, dim3((unsigned)grid), dim3(1024), args, smem, s));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

template <int P, int T>
static int launch_fps_reg(const float* coords, int B, int N, int M, int* idx, float* centers, cudaStream_t s)
{
    const size_t smem = (size_t)3 * N * sizeof(float);
    if (smem > 48 * 1024)
        P2PB_CUDA_OK(cudaFuncSetAttribute(fps_reg_kernel<P, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p2pb_prefer_max_smem((const void*)fps_reg_kernel<P, T>);
    (void)p2pb_launch(fps_reg_kernel<P, T>, dim3(B), dim3(T), (size_t)(smem), s, coords, N, M, idx, centers);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

static int g_fps_cluster = 1;
static int g_fps_shape = 0;     // development aid: alternative (points per thread, threads) shapes of fps_reg_kernel
P2PB_API int p2pb_fps_set_shape(int shape)
{
    g_fps_shape = shape;
    return P2PB_OK;
}
// development aid / cross-check: 0 = never use the cluster kernel, 1 = for whole clouds (default), 2 = also for patches > 2048 points
P2PB_API int p2pb_fps_set_cluster(int on)
{
    g_fps_cluster = on < 0 ? 0 : (on > 3 ? 3 : on);        // 3: development aid -- whole-GPU grid kernel for every cloud of more than 16384 points
    return P2PB_OK;
}

// coords [B,3,N] -> idx int32 [B,M] (+ optional centers [B,3,M] = coords gathered at idx, fused)
// scratch: only needed when N > 16384 (B*N floats)
P2PB_API int p2pb_furthest_point_sampling(const float* coords, int B, int N, int M, int* idx, float* centers,
                                          float* scratch, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(B >= 0 && N > 0 && M >= 0, "fps: bad sizes B=%d N=%d M=%d", B, N, M);
    if (B == 0 || M == 0) return P2PB_OK;
    if (g_fps_shape == 1) {            // tuning: fewer, fatter threads (the per-warp reduction is as costly as 8 points per thread)
        if (N <= 2048 && N > 1024) return launch_fps_reg<16, 128>(coords, B, N, M, idx, centers, s);
        if (N <= 8192 && N > 4096) return launch_fps_reg<16, 512>(coords, B, N, M, idx, centers, s);
    } else if (g_fps_shape == 2) {
        if (N <= 2048 && N > 1024) return launch_fps_reg<4, 512>(coords, B, N, M, idx, centers, s);
        if (N <= 8192 && N > 4096) return launch_fps_reg<32, 256>(coords, B, N, M, idx, centers, s);
    }
    if (N <= 256) return launch_fps_reg<1, 256>(coords, B, N, M, idx, centers, s);
    if (N <= 512) return launch_fps_reg<2, 256>(coords, B, N, M, idx, centers, s);
    if (N <= 1024) return launch_fps_reg<4, 256>(coords, B, N, M, idx, centers, s);
    if (N <= 2048) return launch_fps_reg<8, 256>(coords, B, N, M, idx, centers, s);
    if (g_fps_cluster == 3 && N > 16384 && centers == nullptr && scratch != nullptr) {
        int rc = P2PB_OK;
        for (int b = 0; b < B && rc == P2PB_OK; ++b)
            rc = launch_fps_grid(coords + (size_t)b * 3 * N, N, M, idx + (size_t)b * M, scratch + (size_t)b * N, s);
        if (rc != P2PB_ERR_UNSUPPORTED) return rc;
    }
    if (g_fps_cluster) {
        // Whole clouds (N > 16384): 16 CTAs x 1024 threads.  For patches (N <= 16384) one cluster barrier per iteration costs
        // as much (~1.1 us) as the whole register-resident iteration of fps_reg_kernel, so those keep one CTA per patch
        // (measured on B200: N = 8192 1.15 vs 1.13 us/iteration; N = 149504 3.5 vs 23.9 us/iteration).
        int rc = P2PB_ERR_UNSUPPORTED;
        if (g_fps_cluster == 2 && N > 2048 && N <= 16384) rc = launch_fps_cluster<8>(coords, B, N, M, idx, centers, N <= 8192 ? 256 : 512, s);
        else if (N > 16384 && N <= 16 * 12288) rc = launch_fps_cluster<16>(coords, B, N, M, idx, centers, 1024, s);
        if (rc == P2PB_ERR_UNSUPPORTED && N > 16384 && N <= 8 * 12288) rc = launch_fps_cluster<8>(coords, B, N, M, idx, centers, 1024, s);
        if (rc != P2PB_ERR_UNSUPPORTED) return rc;
    }
    if (N <= 4096) return launch_fps_reg<8, 512>(coords, B, N, M, idx, centers, s);
    // 8192 points: 32 per thread x 256 threads -- the two-stage u64 argmax costs every warp ~80 instructions per iteration, as
    // much as 8 points, so fewer, fatter warps win (measured 0.77 vs 1.14 us/iteration for <8, 1024>)
    if (N <= 8192) return launch_fps_reg<32, 256>(coords, B, N, M, idx, centers, s);
    if (N <= 16384) return launch_fps_reg<16, 1024>(coords, B, N, M, idx, centers, s);
    P2PB_CHECK_ARG(scratch != nullptr, "fps: N=%d > 16384 needs a B*N float scratch buffer", N);
    P2PB_CHECK_ARG(centers == nullptr, "fps: fused centre gather only for N <= 16384");
    if (g_fps_cluster && N > 16 * 12288) {       // room-sized clouds: the whole GPU on one cloud (cooperative grid kernel)
        int rc = P2PB_OK;
        for (int b = 0; b < B && rc == P2PB_OK; ++b)
            rc = launch_fps_grid(coords + (size_t)b * 3 * N, N, M, idx + (size_t)b * M, scratch + (size_t)b * N, s);
        if (rc != P2PB_ERR_UNSUPPORTED) return rc;
    }
    p2pb_prefer_max_smem((const void*)fps_global_kernel);
    (void)p2pb_launch(fps_global_kernel, dim3(B), dim3(1024), (size_t)(0), s, coords, N, M, scratch, idx);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// gather_features: out[b,c,j] = feat[b,c,idx[b,j]]        (pvcnn_sampling_gpu.cu:17-33)
// grouping:        out[b,c,j,k] = feat[b,c,idx[b,j,k]]    (pvcnn_grouping_gpu.cu:18-39)  == gather with M*U indices
// channel-first: thread per (c, j) with j fastest (coalesced index reads and writes).
// ---------------------------------------------------------------------------------------------------------
__global__ void gather_cf_kernel(const float* __restrict__ feat, const int* __restrict__ idx, float* __restrict__ out,
                                 int C, int N, int M)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int src = idx[(size_t)b * M + j];
    const float* f = feat + (size_t)b * C * N;
    float* o = out + (size_t)b * C * M;
    for (int c = blockIdx.y; c < C; c += gridDim.y) o[(size_t)c * M + j] = __ldg(f + (size_t)c * N + src);
}

P2PB_API int p2pb_gather_features(const float* feat, const int* idx, float* out, int B, int C, int N, int M,
                                  void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && C > 0 && N > 0 && M >= 0, "gather: bad sizes");
    if (B == 0 || M == 0) return P2PB_OK;
    dim3 grid(p2pb_cdiv(M, 256), C < 64 ? C : 64, B);
    p2pb_prefer_max_smem((const void*)gather_cf_kernel);
    (void)p2pb_launch(gather_cf_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, idx, out, C, N, M);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

P2PB_API int p2pb_grouping(const float* feat, const int* idx, float* out, int B, int C, int N, int M, int U,
                           void* stream)
{
    return p2pb_gather_features(feat, idx, out, B, C, N, M * U, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Ball query (pvcnn_ball_query_gpu.cu:19-56): first U points (in index order) with d2 < r2, padded with the first
// hit, all-zero row when the ball is empty.  One WARP per centre: 32 points tested per step, ballot + popc
// compaction preserves index order, early exit when U are found.  Points staged in shared memory per CTA.
// ---------------------------------------------------------------------------------------------------------
// Resident CTAs (of 256 threads) per SM that the ALU-bound geometry kernels (ball query, 3-NN search) are launched with.  They run on
// the side stream NEXT TO the main stream's kernels and must leave thread slots for them (a full-occupancy 3-NN search kept a 5 us
// GroupNorm-coefficient kernel of the main stream waiting for 260 us at N = 8192).  Measured at 1 / 2 / 4 CTAs per SM: 397 / 402 / 402
// patches/s on PVDS, 64.7 / 64.4 / 65.6 on PVDL (fewer CTAs make the geometry late for its consumers).
constexpr int SIDE_CTAS_PER_SM = 4;

constexpr int BQ_TILE = 2048;     // points staged per pass (24 KB: a CTA fits next to a persistent GEMM / conv CTA of the main stream)

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ball_query_kernel(const float* __restrict__ centers,
                                                                const float* __restrict__ points, int M, int N,
                                                                float r2, int U, int* __restrict__ out)
{
    P2PB_PDL_SYNC();
    extern __shared__ float s_pts[];  // [3][BQ_TILE]
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    points += (size_t)b * 3 * N;
    centers += (size_t)b * 3 * M;
    // Each warp handles CQ = 4 centres per round: a staged point is read from shared memory once and tested against 4 centres
    // (the scan is bound by the 3 LDS + loop overhead per 32 points, not by the 6 flops of a distance).  The points are staged in
    // tiles of BQ_TILE in index order; a CTA stops staging as soon as all its centres have their U neighbours.
    constexpr int CQ = 4;
    for (int base = blockIdx.x * WARPS * CQ; base < M; base += gridDim.x * WARPS * CQ) {      // uniform over the CTA
        const int j0 = base + warp * CQ;
        float cx[CQ], cy[CQ], cz[CQ];
        int cnt[CQ], first[CQ];
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            const int j = min(j0 + q, M - 1);
            cx[q] = centers[j];
            cy[q] = centers[j + M];
            cz[q] = centers[j + 2 * M];
            cnt[q] = j0 + q < M ? 0 : U;      // centres past the end count as finished
            first[q] = 0;
        }
        for (int t0 = 0; t0 < N; t0 += BQ_TILE) {
            const bool done = cnt[0] >= U && cnt[1] >= U && cnt[2] >= U && cnt[3] >= U;
            if (__syncthreads_and(done)) break;        // (also: every warp is past the previous tile)
            const int tn = min(BQ_TILE, N - t0);
            for (int i = threadIdx.x; i < tn; i += WARPS * 32) {
                s_pts[i] = points[t0 + i];
                s_pts[i + BQ_TILE] = points[N + t0 + i];
                s_pts[i + 2 * BQ_TILE] = points[2 * N + t0 + i];
            }
            __syncthreads();
            if (done) continue;
            for (int k0 = 0; k0 < tn; k0 += 32) {
                if (cnt[0] >= U && cnt[1] >= U && cnt[2] >= U && cnt[3] >= U) break;
                const int k = k0 + lane;
                float px = 0.f, py = 0.f, pz = 0.f;
                if (k < tn) {
                    px = s_pts[k];
                    py = s_pts[k + BQ_TILE];
                    pz = s_pts[k + 2 * BQ_TILE];
                }
#pragma unroll
                for (int q = 0; q < CQ; ++q) {
                    const bool hit = k < tn && cnt[q] < U && sqdist3(cx[q] - px, cy[q] - py, cz[q] - pz) < r2;
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        if (cnt[q] == 0) first[q] = t0 + k0 + __ffs(m) - 1;
                        const int slot = cnt[q] + __popc(m & ((1u << lane) - 1u));
                        if (hit && slot < U) out[((size_t)b * M + j0 + q) * U + slot] = t0 + k;
                        cnt[q] += __popc(m);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            if (j0 + q >= M) continue;
            int* o = out + ((size_t)b * M + j0 + q) * U;
            const int c = min(cnt[q], U);
            // pad with the first hit (or zeros when empty: reference output is zero-initialised)
            for (int v = c + lane; v < U; v += 32) o[v] = first[q];
        }
        __syncthreads();      // the next round restages tile 0
    }
}

P2PB_API int p2pb_ball_query(const float* centers, const float* points, int B, int M, int N, float radius, int U,
                             int* out, void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && M >= 0 && N > 0 && U > 0, "ball_query: bad sizes");
    if (B == 0 || M == 0) return P2PB_OK;
    const float r2 = radius * radius;  // pvcnn_ball_query.cpp:25 (fp32 product)
    const size_t smem = (size_t)3 * BQ_TILE * sizeof(float);
    constexpr int WARPS = 8;
    int gx = p2pb_cdiv(M, WARPS * 4);
    // enough CTAs per patch to fill the chip, but only about SIDE_CTAS_PER_SM per SM: the kernel runs on the geometry stream next
    // to the main stream's kernels
    const int want = p2pb_cdiv(SIDE_CTAS_PER_SM * p2pb_num_sms(), B);
    if (gx > want) gx = want < 1 ? 1 : want;
    p2pb_prefer_max_smem((const void*)ball_query_kernel<WARPS>);
    (void)p2pb_launch(ball_query_kernel<WARPS>, dim3(dim3(gx, B)), dim3(WARPS * 32), (size_t)(smem), (cudaStream_t)stream, centers, points, M, N, r2, U, out);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// 3-NN search (pvcnn_neighbor_interpolate_gpu.cu:20-81): per point the 3 nearest centres (strict <, first wins),
// weights from SQUARED distances clamped to [1e-10, 1e10].  The reference keeps the running bests as doubles
// initialised to 1e40; fp32 with +inf is equivalent (every candidate is an fp32 value, inf and 1e40 both clamp
// to 1e10, and the double products of two fp32 values round to the fp32 product).  Centres staged in smem.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) three_nn_kernel(const float* __restrict__ points, const float* __restrict__ centers,
                                                       int N, int M, int* __restrict__ idx, float* __restrict__ w)
{
    P2PB_PDL_SYNC();
    extern __shared__ __align__(16) float s_c[];  // [3][Mp], Mp = M rounded up to 4 (pad entries are never compared)
    const int b = blockIdx.y;
    const int Mp = (M + 3) & ~3;
    points += (size_t)b * 3 * N;
    centers += (size_t)b * 3 * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        s_c[i] = centers[i];
        s_c[i + Mp] = centers[i + M];
        s_c[i + 2 * Mp] = centers[i + 2 * M];
    }
    __syncthreads();
    int* ix = idx + (size_t)b * 3 * N;
    float* ww = w + (size_t)b * 3 * N;
    // grid-stride over blocks of 256 points: the launcher caps the grid so that this (ALU-bound, side-stream) kernel never fills
    // the thread slots of an SM -- the main stream's small kernels have to be able to start next to it
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const float ux = points[j], uy = points[j + N], uz = points[j + 2 * N];
        float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY;
        int i0 = 0, i1 = 0, i2 = 0;
        auto visit = [&](float cx, float cy, float cz, int k) {
            const float d = sqdist3(ux - cx, uy - cy, uz - cz);
            if (d < b2) {
                b2 = d; i2 = k;
                if (d < b1) {
                    b2 = b1; i2 = i1; b1 = d; i1 = k;
                    if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = k; }
                }
            }
        };
        const int M4 = M & ~3;
        for (int k = 0; k < M4; k += 4) {       // one 16-byte broadcast load per coordinate and four centres, visited in index order
            const float4 cx = *reinterpret_cast<const float4*>(s_c + k);
            const float4 cy = *reinterpret_cast<const float4*>(s_c + Mp + k);
            const float4 cz = *reinterpret_cast<const float4*>(s_c + 2 * Mp + k);
            visit(cx.x, cy.x, cz.x, k);
            visit(cx.y, cy.y, cz.y, k + 1);
            visit(cx.z, cy.z, cz.z, k + 2);
            visit(cx.w, cy.w, cz.w, k + 3);
        }
        for (int k = M4; k < M; ++k) visit(s_c[k], s_c[k + Mp], s_c[k + 2 * Mp], k);
        b0 = fmaxf(fminf(1e10f, b0), 1e-10f);
        b1 = fmaxf(fminf(1e10f, b1), 1e-10f);
        b2 = fmaxf(fminf(1e10f, b2), 1e-10f);
        const float d0d1 = __fmul_rn(b0, b1), d0d2 = __fmul_rn(b0, b2), d1d2 = __fmul_rn(b1, b2);
        const float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
        ww[j] = __fmul_rn(d1d2, inv); ix[j] = i0;
        ww[j + N] = __fmul_rn(d0d2, inv); ix[j + N] = i1;
        ww[j + 2 * N] = __fmul_rn(d0d1, inv); ix[j + 2 * N] = i2;
    }
}

// out[b,c,j] = f[c,i2]*w2 (+fma) f[c,i1]*w1 (+fma) f[c,i3]*w3   (contraction order of the reference build, see oracle)
__global__ void interp_cf_kernel(const float* __restrict__ cfeat, const int* __restrict__ idx, const float* __restrict__ w,
                                 float* __restrict__ out, int C, int N, int M)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int* ix = idx + (size_t)b * 3 * N;
    const float* ww = w + (size_t)b * 3 * N;
    const int i1 = ix[j], i2 = ix[j + N], i3 = ix[j + 2 * N];
    const float w1 = ww[j], w2 = ww[j + N], w3 = ww[j + 2 * N];
    const float* f = cfeat + (size_t)b * C * M;
    float* o = out + (size_t)b * C * N;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const float* fc = f + (size_t)c * M;
        float acc = __fmul_rn(__ldg(fc + i2), w2);
        acc = __fmaf_rn(__ldg(fc + i1), w1, acc);
        acc = __fmaf_rn(__ldg(fc + i3), w3, acc);
        o[(size_t)c * N + j] = acc;
    }
}

P2PB_API int p2pb_three_nn(const float* points, const float* centers, int B, int N, int M, int* idx, float* w,
                           void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && N > 0 && M > 0, "three_nn: bad sizes");
    if (B == 0) return P2PB_OK;
    const size_t smem = (size_t)3 * ((M + 3) & ~3) * sizeof(float);
    P2PB_CHECK_ARG(smem <= 200 * 1024, "three_nn: M=%d too large for shared-memory staging", M);
    if (smem > 48 * 1024)
        P2PB_CUDA_OK(cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p2pb_prefer_max_smem((const void*)three_nn_kernel);
    // about SIDE_CTAS_PER_SM CTAs per SM, in whole passes over the point blocks
    const int nblk = p2pb_cdiv(N, 256);
    int cap = (SIDE_CTAS_PER_SM * p2pb_num_sms()) / B;
    if (cap < 1) cap = 1;
    const int gx = p2pb_cdiv(nblk, p2pb_cdiv(nblk, cap));
    (void)p2pb_launch(three_nn_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(smem), (cudaStream_t)stream, points, centers, N, M, idx, w);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

P2PB_API int p2pb_three_nn_interpolate(const float* points, const float* centers, const float* cfeat, int B, int C,
                                       int N, int M, float* out, int* idx, float* w, void* stream)
{
    int rc = p2pb_three_nn(points, centers, B, N, M, idx, w, stream);
    if (rc != P2PB_OK || B == 0) return rc;
    P2PB_CHECK_ARG(C > 0, "three_nn_interpolate: bad C");
    dim3 grid(p2pb_cdiv(N, 256), C < 64 ? C : 64, B);
    p2pb_prefer_max_smem((const void*)interp_cf_kernel);
    (void)p2pb_launch(interp_cf_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, cfeat, idx, w, out, C, N, M);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
