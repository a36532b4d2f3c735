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
    p2pb_prefer_max_smem((const void*)nm_distance_kernel);
    nm_distance_kernel<<<dim3(gx, B, segs), 128, 0, s>>>(xyz1, xyz2, n, m, seg_len, scratch);
    P2PB_LAUNCH_OK();
    const long long total = (long long)B * n;
    p2pb_prefer_max_smem((const void*)nm_unpack_kernel);
    nm_unpack_kernel<<<p2pb_cdiv(total, 256), 256, 0, s>>>(scratch, dist, idx, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// =========================================================================================================
// Approximate earth mover's distance (forward): replaces approxmatch + matchcost
//   (metrics/PyTorchEMD/cuda/emd_kernel.cu:33-165, 211-253; called by emd_nograd.py:19-44).
// Same annealed soft assignment: levels j = 7..-2 (level = -4^j, 0 at j = -2), three all-pairs passes per level
//   (1) ratioL[k] = remainL[k] / (1e-9 + sum_l e^{level d(k,l)} remainR[l])
//   (2) sumr = remainR[l] sum_k e^{level d} ratioL[k];  ratioR[l] = min(remainR/(sumr+1e-9), 1) remainR;  remainR -= sumr (>= 0)
//   (3) w(k,l) = e^{level d} ratioL[k] ratioR[l];  match[l,k] += w;  remainL[k] -= sum_l w (>= 0)
// and cost = sum_{k,l} d(k,l) match[l,k] with the SQUARED distance d (this fork, :236-237).
// B200 design: the reference runs one 512-thread CTA per cloud (<<<32,512>>>: at most 32 SMs) and materialises
// match [B,m,n] (268 MB per cloud at 8192 points) only to contract it with d afterwards.  Here every pass is its own
// launch over (row tiles x B) CTAs -- a warp per row, lanes striding the other cloud -- and pass (3) accumulates
// sum_l d*w per row directly, so match is never stored: memory O(n+m), all SMs busy.  fp32, __expf like the reference;
// the summation order differs (warp tree vs one thread per row), so results agree to fp tolerance, not bitwise.
// =========================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) emd_pass_kernel(const float* __restrict__ rows_xyz, const float* __restrict__ cols_xyz, int R,
                                                       int C, float level, const float* __restrict__ col_w,
                                                       float* __restrict__ remain_row, float* __restrict__ ratio_row,
                                                       float* __restrict__ cost_row)
{
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + warp;
    if (k >= R) return;
    const float* pr = rows_xyz + ((size_t)b * R + k) * 3;
    const float x1 = pr[0], y1 = pr[1], z1 = pr[2];
    const float* pc = cols_xyz + (size_t)b * C * 3;
    const float* w = col_w + (size_t)b * C;
    float acc = 0.f, dacc = 0.f;
    for (int l = lane; l < C; l += 32) {
        const float dx = pc[l * 3] - x1, dy = pc[l * 3 + 1] - y1, dz = pc[l * 3 + 2] - z1;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float e = __expf(level * d) * w[l];
        acc += e;
        if (MODE == 3) dacc = fmaf(d, e, dacc);
    }
    acc = warp_sum(acc);
    if (MODE == 3) dacc = warp_sum(dacc);
    if (lane != 0) return;
    const size_t o = (size_t)b * R + k;
    if (MODE == 1) {
        ratio_row[o] = remain_row[o] / (1e-9f + acc);
    } else if (MODE == 2) {
        const float rem = remain_row[o];
        const float sumr = acc * rem;
        const float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
        ratio_row[o] = consumption * rem;
        remain_row[o] = fmaxf(0.0f, rem - sumr);
    } else {
        const float rl = ratio_row[o];
        cost_row[o] += dacc * rl;
        remain_row[o] = fmaxf(0.0f, remain_row[o] - acc * rl);
    }
}

__global__ void emd_init_kernel(float* remainL, float* remainR, float* cost_row, int n, int m, float multiL, float multiR,
                                long long totalL, long long totalR)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < totalL) {
        remainL[i] = multiL;
        cost_row[i] = 0.f;
    }
    if (i < totalR) remainR[i] = multiR;
}

__global__ void __launch_bounds__(256) emd_cost_kernel(const float* __restrict__ cost_row, int n, float* __restrict__ cost)
{
    __shared__ double s[8];
    const int b = blockIdx.x;
    double a = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) a += (double)cost_row[(size_t)b * n + k];
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s[i];
        cost[b] = (float)t;
    }
}

// xyz1 [B,n,3], xyz2 [B,m,3] -> cost [B] (un-normalised, like matchcost_forward); scratch: B*(3n+2m) floats
P2PB_API int p2pb_emd_approx(const float* xyz1, const float* xyz2, int B, int n, int m, float* cost, float* scratch, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(B >= 0 && n > 0 && m > 0, "emd_approx: bad sizes");
    P2PB_CHECK_ARG(scratch != nullptr, "emd_approx: scratch of B*(3n+2m) floats required");
    if (B == 0) return P2PB_OK;
    float* remainL = scratch;
    float* ratioL = remainL + (size_t)B * n;
    float* cost_row = ratioL + (size_t)B * n;
    float* remainR = cost_row + (size_t)B * n;
    float* ratioR = remainR + (size_t)B * m;
    const float multiL = n >= m ? 1.f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.f;  // integer division, :38-43
    const long long tl = (long long)B * n, tr = (long long)B * m;
    p2pb_prefer_max_smem((const void*)emd_init_kernel);
    emd_init_kernel<<<p2pb_cdiv(tl > tr ? tl : tr, 256), 256, 0, s>>>(remainL, remainR, cost_row, n, m, multiL, multiR, tl, tr);
    P2PB_LAUNCH_OK();
    const dim3 gl(p2pb_cdiv(n, 8), B), gr(p2pb_cdiv(m, 8), B);
    for (int j = 7; j >= -2; --j) {
        const float level = j == -2 ? 0.f : -powf(4.0f, (float)j);
        p2pb_prefer_max_smem((const void*)emd_pass_kernel<1>);
        emd_pass_kernel<1><<<gl, 256, 0, s>>>(xyz1, xyz2, n, m, level, remainR, remainL, ratioL, nullptr);
        P2PB_LAUNCH_OK();
        p2pb_prefer_max_smem((const void*)emd_pass_kernel<2>);
        emd_pass_kernel<2><<<gr, 256, 0, s>>>(xyz2, xyz1, m, n, level, ratioL, remainR, ratioR, nullptr);
        P2PB_LAUNCH_OK();
        p2pb_prefer_max_smem((const void*)emd_pass_kernel<3>);
        emd_pass_kernel<3><<<gl, 256, 0, s>>>(xyz1, xyz2, n, m, level, ratioR, remainL, ratioL, cost_row);
        P2PB_LAUNCH_OK();
    }
    p2pb_prefer_max_smem((const void*)emd_cost_kernel);
    emd_cost_kernel<<<B, 256, 0, s>>>(cost_row, n, cost);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// =========================================================================================================
// kNN patch extraction: for every query the K nearest points of one cloud, ascending distance (ties: lower index).
// Replaces pytorch3d.ops.knn_points as called by denoise_object.py:90-91 (K = 2048 neighbours of each FPS seed in a
// 10k..150k point cloud; return_sorted default True) -- an un-vendored dependency of the reference: the
// contract reproduced is "squared L2, K smallest, sorted ascending".
// One CTA per query: (1) 4-pass radix select on the fp32 distance bits (non-negative floats order like unsigned ints)
// finds the K-th smallest distance T, (2) an index-ordered compaction takes everything below T plus the first
// (K - #below) points at exactly T, (3) a shared-memory bitonic sort of the packed (distance bits, index) keys orders
// them.  Distances are recomputed per pass (3 loads + 3 FMAs) instead of being stored.
// =========================================================================================================
#define KNN_THREADS 1024

__device__ __forceinline__ unsigned knn_dist_bits(const float* __restrict__ pts, int i, float qx, float qy, float qz)
{
    return __float_as_uint(sqdist3(pts[i * 3] - qx, pts[i * 3 + 1] - qy, pts[i * 3 + 2] - qz));
}

__global__ void __launch_bounds__(KNN_THREADS) knn_select_kernel(const float* __restrict__ queries, const float* __restrict__ pts,
                                                                 int N, int K, int Kp2, int* __restrict__ idx_out,
                                                                 float* __restrict__ dist_out)
{
    extern __shared__ unsigned long long s_keys[];   // [Kp2]
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_scan[KNN_THREADS / 32];
    __shared__ unsigned s_prefix, s_remaining, s_nsel, s_neq;
    const int q = blockIdx.x, t = threadIdx.x;
    const float qx = queries[q * 3], qy = queries[q * 3 + 1], qz = queries[q * 3 + 2];
    // ---- (1) radix select: after the 4 passes `prefix` is the bit pattern of the K-th smallest distance
    if (t == 0) {
        s_prefix = 0;
        s_remaining = (unsigned)K;
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (t < 256) s_hist[t] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = t; i < N; i += KNN_THREADS) {
            const unsigned d = knn_dist_bits(pts, i, qx, qy, qz);
            if ((d & mask) == prefix) atomicAdd(&s_hist[(d >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (t == 0) {
            unsigned rem = s_remaining, bin = 0;
            for (; bin < 256; ++bin) {
                const unsigned c = s_hist[bin];
                if (rem <= c) break;
                rem -= c;
            }
            s_prefix = prefix | (bin << shift);
            s_remaining = rem;          // rank of the K-th smallest inside the selected bin (1-based)
        }
        __syncthreads();
    }
    const unsigned T = s_prefix;
    const unsigned take_eq = s_remaining;    // how many points at exactly T belong to the K nearest
    // ---- (2) compaction in index order (block-wide scans per chunk of KNN_THREADS points)
    if (t == 0) {
        s_nsel = 0;
        s_neq = 0;
    }
    for (int i = t; i < Kp2; i += KNN_THREADS) s_keys[i] = ~0ull;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += KNN_THREADS) {
        const int i = i0 + t;
        unsigned d = 0xffffffffu;
        if (i < N) d = knn_dist_bits(pts, i, qx, qy, qz);
        const bool below = i < N && d < T, eq = i < N && d == T;
        // packed scan: low 16 bits count `below`, high 16 bits count `eq`
        unsigned v = (below ? 1u : 0u) | (eq ? 0x10000u : 0u);
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
            if ((t & 31) >= o) inc += n;
        }
        if ((t & 31) == 31) s_scan[t >> 5] = inc;
        __syncthreads();
        if (t < 32) {
            unsigned w = s_scan[t];
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, winc, o);
                if (t >= o) winc += n;
            }
            s_scan[t] = winc - w;    // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned excl = s_scan[t >> 5] + inc - v;
        const unsigned base_below = s_nsel, base_eq = s_neq;
        // points below T fill [0, K - take_eq) in index order; the first take_eq points at exactly T fill the rest from the end
        if (below) {
            s_keys[base_below + (excl & 0xffffu)] = ((unsigned long long)d << 32) | (unsigned)i;
        } else if (eq) {
            const unsigned rank_eq = base_eq + (excl >> 16);
            if (rank_eq < take_eq) s_keys[(unsigned)K - 1u - rank_eq] = ((unsigned long long)d << 32) | (unsigned)i;
        }
        __syncthreads();
        if (t == KNN_THREADS - 1) {
            s_nsel = base_below + ((excl + v) & 0xffffu);
            s_neq = base_eq + ((excl + v) >> 16);
        }
        __syncthreads();
    }
    // ---- (3) bitonic sort of the Kp2 keys (padding keys are all-ones and sink to the end)
    for (int size = 2; size <= Kp2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < Kp2 / 2; i += KNN_THREADS) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == up) {
                    s_keys[lo] = b;
                    s_keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = t; i < K; i += KNN_THREADS) {
        const unsigned long long key = s_keys[i];
        idx_out[(size_t)q * K + i] = (int)(unsigned)key;
        if (dist_out != nullptr) dist_out[(size_t)q * K + i] = __uint_as_float((unsigned)(key >> 32));
    }
}

// queries [Q,3], pts [N,3] -> idx int32 [Q,K] (ascending distance, ties by index), dist [Q,K] squared (optional)
P2PB_API int p2pb_knn_points(const float* queries, const float* pts, int Q, int N, int K, int* idx, float* dist, void* stream)
{
    P2PB_CHECK_ARG(Q >= 0 && N > 0 && K > 0 && K <= N, "knn_points: need 0 < K <= N (Q=%d N=%d K=%d)", Q, N, K);
    P2PB_CHECK_ARG(K <= 16384, "knn_points: K=%d exceeds the shared-memory sort (16384)", K);
    P2PB_CHECK_ARG(N < (1 << 30), "knn_points: N too large");
    if (Q == 0) return P2PB_OK;
    int Kp2 = 2;
    while (Kp2 < K) Kp2 <<= 1;
    const size_t smem = (size_t)Kp2 * sizeof(unsigned long long);
    static bool attr_set = false;
    if (!attr_set) {
        P2PB_CUDA_OK(cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        attr_set = true;
    }
    p2pb_prefer_max_smem((const void*)knn_select_kernel);
    knn_select_kernel<<<Q, KNN_THREADS, smem, (cudaStream_t)stream>>>(queries, pts, N, K, Kp2, idx, dist);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// =========================================================================================================
// Radius query for room patch creation: all points of the room within `radius` of each patch centre.
// Replaces sklearn.neighbors.KDTree(room_points).query_radius(centers, r) (denoise_room.py:454-465), a CPU KD-tree on
// up to millions of points that left the GPUs idle between chunks.  Result: CSR (offsets, indices)
// with the indices of every centre in ASCENDING point order (sklearn's order is traversal order, i.e. unspecified; the
// reference only indexes with it).  Membership: squared distance <= radius^2 in fp32 with the library's sqdist3 order.
// Two passes of the same brute-force scan (the cloud is L2-resident: 2 M points = 24 MB): count, then -- after an
// exclusive scan of the counts -- an index-ordered block compaction.  One CTA per centre.
// =========================================================================================================
__global__ void __launch_bounds__(1024) radius_count_kernel(const float* __restrict__ centers, const float* __restrict__ pts, int N,
                                                            float r2, int* __restrict__ counts)
{
    __shared__ int s_cnt[32];
    const int c = blockIdx.x, t = threadIdx.x;
    const float cx = centers[c * 3], cy = centers[c * 3 + 1], cz = centers[c * 3 + 2];
    int n = 0;
    for (int i = t; i < N; i += 1024) n += sqdist3(pts[i * 3] - cx, pts[i * 3 + 1] - cy, pts[i * 3 + 2] - cz) <= r2 ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((t & 31) == 0) s_cnt[t >> 5] = n;
    __syncthreads();
    if (t < 32) {
        int v = s_cnt[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (t == 0) counts[c] = v;
    }
}

__global__ void __launch_bounds__(1024) radius_fill_kernel(const float* __restrict__ centers, const float* __restrict__ pts, int N,
                                                           float r2, const long long* __restrict__ offsets, int* __restrict__ indices)
{
    __shared__ int s_scan[32];
    __shared__ int s_base;
    const int c = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float cx = centers[c * 3], cy = centers[c * 3 + 1], cz = centers[c * 3 + 2];
    int* out = indices + offsets[c];
    if (t == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += 1024) {
        const int i = i0 + t;
        const bool in = i < N && sqdist3(pts[i * 3] - cx, pts[i * 3 + 1] - cy, pts[i * 3 + 2] - cz) <= r2;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        const int within = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_scan[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            const int w = s_scan[lane];
            int inc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            s_scan[lane] = inc - w;
        }
        __syncthreads();
        const int base = s_base;
        if (in) out[base + s_scan[warp] + within] = i;
        __syncthreads();
        if (t == 1023) s_base = base + s_scan[31] + __popc(m);
        __syncthreads();
    }
}

// centers [P,3], pts [N,3] -> counts int32 [P]
P2PB_API int p2pb_radius_count(const float* centers, const float* pts, int P, int N, float radius, int* counts, void* stream)
{
    P2PB_CHECK_ARG(P >= 0 && N > 0 && radius >= 0.f, "radius_count: bad sizes");
    if (P == 0) return P2PB_OK;
    radius_count_kernel<<<P, 1024, 0, (cudaStream_t)stream>>>(centers, pts, N, radius * radius, counts);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// offsets int64 [P+1] = exclusive scan of counts (done by the caller); indices int32 [offsets[P]], ascending per centre
P2PB_API int p2pb_radius_fill(const float* centers, const float* pts, int P, int N, float radius, const long long* offsets,
                              int* indices, void* stream)
{
    P2PB_CHECK_ARG(P >= 0 && N > 0 && radius >= 0.f, "radius_fill: bad sizes");
    if (P == 0) return P2PB_OK;
    radius_fill_kernel<<<P, 1024, 0, (cudaStream_t)stream>>>(centers, pts, N, radius * radius, offsets, indices);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// =========================================================================================================
// Point-to-mesh distances for the P2M / P2F evaluation metrics: replaces
//   pytorch3d._C.point_face_dist_forward / face_point_dist_forward as called by metrics/p2m.py:23-160, 307-375
//   (point_mesh_face_distance_custom) from metrics/metrics.py:196-225 (point_face_dist).
// pytorch3d is an un-vendored dependency (requirements.txt, unpinned): parity is UNPINNED at that boundary; what is
// restated here is its published point-triangle distance (pytorch3d/csrc/utils/geometry_utils.cuh, v0.7):
//   n = cross(v2-v0, v1-v0), |n| = norm(n), n /= (|n| + 1e-8);  t = dot(v0-p, n);  p0 = p + t n   (projection on the plane)
//   inside = area(v0,v1,v2) >= min_triangle_area  and  all barycentric coordinates of p0 in [0,1]
//            (w1 = (d11 d20 - d01 d21)/den, w2 = (d00 d21 - d01 d20)/den, w0 = 1-w1-w2, den = d00 d11 - d01^2 + 1e-8)
//   d = t^2 if inside and |n| > 1e-8, else the smallest squared distance to the three edge SEGMENTS
//   (segment: l2 = |v1-v0|^2 <= 1e-8 -> |p-v1|^2, else clamp(dot(v1-v0, p-v0)/l2, 0, 1)).
// min_triangle_area defaults to 5e-3 in THIS fork (metrics/p2m.py:20): smaller triangles are treated as their edges.
// One launch gives both directions: every thread owns a point and scans all triangles (staged in shared memory) for
// point->face; the per-triangle minimum over the CTA's points is reduced in the warp and merged with an integer atomicMin
// on the float bits (distances are >= 0, so the bit pattern orders like the value; min is order-independent => deterministic).
// =========================================================================================================
__device__ __forceinline__ float p2f_dot(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }

__device__ __forceinline__ float p2f_segment(float px, float py, float pz, const float* a, const float* b)
{
    const float ex = b[0] - a[0], ey = b[1] - a[1], ez = b[2] - a[2];
    const float l2 = p2f_dot(ex, ey, ez, ex, ey, ez);
    if (l2 <= 1e-8f) {
        const float dx = px - b[0], dy = py - b[1], dz = pz - b[2];
        return p2f_dot(dx, dy, dz, dx, dy, dz);
    }
    float t = p2f_dot(ex, ey, ez, px - a[0], py - a[1], pz - a[2]) / l2;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float dx = px - (a[0] + t * ex), dy = py - (a[1] + t * ey), dz = pz - (a[2] + t * ez);
    return p2f_dot(dx, dy, dz, dx, dy, dz);
}

__device__ __forceinline__ float p2f_point_triangle(float px, float py, float pz, const float* tri, float min_area)
{
    const float* v0 = tri, *v1 = tri + 3, *v2 = tri + 6;
    const float ax = v2[0] - v0[0], ay = v2[1] - v0[1], az = v2[2] - v0[2];      // v2 - v0
    const float bx = v1[0] - v0[0], by = v1[1] - v0[1], bz = v1[2] - v0[2];      // v1 - v0
    float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;   // cross(v2-v0, v1-v0)
    const float nn = sqrtf(p2f_dot(nx, ny, nz, nx, ny, nz));
    const float inv = 1.0f / (nn + 1e-8f);
    nx *= inv; ny *= inv; nz *= inv;
    const float t = p2f_dot(v0[0] - px, v0[1] - py, v0[2] - pz, nx, ny, nz);
    bool inside = false;
    if (0.5f * nn >= min_area) {                         // area = |cross| / 2
        const float qx = px + t * nx - v0[0], qy = py + t * ny - v0[1], qz = pz + t * nz - v0[2];     // p0 - v0
        const float d00 = p2f_dot(bx, by, bz, bx, by, bz), d01 = p2f_dot(bx, by, bz, ax, ay, az), d11 = p2f_dot(ax, ay, az, ax, ay, az);
        const float d20 = p2f_dot(qx, qy, qz, bx, by, bz), d21 = p2f_dot(qx, qy, qz, ax, ay, az);
        const float den = d00 * d11 - d01 * d01 + 1e-8f;
        const float w1 = (d11 * d20 - d01 * d21) / den, w2 = (d00 * d21 - d01 * d20) / den, w0 = 1.0f - w1 - w2;
        inside = w0 >= 0.0f && w0 <= 1.0f && w1 >= 0.0f && w1 <= 1.0f && w2 >= 0.0f && w2 <= 1.0f;
    }
    if (inside && nn > 1e-8f) return t * t;
    const float e01 = p2f_segment(px, py, pz, v0, v1), e02 = p2f_segment(px, py, pz, v0, v2), e12 = p2f_segment(px, py, pz, v1, v2);
    float d = e01 > e02 ? e02 : e01;
    return d > e12 ? e12 : d;
}

#define P2F_TILE 512
__global__ void __launch_bounds__(256) p2f_kernel(const float* __restrict__ pts, int P, const float* __restrict__ tris, int T, float min_area,
                                                  float* __restrict__ point_dist, unsigned int* __restrict__ face_bits)
{
    __shared__ float s_tri[P2F_TILE * 9];
    const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    const bool ok = i < P;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (ok) {
        px = pts[(size_t)i * 3];
        py = pts[(size_t)i * 3 + 1];
        pz = pts[(size_t)i * 3 + 2];
    }
    float best = INFINITY;
    for (int t0 = 0; t0 < T; t0 += P2F_TILE) {
        const int len = min(P2F_TILE, T - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < len * 9; k += 256) s_tri[k] = tris[(size_t)t0 * 9 + k];
        __syncthreads();
        for (int k = 0; k < len; ++k) {
            float d = ok ? p2f_point_triangle(px, py, pz, s_tri + k * 9, min_area) : INFINITY;
            best = fminf(best, d);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) d = fminf(d, __shfl_xor_sync(0xffffffffu, d, m));
            if (lane == 0 && d < INFINITY) atomicMin(face_bits + t0 + k, __float_as_uint(d));
        }
    }
    if (ok) point_dist[i] = best;
}

// pts [P,3], tris [T,3,3] -> point_dist [P] (squared distance of every point to its closest triangle) and face_dist [T] (squared
// distance of every triangle to its closest point); face_dist must not alias anything, it is initialised here
P2PB_API int p2pb_point_face_dist(const float* pts, int P, const float* tris, int T, float min_triangle_area, float* point_dist,
                                  float* face_dist, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(P > 0 && T > 0, "point_face_dist: bad sizes P=%d T=%d", P, T);
    P2PB_CUDA_OK(cudaMemsetAsync(face_dist, 0x7f, sizeof(float) * (size_t)T, s));      // 0x7f7f7f7f = 3.39e38 as float bits
    p2f_kernel<<<p2pb_cdiv(P, 256), 256, 0, s>>>(pts, P, tris, T, min_triangle_area, point_dist, reinterpret_cast<unsigned int*>(face_dist));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
