// engine_kernels.cu -- fused streaming kernels of the channels-last engine (everything of the PVCNN U-Net evaluation
// that is not a tensor-core contraction).  They replace the ~600 un-fused ATen launches per network evaluation of the
// reference (GroupNorm, AdaGN mul/add, Swish, SE, torch.cat, expand, max, zeros) with a handful of
// HBM-bound passes:  rows are channels-last fp32 [M, ld] (ld multiple of 4), one thread per float4, coalesced.
//
//   voxelize_cl      CSR gather-mean of point rows (+ broadcast time-embedding channels) -> dense grid rows
//   gn_coef          (sum, sum^2) partials of a GEMM/conv epilogue -> per-(sample, channel) affine of GroupNorm/AdaGN
//   affine_act       y = act(x*A[b,c] + B[b,c])      (+ max over K consecutive rows, + max over all rows of a sample)
//   devox_cl         trilinear gather of the raw conv output with AdaGN*SE folded in + point branch (AdaGN+Swish) add
//   group_rows       ball-query neighbourhood gather: [features[idx], xyz[idx]-centre]
//   interp_rows      3-NN inverse-distance interpolation gather
//   linear_small     per-sample small Linear (AdaGN/temb/SE/attention-sized matvecs), warp per output
//   attention_small  bottleneck LinearAttention core (softmax over tokens, 32x32 context per head)
//   bridge_update    pred_x0 = xt - std*eps ; xt <- mu_x0*pred_x0 + mu_xn*xt           (p2pb.py:155-165,190-213)
#include <cuda_fp16.h>

#include "common.cuh"

// Flat element indices are split with 32-bit unsigned divisions (a 64-bit division costs ~10x more ALU work than the
// 16 bytes each thread moves); every launcher checks that the element count fits.
// store 4 consecutive channels as fp32 or as IEEE half (round to nearest): the conv operands of the half path
// Range guard of the IEEE-half operand storage: every value the engine rounds to half is checked against half's largest finite
// value; a hit is counted here (the engine reads the counter after the graph replay and re-runs the call with fp32 / tf32 operand
// storage, engine.py).  The reference's arithmetic (fp32 storage, TF32 multiply) has 8 exponent bits, half has 5.
__device__ unsigned int g_half_overflow = 0;
__device__ __forceinline__ void half_range_check(float m)
{
    if (!(m <= 65504.0f)) atomicAdd(&g_half_overflow, 1u);      // also catches NaN
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v)
{
    half_range_check(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&lo);
    u.y = *reinterpret_cast<const unsigned*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

// values that did not fit IEEE half since the last reset (0 = the half operand storage was exact-range); synchronises `stream`
P2PB_API int p2pb_half_overflow_count(int reset, unsigned int* host_count, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int v = 0;
    P2PB_CUDA_OK(cudaMemcpyFromSymbolAsync(&v, g_half_overflow, sizeof(v), 0, cudaMemcpyDeviceToHost, s));
    P2PB_CUDA_OK(cudaStreamSynchronize(s));
    if (host_count != nullptr) *host_count = v;
    if (reset && v != 0) {
        const unsigned int z = 0;
        P2PB_CUDA_OK(cudaMemcpyToSymbolAsync(g_half_overflow, &z, sizeof(z), 0, cudaMemcpyHostToDevice, s));
        P2PB_CUDA_OK(cudaStreamSynchronize(s));
    }
    return P2PB_OK;
}

#define P2PB_CHECK_U32(total, what) P2PB_CHECK_ARG((total) < 4294967296LL, what ": %lld elements exceed the 32-bit index range, split the batch", (long long)(total))

// ---------------------------------------------------------------------------------------------------------
// coords [B,3,N] -> rows [B*N, ld] columns col0..col0+2 (other columns untouched)
// ---------------------------------------------------------------------------------------------------------
__global__ void coords_to_rows_kernel(const float* __restrict__ coords, float* __restrict__ rows, int N, int ld, int col0,
                                      long long total)
{
    P2PB_PDL_SYNC();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / N;
    const int n = (int)(i - b * N);
    const float* c = coords + b * 3 * N;
    float* o = rows + i * ld + col0;
    o[0] = c[n];
    o[1] = c[n + N];
    o[2] = c[n + 2 * N];
}

__global__ void coords_to_rows_f16_kernel(const float* __restrict__ coords, __half* __restrict__ rows, int N, int ld, int col0,
                                          long long total)
{
    P2PB_PDL_SYNC();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / N;
    const int n = (int)(i - b * N);
    const float* c = coords + b * 3 * N;
    __half* o = rows + i * ld + col0;
    half_range_check(fmaxf(fabsf(c[n]), fmaxf(fabsf(c[n + N]), fabsf(c[n + 2 * N]))));
    o[0] = __float2half_rn(c[n]);
    o[1] = __float2half_rn(c[n + N]);
    o[2] = __float2half_rn(c[n + 2 * N]);
}

// same, IEEE-half rows (A operand of p2pb_gemm_rows_f16)
P2PB_API int p2pb_coords_to_rows_f16(const float* coords, void* rows, int B, int N, int ld, int col0, void* stream)
{
    const long long total = (long long)B * N;
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)coords_to_rows_f16_kernel);
    (void)p2pb_launch(coords_to_rows_f16_kernel, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, coords, reinterpret_cast<__half*>(rows), N, ld,
                                                                                    col0, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

P2PB_API int p2pb_coords_to_rows(const float* coords, float* rows, int B, int N, int ld, int col0, void* stream)
{
    const long long total = (long long)B * N;
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)coords_to_rows_kernel);
    (void)p2pb_launch(coords_to_rows_kernel, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, coords, rows, N, ld, col0, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// Scatter-mean of one voxel for the 4 channels c0..c0+3: sum over the voxel's n points (CSR order = ascending point index, the
// oracle's order; vox_gpu.cu:70-75 adds with atomics in no particular order) of feat[p, c] * (1/n); channels Cf <= c < Cf+E get the
// same mean of the broadcast time embedding, channels beyond that 0.  The loads of 16 points are in flight together and a chunk that
// straddles Cf still uses 16-byte loads (rows are padded to ldf; what is read beyond Cf is discarded): a voxel that holds hundreds
// of points is otherwise a chain of dependent L2 round trips, three of them per point in the straddling chunk (late bridge steps
// concentrate the points: measured 66 -> 270 us for the first voxelisation between the 1st and the 24th step of the bench).
__device__ __forceinline__ float4 voxel_mean4(const float* __restrict__ feat, int ldf, int Cf, const float* __restrict__ temb_b, int E,
                                              const int* __restrict__ ord, int n, int c0)
{
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const float inv = (float)(1.0 / (double)(float)n);
    if (c0 < Cf) {
        if ((ldf & 3) == 0 && c0 + 3 < ldf) {
            int i = 0;
            for (; i + 16 <= n; i += 16) {
                int o[16];
                float4 f[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) o[u] = ord[i + u];
#pragma unroll
                for (int u = 0; u < 16; ++u) f[u] = __ldg(reinterpret_cast<const float4*>(feat + (size_t)o[u] * ldf + c0));
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    a[0] = __fadd_rn(a[0], __fmul_rn(f[u].x, inv));
                    a[1] = __fadd_rn(a[1], __fmul_rn(f[u].y, inv));
                    a[2] = __fadd_rn(a[2], __fmul_rn(f[u].z, inv));
                    a[3] = __fadd_rn(a[3], __fmul_rn(f[u].w, inv));
                }
            }
            for (; i < n; ++i) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(feat + (size_t)ord[i] * ldf + c0));
                a[0] = __fadd_rn(a[0], __fmul_rn(f.x, inv));
                a[1] = __fadd_rn(a[1], __fmul_rn(f.y, inv));
                a[2] = __fadd_rn(a[2], __fmul_rn(f.z, inv));
                a[3] = __fadd_rn(a[3], __fmul_rn(f.w, inv));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (c0 + k >= Cf) continue;
                float s = 0.f;
                for (int i = 0; i < n; ++i) s = __fadd_rn(s, __fmul_rn(__ldg(feat + (size_t)ord[i] * ldf + c0 + k), inv));
                a[k] = s;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        if (c < Cf) continue;
        float s = 0.f;
        if (c < Cf + E) {
            const float t = __fmul_rn(__ldg(temb_b + (c - Cf)), inv);
            for (int i = 0; i < n; ++i) s = __fadd_rn(s, t);
        }
        a[k] = s;
    }
    return make_float4(a[0], a[1], a[2], a[3]);
}

// ---------------------------------------------------------------------------------------------------------
// voxelize_cl: out[b, v, c] for every voxel v and channel c < Cp (each element written exactly once)
//   c <  Cf        : sum over the voxel's points (ascending index) of feat[b, p, c] * (1/cnt)    (vox_gpu.cu:70-75)
//   Cf <= c < Cf+E : the same scatter-mean applied to the broadcast time embedding temb[b, c-Cf]
//   else           : 0 (channel padding for the 32-wide K chunks of the conv)
// ---------------------------------------------------------------------------------------------------------
template <typename OUT>
__global__ void __launch_bounds__(256) voxelize_cl_kernel(const float* __restrict__ feat, int ldf, int Cf,
                                                          const float* __restrict__ temb, int E,
                                                          const int* __restrict__ order, const int* __restrict__ start,
                                                          const int* __restrict__ cnt, OUT* __restrict__ out, int Cp,
                                                          int N, int r3, unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned C4 = Cp >> 2;
    const unsigned vrow = e / C4;  // b*r3 + v
    const int c0 = (int)(e - vrow * C4) * 4;
    const int b = (int)(vrow / (unsigned)r3);
    const int n = cnt[vrow];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > 0)
        acc = voxel_mean4(feat + (size_t)b * N * ldf, ldf, Cf, temb != nullptr ? temb + (size_t)b * E : nullptr, E,
                          order + (size_t)b * N + start[vrow], n, c0);
    store4(out + (size_t)e * 4, acc);
}

P2PB_API int p2pb_voxelize_cl(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order,
                              const int* start, const int* cnt, float* out, int Cp, int B, int N, int r, void* stream)
{
    P2PB_CHECK_ARG(Cp % 4 == 0 && Cf + E <= Cp && Cf > 0, "voxelize_cl: bad channels Cf=%d E=%d Cp=%d", Cf, E, Cp);
    const int r3 = r * r * r;
    const long long total4 = (long long)B * r3 * (Cp / 4);
    P2PB_CHECK_U32(total4, "voxelize_cl");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)voxelize_cl_kernel<float>);
    (void)p2pb_launch(voxelize_cl_kernel<float>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, ldf, Cf, temb, E, order, start, cnt, out,
                                                                                      Cp, N, r3, total4);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// same with an IEEE-half grid (A operand of p2pb_conv3d_cl_f16); fp32 sums, rounded once
P2PB_API int p2pb_voxelize_cl_f16(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order, const int* start,
                                  const int* cnt, void* out, int Cp, int B, int N, int r, void* stream)
{
    P2PB_CHECK_ARG(Cp % 4 == 0 && Cf + E <= Cp && Cf > 0, "voxelize_cl_f16: bad channels Cf=%d E=%d Cp=%d", Cf, E, Cp);
    const int r3 = r * r * r;
    const long long total4 = (long long)B * r3 * (Cp / 4);
    P2PB_CHECK_U32(total4, "voxelize_cl_f16");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)voxelize_cl_kernel<__half>);
    (void)p2pb_launch(voxelize_cl_kernel<__half>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, ldf, Cf, temb, E, order, start, cnt,
                                                                                       reinterpret_cast<__half*>(out), Cp, N, r3, total4);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// gn_coef: GroupNorm (+ AdaGN) statistics -> per-(sample, channel) affine.
//   stats  [B*tiles, C, 2]  per-tile column (sum, sum^2) written by the GEMM epilogue (tiles per sample = tiles)
//   y = ((x - mean_g) * rstd_g * gamma_c + beta_c) * factor_bc + bias_bc     =  x * A[b,c] + Bc[b,c]
//   emd (optional) [B, ld_emd]: factor at column emd_off + c, bias at emd_off + C + c      (modules.py:341-358)
//   ymean (optional) [B, C]: mean over rows of y (= A*mean_c + Bc), the SE squeeze          (modules.py:378)
// One CTA per (sample, group); fp64 accumulation of the partials (deterministic, no atomics anywhere).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_coef_kernel(const float* __restrict__ stats, int tiles, int C, int groups, float count,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ emd, int ld_emd, int emd_off, float eps,
                                                      float* __restrict__ coefA, float* __restrict__ coefB,
                                                      float* __restrict__ ymean)
{
    P2PB_PDL_SYNC();
    // threads = (channel within group) x (tile slice): every thread sums a strided slice of the tiles for one channel in
    // fp64; slices, then channels, are combined by fixed binary trees in shared memory -> deterministic.  The kernel is
    // pure latency (48 launches per evaluation): the affine parameters are fetched before the reduction, the tile loop is
    // unrolled so that its loads are in flight together, and no thread walks a serial list.
    __shared__ double s_part[256][2];
    __shared__ double s_ch[128][2];
    const int b = blockIdx.x / groups, g = blockIdx.x % groups;
    const int cpg = C / groups;
    const int t = threadIdx.x;
    const int slices = 256 / cpg;              // cpg is a power of two <= 128 for every layer of the network
    const int c = t % cpg, sl = t / cpg;
    const int ch = g * cpg + c;
    float p_gamma = 0.f, p_beta = 0.f, p_f = 1.f, p_eb = 0.f;
    if (t < cpg) {
        p_gamma = gamma[ch];
        p_beta = beta[ch];
        if (emd != nullptr) {
            p_f = emd[(size_t)b * ld_emd + emd_off + ch];
            p_eb = emd[(size_t)b * ld_emd + emd_off + C + ch];
        }
    }
    double s = 0.0, q = 0.0;
    {
        const float* sp = stats + ((size_t)b * tiles * C + ch) * 2;
#pragma unroll 4
        for (int tl = sl; tl < tiles; tl += slices) {
            const float2 p = *reinterpret_cast<const float2*>(sp + (size_t)tl * C * 2);
            s += (double)p.x;
            q += (double)p.y;
        }
    }
    s_part[t][0] = s;
    s_part[t][1] = q;
    __syncthreads();
    for (int stride = slices >> 1; stride > 0; stride >>= 1) {       // slices -> channel sums (rows 0..cpg-1)
        if (sl < stride) {
            s_part[t][0] += s_part[t + stride * cpg][0];
            s_part[t][1] += s_part[t + stride * cpg][1];
        }
        __syncthreads();
    }
    if (t < cpg) {
        s_ch[t][0] = s_part[t][0];
        s_ch[t][1] = s_part[t][1];
    }
    __syncthreads();
    for (int stride = cpg >> 1; stride > 0; stride >>= 1) {          // channel sums -> group sums (row 0)
        if (t < stride) {
            s_part[t][0] += s_part[t + stride][0];
            s_part[t][1] += s_part[t + stride][1];
        }
        __syncthreads();
    }
    const double gsum = s_part[0][0], gsq = s_part[0][1];
    if (t < cpg) {
        const double n = (double)count * cpg;
        const double mean = gsum / n;
        double var = gsq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float meanf = (float)mean;
        const int chn = g * cpg + t;
        float a = rstd * p_gamma;
        float bb = p_beta - meanf * a;
        if (emd != nullptr) {
            a *= p_f;
            bb = bb * p_f + p_eb;
        }
        coefA[(size_t)b * C + chn] = a;
        coefB[(size_t)b * C + chn] = bb;
        if (ymean != nullptr) ymean[(size_t)b * C + chn] = a * (float)(s_ch[t][0] / (double)count) + bb;
    }
}

P2PB_API int p2pb_gn_coef(const float* stats, int tiles, int B, int C, int groups, int rows_per_sample, const float* gamma,
                          const float* beta, const float* emd, int ld_emd, int emd_off, float eps, float* coefA, float* coefB,
                          float* ymean, void* stream)
{
    const int cpg = groups > 0 ? C / groups : 0;
    P2PB_CHECK_ARG(groups > 0 && C % groups == 0 && cpg <= 128 && (cpg & (cpg - 1)) == 0,
                   "gn_coef: C=%d groups=%d (group width must be a power of two <= 128)", C, groups);
    if (B == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)gn_coef_kernel);
    (void)p2pb_launch(gn_coef_kernel, dim3(B * groups), dim3(256), (size_t)(0), (cudaStream_t)stream, stats, tiles, C, groups, (float)rows_per_sample, gamma, beta, emd,
                                                                ld_emd, emd_off, eps, coefA, coefB, ymean);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// column (sum, sum^2) of rows for samples whose row count is not a multiple of the 128-row GEMM tile
// (deep U-Net levels: 8..32 points per patch): out [B, C, 2] == the epilogue format with tiles = 1
__global__ void col_stats_kernel(const float* __restrict__ x, int ld, int rows, int C, float* __restrict__ out)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float* p = x + (size_t)b * rows * ld + c;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < rows; ++i) {
        const float v = p[(size_t)i * ld];
        s += v;
        q += v * v;
    }
    out[((size_t)b * C + c) * 2 + 0] = s;
    out[((size_t)b * C + c) * 2 + 1] = q;
}

P2PB_API int p2pb_col_stats(const float* x, int ld, int B, int rows, int C, float* out, void* stream)
{
    if (B == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)col_stats_kernel);
    (void)p2pb_launch(col_stats_kernel, dim3(dim3(p2pb_cdiv(C, 128), B)), dim3(128), (size_t)(0), (cudaStream_t)stream, x, ld, rows, C, out);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

int g_p2pb_act_grid = 0;     // development aid: CTAs per SM of the activation passes' persistent grid (0 = one CTA per 1024 float4)
P2PB_API int p2pb_set_act_grid(int ctas_per_sm)
{
    g_p2pb_act_grid = ctas_per_sm;
    return P2PB_OK;
}
static inline unsigned p2pb_act_grid(long long total4)
{
    const long long full = (total4 + 1023) / 1024;
    if (g_p2pb_act_grid <= 0) return (unsigned)full;
    const long long cap = (long long)p2pb_num_sms() * g_p2pb_act_grid;
    return (unsigned)(full < cap ? full : cap);
}

static inline int p2pb_log2_exact(long long v)      // log2(v) if v is a power of two, else -1
{
    if (v <= 0 || (v & (v - 1)) != 0) return -1;
    int l = 0;
    while ((1LL << l) < v) ++l;
    return l;
}

// ---------------------------------------------------------------------------------------------------------
// affine_act: out[m, c] = act(x[m, c] * A[b, c] + Bc[b, c]),  b = m / rows_per_sample,  act: 0 none, 1 swish
//   pool == 1 : out rows [M, ldo]
//   pool  > 1 : max over `pool` consecutive rows -> out rows [M/pool, ldo]            (neighbour max, pvcnn.py:414)
//   gmax != 0 : additionally atomic-max over all rows of the sample -> gmax[b, c]      (global max-pool, pvcnn.py:923,930)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v)
{
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

template <int ACT>
__device__ __forceinline__ float4 affine4(float4 x, float4 a, float4 b)
{
    float4 y = make_float4(fmaf(x.x, a.x, b.x), fmaf(x.y, a.y, b.y), fmaf(x.z, a.z, b.z), fmaf(x.w, a.w, b.w));
    if (ACT == 1) {
        y.x = swishf(y.x); y.y = swishf(y.y); y.z = swishf(y.z); y.w = swishf(y.w);
    }
    return y;
}

// 4 float4 per thread, all four x loads issued before anything else and ONLY they live while in flight: the (L1-resident)
// coefficient rows are fetched after the x loads return, which keeps the kernel at 40 registers = 6 CTAs per SM = 96 KB of loads in
// flight per SM.  Measured at 64 patches (tools/bench_act.py): 1 load per thread 4.0 TB/s; 4 loads with the coefficients loaded
// alongside (76 registers) 4.9 TB/s; this form 6.4 TB/s.  x is read once -> evict-first loads.
// lC4 / lrps >= 0: C/4 resp. rows_per_sample are powers of two (every layer of the network) -> shifts instead of divisions
__device__ __forceinline__ float4 ld_stream4(const float* p)     // read-once input: evict-first
{
    return __ldcs(reinterpret_cast<const float4*>(p));
}

template <int ACT, typename OUT>
__global__ void __launch_bounds__(256, 5) affine_act_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ A,
                                                            const float* __restrict__ Bc, int rows_per_sample, int C,
                                                            OUT* __restrict__ out, int ldo, unsigned total4, int lC4, int lrps)
{
    P2PB_PDL_SYNC();
    const unsigned C4 = C >> 2;
    // grid-stride over blocks of 4 x blockDim float4 (the launcher picks a persistent grid for large tensors)
    for (unsigned base = blockIdx.x * (blockDim.x * 4); base < total4; base += gridDim.x * (blockDim.x * 4)) {
        const unsigned e0 = base + threadIdx.x;
        float4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned e = e0 + k * blockDim.x;
            if (e < total4) {
                const unsigned mu = lC4 >= 0 ? (e >> lC4) : e / C4;
                xv[k] = ld_stream4(x + (size_t)mu * ldx + (e - mu * C4) * 4);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned e = e0 + k * blockDim.x;
            if (e < total4) {
                const unsigned mu = lC4 >= 0 ? (e >> lC4) : e / C4;
                const int c = (int)(e - mu * C4) * 4;
                const size_t b = lrps >= 0 ? (mu >> lrps) : mu / (unsigned)rows_per_sample;
                const float4 a = __ldg(reinterpret_cast<const float4*>(A + b * C + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(Bc + b * C + c));
                store4(out + (size_t)mu * ldo + c, affine4<ACT>(xv[k], a, bb));
            }
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(256) affine_act_pool_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ A,
                                                              const float* __restrict__ Bc, int rows_per_sample, int C,
                                                              int pool, float* __restrict__ out, int ldo, unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned C4 = C >> 2;
    const unsigned mou = e / C4;  // pooled row
    const int c = (int)(e - mou * C4) * 4;
    const size_t mo = mou;
    const size_t m0 = mo * pool;
    const size_t b = (mou * (unsigned)pool) / (unsigned)rows_per_sample;
    const float4 a = __ldg(reinterpret_cast<const float4*>(A + b * C + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(Bc + b * C + c));
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int i = 0; i < pool; ++i) {
        const float4 y = affine4<ACT>(*reinterpret_cast<const float4*>(x + (m0 + i) * ldx + c), a, bb);
        mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
    }
    *reinterpret_cast<float4*>(out + mo * ldo + c) = mx;
}

// global max over the rows of each sample: grid (row tiles, B); each thread owns 4 channels and walks `rows_per_cta` rows
template <int ACT>
__global__ void __launch_bounds__(256) affine_act_gmax_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ A,
                                                              const float* __restrict__ Bc, int rows_per_sample, int C,
                                                              int rows_per_cta, float* __restrict__ out, int ldo,
                                                              float* __restrict__ gmax)
{
    P2PB_PDL_SYNC();
    const int b = blockIdx.y;
    const int C4 = C >> 2;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
        const int c = c4 * 4;
        const float4 a = __ldg(reinterpret_cast<const float4*>(A + (size_t)b * C + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(Bc + (size_t)b * C + c));
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int i = 0; i < rows_per_cta; ++i) {
            const long long rr = r0 + i;
            if (rr >= rows_per_sample) break;
            const long long m = (long long)b * rows_per_sample + rr;
            const float4 y = affine4<ACT>(*reinterpret_cast<const float4*>(x + m * ldx + c), a, bb);
            if (out != nullptr) *reinterpret_cast<float4*>(out + m * ldo + c) = y;
            mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
        }
        float* g = gmax + (size_t)b * C + c;
        atomic_max_float(g + 0, mx.x); atomic_max_float(g + 1, mx.y); atomic_max_float(g + 2, mx.z); atomic_max_float(g + 3, mx.w);
    }
}

// global max over the rows of a sample of act(x*A + Bc) from the per-tile column (max, min) of x written by the GEMM
// epilogue: x -> fma(x, A, Bc) is monotone and Swish is unimodal (decreasing, then increasing), so the maximum over any
// set of rows is attained at the largest or the smallest x of the column.  colmm [B*tiles, C, 2].
__global__ void __launch_bounds__(256) gmax_minmax_kernel(const float* __restrict__ colmm, int tiles, int C,
                                                          const float* __restrict__ A, const float* __restrict__ Bc, int act,
                                                          float* __restrict__ gmax, int total)
{
    P2PB_PDL_SYNC();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int b = e / C, c = e - b * C;
    float mx = -INFINITY, mn = INFINITY;
    for (int t = 0; t < tiles; ++t) {
        const float2 p = *reinterpret_cast<const float2*>(colmm + (((size_t)b * tiles + t) * C + c) * 2);
        mx = fmaxf(mx, p.x);
        mn = fminf(mn, p.y);
    }
    const float a = A[e], bb = Bc[e];
    float y0 = fmaf(mx, a, bb), y1 = fmaf(mn, a, bb);
    if (act == 1) {
        y0 = swishf(y0);
        y1 = swishf(y1);
    }
    gmax[e] = fmaxf(y0, y1);
}

P2PB_API int p2pb_gmax_minmax(const float* colmm, int tiles, int B, int C, const float* A, const float* Bc, int act, float* gmax,
                              void* stream)
{
    const int total = B * C;
    if (total == 0) return P2PB_OK;
    P2PB_CHECK_ARG(tiles > 0, "gmax_minmax: tiles must be positive");
    p2pb_prefer_max_smem((const void*)gmax_minmax_kernel);
    (void)p2pb_launch(gmax_minmax_kernel, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, colmm, tiles, C, A, Bc, act, gmax, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// Neighbourhood max-pool (max over the K = 32 grouped rows of a centre, pvcnn.py:414) of act(x*A + Bc) from the GEMM
// epilogue's column (max, min) per 32-row block -- one block IS one neighbourhood -- by the same argument as
// gmax_minmax_kernel (monotone affine map, unimodal Swish): the [B*M*32, C] pre-activation of the last shared-MLP layer is
// never written and the pooling pass over it disappears.  colmm [B*M, C, 2]; A, Bc [B, C]; out rows [B*M, ldo].
__global__ void __launch_bounds__(256) pool32_minmax_kernel(const float* __restrict__ colmm, int M, int C,
                                                            const float* __restrict__ A, const float* __restrict__ Bc, int act,
                                                            float* __restrict__ out, int ldo, unsigned total)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const unsigned row = e / (unsigned)C;        // b*M + j
    const int c = (int)(e - row * (unsigned)C);
    const unsigned b = row / (unsigned)M;
    const float2 p = *reinterpret_cast<const float2*>(colmm + (size_t)e * 2);
    const float a = A[(size_t)b * C + c], bb = Bc[(size_t)b * C + c];
    float y0 = fmaf(p.x, a, bb), y1 = fmaf(p.y, a, bb);
    if (act == 1) {
        y0 = swishf(y0);
        y1 = swishf(y1);
    }
    out[(size_t)row * ldo + c] = fmaxf(y0, y1);
}

P2PB_API int p2pb_pool32_minmax(const float* colmm, int B, int M, int C, const float* A, const float* Bc, int act, float* out,
                                int ldo, void* stream)
{
    const long long total = (long long)B * M * C;
    P2PB_CHECK_U32(total, "pool32_minmax");
    if (total == 0) return P2PB_OK;
    P2PB_CHECK_ARG(ldo >= C, "pool32_minmax: ldo < C");
    p2pb_prefer_max_smem((const void*)pool32_minmax_kernel);
    (void)p2pb_launch(pool32_minmax_kernel, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, colmm, M, C, A, Bc,
                      act, out, ldo, (unsigned)total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

__global__ void fill_kernel(float* p, float v, long long n)
{
    P2PB_PDL_SYNC();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

P2PB_API int p2pb_affine_act(const float* x, int ldx, const float* A, const float* Bc, int rows_per_sample, int M, int C, int act,
                             int pool, float* out, int ldo, float* gmax, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && (out == nullptr || ldo % 4 == 0), "affine_act: C/ld must be multiples of 4");
    P2PB_CHECK_ARG(M % rows_per_sample == 0 && pool >= 1 && rows_per_sample % pool == 0, "affine_act: bad row partition");
    if (M == 0) return P2PB_OK;
    const int B = M / rows_per_sample;
    if (gmax != nullptr) {
        P2PB_CHECK_ARG(pool == 1, "affine_act: gmax and pool are exclusive");
        p2pb_prefer_max_smem((const void*)fill_kernel);
        (void)p2pb_launch(fill_kernel, dim3(p2pb_cdiv((long long)B * C, 256)), dim3(256), (size_t)(0), s, gmax, -INFINITY, (long long)B * C);
        P2PB_LAUNCH_OK();
        int rows_per_cta = p2pb_cdiv(rows_per_sample, p2pb_cdiv(4 * p2pb_num_sms(), B));
        if (rows_per_cta < 8) rows_per_cta = 8;
        dim3 grid(p2pb_cdiv(rows_per_sample, rows_per_cta), B);
        if (act) { p2pb_prefer_max_smem((const void*)affine_act_gmax_kernel<1>); (void)p2pb_launch(affine_act_gmax_kernel<1>, dim3(grid), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, rows_per_cta, out, ldo, gmax); }
        else { p2pb_prefer_max_smem((const void*)affine_act_gmax_kernel<0>); (void)p2pb_launch(affine_act_gmax_kernel<0>, dim3(grid), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, rows_per_cta, out, ldo, gmax); }
        P2PB_LAUNCH_OK();
        return P2PB_OK;
    }
    P2PB_CHECK_ARG(out != nullptr, "affine_act: out required");
    if (pool == 1) {
        const long long total4 = (long long)M * (C / 4);
        P2PB_CHECK_U32(total4, "affine_act");
        if (act) { p2pb_prefer_max_smem((const void*)affine_act_kernel<1, float>); (void)p2pb_launch(affine_act_kernel<1, float>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, out, ldo, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(rows_per_sample)); }
        else { p2pb_prefer_max_smem((const void*)affine_act_kernel<0, float>); (void)p2pb_launch(affine_act_kernel<0, float>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, out, ldo, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(rows_per_sample)); }
    } else {
        const long long total4 = (long long)(M / pool) * (C / 4);
        P2PB_CHECK_U32(total4, "affine_act(pool)");
        if (act) { p2pb_prefer_max_smem((const void*)affine_act_pool_kernel<1>); (void)p2pb_launch(affine_act_pool_kernel<1>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, pool, out, ldo, total4); }
        else { p2pb_prefer_max_smem((const void*)affine_act_pool_kernel<0>); (void)p2pb_launch(affine_act_pool_kernel<0>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, pool, out, ldo, total4); }
    }
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// affine_act with IEEE-half output rows (A operand of the half GEMM / conv entry points); no pooling
P2PB_API int p2pb_affine_act_f16(const float* x, int ldx, const float* A, const float* Bc, int rows_per_sample, int M, int C, int act,
                                 void* out, int ldo, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && out != nullptr, "affine_act_f16: C/ld must be multiples of 4");
    P2PB_CHECK_ARG(rows_per_sample > 0 && M % rows_per_sample == 0, "affine_act_f16: bad row partition");
    if (M == 0) return P2PB_OK;
    const long long total4 = (long long)M * (C / 4);
    P2PB_CHECK_U32(total4, "affine_act_f16");
    __half* o = reinterpret_cast<__half*>(out);
    if (act) { p2pb_prefer_max_smem((const void*)affine_act_kernel<1, __half>); (void)p2pb_launch(affine_act_kernel<1, __half>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, o, ldo, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(rows_per_sample)); }
    else { p2pb_prefer_max_smem((const void*)affine_act_kernel<0, __half>); (void)p2pb_launch(affine_act_kernel<0, __half>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), s, x, ldx, A, Bc, rows_per_sample, C, o, ldo, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(rows_per_sample)); }
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// devox_cl: out[b,n,c] = (sum_k w_k * raw[b, corner_k, c]) * (A[b,c]*se[b,c]) + (Bc[b,c]*se[b,c]) * sum_k w_k
//                        + swish(praw[b,n,c] * pA[b,c] + pB[b,c])
// i.e. trilinear devoxelisation (trilinear_devox_gpu.cu:21-109) of  SE(AdaGN(conv2 output))  without ever materialising
// the normalised grid, plus the PVConv point branch (pvcnn.py:324-328).  raw grid rows [B*r^3, ldg].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) devox_cl_kernel(const float* __restrict__ ncoords, const float* __restrict__ raw, int ldg,
                                                       const float* __restrict__ A, const float* __restrict__ Bc,
                                                       const float* __restrict__ se, const float* __restrict__ praw, int ldp,
                                                       const float* __restrict__ pA, const float* __restrict__ pB,
                                                       float* __restrict__ out, int ldo, int C, int N, int r, unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned C4 = C >> 2;
    const unsigned mu = e / C4;  // b*N + n
    const int c = (int)(e - mu * C4) * 4;
    const size_t m = mu;
    const int b = (int)(mu / (unsigned)N);
    const int n = (int)(mu - (unsigned)b * (unsigned)N);
    const float* co = ncoords + (size_t)b * 3 * N;
    const float x = co[n], y = co[n + N], z = co[n + 2 * N];
    const int r2 = r * r;
    const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
    const float xd1 = x - xl, yd1 = y - yl, zd1 = z - zl;
    const float xd0 = 1.0f - xd1, yd0 = 1.0f - yd1, zd0 = 1.0f - zd1;
    const int base = (int)xl * r2 + (int)yl * r + (int)zl;
    const int xh = (xd1 > 0) ? r2 : 0, yh = (yd1 > 0) ? r : 0, zh = (zd1 > 0) ? 1 : 0;
    const float w[8] = {xd0 * yd0 * zd0, xd0 * yd0 * zd1, xd0 * yd1 * zd0, xd0 * yd1 * zd1,
                        xd1 * yd0 * zd0, xd1 * yd0 * zd1, xd1 * yd1 * zd0, xd1 * yd1 * zd1};
    const int off[8] = {0, zh, yh, yh + zh, xh, xh + zh, xh + yh, xh + yh + zh};
    const float* g = raw + ((size_t)b * r2 * r + base) * ldg + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(g + (size_t)off[k] * ldg));
        acc.x = fmaf(w[k], f.x, acc.x); acc.y = fmaf(w[k], f.y, acc.y);
        acc.z = fmaf(w[k], f.z, acc.z); acc.w = fmaf(w[k], f.w, acc.w);
        wsum += w[k];
    }
    float4 a = __ldg(reinterpret_cast<const float4*>(A + (size_t)b * C + c));
    float4 bb = __ldg(reinterpret_cast<const float4*>(Bc + (size_t)b * C + c));
    if (se != nullptr) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(se + (size_t)b * C + c));
        a.x *= s.x; a.y *= s.y; a.z *= s.z; a.w *= s.w;
        bb.x *= s.x; bb.y *= s.y; bb.z *= s.z; bb.w *= s.w;
    }
    float4 o = make_float4(fmaf(acc.x, a.x, bb.x * wsum), fmaf(acc.y, a.y, bb.y * wsum), fmaf(acc.z, a.z, bb.z * wsum),
                           fmaf(acc.w, a.w, bb.w * wsum));
    if (praw != nullptr) {
        const float4 p = *reinterpret_cast<const float4*>(praw + m * ldp + c);
        const float4 pa = __ldg(reinterpret_cast<const float4*>(pA + (size_t)b * C + c));
        const float4 pb = __ldg(reinterpret_cast<const float4*>(pB + (size_t)b * C + c));
        const float4 q = affine4<1>(p, pa, pb);
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    *reinterpret_cast<float4*>(out + m * ldo + c) = o;
}

P2PB_API int p2pb_devox_cl(const float* ncoords, const float* raw, int ldg, const float* A, const float* Bc, const float* se,
                           const float* praw, int ldp, const float* pA, const float* pB, float* out, int ldo, int B, int C, int N,
                           int r, void* stream)
{
    P2PB_CHECK_ARG(C % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0 && (praw == nullptr || ldp % 4 == 0), "devox_cl: alignment");
    const long long total4 = (long long)B * N * (C / 4);
    P2PB_CHECK_U32(total4, "devox_cl");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)devox_cl_kernel);
    (void)p2pb_launch(devox_cl_kernel, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, ncoords, raw, ldg, A, Bc, se, praw, ldp, pA, pB, out,
                                                                            ldo, C, N, r, total4);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// group_rows: out[(b*M + j)*U + k, 0:Cf] = feat[b, idx[b,j,k], 0:Cf] ; out[.., Cf:Cf+3] = xyz[idx] - centre_j
// (pvcnn_grouping_gpu.cu:18-39 twice + the centre subtraction + torch.cat of pvcnn.py:117-126).  Columns >= Cf+3 are
// never written (zero from allocation).  One thread per (row, 4-channel chunk); chunk index Cf/4 carries the xyz part.
// ---------------------------------------------------------------------------------------------------------
template <typename OUT>
__global__ void __launch_bounds__(256) group_rows_kernel(const float* __restrict__ feat, int ldf, int Cf,
                                                         const float* __restrict__ coords, const float* __restrict__ centers,
                                                         const int* __restrict__ idx, OUT* __restrict__ out, int ldo, int N,
                                                         int M, int U, unsigned total)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const unsigned C4 = (Cf >> 2) + 1;
    const unsigned rowu = e / C4;  // (b*M + j)*U + k
    const int c4 = (int)(e - rowu * C4);
    const size_t row = rowu;
    const unsigned bj = rowu / (unsigned)U;
    const int b = (int)(bj / (unsigned)M);
    const int j = (int)(bj - (unsigned)b * (unsigned)M);
    const int src = idx[row];
    if (c4 < (Cf >> 2)) {
        store4(out + row * ldo + c4 * 4, __ldg(reinterpret_cast<const float4*>(feat + ((size_t)b * N + src) * ldf + c4 * 4)));
    } else {
        const float* co = coords + (size_t)b * 3 * N;
        const float* ce = centers + (size_t)b * 3 * M;
        OUT* o = out + row * ldo + Cf;
        o[0] = (OUT)(co[src] - ce[j]);
        o[1] = (OUT)(co[src + N] - ce[j + M]);
        o[2] = (OUT)(co[src + 2 * N] - ce[j + 2 * M]);
    }
}

P2PB_API int p2pb_group_rows(const float* feat, int ldf, int Cf, const float* coords, const float* centers, const int* idx,
                             float* out, int ldo, int B, int N, int M, int U, void* stream)
{
    P2PB_CHECK_ARG(Cf % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0 && ldo >= Cf + 3, "group_rows: alignment (Cf %% 4, ld %% 4)");
    const long long total = (long long)B * M * U * (Cf / 4 + 1);
    P2PB_CHECK_U32(total, "group_rows");
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)group_rows_kernel<float>);
    (void)p2pb_launch(group_rows_kernel<float>, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, ldf, Cf, coords, centers, idx, out, ldo, N, M,
                                                                             U, total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// same with IEEE-half output rows (A operand of p2pb_gemm_rows_f16); ldo in halves
P2PB_API int p2pb_group_rows_f16(const float* feat, int ldf, int Cf, const float* coords, const float* centers, const int* idx,
                                 void* out, int ldo, int B, int N, int M, int U, void* stream)
{
    P2PB_CHECK_ARG(Cf % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0 && ldo >= Cf + 3, "group_rows_f16: alignment (Cf %% 4, ld %% 4)");
    const long long total = (long long)B * M * U * (Cf / 4 + 1);
    P2PB_CHECK_U32(total, "group_rows_f16");
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)group_rows_kernel<__half>);
    (void)p2pb_launch(group_rows_kernel<__half>, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, ldf, Cf,
                      coords, centers, idx, reinterpret_cast<__half*>(out), ldo, N, M, U, (unsigned)total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// group_project: first shared-MLP layer of a set-abstraction module WITHOUT materialising the grouped tensor.
// The layer is linear in its grouped input [features[idx], xyz[idx] - centre] (pvcnn.py:117-126, 174-192), so
//   v[b, j, k, c] = Pf[b, idx[b,j,k], c] + Wx[c] . (xyz[b, idx] - centre[b, j])
// with Pf = features @ Wf^T + bias (+ time-embedding fold) computed ONCE per point by the GEMM (N rows instead of M*32),
// and the 3-channel coordinate part evaluated here in fp32 from the fp32 difference (more accurate than rounding it to a
// tensor-core operand).  One warp per centre, lanes over channels.
//   mode 0: GroupNorm partials (sum, sum^2) over the centre's 32 rows -> stats [B*M, C, 2]   (the 32-row-block format)
//   mode 1: y = swish(v*A + Bc) -> IEEE-half rows [B*M*32, ldo] (A operand of the next layer's GEMM)
// Replaces group_rows + the [B*M*32, C_in] grouped buffer + the first GEMM's [B*M*32, C] output + its activation pass.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 4) group_project_kernel(const float* __restrict__ Pf, int ldp, const float* __restrict__ Wx,
                                                            const float* __restrict__ coords, const float* __restrict__ centers,
                                                            const int* __restrict__ idx, const float* __restrict__ A,
                                                            const float* __restrict__ Bc, float* __restrict__ stats,
                                                            __half* __restrict__ out, int ldo, int C, int N, int M, int total_centers)
{
    P2PB_PDL_SYNC();
    // The kernel is instruction-bound (a gather of L2-resident rows): every lane owns TWO adjacent channels per 64-channel pass,
    // so the four shuffles that broadcast neighbour k's (index, dx, dy, dz) are shared by 64 channels, the gather is one 8-byte
    // load and the half output one 4-byte store per lane.
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= total_centers) return;
    const int b = warp / M, j = warp - b * M;
    const float* co = coords + (size_t)b * 3 * N;
    const float cx = centers[((size_t)b * 3 + 0) * M + j], cy = centers[((size_t)b * 3 + 1) * M + j], cz = centers[((size_t)b * 3 + 2) * M + j];
    const int my_src = idx[(size_t)warp * 32 + lane];
    const float mdx = co[my_src] - cx, mdy = co[my_src + N] - cy, mdz = co[my_src + 2 * N] - cz;
    const float* pf_b = Pf + (size_t)b * N * ldp;
    for (int c0 = 0; c0 < C; c0 += 64) {
        const int c = c0 + 2 * lane;
        const bool act = c < C;                         // C % 32 == 0: the last pass of C = 32 (mod 64) uses half of the lanes
        const int cc = act ? c : 0;
        const float wx0 = Wx[cc * 3], wy0 = Wx[cc * 3 + 1], wz0 = Wx[cc * 3 + 2];
        const float wx1 = Wx[cc * 3 + 3], wy1 = Wx[cc * 3 + 4], wz1 = Wx[cc * 3 + 5];
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        if (MODE == 1) {
            const float2 av = *reinterpret_cast<const float2*>(A + (size_t)b * C + cc), bv = *reinterpret_cast<const float2*>(Bc + (size_t)b * C + cc);
            a0 = av.x; a1 = av.y; b0 = bv.x; b1 = bv.y;
        }
        float s10 = 0.f, s20 = 0.f, s11 = 0.f, s21 = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const int src = __shfl_sync(0xffffffffu, my_src, k);
            const float dx = __shfl_sync(0xffffffffu, mdx, k), dy = __shfl_sync(0xffffffffu, mdy, k), dz = __shfl_sync(0xffffffffu, mdz, k);
            const float2 pv = *reinterpret_cast<const float2*>(pf_b + (size_t)src * ldp + cc);
            float v0 = fmaf(wx0, dx, pv.x), v1 = fmaf(wx1, dx, pv.y);
            v0 = fmaf(wy0, dy, v0);
            v1 = fmaf(wy1, dy, v1);
            v0 = fmaf(wz0, dz, v0);
            v1 = fmaf(wz1, dz, v1);
            if (MODE == 0) {
                s10 += v0;
                s20 = fmaf(v0, v0, s20);
                s11 += v1;
                s21 = fmaf(v1, v1, s21);
            } else {
                const float y0 = swishf(fmaf(v0, a0, b0)), y1 = swishf(fmaf(v1, a1, b1));
                half_range_check(fmaxf(fabsf(y0), fabsf(y1)));
                if (act) *reinterpret_cast<__half2*>(out + ((size_t)warp * 32 + k) * ldo + c) = __floats2half2_rn(y0, y1);
            }
        }
        if (MODE == 0 && act) *reinterpret_cast<float4*>(stats + ((size_t)warp * C + c) * 2) = make_float4(s10, s20, s11, s21);
    }
}

// Pf [B*N, ldp] fp32, Wx [C, 3], coords [B,3,N], centers [B,3,M], idx [B,M,32]; mode 0 -> stats [B*M, C, 2];
// mode 1 (A, Bc [B, C]) -> out half rows [B*M*32, ldo]
P2PB_API int p2pb_group_project(const float* Pf, int ldp, const float* Wx, const float* coords, const float* centers, const int* idx,
                                const float* A, const float* Bc, float* stats, void* out, int ldo, int B, int C, int N, int M, int U,
                                int mode, void* stream)
{
    P2PB_CHECK_ARG(U == 32 && C % 32 == 0 && C > 0, "group_project: needs 32 neighbours per centre and C %% 32 == 0 (U=%d C=%d)", U, C);
    P2PB_CHECK_ARG(mode == 0 ? stats != nullptr : (out != nullptr && A != nullptr && Bc != nullptr && ldo >= C && ldo % 2 == 0), "group_project: bad outputs for mode %d", mode);
    P2PB_CHECK_ARG(ldp % 2 == 0, "group_project: ldp=%d must be even", ldp);
    const long long total = (long long)B * M;
    if (total == 0) return P2PB_OK;
    P2PB_CHECK_U32(total * 32, "group_project");
    const dim3 grid(p2pb_cdiv(total * 32, 256)), block(256);
    if (mode == 0) {
        p2pb_prefer_max_smem((const void*)group_project_kernel<0>);
        (void)p2pb_launch(group_project_kernel<0>, grid, block, (size_t)0, (cudaStream_t)stream, Pf, ldp, Wx, coords, centers, idx, A, Bc,
                          stats, reinterpret_cast<__half*>(out), ldo, C, N, M, (int)total);
    } else {
        p2pb_prefer_max_smem((const void*)group_project_kernel<1>);
        (void)p2pb_launch(group_project_kernel<1>, grid, block, (size_t)0, (cudaStream_t)stream, Pf, ldp, Wx, coords, centers, idx, A, Bc,
                          stats, reinterpret_cast<__half*>(out), ldo, C, N, M, (int)total);
    }
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// interp_rows: out[b*N + n, 0:C] = f[i2]*w2 + f[i1]*w1 + f[i3]*w3 (reference contraction order) from rows [B*M, ldf]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) interp_rows_kernel(const float* __restrict__ f, int ldf, const int* __restrict__ idx,
                                                          const float* __restrict__ w, float* __restrict__ out, int ldo, int C,
                                                          int N, int M, unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned C4 = C >> 2;
    const unsigned mu = e / C4;
    const int c = (int)(e - mu * C4) * 4;
    const size_t m = mu;
    const int b = (int)(mu / (unsigned)N);
    const int n = (int)(mu - (unsigned)b * (unsigned)N);
    const int* ix = idx + (size_t)b * 3 * N;
    const float* ww = w + (size_t)b * 3 * N;
    const int i1 = ix[n], i2 = ix[n + N], i3 = ix[n + 2 * N];
    const float w1 = ww[n], w2 = ww[n + N], w3 = ww[n + 2 * N];
    const float* fb = f + (size_t)b * M * ldf + c;
    const float4 f1 = __ldg(reinterpret_cast<const float4*>(fb + (size_t)i1 * ldf));
    const float4 f2 = __ldg(reinterpret_cast<const float4*>(fb + (size_t)i2 * ldf));
    const float4 f3 = __ldg(reinterpret_cast<const float4*>(fb + (size_t)i3 * ldf));
    float4 o;
    o.x = fmaf(f3.x, w3, fmaf(f1.x, w1, f2.x * w2));
    o.y = fmaf(f3.y, w3, fmaf(f1.y, w1, f2.y * w2));
    o.z = fmaf(f3.z, w3, fmaf(f1.z, w1, f2.z * w2));
    o.w = fmaf(f3.w, w3, fmaf(f1.w, w1, f2.w * w2));
    *reinterpret_cast<float4*>(out + m * ldo + c) = o;
}

P2PB_API int p2pb_interp_rows(const float* f, int ldf, const int* idx, const float* w, float* out, int ldo, int B, int C, int N,
                              int M, void* stream)
{
    P2PB_CHECK_ARG(C % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0, "interp_rows: alignment");
    const long long total4 = (long long)B * N * (C / 4);
    P2PB_CHECK_U32(total4, "interp_rows");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)interp_rows_kernel);
    (void)p2pb_launch(interp_rows_kernel, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, f, ldf, idx, w, out, ldo, C, N, M, total4);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// linear_small: out[b, o] = act(sum_k in[b, k] * W[o, k] + bias[o]);  warp per (b, o);  act: 0 none, 1 swish, 2 relu,
// 3 sigmoid, 4 leaky_relu(0.1).  For the per-sample vectors of the network (time-embedding MLP unet_pvc.py:52-56,
// time-embedding bias folds, SE excitation modules.py:365-370, global-feature bias fold of Pnet2Stage).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ in, int ldi, const float* __restrict__ W,
                                                           int ldw, const float* __restrict__ bias, int K, int O, int act,
                                                           float* __restrict__ out, int ldo, long long total)
{
    P2PB_PDL_SYNC();
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int o = (int)(wid % O);
    const long long b = wid / O;
    const float* x = in + b * ldi;
    const float* wr = W + (size_t)o * ldw;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(x[k], __ldg(wr + k), s);
    s = warp_sum(s);
    if (lane == 0) {
        if (bias != nullptr) s += bias[o];
        if (act == 1) s = swishf(s);
        else if (act == 2) s = fmaxf(s, 0.f);
        else if (act == 3) s = 1.0f / (1.0f + __expf(-s));
        else if (act == 4) s = s > 0.f ? s : 0.1f * s;
        out[b * ldo + o] = s;
    }
}

P2PB_API int p2pb_linear_small(const float* in, int ldi, const float* W, int ldw, const float* bias, int B, int K, int O, int act,
                               float* out, int ldo, void* stream)
{
    const long long total = (long long)B * O;
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)linear_small_kernel);
    (void)p2pb_launch(linear_small_kernel, dim3(p2pb_cdiv(total * 32, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, in, ldi, W, ldw, bias, K, O, act, out, ldo,
                                                                                    total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// step_vectors: everything of one sampling step that depends only on the time embedding, in ONE launch:
//   temb = Linear(LeakyReLU_0.1(Linear(sinusoid)))                                   (embedf, unet_pvc.py:52-56)
//   fold_i = W_i[:, time columns] @ temb  for every layer whose input is cat[features, time_emb] in front of a 1x1 conv
//            (PVConv point branch, SA-module MLP, FP-module first Conv1d: the 64 constant channels are a per-sample bias)
// All fold weights are stacked in Wall [R, E]; row r of the stack writes to  row_ptr[r] + b * row_stride[r]  (each fold keeps
// its own dense [B, cout] buffer, the GEMM epilogue's bias2 operand).  Replaces 2 + (number of folds) linear_small launches.
// grid (ceil(R / 64), B): every CTA recomputes the 2-layer MLP of its sample (8K FMAs) and produces 64 stacked rows.
// ---------------------------------------------------------------------------------------------------------
// dot product of a shared-memory vector with a global weight row, split over TPO consecutive lanes (TPO a power of two <= 32): every
// thread issues its K / TPO loads back to back (one memory latency per layer instead of one per output), then log2(TPO) shuffles
template <int TPO>
__device__ __forceinline__ float dot_tpo(const float* __restrict__ xs, const float* __restrict__ wrow, int K, int sub)
{
    float s = 0.f;
#pragma unroll 4
    for (int k = sub; k < K; k += TPO) s = fmaf(xs[k], __ldg(wrow + k), s);
#pragma unroll
    for (int m = TPO >> 1; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    return s;
}

__global__ void __launch_bounds__(256) step_vectors_kernel(const float* __restrict__ sin, int ld_sin, const float* __restrict__ w0,
                                                           const float* __restrict__ b0, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, int E, const float* __restrict__ Wall, int R,
                                                           const long long* __restrict__ row_ptr, const int* __restrict__ row_stride,
                                                           float* __restrict__ temb_out)
{
    P2PB_PDL_SYNC();
    __shared__ float s_in[128], s_h[128], s_t[128];
    const int b = blockIdx.y, t = threadIdx.x;
    const int sub = t & 3, o4 = t >> 2;                 // 4 threads per output, 64 outputs per pass
    // prefetch this CTA's slice of the stacked fold weights while the two small layers run (independent of them)
    const int r = blockIdx.x * 64 + o4;
    float wreg[32];                                     // E <= 128 -> at most 32 values per thread
    if (r < R) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int k = sub + 4 * i;
            wreg[i] = k < E ? __ldg(Wall + (size_t)r * E + k) : 0.f;
        }
    }
    for (int k = t; k < E; k += 256) s_in[k] = sin[(size_t)b * ld_sin + k];
    __syncthreads();
    for (int o = o4; o < E; o += 64) {
        float s = dot_tpo<4>(s_in, w0 + (size_t)o * E, E, sub) + b0[o];
        if (sub == 0) s_h[o] = s > 0.f ? s : 0.1f * s;
    }
    __syncthreads();
    for (int o = o4; o < E; o += 64) {
        const float s = dot_tpo<4>(s_h, w2 + (size_t)o * E, E, sub) + b2[o];
        if (sub == 0) {
            s_t[o] = s;
            if (blockIdx.x == 0) temb_out[(size_t)b * E + o] = s;
        }
    }
    __syncthreads();
    float s = 0.f;
    if (r < R) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int k = sub + 4 * i;
            if (k < E) s = fmaf(s_t[k], wreg[i], s);
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (r < R && sub == 0) reinterpret_cast<float*>(row_ptr[r])[(size_t)b * row_stride[r]] = s;
}

P2PB_API int p2pb_step_vectors(const float* sin, int ld_sin, const float* w0, const float* b0, const float* w2, const float* b2, int B,
                               int E, const float* Wall, int R, const long long* row_ptr, const int* row_stride, float* temb_out,
                               void* stream)
{
    P2PB_CHECK_ARG(B > 0 && E > 0 && E <= 128 && R >= 0, "step_vectors: bad sizes B=%d E=%d R=%d (E <= 128)", B, E, R);
    p2pb_prefer_max_smem((const void*)step_vectors_kernel);
    const int gx = R > 0 ? p2pb_cdiv(R, 64) : 1;
    (void)p2pb_launch(step_vectors_kernel, dim3(gx, B), dim3(256), (size_t)(0), (cudaStream_t)stream, sin, ld_sin, w0, b0, w2, b2, E, Wall, R,
                      row_ptr, row_stride, temb_out);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// se_excite: squeeze-excitation gate of a PVConv (modules.py:362-378) from the per-sample channel means the GroupNorm
// coefficient kernel already produced: se = sigmoid(W2 @ relu(W0 @ ymean)).  One CTA per sample (both Linears, bias-free).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_excite_kernel(const float* __restrict__ ym, const float* __restrict__ w0,
                                                        const float* __restrict__ w2, int C, int Hd, float* __restrict__ se)
{
    P2PB_PDL_SYNC();
    extern __shared__ float s_se[];       // [C] means, [Hd] hidden
    float* s_m = s_se;
    float* s_hid = s_se + C;
    const int b = blockIdx.x, t = threadIdx.x;
    for (int k = t; k < C; k += 256) s_m[k] = ym[(size_t)b * C + k];
    __syncthreads();
    {   // hidden layer: 8 threads per output, 32 outputs per pass (Hd = C / 8 <= 64)
        const int sub = t & 7;
        for (int o = t >> 3; o < Hd; o += 32) {
            const float s = dot_tpo<8>(s_m, w0 + (size_t)o * C, C, sub);
            if (sub == 0) s_hid[o] = fmaxf(s, 0.f);
        }
    }
    __syncthreads();
    for (int o = t; o < C; o += 256) {     // gate: one thread per channel, Hd independent loads each
        const float* wr = w2 + (size_t)o * Hd;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < Hd; ++k) s = fmaf(s_hid[k], __ldg(wr + k), s);
        se[(size_t)b * C + o] = 1.0f / (1.0f + __expf(-s));
    }
}

P2PB_API int p2pb_se_excite(const float* ymean, const float* w0, const float* w2, int B, int C, int Hd, float* se, void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && C > 0 && Hd > 0 && (C + Hd) * 4 <= 48 * 1024, "se_excite: bad sizes C=%d hidden=%d", C, Hd);
    if (B == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)se_excite_kernel);
    (void)p2pb_launch(se_excite_kernel, dim3(B), dim3(256), (size_t)((C + Hd) * sizeof(float)), (cudaStream_t)stream, ymean, w0, w2, C, Hd, se);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// head_bridge: the end of one sampling step in one pass over the classifier's hidden rows:
//   h = Swish(GroupNorm(raw))  (coefficients A, Bc from gn_coef)  ->  eps = W[3, C] h + bias   (classifier, unet_pvc.py:147-154,
//   263-267; Dropout is the identity in eval)  ->  pred_x0 = xt - std*eps ; (clip) ; xt <- mu_x0*pred_x0 + mu_xn*xt
//   (p2pb.py:155-165, 190-213, same operation order as the reference; scalars from the device table `coef`).
// The [B*N, C] activation is never written, the 128 -> 3 projection runs in fp32 FMAs (the reference's Conv1d is TF32).
// One warp per point: lanes stride the channels (coalesced 128-byte reads), three shuffle reductions.
// ---------------------------------------------------------------------------------------------------------
// A warp owns 32 CONSECUTIVE points of one sample: per point the lanes stride the channels (coalesced 16-byte loads, 4 points in
// flight), the three dot products are reduced by shuffles, and lane j keeps the result of point j -- so the bridge update reads and
// writes xt / pred_x0 [B,3,N] with unit stride across the warp.
__global__ void __launch_bounds__(256) head_bridge_kernel(const float* __restrict__ raw, int ldr, const float* __restrict__ A,
                                                          const float* __restrict__ Bc, const float* __restrict__ W,
                                                          const float* __restrict__ bias, int C, int N, long long total,
                                                          const float* __restrict__ xt, const float* __restrict__ coef, int clip,
                                                          float* __restrict__ xt_next, float* __restrict__ pred_x0,
                                                          float* __restrict__ eps_out, int lde)
{
    P2PB_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const int groups_per_sample = (N + 31) >> 5;
    const long long gw = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;          // warp = (sample, 32-point group)
    const long long b = gw / groups_per_sample;
    if (b * N >= total) return;
    const int n0 = (int)(gw - b * groups_per_sample) << 5;
    const int cnt = min(32, N - n0);
    const float* a = A + b * C;
    const float* bb = Bc + b * C;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;      // lane j: eps of point n0 + j
    for (int c = lane * 4; c < C; c += 128) {
        const float4 av = __ldg(reinterpret_cast<const float4*>(a + c)), bv = __ldg(reinterpret_cast<const float4*>(bb + c));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + c)), w1 = __ldg(reinterpret_cast<const float4*>(W + C + c));
        const float4 w2 = __ldg(reinterpret_cast<const float4*>(W + 2 * C + c));
        const float* x = raw + ((size_t)b * N + n0) * ldr + c;
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)j * ldr);
            const float h0 = swishf(fmaf(xv.x, av.x, bv.x)), h1 = swishf(fmaf(xv.y, av.y, bv.y));
            const float h2 = swishf(fmaf(xv.z, av.z, bv.z)), h3 = swishf(fmaf(xv.w, av.w, bv.w));
            float e0 = fmaf(h0, w0.x, fmaf(h1, w0.y, fmaf(h2, w0.z, h3 * w0.w)));
            float e1 = fmaf(h0, w1.x, fmaf(h1, w1.y, fmaf(h2, w1.z, h3 * w1.w)));
            float e2 = fmaf(h0, w2.x, fmaf(h1, w2.y, fmaf(h2, w2.z, h3 * w2.w)));
            e0 = warp_sum(e0);
            e1 = warp_sum(e1);
            e2 = warp_sum(e2);
            if (lane == j) {
                r0 += e0;
                r1 += e1;
                r2 += e2;
            }
        }
    }
    if (lane < cnt) {
        const int n = n0 + lane;
        const long long pt = b * N + n;
        const float ev[3] = {r0 + bias[0], r1 + bias[1], r2 + bias[2]};
        if (eps_out != nullptr) {
            eps_out[pt * lde] = ev[0];
            eps_out[pt * lde + 1] = ev[1];
            eps_out[pt * lde + 2] = ev[2];
        }
        if (xt != nullptr) {
            const float std_n = coef[0], mu_x0 = coef[1], mu_xn = coef[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const size_t o = ((size_t)b * 3 + k) * N + n;
                const float xv = xt[o];
                float p0 = __fsub_rn(xv, __fmul_rn(std_n, ev[k]));
                if (clip) p0 = fminf(fmaxf(p0, -3.0f), 3.0f);
                if (pred_x0 != nullptr) pred_x0[o] = p0;
                xt_next[o] = __fadd_rn(__fmul_rn(mu_x0, p0), __fmul_rn(mu_xn, xv));
            }
        }
    }
}

// raw [B*N, ldr] (pre-norm classifier hidden rows), A / Bc [B, C], W [3, C], bias [3]; xt [B,3,N] (null: eps only), coef = device
// {std_fwd[n], mu_x0, mu_xn}; xt_next may alias xt; pred_x0 / eps_out (rows [B*N, lde], columns 0..2) optional
P2PB_API int p2pb_head_bridge(const float* raw, int ldr, const float* A, const float* Bc, const float* W, const float* bias, int B, int C,
                              int N, const float* xt, const float* coef, int clip, float* xt_next, float* pred_x0, float* eps_out,
                              int lde, void* stream)
{
    P2PB_CHECK_ARG(B >= 0 && N > 0 && C > 0 && C % 4 == 0 && ldr % 4 == 0 && ldr >= C, "head_bridge: bad sizes C=%d ldr=%d", C, ldr);
    P2PB_CHECK_ARG(xt != nullptr || eps_out != nullptr, "head_bridge: no output requested");
    const long long total = (long long)B * N;
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)head_bridge_kernel);
    const long long warps = (long long)B * ((N + 31) / 32);
    (void)p2pb_launch(head_bridge_kernel, dim3(p2pb_cdiv(warps * 32, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, raw, ldr, A, Bc, W, bias,
                      C, N, total, xt, coef, clip, xt_next, pred_x0, eps_out, lde);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// attention_small: LinearAttention core (modules.py:186-192) on the bottleneck tokens.  qkv rows [B*N, 3*H*32] with
// channel = qkv*H*32 + head*32 + d.  k <- softmax over the N tokens; ctx[d,e] = sum_n k[d,n] v[e,n];
// out[e,n] = sum_d ctx[d,e] q[d,n]  -> out rows [B*N, H*32].  One CTA (32x32 threads) per (sample, head); N <= 64.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) attention_small_kernel(const float* __restrict__ qkv, int ldq, int H, int N,
                                                               float* __restrict__ out, int ldo)
{
    P2PB_PDL_SYNC();
    __shared__ float sq[32][65], sk[32][65], sv[32][65], sctx[32][33];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 32
    const float* base = qkv + (size_t)b * N * ldq;
    for (int n = ty; n < N; n += 32) {
        sq[tx][n] = base[(size_t)n * ldq + 0 * H * 32 + h * 32 + tx];
        sk[tx][n] = base[(size_t)n * ldq + 1 * H * 32 + h * 32 + tx];
        sv[tx][n] = base[(size_t)n * ldq + 2 * H * 32 + h * 32 + tx];
    }
    __syncthreads();
    if (ty == 0) {  // softmax over tokens for row d = tx
        float mx = -INFINITY;
        for (int n = 0; n < N; ++n) mx = fmaxf(mx, sk[tx][n]);
        float s = 0.f;
        for (int n = 0; n < N; ++n) {
            const float ev = expf(sk[tx][n] - mx);
            sk[tx][n] = ev;
            s += ev;
        }
        for (int n = 0; n < N; ++n) sk[tx][n] /= s;
    }
    __syncthreads();
    {  // ctx[d = ty][e = tx]
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(sk[ty][n], sv[tx][n], s);
        sctx[ty][tx] = s;
    }
    __syncthreads();
    for (int n = ty; n < N; n += 32) {  // out[e = tx][n]
        float s = 0.f;
#pragma unroll 8
        for (int d = 0; d < 32; ++d) s = fmaf(sctx[d][tx], sq[d][n], s);
        out[((size_t)b * N + n) * ldo + h * 32 + tx] = s;
    }
}

P2PB_API int p2pb_attention_small(const float* qkv, int ldq, int B, int H, int N, float* out, int ldo, void* stream)
{
    P2PB_CHECK_ARG(N > 0 && N <= 64, "attention_small: N=%d tokens (bottleneck only, <= 64)", N);
    if (B == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)attention_small_kernel);
    (void)p2pb_launch(attention_small_kernel, dim3(B * H), dim3(dim3(32, 32)), (size_t)(0), (cudaStream_t)stream, qkv, ldq, H, N, out, ldo);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// attention_softmax_small: scaled dot-product softmax attention on the bottleneck tokens -- modules.Attention (norm=False, no
// time conditioning, no qk-norm; /root/reference/models/modules.py:197-264) around Attend (:77-162), selected by
// `attention_type: flash` (unet_pvc.py:98-99, 238-241).  qkv rows [B*N, 3*H*32], channel = qkv*H*32 + head*32 + d (to_q, then the
// k and v halves of to_kv, each laid out "(h d)").  out[i, :] = softmax_j(q_i . k_j / sqrt(32)) v_j -> rows [B*N, H*32].
// One CTA (32 x 32 threads) per (sample, head); N <= 64 tokens.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) attention_softmax_small_kernel(const float* __restrict__ qkv, int ldq, int H, int N,
                                                                       float* __restrict__ out, int ldo)
{
    P2PB_PDL_SYNC();
    __shared__ float sq[64][33], sk[64][33], sv[64][33], ss[64][65];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 32
    const float* base = qkv + (size_t)b * N * ldq;
    for (int n = ty; n < N; n += 32) {
        sq[n][tx] = base[(size_t)n * ldq + 0 * H * 32 + h * 32 + tx];
        sk[n][tx] = base[(size_t)n * ldq + 1 * H * 32 + h * 32 + tx];
        sv[n][tx] = base[(size_t)n * ldq + 2 * H * 32 + h * 32 + tx];
    }
    __syncthreads();
    const float scale = 0.17677669529663687f;      // 32^-0.5 (modules.py:139)
    for (int i = ty; i < N; i += 32)
        for (int j = tx; j < N; j += 32) {
            float s = 0.f;
#pragma unroll 8
            for (int d = 0; d < 32; ++d) s = fmaf(sq[i][d], sk[j][d], s);
            ss[i][j] = s * scale;
        }
    __syncthreads();
    for (int i = ty; i < N; i += 32) {             // one warp per row: softmax over j
        float mx = -INFINITY;
        for (int j = tx; j < N; j += 32) mx = fmaxf(mx, ss[i][j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = tx; j < N; j += 32) {
            const float e = expf(ss[i][j] - mx);
            ss[i][j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = tx; j < N; j += 32) ss[i][j] *= inv;
    }
    __syncthreads();
    for (int i = ty; i < N; i += 32) {             // out[i][d = tx]
        float s = 0.f;
        for (int j = 0; j < N; ++j) s = fmaf(ss[i][j], sv[j][tx], s);
        out[((size_t)b * N + i) * ldo + h * 32 + tx] = s;
    }
}

P2PB_API int p2pb_attention_softmax_small(const float* qkv, int ldq, int B, int H, int N, float* out, int ldo, void* stream)
{
    P2PB_CHECK_ARG(N > 0 && N <= 64, "attention_softmax_small: N=%d tokens (bottleneck only, <= 64)", N);
    if (B == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)attention_softmax_small_kernel);
    (void)p2pb_launch(attention_softmax_small_kernel, dim3(B * H), dim3(dim3(32, 32)), (size_t)(0), (cudaStream_t)stream, qkv, ldq, H, N, out, ldo);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// bridge_update (p2pb.py:155-165 + 190-213, ot_ode): eps rows [B*N, lde] (first 3 columns) and xt [B,3,N]
//   pred_x0 = xt - std_n * eps ; (clip +-3) ; xt_next = mu_x0 * pred_x0 + mu_xn * xt      (same op order as the reference)
// Scalars come from a device table coef[step_slot*3 + {0,1,2}] so that one captured graph serves every step.
// ---------------------------------------------------------------------------------------------------------
__global__ void bridge_update_kernel(const float* __restrict__ xt, const float* __restrict__ eps, int lde,
                                     const float* __restrict__ coef, int clip, float* __restrict__ xt_next,
                                     float* __restrict__ pred_x0, int N, long long total)
{
    P2PB_PDL_SYNC();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / N;
    const int n = (int)(i - b * N);
    const float std_n = coef[0], mu_x0 = coef[1], mu_xn = coef[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const size_t o = ((size_t)b * 3 + a) * N + n;
        const float x = xt[o];
        float p0 = __fsub_rn(x, __fmul_rn(std_n, eps[i * lde + a]));
        if (clip) p0 = fminf(fmaxf(p0, -3.0f), 3.0f);
        if (pred_x0 != nullptr) pred_x0[o] = p0;
        xt_next[o] = __fadd_rn(__fmul_rn(mu_x0, p0), __fmul_rn(mu_xn, x));
    }
}

P2PB_API int p2pb_bridge_update(const float* xt, const float* eps, int lde, const float* coef, int clip, float* xt_next,
                                float* pred_x0, int B, int N, void* stream)
{
    const long long total = (long long)B * N;
    if (total == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)bridge_update_kernel);
    (void)p2pb_launch(bridge_update_kernel, dim3(p2pb_cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, xt, eps, lde, coef, clip, xt_next, pred_x0, N,
                                                                                total);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Producers of the zero-bordered row-major conv input  X[b][(r+2)^3][Cp]  (padded-linear voxel index, see
// conv_halo.cu).  Same thread mapping as the dense kernels (one thread per float4, channel chunk fastest: coalesced
// reads and writes); only interior rows are ever written, so the border rows stay zero from allocation.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t padded_row(int b, int v, int r)
{
    const int P = r + 1;     // shared-padding layout, see p2pb_conv_halo_layout (conv_halo.cu)
    const int x = v / (r * r), y = (v / r) % r, z = v % r;
    return (size_t)b * P * P * P + (size_t)(x + 1) * P * P + (size_t)(y + 1) * P + (z + 1);
}

__global__ void __launch_bounds__(256) voxelize_padded_kernel(const float* __restrict__ feat, int ldf, int Cf,
                                                              const float* __restrict__ temb, int E,
                                                              const int* __restrict__ order, const int* __restrict__ start,
                                                              const int* __restrict__ cnt, float* __restrict__ out, int Cp,
                                                              int N, int r, unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned r3 = r * r * r;
    const unsigned C4 = Cp >> 2;
    const unsigned vrow = e / C4;
    const int c0 = (int)(e - vrow * C4) * 4;
    const int b = (int)(vrow / r3);
    const int v = (int)(vrow - (unsigned)b * r3);
    const int n = cnt[vrow];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > 0)
        acc = voxel_mean4(feat + (size_t)b * N * ldf, ldf, Cf, temb != nullptr ? temb + (size_t)b * E : nullptr, E,
                          order + (size_t)b * N + start[vrow], n, c0);
    *reinterpret_cast<float4*>(out + padded_row(b, v, r) * Cp + c0) = acc;
}

P2PB_API int p2pb_voxelize_padded(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order,
                                  const int* start, const int* cnt, float* out, int Cp, int B, int N, int r, void* stream)
{
    P2PB_CHECK_ARG(Cp % 32 == 0 && Cf + E <= Cp && Cf > 0, "voxelize_padded: bad channels Cf=%d E=%d Cp=%d", Cf, E, Cp);
    const long long total4 = (long long)B * r * r * r * (Cp / 4);
    P2PB_CHECK_U32(total4, "voxelize_padded");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)voxelize_padded_kernel);
    (void)p2pb_launch(voxelize_padded_kernel, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, feat, ldf, Cf, temb, E, order, start, cnt, out,
                                                                                   Cp, N, r, total4);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// Sparse form of voxelize_padded for grids that are mostly empty (a 2048-point patch occupies ~5 % of a 32^3 grid):
// the grid is kept all-zero between evaluations; this kernel writes only the occupied voxel rows (clear = 0), and after
// the convolution has consumed the grid the same enumeration zeroes them again (clear = 1).  One thread per
// (sorted point slot, 4 channels); the slot that starts a voxel's CSR range owns the voxel, the others exit.  The sums
// run over the voxel's points in ascending point index exactly like the dense kernel (bit-identical result).
template <typename OUT>
__global__ void __launch_bounds__(256) voxelize_sparse_kernel(const float* __restrict__ feat, int ldf, int Cf,
                                                              const float* __restrict__ temb, int E,
                                                              const int* __restrict__ order, const int* __restrict__ ind,
                                                              const int* __restrict__ start, const int* __restrict__ cnt,
                                                              OUT* __restrict__ out, int Cp, int N, int r, int clear,
                                                              unsigned total4)
{
    P2PB_PDL_SYNC();
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const unsigned r3 = r * r * r;
    const unsigned C4 = Cp >> 2;
    const unsigned slot = e / C4;            // b*N + j
    const int c0 = (int)(e - slot * C4) * 4;
    const int b = (int)(slot / (unsigned)N);
    const int j = (int)(slot - (unsigned)b * N);
    const int p = order[slot];
    const int v = ind[(size_t)b * N + p];
    const size_t vrow = (size_t)b * r3 + v;
    if (start[vrow] != j) return;            // not the first point of its voxel
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!clear)
        acc = voxel_mean4(feat + (size_t)b * N * ldf, ldf, Cf, temb != nullptr ? temb + (size_t)b * E : nullptr, E,
                          order + (size_t)b * N + j, cnt[vrow], c0);
    store4(out + padded_row(b, v, r) * Cp + c0, acc);
}

static int voxelize_sparse_impl(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order, const int* ind,
                                const int* start, const int* cnt, void* out, int Cp, int B, int N, int r, int clear, bool f16,
                                void* stream)
{
    P2PB_CHECK_ARG(Cp % 32 == 0 && Cf + E <= Cp && Cf > 0, "voxelize_padded_sparse: bad channels Cf=%d E=%d Cp=%d", Cf, E, Cp);
    const long long total4 = (long long)B * N * (Cp / 4);
    P2PB_CHECK_U32(total4, "voxelize_padded_sparse");
    if (total4 == 0) return P2PB_OK;
    if (f16) {
        p2pb_prefer_max_smem((const void*)voxelize_sparse_kernel<__half>);
        (void)p2pb_launch(voxelize_sparse_kernel<__half>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
            feat, ldf, Cf, temb, E, order, ind, start, cnt, reinterpret_cast<__half*>(out), Cp, N, r, clear, total4);
    } else {
        p2pb_prefer_max_smem((const void*)voxelize_sparse_kernel<float>);
        (void)p2pb_launch(voxelize_sparse_kernel<float>, dim3(p2pb_cdiv(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
            feat, ldf, Cf, temb, E, order, ind, start, cnt, reinterpret_cast<float*>(out), Cp, N, r, clear, total4);
    }
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

P2PB_API int p2pb_voxelize_padded_sparse(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order,
                                         const int* ind, const int* start, const int* cnt, float* out, int Cp, int B, int N,
                                         int r, int clear, void* stream)
{
    return voxelize_sparse_impl(feat, ldf, Cf, temb, E, order, ind, start, cnt, out, Cp, B, N, r, clear, false, stream);
}

// same, the grid is IEEE half (operand of p2pb_conv3d_halo_f16); sums are formed in fp32 exactly as above, rounded once
P2PB_API int p2pb_voxelize_padded_sparse_f16(const float* feat, int ldf, int Cf, const float* temb, int E, const int* order,
                                             const int* ind, const int* start, const int* cnt, void* out, int Cp, int B, int N,
                                             int r, int clear, void* stream)
{
    return voxelize_sparse_impl(feat, ldf, Cf, temb, E, order, ind, start, cnt, out, Cp, B, N, r, clear, true, stream);
}

// y = swish(x*A + B) of dense conv-output rows [B*r^3, ldx] -> zero-bordered padded input rows of the next conv
// lC4 / lr >= 0: C/4 resp. r are powers of two -> the (sample, x, y, z) split of a voxel row is shifts and masks
template <typename OUT>
__global__ void __launch_bounds__(256, 5) affine_act_padded_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ A,
                                                                   const float* __restrict__ Bc, int C, OUT* __restrict__ out, int ldo,
                                                                   int r, unsigned total4, int lC4, int lr)
{
    P2PB_PDL_SYNC();
    const unsigned r3 = r * r * r;
    const unsigned C4 = C >> 2;
    const int P = r + 1;
    for (unsigned base = blockIdx.x * (blockDim.x * 4); base < total4; base += gridDim.x * (blockDim.x * 4)) {
        const unsigned e0 = base + threadIdx.x;
        float4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // 4 float4 per thread; only the x loads are live while in flight (see affine_act_kernel)
            const unsigned e = e0 + k * blockDim.x;
            if (e < total4) {
                const unsigned vrow = lC4 >= 0 ? (e >> lC4) : e / C4;
                xv[k] = ld_stream4(x + (size_t)vrow * ldx + (e - vrow * C4) * 4);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned e = e0 + k * blockDim.x;
            if (e < total4) {
                const unsigned vrow = lC4 >= 0 ? (e >> lC4) : e / C4;
                const int c = (int)(e - vrow * C4) * 4;
                int b;
                size_t orow;
                if (lr >= 0) {
                    b = (int)(vrow >> (3 * lr));
                    const unsigned v = vrow & (r3 - 1);
                    const int vx = (int)(v >> (2 * lr)), vy = (int)((v >> lr) & (unsigned)(r - 1)), vz = (int)(v & (unsigned)(r - 1));
                    orow = (size_t)b * P * P * P + (size_t)(vx + 1) * P * P + (size_t)(vy + 1) * P + (vz + 1);
                } else {
                    b = (int)(vrow / r3);
                    orow = padded_row(b, (int)(vrow - (unsigned)b * r3), r);
                }
                const float4 a = __ldg(reinterpret_cast<const float4*>(A + (size_t)b * C + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(Bc + (size_t)b * C + c));
                store4(out + orow * ldo + c, affine4<1>(xv[k], a, bb));
            }
        }
    }
}

P2PB_API int p2pb_affine_act_padded(const float* x, int ldx, const float* A, const float* Bc, int B, int C, int r, float* out,
                                    void* stream)
{
    P2PB_CHECK_ARG(C % 32 == 0 && ldx % 4 == 0, "affine_act_padded: C %% 32, ldx %% 4");
    const long long total4 = (long long)B * r * r * r * (C / 4);
    P2PB_CHECK_U32(total4, "affine_act_padded");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)affine_act_padded_kernel<float>);
    (void)p2pb_launch(affine_act_padded_kernel<float>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, ldx, A, Bc, C, out, C, r, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(r));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

// same with IEEE-half output rows of pitch ldo >= C halves (ldo a multiple of 64: the 64-channel chunks of the half conv;
// columns C..ldo-1 are never written and stay zero)
P2PB_API int p2pb_affine_act_padded_f16(const float* x, int ldx, const float* A, const float* Bc, int B, int C, int r, void* out,
                                        int ldo, void* stream)
{
    P2PB_CHECK_ARG(C % 32 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldo >= C, "affine_act_padded_f16: C %% 32, ldx %% 4, ldo >= C");
    const long long total4 = (long long)B * r * r * r * (C / 4);
    P2PB_CHECK_U32(total4, "affine_act_padded_f16");
    if (total4 == 0) return P2PB_OK;
    p2pb_prefer_max_smem((const void*)affine_act_padded_kernel<__half>);
    (void)p2pb_launch(affine_act_padded_kernel<__half>, dim3(p2pb_act_grid(total4)), dim3(256), (size_t)(0), (cudaStream_t)stream,
            x, ldx, A, Bc, C, reinterpret_cast<__half*>(out), ldo, r, total4, p2pb_log2_exact(C / 4), p2pb_log2_exact(r));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}
