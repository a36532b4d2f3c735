// gemm_tf32.cu -- C-ABI entry points of the dense contractions (p2pb_gemm_rows*, p2pb_conv3d_cl) and the FIRST-generation
// one-tile-per-CTA tcgen05 kernel.  The entry points build the TMA descriptors and hand over to the persistent kernel
// (gemm_persist.cu) whenever the shape is inside its envelope; this kernel remains for N = 16 (the 128 -> 3 head), for
// unaligned outputs and as an on-device cross-check (p2pb_debug_set(2)).
//
// One kernel family covers every Conv3d / Conv1d / Conv2d(1x1) / Linear of the hot path
// (reference: cuDNN/cuBLAS library calls behind models/pvcnn.py:174-192,265-284, models/modules.py:337,365-370):
//
//   D[M, N] = A[M, K] * W[N, K]^T (+ bias[N]) (+ bias2[sample(m), N])            fp32 in/out, TF32 MMA, fp32 accumulate
//
//   rows mode : A is up to 3 row-major segments [M, K_i] (channels-last activations; the segments replace torch.cat)
//   conv mode : A is the channels-last voxel grid [B, r, r, r, Cin]; K = 27 taps x Cin.  The im2col matrix is never
//               built: for each tap a 5-D TMA box load shifted by (dx,dy,dz) brings the 128-voxel x 32-channel operand
//               tile, out-of-range voxels are zero-filled by the TMA unit (= the conv's zero padding).
//
// Tile: BM=128 rows (TMEM lanes) x BN<=256 columns (TMEM columns) per CTA, BK=32 fp32 = one 128-byte swizzle span,
// UMMA_K=8 (kind::tf32).  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// thread), warps 2-5 = epilogue (tcgen05.ld -> registers -> bias -> global; optional per-tile column sums
// (sum x, sum x^2) for the GroupNorm that follows, so the normalisation statistics cost no extra pass).
// smem ring of S stages with full/empty mbarriers; accumulator hand-off through one tmem_full mbarrier.
// TF32 is the arithmetic the reference's convolutions run in (cuDNN default).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int UMMA_K = 8;
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 4;  // 16 KiB

struct GemmArgs {
    int dbg;
    int M, ldd, n_total;
    int nseg;
    int seg_chunks[3];
    int conv, bx, by, bz, cin_chunks, tiles_per_sample, r;
    int rows_per_sample;
    const float* bias;
    const float* bias2;
    float* D;
    float* stats;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a converged warp (ptxas then knows a single thread issues the tcgen05/TMA instructions and
// feeds them from uniform registers directly instead of emitting a per-operand R2UR "waterfall" loop)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row x 128 B atoms,
// SBO = 1024 B between atoms along M/N, LBO unused (one atom along K), version 1, layout SWIZZLE_128B (=2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 columns held by each of the 32 lanes (one row per lane) -> lane l ends with the sum over the 32 rows of column l.
// Recursive halving: 16+8+4+2+1 = 31 shuffles.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane)
{
#pragma unroll
    for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            const float send = up ? v[j] : v[j + n / 2];
            const float keep = up ? v[j + n / 2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                 const GemmArgs args, const int stages)
{
    P2PB_PDL_SYNC();
    constexpr int B_STAGE_BYTES = BN * BK * 4;
    constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    constexpr int CW = BN >= 32 ? 32 : 16;  // epilogue chunk width (columns per tcgen05.ld)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A stages][B stages][barriers][tmem ptr][stats scratch]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)stages * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)stages * B_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + stages;
    uint64_t* tmem_full_bar = bars + 2 * stages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = N tile (fastest varying): the CTAs that share one A tile are scheduled together, so A is read from HBM
    // once and served to the others from L2
    const int m_tile = blockIdx.y;
    const int m0 = m_tile * BM;
    const int n0 = blockIdx.x * BN;

    int total_chunks = 0;
    if (args.conv) total_chunks = 27 * args.cin_chunks;
    else
        for (int s = 0; s < args.nseg; ++s) total_chunks += args.seg_chunks[s];

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int cb = 0, cx = 0, cy = 0;
            if (args.conv) {
                cb = m_tile / args.tiles_per_sample;
                const int v0 = (m_tile % args.tiles_per_sample) * BM;  // first voxel of the tile, z fastest
                cx = v0 / (args.r * args.r);
                cy = (v0 / args.r) % args.r;
            }
            int seg = 0, seg_it = 0;
            for (int it = 0; it < total_chunks; ++it) {
                const int st = it % stages;
                const uint32_t ph = (uint32_t)(it / stages) & 1u;
                mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1u);
                const uint32_t fb = smem_u32(&full_bar[st]);
                mbar_expect_tx(fb, A_STAGE_BYTES + B_STAGE_BYTES);
                const uint32_t dstA = smem_u32(sA + (size_t)st * A_STAGE_BYTES);
                if (args.conv) {
                    const int tap = it / args.cin_chunks, kc = it - tap * args.cin_chunks;
                    const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
                    tma_load_5d(dstA, &mapA0, fb, kc * BK, dz, cy + dy, cx + dx, cb);
                } else {
                    while (seg_it >= args.seg_chunks[seg]) {
                        seg_it = 0;
                        ++seg;
                    }
                    const CUtensorMap* mp = seg == 0 ? &mapA0 : (seg == 1 ? &mapA1 : &mapA2);
                    tma_load_2d(dstA, mp, fb, seg_it * BK, m0);
                    ++seg_it;
                }
                tma_load_2d(smem_u32(sB + (size_t)st * B_STAGE_BYTES), &mapB, fb, it * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptor: D=F32, A=B=TF32, K-major both, N>>3 @17, M>>4 @24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int it = 0; it < total_chunks; ++it) {
            const int st = it % stages;
            const uint32_t ph = (uint32_t)(it / stages) & 1u;
            mbar_wait(smem_u32(&full_bar[st]), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t ad = umma_desc_sw128(smem_u32(sA + (size_t)st * A_STAGE_BYTES));
                const uint64_t bd = umma_desc_sw128(smem_u32(sB + (size_t)st * B_STAGE_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advance 32 bytes (8 tf32) inside the 128-byte swizzle span: +2 in the (addr>>4) field
                    umma_tf32(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)((it | k) != 0));
                }
                umma_commit(smem_u32(&empty_bar[st]));
                if (it == total_chunks - 1) umma_commit(smem_u32(tmem_full_bar));
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====================
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < args.M;
        mbar_wait(smem_u32(tmem_full_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* bias2_row = nullptr;
        if (args.bias2 != nullptr && row_ok) bias2_row = args.bias2 + (size_t)(row / args.rows_per_sample) * args.n_total;
        float* drow = args.D + (size_t)row * args.ldd + n0;
#pragma unroll 1
        for (int c = 0; c < BN / CW; ++c) {
            float v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * CW);
            if (CW == 32) tmem_ld32(taddr, v);
            else {
                tmem_ld16(taddr, v);
#pragma unroll
                for (int j = 16; j < 32; ++j) v[j] = 0.f;
            }
            const int nb = n0 + c * CW;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                float x = v[j];
                if (args.bias != nullptr) x += __ldg(args.bias + nb + j);
                if (bias2_row != nullptr) x += __ldg(bias2_row + nb + j);
                v[j] = row_ok ? x : 0.f;
            }
            if (row_ok && !(args.dbg & 1)) {
#pragma unroll
                for (int j = 0; j < CW; j += 4)
                    *reinterpret_cast<float4*>(drow + c * CW + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (args.stats != nullptr) {
                float sq[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
                const float s1 = warp_colsum32(v, lane);
                const float s2 = warp_colsum32(sq, lane);
                // 32-row partials: [(m_tile*4 + q), n_total, 2] (same contract as gemm_persist.cu)
                if (lane < CW)
                    *reinterpret_cast<float2*>(args.stats + ((size_t)(m_tile * 4 + q) * args.n_total + n0 + c * CW + lane) * 2) =
                        make_float2(s1, s2);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box)
{
    PFN_encodeTiled enc = get_encode();
    if (enc == nullptr) {
        p2pb_set_error("cuTensorMapEncodeTiled entry point not available");
        return P2PB_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        p2pb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu/%llu, box %u/%u)", (int)rc, rank,
                       (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return P2PB_ERR_CUDA;
    }
    return P2PB_OK;
}

int g_gemm_dbg = 0;

int pick_bn(int n_total)
{
    const int cands[5] = {256, 128, 64, 32, 16};
    for (int i = 0; i < 5; ++i)
        if (n_total % cands[i] == 0) return cands[i];
    return 0;
}

template <int BN>
int launch_gemm(const CUtensorMap* maps, const GemmArgs& a, cudaStream_t s)
{
    const int b_stage = BN * BK * 4;
    int stages = (200 * 1024) / (A_STAGE_BYTES + b_stage);
    if (stages > 8) stages = 8;
    // short-K GEMMs (1x1 convs on 32..128 channels) need no deep ring: a small footprint lets several CTAs share an SM
    // (TMEM: BN columns each) so that one CTA's epilogue overlaps another's loads/MMAs
    int total_chunks = a.conv ? 27 * a.cin_chunks : a.seg_chunks[0] + a.seg_chunks[1] + a.seg_chunks[2];
    if (stages > total_chunks) stages = total_chunks;
    if (stages < 2) stages = 2;
    const size_t smem = 1024 + (size_t)stages * (A_STAGE_BYTES + b_stage) + (2 * stages + 1) * 8 + 16 + (size_t)4 * BN * 2 * 4;
    static bool attr_set = false;
    if (!attr_set) {
        P2PB_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid(a.n_total / BN, p2pb_cdiv(a.M, BM));
    P2PB_CHECK_ARG(grid.y <= 65535u, "gemm: M=%d needs %u row tiles (> 65535): split the batch", a.M, grid.y);
    p2pb_prefer_max_smem((const void*)gemm_tf32_kernel<BN>);
    (void)p2pb_launch(gemm_tf32_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), (size_t)(smem), s, maps[0], maps[1], maps[2], maps[3], a, stages);
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

int dispatch_gemm(const CUtensorMap* maps, const GemmArgs& a, int bn, cudaStream_t s)
{
    switch (bn) {
        case 256: return launch_gemm<256>(maps, a, s);
        case 128: return launch_gemm<128>(maps, a, s);
        case 64: return launch_gemm<64>(maps, a, s);
        case 32: return launch_gemm<32>(maps, a, s);
        case 16: return launch_gemm<16>(maps, a, s);
    }
    p2pb_set_error("gemm: N=%d must be a multiple of 16", a.n_total);
    return P2PB_ERR_INVALID;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------

// persistent kernel (gemm_persist.cu); returns P2PB_ERR_UNSUPPORTED when the shape is outside its envelope
extern int g_p2pb_gemm_mode;
int p2pb_gemm_persist_try(const CUtensorMap* mapsA, int nseg, const int* seg_chunks, int conv, int cin_chunks, int tiles_per_sample,
                          int r, const float* W, int ktot, const float* bias, const float* bias2, int rows_per_sample, float* D,
                          int ldd, float* stats, float* colmm, int M, int N, cudaStream_t s);

// development aid (tools/, tests): bit 0 = legacy kernel skips its global stores; bit 1 = force the legacy
// one-tile-per-CTA kernel; bit 2 = persistent kernel forms 2-CTA multicast clusters instead of cta_group::2 pairs;
// bits 3, 4 = timing experiments; bit 5 = persistent kernel never forms CTA pairs
P2PB_API int p2pb_debug_set(int flags)
{
    g_gemm_dbg = flags;
    g_p2pb_gemm_mode = flags;
    return P2PB_OK;
}

// rows-mode GEMM:  D[M, N] = sum_i A_i[M, K_i] * W[N, sum K_i]^T + bias + bias2[m / rows_per_sample]
//   A_i : row-major, row pitch lda_i floats (K_i, lda_i multiples of 32 resp. 4), i < nseg <= 3 (replaces torch.cat)
//   W   : [N, Ktot] row-major, Ktot = sum K_i;  N multiple of 16
//   D   : [M, ldd];  stats (optional): [4*ceil(M/128), N, 2] column (sum, sum of squares) of every 32-row block
//   colmm (optional, persistent kernel only): [4*ceil(M/128), N, 2] column (max, min) per 32-row block; D may then be null
P2PB_API int p2pb_gemm_rows_ex(const float* A0, int K0, int lda0, const float* A1, int K1, int lda1, const float* A2, int K2,
                               int lda2, const float* W, const float* bias, const float* bias2, int rows_per_sample, float* D,
                               int ldd, float* stats, float* colmm, int M, int N, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const float* Ap[3] = {A0, A1, A2};
    const int Ks[3] = {K0, K1, K2}, lds[3] = {lda0, lda1, lda2};
    P2PB_CHECK_ARG(M > 0 && N > 0 && N % 16 == 0, "gemm_rows: bad M=%d N=%d (N must be a multiple of 16)", M, N);
    P2PB_CHECK_ARG(D == nullptr || (ldd % 4 == 0 && ldd >= N), "gemm_rows: ldd=%d must be >= N and a multiple of 4", ldd);
    P2PB_CHECK_ARG(D != nullptr || stats != nullptr || colmm != nullptr, "gemm_rows: no output requested");
    P2PB_CHECK_ARG(bias2 == nullptr || rows_per_sample > 0, "gemm_rows: bias2 needs rows_per_sample");
    GemmArgs a = {};
    a.M = M; a.ldd = ldd; a.n_total = N; a.conv = 0; a.dbg = g_gemm_dbg;
    a.bias = bias; a.bias2 = bias2; a.rows_per_sample = rows_per_sample; a.D = D; a.stats = stats;
    CUtensorMap maps[4];
    int ktot = 0, nseg = 0;
    for (int i = 0; i < 3; ++i) {
        if (Ap[i] == nullptr || Ks[i] == 0) break;
        P2PB_CHECK_ARG(Ks[i] % BK == 0 && lds[i] % 4 == 0 && lds[i] >= Ks[i], "gemm_rows: segment %d K=%d lda=%d (K %% 32, lda %% 4)", i, Ks[i], lds[i]);
        P2PB_CHECK_ARG((reinterpret_cast<uintptr_t>(Ap[i]) & 15) == 0, "gemm_rows: segment %d not 16-byte aligned", i);
        cuuint64_t dims[2] = {(cuuint64_t)Ks[i], (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)lds[i] * 4};
        cuuint32_t box[2] = {BK, BM};
        int rc = make_map(&maps[i], Ap[i], 2, dims, str, box);
        if (rc != P2PB_OK) return rc;
        a.seg_chunks[i] = Ks[i] / BK;
        ktot += Ks[i];
        ++nseg;
    }
    P2PB_CHECK_ARG(nseg > 0, "gemm_rows: no A segment");
    for (int i = nseg; i < 3; ++i) maps[i] = maps[0];
    a.nseg = nseg;
    {
        int rc = p2pb_gemm_persist_try(maps, nseg, a.seg_chunks, 0, 0, 0, 0, W, ktot, bias, bias2, rows_per_sample, D, ldd, stats,
                                       colmm, M, N, s);
        if (rc != P2PB_ERR_UNSUPPORTED) return rc;
    }
    P2PB_CHECK_ARG(D != nullptr && colmm == nullptr, "gemm_rows: column max/min and null D need N %% 32 == 0 and an aligned D");
    const int bn = pick_bn(N);
    {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)ktot * 4};
        cuuint32_t box[2] = {BK, (cuuint32_t)bn};
        int rc = make_map(&maps[3], W, 2, dims, str, box);
        if (rc != P2PB_OK) return rc;
    }
    return dispatch_gemm(maps, a, bn, s);
}

P2PB_API int p2pb_gemm_rows(const float* A0, int K0, int lda0, const float* A1, int K1, int lda1, const float* A2, int K2,
                            int lda2, const float* W, const float* bias, const float* bias2, int rows_per_sample, float* D,
                            int ldd, float* stats, int M, int N, void* stream)
{
    return p2pb_gemm_rows_ex(A0, K0, lda0, A1, K1, lda1, A2, K2, lda2, W, bias, bias2, rows_per_sample, D, ldd, stats, nullptr, M, N,
                             stream);
}

// implicit-GEMM 3x3x3 convolution, stride 1, zero padding 1, channels-last:
//   grid [B, r, r, r, Cin] (Cin multiple of 32; x slowest, z fastest = the reference's flat voxel index x*r^2+y*r+z)
//   W    [Cout, 27*Cin] with k = ((kx*3+ky)*3+kz)*Cin + c  (repacked from the reference's [Cout, Cin, 3, 3, 3])
//   D    [B*r^3, ldd] ; stats (optional) [B*r^3/32, Cout, 2]
P2PB_API int p2pb_conv3d_cl(const float* grid, const float* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                            int Cin, int Cout, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    P2PB_CHECK_ARG(B > 0 && Cin % BK == 0 && Cout % 16 == 0, "conv3d: bad B=%d Cin=%d Cout=%d (Cin %% 32, Cout %% 16)", B, Cin, Cout);
    P2PB_CHECK_ARG(r >= 8 && (r & (r - 1)) == 0 && r <= 128, "conv3d: r=%d must be a power of two in [8,128]", r);
    P2PB_CHECK_ARG(ldd % 4 == 0 && ldd >= Cout, "conv3d: bad ldd");
    GemmArgs a = {};
    const int r3 = r * r * r;
    a.M = B * r3; a.ldd = ldd; a.n_total = Cout; a.conv = 1; a.r = r;
    a.bz = r < BM ? r : BM;
    a.by = (BM / a.bz) < r ? (BM / a.bz) : r;
    a.bx = BM / (a.bz * a.by);
    a.cin_chunks = Cin / BK;
    a.tiles_per_sample = r3 / BM;
    a.bias = bias; a.D = D; a.stats = stats;
    CUtensorMap maps[4];
    {
        cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
        cuuint64_t str[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cin * 4 * r, (cuuint64_t)Cin * 4 * r * r, (cuuint64_t)Cin * 4 * r3};
        cuuint32_t box[5] = {BK, (cuuint32_t)a.bz, (cuuint32_t)a.by, (cuuint32_t)a.bx, 1};
        int rc = make_map(&maps[0], grid, 5, dims, str, box);
        if (rc != P2PB_OK) return rc;
        maps[1] = maps[0];
        maps[2] = maps[0];
    }
    {
        int rc = p2pb_gemm_persist_try(maps, 1, nullptr, 1, a.cin_chunks, a.tiles_per_sample, r, W, 27 * Cin, bias, nullptr, 0, D, ldd,
                                       stats, nullptr, a.M, Cout, s);
        if (rc != P2PB_ERR_UNSUPPORTED) return rc;
    }
    const int bn = pick_bn(Cout);
    {
        cuuint64_t dims[2] = {(cuuint64_t)27 * Cin, (cuuint64_t)Cout};
        cuuint64_t str[1] = {(cuuint64_t)27 * Cin * 4};
        cuuint32_t box[2] = {BK, (cuuint32_t)bn};
        int rc = make_map(&maps[3], W, 2, dims, str, box);
        if (rc != P2PB_OK) return rc;
    }
    return dispatch_gemm(maps, a, bn, s);
}
