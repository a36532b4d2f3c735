// conv_halo.cu -- 3x3x3 voxel convolution as a halo-reuse implicit GEMM on tcgen05, for the large-grid / small-channel layers
// (r = 32, 16) where a per-tap TMA kernel (gemm_persist.cu, conv mode) re-reads the activation tile once per tap (27x).
// Replaces cuDNN Conv3d behind models/pvcnn.py:265-284.
//
// Idea: store the conv INPUT as a zero-padded row-major grid  X[b][P^3][Cin]  indexed by the padded-linear voxel
// index q = (x+1)P^2 + (y+1)P + (z+1), P = r+1 (one pad position per row / slab / sample, shared with the neighbour).  In padded-linear space every tap is a CONSTANT row shift
// d = dx*P^2 + dy*P + dz, so for one dx the 9 taps (dy,dz) of a 128-row output tile read nine overlapping 128-row
// windows of ONE contiguous run of W = 128 + 2P + 2 rows.  That run is brought into shared memory ONCE per 128-byte
// channel chunk (32 fp32 / 64 half channels) by a single TMA box load {chunk, W rows} (128B swizzle) and the 9 taps are 9
// UMMA descriptors into the same buffer, start address advanced by whole 128-byte rows (the tensor core applies the 128B
// swizzle to absolute shared-memory address bits, exactly as TMA did when it wrote the window, so no base-offset
// correction is needed -- verified against fp64 convolutions).  Activation traffic drops from 27 to 3*(W/128) tile reads.
//
// Why 128B-swizzled A and not the un-swizzled core-matrix layout (which takes any 16-byte start): measured with
// tools/ubench/mma_rate.cu on B200, one tcgen05.mma (M=128, K=32 bytes) costs max(49, N/2) cycles with a SWIZZLE_128B
// A operand but ~115-125 cycles with an INTERLEAVE (no-swizzle) A operand, independent of the B layout.
//
// Mainloop: weights stream through a ring of SUB-slabs (the 3 dz-taps of one (dx, chunk, dy)) while the activation
// windows of the unit's G tiles stay resident for the whole (dx, chunk) slab:
//   for slab (dx, chunk): for dy: [wait sub-slab] for tile g: [dy == 0: wait window g] 3 taps x 4 MMAs [dy == 2: free window g]
// Separate producer warps feed the two rings so that neither blocks the other; one elected thread issues every MMA with
// 32-bit descriptor arithmetic (at N <= 64 an MMA executes in ~49 cycles, so the issue path is on the critical path);
// accumulators of consecutive units ping-pong between two 256-column halves of TMEM (epilogue overlaps the next mainloop).
//
// Two multipliers on top (template parameters, see the kernel): cta_group::2 CTA pairs (each SM keeps half of the weight
// rows) and IEEE-half operands (kind::f16: same 10-bit mantissa as tf32, twice the channels per byte and per MMA).
// History and measurements of every step: profiles/r01_ncu_conv.md.
//
// Outputs for border positions inside the tile's linear range are computed but masked (not stored, not counted in the
// GroupNorm statistics); the result is written in the dense [B*r^3, Cout] row layout the rest of the engine uses.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int HBM = 128;
constexpr int HALO_THREADS = 352;   // warp 0: window producer, 1: MMA issuer, 2-5: epilogue set 0, 6: weight producer, 7-10: epilogue set 1

struct HaloArgs {
    int B, r, P, P2, P3;      // P = r+1 (shared padding, see p2pb_conv_halo_layout)
    int cin_chunks;           // Cin_p / 32
    int cin_valid;            // channels that can be non-zero (<= Cin_p): trailing K=8 MMAs of the last chunk are skipped
    int cout;
    int W;                    // window rows (odd, >= 128 + 2P + 2)
    int G;                    // tiles per work unit (their accumulators share one TMEM half)
    int halves;               // 2: units ping-pong between two 256-column TMEM halves; 1: one unit owns all 512 columns
    int tiles_per_sample, total_tiles;
    int tile_rows;            // output rows a tile advances by: 128, or 126 in the dz-stacked form (rows 0 and 127 of a tile are halo)
    int q_first, q_last;
    int ldd;
    int a_stages, w_stages;
    const float* X;           // padded row-major input (only used for documentation; loads go through mapX)
    const float* bias;
    float* D;                 // dense rows [B*r^3, ldd]
    float* stats;             // [total_tiles, cout, 2] or null
};

// explicit shared-space accesses for the epilogue (staging tile, boundary rows, statistics): through generic pointers (the dynamic
// shared-memory base is aligned at run time) the compiler emits generic ST.E / LD.E instead of STS / LDS (see gemm_persist.cu)
__device__ __forceinline__ void h_sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 h_lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float h_lds32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void h_sts64(uint32_t addr, float x, float y)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

__device__ __forceinline__ uint32_t h_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one elected lane of a converged warp: ptxas then feeds tcgen05/TMA instructions from uniform registers directly
// (with `if (lane == 0)` it emits an ELECT + R2UR "waterfall" loop around every tcgen05.mma, ~100 cycles each)
__device__ __forceinline__ bool h_elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void h_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void h_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void h_mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void h_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Shared-memory matrix descriptors: K-major SWIZZLE_128B (8-row x 128 B atoms, SBO = 1024 B, version 1), built in the MMA
// loop as a constant high word | (address >> 4).
// A uses the same descriptor with a start address advanced by whole 128-byte rows inside the window (the row shift of
// a tap): the tensor core applies the 128B swizzle to absolute shared-memory address bits, exactly as the TMA unit did
// when it wrote the window, so no base-offset correction is needed (verified against fp64 convolutions on B200;
// setting base_offset = (addr >> 7) & 7 gives wrong results).
__device__ __forceinline__ void h_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::f16 with IEEE half operands: K = 16 per instruction at the cost of a K = 8 tf32 one.  Half has the SAME 10-bit
// mantissa as tf32 (the tensor core drops the low 13 mantissa bits of an fp32 operand in kind::tf32), products are exact
// and accumulation is fp32 in both kinds, so as long as the operands fit half's exponent range (post-GroupNorm/Swish
// activations, O(1) weights) the arithmetic class is unchanged while every operand byte carries twice the channels.
__device__ __forceinline__ void h_umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void h_umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
// ---- cta_group::2 (CTA pair, M = 256 = one tile per CTA) forms; bit 24 of a shared::cluster address selects the CTA in the
// pair, clearing it addresses the leader's barrier (same convention as gemm_persist.cu)
constexpr uint32_t H_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void h_tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void h_umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void h_umma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)0x3)
                 : "memory");
}
__device__ __forceinline__ void h_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void h_umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void h_tmem_ld32(uint32_t taddr, float* v)
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
// PAIR: two CTAs of a cluster form a cta_group::2 pair.  Each MMA has M = 256 = one 128-row tile per CTA against the SAME
// weights; CTA r keeps only weight rows [Cout/2 r, Cout/2 (r+1)) of every tap in its shared memory (the tensor core reads
// the other half from the peer), so per SM the weight bytes fetched from L2, written to and read from shared memory halve.
// At N = Cout <= 64 the shared-memory operand fetch is the binding resource (profiles/r01_ncu_conv.md), at Cout = 128
// the L2 -> SM weight stream was.  The leader CTA issues all MMAs; both CTAs load their own windows / weight halves and
// run their own epilogue.
// F16: operands are IEEE half (X and W); a 128-byte chunk row then holds 64 channels and one MMA covers K = 16.
// STACK (half operands, Cout <= 64): the three dz-taps of one (dx, dy) are ONE MMA with N = 3 Cout.  At N <= 64 an MMA costs
// ~49 cycles whatever N is (the 4 KB A-operand fetch from shared memory, tools/ubench/mma_rate.cu), i.e. 35 % (N = 64) to 67 %
// (N = 32) of the tensor pipe idles; stacking the taps' weight rows -- which already sit contiguously in the sub-slab -- gives
// 3x fewer MMAs at 2x (N = 192) or 1x (N = 96) the cost each.  The three column groups are the PARTIAL outputs
//   P_dz[q'] = sum_{dx,dy} X[q' + dx P^2 + dy P] W(dx,dy,dz)         (A rows WITHOUT the dz shift)
// and the convolution is  out[q] = P_-1[q-1] + P_0[q] + P_+1[q+1]:  a +-1 shift along the TMEM LANES, done in the epilogue with
// two warp shuffles per value (+ a shared-memory hand-over of the two rows at every warp boundary).  Rows 0 and 127 of a tile
// have no complete sum, so tiles advance by 126 rows (1.6 % more tiles).
template <bool PAIR, bool F16, bool STACK>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const HaloArgs a)
{
    P2PB_PDL_SYNC();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NC = PAIR ? 2 : 1;
    constexpr int CHK = F16 ? 64 : 32;                  // channels per 128-byte chunk row
    constexpr int KMMA = F16 ? 16 : 8;                  // channels per MMA
    const int tap_bytes = (a.cout / NC) * 128;          // one tap: [Cout (/2 in pair mode)] rows of one 128B-swizzled chunk
    const int sub_bytes = 3 * tap_bytes;                // sub-slab: the 3 dz-taps of one (dx, chunk, dy)
    const int a_stage_bytes = a.W * 128;                // W rows x 32 channels (one 128-byte swizzle span per row)
    const int a_stage_stride = (a_stage_bytes + 1023) & ~1023;
    uint8_t* sW = smem;
    uint8_t* sA = sW + (size_t)a.w_stages * sub_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)a.a_stages * a_stage_stride);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + a.a_stages;
    uint64_t* w_full = a_empty + a.a_stages;
    uint64_t* w_empty = w_full + a.w_stages;
    uint64_t* tmem_full = w_empty + a.w_stages;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_stats = reinterpret_cast<float*>(tmem_ptr_smem + 2);  // [2 sets][4][cout][2]
    // epilogue staging tiles: 4 warps x (32 rows x 128 B), 1024-byte aligned for the XOR swizzle
    uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_stats + 2 * 4 * a.cout * 2) + 1023) & ~(uintptr_t)1023);
    float* s_edge = reinterpret_cast<float*>(sStage + 8 * 4096);   // STACK: [2 sets][4 warps][2][32] boundary rows of the +-1 lane shift
    const int ncols = STACK ? 3 * a.cout : a.cout;      // accumulator columns of one tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (PAIR) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        h_cluster_sync();
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
        for (int s = 0; s < a.a_stages; ++s) {
            h_mbar_init(h_smem_u32(&a_full[s]), 1);
            h_mbar_init(h_smem_u32(&a_empty[s]), 1);
        }
        for (int s = 0; s < a.w_stages; ++s) {
            h_mbar_init(h_smem_u32(&w_full[s]), 1);
            h_mbar_init(h_smem_u32(&w_empty[s]), 1);
        }
        for (int h = 0; h < 2; ++h) {
            h_mbar_init(h_smem_u32(&tmem_full[h]), 1);
            h_mbar_init(h_smem_u32(&tmem_empty[h]), 8 * NC);   // one arrival per epilogue warp (of both CTAs in pair mode)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_ptr_smem)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_ptr_smem)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) h_cluster_sync();       // the peer's barriers exist before anything is signalled across the pair
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int nslabs = 3 * a.cin_chunks;
    // Balanced static schedule: CTA c owns the contiguous tile range [c*T/grid, (c+1)*T/grid) and walks it in units of up
    // to G tiles, so no CTA gets more than ceil(T/grid) tiles (a unit-granular round-robin loses up to G-1 tile-times).
    // (pair mode: the range belongs to the cluster; slot g of a unit is tile tile0 + 2g + rank, and when the range is odd
    // the last slot of CTA 1 is a dummy that recomputes tile0 and stores nothing)
    const int n_owner = (int)gridDim.x / NC, owner = (int)blockIdx.x / NC;
    const int tile_begin = (int)(((long long)a.total_tiles * owner) / n_owner);
    const int tile_end = (int)(((long long)a.total_tiles * (owner + 1)) / n_owner);
    // Persistent CTA: work units of G tiles, unit u of this CTA accumulates in TMEM half (u & 1) so that the epilogue of
    // one unit overlaps the mainloop of the next (tmem_full / tmem_empty hand-off per half).
    const int half_cols = a.halves == 2 ? 256 : 0;

    if (warp == 0) {
        // ===================== producer A: activation windows (TMA), G per slab =====================
        if (h_elect_one()) {
            int ast = 0;
            uint32_t aph = 0;
            for (int tile0 = tile_begin; tile0 < tile_end; tile0 += NC * a.G) {
                const int ntiles = min(a.G, (tile_end - tile0 + NC - 1) / NC);
                for (int sl = 0; sl < nslabs; ++sl) {
                    const int dx = sl / a.cin_chunks, kc = sl - dx * a.cin_chunks;
                    for (int g = 0; g < ntiles; ++g) {
                        int tile = tile0 + NC * g + (int)rank;
                        if (tile >= tile_end) tile = tile0;       // dummy slot of an odd range (pair mode)
                        const int b = tile / a.tiles_per_sample;
                        const int q0 = a.q_first + (tile - b * a.tiles_per_sample) * a.tile_rows - (STACK ? 1 : 0);
                        const long long qs = (long long)q0 + (long long)(dx - 1) * a.P2 - (a.P + 1);  // window start row (>= 0)
                        h_mbar_wait(h_smem_u32(&a_empty[ast]), aph ^ 1u);
                        if (!PAIR) {
                            const uint32_t fb = h_smem_u32(&a_full[ast]);
                            h_mbar_expect_tx(fb, (uint32_t)a_stage_bytes);
                            h_tma_load_2d(h_smem_u32(sA + (size_t)ast * a_stage_stride), &mapX, fb, kc * CHK,
                                          (int)((long long)b * a.P3 + qs));
                        } else {
                            // both windows of the pair complete on the LEADER's barrier, armed with the bytes of both
                            const uint32_t fb = h_smem_u32(&a_full[ast]) & H_PEER_MASK;
                            if (rank == 0) h_mbar_expect_tx(fb, (uint32_t)(2 * a_stage_bytes));
                            h_tma_load_2d_pair(h_smem_u32(sA + (size_t)ast * a_stage_stride), &mapX, fb, kc * CHK,
                                               (int)((long long)b * a.P3 + qs));
                        }
                        if (++ast == a.a_stages) {
                            ast = 0;
                            aph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 6) {
        // ===================== producer W: weight sub-slabs (3 taps each) =====================
        if (h_elect_one()) {
            int wst = 0;
            uint32_t wph = 0;
            for (int tile0 = tile_begin; tile0 < tile_end; tile0 += NC * a.G) {
                for (int sl = 0; sl < nslabs; ++sl) {
                    const int dx = sl / a.cin_chunks, kc = sl - dx * a.cin_chunks;
                    for (int dy = 0; dy < 3; ++dy) {
                        h_mbar_wait(h_smem_u32(&w_empty[wst]), wph ^ 1u);
                        if (!PAIR) {
                            const uint32_t fb = h_smem_u32(&w_full[wst]);
                            h_mbar_expect_tx(fb, (uint32_t)sub_bytes);
                            for (int dz = 0; dz < 3; ++dz)
                                h_tma_load_2d(h_smem_u32(sW + (size_t)wst * sub_bytes + (size_t)dz * tap_bytes), &mapW, fb,
                                              ((dx * 9 + dy * 3 + dz) * a.cin_chunks + kc) * CHK, 0);
                        } else {
                            // this CTA's half of the output channels of every tap; completes on the leader's barrier
                            const uint32_t fb = h_smem_u32(&w_full[wst]) & H_PEER_MASK;
                            if (rank == 0) h_mbar_expect_tx(fb, (uint32_t)(2 * sub_bytes));
                            for (int dz = 0; dz < 3; ++dz) {
                                // un-stacked: this CTA's half of the rows of every tap.  STACK: the operand is the 3 taps stacked
                                // ([3 Cout] rows, dz-major); this CTA keeps rows [rank * 1.5 Cout, +1.5 Cout) = three half-tap boxes
                                const int hb = STACK ? (int)rank * 3 + dz : 2 * dz + (int)rank;      // half-tap index 0..5
                                h_tma_load_2d_pair(h_smem_u32(sW + (size_t)wst * sub_bytes + (size_t)dz * tap_bytes), &mapW, fb,
                                                   ((dx * 9 + dy * 3 + (hb >> 1)) * a.cin_chunks + kc) * CHK, (hb & 1) * (a.cout / 2));
                            }
                        }
                        if (++wst == a.w_stages) {
                            wst = 0;
                            wph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: ONE thread runs the whole loop =====================
        // At N <= 64 one tcgen05.mma executes in ~49 cycles, so the issue path itself is on the critical path: the loop
        // below keeps per-MMA work to two 32-bit adds (descriptor low words; the high word is a constant) and runs in a
        // single elected thread, so there is no per-step warp re-convergence.
        if ((!PAIR || rank == 0) && h_elect_one()) {
            // instruction descriptor: D = f32; A, B = tf32 (format 2) or f16 (format 0), both K-major; N >> 3 @17, M >> 4 @24
            const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(ncols >> 3) << 17) |
                                   ((uint32_t)((NC * HBM) >> 4) << 24);
            const uint32_t tap_step = (uint32_t)(tap_bytes >> 4);
            const uint32_t sA_lo = (h_smem_u32(sA) & 0x3ffff) >> 4, sW_lo = (h_smem_u32(sW) & 0x3ffff) >> 4;
            const uint32_t a_stride_lo = (uint32_t)(a_stage_stride >> 4), w_stride_lo = (uint32_t)(sub_bytes >> 4);
            constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            int wst = 0, it_unit = 0;
            uint32_t wph = 0;
            int ast0 = 0;          // ring stage of the current slab's first window
            uint32_t aph0 = 0;     // and its phase
            for (int tile0 = tile_begin; tile0 < tile_end; tile0 += NC * a.G, ++it_unit) {
                const int ntiles = min(a.G, (tile_end - tile0 + NC - 1) / NC);
                const int h = a.halves == 2 ? (it_unit & 1) : 0;
                const uint32_t use = (uint32_t)(a.halves == 2 ? (it_unit >> 1) : it_unit);
                h_mbar_wait(h_smem_u32(&tmem_empty[h]), (use & 1u) ^ 1u);     // epilogue(s) have drained this half
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol0 = tmem_base + (uint32_t)(h * half_cols);
                for (int sl = 0; sl < nslabs; ++sl) {
                    const int kc = sl % a.cin_chunks;
                    int nj = (a.cin_valid - kc * CHK + KMMA - 1) / KMMA;     // MMAs that can see a non-zero channel
                    nj = nj > 4 ? 4 : nj;
                    for (int dy = 0; dy < 3; ++dy) {
                        h_mbar_wait(h_smem_u32(&w_full[wst]), wph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_lo = sW_lo + (uint32_t)wst * w_stride_lo;
                        // first tap (dz = -1) of this dy inside a window, in 16-byte descriptor units (128 B per row)
                        const uint32_t row_lo = (uint32_t)(((a.P + 1) + (dy - 1) * a.P - (STACK ? 0 : 1)) * 8);
                        int st = ast0;
                        uint32_t aph = aph0;
                        for (int g = 0; g < ntiles; ++g) {
                            if (dy == 0) {
                                h_mbar_wait(h_smem_u32(&a_full[st]), aph);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            }
                            const uint32_t a_lo = sA_lo + (uint32_t)st * a_stride_lo + row_lo;
                            const uint32_t dcol = dcol0 + (uint32_t)(g * ncols);
                            const uint32_t first = (uint32_t)((sl | dy) != 0);
                            if (STACK) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (j < nj) {
                                        const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(j * 2));
                                        const uint64_t bd = DESC_HI | (uint64_t)(b_lo + (uint32_t)(j * 2));
                                        const uint32_t acc = j != 0 ? 1u : first;
                                        if (PAIR) h_umma_f16_pair(dcol, ad, bd, idesc, acc);
                                        else h_umma_f16(dcol, ad, bd, idesc, acc);
                                    }
                                }
                            } else
#pragma unroll
                            for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (j < nj) {
                                        const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(dz * 8 + j * 2));
                                        const uint64_t bd = DESC_HI | (uint64_t)(b_lo + (uint32_t)dz * tap_step + (uint32_t)(j * 2));
                                        const uint32_t acc = (dz | j) != 0 ? 1u : first;
                                        if (PAIR && F16) h_umma_f16_pair(dcol, ad, bd, idesc, acc);
                                        else if (PAIR) h_umma_tf32_pair(dcol, ad, bd, idesc, acc);
                                        else if (F16) h_umma_f16(dcol, ad, bd, idesc, acc);
                                        else h_umma_tf32(dcol, ad, bd, idesc, acc);
                                    }
                                }
                            }
                            if (dy == 2) {
                                if (PAIR) h_umma_commit_pair(h_smem_u32(&a_empty[st]));
                                else h_umma_commit(h_smem_u32(&a_empty[st]));
                            }
                            if (++st == a.a_stages) {
                                st = 0;
                                aph ^= 1u;
                            }
                        }
                        if (PAIR) h_umma_commit_pair(h_smem_u32(&w_empty[wst]));
                        else h_umma_commit(h_smem_u32(&w_empty[wst]));
                        if (++wst == a.w_stages) {
                            wst = 0;
                            wph ^= 1u;
                        }
                        if (dy == 2) {      // the slab's windows are consumed: the next slab starts after them in the ring
                            ast0 = st;
                            aph0 = aph;
                        }
                    }
                }
                if (PAIR) h_umma_commit_pair(h_smem_u32(&tmem_full[h]));
                else h_umma_commit(h_smem_u32(&tmem_full[h]));
            }
        }
    } else {
        // ===================== epilogue =====================
        // two sets of 4 warps (warp w reads TMEM lanes 32*(w%4)..+31); set 0 takes the even tiles of a unit, set 1 the odd ones
        const int qd = warp & 3;
        const int eset = warp >= 7 ? 1 : 0;
        float* s_stats_set = s_stats + (size_t)eset * 4 * a.cout * 2;
        const int r = a.r, r3 = r * r * r;
        int it_unit = 0;
        int epar = 0;
        for (int tile0 = tile_begin; tile0 < tile_end; tile0 += NC * a.G, ++it_unit) {
            const int ntiles = min(a.G, (tile_end - tile0 + NC - 1) / NC);
            const int h = a.halves == 2 ? (it_unit & 1) : 0;
            const uint32_t use = (uint32_t)(a.halves == 2 ? (it_unit >> 1) : it_unit);
            h_mbar_wait(h_smem_u32(&tmem_full[h]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // two epilogue warp sets alternate over the tiles of a unit -- over the UNITS when a unit is a single tile (G = 1)
            for (int g = (a.G == 1 ? ((it_unit & 1) == eset ? 0 : 1) : eset); g < ntiles; g += 2) {
                const int tile = tile0 + NC * g + (int)rank;
                const bool tile_ok = tile < tile_end;         // false only for the dummy slot of an odd range (pair mode)
                const int b = tile / a.tiles_per_sample;
                const int it = qd * 32 + lane;                // row inside the tile = TMEM lane
                const int q = a.q_first + (tile - b * a.tiles_per_sample) * a.tile_rows - (STACK ? 1 : 0) + it;
                const int x = q / a.P2, rem = q - x * a.P2, y = rem / a.P, z = rem - y * a.P;
                const bool ok = tile_ok && q <= a.q_last && x >= 1 && x <= r && y >= 1 && y <= r && z >= 1 && z <= r &&
                                (!STACK || (it >= 1 && it <= HBM - 2));
                const size_t v = (size_t)b * r3 + (size_t)(x - 1) * r * r + (size_t)(y - 1) * r + (size_t)(z - 1);
                // The valid rows of a tile are CONSECUTIVE dense voxel rows (v enumerates the interior voxels in the same order
                // as q, pads skipped), so the warp's valid rows are compacted into a 128B-swizzled staging tile and leave as
                // full 128-byte row segments (4 rows per store instruction instead of 32 different cache lines), and the
                // GroupNorm partials are read back column-wise from the staging tile instead of two 31-shuffle butterflies.
                const unsigned vmask = __ballot_sync(0xffffffffu, ok);
                const int nv = __popc(vmask);
                const int pos = __popc(vmask & ((1u << lane) - 1u));
                const unsigned long long v0 = __shfl_sync(0xffffffffu, (unsigned long long)v, vmask ? (__ffs(vmask) - 1) : 0);
                float* dbase = a.D + (size_t)v0 * a.ldd;
                uint8_t* stg = sStage + (size_t)(eset * 4 + qd) * 4096;
                for (int c = 0; c < a.cout / 32; ++c) {
                    float vv[32];
                    const uint32_t tcol = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(h * half_cols + g * ncols + c * 32);
                    if (STACK) {
                        // out[row] = P_-1[row-1] + P_0[row] + P_+1[row+1]: column groups 0 / 1 / 2 of the tile, shifted along the lanes
                        float vm[32], vp[32];
                        h_tmem_ld32(tcol, vm);
                        h_tmem_ld32(tcol + (uint32_t)a.cout, vv);
                        h_tmem_ld32(tcol + (uint32_t)(2 * a.cout), vp);
                        float* eb = s_edge + (size_t)((eset * 2 + epar) * 4) * 64;     // [warp][{last row of P_-1, first row of P_+1}][32]
                        const uint32_t eb_s = h_smem_u32(eb);
                        epar ^= 1;         // double-buffered: the barrier of use k+1 separates the reads of use k from the writes of use k+2
                        if (lane == 31) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                h_sts128(eb_s + (qd * 64 + 4 * j) * 4, make_float4(vm[4 * j], vm[4 * j + 1], vm[4 * j + 2], vm[4 * j + 3]));
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                h_sts128(eb_s + (qd * 64 + 32 + 4 * j) * 4, make_float4(vp[4 * j], vp[4 * j + 1], vp[4 * j + 2], vp[4 * j + 3]));
                        }
                        if (eset == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                        else asm volatile("bar.sync 2, 128;" ::: "memory");
                        // branch-free: every lane reads the neighbour warps' boundary rows (broadcast), lanes 0 / 31 select them
                        // (rows 0 and 127 of the tile have no neighbour: they are masked halo rows anyway)
                        const uint32_t pu = eb_s + ((qd > 0 ? qd - 1 : qd) * 64) * 4;
                        const uint32_t pd = eb_s + ((qd < 3 ? qd + 1 : qd) * 64 + 32) * 4;
                        const bool l0 = lane == 0, l31 = lane == 31;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 u = h_lds128(pu + 16 * j4), d = h_lds128(pd + 16 * j4);
                            const float ua[4] = {u.x, u.y, u.z, u.w}, da[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int j = 4 * j4 + k;
                                const float up = __shfl_up_sync(0xffffffffu, vm[j], 1);
                                const float dn = __shfl_down_sync(0xffffffffu, vp[j], 1);
                                vv[j] += (l0 ? ua[k] : up) + (l31 ? da[k] : dn);
                            }
                        }
                    } else {
                        h_tmem_ld32(tcol, vv);
                    }
                    __syncwarp();              // the previous chunk's readers are done with the staging tile
                    const uint32_t stg_s = h_smem_u32(stg);
                    if (ok) {
                        const uint32_t rowp = stg_s + pos * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 t = make_float4(vv[4 * j], vv[4 * j + 1], vv[4 * j + 2], vv[4 * j + 3]);
                            if (a.bias != nullptr) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + c * 32 + 4 * j));
                                t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
                            }
                            h_sts128(rowp + ((j ^ (pos & 7)) << 4), t);
                        }
                    }
                    __syncwarp();
                    // rows [0, nv) of the staging tile -> nv consecutive dense rows, 8 lanes per 128-byte row segment
                    for (int i = lane; i < nv * 8; i += 32) {
                        const int rr = i >> 3, j = i & 7;
                        const float4 t = h_lds128(stg_s + rr * 128 + ((j ^ (rr & 7)) << 4));
                        *reinterpret_cast<float4*>(dbase + (size_t)rr * a.ldd + c * 32 + j * 4) = t;
                    }
                    if (a.stats != nullptr) {
                        // column `lane`: element (rr, lane) sits at rr*128 + (((lane>>2) ^ (rr&7))<<4) + (lane&3)*4 (conflict-free)
                        float s1 = 0.f, s2 = 0.f;
                        const uint32_t colp = stg_s + (lane & 3) * 4;
                        for (int rr = 0; rr < nv; ++rr) {
                            const float xx = h_lds32(colp + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4));
                            s1 += xx;
                            s2 = fmaf(xx, xx, s2);
                        }
                        h_sts64(h_smem_u32(s_stats_set) + (((qd * a.cout) + c * 32 + lane) * 2) * 4, s1, s2);
                    }
                }
                if (a.stats != nullptr) {
                    if (eset == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                    else asm volatile("bar.sync 2, 128;" ::: "memory");
                    const int t = qd * 32 + lane;
                    for (int n = t; tile_ok && n < a.cout; n += 128) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            s1 += h_lds32(h_smem_u32(s_stats_set) + (((w * a.cout) + n) * 2 + 0) * 4);
                            s2 += h_lds32(h_smem_u32(s_stats_set) + (((w * a.cout) + n) * 2 + 1) * 4);
                        }
                        float* o = a.stats + ((size_t)tile * a.cout + n) * 2;
                        o[0] = s1;
                        o[1] = s2;
                    }
                    if (eset == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                    else asm volatile("bar.sync 2, 128;" ::: "memory");
                }
            }
            // this half of TMEM may be overwritten by the MMA warp again
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) {
                if (PAIR) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(h_smem_u32(&tmem_empty[h]) & H_PEER_MASK) : "memory");
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(h_smem_u32(&tmem_empty[h])) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) h_cluster_sync();       // nobody leaves while the peer may still signal / read this CTA
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*PFN_encodeTiled_h)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// Geometry of the padded row-major layout for resolution r: rows per sample, slack rows needed after the last sample.
// SHARED padding: P = r + 1.  Voxel (x,y,z) sits at row q = (x+1)P^2 + (y+1)P + (z+1); position 0 of every z-row, y-row 0 of
// every x-slab and x-slab 0 of every sample are zero and never written.  The "right-hand" pad of a z-row is position 0 of
// the NEXT z-row (z = r+1 = P carries over into y+1), that of a y-row block is y-row 0 of the next slab, that of the last
// slab is slab 0 of the next sample (or the slack rows after the last sample) -- every out-of-range neighbour of an
// interior voxel still lands on a zero row, and taps are still constant row shifts dx*P^2 + dy*P + dz.  Compared with
// padding both sides (P = r + 2) the rows to sweep shrink from (r+2)^3 to (r+1)^3: 17 % fewer tiles at r = 16, 9 % at r = 32.
// dz-stacked form (see the kernel): half operands, Cout = 32, or Cout = 64 with Cin >= 192.  Measured on B200 (64 patches,
// tools/bench_halo.py, stacked vs un-stacked): 64(35)->32 @32^3 234 vs 269 us, 32->32 @32^3 225 vs 253, 192->64 @16^3 129 vs 148;
// but 64->64 @32^3 494 vs 397 and 64->64 @16^3 70 vs 56: with one 192-column tile per TMEM half (G = 1) every tile streams all 27
// taps' weights again, and the L2 -> SM fill (Little's law: ~100 KB in flight per SM at ~1.5 us) cannot feed 185 KB per 1.9 us tile.
static inline bool halo_stacked(bool f16, int Cin, int Cout) { return f16 && (Cout == 32 || (Cout == 64 && Cin >= 192)); }
static int g_halo_stack = 1;      // p2pb_conv_halo_tune(.., .., G): G in [100, 200) = stacking OFF, [200, 300) = stacking forced ON (A/B timing)

// tiles per sample of the kernel variant that will run for (r, Cout, operand type): the stats buffer of p2pb_conv3d_halo* has
// B * tiles rows
P2PB_API int p2pb_conv_halo_tiles(int r, int Cin, int Cout, int f16)
{
    const int P = r + 1, P2 = P * P;
    const int q_first = P2 + P + 1, q_last = r * P2 + r * P + r;
    const int rows = (g_halo_stack == 2 || (g_halo_stack && halo_stacked(f16 != 0, Cin, Cout))) && f16 && (Cout == 32 || Cout == 64) ? HBM - 2 : HBM;
    return (q_last - q_first + 1 + rows - 1) / rows;
}

P2PB_API int p2pb_conv_halo_layout(int r, int* P3_out, int* slack_rows_out, int* tiles_per_sample_out)
{
    const int P = r + 1, P2 = P * P, P3 = P2 * P;
    const int q_first = P2 + P + 1, q_last = r * P2 + r * P + r;
    if (P3_out) *P3_out = P3;
    if (slack_rows_out) *slack_rows_out = P2 + 128 + 2 * P + 8;
    if (tiles_per_sample_out) *tiles_per_sample_out = (q_last - q_first + 1 + HBM - 1) / HBM;
    return P2PB_OK;
}

// X: zero-padded row-major grid [B*(r+1)^3 + slack rows, Cin] (p2pb_conv_halo_layout); W: [Cout, 27*Cin] (k = tap*Cin + c);
// D: dense rows [B*r^3, ldd]; stats (optional): [B*tiles_per_sample, Cout, 2]
// development aid (tools/bench_conv.py): override the pipeline shape; 0 = automatic
static int g_halo_w_stages = 0, g_halo_a_stages = 0, g_halo_G = 0;
static int g_halo_pair = 1;   // 1: cta_group::2 CTA pairs when there is enough work; 0: independent CTAs (p2pb_conv_halo_tune G < 0)
P2PB_API int p2pb_conv_halo_tune(int w_stages, int a_stages, int G)
{
    g_halo_w_stages = w_stages; g_halo_a_stages = a_stages;
    g_halo_stack = 1;
    if (G >= 200) {                    // 200 + G: dz-stacking forced on for every half-operand Cout in {32, 64} shape
        g_halo_stack = 2;
        G -= 200;
    } else if (G >= 100) {             // 100 + G: dz-stacking off (the round-1 kernel), G tiles per unit (0 = automatic)
        g_halo_stack = 0;
        G -= 100;
    }
    g_halo_pair = G < 0 ? 0 : 1;       // a negative G selects the un-paired kernel with |G| (0 = automatic) tiles per unit
    g_halo_G = G < 0 ? (G == -1 ? 0 : -G) : G;
    return P2PB_OK;
}

// cin_valid <= Cin: channels that can be non-zero (the rest is zero padding in X and W): the K=8 MMAs that would only
// multiply padding are skipped (SA0's first conv: 35 real channels in a 64-wide layout -> 5 of 8 MMAs per tap)
static int conv3d_halo_impl(const void* X, const void* W, const float* bias, float* D, int ldd, float* stats, int B, int r, int Cin,
                           int cin_valid, int Cout, bool f16, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const int chk = f16 ? 64 : 32, esz = f16 ? 2 : 4;
    P2PB_CHECK_ARG(B > 0 && Cin % chk == 0 && Cout % 32 == 0 && Cout <= 128, "conv3d_halo: Cin=%d Cout=%d (Cin %% %d, Cout %% 32, Cout<=128)", Cin, Cout, chk);
    P2PB_CHECK_ARG(cin_valid > 0 && cin_valid <= Cin && cin_valid > Cin - chk, "conv3d_halo: cin_valid=%d must lie in the last %d-channel chunk of Cin=%d", cin_valid, chk, Cin);
    P2PB_CHECK_ARG(r >= 8 && r <= 62, "conv3d_halo: r=%d out of range (TMA box rows 128+2(r+2)+2 <= 256)", r);
    P2PB_CHECK_ARG(ldd % 4 == 0 && ldd >= Cout, "conv3d_halo: bad ldd");
    HaloArgs a = {};
    a.B = B; a.r = r; a.P = r + 1; a.P2 = a.P * a.P; a.P3 = a.P2 * a.P;
    a.cin_chunks = Cin / chk; a.cin_valid = cin_valid; a.cout = Cout;
    a.W = 128 + 2 * a.P + 2;
    a.q_first = a.P2 + a.P + 1;
    a.q_last = r * a.P2 + r * a.P + r;
    const bool stack = f16 && (Cout == 32 || Cout == 64) && (g_halo_stack == 2 || (g_halo_stack && halo_stacked(f16, Cin, Cout)));
    a.tile_rows = stack ? HBM - 2 : HBM;
    a.tiles_per_sample = (a.q_last - a.q_first + 1 + a.tile_rows - 1) / a.tile_rows;
    a.total_tiles = B * a.tiles_per_sample;
    const int ncols = stack ? 3 * Cout : Cout;
    const bool pair_hint = g_halo_pair && a.total_tiles >= 2 * p2pb_num_sms();
    // (half operands in CTA pairs at Cout = 64: G = 2 measured faster than 4 -- 459 vs 526 us at 64 -> 64 @ 32^3 -- the
    // weight stream is already a quarter of the tf32 single-CTA one and shorter units balance better)
    // G tiles share every weight sub-slab; their accumulators fit one 256-column half of TMEM (G = 4 at Cout <= 64, 2 at
    // Cout = 128) and units ping-pong between the halves, so the epilogue overlaps the next unit's mainloop (measured at
    // Cout = 128: G = 4 without overlap 467 us, G = 2 with overlap 390 us)
    a.G = (Cout <= 32 || (Cout <= 64 && !(f16 && pair_hint))) ? 4 : 2;
    if (stack) a.G = 256 / ncols;          // 2 tiles (N = 96) or 1 tile (N = 192) per 256-column TMEM half, units ping-pong
    if (g_halo_G > 0 && g_halo_G * ncols <= 512) a.G = g_halo_G;
    a.halves = a.G * ncols <= 256 ? 2 : 1;
    a.ldd = ldd;
    a.X = reinterpret_cast<const float*>(X); a.bias = bias; a.D = D; a.stats = stats;
    const int n_sms = p2pb_num_sms();
    const bool pair = g_halo_pair && a.total_tiles >= 2 * n_sms && n_sms >= 2;
    const int sub_bytes = 3 * (pair ? Cout / 2 : Cout) * 128;     // per CTA: pair mode keeps half of the output channels
    const int a_stage_stride = ((a.W * 128) + 1023) & ~1023;
    const int stage_bytes = 8 * 4096 + 1024 + 4096;      // epilogue staging tiles (+ alignment) + boundary rows of the stacked form
    const int budget = (g_p2pb_smem_budget_kb - 2) * 1024 - 1024 - 256 - 8 * Cout * 2 * 4 - stage_bytes;
    a.w_stages = (3 * sub_bytes + (a.G + 1) * a_stage_stride <= budget) ? 3 : 2;
    a.a_stages = (budget - a.w_stages * sub_bytes) / a_stage_stride;
    int a_cap = 2 * a.G;
    if (stack) {
        // the stacked mainloop consumes a weight sub-slab in G * 4 MMAs (0.2-0.4 us) and a window in 12 (0.3-0.6 us), 3x faster
        // than the un-stacked one, while a TMA round trip stays ~1 us: keep ~64 KB of weights and 4-5 windows in flight
        a_cap = 2 * a.G > 4 ? 2 * a.G : 4;
        a.w_stages = 8;
        while (a.w_stages > 3 && a.w_stages * sub_bytes + a_cap * a_stage_stride > budget) --a.w_stages;
        a.a_stages = (budget - a.w_stages * sub_bytes) / a_stage_stride;
        if (a.a_stages > a_cap + 1) a.a_stages = a_cap + 1;
    }
    if (g_halo_w_stages >= 2) {
        a.w_stages = g_halo_w_stages;
        a.a_stages = (budget - a.w_stages * sub_bytes) / a_stage_stride;
    }
    if (!stack && a.a_stages > a_cap) a.a_stages = a_cap;
    if (g_halo_a_stages > 0 && g_halo_a_stages < a.a_stages) a.a_stages = g_halo_a_stages;
    P2PB_CHECK_ARG(a.a_stages >= a.G + 1, "conv3d_halo: shared memory budget exceeded (Cout=%d r=%d)", Cout, r);
    const size_t smem = 1024 + (size_t)a.w_stages * sub_bytes + (size_t)a.a_stages * a_stage_stride + 256 + (size_t)8 * Cout * 2 * 4 + stage_bytes;
    CUtensorMap mapW, mapX;
    {
        static PFN_encodeTiled_h enc = nullptr;
        if (enc == nullptr) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
                qres == cudaDriverEntryPointSuccess)
                enc = reinterpret_cast<PFN_encodeTiled_h>(p);
        }
        P2PB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
        const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        cuuint64_t dims[2] = {(cuuint64_t)27 * Cin, (cuuint64_t)Cout};
        cuuint64_t str[1] = {(cuuint64_t)27 * Cin * esz};
        cuuint32_t box[2] = {(cuuint32_t)chk, (cuuint32_t)(pair ? Cout / 2 : Cout)};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = enc(&mapW, dt, 2, const_cast<void*>(W), dims, str, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            p2pb_set_error("conv3d_halo: cuTensorMapEncodeTiled(W) failed (%d)", (int)rc);
            return P2PB_ERR_CUDA;
        }
        int slack = 0;
        p2pb_conv_halo_layout(r, nullptr, &slack, nullptr);
        cuuint64_t xdims[2] = {(cuuint64_t)Cin, (cuuint64_t)B * a.P3 + (cuuint64_t)slack};
        cuuint64_t xstr[1] = {(cuuint64_t)Cin * esz};
        cuuint32_t xbox[2] = {(cuuint32_t)chk, (cuuint32_t)a.W};
        rc = enc(&mapX, dt, 2, const_cast<void*>(X), xdims, xstr, xbox, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            p2pb_set_error("conv3d_halo: cuTensorMapEncodeTiled(X) failed (%d)", (int)rc);
            return P2PB_ERR_CUDA;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        P2PB_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_sms & ~1), 1, 1);
        cfg.blockDim = dim3(HALO_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_p2pb_pdl ? 2 : 1;
        if (stack) P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<true, true, true>, mapW, mapX, a));
        else if (f16) P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<true, true, false>, mapW, mapX, a));
        else P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<true, false, false>, mapW, mapX, a));
    } else {
        int grid = n_sms;
        if (grid > a.total_tiles) grid = a.total_tiles;
        if (stack) (void)p2pb_launch(conv_halo_kernel<false, true, true>, dim3(grid), dim3(HALO_THREADS), (size_t)(smem), s, mapW, mapX, a);
        else if (f16) (void)p2pb_launch(conv_halo_kernel<false, true, false>, dim3(grid), dim3(HALO_THREADS), (size_t)(smem), s, mapW, mapX, a);
        else (void)p2pb_launch(conv_halo_kernel<false, false, false>, dim3(grid), dim3(HALO_THREADS), (size_t)(smem), s, mapW, mapX, a);
    }
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

P2PB_API int p2pb_conv3d_halo_ex(const float* X, const float* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                                 int Cin, int cin_valid, int Cout, void* stream)
{
    return conv3d_halo_impl(X, W, bias, D, ldd, stats, B, r, Cin, cin_valid, Cout, false, stream);
}

// IEEE-half operands (X [rows, Cin] and W [Cout, 27*Cin] as __half, Cin a multiple of 64), fp32 accumulate / bias / output:
// same 10-bit operand mantissa as the tf32 path, twice the channels per operand byte and per MMA
P2PB_API int p2pb_conv3d_halo_f16(const void* X, const void* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                                  int Cin, int cin_valid, int Cout, void* stream)
{
    return conv3d_halo_impl(X, W, bias, D, ldd, stats, B, r, Cin, cin_valid, Cout, true, stream);
}

P2PB_API int p2pb_conv3d_halo(const float* X, const float* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                              int Cin, int Cout, void* stream)
{
    return p2pb_conv3d_halo_ex(X, W, bias, D, ldd, stats, B, r, Cin, Cin, Cout, stream);
}
