// gemm_persist.cu -- persistent tcgen05 GEMM / per-tap implicit-GEMM conv (fp32-stored operands -> kind::tf32, IEEE-half operands ->
// kind::f16) and ALL dense-contraction entry points of the C ABI (p2pb_gemm_rows*, p2pb_conv3d_cl*).  One kernel family covers every
// Conv1d / Conv2d(1x1) / Linear and the r = 8 Conv3d of the hot path (reference: cuDNN/cuBLAS behind models/pvcnn.py:174-192,265-284,
// models/modules.py:337,365-370):  D[M,N] = A[M,K] W[N,K]^T (+ bias[N]) (+ bias2[sample(m),N]), fp32 accumulate / output.
// Built around what the round-1 profiles showed about the first-generation one-tile-per-CTA kernel (removed in round 2):
//
//   * its epilogue (tcgen05.ld -> per-lane 128-byte row stores) was not overlapped with anything and every STG.128
//     touched 32 different cache lines (32 LSU wavefronts per instruction);
//   * a 128 x 256 TF32 tile streams (128 + 256) x 32 x 4 B per 512 MMA cycles = 96 B/cycle/SM, above what L2 can
//     deliver to 148 SMs at once (~42 B/cycle/SM), so the big GEMMs were L2-bound at ~550 TFLOP/s.
//
// Design:
//   * one CTA per SM, static round-robin over (m, n) tiles, n fastest (CTAs that share an A tile run together);
//   * two accumulators in TMEM (2 x BN columns): the epilogue of tile i overlaps the mainloop of tile i+1
//     (tmem_full / tmem_empty mbarriers per accumulator);
//   * optional 2-CTA cluster along M: the two CTAs work on vertically adjacent tiles with the SAME weights, each loads
//     half of every B stage and multicasts it to both (cp.async.bulk.tensor ... .multicast::cluster), the MMA warps
//     release a stage in both CTAs with tcgen05.commit ... .multicast::cluster -> weight traffic from L2 halves;
//   * epilogue: TMEM -> registers (+bias, +per-sample bias) -> 128B-swizzled staging tile in shared memory -> one TMA
//     store per 32 x 32 block (full 128-byte lines, clipped at M by the TMA unit); the GroupNorm partials (sum, sum^2)
//     and, on request, the column max / min are read back column-wise from the staging tile (bank-conflict free);
//   * D may be null (statistics only): the global PointNet's last layer needs only max over points of
//     Swish(GN(x)), which is determined by the per-column max and min of x (Swish is unimodal), so its
//     [B*N, 1024] output is never written.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int PBM = 128;
constexpr int PBK = 32;
constexpr int P_THREADS = 320;   // warp 0: TMA producer, 1: MMA issuer, 2-9: epilogue (two warps per TMEM lane quarter)
constexpr int PA_STAGE = PBM * PBK * 4;  // 16 KiB
constexpr int P_SIDE_KB = 28;            // shared memory left free per SM for a co-resident side-stream CTA (see p_launch)

struct PArgs {
    int M, n_total, n_tiles, m_tiles, total_items;
    int nseg;
    int seg_chunks[3];
    int total_chunks;
    int conv, cin_chunks, tiles_per_sample, r;
    int rows_per_sample;
    int stages;
    int store;
    int dbg;
    int f16;        // operands are IEEE half: a 128-byte chunk row holds 64 K-elements, one MMA (kind::f16) covers K = 16
    const float* bias;
    const float* bias2;
    float* stats;   // [4*m_tiles, n_total, 2]  (sum, sum^2) of each 32-row block
    float* colmm;   // [4*m_tiles, n_total, 2]  (max, min)
};

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// explicit shared-space accesses for the epilogue's staging tile: through a generic pointer (the dynamic shared-memory base is carved
// up at run time) the compiler emits generic ST.E / LD.E, which are slower than STS / LDS (ncu source page of the epilogue)
__device__ __forceinline__ void p_sts128(uint32_t addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float p_lds32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ bool p_elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void p_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void p_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p_mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void p_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void p_tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void p_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                              int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void p_tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint64_t p_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void p_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::f16 (IEEE half operands, fp32 accumulate): same 10-bit operand mantissa as kind::tf32, K = 16 per instruction
__device__ __forceinline__ void p_umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void p_umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p_umma_commit_mc(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void p_tmem_ld32(uint32_t taddr, float* v)
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
// ---- cta_group::2 (CTA pair) forms.  The peer bit (bit 24) of a shared::cluster address selects the CTA inside the pair;
// clearing it addresses the LEADER's (rank 0) copy of a barrier from either CTA.
constexpr uint32_t P_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void p_tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void p_tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2,
                                                   int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// M = 256 across the pair: CTA r supplies A rows [128r, 128r+128) and B rows [N/2 r, N/2 (r+1)) from ITS shared memory (same
// offsets in both CTAs) and receives D rows [128r, 128r+128) x N in ITS tensor memory.  Issued by the leader only.
__device__ __forceinline__ void p_umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void p_umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void p_umma_commit_pair_mc(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void p_mbar_arrive_cluster(uint32_t cluster_bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

__device__ __forceinline__ uint32_t p_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void p_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN>
struct PCfg {
    static constexpr int B_STAGE = BN * PBK * 4;
    static constexpr int NACC = BN >= 256 ? 2 : 4;              // accumulators in TMEM (tiles in flight MMA -> epilogue)
    static constexpr int TM_COLS = NACC * BN;                   // BN in {32,64,128,256} -> 128..512 (powers of two)
    static constexpr int NCB = 1;                               // staging buffers per epilogue warp (4 KiB each)
    static constexpr int C_BYTES = 8 * NCB * 4096;
    static constexpr int STAT_BYTES = 0;
};

template <int BN, int CL>
__global__ void __launch_bounds__(512, 1)      // 320 threads are launched; 512 caps the kernel at 128 registers (see p_launch: P_SIDE_KB)
gemm_persist_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapD, const PArgs a)
{
    P2PB_PDL_SYNC();
    using Cfg = PCfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int stages = a.stages;
    uint8_t* sA = smem;
    constexpr int B_STAGE = CL == 3 ? Cfg::B_STAGE / 2 : Cfg::B_STAGE;   // pair mode: each CTA stores only its half of the weight tile
    uint8_t* sB = sA + (size_t)stages * PA_STAGE;
    uint8_t* sC = sB + (size_t)stages * B_STAGE;
    float* s_stats = reinterpret_cast<float*>(sC + Cfg::C_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_stats) + Cfg::STAT_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + stages;
    constexpr int NACC = Cfg::NACC;
    uint64_t* tfull = bars + 2 * stages;    // [NACC]
    uint64_t* tempty = tfull + NACC;        // [NACC]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty + NACC);

    // CL = 1: independent CTAs; 2: 2-CTA cluster, cta_group::1 MMAs, multicast weights; 3: CTA pair, cta_group::2 MMAs (M = 256)
    constexpr bool PAIR = CL == 3;
    constexpr int NCTA = CL == 1 ? 1 : 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CL > 1 ? p_cluster_ctarank() : 0u;
    const int cluster_id = (int)blockIdx.x / NCTA;
    const int n_clusters = (int)gridDim.x / NCTA;
    if (PAIR) p_cluster_sync();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapD) : "memory");
        for (int s = 0; s < stages; ++s) {
            p_mbar_init(p_smem_u32(&full_bar[s]), 1);
            p_mbar_init(p_smem_u32(&empty_bar[s]), CL == 2 ? 2 : 1);   // multicast mode: one tcgen05.commit arrival per CTA
        }
        for (int h = 0; h < NACC; ++h) {
            p_mbar_init(p_smem_u32(&tfull[h]), 1);
            p_mbar_init(p_smem_u32(&tempty[h]), PAIR ? 16 : 8);   // one arrival per epilogue warp (of both CTAs in pair mode)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p_smem_u32(tmem_ptr_smem)),
                         "r"((uint32_t)Cfg::TM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p_smem_u32(tmem_ptr_smem)),
                         "r"((uint32_t)Cfg::TM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) p_cluster_sync();      // peer barriers are initialised before anyone multicasts into this CTA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (p_elect_one()) {
            int st = 0;
            uint32_t ph = 0;
            const int chk = a.f16 ? 64 : 32;     // K-elements per 128-byte chunk row
            const int sc0 = a.seg_chunks[0], sc1 = a.seg_chunks[1], sc2 = a.seg_chunks[2];
            for (int item = cluster_id; item < a.total_items; item += n_clusters) {
                const int n_tile = item % a.n_tiles;
                int m_tile = (item / a.n_tiles) * NCTA + (int)rank;
                if (m_tile >= a.m_tiles) m_tile = a.m_tiles - 1;   // odd tail of a pair: redo a valid tile, epilogue skips it
                const int m0 = m_tile * PBM, n0 = n_tile * BN;
                int cb = 0, cx = 0, cy = 0;
                if (a.conv) {
                    cb = m_tile / a.tiles_per_sample;
                    const int v0 = (m_tile - cb * a.tiles_per_sample) * PBM;
                    cx = v0 / (a.r * a.r);
                    cy = (v0 / a.r) % a.r;
                }
                int seg = 0, seg_it = 0, tap = 0, kc = 0;
                for (int it = 0; it < a.total_chunks; ++it) {
                    p_mbar_wait(p_smem_u32(&empty_bar[st]), ph ^ 1u);
                    // pair mode: every load of both CTAs signals the LEADER's full barrier, which the leader arms with the
                    // bytes of both CTAs (2 A tiles + the two halves of the weight tile)
                    const uint32_t fb = PAIR ? (p_smem_u32(&full_bar[st]) & P_PEER_MASK) : p_smem_u32(&full_bar[st]);
                    if (!PAIR) p_mbar_expect_tx(fb, PA_STAGE + Cfg::B_STAGE);
                    else if (rank == 0) p_mbar_expect_tx(fb, 2 * PA_STAGE + Cfg::B_STAGE);
                    const uint32_t dstA = p_smem_u32(sA + (size_t)st * PA_STAGE);
                    if (a.conv) {
                        const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
                        if (PAIR) p_tma_load_5d_pair(dstA, &mapA0, fb, kc * chk, dz, cy + dy, cx + dx, cb);
                        else p_tma_load_5d(dstA, &mapA0, fb, kc * chk, dz, cy + dy, cx + dx, cb);
                        if (++kc == a.cin_chunks) {
                            kc = 0;
                            ++tap;
                        }
                    } else {
                        while (seg_it >= (seg == 0 ? sc0 : (seg == 1 ? sc1 : sc2))) {
                            seg_it = 0;
                            ++seg;
                        }
                        const CUtensorMap* mp = seg == 0 ? &mapA0 : (seg == 1 ? &mapA1 : &mapA2);
                        if (PAIR) p_tma_load_2d_pair(dstA, mp, fb, seg_it * chk, m0);
                        else p_tma_load_2d(dstA, mp, fb, seg_it * chk, m0);
                        ++seg_it;
                    }
                    const uint32_t dstB = p_smem_u32(sB + (size_t)st * B_STAGE);
                    if (CL == 1) {
                        p_tma_load_2d(dstB, &mapB, fb, it * chk, n0);
                    } else if (CL == 2) {
                        // this CTA's half of the weight tile goes to both CTAs (same offset), and signals both full barriers
                        p_tma_load_2d_mc(dstB + rank * (Cfg::B_STAGE / 2), &mapB, fb, it * chk, n0 + (int)rank * (BN / 2), (uint16_t)0x3);
                    } else {
                        // pair: this CTA keeps only ITS half of the weight rows (the MMA reads the other half from the peer)
                        p_tma_load_2d_pair(dstB, &mapB, fb, it * chk, n0 + (int)rank * (BN / 2));
                    }
                    if (++st == stages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair mode: the leader CTA only) =====================
        const uint32_t fmt = a.f16 ? 0u : 2u;    // operand format: 0 = f16, 2 = tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * PBM : PBM) >> 4) << 24);
        const bool f16 = a.f16 != 0;
        int st = 0;
        uint32_t ph = 0;
        int li = 0;
        if (!PAIR || rank == 0) {
            for (int item = cluster_id; item < a.total_items; item += n_clusters, ++li) {
                const int h = li % NACC;
                const uint32_t use = (uint32_t)(li / NACC);
                p_mbar_wait(p_smem_u32(&tempty[h]), (use & 1u) ^ 1u);     // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_base + (uint32_t)(h * BN);
                for (int it = 0; it < a.total_chunks; ++it) {
                    p_mbar_wait(p_smem_u32(&full_bar[st]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (p_elect_one()) {
                        const uint64_t ad = p_desc_sw128(p_smem_u32(sA + (size_t)st * PA_STAGE));
                        const uint64_t bd = p_desc_sw128(p_smem_u32(sB + (size_t)st * B_STAGE));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adk = ad + (uint64_t)(k * 2), bdk = bd + (uint64_t)(k * 2);
                            const uint32_t acc = (uint32_t)((it | k) != 0);
                            if (PAIR) {
                                if (f16) p_umma_f16_pair(dcol, adk, bdk, idesc, acc);
                                else p_umma_tf32_pair(dcol, adk, bdk, idesc, acc);
                            } else {
                                if (f16) p_umma_f16(dcol, adk, bdk, idesc, acc);
                                else p_umma_tf32(dcol, adk, bdk, idesc, acc);
                            }
                        }
                        if (CL == 1) p_umma_commit(p_smem_u32(&empty_bar[st]));
                        else if (CL == 2) p_umma_commit_mc(p_smem_u32(&empty_bar[st]), (uint16_t)0x3);
                        else p_umma_commit_pair_mc(p_smem_u32(&empty_bar[st]), (uint16_t)0x3);
                        if (it == a.total_chunks - 1) {
                            if (PAIR) p_umma_commit_pair_mc(p_smem_u32(&tfull[h]), (uint16_t)0x3);
                            else p_umma_commit(p_smem_u32(&tfull[h]));
                        }
                    }
                    __syncwarp();
                    if (++st == stages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9; warp w reads TMEM lanes 32*(w%4) .. +31 (hardware rule) and the
        // two warps of a lane quarter take the even / odd 32-column chunks =====================
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        uint8_t* myC = sC + (size_t)(warp - 2) * Cfg::NCB * 4096;
        const bool want_stats = a.stats != nullptr, want_mm = a.colmm != nullptr;
        int li = 0, cbuf = 0;
        for (int item = cluster_id; item < a.total_items; item += n_clusters, ++li) {
            const int n_tile = item % a.n_tiles;
            const int m_tile = (item / a.n_tiles) * NCTA + (int)rank;
            const bool tile_ok = m_tile < a.m_tiles;
            const int m0 = m_tile * PBM, n0 = n_tile * BN;
            const int h = li % NACC;
            const uint32_t use = (uint32_t)(li / NACC);
            const int row0 = m0 + q * 32;
            const int row = row0 + lane;
            const bool row_ok = tile_ok && row < a.M;
            int nvalid = tile_ok ? a.M - row0 : 0;
            nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
            const float* bias2_row = nullptr;
            if (a.bias2 != nullptr && row_ok) bias2_row = a.bias2 + (size_t)(row / a.rows_per_sample) * a.n_total;
            p_mbar_wait(p_smem_u32(&tfull[h]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (BN / 32 < 2 && chalf == 1) {     // single-chunk tiles: the odd-chunk warps have nothing to read
                if (lane == 0) {
                    if (PAIR) p_mbar_arrive_cluster(p_smem_u32(&tempty[h]) & P_PEER_MASK);
                    else p_mbar_arrive(p_smem_u32(&tempty[h]));
                }
                continue;
            }
#pragma unroll 1
            for (int c = chalf; c < BN / 32; c += 2) {
                float v[32];
                p_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * BN + c * 32), v);
                if (c + 2 >= BN / 32) {
                    // this warp's share of the accumulator is read: hand it back to the MMA warp before the rest of the epilogue
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        if (PAIR) p_mbar_arrive_cluster(p_smem_u32(&tempty[h]) & P_PEER_MASK);   // the leader's MMA warp owns the hand-off
                        else p_mbar_arrive(p_smem_u32(&tempty[h]));
                    }
                }
                const int nb = n0 + c * 32;
                // bias / per-sample bias: 16-byte loads (nb is a multiple of 32 floats, the rows are 16-byte aligned: checked on the host)
                if (a.bias != nullptr) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb) + j4);
                        v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
                    }
                }
                if (bias2_row != nullptr) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias2_row + nb) + j4);
                        v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
                    }
                }
                if (!row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                uint8_t* buf = myC + (size_t)cbuf * 4096;
                // the TMA store that last read this staging buffer must have finished reading it
                if (lane == 0) {
                    if (Cfg::NCB == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                __syncwarp();
                const uint32_t buf_s = p_smem_u32(buf);
                {
                    const uint32_t rowp = buf_s + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        p_sts128(rowp + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (a.store && nvalid > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (a.store && nvalid > 0 && lane == 0) {
                    p_tma_store_2d(&mapD, p_smem_u32(buf), nb, row0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (want_stats || want_mm) {
                    // column `lane` of the 32 x 32 block: element (r, lane) sits at r*128 + (((lane>>2) ^ (r&7))<<4) + (lane&3)*4
                    float s1 = 0.f, s2 = 0.f, mx = -INFINITY, mn = INFINITY;
                    const uint32_t colp = buf_s + (lane & 3) * 4;
                    if (nvalid == 32) {
                        float x[32];
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            x[r] = p_lds32(colp + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4));
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            s1 += x[r];
                            s2 = fmaf(x[r], x[r], s2);
                            mx = fmaxf(mx, x[r]);
                            mn = fminf(mn, x[r]);
                        }
                    } else {
                        for (int r = 0; r < nvalid; ++r) {
                            const float x = p_lds32(colp + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4));
                            s1 += x;
                            s2 = fmaf(x, x, s2);
                            mx = fmaxf(mx, x);
                            mn = fminf(mn, x);
                        }
                    }
                    if (tile_ok) {
                        // 32-row partials go straight to global: [(m_tile*4 + q), n_total, 2], coalesced 256 B per warp
                        const size_t o = ((size_t)(m_tile * 4 + q) * a.n_total + nb + lane) * 2;
                        if (want_stats) *reinterpret_cast<float2*>(a.stats + o) = make_float2(s1, s2);
                        if (want_mm) *reinterpret_cast<float2*>(a.colmm + o) = make_float2(mx, mn);
                    }
                }
                if (Cfg::NCB == 2) cbuf ^= 1;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) p_cluster_sync();      // nobody leaves while the peer may still multicast into / signal this CTA
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled_p)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int p_make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, bool f16 = false)
{
    static PFN_encodeTiled_p enc = nullptr;
    if (enc == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_encodeTiled_p>(p);
    }
    if (enc == nullptr) {
        p2pb_set_error("cuTensorMapEncodeTiled entry point not available");
        return P2PB_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                      const_cast<void*>(base), dims, strides_bytes, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        p2pb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu/%llu, box %u/%u)", (int)rc, rank,
                       (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return P2PB_ERR_CUDA;
    }
    return P2PB_OK;
}

template <int BN, int CL>
int p_launch(const CUtensorMap* maps, PArgs& a, cudaStream_t s)
{
    using Cfg = PCfg<BN>;
    const int fixed = 1024 + Cfg::C_BYTES + Cfg::STAT_BYTES + 512;
    constexpr int b_stage = CL == 3 ? Cfg::B_STAGE / 2 : Cfg::B_STAGE;
    // P_SIDE_KB of shared memory (and, through __launch_bounds__, a quarter of every sub-partition's registers) stay free for a CTA of
    // the geometry stream's kernels (FPS, ball query: 24 KB of staged coordinates, <= 64 registers x 256 threads): a persistent GEMM
    // that owns the whole SM would otherwise wait for those CTAs to finish on 64 of the 148 SMs and, with its static tile schedule,
    // take twice as long (tools/bench_overlap.py: 283 -> 193 us for the three global-PointNet GEMMs next to FPS)
    int stages = ((g_p2pb_smem_budget_kb - P_SIDE_KB) * 1024 - fixed) / (PA_STAGE + b_stage);
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    a.stages = stages;
    const size_t smem = (size_t)fixed + (size_t)stages * (PA_STAGE + b_stage);
    static bool attr_set = false;
    if (!attr_set) {
        P2PB_CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        p2pb_prefer_max_smem((const void*)gemm_persist_kernel<BN, CL>);      // same carve-out as every other kernel: CTAs of the
        attr_set = true;                                                     // side-stream geometry kernels can share the SM
    }
    constexpr int NCTA = CL == 1 ? 1 : 2;
    const int m_items = (a.m_tiles + NCTA - 1) / NCTA;
    a.total_items = m_items * a.n_tiles;
    int n_clusters = p2pb_num_sms() / NCTA;
    if (n_clusters > a.total_items) n_clusters = a.total_items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * NCTA), 1, 1);
    cfg.blockDim = dim3(P_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_p2pb_pdl ? 2 : 1;
    P2PB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_persist_kernel<BN, CL>, maps[0], maps[1], maps[2], maps[3], maps[4], a));
    P2PB_LAUNCH_OK();
    return P2PB_OK;
}

}  // namespace

static int g_p2pb_gemm_mode = 0;   // p2pb_gemm_tune: 4 = 2-CTA multicast clusters instead of CTA pairs; 32 = never form CTA pairs

// development aid for tests/ and tools/: how the persistent GEMM groups CTAs for the big shapes.  0 (default) = cta_group::2 CTA
// pairs (M = 256), 4 = 2-CTA clusters with multicast weights, 32 = independent CTAs only.  Results are identical in every mode.
P2PB_API int p2pb_gemm_tune(int mode)
{
    P2PB_CHECK_ARG(mode == 0 || mode == 4 || mode == 32, "gemm_tune: mode %d (0, 4 or 32)", mode);
    g_p2pb_gemm_mode = mode;
    return P2PB_OK;
}

// mapsA: up to 3 prepared A maps (rows mode: box {32 fp32 | 64 half, 128}; conv mode: 5-D box), W: [N, Ktot].
static int gemm_persist_launch(const CUtensorMap* mapsA, int nseg, const int* seg_chunks, int conv, int cin_chunks, int tiles_per_sample,
                               int r, const void* W, int ktot, const float* bias, const float* bias2, int rows_per_sample, float* D,
                               int ldd, float* stats, float* colmm, int M, int N, bool f16, cudaStream_t s)
{
    P2PB_CHECK_ARG(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
    P2PB_CHECK_ARG(D == nullptr || ((reinterpret_cast<uintptr_t>(D) & 15) == 0 && ldd % 4 == 0), "gemm: D must be 16-byte aligned with ldd %% 4 == 0");
    P2PB_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias2) & 15) == 0,
                   "gemm: bias / per-sample bias must be 16-byte aligned (the epilogue reads them with 16-byte loads)");
    const int bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
    PArgs a = {};
    a.M = M; a.n_total = N; a.n_tiles = N / bn; a.m_tiles = p2pb_cdiv(M, PBM);
    a.nseg = nseg;
    int total = 0;
    for (int i = 0; i < 3; ++i) {
        a.seg_chunks[i] = (seg_chunks != nullptr && i < nseg) ? seg_chunks[i] : 0;
        total += a.seg_chunks[i];
    }
    a.conv = conv; a.cin_chunks = cin_chunks; a.tiles_per_sample = tiles_per_sample; a.r = r;
    a.total_chunks = conv ? 27 * cin_chunks : total;
    a.rows_per_sample = rows_per_sample;
    a.store = D != nullptr;
    a.f16 = f16 ? 1 : 0;
    a.bias = bias; a.bias2 = bias2; a.stats = stats; a.colmm = colmm;
    // 2-CTA cluster with multicast weights when the weight tile dominates the operand traffic and there is enough work
    // CTA pairs (cta_group::2, M = 256): each SM fetches only half of the weight tile -> for the tensor-bound shapes
    const bool pair = g_p2pb_gemm_mode == 0 && bn >= 128 && a.m_tiles >= 2 && (long long)a.m_tiles * a.n_tiles >= 2LL * p2pb_num_sms();
    const bool cl2 = g_p2pb_gemm_mode == 4 && bn >= 128 && a.m_tiles >= 2 && (long long)a.m_tiles * a.n_tiles >= 2LL * p2pb_num_sms();
    CUtensorMap maps[5];
    for (int i = 0; i < 3; ++i) maps[i] = mapsA[i < nseg ? i : 0];
    {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)ktot * (f16 ? 2 : 4)};
        cuuint32_t box[2] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)((cl2 || pair) ? bn / 2 : bn)};
        int rc = p_make_map(&maps[3], W, 2, dims, str, box, f16);
        if (rc != P2PB_OK) return rc;
    }
    if (D != nullptr) {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)ldd * 4};
        cuuint32_t box[2] = {32, 32};
        int rc = p_make_map(&maps[4], D, 2, dims, str, box);
        if (rc != P2PB_OK) return rc;
    } else {
        maps[4] = maps[3];
    }
    if (pair) {
        switch (bn) {
            case 256: return p_launch<256, 3>(maps, a, s);
            case 128: return p_launch<128, 3>(maps, a, s);
        }
    }
    if (cl2) {
        switch (bn) {
            case 256: return p_launch<256, 2>(maps, a, s);
            case 128: return p_launch<128, 2>(maps, a, s);
        }
    }
    switch (bn) {
        case 256: return p_launch<256, 1>(maps, a, s);
        case 128: return p_launch<128, 1>(maps, a, s);
        case 64: return p_launch<64, 1>(maps, a, s);
        case 32: return p_launch<32, 1>(maps, a, s);
    }
    return P2PB_ERR_UNSUPPORTED;
}

// ---- fp32-stored (tf32) operand entry points ------------------------------------------------------------------------------------
// rows-mode GEMM:  D[M, N] = sum_i A_i[M, K_i] * W[N, sum K_i]^T + bias + bias2[m / rows_per_sample]
//   A_i : row-major, row pitch lda_i floats (K_i multiples of 32, lda_i of 4), i < nseg <= 3 (replaces torch.cat)
//   W   : [N, Ktot] row-major, Ktot = sum K_i;  N multiple of 32
//   D   : [M, ldd] (16-byte aligned, ldd % 4 == 0) or null;  stats / colmm (optional): [4*ceil(M/128), N, 2] column (sum, sum of
//         squares) / (max, min) of every 32-row block
// Replaces the cuDNN / cuBLAS calls behind models/pvcnn.py:174-192 (Conv1d / Conv2d 1x1) and models/modules.py:337,365-370 (Linear).
P2PB_API int p2pb_gemm_rows_ex(const float* A0, int K0, int lda0, const float* A1, int K1, int lda1, const float* A2, int K2,
                               int lda2, const float* W, const float* bias, const float* bias2, int rows_per_sample, float* D,
                               int ldd, float* stats, float* colmm, int M, int N, void* stream)
{
    const float* Ap[3] = {A0, A1, A2};
    const int Ks[3] = {K0, K1, K2}, lds[3] = {lda0, lda1, lda2};
    P2PB_CHECK_ARG(M > 0 && N > 0 && N % 32 == 0, "gemm_rows: bad M=%d N=%d (N must be a multiple of 32)", M, N);
    P2PB_CHECK_ARG(D == nullptr || (ldd % 4 == 0 && ldd >= N), "gemm_rows: ldd=%d must be >= N and a multiple of 4", ldd);
    P2PB_CHECK_ARG(D != nullptr || stats != nullptr || colmm != nullptr, "gemm_rows: no output requested");
    P2PB_CHECK_ARG(bias2 == nullptr || rows_per_sample > 0, "gemm_rows: bias2 needs rows_per_sample");
    CUtensorMap maps[3];
    int chunks[3] = {0, 0, 0}, ktot = 0, nseg = 0;
    for (int i = 0; i < 3; ++i) {
        if (Ap[i] == nullptr || Ks[i] == 0) break;
        P2PB_CHECK_ARG(Ks[i] % 32 == 0 && lds[i] % 4 == 0 && lds[i] >= Ks[i], "gemm_rows: segment %d K=%d lda=%d (K %% 32, lda %% 4)", i, Ks[i], lds[i]);
        P2PB_CHECK_ARG((reinterpret_cast<uintptr_t>(Ap[i]) & 15) == 0, "gemm_rows: segment %d not 16-byte aligned", i);
        cuuint64_t dims[2] = {(cuuint64_t)Ks[i], (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)lds[i] * 4};
        cuuint32_t box[2] = {32, PBM};
        int rc = p_make_map(&maps[i], Ap[i], 2, dims, str, box);
        if (rc != P2PB_OK) return rc;
        chunks[i] = Ks[i] / 32;
        ktot += Ks[i];
        ++nseg;
    }
    P2PB_CHECK_ARG(nseg > 0, "gemm_rows: no A segment");
    return gemm_persist_launch(maps, nseg, chunks, 0, 0, 0, 0, W, ktot, bias, bias2, rows_per_sample, D, ldd, stats, colmm, M, N, false,
                               (cudaStream_t)stream);
}

P2PB_API int p2pb_gemm_rows(const float* A0, int K0, int lda0, const float* A1, int K1, int lda1, const float* A2, int K2,
                            int lda2, const float* W, const float* bias, const float* bias2, int rows_per_sample, float* D,
                            int ldd, float* stats, int M, int N, void* stream)
{
    return p2pb_gemm_rows_ex(A0, K0, lda0, A1, K1, lda1, A2, K2, lda2, W, bias, bias2, rows_per_sample, D, ldd, stats, nullptr, M, N,
                             stream);
}

// implicit-GEMM 3x3x3 convolution, stride 1, zero padding 1, channels-last (replaces cuDNN Conv3d, models/pvcnn.py:265-284):
//   grid [B, r, r, r, Cin] (Cin multiple of 32; x slowest, z fastest = the reference's flat voxel index x*r^2+y*r+z)
//   W    [Cout, 27*Cin] with k = ((kx*3+ky)*3+kz)*Cin + c  (repacked from the reference's [Cout, Cin, 3, 3, 3]), Cout % 32 == 0
//   D    [B*r^3, ldd] ; stats (optional) [B*r^3/32, Cout, 2]
// The im2col matrix is never built: per tap a 5-D TMA box load shifted by (dx,dy,dz) brings the 128-voxel x 32-channel operand
// tile, out-of-range voxels are zero-filled by the TMA unit (= the conv's zero padding).
P2PB_API int p2pb_conv3d_cl(const float* grid, const float* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                            int Cin, int Cout, void* stream)
{
    P2PB_CHECK_ARG(B > 0 && Cin % 32 == 0 && Cout % 32 == 0, "conv3d_cl: bad B=%d Cin=%d Cout=%d (Cin %% 32, Cout %% 32)", B, Cin, Cout);
    P2PB_CHECK_ARG(r >= 8 && (r & (r - 1)) == 0 && r <= 128, "conv3d_cl: r=%d must be a power of two in [8,128]", r);
    P2PB_CHECK_ARG(ldd % 4 == 0 && ldd >= Cout && (reinterpret_cast<uintptr_t>(D) & 15) == 0, "conv3d_cl: bad D/ldd");
    const int r3 = r * r * r;
    const int bz = r < PBM ? r : PBM;
    const int by = (PBM / bz) < r ? (PBM / bz) : r;
    const int bx = PBM / (bz * by);
    CUtensorMap map;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cin * 4 * r, (cuuint64_t)Cin * 4 * r * r, (cuuint64_t)Cin * 4 * r3};
    cuuint32_t box[5] = {32, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, 1};
    int rc = p_make_map(&map, grid, 5, dims, str, box);
    if (rc != P2PB_OK) return rc;
    return gemm_persist_launch(&map, 1, nullptr, 1, Cin / 32, r3 / PBM, r, W, 27 * Cin, bias, nullptr, 0, D, ldd, stats, nullptr, B * r3,
                               Cout, false, (cudaStream_t)stream);
}

// ---- IEEE-half operand entry points (A segments and W are __half; K_i multiples of 64; bias / D / stats fp32) -------------
// Same contract as p2pb_gemm_rows_ex / p2pb_conv3d_cl otherwise.  Half keeps the 10-bit mantissa a tf32 operand has inside the
// tensor core; one 128-byte operand row and one MMA carry twice the K of the tf32 path.
P2PB_API int p2pb_gemm_rows_f16(const void* A0, int K0, int lda0, const void* A1, int K1, int lda1, const void* A2, int K2, int lda2,
                                const void* W, const float* bias, const float* bias2, int rows_per_sample, float* D, int ldd,
                                float* stats, float* colmm, int M, int N, void* stream)
{
    const void* Ap[3] = {A0, A1, A2};
    const int Ks[3] = {K0, K1, K2}, lds[3] = {lda0, lda1, lda2};
    P2PB_CHECK_ARG(M > 0 && N > 0 && N % 32 == 0, "gemm_rows_f16: bad M=%d N=%d (N must be a multiple of 32)", M, N);
    P2PB_CHECK_ARG(D == nullptr || (ldd % 4 == 0 && ldd >= N && (reinterpret_cast<uintptr_t>(D) & 15) == 0), "gemm_rows_f16: bad D/ldd");
    P2PB_CHECK_ARG(D != nullptr || stats != nullptr || colmm != nullptr, "gemm_rows_f16: no output requested");
    P2PB_CHECK_ARG(bias2 == nullptr || rows_per_sample > 0, "gemm_rows_f16: bias2 needs rows_per_sample");
    CUtensorMap maps[3];
    int chunks[3] = {0, 0, 0}, ktot = 0, nseg = 0;
    for (int i = 0; i < 3; ++i) {
        if (Ap[i] == nullptr || Ks[i] == 0) break;
        P2PB_CHECK_ARG(Ks[i] % 64 == 0 && lds[i] % 8 == 0 && lds[i] >= Ks[i], "gemm_rows_f16: segment %d K=%d lda=%d (K %% 64, lda %% 8)", i, Ks[i], lds[i]);
        P2PB_CHECK_ARG((reinterpret_cast<uintptr_t>(Ap[i]) & 15) == 0, "gemm_rows_f16: segment %d not 16-byte aligned", i);
        cuuint64_t dims[2] = {(cuuint64_t)Ks[i], (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)lds[i] * 2};
        cuuint32_t box[2] = {64, PBM};
        int rc = p_make_map(&maps[i], Ap[i], 2, dims, str, box, true);
        if (rc != P2PB_OK) return rc;
        chunks[i] = Ks[i] / 64;
        ktot += Ks[i];
        ++nseg;
    }
    P2PB_CHECK_ARG(nseg > 0, "gemm_rows_f16: no A segment");
    return gemm_persist_launch(maps, nseg, chunks, 0, 0, 0, 0, W, ktot, bias, bias2, rows_per_sample, D, ldd, stats, colmm, M, N, true,
                               (cudaStream_t)stream);
}

// grid [B, r, r, r, Cin] __half channels-last (Cin multiple of 64), W [Cout, 27*Cin] __half -> D [B*r^3, ldd] fp32
P2PB_API int p2pb_conv3d_cl_f16(const void* grid, const void* W, const float* bias, float* D, int ldd, float* stats, int B, int r,
                                int Cin, int Cout, void* stream)
{
    P2PB_CHECK_ARG(B > 0 && Cin % 64 == 0 && Cout % 32 == 0, "conv3d_cl_f16: bad B=%d Cin=%d Cout=%d (Cin %% 64, Cout %% 32)", B, Cin, Cout);
    P2PB_CHECK_ARG(r >= 8 && (r & (r - 1)) == 0 && r <= 128, "conv3d_cl_f16: r=%d must be a power of two in [8,128]", r);
    P2PB_CHECK_ARG(ldd % 4 == 0 && ldd >= Cout && (reinterpret_cast<uintptr_t>(D) & 15) == 0, "conv3d_cl_f16: bad D/ldd");
    const int r3 = r * r * r;
    const int bz = r < PBM ? r : PBM;
    const int by = (PBM / bz) < r ? (PBM / bz) : r;
    const int bx = PBM / (bz * by);
    CUtensorMap map;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * 2 * r, (cuuint64_t)Cin * 2 * r * r, (cuuint64_t)Cin * 2 * r3};
    cuuint32_t box[5] = {64, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, 1};
    int rc = p_make_map(&map, grid, 5, dims, str, box, true);
    if (rc != P2PB_OK) return rc;
    return gemm_persist_launch(&map, 1, nullptr, 1, Cin / 64, r3 / PBM, r, W, 27 * Cin, bias, nullptr, 0, D, ldd, stats, nullptr, B * r3,
                               Cout, true, (cudaStream_t)stream);
}

