// eosvr_ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the matcher uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), fences.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eosvr {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id()
{
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// ---- mbarrier -----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA ----------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on an mbarrier (bytes).
// c0 = innermost (K element) coordinate, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                            int32_t c0, int32_t c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1)
        : "memory");
}

// Multicast variant: the box lands at the same CTA-relative offset in every CTA of cta_mask, and each of
// those CTAs gets the complete_tx on the mbarrier at the same CTA-relative offset.
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                               int32_t c0, int32_t c1, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// One full warp allocates ncols TMEM columns; the base address is written to *smem_out.
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_out, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16/bf16 inputs, fp32 accumulate); one thread issues.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns (thread i <- lane i).
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t &v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Same, for a load whose wait is separated from its issue by other work: the destination registers are listed as
// read-write operands so the compiler cannot schedule a use of them above the wait.
__device__ __forceinline__ void tmem_ld_wait_x16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// registers -> TMEM: zero this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_zero_x16(uint32_t taddr)
{
    const uint32_t z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(z)
        : "memory");
}
// registers -> TMEM: the same 32-bit value into this warp's 32 lanes x 16 consecutive columns (lane i writes ITS value)
__device__ __forceinline__ void tmem_st_fill_x16(uint32_t taddr, uint32_t bits)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(bits)
        : "memory");
}
// Same for 16 columns from EIGHT caller-held registers (two tcgen05.st.x8).  A tcgen05.st needs its source values in
// consecutive registers: the single-operand form above makes the compiler copy the value into 15 more registers before
// EVERY store (15 moves per 16-column chunk in the screening epilogue); holding 8 copies for a whole tile costs 7 moves
// per tile.  `opaque_copy` keeps the compiler from folding the copies back into one register.
__device__ __forceinline__ uint32_t opaque_copy(uint32_t x)
{
    uint32_t y;
    asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_fill8_x16(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n\t"
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0 + 8], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
        : "memory");
}
// 20 consecutive columns (one 5-way x 4-segment episode of the screening epilogue): an x16 and an x4 access.
__device__ __forceinline__ void tmem_ld_x20(uint32_t taddr, uint32_t (&v)[20])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%20];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%16, %17, %18, %19}, [%20 + 16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_x20(uint32_t (&v)[20])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st_fill8_x20(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n\t"
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0 + 8], {%1, %2, %3, %4, %5, %6, %7, %8};\n\t"
        "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0 + 16], {%1, %2, %3, %4};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- CTA pair (cluster of 2, cta_group::2) -------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// Arrive on a barrier in another CTA of the cluster.  Default (.release.cta) semantics: the data these
// arrivals order (TMEM reads) is fenced with tcgen05.fence, and a cluster-scope release would cost a
// GPU-wide memory barrier per arrival.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are signalled on the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *m, uint32_t leader_bar,
                                                int32_t c0, int32_t c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *smem_out, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 split over the pair; issued by one thread of the leader.
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (once all prior MMAs of this thread completed) on the barrier at this offset in every CTA of mask.
__device__ __forceinline__ void mma_commit_2sm(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}

// ---- register re-partitioning between warp roles (whole warp executes) -----------------------
template <int N>
__device__ __forceinline__ void reg_release()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_acquire()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- misc ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 128 bytes,
// 8-row groups 1024 bytes apart (the layout a SWIZZLE_128B TMA box of 64 16-bit elements
// produces).  k_byte_off selects the 32-byte K=16 slice inside the 128-byte swizzle atom.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t k_byte_off)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>(((smem_addr + k_byte_off) >> 4) & 0x3FFFu);   // start address
    d |= static_cast<uint64_t>(1) << 16;                                      // LBO (unused with swizzle)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                              // SBO = 1024 B
    d |= static_cast<uint64_t>(1) << 46;                                      // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                                      // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, K-major A and B.
// fmt: 0 = fp16, 1 = bf16.
__host__ __device__ inline uint32_t umma_idesc_f16(int fmt, int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;                              // C format = F32
    d |= static_cast<uint32_t>(fmt) << 7;      // A format
    d |= static_cast<uint32_t>(fmt) << 10;     // B format
    d |= static_cast<uint32_t>(N >> 3) << 17;  // N / 8
    d |= static_cast<uint32_t>(M >> 4) << 24;  // M / 16
    return d;
}

}  // namespace ptx
}  // namespace eosvr
