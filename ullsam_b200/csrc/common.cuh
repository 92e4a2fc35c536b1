// ullsam_b200 — sm_100a device-side building blocks (inline PTX).
//
// Everything the tensor-core kernels share: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences) and the UMMA shared-memory and
// instruction descriptors.  Hand-written for sm_100a; no CUTLASS dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ullsam {

#define ULLSAM_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------
ULLSAM_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

ULLSAM_DEVICE bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
ULLSAM_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

ULLSAM_DEVICE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

ULLSAM_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b64 st;\n\t"
      "mbarrier.arrive.shared::cta.b64 st, [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

ULLSAM_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t"
      ".reg .b64 st;\n\t"
      "mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}

// try_wait with a suspend-time hint: the waiting thread sleeps (NANOSLEEP.SYNCS) instead of re-polling, which
// leaves the issue slots of its SM sub-partition to the compute warps (measured: GEMM lin2 1.31 -> 1.44 PF/s).
#ifndef ULLSAM_TRYWAIT_HINT
#define ULLSAM_TRYWAIT_HINT 20000u
#endif
ULLSAM_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#ifndef ULLSAM_TRYWAIT_NO_HINT
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ULLSAM_TRYWAIT_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}

// Non-blocking probe (mbarrier.test_wait never suspends the thread): has the phase with this parity completed?
ULLSAM_DEVICE bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

#ifndef ULLSAM_MBAR_TIMEOUT_NS   // (compute-sanitizer builds raise it: instrumented kernels run 10-100x slower, tools/build_sanitizer_lib.sh)
#define ULLSAM_MBAR_TIMEOUT_NS 4000000000ull
#endif
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) after ~4 s of wall time instead of a hung
// GPU.  The clock is only read once the wait has already failed 256 times, so the normal path costs nothing.
ULLSAM_DEVICE uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
ULLSAM_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > ULLSAM_MBAR_TIMEOUT_NS) {
        printf("ullsam: mbarrier timeout block=(%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y, threadIdx.x,
               smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// Programmatic dependent launch (kernels launched with launch_pdl(), host_util.h): the grid may be scheduled while its
// predecessor in the stream is still draining; pdl_wait() returns once the predecessor has completed and its memory is
// visible — every thread calls it before its first access to global memory.  pdl_trigger() lets the NEXT grid be scheduled
// as soon as every CTA of this one has got this far.  Both are no-ops for ordinary launches.
ULLSAM_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
ULLSAM_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// generic-proxy writes (st.shared) -> async proxy (UMMA / TMA) visibility
ULLSAM_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------
// Per-thread L2 prefetch of the cache line holding gptr.
ULLSAM_DEVICE void l2_prefetch_line(const void* gptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(gptr)) : "memory");
}
ULLSAM_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

ULLSAM_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA store with an fp32 add performed at the memory side (global[tile] += smem[tile]); bulk-group completion
ULLSAM_DEVICE void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
ULLSAM_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
ULLSAM_DEVICE void tma_store_wait_read() {   // all but the N most recent bulk groups have finished READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
ULLSAM_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

ULLSAM_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                               int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
ULLSAM_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
ULLSAM_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ULLSAM_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ULLSAM_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ULLSAM_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
ULLSAM_DEVICE void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
ULLSAM_DEVICE void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
ULLSAM_DEVICE void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]; issued by ONE thread.
ULLSAM_DEVICE void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrives once every tcgen05 op issued so far by this thread has completed.
// (implies tcgen05.fence::before_thread_sync)
ULLSAM_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets row (lane base + t).
ULLSAM_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Scheduling fence for register values: an empty volatile asm that "rewrites" v, so every later use of v is ordered
// after the volatile asms that precede it (e.g. the tcgen05.ld that prefetches the NEXT buffer).  Without it ptxas
// hoists the arithmetic on the current buffer above the prefetch and the TMEM latency is exposed again.
ULLSAM_DEVICE void pin32(uint32_t (&v)[32]) {
  asm volatile(""
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                 "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31]));
}

ULLSAM_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

ULLSAM_DEVICE uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

ULLSAM_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
      "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

ULLSAM_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

ULLSAM_DEVICE void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}

ULLSAM_DEVICE void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2)
// ------------------------------------------------------------------------------------------
ULLSAM_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ULLSAM_DEVICE void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
ULLSAM_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
ULLSAM_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same without the release: the arrive orders nothing but itself.  (The .release form compiles to MEMBAR.ALL.GPU, which
// stalls the warp until every global store it has issued is acknowledged — ncu: 9 % of the projection GEMM's epilogue.)
// For "this TMEM buffer has been read" signals: the tcgen05.ld data is already in registers (tcgen05.wait::ld) and no
// generic-proxy memory is handed over.
ULLSAM_DEVICE void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
ULLSAM_DEVICE void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
ULLSAM_DEVICE void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
ULLSAM_DEVICE void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ULLSAM_DEVICE void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256 x N x 16 MMA across the CTA pair; issued by ONE thread of the leader CTA
ULLSAM_DEVICE void umma_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once every tcgen05 op issued so far has completed
ULLSAM_DEVICE void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ------------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor")
// ------------------------------------------------------------------------------------------
// Shared-memory operand descriptor, 128-byte swizzle.
//   K-major  : tile = rows x 64 bf16 (128 B per row), 8-row groups 1024 B apart (SBO); LBO field = 1.
//   MN-major : tile = k-rows x 64 bf16 contiguous along MN, 8-k groups 1024 B apart (SBO);
//              LBO = byte distance between 64-wide MN atoms (unused when the operand is 64 wide).
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SW128)
ULLSAM_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
// bits: [4,6) D fmt (1=f32) | [7,10) A fmt (1=bf16) | [10,13) B fmt | 15 A major | 16 B major (0 = K-major)
//       [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// numerics helpers
// ------------------------------------------------------------------------------------------
ULLSAM_DEVICE uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot) and the three-input max (FMNMX3).
ULLSAM_DEVICE uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
ULLSAM_DEVICE void unpk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
ULLSAM_DEVICE uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
ULLSAM_DEVICE uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
ULLSAM_DEVICE uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
ULLSAM_DEVICE float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Exact-erf GELU (nn.GELU() default, common.py:18,23).  erf by Abramowitz & Stegun 7.1.28:
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  |error| <= 3e-7 for z >= 0, odd extension for z < 0
// — one MUFU.RCP and ~15 FMA-pipe instructions instead of erff()'s two divergent branches; the error is far
// below the bf16 rounding of the epilogue's output (and below fp32 accumulation noise of the GEMM itself).
// erf(|x| / sqrt(2)) in [0, 1] (A&S 7.1.28 with the coefficients pre-scaled by 2^(-k/2)).  The 16th power and
// the reciprocal are taken on the MUFU pipe as 2^(-16 log2 p): the epilogue that uses this is bound by the FMA pipe,
// the special-function unit is otherwise idle, and lg2/ex2.approx are accurate to ~2^-22.
ULLSAM_DEVICE float erf_abs_scaled(float x) {
  const float a = fabsf(x);
  float p = fmaf(a, 0.0000430638f * 0.125f, 0.0002765672f * 0.17677669529663687f);
  p = fmaf(p, a, 0.0001520143f * 0.25f);
  p = fmaf(p, a, 0.0092705272f * 0.35355339059327373f);
  p = fmaf(p, a, 0.0422820123f * 0.5f);
  p = fmaf(p, a, 0.0705230784f * 0.70710678118654752f);
  p = fmaf(p, a, 1.0f);  // >= 1
  float lg, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(p));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(lg * -16.0f));
  return 1.0f - r;
}

// Same value with the 16th power taken by one MUFU.RCP and four squarings on the FMA pipe (1 MUFU + 12 FMA-pipe
// instructions instead of 2 + 10): a GELU epilogue can mix the two forms to balance the XU and FMA pipes.
ULLSAM_DEVICE float erf_abs_scaled_rcp(float x) {
  const float a = fabsf(x);
  float p = fmaf(a, 0.0000430638f * 0.125f, 0.0002765672f * 0.17677669529663687f);
  p = fmaf(p, a, 0.0001520143f * 0.25f);
  p = fmaf(p, a, 0.0092705272f * 0.35355339059327373f);
  p = fmaf(p, a, 0.0422820123f * 0.5f);
  p = fmaf(p, a, 0.0705230784f * 0.70710678118654752f);
  p = fmaf(p, a, 1.0f);  // >= 1
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
  r *= r; r *= r; r *= r; r *= r;
  return 1.0f - r;
}
ULLSAM_DEVICE float gelu_erf_rcp(float x) {
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs_scaled_rcp(x), hx);
}

ULLSAM_DEVICE float erf_as(float z) { return copysignf(erf_abs_scaled(z * 1.41421356237309505f), z); }

// gelu(x) = x/2 + |x/2| * erf(|x|/sqrt(2))   (erf is odd, so the sign rides on x/2)
ULLSAM_DEVICE float gelu_erf(float x) {
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs_scaled(x), hx);
}

// gelu(acc + bias) with the bias add folded into the first multiply
ULLSAM_DEVICE float gelu_erf_bias(float acc, float bias) {
  const float x = acc + bias;
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs_scaled(x), hx);
}

// erf-GELU through a fitted logistic form (MLPBlock.lin1's epilogue, which is bound by its instruction count):
//   gelu(x) = x * Phi(x) ~= x / (1 + 2^(x * q(min(x^2, 36)))),   q(t) = k0 + t (k1 + t k2)   (the -2 log2(e) folded into k)
// Minimax fit over the whole real line: |error| <= 2.6e-5 absolute (rel-L2 3.6e-6 on N(0, 2^2) inputs), two orders of
// magnitude below the bf16 rounding of the value it produces; no cancellation for negative x (the logistic form gives the
// small tail values to full relative precision, 1 + tanh would not); q is positive and increasing on [0, 36] and the
// clamp keeps it there, so x -> +inf gives x and x -> -inf gives -0 like the exact function.  3 FMUL + 2 FFMA + FMNMX +
// FADD + 2 MUFU (ex2, rcp) = 9 instructions against 13 for the Abramowitz-Stegun erf above.
ULLSAM_DEVICE float gelu_logistic(float x) {
  const float t = fminf(x * x, 36.0f);
  float q = fmaf(t, 0.00101426302f, -0.106775723f);
  q = fmaf(q, t, -2.30112123f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * q));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

ULLSAM_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ULLSAM_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ullsam
