// Thin inline-PTX layer for sm_100a (B200): mbarrier, bulk/TMA copies, tcgen05 (UMMA) and TMEM.
// Hand-written for this project; syntax cross-checked against the CUDA 12.9 cuda::ptx headers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// One lane of the (fully converged) warp returns true.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Suspend-time hint (ns): the hardware parks the waiting thread until the phase completes or the hint expires, so
// waiting warps do not burn issue slots in a poll loop.
#ifndef CDAN_MBAR_SUSPEND_NS
#define CDAN_MBAR_SUSPEND_NS 20000u
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(CDAN_MBAR_SUSPEND_NS)
      : "memory");
  return ok != 0;
}
// Non-blocking phase test.  mbarrier.try_wait with a suspend-time hint compiles to a loop that SLEEPS FIRST (NANOSLEEP.SYNCS,
// woken by the next barrier event on the SM or the time-out) and only then checks the phase, so waiting on a phase that has
// already completed still costs 300-500 cycles (measured per wait in conv_stream2 and dense_fused traces).  test_wait is a
// single SYNCS.PHASECHK: every wait below tries it once before entering the blocking loop.
__device__ __forceinline__ bool mbar_test_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug must fail loudly (trap -> launch failure), never hang the GPU.
// mbar_wait        : tight poll — for the latency-critical MMA issuer (one warp).
// mbar_wait_relaxed: poll with nanosleep back-off — for producer / epilogue warps, so that their polling does not
//                    steal issue slots from the warps doing real work on the same SM sub-partition.
#ifndef CDAN_MBAR_MAX_POLLS
#define CDAN_MBAR_MAX_POLLS (1u << 22)
#endif
static __device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
  printf("cdan_b200: mbarrier timeout block=(%d,%d) thread=%d smem=0x%x parity=%u\n", blockIdx.x, blockIdx.y,
         threadIdx.x, smem_u32(bar), parity);
#ifdef CDAN_MBAR_DEBUG
  asm volatile("exit;");  // debug builds: let the kernel drain so the printf buffer reaches the host
#else
  __trap();
#endif
}
// The poll loop lives in PTX (try_wait with a suspend-time hint, 4 instructions per wake-up): a waiting warp wakes up
// about ten times per wait on B200, and the compiler-generated C++ loop cost ~16 issue slots per wake-up — a third of
// all instructions the streaming convolution executed.  Still bounded: after CDAN_MBAR_MAX_POLLS failed polls it traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_test_a(smem_u32(bar), parity)) return;
  uint32_t timed_out;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "mov.u32 %0, 0;\n"
      "CDAN_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "@p bra CDAN_WAIT_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 p, cnt, %4;\n\t"
      "@p bra CDAN_WAIT_LOOP;\n\t"
      "mov.u32 %0, 1;\n"
      "CDAN_WAIT_DONE:\n\t"
      "}"
      : "=r"(timed_out)
      : "r"(smem_u32(bar)), "r"(parity), "r"(CDAN_MBAR_SUSPEND_NS), "r"(CDAN_MBAR_MAX_POLLS)
      : "memory");
  if (timed_out) mbar_timeout(bar, parity);
}
// Same with the barrier given as a 32-bit shared-window address: hot single-thread roles (MMA issuers) convert their
// barrier arrays once — the generic->shared conversion (S2R + LEA) otherwise sits on the dependency chain of every wait.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
  if (mbar_test_a(bar_addr, parity)) return;
  uint32_t timed_out;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "mov.u32 %0, 0;\n"
      "CDAN_WAITA_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "@p bra CDAN_WAITA_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 p, cnt, %4;\n\t"
      "@p bra CDAN_WAITA_LOOP;\n\t"
      "mov.u32 %0, 1;\n"
      "CDAN_WAITA_DONE:\n\t"
      "}"
      : "=r"(timed_out)
      : "r"(bar_addr), "r"(parity), "r"(CDAN_MBAR_SUSPEND_NS), "r"(CDAN_MBAR_MAX_POLLS)
      : "memory");
  if (timed_out) mbar_timeout(nullptr, parity | (bar_addr << 1));
}
// Bounded wait with nanosleep back-off between polls.  try_wait returns after a short, implementation-defined time when
// the phase is still pending; a kernel in which twenty warps wait at any moment (dense_fused.cu) otherwise spends most
// of its issue slots in poll loops (ncu: 61 % issue-active, 16 % of all instructions branches) and starves the few warps
// that have work.
__device__ __forceinline__ void mbar_wait_sleep_a(uint32_t bar_addr, uint32_t parity, uint32_t sleep_ns) {
  uint32_t timed_out;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "mov.u32 %0, 0;\n"
      "CDAN_WAITS_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "@p bra CDAN_WAITS_DONE;\n\t"
      "nanosleep.u32 %5;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 p, cnt, %4;\n\t"
      "@p bra CDAN_WAITS_LOOP;\n\t"
      "mov.u32 %0, 1;\n"
      "CDAN_WAITS_DONE:\n\t"
      "}"
      : "=r"(timed_out)
      : "r"(bar_addr), "r"(parity), "r"(CDAN_MBAR_SUSPEND_NS), "r"(CDAN_MBAR_MAX_POLLS), "r"(sleep_ns)
      : "memory");
  if (timed_out) mbar_timeout(nullptr, parity | (bar_addr << 1));
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t sleep_ns = 64) {
  if (mbar_test_a(smem_u32(bar), parity)) return;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if (++polls > (CDAN_MBAR_MAX_POLLS >> 2)) mbar_timeout(bar, parity);
  }
}

// ---------------------------------------------------------------- bulk copies (TMA engine)
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes multiple of 16, 16B-aligned).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 4-D tiled tensor load (coordinates innermost first). OOB elements are zero-filled by the hardware.
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------- TMEM / tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, one CTA. Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, shape M x N (K = 16 per instruction).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4)                       // D format: f32
         | (1u << 7)                     // A format: bf16
         | (1u << 10)                    // B format: bf16
         | (uint32_t(N >> 3) << 17)      // N / 8
         | (uint32_t(M >> 4) << 24);     // M / 16
}
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B apart, groups of
// eight rows are `sbo_bytes` apart (1024 for a dense tile). Version field = 1 on sm_100.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;  // LBO (ignored for swizzled K-major), encoded 16 B
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;  // descriptor version
  d |= uint64_t(base_offset & 7) << 49;
  d |= uint64_t(2) << 61;  // SWIZZLE_128B
  return d;
}

// Same for a 32-byte-swizzled K-major operand (one K=16 step per row: rows 32 B apart, groups of eight rows
// `sbo_bytes` apart, 256 for a dense tile; the 16-byte halves of a row are swapped when address bit 7 is set).
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(6) << 61;  // SWIZZLE_32B
  return d;
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (thread t <- lane 32*(warp%4)+t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}


// Zero this warp's 32 lanes x 16 consecutive TMEM columns (used to re-arm ring accumulators after they were read).
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(
          taddr),
      "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16-byte streaming global load: read-only path, no L1 allocation (each activation byte is used once per CTA).
__device__ __forceinline__ uint4 ldg_stream128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// Explicit shared-window accesses (32-bit shared address): keeps the compiler from falling back to generic LD/ST.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// ---------------------------------------------------------------- CTA pairs (cta_group::2, clusters of two CTAs)
// Two CTAs on the SMs of one TPC execute ONE tcgen05.mma of M = 256: CTA r supplies rows [128r, 128r+128) of A and rows
// [N/2 r, N/2 (r+1)) of B from ITS shared memory (same offsets in both CTAs) and receives its 128 accumulator rows in ITS
// tensor memory.  Only the leader (rank 0) issues MMAs and commits; every CTA streams half of the B operand, which halves
// the shared-memory fill traffic per FLOP (the one-CTA kernel saturates the 128 B/clk shared-memory port: 12 KB of operand
// reads + 4 KB of fills per 128-cycle N = 256 MMA).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // all threads of all CTAs of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {  // arrive on a barrier of any CTA of the cluster
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// shared::cluster address of the LEADER CTA's copy of a barrier, for the cta_group::2 TMA forms: in the shared::cluster window
// of a CTA pair, bit 24 of an address selects the CTA; clearing it names the even (leader) CTA.
__device__ __forceinline__ uint32_t leader_bar(const uint64_t* bar) { return smem_u32(bar) & 0xFEFFFFFFu; }
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {  // one warp (same warp id) in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(uint16_t(3))
               : "memory");
}
// tensor-map loads whose completion bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma2_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3, const uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader_bar(bar))
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst_smem, const void* tmap, int c0, int c1, const uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                   "r"(smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(leader_bar(bar))
               : "memory");
}

// ---------------------------------------------------------------- predicated single-thread instructions
// Executed by every lane of a converged warp with warp-uniform operands; only the lane whose `leader` flag is set
// performs the operation.  Keeping the control flow uniform lets ptxas hold descriptors / barrier addresses in uniform
// registers instead of wrapping each UTCHMMA / UTCBAR / UTMALDG in a vote + R2UR.BROADCAST loop.
__device__ __forceinline__ void umma_bf16_if(uint32_t leader, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_if(uint32_t leader, uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b64 st;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_if(uint32_t leader, void* dst_smem, const void* tmap, int c0, int c1, int c2,
                                               int c3, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}" ::
          "r"(smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "r"(leader)
      : "memory");
}
// two fp32 -> packed bf16x2 (lo in bits 0-15) with ReLU folded into the conversion
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Byte offset of the 16-byte chunk `chunk16` (0..7) of row `row` inside a 128B-swizzled K-major tile whose
// base is 1024-byte aligned: chunk index XOR (row mod 8).
__host__ __device__ constexpr uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

}  // namespace ptx
