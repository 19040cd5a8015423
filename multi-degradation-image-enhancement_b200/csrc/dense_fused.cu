// Fused final dense block for sm_100a — ONE kernel for the tail of the decoder (reference models/cdan.py:153-157 with
// DenseBlock :22-53):
//
//     F0 = bilinear_x2(relu(bn4(convT4(.)))) + x          (3 channels, full resolution)
//     G1..G4 = the four growth layers  conv3x3(relu(bn_l(cat(F0, G1..G_l))))  -> 16 channels each
//     y = sigmoid(conv1x1(relu(bn_t(cat(F0, G1..G4)))))   (3 channels, fp32 NCHW)
//
// Layer by layer this block moved 41 GB of HBM traffic per 32 x 1080p step for ~2.9 GB of compulsory I/O (round-1 ncu):
// every layer re-read the growing concat.  Here the concat never leaves the SM: a CTA walks a 128-pixel column strip of one
// image segment top to bottom and runs all five convolutions as a chain of per-layer pipelines that are coupled only
// through shared-memory rings.
//
// Data flow (one image row at a time):
//   loader (4 warps)      computes F0 for the strip row from the half-resolution tensor and x, rounds it to bf16 (the value
//                         the unfused path stores) and writes ONE K=16 operand row G0 that holds the pre-activated versions
//                         relu(s_c*F0+t_c) of the three channels for ALL five consumers (K slot 3c+ch; each consumer's
//                         weight block is zero outside its own three slots).
//   issuer L_c (1 warp)   for input row j: [G0 | V(c,1) .. V(c,c)] (one K=16 step per group) x three horizontal taps (the
//                         A descriptor shifted by s pixels) -> tcgen05.mma with N = 48 = [W(r=2) | W(r=1) | W(r=0)]:
//                         the vertical taps are folded into N and land in the TMEM accumulators of output rows j-1, j,
//                         j+1 (ring of 4 slots + 2 shadow slots per layer, slot = absolute image row mod 4).
//   epilogue E_c (4 warps) drains output row j-1 of layer c (16 channels), adds the bias, rounds to bf16 (= the value the
//                         unfused path stores in HBM) and writes, for every later consumer c' in {c+1..3, T}, the
//                         pre-activated version relu(s_c'*v + t_c') into that consumer's ring V(c', c+1) — 128 px x 32 B,
//                         SWIZZLE_32B K-major, exactly the layout tcgen05 reads.  Pixels outside the image are written as
//                         zero (padding is applied AFTER the activation, models/cdan.py:41-46).
//   issuer T / epilogue T  the 1x1 transition accumulates its five K=16 steps as the versions arrive (fixed program order,
//                         so the summation order never depends on timing), then bias + sigmoid + planar fp32 store.
// Horizontal geometry: ring index q of a strip row <-> image column w0 - 4 + q.  Each 3x3 layer loses one pixel per side,
// so the valid index range of group g is [g, 128 - g) and a strip yields 120 output columns (1920 = 16 x 120).
// Vertical geometry: a segment [h0, h1) re-computes four rows above and below (receptive field of the chain).
// Every wait is a bounded mbarrier wait (traps instead of hanging); all results are independent of batch size, strip
// segmentation and timing (ring slots are functions of the absolute image row; one issuer per accumulator).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dense_fused.cuh"
#include "ptx_sm100.cuh"
#include "stream_common.cuh"

namespace cdan {

struct FusedFdPack {
  uint8_t* d_blob = nullptr;
};

namespace {

constexpr int kRowBytes = 4096;  // one operand row: 128 pixels x 16 channels bf16
constexpr int kValidW = 120;     // output columns per strip
#ifndef CDAN_FUSED_GCONST_LDS
#define CDAN_FUSED_GCONST_LDS 1  // 0: ring constants by select chains instead of a shared-memory table — measured 10.4 vs 7.9 ms (register spills)
#endif
#ifndef CDAN_FUSED_ONE_RELEASE
#define CDAN_FUSED_ONE_RELEASE 1
#endif
#ifndef CDAN_FUSED_SLEEP_NS
#define CDAN_FUSED_SLEEP_NS 32
#endif
constexpr uint32_t kSleepNs = CDAN_FUSED_SLEEP_NS;  // back-off between mbarrier polls
// Ring depths (rows).  G0 is released by all four layers; version ring (c', g) feeds layer c' with group g.  A ring has to
// hold the structural lag between its producer and its consumer (c' - g + 1 rows for the version rings, 3 rows for G0)
// PLUS the rows that pass while the chain in between runs (two hand-offs per layer), hence deeper rings for far consumers.
constexpr int kG0Depth = 11;
__host__ __device__ constexpr int ring_depth(int cp, int g) { return 3 + 2 * (cp - g); }
__host__ __device__ constexpr int ring_base(int cp, int g) {
  int b = kG0Depth;
  for (int c = 1; c <= 3; ++c)
    for (int gg = 1; gg <= c; ++gg) {
      if (c == cp && gg == g) return b;
      b += ring_depth(c, gg);
    }
  return b;
}
__host__ __device__ constexpr int ring_id(int cp, int g) { return cp * (cp - 1) / 2 + (g - 1); }  // 0..5
constexpr int kRowSlots = ring_base(4, 1);  // 37
// closed forms used in the kernel (registers are scarce: 80 per thread), checked against the definitions above
__host__ __device__ constexpr int ring_base_fast(int cp, int g) {
  return kG0Depth + (cp == 1 ? 0 : (cp == 2 ? (g == 1 ? 3 : 8) : (g == 1 ? 11 : (g == 2 ? 18 : 23))));
}
static_assert(ring_base_fast(1, 1) == ring_base(1, 1) && ring_base_fast(2, 1) == ring_base(2, 1) && ring_base_fast(2, 2) == ring_base(2, 2) &&
              ring_base_fast(3, 1) == ring_base(3, 1) && ring_base_fast(3, 2) == ring_base(3, 2) && ring_base_fast(3, 3) == ring_base(3, 3), "ring layout");
// Transition partial sums: fp32 [row][ring index][4] (3 output channels), written by the loader (bias + F0 term), updated by
// the four epilogues in layer order, finished (sigmoid + store) by the last one.
constexpr int kTsDepth = 13, kTsRowBytes = 128 * 16;
// Parameter blob (global -> shared by one bulk copy): weights | activation tables | transition parameters | bias
constexpr int kWLayerBlock = 48 * 32;  // (layer, group, tap): 48 rows x 32 B, SWIZZLE_32B
__host__ __device__ constexpr int w_layer_off(int c, int g, int s) { return ((c * (c + 1) / 2 + g) * 3 + s) * kWLayerBlock; }
constexpr int kTabOff = 30 * kWLayerBlock;   // 6 rings x (sc 32 B | sh 32 B), bf16
constexpr int kG0TabOff = kTabOff + 6 * 64;  // sc 32 B | sh 32 B: K slot 3c+ch of the four layers
constexpr int kTActOff = kG0TabOff + 64;     // transition pre-activation, bf16: sc[80] | sh[80] (physical channels)
constexpr int kTWOff = kTActOff + 320;       // transition weights as fp32 (bf16-rounded values): [5 groups][3 outputs][16 ch]
constexpr int kBiasOff = kTWOff + 5 * 192;   // 5 x 16 fp32 (4 layers, transition)
constexpr int kBlobBytes = kBiasOff + 5 * 16 * 4;
constexpr int kBlobPad = (kBlobBytes + 1023) / 1024 * 1024;
constexpr int kTsOff = kBlobPad + kRowSlots * kRowBytes + 256;
constexpr int kSmemBytes = kTsOff + kTsDepth * kTsRowBytes + 1024;
static_assert(kSmemBytes <= 232448 - 2048, "shared memory budget");
constexpr int kThreads = 24 * 32;  // 4 layer issuers, 4 loader warps, 4 x 4 epilogue warps

struct FParams {
  int N, H, W;
  int strips, SEG, segs, nitems;
  const bf16* t4;  // [N][H/2][W/2][t4_ld] relu(bn4(convT4)), channels 0..2
  int t4_ld;
  const float* x;  // [N][3][H][W]
  float* out;      // [N][3][H][W]
  const uint8_t* blob;
  unsigned long long* trace;  // debug (CDAN_FUSED_TRACE=1): clock64 timeline of CTA 0, first item: [role][kTraceRows]
};
constexpr int kTraceRows = 96, kTraceRoles = 20;
#ifdef CDAN_FUSED_TRACE_BUILD
#define FTRACE(role, row) do { if (P.trace && blockIdx.x == 0 && item == 0 && (row) >= 0 && (row) < kTraceRows) P.trace[(role) * kTraceRows + (row)] = clock64(); } while (0)
#else
#define FTRACE(role, row) do { } while (0)
#endif

struct Seg {
  int n, w0, h0, h1;
};
__device__ __forceinline__ Seg decode_seg(const FParams& P, int item) {
  const int per_img = P.strips * P.segs;
  Seg s;
  s.n = item / per_img;
  const int r = item - s.n * per_img;
  const int seg = r / P.strips;
  s.w0 = (r - seg * P.strips) * kValidW;
  s.h0 = seg * P.SEG;
  s.h1 = min(P.H, s.h0 + P.SEG);
  return s;
}
// rows consumer c reads (c = 0..3: 3x3 layers, 4: transition)
__device__ __forceinline__ int in_lo(const Seg& s, int c) { return c == 4 ? s.h0 : max(0, s.h0 - 4 + c); }
__device__ __forceinline__ int in_hi(const Seg& s, int H, int c) { return c == 4 ? s.h1 : min(H, s.h1 + 4 - c); }

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return (1u << 16) | ((smem_addr & 0x3FFFFu) >> 4); }
__device__ __forceinline__ uint32_t sw32_off(int idx, int half) { return uint32_t(idx) * 32u + (uint32_t((half ^ (idx >> 2)) & 1) << 4); }
__device__ __forceinline__ void arrive_a(uint32_t bar_addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// Loads of the parameter blob (weights, activation tables, group constants): written once before the roles start and never
// again, so these need no ordering against the pipeline's barriers — NOT volatile, no memory clobber: the compiler may hoist
// them out of the row loop or issue them early (the epilogue warps stall on shared-memory latency, ncu short_scoreboard 24 %).
// Their addresses are derived from ordered_zero(), a volatile asm placed AFTER the wait for the blob copy: the loads depend on
// its result, so they cannot be moved above that wait.
__device__ __forceinline__ uint32_t ordered_zero() {
  uint32_t z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z)::"memory");
  return z;
}
__device__ __forceinline__ uint4 ldc128(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ldc32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ldc_f4(uint32_t addr) {
  const uint4 u = ldc128(addr);
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  const uint4 u = ptx::lds128(addr);
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

// Blocking wait, out of line: the hot loops only carry the non-blocking tests (the kernel is sensitive to the size of its
// instruction streams — 24 warps at different program counters share the SM's instruction caches).
static __device__ __noinline__ void wait_slow(uint32_t bar_addr, uint32_t parity) { ptx::mbar_wait_a(bar_addr, parity); }

// One lane polls, the warp follows: a warp-wide mbarrier.try_wait costs every lane a trip through the barrier unit
// (~400 cycles per already-completed wait measured in the first version of this kernel).
__device__ __forceinline__ void wait_l0(uint32_t bar_addr, uint32_t parity, int lane) {
  if (lane == 0 && !ptx::mbar_test_a(bar_addr, parity)) wait_slow(bar_addr, parity);
  __syncwarp();
}

// Named barrier over the four warps of one role group (ids 1..5; 0 is __syncthreads).  Only ONE thread per group polls
// mbarriers: every mbarrier event on the SM wakes every thread parked in try_wait, and with 24 polling warps those wake-ups
// were 70 % of all issued instructions (ncu, profiles/r02_fused_*).  The other 127 threads wait here instead.
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

struct Smem {
  uint32_t blob, rows, ts, full, empty, acc_done, acc_free, ts_full, ts_empty, gconst;  // shared-window addresses
  uint64_t* w_full;
  uint32_t tmem;
};
// barrier arrays: full/empty[kRowSlots]; acc_done/acc_free[4 layers][4 slots]; ts_full[5 stages][kTsDepth] (stage 0 = loader,
// stage c+1 = epilogue c), ts_empty[kTsDepth]

// SlotPhases::claim_a with the poll done by lane 0 only
__device__ __forceinline__ void claim_l0(SlotPhases& fp, uint32_t free_bars, int s, int lane) {
  if ((fp.used >> s) & 1u) {
    if (lane == 0 && !ptx::mbar_test_a(free_bars + 8u * uint32_t(s), (fp.par >> s) & 1u)) wait_slow(free_bars + 8u * uint32_t(s), (fp.par >> s) & 1u);
    fp.par ^= 1u << s;
  }
  fp.used |= 1u << s;
}

// A G0 row this consumer does not read still has to be released (the ring slot is freed by all four layers).
__device__ __forceinline__ void skip_g0(const Smem& S, Ring& g0, int lane) {
  wait_l0(S.full + 8u * uint32_t(g0.i), g0.w & 1, lane);
  if (lane == 0) arrive_a(S.empty + 8u * uint32_t(g0.i));
  g0.step(kG0Depth);
}

// ------------------------------------------------------------------------------------------------ MMA issuer of layer C
// One code path for all four layers (C is a run-time value): the four issuer warps then share their instruction stream —
// with one specialised copy per layer the kernel's hot code was ~10 distinct streams and instruction fetch dominated.
__device__ __forceinline__ void issuer_layer(const FParams& P, const Smem& S, const int C, int lane) {
  const uint32_t idesc = ptx::umma_idesc_bf16(128, 48);
  const uint64_t hi = ptx::umma_desc_sw32(0, 256) & 0xffffffff00000000ull;
  const uint32_t acc_done = S.acc_done + 32u * C, acc_free = S.acc_free + 32u * C;
  const uint32_t wbase = desc_lo(S.blob + w_layer_off(C, 0, 0));
  const uint32_t dbase = S.tmem + 96u * C;
  Ring g0;
  Ring vr[3];
  SlotPhases fp;
  ptx::mbar_wait_relaxed(S.w_full, 0);
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo0 = in_lo(sg, 0), hi0 = in_hi(sg, P.H, 0), lo = in_lo(sg, C), hi_ = in_hi(sg, P.H, C);
    for (int r = lo0; r < lo; ++r) skip_g0(S, g0, lane);
    for (int j = lo; j < hi_; ++j) {
      // all waits of the row by lane 0 (phase bookkeeping is warp-uniform register state): the common case — everything
      // ready — is one batch of non-blocking phase tests issued back to back
      if (j == lo) {
        claim_l0(fp, acc_free, (j - 1) & 3, lane);
        claim_l0(fp, acc_free, j & 3, lane);
      }
      {
        const int sn = (j + 1) & 3;
        const bool need_claim = (fp.used >> sn) & 1u;
        if (lane == 0) {
          bool ok = ptx::mbar_test_a(S.full + 8u * uint32_t(g0.i), g0.w & 1);
          if (need_claim) ok &= ptx::mbar_test_a(acc_free + 8u * uint32_t(sn), (fp.par >> sn) & 1u);
#pragma unroll
          for (int g = 1; g <= 3; ++g)
            if (g <= C) ok &= ptx::mbar_test_a(S.full + 8u * uint32_t(ring_base_fast(C, g) + vr[g - 1].i), vr[g - 1].w & 1);
          if (!ok) {
            if (need_claim) wait_slow(acc_free + 8u * uint32_t(sn), (fp.par >> sn) & 1u);
            wait_slow(S.full + 8u * uint32_t(g0.i), g0.w & 1);
#pragma unroll
            for (int g = 1; g <= 3; ++g)
              if (g <= C) wait_slow(S.full + 8u * uint32_t(ring_base_fast(C, g) + vr[g - 1].i), vr[g - 1].w & 1);
          }
        }
        if (need_claim) fp.par ^= 1u << sn;
        fp.used |= 1u << sn;
      }
      __syncwarp();
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
        const uint32_t dcol = dbase + 16u * uint32_t((j - 1) & 3);
        {
          const uint32_t a = desc_lo(S.rows + uint32_t(g0.i) * kRowBytes);
#pragma unroll
          for (int s = 0; s < 3; ++s) ptx::umma_bf16(dcol, hi | (a + 2u * s), hi | (wbase + uint32_t(s) * (kWLayerBlock >> 4)), idesc, 1u);
        }
#pragma unroll
        for (int g = 1; g <= 3; ++g) {
          if (g > C) break;
          const uint32_t a = desc_lo(S.rows + uint32_t(ring_base_fast(C, g) + vr[g - 1].i) * kRowBytes);
          const uint32_t b = wbase + uint32_t(g) * (3 * kWLayerBlock >> 4);
#pragma unroll
          for (int s = 0; s < 3; ++s) ptx::umma_bf16(dcol, hi | (a + 2u * s), hi | (b + uint32_t(s) * (kWLayerBlock >> 4)), idesc, 1u);
        }
        ptx::umma_commit_a(S.empty + 8u * uint32_t(g0.i));
#pragma unroll
        for (int g = 1; g <= 3; ++g)
          if (g <= C) ptx::umma_commit_a(S.empty + 8u * uint32_t(ring_base_fast(C, g) + vr[g - 1].i));
        ptx::umma_commit_a(acc_done + 8u * uint32_t((j - 1) & 3));
        FTRACE(C, j);
      }
      __syncwarp();
      g0.step(kG0Depth);
#pragma unroll
      for (int g = 1; g <= 3; ++g)
        if (g <= C) vr[g - 1].step(ring_depth(C, g));
    }
    // the last two accumulator rows of the segment receive no further input
    if (ptx::elect_one()) {
      ptx::umma_commit_a(acc_done + 8u * uint32_t((hi_ - 1) & 3));
      ptx::umma_commit_a(acc_done + 8u * uint32_t(hi_ & 3));
    }
    __syncwarp();
    for (int r = hi_; r < hi0; ++r) skip_g0(S, g0, lane);
  }
}

// ------------------------------------------------------------------------------------------------ epilogue of layer C
// Also one code path for all layers.  Per output row: drain the accumulator (+ shadow slot) and round it to bf16; write the
// pre-activated version for every later layer (the layer's bias is folded into the activation shift: relu(s*(v+b)+t) =
// relu(s*v + (s*b+t))); add this group's term of the 1x1 transition to the fp32 partial-sum ring; the last layer's
// epilogue finishes the transition (sigmoid, planar fp32 store, models/cdan.py:157).
// Synchronisation: the group's leader thread does all mbarrier polling and arriving; the four warps meet at a named barrier
// three times per row.  The kernel is bound by instruction issue (24 warps x a few hundred instructions per row), so this
// loop is written for instruction count.
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum_f2(uint64_t v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

__device__ __forceinline__ void epilogue_layer(const FParams& P, const Smem& S, const int C, int q, int lane) {
  const int G = C + 1;   // group this layer produces
  const int NC = 3 - C;  // later layers that read it
  const int bar_id = 1 + C;
  const bool leader = q == C && lane == 0;  // the groups' leaders sit on different SM sub-partitions
  ptx::mbar_wait_relaxed(S.w_full, 0);  // parameter blob resident
  const uint32_t cz = ordered_zero();
  const uint32_t acc_free = S.acc_free + 32u * C, acc_done = S.acc_done + 32u * C;
  // consumer k of this layer's group: layer cp = C + 1 + k reads it through ring (cp, G); base slot and activation-table
  // address come from a small shared-memory table (one LDS with an immediate offset instead of a select chain per use)
  const uint32_t gc = S.gconst + cz + uint32_t(C) * 32u;
#if CDAN_FUSED_GCONST_LDS
  auto rbk = [&](int k) { return int(ldc32(gc + 4u * k)); };
  auto tabk = [&](int k) { return ldc32(gc + 12u + 4u * k); };
#else
  // closed forms of ring_base_fast(C+1+k, C+1) and of the table address of ring_id(C+1+k, C+1): two selects instead of a
  // shared-memory load at the head of every version's dependent chain (k is a compile-time constant at every use)
  (void)gc;
  auto rbk = [&](int k) { return kG0Depth + (k == 0 ? (C == 0 ? 0 : (C == 1 ? 8 : 23)) : (k == 1 ? (C == 0 ? 3 : 18) : 11)); };
  auto tabk = [&](int k) {
    return S.blob + cz + kTabOff + 64u * uint32_t(k == 0 ? (C == 0 ? 0 : (C == 1 ? 2 : 5)) : (k == 1 ? (C == 0 ? 1 : 4) : 3));
  };
#endif
  Ring cr[3];
  Ring tsr;  // transition partial-sum ring position (rows [h0, h1) of every item, in order)
  uint32_t dpar = 0;  // acc_done phase parity per slot
  const int L = q * 32 + lane, idx = L + 1;  // output lane L is centred on ring index L + 1
  const bool idx_ok = idx >= G && idx < 128 - G;
  const uint32_t off0 = sw32_off(idx & 127, 0), off1 = sw32_off(idx & 127, 1);
  const uint32_t lb = S.tmem + (uint32_t(q * 32) << 16) + 96u * C;
  const uint32_t tact_u = S.blob + cz + kTActOff + uint32_t(G) * 32u;  // sc of this group's 16 channels (sh: + 160)
  const uint32_t tw_u = S.blob + cz + kTWOff + uint32_t(G) * 192u;     // [3 outputs][16 ch] fp32 of this group
  const uint32_t ts_in = S.ts_full + 8u * uint32_t(C * kTsDepth), ts_out = S.ts_full + 8u * uint32_t((C + 1) * kTsDepth);
  const uint32_t ts_addr0 = S.ts + uint32_t(idx & 127) * 16u;
  const size_t plane = size_t(P.H) * P.W;
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo = in_lo(sg, C), hi_ = in_hi(sg, P.H, C);
    const int olo = lo == 0 ? 0 : lo + 1, ohi = hi_ == P.H ? P.H : hi_ - 1;
    const int col = sg.w0 - 4 + idx;
    const bool keep = idx_ok && col >= 0 && col < P.W;
    float* orow = P.out + size_t(sg.n) * 3 * plane + col;
    for (int i = lo - 1; i <= hi_; ++i) {
      const int slot = i & 3;
      const bool valid = i >= olo && i < ohi;
      const bool tsrow = i >= sg.h0 && i < sg.h1;  // rows of the transition (always inside the valid range)
      // rows layer cp reads: [max(0, h0-4+cp), min(H, h1+4-cp)); a valid row is inside the image, so the clipping drops out
      bool take[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) take[k] = k < NC && valid && i >= sg.h0 - 3 + C + k && i < sg.h1 + 3 - C - k;
      // 0. the leader waits for everything this row needs: the accumulator, a free slot in every consumer ring, and the
      //    transition partial sums of the earlier stages (all but the first are usually complete already)
      if (leader) {
        bool ok = ptx::mbar_test_a(acc_done + 8u * uint32_t(slot), (dpar >> slot) & 1u);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (take[k]) ok &= ptx::mbar_test_a(S.empty + 8u * uint32_t(rbk(k) + cr[k].i), (cr[k].w & 1) ^ 1);
        if (tsrow) ok &= ptx::mbar_test_a(ts_in + 8u * uint32_t(tsr.i), tsr.w & 1);
        if (!ok) {
          wait_slow(acc_done + 8u * uint32_t(slot), (dpar >> slot) & 1u);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (take[k]) wait_slow(S.empty + 8u * uint32_t(rbk(k) + cr[k].i), (cr[k].w & 1) ^ 1);
          if (tsrow) wait_slow(ts_in + 8u * uint32_t(tsr.i), tsr.w & 1);
        }
      }
      dpar ^= 1u << slot;
      group_sync(bar_id);
      ptx::tc_fence_after_sync();
      if (leader) FTRACE(4 + C, i);
      const bool shadow = slot < 2;
      const uint32_t tm = lb + 16u * uint32_t(slot), ts = lb + 16u * uint32_t(4 + slot);
      // 1. drain the accumulator row (+ its shadow slot), re-zero it
      uint32_t v[16];
      if (valid) {
        ptx::tmem_ld16(tm, v);
        if (shadow) {
          uint32_t v2[8];
          ptx::tmem_ld8(ts, v2);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          ptx::tmem_ld8(ts + 8u, v2);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 + e] = __float_as_uint(__uint_as_float(v[8 + e]) + __uint_as_float(v2[e]));
        } else {
          ptx::tmem_wait_ld();
        }
      }
      ptx::tmem_st16_zero(tm);
      if (shadow) ptx::tmem_st16_zero(ts);
      if (C == 0 && leader) FTRACE(16, i);
      // 2. round to bf16 (the bias is folded into the consumers' activation shifts)
      __nv_bfloat162 raw[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) raw[e] = __floats2bfloat162_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
      // 3. one pre-activated version per later layer; the next layer's comes first and is released at once together with
      //    the accumulator row (it is on the critical path of the chain)
      auto write_version = [&](int k) {
        const uint32_t tb = tabk(k);
        const uint4 sc0 = ldc128(tb), sc1 = ldc128(tb + 16u), sh0 = ldc128(tb + 32u), sh1 = ldc128(tb + 48u);
        const __nv_bfloat162* c0p = reinterpret_cast<const __nv_bfloat162*>(&sc0);
        const __nv_bfloat162* c1p = reinterpret_cast<const __nv_bfloat162*>(&sc1);
        const __nv_bfloat162* h0p = reinterpret_cast<const __nv_bfloat162*>(&sh0);
        const __nv_bfloat162* h1p = reinterpret_cast<const __nv_bfloat162*>(&sh1);
        uint4 o0, o1;
        __nv_bfloat162* a0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* a1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          a0[e] = __hfma2_relu(raw[e], c0p[e], h0p[e]);
          a1[e] = __hfma2_relu(raw[4 + e], c1p[e], h1p[e]);
        }
        if (!keep) {  // outside the image / the layer's valid range: zero AFTER activation
          o0 = make_uint4(0u, 0u, 0u, 0u);
          o1 = make_uint4(0u, 0u, 0u, 0u);
        }
        if (idx < 128) {
          const uint32_t base = S.rows + uint32_t(rbk(k) + cr[k].i) * kRowBytes;
          ptx::sts128(base + off0, o0);
          ptx::sts128(base + off1, o1);
        }
      };
#if CDAN_FUSED_ONE_RELEASE
      // one release point per row (below): the chain of layers is throughput-bound, not latency-bound — the rings absorb the
      // extra lag — and every release costs a proxy fence (drains the row's shared-memory stores) plus a group barrier
      if (take[0]) write_version(0);
#else
      if (take[0]) {
        write_version(0);
        ptx::fence_proxy_async_smem();
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before_sync();
      group_sync(bar_id);
      if (leader) {
        arrive_a(acc_free + 8u * uint32_t(slot));
        if (take[0]) arrive_a(S.full + 8u * uint32_t(rbk(0) + cr[0].i));
        if (C == 0) FTRACE(13, i);
      }
#endif
      if (take[1]) write_version(1);
      if (take[2]) write_version(2);
      if (C == 0 && leader) FTRACE(14, i);
      // 4. this group's term of the 1x1 transition (16 channels x 3 outputs, packed fp32 FMAs) on top of the partial sums so
      //    far (loader: bias + F0 term; earlier layers in order)
      if (tsrow) {
        uint64_t av[8];
        {
          const uint4 sc0 = ldc128(tact_u), sc1 = ldc128(tact_u + 16u), sh0 = ldc128(tact_u + 160u), sh1 = ldc128(tact_u + 176u);
          const __nv_bfloat162* c0p = reinterpret_cast<const __nv_bfloat162*>(&sc0);
          const __nv_bfloat162* c1p = reinterpret_cast<const __nv_bfloat162*>(&sc1);
          const __nv_bfloat162* h0p = reinterpret_cast<const __nv_bfloat162*>(&sh0);
          const __nv_bfloat162* h1p = reinterpret_cast<const __nv_bfloat162*>(&sh1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 x0 = __hfma2_relu(raw[e], c0p[e], h0p[e]), x1 = __hfma2_relu(raw[4 + e], c1p[e], h1p[e]);
            const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&x0), u1 = *reinterpret_cast<const uint32_t*>(&x1);
            av[e] = pack_f2(bflo(u0), bfhi(u0));
            av[4 + e] = pack_f2(bflo(u1), bfhi(u1));
          }
        }
        float tt[3];
#pragma unroll
        for (int co = 0; co < 3; ++co) {  // weights: [output][16 channels] fp32 per group
          uint64_t acc = 0ull;
#pragma unroll
          for (int h4 = 0; h4 < 4; ++h4) {
            const uint4 w = ldc128(tw_u + uint32_t(co) * 64u + uint32_t(h4) * 16u);
            acc = ffma2((uint64_t(w.y) << 32) | w.x, av[2 * h4], acc);
            acc = ffma2((uint64_t(w.w) << 32) | w.z, av[2 * h4 + 1], acc);
          }
          tt[co] = sum_f2(acc);
        }
        const uint32_t addr = ts_addr0 + uint32_t(tsr.i) * kTsRowBytes;
        float4 acc = lds_f4(addr);
        acc.x += tt[0]; acc.y += tt[1]; acc.z += tt[2];
        if (C < 3) {
          if (idx < 128) ptx::sts128(addr, make_uint4(__float_as_uint(acc.x), __float_as_uint(acc.y), __float_as_uint(acc.z), 0u));
        } else if (idx >= 4 && idx < 124 && col < P.W) {
          float* o = orow + size_t(i) * P.W;
          o[0] = 1.0f / (1.0f + __expf(-acc.x));
          o[plane] = 1.0f / (1.0f + __expf(-acc.y));
          o[2 * plane] = 1.0f / (1.0f + __expf(-acc.z));
        }
      }
      if (C == 0 && leader) FTRACE(15, i);
      // 5. release the remaining versions and the partial sums
#if CDAN_FUSED_ONE_RELEASE
      {
        ptx::fence_proxy_async_smem();
        ptx::tmem_wait_st();
        ptx::tc_fence_before_sync();
        group_sync(bar_id);
        // (measured slower: dealing the waits / releases to lane 0 of all four warps, 8.08 vs 7.92 ms; ONE group barrier per row
        //  with the leader polling row i+1 before it releases row i, 9.57 ms — holding a row back costs the chain its slack)
        if (leader) {
          arrive_a(acc_free + 8u * uint32_t(slot));
          if (take[0]) arrive_a(S.full + 8u * uint32_t(rbk(0) + cr[0].i));
          if (take[1]) arrive_a(S.full + 8u * uint32_t(rbk(1) + cr[1].i));
          if (take[2]) arrive_a(S.full + 8u * uint32_t(rbk(2) + cr[2].i));
          if (tsrow) arrive_a(C < 3 ? ts_out + 8u * uint32_t(tsr.i) : S.ts_empty + 8u * uint32_t(tsr.i));
        }
      }
#else
      if (take[1] || tsrow) {
        ptx::fence_proxy_async_smem();
        group_sync(bar_id);
        if (leader) {
          if (take[1]) arrive_a(S.full + 8u * uint32_t(rbk(1) + cr[1].i));
          if (take[2]) arrive_a(S.full + 8u * uint32_t(rbk(2) + cr[2].i));
          if (tsrow) arrive_a(C < 3 ? ts_out + 8u * uint32_t(tsr.i) : S.ts_empty + 8u * uint32_t(tsr.i));
        }
      }
#endif
      if (leader) FTRACE(8 + C, i);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (take[k]) cr[k].step(3 + 2 * k);
      if (tsrow) tsr.step(kTsDepth);
    }
  }
}

// ------------------------------------------------------------------------------------------------ loader (F0 -> G0, TS)
__device__ __forceinline__ void loader(const FParams& P, const Smem& S, int t, int lane) {
  Ring gr, tsr;
  ptx::mbar_wait_relaxed(S.w_full, 0);
  const uint32_t blob_c = S.blob + ordered_zero();  // base of the constant loads (ordered after the wait above)
  const uint32_t off0 = sw32_off(t, 0), off1 = sw32_off(t, 1);
  const int IH = P.H >> 1, IW = P.W >> 1;
  const size_t plane = size_t(P.H) * P.W;
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo0 = in_lo(sg, 0), hi0 = in_hi(sg, P.H, 0);
    const int col = sg.w0 - 4 + t;
    const bool col_ok = col >= 0 && col < P.W;
    // bilinear x2, align_corners = False: even output -> inputs (i-1: .25, i: .75), odd -> (i: .75, i+1: .25), clamped
    const int cx = col >> 1;
    const int c0 = (col & 1) ? cx : max(cx - 1, 0), c1 = (col & 1) ? min(cx + 1, IW - 1) : cx;
    const float wx0 = (col & 1) ? 0.75f : 0.25f, wx1 = (col & 1) ? 0.25f : 0.75f;
    const bf16* tn = P.t4 + size_t(sg.n) * IH * IW * P.t4_ld;
    const float* xn = P.x + size_t(sg.n) * 3 * plane + col;
    for (int r = lo0; r < hi0; r += 2) {  // rows r (even) and r + 1 share the three half-resolution rows j-1, j, j+1
      const int j = r >> 1;
      const int jm = max(j - 1, 0), jp = min(j + 1, IH - 1);
      // horizontally interpolated half-resolution rows (3 channels each), and the six x values of both output rows
      float hz[3][3];
      float xv[2][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) hz[a][ch] = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) xv[dy][ch] = 0.f;
      if (col_ok) {
        uint2 q0[3], q1[3];
        const int rws[3] = {jm, j, jp};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          q0[a] = *reinterpret_cast<const uint2*>(tn + (size_t(rws[a]) * IW + c0) * P.t4_ld);
          q1[a] = *reinterpret_cast<const uint2*>(tn + (size_t(rws[a]) * IW + c1) * P.t4_ld);
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) xv[dy][ch] = xn[ch * plane + size_t(r + dy) * P.W];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          hz[a][0] = wx0 * bflo(q0[a].x) + wx1 * bflo(q1[a].x);
          hz[a][1] = wx0 * bfhi(q0[a].x) + wx1 * bfhi(q1[a].x);
          hz[a][2] = wx0 * bflo(q0[a].y) + wx1 * bflo(q1[a].y);
        }
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int row = r + dy;
        const bool tsrow = row >= sg.h0 && row < sg.h1;
        if (t == 0) {
          bool ok = ptx::mbar_test_a(S.empty + 8u * uint32_t(gr.i), (gr.w & 1) ^ 1);
          if (tsrow) ok &= ptx::mbar_test_a(S.ts_empty + 8u * uint32_t(tsr.i), (tsr.w & 1) ^ 1);
          if (!ok) {
            wait_slow(S.empty + 8u * uint32_t(gr.i), (gr.w & 1) ^ 1);
            if (tsrow) wait_slow(S.ts_empty + 8u * uint32_t(tsr.i), (tsr.w & 1) ^ 1);
          }
        }
        group_sync(5);
        uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = make_uint4(0u, 0u, 0u, 0u);
        float4 tsum = ldc_f4(blob_c + kBiasOff + 4 * 64);  // transition bias
        if (col_ok) {
          const int r0 = dy ? 1 : 0, r1 = dy ? 2 : 1;
          const float wy0 = dy ? 0.75f : 0.25f, wy1 = dy ? 0.25f : 0.75f;
          __nv_bfloat16 f[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) f[ch] = __float2bfloat16_rn((wy0 * hz[r0][ch] + wy1 * hz[r1][ch]) + xv[dy][ch]);
          // K slot k = 3*layer + channel (k = 12..15 unused: scale = shift = 0); tables stay in shared memory
          const uint4 sc0 = ldc128(blob_c + kG0TabOff), sh0 = ldc128(blob_c + kG0TabOff + 32u);
          const uint4 sc1 = ldc128(blob_c + kG0TabOff + 16u), sh1 = ldc128(blob_c + kG0TabOff + 48u);
          const __nv_bfloat162* c0p = reinterpret_cast<const __nv_bfloat162*>(&sc0);
          const __nv_bfloat162* h0p = reinterpret_cast<const __nv_bfloat162*>(&sh0);
          const __nv_bfloat162* c1p = reinterpret_cast<const __nv_bfloat162*>(&sc1);
          const __nv_bfloat162* h1p = reinterpret_cast<const __nv_bfloat162*>(&sh1);
          __nv_bfloat162* a0 = reinterpret_cast<__nv_bfloat162*>(&o0);
          __nv_bfloat162* a1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 rp;
            rp.x = f[(2 * e) % 3];
            rp.y = f[(2 * e + 1) % 3];
            a0[e] = __hfma2_relu(rp, c0p[e], h0p[e]);
            rp.x = f[(8 + 2 * e) % 3];
            rp.y = f[(8 + 2 * e + 1) % 3];
            a1[e] = __hfma2_relu(rp, c1p[e], h1p[e]);
          }
          if (tsrow) {  // F0 term of the transition (physical channels 0..2)
            const uint2 tsc = *reinterpret_cast<const uint2*>(__cvta_shared_to_generic(S.blob + kTActOff));
            const uint2 tsh = *reinterpret_cast<const uint2*>(__cvta_shared_to_generic(S.blob + kTActOff + 160u));
            __nv_bfloat162 p01, p2;
            p01.x = f[0]; p01.y = f[1]; p2.x = f[2]; p2.y = f[2];
            const __nv_bfloat162 a01 = __hfma2_relu(p01, *reinterpret_cast<const __nv_bfloat162*>(&tsc.x), *reinterpret_cast<const __nv_bfloat162*>(&tsh.x));
            const __nv_bfloat162 a2 = __hfma2_relu(p2, *reinterpret_cast<const __nv_bfloat162*>(&tsc.y), *reinterpret_cast<const __nv_bfloat162*>(&tsh.y));
            const float av[3] = {__low2float(a01), __high2float(a01), __low2float(a2)};
#pragma unroll
            const float4 wa = ldc_f4(blob_c + kTWOff), wb = ldc_f4(blob_c + kTWOff + 64u), wc = ldc_f4(blob_c + kTWOff + 128u);
            tsum.x = fmaf(wa.z, av[2], fmaf(wa.y, av[1], fmaf(wa.x, av[0], tsum.x)));
            tsum.y = fmaf(wb.z, av[2], fmaf(wb.y, av[1], fmaf(wb.x, av[0], tsum.y)));
            tsum.z = fmaf(wc.z, av[2], fmaf(wc.y, av[1], fmaf(wc.x, av[0], tsum.z)));
          }
        }
        const uint32_t base = S.rows + uint32_t(gr.i) * kRowBytes;
        ptx::sts128(base + off0, o0);
        ptx::sts128(base + off1, o1);
        if (tsrow) ptx::sts128(S.ts + uint32_t(tsr.i) * kTsRowBytes + uint32_t(t) * 16u,
                               make_uint4(__float_as_uint(tsum.x), __float_as_uint(tsum.y), __float_as_uint(tsum.z), 0u));
        ptx::fence_proxy_async_smem();
        group_sync(5);
        if (t == 0) {
          arrive_a(S.full + 8u * uint32_t(gr.i));
          if (tsrow) arrive_a(S.ts_full + 8u * uint32_t(tsr.i));
          FTRACE(12, row);
        }
        gr.step(kG0Depth);
        if (tsrow) tsr.step(kTsDepth);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) dense_fused_kernel(const FParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kRowSlots], empty[kRowSlots], acc_done[16], acc_free[16], ts_full[5 * kTsDepth], ts_empty[kTsDepth], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t gconst[4 * 8];  // per epilogue group: ring base slot of consumer k = 0..2, activation-table address of k = 0..2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < 12) {
    const int c = tid / 3, k = tid % 3, cp = min(c + 1 + k, 3), g = min(c + 1, cp);
    gconst[c * 8 + k] = uint32_t(ring_base_fast(cp, g));
    gconst[c * 8 + 3 + k] = ptx::smem_u32(smem) + kTabOff + uint32_t(ring_id(cp, g)) * 64u;
  }
  if (tid == 0) {
    for (int i = 0; i < kRowSlots; ++i) {
      ptx::mbar_init(&full[i], 1);  // one arrival by the producing group's leader
      ptx::mbar_init(&empty[i], i < kG0Depth ? 4 : 1);
    }
    for (int i = 0; i < 16; ++i) {
      ptx::mbar_init(&acc_done[i], 1);
      ptx::mbar_init(&acc_free[i], 1);
    }
    for (int i = 0; i < 5 * kTsDepth; ++i) ptx::mbar_init(&ts_full[i], 1);
    for (int i = 0; i < kTsDepth; ++i) ptx::mbar_init(&ts_empty[i], 1);
    ptx::mbar_init(&w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  {  // operand rows start as zeros (ring index 0 of groups 1-3 and the unused K slots are never written afterwards)
    uint4* rows = reinterpret_cast<uint4*>(smem + kBlobPad);
    for (int i = tid; i < (kRowSlots * kRowBytes + 256) / 16; i += kThreads) rows[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  Smem S;
  S.blob = ptx::smem_u32(smem);
  S.rows = S.blob + kBlobPad;
  S.ts = S.blob + kTsOff;
  S.full = ptx::smem_u32(full);
  S.empty = ptx::smem_u32(empty);
  S.acc_done = ptx::smem_u32(acc_done);
  S.acc_free = ptx::smem_u32(acc_free);
  S.ts_full = ptx::smem_u32(ts_full);
  S.ts_empty = ptx::smem_u32(ts_empty);
  S.gconst = ptx::smem_u32(gconst);
  S.w_full = &w_full;
  S.tmem = tmem_base_s;
  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_arrive_expect_tx(&w_full, kBlobBytes);
    for (uint32_t off = 0; off < uint32_t(kBlobBytes); off += 16384)
      ptx::bulk_g2s(smem + off, P.blob + off, min(16384u, uint32_t(kBlobBytes) - off), &w_full);
  }
  if (warp >= 8 && warp < 12) {  // accumulators start at zero (all layers accumulate, rows are re-zeroed when drained)
    const uint32_t lb = S.tmem + (uint32_t((warp & 3) * 32) << 16);
    for (int c = 0; c < 512; c += 16) ptx::tmem_st16_zero(lb + c);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();

  if (warp < 4) issuer_layer(P, S, warp, lane);
  else if (warp < 8) loader(P, S, (warp - 4) * 32 + lane, lane);
  else epilogue_layer(P, S, (warp - 8) >> 2, warp & 3, lane);

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(S.tmem, 512);
  }
}

void put_sw32(uint8_t* blk, int row, int k, float val) {
  const bf16 b = __float2bfloat16_rn(val);
  std::memcpy(blk + row * 32 + ((((k >> 3) ^ (row >> 2)) & 1) << 4) + (k & 7) * 2, &b, 2);
}

}  // namespace

int fused_fd_pack_create(const FusedFdLayer layers[5], FusedFdPack** out) {
  *out = nullptr;
  for (int c = 0; c < 5; ++c) {
    const FusedFdLayer& L = layers[c];
    if (L.Cin != 16 * (c + 1) || L.CoutP != 16 || !L.w || !L.bias || !L.pre_s || !L.pre_t)
      return fail("fused final dense block: unexpected layer geometry");
  }
  std::vector<uint8_t> blob(kBlobBytes, 0);
  auto wat = [&](const FusedFdLayer& L, int tap, int ci, int co) { return L.w[(size_t(tap) * L.Cin + ci) * L.CoutP + co]; };
  for (int c = 0; c < 4; ++c)
    for (int g = 0; g <= c; ++g)
      for (int s = 0; s < 3; ++s) {
        uint8_t* blk = blob.data() + w_layer_off(c, g, s);
        for (int pos = 0; pos < 3; ++pos)  // window position pos <-> kernel row 2 - pos (input row j feeds output row j-1+pos)
          for (int co = 0; co < 16; ++co) {
            const int tap = (2 - pos) * 3 + s;
            if (g == 0) {
              for (int ch = 0; ch < 3; ++ch) put_sw32(blk, pos * 16 + co, 3 * c + ch, wat(layers[c], tap, ch, co));
            } else {
              for (int k = 0; k < 16; ++k) put_sw32(blk, pos * 16 + co, k, wat(layers[c], tap, 16 * g + k, co));
            }
          }
      }
  auto put_b = [&](int off, float v) {
    const bf16 b = __float2bfloat16_rn(v);
    std::memcpy(blob.data() + off, &b, 2);
  };
  for (int cp = 1; cp <= 3; ++cp)
    for (int g = 1; g <= cp; ++g)
      for (int k = 0; k < 16; ++k) {
        // group g = output of layer g-1: its bias is folded into the shift, relu(s*(v+b)+t) = relu(s*v + (s*b+t))
        put_b(kTabOff + ring_id(cp, g) * 64 + k * 2, layers[cp].pre_s[16 * g + k]);
        put_b(kTabOff + ring_id(cp, g) * 64 + 32 + k * 2, layers[cp].pre_s[16 * g + k] * layers[g - 1].bias[k] + layers[cp].pre_t[16 * g + k]);
      }
  for (int cp = 0; cp <= 3; ++cp)
    for (int ch = 0; ch < 3; ++ch) {
      put_b(kG0TabOff + (3 * cp + ch) * 2, layers[cp].pre_s[ch]);
      put_b(kG0TabOff + 32 + (3 * cp + ch) * 2, layers[cp].pre_t[ch]);
    }
  // transition (1x1, 80 physical input channels -> 3): pre-activation tables (bf16) and weights as fp32 values of their
  // bf16 roundings (the products with bf16 activations are then exact in fp32, as on the tensor-core path)
  for (int ci = 0; ci < 80; ++ci) {
    put_b(kTActOff + ci * 2, layers[4].pre_s[ci]);
    put_b(kTActOff + 160 + ci * 2, ci < 16 ? layers[4].pre_t[ci] : layers[4].pre_s[ci] * layers[ci / 16 - 1].bias[ci % 16] + layers[4].pre_t[ci]);
    for (int co = 0; co < 3; ++co) {
      const float v = __bfloat162float(__float2bfloat16_rn(wat(layers[4], 0, ci, co)));
      std::memcpy(blob.data() + kTWOff + ((ci / 16) * 3 + co) * 64 + (ci % 16) * 4, &v, 4);
    }
  }
  for (int c = 0; c < 5; ++c)
    for (int co = 0; co < 16; ++co) {
      const float v = layers[c].bias[co];
      std::memcpy(blob.data() + kBiasOff + (c * 16 + co) * 4, &v, 4);
    }
  FusedFdPack* p = new FusedFdPack();
  if (cudaMalloc(&p->d_blob, blob.size()) != cudaSuccess ||
      cudaMemcpy(p->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    fused_fd_pack_destroy(p);
    return fail("fused final dense block: parameter upload failed");
  }
  *out = p;
  return 0;
}

void fused_fd_pack_destroy(FusedFdPack* p) {
  if (!p) return;
  if (p->d_blob) cudaFree(p->d_blob);
  delete p;
}

int fused_fd_launch(const FusedFdPack& pk, const void* t4, int t4_ld, const float* x, float* y, int N, int H, int W,
                    cudaStream_t stream) {
  if (H % 2 || W % 2 || N <= 0) return fail("fused final dense block: H and W must be even");
  if (t4_ld % 4 != 0 || reinterpret_cast<uintptr_t>(t4) % 8 != 0) return fail("fused final dense block: misaligned half-resolution input");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  FParams P{};
  P.N = N; P.H = H; P.W = W;
  P.t4 = reinterpret_cast<const bf16*>(t4); P.t4_ld = t4_ld;
  P.x = x; P.out = y; P.blob = pk.d_blob;
  P.strips = ceil_div(W, kValidW);
  // Segment height: every segment pays ~8 re-computed rows plus ~10 rows of pipeline fill; more segments balance the
  // persistent CTAs better.  Pick the (even) height with the lowest cost = rounds of items per CTA x rows per item.
  long best_cost = -1;
  for (int segs = 1; segs <= std::max(1, H / 16); ++segs) {
    const int seg = (ceil_div(H, segs) + 1) & ~1;
    const int nseg = ceil_div(H, seg);
    const long items = long(N) * P.strips * nseg;
    static const int seg_over = getenv("CDAN_FUSED_SEG_OVERHEAD") ? atoi(getenv("CDAN_FUSED_SEG_OVERHEAD")) : 18;  // A/B switch
    const long cost = ((items + sms - 1) / sms) * (seg + seg_over);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      P.SEG = seg;
      P.segs = nseg;
    }
  }
  P.nitems = N * P.strips * P.segs;
  CDAN_CUDA_OK(cudaFuncSetAttribute(dense_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  static const bool tracing = getenv("CDAN_FUSED_TRACE") && atoi(getenv("CDAN_FUSED_TRACE")) != 0;
  static unsigned long long* d_trace = nullptr;
  if (tracing) {
    if (!d_trace) cudaMalloc(&d_trace, kTraceRoles * kTraceRows * sizeof(unsigned long long));
    cudaMemsetAsync(d_trace, 0, kTraceRoles * kTraceRows * sizeof(unsigned long long), stream);
    P.trace = d_trace;
  }
  dense_fused_kernel<<<std::min(P.nitems, sms), kThreads, kSmemBytes, stream>>>(P);
  CDAN_CUDA_OK(cudaGetLastError());
  if (tracing) {
    cudaStreamSynchronize(stream);
    std::vector<unsigned long long> t(kTraceRoles * kTraceRows);
    cudaMemcpy(t.data(), d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (auto v : t) if (v && v < t0) t0 = v;
    fprintf(stderr, "FUSED TRACE N=%d H=%d W=%d SEG=%d segs=%d strips=%d items=%d\n", N, H, W, P.SEG, P.segs, P.strips, P.nitems);
    const char* names[kTraceRoles] = {"L0_issued", "L1_issued", "L2_issued", "L3_issued", "E0_start", "E1_start", "E2_start", "E3_start",
                                      "E0_done", "E1_done", "E2_done", "E3_done", "loader", "E0_accfree", "E0_versions", "E0_tsterm", "E0_drained", "", "", ""};
    for (int r = 0; r < 17; ++r) {
      fprintf(stderr, "%-10s", names[r]);
      for (int i = 0; i < kTraceRows; ++i) fprintf(stderr, " %7lld", t[r * kTraceRows + i] ? (long long)(t[r * kTraceRows + i] - t0) : -1ll);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace cdan
