// Streaming tcgen05 convolution for sm_100a: the kernel behind every narrow-output layer of CDAN (the 16 dense-block
// 3x3 layers, the 4 transition 1x1 convs, decoder.conv4 and encoder.conv1) — reference models/cdan.py:8-53,70-98.
//
// Formulation.  A CTA walks a COLUMN STRIP of one image top to bottom.  One input row of the strip (128 pixels incl.
// the horizontal halo, one 64-channel K-chunk = 16 KB, 128-byte-swizzled K-major) is one pipeline stage and one MMA
// M-block: TMEM lane p <-> strip pixel p.  For a 3x3 convolution the three VERTICAL taps are folded into the MMA N
// dimension: input row j is multiplied once by [W(r=2) | W(r=1) | W(r=0)] (N = 3*NT) and the three NT-column blocks
// accumulate straight into the accumulators of output rows j-1, j, j+1, which sit in consecutive slots of a TMEM
// RING.  The three horizontal taps are three A descriptors shifted by one pixel row (128 B) of the same stage.
//   * N = 48 instead of 16 costs 44 instead of 39 cycles per tcgen05.mma (measured, profiles/r01_mma_tmem_probe.log), so a
//     16-channel dense layer issues 3x fewer, equally expensive MMAs: 20 % -> 55 % of the tensor pipe.
//   * every activation row is loaded from HBM exactly once per strip (no vertical halo re-reads), stages are small
//     (16 KB) so 8-12 of them are in flight per SM.
//   * a finished accumulator row is read by the epilogue and immediately re-zeroed (tcgen05.st), so all MMAs use
//     accumulate=1.  Ring wrap: input row G writes slots (G mod R)+{0,1,2}; slots R and R+1 are "shadow" copies of
//     slots 0 and 1, summed by the epilogue — the N=3*NT window never has to be split.
// encoder.conv1 (3 input channels, fp32 NCHW) additionally folds the three HORIZONTAL taps into K (k = s*3+ci, one
// K=16 step), so a whole input row is ONE tcgen05.mma with N = 192, and its epilogue max-pools 2x2 (vertical pair =
// two ring slots of the same thread, horizontal pair = lane^1).
// 1x1 convolutions use the same pipeline with a window of one slot and fresh accumulators (no clearing).
//
// Two kernels share this formulation:
//   conv_stream_kernel<IN,FOLD,EPI>   one image row per stage: conv1 (planar fp32 input, K-fold, fused 2x2 max-pool),
//                                     the TMA-fed layers (conv2 and decoder.conv3 in the WIDE nine-tap form with the
//                                     weights resident, decoder.conv4 as a nine-tap fold) and the A/B variants of the
//                                     dense layers.  Warp roles: 0 TMA producer (+ resident weights), 1 MMA issuer,
//                                     2 TMEM allocator, 4-11 epilogue (two groups), then the A-stage workers.
//   conv_stream2_kernel<FOLD,EPI,GP>  two image rows per stage: all dense-block layers (3x3 row fold, 1x1 transitions).
//                                     GP selects the input layout (0 NHWC, 1 group-planar, 2 hybrid: compact NHWC head
//                                     + 16-channel group planes); up to three MMA issuer warps take row
//                                     pairs round-robin under a completion token.
// All waits are bounded (ptx::mbar_wait traps), so a protocol bug fails loudly instead of hanging.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_umma.cuh"
#include "ptx_sm100.cuh"
#include "stream_common.cuh"

namespace cdan {

struct StreamPack {
  uint8_t* d_w = nullptr;   // fold / 1x1 image: [pass][chunk][s][NMMA rows][128 B swizzled]
  uint8_t* d_wk = nullptr;  // conv1 K-folded image: [192 rows][128 B swizzled] (only when Cin == 3, ks == 3, Cout <= 64)
  uint8_t* d_wr = nullptr;  // 16-wide 3x3, row-fold form: [chunk][s][48 rows = pos*16+co][128 B swizzled]
  uint8_t* d_ww = nullptr;  // wide 3x3 image (Cout 17..128, weights resident): [chunk][tap r*3+s][NTw rows][128 B swizzled]
  float* d_bias = nullptr;  // [npass * NT]
  int Cin = 0, Cout = 0, ks = 3, NT = 0, npass = 1, nchunks = 0;
  int NTw = 0, npass_w = 1;
  size_t pass_bytes = 0, wide_bytes = 0, rfold_bytes = 0;
};

namespace {

// kSPro2 = kSPro in a half-size CTA (512 threads, 256 TMEM columns, <= 113 KB shared memory) so that TWO CTAs share an
// SM: every role of this kernel is a single-warp latency chain, and two independent pipelines per SM overlap them.
enum SIn : int { kSTma = 0, kSPro = 1, kSNchw = 2, kSPro2 = 3 };
__host__ __device__ constexpr bool is_pro(int in_mode) { return in_mode == 1 || in_mode == 3; }
enum SEpi : int { kSStore = 0, kSPool = 1, kSNchwOut = 2 };

constexpr int kStage = 16384;  // one input row: 128 pixels x 64 channels bf16
constexpr int kMaxSA = 12;
constexpr int kMaxSA2 = 16;  // two-row kernel (group-planar stages can be as small as 8 KB)
// conv1 raw fp32 row ring: one stage = 3 channels x 136 pixels starting at column w0-4 (TMA needs the innermost start
// coordinate 16-byte aligned; w0 is a multiple of 4), 2 KB per stage.
constexpr int kRawStages = 8, kRawW = 136, kRawFloats = 512;
constexpr int kMaxR = 32;
// Warp layout: warps 0-3 control, 4-11 epilogue (two groups of four alternating accumulator rows), then the A-stage
// workers (16 warps = two groups alternating stages for the dense pre-activation, 4 for conv1's row builder, none for
// TMA-fed layers).
constexpr int kEpiWarp0 = 4;
__host__ __device__ constexpr int epi_warps(int in_mode) { return in_mode == 3 ? 4 : 8; }
__host__ __device__ constexpr int work_warps(int in_mode) { return in_mode == 0 ? 0 : (in_mode == 1 ? 16 : (in_mode == 3 ? 8 : 4)); }

struct SParams {
  int N, H, W, Cin;
  int pad;       // 1: 3x3, 0: 1x1
  int NT, NMMA;  // output channels per accumulator row / MMA N
  int wide;      // 3x3 with NT > 16: nine separate taps (3 ring slots x 3 shifted views), no shadow slots
  int SW;        // TMEM columns per ring slot (= NT, or 3*NT when the horizontal taps are folded into N as well)
  int R;         // ring slots (excluding the two shadow slots)
  int TW, strips, SEG, segs, nitems;
  int nchunks, nS, SA;
  int wsplit;       // two-row kernel: both worker groups process every stage, one image row each (else alternate stages)
  int nh, ngroups;  // two-row kernel: NHWC stages per row pair (hybrid: the head), planar 16-channel groups
  int npc, gps;     // two-row kernel: stages per row pair; group-planar: 16-channel groups per stage (4, or 5 for Cin = 80)
  int tok_inside;   // two-row kernel: the elected lane waits for the issuer token inside its issue region (A/B switch)
  int ni, pt;       // two-row kernel: number of MMA issuer warps (2 or 3), row pairs per issuer turn (1 or 2)
  int stage_bytes;  // group-planar input only: bytes per two-row stage (8 KB per 16-channel group)
  int relu, sigmoid, Cout;
  uint32_t wbytes;
  const bf16* in;
  int in_ld;
  const float* in_nchw;
  const float* pre_s;
  const float* pre_t;
  const uint8_t* wpack;
  const float* bias;
  bf16* out;
  int out_ld;
  float* out_nchw;
  int ablate;  // debug (CDAN_ABLATE bitmask): 1 skip epilogue global stores, 2 skip worker math, 4 skip MMA issue, 8 skip epilogue TMEM traffic
  unsigned long long* trace;  // timeline of CTA 0 (debug builds with -DCDAN_STREAM_TRACE_BUILD): [role][kTraceN]
};
constexpr int kTraceN = 256;
#ifdef CDAN_STREAM_TRACE_BUILD
#define STRACE(role, idx) do { if (P.trace && blockIdx.x == 0 && (idx) < kTraceN) P.trace[(role) * kTraceN + (idx)] = clock64(); } while (0)
#else
#define STRACE(role, idx) do { } while (0)
#endif

struct Item {
  int n, w0, h0, h1;
};
__device__ __forceinline__ Item decode_item(const SParams& P, int item) {
  const int per_img = P.strips * P.segs;
  Item it;
  it.n = item / per_img;
  const int r = item - it.n * per_img;
  const int seg = r / P.strips;
  it.w0 = (r - seg * P.strips) * P.TW;
  it.h0 = seg * P.SEG;
  it.h1 = min(P.H, it.h0 + P.SEG);
  return it;
}

template <int IN, int FOLD, int EPI>
__global__ void __launch_bounds__(32 * (kEpiWarp0 + epi_warps(IN) + work_warps(IN)), IN == 3 ? 2 : 1) conv_stream_kernel(const __grid_constant__ CUtensorMap tmapA, const SParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = sA + size_t(P.SA) * kStage + 1024;
  float* s_pre_s = reinterpret_cast<float*>(sW + P.wbytes);
  float* s_pre_t = s_pre_s + P.nchunks * 64;
  float* s_bias = s_pre_t + P.nchunks * 64;
  float* s_raw = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_bias + max(P.NT, 64)) + 127) & ~uintptr_t(127));  // kSNchw only

  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], raw_full[kMaxSA], raw_empty[kRawStages], acc_done[kMaxR], acc_free[kMaxR], w_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // FOLD: 1 = 1x1 conv; 3 = 3x3 in the input mode's default form (kSPro: nine-tap fold + shift epilogue, kSNchw: conv1
  // K-fold, kSTma: wide nine-tap); 9 = 3x3 TMA-fed with the nine-tap fold (16-channel outputs, e.g. decoder.conv4)
  //       4 = 3x3 with only the vertical taps folded into N (N = 3*NT) and the horizontal taps as shifted A views
  constexpr int PAD = FOLD != 1 ? 1 : 0;
  constexpr bool RFOLD = FOLD == 4;
  constexpr int kEpiWarps = epi_warps(IN), kWorkWarp0 = kEpiWarp0 + kEpiWarps, kWorkWarps = work_warps(IN);
  constexpr int kEG = kEpiWarps / 4;  // epilogue groups
  // SHIFT: all nine taps folded into N (N = 9*NT); lane l of warp-quarter q holds strip pixel 30*q + l - 1, the
  // epilogue adds the three horizontal-tap column groups of lanes l-1, l, l+1 (valid outputs: l = 1..30).
  constexpr bool SHIFT = (FOLD == 3 && is_pro(IN)) || FOLD == 9;
  constexpr int kGroups = kWorkWarps / 8;  // worker groups (kSPro: 2, kSPro2: 1)
  constexpr uint32_t kTmemCols = IN == kSPro2 ? 256 : 512;
  constexpr bool WIDE = FOLD == 3 && IN == kSTma;
  constexpr bool RELU = !is_pro(IN);  // dense-block layers have no output ReLU; ConvBlock / decoder convs always do
  // Wide pooled layers (conv2: N = 128, four ring slots) leave the MMA warp one spare accumulator row, so MMA and
  // epilogue run back to back; both epilogue groups then drain EVERY row pair, half of the channels each.
  constexpr bool kSplitPool = WIDE && EPI == kSPool && kEG == 2;
  // Pair-granular accumulator hand-offs for the other pooled layers (conv1): the epilogue consumes and releases row pairs
  // (even slot, odd slot) as a unit, so acc_done is committed on the odd slot only and acc_free claimed / arrived on the
  // even slot only — the issuing thread's per-row hand-offs are its critical path (one N=192 MMA per row).
  constexpr bool kPairSync = EPI == kSPool && !kSplitPool;

  if (tid == 0) {
    for (int i = 0; i < kMaxSA; ++i) {
      ptx::mbar_init(&a_full[i], IN == kSTma ? 1 : (is_pro(IN) ? 8 : 4));
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&raw_full[i], 1);
      if (i < kRawStages) ptx::mbar_init(&raw_empty[i], 4);
    }
    for (int i = 0; i < P.R; ++i) {
      ptx::mbar_init(&acc_done[i], 1);
      ptx::mbar_init(&acc_free[i], kSplitPool ? 8 : 4);
    }
    ptx::mbar_init(&w_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmapA);
#ifdef CDAN_MBAR_DEBUG
    if (blockIdx.x == 0)
      printf("barriers: a_full=0x%x a_empty=0x%x raw_full=0x%x raw_empty=0x%x acc_done=0x%x acc_free=0x%x w_full=0x%x\n", ptx::smem_u32(a_full),
             ptx::smem_u32(a_empty), ptx::smem_u32(raw_full), ptx::smem_u32(raw_empty), ptx::smem_u32(acc_done), ptx::smem_u32(acc_free), ptx::smem_u32(&w_full));
#endif
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_s, kTmemCols);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < P.nchunks * 64; i += blockDim.x) {
    const bool ok = is_pro(IN) && i < P.Cin;
    s_pre_s[i] = ok ? P.pre_s[i] : 0.f;
    s_pre_t[i] = ok ? P.pre_t[i] : 0.f;
  }
  for (int i = tid; i < P.NT; i += blockDim.x) s_bias[i] = P.bias[i];
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (PAD && warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {  // ring accumulators start at zero
    const uint32_t lb = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    for (int c = 0; c < int(kTmemCols); c += 16) ptx::tmem_st16_zero(lb + c);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();

  if (warp == 0) {
    // ============================================================ producer: resident weights, then A rows via TMA.
    // The whole warp walks the loop (uniform control flow); the copies are issued by the elected lane only.
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&w_full, P.wbytes);
      for (uint32_t off = 0; off < P.wbytes; off += 32768)
        ptx::bulk_g2s(sW + off, P.wpack + off, min(32768u, P.wbytes - off), &w_full);
    }
    __syncwarp();
    if (IN == kSNchw) {
      Ring rr;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int j0 = max(it.h0 - 1, 0), j1 = min(it.h1 + 1, P.H);
        for (int j = j0; j < j1; ++j) {
          ptx::mbar_wait(&raw_empty[rr.i], (rr.w & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&raw_full[rr.i], kRawW * 3 * 4);
            ptx::tma_load_4d(s_raw + rr.i * kRawFloats, &tmapA, it.w0 - 4, j, 0, it.n, &raw_full[rr.i]);
          }
          __syncwarp();
          rr.step(kRawStages);
        }
      }
    } else {
      Ring st;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int j0 = max(it.h0 - PAD, 0), j1 = min(it.h1 + PAD, P.H);
        for (int j = j0; j < j1; ++j) {
          for (int c = 0; c < P.nchunks; ++c) {
            // kSTma: the row feeds the MMA directly; kSPro: it lands raw and the workers activate it in place
            uint64_t* full = IN == kSTma ? &a_full[st.i] : &raw_full[st.i];
            ptx::mbar_wait(&a_empty[st.i], (st.w & 1) ^ 1);
            if (lane == 0) STRACE(0, st.w * P.SA + st.i);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(full, kStage);
              uint8_t* dst = sA + size_t(st.i) * kStage;
              if (SHIFT) {
#pragma unroll
                for (int q = 0; q < 4; ++q) ptx::tma_load_4d(dst + q * 4096, &tmapA, c * 64, it.w0 - 1 + 30 * q, j, it.n, full);
              } else {
                ptx::tma_load_4d(dst, &tmapA, c * 64, it.w0 - ((WIDE || RFOLD) ? 1 : 0), j, it.n, full);
              }
            }
            __syncwarp();
            st.step(P.SA);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: uniform loop, elected lane issues
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, P.NMMA);
      const uint64_t desc_hi = ptx::umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
      const uint32_t flags = uint32_t(ptx::umma_desc_sw128(0, 1024) & 0xffffffffull);
      const uint32_t a_base = flags | ((ptx::smem_u32(sA) & 0x3FFFFu) >> 4);
      const uint32_t b_base = flags | ((ptx::smem_u32(sW) & 0x3FFFFu) >> 4);
      const uint32_t blk16 = uint32_t(P.NMMA) * 8u;  // one [NMMA x 128 B] weight block in 16-byte units
      const int klast = IN == kSNchw ? 1 : min(4, (P.Cin - (P.nchunks - 1) * 64 + 15) >> 4);
      Ring st;        // A stage
      Ring dr;        // ring slot of the accumulator row this input row's window starts at
      SlotPhases fp;  // acc_free phases
      const uint32_t a_full_u = ptx::smem_u32(a_full), a_empty_u = ptx::smem_u32(a_empty), acc_done_u = ptx::smem_u32(acc_done),
                     acc_free_u = ptx::smem_u32(acc_free);
      ptx::mbar_wait(&w_full, 0);
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int n_in = it.h1 - it.h0 + 2 * PAD;
        dr.i = it.h0 % P.R;  // accumulator row 0 of the segment = image row h0 - 2*PAD
        if (PAD) {           // the first input row also opens the segment's first two accumulator rows
          fp.claim_a(acc_free_u, dr.i);
          if (!kPairSync) fp.claim_a(acc_free_u, dr.i + 1 == P.R ? 0 : dr.i + 1);
        }
        for (int jj = 0; jj < n_in; ++jj) {
          const int j = it.h0 - PAD + jj;
          {
            int newest = dr.i + 2 * PAD;  // newest accumulator row this input row touches
            if (newest >= P.R) newest -= P.R;
            if (!kPairSync || !(newest & 1)) fp.claim_a(acc_free_u, newest);
          }
          ptx::tc_fence_after_sync();
          if (lane == 0) STRACE(7, dr.i);
          if (j >= 0 && j < P.H) {
            const uint32_t dcol = tmem_base + uint32_t(dr.i * P.SW);
            uint32_t b0 = b_base;
            for (int c = 0; c < P.nchunks; ++c) {
              const int ksteps = c == P.nchunks - 1 ? klast : 4;
              ptx::mbar_wait_a(a_full_u + 8u * uint32_t(st.i), st.w & 1);
              ptx::tc_fence_after_sync();
              if (lane == 0) STRACE(3, st.w * P.SA + st.i);
              const uint32_t a0 = a_base + uint32_t(st.i) * (kStage >> 4);
              const uint32_t acc0 = PAD ? 1u : (c != 0 ? 1u : 0u);
              // One elected lane issues (ptxas keeps descriptors in uniform registers inside an elect.sync region;
              // a per-thread predicate on the instruction instead costs a vote + R2UR.BROADCAST sequence per MMA).
              if (ptx::elect_one()) {
                if (P.ablate & 4) {
                } else if (WIDE) {
                  // nine taps: kernel row r feeds accumulator row G + (2 - r) (its own ring slot), the horizontal tap
                  // s is the A view shifted by s pixels; weight blocks are ordered [r*3+s]
#pragma unroll
                  for (int pos = 0; pos < 3; ++pos) {
                    int sl = dr.i + pos;
                    if (sl >= P.R) sl -= P.R;
                    const uint32_t dc = tmem_base + uint32_t(sl * P.SW);
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                      const uint32_t bb = b0 + uint32_t((2 - pos) * 3 + s) * blk16;
#pragma unroll
                      for (int k = 0; k < 4; ++k)
                        if (k < ksteps)
                          ptx::umma_bf16(dc, desc_hi | (a0 + uint32_t(8 * s + 2 * k)), desc_hi | (bb + uint32_t(2 * k)), idesc, 1u);
                    }
                  }
                } else if (RFOLD) {
#pragma unroll
                  for (int s = 0; s < 3; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      if (k < ksteps)
                        ptx::umma_bf16(dcol, desc_hi | (a0 + uint32_t(8 * s + 2 * k)), desc_hi | (b0 + uint32_t(s) * blk16 + uint32_t(2 * k)), idesc, 1u);
                } else {
#pragma unroll
                  for (int k = 0; k < (IN == kSNchw ? 1 : 4); ++k)
                    if (k < ksteps)
                      ptx::umma_bf16(dcol, desc_hi | (a0 + uint32_t(2 * k)), desc_hi | (b0 + uint32_t(2 * k)), idesc, k == 0 ? acc0 : 1u);
                }
                ptx::umma_commit_a(a_empty_u + 8u * uint32_t(st.i));
              }
              if (lane == 0) STRACE(4, st.w * P.SA + st.i);
              __syncwarp();
              st.step(P.SA);
              b0 += WIDE ? 9u * blk16 : (RFOLD ? 3u * blk16 : blk16);
            }
          }
          if (!kPairSync || (dr.i & 1)) {
            if (ptx::elect_one()) ptx::umma_commit_a(acc_done_u + 8u * uint32_t(dr.i));
            __syncwarp();
          }
          dr.step(P.R);
        }
        if (PAD) {  // the two trailing accumulator rows of the segment receive no further input
          if (!kPairSync || (dr.i & 1)) {
            if (ptx::elect_one()) ptx::umma_commit_a(acc_done_u + 8u * uint32_t(dr.i));
            __syncwarp();
          }
          dr.step(P.R);
          if (!kPairSync || (dr.i & 1)) {
            if (ptx::elect_one()) ptx::umma_commit_a(acc_done_u + 8u * uint32_t(dr.i));
            __syncwarp();
          }
          dr.step(P.R);
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ============================================================ epilogue
    const int eg = (warp - kEpiWarp0) >> 2, q = warp & 3;
    const uint32_t lb = tmem_base + (uint32_t(q * 32) << 16);
    const int px = q * 32 + lane;
    SlotPhases dp;  // acc_done phases
    int pairs_seen = 0;  // running count of pooled row pairs (they alternate between the epilogue groups)
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int n_acc = it.h1 - it.h0 + 4 * PAD;
      const int col = SHIFT ? it.w0 - 1 + 30 * q + lane : it.w0 + px;
      const bool col_ok = SHIFT ? (lane >= 1 && lane <= 30 && col < P.W && col < it.w0 + P.TW) : (px < P.TW && col < P.W);
      if (EPI != kSPool) {
        // segments start at even image rows and R is even: accumulator row ii sits in a slot of parity (ii & 1), so a
        // slot always belongs to the same epilogue group
        const int ii0 = kEG == 2 ? eg : 0;
        Ring sr;
        sr.i = it.h0 % P.R;
        sr.add(ii0, P.R);
        int i = it.h0 - 2 * PAD + ii0;
        // output address of (n, i, col): advanced by kEG rows per step
        const size_t row_elems = EPI == kSNchwOut ? size_t(P.W) : size_t(P.W) * P.out_ld;
        bf16* o_b = nullptr;
        float* o_f = nullptr;
        if (EPI == kSNchwOut) o_f = P.out_nchw + (size_t(it.n) * P.Cout * P.H + i) * P.W + col;
        else o_b = P.out + ((size_t(it.n) * P.H + i) * P.W + col) * P.out_ld;
        for (int ii = ii0; ii < n_acc; ii += kEG) {
          const int slot = sr.i;
          dp.wait(acc_done, slot);
          ptx::tc_fence_after_sync();
          if (q == 0 && lane == 0) STRACE(5, sr.i);
          const bool row_ok = i >= it.h0 && i < it.h1 && !(P.ablate & 8);
          const bool shadow = PAD && !WIDE && slot < 2;
          const uint32_t tm = lb + uint32_t(slot * P.SW), ts = lb + uint32_t((P.R + slot) * P.SW);
          // 8 output channels at a time keeps the live register set small (spills are L2 round trips here: with
          // ~224 KB of shared memory in use the L1 has almost no capacity left)
          for (int c0 = 0; c0 < P.NT; c0 += 8) {
            uint32_t v[8];
            if (SHIFT) {
              // slot = [tap s=0 | s=1 | s=2] x 16 channels; out(l) = Y0(l-1) + Y1(l) + Y2(l+1)
              uint32_t y0[8], y2[8];
              if (row_ok) {
                ptx::tmem_ld8(tm + c0, y0);
                ptx::tmem_ld8(tm + 16 + c0, v);
                ptx::tmem_ld8(tm + 32 + c0, y2);
                if (shadow) {
                  uint32_t w0[8], w1[8], w2[8];
                  ptx::tmem_ld8(ts + c0, w0);
                  ptx::tmem_ld8(ts + 16 + c0, w1);
                  ptx::tmem_ld8(ts + 32 + c0, w2);
                  ptx::tmem_wait_ld();
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    y0[e] = __float_as_uint(__uint_as_float(y0[e]) + __uint_as_float(w0[e]));
                    v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(w1[e]));
                    y2[e] = __float_as_uint(__uint_as_float(y2[e]) + __uint_as_float(w2[e]));
                  }
                } else {
                  ptx::tmem_wait_ld();
                }
              }
              ptx::tmem_st8_zero(tm + c0); ptx::tmem_st8_zero(tm + 16 + c0); ptx::tmem_st8_zero(tm + 32 + c0);
              if (shadow) { ptx::tmem_st8_zero(ts + c0); ptx::tmem_st8_zero(ts + 16 + c0); ptx::tmem_st8_zero(ts + 32 + c0); }
              if (row_ok) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float a0 = __shfl_up_sync(0xffffffffu, __uint_as_float(y0[e]), 1);
                  const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(y2[e]), 1);
                  v[e] = __float_as_uint(a0 + __uint_as_float(v[e]) + a2);
                }
              }
            } else {
              if (row_ok) {
                ptx::tmem_ld8(tm + c0, v);
                if (shadow) {
                  uint32_t v2[8];
                  ptx::tmem_ld8(ts + c0, v2);
                  ptx::tmem_wait_ld();
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
                } else {
                  ptx::tmem_wait_ld();
                }
              }
              if (PAD) {
                ptx::tmem_st8_zero(tm + c0);
                if (shadow) ptx::tmem_st8_zero(ts + c0);
              }
            }
            if (row_ok) {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
              float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y, __uint_as_float(v[2]) + b0.z,
                            __uint_as_float(v[3]) + b0.w, __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                            __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
              if (RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
              if (col_ok && !(P.ablate & 1)) {
                if (EPI == kSNchwOut) {
                  const size_t plane = size_t(P.H) * P.W;
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    if (c0 + e < P.Cout) o_f[size_t(c0 + e) * plane] = P.sigmoid ? 1.0f / (1.0f + __expf(-f[e])) : f[e];
                } else if (c0 < P.Cout) {  // channel slices are padded to multiples of 8
                  *reinterpret_cast<uint4*>(o_b + c0) = make_uint4(bf2(f[0], f[1]), bf2(f[2], f[3]), bf2(f[4], f[5]), bf2(f[6], f[7]));
                }
              }
            }
          }
          if (PAD) ptx::tmem_wait_st();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_free[slot]);
          if (q == 0 && lane == 0) STRACE(6, sr.i);
          sr.add(kEG, P.R);
          i += kEG;
          if (EPI == kSNchwOut) o_f += kEG * row_elems; else o_b += kEG * row_elems;
        }
      } else {
        // 2x2 max-pool: accumulator rows (a, a+1) <-> image rows (i, i+1), i even; horizontal partner = lane ^ 1.
        // Row pairs alternate between the two epilogue groups.  (A is even at every item start: segments are even.)
        // every group walks all pairs (to keep every slot's phase bit current) but only drains its own
        Ring sr;
        sr.i = it.h0 % P.R;
        int i = it.h0 - 2 * PAD;
        for (int ii = 0; ii < n_acc; ii += 2, ++pairs_seen, sr.add(2, P.R), i += 2) {
          const int sl0 = sr.i, sl1 = sr.i + 1;  // h0 and R are even, so a pair never straddles the ring end
          if (!kSplitPool && kEG == 2 && (pairs_seen & 1) != eg) {
            if (!kPairSync) dp.skip(sl0);
            dp.skip(sl1);
            continue;
          }
          if (kSplitPool) {
            // Row-wise drain (wide pooled layers have one spare ring slot): the first row of the pair is read as soon as
            // it is complete, kept as bf16(relu(v + bias)) in registers and its slot released at once; the second row is
            // combined with it.  max() commutes with the monotone bias/ReLU/rounding chain, so the result is bitwise the
            // same as pooling first.  Both groups drain every row, half of the channels each.
            const bool row_ok = i >= it.h0 && i < it.h1;
            const int cbeg = eg * (P.NT >> 1);
            const bool odd = lane & 1;
            bf16* o_row = P.out + ((size_t(it.n) * (P.H >> 1) + (i >> 1)) * (P.W >> 1) + (col >> 1)) * P.out_ld;
            uint32_t held[4][8];  // up to 64 channels of row 0 as bf16 pairs
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int sl = rr ? sl1 : sl0;
              dp.wait(acc_done, sl);
              ptx::tc_fence_after_sync();
              if (rr == 0 && q == 0 && lane == 0) STRACE(5, sr.i >> 1);
              const uint32_t tr = lb + uint32_t(sl * P.NT);
#pragma unroll
              for (int ci = 0; ci < 4; ++ci) {
                const int c0 = cbeg + 16 * ci;
                if (16 * ci >= (P.NT >> 1)) break;
                uint32_t v[16];
                if (row_ok) {
                  ptx::tmem_ld16(tr + c0, v);
                  ptx::tmem_wait_ld();
                }
                ptx::tmem_st16_zero(tr + c0);
                if (!row_ok) continue;
                uint32_t y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float2 bb = *reinterpret_cast<const float2*>(s_bias + c0 + 2 * e);
                  float f0 = __uint_as_float(v[2 * e]) + bb.x, f1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                  if (RELU) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                  y[e] = bf2(f0, f1);
                }
                if (rr == 0) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) held[ci][e] = y[e];
                } else {
                  // vertical max, then the horizontal partner (lane ^ 1): even lane keeps channels [c0, c0+8), odd lane
                  // [c0+8, c0+16) of the pooled pixel
                  uint32_t m[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 lo = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&y[e]), *reinterpret_cast<const __nv_bfloat162*>(&held[ci][e]));
                    const __nv_bfloat162 hi = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&y[4 + e]), *reinterpret_cast<const __nv_bfloat162*>(&held[ci][4 + e]));
                    const uint32_t lo_u = *reinterpret_cast<const uint32_t*>(&lo), hi_u = *reinterpret_cast<const uint32_t*>(&hi);
                    const uint32_t mine = odd ? hi_u : lo_u, theirs = odd ? lo_u : hi_u;
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, theirs, 1);
                    const __nv_bfloat162 mm = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&mine), *reinterpret_cast<const __nv_bfloat162*>(&got));
                    m[e] = *reinterpret_cast<const uint32_t*>(&mm);
                  }
                  const int cg = c0 + (odd ? 8 : 0);
                  if (col_ok && cg < P.Cout && !(P.ablate & 1)) *reinterpret_cast<uint4*>(o_row + cg) = make_uint4(m[0], m[1], m[2], m[3]);
                }
              }
              ptx::tmem_wait_st();
              ptx::tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&acc_free[sl]);
            }
            if (q == 0 && lane == 0) STRACE(6, sr.i >> 1);
            continue;
          }
          if (!kPairSync) dp.wait(acc_done, sl0);
          dp.wait(acc_done, sl1);
          ptx::tc_fence_after_sync();
          if (q == 0 && lane == 0) STRACE(5, sr.i >> 1);
          const bool row_ok = i >= it.h0 && i < it.h1;
          const bool sh0 = !WIDE && sl0 < 2, sh1 = !WIDE && sl1 < 2;
          const uint32_t t0 = lb + uint32_t(sl0 * P.NT), t1 = lb + uint32_t(sl1 * P.NT);
          const uint32_t ts0 = lb + uint32_t((P.R + sl0) * P.NT), ts1 = lb + uint32_t((P.R + sl1) * P.NT);
          bf16* o_row = P.out + ((size_t(it.n) * (P.H >> 1) + (i >> 1)) * (P.W >> 1) + (col >> 1)) * P.out_ld;
          const int cbeg = kSplitPool ? eg * (P.NT >> 1) : 0, cend = kSplitPool ? cbeg + (P.NT >> 1) : P.NT;
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            // 16 channels of both rows per TMEM round trip (pooling kernels run with <= 512 threads, i.e. 128 registers)
            uint32_t v0[16], v1[16];
            if (row_ok) {
              ptx::tmem_ld16(t0 + c0, v0);
              ptx::tmem_ld16(t1 + c0, v1);
              if (sh0 || sh1) {
                uint32_t w0[16], w1[16];
                if (sh0) ptx::tmem_ld16(ts0 + c0, w0);
                if (sh1) ptx::tmem_ld16(ts1 + c0, w1);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  if (sh0) v0[e] = __float_as_uint(__uint_as_float(v0[e]) + __uint_as_float(w0[e]));
                  if (sh1) v1[e] = __float_as_uint(__uint_as_float(v1[e]) + __uint_as_float(w1[e]));
                }
              } else {
                ptx::tmem_wait_ld();
              }
            }
            ptx::tmem_st16_zero(t0 + c0);
            ptx::tmem_st16_zero(t1 + c0);
            if (sh0) ptx::tmem_st16_zero(ts0 + c0);
            if (sh1) ptx::tmem_st16_zero(ts1 + c0);
            if (row_ok) {
              // even lane keeps channels [c0, c0+8), odd lane [c0+8, c0+16) of the pooled pixel: exchange the other half
              // with the horizontal partner and take the max (max commutes with the per-channel bias and with ReLU)
              const bool odd = lane & 1;
              float m[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float lo = fmaxf(__uint_as_float(v0[e]), __uint_as_float(v1[e]));
                const float hi = fmaxf(__uint_as_float(v0[8 + e]), __uint_as_float(v1[8 + e]));
                const float mine = odd ? hi : lo, theirs = odd ? lo : hi;
                const float got = __shfl_xor_sync(0xffffffffu, theirs, 1);  // partner's value of MY channel half
                m[e] = fmaxf(mine, got);
              }
              const int cg = c0 + (odd ? 8 : 0);
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cg), b1 = *reinterpret_cast<const float4*>(s_bias + cg + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                m[e] += bb[e];
                if (RELU) m[e] = fmaxf(m[e], 0.f);
              }
              if (col_ok && cg < P.Cout)
                *reinterpret_cast<uint4*>(o_row + cg) = make_uint4(bf2(m[0], m[1]), bf2(m[2], m[3]), bf2(m[4], m[5]), bf2(m[6], m[7]));
            }
          }
          ptx::tmem_wait_st();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(&acc_free[sl0]);
            if (!kPairSync) ptx::mbar_arrive(&acc_free[sl1]);
          }
          if (q == 0 && lane == 0) STRACE(6, sr.i >> 1);
        }
      }
    }
  } else if (warp >= kWorkWarp0) {
    // ============================================================ A-stage workers
    const int aw = warp - kWorkWarp0;
    if (is_pro(IN)) {
      // The raw NHWC bf16 row was delivered by TMA (zero outside the image / beyond Cin).  Apply the dense-block
      // pre-activation relu(s*x+t) in place in fp32; pixels outside the image stay zero, i.e. the conv padding is applied
      // AFTER the activation (reference models/cdan.py:41-46).  Thread -> (16-byte channel group u, pixels qb + 32*i).
      // All four loads are issued before any arithmetic and the four stores follow (the explicit ld/st.shared are
      // volatile asm and would otherwise serialise load -> math -> store per pixel).
      // Two groups of eight warps take alternate stages.
      const int grp = aw >> 3, t = (aw & 7) * 32 + lane, u = t & 7, qb = t >> 3;
      const uint32_t sA_u = ptx::smem_u32(sA);
      uint32_t off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) off[i] = ptx::sw128_offset(uint32_t(qb + 32 * i), uint32_t(u));
      __nv_bfloat162 sc[4], sh[4];
      int cached_c = -1;
      Ring st;
      int turn = 0;  // stage parity: this group works when turn == grp
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const Item it = decode_item(P, item);
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = SHIFT ? it.w0 - 1 + 30 * i + qb : it.w0 - (RFOLD ? 1 : 0) + qb + 32 * i;
          ok[i] = col >= 0 && col < P.W;
        }
        const int j0 = max(it.h0 - PAD, 0), j1 = min(it.h1 + PAD, P.H);
        for (int j = j0; j < j1; ++j) {
          for (int c = 0; c < P.nchunks; ++c, turn ^= 1, st.step(P.SA)) {
            if (kGroups == 2 && turn != grp) continue;
            if (c != cached_c) {
              // BN scale / shift of this thread's 8 channels as packed bf16x2 (rounded once here): the activation is
              // one HFMA2.BF16 with ReLU per channel pair instead of unpack + 2 FFMA + pack
              const float4 fs0 = *reinterpret_cast<const float4*>(s_pre_s + c * 64 + u * 8);
              const float4 fs1 = *reinterpret_cast<const float4*>(s_pre_s + c * 64 + u * 8 + 4);
              const float4 ft0 = *reinterpret_cast<const float4*>(s_pre_t + c * 64 + u * 8);
              const float4 ft1 = *reinterpret_cast<const float4*>(s_pre_t + c * 64 + u * 8 + 4);
              sc[0] = __floats2bfloat162_rn(fs0.x, fs0.y); sc[1] = __floats2bfloat162_rn(fs0.z, fs0.w);
              sc[2] = __floats2bfloat162_rn(fs1.x, fs1.y); sc[3] = __floats2bfloat162_rn(fs1.z, fs1.w);
              sh[0] = __floats2bfloat162_rn(ft0.x, ft0.y); sh[1] = __floats2bfloat162_rn(ft0.z, ft0.w);
              sh[2] = __floats2bfloat162_rn(ft1.x, ft1.y); sh[3] = __floats2bfloat162_rn(ft1.z, ft1.w);
              cached_c = c;
            }
            const bool active = u * 8 < min(64, P.Cin - c * 64);
            ptx::mbar_wait(&raw_full[st.i], st.w & 1);
            if ((aw & 7) == 0 && lane == 0) STRACE(1, st.w * P.SA + st.i);
            if (active && !(P.ablate & 2)) {
              const uint32_t base = sA_u + uint32_t(st.i) * kStage;
              uint4 r[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) r[i] = ptx::lds128(base + off[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162* v = reinterpret_cast<__nv_bfloat162*>(&r[i]);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = __hfma2_relu(v[e], sc[e], sh[e]);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (ok[i]) ptx::sts128(base + off[i], r[i]);  // out-of-image pixels keep TMA's zero fill
            }
            if (!(P.ablate & 16)) ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&a_full[st.i]);
            if ((aw & 7) == 0 && lane == 0) STRACE(2, st.w * P.SA + st.i);
          }
        }
      }
    } else if (IN == kSNchw) {
      // encoder.conv1: the planar fp32 row (3 channels x 136 pixels from column w0-4, zero filled outside the image by TMA) waits in the
      // raw ring; build the K-folded bf16 row A[p][s*3+ci] = x[ci][j][w0+p-1+s], one thread per strip pixel.
      const int p = aw * 32 + lane;
      const uint32_t sA_u = ptx::smem_u32(sA);
      Ring st, rr;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int j0 = max(it.h0 - 1, 0), j1 = min(it.h1 + 1, P.H);
        for (int j = j0; j < j1; ++j) {
          ptx::mbar_wait(&raw_full[rr.i], rr.w & 1);
          if (aw == 0 && lane == 0) STRACE(1, st.w * P.SA + st.i);
          const float* raw = s_raw + rr.i * kRawFloats + p + 3;
          float x[9];
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) x[s * 3 + ci] = raw[ci * kRawW + s];
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&raw_empty[rr.i]);
          ptx::mbar_wait(&a_empty[st.i], (st.w & 1) ^ 1);
          const uint32_t base = sA_u + uint32_t(st.i) * kStage;
          ptx::sts128(base + ptx::sw128_offset(uint32_t(p), 0),
                      make_uint4(bf2(x[0], x[1]), bf2(x[2], x[3]), bf2(x[4], x[5]), bf2(x[6], x[7])));
          ptx::sts128(base + ptx::sw128_offset(uint32_t(p), 1), make_uint4(bf2(x[8], 0.f), 0u, 0u, 0u));
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[st.i]);
          if (aw == 0 && lane == 0) STRACE(2, st.w * P.SA + st.i);
          st.step(P.SA);
          rr.step(kRawStages);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ two rows per stage
// Dense-block layers (pre-activated input; 3x3 in the row-fold form or 1x1) with TWO image rows per pipeline stage.
// Every hand-off in this pipeline (TMA issue, raw_full -> workers, a_full -> MMA, commits, acc_done -> epilogue,
// acc_free -> MMA) is a single-warp latency of 150-200 cycles, and with one row per stage those hand-offs alone cost
// ~930 cycles per row (measured by ablation: all arithmetic, MMAs, TMEM traffic and stores removed).  Here one stage is
// 32 KB = rows (j, j+1) of the strip, accumulator rows are signalled and recycled in pairs, and the two epilogue groups
// each own one row of the pair — half the hand-offs per row, same data path.
//
// GP = group-planar input: the concat buffer is stored as one dense plane [N][H][W][16] per 16-channel group, so a strip
// row of one group is 4 KB of CONTIGUOUS memory.  A pixel line of the NHWC layout (32-128 useful bytes every 256) costs
// the L2 one request per pixel and K-chunk whatever its length, which bounded the NHWC form at ~7 cycles per pixel and
// SM; contiguous 32-byte lines merge into full 128-byte requests (profiles/r01_bulk_probe.log: 4.3-6.3 TB/s for one to
// five groups).  One group = one K=16 MMA step: stage = [group][row][128 px x 32 B] in SWIZZLE_32B, the A descriptor of
// tap s starts s pixels (s * 32 B) into the row.
// GP = 2 (hybrid): the buffer's head (the pooled ConvBlock output, P.nh 64-channel chunks) stays NHWC because the next
// ConvBlock, the decoder skip connection and the stage taps read it as such; only the 16-channel groups the dense layers
// append are planes.  A row pair then takes P.nh NHWC stages (tmapA) followed by planar stages of up to four groups
// (tmapG); the weight image is indexed by 64-channel chunk either way.
template <int FOLD, int EPI, int GP>
__global__ void __launch_bounds__(896, 1) conv_stream2_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                              const __grid_constant__ CUtensorMap tmapG, const SParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t kStage2 = GP == 1 ? uint32_t(P.stage_bytes) : uint32_t(2 * kStage);
  constexpr uint32_t kGroupBytes = 8192, kGroupRow = 4096;  // GP: two rows of one 16-channel group / one row
  uint8_t* sA = smem;
  uint8_t* sW = sA + size_t(P.SA) * kStage2 + 1024;
  float* s_pre_s = reinterpret_cast<float*>(sW + P.wbytes);
  float* s_pre_t = s_pre_s + P.nchunks * 64;
  float* s_bias = s_pre_t + P.nchunks * 64;
  // BatchNorm scale / shift as bf16 tables (packed HFMA2 operands of the planar worker path), behind the bias vector
  __nv_bfloat16* s_sc = reinterpret_cast<__nv_bfloat16*>(s_bias + 128);
  __nv_bfloat16* s_sh = s_sc + P.nchunks * 64;
  // stage c of a row pair is planar?  (GP 0: never, 1: always, 2: after the NHWC head)
  auto planar_stage = [&](int c) { return GP == 1 || (GP == 2 && c >= P.nh); };

  __shared__ uint64_t a_full[kMaxSA2], a_empty[kMaxSA2], raw_full[kMaxSA2], acc_done[kMaxR / 2], acc_free[kMaxR / 2], w_full, mma_turn[3];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool RFOLD = FOLD == 4;
  constexpr int PAD = RFOLD ? 1 : 0;
  constexpr int kEpiWarps = 8, kWorkWarp0 = kEpiWarp0 + kEpiWarps;

  if (tid == 0) {
    for (int i = 0; i < kMaxSA2; ++i) {
      ptx::mbar_init(&a_full[i], P.wsplit ? 16 : 8);
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&raw_full[i], 1);
    }
    for (int i = 0; i < kMaxR / 2; ++i) {
      ptx::mbar_init(&acc_done[i], 1);
      ptx::mbar_init(&acc_free[i], 8);
    }
    ptx::mbar_init(&w_full, 1);
    ptx::mbar_init(&mma_turn[0], 1);
    ptx::mbar_init(&mma_turn[1], 1);
    ptx::mbar_init(&mma_turn[2], 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmapA);
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < P.nchunks * 64; i += blockDim.x) {
    const bool ok = i < P.Cin;
    const float fs = ok ? P.pre_s[i] : 0.f, ft = ok ? P.pre_t[i] : 0.f;
    if (GP != 0) {
      s_sc[i] = __float2bfloat16_rn(fs);
      s_sh[i] = __float2bfloat16_rn(ft);
    }
    if (GP != 1) {
      s_pre_s[i] = fs;
      s_pre_t[i] = ft;
    }
  }
  for (int i = tid; i < P.NT; i += blockDim.x) s_bias[i] = P.bias[i];
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (PAD && warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {  // ring accumulators start at zero
    const uint32_t lb = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    for (int c = 0; c < 512; c += 16) ptx::tmem_st16_zero(lb + c);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();

  if (warp == 0) {
    // ============================================================ producer
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&w_full, P.wbytes);
      for (uint32_t off = 0; off < P.wbytes; off += 32768)
        ptx::bulk_g2s(sW + off, P.wpack + off, min(32768u, P.wbytes - off), &w_full);
    }
    __syncwarp();
    Ring st;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int npairs = (it.h1 - it.h0 + 2 * PAD + 1) >> 1;
      for (int pp = 0; pp < npairs; ++pp) {
        const int j = it.h0 - PAD + 2 * pp;  // rows j, j+1: outside the image -> zero filled by TMA
        for (int c = 0; c < P.npc; ++c) {
          ptx::mbar_wait(&a_empty[st.i], (st.w & 1) ^ 1);
          if (ptx::elect_one()) {
            if (planar_stage(c)) {
              const int g0 = P.gps * (c - P.nh);  // first group of this stage (groups count from the end of the head)
              const int ng = min(P.gps, P.ngroups - g0);
              ptx::mbar_arrive_expect_tx(&raw_full[st.i], uint32_t(ng) * kGroupBytes);
              for (int g = 0; g < ng; ++g)
                ptx::tma_load_4d(sA + size_t(st.i) * kStage2 + size_t(g) * kGroupBytes, GP == 2 ? &tmapG : &tmapA, 0,
                                 it.w0 - PAD, j, it.n + (g0 + g) * P.N, &raw_full[st.i]);
            } else {
              ptx::mbar_arrive_expect_tx(&raw_full[st.i], kStage2);
              ptx::tma_load_4d(sA + size_t(st.i) * kStage2, &tmapA, c * 64, it.w0 - PAD, j, it.n, &raw_full[st.i]);
            }
            STRACE(0, st.w * P.SA + st.i);
          }
          __syncwarp();
          st.step(P.SA);
        }
      }
    }
  } else if (warp >= 1 && warp <= 3 && (warp != 2 || P.ni > 2)) {
    // ============================================================ MMA issuers (P.ni = 2 or 3 warps, row pairs round-robin)
    // One issuing thread spends ~1100 cycles per row pair on hand-offs (two mbarrier waits, the tcgen05 fence, three
    // commits) during which the tensor pipe drains: it queues only a few MMAs, so issue time and hand-off time ADD
    // (timeline traces in profiles/r01_trace_stream2.md).  Two issuers take alternate row pairs; the order of the
    // accumulations is kept by a token: an issuer starts its pair only after the other's MMAs have COMPLETED
    // (tcgen05.commit on mma_turn), so its hand-offs overlap the other's MMAs and results stay bitwise deterministic.
    // With three issuers (warp 2 joins after allocating TMEM) an issuer's loop has three pair times to complete, and a
    // turn is P.pt consecutive pairs so that the hand-off latency (MMA completion + wake-up + fence) is paid once per
    // P.pt pairs.
    const int NI = P.ni, PT = P.pt, HIST = (NI - 1) * PT;  // HIST (<= 4) = pairs the other issuers may still be working on
    const int mw = warp == 1 ? 0 : (warp == 3 ? 1 : 2);
    const int mw_next = mw + 1 == NI ? 0 : mw + 1;
    const uint32_t idesc = ptx::umma_idesc_bf16(128, P.NMMA);
    const uint64_t desc_hi = ptx::umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
    const uint32_t flags = uint32_t(ptx::umma_desc_sw128(0, 1024) & 0xffffffffull);
    const uint64_t desc32_hi = ptx::umma_desc_sw32(0, 256) & 0xffffffff00000000ull;
    const uint32_t a_base = flags | ((ptx::smem_u32(sA) & 0x3FFFFu) >> 4);
    const uint32_t b_base = flags | ((ptx::smem_u32(sW) & 0x3FFFFu) >> 4);
    const uint32_t blk16 = uint32_t(P.NMMA) * 8u;
    const int klast = min(4, (P.Cin - (P.nchunks - 1) * 64 + 15) >> 4);
    Ring st, dr;    // stage; ring slot of the first accumulator row of the current pair (tied to the absolute image row)
    SlotPhases fp;  // acc_free phases, one bit per slot pair (both issuers track every slot)
    uint32_t gp = 0, tok = 0;  // global pair sequence number; tokens consumed by this issuer
    uint32_t pm1 = 0, pm2 = 0, pm3 = 0, pm4 = 0;  // accumulator slot pairs claimed for pairs gp-1 .. gp-4
    int owner = 0, tpos = 0;                      // issuer of pair gp, position of gp inside that issuer's turn
    bool nx_claimed = false, nx_waited = false;   // PT = 2: the turn's second pair was claimed / its stages awaited early
    auto next_pair = [&](uint32_t mask) {
      pm4 = pm3; pm3 = pm2; pm2 = pm1; pm1 = mask;
      ++gp;
      if (++tpos == PT) { tpos = 0; owner = owner + 1 == NI ? 0 : owner + 1; }
    };
    const uint32_t a_full_u = ptx::smem_u32(a_full), a_empty_u = ptx::smem_u32(a_empty), acc_done_u = ptx::smem_u32(acc_done),
                   acc_free_u = ptx::smem_u32(acc_free), my_turn_u = ptx::smem_u32(&mma_turn[mw]),
                   next_turn_u = ptx::smem_u32(&mma_turn[mw_next]);
    ptx::mbar_wait(&w_full, 0);
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int npairs = (it.h1 - it.h0 + 2 * PAD + 1) >> 1;
      dr.i = it.h0 % P.R;  // even: segments start at even rows, R is even
      for (int pp = 0; pp < npairs; ++pp) {
        const bool mine = owner == mw;
        int newest = dr.i + 2 * PAD;  // newest accumulator pair this input pair touches
        if (newest >= P.R) newest -= P.R;
        // slot pairs this input pair opens: the newest one, and at a segment start (3x3) also the segment's first
        const int s1 = newest >> 1, s0 = (PAD && pp == 0) ? (dr.i >> 1) : s1;
        const uint32_t cur_mask = (1u << s1) | (1u << s0);
        if (!mine) {
          fp.note(s1);
          if (s0 != s1) fp.note(s0);
          for (int c = 0; c < P.npc; ++c) st.step(P.SA);
          dr.add(2, P.R);
          next_pair(cur_mask);
          continue;
        }
        const bool first = tpos == 0, last = tpos == PT - 1;
        const bool need_token = first && gp > 0;
        // A slot may be claimed ahead of the token only if its previous use is older than the other issuers' current
        // pairs: otherwise that use's own claim may still be pending and a parity wait one phase ahead returns a false
        // positive (ring positions jump at segment starts, so consecutive pairs can meet in one slot).
        // (Later pairs of a turn follow the token: everything before them is complete or this issuer's own.)
        const uint32_t recent = pm1 | (HIST > 1 ? pm2 : 0u) | (HIST > 2 ? pm3 : 0u) | (HIST > 3 ? pm4 : 0u);
        const uint32_t deferred = need_token ? (cur_mask & recent) : 0u;
        const bool pre_c = !first && nx_claimed, pre_w = !first && nx_waited;  // done while waiting for this turn's token
        nx_claimed = nx_waited = false;
        if (!pre_c) {
          if (!((deferred >> s1) & 1u)) fp.claim_a(acc_free_u, s1);
          if (s0 != s1 && !((deferred >> s0) & 1u)) fp.claim_a(acc_free_u, s0);
        }
        if (PT == 2 && first && pp + 1 < npairs) {
          // Two pairs per turn: the second pair's accumulator slot and stages are serviced here, ahead of the token, under
          // the same two rules (slot not claimed for the other issuers' current pairs; previous use of its stages older
          // than those pairs) — behind the token the turn is then pure MMA issue for both pairs.
          int nn = dr.i + 2 + 2 * PAD;
          while (nn >= P.R) nn -= P.R;
          const uint32_t mn = 1u << (nn >> 1);
          if (!need_token || !(mn & (recent | cur_mask))) {
            fp.claim_a(acc_free_u, nn >> 1);
            nx_claimed = true;
          }
          if (!need_token || P.SA >= (HIST + PT) * P.npc) {
            Ring sx = st;
            for (int c = 0; c < P.npc; ++c) sx.step(P.SA);
            for (int c = 0; c < P.npc; ++c, sx.step(P.SA)) ptx::mbar_wait_a(a_full_u + 8u * uint32_t(sx.i), sx.w & 1);
            nx_waited = true;
          }
        }
        if (mw == 0 && lane == 0) STRACE(7, st.w * P.SA + st.i);
        // With enough stages (SA >= NI*PT stages-per-pair: the previous use of every stage of this pair lies before the
        // other issuers' current pairs) ALL stages of the pair are waited for ahead of the token, so that the chunk loop
        // behind the token is pure MMA issue.
        const bool early_all = need_token && P.npc > 1 && P.SA >= (HIST + PT) * P.npc;
        if (early_all) {
          Ring sx = st;
          for (int c = 0; c < P.npc; ++c, sx.step(P.SA)) ptx::mbar_wait_a(a_full_u + 8u * uint32_t(sx.i), sx.w & 1);
        }
        uint32_t b0 = b_base;
        for (int c = 0; c < P.npc; ++c) {
          const bool pl = planar_stage(c);
          const int ksteps = pl ? min(P.gps, P.ngroups - P.gps * (c - P.nh)) : (GP == 2 || c != P.npc - 1 ? 4 : klast);
          const uint64_t adesc_hi = pl ? desc32_hi : desc_hi;
          const uint32_t a_row = pl ? (kGroupRow >> 4) : uint32_t(kStage >> 4);  // descriptor units (16 B) per image row
          const uint32_t a_tap = pl ? 2u : 8u, a_k = pl ? (kGroupBytes >> 4) : 2u;  // per horizontal tap (one pixel) / per K=16 step
          // The first stage of a pair may be waited for ahead of the token only if its previous use lies before the
          // other issuers' current pairs (SA > HIST * nchunks): otherwise that use may not even be filled yet and a
          // parity wait one phase ahead returns a false positive.
          const bool late = c == 0 && need_token && P.SA <= HIST * P.npc;
          if (!late && !early_all && !pre_w) ptx::mbar_wait_a(a_full_u + 8u * uint32_t(st.i), st.w & 1);
          // Common case: the elected lane alone waits for the token, AFTER its descriptors are set up (the wait is the
          // hand-off chain's critical path, everything hoisted above it is free).
          const bool tok_inside = c == 0 && need_token && !late && deferred == 0u && P.tok_inside;
          if (c == 0 && need_token && !tok_inside) {  // the previous issuer's turn has completed
            ptx::mbar_wait_a(my_turn_u, tok & 1u);
            if (deferred) {
              if ((deferred >> s1) & 1u) fp.claim_a(acc_free_u, s1);
              if (s0 != s1 && ((deferred >> s0) & 1u)) fp.claim_a(acc_free_u, s0);
            }
          }
          if (late) ptx::mbar_wait_a(a_full_u + 8u * uint32_t(st.i), st.w & 1);
          if (mw == 0 && lane == 0) STRACE(3, st.w * P.SA + st.i);
          if (!tok_inside) ptx::tc_fence_after_sync();
          const uint32_t a0 = a_base + uint32_t(st.i) * (kStage2 >> 4);
          const uint32_t tok_par = tok & 1u;
          if (c == 0 && need_token) ++tok;
          if (ptx::elect_one()) {
            // One dynamic loop over the K=16 steps of the chunk, the (row, tap) MMAs of a step unrolled with immediate
            // descriptor offsets (a fully predicated 24-MMA unroll cost ~250 instructions per stage).
            const uint32_t dc0 = tmem_base + uint32_t(dr.i * P.SW);  // dr.i is even and R is even: no wrap inside a pair
            uint32_t ak = a0;
            if (tok_inside) {
              ptx::mbar_wait_a(my_turn_u, tok_par);
              ptx::tc_fence_after_sync();
            }
            if (!(P.ablate & 4)) {
              for (int k = 0; k < ksteps; ++k, ak += a_k) {
                // weights are packed per 64-channel chunk; a group-planar stage may hold five groups (chunk 1, step 0)
                const int gi = P.gps * c + k;
                const uint32_t bk = GP == 1 ? b_base + uint32_t(gi >> 2) * (RFOLD ? 3u * blk16 : blk16) + uint32_t(gi & 3) * 2u
                                            : b0 + 2u * uint32_t(k);  // hybrid: four groups per planar stage = one chunk
                if (RFOLD) {
#pragma unroll
                  for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int s = 0; s < 3; ++s)
                      ptx::umma_bf16(dc0 + uint32_t(r * 16), adesc_hi | (ak + uint32_t(r) * a_row + uint32_t(s) * a_tap),
                                     desc_hi | (bk + uint32_t(s * 48 * 8)), idesc, 1u);
                } else {
                  const uint32_t acc = (c | k) != 0 ? 1u : 0u;
                  ptx::umma_bf16(dc0, adesc_hi | ak, desc_hi | bk, idesc, acc);
                  ptx::umma_bf16(dc0 + uint32_t(P.SW), adesc_hi | (ak + a_row), desc_hi | bk, idesc, acc);
                }
              }
            }
            ptx::umma_commit_a(a_empty_u + 8u * uint32_t(st.i));
            if (c == P.npc - 1) {
              ptx::umma_commit_a(acc_done_u + 8u * uint32_t(dr.i >> 1));
              if (last) ptx::umma_commit_a(next_turn_u);
            }
            if (mw == 0) STRACE(4, st.w * P.SA + st.i);
          }
          __syncwarp();
          st.step(P.SA);
          b0 += RFOLD ? 3u * blk16 : blk16;
        }
        dr.add(2, P.R);
        next_pair(cur_mask);
      }
      if (PAD) {  // the trailing accumulator-row pair of the segment receives no further input
        if (owner == mw) {
          if (tpos == 0) {  // gp > 0 here: every segment has at least one input pair
            ptx::mbar_wait_a(my_turn_u, tok & 1u);
            ++tok;
          }
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {
            ptx::umma_commit_a(acc_done_u + 8u * uint32_t(dr.i >> 1));
            if (tpos == PT - 1) ptx::umma_commit_a(next_turn_u);
          }
          __syncwarp();
        }
        dr.add(2, P.R);
        next_pair(0u);
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ============================================================ epilogue: group eg owns row (pair start + eg)
    const int eg = (warp - kEpiWarp0) >> 2, q = warp & 3;
    const uint32_t lb = tmem_base + (uint32_t(q * 32) << 16);
    const int px = q * 32 + lane;
    SlotPhases dp;  // acc_done phases, one bit per slot pair
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int n_acc_pairs = ((it.h1 - it.h0 + 2 * PAD + 1) >> 1) + PAD;
      const int col = it.w0 + px;
      const bool col_ok = px < P.TW && col < P.W;
      int i = it.h0 - 2 * PAD + eg;
      const size_t row_elems = EPI == kSNchwOut ? size_t(P.W) : size_t(P.W) * P.out_ld;
      bf16* o_b = nullptr;
      float* o_f = nullptr;
      if (EPI == kSNchwOut) o_f = P.out_nchw + (size_t(it.n) * P.Cout * P.H + i) * P.W + col;
      else o_b = P.out + ((size_t(it.n) * P.H + i) * P.W + col) * P.out_ld;
      Ring pr;
      pr.i = it.h0 % P.R;
      for (int pp = 0; pp < n_acc_pairs; ++pp) {
        const int pslot = pr.i >> 1, slot = pr.i + eg;
        dp.wait(acc_done, pslot);
        if (eg == 0 && q == 0 && lane == 0 && item == int(blockIdx.x)) STRACE(5, pp);
        ptx::tc_fence_after_sync();
        const bool row_ok = i >= it.h0 && i < it.h1;
        const bool shadow = PAD && slot < 2;
        const uint32_t tm = lb + uint32_t(slot * P.SW), ts = lb + uint32_t((P.R + slot) * P.SW);
        for (int c0 = 0; c0 < ((P.ablate & 8) ? 0 : P.NT); c0 += 8) {
          uint32_t v[8];
          if (row_ok) {
            ptx::tmem_ld8(tm + c0, v);
            if (shadow) {
              uint32_t v2[8];
              ptx::tmem_ld8(ts + c0, v2);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
            } else {
              ptx::tmem_wait_ld();
            }
          }
          if (PAD) {
            ptx::tmem_st8_zero(tm + c0);
            if (shadow) ptx::tmem_st8_zero(ts + c0);
          }
          if (row_ok && col_ok) {
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
            const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y, __uint_as_float(v[2]) + b0.z,
                                __uint_as_float(v[3]) + b0.w, __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
            if (EPI == kSNchwOut) {
              const size_t plane = size_t(P.H) * P.W;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (c0 + e < P.Cout) o_f[size_t(c0 + e) * plane] = P.sigmoid ? 1.0f / (1.0f + __expf(-f[e])) : f[e];
            } else if (c0 < P.Cout && !(P.ablate & 1)) {  // channel slices are padded to multiples of 8
              *reinterpret_cast<uint4*>(o_b + c0) = make_uint4(bf2(f[0], f[1]), bf2(f[2], f[3]), bf2(f[4], f[5]), bf2(f[6], f[7]));
            }
          }
        }
        if (PAD) ptx::tmem_wait_st();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_free[pslot]);
        if (eg == 0 && q == 0 && lane == 0 && item == int(blockIdx.x)) STRACE(6, pp);
        pr.add(2, P.R);
        i += 2;
        if (EPI == kSNchwOut) o_f += 2 * row_elems; else o_b += 2 * row_elems;
      }
    }
  } else if (warp >= kWorkWarp0) {
    // ============================================================ pre-activation workers (two groups alternate stages)
    const int aw = warp - kWorkWarp0;
    const int grp = aw >> 3, t = (aw & 7) * 32 + lane, u = t & 7, qb = t >> 3;
    const uint32_t sA_u = ptx::smem_u32(sA);
    // planar stages: thread <-> (pixel p, 16-byte half h) of every group row; SWIZZLE_32B swaps the halves of pixels with
    // bit 2 set.  NHWC stages: thread <-> (16-byte channel unit u, pixels qb + 32*i) of both rows.
    const int p = t >> 1, h = t & 1;
    const uint32_t offp = uint32_t(p) * 32u + (uint32_t((h ^ (p >> 2)) & 1) << 4);
    const uint32_t sc_u = ptx::smem_u32(s_sc) + uint32_t(h) * 16u, sh_u = ptx::smem_u32(s_sh) + uint32_t(h) * 16u;
    uint32_t off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) off[i] = ptx::sw128_offset(uint32_t(qb + 32 * i), uint32_t(u));
    __nv_bfloat162 sc[4], sh[4];
    int cached_c = -1;
    Ring st;
    int turn = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int colp = it.w0 - PAD + p;
      const bool okp = colp >= 0 && colp < P.W;
      bool ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = it.w0 - PAD + qb + 32 * i;
        ok[i] = col >= 0 && col < P.W;
      }
      const int npairs = (it.h1 - it.h0 + 2 * PAD + 1) >> 1;
      for (int pp = 0; pp < npairs; ++pp) {
        const int j = it.h0 - PAD + 2 * pp;
        for (int c = 0; c < P.npc; ++c, turn ^= 1, st.step(P.SA)) {
          // wsplit: both groups work on every stage, one image row each (balanced whatever the mix of stage kinds, and
          // every group observes every phase of every slot); else the groups take alternate stages
          if (!P.wsplit && turn != grp) continue;
          const int r_lo = P.wsplit ? grp : 0, r_hi = P.wsplit ? grp + 1 : 2;
          if (planar_stage(c)) {
            const int g0 = P.gps * (c - P.nh);
            const int ng = min(P.gps, P.ngroups - g0);
            ptx::mbar_wait(&raw_full[st.i], st.w & 1);
            if ((aw & 7) == 0 && lane == 0) STRACE(1, st.w * P.SA + st.i);
            if (okp && !(P.ablate & 2)) {
              const uint32_t base = sA_u + uint32_t(st.i) * kStage2 + offp;
              const uint32_t tab = uint32_t(P.nh * 64 + g0 * 16) * 2u;  // table offset of the stage's first channel
#pragma unroll
              for (int g = 0; g < 5; ++g) {
                if (g >= ng) break;
                uint4 x[2];
#pragma unroll
                for (int r = 0; r < 2; ++r)
                  if (r >= r_lo && r < r_hi) x[r] = ptx::lds128(base + uint32_t(g) * kGroupBytes + uint32_t(r) * kGroupRow);
                const uint4 csc = ptx::lds128(sc_u + tab + uint32_t(g) * 32u), csh = ptx::lds128(sh_u + tab + uint32_t(g) * 32u);
                const __nv_bfloat162* gsc = reinterpret_cast<const __nv_bfloat162*>(&csc);
                const __nv_bfloat162* gsh = reinterpret_cast<const __nv_bfloat162*>(&csh);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                  if (r < r_lo || r >= r_hi) continue;
                  __nv_bfloat162* v = reinterpret_cast<__nv_bfloat162*>(&x[r]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] = __hfma2_relu(v[e], gsc[e], gsh[e]);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r)  // rows outside the image keep TMA's zero fill (padding after activation)
                  if (r >= r_lo && r < r_hi && j + r >= 0 && j + r < P.H) ptx::sts128(base + uint32_t(g) * kGroupBytes + uint32_t(r) * kGroupRow, x[r]);
              }
            }
          } else {
            if (c != cached_c) {
              const float4 fs0 = *reinterpret_cast<const float4*>(s_pre_s + c * 64 + u * 8);
              const float4 fs1 = *reinterpret_cast<const float4*>(s_pre_s + c * 64 + u * 8 + 4);
              const float4 ft0 = *reinterpret_cast<const float4*>(s_pre_t + c * 64 + u * 8);
              const float4 ft1 = *reinterpret_cast<const float4*>(s_pre_t + c * 64 + u * 8 + 4);
              sc[0] = __floats2bfloat162_rn(fs0.x, fs0.y); sc[1] = __floats2bfloat162_rn(fs0.z, fs0.w);
              sc[2] = __floats2bfloat162_rn(fs1.x, fs1.y); sc[3] = __floats2bfloat162_rn(fs1.z, fs1.w);
              sh[0] = __floats2bfloat162_rn(ft0.x, ft0.y); sh[1] = __floats2bfloat162_rn(ft0.z, ft0.w);
              sh[2] = __floats2bfloat162_rn(ft1.x, ft1.y); sh[3] = __floats2bfloat162_rn(ft1.z, ft1.w);
              cached_c = c;
            }
            const bool active = u * 8 < min(64, P.Cin - c * 64);
            ptx::mbar_wait(&raw_full[st.i], st.w & 1);
            if (active) {
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                if (r < r_lo || r >= r_hi) continue;
                if (j + r < 0 || j + r >= P.H) continue;  // rows outside the image keep TMA's zero fill (padding after activation)
                const uint32_t base = sA_u + uint32_t(st.i) * kStage2 + uint32_t(r) * kStage;
                uint4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = ptx::lds128(base + off[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162* v = reinterpret_cast<__nv_bfloat162*>(&x[i]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] = __hfma2_relu(v[e], sc[e], sh[e]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (ok[i]) ptx::sts128(base + off[i], x[i]);
              }
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[st.i]);
          if ((aw & 7) == 0 && lane == 0) STRACE(2, st.w * P.SA + st.i);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled stream_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

constexpr int kSmemLimit = 232448 - 2048;  // 227 KB opt-in maximum minus static shared (barriers) and alignment slack

void put_bf16(uint8_t* blk, int row, int k, float val) {
  const bf16 b = __float2bfloat16_rn(val);
  std::memcpy(blk + ptx::sw128_offset(uint32_t(row), uint32_t(k >> 3)) + (k & 7) * 2, &b, 2);
}

int stream_nt(int Cout, int ks) {
  if (Cout <= 16) return 16;
  if (ks == 3) return 0;  // folded 3x3 only for 16-wide outputs (ring = 32 slots); wider 3x3 use the tile kernel
  if (Cout <= 64) return 64;
  return 128;
}

}  // namespace

int stream_pack_create(const float* w, const float* bias, int Cin, int Cout, int CoutP, int ks, StreamPack** out) {
  *out = nullptr;
  if (ks != 1 && ks != 3) return fail("stream_pack: ks must be 1 or 3");
  StreamPack* p = new StreamPack();
  p->Cin = Cin; p->Cout = Cout; p->ks = ks;
  p->nchunks = (Cin + 63) / 64;
  p->NT = stream_nt(Cout, ks);
  std::vector<uint8_t> img;
  std::vector<float> hb;
  if (p->NT) {
    // 3x3: one [144 x 64ch] block per K-chunk, rows n = pos*48 + s*16 + co with window position pos <-> kernel row
    // r = 2 - pos (input row j feeds output rows j-1+pos) and s the horizontal tap; 1x1: rows n = co.
    const int NT = p->NT, NMMA = ks == 3 ? 9 * NT : NT;
    p->npass = (Cout + NT - 1) / NT;
    const size_t block = size_t(NMMA) * 128;
    p->pass_bytes = size_t(p->nchunks) * block;
    img.assign(p->pass_bytes * p->npass, 0);
    hb.assign(size_t(p->npass) * NT, 0.f);
    for (int pass = 0; pass < p->npass; ++pass)
      for (int c = 0; c < p->nchunks; ++c) {
        uint8_t* blk = img.data() + pass * p->pass_bytes + size_t(c) * block;
        for (int pos = 0; pos < (ks == 3 ? 3 : 1); ++pos)
          for (int s = 0; s < (ks == 3 ? 3 : 1); ++s)
            for (int nn = 0; nn < NT; ++nn) {
              const int co = pass * NT + nn;
              if (co >= Cout) continue;
              const int tap = ks == 3 ? (2 - pos) * 3 + s : 0;
              for (int k = 0; k < 64; ++k) {
                const int ci = c * 64 + k;
                if (ci >= Cin) continue;
                put_bf16(blk, (pos * 3 + s) * NT + nn, k, w[(size_t(tap) * Cin + ci) * CoutP + co]);
              }
            }
      }
    if (cudaMalloc(&p->d_w, img.size()) != cudaSuccess ||
        cudaMemcpy(p->d_w, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      stream_pack_destroy(p);
      return fail("stream_pack: weight upload failed");
    }
  }
  if (ks == 3 && Cin == 3 && Cout <= 64) {  // conv1: horizontal taps folded into K, vertical taps into N (NT = 64)
    std::vector<uint8_t> k(size_t(192) * 128, 0);
    for (int pos = 0; pos < 3; ++pos)
      for (int co = 0; co < Cout; ++co)
        for (int s = 0; s < 3; ++s)
          for (int ci = 0; ci < 3; ++ci)
            put_bf16(k.data(), pos * 64 + co, s * 3 + ci, w[(size_t((2 - pos) * 3 + s) * Cin + ci) * CoutP + co]);
    if (cudaMalloc(&p->d_wk, k.size()) != cudaSuccess ||
        cudaMemcpy(p->d_wk, k.data(), k.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      stream_pack_destroy(p);
      return fail("stream_pack: weight upload failed");
    }
  }
  if (ks == 3 && Cout <= 16) {  // row-fold form of the 16-wide 3x3 layers
    const size_t block = size_t(48) * 128;
    std::vector<uint8_t> wr(size_t(p->nchunks) * 3 * block, 0);
    for (int c = 0; c < p->nchunks; ++c)
      for (int sx = 0; sx < 3; ++sx)
        for (int pos = 0; pos < 3; ++pos)
          for (int co = 0; co < Cout; ++co)
            for (int k = 0; k < 64; ++k) {
              const int ci = c * 64 + k;
              if (ci >= Cin) continue;
              put_bf16(wr.data() + (size_t(c) * 3 + sx) * block, pos * 16 + co, k, w[(size_t((2 - pos) * 3 + sx) * Cin + ci) * CoutP + co]);
            }
    p->rfold_bytes = wr.size();
    if (cudaMalloc(&p->d_wr, wr.size()) != cudaSuccess ||
        cudaMemcpy(p->d_wr, wr.data(), wr.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      stream_pack_destroy(p);
      return fail("stream_pack: weight upload failed");
    }
  }
  if (ks == 3) {  // wide form: per 64-channel pass nine [NTw x 64ch] blocks per K-chunk, resident in shared memory
    static const int ntw_big = getenv("CDAN_NTW") ? atoi(getenv("CDAN_NTW")) : 128;  // conv2: one N=128 pass (2.49 ms) beats two N=64 passes (2.82 ms)
    const int NTw = Cout <= 16 ? 16 : (Cout > 64 ? ntw_big : 64), npw = (Cout + NTw - 1) / NTw;
    const size_t block = size_t(NTw) * 128, bytes = size_t(p->nchunks) * 9 * block;
    if (bytes <= 152 * 1024 && npw <= 2) {  // conv3 in four passes measured no faster than the tile kernel
      std::vector<uint8_t> ww(bytes * npw, 0);
      for (int pass = 0; pass < npw; ++pass)
        for (int c = 0; c < p->nchunks; ++c)
          for (int tap = 0; tap < 9; ++tap)
            for (int nn = 0; nn < NTw; ++nn) {
              const int co = pass * NTw + nn;
              if (co >= Cout) continue;
              for (int k = 0; k < 64; ++k) {
                const int ci = c * 64 + k;
                if (ci >= Cin) continue;
                put_bf16(ww.data() + pass * bytes + (size_t(c) * 9 + tap) * block, nn, k, w[(size_t(tap) * Cin + ci) * CoutP + co]);
              }
            }
      p->NTw = NTw;
      p->npass_w = npw;
      p->wide_bytes = bytes;
      if (cudaMalloc(&p->d_ww, ww.size()) != cudaSuccess ||
          cudaMemcpy(p->d_ww, ww.data(), ww.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        stream_pack_destroy(p);
        return fail("stream_pack: weight upload failed");
      }
    }
  }
  {
    hb.assign(std::max<size_t>(std::max<size_t>(128, size_t(Cout)), size_t(p->npass) * std::max(p->NT, 1)), 0.f);
    for (int co = 0; co < Cout; ++co) hb[co] = bias[co];
    if (cudaMalloc(&p->d_bias, hb.size() * 4) != cudaSuccess ||
        cudaMemcpy(p->d_bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      stream_pack_destroy(p);
      return fail("stream_pack: bias upload failed");
    }
  }
  *out = p;
  return 0;
}

void stream_pack_destroy(StreamPack* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  if (p->d_wk) cudaFree(p->d_wk);
  if (p->d_ww) cudaFree(p->d_ww);
  if (p->d_wr) cudaFree(p->d_wr);
  if (p->d_bias) cudaFree(p->d_bias);
  delete p;
}

bool conv_stream_supported(const ConvDesc& d, const StreamPack& pk) {
  if (d.in_nchw) return pk.d_wk && d.Cin == 3 && d.W % 4 == 0 && d.relu && !d.pre_scale && !d.out_nchw && d.out_ld % 8 == 0 && (!d.pool || !((d.H | d.W) & 1));
  if (d.Cin % 8 != 0 || d.in_ld % 8 != 0) return false;
  if (d.in_gstride) {  // group-planar input: dense pre-activation layers of the two-row kernel only
    if (!d.pre_scale || d.relu || d.pool || d.Cin % 16 != 0) return false;
    if (d.Chead ? (d.Chead % 64 != 0 || d.Chead >= d.Cin || !d.in2 || d.out_nchw) : d.in_ld != 16) return false;
    if (d.in_gstride != size_t(d.N) * d.H * d.W * 16) return false;  // planes must be contiguous (one tensor map)
    if (d.ks == 3) return pk.d_wr && !d.out_nchw && d.out_ld % 8 == 0;
    return pk.d_w && pk.NT > 0 && (d.out_nchw ? d.Cout <= 16 : d.out_ld % 8 == 0);
  }
  if ((d.relu != 0) != (d.pre_scale == nullptr)) return false;  // ReLU is compiled in per input mode
  if (d.ks == 3 && !d.pre_scale) {  // TMA-fed 3x3: nine-tap fold for 16-wide outputs, else wide form (weights resident)
    if (d.out_nchw || d.out_ld % 8 != 0) return false;
    if (pk.d_w && pk.NT == 16 && !d.pool) return true;
    return pk.d_ww && (!d.pool || !((d.H | d.W) & 1));
  }
  if (!pk.d_w || pk.NT == 0 || d.pool) return false;
  if (d.out_nchw) return d.Cout <= 16;
  return d.out_ld % 8 == 0;
}

// Number of kernels conv_stream_launch issues for this convolution (output-channel passes).
int conv_stream_kernel_count(const ConvDesc& d, const StreamPack& pk) {
  if (d.in_nchw) return 1;
  if (d.ks == 3 && !d.pre_scale) return (pk.d_w && pk.NT == 16 && !d.pool) ? 1 : pk.npass_w;
  return d.ks == 3 ? 1 : pk.npass;
}

int conv_stream_launch(const ConvDesc& d, const StreamPack& pk, cudaStream_t stream) {
  if (!conv_stream_supported(d, pk)) return fail("conv_stream: unsupported convolution shape");
  const bool kfold = d.in_nchw != nullptr;
  const int in_mode = kfold ? kSNchw : (d.pre_scale ? kSPro : kSTma);
  const int fold = d.ks == 3 ? 3 : 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  SParams P{};
  P.N = d.N; P.H = d.H; P.W = d.W; P.Cin = d.Cin;
  P.pad = fold == 3 ? 1 : 0;
  const bool fold9 = fold == 3 && in_mode == kSTma && pk.d_w && pk.NT == 16 && !d.pool;  // TMA-fed, 16-wide: nine-tap fold
  const bool wide = fold == 3 && in_mode == kSTma && !fold9;
  P.wide = wide ? 1 : 0;
  P.NT = kfold ? 64 : (wide ? pk.NTw : pk.NT);
  // Optional (CDAN_DUAL=1): dense pre-activation layers as two half-size CTAs per SM (kSPro2) when their weights leave
  // room for >= 3 stages in 110 KB; 3x3 layers then use the row-fold form (16 TMEM columns per ring slot fit a 256-column
  // allocation).  Measured 3 % slower than one full-size CTA per SM on B200, so it is off by default.
  static const bool dual_enabled = getenv("CDAN_DUAL") && atoi(getenv("CDAN_DUAL")) != 0;
  constexpr int kSmemLimit2 = 110 * 1024;  // 2 x (dynamic + 1 KB align slack + static barriers + 1 KB reserved) <= 228 KB
  bool dual = false;
  if (dual_enabled && in_mode == kSPro && !d.pool) {
    const size_t wb = fold == 3 ? (pk.d_wr && !d.out_nchw ? pk.rfold_bytes : 0) : (pk.NT <= 64 ? pk.pass_bytes : 0);
    const int tail2 = 2 * pk.nchunks * 64 * 4 + 64 * 4 + 256;
    dual = wb > 0 && (kSmemLimit2 - 1024 - int(wb) - tail2) / kStage >= 3;
  }
  // Two rows per stage (conv_stream2_kernel) is the default for all dense pre-activation layers (3x3 in the row-fold
  // form): the MMA warp's per-iteration latency is on the critical path and is paid once per two rows.  CDAN_RPS=1
  // selects the one-row kernel (nine-tap fold) instead.
  static const int rps_env = getenv("CDAN_RPS") ? atoi(getenv("CDAN_RPS")) : 2;
  bool rps2 = false;
  const int gp = d.in_gstride != 0 ? (d.Chead ? 2 : 1) : 0;  // 1 = group-planar input, 2 = NHWC head + group planes
  if ((rps_env != 1 || gp) && !dual && in_mode == kSPro && !d.pool) {
    const size_t wb = fold == 3 ? (pk.d_wr && !d.out_nchw ? pk.rfold_bytes : 0) : pk.pass_bytes;
    const int tail2 = 2 * pk.nchunks * 64 * 4 + 128 * 4 + 256 + 128 * 4 + pk.nchunks * 64 * 4;
    rps2 = wb > 0 && (kSmemLimit - 1024 - int(wb) - tail2) / (2 * kStage) >= 4;
  }
  static const char* dense_form = getenv("CDAN_DENSE_FORM");  // "rfold" | "shift" (A/B switch), default per layer
  const bool rfold = fold == 3 && in_mode == kSPro && pk.d_wr && !d.out_nchw && (dual || rps2 || (dense_form && !strcmp(dense_form, "rfold")));  // measured equal or slightly slower than the nine-tap fold on B200
  const bool shift = (fold == 3 && in_mode == kSPro && !rfold) || fold9;
  P.NMMA = shift ? 9 * P.NT : (wide ? P.NT : fold * P.NT);  // rfold: 3 * 16
  P.SW = shift ? 3 * P.NT : P.NT;
  const int tmem_cols = dual ? 256 : 512;
  P.R = (fold == 3 && !wide) ? std::min(30, tmem_cols / P.SW - 2) : std::min(kMaxR, tmem_cols / P.NT);
  P.R &= ~1;
  P.nchunks = kfold ? 1 : pk.nchunks;
  P.nS = 1;
  const int tw_max = shift ? 120 : ((wide || rfold) ? 126 : 128);
  P.strips = ceil_div(d.W, tw_max);
  P.TW = std::min(tw_max, kfold ? (ceil_div(d.W, P.strips) + 3) & ~3 : (ceil_div(d.W, P.strips) + 1) & ~1);
  const int want_segs = std::max(1, ceil_div(sms * 8, d.N * P.strips));
  P.SEG = std::max(std::min(32, d.H), (ceil_div(d.H, want_segs) + 1) & ~1);
  // Small launches (a few images, or the 1/8-resolution layers of one image): with the 32-row floor most SMs get no item at
  // all (one 1080p image at 1/8 resolution: 2 strips x 5 segments = 10 items for 148 SMs, 0.22 ms for a transition that takes
  // 0.02 ms per image in a batch of 32).  Then pick the even height with the lowest cost = rounds of items per SM x (rows per
  // item + re-read halo rows + pipeline fill); results do not depend on the segmentation (absolute-row ring slots).
  if (long(d.N) * P.strips * ceil_div(d.H, P.SEG) < 2L * sms && d.H > 8) {
    const int over = (fold == 3 ? 2 : 0) + 6;
    long best = -1;
    for (int seg = 8; seg <= ((d.H + 1) & ~1); seg += 2) {
      const long items = long(d.N) * P.strips * ceil_div(d.H, seg);
      const long cost = ((items + sms - 1) / sms) * (std::min(seg, d.H) + over);
      if (best < 0 || cost <= best) {
        best = cost;
        P.SEG = seg;
      }
    }
  }
  P.segs = ceil_div(d.H, P.SEG);
  P.nitems = d.N * P.strips * P.segs;
  P.relu = d.relu; P.sigmoid = d.sigmoid; P.Cout = d.Cout;
  P.in = reinterpret_cast<const bf16*>(d.in); P.in_ld = d.in_ld;
  P.in_nchw = d.in_nchw; P.pre_s = d.pre_scale; P.pre_t = d.pre_shift;
  P.out_ld = d.out_ld; P.out_nchw = d.out_nchw;
  P.wbytes = uint32_t(kfold ? size_t(192) * 128 : (wide ? pk.wide_bytes : (rfold ? pk.rfold_bytes : pk.pass_bytes)));
  const int tail = 2 * P.nchunks * 64 * 4 + std::max(P.NT, 64) * 4 + 256 + (kfold ? kRawStages * kRawFloats * 4 + 128 : 0) +
                   (rps2 ? 128 * 4 + P.nchunks * 64 * 4 : 0);  // two-row kernel: bf16 scale / shift tables behind a 128-float bias vector
  if (gp && !rps2) return fail("conv_stream: group-planar input needs the two-row kernel (weights too large)");
  static const int ni_env = getenv("CDAN_ISSUERS") ? atoi(getenv("CDAN_ISSUERS")) : 3;
  P.ni = ni_env == 2 ? 2 : 3;
  // two pairs per turn measured slower (the second pair's waits sit inside the token chain): 41.1 vs 40.1 ms per step
  static const int pt_env = getenv("CDAN_PAIRS_PER_TURN") ? atoi(getenv("CDAN_PAIRS_PER_TURN")) : 0;
  P.pt = pt_env == 2 ? 2 : 1;  // 3: adaptive, chosen below once the stage count is known
  static const int tok_env = getenv("CDAN_TOKEN_INSIDE") ? atoi(getenv("CDAN_TOKEN_INSIDE")) : 0;  // measured equal (39.0 vs 39.1 ms): off
  P.tok_inside = tok_env != 0;
  static const int ws_env = getenv("CDAN_WORKER_SPLIT") ? atoi(getenv("CDAN_WORKER_SPLIT")) : 0;  // measured slower (39.7 vs 38.8 ms per step): off
  P.wsplit = ws_env != 0;
  P.nh = gp == 2 ? d.Chead / 64 : 0;
  P.ngroups = gp ? (d.Cin - d.Chead) / 16 : 0;
  P.gps = gp == 1 ? (d.Cin == 80 ? 5 : std::min(4, d.Cin / 16)) : 4;  // the 80-channel transition fits one 40 KB stage per row pair
  P.npc = gp ? P.nh + ceil_div(P.ngroups, P.gps) : P.nchunks;
  P.stage_bytes = gp == 1 ? P.gps * 8192 : 0;
  const int stage_bytes = gp == 1 ? P.stage_bytes : (rps2 ? 2 * kStage : kStage);
  P.SA = std::min(gp == 1 ? kMaxSA2 : kMaxSA, ((dual ? kSmemLimit2 : kSmemLimit) - 1024 - int(P.wbytes) - tail) / stage_bytes);
  // The two worker groups take alternate stages: with an even stage count every ring slot always belongs to the same
  // group.  (With an odd count a group could test a slot's mbarrier parity a full phase ahead of the last phase it
  // observed there — parity waits then return a false positive and the pipeline desynchronises.)
  if (in_mode == kSPro) P.SA &= ~1;
  static const int sa_cap = getenv("CDAN_SA_MAX") ? atoi(getenv("CDAN_SA_MAX")) : kMaxSA;
  P.SA = std::min(P.SA, std::max(3, sa_cap));
  if (P.SA < 3) return fail("conv_stream: weights leave no room for the activation pipeline");
  // CDAN_PAIRS_PER_TURN=3: two row pairs per issuer turn where the stage ring is deep enough for both pairs' stages to
  // be awaited ahead of the token (SA >= issuers * 2 * stages per pair).  Measured slower than one pair per turn even so
  // (final dense layers 1.45/1.56/1.79/2.16 -> 1.54/1.82/2.10/2.40 ms), hence not the default.
  if (pt_env == 3) P.pt = (rps2 && P.SA >= P.ni * 2 * P.npc) ? 2 : 1;
  const int smem_bytes = P.SA * stage_bytes + 1024 + int(P.wbytes) + tail + 1024;

  CUtensorMap tmap, tmapG;
  std::memset(&tmap, 0, sizeof(tmap));
  std::memset(&tmapG, 0, sizeof(tmapG));
  if (in_mode != kSNchw) {
    PFN_encodeTiled enc = stream_get_encode();
    if (!enc) return fail("conv_stream: cuTensorMapEncodeTiled is not available from the driver");
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (gp != 1) {  // NHWC input (gp 2: the head of a hybrid concat buffer)
      if (reinterpret_cast<uintptr_t>(d.in) % 16 != 0) return fail("conv_stream: input pointer must be 16-byte aligned");
      const int Cn = gp == 2 ? d.Chead : d.Cin;
      cuuint64_t gdim[4] = {cuuint64_t(Cn), cuuint64_t(d.W), cuuint64_t(d.H), cuuint64_t(d.N)};
      cuuint64_t gstr[3] = {cuuint64_t(d.in_ld) * 2, cuuint64_t(d.W) * d.in_ld * 2, cuuint64_t(d.H) * d.W * d.in_ld * 2};
      cuuint32_t box[4] = {64, cuuint32_t(shift ? 32 : 128), cuuint32_t(rps2 ? 2 : 1), 1};
      CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.in), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       // 128-byte L2 promotion only when every K-chunk is a full 128-byte line; otherwise 64 bytes (ncu:
                       // with promotion NONE a 32-byte line still pulled 128 bytes from DRAM, with 64B it pulls 64)
                       Cn % 64 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("conv_stream: cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
    }
    if (gp) {
      // group planes: the planes of all groups form one [N * groups][H][W][16] tensor, image n of group g = index n + g*N
      const void* base = gp == 2 ? d.in2 : d.in;
      if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail("conv_stream: plane pointer must be 16-byte aligned");
      cuuint64_t gdim[4] = {16, cuuint64_t(d.W), cuuint64_t(d.H), cuuint64_t(d.N) * cuuint64_t(P.ngroups)};
      cuuint64_t gstr[3] = {32, cuuint64_t(d.W) * 32, cuuint64_t(d.H) * d.W * 32};
      cuuint32_t box[4] = {16, 128, 2, 1};
      CUresult r = enc(&tmapG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("conv_stream: cuTensorMapEncodeTiled (planes) failed with code " + std::to_string(int(r)));
      if (gp == 1) tmap = tmapG;
    }
  }
  if (in_mode == kSNchw) {
    PFN_encodeTiled enc = stream_get_encode();
    if (!enc) return fail("conv_stream: cuTensorMapEncodeTiled is not available from the driver");
    if (reinterpret_cast<uintptr_t>(d.in_nchw) % 16 != 0 || d.W % 4 != 0)
      return fail("conv_stream: planar fp32 input must be 16-byte aligned with W a multiple of 4");
    cuuint64_t gdim[4] = {cuuint64_t(d.W), cuuint64_t(d.H), 3, cuuint64_t(d.N)};
    cuuint64_t gstr[3] = {cuuint64_t(d.W) * 4, cuuint64_t(d.H) * d.W * 4, cuuint64_t(3) * d.H * d.W * 4};
    cuuint32_t box[4] = {cuuint32_t(kRawW), 1, 3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d.in_nchw), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("conv_stream: cuTensorMapEncodeTiled (fp32 input) failed with code " + std::to_string(int(r)));
  }
#ifdef CDAN_STREAM_TRACE_BUILD
  static const int trace_launch = getenv("CDAN_STREAM_TRACE") ? atoi(getenv("CDAN_STREAM_TRACE")) : -1;
  static int launch_ctr = 0;
  static unsigned long long* d_trace = nullptr;
  const bool tracing = trace_launch >= 0 && launch_ctr++ == trace_launch;
  if (tracing) {
    if (!d_trace) cudaMalloc(&d_trace, 8 * kTraceN * sizeof(unsigned long long));
    cudaMemsetAsync(d_trace, 0, 8 * kTraceN * sizeof(unsigned long long), stream);
    P.trace = d_trace;
  }
#endif
  static const int ablate = getenv("CDAN_ABLATE") ? atoi(getenv("CDAN_ABLATE")) : 0;
  P.ablate = ablate;
  const int grid = std::min(P.nitems, dual ? 2 * sms : sms);
  const int threads = rps2 ? 896 : 32 * (kEpiWarp0 + epi_warps(dual ? kSPro2 : in_mode) + work_warps(dual ? kSPro2 : in_mode));
  const int npass = kfold ? 1 : (wide ? pk.npass_w : pk.npass);
  for (int pass = 0; pass < npass; ++pass) {
    P.wpack = kfold ? pk.d_wk : (wide ? pk.d_ww + size_t(pass) * pk.wide_bytes : (rfold ? pk.d_wr : pk.d_w + size_t(pass) * pk.pass_bytes));
    P.bias = pk.d_bias + size_t(pass) * P.NT;
    P.Cout = std::min(P.NT, d.Cout - pass * P.NT);
    P.out = reinterpret_cast<bf16*>(d.out) + size_t(pass) * P.NT;
    auto launch = [&](auto kern) -> int {
      CDAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      kern<<<grid, threads, smem_bytes, stream>>>(tmap, P);
      CDAN_CUDA_OK(cudaGetLastError());
      return 0;
    };
    int rc;
    auto launch2 = [&](auto kern) -> int {
      CDAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      kern<<<grid, threads, smem_bytes, stream>>>(tmap, tmapG, P);
      CDAN_CUDA_OK(cudaGetLastError());
      return 0;
    };
    if (rps2) {
      if (gp == 2) {
        rc = fold == 3 ? launch2(conv_stream2_kernel<4, kSStore, 2>) : launch2(conv_stream2_kernel<1, kSStore, 2>);
      } else if (gp == 1) {
        if (fold == 3) rc = launch2(conv_stream2_kernel<4, kSStore, 1>);
        else rc = d.out_nchw ? launch2(conv_stream2_kernel<1, kSNchwOut, 1>) : launch2(conv_stream2_kernel<1, kSStore, 1>);
      } else if (fold == 3) rc = launch2(conv_stream2_kernel<4, kSStore, 0>);
      else rc = d.out_nchw ? launch2(conv_stream2_kernel<1, kSNchwOut, 0>) : launch2(conv_stream2_kernel<1, kSStore, 0>);
    } else if (kfold) rc = d.pool ? launch(conv_stream_kernel<kSNchw, 3, kSPool>) : launch(conv_stream_kernel<kSNchw, 3, kSStore>);
    else if (fold == 3) {
      if (d.out_nchw) rc = in_mode == kSPro ? launch(conv_stream_kernel<kSPro, 3, kSNchwOut>) : launch(conv_stream_kernel<kSTma, 3, kSNchwOut>);
      else if (in_mode == kSPro && dual) rc = launch(conv_stream_kernel<kSPro2, 4, kSStore>);
      else if (in_mode == kSPro) rc = rfold ? launch(conv_stream_kernel<kSPro, 4, kSStore>) : launch(conv_stream_kernel<kSPro, 3, kSStore>);
      else if (fold9) rc = launch(conv_stream_kernel<kSTma, 9, kSStore>);
      else rc = d.pool ? launch(conv_stream_kernel<kSTma, 3, kSPool>) : launch(conv_stream_kernel<kSTma, 3, kSStore>);
    } else {
      if (dual) rc = d.out_nchw ? launch(conv_stream_kernel<kSPro2, 1, kSNchwOut>) : launch(conv_stream_kernel<kSPro2, 1, kSStore>);
      else if (d.out_nchw) rc = in_mode == kSPro ? launch(conv_stream_kernel<kSPro, 1, kSNchwOut>) : launch(conv_stream_kernel<kSTma, 1, kSNchwOut>);
      else rc = in_mode == kSPro ? launch(conv_stream_kernel<kSPro, 1, kSStore>) : launch(conv_stream_kernel<kSTma, 1, kSStore>);
    }
    if (rc) return rc;
  }
#ifdef CDAN_STREAM_TRACE_BUILD
  if (tracing) {
    cudaStreamSynchronize(stream);
    std::vector<unsigned long long> t(8 * kTraceN);
    cudaMemcpy(t.data(), d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (auto v : t) if (v && v < t0) t0 = v;
    fprintf(stderr, "STREAM TRACE Cin=%d Cout=%d ks=%d H=%d W=%d N=%d SA=%d R=%d nchunks=%d items=%d SEG=%d strips=%d NMMA=%d\n", d.Cin, d.Cout,
            d.ks, d.H, d.W, d.N, P.SA, P.R, P.nchunks, P.nitems, P.SEG, P.strips, P.NMMA);
    const char* names[8] = {"tma_issue", "act_start", "act_done", "mma_start", "mma_issued", "epi_start", "epi_done", "mma_rowgo"};
    for (int r = 0; r < 8; ++r) {
      fprintf(stderr, "%-10s", names[r]);
      for (int i = 0; i < 96; ++i) fprintf(stderr, " %6lld", t[r * kTraceN + i] ? (long long)(t[r * kTraceN + i] - t0) : -1ll);
      fprintf(stderr, "\n");
    }
  }
#endif
  return 0;
}

}  // namespace cdan
