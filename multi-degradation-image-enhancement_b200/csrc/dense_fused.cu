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
constexpr int kG0Depth = 8;      // shared ring of the F0 versions (released by all five consumers)
constexpr int kValidW = 120;     // output columns per strip
// Version rings of groups 1..4, one per (consumer c' in 1..4 [4 = transition], group g in 1..c'), enumerated c' major.
__host__ __device__ constexpr int ring_depth(int cp, int g) { return cp <= 3 ? cp - g + 3 : 3; }
__host__ __device__ constexpr int ring_base(int cp, int g) {
  int b = kG0Depth;
  for (int c = 1; c <= 4; ++c)
    for (int gg = 1; gg <= c; ++gg) {
      if (c == cp && gg == g) return b;
      b += ring_depth(c, gg);
    }
  return b;
}
__host__ __device__ constexpr int ring_id(int cp, int g) { return cp * (cp - 1) / 2 + (g - 1); }
constexpr int kRowSlots = ring_base(5, 1);  // 42
// Parameter blob (global -> shared by one bulk copy): weights | activation tables | bias
constexpr int kWLayerBlock = 48 * 32;                                    // (layer, group, tap): 48 rows x 32 B
__host__ __device__ constexpr int w_layer_off(int c, int g, int s) { return ((c * (c + 1) / 2 + g) * 3 + s) * kWLayerBlock; }
constexpr int kWTOff = 30 * kWLayerBlock;                                 // transition: 5 blocks of 16 rows x 32 B
constexpr int kTabOff = kWTOff + 5 * 512;                                 // 10 rings x (sc 32 B | sh 32 B)
constexpr int kG0TabOff = kTabOff + 10 * 64;                              // sc 32 B | sh 32 B
constexpr int kBiasOff = kG0TabOff + 64;                                  // 5 x 16 fp32
constexpr int kBlobBytes = kBiasOff + 5 * 16 * 4;
constexpr int kBlobPad = (kBlobBytes + 1023) / 1024 * 1024;
constexpr int kSmemBytes = kBlobPad + kRowSlots * kRowBytes + 256 + 1024;
constexpr int kThreads = 29 * 32;
// TMEM: layer c -> columns [96c, 96c + 96): 4 ring slots + 2 shadow slots of 16 columns; transition: 8 slots from 384.
constexpr uint32_t kTCol = 384;

struct FParams {
  int N, H, W;
  int strips, SEG, segs, nitems;
  const bf16* t4;  // [N][H/2][W/2][t4_ld] relu(bn4(convT4)), channels 0..2
  int t4_ld;
  const float* x;  // [N][3][H][W]
  float* out;      // [N][3][H][W]
  const uint8_t* blob;
};

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

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

struct Smem {
  uint32_t blob, rows, full, empty, acc_done, acc_free, accT_done, accT_free;  // shared-window addresses
  uint64_t* w_full;
  uint32_t tmem;
};

// A G0 row this consumer does not read still has to be released (the ring slot is freed by all five consumers).
__device__ __forceinline__ void skip_g0(const Smem& S, Ring& g0, int lane) {
  ptx::mbar_wait_a(S.full + 8u * uint32_t(g0.i), g0.w & 1);
  if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(S.empty + 8u * uint32_t(g0.i)) : "memory");
  g0.step(kG0Depth);
}

// ------------------------------------------------------------------------------------------------ MMA issuer, layer C
template <int C>
__device__ __forceinline__ void issuer_layer(const FParams& P, const Smem& S, int lane) {
  const uint32_t idesc = ptx::umma_idesc_bf16(128, 48);
  const uint64_t hi = ptx::umma_desc_sw32(0, 256) & 0xffffffff00000000ull;
  const uint32_t acc_done = S.acc_done + 32u * C, acc_free = S.acc_free + 32u * C;
  Ring g0;
  Ring vr[C > 0 ? C : 1];
  SlotPhases fp;
  ptx::mbar_wait(S.w_full, 0);
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo0 = in_lo(sg, 0), hi0 = in_hi(sg, P.H, 0), lo = in_lo(sg, C), hi_ = in_hi(sg, P.H, C);
    for (int r = lo0; r < lo; ++r) skip_g0(S, g0, lane);
    for (int j = lo; j < hi_; ++j) {
      if (j == lo) {
        fp.claim_a(acc_free, (j - 1) & 3);
        fp.claim_a(acc_free, j & 3);
      }
      fp.claim_a(acc_free, (j + 1) & 3);
      ptx::mbar_wait_a(S.full + 8u * uint32_t(g0.i), g0.w & 1);
#pragma unroll
      for (int g = 1; g <= C; ++g) ptx::mbar_wait_a(S.full + 8u * uint32_t(ring_base(C, g) + vr[g - 1].i), vr[g - 1].w & 1);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
        const uint32_t dcol = S.tmem + 96u * C + 16u * uint32_t((j - 1) & 3);
        {
          const uint32_t a = desc_lo(S.rows + uint32_t(g0.i) * kRowBytes), b = desc_lo(S.blob + w_layer_off(C, 0, 0));
#pragma unroll
          for (int s = 0; s < 3; ++s) ptx::umma_bf16(dcol, hi | (a + 2u * s), hi | (b + uint32_t(s) * (kWLayerBlock >> 4)), idesc, 1u);
        }
#pragma unroll
        for (int g = 1; g <= C; ++g) {
          const uint32_t a = desc_lo(S.rows + uint32_t(ring_base(C, g) + vr[g - 1].i) * kRowBytes),
                         b = desc_lo(S.blob + w_layer_off(C, g, 0));
#pragma unroll
          for (int s = 0; s < 3; ++s) ptx::umma_bf16(dcol, hi | (a + 2u * s), hi | (b + uint32_t(s) * (kWLayerBlock >> 4)), idesc, 1u);
        }
        ptx::umma_commit_a(S.empty + 8u * uint32_t(g0.i));
#pragma unroll
        for (int g = 1; g <= C; ++g) ptx::umma_commit_a(S.empty + 8u * uint32_t(ring_base(C, g) + vr[g - 1].i));
        ptx::umma_commit_a(acc_done + 8u * uint32_t((j - 1) & 3));
      }
      __syncwarp();
      g0.step(kG0Depth);
#pragma unroll
      for (int g = 1; g <= C; ++g) vr[g - 1].step(ring_depth(C, g));
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

// ------------------------------------------------------------------------------------------------ MMA issuer, transition
// Program order per step t: G1 row t (opens the accumulator, accumulate = 0), G2 row t-1, G3 row t-2, G4 row t-3, G0 row
// t-3, then the row is handed to the epilogue.  The order is fixed, so the fp32 summation order of a row never changes.
__device__ __forceinline__ void issuer_transition(const FParams& P, const Smem& S, int lane) {
  const uint32_t idesc = ptx::umma_idesc_bf16(128, 16);
  const uint64_t hi = ptx::umma_desc_sw32(0, 256) & 0xffffffff00000000ull;
  Ring g0;
  Ring vr[4];
  SlotPhases fp;
  ptx::mbar_wait(S.w_full, 0);
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo0 = in_lo(sg, 0), hi0 = in_hi(sg, P.H, 0);
    for (int r = lo0; r < sg.h0; ++r) skip_g0(S, g0, lane);
    for (int t = sg.h0; t < sg.h1 + 3; ++t) {
#pragma unroll
      for (int g = 1; g <= 4; ++g) {
        const int r = t - (g - 1);
        if (r < sg.h0 || r >= sg.h1) continue;
        if (g == 1) fp.claim_a(S.accT_free, r & 7);
        const uint32_t slot = uint32_t(ring_base(4, g) + vr[g - 1].i);
        ptx::mbar_wait_a(S.full + 8u * slot, vr[g - 1].w & 1);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          ptx::umma_bf16(S.tmem + kTCol + 16u * uint32_t(r & 7), hi | desc_lo(S.rows + slot * kRowBytes),
                         hi | desc_lo(S.blob + kWTOff + g * 512), idesc, g == 1 ? 0u : 1u);
          ptx::umma_commit_a(S.empty + 8u * slot);
        }
        __syncwarp();
        vr[g - 1].step(ring_depth(4, g));
      }
      const int r = t - 3;
      if (r >= sg.h0) {
        ptx::mbar_wait_a(S.full + 8u * uint32_t(g0.i), g0.w & 1);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          ptx::umma_bf16(S.tmem + kTCol + 16u * uint32_t(r & 7), hi | desc_lo(S.rows + uint32_t(g0.i) * kRowBytes),
                         hi | desc_lo(S.blob + kWTOff), idesc, 1u);
          ptx::umma_commit_a(S.empty + 8u * uint32_t(g0.i));
          ptx::umma_commit_a(S.accT_done + 8u * uint32_t(r & 7));
        }
        __syncwarp();
        g0.step(kG0Depth);
      }
    }
    for (int r = sg.h1; r < hi0; ++r) skip_g0(S, g0, lane);
  }
}

// ------------------------------------------------------------------------------------------------ epilogue, layer C
template <int C>
__device__ __forceinline__ void epilogue_layer(const FParams& P, const Smem& S, uint64_t* full, uint64_t* empty, uint64_t* acc_done,
                                               uint64_t* acc_free, int q, int lane) {
  constexpr int G = C + 1;   // group this layer produces
  constexpr int NC = 4 - C;  // consumers: layers C+1..3, then the transition
  Ring cr[NC];
  SlotPhases dp;
  const int L = q * 32 + lane, idx = L + 1;  // output lane L is centred on ring index L + 1
  const bool idx_ok = idx >= G && idx < 128 - G;
  const uint32_t off[2] = {sw32_off(idx & 127, 0), sw32_off(idx & 127, 1)};
  const uint32_t lb = S.tmem + (uint32_t(q * 32) << 16) + 96u * C;
  ptx::mbar_wait(S.w_full, 0);
  float bias[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) bias[e] = 0.f;
  {
    const float* sb = reinterpret_cast<const float*>(__cvta_shared_to_generic(S.blob + kBiasOff)) + C * 16;
#pragma unroll
    for (int e = 0; e < 16; ++e) bias[e] = sb[e];
  }
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int lo = in_lo(sg, C), hi_ = in_hi(sg, P.H, C);
    const int olo = lo == 0 ? 0 : lo + 1, ohi = hi_ == P.H ? P.H : hi_ - 1;
    const int col = sg.w0 - 4 + idx;
    const bool keep = idx_ok && col >= 0 && col < P.W;
    for (int i = lo - 1; i <= hi_; ++i) {
      const int slot = i & 3;
      dp.wait(acc_done, slot);
      ptx::tc_fence_after_sync();
      const bool valid = i >= olo && i < ohi;
      bool take[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int cp = k < NC - 1 ? C + 1 + k : 4;
        take[k] = valid && i >= in_lo(sg, cp) && i < in_hi(sg, P.H, cp);
        if (take[k]) ptx::mbar_wait(&empty[ring_base(cp, G) + cr[k].i], (cr[k].w & 1) ^ 1);
      }
      const bool shadow = slot < 2;
      const uint32_t tm = lb + 16u * uint32_t(slot), ts = lb + 16u * uint32_t(4 + slot);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[8];
        if (valid) {
          ptx::tmem_ld8(tm + 8u * half, v);
          if (shadow) {
            uint32_t v2[8];
            ptx::tmem_ld8(ts + 8u * half, v2);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          } else {
            ptx::tmem_wait_ld();
          }
        }
        ptx::tmem_st8_zero(tm + 8u * half);
        if (shadow) ptx::tmem_st8_zero(ts + 8u * half);
        if (valid) {
          __nv_bfloat162 raw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            raw[e] = __floats2bfloat162_rn(__uint_as_float(v[2 * e]) + bias[8 * half + 2 * e], __uint_as_float(v[2 * e + 1]) + bias[8 * half + 2 * e + 1]);
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            if (!take[k]) continue;
            const int cp = k < NC - 1 ? C + 1 + k : 4;
            const uint32_t tab = S.blob + kTabOff + uint32_t(ring_id(cp, G)) * 64u + uint32_t(half) * 16u;
            const uint4 csc = ptx::lds128(tab), csh = ptx::lds128(tab + 32u);
            const __nv_bfloat162* sc = reinterpret_cast<const __nv_bfloat162*>(&csc);
            const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&csh);
            uint4 o;
            __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = __hfma2_relu(raw[e], sc[e], sh[e]);
            if (!keep) o = make_uint4(0u, 0u, 0u, 0u);  // outside the image / the layer's valid range: zero AFTER activation
            if (idx < 128) ptx::sts128(S.rows + uint32_t(ring_base(cp, G) + cr[k].i) * kRowBytes + off[half], o);
          }
        }
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before_sync();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&acc_free[slot]);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const int cp = k < NC - 1 ? C + 1 + k : 4;
          if (take[k]) ptx::mbar_arrive(&full[ring_base(cp, G) + cr[k].i]);
        }
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int cp = k < NC - 1 ? C + 1 + k : 4;
        if (take[k]) cr[k].step(ring_depth(cp, G));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ epilogue, transition
__device__ __forceinline__ void epilogue_transition(const FParams& P, const Smem& S, uint64_t* accT_done, uint64_t* accT_free, int q,
                                                    int lane) {
  SlotPhases dp;
  const int idx = q * 32 + lane;  // 1x1: output lane = ring index
  const uint32_t lb = S.tmem + (uint32_t(q * 32) << 16) + kTCol;
  ptx::mbar_wait(S.w_full, 0);
  const float* sb = reinterpret_cast<const float*>(__cvta_shared_to_generic(S.blob + kBiasOff)) + 4 * 16;
  const float b0 = sb[0], b1 = sb[1], b2 = sb[2];
  const size_t plane = size_t(P.H) * P.W;
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const Seg sg = decode_seg(P, item);
    const int col = sg.w0 - 4 + idx;
    const bool ok = idx >= 4 && idx < 124 && col < P.W;
    float* o = P.out + size_t(sg.n) * 3 * plane + size_t(sg.h0) * P.W + col;
    for (int r = sg.h0; r < sg.h1; ++r, o += P.W) {
      const int slot = r & 7;
      dp.wait(accT_done, slot);
      ptx::tc_fence_after_sync();
      uint32_t v[4];
      tmem_ld4(lb + 16u * uint32_t(slot), v);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&accT_free[slot]);
      if (ok) {
        const float f0 = __uint_as_float(v[0]) + b0, f1 = __uint_as_float(v[1]) + b1, f2 = __uint_as_float(v[2]) + b2;
        o[0] = 1.0f / (1.0f + __expf(-f0));
        o[plane] = 1.0f / (1.0f + __expf(-f1));
        o[2 * plane] = 1.0f / (1.0f + __expf(-f2));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ loader (F0 -> G0)
__device__ __forceinline__ void loader(const FParams& P, const Smem& S, uint64_t* full, uint64_t* empty, int t, int lane) {
  Ring gr;
  ptx::mbar_wait(S.w_full, 0);
  const uint4 sc0 = ptx::lds128(S.blob + kG0TabOff), sc1 = ptx::lds128(S.blob + kG0TabOff + 16u);
  const uint4 sh0 = ptx::lds128(S.blob + kG0TabOff + 32u), sh1 = ptx::lds128(S.blob + kG0TabOff + 48u);
  const __nv_bfloat162* sc[2] = {reinterpret_cast<const __nv_bfloat162*>(&sc0), reinterpret_cast<const __nv_bfloat162*>(&sc1)};
  const __nv_bfloat162* sh[2] = {reinterpret_cast<const __nv_bfloat162*>(&sh0), reinterpret_cast<const __nv_bfloat162*>(&sh1)};
  const uint32_t off[2] = {sw32_off(t, 0), sw32_off(t, 1)};
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
      const int rows[3] = {max(j - 1, 0), j, min(j + 1, IH - 1)};
      float tv[3][2][3];
      float xv[2][3];
      if (col_ok) {
        uint2 raw[3][2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          raw[a][0] = *reinterpret_cast<const uint2*>(tn + (size_t(rows[a]) * IW + c0) * P.t4_ld);
          raw[a][1] = *reinterpret_cast<const uint2*>(tn + (size_t(rows[a]) * IW + c1) * P.t4_ld);
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) xv[dy][ch] = xn[ch * plane + size_t(r + dy) * P.W];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            tv[a][b][0] = bflo(raw[a][b].x);
            tv[a][b][1] = bfhi(raw[a][b].x);
            tv[a][b][2] = bflo(raw[a][b].y);
          }
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        ptx::mbar_wait(&empty[gr.i], (gr.w & 1) ^ 1);
        uint4 o[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
        if (col_ok) {
          const int r0 = dy ? 1 : 0, r1 = dy ? 2 : 1;
          const float wy0 = dy ? 0.75f : 0.25f, wy1 = dy ? 0.25f : 0.75f;
          __nv_bfloat16 f[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float up = wy0 * (wx0 * tv[r0][0][ch] + wx1 * tv[r0][1][ch]) + wy1 * (wx0 * tv[r1][0][ch] + wx1 * tv[r1][1][ch]);
            f[ch] = __float2bfloat16_rn(up + xv[dy][ch]);
          }
          // K slot k = 3*consumer + channel (k = 15 unused: scale = shift = 0)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&o[hf]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = 8 * hf + 2 * e;
              __nv_bfloat162 rp;
              rp.x = f[k % 3];
              rp.y = f[(k + 1) % 3];
              a[e] = __hfma2_relu(rp, sc[hf][e], sh[hf][e]);
            }
          }
        }
        const uint32_t base = S.rows + uint32_t(gr.i) * kRowBytes;
        ptx::sts128(base + off[0], o[0]);
        ptx::sts128(base + off[1], o[1]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&full[gr.i]);
        gr.step(kG0Depth);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) dense_fused_kernel(const FParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kRowSlots], empty[kRowSlots], acc_done[16], acc_free[16], accT_done[8], accT_free[8], w_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kRowSlots; ++i) {
      ptx::mbar_init(&full[i], 4);
      ptx::mbar_init(&empty[i], i < kG0Depth ? 5 : 1);
    }
    for (int i = 0; i < 16; ++i) {
      ptx::mbar_init(&acc_done[i], 1);
      ptx::mbar_init(&acc_free[i], 4);
    }
    for (int i = 0; i < 8; ++i) {
      ptx::mbar_init(&accT_done[i], 1);
      ptx::mbar_init(&accT_free[i], 4);
    }
    ptx::mbar_init(&w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  {  // operand rows start as zeros (ring index 0 of groups 1-4 and the unused K slot are never written afterwards)
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
  S.full = ptx::smem_u32(full);
  S.empty = ptx::smem_u32(empty);
  S.acc_done = ptx::smem_u32(acc_done);
  S.acc_free = ptx::smem_u32(acc_free);
  S.accT_done = ptx::smem_u32(accT_done);
  S.accT_free = ptx::smem_u32(accT_free);
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

  if (warp == 0) issuer_layer<0>(P, S, lane);
  else if (warp == 1) issuer_layer<1>(P, S, lane);
  else if (warp == 2) issuer_layer<2>(P, S, lane);
  else if (warp == 3) issuer_layer<3>(P, S, lane);
  else if (warp < 8) loader(P, S, full, empty, (warp - 4) * 32 + lane, lane);
  else if (warp < 12) epilogue_layer<0>(P, S, full, empty, acc_done + 0, acc_free + 0, warp & 3, lane);
  else if (warp < 16) epilogue_layer<1>(P, S, full, empty, acc_done + 4, acc_free + 4, warp & 3, lane);
  else if (warp < 20) epilogue_layer<2>(P, S, full, empty, acc_done + 8, acc_free + 8, warp & 3, lane);
  else if (warp < 24) epilogue_layer<3>(P, S, full, empty, acc_done + 12, acc_free + 12, warp & 3, lane);
  else if (warp < 28) epilogue_transition(P, S, accT_done, accT_free, warp & 3, lane);
  else issuer_transition(P, S, lane);

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
  for (int g = 0; g < 5; ++g) {
    uint8_t* blk = blob.data() + kWTOff + g * 512;
    for (int co = 0; co < 16; ++co) {
      if (g == 0) {
        for (int ch = 0; ch < 3; ++ch) put_sw32(blk, co, 12 + ch, wat(layers[4], 0, ch, co));
      } else {
        for (int k = 0; k < 16; ++k) put_sw32(blk, co, k, wat(layers[4], 0, 16 * g + k, co));
      }
    }
  }
  auto put_b = [&](int off, float v) {
    const bf16 b = __float2bfloat16_rn(v);
    std::memcpy(blob.data() + off, &b, 2);
  };
  for (int cp = 1; cp <= 4; ++cp)
    for (int g = 1; g <= cp; ++g)
      for (int k = 0; k < 16; ++k) {
        put_b(kTabOff + ring_id(cp, g) * 64 + k * 2, layers[cp].pre_s[16 * g + k]);
        put_b(kTabOff + ring_id(cp, g) * 64 + 32 + k * 2, layers[cp].pre_t[16 * g + k]);
      }
  for (int cp = 0; cp <= 4; ++cp)
    for (int ch = 0; ch < 3; ++ch) {
      put_b(kG0TabOff + (3 * cp + ch) * 2, layers[cp].pre_s[ch]);
      put_b(kG0TabOff + 32 + (3 * cp + ch) * 2, layers[cp].pre_t[ch]);
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
    const long cost = ((items + sms - 1) / sms) * (seg + 18);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      P.SEG = seg;
      P.segs = nseg;
    }
  }
  P.nitems = N * P.strips * P.segs;
  CDAN_CUDA_OK(cudaFuncSetAttribute(dense_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  dense_fused_kernel<<<std::min(P.nitems, sms), kThreads, kSmemBytes, stream>>>(P);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cdan
