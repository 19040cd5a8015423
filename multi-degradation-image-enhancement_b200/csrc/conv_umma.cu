// tcgen05 implicit-GEMM convolution for sm_100a (3x3 pad 1 / 1x1, stride 1, NHWC bf16, fp32 accumulate in TMEM).
//
// Formulation ("flattened shifted GEMM"): a CTA owns an output tile of TH rows x TW columns of one image.  The halo'd
// input tile ((TH+2) x WP pixels, WP = TW+2) of one 64-channel K-chunk is staged ONCE in shared memory as
// [pixel][64 ch] = 128-byte rows in the canonical K-major SWIZZLE_128B layout.  Because the swizzle is a function of
// the absolute smem address (verified by csrc/probes/umma_probe.cu on B200), the A operand of tap (r,s) is simply the
// same buffer viewed from row offset r*WP+s: nine descriptor offsets replace nine im2col loads.  Output "pixel"
// p = hh*WP + ww lives in TMEM lane p%128 of M-block p/128; columns ww >= TW are garbage and masked in the epilogue.
//
// Warp roles (warp-specialised, mbarrier pipelines, persistent over tiles):
//   warp 0      : producer — bulk-copies pre-swizzled weight blocks (B ring) and, in TMA mode, issues the 4-D
//                 tensor-map tile loads of A (hardware zero fill = conv padding)
//   warp 1      : tcgen05.mma issuer (one elected lane)
//   warp 2      : TMEM allocator
//   warps 4-7   : epilogue — tcgen05.ld, bias(+folded BN), ReLU, optional 2x2 max-pool / sigmoid+NCHW, global stores
//   warps 8-15  : (producer modes only) A-tile builders: global load -> dense-block pre-activation
//                 relu(s*x+t) in fp32 -> zero padding AFTER the activation -> swizzled st.shared, or the fp32 NCHW
//                 3-channel network input -> bf16
#include "conv_umma.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptx_sm100.cuh"

namespace cdan {

struct UmmaPack {
  uint8_t* d_w = nullptr;   // [npass][chunk][tap][NT rows][128 B swizzled]
  float* d_bias = nullptr;  // [npass*NT]
  int Cin = 0, Cout = 0, ks = 3, NT = 0, npass = 1, nchunks = 0;
  StreamPack* stream = nullptr;  // streaming-kernel image of the same weights (narrow-output layers)
};

namespace {

enum InMode : int { kInTma = 0, kInPro = 1, kInNchw3 = 2 };

struct KParams {
  int N, H, W, Cin, in_ld;
  int ks, halo, taps;
  int TH, TW, WP, NMB, NT;
  int tiles_x, tiles_y, ntiles;  // ntiles = N*tiles_y*tiles_x (per N-pass)
  int nchunks, npass, Cout;
  int SA, SB, ACC;
  int a_stage_bytes, b_stage_bytes, tps, bst_per_chunk;  // tps = taps per B stage
  int a_rows;                                             // (TH+2*halo)*WP rows actually produced
  int relu, pool, sigmoid;
  const bf16* in;
  const float* in_nchw;
  const float* pre_s;
  const float* pre_t;
  const uint8_t* wpack;
  const float* bias;
  bf16* out;
  int out_ld;
  float* out_nchw;
};

constexpr int kMaxSA = 4, kMaxSB = 8;

__device__ __forceinline__ void decode_tile(const KParams& P, int t, int& n, int& h0, int& w0) {
  const int per_img = P.tiles_x * P.tiles_y;
  n = t / per_img;
  const int r = t - n * per_img;
  const int ty = r / P.tiles_x;
  h0 = ty * P.TH;
  w0 = (r - ty * P.tiles_x) * P.TW;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ epilogue role
// Shared by the one-CTA kernel and the CTA-pair kernel.  WorkFn(i, &t, &pass, &dup) maps this CTA's i-th work item to a tile
// and N-pass (dup: a padding item whose stores are masked), ReleaseFn(as) hands accumulator buffer `as` back to the MMA issuer.
// kEpiGroups groups of four warps (one per TMEM lane quarter).  The 16-column chunks of a tile's accumulators are dealt
// round-robin to the groups, and a group requests its NEXT chunk from TMEM before it processes the current one, so the
// tcgen05.ld latency is hidden behind bias / pooling / stores (measured before: ~690 cycles per chunk, one group, no
// overlap = 22k cycles per 2-M-block tile that the single-buffered accumulator could not hide behind the MMAs).
template <int kEpiGroups, typename WorkFn, typename ReleaseFn>
__device__ __forceinline__ void epilogue_role(const KParams& P, int warp, int lane, int tid, uint32_t tmem_base, float* s_bias,
                                            float* s_xbuf, uint64_t* acc_full, int nwork, WorkFn work_of, ReleaseFn release) {
  const int q = (warp - 4) & 3;   // TMEM lane quarter
  const int g = (warp - 4) >> 2;  // epilogue group
  const int gt = tid - 128 - g * 128;  // thread index inside the group
  const int bar_id = 1 + g;
  float* bias_g = s_bias + g * 256;
  float* xbuf_g = s_xbuf + g * (2 * 2 * 32 * 16);
  auto gsync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
  for (int i = gt; i < P.NT; i += 128) bias_g[i] = 0.f;
  gsync();
  int as = 0, pacc = 0;
  int cur_pass = -1;
  const int cpm = P.NT >> 4;            // chunks per M-block
  const int nchunk = P.NMB * cpm;       // chunks per tile
  for (int wi = 0; wi < nwork; ++wi) {
    int t, pass;
    bool dup;
    work_of(wi, t, pass, dup);
    int n, h0, w0;
    decode_tile(P, t, n, h0, w0);
    if (pass != cur_pass) {  // (re)load this pass's bias slice
      gsync();
      for (int i = gt; i < P.NT; i += 128) bias_g[i] = P.bias[pass * P.NT + i];
      gsync();
      cur_pass = pass;
    }
    ptx::mbar_wait(&acc_full[as], pacc);
    ptx::tc_fence_after_sync();
    const uint32_t acc_col = tmem_base + uint32_t(as * P.NMB * P.NT) + (uint32_t(q * 32) << 16);
    const int cbase = pass * P.NT;  // first output channel of this pass
    int xpar = 0;
    uint32_t raw[16];
    int ci = g;
    if (ci < nchunk) ptx::tmem_ld16(acc_col + uint32_t(ci << 4), raw);  // chunk ci sits at column ci*16 (mb*NT + c0)
    for (; ci < nchunk; ci += kEpiGroups) {
      const int mb = ci / cpm, c0 = (ci - mb * cpm) << 4;
      const int p = mb * 128 + q * 32 + lane;
      const int hh = p / P.WP, ww = p - hh * P.WP;
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = hh < P.TH && ww < P.TW && h < P.H && w < P.W && !dup;
      ptx::tmem_wait_ld();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float f = __uint_as_float(raw[j]) + bias_g[c0 + j];
        if (P.relu) f = fmaxf(f, 0.f);
        v[j] = f;
      }
      if (ci + kEpiGroups < nchunk) ptx::tmem_ld16(acc_col + uint32_t((ci + kEpiGroups) << 4), raw);
      const int cg = cbase + c0;  // global output channel of v[0]
      if (P.pool && P.WP == 16) {
        // 2x2 max-pool with 16-pixel tile rows: a warp holds two image rows, so both partners are lanes of the same warp
        // (lane^1 and lane^16) — no shared-memory exchange and no group barrier per chunk
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
          v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 16));
        }
        if (valid && !(lane & 17) && cg < P.Cout) {
          bf16* o = P.out + ((size_t(n) * (P.H >> 1) + (h >> 1)) * (P.W >> 1) + (w >> 1)) * P.out_ld + cg;
          uint4 u0, u1;
          u0.x = pack_bf16x2(v[0], v[1]); u0.y = pack_bf16x2(v[2], v[3]);
          u0.z = pack_bf16x2(v[4], v[5]); u0.w = pack_bf16x2(v[6], v[7]);
          u1.x = pack_bf16x2(v[8], v[9]); u1.y = pack_bf16x2(v[10], v[11]);
          u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
          *reinterpret_cast<uint4*>(o) = u0;
          *reinterpret_cast<uint4*>(o + 8) = u1;
        }
      } else if (P.pool) {
        // 2x2 max-pool: horizontal partner = lane^1, vertical partner = lane + WP inside this M-block
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
        const int rows_per_warp_shift = (P.WP == 64) ? 1 : 0;  // WP=64: warps (0,1)=row0,(2,3)=row1; WP=32: warp=row
        const bool upper = rows_per_warp_shift ? (q >= 2) : (q & 1);
        const int pair = rows_per_warp_shift ? (q & 1) : (q >> 1);
        float* xb = xbuf_g + (xpar * 2 + pair) * (32 * 16);
        if (upper) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xb[j * 32 + lane] = v[j];
        }
        gsync();
        if (!upper && valid && !(lane & 1) && cg < P.Cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], xb[j * 32 + lane]);
          bf16* o = P.out + ((size_t(n) * (P.H >> 1) + (h >> 1)) * (P.W >> 1) + (w >> 1)) * P.out_ld + cg;
          uint4 u0, u1;
          u0.x = pack_bf16x2(v[0], v[1]); u0.y = pack_bf16x2(v[2], v[3]);
          u0.z = pack_bf16x2(v[4], v[5]); u0.w = pack_bf16x2(v[6], v[7]);
          u1.x = pack_bf16x2(v[8], v[9]); u1.y = pack_bf16x2(v[10], v[11]);
          u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
          *reinterpret_cast<uint4*>(o) = u0;
          *reinterpret_cast<uint4*>(o + 8) = u1;
        }
        xpar ^= 1;
      } else if (P.out_nchw) {
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cg + j < P.Cout) {
              float f = v[j];
              if (P.sigmoid) f = 1.0f / (1.0f + __expf(-f));
              P.out_nchw[((size_t(n) * P.Cout + cg + j) * P.H + h) * P.W + w] = f;
            }
        }
      } else if (valid && cg < P.Cout) {
        bf16* o = P.out + ((size_t(n) * P.H + h) * P.W + w) * P.out_ld + cg;
        if (cg + 16 <= P.Cout) {
          uint4 u0, u1;
          u0.x = pack_bf16x2(v[0], v[1]); u0.y = pack_bf16x2(v[2], v[3]);
          u0.z = pack_bf16x2(v[4], v[5]); u0.w = pack_bf16x2(v[6], v[7]);
          u1.x = pack_bf16x2(v[8], v[9]); u1.y = pack_bf16x2(v[10], v[11]);
          u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
          *reinterpret_cast<uint4*>(o) = u0;
          *reinterpret_cast<uint4*>(o + 8) = u1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cg + j < P.Cout) o[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) release(as);
    if (++as == P.ACC) { as = 0; pacc ^= 1; }
  }
}

template <int IN_MODE>
__global__ void __launch_bounds__(512, 1) conv_umma_kernel(const __grid_constant__ CUtensorMap tmapA, const KParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + size_t(P.SA) * P.a_stage_bytes;
  uint8_t* sTail = sB + size_t(P.SB) * P.b_stage_bytes + 1024;  // 1 KB slack for garbage-row over-reads
  float* s_bias = reinterpret_cast<float*>(sTail);                // [2 epilogue groups][NT]
  float* s_xbuf = s_bias + 2 * 256;                               // pool exchange per epilogue group: [2][2 pairs][16 cols][32 lanes]

  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], raw_full[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB], acc_full[2],
      acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_work = P.ntiles * P.npass;
  constexpr int kProducerWarps = 8;
  constexpr int kEpiGroups = IN_MODE == kInTma ? 2 : 1;  // TMA mode has no builder warps: warps 8-11 drain accumulators too

  if (tid == 0) {
    for (int i = 0; i < P.SA; ++i) {
      ptx::mbar_init(&a_full[i], IN_MODE == kInTma ? 1 : kProducerWarps);
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&raw_full[i], 1);
    }
    for (int i = 0; i < P.SB; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 4 * kEpiGroups);
    }
    ptx::fence_mbar_init();
    if (IN_MODE != kInNchw3) ptx::prefetch_tmap(&tmapA);
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ============================================================ producer of the weight (B) ring
    if (lane == 0) {
      int sb = 0, pb = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int pass = work % P.npass;
        for (int c = 0; c < P.nchunks; ++c) {
          const uint8_t* wsrc = P.wpack + (size_t(pass) * P.nchunks + c) * P.taps * (size_t(P.NT) * 128);
          for (int j = 0; j < P.bst_per_chunk; ++j) {
            const int ntap = min(P.tps, P.taps - j * P.tps);
            const uint32_t bytes = uint32_t(ntap) * P.NT * 128u;
            ptx::mbar_wait(&b_empty[sb], pb ^ 1);
            ptx::mbar_arrive_expect_tx(&b_full[sb], bytes);
            ptx::bulk_g2s(sB + size_t(sb) * P.b_stage_bytes, wsrc + size_t(j) * P.tps * P.NT * 128, bytes, &b_full[sb]);
            if (++sb == P.SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ============================================================ producer of the activation (A) tiles via TMA.  Its own
    // thread: behind the weight ring in one program order, the tile of chunk c+1 was requested only ~4 weight stages
    // (~2 us) before the MMAs needed it — less than a loaded DRAM round trip — and the tensor pipe idled on a_full.
    if (IN_MODE != kInNchw3 && lane == 0) {
      int sa = 0, pa = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int t = work / P.npass;
        int n, h0, w0;
        decode_tile(P, t, n, h0, w0);
        for (int c = 0; c < P.nchunks; ++c) {
          // TMA mode: the tile feeds the MMA directly; producer mode: it lands "raw" and warps 8-15 activate it
          uint64_t* full = IN_MODE == kInTma ? &a_full[sa] : &raw_full[sa];
          ptx::mbar_wait(&a_empty[sa], pa ^ 1);
          ptx::mbar_arrive_expect_tx(full, uint32_t(P.a_rows) * 128u);
          ptx::tma_load_4d(sA + size_t(sa) * P.a_stage_bytes, &tmapA, c * 64, w0 - P.halo, h0 - P.halo, n, full);
          if (++sa == P.SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    // The whole warp walks the pipeline (so the waits are convergent); one elected lane issues.  Descriptors are
    // kept as {lo, hi} with a constant hi word, so each tcgen05.mma costs two 32-bit adds of issue overhead.
    const uint32_t idesc = ptx::umma_idesc_bf16(128, P.NT);
    const uint64_t desc_hi = ptx::umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
    const uint32_t desc_lo_flags = uint32_t(ptx::umma_desc_sw128(0, 1024) & 0xffffffffull);  // LBO field
    const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
    int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      ptx::mbar_wait(&acc_empty[as], pacc ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t acc_col = tmem_base + uint32_t(as * P.NMB * P.NT);
      for (int c = 0; c < P.nchunks; ++c) {
        const int ksteps = min(4, (P.Cin - c * 64 + 15) >> 4);
        ptx::mbar_wait(&a_full[sa], pa);
        const uint32_t a_base = sA_u + uint32_t(sa) * P.a_stage_bytes;
        for (int j = 0; j < P.bst_per_chunk; ++j) {
          ptx::mbar_wait(&b_full[sb], pb);
          ptx::tc_fence_after_sync();
          const uint32_t b_base = sB_u + uint32_t(sb) * P.b_stage_bytes;
          const int ntap = min(P.tps, P.taps - j * P.tps);
          if (ptx::elect_one()) {
            for (int tl = 0; tl < ntap; ++tl) {
              const int tap = j * P.tps + tl;
              const int r = tap / P.ks, s = tap - r * P.ks;
              // descriptor address fields are in 16-byte units: one pixel row = 8, one M-block = 1024, one K step = 2
              uint32_t a_lo = desc_lo_flags | (((a_base + uint32_t(r * P.WP + s) * 128u) & 0x3FFFFu) >> 4);
              const uint32_t b_lo = desc_lo_flags | (((b_base + uint32_t(tl) * P.NT * 128u) & 0x3FFFFu) >> 4);
              const uint32_t first = (c | tap) != 0 ? 1u : 0u;
              uint32_t d = acc_col;
              for (int mb = 0; mb < P.NMB; ++mb) {
                ptx::umma_bf16(d, desc_hi | a_lo, desc_hi | b_lo, idesc, first);
                if (ksteps > 1) ptx::umma_bf16(d, desc_hi | (a_lo + 2), desc_hi | (b_lo + 2), idesc, 1u);
                if (ksteps > 2) ptx::umma_bf16(d, desc_hi | (a_lo + 4), desc_hi | (b_lo + 4), idesc, 1u);
                if (ksteps > 3) ptx::umma_bf16(d, desc_hi | (a_lo + 6), desc_hi | (b_lo + 6), idesc, 1u);
                a_lo += 1024;
                d += uint32_t(P.NT);
              }
            }
            ptx::umma_commit(&b_empty[sb]);                        // frees the weight stage once these MMAs retire
            if (j == P.bst_per_chunk - 1) {
              ptx::umma_commit(&a_empty[sa]);                      // ... and the activation stage after its last tap
              if (c == P.nchunks - 1) ptx::umma_commit(&acc_full[as]);
            }
          }
          __syncwarp();
          if (++sb == P.SB) { sb = 0; pb ^= 1; }
        }
        if (++sa == P.SA) { sa = 0; pa ^= 1; }
      }
      if (++as == P.ACC) { as = 0; pacc ^= 1; }
    }
  } else if (warp >= 4 && warp < 4 + 4 * kEpiGroups) {
    // ============================================================ epilogue (epilogue_role above)
    const int nwork = (total_work - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    epilogue_role<kEpiGroups>(
        P, warp, lane, tid, tmem_base, s_bias, s_xbuf, acc_full, blockIdx.x < unsigned(total_work) ? nwork : 0,
        [&](int wi, int& t, int& pass, bool& dup) {
          const int work = int(blockIdx.x) + wi * int(gridDim.x);
          t = work / P.npass;
          pass = work - t * P.npass;
          dup = false;
        },
        [&](int as) { ptx::mbar_arrive(&acc_empty[as]); });
  } else if (warp >= 8) {
    // ============================================================ A-tile builders (producer modes; in TMA mode warps 8-11 are
    // the second epilogue group, handled above)
    if (IN_MODE != kInTma) {
      const int pw = warp - 8;
      int sa = 0, pa = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int t = work / P.npass;
        int n, h0, w0;
        decode_tile(P, t, n, h0, w0);
        for (int c = 0; c < P.nchunks; ++c) {
          uint8_t* dst = sA + size_t(sa) * P.a_stage_bytes;
          if (IN_MODE == kInNchw3) {
            // fp32 planar 3-channel network input -> one 16-channel group (3 real + 13 zero) per pixel.
            // Loads of a whole batch of pixels are issued before any is consumed (memory-level parallelism).
            ptx::mbar_wait(&a_empty[sa], pa ^ 1);
            const size_t plane = size_t(P.H) * P.W;
            const float* img = P.in_nchw + size_t(n) * 3 * plane;
            constexpr int U = 4;
            for (int q0 = pw * 32 + lane; q0 < P.a_rows; q0 += kProducerWarps * 32 * U) {
              float x[U][3];
#pragma unroll
              for (int i = 0; i < U; ++i) {
                const int q = q0 + i * kProducerWarps * 32;
                const int rr = q / P.WP, wi = q - rr * P.WP;
                const int h = h0 - P.halo + rr, w = w0 - P.halo + wi;
                const bool ok = q < P.a_rows && h >= 0 && h < P.H && w >= 0 && w < P.W;
                const size_t o = ok ? size_t(h) * P.W + w : 0;
                x[i][0] = ok ? img[o] : 0.f;
                x[i][1] = ok ? img[o + plane] : 0.f;
                x[i][2] = ok ? img[o + 2 * plane] : 0.f;
              }
#pragma unroll
              for (int i = 0; i < U; ++i) {
                const int q = q0 + i * kProducerWarps * 32;
                if (q < P.a_rows) {
                  *reinterpret_cast<uint4*>(dst + ptx::sw128_offset(uint32_t(q), 0)) =
                      make_uint4(pack_bf16x2(x[i][0], x[i][1]), pack_bf16x2(x[i][2], 0.f), 0u, 0u);
                  *reinterpret_cast<uint4*>(dst + ptx::sw128_offset(uint32_t(q), 1)) = make_uint4(0u, 0u, 0u, 0u);
                }
              }
            }
          } else {
            // The raw NHWC bf16 tile was delivered by TMA (zero outside the image / beyond Cin).  Apply the dense
            // block pre-activation relu(s*x+t) IN PLACE in fp32; pixels outside the image stay zero, i.e. the conv
            // padding is applied after the activation (reference models/cdan.py:41-46).
            const int cch = min(64, P.Cin - c * 64);
            const int ksteps = (cch + 15) >> 4;
            const int upp = ksteps <= 1 ? 2 : (ksteps == 2 ? 4 : 8);  // 16-byte units per pixel the MMA reads
            const int u = lane & (upp - 1), psub = lane / upp, ppi = 32 / upp;
            const int ch0 = c * 64 + u * 8;
            float sc[8], sh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const bool cv = ch0 + j < P.Cin;
              sc[j] = cv ? P.pre_s[ch0 + j] : 0.f;
              sh[j] = cv ? P.pre_t[ch0 + j] : 0.f;
            }
            const int step = kProducerWarps * ppi;
            const int step_r = step / P.WP, step_w = step - step_r * P.WP;
            int q = pw * ppi + psub;
            int rr = q / P.WP, wi = q - rr * P.WP;
            ptx::mbar_wait(&raw_full[sa], pa);
            if (ch0 < P.Cin) {
              for (; q < P.a_rows; q += step) {
                const int h = h0 - P.halo + rr, w = w0 - P.halo + wi;
                if (h >= 0 && h < P.H && w >= 0 && w < P.W) {
                  uint4* ptr = reinterpret_cast<uint4*>(dst + ptx::sw128_offset(uint32_t(q), uint32_t(u)));
                  const uint4 raw = *ptr;
                  const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
                  uint32_t ow[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float lo = __uint_as_float(rw[i] << 16), hi = __uint_as_float(rw[i] & 0xffff0000u);
                    ow[i] = pack_bf16x2(fmaxf(fmaf(lo, sc[2 * i], sh[2 * i]), 0.f),
                                        fmaxf(fmaf(hi, sc[2 * i + 1], sh[2 * i + 1]), 0.f));
                  }
                  *ptr = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                wi += step_w;
                rr += step_r;
                if (wi >= P.WP) { wi -= P.WP; ++rr; }
              }
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[sa]);
          if (++sa == P.SA) { sa = 0; pa ^= 1; }
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

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// The TMA-fed tile kernel on CTA pairs (clusters of two, tcgen05 cta_group::2; ptx_sm100.cuh "CTA pairs").  Each CTA of a pair
// owns ITS OWN output tile (tiles 2w and 2w+1) and stages that tile's activations exactly as above; the pair shares the
// weights: CTA r loads rows [r NT/2, (r+1) NT/2) of every tap block through a 2-D tensor map, and one M = 256 MMA issued by
// the leader multiplies both tiles by the full block.  Per CTA the weight fill traffic into shared memory and the B operand
// reads of the tensor pipe are halved.  All TMA completions are counted on the leader's `full` barriers (cta_group::2 loads),
// stage releases and accumulator hand-offs reach both CTAs through multicast commits, and the peer's epilogue warps release
// accumulators with remote arrivals on the leader's barrier.
__global__ void __launch_bounds__(384, 1) conv_umma_pair_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                                 const __grid_constant__ CUtensorMap tmapB, const KParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + size_t(P.SA) * P.a_stage_bytes;
  uint8_t* sTail = sB + size_t(P.SB) * P.b_stage_bytes + 1024;
  float* s_bias = reinterpret_cast<float*>(sTail);
  float* s_xbuf = s_bias + 2 * 256;

  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = int(blockIdx.x >> 1), npairs = int(gridDim.x >> 1);
  const int total_work = ((P.ntiles + 1) >> 1) * P.npass;  // items of the pair: (tile pair, N-pass)
  const int half_rows = P.NT >> 1;

  if (tid == 0) {
    for (int i = 0; i < P.SA; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < P.SB; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 16);  // eight epilogue warps of each CTA
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmapA);
    ptx::prefetch_tmap(&tmapB);
  }
  if (warp == 2) {
    ptx::tmem_alloc2(&tmem_base_s, 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's barriers exist before anything is signalled on them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  auto tile_of = [&](int work, int& t, int& pass, bool& dup) {
    const int tp = work / P.npass;
    pass = work - tp * P.npass;
    t = 2 * tp + int(rank);
    dup = t >= P.ntiles;  // odd tile count: the last pair computes the last tile twice, the copy's stores are masked
    if (dup) t = P.ntiles - 1;
  };

  if (warp == 0) {
    // ============================================================ weights: this CTA's half of every tap block
    if (lane == 0) {
      int sb = 0, pb = 0;
      for (int work = pair; work < total_work; work += npairs) {
        const int pass = work % P.npass;
        for (int c = 0; c < P.nchunks; ++c)
          for (int j = 0; j < P.taps; ++j) {
            ptx::mbar_wait(&b_empty[sb], pb ^ 1);
            if (leader) ptx::mbar_arrive_expect_tx(&b_full[sb], 2u * uint32_t(half_rows) * 128u);
            const int row = ((pass * P.nchunks + c) * P.taps + j) * P.NT + int(rank) * half_rows;
            ptx::tma2_load_2d(sB + size_t(sb) * P.b_stage_bytes, &tmapB, 0, row, &b_full[sb]);
            if (++sb == P.SB) { sb = 0; pb ^= 1; }
          }
      }
    }
  } else if (warp == 3) {
    // ============================================================ activations: this CTA's tile
    if (lane == 0) {
      int sa = 0, pa = 0;
      for (int work = pair; work < total_work; work += npairs) {
        int t, pass, n, h0, w0;
        bool dup;
        tile_of(work, t, pass, dup);
        decode_tile(P, t, n, h0, w0);
        for (int c = 0; c < P.nchunks; ++c) {
          ptx::mbar_wait(&a_empty[sa], pa ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&a_full[sa], 2u * uint32_t(P.a_rows) * 128u);
          ptx::tma2_load_4d(sA + size_t(sa) * P.a_stage_bytes, &tmapA, c * 64, w0 - P.halo, h0 - P.halo, n, &a_full[sa]);
          if (++sa == P.SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (leader CTA only): M = 256 over both tiles
    if (leader) {
      const uint32_t idesc = ptx::umma_idesc_bf16(256, P.NT);
      const uint64_t desc_hi = ptx::umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
      const uint32_t desc_lo_flags = uint32_t(ptx::umma_desc_sw128(0, 1024) & 0xffffffffull);
      const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
      int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
      for (int work = pair; work < total_work; work += npairs) {
        ptx::mbar_wait(&acc_empty[as], pacc ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t acc_col = tmem_base + uint32_t(as * P.NMB * P.NT);
        for (int c = 0; c < P.nchunks; ++c) {
          const int ksteps = min(4, (P.Cin - c * 64 + 15) >> 4);
          ptx::mbar_wait(&a_full[sa], pa);
          const uint32_t a_base = sA_u + uint32_t(sa) * P.a_stage_bytes;
          for (int tap = 0; tap < P.taps; ++tap) {
            ptx::mbar_wait(&b_full[sb], pb);
            ptx::tc_fence_after_sync();
            const uint32_t b_base = sB_u + uint32_t(sb) * P.b_stage_bytes;
            if (ptx::elect_one()) {
              const int r = tap / P.ks, s = tap - r * P.ks;
              uint32_t a_lo = desc_lo_flags | (((a_base + uint32_t(r * P.WP + s) * 128u) & 0x3FFFFu) >> 4);
              const uint32_t b_lo = desc_lo_flags | ((b_base & 0x3FFFFu) >> 4);
              const uint32_t first = (c | tap) != 0 ? 1u : 0u;
              uint32_t d = acc_col;
              for (int mb = 0; mb < P.NMB; ++mb) {
                ptx::umma2_bf16(d, desc_hi | a_lo, desc_hi | b_lo, idesc, first);
                if (ksteps > 1) ptx::umma2_bf16(d, desc_hi | (a_lo + 2), desc_hi | (b_lo + 2), idesc, 1u);
                if (ksteps > 2) ptx::umma2_bf16(d, desc_hi | (a_lo + 4), desc_hi | (b_lo + 4), idesc, 1u);
                if (ksteps > 3) ptx::umma2_bf16(d, desc_hi | (a_lo + 6), desc_hi | (b_lo + 6), idesc, 1u);
                a_lo += 1024;
                d += uint32_t(P.NT);
              }
              ptx::umma2_commit_both(&b_empty[sb]);
              if (tap == P.taps - 1) {
                ptx::umma2_commit_both(&a_empty[sa]);
                if (c == P.nchunks - 1) ptx::umma2_commit_both(&acc_full[as]);
              }
            }
            __syncwarp();
            if (++sb == P.SB) { sb = 0; pb ^= 1; }
          }
          if (++sa == P.SA) { sa = 0; pa ^= 1; }
        }
        if (++as == P.ACC) { as = 0; pacc ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue: each CTA drains its own tile
    const int nwork = pair < total_work ? (total_work - pair + npairs - 1) / npairs : 0;
    const uint32_t leader_acc_empty = ptx::mapa_u32(ptx::smem_u32(&acc_empty[0]), 0);
    epilogue_role<2>(
        P, warp, lane, tid, tmem_base, s_bias, s_xbuf, acc_full, nwork,
        [&](int wi, int& t, int& pass, bool& dup) { tile_of(pair + wi * npairs, t, pass, dup); },
        [&](int as) { ptx::mbar_arrive_cluster(leader_acc_empty + 8u * uint32_t(as)); });
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // no CTA leaves (or frees tensor memory) while its peer may still signal it
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
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

constexpr int kSmemLimit = 232448 - 1024;  // 227 KB opt-in maximum minus alignment slack

int pick_nt(int Cout, int* npass) {
  *npass = 1;
  if (Cout <= 16) return 16;
  if (Cout <= 32) return 32;
  if (Cout <= 64) return 64;
  if (Cout <= 128) return 128;
  static const int nt_cap = getenv("CDAN_UMMA_NT") ? atoi(getenv("CDAN_UMMA_NT")) : 256;  // A/B switch: 128 = more M-blocks per weight pass
  if (nt_cap == 128) {
    *npass = (Cout + 127) / 128;
    return 128;
  }
  *npass = (Cout + 255) / 256;
  return 256;
}

struct TileCfg {
  int TH, TW, WP, NMB, ACC, SA, SB, tps, a_stage_bytes, b_stage_bytes, smem_bytes;
};

// Choose the tile geometry: minimise issued MMA rows (tiles * NMB * 128) subject to shared-memory capacity.
bool choose_tiles(const ConvDesc& d, int NT, int in_mode, TileCfg* out, bool pair = false) {
  const int halo = d.ks / 2, taps = d.ks * d.ks;
  const int nchunks = (d.Cin + 63) / 64;
  TileCfg best{};
  double best_cost = 1e300;
  static const int nmb_max = getenv("CDAN_UMMA_NMB_MAX") ? atoi(getenv("CDAN_UMMA_NMB_MAX")) : 8;  // A/B switch
  static const int nmb_min = getenv("CDAN_UMMA_NMB_MIN") ? atoi(getenv("CDAN_UMMA_NMB_MIN")) : 1;
  for (int NMB = 1; NMB * NT <= 512 && NMB <= nmb_max; NMB *= 2) {
    if (NMB < nmb_min && 2 * NMB * NT <= 512) continue;
    const int ACC = (2 * NMB * NT <= 512) ? 2 : 1;
    std::vector<int> wps;
    if (d.pool) {
      // 16-pixel tile rows keep both pooling partners inside a warp (epilogue_role: no shared-memory exchange, no group
      // barrier), but only 14 of 16 columns are outputs: measured slower for encoder.conv3 (2.26 vs 2.05 ms with 32-pixel
      // rows) — the pooled epilogue is not what limits the pair kernel.  Kept behind the A/B switch CDAN_UMMA_POOL_WP=16.
      static const int pool_wp = getenv("CDAN_UMMA_POOL_WP") ? atoi(getenv("CDAN_UMMA_POOL_WP")) : 0;
      if (pool_wp == 16 || pool_wp == 32 || pool_wp == 64) wps = {pool_wp};
      else wps = {32, 64};
    } else {
      for (int wp = 2 * halo + 2; wp <= std::min(256, d.W + 2 * halo + 6); wp += 2) wps.push_back(wp);
      if (d.W + 2 * halo <= 256) wps.push_back(d.W + 2 * halo);
    }
    for (int WP : wps) {
      int TW = WP - 2 * halo;
      if (TW < 1) continue;
      int TH = (NMB * 128) / WP;
      if (TH < 1) continue;
      if (d.pool && (TH & 1)) continue;
      TH = std::min(TH, d.pool ? ((d.H + 1) & ~1) : d.H);
      if (TH < 1) continue;
      const int a_rows_alloc = NMB * 128 + 2 * halo * WP + 8;
      const int a_stage = int(align_up(size_t(a_rows_alloc) * 128, 1024));
      const int tps = NT <= 32 ? taps : (NT == 64 ? std::min(taps, 3) : 1);
      const int b_stage = tps * (pair ? NT / 2 : NT) * 128;  // CTA pairs: each CTA holds half of a tap block
      const int tail = 1024 + 2 * 256 * 4 + 2 * (2 * 2 * 32 * 16 * 4) + 64;
      // pipeline depth: as many stages as fit, capped
      int SA = (nchunks >= 2 || true) ? 2 : 1, SB = 4;
      auto total = [&](int sa, int sb) { return sa * a_stage + sb * b_stage + tail; };
      while (SB > 2 && total(SA, SB) > kSmemLimit) --SB;
      if (total(SA, SB) > kSmemLimit) continue;
      while (SA < 3 && total(SA + 1, SB) <= kSmemLimit && in_mode != kInTma) ++SA;
      static const int sb_max = getenv("CDAN_UMMA_SB_MAX") ? atoi(getenv("CDAN_UMMA_SB_MAX")) : 6;  // A/B switch
      while (SB < kMaxSB && total(SA, SB + 1) <= kSmemLimit && SB < sb_max) ++SB;
      const double tiles = double((d.W + TW - 1) / TW) * ((d.H + TH - 1) / TH);
      // per-tile time model (cycles): tensor pipe vs L2->smem operand traffic, plus the epilogue when the
      // accumulator is single-buffered (then it cannot overlap the next tile's MMAs)
      const int ksteps_total = (d.Cin + 15) / 16;
      const double mma_cyc = double(NMB) * taps * ksteps_total * std::max(16.0, NT / 2.0);
      const double l2_bytes = double(nchunks) * (double(TH + 2 * halo) * WP * 128.0 + double(taps) * (pair ? NT / 2 : NT) * 128.0);
      // epilogue per 16-column chunk.  The fused 2x2 max-pool (shuffles + a shared-memory exchange and a group barrier per
      // chunk) is expensive enough that it must overlap the next tile's MMAs: encoder.conv3 runs 2.28 ms with double-buffered
      // accumulators (NMB = 1) against 2.67 ms with NMB = 2 (r02 A/B, CDAN_UMMA_NMB_MAX); the un-pooled layers measured the
      // other way round (conv4 1.98 / 2.08, decoder.conv1 1.78 / 2.01 ms: halving the weight traffic into shared memory wins)
      const double epi_cyc = double(NMB) * (NT / 16.0) * (d.pool ? (WP == 16 ? 500.0 : 1200.0) : 70.0) + 300.0;
      // operand bytes per cycle and SM the fills sustain next to the MMA's own operand reads; CTA pairs measured a little
      // better with two M-blocks per weight stage on the un-pooled layers (conv4 1.79 / 1.85 ms), hence the lower rate there
      const double fill_rate = pair ? 30.0 : 36.0;
      const double cost = tiles * (std::max(mma_cyc, l2_bytes / fill_rate) + (ACC == 1 ? epi_cyc : 0.15 * epi_cyc) + 200.0);
      if (cost < best_cost) {
        best_cost = cost;
        best = TileCfg{TH, TW, WP, NMB, ACC, SA, SB, tps, a_stage, b_stage, total(SA, SB) + 1024};
      }
    }
  }
  if (best_cost >= 1e300) return false;
  *out = best;
  return true;
}

}  // namespace

int umma_pack_create(const float* w, const float* bias, int Cin, int Cout, int CoutP, int ks, UmmaPack** out) {
  *out = nullptr;
  if (ks != 1 && ks != 3) return fail("umma_pack: ks must be 1 or 3");
  UmmaPack* p = new UmmaPack();
  p->Cin = Cin; p->Cout = Cout; p->ks = ks;
  p->NT = pick_nt(Cout, &p->npass);
  p->nchunks = (Cin + 63) / 64;
  const int taps = ks * ks, NT = p->NT;
  const size_t block = size_t(NT) * 128;
  std::vector<uint8_t> img(size_t(p->npass) * p->nchunks * taps * block, 0);
  std::vector<float> hb(size_t(p->npass) * NT, 0.f);
  for (int pass = 0; pass < p->npass; ++pass)
    for (int c = 0; c < p->nchunks; ++c)
      for (int t = 0; t < taps; ++t) {
        uint8_t* blk = img.data() + ((size_t(pass) * p->nchunks + c) * taps + t) * block;
        for (int nn = 0; nn < NT; ++nn) {
          const int co = pass * NT + nn;
          if (co >= Cout) continue;
          for (int k = 0; k < 64; ++k) {
            const int ci = c * 64 + k;
            if (ci >= Cin) continue;
            const float val = w[(size_t(t) * Cin + ci) * CoutP + co];
            const bf16 b = __float2bfloat16_rn(val);
            std::memcpy(blk + ptx::sw128_offset(uint32_t(nn), uint32_t(k >> 3)) + (k & 7) * 2, &b, 2);
          }
        }
      }
  for (int co = 0; co < Cout; ++co) hb[co] = bias[co];
  if (cudaMalloc(&p->d_w, img.size()) != cudaSuccess || cudaMalloc(&p->d_bias, hb.size() * 4) != cudaSuccess) {
    umma_pack_destroy(p);
    return fail("umma_pack: cudaMalloc failed");
  }
  if (cudaMemcpy(p->d_w, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    umma_pack_destroy(p);
    return fail("umma_pack: upload failed");
  }
  if (stream_pack_create(w, bias, Cin, Cout, CoutP, ks, &p->stream) != 0) {
    umma_pack_destroy(p);
    return -1;
  }
  *out = p;
  return 0;
}

void umma_pack_destroy(UmmaPack* p) {
  if (!p) return;
  stream_pack_destroy(p->stream);
  if (p->d_w) cudaFree(p->d_w);
  if (p->d_bias) cudaFree(p->d_bias);
  delete p;
}

bool conv_umma_supported(const ConvDesc& d) {
  if (d.ks != 1 && d.ks != 3) return false;
  if (d.in_nchw) return d.Cin <= 16 && !d.pre_scale;  // 3-channel planar network input
  if (d.Cin % 8 != 0 || d.in_ld % 8 != 0) return false;
  if (d.out_nchw) return d.Cout <= 16;
  if (d.out_ld % 8 != 0) return false;
  if (d.pool && ((d.H | d.W) & 1)) return false;
  return true;
}

int conv_umma_kernel_count(const ConvDesc& d, const UmmaPack& pk) {
  static const bool use_stream = !(getenv("CDAN_CONV_STREAM") && atoi(getenv("CDAN_CONV_STREAM")) == 0);
  if (use_stream && pk.stream && conv_stream_supported(d, *pk.stream)) return conv_stream_kernel_count(d, *pk.stream);
  return 1;
}

int conv_umma_launch(const ConvDesc& d, const UmmaPack& pk, cudaStream_t stream) {
  if (!conv_umma_supported(d)) return fail("conv_umma: unsupported convolution shape");
  if (pk.Cin != d.Cin || pk.Cout != d.Cout || pk.ks != d.ks) return fail("conv_umma: weight pack does not match");
  static const bool use_stream = !(getenv("CDAN_CONV_STREAM") && atoi(getenv("CDAN_CONV_STREAM")) == 0);
  if (use_stream && pk.stream && conv_stream_supported(d, *pk.stream)) return conv_stream_launch(d, *pk.stream, stream);
  if (d.in_gstride) return fail("conv_umma: group-planar input is only implemented by the streaming kernel");
  const int in_mode = d.in_nchw ? kInNchw3 : (d.pre_scale ? kInPro : kInTma);
  // CTA pairs (cta_group::2) for the TMA-fed layers with wide outputs; CDAN_UMMA_PAIR=0 keeps the one-CTA kernel (A/B switch)
  static const bool pair_enabled = !(getenv("CDAN_UMMA_PAIR") && atoi(getenv("CDAN_UMMA_PAIR")) == 0);
  const bool pair = pair_enabled && in_mode == kInTma && (pk.NT == 256 || pk.NT == 128);
  TileCfg tc;
  if (!choose_tiles(d, pk.NT, in_mode, &tc, pair)) return fail("conv_umma: no tile configuration fits shared memory");

  KParams P{};
  P.N = d.N; P.H = d.H; P.W = d.W; P.Cin = d.Cin; P.in_ld = d.in_ld;
  P.ks = d.ks; P.halo = d.ks / 2; P.taps = d.ks * d.ks;
  P.TH = tc.TH; P.TW = tc.TW; P.WP = tc.WP; P.NMB = tc.NMB; P.NT = pk.NT;
  P.tiles_x = ceil_div(d.W, tc.TW); P.tiles_y = ceil_div(d.H, tc.TH);
  P.ntiles = d.N * P.tiles_x * P.tiles_y;
  P.nchunks = pk.nchunks; P.npass = pk.npass; P.Cout = d.Cout;
  P.SA = tc.SA; P.SB = tc.SB; P.ACC = tc.ACC;
  P.a_stage_bytes = tc.a_stage_bytes; P.b_stage_bytes = tc.b_stage_bytes;
  P.tps = tc.tps; P.bst_per_chunk = ceil_div(P.taps, tc.tps);
  P.a_rows = (tc.TH + 2 * P.halo) * tc.WP;
  P.relu = d.relu; P.pool = d.pool; P.sigmoid = d.sigmoid;
  P.in = reinterpret_cast<const bf16*>(d.in); P.in_nchw = d.in_nchw;
  P.pre_s = d.pre_scale; P.pre_t = d.pre_shift;
  P.wpack = pk.d_w; P.bias = pk.d_bias;
  P.out = reinterpret_cast<bf16*>(d.out); P.out_ld = d.out_ld; P.out_nchw = d.out_nchw;

  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  if (in_mode != kInNchw3) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail("conv_umma: cuTensorMapEncodeTiled is not available from the driver");
    if (reinterpret_cast<uintptr_t>(d.in) % 16 != 0) return fail("conv_umma: input pointer must be 16-byte aligned");
    cuuint64_t gdim[4] = {cuuint64_t(d.Cin), cuuint64_t(d.W), cuuint64_t(d.H), cuuint64_t(d.N)};
    cuuint64_t gstr[3] = {cuuint64_t(d.in_ld) * 2, cuuint64_t(d.W) * d.in_ld * 2, cuuint64_t(d.H) * d.W * d.in_ld * 2};
    cuuint32_t box[4] = {64, cuuint32_t(tc.WP), cuuint32_t(tc.TH + 2 * P.halo), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("conv_umma: cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (pair) {
    PFN_encodeTiled enc = get_encode();
    CUtensorMap tmapB;
    std::memset(&tmapB, 0, sizeof(tmapB));
    const cuuint64_t rows = cuuint64_t(pk.npass) * pk.nchunks * P.taps * pk.NT;
    cuuint64_t gdim[2] = {64, rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, cuuint32_t(pk.NT / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint8_t*>(pk.d_w), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("conv_umma: cuTensorMapEncodeTiled (weights) failed with code " + std::to_string(int(r)));
    const int pair_work = ((P.ntiles + 1) / 2) * P.npass;
    CDAN_CUDA_OK(cudaFuncSetAttribute(conv_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc.smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * std::max(1, sms / 2));
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = size_t(tc.smem_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent pairs: as many clusters as can be resident at once (both CTAs of a pair sit on one TPC)
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, conv_umma_pair_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
      cudaGetLastError();
      max_clusters = sms / 2;
    }
    cfg.gridDim = dim3(2 * std::max(1, std::min(pair_work, max_clusters)));
    static const bool verbose = getenv("CDAN_UMMA_VERBOSE") != nullptr;
    if (verbose)
      fprintf(stderr, "conv_umma pair: Cin=%d Cout=%d H=%d W=%d N=%d TH=%d TW=%d NMB=%d ACC=%d SA=%d SB=%d clusters=%d (max resident %d)\n", d.Cin,
              d.Cout, d.H, d.W, d.N, tc.TH, tc.TW, tc.NMB, tc.ACC, tc.SA, tc.SB, int(cfg.gridDim.x / 2), max_clusters);
    CDAN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_umma_pair_kernel, tmap, tmapB, P));
    return 0;
  }
  const int grid = std::min(P.ntiles * P.npass, sms);
  const int threads = in_mode == kInTma ? 384 : 512;  // TMA mode: 4 role warps + two epilogue groups
  auto launch = [&](auto kern) -> int {
    CDAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc.smem_bytes));
    kern<<<grid, threads, tc.smem_bytes, stream>>>(tmap, P);
    CDAN_CUDA_OK(cudaGetLastError());
    return 0;
  };
  if (in_mode == kInTma) return launch(conv_umma_kernel<kInTma>);
  if (in_mode == kInPro) return launch(conv_umma_kernel<kInPro>);
  return launch(conv_umma_kernel<kInNchw3>);
}

}  // namespace cdan
