// CBAM attention (models/cbam.py:26-95) as HBM-bound kernels over NHWC activations.
//   pass 1  pool_partial : per-(n,c) sum and max over one pair of image rows (deterministic 2-stage reduction, no float
//                          atomics -> bitwise repeatable like the reference under cudnn.deterministic).  In the decoder
//                          this pass is fused into the producer (glue.cu up_add writes the same partials).
//   pass 2  gate_mlp     : finish the reduction, shared MLP on avg and max, sigmoid  -> gate[n][c]
//   pass 3  compress     : per-pixel channel max / mean of x*gate (warp-shuffle reduction) -> comp[n,h,w,2]
//   pass 4  spatial_gate : sigmoid(bn(conv7x7(comp)))                                  -> sgate[n,h,w]
//   pass 5  apply        : out = ((x*gate)*sgate) [* dense]
// (Fusing passes 3-5 into one tiled or streaming kernel was measured and is slower on B200: the tiles resident at once
// exceed the L2, and a per-strip streaming version is latency-bound between its block barriers.)
#include <algorithm>

#include "band.cuh"
#include "kernels.cuh"

namespace cdan {
namespace {


template <typename T>
__global__ void __launch_bounds__(256) pool_partial_kernel(const T* __restrict__ x, int ld, int C, int HW, int W, int nblk,
                                                            float* __restrict__ psum, float* __restrict__ pmax, size_t img_pixels) {
  extern __shared__ float red[];  // [npl][C] sums then [npl][C] maxes
  const int vecs = C >> 3;
  const int npl = 256 / vecs;  // pixel lanes
  const int vc = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int blk = blockIdx.x, n = blockIdx.y;
  const int chunk = 2 * W;  // one block per pair of image rows (same partition as the fused producer in glue.cu)
  const int p0 = blk * chunk, p1 = min(HW, p0 + chunk);
  float s[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; m[j] = -INFINITY; }
  if (pl < npl) {
    const T* base = x + size_t(n) * img_pixels * ld + vc * 8;  // HW = pixels pooled (a row range in band mode)
    for (int p = p0 + pl; p < p1; p += npl) {
      const F8 v = load8<T>(base + size_t(p) * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v.v[j]; m[j] = fmaxf(m[j], v.v[j]); }
    }
    float* rs = red + (pl * C + vc * 8);
    float* rm = red + (npl * C) + (pl * C + vc * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { rs[j] = s[j]; rm[j] = m[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float ss = 0.f, mm = -INFINITY;
    for (int q = 0; q < npl; ++q) {  // fixed order
      ss += red[q * C + c];
      mm = fmaxf(mm, red[npl * C + q * C + c]);
    }
    psum[(size_t(n) * nblk + blk) * C + c] = ss;
    pmax[(size_t(n) * nblk + blk) * C + c] = mm;
  }
}

// Band mode: finish the reduction over this band's partials -> pooled sums [N][C] | maxima [N][C] (then all-reduced).
__global__ void __launch_bounds__(256) pool_finish_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int nblk,
                                                           int C, int N, float* __restrict__ pooled) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ss = 0.f, mm = -INFINITY;
    for (int b = 0; b < nblk; ++b) {
      ss += psum[(size_t(n) * nblk + b) * C + c];
      mm = fmaxf(mm, pmax[(size_t(n) * nblk + b) * C + c]);
    }
    pooled[size_t(n) * C + c] = ss;
    pooled[size_t(N + n) * C + c] = mm;
  }
}

// One block per image.  att = mlp(avg) + mlp(max), mlp = Linear(C,C/16) -> ReLU -> Linear(C/16,C)
// (models/cbam.py:30-35,41-45,54-57: the second bias is added once per pooled branch).
__global__ void __launch_bounds__(256) gate_mlp_kernel(const float* __restrict__ psum, const float* __restrict__ pmax,
                                                        int nblk, int C, int HW, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float sm[];  // avg[C], mx[C], h_avg[R], h_max[R], red[2 * blockDim] (C < blockDim only)
  const int R = C / 16;
  float* avg = sm;
  float* mx = sm + C;
  float* ha = sm + 2 * C;
  float* hm = ha + R;
  const int n = blockIdx.x;
  if (C < int(blockDim.x)) {
    // fewer channels than threads (C = 64, 128): blockDim / C threads share a channel, each sums every parts-th partial
    // (one 1080p image at 1/2 resolution has 270 partials per channel — a serial chain that long is most of this kernel's
    // 20-60 us), then the parts are combined in a fixed order
    float* red = hm + R;  // [parts][2][C]
    const int parts = int(blockDim.x) / C, c = threadIdx.x % C, part = threadIdx.x / C;
    float ss = 0.f, mm = -INFINITY;
    if (part < parts)
      for (int b = part; b < nblk; b += parts) {
        ss += psum[(size_t(n) * nblk + b) * C + c];
        mm = fmaxf(mm, pmax[(size_t(n) * nblk + b) * C + c]);
      }
    if (part < parts) {
      red[(part * 2 + 0) * C + c] = ss;
      red[(part * 2 + 1) * C + c] = mm;
    }
    __syncthreads();
    if (threadIdx.x < C) {
      float s2 = 0.f, m2 = -INFINITY;
      for (int q = 0; q < parts; ++q) {
        s2 += red[(q * 2 + 0) * C + threadIdx.x];
        m2 = fmaxf(m2, red[(q * 2 + 1) * C + threadIdx.x]);
      }
      avg[threadIdx.x] = s2 / float(HW);
      mx[threadIdx.x] = m2;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float ss = 0.f, mm = -INFINITY;
      for (int b = 0; b < nblk; ++b) {
        ss += psum[(size_t(n) * nblk + b) * C + c];
        mm = fmaxf(mm, pmax[(size_t(n) * nblk + b) * C + c]);
      }
      avg[c] = ss / float(HW);
      mx[c] = mm;
    }
  }
  __syncthreads();
  // hidden layer: one warp per hidden unit (round-robin), lanes stride over C
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  for (int r = warp; r < R; r += nwarp) {
    float da = 0.f, dm = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = w1[r * C + c];
      da = fmaf(w, avg[c], da);
      dm = fmaf(w, mx[c], dm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      da += __shfl_xor_sync(0xffffffffu, da, o);
      dm += __shfl_xor_sync(0xffffffffu, dm, o);
    }
    if (lane == 0) {
      ha[r] = fmaxf(da + b1[r], 0.f);
      hm[r] = fmaxf(dm + b1[r], 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float oa = b2[c], om = b2[c];
    for (int r = 0; r < R; ++r) {
      const float w = w2[c * R + r];
      oa = fmaf(w, ha[r], oa);
      om = fmaf(w, hm[r], om);
    }
    gate[size_t(n) * C + c] = 1.0f / (1.0f + expf(-(oa + om)));
  }
}

// ChannelPool of the channel-gated tensor (models/cbam.py:68-70): comp[...,0] = max_c, comp[...,1] = mean_c.
template <typename T>
__global__ void __launch_bounds__(256) compress_kernel(const T* __restrict__ x, int ld, int C, int HW, size_t npix,
                                                        const float* __restrict__ gate, float* __restrict__ comp) {
  const int vecs = C >> 3;
  const int lpp = vecs < 32 ? vecs : 32;  // lanes per pixel (power of two: C in {64,...,512})
  const int ppw = 32 / lpp;               // pixels per warp
  const int lane = threadIdx.x % 32;
  const int sub = lane % lpp, wp = lane / lpp;
  const size_t warp_global = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) / 32;
  const size_t nwarps = size_t(gridDim.x) * blockDim.x / 32;
  for (size_t p0 = warp_global * ppw; p0 < npix; p0 += nwarps * ppw) {
    const size_t pix = p0 + wp;
    float mx = -INFINITY, sum = 0.f;
    if (pix < npix) {
      const int n = int(pix / HW);
      const T* px = x + pix * ld;
      const float* g = gate + size_t(n) * C;
      for (int v = sub; v < vecs; v += lpp) {
        const F8 a = load8<T>(px + v * 8);
        const float4 g0 = *reinterpret_cast<const float4*>(g + v * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(g + v * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = a.v[j] * gg[j];
          mx = fmaxf(mx, t);
          sum += t;
        }
      }
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (sub == 0 && pix < npix) {
      comp[pix * 2 + 0] = mx;
      comp[pix * 2 + 1] = sum / float(C);
    }
  }
}

// Same reduction (identical summation order) for C <= 512, restructured for bandwidth: one image per blockIdx.y so the
// channel gate of a lane stays in registers, and U pixels per lane and iteration so U x VPL independent 16-byte loads
// are in flight per thread (the one-load-per-iteration form above ran at 2.3 TB/s).
template <typename T, int VPL, int U>
__global__ void __launch_bounds__(256) compress2_kernel(const T* __restrict__ x, int ld, int C, int HW,
                                                         const float* __restrict__ gate, float* __restrict__ comp) {
  const int vecs = C >> 3;
  const int lpp = vecs < 32 ? vecs : 32;  // lanes per pixel
  const int ppw = 32 / lpp;               // pixels per warp and step
  const int lane = threadIdx.x % 32, sub = lane % lpp, wp = lane / lpp;
  const int n = blockIdx.y;
  float gg[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float* g = gate + size_t(n) * C + (sub + i * lpp) * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
    gg[i][0] = g0.x; gg[i][1] = g0.y; gg[i][2] = g0.z; gg[i][3] = g0.w;
    gg[i][4] = g1.x; gg[i][5] = g1.y; gg[i][6] = g1.z; gg[i][7] = g1.w;
  }
  const T* xi = x + size_t(n) * HW * ld;
  float2* ci = reinterpret_cast<float2*>(comp) + size_t(n) * HW;
  const int wi = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, nw = gridDim.x * (blockDim.x / 32);
  const float inv_c = 1.0f / float(C);
  for (int p0 = wi * ppw * U; p0 < HW; p0 += nw * ppw * U) {
    F8 a[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = p0 + u * ppw + wp;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (pix < HW) a[u][i] = load8<T>(xi + size_t(pix) * ld + (sub + i * lpp) * 8);
        else
#pragma unroll
          for (int j = 0; j < 8; ++j) a[u][i].v[j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = p0 + u * ppw + wp;
      float mx = -INFINITY, sum = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = a[u][i].v[j] * gg[i][j];
          mx = fmaxf(mx, t);
          sum += t;
        }
      for (int o = lpp >> 1; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
      }
      if (sub == 0 && pix < HW) ci[pix] = make_float2(mx, sum * inv_c);
    }
  }
}

// SpatialGate (models/cbam.py:72-82): 7x7 conv over the 2-channel map, zero pad 3, no bias, BN(1), sigmoid.
// One 16x64 tile per block: the tile + 3-pixel halo of comp is staged in shared memory once (zeros outside the image),
// then every thread evaluates four pixels from shared memory (98 FMAs each).
constexpr int kSgTH = 16, kSgTW = 64, kSgHH = kSgTH + 6, kSgHW = kSgTW + 6;
__global__ void __launch_bounds__(256) spatial_gate_kernel(const float* __restrict__ comp, const float* __restrict__ w7,
                                                            float bn_a, float bn_b, float* __restrict__ sgate, int H,
                                                            int W) {
  __shared__ float2 s_c[kSgHH * kSgHW];
  __shared__ float w[98];
  const int n = blockIdx.z, y0 = blockIdx.y * kSgTH, x0 = blockIdx.x * kSgTW;
  if (threadIdx.x < 98) w[threadIdx.x] = w7[threadIdx.x];
  const float2* cn = reinterpret_cast<const float2*>(comp) + size_t(n) * H * W;
  for (int q = threadIdx.x; q < kSgHH * kSgHW; q += 256) {
    const int hh = q / kSgHW, ww = q - hh * kSgHW;
    const int y = y0 - 3 + hh, x = x0 - 3 + ww;
    s_c[q] = (y >= 0 && y < H && x >= 0 && x < W) ? cn[size_t(y) * W + x] : make_float2(0.f, 0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int q = threadIdx.x + 256 * k;
    const int hh = q / kSgTW, ww = q - hh * kSgTW;
    const int y = y0 + hh, x = x0 + ww;
    if (y >= H || x >= W) continue;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        const float2 cp = s_c[(hh + r) * kSgHW + ww + c];
        acc = fmaf(w[r * 7 + c], cp.x, acc);
        acc = fmaf(w[49 + r * 7 + c], cp.y, acc);
      }
    sgate[(size_t(n) * H + y) * W + x] = 1.0f / (1.0f + expf(-(bn_a * acc + bn_b)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) apply_kernel(const T* __restrict__ x, int ld, const float* __restrict__ gate,
                                                     const float* __restrict__ sgate, const T* __restrict__ mul,
                                                     int mul_ld, T* __restrict__ out, int out_ld, int C, int HW,
                                                     size_t npix) {
  const int vecs = C >> 3;
  const size_t total = npix * vecs;
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
    const int v = int(idx % vecs);
    const size_t pix = idx / vecs;
    const int n = int(pix / HW);
    const F8 a = load8<T>(x + pix * ld + v * 8);
    const float* g = gate + size_t(n) * C + v * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(g);
    const float4 g1 = *reinterpret_cast<const float4*>(g + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float sg = sgate[pix];
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = (a.v[j] * gg[j]) * sg;
    if (mul) {
      const F8 m = load8<T>(mul + pix * mul_ld + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] *= m.v[j];
    }
    store8<T>(out + pix * out_ld + v * 8, r);
  }
}

inline int grid_for(size_t total, int block = 256, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return int(g < 1 ? 1 : (g > size_t(cap) ? cap : g));
}

template <typename T>
int cbam_typed(const void* x, int x_ld, const void* mul, int mul_ld, void* out, int out_ld, int N, int H, int W, int C,
               const CbamWeights& wt, const CbamScratch& sc, bool pooled, cudaStream_t s, const CbamBand* band) {
  const int HW = H * W;
  const int vecs = C / 8;
  const int npl = 256 / vecs;
  if (band) {
    // statistics of the owned rows only (the partials a fused producer wrote cover halo rows too and are not used)
    const int nblk = cbam_pool_blocks(band->rows);
    pool_partial_kernel<T><<<dim3(nblk, N), 256, 2 * npl * C * sizeof(float), s>>>(
        (const T*)x + size_t(band->row0) * W * x_ld, x_ld, C, band->rows * W, W, nblk, sc.psum, sc.pmax, size_t(HW));
    CDAN_CUDA_OK(cudaGetLastError());
    pool_finish_kernel<<<N, 256, 0, s>>>(sc.psum, sc.pmax, nblk, C, N, sc.pooled);
    CDAN_CUDA_OK(cudaGetLastError());
    CDAN_TRY(band->comm->allreduce_sum_max(sc.pooled, sc.pooled + size_t(N) * C, N * C, s));
    gate_mlp_kernel<<<N, 256, (2 * C + 2 * (C / 16) + 512) * sizeof(float), s>>>(sc.pooled, sc.pooled + size_t(N) * C, 1, C, band->HW_full,
                                                                           wt.w1, wt.b1, wt.w2, wt.b2, sc.gate);
    CDAN_CUDA_OK(cudaGetLastError());
  } else {
    if (!pooled) {
      pool_partial_kernel<T><<<dim3(sc.nblk, N), 256, 2 * npl * C * sizeof(float), s>>>((const T*)x, x_ld, C, HW, W, sc.nblk,
                                                                                      sc.psum, sc.pmax, size_t(HW));
      CDAN_CUDA_OK(cudaGetLastError());
    }
    gate_mlp_kernel<<<N, 256, (2 * C + 2 * (C / 16) + 512) * sizeof(float), s>>>(sc.psum, sc.pmax, sc.nblk, C, HW, wt.w1, wt.b1,
                                                                           wt.w2, wt.b2, sc.gate);
    CDAN_CUDA_OK(cudaGetLastError());
  }
  const size_t npix = size_t(N) * HW;
  const int lpp = vecs < 32 ? vecs : 32;
  if (C <= 512) {  // lanes cover the channels of a pixel in at most two 16-byte vectors each
    // pixels per lane and iteration: 4 with one vector per lane, 2 with two (C = 512); 8 / 4 measured slower
    const int U = vecs <= 32 ? 4 : 2;
    const int ppw = 32 / lpp, bx = std::max(1, std::min(ceil_div(HW, 8 * ppw * U), ceil_div(148 * 8, N)));
    if (vecs <= 32) compress2_kernel<T, 1, 4><<<dim3(bx, N), 256, 0, s>>>((const T*)x, x_ld, C, HW, sc.gate, sc.comp);
    else compress2_kernel<T, 2, 2><<<dim3(bx, N), 256, 0, s>>>((const T*)x, x_ld, C, HW, sc.gate, sc.comp);
  } else {
    compress_kernel<T><<<grid_for(npix * lpp), 256, 0, s>>>((const T*)x, x_ld, C, HW, npix, sc.gate, sc.comp);
  }
  CDAN_CUDA_OK(cudaGetLastError());
  spatial_gate_kernel<<<dim3(ceil_div(W, kSgTW), ceil_div(H, kSgTH), N), 256, 0, s>>>(sc.comp, wt.w7, wt.bn_a, wt.bn_b, sc.sgate, H, W);
  CDAN_CUDA_OK(cudaGetLastError());
  apply_kernel<T><<<grid_for(npix * vecs), 256, 0, s>>>((const T*)x, x_ld, sc.gate, sc.sgate, (const T*)mul, mul_ld,
                                                        (T*)out, out_ld, C, HW, npix);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int cbam_pool_blocks(int H) { return (H + 1) / 2; }  // one partial per pair of image rows

size_t cbam_scratch_floats(int N, int C, int H, int W) {
  return 2 * size_t(N) * cbam_pool_blocks(H) * C + 3 * size_t(N) * C + 3 * size_t(N) * H * W + 64;
}

void cbam_scratch_carve(float* base, int N, int C, int H, int W, CbamScratch* sc) {
  sc->nblk = cbam_pool_blocks(H);
  const size_t part = size_t(N) * sc->nblk * C;
  sc->psum = base;
  sc->pmax = base + part;
  sc->gate = sc->pmax + part;
  sc->pooled = sc->gate + size_t(N) * C;
  size_t off = 2 * part + 3 * size_t(N) * C;
  off = (off + 3) / 4 * 4;  // 16-byte alignment for the float2 reads of comp
  sc->comp = base + off;
  sc->sgate = sc->comp + 2 * size_t(N) * H * W;
}

int cbam_launch(DType dt, const void* x, int x_ld, const void* mul, int mul_ld, void* out, int out_ld, int N, int H,
                int W, int C, const CbamWeights& wt, const CbamScratch& sc, bool pooled, cudaStream_t s, const CbamBand* band) {
  if (C % 64 != 0 || C > 2048) return fail("cbam: gate_channels must be a multiple of 64 (<= 2048) for the CUDA path");
  if ((C & (C - 1)) != 0) return fail("cbam: gate_channels must be a power of two for the CUDA path");
  if (N > 65535) return fail("cbam: batch or image too large for one launch");
  return dt == kF32 ? cbam_typed<float>(x, x_ld, mul, mul_ld, out, out_ld, N, H, W, C, wt, sc, pooled, s, band)
                    : cbam_typed<bf16>(x, x_ld, mul, mul_ld, out, out_ld, N, H, W, C, wt, sc, pooled, s, band);
}

}  // namespace cdan
