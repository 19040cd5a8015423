// Post-processing tail (reference utils/post_processing.py:5-77) on planar fp32 [N,3,H,W] images, and the
// PSNR / SSIM metric reductions (reference utils/metrics_factory.py:74-94 -> torchmetrics defaults, restated;
// parity unpinned because torchmetrics is an unpinned third-party dependency absent from the reference tree).
// All reductions are two-stage with a fixed order (no float atomics) -> bitwise repeatable.
// The reference's `if images.max() > 1.0: images = images / 255.0` host sync becomes a device-side flag.
#include "kernels.cuh"

namespace cdan {
namespace {

constexpr int kPlaneBlocks = 32;  // partial-reduction blocks per (n,c) plane

__device__ __forceinline__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < int(blockDim.x) / 32; ++i) r += sh[i];
  return r;  // valid in thread 0
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = -INFINITY;
  if (threadIdx.x == 0)
    for (int i = 0; i < int(blockDim.x) / 32; ++i) r = fmaxf(r, sh[i]);
  return r;
}

// grid (kPlaneBlocks, planes): partial sum / max of each plane
__global__ void __launch_bounds__(256) plane_partial_kernel(const float* __restrict__ x, int HW, float* __restrict__ psum,
                                                             float* __restrict__ pmax) {
  __shared__ float sh[8];
  const int plane = blockIdx.y, blk = blockIdx.x;
  const int chunk = (HW + kPlaneBlocks - 1) / kPlaneBlocks;
  const int p0 = blk * chunk, p1 = min(HW, p0 + chunk);
  const float* px = x + size_t(plane) * HW;
  float s = 0.f, m = -INFINITY;
  for (int p = p0 + threadIdx.x; p < p1; p += 256) {
    const float v = px[p];
    s += v;
    m = fmaxf(m, v);
  }
  const float bs = block_sum(s, sh);
  const float bm = block_max(m, sh);
  if (threadIdx.x == 0) {
    psum[plane * kPlaneBlocks + blk] = bs;
    pmax[plane * kPlaneBlocks + blk] = bm;
  }
}
// 1 block: mean[plane] (of the unit-range image) and the global unit scale (1 or 1/255)
__global__ void plane_finish_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int planes, int HW,
                                    float* __restrict__ mean, float* __restrict__ scale) {
  __shared__ float sh[8];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < planes * kPlaneBlocks; i += blockDim.x) m = fmaxf(m, pmax[i]);
  const float gm = block_max(m, sh);
  __shared__ float sc;
  if (threadIdx.x == 0) {
    sc = gm > 1.0f ? (1.0f / 255.0f) : 1.0f;
    *scale = sc;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < planes; p += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < kPlaneBlocks; ++b) s += psum[p * kPlaneBlocks + b];
    mean[p] = (s / float(HW)) * sc;
  }
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float unit(float v, float sc) { return sc == 1.0f ? v : v / 255.0f; }

__global__ void __launch_bounds__(256) contrast_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total,
                                                        int HW, const float* __restrict__ mean,
                                                        const float* __restrict__ scale, float f) {
  const float sc = *scale;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const float m = mean[i / HW];
    y[i] = clamp01((unit(x[i], sc) - m) * f + m);
  }
}
__global__ void __launch_bounds__(256) color_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW,
                                                     const float* __restrict__ scale, float f) {
  const float sc = *scale;
  const size_t total = size_t(N) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t n = i / HW, p = i % HW;
    const size_t b = n * 3 * HW + p;
    const float r = unit(x[b], sc), g = unit(x[b + HW], sc), bl = unit(x[b + 2 * size_t(HW)], sc);
    const float gray = 0.2989f * r + 0.5870f * g + 0.1140f * bl;
    y[b] = clamp01(gray + f * (r - gray));
    y[b + HW] = clamp01(gray + f * (g - gray));
    y[b + 2 * size_t(HW)] = clamp01(gray + f * (bl - gray));
  }
}
struct Stencil {
  float k[9];
  float keep, mix;  // y = clamp(keep*x + mix*conv(x))
};
__global__ void __launch_bounds__(256) stencil_kernel(const float* __restrict__ x, float* __restrict__ y, int planes,
                                                       int H, int W, const float* __restrict__ scale, Stencil st) {
  const float sc = *scale;
  const size_t total = size_t(planes) * H * W;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int xx = int(i % W), yy = int((i / W) % H);
    const float* pl = x + (i - (size_t(yy) * W + xx));
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y2 = yy + r - 1;
      if (y2 < 0 || y2 >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int x2 = xx + s - 1;
        if (x2 < 0 || x2 >= W) continue;
        acc = fmaf(st.k[r * 3 + s], unit(pl[size_t(y2) * W + x2], sc), acc);
      }
    }
    y[i] = clamp01(st.keep * unit(x[i], sc) + st.mix * acc);
  }
}

inline int grid_for(size_t total, int block = 256, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return int(g < 1 ? 1 : (g > size_t(cap) ? cap : g));
}

// ------------------------------------------------------------------------------------------------ metrics
// pass 1: per-block partials of sum((p-t)^2), min/max of p and t.   layout: part[blk][5]
__global__ void __launch_bounds__(256) metric_partial_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                              size_t total, float* __restrict__ part) {
  __shared__ float sh[8];
  float sse = 0.f, pmin = INFINITY, pmax = -INFINITY, tmin = INFINITY, tmax = -INFINITY;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const float a = p[i], b = t[i], d = a - b;
    sse = fmaf(d, d, sse);
    pmin = fminf(pmin, a); pmax = fmaxf(pmax, a);
    tmin = fminf(tmin, b); tmax = fmaxf(tmax, b);
  }
  const float r0 = block_sum(sse, sh);
  const float r1 = -block_max(-pmin, sh), r2 = block_max(pmax, sh);
  const float r3 = -block_max(-tmin, sh), r4 = block_max(tmax, sh);
  if (threadIdx.x == 0) {
    float* o = part + size_t(blockIdx.x) * 5;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4;
  }
}
// 1 thread: fold the partials.  stats = {sse, pmin, pmax, tmin, tmax, c1, c2}
__global__ void metric_finish1_kernel(const float* __restrict__ part, int nblk, float* __restrict__ stats) {
  if (threadIdx.x || blockIdx.x) return;
  double sse = 0;
  float pmin = INFINITY, pmax = -INFINITY, tmin = INFINITY, tmax = -INFINITY;
  for (int b = 0; b < nblk; ++b) {
    sse += part[b * 5];
    pmin = fminf(pmin, part[b * 5 + 1]); pmax = fmaxf(pmax, part[b * 5 + 2]);
    tmin = fminf(tmin, part[b * 5 + 3]); tmax = fmaxf(tmax, part[b * 5 + 4]);
  }
  const float dr = fmaxf(pmax - pmin, tmax - tmin);  // SSIM data_range=None -> max of the two ranges
  stats[0] = float(sse); stats[1] = pmin; stats[2] = pmax; stats[3] = tmin; stats[4] = tmax;
  stats[5] = (0.01f * dr) * (0.01f * dr);
  stats[6] = (0.03f * dr) * (0.03f * dr);
}
// pass 2: SSIM map over the valid (H-10)x(W-10) interior, 11x11 Gaussian sigma 1.5; 16x16 outputs per block.
struct Gauss11 { float g[11]; };
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ p, const float* __restrict__ t, int H, int W,
                                                    const float* __restrict__ stats, Gauss11 gw,
                                                    float* __restrict__ part) {
  __shared__ float sp[26][27], stt[26][27];
  __shared__ float sh[8];
  const int plane = blockIdx.z;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  const int VH = H - 10, VW = W - 10;
  const float* pp = p + size_t(plane) * H * W;
  const float* tp = t + size_t(plane) * H * W;
  for (int i = threadIdx.x; i < 26 * 26; i += 256) {
    const int ly = i / 26, lx = i % 26;
    const int gy = oy0 + ly, gx = ox0 + lx;
    const bool ok = gy < H && gx < W;
    sp[ly][lx] = ok ? pp[size_t(gy) * W + gx] : 0.f;
    stt[ly][lx] = ok ? tp[size_t(gy) * W + gx] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x % 16, ly = threadIdx.x / 16;
  float val = 0.f;
  if (ox0 + lx < VW && oy0 + ly < VH) {
    float mp = 0, mt = 0, spp = 0, st2 = 0, spt = 0;
#pragma unroll
    for (int r = 0; r < 11; ++r) {
      float rp = 0, rt = 0, rpp = 0, rtt = 0, rpt = 0;
#pragma unroll
      for (int s = 0; s < 11; ++s) {
        const float a = sp[ly + r][lx + s], b = stt[ly + r][lx + s], w = gw.g[s];
        rp = fmaf(w, a, rp); rt = fmaf(w, b, rt);
        rpp = fmaf(w, a * a, rpp); rtt = fmaf(w, b * b, rtt); rpt = fmaf(w, a * b, rpt);
      }
      const float w = gw.g[r];
      mp = fmaf(w, rp, mp); mt = fmaf(w, rt, mt);
      spp = fmaf(w, rpp, spp); st2 = fmaf(w, rtt, st2); spt = fmaf(w, rpt, spt);
    }
    const float c1 = stats[5], c2 = stats[6];
    const float vp = fmaxf(spp - mp * mp, 0.f), vt = fmaxf(st2 - mt * mt, 0.f), cov = spt - mp * mt;
    val = ((2.f * mp * mt + c1) * (2.f * cov + c2)) / ((mp * mp + mt * mt + c1) * (vp + vt + c2));
  }
  const float bs = block_sum(val, sh);
  if (threadIdx.x == 0) part[(size_t(plane) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = bs;
}
__global__ void metric_finish2_kernel(const float* __restrict__ stats, const float* __restrict__ part, int nblk,
                                      double numel, double nvalid, float* __restrict__ result) {
  if (threadIdx.x || blockIdx.x) return;
  // PSNR (torchmetrics default data_range=None): range from the target's min/max tracked against 0-initialised state
  const double dr = double(fmaxf(stats[4], 0.f)) - double(fminf(stats[3], 0.f));
  const double mse = double(stats[0]) / numel;
  result[0] = float(10.0 * log10(dr * dr / mse));
  double s = 0;
  for (int b = 0; b < nblk; ++b) s += part[b];
  result[1] = float(s / nvalid);
}

}  // namespace

size_t postproc_scratch_floats(int N, int H, int W) {
  (void)H; (void)W;
  return size_t(N) * 3 * (2 * kPlaneBlocks + 1) + 16;
}

int postproc_launch(int op, const float* x, float* y, int N, int H, int W, float arg, float* scratch, cudaStream_t s) {
  const int planes = N * 3, HW = H * W;
  float* psum = scratch;
  float* pmax = psum + size_t(planes) * kPlaneBlocks;
  float* mean = pmax + size_t(planes) * kPlaneBlocks;
  float* scale = mean + planes;
  if (planes > 65535) return fail("postproc: batch too large");
  plane_partial_kernel<<<dim3(kPlaneBlocks, planes), 256, 0, s>>>(x, HW, psum, pmax);
  plane_finish_kernel<<<1, 256, 0, s>>>(psum, pmax, planes, HW, mean, scale);
  const size_t total = size_t(planes) * HW;
  if (op == kContrast) {
    contrast_kernel<<<grid_for(total), 256, 0, s>>>(x, y, total, HW, mean, scale, arg);
  } else if (op == kColor) {
    color_kernel<<<grid_for(size_t(N) * HW), 256, 0, s>>>(x, y, N, HW, scale, arg);
  } else {
    Stencil st;
    if (op == kSharpen) {
      // kernel = [[0,-1,0],[-1,5,-1],[0,-1,0]]*strength + eye(3); kernel /= kernel.sum()  (post_processing.py:40-48)
      const float base[9] = {0, -1, 0, -1, 5, -1, 0, -1, 0};
      float sum = 0.f;
      for (int i = 0; i < 9; ++i) {
        st.k[i] = base[i] * arg + ((i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f);
        sum += st.k[i];
      }
      for (int i = 0; i < 9; ++i) st.k[i] /= sum;
      st.keep = 0.f;
      st.mix = 1.f;
    } else {
      const float base[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
      for (int i = 0; i < 9; ++i) st.k[i] = base[i] / 16.0f;
      st.keep = 1.0f - arg;
      st.mix = arg;
    }
    stencil_kernel<<<grid_for(total), 256, 0, s>>>(x, y, planes, H, W, scale, st);
  }
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

static int metric_blocks(size_t total) { return grid_for(total, 256, 1024); }

size_t metrics_scratch_floats(int planes, int H, int W) {
  const size_t tiles = size_t(ceil_div(W - 10, 16)) * ceil_div(H - 10, 16) * planes;
  return 1024 * 5 + 16 + tiles + 16;
}

// Output quantisation of Model._save_batch_outputs (reference models/model.py:80-83): planar fp32 [N,3,H,W] ->
// interleaved uint8 [N,H,W,3], u8 = trunc(clip(x * 255, 0, 255)) exactly as numpy's `(img * 255).clip(0, 255).astype(uint8)`
// (fp32 product, truncation toward zero).  One thread per four pixels: three coalesced float4 loads, three 32-bit stores.
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, size_t HW,
                                                           size_t quads_per_image, size_t total_quads) {
  for (size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x; q < total_quads; q += size_t(gridDim.x) * blockDim.x) {
    const size_t n = q / quads_per_image, p = (q - n * quads_per_image) * 4;
    const float* xi = x + n * 3 * HW + p;
    const float4 r = *reinterpret_cast<const float4*>(xi), g = *reinterpret_cast<const float4*>(xi + HW),
                 b = *reinterpret_cast<const float4*>(xi + 2 * HW);
    auto cv = [](float v) -> uint32_t { return uint32_t(fminf(fmaxf(v * 255.0f, 0.0f), 255.0f)); };
    const uint32_t px[12] = {cv(r.x), cv(g.x), cv(b.x), cv(r.y), cv(g.y), cv(b.y), cv(r.z), cv(g.z), cv(b.z), cv(r.w), cv(g.w), cv(b.w)};
    uint32_t* o = reinterpret_cast<uint32_t*>(y + (n * HW + p) * 3);
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = px[4 * i] | (px[4 * i + 1] << 8) | (px[4 * i + 2] << 16) | (px[4 * i + 3] << 24);
  }
}

int quantize_u8_launch(const float* x, uint8_t* y, int N, int H, int W, cudaStream_t s) {
  const size_t HW = size_t(H) * W;
  if (HW % 4 != 0) return fail("quantize_u8: H*W must be a multiple of 4");
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(y) % 4 != 0) return fail("quantize_u8: misaligned buffer");
  const size_t qpi = HW / 4, total = qpi * N;
  const int blocks = int(std::min<size_t>((total + 255) / 256, size_t(148) * 16));
  quantize_u8_kernel<<<std::max(blocks, 1), 256, 0, s>>>(x, y, HW, qpi, total);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

// Same-size input normalisation for the uint8 host path (reference data/dataset.py:86-92: the uint8 HWC image goes through
// `A.Normalize(mean 0, std 1, max_pixel_value 255)` + `ToTensorV2`): interleaved uint8 [N,H,W,3] -> planar fp32
// [N,3,H,W], x = float(u8) * float32(1/255).  One thread per four pixels: three 32-bit loads, three float4 stores.
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t HW,
                                                            size_t quads_per_image, size_t total_quads) {
  const float inv255 = 1.0f / 255.0f;
  for (size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x; q < total_quads; q += size_t(gridDim.x) * blockDim.x) {
    const size_t n = q / quads_per_image, p = (q - n * quads_per_image) * 4;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(src + (n * HW + p) * 3);
    const uint32_t w0 = in[0], w1 = in[1], w2 = in[2];
    uint32_t b[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      b[i] = (w0 >> (8 * i)) & 0xffu;
      b[4 + i] = (w1 >> (8 * i)) & 0xffu;
      b[8 + i] = (w2 >> (8 * i)) & 0xffu;
    }
    float* o = dst + n * 3 * HW + p;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(o + c * HW) = make_float4(float(b[c]) * inv255, float(b[3 + c]) * inv255, float(b[6 + c]) * inv255, float(b[9 + c]) * inv255);
  }
}

int normalize_u8_launch(const uint8_t* src, float* dst, int N, int H, int W, cudaStream_t s) {
  const size_t HW = size_t(H) * W;
  if (HW % 4 != 0) return fail("normalize_u8: H*W must be a multiple of 4");
  if (reinterpret_cast<uintptr_t>(dst) % 16 != 0 || reinterpret_cast<uintptr_t>(src) % 4 != 0) return fail("normalize_u8: misaligned buffer");
  const size_t qpi = HW / 4, total = qpi * N;
  const int blocks = int(std::min<size_t>((total + 255) / 256, size_t(148) * 16));
  normalize_u8_kernel<<<std::max(blocks, 1), 256, 0, s>>>(src, dst, HW, qpi, total);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

// Input side (reference data/dataset.py:86-92 + utils/transforms_factory.py:50-86): uint8 HWC images ->
// cv2.resize(INTER_LINEAR) -> Normalize(mean 0, std 1, max 255) -> CHW float32.  The arithmetic is OpenCV's fixed-point
// bilinear for 8-bit images (11-bit weights; resize.cpp HResizeLinear / VResizeLinear), restated and pinned bit-exactly
// against cv2 in oracle/input_oracle.py; the per-column / per-row source indices and weights are computed on the host
// with the same float/double operations OpenCV uses and arrive as tables [x0 | x1 | a0 | a1] and [y0 | y1 | b0 | b1].
__global__ void __launch_bounds__(256) resize_normalize_u8_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                                   float* __restrict__ dst, int Hd, int Wd,
                                                                   const int* __restrict__ xt, const int* __restrict__ yt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= Wd) return;
  const int x0 = xt[x], x1 = xt[Wd + x], a0 = xt[2 * Wd + x], a1 = xt[3 * Wd + x];
  const int y0 = yt[y], y1 = yt[Hd + y], b0 = yt[2 * Hd + y], b1 = yt[3 * Hd + y];
  const uint8_t* im = src + size_t(n) * Hs * Ws * 3;
  const uint8_t *r0 = im + size_t(y0) * Ws * 3, *r1 = im + size_t(y1) * Ws * 3;
  const float inv255 = 1.0f / 255.0f;  // float32 reciprocal, as albumentations.Normalize multiplies by it
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = int(r0[x0 * 3 + c]) * a0 + int(r0[x1 * 3 + c]) * a1;  // horizontal pass, scaled by 2^11
    const int h1 = int(r1[x0 * 3 + c]) * a0 + int(r1[x1 * 3 + c]) * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    dst[((size_t(n) * 3 + c) * Hd + y) * Wd + x] = float(v) * inv255;
  }
}

int resize_normalize_u8_launch(const uint8_t* src, int N, int Hs, int Ws, float* dst, int Hd, int Wd, const int* xt,
                               const int* yt, cudaStream_t s) {
  if (N > 65535 || Hd > 65535) return fail("resize_normalize_u8: batch or output height too large for one launch");
  resize_normalize_u8_kernel<<<dim3((Wd + 255) / 256, Hd, N), 256, 0, s>>>(src, Hs, Ws, dst, Hd, Wd, xt, yt);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

int psnr_ssim_launch(const float* pred, const float* target, int N, int C, int H, int W, float* scratch, float* result,
                     cudaStream_t s) {
  const int planes = N * C;
  if (planes > 65535) return fail("psnr_ssim: too many planes");
  const size_t total = size_t(planes) * H * W;
  float* part1 = scratch;
  float* stats = part1 + 1024 * 5;
  float* part2 = stats + 16;
  const int nb1 = metric_blocks(total);
  metric_partial_kernel<<<nb1, 256, 0, s>>>(pred, target, total, part1);
  metric_finish1_kernel<<<1, 32, 0, s>>>(part1, nb1, stats);
  Gauss11 gw;
  double sum = 0, g[11];
  for (int i = 0; i < 11; ++i) {
    const double d = i - 5;
    g[i] = exp(-(d * d) / (2.0 * 1.5 * 1.5));
    sum += g[i];
  }
  for (int i = 0; i < 11; ++i) gw.g[i] = float(g[i] / sum);
  dim3 grid(ceil_div(W - 10, 16), ceil_div(H - 10, 16), planes);
  ssim_kernel<<<grid, 256, 0, s>>>(pred, target, H, W, stats, gw, part2);
  const int nb2 = int(grid.x * grid.y * grid.z);
  metric_finish2_kernel<<<1, 32, 0, s>>>(stats, part2, nb2, double(total), double(planes) * (H - 10) * (W - 10), result);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cdan
