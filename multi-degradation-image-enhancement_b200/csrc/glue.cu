// Decoder glue and layout conversion: HBM-bound, one 8-channel vector (16 B bf16 / 32 B fp32) per thread,
// consecutive threads on consecutive channel vectors of the same pixel -> fully coalesced NHWC access.
#include "kernels.cuh"

namespace cdan {
namespace {

// Bilinear x2, align_corners=False (models/cdan.py:137,145,153): output index o reads input
// i = o>>1 and its neighbour (i-1 for even o, i+1 for odd o, clamped) with weights 0.75 / 0.25.
__device__ __forceinline__ void up_taps(int o, int n_in, int& i0, int& i1, float& w0, float& w1) {
  const int i = o >> 1;
  if (o & 1) {
    i0 = i; i1 = min(i + 1, n_in - 1); w0 = 0.75f; w1 = 0.25f;
  } else {
    i0 = max(i - 1, 0); i1 = i; w0 = 0.25f; w1 = 0.75f;
  }
}

template <typename T, bool UP>
__global__ void __launch_bounds__(256) up_add_kernel(const T* __restrict__ a, int a_ld, const T* __restrict__ skip,
                                                      int skip_ld, T* __restrict__ out, int out_ld, int N, int OH,
                                                      int OW, int C) {
  const int vecs = C >> 3;
  const size_t total = size_t(N) * OH * OW * vecs;
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
    const int v = int(idx % vecs);
    const size_t pix = idx / vecs;
    const int ox = int(pix % OW);
    const int oy = int((pix / OW) % OH);
    const int n = int(pix / (size_t(OW) * OH));
    F8 r;
    if (UP) {
      const int IH = OH >> 1, IW = OW >> 1;
      int y0, y1, x0, x1;
      float wy0, wy1, wx0, wx1;
      up_taps(oy, IH, y0, y1, wy0, wy1);
      up_taps(ox, IW, x0, x1, wx0, wx1);
      const T* base = a + size_t(n) * IH * IW * a_ld + v * 8;
      const F8 a00 = load8<T>(base + (size_t(y0) * IW + x0) * a_ld);
      const F8 a01 = load8<T>(base + (size_t(y0) * IW + x1) * a_ld);
      const F8 a10 = load8<T>(base + (size_t(y1) * IW + x0) * a_ld);
      const F8 a11 = load8<T>(base + (size_t(y1) * IW + x1) * a_ld);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        r.v[j] = wy0 * (wx0 * a00.v[j] + wx1 * a01.v[j]) + wy1 * (wx0 * a10.v[j] + wx1 * a11.v[j]);
    } else {
      r = load8<T>(a + pix * a_ld + v * 8);
    }
    const F8 sk = load8<T>(skip + pix * skip_ld + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] += sk.v[j];
    store8<T>(out + pix * out_ld + v * 8, r);
  }
}

// Final decoder stage: 3 real channels; writes a zero-padded 16-channel pixel (head of the final dense buffer).
template <typename T>
__global__ void __launch_bounds__(256) up_add_input_kernel(const T* __restrict__ a, int a_ld,
                                                            const float* __restrict__ x, T* __restrict__ out,
                                                            int out_ld, int pad_to, int N, int OH, int OW) {
  const size_t total = size_t(N) * OH * OW;
  const int IH = OH >> 1, IW = OW >> 1;
  for (size_t pix = blockIdx.x * size_t(blockDim.x) + threadIdx.x; pix < total; pix += size_t(gridDim.x) * blockDim.x) {
    const int ox = int(pix % OW);
    const int oy = int((pix / OW) % OH);
    const int n = int(pix / (size_t(OW) * OH));
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    up_taps(oy, IH, y0, y1, wy0, wy1);
    up_taps(ox, IW, x0, x1, wx0, wx1);
    const T* base = a + size_t(n) * IH * IW * a_ld;
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a00 = to_f32<T>(base[(size_t(y0) * IW + x0) * a_ld + c]);
      const float a01 = to_f32<T>(base[(size_t(y0) * IW + x1) * a_ld + c]);
      const float a10 = to_f32<T>(base[(size_t(y1) * IW + x0) * a_ld + c]);
      const float a11 = to_f32<T>(base[(size_t(y1) * IW + x1) * a_ld + c]);
      const float up = wy0 * (wx0 * a00 + wx1 * a01) + wy1 * (wx0 * a10 + wx1 * a11);
      r.v[c] = up + x[((size_t(n) * 3 + c) * OH + oy) * OW + ox];
    }
    T* o = out + pix * out_ld;
    store8<T>(o, r);
    F8 z;
#pragma unroll
    for (int j = 0; j < 8; ++j) z.v[j] = 0.f;
    for (int c = 8; c < pad_to; c += 8) store8<T>(o + c, z);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int dst_ld, int N, int C,
                                    int H, int W) {
  const size_t total = size_t(N) * H * W * C;
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const size_t pix = idx / C;
    const size_t hw = pix % (size_t(H) * W);
    const size_t n = pix / (size_t(H) * W);
    dst[pix * dst_ld + c] = from_f32<T>(src[(n * C + c) * size_t(H) * W + hw]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int src_ld, float* __restrict__ dst, int N, int C,
                                    int H, int W) {
  const size_t total = size_t(N) * H * W * C;
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
    const size_t hw = idx % (size_t(H) * W);
    const int c = int((idx / (size_t(H) * W)) % C);
    const size_t n = idx / (size_t(H) * W * C);
    dst[idx] = to_f32<T>(src[(n * size_t(H) * W + hw) * src_ld + c]);
  }
}

inline int grid_for(size_t total, int block = 256, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return int(g < 1 ? 1 : (g > size_t(cap) ? cap : g));
}

}  // namespace

int up_add_launch(DType dt, const void* a, int a_ld, const void* skip, int skip_ld, void* out, int out_ld, int N,
                  int OH, int OW, int C, int up, cudaStream_t s) {
  if (C % 8) return fail("up_add: C must be a multiple of 8");
  if (up && ((OH | OW) & 1)) return fail("up_add: upsampled extent must be even");
  const size_t total = size_t(N) * OH * OW * (C / 8);
  const int g = grid_for(total);
  if (dt == kF32) {
    if (up) up_add_kernel<float, true><<<g, 256, 0, s>>>((const float*)a, a_ld, (const float*)skip, skip_ld, (float*)out, out_ld, N, OH, OW, C);
    else up_add_kernel<float, false><<<g, 256, 0, s>>>((const float*)a, a_ld, (const float*)skip, skip_ld, (float*)out, out_ld, N, OH, OW, C);
  } else {
    if (up) up_add_kernel<bf16, true><<<g, 256, 0, s>>>((const bf16*)a, a_ld, (const bf16*)skip, skip_ld, (bf16*)out, out_ld, N, OH, OW, C);
    else up_add_kernel<bf16, false><<<g, 256, 0, s>>>((const bf16*)a, a_ld, (const bf16*)skip, skip_ld, (bf16*)out, out_ld, N, OH, OW, C);
  }
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

int up_add_input_launch(DType dt, const void* a, int a_ld, const float* x_nchw, void* out, int out_ld, int pad_to,
                        int N, int OH, int OW, cudaStream_t s) {
  if (pad_to % 8 || pad_to < 8) return fail("up_add_input: pad_to must be a positive multiple of 8");
  const int g = grid_for(size_t(N) * OH * OW);
  if (dt == kF32) up_add_input_kernel<float><<<g, 256, 0, s>>>((const float*)a, a_ld, x_nchw, (float*)out, out_ld, pad_to, N, OH, OW);
  else up_add_input_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)a, a_ld, x_nchw, (bf16*)out, out_ld, pad_to, N, OH, OW);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

int nchw_to_nhwc_launch(DType dt, const float* src, void* dst, int dst_ld, int N, int C, int H, int W, cudaStream_t s) {
  const int g = grid_for(size_t(N) * C * H * W);
  if (dt == kF32) nchw_to_nhwc_kernel<float><<<g, 256, 0, s>>>(src, (float*)dst, dst_ld, N, C, H, W);
  else nchw_to_nhwc_kernel<bf16><<<g, 256, 0, s>>>(src, (bf16*)dst, dst_ld, N, C, H, W);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}
int nhwc_to_nchw_launch(DType dt, const void* src, int src_ld, float* dst, int N, int C, int H, int W, cudaStream_t s) {
  const int g = grid_for(size_t(N) * C * H * W);
  if (dt == kF32) nhwc_to_nchw_kernel<float><<<g, 256, 0, s>>>((const float*)src, src_ld, dst, N, C, H, W);
  else nhwc_to_nchw_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)src, src_ld, dst, N, C, H, W);
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cdan
