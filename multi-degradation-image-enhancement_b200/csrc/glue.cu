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

// out = (UP ? bilinear_x2(a) : a) + skip.  One block per (image, pair of output rows); a thread owns 8 channels of a
// 2x2 output quad: the quad's 3x3 input neighbourhood is loaded once (6 loads per output row pair instead of 8 per
// pixel) and no per-element 64-bit index arithmetic is needed.  Optionally the block also reduces the per-channel sum and
// max of what it wrote (ChannelGate pooling partials, models/cbam.py:41-45) — deterministic, no float atomics.
template <typename T, bool UP>
__global__ void __launch_bounds__(256) up_add_kernel(const T* __restrict__ a, int a_ld, const T* __restrict__ skip,
                                                      int skip_ld, T* __restrict__ out, int out_ld, int OH, int OW,
                                                      int C, float* __restrict__ psum, float* __restrict__ pmax) {
  extern __shared__ float red[];  // [npl][C] sums then [npl][C] maxes (only when pooling)
  const int vecs = C >> 3;
  const int npl = 256 / vecs;
  const int v = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int j = blockIdx.x, n = blockIdx.y;
  const int IH = UP ? OH >> 1 : OH, IW = UP ? OW >> 1 : OW;
  const int QW = (OW + 1) >> 1;
  float s[8], m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; m[e] = -INFINITY; }
  auto emit_sk = [&](int oy, int ox, F8 r, const F8& sk) {  // skip value loaded ahead by the caller
    const size_t pix = (size_t(n) * OH + oy) * OW + ox;
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] += sk.v[e];
    store8<T>(out + pix * out_ld + v * 8, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float q = to_f32<T>(from_f32<T>(r.v[e]));  // pool what was stored
      s[e] += q;
      m[e] = fmaxf(m[e], q);
    }
  };
  auto emit = [&](int oy, int ox, F8 r) {
    const size_t pix = (size_t(n) * OH + oy) * OW + ox;
    const F8 sk = load8<T>(skip + pix * skip_ld + v * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] += sk.v[e];
    store8<T>(out + pix * out_ld + v * 8, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float q = to_f32<T>(from_f32<T>(r.v[e]));  // pool what was stored
      s[e] += q;
      m[e] = fmaxf(m[e], q);
    }
  };
  if (pl < npl) {
    const T* an = a + size_t(n) * IH * IW * a_ld + v * 8;
    for (int i = pl; i < QW; i += npl) {
      if (UP) {
        const int jm = max(j - 1, 0), jp = min(j + 1, IH - 1), im = max(i - 1, 0), ip = min(i + 1, IW - 1);
        const F8 m0 = load8<T>(an + (size_t(j) * IW + im) * a_ld), m1 = load8<T>(an + (size_t(j) * IW + i) * a_ld),
                 m2 = load8<T>(an + (size_t(j) * IW + ip) * a_ld);
        // all four skip vectors of the quad are requested before any arithmetic (one latency instead of four)
        const T* sq = skip + ((size_t(n) * OH + 2 * j) * OW + 2 * i) * skip_ld + v * 8;
        const F8 k00 = load8<T>(sq), k01 = load8<T>(sq + skip_ld), k10 = load8<T>(sq + size_t(OW) * skip_ld),
                 k11 = load8<T>(sq + size_t(OW) * skip_ld + skip_ld);
        const F8 t0 = load8<T>(an + (size_t(jm) * IW + im) * a_ld), t1 = load8<T>(an + (size_t(jm) * IW + i) * a_ld),
                 t2 = load8<T>(an + (size_t(jm) * IW + ip) * a_ld);
        const F8 b0 = load8<T>(an + (size_t(jp) * IW + im) * a_ld), b1 = load8<T>(an + (size_t(jp) * IW + i) * a_ld),
                 b2 = load8<T>(an + (size_t(jp) * IW + ip) * a_ld);
        {
          F8 r0, r1;
#pragma unroll
          for (int e = 0; e < 8; ++e) {  // row 2j: rows (j-1: .25, j: .75); cols even (i-1: .25, i: .75), odd (i: .75, i+1: .25)
            r0.v[e] = 0.25f * (0.25f * t0.v[e] + 0.75f * t1.v[e]) + 0.75f * (0.25f * m0.v[e] + 0.75f * m1.v[e]);
            r1.v[e] = 0.25f * (0.75f * t1.v[e] + 0.25f * t2.v[e]) + 0.75f * (0.75f * m1.v[e] + 0.25f * m2.v[e]);
          }
          emit_sk(2 * j, 2 * i, r0, k00);
          emit_sk(2 * j, 2 * i + 1, r1, k01);
        }
        {
          F8 r0, r1;
#pragma unroll
          for (int e = 0; e < 8; ++e) {  // row 2j+1: rows (j: .75, j+1: .25)
            r0.v[e] = 0.75f * (0.25f * m0.v[e] + 0.75f * m1.v[e]) + 0.25f * (0.25f * b0.v[e] + 0.75f * b1.v[e]);
            r1.v[e] = 0.75f * (0.75f * m1.v[e] + 0.25f * m2.v[e]) + 0.25f * (0.75f * b1.v[e] + 0.25f * b2.v[e]);
          }
          emit_sk(2 * j + 1, 2 * i, r0, k10);
          emit_sk(2 * j + 1, 2 * i + 1, r1, k11);
        }
      } else {
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int oy = 2 * j + dy, ox = 2 * i + dx;
            if (oy < OH && ox < OW) emit(oy, ox, load8<T>(an + (size_t(oy) * IW + ox) * a_ld));
          }
      }
    }
  }
  if (psum == nullptr) return;
  if (pl < npl) {
    float* rs = red + (pl * C + v * 8);
    float* rm = red + (npl * C) + (pl * C + v * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) { rs[e] = s[e]; rm[e] = m[e]; }
  }
  __syncthreads();
  const int nblk = gridDim.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    float ss = 0.f, mm = -INFINITY;
    for (int q = 0; q < npl; ++q) {  // fixed order
      ss += red[q * C + c];
      mm = fmaxf(mm, red[npl * C + q * C + c]);
    }
    psum[(size_t(n) * nblk + j) * C + c] = ss;
    pmax[(size_t(n) * nblk + j) * C + c] = mm;
  }
}

// Final decoder stage (models/cdan.py:153-154): out[n,h,w,0:3] = bilinear_x2(a)[.,0:3] + x_nchw, channels 3..pad_to
// zero (head of the final dense block's concat buffer).  One block per (image, pair of output rows), one thread per
// 2x2 output quad: the quad's 3x3 neighbourhood of `a` is read once, x is read and out written row-contiguously.
template <typename T>
__global__ void __launch_bounds__(256) up_add_input_kernel(const T* __restrict__ a, int a_ld,
                                                            const float* __restrict__ x, T* __restrict__ out,
                                                            int out_ld, int pad_to, int OH, int OW) {
  const int j = blockIdx.x, n = blockIdx.y;
  const int IH = OH >> 1, IW = OW >> 1;
  const int jm = max(j - 1, 0), jp = min(j + 1, IH - 1);
  const T* an = a + size_t(n) * IH * IW * a_ld;
  const size_t plane = size_t(OH) * OW;
  const float* xn = x + size_t(n) * 3 * plane;
  for (int i = threadIdx.x; i < IW; i += 256) {
    const int im = max(i - 1, 0), ip = min(i + 1, IW - 1);
    float t[3][3][3];  // [row jm/j/jp][col im/i/ip][channel]
    const int rows[3] = {jm, j, jp}, cols[3] = {im, i, ip};
    // all loads of the quad are requested before any arithmetic: nine 8-channel vectors of `a` (a_ld >= 8) and the
    // two-pixel fp32 pairs of x for both output rows and the three channels
    F8 av[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) av[r][c] = load8<T>(an + (size_t(rows[r]) * IW + cols[c]) * a_ld);
    float2 xv[2][3];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        xv[dy][ch] = *reinterpret_cast<const float2*>(xn + ch * plane + size_t(2 * j + dy) * OW + 2 * i);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) t[r][c][ch] = av[r][c].v[ch];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        // same tap order and weights as the general kernel: even index -> (i-1: .25, i: .75), odd -> (i: .75, i+1: .25)
        const int r0 = dy ? 1 : 0, r1 = dy ? 2 : 1, c0 = dx ? 1 : 0, c1 = dx ? 2 : 1;
        const float wy0 = dy ? 0.75f : 0.25f, wy1 = dy ? 0.25f : 0.75f, wx0 = dx ? 0.75f : 0.25f, wx1 = dx ? 0.25f : 0.75f;
        const int oy = 2 * j + dy, ox = 2 * i + dx;
        F8 r;
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float up = wy0 * (wx0 * t[r0][c0][ch] + wx1 * t[r0][c1][ch]) + wy1 * (wx0 * t[r1][c0][ch] + wx1 * t[r1][c1][ch]);
          r.v[ch] = up + (dx ? xv[dy][ch].y : xv[dy][ch].x);
        }
        T* o = out + ((size_t(n) * OH + oy) * OW + ox) * out_ld;
        store8<T>(o, r);
        F8 z;
#pragma unroll
        for (int e = 0; e < 8; ++e) z.v[e] = 0.f;
        for (int c = 8; c < pad_to; c += 8) store8<T>(o + c, z);
      }
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
                  int OH, int OW, int C, int up, cudaStream_t s, float* psum, float* pmax) {
  if (C % 8 || C > 2048) return fail("up_add: C must be a multiple of 8 (<= 2048)");
  if (up && ((OH | OW) & 1)) return fail("up_add: upsampled extent must be even");
  if (N > 65535) return fail("up_add: batch too large for one launch");
  const dim3 grid((OH + 1) / 2, N);
  const int npl = 256 / (C / 8);
  const size_t smem = psum ? 2 * size_t(npl) * C * sizeof(float) : 0;
  if (dt == kF32) {
    if (up) up_add_kernel<float, true><<<grid, 256, smem, s>>>((const float*)a, a_ld, (const float*)skip, skip_ld, (float*)out, out_ld, OH, OW, C, psum, pmax);
    else up_add_kernel<float, false><<<grid, 256, smem, s>>>((const float*)a, a_ld, (const float*)skip, skip_ld, (float*)out, out_ld, OH, OW, C, psum, pmax);
  } else {
    if (up) up_add_kernel<bf16, true><<<grid, 256, smem, s>>>((const bf16*)a, a_ld, (const bf16*)skip, skip_ld, (bf16*)out, out_ld, OH, OW, C, psum, pmax);
    else up_add_kernel<bf16, false><<<grid, 256, smem, s>>>((const bf16*)a, a_ld, (const bf16*)skip, skip_ld, (bf16*)out, out_ld, OH, OW, C, psum, pmax);
  }
  CDAN_CUDA_OK(cudaGetLastError());
  return 0;
}

int up_add_input_launch(DType dt, const void* a, int a_ld, const float* x_nchw, void* out, int out_ld, int pad_to,
                        int N, int OH, int OW, cudaStream_t s) {
  if (pad_to % 8 || pad_to < 8) return fail("up_add_input: pad_to must be a positive multiple of 8");
  if (a_ld < 8 || a_ld % 8) return fail("up_add_input: the upsampled tensor must carry a multiple of 8 (>= 8) channels per pixel");
  if ((OH | OW) & 1) return fail("up_add_input: output extent must be even");
  if (N > 65535) return fail("up_add_input: batch too large for one launch");
  const dim3 g(OH / 2, N);
  if (dt == kF32) up_add_input_kernel<float><<<g, 256, 0, s>>>((const float*)a, a_ld, x_nchw, (float*)out, out_ld, pad_to, OH, OW);
  else up_add_input_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)a, a_ld, x_nchw, (bf16*)out, out_ld, pad_to, OH, OW);
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
