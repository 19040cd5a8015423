// Single-operator C-ABI entry points (unit tests): fp32 NCHW device tensors in/out, converted internally to the
// dtype's NHWC layout so that exactly the production kernels run in between.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/cdan_b200.h"
#include "conv.cuh"
#include "conv_umma.cuh"
#include "kernels.cuh"

using namespace cdan;

namespace {

struct Scratch {  // frees on scope exit
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  int alloc(void** out, size_t bytes) {
    CDAN_CUDA_OK(cudaMalloc(out, bytes ? bytes : 16));
    ptrs.push_back(*out);
    return 0;
  }
};

size_t esize(int dtype) { return dtype == CDAN_DTYPE_F32 ? 4 : 2; }

}  // namespace

extern "C" {

int cdan_op_conv2d(int dtype, int impl, void* stream, const float* x, int N, int Cin, int H, int W, const float* w,
                   const float* bias, int Cout, int ks, const float* pre_scale, const float* pre_shift, int relu,
                   int pool, float* y) {
  if (!x || !w || !y) return fail("cdan_op_conv2d: NULL argument");
  if (ks != 1 && ks != 3) return fail("cdan_op_conv2d: ks must be 1 or 3");
  cudaStream_t s = (cudaStream_t)stream;
  const DType dt = DType(dtype);
  Scratch sc;
  // A 3-channel 3x3 convolution on the tcgen05 path reads the caller's planar fp32 tensor directly, exactly like
  // encoder.conv1 does inside cdan_forward (no NHWC staging copy).
  const bool nchw_in = Cin == 3 && ks == 3 && !pre_scale && impl == 0 && dt == kBF16;
  const int CinP = nchw_in ? Cin : int(align_up(size_t(Cin), 16)), CoutP = int(align_up(size_t(Cout), 16));
  const int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W;
  void *xin, *yout;
  float *dw, *db, *dps = nullptr, *dpt = nullptr;
  CDAN_TRY(sc.alloc(&xin, size_t(N) * H * W * CinP * esize(dtype)));
  CDAN_TRY(sc.alloc(&yout, size_t(N) * OH * OW * CoutP * esize(dtype)));
  CDAN_CUDA_OK(cudaMemsetAsync(xin, 0, size_t(N) * H * W * CinP * esize(dtype), s));
  if (!nchw_in) CDAN_TRY(nchw_to_nhwc_launch(dt, x, xin, CinP, N, Cin, H, W, s));
  // pack weights on the host: [taps][CinP][CoutP]
  const int taps = ks * ks;
  std::vector<float> hw(size_t(Cout) * Cin * taps), hb(CoutP, 0.f), pw(size_t(taps) * CinP * CoutP, 0.f);
  CDAN_CUDA_OK(cudaMemcpyAsync(hw.data(), w, hw.size() * 4, cudaMemcpyDefault, s));
  if (bias) CDAN_CUDA_OK(cudaMemcpyAsync(hb.data(), bias, size_t(Cout) * 4, cudaMemcpyDefault, s));
  std::vector<float> hps(CinP, 0.f), hpt(CinP, 0.f);
  if (pre_scale) {
    CDAN_CUDA_OK(cudaMemcpyAsync(hps.data(), pre_scale, size_t(Cin) * 4, cudaMemcpyDefault, s));
    CDAN_CUDA_OK(cudaMemcpyAsync(hpt.data(), pre_shift, size_t(Cin) * 4, cudaMemcpyDefault, s));
  }
  CDAN_CUDA_OK(cudaStreamSynchronize(s));
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) pw[(size_t(t) * CinP + c) * CoutP + o] = hw[(size_t(o) * Cin + c) * taps + t];
  CDAN_TRY(sc.alloc((void**)&dw, pw.size() * 4));
  CDAN_TRY(sc.alloc((void**)&db, hb.size() * 4));
  CDAN_CUDA_OK(cudaMemcpyAsync(dw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice, s));
  CDAN_CUDA_OK(cudaMemcpyAsync(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice, s));
  if (pre_scale) {
    CDAN_TRY(sc.alloc((void**)&dps, size_t(CinP) * 4));
    CDAN_TRY(sc.alloc((void**)&dpt, size_t(CinP) * 4));
    CDAN_CUDA_OK(cudaMemcpyAsync(dps, hps.data(), size_t(CinP) * 4, cudaMemcpyHostToDevice, s));
    CDAN_CUDA_OK(cudaMemcpyAsync(dpt, hpt.data(), size_t(CinP) * 4, cudaMemcpyHostToDevice, s));
  }
  ConvDesc d;
  d.N = N; d.H = H; d.W = W; d.Cin = CinP; d.Cout = Cout; d.ks = ks;
  d.in = xin; d.in_ld = CinP;
  if (nchw_in) d.in_nchw = x;
  d.pre_scale = dps; d.pre_shift = dpt;
  d.w = dw; d.CoutP = CoutP; d.bias = db;
  d.relu = relu; d.pool = pool;
  d.out = yout; d.out_ld = CoutP;
  if (impl == 2 || (impl == 0 && dt == kBF16 && conv_umma_supported(d))) {
    if (dt != kBF16) return fail("cdan_op_conv2d: the tcgen05 path is bf16 only");
    UmmaPack* pk = nullptr;
    CDAN_TRY(umma_pack_create(pw.data(), hb.data(), CinP, Cout, CoutP, ks, &pk));
    int rc = conv_umma_launch(d, *pk, s);
    if (rc == 0) {
      cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) rc = fail(std::string("cdan_op_conv2d: tcgen05 kernel failed: ") + cudaGetErrorString(e));
    }
    umma_pack_destroy(pk);
    if (rc) return rc;
  } else {
    CDAN_TRY(conv_simt_launch(d, dt, s));
  }
  CDAN_TRY(nhwc_to_nchw_launch(dt, yout, CoutP, y, N, Cout, OH, OW, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cdan_op_cbam(int dtype, void* stream, const float* x, int N, int C, int H, int W, const float* w1, const float* b1,
                 const float* w2, const float* b2, const float* w7, const float* bn_host4, const float* mul, float* y) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !w7 || !bn_host4 || !y) return fail("cdan_op_cbam: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const DType dt = DType(dtype);
  Scratch sc;
  void *xin, *min = nullptr, *yout;
  float* scratch;
  const size_t elems = size_t(N) * H * W * C;
  CDAN_TRY(sc.alloc(&xin, elems * esize(dtype)));
  CDAN_TRY(sc.alloc(&yout, elems * esize(dtype)));
  CDAN_TRY(nchw_to_nhwc_launch(dt, x, xin, C, N, C, H, W, s));
  if (mul) {
    CDAN_TRY(sc.alloc(&min, elems * esize(dtype)));
    CDAN_TRY(nchw_to_nhwc_launch(dt, mul, min, C, N, C, H, W, s));
  }
  CDAN_TRY(sc.alloc((void**)&scratch, cbam_scratch_floats(N, C, H, W) * 4));
  CbamScratch cs;
  cbam_scratch_carve(scratch, N, C, H, W, &cs);
  CbamWeights wt;
  wt.w1 = w1; wt.b1 = b1; wt.w2 = w2; wt.b2 = b2; wt.w7 = w7;
  const double a = double(bn_host4[0]) / std::sqrt(double(bn_host4[3]) + 1e-5);
  wt.bn_a = float(a);
  wt.bn_b = float(double(bn_host4[1]) - double(bn_host4[2]) * a);
  CDAN_TRY(cbam_launch(dt, xin, C, min, C, yout, C, N, H, W, C, wt, cs, false, s));
  CDAN_TRY(nhwc_to_nchw_launch(dt, yout, C, y, N, C, H, W, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cdan_op_upsample_add(int dtype, void* stream, const float* a, const float* skip, int N, int C, int H, int W, int up,
                         float* y) {
  if (!a || !skip || !y) return fail("cdan_op_upsample_add: NULL argument");
  if (C % 8) return fail("cdan_op_upsample_add: C must be a multiple of 8");
  cudaStream_t s = (cudaStream_t)stream;
  const DType dt = DType(dtype);
  Scratch sc;
  const int OH = up ? 2 * H : H, OW = up ? 2 * W : W;
  void *ain, *sin, *yout;
  CDAN_TRY(sc.alloc(&ain, size_t(N) * H * W * C * esize(dtype)));
  CDAN_TRY(sc.alloc(&sin, size_t(N) * OH * OW * C * esize(dtype)));
  CDAN_TRY(sc.alloc(&yout, size_t(N) * OH * OW * C * esize(dtype)));
  CDAN_TRY(nchw_to_nhwc_launch(dt, a, ain, C, N, C, H, W, s));
  CDAN_TRY(nchw_to_nhwc_launch(dt, skip, sin, C, N, C, OH, OW, s));
  CDAN_TRY(up_add_launch(dt, ain, C, sin, C, yout, C, N, OH, OW, C, up, s));
  CDAN_TRY(nhwc_to_nchw_launch(dt, yout, C, y, N, C, OH, OW, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cdan_postprocess(void* stream, int op, float arg, const float* x, float* y, int N, int H, int W) {
  if (!x || !y) return fail("cdan_postprocess: NULL argument");
  if (op < 0 || op > 3) return fail("cdan_postprocess: unknown op");
  if ((op == 2 || op == 3) && x == y) return fail("cdan_postprocess: stencil ops cannot run in place");
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;
  float* scratch;
  CDAN_TRY(sc.alloc((void**)&scratch, postproc_scratch_floats(N, H, W) * 4));
  CDAN_TRY(postproc_launch(op, x, y, N, H, W, arg, scratch, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));  // scratch is freed on return
  return 0;
}

namespace {
// Source indices and 11-bit weights of cv::resize INTER_LINEAR for one axis (resize.cpp, resizeGeneric_ set-up): the
// coordinate is evaluated in double, its fraction in float; horizontally the fraction is forced to 0 where the two-tap
// window leaves the image, vertically only the two row indices are clipped.  Layout: [i0 | i1 | w0 | w1], n entries each.
std::vector<int> cv_linear_table(int src, int dst, bool vertical) {
  std::vector<int> t(size_t(4) * dst);
  const double scale = 1.0 / (double(dst) / double(src));
  for (int d = 0; d < dst; ++d) {
    float f = float((d + 0.5) * scale - 0.5);
    int s = int(floorf(f));
    f -= float(s);
    if (!vertical) {
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    t[d] = std::min(std::max(s, 0), src - 1);
    t[dst + d] = std::min(std::max(s + 1, 0), src - 1);
    t[2 * dst + d] = int(lrintf((1.f - f) * 2048.f));  // saturate_cast<short>(float): round half to even
    t[3 * dst + d] = int(lrintf(f * 2048.f));
  }
  return t;
}
}  // namespace

int cdan_resize_normalize_u8(void* stream, const unsigned char* src, int N, int Hs, int Ws, float* dst, int Hd, int Wd) {
  if (!src || !dst) return fail("cdan_resize_normalize_u8: NULL argument");
  if (N <= 0 || Hs <= 0 || Ws <= 0 || Hd <= 0 || Wd <= 0) return fail("cdan_resize_normalize_u8: empty input or output");
  cudaStream_t s = (cudaStream_t)stream;
  const std::vector<int> xt = cv_linear_table(Ws, Wd, false), yt = cv_linear_table(Hs, Hd, true);
  Scratch sc;
  int *d_xt, *d_yt;
  CDAN_TRY(sc.alloc((void**)&d_xt, xt.size() * sizeof(int)));
  CDAN_TRY(sc.alloc((void**)&d_yt, yt.size() * sizeof(int)));
  CDAN_CUDA_OK(cudaMemcpyAsync(d_xt, xt.data(), xt.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CDAN_CUDA_OK(cudaMemcpyAsync(d_yt, yt.data(), yt.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CDAN_TRY(resize_normalize_u8_launch(src, N, Hs, Ws, dst, Hd, Wd, d_xt, d_yt, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));  // the tables are freed on return
  return 0;
}

int cdan_quantize_u8(void* stream, const float* x, unsigned char* y, int N, int H, int W) {
  if (!x || !y) return fail("cdan_quantize_u8: NULL argument");
  if (N <= 0 || H <= 0 || W <= 0) return fail("cdan_quantize_u8: empty input");
  return quantize_u8_launch(x, y, N, H, W, (cudaStream_t)stream);
}

int cdan_psnr_ssim(void* stream, const float* pred, const float* target, int N, int C, int H, int W,
                   float* result_host2) {
  if (!pred || !target || !result_host2) return fail("cdan_psnr_ssim: NULL argument");
  if (H < 11 || W < 11) return fail("cdan_psnr_ssim: images must be at least 11x11 (SSIM window)");
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;
  float *scratch, *res;
  CDAN_TRY(sc.alloc((void**)&scratch, metrics_scratch_floats(N * C, H, W) * 4));
  CDAN_TRY(sc.alloc((void**)&res, 2 * 4));
  CDAN_TRY(psnr_ssim_launch(pred, target, N, C, H, W, scratch, res, s));
  CDAN_CUDA_OK(cudaMemcpyAsync(result_host2, res, 8, cudaMemcpyDeviceToHost, s));
  CDAN_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"
