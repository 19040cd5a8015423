// Launchers of the HBM-bound kernels (CBAM, decoder glue, layout conversion, post-processing, metrics).
#pragma once
#include "common.cuh"

namespace cdan {

// ---- decoder glue (models/cdan.py:130,137-138,145-146,153-154)
// out = (up ? bilinear_x2(a) : a) + skip ; a is [N,OH/2,OW/2,C] when up else [N,OH,OW,C]; all NHWC.
// If psum/pmax are given the kernel also writes the ChannelGate pooling partials of `out` (per image, per pair of
// output rows: [N][cbam_pool_blocks(OH)][C]), saving CBAM's first pass over the tensor.
int up_add_launch(DType dt, const void* a, int a_ld, const void* skip, int skip_ld, void* out, int out_ld, int N,
                  int OH, int OW, int C, int up, cudaStream_t s, float* psum = nullptr, float* pmax = nullptr);
// Final stage: out[n,h,w,0:3] = bilinear_x2(a)[.,0:3] + x_nchw ; out[n,h,w,3:pad_to] = 0.
int up_add_input_launch(DType dt, const void* a, int a_ld, const float* x_nchw, void* out, int out_ld, int pad_to,
                        int N, int OH, int OW, cudaStream_t s);

// ---- layout conversion (tests, stage taps)
int nchw_to_nhwc_launch(DType dt, const float* src, void* dst, int dst_ld, int N, int C, int H, int W, cudaStream_t s);
int nhwc_to_nchw_launch(DType dt, const void* src, int src_ld, float* dst, int N, int C, int H, int W, cudaStream_t s);

// ---- CBAM (models/cbam.py:37-60, 68-82, 91-95)
struct CbamWeights {
  const float* w1 = nullptr;  // [C/16][C]
  const float* b1 = nullptr;  // [C/16]
  const float* w2 = nullptr;  // [C][C/16]
  const float* b2 = nullptr;  // [C]
  const float* w7 = nullptr;  // [2][7][7]  (channel 0 = max map, 1 = mean map)
  float bn_a = 1.f, bn_b = 0.f;  // folded eval BatchNorm2d(1): s = a*conv + b
};
struct CbamScratch {
  float* psum = nullptr;   // [N][nblk][C]  per-(image, row pair) channel sums
  float* pmax = nullptr;   // [N][nblk][C]  ... and maxima
  float* gate = nullptr;   // [N][C]
  float* comp = nullptr;   // [N][H][W][2]
  float* sgate = nullptr;  // [N][H][W]
  float* pooled = nullptr; // [2][N][C]  band mode: channel sums and maxima of the owned rows, all-reduced in place
  int nblk = 0;
};
// Row-tiled forward (band.cuh): the ChannelGate pools over the rows this band OWNS and all-reduces the statistics.
struct BandComm;
struct CbamBand {
  int row0 = 0, rows = 0;   // owned rows of the (extended) tensor
  int HW_full = 0;          // pixels of the whole image at this resolution (divisor of the mean)
  BandComm* comm = nullptr;
};
int cbam_pool_blocks(int H);  // partial-reduction blocks per image: one per pair of image rows
size_t cbam_scratch_floats(int N, int C, int H, int W);
void cbam_scratch_carve(float* base, int N, int C, int H, int W, CbamScratch* sc);
// out = SpatialGate(ChannelGate(x)) [* mul]  (mul = dense-block output of the decoder, models/cdan.py:133,141,149).
// pooled = true: sc.psum / sc.pmax were already written by the kernel that produced x (up_add_launch with pool outputs).
int cbam_launch(DType dt, const void* x, int x_ld, const void* mul, int mul_ld, void* out, int out_ld, int N, int H,
                int W, int C, const CbamWeights& wt, const CbamScratch& sc, bool pooled, cudaStream_t s,
                const CbamBand* band = nullptr);

// ---- post-processing on planar fp32 NCHW images (utils/post_processing.py)
enum PostOp : int { kContrast = 0, kColor = 1, kSharpen = 2, kDenoise = 3 };
// scratch: >= N*3*(1 + blocks) floats + 1 int, see postproc_scratch_floats
size_t postproc_scratch_floats(int N, int H, int W);
int postproc_launch(int op, const float* x, float* y, int N, int H, int W, float arg, float* scratch, cudaStream_t s);

// ---- metrics on planar fp32 NCHW (torchmetrics defaults restated; parity unpinned)
size_t metrics_scratch_floats(int N, int H, int W);
// result[0] = PSNR, result[1] = SSIM (device pointer, 2 floats)
int resize_normalize_u8_launch(const uint8_t* src, int N, int Hs, int Ws, float* dst, int Hd, int Wd, const int* xt,
                               const int* yt, cudaStream_t s);
int normalize_u8_launch(const uint8_t* src, float* dst, int N, int H, int W, cudaStream_t s);
int quantize_u8_launch(const float* x, uint8_t* y, int N, int H, int W, cudaStream_t s);
int psnr_ssim_launch(const float* pred, const float* target, int N, int C, int H, int W, float* scratch, float* result,
                     cudaStream_t s);

}  // namespace cdan
