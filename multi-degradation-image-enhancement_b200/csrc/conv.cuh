// Convolution launch descriptor shared by the CUDA-core (fp32-exact) and tcgen05 (bf16) implementations.
#pragma once
#include "common.cuh"

namespace cdan {

// One stride-1 "same" convolution (3x3 pad 1, or 1x1) over an NHWC activation, optionally
//   * with the dense-block pre-activation prologue a = relu(pre_scale[c]*x[c] + pre_shift[c]) applied to the
//     INPUT before zero padding (reference: models/cdan.py:41-53, BN -> ReLU -> Conv),
//   * with bias (+ folded BatchNorm) and ReLU on the output (models/cdan.py:15-19, 127-129),
//   * with a fused 2x2/stride-2 max-pool of the output (models/cdan.py:75,82,89),
//   * reading the network input as planar fp32 NCHW, or writing the network output as planar fp32 NCHW with
//     a fused sigmoid (models/cdan.py:157).
struct ConvDesc {
  int N = 0, H = 0, W = 0;  // input (= un-pooled output) extent
  int Cin = 0;              // physical input channels consumed (incl. zero-weight pad channels)
  int Cout = 0;
  int ks = 3;               // 3 or 1
  const void* in = nullptr; // NHWC, storage type T
  int in_ld = 0;
  const void* in2 = nullptr;  // hybrid concat buffer (Chead > 0): base of the group planes; `in`/`in_ld` describe the NHWC head
  int Chead = 0;              // channels of the NHWC head (multiple of 64); 0 = all channels are in group planes
  size_t in_gstride = 0;    // bf16 tensor-core path only: if non-zero the input is GROUP-PLANAR — one dense plane
                            // [N][H][W][16] per 16-channel group, planes `in_gstride` elements apart, in_ld = 16
  const float* in_nchw = nullptr;  // if set: planar fp32 input [N][Cin][H][W] (first layer), `in` unused
  const float* pre_scale = nullptr;
  const float* pre_shift = nullptr;
  const float* w = nullptr;  // fp32 [taps][Cin][CoutP], BN folded
  int CoutP = 0;
  const float* bias = nullptr;  // [CoutP]
  int relu = 0;
  int pool = 0;
  void* out = nullptr;  // NHWC, storage type T (already offset to the channel slice)
  int out_ld = 0;
  float* out_nchw = nullptr;  // if set: planar fp32 output [N][Cout][H][W]
  int sigmoid = 0;
};

int conv_simt_launch(const ConvDesc& d, DType dt, cudaStream_t stream);

}  // namespace cdan
