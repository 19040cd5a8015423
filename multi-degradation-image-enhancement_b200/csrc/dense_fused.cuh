// Fused final dense block (dense_fused.cu): up-sample + skip add, four 3x3 growth layers, 1x1 transition and sigmoid of
// the decoder tail (reference models/cdan.py:153-157, DenseBlock :22-53) as ONE tcgen05 kernel.
#pragma once
#include "common.cuh"

namespace cdan {

struct FusedFdPack;  // device-resident parameter blob (SWIZZLE_32B weight blocks, activation tables, biases)

// One layer of the block as packed for the CUDA-core kernel (HOST pointers): w is fp32 [taps][Cin][CoutP] with
// Cin = 16 * (l + 1) physical input channels (3 real + 13 pad, then 16 per earlier layer), pre_s / pre_t the
// pre-activation BatchNorm scale / shift per physical input channel.  layers[0..3] = growth layers, [4] = transition.
struct FusedFdLayer {
  int Cin = 0, CoutP = 0;
  const float* w = nullptr;
  const float* bias = nullptr;
  const float* pre_s = nullptr;
  const float* pre_t = nullptr;
};

int fused_fd_pack_create(const FusedFdLayer layers[5], FusedFdPack** out);
void fused_fd_pack_destroy(FusedFdPack* p);
// t4: bf16 NHWC [N][H/2][W/2][t4_ld] = relu(bn4(convT4(.))) (channels 0..2); x, y: fp32 NCHW [N][3][H][W].
int fused_fd_launch(const FusedFdPack& pack, const void* t4, int t4_ld, const float* x, float* y, int N, int H, int W,
                    cudaStream_t stream);

}  // namespace cdan
