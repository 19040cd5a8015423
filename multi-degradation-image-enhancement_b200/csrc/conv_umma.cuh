// tcgen05 (5th-gen tensor core) implicit-GEMM convolution, bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include "conv.cuh"

namespace cdan {

struct UmmaPack;  // device-resident, pre-swizzled bf16 weight image + fp32 bias

// Build from the CUDA-core packing (fp32 [taps][Cin][CoutP], BN already folded) — host pointers.
int umma_pack_create(const float* w_taps_cin_coutp, const float* bias_coutp, int Cin, int Cout, int CoutP, int ks,
                     UmmaPack** out);
void umma_pack_destroy(UmmaPack* p);

bool conv_umma_supported(const ConvDesc& d);

// Streaming, tap-folded variant (conv_stream.cu) used for the narrow-output layers; conv_umma_launch dispatches to it.
struct StreamPack;
int stream_pack_create(const float* w_taps_cin_coutp, const float* bias_coutp, int Cin, int Cout, int CoutP, int ks,
                       StreamPack** out);
void stream_pack_destroy(StreamPack* p);
bool conv_stream_supported(const ConvDesc& d, const StreamPack& pack);
int conv_stream_launch(const ConvDesc& d, const StreamPack& pack, cudaStream_t stream);
int conv_stream_kernel_count(const ConvDesc& d, const StreamPack& pack);
// Kernels conv_umma_launch issues for this convolution (1, or one per output-channel pass of the streaming kernel).
int conv_umma_kernel_count(const ConvDesc& d, const UmmaPack& pack);
int conv_umma_launch(const ConvDesc& d, const UmmaPack& pack, cudaStream_t stream);

}  // namespace cdan
