// Plan = packed weights + workspace + the layer schedule of one CDAN forward.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "band.cuh"
#include "common.cuh"
#include "conv.cuh"
#include "dense_fused.cuh"
#include "kernels.cuh"

namespace cdan {

enum ConvId : int {
  ENC1 = 0, ENC2, ENC3, ENC4,
  D1L0, D1L1, D1L2, D1L3, D1T,
  D2L0, D2L1, D2L2, D2L3, D2T,
  D3L0, D3L1, D3L2, D3L3, D3T,
  DEC1, DEC2, DEC3, DEC4,
  FDL0, FDL1, FDL2, FDL3, FDT,
  kNumConv
};

struct UmmaPack;  // tcgen05 weight image (conv_umma.cu)

struct ConvLayer {
  std::string name;
  int Cin = 0;    // physical input channels (incl. zero pad channels)
  int Cout = 0;
  int CoutP = 0;  // packed output-channel stride
  int ks = 3;
  int relu = 0;
  float* d_w = nullptr;      // fp32 [taps][Cin][CoutP], post-conv BN folded
  float* d_bias = nullptr;   // fp32 [CoutP]
  float* d_pre_s = nullptr;  // fp32 [Cin] pre-activation scale (dense layers) or null
  float* d_pre_t = nullptr;
  UmmaPack* umma = nullptr;
  // host copies of the packed tensors (kept for the layers a fused kernel re-packs)
  std::vector<float> h_w, h_bias, h_pre_s, h_pre_t;
};

struct CbamLayer {
  int C = 0;
  CbamWeights w;
};

struct Stage {
  const void* p = nullptr;
  int C = 0, ld = 0, h = 0, w = 0;
};

struct Buffers {
  // all NHWC in the plan's storage type; names follow the reference's forward (models/cdan.py:70-98,126-159)
  void *D1, *DN1, *D2, *DN2, *D3, *DN3, *E4, *B0, *T1, *A1, *C1, *T2, *U2, *C2, *T3, *U3, *C3, *T4, *FD;
  float* cbam_scratch;
  size_t total_bytes;
};

}  // namespace cdan

struct cdan_plan {
  int device = 0;
  cdan::DType dt = cdan::kF32;
  int conv_impl = 0;
  bool loaded = false;
  cdan::ConvLayer conv[cdan::kNumConv];
  cdan::CbamLayer cbam[4];
  bool fd_fused_on = true;                // option "fd_fused"
  cdan::FusedFdPack* fd_fused = nullptr;  // final dense block as one kernel (bf16 tensor-core plans)
  // spatial row tiling (band.cuh): transport to the neighbouring bands, counters of the most recent banded forward
  cdan::BandComm* band_comm = nullptr;
  cdan::BandStats band_stats;
  std::vector<void*> owned;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int N = 0, H = 0, W = 0;
  int launches = 0;
  cdan::Buffers buf{};
  std::map<std::string, cdan::Stage> stages;
  cudaStream_t own_stream = nullptr;
  // host-buffer entry point: copy streams, per-slot events and double-buffered device staging (sub-batch pipeline)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  void* host_stage = nullptr;   // [2 slots][fp32 x | fp32 y | u8 x | u8 y]
  size_t host_stage_bytes = 0;
  int host_chunk = 0;           // images per full pipeline step (option "host_chunk"; 0 = auto: 16 x 1080p of pixels)
  // optional per-launch CUDA-event timing ("profile" option): label -> accumulated ms / count
  int profile = 0;
  struct Span { std::string label; cudaEvent_t e0, e1; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
};
