// Spatial row tiling of very large images (SURVEY 8(e) "spatial rows", BASELINE config C5): one band of image rows per GPU.
//
// The CDAN forward is not separable by rows (reference models/cdan.py:70-159, models/cbam.py:37-82): every 3x3 convolution,
// bilinear x2 and SpatialGate 7x7 looks one to three rows across a band boundary and every ChannelGate pools over the whole
// image.  A band therefore carries `halo` extra rows above and below (8 | halo, halo/8 >= 3) at full resolution, halo >> k
// at resolution level k, and runs the UNCHANGED kernels on the extended band as if it were a whole image.  Rows near the
// artificial border are wrong ("dirt") and every layer widens the dirty zone (3x3: +1 row, 7x7: +3, bilinear x2: 2d+1,
// 2x2 max-pool: ceil(d/2)); plan.cu tracks that depth per tensor and, only when the next layer would push it into the owned
// rows, refreshes the tensor's halo rows from the neighbours (7 exchanges per forward at halo 24 instead of one before each
// of the 31 cross-row operators).  ChannelGate statistics are pooled over the owned rows and all-reduced (SUM and MAX of
// [N, C] fp32); the mean divides by the FULL image's pixel count (models/cbam.py:41,44).
//
// Two transports behind one interface: NCCL (one process per GPU; ncclSend/ncclRecv in a group over NVLink, ncclAllReduce),
// loaded at run time so that single-GPU users need no NCCL, and an in-process transport (one host thread per band, device
// copies ordered by events) that makes the whole schedule testable on ONE GPU.
#pragma once
#include "common.cuh"

namespace cdan {

struct HaloMsg {       // the halo rows of one tensor
  char* base = nullptr;     // the tensor inside this band's workspace
  size_t img_stride = 0;    // bytes between images
  int nimg = 0;
  size_t bytes = 0;         // halo depth (rows) x row bytes
  size_t top_send = 0, top_recv = 0, bot_send = 0, bot_recv = 0;  // byte offsets inside an image
};

struct BandComm {
  int rank = 0, nranks = 1;
  virtual ~BandComm() {}
  // rows [top_send, +bytes) go to rank-1 and arrive in its [bot_recv, +bytes); rows [bot_send, +bytes) go to rank+1
  virtual int exchange(const HaloMsg& m, cudaStream_t s) = 0;
  // in place over all bands: sum[i] = SUM, mx[i] = MAX, i < count
  virtual int allreduce_sum_max(float* sum, float* mx, int count, cudaStream_t s) = 0;
};

struct BandStats {
  int exchanges = 0, allreduces = 0;
  long long halo_bytes_received = 0;
};

// Band split: contiguous bands, boundaries at multiples of 8 rows; rows_out = {owned begin, owned end, extended begin, extended end}
int band_rows(int H, int nbands, int rank, int halo, int rows_out[4]);

}  // namespace cdan

// in-process transport shared by the plans (bands) of one process
struct cdan_band_group;
