// Probe: TMA tile load of a planar fp32 image row [3 ch x BW px] without swizzle, for several inner box widths.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx_sm100.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(2);} } while (0)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int bw, int w0, int j, int n, float* out, int nch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, bw * nch * 4);
    ptx::tma_load_4d(smem, &tmap, w0, j, 0, n, &bar);
  }
  { int polls = 0; while (!ptx::mbar_try_wait(&bar, 0)) { if (++polls > 2000000) { if (threadIdx.x == 0) out[4000] = 12345.f; return; } } }
  for (int i = threadIdx.x; i < bw * nch; i += blockDim.x) out[i] = ((float*)smem)[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;  // 0: f32 none; 1: f32 sw128 (bw<=32); 2: f32 none, L2 promo none; 3: int32 none
  PFN_encodeTiled enc; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  for (int W : {64, 32, 24}) {
  const int N = 2, H = 16;
  std::vector<float> h(N * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = float(i);
  float *d, *o; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, 4096 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  for (int bw : {32, 64, 128, 132, 256}) {
    if (variant == 1 && bw > 32) continue;
    CUtensorMap tmap;
    cuuint64_t gdim[4] = {(cuuint64_t)W, H, 3, N}; cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, 1, 3, 1}; cuuint32_t estr[4] = {1, 1, 1, 1};
    if (variant == 4) { gdim[0] *= 4; box[0] *= 4; if (box[0] > 256) continue; }
    const int nch = variant >= 5 ? 1 : 3;
    box[2] = nch;
    CUresult r = enc(&tmap, variant == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (variant == 3 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, variant == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, variant == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("bw=%d encode failed %d\n", bw, (int)r); continue; }
    CK(cudaMemset(o, 0xff, 4096 * 4));
    k<<<1, 128, 16384>>>(tmap, bw, variant == 6 ? 0 : -1, 5, 1, o, nch);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bw=%d kernel failed: %s\n", bw, cudaGetErrorString(e)); return 3; }
    std::vector<float> g(bw * 3); CK(cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < nch; ++c) for (int p = 0; p < bw; ++p) { int col = (variant == 6 ? 0 : -1) + p; float ref = (col >= 0 && col < W) ? h[((1 * 3 + c) * H + 5) * W + col] : 0.f; if (g[c * bw + p] != ref) ++bad; }
    float flag; CK(cudaMemcpy(&flag, o + 4000, 4, cudaMemcpyDeviceToHost));
    printf("W=%d bw=%d %s, mismatches=%d\n", W, bw, flag == 12345.f ? "TIMEOUT" : "ok", bad);
  }
  CK(cudaFree(d)); CK(cudaFree(o));
  }
  return 0;
}
