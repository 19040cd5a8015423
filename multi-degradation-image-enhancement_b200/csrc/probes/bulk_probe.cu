// Bulk-copy streaming probe (run on a B200 via gpurun): how fast can one CTA per SM pull strip rows of a GROUP-PLANAR
// activation layout (G planes of [N][H][W][16] bf16, a strip row of one group = 128 px x 32 B = 4 KB contiguous) into
// shared memory with cp.async.bulk (one request per row and group), compared with 4-D TMA tile loads of the same bytes
// from the NHWC layout (one 32..128-byte line per pixel)?  All waits bounded.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../ptx_sm100.cuh"

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

using bf16 = __nv_bfloat16;

struct PP {
  int N, H, W, strips, rows, S, G;
  const bf16* in;       // planar: plane g at in + g * plane_elems
  size_t plane_elems;
};

// mode 0: cp.async.bulk per (row, group); mode 1: 4-D TMA tile {16 ch, 128 px, rows, 1} per group from the planes
__global__ void __launch_bounds__(320, 1) k_bulk(const __grid_constant__ CUtensorMap tmap, const PP P, int mode,
                                                 unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int stage_bytes = 4096 * P.rows * P.G;
  if (tid == 0) {
    for (int i = 0; i < P.S; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 8); }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmap);
  }
  __syncthreads();
  const int nitems = P.N * P.strips;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int n = item / P.strips, w0 = (item % P.strips) * 126;
        for (int j = 0; j < P.H; j += P.rows) {
          ptx::mbar_wait_relaxed(&empty[s], ph ^ 1, 32);
          ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
          uint8_t* dst = smem + size_t(s) * stage_bytes;
          for (int g = 0; g < P.G; ++g) {
            if (mode == 0) {
              for (int r = 0; r < P.rows; ++r) {
                const bf16* src = P.in + size_t(g) * P.plane_elems + ((size_t(n) * P.H + j + r) * P.W + (w0 - 1)) * 16;
                ptx::bulk_g2s(dst + (g * P.rows + r) * 4096, src, 4096, &full[s]);
              }
            } else {
              ptx::tma_load_4d(dst + g * P.rows * 4096, &tmap, 0, w0 - 1, j, n + g * P.N, &full[s]);
            }
          }
          if (++s == P.S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 2) {
    int s = 0, ph = 0;
    unsigned long long acc = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      for (int j = 0; j < P.H; j += P.rows) {
        ptx::mbar_wait_relaxed(&full[s], ph, 32);
        acc += *(const unsigned*)(smem + size_t(s) * stage_bytes + (tid & 255) * 16);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        if (++s == P.S) { s = 0; ph ^= 1; }
      }
    }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  PFN_encodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  unsigned long long* dSink;
  CK(cudaMalloc(&dSink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int N = 8, H = 1080, W = 1920;
  const size_t plane = size_t(N) * H * W * 16;
  const int GMAX = 5;
  bf16* dBase;
  CK(cudaMalloc(&dBase, (plane * GMAX + 8192) * 2));
  CK(cudaMemset(dBase, 0, (plane * GMAX + 8192) * 2));
  bf16* dIn = dBase + 4096;  // guard in front: strip 0 starts one pixel before the row
  for (int G : {1, 2, 4, 5}) {
    PP P{};
    P.N = N; P.H = H; P.W = W; P.strips = (W + 125) / 126; P.G = G; P.in = dIn; P.plane_elems = plane;
    const double useful = double(N) * P.strips * H * G * 4096.0;
    printf("== group-planar, %d group(s) (%d B/px): %.2f GB useful per pass (x4 for 32 images)\n", G, 32 * G, useful / 1e9);
    for (int rows : {1, 2, 4}) {
      CUtensorMap tmap;
      cuuint64_t gdim[4] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * GMAX};
      cuuint64_t gstr[3] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32};
      cuuint32_t box[4] = {16, 128, (cuuint32_t)rows, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dIn, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 5; }
      for (int S : {4, 8, 12}) {
        const int stage = 4096 * rows * G;
        if (S * stage > 200 * 1024) continue;
        for (int mode : {0, 1}) {
          P.rows = rows; P.S = S;
          const int smem = S * stage + 2048;
          CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          float best = 1e9;
          for (int it = 0; it < 3; ++it) {
            CK(cudaEventRecord(e0));
            k_bulk<<<148, 320, smem>>>(tmap, P, mode, dSink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = fminf(best, ms);
          }
          printf("  rows/stage=%d stages=%2d (%3d KB in flight) %s: %.3f ms -> %.0f GB/s useful\n", rows, S, S * stage / 1024,
                 mode == 0 ? "cp.async.bulk 4 KB" : "TMA tile 32-B lines", best, useful / best / 1e6);
        }
      }
    }
  }
  printf("bulk probe done\n");
  return 0;
}
