// Streaming-load probe (run on a B200 via gpurun): how fast can one CTA per SM pull [128 px x 64 ch] bf16 rows of a
// channel-strided NHWC tensor from HBM into shared memory?
//   mode 0: TMA 4-D tile loads (box {64 ch, 128 px, ROWS, 1}, SWIZZLE_128B) into an S-stage ring, one issuing thread
//   mode 1: LDG.128 by 256 threads (U loads in flight per thread) + swizzled STS.128
//   mode 2: like mode 0 but a TMA L2 prefetch of the row PF stages ahead is issued as well
// Reports useful GB/s (bytes landed in smem / time).  All waits bounded.
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
  int N, H, W, ld, strips, rows_per_op, S, nchunks, pf;
  const bf16* in;
};

__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(320, 1) k_tma(const __grid_constant__ CUtensorMap tmap, const PP P, int mode,
                                                unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int stage_bytes = 16384 * P.rows_per_op;
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
        for (int j = 0; j < P.H; j += P.rows_per_op) {
          for (int c = 0; c < P.nchunks; ++c) {
            if (mode == 2 && j + P.pf * P.rows_per_op < P.H) tma_prefetch_4d(&tmap, c * 64, w0 - 1, j + P.pf * P.rows_per_op, n);
            ptx::mbar_wait_relaxed(&empty[s], ph ^ 1, 32);
            ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
            ptx::tma_load_4d(smem + size_t(s) * stage_bytes, &tmap, c * 64, w0 - 1, j, n, &full[s]);
            if (++s == P.S) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 2) {
    int s = 0, ph = 0;
    unsigned long long acc = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      for (int j = 0; j < P.H; j += P.rows_per_op) {
        for (int c = 0; c < P.nchunks; ++c) {
          ptx::mbar_wait_relaxed(&full[s], ph, 32);
          acc += *(const unsigned*)(smem + size_t(s) * stage_bytes + (tid & 255) * 16);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&empty[s]);
          if (++s == P.S) { s = 0; ph ^= 1; }
        }
      }
    }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
}

template <int U>
__global__ void __launch_bounds__(256, 1) k_ldg(const PP P, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x;
  const int u = tid & 7, qb = tid >> 3;  // 16-byte unit within the pixel's 128 B, pixel qb + 32*i
  const int nitems = P.N * P.strips;
  unsigned long long acc = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int n = item / P.strips, w0 = (item % P.strips) * 126;
    for (int j = 0; j < P.H; j += U / 4 > 0 ? U / 4 : 1) {
      for (int c = 0; c < P.nchunks; ++c) {
        uint4 v[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int row = j + i / 4, q = qb + 32 * (i % 4);
          const int col = w0 - 1 + q;
          const bool ok = col >= 0 && col < P.W && row < P.H;
          const bf16* src = P.in + ((size_t(n) * P.H + (ok ? row : 0)) * P.W + (ok ? col : 0)) * P.ld + c * 64 + u * 8;
          v[i] = ok ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int q = qb + 32 * (i % 4);
          *reinterpret_cast<uint4*>(smem + (i / 4) * 16384 + ptx::sw128_offset(q, u)) = v[i];
        }
        acc += v[0].x;
      }
    }
  }
  if (acc == 0x1234567ull) sink[0] = acc;
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

  struct Case { int N, H, W, ld, C; const char* name; };
  Case cases[] = {{8, 1080, 1920, 80, 64, "final_dense L3 (64 of 80 ch, 8 img)"},
                  {8, 1080, 1920, 64, 64, "dense NHWC ld=64 (contiguous)"},
                  {32, 540, 960, 128, 64, "dense1 L0 (64 of 128 ch)"},
                  {32, 540, 960, 128, 128, "dense1 T (128 of 128 ch, 2 chunks)"}};
  for (auto cs : cases) {
    const size_t elems = size_t(cs.N) * cs.H * cs.W * cs.ld;
    bf16* dIn;
    CK(cudaMalloc(&dIn, elems * 2));
    CK(cudaMemset(dIn, 0, elems * 2));
    PP P{};
    P.N = cs.N; P.H = cs.H; P.W = cs.W; P.ld = cs.ld; P.in = dIn;
    P.strips = (cs.W + 125) / 126;
    P.nchunks = cs.C / 64;
    const double useful = double(cs.N) * P.strips * cs.H * P.nchunks * 16384.0;
    printf("== %s: %.2f GB useful per pass\n", cs.name, useful / 1e9);
    for (int rows : {1, 2, 4}) {
      if (cs.H % rows) continue;
      CUtensorMap tmap;
      cuuint64_t gdim[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)cs.N};
      cuuint64_t gstr[3] = {(cuuint64_t)cs.ld * 2, (cuuint64_t)cs.W * cs.ld * 2, (cuuint64_t)cs.H * cs.W * cs.ld * 2};
      cuuint32_t box[4] = {64, 128, (cuuint32_t)rows, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dIn, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 5; }
      for (int S : {4, 12}) {
        if (S * rows * 16384 > 200 * 1024) continue;
        for (int mode : {0, 2}) {
          P.rows_per_op = rows; P.S = S; P.pf = 24 / rows;
          const int smem = S * rows * 16384 + 2048;
          CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          float best = 1e9;
          for (int it = 0; it < 3; ++it) {
            CK(cudaEventRecord(e0));
            k_tma<<<148, 320, smem>>>(tmap, P, mode, dSink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = fminf(best, ms);
          }
          printf("  TMA rows/op=%d stages=%2d %s: %.3f ms -> %.0f GB/s useful\n", rows, S, mode == 2 ? "+L2 prefetch" : "            ",
                 best, useful / best / 1e6);
        }
      }
    }
    {
      const int smem = 4 * 16384 + 2048;
      auto run = [&](auto kern, int U) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
          CK(cudaEventRecord(e0));
          kern<<<148, 256, smem>>>(P, dSink);
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          best = fminf(best, ms);
        }
        printf("  LDG.128 x%2d in flight per thread (256 thr): %.3f ms -> %.0f GB/s useful\n", U, best, useful / best / 1e6);
      };
      run(k_ldg<4>, 4);
      run(k_ldg<8>, 8);
      run(k_ldg<16>, 16);
    }
    CK(cudaFree(dIn));
  }
  printf("stream probe done\n");
  return 0;
}
