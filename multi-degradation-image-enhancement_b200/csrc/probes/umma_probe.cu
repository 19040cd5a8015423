// Probe for the building blocks of the tcgen05 implicit-GEMM convolution (run on a B200 via gpurun):
//   T1  tcgen05.mma M=128, N in {16,64,256}, K-major SWIZZLE_128B operands written by threads, with the A
//       descriptor starting at an arbitrary ROW offset r0 inside a larger 1024B-aligned swizzled buffer
//       ("shifted view"). Tells us whether the 128B swizzle is a function of absolute smem address bits
//       (base_offset = 0 works for any r0) or needs the descriptor base_offset field.
//   T2  a miniature 3x3 convolution: TMA 4-D tile load of a halo'd NHWC tile (negative / OOB coordinates ->
//       hardware zero fill), bulk copy of host-pre-swizzled weights, 9 shifted-view MMA groups, TMEM epilogue.
// Every wait is bounded (ptx::mbar_wait traps on timeout), so a protocol bug cannot hang the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
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

// ------------------------------------------------------------------------------------------------ T1
// A_g: [RA rows][64] bf16 row-major, B_g: [N][64] bf16 row-major. D_g: [128][N] fp32 = A[r0:r0+128] * B^T
__global__ void __launch_bounds__(128) k_probe_shift(const bf16* A_g, int RA, const bf16* B_g, int N, float* D_g,
                                                     int r0, int base_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                        // RA*128 bytes (RA multiple of 8)
  uint8_t* sB = smem + ((RA * 128 + 1023) / 1024) * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < RA * 8; i += 128) {
    int row = i / 8, ch = i % 8;
    *(uint4*)(sA + ptx::sw128_offset(row, ch)) = *(const uint4*)(A_g + row * 64 + ch * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    int row = i / 8, ch = i % 8;
    *(uint4*)(sB + ptx::sw128_offset(row, ch)) = *(const uint4*)(B_g + row * 64 + ch * 8);
  }
  ptx::fence_proxy_async_smem();
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sA) + r0 * 128 + k * 32, 1024, base_off);
      uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(sB) + k * 32, 1024, 0);
      ptx::umma_bf16(tmem, da, db, idesc, k > 0);
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c, v);
    ptx::tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D_g[(warp * 32 + (tid % 32)) * N + c + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------ T2
// Input x: [1][H][W][64] bf16 NHWC. Weights pre-swizzled on host: 9 taps x [COUT rows][128 B].
// One CTA computes output rows h0..h0+TH-1, all W (W <= WP-2), COUT channels.
constexpr int T2_TH = 4, T2_WP = 32, T2_COUT = 32;
__global__ void __launch_bounds__(128) k_probe_conv(const __grid_constant__ CUtensorMap tmap, const uint8_t* Wsw_g,
                                                    float* D_g /*[128][COUT] flattened (hh*WP+ww)*/, int h0,
                                                    uint8_t* dumpA /* (TH+2)*WP*128 bytes raw smem image */) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = (T2_TH + 2) * T2_WP * 128;  // 24576
  constexpr int B_TAP = T2_COUT * 128;                 // 4096
  uint8_t* sA = smem;
  uint8_t* sPad = smem + A_BYTES;  // 1 KB of slack read by the last (garbage) rows
  uint8_t* sB = sPad + 1024;
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;

  for (int i = tid; i < 1024 / 4; i += 128) ((uint32_t*)sPad)[i] = 0;
  ptx::fence_proxy_async_smem();
  if (tid == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 32);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(&bar_full, A_BYTES + 9 * B_TAP);
    ptx::tma_load_4d(sA, &tmap, 0, -1, h0 - 1, 0, &bar_full);
    ptx::bulk_g2s(sB, Wsw_g, 9 * B_TAP, &bar_full);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after_sync();
    const uint32_t idesc = ptx::umma_idesc_bf16(128, T2_COUT);
    int first = 1;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 4; ++k) {
          uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sA) + (r * T2_WP + s) * 128 + k * 32, 1024, 0);
          uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(sB) + (r * 3 + s) * B_TAP + k * 32, 1024, 0);
          ptx::umma_bf16(tmem, da, db, idesc, first ? 0u : 1u);
          first = 0;
        }
    ptx::umma_commit(&bar_mma);
  }
  ptx::mbar_wait(&bar_mma, 0);
  ptx::tc_fence_after_sync();
  for (int c = 0; c < T2_COUT; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c, v);
    ptx::tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D_g[(warp * 32 + (tid % 32)) * T2_COUT + c + j] = __uint_as_float(v[j]);
  }
  // dump the raw A tile so the host can check TMA layout / zero fill independently of the MMA
  for (int i = tid; i < A_BYTES / 16; i += 128) ((uint4*)dumpA)[i] = ((const uint4*)sA)[i];
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}


// ------------------------------------------------------------------------------------------------ T3
// Issue-rate / throughput of back-to-back tcgen05.mma (M=128, K=16) for several N, one CTA per SM, operands fixed.
__global__ void __launch_bounds__(128) k_probe_rate(int N, int reps, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // 4 M-blocks x 128 rows x 128 B = 64 KB
  uint8_t* sB = smem + 65536;       // 256 rows x 128 B = 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (65536 + 32768) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
      const uint64_t hi = ptx::umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
      const uint32_t fl = uint32_t(ptx::umma_desc_sw128(0, 1024));
      const uint32_t a0 = fl | ((ptx::smem_u32(sA) & 0x3FFFF) >> 4), b0 = fl | ((ptx::smem_u32(sB) & 0x3FFFF) >> 4);
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        uint32_t a = a0 + uint32_t(r & 7) * 8;  // shifted views like the conv taps
        uint32_t d = tmem;
        for (int mb = 0; mb < 4; ++mb) {
          ptx::umma_bf16(d, hi | a, hi | b0, idesc, 1u);
          ptx::umma_bf16(d, hi | (a + 2), hi | (b0 + 2), idesc, 1u);
          ptx::umma_bf16(d, hi | (a + 4), hi | (b0 + 4), idesc, 1u);
          ptx::umma_bf16(d, hi | (a + 6), hi | (b0 + 6), idesc, 1u);
          a += 1024;
          d += (4 * N <= 512) ? N : 0;
        }
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      t1 = clock64();
      cycles_out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------ T4
// TMEM read / clear throughput: `nwarps` warps (warp w owns lane quarter w%4) stream over `cols` columns `reps` times
// with tcgen05.ld.32x32b.x16 (mode 0), .x16 followed by a tcgen05.st of zeros to the same columns (mode 1).
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
}
__global__ void __launch_bounds__(256) k_probe_tmem(int mode, int cols, int reps, long long* cycles_out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t((warp & 3) * 32) << 16);
  const int half = (blockDim.x > 128) ? (warp >> 2) : 0, nhalf = blockDim.x > 128 ? 2 : 1;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int c = half * 16; c < cols; c += 16 * nhalf) {
      uint32_t v[16];
      ptx::tmem_ld16(tmem + c, v);
      ptx::tmem_wait_ld();
      if (mode == 1) {
        tmem_st16_zero(tmem + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += __uint_as_float(v[j]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base_s, 512);
}
// mode 2: two loads in flight before the wait (x16 + x16); mode 3: one x32 load
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__global__ void __launch_bounds__(256) k_probe_tmem32(int cols, int reps, long long* cycles_out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t((warp & 3) * 32) << 16);
  const int half = (blockDim.x > 128) ? (warp >> 2) : 0, nhalf = blockDim.x > 128 ? 2 : 1;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int c = half * 32; c < cols; c += 32 * nhalf) {
      uint32_t v[32];
      tmem_ld32(tmem + c, v);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base_s, 512);
}

// ------------------------------------------------------------------------------------------------ host
static float bf(bf16 v) { return __bfloat162float(v); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  srand(1);

  // ---------------- T1
  {
    const int RA = 208;  // rows in the A buffer
    std::vector<bf16> A(RA * 64), B(256 * 64);
    for (auto& v : A) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : B) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    bf16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    const int smem_bytes = 1024 + 26 * 1024 + 32 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k_probe_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    struct Cfg {
      int N, r0, bo;
    };
    std::vector<Cfg> cfgs = {{64, 0, 0},  {16, 0, 0},  {256, 0, 0}, {64, 8, 0},  {64, 1, 0},  {64, 1, 1},
                             {64, 2, 0},  {64, 2, 2},  {64, 3, 0},  {64, 5, 0},  {64, 5, 5},  {64, 33, 0},
                             {64, 34, 0}, {64, 66, 0}, {64, 67, 0}, {64, 67, 3}, {16, 35, 0}, {256, 77, 0}};
    for (auto c : cfgs) {
      CK(cudaMemset(dD, 0, 128 * 256 * 4));
      k_probe_shift<<<1, 128, smem_bytes>>>(dA, RA, dB, c.N, dD, c.r0, c.bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("T1 N=%d r0=%d bo=%d : CUDA ERROR %s\n", c.N, c.r0, c.bo, cudaGetErrorString(e));
        return 3;
      }
      std::vector<float> D(128 * c.N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < c.N; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)bf(A[(c.r0 + m) * 64 + k]) * bf(B[n * 64 + k]);
          maxerr = fmax(maxerr, fabs(ref - D[m * c.N + n]));
        }
      printf("T1 N=%3d r0=%3d base_off=%d : max_abs_err=%.3e %s\n", c.N, c.r0, c.bo, maxerr,
             maxerr < 1e-3 ? "PASS" : "FAIL");
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
  }

  // ---------------- T2
  {
    const int H = 12, W = 30, C = 64;
    std::vector<bf16> X(H * W * C), Wt(T2_COUT * 9 * C);
    for (auto& v : X) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : Wt) v = __float2bfloat16((rand() % 2001 - 1000) / 4000.0f);  // [cout][tap][c]
    // host pre-swizzle: tap-major blocks of [COUT rows][128B]
    std::vector<uint8_t> Wsw(9 * T2_COUT * 128);
    for (int t = 0; t < 9; ++t)
      for (int n = 0; n < T2_COUT; ++n)
        for (int ch = 0; ch < 8; ++ch)
          memcpy(&Wsw[t * T2_COUT * 128 + ptx::sw128_offset(n, ch)], &Wt[(n * 9 + t) * C + ch * 8], 16);
    bf16* dX;
    uint8_t *dW, *dDump;
    float* dD;
    CK(cudaMalloc(&dX, X.size() * 2));
    CK(cudaMalloc(&dW, Wsw.size()));
    CK(cudaMalloc(&dD, 128 * T2_COUT * 4));
    CK(cudaMalloc(&dDump, (T2_TH + 2) * T2_WP * 128));
    CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, Wsw.data(), Wsw.size(), cudaMemcpyHostToDevice));

    PFN_encodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) {
      printf("no cuTensorMapEncodeTiled\n");
      return 4;
    }
    CUtensorMap tmap;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)T2_WP, (cuuint32_t)(T2_TH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dX, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return 5;
    }
    const int smem_bytes = 1024 + (T2_TH + 2) * T2_WP * 128 + 1024 + 9 * T2_COUT * 128 + 1024;
    CK(cudaFuncSetAttribute(k_probe_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    for (int h0 : {0, 4, 8}) {
      CK(cudaMemset(dD, 0, 128 * T2_COUT * 4));
      k_probe_conv<<<1, 128, smem_bytes>>>(tmap, dW, dD, h0, dDump);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("T2 h0=%d : CUDA ERROR %s\n", h0, cudaGetErrorString(e));
        return 6;
      }
      std::vector<float> D(128 * T2_COUT);
      std::vector<uint8_t> dump((T2_TH + 2) * T2_WP * 128);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(dump.data(), dDump, dump.size(), cudaMemcpyDeviceToHost));
      // (a) TMA tile check: pixel-row p = hh*WP + ww holds x[h0-1+hh][ww-1][:] (zero outside the image)
      int bad = 0;
      for (int hh = 0; hh < T2_TH + 2; ++hh)
        for (int ww = 0; ww < T2_WP; ++ww)
          for (int c = 0; c < C; ++c) {
            int h = h0 - 1 + hh, w = ww - 1;
            float ref = (h >= 0 && h < H && w >= 0 && w < W) ? bf(X[(h * W + w) * C + c]) : 0.f;
            int p = hh * T2_WP + ww;
            bf16 got;
            memcpy(&got, &dump[ptx::sw128_offset(p, c / 8) + (c % 8) * 2], 2);
            if (bf(got) != ref) ++bad;
          }
      // (b) conv check
      double maxerr = 0;
      for (int hh = 0; hh < T2_TH; ++hh)
        for (int ww = 0; ww < W; ++ww)
          for (int n = 0; n < T2_COUT; ++n) {
            double ref = 0;
            for (int rr = 0; rr < 3; ++rr)
              for (int ss = 0; ss < 3; ++ss) {
                int h = h0 + hh + rr - 1, w = ww + ss - 1;
                if (h < 0 || h >= H || w < 0 || w >= W) continue;
                for (int c = 0; c < C; ++c)
                  ref += (double)bf(X[(h * W + w) * C + c]) * bf(Wt[(n * 9 + rr * 3 + ss) * C + c]);
              }
            maxerr = fmax(maxerr, fabs(ref - D[(hh * T2_WP + ww) * T2_COUT + n]));
          }
      printf("T2 h0=%d : tma_tile_mismatches=%d conv_max_abs_err=%.3e %s\n", h0, bad, maxerr,
             (bad == 0 && maxerr < 2e-3) ? "PASS" : "FAIL");
    }
  }

  // ---------------- T3
  {
    long long* dC;
    CK(cudaMalloc(&dC, 148 * 8));
    const int smem_bytes = 1024 + 65536 + 32768 + 1024;
    CK(cudaFuncSetAttribute(k_probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    for (int N : {16, 32, 48, 64, 96, 128, 144, 192, 256}) {
      const int reps = 500;  // 500 * 16 = 8000 MMAs
      for (int it = 0; it < 2; ++it) {
        k_probe_rate<<<148, 128, smem_bytes>>>(N, reps, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T3 N=%d CUDA ERROR %s\n", N, cudaGetErrorString(e)); return 7; }
      }
      std::vector<long long> c(148);
      CK(cudaMemcpy(c.data(), dC, 148 * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto v : c) avg += v; avg /= 148;
      printf("T3 N=%3d : %.1f cycles per MMA (M=128,K=16), ideal %.1f -> %.1f%% of tensor peak\n", N, avg / (reps * 16.0),
             N / 2.0, 100.0 * (N / 2.0) / (avg / (reps * 16.0)));
    }
    cudaFree(dC);
  }

  // ---------------- T4
  {
    long long* dC; float* dS;
    CK(cudaMalloc(&dC, 148 * 8));
    CK(cudaMalloc(&dS, 1024 * 4));
    for (int threads : {128, 256})
      for (int mode : {0, 1, 3}) {
        const int reps = 200, cols = 512;
        for (int it = 0; it < 2; ++it) {
          if (mode == 3) k_probe_tmem32<<<148, threads>>>(cols, reps, dC, dS);
          else k_probe_tmem<<<148, threads>>>(mode, cols, reps, dC, dS);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("T4 CUDA ERROR %s\n", cudaGetErrorString(e)); return 8; }
        }
        std::vector<long long> c(148);
        CK(cudaMemcpy(c.data(), dC, 148 * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : c) avg += v; avg /= 148;
        const double bytes = double(reps) * 128 * cols * 4;
        printf("T4 threads=%d mode=%d (%s): %.1f B/cycle/SM TMEM, %.1f cycles per 128x16 block\n", threads, mode,
               mode == 0 ? "ld.x16" : mode == 1 ? "ld.x16+st.x16 zero" : "ld.x32", bytes / avg, avg / (reps * cols / 16.0));
      }
    cudaFree(dC); cudaFree(dS);
  }
  printf("probe done\n");
  return 0;
}
