// Shared declarations for the CDAN sm_100a kernels: storage types, NHWC views, error plumbing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace cdan {

using bf16 = __nv_bfloat16;

enum DType : int { kF32 = 0, kBF16 = 1 };

// Thread-local error message behind cdan_last_error().
void set_error(const std::string& msg);
int fail(const std::string& msg);  // records msg, returns -1

#define CDAN_CUDA_OK(expr)                                                                              \
  do {                                                                                                  \
    cudaError_t err__ = (expr);                                                                         \
    if (err__ != cudaSuccess)                                                                           \
      return ::cdan::fail(std::string(#expr) + " failed: " + cudaGetErrorString(err__) + " (" + __FILE__ + \
                          ":" + std::to_string(__LINE__) + ")");                                       \
  } while (0)

#define CDAN_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != 0) return rc__;   \
  } while (0)

// NHWC activation view with an explicit channel stride, so a tensor can be a channel slice of a wider
// (dense-block concat) buffer: element (n,h,w,c) lives at p[((n*H + h)*W + w)*ld + c].
struct View {
  void* p = nullptr;
  int C = 0;   // channels visible through the view
  int ld = 0;  // channel stride of the underlying buffer (elements)
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels as fp32 (vector width used by all HBM-bound kernels: 16 B of bf16, 32 B of fp32).
struct F8 {
  float v[8];
};
template <typename T>
__device__ __forceinline__ F8 load8(const T* p);
template <>
__device__ __forceinline__ F8 load8<float>(const float* p) {
  F8 r;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <>
__device__ __forceinline__ F8 load8<bf16>(const bf16* p) {
  F8 r;
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const F8& r);
template <>
__device__ __forceinline__ void store8<float>(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const F8& r) {
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace cdan
