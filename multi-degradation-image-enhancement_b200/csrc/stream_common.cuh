// Device helpers shared by the streaming tcgen05 kernels (conv_stream.cu, dense_fused.cu): ring counters and per-slot
// mbarrier phase bookkeeping.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cdan {
namespace {

__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bflo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfhi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Ring counter: index modulo `n` plus the number of wrap-arounds (mbarrier phase bookkeeping without integer division —
// every role is a single warp running a dependent instruction stream, so per-row instruction count IS the row time).
struct Ring {
  int i = 0, w = 0;
  __device__ __forceinline__ void step(int n) { if (++i == n) { i = 0; ++w; } }
  __device__ __forceinline__ void add(int k, int n) { i += k; while (i >= n) { i -= n; ++w; } }  // small k
  __device__ __forceinline__ void jump(int k, int n) { i += k; const int d = i / n; w += d; i -= d * n; }  // once per item
};

// Accumulator-ring bookkeeping.  A row's ring slot is a function of its ABSOLUTE image row (slot = (row + 2*PAD) mod R),
// not of a running counter: which slot a row lands in decides the association of its three vertical-tap partial sums
// (slots 0 and 1 are completed through the shadow slots), so tying it to the image row makes results independent of the
// batch size and of how the image is cut into segments.  Segments therefore start at arbitrary ring positions, and the
// mbarrier phase of every slot is tracked individually: bit s of `par` = parity of the next completion to wait for.
struct SlotPhases {
  uint32_t par = 0, used = 0;
  // consumer side of a barrier that is armed once per use of the slot
  __device__ __forceinline__ void wait(uint64_t* bars, int s) {
    ptx::mbar_wait(&bars[s], (par >> s) & 1u);
    par ^= 1u << s;
  }
  __device__ __forceinline__ void skip(int s) { par ^= 1u << s; }  // a use observed by someone else
  // same with the barrier array given as a 32-bit shared address (MMA issuers)
  __device__ __forceinline__ void claim_a(uint32_t free_bars, int s) {
    if ((used >> s) & 1u) {
      ptx::mbar_wait_a(free_bars + 8u * uint32_t(s), (par >> s) & 1u);
      par ^= 1u << s;
    }
    used |= 1u << s;
  }
  // producer side, a use claimed by ANOTHER producer thread: same bookkeeping as claim() without the wait
  __device__ __forceinline__ void note(int s) {
    if ((used >> s) & 1u) par ^= 1u << s;
    used |= 1u << s;
  }
  // producer side: before the first write of a new use, wait until the previous use (if any) was drained
  __device__ __forceinline__ void claim(uint64_t* free_bars, int s) {
    if ((used >> s) & 1u) wait(free_bars, s);
    used |= 1u << s;
  }
};

}  // namespace
}  // namespace cdan
