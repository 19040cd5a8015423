// Transports of the row-tiled forward (band.cuh): in-process (threads + events + device copies) and NCCL (dlopen'ed).
#include "band.cuh"

#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/cdan_b200.h"
#include "plan.hpp"

namespace cdan {
namespace {

constexpr int kMaxBands = 16;

__global__ void reduce_bands_kernel(const float* __restrict__ gathered, int nb, int count, float* __restrict__ sum, float* __restrict__ mx) {
  // gathered: [band][2][count]; fixed band order -> every band computes bitwise the same result
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f, m = -INFINITY;
  for (int b = 0; b < nb; ++b) {
    s += gathered[(size_t(b) * 2 + 0) * count + i];
    m = fmaxf(m, gathered[(size_t(b) * 2 + 1) * count + i]);
  }
  sum[i] = s;
  mx[i] = m;
}

}  // namespace
}  // namespace cdan

// ------------------------------------------------------------------------------------------------ in-process transport
struct cdan_band_group {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  unsigned gen = 0;
  bool broken = false;
  struct Slot {
    cdan::HaloMsg msg;
    cudaEvent_t ready = nullptr, done = nullptr;
    const float *sum = nullptr, *mx = nullptr;
    float* tmp = nullptr;  // [n bands][2][count] gathered statistics of this band's reduction
    size_t tmp_floats = 0;
    int device = -1;
    bool attached = false;
  } slot[cdan::kMaxBands];

  // all bands meet here; a band that never arrives (its forward failed) breaks the barrier instead of hanging the others
  int barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return -1;
    const unsigned g = gen;
    if (++waiting == n) {
      waiting = 0;
      ++gen;
      cv.notify_all();
      return 0;
    }
    if (!cv.wait_for(lk, std::chrono::seconds(60), [&] { return gen != g || broken; })) {
      broken = true;
      cv.notify_all();
      return -1;
    }
    return broken ? -1 : 0;
  }
};

namespace cdan {
namespace {

struct LocalComm : BandComm {
  cdan_band_group* g = nullptr;

  int sync(const char* what) {
    if (g->barrier() != 0) return fail(std::string("band group: a band did not reach the ") + what + " rendezvous (timeout or earlier failure)");
    return 0;
  }

  int exchange(const HaloMsg& m, cudaStream_t s) override {
    auto& me = g->slot[rank];
    me.msg = m;
    CDAN_CUDA_OK(cudaEventRecord(me.ready, s));
    CDAN_TRY(sync("halo exchange"));
    for (int d = -1; d <= 1; d += 2) {
      const int nb = rank + d;
      if (nb < 0 || nb >= nranks) continue;
      const HaloMsg& o = g->slot[nb].msg;
      if (o.bytes != m.bytes || o.nimg != m.nimg) return fail("band group: neighbouring bands disagree about a halo message");
      CDAN_CUDA_OK(cudaStreamWaitEvent(s, g->slot[nb].ready, 0));
      for (int i = 0; i < m.nimg; ++i) {
        char* dst = m.base + size_t(i) * m.img_stride + (d < 0 ? m.top_recv : m.bot_recv);
        const char* src = o.base + size_t(i) * o.img_stride + (d < 0 ? o.bot_send : o.top_send);
        CDAN_CUDA_OK(cudaMemcpyAsync(dst, src, m.bytes, cudaMemcpyDefault, s));
      }
    }
    CDAN_CUDA_OK(cudaEventRecord(me.done, s));
    CDAN_TRY(sync("halo exchange (completion)"));
    for (int d = -1; d <= 1; d += 2) {
      const int nb = rank + d;
      if (nb >= 0 && nb < nranks) CDAN_CUDA_OK(cudaStreamWaitEvent(s, g->slot[nb].done, 0));
    }
    return 0;
  }

  int allreduce_sum_max(float* sum, float* mx, int count, cudaStream_t s) override {
    auto& me = g->slot[rank];
    const size_t need = size_t(nranks) * 2 * count;
    if (me.tmp_floats < need) {
      CDAN_CUDA_OK(cudaStreamSynchronize(s));
      if (me.tmp) CDAN_CUDA_OK(cudaFree(me.tmp));
      me.tmp = nullptr;
      CDAN_CUDA_OK(cudaMalloc(&me.tmp, need * sizeof(float)));
      me.tmp_floats = need;
    }
    me.sum = sum;
    me.mx = mx;
    CDAN_CUDA_OK(cudaEventRecord(me.ready, s));
    CDAN_TRY(sync("all-reduce"));
    for (int b = 0; b < nranks; ++b) {
      if (b != rank) CDAN_CUDA_OK(cudaStreamWaitEvent(s, g->slot[b].ready, 0));
      CDAN_CUDA_OK(cudaMemcpyAsync(me.tmp + (size_t(b) * 2 + 0) * count, g->slot[b].sum, count * sizeof(float), cudaMemcpyDefault, s));
      CDAN_CUDA_OK(cudaMemcpyAsync(me.tmp + (size_t(b) * 2 + 1) * count, g->slot[b].mx, count * sizeof(float), cudaMemcpyDefault, s));
    }
    CDAN_CUDA_OK(cudaEventRecord(me.done, s));
    CDAN_TRY(sync("all-reduce (completion)"));
    for (int b = 0; b < nranks; ++b)
      if (b != rank) CDAN_CUDA_OK(cudaStreamWaitEvent(s, g->slot[b].done, 0));  // everybody has read my statistics
    reduce_bands_kernel<<<ceil_div(count, 256), 256, 0, s>>>(me.tmp, nranks, count, sum, mx);
    CDAN_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------ NCCL transport
// Minimal declarations of the NCCL 2.x C API (stable ABI); the library is dlopen'ed so that libcdan_b200.so has no link
// dependency on it.
struct NcclId { char internal[128]; };
using NcclCommT = void*;
constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclSum = 0, kNcclMax = 2;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclCommT*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclCommT) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

int nccl_api(NcclApi** out) {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!api.handle) {
    // the soname first: inside a process that already imported torch this resolves to the copy torch loaded
    const char* names[] = {getenv("CDAN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
      if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(std::string("spatial tiling over NCCL: cannot load libnccl.so.2 (") + dlerror() + "); set CDAN_NCCL_LIB");
    bool ok = true;
    auto sym = [&](const char* n) {
      void* f = dlsym(h, n);
      if (!f) ok = false;
      return f;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) return fail("spatial tiling over NCCL: libnccl lacks a required symbol");
    api.handle = h;
  }
  *out = &api;
  return 0;
}

#define CDAN_NCCL_OK(expr)                                                                                          \
  do {                                                                                                              \
    int r__ = (expr);                                                                                               \
    if (r__ != 0) return fail(std::string(#expr) + " failed: " + (api->GetErrorString ? api->GetErrorString(r__) : "?")); \
  } while (0)

struct NcclComm : BandComm {
  NcclApi* api = nullptr;
  NcclCommT comm = nullptr;
  ~NcclComm() override {
    if (comm && api) api->CommDestroy(comm);
  }
  int exchange(const HaloMsg& m, cudaStream_t s) override {
    CDAN_NCCL_OK(api->GroupStart());
    for (int i = 0; i < m.nimg; ++i) {
      char* img = m.base + size_t(i) * m.img_stride;
      if (rank > 0) {
        CDAN_NCCL_OK(api->Send(img + m.top_send, m.bytes, kNcclInt8, rank - 1, comm, s));
        CDAN_NCCL_OK(api->Recv(img + m.top_recv, m.bytes, kNcclInt8, rank - 1, comm, s));
      }
      if (rank + 1 < nranks) {
        CDAN_NCCL_OK(api->Send(img + m.bot_send, m.bytes, kNcclInt8, rank + 1, comm, s));
        CDAN_NCCL_OK(api->Recv(img + m.bot_recv, m.bytes, kNcclInt8, rank + 1, comm, s));
      }
    }
    CDAN_NCCL_OK(api->GroupEnd());
    return 0;
  }
  int allreduce_sum_max(float* sum, float* mx, int count, cudaStream_t s) override {
    CDAN_NCCL_OK(api->GroupStart());
    CDAN_NCCL_OK(api->AllReduce(sum, sum, size_t(count), kNcclFloat32, kNcclSum, comm, s));
    CDAN_NCCL_OK(api->AllReduce(mx, mx, size_t(count), kNcclFloat32, kNcclMax, comm, s));
    CDAN_NCCL_OK(api->GroupEnd());
    return 0;
  }
};

void detach(cdan_plan* p) {
  delete p->band_comm;
  p->band_comm = nullptr;
}

}  // namespace

// Band split: contiguous bands with boundaries at multiples of 8 rows (the three 2x2 max-pools never straddle a boundary);
// the first H/8 % nbands bands are 8 rows taller.  rows_out = {owned begin, owned end, extended begin, extended end}.
int band_rows(int H, int nbands, int rank, int halo, int rows_out[4]) {
  if (H <= 0 || H % 8) return fail("band split: H must be a positive multiple of 8");
  if (nbands < 1 || nbands > kMaxBands || rank < 0 || rank >= nbands) return fail("band split: bad rank / band count");
  if (halo % 8 || halo < 24) return fail("band split: halo must be a multiple of 8 and at least 24 rows (3 rows at 1/8 resolution for the SpatialGate 7x7)");
  const int units = H / 8, base = units / nbands, extra = units % nbands;
  if (nbands > 1 && base * 8 < halo) return fail("band split: bands of " + std::to_string(base * 8) + " rows are thinner than the " + std::to_string(halo) + "-row halo");
  const int r0 = 8 * (rank * base + std::min(rank, extra)), r1 = r0 + 8 * (base + (rank < extra ? 1 : 0));
  rows_out[0] = r0;
  rows_out[1] = r1;
  rows_out[2] = rank > 0 ? r0 - halo : r0;
  rows_out[3] = rank + 1 < nbands ? r1 + halo : r1;
  return 0;
}

}  // namespace cdan

using namespace cdan;

extern "C" {

int cdan_band_group_create(int nbands, cdan_band_group** out) {
  if (!out) return fail("cdan_band_group_create: NULL argument");
  if (nbands < 1 || nbands > kMaxBands) return fail("cdan_band_group_create: 1..16 bands");
  auto* g = new cdan_band_group();
  g->n = nbands;
  *out = g;
  return 0;
}

int cdan_band_group_destroy(cdan_band_group* g) {
  if (!g) return 0;
  for (int i = 0; i < g->n; ++i) {
    auto& sl = g->slot[i];
    if (sl.device >= 0) cudaSetDevice(sl.device);
    if (sl.ready) cudaEventDestroy(sl.ready);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.tmp) cudaFree(sl.tmp);
  }
  delete g;
  return 0;
}

int cdan_plan_band_attach_local(cdan_plan* p, cdan_band_group* g, int rank) {
  if (!p || !g) return fail("cdan_plan_band_attach_local: NULL argument");
  if (rank < 0 || rank >= g->n) return fail("cdan_plan_band_attach_local: rank out of range");
  if (g->slot[rank].attached) return fail("cdan_plan_band_attach_local: this rank already has a plan");
  detach(p);
  cudaSetDevice(p->device);
  auto& sl = g->slot[rank];
  CDAN_CUDA_OK(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
  CDAN_CUDA_OK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  sl.device = p->device;
  sl.attached = true;
  auto* c = new LocalComm();
  c->g = g;
  c->rank = rank;
  c->nranks = g->n;
  p->band_comm = c;
  return 0;
}

int cdan_band_nccl_unique_id(void* id_out, size_t len) {
  if (!id_out || len < sizeof(NcclId)) return fail("cdan_band_nccl_unique_id: need a 128-byte buffer");
  NcclApi* api = nullptr;
  CDAN_TRY(nccl_api(&api));
  NcclId id;
  CDAN_NCCL_OK(api->GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return 0;
}

int cdan_plan_band_attach_nccl(cdan_plan* p, int rank, int nranks, const void* id_bytes, size_t len) {
  if (!p || !id_bytes || len < sizeof(NcclId)) return fail("cdan_plan_band_attach_nccl: NULL argument or short id");
  if (nranks < 1 || nranks > kMaxBands || rank < 0 || rank >= nranks) return fail("cdan_plan_band_attach_nccl: bad rank / nranks");
  NcclApi* api = nullptr;
  CDAN_TRY(nccl_api(&api));
  detach(p);
  cudaSetDevice(p->device);
  NcclId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  auto* c = new NcclComm();
  c->api = api;
  c->rank = rank;
  c->nranks = nranks;
  int r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    c->comm = nullptr;
    delete c;
    return fail(std::string("ncclCommInitRank failed: ") + api->GetErrorString(r));
  }
  p->band_comm = c;
  return 0;
}

int cdan_plan_band_detach(cdan_plan* p) {
  if (p) detach(p);
  return 0;
}

int cdan_band_rows(int H, int nbands, int rank, int halo, int rows_out[4]) {
  if (!rows_out) return fail("cdan_band_rows: NULL argument");
  return band_rows(H, nbands, rank, halo, rows_out);
}

int cdan_band_stats(cdan_plan* p, long long out[3]) {
  if (!p || !out) return fail("cdan_band_stats: NULL argument");
  out[0] = p->band_stats.exchanges;
  out[1] = p->band_stats.halo_bytes_received;
  out[2] = p->band_stats.allreduces;
  return 0;
}

}  // extern "C"
