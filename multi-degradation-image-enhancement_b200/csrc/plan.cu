// C-ABI entry points: plan life cycle, weight folding/packing, the CDAN forward schedule, stage taps.
#include "plan.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "../../include/cdan_b200.h"
#include "conv_umma.cuh"

namespace cdan {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(const std::string& msg) {
  g_error = msg;
  return -1;
}

namespace {

constexpr double kBnEps = 1e-5;  // nn.BatchNorm2d default eps (models/cdan.py:12; models/cbam.py:11)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    // always: since CUDA 12 cudaSetDevice also binds the primary context to the calling thread, which the driver-API calls
    // of the kernels' launchers (cuTensorMapEncodeTiled) need in a host thread that has made no runtime call yet
    cudaSetDevice(dev);
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

using HostDict = std::map<std::string, std::vector<float>>;

int need(const HostDict& sd, const std::string& key, size_t numel, const float** out) {
  auto it = sd.find(key);
  if (it == sd.end()) return fail("state_dict is missing key '" + key + "'");
  if (it->second.size() != numel)
    return fail("state_dict key '" + key + "' has " + std::to_string(it->second.size()) + " elements, expected " +
                std::to_string(numel));
  *out = it->second.data();
  return 0;
}

// eval-mode BatchNorm as y = s*x + t
int bn_affine(const HostDict& sd, const std::string& prefix, int C, std::vector<double>& s, std::vector<double>& t) {
  const float *g, *b, *m, *v;
  CDAN_TRY(need(sd, prefix + ".weight", C, &g));
  CDAN_TRY(need(sd, prefix + ".bias", C, &b));
  CDAN_TRY(need(sd, prefix + ".running_mean", C, &m));
  CDAN_TRY(need(sd, prefix + ".running_var", C, &v));
  s.resize(C);
  t.resize(C);
  for (int c = 0; c < C; ++c) {
    s[c] = double(g[c]) / std::sqrt(double(v[c]) + kBnEps);
    t[c] = double(b[c]) - double(m[c]) * s[c];
  }
  return 0;
}

int upload(cdan_plan* p, const std::vector<float>& h, float** d) {
  CDAN_CUDA_OK(cudaMalloc(d, h.size() * sizeof(float)));
  p->owned.push_back(*d);
  CDAN_CUDA_OK(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// Pack one convolution for the CUDA-core kernel: fp32 [taps][CinPhys][CoutP].
//   w         : Conv2d weight [Cout][Cin][ks][ks], or ConvTranspose2d weight [Cin][Cout][ks][ks] when `transposed`
//               (then the equivalent correlation kernel is W'[o][i][u][v] = W[i][o][ks-1-u][ks-1-v])
//   post_s/t  : per-output-channel affine of a FOLLOWING BatchNorm (folded into weight and bias), may be empty
//   cmap      : logical input channel -> physical channel of the NHWC buffer
int pack_conv(cdan_plan* p, ConvLayer& L, const float* w, const float* bias, int Cout, int CinLogical, int ks,
              bool transposed, const std::vector<double>& post_s, const std::vector<double>& post_t,
              const std::vector<int>& cmap, int CinPhys) {
  L.Cin = CinPhys;
  L.Cout = Cout;
  L.ks = ks;
  L.CoutP = int(align_up(size_t(Cout), 16));
  const int taps = ks * ks;
  std::vector<float> hw(size_t(taps) * CinPhys * L.CoutP, 0.f), hb(L.CoutP, 0.f);
  for (int o = 0; o < Cout; ++o) {
    const double s = post_s.empty() ? 1.0 : post_s[o];
    for (int ci = 0; ci < CinLogical; ++ci)
      for (int u = 0; u < ks; ++u)
        for (int v = 0; v < ks; ++v) {
          const double val = transposed ? w[((size_t(ci) * Cout + o) * ks + (ks - 1 - u)) * ks + (ks - 1 - v)]
                                        : w[((size_t(o) * CinLogical + ci) * ks + u) * ks + v];
          hw[(size_t(u * ks + v) * CinPhys + cmap[ci]) * L.CoutP + o] = float(val * s);
        }
    hb[o] = float(post_s.empty() ? double(bias[o]) : double(bias[o]) * s + post_t[o]);
  }
  CDAN_TRY(upload(p, hw, &L.d_w));
  CDAN_TRY(upload(p, hb, &L.d_bias));
  if (Cout <= 16) {  // narrow layers: keep the host image for fused re-packing (final dense block)
    L.h_w = hw;
    L.h_bias = hb;
  }
  if (p->dt == kBF16) CDAN_TRY(umma_pack_create(hw.data(), hb.data(), CinPhys, Cout, L.CoutP, ks, &L.umma));
  return 0;
}

int pack_pre(cdan_plan* p, ConvLayer& L, const std::vector<double>& s, const std::vector<double>& t,
             const std::vector<int>& cmap, int CinPhys) {
  std::vector<float> hs(CinPhys, 0.f), ht(CinPhys, 0.f);  // pad channels: relu(0*x + 0) = 0
  for (size_t c = 0; c < s.size(); ++c) {
    hs[cmap[c]] = float(s[c]);
    ht[cmap[c]] = float(t[c]);
  }
  CDAN_TRY(upload(p, hs, &L.d_pre_s));
  CDAN_TRY(upload(p, ht, &L.d_pre_t));
  L.h_pre_s = hs;
  L.h_pre_t = ht;
  return 0;
}

std::vector<int> dense_cmap(int C, int Cpad, int n_logical) {
  std::vector<int> m(n_logical);
  for (int j = 0; j < n_logical; ++j) m[j] = j < C ? j : Cpad + (j - C);
  return m;
}

int load_conv_block(cdan_plan* p, const HostDict& sd, ConvId id, const std::string& prefix, int Ci, int Co) {
  const float *w, *b;
  CDAN_TRY(need(sd, prefix + ".conv.weight", size_t(Co) * Ci * 9, &w));
  CDAN_TRY(need(sd, prefix + ".conv.bias", Co, &b));
  std::vector<double> s, t;
  CDAN_TRY(bn_affine(sd, prefix + ".bn", Co, s, t));
  ConvLayer& L = p->conv[id];
  L.name = prefix;
  L.relu = 1;
  return pack_conv(p, L, w, b, Co, Ci, 3, false, s, t, dense_cmap(Ci, Ci, Ci), Ci);
}

int load_dense_block(cdan_plan* p, const HostDict& sd, ConvId first, const std::string& prefix, int C, int Cpad,
                     int Cout) {
  for (int l = 0; l < 4; ++l) {
    const int ci = C + 16 * l, ciPhys = Cpad + 16 * l;
    const std::string lp = prefix + ".layers." + std::to_string(l);
    const float *w, *b;
    CDAN_TRY(need(sd, lp + ".2.weight", size_t(16) * ci * 9, &w));
    CDAN_TRY(need(sd, lp + ".2.bias", 16, &b));
    std::vector<double> s, t;
    CDAN_TRY(bn_affine(sd, lp + ".0", ci, s, t));
    ConvLayer& L = p->conv[first + l];
    L.name = lp;
    L.relu = 0;
    const std::vector<int> cmap = dense_cmap(C, Cpad, ci);
    CDAN_TRY(pack_conv(p, L, w, b, 16, ci, 3, false, {}, {}, cmap, ciPhys));
    CDAN_TRY(pack_pre(p, L, s, t, cmap, ciPhys));
  }
  const int ci = C + 64, ciPhys = Cpad + 64;
  const std::string tp = prefix + ".transition_layer";
  const float *w, *b;
  CDAN_TRY(need(sd, tp + ".2.weight", size_t(Cout) * ci, &w));
  CDAN_TRY(need(sd, tp + ".2.bias", Cout, &b));
  std::vector<double> s, t;
  CDAN_TRY(bn_affine(sd, tp + ".0", ci, s, t));
  ConvLayer& L = p->conv[first + 4];
  L.name = tp;
  L.relu = 0;
  const std::vector<int> cmap = dense_cmap(C, Cpad, ci);
  CDAN_TRY(pack_conv(p, L, w, b, Cout, ci, 1, false, {}, {}, cmap, ciPhys));
  CDAN_TRY(pack_pre(p, L, s, t, cmap, ciPhys));
  return 0;
}

int load_decoder_conv(cdan_plan* p, const HostDict& sd, ConvId id, int idx, int Ci, int Co) {
  const std::string cp = "decoder.conv" + std::to_string(idx), bp = "decoder.bn" + std::to_string(idx);
  const float *w, *b;
  CDAN_TRY(need(sd, cp + ".weight", size_t(Ci) * Co * 9, &w));
  CDAN_TRY(need(sd, cp + ".bias", Co, &b));
  std::vector<double> s, t;
  CDAN_TRY(bn_affine(sd, bp, Co, s, t));
  ConvLayer& L = p->conv[id];
  L.name = cp;
  L.relu = 1;
  return pack_conv(p, L, w, b, Co, Ci, 3, true, s, t, dense_cmap(Ci, Ci, Ci), Ci);
}

int load_cbam(cdan_plan* p, const HostDict& sd, int slot, const std::string& prefix, int C) {
  const int R = C / 16;
  const float *w1, *b1, *w2, *b2, *w7;
  CDAN_TRY(need(sd, prefix + ".ChannelGate.mlp.1.weight", size_t(R) * C, &w1));
  CDAN_TRY(need(sd, prefix + ".ChannelGate.mlp.1.bias", R, &b1));
  CDAN_TRY(need(sd, prefix + ".ChannelGate.mlp.3.weight", size_t(C) * R, &w2));
  CDAN_TRY(need(sd, prefix + ".ChannelGate.mlp.3.bias", C, &b2));
  CDAN_TRY(need(sd, prefix + ".SpatialGate.spatial.conv.weight", 98, &w7));
  std::vector<double> s, t;
  CDAN_TRY(bn_affine(sd, prefix + ".SpatialGate.spatial.bn", 1, s, t));
  CbamLayer& L = p->cbam[slot];
  L.C = C;
  float* d;
  CDAN_TRY(upload(p, std::vector<float>(w1, w1 + size_t(R) * C), &d)); L.w.w1 = d;
  CDAN_TRY(upload(p, std::vector<float>(b1, b1 + R), &d)); L.w.b1 = d;
  CDAN_TRY(upload(p, std::vector<float>(w2, w2 + size_t(C) * R), &d)); L.w.w2 = d;
  CDAN_TRY(upload(p, std::vector<float>(b2, b2 + C), &d)); L.w.b2 = d;
  CDAN_TRY(upload(p, std::vector<float>(w7, w7 + 98), &d)); L.w.w7 = d;
  L.w.bn_a = float(s[0]);
  L.w.bn_b = float(t[0]);
  return 0;
}

void free_weights(cdan_plan* p) {
  for (void* d : p->owned) cudaFree(d);
  p->owned.clear();
  for (auto& L : p->conv) {
    umma_pack_destroy(L.umma);
    L = ConvLayer{};
  }
  fused_fd_pack_destroy(p->fd_fused);
  p->fd_fused = nullptr;
  p->loaded = false;
}

// Channel stride of the final dense block's concat buffer: 80 channels are used, the pixel is padded to 128 channels
// (256 B) so that every channel-prefix read starts and ends on a 64-byte DRAM atom (measured: 160-byte pixels make the
// 128-byte prefix of layer 3 fetch the whole buffer and cost ~6 % of the step; 192-byte pixels are worse still).
// CDAN_FD_PLANAR=0 keeps the final dense block NHWC (A/B switch); the planar form needs the streaming kernels.
bool fd_planar_enabled() {
  static const bool v = !(getenv("CDAN_FD_PLANAR") && atoi(getenv("CDAN_FD_PLANAR")) == 0) &&
                        !(getenv("CDAN_CONV_STREAM") && atoi(getenv("CDAN_CONV_STREAM")) == 0);
  return v;
}
// CDAN_DENSE_HYBRID=0 keeps the concat buffers of dense blocks 1-3 NHWC (A/B switch).
bool dense_hybrid_enabled() {
  static const bool v = !(getenv("CDAN_DENSE_HYBRID") && atoi(getenv("CDAN_DENSE_HYBRID")) == 0) &&
                        !(getenv("CDAN_CONV_STREAM") && atoi(getenv("CDAN_CONV_STREAM")) == 0);
  return v;
}
// CDAN_FD_FUSED=0 runs the final dense block layer by layer (A/B switch and the source of the "dec.final_in" stage tap).
bool fd_fused_enabled() {
  static const bool v = !(getenv("CDAN_FD_FUSED") && atoi(getenv("CDAN_FD_FUSED")) == 0);
  return v;
}
int fd_ld() {
  static const int v = getenv("CDAN_FD_LD") ? atoi(getenv("CDAN_FD_LD")) : 128;
  return v;
}

// ------------------------------------------------------------------------------------------ workspace
size_t carve(Buffers& b, char* base, DType dt, int N, int H, int W, bool fd_fused) {
  const size_t es = dt == kF32 ? 4 : 2;
  const size_t p1 = size_t(N) * H * W, p2 = p1 / 4, p4 = p1 / 16, p8 = p1 / 64;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  b.D1 = take(p2 * 128 * es);
  b.DN1 = take(p2 * 64 * es);
  b.D2 = take(p4 * 192 * es);
  b.DN2 = take(p4 * 128 * es);
  b.D3 = take(p8 * 320 * es);
  b.DN3 = take(p8 * 256 * es);
  b.E4 = take(p8 * 512 * es);
  b.B0 = take(p8 * 512 * es);
  b.T1 = take(p8 * 256 * es);
  b.A1 = take(p8 * 256 * es);
  b.C1 = take(p8 * 256 * es);
  b.T2 = take(p8 * 128 * es);
  b.U2 = take(p4 * 128 * es);
  b.C2 = take(p4 * 128 * es);
  b.T3 = take(p4 * 64 * es);
  b.U3 = take(p2 * 64 * es);
  b.C3 = take(p2 * 64 * es);
  b.T4 = take(p2 * 8 * es);
  b.FD = take(fd_fused ? 0 : p1 * size_t(fd_ld()) * es);  // final dense block concat: 3 input channels padded to 16, then 4 x 16
  size_t sc = 0;
  sc = std::max(sc, cbam_scratch_floats(N, 512, H / 8, W / 8));
  sc = std::max(sc, cbam_scratch_floats(N, 256, H / 8, W / 8));
  sc = std::max(sc, cbam_scratch_floats(N, 128, H / 4, W / 4));
  sc = std::max(sc, cbam_scratch_floats(N, 64, H / 2, W / 2));
  b.cbam_scratch = (float*)take(sc * sizeof(float));
  b.total_bytes = off;
  return off;
}

bool use_fd_fused(const cdan_plan* p) { return p->dt == kBF16 && p->conv_impl == 0 && p->fd_fused && p->fd_fused_on && fd_fused_enabled(); }

int ensure_workspace(cdan_plan* p, int N, int H, int W) {
  Buffers probe{};
  const bool fused = use_fd_fused(p);
  const size_t bytes = carve(probe, nullptr, p->dt, N, H, W, fused);
  if (bytes > p->ws_bytes) {
    if (p->ws) CDAN_CUDA_OK(cudaFree(p->ws));
    p->ws = nullptr;
    p->ws_bytes = 0;
    CDAN_CUDA_OK(cudaMalloc(&p->ws, bytes));
    p->ws_bytes = bytes;
  }
  carve(p->buf, (char*)p->ws, p->dt, N, H, W, fused);
  return 0;
}

// ------------------------------------------------------------------------------------------ profiling spans
cudaEvent_t get_event(cdan_plan* p) {
  if (!p->event_pool.empty()) {
    cudaEvent_t e = p->event_pool.back();
    p->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct SpanGuard {  // records a pair of events around the launches issued in its scope
  cdan_plan* p;
  cudaStream_t s;
  size_t idx = size_t(-1);
  SpanGuard(cdan_plan* plan, cudaStream_t stream, const std::string& label) : p(plan), s(stream) {
    if (!p->profile) return;
    p->spans.push_back({label, get_event(p), get_event(p)});
    idx = p->spans.size() - 1;
    cudaEventRecord(p->spans[idx].e0, s);
  }
  ~SpanGuard() {
    if (idx != size_t(-1)) cudaEventRecord(p->spans[idx].e1, s);
  }
};

// ------------------------------------------------------------------------------------------ schedule
int conv_dispatch(cdan_plan* p, const ConvLayer& L, const ConvDesc& d, cudaStream_t s);  // below

int run_conv(cdan_plan* p, ConvId id, int N, int H, int W, const void* in, int in_ld, void* out, int out_ld,
             int pool, cudaStream_t s, const float* in_nchw = nullptr, float* out_nchw = nullptr, int sigmoid = 0,
             size_t in_gstride = 0, const void* in2 = nullptr, int Chead = 0) {
  const ConvLayer& L = p->conv[id];
  ConvDesc d;
  d.N = N; d.H = H; d.W = W;
  d.Cin = L.Cin; d.Cout = L.Cout; d.ks = L.ks;
  d.in = in; d.in_ld = in_ld; d.in_nchw = in_nchw; d.in_gstride = in_gstride; d.in2 = in2; d.Chead = Chead;
  d.pre_scale = L.d_pre_s; d.pre_shift = L.d_pre_t;
  d.w = L.d_w; d.CoutP = L.CoutP; d.bias = L.d_bias;
  d.relu = L.relu; d.pool = pool;
  d.out = out; d.out_ld = out_ld; d.out_nchw = out_nchw; d.sigmoid = sigmoid;
  const bool umma = p->dt == kBF16 && p->conv_impl == 0 && L.umma && conv_umma_supported(d);
  SpanGuard span(p, s, std::string("conv|") + L.name + (umma ? (d.pre_scale || d.in_nchw ? "|umma_pro" : "|umma_tma") : "|simt"));
  CDAN_TRY(conv_dispatch(p, L, d, s));
  static const bool debug_sync = getenv("CDAN_DEBUG_SYNC") != nullptr;
  if (debug_sync) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail("kernel of layer '" + L.name + "' failed: " + cudaGetErrorString(e));
  }
  p->launches += umma ? conv_umma_kernel_count(d, *L.umma) : 1;
  return 0;
}

// bf16 plans run every convolution on the tcgen05 kernel; fp32 plans (and conv_impl=1) use the CUDA-core kernel.
int conv_dispatch(cdan_plan* p, const ConvLayer& L, const ConvDesc& d, cudaStream_t s) {
  if (p->dt == kBF16 && p->conv_impl == 0 && L.umma && conv_umma_supported(d)) return conv_umma_launch(d, *L.umma, s);
  return conv_simt_launch(d, p->dt, s);
}

inline char* at(void* base, size_t elems, DType dt) { return (char*)base + elems * (dt == kF32 ? 4 : 2); }

// `pre(l)` runs before layer l (0-3: growth layers, 4: transition) is launched: the row-tiled forward refreshes halo rows there.
using LayerHook = std::function<int(int)>;

int run_dense(cdan_plan* p, ConvId first, int N, int h, int w, void* D, int ld, int Cpad, void* DN, int Cout,
              cudaStream_t s, float* out_nchw = nullptr, const LayerHook& pre = nullptr) {
  for (int l = 0; l < 4; ++l) {
    if (pre) CDAN_TRY(pre(l));
    CDAN_TRY(run_conv(p, ConvId(first + l), N, h, w, D, ld, at(D, Cpad + 16 * l, p->dt), ld, 0, s));
  }
  (void)Cout;
  if (pre) CDAN_TRY(pre(4));
  return run_conv(p, ConvId(first + 4), N, h, w, D, ld, DN, DN ? p->conv[first + 4].Cout : 0, 0, s, nullptr, out_nchw,
                  out_nchw ? 1 : 0);
}

// final dense block in the GROUP-PLANAR layout (bf16 tensor-core plans): plane g of D holds channels [16g, 16g+16) as a
// dense [N][h][w][16] tensor; layer l reads planes 0..l and writes plane l+1, the transition reads all five.
int run_dense_planar(cdan_plan* p, ConvId first, int N, int h, int w, void* D, cudaStream_t s, float* out_nchw,
                     const LayerHook& pre = nullptr) {
  const size_t gs = size_t(N) * h * w * 16;
  for (int l = 0; l < 4; ++l) {
    if (pre) CDAN_TRY(pre(l));
    CDAN_TRY(run_conv(p, ConvId(first + l), N, h, w, D, 16, at(D, gs * (1 + l), p->dt), 16, 0, s, nullptr, nullptr, 0, gs));
  }
  if (pre) CDAN_TRY(pre(4));
  return run_conv(p, ConvId(first + 4), N, h, w, D, 16, nullptr, 0, 0, s, nullptr, out_nchw, 1, gs);
}

// dense blocks 1-3 on a HYBRID concat buffer (bf16 tensor-core plans): the pooled ConvBlock output stays a compact NHWC
// head of Cpre channels (the next ConvBlock, the decoder skip connection and the stage taps read it), the four 16-channel
// groups the layers append are dense planes behind it.  Layer 0 reads only the head.
int run_dense_hybrid(cdan_plan* p, ConvId first, int N, int h, int w, void* D, int Cpre, void* DN, cudaStream_t s,
                     const LayerHook& pre = nullptr) {
  const size_t gs = size_t(N) * h * w * 16;
  void* planes = at(D, size_t(N) * h * w * Cpre, p->dt);
  if (pre) CDAN_TRY(pre(0));
  CDAN_TRY(run_conv(p, first, N, h, w, D, Cpre, planes, 16, 0, s));
  for (int l = 1; l < 4; ++l) {
    if (pre) CDAN_TRY(pre(l));
    CDAN_TRY(run_conv(p, ConvId(first + l), N, h, w, D, Cpre, at(planes, gs * l, p->dt), 16, 0, s, nullptr, nullptr, 0, gs, planes, Cpre));
  }
  if (pre) CDAN_TRY(pre(4));
  return run_conv(p, ConvId(first + 4), N, h, w, D, Cpre, DN, p->conv[first + 4].Cout, 0, s, nullptr, nullptr, 0, gs, planes, Cpre);
}

int run_cbam(cdan_plan* p, int slot, const void* x, const void* mul, void* out, int N, int h, int w, bool pooled,
             cudaStream_t s, const CbamBand* band = nullptr) {
  const CbamLayer& L = p->cbam[slot];
  CbamScratch sc;
  cbam_scratch_carve(p->buf.cbam_scratch, N, L.C, h, w, &sc);
  SpanGuard span(p, s, "cbam|C" + std::to_string(L.C));
  CDAN_TRY(cbam_launch(p->dt, x, L.C, mul, L.C, out, L.C, N, h, w, L.C, L.w, sc, pooled, s, band));
  p->launches += band ? 6 : (pooled ? 4 : 5);
  if (band) p->band_stats.allreduces += 2;
  return 0;
}

// decoder stage glue (models/cdan.py:130,137-138,145-146): out = [up](a) + skip, with the following CBAM's pooling
// partials written by the same kernel
int run_up_add(cdan_plan* p, int cbam_slot, const void* a, int a_ld, const void* skip, int skip_ld, void* out, int N,
               int oh, int ow, int up, cudaStream_t s) {
  const int C = p->cbam[cbam_slot].C;
  CbamScratch sc;
  cbam_scratch_carve(p->buf.cbam_scratch, N, C, oh, ow, &sc);
  SpanGuard span(p, s, "glue|up_add_C" + std::to_string(C));
  CDAN_TRY(up_add_launch(p->dt, a, a_ld, skip, skip_ld, out, C, N, oh, ow, C, up, s, sc.psum, sc.pmax));
  p->launches += 1;
  return 0;
}

void fill_stages(cdan_plan* p, int ld1, int ld2, int ld3, int fdl) {
  Buffers& b = p->buf;
  const int H = p->H, W = p->W;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
  auto& st = p->stages;
  st.clear();
  st["enc.out1"] = {b.D1, 64, ld1, H2, W2};
  st["enc.dense1"] = {b.DN1, 64, 64, H2, W2};
  st["enc.out2"] = {b.D2, 128, ld2, H4, W4};
  st["enc.dense2"] = {b.DN2, 128, 128, H4, W4};
  st["enc.out3"] = {b.D3, 256, ld3, H8, W8};
  st["enc.dense3"] = {b.DN3, 256, 256, H8, W8};
  st["enc.conv4"] = {b.E4, 512, 512, H8, W8};
  st["bottleneck"] = {b.B0, 512, 512, H8, W8};
  st["dec.bn1"] = {b.T1, 256, 256, H8, W8};
  st["dec.gated1"] = {b.C1, 256, 256, H8, W8};
  st["dec.bn2"] = {b.T2, 128, 128, H8, W8};
  st["dec.gated2"] = {b.C2, 128, 128, H4, W4};
  st["dec.bn3"] = {b.T3, 64, 64, H4, W4};
  st["dec.gated3"] = {b.C3, 64, 64, H2, W2};
  st["dec.bn4"] = {b.T4, 3, 8, H2, W2};
  if (fdl) st["dec.final_in"] = {b.FD, 3, fdl, H, W};  // not materialised by the fused final dense block
}

// ------------------------------------------------------------------------------------------ row-tiled forward (band.cuh)
// A tensor of the schedule as the band bookkeeping sees it: where its rows live and how many rows next to an artificial
// (band) border are wrong.  Aliased tensors (the groups of an NHWC concat buffer) share one Ten.
struct Ten {
  void* p = nullptr;
  int ld = 0;    // elements between pixels
  int lvl = 0;   // resolution level: rows = H >> lvl
  int dirt = 0;  // wrong rows at each artificial border (at this tensor's resolution)
};
struct BandRun {
  cdan_plan* p;
  cudaStream_t s;
  int N, Hext, W, halo, dtop, dbot, Hfull;  // extended band; dtop/dbot = halo rows above / below (0 at the image border)
  int D(int lvl) const { return halo >> lvl; }
  // replace the halo rows of t by the neighbours' owned rows
  int refresh(Ten& t) {
    const size_t es = p->dt == kF32 ? 4 : 2;
    const int d = D(t.lvl), h = Hext >> t.lvl;
    const size_t row = size_t(W >> t.lvl) * t.ld * es;
    HaloMsg m;
    m.base = (char*)t.p;
    m.img_stride = size_t(h) * row;
    m.nimg = N;
    m.bytes = size_t(d) * row;
    m.top_recv = 0;
    m.top_send = size_t(dtop >> t.lvl) * row;
    m.bot_recv = size_t(h - (dbot >> t.lvl)) * row;
    m.bot_send = m.bot_recv - m.bytes;
    {
      SpanGuard span(p, s, "band|halo_refresh");
      CDAN_TRY(p->band_comm->exchange(m, s));
    }
    const int nbrs = (p->band_comm->rank > 0) + (p->band_comm->rank + 1 < p->band_comm->nranks);
    p->band_stats.exchanges += 1;
    p->band_stats.halo_bytes_received += (long long)(m.bytes) * N * nbrs;
    t.dirt = 0;
    return 0;
  }
};

// The CDAN forward.  br == nullptr: a whole image.  Otherwise x / y are the rows of the EXTENDED band and the schedule
// interleaves halo refreshes: `need(t, e)` before a layer that looks e rows across the border (3x3: 1, 7x7: 3, bilinear x2:
// checked as 2*dirt+1) refreshes t only if the dirty zone would otherwise grow past the halo into the owned rows.
int forward_impl(cdan_plan* p, cudaStream_t s, const float* x, float* y, int N, int H, int W, BandRun* br = nullptr) {
  if (!p->loaded) return fail("cdan_forward: no weights loaded (call cdan_plan_load_weights first)");
  if (N <= 0 || H <= 0 || W <= 0) return fail("cdan_forward: empty input");
  if (H % 8 || W % 8)
    return fail("cdan_forward: H and W must be multiples of 8 (got " + std::to_string(H) + "x" + std::to_string(W) +
                "); the reference fails with a size mismatch for such inputs (models/cdan.py:75-89,137-154)");
  CDAN_TRY(ensure_workspace(p, N, H, W));
  Buffers& b = p->buf;
  const DType dt = p->dt;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
  p->launches = 0;
  p->N = N; p->H = H; p->W = W;

  auto need = [&](Ten& t, int e) -> int { return br && t.dirt + e > br->D(t.lvl) ? br->refresh(t) : 0; };
  auto need_up = [&](Ten& t) -> int { return br && 2 * t.dirt + 1 > br->D(t.lvl - 1) ? br->refresh(t) : 0; };
  CbamBand cb[4];
  auto band_of = [&](int slot, int lvl) -> const CbamBand* {
    if (!br) return nullptr;
    cb[slot].row0 = br->dtop >> lvl;
    cb[slot].rows = (br->Hext - br->dtop - br->dbot) >> lvl;
    cb[slot].HW_full = (br->Hfull >> lvl) * (W >> lvl);
    cb[slot].comm = p->band_comm;
    return &cb[slot];
  };

  // ---- Encoder (models/cdan.py:70-98).  Pooled ConvBlock outputs land in channels [0,C) of the dense-block
  //      concat buffers, so they double as skip connections and as the first `features` entry.
  // Concat buffers of dense blocks 1-3: hybrid (compact NHWC head + group planes) on the tensor-core path, else NHWC.
  const bool hyb = dt == kBF16 && p->conv_impl == 0 && dense_hybrid_enabled();
  const int ld1 = hyb ? 64 : 128, ld2 = hyb ? 128 : 192, ld3 = hyb ? 256 : 320;
  // dense block on concat buffer D at level lvl: head = the block input, g[1..4] = the groups the layers append (hybrid:
  // planes behind the compact head; NHWC: the same rows as the head, one Ten)
  struct DenseTens {
    Ten head, plane[4], out;
    Ten* g[5];
  };
  auto dense_tens = [&](DenseTens& d, void* D, int Cpre, int ld, int lvl, int h, int w, bool planar, void* DN, int Cdn, int head_dirt) {
    d.head = Ten{D, ld, lvl, head_dirt};
    d.g[0] = &d.head;
    for (int l = 0; l < 4; ++l) {
      d.plane[l] = Ten{at(at(D, size_t(N) * h * w * Cpre, dt), size_t(N) * h * w * 16 * l, dt), 16, lvl, 0};
      d.g[1 + l] = planar ? &d.plane[l] : &d.head;
    }
    d.out = Ten{DN, Cdn, lvl, 0};
  };
  auto dense_hook = [&](DenseTens& d) -> LayerHook {
    if (!br) return nullptr;
    return [&d, &need](int l) -> int {  // layer l reads groups 0..l (3x3: one row across the border; transition l = 4: none)
      const int e = l < 4 ? 1 : 0, last = l < 4 ? l : 4;
      int worst = 0;
      for (int i = 0; i <= last; ++i) {
        CDAN_TRY(need(*d.g[i], e));
        worst = std::max(worst, d.g[i]->dirt);
      }
      Ten& o = l < 4 ? *d.g[l + 1] : d.out;
      o.dirt = std::max(o.p == d.head.p && l < 4 ? o.dirt : 0, worst + e);
      return 0;
    };
  };
  DenseTens t1, t2, t3;
  CDAN_TRY(run_conv(p, ENC1, N, H, W, nullptr, 0, b.D1, ld1, 1, s, x));  // input rows are never dirty: conv (+1), pool (ceil /2)
  dense_tens(t1, b.D1, 64, ld1, 1, H2, W2, hyb, b.DN1, 64, 1);
  if (hyb) CDAN_TRY(run_dense_hybrid(p, D1L0, N, H2, W2, b.D1, 64, b.DN1, s, dense_hook(t1)));
  else CDAN_TRY(run_dense(p, D1L0, N, H2, W2, b.D1, 128, 64, b.DN1, 64, s, nullptr, dense_hook(t1)));
  CDAN_TRY(need(t1.head, 1));
  CDAN_TRY(run_conv(p, ENC2, N, H2, W2, b.D1, ld1, b.D2, ld2, 1, s));
  dense_tens(t2, b.D2, 128, ld2, 2, H4, W4, hyb, b.DN2, 128, (t1.head.dirt + 2) / 2);
  if (hyb) CDAN_TRY(run_dense_hybrid(p, D2L0, N, H4, W4, b.D2, 128, b.DN2, s, dense_hook(t2)));
  else CDAN_TRY(run_dense(p, D2L0, N, H4, W4, b.D2, 192, 128, b.DN2, 128, s, nullptr, dense_hook(t2)));
  CDAN_TRY(need(t2.head, 1));
  CDAN_TRY(run_conv(p, ENC3, N, H4, W4, b.D2, ld2, b.D3, ld3, 1, s));
  dense_tens(t3, b.D3, 256, ld3, 3, H8, W8, hyb, b.DN3, 256, (t2.head.dirt + 2) / 2);
  if (hyb) CDAN_TRY(run_dense_hybrid(p, D3L0, N, H8, W8, b.D3, 256, b.DN3, s, dense_hook(t3)));
  else CDAN_TRY(run_dense(p, D3L0, N, H8, W8, b.D3, 320, 256, b.DN3, 256, s, nullptr, dense_hook(t3)));
  CDAN_TRY(need(t3.head, 1));
  CDAN_TRY(run_conv(p, ENC4, N, H8, W8, b.D3, ld3, b.E4, 512, 0, s));
  Ten E4{b.E4, 512, 3, t3.head.dirt + 1};
  // ---- bottleneck CBAM(512) (models/cdan.py:173)
  CDAN_TRY(need(E4, 3));
  CDAN_TRY(run_cbam(p, 0, b.E4, nullptr, b.B0, N, H8, W8, false, s, band_of(0, 3)));
  Ten B0{b.B0, 512, 3, E4.dirt + 3};
  // ---- Decoder (models/cdan.py:126-159)
  CDAN_TRY(need(B0, 1));
  CDAN_TRY(run_conv(p, DEC1, N, H8, W8, b.B0, 512, b.T1, 256, 0, s));
  CDAN_TRY(run_up_add(p, 1, b.T1, 256, b.D3, ld3, b.A1, N, H8, W8, 0, s));
  Ten A1{b.A1, 256, 3, std::max(B0.dirt + 1, t3.head.dirt)};
  CDAN_TRY(need(A1, 3));
  CDAN_TRY(run_cbam(p, 1, b.A1, b.DN3, b.C1, N, H8, W8, true, s, band_of(1, 3)));
  Ten C1{b.C1, 256, 3, std::max(A1.dirt + 3, t3.out.dirt)};
  CDAN_TRY(need(C1, 1));
  CDAN_TRY(run_conv(p, DEC2, N, H8, W8, b.C1, 256, b.T2, 128, 0, s));
  Ten T2{b.T2, 128, 3, C1.dirt + 1};
  CDAN_TRY(need_up(T2));
  CDAN_TRY(run_up_add(p, 2, b.T2, 128, b.D2, ld2, b.U2, N, H4, W4, 1, s));
  Ten U2{b.U2, 128, 2, std::max(2 * T2.dirt + 1, t2.head.dirt)};
  CDAN_TRY(need(U2, 3));
  CDAN_TRY(run_cbam(p, 2, b.U2, b.DN2, b.C2, N, H4, W4, true, s, band_of(2, 2)));
  Ten C2{b.C2, 128, 2, std::max(U2.dirt + 3, t2.out.dirt)};
  CDAN_TRY(need(C2, 1));
  CDAN_TRY(run_conv(p, DEC3, N, H4, W4, b.C2, 128, b.T3, 64, 0, s));
  Ten T3{b.T3, 64, 2, C2.dirt + 1};
  CDAN_TRY(need_up(T3));
  CDAN_TRY(run_up_add(p, 3, b.T3, 64, b.D1, ld1, b.U3, N, H2, W2, 1, s));
  Ten U3{b.U3, 64, 1, std::max(2 * T3.dirt + 1, t1.head.dirt)};
  CDAN_TRY(need(U3, 3));
  CDAN_TRY(run_cbam(p, 3, b.U3, b.DN1, b.C3, N, H2, W2, true, s, band_of(3, 1)));
  Ten C3{b.C3, 64, 1, std::max(U3.dirt + 3, t1.out.dirt)};
  CDAN_TRY(need(C3, 1));
  CDAN_TRY(run_conv(p, DEC4, N, H2, W2, b.C3, 64, b.T4, 8, 0, s));
  Ten T4{b.T4, 8, 1, C3.dirt + 1};
  int out_dirt = 0;
  if (use_fd_fused(p)) {
    // bilinear x2 + x, the final DenseBlock(3,3,16,4) and the sigmoid as ONE kernel (dense_fused.cu): the 67-channel
    // full-resolution concat never exists in HBM.  Band mode: bilinear (2d+1) and four 3x3 layers (+4) inside the kernel.
    if (br && 2 * T4.dirt + 5 > br->D(0)) CDAN_TRY(br->refresh(T4));
    out_dirt = 2 * T4.dirt + 5;
    SpanGuard span(p, s, "conv|decoder.final_dense|fused");
    CDAN_TRY(fused_fd_launch(*p->fd_fused, b.T4, 8, x, y, N, H, W, s));
    p->launches += 1;
    fill_stages(p, ld1, ld2, ld3, 0);
  } else {
    // The final dense block's concat buffer is group-planar on the tensor-core path, NHWC otherwise.
    const bool fd_planar = dt == kBF16 && p->conv_impl == 0 && fd_planar_enabled();
    const int fdl = fd_planar ? 16 : fd_ld();
    CDAN_TRY(need_up(T4));
    { SpanGuard span(p, s, "glue|up_add_input"); CDAN_TRY(up_add_input_launch(dt, b.T4, 8, x, b.FD, fdl, 16, N, H, W, s)); }
    p->launches += 1;
    DenseTens tf;
    dense_tens(tf, b.FD, 16, fdl, 0, H, W, fd_planar, nullptr, 0, 2 * T4.dirt + 1);
    // final DenseBlock(3,3,16,4) + sigmoid, written straight to the caller's fp32 NCHW output
    if (fd_planar) CDAN_TRY(run_dense_planar(p, FDL0, N, H, W, b.FD, s, y, dense_hook(tf)));
    else CDAN_TRY(run_dense(p, FDL0, N, H, W, b.FD, fd_ld(), 16, nullptr, 3, s, y, dense_hook(tf)));
    out_dirt = tf.out.dirt;
    fill_stages(p, ld1, ld2, ld3, fdl);
  }
  if (br && out_dirt > br->D(0)) return fail("row-tiled forward: internal error, the output's dirty zone reaches the owned rows");
  return 0;
}

int forward_band_impl(cdan_plan* p, cudaStream_t s, const float* x_ext, float* y_ext, int N, int H, int W, int halo) {
  if (!p->band_comm) return fail("cdan_forward_band: no band transport attached (cdan_plan_band_attach_local / _nccl)");
  int rows[4];
  CDAN_TRY(band_rows(H, p->band_comm->nranks, p->band_comm->rank, halo, rows));
  BandRun br{p, s, N, rows[3] - rows[2], W, halo, rows[0] - rows[2], rows[3] - rows[1], H};
  p->band_stats = BandStats{};
  return forward_impl(p, s, x_ext, y_ext, N, br.Hext, W, &br);
}

}  // namespace
}  // namespace cdan

using namespace cdan;

extern "C" {

const char* cdan_last_error(void) { return g_error.c_str(); }
const char* cdan_version(void) { return "cdan_b200 0.1 sm_100a"; }

int cdan_plan_create(int device, int dtype, cdan_plan** plan_out) {
  if (!plan_out) return fail("cdan_plan_create: plan_out is NULL");
  if (dtype != CDAN_DTYPE_F32 && dtype != CDAN_DTYPE_BF16) return fail("cdan_plan_create: unknown dtype");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(std::string("cdan_plan_create: no CUDA device available (") + cudaGetErrorString(e) +
                "); this library has no CPU fallback");
  if (device < 0 || device >= count) return fail("cdan_plan_create: device index out of range");
  cudaDeviceProp prop;
  CDAN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(std::string("cdan_plan_create: built for sm_100a (B200) only, device is sm_") +
                std::to_string(prop.major) + std::to_string(prop.minor));
  cdan_plan* p = new cdan_plan();
  p->device = device;
  p->dt = DType(dtype);
  *plan_out = p;
  return 0;
}

int cdan_plan_destroy(cdan_plan* p) {
  if (!p) return 0;
  DeviceGuard g(p->device);
  free_weights(p);
  delete p->band_comm;
  if (p->ws) cudaFree(p->ws);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_h2d[i]) cudaEventDestroy(p->ev_h2d[i]);
    if (p->ev_comp[i]) cudaEventDestroy(p->ev_comp[i]);
    if (p->ev_d2h[i]) cudaEventDestroy(p->ev_d2h[i]);
  }
  if (p->host_stage) cudaFree(p->host_stage);
  delete p;
  return 0;
}

int cdan_plan_set_option(cdan_plan* p, const char* name, int value) {
  if (!p || !name) return fail("cdan_plan_set_option: NULL argument");
  if (!strcmp(name, "profile")) {
    p->profile = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "host_chunk")) {
    if (value < 0) return fail("host_chunk must be >= 0 (0 = auto)");
    p->host_chunk = value;
    return 0;
  }
  if (!strcmp(name, "fd_fused")) {
    p->fd_fused_on = value != 0;
    return 0;
  }
  if (!strcmp(name, "conv_impl")) {
    if (value < 0 || value > 1) return fail("conv_impl must be 0 (auto) or 1 (CUDA cores)");
    p->conv_impl = value;
    return 0;
  }
  return fail(std::string("unknown option '") + name + "'");
}

int cdan_plan_load_weights(cdan_plan* p, int n, const char* const* keys, const void* const* ptrs,
                           const int64_t* numels) {
  if (!p || !keys || !ptrs || !numels) return fail("cdan_plan_load_weights: NULL argument");
  DeviceGuard g(p->device);
  HostDict sd;
  for (int i = 0; i < n; ++i) {
    const std::string key = keys[i];
    const std::string tail = "num_batches_tracked";
    if (key.size() >= tail.size() && key.compare(key.size() - tail.size(), tail.size(), tail) == 0) continue;
    std::vector<float> h(static_cast<size_t>(numels[i]), 0.f);
    CDAN_CUDA_OK(cudaMemcpy(h.data(), ptrs[i], h.size() * sizeof(float), cudaMemcpyDefault));
    sd.emplace(key, std::move(h));
  }
  free_weights(p);
  int rc = 0;
  const int enc_c[5] = {3, 64, 128, 256, 512};
  for (int i = 1; i <= 4 && !rc; ++i)
    rc = load_conv_block(p, sd, ConvId(ENC1 + i - 1), "encoder.conv" + std::to_string(i), enc_c[i - 1], enc_c[i]);
  if (!rc) rc = load_dense_block(p, sd, D1L0, "encoder.dense1", 64, 64, 64);
  if (!rc) rc = load_dense_block(p, sd, D2L0, "encoder.dense2", 128, 128, 128);
  if (!rc) rc = load_dense_block(p, sd, D3L0, "encoder.dense3", 256, 256, 256);
  if (!rc) rc = load_dense_block(p, sd, FDL0, "decoder.final_dense", 3, 16, 3);
  const int dec_c[5] = {512, 256, 128, 64, 3};
  for (int i = 1; i <= 4 && !rc; ++i) rc = load_decoder_conv(p, sd, ConvId(DEC1 + i - 1), i, dec_c[i - 1], dec_c[i]);
  if (!rc) rc = load_cbam(p, sd, 0, "bottleneck", 512);
  if (!rc) rc = load_cbam(p, sd, 1, "decoder.cbam1", 256);
  if (!rc) rc = load_cbam(p, sd, 2, "decoder.cbam2", 128);
  if (!rc) rc = load_cbam(p, sd, 3, "decoder.cbam3", 64);
  if (!rc && p->dt == kBF16) {
    FusedFdLayer fl[5];
    for (int i = 0; i < 5; ++i) {
      const ConvLayer& L = p->conv[FDL0 + i];
      fl[i].Cin = L.Cin; fl[i].CoutP = L.CoutP;
      fl[i].w = L.h_w.data(); fl[i].bias = L.h_bias.data(); fl[i].pre_s = L.h_pre_s.data(); fl[i].pre_t = L.h_pre_t.data();
    }
    rc = fused_fd_pack_create(fl, &p->fd_fused);
  }
  if (rc) {
    free_weights(p);
    return rc;
  }
  p->loaded = true;
  return 0;
}

int cdan_workspace_bytes(cdan_plan* p, int N, int H, int W, size_t* bytes_out) {
  if (!p || !bytes_out) return fail("cdan_workspace_bytes: NULL argument");
  if (N <= 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return fail("cdan_workspace_bytes: H and W must be positive multiples of 8");
  Buffers b{};
  *bytes_out = carve(b, nullptr, p->dt, N, H, W, use_fd_fused(p));
  return 0;
}

int cdan_forward(cdan_plan* p, void* stream, const float* x, float* y, int N, int H, int W) {
  if (!p || !x || !y) return fail("cdan_forward: NULL argument");
  DeviceGuard g(p->device);
  return forward_impl(p, (cudaStream_t)stream, x, y, N, H, W);
}

// Host-buffer pipeline shared by cdan_forward_host (fp32 NCHW buffers) and cdan_forward_host_u8 (uint8 NHWC buffers).
static int forward_host_impl(cdan_plan* p, const void* x_host, void* y_host, int N, int H, int W, bool u8) {
  DeviceGuard g(p->device);
  if (H % 8 || W % 8 || N <= 0) return fail("cdan_forward_host: H and W must be multiples of 8");
  // Sub-batch pipeline over three streams: the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the forward
  // of chunk k (PCIe is full duplex), with double-buffered device staging.
  if (!p->own_stream) {
    CDAN_CUDA_OK(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    CDAN_CUDA_OK(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
    CDAN_CUDA_OK(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CDAN_CUDA_OK(cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming));
      CDAN_CUDA_OK(cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
      CDAN_CUDA_OK(cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  // images per full pipeline step: option "host_chunk", or (0 = auto) as many images as make up 16 x 1080p worth of
  // pixels, so that small images are not run as launch-bound forwards of a handful of images
  int chunk = p->host_chunk;
  if (chunk <= 0) chunk = int(std::max<size_t>(1, (size_t(16) * 1080 * 1920) / (size_t(H) * W)));
  if (chunk >= N && N >= 8) chunk = (N + 1) / 2;  // keep at least two steps so that copies overlap compute
  const int cb = std::max(1, std::min(N, chunk));
  const size_t img_floats = size_t(3) * H * W, slot_floats = size_t(cb) * img_floats;
  // device staging per slot: fp32 x | fp32 y | (u8 path) u8 x | u8 y
  const size_t slot_bytes = 2 * slot_floats * sizeof(float) + (u8 ? 2 * align_up(slot_floats, 256) : 0);
  if (p->host_stage_bytes < 2 * slot_bytes) {
    CDAN_CUDA_OK(cudaDeviceSynchronize());
    if (p->host_stage) CDAN_CUDA_OK(cudaFree(p->host_stage));
    p->host_stage = nullptr;
    p->host_stage_bytes = 0;
    CDAN_CUDA_OK(cudaMalloc(&p->host_stage, 2 * slot_bytes));
    p->host_stage_bytes = 2 * slot_bytes;
  }
  CDAN_TRY(ensure_workspace(p, cb, H, W));
  // Chunk schedule.  Three engines run concurrently (H2D of chunk k+1, forward of chunk k, D2H of chunk k-1) on double-buffered
  // staging; what stays exposed is the first H2D, the last D2H, any copy that does not fit behind its neighbouring forward, and
  // the lower efficiency of small forwards (~0.45 ms per extra chunk at 1080p).  A symmetric ramp solves that: start with one
  // 1080p-equivalent of pixels (two on the uint8 path), grow by the ratio of forward time to copy time per image — about 2-3 for
  // fp32 buffers (0.57 ms of PCIe per 1080p image and direction against 1.05 ms of compute), 6 for uint8 — up to the middle,
  // then mirror it.  32 x 1080p: fp32 [1,3,6,12,6,3,1] (measured 38.4 ms against 39.2 ms for [2,6,16,6,2]; device-resident forward 33.7 ms),
  // uint8 [2,14,14,2] (36.2 ms).
  // Results do not depend on the schedule (the forward is batch-independent, bitwise).
  std::vector<int> sched;
  {
    const int unit = int(std::max<size_t>(1, (size_t(1080) * 1920) / (size_t(H) * W)));  // images per 1080p of pixels
    const int first = (u8 ? 2 : 1) * unit;
    std::vector<int> ramp;
    int sum = 0;
    if (p->host_chunk <= 0 || p->host_chunk >= 4) {
      for (int c = first, k = 0; c < cb && 2 * (sum + c) < N - c; ++k) {
        ramp.push_back(c);
        sum += c;
        c *= u8 ? 6 : (k == 0 ? 3 : 2);
      }
    }
    if (!ramp.empty()) {
      sched = ramp;
      int middle = N - 2 * sum;
      const int parts = (middle + cb - 1) / cb;
      for (int i = 0; i < parts; ++i) {
        const int c = (middle + (parts - i) - 1) / (parts - i);
        sched.push_back(c);
        middle -= c;
      }
      sched.insert(sched.end(), ramp.rbegin(), ramp.rend());
    } else {
      for (int rest = N; rest > 0; rest -= cb) sched.push_back(std::min(cb, rest));
    }
  }
  const size_t img_host_bytes = u8 ? img_floats : img_floats * sizeof(float);
  int k = 0, n0 = 0;
  for (size_t ci = 0; ci < sched.size(); n0 += sched[ci], ++ci, ++k) {
    const int nb = sched[ci], slot = k & 1;
    char* base = reinterpret_cast<char*>(p->host_stage) + size_t(slot) * slot_bytes;
    float* xs = reinterpret_cast<float*>(base);
    float* ys = xs + slot_floats;
    uint8_t* xu = reinterpret_cast<uint8_t*>(ys + slot_floats);
    uint8_t* yu = xu + align_up(slot_floats, 256);
    const size_t bytes = size_t(nb) * img_host_bytes;
    if (k >= 2) CDAN_CUDA_OK(cudaStreamWaitEvent(p->h2d_stream, p->ev_comp[slot], 0));  // x slot free again
    CDAN_CUDA_OK(cudaMemcpyAsync(u8 ? (void*)xu : (void*)xs, (const char*)x_host + size_t(n0) * img_host_bytes, bytes,
                                 cudaMemcpyHostToDevice, p->h2d_stream));
    CDAN_CUDA_OK(cudaEventRecord(p->ev_h2d[slot], p->h2d_stream));
    CDAN_CUDA_OK(cudaStreamWaitEvent(p->own_stream, p->ev_h2d[slot], 0));
    if (k >= 2) CDAN_CUDA_OK(cudaStreamWaitEvent(p->own_stream, p->ev_d2h[slot], 0));      // y slot drained
    if (u8) CDAN_TRY(normalize_u8_launch(xu, xs, nb, H, W, p->own_stream));
    CDAN_TRY(forward_impl(p, p->own_stream, xs, ys, nb, H, W));
    if (u8) CDAN_TRY(quantize_u8_launch(ys, yu, nb, H, W, p->own_stream));
    CDAN_CUDA_OK(cudaEventRecord(p->ev_comp[slot], p->own_stream));
    CDAN_CUDA_OK(cudaStreamWaitEvent(p->d2h_stream, p->ev_comp[slot], 0));
    CDAN_CUDA_OK(cudaMemcpyAsync((char*)y_host + size_t(n0) * img_host_bytes, u8 ? (const void*)yu : (const void*)ys, bytes,
                                 cudaMemcpyDeviceToHost, p->d2h_stream));
    CDAN_CUDA_OK(cudaEventRecord(p->ev_d2h[slot], p->d2h_stream));
  }
  CDAN_CUDA_OK(cudaStreamSynchronize(p->d2h_stream));
  return 0;
}

int cdan_forward_band(cdan_plan* p, void* stream, const float* x_ext, float* y_ext, int N, int H, int W, int halo) {
  if (!p || !x_ext || !y_ext) return fail("cdan_forward_band: NULL argument");
  DeviceGuard g(p->device);
  return forward_band_impl(p, (cudaStream_t)stream, x_ext, y_ext, N, H, W, halo);
}

int cdan_forward_host(cdan_plan* p, const float* x_host, float* y_host, int N, int H, int W) {
  if (!p || !x_host || !y_host) return fail("cdan_forward_host: NULL argument");
  return forward_host_impl(p, x_host, y_host, N, H, W, false);
}

int cdan_forward_host_u8(cdan_plan* p, const unsigned char* x_host, unsigned char* y_host, int N, int H, int W) {
  if (!p || !x_host || !y_host) return fail("cdan_forward_host_u8: NULL argument");
  return forward_host_impl(p, x_host, y_host, N, H, W, true);
}

int cdan_stage_read(cdan_plan* p, void* stream, const char* name, float* dst, int64_t shape_out[4]) {
  if (!p || !name) return fail("cdan_stage_read: NULL argument");
  DeviceGuard g(p->device);
  auto it = p->stages.find(name);
  if (it == p->stages.end()) return fail(std::string("cdan_stage_read: unknown stage '") + name + "' (run cdan_forward first)");
  const Stage& st = it->second;
  if (shape_out) {
    shape_out[0] = p->N; shape_out[1] = st.C; shape_out[2] = st.h; shape_out[3] = st.w;
  }
  if (!dst) return 0;
  return nhwc_to_nchw_launch(p->dt, st.p, st.ld, dst, p->N, st.C, st.h, st.w, (cudaStream_t)stream);
}

int cdan_last_launch_count(cdan_plan* p) { return p ? p->launches : 0; }

int cdan_profile_read(cdan_plan* p, char* buf, size_t buflen) {
  if (!p || !buf || buflen == 0) return fail("cdan_profile_read: NULL argument");
  DeviceGuard g(p->device);
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (auto& sp : p->spans) {
    CDAN_CUDA_OK(cudaEventSynchronize(sp.e1));
    float ms = 0.f;
    CDAN_CUDA_OK(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
    if (!acc.count(sp.label)) order.push_back(sp.label);
    acc[sp.label].first += ms;
    acc[sp.label].second += 1;
    p->event_pool.push_back(sp.e0);
    p->event_pool.push_back(sp.e1);
  }
  p->spans.clear();
  std::string out;
  for (auto& k : order) out += k + " " + std::to_string(acc[k].first) + " " + std::to_string(acc[k].second) + "\n";
  if (out.size() + 1 > buflen) return fail("cdan_profile_read: buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

}  // extern "C"
