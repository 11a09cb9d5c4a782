// C ABI of libdiartb200.so (include/diart_b200.h): handles, weight preparation, workspaces and the
// launch sequences of the two networks, the clustering step and the fused pipeline step.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <map>
#include <memory>
#include <sstream>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/diart_b200.h"
#include "dg_common.cuh"

namespace dg {

static thread_local std::string g_err;
std::atomic<long long> g_launches{0};
thread_local int g_sm_limit = 0;   // per host thread: distinct handles driven by distinct threads stay independent
void set_error(const std::string& msg) { g_err = msg; }

struct ProfRec {
  std::string name;
  cudaEvent_t a, b;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
// DG_TRACE_LAUNCHES=1: every scope prints "dg-trace <tag> <first launch ordinal> <one past the last>" on stderr, so that a
// profiler's launch list (tools/ncu_summary.py) can be labelled with the tags bench.py reports
static const bool g_trace = getenv("DG_TRACE_LAUNCHES") && getenv("DG_TRACE_LAUNCHES")[0] == '1';
ProfScope::ProfScope(const char* name_, cudaStream_t st_) : on(g_prof), st(st_), a(nullptr), b(nullptr), name(name_) {
  if (g_trace) first = g_launches.load();
  if (!on) return;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
}
ProfScope::~ProfScope() {
  if (g_trace) fprintf(stderr, "dg-trace %s %lld %lld\n", name, first, (long long)g_launches.load());
  if (!on) return;
  cudaEventRecord(b, st);
  g_recs.push_back({name, a, b});
}

int launch_stats_pool_ex(const float* x, int stride, int T, int C, const float* w, int F, int K, int layout,
                         int n_groups, const int* grp_item, const int* grp_q0, const int* grp_nq, const int* idx0,
                         const int* idx1, const float* lam1, float eps, float* pooled, cudaStream_t st,
                         long long item_pitch = 0, int row_pitch = 0);

// ------------------------------------------------------------------------------ small utilities
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  float wscale = 1.f;   // weight hi planes: the power-of-two factor both planes were multiplied by (upload_split)
  int ensure(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    // (re)allocation is rare (first step at a given batch size).  The handles drive several non-blocking
    // streams, which do not order against the legacy stream this memset runs on: drain the device on both sides.
    DG_CUDA(cudaDeviceSynchronize());
    DG_CUDA(cudaMalloc(&p, n));
    DG_CUDA(cudaMemset(p, 0, n));
    DG_CUDA(cudaDeviceSynchronize());
    bytes = n;
    return 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

// A model handle owns ONE set of activation buffers per scratch lane.  A new user of a lane -- another pipeline built on the same
// handle, or a block-level call on another stream -- first waits, stream-ordered, for the previous user's last kernel; without it
// two users in flight would silently overwrite each other's activations.  (Host threads: a handle is single-threaded.)
struct UseGuard {
  cudaEvent_t e = nullptr;
  const void* owner = nullptr;
  ~UseGuard() {
    if (e) cudaEventDestroy(e);
  }
};
static thread_local bool g_in_pipeline = false;      // the fused pipeline brackets its own uses (per lane)
static int use_begin(UseGuard& u, const void* owner, cudaStream_t st) {
  if (u.e && u.owner != owner) DG_CUDA(cudaStreamWaitEvent(st, u.e, 0));
  return 0;
}
static int use_end(UseGuard& u, const void* owner, cudaStream_t st) {
  if (!u.e) DG_CUDA(cudaEventCreateWithFlags(&u.e, cudaEventDisableTiming));
  DG_CUDA(cudaEventRecord(u.e, st));
  u.owner = owner;
  return 0;
}

struct Tensors {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  Tensors(const dg_tensor* t, int n) {
    for (int i = 0; i < n; i++)
      if (t[i].name) m[t[i].name] = {t[i].data, t[i].numel};
  }
  const float* get(const std::string& name, int64_t numel) const {
    auto it = m.find(name);
    if (it == m.end()) {
      set_error("missing tensor '" + name + "' in state dict");
      return nullptr;
    }
    if (it->second.second != numel || !it->second.first) {
      set_error("tensor '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
                std::to_string(numel));
      return nullptr;
    }
    return it->second.first;
  }
  int64_t numel(const std::string& name) const {
    auto it = m.find(name);
    return it == m.end() ? -1 : it->second.second;
  }
};

// DG_SIMT=1 forces the float32 SIMT GEMMs everywhere (A/B switch for the parity tests and bench)
static bool use_tensor_cores() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DG_SIMT");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static int upload_u16(DevBuf& b, const std::vector<uint16_t>& h) {
  if (b.ensure(h.size() * 2)) return -2;
  DG_CUDA(cudaMemcpy(b.p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  return 0;
}

// float32 [N][K] host weights -> zero-padded 16-bit hi/lo device planes [Npad][K]
static int upload_split(DevBuf& hi, DevBuf& lo, const std::vector<float>& w_nk, int N, int Npad, int K) {
  std::vector<uint16_t> h((size_t)Npad * K), l((size_t)Npad * K);
  hi.wscale = lo.wscale = weight_plane_scale(w_nk.data(), (size_t)N * K, split_f16());
  split_weights_host(w_nk.data(), N, Npad, K, h.data(), l.data(), split_f16(), hi.wscale);
  return (upload_u16(hi, h) || upload_u16(lo, l)) ? DG_ECUDA : 0;
}

static int upload(DevBuf& b, const std::vector<float>& h) {
  if (b.ensure(h.size() * sizeof(float))) return -2;
  DG_CUDA(cudaMemcpy(b.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------ SincNet front end
struct SincWeights {
  float wn_gamma = 1.f, wn_beta = 0.f;
  DevBuf filt, g0, b0, w1, bias1, g1, b1, w2, bias2, g2, b2;
  DevBuf w1_hi, w1_lo, w2_hi, w2_lo;   // tcgen05 path: 16-bit hi/lo planes [64][448] (taps folded into K: 5 x 80 + pad) and [64][5*64]
  DevBuf filt_planes;                  // sinc filter bank as 16-bit planes [3][80][256] (hi, lo; lo2 for the bf16 mode)
  DevBuf cf;                           // folded wav-norm affine: beta * sum_k h[f][k]
  DevBuf hsum;                         // sum_k h[f][k] (stream form of the sinc layer)
};

// ParamSincFB.filters() in float32, as asteroid-filterbanks computes it with torch
static void sinc_filters(const float* low_hz_, const float* band_hz_, std::vector<float>& filt /*[251][80]*/) {
  filt.assign(251 * 80, 0.f);
  float n_[125], win[125];
  for (int i = 0; i < 125; i++) {
    const float t = (float)(i - 125) / 16000.0f;
    n_[i] = 6.283185307179586f * t;
    win[i] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * i / 250.0));
  }
  for (int f = 0; f < 40; f++) {
    const float low = 50.f + fabsf(low_hz_[f]);
    float high = low + 50.f + fabsf(band_hz_[f]);
    high = fminf(fmaxf(high, 50.f), 8000.f);
    const float band = high - low, two_band = 2.f * band;
    for (int i = 0; i < 125; i++) {
      const float ft_low = low * n_[i], ft_high = high * n_[i], half_n = n_[i] / 2.f;
      const float lc = ((sinf(ft_high) - sinf(ft_low)) / half_n) * win[i];
      const float ls = ((cosf(ft_low) - cosf(ft_high)) / half_n) * win[i];
      filt[i * 80 + f] = lc / two_band;
      filt[(250 - i) * 80 + f] = lc / two_band;
      filt[i * 80 + 40 + f] = ls / two_band;
      filt[(250 - i) * 80 + 40 + f] = (-ls) / two_band;
    }
    filt[125 * 80 + f] = two_band / two_band;
    filt[125 * 80 + 40 + f] = 0.f / two_band;
  }
}

static int prep_sincnet(const Tensors& t, const std::string& pre, SincWeights& w) {
  const float *g, *b;
  if (!(g = t.get(pre + "wav_norm1d.weight", 1)) || !(b = t.get(pre + "wav_norm1d.bias", 1))) return DG_EWEIGHT;
  w.wn_gamma = g[0];
  w.wn_beta = b[0];
  const float* lo = t.get(pre + "conv1d.0.filterbank.low_hz_", 40);
  const float* bd = t.get(pre + "conv1d.0.filterbank.band_hz_", 40);
  if (!lo || !bd) return DG_EWEIGHT;
  std::vector<float> h;
  sinc_filters(lo, bd, h);
  if (upload(w.filt, h)) return DG_ECUDA;
  {
    std::vector<uint16_t> fp(3 * 80 * 256);
    sinc_tc_pack_filters(h.data(), fp.data(), split_f16());
    if (upload_u16(w.filt_planes, fp)) return DG_ECUDA;
    std::vector<float> cf(80);
    sinc_tc_affine_consts(h.data(), w.wn_beta, cf.data());
    if (upload(w.cf, cf)) return DG_ECUDA;
    std::vector<float> hs(80);
    sinc_tc_affine_consts(h.data(), 1.f, hs.data());
    if (upload(w.hsum, hs)) return DG_ECUDA;
  }
  auto pad_vec = [&](const std::string& name, int n, int npad, DevBuf& dst) -> int {
    const float* s = t.get(name, n);
    if (!s) return DG_EWEIGHT;
    std::vector<float> v(npad, 0.f);
    memcpy(v.data(), s, n * sizeof(float));
    return upload(dst, v) ? DG_ECUDA : 0;
  };
  int rc;
  if ((rc = pad_vec(pre + "norm1d.0.weight", 80, 80, w.g0)) || (rc = pad_vec(pre + "norm1d.0.bias", 80, 80, w.b0)) ||
      (rc = pad_vec(pre + "norm1d.1.weight", 60, 64, w.g1)) || (rc = pad_vec(pre + "norm1d.1.bias", 60, 64, w.b1)) ||
      (rc = pad_vec(pre + "norm1d.2.weight", 60, 64, w.g2)) || (rc = pad_vec(pre + "norm1d.2.bias", 60, 64, w.b2)) ||
      (rc = pad_vec(pre + "conv1d.1.bias", 60, 64, w.bias1)) || (rc = pad_vec(pre + "conv1d.2.bias", 60, 64, w.bias2)))
    return rc;
  // Conv1d weights [out][in][k] -> shifted-window GEMM layout [(tap*Cin_pad + c)][out_pad]
  auto conv_w = [&](const std::string& name, int out, int in, int k, int in_pad, int out_pad, DevBuf& dst) -> int {
    const float* s = t.get(name, (int64_t)out * in * k);
    if (!s) return DG_EWEIGHT;
    std::vector<float> v((size_t)k * in_pad * out_pad, 0.f);
    for (int o = 0; o < out; o++)
      for (int c = 0; c < in; c++)
        for (int j = 0; j < k; j++) v[((size_t)j * in_pad + c) * out_pad + o] = s[((size_t)o * in + c) * k + j];
    return upload(dst, v) ? DG_ECUDA : 0;
  };
  if ((rc = conv_w(pre + "conv1d.1.weight", 60, 80, 5, 80, 64, w.w1)) ||
      (rc = conv_w(pre + "conv1d.2.weight", 60, 60, 5, 64, 64, w.w2)))
    return rc;
  auto conv_w_tc = [&](const std::string& name, int out, int in, int k, int in_pad, DevBuf& hi, DevBuf& lo) -> int {
    const float* s = t.get(name, (int64_t)out * in * k);
    if (!s) return DG_EWEIGHT;
    std::vector<float> w_nk((size_t)out * k * in_pad, 0.f);
    for (int o = 0; o < out; o++)
      for (int c = 0; c < in; c++)
        for (int j = 0; j < k; j++) w_nk[(size_t)o * k * in_pad + j * in_pad + c] = s[((size_t)o * in + c) * k + j];
    return upload_split(hi, lo, w_nk, out, 64, k * in_pad);
  };
  {
    // Conv1d(80, 60, 5) with its taps folded into K: the input planes are 80-channel rows (pitch 160 B), so the im2col row of
    // output row m is the 400 CONTIGUOUS values starting at row m -- read through an overlapping-row TMA view, K = 448
    const float* s1 = t.get(pre + "conv1d.1.weight", (int64_t)60 * 80 * 5);
    if (!s1) return DG_EWEIGHT;
    std::vector<float> w_nk((size_t)60 * 448, 0.f);
    for (int o = 0; o < 60; o++)
      for (int c = 0; c < 80; c++)
        for (int j = 0; j < 5; j++) w_nk[(size_t)o * 448 + j * 80 + c] = s1[((size_t)o * 80 + c) * 5 + j];
    if (upload_split(w.w1_hi, w.w1_lo, w_nk, 60, 64, 448)) return DG_ECUDA;
  }
  if ((rc = conv_w_tc(pre + "conv1d.2.weight", 60, 60, 5, 64, w.w2_hi, w.w2_lo))) return rc;
  return 0;
}

// waveform statistics and the standardised-waveform planes of a batch; both networks' SincNets read the same
// ones, so the fused pipeline computes them once per step
struct SincPrep {
  DevBuf wmean, wrstd, wh, wl;
  // stream form (needs a hop hint; DG_STREAM_SINC=0 disables it): planes of the raw stream and the device flag "this batch is a run of
  // overlapping windows"; `hop` > 0 means the stream-form launches were enqueued for this batch
  DevBuf swh, swl, flag, spart;
  int hop = 0;
  int ensure(int B, const Geom& g) {
    const size_t bytes = 4 * sinc_tc_plane_elems(B, g) * 2;
    return (wmean.ensure(B * 4) || wrstd.ensure(B * 4) || wh.ensure(bytes) || wl.ensure(bytes)) ? DG_ECUDA : 0;
  }
  int ensure_stream(int B, const Geom& g, int hop_) {
    const size_t bytes = 4 * sinc_stream_geom(B, g, hop_).plane * 2;
    return (swh.ensure(bytes) || swl.ensure(bytes) || flag.ensure(16)) ? DG_ECUDA : 0;
  }
};
struct SincWork {
  DevBuf wmean, wrstd, p0, sc0, sh0, p1, sc1, sh1, p2, sc2, sh2;
  DevBuf a0h, a0l, c1, a1h, a1l, c2;   // tcgen05 path: 16-bit planes of the conv inputs, un-pooled conv outputs
  DevBuf craw, part;                   // stream form: raw convolution of the stream [P][80], statistics partials
  DevBuf part3;                        // per-tile InstanceNorm partial sums of the pooling GEMM epilogues (conv1, conv2)
  SincPrep own_prep;                   // statistics + waveform planes when no shared ones are supplied
  const float* out = nullptr;          // conv2 output that the next layer normalises on load ...
  int out_pool = 0;                    // ... 1: still un-pooled (rows = 3x), MaxPool1d(3) is applied on load
  int ensure_tc(int B, const Geom& g) {
    const size_t tail = 64;
    if (a0h.ensure(((size_t)B * g.S0 + tail) * 128 * 2) || a0l.ensure(((size_t)B * g.S0 + tail) * 128 * 2) ||
        c1.ensure(((size_t)B * g.S0 + tail) * 64 * 4) || a1h.ensure(((size_t)B * g.S1 + tail) * 64 * 2) ||
        a1l.ensure(((size_t)B * g.S1 + tail) * 64 * 2) || c2.ensure(((size_t)B * g.S1 + tail) * 64 * 4))
      return DG_ECUDA;
    return 0;
  }
  int ensure(int B, const Geom& g) {
    const size_t tail = 64;  // spare rows so shifted windows of the last tile stay in bounds
    if (wmean.ensure(B * 4) || wrstd.ensure(B * 4) || p0.ensure(((size_t)B * g.S0 + tail) * 80 * 4) ||
        sc0.ensure((size_t)B * 80 * 4) || sh0.ensure((size_t)B * 80 * 4) ||
        p1.ensure(((size_t)B * g.S1 + tail) * 64 * 4) || sc1.ensure((size_t)B * 64 * 4) ||
        sh1.ensure((size_t)B * 64 * 4) || p2.ensure(((size_t)B * g.S2 + tail) * 64 * 4) ||
        sc2.ensure((size_t)B * 64 * 4) || sh2.ensure((size_t)B * 64 * 4))
      return DG_ECUDA;
    return 0;
  }
};

static int run_sinc_prep(SincPrep& p, const float* wav, int B, const Geom& g, cudaStream_t st, int hop = 0,
                         bool overlap_known = false) {
  int rc;
  if ((rc = p.ensure(B, g))) return rc;
  // stream form of the sinc layer: on by default (DG_STREAM_SINC=0 disables it), only with a hop hint from the caller;
  // the device flag written by overlap_check decides per batch, so a wrong hint costs a few empty launches, never a
  // wrong result
  static const bool stream_on = !(getenv("DG_STREAM_SINC") && getenv("DG_STREAM_SINC")[0] == '0');
  p.hop = 0;
  if (stream_on && hop > 0 && B >= 4 && hop % 40 == 0 && g.S % 4 == 0 && hop < g.S && ((uintptr_t)wav & 15) == 0) {
    if ((rc = p.ensure_stream(B, g, hop))) return rc;
    if (overlap_known) {   // the batch was formed on the device from ONE stream (dg_stream): nothing to verify
      DG_CUDA(cudaMemsetAsync(p.flag.p, 1, sizeof(int), st));
    } else if ((rc = launch_overlap_check(wav, B, g.S, hop, p.flag.as<int>(), st))) {
      return rc;
    }
    p.hop = hop;
  }
  const bool fast_stats = p.hop && stream_stats_ok(g.S, p.hop);
  if (fast_stats) {
    if (p.spart.ensure(stream_stats_doubles(B, g.S, p.hop) * 8)) return DG_ECUDA;
    if ((rc = launch_stream_stats(wav, B, g.S, p.hop, p.spart.as<double>(), p.wmean.as<float>(), p.wrstd.as<float>(),
                                  p.flag.as<int>(), st)))
      return rc;
  }
  if ((rc = launch_wave_stats(wav, B, g.S, p.wmean.as<float>(), p.wrstd.as<float>(), st, fast_stats ? p.flag.as<int>() : nullptr)))
    return rc;
  if (p.hop && (rc = launch_stream_prep(wav, B, g, hop, p.swh.p, p.swl.p, p.flag.as<int>(), st))) return rc;
  return launch_sinc_prep(wav, p.wmean.as<float>(), p.wrstd.as<float>(), B, g, p.wh.p, p.wl.p, st,
                          p.hop ? p.flag.as<int>() : nullptr);
}

// waveform [B,S] -> k.out (pre-norm conv2 output, pooled [B*S2,64] or un-pooled [B*S1,64]) + its
// InstanceNorm scale/shift (k.sc2, k.sh2)
static int run_sincnet(const SincWeights& w, SincWork& k, const float* wav, int B, const Geom& g, cudaStream_t st,
                       const SincPrep* shared = nullptr) {
  int rc;
  if ((rc = k.ensure(B, g))) return rc;
  if (use_tensor_cores()) {
    if ((rc = k.ensure_tc(B, g))) return rc;
    const int* stream_flag = nullptr;      // device flag "the stream form produced the conv1 operand planes of this batch"
    static const bool sinc_simt = getenv("DG_SINC_SIMT") && getenv("DG_SINC_SIMT")[0] == '1';
    if (sinc_simt) {
      if ((rc = launch_wave_stats(wav, B, g.S, k.wmean.as<float>(), k.wrstd.as<float>(), st))) return rc;
      rc = launch_sinc0(wav, k.wmean.as<float>(), k.wrstd.as<float>(), w.wn_gamma, w.wn_beta, w.filt.as<float>(), B,
                        g, k.p0.as<float>(), st);
    } else {
      const SincPrep* prep = shared;
      if (!prep) {
        if ((rc = run_sinc_prep(k.own_prep, wav, B, g, st))) return rc;
        prep = &k.own_prep;
      }
      if (prep->hop) {   // stream form: one convolution of the unique samples + a per-window affine / |.| / pool pass
        const SincStreamGeom sg = sinc_stream_geom(B, g, prep->hop);
        if (k.craw.ensure(((size_t)sg.P + 16) * 80 * 4) || k.part.ensure(sinc_pool_part_floats(B, g, prep->hop) * 4)) return DG_ECUDA;
        // raw convolution of the stream, then statistics and normalised operand planes straight from it (p0 is never written)
        if ((rc = launch_sinc0_tc_stream(w.filt_planes.p, B, g, prep->hop, prep->swh.p, prep->swl.p, k.craw.as<float>(),
                                         prep->flag.as<int>(), st)) ||
            (rc = launch_sinc_pool_fused(k.craw.as<float>(), prep->wmean.as<float>(), prep->wrstd.as<float>(), w.cf.as<float>(),
                                         w.hsum.as<float>(), w.wn_gamma, B, g, prep->hop, w.g0.as<float>(), w.b0.as<float>(),
                                         k.part.as<float>(), k.sc0.as<float>(), k.sh0.as<float>(), k.a0h.p, k.a0l.p,
                                         prep->flag.as<int>(), st)))
          return rc;
        stream_flag = prep->flag.as<int>();
      }
      rc = launch_sinc0_tc(w.wn_gamma, w.cf.as<float>(), w.filt_planes.p, B, g, prep->wh.p, prep->wl.p,
                           k.p0.as<float>(), st, prep->hop ? prep->flag.as<int>() : nullptr);
    }
    if (rc) return rc;
    if ((rc = launch_instnorm_stats(k.p0.as<float>(), B, g.S0, g.T0, 80, 80, w.g0.as<float>(), w.b0.as<float>(),
                                    k.sc0.as<float>(), k.sh0.as<float>(), st, 0, stream_flag)))
      return rc;
    // Conv1d(80,60,5): normalised input as 16-bit hi/lo planes (80-channel rows), un-pooled float32 output
    const long long M0 = (long long)B * g.S0, M1 = (long long)B * g.S1;
    if ((rc = launch_split_ex(k.p0.as<float>(), M0, 80, 80, 80, 0, g.S0, k.sc0.as<float>(), k.sh0.as<float>(),
                              k.a0h.p, k.a0l.p, st, stream_flag)))
      return rc;
    // conv1 / conv2 with MaxPool1d(3) and the InstanceNorm partial sums in the GEMM epilogue (TC_MAXPOOL3): the un-pooled maps are
    // never written, the statistics pass reads 2 x 2 x 64 floats per tile.  Needs a tile of 96..126 rows that divides the item at both stages;
    // DG_NO_POOL3_FUSE=1 = the round-1 path (un-pooled float32 map -> instnorm_stats -> split with pooling on load)
    static const bool pool3_on = !(getenv("DG_NO_POOL3_FUSE") && getenv("DG_NO_POOL3_FUSE")[0] == '1');
    const int tr0 = gemm_tc_pool3_tile_rows(g.S0), tr1 = gemm_tc_pool3_tile_rows(g.S1);
    if (pool3_on && tr0 && tr1) {
      if (k.part3.ensure((size_t)(M0 / tr0) * 2 * 2 * 64 * 4)) return DG_ECUDA;
      TcGemm t{};
      t.A_hi = k.a0h.p; t.A_lo = k.a0l.p; t.lda = 80; t.Cin = 448; t.KW = 1; t.dil = 1; t.Mtot = M0; t.M = M0;
      t.W_hi = w.w1_hi.p; t.w_scale = w.w1_hi.wscale; t.W_lo = w.w1_lo.p; t.Npad = 64; t.N = 64; t.bias = w.bias1.as<float>();
      t.out_f32 = k.p1.as<float>(); t.ldc = 64; t.epi = 5; t.tag = "sinc_conv1";
      t.pool_part = k.part3.as<float>(); t.pool_item_rows = g.S0; t.pool3_T = g.T1; t.pool3_tile_rows = tr0;
      if ((rc = launch_gemm_tc(t, st)) ||
          (rc = launch_instnorm_finalize(k.part3.as<float>(), B, g.S0, tr0, g.T1, 64, 64, w.bias1.as<float>(), w.g1.as<float>(),
                                         w.b1.as<float>(), k.sc1.as<float>(), k.sh1.as<float>(), 64, st)) ||
          (rc = launch_split_ex(k.p1.as<float>(), M1, 64, 64, 64, 0, g.S1, k.sc1.as<float>(), k.sh1.as<float>(), k.a1h.p, k.a1l.p, st)))
        return rc;
      t.A_hi = k.a1h.p; t.A_lo = k.a1l.p; t.lda = 64; t.Cin = 64; t.KW = 5; t.Mtot = M1; t.M = M1;
      t.W_hi = w.w2_hi.p; t.w_scale = w.w2_hi.wscale; t.W_lo = w.w2_lo.p; t.bias = w.bias2.as<float>();
      t.out_f32 = k.p2.as<float>(); t.tag = "sinc_conv2";
      t.pool_item_rows = g.S1; t.pool3_T = g.T2; t.pool3_tile_rows = tr1;
      if ((rc = launch_gemm_tc(t, st))) return rc;
      k.out = k.p2.as<float>();
      k.out_pool = 0;
      return launch_instnorm_finalize(k.part3.as<float>(), B, g.S1, tr1, g.T2, 64, 64, w.bias2.as<float>(), w.g2.as<float>(),
                                      w.b2.as<float>(), k.sc2.as<float>(), k.sh2.as<float>(), 64, st);
    }
    TcGemm t{};
    t.A_hi = k.a0h.p; t.A_lo = k.a0l.p; t.lda = 80; t.Cin = 448; t.KW = 1; t.dil = 1; t.Mtot = M0; t.M = M0;
    t.W_hi = w.w1_hi.p; t.w_scale = w.w1_hi.wscale; t.W_lo = w.w1_lo.p; t.Npad = 64; t.N = 64; t.bias = w.bias1.as<float>();
    t.out_f32 = k.c1.as<float>(); t.ldc = 64; t.epi = 0; t.tag = "sinc_conv1";
    if ((rc = launch_gemm_tc(t, st))) return rc;
    if ((rc = launch_instnorm_stats(k.c1.as<float>(), B, g.S0, g.T1, 64, 64, w.g1.as<float>(), w.b1.as<float>(),
                                    k.sc1.as<float>(), k.sh1.as<float>(), st, 1)))
      return rc;
    // Conv1d(60,60,5) on MaxPool(conv1) -> norm -> leaky, again un-pooled output
    if ((rc = launch_split_ex(k.c1.as<float>(), M1, 64, 64, 64, 1, g.S1, k.sc1.as<float>(), k.sh1.as<float>(),
                              k.a1h.p, k.a1l.p, st)))
      return rc;
    t.A_hi = k.a1h.p; t.A_lo = k.a1l.p; t.lda = 64; t.Cin = 64; t.KW = 5; t.Mtot = M1; t.M = M1;
    t.W_hi = w.w2_hi.p; t.w_scale = w.w2_hi.wscale; t.W_lo = w.w2_lo.p; t.bias = w.bias2.as<float>();
    t.out_f32 = k.c2.as<float>(); t.tag = "sinc_conv2";
    if ((rc = launch_gemm_tc(t, st))) return rc;
    k.out = k.c2.as<float>();
    k.out_pool = 1;
    return launch_instnorm_stats(k.c2.as<float>(), B, g.S1, g.T2, 64, 64, w.g2.as<float>(), w.b2.as<float>(),
                                 k.sc2.as<float>(), k.sh2.as<float>(), st, 1);
  }
  k.out = k.p2.as<float>();
  k.out_pool = 0;
  if ((rc = launch_wave_stats(wav, B, g.S, k.wmean.as<float>(), k.wrstd.as<float>(), st))) return rc;
  if ((rc = launch_sinc0(wav, k.wmean.as<float>(), k.wrstd.as<float>(), w.wn_gamma, w.wn_beta, w.filt.as<float>(), B,
                         g, k.p0.as<float>(), st)))
    return rc;
  if ((rc = launch_instnorm_stats(k.p0.as<float>(), B, g.S0, g.T0, 80, 80, w.g0.as<float>(), w.b0.as<float>(),
                                  k.sc0.as<float>(), k.sh0.as<float>(), st)))
    return rc;
  GemmArgs a{};
  a.A = k.p0.as<float>(); a.lda = 80; a.Cin = 80; a.KW = 5; a.dil = 1;
  a.Mtot = (long long)B * g.S0; a.M = (long long)B * g.S0;
  a.W = w.w1.as<float>(); a.ldw = 64; a.N = 64; a.bias = w.bias1.as<float>();
  a.in_sc = k.sc0.as<float>(); a.in_sh = k.sh0.as<float>(); a.item_rows = g.S0;
  a.C = k.p1.as<float>(); a.ldc = 64; a.epi = EPI_BIAS_POOL3; a.tag = "sinc_conv1";
  if ((rc = launch_gemm(a, st))) return rc;
  if ((rc = launch_instnorm_stats(k.p1.as<float>(), B, g.S1, g.T1, 64, 64, w.g1.as<float>(), w.b1.as<float>(),
                                  k.sc1.as<float>(), k.sh1.as<float>(), st)))
    return rc;
  a.A = k.p1.as<float>(); a.lda = 64; a.Cin = 64;
  a.Mtot = (long long)B * g.S1; a.M = (long long)B * g.S1;
  a.W = w.w2.as<float>(); a.bias = w.bias2.as<float>();
  a.in_sc = k.sc1.as<float>(); a.in_sh = k.sh1.as<float>(); a.item_rows = g.S1;
  a.C = k.p2.as<float>(); a.tag = "sinc_conv2";
  if ((rc = launch_gemm(a, st))) return rc;
  return launch_instnorm_stats(k.p2.as<float>(), B, g.S2, g.T2, 64, 64, w.g2.as<float>(), w.b2.as<float>(),
                               k.sc2.as<float>(), k.sh2.as<float>(), st);
}

}  // namespace dg

using namespace dg;

// ================================================================================== segmentation
struct dg_seg {
  int device = 0, K = 3;           // K = classifier outputs (local speakers; powerset classes for powerset models)
  int ps_speakers = 0;             // > 0: powerset model with this many local speakers (dg_seg_set_powerset)
  DevBuf ps_masks;                 // speaker bit set of every powerset class
  SincWeights sw;
  DevBuf wih[4], bih[4], whh[4];   // input projections [in_pad][1024], bias [1024], packed W_hh
  DevBuf wih_hi[4], wih_lo[4];     // the same as 16-bit hi/lo planes [1024][in_pad] for the tcgen05 path
  DevBuf whh_hi[4], whh_lo[4];     // W_hh as 16-bit hi/lo planes [2][512][128] for the tcgen05 recurrence
  DevBuf l1w, l1b, l2w, l2b, cw, cb;
  DevBuf l1_hi, l1_lo, l2_hi, l2_lo, ones128, zeros128;   // head Linears as 16-bit hi/lo planes [128][in] (tcgen05 path)
  // activations: two independent sets ("lanes") so that the fused pipeline can run the segmentation chains of
  // two consecutive steps concurrently (the recurrence occupies only 32 SMs)
  struct Scratch {
    SincWork work;
    DevBuf gx, hA, hB, y1, y2;
    DevBuf xh, xl;                 // bf16 hi/lo planes of the current in-projection input
    DevBuf y1h, y1l;               // bf16 planes of the first head Linear's output
  } scr[2];
  int lane = 0;
  const SincPrep* shared_prep = nullptr;   // set by the fused pipeline: statistics + planes computed once per step
  UseGuard guard[2];                       // per scratch lane
};

static int seg_prepare(dg_seg* h, const Tensors& t) {
  int rc;
  if ((rc = prep_sincnet(t, "sincnet.", h->sw))) return rc;
  for (int L = 0; L < 4; L++) {
    const int in = L == 0 ? 60 : 256, in_pad = L == 0 ? 64 : 256;
    std::vector<float> w((size_t)in_pad * 1024, 0.f), b(1024, 0.f), packed(lstm_whh_packed_floats());
    const float* hh[2];
    for (int d = 0; d < 2; d++) {
      const std::string sfx = "_l" + std::to_string(L) + (d ? "_reverse" : "");
      const float* wi = t.get("lstm.weight_ih" + sfx, (int64_t)512 * in);
      const float* bi = t.get("lstm.bias_ih" + sfx, 512);
      const float* bh = t.get("lstm.bias_hh" + sfx, 512);
      hh[d] = t.get("lstm.weight_hh" + sfx, 512 * 128);
      if (!wi || !bi || !bh || !hh[d]) return DG_EWEIGHT;
      for (int r = 0; r < 512; r++) {
        for (int c = 0; c < in; c++) w[(size_t)c * 1024 + d * 512 + r] = wi[(size_t)r * in + c];
        b[d * 512 + r] = bi[r] + bh[r];
      }
    }
    lstm_pack_whh(hh[0], hh[1], packed.data());
    if (upload(h->wih[L], w) || upload(h->bih[L], b) || upload(h->whh[L], packed)) return DG_ECUDA;
    std::vector<float> w_nk((size_t)1024 * in_pad, 0.f);
    for (int n = 0; n < 1024; n++)
      for (int c = 0; c < in; c++) w_nk[(size_t)n * in_pad + c] = w[(size_t)c * 1024 + n];
    if (upload_split(h->wih_hi[L], h->wih_lo[L], w_nk, 1024, 1024, in_pad)) return DG_ECUDA;
    {
      std::vector<uint16_t> rh(lstm_tc_plane_elems()), rl(lstm_tc_plane_elems());
      h->whh_hi[L].wscale = lstm_tc_pack_whh(hh[0], hh[1], rh.data(), rl.data(), split_f16());
      if (upload_u16(h->whh_hi[L], rh) || upload_u16(h->whh_lo[L], rl)) return DG_ECUDA;
    }
  }
  auto linear_t = [&](const std::string& name, int out, int in, DevBuf& dw, DevBuf& db) -> int {
    const float* w = t.get(name + ".weight", (int64_t)out * in);
    const float* b = t.get(name + ".bias", out);
    if (!w || !b) return DG_EWEIGHT;
    std::vector<float> wt((size_t)in * out), bv(b, b + out);
    for (int o = 0; o < out; o++)
      for (int c = 0; c < in; c++) wt[(size_t)c * out + o] = w[(size_t)o * in + c];
    return (upload(dw, wt) || upload(db, bv)) ? DG_ECUDA : 0;
  };
  if ((rc = linear_t("linear.0", 128, 256, h->l1w, h->l1b)) || (rc = linear_t("linear.1", 128, 128, h->l2w, h->l2b)))
    return rc;
  {
    const float* w0 = t.get("linear.0.weight", 128 * 256);
    const float* w1 = t.get("linear.1.weight", 128 * 128);
    if (!w0 || !w1) return DG_EWEIGHT;
    if (upload_split(h->l1_hi, h->l1_lo, std::vector<float>(w0, w0 + 128 * 256), 128, 128, 256) ||
        upload_split(h->l2_hi, h->l2_lo, std::vector<float>(w1, w1 + 128 * 128), 128, 128, 128) ||
        upload(h->ones128, std::vector<float>(128, 1.f)) || upload(h->zeros128, std::vector<float>(128, 0.f)))
      return DG_ECUDA;
  }
  const int64_t cn = t.numel("classifier.bias");
  if (cn < 1 || cn > 8) {
    set_error("classifier.bias missing or more than 8 local speakers");
    return DG_EWEIGHT;
  }
  h->K = (int)cn;
  const float* cw = t.get("classifier.weight", cn * 128);
  const float* cb = t.get("classifier.bias", cn);
  if (!cw || !cb) return DG_EWEIGHT;
  if (upload(h->cw, std::vector<float>(cw, cw + cn * 128)) || upload(h->cb, std::vector<float>(cb, cb + cn)))
    return DG_ECUDA;
  return 0;
}

extern "C" const char* dg_last_error(void) { return g_err.c_str(); }
extern "C" int dg_version(void) { return 100; }
extern "C" int64_t dg_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int dg_profile_enable(int enable) {
  g_prof = enable != 0;
  return DG_OK;
}

// JSON {"name": {"count": n, "ms": total}, ...} of everything recorded since the last report.
// Synchronises the device.  Returns the number of bytes written (excluding the terminator).
extern "C" int dg_profile_report(char* buf, int cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    auto& e = agg[r.name];
    e.first++;
    e.second += ms;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  std::ostringstream os;
  os << "{";
  bool first = true;
  for (auto& kv : agg) {
    if (!first) os << ", ";
    first = false;
    os << "\"" << kv.first << "\": {\"count\": " << kv.second.first << ", \"ms\": " << kv.second.second << "}";
  }
  os << "}";
  const std::string s = os.str();
  if (!buf || cap <= (int)s.size()) {
    set_error("dg_profile_report: buffer too small");
    return DG_EINVAL;
  }
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

extern "C" int dg_seg_create(const dg_tensor* tensors, int n, int device, dg_seg** out) {
  if (!tensors || !out) {
    set_error("dg_seg_create: null argument");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(device));
  std::unique_ptr<dg_seg> h(new dg_seg());
  h->device = device;
  Tensors t(tensors, n);
  int rc = seg_prepare(h.get(), t);
  if (rc) return rc;
  *out = h.release();
  return DG_OK;
}

extern "C" int dg_seg_dims(const dg_seg* h, int num_samples, int* frames, int* speakers) {
  if (!h || num_samples < 3000) {
    set_error("dg_seg_dims: bad arguments");
    return DG_EINVAL;
  }
  Geom g = make_geom(num_samples);
  if (frames) *frames = g.T2;
  if (speakers) *speakers = h->ps_speakers ? h->ps_speakers : h->K;
  return DG_OK;
}

// Declares the model a powerset model (pyannote/segmentation-3.0 style): its classifier has one output per subset of
// the `num_speakers` local speakers of size <= `max_per_frame`, in itertools.combinations order (pyannote
// Powerset.build_mapping); the forward then returns hard multilabel scores (reference models.py:29-39).
extern "C" int dg_seg_set_powerset(dg_seg* h, int num_speakers, int max_per_frame) {
  if (!h || num_speakers < 1 || num_speakers > 8 || max_per_frame < 0 || max_per_frame > num_speakers) {
    set_error("dg_seg_set_powerset: bad arguments");
    return DG_EINVAL;
  }
  std::vector<uint32_t> masks;
  for (int size = 0; size <= max_per_frame; size++)          // subsets by size, each size in lexicographic order
    for (uint32_t m = 0; m < (1u << num_speakers); m++) {
      if (__builtin_popcount(m) != size) continue;
      masks.push_back(m);
    }
  // lexicographic order of combinations (0,1) < (0,2) < (1,2) is NOT numeric order of the bit masks in general: sort each
  // size class by the sorted member tuples
  auto members = [&](uint32_t m) {
    std::vector<int> v;
    for (int i = 0; i < num_speakers; i++)
      if (m >> i & 1u) v.push_back(i);
    return v;
  };
  std::stable_sort(masks.begin(), masks.end(), [&](uint32_t a, uint32_t b) {
    const int sa = __builtin_popcount(a), sb = __builtin_popcount(b);
    if (sa != sb) return sa < sb;
    return members(a) < members(b);
  });
  if ((int)masks.size() != h->K) {
    set_error("dg_seg_set_powerset: the classifier has " + std::to_string(h->K) + " outputs but the powerset has " +
              std::to_string(masks.size()) + " classes");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  if (h->ps_masks.ensure(masks.size() * 4)) return DG_ECUDA;
  DG_CUDA(cudaMemcpy(h->ps_masks.p, masks.data(), masks.size() * 4, cudaMemcpyHostToDevice));
  h->ps_speakers = num_speakers;
  return DG_OK;
}

// classifier + sigmoid, or classifier + powerset decoding
static int seg_head_final(dg_seg* h, const float* y2, int B, const Geom& g, float* seg, cudaStream_t st) {
  if (h->ps_speakers)
    return launch_seg_powerset(y2, h->cw.as<float>(), h->cb.as<float>(), B, g.T2, g.S2, h->K, h->ps_speakers,
                               h->ps_masks.as<unsigned>(), seg, st);
  return launch_seg_final(y2, h->cw.as<float>(), h->cb.as<float>(), B, g.T2, g.S2, h->K, seg, st);
}

static int seg_forward_impl(dg_seg* h, const float* wav, int B, int S, float* seg, void* stream);
extern "C" int dg_seg_forward(dg_seg* h, const float* wav, int B, int S, float* seg, void* stream) {
  if (!h || !wav || !seg || B < 1 || S < 3000) {
    set_error("dg_seg_forward: bad arguments (need B >= 1, S >= 3000)");
    return DG_EINVAL;
  }
  if (g_in_pipeline) return seg_forward_impl(h, wav, B, S, seg, stream);
  DG_CUDA(cudaSetDevice(h->device));
  UseGuard& u = h->guard[h->lane & 1];
  int rc = use_begin(u, stream ? stream : (void*)h, (cudaStream_t)stream);
  if (!rc) rc = seg_forward_impl(h, wav, B, S, seg, stream);
  if (!rc) rc = use_end(u, stream ? stream : (void*)h, (cudaStream_t)stream);
  return rc;
}

static int seg_forward_impl(dg_seg* h, const float* wav, int B, int S, float* seg, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_CUDA(cudaSetDevice(h->device));
  dg_seg::Scratch& w = h->scr[h->lane & 1];
  const Geom g = make_geom(S);
  int rc;
  if ((rc = run_sincnet(h->sw, w.work, wav, B, g, st, h->shared_prep))) return rc;
  const size_t rows = (size_t)B * g.S2 + 64;
  if (w.gx.ensure(rows * 1024 * 4) || w.hA.ensure(rows * 256 * 4) || w.hB.ensure(rows * 256 * 4) ||
      w.y1.ensure(rows * 128 * 4) || w.y2.ensure(rows * 128 * 4))
    return DG_ECUDA;
  const long long M = (long long)B * g.S2;
  float* hin = nullptr;
  float* hbuf[2] = {w.hA.as<float>(), w.hB.as<float>()};
  const bool tc = use_tensor_cores();
  if (tc && (w.xh.ensure(rows * 256 * 2) || w.xl.ensure(rows * 256 * 2))) return DG_ECUDA;
  for (int L = 0; L < 4; L++) {
    if (tc) {
      const int cin = L == 0 ? 64 : 256;
      static const bool lstm_simt = getenv("DG_LSTM_SIMT") && getenv("DG_LSTM_SIMT")[0] == '1';
      if (L == 0)
        rc = launch_split_ex(w.work.out, M, 64, 64, 64, w.work.out_pool, g.S2, w.work.sc2.as<float>(),
                             w.work.sh2.as<float>(), w.xh.p, w.xl.p, st);
      else if (lstm_simt)
        rc = launch_split(hin, M, 256, g.S2, nullptr, nullptr, w.xh.p, w.xl.p, st);
      if (rc) return rc;
      TcGemm t{};
      t.A_hi = w.xh.p; t.A_lo = w.xl.p; t.lda = cin; t.Cin = cin; t.KW = 1; t.dil = 1; t.Mtot = M; t.M = M;
      t.W_hi = h->wih_hi[L].p; t.w_scale = h->wih_hi[L].wscale; t.W_lo = h->wih_lo[L].p; t.Npad = 1024; t.N = 1024; t.bias = h->bih[L].as<float>();
      t.out_f32 = w.gx.as<float>(); t.ldc = 1024; t.epi = 0; t.tag = "lstm_inproj";
      if ((rc = launch_gemm_tc(t, st))) return rc;
      float* hout = hbuf[L & 1];
      // the tcgen05 recurrence writes h_t straight into the operand planes of the next GEMM (the in-projection that
      // read them has completed in stream order); the SIMT recurrence (DG_LSTM_SIMT=1) goes through float32 + split
      if (lstm_simt)
        rc = launch_lstm_layer(w.gx.as<float>(), h->whh[L].as<float>(), B, g.T2, g.S2, hout, st);
      else
        rc = launch_lstm_layer_tc(w.gx.as<float>(), h->whh_hi[L].p, h->whh_lo[L].p, h->whh_hi[L].wscale, B, g.T2, g.S2, nullptr, w.xh.p, w.xl.p, st);
      if (rc) return rc;
      hin = hout;
      continue;
    }
    GemmArgs a{};
    if (L == 0) {
      a.A = w.work.p2.as<float>(); a.lda = 64; a.Cin = 64;
      a.in_sc = w.work.sc2.as<float>(); a.in_sh = w.work.sh2.as<float>(); a.item_rows = g.S2;
    } else {
      a.A = hin; a.lda = 256; a.Cin = 256;
    }
    a.KW = 1; a.dil = 1; a.Mtot = M; a.M = M;
    a.W = h->wih[L].as<float>(); a.ldw = 1024; a.N = 1024; a.bias = h->bih[L].as<float>();
    a.C = w.gx.as<float>(); a.ldc = 1024; a.epi = EPI_BIAS; a.tag = "lstm_inproj";
    if ((rc = launch_gemm(a, st))) return rc;
    float* hout = hbuf[L & 1];
    if ((rc = launch_lstm_layer(w.gx.as<float>(), h->whh[L].as<float>(), B, g.T2, g.S2, hout, st))) return rc;
    hin = hout;
  }
  if (tc) {
    // Linear(256,128) -> leaky -> Linear(128,128) -> leaky on the tcgen05 GEMM (identity "BatchNorm")
    if (w.y1h.ensure(rows * 128 * 2) || w.y1l.ensure(rows * 128 * 2)) return DG_ECUDA;
    static const bool lstm_simt2 = getenv("DG_LSTM_SIMT") && getenv("DG_LSTM_SIMT")[0] == '1';
    if (lstm_simt2 && (rc = launch_split(hin, M, 256, g.S2, nullptr, nullptr, w.xh.p, w.xl.p, st))) return rc;
    TcGemm t{};
    t.A_hi = w.xh.p; t.A_lo = w.xl.p; t.lda = 256; t.Cin = 256; t.KW = 1; t.dil = 1; t.Mtot = M; t.M = M;
    t.W_hi = h->l1_hi.p; t.w_scale = h->l1_hi.wscale; t.W_lo = h->l1_lo.p; t.Npad = 128; t.N = 128; t.bias = h->l1b.as<float>();
    t.bn_scale = h->ones128.as<float>(); t.bn_shift = h->zeros128.as<float>();
    t.out_hi = w.y1h.p; t.out_lo = w.y1l.p; t.ldc = 128; t.epi = 1; t.tag = "seg_linear";
    if ((rc = launch_gemm_tc(t, st))) return rc;
    t.A_hi = w.y1h.p; t.A_lo = w.y1l.p; t.lda = 128; t.Cin = 128;
    t.W_hi = h->l2_hi.p; t.w_scale = h->l2_hi.wscale; t.W_lo = h->l2_lo.p; t.bias = h->l2b.as<float>();
    t.out_hi = nullptr; t.out_lo = nullptr; t.out_f32 = w.y2.as<float>(); t.epi = 2;
    if ((rc = launch_gemm_tc(t, st))) return rc;
    return seg_head_final(h, w.y2.as<float>(), B, g, seg, st);
  }
  GemmArgs a{};
  a.A = hin; a.lda = 256; a.Cin = 256; a.KW = 1; a.dil = 1; a.Mtot = M; a.M = M;
  a.W = h->l1w.as<float>(); a.ldw = 128; a.N = 128; a.bias = h->l1b.as<float>();
  a.C = w.y1.as<float>(); a.ldc = 128; a.epi = EPI_BIAS_LEAKY; a.tag = "seg_linear";
  if ((rc = launch_gemm(a, st))) return rc;
  a.A = w.y1.as<float>(); a.lda = 128; a.Cin = 128;
  a.W = h->l2w.as<float>(); a.bias = h->l2b.as<float>(); a.C = w.y2.as<float>();
  if ((rc = launch_gemm(a, st))) return rc;
  return seg_head_final(h, w.y2.as<float>(), B, g, seg, st);
}

extern "C" int dg_seg_destroy(dg_seg* h) {
  delete h;
  return DG_OK;
}

// ===================================================================================== embedding
struct dg_emb {
  int device = 0, pool_mode = 31, D = 512;
  SincWeights sw;
  DevBuf tw[5], tb[5], bns[5], bnh[5];
  DevBuf tw_hi[5], tw_lo[5];             // bf16 hi/lo planes [Npad][K] for the tcgen05 path
  DevBuf ew_hi, ew_lo, ph, pl;           // Linear(3000, D): weights [Dpad][3008], pooled statistics planes
  DevBuf xh, xl, aH, aL, bH, bL;         // bf16 hi/lo activation planes
  DevBuf ew, eb;
  SincWork work;
  UseGuard guard;
  DevBuf tA, tB, t5, pooled, eraw;
  DevBuf idx0, idx1, lam1;
  int tab_F = -1, tab_T = -1;
  DevBuf flags, uniq, grp, gathered;   // compatibility path
  const SincPrep* shared_prep = nullptr;
  // what the pooling reads after a trunk pass: x(item, t, c) = pool_x[item * pool_item_pitch + t * pool_row_pitch + c], c < pool_C
  const float* pool_x = nullptr;
  long long pool_item_pitch = 0;
  int pool_row_pitch = 0, pool_C = 1500;
  const void *t4h = nullptr, *t4l = nullptr;   // operand planes of TDNN5 after a trunk pass that stopped before it
  DevBuf pool_rw, pool_vs, pool_part;           // fused TDNN5 + pooling: row weights, weight sums, per-tile partial sums
  int variant = 0;                     // 0: XVectorSincNet (pyannote/embedding), 1: WeSpeaker ResNet34 (variant B)
  std::unique_ptr<struct ResNet> rn;
};

// ---- variant B: WeSpeaker ResNet34 (SURVEY.md 8(a) A8'; kernels in resnet.cu + the Conv2d epilogue of gemm_tc.cu)
struct ResConv {                       // Conv2d (3x3 pad 1 or 1x1, no bias) + folded BatchNorm2d(eval)
  int cin = 0, cout = 0, ksize = 3, stride = 1;
  int KW = 9, cin_gemm = 0, lda = 0;   // GEMM view: taps, channels consumed per tap, row pitch of the input planes
  DevBuf w_hi, w_lo, sc, sh;
};
struct ResBlock {
  ResConv c1, c2, sc;
  bool has_sc = false;
};
struct ResNet {
  DevBuf fb_hi, fb_lo, banks, k_lo, k_hi, stem_w, stem_sc, stem_sh;
  std::vector<ResBlock> blocks;
  int stage_of[16];
  // work buffers: planes of the waveform, spectrum, log-mel map, three plane pairs per stage, float32 final map
  DevBuf wav_hi, wav_lo, spec, logmel, mean, act[4][3][2], fin;
  int last_S = 0;                // the padding rings are only valid for one geometry: buffers are cleared when it changes
  int stop_after = 99;           // test hook (dg_emb_debug_trunk): stop after the stem (-1) / after block k
  int dbg_stage = 0, dbg_buf = 0;
};
static const int RN_CH[4] = {32, 64, 128, 256};
static const int RN_BLOCKS[4] = {3, 4, 6, 3};

static const int TD_OUT[5] = {512, 512, 512, 512, 1500};
static const int TD_K[5] = {5, 3, 3, 1, 1};
static const int TD_DIL[5] = {1, 2, 3, 1, 1};

static int resnet_prepare(dg_emb* h, const Tensors& t);

static int emb_prepare(dg_emb* h, const Tensors& t) {
  int rc;
  if (t.numel("resnet.conv1.weight") > 0) return resnet_prepare(h, t);     // variant B checkpoint
  if ((rc = prep_sincnet(t, "sincnet.", h->sw))) return rc;
  int in = 60, in_pad = 64;
  for (int L = 0; L < 5; L++) {
    const int out = TD_OUT[L], k = TD_K[L];
    const std::string cv = "tdnns." + std::to_string(3 * L), bn = "tdnns." + std::to_string(3 * L + 2);
    const float* w = t.get(cv + ".weight", (int64_t)out * in * k);
    const float* b = t.get(cv + ".bias", out);
    const float* gm = t.get(bn + ".weight", out);
    const float* bt = t.get(bn + ".bias", out);
    const float* rm = t.get(bn + ".running_mean", out);
    const float* rv = t.get(bn + ".running_var", out);
    if (!w || !b || !gm || !bt || !rm || !rv) return DG_EWEIGHT;
    std::vector<float> wt((size_t)k * in_pad * out, 0.f), bv(b, b + out), sc(out), sf(out);
    for (int o = 0; o < out; o++) {
      for (int c = 0; c < in; c++)
        for (int j = 0; j < k; j++) wt[((size_t)j * in_pad + c) * out + o] = w[((size_t)o * in + c) * k + j];
      // BatchNorm1d(eval): (x - mean) / sqrt(var + 1e-5) * gamma + beta  ==  x * sc + sf
      sc[o] = gm[o] / sqrtf(rv[o] + 1e-5f);
      sf[o] = bt[o] - rm[o] * sc[o];
    }
    if (upload(h->tw[L], wt) || upload(h->tb[L], bv) || upload(h->bns[L], sc) || upload(h->bnh[L], sf)) return DG_ECUDA;
    {
      const int K = k * in_pad, npad = (out + 255) / 256 * 256;
      std::vector<float> w_nk((size_t)out * K, 0.f);
      for (int o = 0; o < out; o++)
        for (int c = 0; c < in; c++)
          for (int j = 0; j < k; j++) w_nk[(size_t)o * K + j * in_pad + c] = w[((size_t)o * in + c) * k + j];
      if (upload_split(h->tw_hi[L], h->tw_lo[L], w_nk, out, npad, K)) return DG_ECUDA;
    }
    in = out;
    in_pad = out;
  }
  const int64_t dn = t.numel("embedding.bias");
  if (dn < 4 || dn % 4) {
    set_error("embedding.bias missing or dimension not a multiple of 4");
    return DG_EWEIGHT;
  }
  h->D = (int)dn;
  const float* ew = t.get("embedding.weight", dn * 3000);
  const float* eb = t.get("embedding.bias", dn);
  if (!ew || !eb) return DG_EWEIGHT;
  std::vector<float> wt((size_t)3000 * dn);
  for (int o = 0; o < dn; o++)
    for (int c = 0; c < 3000; c++) wt[(size_t)c * dn + o] = ew[(size_t)o * 3000 + c];
  if (upload(h->ew, wt) || upload(h->eb, std::vector<float>(eb, eb + dn))) return DG_ECUDA;
  {
    std::vector<float> w_nk((size_t)dn * 3008, 0.f);
    for (int o = 0; o < dn; o++) memcpy(&w_nk[(size_t)o * 3008], ew + (size_t)o * 3000, 3000 * sizeof(float));
    if (upload_split(h->ew_hi, h->ew_lo, w_nk, (int)dn, ((int)dn + 255) / 256 * 256, 3008)) return DG_ECUDA;
  }
  return 0;
}

// Conv2d weight [co][ci][kh (mel)][kw (time)] + BatchNorm2d -> GEMM weight planes [Npad][K] (tap-major K) + scale / shift.
// Maps are [item][w = time][h = mel][C]: tap (dw, dh) multiplies w[co][ci][dh][dw].  With 32 input channels the three dh
// taps of one dw are 96 CONTIGUOUS values of the input planes (rows h-1, h, h+1 follow each other in memory), so they are
// read as one 128-wide K slab through an overlapping-row view (row pitch 32): 3 taps x 128 instead of 9 taps x 64.
static int resnet_conv_prepare(const Tensors& t, const std::string& conv, const std::string& bn, int cin, int cout, int ksize,
                               int stride, ResConv& c) {
  const float* w = t.get(conv + ".weight", (int64_t)cout * cin * ksize * ksize);
  const float* gm = t.get(bn + ".weight", cout);
  const float* bt = t.get(bn + ".bias", cout);
  const float* rm = t.get(bn + ".running_mean", cout);
  const float* rv = t.get(bn + ".running_var", cout);
  if (!w || !gm || !bt || !rm || !rv) return DG_EWEIGHT;
  c.cin = cin; c.cout = cout; c.ksize = ksize; c.stride = stride;
  const bool narrow = cin == 32;
  if (ksize == 3) {
    c.KW = narrow ? 3 : 9;
    c.cin_gemm = narrow ? 128 : cin;
  } else {
    c.KW = 1;
    c.cin_gemm = narrow ? 64 : cin;
  }
  c.lda = cin;
  const int K = c.KW * c.cin_gemm;
  const int npad = cout <= 64 ? cout : (cout + 127) / 128 * 128;
  std::vector<float> w_nk((size_t)cout * K, 0.f), sc(cout), sh(cout);
  for (int o = 0; o < cout; o++) {
    for (int ci = 0; ci < cin; ci++)
      for (int dh = 0; dh < ksize; dh++)
        for (int dw = 0; dw < ksize; dw++) {
          const float v = w[(((size_t)o * cin + ci) * ksize + dh) * ksize + dw];
          size_t k;
          if (ksize == 1) k = ci;
          else if (narrow) k = (size_t)dw * 128 + dh * 32 + ci;
          else k = (size_t)(dw * 3 + dh) * cin + ci;
          w_nk[(size_t)o * K + k] = v;
        }
    sc[o] = gm[o] / sqrtf(rv[o] + 1e-5f);
    sh[o] = bt[o] - rm[o] * sc[o];
  }
  if (upload_split(c.w_hi, c.w_lo, w_nk, cout, npad, K) || upload(c.sc, sc) || upload(c.sh, sh)) return DG_ECUDA;
  return 0;
}

static int resnet_prepare(dg_emb* h, const Tensors& t) {
  int rc;
  h->variant = 1;
  h->rn.reset(new ResNet());
  ResNet& r = *h->rn;
  {
    std::vector<float> op;
    fbank_frame_operator(op);                                   // [514][400]
    std::vector<float> w_nk((size_t)514 * 448, 0.f);
    for (int n = 0; n < 514; n++) memcpy(&w_nk[(size_t)n * 448], &op[(size_t)n * 400], 400 * sizeof(float));
    if (upload_split(r.fb_hi, r.fb_lo, w_nk, 514, 640, 448)) return DG_ECUDA;
    std::vector<float> banks;
    std::vector<int> lo, hi;
    fbank_mel_banks(banks, lo, hi);
    if (upload(r.banks, banks) || r.k_lo.ensure(80 * 4) || r.k_hi.ensure(80 * 4)) return DG_ECUDA;
    DG_CUDA(cudaMemcpy(r.k_lo.p, lo.data(), 80 * 4, cudaMemcpyHostToDevice));
    DG_CUDA(cudaMemcpy(r.k_hi.p, hi.data(), 80 * 4, cudaMemcpyHostToDevice));
  }
  {
    const float* w = t.get("resnet.conv1.weight", 32 * 9);
    const float* gm = t.get("resnet.bn1.weight", 32);
    const float* bt = t.get("resnet.bn1.bias", 32);
    const float* rm = t.get("resnet.bn1.running_mean", 32);
    const float* rv = t.get("resnet.bn1.running_var", 32);
    if (!w || !gm || !bt || !rm || !rv) return DG_EWEIGHT;
    std::vector<float> sc(32), sh(32);
    for (int o = 0; o < 32; o++) {
      sc[o] = gm[o] / sqrtf(rv[o] + 1e-5f);
      sh[o] = bt[o] - rm[o] * sc[o];
    }
    if (upload(r.stem_w, std::vector<float>(w, w + 288)) || upload(r.stem_sc, sc) || upload(r.stem_sh, sh)) return DG_ECUDA;
  }
  int in_planes = 32, bi = 0;
  r.blocks.resize(16);
  for (int st = 0; st < 4; st++)
    for (int b = 0; b < RN_BLOCKS[st]; b++, bi++) {
      const int planes = RN_CH[st], stride = (b == 0 && st > 0) ? 2 : 1;
      const std::string pre = "resnet.layer" + std::to_string(st + 1) + "." + std::to_string(b) + ".";
      ResBlock& blk = r.blocks[bi];
      r.stage_of[bi] = st;
      if ((rc = resnet_conv_prepare(t, pre + "conv1", pre + "bn1", in_planes, planes, 3, stride, blk.c1)) ||
          (rc = resnet_conv_prepare(t, pre + "conv2", pre + "bn2", planes, planes, 3, 1, blk.c2)))
        return rc;
      blk.has_sc = stride != 1 || in_planes != planes;
      if (blk.has_sc && (rc = resnet_conv_prepare(t, pre + "shortcut.0", pre + "shortcut.1", in_planes, planes, 1, stride, blk.sc)))
        return rc;
      in_planes = planes;
    }
  // Linear(5120, D): pyannote's feature order is (channel, mel) -- "batch (dimension channel) frames" -- ours (mel, channel)
  const int64_t dn = t.numel("resnet.seg_1.bias");
  if (dn < 4 || dn % 4) {
    set_error("resnet.seg_1.bias missing or dimension not a multiple of 4");
    return DG_EWEIGHT;
  }
  h->D = (int)dn;
  const float* ew = t.get("resnet.seg_1.weight", dn * 5120);
  const float* eb = t.get("resnet.seg_1.bias", dn);
  if (!ew || !eb) return DG_EWEIGHT;
  std::vector<float> w_nk((size_t)dn * 5120);
  for (int o = 0; o < dn; o++)
    for (int half = 0; half < 2; half++)
      for (int hh = 0; hh < 10; hh++)
        for (int c = 0; c < 256; c++) w_nk[(size_t)o * 5120 + half * 2560 + hh * 256 + c] = ew[(size_t)o * 5120 + half * 2560 + c * 10 + hh];
  if (upload_split(h->ew_hi, h->ew_lo, w_nk, (int)dn, ((int)dn + 255) / 256 * 256, 5120) ||
      upload(h->eb, std::vector<float>(eb, eb + dn)))
    return DG_ECUDA;
  h->pool_C = 2560;
  return 0;
}

// geometry of variant B for S samples: fbank frames and the four map sizes (time x mel)
struct ResGeom {
  int T0, W[4], H[4];
};
static int resnet_geom(int S, ResGeom& g) {
  if (S < 800 || S % 160) {
    set_error("WeSpeaker embedding: chunk length must be a multiple of 160 samples (>= 800)");
    return DG_EINVAL;
  }
  g.T0 = S / 160 - 2;                          // 1 + (S - 400) / 160, snip_edges
  g.W[0] = g.T0;
  g.H[0] = 80;
  for (int s = 1; s < 4; s++) {
    g.W[s] = (g.W[s - 1] - 1) / 2 + 1;
    g.H[s] = (g.H[s - 1] - 1) / 2 + 1;
  }
  return 0;
}

static int resnet_conv(const ResConv& c, const void* in_hi, const void* in_lo, int U, int Wp, int Hp, int Wop, int Hop,
                       void* out_hi, void* out_lo, float* out_f32, const void* res_hi, const void* res_lo, int relu,
                       const char* tag, cudaStream_t st) {
  int taps[9];
  if (c.ksize == 1) taps[0] = 0;
  else if (c.KW == 3)
    for (int dw = 0; dw < 3; dw++) taps[dw] = (dw - 1) * Hp - 1;           // three dh taps folded into one K slab
  else
    for (int dw = 0; dw < 3; dw++)
      for (int dh = 0; dh < 3; dh++) taps[dw * 3 + dh] = (dw - 1) * Hp + (dh - 1);
  TcGemm t{};
  const long long rows = (long long)U * Wp * Hp;
  t.A_hi = in_hi; t.A_lo = in_lo; t.lda = c.lda; t.Cin = c.cin_gemm; t.KW = c.KW; t.dil = 1; t.Mtot = rows; t.M = rows;
  t.W_hi = c.w_hi.p; t.w_scale = c.w_hi.wscale; t.W_lo = c.w_lo.p; t.Npad = c.cout <= 64 ? c.cout : (c.cout + 127) / 128 * 128; t.N = c.cout;
  t.bn_scale = c.sc.as<float>(); t.bn_shift = c.sh.as<float>();
  t.out_hi = out_hi; t.out_lo = out_lo; t.out_f32 = out_f32; t.ldc = c.cout; t.epi = 3; t.tag = tag;
  t.tap_off = taps; t.Wp = Wp; t.Hp = Hp; t.Wop = Wop; t.Hop = Hop; t.stride2 = c.stride == 2; t.relu = relu;
  t.res_hi = res_hi; t.res_lo = res_lo;
  return launch_gemm_tc(t, st);
}

// waveform [U,S] -> float32 final map [U][W3 + 2][H3 + 2][256] (h->pool_x descriptor), frames W3
static int resnet_trunk(dg_emb* h, const float* wav, int U, int S, cudaStream_t st, int* T_out) {
  int rc;
  ResNet& r = *h->rn;
  ResGeom g;
  if ((rc = resnet_geom(S, g))) return rc;
  const int rpi = S / 160;                                      // spectrum rows per item (the last two are not frames)
  const long long n = (long long)U * S;
  if (r.wav_hi.ensure(((size_t)n + 1024) * 2) || r.wav_lo.ensure(((size_t)n + 1024) * 2) ||
      r.spec.ensure(((size_t)U * rpi + 128) * 640 * 4) || r.logmel.ensure((size_t)U * g.T0 * 80 * 4) ||
      r.mean.ensure((size_t)U * 80 * 4))
    return DG_ECUDA;
  for (int s = 0; s < 4; s++) {
    const size_t rows = (size_t)U * (g.W[s] + 2) * (g.H[s] + 2) + 256;      // + tail: overlapping-row reads of the last rows
    for (int b = 0; b < 3; b++)
      for (int p = 0; p < 2; p++)
        if (r.act[s][b][p].ensure(rows * RN_CH[s] * 2)) return DG_ECUDA;    // zero-initialised: the padding ring stays zero
  }
  if (r.fin.ensure(((size_t)U * (g.W[3] + 2) * (g.H[3] + 2) + 64) * 256 * 4)) return DG_ECUDA;
  if (r.last_S != S) {
    if (r.last_S)
      for (int s = 0; s < 4; s++)
        for (int b = 0; b < 3; b++)
          for (int p = 0; p < 2; p++) DG_CUDA(cudaMemsetAsync(r.act[s][b][p].p, 0, r.act[s][b][p].bytes, st));
    r.last_S = S;
  }
  // ---- kaldi fbank: planes of x * 2^15, [rows, 448] x [448, 640] on the tensor cores, power -> mel -> log, time mean
  if ((rc = launch_fb_planes(wav, n, r.wav_hi.p, r.wav_lo.p, st))) return rc;
  {
    TcGemm t{};
    t.A_hi = r.wav_hi.p; t.A_lo = r.wav_lo.p; t.lda = 160; t.Cin = 448; t.KW = 1; t.dil = 1;
    t.Mtot = (long long)U * rpi; t.M = (long long)U * rpi;
    t.W_hi = r.fb_hi.p; t.w_scale = r.fb_hi.wscale; t.W_lo = r.fb_lo.p; t.Npad = 640; t.N = 640; t.out_f32 = r.spec.as<float>(); t.ldc = 640; t.epi = 0;
    t.tag = "fbank_dft";
    if ((rc = launch_gemm_tc(t, st))) return rc;
  }
  if ((rc = launch_fb_mel(r.spec.as<float>(), 640, rpi, g.T0, U, r.banks.as<float>(), r.k_lo.as<int>(), r.k_hi.as<int>(),
                          r.logmel.as<float>(), st)) ||
      (rc = launch_fb_mean(r.logmel.as<float>(), U, g.T0, r.mean.as<float>(), st)) ||
      (rc = launch_rn_stem(r.logmel.as<float>(), r.mean.as<float>(), U, g.T0, r.stem_w.as<float>(), r.stem_sc.as<float>(),
                           r.stem_sh.as<float>(), r.act[0][0][0].p, r.act[0][0][1].p, st)))
    return rc;
  // ---- 16 BasicBlocks: y = relu(bn1(conv1(x))); out = relu(bn2(conv2(y)) + shortcut(x))
  int cur = 0;                      // buffer (0 / 2) of the current stage that holds x
  int prev_stage = 0;
  static const char* kTags[4] = {"resnet_l1", "resnet_l2", "resnet_l3", "resnet_l4"};
  r.dbg_stage = 0;
  r.dbg_buf = 0;
  for (size_t bi = 0; bi < r.blocks.size() && (int)bi <= r.stop_after; bi++) {
    const ResBlock& blk = r.blocks[bi];
    const int s = r.stage_of[bi];
    const int Wp = g.W[s] + 2, Hp = g.H[s] + 2;
    DevBuf* x = r.act[prev_stage][cur];
    const int xWp = g.W[prev_stage] + 2, xHp = g.H[prev_stage] + 2;
    if (s != prev_stage) cur = 0;   // first block of a stage: x comes from the previous stage, the output goes to buffer 0
    DevBuf* y = r.act[s][1];
    DevBuf* out = s != prev_stage ? r.act[s][0] : r.act[s][cur ^ 2];
    const void *res_hi = x[0].p, *res_lo = x[1].p;
    if (blk.has_sc) {               // BatchNorm(Conv1x1 stride 2 (x)) into buffer 2 of this stage
      DevBuf* z = r.act[s][2];
      if ((rc = resnet_conv(blk.sc, x[0].p, x[1].p, U, xWp, xHp, Wp, Hp, z[0].p, z[1].p, nullptr, nullptr, nullptr, 0, kTags[s], st)))
        return rc;
      res_hi = z[0].p;
      res_lo = z[1].p;
    }
    if ((rc = resnet_conv(blk.c1, x[0].p, x[1].p, U, xWp, xHp, Wp, Hp, y[0].p, y[1].p, nullptr, nullptr, nullptr, 1, kTags[s], st)))
      return rc;
    const bool last = bi + 1 == r.blocks.size();
    if ((rc = resnet_conv(blk.c2, y[0].p, y[1].p, U, Wp, Hp, Wp, Hp, last ? nullptr : out[0].p, last ? nullptr : out[1].p,
                          last ? r.fin.as<float>() : nullptr, res_hi, res_lo, 1, kTags[s], st)))
      return rc;
    if (s == prev_stage) cur ^= 2;
    prev_stage = s;
    r.dbg_stage = s;
    r.dbg_buf = cur;
  }
  const int Wp3 = g.W[3] + 2, Hp3 = g.H[3] + 2;
  h->pool_x = r.fin.as<float>() + ((size_t)1 * Hp3 + 1) * 256;       // position (w = 1, h = 1) of item 0
  h->pool_item_pitch = (long long)Wp3 * Hp3 * 256;
  h->pool_row_pitch = Hp3 * 256;
  h->pool_C = g.H[3] * 256;
  *T_out = g.W[3];
  return 0;
}

// test hook: runs the variant-B trunk up to a given point and returns the intermediate map as float32 on the host.
// stop_after = -2: log-mel features [U][T0][80] (before mean normalisation), -1: stem output, k >= 0: output of BasicBlock k
// (dims = {U, W, H, C}, un-padded, layout [item][w = time][h = mel][channel]); 15 = the final map.
extern "C" int dg_emb_debug_trunk(dg_emb* h, const float* wav_dev, int U, int S, int stop_after, float* out_host, int64_t cap,
                                  int* dims) {
  if (!h || h->variant != 1 || !wav_dev || !out_host || !dims || U < 1) {
    set_error("dg_emb_debug_trunk: needs a WeSpeaker (variant B) handle");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  ResNet& r = *h->rn;
  ResGeom g;
  int rc, T = 0;
  if ((rc = resnet_geom(S, g))) return rc;
  r.stop_after = stop_after < -1 ? -1 : stop_after;
  rc = resnet_trunk(h, wav_dev, U, S, nullptr, &T);
  r.stop_after = 99;
  if (rc) return rc;
  DG_CUDA(cudaDeviceSynchronize());
  if (stop_after == -2) {
    dims[0] = U; dims[1] = g.T0; dims[2] = 80; dims[3] = 1;
    const int64_t n = (int64_t)U * g.T0 * 80;
    if (n > cap) return DG_EINVAL;
    DG_CUDA(cudaMemcpy(out_host, r.logmel.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return DG_OK;
  }
  const int s = r.dbg_stage, W = g.W[s], H = g.H[s], C = RN_CH[s], Wp = W + 2, Hp = H + 2;
  dims[0] = U; dims[1] = W; dims[2] = H; dims[3] = C;
  const int64_t n = (int64_t)U * W * H * C;
  if (n > cap) {
    set_error("dg_emb_debug_trunk: buffer too small");
    return DG_EINVAL;
  }
  const size_t rows = (size_t)U * Wp * Hp;
  std::vector<float> full(rows * C);
  if (stop_after >= 15) {
    DG_CUDA(cudaMemcpy(full.data(), r.fin.p, rows * C * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> hi(rows * C), lo(rows * C);
    DG_CUDA(cudaMemcpy(hi.data(), r.act[s][r.dbg_buf][0].p, rows * C * 2, cudaMemcpyDeviceToHost));
    DG_CUDA(cudaMemcpy(lo.data(), r.act[s][r.dbg_buf][1].p, rows * C * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < rows * C; i++) full[i] = host_h16_to_f32(hi[i], split_f16()) + host_h16_to_f32(lo[i], split_f16());
  }
  for (int u = 0; u < U; u++)
    for (int w = 0; w < W; w++)
      for (int hh = 0; hh < H; hh++)
        memcpy(out_host + (((size_t)u * W + w) * H + hh) * C, &full[(((size_t)u * Wp + w + 1) * Hp + hh + 1) * C], (size_t)C * 4);
  return DG_OK;
}

extern "C" int dg_emb_create(const dg_tensor* tensors, int n, int pool_mode, int device, dg_emb** out) {
  if (!tensors || !out || (pool_mode != 31 && pool_mode != 21)) {
    set_error("dg_emb_create: bad arguments (pool_mode must be 31 or 21)");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(device));
  std::unique_ptr<dg_emb> h(new dg_emb());
  h->device = device;
  h->pool_mode = pool_mode;
  Tensors t(tensors, n);
  int rc = emb_prepare(h.get(), t);
  if (rc) return rc;
  *out = h.release();
  return DG_OK;
}

extern "C" int dg_emb_dims(const dg_emb* h, int num_samples, int* frames, int* dimension) {
  if (!h || num_samples < 3000) {
    set_error("dg_emb_dims: bad arguments");
    return DG_EINVAL;
  }
  if (h->variant == 1) {
    ResGeom rg;
    int rc = resnet_geom(num_samples, rg);
    if (rc) return rc;
    if (frames) *frames = rg.W[3];
    if (dimension) *dimension = h->D;
    return DG_OK;
  }
  Geom g = make_geom(num_samples);
  if (frames) *frames = g.T2 - 14;
  if (dimension) *dimension = h->D;
  return DG_OK;
}

// F.interpolate index tables, computed in float32 exactly like ATen's upsample kernels
static int build_tables(dg_emb* h, int F, int T, cudaStream_t st) {
  if (h->tab_F == F && h->tab_T == T) return 0;
  std::vector<int> i0(T), i1(T);
  std::vector<float> l1(T);
  const float scale = (float)F / (float)T;
  for (int t = 0; t < T; t++) {
    if (F == T) {
      i0[t] = i1[t] = t;
      l1[t] = 0.f;
    } else if (h->pool_mode == 31) {   // mode="nearest": min(floor(dst * scale), F - 1)
      int s = (int)floorf((float)t * scale);
      if (s > F - 1) s = F - 1;
      i0[t] = i1[t] = s;
      l1[t] = 0.f;
    } else {                           // mode="linear", align_corners=False
      float src = scale * ((float)t + 0.5f) - 0.5f;
      if (src < 0.f) src = 0.f;
      int a = (int)src;
      if (a > F - 1) a = F - 1;
      i0[t] = a;
      i1[t] = a + (a < F - 1 ? 1 : 0);
      l1[t] = src - (float)a;
    }
  }
  if (h->idx0.ensure(T * 4) || h->idx1.ensure(T * 4) || h->lam1.ensure(T * 4)) return DG_ECUDA;
  DG_CUDA(cudaStreamSynchronize(st));
  DG_CUDA(cudaMemcpy(h->idx0.p, i0.data(), T * 4, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->idx1.p, i1.data(), T * 4, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->lam1.p, l1.data(), T * 4, cudaMemcpyHostToDevice));
  h->tab_F = F;
  h->tab_T = T;
  return 0;
}

// waveform [U,S] -> t5 [U*S2, 1500]; returns the number of valid frames
// DG_NO_POOL_FUSE=1: A/B switch, TDNN5 writes its map and stats_pool reads it back (the round-1 path)
static bool pool_fusion_on() {
  static const bool off = getenv("DG_NO_POOL_FUSE") && getenv("DG_NO_POOL_FUSE")[0] == '1';
  return !off && use_tensor_cores();
}

// `defer_last`: stop before TDNN5 (its operand planes are left in h->t4h / t4l) -- the caller runs it fused with the pooling
static int emb_trunk(dg_emb* h, const float* wav, int U, const Geom& g, cudaStream_t st, int* T_out, bool defer_last = false) {
  int rc;
  if (h->variant == 1) return resnet_trunk(h, wav, U, g.S, st, T_out);
  if ((rc = run_sincnet(h->sw, h->work, wav, U, g, st, h->shared_prep))) return rc;
  const size_t rows = (size_t)U * g.S2 + 64;
  if (h->tA.ensure(rows * 512 * 4) || h->tB.ensure(rows * 512 * 4) || h->t5.ensure(rows * 1500 * 4)) return DG_ECUDA;
  h->pool_x = h->t5.as<float>();
  h->pool_item_pitch = (long long)g.S2 * 1500;
  h->pool_row_pitch = 1500;
  h->pool_C = 1500;
  const long long M = (long long)U * g.S2;
  if (use_tensor_cores()) {
    if (h->xh.ensure(rows * 64 * 2) || h->xl.ensure(rows * 64 * 2) || h->aH.ensure(rows * 512 * 2) ||
        h->aL.ensure(rows * 512 * 2) || h->bH.ensure(rows * 512 * 2) || h->bL.ensure(rows * 512 * 2))
      return DG_ECUDA;
    if ((rc = launch_split_ex(h->work.out, M, 64, 64, 64, h->work.out_pool, g.S2, h->work.sc2.as<float>(),
                              h->work.sh2.as<float>(), h->xh.p, h->xl.p, st)))
      return rc;
    const void *ih = h->xh.p, *il = h->xl.p;
    int cin = 64, T = g.T2;
    void* oh[2] = {h->aH.p, h->bH.p};
    void* ol[2] = {h->aL.p, h->bL.p};
    static const char* kTags[5] = {"tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5"};
    for (int L = 0; L < 5; L++) {
      if (L == 4 && defer_last) {
        h->t4h = ih;
        h->t4l = il;
        T -= (TD_K[L] - 1) * TD_DIL[L];
        break;
      }
      TcGemm t{};
      t.A_hi = ih; t.A_lo = il; t.lda = cin; t.Cin = cin; t.KW = TD_K[L]; t.dil = TD_DIL[L]; t.Mtot = M; t.M = M;
      t.W_hi = h->tw_hi[L].p; t.w_scale = h->tw_hi[L].wscale; t.W_lo = h->tw_lo[L].p; t.Npad = (TD_OUT[L] + 255) / 256 * 256; t.N = TD_OUT[L];
      t.bias = h->tb[L].as<float>(); t.bn_scale = h->bns[L].as<float>(); t.bn_shift = h->bnh[L].as<float>();
      t.tag = kTags[L];
      if (L == 4) {
        t.out_f32 = h->t5.as<float>(); t.ldc = 1500; t.epi = 2;
      } else {
        t.out_hi = oh[L & 1]; t.out_lo = ol[L & 1]; t.ldc = 512; t.epi = 1;
      }
      if ((rc = launch_gemm_tc(t, st))) return rc;
      ih = oh[L & 1]; il = ol[L & 1];
      cin = TD_OUT[L];
      T -= (TD_K[L] - 1) * TD_DIL[L];
    }
    *T_out = T;
    return 0;
  }
  const float* in = h->work.p2.as<float>();
  int cin = 64, T = g.T2;
  float* bufs[2] = {h->tA.as<float>(), h->tB.as<float>()};
  for (int L = 0; L < 5; L++) {
    GemmArgs a{};
    a.A = in; a.lda = cin; a.Cin = cin; a.KW = TD_K[L]; a.dil = TD_DIL[L];
    a.Mtot = M; a.M = M;
    a.W = h->tw[L].as<float>(); a.ldw = TD_OUT[L]; a.N = TD_OUT[L]; a.bias = h->tb[L].as<float>();
    a.bn_scale = h->bns[L].as<float>(); a.bn_shift = h->bnh[L].as<float>();
    if (L == 0) {
      a.in_sc = h->work.sc2.as<float>(); a.in_sh = h->work.sh2.as<float>(); a.item_rows = g.S2;
    }
    float* out = L == 4 ? h->t5.as<float>() : bufs[L & 1];
    a.C = out; a.ldc = TD_OUT[L]; a.epi = EPI_BIAS_LEAKY_BN;
    static const char* kTags[5] = {"tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5"};
    a.tag = kTags[L];
    if ((rc = launch_gemm(a, st))) return rc;
    in = out;
    cin = TD_OUT[L];
    T -= (TD_K[L] - 1) * TD_DIL[L];
  }
  *T_out = T;
  return 0;
}

// TDNN5 (Conv1d(512, 1500, 1) -> LeakyReLU -> BatchNorm) fused with the K weighted statistics poolings: the [rows, 1500] map
// (455 MB at B = 256) is never written; the epilogue leaves per-tile partial sums, pool_finalize turns them into mean / std.
static int emb_tdnn5_pool(dg_emb* h, int U, const Geom& g, const float* weights, int F, int K, int T, cudaStream_t st) {
  int rc;
  const long long M = (long long)U * g.S2;
  const int m_tiles = (int)((M + 127) / 128);
  const float eps = h->pool_mode == 31 ? 1e-8f : 0.f;
  if (h->pool_rw.ensure(((size_t)M + 128) * 16) || h->pool_vs.ensure((size_t)U * K * 8) ||
      h->pool_part.ensure((size_t)m_tiles * 2 * 8 * 1500 * 4) || h->pooled.ensure((size_t)U * K * 3000 * 4))
    return DG_ECUDA;
  if ((rc = launch_pool_weights(weights, U, F, K, g.S2, T, h->idx0.as<int>(), h->idx1.as<int>(), h->lam1.as<float>(), eps,
                                h->pool_rw.as<float>(), h->pool_vs.as<float>(), st)))
    return rc;
  TcGemm t{};
  t.A_hi = h->t4h; t.A_lo = h->t4l; t.lda = 512; t.Cin = 512; t.KW = 1; t.dil = 1; t.Mtot = M; t.M = M;
  t.W_hi = h->tw_hi[4].p; t.w_scale = h->tw_hi[4].wscale; t.W_lo = h->tw_lo[4].p; t.Npad = 1536; t.N = 1500;
  t.bias = h->tb[4].as<float>(); t.bn_scale = h->bns[4].as<float>(); t.bn_shift = h->bnh[4].as<float>();
  t.ldc = 1500; t.epi = 4; t.tag = "tdnn5";
  t.pool_w = h->pool_rw.as<float>(); t.pool_part = h->pool_part.as<float>(); t.pool_item_rows = g.S2; t.pool_K = K;
  if ((rc = launch_gemm_tc(t, st))) return rc;
  h->pool_C = 1500;
  return launch_pool_finalize(h->pool_part.as<float>(), h->pool_vs.as<float>(), h->bnh[4].as<float>(), U, K, 1500, g.S2, T, eps,
                              h->pooled.as<float>(), st);
}

static int emb_project(dg_emb* h, int rows, int normalize, float norm, float* out, cudaStream_t st) {
  if (use_tensor_cores() || h->variant == 1) {
    int rc;
    const int nfeat = 2 * h->pool_C, kpad = (nfeat + 63) / 64 * 64;     // 3000 -> 3008, 5120 -> 5120
    if (h->ph.ensure(((size_t)rows + 128) * kpad * 2) || h->pl.ensure(((size_t)rows + 128) * kpad * 2)) return DG_ECUDA;
    if ((rc = launch_split_ex(h->pooled.as<float>(), rows, nfeat, nfeat, kpad, 0, 1, nullptr, nullptr, h->ph.p, h->pl.p, st)))
      return rc;
    float* dst = out;
    if (normalize) {
      if (h->eraw.ensure((size_t)rows * h->D * 4)) return DG_ECUDA;
      dst = h->eraw.as<float>();
    }
    TcGemm t{};
    t.A_hi = h->ph.p; t.A_lo = h->pl.p; t.lda = kpad; t.Cin = kpad; t.KW = 1; t.dil = 1; t.Mtot = rows; t.M = rows;
    t.W_hi = h->ew_hi.p; t.w_scale = h->ew_hi.wscale; t.W_lo = h->ew_lo.p; t.Npad = (h->D + 255) / 256 * 256; t.N = h->D;
    t.bias = h->eb.as<float>(); t.out_f32 = dst; t.ldc = h->D; t.epi = 0; t.tag = "emb_linear";
    if ((rc = launch_gemm_tc(t, st))) return rc;
    return normalize ? launch_l2norm(dst, rows, h->D, norm, out, st) : 0;
  }
  GemmArgs a{};
  a.A = h->pooled.as<float>(); a.lda = 3000; a.Cin = 3000; a.KW = 1; a.dil = 1; a.Mtot = rows; a.M = rows;
  a.W = h->ew.as<float>(); a.ldw = h->D; a.N = h->D; a.bias = h->eb.as<float>();
  a.ldc = h->D; a.epi = EPI_BIAS; a.tag = "emb_linear";
  if (!normalize) {
    a.C = out;
    return launch_gemm(a, st);
  }
  if (h->eraw.ensure((size_t)rows * h->D * 4)) return DG_ECUDA;
  a.C = h->eraw.as<float>();
  int rc;
  if ((rc = launch_gemm(a, st))) return rc;
  return launch_l2norm(h->eraw.as<float>(), rows, h->D, norm, out, st);
}

extern "C" int dg_emb_forward(dg_emb* h, const float* wav, const float* weights, int B, int S, int F, int K,
                              int normalize, float norm, float* out, void* stream) {
  if (!h || !wav || !out || B < 1 || S < 3000 || K < 1 || (!weights && K != 1) || (weights && F < 1)) {
    set_error("dg_emb_forward: bad arguments");
    return DG_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DG_CUDA(cudaSetDevice(h->device));
  const Geom g = make_geom(S);
  int rc, T = 0;
  const void* me = stream ? stream : (void*)h;
  struct Done {              // every exit of this call marks the end of the use
    dg_emb* h; const void* me; cudaStream_t st; bool on;
    ~Done() { if (on) use_end(h->guard, me, st); }
  } done{h, me, st, !g_in_pipeline};
  if (!g_in_pipeline && (rc = use_begin(h->guard, me, st))) return rc;
  const bool fuse = weights && h->variant == 0 && K <= 4 && g.S2 >= 128 && pool_fusion_on();
  if ((rc = emb_trunk(h, wav, B, g, st, &T, fuse))) return rc;
  if (weights && (rc = build_tables(h, F, T, st))) return rc;
  if (fuse) {
    if ((rc = emb_tdnn5_pool(h, B, g, weights, F, K, T, st))) return rc;
    return emb_project(h, B * K, normalize, norm, out, st);
  }
  if (h->pooled.ensure((size_t)B * K * 2 * h->pool_C * 4)) return DG_ECUDA;
  const float eps = h->pool_mode == 31 ? 1e-8f : 0.f;
  if ((rc = launch_stats_pool(h->pool_x, B, g.S2, T, h->pool_C, weights, F, K, h->idx0.as<int>(),
                              h->idx1.as<int>(), h->lam1.as<float>(), weights ? eps : 0.f,
                              h->pooled.as<float>(), st, h->pool_item_pitch, h->pool_row_pitch)))
    return rc;
  return emb_project(h, B * K, normalize, norm, out, st);
}

extern "C" int dg_emb_forward_rows(dg_emb* h, const float* wav, const float* weights, int N, int S, int F, float* out,
                                   void* stream) {
  if (!h || !wav || !out || N < 1 || S < 3000 || (weights && F < 1)) {
    set_error("dg_emb_forward_rows: bad arguments");
    return DG_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DG_CUDA(cudaSetDevice(h->device));
  const Geom g = make_geom(S);
  int rc, T = 0;
  const void* me = stream ? stream : (void*)h;
  struct Done {
    dg_emb* h; const void* me; cudaStream_t st; bool on;
    ~Done() { if (on) use_end(h->guard, me, st); }
  } done{h, me, st, !g_in_pipeline};
  if (!g_in_pipeline && (rc = use_begin(h->guard, me, st))) return rc;
  // consecutive identical rows (the reference repeats each waveform once per local speaker,
  // src/diart/blocks/embedding.py:57-59) share one trunk pass
  if (h->flags.ensure((size_t)N * 4)) return DG_ECUDA;
  if ((rc = launch_row_equal_flags(wav, N, S, h->flags.as<int>(), st))) return rc;
  std::vector<int> flags(N);
  DG_CUDA(cudaMemcpyAsync(flags.data(), h->flags.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaStreamSynchronize(st));
  std::vector<int> uniq, gi, gq0, gnq;
  for (int n = 0; n < N; n++) {
    if (!flags[n]) uniq.push_back(n);
    const int item = (int)uniq.size() - 1;
    if (!flags[n] || gnq.back() == 4) {
      gi.push_back(item);
      gq0.push_back(n);
      gnq.push_back(1);
    } else {
      gnq.back()++;
    }
  }
  const int U = (int)uniq.size(), G = (int)gi.size();
  const float* trunk_in = wav;
  if (U != N) {
    if (h->uniq.ensure((size_t)U * 4) || h->gathered.ensure((size_t)U * S * 4)) return DG_ECUDA;
    DG_CUDA(cudaMemcpyAsync(h->uniq.p, uniq.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
    if ((rc = launch_gather_rows(wav, h->uniq.as<int>(), U, S, h->gathered.as<float>(), st))) return rc;
    trunk_in = h->gathered.as<float>();
  }
  if (h->grp.ensure((size_t)3 * G * 4)) return DG_ECUDA;
  std::vector<int> packed(3 * G);
  memcpy(packed.data(), gi.data(), G * 4);
  memcpy(packed.data() + G, gq0.data(), G * 4);
  memcpy(packed.data() + 2 * G, gnq.data(), G * 4);
  DG_CUDA(cudaMemcpyAsync(h->grp.p, packed.data(), (size_t)3 * G * 4, cudaMemcpyHostToDevice, st));
  if ((rc = emb_trunk(h, trunk_in, U, g, st, &T))) return rc;
  if (weights && (rc = build_tables(h, F, T, st))) return rc;
  if (h->pooled.ensure((size_t)N * 2 * h->pool_C * 4)) return DG_ECUDA;
  const float eps = (weights && h->pool_mode == 31) ? 1e-8f : 0.f;
  const int* gp = h->grp.as<int>();
  if ((rc = launch_stats_pool_ex(h->pool_x, g.S2, T, h->pool_C, weights, F, 1, 1, G, gp, gp + G, gp + 2 * G,
                                 h->idx0.as<int>(), h->idx1.as<int>(), h->lam1.as<float>(), eps,
                                 h->pooled.as<float>(), st, h->pool_item_pitch, h->pool_row_pitch)))
    return rc;
  rc = emb_project(h, N, 0, 1.f, out, st);
  DG_CUDA(cudaStreamSynchronize(st));   // host staging vectors above must outlive the async copies
  return rc;
}

extern "C" int dg_emb_destroy(dg_emb* h) {
  delete h;
  return DG_OK;
}

// =========================================================================== element-wise blocks
extern "C" int dg_osp(const float* seg, int B, int F, int K, float gamma, float beta, int normalize, float* out,
                      void* stream) {
  if (!seg || !out || B < 1 || F < 1 || K < 1) {
    set_error("dg_osp: bad arguments");
    return DG_EINVAL;
  }
  return launch_osp(seg, B, F, K, gamma, beta, normalize, out, (cudaStream_t)stream);
}

extern "C" int dg_normalize_embeddings(const float* emb, int rows, int D, float norm, float* out, void* stream) {
  if (!emb || !out || rows < 1 || D < 1) {
    set_error("dg_normalize_embeddings: bad arguments");
    return DG_EINVAL;
  }
  return launch_l2norm(emb, rows, D, norm, out, (cudaStream_t)stream);
}

// ==================================================================================== clustering
struct dg_cluster {
  int device = 0;
  ClusterParams p;
  DevBuf centers, active, init, prep, prep_d, record;
  DevBuf base, base_active, relabel;   // shared-identity mode: table at the last merge, relabel of created centres
};

extern "C" int dg_cluster_create(int max_speakers, int dim, double tau, double rho, double delta, int device,
                                 dg_cluster** out) {
  if (!out || max_speakers < 1 || max_speakers > 32 || dim < 1) {
    set_error("dg_cluster_create: need 1 <= max_speakers <= 32 and dim >= 1");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(device));
  std::unique_ptr<dg_cluster> h(new dg_cluster());
  h->device = device;
  h->p.M = max_speakers;
  h->p.D = dim;
  // numpy compares a float32 array with a Python float in float32 (weak scalar promotion)
  h->p.tau_f = (float)tau;
  h->p.rho_f = (float)rho;
  h->p.delta = delta;
  h->p.metric = 0;
  if (h->centers.ensure((size_t)max_speakers * dim * 8) || h->active.ensure(32 * 4) || h->init.ensure(2 * 4) ||
      h->base.ensure((size_t)max_speakers * dim * 8) || h->base_active.ensure(32 * 4) || h->relabel.ensure(32 * 4))
    return DG_ECUDA;
  *out = h.release();
  return DG_OK;
}

extern "C" int dg_cluster_set_metric(dg_cluster* h, int metric) {
  if (!h || metric < 0 || metric > 4) {
    set_error("dg_cluster_set_metric: 0 cosine, 1 euclidean, 2 sqeuclidean, 3 cityblock, 4 chebyshev");
    return DG_EINVAL;
  }
  h->p.metric = metric;
  return DG_OK;
}

extern "C" int dg_cluster_step(dg_cluster* h, const float* seg, const float* emb, int B, int F, int K, int32_t* map,
                               float* permuted, void* stream) {
  if (!h || !seg || !emb || !map || B < 0 || F < 1 || K < 1) {
    set_error("dg_cluster_step: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  if (h->prep.ensure(cluster_prep_floats(B, K) * 4 + 16) || h->prep_d.ensure(cluster_prep_doubles(B, K) * 8 + 16))
    return DG_ECUDA;
  return launch_cluster_step(h->p, seg, emb, B, F, K, h->centers.as<double>(), h->active.as<int>(),
                             h->init.as<int>(), h->prep.as<float>(), h->prep_d.as<double>(), map, permuted,
                             (cudaStream_t)stream);
}

extern "C" int dg_cluster_reset(dg_cluster* h) {
  if (!h) return DG_EINVAL;
  DG_CUDA(cudaSetDevice(h->device));
  DG_CUDA(cudaDeviceSynchronize());
  DG_CUDA(cudaMemset(h->centers.p, 0, h->centers.bytes));
  DG_CUDA(cudaMemset(h->active.p, 0, h->active.bytes));
  DG_CUDA(cudaMemset(h->init.p, 0, h->init.bytes));
  DG_CUDA(cudaMemset(h->base.p, 0, h->base.bytes));
  DG_CUDA(cudaMemset(h->base_active.p, 0, h->base_active.bytes));
  return DG_OK;
}

extern "C" int dg_cluster_get_state(dg_cluster* h, double* centers, int32_t* active, int* initialized) {
  if (!h) return DG_EINVAL;
  DG_CUDA(cudaSetDevice(h->device));
  DG_CUDA(cudaDeviceSynchronize());
  int init[2] = {0, 0};
  DG_CUDA(cudaMemcpy(init, h->init.p, 8, cudaMemcpyDeviceToHost));
  if (init[1]) {
    set_error("Cannot update unknown centers");   // reference clustering.py:98 (AssertionError)
    return DG_EINVAL;
  }
  if (centers) DG_CUDA(cudaMemcpy(centers, h->centers.p, (size_t)h->p.M * h->p.D * 8, cudaMemcpyDeviceToHost));
  if (active) DG_CUDA(cudaMemcpy(active, h->active.p, (size_t)h->p.M * 4, cudaMemcpyDeviceToHost));
  if (initialized) *initialized = init[0];
  return DG_OK;
}

extern "C" int dg_cluster_set_state(dg_cluster* h, const double* centers, const int32_t* active, int initialized) {
  if (!h || !centers || !active) return DG_EINVAL;
  DG_CUDA(cudaSetDevice(h->device));
  DG_CUDA(cudaDeviceSynchronize());
  int init[2] = {initialized ? 1 : 0, 0};
  DG_CUDA(cudaMemcpy(h->centers.p, centers, (size_t)h->p.M * h->p.D * 8, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->active.p, active, (size_t)h->p.M * 4, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->init.p, init, 8, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->base.p, centers, (size_t)h->p.M * h->p.D * 8, cudaMemcpyHostToDevice));
  DG_CUDA(cudaMemcpy(h->base_active.p, active, (size_t)h->p.M * 4, cudaMemcpyHostToDevice));
  return DG_OK;
}

extern "C" int dg_cluster_destroy(dg_cluster* h) {
  delete h;
  return DG_OK;
}

// shared-identity extension (BASELINE config 5); kernels and rule in cluster.cu
extern "C" int dg_cluster_record_len(const dg_cluster* h) { return h ? h->p.M * h->p.D + h->p.M + 2 : 0; }

extern "C" int dg_cluster_export_delta(dg_cluster* h, double* record_dev, void* stream) {
  if (!h || !record_dev) {
    set_error("dg_cluster_export_delta: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  return launch_cluster_export(h->centers.as<double>(), h->active.as<int>(), h->base.as<double>(),
                               h->base_active.as<int>(), h->p.M, h->p.D, record_dev, (cudaStream_t)stream);
}

extern "C" int dg_cluster_merge(dg_cluster* h, const double* records_dev, int world, int rank, int32_t* maps_dev,
                                int n_maps, void* stream) {
  if (!h || !records_dev || world < 1 || rank < 0 || rank >= world) {
    set_error("dg_cluster_merge: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = launch_cluster_merge(records_dev, world, rank, h->p, dg_cluster_record_len(h), h->centers.as<double>(),
                                 h->active.as<int>(), h->base.as<double>(), h->base_active.as<int>(),
                                 h->init.as<int>(), h->relabel.as<int32_t>(), (cudaStream_t)stream)))
    return rc;
  if (maps_dev && n_maps > 0) return launch_relabel_maps(maps_dev, n_maps, h->relabel.as<int32_t>(), (cudaStream_t)stream);
  return DG_OK;
}

// ================================================================================== self test
extern "C" int dg_selftest_split_host(const float* x, long long n, int f16, unsigned short* hi, unsigned short* lo) {
  if (!x || !hi || !lo || n < 0) {
    set_error("dg_selftest_split_host: null argument");
    return DG_EINVAL;
  }
  for (long long i = 0; i < n; i++) {
    hi[i] = host_f32_to_h16(x[i], f16);
    lo[i] = host_f32_to_h16(x[i] - host_h16_to_f32(hi[i], f16), f16);
  }
  return DG_OK;
}

// Runs the same shifted-window GEMM through the float32 SIMT kernel and through the tcgen05 split-precision
// kernel on seeded random data and reports the largest absolute difference and the output scale.
extern "C" int dg_selftest_gemm_tc(int M, int Cin, int KW, int dil, int N, int epi, float* max_abs_diff,
                                   float* out_rms) {
  if (M < 1 || Cin % 64 || KW < 1 || N % 4 || !max_abs_diff || !out_rms) {
    set_error("dg_selftest_gemm_tc: bad arguments");
    return DG_EINVAL;
  }
  const int K = KW * Cin, npad = N == 64 ? 64 : (N + 255) / 256 * 256;
  const long long Mtot = M;
  std::vector<float> A((size_t)Mtot * Cin), Wkn((size_t)K * N), Wnk((size_t)N * K), bias(N), bsc(N), bsh(N);
  uint32_t seed = 12345u;
  auto rnd = [&]() {
    seed = seed * 1664525u + 1013904223u;
    return ((seed >> 8) & 0xFFFF) / 65536.f - 0.5f;
  };
  // DG_SELFTEST_AMP: amplitude of the A operand (default 2): small values put the whole lo plane into fp16's subnormal range
  static const float amp = getenv("DG_SELFTEST_AMP") ? (float)atof(getenv("DG_SELFTEST_AMP")) : 2.f;
  for (auto& v : A) v = amp * rnd();
  for (int k = 0; k < K; k++)
    for (int n = 0; n < N; n++) {
      const float w = rnd() * 0.25f;
      Wkn[(size_t)k * N + n] = w;
      Wnk[(size_t)n * K + k] = w;
    }
  for (int n = 0; n < N; n++) {
    bias[n] = rnd();
    bsc[n] = 1.f + rnd();
    bsh[n] = rnd();
  }
  DevBuf dA, dWkn, dWh, dWl, dB, dS, dH, dAh, dAl, dC0, dC1, dOh, dOl;
  if (upload(dA, A) || upload(dWkn, Wkn) || upload(dB, bias) || upload(dS, bsc) || upload(dH, bsh) ||
      upload_split(dWh, dWl, Wnk, N, npad, K))
    return DG_ECUDA;
  if (dAh.ensure((size_t)Mtot * Cin * 2) || dAl.ensure((size_t)Mtot * Cin * 2) || dC0.ensure((size_t)M * N * 4) ||
      dC1.ensure((size_t)M * N * 4) || dOh.ensure((size_t)M * N * 2) || dOl.ensure((size_t)M * N * 2))
    return DG_ECUDA;
  int rc;
  GemmArgs g{};
  g.A = dA.as<float>(); g.lda = Cin; g.Cin = Cin; g.KW = KW; g.dil = dil; g.Mtot = Mtot; g.M = M;
  g.W = dWkn.as<float>(); g.ldw = N; g.N = N; g.bias = dB.as<float>(); g.bn_scale = dS.as<float>();
  g.bn_shift = dH.as<float>(); g.C = dC0.as<float>(); g.ldc = N; g.epi = epi == 0 ? EPI_BIAS : EPI_BIAS_LEAKY_BN;
  g.tag = "selftest_simt";
  if ((rc = launch_gemm(g, nullptr))) return rc;
  if ((rc = launch_split(dA.as<float>(), Mtot, Cin, 1, nullptr, nullptr, dAh.p, dAl.p, nullptr))) return rc;
  TcGemm t{};
  t.A_hi = dAh.p; t.A_lo = dAl.p; t.lda = Cin; t.Cin = Cin; t.KW = KW; t.dil = dil; t.Mtot = Mtot; t.M = M;
  t.W_hi = dWh.p; t.w_scale = dWh.wscale; t.W_lo = dWl.p; t.Npad = npad; t.N = N; t.bias = dB.as<float>(); t.bn_scale = dS.as<float>();
  t.bn_shift = dH.as<float>(); t.out_f32 = dC1.as<float>(); t.out_hi = dOh.p; t.out_lo = dOl.p; t.ldc = N;
  t.epi = epi; t.tag = "selftest_tc";
  if ((rc = launch_gemm_tc(t, nullptr))) return rc;
  DG_CUDA(cudaDeviceSynchronize());
  std::vector<float> c0((size_t)M * N), c1((size_t)M * N);
  DG_CUDA(cudaMemcpy(c0.data(), dC0.p, c0.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint16_t> oh, ol;
  if (epi == 1) {
    oh.resize((size_t)M * N);
    ol.resize((size_t)M * N);
    DG_CUDA(cudaMemcpy(oh.data(), dOh.p, oh.size() * 2, cudaMemcpyDeviceToHost));
    DG_CUDA(cudaMemcpy(ol.data(), dOl.p, ol.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < c1.size(); i++) c1[i] = host_h16_to_f32(oh[i], split_f16()) + host_h16_to_f32(ol[i], split_f16());
  } else {
    DG_CUDA(cudaMemcpy(c1.data(), dC1.p, c1.size() * 4, cudaMemcpyDeviceToHost));
  }
  double md = 0, ss = 0;
  size_t worst = 0, n_big = 0;
  for (size_t i = 0; i < c0.size(); i++) {
    const double d = fabs((double)c0[i] - (double)c1[i]);
    if (!(d <= md)) {     // NaN-propagating max
      md = d;
      worst = i;
    }
    if (!(d <= 1e-2)) n_big++;
    ss += (double)c0[i] * c0[i];
  }
  if (n_big)
    fprintf(stderr, "dg_selftest_gemm_tc: %zu of %zu outputs differ by more than 1e-2; worst at row %zu col %zu: simt %g, tcgen05 %g\n",
            n_big, c0.size(), worst / N, worst % N, c0[worst], c1[worst]);
  if (n_big && epi == 1) {
    size_t shown = 0;
    for (size_t i = 0; i < c0.size() && shown < 12; i++)
      if (!(fabs((double)c0[i] - (double)c1[i]) <= 1e-2)) {
        fprintf(stderr, "   row %zu col %zu: simt %g, planes hi 0x%04x lo 0x%04x\n", i / N, i % N, c0[i], oh[i], ol[i]);
        shown++;
      }
  }
  *max_abs_diff = (float)md;
  *out_rms = (float)sqrt(ss / c0.size());
  return DG_OK;
}

// ================================================================================ fused pipeline
// Persistent worker threads for the host-side gather of dg_pipeline_call_host (B separate pageable windows -> pinned staging):
// created once per pipeline handle; a job is one callable that every worker runs concurrently (the callable hands out work
// items through its own atomic counter).
class GatherPool {
 public:
  explicit GatherPool(int n) {
    for (int i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
  }
  ~GatherPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  void start(std::function<void()> fn) {       // returns at once; wait() returns when every worker has finished fn
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = std::move(fn);
      generation_++;
      active_ = (int)th_.size();
    }
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return active_ == 0; });
  }

 private:
  void loop() {
    int seen = 0;
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        fn = job_;
      }
      fn();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--active_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<void()> job_;
  int generation_ = 0, active_ = 0;
  bool stop_ = false;
};

struct dg_pipeline {
  dg_seg* seg;
  dg_emb* emb;
  dg_cluster* clu;
  float gamma, beta;
  int normalize_weights;
  int hop = 0;          // samples between consecutive windows of a batch (hint, dg_pipeline_set_hop); 0 = unknown
  DevBuf osp, wav, segd, embd, mapd, permd;
  cudaStream_t st = nullptr;
  // two-stream overlap inside a step: the segmentation chain (critical path, high priority) and the
  // embedding trunk (independent of it until the pooling weights exist) run concurrently
  cudaStream_t s_seg = nullptr, s_seg2 = nullptr, s_emb = nullptr, s_clu = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  DevBuf osp2;
  cudaEvent_t e_osp2 = nullptr;
  SincPrep prep[2];
  cudaEvent_t e_prep[2] = {nullptr, nullptr};
  cudaEvent_t e_start = nullptr, e_osp = nullptr, e_emb = nullptr, e_done = nullptr;
  // pipelining (dg_pipeline_submit* / collect*): up to DG_MAX_INFLIGHT steps outstanding.  Step n uses result / input
  // slot n % 3 and scratch lane n & 1: two steps compute concurrently (lanes), the third slot lets the host upload the
  // waveforms of step n+2 while steps n and n+1 are on the device
  DevBuf slot_wav[3], slot_seg[3], slot_emb[3], slot_map[3];
  cudaEvent_t e_h2d[3] = {nullptr, nullptr, nullptr}, e_slot_done[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t e_lane_done[2] = {nullptr, nullptr};
  int slot_B[3] = {0, 0, 0}, slot_S[3] = {0, 0, 0}, outstanding = 0;
  long long next_step = 0;
  // shared-identity mode inside the pipelined flow: export / merge run on the clustering stream, in order with the clustering
  // of the submitted steps, so the networks of the next steps keep running meanwhile
  cudaEvent_t e_ident = nullptr, e_ident_in = nullptr;
  long long ident_merged_upto = 0;      // steps below this index have had their maps relabelled by a merge
  bool overlap_known = false;     // the current dg_pipeline_step batch was formed from a dg_stream (windows overlap by construction)
  void* pin_wav = nullptr;        // pinned staging of dg_pipeline_call_host (B separate host windows -> one upload)
  size_t pin_wav_bytes = 0;
  std::unique_ptr<GatherPool> gather;   // worker threads of the host gather (created at the first dg_pipeline_call_host)
  DevBuf call_stream;                   // device image of the stream a dg_pipeline_call_host batch was cut from
  long long call_h2d_bytes = 0;         // bytes the last dg_pipeline_call_host uploaded
};

extern "C" int dg_pipeline_create(dg_seg* seg, dg_emb* emb, dg_cluster* clu, float gamma, float beta,
                                  int normalize_weights, dg_pipeline** out) {
  if (!seg || !emb || !clu || !out) {
    set_error("dg_pipeline_create: null handle");
    return DG_EINVAL;
  }
  if (seg->device != emb->device || seg->device != clu->device) {
    set_error("dg_pipeline_create: handles live on different devices");
    return DG_EINVAL;
  }
  if (clu->p.D != emb->D) {
    set_error("dg_pipeline_create: clustering dimension != embedding dimension");
    return DG_EINVAL;
  }
  std::unique_ptr<dg_pipeline> h(new dg_pipeline());
  h->seg = seg; h->emb = emb; h->clu = clu;
  h->gamma = gamma; h->beta = beta; h->normalize_weights = normalize_weights;
  DG_CUDA(cudaSetDevice(seg->device));
  DG_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  int lo = 0, hi = 0;
  DG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  DG_CUDA(cudaStreamCreateWithPriority(&h->s_seg, cudaStreamNonBlocking, hi));
  DG_CUDA(cudaStreamCreateWithPriority(&h->s_emb, cudaStreamNonBlocking, lo));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_start, cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_osp, cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_emb, cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_done, cudaEventDisableTiming));
  DG_CUDA(cudaStreamCreateWithPriority(&h->s_clu, cudaStreamNonBlocking, hi));
  DG_CUDA(cudaStreamCreateWithPriority(&h->s_seg2, cudaStreamNonBlocking, hi));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_osp2, cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_prep[0], cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_prep[1], cudaEventDisableTiming));
  DG_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
  DG_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 3; i++) {
    DG_CUDA(cudaEventCreateWithFlags(&h->e_h2d[i], cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&h->e_slot_done[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 2; i++) DG_CUDA(cudaEventCreateWithFlags(&h->e_lane_done[i], cudaEventDisableTiming));
  DG_CUDA(cudaEventRecord(h->e_emb, h->s_emb));   // so that the first step's wait on it is well defined
  *out = h.release();
  return DG_OK;
}

// segmentation chain on s_seg and embedding chain on s_emb, both starting after `start`; on return
// e_emb (recorded on s_emb) marks seg, osp and emb complete
// DG_CALL_TIMING=1: device time stamps of the sub-batches of dg_pipeline_call_host (diagnostic)
struct CallDiag {
  cudaEvent_t t0 = nullptr, up[3], prep[3], trunk[3], seg[3], emb[3], clu[3];
  int j = 0;
  void create() {
    if (t0) return;
    cudaEventCreate(&t0);
    for (int i = 0; i < 3; i++)
      for (cudaEvent_t* e : {&up[i], &prep[i], &trunk[i], &seg[i], &emb[i], &clu[i]}) cudaEventCreate(e);
  }
};
static thread_local CallDiag* g_diag = nullptr;
#define DG_DIAG(field, stream)                                        \
  do {                                                                \
    if (g_diag) cudaEventRecord(g_diag->field[g_diag->j], stream);    \
  } while (0)

static int pipeline_nets(dg_pipeline* h, const float* wav, int B, int S, int F, int K, float* seg, float* emb,
                         cudaEvent_t start, int lane = 0, int stream_hop = 0) {
  int rc;
  const Geom g = make_geom(S);
  // lane 0 / 1: segmentation stream, scratch set, OSP buffer and event of this step (consecutive pipelined steps
  // alternate, so step i+1's segmentation chain can start while step i's is still in its recurrence)
  cudaStream_t s_seg = lane ? h->s_seg2 : h->s_seg;
  DevBuf& osp = lane ? h->osp2 : h->osp;
  cudaEvent_t e_osp = lane ? h->e_osp2 : h->e_osp;
  if (osp.ensure((size_t)B * F * K * 4)) return DG_ECUDA;
  DG_CUDA(cudaStreamWaitEvent(s_seg, start, 0));
  DG_CUDA(cudaStreamWaitEvent(h->s_emb, start, 0));
  // another pipeline (or a block-level call) that used these model handles' scratch last: stream-ordered hand-over
  if ((rc = use_begin(h->seg->guard[lane], h, s_seg)) || (rc = use_begin(h->emb->guard, h, h->s_emb))) return rc;
  struct InPipeline {
    InPipeline() { g_in_pipeline = true; }
    ~InPipeline() { g_in_pipeline = false; }
  } in_pipeline;
  // waveform statistics + standardised 16-bit planes once, for both networks' SincNets
  static const bool sinc_simt = getenv("DG_SINC_SIMT") && getenv("DG_SINC_SIMT")[0] == '1';
  const SincPrep* shared = nullptr;
  if (!sinc_simt) {
    if ((rc = run_sinc_prep(h->prep[lane], wav, B, g, s_seg, stream_hop ? stream_hop : h->hop, stream_hop != 0))) return rc;
    DG_CUDA(cudaEventRecord(h->e_prep[lane], s_seg));
    DG_DIAG(prep, s_seg);
    DG_CUDA(cudaStreamWaitEvent(h->s_emb, h->e_prep[lane], 0));
    shared = &h->prep[lane];
  }
  // embedding trunk first in host order (low-priority stream, grid capped to the SMs the LSTM leaves free)
  int T = 0;
  const bool fuse_pool = h->emb->variant == 0 && K <= 4 && g.S2 >= 128 && pool_fusion_on();
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->seg->device);
    const int lstm_ctas = lstm_tc_ctas(B);
    g_sm_limit = sms - lstm_ctas > sms / 2 ? sms - lstm_ctas : 0;
    h->emb->shared_prep = shared;
    rc = emb_trunk(h->emb, wav, B, g, h->s_emb, &T, fuse_pool);
    h->emb->shared_prep = nullptr;
    g_sm_limit = 0;
    if (rc) return rc;
    DG_DIAG(trunk, h->s_emb);
  }
  h->seg->lane = lane;
  h->seg->shared_prep = shared;
  rc = dg_seg_forward(h->seg, wav, B, S, seg, s_seg);
  h->seg->lane = 0;
  h->seg->shared_prep = nullptr;
  if (rc) return rc;
  if ((rc = dg_osp(seg, B, F, K, h->gamma, h->beta, h->normalize_weights, osp.as<float>(), s_seg))) return rc;
  DG_CUDA(cudaEventRecord(e_osp, s_seg));
  if ((rc = use_end(h->seg->guard[lane], h, s_seg))) return rc;
  DG_DIAG(seg, s_seg);
  DG_CUDA(cudaStreamWaitEvent(h->s_emb, e_osp, 0));
  if ((rc = build_tables(h->emb, F, T, h->s_emb))) return rc;
  if (fuse_pool) {
    // TDNN5 needs the pooling weights: it runs here, after the segmentation of this step, fused with the pooling
    // (persistent grid capped like the trunk's: the other lane's recurrence may hold 2 x ceil(B/16) SMs at this point)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->seg->device);
    const int lstm_ctas = lstm_tc_ctas(B);
    g_sm_limit = sms - lstm_ctas > sms / 2 ? sms - lstm_ctas : 0;
    rc = emb_tdnn5_pool(h->emb, B, g, osp.as<float>(), F, K, T, h->s_emb);
    if (!rc) rc = emb_project(h->emb, B * K, 1, 1.f, emb, h->s_emb);
    g_sm_limit = 0;
    if (rc) return rc;
    DG_CUDA(cudaEventRecord(h->e_emb, h->s_emb));
    DG_DIAG(emb, h->s_emb);
    return use_end(h->emb->guard, h, h->s_emb);
  }
  if (h->emb->pooled.ensure((size_t)B * K * 2 * h->emb->pool_C * 4)) return DG_ECUDA;
  if ((rc = launch_stats_pool(h->emb->pool_x, B, g.S2, T, h->emb->pool_C, osp.as<float>(), F, K,
                              h->emb->idx0.as<int>(), h->emb->idx1.as<int>(), h->emb->lam1.as<float>(),
                              h->emb->pool_mode == 31 ? 1e-8f : 0.f, h->emb->pooled.as<float>(), h->s_emb,
                              h->emb->pool_item_pitch, h->emb->pool_row_pitch)))
    return rc;
  if ((rc = emb_project(h->emb, B * K, 1, 1.f, emb, h->s_emb))) return rc;
  DG_CUDA(cudaEventRecord(h->e_emb, h->s_emb));
  return use_end(h->emb->guard, h, h->s_emb);
}

extern "C" int dg_pipeline_set_hop(dg_pipeline* h, int hop_samples) {
  if (!h || hop_samples < 0) {
    set_error("dg_pipeline_set_hop: bad arguments");
    return DG_EINVAL;
  }
  h->hop = hop_samples;
  return DG_OK;
}

extern "C" int dg_pipeline_step(dg_pipeline* h, const float* wav, int B, int S, float* seg, float* emb, int32_t* map,
                                float* permuted, void* stream) {
  if (!h || !wav || !seg || !emb || !map || B < 1) {
    set_error("dg_pipeline_step: bad arguments");
    return DG_EINVAL;
  }
  if (h->outstanding) {
    set_error("dg_pipeline_step: submitted steps are outstanding; collect them first");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  if (h->osp.ensure((size_t)B * F * K * 4)) return DG_ECUDA;
  static const bool serial = getenv("DG_NO_OVERLAP") && getenv("DG_NO_OVERLAP")[0] == '1';
  if (serial || !use_tensor_cores()) {
    if ((rc = dg_seg_forward(h->seg, wav, B, S, seg, stream))) return rc;
    if ((rc = dg_osp(seg, B, F, K, h->gamma, h->beta, h->normalize_weights, h->osp.as<float>(), stream))) return rc;
    if ((rc = dg_emb_forward(h->emb, wav, h->osp.as<float>(), B, S, F, K, 1, 1.f, emb, stream))) return rc;
    return dg_cluster_step(h->clu, seg, emb, B, F, K, map, permuted, stream);
  }
  cudaStream_t st = (cudaStream_t)stream;
  DG_CUDA(cudaSetDevice(h->seg->device));
  DG_CUDA(cudaEventRecord(h->e_start, st));
  if ((rc = pipeline_nets(h, wav, B, S, F, K, seg, emb, h->e_start, 0, h->overlap_known ? h->hop : 0))) return rc;
  DG_CUDA(cudaStreamWaitEvent(h->s_clu, h->e_emb, 0));
  if ((rc = dg_cluster_step(h->clu, seg, emb, B, F, K, map, permuted, h->s_clu))) return rc;
  DG_CUDA(cudaEventRecord(h->e_done, h->s_clu));
  DG_CUDA(cudaStreamWaitEvent(st, h->e_done, 0));
  return DG_OK;
}

// ---- pipelined variants (up to three steps outstanding, two computing): the sequential clustering of step i and the host copies overlap the
//      networks of step i+1.  Per stream the chunk order is preserved: clustering runs on one stream.
static int pipeline_slot_prepare(dg_pipeline* h, int slot, int B, int S, int F, int K, bool host_in) {
  const int D = h->emb->D;
  if (h->osp.ensure((size_t)B * F * K * 4) || h->slot_seg[slot].ensure((size_t)B * F * K * 4) ||
      h->slot_emb[slot].ensure((size_t)B * K * D * 4) || h->slot_map[slot].ensure((size_t)B * K * 4) ||
      (host_in && h->slot_wav[slot].ensure((size_t)B * S * 4)))
    return DG_ECUDA;
  return DG_OK;
}

static const int DG_MAX_INFLIGHT = 3;

static int pipeline_submit_common(dg_pipeline* h, const float* wav_dev, int B, int S, int F, int K, int slot,
                                  cudaEvent_t start, int stream_hop = 0) {
  int rc;
  const int lane = (int)(h->next_step & 1);
  cudaStream_t s_lane = lane ? h->s_seg2 : h->s_seg;
  // the slot's previous occupant (three submits ago) and the lane's previous user (two submits ago) must be fully
  // clustered before their buffers are rewritten
  DG_CUDA(cudaStreamWaitEvent(s_lane, h->e_slot_done[slot], 0));
  DG_CUDA(cudaStreamWaitEvent(h->s_emb, h->e_slot_done[slot], 0));
  DG_CUDA(cudaStreamWaitEvent(s_lane, h->e_lane_done[lane], 0));
  DG_CUDA(cudaStreamWaitEvent(h->s_emb, h->e_lane_done[lane], 0));
  if ((rc = pipeline_nets(h, wav_dev, B, S, F, K, h->slot_seg[slot].as<float>(), h->slot_emb[slot].as<float>(), start,
                          lane, stream_hop)))
    return rc;
  // the lane's scratch (waveform planes, segmentation activations, OSP weights) is free as soon as this step's embeddings
  // exist -- the clustering reads only the slot buffers -- so the step after next may start before this one is clustered
  DG_CUDA(cudaEventRecord(h->e_lane_done[lane], h->s_emb));
  DG_CUDA(cudaStreamWaitEvent(h->s_clu, h->e_emb, 0));
  if ((rc = dg_cluster_step(h->clu, h->slot_seg[slot].as<float>(), h->slot_emb[slot].as<float>(), B, F, K,
                            h->slot_map[slot].as<int32_t>(), nullptr, h->s_clu)))
    return rc;
  DG_CUDA(cudaEventRecord(h->e_slot_done[slot], h->s_clu));
  DG_DIAG(clu, h->s_clu);
  h->slot_B[slot] = B;
  h->slot_S[slot] = S;
  h->next_step++;
  h->outstanding++;
  return DG_OK;
}

extern "C" int dg_pipeline_submit(dg_pipeline* h, const float* wav_dev, int B, int S, void* stream) {
  if (!h || !wav_dev || B < 1) {
    set_error("dg_pipeline_submit: bad arguments");
    return DG_EINVAL;
  }
  if (h->outstanding >= DG_MAX_INFLIGHT) {
    set_error("dg_pipeline_submit: three steps are already outstanding; collect one first");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  DG_CUDA(cudaSetDevice(h->seg->device));
  const int slot = (int)(h->next_step % 3);
  if ((rc = pipeline_slot_prepare(h, slot, B, S, F, K, false))) return rc;
  DG_CUDA(cudaEventRecord(h->e_start, (cudaStream_t)stream));
  return pipeline_submit_common(h, wav_dev, B, S, F, K, slot, h->e_start);
}

extern "C" int dg_pipeline_collect(dg_pipeline* h, const float** seg_dev, const float** emb_dev,
                                   const int32_t** map_dev, void* stream) {
  if (!h || h->outstanding < 1) {
    set_error("dg_pipeline_collect: nothing outstanding");
    return DG_EINVAL;
  }
  const int slot = (int)((h->next_step - h->outstanding) % 3);
  DG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->e_slot_done[slot], 0));
  if (seg_dev) *seg_dev = h->slot_seg[slot].as<float>();
  if (emb_dev) *emb_dev = h->slot_emb[slot].as<float>();
  if (map_dev) *map_dev = h->slot_map[slot].as<int32_t>();
  h->outstanding--;
  return DG_OK;
}

extern "C" int dg_pipeline_collect_copy(dg_pipeline* h, float* seg_dev, float* emb_dev, int32_t* map_dev,
                                        void* stream) {
  if (!h || h->outstanding < 1) {
    set_error("dg_pipeline_collect_copy: nothing outstanding");
    return DG_EINVAL;
  }
  const int slot = (int)((h->next_step - h->outstanding) % 3);
  int F = 0, K = 0;
  const int B = h->slot_B[slot], D = h->emb->D;
  dg_seg_dims(h->seg, h->slot_S[slot], &F, &K);
  cudaStream_t st = (cudaStream_t)stream;
  DG_CUDA(cudaStreamWaitEvent(st, h->e_slot_done[slot], 0));
  if (seg_dev) DG_CUDA(cudaMemcpyAsync(seg_dev, h->slot_seg[slot].p, (size_t)B * F * K * 4, cudaMemcpyDeviceToDevice, st));
  if (emb_dev) DG_CUDA(cudaMemcpyAsync(emb_dev, h->slot_emb[slot].p, (size_t)B * K * D * 4, cudaMemcpyDeviceToDevice, st));
  if (map_dev) DG_CUDA(cudaMemcpyAsync(map_dev, h->slot_map[slot].p, (size_t)B * K * 4, cudaMemcpyDeviceToDevice, st));
  h->outstanding--;
  return DG_OK;
}

extern "C" int dg_pipeline_submit_host(dg_pipeline* h, const float* wav_host, int B, int S) {
  if (!h || !wav_host || B < 1) {
    set_error("dg_pipeline_submit_host: bad arguments");
    return DG_EINVAL;
  }
  if (h->outstanding >= DG_MAX_INFLIGHT) {
    set_error("dg_pipeline_submit_host: three steps are already outstanding; collect one first");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  DG_CUDA(cudaSetDevice(h->seg->device));
  const int slot = (int)(h->next_step % 3);
  if ((rc = pipeline_slot_prepare(h, slot, B, S, F, K, true))) return rc;
  DG_CUDA(cudaStreamWaitEvent(h->s_h2d, h->e_slot_done[slot], 0));
  DG_CUDA(cudaMemcpyAsync(h->slot_wav[slot].p, wav_host, (size_t)B * S * 4, cudaMemcpyHostToDevice, h->s_h2d));
  DG_CUDA(cudaEventRecord(h->e_h2d[slot], h->s_h2d));
  return pipeline_submit_common(h, h->slot_wav[slot].as<float>(), B, S, F, K, slot, h->e_h2d[slot]);
}

extern "C" int dg_pipeline_collect_host(dg_pipeline* h, float* seg_host, float* emb_host, int32_t* map_host) {
  if (!h || h->outstanding < 1) {
    set_error("dg_pipeline_collect_host: nothing outstanding");
    return DG_EINVAL;
  }
  const int slot = (int)((h->next_step - h->outstanding) % 3);
  int F = 0, K = 0;
  const int B = h->slot_B[slot], D = h->emb->D;
  dg_seg_dims(h->seg, h->slot_S[slot], &F, &K);
  DG_CUDA(cudaStreamWaitEvent(h->s_d2h, h->e_slot_done[slot], 0));
  if (seg_host)
    DG_CUDA(cudaMemcpyAsync(seg_host, h->slot_seg[slot].p, (size_t)B * F * K * 4, cudaMemcpyDeviceToHost, h->s_d2h));
  if (emb_host)
    DG_CUDA(cudaMemcpyAsync(emb_host, h->slot_emb[slot].p, (size_t)B * K * D * 4, cudaMemcpyDeviceToHost, h->s_d2h));
  if (map_host)
    DG_CUDA(cudaMemcpyAsync(map_host, h->slot_map[slot].p, (size_t)B * K * 4, cudaMemcpyDeviceToHost, h->s_d2h));
  DG_CUDA(cudaStreamSynchronize(h->s_d2h));
  h->outstanding--;
  return DG_OK;
}

extern "C" int dg_pipeline_step_host(dg_pipeline* h, const float* wav_host, int B, int S, float* seg_host,
                                     float* emb_host, int32_t* map_host, float* permuted_host) {
  if (!h || !wav_host || B < 1) {
    set_error("dg_pipeline_step_host: bad arguments");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  const int D = h->emb->D, M = h->clu->p.M;
  DG_CUDA(cudaSetDevice(h->seg->device));
  if (h->wav.ensure((size_t)B * S * 4) || h->segd.ensure((size_t)B * F * K * 4) ||
      h->embd.ensure((size_t)B * K * D * 4) || h->mapd.ensure((size_t)B * K * 4) ||
      (permuted_host && h->permd.ensure((size_t)B * F * M * 4)))
    return DG_ECUDA;
  DG_CUDA(cudaMemcpyAsync(h->wav.p, wav_host, (size_t)B * S * 4, cudaMemcpyHostToDevice, h->st));
  if ((rc = dg_pipeline_step(h, h->wav.as<float>(), B, S, h->segd.as<float>(), h->embd.as<float>(),
                             h->mapd.as<int32_t>(), permuted_host ? h->permd.as<float>() : nullptr, h->st)))
    return rc;
  if (seg_host) DG_CUDA(cudaMemcpyAsync(seg_host, h->segd.p, (size_t)B * F * K * 4, cudaMemcpyDeviceToHost, h->st));
  if (emb_host) DG_CUDA(cudaMemcpyAsync(emb_host, h->embd.p, (size_t)B * K * D * 4, cudaMemcpyDeviceToHost, h->st));
  if (map_host) DG_CUDA(cudaMemcpyAsync(map_host, h->mapd.p, (size_t)B * K * 4, cudaMemcpyDeviceToHost, h->st));
  if (permuted_host)
    DG_CUDA(cudaMemcpyAsync(permuted_host, h->permd.p, (size_t)B * F * M * 4, cudaMemcpyDeviceToHost, h->st));
  DG_CUDA(cudaStreamSynchronize(h->st));
  return DG_OK;
}

// ======================================================================== device-side audio stream
// rearrange_audio_stream (reference src/diart/operators.py:44-100) on the device: the host pushes each sample ONCE
// (8 000 new samples per chunk instead of the 80 000 of a stacked window: 8.2 MB instead of 82 MB per 256-chunk step),
// windows are formed from a circular ring in HBM.
struct dg_stream {
  int device = 0, S = 0, hop = 0, C = 0;
  long long wpos = 0, rpos = 0;          // absolute sample counters: pushed / start of the next window
  DevBuf ring;
  float* pin = nullptr;                  // pinned mirror of the ring (staging for the uploads)
  cudaStream_t st = nullptr;             // uploads
  cudaEvent_t e_up = nullptr, e_read = nullptr;
  // uploads still reading the pinned mirror: (first absolute sample, event); a region of the mirror is rewritten only
  // after the upload that last used it has completed
  std::deque<std::pair<long long, cudaEvent_t>> inflight;
  std::vector<cudaEvent_t> spare;
  ~dg_stream() {
    for (auto& e : inflight) cudaEventDestroy(e.second);
    for (auto& e : spare) cudaEventDestroy(e);
    if (pin) cudaFreeHost(pin);
    if (st) cudaStreamDestroy(st);
    if (e_up) cudaEventDestroy(e_up);
    if (e_read) cudaEventDestroy(e_read);
  }
};

extern "C" int dg_stream_create(int chunk_samples, int step_samples, int max_windows, int device, dg_stream** out) {
  if (!out || chunk_samples < 4 || step_samples < 4 || chunk_samples % 4 || step_samples % 4 || max_windows < 1 ||
      step_samples > chunk_samples) {
    set_error("dg_stream_create: chunk and step must be positive multiples of 4 samples, step <= chunk");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(device));
  std::unique_ptr<dg_stream> h(new dg_stream());
  h->device = device; h->S = chunk_samples; h->hop = step_samples;
  // room for the windows being read, a full batch being uploaded meanwhile, and the overlap tail
  h->C = ((chunk_samples + 2 * max_windows * step_samples + 1023) / 1024) * 1024;
  if (h->ring.ensure((size_t)h->C * 4)) return DG_ECUDA;
  DG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->pin), (size_t)h->C * 4, cudaHostAllocDefault));
  DG_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_up, cudaEventDisableTiming));
  DG_CUDA(cudaEventCreateWithFlags(&h->e_read, cudaEventDisableTiming));
  DG_CUDA(cudaEventRecord(h->e_read, h->st));
  *out = h.release();
  return DG_OK;
}

extern "C" int dg_stream_destroy(dg_stream* h) {
  delete h;
  return DG_OK;
}

extern "C" int dg_stream_reset(dg_stream* h) {
  if (!h) return DG_EINVAL;
  DG_CUDA(cudaSetDevice(h->device));
  DG_CUDA(cudaStreamSynchronize(h->st));
  h->wpos = h->rpos = 0;
  for (auto& e : h->inflight) h->spare.push_back(e.second);
  h->inflight.clear();
  return DG_OK;
}

// complete windows that have been pushed but not yet consumed
extern "C" int dg_stream_available(const dg_stream* h) {
  if (!h) return 0;
  const long long have = h->wpos - h->rpos;
  return have < h->S ? 0 : (int)((have - h->S) / h->hop + 1);
}

// appends n samples (host memory, any kind) to the stream; returns once they are staged (the upload is asynchronous)
extern "C" int dg_stream_push_host(dg_stream* h, const float* samples, int n) {
  if (!h || !samples || n < 0) {
    set_error("dg_stream_push_host: bad arguments");
    return DG_EINVAL;
  }
  if (h->wpos + n - h->rpos > h->C) {
    set_error("dg_stream_push_host: ring full (" + std::to_string(h->wpos - h->rpos) + " samples buffered, capacity " +
              std::to_string(h->C) + "): consume windows first");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  // samples older than rpos may be overwritten: uploads are ordered after the last kernel that read the ring
  DG_CUDA(cudaStreamWaitEvent(h->st, h->e_read, 0));
  // the mirror region [wpos, wpos + n) was last used by the uploads of samples one lap earlier: wait for those
  while (!h->inflight.empty() && h->inflight.front().first < h->wpos + n - h->C) {
    DG_CUDA(cudaEventSynchronize(h->inflight.front().second));
    h->spare.push_back(h->inflight.front().second);
    h->inflight.pop_front();
  }
  int done = 0;
  while (done < n) {
    const int at = (int)((h->wpos + done) % h->C);
    const int len = std::min(n - done, h->C - at);
    memcpy(h->pin + at, samples + done, (size_t)len * 4);
    DG_CUDA(cudaMemcpyAsync(h->ring.as<float>() + at, h->pin + at, (size_t)len * 4, cudaMemcpyHostToDevice, h->st));
    done += len;
  }
  cudaEvent_t ev;
  if (!h->spare.empty()) {
    ev = h->spare.back();
    h->spare.pop_back();
  } else {
    DG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  DG_CUDA(cudaEventRecord(ev, h->st));
  h->inflight.emplace_back(h->wpos, ev);
  h->wpos += n;
  DG_CUDA(cudaEventRecord(h->e_up, h->st));
  return DG_OK;
}

// materialises the next B windows as a dense [B, S] batch on `st` and advances the stream by B steps
static int stream_expand(dg_stream* h, int B, float* wav_dev, cudaStream_t st) {
  if (dg_stream_available(h) < B) {
    set_error("dg_stream: " + std::to_string(B) + " windows requested, " + std::to_string(dg_stream_available(h)) + " available");
    return DG_EINVAL;
  }
  if (h->rpos % 4) {
    set_error("dg_stream: window start is not 16-byte aligned");
    return DG_EINVAL;
  }
  DG_CUDA(cudaStreamWaitEvent(st, h->e_up, 0));
  int rc;
  if ((rc = launch_expand_windows(h->ring.as<float>(), h->rpos, h->C, h->hop, h->S, B, wav_dev, st))) return rc;
  DG_CUDA(cudaEventRecord(h->e_read, st));
  h->rpos += (long long)B * h->hop;
  return 0;
}

extern "C" int dg_stream_windows(dg_stream* h, int B, float* wav_dev, void* stream) {
  if (!h || !wav_dev || B < 1) {
    set_error("dg_stream_windows: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  return stream_expand(h, B, wav_dev, (cudaStream_t)stream);
}

// =============================================================================== device post-path
// DelayedAggregation (hamming, loose) + Binarize of reference diarization.py:205-232 on the device (post.cu).  The handle keeps
// the scores and speaker maps of the last `num_windows - 1` chunks (the reference's pred_buffer) on the device.
struct dg_post {
  int device = 0, F = 0, K = 0, M = 0, nw = 1;
  double tau = 0.5;
  DevBuf hamming, hist_seg[2], hist_map[2], plan, header, turns, total;
  int cur = 0, n_hist = 0, cap_B = 0;
  int turn_cap = 0;
  void* pin = nullptr;            // pinned staging: plan in, header + total + turn prefix out
  size_t pin_bytes = 0;
  ~dg_post() {
    if (pin) cudaFreeHost(pin);
  }
};

static const int DG_POST_PREFIX = 16384;   // turns copied back together with the header (one D2H in the common case)

extern "C" int dg_post_create(int frames, int local_speakers, int max_speakers, int num_windows, const double* hamming_host,
                              double tau, int device, dg_post** out) {
  if (!out || !hamming_host || frames < 1 || frames > 1023 || local_speakers < 1 || max_speakers < 1 || max_speakers > 64 ||
      num_windows < 1 || num_windows > 256) {
    set_error("dg_post_create: need 1 <= frames <= 1023, 1 <= max_speakers <= 64, 1 <= num_windows <= 256");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(device));
  std::unique_ptr<dg_post> h(new dg_post());
  h->device = device; h->F = frames; h->K = local_speakers; h->M = max_speakers; h->nw = num_windows; h->tau = tau;
  if (h->hamming.ensure((size_t)frames * 8) || h->total.ensure(16)) return DG_ECUDA;
  DG_CUDA(cudaMemcpy(h->hamming.p, hamming_host, (size_t)frames * 8, cudaMemcpyHostToDevice));
  const size_t hs = (size_t)std::max(1, num_windows - 1);
  for (int i = 0; i < 2; i++)
    if (h->hist_seg[i].ensure(hs * frames * local_speakers * 4) || h->hist_map[i].ensure(hs * local_speakers * 4)) return DG_ECUDA;
  *out = h.release();
  return DG_OK;
}

extern "C" int dg_post_reset(dg_post* h) {
  if (!h) return DG_EINVAL;
  h->n_hist = 0;
  return DG_OK;
}

extern "C" int dg_post_destroy(dg_post* h) {
  delete h;
  return DG_OK;
}

static int post_ensure(dg_post* h, int B) {
  if (B <= h->cap_B) return 0;
  const int stride = 4 + h->nw;
  // worst case: every second frame of every speaker starts a turn
  h->turn_cap = B * h->M * ((h->F + 1) / 2);
  if (h->plan.ensure((size_t)B * stride * 4) || h->header.ensure((size_t)B * 16 + 16) ||
      h->turns.ensure((size_t)h->turn_cap * 4))
    return DG_ECUDA;
  const size_t need = (size_t)B * stride * 4 + (size_t)B * 16 + 16 + (size_t)DG_POST_PREFIX * 4;
  if (need > h->pin_bytes) {
    if (h->pin) cudaFreeHost(h->pin);
    h->pin = nullptr;
    DG_CUDA(cudaHostAlloc(&h->pin, need, cudaHostAllocDefault));
    h->pin_bytes = need;
  }
  h->cap_B = B;
  return 0;
}

// enqueues plan upload, aggregation + binarisation + run-length kernel, history update and the D2H of the results on `st`
static int post_enqueue(dg_post* h, const float* seg_dev, const int32_t* map_dev, int B, const int32_t* plan_host,
                        cudaStream_t st) {
  int rc;
  if ((rc = post_ensure(h, B))) return rc;
  const int stride = 4 + h->nw;
  unsigned char* pin = reinterpret_cast<unsigned char*>(h->pin);
  const size_t plan_bytes = (size_t)B * stride * 4;
  memcpy(pin, plan_host, plan_bytes);
  DG_CUDA(cudaMemcpyAsync(h->plan.p, pin, plan_bytes, cudaMemcpyHostToDevice, st));
  DG_CUDA(cudaMemsetAsync(h->total.p, 0, 4, st));
  if ((rc = launch_post(seg_dev, map_dev, h->hist_seg[h->cur].as<float>(), h->hist_map[h->cur].as<int32_t>(), h->n_hist, B,
                        h->F, h->K, h->M, h->nw, h->plan.as<int32_t>(), stride, h->hamming.as<double>(), h->tau,
                        h->header.as<int32_t>(), h->turns.as<uint32_t>(), h->turn_cap, h->total.as<unsigned int>(), st)))
    return rc;
  const int keep = std::min(h->nw - 1, h->n_hist + B);
  if (keep > 0) {
    if ((rc = launch_post_history(seg_dev, map_dev, h->hist_seg[h->cur].as<float>(), h->hist_map[h->cur].as<int32_t>(),
                                  h->n_hist, B, h->F, h->K, keep, h->hist_seg[h->cur ^ 1].as<float>(),
                                  h->hist_map[h->cur ^ 1].as<int32_t>(), st)))
      return rc;
    h->cur ^= 1;
  }
  h->n_hist = keep;
  unsigned char* out = pin + plan_bytes;
  DG_CUDA(cudaMemcpyAsync(out, h->header.p, (size_t)B * 16, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaMemcpyAsync(out + (size_t)B * 16, h->total.p, 4, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaMemcpyAsync(out + (size_t)B * 16 + 16, h->turns.p, (size_t)std::min(DG_POST_PREFIX, h->turn_cap) * 4,
                          cudaMemcpyDeviceToHost, st));
  return 0;
}

// after `st` has been synchronised: hands the results to the caller
static int post_finish(dg_post* h, int B, int32_t* header_host, uint32_t* turns_host, int turn_cap_host, int* n_turns,
                       cudaStream_t st) {
  const int stride = 4 + h->nw;
  unsigned char* out = reinterpret_cast<unsigned char*>(h->pin) + (size_t)B * stride * 4;
  unsigned int total = 0;
  memcpy(&total, out + (size_t)B * 16, 4);
  if (n_turns) *n_turns = (int)total;
  memcpy(header_host, out, (size_t)B * 16);
  if ((int)total > turn_cap_host) {
    set_error("dg_post_step: turn buffer too small (" + std::to_string(total) + " turns)");
    return DG_EINVAL;
  }
  const unsigned int pre = std::min<unsigned int>(total, (unsigned int)DG_POST_PREFIX);
  memcpy(turns_host, out + (size_t)B * 16 + 16, (size_t)pre * 4);
  if (total > pre) {
    DG_CUDA(cudaMemcpyAsync(turns_host + pre, h->turns.as<uint32_t>() + pre, (size_t)(total - pre) * 4,
                            cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
  }
  return DG_OK;
}

extern "C" int dg_post_step(dg_post* h, const float* seg_dev, const int32_t* map_dev, int B, const int32_t* plan_host,
                            int32_t* header_host, uint32_t* turns_host, int turn_cap_host, int* n_turns, void* stream) {
  if (!h || !seg_dev || !map_dev || !plan_host || !header_host || !turns_host || B < 1) {
    set_error("dg_post_step: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = post_enqueue(h, seg_dev, map_dev, B, plan_host, st))) return rc;
  DG_CUDA(cudaStreamSynchronize(st));
  return post_finish(h, B, header_host, turns_host, turn_cap_host, n_turns, st);
}

// ---- the whole body of SpeakerDiarization.__call__ (reference diarization.py:172-232) in one call: B separate host windows
//      (as rearrange_audio_stream emits them) are gathered into pinned staging by worker threads while earlier rows are
//      already on their way to the device, then fused step + post-path, one D2H of the turn list.
static int upload_rows(dg_pipeline* h, const float* const* rows, int B, int S, float* pin, float* dst_dev, cudaStream_t st) {
  const int R = 4;                                    // rows per work item (1.3 MB at S = 80000)
  const int items = (B + R - 1) / R;
  if (!h->gather) {
    int n = (int)std::thread::hardware_concurrency();
    static const int env_threads = getenv("DG_GATHER_THREADS") ? atoi(getenv("DG_GATHER_THREADS")) : 0;
    n = env_threads > 0 ? env_threads : std::max(1, std::min(n - 2, 24));
    h->gather.reset(new GatherPool(n));
  }
  std::vector<std::atomic<int>> done(items);
  for (auto& d : done) d.store(0, std::memory_order_relaxed);
  std::atomic<int> next{0};
  h->gather->start([&]() {
    for (;;) {
      const int it = next.fetch_add(1, std::memory_order_relaxed);
      if (it >= items) return;
      const int r0 = it * R, r1 = std::min(B, r0 + R);
      for (int r = r0; r < r1; r++) memcpy(pin + (size_t)r * S, rows[r], (size_t)S * 4);
      done[it].store(1, std::memory_order_release);
    }
  });
  // the calling thread forwards finished items, in order, in runs of up to 8 (~10 MB per copy)
  cudaError_t err = cudaSuccess;
  int sent = 0;
  while (sent < items) {
    int upto = sent;
    while (upto < items && upto - sent < 8 && done[upto].load(std::memory_order_acquire)) upto++;
    if (upto == sent) {
      std::this_thread::yield();
      continue;
    }
    const int r0 = sent * R, r1 = std::min(B, upto * R);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(dst_dev + (size_t)r0 * S, pin + (size_t)r0 * S, (size_t)(r1 - r0) * S * 4, cudaMemcpyHostToDevice, st);
    sent = upto;
  }
  h->gather->wait();          // (`next` and `done` live on this frame)
  DG_CUDA(err);
  return 0;
}

// Windows that are consecutive hops of ONE stream -- what the reference's rearrange_audio_stream emits (operators.py:44-100) --
// share S - hop samples with their neighbour.  The workers compare every window with its predecessor (memcmp of the shared
// samples, exact) and pack the `hop` new samples of each into the pinned stream image; the caller then uploads
// S + (B - 1) hop samples instead of B S and forms the windows on the device.  Returns 1 if windows [r0, r0 + nb) continue the
// stream (pin_stream[0 .. S + (r0 + nb - 1) hop) is then valid), 0 if some window does not (the caller falls back to the
// full gather for this and the following sub-batches).
static int pack_stream_rows(dg_pipeline* h, const float* const* rows, int r0, int nb, int S, int hop, float* pin_stream) {
  if (!h->gather) {
    int n = (int)std::thread::hardware_concurrency();
    static const int env_threads = getenv("DG_GATHER_THREADS") ? atoi(getenv("DG_GATHER_THREADS")) : 0;
    n = env_threads > 0 ? env_threads : std::max(1, std::min(n - 2, 24));
    h->gather.reset(new GatherPool(n));
  }
  std::atomic<int> next{r0}, bad{0};
  h->gather->start([&]() {
    for (;;) {
      const int r = next.fetch_add(1, std::memory_order_relaxed);
      if (r >= r0 + nb || bad.load(std::memory_order_relaxed)) return;
      if (r == 0) {
        memcpy(pin_stream, rows[0], (size_t)S * 4);
      } else if (memcmp(rows[r - 1] + hop, rows[r], (size_t)(S - hop) * 4) != 0) {
        bad.store(1, std::memory_order_relaxed);
      } else {
        memcpy(pin_stream + (size_t)S + (size_t)(r - 1) * hop, rows[r] + (S - hop), (size_t)hop * 4);
      }
    }
  });
  h->gather->wait();
  return bad.load() ? 0 : 1;
}

extern "C" int dg_pipeline_call_host(dg_pipeline* h, dg_post* post, const float* const* rows_host, int B, int S,
                                     const int32_t* plan_host, int32_t* header_host, uint32_t* turns_host, int turn_cap_host,
                                     int* n_turns, float* seg_host, int32_t* map_host) {
  if (!h || !post || !rows_host || !plan_host || !header_host || !turns_host || B < 1) {
    set_error("dg_pipeline_call_host: bad arguments");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  if (F != post->F || K != post->K || h->clu->p.M != post->M || post->device != h->seg->device) {
    set_error("dg_pipeline_call_host: post handle was created for other dimensions");
    return DG_EINVAL;
  }
  const int D = h->emb->D;
  (void)D;
  if (h->outstanding) {
    set_error("dg_pipeline_call_host: submitted steps are outstanding; collect them first");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->seg->device));
  // DG_CALL_TIMING=1: host wall-clock phases of the call on stderr (diagnostic)
  static const bool call_timing = getenv("DG_CALL_TIMING") && getenv("DG_CALL_TIMING")[0] == '1';
  const auto tc0 = std::chrono::steady_clock::now();
  static thread_local CallDiag diag;
  if (call_timing) {
    diag.create();
    cudaEventRecord(diag.t0, h->s_h2d);
    g_diag = &diag;
  }
  if (h->segd.ensure((size_t)B * F * K * 4) || h->mapd.ensure((size_t)B * K * 4)) return DG_ECUDA;
  const size_t bytes = (size_t)B * S * 4;
  if (bytes > h->pin_wav_bytes) {
    if (h->pin_wav) cudaFreeHost(h->pin_wav);
    h->pin_wav = nullptr;
    DG_CUDA(cudaHostAlloc(&h->pin_wav, bytes, cudaHostAllocDefault));
    h->pin_wav_bytes = bytes;
  }
  // The batch runs as up to three sub-batches through the pipelined machinery (dg_pipeline_submit_host): the upload of
  // sub-batch j+1 and its front end overlap the recurrence of sub-batch j; clustering stays in chunk order on its one stream,
  // so the result is exactly that of one step over the whole batch.  DG_CALL_SPLIT = 1..3 (default 2 from 128 windows on).
  // sub-batch sizes: DG_CALL_PLAN="n1,n2[,n3]" (windows; must add up to B) or DG_CALL_SPLIT = 1..3 equal parts; default
  // from 192 windows on: three parts -- a short first one so that the device starts early and a short last one, because
  // its dependent chain (1172 recurrence steps + its share of the clustering) is what the caller waits for at the end
  int plan[DG_MAX_INFLIGHT] = {B, 0, 0}, ns = 1;
  {
    static const int split_env = getenv("DG_CALL_SPLIT") ? atoi(getenv("DG_CALL_SPLIT")) : 0;
    static const std::string plan_env = getenv("DG_CALL_PLAN") ? getenv("DG_CALL_PLAN") : "";
    int vals[DG_MAX_INFLIGHT] = {0, 0, 0}, nv = 0, sum = 0;
    if (!plan_env.empty()) {
      std::stringstream ss(plan_env);
      std::string tok;
      while (nv < DG_MAX_INFLIGHT && std::getline(ss, tok, ',')) {
        vals[nv] = atoi(tok.c_str());
        sum += vals[nv];
        if (vals[nv++] < 1) sum = -1 << 20;
      }
    }
    if (nv > 0 && sum == B) {
      ns = nv;
      for (int j = 0; j < nv; j++) plan[j] = vals[j];
    } else if (split_env > 0) {
      ns = std::max(1, std::min({split_env, DG_MAX_INFLIGHT, B / 8 > 0 ? B / 8 : 1}));
      const int Bs = (B + ns - 1) / ns;
      for (int j = 0, left = B; j < ns; j++, left -= Bs) plan[j] = std::min(Bs, left);
      while (ns > 1 && plan[ns - 1] <= 0) ns--;
    } else if (B >= 192) {
      ns = 3;
      plan[0] = (B * 5 / 16 + 7) / 8 * 8;
      plan[2] = (B * 4 / 16 + 7) / 8 * 8;
      plan[1] = B - plan[0] - plan[2];
    } else if (B >= 64) {
      ns = 2;
      plan[0] = (B / 2 + 7) / 8 * 8;
      plan[1] = B - plan[0];
    }
  }
  // consecutive windows of one stream (hop known from dg_pipeline_set_hop): verified on the host, uploaded once (see
  // pack_stream_rows); DG_CALL_NO_DEDUP=1 = always the full gather
  static const bool no_dedup = getenv("DG_CALL_NO_DEDUP") && getenv("DG_CALL_NO_DEDUP")[0] == '1';
  const int hop = h->hop;
  bool as_stream = !no_dedup && hop > 0 && hop < S && hop % 4 == 0 && S % 4 == 0 && B >= 2;
  const size_t stream_len = (size_t)S + (size_t)(B - 1) * (hop > 0 ? hop : 0);
  if (as_stream && h->call_stream.ensure((stream_len + 64) * 4)) return DG_ECUDA;
  float* pin = reinterpret_cast<float*>(h->pin_wav);
  h->call_h2d_bytes = 0;
  int slots[DG_MAX_INFLIGHT], nbs[DG_MAX_INFLIGHT];
  for (int j = 0, r0 = 0; j < ns; r0 += plan[j], j++) {
    const int nb = plan[j];
    const int slot = (int)(h->next_step % 3);
    if ((rc = pipeline_slot_prepare(h, slot, nb, S, F, K, true))) return rc;
    DG_CUDA(cudaStreamWaitEvent(h->s_h2d, h->e_slot_done[slot], 0));
    if (as_stream && !pack_stream_rows(h, rows_host, r0, nb, S, hop, pin)) as_stream = false;
    int stream_hop = 0;
    if (as_stream) {
      // the samples this sub-batch adds to the device image of the stream, then its windows from that image
      const size_t lo = r0 == 0 ? 0 : (size_t)S + (size_t)(r0 - 1) * hop, hi = (size_t)S + (size_t)(r0 + nb - 1) * hop;
      DG_CUDA(cudaMemcpyAsync(h->call_stream.as<float>() + lo, pin + lo, (hi - lo) * 4, cudaMemcpyHostToDevice, h->s_h2d));
      h->call_h2d_bytes += (long long)(hi - lo) * 4;
      const long long cap = (long long)((stream_len + 3) / 4 * 4 + 4);     // linear image: the ring index never wraps
      if ((rc = launch_expand_windows(h->call_stream.as<float>(), (long long)r0 * hop, (int)cap, hop, S, nb,
                                      h->slot_wav[slot].as<float>(), h->s_h2d)))
        return rc;
      stream_hop = hop;
    } else {
      // (after a failed stream check the pinned buffer is reused as the [B, S] staging: earlier sub-batches are already on the device)
      if (h->call_h2d_bytes) DG_CUDA(cudaStreamSynchronize(h->s_h2d));
      if ((rc = upload_rows(h, rows_host + r0, nb, S, pin + (size_t)r0 * S, h->slot_wav[slot].as<float>(), h->s_h2d))) return rc;
      h->call_h2d_bytes += (long long)nb * S * 4;
    }
    DG_CUDA(cudaEventRecord(h->e_h2d[slot], h->s_h2d));
    if (g_diag) g_diag->j = j;
    DG_DIAG(up, h->s_h2d);
    if ((rc = pipeline_submit_common(h, h->slot_wav[slot].as<float>(), nb, S, F, K, slot, h->e_h2d[slot], stream_hop))) return rc;
    slots[j] = slot;
    nbs[j] = nb;
  }
  for (int j = 0, r0 = 0; j < ns; r0 += nbs[j], j++) {     // gather the sub-batches' scores / maps, in order
    DG_CUDA(cudaStreamWaitEvent(h->st, h->e_slot_done[slots[j]], 0));
    DG_CUDA(cudaMemcpyAsync(h->segd.as<float>() + (size_t)r0 * F * K, h->slot_seg[slots[j]].p, (size_t)nbs[j] * F * K * 4,
                            cudaMemcpyDeviceToDevice, h->st));
    DG_CUDA(cudaMemcpyAsync(h->mapd.as<int32_t>() + (size_t)r0 * K, h->slot_map[slots[j]].p, (size_t)nbs[j] * K * 4,
                            cudaMemcpyDeviceToDevice, h->st));
    h->outstanding--;
  }
  const auto tc1 = std::chrono::steady_clock::now();
  if ((rc = post_enqueue(post, h->segd.as<float>(), h->mapd.as<int32_t>(), B, plan_host, h->st))) return rc;
  if (seg_host) DG_CUDA(cudaMemcpyAsync(seg_host, h->segd.p, (size_t)B * F * K * 4, cudaMemcpyDeviceToHost, h->st));
  if (map_host) DG_CUDA(cudaMemcpyAsync(map_host, h->mapd.p, (size_t)B * K * 4, cudaMemcpyDeviceToHost, h->st));
  DG_CUDA(cudaStreamSynchronize(h->st));
  const auto tc2 = std::chrono::steady_clock::now();
  rc = post_finish(post, B, header_host, turns_host, turn_cap_host, n_turns, h->st);
  g_diag = nullptr;
  if (call_timing) {
    static int shown = 0;
    if (rc == 0 && shown++ % 4 == 3) {
      for (int j = 0; j < ns; j++) {
        float t[6] = {0, 0, 0, 0, 0, 0};
        cudaEvent_t ev[6] = {diag.up[j], diag.prep[j], diag.trunk[j], diag.seg[j], diag.emb[j], diag.clu[j]};
        for (int q = 0; q < 6; q++) cudaEventElapsedTime(&t[q], diag.t0, ev[q]);
        fprintf(stderr, "  sub-batch %d (%d windows), ms after entry: uploaded %.2f | front end %.2f | embedding trunk %.2f | segmentation + "
                        "OSP %.2f | embeddings %.2f | clustered %.2f\n", j, nbs[j], t[0], t[1], t[2], t[3], t[4], t[5]);
      }
    }
    static double acc[3] = {0, 0, 0};
    static int calls = 0;
    const auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count(); };
    acc[0] += ms(tc0, tc1);
    acc[1] += ms(tc1, tc2);
    acc[2] += ms(tc2, std::chrono::steady_clock::now());
    if (++calls % 4 == 0) {
      fprintf(stderr, "dg_pipeline_call_host (B=%d, %d sub-batches): gather + upload + enqueue %.2f ms | wait for the device %.2f ms | "
                      "turn list %.2f ms (mean of 4 calls)\n", B, ns, acc[0] / 4, acc[1] / 4, acc[2] / 4);
      acc[0] = acc[1] = acc[2] = 0;
    }
  }
  return rc;
}

extern "C" int64_t dg_pipeline_last_call_h2d_bytes(const dg_pipeline* h) { return h ? (int64_t)h->call_h2d_bytes : 0; }

// ---- shared-identity mode (BASELINE config 5) without leaving the pipelined flow.  After dg_pipeline_submit*:
//   dg_pipeline_identity_export  enqueues the export of this rank's centroid changes behind the clustering of every submitted
//                                step (clustering stream) and makes `stream` wait for it -> the caller all-gathers the records
//   dg_pipeline_identity_merge   makes the clustering stream wait for `stream` (the all-gather), merges all ranks' records and
//                                relabels the speaker maps of the steps clustered since the previous merge (still on the device)
// The clustering of the NEXT submitted step is ordered behind the merge, exactly as in the one-step-at-a-time protocol; only
// the networks of the next steps overlap the exchange.  Call the pair once after every submit, before collecting that step.
extern "C" int dg_pipeline_identity_export(dg_pipeline* h, double* record_dev, void* stream) {
  if (!h || !record_dev) {
    set_error("dg_pipeline_identity_export: bad arguments");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->seg->device));
  if (!h->e_ident) {
    DG_CUDA(cudaEventCreateWithFlags(&h->e_ident, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&h->e_ident_in, cudaEventDisableTiming));
  }
  int rc;
  if ((rc = dg_cluster_export_delta(h->clu, record_dev, h->s_clu))) return rc;
  DG_CUDA(cudaEventRecord(h->e_ident, h->s_clu));
  DG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->e_ident, 0));
  return DG_OK;
}

extern "C" int dg_pipeline_identity_merge(dg_pipeline* h, const double* records_dev, int world, int rank, void* stream) {
  if (!h || !records_dev || world < 1 || rank < 0 || rank >= world || !h->e_ident) {
    set_error("dg_pipeline_identity_merge: bad arguments (or no export before it)");
    return DG_EINVAL;
  }
  DG_CUDA(cudaSetDevice(h->seg->device));
  DG_CUDA(cudaEventRecord(h->e_ident_in, (cudaStream_t)stream));
  DG_CUDA(cudaStreamWaitEvent(h->s_clu, h->e_ident_in, 0));
  int rc;
  long long first = h->ident_merged_upto;
  if (first < h->next_step - DG_MAX_INFLIGHT) first = h->next_step - DG_MAX_INFLIGHT;
  bool merged = false;
  for (long long step = first; step < h->next_step; step++) {
    const int slot = (int)(step % 3);
    int F = 0, K = 0;
    dg_seg_dims(h->seg, h->slot_S[slot], &F, &K);
    int32_t* maps = h->slot_map[slot].as<int32_t>();
    const int n = h->slot_B[slot] * K;
    if (!merged) {
      if ((rc = dg_cluster_merge(h->clu, records_dev, world, rank, maps, n, h->s_clu))) return rc;
      merged = true;
    } else if ((rc = launch_relabel_maps(maps, n, h->clu->relabel.as<int32_t>(), h->s_clu))) {
      return rc;
    }
    DG_CUDA(cudaEventRecord(h->e_slot_done[slot], h->s_clu));      // collect must see the relabelled maps
  }
  if (!merged && (rc = dg_cluster_merge(h->clu, records_dev, world, rank, nullptr, 0, h->s_clu))) return rc;
  h->ident_merged_upto = h->next_step;
  return DG_OK;
}

// pipelined step whose batch is the next B windows of a device-side stream (no window upload at all; the sinc layer takes
// its stream form without the overlap check: the windows overlap by construction)
extern "C" int dg_pipeline_submit_stream(dg_pipeline* h, dg_stream* s, int B) {
  if (!h || !s || B < 1) {
    set_error("dg_pipeline_submit_stream: bad arguments");
    return DG_EINVAL;
  }
  if (h->outstanding >= DG_MAX_INFLIGHT) {
    set_error("dg_pipeline_submit_stream: three steps are already outstanding; collect one first");
    return DG_EINVAL;
  }
  if (s->device != h->seg->device) {
    set_error("dg_pipeline_submit_stream: stream and pipeline live on different devices");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  const int S = s->S;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  DG_CUDA(cudaSetDevice(h->seg->device));
  const int slot = (int)(h->next_step % 3);
  if ((rc = pipeline_slot_prepare(h, slot, B, S, F, K, true))) return rc;
  DG_CUDA(cudaStreamWaitEvent(h->s_h2d, h->e_slot_done[slot], 0));
  if ((rc = stream_expand(s, B, h->slot_wav[slot].as<float>(), h->s_h2d))) return rc;
  DG_CUDA(cudaEventRecord(h->e_h2d[slot], h->s_h2d));
  return pipeline_submit_common(h, h->slot_wav[slot].as<float>(), B, S, F, K, slot, h->e_h2d[slot], s->hop);
}

// SpeakerDiarization.__call__ for the next B windows of a device-side stream: fused step + post-path, synchronous
extern "C" int dg_pipeline_call_stream(dg_pipeline* h, dg_post* post, dg_stream* s, int B, const int32_t* plan_host,
                                       int32_t* header_host, uint32_t* turns_host, int turn_cap_host, int* n_turns,
                                       float* seg_host, int32_t* map_host) {
  if (!h || !post || !s || !plan_host || !header_host || !turns_host || B < 1) {
    set_error("dg_pipeline_call_stream: bad arguments");
    return DG_EINVAL;
  }
  if (h->outstanding) {
    set_error("dg_pipeline_call_stream: submitted steps are outstanding; collect them first");
    return DG_EINVAL;
  }
  int rc, F = 0, K = 0;
  const int S = s->S;
  if ((rc = dg_seg_dims(h->seg, S, &F, &K))) return rc;
  if (F != post->F || K != post->K || h->clu->p.M != post->M || post->device != h->seg->device || s->device != h->seg->device) {
    set_error("dg_pipeline_call_stream: handles were created for other dimensions / devices");
    return DG_EINVAL;
  }
  const int D = h->emb->D;
  DG_CUDA(cudaSetDevice(h->seg->device));
  if (h->wav.ensure((size_t)B * S * 4) || h->segd.ensure((size_t)B * F * K * 4) || h->embd.ensure((size_t)B * K * D * 4) ||
      h->mapd.ensure((size_t)B * K * 4))
    return DG_ECUDA;
  if ((rc = stream_expand(s, B, h->wav.as<float>(), h->st))) return rc;
  const int hop_saved = h->hop;
  h->hop = s->hop;
  h->overlap_known = true;
  rc = dg_pipeline_step(h, h->wav.as<float>(), B, S, h->segd.as<float>(), h->embd.as<float>(), h->mapd.as<int32_t>(), nullptr, h->st);
  h->overlap_known = false;
  h->hop = hop_saved;
  if (rc) return rc;
  if ((rc = post_enqueue(post, h->segd.as<float>(), h->mapd.as<int32_t>(), B, plan_host, h->st))) return rc;
  if (seg_host) DG_CUDA(cudaMemcpyAsync(seg_host, h->segd.p, (size_t)B * F * K * 4, cudaMemcpyDeviceToHost, h->st));
  if (map_host) DG_CUDA(cudaMemcpyAsync(map_host, h->mapd.p, (size_t)B * K * 4, cudaMemcpyDeviceToHost, h->st));
  DG_CUDA(cudaStreamSynchronize(h->st));
  return post_finish(post, B, header_host, turns_host, turn_cap_host, n_turns, h->st);
}

extern "C" int dg_pipeline_destroy(dg_pipeline* h) {
  if (h) {
    if (h->s_seg) cudaStreamDestroy(h->s_seg);
    if (h->s_emb) cudaStreamDestroy(h->s_emb);
    if (h->s_clu) cudaStreamDestroy(h->s_clu);
    if (h->s_seg2) cudaStreamDestroy(h->s_seg2);
    if (h->e_osp2) cudaEventDestroy(h->e_osp2);
    if (h->e_prep[0]) cudaEventDestroy(h->e_prep[0]);
    if (h->e_prep[1]) cudaEventDestroy(h->e_prep[1]);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    for (cudaEvent_t e : {h->e_start, h->e_osp, h->e_emb, h->e_done, h->e_h2d[0], h->e_h2d[1], h->e_h2d[2],
                          h->e_slot_done[0], h->e_slot_done[1], h->e_slot_done[2], h->e_lane_done[0], h->e_lane_done[1],
                          h->e_ident, h->e_ident_in})
      if (e) cudaEventDestroy(e);
  }
  if (h && h->st) cudaStreamDestroy(h->st);
  if (h && h->pin_wav) cudaFreeHost(h->pin_wav);
  delete h;
  return DG_OK;
}
