// Shared declarations for the diart_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <vector>

namespace dg {

// ------------------------------------------------------------------ error plumbing
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;
// persistent kernels size their grid to the SM count; while two streams overlap (fused pipeline) the
// kernels of the lower-priority stream are capped so that they never wait for SMs held by the other one
extern thread_local int g_sm_limit;
inline int usable_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (g_sm_limit > 0 && g_sm_limit < sms) ? g_sm_limit : sms;
}

#define DG_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      dg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

// cudaFuncSetAttribute is per DEVICE: `flags` is a function-local static bool[64]; returns true the first time it is asked
// about the current device (the caller then sets the attribute)
inline bool first_use_on_device(bool* flags) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// every kernel launch goes through this so bench.py can report `gpu_launches`
#define DG_LAUNCHED()                                                                   \
  do {                                                                                  \
    dg::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      dg::set_error(std::string("kernel launch failed at ") + __FILE__ + ":" +          \
                    std::to_string(__LINE__) + ": " + cudaGetErrorString(_e));          \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

// per-kernel CUDA-event timing (bench.py's roofline leg): when enabled through dg_profile_enable(),
// every launcher brackets its launch with two events on the launching stream.
struct ProfScope {
  bool on;
  cudaStream_t st;
  cudaEvent_t a, b;
  const char* name;
  long long first = 0;
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
};

// ------------------------------------------------------------------ geometry of the path
// 80000 samples -> sinc(k251,s10) 7975 -> pool3 2658 -> k5 2654 -> pool3 884 -> k5 880 -> pool3 293.
// Activations are stored time-major / channels-last, [item][row][channel], with a fixed row
// stride per item that is divisible by 9 so that the two pool-by-3 stages keep rows aligned:
// a conv layer is then a pure shifted-window GEMM over the flattened [B*stride, C] matrix and
// rows past an item's valid length are harmless finite garbage that no consumer reads.
struct Geom {
  int S;        // samples per chunk
  int T0c;      // sinc conv outputs              (7975)
  int T0;       // after pool                      (2658)
  int T1;       // conv1+pool valid                (884)
  int T2;       // conv2+pool valid = frames       (293)
  int S0, S1, S2;  // row strides per item         (2664, 888, 296)
};
inline Geom make_geom(int S) {
  Geom g;
  g.S = S;
  g.T0c = (S - 251) / 10 + 1;
  g.T0 = g.T0c / 3;
  g.T1 = (g.T0 - 4) / 3;
  g.T2 = (g.T1 - 4) / 3;
  g.S2 = g.T2 + 3;                      // >= T2, room for pool alignment
  while ((g.S2 * 9) < g.T0 + 0 || (g.S2 * 3) < g.T1) g.S2++;
  g.S1 = g.S2 * 3;
  g.S0 = g.S1 * 3;
  return g;
}

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

// ------------------------------------------------------------------ launchers (defined per .cu)
// sincnet.cu
int launch_wave_stats(const float* wav, int B, int S, float* mean, float* rstd, cudaStream_t st, const int* skip_flag = nullptr);
bool stream_stats_ok(int S, int hop);
size_t stream_stats_doubles(int B, int S, int hop);
int launch_stream_stats(const float* wav, int B, int S, int hop, double* part, float* mean, float* rstd, const int* flag,
                        cudaStream_t st);
int launch_sinc0(const float* wav, const float* mean, const float* rstd, float wn_gamma, float wn_beta,
                 const float* filt /*[251][80]*/, int B, const Geom& g, float* p0 /*[B,S0,80]*/, cudaStream_t st);
int launch_instnorm_stats(const float* x, int B, int stride_rows, int T, int C, int ldc, const float* gamma,
                          const float* beta, float* sc, float* sh, cudaStream_t st, int pool = 0, const int* skip_flag = nullptr);
int launch_instnorm_finalize(const float* part, int B, int item_rows, int tile_rows, int T, int C, int N, const float* bias,
                             const float* gamma, const float* beta, float* sc, float* sh, int ld, cudaStream_t st);
// gemm.cu
enum Epi { EPI_BIAS = 0, EPI_BIAS_LEAKY = 1, EPI_BIAS_LEAKY_BN = 2, EPI_BIAS_POOL3 = 3 };
struct GemmArgs {
  const float* A;      // [Mrows_in, lda]
  int lda;             // channel stride of A (>= Cin)
  int Cin;             // channels consumed per tap (multiple of 4)
  int KW, dil;         // taps, dilation (rows)
  long long Mtot;      // rows of A that exist (reads past it return 0)
  long long M;         // output rows to produce (before pooling)
  const float* W;      // [KW*Cin, ldw]
  int ldw;             // column stride of W (>= N, multiple of 4)
  int N;               // valid output channels
  const float* bias;   // [N] or null
  const float* bn_scale;  // [N] (EPI_BIAS_LEAKY_BN)
  const float* bn_shift;
  const float* in_sc;  // per-(item, channel) instance-norm scale/shift applied (+leaky) on load, or null
  const float* in_sh;
  int item_rows;       // rows per item (for in_sc indexing)
  float* C;            // [M or M/3, ldc]
  int ldc;
  int epi;
  const char* tag;   // kernel label for profiling (layer name)
};
int launch_gemm(const GemmArgs& a, cudaStream_t st);
// gemm_tc.cu -- tcgen05 / TMEM / TMA path (hi/lo split precision, three products)
struct TcGemm {
  const void* A_hi;    // bf16 [Mtot, lda]
  const void* A_lo;
  int lda, Cin, KW, dil;
  long long Mtot, M;
  const void* W_hi;    // bf16 [Npad, KW*Cin]  (n-major: row n holds its K weights, tap-major)
  const void* W_lo;
  float w_scale;       // power-of-two factor the W planes were multiplied by (0 = 1): undone on the accumulator
  int Npad, N;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  float* out_f32;      // [M, ldc] (float32 epilogues)
  void* out_hi;        // bf16 [M, ldc] (split epilogue)
  void* out_lo;
  int ldc;
  int epi;             // 0 bias -> f32, 1 bias+leaky+bn -> hi/lo planes, 2 bias+leaky+bn -> f32, 3 conv2d (see below)
  const char* tag;
  const int* tap_off;  // host array [KW]: row offset of every tap; null = j * dil (Conv1d)
  int sm_limit;        // > 0: cap of the persistent grid (a lower-priority stream leaves SMs to the other one)
  // epi 3 (Conv2d on zero-padded channels-last maps, rows = (item, w, h) of a [Wp][Hp] map): BatchNorm2d affine (bn_scale,
  // bn_shift) -> + residual planes -> ReLU -> hi/lo planes (and / or float32) at the same position of the [Wop][Hop] map;
  // stride2: the convolution is evaluated at every centre and only the odd (w, h) are kept
  int Wp, Hp, Wop, Hop, stride2, relu;
  const void* res_hi;
  const void* res_lo;
  // epi 4 (TDNN5 + weighted statistics pooling, heads.cu: pool_finalize): bias -> LeakyReLU -> BatchNorm, then per 128-row tile
  // and item the sums  S1 = sum_t w_k[t] d,  S2 = sum_t w_k[t] d^2  of d = x - bn_shift  for the K local speakers
  const float* pool_w;   // [Mtot][4]
  float* pool_part;      // [m_tiles][2][4][2][N]
  int pool_item_rows, pool_K;
  // epi 5 (SincNet Conv1d + MaxPool1d(3) + InstanceNorm statistics): out_f32 = bias + max over row triplets ([M / 3, ldc]),
  // pool_part = per-tile partial sums [M / pool3_tile_rows][2][2][N], reduced by launch_instnorm_finalize;
  // pool_item_rows = un-pooled rows per item (multiple of 3), pool3_T = valid pooled frames per item
  int pool3_T, pool3_tile_rows;
};
// rows an m-tile of the pooling epilogue advances by: the largest multiple of 3 up to 126 that divides the item's rows
// (tiles never straddle items: an item's statistics are grouped identically wherever it sits in the batch); 0 if none >= 96
inline int gemm_tc_pool3_tile_rows(int item_rows) {
  for (int t = 126; t >= 96; t -= 3)
    if (item_rows % t == 0) return t;
  return 0;
}
int launch_gemm_tc(const TcGemm& g, cudaStream_t st);
int launch_split(const float* x, long long rows, int C, int item_rows, const float* sc, const float* sh, void* hi,
                 void* lo, cudaStream_t st);
int launch_split_ex(const float* x, long long rows_out, int C, int ld_in, int ld_out, int pool, int item_rows,
                    const float* sc, const float* sh, void* hi, void* lo, cudaStream_t st, const int* skip_flag = nullptr);
// element type of the 16-bit operand planes: 1 = fp16 (default), 0 = bf16 (DG_SPLIT_BF16=1); fixed at first use
int split_f16();
void split_weights_host(const float* w, int N, int Npad, int K, uint16_t* hi, uint16_t* lo, int f16, float scale = 1.f);
float weight_plane_scale(const float* w, size_t n, int f16);   // power of two that keeps the lo plane of small weights normal
uint16_t host_f32_to_h16(float f, int f16);
float host_h16_to_f32(uint16_t h, int f16);
// sinc_tc.cu -- SincNet stage 0 on tcgen05 (overlapping-row TMA view of the waveform)
int sinc_tc_rows_per_item(const Geom& g);
size_t sinc_tc_plane_elems(int B, const Geom& g);
void sinc_tc_pack_filters(const float* filt, uint16_t* planes /*[3][80][256]*/, int f16);
void sinc_tc_affine_consts(const float* filt, float beta, float* cf);
int launch_sinc_prep(const float* wav, const float* mean, const float* rstd, int B, const Geom& g, void* planes_hi,
                     void* planes_lo, cudaStream_t st, const int* skip_flag = nullptr);
int launch_sinc0_tc(float gamma, const float* cf_dev, const void* w_planes, int B, const Geom& g, const void* planes_hi,
                    const void* planes_lo, float* p0, cudaStream_t st, const int* skip_flag = nullptr);
// stream form of the sinc layer (sinc_tc.cu): the batch is B windows of one stream, `hop` samples apart
struct SincStreamGeom {
  int Ls;        // unique samples (B-1)*hop + S
  int P;         // conv positions of the stream
  int rows;      // rows of the overlapping-row view (one "item")
  size_t plane;  // elements per shifted copy
};
SincStreamGeom sinc_stream_geom(int B, const Geom& g, int hop);
int launch_overlap_check(const float* wav, int B, int S, int hop, int* flag, cudaStream_t st);
int launch_stream_prep(const float* wav, int B, const Geom& g, int hop, void* planes_hi, void* planes_lo, const int* flag,
                       cudaStream_t st);
int launch_sinc0_tc_stream(const void* w_planes, int B, const Geom& g, int hop, const void* planes_hi, const void* planes_lo,
                           float* craw, const int* flag, cudaStream_t st);
int launch_sinc_pool(const float* craw, const float* mean, const float* rstd, const float* cf, const float* hsum, float gamma,
                     int B, const Geom& g, int hop, float* p0, const int* flag, cudaStream_t st);
size_t sinc_pool_part_floats(int B, const Geom& g, int hop);
int launch_sinc_pool_fused(const float* craw, const float* mean, const float* rstd, const float* cf, const float* hsum, float gamma,
                           int B, const Geom& g, int hop, const float* g0, const float* b0, float* part, float* sc, float* sh,
                           void* planes_hi, void* planes_lo, const int* flag, cudaStream_t st);
// lstm.cu
int launch_lstm_layer(const float* gx /*[B*stride,1024]*/, const float* whh_packed, int B, int T, int stride,
                      float* hout /*[B*stride,256]*/, cudaStream_t st);
size_t lstm_whh_packed_floats();
void lstm_pack_whh(const float* whh_fwd /*[512][128]*/, const float* whh_bwd, float* packed);
// lstm_tc.cu -- recurrence on tcgen05 (W_hh hi plane in shared memory, lo plane in tensor memory)
size_t lstm_tc_plane_elems();
int lstm_tc_ctas(int B);   // CTAs (= SMs) one recurrence launch occupies at batch B
float lstm_tc_pack_whh(const float* whh_fwd, const float* whh_bwd, uint16_t* hi, uint16_t* lo, int f16);   // -> plane scale
// hout (float32) and / or out_hi, out_lo (16-bit planes of the next GEMM's operand) receive h_t
int launch_lstm_layer_tc(const float* gx, const void* whh_hi, const void* whh_lo, float w_scale, int B, int T, int stride,
                         float* hout, void* out_hi, void* out_lo, cudaStream_t st);
// heads.cu
int launch_seg_final(const float* y /*[B*stride,128]*/, const float* wc /*[K][128]*/, const float* bc, int B, int T,
                     int stride, int K, float* seg /*[B,T,K]*/, cudaStream_t st);
int launch_seg_powerset(const float* y, const float* wc, const float* bc, int B, int T, int stride, int C, int num_speakers,
                        const unsigned* masks_dev, float* seg /*[B,T,num_speakers]*/, cudaStream_t st);
int launch_osp(const float* seg, int B, int F, int K, float gamma, float beta, int normalize, float* out,
               cudaStream_t st);
int launch_stats_pool(const float* x /*[B*stride,C]*/, int B, int stride, int T, int C, const float* w /*[B,F,K]*/,
                      int F, int K, const int* idx0, const int* idx1, const float* lam1, float eps,
                      float* pooled /*[B*K, 2C]*/, cudaStream_t st, long long item_pitch = 0, int row_pitch = 0);
// fused pooling (epi 4 of gemm_tc): row weights + their sums, and the final mean / std from the per-tile partial sums
int launch_pool_weights(const float* w /*[B,F,K]*/, int B, int F, int K, int item_rows, int T, const int* idx0, const int* idx1,
                        const float* lam1, float eps, float* row_w /*[B*item_rows][4]*/, float* vsum /*[B*K][2]*/, cudaStream_t st);
int launch_pool_finalize(const float* part, const float* vsum, const float* pivot, int B, int K, int C, int item_rows, int T,
                         float eps, float* pooled /*[B*K][2C]*/, cudaStream_t st);
int launch_l2norm(const float* in, int rows, int D, float norm, float* out, cudaStream_t st);
int launch_row_equal_flags(const float* wav, int N, int S, int* flags, cudaStream_t st);
int launch_gather_rows(const float* src, const int* index, int rows, int cols, float* dst, cudaStream_t st);
// resnet.cu -- variant B of the embedding (WeSpeaker ResNet34): fbank front end and stem around the Conv2d GEMMs
int launch_fb_planes(const float* wav, long long n, void* hi, void* lo, cudaStream_t st);
int launch_fb_mel(const float* spec, int ld, int rows_per_item, int T, int B, const float* banks, const int* k_lo, const int* k_hi,
                  float* logmel, cudaStream_t st);
int launch_fb_mean(const float* logmel, int B, int T, float* mean, cudaStream_t st);
int launch_rn_stem(const float* logmel, const float* mean, int B, int T, const float* w, const float* sc, const float* sh,
                   void* hi, void* lo, cudaStream_t st);
// cluster.cu
struct ClusterParams {
  int M, D;
  float tau_f, rho_f;   // thresholds as numpy compares them (float32, see cluster.cu)
  double delta;
  int metric;           // 0 cosine (default), 1 euclidean, 2 sqeuclidean, 3 cityblock, 4 chebyshev (scipy cdist names)
};
int launch_cluster_step(const ClusterParams& p, const float* seg, const float* emb, int B, int F, int K,
                        double* centers, int* active, int* initialized, float* prep /*scratch*/,
                        double* prep_d /*scratch*/, int32_t* map, float* permuted, cudaStream_t st);
int launch_cluster_export(const double* centers, const int* active, const double* base, const int* base_active, int M,
                          int D, double* record, cudaStream_t st);
int launch_cluster_merge(const double* records, int world, int rank, const ClusterParams& p, int rec_len, double* centers,
                         int* active, double* base, int* base_active, int* initialized, int32_t* relabel,
                         cudaStream_t st);
int launch_relabel_maps(int32_t* maps, int n, const int32_t* relabel, cudaStream_t st);
size_t cluster_prep_floats(int B, int K);
// post.cu -- aggregation + binarisation + run-length turns (reference diarization.py:205-232)
int launch_post(const float* seg, const int32_t* map, const float* hist_seg, const int32_t* hist_map, int n_hist, int B,
                int F, int K, int M, int nw, const int32_t* plan, int plan_stride, const double* hamming, double tau,
                int32_t* header, uint32_t* turns, int turn_cap, unsigned int* total, cudaStream_t st);
int launch_expand_windows(const float* ring, long long r0, int C, int hop, int S, int B, float* wav, cudaStream_t st);
int launch_post_history(const float* seg, const int32_t* map, const float* hist_seg, const int32_t* hist_map, int n_hist,
                        int B, int F, int K, int keep, float* new_seg, int32_t* new_map, cudaStream_t st);
size_t cluster_prep_doubles(int B, int K);

void fbank_frame_operator(std::vector<float>& op /*[514][400]*/);
void fbank_mel_banks(std::vector<float>& banks /*[80][257]*/, std::vector<int>& k_lo, std::vector<int>& k_hi);

}  // namespace dg
