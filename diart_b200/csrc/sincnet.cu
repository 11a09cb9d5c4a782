// SincNet front end, stage 0 (shared by the segmentation and the embedding net, separate weights):
//   InstanceNorm1d(1, affine) on the waveform -> ParamSincFB 80 x k251 stride 10 -> |.| -> MaxPool1d(3)
// plus the per-(item, channel) InstanceNorm statistics that the next layer applies on load.
// Restates pyannote.audio's SincNet.forward; reached from the reference
// through src/diart/models.py:131-133.
#include "dg_common.cuh"

namespace dg {

// ---------------------------------------------------------------------------------------------
// waveform statistics: mean and 1/sqrt(biased var + 1e-5) per item.  One CTA per item, two passes
// (the second one hits L1/L2), per-thread float partials combined in double.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < nw; i++) t += sm[i];
  return t;
}

__global__ void __launch_bounds__(512) wave_stats_kernel(const float* __restrict__ wav, int S, float* __restrict__ mean,
                                                         float* __restrict__ rstd, const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;     // the stream form computes the statistics from per-hop partial sums
  __shared__ double sm[32];
  const float* x = wav + (size_t)blockIdx.x * S;
  float s = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += x[i];
  double m = block_sum_d((double)s, sm) / S;
  float mf = (float)m, q = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float d = x[i] - mf;
    q += d * d;
  }
  double var = block_sum_d((double)q, sm) / S;
  if (threadIdx.x == 0) {
    mean[blockIdx.x] = mf;
    rstd[blockIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
}

int launch_wave_stats(const float* wav, int B, int S, float* mean, float* rstd, cudaStream_t st, const int* skip_flag) {
  ProfScope _ps("wave_stats", st);
  wave_stats_kernel<<<B, 512, 0, st>>>(wav, S, mean, rstd, skip_flag);
  DG_LAUNCHED();
  return 0;
}

// ---- stream form: the B windows are a run of one stream (window b = samples [b*hop, b*hop + S)), so every sample is summed
// ONCE: partial (sum x, sum x^2) per quarter hop in double, then each window adds its 4 S / hop partials.
// (B x S = 82 MB read by 256 CTAs becomes 8.5 MB read by ~1000.)
__global__ void __launch_bounds__(256) stream_sums_kernel(const float* __restrict__ wav, int B, int S, int hop, int sub,
                                                          double* __restrict__ part, const int* __restrict__ flag) {
  if (*flag == 0) return;
  __shared__ double sm[32];
  const long long first = (long long)blockIdx.x * sub;           // first stream sample of this block
  int b = (int)(first / hop);
  if (b > B - 1) b = B - 1;
  const float4* x = reinterpret_cast<const float4*>(wav + (size_t)b * S + (first - (long long)b * hop));
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < (sub >> 2); i += blockDim.x) {
    const float4 v = x[i];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
  }
  const double t1 = block_sum_d((double)s1, sm);
  const double t2 = block_sum_d((double)s2, sm);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = t1;
    part[2 * blockIdx.x + 1] = t2;
  }
}

__global__ void stream_stats_kernel(const double* __restrict__ part, int B, int S, int hop, int sub, float* __restrict__ mean,
                                    float* __restrict__ rstd, const int* __restrict__ flag) {
  if (*flag == 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int per_hop = hop / sub, n = S / sub;
  double t1 = 0, t2 = 0;
  for (int i = 0; i < n; i++) {
    t1 += part[2 * (b * per_hop + i)];
    t2 += part[2 * (b * per_hop + i) + 1];
  }
  const double m = t1 / S;
  double var = t2 / S - m * m;
  if (var < 0) var = 0;
  mean[b] = (float)m;
  rstd[b] = (float)(1.0 / sqrt(var + 1e-5));
}

bool stream_stats_ok(int S, int hop) { return hop % 16 == 0 && S % (hop / 4) == 0; }
size_t stream_stats_doubles(int B, int S, int hop) { return 2 * ((size_t)(B - 1) * 4 + (size_t)S / (hop / 4)) + 8; }

int launch_stream_stats(const float* wav, int B, int S, int hop, double* part, float* mean, float* rstd, const int* flag,
                        cudaStream_t st) {
  ProfScope _ps("wave_stats", st);
  const int sub = hop / 4;
  const int blocks = (B - 1) * 4 + S / sub;
  stream_sums_kernel<<<blocks, 256, 0, st>>>(wav, B, S, hop, sub, part, flag);
  DG_LAUNCHED();
  stream_stats_kernel<<<(B + 127) / 128, 128, 0, st>>>(part, B, S, hop, sub, mean, rstd, flag);
  DG_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// sinc0: one CTA tile = 64 pooled outputs (192 conv positions) x 80 filters of one item.
// Persistent CTAs keep the 80 KB filter bank in shared memory and walk the (item, tile) list.
// Thread tile: 1 pooled output (3 conv positions) x 8 filters; lanes run over pooled positions,
// warps over (position half, filter group).  fp32 FMA throughout.
// ---------------------------------------------------------------------------------------------
constexpr int SINC_K = 251, SINC_F = 80, SINC_TP = 64, SINC_THREADS = 640;
constexpr int SINC_XSEG = 30 * (SINC_TP - 1) + 20 + SINC_K;  // 2161 samples feed one tile

__global__ void __launch_bounds__(SINC_THREADS, 1)
sinc0_kernel(const float* __restrict__ wav, const float* __restrict__ mean, const float* __restrict__ rstd,
             float wn_gamma, float wn_beta, const float* __restrict__ filt, int B, int S, int T0, int S0,
             int tiles_per_item, float* __restrict__ p0) {
  extern __shared__ float smem[];
  float* hs = smem;                       // [251][80]
  float* xs = smem + SINC_K * SINC_F;     // [SINC_XSEG padded]
  for (int i = threadIdx.x; i < SINC_K * SINC_F / 4; i += blockDim.x)
    reinterpret_cast<float4*>(hs)[i] = reinterpret_cast<const float4*>(filt)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int phalf = warp & 1, fg = warp >> 1;          // 2 position halves x 10 filter groups
  const int pl = phalf * 32 + lane;                    // pooled position within the tile
  const int total = B * tiles_per_item;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int b = tile / tiles_per_item, p_base = (tile - b * tiles_per_item) * SINC_TP;
    const float mu = mean[b], sc = rstd[b] * wn_gamma;
    const float* x = wav + (size_t)b * S;
    const int x0 = 30 * p_base;
    __syncthreads();
    for (int i = threadIdx.x; i < SINC_XSEG; i += blockDim.x) {
      int gi = x0 + i;
      // InstanceNorm1d(1, affine): (x - mean) * rstd * gamma + beta
      xs[i] = gi < S ? (x[gi] - mu) * sc + wn_beta : 0.f;
    }
    __syncthreads();
    const int p = p_base + pl;
    float acc[3][8];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int f = 0; f < 8; f++) acc[j][f] = 0.f;
    const float* xp = xs + 30 * pl;
    const float* hp = hs + fg * 8;
#pragma unroll 2
    for (int k = 0; k < SINC_K; k++) {
      const float4 h0 = *reinterpret_cast<const float4*>(hp + k * SINC_F);
      const float4 h1 = *reinterpret_cast<const float4*>(hp + k * SINC_F + 4);
      const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float xv = xp[10 * j + k];
#pragma unroll
        for (int f = 0; f < 8; f++) acc[j][f] = fmaf(xv, hv[f], acc[j][f]);
      }
    }
    if (p < T0) {
      float v[8];
#pragma unroll
      for (int f = 0; f < 8; f++) v[f] = fmaxf(fmaxf(fabsf(acc[0][f]), fabsf(acc[1][f])), fabsf(acc[2][f]));
      float* o = p0 + ((size_t)b * S0 + p) * SINC_F + fg * 8;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

int launch_sinc0(const float* wav, const float* mean, const float* rstd, float wn_gamma, float wn_beta,
                 const float* filt, int B, const Geom& g, float* p0, cudaStream_t st) {
  ProfScope _ps("sinc0", st);
  static bool attr_done[64] = {};
  const size_t smem = (size_t)(SINC_K * SINC_F + ((SINC_XSEG + 3) & ~3)) * sizeof(float);
  if (first_use_on_device(attr_done))
    DG_CUDA(cudaFuncSetAttribute(sinc0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_per_item = (g.T0 + SINC_TP - 1) / SINC_TP;
  const int total = B * tiles_per_item;
  const int grid = total < 2 * sms ? total : 2 * sms;
  sinc0_kernel<<<grid, SINC_THREADS, smem, st>>>(wav, mean, rstd, wn_gamma, wn_beta, filt, B, g.S, g.T0, g.S0,
                                                  tiles_per_item, p0);
  DG_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm1d(C, affine) statistics over the T valid rows of each item of x[B, stride, ldc]:
// emits sc = gamma * rstd, sh = beta - mean * gamma * rstd so the consumer computes
// leaky(x * sc + sh) on load.  CTA = (item, 32-channel group); 8 warps stride over rows; sums are
// taken around the first row's value (pivot) to avoid E[x^2]-E[x]^2 cancellation, combined in double.
// ---------------------------------------------------------------------------------------------
// pool != 0: x holds the un-pooled conv output (stride_rows rows per item) and the statistics are taken over
// MaxPool1d(3) of it, i.e. value(t) = max(x[3t], x[3t+1], x[3t+2]) for t < T.
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const float* __restrict__ x, int stride_rows, int T, int C,
                                                             int ldc, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ sc,
                                                             float* __restrict__ sh, int pool,
                                                             const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;       // the fused stream-form tail produced these statistics
  __shared__ double s1[8][32], s2[8][32];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  const bool ok = c < C;
  const float* xb = x + (size_t)b * stride_rows * ldc;
  auto val = [&](int t) -> float {
    if (!pool) return xb[(size_t)t * ldc + c];
    const float* p = xb + (size_t)(3 * t) * ldc + c;
    return fmaxf(fmaxf(p[0], p[ldc]), p[2 * ldc]);
  };
  const float pivot = ok ? val(0) : 0.f;
  float a1 = 0.f, a2 = 0.f;
  if (ok)
    for (int t = w; t < T; t += 8) {
      float d = val(t) - pivot;
      a1 += d;
      a2 = fmaf(d, d, a2);
    }
  s1[w][threadIdx.x & 31] = a1;
  s2[w][threadIdx.x & 31] = a2;
  __syncthreads();
  if (w == 0 && ok) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < 8; i++) {
      t1 += s1[i][threadIdx.x];
      t2 += s2[i][threadIdx.x];
    }
    double m = t1 / T, var = t2 / T - m * m;
    if (var < 0) var = 0;
    double mean = (double)pivot + m;
    float r = (float)(1.0 / sqrt(var + 1e-5));
    float gsc = gamma[c] * r;
    sc[(size_t)b * C + c] = gsc;
    sh[(size_t)b * C + c] = beta[c] - (float)mean * gsc;
  }
}

int launch_instnorm_stats(const float* x, int B, int stride_rows, int T, int C, int ldc, const float* gamma,
                          const float* beta, float* sc, float* sh, cudaStream_t st, int pool, const int* skip_flag) {
  ProfScope _ps("instnorm_stats", st);
  dim3 grid((C + 31) / 32, B);
  instnorm_stats_kernel<<<grid, 256, 0, st>>>(x, stride_rows, T, C, ldc, gamma, beta, sc, sh, pool, skip_flag);
  DG_LAUNCHED();
  return 0;
}

// InstanceNorm1d(affine) scale / shift from the per-tile partial sums of gemm_tc's TC_MAXPOOL3 epilogue (tiles of `tile_rows`
// un-pooled rows, a divisor of the item's rows: slot 0 of the item's own tiles).  The sums are of the pooled values BEFORE the
// bias, which serves as the pivot; the tiles of an item are added in tile order, in double.
__global__ void __launch_bounds__(256) instnorm_finalize_kernel(const float* __restrict__ part, int tiles_per_item, int T, int C, int N,
                                                                const float* __restrict__ bias, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ sc,
                                                                float* __restrict__ sh, int ld) {
  // 4 tile groups x 64 channels: the loads of a group are independent, the groups are added in a fixed order
  __shared__ double s1[4][64], s2[4][64];
  const int b = blockIdx.x, c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const long long mt0 = (long long)b * tiles_per_item;
  const int per = (tiles_per_item + 3) / 4, lo = grp * per, hi = min(tiles_per_item, lo + per);
  double t1 = 0, t2 = 0;
  if (c < C)
    for (int i = lo; i < hi; i++) {
      t1 += part[(((mt0 + i) * 2 + 0) * 2 + 0) * N + c];
      t2 += part[(((mt0 + i) * 2 + 0) * 2 + 1) * N + c];
    }
  s1[grp][c] = t1;
  s2[grp][c] = t2;
  __syncthreads();
  if (grp != 0 || c >= C) return;
  t1 = ((s1[0][c] + s1[1][c]) + s1[2][c]) + s1[3][c];
  t2 = ((s2[0][c] + s2[1][c]) + s2[2][c]) + s2[3][c];
  const double m = t1 / T;
  double var = t2 / T - m * m;
  if (var < 0) var = 0;
  const double mean = m + (double)bias[c];
  const float r = (float)(1.0 / sqrt(var + 1e-5));
  const float gsc = gamma[c] * r;
  sc[(size_t)b * ld + c] = gsc;
  sh[(size_t)b * ld + c] = beta[c] - (float)mean * gsc;
}

int launch_instnorm_finalize(const float* part, int B, int item_rows, int tile_rows, int T, int C, int N, const float* bias,
                             const float* gamma, const float* beta, float* sc, float* sh, int ld, cudaStream_t st) {
  ProfScope _ps("instnorm_finalize", st);
  if (C > 64 || tile_rows < 1 || item_rows % tile_rows) {
    set_error("instnorm_finalize: at most 64 channels, tiles must divide the item");
    return -1;
  }
  instnorm_finalize_kernel<<<B, 256, 0, st>>>(part, item_rows / tile_rows, T, C, N, bias, gamma, beta, sc, sh, ld);
  DG_LAUNCHED();
  return 0;
}

}  // namespace dg
