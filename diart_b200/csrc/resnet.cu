// Variant B of the embedding row (SURVEY.md 8(a) A8'): pyannote/wespeaker-voxceleb-resnet34-LM -- the small kernels around
// the tensor-core convolutions.  The reference reaches this network through src/diart/models.py:50,59 (README.md:172-173);
// its arithmetic lives in pyannote.audio's WeSpeakerResNet34 + torchaudio.compliance.kaldi.fbank (restated in oracle/nets.py).
//
//   waveform * 2^15 -> kaldi fbank (25 ms / 10 ms frames, DC removal, pre-emphasis 0.97, Hamming, 512-point power spectrum,
//   80 mel bins, log) -> mean normalisation over time -> Conv2d(1, 32, 3) + BN + ReLU -> 16 BasicBlocks -> TSTP -> Linear
//
// Mapping (prototyped against the oracle on the CPU in oracle/fbank_linear.py and oracle/resnet_gemm_form.py):
//   * every linear step of the fbank before the power spectrum collapses into ONE [514, 400] operator applied to an
//     OVERLAPPING-ROW view of the waveform (row pitch 160 samples): a plain launch of the split-precision tcgen05 GEMM of
//     gemm_tc.cu (K = 400 padded to 448, N = 514 padded to 640), no frame matrix is ever materialised;
//   * maps are channels-last on a zero-padded grid, [item][w = time + 1][h = mel + 1][C], so that a 3x3 convolution is a
//     shifted-window GEMM whose nine taps are row offsets (dw-1) * Hp + (dh-1) and padding is simply the zeros already in
//     memory (gemm_tc.cu, TC_CONV2D epilogue: BatchNorm affine, residual, ReLU, hi/lo planes of the next layer);
//   * this file: waveform -> 16-bit planes, power spectrum -> mel -> log, the time mean, the one-channel stem convolution.
#include <math.h>

#include <vector>

#include "dg_common.cuh"
#include "tc_ptx.cuh"

namespace dg {

// waveform [B, S] float32 -> hi / lo 16-bit planes of x * 2^15 (what pyannote feeds kaldi.fbank), [B * S (+ tail zeros)]
__global__ void __launch_bounds__(256) fb_planes_kernel(const float* __restrict__ wav, long long n, uint16_t* __restrict__ hi,
                                                        uint16_t* __restrict__ lo, int f16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(wav)[i];
    uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_h16(v.x * 32768.f, f16, h0, l0);
    split_h16(v.y * 32768.f, f16, h1, l1);
    split_h16(v.z * 32768.f, f16, h2, l2);
    split_h16(v.w * 32768.f, f16, h3, l3);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_u16x2(h0, h1), pack_u16x2(h2, h3));
    reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack_u16x2(l0, l1), pack_u16x2(l2, l3));
  }
}

int launch_fb_planes(const float* wav, long long n, void* hi, void* lo, cudaStream_t st) {
  ProfScope _ps("fbank_planes", st);
  if (n % 4) {
    set_error("fbank: sample count must be a multiple of 4");
    return -1;
  }
  const long long want = ((n >> 2) + 255) / 256;
  fb_planes_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, st>>>(wav, n, reinterpret_cast<uint16_t*>(hi),
                                                                             reinterpret_cast<uint16_t*>(lo), split_f16());
  DG_LAUNCHED();
  return 0;
}

// spectrum rows [B * rows_per_item][ld] (re 0..256 | im 257..513) -> log mel energies [B][T][80]; one warp per frame
__global__ void __launch_bounds__(256) fb_mel_kernel(const float* __restrict__ spec, int ld, int rows_per_item, int T, int B,
                                                     const float* __restrict__ banks /*[80][257]*/, const int* __restrict__ k_lo,
                                                     const int* __restrict__ k_hi, float* __restrict__ logmel) {
  __shared__ float pw[8][260];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long frame = (long long)blockIdx.x * 8 + warp;
  if (frame >= (long long)B * T) return;
  const int b = (int)(frame / T), t = (int)(frame - (long long)b * T);
  const float* row = spec + ((size_t)b * rows_per_item + t) * ld;
  for (int k = lane; k < 257; k += 32) {
    const float re = row[k], im = row[257 + k];
    pw[warp][k] = re * re + im * im;
  }
  __syncwarp();
  for (int m = lane; m < 80; m += 32) {
    float acc = 0.f;
    const float* bk = banks + m * 257;
    for (int k = k_lo[m]; k < k_hi[m]; k++) acc = fmaf(bk[k], pw[warp][k], acc);
    logmel[(size_t)frame * 80 + m] = logf(fmaxf(acc, 1.1920928955078125e-07f));     // max(mel, float32 eps)
  }
}

int launch_fb_mel(const float* spec, int ld, int rows_per_item, int T, int B, const float* banks, const int* k_lo, const int* k_hi,
                  float* logmel, cudaStream_t st) {
  ProfScope _ps("fbank_mel", st);
  const long long frames = (long long)B * T;
  fb_mel_kernel<<<(int)((frames + 7) / 8), 256, 0, st>>>(spec, ld, rows_per_item, T, B, banks, k_lo, k_hi, logmel);
  DG_LAUNCHED();
  return 0;
}

// per (item, mel bin): mean over the T frames (feats - feats.mean(dim=1), float32 like torch)
__global__ void __launch_bounds__(320) fb_mean_kernel(const float* __restrict__ logmel, int T, float* __restrict__ mean) {
  __shared__ float part[4][80];
  const int b = blockIdx.x, m = threadIdx.x % 80, q = threadIdx.x / 80;       // 4 frame groups x 80 bins
  const float* x = logmel + (size_t)b * T * 80 + m;
  float s = 0.f;
  for (int t = q; t < T; t += 4) s += x[(size_t)t * 80];
  part[q][m] = s;
  __syncthreads();
  if (q == 0) mean[b * 80 + m] = (part[0][m] + part[1][m] + part[2][m] + part[3][m]) / (float)T;
}

int launch_fb_mean(const float* logmel, int B, int T, float* mean, cudaStream_t st) {
  ProfScope _ps("fbank_mean", st);
  fb_mean_kernel<<<B, 320, 0, st>>>(logmel, T, mean);
  DG_LAUNCHED();
  return 0;
}

// stem: Conv2d(1, 32, 3, padding 1) on the mean-normalised [T][80] map + BatchNorm2d + ReLU -> planes of the padded
// [T + 2][82][32] map.  w [32][3 (dh: mel)][3 (dw: time)] as torch stores it; one thread per (w, h) position, 32 channels.
__global__ void __launch_bounds__(256) rn_stem_kernel(const float* __restrict__ logmel, const float* __restrict__ mean, int T,
                                                      const float* __restrict__ w, const float* __restrict__ sc,
                                                      const float* __restrict__ sh, uint16_t* __restrict__ hi,
                                                      uint16_t* __restrict__ lo, int f16) {
  __shared__ float ws[32 * 9], ss[32], bs[32];
  for (int i = threadIdx.x; i < 288; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 32) {
    ss[threadIdx.x] = sc[threadIdx.x];
    bs[threadIdx.x] = sh[threadIdx.x];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= T * 80) return;
  const int t = pos / 80, f = pos - t * 80;
  const float* x = logmel + (size_t)b * T * 80;
  const float* mu = mean + b * 80;
  float in[3][3];     // [dh][dw]
#pragma unroll
  for (int dh = 0; dh < 3; dh++)
#pragma unroll
    for (int dw = 0; dw < 3; dw++) {
      const int ff = f + dh - 1, tt = t + dw - 1;
      in[dh][dw] = (ff >= 0 && ff < 80 && tt >= 0 && tt < T) ? x[(size_t)tt * 80 + ff] - mu[ff] : 0.f;
    }
  const size_t row = ((size_t)b * (T + 2) + t + 1) * 82 + f + 1;
  uint32_t oh[16], ol[16];
#pragma unroll
  for (int c2 = 0; c2 < 16; c2++) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int c = 2 * c2 + e;
      float acc = 0.f;
#pragma unroll
      for (int dh = 0; dh < 3; dh++)
#pragma unroll
        for (int dw = 0; dw < 3; dw++) acc = fmaf(ws[c * 9 + dh * 3 + dw], in[dh][dw], acc);
      v[e] = fmaxf(fmaf(acc, ss[c], bs[c]), 0.f);
    }
    uint16_t h0, l0, h1, l1;
    split_h16(v[0], f16, h0, l0);
    split_h16(v[1], f16, h1, l1);
    oh[c2] = pack_u16x2(h0, h1);
    ol[c2] = pack_u16x2(l0, l1);
  }
  uint4* ph = reinterpret_cast<uint4*>(hi + row * 32);
  uint4* pl = reinterpret_cast<uint4*>(lo + row * 32);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    ph[i] = make_uint4(oh[4 * i], oh[4 * i + 1], oh[4 * i + 2], oh[4 * i + 3]);
    pl[i] = make_uint4(ol[4 * i], ol[4 * i + 1], ol[4 * i + 2], ol[4 * i + 3]);
  }
}

int launch_rn_stem(const float* logmel, const float* mean, int B, int T, const float* w, const float* sc, const float* sh,
                   void* hi, void* lo, cudaStream_t st) {
  ProfScope _ps("resnet_stem", st);
  dim3 grid((T * 80 + 255) / 256, B);
  rn_stem_kernel<<<grid, 256, 0, st>>>(logmel, mean, T, w, sc, sh, reinterpret_cast<uint16_t*>(hi), reinterpret_cast<uint16_t*>(lo),
                                       split_f16());
  DG_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------- host-side constants
// kaldi frame operator (oracle/fbank_linear.py: frame_operator): rows 0..256 real, 257..513 imaginary part of the 512-point
// DFT of hamming * preemphasis(frame - mean(frame)); float64 arithmetic, returned as float32 [514][400]
void fbank_frame_operator(std::vector<float>& op) {
  const int n = 400, padded = 512, nb = padded / 2 + 1;
  std::vector<double> m((size_t)n * n);       // m = diag(window) * pre * dc
  // dc = I - 1/n; pre = I - 0.97 * shift (x[0] - 0.97 x[0] for the first sample: replicate padding)
  // (pre * dc)[j][i] = dc[j][i] - 0.97 * dc[max(j-1,0)][i]
  for (int j = 0; j < n; j++) {
    const double win = 0.54 - 0.46 * cos(2.0 * M_PI * j / (n - 1));
    const int jp = j > 0 ? j - 1 : 0;
    for (int i = 0; i < n; i++) {
      const double dj = (i == j ? 1.0 : 0.0) - 1.0 / n, dp = (i == jp ? 1.0 : 0.0) - 1.0 / n;
      m[(size_t)j * n + i] = win * (dj - 0.97 * dp);
    }
  }
  op.assign((size_t)2 * nb * n, 0.f);
  std::vector<double> cs(padded), sn(padded);
  for (int k = 0; k < padded; k++) {
    cs[k] = cos(2.0 * M_PI * k / padded);
    sn[k] = sin(2.0 * M_PI * k / padded);
  }
  std::vector<double> re(n), im(n);
  for (int k = 0; k < nb; k++) {
    for (int i = 0; i < n; i++) re[i] = im[i] = 0.0;
    for (int j = 0; j < n; j++) {
      const int ph = (int)(((long long)k * j) % padded);
      const double c = cs[ph], s = -sn[ph];
      const double* mj = &m[(size_t)j * n];
      for (int i = 0; i < n; i++) {
        re[i] += c * mj[i];
        im[i] += s * mj[i];
      }
    }
    for (int i = 0; i < n; i++) {
      op[(size_t)k * n + i] = (float)re[i];
      op[(size_t)(nb + k) * n + i] = (float)im[i];
    }
  }
}

// kaldi get_mel_banks without VTLN (oracle/fbank_linear.py: mel_banks): [80][257] float32 + the non-zero range of every filter
void fbank_mel_banks(std::vector<float>& banks, std::vector<int>& k_lo, std::vector<int>& k_hi) {
  const int num_bins = 80, padded = 512, nfft = padded / 2;
  const double sf = 16000.0, low = 20.0, high = 0.5 * sf;
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double mlow = mel(low), mhigh = mel(high), delta = (mhigh - mlow) / (num_bins + 1), bw = sf / padded;
  banks.assign((size_t)num_bins * 257, 0.f);
  k_lo.assign(num_bins, 0);
  k_hi.assign(num_bins, 0);
  for (int b = 0; b < num_bins; b++) {
    const double left = mlow + b * delta, center = mlow + (b + 1.0) * delta, right = mlow + (b + 2.0) * delta;
    int lo = 257, hi = 0;
    for (int k = 0; k < nfft; k++) {
      const double m = mel(bw * k);
      const double up = (m - left) / (center - left), down = (right - m) / (right - center);
      const double v = fmax(0.0, fmin(up, down));
      banks[(size_t)b * 257 + k] = (float)v;
      if (v > 0.0) {
        lo = k < lo ? k : lo;
        hi = k + 1 > hi ? k + 1 : hi;
      }
    }
    k_lo[b] = lo < hi ? lo : 0;
    k_hi[b] = lo < hi ? hi : 0;
  }
}

}  // namespace dg
