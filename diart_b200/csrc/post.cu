// Post-path of SpeakerDiarization.__call__ on the device (reference src/diart/blocks/diarization.py:205-232):
//
//   SpeakerMap.apply           permuted[:, g] = seg[:, k] for every mapped local speaker      mapping.py:341-360
//   DelayedAggregation         Hamming-weighted average of the `latency / step` most recent permuted buffers over the
//                              region that ends `latency` before the newest buffer's end        aggregation.py:73-92,120-218
//                              (+ the first-buffer prepend rule, aggregation.py:188-212)
//   Binarize                   scores > tau, run-length encoded into (speaker, on, off) turns    blocks/utils.py:11-59
//
// The frame ranges (`SlidingWindow.crop(mode="loose", fixed=...)`, pyannote.core) are float64 index arithmetic on chunk
// start times that only the host knows; the host passes them per chunk as a small integer plan (diart_b200/blocks/post.py),
// the device does everything that touches scores.  Arithmetic follows numpy statement by statement in float64 WITHOUT
// fused multiply-add -- np.sum(ham * val, axis=0) / np.sum(ham, axis=0) adds the buffers in order -- so the thresholded
// result is bit-identical to the reference's, not merely close.
//
// One CTA per chunk.  Output: per chunk {offset, count, frames} and a packed turn list (speaker << 20 | on << 10 | off),
// each chunk's turns contiguous, ordered by speaker then time (the order Binarize emits them).
#include "dg_common.cuh"

namespace dg {

constexpr int POST_THREADS = 256;

// plan row (int32): [0] nb buffers aggregated, [1] nf frames of the region crop, [2] first_nf (> 0: first buffer of a
// stream, output = crop of [0, region.end) with its last nf frames replaced), [3] first_lo, [4 ..] lo of each buffer
__global__ void __launch_bounds__(POST_THREADS)
post_kernel(const float* __restrict__ seg, const int32_t* __restrict__ map, const float* __restrict__ hist_seg,
            const int32_t* __restrict__ hist_map, int n_hist, int B, int F, int K, int M, int nw,
            const int32_t* __restrict__ plan, int plan_stride, const double* __restrict__ hamming, double tau,
            int32_t* __restrict__ header /*[B][4]*/, uint32_t* __restrict__ turns, int turn_cap,
            unsigned int* __restrict__ total) {
  extern __shared__ unsigned char sm_raw[];
  const int c = blockIdx.x;
  const int32_t* pl = plan + (size_t)c * plan_stride;
  const int nb = pl[0], nf = pl[1], first_nf = pl[2], first_lo = pl[3];
  const int nfo = first_nf > 0 ? first_nf : nf;
  signed char* inv = reinterpret_cast<signed char*>(sm_raw);           // [nb][M]: local speaker of global g, or -1
  unsigned char* act = sm_raw + ((nw * M + 15) & ~15);                    // [nfo][M]
  __shared__ int cnt[64], off[65];
  __shared__ unsigned int base_s;

  // buffer j of this chunk = virtual chunk v = c - (nb - 1) + j; v < 0 lives in the history (last n_hist chunks seen)
  auto buf_seg = [&](int j) -> const float* {
    const int v = c - (nb - 1) + j;
    return v >= 0 ? seg + (size_t)v * F * K : hist_seg + (size_t)(n_hist + v) * F * K;
  };
  auto buf_map = [&](int j) -> const int32_t* {
    const int v = c - (nb - 1) + j;
    return v >= 0 ? map + (size_t)v * K : hist_map + (size_t)(n_hist + v) * K;
  };
  for (int i = threadIdx.x; i < nb * M; i += POST_THREADS) inv[i] = -1;
  __syncthreads();
  if (threadIdx.x < nb) {
    const int32_t* mp = buf_map(threadIdx.x);
    for (int k = 0; k < K; k++) {          // ascending k: a later local speaker overwrites (as the reference's loop would)
      const int g = mp[k];
      if (g >= 0 && g < M) inv[threadIdx.x * M + g] = (signed char)k;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nfo * M; i += POST_THREADS) {
    const int fo = i / M, g = i - fo * M;
    double v;
    const int fa = fo - (nfo - nf);          // frame of the aggregated part
    if (fa < 0) {                             // prepended part of the very first buffer: raw permuted scores
      int idx = first_lo + fo;
      idx = idx < 0 ? 0 : (idx > F - 1 ? F - 1 : idx);
      const int k = inv[g];
      v = k >= 0 ? (double)buf_seg(0)[(size_t)idx * K + k] : 0.0;
    } else {
      double num = 0.0, den = 0.0;
      for (int j = 0; j < nb; j++) {
        int idx = pl[4 + j] + fa;
        idx = idx < 0 ? 0 : (idx > F - 1 ? F - 1 : idx);    // `fixed` crops are edge-padded
        const int k = inv[j * M + g];
        const double val = k >= 0 ? (double)buf_seg(j)[(size_t)idx * K + k] : 0.0;
        const double h = hamming[idx];
        const double p = __dmul_rn(h, val);
        num = j ? __dadd_rn(num, p) : p;
        den = j ? __dadd_rn(den, h) : h;
      }
      v = __ddiv_rn(num, den);
    }
    act[i] = v > tau ? 1 : 0;
  }
  __syncthreads();
  // run-length encode per speaker (thread g), two passes around a prefix sum
  const int g = threadIdx.x;
  int n = 0;
  if (g < M) {
    int prev = 0;
    for (int f = 0; f < nfo; f++) {
      const int a = act[f * M + g];
      n += (a && !prev);
      prev = a;
    }
    cnt[g] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < M; i++) {
      off[i] = s;
      s += cnt[i];
    }
    off[M] = s;
    base_s = atomicAdd(total, (unsigned int)s);
    header[c * 4 + 0] = (int32_t)base_s;
    header[c * 4 + 1] = s;
    header[c * 4 + 2] = nfo;
    header[c * 4 + 3] = 0;
  }
  __syncthreads();
  if (g < M && n > 0) {
    size_t o = (size_t)base_s + off[g];
    int prev = 0, on = 0;
    for (int f = 0; f <= nfo; f++) {
      const int a = f < nfo ? act[f * M + g] : 0;
      if (a && !prev) on = f;
      if (!a && prev) {
        if (o < (size_t)turn_cap) turns[o] = ((uint32_t)g << 20) | ((uint32_t)on << 10) | (uint32_t)f;
        o++;
      }
      prev = a;
    }
  }
}

// the last `keep` chunks seen (history followed by this batch) become the new history
__global__ void post_history_kernel(const float* __restrict__ seg, const int32_t* __restrict__ map,
                                    const float* __restrict__ hist_seg, const int32_t* __restrict__ hist_map, int n_hist,
                                    int B, int FK, int K, int keep, float* __restrict__ new_seg,
                                    int32_t* __restrict__ new_map) {
  const int i = blockIdx.x;                       // new history slot
  const int v = B - keep + i;                     // virtual chunk
  const float* s = v >= 0 ? seg + (size_t)v * FK : hist_seg + (size_t)(n_hist + v) * FK;
  const int32_t* m = v >= 0 ? map + (size_t)v * K : hist_map + (size_t)(n_hist + v) * K;
  for (int e = threadIdx.x; e < FK; e += blockDim.x) new_seg[(size_t)i * FK + e] = s[e];
  for (int e = threadIdx.x; e < K; e += blockDim.x) new_map[(size_t)i * K + e] = m[e];
}

int launch_post(const float* seg, const int32_t* map, const float* hist_seg, const int32_t* hist_map, int n_hist, int B,
                int F, int K, int M, int nw, const int32_t* plan, int plan_stride, const double* hamming, double tau,
                int32_t* header, uint32_t* turns, int turn_cap, unsigned int* total, cudaStream_t st) {
  ProfScope _ps("post_aggregate", st);
  if (M > 64 || F > 1023 || K > 127) {
    set_error("post: at most 64 global speakers, 1023 frames");
    return -1;
  }
  const size_t smem = ((size_t)(nw * M + 15) & ~(size_t)15) + (size_t)F * M;
  if (smem > 200 * 1024) {
    set_error("post: latency / step too large for the shared-memory plan");
    return -1;
  }
  if (smem > 48 * 1024) {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool done[64] = {};
    if (dev < 64 && !done[dev]) {
      DG_CUDA(cudaFuncSetAttribute(post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      done[dev] = true;
    }
  }
  post_kernel<<<B, POST_THREADS, smem, st>>>(seg, map, hist_seg, hist_map, n_hist, B, F, K, M, nw, plan, plan_stride, hamming,
                                            tau, header, turns, turn_cap, total);
  DG_LAUNCHED();
  return 0;
}

int launch_post_history(const float* seg, const int32_t* map, const float* hist_seg, const int32_t* hist_map, int n_hist,
                        int B, int F, int K, int keep, float* new_seg, int32_t* new_map, cudaStream_t st) {
  if (keep < 1) return 0;
  post_history_kernel<<<keep, 256, 0, st>>>(seg, map, hist_seg, hist_map, n_hist, B, F * K, K, keep, new_seg, new_map);
  DG_LAUNCHED();
  return 0;
}

// ---- device-side rearrange_audio_stream (reference src/diart/operators.py:44-100): windows of a circular sample ring
// window b = samples [r0 + b*hop, r0 + b*hop + S) of the stream, ring index = absolute sample index mod C (C % 4 == 0,
// r0 % 4 == 0, hop % 4 == 0, S % 4 == 0: 16-byte accesses never straddle the wrap)
__global__ void __launch_bounds__(256) expand_windows_kernel(const float* __restrict__ ring, long long r0, int C, int hop, int S,
                                                             float* __restrict__ wav) {
  const int b = blockIdx.y;
  const long long base = r0 + (long long)b * hop;
  float4* dst = reinterpret_cast<float4*>(wav + (size_t)b * S);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (S >> 2); i += gridDim.x * blockDim.x) {
    const int idx = (int)((base + 4LL * i) % C);
    dst[i] = *reinterpret_cast<const float4*>(ring + idx);
  }
}

int launch_expand_windows(const float* ring, long long r0, int C, int hop, int S, int B, float* wav, cudaStream_t st) {
  ProfScope _ps("expand_windows", st);
  dim3 grid(8, B);
  expand_windows_kernel<<<grid, 256, 0, st>>>(ring, r0, C, hop, S, wav);
  DG_LAUNCHED();
  return 0;
}

}  // namespace dg
