// Bidirectional LSTM recurrence of PyanNet (nn.LSTM(60,128,num_layers=4,bidirectional), SURVEY.md
// Appendix A.3; reached from the reference through src/diart/models.py:131-133).
//
// The input projections W_ih x_t + b_ih + b_hh of a whole layer are hoisted into one GEMM (gemm.cu);
// this kernel runs the 293 dependent steps.  One 2-CTA cluster owns R batch rows of one direction.
// W_hh (512 x 128 fp32 = 256 KB) does not fit one SM's shared memory, so it lives in REGISTERS:
// each CTA holds the gate rows of 64 hidden units (4 gates x 64 units x 128 = 32768 weights =
// 64 per thread x 512 threads), and the two CTAs exchange their halves of h_t through distributed
// shared memory once per step.  Gate order i,f,g,o as in PyTorch.
#include <cooperative_groups.h>

#include "dg_common.cuh"

namespace cg = cooperative_groups;

namespace dg {

constexpr int H = 128, LSTM_THREADS = 512;
// h rows are stored with 4 floats of padding after every 32 (k-slice s starts at 36*s): the four
// k-slices read by the lanes of a warp then fall into different banks (no 4-way conflict)
constexpr int HP = 144;
__host__ __device__ constexpr int hidx(int u) { return u + 4 * (u >> 5); }

// packed layout: [dir][cta][thread][64]
size_t lstm_whh_packed_floats() { return (size_t)2 * 2 * LSTM_THREADS * 64; }

// thread (cta c, tid): pair pr = tid/4, k-slice ks = tid%4; local gate rows lr = 2*pr + j, j<2;
// local row lr -> gate g = lr/64, unit u = 64*c + lr%64 -> torch row g*128 + u; k = 32*ks + i.
void lstm_pack_whh(const float* whh_fwd, const float* whh_bwd, float* packed) {
  for (int d = 0; d < 2; d++) {
    const float* w = d == 0 ? whh_fwd : whh_bwd;
    for (int c = 0; c < 2; c++)
      for (int tid = 0; tid < LSTM_THREADS; tid++) {
        const int pr = tid / 4, ks = tid % 4;
        float* o = packed + (((size_t)d * 2 + c) * LSTM_THREADS + tid) * 64;
        for (int j = 0; j < 2; j++) {
          const int lr = 2 * pr + j, g = lr / 64, u = 64 * c + lr % 64;
          for (int i = 0; i < 32; i++) o[j * 32 + i] = w[(size_t)(g * H + u) * H + 32 * ks + i];
        }
      }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int R>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
lstm_rec_kernel(const float* __restrict__ gx, const float* __restrict__ whh_packed, int B, int T, int stride,
                int clusters_per_dir, float* __restrict__ hout) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int cl = blockIdx.x >> 1;
  const int dir = cl / clusters_per_dir;
  const int b0 = (cl - dir * clusters_per_dir) * R;
  const int tid = threadIdx.x, pr = tid >> 2, ks = tid & 3;

  __shared__ __align__(16) float hbuf[2][R][HP];     // h_{t-1} of all 128 units, double buffered
  __shared__ float gates[R][256];                   // this CTA's 256 gate pre-activations
  float* peer_h = cluster.map_shared_rank(&hbuf[0][0][0], crank ^ 1);

  float w[2][32];
  {
    const float4* src = reinterpret_cast<const float4*>(whh_packed + (((size_t)dir * 2 + crank) * LSTM_THREADS + tid) * 64);
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float4 v = src[j * 8 + i];
        w[j][4 * i + 0] = v.x; w[j][4 * i + 1] = v.y; w[j][4 * i + 2] = v.z; w[j][4 * i + 3] = v.w;
      }
  }
  for (int i = tid; i < 2 * R * HP; i += LSTM_THREADS) (&hbuf[0][0][0])[i] = 0.f;
  // activation-phase role: thread a < 64*R handles (row ar, local unit au)
  const int ar = tid >> 6, au = tid & 63;
  const bool act = tid < 64 * R && (b0 + ar) < B;
  float cstate = 0.f;
  // gate rows this thread finalises after the shuffle reduction (ks == 0 lanes): lr0 = 2*pr, lr0+1
  const int lr0 = 2 * pr;
  const int gcol0 = dir * 512 + (lr0 >> 6) * H + 64 * crank + (lr0 & 63);  // column in gx of local row lr0
  cluster.sync();

  for (int step = 0; step < T; step++) {
    const int t = dir == 0 ? step : T - 1 - step;
    const int cur = step & 1, nxt = cur ^ 1;
    // prefetch the input projections of this step (rows lr0, lr0+1 are adjacent columns)
    float2 gxv[R];
    if (ks == 0) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int b = b0 + r;
        gxv[r] = b < B ? *reinterpret_cast<const float2*>(gx + ((size_t)b * stride + t) * 1024 + gcol0)
                       : make_float2(0.f, 0.f);
      }
    }
    float acc[2][R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[0][r] = acc[1][r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float4 hv[R];
#pragma unroll
      for (int r = 0; r < R; r++) hv[r] = *reinterpret_cast<const float4*>(&hbuf[cur][r][36 * ks + 4 * i]);
#pragma unroll
      for (int r = 0; r < R; r++) {
        acc[0][r] = fmaf(w[0][4 * i + 0], hv[r].x, acc[0][r]);
        acc[1][r] = fmaf(w[1][4 * i + 0], hv[r].x, acc[1][r]);
        acc[0][r] = fmaf(w[0][4 * i + 1], hv[r].y, acc[0][r]);
        acc[1][r] = fmaf(w[1][4 * i + 1], hv[r].y, acc[1][r]);
        acc[0][r] = fmaf(w[0][4 * i + 2], hv[r].z, acc[0][r]);
        acc[1][r] = fmaf(w[1][4 * i + 2], hv[r].z, acc[1][r]);
        acc[0][r] = fmaf(w[0][4 * i + 3], hv[r].w, acc[0][r]);
        acc[1][r] = fmaf(w[1][4 * i + 3], hv[r].w, acc[1][r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        float v = acc[j][r];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        acc[j][r] = v;
      }
      if (ks == 0) {
        gates[r][lr0] = acc[0][r] + gxv[r].x;
        gates[r][lr0 + 1] = acc[1][r] + gxv[r].y;
      }
    }
    __syncthreads();
    float hval = 0.f;
    if (act) {
      const float gi = gates[ar][au], gf = gates[ar][64 + au], gg = gates[ar][128 + au], go = gates[ar][192 + au];
      const float i_ = sigmoidf_(gi), f_ = sigmoidf_(gf), g_ = tanhf(gg), o_ = sigmoidf_(go);
      cstate = fmaf(f_, cstate, i_ * g_);
      const float h = o_ * tanhf(cstate);
      const int u = 64 * crank + au;
      hbuf[nxt][ar][hidx(u)] = h;
      peer_h[(nxt * R + ar) * HP + hidx(u)] = h;
      hval = h;
    }
    // split barrier: the release only has to cover the (distributed) shared-memory writes of h; the global
    // store of this step's output is issued between arrive and wait so that its latency is off the chain
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (act) hout[((size_t)(b0 + ar) * stride + t) * 256 + dir * H + 64 * crank + au] = hval;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int R>
static int launch_R(const float* gx, const float* whh, int B, int T, int stride, float* hout, cudaStream_t st) {
  const int cpd = (B + R - 1) / R;
  lstm_rec_kernel<R><<<2 * 2 * cpd, LSTM_THREADS, 0, st>>>(gx, whh, B, T, stride, cpd, hout);
  DG_LAUNCHED();
  return 0;
}

int launch_lstm_layer(const float* gx, const float* whh_packed, int B, int T, int stride, float* hout,
                      cudaStream_t st) {
  ProfScope _ps("lstm_rec", st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // rows per cluster: smallest R such that 2 directions x ceil(B/R) clusters fit in one wave
  const int clusters = sms / 2;
  int R = 1;
  while (R < 8 && 2 * ((B + R - 1) / R) > clusters) R++;
  switch (R) {
    case 1: return launch_R<1>(gx, whh_packed, B, T, stride, hout, st);
    case 2: return launch_R<2>(gx, whh_packed, B, T, stride, hout, st);
    case 3: return launch_R<3>(gx, whh_packed, B, T, stride, hout, st);
    case 4: return launch_R<4>(gx, whh_packed, B, T, stride, hout, st);
    case 5: return launch_R<5>(gx, whh_packed, B, T, stride, hout, st);
    case 6: return launch_R<6>(gx, whh_packed, B, T, stride, hout, st);
    case 7: return launch_R<7>(gx, whh_packed, B, T, stride, hout, st);
    default: return launch_R<8>(gx, whh_packed, B, T, stride, hout, st);
  }
}

}  // namespace dg
