// Shifted-window GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hi/lo split precision.
//
//     C[m, n] = epi( sum_{j<KW} sum_{c<Cin} A[m + j*dil, c] * W[n, j*Cin + c] + bias[n] )
//
// The dense contractions of the path -- the five TDNN layers of XVectorSincNet (dilated Conv1d ->
// LeakyReLU -> BatchNorm1d(eval)) and the LSTM input projections of PyanNet (SURVEY.md Appendix
// A.3/A.4; reached from the reference through src/diart/models.py:131-133) -- must stay at float32-level
// accuracy because their outputs feed hard thresholds (tau_active, rho_update, delta_new).  Each float32
// operand x is therefore carried as two 16-bit planes, hi = rn16(x) and lo = rn16(x - hi) -- fp16 by default
// (22 significand bits for the pair), bf16 with DG_SPLIT_BF16=1 (16 bits) -- and every k-step issues three
// tcgen05.mma (hi*hi + lo*hi + hi*lo) into the same float32 TMEM accumulator.  Measured against the float32
// SIMT GEMM: < 1e-5 relative (tests/test_gpu_gemm_tc.py).
//
// Because activations are stored time-major ([item][row][channel]) a Conv1d tap is just a TMA box whose
// row coordinate is shifted by j*dil: no im2col is ever materialised.
//
// CTA = 192 threads, persistent over (m_tile, n_tile):
//   warp 0      TMA producer: per k-block four boxes (A_hi, A_lo: 128 rows x 64 ch; W_hi, W_lo: BN x 64)
//               into a 128B-swizzled, NSTAGE-deep shared-memory ring, completion on full[] mbarriers
//   warp 1      MMA issuer (one elected lane): 12 tcgen05.mma per k-block into one of two TMEM
//               accumulators (128 lanes x BN fp32 columns each); tcgen05.commit frees the smem slot
//   warps 2-5   epilogue: tcgen05.ld the finished accumulator (one TMEM lane quadrant per warp), bias,
//               LeakyReLU, BatchNorm affine, then either float32 rows or the next layer's hi/lo 16-bit
//               planes; overlaps the next tile's MMAs through the second accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "dg_common.cuh"
#include "tc_ptx.cuh"

namespace dg {

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 192;

struct TcArgs {
  long long M;          // output rows
  int N;                // valid output channels
  int m_tiles, n_tiles;
  int KW, dil, cin_blocks;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  float* out_f32;       // EPI_F32*: [M, ldc]
  __nv_bfloat16* out_hi;   // EPI_*_SPLIT: [M, ldc] each
  __nv_bfloat16* out_lo;
  int ldc;
  int f16;              // operand planes are fp16 (1) or bf16 (0)
  float acc_scale;      // 1 / (power-of-two scale of the weight planes): applied to the accumulator in the epilogue
  int vec8;             // output rows are 32-byte aligned: 256-bit stores
  int tma_out;          // TC_BIAS_F32 on CTA pairs: rows leave through bulk tensor stores (tmOut)
  int tap_off[9];       // row offset of every tap (Conv1d: j * dil; Conv2d on a zero-padded map: (dw-1) * Hp + (dh-1))
  // TC_CONV2D: rows are positions (item, w, h) of a zero-padded [Wp][Hp] map; outputs go to the padded map [Wop][Hop] of the
  // next layer (stride 1: same geometry; stride 2: computed at every centre, only odd (w, h) are kept)
  int Wp, Hp, Wop, Hop, stride2, relu;
  const __nv_bfloat16* res_hi;   // residual planes in the output geometry, or null
  const __nv_bfloat16* res_lo;
  // TC_POOL: weighted statistics pooling fused into the epilogue (the activation map is never written)
  const float* pool_w;     // [rows][4]: pooling weight of (row, speaker), zero for rows past an item's valid frames
  float* pool_part;        // [m_tiles][2 (item of the tile)][4 (speaker)][2 (sum w d, sum w d^2)][N]
  int pool_item_rows, pool_K;
  // TC_MAXPOOL3: m-tiles advance by `tile_rows` <= 126 rows (whole pooling windows, a divisor of the item's rows, so that every
  // item is summed in the same grouping wherever it sits in the batch; the MMA still covers 128 rows), out_f32 receives
  // bias + MaxPool1d(3) over rows ([M / 3, ldc]), pool_part the per-tile InstanceNorm partial sums of the pre-bias pooled values:
  // [m_tiles][2 (item of the tile)][2 (sum, sum of squares)][N] over the pooled frames < pool3_T of an item
  int tile_rows, pool3_T;
};

enum TcEpi { TC_BIAS_F32 = 0, TC_LEAKY_BN_SPLIT = 1, TC_LEAKY_BN_F32 = 2, TC_CONV2D = 3, TC_POOL = 4, TC_MAXPOOL3 = 5 };

// ------------------------------------------------------------------------------------ the kernel
// WRES: the whole W operand (all k-blocks, hi + lo planes) stays in shared memory for the lifetime of the persistent CTA -- the
// 64-wide SincNet convolutions re-fetched 112 / 80 KB of W per 111-row tile and are bound by the L2->SM rate.  Layout:
// [W: TC_WRES_KB k-blocks x (hi, lo)] [NSTAGE stages of A (hi, lo)] [parameters] [barriers] [epilogue staging]
constexpr int TC_WRES_KB = 7;
template <int BN, bool WRES = false>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;     // 16 KB per plane
  static constexpr int W_BYTES = BN * TC_BK * 2;
  static constexpr int WRES_BYTES = WRES ? TC_WRES_KB * 2 * W_BYTES : 0;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + (WRES ? 0 : 2 * W_BYTES);
  static constexpr int NSTAGE = WRES ? 3 : (BN == 256 ? 2 : (BN == 128 ? 3 : 4));   // BN = 64 / 32: 4 stages
  static constexpr int PARAM_BYTES = 3 * BN * 4;
  // TC_POOL: activation chunk [128][33] + row weights [128][4] + cross-row-group staging [4][2][8][32], all float
  static constexpr int POOL_BYTES = (128 * 33 + 128 * 4 + 4 * 2 * 8 * 32) * 4;
  // TC_MAXPOOL3: accumulator chunk [128][33] + cross-row-group staging [4][2][2][32], all float
  static constexpr int POOL3_BYTES = (128 * 33 + 4 * 2 * 2 * 32) * 4;
  static constexpr int TOTAL = WRES_BYTES + NSTAGE * STAGE_BYTES + PARAM_BYTES + 256 + 1024;   // + barriers + alignment slack
  static constexpr int extra(int epi) { return epi == 4 ? POOL_BYTES : (epi == 5 ? POOL3_BYTES : 0); }
};

// ------------------------------------------------------------------------------------ the epilogue of one 128-row tile
// Executed by the 128 epilogue threads of a CTA (four warps, TMEM lane quadrant = warp % 4).  `mt` = index of the 128-row tile
// (rows mt * 128 ..), `tmem_acc` = TMEM address of the tile's accumulator (lane 0), `acc_full_bar` is waited for before the
// accumulator is read (after the parameter staging and the residual prefetch).
// `tm_out` (TC_BIAS_F32 on CTA pairs): the float32 rows leave through bulk tensor stores -- each warp stages its 32 rows x 32
// columns in shared memory (`out_stage`, 1024-byte aligned, 4 KB per warp, 128-byte swizzle) and one lane issues the store;
// per-lane 32-byte global stores to 32 different rows kept this epilogue at 1450 cycles per 32-column chunk (lstm_inproj)
template <int BN, int EPI>
__device__ __forceinline__ void tc_epilogue_tile(const TcArgs& a, float* params, float* pool_stage, uint32_t tmem_acc, long long mt,
                                                 int n0, int quad, int lane, int et, bool stage_params, uint64_t* acc_full_bar,
                                                 int acc_phase, const CUtensorMap* tm_out = nullptr,
                                                 unsigned char* out_stage = nullptr) {
    const long long m = mt * (EPI == TC_MAXPOOL3 ? a.tile_rows : TC_BM) + quad * 32 + lane;
    // stage the per-column parameters of this tile (named barrier 1: the 128 epilogue threads only); with a single
    // column tile they are the same for every tile of this CTA: staged once
    if (stage_params) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < BN; i += 128) {
        const int n = n0 + i;
        const bool ok = n < a.N;
        params[i] = (ok && a.bias) ? a.bias[n] : 0.f;
        constexpr bool has_bn = EPI != TC_BIAS_F32 && EPI != TC_MAXPOOL3;
        params[BN + i] = (ok && has_bn) ? a.bn_scale[n] * (EPI == TC_CONV2D ? a.acc_scale : 1.f) : 1.f;   // (2^-k: exact)
        params[2 * BN + i] = (ok && has_bn) ? a.bn_shift[n] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // TC_CONV2D: output position of this row, and the residual of the first 32 columns requested before the wait
    long long mo = m;                 // output row
    bool row_ok = m < a.M;
    uint4 rh0[4], rl0[4];
    if (EPI == TC_CONV2D) {
      const unsigned mu = (unsigned)m, per = (unsigned)(a.Wp * a.Hp);     // (the launcher checks M < 2^31)
      const unsigned item = mu / per, rem = mu - item * per;
      const int w = (int)(rem / (unsigned)a.Hp), h = (int)(rem - (unsigned)w * (unsigned)a.Hp);
      row_ok = row_ok && w >= 1 && w <= a.Wp - 2 && h >= 1 && h <= a.Hp - 2;     // a centre inside the un-padded map
      if (a.stride2) {
        row_ok = row_ok && (w & 1) && (h & 1);
        mo = ((long long)item * a.Wop + ((w - 1) >> 1) + 1) * a.Hop + ((h - 1) >> 1) + 1;
      }
      if (row_ok && a.res_hi) {
        const uint4* rh = reinterpret_cast<const uint4*>(a.res_hi + mo * a.ldc + n0);
        const uint4* rl = reinterpret_cast<const uint4*>(a.res_lo + mo * a.ldc + n0);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          rh0[q] = rh[q];
          rl0[q] = rl[q];
        }
      }
    }
    mbar_wait(acc_full_bar, acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16);
    // TC_POOL: the rows' pooling weights go to shared memory; `brow` = first row of the tile that belongs to the NEXT item
    // (a 128-row tile covers at most two items)
    int brow = TC_BM;
    if (EPI == TC_POOL) {
      float4 pw = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) pw = *reinterpret_cast<const float4*>(a.pool_w + m * 4);
      reinterpret_cast<float4*>(pool_stage + 128 * 33)[quad * 32 + lane] = pw;     // row of the tile = TMEM lane
      const long long first = (long long)mt * TC_BM;
      const long long nxt = (first / a.pool_item_rows + 1) * a.pool_item_rows;
      brow = nxt - first < TC_BM ? (int)(nxt - first) : TC_BM;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      if (EPI != TC_POOL && EPI != TC_MAXPOOL3 && n0 + c >= a.N) continue;
      float v[32];
      if (EPI == TC_POOL) {
        // bias -> LeakyReLU -> BatchNorm affine, then the deviation from the per-channel pivot (the BatchNorm shift) goes to
        // shared memory; thread (row group rg, column col) then sums its 32 rows for the K speakers -- independent
        // accumulators, no cross-lane traffic -- split at `brow` between the tile's two items
        float* dsm = pool_stage;                    // [128][33]
        const float4* wsm = reinterpret_cast<const float4*>(pool_stage + 128 * 33);
        float* stg = pool_stage + 128 * 33 + 128 * 4;   // [rg 4][item 2][8][32]
#pragma unroll
        for (int q4 = 0; q4 < 8; q4++) {       // parameters as 128-bit loads
          const float4 b4 = *reinterpret_cast<const float4*>(params + c + 4 * q4);
          const float4 s4 = *reinterpret_cast<const float4*>(params + BN + c + 4 * q4);
          const float4 h4 = *reinterpret_cast<const float4*>(params + 2 * BN + c + 4 * q4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float x = leaky(fmaf(__uint_as_float(r[4 * q4 + e]), a.acc_scale, bb[e]));
            dsm[(quad * 32 + lane) * 33 + 4 * q4 + e] = fmaf(x, ss[e], hh[e]) - hh[e];
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
          const int rg = et >> 5, col = et & 31, r_lo = rg * 32, r_hi = r_lo + 32;
#pragma unroll
          for (int sg = 0; sg < 2; sg++) {
            const int lo = sg == 0 ? r_lo : max(r_lo, brow), hi = sg == 0 ? min(r_hi, brow) : r_hi;
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            if (a.pool_K <= 3) {               // the usual three local speakers: the fourth weight is not touched
#pragma unroll 8
              for (int rr = lo; rr < hi; rr++) {
                const float dv = dsm[rr * 33 + col];
                const float4 w4 = wsm[rr];
                const float a0 = w4.x * dv, a1 = w4.y * dv, a2 = w4.z * dv;
                s1[0] += a0; s1[1] += a1; s1[2] += a2;
                s2[0] = fmaf(a0, dv, s2[0]); s2[1] = fmaf(a1, dv, s2[1]); s2[2] = fmaf(a2, dv, s2[2]);
              }
            } else {
#pragma unroll 8
              for (int rr = lo; rr < hi; rr++) {
                const float dv = dsm[rr * 33 + col];
                const float4 w4 = wsm[rr];
                const float a0 = w4.x * dv, a1 = w4.y * dv, a2 = w4.z * dv, a3 = w4.w * dv;
                s1[0] += a0; s1[1] += a1; s1[2] += a2; s1[3] += a3;
                s2[0] = fmaf(a0, dv, s2[0]); s2[1] = fmaf(a1, dv, s2[1]); s2[2] = fmaf(a2, dv, s2[2]); s2[3] = fmaf(a3, dv, s2[3]);
              }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
              stg[((rg * 2 + sg) * 8 + 2 * k) * 32 + col] = s1[k];
              stg[((rg * 2 + sg) * 8 + 2 * k + 1) * 32 + col] = s2[k];
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 2 items x K speakers x 2 sums x 32 columns: the four row groups' totals are added in a fixed order
        {
          const int col = et & 31, twoK = 2 * a.pool_K;
          for (int q = et >> 5; q < 2 * twoK; q += 4) {
            const int sg = q >= twoK ? 1 : 0, j = q - sg * twoK;
            const float tot = ((stg[((0 * 2 + sg) * 8 + j) * 32 + col] + stg[((1 * 2 + sg) * 8 + j) * 32 + col]) +
                               stg[((2 * 2 + sg) * 8 + j) * 32 + col]) + stg[((3 * 2 + sg) * 8 + j) * 32 + col];
            const int n = n0 + c + col;
            if (n < a.N && mt < a.m_tiles) a.pool_part[(((size_t)mt * 2 + sg) * 8 + j) * a.N + n] = tot;
          }
        }
        continue;
      }
      if (EPI == TC_MAXPOOL3) {
        // bias + MaxPool1d(3) over the rows of the tile (42 windows of three TMEM lanes: through shared memory), and the
        // InstanceNorm partial sums of the pooled values, split at `brow3` between the tile's two items
        float* dsm = pool_stage;                    // [128][33]
        float* stg = pool_stage + a.tile_rows * 33; // [rg 4][item 2][2][32]
        if (quad * 32 + lane < a.tile_rows) {        // (the staging buffer holds tile_rows rows)
#pragma unroll
          for (int i = 0; i < 32; i++) dsm[(quad * 32 + lane) * 33 + i] = __uint_as_float(r[i]) * a.acc_scale;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
          const int col = et & 31, rg = et >> 5, n = n0 + c + col;
          const long long first = mt * (long long)a.tile_rows;                  // first un-pooled row of the tile
          const long long item0 = first / a.pool_item_rows;
          const long long nxt = (item0 + 1) * a.pool_item_rows;
          const int brow3 = nxt - first < a.tile_rows ? (int)(nxt - first) / 3 : a.tile_rows / 3;   // first window of the next item
          const long long p_first = first / 3;                                  // first pooled row of the tile
          const int f0 = (int)(p_first - item0 * (a.pool_item_rows / 3));       // its frame index inside item0
          const long long Mp = a.M / 3;
          const float bias = params[c + col];
          float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
          for (int pr = rg; pr < a.tile_rows / 3; pr += 4) {
            const float v = fmaxf(fmaxf(dsm[(3 * pr) * 33 + col], dsm[(3 * pr + 1) * 33 + col]), dsm[(3 * pr + 2) * 33 + col]);
            const int sg = pr >= brow3 ? 1 : 0;
            const int frame = sg ? pr - brow3 : f0 + pr;
            const long long P = p_first + pr;
            if (P < Mp) {
              if (n < a.N) a.out_f32[P * a.ldc + n] = v + bias;
              if (frame < a.pool3_T) {
                s1[sg] += v;
                s2[sg] = fmaf(v, v, s2[sg]);
              }
            }
          }
#pragma unroll
          for (int sg = 0; sg < 2; sg++) {
            stg[((rg * 2 + sg) * 2 + 0) * 32 + col] = s1[sg];
            stg[((rg * 2 + sg) * 2 + 1) * 32 + col] = s2[sg];
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {   // 2 items x 2 sums x 32 columns = 128 values, one per thread: the four row groups in a fixed order
          const int col = et & 31, q = et >> 5, sg = q >> 1, j = q & 1, n = n0 + c + col;
          const float tot = ((stg[((0 * 2 + sg) * 2 + j) * 32 + col] + stg[((1 * 2 + sg) * 2 + j) * 32 + col]) +
                             stg[((2 * 2 + sg) * 2 + j) * 32 + col]) + stg[((3 * 2 + sg) * 2 + j) * 32 + col];
          if (n < a.N) a.pool_part[(((size_t)mt * 2 + sg) * 2 + j) * a.N + n] = tot;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the chunk buffer is rewritten by the next 32 columns
        continue;
      }
      if (EPI == TC_CONV2D) {
        // BatchNorm2d(eval) affine -> (+ residual) -> ReLU
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = fmaf(__uint_as_float(r[i]), params[BN + c + i], params[2 * BN + c + i]);
        if (row_ok && a.res_hi) {
          const uint4* rh = reinterpret_cast<const uint4*>(a.res_hi + mo * a.ldc + n0 + c);
          const uint4* rl = reinterpret_cast<const uint4*>(a.res_lo + mo * a.ldc + n0 + c);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint4 hq = c == 0 ? rh0[q] : rh[q], lq = c == 0 ? rl0[q] : rl[q];
            const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w}, lw[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              v[8 * q + 2 * e] += h16_to_f32((uint16_t)(hw[e] & 0xFFFFu), a.f16) + h16_to_f32((uint16_t)(lw[e] & 0xFFFFu), a.f16);
              v[8 * q + 2 * e + 1] += h16_to_f32((uint16_t)(hw[e] >> 16), a.f16) + h16_to_f32((uint16_t)(lw[e] >> 16), a.f16);
            }
          }
        }
        if (a.relu) {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.f);
        }
        if (row_ok) {
          if (a.out_f32) {
            float* po = a.out_f32 + mo * a.ldc + n0 + c;
#pragma unroll
            for (int i = 0; i < 8; i++)
              reinterpret_cast<float4*>(po)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (a.out_hi) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
              uint16_t h0, l0, h1, l1;
              split_h16(v[2 * i], a.f16, h0, l0);
              split_h16(v[2 * i + 1], a.f16, h1, l1);
              hi[i] = pack_u16x2(h0, h1);
              lo[i] = pack_u16x2(l0, l1);
            }
            uint4* ph = reinterpret_cast<uint4*>(a.out_hi + mo * a.ldc + n0 + c);
            uint4* pl = reinterpret_cast<uint4*>(a.out_lo + mo * a.ldc + n0 + c);
#pragma unroll
            for (int i = 0; i < 4; i++) {
              ph[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              pl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        }
        continue;
      }
      if (EPI == TC_BIAS_F32 && tm_out) {
        unsigned char* wst = out_stage + quad * 4096;           // this warp's [32 rows][128 B]
        if (lane == 0) tma_store_wait_read();                   // the previous chunk's store has read the buffer
        __syncwarp();
        const uint32_t row_addr = smem_u32(wst) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float x0 = fmaf(__uint_as_float(r[4 * j]), a.acc_scale, params[c + 4 * j]);
          const float x1 = fmaf(__uint_as_float(r[4 * j + 1]), a.acc_scale, params[c + 4 * j + 1]);
          const float x2 = fmaf(__uint_as_float(r[4 * j + 2]), a.acc_scale, params[c + 4 * j + 2]);
          const float x3 = fmaf(__uint_as_float(r[4 * j + 3]), a.acc_scale, params[c + 4 * j + 3]);
          st_shared_v4(row_addr + ((j ^ (lane & 7)) << 4), __float_as_uint(x0), __float_as_uint(x1), __float_as_uint(x2),
                       __float_as_uint(x3));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tm_out, wst, n0 + c, (int)(mt * TC_BM) + quad * 32);
          tma_store_commit();
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 32; i++) {
        float x = fmaf(__uint_as_float(r[i]), a.acc_scale, params[c + i]);
        if (EPI != TC_BIAS_F32) {
          x = leaky(x);
          x = fmaf(x, params[BN + c + i], params[2 * BN + c + i]);
        }
        v[i] = x;
      }
      if (m < a.M) {
        if (EPI == TC_LEAKY_BN_SPLIT) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; i++) {
            uint16_t h0, l0, h1, l1;
            split_h16(v[2 * i], a.f16, h0, l0);
            split_h16(v[2 * i + 1], a.f16, h1, l1);
            hi[i] = pack_u16x2(h0, h1);
            lo[i] = pack_u16x2(l0, l1);
          }
          uint4* ph = reinterpret_cast<uint4*>(a.out_hi + m * a.ldc + n0 + c);
          uint4* pl = reinterpret_cast<uint4*>(a.out_lo + m * a.ldc + n0 + c);
          if (a.vec8) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              st_global_v8(ph + 2 * i, hi[8 * i], hi[8 * i + 1], hi[8 * i + 2], hi[8 * i + 3], hi[8 * i + 4], hi[8 * i + 5],
                           hi[8 * i + 6], hi[8 * i + 7]);
              st_global_v8(pl + 2 * i, lo[8 * i], lo[8 * i + 1], lo[8 * i + 2], lo[8 * i + 3], lo[8 * i + 4], lo[8 * i + 5],
                           lo[8 * i + 6], lo[8 * i + 7]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              ph[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              pl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        } else {
          float* po = a.out_f32 + m * a.ldc + n0 + c;
          if (n0 + c + 32 <= a.N && a.vec8) {
#pragma unroll
            for (int i = 0; i < 4; i++)
              st_global_v8(po + 8 * i, __float_as_uint(v[8 * i]), __float_as_uint(v[8 * i + 1]), __float_as_uint(v[8 * i + 2]),
                           __float_as_uint(v[8 * i + 3]), __float_as_uint(v[8 * i + 4]), __float_as_uint(v[8 * i + 5]),
                           __float_as_uint(v[8 * i + 6]), __float_as_uint(v[8 * i + 7]));
          } else if (n0 + c + 32 <= a.N) {
#pragma unroll
            for (int i = 0; i < 8; i++)
              reinterpret_cast<float4*>(po)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (n0 + c + i < a.N) po[i] = v[i];
          }
        }
      }
    }
}

template <int BN, int EPI, bool WRES = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, TcArgs a) {
  using S = TcSmem<BN, WRES>;
  constexpr int NSTAGE = S::NSTAGE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem_base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wres = smem_base;                       // WRES: [k-block][hi 64 x 64 | lo 64 x 64], resident
  unsigned char* smem = smem_base + S::WRES_BYTES;       // the stage ring
  float* params = reinterpret_cast<float*>(smem + NSTAGE * S::STAGE_BYTES);          // bias | bn_scale | bn_shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * S::STAGE_BYTES + S::PARAM_BYTES);
  uint64_t* full = bars;                 // [NSTAGE] TMA -> MMA
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE] MMA -> TMA
  uint64_t* acc_full = bars + 2 * NSTAGE;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * NSTAGE + 2; // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  uint64_t* w_full = bars + 2 * NSTAGE + 5;    // WRES: the resident W operand has landed
  float* pool_stage = reinterpret_cast<float*>(smem + NSTAGE * S::STAGE_BYTES + S::PARAM_BYTES + 256);   // TC_POOL / TC_MAXPOOL3 only

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const int num_tiles = a.m_tiles * a.n_tiles;
  const int kblocks = a.KW * a.cin_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);     // one arrive per epilogue warp
    }
    if (WRES) mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: two accumulators of BN fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * BN < 32 ? 32 : 2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      if (WRES) {   // the whole W operand once (a single column tile: n0 = 0)
        mbar_expect_tx(w_full, kblocks * 2 * S::W_BYTES);
        for (int kb = 0; kb < kblocks; kb++) {
          tma_load_2d(wres + kb * 2 * S::W_BYTES, &tmW_hi, kb * TC_BK, 0, w_full);
          tma_load_2d(wres + kb * 2 * S::W_BYTES + S::W_BYTES, &tmW_lo, kb * TC_BK, 0, w_full);
        }
      }
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
        const int m0 = mt * (EPI == TC_MAXPOOL3 ? a.tile_rows : TC_BM), n0 = nt * BN;
        for (int j = 0; j < a.KW; j++) {
          for (int cb = 0; cb < a.cin_blocks; cb++) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* st = smem + stage * S::STAGE_BYTES;
            mbar_expect_tx(&full[stage], S::STAGE_BYTES);
            const int kcol = (j * a.cin_blocks + cb) * TC_BK;
            tma_load_2d(st, &tmA_hi, cb * TC_BK, m0 + a.tap_off[j], &full[stage]);
            tma_load_2d(st + S::A_BYTES, &tmA_lo, cb * TC_BK, m0 + a.tap_off[j], &full[stage]);
            if (!WRES) {
              tma_load_2d(st + 2 * S::A_BYTES, &tmW_hi, kcol, n0, &full[stage]);
              tma_load_2d(st + 2 * S::A_BYTES + S::W_BYTES, &tmW_lo, kcol, n0, &full[stage]);
            }
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // (elect_one, not lane == 0: descriptors stay in uniform registers and the tcgen05.mma issue back to back)
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=fp16 (or bf16), both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | idesc_ab_format(a.f16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      if (WRES) {
        mbar_wait(w_full, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sw = WRES ? smem_u32(wres + kb * 2 * S::W_BYTES) : sa + 2 * S::A_BYTES;
          const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + S::A_BYTES);
          const uint64_t w_hi = umma_desc(sw), w_lo = umma_desc(sw + S::W_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ks++) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // +32 bytes per 16-element k-step
            umma_f16(tmem_c, a_lo + adv, w_hi + adv, idesc, (kb | ks) != 0);
            umma_f16(tmem_c, a_hi + adv, w_lo + adv, idesc, 1);
            umma_f16(tmem_c, a_hi + adv, w_hi + adv, idesc, 1);
          }
          umma_commit(&empty[stage]);           // smem slot reusable once these MMAs have read it
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&acc_full[acc]);            // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int et = threadIdx.x - 64;            // 0..127
    int acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
      tc_epilogue_tile<BN, EPI>(a, params, pool_stage, tmem_base + acc * BN, mt, nt * BN, quad, lane, et,
                                a.n_tiles > 1 || tile == (int)blockIdx.x, &acc_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------ CTA-pair variant
// The same GEMM on PAIRS of CTAs (cluster of two, tcgen05 cta_group::2): a pair owns a 256 x BN tile; each CTA loads the A rows of
// its 128-row half and HALF of the W tile (BN / 2 rows), the even CTA issues M = 256 MMAs that read A from both CTAs' shared
// memory and the two W halves as one N = BN operand, each CTA's tensor memory receives its 128 rows.  Per SM and k-block that is
// 64 KB instead of 96 KB from L2 and 8 KB instead of 12 KB of operand reads per MMA pair -- the three-product GEMM at 128-row
// tiles is bound by exactly these two rates -- and three pipeline stages instead of two.
template <int BN>
struct TcSmem2 {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;          // 16 KB per plane (this CTA's 128 rows)
  static constexpr int W_BYTES = (BN / 2) * TC_BK * 2;       // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int NSTAGE = BN == 256 ? 3 : 4;
  static constexpr int PARAM_BYTES = 3 * BN * 4;
  static constexpr int POOL_BYTES = (128 * 33 + 128 * 4 + 4 * 2 * 8 * 32) * 4;
  static constexpr int OUT_STAGE_BYTES = 4 * 4096 + 1024;     // bulk-store epilogue: 4 warps x [32][128 B] + alignment
  static constexpr int TOTAL = NSTAGE * STAGE_BYTES + PARAM_BYTES + 256 + 1024;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                const __grid_constant__ CUtensorMap tmOut, TcArgs a) {
  using S = TcSmem2<BN>;
  constexpr int NSTAGE = S::NSTAGE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* params = reinterpret_cast<float*>(smem + NSTAGE * S::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * S::STAGE_BYTES + S::PARAM_BYTES);
  uint64_t* full = bars;                       // [NSTAGE] used in the even CTA: both CTAs' TMA -> MMA
  uint64_t* empty = bars + NSTAGE;             // [NSTAGE] in each CTA: MMA (multicast commit) -> its TMA producer
  uint64_t* acc_full = bars + 2 * NSTAGE;      // [2] in each CTA: MMA (multicast commit) -> its epilogue
  uint64_t* acc_empty = bars + 2 * NSTAGE + 2; // [2] used in the even CTA: epilogue warps of both CTAs -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  float* pool_stage = reinterpret_cast<float*>(smem + NSTAGE * S::STAGE_BYTES + S::PARAM_BYTES + 256);
  // staging of the bulk-store epilogue (TC_BIAS_F32): 4 warps x 4 KB, 1024-byte aligned (128-byte swizzle atoms)
  unsigned char* out_stage = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pool_stage) + 1023) & ~uintptr_t(1023));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  const int m_pairs = (a.m_tiles + 1) >> 1;
  const int num_tiles = m_pairs * a.n_tiles;
  const int kblocks = a.KW * a.cin_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);     // four epilogue warps in each of the two CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += pairs) {
        const int mp = tile / a.n_tiles, nt = tile - mp * a.n_tiles;
        const int m0 = (mp * 2 + rank) * TC_BM, n0 = nt * BN + rank * (BN / 2);
        for (int j = 0; j < a.KW; j++) {
          for (int cb = 0; cb < a.cin_blocks; cb++) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* st = smem + stage * S::STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);       // this CTA's bytes + the peer's
            const int kcol = (j * a.cin_blocks + cb) * TC_BK;
            tma_load_2d_2sm(st, &tmA_hi, cb * TC_BK, m0 + a.tap_off[j], &full[stage]);
            tma_load_2d_2sm(st + S::A_BYTES, &tmA_lo, cb * TC_BK, m0 + a.tap_off[j], &full[stage]);
            tma_load_2d_2sm(st + 2 * S::A_BYTES, &tmW_hi, kcol, n0, &full[stage]);
            tma_load_2d_2sm(st + 2 * S::A_BYTES + S::W_BYTES, &tmW_lo, kcol, n0, &full[stage]);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (even CTA only)
    if (rank == 0 && elect_one()) {
      // D = f32, A = B = fp16 (or bf16), both K-major, N = BN, M = 256 across the pair
      const uint32_t idesc = (1u << 4) | idesc_ab_format(a.f16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += pairs) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + S::A_BYTES);
          const uint64_t w_hi = umma_desc(sa + 2 * S::A_BYTES), w_lo = umma_desc(sa + 2 * S::A_BYTES + S::W_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ks++) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
            umma_f16_2sm(tmem_c, a_lo + adv, w_hi + adv, idesc, (kb | ks) != 0);
            umma_f16_2sm(tmem_c, a_hi + adv, w_lo + adv, idesc, 1);
            umma_f16_2sm(tmem_c, a_hi + adv, w_hi + adv, idesc, 1);
          }
          umma_commit_2sm(&empty[stage]);         // both CTAs' producers may refill the slot
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&acc_full[acc]);          // both CTAs' epilogues
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5 of both CTAs)
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;
    int acc = 0, acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += pairs) {
      const int mp = tile / a.n_tiles, nt = tile - mp * a.n_tiles;
      tc_epilogue_tile<BN, EPI>(a, params, pool_stage, tmem_base + acc * BN, (long long)mp * 2 + rank, nt * BN, quad, lane, et,
                                a.n_tiles > 1 || tile == pair, &acc_full[acc], acc_phase, a.tma_out ? &tmOut : nullptr, out_stage);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (EPI == TC_BIAS_F32 && a.tma_out && lane == 0) tma_store_wait_all();   // this lane's bulk stores are complete
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nothing of the peer is touched after this point
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
// bf16 matrix [rows, cols] row-major (cols contiguous, row pitch `ld` elements); box = 64 cols x box_rows
static int make_map(CUtensorMap* m, const void* base, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return -2;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

// DG_GEMM_1CTA=1: A/B switch, 256-wide tiles on single CTAs (the round-1 kernel) instead of CTA pairs
static bool gemm_pairs_on() {
  static const bool off = getenv("DG_GEMM_1CTA") && getenv("DG_GEMM_1CTA")[0] == '1';
  return !off;
}

template <int BN, int EPI>
static int launch_tc(const TcGemm& g, cudaStream_t st) {
  using S = TcSmem<BN>;
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  const int Ktot = g.KW * g.Cin;
  const int n_tiles = (g.N + BN - 1) / BN;
  if (make_map(&ta_hi, g.A_hi, g.Mtot, g.Cin, g.lda, TC_BM) || make_map(&ta_lo, g.A_lo, g.Mtot, g.Cin, g.lda, TC_BM) ||
      make_map(&tw_hi, g.W_hi, g.Npad, Ktot, Ktot, BN) || make_map(&tw_lo, g.W_lo, g.Npad, Ktot, Ktot, BN))
    return -2;
  TcArgs a{};
  a.M = g.M; a.N = g.N; a.n_tiles = n_tiles;
  a.tile_rows = EPI == TC_MAXPOOL3 ? g.pool3_tile_rows : TC_BM;
  a.pool3_T = g.pool3_T;
  a.m_tiles = (int)((g.M + a.tile_rows - 1) / a.tile_rows);
  a.KW = g.KW; a.dil = g.dil; a.cin_blocks = g.Cin / TC_BK;
  a.bias = g.bias; a.bn_scale = g.bn_scale; a.bn_shift = g.bn_shift;
  a.out_f32 = g.out_f32; a.out_hi = reinterpret_cast<__nv_bfloat16*>(g.out_hi);
  a.out_lo = reinterpret_cast<__nv_bfloat16*>(g.out_lo); a.ldc = g.ldc;
  a.f16 = split_f16();
  a.acc_scale = g.w_scale > 0.f ? 1.f / g.w_scale : 1.f;
  for (int j = 0; j < 9; j++) a.tap_off[j] = j < g.KW ? (g.tap_off ? g.tap_off[j] : j * g.dil) : 0;
  a.Wp = g.Wp; a.Hp = g.Hp; a.Wop = g.Wop; a.Hop = g.Hop; a.stride2 = g.stride2; a.relu = g.relu;
  a.res_hi = reinterpret_cast<const __nv_bfloat16*>(g.res_hi);
  a.res_lo = reinterpret_cast<const __nv_bfloat16*>(g.res_lo);
  a.pool_w = g.pool_w; a.pool_part = g.pool_part; a.pool_item_rows = g.pool_item_rows; a.pool_K = g.pool_K;
  {
    static const bool no_v8 = getenv("DG_NO_V8") && getenv("DG_NO_V8")[0] == '1';     // A/B switch
    const bool planes = EPI == TC_LEAKY_BN_SPLIT;
    const uintptr_t base = planes ? ((uintptr_t)g.out_hi | (uintptr_t)g.out_lo) : (uintptr_t)g.out_f32;
    a.vec8 = !no_v8 && base % 32 == 0 && (g.ldc * (planes ? 2 : 4)) % 32 == 0;
  }
  {   // function attributes are per device: one flag per (instantiation, device)
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      DG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   S::TOTAL + S::extra(EPI)));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  const int sms = g.sm_limit > 0 ? std::min(g.sm_limit, usable_sms()) : usable_sms();
  const int tiles = a.m_tiles * a.n_tiles;
  if constexpr (BN == 256 && (EPI == TC_BIAS_F32 || EPI == TC_LEAKY_BN_SPLIT || EPI == TC_LEAKY_BN_F32 || EPI == TC_POOL)) {
    if (gemm_pairs_on() && a.m_tiles >= 4) {
      using S2 = TcSmem2<BN>;
      // the pair's CTAs each load HALF of the W tile: a second pair of maps with BN / 2-row boxes
      CUtensorMap tw2_hi, tw2_lo;
      if (make_map(&tw2_hi, g.W_hi, g.Npad, Ktot, Ktot, BN / 2) || make_map(&tw2_lo, g.W_lo, g.Npad, Ktot, Ktot, BN / 2)) return -2;
      // float32 rows through bulk tensor stores (DG_NO_TMA_STORE=1: per-lane 256-bit stores)
      CUtensorMap t_out;
      memset(&t_out, 0, sizeof(t_out));
      if constexpr (EPI == TC_BIAS_F32) {
        static const bool no_tma_store = getenv("DG_NO_TMA_STORE") && getenv("DG_NO_TMA_STORE")[0] == '1';
        if (!no_tma_store && g.N % 32 == 0 && (g.ldc * 4) % 16 == 0 && (uintptr_t)g.out_f32 % 16 == 0) {
          EncodeTiledFn fn = encode_fn();
          cuuint64_t dims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
          cuuint64_t strides[1] = {(cuuint64_t)g.ldc * 4};
          cuuint32_t box[2] = {32, 32};
          cuuint32_t estr[2] = {1, 1};
          if (fn && fn(&t_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g.out_f32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            a.tma_out = 1;
        }
      }
      const int smem2 = S2::TOTAL + (EPI == TC_POOL ? S2::POOL_BYTES : 0) + (a.tma_out ? S2::OUT_STAGE_BYTES : 0);
      static bool attr2_done[64] = {};
      if (first_use_on_device(attr2_done))
        DG_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     S2::TOTAL + (EPI == TC_POOL ? S2::POOL_BYTES : 0) + (EPI == TC_BIAS_F32 ? S2::OUT_STAGE_BYTES : 0)));
      const int pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
      const int pairs = std::min(pair_tiles, sms / 2);
      gemm_tc2_kernel<BN, EPI><<<2 * pairs, TC_THREADS, smem2, st>>>(ta_hi, ta_lo, tw2_hi, tw2_lo, t_out, a);
      DG_LAUNCHED();
      return 0;
    }
  }
  const int grid = tiles < sms ? tiles : sms;
  if constexpr (BN == 64 && EPI == TC_MAXPOOL3) {
    // W resident in shared memory (DG_NO_WRES=1: off): needs one column tile, <= 7 k-blocks and a staging buffer that still fits
    static const bool no_wres = getenv("DG_NO_WRES") && getenv("DG_NO_WRES")[0] == '1';
    using SW = TcSmem<64, true>;
    const int smem_w = SW::WRES_BYTES + SW::NSTAGE * SW::STAGE_BYTES + SW::PARAM_BYTES + 256 + (a.tile_rows * 33 + 512) * 4 + 1024;
    if (!no_wres && a.n_tiles == 1 && a.KW * a.cin_blocks <= TC_WRES_KB && smem_w <= 227 * 1024) {
      static bool attrw_done[64] = {};
      if (first_use_on_device(attrw_done))
        DG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64, TC_MAXPOOL3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gemm_tc_kernel<64, TC_MAXPOOL3, true><<<grid, TC_THREADS, smem_w, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, a);
      DG_LAUNCHED();
      return 0;
    }
  }
  gemm_tc_kernel<BN, EPI><<<grid, TC_THREADS, S::TOTAL + S::extra(EPI), st>>>(ta_hi, ta_lo, tw_hi, tw_lo, a);
  DG_LAUNCHED();
  return 0;
}

int launch_gemm_tc(const TcGemm& g, cudaStream_t st) {
  ProfScope _ps(g.tag ? g.tag : "gemm_tc", st);
  if (g.Cin % TC_BK || g.lda % 8 || g.ldc % (g.epi == TC_LEAKY_BN_SPLIT ? 8 : 4) || (g.Npad % 128 && g.Npad != 64 && g.Npad != 32) ||
      g.KW < 1 || g.KW > 9) {
    set_error("gemm_tc: Cin must be a multiple of 64, A pitch a multiple of 8, output pitch a multiple of 4 "
              "(8 for bf16 planes), padded N 32, 64 or a multiple of 128, at most 9 taps");
    return -1;
  }
  const bool wide = g.Npad % 256 == 0;
  if (g.epi == TC_CONV2D) {
    if (g.ldc % 32 || g.N % 32 || g.Wp < 3 || g.Hp < 3 || (!g.out_hi && !g.out_f32) || g.M >= (1LL << 31)) {
      set_error("gemm_tc (conv2d): channel counts must be multiples of 32");
      return -1;
    }
    if (g.Npad == 32) return launch_tc<32, TC_CONV2D>(g, st);
    if (g.Npad == 64) return launch_tc<64, TC_CONV2D>(g, st);
    return wide ? launch_tc<256, TC_CONV2D>(g, st) : launch_tc<128, TC_CONV2D>(g, st);
  }
  if (g.epi == TC_POOL) {
    if (!wide || !g.pool_w || !g.pool_part || g.pool_K < 1 || g.pool_K > 4 || g.pool_item_rows < TC_BM) {
      set_error("gemm_tc (pool): needs 256-wide tiles, 1..4 speakers and items of at least 128 rows");
      return -1;
    }
    return launch_tc<256, TC_POOL>(g, st);
  }
  if (g.epi == TC_MAXPOOL3) {
    if (g.Npad != 64 || !g.out_f32 || !g.pool_part || g.pool3_T < 1 || g.pool3_tile_rows < 3 || g.pool3_tile_rows > 126 ||
        g.pool3_tile_rows % 3 || g.pool_item_rows % g.pool3_tile_rows || g.M % g.pool_item_rows) {
      set_error("gemm_tc (maxpool3): needs 64 output channels and tiles of 3..126 rows (a multiple of 3) that divide the item");
      return -1;
    }
    return launch_tc<64, TC_MAXPOOL3>(g, st);
  }
  if (g.Npad == 64 && g.epi == TC_BIAS_F32) return launch_tc<64, TC_BIAS_F32>(g, st);
  switch (g.epi) {
    case TC_BIAS_F32:
      return wide ? launch_tc<256, TC_BIAS_F32>(g, st) : launch_tc<128, TC_BIAS_F32>(g, st);
    case TC_LEAKY_BN_SPLIT:
      return wide ? launch_tc<256, TC_LEAKY_BN_SPLIT>(g, st) : launch_tc<128, TC_LEAKY_BN_SPLIT>(g, st);
    default:
      return wide ? launch_tc<256, TC_LEAKY_BN_F32>(g, st) : launch_tc<128, TC_LEAKY_BN_F32>(g, st);
  }
}

// ------------------------------------------------------------------------------------ 16-bit hi/lo split
// x [rows_in, ld_in] float32 -> hi/lo 16-bit planes [rows_out, ld_out]; optionally MaxPool1d(3) over rows
// (out row r <- max of in rows 3r..3r+2) and the previous InstanceNorm1d + LeakyReLU (scale/shift per
// (item, channel), item = out row / item_rows); channels [C, ld_out) are written as zeros.
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, long long rows_out, int C, int ld_in,
                                                    int ld_out, int pool, int item_rows, const float* __restrict__ sc,
                                                    const float* __restrict__ sh, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo, int f16, const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  const int q_per_row = ld_out >> 2;
  const long long n4 = rows_out * q_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / q_per_row;
    const int c = (int)(i - row * q_per_row) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
      if (pool) {
        const float* p = x + (row * 3) * ld_in + c;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + ld_in),
                     d = *reinterpret_cast<const float4*>(p + 2 * ld_in);
        v = make_float4(fmaxf(fmaxf(a.x, b.x), d.x), fmaxf(fmaxf(a.y, b.y), d.y), fmaxf(fmaxf(a.z, b.z), d.z),
                        fmaxf(fmaxf(a.w, b.w), d.w));
      } else {
        v = *reinterpret_cast<const float4*>(x + row * ld_in + c);
      }
      if (sc) {
        const long long item = row / item_rows;
        const float4 s = *reinterpret_cast<const float4*>(sc + item * ld_in + c);
        const float4 h = *reinterpret_cast<const float4*>(sh + item * ld_in + c);
        v.x = leaky(fmaf(v.x, s.x, h.x)); v.y = leaky(fmaf(v.y, s.y, h.y));
        v.z = leaky(fmaf(v.z, s.z, h.z)); v.w = leaky(fmaf(v.w, s.w, h.w));
      }
    }
    uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_h16(v.x, f16, h0, l0);
    split_h16(v.y, f16, h1, l1);
    split_h16(v.z, f16, h2, l2);
    split_h16(v.w, f16, h3, l3);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_u16x2(h0, h1), pack_u16x2(h2, h3));
    reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack_u16x2(l0, l1), pack_u16x2(l2, l3));
  }
}

int launch_split_ex(const float* x, long long rows_out, int C, int ld_in, int ld_out, int pool, int item_rows,
                    const float* sc, const float* sh, void* hi, void* lo, cudaStream_t st, const int* skip_flag) {
  ProfScope _ps("split16", st);
  if (C % 4 || ld_in % 4 || ld_out % 4) {
    set_error("split: channel counts must be multiples of 4");
    return -1;
  }
  const long long n4 = rows_out * (ld_out / 4);
  const long long want = (n4 + 255) / 256;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  split_kernel<<<grid, 256, 0, st>>>(x, rows_out, C, ld_in, ld_out, pool, item_rows, sc, sh,
                                     reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), split_f16(), skip_flag);
  DG_LAUNCHED();
  return 0;
}

int launch_split(const float* x, long long rows, int C, int item_rows, const float* sc, const float* sh, void* hi,
                 void* lo, cudaStream_t st) {
  return launch_split_ex(x, rows, C, C, C, 0, item_rows, sc, sh, hi, lo, st);
}

int split_f16() {
  static const int f16 = !(getenv("DG_SPLIT_BF16") && getenv("DG_SPLIT_BF16")[0] == '1');
  return f16;
}

// host-side conversions, round to nearest even (fp16: subnormals kept, finite overflow saturates like cvt.satfinite)
uint16_t host_f32_to_h16(float f, int f16) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if (!f16) return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t au = u & 0x7FFFFFFFu;
  if (au > 0x7F800000u) return (uint16_t)(sign | 0x7FFFu);                  // NaN
  if (au >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);                 // >= 65520 (or inf): largest finite
  if (au < 0x33000001u) return (uint16_t)sign;                              // <= 2^-25: rounds to zero
  const int e = (int)(au >> 23) - 127;                                      // unbiased exponent
  uint32_t mant = (au & 0x7FFFFFu) | 0x800000u;                             // 24-bit significand
  int shift = e >= -14 ? 13 : 13 + (-14 - e);                               // bits dropped
  uint32_t q = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) q++;
  // normal: q has the implicit bit at position 10 -> exponent field e + 15 (a carry out of rounding propagates by itself)
  const uint32_t bits = e >= -14 ? (uint32_t)((e + 14) << 10) + q : q;
  return (uint16_t)(sign | bits);
}
float host_h16_to_f32(uint16_t h, int f16) {
  uint32_t u;
  if (!f16) {
    u = (uint32_t)h << 16;
  } else {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
    if (e == 0) {
      const float v = (float)m * 5.9604644775390625e-08f;                  // m * 2^-24
      float r = sign ? -v : v;
      return r;
    }
    u = e == 31 ? (sign | 0x7F800000u | (m << 13)) : (sign | ((e + 112u) << 23) | (m << 13));
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// host: float32 [N][K] -> zero-padded 16-bit hi/lo planes [Npad][K]
void split_weights_host(const float* w, int N, int Npad, int K, uint16_t* hi, uint16_t* lo, int f16, float scale) {
  for (size_t i = 0; i < (size_t)Npad * K; i++) hi[i] = lo[i] = 0;
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      const float f = w[(size_t)n * K + k] * scale;          // power of two: exact
      const uint16_t h = host_f32_to_h16(f, f16);
      hi[(size_t)n * K + k] = h;
      lo[(size_t)n * K + k] = host_f32_to_h16(f - host_h16_to_f32(h, f16), f16);
    }
}

// Power-of-two scale of a weight tensor's fp16 planes: the largest magnitude lands in [2^12, 2^13), so that the lo plane
// (|lo| <= 2^-11 |w|) of every weight down to 2^-15 of the largest one stays a NORMAL fp16 number (un-scaled, lo goes
// subnormal below |w| = 0.125 and the pair keeps only an absolute 2^-25).  The accumulator is multiplied by 1 / scale in
// the epilogue (an exact operation).  bf16 planes have float32's exponent range: scale 1.  DG_NO_WSCALE=1 disables it (A/B).
float weight_plane_scale(const float* w, size_t n, int f16) {
  static const bool off = getenv("DG_NO_WSCALE") && getenv("DG_NO_WSCALE")[0] == '1';
  if (!f16 || off) return 1.f;
  float mx = 0.f;
  for (size_t i = 0; i < n; i++) {
    const float v = fabsf(w[i]);
    if (v > mx && v < 3.0e38f) mx = v;
  }
  if (!(mx > 0.f)) return 1.f;
  int e = 0;
  frexpf(mx, &e);                       // mx = m * 2^e, m in [0.5, 1)
  int k = 13 - e;                       // mx * 2^k in [2^12, 2^13)
  k = k > 40 ? 40 : (k < -40 ? -40 : k);
  return ldexpf(1.f, k);
}

}  // namespace dg
