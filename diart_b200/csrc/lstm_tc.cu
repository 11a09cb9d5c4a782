// Bidirectional LSTM recurrence of PyanNet on the tensor cores (tcgen05), bf16x3 split precision.
// (nn.LSTM(60,128,num_layers=4,bidirectional); reached from the reference through
// src/diart/models.py:131-133.)  The input projections are hoisted into gemm_tc.cu; this kernel runs the 293
// dependent steps of one layer.
//
// One CTA owns 16 batch rows of one direction and ALL 4 x 128 gate rows, so a step is
//     gates^T[512, 16] = W_hh[512, 128] . h_{t-1}^T[128, 16]          (M = gate rows, N = batch rows)
// i.e. four M=128 tiles, one per gate: TMEM lane u of the four accumulators holds i, f, g, o of hidden unit u,
// and the thread that owns that lane updates c and h of the unit with no cross-thread exchange.
// W_hh must be resident for the whole sequence: its bf16 hi plane (128 KB) lives in shared memory
// (A operand from a descriptor), its lo plane (another 128 KB) in TENSOR MEMORY (A operand from TMEM, 256 of
// the 512 columns) -- the only place left on the SM.  h_{t-1} is re-written every step by the epilogue threads
// as the B operand (hi/lo planes, 128B-swizzled K-major rows).  Products per k-step: Whi.hlo, Wlo.hhi, Whi.hhi.
//
// 288 threads: warps 0-7 epilogue (TMEM lane quadrant = warp % 4, batch columns 8*(warp/4) ..), warp 8 issues
// the TMA load of W_hi once and the 96 tcgen05.mma of every step.
#include <stdlib.h>
#include <string.h>

#include "dg_common.cuh"
#include "tc_ptx.cuh"

namespace dg {

constexpr int LT_NB = 16, LT_THREADS = 384;   // 8 epilogue warps + 4 MMA-issuing warps (one per gate)
constexpr int LT_W_BYTES = 4 * 2 * 128 * 128;          // W_hi: 4 gates x 2 k-blocks x (128 rows x 128 B)
constexpr int LT_H_TILE = LT_NB * 128;                 // one k-block of h^T: 16 rows x 128 B
constexpr int LT_H_BYTES = 2 * 2 * 2 * LT_H_TILE;      // [buffer][plane][k-block]
constexpr int LT_SMEM = LT_W_BYTES + LT_H_BYTES + 256 + 1024;
// TMEM columns: 8 accumulators of 16 (gate x k-half: eight independent accumulation chains hide the
// latency between dependent tcgen05.mma; the halves are summed in the epilogue), then 4 x 64 columns of W_lo
constexpr uint32_t LT_COL_D = 0, LT_COL_WLO = 128;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_c), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// accurate expf (<= 2 ulp) with a 2-ulp reciprocal: absolute error ~1e-7 on both gates, a third of the
// instructions of tanhf + IEEE division (the gate math is issue-bound: 2048 cells x 5 transcendentals per step)
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.f, 1.f + expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) { return 1.f - __fdividef(2.f, expf(2.f * x) + 1.f); }

__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tm_whi /*smem-resident plane (W_lo)*/,
               const uint16_t* __restrict__ w_lo /*TMEM-resident plane (W_hi), [2][512][128] bf16*/,
               const float* __restrict__ gx, int B, int T, int stride, int groups_per_dir, float* __restrict__ hout,
               int f16) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wsm = smem;                         // [gate][k-block][128 x 128 B]
  unsigned char* hsm = smem + LT_W_BYTES;            // [buffer][plane][k-block][16 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LT_W_BYTES + LT_H_BYTES);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* h_ready = bars + 2;                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.x / groups_per_dir;
  const int b0 = (blockIdx.x - dir * groups_per_dir) * LT_NB;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(mma_done, 4);
    mbar_init(&h_ready[0], 8);
    mbar_init(&h_ready[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < LT_H_BYTES / 4; i += LT_THREADS) reinterpret_cast<uint32_t*>(hsm)[i] = 0u;   // h_0 = 0
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ================================================================ TMA (once)
    if (lane == 0) {
      mbar_expect_tx(w_full, LT_W_BYTES);
      for (int g = 0; g < 4; g++)
        for (int kb = 0; kb < 2; kb++)
          tma_load_2d(wsm + (g * 2 + kb) * 16384, &tm_whi, kb * 64, dir * 512 + g * 128, w_full);
      mbar_wait(w_full, 0);
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================================================ W_lo -> tensor memory (A operand)
    // lane r of gate tile g holds W_lo[g*128 + r][0..127] as 64 packed bf16 pairs (low half = even k)
    const int r = warp * 32 + lane;
    for (int g = 0; g < 4; g++) {
      const uint4* src = reinterpret_cast<const uint4*>(w_lo + ((size_t)dir * 512 + g * 128 + r) * 128);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint4 q = src[half * 8 + i];
          v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
        }
        tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + LT_COL_WLO + g * 64 + half * 32, v);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zeroed h buffers -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 8) {
    // issuing a tcgen05.mma costs the issuing thread ~50 cycles (descriptor moves into uniform registers);
    // with N = 16 the math is only 8 cycles, so the four gates are issued by four warps in parallel
    const int g = warp - 8;
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | idesc_ab_format(f16) | ((uint32_t)(LT_NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // all descriptors are affine in (gate, k-step): one base each, compile-time offsets in 16-byte units
      const uint64_t a0 = umma_desc(smem_u32(wsm));
      const uint64_t bb0 = umma_desc(smem_u32(hsm)), bb1 = umma_desc(smem_u32(hsm + LT_H_BYTES / 2));
      const uint32_t alo0 = tmem_base + LT_COL_WLO, d0 = tmem_base + LT_COL_D;
      for (int step = 0; step < T; step++) {
        const int buf = step & 1;
        if (step > 0) {
          mbar_wait(&h_ready[buf], ((step - 1) >> 1) & 1);
          tc_fence_after();
        }
        const uint64_t b0d = buf ? bb1 : bb0;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
#pragma unroll
          for (int prod = 0; prod < 3; prod++) {
#pragma unroll
            for (int kb = 0; kb < 2; kb++) {
              {
                constexpr int kTile = 16384 >> 4, kHTile = LT_H_TILE >> 4;
                const int ks = kb * 4 + kk;
                const uint64_t a_hi = a0 + (uint64_t)((g * 2 + kb) * kTile + kk * 2);
                const uint64_t b_hi = b0d + (uint64_t)(kb * kHTile + kk * 2);
                const uint64_t b_lo = b0d + (uint64_t)(2 * kHTile + kb * kHTile + kk * 2);
                const uint32_t d = d0 + (g * 2 + kb) * LT_NB;
                // the HI plane of W_hh is the one in tensor memory: two of the three products then read their A
                // operand from TMEM and only one (W_lo . h_hi) pays the 4 KB shared-memory read of an SS MMA
                if (prod == 0) umma_bf16_ts(d, alo0 + g * 64 + ks * 8, b_lo, idesc, kk != 0);
                else if (prod == 1) umma_bf16(d, a_hi, b_hi, idesc, 1);
                else umma_bf16_ts(d, alo0 + g * 64 + ks * 8, b_hi, idesc, 1);
              }
            }
          }
        }
        umma_commit(mma_done);
      }
    }
  } else if (warp < 8) {
    // ================================================================ gate math / state update (warps 0..7)
    const int quad = warp & 3, ch = warp >> 2;
    const int u = quad * 32 + lane;                 // hidden unit == TMEM lane
    float c[8];
#pragma unroll
    for (int n = 0; n < 8; n++) c[n] = 0.f;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + LT_COL_D + ch * 8;
    const int kb_u = u >> 6, kq = u & 63, chunk = kq >> 3, e2 = (kq & 7) * 2;
    for (int step = 0; step < T; step++) {
      const int t = dir == 0 ? step : T - 1 - step;
      const int nxt = (step + 1) & 1;
      float xg[4][8];
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const int b = b0 + ch * 8 + n;
        const float* p = gx + ((size_t)(b < B ? b : 0) * stride + t) * 1024 + dir * 512 + u;
#pragma unroll
        for (int g = 0; g < 4; g++) xg[g][n] = b < B ? p[g * 128] : 0.f;
      }
      mbar_wait(mma_done, step & 1);
      tc_fence_after();
      uint32_t ri[8], rf[8], rg[8], ro[8], si[8], sf[8], sg[8], so[8];
      tmem_ld8(tlane + 0 * LT_NB, ri);
      tmem_ld8(tlane + 1 * LT_NB, si);
      tmem_ld8(tlane + 2 * LT_NB, rf);
      tmem_ld8(tlane + 3 * LT_NB, sf);
      tmem_ld8(tlane + 4 * LT_NB, rg);
      tmem_ld8(tlane + 5 * LT_NB, sg);
      tmem_ld8(tlane + 6 * LT_NB, ro);
      tmem_ld8(tlane + 7 * LT_NB, so);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < 8; n++) {   // sum the two k-half accumulators
        ri[n] = __float_as_uint(__uint_as_float(ri[n]) + __uint_as_float(si[n]));
        rf[n] = __float_as_uint(__uint_as_float(rf[n]) + __uint_as_float(sf[n]));
        rg[n] = __float_as_uint(__uint_as_float(rg[n]) + __uint_as_float(sg[n]));
        ro[n] = __float_as_uint(__uint_as_float(ro[n]) + __uint_as_float(so[n]));
      }
      unsigned char* hdst = hsm + nxt * (LT_H_BYTES / 2) + kb_u * LT_H_TILE;
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const float i_ = sigmoid_acc(__uint_as_float(ri[n]) + xg[0][n]);
        const float f_ = sigmoid_acc(__uint_as_float(rf[n]) + xg[1][n]);
        const float g_ = tanh_acc(__uint_as_float(rg[n]) + xg[2][n]);
        const float o_ = sigmoid_acc(__uint_as_float(ro[n]) + xg[3][n]);
        c[n] = fmaf(f_, c[n], i_ * g_);
        const float h = o_ * tanh_acc(c[n]);
        const int row = ch * 8 + n, b = b0 + row;
        if (b < B) hout[((size_t)b * stride + t) * 256 + dir * 128 + u] = h;
        uint16_t hh, hl;
        split_h16(h, f16, hh, hl);
        const int off = row * 128 + ((chunk ^ (row & 7)) << 4) + e2;       // 128B swizzle of the K-major row
        *reinterpret_cast<uint16_t*>(hdst + off) = hh;
        *reinterpret_cast<uint16_t*>(hdst + 2 * LT_H_TILE + off) = hl;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[nxt]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Version 3 of the recurrence (default).  Same decomposition (CTA = 16 batch rows x one direction x 512 gate rows,
// TMEM lane u = hidden unit u), but
//  * both planes of W_hh live in TENSOR MEMORY as far as they fit: hi plane of all four gates (256 columns) + lo plane of
//    gates i, f, g (192 columns) + four 16-column accumulators = 512 columns; only the lo plane of gate o stays in
//    shared memory, so 88 of the 96 tcgen05.mma of a step read their A operand from TMEM and only 8 pay the 4 KB
//    shared-memory operand read of an SS MMA (32 in version 2 -- the MMA phase was bound by exactly those reads);
//  * the MMAs are issued by ONE elected lane from warp-uniform descriptors (uniform registers, back-to-back UTCHMMA);
//  * h_t is the B operand in MN-MAJOR (batch-contiguous) layout: the thread that owns hidden unit u holds h_t[u] of its
//    8 batch rows, which is exactly one 16-byte unit of an MN-major core matrix -> one 16-byte shared store per plane
//    and thread instead of 16 scattered 2-byte stores (the fence.proxy.async that follows costs a MEMBAR whose latency
//    grows with the stores in flight);
//  * the cell update needs 7 MUFU operations instead of 10 (one reciprocal for f*c + i*g, one for o*tanh(c)), ~45
//    instead of ~135 instructions per cell, and is branch-free for full tiles so the 8 cells of a thread overlap;
//  * rows past the batch are skipped (batch-1 latency).
constexpr int L3_THREADS = 288;                        // 8 cell-update warps + 1 issuing warp
constexpr int L3_WS_BYTES = 2 * 128 * 128;             // W_lo of gate o: 2 k-blocks x (128 rows x 128 B)
constexpr int L3_SMEM = L3_WS_BYTES + LT_H_BYTES + 256 + 1024;
constexpr uint32_t L3_COL_D = 0, L3_COL_WHI = 64, L3_COL_WLO = 320;
constexpr int L3_PLANE = 128 * LT_NB * 2;              // one plane of h_t: 128 units x 16 rows x 2 B = 4 KB

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct L3Cell {
  const float* gp;     // gate pre-activations of (first row, t, dir, u)
  float* hp;           // output h of (first row, t, dir, u)
  size_t row_gx, row_h;
  ptrdiff_t dgx, dh;
  uint32_t tlane;      // TMEM address of this thread's accumulator columns
  uint32_t h_addr;     // shared address of this thread's 16-byte unit in buffer 0, hi plane (MN-major) / row 0 (K-major)
  int rows, ch;
};

// the 293 dependent cell updates of one thread; FULL = all 8 batch rows of this warp are valid (no branches: the
// eight independent dependency chains overlap)
template <bool F16, bool BMN, bool FULL>
__device__ __forceinline__ void l3_cell_loop(L3Cell s, int T, uint64_t* mma_done, uint64_t* h_ready, int lane) {
  constexpr int f16 = F16 ? 1 : 0;
  const float L2E = 1.4426950408889634f;
  float c[8];
#pragma unroll
  for (int n = 0; n < 8; n++) c[n] = 0.f;
  for (int step = 0; step < T; step++) {
    const int nxt = (step + 1) & 1;
    float xg[4][8];
#pragma unroll
    for (int n = 0; n < 8; n++) {
      if (FULL || n < s.rows) {
#pragma unroll
        for (int g = 0; g < 4; g++) xg[g][n] = __ldg(s.gp + n * s.row_gx + g * 128);
      } else {
#pragma unroll
        for (int g = 0; g < 4; g++) xg[g][n] = 0.f;
      }
    }
    mbar_wait(mma_done, step & 1);
    tc_fence_after();
    uint32_t ri[8], rf[8], rg[8], ro[8];
    tmem_ld8(s.tlane + 0 * LT_NB, ri);
    tmem_ld8(s.tlane + 1 * LT_NB, rf);
    tmem_ld8(s.tlane + 2 * LT_NB, rg);
    tmem_ld8(s.tlane + 3 * LT_NB, ro);
    tmem_ld_wait();
    float h[8];
    uint16_t hh[8], hl[8];
#pragma unroll
    for (int n = 0; n < 8; n++) {
      if (FULL || n < s.rows) {
        // e^-i, e^-f, e^2g, e^-o (exponents capped at 2^40: sigmoid floor 9e-13, products stay below 2^127)
        const float ei = ex2_approx(fminf((__uint_as_float(ri[n]) + xg[0][n]) * -L2E, 40.f));
        const float ef = ex2_approx(fminf((__uint_as_float(rf[n]) + xg[1][n]) * -L2E, 40.f));
        const float eg = ex2_approx(fminf((__uint_as_float(rg[n]) + xg[2][n]) * (2.f * L2E), 40.f));
        const float eo = ex2_approx(fminf((__uint_as_float(ro[n]) + xg[3][n]) * -L2E, 40.f));
        // c' = c / (1 + ef) + (eg - 1) / ((1 + ei)(1 + eg))  over one common denominator
        const float df = 1.f + ef, p = (1.f + ei) * (1.f + eg);
        c[n] = fmaf(c[n], p, (eg - 1.f) * df) * rcp_approx(p * df);
        // h = tanh(c') / (1 + eo)
        const float ec = ex2_approx(fminf(c[n] * (2.f * L2E), 40.f));
        h[n] = (ec - 1.f) * rcp_approx((1.f + eo) * (ec + 1.f));
        split_h16(h[n], f16, hh[n], hl[n]);
      } else {
        h[n] = 0.f;
        hh[n] = hl[n] = 0;
      }
    }
    const uint32_t dst = s.h_addr + nxt * (LT_H_BYTES / 2);
    if (BMN) {
      st_shared_v4(dst, pack_u16x2(hh[0], hh[1]), pack_u16x2(hh[2], hh[3]), pack_u16x2(hh[4], hh[5]), pack_u16x2(hh[6], hh[7]));
      st_shared_v4(dst + L3_PLANE, pack_u16x2(hl[0], hl[1]), pack_u16x2(hl[2], hl[3]), pack_u16x2(hl[4], hl[5]),
                   pack_u16x2(hl[6], hl[7]));
    } else {
#pragma unroll
      for (int n = 0; n < 8; n++) {
        if (FULL || n < s.rows) {
          const int row = s.ch * 8 + n;
          // s.h_addr already holds k-block and the (chunk, element) position; row and the 128B swizzle are added here
          const uint32_t a = dst + row * 128;
          st_shared_u16(a ^ ((uint32_t)(row & 7) << 4), hh[n]);
          st_shared_u16((a + L3_PLANE) ^ ((uint32_t)(row & 7) << 4), hl[n]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&h_ready[nxt]);
    // the float32 copy of h_t for the next layer leaves after the hand-off: it is not on the recurrence's critical path
#pragma unroll
    for (int n = 0; n < 8; n++)
      if (FULL || n < s.rows) s.hp[n * s.row_h] = h[n];
    s.gp += s.dgx;
    s.hp += s.dh;
  }
}

template <bool F16, bool BMN>
__global__ void __launch_bounds__(L3_THREADS, 1)
lstm_tc3_kernel(const __grid_constant__ CUtensorMap tm_wlo, const uint16_t* __restrict__ w_hi,
                const uint16_t* __restrict__ w_lo /*both [2][512][128]*/, const float* __restrict__ gx, int B, int T,
                int stride, int groups_per_dir, float* __restrict__ hout, int mn_swap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wsm = smem;                         // [k-block][128 x 128 B]   (W_lo, gate o)
  unsigned char* hsm = smem + L3_WS_BYTES;           // [buffer][plane][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L3_WS_BYTES + LT_H_BYTES);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* h_ready = bars + 2;                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int dir = blockIdx.x / groups_per_dir;
  const int b0 = (blockIdx.x - dir * groups_per_dir) * LT_NB;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(&h_ready[0], 8);
    mbar_init(&h_ready[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < LT_H_BYTES / 4; i += L3_THREADS) reinterpret_cast<uint32_t*>(hsm)[i] = 0u;   // h_0 = 0
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(w_full, L3_WS_BYTES);
      for (int kb = 0; kb < 2; kb++) tma_load_2d(wsm + kb * 16384, &tm_wlo, kb * 64, dir * 512 + 3 * 128, w_full);
      mbar_wait(w_full, 0);
    }
    __syncwarp();
  } else if (warp < 4) {
    // lane r of gate tile g holds W[g*128 + r][0..127] as 64 packed 16-bit pairs (low half = even k)
    const int r = warp * 32 + lane;
    for (int plane = 0; plane < 2; plane++) {
      for (int g = 0; g < (plane ? 3 : 4); g++) {
        const uint4* src = reinterpret_cast<const uint4*>((plane ? w_lo : w_hi) + ((size_t)dir * 512 + g * 128 + r) * 128);
        const uint32_t col = (plane ? L3_COL_WLO : L3_COL_WHI) + g * 64;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint4 q = src[half * 8 + i];
            v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
          }
          tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + col + half * 32, v);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zeroed h buffers -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 8) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | idesc_ab_format(F16 ? 1 : 0) | (BMN ? (1u << 16) : 0u) |
                             ((uint32_t)(LT_NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_s = umma_desc(smem_u32(wsm));
      // B operand (h_t): MN-major = [k-group of 8][row-group of 8][8 k x 16 B]: LBO (k-groups) 256 B, SBO (row groups) 128 B,
      // one k-step (16 k) = 512 B;  K-major = 128B-swizzled rows, two k-blocks of 16 rows x 128 B per plane
      const uint32_t h0 = smem_u32(hsm);
      const uint32_t lbo = mn_swap ? 128 : 256, sbo = mn_swap ? 256 : 128;     // (mn_swap: diagnostic A/B only)
      const uint64_t bb0 = BMN ? umma_desc_mn(h0, lbo, sbo) : umma_desc(h0);
      const uint64_t bb1 = BMN ? umma_desc_mn(h0 + LT_H_BYTES / 2, lbo, sbo) : umma_desc(h0 + LT_H_BYTES / 2);
      for (int step = 0; step < T; step++) {
        const int buf = step & 1;
        if (step > 0) {
          mbar_wait(&h_ready[buf], ((step - 1) >> 1) & 1);
          tc_fence_after();
        }
        const uint64_t b0d = buf ? bb1 : bb0;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          constexpr int kTile = 16384 >> 4, kHTile = LT_H_TILE >> 4;
          const int kb = ks >> 2, kk = ks & 3;
          const uint64_t b_hi = b0d + (uint64_t)(BMN ? ks * (512 >> 4) : kb * kHTile + kk * 2);
          const uint64_t b_lo = b_hi + (uint64_t)(L3_PLANE >> 4);
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const uint32_t d = tmem_base + L3_COL_D + g * LT_NB;
            const uint32_t a_hi = tmem_base + L3_COL_WHI + g * 64 + ks * 8;
            umma_bf16_ts(d, a_hi, b_lo, idesc, ks != 0);                                     // W_hi . h_lo
            if (g < 3) umma_bf16_ts(d, tmem_base + L3_COL_WLO + g * 64 + ks * 8, b_hi, idesc, 1);   // W_lo . h_hi
            else umma_bf16(d, a_s + (uint64_t)(kb * kTile + kk * 2), b_hi, idesc, 1);
            umma_bf16_ts(d, a_hi, b_hi, idesc, 1);                                           // W_hi . h_hi
          }
        }
        umma_commit(mma_done);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ cell update (warps 0..7)
    const int quad = warp & 3, ch = warp >> 2;
    const int u = quad * 32 + lane;                 // hidden unit == TMEM lane
    L3Cell s;
    s.rows = min(8, max(0, B - (b0 + ch * 8)));     // valid batch rows of this warp's 8 columns
    s.ch = ch;
    s.tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + L3_COL_D + ch * 8;
    s.row_gx = (size_t)stride * 1024;
    s.row_h = (size_t)stride * 256;
    const int t0 = dir == 0 ? 0 : T - 1;
    s.gp = gx + ((size_t)(b0 + ch * 8) * stride + t0) * 1024 + dir * 512 + u;
    s.hp = hout + ((size_t)(b0 + ch * 8) * stride + t0) * 256 + dir * 128 + u;
    s.dgx = dir == 0 ? 1024 : -1024;
    s.dh = dir == 0 ? 256 : -256;
    if (BMN) {
      s.h_addr = smem_u32(hsm) + (u >> 3) * 256 + ch * 128 + (u & 7) * 16;
    } else {
      const int kb_u = u >> 6, kq = u & 63;
      s.h_addr = smem_u32(hsm) + kb_u * LT_H_TILE + ((kq >> 3) << 4) + (kq & 7) * 2;
    }
    if (s.rows == 8) l3_cell_loop<F16, BMN, true>(s, T, mma_done, h_ready, lane);
    else l3_cell_loop<F16, BMN, false>(s, T, mma_done, h_ready, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

size_t lstm_tc_plane_elems() { return (size_t)2 * 512 * 128; }

// torch weight_hh_l{L}[_reverse] ([512][128], gate order i,f,g,o) -> bf16 hi / lo planes [2][512][128]
void lstm_tc_pack_whh(const float* whh_fwd, const float* whh_bwd, uint16_t* hi, uint16_t* lo, int f16) {
  split_weights_host(whh_fwd, 512, 512, 128, hi, lo, f16);
  split_weights_host(whh_bwd, 512, 512, 128, hi + 512 * 128, lo + 512 * 128, f16);
}

int launch_lstm_layer_tc(const float* gx, const void* whh_hi, const void* whh_lo, int B, int T, int stride, float* hout,
                         cudaStream_t st) {
  ProfScope _ps("lstm_rec", st);
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return -2;
  }
  CUtensorMap tm;
  cuuint64_t dims[2] = {128, 1024};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  // smem-resident plane = lo, TMEM-resident plane = hi (see the MMA sequence in the kernel)
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(whh_lo), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed for W_hh");
    return -2;
  }
  static bool attr_done = false;
  static int version = 3, bmn = 1, mn_swap = 0;
  if (!attr_done) {
    DG_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    DG_CUDA(cudaFuncSetAttribute(lstm_tc3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
    DG_CUDA(cudaFuncSetAttribute(lstm_tc3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
    DG_CUDA(cudaFuncSetAttribute(lstm_tc3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
    DG_CUDA(cudaFuncSetAttribute(lstm_tc3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
    const char* e = getenv("DG_LSTM_V2");      // A/B switches: the version-2 kernel; K-major (scattered) h operand
    if (e && e[0] == '1') version = 2;
    e = getenv("DG_LSTM_KMAJOR");
    if (e && e[0] == '1') bmn = 0;
    e = getenv("DG_LSTM_SWAP");
    if (e && e[0] == '1') mn_swap = 1;
    attr_done = true;
  }
  const int gpd = (B + LT_NB - 1) / LT_NB;
  if (version == 3) {
    const uint16_t* ph = reinterpret_cast<const uint16_t*>(whh_hi);
    const uint16_t* pl = reinterpret_cast<const uint16_t*>(whh_lo);
    const int f16 = split_f16();
#define DG_L3(F, M) lstm_tc3_kernel<F, M><<<2 * gpd, L3_THREADS, L3_SMEM, st>>>(tm, ph, pl, gx, B, T, stride, gpd, hout, mn_swap)
    if (f16 && bmn) DG_L3(true, true);
    else if (f16) DG_L3(true, false);
    else if (bmn) DG_L3(false, true);
    else DG_L3(false, false);
#undef DG_L3
    DG_LAUNCHED();
    return 0;
  }
  lstm_tc_kernel<<<2 * gpd, LT_THREADS, LT_SMEM, st>>>(tm, reinterpret_cast<const uint16_t*>(whh_hi), gx, B, T, stride,
                                                        gpd, hout, split_f16());
  DG_LAUNCHED();
  return 0;
}

}  // namespace dg
