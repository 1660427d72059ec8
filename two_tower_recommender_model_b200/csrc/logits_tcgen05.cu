// In-batch softmax on the tensor cores: the [B, B] logits S = Q C^T live only in TMEM.
//
//   forward : CTA = 128 query rows (Q tile resident in smem).  C streams through a TMA ring
//             in 128-row tiles; tcgen05.mma writes S tiles into a FOUR-stage TMEM accumulator
//             (4 x 128 columns = all of TMEM), so the MMA issuer runs up to four tiles ahead.
//             Two softmax warpgroups ping-pong: group g takes the tiles t = g (mod 2), i.e. TMEM
//             stages g and g+2; inside a group thread = row, one warp per TMEM lane quarter.
//             Each runs an online log-sum-exp in the log2 domain; the two partial (max, sum)
//             pairs of a row are merged at the end.  Output: lse[B] + per-CTA loss partials.
//   backward: "fixed-normaliser attention".  out[r,:] = scale * (sum_t P(r,t) Y_t - Y_r),
//             P = exp(X_r.Y_t/T - lse).  S tile -> TMEM -> registers -> P (bf16) written back
//             IN PLACE into TMEM -> second tcgen05.mma with A = P from TMEM (TS form) and
//             B = Y^T tile from smem accumulates O in TMEM.  Run twice: (X,Y)=(Q,C) with the
//             normaliser indexed by row gives dQ; (X,Y)=(C,Q) indexed by column gives dC.
//
// warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer   warps 2-5 / 6-9: softmax groups 0 / 1.
// Both kernels are bound by the MUFU (ex2) pipe, not the tensor pipe, at d = 64: a 128 x 256
// logits tile costs 512 tensor cycles but 2048 MUFU cycles (16 ex2 / clk / SM).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace tt {
namespace tc {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// named barrier private to one softmax group (ids 1 and 2), 128 threads
__device__ __forceinline__ void group_bar_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void softmax_bar_sync_all(int nthreads) { asm volatile("bar.sync 5, %0;" ::"r"(nthreads) : "memory"); }
// one MUFU op; inputs are finite or -inf, flush-to-zero is what we want for tiny probabilities
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + minimax polynomial, x <= ~0): the softmax loops
// are bound by the 16-per-clock MUFU pipe, so a fixed fraction of the exponentials is computed
// here instead (the FlashAttention-4 trick).  DEG 3: max rel. error 1.0e-4 (below the bf16
// rounding P gets anyway); DEG 4: 2.8e-6 (forward log-sum-exp).
template <int DEG>
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: low mantissa bits = round(x)
  const float f = x - (t - 12582912.f);           // in [-0.5, 0.5]
  float p;
  if (DEG == 3) p = fmaf(fmaf(fmaf(0.05500893f, f, 0.24221095f), f, 0.6932829f), f, 1.0f);
  else p = fmaf(fmaf(fmaf(fmaf(0.009582853f, f, 0.055906426f), f, 0.24024099f), f, 0.69312418f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // * 2^round(x)
}
// The same on two values at once with the packed fp32x2 FMA / ADD of sm_100 (FFMA2 / FADD2: one issue slot
// for two lanes' worth of work -- the softmax loops are short of issue slots as well as of MUFU throughput).
template <int DEG>
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(r, make_float2(-1.f, -1.f), x);
  float2 p;
  if (DEG == 3) {
    p = __ffma2_rn(make_float2(0.05500893f, 0.05500893f), f, make_float2(0.24221095f, 0.24221095f));
    p = __ffma2_rn(p, f, make_float2(0.6932829f, 0.6932829f));
  } else {
    p = __ffma2_rn(make_float2(0.009582853f, 0.009582853f), f, make_float2(0.055906426f, 0.055906426f));
    p = __ffma2_rn(p, f, make_float2(0.24024099f, 0.24024099f));
    p = __ffma2_rn(p, f, make_float2(0.69312418f, 0.69312418f));
  }
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
  return r;
}
// exp2 of 8 consecutive scaled logits x = v * scale + shift, N8 of them (in pairs) on the FMA pipe
template <int N8, int DEG>
__device__ __forceinline__ void ex2_8(const uint32_t* v, float2 sc, float2 sh, float2 (&e)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), sc, sh);
    const bool both = (N8 >= 2 && k == 3) || (N8 >= 4 && k == 1) || (N8 >= 6 && k == 2);
    const bool one = (N8 == 1 && k == 3) || (N8 == 3 && k == 1) || (N8 == 5 && k == 2);
    if (both) e[k] = ex2_poly2<DEG>(x);
    else if (one) e[k] = make_float2(ex2(x.x), ex2_poly<DEG>(x.y));
    else e[k] = make_float2(ex2(x.x), ex2(x.y));
  }
}
// which of every 8 consecutive elements go to the polynomial: N8 of 8
template <int N8>
__device__ __forceinline__ constexpr bool use_poly(int j) {
  return N8 >= 3 ? ((j & 7) == 2 || (j & 7) == 5 || (j & 7) == 7) : N8 == 2 ? ((j & 7) == 3 || (j & 7) == 7) : N8 == 1 ? (j & 7) == 7 : false;
}
#ifndef TT_POLY_FWD
#define TT_POLY_FWD 2
#endif
#ifndef TT_POLY_BWD
#define TT_POLY_BWD 3
#endif
#ifndef TT_MMA_SLEEP_NS
#define TT_MMA_SLEEP_NS 64
#endif
constexpr int kPolyFwd = TT_POLY_FWD, kPolyBwd = TT_POLY_BWD;

__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float a[10], b[4];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  b[0] = fmax3(a[0], a[1], a[2]);
  b[1] = fmax3(a[3], a[4], a[5]);
  b[2] = fmax3(a[6], a[7], a[8]);
  b[3] = fmax3(a[9], __uint_as_float(v[30]), __uint_as_float(v[31]));
  return fmaxf(fmax3(b[0], b[1], b[2]), b[3]);
}
// Producer / MMA-issuer waits: these warps share a scheduler with softmax warps, so back off
// instead of burning issue slots in a try_wait spin.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (TT_MMA_SLEEP_NS > 0) __nanosleep(TT_MMA_SLEEP_NS);
  }
  printf("tt_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}

// ------------------------------------------------------------------ forward
template <int KB>
struct FwdCfg {
#ifndef TT_FWD_GROUPS
#define TT_FWD_GROUPS 4
#endif
  static constexpr int NG = TT_FWD_GROUPS;              // softmax warpgroups; group g takes tiles t = g (mod NG)
  static constexpr int NT = (KB <= 2 && NG == 2) ? 256 : 128;   // C rows (logit columns) per tile
  static constexpr int NACC = 512 / NT;                 // TMEM accumulator stages (NACC x NT = 512 columns)
  static constexpr int THREADS = 64 + NG * 128;
  static constexpr int STAGES = KB == 1 ? 4 : (KB == 3 ? 3 : 2);
  static_assert(NACC % NG == 0 || NG % NACC == 0, "stage ownership");
  static constexpr int Q_BYTES = KB * 128 * 128;       // KB sub-tiles of [128 x 64] bf16
  static constexpr int C_BYTES = KB * NT * 128;        // KB sub-tiles of [NT x 64] bf16
  static constexpr int SMEM = Q_BYTES + STAGES * C_BYTES + 1024 + 256 + 4096 + 256;  // + merge scratch [NG-1][128][2] + path flags
};

// running (max, sum) update over one 32-column chunk; TAIL masks columns >= B
template <bool TAIL>
__device__ __forceinline__ void lse_chunk(const uint32_t (&v)[32], int col0, int B, float scale2, float& m, float& l) {
  float mx;
  if (!TAIL) {
    mx = max32(v) * scale2;
  } else {
    mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < B) mx = fmaxf(mx, __uint_as_float(v[j]) * scale2);
    if (mx == -INFINITY) return;  // chunk entirely past the last column
  }
  const float mn = fmaxf(m, mx);
  const float2 sc = make_float2(scale2, scale2), sh = make_float2(-mn, -mn);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    float2 e[4];
    ex2_8<kPolyFwd, 4>(&v[j], sc, sh, e);
    if (TAIL) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e[k].x = (col0 + j + 2 * k < B) ? e[k].x : 0.f;
        e[k].y = (col0 + j + 2 * k + 1 < B) ? e[k].y : 0.f;
      }
    }
    acc0 = __fadd2_rn(acc0, __fadd2_rn(e[0], e[1]));
    acc1 = __fadd2_rn(acc1, __fadd2_rn(e[2], e[3]));
  }
  const float a0 = acc0.x, a1 = acc0.y, a2 = acc1.x, a3 = acc1.y;
  l = l * ex2(m - mn) + ((a0 + a1) + (a2 + a3));
  m = mn;
}

// The same with a shift fixed for the whole row (an upper bound of its logits): no maximum, no rescaling, no
// dependency between chunks -- just sum += 2^(s*scale - shift).
#ifndef TT_POLY_FWD_FIXED
#define TT_POLY_FWD_FIXED 3
#endif
#ifndef TT_POLY_FWD_FIXED_DEG
#define TT_POLY_FWD_FIXED_DEG 4
#endif
template <bool TAIL>
__device__ __forceinline__ void lse_chunk_fixed(const uint32_t (&v)[32], int col0, int B, float scale2, float shift, float& l) {
  const float2 sc = make_float2(scale2, scale2), sh = make_float2(-shift, -shift);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    float2 e[4];
    ex2_8<TT_POLY_FWD_FIXED, TT_POLY_FWD_FIXED_DEG>(&v[j], sc, sh, e);
    if (TAIL) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e[k].x = (col0 + j + 2 * k < B) ? e[k].x : 0.f;
        e[k].y = (col0 + j + 2 * k + 1 < B) ? e[k].y : 0.f;
      }
    }
    acc0 = __fadd2_rn(acc0, __fadd2_rn(e[0], e[1]));
    acc1 = __fadd2_rn(acc1, __fadd2_rn(e[2], e[3]));
  }
  l += (acc0.x + acc0.y) + (acc1.x + acc1.y);
}

template <int KB>
__global__ void __launch_bounds__(FwdCfg<KB>::THREADS, 1)
tc_softmax_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, int B,
                      float scale2, const float* __restrict__ diag, float* __restrict__ lse,
                      float* __restrict__ partial_loss, float* __restrict__ ml_out,
                      const float* __restrict__ qn2, const unsigned int* __restrict__ cmax2) {
  // gridDim.y > 1: the logit columns are split between gridDim.y CTAs per row block (so that the
  // grid is ~7 waves of 148 instead of 3.46); each writes its (max, sum) pair to ml_out and
  // lse_merge_kernel finishes the job.
  using Cfg = FwdCfg<KB>;
  constexpr int NT = Cfg::NT, S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sC = smem + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + S * Cfg::C_BYTES);
  uint64_t* q_full = bars;
  uint64_t* c_full = bars + 1;            // [S]
  uint64_t* c_empty = c_full + S;         // [S]
  uint64_t* acc_full = c_empty + S;       // [NACC]
  uint64_t* acc_empty = acc_full + Cfg::NACC;   // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::NACC);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int Tall = (B + NT - 1) / NT;
  const int Tper = (Tall + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * Tper;                       // first tile of this CTA
  const int T = max(0, min(Tper, Tall - t0));             // tiles of this CTA (local index t)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmC);
    mbar_init(q_full, 1);
    for (int s = 0; s < S; ++s) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], 1); }
    for (int s = 0; s < Cfg::NACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], (Cfg::NG == 4 && NT == 128) ? 256 : 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sQ + kb * 128 * 128, &tmQ, q_full, kb * 64, m0);
      for (int t = 0; t < T; ++t) {
        const int s = t % S;
        mbar_wait_relaxed(&c_empty[s], ((t / S) & 1) ^ 1);
        mbar_expect_tx(&c_full[s], Cfg::C_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sC + s * Cfg::C_BYTES + kb * NT * 128, &tmC, &c_full[s], kb * 64, (t0 + t) * NT);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(128, NT);
    if (elect_one()) {
      mbar_wait_relaxed(q_full, 0);
      for (int t = 0; t < T; ++t) {
        const int s = t % S, as = t % Cfg::NACC;
        mbar_wait_relaxed(&acc_empty[as], ((t / Cfg::NACC) & 1) ^ 1);
        mbar_wait_relaxed(&c_full[s], (t / S) & 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = smem_desc_k_sw128(smem_u32(sQ + kb * 128 * 128));
          const uint64_t db = smem_desc_k_sw128(smem_u32(sC + s * Cfg::C_BYTES + kb * NT * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem_base + as * NT, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        tc_commit(&c_empty[s]);
        tc_commit(&acc_full[as]);
      }
    }
  } else {
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;                  // softmax group = TMEM stage it owns
    const int r_in = q * 32 + lane;
    const int row = m0 + r_in;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m = -INFINITY, l = 0.f;                   // running max / sum, log2 domain
    if (Cfg::NG == 4 && NT == 128) {
      // Groups work in pairs: pair g/2 takes the tiles t = g/2 (mod 2), group g%2 of the pair the 64-column half
      // g%2 of each.  A thread pulls its 64 columns into registers in one go and releases the TMEM stage at once,
      // so the MMA issuer refills it while the exponentials run (with four stages it stays two tiles ahead).
      const int h = g & 1;
      // Fixed-shift path: valid when, for EVERY row of the CTA, bound - S_ii <= 60 (log2 units): the diagonal term alone
      // then keeps the sum above 2^-60, far from underflow, and what the polynomial's clamp at 2^-125 adds is below
      // 2^-49 of it.  Decided per CTA (the four groups see the same 128 rows, so they agree); otherwise the
      // online-max path below.
      bool fixed = false;
      float shift = 0.f;
      if (qn2 != nullptr) {
        shift = row < B ? sqrtf(qn2[row] * __uint_as_float(*cmax2)) * fabsf(scale2) : 0.f;
        const bool ok = row >= B || (shift - diag[row] * kLog2e <= 60.f && shift < 1e30f);
        int* flags = reinterpret_cast<int*>(bars + 32) + 1024;       // [NG][4], past the merge scratch
        const int all_ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) flags[g * 4 + q] = all_ok;
        group_bar_sync(g);
        fixed = (flags[g * 4] & flags[g * 4 + 1] & flags[g * 4 + 2] & flags[g * 4 + 3]) != 0;
      }
      if (fixed) {
        m = shift;
        for (int t = g >> 1; t < T; t += 2) {
          const int as = t % Cfg::NACC;
          const uint32_t tcol = trow + as * NT + h * 64;
          mbar_wait(&acc_full[as], (t / Cfg::NACC) & 1);
          tc_fence_after();
          const int n0 = (t0 + t) * NT + h * 64;
          uint32_t va[32], vb[32];
          tmem_ld32(tcol, va);
          tmem_ld32(tcol + 32, vb);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
          if (n0 + 64 <= B) {
            lse_chunk_fixed<false>(va, n0, B, scale2, shift, l);
            lse_chunk_fixed<false>(vb, n0 + 32, B, scale2, shift, l);
          } else {
            lse_chunk_fixed<true>(va, n0, B, scale2, shift, l);
            lse_chunk_fixed<true>(vb, n0 + 32, B, scale2, shift, l);
          }
        }
        if (l == 0.f) m = -INFINITY;      // this thread saw no valid column (tail): neutral element of the merge
      } else
      for (int t = g >> 1; t < T; t += 2) {
        const int as = t % Cfg::NACC;
        const uint32_t tcol = trow + as * NT + h * 64;
        mbar_wait(&acc_full[as], (t / Cfg::NACC) & 1);
        tc_fence_after();
        const int n0 = (t0 + t) * NT + h * 64;
        const bool tail = n0 + 64 > B;
        uint32_t va[32], vb[32];
        tmem_ld32(tcol, va);
        tmem_ld32(tcol + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[as]);
        if (!tail) {
          lse_chunk<false>(va, n0, B, scale2, m, l);
          lse_chunk<false>(vb, n0 + 32, B, scale2, m, l);
        } else {
          lse_chunk<true>(va, n0, B, scale2, m, l);
          lse_chunk<true>(vb, n0 + 32, B, scale2, m, l);
        }
      }
    } else
    for (int t = g; t < T; t += Cfg::NG) {
      const int as = t % Cfg::NACC;
      const uint32_t tcol = trow + as * NT;
      mbar_wait(&acc_full[as], (t / Cfg::NACC) & 1);
      tc_fence_after();
      const int n0 = (t0 + t) * NT;
      const bool tail = n0 + NT > B;
      uint32_t va[32], vb[32];
      tmem_ld32(tcol, va);
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 64) {
        tmem_ld_wait();
        tmem_ld32(tcol + c0 + 32, vb);
        if (!tail) lse_chunk<false>(va, n0 + c0, B, scale2, m, l); else lse_chunk<true>(va, n0 + c0, B, scale2, m, l);
        tmem_ld_wait();
        if (c0 + 64 < NT) tmem_ld32(tcol + c0 + 64, va);
        if (!tail) lse_chunk<false>(vb, n0 + c0 + 32, B, scale2, m, l); else lse_chunk<true>(vb, n0 + c0 + 32, B, scale2, m, l);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    // merge the two groups' partial (max, sum) of every row, then this CTA's loss partial
    float* mrg = reinterpret_cast<float*>(bars + 32);  // [128][2], own scratch (the C ring may still be in use)
    if (g > 0) { mrg[((g - 1) * 128 + r_in) * 2] = m; mrg[((g - 1) * 128 + r_in) * 2 + 1] = l; }
    softmax_bar_sync_all(Cfg::NG * 128);
    if (g == 0) {
      float mn = m, lt = l;
#pragma unroll
      for (int og = 1; og < Cfg::NG; ++og) {
        const float m1 = mrg[((og - 1) * 128 + r_in) * 2], l1 = mrg[((og - 1) * 128 + r_in) * 2 + 1];
        const float m2 = fmaxf(mn, m1);
        lt = ((mn == -INFINITY) ? 0.f : lt * ex2(mn - m2)) + ((m1 == -INFINITY) ? 0.f : l1 * ex2(m1 - m2));
        mn = m2;
      }
      float contrib = 0.f;
      if (row < B) {
        if (gridDim.y == 1) {
          const float L = (mn + log2f(lt)) * kLn2;
          lse[row] = L;
          contrib = L - diag[row];
        } else {
          ml_out[((int64_t)blockIdx.y * B + row) * 2] = mn;
          ml_out[((int64_t)blockIdx.y * B + row) * 2 + 1] = lt;
        }
      }
      contrib = warp_sum(contrib);
      if (lane == 0) red[q] = contrib;
    }
    softmax_bar_sync_all(Cfg::NG * 128);
    if (warp == 2 && lane == 0 && gridDim.y == 1) partial_loss[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ backward
template <int KB>
struct BwdCfg {
  static constexpr int D = KB * 64;
#ifdef TT_BWD_NT64
  static constexpr int NT = 64;
  static constexpr int NSP = 4;
#else
  static constexpr int NT = KB <= 2 ? 128 : 64;          // Y rows per tile = S/P columns per TMEM stage
  static constexpr int NSP = KB <= 2 ? 3 : 4;            // TMEM S/P stages: columns [NT*a, NT*a + NT)
#endif
  static constexpr int NKB2 = NT / 64;                   // K blocks of the second GEMM
#ifndef TT_BWD_GROUPS
#define TT_BWD_GROUPS 3
#endif
  static constexpr int NG = KB <= 2 ? TT_BWD_GROUPS : 2;  // softmax warpgroups (tiles t = g mod NG)
  static constexpr int THREADS = 64 + NG * 128;
  static constexpr int STAGES = KB == 1 ? 4 : 2;
  static constexpr int LA0 = NSP - 1;                    // S tiles issued ahead of the P they wait for:
  static constexpr int LOOKAHEAD = STAGES - 1 < LA0 ? STAGES - 1 : LA0;   // bounded by TMEM and smem stages
  static constexpr int X_BYTES = KB * 128 * 128;
  static constexpr int Y_BYTES = KB * NT * 128;          // KB sub-tiles [NT x 64]  (GEMM1 B operand)
  static constexpr int YT_BYTES = NKB2 * D * 128;        // NKB2 sub-tiles [D x 64] (GEMM2 B operand)
  static constexpr int STAGE_BYTES = Y_BYTES + YT_BYTES;
  static constexpr int SMEM = X_BYTES + STAGES * STAGE_BYTES + 1024 + 256 + 4 * 128 * 4;
  static constexpr int O_COL = NSP * NT;                 // O accumulator columns [O_COL, O_COL + D) (<= 512)
  static_assert(O_COL + D <= 512, "TMEM budget");
};

// one 32-column chunk of S -> P (bf16 pairs), written back over S in TMEM
template <bool ROW, bool TAIL>
__device__ __forceinline__ void p_chunk(const uint32_t (&v)[32], uint32_t tdst, const float* lse_cols, float lrow,
                                        float scale2, int col0, int B) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float l0 = lrow, l1 = lrow, l2 = lrow, l3 = lrow;
    if (!ROW) {
      const float4 lv = *reinterpret_cast<const float4*>(lse_cols + j);
      l0 = lv.x; l1 = lv.y; l2 = lv.z; l3 = lv.w;
    }
    const float x0 = fmaf(__uint_as_float(v[j]), scale2, -l0), x1 = fmaf(__uint_as_float(v[j + 1]), scale2, -l1);
    const float x2 = fmaf(__uint_as_float(v[j + 2]), scale2, -l2), x3 = fmaf(__uint_as_float(v[j + 3]), scale2, -l3);
    float p0 = use_poly<kPolyBwd>(j) ? ex2_poly<3>(x0) : ex2(x0);
    float p1 = use_poly<kPolyBwd>(j + 1) ? ex2_poly<3>(x1) : ex2(x1);
    float p2 = use_poly<kPolyBwd>(j + 2) ? ex2_poly<3>(x2) : ex2(x2);
    float p3 = use_poly<kPolyBwd>(j + 3) ? ex2_poly<3>(x3) : ex2(x3);
    if (TAIL) {
      p0 = (col0 + j < B) ? p0 : 0.f;
      p1 = (col0 + j + 1 < B) ? p1 : 0.f;
      p2 = (col0 + j + 2 < B) ? p2 : 0.f;
      p3 = (col0 + j + 3 < B) ? p3 : 0.f;
    }
    pk[j >> 1] = pack_bf16(p0, p1);
    pk[(j >> 1) + 1] = pack_bf16(p2, p3);
  }
  tmem_st16(tdst, pk);
}

template <int KB, bool ROW>
__global__ void __launch_bounds__(BwdCfg<KB>::THREADS, 1)
tc_softmax_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                      const __grid_constant__ CUtensorMap tmYt, int B, int d, float scale2,
                      const float* __restrict__ lse, const float* __restrict__ yf, int64_t ld_yf,
                      const float* __restrict__ relu_mask, int64_t ld_mask, float out_scale,
                      float* __restrict__ out, int64_t ld_out) {
  using Cfg = BwdCfg<KB>;
  constexpr int NT = Cfg::NT, S = Cfg::STAGES, D = Cfg::D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + Cfg::X_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + S * Cfg::STAGE_BYTES);
  uint64_t* x_full = bars;
  uint64_t* y_full = bars + 1;           // [S]
  uint64_t* y_empty = y_full + S;        // [S]
  uint64_t* s_full = y_empty + S;        // [NSP]
  uint64_t* p_full = s_full + Cfg::NSP;  // [NSP]
  uint64_t* o_full = p_full + Cfg::NSP;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* lse_tile = reinterpret_cast<float*>(bars + 32);      // [NG][128], 16-byte aligned (float4 reads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int Tall = (B + NT - 1) / NT;
  const int Tper = (Tall + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * Tper;                       // column split: see the forward kernel
  const int T = max(0, min(Tper, Tall - t0));             // the host guarantees T >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmYt);
    mbar_init(x_full, 1);
    for (int s = 0; s < S; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < Cfg::NSP; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(x_full, Cfg::X_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sX + kb * 128 * 128, &tmX, x_full, kb * 64, m0);
      for (int t = 0; t < T; ++t) {
        const int s = t % S;
        uint8_t* st = sY + s * Cfg::STAGE_BYTES;
        mbar_wait_relaxed(&y_empty[s], ((t / S) & 1) ^ 1);
        mbar_expect_tx(&y_full[s], Cfg::STAGE_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(st + kb * NT * 128, &tmY, &y_full[s], kb * 64, (t0 + t) * NT);
        for (int k2 = 0; k2 < Cfg::NKB2; ++k2)
          tma_load_2d(st + Cfg::Y_BYTES + k2 * D * 128, &tmYt, &y_full[s], (t0 + t) * NT + k2 * 64, 0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc1 = idesc_bf16_f32(128, NT);
    constexpr uint32_t idesc2 = idesc_bf16_f32(128, D);
    if (elect_one()) {
      auto issue_gemm1 = [&](int t) {
        const int s = t % S, as = t % Cfg::NSP;
        mbar_wait_relaxed(&y_full[s], (t / S) & 1);
        tc_fence_after();
        uint8_t* st = sY + s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = smem_desc_k_sw128(smem_u32(sX + kb * 128 * 128));
          const uint64_t db = smem_desc_k_sw128(smem_u32(st + kb * NT * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem_base + as * NT, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
        }
        tc_commit(&s_full[as]);
      };
      mbar_wait_relaxed(x_full, 0);
      for (int i = 0; i < Cfg::LOOKAHEAD && i < T; ++i) issue_gemm1(i);
      for (int t = 0; t < T; ++t) {
        const int s = t % S, as = t % Cfg::NSP;
        // S tiles run LOOKAHEAD ahead of the P tile awaited below, so a softmax group always finds
        // its next S ready.  TMEM stage (t+LOOKAHEAD)%NSP is free: the last P that lived there was
        // consumed by an MMA2 issued earlier, and tcgen05.mma executes in issue order.
        if (t + Cfg::LOOKAHEAD < T) issue_gemm1(t + Cfg::LOOKAHEAD);
        mbar_wait_relaxed(&p_full[as], (t / Cfg::NSP) & 1);
        tc_fence_after();
        uint8_t* yt = sY + s * Cfg::STAGE_BYTES + Cfg::Y_BYTES;
#pragma unroll
        for (int k2 = 0; k2 < Cfg::NKB2; ++k2) {
          const uint64_t db = smem_desc_k_sw128(smem_u32(yt + k2 * D * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ts(tmem_base + Cfg::O_COL, tmem_base + as * NT + (k2 * 4 + k) * 8, db + 2 * k, idesc2,
                   (t | k2 | k) != 0);
        }
        tc_commit(&y_empty[s]);
        if (t == T - 1) tc_commit(o_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;                   // softmax group = TMEM S/P stage it owns
    const int r_in = q * 32 + lane;
    const int row = m0 + r_in;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float lrow = (ROW && row < B) ? lse[row] * kLog2e : 0.f;
    float* lse_g = lse_tile + g * 128;
    for (int t = g; t < T; t += Cfg::NG) {
      const int n0 = (t0 + t) * NT;
      const int as = t % Cfg::NSP;
      const uint32_t tsp = trow + as * NT;
      if (!ROW) {
        if (r_in < NT) lse_g[r_in] = (n0 + r_in < B) ? lse[n0 + r_in] * kLog2e : 0.f;
        group_bar_sync(g);
      }
      mbar_wait(&s_full[as], (t / Cfg::NSP) & 1);
      tc_fence_after();
      const bool tail = n0 + NT > B;
      // P (bf16) overwrites S (fp32) in place.  P chunk c lands on S columns [16c, 16c+16), which
      // belong to S chunk floor(c/2) <= c: already in registers when the store is issued.
      uint32_t va[32], vb[32];
      tmem_ld32(tsp, va);
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 64) {
        tmem_ld_wait();
        tmem_ld32(tsp + c0 + 32, vb);
        if (!tail) p_chunk<ROW, false>(va, tsp + (c0 >> 1), lse_g + c0, lrow, scale2, n0 + c0, B);
        else p_chunk<ROW, true>(va, tsp + (c0 >> 1), lse_g + c0, lrow, scale2, n0 + c0, B);
        tmem_ld_wait();
        if (c0 + 64 < NT) tmem_ld32(tsp + c0 + 64, va);
        if (!tail) p_chunk<ROW, false>(vb, tsp + ((c0 + 32) >> 1), lse_g + c0 + 32, lrow, scale2, n0 + c0 + 32, B);
        else p_chunk<ROW, true>(vb, tsp + ((c0 + 32) >> 1), lse_g + c0 + 32, lrow, scale2, n0 + c0 + 32, B);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[as]);
      if (!ROW) group_bar_sync(g);   // lse_g is rewritten at the top of the next iteration
    }
    // ---- O -> out = out_scale * (O - Y_r), optionally gated by relu_mask > 0 (groups alternate chunks)
    mbar_wait(o_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = g * 32; c0 < D; c0 += Cfg::NG * 32) {
      uint32_t v[32];
      tmem_ld32(trow + Cfg::O_COL + c0, v);
      tmem_ld_wait();
      if (row < B) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col < d) {
            float o = __uint_as_float(v[j]);
            if (blockIdx.y == 0) o -= yf[(int64_t)row * ld_yf + col];
            o *= out_scale;
            if (relu_mask != nullptr && !(relu_mask[(int64_t)row * ld_mask + col] > 0.f)) o = 0.f;
            if (gridDim.y == 1) out[(int64_t)row * ld_out + col] = o;
            else atomicAdd(&out[(int64_t)row * ld_out + col], o);   // two addends: order-independent
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ fused backward (d <= 64)
// ONE pass over the [B, B] probabilities gives both gradients: every P tile is computed once
// (one exponential per logit instead of two) and feeds two tensor-core products,
//     accX[r, :] += sum_c P(r, c) Y[c, :]        accY[c, :] += sum_r P(r, c) X[r, :].
// A CTA owns RT = 2 row tiles of X (256 rows, accX resident in TMEM) and a chunk of the column tiles;
// accY of each column tile is complete within the CTA after its two row tiles and leaves through a
// TMA reduce-add (fp32 add in L2), accX leaves the same way once per CTA.  P goes TMEM -> registers ->
// bf16 in shared memory (128-byte swizzle) and is read by the tensor core twice: K-major as the A
// operand of P.Y and MN-major ("transposed") as the A operand of P^T.X; the X and Y tiles TMA loaded
// for S = X Y^T double as the MN-major B operands, so no transposed copies of X / Y exist.
//   warpgroup 0: warp 0 TMA, warp 1 MMA issuer | warpgroups 1-4: softmax groups (row tile i, 64-column half h) |
//   warpgroup 5: flush.  A softmax thread pulls its 64 columns of the S row into registers at once, which frees
//   the TMEM stage for the next S a whole tile early; setmaxnreg moves registers from warpgroup 0 to the others.
//   TMEM: S stage of row tile i [128i, 128i+128) | accX_i [256+64i, +64) | accY double buffer [384+64b, +64)
// The sums are finished by softmax_bwd_finalize_kernel: out = gate * out_scale * (acc - other side's row).
// Measured dead ends of round 2 (tools/mma_rate.cu, profiles/r02_mma_rate.txt, profiles/r02_softmax_bwd_ts_form_attempts.txt):
// a tcgen05.mma with M = 128, K = 16 never takes less than ~48 (TS) / ~53 (SS) clocks whatever N is, so the N = 64
// gradient products run at 60-66 % of the tensor rate in either form and smaller S tiles only add instructions.  A TS-form
// variant (X resident in TMEM, P written back in place as the A operand of P.Y: 152 KB instead of 200 KB of shared-memory
// traffic per tile) was built, passed every parity test and was SLOWER (1.90 ms with one 64-column stage per softmax group,
// 2.39 / 2.01 ms with two 32-column stages; this kernel: 1.58 ms): in-place P keeps the S stage busy until P.Y has run, the
// look-ahead below is lost, and more, narrower MMAs cost 48 clocks each.
struct FusedCfg {
  static constexpr int RT = 2, NT = 128, YS = 3, PB = 4;     // PB: P buffers, two private to each softmax group
  static constexpr int CH = 2;                           // softmax groups per row tile (each takes NT/CH columns of the row)
  static constexpr int THREADS = 128 * (2 + RT * CH);    // warpgroups: {TMA, MMA, -, -} | RT*CH softmax groups | flush
  static constexpr int REGS_CTRL = 40, REGS_SOFTMAX = 96, REGS_FLUSH = 56;   // setmaxnreg: 40 + 4*96 + 56 = 6 * 80 (launch value)
  static constexpr int REGS_LAUNCH = (65536 / THREADS) / 8 * 8;   // what every warp owns when the kernel starts
  static_assert(REGS_CTRL + RT * CH * REGS_SOFTMAX + REGS_FLUSH <= REGS_LAUNCH * (THREADS / 128),
                "setmaxnreg.inc only draws from what setmaxnreg.dec released inside the CTA");
  static constexpr int TILE = 128 * 128;                 // one [128 x 64] bf16 tile, bytes
  static constexpr int X_BYTES = RT * TILE;
  static constexpr int P_BYTES = 2 * TILE;               // two panels [128 r x 64 c]
  static constexpr int STG_BYTES = TILE;                 // one box [128 rows x 32 fp32]: accumulators leave in two halves
  static constexpr int SMEM = X_BYTES + YS * TILE + PB * P_BYTES + STG_BYTES + 1024 + 512;
  static constexpr int ACCX_COL = 256, ACCY_COL = 384;
};

// one 32-column chunk of S -> P -> four 16-byte pieces of the swizzled bf16 row in shared memory
__device__ __forceinline__ void p_chunk_smem(const uint32_t (&v)[32], uint32_t row_addr, int slot0, int r7, float lrow,
                                             float scale2) {
  const float2 sc = make_float2(scale2, scale2), sh = make_float2(-lrow, -lrow);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float2 e[4];
    ex2_8<kPolyBwd, 3>(&v[m * 8], sc, sh, e);
    st_shared_v4(row_addr + (((slot0 + m) ^ r7) << 4), pack_bf16(e[0].x, e[0].y), pack_bf16(e[1].x, e[1].y),
                 pack_bf16(e[2].x, e[2].y), pack_bf16(e[3].x, e[3].y));
  }
}

__device__ __forceinline__ void flush_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// 64 fp32 accumulator columns of this thread's TMEM lane -> swizzled staging box -> TMA reduce-add, 32 columns at a time
__device__ __forceinline__ void flush_acc(uint32_t taddr, uint32_t stg, int r_in, bool leader, const CUtensorMap* tm,
                                          int row0, int d, uint64_t* release_bar) {
  const uint32_t rowa = stg + r_in * 128;
  const int r7 = r_in & 7;
  const int halves = d > 32 ? 2 : 1;
#pragma unroll 1
  for (int h = 0; h < halves; ++h) {
    uint32_t a[32];
    tmem_ld32(taddr + 32 * h, a);
    tmem_ld_wait();
    if (h == halves - 1 && release_bar != nullptr) {
      tc_fence_before();
      mbar_arrive(release_bar);
    }
    if (leader) bulk_wait_group_read0();   // the previous reduce has finished reading the staging box
    flush_bar_sync();
#pragma unroll
    for (int m = 0; m < 8; ++m) st_shared_v4(rowa + ((m ^ r7) << 4), a[4 * m], a[4 * m + 1], a[4 * m + 2], a[4 * m + 3]);
    fence_proxy_async_smem();
    flush_bar_sync();
#ifndef TT_FUSED_NO_REDUCE   // diagnostic builds only: time the kernel without the L2 reduce traffic
    if (leader) {
      tma_reduce_add_2d(tm, reinterpret_cast<const void*>(__cvta_shared_to_generic(stg)), 32 * h, row0);
      bulk_commit_group();
    }
#endif
  }
}

__global__ void __launch_bounds__(FusedCfg::THREADS, 1)
tc_softmax_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                            const __grid_constant__ CUtensorMap tmAccX, const __grid_constant__ CUtensorMap tmAccY,
                            int B, int d, float scale2, const float* __restrict__ lse) {
  using Cfg = FusedCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + Cfg::X_BYTES;
  uint8_t* sP = sY + Cfg::YS * Cfg::TILE;
  uint8_t* sStg = sP + Cfg::PB * Cfg::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + Cfg::STG_BYTES);
  uint64_t* x_full = bars;
  uint64_t* y_full = x_full + 1;            // [YS]
  uint64_t* y_empty = y_full + Cfg::YS;     // [YS]
  uint64_t* s_full = y_empty + Cfg::YS;     // [RT]
  uint64_t* s_empty = s_full + Cfg::RT;     // [RT]
  uint64_t* p_full = s_empty + Cfg::RT;     // [PB]
  uint64_t* p_empty = p_full + Cfg::PB;     // [PB]
  uint64_t* ay_full = p_empty + Cfg::PB;    // [2]
  uint64_t* ay_empty = ay_full + 2;         // [2]
  uint64_t* ax_full = ay_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ax_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * (Cfg::RT * 128);
  const int Tall = (B + 127) / 128;
  const int Tper = (Tall + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * Tper;
  const int T = max(0, min(Tper, Tall - t0));       // the host guarantees T >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmAccX);
    prefetch_tmap(&tmAccY);
    mbar_init(x_full, 1);
    for (int s = 0; s < Cfg::YS; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < Cfg::RT; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128 * Cfg::CH); }
    for (int s = 0; s < Cfg::PB; ++s) { mbar_init(&p_full[s], 128 * Cfg::CH); mbar_init(&p_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ay_full[s], 1); mbar_init(&ay_empty[s], 128); }
    mbar_init(ax_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int wg = warp >> 2;
  if (wg == 0) {
    setmaxnreg_dec<Cfg::REGS_CTRL>();
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(x_full, Cfg::X_BYTES);
        for (int i = 0; i < Cfg::RT; ++i) tma_load_2d(sX + i * Cfg::TILE, &tmX, x_full, 0, m0 + i * 128);
        for (int t = 0; t < T; ++t) {
          const int s = t % Cfg::YS;
          mbar_wait_relaxed(&y_empty[s], ((t / Cfg::YS) & 1) ^ 1);
          mbar_expect_tx(&y_full[s], Cfg::TILE);
          tma_load_2d(sY + s * Cfg::TILE, &tmY, &y_full[s], 0, (t0 + t) * 128);
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idescS = idesc_bf16_f32(128, 128);
      constexpr uint32_t idescX = idesc_bf16_f32(128, 64) | kIdescBMnMajor;                    // P . Y
      constexpr uint32_t idescY = idesc_bf16_f32(128, 64) | kIdescAMnMajor | kIdescBMnMajor;   // P^T . X
      if (elect_one()) {
        const uint32_t aX = smem_u32(sX), aY = smem_u32(sY), aP = smem_u32(sP);
        auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {   // non-blocking phase test
          uint32_t ok;
          asm volatile(
              "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
              : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
          return ok != 0;
        };
        auto issue_s = [&](int i, int j) {
          const int ys = j % Cfg::YS;
          tc_fence_after();
          const uint64_t da = smem_desc_k_sw128(aX + i * Cfg::TILE);
          const uint64_t db = smem_desc_k_sw128(aY + ys * Cfg::TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem_base + i * 128, da + 2 * k, db + 2 * k, idescS, k != 0);
          tc_commit(&s_full[i]);
        };
        // products of P(i, j); `first` / `last`: first / second of the two row tiles to reach column tile j
        auto consume = [&](int i, int j, bool first, bool last) {
          const int n = 2 * j + i, pb = n % Cfg::PB, b = j & 1, ys = j % Cfg::YS;
          tc_fence_after();
          const uint32_t P = aP + pb * Cfg::P_BYTES, Y = aY + ys * Cfg::TILE, X = aX + i * Cfg::TILE;
#ifndef TT_FUSED_SKIP_DX
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // accX_i += P[r, 16kk..] . Y[16kk.., :]
            const uint64_t da = smem_desc_k_sw128(P + (kk >> 2) * Cfg::TILE) + 2 * (kk & 3);
            const uint64_t db = smem_desc_mn_sw128(Y + kk * 2048, 1024, 1024);
            mma_ss(tmem_base + Cfg::ACCX_COL + 64 * i, da, db, idescX, (j | kk) != 0);
          }
#endif
#ifndef TT_FUSED_SKIP_DY
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // accY_j += P[16kk.., c]^T . X[16kk.., :]
            const uint64_t da = smem_desc_mn_sw128(P + kk * 2048, Cfg::TILE, 1024);
            const uint64_t db = smem_desc_mn_sw128(X + kk * 2048, 1024, 1024);
            mma_ss(tmem_base + Cfg::ACCY_COL + 64 * b, da, db, idescY, (first ? kk : 1) != 0);
          }
#endif
          tc_commit(&p_empty[pb]);
          if (last) {
            tc_commit(&ay_full[b]);
            tc_commit(&y_empty[ys]);
          }
        };
        mbar_wait_relaxed(x_full, 0);
        // Event driven: issue whatever is ready, S tiles first (a softmax group frees its S stage at the
        // start of a tile, so S(i, j+1) can be a whole tile early); tcgen05.mma executes in issue order.
        int js[2] = {0, 0}, jc[2] = {0, 0};
        while (jc[0] < T || jc[1] < T) {
          bool progress = false;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int j = js[i];
            if (j < T && ready(&y_full[j % Cfg::YS], (j / Cfg::YS) & 1) && ready(&s_empty[i], (j & 1) ^ 1)) {
              issue_s(i, j);
              js[i] = j + 1;
              progress = true;
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int j = jc[i];
            if (j >= T) continue;
            const int n = 2 * j + i;
            const bool first = jc[1 - i] <= j;         // the other row tile has not reached column tile j yet
            if (!ready(&p_full[n % Cfg::PB], (n / Cfg::PB) & 1)) continue;
            if (first && !ready(&ay_empty[j & 1], ((j >> 1) & 1) ^ 1)) continue;   // flush group still reads accY(j-2)
            consume(i, j, first, !first);
            jc[i] = j + 1;
            progress = true;
          }
          if (!progress) __nanosleep(32);
        }
        tc_commit(ax_full);
      }
    }
  } else if (wg <= Cfg::RT * Cfg::CH) {
    setmaxnreg_inc<Cfg::REGS_SOFTMAX>();
    const int q = warp & 3;
    const int i = (wg - 1) / Cfg::CH;                // row tile of this softmax group
    const int h = (wg - 1) % Cfg::CH;                // its 64-column half of every S tile = panel h of P
    const int r_in = q * 32 + lane;
    const int row = m0 + i * 128 + r_in;
    const uint32_t ts = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + i * 128 + h * 64;
    const float lrow = row < B ? lse[row] * kLog2e : 0.f;   // rows >= B: X row is zero, contributes nothing
    const int r7 = r_in & 7;
    const uint32_t prow0 = smem_u32(sP) + h * Cfg::TILE + r_in * 128;
    for (int j = 0; j < T; ++j) {
      const int n = 2 * j + i, pb = n % Cfg::PB;
      const uint32_t prow = prow0 + pb * Cfg::P_BYTES;
      uint32_t va[32], vb[32];
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      tmem_ld32(ts, va);
      tmem_ld32(ts + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[i]);                            // the row is in registers: S stage free for tile j+1
      mbar_wait(&p_empty[pb], ((n / Cfg::PB) & 1) ^ 1);    // the products of the tile that used this buffer are done
      p_chunk_smem(va, prow, 0, r7, lrow, scale2);
      p_chunk_smem(vb, prow, 4, r7, lrow, scale2);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[pb]);
    }
  } else {
    setmaxnreg_dec<Cfg::REGS_FLUSH>();
    const int q = warp & 3;
    const int r_in = q * 32 + lane;
    const bool leader = threadIdx.x == (1 + Cfg::RT * Cfg::CH) * 128;
    const uint32_t tl = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t stg = smem_u32(sStg);
    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      mbar_wait_relaxed(&ay_full[b], (j >> 1) & 1);
      tc_fence_after();
      flush_acc(tl + Cfg::ACCY_COL + 64 * b, stg, r_in, leader, &tmAccY, (t0 + j) * 128, d, &ay_empty[b]);
    }
    mbar_wait_relaxed(ax_full, 0);
    tc_fence_after();
    for (int i = 0; i < Cfg::RT; ++i)
      if (m0 + i * 128 < B) flush_acc(tl + Cfg::ACCX_COL + 64 * i, stg, r_in, leader, &tmAccX, m0 + i * 128, d, nullptr);
    if (leader) bulk_wait_group0();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// out = gate * out_scale * (acc - other), in place over the accumulated sums (4 columns per thread)
template <int VEC>
__global__ void __launch_bounds__(256)
softmax_bwd_finalize_kernel(float* __restrict__ out, int64_t ld_out, const float* __restrict__ other, int64_t ld_other,
                            const float* __restrict__ mask, int64_t ld_mask, float out_scale, const float* __restrict__ scale_dev,
                            int B, int d) {
  if (scale_dev != nullptr) out_scale *= *scale_dev;      // the incoming dLoss (a device scalar): no separate multiply kernel
  const int per_row = (d + VEC - 1) / VEC;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)B * per_row) return;
  const int row = (int)(idx / per_row), c0 = (int)(idx % per_row) * VEC;
  if (VEC == 4) {
    float4 a = *reinterpret_cast<const float4*>(out + (int64_t)row * ld_out + c0);
    const float4 o = *reinterpret_cast<const float4*>(other + (int64_t)row * ld_other + c0);
    a.x = out_scale * (a.x - o.x); a.y = out_scale * (a.y - o.y); a.z = out_scale * (a.z - o.z); a.w = out_scale * (a.w - o.w);
    if (mask != nullptr) {
      const float4 g = *reinterpret_cast<const float4*>(mask + (int64_t)row * ld_mask + c0);
      a.x = g.x > 0.f ? a.x : 0.f; a.y = g.y > 0.f ? a.y : 0.f; a.z = g.z > 0.f ? a.z : 0.f; a.w = g.w > 0.f ? a.w : 0.f;
    }
    *reinterpret_cast<float4*>(out + (int64_t)row * ld_out + c0) = a;
  } else {
    float a = out_scale * (out[(int64_t)row * ld_out + c0] - other[(int64_t)row * ld_other + c0]);
    if (mask != nullptr && !(mask[(int64_t)row * ld_mask + c0] > 0.f)) a = 0.f;
    out[(int64_t)row * ld_out + c0] = a;
  }
}

// diag[b] = inv_t * sum_k bf16(q[b,k]) * bf16(c[b,k])   (what the tensor core computes for S_bb)
__global__ void __launch_bounds__(256)
// Also |q_b|^2 per row and max_b |c_b|^2 (atomicMax on the bit pattern of a non-negative float; zeroed by the caller):
// by Cauchy-Schwarz |S_bj| <= |q_b| max|c|, a row-wise upper bound the forward uses as a FIXED softmax shift.
rowdot_bf16_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ c, int64_t ldc,
                   int B, int d, float inv_t, float* __restrict__ diag, float* __restrict__ qn2, unsigned int* __restrict__ cmax2) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float s = 0.f, sq = 0.f, scn = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float qv = __bfloat162float(q[(int64_t)b * ldq + k]), cv = __bfloat162float(c[(int64_t)b * ldc + k]);
    s = fmaf(qv, cv, s);
    sq = fmaf(qv, qv, sq);
    scn = fmaf(cv, cv, scn);
  }
  s = warp_sum(s);
  sq = warp_sum(sq);
  scn = warp_sum(scn);
  if (lane == 0) {
    diag[b] = s * inv_t;
    qn2[b] = sq;
    // running maximum: 65 536 atomics on one address cost 35 us; a (possibly stale) read first leaves O(log B) of them
    const unsigned int bits = __float_as_uint(scn);
    if (bits > *reinterpret_cast<volatile unsigned int*>(cmax2)) atomicMax(cmax2, bits);
  }
}

__global__ void __launch_bounds__(1024)
loss_final_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(red[threadIdx.x]);
    if (threadIdx.x == 0) *out = t * scale;
  }
}

// column-split forward: combine the per-split (max, sum) pairs of a row; block loss partials
__global__ void __launch_bounds__(256)
lse_merge_kernel(const float* __restrict__ ml, int splits, int B, const float* __restrict__ diag,
                 float* __restrict__ lse, float* __restrict__ partial_loss) {
  __shared__ float red[8];
  const int row = blockIdx.x * 256 + threadIdx.x;
  float contrib = 0.f;
  if (row < B) {
    float m = -INFINITY, l = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float mi = ml[((int64_t)sp * B + row) * 2], li = ml[((int64_t)sp * B + row) * 2 + 1];
      const float mn = fmaxf(m, mi);
      l = ((m == -INFINITY) ? 0.f : l * exp2f(m - mn)) + ((mi == -INFINITY) ? 0.f : li * exp2f(mi - mn));
      m = mn;
    }
    const float L = (m + log2f(l)) * kLn2;
    lse[row] = L;
    contrib = L - diag[row];
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partial_loss[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------ wide embeddings (64 < d <= 256): A operand resident in TMEM
// At d = 256 a [128 x 256] bf16 operand tile is 64 KB.  The SS-form kernels above re-read the CTA's own Q / X tile from
// shared memory for every column tile (A = 4 KB per tcgen05.mma) next to the B reads.  Here the CTA's rows are stored ONCE
// into TMEM as bf16 (D/2 columns per 128 rows) and every S product is a TS-form MMA (A from TMEM), so shared memory carries
// only the streamed operand and the 64 KB the X tile occupied become two more pipeline stages:
// (The forward keeps the SS form: its N = 128 products read 8 KB per 64 tensor clocks, exactly what the pipe delivers, and it
// measures 1330-1370 TFLOP/s at d = 256 = 95 % of the sustained bf16 peak.  The SS-form backward measured 51 %: its N = 64
// S products are operand-bound and, worse, 64 KB stages left room for only two of them, so every tile waited out a TMA
// round trip.)
//   backward: one row tile per CTA; TMEM = O [0, D) | X [D, 3D/2) | S/P stages.  The streamed Y tile [NT x D] (K-major B of
//             S = X Y^T) is ALSO the B operand of O += P Y through an MN-major descriptor, so neither a transposed copy of
//             Y in HBM nor a second TMA load exists: 32 KB write + 64 KB reads per 1024 tensor clocks (d = 256, NT = 64).
// The two softmax groups split every S tile by columns (group h takes columns [h NT/2, (h+1) NT/2)) so that P is back in
// TMEM half a tile time after S lands; P is written in place over the group's OWN S columns.
template <int KB>
struct WideBwdCfg {
  static constexpr int D = KB * 64;
  static constexpr int NT = KB <= 2 ? 128 : 64;            // Y rows per tile = S columns per TMEM stage
  static constexpr int NSP = KB == 3 ? 3 : 2;              // S/P stages
  static constexpr int O_COL = 0, X_COL = D;
  static constexpr int S_COL = 512 - NSP * NT;
  static_assert(X_COL + D / 2 <= S_COL, "TMEM budget");
  static constexpr int THREADS = 64 + 2 * 128;
  static constexpr int Y_BYTES = KB * NT * 128;            // KB sub-tiles [NT x 64] bf16, 128-byte swizzle
  static constexpr int STAGES = 4;
  static constexpr int LOOKAHEAD = NSP - 1;
  static constexpr int SMEM = STAGES * Y_BYTES + 1024 + 256 + 2 * 128 * 4 + 256;
};

// this thread's row of a row-major bf16 matrix -> TMEM columns [tcol, tcol + ncols/2) of its lane, ncols a multiple of 32;
// elements at or past `d` (and whole rows when !valid) are zero
__device__ __forceinline__ void row_to_tmem(const __nv_bfloat16* __restrict__ row, bool valid, int d, int col0, int ncols,
                                            uint32_t tcol, bool vec_ok) {
#pragma unroll 1
  for (int c = col0; c < col0 + ncols; c += 32) {
    uint32_t w[16];
    if (valid && vec_ok && c + 32 <= d) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 x = *reinterpret_cast<const uint4*>(row + c + v * 8);
        w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int e = c + 2 * j;
        const uint32_t lo = (valid && e < d) ? static_cast<uint32_t>(__bfloat16_as_ushort(row[e])) : 0u;
        const uint32_t hi = (valid && e + 1 < d) ? static_cast<uint32_t>(__bfloat16_as_ushort(row[e + 1])) : 0u;
        w[j] = lo | (hi << 16);
      }
    }
    tmem_st16(tcol + ((c - col0) >> 1), w);
  }
}

template <int KB, bool ROW>
__global__ void __launch_bounds__(WideBwdCfg<KB>::THREADS, 1)
tc_softmax_bwd_wide_kernel(const __grid_constant__ CUtensorMap tmY, const __nv_bfloat16* __restrict__ x_rows, int64_t ldx,
                           int B, int d, float scale2, const float* __restrict__ lse, const float* __restrict__ yf,
                           int64_t ld_yf, const float* __restrict__ relu_mask, int64_t ld_mask, float out_scale,
                           float* __restrict__ out, int64_t ld_out) {
  using Cfg = WideBwdCfg<KB>;
  constexpr int NT = Cfg::NT, S = Cfg::STAGES, D = Cfg::D, HC = NT / 2;   // HC: S columns per softmax group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sY = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + S * Cfg::Y_BYTES);
  uint64_t* x_ready = bars;
  uint64_t* y_full = bars + 1;           // [S]
  uint64_t* y_empty = y_full + S;        // [S]
  uint64_t* s_full = y_empty + S;        // [NSP]
  uint64_t* p_full = s_full + Cfg::NSP;  // [NSP]
  uint64_t* o_full = p_full + Cfg::NSP;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* lse_tile = reinterpret_cast<float*>(bars + 32);      // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int Tall = (B + NT - 1) / NT;
  const int Tper = (Tall + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * Tper;
  const int T = max(0, min(Tper, Tall - t0));             // the host guarantees T >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    mbar_init(x_ready, 256);
    for (int s = 0; s < S; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < Cfg::NSP; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 256); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int t = 0; t < T; ++t) {
        const int s = t % S;
        uint8_t* st = sY + s * Cfg::Y_BYTES;
        mbar_wait_relaxed(&y_empty[s], ((t / S) & 1) ^ 1);
        mbar_expect_tx(&y_full[s], Cfg::Y_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(st + kb * NT * 128, &tmY, &y_full[s], kb * 64, (t0 + t) * NT);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc1 = idesc_bf16_f32(128, NT);
    constexpr uint32_t idesc2 = idesc_bf16_f32(128, D) | kIdescBMnMajor;
    if (elect_one()) {
      auto issue_gemm1 = [&](int t) {
        const int s = t % S, as = t % Cfg::NSP;
        mbar_wait_relaxed(&y_full[s], (t / S) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(sY + s * Cfg::Y_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t db = smem_desc_k_sw128(st + kb * NT * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ts(tmem_base + Cfg::S_COL + as * NT, tmem_base + Cfg::X_COL + kb * 32 + k * 8, db + 2 * k, idesc1, (kb | k) != 0);
        }
        tc_commit(&s_full[as]);
      };
      mbar_wait_relaxed(x_ready, 0);
      tc_fence_after();
      for (int i = 0; i < Cfg::LOOKAHEAD && i < T; ++i) issue_gemm1(i);
      for (int t = 0; t < T; ++t) {
        const int s = t % S, as = t % Cfg::NSP;
        // S(t + LOOKAHEAD) overwrites the stage whose P was consumed by the O-product issued one iteration ago
        // (tcgen05.mma executes in issue order), and keeps the tensor pipe busy while the softmax groups turn S(t) into P(t)
        if (t + Cfg::LOOKAHEAD < T) issue_gemm1(t + Cfg::LOOKAHEAD);
        mbar_wait_relaxed(&p_full[as], (t / Cfg::NSP) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(sY + s * Cfg::Y_BYTES);
#pragma unroll
        for (int kk = 0; kk < NT / 16; ++kk) {     // O[128, D] += P[:, 16kk..16kk+16) . Y[16kk..16kk+16, :]
          // P columns of softmax group h live at stage columns [h * HC, h * HC + HC / 2) (bf16 pairs)
          const uint32_t pa = tmem_base + Cfg::S_COL + as * NT + (kk / (HC / 16)) * HC + (kk % (HC / 16)) * 8;
          const uint64_t db = smem_desc_mn_sw128(st + kk * 2048, NT * 128, 1024);
          mma_ts(tmem_base + Cfg::O_COL, pa, db, idesc2, (t | kk) != 0);
        }
        tc_commit(&y_empty[s]);
        if (t == T - 1) tc_commit(o_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;                   // softmax group = column half of every S tile
    const int r_in = q * 32 + lane;
    const int row = m0 + r_in;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // ---- this CTA's rows -> TMEM (bf16, A operand of every S product): group h stores column half h
    {
      const bool vec_ok = (ldx % 8) == 0 && (reinterpret_cast<uintptr_t>(x_rows) & 15) == 0;
      row_to_tmem(x_rows + (int64_t)row * ldx, row < B, d, h * (D / 2), D / 2, trow + Cfg::X_COL + h * (D / 4), vec_ok);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x_ready);
    }
    const float lrow = (ROW && row < B) ? lse[row] * kLog2e : 0.f;
    float* lse_g = lse_tile + h * 128;
    for (int t = 0; t < T; ++t) {
      const int n0 = (t0 + t) * NT + h * HC;          // first logit column of this group
      const int as = t % Cfg::NSP;
      const uint32_t tsp = trow + Cfg::S_COL + as * NT + h * HC;
      if (!ROW) {
        if (r_in < HC) lse_g[r_in] = (n0 + r_in < B) ? lse[n0 + r_in] * kLog2e : 0.f;
        group_bar_sync(h);
      }
      mbar_wait(&s_full[as], (t / Cfg::NSP) & 1);
      tc_fence_after();
      const bool tail = n0 + HC > B;
      uint32_t va[32];
      if (HC == 32) {
        tmem_ld32(tsp, va);
        tmem_ld_wait();
        if (!tail) p_chunk<ROW, false>(va, tsp, lse_g, lrow, scale2, n0, B);
        else p_chunk<ROW, true>(va, tsp, lse_g, lrow, scale2, n0, B);
      } else {
        uint32_t vb[32];
        tmem_ld32(tsp, va);
        tmem_ld32(tsp + 32, vb);
        tmem_ld_wait();                                // both chunks are in registers before P overwrites their columns
        if (!tail) {
          p_chunk<ROW, false>(va, tsp, lse_g, lrow, scale2, n0, B);
          p_chunk<ROW, false>(vb, tsp + 16, lse_g + 32, lrow, scale2, n0 + 32, B);
        } else {
          p_chunk<ROW, true>(va, tsp, lse_g, lrow, scale2, n0, B);
          p_chunk<ROW, true>(vb, tsp + 16, lse_g + 32, lrow, scale2, n0 + 32, B);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[as]);
      if (!ROW) group_bar_sync(h);   // lse_g is rewritten at the top of the next iteration
    }
    // ---- O -> out = out_scale * (O - Y_r), optionally gated by relu_mask > 0 (the groups alternate 32-column chunks)
    mbar_wait(o_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = h * 32; c0 < D; c0 += 64) {
      if (c0 >= d) break;
      uint32_t v[32];
      tmem_ld32(trow + Cfg::O_COL + c0, v);
      tmem_ld_wait();
      if (row < B) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col < d) {
            float o = __uint_as_float(v[j]);
            if (blockIdx.y == 0) o -= yf[(int64_t)row * ld_yf + col];
            o *= out_scale;
            if (relu_mask != nullptr && !(relu_mask[(int64_t)row * ld_mask + col] > 0.f)) o = 0.f;
            if (gridDim.y == 1) out[(int64_t)row * ld_out + col] = o;
            else atomicAdd(&out[(int64_t)row * ld_out + col], o);   // two addends: order-independent
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Two column splits once there are at least two tiles per split and enough row blocks to matter.
static int column_splits(int64_t B, int NT) {
  const int64_t tall = (B + NT - 1) / NT;
  return (tall >= 4 && B >= 2048) ? 2 : 1;
}

template <int KB>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tcm, int B, float scale2, const float* diag, float* lse,
                      float* partial, float* ml, int splits, const float* qn2, const unsigned int* cmax2, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_softmax_fwd_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FwdCfg<KB>::SMEM);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_fwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((B + 127) / 128, splits);
  tc_softmax_fwd_kernel<KB><<<grid, FwdCfg<KB>::THREADS, FwdCfg<KB>::SMEM, s>>>(tq, tcm, B, scale2, diag, lse, partial, ml, qn2, cmax2);
  TT_CHECK_LAUNCH("tc_softmax_fwd");
  return TT_OK;
}

template <int KB, bool ROW>
static int launch_bwd(const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& tyt, int B, int d, float scale2,
                      const float* lse, const float* yf, int64_t ld_yf, const float* mask, int64_t ld_mask,
                      float out_scale, float* out, int64_t ld_out, int splits, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_softmax_bwd_kernel<KB, ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         BwdCfg<KB>::SMEM);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  if (splits > 1) {  // the splits accumulate into `out`
    cudaError_t e = cudaMemset2DAsync(out, (size_t)ld_out * 4, 0, (size_t)d * 4, (size_t)B, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd memset: %s", cudaGetErrorString(e));
  }
  dim3 grid((B + 127) / 128, splits);
  tc_softmax_bwd_kernel<KB, ROW><<<grid, BwdCfg<KB>::THREADS, BwdCfg<KB>::SMEM, s>>>(
      tx, ty, tyt, B, d, scale2, lse, yf, ld_yf, mask, ld_mask, out_scale, out, ld_out);
  TT_CHECK_LAUNCH("tc_softmax_bwd");
  return TT_OK;
}

template <int KB, bool ROW>
static int launch_bwd_wide(const CUtensorMap& ty, const void* x_bf16, int64_t ldx, int B, int d, float scale2, const float* lse,
                           const float* yf, int64_t ld_yf, const float* mask, int64_t ld_mask, float out_scale, float* out,
                           int64_t ld_out, int splits, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_softmax_bwd_wide_kernel<KB, ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WideBwdCfg<KB>::SMEM);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd_wide smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  if (splits > 1) {  // the splits accumulate into `out`
    cudaError_t e = cudaMemset2DAsync(out, (size_t)ld_out * 4, 0, (size_t)d * 4, (size_t)B, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd_wide memset: %s", cudaGetErrorString(e));
  }
  dim3 grid((B + 127) / 128, splits);
  tc_softmax_bwd_wide_kernel<KB, ROW><<<grid, WideBwdCfg<KB>::THREADS, WideBwdCfg<KB>::SMEM, s>>>(
      ty, static_cast<const __nv_bfloat16*>(x_bf16), ldx, B, d, scale2, lse, yf, ld_yf, mask, ld_mask, out_scale, out, ld_out);
  TT_CHECK_LAUNCH("tc_softmax_bwd_wide");
  return TT_OK;
}

// Column chunks of the fused backward: minimise waves x (tiles per CTA + fixed cost), fixed cost ~ 3 tiles
static int fused_chunks(int64_t B) {
  const int64_t sb = (B + 255) / 256, tall = (B + 127) / 128;
  int best = 1;
  double best_cost = 1e30;
  for (int ch = 1; ch <= 64 && ch <= tall; ++ch) {
    const int64_t tper = (tall + ch - 1) / ch;
    if ((int64_t)(ch - 1) * tper >= tall) continue;     // an empty chunk
    const int64_t waves = (sb * ch + kNumSMs - 1) / kNumSMs;
    const double cost = (double)waves * (double)(tper + 3);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ch; }
  }
  return best;
}

static int finalize_grad(float* out, int64_t ld_out, const float* other, int64_t ld_other, const float* mask, int64_t ld_mask,
                         float out_scale, const float* scale_dev, int B, int d, cudaStream_t s) {
  const bool vec = (d % 4) == 0 && (ld_out % 4) == 0 && (ld_other % 4) == 0 && (mask == nullptr || (ld_mask % 4) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(other) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
  if (vec) {
    const int64_t n = (int64_t)B * (d / 4);
    softmax_bwd_finalize_kernel<4><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, ld_out, other, ld_other, mask, ld_mask, out_scale, scale_dev, B, d);
  } else {
    const int64_t n = (int64_t)B * d;
    softmax_bwd_finalize_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, ld_out, other, ld_other, mask, ld_mask, out_scale, scale_dev, B, d);
  }
  TT_CHECK_LAUNCH("softmax_bwd_finalize");
  return TT_OK;
}

static int launch_bwd_fused(const void* q_bf16, int64_t ldq, const void* c_bf16, int64_t ldc, const float* q_f32, int64_t ldqf,
                            const float* c_f32, int64_t ldcf, const float* lse, int B, int d, float scale2, float out_scale,
                            const float* gate_q, const float* gate_c, float* dq, int64_t lddq, float* dc, int64_t lddc,
                            const float* scale_dev, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_softmax_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg::SMEM);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd_fused smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  CUtensorMap tq, tcm, tdq, tdc;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, q_bf16, B, d, ldq, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tcm, c_bf16, B, d, ldc, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&tdq, dq, B, d, lddq, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&tdc, dc, B, d, lddc, 128))) return rc;
  cudaError_t e = cudaMemset2DAsync(dq, (size_t)lddq * 4, 0, (size_t)d * 4, (size_t)B, s);
  if (e == cudaSuccess) e = cudaMemset2DAsync(dc, (size_t)lddc * 4, 0, (size_t)d * 4, (size_t)B, s);
  if (e != cudaSuccess) return fail(TT_ERR_CUDA, "softmax_bwd_fused memset: %s", cudaGetErrorString(e));
  dim3 grid((unsigned)((B + 255) / 256), (unsigned)fused_chunks(B));
  tc_softmax_bwd_fused_kernel<<<grid, FusedCfg::THREADS, FusedCfg::SMEM, s>>>(tq, tcm, tdq, tdc, B, d, scale2, lse);
  TT_CHECK_LAUNCH("tc_softmax_bwd_fused");
  if ((rc = finalize_grad(dq, lddq, c_f32, ldcf, gate_q, ldqf, out_scale, scale_dev, B, d, s))) return rc;
  return finalize_grad(dc, lddc, q_f32, ldqf, gate_c, ldcf, out_scale, scale_dev, B, d, s);
}

}  // namespace tc
}  // namespace tt

using namespace tt;
using namespace tt::tc;

static int g_softmax_bwd_mode = [] { const char* v = getenv("TT_SOFTMAX_BWD"); return (v != nullptr && strcmp(v, "split") == 0) ? 1 : 0; }();

// 64 < d <= 256: the TS-form kernels (A operand resident in TMEM); TT_SOFTMAX_WIDE=0 keeps the SS-form kernels (A/B runs)
static int g_softmax_wide = [] { const char* v = getenv("TT_SOFTMAX_WIDE"); return (v != nullptr && strcmp(v, "0") == 0) ? 0 : 1; }();

extern "C" {

int tt_set_softmax_wide_mode(int32_t on) {
  TT_CHECK_ARG(on == 0 || on == 1, "set_softmax_wide_mode: 0 (SS-form kernels) or 1 (TS-form kernels for d > 64)");
  g_softmax_wide = on;
  return TT_OK;
}

int tt_set_softmax_backward_mode(int32_t mode) {
  TT_CHECK_ARG(mode == 0 || mode == 1, "set_softmax_backward_mode: mode must be 0 (auto) or 1 (two-pass, deterministic)");
  g_softmax_bwd_mode = mode;
  return TT_OK;
}

size_t tt_inbatch_softmax_bf16_workspace_bytes(int64_t B) {
  return align_up((size_t)((B + 127) / 128 + 1) * 4, 256) + align_up((size_t)B * 2 * 2 * 4, 256) + align_up((size_t)B * 4, 256) + 512;
}

int tt_inbatch_softmax_forward_bf16(const void* q_bf16, int64_t ldq, const void* c_bf16, int64_t ldc, int64_t B,
                                    int64_t d, float inv_t, float* lse, float* diag, float* loss, void* ws,
                                    size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(B > 0 && d > 0 && q_bf16 && c_bf16 && lse && diag && loss, "inbatch_softmax_forward_bf16: bad args");
  if (d > 256) return fail(TT_ERR_UNSUPPORTED, "inbatch_softmax_forward_bf16: d > 256");
  if (B >= ((int64_t)1 << 30)) return fail(TT_ERR_UNSUPPORTED, "inbatch_softmax_forward_bf16: B too large");
  int blocks = (int)((B + 127) / 128);
  if (!ws || ws_bytes < tt_inbatch_softmax_bf16_workspace_bytes(B)) return fail(TT_ERR_WORKSPACE, "inbatch_softmax_bf16: workspace too small");
  cudaStream_t s = as_stream(stream);
  const int KB = (int)((d + 63) / 64);
  const int NT = KB <= 2 ? FwdCfg<1>::NT : 128;
  const int splits = column_splits(B, NT);
  CUtensorMap tq, tcm;
  int rc = make_tmap_bf16_2d(&tq, q_bf16, B, d, ldq, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tcm, c_bf16, B, d, ldc, NT);
  if (rc) return rc;
  float* partial = static_cast<float*>(ws);
  float* ml = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up((size_t)(blocks + 1) * 4, 256));
  float* qn2 = reinterpret_cast<float*>(reinterpret_cast<char*>(ml) + align_up((size_t)B * 2 * 2 * 4, 256));
  unsigned int* cmax2 = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(qn2) + align_up((size_t)B * 4, 256));
  static const bool no_fixed = [] { const char* v = getenv("TT_SOFTMAX_FWD"); return v != nullptr && strcmp(v, "online") == 0; }();
  if (cudaMemsetAsync(cmax2, 0, 4, s) != cudaSuccess) return fail(TT_ERR_CUDA, "inbatch_softmax_forward_bf16: memset failed");
  rowdot_bf16_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(q_bf16), ldq,
                                                            static_cast<const __nv_bfloat16*>(c_bf16), ldc, (int)B,
                                                            (int)d, inv_t, diag, qn2, cmax2);
  TT_CHECK_LAUNCH("rowdot_bf16");
  const float* qn2_arg = no_fixed ? nullptr : qn2;      // TT_SOFTMAX_FWD=online: always the online-max path (A/B runs)
  const float scale2 = inv_t * kLog2e;
  switch (KB) {
    case 1: rc = launch_fwd<1>(tq, tcm, (int)B, scale2, diag, lse, partial, ml, splits, qn2_arg, cmax2, s); break;
    case 2: rc = launch_fwd<2>(tq, tcm, (int)B, scale2, diag, lse, partial, ml, splits, qn2_arg, cmax2, s); break;
    case 3: rc = launch_fwd<3>(tq, tcm, (int)B, scale2, diag, lse, partial, ml, splits, qn2_arg, cmax2, s); break;
    default: rc = launch_fwd<4>(tq, tcm, (int)B, scale2, diag, lse, partial, ml, splits, qn2_arg, cmax2, s); break;
  }
  if (rc) return rc;
  if (splits > 1) {
    blocks = (int)((B + 255) / 256);
    lse_merge_kernel<<<blocks, 256, 0, s>>>(ml, splits, (int)B, diag, lse, partial);
    TT_CHECK_LAUNCH("lse_merge");
  }
  loss_final_kernel<<<1, 1024, 0, s>>>(partial, blocks, 1.0f / (float)B, loss);
  TT_CHECK_LAUNCH("loss_final");
  return TT_OK;
}

int tt_inbatch_softmax_backward_bf16(const void* q_bf16, int64_t ldq, const void* c_bf16, int64_t ldc,
                                     const void* qt_bf16, int64_t ldqt, const void* ct_bf16, int64_t ldct,
                                     const float* q_f32, int64_t ldqf, const float* c_f32, int64_t ldcf,
                                     const float* lse, int64_t B, int64_t d, float inv_t, float grad_scale,
                                     int32_t relu_gate, float* dq, int64_t lddq, float* dc, int64_t lddc,
                                     const float* grad_scale_dev, void* stream) {
  TT_CHECK_ARG(B > 0 && d > 0 && q_bf16 && c_bf16 && q_f32 && c_f32 && lse && dq && dc,
               "inbatch_softmax_backward_bf16: bad args");
  if (d > 256) return fail(TT_ERR_UNSUPPORTED, "inbatch_softmax_backward_bf16: d > 256");
  cudaStream_t s = as_stream(stream);
  const int KB = (int)((d + 63) / 64);
  const int D = KB * 64;
  const int NT = KB <= 2 ? BwdCfg<1>::NT : 64;
  const int splits = column_splits(B, NT);
  const float scale2 = inv_t * kLog2e;
  const float out_scale = grad_scale * inv_t / (float)B;
  // d <= 64: one pass over P for both gradients, unless the deterministic two-pass kernels were asked for
  if (KB == 1 && g_softmax_bwd_mode == 0 && (lddq % 4) == 0 && (lddc % 4) == 0 &&
      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dc)) & 15) == 0)
    return launch_bwd_fused(q_bf16, ldq, c_bf16, ldc, q_f32, ldqf, c_f32, ldcf, lse, (int)B, (int)d, scale2, out_scale,
                            relu_gate ? q_f32 : nullptr, relu_gate ? c_f32 : nullptr, dq, lddq, dc, lddc, grad_scale_dev, s);
  if (grad_scale_dev != nullptr) return fail(TT_ERR_UNSUPPORTED, "inbatch_softmax_backward_bf16: grad_scale_dev needs the one-pass path (d <= 64)");
  if (KB >= 2 && g_softmax_wide) {
    // TS-form two-pass kernels: row-major operands only (the streamed tile doubles as the MN-major B of the second product)
    const int NTW = KB <= 2 ? 128 : 64;
    const int sp = column_splits(B, NTW);
    CUtensorMap tqn, tcn;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tqn, q_bf16, B, d, ldq, NTW))) return rc;
    if ((rc = make_tmap_bf16_2d(&tcn, c_bf16, B, d, ldc, NTW))) return rc;
    const float* gq = relu_gate ? q_f32 : nullptr;
    const float* gc = relu_gate ? c_f32 : nullptr;
#define TT_LBW(KBV)                                                                                                          \
  do {                                                                                                                       \
    rc = launch_bwd_wide<KBV, true>(tcn, q_bf16, ldq, (int)B, (int)d, scale2, lse, c_f32, ldcf, gq, ldqf, out_scale, dq, lddq, sp, s); \
    if (rc) return rc;                                                                                                       \
    rc = launch_bwd_wide<KBV, false>(tqn, c_bf16, ldc, (int)B, (int)d, scale2, lse, q_f32, ldqf, gc, ldcf, out_scale, dc, lddc, sp, s); \
  } while (0)
    switch (KB) {
      case 2: TT_LBW(2); break;
      case 3: TT_LBW(3); break;
      default: TT_LBW(4); break;
    }
#undef TT_LBW
    return rc;
  }
  TT_CHECK_ARG(qt_bf16 && ct_bf16, "inbatch_softmax_backward_bf16: the two-pass kernels need the transposed copies");
  CUtensorMap tq128, tc128, tqn, tcn, tqt, tct;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq128, q_bf16, B, d, ldq, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tc128, c_bf16, B, d, ldc, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tqn, q_bf16, B, d, ldq, NT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tcn, c_bf16, B, d, ldc, NT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tqt, qt_bf16, d, B, ldqt, D))) return rc;
  if ((rc = make_tmap_bf16_2d(&tct, ct_bf16, d, B, ldct, D))) return rc;
  const float* gate_q = relu_gate ? q_f32 : nullptr;
  const float* gate_c = relu_gate ? c_f32 : nullptr;
#define TT_LB(KBV)                                                                                              \
  do {                                                                                                          \
    rc = launch_bwd<KBV, true>(tq128, tcn, tct, (int)B, (int)d, scale2, lse, c_f32, ldcf, gate_q, ldqf, out_scale, dq, lddq, splits, s); \
    if (rc) return rc;                                                                                          \
    rc = launch_bwd<KBV, false>(tc128, tqn, tqt, (int)B, (int)d, scale2, lse, q_f32, ldqf, gate_c, ldcf, out_scale, dc, lddc, splits, s); \
  } while (0)
  switch (KB) {
    case 1: TT_LB(1); break;
    case 2: TT_LB(2); break;
    case 3: TT_LB(3); break;
    default: TT_LB(4); break;
  }
#undef TT_LB
  return rc;
}

}  // extern "C"
