// The two ReLU towers of the model (query_proj / candidate_proj, reference utils/model_training.py:95-96:
// MLP(embedding_dim, layer_sizes) = Linear+ReLU, Linear+ReLU) as TWO kernels per step instead of ~40:
//
//   towers_fwd_fused_kernel   x (fp32 window of the pooled embeddings) -> bf16 -> tcgen05 GEMM 1 -> +b1, ReLU -> h (bf16, stays
//                             in shared memory as the A operand of) GEMM 2 -> +b2, ReLU -> y (fp32) and its bf16 copy
//   towers_bwd_fused_kernel   dy -> dz2 = dy * (y > 0) -> dh = dz2 W2 -> dz1 = dh * (h > 0) -> dx = dz1 W1, and in the same pass
//                             dW2^T += h^T dz2, dW1 += dz1^T x (accumulated in TMEM over all row tiles of the CTA),
//                             db2 / db1 += column sums of dz2 / dz1 (warp transpose-reduce, registers)
//   towers_grad_reduce_kernel adds the per-CTA partial weight / bias gradients in CTA order (deterministic)
//
// Both towers run in one launch (blockIdx.y = tower); a CTA walks row tiles of 128 samples.  Every operand the
// tensor core needs "transposed" (W2 for dh, W1 for dx, h / dz / x for the weight gradients) is the SAME shared
// memory tile read through an MN-major descriptor, so no transposed copy is ever made.  The kernels are HBM-bound:
// per sample the forward moves 4*in + 2*(64 + 128 + 64) + 4*out bytes, the backward 4*out + 2*(64+128+64) + 4*in.
// Shapes: in <= 64, hidden <= 128, out <= 64, all multiples of 8 (wider towers take the per-layer GEMM path).
//
// warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer | warps 2-5: workers, thread = row of the tile.
#include <string.h>

#include "tc_common.cuh"

namespace tt {
namespace tc {

constexpr int kTwThreads = 192;
constexpr int kPanel = 128 * 128;      // one [128 rows x 64] bf16 panel, 128-byte swizzle: 16 KB
constexpr int kMaxTowers = TT_MAX_TOWERS;

__device__ __forceinline__ void worker_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
// Waits of the single-thread roles: back off instead of burning issue slots next to the workers.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    __nanosleep(64);
  }
  printf("tt_b200: towers mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
// bf16 > 0 for the low / high half of a packed pair
__device__ __forceinline__ bool bf16_pos_lo(uint32_t w) { return (w & 0x8000u) == 0 && (w & 0x7fffu) != 0; }
__device__ __forceinline__ bool bf16_pos_hi(uint32_t w) { return (w & 0x80000000u) == 0 && (w & 0x7fff0000u) != 0; }

// v[i] (this lane = one row) -> v[0] = sum over the 32 lanes of column `lane`
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------ forward
struct TowerFwdArgs {
  const float* x; int64_t ldx;
  const float* b1; const float* b2;
  __nv_bfloat16* xb; __nv_bfloat16* hb; __nv_bfloat16* yb;   // row pitches 64, 128, 64 (zero padded columns)
  float* y; int64_t ldy;
};
struct TowerFwdParams {
  CUtensorMap w1[kMaxTowers], w2[kMaxTowers];
  CUtensorMap xb[kMaxTowers], hb[kMaxTowers], yb[kMaxTowers], y[kMaxTowers];   // outputs leave through TMA stores
  TowerFwdArgs a[kMaxTowers];
  int B, in_dim, hidden, out_dim, tiles;
};
constexpr int kFwdSmem = 2 * kPanel /*W1, W2*/ + kPanel /*x*/ + 2 * kPanel /*h*/ + 1024 /*biases*/ + 256 /*barriers*/ + 1024;

__global__ void __launch_bounds__(kTwThreads, 2)
towers_fwd_fused_kernel(const __grid_constant__ TowerFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;                 // [128 hidden rows x 64 k]           K-major B of GEMM 1
  uint8_t* sW2 = sW1 + kPanel;         // 2 x [64 out rows x 64 k] (8 KB)    K-major B of GEMM 2
  uint8_t* sX = sW2 + kPanel;          // [128 rows x 64]                    K-major A of GEMM 1
  uint8_t* sH = sX + kPanel;           // 2 x [128 rows x 64]                K-major A of GEMM 2
  float* sB1 = reinterpret_cast<float*>(sH + 2 * kPanel);   // [128]
  float* sB2 = sB1 + 128;                                   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + 256);
  uint64_t* w_full = bars;
  uint64_t* x_ready = bars + 1;   // workers -> MMA: bf16 x tile is in smem
  uint64_t* h_full = bars + 2;    // MMA -> workers: GEMM 1 accumulator complete
  uint64_t* h_ready = bars + 3;   // workers -> MMA: bf16 h tile is in smem
  uint64_t* y_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.y;
  const TowerFwdArgs& a = p.a[tw];
  const int n_it = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.w1[tw]);
    prefetch_tmap(&p.w2[tw]);
    prefetch_tmap(&p.xb[tw]); prefetch_tmap(&p.hb[tw]); prefetch_tmap(&p.yb[tw]); prefetch_tmap(&p.y[tw]);
    mbar_init(w_full, 1);
    mbar_init(x_ready, 128);
    mbar_init(h_full, 1);
    mbar_init(h_ready, 128);
    mbar_init(y_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, 2 * kPanel);
      tma_load_2d(sW1, &p.w1[tw], w_full, 0, 0);
      tma_load_2d(sW2, &p.w2[tw], w_full, 0, 0);
      tma_load_2d(sW2 + kPanel / 2, &p.w2[tw], w_full, 64, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc1 = idesc_bf16_f32(128, 128), idesc2 = idesc_bf16_f32(128, 64);
      const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aX = smem_u32(sX), aH = smem_u32(sH);
      mbar_wait_sleep(w_full, 0);
      for (int it = 0; it < n_it; ++it) {
        mbar_wait_sleep(x_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_ss(tmem_base, smem_desc_k_sw128(aX) + 2 * k, smem_desc_k_sw128(aW1) + 2 * k, idesc1, k != 0);
        tc_commit(h_full);
        mbar_wait_sleep(h_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          mma_ss(tmem_base + 128, smem_desc_k_sw128(aH + (kk >> 2) * kPanel) + 2 * (kk & 3),
                 smem_desc_k_sw128(aW2 + (kk >> 2) * (kPanel / 2)) + 2 * (kk & 3), idesc2, kk != 0);
        tc_commit(y_full);
      }
    }
  } else {
    // All global traffic of the workers is coalesced: x is read a row per warp instruction (lane = column pair), and
    // the outputs leave as TMA stores of tiles that sit in shared memory anyway -- the bf16 operand tiles of the two
    // GEMMs ARE xb and hb; y / yb are staged over them once the GEMMs are done.  (Thread-per-row global accesses cost
    // 32 L1 wavefronts per request: the first version of this kernel ran at 18 % of DRAM bandwidth because of it.)
    const int q = warp & 3;
    const int r = q * 32 + lane, r7 = r & 7;
    const int wq = warp - 2;                          // row block of the coalesced x load
    const bool leader = threadIdx.x == 64;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t aX = smem_u32(sX), aH = smem_u32(sH);
    const uint32_t xrow = aX + r * 128, hrow = aH + r * 128;
    sB1[r] = (r < p.hidden && a.b1 != nullptr) ? a.b1[r] : 0.f;
    if (r < 64) sB2[r] = (r < p.out_dim && a.b2 != nullptr) ? a.b2[r] : 0.f;
    worker_bar_sync();
    for (int it = 0; it < n_it; ++it) {
      const int row0 = (blockIdx.x + it * gridDim.x) * 128;
      // ---- x (fp32) -> bf16 tile in shared memory = A of GEMM 1 = the xb output
      if (leader) bulk_wait_group_read0();            // the y / yb stores of the previous tile have read sH / sX
      worker_bar_sync();
#pragma unroll
      for (int rr = 0; rr < 32; rr += 8) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t row = (int64_t)row0 + wq * 32 + rr + u;
          v[u] = make_float2(0.f, 0.f);
          if (row < p.B && 2 * lane < p.in_dim) v[u] = __ldg(reinterpret_cast<const float2*>(a.x + row * a.ldx + 2 * lane));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rl = wq * 32 + rr + u;
          const uint32_t addr = aX + rl * 128 + ((((lane >> 2) ^ (rl & 7))) << 4) + (lane & 3) * 4;
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16(v[u].x, v[u].y)) : "memory");
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(x_ready);
      worker_bar_sync();
      if (leader) {
        tma_store_2d(&p.xb[tw], aX, 0, row0);
        bulk_commit_group();
      }
      // ---- h = relu(acc1 + b1) -> bf16 tile = A of GEMM 2 = the hb output
      mbar_wait(h_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + 32 * c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(sB1 + 32 * c + j);
          pk[j >> 1] = pack_bf16(fmaxf(__uint_as_float(v[j]) + b.x, 0.f), fmaxf(__uint_as_float(v[j + 1]) + b.y, 0.f));
          pk[(j >> 1) + 1] = pack_bf16(fmaxf(__uint_as_float(v[j + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[j + 3]) + b.w, 0.f));
        }
        const uint32_t prow = hrow + (c >> 1) * kPanel;
#pragma unroll
        for (int m = 0; m < 4; ++m)
          st_shared_v4(prow + ((((c & 1) * 4 + m) ^ r7) << 4), pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(h_ready);
      worker_bar_sync();
      if (leader) {
        tma_store_2d(&p.hb[tw], aH, 0, row0);
        if (p.hidden > 64) tma_store_2d(&p.hb[tw], aH + kPanel, 64, row0);
        bulk_commit_group();
      }
      // ---- y = relu(acc2 + b2): staged over the (now idle) operand tiles, fp32 in the h panels, bf16 in the x tile
      mbar_wait(y_full, it & 1);
      tc_fence_after();
      if (leader) bulk_wait_group_read0();            // the xb / hb stores have read the tiles
      worker_bar_sync();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + 128 + 32 * c, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(sB2 + 32 * c + j);
          f[j] = fmaxf(__uint_as_float(v[j]) + b.x, 0.f);
          f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + b.y, 0.f);
          f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + b.z, 0.f);
          f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + b.w, 0.f);
        }
#pragma unroll
        for (int m = 0; m < 8; ++m)      // fp32 box c: [128 rows x 32 columns], 128-byte swizzle
          st_shared_v4(hrow + c * kPanel + ((m ^ r7) << 4), __float_as_uint(f[4 * m]), __float_as_uint(f[4 * m + 1]),
                       __float_as_uint(f[4 * m + 2]), __float_as_uint(f[4 * m + 3]));
#pragma unroll
        for (int m = 0; m < 4; ++m)      // bf16 tile: columns 32c .. 32c+31 are 16-byte slots 4c .. 4c+3
          st_shared_v4(xrow + (((4 * c + m) ^ r7) << 4), pack_bf16(f[8 * m], f[8 * m + 1]), pack_bf16(f[8 * m + 2], f[8 * m + 3]),
                       pack_bf16(f[8 * m + 4], f[8 * m + 5]), pack_bf16(f[8 * m + 6], f[8 * m + 7]));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      worker_bar_sync();
      if (leader) {
        tma_store_2d(&p.y[tw], aH, 0, row0);
        if (p.out_dim > 32) tma_store_2d(&p.y[tw], aH + kPanel, 32, row0);
        tma_store_2d(&p.yb[tw], aX, 0, row0);
        bulk_commit_group();
      }
    }
    if (leader) bulk_wait_group0();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------ backward
struct TowerBwdArgs {
  const float* dy; int64_t lddy;   // [B, out] gradient of the tower output (after its ReLU)
  float* dx; int64_t lddx;         // [B, in]  gradient of the tower input, or null
  float* ws;                       // partials: [ctas][128][64] dW1, [ctas][128][64] dW2^T, [ctas][4][192] bias sums
};
struct TowerBwdParams {
  CUtensorMap w1[kMaxTowers], w2[kMaxTowers], yb[kMaxTowers], hb[kMaxTowers], xb[kMaxTowers], dx[kMaxTowers];
  TowerBwdArgs a[kMaxTowers];
  int B, in_dim, hidden, out_dim, tiles;
};
constexpr int kBwdSmem = 2 * kPanel /*W1, W2*/ + kPanel /*y -> dz2*/ + 2 * kPanel /*h -> dz1*/ + kPanel /*x*/ + 256 + 1024;
constexpr int kWsW = 128 * 64, kWsB = 4 * 192;    // floats per CTA

__global__ void __launch_bounds__(kTwThreads, 2)
towers_bwd_fused_kernel(const __grid_constant__ TowerBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;                 // [128 hidden rows x 64 in]: MN-major B of dx = dz1 W1 (N = in, K = hidden)
  uint8_t* sW2 = sW1 + kPanel;         // 2 x [64 out rows x 64 hidden]: MN-major B of dh = dz2 W2 (N = hidden, K = out)
  uint8_t* sZ2 = sW2 + kPanel;         // y tile (mask), overwritten by dz2: K-major A of dh, MN-major B of dW2^T
  uint8_t* sH = sZ2 + kPanel;          // h tile: MN-major A of dW2^T, mask; overwritten by dz1: K-major A of dx, MN-major A of dW1
  uint8_t* sX = sH + 2 * kPanel;       // x tile: MN-major B of dW1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + kPanel);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;    // TMA -> workers
  uint64_t* in_empty = bars + 2;   // workers' leader -> TMA: products done AND the dx store has read its staging (the h tile)
  uint64_t* z2_ready = bars + 3;   // workers -> MMA
  uint64_t* dh_full = bars + 4;    // MMA -> workers
  uint64_t* w2_done = bars + 5;    // MMA -> workers: h may be overwritten by dz1
  uint64_t* z1_ready = bars + 6;
  uint64_t* dx_full = bars + 7;    // MMA -> workers: ALL products of the tile are done (dx complete, operand tiles idle)
  uint64_t* all_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  constexpr uint32_t kWork = 0, kAccW1 = 128, kAccW2 = 192;   // TMEM columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.y;
  const TowerBwdArgs& a = p.a[tw];
  const int n_it = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.w1[tw]); prefetch_tmap(&p.w2[tw]); prefetch_tmap(&p.yb[tw]); prefetch_tmap(&p.hb[tw]); prefetch_tmap(&p.xb[tw]);
    if (p.a[tw].dx != nullptr) prefetch_tmap(&p.dx[tw]);
    mbar_init(w_full, 1);
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    mbar_init(z2_ready, 128);
    mbar_init(dh_full, 1);
    mbar_init(w2_done, 1);
    mbar_init(z1_ready, 128);
    mbar_init(dx_full, 1);
    mbar_init(all_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, 2 * kPanel);
      tma_load_2d(sW1, &p.w1[tw], w_full, 0, 0);
      tma_load_2d(sW2, &p.w2[tw], w_full, 0, 0);
      tma_load_2d(sW2 + kPanel / 2, &p.w2[tw], w_full, 64, 0);
      for (int it = 0; it < n_it; ++it) {
        const int row0 = (blockIdx.x + it * gridDim.x) * 128;
        mbar_wait_sleep(in_empty, (it & 1) ^ 1);
        mbar_expect_tx(in_full, 4 * kPanel);
        tma_load_2d(sZ2, &p.yb[tw], in_full, 0, row0);
        tma_load_2d(sH, &p.hb[tw], in_full, 0, row0);
        tma_load_2d(sH + kPanel, &p.hb[tw], in_full, 64, row0);
        tma_load_2d(sX, &p.xb[tw], in_full, 0, row0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idescDh = idesc_bf16_f32(128, 128) | kIdescBMnMajor;
      constexpr uint32_t idescDx = idesc_bf16_f32(128, 64) | kIdescBMnMajor;
      constexpr uint32_t idescDw = idesc_bf16_f32(128, 64) | kIdescAMnMajor | kIdescBMnMajor;
      const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aZ2 = smem_u32(sZ2), aH = smem_u32(sH), aX = smem_u32(sX);
      mbar_wait_sleep(w_full, 0);
      for (int it = 0; it < n_it; ++it) {
        mbar_wait_sleep(z2_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)       // dh[r, h] = sum_o dz2[r, o] W2[o, h]
          mma_ss(tmem_base + kWork, smem_desc_k_sw128(aZ2) + 2 * k, smem_desc_mn_sw128(aW2 + k * 2048, kPanel / 2, 1024), idescDh, k != 0);
        tc_commit(dh_full);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)    // dW2^T[h, o] += sum_r h[r, h] dz2[r, o]
          mma_ss(tmem_base + kAccW2, smem_desc_mn_sw128(aH + kk * 2048, kPanel, 1024), smem_desc_mn_sw128(aZ2 + kk * 2048, 1024, 1024),
                 idescDw, (it | kk) != 0);
        tc_commit(w2_done);
        mbar_wait_sleep(z1_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)    // dx[r, i] = sum_h dz1[r, h] W1[h, i]
          mma_ss(tmem_base + kWork, smem_desc_k_sw128(aH + (kk >> 2) * kPanel) + 2 * (kk & 3),
                 smem_desc_mn_sw128(aW1 + kk * 2048, 1024, 1024), idescDx, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)    // dW1[h, i] += sum_r dz1[r, h] x[r, i]
          mma_ss(tmem_base + kAccW1, smem_desc_mn_sw128(aH + kk * 2048, kPanel, 1024), smem_desc_mn_sw128(aX + kk * 2048, 1024, 1024),
                 idescDw, (it | kk) != 0);
        tc_commit(dx_full);
      }
      tc_commit(all_done);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane, r7 = r & 7;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t zrow = smem_u32(sZ2) + r * 128, hrow = smem_u32(sH) + r * 128;
    float db[4] = {0.f, 0.f, 0.f, 0.f};             // column sums of dz1: lane = column of chunk c
    float db2_lo = 0.f, db2_hi = 0.f;                // column sums of dz2: lane = column pair (2 lane, 2 lane + 1)
    const int wq = warp - 2;                          // row block of the coalesced dy pass
    const bool leader = threadIdx.x == 64;
    const uint32_t aZ2 = smem_u32(sZ2), aHs = smem_u32(sH);
    for (int it = 0; it < n_it; ++it) {
      const int row0 = (blockIdx.x + it * gridDim.x) * 128;
      mbar_wait(in_full, it & 1);
      // ---- dz2 = dy * (y > 0) -> bf16, in place over the y tile.  Coalesced: one dy row per warp instruction
      // (lane = column pair), the matching bf16 pair of y sits at a bank-conflict-free word of the swizzled tile.
#pragma unroll
      for (int rr = 0; rr < 32; rr += 8) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t row = (int64_t)row0 + wq * 32 + rr + u;
          v[u] = make_float2(0.f, 0.f);
          if (row < p.B && 2 * lane < p.out_dim) v[u] = __ldg(reinterpret_cast<const float2*>(a.dy + row * a.lddy + 2 * lane));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rl = wq * 32 + rr + u;
          const uint32_t addr = aZ2 + rl * 128 + (((lane >> 2) ^ (rl & 7)) << 4) + (lane & 3) * 4;
          uint32_t yv;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(yv) : "r"(addr));
          const uint32_t pk = pack_bf16(bf16_pos_lo(yv) ? v[u].x : 0.f, bf16_pos_hi(yv) ? v[u].y : 0.f);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pk) : "memory");
          db2_lo += __uint_as_float(pk << 16);          // the bias gradient sums what the tensor core sees (bf16)
          db2_hi += __uint_as_float(pk & 0xffff0000u);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(z2_ready);
      // ---- dz1 = dh * (h > 0) -> bf16, in place over the h tile (once dW2^T has read it)
      mbar_wait(dh_full, it & 1);
      mbar_wait(w2_done, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + kWork + 32 * c, v);
        tmem_ld_wait();
        float f[32];
        const uint32_t prow = hrow + (c >> 1) * kPanel;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const uint32_t addr = prow + ((((c & 1) * 4 + m) ^ r7) << 4);
          const uint4 hv = ld_shared_v4(addr);
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = bf16_pos_lo(hw[e]) ? __uint_as_float(v[m * 8 + 2 * e]) : 0.f;
            const float hi = bf16_pos_hi(hw[e]) ? __uint_as_float(v[m * 8 + 2 * e + 1]) : 0.f;
            pk[e] = pack_bf16(lo, hi);
            f[m * 8 + 2 * e] = __uint_as_float(pk[e] << 16);
            f[m * 8 + 2 * e + 1] = __uint_as_float(pk[e] & 0xffff0000u);
          }
          st_shared_v4(addr, pk[0], pk[1], pk[2], pk[3]);
        }
        db[c] += warp_column_sums(f, lane);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(z1_ready);
      // ---- dx out: staged as two fp32 boxes over the (now idle) h tile, one TMA store; the producer may refill the
      // operand tiles only after that store has read its staging
      mbar_wait(dx_full, it & 1);
      tc_fence_after();
      if (a.dx != nullptr) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(trow + kWork + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < 8; ++m)
            st_shared_v4(hrow + c * kPanel + ((m ^ r7) << 4), v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        worker_bar_sync();
        if (leader) {
          tma_store_2d(&p.dx[tw], aHs, 0, row0);
          if (p.in_dim > 32) tma_store_2d(&p.dx[tw], aHs + kPanel, 32, row0);
          bulk_commit_group();
          bulk_wait_group_read0();
          mbar_arrive(in_empty);
        }
      } else {
        worker_bar_sync();
        if (leader) mbar_arrive(in_empty);
      }
    }
    if (leader) bulk_wait_group0();
    // ---- this CTA's weight / bias gradient partials
    mbar_wait(all_done, 0);
    tc_fence_after();
    float* wsw1 = a.ws + (int64_t)blockIdx.x * kWsW + (int64_t)r * 64;
    float* wsw2 = a.ws + (int64_t)gridDim.x * kWsW + (int64_t)blockIdx.x * kWsW + (int64_t)r * 64;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + kAccW1 + 32 * c, v);      // columns [128, 256): dW1 then dW2^T
      tmem_ld_wait();
      float* o = (c < 2 ? wsw1 : wsw2) + 32 * (c & 1);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    float* wsb = a.ws + 2 * (int64_t)gridDim.x * kWsW + (int64_t)blockIdx.x * kWsB + q * 192;
    wsb[2 * lane] = db2_lo;
    wsb[2 * lane + 1] = db2_hi;
#pragma unroll
    for (int c = 0; c < 4; ++c) wsb[64 + 32 * c + lane] = db[c];
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// dW1[h, i], dW2[o, h], db1[h], db2[o] = sums of the per-CTA partials, in CTA order.
// 256 threads = 64 outputs x 4 interleaved CTA lanes; the four lane sums are added in a fixed order.
struct TowerGradOut {
  const float* ws; float* dw1; float* dw2; float* db1; float* db2;
};
struct TowerGradParams { TowerGradOut t[kMaxTowers]; int ctas, in_dim, hidden, out_dim; };

__global__ void __launch_bounds__(256)
towers_grad_reduce_kernel(const __grid_constant__ TowerGradParams p) {
  __shared__ float red[4][64];
  const TowerGradOut& t = p.t[blockIdx.y];
  const int n1 = p.hidden * p.in_dim, n2 = p.out_dim * p.hidden, n3 = p.hidden, n4 = p.out_dim;
  const int e = blockIdx.x * 64 + (threadIdx.x & 63), z = threadIdx.x >> 6;
  float s = 0.f;
  if (e < n1 + n2 + n3 + n4) {
    if (e < n1 + n2) {
      const float* base;
      if (e < n1) base = t.ws + (int64_t)(e / p.in_dim) * 64 + (e % p.in_dim);
      else { const int o = (e - n1) / p.hidden, h = (e - n1) % p.hidden; base = t.ws + (int64_t)p.ctas * kWsW + (int64_t)h * 64 + o; }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four loads in flight; fixed order, so the result is reproducible
      int c = z;
      for (; c + 12 < p.ctas; c += 16) {
        s0 += base[(int64_t)c * kWsW]; s1 += base[(int64_t)(c + 4) * kWsW];
        s2 += base[(int64_t)(c + 8) * kWsW]; s3 += base[(int64_t)(c + 12) * kWsW];
      }
      for (; c < p.ctas; c += 4) s0 += base[(int64_t)c * kWsW];
      s = (s0 + s1) + (s2 + s3);
    } else {
      const int col = e < n1 + n2 + n3 ? 64 + (e - n1 - n2) : (e - n1 - n2 - n3);
      const float* base = t.ws + 2 * (int64_t)p.ctas * kWsW + col;
      for (int c = z; c < p.ctas; c += 4) {
        const float* b = base + (int64_t)c * kWsB;
        s += (b[0] + b[192]) + (b[384] + b[576]);
      }
    }
  }
  red[z][threadIdx.x & 63] = s;
  __syncthreads();
  if (z == 0 && e < n1 + n2 + n3 + n4) {
    const int i = threadIdx.x;
    const float v = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
    if (e < n1) { if (t.dw1) t.dw1[e] = v; }
    else if (e < n1 + n2) { if (t.dw2) t.dw2[e - n1] = v; }
    else if (e < n1 + n2 + n3) { if (t.db1) t.db1[e - n1 - n2] = v; }
    else if (t.db2) t.db2[e - n1 - n2 - n3] = v;
  }
}

static int towers_grid(int tiles) { return tiles < kNumSMs ? tiles : kNumSMs; }

static int check_shapes(int32_t n, int64_t B, int32_t in_dim, int32_t hidden, int32_t out_dim) {
  if (n < 1 || n > kMaxTowers) return fail(TT_ERR_INVALID, "towers: 1..%d towers", kMaxTowers);
  if (B <= 0 || B >= ((int64_t)1 << 31) - 128) return fail(TT_ERR_INVALID, "towers: bad batch size");
  if (in_dim < 8 || in_dim > 64 || hidden < 8 || hidden > 128 || out_dim < 8 || out_dim > 64 || (in_dim | hidden | out_dim) % 8 != 0)
    return fail(TT_ERR_UNSUPPORTED, "towers: fused path needs in <= 64, hidden <= 128, out <= 64, multiples of 8");
  return TT_OK;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tc
}  // namespace tt

using namespace tt;
using namespace tt::tc;

extern "C" {

size_t tt_towers_backward_workspace_bytes(int64_t B) {
  const int ctas = towers_grid((int)((B + 127) / 128));
  return (size_t)kMaxTowers * ((size_t)ctas * (2 * kWsW + kWsB) * 4 + 256);
}

int tt_towers_forward_fused(const tt_tower_forward* towers, int32_t n_towers, int64_t B, int32_t in_dim, int32_t hidden,
                            int32_t out_dim, void* stream) {
  TT_CHECK_ARG(towers != nullptr, "towers_forward: null");
  int rc = check_shapes(n_towers, B, in_dim, hidden, out_dim);
  if (rc) return rc;
  TowerFwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = (int)B; p.in_dim = in_dim; p.hidden = hidden; p.out_dim = out_dim; p.tiles = (int)((B + 127) / 128);
  for (int t = 0; t < n_towers; ++t) {
    const tt_tower_forward& s = towers[t];
    TT_CHECK_ARG(s.x && s.w1_bf16 && s.w2_bf16 && s.xb && s.hb && s.yb && s.y, "towers_forward: null pointer");
    if (!aligned16(s.x) || (s.ldx % 4) != 0 || !aligned16(s.y) || (s.ldy % 4) != 0 || !aligned16(s.xb) || !aligned16(s.hb) || !aligned16(s.yb))
      return fail(TT_ERR_INVALID, "towers_forward: x / y must be 16-byte aligned with pitches that are multiples of 4");
    if ((rc = make_tmap_bf16_2d(&p.w1[t], s.w1_bf16, hidden, in_dim, s.ldw1, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.w2[t], s.w2_bf16, out_dim, hidden, s.ldw2, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.xb[t], s.xb, B, 64, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.hb[t], s.hb, B, 128, 128, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.yb[t], s.yb, B, 64, 64, 128))) return rc;
    if ((rc = make_tmap_f32_2d(&p.y[t], s.y, B, out_dim, s.ldy, 128))) return rc;
    p.a[t] = TowerFwdArgs{s.x, s.ldx, s.b1, s.b2, static_cast<__nv_bfloat16*>(s.xb), static_cast<__nv_bfloat16*>(s.hb),
                          static_cast<__nv_bfloat16*>(s.yb), s.y, s.ldy};
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(towers_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "towers_fwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(towers_grid(p.tiles), n_towers);
  towers_fwd_fused_kernel<<<grid, kTwThreads, kFwdSmem, as_stream(stream)>>>(p);
  TT_CHECK_LAUNCH("towers_fwd_fused");
  return TT_OK;
}

int tt_towers_backward_fused(const tt_tower_backward* towers, int32_t n_towers, int64_t B, int32_t in_dim, int32_t hidden,
                             int32_t out_dim, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(towers != nullptr, "towers_backward: null");
  int rc = check_shapes(n_towers, B, in_dim, hidden, out_dim);
  if (rc) return rc;
  if (!ws || ws_bytes < tt_towers_backward_workspace_bytes(B)) return fail(TT_ERR_WORKSPACE, "towers_backward: workspace too small");
  TowerBwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = (int)B; p.in_dim = in_dim; p.hidden = hidden; p.out_dim = out_dim; p.tiles = (int)((B + 127) / 128);
  const int ctas = towers_grid(p.tiles);
  TowerGradParams g;
  memset(&g, 0, sizeof(g));
  g.ctas = ctas; g.in_dim = in_dim; g.hidden = hidden; g.out_dim = out_dim;
  const size_t per_tower = ((size_t)ctas * (2 * kWsW + kWsB) * 4 + 255) / 256 * 256;
  for (int t = 0; t < n_towers; ++t) {
    const tt_tower_backward& s = towers[t];
    TT_CHECK_ARG(s.dy && s.w1_bf16 && s.w2_bf16 && s.xb && s.hb && s.yb, "towers_backward: null pointer");
    if (!aligned16(s.dy) || (s.lddy % 4) != 0 || (s.dx && (!aligned16(s.dx) || (s.lddx % 4) != 0)))
      return fail(TT_ERR_INVALID, "towers_backward: dy / dx must be 16-byte aligned with pitches that are multiples of 4");
    if ((rc = make_tmap_bf16_2d(&p.w1[t], s.w1_bf16, hidden, in_dim, s.ldw1, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.w2[t], s.w2_bf16, out_dim, hidden, s.ldw2, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.yb[t], s.yb, B, 64, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.hb[t], s.hb, B, 128, 128, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&p.xb[t], s.xb, B, 64, 64, 128))) return rc;
    if (s.dx && (rc = make_tmap_f32_2d(&p.dx[t], s.dx, B, in_dim, s.lddx, 128))) return rc;
    float* wst = reinterpret_cast<float*>(static_cast<char*>(ws) + t * per_tower);
    p.a[t] = TowerBwdArgs{s.dy, s.lddy, s.dx, s.lddx, wst};
    g.t[t] = TowerGradOut{wst, s.dw1, s.dw2, s.db1, s.db2};
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(towers_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "towers_bwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaStream_t s = as_stream(stream);
  dim3 grid(ctas, n_towers);
  towers_bwd_fused_kernel<<<grid, kTwThreads, kBwdSmem, s>>>(p);
  TT_CHECK_LAUNCH("towers_bwd_fused");
  const int outputs = hidden * in_dim + out_dim * hidden + hidden + out_dim;
  dim3 rgrid((outputs + 63) / 64, n_towers);
  towers_grad_reduce_kernel<<<rgrid, 256, 0, s>>>(g);
  TT_CHECK_LAUNCH("towers_grad_reduce");
  return TT_OK;
}

}  // extern "C"
