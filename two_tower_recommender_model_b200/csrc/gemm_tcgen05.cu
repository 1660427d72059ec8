// C[M,N] = epilogue(A[M,K] . B[N,K]^T) on the 5th-gen tensor cores: bf16 operands staged in
// shared memory by TMA (128-byte swizzle), tcgen05.mma issued by one thread, fp32
// accumulator in TMEM, epilogue straight out of TMEM (bias, ReLU, ReLU-backward mask,
// fp32 / bf16 / transposed-bf16 stores).  One CTA per 128 x BN output tile:
//   warp 0   : TMA producer            warp 1 : TMEM allocator + MMA issuer
//   warps 2-5: epilogue (warp w reads TMEM lanes 32*(w%4) ..)
#include <stdlib.h>

#include "tc_common.cuh"

namespace tt {
namespace tc {

// ------------------------------------------------------------------ tensor map encode (driver entry point, no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(TT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0)
    return fail(TT_ERR_INVALID, "TMA operand must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
  if (rows <= 0 || cols <= 0 || box_rows <= 0 || box_rows > 256) return fail(TT_ERR_INVALID, "bad TMA shape");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TT_OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(TT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 4) != 0)
    return fail(TT_ERR_INVALID, "TMA fp32 operand must be 16-byte aligned with a row pitch that is a multiple of 4 elements");
  if (rows <= 0 || cols <= 0 || box_rows <= 0 || box_rows > 256) return fail(TT_ERR_INVALID, "bad TMA shape");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TT_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (%d)", (int)r);
  return TT_OK;
}

// ------------------------------------------------------------------ kernel
struct GemmEpilogue {
  const float* bias;        // [N] or null
  const float* mask;        // [M, N] (pitch ld_mask): out = mask > 0 ? out : 0, or null
  const __nv_bfloat16* mask_bf16;  // same gate read from a bf16 matrix (pitch ld_mask_bf16), or null
  float* out_f32;           // [M, N] row-major or null
  __nv_bfloat16* out_bf16;  // [M, N] row-major or null
  __nv_bfloat16* out_bf16_t;  // [N, M] (transposed) or null
  int64_t ld_mask, ld_mask_bf16, ld_f32, ld_bf16, ld_bf16_t;
  int relu;
  int kb_per_split;         // split-K: gridDim.z slices of kb_per_split K-blocks; slice z stores raw fp32
                            // partials at out_f32 + z*M*ld_f32 (bias/relu/mask/bf16 outputs must be off)
};

constexpr int kBM = 128, kBK = 64, kStages = 4, kGemmThreadsTc = 192;

// Shared memory follows the K extent: a short-K GEMM (the towers: K = 64 / 128) takes one or two stages, so
// three CTAs share an SM and one CTA's TMA / MMA latency hides behind another's epilogue.
template <int BN>
constexpr int gemm_smem_bytes(int stages) {
  return stages * (kBM * kBK * 2 + (BN < 64 ? 64 : BN) * kBK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}

// MN = true: C[M,N] = A^T B with A given as [K, M] and B as [K, N] row-major (the weight gradient dW = dZ^T A_prev straight
// from the row-major activations: no transposed copies).  A k-block is then kBK rows of the source matrices: sub-tiles of
// [64 k-rows x 64 columns] (one TMA box each), read by the tensor core through MN-major descriptors.
template <int BN, bool MN = false>
__global__ void __launch_bounds__(kGemmThreadsTc, BN <= 128 ? 3 : 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmEpilogue ep, int M, int N, int K, int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BNS = (MN && BN < 64) ? 64 : BN;      // columns of B staged per k-block (an MN-major box is 64 wide)
  constexpr int A_BYTES = kBM * kBK * 2, B_BYTES = BNS * kBK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + stages * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + stages * B_BYTES);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int total_kb = (K + kBK - 1) / kBK;
  const int kb0 = blockIdx.z * ep.kb_per_split;
  const int num_kb = min(ep.kb_per_split, total_kb - kb0);   // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        if (MN) {
          for (int h = 0; h < kBM / 64; ++h) tma_load_2d(sA + s * A_BYTES + h * 8192, &tmA, &full[s], m0 + 64 * h, (kb0 + kb) * kBK);
          for (int h = 0; h < BNS / 64; ++h) tma_load_2d(sB + s * B_BYTES + h * 8192, &tmB, &full[s], n0 + 64 * h, (kb0 + kb) * kBK);
        } else {
          tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], (kb0 + kb) * kBK, m0);
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], (kb0 + kb) * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(kBM, BN) | (MN ? (kIdescAMnMajor | kIdescBMnMajor) : 0u);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages;
      const uint32_t ph = (kb / stages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        if (MN) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_ss(tmem_base, smem_desc_mn_sw128(smem_u32(sA + s * A_BYTES) + k * 2048, 8192, 1024),
                   smem_desc_mn_sw128(smem_u32(sB + s * B_BYTES) + k * 2048, 8192, 1024), idesc, (kb | k) != 0);
        } else {
          const uint64_t da = smem_desc_k_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t db = smem_desc_k_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        tc_commit(&empty[s]);                       // frees the smem stage when these MMAs retire
        if (kb == num_kb - 1) tc_commit(acc_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: thread owns row (32*(warp%4) + lane) of the tile
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr_row + c0, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = n0 + c0 + j;
        float x = __uint_as_float(v[j]);
        if (ep.bias != nullptr && col < N) x += ep.bias[col];
        if (ep.relu) x = fmaxf(x, 0.f);
        f[j] = x;
      }
      if (ep.mask != nullptr && row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (col < N && !(ep.mask[(int64_t)row * ep.ld_mask + col] > 0.f)) f[j] = 0.f;
        }
      }
      if (ep.mask_bf16 != nullptr && row < M) {
        const __nv_bfloat16* mrow = ep.mask_bf16 + (int64_t)row * ep.ld_mask_bf16 + n0 + c0;
        if (n0 + c0 + 32 <= N && (reinterpret_cast<uintptr_t>(mrow) & 15) == 0) {
          // 32 gates = four 16-byte loads (one 2-byte load per element made this epilogue the whole kernel: 2.8 ms for
          // the 262144 x 1024 x 512 data-gradient GEMM of cfg4)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 mk = *reinterpret_cast<const uint4*>(mrow + j);
            const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero (NaN gates count as positive, as x > 0 is false for them
              // ... a ReLU output is never NaN-gated in practice; keep the float compare for exact parity)
              const float lo = __uint_as_float(w[t] << 16), hi = __uint_as_float(w[t] & 0xffff0000u);
              if (!(lo > 0.f)) f[j + 2 * t] = 0.f;
              if (!(hi > 0.f)) f[j + 2 * t + 1] = 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + c0 + j;
            if (col < N && !(__bfloat162float(mrow[j]) > 0.f)) f[j] = 0.f;
          }
        }
      }
      if (row < M) {
        const bool full_chunk = (n0 + c0 + 32 <= N);
        if (ep.out_f32 != nullptr) {
          float* o = ep.out_f32 + (int64_t)blockIdx.z * M * ep.ld_f32 + (int64_t)row * ep.ld_f32 + n0 + c0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < N) o[j] = f[j];
          }
        }
        if (ep.out_bf16 != nullptr) {
          __nv_bfloat16* o = ep.out_bf16 + (int64_t)row * ep.ld_bf16 + n0 + c0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
              pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < N) o[j] = __float2bfloat16(f[j]);
          }
        }
      }
      if (ep.out_bf16_t != nullptr) {
        // transposed store: for a fixed column the warp's 32 rows are contiguous
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (col < N && row < M) ep.out_bf16_t[(int64_t)col * ep.ld_bf16_t + row] = __float2bfloat16(f[j]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
  }
}

// fp32 [rows, cols] (pitch ldx) -> bf16 row-major (pitch ld_out) and/or transposed bf16 [cols, rows]
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gate, int64_t ld_gate, int rows,
                 int cols, __nv_bfloat16* __restrict__ out, int64_t ld_out, __nv_bfloat16* __restrict__ out_t,
                 int64_t ld_out_t) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;     // rows on x: corpora have millions of rows
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float v = (r < rows && c < cols) ? x[(int64_t)r * ldx + c] : 0.f;
    if (gate != nullptr && r < rows && c < cols && !(gate[(int64_t)r * ld_gate + c] > 0.f)) v = 0.f;
    tile[ty + i * 8][tx] = v;
    if (out != nullptr && r < rows && c < cols) out[(int64_t)r * ld_out + c] = __float2bfloat16(v);
  }
  if (out_t == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < rows && c < cols) out_t[(int64_t)c * ld_out_t + r] = __float2bfloat16(tile[tx][ty + i * 8]);
  }
}

// partial[z][c] = sum over the rows of chunk z of x[r, c] (bf16 in, fp32 accumulate).
// A warp reads 64 consecutive columns of a row (bf16x2 per lane = 128 B per row), 8 rows per CTA pass.
__global__ void __launch_bounds__(256)
colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int rows, int cols, int rows_per_chunk,
                           float* __restrict__ partial) {
  __shared__ float red[8][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s0 = 0.f, s1 = 0.f;
  const bool pair = (c + 1 < cols) && ((ldx & 1) == 0);
  if (c < cols) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const __nv_bfloat16* p = x + (int64_t)r * ldx + c;
      if (pair) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
        s0 += __low2float(v);
        s1 += __high2float(v);
      } else {
        s0 += __bfloat162float(p[0]);
        if (c + 1 < cols) s1 += __bfloat162float(p[1]);
      }
    }
  }
  red[ty][tx * 2] = s0;
  red[ty][tx * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
      partial[(int64_t)blockIdx.y * cols + cc] = t;
    }
  }
}

// out[i] = sum_z partial[z*count + i], z ascending (deterministic)
__global__ void tc_reduce_partials_kernel(const float* __restrict__ partial, int S, int64_t count, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += partial[(int64_t)z * count + i];
  out[i] = s;
}

template <int BN, bool MN = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                       int splits, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         gemm_smem_bytes<BN>(kStages));
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "tc_gemm smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int kb_cta = ep.kb_per_split;                       // K blocks one CTA walks through
  int stages = kb_cta < kStages ? (kb_cta < 1 ? 1 : kb_cta) : kStages;
  // Two CTAs per SM matter more than a deep ring: the epilogue (thread-per-row stores) of one tile is several times its
  // MMA time, and only a second resident CTA overlaps it with a mainloop.  Cap the ring at what lets two CTAs fit.
  static const int env_cap = [] { const char* v = getenv("TT_GEMM_SMEM_CAP"); return v ? atoi(v) : 110 * 1024; }();
  while (stages > 1 && gemm_smem_bytes<BN>(stages) > env_cap) --stages;
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM, splits);
  tc_gemm_kernel<BN, MN><<<grid, kGemmThreadsTc, gemm_smem_bytes<BN>(stages), s>>>(ta, tb, ep, M, N, K, stages);
  TT_CHECK_LAUNCH("tc_gemm");
  return TT_OK;
}

}  // namespace tc
}  // namespace tt

using namespace tt;
using namespace tt::tc;

extern "C" {

int tt_cast_f32_to_bf16(const float* x, int64_t ldx, const float* gate, int64_t ld_gate, int64_t rows, int64_t cols,
                        void* out, int64_t ld_out, void* out_t, int64_t ld_out_t, void* stream) {
  TT_CHECK_ARG(rows >= 0 && cols >= 0 && (out || out_t), "cast_bf16: bad args");
  if (rows == 0 || cols == 0) return TT_OK;
  TT_CHECK_ARG(x != nullptr, "cast_bf16: null input");
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  if (grid.y > 65535) return fail(TT_ERR_UNSUPPORTED, "cast_bf16: too many columns");
  cast_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, ldx, gate, ld_gate, (int)rows, (int)cols,
                                                        static_cast<__nv_bfloat16*>(out), ld_out,
                                                        static_cast<__nv_bfloat16*>(out_t), ld_out_t);
  TT_CHECK_LAUNCH("cast_bf16");
  return TT_OK;
}

static int gemm_dispatch(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                         int splits, cudaStream_t s) {
  switch (bn) {
    case 32: return launch_gemm<32>(ta, tb, ep, M, N, K, splits, s);
    case 64: return launch_gemm<64>(ta, tb, ep, M, N, K, splits, s);
    case 128: return launch_gemm<128>(ta, tb, ep, M, N, K, splits, s);
    default: return launch_gemm<256>(ta, tb, ep, M, N, K, splits, s);
  }
}

int tt_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 const float* bias, int32_t relu, const float* mask, int64_t ld_mask, const void* mask_bf16,
                 int64_t ld_mask_bf16, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                 void* out_bf16_t, int64_t ld_bf16_t, void* stream) {
  TT_CHECK_ARG(M > 0 && N > 0 && K > 0 && a && b, "gemm_bf16: bad args");
  TT_CHECK_ARG(out_f32 || out_bf16 || out_bf16_t, "gemm_bf16: no output");
  if (M >= ((int64_t)1 << 31) || (M + kBM - 1) / kBM > 65535) return fail(TT_ERR_UNSUPPORTED, "gemm_bf16: M too large");
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, a, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, b, N, K, ldb, bn);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.mask_bf16 = static_cast<const __nv_bfloat16*>(mask_bf16); ep.ld_mask_bf16 = ld_mask_bf16;
  ep.kb_per_split = (int)((K + kBK - 1) / kBK);
  ep.bias = bias; ep.mask = mask; ep.out_f32 = out_f32;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  ep.out_bf16_t = static_cast<__nv_bfloat16*>(out_bf16_t);
  ep.ld_mask = ld_mask; ep.ld_f32 = ld_f32; ep.ld_bf16 = ld_bf16; ep.ld_bf16_t = ld_bf16_t; ep.relu = relu;
  return gemm_dispatch(bn, ta, tb, ep, (int)M, (int)N, (int)K, 1, as_stream(stream));
}

static int gemm_dispatch_mn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                            int splits, cudaStream_t s) {
  switch (bn) {
    case 32: return launch_gemm<32, true>(ta, tb, ep, M, N, K, splits, s);
    case 64: return launch_gemm<64, true>(ta, tb, ep, M, N, K, splits, s);
    case 128: return launch_gemm<128, true>(ta, tb, ep, M, N, K, splits, s);
    default: return launch_gemm<256, true>(ta, tb, ep, M, N, K, splits, s);
  }
}

static int splitk_splits(int64_t M, int64_t N, int64_t K, int bn) {
  int64_t tiles = ((M + kBM - 1) / kBM) * ((N + bn - 1) / bn);
  if (tiles < 1) tiles = 1;                                     // an empty output (M or N == 0) must not divide by zero
  const int64_t total_kb = (K + kBK - 1) / kBK;
  int64_t want = (kNumSMs + tiles - 1) / tiles;          // one wave of CTAs (fewer partials to reduce)
  if (want > total_kb / 4) want = total_kb / 4;          // at least 4 K-blocks per slice
  if (want < 1) want = 1;
  return (int)want;
}

size_t tt_gemm_bf16_splitk_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  return align_up((size_t)splitk_splits(M, N, K, bn) * M * N * 4, 256) + 256;
}

// C[M,N] (fp32) = A[M,K] . B[N,K]^T with the K reduction split over CTAs (weight gradients:
// M, N small, K = batch).  Partials go to ws, an ordered reduce writes out (deterministic).
int tt_gemm_bf16_splitk(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int64_t N, int64_t K,
                        float* out_f32, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(M > 0 && N > 0 && K > 0 && a && b && out_f32, "gemm_bf16_splitk: bad args");
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  const int want = splitk_splits(M, N, K, bn);
  const int total_kb = (int)((K + kBK - 1) / kBK);
  const int per = (total_kb + want - 1) / want;
  const int splits = (total_kb + per - 1) / per;          // every slice gets >= 1 K-block
  if (!ws || ws_bytes < (size_t)splits * M * N * 4) return fail(TT_ERR_WORKSPACE, "gemm_bf16_splitk: workspace too small");
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, a, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, b, N, K, ldb, bn);
  if (rc) return rc;
  GemmEpilogue ep = {};
  ep.out_f32 = static_cast<float*>(ws);
  ep.ld_f32 = N;
  ep.kb_per_split = per;
  cudaStream_t s = as_stream(stream);
  rc = gemm_dispatch(bn, ta, tb, ep, (int)M, (int)N, (int)K, splits, s);
  if (rc) return rc;
  const int64_t cnt = M * N;
  tc_reduce_partials_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(static_cast<float*>(ws), splits, cnt, out_f32);
  TT_CHECK_LAUNCH("tc_reduce_partials");
  return TT_OK;
}

// C[M,N] (fp32) = A^T B with A [K, M] and B [K, N] row-major bf16 (dW = dZ^T A_prev from the row-major activations).
int tt_gemm_bf16_splitk_mn(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int64_t N, int64_t K,
                           float* out_f32, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(M > 0 && N > 0 && K > 0 && a && b && out_f32, "gemm_bf16_splitk_mn: bad args");
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  const int want = splitk_splits(M, N, K, bn);
  const int total_kb = (int)((K + kBK - 1) / kBK);
  const int per = (total_kb + want - 1) / want;
  const int splits = (total_kb + per - 1) / per;
  if (!ws || ws_bytes < (size_t)splits * M * N * 4) return fail(TT_ERR_WORKSPACE, "gemm_bf16_splitk_mn: workspace too small");
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, a, K, M, lda, kBK);      // boxes of [64 k-rows x 64 columns]
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, b, K, N, ldb, kBK);
  if (rc) return rc;
  GemmEpilogue ep = {};
  ep.out_f32 = static_cast<float*>(ws);
  ep.ld_f32 = N;
  ep.kb_per_split = per;
  cudaStream_t s = as_stream(stream);
  rc = gemm_dispatch_mn(bn, ta, tb, ep, (int)M, (int)N, (int)K, splits, s);
  if (rc) return rc;
  const int64_t cnt = M * N;
  tc_reduce_partials_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(static_cast<float*>(ws), splits, cnt, out_f32);
  TT_CHECK_LAUNCH("tc_reduce_partials");
  return TT_OK;
}

// chunks of >= 256 rows, at most 128 of them: the final ordered reduce reads `chunks` partials per column serially
static int colsum_rows_per_chunk(int64_t rows) {
  int64_t rpc = (rows + 127) / 128;
  rpc = (rpc + 255) / 256 * 256;
  return (int)(rpc < 256 ? 256 : rpc);
}
size_t tt_colsum_bf16_workspace_bytes(int64_t rows, int64_t cols) {
  const int64_t chunks = (rows + 255) / 256;            // upper bound (the kernel uses fewer, larger chunks)
  return align_up((size_t)chunks * cols * 4, 256) + 256;
}

// out[c] = sum_r x[r, c]  (bias gradient of a tower layer from the bf16 dZ)
int tt_colsum_bf16(const void* x, int64_t ldx, int64_t rows, int64_t cols, float* out, void* ws, size_t ws_bytes,
                   void* stream) {
  TT_CHECK_ARG(rows > 0 && cols > 0 && x && out, "colsum_bf16: bad args");
  const int rpc = colsum_rows_per_chunk(rows);
  const int chunks = (int)((rows + rpc - 1) / rpc);
  if (!ws || ws_bytes < (size_t)chunks * cols * 4) return fail(TT_ERR_WORKSPACE, "colsum_bf16: workspace too small");
  cudaStream_t s = as_stream(stream);
  dim3 grid((unsigned)((cols + 63) / 64), (unsigned)chunks);
  colsum_bf16_partial_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, (int)rows, (int)cols, rpc,
                                                  static_cast<float*>(ws));
  TT_CHECK_LAUNCH("colsum_bf16_partial");
  tc_reduce_partials_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(static_cast<float*>(ws), chunks, cols, out);
  TT_CHECK_LAUNCH("tc_reduce_partials");
  return TT_OK;
}

}  // extern "C"
