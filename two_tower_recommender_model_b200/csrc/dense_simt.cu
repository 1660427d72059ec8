// fp32 CUDA-core dense kernels: the exact-fp32 tower path (any shape), the
// reference's row-wise-dot BCE loss, an fp32 in-batch softmax that never writes
// the [B,B] logits, and flat Adam.  The tensor-core (tcgen05) versions of the
// GEMM-shaped ops live in gemm_tcgen05.cu; these kernels serve fp32-exact parity
// and shapes the tensor-core path does not take (dims not a multiple of 16).
#include "common.cuh"

namespace tt {

constexpr int kTile = 64;
constexpr int kBK = 16;
constexpr int kPad = 68;  // 64 + 4: keeps float4 alignment, breaks bank regularity
constexpr int kGemmThreads = 256;

// acc[i][j] += sum_{k in [k0,k1)} A(m0 + ty*4+i, k) * B(n0 + tx*4+j, k)
// loadA(r, k) / loadB(r, k) return 0 outside bounds.  KFAST_x: k is the
// contiguous index of that operand in memory (picks the coalesced load order).
template <bool KFAST_A, bool KFAST_B, class LA, class LB>
__device__ __forceinline__ void tile_mma(LA loadA, LB loadB, int k0, int k1, float (&acc)[4][4],
                                         float (*As)[kPad], float (*Bs)[kPad]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  for (int kk = k0; kk < k1; kk += kBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * kGemmThreads;
      {
        const int r = KFAST_A ? idx >> 4 : idx & 63;
        const int c = KFAST_A ? idx & 15 : idx >> 6;
        As[c][r] = (kk + c < k1) ? loadA(r, kk + c) : 0.f;
      }
      {
        const int r = KFAST_B ? idx >> 4 : idx & 63;
        const int c = KFAST_B ? idx & 15 : idx >> 6;
        Bs[c][r] = (kk + c < k1) ? loadB(r, kk + c) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// C[m, n] = epilogue( sum_k A(m,k) * B(n,k) ), with
//   A(m,k) = A_T ? A[k*lda+m] : A[m*lda+k]   (times (Amask>0) when Amask != null)
//   B(n,k) = B_T ? B[k*ldb+n] : B[n*ldb+k]
// gridDim.z > 1: split-K, slice z writes its partial tile to C + z*M*ldc.
template <bool A_T, bool B_T>
__global__ void __launch_bounds__(kGemmThreads)
simt_gemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Amask,
                 const float* __restrict__ B, int64_t ldb, const float* __restrict__ bias,
                 float* __restrict__ C, int64_t ldc, int M, int N, int K, int relu, int k_chunk) {
  __shared__ __align__(16) float As[kBK][kPad];
  __shared__ __align__(16) float Bs[kBK][kPad];
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int k0 = blockIdx.z * k_chunk;
  const int k1 = min(K, k0 + k_chunk);
  float acc[4][4] = {};
  auto la = [&](int r, int k) -> float {
    const int m = m0 + r;
    if (m >= M) return 0.f;
    const int64_t o = A_T ? (int64_t)k * lda + m : (int64_t)m * lda + k;
    float v = A[o];
    if (Amask != nullptr && !(Amask[o] > 0.f)) v = 0.f;
    return v;
  };
  auto lb = [&](int r, int k) -> float {
    const int n = n0 + r;
    if (n >= N) return 0.f;
    return B_T ? B[(int64_t)k * ldb + n] : B[(int64_t)n * ldb + k];
  };
  tile_mma<!A_T, !B_T>(la, lb, k0, k1, acc, As, Bs);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* Cz = C + (int64_t)blockIdx.z * M * ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      Cz[(int64_t)m * ldc + n] = v;
    }
  }
}

// out[i] = sum_s partial[s*count + i], s ascending (deterministic).
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int S, int64_t count,
                                       float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += partial[(int64_t)z * count + i];
  out[i] = s;
}

// partial[z][n] = sum over rows of chunk z of dy[m,n] * (y[m,n] > 0)
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const float* __restrict__ dy, const float* __restrict__ y, int M, int N,
                         int relu, int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  if (n < N)
    for (int m = r0 + ty; m < r1; m += 8) {
      float v = dy[(int64_t)m * N + n];
      if (relu && !(y[(int64_t)m * N + n] > 0.f)) v = 0.f;
      s += v;
    }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    partial[(int64_t)blockIdx.y * N + n] = t;
  }
}

// ---------------------------------------------------------------------------
// row-wise dot + BCEWithLogits (utils/model_training.py:136-140), fwd + bwd
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dot_bce_kernel(const float* __restrict__ q, const float* __restrict__ c, const int32_t* __restrict__ labels,
               int B, int d, float* __restrict__ logits, float* __restrict__ partial_loss,
               float* __restrict__ dq, float* __restrict__ dc, float gscale) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + warp;
  float li = 0.f;
  if (b < B) {
    const float* qb = q + (int64_t)b * d;
    const float* cb = c + (int64_t)b * d;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) s = fmaf(qb[k], cb[k], s);
    s = warp_sum(s);
    const float y = (float)labels[b];
    li = fmaxf(s, 0.f) - s * y + log1pf(expf(-fabsf(s)));
    if (lane == 0) logits[b] = s;
    if (dq != nullptr) {
      const float sig = 1.f / (1.f + expf(-s));
      const float g = (sig - y) * gscale;
      for (int k = lane; k < d; k += 32) {
        const float qv = qb[k], cv = cb[k];
        dq[(int64_t)b * d + k] = g * cv;
        dc[(int64_t)b * d + k] = g * qv;
      }
    }
  }
  if (lane == 0) red[warp] = li;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partial_loss[blockIdx.x] = t;
  }
}

// loss = scale * sum(partial) with one block, ordered tree (deterministic)
__global__ void __launch_bounds__(1024)
final_sum_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = red[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = t * scale;
  }
}

// ---------------------------------------------------------------------------
// in-batch softmax, fp32 CUDA cores.  One CTA owns 64 rows of X and streams the
// 64-row tiles of Y; the 64x64 logits tile lives in registers only.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads)
inbatch_softmax_fwd_kernel(const float* __restrict__ q, const float* __restrict__ c, int B, int d,
                           float inv_t, float* __restrict__ lse, float* __restrict__ diag,
                           float* __restrict__ partial_loss) {
  __shared__ __align__(16) float As[kBK][kPad];
  __shared__ __align__(16) float Bs[kBK][kPad];
  __shared__ float wred[8];
  const int m0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float rmax[4], rsum[4], rdiag[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { rmax[i] = -INFINITY; rsum[i] = 0.f; rdiag[i] = 0.f; }
  auto la = [&](int r, int k) -> float { return (m0 + r < B) ? q[(int64_t)(m0 + r) * d + k] : 0.f; };
  for (int n0 = 0; n0 < B; n0 += kTile) {
    float acc[4][4] = {};
    auto lb = [&](int r, int k) -> float { return (n0 + r < B) ? c[(int64_t)(n0 + r) * d + k] : 0.f; };
    tile_mma<true, true>(la, lb, 0, d, acc, As, Bs);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        acc[i][j] = (n < B) ? acc[i][j] * inv_t : -INFINITY;
        tmax = fmaxf(tmax, acc[i][j]);
        if (n == m) rdiag[i] = acc[i][j];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float nmax = fmaxf(rmax[i], tmax);
      float ts = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) ts += expf(acc[i][j] - nmax);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
      rsum[i] = rsum[i] * expf(rmax[i] - nmax) + ts;
      rmax[i] = nmax;
    }
  }
  // the diagonal element sits in exactly one of the 16 lanes that share a row
  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float dg = rdiag[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o);
    const int m = m0 + ty * 4 + i;
    if (tx == 0 && m < B) {
      const float l = rmax[i] + logf(rsum[i]);
      lse[m] = l;
      diag[m] = dg;
      lsum += l - dg;
    }
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wred[i];
    partial_loss[blockIdx.x] = t;
  }
}

// out[r,:] = scale * ( sum_t exp(x_r.y_t*inv_t - lse[ROW ? r : t]) * y_t  -  y_r )
// NO = ceil(d/64): number of 64-wide output column chunks held in registers.
template <int NO, bool ROW>
__global__ void __launch_bounds__(kGemmThreads)
inbatch_softmax_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                           const float* __restrict__ lse, int B, int d, float inv_t, float scale,
                           float* __restrict__ out) {
  __shared__ __align__(16) float As[kBK][kPad];
  __shared__ __align__(16) float Bs[kBK][kPad];
  __shared__ __align__(16) float Ps[kTile][kPad];   // P tile  [row][t]
  __shared__ __align__(16) float Ys[kTile][kPad];   // Y chunk [t][dcol]
  const int m0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float o[NO][4][4] = {};
  float lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) lrow[i] = (ROW && m0 + ty * 4 + i < B) ? lse[m0 + ty * 4 + i] : 0.f;
  auto la = [&](int r, int k) -> float { return (m0 + r < B) ? x[(int64_t)(m0 + r) * d + k] : 0.f; };
  for (int n0 = 0; n0 < B; n0 += kTile) {
    float acc[4][4] = {};
    auto lb = [&](int r, int k) -> float { return (n0 + r < B) ? y[(int64_t)(n0 + r) * d + k] : 0.f; };
    tile_mma<true, true>(la, lb, 0, d, acc, As, Bs);
    float lcol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lcol[j] = (!ROW && n0 + tx * 4 + j < B) ? lse[n0 + tx * 4 + j] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (m0 + ty * 4 + i < B) && (n0 + tx * 4 + j < B);
        Ps[ty * 4 + i][tx * 4 + j] = ok ? expf(acc[i][j] * inv_t - (ROW ? lrow[i] : lcol[j])) : 0.f;
      }
#pragma unroll
    for (int ch = 0; ch < NO; ++ch) {
      __syncthreads();  // Ps complete (ch==0) / previous Ys consumed
      for (int idx = threadIdx.x; idx < kTile * kTile; idx += kGemmThreads) {
        const int t = idx >> 6, dc = idx & 63;
        const int col = ch * 64 + dc;
        Ys[t][dc] = (n0 + t < B && col < d) ? y[(int64_t)(n0 + t) * d + col] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int t = 0; t < kTile; ++t) {
        const float4 b = *reinterpret_cast<const float4*>(&Ys[t][tx * 4]);
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = Ps[ty * 4 + i][t];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[ch][i][j] = fmaf(a, bv[j], o[ch][i][j]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ch = 0; ch < NO; ++ch)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= B) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = ch * 64 + tx * 4 + j;
        if (col < d) out[(int64_t)m * d + col] = scale * (o[ch][i][j] - y[(int64_t)m * d + col]);
      }
    }
}

__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                 float bc1, float bc2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// Adam with the step counter on the device: the launch arguments stay constant from step to step,
// so the call can be captured in a CUDA graph.  step_inc_kernel runs first on the same stream.
__global__ void step_inc_kernel(float* step) { *step += 1.0f; }

__global__ void adam_flat_devstep_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                         float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                         const float* __restrict__ step) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = *step;
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}

static int launch_gemm(bool a_t, bool b_t, const float* A, int64_t lda, const float* Amask, const float* B,
                       int64_t ldb, const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu,
                       int splits, int k_chunk, cudaStream_t s) {
  dim3 grid((N + kTile - 1) / kTile, (M + kTile - 1) / kTile, splits);
  if (grid.y > 65535) return fail(TT_ERR_UNSUPPORTED, "simt_gemm: M too large (%d)", M);
#define TT_G(AT, BT) \
  simt_gemm_kernel<AT, BT><<<grid, kGemmThreads, 0, s>>>(A, lda, Amask, B, ldb, bias, C, ldc, M, N, K, relu, k_chunk)
  if (!a_t && !b_t) TT_G(false, false);
  else if (!a_t && b_t) TT_G(false, true);
  else if (a_t && !b_t) TT_G(true, false);
  else TT_G(true, true);
#undef TT_G
  TT_CHECK_LAUNCH("simt_gemm");
  return TT_OK;
}

static int wgrad_splits(int64_t M, int64_t N, int64_t K) {
  int64_t s = (M + 511) / 512;
  if (s > 128) s = 128;
  int64_t cap = ((int64_t)64 << 20) / (N * K * 4 > 0 ? N * K * 4 : 1);
  if (cap < 1) cap = 1;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_linear_forward_f32(const float* x, int64_t ldx, const float* w, const float* bias, float* y,
                          int64_t M, int64_t N, int64_t K, int32_t relu, void* stream) {
  TT_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldx >= K, "linear_forward: bad shape");
  if (M == 0) return TT_OK;
  TT_CHECK_ARG(x && w && y, "linear_forward: null pointer");
  TT_CHECK_ARG(M < ((int64_t)1 << 22), "linear_forward: M too large for the fp32 path");
  return launch_gemm(false, false, x, ldx, nullptr, w, K, bias, y, N, (int)M, (int)N, (int)K, relu, 1, (int)K,
                     as_stream(stream));
}

size_t tt_linear_backward_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int s = wgrad_splits(M, N, K);
  return align_up((size_t)s * N * K * 4, 256) + align_up((size_t)s * N * 4, 256) + 512;
}

int tt_linear_backward_f32(const float* x, int64_t ldx, const float* w, const float* y, const float* dy,
                           float* dx, int64_t lddx, float* dw, float* db, int64_t M, int64_t N, int64_t K,
                           int32_t relu, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear_backward: bad shape");
  TT_CHECK_ARG(x && w && dy && (!relu || y), "linear_backward: null pointer");
  TT_CHECK_ARG(M < ((int64_t)1 << 22), "linear_backward: M too large for the fp32 path");
  cudaStream_t s = as_stream(stream);
  const float* mask = relu ? y : nullptr;
  if (M == 0) {
    if (dw) cudaMemsetAsync(dw, 0, (size_t)N * K * 4, s);
    if (db) cudaMemsetAsync(db, 0, (size_t)N * 4, s);
    return TT_OK;
  }
  int rc;
  if (dx) {  // dx[m,k] = sum_n dz[m,n] * w[n,k]  ->  A = dz (not T), B(k, n) = w[n*K+k] (T)
    rc = launch_gemm(false, true, dy, N, mask, w, K, nullptr, dx, lddx, (int)M, (int)K, (int)N, 0, 1, (int)N, s);
    if (rc) return rc;
  }
  const int S = wgrad_splits(M, N, K);
  const int chunk = (int)((M + S - 1) / S);
  Workspace wk(ws, ws_bytes);
  float* pw = wk.take<float>((size_t)S * N * K);
  float* pb = wk.take<float>((size_t)S * N);
  if (!pw || !pb) return fail(TT_ERR_WORKSPACE, "linear_backward: workspace too small");
  if (dw) {  // dw[n,k] = sum_m dz[m,n] * x[m,k] -> A(n, m) = dz[m*N+n] (T), B(k, m) = x[m*ldx+k] (T)
    rc = launch_gemm(true, true, dy, N, mask, x, ldx, nullptr, pw, K, (int)N, (int)K, (int)M, 0, S, chunk, s);
    if (rc) return rc;
    int64_t cnt = N * K;
    reduce_partials_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(pw, S, cnt, dw);
    TT_CHECK_LAUNCH("reduce_partials(dw)");
  }
  if (db) {
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)S);
    bias_grad_partial_kernel<<<grid, 256, 0, s>>>(dy, y, (int)M, (int)N, relu, chunk, pb);
    TT_CHECK_LAUNCH("bias_grad_partial");
    reduce_partials_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(pb, S, N, db);
    TT_CHECK_LAUNCH("reduce_partials(db)");
  }
  return TT_OK;
}

size_t tt_dot_bce_workspace_bytes(int64_t B) { return align_up((size_t)((B + 7) / 8 + 1) * 4, 256); }

int tt_dot_bce(const float* q, const float* c, const int32_t* labels, int64_t B, int64_t d, float* logits,
               float* loss, float* dq, float* dc, float grad_scale, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(B > 0 && d > 0 && q && c && labels && logits && loss, "dot_bce: bad args");
  TT_CHECK_ARG((dq == nullptr) == (dc == nullptr), "dot_bce: dq and dc must both be set or both be null");
  const int blocks = (int)((B + 7) / 8);
  if (ws_bytes < (size_t)blocks * 4 || !ws) return fail(TT_ERR_WORKSPACE, "dot_bce: workspace too small");
  cudaStream_t s = as_stream(stream);
  float* partial = static_cast<float*>(ws);
  dot_bce_kernel<<<blocks, 256, 0, s>>>(q, c, labels, (int)B, (int)d, logits, partial, dq, dc,
                                        grad_scale / (float)B);
  TT_CHECK_LAUNCH("dot_bce");
  final_sum_kernel<<<1, 1024, 0, s>>>(partial, blocks, 1.0f / (float)B, loss);
  TT_CHECK_LAUNCH("final_sum");
  return TT_OK;
}

size_t tt_inbatch_softmax_workspace_bytes(int64_t B) { return align_up((size_t)((B + 63) / 64 + 1) * 4, 256); }

int tt_inbatch_softmax_forward_f32(const float* q, const float* c, int64_t B, int64_t d, float inv_t,
                                   float* lse, float* diag, float* loss, void* ws, size_t ws_bytes,
                                   void* stream) {
  TT_CHECK_ARG(B > 0 && d > 0 && q && c && lse && diag && loss, "inbatch_softmax_forward: bad args");
  const int blocks = (int)((B + 63) / 64);
  if (!ws || ws_bytes < (size_t)blocks * 4) return fail(TT_ERR_WORKSPACE, "inbatch_softmax: workspace too small");
  cudaStream_t s = as_stream(stream);
  float* partial = static_cast<float*>(ws);
  inbatch_softmax_fwd_kernel<<<blocks, kGemmThreads, 0, s>>>(q, c, (int)B, (int)d, inv_t, lse, diag, partial);
  TT_CHECK_LAUNCH("inbatch_softmax_fwd");
  final_sum_kernel<<<1, 1024, 0, s>>>(partial, blocks, 1.0f / (float)B, loss);
  TT_CHECK_LAUNCH("final_sum");
  return TT_OK;
}

int tt_inbatch_softmax_backward_f32(const float* q, const float* c, const float* lse, int64_t B, int64_t d,
                                    float inv_t, float grad_scale, float* dq, float* dc, void* stream) {
  TT_CHECK_ARG(B > 0 && d > 0 && q && c && lse && dq && dc, "inbatch_softmax_backward: bad args");
  if (d > 256) return fail(TT_ERR_UNSUPPORTED, "inbatch_softmax_backward_f32: d > 256");
  cudaStream_t s = as_stream(stream);
  const int blocks = (int)((B + 63) / 64);
  const float scale = grad_scale * inv_t / (float)B;
#define TT_SB(NO)                                                                                              \
  do {                                                                                                         \
    inbatch_softmax_bwd_kernel<NO, true><<<blocks, kGemmThreads, 0, s>>>(q, c, lse, (int)B, (int)d, inv_t, scale, dq);  \
    inbatch_softmax_bwd_kernel<NO, false><<<blocks, kGemmThreads, 0, s>>>(c, q, lse, (int)B, (int)d, inv_t, scale, dc); \
  } while (0)
  if (d <= 64) TT_SB(1);
  else if (d <= 128) TT_SB(2);
  else TT_SB(4);
#undef TT_SB
  ++g_kernel_launches;  // two launches above, one check below
  TT_CHECK_LAUNCH("inbatch_softmax_bwd");
  return TT_OK;
}

int tt_adam_flat_devstep(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, float* step, void* stream) {
  TT_CHECK_ARG(n >= 0 && step, "adam_flat_devstep: bad args");
  cudaStream_t s = as_stream(stream);
  step_inc_kernel<<<1, 1, 0, s>>>(step);
  TT_CHECK_LAUNCH("step_inc");
  if (n == 0) return TT_OK;
  TT_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_flat_devstep: null pointer");
  adam_flat_devstep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                     beta2, eps, step);
  TT_CHECK_LAUNCH("adam_flat_devstep");
  return TT_OK;
}

int tt_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                 float beta1, float beta2, float eps, float bc1, float bc2, void* stream) {
  TT_CHECK_ARG(n >= 0, "adam_flat: negative n");
  if (n == 0) return TT_OK;
  TT_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_flat: null pointer");
  adam_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                             lr, beta1, beta2, eps, bc1, bc2);
  TT_CHECK_LAUNCH("adam_flat");
  return TT_OK;
}

}  // extern "C"
