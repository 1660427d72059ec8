// Retrieval: exact dot-product top-k (04_evaluate_retrieval.py:134-141, k=100).
// fp32 CUDA-core scoring with the per-block top-k fused into the scoring CTA:
// the [Q, N] score matrix is never written.  A CTA owns 64 queries and one
// contiguous range of items, walks the range in 64-item tiles, and keeps a sorted
// top-k list per query in shared memory; candidates enter only when they beat
// the list's current k-th score.  A second kernel merges the per-range lists.
// Order: descending score, ties -> lower item index (strict ">" on entry keeps it
// because items are visited in ascending index order).
#include "common.cuh"

namespace tt {

constexpr int kTkTile = 64;
constexpr int kTkBK = 16;
constexpr int kTkPad = 68;
constexpr int kTkThreads = 256;
constexpr int kTkMaxK = 128;

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Insert (s, id) into the warp-shared sorted list (score desc; existing entries
// always carry lower ids than the newcomer).  Returns the new count.
__device__ __forceinline__ int warp_insert(float* ls, int* li, int cnt, int k, float s, int id, int lane) {
  int c = 0;
  for (int e = lane; e < cnt; e += 32) c += (ls[e] >= s) ? 1 : 0;
  const int pos = warp_sum_int(c);
  if (pos >= k) return cnt;
  const int last = min(cnt, k - 1);  // entries [pos, last) move one slot right
  float ts[kTkMaxK / 32];
  int ti[kTkMaxK / 32];
#pragma unroll
  for (int t = 0; t < kTkMaxK / 32; ++t) {
    const int e = lane + 32 * t;
    if (e >= pos && e < last) { ts[t] = ls[e]; ti[t] = li[e]; }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < kTkMaxK / 32; ++t) {
    const int e = lane + 32 * t;
    if (e >= pos && e < last) { ls[e + 1] = ts[t]; li[e + 1] = ti[t]; }
  }
  if (lane == 0) { ls[pos] = s; li[pos] = id; }
  __syncwarp();
  return min(cnt + 1, k);
}

__global__ void __launch_bounds__(kTkThreads)
score_topk_kernel(const float* __restrict__ queries, const float* __restrict__ items, int Q, int N, int d,
                  int k, int items_per_split, float* __restrict__ cand_scores, int* __restrict__ cand_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*As)[kTkPad] = reinterpret_cast<float (*)[kTkPad]>(smem_raw);
  float (*Bs)[kTkPad] = As + kTkBK;
  float (*St)[kTkTile + 1] = reinterpret_cast<float (*)[kTkTile + 1]>(Bs + kTkBK);
  float* lscore = reinterpret_cast<float*>(St + kTkTile);       // [64][kTkMaxK]
  int* lidx = reinterpret_cast<int*>(lscore + kTkTile * kTkMaxK);

  const int q0 = blockIdx.x * kTkTile;
  const int split = blockIdx.y;
  const int n_begin = split * items_per_split;
  const int n_end = min(N, n_begin + items_per_split);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  int cnt[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) cnt[r] = 0;

  for (int n0 = n_begin; n0 < n_end; n0 += kTkTile) {
    float acc[4][4] = {};
    for (int kk = 0; kk < d; kk += kTkBK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * kTkThreads;
        const int r = idx >> 4, c = idx & 15;
        As[c][r] = (q0 + r < Q && kk + c < d) ? queries[(int64_t)(q0 + r) * d + kk + c] : 0.f;
        Bs[c][r] = (n0 + r < n_end && kk + c < d) ? items[(int64_t)(n0 + r) * d + kk + c] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kq = 0; kq < kTkBK; ++kq) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kq][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kq][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) St[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();
    // warp w merges rows 8w .. 8w+7
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = warp * 8 + r;
      float* ls = lscore + row * kTkMaxK;
      int* li = lidx + row * kTkMaxK;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = lane + 32 * half;
        const float s = St[row][col];
        const bool valid = (n0 + col < n_end) && (q0 + row < Q);
        const float thr = cnt[r] == k ? ls[k - 1] : -INFINITY;
        unsigned pending = __ballot_sync(0xffffffffu, valid && (cnt[r] < k || s > thr));
        while (pending) {
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const float cs = __shfl_sync(0xffffffffu, s, src);
          if (cnt[r] < k || cs > ls[k - 1]) cnt[r] = warp_insert(ls, li, cnt[r], k, cs, n0 + 32 * half + src, lane);
        }
      }
    }
    __syncthreads();
  }
  // write this split's list: [Q][num_splits][k]
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = warp * 8 + r;
    if (q0 + row >= Q) continue;
    const int64_t base = ((int64_t)(q0 + row) * gridDim.y + split) * k;
    for (int e = lane; e < k; e += 32) {
      const bool have = e < cnt[r];
      cand_scores[base + e] = have ? lscore[row * kTkMaxK + e] : -INFINITY;
      cand_idx[base + e] = have ? lidx[row * kTkMaxK + e] : -1;
    }
  }
}

// One warp per query merges its `splits` sorted lists (ascending item ranges).
__global__ void __launch_bounds__(kTkThreads)
topk_merge_kernel(const float* __restrict__ cand_scores, const int* __restrict__ cand_idx, int Q, int splits,
                  int k, int64_t index_base, float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
  __shared__ float ls[kTkThreads / 32][kTkMaxK];
  __shared__ int li[kTkThreads / 32][kTkMaxK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (kTkThreads / 32) + warp;
  if (q >= Q) return;
  int cnt = 0;
  for (int sp = 0; sp < splits; ++sp) {
    const int64_t base = ((int64_t)q * splits + sp) * k;
    for (int e = 0; e < k; ++e) {
      const int id = cand_idx[base + e];
      if (id < 0) break;                     // lists are dense prefixes
      const float s = cand_scores[base + e];
      if (cnt == k && !(s > ls[warp][k - 1])) break;  // sorted: nothing later can enter either
      cnt = warp_insert(ls[warp], li[warp], cnt, k, s, id, lane);
    }
  }
  for (int e = lane; e < k; e += 32) {
    const bool have = e < cnt;
    out_scores[(int64_t)q * k + e] = have ? ls[warp][e] : -INFINITY;
    out_idx[(int64_t)q * k + e] = have ? (int64_t)li[warp][e] + index_base : (int64_t)-1;
  }
}

static size_t score_topk_smem() {
  return (size_t)2 * kTkBK * kTkPad * 4 + (size_t)kTkTile * (kTkTile + 1) * 4 + (size_t)kTkTile * kTkMaxK * 8;
}

static int topk_splits(int64_t Q, int64_t N) {
  int64_t qtiles = (Q + kTkTile - 1) / kTkTile;
  int64_t want = (2 * kNumSMs + qtiles - 1) / qtiles;
  int64_t max_splits = (N + 4 * kTkTile - 1) / (4 * kTkTile);  // at least 256 items per split
  if (want > max_splits) want = max_splits;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace tt

using namespace tt;

extern "C" {

size_t tt_topk_workspace_bytes(int64_t Q, int64_t N, int64_t k) {
  int s = topk_splits(Q, N);
  return 2 * align_up((size_t)Q * s * k * 4, 256) + 512;
}

int tt_score_topk_f32(const float* queries, const float* items, int64_t Q, int64_t N, int64_t d, int64_t k,
                      int64_t item_index_base, float* out_scores, int64_t* out_indices, void* ws,
                      size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(Q >= 0 && N >= 0 && d > 0 && k > 0, "score_topk: bad shape");
  if (k > kTkMaxK) return fail(TT_ERR_UNSUPPORTED, "score_topk: k=%lld > %d", (long long)k, kTkMaxK);
  if (N >= ((int64_t)1 << 31) || Q >= ((int64_t)1 << 31)) return fail(TT_ERR_UNSUPPORTED, "score_topk: too large");
  if (Q == 0) return TT_OK;
  TT_CHECK_ARG(queries && out_scores && out_indices && (N == 0 || items), "score_topk: null pointer");
  cudaStream_t s = as_stream(stream);
  const int splits = topk_splits(Q, N);
  Workspace w(ws, ws_bytes);
  float* cs = w.take<float>((size_t)Q * splits * k);
  int* ci = w.take<int>((size_t)Q * splits * k);
  if (!cs || !ci) return fail(TT_ERR_WORKSPACE, "score_topk: workspace too small");
  const int per = (int)(((N + splits - 1) / splits + kTkTile - 1) / kTkTile * kTkTile);
  const size_t smem = score_topk_smem();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "score_topk: smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)((Q + kTkTile - 1) / kTkTile), (unsigned)splits);
  score_topk_kernel<<<grid, kTkThreads, smem, s>>>(queries, items, (int)Q, (int)N, (int)d, (int)k, per > 0 ? per : kTkTile, cs, ci);
  TT_CHECK_LAUNCH("score_topk");
  topk_merge_kernel<<<(unsigned)((Q + 7) / 8), kTkThreads, 0, s>>>(cs, ci, (int)Q, splits, (int)k, item_index_base,
                                                                 out_scores, out_indices);
  TT_CHECK_LAUNCH("topk_merge");
  return TT_OK;
}

}  // extern "C"
