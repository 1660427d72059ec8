// Retrieval: exact dot-product top-k (04_evaluate_retrieval.py:134-141, k=100).
// fp32 CUDA-core scoring with the per-block top-k fused into the scoring CTA:
// the [Q, N] score matrix is never written.  A CTA owns 64 queries and one
// contiguous range of items, walks the range in 64-item tiles, and keeps a sorted
// top-k list per query in shared memory; candidates enter only when they beat
// the list's current k-th score.  A second kernel merges the per-range lists.
// Order: descending score, ties -> lower item index (strict ">" on entry keeps it
// because items are visited in ascending index order).
#include "common.cuh"
#include "tc_common.cuh"

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


// ---------------------------------------------------------------------------
// Tensor-core scoring (tcgen05) with the top-k fused into the epilogue.
//   CTA = 128 queries x one contiguous item range.  Q tile resident in smem (bf16, TMA);
//   item tiles of 128 rows stream through a TMA ring; tcgen05.mma (M=128, N=128, K=64) writes
//   score tiles into a 4-stage TMEM ring.  Epilogue: thread = query row.  A row keeps its
//   current k-th best score in a register; a 32-score chunk whose maximum does not beat it is
//   dropped after one max-tree (the common case); survivors are appended to a small per-row
//   buffer in smem, and when any buffer of the warp runs full the warp merges all 32 buffers
//   into the rows' sorted lists cooperatively (warp_insert).  Same tie rule as the fp32 path.
// ---------------------------------------------------------------------------
namespace tc {

// warp_insert on __shared__ lists (the pointers keep their address space, so the compiler emits
// LDS/STS instead of generic accesses) with the rank found by ballots over the lanes' entries
// instead of a shuffle reduction: the insert is a latency chain, so fewer dependent steps matter.
__device__ __forceinline__ int warp_insert_smem(float* __restrict__ ls, int* __restrict__ li, int cnt, int k, float s,
                                                int id, int lane) {
  float es[kTkMaxK / 32];
  int ei[kTkMaxK / 32];
  int pos = 0;
#pragma unroll
  for (int t = 0; t < kTkMaxK / 32; ++t) {
    const int e = lane + 32 * t;
    const bool in = e < cnt;
    es[t] = in ? ls[e] : -INFINITY;
    ei[t] = in ? li[e] : 0;
    pos += __popc(__ballot_sync(0xffffffffu, in && es[t] >= s));
  }
  if (pos >= k) return cnt;
  const int last = min(cnt, k - 1);  // entries [pos, last) move one slot right
  __syncwarp();
#pragma unroll
  for (int t = 0; t < kTkMaxK / 32; ++t) {
    const int e = lane + 32 * t;
    if (e >= pos && e < last) { ls[e + 1] = es[t]; li[e + 1] = ei[t]; }
  }
  if (lane == 0) { ls[pos] = s; li[pos] = id; }
  __syncwarp();
  return min(cnt + 1, k);
}

constexpr int kTcTkThreads = 192;
constexpr int kTcTkNT = 128;        // items per tile
constexpr int kTcTkAcc = 4;         // TMEM ring (4 x 128 columns)
constexpr int kTcTkCap = 40;        // candidate buffer entries per row (flushed when > 8 are waiting)

__host__ __device__ inline size_t tc_topk_smem(int k, int stages) {
  return (size_t)128 * 128 /*Q*/ + (size_t)stages * kTcTkNT * 128 /*ring*/ + 1024 /*align*/ + 256 /*barriers*/ +
         (size_t)128 * k * 8 /*lists*/ + (size_t)128 * kTcTkCap * 8 /*buffers*/;
}

template <int kTcTkStages>   // smem ring depth: 3, or 2 when k > 100 needs the room
__global__ void __launch_bounds__(kTcTkThreads, 1)
tc_score_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmI, int Q, int N,
                     int k, int items_per_split, float* __restrict__ cand_scores, int* __restrict__ cand_idx) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // lists and candidate buffers first, addressed straight off the __shared__ array (LDS/STS);
  // the TMA tiles follow at the next 1024-byte boundary.
  float* lscore = reinterpret_cast<float*>(smem_raw);             // [128][k]
  int* lidx = reinterpret_cast<int*>(lscore + 128 * k);           // [128][k]
  float* bscore = reinterpret_cast<float*>(lidx + 128 * k);       // [128][cap]
  int* bidx = reinterpret_cast<int*>(bscore + 128 * kTcTkCap);    // [128][cap]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bidx + 128 * kTcTkCap);
  uint64_t* q_full = bars;
  uint64_t* i_full = bars + 1;                    // [stages]
  uint64_t* i_empty = i_full + kTcTkStages;       // [stages]
  uint64_t* acc_full = i_empty + kTcTkStages;     // [4]
  uint64_t* acc_empty = acc_full + kTcTkAcc;      // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kTcTkAcc);
  uint8_t* sQ = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars + 32) + 1023) & ~uintptr_t(1023));
  uint8_t* sI = sQ + 128 * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int split = blockIdx.y;
  const int n_begin = split * items_per_split;
  const int n_end = min(N, n_begin + items_per_split);
  const int T = n_end > n_begin ? (n_end - n_begin + kTcTkNT - 1) / kTcTkNT : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmI);
    mbar_init(q_full, 1);
    for (int s = 0; s < kTcTkStages; ++s) { mbar_init(&i_full[s], 1); mbar_init(&i_empty[s], 1); }
    for (int s = 0; s < kTcTkAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 128 * 128);
      tma_load_2d(sQ, &tmQ, q_full, 0, q0);
      for (int t = 0; t < T; ++t) {
        const int s = t % kTcTkStages;
        mbar_wait(&i_empty[s], ((t / kTcTkStages) & 1) ^ 1);
        mbar_expect_tx(&i_full[s], kTcTkNT * 128);
        tma_load_2d(sI + s * kTcTkNT * 128, &tmI, &i_full[s], 0, n_begin + t * kTcTkNT);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(128, kTcTkNT);
    if (elect_one()) {
      mbar_wait(q_full, 0);
      for (int t = 0; t < T; ++t) {
        const int s = t % kTcTkStages, as = t % kTcTkAcc;
        mbar_wait(&acc_empty[as], ((t / kTcTkAcc) & 1) ^ 1);
        mbar_wait(&i_full[s], (t / kTcTkStages) & 1);
        tc_fence_after();
        const uint64_t da = smem_desc_k_sw128(smem_u32(sQ));
        const uint64_t db = smem_desc_k_sw128(smem_u32(sI + s * kTcTkNT * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma_ss(tmem_base + as * kTcTkNT, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
        tc_commit(&i_empty[s]);
        tc_commit(&acc_full[as]);
      }
    }
  } else {
    const int qd = warp & 3;
    const int r_in = qd * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    float* my_bs = bscore + r_in * kTcTkCap;
    int* my_bi = bidx + r_in * kTcTkCap;
    float thr = -INFINITY;
    int bcnt = 0, lcnt = 0;
    // merge every row's pending candidates into its sorted list (whole warp, row by row)
    auto flush = [&]() {
      __syncwarp();
#pragma unroll 1
      for (int r = 0; r < 32; ++r) {
        const int c = __shfl_sync(0xffffffffu, bcnt, r);
        if (c == 0) continue;
        int lc = __shfl_sync(0xffffffffu, lcnt, r);
        const int rr = qd * 32 + r;
        float* ls = lscore + rr * k;
        int* li = lidx + rr * k;
        const float* bs = bscore + rr * kTcTkCap;
        const int* bi = bidx + rr * kTcTkCap;
        for (int e = 0; e < c; ++e) {
          const float s = bs[e];
          if (lc < k || s > ls[k - 1]) lc = warp_insert_smem(ls, li, lc, k, s, bi[e], lane);
        }
        if (lane == r) lcnt = lc;
      }
      __syncwarp();
      bcnt = 0;
      thr = (lcnt == k) ? lscore[r_in * k + k - 1] : -INFINITY;
    };
    for (int t = 0; t < T; ++t) {
      const int as = t % kTcTkAcc;
      mbar_wait(&acc_full[as], (t / kTcTkAcc) & 1);
      tc_fence_after();
      const uint32_t tcol = trow + as * kTcTkNT;
      const int nt0 = n_begin + t * kTcTkNT;
      uint32_t va[32], vb[32];
      tmem_ld32(tcol, va);
#pragma unroll 1
      for (int c0 = 0; c0 < kTcTkNT; c0 += 64) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t(&v)[32] = h == 0 ? va : vb;
          tmem_ld_wait();
          if (h == 0) tmem_ld32(tcol + c0 + 32, vb);
          else if (c0 + 64 < kTcTkNT) tmem_ld32(tcol + c0 + 64, va);
          if (__any_sync(0xffffffffu, bcnt > kTcTkCap - 32)) flush();
          const int n0 = nt0 + c0 + 32 * h;
          if (n0 + 32 <= n_end) {
            float a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 16]));
#pragma unroll
            for (int w = 8; w > 0; w >>= 1)
#pragma unroll
              for (int i = 0; i < w; ++i) a[i] = fmaxf(a[i], a[i + w]);
            if (a[0] > thr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float sc = __uint_as_float(v[j]);
                if (sc > thr) { my_bs[bcnt] = sc; my_bi[bcnt] = n0 + j; ++bcnt; }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float sc = __uint_as_float(v[j]);
              if (n0 + j < n_end && sc > thr) { my_bs[bcnt] = sc; my_bi[bcnt] = n0 + j; ++bcnt; }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    flush();
    if (q0 + r_in < Q) {
      const int64_t base = ((int64_t)(q0 + r_in) * gridDim.y + split) * k;
      for (int e = 0; e < k; ++e) {
        const bool have = e < lcnt;
        cand_scores[base + e] = have ? lscore[r_in * k + e] : -INFINITY;
        cand_idx[base + e] = have ? lidx[r_in * k + e] : -1;
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

}  // namespace tc

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

static void tc_topk_plan(int64_t Q, int64_t N, int64_t* splits_out, int64_t* per_out) {
  // ~2 waves of CTAs; item ranges are multiples of the 128-item tile, at least 512 items each
  const int64_t qtiles = (Q + 127) / 128;
  int64_t splits = (2 * kNumSMs + qtiles - 1) / qtiles;
  const int64_t max_splits = (N + 4 * 128 - 1) / (4 * 128);
  if (splits > max_splits) splits = max_splits;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  const int64_t per = ((N + splits - 1) / splits + 127) / 128 * 128;
  *splits_out = per > 0 ? (N + per - 1) / per : 1;
  *per_out = per > 0 ? per : 128;
}

size_t tt_topk_bf16_workspace_bytes(int64_t Q, int64_t N, int64_t k) {
  int64_t splits, per;
  tc_topk_plan(Q, N > 0 ? N : 1, &splits, &per);
  return 2 * align_up((size_t)Q * splits * k * 4, 256) + 512;
}

// Tensor-core variant: queries / items are bf16 copies (tt_cast_f32_to_bf16), d <= 64, k <= 128.
int tt_score_topk_bf16(const void* queries_bf16, int64_t ldq, const void* items_bf16, int64_t ldi, int64_t Q, int64_t N,
                       int64_t d, int64_t k, int64_t item_index_base, float* out_scores, int64_t* out_indices, void* ws,
                       size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(Q >= 0 && N >= 0 && d > 0 && k > 0, "score_topk_bf16: bad shape");
  if (d > 64) return fail(TT_ERR_UNSUPPORTED, "score_topk_bf16: d=%lld > 64 (use the fp32 path)", (long long)d);
  if (k > kTkMaxK) return fail(TT_ERR_UNSUPPORTED, "score_topk_bf16: k=%lld > %d", (long long)k, kTkMaxK);
  if (N >= ((int64_t)1 << 31) || Q >= ((int64_t)1 << 31)) return fail(TT_ERR_UNSUPPORTED, "score_topk_bf16: too large");
  if (Q == 0) return TT_OK;
  TT_CHECK_ARG(queries_bf16 && out_scores && out_indices && (N == 0 || items_bf16), "score_topk_bf16: null pointer");
  if (N == 0) return fail(TT_ERR_INVALID, "score_topk_bf16: empty corpus");
  cudaStream_t s = as_stream(stream);
  const int64_t qtiles = (Q + 127) / 128;
  int64_t splits, per;
  tc_topk_plan(Q, N, &splits, &per);
  Workspace w(ws, ws_bytes);
  float* cs = w.take<float>((size_t)Q * splits * k);
  int* ci = w.take<int>((size_t)Q * splits * k);
  if (!cs || !ci) return fail(TT_ERR_WORKSPACE, "score_topk_bf16: workspace too small");
  CUtensorMap tq, ti;
  int rc = tc::make_tmap_bf16_2d(&tq, queries_bf16, Q, d, ldq, 128);
  if (rc) return rc;
  rc = tc::make_tmap_bf16_2d(&ti, items_bf16, N, d, ldi, tc::kTcTkNT);
  if (rc) return rc;
  const int stages = k <= 100 ? 3 : 2;
  const size_t smem = tc::tc_topk_smem((int)k, stages);
  dim3 grid((unsigned)qtiles, (unsigned)splits);
  cudaError_t e;
  if (stages == 3) {
    e = cudaFuncSetAttribute(tc::tc_score_topk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TT_ERR_CUDA, "score_topk_bf16 smem attr (%zu B): %s", smem, cudaGetErrorString(e)); }
    tc::tc_score_topk_kernel<3><<<grid, tc::kTcTkThreads, smem, s>>>(tq, ti, (int)Q, (int)N, (int)k, (int)per, cs, ci);
  } else {
    e = cudaFuncSetAttribute(tc::tc_score_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TT_ERR_CUDA, "score_topk_bf16 smem attr (%zu B): %s", smem, cudaGetErrorString(e)); }
    tc::tc_score_topk_kernel<2><<<grid, tc::kTcTkThreads, smem, s>>>(tq, ti, (int)Q, (int)N, (int)k, (int)per, cs, ci);
  }
  TT_CHECK_LAUNCH("tc_score_topk");
  topk_merge_kernel<<<(unsigned)((Q + 7) / 8), kTkThreads, 0, s>>>(cs, ci, (int)Q, (int)splits, (int)k, item_index_base,
                                                                 out_scores, out_indices);
  TT_CHECK_LAUNCH("topk_merge");
  return TT_OK;
}

}  // extern "C"
