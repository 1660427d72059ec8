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
  if (splits == 1) {                           // one range: its sorted list is the answer
    for (int e = lane; e < k; e += 32) {
      const int id = cand_idx[(int64_t)q * k + e];
      out_scores[(int64_t)q * k + e] = cand_scores[(int64_t)q * k + e];
      out_idx[(int64_t)q * k + e] = id >= 0 ? (int64_t)id + index_base : (int64_t)-1;
    }
    return;
  }
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
//   CTA = 256 queries (two 128-row tiles) x one contiguous item range.  The query tiles are RESIDENT IN TMEM
//   (bf16 pairs, 32 columns each) and feed tcgen05.mma as the A operand from there (TS form): only the item tile
//   is read from shared memory, 2 KB per 128x64x16 MMA = 64 B/clk, half of the shared-memory pipe (SS-form
//   N=64 MMAs would need 192 B/clk and stall on it).  Item tiles of 128 rows stream through a 4-stage TMA ring
//   and are used TWICE (one MMA group per query tile, M=128, N=2x64, K=64), which halves the L2 -> SM operand
//   traffic per flop; scores land in a TMEM ring of 3 stages of 64 columns per query tile.  A stage goes back to the MMA issuer as soon as its 64 columns sit in
//   the epilogue's registers (before they are processed), so the issuer runs up to two item tiles ahead.
//   Epilogue: 8 warps (two per scheduler), thread = query row.  Per 32-score chunk a 3-input max tree (FMNMX3,
//   ~0.5 instruction per score) gives four 8-score group maxima; only groups that beat the row's threshold are
//   scanned.  Survivors are APPENDED to an unsorted per-row buffer (128 entries, global memory, L2 resident) --
//   no sorted insertion.  The threshold is the row's k-th best score as of the last compaction, so it is stale
//   by at most one buffer; the number of appends over a pass stays O(k log(n/k)).  When a buffer of the warp
//   is about to run full the warp compacts all 32 rows cooperatively: bitonic sort of the 128 pending entries
//   in registers (64-bit keys = ordered score | inverted index, so ties go to the lower index), bitonic merge
//   with the row's sorted top-k list, write back, new threshold.  Same tie rule as the fp32 path.
// ---------------------------------------------------------------------------
namespace tc {

constexpr int kTcTkThreads = 320;   // warp 0: TMA, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int kTcTkNT = 128;        // items per tile
constexpr int kTcTkQT = 2;          // query tiles (128 rows each) per CTA
constexpr int kTcTkHalf = 64;        // columns per TMEM stage: the 128-item tile is scored as two N=64 MMA groups
constexpr int kTcTkAcc = 3;         // TMEM stages (64 columns each) per query tile (2 x 3 x 64 = 384 columns)
constexpr int kTcTkACol = kTcTkQT * kTcTkAcc * kTcTkHalf;   // the query tiles themselves live in TMEM from here (32 columns each)
constexpr int kTcTkStages = 4;      // smem ring depth for item tiles
constexpr int kTcTkPend = 128;      // pending-buffer entries per row
constexpr int kTcTkRows = 128 * kTcTkQT;

__host__ __device__ inline size_t tc_topk_smem() {
  return (size_t)kTcTkStages * kTcTkNT * 128 /*ring*/ + 1024 /*align*/ + 256 /*barriers*/;
}

// monotone map float -> uint32 (larger float <-> larger integer), and back
__device__ __forceinline__ uint32_t ord_of(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of_ord(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// key of a candidate: descending key order == (descending score, ascending index); 0 = "no entry"
__device__ __forceinline__ unsigned long long key_of(float s, int idx) {
  return (static_cast<unsigned long long>(ord_of(s)) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(idx));
}

// Bitonic network over 128 keys held by a warp, element E = lane * 4 + i.  Sorts blocks of FIRST_BLK..128
// (FIRST_BLK = 2: full sort; FIRST_BLK = 128: merge of a bitonic sequence), result descending.
template <int FIRST_BLK>
__device__ __forceinline__ void bitonic128_desc(unsigned long long (&key)[4], int lane) {
#pragma unroll
  for (int blk = FIRST_BLK; blk <= 128; blk <<= 1) {
#pragma unroll
    for (int st = blk >> 1; st > 0; st >>= 1) {
      if (st < 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((i & st) == 0) {
            const int j = i | st;
            const bool desc = blk >= 4 ? ((lane * 4) & blk) == 0 : (i & blk) == 0;
            const unsigned long long a = key[i], b = key[j];
            const unsigned long long mx = a > b ? a : b, mn = a > b ? b : a;
            key[i] = desc ? mx : mn;
            key[j] = desc ? mn : mx;
          }
        }
      } else {
        const int ls = st >> 2;
        const bool lower = (lane & ls) == 0;
        const bool desc = ((lane * 4) & blk) == 0;
        const bool take_max = lower == desc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, key[i], ls);
          const bool gt = key[i] > o;
          key[i] = (gt == take_max) ? key[i] : o;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kTcTkThreads, 1)
tc_score_topk_kernel(const __nv_bfloat16* __restrict__ queries, int ldq, int d, const __grid_constant__ CUtensorMap tmI,
                     int Q, int N, int k, int items_per_split, float* __restrict__ cand_scores, int* __restrict__ cand_idx,
                     float2* __restrict__ pend) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars;
  uint64_t* i_full = bars + 1;                         // [stages]
  uint64_t* i_empty = i_full + kTcTkStages;            // [stages]
  uint64_t* acc_full = i_empty + kTcTkStages;          // [QT][Acc]
  uint64_t* acc_empty = acc_full + kTcTkQT * kTcTkAcc; // [QT][Acc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kTcTkQT * kTcTkAcc);
  uint8_t* sI = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars + 32) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTcTkRows;
  const int split = blockIdx.y;
  const int n_begin = split * items_per_split;
  const int n_end = min(N, n_begin + items_per_split);
  const int T = n_end > n_begin ? (n_end - n_begin + kTcTkNT - 1) / kTcTkNT : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmI);
    mbar_init(q_full, kTcTkRows);          // every epilogue thread arrives once its query row sits in TMEM
    for (int s = 0; s < kTcTkStages; ++s) { mbar_init(&i_full[s], 1); mbar_init(&i_empty[s], 1); }
    for (int s = 0; s < kTcTkQT * kTcTkAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
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
        const int s = t % kTcTkStages;
        mbar_wait(&i_empty[s], ((t / kTcTkStages) & 1) ^ 1);
        mbar_expect_tx(&i_full[s], kTcTkNT * 128);
        tma_load_2d(sI + s * kTcTkNT * 128, &tmI, &i_full[s], 0, n_begin + t * kTcTkNT);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(128, kTcTkHalf);
    if (elect_one()) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int s = t % kTcTkStages;
        mbar_wait(&i_full[s], (t / kTcTkStages) & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int hs = (2 * t + h) % kTcTkAcc;
          const uint32_t par = (((2 * t + h) / kTcTkAcc) & 1) ^ 1;
          // item rows [64h, 64h + 64) of the tile: 64 rows x 128 B = 8 KB further into the stage
          const uint64_t db = smem_desc_k_sw128(smem_u32(sI + s * kTcTkNT * 128 + h * kTcTkHalf * 128));
#pragma unroll
          for (int qt = 0; qt < kTcTkQT; ++qt) {
            mbar_wait(&acc_empty[qt * kTcTkAcc + hs], par);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma_ts(tmem_base + (qt * kTcTkAcc + hs) * kTcTkHalf, tmem_base + kTcTkACol + qt * 32 + kk * 8, db + 2 * kk, idesc,
                     kk != 0);
            tc_commit(&acc_full[qt * kTcTkAcc + hs]);
          }
        }
        tc_commit(&i_empty[s]);
      }
    }
  } else {
    const int ew = warp - 2;                // 0..7
    const int qt = ew >> 2;                 // query tile of this warp
    const int qd = warp & 3;                // TMEM lane quarter this warp may read
    const int r_in = qd * 32 + lane;
    const int wrow0 = q0 + qt * 128 + qd * 32;          // first query row of this warp
    const int qrow = wrow0 + lane;
    const bool row_ok = qrow < Q;
    const int64_t slot_me = (int64_t)qrow * gridDim.y + split;      // (query, split) slot
    float2* my_pend = pend + slot_me * kTcTkPend;     // unsorted (score, index bits) entries
    {
      // this thread's query row -> TMEM (A operand of every MMA of the CTA): 64 bf16 = 32 columns, zero beyond d / Q
      uint32_t a[32];
      const uint4* src = reinterpret_cast<const uint4*>(queries + (int64_t)qrow * ldq);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && j * 8 < d) x = __ldg(src + j);
        a[4 * j] = x.x; a[4 * j + 1] = x.y; a[4 * j + 2] = x.z; a[4 * j + 3] = x.w;
      }
      if (d & 7) {                                  // the pad elements of the last 16-byte group are not zero in memory
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (2 * c + 1 >= d) a[c] = 2 * c < d ? (a[c] & 0xffffu) : 0u;
        }
      }
      const uint32_t tdst = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + kTcTkACol + qt * 32;
      tmem_st16(tdst, reinterpret_cast<const uint32_t(&)[16]>(a[0]));
      tmem_st16(tdst + 16, reinterpret_cast<const uint32_t(&)[16]>(a[16]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_full);
    }
    float thr = row_ok ? -INFINITY : INFINITY;     // rows beyond Q never fire
    int pcnt = 0, lcnt = 0;

    // Merge every row's pending entries into its sorted top-k list (whole warp, one row at a time).
    // final = true also writes the (-inf, -1) padding of short lists that the cross-split merge expects.
    auto compact = [&](bool final) {
      __syncwarp();
#pragma unroll 1
      for (int r = 0; r < 32; ++r) {
        const int c = __shfl_sync(0xffffffffu, pcnt, r);
        const int lc = __shfl_sync(0xffffffffu, lcnt, r);
        if (wrow0 + r >= Q) break;
        if (c == 0 && !final) continue;
        const int64_t slot = (int64_t)(wrow0 + r) * gridDim.y + split;
        const float4* pp = reinterpret_cast<const float4*>(pend + slot * kTcTkPend);
        float* ls = cand_scores + slot * k;
        int* li = cand_idx + slot * k;
        unsigned long long key[4], old[4];
        {
          const float4 p01 = __ldcg(pp + lane * 2), p23 = __ldcg(pp + lane * 2 + 1);   // entries 4 lane .. 4 lane + 3
          const int e = lane * 4;
          key[0] = e + 0 < c ? key_of(p01.x, __float_as_int(p01.y)) : 0ull;
          key[1] = e + 1 < c ? key_of(p01.z, __float_as_int(p01.w)) : 0ull;
          key[2] = e + 2 < c ? key_of(p23.x, __float_as_int(p23.y)) : 0ull;
          key[3] = e + 3 < c ? key_of(p23.z, __float_as_int(p23.w)) : 0ull;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = lane * 4 + i;
          old[i] = e < lc ? key_of(__ldcg(ls + e), __ldcg(li + e)) : 0ull;
        }
        bitonic128_desc<2>(key, lane);
        // old list is sorted descending: max(new[E], old[127 - E]) is a bitonic sequence holding the best 128
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned long long o = __shfl_sync(0xffffffffu, old[3 - i], 31 - lane);
          key[i] = key[i] > o ? key[i] : o;
        }
        bitonic128_desc<128>(key, lane);
        const int nc = min(lc + c, k);
        float kth = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = lane * 4 + i;
          const float sc = float_of_ord(static_cast<uint32_t>(key[i] >> 32));
          const int id = static_cast<int>(~static_cast<uint32_t>(key[i]));
          if (e < nc) { __stcg(ls + e, sc); __stcg(li + e, id); }
          else if (final && e < k) { __stcg(ls + e, -INFINITY); __stcg(li + e, -1); }
          if (e == k - 1) kth = sc;
        }
        kth = __shfl_sync(0xffffffffu, kth, (k - 1) >> 2);
        if (lane == r) {
          lcnt = nc;
          pcnt = 0;
          if (nc == k) thr = kth;
        }
      }
      __syncwarp();
    };

    const uint32_t trow = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + qt * kTcTkAcc * kTcTkHalf;
    uint64_t* my_full = acc_full + qt * kTcTkAcc;
    uint64_t* my_empty = acc_empty + qt * kTcTkAcc;

    // One 32-score chunk: group maxima by a 3-input max tree; groups that beat the threshold are scanned.
    auto process = [&](uint32_t(&v)[32], int n0) {
      if (n0 + 32 > n_end) {                     // last (partial) tile: columns beyond the range do not exist
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j >= n_end) v[j] = __float_as_uint(-INFINITY);
      }
      float m[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float* f = reinterpret_cast<const float*>(&v[g * 8]);
        const float a = fmaxf(fmaxf(f[0], f[1]), f[2]);
        const float b = fmaxf(fmaxf(f[3], f[4]), f[5]);
        m[g] = fmaxf(fmaxf(a, b), fmaxf(f[6], f[7]));
      }
      const float top = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
      if (__any_sync(0xffffffffu, top > thr)) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (m[g] > thr) {
            const float* f = reinterpret_cast<const float*>(&v[g * 8]);
            unsigned hit = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) hit |= (f[j] > thr ? 1u : 0u) << j;
            if ((hit & (hit - 1)) == 0) {
              // exactly one score of the group beats the threshold (the usual case): it is the group maximum
              __stcg(my_pend + pcnt, make_float2(m[g], __int_as_float(n0 + g * 8 + (__ffs(hit) - 1))));
              ++pcnt;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (hit & (1u << j)) {
                  __stcg(my_pend + pcnt, make_float2(f[j], __int_as_float(n0 + g * 8 + j)));
                  ++pcnt;
                }
              }
            }
          }
        }
        if (__any_sync(0xffffffffu, pcnt > kTcTkPend - 32)) compact(false);
      }
    };
    // 64-column stage hs of this query tile -> chunk pair; the flat chunk index ci walks 4 chunks per item tile
    const int total_chunks = 4 * T;
    uint32_t va[32], vb[32];
    if (total_chunks > 0) {
      mbar_wait(&my_full[0], 0);
      tc_fence_after();
      tmem_ld32(trow, va);
    }
#pragma unroll 1
    for (int ci = 0; ci < total_chunks; ci += 2) {
      const int half = ci >> 1;                         // 64-column half tile index = 2 t + h
      const int hs = half % kTcTkAcc;
      const int n0 = n_begin + half * kTcTkHalf;
      tmem_ld_wait();                                   // chunk ci is in va
      tmem_ld32(trow + hs * kTcTkHalf + 32, vb);
      process(va, n0);
      tmem_ld_wait();                                   // chunk ci + 1 is in vb: the whole stage is in registers
      tc_fence_before();
      mbar_arrive(&my_empty[hs]);
      if (ci + 2 < total_chunks) {
        const int nh = half + 1;
        mbar_wait(&my_full[nh % kTcTkAcc], (nh / kTcTkAcc) & 1);
        tc_fence_after();
        tmem_ld32(trow + (nh % kTcTkAcc) * kTcTkHalf, va);
      }
      process(vb, n0 + 32);
    }
    compact(true);
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
  // One CTA = 256 queries x one item range.  Few, long ranges: every range restarts its thresholds at -inf, and
  // the appends / compactions a row pays grow only with log(range length).  Take the smallest number of ranges
  // that fills >= 85 % of the SMs in the last wave; ranges are multiples of the 128-item tile, >= 1024 items.
  const int64_t qtiles = (Q + tc::kTcTkRows - 1) / tc::kTcTkRows;
  int64_t max_splits = N / 1024;
  if (max_splits > 64) max_splits = 64;
  if (max_splits < 1) max_splits = 1;
  int64_t splits = max_splits;
  double best_eff = -1.0;
  for (int64_t sp = 1; sp <= max_splits; ++sp) {
    const int64_t ctas = qtiles * sp;
    const double eff = (double)ctas / (double)(((ctas + kNumSMs - 1) / kNumSMs) * kNumSMs);
    if (eff >= 0.85) { splits = sp; best_eff = eff; break; }
    if (eff > best_eff + 1e-9) { best_eff = eff; splits = sp; }
  }
  const int64_t per = ((N + splits - 1) / splits + 127) / 128 * 128;
  *splits_out = per > 0 ? (N + per - 1) / per : 1;
  *per_out = per > 0 ? per : 128;
}

size_t tt_topk_bf16_workspace_bytes(int64_t Q, int64_t N, int64_t k) {
  int64_t splits, per;
  tc_topk_plan(Q, N > 0 ? N : 1, &splits, &per);
  return 2 * align_up((size_t)Q * splits * k * 4, 256) + align_up((size_t)Q * splits * tc::kTcTkPend * 8, 256) + 512;
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
  const int64_t qtiles = (Q + tc::kTcTkRows - 1) / tc::kTcTkRows;
  int64_t splits, per;
  tc_topk_plan(Q, N, &splits, &per);
  Workspace w(ws, ws_bytes);
  float* cs = w.take<float>((size_t)Q * splits * k);
  int* ci = w.take<int>((size_t)Q * splits * k);
  float2* pend = w.take<float2>((size_t)Q * splits * tc::kTcTkPend);
  if (!cs || !ci || !pend) return fail(TT_ERR_WORKSPACE, "score_topk_bf16: workspace too small");
  TT_CHECK_ARG(ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(queries_bf16) & 15) == 0,
               "score_topk_bf16: query rows must be 16-byte aligned (pitch a multiple of 8)");
  CUtensorMap ti;
  int rc = tc::make_tmap_bf16_2d(&ti, items_bf16, N, d, ldi, tc::kTcTkNT);
  if (rc) return rc;
  const size_t smem = tc::tc_topk_smem();
  dim3 grid((unsigned)qtiles, (unsigned)splits);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::tc_score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TT_ERR_CUDA, "score_topk_bf16 smem attr (%zu B): %s", smem, cudaGetErrorString(e)); }
    attr_set = true;
  }
  tc::tc_score_topk_kernel<<<grid, tc::kTcTkThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(queries_bf16), (int)ldq, (int)d, ti,
                                                                 (int)Q, (int)N, (int)k, (int)per, cs, ci, pend);
  TT_CHECK_LAUNCH("tc_score_topk");
  topk_merge_kernel<<<(unsigned)((Q + 7) / 8), kTkThreads, 0, s>>>(cs, ci, (int)Q, (int)splits, (int)k, item_index_base,
                                                                 out_scores, out_indices);
  TT_CHECK_LAUNCH("topk_merge");
  return TT_OK;
}

}  // extern "C"
