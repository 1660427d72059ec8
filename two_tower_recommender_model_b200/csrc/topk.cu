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
//   CTA = 256 queries (two 128-row tiles, resident in smem as bf16) x one contiguous item range.  Item tiles of
//   128 rows stream through a 4-stage TMA ring and are used TWICE (one tcgen05.mma group per query tile,
//   M=128, N=128, K=64), which halves the L2 -> SM operand traffic per flop; score tiles land in a TMEM ring of
//   2 stages per query tile (4 x 128 columns).
//   Warpgroup 0 = {TMA producer, MMA issuer}; warpgroups 1-2 = 8 epilogue warps (two per scheduler), thread =
//   query row; setmaxnreg moves registers from warpgroup 0 to the epilogue.  An epilogue thread pulls ALL 128
//   scores of its row into registers and hands the TMEM stage back BEFORE it looks at them, so the issuer refills
//   the stage while the scores are filtered (one barrier round trip per 128 scores, and it overlaps the work).
//   Filter: a 3-input max tree (FMNMX3, ~0.5 instruction per score) gives sixteen 8-score group maxima and the
//   row maximum; one warp vote per tile; only groups that beat the row's threshold are scanned.  Survivors are
//   APPENDED to an unsorted per-row buffer (256 entries, global memory, L2 resident) -- no sorted insertion.  The
//   threshold is the row's k-th best score as of the last compaction, so it is stale by at most one buffer; the
//   number of appends over a pass stays O(k log(n/k)).  When a buffer of the warp passes 96 entries the warp
//   compacts all 32 rows cooperatively: bitonic sort of 128 pending entries in registers (64-bit keys = ordered
//   score | inverted index, so ties go to the lower index), bitonic merge with the row's sorted top-k list, write
//   back, new threshold.  Same tie rule as the fp32 path.
// ---------------------------------------------------------------------------
namespace tc {

// Per-role cycle accounting (build with -DTT_TOPK_PROFILE; read with tt_debug_read_counters): where the
// epilogue warps, the MMA issuer and the TMA producer spend their time.  Compiled out otherwise.
__device__ unsigned long long g_topk_prof[16];
#ifdef TT_TOPK_PROFILE
#define TT_PROF_DECL(n) long long n = 0
#define TT_PROF_T0(t) const long long t = clock64()
#define TT_PROF_ADD(acc, t) acc += clock64() - t
#define TT_PROF_FLUSH(i, acc) atomicAdd(&g_topk_prof[i], (unsigned long long)(acc))
#else
#define TT_PROF_DECL(n)
#define TT_PROF_T0(t)
#define TT_PROF_ADD(acc, t)
#define TT_PROF_FLUSH(i, acc)
#endif

constexpr int kTcTkThreads = 384;   // warpgroup 0: warp 0 TMA, warps 1-2 MMA issuers (one per query tile); warpgroups 1-2: epilogue
constexpr int kTcTkNT = 128;        // items per tile = columns per TMEM stage
constexpr int kTcTkQT = 2;          // query tiles (128 rows each) per CTA
constexpr int kTcTkAcc = 2;         // TMEM stages per query tile (2 x 2 x 128 = 512 columns)
constexpr int kTcTkStages = 8;      // smem ring depth for item tiles (16 KB each): covers the L2 / HBM latency of the stream
constexpr int kTcTkPend = 256;      // pending-buffer entries per row (a tile adds at most 128)
constexpr int kTcTkTrig = 96;       // compact when a row of the warp holds more pending entries than this
constexpr int kTcTkRows = 128 * kTcTkQT;
constexpr int kTcTkRegsCtrl = 56, kTcTkRegsEpi = 224;   // setmaxnreg: 56 + 2 * 224 = 504 <= 3 * 168 (launch value)
static_assert(kTcTkRegsCtrl + 2 * kTcTkRegsEpi <= 3 * ((65536 / kTcTkThreads) / 8 * 8), "setmaxnreg budget");

__host__ __device__ inline size_t tc_topk_smem() {
  return (size_t)kTcTkQT * 128 * 128 /*Q*/ + (size_t)kTcTkStages * kTcTkNT * 128 /*ring*/ + 1024 /*align*/ + 256 /*barriers*/;
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

struct RowState {
  float thr;      // the row's k-th best score as of the last compaction (-inf until the list holds k entries)
  int pcnt;       // pending (unsorted) entries
  int lcnt;       // entries of the sorted list
};

// Merge every row's pending entries into its sorted top-k list (whole warp, one row at a time, 128 pending
// entries per round).  final = true also writes the (-inf, -1) padding of short lists that the cross-split
// merge expects.  One copy in the binary (noinline): the bitonic networks are ~5000 instructions, and inlined
// at every call site they pushed the scoring loop out of the instruction cache.
__device__ __noinline__ RowState compact_rows(RowState st, bool final, int wrow0, int Q, int k, int splits, int split, int lane,
                                              float* __restrict__ cand_scores, int* __restrict__ cand_idx,
                                              const float2* __restrict__ pend) {
  float thr = st.thr;
  int pcnt = st.pcnt, lcnt = st.lcnt;
  __syncwarp();
#pragma unroll 1
  for (int r = 0; r < 32; ++r) {
    const int c = __shfl_sync(0xffffffffu, pcnt, r);
    const int lc = __shfl_sync(0xffffffffu, lcnt, r);
    if (wrow0 + r >= Q) break;
    if (c == 0 && !final) continue;
    const int64_t slot = (int64_t)(wrow0 + r) * splits + split;
    float* ls = cand_scores + slot * k;
    int* li = cand_idx + slot * k;
    unsigned long long key[4], lst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane * 4 + i;
      lst[i] = e < lc ? key_of(__ldcg(ls + e), __ldcg(li + e)) : 0ull;
    }
#pragma unroll 1
    for (int base = 0; base < c; base += 128) {
      const float4* pp = reinterpret_cast<const float4*>(pend + slot * kTcTkPend + base);
      const float4 p01 = __ldcg(pp + lane * 2), p23 = __ldcg(pp + lane * 2 + 1);   // entries 4 lane .. 4 lane + 3
      const int e = base + lane * 4;
      key[0] = e + 0 < c ? key_of(p01.x, __float_as_int(p01.y)) : 0ull;
      key[1] = e + 1 < c ? key_of(p01.z, __float_as_int(p01.w)) : 0ull;
      key[2] = e + 2 < c ? key_of(p23.x, __float_as_int(p23.y)) : 0ull;
      key[3] = e + 3 < c ? key_of(p23.z, __float_as_int(p23.w)) : 0ull;
      bitonic128_desc<2>(key, lane);
      // the list is sorted descending: max(new[E], list[127 - E]) is a bitonic sequence holding the best 128
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned long long o = __shfl_sync(0xffffffffu, lst[3 - i], 31 - lane);
        key[i] = key[i] > o ? key[i] : o;
      }
      bitonic128_desc<128>(key, lane);
#pragma unroll
      for (int i = 0; i < 4; ++i) lst[i] = key[i];
    }
    const int nc = min(lc + c, k);
    float kth = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane * 4 + i;
      const float sc = float_of_ord(static_cast<uint32_t>(lst[i] >> 32));
      const int id = static_cast<int>(~static_cast<uint32_t>(lst[i]));
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
  return RowState{thr, pcnt, lcnt};
}

template <int N>
__device__ __forceinline__ void tk_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void tk_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(kTcTkThreads, 1)
tc_score_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmI, int Q, int N,
                     int k, int items_per_split, float* __restrict__ cand_scores, int* __restrict__ cand_idx,
                     float2* __restrict__ pend) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars;
  uint64_t* i_full = bars + 1;                         // [stages]
  uint64_t* i_empty = i_full + kTcTkStages;            // [stages]
  uint64_t* acc_full = i_empty + kTcTkStages;          // [QT][Acc]
  uint64_t* acc_empty = acc_full + kTcTkQT * kTcTkAcc; // [QT][Acc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kTcTkQT * kTcTkAcc);
  volatile uint32_t* compact_epoch = tmem_slot + 1;   // bumped by a warp that must compact: all epilogue warps follow
  uint8_t* sQ = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars + 32) + 1023) & ~uintptr_t(1023));
  uint8_t* sI = sQ + kTcTkQT * 128 * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTcTkRows;
  const int split = blockIdx.y;
  const int n_begin = split * items_per_split;
  const int n_end = min(N, n_begin + items_per_split);
  const int T = n_end > n_begin ? (n_end - n_begin + kTcTkNT - 1) / kTcTkNT : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmI);
    mbar_init(q_full, 1);
    *compact_epoch = 0;
    for (int s = 0; s < kTcTkStages; ++s) { mbar_init(&i_full[s], 1); mbar_init(&i_empty[s], kTcTkQT); }
    for (int s = 0; s < kTcTkQT * kTcTkAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    tk_setmaxnreg_dec<kTcTkRegsCtrl>();
    if (warp == 0) {
      if (elect_one()) {
        TT_PROF_DECL(p_wait);
        TT_PROF_T0(p_all);
        mbar_expect_tx(q_full, kTcTkQT * 128 * 128);
#pragma unroll
        for (int qt = 0; qt < kTcTkQT; ++qt) tma_load_2d(sQ + qt * 128 * 128, &tmQ, q_full, 0, q0 + qt * 128);
        for (int t = 0; t < T; ++t) {
          const int s = t % kTcTkStages;
          TT_PROF_T0(p0);
          mbar_wait(&i_empty[s], ((t / kTcTkStages) & 1) ^ 1);
          TT_PROF_ADD(p_wait, p0);
          mbar_expect_tx(&i_full[s], kTcTkNT * 128);
          tma_load_2d(sI + s * kTcTkNT * 128, &tmI, &i_full[s], 0, n_begin + t * kTcTkNT);
        }
        TT_PROF_FLUSH(12, p_wait);
        TT_PROF_FLUSH(13, clock64() - p_all);
      }
    } else if (warp <= kTcTkQT) {
      // one MMA issuer per query tile (warps 1 and 2): each waits only for ITS TMEM stages, so a query tile whose
      // epilogue is busy (compaction) does not hold the other one back, and the serial wait -> issue -> commit
      // chain of a single thread (about as long as the MMAs themselves) is split in two
      constexpr uint32_t idesc = idesc_bf16_f32(128, kTcTkNT);
      const int qt = warp - 1;
      if (elect_one()) {
        mbar_wait(q_full, 0);
        TT_PROF_DECL(m_wi);
        TT_PROF_DECL(m_we);
        TT_PROF_T0(m_all);
        const uint64_t da = smem_desc_k_sw128(smem_u32(sQ + qt * 128 * 128));
        for (int t = 0; t < T; ++t) {
          const int s = t % kTcTkStages, as = t % kTcTkAcc;
          TT_PROF_T0(m0);
          mbar_wait(&i_full[s], (t / kTcTkStages) & 1);
          TT_PROF_ADD(m_wi, m0);
          const uint64_t db = smem_desc_k_sw128(smem_u32(sI + s * kTcTkNT * 128));
          TT_PROF_T0(m1);
          mbar_wait(&acc_empty[qt * kTcTkAcc + as], ((t / kTcTkAcc) & 1) ^ 1);
          TT_PROF_ADD(m_we, m1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss(tmem_base + (qt * kTcTkAcc + as) * kTcTkNT, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
          tc_commit(&acc_full[qt * kTcTkAcc + as]);
          tc_commit(&i_empty[s]);          // the stage is free once BOTH issuers' MMAs on it are done (count 2)
        }
        if (qt == 0) {
          TT_PROF_FLUSH(8, m_wi);
          TT_PROF_FLUSH(9, m_we);
          TT_PROF_FLUSH(11, clock64() - m_all);
        }
      }
    }
  } else {
    tk_setmaxnreg_inc<kTcTkRegsEpi>();
    const int qt = (warp - 4) >> 2;         // query tile of this warp
    const int qd = warp & 3;                // TMEM lane quarter this warp may read
    const int wrow0 = q0 + qt * 128 + qd * 32;          // first query row of this warp
    const int qrow = wrow0 + lane;
    const bool row_ok = qrow < Q;
    const int64_t slot_me = (int64_t)qrow * gridDim.y + split;      // (query, split) slot
    float2* my_pend = pend + slot_me * kTcTkPend;     // unsorted (score, index bits) entries
    float thr = row_ok ? -INFINITY : INFINITY;     // rows beyond Q never fire
    int pcnt = 0, lcnt = 0;
    uint32_t seen_epoch = 0;
    TT_PROF_DECL(e_wf);
    TT_PROF_DECL(e_ld);
    TT_PROF_DECL(e_cp);

    auto compact = [&](bool final) {
      TT_PROF_T0(c0);
      const RowState ns = compact_rows(RowState{thr, pcnt, lcnt}, final, wrow0, Q, k, (int)gridDim.y, split, lane, cand_scores, cand_idx, pend);
      thr = ns.thr; pcnt = ns.pcnt; lcnt = ns.lcnt;
      TT_PROF_ADD(e_cp, c0);
    };

    const uint32_t trow = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + qt * kTcTkAcc * kTcTkNT;
    uint64_t* my_full = acc_full + qt * kTcTkAcc;
    uint64_t* my_empty = acc_empty + qt * kTcTkAcc;
    TT_PROF_T0(e_all);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int as = t % kTcTkAcc;
      const int n0 = n_begin + t * kTcTkNT;
      uint32_t v[128];
      {
        TT_PROF_T0(w0);
        mbar_wait(&my_full[as], (t / kTcTkAcc) & 1);
        TT_PROF_ADD(e_wf, w0);
      }
      tc_fence_after();
      {
        TT_PROF_T0(l0);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(trow + as * kTcTkNT + c * 32, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
        tmem_ld_wait();
        TT_PROF_ADD(e_ld, l0);
      }
      tc_fence_before();
      mbar_arrive(&my_empty[as]);                // the whole row slice is in registers: hand the stage back
      if (n0 + kTcTkNT > n_end) {                // last (partial) tile: columns beyond the range do not exist
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (n0 + j >= n_end) v[j] = __float_as_uint(-INFINITY);
      }
      float m[16], qm[4];
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const float* f = reinterpret_cast<const float*>(&v[g * 8]);
        const float a = fmaxf(fmaxf(f[0], f[1]), f[2]);
        const float b = fmaxf(fmaxf(f[3], f[4]), f[5]);
        m[g] = fmaxf(fmaxf(a, b), fmaxf(f[6], f[7]));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) qm[i] = fmaxf(fmaxf(m[4 * i], m[4 * i + 1]), fmaxf(m[4 * i + 2], m[4 * i + 3]));
      const float top = fmaxf(fmaxf(qm[0], qm[1]), fmaxf(qm[2], qm[3]));
      if (__any_sync(0xffffffffu, top > thr)) {
        // usually one or two (lane, group) pairs of the warp fire: descend by warp-uniform votes (32-score quads,
        // then 8-score groups) so that the groups nobody needs cost one vote, not a divergent branch each
        bool uq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) uq[i] = __any_sync(0xffffffffu, qm[i] > thr);     // four independent votes: they pipeline
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!uq[i]) continue;
          bool ug[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) ug[g] = __any_sync(0xffffffffu, m[4 * i + g] > thr);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int g = 4 * i + gg;
            if (!ug[gg]) continue;
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
        }
        // A compaction takes as long as dozens of tiles and stalls the CTA's MMA pipeline (the stage this warp
        // holds is not refilled).  All warps fill their buffers at the same rate, so make them compact TOGETHER:
        // one stall per round instead of one per warp.  The warp that must compact bumps a shared epoch; every
        // warp looks at the epoch whenever one of its rows took a candidate (in the regime where compactions
        // matter that is nearly every tile; a tile without candidates pays nothing).
        const bool must = __any_sync(0xffffffffu, pcnt > kTcTkTrig);
        if (must && lane == 0) atomicAdd(const_cast<uint32_t*>(compact_epoch), 1u);
        __syncwarp();
        const uint32_t ep = *compact_epoch;
        if (must || ep != seen_epoch) {
          seen_epoch = ep;
          compact(false);
        }
      }
    }
    compact(true);
#ifdef TT_TOPK_PROFILE
    if (lane == 0) {
      TT_PROF_FLUSH(0, e_wf);
      TT_PROF_FLUSH(1, e_ld);
      TT_PROF_FLUSH(3, e_cp);
      TT_PROF_FLUSH(4, clock64() - e_all);
    }
#endif
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
  if (qtiles < 1) qtiles = 1;                                   // Q == 0 (an empty query batch) must not divide by zero
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

int tt_debug_read_counters(uint64_t* h_out, int32_t n, int32_t reset) {
  TT_CHECK_ARG(h_out && n > 0 && n <= 16, "debug_read_counters: bad arguments");
  unsigned long long tmp[16];
  cudaError_t e = cudaMemcpyFromSymbol(tmp, tc::g_topk_prof, sizeof(tmp));
  if (e != cudaSuccess) return fail(TT_ERR_CUDA, "debug_read_counters: %s", cudaGetErrorString(e));
  for (int i = 0; i < n; ++i) h_out[i] = tmp[i];
  if (reset) {
    memset(tmp, 0, sizeof(tmp));
    e = cudaMemcpyToSymbol(tc::g_topk_prof, tmp, sizeof(tmp));
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "debug_read_counters: %s", cudaGetErrorString(e));
  }
  return TT_OK;
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
  CUtensorMap tq, ti;
  int rc = tc::make_tmap_bf16_2d(&tq, queries_bf16, Q, d, ldq, 128);
  if (rc) return rc;
  rc = tc::make_tmap_bf16_2d(&ti, items_bf16, N, d, ldi, tc::kTcTkNT);
  if (rc) return rc;
  const size_t smem = tc::tc_topk_smem();
  dim3 grid((unsigned)qtiles, (unsigned)splits);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::tc_score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TT_ERR_CUDA, "score_topk_bf16 smem attr (%zu B): %s", smem, cudaGetErrorString(e)); }
    attr_set = true;
  }
  tc::tc_score_topk_kernel<<<grid, tc::kTcTkThreads, smem, s>>>(tq, ti, (int)Q, (int)N, (int)k, (int)per, cs, ci, pend);
  TT_CHECK_LAUNCH("tc_score_topk");
  topk_merge_kernel<<<(unsigned)((Q + 7) / 8), kTkThreads, 0, s>>>(cs, ci, (int)Q, (int)splits, (int)k, item_index_base,
                                                                 out_scores, out_indices);
  TT_CHECK_LAUNCH("topk_merge");
  return TT_OK;
}

}  // extern "C"
