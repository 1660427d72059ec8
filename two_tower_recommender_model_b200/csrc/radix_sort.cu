// Stable LSD radix sort of (u32 key, u32 payload) pairs, 8-bit digits (9-bit digits when that saves a pass:
// 25..27-bit keys, e.g. the 20 M linearised rows of BASELINE configs[1], take 3 passes instead of 4).
//
// Per pass: (1) per-block digit histogram, stored digit-major [256][nb] so one
// linear exclusive scan yields every (digit, block) output base; (2) that scan;
// (3) stable scatter.  Stability inside a block: warp w owns the contiguous
// chunk [w*256, (w+1)*256) of the block's tile and ranks its 8 rows of 32 keys
// in order with __match_any_sync, so rank order == index order.
#include <stdlib.h>

#include "common.cuh"

namespace tt {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per block
constexpr int kMaxRadix = 512;

template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int nb,
               int32_t* __restrict__ block_hist) {
  constexpr int kRadix = 1 << BITS;
  __shared__ int hist[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) hist[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * kSortThreads + threadIdx.x;
    if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) block_hist[(int64_t)d * nb + blockIdx.x] = hist[d];
}

// Lanes of `active` that hold the same BITS-bit digit as this lane.  BITS ballots (independent, they pipeline) instead
// of match.any, which serialises over the distinct values of the warp (32 distinct 9-bit digits = 32 rounds).
template <int BITS>
__device__ __forceinline__ unsigned same_digit_lanes(unsigned active, int d) {
  unsigned peers = active;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const unsigned set = __ballot_sync(0xffffffffu, (d >> b) & 1);
    peers &= ((d >> b) & 1) ? set : ~set;
  }
  return peers;
}

template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                  int shift, int nb, const int32_t* __restrict__ scanned_hist) {
  constexpr int kRadix = 1 << BITS;
  __shared__ int warp_cnt[kSortWarps][kRadix];
  __shared__ int digit_base[kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) digit_base[d] = scanned_hist[(int64_t)d * nb + blockIdx.x];
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * kSortTile + warp * (kSortItems * 32);
  uint32_t k[kSortItems], v[kSortItems];
  int rank[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    bool valid = idx < n;
    k[j] = valid ? keys_in[idx] : 0xffffffffu;
    v[j] = valid ? vals_in[idx] : 0u;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    bool valid = idx < n;
    unsigned active = __ballot_sync(0xffffffffu, valid);
    int d = (k[j] >> shift) & (kRadix - 1);
    const unsigned peers = same_digit_lanes<BITS>(active, d);
    if (valid) {
      int before = __popc(peers & ((1u << lane) - 1u));
      int prev = warp_cnt[warp][d];
      rank[j] = prev + before;
      __syncwarp(active);
      if (before == 0) warp_cnt[warp][d] = prev + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps, per digit (thread t owns digits t, t + 256, ...)
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
    int run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      int c = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    if (idx < n) {
      int d = (k[j] >> shift) & (kRadix - 1);
      int64_t dst = (int64_t)digit_base[d] + warp_cnt[warp][d] + rank[j];
      keys_out[dst] = k[j];
      vals_out[dst] = v[j];
    }
  }
}

// ---------------------------------------------------------------------------
// Small inputs (<= kFusedMaxBlocks tiles, e.g. the 131 072 ids of BASELINE configs[1]): launch latency is the cost,
// so a pass is ONE kernel.  The per-tile digit counts of a pass live in a [nb][radix] matrix (tile-major: a
// thread that owns digit d reads column d of every row, coalesced).  Each scatter block derives its own output
// bases from that matrix (exclusive scan over digits of the column totals + the counts of the tiles before it;
// <= 128 x 512 ints = 256 KB of L2 reads per block), ranks its keys as the generic kernel does, scatters, and adds
// the NEXT pass's digit of every key to the next matrix at the tile the key lands in (red.global.add) -- so no
// histogram kernel and no scan kernel exist between passes.  The first matrix is filled by whoever produces
// the keys (rs_hist_tiles_kernel here, or the embedding backward's key builder).
// ---------------------------------------------------------------------------
constexpr int kFusedMaxBlocks = 128;

template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
rs_hist_tiles_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ tile_hist) {
  constexpr int kRadix = 1 << BITS;
  __shared__ int hist[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) hist[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * kSortThreads + threadIdx.x;
    if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) tile_hist[(int64_t)blockIdx.x * kRadix + d] = hist[d];
}

template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
rs_scatter_fused_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                        uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift, int nb,
                        const int32_t* __restrict__ tile_hist, int32_t* __restrict__ next_hist, int next_shift) {
  constexpr int kRadix = 1 << BITS;
  constexpr int kPer = kRadix / kSortThreads;       // digits per thread in the scan over digits (1 or 2)
  constexpr int kVec = kRadix / 32 / 4;             // int4 loads per lane and matrix row (2 at 8 bits, 4 at 9)
  __shared__ int scratch[2][kSortWarps][kRadix];    // phase 1: per-warp partial column sums; phase 2: warp_cnt = scratch[0]
  __shared__ int digit_base[kRadix];
  __shared__ int warp_tot[kSortWarps];
  int (*warp_cnt)[kRadix] = scratch[0];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- output bases of this tile.  Column sums of the [nb][radix] matrix: warp w takes rows w, w + 8, ...; a lane
  // reads its 4 * kVec digits of a row with 128-bit loads.  All loads of a thread are independent (one or two L2
  // round trips in all); a thread-per-digit walk down the rows was a chain of nb dependent-latency loads.
  {
    int tot[4 * kVec], bef[4 * kVec];
#pragma unroll
    for (int e = 0; e < 4 * kVec; ++e) { tot[e] = 0; bef[e] = 0; }
    for (int b = warp; b < nb; b += kSortWarps) {
      const int4* row = reinterpret_cast<const int4*>(tile_hist + (int64_t)b * kRadix) + lane * kVec;
      const bool prior = b < (int)blockIdx.x;
#pragma unroll
      for (int q = 0; q < kVec; ++q) {
        const int4 c = __ldg(row + q);
        tot[4 * q] += c.x; tot[4 * q + 1] += c.y; tot[4 * q + 2] += c.z; tot[4 * q + 3] += c.w;
        if (prior) { bef[4 * q] += c.x; bef[4 * q + 1] += c.y; bef[4 * q + 2] += c.z; bef[4 * q + 3] += c.w; }
      }
    }
#pragma unroll
    for (int e = 0; e < 4 * kVec; ++e) {
      scratch[0][warp][lane * 4 * kVec + e] = tot[e];
      scratch[1][warp][lane * 4 * kVec + e] = bef[e];
    }
  }
  __syncthreads();
  int tot[kPer], before[kPer];
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    const int d = threadIdx.x * kPer + e;            // consecutive digits per thread: the scan below is in digit order
    tot[e] = 0; before[e] = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) { tot[e] += scratch[0][w][d]; before[e] += scratch[1][w][d]; }
  }
  int mine = 0;
#pragma unroll
  for (int e = 0; e < kPer; ++e) mine += tot[e];
  int incl = mine;                                   // inclusive scan over threads (digit order)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();                                   // also: every read of scratch is done
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
  int run = wbase + incl - mine;
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    digit_base[threadIdx.x * kPer + e] = run + before[e];
    run += tot[e];
  }
  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();

  // ---- stable ranks inside the tile (same scheme as rs_scatter_kernel)
  const int64_t base = (int64_t)blockIdx.x * kSortTile + warp * (kSortItems * 32);
  uint32_t k[kSortItems], v[kSortItems];
  int rank[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    bool valid = idx < n;
    k[j] = valid ? keys_in[idx] : 0xffffffffu;
    v[j] = valid ? vals_in[idx] : 0u;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    bool valid = idx < n;
    unsigned active = __ballot_sync(0xffffffffu, valid);
    int d = (k[j] >> shift) & (kRadix - 1);
    const unsigned peers = same_digit_lanes<BITS>(active, d);
    if (valid) {
      int bef = __popc(peers & ((1u << lane) - 1u));
      int prev = warp_cnt[warp][d];
      rank[j] = prev + bef;
      __syncwarp(active);
      if (bef == 0) warp_cnt[warp][d] = prev + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
    int r = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      int c = warp_cnt[w][d];
      warp_cnt[w][d] = r;
      r += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int64_t idx = base + j * 32 + lane;
    if (idx < n) {
      int d = (k[j] >> shift) & (kRadix - 1);
      int64_t dst = (int64_t)digit_base[d] + warp_cnt[warp][d] + rank[j];
      keys_out[dst] = k[j];
      vals_out[dst] = v[j];
      if (next_hist != nullptr)
        atomicAdd(&next_hist[(dst / kSortTile) * kRadix + ((k[j] >> next_shift) & (kRadix - 1))], 1);
    }
  }
}

int sort_digit_bits(int key_bits) { return (key_bits + 8) / 9 < (key_bits + 7) / 8 ? 9 : 8; }   // 9-bit digits only when they save a pass
bool sort_uses_fused_path(int64_t n) {
  static const bool off = getenv("TT_SORT_NO_FUSED") != nullptr;     // A/B switch for tuning
  return !off && (n + kSortTile - 1) / kSortTile <= kFusedMaxBlocks;
}
size_t sort_fused_hist_ints(int64_t n, int key_bits) {
  const int bits = sort_digit_bits(key_bits);
  const int passes = (key_bits + bits - 1) / bits;
  const int64_t nb = (n + kSortTile - 1) / kSortTile;
  return (size_t)passes * nb * (1 << bits);
}

// Fused path.  tile_hists: [passes][nb][radix] int32; the caller zeroed it, and filled matrix 0 when prefilled0.
int sort_pairs_u32_fused(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                         uint32_t* tmp_k, uint32_t* tmp_v, int64_t n, int key_bits, int32_t* tile_hists, bool prefilled0,
                         cudaStream_t stream) {
  const int bits = sort_digit_bits(key_bits);
  const int passes = (key_bits + bits - 1) / bits;
  const int radix = 1 << bits;
  const int nb = (int)((n + kSortTile - 1) / kSortTile);
  if (!prefilled0) {
    if (bits == 9) rs_hist_tiles_kernel<9><<<nb, kSortThreads, 0, stream>>>(keys_in, n, 0, tile_hists);
    else rs_hist_tiles_kernel<8><<<nb, kSortThreads, 0, stream>>>(keys_in, n, 0, tile_hists);
    TT_CHECK_LAUNCH("rs_hist_tiles");
  }
  const uint32_t* src_k = keys_in;
  const uint32_t* src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t* dst_k = to_out ? keys_out : tmp_k;
    uint32_t* dst_v = to_out ? vals_out : tmp_v;
    const int32_t* h = tile_hists + (size_t)p * nb * radix;
    int32_t* hn = p + 1 < passes ? tile_hists + (size_t)(p + 1) * nb * radix : nullptr;
    if (bits == 9) rs_scatter_fused_kernel<9><<<nb, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, p * bits, nb, h, hn, (p + 1) * bits);
    else rs_scatter_fused_kernel<8><<<nb, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, p * bits, nb, h, hn, (p + 1) * bits);
    TT_CHECK_LAUNCH("rs_scatter_fused");
    src_k = dst_k;
    src_v = dst_v;
  }
  return TT_OK;
}

size_t sort_workspace_bytes(int64_t n) {
  int64_t nb = (n + kSortTile - 1) / kSortTile;
  if (nb < 1) nb = 1;
  size_t hist = align_up((size_t)kMaxRadix * nb * sizeof(int32_t), 256);
  return 4 * hist + 2 * align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256) +
         scan_workspace_bytes((int64_t)kMaxRadix * nb) + 1024;
}

int sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, int64_t n, int key_bits, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
  if (n < 0 || n >= ((int64_t)1 << 31)) return fail(TT_ERR_INVALID, "sort: n out of range");
  if (key_bits < 1 || key_bits > 32) return fail(TT_ERR_INVALID, "sort: key_bits out of range");
  if (n == 0) return TT_OK;
  const int bits = sort_digit_bits(key_bits);
  const int passes = (key_bits + bits - 1) / bits;
  if (sort_uses_fused_path(n)) {
    Workspace wf(ws, ws_bytes);
    const size_t ints = sort_fused_hist_ints(n, key_bits);
    int32_t* th = wf.take<int32_t>(ints);
    uint32_t* tk = wf.take<uint32_t>(n);
    uint32_t* tv = wf.take<uint32_t>(n);
    if (!th || !tk || !tv) return fail(TT_ERR_WORKSPACE, "sort: workspace too small");
    cudaError_t e = cudaMemsetAsync(th, 0, ints * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "sort memset: %s", cudaGetErrorString(e));
    return sort_pairs_u32_fused(keys_in, vals_in, keys_out, vals_out, tk, tv, n, key_bits, th, false, stream);
  }
  const int radix = 1 << bits;
  const int nb = (int)((n + kSortTile - 1) / kSortTile);
  Workspace w(ws, ws_bytes);
  int32_t* hist = w.take<int32_t>((size_t)kMaxRadix * nb);
  int32_t* scanned = w.take<int32_t>((size_t)kMaxRadix * nb);
  uint32_t* tmp_k = w.take<uint32_t>(n);
  uint32_t* tmp_v = w.take<uint32_t>(n);
  if (!hist || !scanned || !tmp_k || !tmp_v) return fail(TT_ERR_WORKSPACE, "sort: workspace too small");
  void* scan_ws = w.base + w.used;
  size_t scan_ws_bytes = w.size - w.used;

  const uint32_t* src_k = keys_in;
  const uint32_t* src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t* dst_k = to_out ? keys_out : tmp_k;
    uint32_t* dst_v = to_out ? vals_out : tmp_v;
    const int shift = p * bits;
    if (bits == 9) rs_hist_kernel<9><<<nb, kSortThreads, 0, stream>>>(src_k, n, shift, nb, hist);
    else rs_hist_kernel<8><<<nb, kSortThreads, 0, stream>>>(src_k, n, shift, nb, hist);
    TT_CHECK_LAUNCH("rs_hist");
    int rc = exclusive_scan_i32(hist, scanned, (int64_t)radix * nb, nullptr, scan_ws, scan_ws_bytes, stream);
    if (rc) return rc;
    if (bits == 9) rs_scatter_kernel<9><<<nb, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, shift, nb, scanned);
    else rs_scatter_kernel<8><<<nb, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, shift, nb, scanned);
    TT_CHECK_LAUNCH("rs_scatter");
    src_k = dst_k;
    src_v = dst_v;
  }
  return TT_OK;
}

}  // namespace tt

extern "C" {

size_t tt_sort_pairs_workspace_bytes(int64_t n) { return tt::sort_workspace_bytes(n); }

int tt_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                      uint32_t* vals_out, int64_t n, int32_t key_bits, void* ws, size_t ws_bytes,
                      void* stream) {
  TT_CHECK_ARG(n == 0 || (keys_in && vals_in && keys_out && vals_out), "sort_pairs: null pointer");
  TT_CHECK_ARG(keys_in != keys_out && vals_in != vals_out, "sort_pairs: in-place not supported");
  return tt::sort_pairs_u32(keys_in, vals_in, keys_out, vals_out, n, key_bits, ws, ws_bytes,
                            tt::as_stream(stream));
}

}  // extern "C"
