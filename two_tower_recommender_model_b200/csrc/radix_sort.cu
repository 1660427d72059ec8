// Stable LSD radix sort of (u32 key, u32 payload) pairs, 8-bit digits (9-bit digits when that saves a pass:
// 25..27-bit keys, e.g. the 20 M linearised rows of BASELINE configs[1], take 3 passes instead of 4).
//
// Per pass: (1) per-block digit histogram, stored digit-major [256][nb] so one
// linear exclusive scan yields every (digit, block) output base; (2) that scan;
// (3) stable scatter.  Stability inside a block: warp w owns the contiguous
// chunk [w*256, (w+1)*256) of the block's tile and ranks its 8 rows of 32 keys
// in order with __match_any_sync, so rank order == index order.
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
    if (valid) {
      unsigned peers = __match_any_sync(active, d);
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

size_t sort_workspace_bytes(int64_t n) {
  int64_t nb = (n + kSortTile - 1) / kSortTile;
  if (nb < 1) nb = 1;
  size_t hist = align_up((size_t)kMaxRadix * nb * sizeof(int32_t), 256);
  return 2 * hist + 2 * align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256) +
         scan_workspace_bytes((int64_t)kMaxRadix * nb) + 1024;
}

int sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, int64_t n, int key_bits, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
  if (n < 0 || n >= ((int64_t)1 << 31)) return fail(TT_ERR_INVALID, "sort: n out of range");
  if (key_bits < 1 || key_bits > 32) return fail(TT_ERR_INVALID, "sort: key_bits out of range");
  if (n == 0) return TT_OK;
  const int bits = (key_bits + 8) / 9 < (key_bits + 7) / 8 ? 9 : 8;     // 9-bit digits only when they save a pass
  const int passes = (key_bits + bits - 1) / bits;
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
