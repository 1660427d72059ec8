// KeyedJaggedTensor bookkeeping kernels: complete cumsum, GPU batch construction,
// permute_2D_sparse_data, block_bucketize_sparse_features.  All integer work,
// bit-exact against oracle/kjt.py.
#include <stdarg.h>

#include "common.cuh"

namespace tt {

unsigned long long g_kernel_launches = 0;
static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------------------
// Device-wide exclusive scan (int32).  out has n+1 entries when `complete`.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048
constexpr int kSingleBlockScanMax = 32768;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*>=33*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int r = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return r;
}

// One block scans everything (n <= kSingleBlockScanMax): a single launch.
__global__ void __launch_bounds__(1024) scan_single_block_kernel(const int32_t* __restrict__ in,
                                                                  int32_t* __restrict__ out, int n,
                                                                  int32_t* total_out, int complete) {
  __shared__ int smem[33];
  int carry = 0;
  // 16 consecutive elements per thread and pass (four 16-byte loads in flight): the radix sort's 256 x 64 block
  // histograms are ONE pass, i.e. one round of loads, one block scan, one round of stores
  constexpr int kItems = 16;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (int base = 0; base < n; base += 1024 * kItems) {
    const int idx = base + threadIdx.x * kItems;
    int v[kItems];
    if (vec && idx + kItems <= n) {
#pragma unroll
      for (int j = 0; j < kItems; j += 4) {
        const int4 t = *reinterpret_cast<const int4*>(in + idx + j);
        v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j) v[j] = (idx + j < n) ? in[idx + j] : 0;
    }
    int tsum = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j) tsum += v[j];
    int total;
    int ex = block_exclusive_scan(tsum, &total, smem) + carry;
    if (vec && idx + kItems <= n) {
#pragma unroll
      for (int j = 0; j < kItems; j += 4) {
        int4 t;
        t.x = ex; t.y = ex + v[j]; t.z = t.y + v[j + 1]; t.w = t.z + v[j + 2];
        ex = t.w + v[j + 3];
        *reinterpret_cast<int4*>(out + idx + j) = t;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        if (idx + j < n) out[idx + j] = ex;
        ex += v[j];
      }
    }
    carry += total;
  }
  if (threadIdx.x == 0) {
    if (complete) out[n] = carry;
    if (total_out) *total_out = carry;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t* __restrict__ in,
                                                                    int32_t* __restrict__ block_sums,
                                                                    int64_t n) {
  __shared__ int smem[33];
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t idx = base + j * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += smem[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ in,
                                                                   const int32_t* __restrict__ block_offsets,
                                                                   int32_t* __restrict__ out, int64_t n,
                                                                   int32_t* total_out, int complete,
                                                                   int num_blocks) {
  __shared__ int smem[33];
  int64_t idx = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int tsum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (idx + j < n) ? in[idx + j] : 0;
    tsum += v[j];
  }
  int total;
  int ex = block_exclusive_scan(tsum, &total, smem) + block_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (idx + j < n) out[idx + j] = ex;
    ex += v[j];
  }
  if (blockIdx.x == num_blocks - 1 && threadIdx.x == 0) {
    int grand = block_offsets[blockIdx.x] + total;
    if (complete) out[n] = grand;
    if (total_out) *total_out = grand;
  }
}

size_t scan_workspace_bytes(int64_t n) {
  int64_t nb = (n + kScanTile - 1) / kScanTile;
  return align_up((size_t)(nb + 1) * sizeof(int32_t), 256) * 2;
}

static int scan_impl(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out, int complete,
                     void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n < 0 || n >= (int64_t)1 << 31) return fail(TT_ERR_INVALID, "scan: n out of range");
  if (n <= kSingleBlockScanMax) {
    scan_single_block_kernel<<<1, 1024, 0, stream>>>(in, out, (int)n, total_out, complete);
    TT_CHECK_LAUNCH("scan_single_block");
    return TT_OK;
  }
  int64_t nb = (n + kScanTile - 1) / kScanTile;
  Workspace w(ws, ws_bytes);
  int32_t* sums = w.take<int32_t>(nb + 1);
  int32_t* offs = w.take<int32_t>(nb + 1);
  if (!sums || !offs) return fail(TT_ERR_WORKSPACE, "scan: workspace too small");
  if (nb > kSingleBlockScanMax) return fail(TT_ERR_UNSUPPORTED, "scan: n too large");
  scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, sums, n);
  TT_CHECK_LAUNCH("scan_reduce");
  scan_single_block_kernel<<<1, 1024, 0, stream>>>(sums, offs, (int)nb, nullptr, 0);
  TT_CHECK_LAUNCH("scan_block_sums");
  scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, offs, out, n, total_out, complete, (int)nb);
  TT_CHECK_LAUNCH("scan_apply");
  return TT_OK;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out, void* ws,
                       size_t ws_bytes, cudaStream_t stream) {
  return scan_impl(in, out, n, total_out, 0, ws, ws_bytes, stream);
}

// ---------------------------------------------------------------------------
// GPU batch construction (transform_to_torchrec_batch, utils/model_training.py:43-61)
// ---------------------------------------------------------------------------
__global__ void columns_lengths_kernel(const int64_t* __restrict__ ids, int32_t* __restrict__ lengths,
                                       int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lengths[i] = ids[i] != 0 ? 1 : 0;
}

__device__ __forceinline__ int64_t python_mod(int64_t id, int64_t R) {
  int64_t r = id % R;
  if (r != 0 && ((r < 0) != (R < 0))) r += R;
  return r;
}

// row-wise shard: 1 where the (modulo) id falls into this rank's row range
__global__ void columns_lengths_range_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ num_embeddings,
                                             const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                                             int32_t* __restrict__ lengths, int64_t batch, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  int len = 0;
  if (id != 0) {
    const int64_t f = i / batch, r = python_mod(id, num_embeddings[f]);
    len = (r >= lo[f] && r < hi[f]) ? 1 : 0;
  }
  lengths[i] = len;
}

__global__ void columns_values_range_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ num_embeddings,
                                            const int64_t* __restrict__ lo, const int32_t* __restrict__ lengths,
                                            const int32_t* __restrict__ offsets, int64_t* __restrict__ values,
                                            int64_t batch, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || lengths[i] == 0) return;
  const int64_t f = i / batch;
  values[offsets[i]] = python_mod(ids[i], num_embeddings[f]) - lo[f];
}

__global__ void columns_values_kernel(const int64_t* __restrict__ ids,
                                      const int64_t* __restrict__ num_embeddings,
                                      const int32_t* __restrict__ offsets, int64_t* __restrict__ values,
                                      int64_t batch, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t id = ids[i];
  if (id == 0) return;
  int64_t R = num_embeddings[i / batch];
  int64_t r = id % R;
  if (r != 0 && ((r < 0) != (R < 0))) r += R;  // Python modulo
  values[offsets[i]] = r;
}

// ---------------------------------------------------------------------------
// Row-wise shard of W gathered KJTs (sync-free multi-hot input dist): every rank contributes its key-major KJT
// (values padded to a fixed capacity, offsets [F*B+1]); the output is this rank's key-major KJT over the GLOBAL
// batch -- bag (f, r*B + b) keeps, in order, the ids of source bag (r, f, b) that fall into [lo[f], hi[f]),
// rebased to the shard.  Equals bucket `rank` of block_bucketize_sparse_features on the concatenated batch.
// One thread per output bag: bags are short (pooling factor ~20), ids of a bag are contiguous.
// ---------------------------------------------------------------------------
__global__ void gathered_range_count_kernel(const int64_t* __restrict__ values, const int32_t* __restrict__ offsets,
                                            const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                                            int32_t* __restrict__ lengths, int64_t W, int64_t F, int64_t B, int64_t cap) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // output bag: f * (W*B) + r * B + b
  if (o >= F * W * B) return;
  const int64_t f = o / (W * B), rb = o - f * (W * B), r = rb / B, b = rb - r * B;
  const int32_t* off = offsets + r * (F * B + 1) + f * B + b;
  const int64_t* v = values + r * cap;
  const int64_t l = lo[f], h = hi[f];
  int cnt = 0;
  for (int p = off[0]; p < off[1]; ++p) {
    const int64_t id = v[p];
    cnt += (id >= l && id < h) ? 1 : 0;
  }
  lengths[o] = cnt;
}

__global__ void gathered_range_scatter_kernel(const int64_t* __restrict__ values, const int32_t* __restrict__ offsets,
                                              const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                                              const int32_t* __restrict__ out_offsets, int64_t* __restrict__ out_values,
                                              int64_t W, int64_t F, int64_t B, int64_t cap) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= F * W * B) return;
  const int64_t f = o / (W * B), rb = o - f * (W * B), r = rb / B, b = rb - r * B;
  const int32_t* off = offsets + r * (F * B + 1) + f * B + b;
  const int64_t* v = values + r * cap;
  const int64_t l = lo[f], h = hi[f];
  int dst = out_offsets[o];
  for (int p = off[0]; p < off[1]; ++p) {
    const int64_t id = v[p];
    if (id >= l && id < h) out_values[dst++] = id - l;
  }
}

// ---------------------------------------------------------------------------
// permute_2D_sparse_data
// ---------------------------------------------------------------------------
__global__ void permute_lengths_kernel(const int32_t* __restrict__ permute,
                                       const int32_t* __restrict__ lengths,
                                       int32_t* __restrict__ out_lengths, int64_t batch, int64_t n_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  int64_t seg = i / batch, b = i - seg * batch;
  out_lengths[i] = lengths[(int64_t)permute[seg] * batch + b];
}

__global__ void permute_values_kernel(const int32_t* __restrict__ permute,
                                      const int32_t* __restrict__ in_offsets,
                                      const int32_t* __restrict__ out_offsets,
                                      const int64_t* __restrict__ values, int64_t* __restrict__ out_values,
                                      int64_t batch) {
  const int seg = blockIdx.y;
  const int64_t src_seg = permute[seg];
  const int64_t src0 = in_offsets[src_seg * batch];
  const int64_t len = (int64_t)in_offsets[(src_seg + 1) * batch] - src0;
  const int64_t dst0 = out_offsets[(int64_t)seg * batch];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len;
       i += (int64_t)gridDim.x * blockDim.x)
    out_values[dst0 + i] = values[src0 + i];
}

// ---------------------------------------------------------------------------
// block_bucketize_sparse_features: one thread owns one input bag (f,b) and with
// it every output bag (w,f,b), so counters need no atomics and order is stable.
// Ids are taken as UNSIGNED, as fbgemm's uindex_t does: an id past block*W (or a
// negative one) goes to bucket id % W with local index id / W -- fbgemm's
// fallback -- so the bucket always lies in [0, W) and nothing is written out of bounds.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bucket_of_id(int64_t id, int64_t block, int64_t W, int64_t* w, int64_t* local) {
  const uint64_t u = (uint64_t)id, ub = (uint64_t)block, uw = (uint64_t)W;
  if (ub != 0 && u < ub * uw) {
    *w = (int64_t)(u / ub);
    *local = (int64_t)(u - (u / ub) * ub);
  } else {
    *w = (int64_t)(u % uw);
    *local = (int64_t)(u / uw);
  }
}

__global__ void bucketize_count_kernel(const int32_t* __restrict__ offsets,
                                       const int64_t* __restrict__ values,
                                       const int64_t* __restrict__ num_rows, int64_t F, int64_t B,
                                       int64_t W, int32_t* __restrict__ new_lengths,
                                       int32_t* __restrict__ bucket_of) {
  int64_t bag = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bag >= F * B) return;
  int64_t f = bag / B, b = bag - f * B;
  int64_t block = (num_rows[f] + W - 1) / W;
  for (int p = offsets[bag]; p < offsets[bag + 1]; ++p) {
    int64_t w, local;
    bucket_of_id(values[p], block, W, &w, &local);
    bucket_of[p] = (int32_t)w;
    new_lengths[(w * F + f) * B + b] += 1;
  }
}

__global__ void bucketize_scatter_kernel(const int32_t* __restrict__ offsets,
                                         const int64_t* __restrict__ values,
                                         const int64_t* __restrict__ num_rows, int64_t F, int64_t B,
                                         int64_t W, const int32_t* __restrict__ bucket_of,
                                         int32_t* __restrict__ cursor, int64_t* __restrict__ new_values,
                                         int64_t* __restrict__ unbucketize) {
  int64_t bag = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bag >= F * B) return;
  int64_t f = bag / B, b = bag - f * B;
  int64_t block = (num_rows[f] + W - 1) / W;
  for (int p = offsets[bag]; p < offsets[bag + 1]; ++p) {
    int64_t w, local;
    bucket_of_id(values[p], block, W, &w, &local);
    int64_t slot = (w * F + f) * B + b;
    int dst = cursor[slot];
    cursor[slot] = dst + 1;
    new_values[dst] = local;
    if (unbucketize) unbucketize[p] = dst;
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_abi_version(void) { return TT_ABI_VERSION; }
const char* tt_last_error(void) { return last_error_buf(); }
const char* tt_build_arch(void) { return "sm_100a"; }
uint64_t tt_kernel_launch_count(void) { return g_kernel_launches; }

size_t tt_kjt_offsets_workspace_bytes(int64_t n) { return scan_workspace_bytes(n) + 256; }

int tt_kjt_lengths_to_offsets(const int32_t* lengths, int32_t* offsets, int64_t n, void* ws,
                              size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(offsets != nullptr && n >= 0, "lengths_to_offsets: bad args");
  TT_CHECK_ARG(n == 0 || lengths != nullptr, "lengths_to_offsets: null lengths");
  return scan_impl(lengths, offsets, n, nullptr, 1, ws, ws_bytes, as_stream(stream));
}

size_t tt_kjt_from_columns_workspace_bytes(int64_t F, int64_t B) { return scan_workspace_bytes(F * B) + 256; }

int tt_kjt_from_columns(const int64_t* ids, const int64_t* num_embeddings, int64_t F, int64_t B,
                        int64_t* values, int32_t* lengths, int32_t* offsets, void* ws,
                        size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(F >= 0 && B >= 0 && lengths && offsets, "from_columns: bad args");
  cudaStream_t s = as_stream(stream);
  int64_t n = F * B;
  if (n > 0) {
    columns_lengths_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, lengths, n);
    TT_CHECK_LAUNCH("columns_lengths");
  }
  int rc = scan_impl(lengths, offsets, n, nullptr, 1, ws, ws_bytes, s);
  if (rc) return rc;
  if (n > 0) {
    columns_values_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, num_embeddings, offsets, values, B, n);
    TT_CHECK_LAUNCH("columns_values");
  }
  return TT_OK;
}

int tt_kjt_from_columns_range(const int64_t* ids, const int64_t* num_embeddings, const int64_t* row_lo,
                              const int64_t* row_hi, int64_t F, int64_t B, int64_t* values, int32_t* lengths,
                              int32_t* offsets, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(F >= 0 && B >= 0 && lengths && offsets && num_embeddings && row_lo && row_hi, "from_columns_range: bad args");
  cudaStream_t s = as_stream(stream);
  int64_t n = F * B;
  if (n > 0) {
    columns_lengths_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, num_embeddings, row_lo, row_hi, lengths, B, n);
    TT_CHECK_LAUNCH("columns_lengths_range");
  }
  int rc = scan_impl(lengths, offsets, n, nullptr, 1, ws, ws_bytes, s);
  if (rc) return rc;
  if (n > 0) {
    columns_values_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, num_embeddings, row_lo, lengths, offsets, values, B, n);
    TT_CHECK_LAUNCH("columns_values_range");
  }
  return TT_OK;
}

size_t tt_kjt_gathered_range_workspace_bytes(int64_t W, int64_t F, int64_t B) { return scan_workspace_bytes(W * F * B) + 256; }

int tt_kjt_gathered_range(const int64_t* values, int64_t capacity, const int32_t* offsets, const int64_t* row_lo,
                          const int64_t* row_hi, int64_t W, int64_t F, int64_t B, int64_t* out_values,
                          int32_t* out_lengths, int32_t* out_offsets, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(W >= 1 && F >= 0 && B >= 0 && capacity >= 0 && offsets && row_lo && row_hi && out_lengths && out_offsets,
               "kjt_gathered_range: bad args");
  cudaStream_t s = as_stream(stream);
  const int64_t n = W * F * B;
  if (n >= ((int64_t)1 << 31) || W * capacity >= ((int64_t)1 << 31)) return fail(TT_ERR_UNSUPPORTED, "kjt_gathered_range: too large");
  if (n > 0) {
    gathered_range_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(values, offsets, row_lo, row_hi, out_lengths, W, F, B, capacity);
    TT_CHECK_LAUNCH("gathered_range_count");
  }
  int rc = scan_impl(out_lengths, out_offsets, n, nullptr, 1, ws, ws_bytes, s);
  if (rc) return rc;
  if (n > 0) {
    gathered_range_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(values, offsets, row_lo, row_hi, out_offsets, out_values,
                                                                             W, F, B, capacity);
    TT_CHECK_LAUNCH("gathered_range_scatter");
  }
  return TT_OK;
}

size_t tt_kjt_permute_workspace_bytes(int64_t T_out, int64_t B) { return scan_workspace_bytes(T_out * B) + 256; }

int tt_kjt_permute_2d(const int32_t* permute, int64_t T_out, int64_t T_in, int64_t B,
                      const int32_t* lengths, const int32_t* in_offsets, const int64_t* values,
                      int32_t* out_lengths, int32_t* out_offsets, int64_t* out_values, void* ws,
                      size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(T_out >= 0 && T_in >= 0 && B >= 0 && out_offsets, "permute_2d: bad args");
  cudaStream_t s = as_stream(stream);
  int64_t n_out = T_out * B;
  if (n_out > 0) {
    permute_lengths_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, s>>>(permute, lengths, out_lengths, B, n_out);
    TT_CHECK_LAUNCH("permute_lengths");
  }
  int rc = scan_impl(out_lengths, out_offsets, n_out, nullptr, 1, ws, ws_bytes, s);
  if (rc) return rc;
  if (n_out > 0 && values && out_values) {
    dim3 grid(64, (unsigned)T_out);
    permute_values_kernel<<<grid, 256, 0, s>>>(permute, in_offsets, out_offsets, values, out_values, B);
    TT_CHECK_LAUNCH("permute_values");
  }
  return TT_OK;
}

size_t tt_kjt_bucketize_workspace_bytes(int64_t F, int64_t B, int64_t W, int64_t num_values) {
  // cursor [W*F*B] + bucket_of [num_values] + scan scratch
  return scan_workspace_bytes(W * F * B) + align_up((size_t)(W * F * B + 1) * 4, 256) +
         align_up((size_t)(num_values + 1) * 4, 256) + 512;
}

int tt_kjt_block_bucketize(const int32_t* lengths, const int32_t* offsets, const int64_t* values,
                           int64_t num_values, const int64_t* num_rows, int64_t F, int64_t B,
                           int64_t W, int32_t* new_lengths, int32_t* new_offsets,
                           int64_t* new_values, int64_t* unbucketize, void* ws, size_t ws_bytes,
                           void* stream) {
  (void)lengths;
  TT_CHECK_ARG(F >= 0 && B >= 0 && W >= 1 && new_lengths && new_offsets && num_values >= 0,
               "block_bucketize: bad args");
  cudaStream_t s = as_stream(stream);
  int64_t n_out = W * F * B;
  Workspace w(ws, ws_bytes);
  int32_t* cursor = w.take<int32_t>(n_out > 0 ? n_out : 1);
  int32_t* bucket_of = w.take<int32_t>(num_values > 0 ? num_values : 1);
  if (!cursor || !bucket_of) return fail(TT_ERR_WORKSPACE, "block_bucketize: workspace too small");
  if (n_out > 0) {
    cudaError_t e = cudaMemsetAsync(new_lengths, 0, (size_t)n_out * 4, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  int64_t bags = F * B;
  if (bags > 0) {
    bucketize_count_kernel<<<(unsigned)((bags + 127) / 128), 128, 0, s>>>(offsets, values, num_rows, F, B, W,
                                                                         new_lengths, bucket_of);
    TT_CHECK_LAUNCH("bucketize_count");
  }
  int rc = scan_impl(new_lengths, new_offsets, n_out, nullptr, 1, w.base + w.used, w.size - w.used, s);
  if (rc) return rc;
  if (bags > 0) {
    cudaError_t e = cudaMemcpyAsync(cursor, new_offsets, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "memcpy: %s", cudaGetErrorString(e));
    bucketize_scatter_kernel<<<(unsigned)((bags + 127) / 128), 128, 0, s>>>(
        offsets, values, num_rows, F, B, W, bucket_of, cursor, new_values, unbucketize);
    TT_CHECK_LAUNCH("bucketize_scatter");
  }
  return TT_OK;
}

}  // extern "C"
