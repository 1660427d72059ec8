// EmbeddingBagCollection kernels (HBM-bound gather / scatter work; no tensor cores).
//
//   forward : table-batched jagged gather + sum/mean pooling.  One CTA owns a
//             tile of kBagsPerCta consecutive bags of one feature; the tile's
//             offsets are staged in shared memory; a group of G lanes owns a bag
//             and reads each row with 128-bit loads, UB bags in flight per group.
//   backward: linearised (table,row) keys -> stable radix sort -> every run of
//             equal keys is reduced by the group that sits on the run's head and
//             the row-wise optimizer is applied in place (no dense gradient, no
//             atomics, deterministic summation order).
#include "common.cuh"

namespace tt {

constexpr int kBagsPerCta = 128;
constexpr int kMaxBagsPerCta = 512;     // forward only: one-id bags take larger tiles (fewer, longer-lived CTAs: one wave)
#ifndef TT_EBC_UB1
#define TT_EBC_UB1 4
#endif
constexpr int kEbcThreads = 256;

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load_stream(const float* p) { v = ld_stream_f4(p); }
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void load_cg(const float* p) { v = __ldcg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void atomic_add(float* p) const { atomicAdd(reinterpret_cast<float4*>(p), v); }   // red.global.add.v4.f32
  __device__ __forceinline__ void add(const Vec& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
  __device__ __forceinline__ void add_scaled(const Vec& o, float s) {
    v.x += s * o.v.x; v.y += s * o.v.y; v.z += s * o.v.z; v.w += s * o.v.w;
  }
  __device__ __forceinline__ void div(float s) { v.x /= s; v.y /= s; v.z /= s; v.w /= s; }
  __device__ __forceinline__ void scale(float s) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
  __device__ __forceinline__ float sumsq() const { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
  template <typename F>
  __device__ __forceinline__ void map2(const Vec& a, F f) {  // this = f(this, a) elementwise
    v.x = f(v.x, a.v.x); v.y = f(v.y, a.v.y); v.z = f(v.z, a.v.z); v.w = f(v.w, a.v.w);
  }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load_stream(const float* p) { v = __ldg(p); }
  __device__ __forceinline__ void load(const float* p) { v = *p; }
  __device__ __forceinline__ void load_cg(const float* p) { v = __ldcg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ void atomic_add(float* p) const { atomicAdd(p, v); }
  __device__ __forceinline__ void add(const Vec& o) { v += o.v; }
  __device__ __forceinline__ void add_scaled(const Vec& o, float s) { v += s * o.v; }
  __device__ __forceinline__ void div(float s) { v /= s; }
  __device__ __forceinline__ void scale(float s) { v *= s; }
  __device__ __forceinline__ float sumsq() const { return v * v; }
  template <typename F>
  __device__ __forceinline__ void map2(const Vec& a, F f) { v = f(v, a.v); }
};

// Shape class of a slot: VEC floats per access, G lanes per bag, NV accesses per lane.
struct ShapeClass {
  int vec, g, nv;
};
// `lookup`: the forward takes half as many lanes per row and two vectors per lane for 9..32 vector units (D = 64 / 128 with
// float4): twice as many bags per warp, i.e. twice the row reads in flight per SM at the same occupancy -- measured
// L=20/D=128 5.1 -> 6.1 TB/s, L=20/D=64 5.2 -> 6.2, L=5/D=128 4.1 -> 5.4 (profiles/r02_ebc_lookup_variants.txt).  The
// fused update keeps one vector per lane: its occupancy is register-bound (32 registers), and the same change cost it 50 %.
static bool classify(int dim, bool all_vec4, ShapeClass* out, bool lookup = false) {
  int vec = all_vec4 ? 4 : 1;
  int units = dim / vec;
  if (units * vec != dim || units <= 0) return false;
  static const int gs[9] = {1, 2, 4, 8, 16, 32, 32, 32, 32};
  static const int nvs[9] = {1, 1, 1, 1, 1, 1, 2, 4, 8};
  static const int gs_l[9] = {1, 2, 4, 8, 8, 16, 32, 32, 32};
  static const int nvs_l[9] = {1, 1, 1, 1, 2, 2, 2, 4, 8};
  const int* G = (lookup && vec == 4) ? gs_l : gs;
  const int* NV = (lookup && vec == 4) ? nvs_l : nvs;
  for (int i = 0; i < 9; ++i)
    if (units <= G[i] * NV[i]) {
      *out = {vec, G[i], NV[i]};
      return true;
    }
  return false;
}
__host__ __device__ constexpr int class_id(int vec, int g, int nv) { return vec * 10000 + g * 100 + nv; }

struct SlotClasses {
  int id[TT_MAX_FEATURES];
};

// Row `row` of a [rows, stride] fp32 matrix that is either local (peers.world == 0) or striped
// over the GPUs of the box in blocks of rows_per_peer rows: peers.ptr[s] is rank s's buffer, mapped
// into this process (NVLink peer memory), so a plain ld/st reaches it.
__device__ __forceinline__ float* peer_row(const tt_peer_buffers& peers, float* local_base, int64_t row, int stride) {
  if (peers.world == 0) return local_base + row * stride;
  const int s = (int)(row / peers.rows_per_peer);
  const int64_t b = row - (int64_t)s * peers.rows_per_peer;
  return static_cast<float*>(peers.ptr[s]) + b * stride;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int VEC, int G, int NV>
__global__ void __launch_bounds__(kEbcThreads)
ebc_forward_kernel(const __grid_constant__ tt_ebc_plan plan, const __grid_constant__ SlotClasses cls,
                   const __grid_constant__ tt_peer_buffers peers, const int64_t* __restrict__ values,
                   const int32_t* __restrict__ offsets, float* __restrict__ pooled, int tiles_per_slot,
                   int bags_per_cta) {
  constexpr int NG = kEbcThreads / G;                   // bag groups per CTA
  constexpr int UB = NV >= 4 ? 1 : (NV == 2 ? 2 : TT_EBC_UB1);   // bags in flight per group
  const int slot = blockIdx.x / tiles_per_slot;
  if (cls.id[slot] != class_id(VEC, G, NV)) return;
  const int tile = blockIdx.x - slot * tiles_per_slot;
  const int B = plan.batch_size;
  const int bag0 = tile * bags_per_cta;
  const int nb = min(bags_per_cta, B - bag0);
  if (nb <= 0) return;

  __shared__ int s_off[kMaxBagsPerCta + 1];
  const int32_t* off = offsets + (int64_t)plan.kjt_index[slot] * B + bag0;
  for (int i = threadIdx.x; i <= nb; i += kEbcThreads) s_off[i] = off[i];
  __syncthreads();

  const int D = plan.dim[slot];
  const int units = D / VEC;
  const float* __restrict__ W = static_cast<const float*>(plan.weights[slot]);
  const uint64_t R = (uint64_t)plan.num_rows[slot];
  const bool mean = plan.pooling[slot] == TT_POOL_MEAN;
  const int ocol = plan.out_col[slot];
  const int g = threadIdx.x / G, l = threadIdx.x % G;
  const bool scatter_add = (peers.flags & TT_PEER_SCATTER_ADD) != 0;

  for (int bb = g; bb < nb; bb += NG * UB) {
    int s[UB], len[UB];
    Vec<VEC> acc[UB][NV];
    int maxlen = 0;
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int bi = bb + u * NG;
      s[u] = bi < nb ? s_off[bi] : 0;
      len[u] = bi < nb ? s_off[bi + 1] - s[u] : 0;
      maxlen = max(maxlen, len[u]);
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[u][v].zero();
    }
    if (maxlen <= 1) {
      // one id per bag (the reference's shape): the UB ids first (independent loads), then UB rows in flight per group
      int64_t id[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) id[u] = len[u] > 0 ? values[s[u]] : (int64_t)-1;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if ((uint64_t)id[u] < R) {
          const float* row = W + id[u] * D;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = l + v * G;
            if (c < units) acc[u][v].load_stream(row + c * VEC);
          }
        }
      }
    } else {
      // multi-hot bags: one bag at a time per group; the group reads G ids with one coalesced
      // load, then keeps JU row reads in flight (ids broadcast by shuffle) and accumulates in id
      // order (the oracle's order).
      constexpr int JU = NV == 1 ? 4 : (NV == 2 ? 2 : 1);
      unsigned gmask = 0xffffffffu;
      if constexpr (G < 32) gmask = ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G);
#pragma unroll 1
      for (int u = 0; u < UB; ++u) {
        const int bi = bb + u * NG;
        if (bi >= nb) break;
        const int s0 = s_off[bi], ln = s_off[bi + 1] - s0;
        Vec<VEC> a1[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) a1[v].zero();
        for (int base = 0; base < ln; base += G) {
          const int cnt = min(G, ln - base);
          const long long myid = (l < cnt) ? values[s0 + base + l] : -1;
          for (int j = 0; j < cnt; j += JU) {
            Vec<VEC> x[JU][NV];
            bool ok[JU];
#pragma unroll
            for (int t = 0; t < JU; ++t) {
              const long long id = __shfl_sync(gmask, myid, (j + t) < G ? (j + t) : 0, G);
              ok[t] = (j + t < cnt) && ((uint64_t)id < R);
              if (ok[t]) {
                const float* row = W + id * D;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const int c = l + v * G;
                  if (c < units) x[t][v].load_stream(row + c * VEC);
                }
              }
            }
#pragma unroll
            for (int t = 0; t < JU; ++t) {
              if (ok[t]) {
#pragma unroll
                for (int v = 0; v < NV; ++v)
                  if (l + v * G < units) a1[v].add(x[t][v]);
              }
            }
          }
        }
        if (scatter_add && ln == 0) continue;      // row-wise shard: another rank holds this bag's rows
        float* o = peer_row(peers, pooled, bag0 + bi, plan.out_stride) + ocol;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = l + v * G;
          if (c < units) {
            if (mean && ln > 1) a1[v].div((float)ln);
            if (scatter_add) a1[v].atomic_add(o + c * VEC); else a1[v].store(o + c * VEC);
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int bi = bb + u * NG;
      if (bi < nb && !(scatter_add && len[u] == 0)) {
        float* o = peer_row(peers, pooled, bag0 + bi, plan.out_stride) + ocol;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = l + v * G;
          if (c < units) {
            if (mean && len[u] > 1) acc[u][v].div((float)len[u]);
            if (scatter_add) acc[u][v].atomic_add(o + c * VEC); else acc[u][v].store(o + c * VEC);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward: key construction
// ---------------------------------------------------------------------------
// hist0 != nullptr (small inputs, fused sort): the kernel also counts the first radix digit of every key into the
// [tile][radix] matrix of the sort's first pass and parks the unused tail of an over-allocated values array on
// the sentinel (extra CTAs beyond the bag tiles), so that neither a histogram nor a tail kernel is launched.
__global__ void __launch_bounds__(kEbcThreads)
ebc_backward_keys_kernel(const __grid_constant__ tt_ebc_plan plan, const int64_t* __restrict__ values,
                         const int32_t* __restrict__ offsets, uint32_t* __restrict__ keys,
                         uint32_t* __restrict__ payload, int tiles_per_key, int64_t n,
                         float* __restrict__ adam_step, float beta1, float beta2, float* __restrict__ adam_bc,
                         int32_t* __restrict__ hist0, int radix_mask, int radix) {
  if (adam_step != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    // device-side Adam step counter (no host-computed bias correction in the launch arguments)
    const float st = *adam_step + 1.0f;
    *adam_step = st;
    adam_bc[0] = (float)(1.0 - pow((double)beta1, (double)st));
    adam_bc[1] = (float)(1.0 - pow((double)beta2, (double)st));
  }
  const int B = plan.batch_size;
  const uint32_t sentinel = (uint32_t)plan.total_rows;
  const int bag_ctas = tiles_per_key * plan.num_kjt_keys;
  if ((int)blockIdx.x >= bag_ctas) {
    // values may be over-allocated (KeyedJaggedTensor.from_id_columns keeps capacity F*B and the live count on
    // the device): park the unused tail on the sentinel
    const int64_t live = offsets[(int64_t)plan.num_kjt_keys * B];
    const int extra = gridDim.x - bag_ctas;
    for (int64_t p = live + (int64_t)(blockIdx.x - bag_ctas) * kEbcThreads + threadIdx.x; p < n; p += (int64_t)extra * kEbcThreads) {
      keys[p] = sentinel;
      payload[p] = 0;
      if (hist0 != nullptr) atomicAdd(&hist0[(p / kSortTileKeys) * radix + (sentinel & radix_mask)], 1);
    }
    return;
  }
  const int f = blockIdx.x / tiles_per_key;  // KJT key index
  const int tile = blockIdx.x - f * tiles_per_key;
  const int bag0 = tile * kBagsPerCta;
  const int nb = min(kBagsPerCta, B - bag0);
  if (nb <= 0) return;
  __shared__ int s_off[kBagsPerCta + 1];
  const int32_t* off = offsets + (int64_t)f * B + bag0;
  for (int i = threadIdx.x; i <= nb; i += kEbcThreads) s_off[i] = off[i];
  __syncthreads();
  const int p0 = s_off[0], p1 = s_off[nb];
  const int slot = plan.slot_of_kjt[f];
  for (int p = p0 + threadIdx.x; p < p1; p += kEbcThreads) {
    uint32_t key = sentinel, pay = 0;
    if (slot >= 0) {
      int lo = 0, hi = nb;  // invariant: s_off[lo] <= p < s_off[hi]
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (s_off[mid] <= p) lo = mid; else hi = mid;
      }
      const int64_t id = values[p];
      const bool ok = (uint64_t)id < (uint64_t)plan.num_rows[slot];
      key = ok ? (uint32_t)(plan.row_base[slot] + id) : sentinel;
      pay = (uint32_t)(slot * B + bag0 + lo);
    }
    keys[p] = key;
    payload[p] = pay;
    if (hist0 != nullptr) atomicAdd(&hist0[(p / kSortTileKeys) * radix + (key & radix_mask)], 1);
  }
}

// Generic path: park the unused tail (see above) with a separate small launch.
__global__ void ebc_backward_tail_kernel(const int32_t* __restrict__ offsets, int64_t num_bags, int64_t n,
                                         uint32_t sentinel, uint32_t* __restrict__ keys,
                                         uint32_t* __restrict__ payload) {
  const int64_t live = offsets[num_bags];
  for (int64_t p = live + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    keys[p] = sentinel;
    payload[p] = 0;
  }
}

// ---------------------------------------------------------------------------
// backward: segmented reduce over sorted keys + in-place row-wise optimizer
// ---------------------------------------------------------------------------
#ifndef TT_EBC_UPD_MINB
#define TT_EBC_UPD_MINB 8
#endif
constexpr int kHotSeg = 128;          // positions per segment of the fused backward (power of two)
constexpr int kHotRowFloats = 1024;   // widest embedding row the scratch is sized for

// One row's update from its summed gradient g (held across the G lanes of the group): row-wise Adagrad / Adam, SGD,
// or accumulation into a dense gradient.
template <int VEC, int G, int NV>
__device__ __forceinline__ void apply_row(const tt_ebc_plan& plan, const tt_sparse_optimizer& opt, int slot, int64_t row,
                                          int l, unsigned mask, Vec<VEC> (&g)[NV], const float* __restrict__ adam_bc) {
  const int D = plan.dim[slot];
  const int units = D / VEC;
  if (opt.grad_scale != 0.f && opt.grad_scale != 1.f) {
#pragma unroll
    for (int v = 0; v < NV; ++v) g[v].scale(opt.grad_scale);
  }
  // mean over D of g^2 (group reduction)
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
    if (l + v * G < units) ss += g[v].sumsq();
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(mask, ss, o);
  const float msq = ss / (float)D;

  float* W = static_cast<float*>(plan.weights[slot]) + row * D;
  if (opt.kind == TT_OPT_ROWWISE_ADAGRAD) {
    float* st = static_cast<float*>(plan.state0[slot]) + row;
    const float snew = *st + msq;
    __syncwarp(mask);
    if (l == 0) *st = snew;
    const float stdv = sqrtf(snew) + opt.eps;
    const float nlr = -opt.lr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      if (c < units) {
        Vec<VEC> w;
        w.load(W + c * VEC);
        w.map2(g[v], [nlr, stdv](float wv, float gv) { return wv + (nlr * gv) / stdv; });
        w.store(W + c * VEC);
      }
    }
  } else if (opt.kind == TT_OPT_ROWWISE_ADAM) {
    float* vst = static_cast<float*>(plan.state0[slot]) + row;
    float* M = static_cast<float*>(plan.state1[slot]) + row * D;
    const float vnew = opt.beta2 * (*vst) + (1.0f - opt.beta2) * msq;
    __syncwarp(mask);
    if (l == 0) *vst = vnew;
    const float bc2 = opt.step_dev != nullptr ? adam_bc[1] : opt.bias_correction2;
    const float denom = sqrtf(vnew / bc2) + opt.eps;
    const float b1 = opt.beta1, omb1 = 1.0f - opt.beta1, lr = opt.lr;
    const float bc1 = opt.step_dev != nullptr ? adam_bc[0] : opt.bias_correction1;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      if (c < units) {
        Vec<VEC> m, w;
        m.load(M + c * VEC);
        m.map2(g[v], [b1, omb1](float mv, float gv) { return b1 * mv + omb1 * gv; });
        m.store(M + c * VEC);
        w.load(W + c * VEC);
        w.map2(m, [lr, bc1, denom](float wv, float mv) { return wv - lr * (mv / bc1) / denom; });
        w.store(W + c * VEC);
      }
    }
  } else if (opt.kind == TT_OPT_SGD) {
    const float lr = opt.lr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      if (c < units) {
        Vec<VEC> w;
        w.load(W + c * VEC);
        w.map2(g[v], [lr](float wv, float gv) { return wv - lr * gv; });
        w.store(W + c * VEC);
      }
    }
  } else {  // TT_OPT_DENSE_GRAD: grad[row] += g
    float* Gd = static_cast<float*>(plan.state1[slot]) + row * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      if (c < units) {
        Vec<VEC> w;
        w.load(Gd + c * VEC);
        w.add(g[v]);
        w.store(Gd + c * VEC);
      }
    }
  }
}

template <int VEC, int G, int NV>
__device__ __forceinline__ void update_body(const tt_ebc_plan& plan, const tt_sparse_optimizer& opt,
                                            const uint32_t* __restrict__ keys, const uint32_t* __restrict__ payload,
                                            int64_t n, const int32_t* __restrict__ offsets,
                                            const float* __restrict__ grad_out, const tt_peer_buffers& peers,
                                            float* __restrict__ hot, int hot_stride, const float* __restrict__ adam_bc,
                                            int32_t* __restrict__ worklist);

template <int VEC, int G, int NV>
__global__ void __launch_bounds__(kEbcThreads, NV == 1 ? TT_EBC_UPD_MINB : 1)
ebc_backward_update_kernel(const __grid_constant__ tt_ebc_plan plan,
                           const __grid_constant__ tt_sparse_optimizer opt,
                           const uint32_t* __restrict__ keys, const uint32_t* __restrict__ payload,
                           int64_t n, const int32_t* __restrict__ offsets,
                           const float* __restrict__ grad_out, const __grid_constant__ tt_peer_buffers peers,
                           float* __restrict__ hot, int hot_stride, const float* __restrict__ adam_bc,
                           int32_t* __restrict__ worklist /* null: ebc_backward_hot_apply_kernel finishes hot runs */) {
  update_body<VEC, G, NV>(plan, opt, keys, payload, n, offsets, grad_out, peers, hot, hot_stride, adam_bc, worklist);
  if (worklist == nullptr) return;
  // Small inputs (one launch less): the LAST CTA to finish applies the optimizer for the (few) runs that crossed a
  // segment boundary; their heads queued the boundary index in the worklist.  worklist[0] = ticket, [1] = count.
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&worklist[0], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int count = *reinterpret_cast<volatile int32_t*>(&worklist[1]);
  const int l = threadIdx.x % G;
  unsigned mask = 0xffffffffu;
  if constexpr (G < 32) mask = ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G);
  for (int w = threadIdx.x / G; w < count; w += kEbcThreads / G) {
    const int64_t b = *reinterpret_cast<volatile int32_t*>(&worklist[2 + w]);
    const uint32_t key = keys[b * kHotSeg];
    int slot = 0;
    for (int i = 0; i < plan.num_slots; ++i)
      if ((int64_t)key >= plan.row_base[i] && (int64_t)key < plan.row_base[i] + plan.num_rows[i]) {
        slot = i;
        break;
      }
    const int64_t row = (int64_t)key - plan.row_base[slot];
    const int units = plan.dim[slot] / VEC;
    Vec<VEC> g[NV];
    const float* acc = hot + b * (int64_t)hot_stride;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      g[v].zero();
      if (c < units) g[v].load_cg(acc + c * VEC);
    }
    apply_row<VEC, G, NV>(plan, opt, slot, row, l, mask, g, adam_bc);
  }
}

template <int VEC, int G, int NV>
__device__ __forceinline__ void update_body(const tt_ebc_plan& plan, const tt_sparse_optimizer& opt,
                                            const uint32_t* __restrict__ keys, const uint32_t* __restrict__ payload,
                                            int64_t n, const int32_t* __restrict__ offsets,
                                            const float* __restrict__ grad_out, const tt_peer_buffers& peers,
                                            float* __restrict__ hot, int hot_stride, const float* __restrict__ adam_bc,
                                            int32_t* __restrict__ worklist) {
  constexpr int64_t hot_seg = kHotSeg;
  const int64_t gid = ((int64_t)blockIdx.x * kEbcThreads + threadIdx.x) / G;
  const int l = threadIdx.x % G;
  if (gid >= n) return;
  const uint32_t key = keys[gid];
  if (key >= (uint32_t)plan.total_rows) return;        // sentinel (unused key / bad id)
  // A group works on one SEGMENT of a run: from the run's head, or from a position that is a multiple of hot_seg
  // inside a run, up to the next such boundary.  Bounded work per group whatever the id distribution: a row hit
  // by 100 000 ids of the batch (Zipf) is summed by ~800 groups all over the GPU instead of by one.
  if (gid > 0 && keys[gid - 1] == key && (gid & (hot_seg - 1)) != 0) return;   // neither a head nor on a boundary
  const int64_t lim = min(n, (gid | (hot_seg - 1)) + 1);
  int slot = 0;
  for (int i = 0; i < plan.num_slots; ++i)
    if ((int64_t)key >= plan.row_base[i] && (int64_t)key < plan.row_base[i] + plan.num_rows[i]) {
      slot = i;
      break;
    }
  const int64_t row = (int64_t)key - plan.row_base[slot];
  const int D = plan.dim[slot];
  const int units = D / VEC;
  const int B = plan.batch_size;

  // The weight row (and Adam's first-moment row) is a random DRAM read whose address is known as
  // soon as the key is: start pulling it into L2 now, so its latency overlaps the walk over the
  // run's gradient rows below (prefetch costs no registers, unlike an early load).
#ifdef TT_EBC_UPD_PREFETCH
  {
    const char* wrow = reinterpret_cast<const char*>(static_cast<const float*>(plan.weights[slot]) + row * D);
    const int bytes = D * 4;
    for (int o = l * 128; o < bytes; o += G * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(wrow + o));
    if (opt.kind == TT_OPT_ROWWISE_ADAM || opt.kind == TT_OPT_DENSE_GRAD) {
      const char* mrow = reinterpret_cast<const char*>(static_cast<const float*>(plan.state1[slot]) + row * D);
      for (int o = l * 128; o < bytes; o += G * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow + o));
    }
    if (l == 0 && (opt.kind == TT_OPT_ROWWISE_ADAGRAD || opt.kind == TT_OPT_ROWWISE_ADAM))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const float*>(plan.state0[slot]) + row));
  }
#endif

  Vec<VEC> g[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) g[v].zero();
  unsigned mask = 0xffffffffu;
  int gshift = 0;
  if constexpr (G < 32) {
    gshift = (threadIdx.x & 31) / G * G;
    mask = ((1u << (G & 31)) - 1u) << gshift;
  }
  // geometry of one occurrence: where its gradient row starts and its mean divisor (0 = none)
  auto locate = [&](uint32_t bag, int64_t& goff, float& flen) {
    const int sl = bag / B;
    const int b = bag - sl * B;
    flen = 0.f;
    if (plan.pooling[sl] == TT_POOL_MEAN) {
      const int64_t kb = (int64_t)plan.kjt_index[sl] * B + b;
      const int len = offsets[kb + 1] - offsets[kb];
      if (len > 1) flen = (float)len;
    }
    // "offset" of the gradient row, as a pointer value: the row may live on another GPU
    goff = reinterpret_cast<int64_t>(peer_row(peers, const_cast<float*>(grad_out), b, plan.out_stride) + plan.out_col[sl]);
  };
  auto accumulate = [&](int64_t goff, float flen) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = l + v * G;
      if (c < units) {
        Vec<VEC> x;
        x.load(reinterpret_cast<const float*>(goff) + c * VEC);
        if (flen > 0.f) x.div(flen);
        g[v].add(x);
      }
    }
  };
  {  // the head occurrence
    int64_t goff;
    float flen;
    locate(payload[gid], goff, flen);
    accumulate(goff, flen);
  }
  int64_t j = gid + 1;
  int64_t run_end = j;                                 // one past the last occurrence this group summed
  if (j < lim && keys[j] == key) {
    // The run continues (duplicate ids; hot rows of small tables have thousands of occurrences).
    // The group reads G (key, bag) pairs with one coalesced load, every lane resolves the geometry
    // of ITS occurrence, and the gradient rows are then fetched two at a time in run order: the
    // per-occurrence chain key -> bag -> offsets -> row becomes one round trip per G occurrences.
    constexpr unsigned full = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    bool more = true;
    while (more) {
      const int64_t pos = j + l;
      const bool mine = pos < lim && keys[pos] == key;
      int64_t goff = 0;
      float flen = 0.f;
      if (mine) locate(payload[pos], goff, flen);
      const unsigned same = (__ballot_sync(mask, mine) >> gshift) & full;
      const int cnt = same == full ? G : __ffs(~same) - 1;   // length of the leading run of matches
      more = cnt == G;
      run_end = j + cnt;
      for (int e = 0; e < cnt; e += 2) {
        const int64_t g0 = __shfl_sync(mask, goff, e, G);
        const float f0 = __shfl_sync(mask, flen, e, G);
        const int e1 = e + 1 < G ? e + 1 : e;
        const int64_t g1 = __shfl_sync(mask, goff, e1, G);
        const float f1 = __shfl_sync(mask, flen, e1, G);
        Vec<VEC> x0[NV], x1[NV];
        const bool two = e + 1 < cnt;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = l + v * G;
          if (c < units) {
            x0[v].load(reinterpret_cast<const float*>(g0) + c * VEC);
            if (two) x1[v].load(reinterpret_cast<const float*>(g1) + c * VEC);
          }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (l + v * G < units) {
            if (f0 > 0.f) x0[v].div(f0);
            g[v].add(x0[v]);
            if (two) {
              if (f1 > 0.f) x1[v].div(f1);
              g[v].add(x1[v]);
            }
          }
        }
      }
      j += G;
    }
  }

  // A run that crosses an aligned segment boundary is summed by several groups: every part is added to the
  // run's scratch row (red.global.add) and ebc_backward_hot_apply_kernel applies the optimizer once.
  const bool head = gid == 0 || keys[gid - 1] != key;
  const int64_t seg_end = (gid | (hot_seg - 1)) + 1;
  // only a segment filled up to its boundary can continue in the next one (one extra key load, rarely)
  const bool continues = run_end == seg_end && seg_end < n && keys[seg_end] == key;
  if (head && !continues) {
    apply_row<VEC, G, NV>(plan, opt, slot, row, l, mask, g, adam_bc);
    return;
  }
  int64_t h = gid;                                   // position of the run's head
  if (!head) {
    if (gid > hot_seg && keys[gid - hot_seg - 1] == key) {   // older than the previous segment: lower_bound(key)
      int64_t lo = 0, hi = gid - hot_seg - 1;          // keys[hi] == key
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
      }
      h = lo;
    } else {
      h = gid - hot_seg;                               // any position of the previous segment gives the same scratch row
    }
  }
  float* acc = hot + (h / hot_seg + 1) * (int64_t)hot_stride;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = l + v * G;
    if (c < units) g[v].atomic_add(acc + c * VEC);
  }
  if (worklist != nullptr && head && l == 0) worklist[2 + atomicAdd(&worklist[1], 1)] = (int32_t)(h / hot_seg + 1);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int validate_plan(const tt_ebc_plan* p, bool* all_vec4) {
  if (!p) return fail(TT_ERR_INVALID, "ebc: null plan");
  if (p->num_slots < 0 || p->num_slots > TT_MAX_FEATURES || p->num_kjt_keys < 0 ||
      p->num_kjt_keys > TT_MAX_FEATURES)
    return fail(TT_ERR_INVALID, "ebc: num_slots/num_kjt_keys out of range (max %d)", TT_MAX_FEATURES);
  if (p->batch_size < 0) return fail(TT_ERR_INVALID, "ebc: negative batch");
  if (p->total_rows < 0 || p->total_rows >= 0xffffffffLL)
    return fail(TT_ERR_UNSUPPORTED, "ebc: total_rows must be < 2^32-1");
  *all_vec4 = true;
  for (int s = 0; s < p->num_slots; ++s) {
    if (!p->weights[s]) return fail(TT_ERR_INVALID, "ebc: null weights for slot %d", s);
    if (p->dim[s] <= 0) return fail(TT_ERR_INVALID, "ebc: bad dim for slot %d", s);
    if (p->kjt_index[s] < 0 || p->kjt_index[s] >= p->num_kjt_keys)
      return fail(TT_ERR_INVALID, "ebc: kjt_index out of range for slot %d", s);
    if (p->dim[s] % 4 != 0 || p->out_col[s] % 4 != 0) *all_vec4 = false;
  }
  if (p->out_stride % 4 != 0) *all_vec4 = false;
  return TT_OK;
}

template <typename Launch>
static int for_each_class(const tt_ebc_plan* p, bool all_vec4, SlotClasses* cls, bool lookup, Launch launch) {
  int seen[TT_MAX_FEATURES];
  int nseen = 0;
  for (int s = 0; s < p->num_slots; ++s) {
    ShapeClass c;
    if (!classify(p->dim[s], all_vec4, &c, lookup))
      return fail(TT_ERR_UNSUPPORTED, "ebc: embedding_dim %d not supported (max %d)", p->dim[s],
                  all_vec4 ? 1024 : 256);
    cls->id[s] = class_id(c.vec, c.g, c.nv);
  }
  for (int s = 0; s < p->num_slots; ++s) {
    bool dup = false;
    for (int i = 0; i < nseen; ++i) dup |= seen[i] == cls->id[s];
    if (dup) continue;
    seen[nseen++] = cls->id[s];
    int rc = launch(cls->id[s]);
    if (rc) return rc;
  }
  return TT_OK;
}

#define TT_DISPATCH_CLASS(ID, MACRO) TT_DISPATCH_CLASS_(ID, MACRO, )
#define TT_DISPATCH_CLASS_LOOKUP(ID, MACRO) \
  TT_DISPATCH_CLASS_(ID, MACRO, case class_id(4, 8, 2): MACRO(4, 8, 2); break; case class_id(4, 16, 2): MACRO(4, 16, 2); break;)
#define TT_DISPATCH_CLASS_(ID, MACRO, EXTRA)         \
  switch (ID) {                                      \
    EXTRA                                            \
    case class_id(4, 1, 1): MACRO(4, 1, 1); break;   \
    case class_id(4, 2, 1): MACRO(4, 2, 1); break;   \
    case class_id(4, 4, 1): MACRO(4, 4, 1); break;   \
    case class_id(4, 8, 1): MACRO(4, 8, 1); break;   \
    case class_id(4, 16, 1): MACRO(4, 16, 1); break; \
    case class_id(4, 32, 1): MACRO(4, 32, 1); break; \
    case class_id(4, 32, 2): MACRO(4, 32, 2); break; \
    case class_id(4, 32, 4): MACRO(4, 32, 4); break; \
    case class_id(4, 32, 8): MACRO(4, 32, 8); break; \
    case class_id(1, 1, 1): MACRO(1, 1, 1); break;   \
    case class_id(1, 2, 1): MACRO(1, 2, 1); break;   \
    case class_id(1, 4, 1): MACRO(1, 4, 1); break;   \
    case class_id(1, 8, 1): MACRO(1, 8, 1); break;   \
    case class_id(1, 16, 1): MACRO(1, 16, 1); break; \
    case class_id(1, 32, 1): MACRO(1, 32, 1); break; \
    case class_id(1, 32, 2): MACRO(1, 32, 2); break; \
    case class_id(1, 32, 4): MACRO(1, 32, 4); break; \
    case class_id(1, 32, 8): MACRO(1, 32, 8); break; \
    default: return fail(TT_ERR_UNSUPPORTED, "ebc: no kernel for shape class %d", ID); \
  }

// Runs that crossed a segment boundary: slot b (1-based) holds the complete gradient sum of the run whose FIRST
// crossed boundary is b * hot_seg.  One group per boundary applies the optimizer for it.
template <int VEC, int G, int NV>
__global__ void __launch_bounds__(kEbcThreads)
ebc_backward_hot_apply_kernel(const __grid_constant__ tt_ebc_plan plan, const __grid_constant__ tt_sparse_optimizer opt,
                              const uint32_t* __restrict__ keys, int64_t n, const float* __restrict__ hot, int hot_stride,
                              int64_t num_bounds, const float* __restrict__ adam_bc) {
  constexpr int64_t hot_seg = kHotSeg;
  const int64_t b = ((int64_t)blockIdx.x * kEbcThreads + threadIdx.x) / G + 1;
  const int l = threadIdx.x % G;
  if (b > num_bounds) return;
  const int64_t q = b * hot_seg;
  if (q >= n) return;
  const uint32_t key = keys[q];
  if (key >= (uint32_t)plan.total_rows || keys[q - 1] != key) return;           // no run crosses this boundary
  if (b > 1 && keys[q - hot_seg - 1] == key) return;                            // the run crossed an earlier boundary first
  int slot = 0;
  for (int i = 0; i < plan.num_slots; ++i)
    if ((int64_t)key >= plan.row_base[i] && (int64_t)key < plan.row_base[i] + plan.num_rows[i]) {
      slot = i;
      break;
    }
  const int64_t row = (int64_t)key - plan.row_base[slot];
  const int units = plan.dim[slot] / VEC;
  unsigned mask = 0xffffffffu;
  if constexpr (G < 32) mask = ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G);
  Vec<VEC> g[NV];
  const float* acc = hot + b * (int64_t)hot_stride;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = l + v * G;
    g[v].zero();
    if (c < units) g[v].load(acc + c * VEC);
  }
  apply_row<VEC, G, NV>(plan, opt, slot, row, l, mask, g, adam_bc);
}


// ---------------------------------------------------------------------------
// dedup (parity surface): unique linearised (table,row) keys of a batch, ascending, with counts and the inverse
// map -- torch.unique(sorted=True, return_inverse=True, return_counts=True) -- built from the SAME key
// construction and radix sort the fused backward uses.
// ---------------------------------------------------------------------------
__global__ void dedup_iota_kernel(uint32_t* __restrict__ payload, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    payload[i] = (uint32_t)i;
}
__global__ void dedup_flag_kernel(const uint32_t* __restrict__ skeys, int64_t n, uint32_t sentinel, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (skeys[i] < sentinel && (i == 0 || skeys[i - 1] != skeys[i])) ? 1 : 0;
}
__global__ void dedup_emit_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ spos,
                                  const int32_t* __restrict__ flag, const int32_t* __restrict__ excl, int64_t n,
                                  uint32_t sentinel, int64_t* __restrict__ unique_keys, int32_t* __restrict__ counts,
                                  int32_t* __restrict__ inverse) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = skeys[i];
    if (key >= sentinel) {
      inverse[spos[i]] = -1;                           // id outside its table: no row
      continue;
    }
    const int u = excl[i] + flag[i] - 1;               // index of this key among the unique keys
    if (flag[i]) unique_keys[u] = (int64_t)key;
    atomicAdd(&counts[u], 1);
    inverse[spos[i]] = u;
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

static int check_peers(const tt_peer_buffers* p, tt_peer_buffers* out, bool* all_vec4) {
  memset(out, 0, sizeof(*out));
  if (!p) return TT_OK;
  if (p->world < 1 || p->world > TT_MAX_PEERS || p->rows_per_peer < 1) return fail(TT_ERR_INVALID, "peer buffers: bad world / rows_per_peer");
  if ((p->flags & ~TT_PEER_SCATTER_ADD) != 0) return fail(TT_ERR_INVALID, "peer buffers: unknown flags");
  for (int i = 0; i < p->world; ++i) {
    if (!p->ptr[i]) return fail(TT_ERR_INVALID, "peer buffers: null pointer for rank %d", i);
    if ((reinterpret_cast<uintptr_t>(p->ptr[i]) & 15) != 0) *all_vec4 = false;
  }
  *out = *p;
  return TT_OK;
}

static int ebc_forward_impl(const tt_ebc_plan* h_plan, const int64_t* values, const int32_t* offsets,
                            float* pooled, const tt_peer_buffers* h_peers, void* stream) {
  bool all_vec4;
  int rc = validate_plan(h_plan, &all_vec4);
  if (rc) return rc;
  tt_peer_buffers peers;
  rc = check_peers(h_peers, &peers, &all_vec4);
  if (rc) return rc;
  TT_CHECK_ARG(offsets && (pooled || h_peers), "ebc_forward: null pointer");
  if (h_plan->num_slots == 0 || h_plan->batch_size == 0) return TT_OK;
  if (!h_peers && (reinterpret_cast<uintptr_t>(pooled) & 15) != 0) all_vec4 = false;
  cudaStream_t s = as_stream(stream);
  static const int env_bags = [] { const char* v = getenv("TT_EBC_FWD_BAGS"); return v ? atoi(v) : 0; }();
  int bags = kBagsPerCta;
  if (env_bags >= 16 && env_bags <= kMaxBagsPerCta) bags = env_bags;
  const int tiles = (h_plan->batch_size + bags - 1) / bags;
  const unsigned grid = (unsigned)(tiles * h_plan->num_slots);
  SlotClasses cls;
  return for_each_class(h_plan, all_vec4, &cls, true, [&](int id) -> int {
#define TT_LAUNCH_FWD(V, G, N)                                                                       \
  ebc_forward_kernel<V, G, N><<<grid, kEbcThreads, 0, s>>>(*h_plan, cls, peers, values, offsets, pooled, tiles, bags)
    TT_DISPATCH_CLASS_LOOKUP(id, TT_LAUNCH_FWD);
#undef TT_LAUNCH_FWD
    TT_CHECK_LAUNCH("ebc_forward");
    return TT_OK;
  });
}

int tt_ebc_forward(const tt_ebc_plan* h_plan, const int64_t* values, const int32_t* offsets,
                   float* pooled, void* stream) {
  return ebc_forward_impl(h_plan, values, offsets, pooled, nullptr, stream);
}

int tt_ebc_forward_peer(const tt_ebc_plan* h_plan, const int64_t* values, const int32_t* offsets,
                        const tt_peer_buffers* h_peers, void* stream) {
  TT_CHECK_ARG(h_peers != nullptr, "ebc_forward_peer: null peer buffers");
  return ebc_forward_impl(h_plan, values, offsets, nullptr, h_peers, stream);
}

size_t tt_ebc_backward_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return 4 * align_up((size_t)n * 4, 256) + sort_workspace_bytes(n) + 2048 + 256 /* adam bias corrections */ +
         2 * align_up((size_t)(n / kHotSeg + 2) * kHotRowFloats * 4, 256) + align_up((size_t)(n / kHotSeg + 8) * 4, 256);
}

static int ebc_backward_impl(const tt_ebc_plan* h_plan, const tt_sparse_optimizer* h_opt,
                             const int64_t* values, int64_t n, const int32_t* offsets,
                             const float* grad_out, const tt_peer_buffers* h_peers, void* ws, size_t ws_bytes,
                             void* stream) {
  bool all_vec4;
  int rc = validate_plan(h_plan, &all_vec4);
  if (rc) return rc;
  tt_peer_buffers peers;
  rc = check_peers(h_peers, &peers, &all_vec4);
  if (rc) return rc;
  TT_CHECK_ARG(h_opt && offsets && (grad_out || h_peers) && n >= 0, "ebc_backward: bad args");
  TT_CHECK_ARG(h_opt->kind >= TT_OPT_DENSE_GRAD && h_opt->kind <= TT_OPT_SGD, "ebc_backward: bad optimizer kind");
  TT_CHECK_ARG(h_opt->grad_scale >= 0.0f, "ebc_backward: negative grad_scale");
  if (n == 0 || h_plan->num_slots == 0 || h_plan->batch_size == 0) return TT_OK;
  if (n >= ((int64_t)1 << 31)) return fail(TT_ERR_UNSUPPORTED, "ebc_backward: too many ids");
  if ((int64_t)h_plan->num_slots * h_plan->batch_size >= ((int64_t)1 << 32))
    return fail(TT_ERR_UNSUPPORTED, "ebc_backward: slots*batch must be < 2^32");
  for (int sl = 0; sl < h_plan->num_slots; ++sl) {
    if (h_opt->kind == TT_OPT_ROWWISE_ADAGRAD || h_opt->kind == TT_OPT_ROWWISE_ADAM)
      TT_CHECK_ARG(h_plan->state0[sl] != nullptr, "ebc_backward: null state0 for slot %d", sl);
    if (h_opt->kind == TT_OPT_ROWWISE_ADAM || h_opt->kind == TT_OPT_DENSE_GRAD)
      TT_CHECK_ARG(h_plan->state1[sl] != nullptr, "ebc_backward: null state1 for slot %d", sl);
  }
  if (!h_peers && (reinterpret_cast<uintptr_t>(grad_out) & 15) != 0) all_vec4 = false;
  cudaStream_t s = as_stream(stream);
  Workspace w(ws, ws_bytes);
  uint32_t* keys = w.take<uint32_t>(n);
  uint32_t* payload = w.take<uint32_t>(n);
  uint32_t* skeys = w.take<uint32_t>(n);
  uint32_t* spayload = w.take<uint32_t>(n);
  float* hot = w.take<float>((size_t)(n / kHotSeg + 2) * kHotRowFloats);   // generic path (the fused path carves its own, zeroed with the rest)
  float* adam_bc = w.take<float>(2);
  if (!keys || !payload || !skeys || !spayload || !hot || !adam_bc) return fail(TT_ERR_WORKSPACE, "ebc_backward: workspace too small");

  int key_bits = 1;
  while (key_bits < 32 && ((uint64_t)h_plan->total_rows >> key_bits) != 0) ++key_bits;  // sentinel == total_rows
  const int tiles = (h_plan->batch_size + kBagsPerCta - 1) / kBagsPerCta;
  const bool fused_sort = sort_uses_fused_path(n);
  int max_dim0 = 0;
  for (int sl = 0; sl < h_plan->num_slots; ++sl) max_dim0 = h_plan->dim[sl] > max_dim0 ? h_plan->dim[sl] : max_dim0;
  const int64_t num_bounds0 = (n - 1) / kHotSeg;
  int32_t* worklist = nullptr;
  if (fused_sort) {
    // Small inputs: 5 launches + 1 memset in all (keys+histogram+tail | one kernel per sort pass | update+hot apply).
    // One memset clears the sort's digit-count matrices, the hot-row scratch and the worklist header together.
    const size_t hist_ints = sort_fused_hist_ints(n, key_bits);
    const size_t hot_floats = (size_t)(num_bounds0 + 1) * ((max_dim0 + 3) / 4 * 4);
    Workspace wz(w.base + w.used, w.size - w.used);
    int32_t* th = wz.take<int32_t>(hist_ints);
    float* hot_z = wz.take<float>(hot_floats > 0 ? hot_floats : 1);
    worklist = wz.take<int32_t>((size_t)num_bounds0 + 4);
    uint32_t* tk = wz.take<uint32_t>(n);
    uint32_t* tv = wz.take<uint32_t>(n);
    if (!th || !hot_z || !worklist || !tk || !tv) return fail(TT_ERR_WORKSPACE, "ebc_backward: workspace too small");
    hot = hot_z;
    const size_t zero_bytes = reinterpret_cast<char*>(worklist) + 8 - reinterpret_cast<char*>(th);   // hists, scratch, ticket + count
    cudaError_t e = cudaMemsetAsync(th, 0, zero_bytes, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "ebc_backward memset: %s", cudaGetErrorString(e));
    const int bits = sort_digit_bits(key_bits);
    ebc_backward_keys_kernel<<<(unsigned)(tiles * h_plan->num_kjt_keys + 16), kEbcThreads, 0, s>>>(
        *h_plan, values, offsets, keys, payload, tiles, n,
        h_opt->kind == TT_OPT_ROWWISE_ADAM ? h_opt->step_dev : nullptr, h_opt->beta1, h_opt->beta2, adam_bc,
        th, (1 << bits) - 1, 1 << bits);
    TT_CHECK_LAUNCH("ebc_backward_keys");
    rc = sort_pairs_u32_fused(keys, payload, skeys, spayload, tk, tv, n, key_bits, th, true, s);
    if (rc) return rc;
  } else {
    ebc_backward_keys_kernel<<<(unsigned)(tiles * h_plan->num_kjt_keys), kEbcThreads, 0, s>>>(
        *h_plan, values, offsets, keys, payload, tiles, n,
        h_opt->kind == TT_OPT_ROWWISE_ADAM ? h_opt->step_dev : nullptr, h_opt->beta1, h_opt->beta2, adam_bc, nullptr, 0, 0);
    TT_CHECK_LAUNCH("ebc_backward_keys");
    ebc_backward_tail_kernel<<<64, 256, 0, s>>>(offsets, (int64_t)h_plan->num_kjt_keys * h_plan->batch_size, n,
                                                (uint32_t)h_plan->total_rows, keys, payload);
    TT_CHECK_LAUNCH("ebc_backward_tail");
    rc = sort_pairs_u32(keys, payload, skeys, spayload, n, key_bits, w.base + w.used, w.size - w.used, s);
    if (rc) return rc;
  }

  // one shape class for the whole launch: the widest table decides
  int max_dim = 0;
  for (int sl = 0; sl < h_plan->num_slots; ++sl) max_dim = h_plan->dim[sl] > max_dim ? h_plan->dim[sl] : max_dim;
  ShapeClass c;
  if (!classify(max_dim, all_vec4, &c))
    return fail(TT_ERR_UNSUPPORTED, "ebc_backward: embedding_dim %d not supported", max_dim);
  const int id = class_id(c.vec, c.g, c.nv);
  const int64_t threads = n * c.g;
  const unsigned grid = (unsigned)((threads + kEbcThreads - 1) / kEbcThreads);
  if (max_dim > kHotRowFloats) return fail(TT_ERR_UNSUPPORTED, "ebc_backward: embedding_dim %d > %d", max_dim, kHotRowFloats);
  const int hot_stride = (max_dim + 3) / 4 * 4;
  const int64_t num_bounds = (n - 1) / kHotSeg;          // boundaries kHotSeg, 2*kHotSeg, ... inside [1, n)
  if (num_bounds > 0 && !fused_sort) {
    cudaError_t e = cudaMemsetAsync(hot, 0, (size_t)(num_bounds + 1) * hot_stride * 4, s);
    if (e != cudaSuccess) return fail(TT_ERR_CUDA, "ebc_backward memset: %s", cudaGetErrorString(e));
  }
#define TT_LAUNCH_BWD(V, G, N)                                                                        \
  ebc_backward_update_kernel<V, G, N><<<grid, kEbcThreads, 0, s>>>(*h_plan, *h_opt, skeys, spayload, n, \
                                                                   offsets, grad_out, peers, hot, hot_stride, adam_bc, worklist)
  TT_DISPATCH_CLASS(id, TT_LAUNCH_BWD);
#undef TT_LAUNCH_BWD
  TT_CHECK_LAUNCH("ebc_backward_update");
  if (num_bounds > 0 && !fused_sort) {
    const unsigned hgrid = (unsigned)((num_bounds * c.g + kEbcThreads - 1) / kEbcThreads);
#define TT_LAUNCH_HOT(V, G, N)                                                                              \
  ebc_backward_hot_apply_kernel<V, G, N><<<hgrid, kEbcThreads, 0, s>>>(*h_plan, *h_opt, skeys, n, hot, hot_stride, \
                                                                       num_bounds, adam_bc)
    TT_DISPATCH_CLASS(id, TT_LAUNCH_HOT);
#undef TT_LAUNCH_HOT
    TT_CHECK_LAUNCH("ebc_backward_hot_apply");
  }
  return TT_OK;
}

size_t tt_ebc_dedup_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return 4 * align_up((size_t)n * 4, 256) + 2 * align_up((size_t)(n + 1) * 4, 256) + sort_workspace_bytes(n) +
         scan_workspace_bytes(n) + 2048;
}

int tt_ebc_dedup(const tt_ebc_plan* h_plan, const int64_t* values, int64_t n, const int32_t* offsets,
                 int64_t* unique_keys, int32_t* counts, int32_t* inverse, int32_t* num_unique, void* ws,
                 size_t ws_bytes, void* stream) {
  bool all_vec4;
  int rc = validate_plan(h_plan, &all_vec4);
  if (rc) return rc;
  TT_CHECK_ARG(offsets && unique_keys && counts && inverse && num_unique && n >= 0, "ebc_dedup: bad args");
  cudaStream_t s = as_stream(stream);
  if (n == 0 || h_plan->batch_size == 0) {
    cudaError_t e = cudaMemsetAsync(num_unique, 0, 4, s);
    return e == cudaSuccess ? TT_OK : fail(TT_ERR_CUDA, "ebc_dedup: %s", cudaGetErrorString(e));
  }
  if (n >= ((int64_t)1 << 31)) return fail(TT_ERR_UNSUPPORTED, "ebc_dedup: too many ids");
  Workspace w(ws, ws_bytes);
  uint32_t* keys = w.take<uint32_t>(n);
  uint32_t* pos = w.take<uint32_t>(n);
  uint32_t* skeys = w.take<uint32_t>(n);
  uint32_t* spos = w.take<uint32_t>(n);
  int32_t* flag = w.take<int32_t>(n + 1);
  int32_t* excl = w.take<int32_t>(n + 1);
  if (!keys || !pos || !skeys || !spos || !flag || !excl) return fail(TT_ERR_WORKSPACE, "ebc_dedup: workspace too small");
  const int tiles = (h_plan->batch_size + kBagsPerCta - 1) / kBagsPerCta;
  ebc_backward_keys_kernel<<<(unsigned)(tiles * h_plan->num_kjt_keys + 16), kEbcThreads, 0, s>>>(
      *h_plan, values, offsets, keys, pos, tiles, n, nullptr, 0.f, 0.f, nullptr, nullptr, 0, 0);
  TT_CHECK_LAUNCH("ebc_dedup_keys");
  dedup_iota_kernel<<<148, 256, 0, s>>>(pos, n);
  TT_CHECK_LAUNCH("ebc_dedup_iota");
  int key_bits = 1;
  while (key_bits < 32 && ((uint64_t)h_plan->total_rows >> key_bits) != 0) ++key_bits;
  rc = sort_pairs_u32(keys, pos, skeys, spos, n, key_bits, w.base + w.used, w.size - w.used, s);
  if (rc) return rc;
  dedup_flag_kernel<<<148, 256, 0, s>>>(skeys, n, (uint32_t)h_plan->total_rows, flag);
  TT_CHECK_LAUNCH("ebc_dedup_flag");
  Workspace w2(w.base + w.used, w.size - w.used);
  rc = exclusive_scan_i32(flag, excl, n, num_unique, w2.base, w2.size, s);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)n * 4, s);
  if (e != cudaSuccess) return fail(TT_ERR_CUDA, "ebc_dedup memset: %s", cudaGetErrorString(e));
  dedup_emit_kernel<<<148, 256, 0, s>>>(skeys, spos, flag, excl, n, (uint32_t)h_plan->total_rows, unique_keys, counts, inverse);
  TT_CHECK_LAUNCH("ebc_dedup_emit");
  return TT_OK;
}

int tt_ebc_backward_fused(const tt_ebc_plan* h_plan, const tt_sparse_optimizer* h_opt,
                          const int64_t* values, int64_t n, const int32_t* offsets,
                          const float* grad_out, void* ws, size_t ws_bytes, void* stream) {
  return ebc_backward_impl(h_plan, h_opt, values, n, offsets, grad_out, nullptr, ws, ws_bytes, stream);
}

int tt_ebc_backward_fused_peer(const tt_ebc_plan* h_plan, const tt_sparse_optimizer* h_opt,
                               const int64_t* values, int64_t n, const int32_t* offsets,
                               const tt_peer_buffers* h_peer_grads, void* ws, size_t ws_bytes, void* stream) {
  TT_CHECK_ARG(h_peer_grads != nullptr, "ebc_backward_fused_peer: null peer buffers");
  return ebc_backward_impl(h_plan, h_opt, values, n, offsets, nullptr, h_peer_grads, ws, ws_bytes, stream);
}

}  // extern "C"
