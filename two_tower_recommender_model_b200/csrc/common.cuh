// Shared helpers for libtt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tt_b200.h"

namespace tt {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

char* last_error_buf();  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define TT_CHECK_ARG(cond, ...)                          \
  do {                                                   \
    if (!(cond)) return ::tt::fail(TT_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern unsigned long long g_kernel_launches;  // every <<<>>> issued by this library

#define TT_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    ++::tt::g_kernel_launches;                                                       \
    cudaError_t e__ = cudaGetLastError(); /* also clears it: errors never leak into later calls */ \
    if (e__ != cudaSuccess)                                                          \
      return ::tt::fail(TT_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__));       \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t used = 0;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (used + bytes > size) return nullptr;
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exclusive scan of int32 (device-wide, 3 launches).  Implemented in kjt.cu.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out /*may be null*/,
                       void* ws, size_t ws_bytes, cudaStream_t stream);

// Radix sort (radix_sort.cu).
size_t sort_workspace_bytes(int64_t n);
constexpr int kSortTileKeys = 2048;                  // keys per sort tile (= kSortTile)
int sort_digit_bits(int key_bits);
bool sort_uses_fused_path(int64_t n);                // small inputs: one kernel per pass, no scan kernels
size_t sort_fused_hist_ints(int64_t n, int key_bits);
int sort_pairs_u32_fused(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                         uint32_t* tmp_k, uint32_t* tmp_v, int64_t n, int key_bits, int32_t* tile_hists, bool prefilled0,
                         cudaStream_t stream);
int sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, int64_t n, int key_bits, void* ws, size_t ws_bytes,
                   cudaStream_t stream);

}  // namespace tt
