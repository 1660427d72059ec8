// Microbenchmark: TMEM -> register read throughput (tcgen05.ld) on one SM, by warp count and shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu
// The number decides how far the "read S out of TMEM" epilogues (softmax, top-k) can be from the MMA floor:
// a 128x128 fp32 tile is 64 KB of TMEM reads per 128x128x64 bf16 MMA group (256 clk at 4096 MAC/clk/SM).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ uint32_t ld_32x32b(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_32x32b<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  return v[0] ^ v[7];
}
template <>
__device__ __forceinline__ uint32_t ld_32x32b<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  return v[0] ^ v[7];
}
template <>
__device__ __forceinline__ uint32_t ld_32x32b<64>(uint32_t taddr) {
  uint32_t v[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  return v[0] ^ v[7];
}
// 16 lanes x 256 bit, repeated x8: 32 registers per thread (64 columns of 16 lanes... the warp still gets 4 KB)
__device__ __forceinline__ uint32_t ld_16x256b_x8(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  return v[0] ^ v[7];
}

// mode: 16/32/64 = 32x32b.xN; 256 = 16x256b.x8.  Each warp reads 128 columns of its 32-lane quarter per iteration.
template <int MODE, int WAIT_EVERY>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, long long* clk_out, uint32_t* sink_out) {
  uint32_t sink = 0;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 16) {
#pragma unroll
      for (int c = 0; c < 128; c += 16) sink ^= ld_32x32b<16>(base + c);
    } else if (MODE == 32) {
#pragma unroll
      for (int c = 0; c < 128; c += 32) sink ^= ld_32x32b<32>(base + c);
    } else if (MODE == 64) {
#pragma unroll
      for (int c = 0; c < 128; c += 64) sink ^= ld_32x32b<64>(base + c);
    } else {
#pragma unroll
      for (int c = 0; c < 128; c += 64) { sink ^= ld_16x256b_x8(base + c); sink ^= ld_16x256b_x8(base + c + (16u << 16)); }
    }
    if ((it % WAIT_EVERY) == WAIT_EVERY - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) sink_out[threadIdx.x] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_slot));
  }
}

template <int MODE, int wait_every>
static void run(const char* name, int warps, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  const int iters = 4096;
  tmem_read_kernel<MODE, wait_every><<<grid, warps * 32>>>(64, d, nullptr);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  tmem_read_kernel<MODE, wait_every><<<grid, warps * 32>>>(iters, d, nullptr);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[256];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = (double)warps * 32 * 128 * 4 * iters;   // per CTA
  printf("%-14s warps=%2d grid=%3d wait_every=%d: %8lld clk  %7.1f B/clk/SM  %6.1f clk per 16KB warp-read  (%.3f ms, %s)\n", name, warps,
         grid, wait_every, mx, bytes / mx, (double)mx / iters, ms, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int warps : {1, 2, 4, 8, 16}) {
      run<32, 1>("32x32b.x32", warps, grid);
      run<32, 4>("32x32b.x32", warps, grid);
    }
    for (int warps : {4, 8}) {
      run<16, 1>("32x32b.x16", warps, grid);
      run<64, 1>("32x32b.x64", warps, grid);
      run<256, 1>("16x256b.x8", warps, grid);
    }
  }
  return 0;
}
