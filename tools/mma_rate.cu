// Microbenchmark: clocks per tcgen05.mma (M = 128, K = 16, bf16 -> fp32) as issued by ONE thread, by form (SS: A from
// shared memory, TS: A from TMEM), N, operand majorness and accumulator pattern.  Operands are zeros: only timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Itwo_tower_recommender_model_b200/csrc -Iinclude \
//        -o tools/mma_rate tools/mma_rate.cu
// The numbers decide the design of the one-pass softmax backward (N = 64 products) -- see DESIGN.md section 4.
#include <cstdio>

#include "tc_common.cuh"

using namespace tt::tc;

namespace tt { int fail(int code, const char*, ...) { return code; } }

template <int N, bool TS, bool AMN, bool BMN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int ndst, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = idesc_bf16_f32(128, N) | (AMN ? kIdescAMnMajor : 0u) | (BMN ? kIdescBMnMajor : 0u);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      da[k] = AMN ? smem_desc_mn_sw128(a0 + k * 2048, 16384, 1024) : smem_desc_k_sw128(a0) + 2 * k;
      db[k] = BMN ? smem_desc_mn_sw128(b0 + k * 2048, 16384, 1024) : smem_desc_k_sw128(b0) + 2 * k;
    }
    const uint32_t d0 = tm + 256, d1 = ndst > 1 ? tm + 256 + (N < 128 ? 128 : 0) + 0 : d0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {     // straight-line issue: operands are loop invariants
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = (k & 4) ? d1 : d0;
        if (TS) mma_ts(d, tm + 8 * (k & 3), db[k & 3], idesc, 1u);
        else mma_ss(d, da[k & 3], db[k & 3], idesc, 1u);
      }
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, bool TS, bool AMN, bool BMN>
static void run(const char* name, int ndst, long long* dout) {
  const int smem = 98 * 1024 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<N, TS, AMN, BMN><<<148, 128, smem>>>(iters, ndst, dout);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d dst=%d: issue %7.1f clk/mma, complete %7.1f clk/mma (math floor %3d)  %s\n", name, N, ndst, (double)h[0] / iters,
         (double)h[1] / iters, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 16);
  run<128, false, false, false>("SS  A K-major  B K-major", 1, dout);
  run<128, false, false, false>("SS  A K-major  B K-major", 2, dout);
  run<64, false, false, false>("SS  A K-major  B K-major", 1, dout);
  run<64, false, false, false>("SS  A K-major  B K-major", 2, dout);
  run<64, false, false, true>("SS  A K-major  B MN-major", 1, dout);
  run<64, false, true, true>("SS  A MN-major B MN-major", 1, dout);
  run<64, false, true, true>("SS  A MN-major B MN-major", 2, dout);
  run<32, false, false, false>("SS  A K-major  B K-major", 1, dout);
  run<256, false, false, false>("SS  A K-major  B K-major", 1, dout);
  run<128, true, false, false>("TS  A TMEM     B K-major", 1, dout);
  run<64, true, false, false>("TS  A TMEM     B K-major", 1, dout);
  run<64, true, false, false>("TS  A TMEM     B K-major", 2, dout);
  run<64, true, false, true>("TS  A TMEM     B MN-major", 1, dout);
  run<32, true, false, false>("TS  A TMEM     B K-major", 1, dout);
  run<32, true, false, false>("TS  A TMEM     B K-major", 2, dout);
  run<256, true, false, false>("TS  A TMEM     B K-major", 1, dout);
  run<256, true, false, true>("TS  A TMEM     B MN-major", 1, dout);
  return 0;
}
