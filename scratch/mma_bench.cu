// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M=128, K=8) as a function of N, operand source (SS / TS) and
// whether consecutive MMAs accumulate into the SAME TMEM accumulator or rotate over several.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ddrl4nav_b200/csrc scratch/mma_bench.cu -o scratch/mma_bench
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ddrl4nav_b200/csrc/tc_ptx.cuh"

using namespace ddrl;

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
      // warm
      for (int i = 0; i < 8; ++i) {
        const uint64_t da = umma_desc(a_s + (i & 3) * 32, 16, 1024, 2), db = umma_desc(b_s + (i & 3) * 32, 16, 1024, 2);
        if (TS) umma_tf32_ts(tm + (i % NACC) * N, tm + 448 + (i & 3) * 8, db, idesc, 0u);
        else umma_tf32(tm + (i % NACC) * N, da, db, idesc, 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (lane == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint64_t da = umma_desc(a_s + (j & 3) * 32, 16, 1024, 2), db = umma_desc(b_s + (j & 3) * 32, 16, 1024, 2);
          if (TS) umma_tf32_ts(tm + (j % NACC) * N, tm + 448 + (j & 3) * 8, db, idesc, 1u);
          else umma_tf32(tm + (j % NACC) * N, da, db, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 1);
    if (lane == 0) {
      const long long t2 = clock64();
      out[0] = t1 - t0;      // issue time
      out[1] = t2 - t0;      // until complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N, bool TS, int NACC>
void run(long long* d) {
  const int iters = 200;
  cudaFuncSetAttribute(bench<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  bench<N, TS, NACC><<<1, 128, 64 * 1024>>>(d, iters);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d %s NACC=%d : issue %.1f clk/MMA, complete %.1f clk/MMA  (floor N/2 = %d)  %s\n", N, TS ? "TS" : "SS", NACC,
         (double)h[0] / (iters * 12), (double)h[1] / (iters * 12), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<32, false, 1>(d); run<32, false, 2>(d); run<32, false, 4>(d);
  run<32, true, 1>(d); run<32, true, 2>(d); run<32, true, 4>(d); run<32, true, 6>(d);
  run<64, false, 1>(d); run<64, false, 2>(d); run<64, false, 3>(d);
  run<64, true, 1>(d); run<64, true, 2>(d); run<64, true, 3>(d); run<64, true, 6>(d);
  run<128, false, 1>(d); run<128, true, 1>(d); run<128, true, 3>(d);
  run<256, false, 1>(d); run<256, true, 1>(d);
  return 0;
}
