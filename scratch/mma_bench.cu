// Micro-benchmark of the tcgen05 issue / completion rate on one SM and on the whole chip.
//   kind::tf32 (M=128, K=8) and kind::f16 (M=128, K=16) as a function of N, operand source (SS / TS) and the number of
//   accumulators the stream rotates over.  Single-CTA numbers are cycles per MMA; the chip-wide pass runs one CTA per SM
//   for a fixed number of MMAs and reports the sustained TFLOP/s under the power cap (the denominator `bench.py` uses
//   for `roofline.frac_tf32x3` / `frac_f16x3`).  Output: one JSON object on stdout (profiles/tf32_peak.json).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ddrl4nav_b200/csrc scratch/mma_bench.cu -o scratch/mma_bench
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ddrl4nav_b200/csrc/tc_ptx.cuh"

using namespace ddrl;

template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(ta),
                 "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(ta),
                 "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// out[2*cta] = issue cycles, out[2*cta+1] = cycles until the last MMA completed
template <int KIND, int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = KIND == 0 ? 0x3f800000u : 0x3c003c00u;
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
  {
    // operand columns 448..511 of every lane hold 1.0 (tf32) / (1.0, 1.0) (f16): data-dependent power is part of "sustained"
    uint32_t ones[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ones[i] = KIND == 0 ? 0x3f800000u : 0x3c003c00u;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16) + 448;
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_st16(tl + c * 16, ones);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 1) {
    const uint32_t fmt = KIND == 0 ? 2u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
    constexpr int ACC0 = 0, A0 = 448;          // accumulators from column 0, TMEM A operand at column 448
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
      for (int i = 0; i < 8; ++i) {
        const uint64_t da = umma_desc(a_s + (i & 3) * 32, 16, 1024, 2), db = umma_desc(b_s + (i & 3) * 32, 16, 1024, 2);
        if (TS) mma_ts<KIND>(tm + ACC0 + (i % NACC) * N, tm + A0 + (i & 3) * 8, db, idesc, 0u);
        else mma_ss<KIND>(tm + ACC0 + (i % NACC) * N, da, db, idesc, 0u);
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
          if (TS) mma_ts<KIND>(tm + ACC0 + (j % NACC) * N, tm + A0 + (j & 3) * 8, db, idesc, 1u);
          else mma_ss<KIND>(tm + ACC0 + (j % NACC) * N, da, db, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 1);
    if (lane == 0) {
      const long long t2 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// mbarrier fast-path latency: cycles per try_wait on an already-completed phase, back to back, one thread
__global__ void mbar_lat(long long* out) {
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1);
    mbar_arrive(smem_u32(&bar[0])); mbar_arrive(smem_u32(&bar[1]));
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) mbar_wait(smem_u32(&bar[i & 1]), 0);
    long long t1 = clock64();
    out[0] = (t1 - t0) / 256;
    // two independent probes issued together
    t0 = clock64();
    uint32_t ok = 0;
    for (int i = 0; i < 256; ++i) {
      uint32_t p0, p1;
      asm volatile("{\n.reg .pred P1, P2;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%2], %4;\n"
                   "mbarrier.try_wait.parity.shared::cta.b64 P2, [%3], %4;\nselp.u32 %0, 1, 0, P1;\nselp.u32 %1, 1, 0, P2;\n}"
                   : "=r"(p0), "=r"(p1) : "r"(smem_u32(&bar[0])), "r"(smem_u32(&bar[1])), "r"(0u) : "memory");
      ok += p0 + p1;
    }
    t1 = clock64();
    out[1] = (t1 - t0) / 256;
    out[2] = ok;
  }
}

static long long* g_d;
static int g_first = 1;

template <int KIND, int N, bool TS, int NACC>
void run() {
  const int iters = 200;
  cudaFuncSetAttribute(bench<KIND, N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  bench<KIND, N, TS, NACC><<<1, 128, 64 * 1024>>>(g_d, iters);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, g_d, sizeof(h), cudaMemcpyDeviceToHost);
  // chip-wide sustained: 148 CTAs, ~0.3 s
  const int kk = KIND == 0 ? 8 : 16;
  const double flop_per_mma = 2.0 * 128 * N * kk;
  int big = (int)(0.3 * 1.5e9 / (12.0 * (N / 2 > 32 ? N / 2 : 32)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<KIND, N, TS, NACC><<<148, 128, 64 * 1024>>>(g_d, big / 8);
  cudaEventRecord(e0);
  bench<KIND, N, TS, NACC><<<148, 128, 64 * 1024>>>(g_d, big);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double tflops = 148.0 * big * 12.0 * flop_per_mma / (ms * 1e-3) / 1e12;
  printf("%s  {\"kind\": \"%s\", \"N\": %d, \"a\": \"%s\", \"nacc\": %d, \"issue_clk\": %.1f, \"complete_clk\": %.1f, \"chip_tflops_sustained\": %.1f, \"chip_ms\": %.1f%s}",
         g_first ? "" : ",\n", KIND == 0 ? "tf32" : "f16", N, TS ? "tmem" : "smem", NACC, (double)h[0] / (iters * 12),
         (double)h[1] / (iters * 12), tflops, ms, e == cudaSuccess ? "" : ", \"error\": 1");
  g_first = 0;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
  cudaMalloc(&g_d, 148 * 16);
  long long h[3];
  mbar_lat<<<1, 32>>>(g_d);
  cudaMemcpy(h, g_d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"mbarrier_try_wait_complete_clk\": %lld, \"two_probes_clk\": %lld,\n \"mma\": [\n", h[0], h[1]);
  run<0, 32, true, 2>();
  run<0, 64, false, 2>(); run<0, 64, true, 1>(); run<0, 64, true, 2>();
  run<0, 128, false, 2>(); run<0, 128, true, 2>();
  run<0, 192, true, 2>();
  run<0, 256, false, 1>(); run<0, 256, true, 1>();
  run<1, 32, true, 2>();
  run<1, 64, false, 2>(); run<1, 64, true, 1>(); run<1, 64, true, 2>();
  run<1, 128, false, 2>(); run<1, 128, true, 2>();
  run<1, 192, true, 2>();
  run<1, 256, false, 1>(); run<1, 256, true, 1>();
  printf("\n ]}\n");
  return 0;
}
