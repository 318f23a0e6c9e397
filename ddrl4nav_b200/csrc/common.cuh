// Shared helpers for the ddrl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/ddrl_b200.h"

namespace ddrl {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local char g_cuda_err[256];
extern long long g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return DDRL_E_CUDA;
}

#define DDRL_CUDA(call)                                        \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return ::ddrl::cuda_fail(_e, #call); \
  } while (0)

// Optional per-kernel-class device timing (ddrl_prof_start/stop, capi.cu): one event after every launch;
// with the stream kept busy the gap between consecutive events is the kernel's duration.
extern bool g_prof_on;
extern double g_prof_work;      // algorithmic work (flops or bytes) of the NEXT recorded launch
extern bool g_prof_shapes;     // DDRL_PROF_SHAPES=1: GEMM launches are recorded under a name that carries their shape
void prof_record(const char* name);
const char* prof_intern(const std::string& s);
inline void prof_work(double w) { if (g_prof_on) g_prof_work = w; }

// after a <<<>>> launch: count it and surface launch-configuration errors
#define DDRL_LAUNCHED(name)                                       \
  do {                                                            \
    ::ddrl::g_launches++;                                         \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ::ddrl::cuda_fail(_e, name);    \
    if (::ddrl::g_prof_on) ::ddrl::prof_record(name);             \
  } while (0)

// function attributes (dynamic shared-memory opt-in) are per DEVICE: a once-per-process flag would leave a second GPU of the
// same process unconfigured ("invalid argument" at its first launch)
inline int current_device_index() {
  int d = 0;
  return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64 ? d : 0;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  scratch: >= 32 elements of shared memory.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? scratch[lane] : T(0);
    v = warp_sum(v);
  }
  return v;
}

// max|x| of a block -> one atomicMax on the bits of *slot (positive floats order like their bit patterns), and only when it
// can raise the slot: tens of thousands of unconditional atomics on ONE address serialise in L2 (measured: +25 us on a
// 56 us pooling kernel).  Every thread of the block must call it.
__device__ __forceinline__ void amax_commit(unsigned int* slot, float run_max) {
  __shared__ float amax_part[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) amax_part[w] = run_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int k = 0; k < nw; ++k) m = fmaxf(m, amax_part[k]);
    const unsigned int bits = __float_as_uint(m);
    if (m > 0.f && bits > *reinterpret_cast<volatile unsigned int*>(slot)) atomicMax(slot, bits);
  }
}

// ---- deterministic mode (ddrl_set_deterministic / DDRL_DETERMINISTIC=1) ----------------------------------------------------
// Every cross-block floating-point accumulation of the tc3 path (split-K weight gradients, bias gradients, loss sums, the
// gradient norm) adds its block partials with atomicAdd: the ORDER of those adds, hence the last bits of the sum, changes
// from run to run.  In deterministic mode the blocks that add to the same destination take turns in block order: block
// `rank` waits until the destination's turn counter reads `rank`, adds, passes the turn on (the last one resets it).  Blocks
// are dispatched in linear block-id order and the reduction dimension is never the fastest grid dimension, so the block a
// waiter waits for is always resident or done.  Launchers cap the number of parts (chain length) in this mode.
struct DetSeq { unsigned int* ctr; };          // ctr == nullptr: unordered atomics (default)
extern bool g_deterministic;
unsigned int* det_counters(int n);             // n zeroed, self-cleaning turn counters (device pointer; nullptr on failure)
inline DetSeq det_seq(int n) { DetSeq d{nullptr}; if (g_deterministic) d.ctr = det_counters(n); return d; }
constexpr int kDetMaxParts = 64;

__device__ __forceinline__ void det_enter(unsigned int* c, unsigned int rank) {      // ONE thread; follow with a block barrier
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
    if (v != rank) __nanosleep(32);
  } while (v != rank);
}
__device__ __forceinline__ void det_leave(unsigned int* c, unsigned int rank, unsigned int nparts) {   // after a block barrier; ONE thread
  __threadfence();
  const unsigned int nxt = rank + 1 == nparts ? 0u : rank + 1;
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(c), "r"(nxt) : "memory");
}

// the add itself: inside a turn the destination belongs to this block alone, so it is a plain L2 read-modify-write (stores that
// the block barrier + the leader's fence publish, like any grid-synchronisation pattern); outside deterministic mode, atomicAdd
template <typename T>
__device__ __forceinline__ void det_add(bool ordered, T* p, T v) {
  if (ordered) __stcg(p, __ldcg(p) + v);
  else atomicAdd(p, v);
}

// streaming (read-once) loads/stores: keep L1 clean for the data that is reused
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

}  // namespace ddrl
