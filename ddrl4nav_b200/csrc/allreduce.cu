// Gradient all-reduce of the data-parallel learner over NVLink / NVSwitch peer memory -- an own kernel, not a library call.
//
// Every rank's flat gradient buffer lives in memory that all ranks of the box have mapped (peer pointers; with NVSwitch also
// ONE multicast address that reaches every rank's copy).  Rank r owns slice r of the buffer:
//   barrier A (flags in peer memory): every rank's gradients are complete
//   multicast path:  v = multimem.ld_reduce.add [mc + i]   -- the SWITCH adds the W copies of element i on the way in
//                    multimem.st [mc + i], v               -- and fans the result out to all W copies on the way back
//   peer path:       v = sum_p ld [peer_p + i] in rank order; st [peer_p + i], v for every p
//   barrier B: every slice has been written everywhere
// Each element is reduced exactly once, by its owner, and every rank stores the owner's bits: replicas stay bit-identical.
// Per rank 2 x (count / W) elements cross NVLink (1.7 MB each way for the 13.5 MB Pong gradient at W = 8); NCCL's ring /
// tree kernels, their proxy hand-shakes and their launch cost are out of the iteration.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace ddrl {

constexpr int kArMaxWorld = 8;
constexpr int kArBlocks = 148;        // flag slots per phase: one per block of the widest launch (one block per SM)
constexpr int kArThreads = 1024;

struct ArArgs {
  float* peer[kArMaxWorld];        // this element range in every rank's buffer (peer[rank] = the local one)
  float* mc;                       // multicast address of the same range (nullptr: peer path)
  uint32_t* flags[kArMaxWorld];    // barrier flags in every rank's memory: [2 phases][kArBlocks][kArMaxWorld]
  long long n4;                    // float4 elements
  int rank, world;
  uint32_t seq;                    // strictly increasing per call (> 0)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Block b of every rank meets block b of every other rank: thread t < world posts `seq` into rank t's flag
// [phase][b][my rank] and waits for rank t's post in its own [phase][b][t].  Safe to reuse the slots call after call: a rank
// can only post call n + 1's phase-A flags after it left call n's phase B, which every rank entered after leaving phase A.
__device__ __forceinline__ void ar_barrier(const ArArgs& a, int phase) {
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int t = threadIdx.x;
    const size_t slot = ((size_t)phase * kArBlocks + blockIdx.x) * kArMaxWorld;
    __threadfence_system();
    st_release_sys(a.flags[t] + slot + a.rank, a.seq);
    const uint32_t* mine = a.flags[a.rank] + slot + t;
    while ((int32_t)(ld_acquire_sys(mine) - a.seq) < 0) {}
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kArThreads) peer_allreduce_kernel(ArArgs a) {
  ar_barrier(a, 0);
  const long long chunk = (a.n4 + a.world - 1) / a.world;
  const long long lo = (long long)a.rank * chunk, hi = min(a.n4, lo + chunk);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a.mc != nullptr) {
    // 4 independent element groups in flight per thread: with one 1024-thread block per SM a whole 6.7 MB slice (2 ranks) is
    // on the wire at once -- the transfer is bound by NVLink latency x bytes in flight, not by issue rate
    for (; i + 3 * stride < hi; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = mc_ld_reduce4(a.mc + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) mc_st4(a.mc + 4 * (i + u * stride), v[u]);
    }
    for (; i < hi; i += stride) mc_st4(a.mc + 4 * i, mc_ld_reduce4(a.mc + 4 * i));
  } else {
    // peer path: the owner pulls its slice from every rank (all loads of a group in flight together), adds in rank order,
    // pushes the sum to every rank
    for (; i < hi; i += 2 * stride) {
      const bool two = i + stride < hi;
      float4 v0[kArMaxWorld], v1[kArMaxWorld];
#pragma unroll
      for (int p = 0; p < kArMaxWorld; ++p) {
        if (p < a.world) {
          v0[p] = ld_sys4(a.peer[p] + 4 * i);
          if (two) v1[p] = ld_sys4(a.peer[p] + 4 * (i + stride));
        }
      }
      float4 s0 = v0[0], s1 = v1[0];
#pragma unroll
      for (int p = 1; p < kArMaxWorld; ++p) {
        if (p < a.world) {
          s0.x += v0[p].x; s0.y += v0[p].y; s0.z += v0[p].z; s0.w += v0[p].w;
          if (two) { s1.x += v1[p].x; s1.y += v1[p].y; s1.z += v1[p].z; s1.w += v1[p].w; }
        }
      }
#pragma unroll
      for (int p = 0; p < kArMaxWorld; ++p) {
        if (p < a.world) {
          st_sys4(a.peer[p] + 4 * i, s0);
          if (two) st_sys4(a.peer[p] + 4 * (i + stride), s1);
        }
      }
    }
  }
  ar_barrier(a, 1);
}

}  // namespace ddrl

extern "C" int ddrl_peer_allreduce_flag_bytes(void) {
  return (int)(2 * ddrl::kArBlocks * ddrl::kArMaxWorld * sizeof(uint32_t));
}

extern "C" int ddrl_peer_allreduce_f32(float* const* peer_bufs, float* multicast_buf, void* const* peer_flags, int rank, int world,
                                       int64_t offset, int64_t count, uint32_t seq, void* stream) {
  using namespace ddrl;
  if (!peer_bufs || !peer_flags || world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || offset < 0 || count < 0 || seq == 0)
    return DDRL_E_ARG;
  if (offset % 4 != 0 || count % 4 != 0) return DDRL_E_ARG;            // float4 granularity (16-byte aligned slices)
  if (count == 0) return DDRL_OK;
  ArArgs a{};
  for (int p = 0; p < world; ++p) {
    if (!peer_bufs[p] || !peer_flags[p] || (reinterpret_cast<uintptr_t>(peer_bufs[p]) & 15)) return DDRL_E_ARG;
    a.peer[p] = peer_bufs[p] + offset;
    a.flags[p] = static_cast<uint32_t*>(peer_flags[p]);
  }
  if (multicast_buf && (reinterpret_cast<uintptr_t>(multicast_buf) & 15)) return DDRL_E_ARG;
  a.mc = multicast_buf ? multicast_buf + offset : nullptr;
  a.n4 = count / 4;
  a.rank = rank; a.world = world; a.seq = seq;
  // every rank launches the SAME grid (block b meets block b) and all blocks must be co-resident: at most one per SM.
  // Small buffers take fewer blocks (the barrier cost is per block); DDRL_AR_BLOCKS / DDRL_AR_THREADS override for sweeps.
  static const int env_blocks = [] { const char* e = getenv("DDRL_AR_BLOCKS"); return e ? atoi(e) : 0; }();
  static const int env_threads = [] { const char* e = getenv("DDRL_AR_THREADS"); return e ? atoi(e) : 0; }();
  const int threads = env_threads >= 32 && env_threads <= kArThreads ? (env_threads / 32) * 32 : kArThreads;
  const long long per_rank = (a.n4 + world - 1) / world;
  int blocks = (int)std::min<long long>(kArBlocks, std::max<long long>(1, (per_rank + 4LL * threads - 1) / (4LL * threads)));
  if (env_blocks >= 1 && env_blocks <= kArBlocks) blocks = env_blocks;
  peer_allreduce_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
  DDRL_LAUNCHED("peer_allreduce_kernel");
  return DDRL_OK;
}
