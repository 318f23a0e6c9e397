// Multi-job weight-preparation kernel: ONE launch runs a whole phase of the per-layer re-packing / amax / split work that
// follows every optimiser step (and the gradient un-permutation that ends every backward pass) from a device-resident job
// table.  Replaces ~45 launches of 8-15 us each per PPO iteration by 4 (three dependent phases + the un-permutation).
// The job bodies are the very functions the stand-alone kernels run (prep_kernels.cuh), so both paths are bit-identical.
#include <algorithm>

#include "layer_ops.h"
#include "prep_kernels.cuh"

namespace ddrl {

thread_local PrepRecorder* g_prep_rec = nullptr;

__global__ void __launch_bounds__(256) prep_kernel(const PrepJob* __restrict__ jobs, const int* __restrict__ starts, int njobs) {
  // starts[j] <= blockIdx.x < starts[j + 1]
  int lo = 0, hi = njobs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (starts[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const PrepJob& J = jobs[lo];
  const unsigned vb = blockIdx.x - starts[lo], nvb = J.vblocks;
  const int* p = J.i;
  switch (J.type) {
    case PREP_PACK:
      pack_body(static_cast<const float*>(J.a), static_cast<float*>(J.b), p[0], p[1], p[2], p[3], vb, nvb);
      break;
    case PREP_UNPACK:
      unpack_body(static_cast<const float*>(J.a), static_cast<float*>(J.b), p[0], p[1], p[2], p[3], vb, nvb);
      break;
    case PREP_PACK_S2D:
      pack_s2d_body(static_cast<const float*>(J.a), static_cast<float*>(J.b), p[0], p[1], p[2], p[3], p[4], p[5], p[6], vb, nvb);
      break;
    case PREP_PACK_DGRAD:
      pack_dgrad_body(static_cast<const float*>(J.a), static_cast<float*>(J.b), p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8],
                      vb, nvb);
      break;
    case PREP_PACK_DGRAD_FUSED:
      pack_dgrad_fused_body(static_cast<const float*>(J.a), static_cast<float*>(J.b), p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                            p[8], p[9], p[10], p[11], p[12], p[13], p[14], J.total, vb, nvb);
      break;
    case PREP_SPLIT_HILO:
      split_hi_lo_body(static_cast<const float4*>(J.a), static_cast<uint4*>(J.b), static_cast<uint4*>(J.c), J.total, vb, nvb);
      break;
    case PREP_AMAX:
      amax_body(static_cast<const float*>(J.a), J.total, p[0], (long long)p[1], static_cast<unsigned int*>(J.b), vb, nvb);
      break;
    case PREP_SPLIT_F16:
      split_f16_body(static_cast<const float*>(J.a), p[0], p[1], p[2], J.amax, static_cast<__half*>(J.b), static_cast<__half*>(J.c),
                     p[3], static_cast<__half*>(J.d), static_cast<__half*>(J.e), p[4], vb, nvb);
      break;
    case PREP_COPY: {
      const float* src = static_cast<const float*>(J.a);
      float* dst = static_cast<float*>(J.b);
      PREP_FOR(t, J.total) dst[t] = src[t];
      break;
    }
    case PREP_ZERO: {
      float* dst = static_cast<float*>(J.b);
      PREP_FOR(t, J.total) dst[t] = 0.f;
      break;
    }
    default: break;
  }
}

int PrepTable::upload(const std::vector<PrepJob>& jobs) {
  clear();
  njobs = (int)jobs.size();
  if (!njobs) return DDRL_OK;
  std::vector<int> starts(njobs + 1, 0);
  for (int j = 0; j < njobs; ++j) starts[j + 1] = starts[j] + std::max(1, jobs[j].vblocks);
  total_blocks = starts[njobs];
  DDRL_CUDA(cudaMalloc(&dev_jobs, sizeof(PrepJob) * njobs));
  DDRL_CUDA(cudaMalloc(&dev_starts, sizeof(int) * (njobs + 1)));
  DDRL_CUDA(cudaMemcpy(dev_jobs, jobs.data(), sizeof(PrepJob) * njobs, cudaMemcpyHostToDevice));
  DDRL_CUDA(cudaMemcpy(dev_starts, starts.data(), sizeof(int) * (njobs + 1), cudaMemcpyHostToDevice));
  return DDRL_OK;
}
void PrepTable::clear() {
  if (dev_jobs) cudaFree(dev_jobs);
  if (dev_starts) cudaFree(dev_starts);
  dev_jobs = nullptr; dev_starts = nullptr; njobs = 0; total_blocks = 0;
}
int PrepTable::launch(const char* name, cudaStream_t s) const {
  if (!njobs) return DDRL_OK;
  prep_kernel<<<total_blocks, 256, 0, s>>>(static_cast<const PrepJob*>(dev_jobs), dev_starts, njobs);
  DDRL_LAUNCHED(name);
  return DDRL_OK;
}

}  // namespace ddrl
