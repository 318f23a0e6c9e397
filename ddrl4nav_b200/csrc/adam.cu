// K7 -- fused global-norm gradient clip + Adam over the flat parameter buffer.
//
// Replaces torch.nn.utils.clip_grad_norm_(self.parameters(), 0.5)    USTC_lab/nn/ppo.py:115,126
//      and torch.optim.Adam.step() of actor_optim / critic_optim / optim  ppo.py:117,128-129
// (Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad; ppo.py:40-42).
// One flat fp32 buffer in the reference's named_parameters() order; the two optimisers of the
// unshared mode are two contiguous segments with their own learning rate, sharing ONE clip
// coefficient computed from the norm over all parameters (after the all-reduce in a
// data-parallel learner).  HBM-bound: pass 1 reads g (4 B/param); pass 2 reads p,g,m,v and
// writes p,m,v (28 B/param).  No host synchronisation: the norm stays on the device.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "layer_ops.h"

namespace ddrl {

constexpr int kMaxSeg = 8;
struct AdamSegs {
  long long begin[kMaxSeg + 1];
  float step_size[kMaxSeg];   // lr / (1 - beta1^t), computed in double on the host like torch
  int n;
};

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out, DetSeq det) {
  __shared__ double scratch[32];
  double acc = 0.0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = g4[i];
    acc += (double)(x.x * x.x + x.y * x.y) + (double)(x.z * x.z + x.w * x.w);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)(g[i] * g[i]);
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    if (det.ctr) det_enter(det.ctr, blockIdx.x);
    det_add(det.ctr != nullptr, out, acc);
    if (det.ctr) det_leave(det.ctr, blockIdx.x, gridDim.x);
  }
}

struct AdamScalars {
  float beta1_w;     // 1 - beta1  (lerp weight)
  float beta2;
  float one_m_beta2;
  float bc2_sqrt;    // sqrt(1 - beta2^t)
  float eps;
  float max_norm;
  int clip;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float coef, float step_size,
                                          const AdamScalars& sc) {
  g *= coef;                                      // clip_grad_norm_: always multiplied
  m = m + sc.beta1_w * (g - m);                   // exp_avg.lerp_(grad, 1 - beta1)
  v = v * sc.beta2;                               // exp_avg_sq.mul_(beta2)
  v = v + (sc.one_m_beta2 * g) * g;               //            .addcmul_(grad, grad, value=1-beta2)
  const float denom = sqrtf(v) / sc.bc2_sqrt + sc.eps;
  p = p - step_size * (m / denom);                // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        AdamSegs segs, AdamScalars sc, const double* __restrict__ sumsq,
                                                        float* __restrict__ norm_out) {
  float coef = 1.f;
  const float norm = (float)sqrt(*sumsq);
  if (sc.clip) coef = fminf(sc.max_norm / (norm + 1e-6f), 1.0f);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(p + i) & 15) == 0)) {
      float4 pp = *reinterpret_cast<float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float ss[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int s = 0;
        while (s + 1 < segs.n && i + k >= segs.begin[s + 1]) ++s;
        ss[k] = segs.step_size[s];
      }
      adam_elem(pp.x, gg.x, mm.x, vv.x, coef, ss[0], sc);
      adam_elem(pp.y, gg.y, mm.y, vv.y, coef, ss[1], sc);
      adam_elem(pp.z, gg.z, mm.z, vv.z, coef, ss[2], sc);
      adam_elem(pp.w, gg.w, mm.w, vv.w, coef, ss[3], sc);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) {
        int s = 0;
        while (s + 1 < segs.n && j >= segs.begin[s + 1]) ++s;
        float pj = p[j], mj = m[j], vj = v[j];
        adam_elem(pj, g[j], mj, vj, coef, segs.step_size[s], sc);
        p[j] = pj; m[j] = mj; v[j] = vj;
      }
    }
  }
}

// scratch double for the norm of callers that bring none (the stand-alone ddrl_clip_adam entry point): one per device
static double* g_sumsq[64] = {};

static void adam_host_scalars(const long long* seg_begin, const float* seg_lr, int nseg, int step, const ddrl_ppo_hparams* hp,
                              AdamSegs& segs, AdamScalars& sc) {
  segs.n = nseg;
  const double bc1 = 1.0 - pow((double)hp->beta1, (double)step);
  const double bc2 = 1.0 - pow((double)hp->beta2, (double)step);
  for (int i = 0; i < nseg; ++i) {
    segs.begin[i] = seg_begin[i];
    segs.step_size[i] = (float)((double)seg_lr[i] / bc1);
  }
  segs.begin[nseg] = seg_begin[nseg];
  sc.beta1_w = (float)(1.0 - (double)hp->beta1);
  sc.beta2 = hp->beta2;
  sc.one_m_beta2 = (float)(1.0 - (double)hp->beta2);
  sc.bc2_sqrt = (float)sqrt(bc2);
  sc.eps = hp->adam_eps;
  sc.max_norm = hp->max_grad_norm;
  sc.clip = hp->clip_grad;
}

// sumsq_scratch: a device double owned by the caller (per net, so two nets / devices / streams never share it); nullptr
// falls back to a lazily allocated per-device scalar
int clip_adam_launch(float* params, const float* grads, float* m, float* v, long long n, const long long* seg_begin,
                     const float* seg_lr, int nseg, int step, const ddrl_ppo_hparams* hp, float* norm_out,
                     cudaStream_t s, double* sumsq_scratch) {
  if (!sumsq_scratch) {
    int dev = 0;
    DDRL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DDRL_E_ARG;
    if (!g_sumsq[dev]) DDRL_CUDA(cudaMalloc(&g_sumsq[dev], sizeof(double)));
    sumsq_scratch = g_sumsq[dev];
  }
  AdamSegs segs;
  AdamScalars sc;
  adam_host_scalars(seg_begin, seg_lr, nseg, step, hp, segs, sc);
  DDRL_CUDA(cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), s));
  const DetSeq det = det_seq(1);
  const int blocks = (int)std::min<long long>(ceil_div64(n, 256 * 4), det.ctr ? kDetMaxParts : 4 * kNumSMs);
  sumsq_kernel<<<blocks, 256, 0, s>>>(grads, n, sumsq_scratch, det);
  prof_work(4.0 * n);
  DDRL_LAUNCHED("sumsq_kernel");
  const int blocks2 = (int)std::min<long long>(ceil_div64(n, 256 * 4), 8 * kNumSMs);
  clip_adam_kernel<<<blocks2, 256, 0, s>>>(params, grads, m, v, n, segs, sc, sumsq_scratch, norm_out);
  prof_work(28.0 * n);
  DDRL_LAUNCHED("clip_adam_kernel");
  return DDRL_OK;
}

// ---- CUDA-graph support: the clip+Adam kernel node of a captured optimiser step carries the step-dependent scalars
// (lr / (1 - beta1^t), sqrt(1 - beta2^t)) BY VALUE; a replay for another step rewrites exactly those two arguments.
const void* clip_adam_kernel_func() { return reinterpret_cast<const void*>(&clip_adam_kernel); }

int clip_adam_update_node(cudaGraphExec_t exec, cudaGraphNode_t node, const long long* seg_begin, const float* seg_lr, int nseg,
                          int step, const ddrl_ppo_hparams* hp) {
  cudaKernelNodeParams np;
  DDRL_CUDA(cudaGraphKernelNodeGetParams(node, &np));
  AdamSegs segs;
  AdamScalars sc;
  adam_host_scalars(seg_begin, seg_lr, nseg, step, hp, segs, sc);
  void* args[9];
  for (int i = 0; i < 9; ++i) args[i] = np.kernelParams[i];
  args[5] = &segs;
  args[6] = &sc;
  np.kernelParams = args;
  DDRL_CUDA(cudaGraphExecKernelNodeSetParams(exec, node, &np));
  return DDRL_OK;
}

}  // namespace ddrl

extern "C" int ddrl_clip_adam(float* params, float* grads, float* m, float* v, int64_t n, const int64_t* seg_begin_host,
                              const float* seg_lr_host, int nseg, int step, const ddrl_ppo_hparams* hp, float* norm_out,
                              void* stream) {
  using namespace ddrl;
  if (n < 0 || nseg < 1 || nseg > kMaxSeg || step < 1 || !hp || !seg_begin_host || !seg_lr_host) return DDRL_E_ARG;
  if (n == 0) return DDRL_OK;
  if (!params || !grads || !m || !v) return DDRL_E_ARG;
  long long sb[kMaxSeg + 1];
  for (int i = 0; i <= nseg; ++i) sb[i] = seg_begin_host[i];
  return clip_adam_launch(params, grads, m, v, n, sb, seg_lr_host, nseg, step, hp, norm_out, (cudaStream_t)stream, nullptr);
}
