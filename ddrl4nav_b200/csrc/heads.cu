// K4 -- distribution heads: softmax / Categorical re-normalisation / inverse-CDF sampling /
// log-prob, and the Gaussian counterpart.
//
// Replaces  random_choice_prob_index / select_action   USTC_lab/server/utils.py:20-47
//           CategoricalActor._distribution + log_prob     nn/actor.py:90-101
//           GaussionActor._distribution + log_prob        nn/actor.py:58-70
//           sample / log_prob / play-mode lines           server/forward.py:137-144
// The random draw is an INPUT (uniforms / standard normals), so results are reproducible and
// the (probs, u) -> action map is bit-exact with the reference's numpy expression:
//     (p.cumsum(axis=1) > u[:,None]).argmax(axis=1)
// i.e. sequential fp32 prefix sum (__fadd_rn, never reassociated), strict '>', first hit,
// no hit => 0.  These kernels are tiny (<= 8A+16 B per row); one thread per row.
#include "common.cuh"

namespace ddrl {

constexpr float kEps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
constexpr float kLogSqrt2Pi = 0.918938533204672741780329736406f;

__device__ __forceinline__ int inverse_cdf(const float* __restrict__ p, int A, float u) {
  float c = 0.f;
  for (int j = 0; j < A; ++j) {
    c = __fadd_rn(c, p[j]);
    if (c > u) return j;
  }
  return 0;   // numpy argmax of an all-False row
}

__device__ __forceinline__ int argmax_first(const float* __restrict__ p, int A) {
  int best = 0;
  float bv = p[0];
  for (int j = 1; j < A; ++j)
    if (p[j] > bv) { bv = p[j]; best = j; }
  return best;
}

__global__ void sample_probs_kernel(const float* __restrict__ probs, int ld, const float* __restrict__ u, int B, int A,
                                    float* __restrict__ action, float* __restrict__ logp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = probs + (size_t)b * ld;
  if (u) {
    const int a = inverse_cdf(p, A, u[b]);
    action[b] = (float)a;
    if (logp) logp[b] = logf(p[a]);            // server/utils.py:43
  } else {
    action[b] = (float)argmax_first(p, A);     // server/utils.py:45 / server/forward.py:143
    if (logp) logp[b] = 0.f;
  }
}

constexpr int kMaxA = 64;

// logits -> softmax -> Categorical(probs): q = p / sum(p); logits' = log(clamp(q, eps, 1-eps))
__global__ void categorical_head_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ u, int B, int A,
                                        float* __restrict__ action, float* __restrict__ logp,
                                        float* __restrict__ probs_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* x = logits + (size_t)b * ld;
  float q[kMaxA];
  float m = x[0];
  for (int j = 1; j < A; ++j) m = fmaxf(m, x[j]);
  float s = 0.f;
  for (int j = 0; j < A; ++j) { q[j] = expf(x[j] - m); s += q[j]; }
  float s2 = 0.f;
  for (int j = 0; j < A; ++j) { q[j] = q[j] / s; s2 += q[j]; }     // F.softmax (nn/actor.py:94)
  if (u) {
    for (int j = 0; j < A; ++j) q[j] = q[j] / s2;                   // categorical.py:70
    const int a = inverse_cdf(q, A, u[b]);
    action[b] = (float)a;
    logp[b] = logf(fminf(fmaxf(q[a], kEps), 1.f - kEps));
    if (probs_out) for (int j = 0; j < A; ++j) probs_out[(size_t)b * A + j] = q[j];
  } else {
    action[b] = (float)argmax_first(q, A);    // play mode works on the raw softmax output
    logp[b] = 0.f;
    if (probs_out) for (int j = 0; j < A; ++j) probs_out[(size_t)b * A + j] = q[j];
  }
}

__global__ void gaussian_head_kernel(const float* __restrict__ mu, int ld, const float* __restrict__ log_std,
                                     const float* __restrict__ eps, int B, int A, float* __restrict__ action,
                                     float* __restrict__ logp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lp = 0.f;
  for (int j = 0; j < A; ++j) {
    const float m = mu[(size_t)b * ld + j];
    if (eps) {
      const float sd = expf(log_std[j]);
      const float a = __fadd_rn(m, __fmul_rn(sd, eps[(size_t)b * A + j]));   // torch.normal: mul then add
      action[(size_t)b * A + j] = a;
      const float z = a - m;
      lp += -(z * z) / (2.f * (sd * sd)) - logf(sd) - kLogSqrt2Pi;          // normal.py:87-102
    } else {
      action[(size_t)b * A + j] = m;                                         // server/forward.py:141
    }
  }
  logp[b] = lp;
}

}  // namespace ddrl

using namespace ddrl;

extern "C" int ddrl_sample_categorical_probs(const float* probs, int ld, const float* u, int B, int A, float* action,
                                             float* logp, void* stream) {
  if (B < 0 || A < 1 || ld < A) return DDRL_E_ARG;
  if (B == 0) return DDRL_OK;
  if (!probs || !action) return DDRL_E_ARG;
  sample_probs_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(probs, ld, u, B, A, action, logp);
  DDRL_LAUNCHED("sample_probs_kernel");
  return DDRL_OK;
}

extern "C" int ddrl_categorical_head(const float* logits, int ld, const float* u, int B, int A, float* action,
                                     float* logp, float* probs_out, void* stream) {
  if (B < 0 || A < 1 || A > kMaxA || ld < A) return DDRL_E_ARG;
  if (B == 0) return DDRL_OK;
  if (!logits || !action || !logp) return DDRL_E_ARG;
  categorical_head_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, ld, u, B, A, action, logp, probs_out);
  DDRL_LAUNCHED("categorical_head_kernel");
  return DDRL_OK;
}

extern "C" int ddrl_gaussian_head(const float* mu, int ld, const float* log_std, const float* eps, int B, int A,
                                  float* action, float* logp, void* stream) {
  if (B < 0 || A < 1 || ld < A) return DDRL_E_ARG;
  if (B == 0) return DDRL_OK;
  if (!mu || !log_std || !action || !logp) return DDRL_E_ARG;
  gaussian_head_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(mu, ld, log_std, eps, B, A, action, logp);
  DDRL_LAUNCHED("gaussian_head_kernel");
  return DDRL_OK;
}
