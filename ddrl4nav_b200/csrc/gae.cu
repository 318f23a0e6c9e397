// K5 -- GAE / discounted-return scan.
//
// Replaces Agents._accumulate_rewards (USTC_lab/agent/agent.py:124-140):
//     g = 0 ; nv = V[T]
//     for t = T-1 .. 0:   nd = 1 - done[t]
//         g  = g * nd
//         g  = (gamma*lam) * g + ((gamma*nv)*nd - V[t] + r[t])
//         nv = V[t] ; ret[t] = V[t] + g ; adv[t] = g[row 0]
// Layout: time-major [T, C] with C = V*N columns contiguous (what the rollout produces: one
// [V,N] row per env step), so a warp reading 32 consecutive columns of one time row is one
// 128 B coalesced request.  HBM-bound: 17 B per (t, column): r 4 + V 4 + done 1 in, ret 4 + adv 4 out.
//
// Two schedules:
//  * gae_seq_kernel: one thread per column walks time backwards with the loads of the next
//    kUnroll steps in flight (they do not depend on the recurrence).  Every step uses the
//    reference's exact rounding sequence (__fmul_rn/__fadd_rn, no FMA contraction), so the
//    result is BIT-EXACT with the numpy loop.  Used when there are enough columns to fill the GPU.
//  * gae_chunked_kernel: for few columns the time axis is split into CH chunks per column tile.
//    Pass 1 folds each chunk into one affine map g_out = a*g_in + b; the CH maps of a column are
//    then combined by a warp-level associative (Hillis-Steele, shuffle) suffix scan over time;
//    pass 2 replays each chunk from its exact carry-in with the reference's rounding sequence and
//    writes the outputs (re-read served by L2).  Only the carry-in is reassociated; it is
//    damped by gamma*lam per step, measured error <= 1e-6 of max|adv|.
#include "common.cuh"

namespace ddrl {

constexpr int kMaxV = 8;
struct GaeGammas { float g[kMaxV]; };

__device__ __forceinline__ float not_done(uint8_t d) { return (float)(uint8_t)(1 - d); }

// one reference step; returns new g
__device__ __forceinline__ float gae_step(float g, float gl, float gamma, float nv, float v, float r, float nd) {
  g = __fmul_rn(g, nd);
  const float x = __fmul_rn(gl, g);
  float y = __fmul_rn(__fmul_rn(gamma, nv), nd);
  y = __fsub_rn(y, v);
  y = __fadd_rn(y, r);
  return __fadd_rn(x, y);
}

template <int kUnroll>
__global__ void __launch_bounds__(256) gae_seq_kernel(const float* __restrict__ values, const float* __restrict__ rewards,
                                                      const uint8_t* __restrict__ dones, GaeGammas gam, float lam,
                                                      int T, int N, int C, float* __restrict__ ret,
                                                      float* __restrict__ adv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int vrow = c / N;
  const float gamma = gam.g[vrow];
  const float gl = __fmul_rn(gamma, lam);
  float g = 0.f;
  float nv = values[(size_t)T * C + c];
  int t = T;
  while (t > 0) {
    float rr[kUnroll], vv[kUnroll];
    uint8_t dd[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int ti = t - 1 - i;
      if (ti >= 0) {
        const size_t o = (size_t)ti * C + c;
        rr[i] = ld_stream(rewards + o);
        vv[i] = ld_stream(values + o);
        dd[i] = __ldg(dones + o);
      }
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int ti = t - 1 - i;
      if (ti >= 0) {
        g = gae_step(g, gl, gamma, nv, vv[i], rr[i], not_done(dd[i]));
        nv = vv[i];
        const size_t o = (size_t)ti * C + c;
        ret[o] = __fadd_rn(vv[i], g);
        if (vrow == 0) adv[(size_t)ti * N + c] = g;
      }
    }
    t -= kUnroll;
  }
}

// blockDim = (COLS, CH); CH power of two <= 32; COLS*CH multiple of 32.
__global__ void __launch_bounds__(1024) gae_chunked_kernel(const float* __restrict__ values, const float* __restrict__ rewards,
                                                           const uint8_t* __restrict__ dones, GaeGammas gam, float lam,
                                                           int T, int N, int C, float* __restrict__ ret,
                                                           float* __restrict__ adv) {
  extern __shared__ float sm[];
  const int COLS = blockDim.x, CH = blockDim.y;
  float* sa = sm;                 // [CH][COLS]
  float* sb = sm + CH * COLS;     // [CH][COLS]  (reused for the carry-in)
  const int x = threadIdx.x, ch = threadIdx.y;
  const int c = blockIdx.x * COLS + x;
  const bool live = c < C;
  const int Lc = (T + CH - 1) / CH;
  const int t0 = min(T, ch * Lc), t1 = min(T, t0 + Lc);
  float gamma = 0.f, gl = 0.f;
  int vrow = 0;
  if (live) {
    vrow = c / N;
    gamma = gam.g[vrow];
    gl = __fmul_rn(gamma, lam);
  }
  // ---- pass 1: fold the chunk [t0, t1) into g(t0) = a * g(t1) + b
  float a = 1.f, b = 0.f;
  if (live && t1 > t0) {
    float nv = values[(size_t)t1 * C + c];
    for (int t = t1 - 1; t >= t0; --t) {
      const size_t o = (size_t)t * C + c;
      const float v = values[o], r = rewards[o], nd = not_done(dones[o]);
      const float A = gl * nd;
      const float Bt = (gamma * nv) * nd - v + r;
      b = fmaf(A, b, Bt);
      a = A * a;
      nv = v;
    }
  }
  sa[ch * COLS + x] = a;
  sb[ch * COLS + x] = b;
  __syncthreads();
  // ---- warp-level associative suffix scan over the CH chunk maps of each column
  {
    const int tid = ch * COLS + x;
    const int col = tid / CH, k = tid % CH;      // CH consecutive lanes = the chunks of one column
    float pa = sa[k * COLS + col], pb = sb[k * COLS + col];
    // S_k = F_k o F_{k+1} o ... o F_{CH-1}  (later chunks are applied first)
    for (int off = 1; off < CH; off <<= 1) {
      const float qa = __shfl_down_sync(0xffffffffu, pa, off, CH);
      const float qb = __shfl_down_sync(0xffffffffu, pb, off, CH);
      if (k + off < CH) {
        pb = fmaf(pa, qb, pb);
        pa = pa * qa;
      }
    }
    // carry-in of chunk k = S_{k+1}(0) = b part of S_{k+1}; the last chunk starts from g = 0
    float carry = __shfl_down_sync(0xffffffffu, pb, 1, CH);
    if (k == CH - 1) carry = 0.f;
    __syncthreads();
    sb[k * COLS + col] = carry;
  }
  __syncthreads();
  // ---- pass 2: replay with the reference's exact per-step rounding
  if (live && t1 > t0) {
    float g = sb[ch * COLS + x];
    float nv = values[(size_t)t1 * C + c];
    for (int t = t1 - 1; t >= t0; --t) {
      const size_t o = (size_t)t * C + c;
      const float v = values[o], r = rewards[o], nd = not_done(dones[o]);
      g = gae_step(g, gl, gamma, nv, v, r, nd);
      nv = v;
      ret[o] = __fadd_rn(v, g);
      if (vrow == 0) adv[(size_t)t * N + c] = g;
    }
  }
}

}  // namespace ddrl

extern "C" int ddrl_gae_f32(const float* values, const float* rewards, const uint8_t* dones,
                            const float* gamma_host, float lambda, int T, int V, int N, float* ret, float* adv,
                            int algo, void* stream) {
  using namespace ddrl;
  if (T < 0 || V < 1 || V > kMaxV || N < 0 || !gamma_host) return DDRL_E_ARG;
  if (T == 0 || N == 0) return DDRL_OK;      // empty rollout: agent.py:125-126 returns []
  if (!values || !rewards || !dones || !ret || !adv) return DDRL_E_ARG;
  GaeGammas gam;
  for (int i = 0; i < kMaxV; ++i) gam.g[i] = i < V ? gamma_host[i] : 0.f;
  const int C = V * N;
  cudaStream_t s = (cudaStream_t)stream;
  if (algo == 0) algo = (C >= 32768 || T < 8) ? 1 : 2;
  if (algo == 1) {
    const int threads = 128;
    gae_seq_kernel<8><<<ceil_div(C, threads), threads, 0, s>>>(values, rewards, dones, gam, lambda, T, N, C, ret, adv);
    prof_work(17.0 * T * (double)C);
    DDRL_LAUNCHED("gae_seq_kernel");
  } else if (algo == 2) {
    int CH = 32;
    while (CH > 1 && CH * 4 > T) CH >>= 1;       // at least ~4 steps per chunk
    int COLS = 32;
    while (COLS > 8 && ceil_div(C, COLS) < 2 * kNumSMs) COLS >>= 1;
    while (COLS * CH < 32) COLS <<= 1;
    dim3 block(COLS, CH);
    const size_t smem = 2 * sizeof(float) * COLS * CH;
    gae_chunked_kernel<<<ceil_div(C, COLS), block, smem, s>>>(values, rewards, dones, gam, lambda, T, N, C, ret, adv);
    prof_work(17.0 * T * (double)C);
    DDRL_LAUNCHED("gae_chunked_kernel");
  } else {
    return DDRL_E_ARG;
  }
  return DDRL_OK;
}
