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
// Schedules (algo argument of ddrl_gae_f32):
//  * gae_seq_kernel (1) / gae_seq_vec_kernel (4): one thread per column (per 2 adjacent columns) walks time backwards
//    with the loads of the next steps in flight (they do not depend on the recurrence).  Every step uses the reference's
//    exact rounding sequence (__fmul_rn/__fadd_rn, no FMA contraction), so the result is BIT-EXACT with the numpy loop.
//  * gae_tiled_kernel (3/5/7, the default): single pass, time-parallel.  Each thread keeps LC steps of VEC columns in
//    registers, folds them into an affine map, the maps of a column are combined by a warp-level associative scan
//    (Hillis-Steele over shuffles), and the steps are replayed from registers with the exact rounding sequence.  Only
//    the carry-in of each chunk is reassociated (damped by gamma*lam per step; measured <= 3e-7 of max|adv|).
//    Measured on B200 (profiles/r1_gae_sweep.json): what decides the achieved HBM fraction is the number of
//    CONTIGUOUS bytes a warp requests per time row (rows are N*4 bytes apart, every request opens another DRAM page):
//    128 B/warp-row (VEC=1) tops out at 0.59-0.68 of the copy peak, 512 B/warp-row (VEC=4) reaches 0.88.
//  * gae_chunked_kernel (2): two-pass predecessor of the tiled schedule (fold, scan, replay from L2); kept for A/B.
#include <stdlib.h>

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


// Vectorised bit-exact schedule: one thread walks VEC adjacent columns (16-byte loads for VEC = 4), so a warp touches
// 32*VEC*4 contiguous bytes of every time row (DRAM-page friendly) and carries VEC independent recurrences (ILP).
// Double-buffered: the loads of batch k+1 are issued BEFORE batch k is consumed (two register buffers of U steps), so a
// thread always has U..2U steps of independent loads in flight while it walks the dependent chain.  Requires N % VEC == 0 (the VEC columns share one value row / gamma).
template <int VEC> struct GaeVec;
template <> struct GaeVec<2> { using F = float2; using B = uint16_t; };

template <int VEC, int U>
__global__ void __launch_bounds__(64) gae_seq_vec_kernel(const float* __restrict__ values, const float* __restrict__ rewards,
                                                         const uint8_t* __restrict__ dones, GaeGammas gam, float lam,
                                                         int T, int N, int C, float* __restrict__ ret,
                                                         float* __restrict__ adv) {
  using F = typename GaeVec<VEC>::F;
  using B = typename GaeVec<VEC>::B;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (c >= C) return;
  const int vrow = c / N;
  const float gamma = gam.g[vrow];
  const float gl = __fmul_rn(gamma, lam);
  float g[VEC], nv[VEC];
  {
    const F b = *reinterpret_cast<const F*>(values + (size_t)T * C + c);
    const float* bp = reinterpret_cast<const float*>(&b);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { g[k] = 0.f; nv[k] = bp[k]; }
  }
  F rA[U], vA[U], rB[U], vB[U];
  B dA[U], dB[U];
#define GAE_LOADV(R, V, D, tt)                                                 \
  _Pragma("unroll") for (int i = 0; i < U; ++i) {                              \
    const int ti = (tt)-1 - i;                                                 \
    if (ti >= 0) {                                                             \
      const size_t o = (size_t)ti * C + c;                                     \
      R[i] = __ldcs(reinterpret_cast<const F*>(rewards + o));                  \
      V[i] = __ldcs(reinterpret_cast<const F*>(values + o));                   \
      D[i] = __ldcs(reinterpret_cast<const B*>(dones + o));                    \
    }                                                                          \
  }
#define GAE_CONSUMEV(R, V, D, tt)                                              \
  _Pragma("unroll") for (int i = 0; i < U; ++i) {                              \
    const int ti = (tt)-1 - i;                                                 \
    if (ti >= 0) {                                                             \
      const float* rp = reinterpret_cast<const float*>(&R[i]);                 \
      const float* vp = reinterpret_cast<const float*>(&V[i]);                 \
      const uint32_t dw = D[i];                                                \
      F orr, oa;                                                               \
      float* orp = reinterpret_cast<float*>(&orr);                             \
      float* oap = reinterpret_cast<float*>(&oa);                              \
      _Pragma("unroll") for (int k = 0; k < VEC; ++k) {                        \
        g[k] = gae_step(g[k], gl, gamma, nv[k], vp[k], rp[k], not_done((uint8_t)(dw >> (8 * k)))); \
        nv[k] = vp[k];                                                         \
        orp[k] = __fadd_rn(vp[k], g[k]);                                       \
        oap[k] = g[k];                                                         \
      }                                                                        \
      const size_t o = (size_t)ti * C + c;                                     \
      __stcs(reinterpret_cast<F*>(ret + o), orr);                              \
      if (vrow == 0) __stcs(reinterpret_cast<F*>(adv + (size_t)ti * N + c), oa); \
    }                                                                          \
  }
  GAE_LOADV(rA, vA, dA, T)
  for (int t = T; t > 0; t -= 2 * U) {
    GAE_LOADV(rB, vB, dB, t - U)
    GAE_CONSUMEV(rA, vA, dA, t)
    GAE_LOADV(rA, vA, dA, t - 2 * U)
    GAE_CONSUMEV(rB, vB, dB, t - U)
  }
#undef GAE_LOADV
#undef GAE_CONSUMEV
}

// Time-parallel single-pass schedule: blockDim = (COLS, CH), COLS*CH = 256.  The block walks the time axis backwards
// in super-chunks of CH*LC steps.  Thread (x, ch) owns LC consecutive steps of VEC adjacent columns, kept in REGISTERS:
//   load (all LC steps independent, in flight together; 16-byte loads for VEC = 4, so a warp touches 512 contiguous
//   bytes of every time row) -> fold into one affine map g_out = a g_in + b per column ->
//   warp-level associative scan (Hillis-Steele over shuffles) of the CH maps of each column, seeded with the carry of
//   the previous super-chunk -> replay the LC steps from registers with the reference's exact rounding sequence.
// HBM sees every input byte once (17 B per (t, column)) for any T; only the carry-in of each chunk is reassociated.
template <int VEC> struct GaeTile;
template <> struct GaeTile<1> { using F = float; using B = uint8_t; };
template <> struct GaeTile<4> { using F = float4; using B = uint32_t; };

template <int VEC, int LC>
__global__ void __launch_bounds__(256, 2) gae_tiled_kernel(const float* __restrict__ values, const float* __restrict__ rewards,
                                                           const uint8_t* __restrict__ dones, GaeGammas gam, float lam,
                                                           int T, int N, int C, float* __restrict__ ret,
                                                           float* __restrict__ adv) {
  using F = typename GaeTile<VEC>::F;
  using B = typename GaeTile<VEC>::B;
  extern __shared__ float sm[];
  const int COLS = blockDim.x, CH = blockDim.y, W = COLS * VEC, LD = W + 4;
  float* sa = sm;                  // [CH][LD]
  float* sb = sm + CH * LD;        // [CH][LD]  maps, then carry-ins
  float* sc = sm + 2 * CH * LD;    // [W]       g carried across super-chunks
  const int x = threadIdx.x, ch = threadIdx.y;
  const int c = (blockIdx.x * COLS + x) * VEC;
  const bool live = c < C;
  int vrow = 0;
  float gamma = 0.f, gl = 0.f;
  if (live) {
    vrow = c / N;
    gamma = gam.g[vrow];
    gl = __fmul_rn(gamma, lam);
  }
  if (ch == 0) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) sc[x * VEC + k] = 0.f;
  }
  const int tid = ch * COLS + x;
  const int scol = tid / CH, sk = tid % CH;       // scan role: CH consecutive lanes = the chunks of one column
  const int S = CH * LC;
  for (int tend = T; tend > 0; tend -= S) {
    const int t1 = tend - ch * LC;                // this thread's steps: t1-1, t1-2, ..., t1-LC (those >= 0)
    const bool work = live && t1 > 0;
    F v[LC], r[LC];
    B dn[LC];
    float nv0[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) nv0[k] = 0.f;
    if (work) {
      const F b0 = __ldg(reinterpret_cast<const F*>(values + (size_t)t1 * C + c));
#pragma unroll
      for (int k = 0; k < VEC; ++k) nv0[k] = reinterpret_cast<const float*>(&b0)[k];
#pragma unroll
      for (int i = 0; i < LC; ++i) {
        const int t = t1 - 1 - i;
        if (t >= 0) {
          const size_t o = (size_t)t * C + c;
          v[i] = __ldcs(reinterpret_cast<const F*>(values + o));
          r[i] = __ldcs(reinterpret_cast<const F*>(rewards + o));
          dn[i] = __ldcs(reinterpret_cast<const B*>(dones + o));
        }
      }
    }
    float a[VEC], b[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { a[k] = 1.f; b[k] = 0.f; }
    if (work) {
      float nv[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) nv[k] = nv0[k];
#pragma unroll
      for (int i = 0; i < LC; ++i) {
        if (t1 - 1 - i >= 0) {
          const uint32_t dw = dn[i];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float nd = not_done((uint8_t)(dw >> (8 * k)));
            const float vk = reinterpret_cast<const float*>(&v[i])[k], rk = reinterpret_cast<const float*>(&r[i])[k];
            const float A = gl * nd;
            const float Bt = (gamma * nv[k]) * nd - vk + rk;
            b[k] = fmaf(A, b[k], Bt);
            a[k] = A * a[k];
            nv[k] = vk;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      sa[ch * LD + x * VEC + k] = a[k];
      sb[ch * LD + x * VEC + k] = b[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      // P_j = F_j o ... o F_0 (chunk 0 is the latest in time and is applied first)
      const int col = scol * VEC + k;
      float pa = sa[sk * LD + col], pb = sb[sk * LD + col];
      for (int off = 1; off < CH; off <<= 1) {
        const float qa = __shfl_up_sync(0xffffffffu, pa, off, CH);
        const float qb = __shfl_up_sync(0xffffffffu, pb, off, CH);
        if (sk >= off) {
          pb = fmaf(pa, qb, pb);
          pa = pa * qa;
        }
      }
      const float g0 = sc[col];
      const float ua = __shfl_up_sync(0xffffffffu, pa, 1, CH);
      const float ub = __shfl_up_sync(0xffffffffu, pb, 1, CH);
      sb[sk * LD + col] = sk == 0 ? g0 : fmaf(ua, g0, ub);      // carry-in of chunk sk
    }
    __syncthreads();
    if (work) {
      float g[VEC], nv[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) { g[k] = sb[ch * LD + x * VEC + k]; nv[k] = nv0[k]; }
#pragma unroll
      for (int i = 0; i < LC; ++i) {
        const int t = t1 - 1 - i;
        if (t >= 0) {
          const uint32_t dw = dn[i];
          F orr, oa;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float vk = reinterpret_cast<const float*>(&v[i])[k], rk = reinterpret_cast<const float*>(&r[i])[k];
            g[k] = gae_step(g[k], gl, gamma, nv[k], vk, rk, not_done((uint8_t)(dw >> (8 * k))));
            nv[k] = vk;
            reinterpret_cast<float*>(&orr)[k] = __fadd_rn(vk, g[k]);
            reinterpret_cast<float*>(&oa)[k] = g[k];
          }
          const size_t o = (size_t)t * C + c;
          __stcs(reinterpret_cast<F*>(ret + o), orr);
          if (vrow == 0) __stcs(reinterpret_cast<F*>(adv + (size_t)t * N + c), oa);
        }
      }
      if (ch == CH - 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) sc[x * VEC + k] = g[k];
      }
    }
  }
}

template <int VEC, int LC>
static int launch_tiled(const float* values, const float* rewards, const uint8_t* dones, const GaeGammas& gam, float lambda, int T,
                        int N, int C, float* ret, float* adv, cudaStream_t s) {
  const int cols = C / VEC;
  static int min_blocks = -1;
  if (min_blocks < 0) { const char* e = getenv("DDRL_GAE_MINBLOCKS"); min_blocks = e ? atoi(e) : kNumSMs; }
  int COLS = 32;
  while (COLS > 8 && ceil_div(cols, COLS) < min_blocks) COLS >>= 1;
  int CH = 256 / COLS;
  while (CH > 1 && (CH / 2) * LC >= T) CH >>= 1;     // no more chunks than the time axis has
  while (COLS * CH < 32) CH <<= 1;
  dim3 block(COLS, CH);
  const size_t smem = sizeof(float) * (2 * CH * (COLS * VEC + 4) + COLS * VEC);
  gae_tiled_kernel<VEC, LC><<<ceil_div(cols, COLS), block, smem, s>>>(values, rewards, dones, gam, lambda, T, N, C, ret, adv);
  prof_work(17.0 * T * (double)C);
  DDRL_LAUNCHED("gae_tiled_kernel");
  return DDRL_OK;
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


// ---- tempo-GAE: Agents._accumulate_tempo_rewards (USTC_lab/agent/agent.py:142-160) ------------------------------------
// Per-step discount td[t] = tempo_discounts[durations[t]] from the float64 table np.logspace(0, 100, 101, base=gamma)
// (agent.py:119).  td is an np.float64 scalar, so under NumPy-2 promotion the whole recurrence runs in float64 in the
// reference; this kernel does the same (explicit __dmul_rn/__dadd_rn, no FMA contraction) => BIT-EXACT f64 results.
// One thread per column walks time backwards with kUnroll steps of loads in flight.  OUT = double (the reference's
// arrays) or float (one final rounding, what Experience.to_tensor does downstream; halves the write traffic).
struct TempoTable { double d[101]; };

template <typename OUT, int kUnroll>
__global__ void __launch_bounds__(128) gae_tempo_kernel(const float* __restrict__ values, const float* __restrict__ rewards,
                                                        const uint8_t* __restrict__ dones, const int* __restrict__ durations,
                                                        const __grid_constant__ TempoTable table, double lam, int T, int N,
                                                        int C, OUT* __restrict__ ret, OUT* __restrict__ adv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const bool row0 = c < N;
  double g = 0.0;
  double nv = (double)values[(size_t)T * C + c];
  int t = T;
  while (t > 0) {
    float rr[kUnroll], vv[kUnroll];
    uint8_t dd[kUnroll];
    int du[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int ti = t - 1 - i;
      if (ti >= 0) {
        const size_t o = (size_t)ti * C + c;
        rr[i] = ld_stream(rewards + o);
        vv[i] = ld_stream(values + o);
        dd[i] = __ldg(dones + o);
        du[i] = __ldg(durations + ti);
      }
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int ti = t - 1 - i;
      if (ti >= 0) {
        const double td = table.d[du[i]];
        const double nd = (double)(uint8_t)(1 - dd[i]);
        const double v = (double)vv[i];
        g = __dmul_rn(g, nd);
        const double x = __dmul_rn(__dmul_rn(td, lam), g);
        double y = __dmul_rn(__dmul_rn(td, nv), nd);
        y = __dsub_rn(y, v);
        y = __dadd_rn(y, (double)rr[i]);
        g = __dadd_rn(x, y);
        nv = v;
        const size_t o = (size_t)ti * C + c;
        ret[o] = (OUT)__dadd_rn(v, g);
        if (row0) adv[(size_t)ti * N + c] = (OUT)g;
      }
    }
    t -= kUnroll;
  }
}

}  // namespace ddrl

extern "C" int ddrl_gae_tempo(const float* values, const float* rewards, const uint8_t* dones, const int* durations,
                              const double* table_host, int table_len, double lambda, int T, int V, int N, void* ret,
                              void* adv, int out_f64, void* stream) {
  using namespace ddrl;
  if (T < 0 || V < 1 || N < 0 || !table_host || table_len < 1 || table_len > 101) return DDRL_E_ARG;
  if (T == 0 || N == 0) return DDRL_OK;      // empty rollout: agent.py:143-144 returns []
  if (!values || !rewards || !dones || !durations || !ret || !adv) return DDRL_E_ARG;
  TempoTable tab;
  for (int i = 0; i < 101; ++i) tab.d[i] = table_host[i < table_len ? i : table_len - 1];
  const int C = V * N, threads = 128;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_f64)
    gae_tempo_kernel<double, 8><<<ceil_div(C, threads), threads, 0, s>>>(values, rewards, dones, durations, tab, lambda, T, N, C,
                                                                         (double*)ret, (double*)adv);
  else
    gae_tempo_kernel<float, 8><<<ceil_div(C, threads), threads, 0, s>>>(values, rewards, dones, durations, tab, lambda, T, N, C,
                                                                        (float*)ret, (float*)adv);
  prof_work((out_f64 ? 25.0 : 17.0) * T * (double)C);
  DDRL_LAUNCHED("gae_tempo_kernel");
  return DDRL_OK;
}

namespace ddrl {
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
  // algo: 0 auto | 1 sequential scalar (bit-exact) | 2 two-pass chunked | 3 time-parallel single pass, widest vector that fits |
  //       4 sequential vectorised (bit-exact) | 5 time-parallel VEC=1 | 7 time-parallel VEC=4
  const bool al16 = (reinterpret_cast<uintptr_t>(values) | reinterpret_cast<uintptr_t>(rewards) | reinterpret_cast<uintptr_t>(ret) |
                     reinterpret_cast<uintptr_t>(adv)) % 16 == 0 && reinterpret_cast<uintptr_t>(dones) % 4 == 0;
  if (algo == 0) algo = T < 8 ? 1 : 3;
  if (algo == 3) algo = (N % 4 == 0 && al16 && C >= 8192) ? 7 : 5;
  if (algo == 4 && (N % 2 != 0 || !al16)) algo = 1;
  if (algo == 1) {
    const int threads = 128;
    gae_seq_kernel<8><<<ceil_div(C, threads), threads, 0, s>>>(values, rewards, dones, gam, lambda, T, N, C, ret, adv);
    prof_work(17.0 * T * (double)C);
    DDRL_LAUNCHED("gae_seq_kernel");
  } else if (algo == 4) {
    const int threads = 64, cols = C / 2;
    gae_seq_vec_kernel<2, 8><<<ceil_div(cols, threads), threads, 0, s>>>(values, rewards, dones, gam, lambda, T, N, C, ret, adv);
    prof_work(17.0 * T * (double)C);
    DDRL_LAUNCHED("gae_seq_vec_kernel");
  } else if (algo == 5) {
    return launch_tiled<1, 16>(values, rewards, dones, gam, lambda, T, N, C, ret, adv, s);
  } else if (algo == 7) {
    if (N % 4 != 0 || !al16) return DDRL_E_UNSUPPORTED;
    return launch_tiled<4, 8>(values, rewards, dones, gam, lambda, T, N, C, ret, adv, s);
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
