// Bodies of the small per-layer weight-preparation kernels (re-packing for the engines, amax, fp16 / tf32 splits, gradient
// un-permutation).  Each is written against a VIRTUAL grid (vb = block index, nvb = blocks of this job): the stand-alone
// __global__ wrappers pass (blockIdx.x, gridDim.x); prep.cu runs MANY of them as one launch from a device-resident job
// table -- after every optimiser step the learner re-prepares ~45 weight operands, which as separate launches cost ~0.7 ms
// of a 7 ms iteration although they move < 100 MB.
#pragma once
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ddrl {

// ---- tc3 operand scaling (shared with tc3.cu) ----------------------------------------------------------------------
// lo' = lo * 2^11 keeps the residual in the fp16 normal range down to |x s| = 2^-14, i.e. 2^-27 of the tensor's amax.
constexpr float T3_LO = 2048.f, T3_LO_INV = 1.f / 2048.f;

// power-of-two scale that maps amax into [2^13, 2^14) and its inverse, from the exponent bits (amax = 0 or denormal: the
// clamp keeps both finite; inf / nan inputs poison the result either way)
__host__ __device__ __forceinline__ void t3_scale(float amax, float& s, float& inv) {
#ifdef __CUDA_ARCH__
  int e = (__float_as_int(amax) >> 23) & 0xff;
#else
  uint32_t bits; memcpy(&bits, &amax, 4);
  int e = (int)((bits >> 23) & 0xff);
#endif
  if (amax == 0.f) e = 127 + 13;
  e = e < 14 ? 14 : (e > 253 ? 253 : e);
  const uint32_t sb = (uint32_t)(267 - e) << 23, ib = (uint32_t)(e - 13) << 23;
#ifdef __CUDA_ARCH__
  s = __uint_as_float(sb); inv = __uint_as_float(ib);
#else
  memcpy(&s, &sb, 4); memcpy(&inv, &ib, 4);
#endif
}

#define PREP_FOR(t, total) for (long long t = vb * 256LL + threadIdx.x; t < (total); t += (long long)nvb * 256LL)

// packed[o*ld + i*J + j] = src[o*I*J + j*I + i]   (inner [J][I] -> [I][J] transpose; I = 1: row-stride change)
__device__ __forceinline__ void pack_body(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int J, int ld,
                                          unsigned vb, unsigned nvb) {
  const long long total = (long long)O * I * J;
  PREP_FOR(t, total) {
    const int j = (int)(t % J);
    const long long r = t / J;
    const int i = (int)(r % I);
    const long long o = r / I;
    dst[o * ld + (long long)i * J + j] = src[o * I * J + (long long)j * I + i];
  }
}
// grad[o*I*J + j*I + i] = packed_grad[o*ld + i*J + j]
__device__ __forceinline__ void unpack_body(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int J, int ld,
                                            unsigned vb, unsigned nvb) {
  const long long total = (long long)O * I * J;
  PREP_FOR(t, total) {
    const int i = (int)(t % I);
    const long long r = t / I;
    const int j = (int)(r % J);
    const long long o = r / J;
    dst[t] = src[o * ld + (long long)i * J + j];
  }
}
// packed[o*ld + ((a*KW2 + b)*s*s + i*s + j)*C + c] = w[o, c, s*a + i, s*b + j]      (w: reference OIHW)
__device__ __forceinline__ void pack_s2d_body(const float* __restrict__ w, float* __restrict__ dst, int O, int C, int KH, int KW,
                                              int s, int ld, int unpack, unsigned vb, unsigned nvb) {
  const long long total = (long long)O * C * KH * KW;
  const int KW2 = KW / s;
  PREP_FOR(t, total) {
    const int kw = (int)(t % KW);
    long long r = t / KW;
    const int kh = (int)(r % KH); r /= KH;
    const int c = (int)(r % C);
    const long long o = r / C;
    const int a = kh / s, i = kh - a * s, b = kw / s, j = kw - b * s;
    const long long pk = o * ld + ((long long)(a * KW2 + b) * s * s + i * s + j) * C + c;
    if (unpack) dst[t] = w[pk]; else dst[pk] = w[t];
  }
}
// Wd[c, (th*ntx + tw)*Cout + o] = w[o, c, kh, kw]  with kh = ry + s*(nty-1-th), kw = rx + s*(ntx-1-tw)   (w: reference OIHW)
__device__ __forceinline__ void pack_dgrad_body(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin, int KH,
                                                int KW, int s, int ry, int rx, int nty, int ntx, unsigned vb, unsigned nvb) {
  const long long total = (long long)Cin * nty * ntx * Cout;
  PREP_FOR(t, total) {
    const int o = (int)(t % Cout);
    long long r = t / Cout;
    const int tw = (int)(r % ntx); r /= ntx;
    const int th = (int)(r % nty);
    const int c = (int)(r / nty);
    const int kh = ry + s * (nty - 1 - th), kw = rx + s * (ntx - 1 - tw);
    wd[t] = w[(((long long)o * Cin + c) * KH + kh) * KW + kw];
  }
}
// Wd[(cy*ncx + cx)*Cin + c, (th*ntx + tw)*Cout + o] = w[o, c, kh, kw],  kh = cy + s*(q0y[cy] + pady - th) (0 if no such tap)
__device__ __forceinline__ void pack_dgrad_fused_body(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin,
                                                      int KH, int KW, int s, int sx, int ncx, int nty, int ntx, int pady, int padx,
                                                      int q0y0, int q0y1, int q0x0, int q0x1, long long total, unsigned vb,
                                                      unsigned nvb) {
  PREP_FOR(t, total) {
    const int o = (int)(t % Cout);
    long long r = t / Cout;
    const int tw = (int)(r % ntx); r /= ntx;
    const int th = (int)(r % nty); r /= nty;
    const int c = (int)(r % Cin); r /= Cin;
    const int cx = (int)(r % ncx), cy = (int)(r / ncx);
    const int uy = (cy ? q0y1 : q0y0) + pady - th, ux = (cx ? q0x1 : q0x0) + padx - tw;
    const int kh = cy + s * uy, kw = cx + sx * ux;
    float v = 0.f;
    if (uy >= 0 && ux >= 0 && kh < KH && kw < KW) v = w[(((long long)o * Cin + c) * KH + kh) * KW + kw];
    wd[t] = v;
  }
}
// hi = rn_tf32(w), lo = rn_tf32(w - hi): the weight operand's split of the tf32 engines
__device__ __forceinline__ void split_hi_lo_body(const float4* __restrict__ w, uint4* __restrict__ hi, uint4* __restrict__ lo,
                                                 long long n4, unsigned vb, unsigned nvb) {
  PREP_FOR(i, n4) {
    const float4 x = w[i];
    uint4 h, l;
    h.x = tf32_rn(x.x); h.y = tf32_rn(x.y); h.z = tf32_rn(x.z); h.w = tf32_rn(x.w);
    l.x = tf32_rn(x.x - __uint_as_float(h.x)); l.y = tf32_rn(x.y - __uint_as_float(h.y));
    l.z = tf32_rn(x.z - __uint_as_float(h.z)); l.w = tf32_rn(x.w - __uint_as_float(h.w));
    hi[i] = h; lo[i] = l;
  }
}
// amax|x| over a [rows, cols] view (row stride ld) -> atomicMax on the bits of *slot (zeroed beforehand)
__device__ __forceinline__ void amax_body(const float* __restrict__ x, long long rows, int cols, long long ld,
                                          unsigned int* __restrict__ slot, unsigned vb, unsigned nvb) {
  float m = 0.f;
  if (cols == ld && (cols & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long n4 = rows * cols / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    PREP_FOR(i, n4) {
      const float4 v = x4[i];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const long long n = rows * cols;
    PREP_FOR(i, n) {
      const long long r = i / cols;
      m = fmaxf(m, fabsf(x[r * ld + (i - r * cols)]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    if (m > 0.f) atomicMax(slot, __float_as_uint(m));
  }
}
// weights w [N, ldw] fp32 (K valid columns) -> hi / lo' [N, ld16] fp16 with the scale of *amax; optionally the transposed
// pair hiT / loT [K, ldT16] (the K-major weight operand of a linear layer's data gradient).  Padding columns are zero.
__device__ __forceinline__ void split_f16_body(const float* __restrict__ w, int N, int K, int ldw, const float* __restrict__ amax,
                                               __half* __restrict__ hi, __half* __restrict__ lo, int ld16, __half* __restrict__ hiT,
                                               __half* __restrict__ loT, int ldT16, unsigned vb, unsigned nvb) {
  float s, inv;
  t3_scale(*amax, s, inv);
  const long long n = (long long)N * ld16;
  PREP_FOR(i, n) {
    const int r = (int)(i / ld16), k = (int)(i - (long long)r * ld16);
    __half h = __float2half_rn(0.f), l = h;
    if (k < K) {
      const float y = w[(long long)r * ldw + k] * s;
      h = __float2half_rn(y);
      l = __float2half_rn((y - __half2float(h)) * T3_LO);
      if (hiT) { hiT[(long long)k * ldT16 + r] = h; loT[(long long)k * ldT16 + r] = l; }
    }
    hi[i] = h; lo[i] = l;
  }
}

// ---- job table ----------------------------------------------------------------------------------------------------------
enum PrepType {
  PREP_PACK = 0, PREP_UNPACK, PREP_PACK_S2D, PREP_PACK_DGRAD, PREP_PACK_DGRAD_FUSED, PREP_SPLIT_HILO, PREP_AMAX, PREP_SPLIT_F16,
  PREP_COPY, PREP_ZERO
};
struct PrepJob {
  int type, vblocks;
  const void* a;          // source
  void *b, *c, *d, *e;    // destinations
  const float* amax;
  long long total;
  int i[16];
};
// host side (prep.cu): while a recorder is installed the weight-preparation wrappers append jobs instead of launching
struct PrepRecorder { std::vector<PrepJob> jobs; };
extern thread_local PrepRecorder* g_prep_rec;
inline bool prep_record(const PrepJob& j) {
  if (!g_prep_rec) return false;
  g_prep_rec->jobs.push_back(j);
  return true;
}
inline int prep_blocks(long long total, int per_block = 256) {
  const long long b = (total + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : (b > 4LL * kNumSMs ? 4LL * kNumSMs : b));
}

}  // namespace ddrl
