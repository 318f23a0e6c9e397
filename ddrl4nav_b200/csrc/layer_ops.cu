// Layer glue kernels around the GEMM engines: im2col / col2im (conv as GEMM, NHWC internal
// activations), fused ReLU+2x2 max-pool forward/backward, activation backward, bias-gradient
// column sums, weight (un)packing between the reference's OIHW / C-major-flatten layouts and the
// engine's K-contiguous packed layouts, and the skinny (N <= 32) head linears.
//
// Reference semantics reproduced:
//   conv = cross-correlation, OIHW weights, NCHW tensors      nn/atari_encoder.py:16-18,26-28
//   max_pool2d(relu(conv(x)), 2, stride=2), floor mode         nn/nav_encoder.py:29-31,100-102
//     tie -> gradient to the FIRST max in row-major window order; relu'(0) = 0
//   leaky_relu slope 0.01, gradient at exactly 0 is 0.01       nn/atari_encoder.py:26-28
//   flatten before the first Linear is C-major (c*H*W + h*W + w) nn/atari_encoder.py:30
// All of these are HBM-bound streaming kernels: coalesced along the channel / K axis.
#include <algorithm>

#include "common.cuh"
#include "layer_ops.h"
#include "prep_kernels.cuh"

namespace ddrl {

// ---------------------------------------------------------------- im2col
// cols[(b,ho,wo), k] with k = (kh*KW + kw)*C + c   (order 0, NHWC-friendly)
//                     or k = (c*KH + kh)*KW + kw   (order 1, = OIHW weight order, for NCHW inputs)
// columns [K, ldc) are zero padding (written every time; ldc - K < 4).
__global__ void __launch_bounds__(256) im2col_kernel(ConvGeom g, const float* __restrict__ x, float* __restrict__ cols,
                                                     long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % g.ldc);
    const long long row = i / g.ldc;
    float v = 0.f;
    if (k < g.K) {
      int c, kh, kw;
      if (g.order == 0) { c = k % g.C; const int t = k / g.C; kw = t % g.KW; kh = t / g.KW; }
      else { kw = k % g.KW; const int t = k / g.KW; kh = t % g.KH; c = t / g.KH; }
      const int wo = (int)(row % g.Wo);
      const long long t2 = row / g.Wo;
      const int ho = (int)(t2 % g.Ho);
      const long long b = t2 / g.Ho;
      const int h = ho * g.stride - g.pad + kh, w = wo * g.stride - g.pad + kw;
      if (h >= 0 && h < g.H && w >= 0 && w < g.W) v = x[b * g.sb + h * g.sh + w * g.sw + c * g.sc];
    }
    cols[i] = v;
  }
}

// Row-cooperative variant for the small-K first layers (K <= 1024, any order): a group of KT = 2^j >= min(ldc, 32) lanes owns
// one output row -- the (image, ho, wo) decode happens once per row in 32-bit arithmetic, the per-k tap offsets come from a
// table in shared memory, and the lanes' stores are consecutive floats of the row.
__global__ void __launch_bounds__(256) im2col_rows_kernel(ConvGeom g, const float* __restrict__ x, float* __restrict__ cols,
                                                          long long rows, int kt) {
  extern __shared__ int tab[];                          // [ldc] input offset | [ldc] kh | [ldc] kw   (k >= K: kh = -2^20)
  int* koff = tab; int* kkh = tab + g.ldc; int* kkw = tab + 2 * g.ldc;
  for (int k = threadIdx.x; k < g.ldc; k += blockDim.x) {
    int c = 0, kh = -(1 << 20), kw = 0;
    if (k < g.K) {
      if (g.order == 0) { c = k % g.C; const int t = k / g.C; kw = t % g.KW; kh = t / g.KW; }
      else { kw = k % g.KW; const int t = k / g.KW; kh = t % g.KH; c = t / g.KH; }
    }
    koff[k] = (int)(c * g.sc + (long long)kh * g.sh + (long long)kw * g.sw);
    kkh[k] = kh; kkw[k] = kw;
  }
  __syncthreads();
  const int rpb = blockDim.x / kt;                      // rows per block iteration
  const int kl = threadIdx.x % kt, rl = threadIdx.x / kt;
  const int hw = g.Ho * g.Wo;
  for (long long row = (long long)blockIdx.x * rpb + rl; row < rows; row += (long long)gridDim.x * rpb) {
    const long long b = row / hw;
    const int p = (int)(row - b * hw);
    const int ho = p / g.Wo, wo = p - ho * g.Wo;
    const int h0 = ho * g.stride - g.pad, w0 = wo * g.stride - g.pad;
    const float* xb = x + b * g.sb + (long long)h0 * g.sh + (long long)w0 * g.sw;
    float* cr = cols + row * g.ldc;
    for (int k = kl; k < g.ldc; k += kt) {
      const int h = h0 + kkh[k], w = w0 + kkw[k];
      cr[k] = (h >= 0 && h < g.H && w >= 0 && w < g.W) ? __ldg(xb + koff[k]) : 0.f;
    }
  }
}

// float4 variant: 4 consecutive k of one row share their (kh,kw) tap [order 0, C % 4 == 0] or their (c,kh) and
// cover 4 adjacent input pixels [order 1, KW % 4 == 0, sw == 1]: one 16 B load, one 16 B store, a quarter of the
// index arithmetic.  Requires K % 4 == 0 (ldc == K) and 16 B aligned sources.
__global__ void __launch_bounds__(256) im2col_vec4_kernel(ConvGeom g, const float* __restrict__ x, float4* __restrict__ cols,
                                                          long long total4) {
  const int k4n = g.ldc >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % k4n) << 2;
    const long long row = i / k4n;
    int c, kh, kw;
    if (g.order == 0) { c = k % g.C; const int t = k / g.C; kw = t % g.KW; kh = t / g.KW; }
    else { kw = k % g.KW; const int t = k / g.KW; kh = t % g.KH; c = t / g.KH; }
    const int wo = (int)(row % g.Wo);
    const long long t2 = row / g.Wo;
    const int ho = (int)(t2 % g.Ho);
    const long long b = t2 / g.Ho;
    const int h = ho * g.stride - g.pad + kh, w = wo * g.stride - g.pad + kw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h >= 0 && h < g.H) {
      const float* p = x + b * g.sb + h * g.sh + w * g.sw + c * g.sc;
      if (g.order == 0) {
        if (w >= 0 && w < g.W) v = *reinterpret_cast<const float4*>(p);
      } else {
        if (w >= 0 && w + 3 < g.W) v = *reinterpret_cast<const float4*>(p);
        else {
          if (w >= 0 && w < g.W) v.x = p[0];
          if (w + 1 >= 0 && w + 1 < g.W) v.y = p[1];
          if (w + 2 >= 0 && w + 2 < g.W) v.z = p[2];
          if (w + 3 >= 0 && w + 3 < g.W) v.w = p[3];
        }
      }
    }
    cols[i] = v;
  }
}

// float4 col2im for order-0 (NHWC, C % 4 == 0) geometries
__global__ void __launch_bounds__(256) col2im_vec4_kernel(ConvGeom g, const float* __restrict__ dcols, float4* __restrict__ dx,
                                                          long long total4) {
  const int c4n = g.C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) << 2;
    long long t = i / c4n;
    const int w = (int)(t % g.W); t /= g.W;
    const int h = (int)(t % g.H);
    const long long b = t / g.H;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < g.KH; ++kh) {
      const int hn = h + g.pad - kh;
      if (hn < 0 || hn % g.stride) continue;
      const int ho = hn / g.stride;
      if (ho >= g.Ho) continue;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int wn = w + g.pad - kw;
        if (wn < 0 || wn % g.stride) continue;
        const int wo = wn / g.stride;
        if (wo >= g.Wo) continue;
        const float4 v = *reinterpret_cast<const float4*>(dcols + ((b * g.Ho + ho) * g.Wo + wo) * g.ldc + (kh * g.KW + kw) * g.C + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    dx[i] = acc;
  }
}

// ---------------------------------------------------------------- col2im (gather form, no atomics)
// dx[b,h,w,c] (dense NHWC) = sum over (kh,kw) with (h+pad-kh) % stride == 0 of dcols[(b,ho,wo), (kh,kw,c)]
__global__ void __launch_bounds__(256) col2im_kernel(ConvGeom g, const float* __restrict__ dcols, float* __restrict__ dx,
                                                     long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    long long t = i / g.C;
    const int w = (int)(t % g.W); t /= g.W;
    const int h = (int)(t % g.H);
    const long long b = t / g.H;
    float acc = 0.f;
    for (int kh = 0; kh < g.KH; ++kh) {
      const int hn = h + g.pad - kh;
      if (hn < 0 || hn % g.stride) continue;
      const int ho = hn / g.stride;
      if (ho >= g.Ho) continue;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int wn = w + g.pad - kw;
        if (wn < 0 || wn % g.stride) continue;
        const int wo = wn / g.stride;
        if (wo >= g.Wo) continue;
        const int k = g.order == 0 ? (kh * g.KW + kw) * g.C + c : (c * g.KH + kh) * g.KW + kw;
        acc += dcols[((b * g.Ho + ho) * g.Wo + wo) * g.ldc + k];
      }
    }
    dx[i] = acc;
  }
}

// ---------------------------------------------------------------- relu + maxpool 2x2/2 (NHWC)
// a: post-ReLU activations [B,H,W,C] (ReLU applied in the GEMM epilogue); out [B,H/2,W/2,C]; idx: argmax 0..3
__global__ void __launch_bounds__(256) pool_fwd_kernel(const float* __restrict__ a, float* __restrict__ out,
                                                       uint8_t* __restrict__ idx, int H, int W, int C, long long total) {
  const int Ho = H / 2, Wo = W / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const long long b = t / Ho;
    const float* p = a + ((b * H + 2 * ho) * W + 2 * wo) * C + c;
    float best = p[0];
    int bi = 0;
    const float v1 = p[C], v2 = p[(long long)W * C], v3 = p[(long long)W * C + C];
    if (v1 > best) { best = v1; bi = 1; }
    if (v2 > best) { best = v2; bi = 2; }
    if (v3 > best) { best = v3; bi = 3; }
    out[i] = best;
    if (idx) idx[i] = (uint8_t)(best > 0.f ? bi : 4);       // 4: window maximum <= 0, ReLU passes no gradient
  }
}

// da[b,h,w,c] = (argmax of its window == this position && a > 0) ? dout : 0    (dense write of da)
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ idx,
                                                       const float* __restrict__ a, float* __restrict__ da, int H, int W,
                                                       int C, long long total) {
  const int Ho = H / 2, Wo = W / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const long long b = t / H;
    float v = 0.f;
    const int ho = h >> 1, wo = w >> 1;
    if (ho < Ho && wo < Wo) {
      const long long o = ((b * Ho + ho) * Wo + wo) * C + c;
      const int me = (h & 1) * 2 + (w & 1);
      if (idx[o] == me) v = dout[o];                          // idx == 4 encodes relu'(max) == 0
    }
    da[i] = v;
  }
}

// float4 variants (C % 4 == 0, H and W even, < 2^31 pooled elements): one thread per (pooled pixel, 4 channels).
// The ReLU derivative is folded into the index (4 = "no gradient": the window maximum is <= 0), so the backward pass
// reads only dout (a quarter of da) and one index byte per element -- 21 B per 4 da elements instead of 37.
__global__ void __launch_bounds__(256) pool_fwd_vec4_kernel(const float4* __restrict__ a, float4* __restrict__ out,
                                                            uchar4* __restrict__ idx, int H, int W, int C4, unsigned total,
                                                            unsigned int* __restrict__ amax_slot) {
  const int Ho = H / 2, Wo = W / 2;
  float run_max = 0.f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % C4;
    unsigned t = i / C4;
    const unsigned wo = t % Wo; t /= Wo;
    const unsigned ho = t % Ho;
    const unsigned b = t / Ho;
    const float4* p = a + ((size_t)(b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 v0 = p[0], v1 = p[C4], v2 = p[(size_t)W * C4], v3 = p[(size_t)W * C4 + C4];
    float4 best;
    uchar4 bi;
#define DDRL_POOL1(f)                                           \
    {                                                           \
      float m = v0.f; int k = 0;                                \
      if (v1.f > m) { m = v1.f; k = 1; }                        \
      if (v2.f > m) { m = v2.f; k = 2; }                        \
      if (v3.f > m) { m = v3.f; k = 3; }                        \
      best.f = m; bi.f = (unsigned char)(m > 0.f ? k : 4);      \
    }
    DDRL_POOL1(x) DDRL_POOL1(y) DDRL_POOL1(z) DDRL_POOL1(w)
#undef DDRL_POOL1
    out[i] = best;
    if (idx) idx[i] = bi;
    run_max = fmaxf(run_max, fmaxf(fmaxf(fabsf(best.x), fabsf(best.y)), fmaxf(fabsf(best.z), fabsf(best.w))));
  }
  if (amax_slot != nullptr) amax_commit(amax_slot, run_max);
}
__global__ void __launch_bounds__(256) pool_bwd_vec4_kernel(const float4* __restrict__ dout, const uchar4* __restrict__ idx,
                                                            float4* __restrict__ da, int H, int W, int C4, unsigned total,
                                                            unsigned int* __restrict__ amax_slot) {
  const int Ho = H / 2, Wo = W / 2;
  float run_max = 0.f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % C4;
    unsigned t = i / C4;
    const unsigned wo = t % Wo; t /= Wo;
    const unsigned ho = t % Ho;
    const unsigned b = t / Ho;
    const float4 g = dout[i];
    const uchar4 k = idx[i];
    float4* p = da + ((size_t)(b * H + 2 * ho) * W + 2 * wo) * C4 + c;
#define DDRL_SEL(pos) make_float4(k.x == pos ? g.x : 0.f, k.y == pos ? g.y : 0.f, k.z == pos ? g.z : 0.f, k.w == pos ? g.w : 0.f)
    p[0] = DDRL_SEL(0);
    p[C4] = DDRL_SEL(1);
    p[(size_t)W * C4] = DDRL_SEL(2);
    p[(size_t)W * C4 + C4] = DDRL_SEL(3);
#undef DDRL_SEL
    // what reaches da is g where the argmax survived the ReLU (k < 4), zero elsewhere
    run_max = fmaxf(run_max, fmaxf(fmaxf(k.x < 4 ? fabsf(g.x) : 0.f, k.y < 4 ? fabsf(g.y) : 0.f),
                                   fmaxf(k.z < 4 ? fabsf(g.z) : 0.f, k.w < 4 ? fabsf(g.w) : 0.f)));
  }
  if (amax_slot != nullptr) amax_commit(amax_slot, run_max);
}

// dy *= act'(y) from the activation OUTPUT y: relu: y>0 ; leaky: y>0 ? 1 : 0.01 (y==0 -> 0.01)
__global__ void __launch_bounds__(256) act_bwd_kernel(float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y,
                                                      int rows, int colsN, int act) {
  const long long total = (long long)rows * colsN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % colsN);
    const long long r = i / colsN;
    const float yy = y[r * ld_y + c];
    float* d = dy + r * ld_dy + c;
    if (act == 1) { if (!(yy > 0.f)) *d = 0.f; }
    else if (act == 2) { if (!(yy > 0.f)) *d *= 0.01f; }
  }
}

// ---------------------------------------------------------------- bias gradient: db[n] += sum_rows dy[r, n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, int ld, long long rows, int N,
                                                     float* __restrict__ db, long long rows_per_block, DetSeq det) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int n = blockIdx.x * 32 + tx;
  const long long r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (n < N)
    for (long long r = r0 + ty; r < r1; r += 8) acc += dy[r * ld + n];
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][tx];
    part[0][tx] = s;
  }
  // deterministic mode: the row blocks of one column group add in block order (turn counter per blockIdx.x)
  if (det.ctr && threadIdx.x == 0) det_enter(det.ctr + blockIdx.x, blockIdx.y);
  __syncthreads();
  if (ty == 0 && n < N) det_add(det.ctr != nullptr, db + n, part[0][tx]);
  if (det.ctr) {
    __syncthreads();
    if (threadIdx.x == 0) det_leave(det.ctr + blockIdx.x, blockIdx.y, gridDim.y);
  }
}

// float4 variant: thread = (row lane, 4 columns); 4 rows in flight per thread; N % 4 == 0, N <= 1024, 16-byte aligned rows
__global__ void __launch_bounds__(256) colsum4_kernel(const float4* __restrict__ dy, int ld4, long long rows, int n4,
                                                      float* __restrict__ db, long long rows_per_block,
                                                      float* __restrict__ db2, int split4, DetSeq det) {
  __shared__ float4 part[256];
  const int ry = 256 / n4;                                // row lanes per block (n4 is a power-of-two divisor of 256 or < 256)
  const int tx = threadIdx.x % n4, ty = threadIdx.x / n4;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ty < ry) {
    long long r = r0 + ty;
    for (; r + 3LL * ry < r1; r += 4LL * ry) {
      const float4 a = __ldg(dy + r * ld4 + tx), b = __ldg(dy + (r + ry) * ld4 + tx), c = __ldg(dy + (r + 2LL * ry) * ld4 + tx),
                   d = __ldg(dy + (r + 3LL * ry) * ld4 + tx);
      acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += ry) {
      const float4 a = __ldg(dy + r * ld4 + tx);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  part[threadIdx.x] = acc;
  if (det.ctr && threadIdx.x == 0) det_enter(det.ctr, blockIdx.x);      // deterministic mode: row blocks add in block order
  __syncthreads();
  if (ty == 0) {
    float4 sum = part[tx];
    for (int k = 1; k < ry; ++k) {
      const float4 v = part[k * n4 + tx];
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    // columns [0, 4 split4) -> db, the rest -> db2 (two layers' biases behind ONE pass over a fused [rows, N0 + N1] gradient)
    float* out = (db2 != nullptr && tx >= split4) ? db2 + 4 * (tx - split4) : db + 4 * tx;
    const bool ord = det.ctr != nullptr;
    det_add(ord, out, sum.x); det_add(ord, out + 1, sum.y);
    det_add(ord, out + 2, sum.z); det_add(ord, out + 3, sum.w);
  }
  if (det.ctr) {
    __syncthreads();
    if (threadIdx.x == 0) det_leave(det.ctr, blockIdx.x, gridDim.x);
  }
}

// ---------------------------------------------------------------- weight packing
// bodies: prep_kernels.cuh (shared with the multi-job kernel of prep.cu)
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int J,
                                                   int ld) {
  pack_body(src, dst, O, I, J, ld, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I,
                                                     int J, int ld) {
  unpack_body(src, dst, O, I, J, ld, blockIdx.x, gridDim.x);
}

// dst[r*ld_d + c] = src[r*ld_s + c]
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, int ld_s, float* __restrict__ dst, int ld_d,
                                                     long long rows, int colsN) {
  const long long total = rows * colsN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % colsN);
    const long long r = i / colsN;
    dst[r * ld_d + c] = src[r * ld_s + c];
  }
}

// ---------------------------------------------------------------- skinny linears (heads, N <= 32)
// y[b, n] = x[b,:] . W[n,:] + bias[n]; one warp per row, lanes stride K (coalesced), N warp reductions.
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int B, int N, int K,
                                                         float* __restrict__ y, int ldy) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* xr = x + (size_t)warp * ldx;
  for (int n = 0; n < N; ++n) {
    const float* w = W + (size_t)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(xr[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[(size_t)warp * ldy + n] = acc + (bias ? bias[n] : 0.f);
  }
}
// dx[b, k] (+)= sum_n dy[b, n] * W[n, k]
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ W,
                                                           int B, int N, int K, float* __restrict__ dx, int ldx,
                                                           int accumulate) {
  const long long total = (long long)B * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long b = i / K;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(dy[b * ldy + n], W[(size_t)n * K + k], acc);
    float* d = dx + b * ldx + k;
    *d = accumulate ? *d + acc : acc;
  }
}
// dW[n, k] += sum_b dy[b, n] * x[b, k] ; db[n] += sum_b dy[b, n].  grid = (ceil(K/256), row chunks, N)
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x,
                                                           int ldx, int B, int N, int K, float* __restrict__ dW,
                                                           float* __restrict__ db, int rows_per_block, DetSeq det) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(B, r0 + rows_per_block);
  float acc = 0.f, accb = 0.f;
  if (k < K) {
    for (int r = r0; r < r1; ++r) {
      const float d = dy[(size_t)r * ldy + n];
      acc = fmaf(d, x[(size_t)r * ldx + k], acc);
      accb += d;
    }
  }
  // deterministic mode: the row chunks of one (k block, n) add in chunk order
  unsigned int* turn = det.ctr ? det.ctr + (size_t)blockIdx.z * gridDim.x + blockIdx.x : nullptr;
  if (turn) {
    if (threadIdx.x == 0) det_enter(turn, blockIdx.y);
    __syncthreads();
  }
  if (k < K) {
    det_add(turn != nullptr, dW + (size_t)n * K + k, acc);
    if (db && k == 0) det_add(turn != nullptr, db + n, accb);
  }
  if (turn) {
    __syncthreads();
    if (threadIdx.x == 0) det_leave(turn, blockIdx.y, gridDim.y);
  }
}

// Register-tiled forms of the three head kernels for the common shapes (K <= 512, N <= 8 -- the 6-way / 2-d / 1-wide heads on
// the 512-wide feature): every operand element is read ONCE per block (the general kernels above re-read x per output
// column and decompose a flat index with 64-bit divisions per element: 15 - 30 us per launch on 8192 rows, 2.3 % of a
// Pong step).
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_fwd_reg_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                                             const float* __restrict__ bias, int B, int N, int K,
                                                             float* __restrict__ y, int ldy) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* xr = x + (size_t)warp * ldx;
  float xv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) xv[j] = (lane + 32 * j) < K ? xr[lane + 32 * j] : 0.f;
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    if (n < N) {
      const float* w = W + (size_t)n * K;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if ((lane + 32 * j) < K) acc = fmaf(xv[j], __ldg(w + lane + 32 * j), acc);
      acc = warp_sum(acc);
      if (lane == 0) y[(size_t)warp * ldy + n] = acc + (bias ? bias[n] : 0.f);
    }
  }
}
// dx[b, 4 k4 ..] (+)= sum_n dy[b, n] * W[n, 4 k4 ..]: thread = (row lane, 4 consecutive k), its N x 4 weights in registers
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_dgrad_reg_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ W,
                                                               int B, int N, int K, float* __restrict__ dx, int ldx, int accumulate) {
  const int k4n = K >> 2;                                 // float4 columns per row (K % 4 == 0, k4n <= 256)
  const int k4 = threadIdx.x % k4n, rl = threadIdx.x / k4n, rpb = 256 / k4n;
  if (rl >= rpb) return;
  float4 w[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    w[n] = n < N ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long b = (long long)blockIdx.x * rpb + rl; b < B; b += (long long)gridDim.x * rpb) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < N) {
        const float d = __ldg(dy + b * ldy + n);
        acc.x = fmaf(d, w[n].x, acc.x); acc.y = fmaf(d, w[n].y, acc.y); acc.z = fmaf(d, w[n].z, acc.z); acc.w = fmaf(d, w[n].w, acc.w);
      }
    }
    float4* p = reinterpret_cast<float4*>(dx + b * ldx) + k4;
    if (accumulate) { const float4 o = *p; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *p = acc;
  }
}
// dW[n, k] += sum_b dy[b, n] * x[b, k]; db[n] += sum_b dy[b, n]: thread = one k, N accumulators; grid = (K / 256, row chunks)
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_wgrad_reg_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x,
                                                               int ldx, int B, int N, int K, float* __restrict__ dW,
                                                               float* __restrict__ db, int rows_per_block, DetSeq det) {
  __shared__ float sdy[64 * NMAX];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(B, r0 + rows_per_block);
  float acc[NMAX], accb = 0.f;
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  for (int t0 = r0; t0 < r1; t0 += 64) {
    const int nr = min(64, r1 - t0);
    for (int i = threadIdx.x; i < nr * NMAX; i += 256) {
      const int r = i / NMAX, n = i - r * NMAX;
      sdy[i] = n < N ? __ldg(dy + (size_t)(t0 + r) * ldy + n) : 0.f;
    }
    __syncthreads();
    if (k < K) {
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const float xv = __ldg(x + (size_t)(t0 + r) * ldx + k);
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[n] = fmaf(sdy[r * NMAX + n], xv, acc[n]);
      }
    }
    if (db && blockIdx.x == 0 && (int)threadIdx.x < N)
      for (int r = 0; r < nr; ++r) accb += sdy[r * NMAX + threadIdx.x];
    __syncthreads();
  }
  // deterministic mode: the row chunks of one k block add in chunk order
  unsigned int* turn = det.ctr ? det.ctr + blockIdx.x : nullptr;
  if (turn) {
    if (threadIdx.x == 0) det_enter(turn, blockIdx.y);
    __syncthreads();
  }
  if (k < K) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) det_add(turn != nullptr, dW + (size_t)n * K + k, acc[n]);
  }
  if (db && blockIdx.x == 0 && (int)threadIdx.x < N) det_add(turn != nullptr, db + threadIdx.x, accb);
  if (turn) {
    __syncthreads();
    if (threadIdx.x == 0) det_leave(turn, blockIdx.y, gridDim.y);
  }
}

// ---------------------------------------------------------------- space-to-depth (first-layer strided convs)
// A valid (pad 0) convolution of an NCHW observation whose kernel extents and image extents are multiples of its stride s
// (NatureCNN conv1: 8x8 stride 4 on 84x84, nn/atari_encoder.py:16) equals a stride-1 convolution with a (KH/s) x (KW/s)
// kernel over the space-to-depth tensor  X2[b, Y, X, (i*s + j)*C + c] = x[b, c, s*Y + i, s*X + j]  -- an NHWC tensor of
// exactly the observation's size (the im2col matrix it replaces is KH*KW/s^2 times larger), which the implicit-GEMM
// (tap-TMA) path consumes directly.  One block per (b, Y) band: coalesced row reads -> shared memory -> one contiguous
// W2*C2 run of the output.
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W,
                                                  int s, long long sb, long long bands) {
  extern __shared__ float band[];                       // [C][s][W]
  const int H2 = H / s, W2 = W / s, C2 = s * s * C;
  const int n_in = C * s * W;
  for (long long bd = blockIdx.x; bd < bands; bd += gridDim.x) {
    const long long b = bd / H2;
    const int Y = (int)(bd - b * H2);
    const float* xb = x + b * sb + (long long)Y * s * W;
    for (int e = threadIdx.x; e < n_in; e += blockDim.x) {
      const int w = e % W, ci = e / W;                  // ci = c*s + i
      const int c = ci / s, i = ci - c * s;
      band[e] = xb[(long long)c * H * W + (long long)i * W + w];
    }
    __syncthreads();
    float* ob = out + bd * (long long)W2 * C2;
    for (int e = threadIdx.x; e < W2 * C2; e += blockDim.x) {
      const int ch = e % C2, X = e / C2;
      const int c = ch % C, ij = ch / C;
      const int j = ij % s, i = ij / s;
      ob[e] = band[(c * s + i) * W + X * s + j];
    }
    __syncthreads();
  }
}
// C == 4, W % 4 == 0 (NatureCNN frames): 128-bit loads and stores, BANDS bands in flight per block iteration.  One output
// float4 = the 4 channels of one (i, j) position.
// CH / CW / CS > 0: the frame size and stride as compile-time constants (NatureCNN: 84 x 84, stride 4).  The kernel decomposes a
// flat slot index with seven integer divisions per 16-byte access; with run-time divisors that was 45 % of its issue slots
// (64 % busy) and it ran at 2.9 TB/s (profiles/r3g_*) -- constant divisors compile to multiply-shift.
template <int BANDS, int CH = 0, int CW = 0, int CS = 0>
__global__ void __launch_bounds__(256) s2d_c4_kernel(const float* __restrict__ x, float* __restrict__ out, int H_, int W_, int s_,
                                                     long long sb, long long bands, unsigned int* __restrict__ amax_slot) {
  extern __shared__ float4 band4[];                     // [BANDS][4*s][W/4]
  float* band = reinterpret_cast<float*>(band4);
  const int H = CH > 0 ? CH : H_, W = CW > 0 ? CW : W_, s = CS > 0 ? CS : s_;
  const int H2 = H / s, W2 = W / s, W4 = W / 4, rows = 4 * s, C2 = s * s * 4;
  const int n_in4 = rows * W4, n_out4 = W2 * s * s;
  float run_max = 0.f;
  for (long long bd0 = (long long)blockIdx.x * BANDS; bd0 < bands; bd0 += (long long)gridDim.x * BANDS) {
    const int nb = (int)min((long long)BANDS, bands - bd0);
    for (int e = threadIdx.x; e < nb * n_in4; e += blockDim.x) {
      const int k = e / n_in4, r4 = e - k * n_in4;
      const int ci = r4 / W4, w4 = r4 - ci * W4;        // ci = c*s + i
      const int c = ci / s, i = ci - c * s;
      const long long bd = bd0 + k;
      const long long b = CH > 0 ? (long long)((unsigned int)bd / (unsigned int)H2) : bd / H2;      // (bands < 2^31 whenever CH > 0)
      const int Y = (int)(bd - b * H2);
      band4[e] = __ldg(reinterpret_cast<const float4*>(x + b * sb + (long long)c * H * W + (long long)(Y * s + i) * W) + w4);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb * n_out4; e += blockDim.x) {
      const int k = e / n_out4, r4 = e - k * n_out4;
      const int X = r4 / (s * s), ij = r4 - X * (s * s);
      const int i = ij / s, j = ij - i * s;
      const float* bp = band + (size_t)k * rows * W + i * W + X * s + j;
      const float4 v = make_float4(bp[0], bp[(size_t)s * W], bp[(size_t)2 * s * W], bp[(size_t)3 * s * W]);
      reinterpret_cast<float4*>(out + (bd0 + k) * (long long)W2 * C2)[r4] = v;
      run_max = fmaxf(run_max, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    __syncthreads();
  }
  if (amax_slot != nullptr) amax_commit(amax_slot, run_max);
}
__global__ void __launch_bounds__(256) pack_s2d_kernel(const float* __restrict__ w, float* __restrict__ dst, int O, int C, int KH,
                                                       int KW, int s, int ld, int unpack) {
  pack_s2d_body(w, dst, O, C, KH, KW, s, ld, unpack, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------- thin-K layers (first convs on 1..4-channel maps)
// y[m, n] = act(sum_k x[m, k] W[n, k] + b[n]) and dW[n, k] += sum_m dy[m, n] x[m, k] for K <= 36 (K4 = ldx/4 float4 per
// row), N = 4*NC4 <= 64.  These layers are pure HBM streams (y / dy are 64 floats per row, x is 12): a tensor-core tile
// would spend one whole K block and a full epilogue on 9 useful k values.  Thread = (row lane, 4 output channels); its
// 4 x K weights (forward) or 4 x K partial sums (weight gradient) live in registers for the whole launch.
template <int K4, int NC4>
__global__ void __launch_bounds__(256) thin_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ W, int ldw,
                                                       const float* __restrict__ bias, float4* __restrict__ y, long long M,
                                                       int K, int act) {
  constexpr int RL = 256 / NC4;                         // rows in flight per block
  const int c4 = threadIdx.x % NC4, rl = threadIdx.x / NC4;
  float w[4][K4 * 4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K4 * 4; ++k) w[j][k] = k < K ? W[(size_t)(c4 * 4 + j) * ldw + k] : 0.f;
  const float4 b = bias ? make_float4(bias[c4 * 4], bias[c4 * 4 + 1], bias[c4 * 4 + 2], bias[c4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  // UR rows per thread and iteration: all their x loads are issued before the first FMA (memory-level parallelism; with
  // ~60 live registers only 3-4 blocks fit an SM, so one row in flight per thread leaves HBM idle)
  constexpr int UR = K4 <= 3 ? 4 : 2;
  const long long step = (long long)gridDim.x * RL;
  for (long long m0 = (long long)blockIdx.x * RL + rl; m0 < M; m0 += step * UR) {
    float4 v[UR][K4];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long m = m0 + u * step;
#pragma unroll
      for (int q = 0; q < K4; ++q) v[u][q] = m < M ? __ldg(x + m * K4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long m = m0 + u * step;
      if (m >= M) break;
      float4 acc = b;
#pragma unroll
      for (int q = 0; q < K4; ++q) {
        const float xv[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc.x = fmaf(xv[e], w[0][q * 4 + e], acc.x); acc.y = fmaf(xv[e], w[1][q * 4 + e], acc.y);
          acc.z = fmaf(xv[e], w[2][q * 4 + e], acc.z); acc.w = fmaf(xv[e], w[3][q * 4 + e], acc.w);
        }
      }
      if (act == 1) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      else if (act == 2) {
        acc.x = acc.x > 0.f ? acc.x : 0.01f * acc.x; acc.y = acc.y > 0.f ? acc.y : 0.01f * acc.y;
        acc.z = acc.z > 0.f ? acc.z : 0.01f * acc.z; acc.w = acc.w > 0.f ? acc.w : 0.01f * acc.w;
      }
      y[m * NC4 + c4] = acc;
    }
  }
}
template <int K4, int NC4>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                         float* __restrict__ dW, int ldw, long long M, int K, DetSeq det) {
  constexpr int RL = 256 / NC4;
  __shared__ float red[NC4 * 4 * K4 * 4];
  const int c4 = threadIdx.x % NC4, rl = threadIdx.x / NC4;
  float acc[4][K4 * 4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K4 * 4; ++k) acc[j][k] = 0.f;
  constexpr int UR = K4 <= 3 ? 4 : 2;
  const long long step = (long long)gridDim.x * RL;
  for (long long m0 = (long long)blockIdx.x * RL + rl; m0 < M; m0 += step * UR) {
    float4 v[UR][K4], g[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long m = m0 + u * step;
      g[u] = m < M ? __ldg(dy + m * NC4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < K4; ++q) v[u][q] = m < M ? __ldg(x + m * K4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
#pragma unroll
      for (int q = 0; q < K4; ++q) {
        const float xv[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[0][q * 4 + e] = fmaf(g[u].x, xv[e], acc[0][q * 4 + e]); acc[1][q * 4 + e] = fmaf(g[u].y, xv[e], acc[1][q * 4 + e]);
          acc[2][q * 4 + e] = fmaf(g[u].z, xv[e], acc[2][q * 4 + e]); acc[3][q * 4 + e] = fmaf(g[u].w, xv[e], acc[3][q * 4 + e]);
        }
      }
    }
  }
  // block reduction without shared-memory atomics: lanes l and l ^ 16 (NC4 = 16) / l ^ 8, l ^ 16 (NC4 = 8) hold the same
  // channels -> shuffle; then one partial per warp in shared memory, summed by the first NC4*16*K4 threads
  constexpr int NV = NC4 * 4 * K4 * 4;                  // outputs per block
  constexpr int NW = 8;                                 // warps
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K4 * 4; ++k) {
      float v = acc[j][k];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (NC4 == 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
      acc[j][k] = v;
    }
  // the 8 warps add their partials into ONE [NV] buffer in turn (8 barriers, once per block)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int w = 0; w < NW; ++w) {
    if (wid == w && lane < NC4) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < K4 * 4; ++k) {
          float* p = red + (lane * 4 + j) * (K4 * 4) + k;
          *p = (w == 0 ? 0.f : *p) + acc[j][k];
        }
    }
    __syncthreads();
  }
  if (det.ctr) {                                         // deterministic mode: the blocks add in block order
    if (threadIdx.x == 0) det_enter(det.ctr, blockIdx.x);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NV; i += 256) {
    const int n = i / (K4 * 4), k = i % (K4 * 4);
    if (k < K) det_add(det.ctr != nullptr, dW + (size_t)n * ldw + k, red[i]);
  }
  if (det.ctr) {
    __syncthreads();
    if (threadIdx.x == 0) det_leave(det.ctr, blockIdx.x, gridDim.x);
  }
}

// ---------------------------------------------------------------- launchers
static inline int grid_for(long long total, int threads = 256) {
  return (int)std::min<long long>((total + threads - 1) / threads, 32LL * kNumSMs);
}

int im2col(const ConvGeom& g, const float* x, float* cols, int B, cudaStream_t s) {
  const long long total = (long long)B * g.Ho * g.Wo * g.ldc;
  if (total == 0) return DDRL_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(cols)) & 15) == 0 && g.ldc == g.K;
  const bool v0 = g.order == 0 && g.C % 4 == 0;
  const bool v1 = g.order == 1 && g.KW % 4 == 0 && g.sw == 1 && g.stride % 4 == 0 && g.pad % 4 == 0 && g.sh % 4 == 0 &&
                  g.sc % 4 == 0 && g.sb % 4 == 0;
  prof_work(4.0 * total + 4.0 * (double)B * g.H * g.W * g.C);      // write cols + read input once
  if (aligned && (v0 || v1)) {
    im2col_vec4_kernel<<<grid_for(total / 4), 256, 0, s>>>(g, x, reinterpret_cast<float4*>(cols), total / 4);
    DDRL_LAUNCHED("im2col_vec4_kernel");
    return DDRL_OK;
  }
  if (g.ldc <= 1024 && (long long)g.C * g.sc + (long long)g.KH * g.sh + (long long)g.KW * g.sw < (1LL << 30)) {
    int kt = 1;
    while (kt < g.ldc && kt < 32) kt <<= 1;
    const long long rows = (long long)B * g.Ho * g.Wo;
    const int rpb = 256 / kt;
    im2col_rows_kernel<<<(int)std::min<long long>((rows + rpb - 1) / rpb, 16LL * kNumSMs), 256, 3 * g.ldc * sizeof(int), s>>>(
        g, x, cols, rows, kt);
    DDRL_LAUNCHED("im2col_kernel");
    return DDRL_OK;
  }
  im2col_kernel<<<grid_for(total), 256, 0, s>>>(g, x, cols, total);
  DDRL_LAUNCHED("im2col_kernel");
  return DDRL_OK;
}
int col2im(const ConvGeom& g, const float* dcols, float* dx, int B, cudaStream_t s) {
  const long long total = (long long)B * g.H * g.W * g.C;
  if (total == 0) return DDRL_OK;
  prof_work(4.0 * total + 4.0 * (double)B * g.Ho * g.Wo * g.ldc);
  if (g.order == 0 && g.C % 4 == 0 && g.ldc % 4 == 0 && ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dcols)) & 15) == 0) {
    col2im_vec4_kernel<<<grid_for(total / 4), 256, 0, s>>>(g, dcols, reinterpret_cast<float4*>(dx), total / 4);
    DDRL_LAUNCHED("col2im_vec4_kernel");
    return DDRL_OK;
  }
  col2im_kernel<<<grid_for(total), 256, 0, s>>>(g, dcols, dx, total);
  DDRL_LAUNCHED("col2im_kernel");
  return DDRL_OK;
}
static inline bool pool_vec_ok(const void* p0, const void* p1, const void* p2, int H, int W, int C, long long total) {
  return C % 4 == 0 && H % 2 == 0 && W % 2 == 0 && total / 4 < 0x7fffffffLL &&
         ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p2) & 3) == 0;
}
int pool_fwd(const float* a, float* out, uint8_t* idx, int B, int H, int W, int C, cudaStream_t s, float* amax_slot) {
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  if (total == 0) return DDRL_OK;
  prof_work(4.0 * (double)B * H * W * C + 5.0 * total);
  if (pool_vec_ok(a, out, idx, H, W, C, total)) {
    pool_fwd_vec4_kernel<<<grid_for(total / 4), 256, 0, s>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<float4*>(out),
                                                             reinterpret_cast<uchar4*>(idx), H, W, C / 4, (unsigned)(total / 4),
                                                             reinterpret_cast<unsigned int*>(amax_slot));
    DDRL_LAUNCHED("pool_fwd_kernel");
    return DDRL_OK;
  }
  pool_fwd_kernel<<<grid_for(total), 256, 0, s>>>(a, out, idx, H, W, C, total);
  DDRL_LAUNCHED("pool_fwd_kernel");
  if (amax_slot) return amax_f32(out, total / C, C, C, amax_slot, false, s);
  return DDRL_OK;
}
// `a` (the pooled layer's input) is no longer read: relu'(a) at the argmax is encoded in idx by pool_fwd
int pool_bwd(const float* dout, const uint8_t* idx, const float* a, float* da, int B, int H, int W, int C, cudaStream_t s,
             float* amax_slot) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return DDRL_OK;
  const long long pooled = (long long)B * (H / 2) * (W / 2) * C;
  prof_work(4.0 * (double)total + 5.0 * pooled);
  if (pool_vec_ok(dout, da, idx, H, W, C, pooled)) {
    pool_bwd_vec4_kernel<<<grid_for(pooled / 4), 256, 0, s>>>(reinterpret_cast<const float4*>(dout),
                                                              reinterpret_cast<const uchar4*>(idx), reinterpret_cast<float4*>(da),
                                                              H, W, C / 4, (unsigned)(pooled / 4),
                                                              reinterpret_cast<unsigned int*>(amax_slot));
    DDRL_LAUNCHED("pool_bwd_kernel");
    return DDRL_OK;
  }
  pool_bwd_kernel<<<grid_for(total), 256, 0, s>>>(dout, idx, a, da, H, W, C, total);
  DDRL_LAUNCHED("pool_bwd_kernel");
  if (amax_slot) return amax_f32(da, total / C, C, C, amax_slot, false, s);
  return DDRL_OK;
}
int act_bwd(float* dy, int ld_dy, const float* y, int ld_y, long long rows, int colsN, int act, cudaStream_t s) {
  if (act == 0 || rows * colsN == 0) return DDRL_OK;
  act_bwd_kernel<<<grid_for(rows * colsN), 256, 0, s>>>(dy, ld_dy, y, ld_y, (int)rows, colsN, act);
  DDRL_LAUNCHED("act_bwd_kernel");
  return DDRL_OK;
}
int colsum_add(const float* dy, int ld, long long rows, int N, float* db, cudaStream_t s, float* db2, int split) {
  if (rows == 0 || N == 0) return DDRL_OK;
  if (db2 && !(N % 4 == 0 && split % 4 == 0 && ld % 4 == 0 && N <= 1024 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0)) {
    { const int rc = colsum_add(dy, ld, rows, split, db, s, nullptr, 0); if (rc != DDRL_OK) return rc; }
    return colsum_add(dy + split, ld, rows, N - split, db2, s, nullptr, 0);
  }
  prof_work(4.0 * (double)rows * N);
  if (N % 4 == 0 && ld % 4 == 0 && N <= 1024 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const int n4 = N / 4;
    const int ry = std::max(1, 256 / n4);
    const DetSeq det = det_seq(1);
    long long chunks = std::max<long long>(1, std::min<long long>((rows + 16LL * ry - 1) / (16LL * ry), det.ctr ? kDetMaxParts : 8LL * kNumSMs));
    const long long rpb = (rows + chunks - 1) / chunks;
    chunks = (rows + rpb - 1) / rpb;
    if (n4 <= 256) {
      colsum4_kernel<<<(unsigned)chunks, 256, 0, s>>>(reinterpret_cast<const float4*>(dy), ld / 4, rows, n4, db, rpb, db2, split / 4, det);
      DDRL_LAUNCHED("colsum_kernel");
      return DDRL_OK;
    }
  }
  const int nb = ceil_div(N, 32);
  const DetSeq det = det_seq(nb);
  long long chunks = std::max<long long>(1, std::min<long long>((rows + 255) / 256, det.ctr ? kDetMaxParts : (4LL * kNumSMs + nb - 1) / nb));
  const long long rpb = (rows + chunks - 1) / chunks;
  chunks = (rows + rpb - 1) / rpb;
  colsum_kernel<<<dim3(nb, (unsigned)chunks), 256, 0, s>>>(dy, ld, rows, N, db, rpb, det);
  DDRL_LAUNCHED("colsum_kernel");
  return DDRL_OK;
}
int pack_weight(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s) {
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_PACK; j.a = src; j.b = dst; j.total = (long long)O * I * J; j.vblocks = prep_blocks(j.total);
    j.i[0] = O; j.i[1] = I; j.i[2] = J; j.i[3] = ld;
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  pack_kernel<<<grid_for((long long)O * I * J), 256, 0, s>>>(src, dst, O, I, J, ld);
  DDRL_LAUNCHED("pack_kernel");
  return DDRL_OK;
}
int unpack_grad(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s) {
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_UNPACK; j.a = src; j.b = dst; j.total = (long long)O * I * J; j.vblocks = prep_blocks(j.total);
    j.i[0] = O; j.i[1] = I; j.i[2] = J; j.i[3] = ld;
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  unpack_kernel<<<grid_for((long long)O * I * J), 256, 0, s>>>(src, dst, O, I, J, ld);
  DDRL_LAUNCHED("unpack_kernel");
  return DDRL_OK;
}
bool thin_supported(long long M, int N, int K, const float* x, int ldx, const float* y, int ldy) {
  if (M < 1 || ldy != N || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0) return false;
  const int k4 = ldx / 4, nc4 = N / 4;
  if (ldx % 4 != 0 || N % 4 != 0 || K > ldx) return false;
  return (nc4 == 16 && (k4 == 3 || k4 == 9)) || (nc4 == 8 && (k4 == 2 || k4 == 4));
}
#define DDRL_THIN_DISPATCH(KERNEL, ...)                                                        \
  do {                                                                                         \
    const int k4 = ldx / 4, nc4 = N / 4;                                                       \
    const int grid = (int)std::min<long long>((M + 256 / nc4 - 1) / (256 / nc4), thin_grid_cap); \
    if (nc4 == 16 && k4 == 3) KERNEL<3, 16><<<grid, 256, 0, s>>>(__VA_ARGS__);                \
    else if (nc4 == 16 && k4 == 9) KERNEL<9, 16><<<grid, 256, 0, s>>>(__VA_ARGS__);           \
    else if (nc4 == 8 && k4 == 4) KERNEL<4, 8><<<grid, 256, 0, s>>>(__VA_ARGS__);  /* 3 x 960 laser variant: K = 15 */ \
    else KERNEL<2, 8><<<grid, 256, 0, s>>>(__VA_ARGS__);                                      \
  } while (0)
int thin_fwd(const float* x, int ldx, const float* W, int ldw, const float* bias, float* y, long long M, int N, int K, int act,
             cudaStream_t s) {
  if (!thin_supported(M, N, K, x, ldx, y, N)) return DDRL_E_UNSUPPORTED;
  prof_work(4.0 * (double)M * (ldx + N));
  const long long thin_grid_cap = 8LL * kNumSMs;
  DDRL_THIN_DISPATCH(thin_fwd_kernel, reinterpret_cast<const float4*>(x), W, ldw, bias, reinterpret_cast<float4*>(y), M, K, act);
  DDRL_LAUNCHED("thin_fwd_kernel");
  return DDRL_OK;
}
int thin_wgrad(const float* x, int ldx, const float* dy, float* dW, int ldw, long long M, int N, int K, cudaStream_t s) {
  if (!thin_supported(M, N, K, x, ldx, dy, N)) return DDRL_E_UNSUPPORTED;
  prof_work(4.0 * (double)M * (ldx + N));
  const DetSeq det = det_seq(1);
  const long long thin_grid_cap = det.ctr ? kDetMaxParts : 8LL * kNumSMs;
  DDRL_THIN_DISPATCH(thin_wgrad_kernel, reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), dW, ldw, M, K, det);
  DDRL_LAUNCHED("thin_wgrad_kernel");
  return DDRL_OK;
}
#undef DDRL_THIN_DISPATCH
int space_to_depth(const ConvGeom& g, const float* x, float* out, int B, cudaStream_t s, float* amax_slot) {
  if (g.order != 1 || g.sw != 1 || g.stride < 1 || g.H % g.stride || g.W % g.stride) return DDRL_E_ARG;
  const long long bands = (long long)B * (g.H / g.stride);
  if (bands == 0) return DDRL_OK;
  const size_t smem = sizeof(float) * (size_t)g.C * g.stride * g.W;
  if (smem > 48 * 1024) return DDRL_E_UNSUPPORTED;
  prof_work(8.0 * (double)B * g.C * g.H * g.W);
  unsigned int* am = reinterpret_cast<unsigned int*>(amax_slot);
  if (g.C == 4 && g.W % 4 == 0 && g.sb % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
      4 * smem <= 48 * 1024) {
    constexpr int BANDS = 4;
    const long long groups = (bands + BANDS - 1) / BANDS;
    const int grid = (int)std::min<long long>(groups, 8LL * kNumSMs);
    if (g.H == 84 && g.W == 84 && g.stride == 4 && bands < (1LL << 31))
      s2d_c4_kernel<BANDS, 84, 84, 4><<<grid, 256, BANDS * smem, s>>>(x, out, g.H, g.W, g.stride, g.sb, bands, am);
    else
      s2d_c4_kernel<BANDS><<<grid, 256, BANDS * smem, s>>>(x, out, g.H, g.W, g.stride, g.sb, bands, am);
    DDRL_LAUNCHED("s2d_kernel");
    return DDRL_OK;
  }
  s2d_kernel<<<(int)std::min<long long>(bands, 16LL * kNumSMs), 256, smem, s>>>(x, out, g.C, g.H, g.W, g.stride, g.sb, bands);
  DDRL_LAUNCHED("s2d_kernel");
  if (amax_slot) {                               // generic layout: the reduction runs as a pass of its own
    const int rc = amax_f32(out, (long long)B * (g.H / g.stride) * (g.W / g.stride), g.C * g.stride * g.stride,
                            (long long)g.C * g.stride * g.stride, amax_slot, false, s);
    if (rc != DDRL_OK) return rc;
  }
  return DDRL_OK;
}
static int record_s2d(const float* src, float* dst, int O, int C, int KH, int KW, int stride, int ld, int unpack) {
  PrepJob j{}; j.type = PREP_PACK_S2D; j.a = src; j.b = dst; j.total = (long long)O * C * KH * KW; j.vblocks = prep_blocks(j.total);
  j.i[0] = O; j.i[1] = C; j.i[2] = KH; j.i[3] = KW; j.i[4] = stride; j.i[5] = ld; j.i[6] = unpack;
  return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
}
int pack_weight_s2d(const float* w_oihw, float* dst, int O, int C, int KH, int KW, int stride, int ld, cudaStream_t s) {
  if (g_prep_rec) return record_s2d(w_oihw, dst, O, C, KH, KW, stride, ld, 0);
  pack_s2d_kernel<<<grid_for((long long)O * C * KH * KW), 256, 0, s>>>(w_oihw, dst, O, C, KH, KW, stride, ld, 0);
  DDRL_LAUNCHED("pack_kernel");
  return DDRL_OK;
}
int unpack_grad_s2d(const float* src, float* dst_oihw, int O, int C, int KH, int KW, int stride, int ld, cudaStream_t s) {
  if (g_prep_rec) return record_s2d(src, dst_oihw, O, C, KH, KW, stride, ld, 1);
  pack_s2d_kernel<<<grid_for((long long)O * C * KH * KW), 256, 0, s>>>(src, dst_oihw, O, C, KH, KW, stride, ld, 1);
  DDRL_LAUNCHED("unpack_kernel");
  return DDRL_OK;
}
int copy2d(const float* src, int ld_s, float* dst, int ld_d, long long rows, int colsN, cudaStream_t s) {
  if (rows * colsN == 0) return DDRL_OK;
  copy2d_kernel<<<grid_for(rows * colsN), 256, 0, s>>>(src, ld_s, dst, ld_d, rows, colsN);
  DDRL_LAUNCHED("copy2d_kernel");
  return DDRL_OK;
}
static const bool g_skinny_reg = [] { const char* e = getenv("DDRL_SKINNY_GENERAL"); return !(e && e[0] == '1'); }();
int skinny_fwd(const float* x, int ldx, const float* W, const float* bias, int B, int N, int K, float* y, int ldy,
               cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  if (g_skinny_reg && K <= 512 && N <= 8) {
    if (N == 1) skinny_fwd_reg_kernel<1><<<ceil_div(B, 8), 256, 0, s>>>(x, ldx, W, bias, B, N, K, y, ldy);
    else skinny_fwd_reg_kernel<8><<<ceil_div(B, 8), 256, 0, s>>>(x, ldx, W, bias, B, N, K, y, ldy);
    DDRL_LAUNCHED("skinny_fwd_kernel");
    return DDRL_OK;
  }
  skinny_fwd_kernel<<<ceil_div(B, 8), 256, 0, s>>>(x, ldx, W, bias, B, N, K, y, ldy);
  DDRL_LAUNCHED("skinny_fwd_kernel");
  return DDRL_OK;
}
int skinny_dgrad(const float* dy, int ldy, const float* W, int B, int N, int K, float* dx, int ldx, int accumulate,
                 cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  const bool al = ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 && ldx % 4 == 0;
  if (g_skinny_reg && al && K % 4 == 0 && K <= 1024 && N <= 8) {
    const int rpb = 256 / (K / 4);
    const int grid = (int)std::min<long long>(ceil_div(B, rpb), 16LL * kNumSMs);
    if (N == 1) skinny_dgrad_reg_kernel<1><<<grid, 256, 0, s>>>(dy, ldy, W, B, N, K, dx, ldx, accumulate);
    else skinny_dgrad_reg_kernel<8><<<grid, 256, 0, s>>>(dy, ldy, W, B, N, K, dx, ldx, accumulate);
    DDRL_LAUNCHED("skinny_dgrad_kernel");
    return DDRL_OK;
  }
  skinny_dgrad_kernel<<<grid_for((long long)B * K), 256, 0, s>>>(dy, ldy, W, B, N, K, dx, ldx, accumulate);
  DDRL_LAUNCHED("skinny_dgrad_kernel");
  return DDRL_OK;
}
int skinny_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int N, int K, float* dW, float* db,
                 cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  const int kb = ceil_div(K, 256);
  if (g_skinny_reg && N <= 8) {
    const DetSeq det = det_seq(kb);
    int chunks = std::max(1, std::min(ceil_div(B, 64), det.ctr ? kDetMaxParts : ceil_div(4 * kNumSMs, kb)));
    const int rpb = ceil_div(ceil_div(B, chunks), 64) * 64;
    chunks = ceil_div(B, rpb);
    if (N == 1) skinny_wgrad_reg_kernel<1><<<dim3(kb, chunks), 256, 0, s>>>(dy, ldy, x, ldx, B, N, K, dW, db, rpb, det);
    else skinny_wgrad_reg_kernel<8><<<dim3(kb, chunks), 256, 0, s>>>(dy, ldy, x, ldx, B, N, K, dW, db, rpb, det);
    DDRL_LAUNCHED("skinny_wgrad_kernel");
    return DDRL_OK;
  }
  const DetSeq det = det_seq(kb * N);
  int chunks = std::max(1, std::min(ceil_div(B, 64), det.ctr ? kDetMaxParts : ceil_div(4 * kNumSMs, kb * N)));
  const int rpb = ceil_div(B, chunks);
  chunks = ceil_div(B, rpb);
  skinny_wgrad_kernel<<<dim3(kb, chunks, N), 256, 0, s>>>(dy, ldy, x, ldx, B, N, K, dW, db, rpb, det);
  DDRL_LAUNCHED("skinny_wgrad_kernel");
  return DDRL_OK;
}

}  // namespace ddrl
