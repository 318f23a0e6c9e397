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
    if (idx) idx[i] = (uint8_t)bi;
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
      if (idx[o] == me && a[i] > 0.f) v = dout[o];
    }
    da[i] = v;
  }
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
                                                     float* __restrict__ db, long long rows_per_block) {
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
    atomicAdd(db + n, s);
  }
}

// ---------------------------------------------------------------- weight packing
// packed[o*ld + i*J + j] = src[o*I*J + j*I + i]   (inner [J][I] -> [I][J] transpose; I = 1: row-stride change)
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int J,
                                                   int ld) {
  const long long total = (long long)O * I * J;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % J);
    const long long r = t / J;
    const int i = (int)(r % I);
    const long long o = r / I;
    dst[o * ld + (long long)i * J + j] = src[o * I * J + (long long)j * I + i];
  }
}
// grad[o*I*J + j*I + i] (+)= packed_grad[o*ld + i*J + j]
__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I,
                                                     int J, int ld) {
  const long long total = (long long)O * I * J;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % I);
    const long long r = t / I;
    const int j = (int)(r % J);
    const long long o = r / J;
    dst[t] = src[o * ld + (long long)i * J + j];
  }
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
                                                           float* __restrict__ db, int rows_per_block) {
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
    atomicAdd(dW + (size_t)n * K + k, acc);
    if (db && k == 0) atomicAdd(db + n, accb);
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
int pool_fwd(const float* a, float* out, uint8_t* idx, int B, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  if (total == 0) return DDRL_OK;
  pool_fwd_kernel<<<grid_for(total), 256, 0, s>>>(a, out, idx, H, W, C, total);
  DDRL_LAUNCHED("pool_fwd_kernel");
  return DDRL_OK;
}
int pool_bwd(const float* dout, const uint8_t* idx, const float* a, float* da, int B, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return DDRL_OK;
  pool_bwd_kernel<<<grid_for(total), 256, 0, s>>>(dout, idx, a, da, H, W, C, total);
  DDRL_LAUNCHED("pool_bwd_kernel");
  return DDRL_OK;
}
int act_bwd(float* dy, int ld_dy, const float* y, int ld_y, long long rows, int colsN, int act, cudaStream_t s) {
  if (act == 0 || rows * colsN == 0) return DDRL_OK;
  act_bwd_kernel<<<grid_for(rows * colsN), 256, 0, s>>>(dy, ld_dy, y, ld_y, (int)rows, colsN, act);
  DDRL_LAUNCHED("act_bwd_kernel");
  return DDRL_OK;
}
int colsum_add(const float* dy, int ld, long long rows, int N, float* db, cudaStream_t s) {
  if (rows == 0 || N == 0) return DDRL_OK;
  const int nb = ceil_div(N, 32);
  long long chunks = std::max<long long>(1, std::min<long long>((rows + 255) / 256, (4LL * kNumSMs + nb - 1) / nb));
  const long long rpb = (rows + chunks - 1) / chunks;
  chunks = (rows + rpb - 1) / rpb;
  colsum_kernel<<<dim3(nb, (unsigned)chunks), 256, 0, s>>>(dy, ld, rows, N, db, rpb);
  DDRL_LAUNCHED("colsum_kernel");
  return DDRL_OK;
}
int pack_weight(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s) {
  pack_kernel<<<grid_for((long long)O * I * J), 256, 0, s>>>(src, dst, O, I, J, ld);
  DDRL_LAUNCHED("pack_kernel");
  return DDRL_OK;
}
int unpack_grad(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s) {
  unpack_kernel<<<grid_for((long long)O * I * J), 256, 0, s>>>(src, dst, O, I, J, ld);
  DDRL_LAUNCHED("unpack_kernel");
  return DDRL_OK;
}
int copy2d(const float* src, int ld_s, float* dst, int ld_d, long long rows, int colsN, cudaStream_t s) {
  if (rows * colsN == 0) return DDRL_OK;
  copy2d_kernel<<<grid_for(rows * colsN), 256, 0, s>>>(src, ld_s, dst, ld_d, rows, colsN);
  DDRL_LAUNCHED("copy2d_kernel");
  return DDRL_OK;
}
int skinny_fwd(const float* x, int ldx, const float* W, const float* bias, int B, int N, int K, float* y, int ldy,
               cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  skinny_fwd_kernel<<<ceil_div(B, 8), 256, 0, s>>>(x, ldx, W, bias, B, N, K, y, ldy);
  DDRL_LAUNCHED("skinny_fwd_kernel");
  return DDRL_OK;
}
int skinny_dgrad(const float* dy, int ldy, const float* W, int B, int N, int K, float* dx, int ldx, int accumulate,
                 cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  skinny_dgrad_kernel<<<grid_for((long long)B * K), 256, 0, s>>>(dy, ldy, W, B, N, K, dx, ldx, accumulate);
  DDRL_LAUNCHED("skinny_dgrad_kernel");
  return DDRL_OK;
}
int skinny_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int N, int K, float* dW, float* db,
                 cudaStream_t s) {
  if (B == 0) return DDRL_OK;
  const int kb = ceil_div(K, 256);
  int chunks = std::max(1, std::min(ceil_div(B, 64), ceil_div(4 * kNumSMs, kb * N)));
  const int rpb = ceil_div(B, chunks);
  chunks = ceil_div(B, rpb);
  skinny_wgrad_kernel<<<dim3(kb, chunks, N), 256, 0, s>>>(dy, ldy, x, ldx, B, N, K, dW, db, rpb);
  DDRL_LAUNCHED("skinny_wgrad_kernel");
  return DDRL_OK;
}

}  // namespace ddrl
