// Convolution layers on the implicit-GEMM tcgen05 path (gemm_tc.cu: conv_tc_fwd / conv_tc_wgrad).
//
// The reference runs nn.Conv2d / nn.Conv1d through autograd (nn/atari_encoder.py:16-28, nn/nav_encoder.py:17-31,
// 85-93): forward cross-correlation, and in backward the data gradient (a transposed convolution) and the weight
// gradient.  Here all three are GEMMs whose activation operand is fetched tap by tap with 4-D TMA boxes straight
// from the NHWC activation tensor:
//   forward   y[pix, o]      = sum_{kh,kw,c} x[pix*s + (kh,kw) - p, c] * Wp[o, (kh,kw,c)]
//   dgrad     dx[pix', c]    = sum over the taps that reach pix' of dy[...] * W   -- split by the parity class of
//             (pix' + p) mod s: each class is a stride-1 convolution over dy with ceil((K - r)/s) taps per axis
//             and its own re-packed (flipped, transposed) weights; the classes write interleaved output pixels
//   wgrad     dWp[o, (kh,kw,c)] = sum_pix dy[pix, o] * x[pix*s + (kh,kw) - p, c]
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "layer_ops.h"
#include "prep_kernels.cuh"

namespace ddrl {

// body: prep_kernels.cuh
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin,
                                                         int KH, int KW, int s, int ry, int rx, int nty, int ntx) {
  pack_dgrad_body(w, wd, Cout, Cin, KH, KW, s, ry, rx, nty, ntx, blockIdx.x, gridDim.x);
}

// 1-D convolutions (H == 1) are run with the pixel axis on the box's row dimension
static void oriented(const ConvGeom& g, int& H, int& W, int& KH, int& KW, int& Ho, int& Wo) {
  if (g.H == 1 && g.KH == 1) { H = g.W; W = 1; KH = g.KW; KW = 1; Ho = g.Wo; Wo = 1; }
  else { H = g.H; W = g.W; KH = g.KH; KW = g.KW; Ho = g.Ho; Wo = g.Wo; }
}

ConvOp conv_op_fwd(const ConvGeom& g, const float* x, int Ctot, int c_off, int B) {
  ConvOp o;
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  o.a = x; o.Hin = H; o.Win = W; o.Ctot = Ctot; o.c_off = c_off; o.Cin = g.C;
  o.KH = KH; o.KW = KW; o.sy = g.stride; o.sx = W == 1 ? 1 : g.stride; o.py = g.pad; o.px = W == 1 ? 0 : g.pad;
  o.Yn = Ho; o.Xn = Wo; o.Bn = B;
  return o;
}

int conv_dgrad_plan(const ConvGeom& g, int Cout, std::vector<DgradClass>& out) {
  out.clear();
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const int s = g.stride, py = g.pad, px = W == 1 ? 0 : g.pad;
  const int sxn = W == 1 ? 1 : s;
  for (int ry = 0; ry < s; ++ry)
    for (int rx = 0; rx < sxn; ++rx) {
      DgradClass c;
      c.ry = ry; c.rx = rx;
      c.nty = (KH - ry + s - 1) / s;
      c.ntx = (KW - rx + sxn - 1) / sxn;
      if (c.nty < 1 || c.ntx < 1) return DDRL_E_UNSUPPORTED;
      c.iy0 = ((ry - py) % s + s) % s;
      c.ix0 = ((rx - px) % sxn + sxn) % sxn;
      if (c.iy0 >= H || c.ix0 >= W) continue;
      c.Yn = (H - c.iy0 + s - 1) / s;
      c.Xn = (W - c.ix0 + sxn - 1) / sxn;
      c.pady = c.nty - 1 - (c.iy0 + py - ry) / s;
      c.padx = c.ntx - 1 - (c.ix0 + px - rx) / sxn;
      c.K = c.nty * c.ntx * Cout;
      c.wd = c.wd_hi = c.wd_lo = nullptr;
      out.push_back(c);
    }
  return DDRL_OK;
}

int pack_dgrad(const float* w_oihw, const ConvGeom& g, int Cout, const DgradClass& c, cudaStream_t s) {
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const long long total = (long long)g.C * c.K;
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_PACK_DGRAD; j.a = w_oihw; j.b = c.wd; j.total = total; j.vblocks = prep_blocks(total);
    const int v[9] = {Cout, g.C, KH, KW, g.stride, c.ry, c.rx, c.nty, c.ntx};
    for (int k = 0; k < 9; ++k) j.i[k] = v[k];
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
  // the reference weight is [Cout, Cin, g.KH, g.KW]; in the transposed (1-D) orientation KH/KW swap with it
  pack_dgrad_kernel<<<blocks, 256, 0, s>>>(w_oihw, c.wd, Cout, g.C, KH, KW, g.stride, c.ry, c.rx, c.nty, c.ntx);
  DDRL_LAUNCHED("pack_dgrad_kernel");
  return DDRL_OK;
}

// dx (dense NHWC [B, H, W, C]) = sum over classes; every input pixel belongs to exactly one class
int conv_dgrad_tc(const ConvGeom& g, int Cout, const std::vector<DgradClass>& cls, const float* dy, int dy_ctot, int dy_coff,
                  float* dx, int act, const float* mask, int B, cudaStream_t s, bool tmem_engine, const Tc3Ctx* t3) {
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const int sy = g.stride, sx = W == 1 ? 1 : g.stride;
  for (const DgradClass& c : cls) {
    ConvOp o;
    o.a = dy; o.Hin = Ho; o.Win = Wo; o.Ctot = dy_ctot; o.c_off = dy_coff; o.Cin = Cout;
    o.KH = c.nty; o.KW = c.ntx; o.sy = 1; o.sx = 1; o.py = c.pady; o.px = c.padx;
    o.Yn = c.Yn; o.Xn = c.Xn; o.Bn = B;
    const long long off = ((long long)c.iy0 * W + c.ix0) * g.C;
    int r;
    if (t3)
      r = tc3_conv_fwd(o, c.w16.hi, c.w16.lo, c.w16.ld, g.C, t3->amax_a, c.w16.amax, nullptr, act, mask ? mask + off : nullptr,
                       dx + off, (long long)H * W * g.C, (long long)sy * W * g.C, (long long)sx * g.C, t3->amax_out, s);
    else if (tmem_engine)
      r = tc2_conv_fwd(o, c.wd_hi, c.wd_lo, c.K, g.C, nullptr, act, mask ? mask + off : nullptr, dx + off,
                       (long long)H * W * g.C, (long long)sy * W * g.C, (long long)sx * g.C, s);
    else
      r = conv_tc_fwd(o, c.wd, c.K, g.C, nullptr, act, mask ? mask + off : nullptr, dx + off, (long long)H * W * g.C,
                      (long long)sy * W * g.C, (long long)sx * g.C, s);
    if (r != DDRL_OK) return r;
  }
  return DDRL_OK;
}

// ---- fused parity classes (tc2 engine) ----------------------------------------------------------------------------
// body: prep_kernels.cuh
__global__ void __launch_bounds__(256) pack_dgrad_fused_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout,
                                                               int Cin, int KH, int KW, int s, int sx, int ncx, int nty, int ntx,
                                                               int pady, int padx, int q0y0, int q0y1, int q0x0, int q0x1,
                                                               long long total) {
  pack_dgrad_fused_body(w, wd, Cout, Cin, KH, KW, s, sx, ncx, nty, ntx, pady, padx, q0y0, q0y1, q0x0, q0x1, total, blockIdx.x,
                        gridDim.x);
}

int conv_dgrad_fused_plan(const ConvGeom& g, int Cout, DgradFused& f) {
  memset(&f, 0, sizeof(f));
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const int s = g.stride, sx = W == 1 ? 1 : s;
  const int py = g.pad, px = W == 1 ? 0 : g.pad;
  if (s < 2 || s > 2) return DDRL_E_UNSUPPORTED;                // stride 1 has a single class: the plain path
  if (g.order != 0 || g.C % 4 != 0 || Cout % 32 != 0) return DDRL_E_UNSUPPORTED;
  f.s = s; f.ncy = s; f.ncx = sx;
  int pady = 0, padx = 0, topy = 0, topx = 0;
  for (int r = 0; r < s; ++r) {
    const int nt = (KH - r + s - 1) / s;
    if (nt < 1) return DDRL_E_UNSUPPORTED;
    f.iy0[r] = ((r - py) % s + s) % s;
    f.q0y[r] = (f.iy0[r] + py - r) / s;
    pady = std::max(pady, nt - 1 - f.q0y[r]);
    topy = std::max(topy, f.q0y[r]);
  }
  for (int r = 0; r < sx; ++r) {
    const int nt = (KW - r + sx - 1) / sx;
    if (nt < 1) return DDRL_E_UNSUPPORTED;
    f.ix0[r] = ((r - px) % sx + sx) % sx;
    f.q0x[r] = (f.ix0[r] + px - r) / sx;
    padx = std::max(padx, nt - 1 - f.q0x[r]);
    topx = std::max(topx, f.q0x[r]);
  }
  f.pady = pady; f.padx = padx;
  f.nty = topy + pady + 1; f.ntx = topx + padx + 1;
  f.Jy = (H + s - 1) / s; f.Jx = (W + sx - 1) / sx;
  f.K = f.nty * f.ntx * Cout;
  f.N = f.ncy * f.ncx * g.C;
  if (f.N > 256) return DDRL_E_UNSUPPORTED;
  ConvOp o;
  o.a = reinterpret_cast<const float*>(uintptr_t(256)); o.Hin = Ho; o.Win = Wo; o.Ctot = Cout; o.c_off = 0; o.Cin = Cout;
  o.KH = f.nty; o.KW = f.ntx; o.sy = 1; o.sx = 1; o.py = f.pady; o.px = f.padx; o.Yn = f.Jy; o.Xn = f.Jx; o.Bn = 1;
  if (!conv_tc_supported(o, false)) return DDRL_E_UNSUPPORTED;
  f.on = true;
  return DDRL_OK;
}

int pack_dgrad_fused(const float* w_oihw, const ConvGeom& g, int Cout, const DgradFused& f, cudaStream_t s) {
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const long long total = (long long)f.N * f.K;
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_PACK_DGRAD_FUSED; j.a = w_oihw; j.b = f.wd; j.total = total; j.vblocks = prep_blocks(total);
    const int v[15] = {Cout, g.C, KH, KW, f.s, W == 1 ? 1 : f.s, f.ncx, f.nty, f.ntx, f.pady, f.padx, f.q0y[0], f.q0y[1], f.q0x[0], f.q0x[1]};
    for (int k = 0; k < 15; ++k) j.i[k] = v[k];
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
  pack_dgrad_fused_kernel<<<blocks, 256, 0, s>>>(w_oihw, f.wd, Cout, g.C, KH, KW, f.s, W == 1 ? 1 : f.s, f.ncx, f.nty, f.ntx,
                                                 f.pady, f.padx, f.q0y[0], f.q0y[1], f.q0x[0], f.q0x[1], total);
  DDRL_LAUNCHED("pack_dgrad_fused_kernel");
  return DDRL_OK;
}

// out_ctot / out_coff: dx (and the mask, addressed like dx) are channels [out_coff, out_coff + C) of an out_ctot-wide
// NHWC tensor (0 = dense [B, H, W, C])
int conv_dgrad_fused_tc2(const ConvGeom& g, int Cout, const DgradFused& f, const float* dy, float* dx, int act,
                         const float* mask, int B, cudaStream_t s, int out_ctot, int out_coff, const Tc3Ctx* t3) {
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  const int sx = W == 1 ? 1 : f.s;
  const long long ct = out_ctot > 0 ? out_ctot : g.C;
  if (out_ctot > 0) { dx += out_coff; if (mask) mask += out_coff; }
  ConvOp o;
  o.a = dy; o.Hin = Ho; o.Win = Wo; o.Ctot = Cout; o.c_off = 0; o.Cin = Cout;
  o.KH = f.nty; o.KW = f.ntx; o.sy = 1; o.sx = 1; o.py = f.pady; o.px = f.padx;
  o.Yn = f.Jy; o.Xn = f.Jx; o.Bn = B;
  TcTap cls;
  memset(&cls, 0, sizeof(cls));
  cls.ncls = f.ncy * f.ncx; cls.cls_cols = g.C; cls.out_s = f.s; cls.out_H = H; cls.out_W = W;
  cls.work_scale = (float)(KH * KW) / (float)(f.ncy * f.ncx * f.nty * f.ntx);
  for (int cy = 0; cy < f.ncy; ++cy)
    for (int cx = 0; cx < f.ncx; ++cx) {
      const int q = cy * f.ncx + cx;
      cls.cls_iy[q] = f.iy0[cy]; cls.cls_ix[q] = f.ix0[cx];
      cls.cls_off[q] = ((long long)f.iy0[cy] * W + f.ix0[cx]) * ct;
    }
  // tile pixel (jy, jx) -> base input pixel (s*jy, sx*jx); out_s scales both axes, so a 1-D layer (W == 1) keeps x = 0
  if (t3)
    return tc3_conv_fwd(o, f.w16.hi, f.w16.lo, f.w16.ld, f.N, t3->amax_a, f.w16.amax, nullptr, act, mask, dx, (long long)H * W * ct,
                        (long long)f.s * W * ct, (long long)sx * ct, t3->amax_out, s, &cls);
  return tc2_conv_fwd(o, f.wd_hi, f.wd_lo, f.K, f.N, nullptr, act, mask, dx, (long long)H * W * ct, (long long)f.s * W * ct,
                      (long long)sx * ct, s, &cls);
}

bool conv_dgrad_supported(const ConvGeom& g, int Cout, const std::vector<DgradClass>& cls, const float* dy, int dy_ctot,
                          int dy_coff, int B) {
  int H, W, KH, KW, Ho, Wo;
  oriented(g, H, W, KH, KW, Ho, Wo);
  if (g.order != 0 || g.C % 4 != 0) return false;
  for (const DgradClass& c : cls) {
    ConvOp o;
    o.a = dy; o.Hin = Ho; o.Win = Wo; o.Ctot = dy_ctot; o.c_off = dy_coff; o.Cin = Cout;
    o.KH = c.nty; o.KW = c.ntx; o.sy = 1; o.sx = 1; o.py = c.pady; o.px = c.padx;
    o.Yn = c.Yn; o.Xn = c.Xn; o.Bn = B;
    if (!conv_tc_supported(o, false)) return false;
  }
  return true;
}

}  // namespace ddrl

// =========================================================================================
// C ABI: stand-alone convolution entry (parity tests of the implicit-GEMM path)
// =========================================================================================
using namespace ddrl;

extern "C" int ddrl_conv_nhwc_f32(int mode, int op, const ddrl_conv_desc* d, const float* x, const float* w, const float* bias,
                                  const float* dy, int act, const float* mask, float* out, void* stream) {
  if (!d || !out || !w || op < 0 || op > 2) return DDRL_E_ARG;
  if (mode != DDRL_GEMM_TC_3XTF32 && mode != DDRL_GEMM_TC2_TMEM && mode != DDRL_GEMM_TC3_F16) return DDRL_E_ARG;
  const bool v3 = mode == DDRL_GEMM_TC3_F16;
  const bool v2 = mode == DDRL_GEMM_TC2_TMEM;
  if (d->B < 1 || d->H < 1 || d->W < 1 || d->Cin < 1 || d->Cout < 1 || d->KH < 1 || d->KW < 1 || d->stride < 1 || d->pad < 0)
    return DDRL_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  ConvGeom g;
  g.H = d->H; g.W = d->W; g.C = d->Cin;
  g.sc = 1; g.sw = d->Cin; g.sh = (long long)d->W * d->Cin; g.sb = (long long)d->H * d->W * d->Cin; g.order = 0;
  g.KH = d->KH; g.KW = d->KW; g.stride = d->stride; g.pad = d->pad;
  // H == 1 && KH == 1 is a Conv1d: stride / padding act on the W axis only
  g.Ho = (d->H == 1 && d->KH == 1) ? 1 : (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  g.Wo = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  g.K = d->Cin * d->KH * d->KW; g.ldc = g.K;
  if (g.Ho < 1 || g.Wo < 1) return DDRL_E_ARG;
  const int taps = d->KH * d->KW;
  int rc = DDRL_OK;
  float* tmp = nullptr;
  if (v3 && op != 2) {
    // tc3: pack the weights (forward layout, or the fused / per-class data-gradient layouts), split them to scaled fp16,
    // take amax of the activation operand, run
    char* raw = nullptr;
    auto finish = [&](int r) { cudaStreamSynchronize(s); if (raw) cudaFree(raw); return r; };
    auto split_into = [&](const float* wp, int rows, int K, W16& w16, char* h16, float* slot) {
      const int ld16 = (K + 7) & ~7;
      w16.hi = h16; w16.lo = h16 + (size_t)rows * ld16 * 2; w16.ld = ld16; w16.amax = slot;
      int r = amax_f32(wp, rows, K, K, slot, true, s);
      if (r == DDRL_OK) r = split_f16(wp, rows, K, K, slot, const_cast<void*>(w16.hi), const_cast<void*>(w16.lo), ld16, nullptr, nullptr, 0, s);
      return r;
    };
    if (op == 0) {
      if (!x) return DDRL_E_ARG;
      const size_t nw = (size_t)d->Cout * g.K, n16 = (size_t)d->Cout * ((g.K + 7) & ~7);
      DDRL_CUDA(cudaMalloc(&raw, 256 + 4 * nw + 4 * n16 + 256));
      float* slots = reinterpret_cast<float*>(raw);
      float* wp = reinterpret_cast<float*>(raw + 256);
      DDRL_CUDA(cudaMemsetAsync(raw, 0, 256, s));
      rc = pack_weight(w, wp, d->Cout, taps, d->Cin, g.K, s);
      ConvOp o = conv_op_fwd(g, x, d->Cin, 0, d->B);
      if (rc == DDRL_OK && !conv_tc_supported(o, false)) rc = DDRL_E_UNSUPPORTED;
      W16 w16;
      if (rc == DDRL_OK) rc = split_into(wp, d->Cout, g.K, w16, raw + 256 + 4 * nw, slots + 1);
      if (rc == DDRL_OK) rc = amax_f32(x, (long long)d->B * d->H * d->W, d->Cin, d->Cin, slots, false, s);
      if (rc == DDRL_OK)
        rc = tc3_conv_fwd(o, w16.hi, w16.lo, w16.ld, d->Cout, slots, slots + 1, bias, act, mask, out, (long long)o.Yn * o.Xn * d->Cout,
                          (long long)o.Xn * d->Cout, d->Cout, slots + 2, s);
      return finish(rc);
    }
    if (!dy) return DDRL_E_ARG;
    DgradFused fz;
    std::vector<DgradClass> cls;
    const bool fused = conv_dgrad_fused_plan(g, d->Cout, fz) == DDRL_OK;
    size_t tot = 0, tot16 = 0;
    if (fused) { tot = (size_t)fz.N * fz.K; tot16 = (size_t)fz.N * ((fz.K + 7) & ~7); }
    else {
      rc = conv_dgrad_plan(g, d->Cout, cls);
      if (rc != DDRL_OK) return rc;
      for (auto& c : cls) { tot += (size_t)g.C * c.K; tot16 += (size_t)g.C * ((c.K + 7) & ~7); }
    }
    DDRL_CUDA(cudaMalloc(&raw, 1024 + 4 * tot + 4 * tot16 + 256));
    DDRL_CUDA(cudaMemsetAsync(raw, 0, 1024, s));
    float* slots = reinterpret_cast<float*>(raw);
    float* wp = reinterpret_cast<float*>(raw + 1024);
    char* h16 = raw + 1024 + 4 * tot;
    rc = amax_f32(dy, (long long)d->B * g.Ho * g.Wo, d->Cout, d->Cout, slots, false, s);
    Tc3Ctx ctx{slots, slots + 1};
    if (fused) {
      fz.wd = wp;
      if (rc == DDRL_OK) rc = pack_dgrad_fused(w, g, d->Cout, fz, s);
      if (rc == DDRL_OK) rc = split_into(wp, fz.N, fz.K, fz.w16, h16, slots + 2);
      if (rc == DDRL_OK) rc = conv_dgrad_fused_tc2(g, d->Cout, fz, dy, out, act, mask, d->B, s, 0, 0, &ctx);
      return finish(rc);
    }
    size_t off = 0, off16 = 0;
    int si = 2;
    for (auto& c : cls) {
      c.wd = wp + off;
      if (rc == DDRL_OK) rc = pack_dgrad(w, g, d->Cout, c, s);
      if (rc == DDRL_OK) rc = split_into(c.wd, g.C, c.K, c.w16, h16 + 4 * off16, slots + si++);
      off += (size_t)g.C * c.K; off16 += (size_t)g.C * ((c.K + 7) & ~7);
    }
    if (rc == DDRL_OK && !conv_dgrad_supported(g, d->Cout, cls, dy, d->Cout, 0, d->B)) rc = DDRL_E_UNSUPPORTED;
    if (rc == DDRL_OK) rc = conv_dgrad_tc(g, d->Cout, cls, dy, d->Cout, 0, out, act, mask, d->B, s, true, &ctx);
    return finish(rc);
  }
  if (op == 0) {
    if (!x) return DDRL_E_ARG;
    const size_t nw = ((size_t)d->Cout * g.K + 3) & ~size_t(3);
    DDRL_CUDA(cudaMalloc(&tmp, sizeof(float) * 3 * nw));
    DDRL_CUDA(cudaMemsetAsync(tmp, 0, sizeof(float) * 3 * nw, s));
    rc = pack_weight(w, tmp, d->Cout, taps, d->Cin, g.K, s);
    ConvOp o = conv_op_fwd(g, x, d->Cin, 0, d->B);
    if (rc == DDRL_OK && !conv_tc_supported(o, false)) rc = DDRL_E_UNSUPPORTED;
    if (rc == DDRL_OK && v2) rc = split_hi_lo(tmp, tmp + nw, tmp + 2 * nw, (long long)nw, s);
    if (rc == DDRL_OK) {
      if (v2)
        rc = tc2_conv_fwd(o, tmp + nw, tmp + 2 * nw, g.K, d->Cout, bias, act, mask, out, (long long)o.Yn * o.Xn * d->Cout,
                          (long long)o.Xn * d->Cout, d->Cout, s);
      else
        rc = conv_tc_fwd(o, tmp, g.K, d->Cout, bias, act, mask, out, (long long)o.Yn * o.Xn * d->Cout,
                         (long long)o.Xn * d->Cout, d->Cout, s);
    }
  } else if (op == 1) {
    if (!dy) return DDRL_E_ARG;
    DgradFused fz;
    if (v2 && conv_dgrad_fused_plan(g, d->Cout, fz) == DDRL_OK) {
      const size_t tot = (size_t)fz.N * fz.K;
      DDRL_CUDA(cudaMalloc(&tmp, sizeof(float) * 3 * tot));
      fz.wd = tmp; fz.wd_hi = tmp + tot; fz.wd_lo = tmp + 2 * tot;
      rc = pack_dgrad_fused(w, g, d->Cout, fz, s);
      if (rc == DDRL_OK) rc = split_hi_lo(tmp, tmp + tot, tmp + 2 * tot, (long long)tot, s);
      if (rc == DDRL_OK) rc = conv_dgrad_fused_tc2(g, d->Cout, fz, dy, out, act, mask, d->B, s);
      cudaStreamSynchronize(s);
      cudaFree(tmp);
      return rc;
    }
    std::vector<DgradClass> cls;
    rc = conv_dgrad_plan(g, d->Cout, cls);
    size_t tot = 0;
    for (auto& c : cls) tot += (size_t)g.C * c.K;                 // Cout % 32 == 0 on this path: multiples of 4
    if (rc == DDRL_OK) DDRL_CUDA(cudaMalloc(&tmp, sizeof(float) * 3 * tot));
    size_t off = 0;
    for (auto& c : cls) { c.wd = tmp + off; c.wd_hi = c.wd + tot; c.wd_lo = c.wd + 2 * tot; off += (size_t)g.C * c.K; }
    for (auto& c : cls) if (rc == DDRL_OK) rc = pack_dgrad(w, g, d->Cout, c, s);
    if (rc == DDRL_OK && !conv_dgrad_supported(g, d->Cout, cls, dy, d->Cout, 0, d->B)) rc = DDRL_E_UNSUPPORTED;
    if (rc == DDRL_OK && v2) rc = split_hi_lo(tmp, tmp + tot, tmp + 2 * tot, (long long)tot, s);
    if (rc == DDRL_OK) rc = conv_dgrad_tc(g, d->Cout, cls, dy, d->Cout, 0, out, act, mask, d->B, s, v2);
  } else {
    if (!x || !dy) return DDRL_E_ARG;
    DDRL_CUDA(cudaMalloc(&tmp, sizeof(float) * (size_t)d->Cout * g.K));
    DDRL_CUDA(cudaMemsetAsync(tmp, 0, sizeof(float) * (size_t)d->Cout * g.K, s));
    ConvOp o = conv_op_fwd(g, x, d->Cin, 0, d->B);
    if (v3) {
      float* slots = nullptr;
      if (!tc3_conv_wgrad_supported(o)) rc = DDRL_E_UNSUPPORTED;
      if (rc == DDRL_OK && cudaMalloc(&slots, 256) != cudaSuccess) rc = DDRL_E_CUDA;
      if (rc == DDRL_OK && cudaMemsetAsync(slots, 0, 256, s) != cudaSuccess) rc = DDRL_E_CUDA;
      if (rc == DDRL_OK) rc = amax_f32(x, (long long)d->B * d->H * d->W, d->Cin, d->Cin, slots, false, s);
      if (rc == DDRL_OK) rc = amax_f32(dy, (long long)d->B * g.Ho * g.Wo, d->Cout, d->Cout, slots + 1, false, s);
      if (rc == DDRL_OK) rc = tc3_conv_wgrad(o, dy, d->Cout, d->Cout, slots, slots + 1, tmp, g.K, s);
      if (rc == DDRL_OK) rc = unpack_grad(tmp, out, d->Cout, taps, d->Cin, g.K, s);
      cudaStreamSynchronize(s);
      if (slots) cudaFree(slots);
      if (tmp) cudaFree(tmp);
      return rc;
    }
    if (!conv_tc_supported(o, true)) rc = DDRL_E_UNSUPPORTED;
    if (rc == DDRL_OK) rc = v2 ? tc2_conv_wgrad(o, dy, d->Cout, d->Cout, tmp, g.K, s) : conv_tc_wgrad(o, dy, d->Cout, d->Cout, tmp, g.K, s);
    if (rc == DDRL_OK) rc = unpack_grad(tmp, out, d->Cout, taps, d->Cin, g.K, s);
  }
  cudaStreamSynchronize(s);
  if (tmp) cudaFree(tmp);
  return rc;
}
