// Internal (non-ABI) declarations shared by the engine translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/ddrl_b200.h"

namespace ddrl {

struct ConvGeom {
  int H, W, C;                   // input extent
  long long sb, sh, sw, sc;      // input element strides (NCHW or NHWC)
  int KH, KW, stride, pad;
  int Ho, Wo;
  int K;                         // C*KH*KW
  int ldc;                       // row stride of the im2col matrix (K rounded up to 4)
  int order;                     // 0: k=(kh,kw,c)   1: k=(c,kh,kw)
};

// Implicit-GEMM operand description for the tcgen05 engine (gemm_tc.cu).  mode 0 = plain 2-D operands.
// mode 1: the activation operand is fetched as 4-D TMA boxes (32 channels x Xn pixels x ny rows x nb images) of an
// NHWC tensor, one box per (kernel tap, 32-channel chunk): no im2col matrix exists in memory.
struct TcTap {
  int mode;
  int Xn, Yn, Bn;                // output pixel grid: Xn x Yn per image, Bn images
  int ny, nb, tpi;               // box = nb images x ny rows x Xn pixels; boxes per image
  int rows;                      // Xn*ny*nb   (forward/dgrad: <= 128 = the M tile; wgrad: <= 32 = one K block)
  // two-phase tiling (tc2 forward/dgrad): tiles [0, tiles1) are phase-1 boxes (rows [0, y2) of each image in blocks of ny,
  // nb images per box); tiles >= tiles1 cover the remaining rows [y2, Yn) with nb2 images per box (their own tensor map).
  // 9x9 images: 7 rows x 2 images + 2 rows x 7 images = 126-row tiles instead of one 81-row image per 128-row tile.
  int tiles1, y2, ny2, nb2, rows2;
  // weight gradient, two pixel-box classes per image (K blocks of exactly 32 pixels): columns [0, wxw0) in boxes of
  // wxw0 x wyh0 pixels (wnb0 per image), columns [wxw0, Xn) in boxes of wxw1 x wyh1 (tpi - wnb0 per image).  A 20-pixel-wide
  // map: 16x2 + 4x8 boxes = 13 full K blocks per image instead of 20 blocks of one 20-pixel row.
  int w2on, wxw0, wyh0, wnb0, wxw1, wyh1;
  int kpad;                      // rows rounded up to 8 (wgrad: MMA K steps per block)
  int KW, cpb;                   // tap index = kh*KW + kw; 32-channel chunks per tap
  int nslices;                   // taps * cpb
  int sx, sy, px, py;            // input coordinate = out*s + k - p
  int c_off;                     // first channel of the operand inside the tensor's channel axis
  long long osb, osy, osx;       // output element strides per (image, row, pixel)   (forward/dgrad)
  // fused stride-parity data gradient: the N axis is ncls groups of cls_cols channels; group q writes the input pixel
  // (out_s*y + cls_iy[q], out_s*x + cls_ix[q]) of its tile pixel (y, x)  [ncls == 0: plain channel axis]
  int ncls, cls_cols, out_s, out_H, out_W;
  int cls_iy[4], cls_ix[4];
  long long cls_off[4];
  float work_scale;              // algorithmic / issued flops of the launch (fused classes pad missing taps with zeros); 0 = 1
};

// tc3 engine: one weight operand as scaled fp16 hi / lo' rows (split_f16), its transposed twin (linear layers: the K-major
// operand of the data gradient) and the device scalar holding amax|w|
struct W16 {
  const void *hi = nullptr, *lo = nullptr;
  int ld = 0;
  const void *hiT = nullptr, *loT = nullptr;
  int ldT = 0;
  const float* amax = nullptr;
};
// per-launch operand scalars of the tc3 engine: amax of the activation operand (device), optional amax accumulator of the output
struct Tc3Ctx {
  const float* amax_a;
  float* amax_out;
};

// One implicit-GEMM convolution launch over an NHWC tensor a[Bn, Hin, Win, Ctot] (channels [c_off, c_off+Cin)).
struct ConvOp {
  const float* a;
  int Hin, Win, Ctot, c_off, Cin;
  int KH, KW, sy, sx, py, px;
  int Yn, Xn, Bn;                // output pixel grid
};
// out[pixel, n] = epilogue( sum_{tap,c} a[pixel*s + tap - p, c] * Wp[n, (tap, c)] );  pixel -> out + b*osb + y*osy + x*osx
// act: 0 none, 1 relu, 2 leaky, 3 multiply by relu'(mask), 4 multiply by leaky'(mask)  (mask addressed like out)
int conv_tc_fwd(const ConvOp& o, const float* Wp, int ldw, int N, const float* bias, int act, const float* mask, float* out,
                long long osb, long long osy, long long osx, cudaStream_t s);
// dWp[n, (tap, c)] += sum_pixels dy[pixel, n] * a[pixel*s + tap - p, c]      (dy rows = pixels in (b, y, x) order)
int conv_tc_wgrad(const ConvOp& o, const float* dy, int ldy, int N, float* dWp, int ldw, cudaStream_t s);
bool conv_tc_supported(const ConvOp& o, bool wgrad);

// tc2.cu: persistent TMEM-operand engine (pre-split weights Bhi / Blo = split_hi_lo of the packed weights)
bool tc2_gemm_supported(int form, int M, int N, int K, const float* A, int lda, const float* Bhi, const float* Blo, int ldb);
int tc2_gemm(int form, int M, int N, int K, const float* A, int lda, const float* Bhi, const float* Blo, int ldb, float* C,
             int ldc, const float* bias, int act, const float* mask, cudaStream_t s);
int tc2_conv_fwd(const ConvOp& o, const float* Whi, const float* Wlo, int ldw, int N, const float* bias, int act,
                 const float* mask, float* out, long long osb, long long osy, long long osx, cudaStream_t s,
                 const TcTap* cls = nullptr);     // cls: only its ncls.. fields are read (fused data gradient)
int tc2_wgrad(int Kx, int N, long long rows, const float* x, int ldx, const float* dy, int ldy, float* dW, int ldw,
              cudaStream_t s);
int tc2_conv_wgrad(const ConvOp& o, const float* dy, int ldy, int N, float* dWp, int ldw, cudaStream_t s);
int split_hi_lo(const float* w, float* hi, float* lo, long long n, cudaStream_t s);

// tc3.cu: scaled-fp16 3-product engine (kind::f16, 64-wide K blocks); weights pre-split by split_f16, operand amax scalars
bool tc3_gemm_supported(int M, int N, int K, const float* A, int lda, const void* Bhi, const void* Blo, int ldb16);
int tc3_gemm(int M, int N, int K, const float* A, int lda, const void* Bhi, const void* Blo, int ldb16, const float* amax_a,
             const float* amax_b, float* C, int ldc, const float* bias, int act, const float* mask, float* amax_out,
             cudaStream_t s);
int tc3_conv_fwd(const ConvOp& o, const void* Whi, const void* Wlo, int ldw16, int N, const float* amax_a, const float* amax_b,
                 const float* bias, int act, const float* mask, float* out, long long osb, long long osy, long long osx,
                 float* amax_out, cudaStream_t s, const TcTap* cls = nullptr, const void* a_hi16 = nullptr, const void* a_lo16 = nullptr);
// a_hi16 / a_lo16 (optional, conv launches): the activation operand pre-split into fp16 planes by tc3_presplit (same layout as
// o.a with 2-byte elements, scaled by t3_scale(*amax_a)); needs Cin % 64 == 0.  Both MMA operands are then plain TMA loads.
int tc3_presplit(const float* x, long long n, const float* amax, void* hi, void* lo, cudaStream_t s);
// sign-bit tensors (one bit per element, set where the activation is > 0): a relu / leaky forward launch that writes a
// registered buffer densely also writes its bits; a data gradient whose mask points into a buffer with fresh bits reads those
// instead of the fp32 activation (the net registers nothing when DDRL_NO_SIGNBITS=1)
void tc3_signbits_register(const float* base, size_t elems, unsigned int* bits);
void tc3_signbits_unregister(const float* base);
// db (optional): bias gradient db[n] += sum_r dy[r, n], fused into the kernel's dy conversion (no separate column-sum pass)
int tc3_wgrad(int Kx, int N, long long rows, const float* x, int ldx, const float* dy, int ldy, const float* amax_x,
              const float* amax_dy, float* dW, int ldw, cudaStream_t s, float* db = nullptr, float* db2 = nullptr, int db_split = 0);
bool tc3_conv_wgrad_supported(const ConvOp& o);
int tc3_conv_wgrad(const ConvOp& o, const float* dy, int ldy, int N, const float* amax_x, const float* amax_dy, float* dWp, int ldw,
                   cudaStream_t s, float* db = nullptr, float* db2 = nullptr, int db_split = 0, const void* a_hi16 = nullptr,
                   const void* a_lo16 = nullptr);
int amax_f32(const float* x, long long rows, int cols, long long ld, float* slot, bool zero_first, cudaStream_t s);
int split_f16(const float* w, int N, int K, int ldw, const float* amax, void* hi, void* lo, int ld16, void* hiT, void* loT,
              int ldT16, cudaStream_t s);

// conv_ops.cu: convolution layers on the implicit-GEMM path
struct DgradClass {              // one parity class (pix + pad) mod stride of the data gradient
  int ry, rx;                    // class residues
  int nty, ntx;                  // taps per axis reaching this class
  int iy0, ix0;                  // first input pixel of the class
  int Yn, Xn;                    // pixels of the class per image
  int pady, padx;                // padding of the equivalent stride-1 convolution over dy
  int K;                         // nty*ntx*Cout
  float* wd;                     // packed weights [Cin, K]
  float *wd_hi, *wd_lo;          // their tf32 hi / lo split (tc2 engine)
  W16 w16;                       // scaled fp16 split (tc3 engine)
};
struct DgradFused {              // all parity classes of a strided data gradient as ONE GEMM (tc2 engine)
  bool on;
  int s, ncy, ncx;               // stride, classes per axis
  int nty, ntx, pady, padx;      // unified taps / padding of the stride-1 convolution over dy
  int Jy, Jx;                    // unified pixel grid per image (ceil(H/s), ceil(W/s))
  int K, N;                      // nty*ntx*Cout, ncy*ncx*Cin
  int iy0[2], ix0[2], q0y[2], q0x[2];
  float *wd, *wd_hi, *wd_lo;     // [N, K] (class-major rows), zero where a class has no tap
  W16 w16;                       // scaled fp16 split (tc3 engine)
};
int conv_dgrad_fused_plan(const ConvGeom& g, int Cout, DgradFused& f);
int pack_dgrad_fused(const float* w_oihw, const ConvGeom& g, int Cout, const DgradFused& f, cudaStream_t s);
int conv_dgrad_fused_tc2(const ConvGeom& g, int Cout, const DgradFused& f, const float* dy, float* dx, int act,
                         const float* mask, int B, cudaStream_t s, int out_ctot = 0, int out_coff = 0,
                         const Tc3Ctx* t3 = nullptr);          // t3 != nullptr: the tc3 engine (f.w16)
ConvOp conv_op_fwd(const ConvGeom& g, const float* x, int Ctot, int c_off, int B);
int conv_dgrad_plan(const ConvGeom& g, int Cout, std::vector<DgradClass>& out);
int pack_dgrad(const float* w_oihw, const ConvGeom& g, int Cout, const DgradClass& c, cudaStream_t s);
int conv_dgrad_tc(const ConvGeom& g, int Cout, const std::vector<DgradClass>& cls, const float* dy, int dy_ctot, int dy_coff,
                  float* dx, int act, const float* mask, int B, cudaStream_t s, bool tmem_engine = false,
                  const Tc3Ctx* t3 = nullptr);                  // t3 != nullptr: the tc3 engine (c.w16)
bool conv_dgrad_supported(const ConvGeom& g, int Cout, const std::vector<DgradClass>& cls, const float* dy, int dy_ctot,
                          int dy_coff, int B);

// GEMM engines --------------------------------------------------------------------------
int gemm_simt(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
              const float* bias, int act, int beta, int trans_c, cudaStream_t s);
int gemm_tc(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, int act, int beta, int trans_c, cudaStream_t s, const float* mask = nullptr);
bool gemm_tc_supported(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, const float* C,
                       int ldc, int trans_c);

// layer glue ----------------------------------------------------------------------------
int im2col(const ConvGeom& g, const float* x, float* cols, int B, cudaStream_t s);
int col2im(const ConvGeom& g, const float* dcols, float* dx, int B, cudaStream_t s);
// amax_slot (optional, tc3 engine): max|output| is accumulated there on the way (atomicMax on the bits; zeroed by the caller)
int pool_fwd(const float* a, float* out, uint8_t* idx, int B, int H, int W, int C, cudaStream_t s, float* amax_slot = nullptr);
int pool_bwd(const float* dout, const uint8_t* idx, const float* a, float* da, int B, int H, int W, int C, cudaStream_t s,
             float* amax_slot = nullptr);
int act_bwd(float* dy, int ld_dy, const float* y, int ld_y, long long rows, int colsN, int act, cudaStream_t s);
// db2 != nullptr: columns [0, split) accumulate into db, [split, N) into db2 (one pass over a fused two-layer gradient)
int colsum_add(const float* dy, int ld, long long rows, int N, float* db, cudaStream_t s, float* db2 = nullptr, int split = 0);
int pack_weight(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s);
int unpack_grad(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s);
// thin-K layers (K <= 36, N <= 64): register-resident weights / partial sums, HBM-streaming (layer_ops.cu)
bool thin_supported(long long M, int N, int K, const float* x, int ldx, const float* y, int ldy);
int thin_fwd(const float* x, int ldx, const float* W, int ldw, const float* bias, float* y, long long M, int N, int K, int act,
             cudaStream_t s);
int thin_wgrad(const float* x, int ldx, const float* dy, float* dW, int ldw, long long M, int N, int K, cudaStream_t s);
// first-layer strided valid convs as stride-1 convs over the space-to-depth observation (layer_ops.cu)
// amax_slot (optional, tc3 engine): the kernel also leaves max|x| of the frames it moves there (atomicMax on the bits; the
// caller zeroes the slot) -- the scale of the first conv's activation operand, without a pass of its own over the tensor
int space_to_depth(const ConvGeom& g, const float* x, float* out, int B, cudaStream_t s, float* amax_slot = nullptr);
int pack_weight_s2d(const float* w_oihw, float* dst, int O, int C, int KH, int KW, int stride, int ld, cudaStream_t s);
int unpack_grad_s2d(const float* src, float* dst_oihw, int O, int C, int KH, int KW, int stride, int ld, cudaStream_t s);
int copy2d(const float* src, int ld_s, float* dst, int ld_d, long long rows, int colsN, cudaStream_t s);
int skinny_fwd(const float* x, int ldx, const float* W, const float* bias, int B, int N, int K, float* y, int ldy,
               cudaStream_t s);
int skinny_dgrad(const float* dy, int ldy, const float* W, int B, int N, int K, float* dx, int ldx, int accumulate,
                 cudaStream_t s);
int skinny_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int N, int K, float* dW, float* db,
                 cudaStream_t s);

// device-resident job table of one weight-preparation phase (prep.cu)
struct PrepJob;
struct PrepTable {
  void* dev_jobs = nullptr;
  int* dev_starts = nullptr;
  int njobs = 0, total_blocks = 0;
  int upload(const std::vector<PrepJob>& jobs);       // synchronous (cudaMalloc + cudaMemcpy): never inside a capture
  int launch(const char* name, cudaStream_t s) const;
  void clear();
};

// K7 (adam.cu)
int clip_adam_launch(float* params, const float* grads, float* m, float* v, long long n, const long long* seg_begin,
                     const float* seg_lr, int nseg, int step, const ddrl_ppo_hparams* hp, float* norm_out,
                     cudaStream_t s, double* sumsq_scratch);
// graph replay of an optimiser step: rewrites the step-dependent arguments of the captured clip_adam_kernel node
const void* clip_adam_kernel_func();
int clip_adam_update_node(cudaGraphExec_t exec, cudaGraphNode_t node, const long long* seg_begin, const float* seg_lr, int nseg,
                          int step, const ddrl_ppo_hparams* hp);

}  // namespace ddrl
