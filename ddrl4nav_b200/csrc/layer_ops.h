// Internal (non-ABI) declarations shared by the engine translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

// GEMM engines --------------------------------------------------------------------------
int gemm_simt(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
              const float* bias, int act, int beta, int trans_c, cudaStream_t s);
int gemm_tc(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, int act, int beta, int trans_c, cudaStream_t s);
bool gemm_tc_supported(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, const float* C,
                       int ldc, int trans_c);

// layer glue ----------------------------------------------------------------------------
int im2col(const ConvGeom& g, const float* x, float* cols, int B, cudaStream_t s);
int col2im(const ConvGeom& g, const float* dcols, float* dx, int B, cudaStream_t s);
int pool_fwd(const float* a, float* out, uint8_t* idx, int B, int H, int W, int C, cudaStream_t s);
int pool_bwd(const float* dout, const uint8_t* idx, const float* a, float* da, int B, int H, int W, int C, cudaStream_t s);
int act_bwd(float* dy, int ld_dy, const float* y, int ld_y, long long rows, int colsN, int act, cudaStream_t s);
int colsum_add(const float* dy, int ld, long long rows, int N, float* db, cudaStream_t s);
int pack_weight(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s);
int unpack_grad(const float* src, float* dst, int O, int I, int J, int ld, cudaStream_t s);
int copy2d(const float* src, int ld_s, float* dst, int ld_d, long long rows, int colsN, cudaStream_t s);
int skinny_fwd(const float* x, int ldx, const float* W, const float* bias, int B, int N, int K, float* y, int ldy,
               cudaStream_t s);
int skinny_dgrad(const float* dy, int ldy, const float* W, int B, int N, int K, float* dx, int ldx, int accumulate,
                 cudaStream_t s);
int skinny_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int N, int K, float* dW, float* db,
                 cudaStream_t s);

// K7 (adam.cu)
int clip_adam_launch(float* params, const float* grads, float* m, float* v, long long n, const long long* seg_begin,
                     const float* seg_lr, int nseg, int step, const ddrl_ppo_hparams* hp, float* norm_out,
                     cudaStream_t s);

}  // namespace ddrl
