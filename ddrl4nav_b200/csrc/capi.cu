// Library-level entry points of the C ABI (include/ddrl_b200.h).
#include "common.cuh"
#include "layer_ops.h"

namespace ddrl {
thread_local char g_cuda_err[256] = {0};
long long g_launches = 0;
}  // namespace ddrl

extern "C" int ddrl_version(void) { return 100; }

extern "C" const char* ddrl_error_string(int code) {
  switch (code) {
    case DDRL_OK: return "ok";
    case DDRL_E_ARG: return "bad argument";
    case DDRL_E_CUDA: return "CUDA runtime error (see ddrl_last_cuda_error)";
    case DDRL_E_STATE: return "call order / unbound buffers";
    case DDRL_E_UNSUPPORTED: return "unsupported configuration";
    case DDRL_E_NOMEM: return "device out of memory";
  }
  return "unknown error";
}

extern "C" const char* ddrl_last_cuda_error(void) { return ddrl::g_cuda_err; }
extern "C" int64_t ddrl_launch_count(void) { return ddrl::g_launches; }
extern "C" void ddrl_launch_count_reset(void) { ddrl::g_launches = 0; }

extern "C" int ddrl_gemm_f32(int mode, int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                             float* C, int ldc, const float* bias, int act, int beta, void* stream) {
  using namespace ddrl;
  if (M < 0 || N < 0 || K < 0 || form < 0 || form > 2 || act < 0 || act > 2) return DDRL_E_ARG;
  if (M == 0 || N == 0) return DDRL_OK;
  if (!A || !B || !C) return DDRL_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == DDRL_GEMM_TC_3XTF32) {
    if (!gemm_tc_supported(form, M, N, K, A, lda, B, ldb, C, ldc, 0)) return DDRL_E_UNSUPPORTED;
    return gemm_tc(form, M, N, K, A, lda, B, ldb, C, ldc, bias, act, beta, 0, s);
  }
  if (mode != DDRL_GEMM_SIMT_F32) return DDRL_E_ARG;
  return gemm_simt(form, M, N, K, A, lda, B, ldb, C, ldc, bias, act, beta, 0, s);
}
