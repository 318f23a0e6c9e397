// tcgen05 3xTF32 GEMM engine -- placeholder until the kernel lands (reports "unsupported" so the
// engine stays on the fp32 SIMT path; nothing is emulated or faked).
#include "common.cuh"
#include "layer_ops.h"

namespace ddrl {
bool gemm_tc_supported(int, int, int, int, const float*, int, const float*, int, const float*, int, int) { return false; }
int gemm_tc(int, int, int, int, const float*, int, const float*, int, float*, int, const float*, int, int, int, cudaStream_t) {
  return DDRL_E_UNSUPPORTED;
}
}  // namespace ddrl
