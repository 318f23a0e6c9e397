// Library-level entry points of the C ABI (include/ddrl_b200.h).
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "layer_ops.h"

namespace ddrl {
thread_local char g_cuda_err[256] = {0};
long long g_launches = 0;

// ---- deterministic mode ------------------------------------------------------------------
bool g_deterministic = [] { const char* e = getenv("DDRL_DETERMINISTIC"); return e && e[0] == '1'; }();
static unsigned int* g_det_pool[64] = {};
static size_t g_det_next[64] = {};
constexpr size_t kDetPool = 1 << 16;
unsigned int* det_counters(int n) {
  const int dev = current_device_index();
  if (n < 1 || (size_t)n > kDetPool) return nullptr;
  if (!g_det_pool[dev]) {
    if (cudaMalloc(&g_det_pool[dev], kDetPool * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(g_det_pool[dev], 0, kDetPool * sizeof(unsigned int));
  }
  // round robin: counters reset themselves when their chain completes, and a range comes round again only after 65536
  // later counters have been handed to kernels that run after this one (stream order)
  if (g_det_next[dev] + n > kDetPool) g_det_next[dev] = 0;
  unsigned int* p = g_det_pool[dev] + g_det_next[dev];
  g_det_next[dev] += n;
  return p;
}

// ---- per-kernel-class device timing ------------------------------------------------------
bool g_prof_on = false;
double g_prof_work = 0.0;
static cudaStream_t g_prof_stream = nullptr;
struct ProfRec { cudaEvent_t ev; const char* name; double work; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}

// shape-tagged kernel names (DDRL_PROF_SHAPES=1): interned so the recorded pointer stays valid
bool g_prof_shapes = false;
const char* prof_intern(const std::string& s) {
  static std::map<std::string, int> pool;
  return pool.emplace(s, 0).first->first.c_str();
}

void prof_record(const char* name) {
  ProfRec r{prof_event(), name, g_prof_work};
  g_prof_work = 0.0;
  cudaEventRecord(r.ev, g_prof_stream);
  g_prof_recs.push_back(r);
}
}  // namespace ddrl

extern "C" int ddrl_version(void) { return 100; }

extern "C" int ddrl_set_deterministic(int on) {
  const int was = ddrl::g_deterministic ? 1 : 0;
  ddrl::g_deterministic = on != 0;
  if (on) ddrl::det_counters(1);     // allocate the current device's counter pool now: never inside a graph capture
  return was;
}

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

extern "C" int ddrl_prof_start(void* stream) {
  using namespace ddrl;
  for (auto& r : g_prof_recs) g_prof_pool.push_back(r.ev);
  g_prof_recs.clear();
  g_prof_stream = (cudaStream_t)stream;
  g_prof_on = true;
  const char* e = getenv("DDRL_PROF_SHAPES");
  g_prof_shapes = e && e[0] == '1';
  prof_record("__start__");
  return DDRL_OK;
}

// Writes "name ms launches work\n" lines (work = sum of the algorithmic flops/bytes the launchers declared).
extern "C" int ddrl_prof_stop(char* out, int cap) {
  using namespace ddrl;
  g_prof_on = false;
  if (g_prof_recs.empty()) return DDRL_E_STATE;
  DDRL_CUDA(cudaEventSynchronize(g_prof_recs.back().ev));
  struct Acc { double ms = 0, work = 0; long n = 0; };
  std::map<std::string, Acc> acc;
  for (size_t i = 1; i < g_prof_recs.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_prof_recs[i - 1].ev, g_prof_recs[i].ev);
    Acc& a = acc[g_prof_recs[i].name];
    a.ms += ms; a.work += g_prof_recs[i].work; a.n += 1;
  }
  std::string s;
  char line[256];
  for (auto& kv : acc) {
    snprintf(line, sizeof(line), "%s %.6f %ld %.6e\n", kv.first.c_str(), kv.second.ms, kv.second.n, kv.second.work);
    s += line;
  }
  if (out && cap > 0) { strncpy(out, s.c_str(), cap - 1); out[cap - 1] = 0; }
  return DDRL_OK;
}

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
  if (mode == DDRL_GEMM_TC2_TMEM) {
    if (form == 2) {
      // C[m,n] (+)= sum_k A[k,m] B[k,n]: the engine's weight-gradient form with dy := A, x := B (output row = dy column)
      if (bias || act) return DDRL_E_UNSUPPORTED;
      if (!beta) DDRL_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, s));
      return tc2_wgrad(N, M, K, B, ldb, A, lda, C, ldc, s);
    }
    if (beta) return DDRL_E_UNSUPPORTED;
    const size_t nb = (size_t)(form == 0 ? N : K) * ldb;
    const size_t nb4 = (nb + 3) & ~size_t(3);
    float* tmp = nullptr;
    DDRL_CUDA(cudaMalloc(&tmp, sizeof(float) * 3 * nb4));
    int r = DDRL_OK;
    if (cudaMemsetAsync(tmp, 0, sizeof(float) * 3 * nb4, s) != cudaSuccess ||
        cudaMemcpyAsync(tmp, B, sizeof(float) * nb, cudaMemcpyDeviceToDevice, s) != cudaSuccess) r = DDRL_E_CUDA;
    if (r == DDRL_OK) r = split_hi_lo(tmp, tmp + nb4, tmp + 2 * nb4, (long long)nb4, s);
    if (r == DDRL_OK) r = tc2_gemm(form, M, N, K, A, lda, tmp + nb4, tmp + 2 * nb4, ldb, C, ldc, bias, act, nullptr, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    return r;
  }
  if (mode == DDRL_GEMM_TC3_F16) {
    if (form == 2) {
      // C[m,n] (+)= sum_k A[k,m] B[k,n]: the engine's weight-gradient form with dy := A, x := B (output row = dy column)
      if (bias || act) return DDRL_E_UNSUPPORTED;
      if (!beta) DDRL_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, s));
      float* slots = nullptr;
      DDRL_CUDA(cudaMalloc(&slots, 256));
      int r = DDRL_OK;
      if (cudaMemsetAsync(slots, 0, 256, s) != cudaSuccess) r = DDRL_E_CUDA;
      if (r == DDRL_OK) r = amax_f32(B, K, N, ldb, slots, false, s);
      if (r == DDRL_OK) r = amax_f32(A, K, M, lda, slots + 1, false, s);
      if (r == DDRL_OK) r = tc3_wgrad(N, M, K, B, ldb, A, lda, slots, slots + 1, C, ldc, s);
      cudaStreamSynchronize(s);
      cudaFree(slots);
      return r;
    }
    if (beta) return DDRL_E_UNSUPPORTED;
    // split B on the fly: K-major hi / lo' [N, ld16] (form 1: through the transposing output of the split)
    const int ld16 = (K + 7) & ~7, ldN16 = (N + 7) & ~7;
    const size_t halfs = (size_t)N * ld16 + (form == 1 ? (size_t)K * ldN16 : 0);
    char* tmp = nullptr;
    DDRL_CUDA(cudaMalloc(&tmp, 256 + 4 * halfs));
    float* slots = reinterpret_cast<float*>(tmp);
    char* h0 = tmp + 256;
    int r = DDRL_OK;
    if (cudaMemsetAsync(tmp, 0, 256 + 4 * halfs, s) != cudaSuccess) r = DDRL_E_CUDA;
    if (r == DDRL_OK) r = amax_f32(A, M, K, lda, slots, false, s);
    if (form == 0) {
      char *hi = h0, *lo = h0 + 2 * halfs;
      if (r == DDRL_OK) r = amax_f32(B, N, K, ldb, slots + 1, false, s);
      if (r == DDRL_OK) r = split_f16(B, N, K, ldb, slots + 1, hi, lo, ld16, nullptr, nullptr, 0, s);
      if (r == DDRL_OK) r = tc3_gemm(M, N, K, A, lda, hi, lo, ld16, slots, slots + 1, C, ldc, bias, act, nullptr, slots + 2, s);
    } else {
      // B is [K, N]: split it as a [K rows, N cols] matrix and use the transposed pair [N, ld16]
      char *hiT = h0, *loT = h0 + 2 * (size_t)N * ld16;
      char *hi = h0 + 4 * (size_t)N * ld16, *lo = hi + 2 * (size_t)K * ldN16;
      if (r == DDRL_OK) r = amax_f32(B, K, N, ldb, slots + 1, false, s);
      if (r == DDRL_OK) r = split_f16(B, K, N, ldb, slots + 1, hi, lo, ldN16, hiT, loT, ld16, s);
      if (r == DDRL_OK) r = tc3_gemm(M, N, K, A, lda, hiT, loT, ld16, slots, slots + 1, C, ldc, bias, act, nullptr, slots + 2, s);
    }
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    return r;
  }
  if (mode != DDRL_GEMM_SIMT_F32) return DDRL_E_ARG;
  return gemm_simt(form, M, N, K, A, lda, B, ldb, C, ldc, bias, act, beta, 0, s);
}
