// fp32 CUDA-core GEMM (FFMA), the numerically plain engine mode (DDRL_GEMM_SIMT_F32).
//
// It is the in-repo fp32 yardstick the tcgen05 3xTF32 kernels (gemm_tc.cu) are tested against,
// and the engine used for the skinny head GEMMs (N = 1..28) where a 128-wide tensor-core tile
// would idle.  Three operand forms cover forward, data-gradient and weight-gradient of every
// Linear / im2col'd Conv layer of the reference encoders (nn/atari_encoder.py, nn/nav_encoder.py):
//   form 0  C[m,n] = sum_k A[m,k] B[n,k]   (A [M,K] row-major, B [N,K] row-major)   y = x W^T
//   form 1  C[m,n] = sum_k A[m,k] B[k,n]   (B [K,N] row-major)                       dx = dy W
//   form 2  C[m,n] = sum_k A[k,m] B[k,n]   (A [K,M], B [K,N] row-major)              dW = dy^T x
// Tiling: 128 x BN x 16 block tile, 256 threads as 16x16, 8 x (BN/16) register tile,
// register-prefetched double buffering through shared memory, float4 global loads where the
// contiguous axis allows.  Split-K (grid.z) with atomic accumulation for the weight-gradient
// shapes whose M*N is small and K = batch*pixels is huge.
#include <algorithm>

#include "common.cuh"

namespace ddrl {

constexpr int BM = 128, BK = 16, PAD = 4;

struct GemmArgs {
  const float* A; const float* B; float* C; const float* bias;
  int M, N, K;
  long long sAm, sAk, sBn, sBk;
  long long sCm, sCn;     // C element strides (sCn = 1 normally; swapped for a transposed store)
  int act, beta, k_per_split, atomic;
};

// MODE 0: scalar, generic strides. MODE 1: contiguous along k (float4 along k).
// MODE 2: contiguous along the m/n axis (float4 along m/n).
template <int ROWS, int MODE>
struct TileLoader {
  static constexpr int ELEMS = ROWS * BK;
  static constexpr int PER_THREAD = (ELEMS + 255) / 256;        // scalars
  static constexpr int V4 = (ELEMS / 4 + 255) / 256;            // float4 slots
  float r[MODE == 0 ? PER_THREAD : V4 * 4];

  __device__ __forceinline__ void load(const float* __restrict__ base, long long s_row, long long s_k, int row0,
                                       int nrows, int k0, int kend, int tid) {
    if (MODE == 0) {
#pragma unroll
      for (int s = 0; s < PER_THREAD; ++s) {
        const int idx = tid + s * 256;
        const int row = idx % ROWS, k = idx / ROWS;
        float v = 0.f;
        if (idx < ELEMS && row0 + row < nrows && k0 + k < kend) v = base[(row0 + row) * s_row + (k0 + k) * s_k];
        r[s] = v;
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int s = 0; s < V4; ++s) {
        const int idx = tid + s * 256;
        const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ELEMS / 4 && row0 + row < nrows) {
          const float* p = base + (row0 + row) * s_row + (k0 + kq);
          if (k0 + kq + 3 < kend) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (k0 + kq + 0 < kend) v.x = p[0];
            if (k0 + kq + 1 < kend) v.y = p[1];
            if (k0 + kq + 2 < kend) v.z = p[2];
          }
        }
        r[s * 4 + 0] = v.x; r[s * 4 + 1] = v.y; r[s * 4 + 2] = v.z; r[s * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int s = 0; s < V4; ++s) {
        const int idx = tid + s * 256;
        const int k = idx / (ROWS / 4), rq = (idx % (ROWS / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ELEMS / 4 && k0 + k < kend) {
          const float* p = base + (k0 + k) * s_k + (row0 + rq);
          if (row0 + rq + 3 < nrows) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (row0 + rq + 0 < nrows) v.x = p[0];
            if (row0 + rq + 1 < nrows) v.y = p[1];
            if (row0 + rq + 2 < nrows) v.z = p[2];
          }
        }
        r[s * 4 + 0] = v.x; r[s * 4 + 1] = v.y; r[s * 4 + 2] = v.z; r[s * 4 + 3] = v.w;
      }
    }
  }

  // shared tile is [BK][ROWS + PAD]
  __device__ __forceinline__ void store(float* __restrict__ sm, int tid) const {
    constexpr int LD = ROWS + PAD;
    if (MODE == 0) {
#pragma unroll
      for (int s = 0; s < PER_THREAD; ++s) {
        const int idx = tid + s * 256;
        if (idx < ELEMS) sm[(idx / ROWS) * LD + (idx % ROWS)] = r[s];
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int s = 0; s < V4; ++s) {
        const int idx = tid + s * 256;
        if (idx < ELEMS / 4) {
          const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) sm[(kq + j) * LD + row] = r[s * 4 + j];
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < V4; ++s) {
        const int idx = tid + s * 256;
        if (idx < ELEMS / 4) {
          const int k = idx / (ROWS / 4), rq = (idx % (ROWS / 4)) * 4;
          *reinterpret_cast<float4*>(sm + k * LD + rq) = make_float4(r[s * 4], r[s * 4 + 1], r[s * 4 + 2], r[s * 4 + 3]);
        }
      }
    }
  }
};

template <int BN, int AMODE, int BMODE>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  constexpr int TM = 8, TN = BN / 16;
  constexpr int LDA = BM + PAD, LDB = BN + PAD;
  __shared__ __align__(16) float As[2][BK * LDA];
  __shared__ __align__(16) float Bs[2][BK * LDB];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  TileLoader<BM, AMODE> la;
  TileLoader<BN, BMODE> lb;
  const int ntiles = (kend - kbeg + BK - 1) / BK;
  if (ntiles > 0) {
    la.load(g.A, g.sAm, g.sAk, m0, g.M, kbeg, kend, tid);
    lb.load(g.B, g.sBn, g.sBk, n0, g.N, kbeg, kend, tid);
    la.store(As[0], tid);
    lb.store(Bs[0], tid);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < ntiles) {
      la.load(g.A, g.sAm, g.sAk, m0, g.M, kbeg + (t + 1) * BK, kend, tid);
      lb.load(g.B, g.sBn, g.sBk, n0, g.N, kbeg + (t + 1) * BK, kend, tid);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(as + kk * LDA + ty * TM);
      const float4 a1 = *reinterpret_cast<const float4*>(as + kk * LDA + ty * TM + 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * LDB + tx * TN);
        const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * LDB + tx * TN + 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * LDB + tx * TN);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = bs[kk * LDB + tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      la.store(As[cur ^ 1], tid);
      lb.store(Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

  // epilogue
  const bool add_bias = g.bias != nullptr && (!g.atomic || blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (add_bias) v += g.bias[n];
      float* c = g.C + m * g.sCm + n * g.sCn;
      if (g.atomic) {
        atomicAdd(c, v);
      } else {
        if (g.beta) v += *c;
        if (g.act == 1) v = fmaxf(v, 0.f);
        else if (g.act == 2) v = v > 0.f ? v : 0.01f * v;
        *c = v;
      }
    }
  }
}

template <int BN>
static int launch_bn(const GemmArgs& g, int form, bool vec, dim3 grid, cudaStream_t s) {
  if (!vec) {
    sgemm_kernel<BN, 0, 0><<<grid, 256, 0, s>>>(g);
  } else if (form == 0) {
    sgemm_kernel<BN, 1, 1><<<grid, 256, 0, s>>>(g);
  } else if (form == 1) {
    sgemm_kernel<BN, 1, 2><<<grid, 256, 0, s>>>(g);
  } else {
    sgemm_kernel<BN, 2, 2><<<grid, 256, 0, s>>>(g);
  }
  prof_work(2.0 * g.M * (double)g.N * g.K);
  DDRL_LAUNCHED("sgemm_kernel");
  return DDRL_OK;
}

// trans_c: store C[m,n] at C[n*ldc + m] (bias is then indexed by m is NOT supported: bias must be NULL)
int gemm_simt(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
              const float* bias, int act, int beta, int trans_c, cudaStream_t s) {
  if (M <= 0 || N <= 0) return DDRL_OK;
  if (trans_c && bias) return DDRL_E_ARG;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.bias = bias; g.M = M; g.N = N; g.K = K; g.act = act; g.beta = beta;
  g.sCm = trans_c ? 1 : ldc; g.sCn = trans_c ? ldc : 1;
  if (form == 0) { g.sAm = lda; g.sAk = 1; g.sBn = ldb; g.sBk = 1; }
  else if (form == 1) { g.sAm = lda; g.sAk = 1; g.sBn = 1; g.sBk = ldb; }
  else if (form == 2) { g.sAm = 1; g.sAk = lda; g.sBn = 1; g.sBk = ldb; }
  else return DDRL_E_ARG;
  const bool vec = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0;
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : (N > 16 ? 32 : 16));
  const int tiles = ceil_div(M, BM) * ceil_div(N, bn);
  int splits = 1;
  if (act == 0 && K >= 4096 && tiles < 2 * kNumSMs) {
    splits = std::min(std::min(ceil_div(4 * kNumSMs, tiles), ceil_div(K, 1024)), 512);
  }
  int kps = ceil_div(ceil_div(K, splits), BK) * BK;
  splits = ceil_div(K, kps);
  g.k_per_split = kps;
  g.atomic = splits > 1;
  if (g.atomic && !beta) {
    // zero the destination tile rows first (C may be a strided sub-block)
    const int rows = trans_c ? N : M, cols = trans_c ? M : N;
    if (ldc == cols) DDRL_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)rows * cols, s));
    else DDRL_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * cols, rows, s));
  }
  dim3 grid(ceil_div(M, BM), ceil_div(N, bn), splits);
  if (grid.y > 65535 || grid.z > 65535) return DDRL_E_UNSUPPORTED;
  switch (bn) {
    case 128: return launch_bn<128>(g, form, vec, grid, s);
    case 64: return launch_bn<64>(g, form, vec, grid, s);
    case 32: return launch_bn<32>(g, form, vec, grid, s);
    default: return launch_bn<16>(g, form, vec, grid, s);
  }
}

}  // namespace ddrl
