// tcgen05 GEMM engine, fp32 in / fp32 out, 3xTF32 split ("parity mode", DDRL_GEMM_TC_3XTF32).
//
// Every Linear / Conv (as GEMM) of the reference encoders runs fp32 on the CPU (nn/atari_encoder.py,
// nn/nav_encoder.py, autograd for backward).  A single TF32 pass (10-bit mantissa) cannot meet the 1e-5
// parity target, so each fp32 operand x is split ON CHIP into hi = rn_tf32(x), lo = rn_tf32(x - hi) and the
// product is accumulated as lo*hi + hi*lo + hi*hi in an fp32 TMEM accumulator (error ~2^-22 per product,
// unbiased) -- three tcgen05.mma.kind::tf32 per K-step.
//
// Kernel anatomy (one 128 x BN output tile per CTA, optional split-K over grid.z):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B-swizzled) of raw fp32 A and B into a smem ring
//   warps 4-11  splitters: rewrite each landed tile in place as hi and write lo to its twin buffer
//               (element-wise, so the swizzled layout is preserved), fence.proxy.async, signal
//   warp 1      MMA issuer: one elected thread issues 3 x (BK/8) tcgen05.mma per stage, tcgen05.commit frees
//               the stage; final commit publishes the accumulator
//   warps 4-11  epilogue: tcgen05.ld the accumulator (TMEM lane = output row), bias + activation, store
//               (plain / transposed / atomic for split-K and gradient accumulation)
//   warp 2      TMEM allocator
// Operand forms (same as gemm_simt.cu): K-major operands use one TMA box [rows x 32 floats]; MN-major
// operands (weight-gradient and data-gradient GEMMs) use 32x32 boxes with the 128B/32B-atom TMA swizzle so that
// the shared-memory image is the canonical UMMA MN-major SWIZZLE_128B_BASE32B layout (LBO = 4096 B between
// 32-wide groups, SBO = 512 B between 4-deep K groups).  Out-of-range rows/columns/K are zero-filled by TMA.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "layer_ops.h"
#include "tc_ptx.cuh"

namespace ddrl {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                     // floats per stage along K (= 128 B = one swizzle row)
constexpr int TC_THREADS = 384;
constexpr int TC_SPLIT_WARPS = 8;

struct TcArgs {
  float* C;
  const float* bias;
  const float* mask;                           // act 3/4: C = acc * act'(mask[same address])  (fused activation backward)
  long long sCm, sCn;
  int M, N, K;
  int kb_total;                                // K blocks in all (plain: ceil(K/32); wgrad taps: pixel blocks)
  int kb_per_split;                            // K blocks per grid.z slice
  int act, atomic;
  int vec_store;                               // plain row-major C with 16-byte aligned rows: float4 epilogue stores
  TcTap tap;                                   // tap.mode != 0: implicit-GEMM convolution operands (layer_ops.h)
};

// The tensor core adds into its fp32 accumulator with TRUNCATION (measured on B200: all-positive tf32-exact
// inputs lose ~1 ulp per accumulating MMA, -7e-6 relative after 128 MMAs).  To stay at fp32-FFMA accuracy the
// main hi*hi product is accumulated in TMEM only over TC_CHUNK stages (K = 128, 16 MMAs), then drained and added
// with round-to-nearest into fp32 registers of the splitter warps (two TMEM buffers, so the MMA pipe never waits);
// the small lo*hi + hi*lo terms go to a third TMEM accumulator whose truncation is 2^-11 smaller.
constexpr int TC_CHUNK = 4;

template <int BN>
struct TcCfg {
  static constexpr int BNS = (BN + 31) / 32 * 32;                // smem rows of the B tile
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
  static constexpr int B_BYTES = BNS * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);    // hi(raw) + lo for A and B
  // BN = 128: one CTA per SM with a 3-deep ring; BN <= 64: two co-resident CTAs (2-deep rings) so that one CTA's
  // prologue / epilogue overlaps the other's main loop (the conv GEMMs of this path have K <= 1600)
  static constexpr int CTAS_PER_SM = BN <= 64 ? 2 : 1;
  static constexpr int STAGES = BN <= 64 ? 2 : 3;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN <= 32 ? 128 : (BN <= 64 ? 256 : 512);   // main0 | main1 | corr
  static constexpr int COLS = BN / 2;                            // accumulator columns held by one splitter warp
};

// barrier slots: [0,S) full (TMA landed)  [S,2S) split done  [2S,3S) empty (MMA consumed)
//                [3S,3S+2) accumulator buffer ready  [3S+2,3S+4) accumulator buffer drained
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<BN>::CTAS_PER_SM)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs g) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_dyn[];
  // 1024-byte alignment by POINTER arithmetic on the __shared__ array: an integer round-trip hides the address space from
  // the compiler, which then emits generic LD / ST (long-scoreboard, L1TEX path) for every shared-memory access below
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  const TcTap& tp = g.tap;
  const bool tapA = tp.mode != 0;
  // implicit-GEMM forward / data-gradient: this CTA's M tile is a box of output pixels (nb images x ny rows x Xn)
  int b0 = 0, y0 = 0;
  if (tapA && !A_MN) { b0 = (blockIdx.x / tp.tpi) * tp.nb; y0 = (blockIdx.x % tp.tpi) * tp.ny; }
  const int kb_total = g.kb_total;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(kb_total, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;
  const int nchunks = (nkb + TC_CHUNK - 1) / TC_CHUNK;
  // implicit-GEMM weight gradient: a K block is a box of <= 32 pixels; rows the boxes do not cover must read as zero
  if (tapA && A_MN) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < S * Cfg::STAGE_BYTES / 16; i += TC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(bars + s), 1);
      mbar_init(smem_u32(bars + S + s), TC_SPLIT_WARPS);
      mbar_init(smem_u32(bars + 2 * S + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(bars + 3 * S + b), 1);
      mbar_init(smem_u32(bars + 3 * S + 2 + b), TC_SPLIT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        const int kbi = kb0 + i;
        mbar_wait(smem_u32(bars + 2 * S + s), ph ^ 1);
        const uint32_t full = smem_u32(bars + s);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const uint32_t a_dst = smem_u32(st), b_dst = smem_u32(st + 2 * Cfg::A_BYTES);
        const int k = kbi * TC_BK;
        if (!tapA) {
          mbar_expect_tx(full, Cfg::A_BYTES + Cfg::B_BYTES);
          if (!A_MN) {
            tma_load_2d(&tmA, full, a_dst, k, m0);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / 32; ++j) tma_load_2d(&tmA, full, a_dst + j * 4096, m0 + j * 32, k);
          }
          if (!B_MN) {
            tma_load_2d(&tmB, full, b_dst, k, n0);
          } else {
#pragma unroll
            for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_2d(&tmB, full, b_dst + j * 4096, n0 + j * 32, k);
          }
        } else if (!A_MN) {
          // forward / data-gradient: K block = (tap, 32-channel chunk); one 4-D box of the NHWC activation tensor
          const int tap = kbi / tp.cpb, cc = kbi - tap * tp.cpb;
          const int kh = tap / tp.KW, kw = tap - kh * tp.KW;
          mbar_expect_tx(full, tp.rows * 128 + Cfg::B_BYTES);
          tma_load_4d(&tmA, full, a_dst, tp.c_off + cc * 32, kw - tp.px, y0 * tp.sy + kh - tp.py, b0);
          if (!B_MN) {
            tma_load_2d(&tmB, full, b_dst, k, n0);
          } else {
#pragma unroll
            for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_2d(&tmB, full, b_dst + j * 4096, n0 + j * 32, k);
          }
        } else {
          // weight gradient: K block = box of pixels (image b, rows yy0..); M = 4 (tap, chunk) slices of the im2col K axis
          const int b = kbi / tp.tpi, yy0 = (kbi - b * tp.tpi) * tp.ny;
          int na = 0;
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j) na += (m0 / 32 + j) < tp.nslices ? 1 : 0;
          mbar_expect_tx(full, (na + Cfg::BNS / 32) * tp.rows * 128);
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j) {
            const int sl = m0 / 32 + j;
            if (sl < tp.nslices) {
              const int tap = sl / tp.cpb, cc = sl - tap * tp.cpb;
              const int kh = tap / tp.KW, kw = tap - kh * tp.KW;
              tma_load_4d(&tmA, full, a_dst + j * 4096, tp.c_off + cc * 32, kw - tp.px, yy0 * tp.sy + kh - tp.py, b);
            }
          }
          // dy as [image][pixel of the image][N]: pixel rows past the image's last row read as zero (the x boxes of
          // those phantom rows may hold valid input pixels)
#pragma unroll
          for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_3d(&tmB, full, b_dst + j * 4096, n0 + j * 32, yy0 * tp.Xn, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, a_major bit15, b_major bit16,
    // N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t tmem_corr = tmem_base + 2 * BN;
    const int ksteps = (tapA && A_MN) ? tp.kpad / 8 : TC_BK / 8;     // pixel-box K blocks may be shorter than 32
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S;
      const uint32_t ph = (i / S) & 1;
      const int c = i / TC_CHUNK, buf = c & 1;
      const bool first_in_chunk = (i % TC_CHUNK) == 0;
      if (first_in_chunk) mbar_wait(smem_u32(bars + 3 * S + 2 + buf), ((c >> 1) & 1) ^ 1);   // buffer drained
      mbar_wait(smem_u32(bars + S + s), ph);
      tc_fence_after();
      if (lane == 0) {
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const uint32_t a_hi = smem_u32(st), a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t b_hi = smem_u32(st + 2 * Cfg::A_BYTES), b_lo = b_hi + Cfg::B_BYTES;
        // MN-major: 32-wide groups 4096 B apart (LBO), 4-deep K groups 512 B apart (SBO); K-major: 8-row groups 1024 B apart
        const uint32_t lbo = 4096;
        const uint32_t tmem_main = tmem_base + buf * BN;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_base = pass == 0 ? a_lo : a_hi;      // lo*hi, hi*lo -> corr ; hi*hi -> main[buf]
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 8; ++k4) {
            if (k4 >= ksteps) break;
            const uint64_t da = umma_desc(a_base + (A_MN ? k4 * 1024 : k4 * 32), A_MN ? lbo : 16, A_MN ? 512 : 1024, A_MN ? 1 : 2);
            const uint64_t db = umma_desc(b_base + (B_MN ? k4 * 1024 : k4 * 32), B_MN ? lbo : 16, B_MN ? 512 : 1024, B_MN ? 1 : 2);
            if (pass < 2) umma_tf32(tmem_corr, da, db, idesc, (i | pass | k4) != 0 ? 1u : 0u);
            else umma_tf32(tmem_main, da, db, idesc, (!first_in_chunk || k4 != 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(bars + 2 * S + s));                 // stage reusable once these MMAs retire
        if ((i % TC_CHUNK) == TC_CHUNK - 1 || i == nkb - 1) umma_commit(smem_u32(bars + 3 * S + buf));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ hi/lo splitters + fp32 accumulators
    const int t = threadIdx.x - 128;                              // 0..255
    const int q = warp & 3;                                       // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;                             // column half
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    constexpr int A_V4 = Cfg::A_BYTES / 16, B_V4 = Cfg::B_BYTES / 16;
    float acc[Cfg::COLS];
#pragma unroll
    for (int j = 0; j < Cfg::COLS; ++j) acc[j] = 0.f;
    int drained = 0;
    auto drain = [&](int c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(bars + 3 * S + buf), (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int j0 = 0; j0 < Cfg::COLS; j0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + t_lane + buf * BN + half * Cfg::COLS + j0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j0 + j] += v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + 3 * S + 2 + buf));
    };
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S;
      const uint32_t ph = (i / S) & 1;
      mbar_wait(smem_u32(bars + s), ph);
      uint8_t* st = smem + s * Cfg::STAGE_BYTES;
#pragma unroll 4
      for (int v = t; v < A_V4 + B_V4; v += TC_SPLIT_WARPS * 32) {
        uint8_t* hi_p = v < A_V4 ? st + v * 16 : st + 2 * Cfg::A_BYTES + (v - A_V4) * 16;
        uint8_t* lo_p = hi_p + (v < A_V4 ? Cfg::A_BYTES : Cfg::B_BYTES);
        const float4 x = *reinterpret_cast<const float4*>(hi_p);
        uint4 h, l;
        h.x = tf32_rn(x.x); h.y = tf32_rn(x.y); h.z = tf32_rn(x.z); h.w = tf32_rn(x.w);
        l.x = tf32_rn(x.x - __uint_as_float(h.x)); l.y = tf32_rn(x.y - __uint_as_float(h.y));
        l.z = tf32_rn(x.z - __uint_as_float(h.z)); l.w = tf32_rn(x.w - __uint_as_float(h.w));
        *reinterpret_cast<uint4*>(hi_p) = h;
        *reinterpret_cast<uint4*>(lo_p) = l;
      }
      fence_async_smem();                                         // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + S + s));
      // the chunk BEFORE the one whose last stage was just split has certainly been issued: drain it now
      if ((i % TC_CHUNK) == TC_CHUNK - 1 && i / TC_CHUNK >= 1) drain(drained++);
    }
    while (drained < nchunks) drain(drained++);
    // ------------------------------------------------------------ epilogue (from registers)
    if (nkb > 0) {
      // output row of this thread: plain GEMM row, or output pixel (image, y, x) of the tile's pixel box
      const int r = q * 32 + lane;
      bool rvalid;
      long long roff;
      if (tapA && !A_MN) {
        const int x = r % tp.Xn, t2 = r / tp.Xn;
        const int yy = t2 % tp.ny, bb = t2 / tp.ny;
        rvalid = r < tp.rows && (b0 + bb) < tp.Bn && (y0 + yy) < tp.Yn;
        roff = (long long)(b0 + bb) * tp.osb + (long long)(y0 + yy) * tp.osy + (long long)x * tp.osx;
      } else {
        rvalid = (m0 + r) < g.M;
        roff = (long long)(m0 + r) * g.sCm;
      }
      const float neg_slope = g.act == 3 ? 0.f : 0.01f;            // act 3: relu' of mask, act 4: leaky'
#pragma unroll
      for (int j0 = 0; j0 < Cfg::COLS; j0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + t_lane + 2 * BN + half * Cfg::COLS + j0, v);     // lo*hi + hi*lo correction
        const int colv = n0 + half * Cfg::COLS + j0;
        if (rvalid && g.vec_store && colv + 16 <= g.N) {
          // 16 consecutive columns of this thread's row: four 16-byte stores
          float* p = g.C + roff + colv;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o, mk = make_float4(1.f, 1.f, 1.f, 1.f);
            if (g.act >= 3) mk = *reinterpret_cast<const float4*>(g.mask + roff + colv + j);
            const float* mv = reinterpret_cast<const float*>(&mk);
            float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = acc[j0 + j + e] + v[j + e];
              if (g.bias != nullptr) x += g.bias[colv + j + e];
              if (g.act == 1) x = fmaxf(x, 0.f);
              else if (g.act == 2) x = x > 0.f ? x : 0.01f * x;
              else if (g.act >= 3) x = mv[e] > 0.f ? x : neg_slope * x;
              ov[e] = x;
            }
            *reinterpret_cast<float4*>(p + j) = o;
          }
        } else if (rvalid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = n0 + half * Cfg::COLS + j0 + j;
            if (col < g.N) {
              float x = acc[j0 + j] + v[j];
              if (g.bias != nullptr && (!g.atomic || blockIdx.z == 0)) x += g.bias[col];
              float* p = g.C + roff + col * g.sCn;
              if (g.atomic) {
                atomicAdd(p, x);
              } else {
                if (g.act == 1) x = fmaxf(x, 0.f);
                else if (g.act == 2) x = x > 0.f ? x : 0.01f * x;
                else if (g.act >= 3) x = g.mask[roff + col * g.sCn] > 0.f ? x : neg_slope * x;
                *p = x;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int tc_get_encode() {
  if (g_encode) return DDRL_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) return cuda_fail(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return DDRL_OK;
}

// generic tiled map (tc3.cu / tc4.cu build their fp16 maps with it); swizzle: 0 SWIZZLE_128B_ATOM_32B, 1 SWIZZLE_128B, 2 SWIZZLE_64B
int tc_encode_tiled(CUtensorMap* m, bool f16, int rank, const void* base, const unsigned long long* dims,
                    const unsigned long long* strides_bytes, const unsigned* box, const unsigned* estr, int swizzle) {
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = g_encode(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(base), d, st, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled(rank %d%s) failed (%d)", rank, f16 ? ", f16" : "", (int)r);
    return DDRL_E_CUDA;
  }
  return DDRL_OK;
}

// 2-D fp32 tensor [outer, inner] with row stride ld (floats); box = [box_outer, 32 floats], 128B swizzle
int tc_make_map(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_outer,
                bool mn_major) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return DDRL_E_CUDA;
  }
  return DDRL_OK;
}

bool gemm_tc_supported(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, const float* C,
                       int ldc, int trans_c) {
  if (M < 1 || N < 16 || K < 1) return false;
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0) return false;
  if (lda % 4 != 0 || ldb % 4 != 0) return false;
  return true;
}

template <int BN>
static int launch_tc(int form, const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, dim3 grid, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static bool attr_done_dev[64][3] = {};
  bool* attr_done = attr_done_dev[current_device_index()];
#define TC_LAUNCH(AMN, BMN)                                                                                              \
  do {                                                                                                                   \
    if (!attr_done[form]) {                                                                                              \
      DDRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); \
      attr_done[form] = true;                                                                                            \
    }                                                                                                                    \
    gemm_tc_kernel<BN, AMN, BMN><<<grid, TC_THREADS, Cfg::SMEM, s>>>(ta, tb, g);                                         \
  } while (0)
  if (form == 0) TC_LAUNCH(false, false);
  else if (form == 1) TC_LAUNCH(false, true);
  else TC_LAUNCH(true, true);
#undef TC_LAUNCH
  prof_work(2.0 * g.M * (double)g.N * g.K);
  if (g_prof_on && g_prof_shapes) {
    char nm[96];
    snprintf(nm, sizeof(nm), "%s[f%d,M=%d,N=%d,K=%d,z=%d]", g.tap.mode ? "conv_tc" : "gemm_tc", form, g.M, g.N, g.K, (int)grid.z);
    DDRL_LAUNCHED(prof_intern(nm));
    return DDRL_OK;
  }
  DDRL_LAUNCHED("gemm_tc_kernel");
  return DDRL_OK;
}

int gemm_tc(int form, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, int act, int beta, int trans_c, cudaStream_t s, const float* mask) {
  if (trans_c && bias) return DDRL_E_ARG;
  if (act >= 3 && (!mask || trans_c || beta)) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  CUtensorMap ta, tb;
  // form 0: A [M,K] K-major, B [N,K] K-major | form 1: B [K,N] MN-major | form 2: A [K,M], B [K,N] MN-major
  if (form == 2) r = tc_make_map(&ta, A, M, K, lda, 32, true); else r = tc_make_map(&ta, A, K, M, lda, TC_BM, false);
  if (r != DDRL_OK) return r;
  if (form == 0) r = tc_make_map(&tb, B, K, N, ldb, (bn + 31) / 32 * 32, false); else r = tc_make_map(&tb, B, N, K, ldb, 32, true);
  if (r != DDRL_OK) return r;
  TcArgs g;
  memset(&g, 0, sizeof(g));
  g.C = C; g.bias = bias; g.mask = mask; g.M = M; g.N = N; g.K = K; g.act = act;
  g.sCm = trans_c ? 1 : ldc; g.sCn = trans_c ? ldc : 1;
  const int kb_total = ceil_div(K, TC_BK);
  g.kb_total = kb_total;
  const int tiles = ceil_div(M, TC_BM) * ceil_div(N, bn);
  int splits = 1;
  if (act == 0 && kb_total >= 64 && tiles < kNumSMs) splits = std::min(std::min(ceil_div(2 * kNumSMs, tiles), kb_total / 16), 1024);
  int kbps = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kbps);
  g.kb_per_split = kbps;
  g.atomic = (splits > 1 || beta) ? 1 : 0;
  g.vec_store = (!g.atomic && !trans_c && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                 (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) ? 1 : 0;
  if (splits > 1 && !beta) {
    const int rows = trans_c ? N : M, cols = trans_c ? M : N;
    if (ldc == cols) DDRL_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)rows * cols, s));
    else DDRL_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * cols, rows, s));
  }
  dim3 grid(ceil_div(M, TC_BM), ceil_div(N, bn), splits);
  if (grid.y > 65535 || grid.z > 65535) return DDRL_E_UNSUPPORTED;
  switch (bn) {
    case 128: return launch_tc<128>(form, ta, tb, g, grid, s);
    case 64: return launch_tc<64>(form, ta, tb, g, grid, s);
    default: return launch_tc<32>(form, ta, tb, g, grid, s);
  }
}

// ---------------------------------------------------------------- implicit-GEMM convolutions
// 4-D map over an NHWC tensor [Bn, H, W, C]: box = 32 channels x nx pixels (every sx-th) x ny rows (every sy-th) x nb
int tc_make_map_nhwc(CUtensorMap* m, const ConvOp& o, int nx, int ny, int nb, bool mn_major) {
  cuuint64_t dims[4] = {(cuuint64_t)o.Ctot, (cuuint64_t)o.Win, (cuuint64_t)o.Hin, (cuuint64_t)o.Bn};
  cuuint64_t strides[3] = {(cuuint64_t)o.Ctot * 4, (cuuint64_t)o.Win * o.Ctot * 4, (cuuint64_t)o.Hin * o.Win * o.Ctot * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)((nx - 1) * o.sx + 1), (cuuint32_t)((ny - 1) * o.sy + 1), (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, (cuuint32_t)o.sx, (cuuint32_t)o.sy, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(o.a), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled(nhwc box %d x %d x %d) failed (%d)", nx, ny, nb, (int)r);
    return DDRL_E_CUDA;
  }
  return DDRL_OK;
}

// dy [image][pixel of the image][N] (MN-major operand): boxes of `rows` pixel rows x 32 columns; rows past the image's
// last pixel read as zero
int tc_make_map_dy3(CUtensorMap* m, const float* dy, int ldy, int N, long long ipix, int Bn, int rows) {
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)ipix, (cuuint64_t)Bn};
  cuuint64_t strides[2] = {(cuuint64_t)ldy * 4, (cuuint64_t)ipix * ldy * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dy), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled(dy box) failed (%d)", (int)cr);
    return DDRL_E_CUDA;
  }
  return DDRL_OK;
}

// dy [image][y][x][N] (MN-major operand): boxes of xw x yh pixels x 32 columns; pixels outside the image read as zero
int tc_make_map_dy4(CUtensorMap* m, const float* dy, int ldy, int N, int Xn, int Yn, int Bn, int xw, int yh) {
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)Xn, (cuuint64_t)Yn, (cuuint64_t)Bn};
  cuuint64_t strides[3] = {(cuuint64_t)ldy * 4, (cuuint64_t)Xn * ldy * 4, (cuuint64_t)Yn * Xn * ldy * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)xw, (cuuint32_t)yh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled(dy %d x %d box) failed (%d)", xw, yh, (int)cr);
    return DDRL_E_CUDA;
  }
  return DDRL_OK;
}

bool conv_tc_supported(const ConvOp& o, bool wgrad) {
  if (o.Cin % 32 != 0 || o.Ctot % 4 != 0 || o.c_off % 4 != 0 || o.c_off + o.Cin > o.Ctot) return false;
  if ((reinterpret_cast<uintptr_t>(o.a) & 15) != 0) return false;
  if (o.sx < 1 || o.sx > 8 || o.sy < 1 || o.sy > 8) return false;
  if (o.Xn < 1 || o.Yn < 1 || o.Bn < 1) return false;
  const int cap = wgrad ? 32 : TC_BM;
  if (o.Xn > cap) return false;
  if ((o.Xn - 1) * o.sx + 1 > 256) return false;
  const int ny = std::min(o.Yn, std::max(1, cap / o.Xn));
  if ((ny - 1) * o.sy + 1 > 256) return false;
  return true;
}

void tc_tap_common(TcTap& t, const ConvOp& o, int cap, bool two_phase) {
  memset(&t, 0, sizeof(t));
  t.mode = 1;
  t.Xn = o.Xn; t.Yn = o.Yn; t.Bn = o.Bn;
  if (o.Xn * o.Yn <= cap) {                     // whole images per box
    t.ny = o.Yn; t.nb = std::min(std::max(1, cap / (o.Xn * o.Yn)), 256); t.tpi = 1;
  } else {                                      // row blocks of one image
    t.ny = std::max(1, cap / o.Xn); t.nb = 1; t.tpi = ceil_div(o.Yn, t.ny);
  }
  t.tiles1 = 0x7fffffff;
  if (two_phase) {
    // tiles per image of the plan above, against: k full blocks of ny_a rows (nb_a images per box) + one box class for
    // the remaining rows
    // (taller boxes first: a plan must beat the incumbent by 5 % to replace it, so equal tile counts keep the more
    // local one)
    double best = (double)t.tpi / t.nb;
    for (int ny_a = o.Yn; ny_a >= 1; --ny_a) {
      const int nb_a = std::min(cap / (o.Xn * ny_a), 256);
      if (nb_a < 1) continue;
      if ((ny_a - 1) * o.sy + 1 > 256) continue;
      const int k = o.Yn / ny_a, rem = o.Yn - k * ny_a;
      if (rem == 0) {
        const double c = (double)k / nb_a;
        if (c < best * 0.95) { best = c; t.ny = ny_a; t.nb = nb_a; t.tpi = k; t.y2 = 0; t.ny2 = 0; t.nb2 = 0; }
        continue;
      }
      const int nb_b = std::min(cap / (o.Xn * rem), 256);
      if (nb_b < 1 || (rem - 1) * o.sy + 1 > 256) continue;
      const double c = (double)k / nb_a + 1.0 / nb_b;
      if (c < best * 0.95) { best = c; t.ny = ny_a; t.nb = nb_a; t.tpi = k; t.y2 = k * ny_a; t.ny2 = rem; t.nb2 = nb_b; }
    }
    if (t.ny2 > 0) {
      t.tiles1 = ceil_div(o.Bn, t.nb) * t.tpi;
      t.rows2 = o.Xn * t.ny2 * t.nb2;
    }
  }
  t.rows = o.Xn * t.ny * t.nb;
  t.kpad = (t.rows + 7) & ~7;
  t.KW = o.KW; t.cpb = o.Cin / 32; t.nslices = o.KH * o.KW * t.cpb;
  t.sx = o.sx; t.sy = o.sy; t.px = o.px; t.py = o.py; t.c_off = o.c_off;
}

int conv_tc_fwd(const ConvOp& o, const float* Wp, int ldw, int N, const float* bias, int act, const float* mask, float* out,
                long long osb, long long osy, long long osx, cudaStream_t s) {
  if (!conv_tc_supported(o, false) || N < 1 || ldw % 4 != 0 || (reinterpret_cast<uintptr_t>(Wp) & 15) != 0) return DDRL_E_UNSUPPORTED;
  if (act >= 3 && !mask) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  TcArgs g;
  memset(&g, 0, sizeof(g));
  tc_tap_common(g.tap, o, TC_BM);
  g.tap.osb = osb; g.tap.osy = osy; g.tap.osx = osx;
  const int K = o.KH * o.KW * o.Cin;
  CUtensorMap ta, tb;
  r = tc_make_map_nhwc(&ta, o, o.Xn, g.tap.ny, g.tap.nb, false);
  if (r != DDRL_OK) return r;
  r = tc_make_map(&tb, Wp, K, N, ldw, (bn + 31) / 32 * 32, false);
  if (r != DDRL_OK) return r;
  g.C = out; g.bias = bias; g.mask = mask; g.act = act;
  g.M = o.Bn * o.Yn * o.Xn; g.N = N; g.K = K;
  g.sCm = 0; g.sCn = 1;
  g.kb_total = g.tap.nslices; g.kb_per_split = g.kb_total;
  g.atomic = 0;
  g.vec_store = (osb % 4 == 0 && osy % 4 == 0 && osx % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) ? 1 : 0;
  dim3 grid(ceil_div(o.Bn, g.tap.nb) * g.tap.tpi, ceil_div(N, bn), 1);
  switch (bn) {
    case 128: return launch_tc<128>(0, ta, tb, g, grid, s);
    case 64: return launch_tc<64>(0, ta, tb, g, grid, s);
    default: return launch_tc<32>(0, ta, tb, g, grid, s);
  }
}

int conv_tc_wgrad(const ConvOp& o, const float* dy, int ldy, int N, float* dWp, int ldw, cudaStream_t s) {
  if (!conv_tc_supported(o, true) || N < 16 || ldy % 4 != 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0) return DDRL_E_UNSUPPORTED;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  TcArgs g;
  memset(&g, 0, sizeof(g));
  tc_tap_common(g.tap, o, 32);
  if (g.tap.nb != 1) { g.tap.nb = 1; g.tap.rows = o.Xn * g.tap.ny; g.tap.kpad = (g.tap.rows + 7) & ~7; }
  const int K = o.KH * o.KW * o.Cin;             // = M of this GEMM (the im2col K axis)
  CUtensorMap ta, tb;
  r = tc_make_map_nhwc(&ta, o, o.Xn, g.tap.ny, 1, true);
  if (r != DDRL_OK) return r;
  r = tc_make_map_dy3(&tb, dy, ldy, N, (long long)o.Yn * o.Xn, o.Bn, g.tap.rows);
  if (r != DDRL_OK) return r;
  g.C = dWp; g.bias = nullptr; g.mask = nullptr; g.act = 0;
  g.M = K; g.N = N; g.K = o.Bn * o.Yn * o.Xn;
  g.sCm = 1; g.sCn = ldw;                        // C[m = im2col k, n = cout] -> dWp[n*ldw + m]
  g.kb_total = o.Bn * g.tap.tpi;
  const int tiles = ceil_div(K, TC_BM) * ceil_div(N, bn);
  int splits = std::max(1, std::min(std::min(ceil_div(2 * kNumSMs, tiles), g.kb_total / 16), 1024));
  int kbps = ceil_div(g.kb_total, splits);
  splits = ceil_div(g.kb_total, kbps);
  g.kb_per_split = kbps;
  g.atomic = 1;                                  // accumulates into the packed gradient (zeroed per backward)
  g.vec_store = 0;
  dim3 grid(ceil_div(K, TC_BM), ceil_div(N, bn), splits);
  switch (bn) {
    case 128: return launch_tc<128>(2, ta, tb, g, grid, s);
    case 64: return launch_tc<64>(2, ta, tb, g, grid, s);
    default: return launch_tc<32>(2, ta, tb, g, grid, s);
  }
}

}  // namespace ddrl
