// Second-generation tcgen05 engine: persistent, warp-specialised, activation operand split into TENSOR MEMORY.
//
// Same arithmetic as gemm_tc.cu (fp32 in / fp32 out, 3xTF32: lo*hi + hi*lo + hi*hi, the hi*hi term accumulated in
// TMEM only over 16 MMAs and then added round-to-nearest into fp32 registers), re-organised around what the narrow
// GEMMs of this path (N = 32 / 64 output channels, K = 256..1600) are bound by: shared-memory bandwidth and per-tile
// fixed cost, not the MMA rate.
//   * the activation operand A (128 rows x 32 k) is read from shared memory ONCE per K block by the splitter warps,
//     split into hi / lo in registers and written to TMEM (tcgen05.st); all three MMA passes read A from TMEM
//     (tcgen05.mma with a TMEM A operand), so the tensor core only reads the small B tile from shared memory;
//   * the weight operand B is split once per optimiser step on the device (split_hi_lo) and arrives by TMA as two
//     ready-made tiles; nothing is rewritten in shared memory on the forward / data-gradient path;
//   * CTAs are persistent (one per SM) and loop over output tiles; dedicated epilogue warps drain the accumulators,
//     so the stores of tile t overlap the MMAs of tile t+1;
//   * weight gradient: A^T is what the MMA needs (M = im2col K axis, reduction over pixels); the transposition is free
//     because each splitter thread gathers one k column of the landed [pixels x 32 k] box into its TMEM lane.
// Warp roles: 0 TMA producer | 1 MMA issuer for the chunked accumulators | 3 MMA issuer for the whole-tile correction
// accumulator (two independent in-order MMA streams: one thread can only issue a tf32 MMA every ~45 cycles) |
// 2 TMEM allocator | splitter groups of 4 warps (TMEM lane quadrant = warp % 4; groups alternate K blocks) |
// epilogue warps (4 for BN = 32, 8 otherwise: the epilogue is ~2000 dependent instructions per 64-column tile and
// warp -- with one warp per scheduler it, not the MMA stream, bounded the N = 64 layers; profiles/r1b_ncu_tc2.md).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "layer_ops.h"
#include "prep_kernels.cuh"
#include "tc_ptx.cuh"

namespace ddrl {

constexpr int T2_BM = 128;
constexpr int T2_BK = 32;
constexpr int T2_CHUNK = 4;                   // K blocks per TMEM main-accumulator chunk (16 accumulating MMAs)

// Optional role-level wait accounting (build with -DTC2_TIMING): cycles each warp role spends blocked on each of its
// mbarriers, accumulated over every launch; read with ddrl_tc2_timing_read.  [role*4 + k], k = 3 is the role's lifetime.
// roles: 0 TMA producer {empty} | 1 chunk MMA {mfree, full, aready} | 2 corr MMA {cfree, full, aready} |
//        3 splitter warp 4 {full, afree} | 4 epilogue warp 0 {mfull, cfull, stores}
__device__ unsigned long long g_tc2_wait[32];
#ifdef TC2_TIMING
#define T2_WAIT(bar, par, acc) do { const long long _t0 = clock64(); mbar_wait(bar, par); acc += clock64() - _t0; } while (0)
#define T2_ROLE_BEGIN long long w0 = 0, w1 = 0, w2 = 0; const long long role_t0 = clock64();
#define T2_ROLE_END(role, cond) do { if ((cond) && lane == 0) { \
    atomicAdd(&g_tc2_wait[(role) * 4 + 0], (unsigned long long)w0); atomicAdd(&g_tc2_wait[(role) * 4 + 1], (unsigned long long)w1); \
    atomicAdd(&g_tc2_wait[(role) * 4 + 2], (unsigned long long)w2); \
    atomicAdd(&g_tc2_wait[(role) * 4 + 3], (unsigned long long)(clock64() - role_t0)); } } while (0)
#else
#define T2_WAIT(bar, par, acc) mbar_wait(bar, par)
#define T2_ROLE_BEGIN
#define T2_ROLE_END(role, cond)
#endif

struct Tc2Args {
  float* C;
  const float* bias;
  const float* mask;
  long long sCm, sCn;
  int M, N, K;
  int kb_total, kb_per_split;
  int act, atomic, vec_store;
  int m_tiles, n_tiles;                       // mode 0: persistent tile space
  TcTap tap;
};

template <int BN>
struct T2Cfg {
  static constexpr int BNS = (BN + 31) / 32 * 32;
  static constexpr int A_BYTES = T2_BM * T2_BK * 4;              // 16 KB raw activation tile (or 4 transposed slices)
  static constexpr int B_BYTES = BNS * T2_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;      // A raw | B hi | B lo
  static constexpr int STAGES = BN <= 32 ? 6 : (BN <= 64 ? 5 : 4);
  // BN <= 64: hi*hi and hi*lo are ONE MMA of N' = 2 BN (B hi | B lo tiles are adjacent in shared memory) writing the
  // adjacent accumulators [main | corrB] of the current chunk buffer; lo*hi goes to corrA.  2 MMAs per k step instead
  // of 3 (every tf32 MMA with N <= 64 occupies the tensor pipe for ~45 cycles regardless of N, scratch/mma_bench.cu).
  // (Measured alternatives -- un-folded accumulators with more A slots, one MMA stream, a single chunk accumulator for
  // BN = 128 -- are recorded in profiles/r1d_tc2_skeleton.md / r1c_tc2_experiments.txt; none was adopted.)
  static constexpr bool FOLD = BN <= 64;
  static constexpr int SA = BN <= 32 ? 4 : (BN <= 64 ? 3 : 2);   // TMEM A slots (64 columns each: hi | lo)
  static constexpr int NEPI = BN <= 32 ? 4 : 8;                  // 32 accumulator columns per epilogue thread (64 for BN = 128)
  static constexpr int NSG = BN <= 64 ? 2 : 1;                   // splitter groups (4 warps each), K blocks round-robin
  static constexpr int EPI0 = 4 + 4 * NSG;                       // first epilogue warp
  static constexpr int THREADS = (EPI0 + NEPI) * 32;
  static constexpr int COLS = BN / (NEPI / 4);                   // accumulator columns per epilogue thread
  // FOLD: [main0 | corrB0 | main1 | corrB1 | corrA | A slots]; else [main0 | main1 | corr | A slots]
  static constexpr int TM_MAIN0 = 0, TM_MAIN1 = FOLD ? 2 * BN : BN, TM_CORR = FOLD ? 4 * BN : 2 * BN, TM_A = FOLD ? 5 * BN : 3 * BN;
  static constexpr int TMEM_COLS = 512;
  static constexpr int NBARS = 2 * STAGES + SA + 6;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES + 256;      // epilogue staging: one swizzled 32 x 32 fp32 panel per warp
  static constexpr int SMEM = 1024 + STG_OFF + NEPI * 4096;
  static_assert(NBARS * 8 + 16 <= 256, "barrier block");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(TM_A + SA * 64 <= 512, "TMEM budget");
  static_assert(STAGES > SA, "the stage barrier of K block it - SA releases TMEM slot it % SA");
};

// MODE 0: C[M,N] = epi(A[M,K] . B)   A K-major (plain 2-D or tap boxes), B = pre-split weights (K-major or MN-major)
// MODE 1: C[m,n] += sum_r A[r,m] * B[r,n]   (weight gradient; A = activations [rows, k], B = dy [rows, n], both raw)
template <int BN, int MODE, bool B_MN>
__global__ void __launch_bounds__(T2Cfg<BN>::THREADS, 1)
tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
           const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmA2, Tc2Args g) {
  using Cfg = T2Cfg<BN>;
  constexpr int S = Cfg::STAGES, SA = Cfg::SA;
  extern __shared__ uint8_t smem_dyn[];
  // 1024-byte alignment by POINTER arithmetic on the __shared__ array: an integer round-trip hides the address space from
  // the compiler, which then emits generic LD / ST (long-scoreboard, L1TEX path) for every shared-memory access below
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* bar_full = bars;                  // [S]  TMA landed
  uint64_t* bar_empty = bars + S;             // [S]  MMAs that read the stage retired
  uint64_t* bar_aready = bars + 2 * S;        // [SA] TMEM A slot written
  uint64_t* bar_mfull = bars + 2 * S + SA;          // [2] main accumulator chunk complete
  uint64_t* bar_mfree = bar_mfull + 2;              // [2] drained
  uint64_t* bar_cfull = bar_mfull + 4;              // correction accumulator complete (tile end)
  uint64_t* bar_cfree = bar_mfull + 5;              // read by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcTap& tp = g.tap;
  const bool tapA = tp.mode != 0;

  if (MODE == 1 && tapA) {
    // pixel-box K blocks shorter than 32 rows: the rows no box covers must read as zero
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < S * Cfg::STAGE_BYTES / 16; i += Cfg::THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  if (warp == 0 && lane == 0) {
    // stages and A slots are released by BOTH MMA issuers' commits
    for (int s = 0; s < S; ++s) { mbar_init(smem_u32(bar_full + s), 1); mbar_init(smem_u32(bar_empty + s), 2); }
    for (int a = 0; a < SA; ++a) mbar_init(smem_u32(bar_aready + a), 4);
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(bar_mfull + b), 1); mbar_init(smem_u32(bar_mfree + b), Cfg::NEPI); }
    mbar_init(smem_u32(bar_cfull), 1);
    mbar_init(smem_u32(bar_cfree), Cfg::NEPI);
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

  // ---- work enumeration, identical in every role --------------------------------------------------------------
  // MODE 0: tiles blockIdx.x, +gridDim.x, ... of the (m_tiles x n_tiles) space, full K each
  // MODE 1: one unit per CTA: (blockIdx.x = M block, blockIdx.y = N tile, blockIdx.z = K split)
  const int total_tiles = MODE == 0 ? g.m_tiles * g.n_tiles : 1;
  const int tile_step = MODE == 0 ? (int)gridDim.x : 1;
  const int tile_first = MODE == 0 ? (int)blockIdx.x : 0;
  int kb0 = 0, nkb = g.kb_total;
  if (MODE == 1) {
    kb0 = blockIdx.z * g.kb_per_split;
    nkb = min(g.kb_total, kb0 + g.kb_per_split) - kb0;
  }
  const int ksteps = (MODE == 1 && tapA) ? tp.kpad / 8 : T2_BK / 8;

  if (warp == 0) {
    // ============================================================ TMA producer
    // The whole warp runs the loop convergently; one elected lane issues (keeps every operand in uniform registers).
    // Tap / pixel-block coordinates advance incrementally: no integer division per K block.
    uint32_t it = 0;
    T2_ROLE_BEGIN
    // weight gradient, implicit operand: the 4 (tap, chunk) slices of this CTA's M block are fixed
    int sl_c[4] = {0, 0, 0, 0}, sl_x[4] = {0, 0, 0, 0}, sl_y[4] = {0, 0, 0, 0}, na = 0;
    if (MODE == 1 && tapA) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sl = (int)blockIdx.x * 4 + j;
        if (sl < tp.nslices) {
          const int tap = sl / tp.cpb, cc = sl - tap * tp.cpb;
          const int kh = tap / tp.KW, kw = tap - kh * tp.KW;
          sl_c[j] = tp.c_off + cc * 32; sl_x[j] = kw - tp.px; sl_y[j] = kh - tp.py;
          na = j + 1;
        }
      }
    }
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int mt = MODE == 0 ? tile / g.n_tiles : (int)blockIdx.x;
      const int nt = MODE == 0 ? tile - mt * g.n_tiles : (int)blockIdx.y;
      const int m0 = mt * T2_BM, n0 = nt * BN;
      int b0 = 0, y0 = 0;
      const bool ph2 = MODE == 0 && tapA && mt >= tp.tiles1;       // second tiling phase: the images' remaining rows
      if (MODE == 0 && tapA) {
        if (!ph2) { b0 = (mt / tp.tpi) * tp.nb; y0 = (mt % tp.tpi) * tp.ny * tp.sy - tp.py; }
        else { b0 = (mt - tp.tiles1) * tp.nb2; y0 = tp.y2 * tp.sy - tp.py; }
      }
      int cc = 0, kw = 0, kh = 0;                     // forward taps: K block = (kh, kw, chunk), chunk fastest
      int pb = 0, pj = 0;                             // wgrad pixel blocks: image, row block within the image
      if (MODE == 1 && tapA) { pb = kb0 / tp.tpi; pj = kb0 - pb * tp.tpi; }
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S;
        T2_WAIT(smem_u32(bar_empty + s), ((it / S) & 1) ^ 1, w0);
        if (elect_one()) {
          const uint32_t full = smem_u32(bar_full + s);
          const uint32_t a_dst = smem_u32(smem) + s * Cfg::STAGE_BYTES, bh_dst = a_dst + Cfg::A_BYTES,
                         bl_dst = bh_dst + Cfg::B_BYTES;
          const int k = (kb0 + i) * T2_BK;
          if (MODE == 0) {
            const uint32_t bbytes = 2u * Cfg::B_BYTES;
            if (!tapA) {
              mbar_expect_tx(full, Cfg::A_BYTES + bbytes);
              tma_load_2d(&tmA, full, a_dst, k, m0);
            } else {
              mbar_expect_tx(full, (ph2 ? tp.rows2 : tp.rows) * 128 + bbytes);
              tma_load_4d(ph2 ? &tmA2 : &tmA, full, a_dst, tp.c_off + cc * 32, kw - tp.px, y0 + kh, b0);
            }
            if (!B_MN) {
              tma_load_2d(&tmBhi, full, bh_dst, k, n0);
              tma_load_2d(&tmBlo, full, bl_dst, k, n0);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::BNS / 32; ++j) {
                tma_load_2d(&tmBhi, full, bh_dst + j * 4096, n0 + j * 32, k);
                tma_load_2d(&tmBlo, full, bl_dst + j * 4096, n0 + j * 32, k);
              }
            }
          } else if (!tapA) {
            // A slices: [32 rows x 32 k] boxes of x[rows, Kx] at k = m0 + 32 j; B: dy[rows, N] boxes of 32 rows x 32 n
            mbar_expect_tx(full, Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(&tmA, full, a_dst + j * 4096, m0 + j * 32, k);
#pragma unroll
            for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_2d(&tmBhi, full, bh_dst + j * 4096, n0 + j * 32, k);
          } else if (tp.w2on) {
            // two pixel-box classes (32 pixels each): class 1 covers the columns right of wxw0
            const bool c1 = pj >= tp.wnb0;
            const int x0 = c1 ? tp.wxw0 : 0, yy0 = c1 ? (pj - tp.wnb0) * tp.wyh1 : pj * tp.wyh0;
            mbar_expect_tx(full, (na + Cfg::BNS / 32) * 32 * 128);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < na) tma_load_4d(c1 ? &tmA2 : &tmA, full, a_dst + j * 4096, sl_c[j], x0 * tp.sx + sl_x[j], yy0 * tp.sy + sl_y[j], pb);
#pragma unroll
            for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_4d(c1 ? &tmBlo : &tmBhi, full, bh_dst + j * 4096, n0 + j * 32, x0, yy0, pb);
          } else {
            const int yy0 = pj * tp.ny;
            mbar_expect_tx(full, (na + Cfg::BNS / 32) * tp.rows * 128);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < na) tma_load_4d(&tmA, full, a_dst + j * 4096, sl_c[j], sl_x[j], yy0 * tp.sy + sl_y[j], pb);
#pragma unroll
            for (int j = 0; j < Cfg::BNS / 32; ++j) tma_load_3d(&tmBhi, full, bh_dst + j * 4096, n0 + j * 32, yy0 * tp.Xn, pb);
          }
        }
        __syncwarp();
        if (MODE == 0) { if (++cc == tp.cpb) { cc = 0; if (++kw == tp.KW) { kw = 0; ++kh; } } }
        else if (++pj == tp.tpi) { pj = 0; ++pb; }
      }
    }
    T2_ROLE_END(0, true);
  } else if (warp == 1 || warp == 3) {
    // ============================================================ MMA issuers (one elected thread each)
    // warp 1: chunk buffers   main (+)= A_hi . B_hi          [FOLD: [main | corrB] (+)= A_hi . [B_hi ; B_lo], N' = 2 BN]
    // warp 3: whole tile      corr  += A_lo . B_hi           [!FOLD: ... + A_hi . B_lo]
    // The two streams write disjoint accumulators, so they need no ordering between them.
    const bool chunk_role = warp == 1;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                            ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t t_corr = tmem_base + Cfg::TM_CORR;
    constexpr uint32_t kstep = B_MN ? (1024 >> 4) : (32 >> 4);      // start-address field increment per k step
    uint32_t it = 0, ch = 0, tl = 0;
    T2_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tl) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S, a = it % SA;
        const uint32_t buf = ch & 1;
        const bool first_in_chunk = (i % T2_CHUNK) == 0;
        const bool last_in_chunk = (i % T2_CHUNK) == T2_CHUNK - 1 || i == nkb - 1;
        if (chunk_role) {
          if (first_in_chunk) {
            T2_WAIT(smem_u32(bar_mfree + buf), ((ch >> 1) & 1) ^ 1, w0);
          }
        }
        else if (i == 0) T2_WAIT(smem_u32(bar_cfree), (tl & 1) ^ 1, w0);
        T2_WAIT(smem_u32(bar_full + s), (it / S) & 1, w1);
        T2_WAIT(smem_u32(bar_aready + a), (it / SA) & 1, w2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
          const uint32_t a_hi = tmem_base + Cfg::TM_A + a * 64, a_lo = a_hi + 32;
          const uint64_t dbh0 = B_MN ? umma_desc(b_hi, 4096, 512, 1) : umma_desc(b_hi, 16, 1024, 2);
          if (chunk_role) {
            const uint32_t t_main = tmem_base + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0);
#pragma unroll
            for (int k4 = 0; k4 < T2_BK / 8; ++k4) {
              if (k4 >= ksteps) break;
              umma_tf32_ts(t_main, a_hi + k4 * 8, dbh0 + k4 * kstep, Cfg::FOLD ? idesc2 : idesc,
                           (!first_in_chunk || k4 != 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(bar_empty + s));            // retires the stage AND the TMEM A slot of this K block
            if (last_in_chunk) umma_commit(smem_u32(bar_mfull + buf));
          } else {
            const uint64_t dbl0 = B_MN ? umma_desc(b_lo, 4096, 512, 1) : umma_desc(b_lo, 16, 1024, 2);
#pragma unroll
            for (int k4 = 0; k4 < T2_BK / 8; ++k4) {
              if (k4 >= ksteps) break;
              umma_tf32_ts(t_corr, a_lo + k4 * 8, dbh0 + k4 * kstep, idesc, (i | k4) != 0 ? 1u : 0u);
              if (!Cfg::FOLD) umma_tf32_ts(t_corr, a_hi + k4 * 8, dbl0 + k4 * kstep, idesc, 1u);
            }
            umma_commit(smem_u32(bar_empty + s));
            if (i == nkb - 1) umma_commit(smem_u32(bar_cfull));
          }
        }
        __syncwarp();
        if (last_in_chunk) ++ch;
      }
    }
    T2_ROLE_END(chunk_role ? 1 : 2, true);
  } else if (warp >= 4 && warp < Cfg::EPI0) {
    // ============================================================ splitters: smem A -> hi / lo -> TMEM
    const int q = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const int st_tid = (threadIdx.x - 128) & 127;                // thread index within the group
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t it = 0;
    T2_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % S, a = it % SA;
        if ((int)(it % Cfg::NSG) != grp) {                         // the groups take K blocks round-robin
          // every group follows every phase of the stage barrier: a parity wait that skips a phase (S not a multiple of
          // NSG) is satisfied by the phase before the skipped one (see tc3.cu)
          if (S % Cfg::NSG != 0) T2_WAIT(smem_u32(bar_full + s), (it / S) & 1, w0);
          continue;
        }
        T2_WAIT(smem_u32(bar_full + s), (it / S) & 1, w0);
        const uint8_t* st = smem + s * Cfg::STAGE_BYTES;
#ifdef TC2_TIMING
        const long long sp0 = clock64();
#endif
        if (MODE == 1) {
          // raw dy tile: hi in place + lo twin; element-wise, so the TMA swizzle is preserved
          uint8_t* bh = const_cast<uint8_t*>(st) + Cfg::A_BYTES;
          for (int v = st_tid; v < Cfg::B_BYTES / 16; v += 128) {
            const float4 x = *reinterpret_cast<const float4*>(bh + v * 16);
            uint4 h, l;
            split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
            *reinterpret_cast<uint4*>(bh + v * 16) = h;
            *reinterpret_cast<uint4*>(bh + Cfg::B_BYTES + v * 16) = l;
          }
          fence_async_smem();
        }
        const uint32_t ta = tmem_base + t_lane + Cfg::TM_A + a * 64;
        // two halves of 16 k values (keeps the live split registers at 32: the CTA runs 640 threads for BN = 64)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t hi[16], lo[16];
          if (MODE == 0) {
            // thread = tile row; its 32 k values are the row's eight 16-byte chunks (128B swizzle: chunk ^ (row & 7))
            const int row = q * 32 + lane;
            const uint8_t* rp = st + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(rp + (((hf * 4 + c) ^ (row & 7)) << 4));
              const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) split_tf32(x[e], hi[c * 4 + e], lo[c * 4 + e]);
            }
          } else {
            // thread = k column `lane` of slice q; it gathers that column over the (<= 32) pixel rows of the box
            const uint8_t* sp = st + q * 4096 + (lane & 3) * 4;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const int pp = hf * 16 + p;
              const float x = *reinterpret_cast<const float*>(sp + pp * 128 + (((lane >> 2) ^ (pp & 7)) << 4));
              split_tf32(x, hi[p], lo[p]);
            }
          }
          if (hf == 0) {
#ifdef TC2_TIMING
            w2 += clock64() - sp0;
#endif
            // TMEM slot a was last read by K block it - SA: its stage barrier (both MMA streams commit to it) doubles as
            // the slot's release -- S > SA, so that barrier cannot be a second phase ahead when we look at it
            if (it >= (uint32_t)SA) {
              const uint32_t j = it - SA;
              T2_WAIT(smem_u32(bar_empty + (j % S)), (j / S) & 1, w1);
            }
            tc_fence_after();
          }
          tmem_st16(ta + hf * 16, hi);
          tmem_st16(ta + 32 + hf * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_aready + a));
      }
    }
    T2_ROLE_END(3, warp == 4);
  } else if (warp >= Cfg::EPI0) {
    // ============================================================ drain + epilogue
    const int e = warp - Cfg::EPI0;
    const int q = e & 3, half = e >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const uint32_t col0 = half * Cfg::COLS;
    uint32_t ch = 0, tl = 0;
    T2_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tl) {
      const int mt = MODE == 0 ? tile / g.n_tiles : (int)blockIdx.x;
      const int nt = MODE == 0 ? tile - mt * g.n_tiles : (int)blockIdx.y;
      const int m0 = mt * T2_BM, n0 = nt * BN;
      float acc[Cfg::COLS];
#pragma unroll
      for (int j = 0; j < Cfg::COLS; ++j) acc[j] = 0.f;
      const int nch = (nkb + T2_CHUNK - 1) / T2_CHUNK;
      for (int c = 0; c < nch; ++c, ++ch) {
        const int buf = ch & 1;
        T2_WAIT(smem_u32(bar_mfull + buf), (ch >> 1) & 1, w0);
        tc_fence_after();
#pragma unroll
        for (int j0 = 0; j0 < Cfg::COLS; j0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + col0 + j0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j0 + j] += v[j];
          if (Cfg::FOLD) {
            // the hi*lo term of this chunk sits BN columns further (second half of the folded N' = 2 BN MMA)
            tmem_ld32(tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + BN + col0 + j0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j0 + j] += v[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_mfree + buf));
      }
      {
        T2_WAIT(smem_u32(bar_cfull), tl & 1, w1);
        tc_fence_after();
#pragma unroll
        for (int j0 = 0; j0 < Cfg::COLS; j0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + t_lane + Cfg::TM_CORR + col0 + j0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j0 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_cfree));
      }
      // ---- stores
#ifdef TC2_TIMING
      const long long st0 = clock64();
#endif
      const int r = q * 32 + lane;
      bool rvalid;
      long long roff;
      if (MODE == 0 && tapA) {
        const bool ph2 = mt >= tp.tiles1;
        const int b0 = ph2 ? (mt - tp.tiles1) * tp.nb2 : (mt / tp.tpi) * tp.nb, y0 = ph2 ? tp.y2 : (mt % tp.tpi) * tp.ny;
        const int nyp = ph2 ? tp.ny2 : tp.ny;
        const int x = r % tp.Xn, t2 = r / tp.Xn;
        const int yy = t2 % nyp, bb = t2 / nyp;
        rvalid = r < (ph2 ? tp.rows2 : tp.rows) && (b0 + bb) < tp.Bn && (y0 + yy) < tp.Yn;
        roff = (long long)(b0 + bb) * tp.osb + (long long)(y0 + yy) * tp.osy + (long long)x * tp.osx;
      } else {
        rvalid = (m0 + r) < g.M;
        roff = (long long)(m0 + r) * g.sCm;
      }
      const float neg_slope = g.act == 3 ? 0.f : 0.01f;
      // fused parity classes: tile pixel (py, px) -> input pixel (out_s*py + cls_iy, out_s*px + cls_ix) per column group
      const bool fused = MODE == 0 && tapA && tp.ncls > 1;
      int py = 0, px = 0;
      if (fused) {
        const bool ph2 = mt >= tp.tiles1;
        const int y0 = ph2 ? tp.y2 : (mt % tp.tpi) * tp.ny;
        px = (r % tp.Xn) * tp.out_s;
        py = (y0 + (r / tp.Xn) % (ph2 ? tp.ny2 : tp.ny)) * tp.out_s;
      }
      if (MODE == 0 && g.vec_store) {
        // Coalesced path: the warp's 32 rows x 32 columns go through a swizzled 4 KB staging panel, then every store
        // (and activation-mask load) instruction covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
        float* stg = reinterpret_cast<float*>(smem + Cfg::STG_OFF) + e * 1024;
        const int cidx = lane & 7, rsub = lane >> 3;
        const int pyx = (py << 16) | px;
#pragma unroll
        for (int p0 = 0; p0 < Cfg::COLS; p0 += 32) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) =
                make_float4(acc[p0 + 4 * c], acc[p0 + 4 * c + 1], acc[p0 + 4 * c + 2], acc[p0 + 4 * c + 3]);
          __syncwarp();
          const int colv = n0 + col0 + p0 + cidx * 4;
          const bool cok = colv < g.N;
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (g.bias != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (colv + k < g.N) bv[k] = g.bias[colv + k];
          }
          long long cadd = 0;
          int ciy = 0, cix = 0;
          if (fused && cok) {
            const int qq = colv / tp.cls_cols;
            ciy = tp.cls_iy[qq]; cix = tp.cls_ix[qq];
            cadd = tp.cls_off[qq] - (long long)qq * tp.cls_cols;     // column colv of group qq lands at channel colv - qq*cls_cols
          }
          // pass 1: addresses + all activation-mask loads of the panel in flight together (the mask may alias nothing
          // the compiler can prove, so loads interleaved with the stores would serialise on L2 latency)
          long long off[8];
          float4 mk[8];
          uint32_t okm = 0;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int rr = i8 * 4 + rsub;
            const int ok = __shfl_sync(0xffffffffu, (int)rvalid, rr);
            const long long ro = __shfl_sync(0xffffffffu, roff, rr);
            const int ryx = __shfl_sync(0xffffffffu, pyx, rr);
            bool live = ok && cok;
            if (fused && ((ryx >> 16) + ciy >= tp.out_H || (ryx & 0xffff) + cix >= tp.out_W)) live = false;
            off[i8] = ro + cadd + colv;
            mk[i8] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (live) {
              okm |= 1u << i8;
              if (g.act >= 3) {
                if (colv + 4 <= g.N) mk[i8] = __ldg(reinterpret_cast<const float4*>(g.mask + off[i8]));
                else {
                  float* mv = reinterpret_cast<float*>(&mk[i8]);
#pragma unroll
                  for (int k = 0; k < 4; ++k) if (colv + k < g.N) mv[k] = __ldg(g.mask + off[i8] + k);
                }
              }
            }
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            if (!((okm >> i8) & 1u)) continue;
            const int rr = i8 * 4 + rsub;
            const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cidx ^ (rr & 7)) << 2));
            const float xv[4] = {v4.x, v4.y, v4.z, v4.w};
            const float* mv = reinterpret_cast<const float*>(&mk[i8]);
            float4 o;
            float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x = xv[k] + bv[k];
              if (g.act == 1) x = fmaxf(x, 0.f);
              else if (g.act == 2) x = x > 0.f ? x : 0.01f * x;
              else if (g.act >= 3) x = mv[k] > 0.f ? x : neg_slope * x;
              ov[k] = x;
            }
            if (colv + 4 <= g.N) {
              *reinterpret_cast<float4*>(g.C + off[i8]) = o;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) if (colv + k < g.N) g.C[off[i8] + k] = ov[k];
            }
          }
          __syncwarp();
        }
      } else if (rvalid) {
#pragma unroll
        for (int j0 = 0; j0 < Cfg::COLS; j0 += 4) {
          const int colv = n0 + col0 + j0;
          long long coff = roff;
          if (fused) {
            if (colv >= g.N) continue;
            const int qq = colv / tp.cls_cols;
            if (py + tp.cls_iy[qq] >= tp.out_H || px + tp.cls_ix[qq] >= tp.out_W) continue;
            coff += tp.cls_off[qq] - (long long)qq * tp.cls_cols;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int col = colv + k;
            if (col < g.N) {
              float x = acc[j0 + k];
              float* p = g.C + coff + col * g.sCn;
              if (g.atomic) {
                atomicAdd(p, x);
              } else {
                if (g.bias != nullptr) x += g.bias[col];
                if (g.act == 1) x = fmaxf(x, 0.f);
                else if (g.act == 2) x = x > 0.f ? x : 0.01f * x;
                else if (g.act >= 3) x = g.mask[coff + col * g.sCn] > 0.f ? x : neg_slope * x;
                *p = x;
              }
            }
          }
        }
      }
#ifdef TC2_TIMING
      w2 += clock64() - st0;
#endif
    }
    T2_ROLE_END(4, warp == Cfg::EPI0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
template <int BN, int MODE, bool B_MN>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, const Tc2Args& g, dim3 grid,
                   cudaStream_t s, const CUtensorMap* ta2 = nullptr) {
  using Cfg = T2Cfg<BN>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_index()];
  if (!attr_done) {
    DDRL_CUDA(cudaFuncSetAttribute(tc2_kernel<BN, MODE, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  tc2_kernel<BN, MODE, B_MN><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(ta, tbh, tbl, ta2 ? *ta2 : ta, g);
  prof_work(2.0 * g.M * (double)g.N * g.K * (g.tap.work_scale > 0.f ? g.tap.work_scale : 1.f));   // algorithmic flops
  if (g_prof_on && g_prof_shapes) {
    char nm[96];
    snprintf(nm, sizeof(nm), "%s2[%s,M=%d,N=%d,K=%d,g=%d]", g.tap.mode ? "conv_tc" : "gemm_tc",
             MODE ? "wgrad" : (B_MN ? "dgrad" : "fwd"), g.M, g.N, g.K, (int)(grid.x * grid.y * grid.z));
    DDRL_LAUNCHED(prof_intern(nm));
    return DDRL_OK;
  }
  DDRL_LAUNCHED("tc2_kernel");
  return DDRL_OK;
}

template <int MODE, bool B_MN>
static int launch2_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, const Tc2Args& g, dim3 grid,
                      cudaStream_t s, const CUtensorMap* ta2 = nullptr) {
  switch (bn) {
    case 128: return launch2<128, MODE, B_MN>(ta, tbh, tbl, g, grid, s, ta2);
    case 64: return launch2<64, MODE, B_MN>(ta, tbh, tbl, g, grid, s, ta2);
    default: return launch2<32, MODE, B_MN>(ta, tbh, tbl, g, grid, s, ta2);
  }
}

static inline int pick_bn(int N) { return N > 64 ? 128 : (N > 32 ? 64 : 32); }
// persistent grid of the forward / data-gradient launches (DDRL_TC2_GRID: experiments on L2 vs per-SM ingress limits)
static inline int persistent_ctas() {
  static const int n = [] { const char* e = getenv("DDRL_TC2_GRID"); const int v = e ? atoi(e) : 0; return v > 0 && v < kNumSMs ? v : kNumSMs; }();
  return n;
}
// Implicit convolutions with a short K loop are bounded by the epilogue (one 64-column pass per warp and tile), not by
// the MMA stream: DDRL_TC2_CONV_MAXBN=64 runs their N = 128 layers as two 64-wide tiles (twice the epilogue warps per
// output column, FOLD MMAs) at the price of splitting the activation tile twice.
static inline int pick_bn_conv(int N, int kb_total) {
  static const int maxbn = [] { const char* e = getenv("DDRL_TC2_CONV_MAXBN"); return e ? atoi(e) : 128; }();
  static const int maxkb = [] { const char* e = getenv("DDRL_TC2_CONV_MAXBN_KB"); return e ? atoi(e) : 1 << 30; }();
  int bn = pick_bn(N);
  if (bn > maxbn && kb_total <= maxkb && maxbn >= 32) bn = maxbn;
  return bn;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool tc2_gemm_supported(int form, int M, int N, int K, const float* A, int lda, const float* Bhi, const float* Blo, int ldb) {
  if (form != 0 && form != 1) return false;
  if (M < 1 || N < 1 || K < 1) return false;
  if (!al16(A) || !al16(Bhi) || !al16(Blo) || lda % 4 != 0 || ldb % 4 != 0) return false;
  return true;
}

// forward / data-gradient GEMM with pre-split weights.  form 0: B [N,K] K-major; form 1: B [K,N] MN-major.
int tc2_gemm(int form, int M, int N, int K, const float* A, int lda, const float* Bhi, const float* Blo, int ldb, float* C,
             int ldc, const float* bias, int act, const float* mask, cudaStream_t s) {
  if (!tc2_gemm_supported(form, M, N, K, A, lda, Bhi, Blo, ldb)) return DDRL_E_UNSUPPORTED;
  if (act >= 3 && !mask) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn(N);
  const int bns = (bn + 31) / 32 * 32;
  CUtensorMap ta, tbh, tbl;
  r = tc_make_map(&ta, A, K, M, lda, T2_BM, false);
  if (r != DDRL_OK) return r;
  if (form == 0) { r = tc_make_map(&tbh, Bhi, K, N, ldb, bns, false); if (r == DDRL_OK) r = tc_make_map(&tbl, Blo, K, N, ldb, bns, false); }
  else { r = tc_make_map(&tbh, Bhi, N, K, ldb, 32, true); if (r == DDRL_OK) r = tc_make_map(&tbl, Blo, N, K, ldb, 32, true); }
  if (r != DDRL_OK) return r;
  Tc2Args g;
  memset(&g, 0, sizeof(g));
  g.C = C; g.bias = bias; g.mask = mask; g.act = act;
  g.M = M; g.N = N; g.K = K; g.sCm = ldc; g.sCn = 1;
  g.kb_total = ceil_div(K, T2_BK); g.kb_per_split = g.kb_total;
  g.m_tiles = ceil_div(M, T2_BM); g.n_tiles = ceil_div(N, bn);
  g.vec_store = (ldc % 4 == 0 && al16(C) && (!mask || al16(mask))) ? 1 : 0;
  dim3 grid(std::min(g.m_tiles * g.n_tiles, persistent_ctas()), 1, 1);
  return form == 0 ? launch2_bn<0, false>(bn, ta, tbh, tbl, g, grid, s) : launch2_bn<0, true>(bn, ta, tbh, tbl, g, grid, s);
}

int tc2_conv_fwd(const ConvOp& o, const float* Whi, const float* Wlo, int ldw, int N, const float* bias, int act,
                 const float* mask, float* out, long long osb, long long osy, long long osx, cudaStream_t s,
                 const TcTap* cls) {
  if (!conv_tc_supported(o, false) || N < 1 || ldw % 4 != 0 || !al16(Whi) || !al16(Wlo)) return DDRL_E_UNSUPPORTED;
  if (act >= 3 && !mask) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn_conv(N, o.KH * o.KW * (o.Cin / 32));
  const int bns = (bn + 31) / 32 * 32;
  Tc2Args g;
  memset(&g, 0, sizeof(g));
  static const bool one_phase = [] { const char* e = getenv("DDRL_TC2_ONE_PHASE"); return e && e[0] == '1'; }();
  tc_tap_common(g.tap, o, T2_BM, !one_phase);
  g.tap.osb = osb; g.tap.osy = osy; g.tap.osx = osx;
  if (cls && cls->ncls > 1) {
    if (cls->ncls > 4 || N != cls->ncls * cls->cls_cols || cls->cls_cols % 4 != 0) return DDRL_E_ARG;
    g.tap.ncls = cls->ncls; g.tap.cls_cols = cls->cls_cols; g.tap.out_s = cls->out_s; g.tap.out_H = cls->out_H;
    g.tap.out_W = cls->out_W; g.tap.work_scale = cls->work_scale;
    for (int q = 0; q < cls->ncls; ++q) {
      g.tap.cls_iy[q] = cls->cls_iy[q]; g.tap.cls_ix[q] = cls->cls_ix[q]; g.tap.cls_off[q] = cls->cls_off[q];
      if (cls->cls_off[q] % 4 != 0) return DDRL_E_ARG;
    }
  }
  const int K = o.KH * o.KW * o.Cin;
  CUtensorMap ta, tbh, tbl, ta2;
  r = tc_make_map_nhwc(&ta, o, o.Xn, g.tap.ny, g.tap.nb, false);
  const bool ph2 = g.tap.ny2 > 0;
  if (r == DDRL_OK && ph2) r = tc_make_map_nhwc(&ta2, o, o.Xn, g.tap.ny2, g.tap.nb2, false);
  if (r == DDRL_OK) r = tc_make_map(&tbh, Whi, K, N, ldw, bns, false);
  if (r == DDRL_OK) r = tc_make_map(&tbl, Wlo, K, N, ldw, bns, false);
  if (r != DDRL_OK) return r;
  g.C = out; g.bias = bias; g.mask = mask; g.act = act;
  g.M = o.Bn * o.Yn * o.Xn; g.N = N; g.K = K; g.sCm = 0; g.sCn = 1;
  g.kb_total = g.tap.nslices; g.kb_per_split = g.kb_total;
  g.m_tiles = ceil_div(o.Bn, g.tap.nb) * g.tap.tpi + (ph2 ? ceil_div(o.Bn, g.tap.nb2) : 0);
  g.n_tiles = ceil_div(N, bn);
  g.vec_store = (osb % 4 == 0 && osy % 4 == 0 && osx % 4 == 0 && al16(out) && (!mask || al16(mask))) ? 1 : 0;
  dim3 grid(std::min(g.m_tiles * g.n_tiles, persistent_ctas()), 1, 1);
  return launch2_bn<0, false>(bn, ta, tbh, tbl, g, grid, s, ph2 ? &ta2 : nullptr);
}

// K splits of the weight gradient: one CTA per (tile, split) and one CTA per SM at a time, so the launch runs in
// ceil(tiles*splits / 148) waves of equal-length CTAs.  Pick the split count whose last wave is fullest (a grid of 300
// CTAs is THREE waves: 148 + 148 + 4), preferring fewer splits (less atomic traffic) among near-equals; every split keeps
// >= 8 K blocks so the pipeline prologue stays amortised.
static void wgrad_splits(Tc2Args& g, int tiles) {
  const int max_splits = std::max(1, std::min(g.kb_total / 8, 1024));
  int best = 1;
  double best_cost = 1e300;
  for (int sp = 1; sp <= max_splits && (long long)tiles * sp <= 4LL * kNumSMs; ++sp) {
    const int kbps = ceil_div(g.kb_total, sp);
    const int waves = ceil_div(tiles * ceil_div(g.kb_total, kbps), kNumSMs);
    const double cost = (double)waves * (kbps + 6);                 // K blocks per CTA + fixed prologue/epilogue
    if (cost < best_cost * 0.98) { best_cost = cost; best = sp; }
  }
  g.kb_per_split = ceil_div(g.kb_total, best);
}

// dW[n*ldw + k] += sum_r dy[r, n] * x[r, k]     (x [rows, Kx], dy [rows, N]; accumulates atomically)
int tc2_wgrad(int Kx, int N, long long rows, const float* x, int ldx, const float* dy, int ldy, float* dW, int ldw,
              cudaStream_t s) {
  if (Kx < 1 || N < 1 || rows < 1 || !al16(x) || !al16(dy) || ldx % 4 != 0 || ldy % 4 != 0 || rows > 0x7fffffffLL)
    return DDRL_E_UNSUPPORTED;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  r = tc_make_map(&ta, x, Kx, rows, ldx, 32, false);             // [32 rows x 32 k] boxes, plain 128B swizzle
  if (r == DDRL_OK) r = tc_make_map(&tb, dy, N, rows, ldy, 32, true);
  if (r != DDRL_OK) return r;
  Tc2Args g;
  memset(&g, 0, sizeof(g));
  g.C = dW; g.M = Kx; g.N = N; g.K = (int)rows; g.sCm = 1; g.sCn = ldw; g.atomic = 1;
  g.kb_total = (int)ceil_div64(rows, T2_BK);
  const int tiles = ceil_div(Kx, T2_BM) * ceil_div(N, bn);
  wgrad_splits(g, tiles);
  dim3 grid(ceil_div(Kx, T2_BM), ceil_div(N, bn), ceil_div(g.kb_total, g.kb_per_split));
  return launch2_bn<1, true>(bn, ta, tb, tb, g, grid, s);
}

int tc2_conv_wgrad(const ConvOp& o, const float* dy, int ldy, int N, float* dWp, int ldw, cudaStream_t s) {
  if (!conv_tc_supported(o, true) || N < 1 || ldy % 4 != 0 || !al16(dy)) return DDRL_E_UNSUPPORTED;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn(N);
  Tc2Args g;
  memset(&g, 0, sizeof(g));
  tc_tap_common(g.tap, o, 32);
  if (g.tap.nb != 1) { g.tap.nb = 1; g.tap.rows = o.Xn * g.tap.ny; g.tap.kpad = (g.tap.rows + 7) & ~7; }
  const int K = o.KH * o.KW * o.Cin;
  CUtensorMap ta, tb, ta2, tb2;
  // K blocks of exactly 32 pixels from two box classes when the map width is a sum of two powers of two and that packs
  // the image into >= 10 % fewer K blocks than whole rows (the launch is bound by the per-K-block hand-off chain)
  static const bool no_w2 = [] { const char* e = getenv("DDRL_TC2_NO_WGRAD_BOXES"); return e && e[0] == '1'; }();
  if (!no_w2 && ldy == N) {
    for (int xw0 = 32; xw0 >= 2; xw0 >>= 1) {
      const int xw1 = o.Xn - xw0;
      if (xw1 < 1 || xw1 > xw0 || (xw1 & (xw1 - 1)) != 0) continue;
      const int yh0 = 32 / xw0, yh1 = 32 / xw1;
      if ((yh0 - 1) * o.sy + 1 > 256 || (yh1 - 1) * o.sy + 1 > 256) continue;
      const int nb0 = ceil_div(o.Yn, yh0), nb1 = ceil_div(o.Yn, yh1);
      if ((nb0 + nb1) * 10 > g.tap.tpi * 9) continue;
      g.tap.w2on = 1; g.tap.wxw0 = xw0; g.tap.wyh0 = yh0; g.tap.wnb0 = nb0; g.tap.wxw1 = xw1; g.tap.wyh1 = yh1;
      g.tap.tpi = nb0 + nb1; g.tap.rows = 32; g.tap.kpad = 32;
      break;
    }
  }
  if (g.tap.w2on) {
    r = tc_make_map_nhwc(&ta, o, g.tap.wxw0, g.tap.wyh0, 1, false);
    if (r == DDRL_OK) r = tc_make_map_nhwc(&ta2, o, g.tap.wxw1, g.tap.wyh1, 1, false);
    if (r == DDRL_OK) r = tc_make_map_dy4(&tb, dy, ldy, N, o.Xn, o.Yn, o.Bn, g.tap.wxw0, g.tap.wyh0);
    if (r == DDRL_OK) r = tc_make_map_dy4(&tb2, dy, ldy, N, o.Xn, o.Yn, o.Bn, g.tap.wxw1, g.tap.wyh1);
  } else {
    r = tc_make_map_nhwc(&ta, o, o.Xn, g.tap.ny, 1, false);
    if (r == DDRL_OK) r = tc_make_map_dy3(&tb, dy, ldy, N, o.Yn * o.Xn, o.Bn, g.tap.rows);
  }
  if (r != DDRL_OK) return r;
  g.C = dWp; g.M = K; g.N = N; g.K = o.Bn * o.Yn * o.Xn; g.sCm = 1; g.sCn = ldw; g.atomic = 1;
  g.kb_total = o.Bn * g.tap.tpi;
  const int tiles = ceil_div(K, T2_BM) * ceil_div(N, bn);
  wgrad_splits(g, tiles);
  dim3 grid(ceil_div(K, T2_BM), ceil_div(N, bn), ceil_div(g.kb_total, g.kb_per_split));
  if (g.tap.w2on) return launch2_bn<1, true>(bn, ta, tb, tb2, g, grid, s, &ta2);
  return launch2_bn<1, true>(bn, ta, tb, tb, g, grid, s);
}

// hi = rn_tf32(w), lo = rn_tf32(w - hi): the weight operand's split, once per optimiser step
__global__ void __launch_bounds__(256) split_hi_lo_kernel(const float4* __restrict__ w, uint4* __restrict__ hi,
                                                          uint4* __restrict__ lo, long long n4) {
  split_hi_lo_body(w, hi, lo, n4, blockIdx.x, gridDim.x);
}
int split_hi_lo(const float* w, float* hi, float* lo, long long n, cudaStream_t s) {
  if (n % 4 != 0 || !al16(w) || !al16(hi) || !al16(lo)) return DDRL_E_ARG;
  if (n == 0) return DDRL_OK;
  const long long n4 = n / 4;
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_SPLIT_HILO; j.a = w; j.b = hi; j.c = lo; j.total = n4; j.vblocks = prep_blocks(n4);
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 8LL * kNumSMs);
  split_hi_lo_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(w), reinterpret_cast<uint4*>(hi),
                                            reinterpret_cast<uint4*>(lo), n4);
  DDRL_LAUNCHED("split_hi_lo_kernel");
  return DDRL_OK;
}

}  // namespace ddrl

// debugging aid (not part of the public header): role wait counters of the tc2 engine (all zero unless built -DTC2_TIMING)
extern "C" int ddrl_tc2_timing_read(unsigned long long* out32, int reset) {
  if (out32) DDRL_CUDA(cudaMemcpyFromSymbol(out32, ddrl::g_tc2_wait, sizeof(unsigned long long) * 32));
  if (reset) {
    unsigned long long z[32] = {0};
    DDRL_CUDA(cudaMemcpyToSymbol(ddrl::g_tc2_wait, z, sizeof(z)));
  }
  return DDRL_OK;
}
