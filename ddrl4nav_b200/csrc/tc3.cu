// Third-generation tcgen05 engine: fp32 in / fp32 out through THREE kind::f16 products of scaled fp16 splits.
//
// What the round-1 engine (tc2.cu, 3xTF32) was bound by, measured on B200 (profiles/tf32_peak.json, r1d_tc2_skeleton.md):
//   * every M = 128 MMA with N <= 64 holds the tensor pipe for 45 clk and kind::tf32 covers only K = 8 per instruction;
//   * the single issuing thread paid two mbarrier round trips per 32-wide K block;
//   * the tf32 weight tiles are 4 bytes per element of shared-memory traffic on every MMA pass.
// This engine keeps tc2's structure (persistent CTAs, activation operand split by dedicated warps into TENSOR MEMORY,
// implicit-GEMM tap boxes, fused parity-class data gradient, swizzled staging epilogue) and changes the arithmetic and
// the hand-off granularity:
//   * operands are split as  x*s = hi + lo * 2^-11,  hi = rn_f16(x*s), lo = rn_f16((x*s - hi) * 2^11)  with a per-tensor
//     power-of-two scale s = 2^(13 - floor(log2 amax|x|)) (the fp16 exponent range is the only thing lost against tf32:
//     the 22 significant bits are the same, and elements below amax * 2^-27 only lose RELATIVE precision).  Products
//     hi*hi -> main accumulator, hi*lo' + lo'*hi -> correction accumulators, result = (main + corr * 2^-11) / (sA sB).
//     kind::f16 is K = 16 per MMA in the clocks kind::tf32 needs for K = 8: half the tensor time, half the weight bytes.
//   * K blocks are 64 wide (two 32-float TMA boxes of the raw activation tile; one 128-byte swizzle row of fp16 weights):
//     half the barrier round trips per FLOP; the MMA issuers wait on ONE barrier per K block (the splitter's "A slot
//     ready" arrival happens after the stage's TMA barrier, so it covers the weight tile as well).
//   * amax|x| of every GEMM operand is a device scalar maintained by its producer (this engine's epilogue tracks the
//     running max of what it stores: one atomicMax per epilogue warp per launch) or by `amax_f32` for foreign tensors.
//   * weights are split, scaled and -- for the data gradient of linear layers -- transposed once per optimiser step, so
//     the weight operand is always K-major.
// Accumulation: the tensor core adds into fp32 with truncation (scratch/tc_acc.py), so `main` lives in TMEM only for 16
// MMAs (4 K blocks) and is then added round-to-nearest into fp32 registers, exactly as in tc2.
// Warp roles: 0 TMA producer | 1 MMA issuer (chunk accumulators) | 3 MMA issuer (whole-tile correction accumulator) |
// 2 TMEM allocator, then STORE WARP of the TMA epilogue (read-out waits, activation-mask loads, tile stores) | splitter
// groups of 4 warps | epilogue warps.
// Later additions (end of round 2; DESIGN.md section 5.1, profiles/r4_notes.md):
//   * pre-split activation operands (Tc3Args::presplit): fp16 hi | lo' planes written once by tc3_presplit -> both MMA operands
//     are plain TMA loads (SS MMAs; MN-major A in the weight gradient), the splitter warps idle;
//   * resident weight tiles when the K blocks of a tile map one-to-one onto the stages (Tc3Args::b_resident);
//   * sign-bit activation masks (Tc3Args::bits_out / bits_in): forward launches leave one bit per stored element, data
//     gradients read those words instead of the fp32 activation;
//   * weight gradient: two-phase K tiling of small maps (multi-image pixel boxes), conversion-task offsets hoisted.
// What bounds the N = 64 launches now is L2 -> SM bandwidth (the fp32 tile is fetched once per tap): ncu shows 9.4-10 TB/s
// on them (profiles/r4z_ncu_tc3_table.md).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "layer_ops.h"
#include "prep_kernels.cuh"
#include "tc_ptx.cuh"

namespace ddrl {

constexpr int T3_BM = 128;
constexpr int T3_BK = 64;                     // k per K block (two 32-float boxes of A; 64 halfs = one swizzle row of B)
constexpr int T3_CHUNK = 4;                   // K blocks per TMEM main-accumulator chunk (16 accumulating MMAs)
// lo' = lo * 2^11 keeps the residual in the fp16 normal range down to |x s| = 2^-14, i.e. 2^-27 of the tensor's amax.
// Measured alternative (profiles/r2h_*): an unscaled residual and s = 1 for tensors already inside the fp16 range save
// two of ten splitter instructions per element pair (forward convs 5 % faster), but rows 2^-20 below the tensor's amax
// drop from 22 to ~10 significant bits and micro-batched runs stop being bit-identical to single-shot ones (the power-of-two
// scale is otherwise exact, so the split mantissas do not depend on it): not adopted.
// (T3_LO / T3_LO_INV and t3_scale live in prep_kernels.cuh: the weight-preparation kernels share them)

// Optional role-level accounting (build with -DTC3_TIMING, scratch/tc3_roles.py): cycles each warp role of the forward
// kernel spends in each of its phases, accumulated over every launch.  [role*8 + k]; k = 7 is the role's lifetime.
// roles: 0 producer {empty} | 1 chunk MMA {mfree, aready, issue+commit} | 2 corr MMA {cfree, aready, issue+commit} |
//        3 splitter warp 4 {full, afree, lds+split, st+wait::st} | 4 epilogue warp 0 {mfull, cfull, stores, drain}
__device__ unsigned long long g_tc3_wait[96];   // weight-gradient kernel: roles 5 producer | 6 chunk MMA | 7 splitter warp 4 | 8 epilogue warp 0
#ifdef TC3_TIMING
#define T3_T0 const long long _t0 = clock64()
#define T3_ACC(acc) acc += clock64() - _t0
#define T3_WAIT(bar, par, acc) do { T3_T0; mbar_wait(bar, par); T3_ACC(acc); } while (0)
#define T3_WAITL(bar, par, acc) do { T3_T0; mbar_wait_long(bar, par); T3_ACC(acc); } while (0)
#define T3_ROLE_BEGIN long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0; const long long role_t0 = clock64();
#ifdef TC3_TIMING_PS            // account only the launches whose activation operand is pre-split
#define T3_TCOND (g.presplit != 0)
#else
#define T3_TCOND true
#endif
#define T3_ROLE_END(role, cond) do { if ((cond) && T3_TCOND && lane == 0) { \
    atomicAdd(&g_tc3_wait[(role) * 8 + 0], (unsigned long long)w0); atomicAdd(&g_tc3_wait[(role) * 8 + 1], (unsigned long long)w1); \
    atomicAdd(&g_tc3_wait[(role) * 8 + 2], (unsigned long long)w2); atomicAdd(&g_tc3_wait[(role) * 8 + 3], (unsigned long long)w3); \
    atomicAdd(&g_tc3_wait[(role) * 8 + 4], (unsigned long long)w4); atomicAdd(&g_tc3_wait[(role) * 8 + 5], (unsigned long long)w5); \
    atomicAdd(&g_tc3_wait[(role) * 8 + 6], (unsigned long long)w6); \
    atomicAdd(&g_tc3_wait[(role) * 8 + 7], (unsigned long long)(clock64() - role_t0)); } } while (0)
#define T3_SECTION_BEGIN const long long _s0 = clock64()
#define T3_SECTION_END(acc) acc += clock64() - _s0
#else
#define T3_WAIT(bar, par, acc) mbar_wait(bar, par)
#define T3_WAITL(bar, par, acc) mbar_wait_long(bar, par)
#define T3_ROLE_BEGIN
#define T3_ROLE_END(role, cond)
#define T3_SECTION_BEGIN
#define T3_SECTION_END(acc)
#endif

struct Tc3Args {
  float* C;
  const float* bias;
  const float* mask;
  long long sCm, sCn;
  int M, N, K;
  int kb_total, kb_per_split;                 // K blocks of 64 (tap mode: pairs of 32-channel slices)
  int act, atomic, vec_store;
  int tma_store;                              // outputs leave through TMA tile stores of the staged panels (tmC / tmC2)
  int tma_mask;                               // ... and the activation mask (act >= 3) ARRIVES through TMA, into the same panels (tmM / tmM2)
  // The activation operand was split ONCE, by its producer, into two fp16 planes (hi | lo', the tensor's layout with 2-byte
  // elements, scaled by t3_scale(*amax_a)): both MMA operands are plain TMA loads, the A tiles stay in shared memory (SS
  // MMAs) and the splitter warps have nothing to do.  Forward kernel: tmA / tmA2 = hi plane, tmM / tmM2 = lo' plane (such
  // launches carry no activation mask); K blocks are one 64-channel slice of one tap.  Weight gradient: tmA / tmA2 = hi plane,
  // tmAl / tmAl2 = lo' plane, the landed [64 pixels x 64 channels] boxes ARE the MN-major A tiles.
  int presplit;
  // forward kernel, one N tile and a K-block count that divides the stage count: K block i always lands in stage i mod nkb, so
  // its weight tiles are loaded ONCE per CTA and stay there -- later tiles only stream the activation tiles (the short-K first
  // conv moves 184 KB per tile through L2 otherwise and sits at the L2 throughput cap, profiles/r4c_roles_ps.txt)
  int b_resident;
  int m_tiles, n_tiles;
  const float* amax_a;                        // device scalars: amax of the activation operand / of the weight operand
  const float* amax_b;
  float* amax_out;                            // optional: running amax of the stored output (atomicMax on the bits)
  // Sign bits of an activation tensor (TMA epilogue): one bit per element, word = element offset / 32, set where the stored
  // activation is > 0.  A forward launch whose output has a registered bit tensor writes it (bits_out, one word per lane and
  // 32-column round); the data gradient that needs that activation's derivative reads the words (bits_in, addressed like
  // its own output) instead of loading the whole fp32 activation through the staging panels: 1/32 of the mask bytes and no
  // TMA round trip inside the tile.
  unsigned int* bits_out;
  const unsigned int* bits_in;
  float* colsum;                              // weight gradient: optional db[n] += sum_r dy[r, n] (bias gradient), fused into the dy conversion
  float* colsum2;                             // columns >= colsum_split go to colsum2[n - colsum_split] (two layers behind one fused dy)
  int colsum_split;
  unsigned int* det_ctr;                      // weight gradient, deterministic mode: turn counters, one per (M block, N tile)
  TcTap tap;
};

// (x0, x1) * s -> packed fp16 hi pair and packed fp16 lo' pair (saturating: a stale amax gives a wrong, finite result)
template <bool SCALED = true>
__device__ __forceinline__ void t3_split2(float x0, float x1, float s, uint32_t& hi, uint32_t& lo) {
  const float y0 = SCALED ? x0 * s : x0, y1 = SCALED ? x1 * s : x1;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y1), "f"(y0));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float r0 = (y0 - f.x) * T3_LO, r1 = (y1 - f.y) * T3_LO;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BN>
struct T3Cfg {
  static constexpr int A_SUB = T3_BM * 128;                      // one 32-float box of the raw activation tile: 16 KB
  static constexpr int A_BYTES = 2 * A_SUB;
  static constexpr int B_BYTES = BN * 128;                       // BN rows x 64 halfs
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;      // A raw (2 boxes) | B hi | B lo
  static constexpr int STAGES = BN <= 32 ? 5 : (BN <= 64 ? 4 : 3);
  // BN <= 64: hi*hi and hi*lo are ONE MMA of N' = 2 BN over the adjacent [B_hi ; B_lo] tiles (N' = 128 runs at the full
  // N/2-clk rate, two N = 64 MMAs would cost 45 clk each); lo*hi goes to the whole-tile correction accumulator.
  static constexpr bool FOLD = BN <= 64;
  static constexpr int SA = BN <= 32 ? 4 : (BN <= 64 ? 3 : 2);   // TMEM A slots: 64 columns each = [hi 32 | lo 32] for 64 k
  static constexpr int NEPI = BN <= 32 ? 4 : 8;
  static constexpr int NSG = BN <= 64 ? 2 : 1;                   // splitter groups (4 warps each), K blocks round-robin
  static constexpr int EPI0 = 4 + 4 * NSG;
  static constexpr int THREADS = (EPI0 + NEPI) * 32;
  static constexpr int COLS = BN / (NEPI / 4);
  // FOLD: [main0 | corrB0 | main1 | corrB1 | corrA | A slots]; else [main0 | main1 | corr | A slots]
  static constexpr int TM_MAIN0 = 0, TM_MAIN1 = FOLD ? 2 * BN : BN, TM_CORR = FOLD ? 4 * BN : 2 * BN, TM_A = FOLD ? 5 * BN : 3 * BN;
  static constexpr int TMEM_COLS = 512;
  static constexpr int NBARS = 2 * STAGES + SA + 7;
  // epilogue staging: NEPI/4 panels of 128 rows x 32 fp32 columns (128-byte rows, 128B swizzle = the box layout of a TMA
  // tile store; warp e owns rows [32 (e & 3), +32) of panel e >> 2), 1024-byte aligned
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STG_OFF + NEPI * 4096;
  static constexpr int BIAS_OFF = BAR_OFF + 256;                 // the tile's BN bias values (zero past N), TMA epilogue
  static constexpr int SMEM = 1024 + BIAS_OFF + BN * 4;
  static_assert(NBARS * 8 + 16 <= 256, "barrier block");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(TM_A + SA * 64 <= 512, "TMEM budget");
  static_assert(STAGES > SA, "the stage barrier of K block it - SA releases TMEM slot it % SA");
};

// C[M,N] = epi( (A[M,K] . B[N,K]^T) )   A: fp32, plain 2-D or tap boxes; B: pre-split fp16 hi / lo', K-major
template <int BN>
__global__ void __launch_bounds__(T3Cfg<BN>::THREADS, 1)
tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
           const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmA2,
           const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
           const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmM2, Tc3Args g) {
  using Cfg = T3Cfg<BN>;
  constexpr int S = Cfg::STAGES, SA = Cfg::SA;
  extern __shared__ uint8_t smem_dyn[];
  // 1024-byte alignment by POINTER arithmetic on the __shared__ array: an integer round-trip hides the address space from
  // the compiler, which then emits generic LD / ST (long-scoreboard, L1TEX path) for every shared-memory access below
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* bar_full = bars;                  // [S]  TMA landed
  uint64_t* bar_empty = bars + S;             // [S]  MMAs that read the stage (and its TMEM slot) retired
  uint64_t* bar_aready = bars + 2 * S;        // [SA] TMEM A slot written
  uint64_t* bar_mfull = bars + 2 * S + SA;    // [2] main accumulator chunk complete
  uint64_t* bar_mfree = bar_mfull + 2;        // [2] drained
  uint64_t* bar_cfull = bar_mfull + 4;        // correction accumulator complete (tile end)
  uint64_t* bar_cfree = bar_mfull + 5;        // read by the epilogue
  uint64_t* bar_mask = bar_mfull + 6;         // activation-mask panels landed in the staging area (TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcTap& tp = g.tap;
  const bool tapA = tp.mode != 0;

  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(bar_mask), 1);
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

  // ---- work enumeration, identical in every role: tiles blockIdx.x, +gridDim.x, ... of the (m_tiles x n_tiles) space
  const int total_tiles = g.m_tiles * g.n_tiles;
  const int tile_step = (int)gridDim.x, tile_first = (int)blockIdx.x;
  const int nkb = g.kb_total;
  // tap mode with an odd slice count: the last K block holds ONE 32-channel slice (2 k steps)
  const bool ps = g.presplit != 0;
  const bool odd_tail = tapA && !ps && (tp.nslices & 1);

  if (warp == 0) {
    // ============================================================ TMA producer
    uint32_t it = 0;
    T3_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
      const int m0 = mt * T3_BM, n0 = nt * BN;
      int b0 = 0, y0 = 0;
      const bool ph2 = tapA && mt >= tp.tiles1;       // second tiling phase: the images' remaining rows
      if (tapA) {
        if (!ph2) { b0 = (mt / tp.tpi) * tp.nb; y0 = (mt % tp.tpi) * tp.ny * tp.sy - tp.py; }
        else { b0 = (mt - tp.tiles1) * tp.nb2; y0 = tp.y2 * tp.sy - tp.py; }
      }
      int cc = 0, kw = 0, kh = 0;                     // tap slices advance (chunk fastest, then kw, then kh)
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S;
        T3_WAITL(smem_u32(bar_empty + s), ((it / S) & 1) ^ 1, w0);
        const bool load_b = !g.b_resident || it < (uint32_t)S;
        const uint32_t b_tx = load_b ? 2u * Cfg::B_BYTES : 0u;
        if (elect_one()) {
          const uint32_t full = smem_u32(bar_full + s);
          const uint32_t a_dst = smem_u32(smem) + s * Cfg::STAGE_BYTES, bh_dst = a_dst + Cfg::A_BYTES, bl_dst = bh_dst + Cfg::B_BYTES;
          const int k = i * T3_BK;
          if (ps) {
            // pre-split operand: [hi tile | lo' tile], 128 rows x 64 halfs each (tap mode: tp.cpb counts 64-channel slices)
            if (!tapA) {
              mbar_expect_tx(full, Cfg::A_BYTES + b_tx);
              tma_load_2d(&tmA, full, a_dst, k, m0);
              tma_load_2d(&tmM, full, a_dst + Cfg::A_SUB, k, m0);
            } else {
              const uint32_t box = (uint32_t)(ph2 ? tp.rows2 : tp.rows) * 128u;
              mbar_expect_tx(full, 2u * box + b_tx);
              tma_load_4d(ph2 ? &tmA2 : &tmA, full, a_dst, tp.c_off + cc * 64, kw - tp.px, y0 + kh, b0);
              tma_load_4d(ph2 ? &tmM2 : &tmM, full, a_dst + Cfg::A_SUB, tp.c_off + cc * 64, kw - tp.px, y0 + kh, b0);
            }
          } else if (!tapA) {
            mbar_expect_tx(full, Cfg::A_BYTES + b_tx);
            tma_load_2d(&tmA, full, a_dst, k, m0);
            tma_load_2d(&tmA, full, a_dst + Cfg::A_SUB, k + 32, m0);
          } else {
            const bool two = !(odd_tail && i == nkb - 1);
            const uint32_t box = (uint32_t)(ph2 ? tp.rows2 : tp.rows) * 128u;
            mbar_expect_tx(full, (two ? 2u : 1u) * box + b_tx);
            tma_load_4d(ph2 ? &tmA2 : &tmA, full, a_dst, tp.c_off + cc * 32, kw - tp.px, y0 + kh, b0);
            if (two) {
              int cc2 = cc + 1, kw2 = kw, kh2 = kh;
              if (cc2 == tp.cpb) { cc2 = 0; if (++kw2 == tp.KW) { kw2 = 0; ++kh2; } }
              tma_load_4d(ph2 ? &tmA2 : &tmA, full, a_dst + Cfg::A_SUB, tp.c_off + cc2 * 32, kw2 - tp.px, y0 + kh2, b0);
            }
          }
          if (load_b) {
            tma_load_2d(&tmBhi, full, bh_dst, k, n0);
            tma_load_2d(&tmBlo, full, bl_dst, k, n0);
          }
        }
        __syncwarp();
        if (tapA) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (ps && j == 1) break;                  // pre-split: one (64-channel) slice per K block
            if (++cc == tp.cpb) { cc = 0; if (++kw == tp.KW) { kw = 0; ++kh; } }
          }
        }
      }
    }
    T3_ROLE_END(0, true);
  } else if (warp == 1 || warp == 3) {
    // ============================================================ MMA issuers (one elected thread each)
    // warp 1: chunk buffers   main (+)= A_hi . B_hi          [FOLD: [main | corrB] (+)= A_hi . [B_hi ; B_lo], N' = 2 BN]
    // warp 3: whole tile      corr  += A_lo . B_hi           [!FOLD: ... + A_hi . B_lo]
    const bool chunk_role = warp == 1;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);        // f16 x f16 -> f32
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t t_corr = tmem_base + Cfg::TM_CORR;
    uint32_t it = 0, ch = 0, tl = 0;
    T3_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tl) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S, a = it % SA;
        const uint32_t buf = ch & 1;
        const bool first_in_chunk = (i % T3_CHUNK) == 0;
        const bool last_in_chunk = (i % T3_CHUNK) == T3_CHUNK - 1 || i == nkb - 1;
        const int ksteps = (odd_tail && i == nkb - 1) ? 2 : 4;
        if (chunk_role) { if (first_in_chunk) T3_WAIT(smem_u32(bar_mfree + buf), ((ch >> 1) & 1) ^ 1, w0); }
        else if (i == 0) T3_WAIT(smem_u32(bar_cfree), (tl & 1) ^ 1, w0);
        // ONE wait per K block: the splitters arrive on `aready` after they have seen the stage's TMA barrier, so the
        // weight tile of the stage is complete (and visible through the same release / acquire chain) as well
        // (pre-split operand: nobody stands between the TMA and the MMAs -- both issuers wait for the stage itself)
        if (ps) T3_WAIT(smem_u32(bar_full + s), (it / S) & 1, w1);
        else T3_WAIT(smem_u32(bar_aready + a), (it / SA) & 1, w1);
        tc_fence_after();
        T3_SECTION_BEGIN;
        if (elect_one()) {
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
          const uint32_t a_hi = tmem_base + Cfg::TM_A + a * 64, a_lo = a_hi + 32;
          const uint64_t dbh0 = umma_desc(b_hi, 16, 1024, 2);
          if (ps) {
            // A tiles in shared memory: K-major, 128 rows x 64 halfs, 128B swizzle -- the weight tile's own layout
            const uint32_t sa_hi = smem_base + s * Cfg::STAGE_BYTES;
            const uint64_t dah0 = umma_desc(sa_hi, 16, 1024, 2), dal0 = umma_desc(sa_hi + Cfg::A_SUB, 16, 1024, 2);
            if (chunk_role) {
              const uint32_t t_main = tmem_base + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_f16_ss(t_main, dah0 + k4 * 2, dbh0 + k4 * 2, Cfg::FOLD ? idesc2 : idesc, (!first_in_chunk || k4 != 0) ? 1u : 0u);
              umma_commit(smem_u32(bar_empty + s));
              if (last_in_chunk) umma_commit(smem_u32(bar_mfull + buf));
            } else {
              const uint64_t dbl0 = umma_desc(b_lo, 16, 1024, 2);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_f16_ss(t_corr, dal0 + k4 * 2, dbh0 + k4 * 2, idesc, (i | k4) != 0 ? 1u : 0u);
                if (!Cfg::FOLD) umma_f16_ss(t_corr, dah0 + k4 * 2, dbl0 + k4 * 2, idesc, 1u);
              }
              umma_commit(smem_u32(bar_empty + s));
              if (i == nkb - 1) umma_commit(smem_u32(bar_cfull));
            }
          } else if (chunk_role) {
            const uint32_t t_main = tmem_base + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (k4 >= ksteps) break;
              umma_f16_ts(t_main, a_hi + k4 * 8, dbh0 + k4 * 2, Cfg::FOLD ? idesc2 : idesc, (!first_in_chunk || k4 != 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(bar_empty + s));            // retires the stage AND the TMEM A slot of this K block
            if (last_in_chunk) umma_commit(smem_u32(bar_mfull + buf));
          } else {
            const uint64_t dbl0 = umma_desc(b_lo, 16, 1024, 2);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (k4 >= ksteps) break;
              umma_f16_ts(t_corr, a_lo + k4 * 8, dbh0 + k4 * 2, idesc, (i | k4) != 0 ? 1u : 0u);
              if (!Cfg::FOLD) umma_f16_ts(t_corr, a_hi + k4 * 8, dbl0 + k4 * 2, idesc, 1u);
            }
            umma_commit(smem_u32(bar_empty + s));
            if (i == nkb - 1) umma_commit(smem_u32(bar_cfull));
          }
        }
        __syncwarp();
        T3_SECTION_END(w2);
        if (last_in_chunk) ++ch;
      }
    }
    T3_ROLE_END(chunk_role ? 1 : 2, true);
  } else if (warp >= 4 && warp < Cfg::EPI0) {
    // ============================================================ splitters: smem A (fp32) -> scaled fp16 hi / lo' -> TMEM
    const int q = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    float sA, sA_inv;
    t3_scale(__ldg(g.amax_a), sA, sA_inv);
    const bool unit_scale = sA == 1.f;
    const int row = q * 32 + lane;
    uint32_t it = 0;
    T3_ROLE_BEGIN
    for (int tile = ps ? total_tiles : tile_first; tile < total_tiles; tile += tile_step) {    // (pre-split operand: nothing to do)
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % S, a = it % SA;
        if ((int)(it % Cfg::NSG) != grp) {                         // the groups take K blocks round-robin
          // ... but every group follows EVERY phase of the stage barriers: when S is not a multiple of NSG a group meets a
          // stage only on every other use, and a parity wait that skips a phase is satisfied by the phase BEFORE the skipped
          // one -- the group would read the stage while the skipped block's (or its own block's) TMA is still in flight
          if (S % Cfg::NSG != 0) T3_WAITL(smem_u32(bar_full + s), (it / S) & 1, w0);
          continue;
        }
        T3_WAITL(smem_u32(bar_full + s), (it / S) & 1, w0);
        T3_SECTION_BEGIN;
        const uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const uint32_t ta = tmem_base + t_lane + Cfg::TM_A + a * 64;
        const int nsub = (odd_tail && i == nkb - 1) ? 1 : 2;
        // thread = tile row; sub-tile j holds k [32 j, 32 j + 32) of the row as eight 16-byte chunks (128B swizzle)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j >= nsub) break;
          uint32_t hi[16], lo[16];
          const uint8_t* rp = st + j * Cfg::A_SUB + row * 128;
          if (unit_scale) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
              t3_split2<false>(v.x, v.y, 1.f, hi[2 * c], lo[2 * c]);
              t3_split2<false>(v.z, v.w, 1.f, hi[2 * c + 1], lo[2 * c + 1]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
              t3_split2<true>(v.x, v.y, sA, hi[2 * c], lo[2 * c]);
              t3_split2<true>(v.z, v.w, sA, hi[2 * c + 1], lo[2 * c + 1]);
            }
          }
          if (j == 0) {
            // TMEM slot a was last read by K block it - SA: its stage barrier (both MMA streams commit to it) doubles as
            // the slot's release -- S > SA, so that barrier cannot be a second phase ahead when we look at it
            if (it >= (uint32_t)SA) {
              const uint32_t jj = it - SA;
              T3_WAIT(smem_u32(bar_empty + (jj % S)), (jj / S) & 1, w1);
            }
            tc_fence_after();
          }
          tmem_st16(ta + j * 16, hi);
          tmem_st16(ta + 32 + j * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_aready + a));
        T3_SECTION_END(w2);                                        // includes the afree wait (subtract w1)
      }
    }
    T3_ROLE_END(3, warp == 4);
  } else if (warp == 2) {
    // ============================================================ store warp (TMA epilogue only)
    // Lane 0 owns every bulk-async group of the epilogue: per tile and 32-column round it waits until the previous stores have
    // READ the staging panels, requests the round's activation-mask panels (data gradients; the first round's travel while the
    // epilogue warps still drain the accumulators), meets the epilogue warps at "panels free", again at "panels complete", and
    // stores the panels with tensor-map boxes that mirror the load-side boxes.
    if (g.tma_store) {
      const bool masked = g.tma_mask != 0;
      const bool fusedT = tapA && tp.ncls > 1;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
        const int m0 = mt * T3_BM, n0 = nt * BN;
        int cy = 0, cb = 0;
        const bool ph2 = tapA && mt >= tp.tiles1;
        if (tapA) {
          cb = ph2 ? (mt - tp.tiles1) * tp.nb2 : (mt / tp.tpi) * tp.nb;
          cy = ph2 ? tp.y2 : (mt % tp.tpi) * tp.ny;
        }
        const uint32_t box_bytes = tapA ? (uint32_t)(ph2 ? tp.rows2 : tp.rows) * 128u : 16384u;
        // global coordinates of the panel that starts at column `col`: a panel of the fused stride-parity gradient belongs to
        // one pixel class, whose offset is the start of a strided box
        auto coords = [&](int col, int& c0, int& c1, int& c2, int& c3) {
          c0 = col; c1 = 0; c2 = cy; c3 = cb;
          if (fusedT) {
            const int qq = col / tp.cls_cols;
            c0 = col - qq * tp.cls_cols; c1 = tp.cls_ix[qq]; c2 = tp.out_s * cy + tp.cls_iy[qq];
          }
        };
#pragma unroll 1
        for (int p0 = 0; p0 < Cfg::COLS; p0 += 32) {
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (masked) {
              uint32_t live = 0;
#pragma unroll
              for (int h = 0; h < Cfg::NEPI / 4; ++h) live += (n0 + h * Cfg::COLS + p0 < g.N) ? 1u : 0u;
              mbar_expect_tx(smem_u32(bar_mask), live * box_bytes);
#pragma unroll
              for (int h = 0; h < Cfg::NEPI / 4; ++h) {
                const int col = n0 + h * Cfg::COLS + p0;
                if (col >= g.N) continue;
                const uint32_t dst = smem_u32(smem + Cfg::STG_OFF) + h * 16384;
                if (!tapA) tma_load_2d(&tmM, smem_u32(bar_mask), dst, col, m0);
                else {
                  int c0, c1, c2, c3;
                  coords(col, c0, c1, c2, c3);
                  tma_load_4d(ph2 ? &tmM2 : &tmM, smem_u32(bar_mask), dst, c0, c1, c2, c3);
                }
              }
            }
          }
          __syncwarp();
          asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32 + 32) : "memory");       // panels free
          asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32 + 32) : "memory");       // panels complete
          if (lane == 0) {
#pragma unroll
            for (int h = 0; h < Cfg::NEPI / 4; ++h) {
              const int col = n0 + h * Cfg::COLS + p0;
              if (col >= g.N) continue;
              const uint32_t src = smem_u32(smem + Cfg::STG_OFF) + h * 16384;
              if (!tapA)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(src), "r"(col), "r"(m0) : "memory");
              else {
                int c0, c1, c2, c3;
                coords(col, c0, c1, c2, c3);
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(ph2 ? &tmC2 : &tmC)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                             : "memory");
              }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  } else if (warp >= Cfg::EPI0) {
    // ============================================================ drain + epilogue
    const int e = warp - Cfg::EPI0;
    const int q = e & 3, half = e >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const uint32_t col0 = half * Cfg::COLS;
    float sA, sA_inv, sB, sB_inv;
    t3_scale(__ldg(g.amax_a), sA, sA_inv);
    t3_scale(__ldg(g.amax_b), sB, sB_inv);
    float run_max = 0.f;
    uint32_t ch = 0, tl = 0, mround = 0;
    // result scale: both factors are powers of two, so one multiplication by their product is exact -- unless the product
    // itself leaves the normal range (operands near 1e+-19), where the two-step form is kept
    float s_pre = 1.f, s_out = sA_inv * sB_inv;
    if (!(s_out >= 1e-30f && s_out <= 1e30f)) { s_pre = sA_inv; s_out = sB_inv; }
    // tile row -> (pixel, row, image) inside the tile's box, for both tiling phases: once per kernel, not per tile (the four
    // integer divisions were ~100 instructions of every tile's epilogue)
    const int r = q * 32 + lane;
    int ryb1 = 0, ryb2 = 0;                              // row | image << 16
    const bool need_off = !g.tma_store || g.bits_out != nullptr || g.bits_in != nullptr;      // per-row element offsets
    const int rx = (tapA && need_off) ? r % tp.Xn : 0;
    if (tapA) {
      const int t1 = r / tp.Xn;
      ryb1 = (t1 % tp.ny) | ((t1 / tp.ny) << 16);
      if (tp.ny2 > 0) ryb2 = (t1 % tp.ny2) | ((t1 / tp.ny2) << 16);
    }
    float* bias_s = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF);
    const int et = e * 32 + lane;
    T3_ROLE_BEGIN
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tl) {
      const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
      const int m0 = mt * T3_BM, n0 = nt * BN;
      // tile-level coordinates of the TMA epilogue (stores, and -- data gradients -- the mask panels that arrive the same way)
      int cy = 0, cb = 0;
      if (tapA) {
        const bool ph2 = mt >= tp.tiles1;
        cb = ph2 ? (mt - tp.tiles1) * tp.nb2 : (mt / tp.tpi) * tp.nb;
        cy = ph2 ? tp.y2 : (mt % tp.tpi) * tp.ny;
      }
      const bool masked = g.tma_store && g.tma_mask;
      // the tile's bias values: requested now, parked in shared memory after the drain (the load's latency hides behind it)
      float bias_r = 0.f;
      if (g.tma_store && et < BN && g.bias != nullptr && n0 + et < g.N) bias_r = __ldg(g.bias + n0 + et);
      const bool ph2e = tapA && mt >= tp.tiles1;
      const int ryb = ph2e ? ryb2 : ryb1;
      const bool rvalid = tapA ? (r < (ph2e ? tp.rows2 : tp.rows) && cb + (ryb >> 16) < tp.Bn && cy + (ryb & 0xffff) < tp.Yn)
                               : (m0 + r) < g.M;
      // fused parity classes: tile pixel (py, px) -> input pixel (out_s*py + cls_iy, out_s*px + cls_ix) per column group
      const bool fused = tapA && tp.ncls > 1;
      long long roff = 0;                              // element offset of this thread's row in the output (mask) tensor
      int py = 0, px = 0;
      if (need_off) {
        if (tapA) {
          const int yy = ryb & 0xffff, bb = ryb >> 16;
          roff = (long long)(cb + bb) * tp.osb + (long long)(cy + yy) * tp.osy + (long long)rx * tp.osx;
          if (fused) { px = rx * tp.out_s; py = (cy + yy) * tp.out_s; }
        } else {
          roff = (long long)(m0 + r) * g.sCm;
        }
      }
      // activation-derivative bits of the row's 32-column groups (data gradients): requested before the drain as well
      uint32_t mw[Cfg::COLS / 32];
#pragma unroll
      for (int p = 0; p < Cfg::COLS / 32; ++p) mw[p] = 0u;
      if (g.bits_in != nullptr && rvalid) {
#pragma unroll
        for (int p = 0; p < Cfg::COLS / 32; ++p) {
          const int col = n0 + col0 + p * 32;
          if (col >= g.N) continue;
          long long eoff = roff + col;
          bool inb = true;
          if (fused) {
            const int qq = col / tp.cls_cols;
            inb = py + tp.cls_iy[qq] < tp.out_H && px + tp.cls_ix[qq] < tp.out_W;
            eoff += tp.cls_off[qq] - (long long)qq * tp.cls_cols;
          }
          if (inb) mw[p] = __ldg(g.bits_in + (eoff >> 5));
        }
      }
      float acc[Cfg::COLS];
#pragma unroll
      for (int j = 0; j < Cfg::COLS; ++j) acc[j] = 0.f;
      const int nch = (nkb + T3_CHUNK - 1) / T3_CHUNK;
      for (int c = 0; c < nch; ++c, ++ch) {
        const int buf = ch & 1;
        T3_WAITL(smem_u32(bar_mfull + buf), (ch >> 1) & 1, w0);
        tc_fence_after();
#pragma unroll
        for (int j0 = 0; j0 < Cfg::COLS; j0 += 32) {
          float v[32];
          if (Cfg::FOLD && c == 0) {
            // first chunk of a tile: the main term lands in the (empty) accumulator registers directly, together with the
            // hi*lo' term -- both loads in flight behind one wait
            tmem_ld32x2(tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + col0 + j0,
                        tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + BN + col0 + j0, acc + j0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j0 + j] = fmaf(v[j], T3_LO_INV, acc[j0 + j]);
            continue;
          }
          tmem_ld32(tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + col0 + j0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j0 + j] += v[j];
          if (Cfg::FOLD) {
            // the hi*lo' term of this chunk sits BN columns further (second half of the folded N' = 2 BN MMA)
            tmem_ld32(tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + BN + col0 + j0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j0 + j] = fmaf(v[j], T3_LO_INV, acc[j0 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_mfree + buf));
      }
      T3_WAITL(smem_u32(bar_cfull), tl & 1, w1);
      tc_fence_after();
#pragma unroll
      for (int j0 = 0; j0 < Cfg::COLS; j0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + t_lane + Cfg::TM_CORR + col0 + j0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j0 + j] = fmaf(v[j], T3_LO_INV, acc[j0 + j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_cfree));
      if (s_pre != 1.f) {
#pragma unroll
        for (int j = 0; j < Cfg::COLS; ++j) acc[j] *= s_pre;
      }
      // ---- stores
      T3_SECTION_BEGIN;
      const float neg_slope = g.act == 3 ? 0.f : 0.01f;
      if (!g.tma_store) {
#pragma unroll
        for (int j = 0; j < Cfg::COLS; ++j) acc[j] *= s_out;
      }
      if (g.tma_store) {
        // TMA path: bias + activation in registers, the warp's 32 rows x 32 columns into its slice of the swizzled panel,
        // then ONE elected thread stores the whole tile panel(s) with a tensor-map box that mirrors the load-side box
        // (rows past the tensor's extents are clipped by the TMA unit: no per-row addressing, no per-element stores).
        // Data gradients (g.tma_mask): the activation derivative needs the layer input at the very positions the panel is
        // stored to, so the SAME box is first loaded into the panel (tmM / tmM2), transformed in place and stored back; the
        // fused stride-parity gradient is one strided box per 32-column panel (a panel belongs to one pixel class).
        float* stg = reinterpret_cast<float*>(smem + Cfg::STG_OFF) + e * 1024;
        // Instruction count is what this path is bound by (profiles/r4b_roles_ps.txt: the epilogue warps of the short-K first
        // conv are 94 % busy, two thirds of it here), so per element it is: one FFMA (result scale + bias), the activation as
        // max(x, slope x), a share of an FMNMX3 for the running amax.  The tile's bias values sit in shared memory (zero past N:
        // the accumulators of those columns are zero, so they store and max as zero without a per-element column test).
        const bool relu = g.act == 1;
        const float slope = g.act == 2 ? 0.01f : 1.f;
        const bool has_bias = g.bias != nullptr;
        if (et < BN) bias_s[et] = bias_r;               // read after the panel barrier below
        // The TMA instructions themselves (read-out wait of the previous stores, mask-panel loads, tile stores) are issued by
        // the STORE WARP (warp 2), which meets these warps at the two panel barriers of every round: ~150 single-lane
        // instructions per tile that used to sit on the critical path of epilogue warp 0 -- and, through the accumulator
        // barriers that need all epilogue warps, of the whole tile (profiles/r4e_roles_ps.txt).
#pragma unroll
        for (int p0 = 0; p0 < Cfg::COLS; p0 += 32) {
#ifdef TC3_TIMING
          const long long _r0 = clock64();
#endif
          asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32 + 32) : "memory");       // panels free (and mask loads issued)
#ifdef TC3_TIMING
          w3 += clock64() - _r0;                          // read-out wait of the previous tile's TMA stores + panel barrier
#endif
          const float4* bias4 = reinterpret_cast<const float4*>(bias_s + col0 + p0);
          float tmax = 0.f;
          if (g.bits_in != nullptr) {
            // data gradient, activation derivative from the sign bits: no mask panel, the staging panel is write-only
            const uint32_t w = mw[p0 / 32];
            float4 bq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bq[c] = has_bias ? bias4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float* slot = stg + lane * 32 + ((c ^ (lane & 7)) << 2);
              const float bv[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float x = fmaf(acc[p0 + 4 * c + k], s_out, bv[k]);
                o[k] = ((w >> (4 * c + k)) & 1u) ? x : neg_slope * x;
              }
              tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))));
              *reinterpret_cast<float4*>(slot) = make_float4(o[0], o[1], o[2], o[3]);
            }
          } else if (g.bits_out != nullptr) {
            // forward launch that also leaves the sign bits of what it stores (relu / leaky outputs)
            uint32_t w = 0u;
            float4 bq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bq[c] = bias4[c];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float* slot = stg + lane * 32 + ((c ^ (lane & 7)) << 2);
              const float bv[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float x = fmaf(acc[p0 + 4 * c + k], s_out, bv[k]);
                const bool pos = x > 0.f;
                o[k] = pos ? x : (relu ? 0.f : slope * x);
                if (pos) w |= 1u << (4 * c + k);
              }
              tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))));
              *reinterpret_cast<float4*>(slot) = make_float4(o[0], o[1], o[2], o[3]);
            }
            const int col = n0 + col0 + p0;
            if (rvalid && col < g.N) g.bits_out[(roff + col) >> 5] = w;
          } else
          // operands of the whole round first (8 independent 16-byte loads in flight), then arithmetic and panel writes: the
          // compiler does not move a shared-memory load above an earlier shared-memory store, so a load inside the write loop
          // costs its full latency in every iteration
          if (masked) {
            mbar_wait(smem_u32(bar_mask), mround & 1);
            ++mround;
            float4 mq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) mq[c] = *reinterpret_cast<const float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2));
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float* slot = stg + lane * 32 + ((c ^ (lane & 7)) << 2);
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_bias) b4 = bias4[c];
              const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, mk[4] = {mq[c].x, mq[c].y, mq[c].z, mq[c].w};
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float x = fmaf(acc[p0 + 4 * c + k], s_out, bv[k]);
                o[k] = mk[k] > 0.f ? x : neg_slope * x;
              }
              tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))));
              *reinterpret_cast<float4*>(slot) = make_float4(o[0], o[1], o[2], o[3]);
            }
          } else {
            float4 bq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bq[c] = bias4[c];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float* slot = stg + lane * 32 + ((c ^ (lane & 7)) << 2);
              const float bv[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float x = fmaf(acc[p0 + 4 * c + k], s_out, bv[k]);
                o[k] = fmaxf(x, relu ? 0.f : slope * x);          // relu / leaky (slope < 1) / identity (slope = 1)
              }
              tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))));
              *reinterpret_cast<float4*>(slot) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          run_max = fmaxf(run_max, rvalid ? tmax : 0.f);           // rows past the tile's box hold stale operands
#ifdef TC3_TIMING
          const long long _r1 = clock64();
          w4 += _r1 - _r0;                                // ... + bias / activation / panel writes
#endif
          fence_async_smem();
#ifdef TC3_TIMING
          const long long _r2 = clock64();
          w5 += _r2 - _r1;                                // proxy fence
#endif
          asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32 + 32) : "memory");       // panels complete: the store warp takes over
#ifdef TC3_TIMING
          w6 += clock64() - _r2;                          // panel-complete barrier
#endif
        }
      } else if (g.vec_store) {
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
          // pass 1: addresses + all activation-mask loads of the panel in flight together
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
              run_max = fmaxf(run_max, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) if (colv + k < g.N) { g.C[off[i8] + k] = ov[k]; run_max = fmaxf(run_max, fabsf(ov[k])); }
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
                run_max = fmaxf(run_max, fabsf(x));
              }
            }
          }
        }
      }
      T3_SECTION_END(w2);
    }
    T3_ROLE_END(4, warp == Cfg::EPI0);
    if (g.amax_out != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
      if (lane == 0 && run_max > 0.f) atomicMax(reinterpret_cast<unsigned int*>(g.amax_out), __float_as_uint(run_max));
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
struct Tc3Maps { CUtensorMap a, bhi, blo, a2, c, c2, m, m2; };

template <int BN>
static int launch3(const Tc3Maps& m, const Tc3Args& g, dim3 grid, cudaStream_t s) {
  using Cfg = T3Cfg<BN>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_index()];
  if (!attr_done) {
    DDRL_CUDA(cudaFuncSetAttribute(tc3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  static const bool no_bres = [] { const char* e = getenv("DDRL_TC3_NO_BRES"); return e && e[0] == '1'; }();
  Tc3Args ga = g;
  ga.b_resident = (!no_bres && g.n_tiles == 1 && g.kb_total >= 1 && Cfg::STAGES % g.kb_total == 0) ? 1 : 0;
  tc3_kernel<BN><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(m.a, m.bhi, m.blo, m.a2, m.c, m.c2, m.m, m.m2, ga);
  prof_work(2.0 * g.M * (double)g.N * g.K * (g.tap.work_scale > 0.f ? g.tap.work_scale : 1.f));   // algorithmic flops
  if (g_prof_on && g_prof_shapes) {
    char nm[96];
    snprintf(nm, sizeof(nm), "%s3[fwd,M=%d,N=%d,K=%d,g=%d]", g.tap.mode ? "conv_tc" : "gemm_tc", g.M, g.N, g.K,
             (int)(grid.x * grid.y * grid.z));
    DDRL_LAUNCHED(prof_intern(nm));
    return DDRL_OK;
  }
  DDRL_LAUNCHED("tc3_kernel");
  return DDRL_OK;
}

static int launch3_bn(int bn, const Tc3Maps& m, const Tc3Args& g, dim3 grid, cudaStream_t s) {
  switch (bn) {
    case 128: return launch3<128>(m, g, grid, s);
    case 64: return launch3<64>(m, g, grid, s);
    default: return launch3<32>(m, g, grid, s);
  }
}
// Forward outputs (no activation mask, one pixel class) leave through TMA tile stores: ~120 instead of ~700 instructions
// per warp and 32 x 32 panel -- the N = 64 conv kernels issue 2.5 of 4 instructions per clock, so epilogue instructions are
// the currency (profiles/r2u_*).  The first version of this path (profiles/r2f_*) lost 10-15 % to per-element bias loads that
// compiled to 32 serialised LDG / branch blocks and to waiting for the TMA read-out inside the tile.  DDRL_TC3_TMA_STORE=0
// selects the per-warp coalesced stores.
static const bool g_t3_tma_store = [] { const char* e = getenv("DDRL_TC3_TMA_STORE"); return !(e && e[0] == '0'); }();
// ... and so do data gradients: the activation mask travels INTO the staging panels by TMA (same boxes as the stores), the
// fused stride-parity gradient stores one strided box per pixel class.  DDRL_TC3_TMA_DGRAD=0: per-warp mask loads + stores.
static const bool g_t3_tma_dgrad = [] { const char* e = getenv("DDRL_TC3_TMA_DGRAD"); return !(e && e[0] == '0'); }();

static inline int pick_bn3(int N) { return N > 64 ? 128 : (N > 32 ? 64 : 32); }
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// pre-split weight operand [N, ldb16 halfs] (K-major): box = 64 halfs x bn rows, 128B swizzle
static int make_map_b16(CUtensorMap* m, const void* base, int K, int N, int ldb16, int bn) {
  const unsigned long long dims[2] = {(unsigned long long)K, (unsigned long long)N};
  const unsigned long long strides[1] = {(unsigned long long)ldb16 * 2};
  const unsigned box[2] = {64u, (unsigned)bn}, estr[2] = {1u, 1u};
  return tc_encode_tiled(m, true, 2, base, dims, strides, box, estr, true);
}

// 4-D map over one fp16 plane of a pre-split NHWC tensor [Bn, H, W, Ctot]: box = 64 channels x nx pixels (every sx-th) x ny rows
// (every sy-th) x nb images, 128B swizzle: the box lands as rows of 64 halfs = the K-major (forward) / MN-major (weight gradient)
// fp16 operand tile
static int make_map_nhwc16(CUtensorMap* m, const void* plane, const ConvOp& o, int nx, int ny, int nb) {
  const unsigned long long dims[4] = {(unsigned long long)o.Ctot, (unsigned long long)o.Win, (unsigned long long)o.Hin, (unsigned long long)o.Bn};
  const unsigned long long strides[3] = {(unsigned long long)o.Ctot * 2, (unsigned long long)o.Win * o.Ctot * 2,
                                         (unsigned long long)o.Hin * o.Win * o.Ctot * 2};
  const unsigned box[4] = {64u, (unsigned)((nx - 1) * o.sx + 1), (unsigned)((ny - 1) * o.sy + 1), (unsigned)nb};
  const unsigned estr[4] = {1u, (unsigned)o.sx, (unsigned)o.sy, 1u};
  return tc_encode_tiled(m, true, 4, plane, dims, strides, box, estr, true);
}
static inline bool presplit_ok(const ConvOp& o, const void* hi, const void* lo) {
  return hi && lo && al16(hi) && al16(lo) && o.Cin % 64 == 0 && o.Ctot % 8 == 0 && o.c_off % 8 == 0;
}

// x [n] fp32 -> fp16 planes hi [n] | lo' [n] of x * t3_scale(*amax): the split every tc3 consumer of x would otherwise repeat per
// use (forward taps, weight gradient), done once.  8 elements per thread and step.
__global__ void __launch_bounds__(256) presplit_kernel(const float* __restrict__ x, long long n8, const float* __restrict__ amax,
                                                       uint4* __restrict__ hi, uint4* __restrict__ lo) {
  float sA, sA_inv;
  t3_scale(__ldg(amax), sA, sA_inv);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x4 + 2 * i), b = __ldg(x4 + 2 * i + 1);
    uint4 h, l;
    t3_split2(a.x, a.y, sA, h.x, l.x); t3_split2(a.z, a.w, sA, h.y, l.y);
    t3_split2(b.x, b.y, sA, h.z, l.z); t3_split2(b.z, b.w, sA, h.w, l.w);
    hi[i] = h; lo[i] = l;
  }
}
int tc3_presplit(const float* x, long long n, const float* amax, void* hi, void* lo, cudaStream_t s) {
  if (!x || !amax || !hi || !lo || n < 0 || n % 8 != 0 || !al16(x) || !al16(hi) || !al16(lo)) return DDRL_E_ARG;
  if (n == 0) return DDRL_OK;
  const long long n8 = n / 8;
  const int blocks = (int)std::min<long long>((n8 + 255) / 256, 16LL * kNumSMs);
  presplit_kernel<<<blocks, 256, 0, s>>>(x, n8, amax, reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
  DDRL_LAUNCHED("presplit_kernel");
  return DDRL_OK;
}

// ---------------------------------------------------------------- sign-bit tensors of activations (see Tc3Args::bits_out)
// The net registers a bit tensor per activation buffer whose derivative a data gradient will need.  `fresh` says that the
// buffer's CURRENT contents were written by a launch that also wrote the bits: a consumer only trusts fresh bits, and any
// tc3 launch that writes the buffer through another epilogue path clears the flag.
struct SignBits { const float* base; size_t elems; unsigned int* bits; bool fresh; };
static std::vector<SignBits> g_signbits;
static std::mutex g_signbits_mu;
void tc3_signbits_register(const float* base, size_t elems, unsigned int* bits) {
  std::lock_guard<std::mutex> lk(g_signbits_mu);
  for (auto& e : g_signbits)
    if (e.base == base) { e.elems = elems; e.bits = bits; e.fresh = false; return; }
  g_signbits.push_back({base, elems, bits, false});
}
void tc3_signbits_unregister(const float* base) {
  std::lock_guard<std::mutex> lk(g_signbits_mu);
  for (size_t i = 0; i < g_signbits.size(); ++i)
    if (g_signbits[i].base == base) { g_signbits.erase(g_signbits.begin() + i); return; }
}
// producer side: `out` is exactly a registered buffer, written densely with row length N (a multiple of 32) through the TMA
// epilogue of a relu / leaky launch -> its bit tensor (and the buffer counts as fresh); otherwise nullptr (and stale)
static unsigned int* signbits_for_output(const float* out, bool eligible) {
  std::lock_guard<std::mutex> lk(g_signbits_mu);
  for (auto& e : g_signbits)
    if (out >= e.base && out < e.base + e.elems) {
      e.fresh = eligible && out == e.base;
      return e.fresh ? e.bits : nullptr;
    }
  return nullptr;
}
// consumer side: `mask` points into a registered buffer with fresh bits, at a 32-element boundary -> the word that holds the
// bit of element mask[0]
static const unsigned int* signbits_for_mask(const float* mask) {
  std::lock_guard<std::mutex> lk(g_signbits_mu);
  if (!mask) return nullptr;
  for (auto& e : g_signbits)
    if (mask >= e.base && mask < e.base + e.elems) {
      const size_t off = (size_t)(mask - e.base);
      return (e.fresh && off % 32 == 0) ? e.bits + off / 32 : nullptr;
    }
  return nullptr;
}

bool tc3_gemm_supported(int M, int N, int K, const float* A, int lda, const void* Bhi, const void* Blo, int ldb16) {
  if (M < 1 || N < 1 || K < 1) return false;
  if (!al16(A) || !al16(Bhi) || !al16(Blo) || lda % 4 != 0 || ldb16 % 8 != 0) return false;
  return true;
}

// C[M,N] = epi(A[M,K] . B[N,K]^T)   B = pre-split fp16 hi / lo' rows of ldb16 halfs (t3 split of the fp32 weights)
int tc3_gemm(int M, int N, int K, const float* A, int lda, const void* Bhi, const void* Blo, int ldb16, const float* amax_a,
             const float* amax_b, float* C, int ldc, const float* bias, int act, const float* mask, float* amax_out,
             cudaStream_t s) {
  if (!tc3_gemm_supported(M, N, K, A, lda, Bhi, Blo, ldb16)) return DDRL_E_UNSUPPORTED;
  if ((act >= 3 && !mask) || !amax_a || !amax_b) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn3(N);
  Tc3Maps mp;
  r = tc_make_map(&mp.a, A, K, M, lda, T3_BM, false);
  if (r == DDRL_OK) r = make_map_b16(&mp.bhi, Bhi, K, N, ldb16, bn);
  if (r == DDRL_OK) r = make_map_b16(&mp.blo, Blo, K, N, ldb16, bn);
  if (r != DDRL_OK) return r;
  mp.a2 = mp.a;
  Tc3Args g;
  memset(&g, 0, sizeof(g));
  g.C = C; g.bias = bias; g.mask = mask; g.act = act;
  g.M = M; g.N = N; g.K = K; g.sCm = ldc; g.sCn = 1;
  g.kb_total = ceil_div(K, T3_BK); g.kb_per_split = g.kb_total;
  g.m_tiles = ceil_div(M, T3_BM); g.n_tiles = ceil_div(N, bn);
  g.amax_a = amax_a; g.amax_b = amax_b; g.amax_out = amax_out;
  g.vec_store = (ldc % 4 == 0 && al16(C) && (!mask || al16(mask))) ? 1 : 0;
  const bool tma_ok = g.vec_store && g_t3_tma_store && (act < 3 || g_t3_tma_dgrad);
  // data gradient whose activation has fresh sign bits (and a word-aligned layout): the mask is 1/32 of the bytes
  const unsigned int* mbits = (tma_ok && act >= 3 && ldc % 32 == 0) ? signbits_for_mask(mask) : nullptr;
  if (tma_ok) {
    // output tiles leave through TMA: boxes of 128 rows x 32 columns of C [M, N]; the mask of a data gradient (same
    // layout as C) arrives through the same boxes
    const unsigned long long dims[2] = {(unsigned long long)N, (unsigned long long)M};
    const unsigned long long strides[1] = {(unsigned long long)ldc * 4};
    const unsigned box[2] = {32u, (unsigned)T3_BM}, estr[2] = {1u, 1u};
    int rs = tc_encode_tiled(&mp.c, false, 2, C, dims, strides, box, estr, true);
    if (rs == DDRL_OK && act >= 3 && !mbits) rs = tc_encode_tiled(&mp.m, false, 2, mask, dims, strides, box, estr, true);
    if (rs == DDRL_OK) { g.tma_store = 1; g.tma_mask = (act >= 3 && !mbits) ? 1 : 0; g.bits_in = mbits; }
  }
  g.bits_out = signbits_for_output(C, g.tma_store && (act == 1 || act == 2) && ldc == N && N % 32 == 0);
  if (!g.tma_store) mp.c = mp.a;
  mp.c2 = mp.c;
  if (!g.tma_mask) mp.m = mp.c;
  mp.m2 = mp.m;
  dim3 grid(std::min(g.m_tiles * g.n_tiles, kNumSMs), 1, 1);
  return launch3_bn(bn, mp, g, grid, s);
}

int tc3_conv_fwd(const ConvOp& o, const void* Whi, const void* Wlo, int ldw16, int N, const float* amax_a, const float* amax_b,
                 const float* bias, int act, const float* mask, float* out, long long osb, long long osy, long long osx,
                 float* amax_out, cudaStream_t s, const TcTap* cls, const void* a_hi16, const void* a_lo16) {
  if (!conv_tc_supported(o, false) || N < 1 || ldw16 % 8 != 0 || !al16(Whi) || !al16(Wlo)) return DDRL_E_UNSUPPORTED;
  if ((act >= 3 && !mask) || !amax_a || !amax_b) return DDRL_E_ARG;
  const bool ps = a_hi16 != nullptr || a_lo16 != nullptr;
  if (ps && (!presplit_ok(o, a_hi16, a_lo16) || act >= 3)) return DDRL_E_UNSUPPORTED;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = pick_bn3(N);
  Tc3Args g;
  memset(&g, 0, sizeof(g));
  tc_tap_common(g.tap, o, T3_BM, true);
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
  Tc3Maps mp;
  const bool ph2 = g.tap.ny2 > 0;
  if (ps) {
    g.presplit = 1;
    g.tap.cpb = o.Cin / 64; g.tap.nslices = o.KH * o.KW * g.tap.cpb;     // 64-channel slices, one per K block
    r = make_map_nhwc16(&mp.a, a_hi16, o, o.Xn, g.tap.ny, g.tap.nb);
    if (r == DDRL_OK && ph2) r = make_map_nhwc16(&mp.a2, a_hi16, o, o.Xn, g.tap.ny2, g.tap.nb2);
  } else {
    r = tc_make_map_nhwc(&mp.a, o, o.Xn, g.tap.ny, g.tap.nb, false);
    if (r == DDRL_OK && ph2) r = tc_make_map_nhwc(&mp.a2, o, o.Xn, g.tap.ny2, g.tap.nb2, false);
  }
  if (r == DDRL_OK) r = make_map_b16(&mp.bhi, Whi, K, N, ldw16, bn);
  if (r == DDRL_OK) r = make_map_b16(&mp.blo, Wlo, K, N, ldw16, bn);
  if (r != DDRL_OK) return r;
  if (!ph2) mp.a2 = mp.a;
  g.C = out; g.bias = bias; g.mask = mask; g.act = act;
  g.M = o.Bn * o.Yn * o.Xn; g.N = N; g.K = K; g.sCm = 0; g.sCn = 1;
  g.kb_total = ps ? g.tap.nslices : (g.tap.nslices + 1) / 2; g.kb_per_split = g.kb_total;
  g.m_tiles = ceil_div(o.Bn, g.tap.nb) * g.tap.tpi + (ph2 ? ceil_div(o.Bn, g.tap.nb2) : 0);
  g.n_tiles = ceil_div(N, bn);
  g.amax_a = amax_a; g.amax_b = amax_b; g.amax_out = amax_out;
  g.vec_store = (osb % 4 == 0 && osy % 4 == 0 && osx % 4 == 0 && al16(out) && (!mask || al16(mask))) ? 1 : 0;
  const bool classes = g.tap.ncls > 1;
  // fused stride-parity gradient through TMA: every 32-column panel belongs to ONE pixel class (cls_cols % 32 == 0), whose
  // pixels (out_s*y + iy, out_s*x + ix) are a strided box of the output; needs the dense [B, out_H, out_W, ct] layout
  const long long ct = classes && g.tap.out_s > 0 ? osx / g.tap.out_s : 0;
  const bool classes_ok = classes && g_t3_tma_dgrad && g.tap.cls_cols % 32 == 0 && g.tap.out_W > 1 && g.tap.ncls == g.tap.out_s * g.tap.out_s &&
                          ct > 0 && osx == ct * g.tap.out_s && osy == (long long)g.tap.out_s * g.tap.out_W * ct &&
                          osb == (long long)g.tap.out_H * g.tap.out_W * ct;
  const bool tma_ok = g.vec_store && g_t3_tma_store && (act < 3 || g_t3_tma_dgrad) && (!classes || classes_ok);
  const unsigned int* mbits = (tma_ok && act >= 3 && osb % 32 == 0 && osy % 32 == 0 && osx % 32 == 0 && (!classes || g.tap.cls_cols % 32 == 0))
                                  ? signbits_for_mask(mask) : nullptr;
  if (mbits && classes)
    for (int q = 0; q < g.tap.ncls; ++q) if (g.tap.cls_off[q] % 32 != 0) mbits = nullptr;
  if (tma_ok) {
    // output tiles leave through TMA: the store box (32 columns x Xn pixels x ny rows x nb images of out[b, y, x, n])
    // mirrors the load-side pixel box, one map per tiling phase; the mask of a data gradient arrives through the same boxes
    const int es = classes ? g.tap.out_s : 1;
    const unsigned long long dims[4] = {(unsigned long long)(classes ? g.tap.cls_cols : N), (unsigned long long)(classes ? g.tap.out_W : o.Xn),
                                        (unsigned long long)(classes ? g.tap.out_H : o.Yn), (unsigned long long)o.Bn};
    const unsigned long long strides[3] = {(unsigned long long)(osx / es) * 4, (unsigned long long)(osy / es) * 4, (unsigned long long)osb * 4};
    const unsigned estr[4] = {1u, (unsigned)es, (unsigned)es, 1u};
    auto boxdim = [&](int n) { return (unsigned)((n - 1) * es + 1); };
    const unsigned box1[4] = {32u, boxdim(o.Xn), boxdim(g.tap.ny), (unsigned)g.tap.nb};
    const unsigned box2[4] = {32u, boxdim(o.Xn), boxdim(g.tap.ny2 > 0 ? g.tap.ny2 : 1), (unsigned)(g.tap.nb2 > 0 ? g.tap.nb2 : 1)};
    int rs = tc_encode_tiled(&mp.c, false, 4, out, dims, strides, box1, estr, true);
    if (rs == DDRL_OK && ph2) rs = tc_encode_tiled(&mp.c2, false, 4, out, dims, strides, box2, estr, true);
    if (rs == DDRL_OK && act >= 3 && !mbits) {
      rs = tc_encode_tiled(&mp.m, false, 4, mask, dims, strides, box1, estr, true);
      if (rs == DDRL_OK && ph2) rs = tc_encode_tiled(&mp.m2, false, 4, mask, dims, strides, box2, estr, true);
    }
    if (rs == DDRL_OK) { g.tma_store = 1; g.tma_mask = (act >= 3 && !mbits) ? 1 : 0; g.bits_in = mbits; }
  }
  g.bits_out = signbits_for_output(out, g.tma_store && !classes && (act == 1 || act == 2) && N % 32 == 0 && osx == N &&
                                            osy == (long long)o.Xn * N && osb == (long long)o.Yn * o.Xn * N);
  if (!g.tma_store) mp.c = mp.a;
  if (!g.tma_store || !ph2) mp.c2 = mp.c;
  if (!g.tma_mask) mp.m = mp.c;
  if (!g.tma_mask || !ph2) mp.m2 = mp.m;
  if (ps) {                                       // the mask slots carry the lo' plane
    r = make_map_nhwc16(&mp.m, a_lo16, o, o.Xn, g.tap.ny, g.tap.nb);
    if (r == DDRL_OK && ph2) r = make_map_nhwc16(&mp.m2, a_lo16, o, o.Xn, g.tap.ny2, g.tap.nb2);
    if (r != DDRL_OK) return r;
    if (!ph2) mp.m2 = mp.m;
  }
  dim3 grid(std::min(g.m_tiles * g.n_tiles, kNumSMs), 1, 1);
  return launch3_bn(bn, mp, g, grid, s);
}

// ================================================================ weight gradient
// C[m, n] += sum_r A[r, m] * B[r, n]     A = activations [rows, k] (plain 2-D or tap boxes), B = dy [rows, n], both raw fp32.
// The GEMM's M axis is the im2col K axis, so A must be transposed: free, because each splitter thread gathers one k
// column of the landed [64 rows x 32 k] slice into its TMEM lane (pairs of rows packed as fp16x2).  The dy tile is
// converted by the same warps into MN-major fp16 hi / lo' tiles (64 rows x 128 B, 128B swizzle) that belong to the TMEM
// slot, so the raw TMA stage is released as soon as it has been read (before the MMAs run).
//   barriers: full[S] TMA landed | empty[S] splitter warps done with the raw stage | aready[SA] slot + B tiles written |
//             afree[SA] both MMA streams retired the slot | mfull/mfree[2], cfull as in the forward kernel
template <int BN>
struct T3WCfg {
  static_assert(BN == 64 || BN == 128, "weight-gradient tiles are 64 or 128 columns");
  static constexpr int R_SUB = T3_BK * 128;                      // one raw slab: 64 rows x 32 fp32 = 8 KB
  static constexpr int A_SUB = R_SUB;                            // one A slice: 64 rows x 32 k fp32, or 64 pixels x 64 halfs of a pre-split plane
  static constexpr int A_BYTES = 4 * A_SUB;
  static constexpr int BRAW_BYTES = (BN / 32) * R_SUB;           // dy tile raw: slabs of 64 rows x 32 n
  static constexpr int STAGE_BYTES = A_BYTES + BRAW_BYTES;
  static constexpr int B16_TILE = T3_BK * 128;                   // 64 rows x 64 n halfs = 8 KB
  static constexpr int B16_BYTES = 2 * (BN / 64) * B16_TILE;     // [hi groups | lo groups]
  static constexpr bool FOLD = BN <= 64;
  static constexpr int SA = BN <= 64 ? 3 : 2;
  static constexpr int STAGES = BN <= 64 ? 3 : 2;
  static constexpr int NEPI = 8;
  static constexpr int NSG = 2;
  static constexpr int EPI0 = 4 + 4 * NSG;
  static constexpr int THREADS = (EPI0 + NEPI) * 32;
  static constexpr int COLS = BN / 2;
  static constexpr int TM_MAIN0 = 0, TM_MAIN1 = FOLD ? 2 * BN : BN, TM_CORR = FOLD ? 4 * BN : 2 * BN, TM_A = FOLD ? 5 * BN : 3 * BN;
  static constexpr int TMEM_COLS = 512;
  static constexpr int NBARS = 2 * STAGES + 2 * SA + 6;
  static constexpr int B16_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = B16_OFF + SA * B16_BYTES;
  static constexpr int CS_OFF = BAR_OFF + 256;                    // deterministic mode: [NEPI][BN] bias-gradient partials
  static constexpr int SMEM = 1024 + CS_OFF + NEPI * BN * 4;
  static_assert(NBARS * 8 + 16 <= 256, "barrier block");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(TM_A + SA * 64 <= 512, "TMEM budget");
};

template <int BN>
__global__ void __launch_bounds__(T3WCfg<BN>::THREADS, 1)
tc3_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmAl2, Tc3Args g) {
  using Cfg = T3WCfg<BN>;
  constexpr int S = Cfg::STAGES, SA = Cfg::SA;
  extern __shared__ uint8_t smem_dyn[];
  // 1024-byte alignment by POINTER arithmetic on the __shared__ array: an integer round-trip hides the address space from
  // the compiler, which then emits generic LD / ST (long-scoreboard, L1TEX path) for every shared-memory access below
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + S;
  uint64_t* bar_aready = bars + 2 * S;
  uint64_t* bar_afree = bars + 2 * S + SA;
  uint64_t* bar_mfull = bars + 2 * S + 2 * SA;
  uint64_t* bar_mfree = bar_mfull + 2;
  uint64_t* bar_cfull = bar_mfull + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcTap& tp = g.tap;
  const bool tapA = tp.mode != 0;
  // pre-split activation operand (fp16 hi / lo' planes): the landed boxes are the MMA's MN-major A tiles, so the stage is
  // released by the MMAs (2 commits) and the conversion warps, and the gather warps have nothing to do
  const bool ps = g.presplit != 0;

  if (tapA) {
    // pixel-box K blocks shorter than 64 rows: the rows no box covers must read as zero
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < S * Cfg::STAGE_BYTES / 16; i += Cfg::THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  if (warp == 0 && lane == 0) {
    // a stage is consumed / a slot is produced by the 4 gather warps of one splitter group AND the NEPI conversion warps
    for (int s = 0; s < S; ++s) { mbar_init(smem_u32(bar_full + s), 1); mbar_init(smem_u32(bar_empty + s), (ps ? 2 : 4) + Cfg::NEPI); }
    for (int a = 0; a < SA; ++a) { mbar_init(smem_u32(bar_aready + a), (ps ? 0 : 4) + Cfg::NEPI); mbar_init(smem_u32(bar_afree + a), 2); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(bar_mfull + b), 1); mbar_init(smem_u32(bar_mfree + b), Cfg::NEPI); }
    mbar_init(smem_u32(bar_cfull), 1);
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

  // one unit per CTA: (blockIdx.x = M block, blockIdx.y = N tile, blockIdx.z = K split)
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int nkb = min(g.kb_total, kb0 + g.kb_per_split) - kb0;
  const int ksteps = tapA ? tp.kpad / 16 : 4;
  const int m0 = (int)blockIdx.x * T3_BM, n0 = (int)blockIdx.y * BN;

  if (warp == 0) {
    // ============================================================ TMA producer
    int sl_c[4] = {0, 0, 0, 0}, sl_x[4] = {0, 0, 0, 0}, sl_y[4] = {0, 0, 0, 0}, na = 0;
    if (tapA) {
      // (pre-split: slices are 64 channels wide, two per M block; tp.cpb / tp.nslices count those)
      const int spb = ps ? 2 : 4, sw = ps ? 64 : 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sl = (int)blockIdx.x * spb + j;
        if (j < spb && sl < tp.nslices) {
          const int tap = sl / tp.cpb, cc = sl - tap * tp.cpb;
          const int kh = tap / tp.KW, kw = tap - kh * tp.KW;
          sl_c[j] = tp.c_off + cc * sw; sl_x[j] = kw - tp.px; sl_y[j] = kh - tp.py;
          na = j + 1;
        }
      }
    }
    int pb = 0, pj = 0;
    if (tapA) { pb = kb0 / tp.tpi; pj = kb0 - pb * tp.tpi; }
    T3_ROLE_BEGIN
    for (int i = 0; i < nkb; ++i) {
      const uint32_t s = i % S;
      T3_WAITL(smem_u32(bar_empty + s), ((i / S) & 1) ^ 1, w0);
      if (elect_one()) {
        const uint32_t full = smem_u32(bar_full + s);
        const uint32_t a_dst = smem_u32(smem) + s * Cfg::STAGE_BYTES, b_dst = a_dst + Cfg::A_BYTES;
        const int k = (kb0 + i) * T3_BK;
        if (!tapA) {
          mbar_expect_tx(full, Cfg::A_BYTES + Cfg::BRAW_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(&tmA, full, a_dst + j * Cfg::A_SUB, m0 + j * 32, k);
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmB, full, b_dst + j * Cfg::R_SUB, n0 + j * 32, k);
        } else if (tp.w2on) {
          // two pixel-box classes (64 pixels each): class 1 covers the columns right of wxw0
          const bool c1 = pj >= tp.wnb0;
          const int x0 = c1 ? tp.wxw0 : 0, yy0 = c1 ? (pj - tp.wnb0) * tp.wyh1 : pj * tp.wyh0;
          mbar_expect_tx(full, ((ps ? 2 * na : na) + BN / 32) * 64 * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < na) {
              tma_load_4d(c1 ? &tmA2 : &tmA, full, a_dst + j * Cfg::A_SUB, sl_c[j], x0 * tp.sx + sl_x[j], yy0 * tp.sy + sl_y[j], pb);
              if (ps) tma_load_4d(c1 ? &tmAl2 : &tmAl, full, a_dst + (2 + j) * Cfg::A_SUB, sl_c[j], x0 * tp.sx + sl_x[j], yy0 * tp.sy + sl_y[j], pb);
            }
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_4d(c1 ? &tmB2 : &tmB, full, b_dst + j * Cfg::R_SUB, n0 + j * 32, x0, yy0, pb);
        } else {
          // K blocks [0, tiles1): boxes of ny rows x nb images, tpi per image group; K blocks >= tiles1 (two-phase K tiling of
          // small maps): the images' remaining rows [y2, Yn), nb2 images per box -- 9x9 maps: 7 rows x 1 image + 2 rows x 3
          // images = 4 K blocks per 3 images instead of 6 (one of 63 and one of 18 pixels per image)
          const int kb = kb0 + i;
          const bool p2 = kb >= tp.tiles1;
          int bb0, yy0;
          if (!p2) { const int grp = kb / tp.tpi; bb0 = grp * tp.nb; yy0 = (kb - grp * tp.tpi) * tp.ny; }
          else { bb0 = (kb - tp.tiles1) * tp.nb2; yy0 = tp.y2; }
          mbar_expect_tx(full, ((ps ? 2 * na : na) + BN / 32) * (p2 ? tp.rows2 : tp.rows) * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < na) {
              tma_load_4d(p2 ? &tmA2 : &tmA, full, a_dst + j * Cfg::A_SUB, sl_c[j], sl_x[j], yy0 * tp.sy + sl_y[j], bb0);
              if (ps) tma_load_4d(p2 ? &tmAl2 : &tmAl, full, a_dst + (2 + j) * Cfg::A_SUB, sl_c[j], sl_x[j], yy0 * tp.sy + sl_y[j], bb0);
            }
          if (tp.nb == 1 && tp.ny2 == 0) {             // one image per box: dy rows are one run of pixels (3-D map)
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_3d(&tmB, full, b_dst + j * Cfg::R_SUB, n0 + j * 32, yy0 * tp.Xn, bb0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_4d(p2 ? &tmB2 : &tmB, full, b_dst + j * Cfg::R_SUB, n0 + j * 32, 0, yy0, bb0);
          }
        }
      }
      __syncwarp();
      if (tapA && ++pj == tp.tpi) { pj = 0; ++pb; }
    }
    T3_ROLE_END(5, true);
  } else if (warp == 1 || warp == 3) {
    // ============================================================ MMA issuers
    const bool chunk_role = warp == 1;
    // f16 x f16 -> f32, B MN-major (bit 16)
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
    const uint32_t b16_base = smem_u32(smem) + Cfg::B16_OFF;
    const uint32_t t_corr = tmem_base + Cfg::TM_CORR;
    uint32_t ch = 0;
    T3_ROLE_BEGIN
    for (int i = 0; i < nkb; ++i) {
      const uint32_t a = i % SA;
      const uint32_t buf = ch & 1;
      const bool first_in_chunk = (i % T3_CHUNK) == 0;
      const bool last_in_chunk = (i % T3_CHUNK) == T3_CHUNK - 1 || i == nkb - 1;
      if (chunk_role && first_in_chunk) T3_WAIT(smem_u32(bar_mfree + buf), ((ch >> 1) & 1) ^ 1, w0);
      T3_WAIT(smem_u32(bar_aready + a), (i / SA) & 1, w1);
      tc_fence_after();
      T3_SECTION_BEGIN;
      if (elect_one()) {
        // MN-major 128B-swizzled fp16 tiles: 64-column groups 8 KB apart (LBO), 8-row K groups 1 KB apart (SBO);
        // one k step = 16 rows = 2 KB
        const uint32_t b_hi = b16_base + a * Cfg::B16_BYTES, b_lo = b_hi + (BN / 64) * Cfg::B16_TILE;
        const uint32_t a_hi = tmem_base + Cfg::TM_A + a * 64, a_lo = a_hi + 32;
        const uint64_t dbh0 = umma_desc(b_hi, Cfg::B16_TILE, 1024, 2);
        if (ps) {
          // A from shared memory, MN-major (bit 15): the stage holds [hi slice 0 | hi slice 1 | lo' slice 0 | lo' slice 1], each
          // 64 pixels (k) x 64 channels (m) in the dy tiles' own layout: 64-channel groups 8 KB apart, one k step = 2 KB
          const uint32_t st_a = smem_u32(smem) + (i % S) * Cfg::STAGE_BYTES;
          const uint64_t dah0 = umma_desc(st_a, Cfg::A_SUB, 1024, 2), dal0 = umma_desc(st_a + 2 * Cfg::A_SUB, Cfg::A_SUB, 1024, 2);
          const uint32_t am = 1u << 15;
          if (chunk_role) {
            const uint32_t t_main = tmem_base + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (k4 >= ksteps) break;
              umma_f16_ss(t_main, dah0 + k4 * (2048 >> 4), dbh0 + k4 * (2048 >> 4), (Cfg::FOLD ? idesc2 : idesc) | am, (!first_in_chunk || k4 != 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(bar_empty + (i % S)));
            umma_commit(smem_u32(bar_afree + a));
            if (last_in_chunk) umma_commit(smem_u32(bar_mfull + buf));
          } else {
            const uint64_t dbl0 = umma_desc(b_lo, Cfg::B16_TILE, 1024, 2);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (k4 >= ksteps) break;
              umma_f16_ss(t_corr, dal0 + k4 * (2048 >> 4), dbh0 + k4 * (2048 >> 4), idesc | am, (i | k4) != 0 ? 1u : 0u);
              if (!Cfg::FOLD) umma_f16_ss(t_corr, dah0 + k4 * (2048 >> 4), dbl0 + k4 * (2048 >> 4), idesc | am, 1u);
            }
            umma_commit(smem_u32(bar_empty + (i % S)));
            umma_commit(smem_u32(bar_afree + a));
            if (i == nkb - 1) umma_commit(smem_u32(bar_cfull));
          }
        } else if (chunk_role) {
          const uint32_t t_main = tmem_base + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (k4 >= ksteps) break;
            umma_f16_ts(t_main, a_hi + k4 * 8, dbh0 + k4 * (2048 >> 4), Cfg::FOLD ? idesc2 : idesc, (!first_in_chunk || k4 != 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(bar_afree + a));
          if (last_in_chunk) umma_commit(smem_u32(bar_mfull + buf));
        } else {
          const uint64_t dbl0 = umma_desc(b_lo, Cfg::B16_TILE, 1024, 2);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (k4 >= ksteps) break;
            umma_f16_ts(t_corr, a_lo + k4 * 8, dbh0 + k4 * (2048 >> 4), idesc, (i | k4) != 0 ? 1u : 0u);
            if (!Cfg::FOLD) umma_f16_ts(t_corr, a_hi + k4 * 8, dbl0 + k4 * (2048 >> 4), idesc, 1u);
          }
          umma_commit(smem_u32(bar_afree + a));
          if (i == nkb - 1) umma_commit(smem_u32(bar_cfull));
        }
      }
      __syncwarp();
      T3_SECTION_END(w2);
      if (last_in_chunk) ++ch;
    }
    T3_ROLE_END(6, chunk_role);
  } else if (warp >= 4 && warp < Cfg::EPI0) {
    // ============================================================ splitters: A columns (transposing gather) -> TMEM
    // Measured (profiles/r2p_tc3_roles.txt): with the dy conversion in these warps they were busy 87 % of the kernel while the
    // MMA issuers waited 68 % and the epilogue warps 96 %: the conversion now runs in the epilogue warps.
    const int q = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    float sA, sA_inv;
    t3_scale(__ldg(g.amax_a), sA, sA_inv);
    T3_ROLE_BEGIN
    for (int i = ps ? nkb : 0; i < nkb; ++i) {                     // (pre-split operand: nothing to gather)
      const int s = i % S, a = i % SA;
      if ((i % Cfg::NSG) != grp) {
        // follow every phase of the stage barrier (see the forward kernel): with S = 3 stages and 2 groups a group meets a
        // stage on every other use, and its parity wait would be satisfied by the phase before the one it skipped
        if (S % Cfg::NSG != 0) T3_WAIT(smem_u32(bar_full + s), (i / S) & 1, w0);
        continue;
      }
      T3_WAIT(smem_u32(bar_full + s), (i / S) & 1, w0);
      // the TMEM slot was last read by K block i - SA
      if (i >= SA) T3_WAIT(smem_u32(bar_afree + a), ((i / SA) - 1) & 1, w1);
      tc_fence_after();
      T3_SECTION_BEGIN;
      const uint8_t* st = smem + s * Cfg::STAGE_BYTES;
      // A: thread = k column `lane` of slice q; gathers it over the 64 rows of the box, pairs of rows -> fp16x2
      const uint32_t ta = tmem_base + t_lane + Cfg::TM_A + a * 64;
      const uint8_t* sp = st + q * Cfg::A_SUB + (lane & 3) * 4;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int pp = hf * 32 + 2 * p;
          const float x0 = *reinterpret_cast<const float*>(sp + pp * 128 + (((lane >> 2) ^ (pp & 7)) << 4));
          const float x1 = *reinterpret_cast<const float*>(sp + (pp + 1) * 128 + (((lane >> 2) ^ ((pp + 1) & 7)) << 4));
          // always the scaled form (s = 1 multiplies exactly): a per-pair `unit scale` branch made every pair its own basic
          // block -- 2 LDS, branch, a 6-deep dependent conversion chain, branch -- with no overlap between pairs: ~80 clk per
          // pair, 5000 clk per K block, the splitters 85 % busy and the tensor pipe 18 % (profiles/r2p_tc3_roles.txt)
          t3_split2<true>(x0, x1, sA, hi[p], lo[p]);
        }
#ifdef TC3_TIMING
        { const long long _q0 = clock64();
#endif
        tmem_st16(ta + hf * 16, hi);
        tmem_st16(ta + 32 + hf * 16, lo);
#ifdef TC3_TIMING
          w3 += clock64() - _q0; }
#endif
      }
#ifdef TC3_TIMING
      const long long _q1 = clock64();
#endif
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(bar_empty + s)); mbar_arrive(smem_u32(bar_aready + a)); }
#ifdef TC3_TIMING
      w3 += clock64() - _q1;
#endif
      T3_SECTION_END(w2);
    }
    T3_ROLE_END(7, warp == 4);
  } else if (warp >= Cfg::EPI0) {
    // ============================================================ dy conversion + drain + atomic accumulate
    // Per K block these 8 warps turn the landed dy tile (fp32) into the MMA's fp16 hi / lo' tiles; after the last K block of
    // chunk c they drain chunk c - 1 (its MMAs only depend on conversions that are already done, so the wait cannot
    // deadlock, and the chunk MMA stream gets its buffer back one chunk ahead of needing it).
    const int e = warp - Cfg::EPI0;
    const int q = e & 3, half = e >> 2;
    const int et = e * 32 + lane;                               // conversion thread 0 .. 255
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const uint32_t col0 = half * Cfg::COLS;
    float sA, sA_inv, sB, sB_inv;
    t3_scale(__ldg(g.amax_a), sA, sA_inv);
    t3_scale(__ldg(g.amax_b), sB, sB_inv);
    // bias gradient: every conversion task of this thread covers the same 8 columns (256 threads, BN / 8 tasks per row),
    // so the column sums of the dy tiles ride along in 8 registers (only the CTAs of the first M block contribute)
    const bool do_colsum = g.colsum != nullptr && blockIdx.x == 0;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float acc[Cfg::COLS];
#pragma unroll
    for (int j = 0; j < Cfg::COLS; ++j) acc[j] = 0.f;
    const int nch = (nkb + T3_CHUNK - 1) / T3_CHUNK;
    T3_ROLE_BEGIN
    // BN = 128: 64 accumulator registers per thread stay live across the conversion loop, so the drain reads TMEM 16
    // columns at a time (32 would not fit the 96-register budget of a 640-thread CTA without spilling)
    constexpr int LDW = BN <= 64 ? 32 : 16;
    auto drain = [&](int c) {
      const int buf = c & 1;
      T3_WAITL(smem_u32(bar_mfull + buf), (c >> 1) & 1, w0);
      tc_fence_after();
#pragma unroll
      for (int j0 = 0; j0 < Cfg::COLS; j0 += LDW) {
        float v[LDW];
        const uint32_t ta = tmem_base + t_lane + (buf ? Cfg::TM_MAIN1 : Cfg::TM_MAIN0) + col0 + j0;
        if (LDW == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v);
#pragma unroll
        for (int j = 0; j < LDW; ++j) acc[j0 + j] += v[j];
        if (Cfg::FOLD) {
          if (LDW == 32) tmem_ld32(ta + BN, v); else tmem_ld16(ta + BN, v);
#pragma unroll
          for (int j = 0; j < LDW; ++j) acc[j0 + j] = fmaf(v[j], T3_LO_INV, acc[j0 + j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_mfree + buf));
    };
    constexpr int NT = T3_BK * (BN / 8) / (Cfg::NEPI * 32);       // conversion tasks per thread and K block
    int t_src[NT], t_dst[NT];
    // two-phase K tiling with boxes of different heights: a stage's rows past the current box still hold the previous K
    // block's dy pixels -- converted as zeros (the activation rows they meet are finite, so the products vanish)
    const bool mask_rows = tapA && !tp.w2on && tp.ny2 > 0 && tp.rows2 != tp.rows;
    int t_row[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int v = et + t * Cfg::NEPI * 32;
      const int r = v / (BN / 8), c8 = v - r * (BN / 8);
      t_row[t] = r;
      t_src[t] = (c8 >> 2) * Cfg::R_SUB + r * 128 + ((((c8 & 3) * 2) ^ (r & 7)) << 4);      // second chunk: this offset ^ 16
      t_dst[t] = (c8 >> 3) * Cfg::B16_TILE + r * 128 + (((c8 & 7) ^ (r & 7)) << 4);
    }
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S, a = i % SA;
      T3_WAIT(smem_u32(bar_full + s), (i / S) & 1, w3);
      // the fp16 dy tiles of slot a were last read by K block i - SA
      if (i >= SA) T3_WAIT(smem_u32(bar_afree + a), ((i / SA) - 1) & 1, w3);
      const uint8_t* st = smem + s * Cfg::STAGE_BYTES;
      uint8_t* b16 = smem + Cfg::B16_OFF + a * Cfg::B16_BYTES;
      const int rows_blk = (kb0 + i) >= tp.tiles1 ? tp.rows2 : tp.rows;
      // dy tile: task = (row r, 8 consecutive columns): 2 swizzled 16-byte fp32 chunks -> 1 chunk of hi + 1 chunk of lo'
      // (a thread's tasks sit at the same tile positions in every K block: offsets computed once, before the loop -- the
      // per-task index arithmetic was 12 % of every weight-gradient launch, profiles/r4i_*)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const uint8_t* src = st + Cfg::A_BYTES + t_src[t];
        float4 x0 = *reinterpret_cast<const float4*>(src);
        float4 x1 = *reinterpret_cast<const float4*>(st + Cfg::A_BYTES + (t_src[t] ^ 16));
        if (mask_rows && t_row[t] >= rows_blk) { x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0; }
        uint4 h, l;
        t3_split2(x0.x, x0.y, sB, h.x, l.x); t3_split2(x0.z, x0.w, sB, h.y, l.y);
        t3_split2(x1.x, x1.y, sB, h.z, l.z); t3_split2(x1.z, x1.w, sB, h.w, l.w);
        if (do_colsum) {
          cs[0] += x0.x; cs[1] += x0.y; cs[2] += x0.z; cs[3] += x0.w; cs[4] += x1.x; cs[5] += x1.y; cs[6] += x1.z; cs[7] += x1.w;
        }
        uint8_t* dst = b16 + t_dst[t];
        *reinterpret_cast<uint4*>(dst) = h;
        *reinterpret_cast<uint4*>(dst + (BN / 64) * Cfg::B16_TILE) = l;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(bar_empty + s)); mbar_arrive(smem_u32(bar_aready + a)); }
      if ((i % T3_CHUNK) == T3_CHUNK - 1 && i >= T3_CHUNK) drain(i / T3_CHUNK - 1);
    }
    // chunks not drained inside the loop: the last full-or-partial one, and the one before it when the tail was partial
    {
      const int drained = nkb >= 2 * T3_CHUNK ? nkb / T3_CHUNK - 1 : 0;      // chunks 0 .. drained - 1 are done
      for (int c = drained; c < nch; ++c) drain(c);
    }
    // deterministic mode (g.det_ctr): the K splits of one tile add in split order -- a turn counter per (M block, N tile) --
    // and the 8 conversion warps' bias-gradient partials meet in shared memory in warp order instead of by atomics
    unsigned int* turn = g.det_ctr ? g.det_ctr + (size_t)blockIdx.y * gridDim.x + blockIdx.x : nullptr;
    float* cs_red = reinterpret_cast<float*>(smem + Cfg::CS_OFF);
    if (do_colsum) {
      // threads with equal (et mod BN/8) hold partial sums of the same 8 columns: lanes l, l + BN/8, ... of a warp
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = cs[k];
#pragma unroll
        for (int o = 16; o >= BN / 8; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        cs[k] = v;
      }
      if (lane < BN / 8) {
        const int col = n0 + lane * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (turn) cs_red[e * BN + lane * 8 + k] = cs[k];
          else if (col + k < g.N) {
            const int c = col + k;
            atomicAdd((g.colsum2 != nullptr && c >= g.colsum_split) ? g.colsum2 + (c - g.colsum_split) : g.colsum + c, cs[k]);
          }
        }
      }
    }
    if (nkb > 0) {
      T3_WAIT(smem_u32(bar_cfull), 0, w1);
      tc_fence_after();
      T3_SECTION_BEGIN;
#pragma unroll
      for (int j0 = 0; j0 < Cfg::COLS; j0 += LDW) {
        float v[LDW];
        if (LDW == 32) tmem_ld32(tmem_base + t_lane + Cfg::TM_CORR + col0 + j0, v);
        else tmem_ld16(tmem_base + t_lane + Cfg::TM_CORR + col0 + j0, v);
#pragma unroll
        for (int j = 0; j < LDW; ++j) acc[j0 + j] = fmaf(v[j], T3_LO_INV, acc[j0 + j]);
      }
      if (turn) {
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32) : "memory");        // cs_red complete
        if (e == 0 && lane == 0) det_enter(turn, blockIdx.z);
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32) : "memory");
        if (do_colsum && et < BN && n0 + et < g.N) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < Cfg::NEPI; ++w) v += cs_red[w * BN + et];
          const int c = n0 + et;
          det_add(true, (g.colsum2 != nullptr && c >= g.colsum_split) ? g.colsum2 + (c - g.colsum_split) : g.colsum + c, v);
        }
      }
      const int r = q * 32 + lane;
      if (m0 + r < g.M) {
        float* crow = g.C + (long long)(m0 + r) * g.sCm;
#pragma unroll
        for (int j = 0; j < Cfg::COLS; ++j) {
          const int col = n0 + col0 + j;
          if (col < g.N) det_add(turn != nullptr, crow + (long long)col * g.sCn, (acc[j] * sA_inv) * sB_inv);
        }
      }
      if (turn) {
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NEPI * 32) : "memory");
        if (e == 0 && lane == 0) det_leave(turn, blockIdx.z, gridDim.z);
      }
      T3_SECTION_END(w2);
    }
    T3_ROLE_END(8, warp == Cfg::EPI0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int BN>
static int launch3w(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2, const Tc3Args& g,
                    dim3 grid, cudaStream_t s, const CUtensorMap* tal = nullptr, const CUtensorMap* tal2 = nullptr) {
  using Cfg = T3WCfg<BN>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_index()];
  if (!attr_done) {
    DDRL_CUDA(cudaFuncSetAttribute(tc3_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  tc3_wgrad_kernel<BN><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(ta, tb, ta2, tb2, tal ? *tal : ta, tal2 ? *tal2 : ta2, g);
  prof_work(2.0 * g.M * (double)g.N * g.K);
  if (g_prof_on && g_prof_shapes) {
    char nm[96];
    snprintf(nm, sizeof(nm), "%s3[wgrad,M=%d,N=%d,K=%d,g=%d]", g.tap.mode ? "conv_tc" : "gemm_tc", g.M, g.N, g.K,
             (int)(grid.x * grid.y * grid.z));
    DDRL_LAUNCHED(prof_intern(nm));
    return DDRL_OK;
  }
  DDRL_LAUNCHED("tc3_wgrad_kernel");
  return DDRL_OK;
}

// K splits: one CTA per (tile, split), one CTA per SM at a time -> waves of equal-length CTAs; pick the split count whose
// last wave is fullest, preferring fewer splits among near-equals; every split keeps >= 4 K blocks
static void wgrad3_splits(Tc3Args& g, int tiles) {
  const int max_splits = std::max(1, std::min(g.kb_total / 4, 1024));
  int best = 1;
  double best_cost = 1e300;
  for (int sp = 1; sp <= max_splits && (long long)tiles * sp <= 4LL * kNumSMs; ++sp) {
    const int kbps = ceil_div(g.kb_total, sp);
    const int waves = ceil_div(tiles * ceil_div(g.kb_total, kbps), kNumSMs);
    const double cost = (double)waves * (kbps + 4);
    if (cost < best_cost * 0.98) { best_cost = cost; best = sp; }
  }
  g.kb_per_split = ceil_div(g.kb_total, best);
}

// dW[n*ldw + k] += sum_r dy[r, n] * x[r, k]     (x [rows, Kx], dy [rows, N]; accumulates atomically)
int tc3_wgrad(int Kx, int N, long long rows, const float* x, int ldx, const float* dy, int ldy, const float* amax_x,
              const float* amax_dy, float* dW, int ldw, cudaStream_t s, float* db, float* db2, int db_split) {
  if (Kx < 1 || N < 1 || rows < 1 || !al16(x) || !al16(dy) || ldx % 4 != 0 || ldy % 4 != 0 || rows > 0x7fffffffLL)
    return DDRL_E_UNSUPPORTED;
  if (!amax_x || !amax_dy) return DDRL_E_ARG;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = N > 64 ? 128 : 64;
  CUtensorMap ta, tb;
  r = tc_make_map(&ta, x, Kx, rows, ldx, T3_BK, false);           // [64 rows x 32 k] boxes, plain 128B swizzle
  if (r == DDRL_OK) r = tc_make_map(&tb, dy, N, rows, ldy, T3_BK, false);
  if (r != DDRL_OK) return r;
  Tc3Args g;
  memset(&g, 0, sizeof(g));
  g.C = dW; g.M = Kx; g.N = N; g.K = (int)rows; g.sCm = 1; g.sCn = ldw; g.atomic = 1;
  g.kb_total = (int)ceil_div64(rows, T3_BK);
  g.amax_a = amax_x; g.amax_b = amax_dy; g.colsum = db; g.colsum2 = db2; g.colsum_split = db_split;
  const int tiles = ceil_div(Kx, T3_BM) * ceil_div(N, bn);
  wgrad3_splits(g, tiles);
  dim3 grid(ceil_div(Kx, T3_BM), ceil_div(N, bn), ceil_div(g.kb_total, g.kb_per_split));
  g.det_ctr = getenv("DDRL_DET_SKIP_WGRAD") ? nullptr : det_seq((int)(grid.x * grid.y)).ctr;
  return bn == 128 ? launch3w<128>(ta, tb, ta, tb, g, grid, s) : launch3w<64>(ta, tb, ta, tb, g, grid, s);
}

bool tc3_conv_wgrad_supported(const ConvOp& o) {
  if (o.Cin % 32 != 0 || o.Ctot % 4 != 0 || o.c_off % 4 != 0 || o.c_off + o.Cin > o.Ctot || !al16(o.a)) return false;
  if (o.sx < 1 || o.sx > 8 || o.sy < 1 || o.sy > 8 || o.Xn < 1 || o.Yn < 1 || o.Bn < 1) return false;
  if (o.Xn > T3_BK || (o.Xn - 1) * o.sx + 1 > 256) return false;
  const int ny = std::min(o.Yn, std::max(1, T3_BK / o.Xn));
  return (ny - 1) * o.sy + 1 <= 256;
}

// dy boxes of the tap weight gradient: swizzle 128B (the tile is converted in shared memory, not fed to the MMA)
static int make_map_dy3w(CUtensorMap* m, const float* dy, int ldy, int N, long long ipix, int Bn, int rows) {
  const unsigned long long dims[3] = {(unsigned long long)N, (unsigned long long)ipix, (unsigned long long)Bn};
  const unsigned long long strides[2] = {(unsigned long long)ldy * 4, (unsigned long long)ipix * ldy * 4};
  const unsigned box[3] = {32u, (unsigned)rows, 1u}, estr[3] = {1u, 1u, 1u};
  return tc_encode_tiled(m, false, 3, dy, dims, strides, box, estr, true);
}
static int make_map_dy4w(CUtensorMap* m, const float* dy, int ldy, int N, int Xn, int Yn, int Bn, int xw, int yh, int nb = 1) {
  const unsigned long long dims[4] = {(unsigned long long)N, (unsigned long long)Xn, (unsigned long long)Yn, (unsigned long long)Bn};
  const unsigned long long strides[3] = {(unsigned long long)ldy * 4, (unsigned long long)Xn * ldy * 4, (unsigned long long)Yn * Xn * ldy * 4};
  const unsigned box[4] = {32u, (unsigned)xw, (unsigned)yh, (unsigned)nb}, estr[4] = {1u, 1u, 1u, 1u};
  return tc_encode_tiled(m, false, 4, dy, dims, strides, box, estr, true);
}

int tc3_conv_wgrad(const ConvOp& o, const float* dy, int ldy, int N, const float* amax_x, const float* amax_dy, float* dWp, int ldw,
                   cudaStream_t s, float* db, float* db2, int db_split, const void* a_hi16, const void* a_lo16) {
  if (!tc3_conv_wgrad_supported(o) || N < 1 || ldy % 4 != 0 || !al16(dy)) return DDRL_E_UNSUPPORTED;
  if (!amax_x || !amax_dy) return DDRL_E_ARG;
  const bool ps = a_hi16 != nullptr || a_lo16 != nullptr;
  if (ps && !presplit_ok(o, a_hi16, a_lo16)) return DDRL_E_UNSUPPORTED;
  int r = tc_get_encode();
  if (r != DDRL_OK) return r;
  const int bn = N > 64 ? 128 : 64;
  Tc3Args g;
  memset(&g, 0, sizeof(g));
  tc_tap_common(g.tap, o, T3_BK);
  if (g.tap.nb != 1) { g.tap.nb = 1; g.tap.ny = std::min(o.Yn, std::max(1, T3_BK / o.Xn)); g.tap.tpi = ceil_div(o.Yn, g.tap.ny); }
  g.tap.rows = o.Xn * g.tap.ny;
  g.tap.kpad = (g.tap.rows + 15) & ~15;
  const int K = o.KH * o.KW * o.Cin;
  CUtensorMap ta, tb, ta2, tb2, tal, tal2;
  if (ps) { g.presplit = 1; g.tap.cpb = o.Cin / 64; g.tap.nslices = o.KH * o.KW * g.tap.cpb; }
  auto map_a = [&](CUtensorMap* hi, CUtensorMap* lo, int nx, int ny, int nb = 1) {
    if (!ps) return tc_make_map_nhwc(hi, o, nx, ny, nb, false);
    const int rr = make_map_nhwc16(hi, a_hi16, o, nx, ny, nb);
    return rr != DDRL_OK ? rr : make_map_nhwc16(lo, a_lo16, o, nx, ny, nb);
  };
  // K blocks of exactly 64 pixels from two box classes when the map width is a sum of two powers of two and that packs
  // the image into >= 10 % fewer K blocks than whole rows
  static const bool no_w2 = [] { const char* e = getenv("DDRL_TC2_NO_WGRAD_BOXES"); return e && e[0] == '1'; }();
  if (!no_w2 && ldy == N) {
    for (int xw0 = 64; xw0 >= 2; xw0 >>= 1) {
      const int xw1 = o.Xn - xw0;
      if (xw1 < 1 || xw1 > xw0 || (xw1 & (xw1 - 1)) != 0) continue;
      const int yh0 = 64 / xw0, yh1 = 64 / xw1;
      if ((yh0 - 1) * o.sy + 1 > 256 || (yh1 - 1) * o.sy + 1 > 256 || yh1 > 256) continue;
      const int nb0 = ceil_div(o.Yn, yh0), nb1 = ceil_div(o.Yn, yh1);
      if ((nb0 + nb1) * 10 > g.tap.tpi * 9) continue;
      g.tap.w2on = 1; g.tap.wxw0 = xw0; g.tap.wyh0 = yh0; g.tap.wnb0 = nb0; g.tap.wxw1 = xw1; g.tap.wyh1 = yh1;
      g.tap.tpi = nb0 + nb1; g.tap.rows = 64; g.tap.kpad = 64;
      break;
    }
  }
  if (g.tap.w2on) {
    r = map_a(&ta, &tal, g.tap.wxw0, g.tap.wyh0);
    if (r == DDRL_OK) r = map_a(&ta2, &tal2, g.tap.wxw1, g.tap.wyh1);
    if (r == DDRL_OK) r = make_map_dy4w(&tb, dy, ldy, N, o.Xn, o.Yn, o.Bn, g.tap.wxw0, g.tap.wyh0);
    if (r == DDRL_OK) r = make_map_dy4w(&tb2, dy, ldy, N, o.Xn, o.Yn, o.Bn, g.tap.wxw1, g.tap.wyh1);
  } else {
    // two-phase K tiling (tc_tap_common's plan with 64-row boxes): multi-image boxes / a second box class for the rows a
    // whole number of row blocks leaves over, when that packs the pixels into >= 5 % fewer K blocks
    const char* e_k2 = getenv("DDRL_TC3_NO_WGRAD_K2");
    const bool no_k2 = e_k2 && e_k2[0] == '1';
    if (!no_k2 && ldy == N) {
      TcTap t2;
      tc_tap_common(t2, o, T3_BK, true);
      if (t2.nb > 1 || t2.ny2 > 0) {
        g.tap.ny = t2.ny; g.tap.nb = t2.nb; g.tap.tpi = t2.tpi; g.tap.rows = t2.rows;
        g.tap.y2 = t2.y2; g.tap.ny2 = t2.ny2; g.tap.nb2 = t2.nb2; g.tap.rows2 = t2.rows2; g.tap.tiles1 = t2.tiles1;
        g.tap.kpad = (std::max(t2.rows, t2.ny2 > 0 ? t2.rows2 : 0) + 15) & ~15;
      }
    }
    const bool p2 = g.tap.ny2 > 0;
    r = map_a(&ta, &tal, o.Xn, g.tap.ny, g.tap.nb);
    if (r == DDRL_OK)
      r = (g.tap.nb == 1 && !p2) ? make_map_dy3w(&tb, dy, ldy, N, (long long)o.Yn * o.Xn, o.Bn, g.tap.rows)
                                 : make_map_dy4w(&tb, dy, ldy, N, o.Xn, o.Yn, o.Bn, o.Xn, g.tap.ny, g.tap.nb);
    if (r == DDRL_OK && p2) r = map_a(&ta2, &tal2, o.Xn, g.tap.ny2, g.tap.nb2);
    if (r == DDRL_OK && p2) r = make_map_dy4w(&tb2, dy, ldy, N, o.Xn, o.Yn, o.Bn, o.Xn, g.tap.ny2, g.tap.nb2);
    if (!p2) { ta2 = ta; tb2 = tb; tal2 = tal; }
  }
  if (r != DDRL_OK) return r;
  g.C = dWp; g.M = K; g.N = N; g.K = o.Bn * o.Yn * o.Xn; g.sCm = 1; g.sCn = ldw; g.atomic = 1;
  g.kb_total = g.tap.w2on ? o.Bn * g.tap.tpi : ceil_div(o.Bn, g.tap.nb) * g.tap.tpi + (g.tap.ny2 > 0 ? ceil_div(o.Bn, g.tap.nb2) : 0);
  g.amax_a = amax_x; g.amax_b = amax_dy; g.colsum = db; g.colsum2 = db2; g.colsum_split = db_split;
  const int tiles = ceil_div(K, T3_BM) * ceil_div(N, bn);
  wgrad3_splits(g, tiles);
  dim3 grid(ceil_div(K, T3_BM), ceil_div(N, bn), ceil_div(g.kb_total, g.kb_per_split));
  g.det_ctr = det_seq((int)(grid.x * grid.y)).ctr;
  if (ps) return bn == 128 ? launch3w<128>(ta, tb, ta2, tb2, g, grid, s, &tal, &tal2) : launch3w<64>(ta, tb, ta2, tb2, g, grid, s, &tal, &tal2);
  return bn == 128 ? launch3w<128>(ta, tb, ta2, tb2, g, grid, s) : launch3w<64>(ta, tb, ta2, tb2, g, grid, s);
}

// ---------------------------------------------------------------- operand preparation
// amax|x| over a [rows, cols] view (row stride ld) -> atomicMax on the bits of *slot (the caller zeroes the slot); body in
// prep_kernels.cuh
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long rows, int cols, long long ld,
                                                   unsigned int* __restrict__ slot) {
  amax_body(x, rows, cols, ld, slot, blockIdx.x, gridDim.x);
}
int amax_f32(const float* x, long long rows, int cols, long long ld, float* slot, bool zero_first, cudaStream_t s) {
  if (!slot || rows < 0 || cols < 0) return DDRL_E_ARG;
  if (g_prep_rec) {
    // recorded for the multi-job kernel: the zeroing of the slot belongs to the PREVIOUS phase (the caller keeps a second
    // recorder for it), so only the reduction is recorded here
    if (rows == 0 || cols == 0) return DDRL_OK;
    if (ld > 0x7fffffffLL) return DDRL_E_ARG;
    PrepJob j{}; j.type = PREP_AMAX; j.a = x; j.b = slot; j.total = rows; j.i[0] = cols; j.i[1] = (int)ld;
    j.vblocks = prep_blocks(rows * cols, 1024);
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  if (zero_first) DDRL_CUDA(cudaMemsetAsync(slot, 0, sizeof(float), s));
  if (rows == 0 || cols == 0) return DDRL_OK;
  if (!x) return DDRL_E_ARG;
  const long long work = (rows * cols + 1023) / 1024;
  const int blocks = (int)std::min<long long>(std::max<long long>(work, 1), 8LL * kNumSMs);
  amax_kernel<<<blocks, 256, 0, s>>>(x, rows, cols, ld, reinterpret_cast<unsigned int*>(slot));
  DDRL_LAUNCHED("amax_kernel");
  return DDRL_OK;
}

// weights w [N, ldw] fp32 -> scaled fp16 hi / lo' (+ transposed pair); body in prep_kernels.cuh
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ w, int N, int K, int ldw, const float* __restrict__ amax,
                                                        __half* __restrict__ hi, __half* __restrict__ lo, int ld16,
                                                        __half* __restrict__ hiT, __half* __restrict__ loT, int ldT16) {
  split_f16_body(w, N, K, ldw, amax, hi, lo, ld16, hiT, loT, ldT16, blockIdx.x, gridDim.x);
}
int split_f16(const float* w, int N, int K, int ldw, const float* amax, void* hi, void* lo, int ld16, void* hiT, void* loT, int ldT16,
              cudaStream_t s) {
  if (!w || !amax || !hi || !lo || N < 1 || K < 1 || ld16 < K) return DDRL_E_ARG;
  const long long n = (long long)N * ld16;
  if (g_prep_rec) {
    PrepJob j{}; j.type = PREP_SPLIT_F16; j.a = w; j.b = hi; j.c = lo; j.d = hiT; j.e = loT; j.amax = amax; j.total = n;
    j.i[0] = N; j.i[1] = K; j.i[2] = ldw; j.i[3] = ld16; j.i[4] = ldT16; j.vblocks = prep_blocks(n);
    return prep_record(j) ? DDRL_OK : DDRL_E_STATE;
  }
  const int blocks = (int)std::min<long long>((n + 255) / 256, 8LL * kNumSMs);
  split_f16_kernel<<<blocks, 256, 0, s>>>(w, N, K, ldw, amax, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ld16,
                                          reinterpret_cast<__half*>(hiT), reinterpret_cast<__half*>(loT), ldT16);
  DDRL_LAUNCHED("split_f16_kernel");
  return DDRL_OK;
}

}  // namespace ddrl

// debugging aid (not part of the public header): role counters of the tc3 forward kernel (all zero unless built -DTC3_TIMING)
extern "C" int ddrl_tc3_timing_read(unsigned long long* out64, int reset) {
  if (out64) DDRL_CUDA(cudaMemcpyFromSymbol(out64, ddrl::g_tc3_wait, sizeof(unsigned long long) * 96));
  if (reset) {
    unsigned long long z[64] = {0};
    DDRL_CUDA(cudaMemcpyToSymbol(ddrl::g_tc3_wait, z, sizeof(z)));
  }
  return DDRL_OK;
}
