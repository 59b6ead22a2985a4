// Tensor-core (tcgen05 / TMEM) engine of the learned-coefficient row integrator, sm_100a.
//
// The conv stack is 96 % of the FLOPs and is an implicit GEMM per 128-position tile:
//   hidden layer : D[128 x 32] += sum_{tap k, ci-block} A_k[128 x 16] * B_k[16 x 32]   (K = 5*32)
//   last layer   : D[128 x NL] += ...                                                   (NL = 16 | 32)
// A = activations kept in shared memory as K-major fp16 "chunk planes" [ci/8][position][8 halfs] (no
// swizzle), so the tap shift k is just +16 B on the descriptor start address and the periodic halo is two
// extra positions per plane.  B = filters pre-packed on the host in the same canonical layout.  The
// polynomial-accuracy projection is folded into the last layer's filters on the host (W3' = W3 . nullspace,
// window form), so the last epilogue reads stencil coefficients straight out of TMEM.
//
// Operand precision (template parameter PREC; ddd1d.h DDD1D_ENGINE_TENSOR*):
//   PREC = 3  x * s = hi + lo (both fp16, s a power of two from a verified bound on the row), filters
//             W * sw = Wh + Wl, products hi*Wh + hi*Wl + lo*Wh accumulated in FP32 in TMEM (the dropped lo*Wl is
//             2^-22 relative): FP32-faithful.  B holds [Wh | Wl] side by side, so one MMA of width 2*NB gives
//             hi*Wh ("main" columns) and hi*Wl ("cross" columns) and a second of width NB adds lo*Wh to the
//             cross columns.  The tensor core truncates when it adds into an accumulator (measured,
//             profiles/r01/tc_precision.txt), so the small cross terms keep their own columns and the
//             epilogue adds main + cross in FP32.
//   PREC = 2  hi * [Wh | Wl]: activations rounded to fp16 (11 bits), filters at 22 bits; no lo planes.
//   PREC = 1  hi * Wh: plain fp16 operands, FP32 accumulate.
//
// Geometry (compile time).  A CTA (one per SM, persistent) has R row teams of 128 * TILES threads -- thread
// <-> position of a tile, so a warp reads exactly its own TMEM lane quadrant -- and every team keeps two
// "slots" in flight.  A slot is one row of N = 128 * TILES points, or, for the reference's small grids,
// RPT = 2 | 4 rows of N = 64 | 32 points packed into one 128-position tile (then every group of 8 positions
// is stored with its own two-position halos and the descriptor's stride between 8-row groups (SBO) steps
// over them, which keeps the tap shift a plain start-address offset although rows wrap inside the tile).
// The two slots of a team take turns: while one slot's MMAs run on the tensor pipe, the team's threads do
// the other slot's CUDA-core phase.  MMAs are issued by dedicated warps (tcgen05.mma blocks its issuer for
// about as long as the pipe is busy): an issuer serves the slots of one team in the fixed order in which
// the team requests them, so it sleeps in mbarrier.try_wait instead of polling.
//
// Per right-hand side a slot goes through
//   phase 0  stage value (float64 state + dt * sum a k, rounded to float32), / sigma, exchange with the
//            neighbours through an L1-resident global scratch, first conv layer 1 -> 32 on the CUDA cores
//            (filters are constant-bank FFMA operands), ReLU, fp16 planes, request the hidden layer's MMAs;
//   phase 1  (per hidden layer) wait for the tile's tcgen05.commit, TMEM -> registers, bias, ReLU, planes
//            rewritten in place, request the next layer;
//   phase 2  wait, stencil coefficients from TMEM, 7-point window dot products on the un-normalised row,
//            equation of motion, flux difference, forcing, stage derivative; after the last stage the
//            Runge-Kutta update and the snapshot.
// Shared memory holds only what the tensor pipe reads (shared-memory LOADS stall while MMAs stream their
// operands: profiles/r01/tc_overlap.txt); per-thread state lives in registers.
//
// What binds the kernel (profiles/r02/README.md): the tensor core's shared-memory operand fetch.  An MMA of M = 128,
// K = 16 costs 32 + N/4 clocks of wavefronts for N/2 clocks of math, so the N = 64 + 32 MMAs of a 32-channel conv keep
// the unit busy 85-93 % of the cycles at 38 % math.  The CUDA-core side has four team warps per scheduler and is
// bound by dependent-latency chains, not by its instruction count: the step loop therefore has no jump table (the
// plain equations and `integrating` are predicates of their own), no division, no stack array (per-call exports take
// values), and the stage derivatives rotate instead of being dispatched on the stage index.
#pragma once
#include <type_traits>

#include "ddd1d_tc_common.cuh"

namespace ddd1d {
namespace tc {

constexpr int kFsStride = 2 * kMaxModes;     // forcing mode amplitudes per RK stage: [sine 0..7 | cosine 0..7]
constexpr int kFsBuffers = 4;                // amplitude sets in rotation (three are needed: written one step ahead of their use; four makes the index a mask)
constexpr int kFsWords = kFsBuffers * kMaxStages * kFsStride;
constexpr int kMaxRpt = 4;

// Everything the kernel reads that is not a tensor-pipe operand: kernel parameter = constant bank.
struct TcParams {
  int eq, D, S, wshift, nhid;       // nhid: tensor layers between the first and the last conv (0 | 1)
  int plain_burgers, plain_kdv, plain_ks;   // eq as three flags: the common forms are tested one by one, a switch over
                                            // eq compiles to a jump table (two dependent constant loads) per right-hand side
  int M, P, fcap;                   // forcing: modes, terms per sample, samples set
  int debug;                        // TIMING EXPERIMENTS ONLY (DDD1D_TC_DEBUG): bit 0 no MMAs, bit 2 teams do not wait
                                    // for their MMAs (garbage results), bit 6 MMA stream alone
  float sigma, eta, inv_dx;
  float w1abs, b1abs, whabs, bhabs;         // operator-norm bounds behind the fp16 plane scales
  float inv_sw_hid, inv_sw_last;            // 1 / power-of-two filter scales
  const float* blob;                // B planes: hidden [20][64 rows][16 B] then last [20][2 NL rows][16 B]
  const float* fparams;             // [fcap][4][P]: a, omega, phi, signed k
  const float* fbasis;              // [2M][N]
  float* scratch;                   // [grid][slots of the CTA][sc_stride] floats (L1 / L2 resident)
  alignas(16) float w1[kTaps * kF]; // first layer [tap][channel]
  alignas(16) float b1[kF];         // biases: first layer,
  alignas(16) float bh[kF];         //         hidden tensor layer,
  alignas(16) float bl[kF];         //         folded last layer
};

#ifndef DDD1D_TC_SPLIT_REQ
#define DDD1D_TC_SPLIT_REQ 0       // 1: request a layer per ci-block (measured: correct, 2.5 % slower at C2)
#endif

#ifndef DDD1D_TC_TILE_REQ
#define DDD1D_TC_TILE_REQ 0        // 1: rows of several tiles request their layers per tile (measured: bit-identical,
                                   // C2 -2.3 %, C4 -5 %: finer requests cost the issuers more waits than they start MMAs early)
#endif

#ifndef DDD1D_TC_WS
#define DDD1D_TC_WS 0              // 1: the wide MMAs of rows of several tiles as tcgen05.mma.ws, B kept in a collector buffer
                                   // (measured: bit-identical results, C2 -0.9 %, C4 -7 %: tiles finish together instead of
                                   // one after the other, which costs more than the saved B fetches)
#endif

#ifndef DDD1D_TC_ISSUER_FORCING
#define DDD1D_TC_ISSUER_FORCING 0  // 1: the issuer warps compute the next step's forcing amplitudes (measured: correct,
                                   // C2 6.1e9 -> 4.8e9: anything an issuer does between two requests delays the MMA stream)
#endif

// SL_: rows ("slots") a team keeps in flight.  2: every slot owns a TMEM accumulator block.  3: the team's slots
// share its TWO blocks as a pool (a block is only needed from a layer's request to the tcgen05.ld of its epilogue),
// so a slot has two turns of the other slots between a request and the need for its result instead of one.
template <int TILES_, int RPT_, int NL_, int PREC_, int SL_ = 2>
struct Geo {
  static constexpr int TILES = TILES_, RPT = RPT_, NL = NL_, PREC = PREC_, SL = SL_;
  static_assert(SL == 2 || SL == 3, "two or three slots per team");
  static_assert(TILES == 1 || RPT == 1, "packed rows live in one tile");
  static constexpr int TEAM = 128 * TILES;            // threads of a team = positions of a slot
  static constexpr int N = TEAM / RPT;                // points of a row
  // row teams per CTA (packed rows with NL = 32 at full precision: three, to fit shared memory)
  static constexpr int R = (RPT > 1 && NL == 32 && PREC == 3) ? 3 : 4 / TILES;
  static constexpr int TS = SL * R;                   // slots per CTA
  static constexpr int BLOCKS = 2 * R;                // TMEM accumulator blocks (of TILES tiles) per CTA
  static_assert(SL == 2 || R == 2, "the block pool is written for one issuer per team");
  static constexpr int ISSUERS = R > 2 ? R : 2;
  static constexpr int SPI = 2 * R / ISSUERS;         // slots per issuer with SL = 2 (consecutive: one team's, or one)
  static constexpr int TEAM_WARPS = TEAM / 32;
  static constexpr int THREADS = R * TEAM + 32 * ISSUERS;
  static constexpr int GROUP = RPT == 1 ? 8 : 12;     // plane positions stored per 8 positions of a tile
  static constexpr uint32_t SBO = GROUP * 16;         // bytes between the 8-row groups of an MMA operand
  static constexpr uint32_t PLANE = RPT == 1 ? (TEAM + 4) * 16 : 16 * 12 * 16;           // bytes of one chunk plane
  static constexpr int PLANES = 4 * (PREC == 3 ? 2 : 1);                              // hi (+ lo), 4 chunk planes each
  static constexpr uint32_t SLOT_BYTES = ((PLANES * PLANE + 127) / 128) * 128;
  static constexpr uint32_t BH_BYTES = kTaps * 4 * 64 * 16;                           // hidden B planes [20][Wh 32 | Wl 32][16 B]
  static constexpr uint32_t BL_BYTES = kTaps * 4 * 2 * NL * 16;
  // Split requests: a layer's MMAs over the first ci-block (chunk planes 0, 1) are requested as soon as those
  // planes are stored, half a phase before the rest, so only half a layer is still to run when the phase ends.
  static constexpr bool SPLIT = DDD1D_TC_SPLIT_REQ != 0;
  // TREQ: a row of several tiles requests a layer PER TILE: tile m's MMAs start when its own four warps and the two
  // neighbouring edge warps (whose positions are its halo) have stored their planes, not when the whole row has
  static constexpr bool TREQ = DDD1D_TC_TILE_REQ != 0 && TILES >= 2 && SL == 2 && !SPLIT;
  static constexpr int REQS = SPLIT ? 2 : TREQ ? TILES : 1;   // request barriers per slot (per ci-block | per tile)
  // mbarriers: [0] blob copy | [BAR_REQ + kb * TS + ts] "planes of ci-block kb stored" (one arrival per team warp)
  // | [BAR_DONE + blk * TILES + m] "tile's MMAs done" | pool only: [BAR_READ + blk] "block read"
  static constexpr int BAR_REQ = 1, BAR_DONE = BAR_REQ + REQS * TS, BAR_READ = BAR_DONE + BLOCKS * TILES;
  // AMP (opt-in experiment, off): the forcing amplitudes of the NEXT step computed by the issuer warps instead of
  // three warps of the team; [BAR_AMP + ts] "next step's amplitudes written" (one arrival per step, by the issuer)
  static constexpr bool AMP = DDD1D_TC_ISSUER_FORCING != 0 && RPT == 1 && SL == 2;
  static constexpr int BAR_AMP = BAR_READ + (SL == 3 ? BLOCKS : 0);
  static constexpr int NBARS = BAR_AMP + (AMP ? TS : 0);
  static constexpr uint32_t OFF_BAR = 0, OFF_TMEM = 240, OFF_BLOB = 256;
  static_assert(NBARS * 8 <= (int)OFF_TMEM, "mbarriers overlap the TMEM address slot");
  static constexpr uint32_t OFF_SLOTS = OFF_BLOB + BH_BYTES + BL_BYTES;
  static constexpr uint32_t SMEM = OFF_SLOTS + TS * SLOT_BYTES;
  static constexpr int COLS = PREC == 1 ? 32 : 64;    // TMEM columns of one tile's accumulator block
  static constexpr int TMEM_USED = BLOCKS * TILES * COLS;  // 512 (PREC >= 2) or 256 (384 with three teams)
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;     // allocations are powers of two
  // global scratch of one slot, in floats
  static constexpr int ROWBUF = N + 2 * kHalo + 2;    // one row with its halo
  static constexpr int SC_ROWS = 0;                   // [stage parity][raw | normalised][RPT][ROWBUF]
  static constexpr int SC_FLUX = 4 * RPT * ROWBUF;
  static constexpr int SC_UMAX = SC_FLUX + TEAM;
  static constexpr int SC_BAD = SC_UMAX + 4;          // [RPT] first non-finite step
  static constexpr int SC_FS = SC_BAD + 4;            // [RPT][kFsWords]
  static constexpr int SC_STRIDE = ((SC_FS + RPT * kFsWords + 31) / 32) * 32;
};

// Mode amplitudes of one sample's forcing at time t (equations.py:196-219), by one warp: lane q holds term q
// (terms beyond 32 in further rounds), the amplitude of mode m is the warp sum of the terms with |k| == m.
// Writes fs[0..8) = sum a sin(w t + phi) (the factors of the cosine basis) and fs[8..16) = sum sgn(k) a
// cos(w t + phi) (the factors of the sine basis); modes beyond M are zero.
static __device__ __noinline__ void forcing_amplitudes_t(const TcParams& P, float* fs, int sample, float t, int lane) {
  float ps[kMaxModes], pc[kMaxModes];
#pragma unroll
  for (int m = 0; m < kMaxModes; ++m) ps[m] = pc[m] = 0.f;
  const float* fp = P.fparams + (size_t)sample * 4 * P.P;
  for (int q = lane; q < P.P; q += 32) {
    const float a = __ldg(fp + q), w = __ldg(fp + P.P + q), phi = __ldg(fp + 2 * P.P + q), k = __ldg(fp + 3 * P.P + q);
    float sn, cs;
    sincosf(fmaf(w, t, phi), &sn, &cs);
    const float a_sin = a * sn, a_cos = (k < 0.f ? -a : a) * cs, ka = fabsf(k);
#pragma unroll
    for (int m = 0; m < kMaxModes; ++m)
      if (ka == (float)(m + 1)) { ps[m] += a_sin; pc[m] += a_cos; }
  }
  float mine_s = 0.f, mine_c = 0.f;
#pragma unroll
  for (int m = 0; m < kMaxModes; ++m) {
    if (m >= P.M) break;
    float a = ps[m], b = pc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == m) { mine_s = a; mine_c = b; }
  }
  if (lane < kMaxModes) {
    fs[lane] = mine_s;
    fs[kMaxModes + lane] = mine_c;
  }
}

// ---- issue ---------------------------------------------------------------------------------------
// Every MMA of one layer for one 128-position tile, by one elected lane of a converged warp.  All strides are
// compile-time, so the descriptors are immediates added to two uniform registers.
//   a_hi / a_lo : shared addresses (>> 4) of the tile's first plane, position 0
//   b           : shared address (>> 4) of the layer's [Wh | Wl] planes (2 * NB rows of 16 B per plane)
//   d           : TMEM address of the tile's block, layout [main NB | cross NB]
__device__ __forceinline__ void mma_f16_ab(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <class G, int NB>
__device__ __forceinline__ void issue_tile(uint32_t a_hi, uint32_t a_lo, uint32_t b, uint32_t d) {
  constexpr uint32_t plane16 = G::PLANE >> 4, bplane16 = (2u * NB * 16u) >> 4;
  const uint32_t desc_hi = (G::SBO >> 4) | (1u << 14);          // A: stride between 8-row groups; descriptor version 1
  const uint32_t bdesc_hi = (128u >> 4) | (1u << 14);           // B: filter rows are contiguous
  const uint32_t ah0 = (a_hi & 0x3FFFu) | (plane16 << 16);       // leading byte offset = plane pitch (the two K chunks)
  const uint32_t al0 = (a_lo & 0x3FFFu) | (plane16 << 16);
  const uint32_t b0 = (b & 0x3FFFu) | (bplane16 << 16);
  const uint32_t idesc_wide = instr_desc_f16(128, 2 * NB), idesc_narrow = instr_desc_f16(128, NB);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
        const uint32_t bo = (uint32_t)(k * 4 + 2 * kb) * bplane16;
        const uint32_t acc = (k == 0 && kb == 0) ? 0u : 1u;
        if (G::PREC == 1) {
          mma_f16_ab(d, ah0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_narrow, acc);      // hi * Wh
        } else {
          mma_f16_ab(d, ah0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_wide, acc);        // hi * [Wh | Wl] -> main | cross
          if (G::PREC == 3) mma_f16_ab(d + NB, al0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_narrow, 1u);   // lo * Wh -> cross
        }
      }
    }
  }
  __syncwarp();
}

// The same for one ci-block kb (run-time: one copy of the code serves both halves of a split request): the five
// taps of K chunk pair kb; the very first MMA of the layer (kb = 0, tap 0) overwrites the accumulators.
template <class G, int NB>
__device__ __forceinline__ void issue_tile_half(uint32_t a_hi, uint32_t a_lo, uint32_t b, uint32_t d, uint32_t kb) {
  constexpr uint32_t plane16 = G::PLANE >> 4, bplane16 = (2u * NB * 16u) >> 4;
  const uint32_t desc_hi = (G::SBO >> 4) | (1u << 14);
  const uint32_t bdesc_hi = (128u >> 4) | (1u << 14);
  const uint32_t ah0 = ((a_hi + 2u * kb * plane16) & 0x3FFFu) | (plane16 << 16);
  const uint32_t al0 = ((a_lo + 2u * kb * plane16) & 0x3FFFu) | (plane16 << 16);
  const uint32_t b0 = ((b + 2u * kb * bplane16) & 0x3FFFu) | (bplane16 << 16);
  const uint32_t idesc_wide = instr_desc_f16(128, 2 * NB), idesc_narrow = instr_desc_f16(128, NB);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
      const uint32_t ao = (uint32_t)k, bo = (uint32_t)(k * 4) * bplane16;
      const uint32_t acc = k == 0 ? kb : 1u;
      if (G::PREC == 1) {
        mma_f16_ab(d, ah0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_narrow, acc);
      } else {
        mma_f16_ab(d, ah0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_wide, acc);
        if (G::PREC == 3) mma_f16_ab(d + NB, al0 + ao, desc_hi, b0 + bo, bdesc_hi, idesc_narrow, 1u);
      }
    }
  }
  __syncwarp();
}

// Weight-stationary form of the wide MMA (tcgen05.mma.ws, N = 64): B stays in collector buffer b<BUF> from the tile
// that fills it to the tile that uses it last, so only the first tile of a row fetches the filter planes (measured:
// 46.4 instead of 53.0 clk per MMA of a pair, scripts/tc_reuse_rate.py).  One buffer per issuer warp.
//   OP: 0 fill, 1 use, 2 lastuse
template <int BUF, int OP>
__device__ __forceinline__ void mma_f16_ws(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
#define DDD1D_WS_ASM(BN, ON)                                                                                        \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"      \
               "setp.ne.b32 p, %6, 0;\n\t"                                                                          \
               "tcgen05.mma.ws.cta_group::1.kind::f16.collector::" BN "::" ON " [%0], da, db, %5, p;\n\t}" ::"r"(    \
                   tmem_d),                                                                                         \
               "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)                              \
               : "memory")
  if (BUF == 0) {
    if (OP == 0) DDD1D_WS_ASM("b0", "fill"); else if (OP == 1) DDD1D_WS_ASM("b0", "use"); else DDD1D_WS_ASM("b0", "lastuse");
  } else {
    if (OP == 0) DDD1D_WS_ASM("b1", "fill"); else if (OP == 1) DDD1D_WS_ASM("b1", "use"); else DDD1D_WS_ASM("b1", "lastuse");
  }
#undef DDD1D_WS_ASM
}

// Every MMA of one layer for ALL tiles of a slot, tiles innermost: per (tap, ci-block) the wide MMAs of the tiles share
// their B planes through the collector, then the narrow ones follow.  slot16: shared address (>> 4) of the slot's
// first plane; d0: TMEM address of the slot's first tile block.
template <class G, int NB, int BUF>
__device__ __forceinline__ void issue_slot_ws(uint32_t slot16, uint32_t b, uint32_t d0) {
  static_assert(G::TILES >= 2 && NB == 32 && G::PREC >= 2, "wide MMAs of width 64 over several tiles");
  constexpr uint32_t plane16 = G::PLANE >> 4, bplane16 = (2u * NB * 16u) >> 4;
  const uint32_t desc_hi = (G::SBO >> 4) | (1u << 14);
  const uint32_t bdesc_hi = (128u >> 4) | (1u << 14);
  const uint32_t ah0 = (slot16 & 0x3FFFu) | (plane16 << 16);
  const uint32_t al0 = ((slot16 + ((4u * G::PLANE) >> 4)) & 0x3FFFu) | (plane16 << 16);
  const uint32_t b0 = (b & 0x3FFFu) | (bplane16 << 16);
  const uint32_t idesc_wide = instr_desc_f16(128, 2 * NB), idesc_narrow = instr_desc_f16(128, NB);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
        const uint32_t bo = (uint32_t)(k * 4 + 2 * kb) * bplane16;
        const uint32_t acc = (k == 0 && kb == 0) ? 0u : 1u;
#pragma unroll
        for (int m = 0; m < G::TILES; ++m) {
          const uint32_t d = d0 + (uint32_t)(m * G::COLS), am = ah0 + ao + (uint32_t)m * 128u;
          if (m == 0) mma_f16_ws<BUF, 0>(d, am, desc_hi, b0 + bo, bdesc_hi, idesc_wide, acc);
          else if (m == G::TILES - 1) mma_f16_ws<BUF, 2>(d, am, desc_hi, b0 + bo, bdesc_hi, idesc_wide, acc);
          else mma_f16_ws<BUF, 1>(d, am, desc_hi, b0 + bo, bdesc_hi, idesc_wide, acc);
        }
        if (G::PREC == 3) {
#pragma unroll
          for (int m = 0; m < G::TILES; ++m)
            mma_f16_ab(d0 + (uint32_t)(m * G::COLS) + NB, al0 + ao + (uint32_t)m * 128u, desc_hi, b0 + bo, bdesc_hi,
                       idesc_narrow, 1u);
        }
      }
    }
  }
  __syncwarp();
}

// ---- plane stores ----------------------------------------------------------------------------------
// byte offset of tile position p inside a chunk plane
template <class G>
__device__ __forceinline__ uint32_t plane_pos(int p) {
  return G::RPT == 1 ? (uint32_t)(p + 2) * 16u : (uint32_t)((p >> 3) * 12 + 2 + (p & 7)) * 16u;
}

// Where a thread's activations also go (the halo copies), as byte offsets relative to its own position;
// 0 = no copy.  RPT == 1: the first / last two positions of the row wrap around.  RPT > 1: every group of 8
// carries its own halos, so positions 0, 1 of a group are also the right halo of the previous group of the
// same row and positions 6, 7 the left halo of the next one.
template <class G>
__device__ __forceinline__ int halo_copy_offset(int p) {
  if (G::RPT == 1) {
    if (p < 2) return G::TEAM * 16;
    if (p >= G::TEAM - 2) return -G::TEAM * 16;
    return 0;
  }
  constexpr int groups = G::N / 8;                 // groups per row
  const int g = p >> 3, j = p & 7, gr = g % groups;
  if (j < 2) {                                     // -> slot 10 + j of the previous group of the row
    const int pg = gr == 0 ? g + groups - 1 : g - 1;
    return ((pg * 12 + 10 + j) - (g * 12 + 2 + j)) * 16;
  }
  if (j >= 6) {                                    // -> slot j - 6 of the next group of the row
    const int ng = gr == groups - 1 ? g - (groups - 1) : g + 1;
    return ((ng * 12 + j - 6) - (g * 12 + 2 + j)) * 16;
  }
  return 0;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- packed FP32x2 arithmetic (FFMA2 / FMUL2 / FADD2: two lanes per issue slot) -------------------------
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long b) { return *reinterpret_cast<float2*>(&b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
// two consecutive entries of a constant-bank table (i even)
__device__ __forceinline__ float2 pair_at(const float* table, int i) { return *reinterpret_cast<const float2*>(table + i); }

// mbarrier wait for the hot path: plain try_wait loop; a protocol bug traps after ~2^27 timeouts instead of hanging
// (barriers are addressed by their 32-bit shared address, computed once per thread: converting the generic pointer at
// every wait costs an S2UR and four more instructions in front of each try_wait)
__device__ __forceinline__ void mbar_arrive_s(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(uint32_t addr, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(100000u)
        : "memory");
    if (done) return;
    if (++spins > (1u << 27)) asm volatile("trap;");
  }
}

// Slow path of the activation bound: the slot's maximum of |u / sigma| by an atomic performed in L2, read past
// L1; returns twice the maximum.  Runs for a slot's first stage and when a row outgrows (or decays 256x below) its
// bound.
static __device__ __noinline__ float recalibrate(unsigned int* umax_w, float usn, int team, int team_threads, int lane,
                                                 bool leader) {
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(usn)));
  if (lane == 0) atomicMax(umax_w, wmax);
  team_sync(team, team_threads);
  const float m = __uint_as_float(__ldcg(umax_w));
  team_sync(team, team_threads);
  if (leader) *umax_w = 0u;                       // next use is at least one barrier away
  return 2.f * m;
}

// sixteen accumulator columns of one position as eight pairs: main (+ cross)
template <int PREC>
__device__ __forceinline__ void tmem_read_pairs(uint32_t t_main, uint32_t t_cross, float2 (&v)[8]) {
  uint32_t a[16];
  tmem_ld16_issue(t_main, a);
  if (PREC == 1) {
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1]));
  } else {
    uint32_t b[16];
    tmem_ld16_issue(t_cross, b);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = fadd2(make_float2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1])),
                   make_float2(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1])));
  }
}

// relu(x) = hi + lo as two fp16 pairs, ReLU included.
//   DDD1D_SPLIT_RZ (4 instructions per pair): hi = fp16(x) rounded TOWARDS ZERO and clamped at zero
//     (F2FP.RELU.RZ), the exact remainders x - hi by the mixed-precision FMA (FHFMA: f16 * f16 + f32), lo =
//     fp16(remainder) clamped at zero (F2FP.RELU): with hi rounded towards zero the remainder of a positive x is
//     never negative, so the second clamp only acts where x < 0.  The remainder spans a whole ulp of hi: 2^-23.
//   default (6 per pair): hi = fp16(x) rounded to nearest, clamped (F2FP.RELU); remainders max(x, 0) - hi (two
//     FMNMX, two FHFMA), lo = fp16(remainder): 2^-24, the float32 operand precision.
__device__ __forceinline__ void split_pair_relu(float2 v, uint32_t& hi, uint32_t& lo) {
#ifdef DDD1D_SPLIT_RZ
  asm("{\n\t.reg .b16 h0, h1, m1;\n\t.reg .f32 d0, d1;\n\t"
      "cvt.rz.relu.f16x2.f32 %0, %3, %2;\n\t"
      "mov.b32 {h0, h1}, %0;\n\t"
      "mov.b16 m1, 0xBC00;\n\t"                       // -1.0
      "fma.rn.f32.f16 d0, h0, m1, %2;\n\t"
      "fma.rn.f32.f16 d1, h1, m1, %3;\n\t"
      "cvt.rn.relu.f16x2.f32 %1, d1, d0;\n\t}"
      : "=&r"(hi), "=&r"(lo)
      : "f"(v.x), "f"(v.y));
#else
  const float rx = fmaxf(v.x, 0.f), ry = fmaxf(v.y, 0.f);
  asm("{\n\t.reg .b16 h0, h1, m1;\n\t.reg .f32 d0, d1;\n\t"
      "cvt.rn.f16x2.f32 %0, %3, %2;\n\t"
      "mov.b32 {h0, h1}, %0;\n\t"
      "mov.b16 m1, 0xBC00;\n\t"                       // -1.0
      "fma.rn.f32.f16 d0, h0, m1, %2;\n\t"
      "fma.rn.f32.f16 d1, h1, m1, %3;\n\t"
      "cvt.rn.f16x2.f32 %1, d1, d0;\n\t}"
      : "=&r"(hi), "=&r"(lo)
      : "f"(rx), "f"(ry));
#endif
}
// relu(x) rounded to one fp16 pair (PREC < 3)
__device__ __forceinline__ uint32_t pack_half2_relu(float2 v) {
  uint32_t h;
  asm("cvt.rn.relu.f16x2.f32 %0, %2, %1;" : "=r"(h) : "f"(v.x), "f"(v.y));
  return h;
}

// relu of eight consecutive channels of one position (four pairs) -> one 16-byte chunk of the hi plane (and of the
// lo plane)
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// `mine`: 32-bit shared address of the thread's position in the slot's first plane
template <class G>
__device__ __forceinline__ void store_chunk8(uint32_t mine, int c8, int copy_off, bool has_copy,
                                             const float2 (&v)[4]) {
  uint4 h, l;
  if (G::PREC == 3) {
    split_pair_relu(v[0], h.x, l.x);
    split_pair_relu(v[1], h.y, l.y);
    split_pair_relu(v[2], h.z, l.z);
    split_pair_relu(v[3], h.w, l.w);
  } else {
    h.x = pack_half2_relu(v[0]);
    h.y = pack_half2_relu(v[1]);
    h.z = pack_half2_relu(v[2]);
    h.w = pack_half2_relu(v[3]);
  }
  const uint32_t hp = mine + (uint32_t)c8 * G::PLANE;
  sts128(hp, h);
  if (G::PREC == 3) sts128(hp + 4 * G::PLANE, l);
  if (has_copy) {
    sts128(hp + copy_off, h);
    if (G::PREC == 3) sts128(hp + 4 * G::PLANE + copy_off, l);
  }
}

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// NB accumulator columns of one position: main (+ cross)
template <int NB, int PREC>
__device__ __forceinline__ void tmem_read(uint32_t taddr, float (&v)[NB]) {
  if (NB == 32) {
    uint32_t a[32];
    tmem_ld32_issue(taddr, a);
    if (PREC == 1) {
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]);
    } else {
      uint32_t b[32];
      tmem_ld32_issue(taddr + 32, b);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(b[i]);
    }
  } else {
    uint32_t a[16];
    tmem_ld16_issue(taddr, a);
    if (PREC == 1) {
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]);
    } else {
      uint32_t b[16];
      tmem_ld16_issue(taddr + 16, b);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(b[i]);
    }
  }
}

// per-thread state of one slot that outlives a phase; slots are unrolled, so this lives in registers
// a + b = s + e exactly (Knuth's TwoSum)
__device__ __forceinline__ void two_sum(float a, float b, float& s, float& e) {
  s = a + b;
  const float bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}

struct SlotState {
  // The solution, float64-equivalent (integrate.py:154: SciPy carries y in float64), as an unevaluated sum of two
  // floats yh + yl with yh = float(yh + yl): ~48 significant bits and no float64 instruction in the loop
  // (conversions to / from float64 run at 16 lanes per clock per SM).
  float yh, yl;
  // Stage derivatives of the step in progress, newest first: after stage s, k0 = k_s, k1 = k_{s-1}, ...  (they rotate
  // on every stage, so nothing dispatches on the stage index; the tableau rows are fetched in the same order)
  float k0, k1, k2, k3;
  float umax;        // verified bound on the slot's max |u / sigma|; < 0: not calibrated yet
};

// The two slots of a team share ONE copy of the phase code (the steady-state loop has to stay inside the
// instruction cache: two unrolled copies were 80 KB), selected by the run-time slot index: a field of the slot
// in turn is one SEL to read and two predicated moves to write (exchanging the two register sets after every
// turn cost ~40 moves).
template <int SL>
struct SlotSet;
template <>
struct SlotSet<2> {
  SlotState a, b;
};
template <>
struct SlotSet<3> {
  SlotState a, b, c;
};
__device__ __forceinline__ float slot_get(const SlotSet<2>& s, int sl, float SlotState::*f) { return sl ? s.b.*f : s.a.*f; }
__device__ __forceinline__ float slot_get(const SlotSet<3>& s, int sl, float SlotState::*f) {
  return sl == 0 ? s.a.*f : sl == 1 ? s.b.*f : s.c.*f;
}
__device__ __forceinline__ void slot_put(SlotSet<2>& s, int sl, float SlotState::*f, float v) {
  if (sl) s.b.*f = v;
  else s.a.*f = v;
}
__device__ __forceinline__ void slot_put(SlotSet<3>& s, int sl, float SlotState::*f, float v) {
  if (sl == 0) s.a.*f = v;
  else if (sl == 1) s.b.*f = v;
  else s.c.*f = v;
}
#define DDD1D_SLOT_GET(SS, sl, f) slot_get(SS, sl, &SlotState::f)
#define DDD1D_SLOT_PUT(SS, sl, f, v) slot_put(SS, sl, &SlotState::f, v)

// OP_COEF export of the last epilogue (per-call parity hook, not on the integration path): window column q of one
// grid point.  One value per call, so that the caller's sixteen columns stay in registers (an array argument put
// them on the stack of every right-hand side, export or not).
static __device__ __noinline__ void export_coefficient(const TcParams& P, const Work& W, float c, int q, size_t point) {
  const int d = q / kWin, slot = q % kWin - P.wshift;
  if (d < P.D && slot >= 0 && slot < P.S) W.out[(point * P.D + d) * P.S + slot] = c;
}

// OP_DERIV export (per-call parity hook): out of line, the integration loop has to stay small; by value, so that
// the derivatives stay in registers
static __device__ __noinline__ void export_derivatives(const TcParams& P, const Work& W, float d0, float d1, float d2,
                                                       float d3, size_t point) {
  static_assert(kMaxD == 4, "one argument per derivative");
  W.out[point * P.D] = d0;
  if (P.D > 1) W.out[point * P.D + 1] = d1;
  if (P.D > 2) W.out[point * P.D + 2] = d2;
  if (P.D > 3) W.out[point * P.D + 3] = d3;
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <int TILES, int RPT, int NL, int PREC, int SL = 2>
__global__ void __launch_bounds__(Geo<TILES, RPT, NL, PREC, SL>::THREADS, 1)
tc_row_kernel(const __grid_constant__ TcParams P, const __grid_constant__ Work W, const __grid_constant__ Tableau tab) {
  using G = Geo<TILES, RPT, NL, PREC, SL>;
  constexpr int N = G::N, TEAM = G::TEAM, R = G::R, TS = G::TS, BLOCKS = G::BLOCKS;
  constexpr bool POOL = SL == 3;
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler (role branches, uniform registers)
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem_raw + G::OFF_BAR);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + G::OFF_TMEM);
  // barrier map: Geo (blk = ts with two slots per team; "block read": one arrival per team warp once the epilogue
  // has its accumulators in registers, the block may be overwritten)
  constexpr bool SPLIT = G::SPLIT;
  constexpr int BAR_REQ = G::BAR_REQ, BAR_DONE = G::BAR_DONE, BAR_READ = G::BAR_READ;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    for (int t = 0; t < G::REQS * TS; ++t) mbar_init(&bars[BAR_REQ + t], G::TREQ ? 6u : (uint32_t)G::TEAM_WARPS);
    for (int t = 0; t < BLOCKS * TILES; ++t) mbar_init(&bars[BAR_DONE + t], 1);
    if (POOL)
      for (int t = 0; t < BLOCKS; ++t) mbar_init(&bars[BAR_READ + t], (uint32_t)G::TEAM_WARPS);
    if (G::AMP)
      for (int t = 0; t < TS; ++t) mbar_init(&bars[G::BAR_AMP + t], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    constexpr uint32_t bytes = G::BH_BYTES + G::BL_BYTES;
    mbar_expect_tx(&bars[0], bytes);
    bulk_copy_g2s(smem_raw + G::OFF_BLOB, P.blob, bytes, &bars[0]);
  }
  if (warp == 0) tmem_alloc(tmem_slot, G::TMEM_COLS);
  mbar_wait_guarded(&bars[0], 0);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_teams = gridDim.x * R;
  const int units = (W.batch + RPT - 1) / RPT;               // slots' worth of rows in the batch
  const int nsteps = (W.op == OP_INTEGRATE) ? W.nsteps : 1;
  const int nstages = (W.op == OP_INTEGRATE) ? tab.stages : 1;
  const int nhid = P.nhid;
  const uint32_t smem_s = smem_u32(smem_raw);

  if (warp >= R * G::TEAM_WARPS) {
    // ---------------- issuer warps ----------------
    // Issuer i serves slots [i * SPI, (i + 1) * SPI): both slots of one team (or one slot).  A team requests
    // its slots' layers in a fixed order, which this loop mirrors, so every wait is a blocking try_wait.
    const int issuer = warp - R * G::TEAM_WARPS;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t bh16 = (smem_s + G::OFF_BLOB) >> 4, bl16 = (smem_s + G::OFF_BLOB + G::BH_BYTES) >> 4;
    if constexpr (POOL) {
      // Block pool: issuer = team.  The team's requests are served in the order it makes them; request number rq
      // goes to block rq & 1 of the team, which the epilogue of request rq - 2 must have read ("block read").
      const int team = issuer;
      uint32_t parity = 0, rq = 0;
      for (int unit0 = blockIdx.x * R + team; unit0 < units; unit0 += SL * total_teams) {
        const int nslots = min(SL, (units - unit0 + total_teams - 1) / total_teams);
        for (int it = 0; it < nsteps * nstages; ++it) {
          for (int layer = 0; layer <= nhid; ++layer) {
#pragma unroll 1
            for (int q = 0; q < nslots; ++q) {
              const int ts = team * SL + q;
              const int blk = team * 2 + (int)(rq & 1u);
              const uint32_t use = rq >> 1;
              const uint32_t slot16 = (smem_s + G::OFF_SLOTS + (uint32_t)ts * G::SLOT_BYTES) >> 4;
              if constexpr (SPLIT) {
#pragma unroll 1
                for (uint32_t kb = 0; kb < 2; ++kb) {
                  if (!(P.debug & 64)) {
                    mbar_wait_guarded(&bars[BAR_REQ + kb * TS + ts], parity);
                    if (kb == 0 && use >= 1u) mbar_wait_guarded(&bars[BAR_READ + blk], (use - 1u) & 1u);
                  }
                  fence_after();
#pragma unroll 1
                  for (int m = 0; m < TILES; ++m) {
                    const uint32_t a_hi = slot16 + (uint32_t)m * 128u, a_lo = a_hi + ((4u * G::PLANE) >> 4);
                    const uint32_t d = tmem_u + (uint32_t)((blk * TILES + m) * G::COLS);
                    if (!(P.debug & 1)) {
                      if (layer < nhid) issue_tile_half<G, 32>(a_hi, a_lo, bh16, d, kb);
                      else issue_tile_half<G, NL>(a_hi, a_lo, bl16, d, kb);
                    }
                    if (kb == 1 && elect_one()) mma_commit(&bars[BAR_DONE + blk * TILES + m]);
                    __syncwarp();
                  }
                }
              } else {
              if (!(P.debug & 64)) {
                mbar_wait_guarded(&bars[BAR_REQ + ts], parity);
                if (use >= 1u) mbar_wait_guarded(&bars[BAR_READ + blk], (use - 1u) & 1u);
              }
              fence_after();
#pragma unroll 1
              for (int m = 0; m < TILES; ++m) {
                const uint32_t a_hi = slot16 + (uint32_t)m * 128u, a_lo = a_hi + ((4u * G::PLANE) >> 4);
                const uint32_t d = tmem_u + (uint32_t)((blk * TILES + m) * G::COLS);
                if (!(P.debug & 1)) {
                  if (layer < nhid) issue_tile<G, 32>(a_hi, a_lo, bh16, d);
                  else issue_tile<G, NL>(a_hi, a_lo, bl16, d);
                }
                if (elect_one()) mma_commit(&bars[BAR_DONE + blk * TILES + m]);
                __syncwarp();
              }
              }
              ++rq;
            }
            parity ^= 1u;
          }
        }
      }
    } else {
    const int ts0 = issuer * G::SPI;
    const int team = ts0 >> 1;
    const bool issuer_forcing = eq_forced(P.eq) && P.P > 0 && W.integrating && !(P.debug & 64);
    const int g = blockIdx.x * R + team;
    uint32_t parity = 0;                                       // all served slots flip together
    for (int unit0 = g; unit0 < units; unit0 += 2 * total_teams) {
      const int nslots = unit0 + total_teams < units ? 2 : 1;
      for (int it = 0, step_i = 0, s_i = 0; it < nsteps * nstages; ++it) {
        for (int layer = 0; layer <= nhid; ++layer) {
#pragma unroll 1
          for (int q = 0; q < G::SPI; ++q) {              // rolled: one copy of the issue code per layer kind
            const int ts = ts0 + q;
            if ((ts & 1) >= nslots) continue;
            const uint32_t slot16 = (smem_s + G::OFF_SLOTS + (uint32_t)ts * G::SLOT_BYTES) >> 4;
            if constexpr (SPLIT) {
#pragma unroll 1
              for (uint32_t kb = 0; kb < 2; ++kb) {           // first ci-block as soon as its planes are stored
                if (!(P.debug & 64)) mbar_wait_guarded(&bars[BAR_REQ + kb * TS + ts], parity);
                fence_after();
#pragma unroll 1
                for (int m = 0; m < TILES; ++m) {
                  const uint32_t a_hi = slot16 + (uint32_t)m * 128u, a_lo = a_hi + ((4u * G::PLANE) >> 4);
                  const uint32_t d = tmem_u + (uint32_t)((ts * TILES + m) * G::COLS);
                  if (!(P.debug & 1)) {
                    if (layer < nhid) issue_tile_half<G, 32>(a_hi, a_lo, bh16, d, kb);
                    else issue_tile_half<G, NL>(a_hi, a_lo, bl16, d, kb);
                  }
                  if (kb == 1 && elect_one()) mma_commit(&bars[BAR_DONE + ts * TILES + m]);
                  __syncwarp();
                }
              }
            } else {
            if (!G::TREQ && !(P.debug & 64)) mbar_wait_guarded(&bars[BAR_REQ + ts], parity);
            fence_after();
            constexpr bool WS = DDD1D_TC_WS != 0 && TILES >= 2 && PREC >= 2 && !G::TREQ;
            bool ws_done = false;
            if constexpr (WS) {
              // all tiles of the slot in one pass, the filter planes fetched once per (tap, ci-block); one collector
              // buffer per issuer warp
              if (layer < nhid || NL == 32) {
                const uint32_t d0 = tmem_u + (uint32_t)(ts * TILES * G::COLS);
                if (!(P.debug & 1)) {
                  const uint32_t bb = layer < nhid ? bh16 : bl16;
                  if (issuer == 0) issue_slot_ws<G, 32, 0>(slot16, bb, d0);
                  else issue_slot_ws<G, 32, 1>(slot16, bb, d0);
                }
                if (elect_one())
                  for (int m = 0; m < TILES; ++m) mma_commit(&bars[BAR_DONE + ts * TILES + m]);
                __syncwarp();
                ws_done = true;
              }
            }
            if (!ws_done) {
#pragma unroll 1
            for (int m = 0; m < TILES; ++m) {
              if (G::TREQ) {
                if (!(P.debug & 64)) mbar_wait_guarded(&bars[BAR_REQ + m * TS + ts], parity);
                fence_after();
              }
              const uint32_t a_hi = slot16 + (uint32_t)m * 128u, a_lo = a_hi + ((4u * G::PLANE) >> 4);
              const uint32_t d = tmem_u + (uint32_t)((ts * TILES + m) * G::COLS);
              if (!(P.debug & 1)) {
                if (layer < nhid) issue_tile<G, 32>(a_hi, a_lo, bh16, d);
                else issue_tile<G, NL>(a_hi, a_lo, bl16, d);
              }
              if (elect_one()) mma_commit(&bars[BAR_DONE + ts * TILES + m]);
              __syncwarp();
            }
            }
            }
            if constexpr (G::AMP) {
              // With this slot's first layer of (step, stage s) on the pipe: the slot's forcing amplitudes of stage s
              // of the NEXT step, into the set the team reads a step from now (same expression, same warp-level sums
              // as the team's own first-step code).  After the last stage the three sets are announced.
              if (layer == 0 && issuer_forcing && step_i + 1 < nsteps) {
                const int frow = unit0 + (ts & 1) * total_teams;
                float* const sc = P.scratch + ((size_t)blockIdx.x * TS + ts) * G::SC_STRIDE;
                const float tq = (float)(W.t0 + (double)(step_i + 1) * W.dt + tab.c[s_i] * W.dt);
                forcing_amplitudes_t(P, sc + G::SC_FS + (((step_i + 1) % kFsBuffers) * kMaxStages + s_i) * kFsStride,
                                     W.sample_offset + frow, tq, lane);
                __threadfence_block();
                __syncwarp();
                if (s_i == nstages - 1 && lane == 0) mbar_arrive(&bars[G::BAR_AMP + ts]);
              }
            }
          }
          parity ^= 1u;
        }
        if (++s_i == nstages) { s_i = 0; ++step_i; }
      }
    }
    }
  } else if (!(P.debug & 64)) {
    // ---------------- row teams ----------------
    const int team = warp / G::TEAM_WARPS;
    const int p = tid - team * TEAM;                   // position in the slot's tile(s)
    const int rr = RPT == 1 ? 0 : p / N;               // row within the slot
    const int x = RPT == 1 ? p : p - rr * N;           // grid point
    const int tile = p >> 7;
    const int warp_in_team = __shfl_sync(0xffffffffu, warp - team * G::TEAM_WARPS, 0);
    const bool cons = eq_conservative(P.eq);
    const bool forced = eq_forced(P.eq) && P.P > 0 && (W.op == OP_RHS || W.op == OP_INTEGRATE);
    const bool fast_op = W.op == OP_RHS || W.op == OP_INTEGRATE;
    const bool integrating = W.integrating != 0;       // == (W.op == OP_INTEGRATE), see Work
    const bool nowait = (P.debug & 4) != 0;
    // halo duties: the stage row's 3-point halo (scratch) and the planes' halo copies
    const bool halo_warp = RPT > 1 || warp_in_team == 0 || warp_in_team == G::TEAM_WARPS - 1;
    const int copy_off = halo_copy_offset<G>(p);
    const bool has_copy = halo_warp && copy_off != 0;
    // a tile's MMAs also read two positions of the neighbouring tiles (and the wrapped halo), so the first /
    // last warp of a tile waits for that tile's MMAs too before it overwrites its planes
    const int wit = warp_in_team & 3;
    const int nb_tile = TILES == 1 ? -1 : wit == 0 ? (tile + TILES - 1) % TILES : wit == 3 ? (tile + 1) % TILES : -1;

    const int ts_a = team * SL;                         // the team's first slot
    const int blk_a = team * 2;                         // ... and first accumulator block (= slot with two slots per team)
    const uint32_t mine = smem_s + G::OFF_SLOTS + (uint32_t)ts_a * G::SLOT_BYTES + plane_pos<G>(p);   // shared address
    float* const sc0 = P.scratch + ((size_t)blockIdx.x * TS + ts_a) * G::SC_STRIDE;
    const uint32_t bars_s = smem_u32(bars);
    const uint32_t req0 = bars_s + 8u * (uint32_t)(BAR_REQ + ts_a);            // ci-block 1 of a split request: + TS
    const uint32_t done0 = bars_s + 8u * (uint32_t)(BAR_DONE + blk_a * TILES + tile);
    const uint32_t done_nb0 = bars_s + 8u * (uint32_t)(BAR_DONE + blk_a * TILES + (nb_tile < 0 ? tile : nb_tile));
    const uint32_t read0 = bars_s + 8u * (uint32_t)(BAR_READ + blk_a);         // pool: "block read"
    const uint32_t amp0 = bars_s + 8u * (uint32_t)(G::BAR_AMP + ts_a);         // "next step's amplitudes written"
    uint32_t amp_base = 0;              // arrivals on the slots' amplitude barriers before this unit (their parity)
    const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((blk_a * TILES + tile) * G::COLS);
    uint32_t cq = 0;      // pool: requests of this team consumed so far; request cq sits in block cq & 1, use cq >> 1
    const float* const fbasis_x = P.fbasis + x;         // this point's column of the forcing basis (L1 resident)

    SlotSet<SL> SS;
    const int g = blockIdx.x * R + team;

    for (int unit0 = g; unit0 < units; unit0 += SL * total_teams) {
      const int nslots = min(SL, (units - unit0 + total_teams - 1) / total_teams);
      // ---- load the rows ----
      auto load_slot = [&](int sl, SlotState& S) {
        const int row = (unit0 + sl * total_teams) * RPT + rr;
        const bool live = sl < nslots && row < W.batch;
        if (!live) {
          S.yh = S.yl = 0.f;
        } else if (W.u64) {
          const double v = W.u64[(size_t)row * N + x];
          S.yh = (float)v;
          S.yl = (float)(v - (double)S.yh);
        } else {
          S.yh = __ldg(W.u + (size_t)row * N + x);
          S.yl = 0.f;
        }
        S.k0 = S.k1 = S.k2 = S.k3 = 0.f;
        S.umax = -1.f;                                   // no bound yet: the first stage calibrates
        if (sl >= nslots) return;
        float* sc = sc0 + sl * G::SC_STRIDE;
        if (p < RPT) reinterpret_cast<unsigned int*>(sc + G::SC_BAD)[p] = 0xffffffffu;
        if (p == 0) *reinterpret_cast<unsigned int*>(sc + G::SC_UMAX) = 0u;
      };
      load_slot(0, SS.a);
      load_slot(1, SS.b);
      if constexpr (SL == 3) load_slot(2, SS.c);
      team_sync(team, TEAM);

      // A right-hand side is "started" (phase 0) and "finished" (phase 2) in different turns: a slot's turn is
      // [finish the previous right-hand side | start the next one], so the short phase 2 never stands alone
      // between two waits -- the other slot's last-layer MMAs run under this slot's long phase 0.
      uint32_t done_parity = 0, stage_par = 0;
      bool have_prev = false;
      int prev_step = 0, prev_s = 0;

      // ---- phase 2 (+ the Runge-Kutta update after the last stage) of right-hand side (fstep, fs) ----
      // snap: index of the snapshot this right-hand side completes (-1: none)
      auto finish = [&](int sl, int fstep, int fs, uint32_t fpar, int snap) {
        float* const sc = sc0 + sl * G::SC_STRIDE;
        const float* const rowbuf = sc + G::SC_ROWS + (fpar * 2u * RPT + rr) * G::ROWBUF;
        const int row = (unit0 + sl * total_teams) * RPT + rr;
        const bool live = row < W.batch;
        float u7[kWin];
#pragma unroll
        for (int j = 0; j < kWin; ++j) u7[j] = rowbuf[x + j];
        float f = 0.f;
        if (G::AMP && forced && integrating && fs == 0 && fstep > 0)      // written by the issuer during step fstep - 1
          mbar_wait_spin(amp0 + 8u * (uint32_t)sl, (amp_base + (uint32_t)(fstep - 1)) & 1u);
        if (forced) {
          const float* amp = sc + G::SC_FS + rr * kFsWords + ((fstep % kFsBuffers) * kMaxStages + fs) * kFsStride;
          if (P.M <= 4) {
            const float4 a = *reinterpret_cast<const float4*>(amp), b = *reinterpret_cast<const float4*>(amp + kMaxModes);
            float c[4], d[4];                                  // amplitudes beyond M are zero; their basis loads are skipped
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              c[m] = m < P.M ? __ldg(fbasis_x + (size_t)m * N) : 0.f;
              d[m] = m < P.M ? __ldg(fbasis_x + (size_t)(P.M + m) * N) : 0.f;
            }
            f = fmaf(a.x, c[0], f); f = fmaf(b.x, d[0], f);
            f = fmaf(a.y, c[1], f); f = fmaf(b.y, d[1], f);
            f = fmaf(a.z, c[2], f); f = fmaf(b.z, d[2], f);
            f = fmaf(a.w, c[3], f); f = fmaf(b.w, d[3], f);
          } else {
#pragma unroll 1
            for (int m = 0; m < P.M; ++m) {
              f = fmaf(amp[m], __ldg(fbasis_x + (size_t)m * N), f);
              f = fmaf(amp[kMaxModes + m], __ldg(fbasis_x + (size_t)(P.M + m) * N), f);
            }
          }
        }
        const float bound1 = fmaf(P.w1abs, DDD1D_SLOT_GET(SS, sl, umax), P.b1abs);
        const float s_last = nhid > 0 ? scale_for(fmaf(P.whabs, bound1, P.bhabs)) : scale_for(bound1);
        const float inv_last = pow2_inverse(s_last) * P.inv_sw_last;
        // where the slot's accumulators are: its own block, or (pool) the block its request was dealt
        const int blk = POOL ? (int)(cq & 1u) : sl;
        const uint32_t dpar = POOL ? ((cq >> 1) & 1u) : done_parity;
        if (!nowait) mbar_wait_spin(done0 + 8u * (uint32_t)(blk * TILES), dpar);
        if (nb_tile >= 0 && !nowait) mbar_wait_spin(done_nb0 + 8u * (uint32_t)(blk * TILES), dpar);     // before the next stage rewrites the planes
        fence_after();
        // window coefficients = accumulators + folded bias, then the stencil dot products (model.py:536-548);
        // sixteen columns at a time keeps the register peak low
        float dv[kMaxD];
#pragma unroll
        for (int d = 0; d < kMaxD; ++d) dv[d] = 0.f;
        const uint32_t taddr = taddr0 + (uint32_t)(blk * TILES * G::COLS);
#pragma unroll
        for (int half = 0; half < NL / 16; ++half) {
          float2 v[8];
          tmem_read_pairs<PREC>(taddr + 16 * half, taddr + NL + 16 * half, v);
          float c16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int q = 16 * half + i;                    // column = derivative * 7 + window slot
            c16[i] = fmaf(i & 1 ? v[i >> 1].y : v[i >> 1].x, inv_last, P.bl[q]);
            if (q < kMaxD * kWin) dv[q / kWin] = fmaf(c16[i], u7[q % kWin], dv[q / kWin]);
          }
          if (!integrating && W.op == OP_COEF && live) {
#pragma unroll
            for (int i = 0; i < 16; ++i) export_coefficient(P, W, c16[i], 16 * half + i, (size_t)row * N + x);
          }
        }
        fence_before();
        if constexpr (POOL) {                     // the block may be overwritten: its next request can be served
          __syncwarp();
          if (lane == 0) mbar_arrive_s(read0 + 8u * (uint32_t)blk);
          ++cq;
        }
        if (!integrating && !fast_op) {
          if (W.op == OP_DERIV && live) export_derivatives(P, W, dv[0], dv[1], dv[2], dv[3], (size_t)row * N + x);
          return;
        }
        float r;                      // equations.py:269-587, op by op like the float32 graph (equation_point)
        if (P.plain_burgers) r = __fsub_rn(__fmul_rn(P.eta, dv[1]), __fmul_rn(u7[kHalo], dv[0]));
        else if (P.plain_kdv) r = __fsub_rn(__fmul_rn(__fmul_rn(-6.f, u7[kHalo]), dv[0]), dv[1]);
        else if (P.plain_ks) r = __fsub_rn(__fsub_rn(__fmul_rn(-u7[kHalo], dv[0]), dv[2]), dv[1]);
        else r = equation_point(P.eq, u7[kHalo], dv, P.eta);
        if (cons) {
          float* const flux = sc + G::SC_FLUX;
          flux[p] = r;
          team_sync(team, TEAM);
          const float fwd = flux[x + 1 == N ? p + 1 - N : p + 1];
          r = -__fmul_rn(P.inv_dx, __fsub_rn(fwd, r));
        }
        if (forced) r = __fadd_rn(r, f);
        if (!integrating) {           // OP_RHS
          if (live) {
            if (W.out64) W.out64[(size_t)row * N + x] = (double)r;
            else W.out[(size_t)row * N + x] = r;
          }
          return;
        }
        // the stage derivatives rotate: newest first
        const float kn1 = DDD1D_SLOT_GET(SS, sl, k0), kn2 = DDD1D_SLOT_GET(SS, sl, k1), kn3 = DDD1D_SLOT_GET(SS, sl, k2);
        DDD1D_SLOT_PUT(SS, sl, k3, kn3);
        DDD1D_SLOT_PUT(SS, sl, k2, kn2);
        DDD1D_SLOT_PUT(SS, sl, k1, kn1);
        DDD1D_SLOT_PUT(SS, sl, k0, r);
        if (fs != nstages - 1) return;
        // ---- the step is complete: y += dt * sum b k ----
        // increment dt * b_j * k_j as a float pair: exact products (FMA residual) of the float-float constants,
        // highs summed with TwoSum -- the float64 sum of the reference to ~2^-48
        float ih = 0.f, il = 0.f;
        // k_j in stage order: the newest is k_{nstages-1}
        float kk[kMaxStages];
        if (nstages == 3) { kk[0] = kn2; kk[1] = kn1; kk[2] = r; kk[3] = 0.f; }
        else if (nstages == 2) { kk[0] = kn1; kk[1] = r; kk[2] = kk[3] = 0.f; }
        else if (nstages == 4) { kk[0] = kn3; kk[1] = kn2; kk[2] = kn1; kk[3] = r; }
        else { kk[0] = r; kk[1] = kk[2] = kk[3] = 0.f; }
#pragma unroll
        for (int j = 0; j < kMaxStages; ++j) {
          if (j >= nstages) break;
          const float ph = W.bdt_hi[j] * kk[j];
          const float pl = fmaf(W.bdt_hi[j], kk[j], -ph) + W.bdt_lo[j] * kk[j];
          float sh, e;
          two_sum(ih, ph, sh, e);
          ih = sh;
          il += e + pl;
        }
        float yh, e;
        two_sum(DDD1D_SLOT_GET(SS, sl, yh), ih, yh, e);
        const float yl = DDD1D_SLOT_GET(SS, sl, yl) + (e + il);
        const float y = yh + yl;                       // renormalise: y = float(yh + yl)
        DDD1D_SLOT_PUT(SS, sl, yl, W.state_f32 ? 0.f : yl - (y - yh));      // (float32 carry: tf odeint_fixed, model.py:138-159)
        DDD1D_SLOT_PUT(SS, sl, yh, y);
        if (!isfinite(y))       // first step at which the row left the finite range (rare, so an atomic is fine)
          atomicMin(reinterpret_cast<unsigned int*>(sc + G::SC_BAD) + rr, (unsigned int)fstep);
        if (snap >= 0 && live) W.snaps[((size_t)snap * W.batch + row) * N + x] = y;
      };

      // ---- phase 0 of right-hand side (step, s): stage value, exchange, first layer, planes, request ----
      // aN, aP, aQ: dt * a[s][s-1], a[s][s-2], a[s][s-3] (0 beyond the stage): the tableau row, newest derivative first
      auto start = [&](int sl, int step, int s, float aN, float aP, float aQ) {
        float* const sc = sc0 + sl * G::SC_STRIDE;
        float* const rowbuf = sc + G::SC_ROWS + (stage_par * 2u * RPT + rr) * G::ROWBUF;   // raw; normalised at + RPT * ROWBUF
        // stage value y + dt * sum a_j k_j rounded to float32 (integrate.py:57-60,71): the increment in float32
        // (its rounding is 1e-7 of an increment that is itself far below half an ulp of y), added low part first
        // sum in stage order (a[s][0] k_0 first), as before the derivatives rotated: at stage 1 the two inner terms are
        // exact zeros, at stage 2 the innermost one
        const float inc = s == 1 ? aN * DDD1D_SLOT_GET(SS, sl, k0)
                        : s == 2 ? fmaf(aN, DDD1D_SLOT_GET(SS, sl, k0), aP * DDD1D_SLOT_GET(SS, sl, k1))
                                 : fmaf(aN, DDD1D_SLOT_GET(SS, sl, k0),
                                        fmaf(aP, DDD1D_SLOT_GET(SS, sl, k1), aQ * DDD1D_SLOT_GET(SS, sl, k2)));
        const float yh0 = DDD1D_SLOT_GET(SS, sl, yh);
        const float us = s == 0 ? yh0 : yh0 + (DDD1D_SLOT_GET(SS, sl, yl) + inc);
        const float usn = __fdiv_rn(us, P.sigma);            // model.py:450-451
        rowbuf[x + kHalo] = us;
        rowbuf[RPT * G::ROWBUF + x + kHalo] = usn;
        if (halo_warp) {
          if (x < kHalo) { rowbuf[x + kHalo + N] = us; rowbuf[RPT * G::ROWBUF + x + kHalo + N] = usn; }
          if (x >= N - kHalo) { rowbuf[x + kHalo - N] = us; rowbuf[RPT * G::ROWBUF + x + kHalo - N] = usn; }
        }
        if (forced && s == 0 && step == 0) {
          // the first step's amplitudes; later steps get theirs one step ahead, after the planes are stored
          const double t0 = W.t0;
          for (int task = warp_in_team; task < RPT * nstages; task += G::TEAM_WARPS) {
            int fr = 0, sq = task;                     // task = fr * nstages + sq without a division
            if (RPT > 1) while (sq >= nstages) { sq -= nstages; ++fr; }
            const int frow = (unit0 + sl * total_teams) * RPT + fr;
            const float tq = (float)(W.op == OP_INTEGRATE ? t0 + tab.c[sq] * W.dt : W.t0);
            if (frow < W.batch)
              forcing_amplitudes_t(P, sc + G::SC_FS + fr * kFsWords + sq * kFsStride, W.sample_offset + frow, tq, lane);
          }
        }
        float umax = DDD1D_SLOT_GET(SS, sl, umax);
        bool keep = team_sync_all(team, TEAM, fabsf(usn) <= umax);    // team-uniform; NaN rows fail every stage
        if (keep && s == 0 && (step & 15) == 15)                      // now and then: has the slot decayed far below
          keep = !team_sync_all(team, TEAM, fabsf(usn) < umax * (1.f / 256.f));   // its bound (lo planes would thin out)?
        if (!keep) {
          umax = recalibrate(reinterpret_cast<unsigned int*>(sc + G::SC_UMAX), usn, team, TEAM, lane, p == 0);
          DDD1D_SLOT_PUT(SS, sl, umax, umax);
        }
        const float bound1 = fmaf(P.w1abs, umax, P.b1abs);   // |h1| <= |b1| + sum|W1| * max|u/sigma|
        const float s_act = scale_for(bound1);

        // ---- first layer 1 -> 32 on the CUDA cores: relu(s * (W u + b)) = s * relu(W u + b), s a power of two;
        //      packed FP32x2 arithmetic (FFMA2 with the input broadcast and the filter pair from the constant bank) ----
        float un[kTaps];
#pragma unroll
        for (int k = 0; k < kTaps; ++k) un[k] = rowbuf[RPT * G::ROWBUF + x + k + 1] * s_act;
        const uint32_t my = mine + (uint32_t)sl * G::SLOT_BYTES;
        const float2 s2 = make_float2(s_act, s_act);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float2 h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = fmul2(pair_at(P.b1, 8 * c8 + 2 * i), s2);
#pragma unroll
          for (int k = 0; k < kTaps; ++k) {
            const float2 u2 = make_float2(un[k], un[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = ffma2(u2, pair_at(P.w1, k * kF + 8 * c8 + 2 * i), h[i]);
          }
          store_chunk8<G>(my, c8, copy_off, has_copy, h);        // (ReLU inside the conversion)
          if (SPLIT && c8 == 1) {                                // the first ci-block's planes are complete
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_s(req0 + 8u * (uint32_t)sl);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (G::TREQ) {
          if (lane == 0) {
            mbar_arrive_s(req0 + 8u * (uint32_t)(tile * TS + sl));
            if (nb_tile >= 0) mbar_arrive_s(req0 + 8u * (uint32_t)(nb_tile * TS + sl));
          }
        } else if (lane == 0) {
          mbar_arrive_s(req0 + 8u * (uint32_t)((SPLIT ? TS : 0) + sl));
        }
        if (!G::AMP && forced && s == 0 && W.op == OP_INTEGRATE && step + 1 < nsteps) {
          // Forcing amplitudes of the NEXT step, off the critical path (the MMAs just requested are running).
          // One warp per (row, stage): one forcing term per lane, mode amplitudes by warp sums.  Three sets
          // rotate: set (step + 1) % 3 was last read in step - 2, and every warp that gets here has passed a
          // barrier of step `step`, which no warp reaches before it has finished step - 1.
          for (int task = warp_in_team; task < RPT * nstages; task += G::TEAM_WARPS) {
            int fr = 0, sq = task;
            if (RPT > 1) while (sq >= nstages) { sq -= nstages; ++fr; }
            const int frow = (unit0 + sl * total_teams) * RPT + fr;
            const float tq = (float)(W.t0 + (double)(step + 1) * W.dt + tab.c[sq] * W.dt);
            if (frow < W.batch)
              forcing_amplitudes_t(P, sc + G::SC_FS + fr * kFsWords + (((step + 1) % kFsBuffers) * kMaxStages + sq) * kFsStride,
                                   W.sample_offset + frow, tq, lane);
          }
        }
      };

      // ---- phase 1: epilogue of a hidden tensor layer, planes rewritten in place, next layer requested ----
      auto hidden = [&](int sl) {
        const float bound1 = fmaf(P.w1abs, DDD1D_SLOT_GET(SS, sl, umax), P.b1abs);
        // accumulators carry (activation scale x filter scale); the next planes get their own scale
        const float inv = pow2_inverse(scale_for(bound1)) * P.inv_sw_hid;
        const float s_act = scale_for(fmaf(P.whabs, bound1, P.bhabs));   // |h2| <= |b2| + sum|W2| max|h1|
        const float2 inv2 = make_float2(inv, inv), s2 = make_float2(s_act, s_act);
        const uint32_t my = mine + (uint32_t)sl * G::SLOT_BYTES;
        const int blk = POOL ? (int)(cq & 1u) : sl;
        const uint32_t dpar = POOL ? ((cq >> 1) & 1u) : done_parity;
        const uint32_t taddr = taddr0 + (uint32_t)(blk * TILES * G::COLS);
        if (!nowait) mbar_wait_spin(done0 + 8u * (uint32_t)(blk * TILES), dpar);
        fence_after();
        if (nb_tile >= 0 && !nowait) mbar_wait_spin(done_nb0 + 8u * (uint32_t)(blk * TILES), dpar);
        if constexpr (SPLIT) {
          // The last layer's first ci-block is requested after half of the planes: its MMAs overwrite this block,
          // so all 32 (+ 32) accumulator columns are in registers before the first request.
          float2 acc[2][8];
          tmem_read_pairs<PREC>(taddr, taddr + 32, acc[0]);
          tmem_read_pairs<PREC>(taddr + 16, taddr + 48, acc[1]);
          fence_before();
          if (POOL) {
            __syncwarp();
            if (lane == 0) mbar_arrive_s(read0 + 8u * (uint32_t)blk);
            ++cq;
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float2 h[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 v = ffma2(acc[half][4 * c + i], inv2, pair_at(P.bh, 16 * half + 8 * c + 2 * i));
                h[i] = fmul2(v, s2);                    // relu(v) * s = relu(v * s); the conversion clamps
              }
              store_chunk8<G>(my, 2 * half + c, copy_off, has_copy, h);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_s(req0 + 8u * (uint32_t)(half * TS + sl));
          }
        } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {            // 16 channels at a time keeps the register peak low
          float2 acc[8];
          tmem_read_pairs<PREC>(taddr + 16 * half, taddr + 32 + 16 * half, acc);
          if (POOL && half == 1) {                        // all accumulators are in registers: release the block early
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_s(read0 + 8u * (uint32_t)blk);
            ++cq;
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float2 h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 v = ffma2(acc[4 * c + i], inv2, pair_at(P.bh, 16 * half + 8 * c + 2 * i));
              h[i] = fmul2(v, s2);                    // relu(v) * s = relu(v * s); the conversion clamps
            }
            store_chunk8<G>(my, 2 * half + c, copy_off, has_copy, h);
          }
        }
        fence_before();
        fence_async_smem();
        __syncwarp();
        if (G::TREQ) {
          if (lane == 0) {
            mbar_arrive_s(req0 + 8u * (uint32_t)(tile * TS + sl));
            if (nb_tile >= 0) mbar_arrive_s(req0 + 8u * (uint32_t)(nb_tile * TS + sl));
          }
        } else if (lane == 0) {
          mbar_arrive_s(req0 + 8u * (uint32_t)sl);
        }
        }
      };

      int until_save = W.save_every;       // snapshots: a down-counter; the index is divided out only when one is due
      for (int step = 0; step < nsteps; ++step) {
        for (int s = 0; s < nstages; ++s) {
          // this stage's row of the tableau, newest derivative first, fetched once for both slots
          const float aN = s > 0 ? W.adt[s][s - 1] : 0.f, aP = s > 1 ? W.adt[s][s - 2] : 0.f,
                      aQ = s > 2 ? W.adt[s][s - 3] : 0.f;
          int snap = -1;
          if (have_prev && s == 0 && --until_save == 0) {               // the previous right-hand side completed a step
            until_save = W.save_every;
            snap = step / W.save_every - 1;                             // `step` steps are complete
          }
#pragma unroll 1
          for (int sl = 0; sl < nslots; ++sl) {
            if (have_prev) finish(sl, prev_step, prev_s, stage_par ^ 1u, snap);
            start(sl, step, s, aN, aP, aQ);
          }
          if (have_prev) done_parity ^= 1u;
          for (int l = 0; l < nhid; ++l) {
#pragma unroll 1
            for (int sl = 0; sl < nslots; ++sl) {
              hidden(sl);
            }
            done_parity ^= 1u;
          }
          have_prev = true;
          prev_step = step;
          prev_s = s;
          stage_par ^= 1u;
        }
      }
      {
        const int snap = (W.op == OP_INTEGRATE && --until_save == 0) ? nsteps / W.save_every - 1 : -1;
#pragma unroll 1
        for (int sl = 0; sl < nslots; ++sl) {
          finish(sl, prev_step, prev_s, stage_par ^ 1u, snap);
        }
      }
      if (W.op == OP_INTEGRATE && W.first_bad) {
        team_sync(team, TEAM);
#pragma unroll
        for (int sl = 0; sl < SL; ++sl) {
          if (sl >= nslots) continue;
          const int row = (unit0 + sl * total_teams) * RPT + rr;
          const unsigned int fb = reinterpret_cast<unsigned int*>(sc0 + sl * G::SC_STRIDE + G::SC_BAD)[rr];
          if (x == 0 && row < W.batch) W.first_bad[row] = (fb == 0xffffffffu) ? -1 : (int)fb;
        }
      }
      team_sync(team, TEAM);       // the next rows reuse the slot regions
      if (G::AMP && forced && integrating) amp_base += (uint32_t)(nsteps - 1);
    }
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

}  // namespace tc
}  // namespace ddd1d
