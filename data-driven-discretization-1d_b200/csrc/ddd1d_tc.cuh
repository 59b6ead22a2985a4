// Tensor-core (tcgen05 / TMEM) version of the learned-coefficient row kernel, sm_100a.
//
// The conv stack is 96 % of the FLOPs and is an implicit GEMM per 128-position tile:
//   hidden layer : D[128 x 32] += sum_{tap k, ci-block} A_k[128 x 8] * B_k[8 x 32]   (K = 5*32)
//   last layer   : D[128 x NL] += ...                                                 (NL = 16 | 32)
// A = activations kept in shared memory as K-major "chunk planes" [ci/4][position][4 floats]
// (no swizzle), so the tap shift k is just +16 B on the descriptor start address and the periodic
// halo is two extra positions per plane.  B = filters pre-packed on the host in the same canonical
// layout.  FP32 fidelity on a TF32 pipe comes from the 3xTF32 split: x = hi + lo (both rounded to
// TF32) and hi*Whi + lo*Whi + hi*Wlo accumulated in FP32 in TMEM (the dropped lo*Wlo term is 2^-22
// relative).  B holds [Whi | Wlo] side by side, so one MMA of width 2*NB produces hi*Whi (the "main"
// columns) and hi*Wlo (the "cross" columns) and a second one of width NB adds lo*Whi to the cross
// columns: two instructions per (tap, ci-block) instead of three.  The tensor core truncates when it
// adds into an accumulator (measured: error grows linearly with the number of accumulate steps, see
// profiles/r01/tc_precision.txt), so the large main terms get their own columns, split once more
// into even and odd taps (12 + 8 steps), and the epilogue adds the four partial sums in FP32.  The polynomial-accuracy projection is folded into the last layer's filters on
// the host (W3' = W3 . nullspace, window form), so the last epilogue reads stencil coefficients
// straight out of TMEM.
//
// Warp roles (one CTA per SM, persistent): R "row teams" of N threads (thread <-> grid point; the
// team's warps are 4-aligned so each warp reads its own TMEM lane quadrant) run the whole
// Runge-Kutta program of their row.  The first warp of each 128-position tile issues that tile's
// tcgen05.mma once the team's planes are complete and signals completion with
// tcgen05.commit -> mbarrier; 16 warps per CTA keep 128 registers per thread (no spills).  While one team runs an epilogue on the CUDA
// cores, the tensor pipe works on another team's tile.
#pragma once
#include <cuda_fp16.h>

#include "ddd1d_device.cuh"

namespace ddd1d {
namespace tc {

constexpr int kF = 32;             // hidden width this path is built for
constexpr int kTaps = 5;
constexpr int kChunks = kF / 4;    // 16-byte chunks along ci
constexpr int kForcingStride = 2 * kMaxModes + 3 * kMaxForcing;   // floats of forcing scratch per RK stage
constexpr long long kSpinCycles = 4000000000ll;   // ~2 s at 1.9 GHz: a protocol bug traps instead of hanging

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) leading byte offset>>4 (between the two 16-B K chunks of one MMA)
//   [32,46) stride byte offset>>4 (between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::tf32, FP32 accumulate, A and B K-major (cute::UMMA::InstrDescriptor):
//   c_format F32 = 1 @4 | a_format TF32 = 2 @7 | b_format TF32 = 2 @10 | N>>3 @17 | M>>4 @24
__device__ __forceinline__ uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
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
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// mbarrier helpers with a spin guard: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long start = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > kSpinCycles) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void team_sync(int team, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(threads) : "memory");
}

// x = hi + lo with hi = x rounded to TF32 (round-half-up on the magnitude) and lo = the exact
// remainder, itself rounded to TF32, so the tensor core's truncation of its inputs never acts.
__device__ __forceinline__ float round_tf32(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = round_tf32(v);
  lo = round_tf32(v - hi);
}

// ---- shared layouts ----------------------------------------------------------------------------
// activation planes of one team: plane c (ci = 4c..4c+3), position x at byte (x + 2) * 16
// filters: hidden  Bh[(tap*8 + chunk) * 512 + co*16 + (ci%4)*4],  last  Bl[(tap*8 + chunk) * NL*16 + ...]

struct TcView {
  uint64_t* bars;        // [0] blob copy, [1+t] request (count N), [1+R+t] done (count 1)
  uint32_t* tmem_slot;
  float* blob;
  unsigned char* team_base;
};

// One elected lane of a converged warp (CUTLASS's elect_one_sync): keeps the surrounding values in
// uniform registers, so tcgen05.mma takes its descriptors without per-instruction R2UR shuffles.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mma_tf32_split(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint32_t instr_desc_f16(int m, int n) {   // kind::f16, fp16 x fp16 -> fp32
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Issue every MMA of one layer for one 128-position tile.  Must be called by a converged warp with
// warp-uniform arguments; one elected lane issues.
//   act_hi/act_lo : shared addresses of the tile's team planes (position 0 of the row)
//   b             : shared address of the layer's [Whi | Wlo] planes, b_plane_bytes = 2*NB*16
//   d_col         : TMEM column of the tile's accumulator block; layout
//                   [even taps: main NB | cross NB][odd taps: main NB | cross NB]
// F16 = false: TF32 planes, 8 chunk planes of 4 floats, K = 8 per MMA (4 ci-blocks per tap);
// F16 = true : fp16 planes, 4 chunk planes of 8 halfs, K = 16 per MMA (2 ci-blocks per tap).
template <bool F16>
__device__ __forceinline__ void issue_layer(uint32_t act_hi, uint32_t act_lo, uint32_t plane_bytes, uint32_t b,
                                            uint32_t b_plane_bytes, int tile, uint32_t d_col, int nb) {
  constexpr int kPlanes = F16 ? kChunks / 2 : kChunks;
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);          // SBO = 128 B, version 1 (bits 32..47)
  const uint32_t plane16 = plane_bytes >> 4, bplane16 = b_plane_bytes >> 4;
  const uint32_t ah0 = (((act_hi >> 4) + (uint32_t)tile * 128u) & 0x3FFFu) | (plane16 << 16);
  const uint32_t al0 = (((act_lo >> 4) + (uint32_t)tile * 128u) & 0x3FFFu) | (plane16 << 16);
  const uint32_t b0 = ((b >> 4) & 0x3FFFu) | (bplane16 << 16);
  const uint32_t idesc_wide = F16 ? instr_desc_f16(128, 2 * nb) : instr_desc_tf32(128, 2 * nb);
  const uint32_t idesc_narrow = F16 ? instr_desc_f16(128, nb) : instr_desc_tf32(128, nb);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
      const uint32_t d_main = d_col + (uint32_t)((k & 1) * 2 * nb);
      const uint32_t d_cross = d_main + (uint32_t)nb;
#pragma unroll
      for (int kb = 0; kb < kPlanes / 2; ++kb) {
        const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
        const uint32_t bo = (uint32_t)(k * kPlanes + 2 * kb) * bplane16;
        const uint32_t first = (k < 2 && kb == 0) ? 0u : 1u;     // first touch of the even / odd block
        if (F16) {
          mma_f16_split(d_main, ah0 + ao, b0 + bo, desc_hi, idesc_wide, first);    // hi*[Wh|Wl'] -> main | cross
          mma_f16_split(d_cross, al0 + ao, b0 + bo, desc_hi, idesc_narrow, 1u);    // lo'*Wh      -> cross
        } else {
          mma_tf32_split(d_main, ah0 + ao, b0 + bo, desc_hi, idesc_wide, first);   // hi*[Whi|Wlo] -> main | cross
          mma_tf32_split(d_cross, al0 + ao, b0 + bo, desc_hi, idesc_narrow, 1u);   // lo*Whi       -> cross
        }
      }
    }
  }
  __syncwarp();
}

// ---- fp16 x 2 planes -----------------------------------------------------------------------------
// v (already multiplied by the layer's power-of-two scale) = hi + lo' * 2^-11 with hi = fp16(v) and
// lo' = fp16((v - hi) * 2^11): 22 significant bits like the 3xTF32 split, but 2 bytes per element, so one
// 4 KB A read covers K = 16.  Static bounds on the activations (operator norms x the row's max |u/sigma|)
// keep v below 2^14, far from fp16's range limits; scales are powers of two, i.e. exact.
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  return (uint32_t)__half_as_ushort(__float2half_rn(a)) | ((uint32_t)__half_as_ushort(__float2half_rn(b)) << 16);
}
__device__ __forceinline__ float lo_part(float v) { return (v - __half2float(__float2half_rn(v))) * 2048.f; }

__device__ __forceinline__ void store_split_f16(unsigned char* hi_plane, unsigned char* lo_plane, int x, int N,
                                                bool edge, const float (&v)[8]) {
  uint4 h, l;
  h.x = pack_half2(v[0], v[1]); h.y = pack_half2(v[2], v[3]); h.z = pack_half2(v[4], v[5]); h.w = pack_half2(v[6], v[7]);
  l.x = pack_half2(lo_part(v[0]), lo_part(v[1])); l.y = pack_half2(lo_part(v[2]), lo_part(v[3]));
  l.z = pack_half2(lo_part(v[4]), lo_part(v[5])); l.w = pack_half2(lo_part(v[6]), lo_part(v[7]));
  *reinterpret_cast<uint4*>(hi_plane + (size_t)(x + 2) * 16) = h;
  *reinterpret_cast<uint4*>(lo_plane + (size_t)(x + 2) * 16) = l;
  if (edge) {
    if (x < 2) {
      *reinterpret_cast<uint4*>(hi_plane + (size_t)(x + 2 + N) * 16) = h;
      *reinterpret_cast<uint4*>(lo_plane + (size_t)(x + 2 + N) * 16) = l;
    }
    if (x >= N - 2) {
      *reinterpret_cast<uint4*>(hi_plane + (size_t)(x + 2 - N) * 16) = h;
      *reinterpret_cast<uint4*>(lo_plane + (size_t)(x + 2 - N) * 16) = l;
    }
  }
}
// largest power of two s with bound * s < 2^14 (bound > 0), capped so that tiny bounds stay finite
__device__ __forceinline__ float scale_for(float bound) {
  int e;
  frexpf(fmaxf(bound, 1e-30f), &e);            // bound = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, min(14 - e, 60));
}

// store 4 consecutive channels of one position into a plane (+ its wrapped halo copy).  `edge` is
// warp-uniform: only the first and last warp of a team own positions that feed the halo.
__device__ __forceinline__ void store_chunk(unsigned char* plane, int x, int N, bool edge, float4 v) {
  *reinterpret_cast<float4*>(plane + (size_t)(x + 2) * 16) = v;
  if (edge) {
    if (x < 2) *reinterpret_cast<float4*>(plane + (size_t)(x + 2 + N) * 16) = v;
    if (x >= N - 2) *reinterpret_cast<float4*>(plane + (size_t)(x + 2 - N) * 16) = v;
  }
}

__device__ __forceinline__ void store_split(unsigned char* hi_plane, unsigned char* lo_plane, int x, int N,
                                            bool edge, float a, float b, float c, float d) {
  float4 h, l;
  split_tf32(a, h.x, l.x);
  split_tf32(b, h.y, l.y);
  split_tf32(c, h.z, l.z);
  split_tf32(d, h.w, l.w);
  store_chunk(hi_plane, x, N, edge, h);
  store_chunk(lo_plane, x, N, edge, l);
}


__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// sixteen outputs = (even main + odd main) + (even cross + odd cross); two loads in flight at a time
// keeps the register peak at 48
// cross_scale: 1 for the TF32 planes, 2^-11 for the fp16 planes (their cross terms carry a 2^11 factor)
__device__ __forceinline__ void tmem_sum4x16(uint32_t t_em, uint32_t t_om, uint32_t t_ec, uint32_t t_oc, float* v,
                                             float cross_scale = 1.f) {
  uint32_t a[16], b[16];
  tmem_ld16_issue(t_em, a);
  tmem_ld16_issue(t_om, b);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(b[i]);
  tmem_ld16_issue(t_ec, a);
  tmem_ld16_issue(t_oc, b);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(a[i]) + __uint_as_float(b[i]), cross_scale, v[i]);
}
// NB = 32: block layout [even main 32 | even cross 32 | odd main 32 | odd cross 32]
__device__ __forceinline__ void tmem_sum32(uint32_t taddr, float (&v)[32], float cross_scale = 1.f) {
  tmem_sum4x16(taddr, taddr + 64, taddr + 32, taddr + 96, v, cross_scale);
  tmem_sum4x16(taddr + 16, taddr + 80, taddr + 48, taddr + 112, v + 16, cross_scale);
}
// NB = 16: block layout [even main 16 | even cross 16 | odd main 16 | odd cross 16]
__device__ __forceinline__ void tmem_sum16(uint32_t taddr, float (&v)[16], float cross_scale = 1.f) {
  tmem_sum4x16(taddr, taddr + 32, taddr + 16, taddr + 48, v, cross_scale);
}

// Last-layer epilogue for one grid point: window coefficients = TMEM accumulators + folded bias,
// then the stencil dot products (model.py:536-548).  NLV = TMEM columns of the last layer.
template <int NLV>
__device__ __forceinline__ void last_epilogue(const Params& P, const Work& W, uint32_t taddr,
                                              const float (&u7)[kWin],
                                              int row, int x, float (&dv)[kMaxD], float cross_scale, float inv_scale) {
  float cfv[NLV];
  if (NLV == 16) tmem_sum16(taddr, reinterpret_cast<float(&)[16]>(cfv), cross_scale);
  else tmem_sum32(taddr, reinterpret_cast<float(&)[32]>(cfv), cross_scale);
  fence_before();
  const int N = P.N;
#pragma unroll
  for (int d = 0; d < kMaxD; ++d) {
    dv[d] = 0.f;
    if (d * kWin + kWin > NLV || d >= P.D) continue;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kWin; ++j) {
      const float cf = fmaf(cfv[d * kWin + j], inv_scale, P.tc_bl[d * kWin + j]);
      sum = fmaf(cf, u7[j], sum);
      if (W.op == OP_COEF) {
        const int i = j - P.wshift;
        if (i >= 0 && i < P.S) W.out[(((size_t)row * N + x) * P.D + d) * P.S + i] = cf;
      }
    }
    dv[d] = sum;
    if (W.op == OP_DERIV) W.out[((size_t)row * N + x) * P.D + d] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(512, 1) tc_row_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P.N, R = P.tc_teams, tiles = N / 128;
  const int team_warps = N / 32;
  const bool is_alloc_warp = warp == 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + P.tc_off_slot);
  float* blob = reinterpret_cast<float*>(smem_raw + P.off_blob);
  const uint32_t plane_bytes = (uint32_t)(N + 4) * 16u;
  const int NL = P.tc_nlast;                         // 16 or 32 columns for the last layer
  const int hidden_tc_layers = P.nlayers - 2;        // layers between the first and the last

  Tableau* tab_s = reinterpret_cast<Tableau*>(smem_raw + P.tc_off_tab);
  if (tid == 0) {
    *tab_s = make_tableau(W.scheme);
    reinterpret_cast<uint32_t*>(smem_raw + P.tc_off_slot + 4)[0] = 0u;
    reinterpret_cast<uint32_t*>(smem_raw + P.tc_off_slot + 4)[1] = 0u;
    mbar_init(&bars[0], 1);
    for (int t = 0; t < R; ++t) {
      mbar_init(&bars[1 + t], (uint32_t)N);
      mbar_init(&bars[1 + R + t], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)P.blob_floats * 4u;
    mbar_expect_tx(&bars[0], bytes);
    bulk_copy_g2s(blob, P.blob, bytes, &bars[0]);
  }
  if (is_alloc_warp) tmem_alloc(tmem_slot, 512);      // 4 tiles x 128 columns
  mbar_wait_guarded(&bars[0], 0);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_teams = gridDim.x * R;

  {
    // ---------------- row team ----------------
    const int team = warp / team_warps;
    const int x = tid - team * N;                      // this thread's grid point
    unsigned char* tb = smem_raw + P.tc_off_team0 + (size_t)team * P.tc_team_stride;
    unsigned char* act_hi = tb + P.tc_t_act_hi;
    unsigned char* act_lo = tb + P.tc_t_act_lo;
    float* ust = reinterpret_cast<float*>(tb + P.tc_t_ust);
    float* unr = ust + (N + 2 * kHalo + 2);              // the same row divided by sigma
    uint32_t* umax_w = reinterpret_cast<uint32_t*>(tb + P.tc_t_umax);   // per-warp max |u/sigma| (fp16 planes)
    float* kst = reinterpret_cast<float*>(tb + P.tc_t_k);
    float* flux = reinterpret_cast<float*>(tb + P.tc_t_flux);
    float* fs = reinterpret_cast<float*>(tb + P.tc_t_fs);
    uint64_t* req = &bars[1 + team];
    uint64_t* done = &bars[1 + R + team];
    uint32_t done_parity = 0;
    const int tile = x >> 7;
    const int warp_in_team = __shfl_sync(0xffffffffu, warp - team * team_warps, 0);
    const bool edge = warp_in_team == 0 || warp_in_team == team_warps - 1;
    // The first warp of every 128-position tile also issues that tile's MMAs (asynchronous: it then
    // waits for completion like everybody else).  All issue-side values are warp-uniform.
    const bool issuer = warp_in_team == 0;
    const int team_u = __shfl_sync(0xffffffffu, team, 0);
    const uint32_t smem_s = smem_u32(dyn_smem);
    const uint32_t blob_s = smem_s + (uint32_t)P.off_blob;
    const uint32_t team_s = smem_s + (uint32_t)P.tc_off_team0 + (uint32_t)team_u * (uint32_t)P.tc_team_stride;
    const uint32_t act_hi_s = team_s + (uint32_t)P.tc_t_act_hi, act_lo_s = team_s + (uint32_t)P.tc_t_act_lo;
    const uint32_t d_col0 = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(team_u * tiles * 128);
    // ticket lock on the tensor pipe: one team's MMAs in flight at a time, so they run at full speed
    // while the other teams are in their CUDA-core phases (alternation instead of lockstep)
    volatile uint32_t* lock = reinterpret_cast<volatile uint32_t*>(smem_raw + P.tc_off_slot + 4);  // [0] next ticket, [1] now serving
    uint32_t req_parity = 0;
    bool holding = false;
    auto post_layer = [&](int layer_idx) {
      if (issuer) {
        mbar_wait_guarded(req, req_parity);          // every thread of the team has stored its planes
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(const_cast<uint32_t*>(lock), 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        long long start = 0;
        uint32_t spins = 0;
        while (lock[1] != ticket) {
          if ((++spins & 1023u) == 0) {
            const long long now = clock64();
            if (start == 0) start = now;
            else if (now - start > kSpinCycles) asm volatile("trap;");
          }
        }
        fence_after();
        if (P.tc_debug & 1) {
          // timing experiment: no MMAs
        } else if (layer_idx != hidden_tc_layers) {
          const uint32_t off = (uint32_t)(P.tc_bhid_off + layer_idx * P.tc_bhid_stride) * 4u;
          for (int m = 0; m < tiles; ++m)
            issue_layer<F16>(act_hi_s, act_lo_s, plane_bytes, blob_s + off, 2u * 32u * 16u, m, d_col0 + (uint32_t)m * 128u, 32);
        } else {
          for (int m = 0; m < tiles; ++m)
            issue_layer<F16>(act_hi_s, act_lo_s, plane_bytes, blob_s + (uint32_t)P.tc_blast_off * 4u,
                        2u * (uint32_t)NL * 16u, m, d_col0 + (uint32_t)m * 128u, NL);
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
        holding = true;
      }
      req_parity ^= 1u;
    };
    // called right after a wait on `done`: the team's MMAs have completed, pass the pipe on
    auto release_pipe = [&]() {
      if (issuer && holding) {
        if (lane == 0) atomicAdd(const_cast<uint32_t*>(lock + 1), 1u);
        holding = false;
      }
    };
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((team * tiles + tile) * 128);
    const Tableau& tab = *tab_s;
    const bool cons = eq_conservative(P.eq);
    const bool forced_eq = eq_forced(P.eq) && P.P > 0;

    const int g = blockIdx.x * R + team;
    for (int row = g; row < W.batch; row += total_teams) {
      const int sample = W.sample_offset + row;
      const ForcingTerm fterm = load_forcing_term(P, sample, x);
      double y = W.u64 ? W.u64[(size_t)row * N + x] : (double)__ldg(W.u + (size_t)row * N + x);
      int first_bad = -1, save_idx = 0;
      const int nsteps = (W.op == OP_INTEGRATE) ? W.nsteps : 1;
      for (int step = 0; step < nsteps; ++step) {
        const double t0 = W.t0 + (double)step * W.dt;
        const int nstages = (W.op == OP_INTEGRATE) ? tab.stages : 1;
        for (int s = 0; s < nstages; ++s) {
          // ---- stage value, rounded to float32 (integrate.py:57-60,71) ----
          double accd = 0.0;
#pragma unroll
          for (int j = 0; j < kMaxStages; ++j)
            if (j < s && tab.a[s][j] != 0.0) accd += tab.a[s][j] * (double)kst[j * N + x];
          const float us = (float)(s == 0 ? y : y + W.dt * accd);
          const float usn = __fdiv_rn(us, P.sigma);            // model.py:450-451
          ust[x + kHalo] = us;
          unr[x + kHalo] = usn;
          if (F16) {      // row maximum of |u / sigma| for the activation bounds; rides on the stage barrier
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(usn)));
            if (lane == 0) umax_w[warp_in_team] = wmax;
          }
          if (edge) {
            if (x < kHalo) { ust[x + kHalo + N] = us; unr[x + kHalo + N] = usn; }
            if (x >= N - kHalo) { ust[x + kHalo - N] = us; unr[x + kHalo - N] = usn; }
          }
          const bool forced = forced_eq && (W.op == OP_RHS || W.op == OP_INTEGRATE);
          if (forced && s == 0) {
            // the sincos of every stage of this step, spread over nstages * P threads
            const int sq = x / P.P, q = x - sq * P.P;
            if (sq < nstages) {
              const ForcingTerm fq = (sq == 0) ? fterm : load_forcing_term(P, sample, q);
              const float ts = (float)(W.op == OP_INTEGRATE ? t0 + tab.c[sq] * W.dt : W.t0);
              forcing_terms(P, fs + sq * kForcingStride, fq, q, ts);
            }
          }
          team_sync(team, N);
          if (forced) forcing_reduce(P, fs + s * kForcingStride, x);   // visible to the team after the mbarrier rounds below
          float s_act = 1.f, bound = 0.f;      // scale of the planes being written / bound on their values
          if (F16) {
            uint32_t m = 0;
            for (int w = 0; w < team_warps; ++w) m = max(m, umax_w[w]);
            bound = fmaf(P.tc_w1abs, __uint_as_float(m), P.tc_b1abs);   // |h1| <= |b1| + sum|W1| * max|u/sigma|
            s_act = scale_for(bound);
          }
          float u7[kWin];
#pragma unroll
          for (int j = 0; j < kWin; ++j) u7[j] = ust[x + j];

          // ---- first layer 1 -> 32 on the CUDA cores, split and written as A planes ----
          {
            float un[kTaps];
#pragma unroll
            for (int k = 0; k < kTaps; ++k) un[k] = unr[x + k + 1];
            float pend[8];      // fp16 planes hold 8 channels per 16-byte chunk
#pragma unroll
            for (int c4 = 0; c4 < kChunks; ++c4) {
              float4 h = make_float4(P.tc_b1[4 * c4], P.tc_b1[4 * c4 + 1], P.tc_b1[4 * c4 + 2], P.tc_b1[4 * c4 + 3]);
#pragma unroll
              for (int k = 0; k < kTaps; ++k) {      // filters are constant-bank operands of the FFMAs
                h.x = fmaf(un[k], P.tc_w1[k * kF + 4 * c4], h.x);
                h.y = fmaf(un[k], P.tc_w1[k * kF + 4 * c4 + 1], h.y);
                h.z = fmaf(un[k], P.tc_w1[k * kF + 4 * c4 + 2], h.z);
                h.w = fmaf(un[k], P.tc_w1[k * kF + 4 * c4 + 3], h.w);
              }
              // hidden activations are ReLU on this engine (other nonlinearities use the FFMA engine)
              if (F16) {
                const int o = (c4 & 1) * 4;
                pend[o] = fmaxf(h.x, 0.f) * s_act; pend[o + 1] = fmaxf(h.y, 0.f) * s_act;
                pend[o + 2] = fmaxf(h.z, 0.f) * s_act; pend[o + 3] = fmaxf(h.w, 0.f) * s_act;
                if (c4 & 1)
                  store_split_f16(act_hi + (size_t)(c4 >> 1) * plane_bytes, act_lo + (size_t)(c4 >> 1) * plane_bytes,
                                  x, N, edge, pend);
              } else {
                store_split(act_hi + (size_t)c4 * plane_bytes, act_lo + (size_t)c4 * plane_bytes, x, N, edge,
                            fmaxf(h.x, 0.f), fmaxf(h.y, 0.f), fmaxf(h.z, 0.f), fmaxf(h.w, 0.f));
              }
            }
          }
          fence_async_smem();
          mbar_arrive(req);
          post_layer(0);

          // ---- hidden layers on the tensor pipe; epilogue rewrites the planes in place ----
          for (int l = 0; l < hidden_tc_layers; ++l) {
            mbar_wait_guarded(done, done_parity);
            done_parity ^= 1u;
            release_pipe();
            fence_after();
            float acc[32];
            tmem_sum32(taddr, acc, F16 ? (1.f / 2048.f) : 1.f);
            fence_before();
            // accumulators carry (activation scale x filter scale); the next planes get their own scale
            const float inv = F16 ? 1.f / (s_act * P.tc_sw_hid) : 1.f;
            if (F16) {
              bound = fmaf(P.tc_whabs, bound, P.tc_bhabs);      // |h2| <= |b2| + max_co sum|W2| * max|h1|
              s_act = scale_for(bound);
            }
            float pend[8];
#pragma unroll
            for (int c4 = 0; c4 < kChunks; ++c4) {
              const float4 b = make_float4(P.tc_bh[4 * c4], P.tc_bh[4 * c4 + 1], P.tc_bh[4 * c4 + 2], P.tc_bh[4 * c4 + 3]);
              if (F16) {
                const int o = (c4 & 1) * 4;
                pend[o] = fmaxf(fmaf(acc[4 * c4], inv, b.x), 0.f) * s_act;
                pend[o + 1] = fmaxf(fmaf(acc[4 * c4 + 1], inv, b.y), 0.f) * s_act;
                pend[o + 2] = fmaxf(fmaf(acc[4 * c4 + 2], inv, b.z), 0.f) * s_act;
                pend[o + 3] = fmaxf(fmaf(acc[4 * c4 + 3], inv, b.w), 0.f) * s_act;
                if (c4 & 1)
                  store_split_f16(act_hi + (size_t)(c4 >> 1) * plane_bytes, act_lo + (size_t)(c4 >> 1) * plane_bytes,
                                  x, N, edge, pend);
              } else {
                store_split(act_hi + (size_t)c4 * plane_bytes, act_lo + (size_t)c4 * plane_bytes, x, N, edge,
                            fmaxf(acc[4 * c4] + b.x, 0.f), fmaxf(acc[4 * c4 + 1] + b.y, 0.f),
                            fmaxf(acc[4 * c4 + 2] + b.z, 0.f), fmaxf(acc[4 * c4 + 3] + b.w, 0.f));
              }
            }
            fence_async_smem();
            mbar_arrive(req);
            post_layer(l + 1);
          }

          // ---- last layer: stencil coefficients (projection folded in) straight from TMEM ----
          mbar_wait_guarded(done, done_parity);
          done_parity ^= 1u;
          release_pipe();
          fence_after();
          float dv[kMaxD];
          const float inv_last = F16 ? 1.f / (s_act * P.tc_sw_last) : 1.f;
          const float cross_scale = F16 ? (1.f / 2048.f) : 1.f;
          if (NL == 16) last_epilogue<16>(P, W, taddr, u7, row, x, dv, cross_scale, inv_last);
          else last_epilogue<32>(P, W, taddr, u7, row, x, dv, cross_scale, inv_last);
          if (W.op == OP_COEF || W.op == OP_DERIV) continue;
          float r = equation_point(P.eq, u7[kHalo], dv, P.eta);
          if (cons) {
            flux[x] = r;
            team_sync(team, N);
            const float fwd = flux[x + 1 == N ? 0 : x + 1];
            r = -__fmul_rn(P.inv_dx, __fsub_rn(fwd, r));
          }
          if (forced) {
            float f = 0.f;
            for (int m = 0; m < P.M; ++m) {
              f = fmaf(fs[s * kForcingStride + m], __ldg(P.fbasis + (size_t)m * N + x), f);
              f = fmaf(fs[s * kForcingStride + P.M + m], __ldg(P.fbasis + (size_t)(P.M + m) * N + x), f);
            }
            r = __fadd_rn(r, f);
          }
          if (W.op == OP_RHS) {
            if (W.out64) W.out64[(size_t)row * N + x] = (double)r;
            else W.out[(size_t)row * N + x] = r;
          } else {
            kst[s * N + x] = r;
          }
        }
        if (W.op != OP_INTEGRATE) continue;
        double accd = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxStages; ++j)
          if (j < tab.stages && tab.b[j] != 0.0) accd += tab.b[j] * (double)kst[j * N + x];
        y = y + W.dt * accd;
        if (first_bad < 0 && !isfinite(y)) first_bad = step;
        if (((step + 1) % W.save_every) == 0) {
          W.snaps[((size_t)save_idx * W.batch + row) * N + x] = (float)y;
          ++save_idx;
        }
      }
      if (W.op == OP_INTEGRATE && W.first_bad) {
        unsigned int* slot = reinterpret_cast<unsigned int*>(fs + kMaxStages * kForcingStride);
        if (x == 0) *slot = 0xffffffffu;
        team_sync(team, N);
        atomicMin(slot, first_bad < 0 ? 0xffffffffu : (unsigned int)first_bad);
        team_sync(team, N);
        if (x == 0) W.first_bad[row] = (*slot == 0xffffffffu) ? -1 : (int)*slot;
        team_sync(team, N);
      }
    }
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (is_alloc_warp) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Probe: one 128-position tile of a 32 -> NOUT, 5-tap periodic-free conv through the same
// descriptor / split / TMEM path.  Used by tests to validate layouts in isolation.
//   x     [132][32] float  (positions -2..129)
//   w_cat packed B planes [5*8][2*NOUT][4]: rows 0..NOUT-1 = Whi, NOUT..2*NOUT-1 = Wlo
//   out   [128][NOUT] float
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) tc_probe_kernel(const float* __restrict__ xin,
                                                          const float* __restrict__ w_cat, float* __restrict__ out,
                                                          int nout) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t plane_bytes = 132u * 16u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  unsigned char* a_hi = smem_raw + 128;
  unsigned char* a_lo = a_hi + kChunks * plane_bytes;
  float* b_cat = reinterpret_cast<float*>(a_lo + kChunks * plane_bytes);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(slot, 128);
  for (int i = tid; i < 132 * kChunks; i += blockDim.x) {
    const int pos = i / kChunks, c4 = i % kChunks;
    const float4 v = *reinterpret_cast<const float4*>(xin + (size_t)pos * kF + 4 * c4);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(a_hi + (size_t)c4 * plane_bytes + (size_t)pos * 16) = h;
    *reinterpret_cast<float4*>(a_lo + (size_t)c4 * plane_bytes + (size_t)pos * 16) = l;
  }
  for (int i = tid; i < kTaps * kChunks * 2 * nout * 4; i += blockDim.x) b_cat[i] = w_cat[i];
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  if (warp == 4) {
    const uint32_t base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    issue_layer<false>(smem_u32(a_hi), smem_u32(a_lo), plane_bytes, smem_u32(b_cat), 2u * (uint32_t)nout * 16u, 0, base_u,
                nout);
    if (elect_one()) mma_commit(bar);
    __syncwarp();
  } else {
    mbar_wait_guarded(bar, 0);
    fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (nout == 16) {
      float v[16];
      tmem_sum16(taddr, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) out[(size_t)tid * 16 + i] = v[i];
    } else {
      float v[32];
      tmem_sum32(taddr, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) out[(size_t)tid * 32 + i] = v[i];
    }
    fence_before();
  }
  __syncthreads();
  fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark (debug): one warp issues reps x 20 (tap, ci-block) steps on planes of
// arbitrary data; the CTA measures the clocks until tcgen05.commit fires.  Per step up to two MMAs:
//   first : A = plane set 0, N = n1, D columns at d_off1 (+ 128 * (k & 1) when alt != 0)
//   second: A = plane set `a2`, N = n2, D columns at d_off2 (same alternation)
// n == 0 skips that MMA.  kind 0 = tf32 (K = 8, 4-byte elements), 1 = bf16 (kind::f16, K = 16).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t instr_desc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int KIND, int N1, int N2, int BROWS, int ALT>
__global__ void __launch_bounds__(128, 1) tc_rate_kernel(int reps, long long* __restrict__ cycles) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t plane_bytes = 516u * 16u;
  constexpr uint32_t b_plane_bytes = (uint32_t)BROWS * 16u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  unsigned char* a_0 = smem_raw + 128;
  unsigned char* a_1 = a_0 + kChunks * plane_bytes;
  unsigned char* b_cat = a_1 + kChunks * plane_bytes;
  for (uint32_t i = tid; i < (2 * kChunks * plane_bytes + kTaps * kChunks * b_plane_bytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(a_0)[i] = 0x3c003c00u + (i & 63u);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  long long t0 = 0;
  if (warp == 0) {
    const uint32_t base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t plane16 = plane_bytes >> 4, bplane16 = b_plane_bytes >> 4;
    const uint32_t a00 = ((smem_u32(a_0) >> 4) & 0x3FFFu) | (plane16 << 16);
    const uint32_t a10 = ((smem_u32(a_1) >> 4) & 0x3FFFu) | (plane16 << 16);
    const uint32_t b0 = ((smem_u32(b_cat) >> 4) & 0x3FFFu) | (bplane16 << 16);
    const uint32_t id1 = KIND ? instr_desc_bf16(128, N1 ? N1 : 16) : instr_desc_tf32(128, N1 ? N1 : 16);
    const uint32_t id2 = KIND ? instr_desc_bf16(128, N2 ? N2 : 16) : instr_desc_tf32(128, N2 ? N2 : 16);
    t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
#pragma unroll
          for (int kb = 0; kb < kChunks / 2; ++kb) {
            const uint32_t dsel = ALT ? (uint32_t)(k & 1) * 256u : 0u;
            const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
            const uint32_t bo = (uint32_t)(k * kChunks + 2 * kb) * bplane16;
            if (N1) {
              if (KIND) mma_bf16_split(base_u + dsel, a00 + ao, b0 + bo, desc_hi, id1, 1u);
              else mma_tf32_split(base_u + dsel, a00 + ao, b0 + bo, desc_hi, id1, 1u);
            }
            if (N2) {
              if (KIND) mma_bf16_split(base_u + dsel + 128u, a10 + ao, b0 + bo, desc_hi, id2, 1u);
              else mma_tf32_split(base_u + dsel + 128u, a10 + ao, b0 + bo, desc_hi, id2, 1u);
            }
          }
        }
      }
    }
    __syncwarp();
    if (elect_one()) mma_commit(bar);
    __syncwarp();
    mbar_wait_guarded(bar, 0);
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace ddd1d
