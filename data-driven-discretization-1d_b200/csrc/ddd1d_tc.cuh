// Tensor-core (tcgen05 / TMEM) version of the learned-coefficient row kernel, sm_100a.
//
// The conv stack is 96 % of the FLOPs and is an implicit GEMM per 128-position tile:
//   hidden layer : D[128 x 32] += sum_{tap k, ci-block} A_k[128 x 16] * B_k[16 x 32]   (K = 5*32)
//   last layer   : D[128 x NL] += ...                                                   (NL = 16 | 32)
// A = activations kept in shared memory as K-major "chunk planes" [ci/8][position][8 halfs] (no swizzle), so
// the tap shift k is just +16 B on the descriptor start address and the periodic halo is two extra positions
// per plane.  B = filters pre-packed on the host in the same canonical layout.  FP32 fidelity on an fp16 pipe
// comes from a two-term split: x * s = hi + lo (both fp16, s a power of two from a verified bound on the
// row) and hi*Wh + lo*Wh + hi*Wl accumulated in FP32 in TMEM (the dropped lo*Wl term is 2^-22 relative).
// B holds [Wh | Wl] side by side, so one MMA of width 2*NB produces hi*Wh (the "main" columns) and hi*Wl (the
// "cross" columns) and a second one of width NB adds lo*Wh to the cross columns: two instructions per
// (tap, ci-block) instead of three.  The tensor core truncates when it adds into an accumulator (measured,
// profiles/r01/tc_precision.txt), so the small cross terms keep their own columns and the epilogue adds
// main + cross in FP32.  The polynomial-accuracy projection is folded into the last layer's filters on the
// host (W3' = W3 . nullspace, window form), so the last epilogue reads stencil coefficients straight out of
// TMEM.  (issue_layer<false, ...> is the TF32 / 3xTF32 form of the same scheme; the probe kernel and the
// precision microbenchmark use it.)
//
// Warp roles (one CTA per SM, persistent): R "row teams" of N threads (thread <-> grid point; the team's
// warps are 4-aligned so each warp reads its own TMEM lane quadrant) run the whole Runge-Kutta program of two
// rows each ("slots"), and P.tc_issuers further warps do nothing but issue tcgen05.mma for the slots they
// serve and signal completion with tcgen05.commit -> mbarrier.  While one row's MMAs run, its team works on
// its other row.  See the comment above tc_row_kernel and DESIGN.md section 4.1 for what bounds the kernel.
#pragma once
#include <cuda_fp16.h>

#include "ddd1d_device.cuh"

namespace ddd1d {
namespace tc {

constexpr int kF = 32;             // hidden width this path is built for
constexpr int kTaps = 5;
constexpr int kChunks = kF / 4;    // 16-byte chunks along ci
constexpr int kFsStride = 2 * kMaxModes;     // forcing mode amplitudes per RK stage (sine | cosine)
constexpr int kFsBuffers = 3;                // amplitude sets in rotation: written one step ahead of their use
constexpr int kFsWords = kFsBuffers * kMaxStages * kFsStride;     // + 1 word: first non-finite step of the row
constexpr int kTraceCap = 8192;     // debug trace records per stream
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // suspends up to the hint (ns)
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > kSpinCycles) asm volatile("trap;");
    }
  }
}
// debug event trace: (tag << 48 | clock64) records, one stream per traced thread (null = off)
#ifdef DDD1D_TRACE
__device__ __forceinline__ void trace_ev(long long* base, int& n, int tag) {
  if (base && n < kTraceCap) base[n++] = ((long long)tag << 48) | (clock64() & 0xffffffffffffll);
}
#else
#define trace_ev(base, n, tag) ((void)0)
#endif
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void team_sync(int team, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(threads) : "memory");
}
// the same barrier carrying a vote: true iff `ok` holds on every thread of the team
__device__ __forceinline__ bool team_sync_all(int team, int threads, bool ok) {
  uint32_t all;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "barrier.cta.red.and.pred q, %1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(all)
      : "r"(team + 1), "r"(threads), "r"((uint32_t)ok)
      : "memory");
  return all != 0;
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
// EO = true : taps accumulate alternately into two D blocks [even main|cross][odd main|cross] (halves the
//             accumulate chain; the TF32 probe uses it);  EO = false: one block [main | cross].
template <bool F16, bool EO>
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
      const uint32_t d_main = d_col + (EO ? (uint32_t)((k & 1) * 2 * nb) : 0u);
      const uint32_t d_cross = d_main + (uint32_t)nb;
#pragma unroll
      for (int kb = 0; kb < kPlanes / 2; ++kb) {
        const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
        const uint32_t bo = (uint32_t)(k * kPlanes + 2 * kb) * bplane16;
        const uint32_t first = (k < (EO ? 2 : 1) && kb == 0) ? 0u : 1u;     // first touch of a D block
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
// v (already multiplied by the layer's power-of-two scale) = hi + lo with hi = fp16(v) and
// lo = fp16(v - hi): 22 significant bits like the 3xTF32 split, but 2 bytes per element, so one 4 KB A read
// covers K = 16.  Static bounds on the activations (operator norms x a verified bound on the row's
// max |u/sigma|) put the largest v in [2^12, 2^14), far from fp16's range limits; scales are powers of two,
// i.e. exact.  lo is at most half an ulp of hi; where it falls into fp16's subnormal range (|v| < 2^-3) its
// absolute error 2^-25 is 2^-37 of the row's bound.  kLoScale = 2048 would keep lo normal everywhere at the
// price of two more multiplies per pair (the filters' lo rows carry the same factor: ddd1d_api.cu).
constexpr float kLoScale = 1.f;
__device__ __forceinline__ void split_half2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);                  // one packed conversion
  const float2 f = __half22float2(h);
  const __half2 l = kLoScale == 1.f ? __floats2half2_rn(a - f.x, b - f.y)     // exact remainder
                                    : __floats2half2_rn((a - f.x) * kLoScale, (b - f.y) * kLoScale);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void store_split_f16(unsigned char* hi_plane, unsigned char* lo_plane, int x, int N,
                                                bool edge, const float (&v)[8]) {
  uint4 h, l;
  split_half2(v[0], v[1], h.x, l.x);
  split_half2(v[2], v[3], h.y, l.y);
  split_half2(v[4], v[5], h.z, l.z);
  split_half2(v[6], v[7], h.w, l.w);
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
  // bound in [2^(eb-127), 2^(eb-126))  ->  s = 2^(140 - eb), i.e. bound * s in [2^13, 2^14)
  const int eb = (int)((__float_as_uint(fmaxf(bound, 1e-30f)) >> 23) & 0xffu);
  return __uint_as_float((uint32_t)(min(140 - eb, 60) + 127) << 23);
}
// 1 / s for a power of two s (exact)
__device__ __forceinline__ float pow2_inverse(float s) {
  return __uint_as_float((254u << 23) - __float_as_uint(s));
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
// sixteen outputs = main + cross * cross_scale
__device__ __forceinline__ void tmem_pair16(uint32_t t_main, uint32_t t_cross, float* v, float cross_scale) {
  uint32_t a[16], b[16];
  tmem_ld16_issue(t_main, a);
  tmem_ld16_issue(t_cross, b);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(b[i]), cross_scale, __uint_as_float(a[i]));
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
  tmem_pair16(taddr, taddr + NLV, cfv, cross_scale);                       // [main NLV | cross NLV]
  if (NLV == 32) tmem_pair16(taddr + 16, taddr + NLV + 16, cfv + 16, cross_scale);
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
//
// A CTA holds R row teams of N threads (thread <-> grid point); every team keeps SLOTS rows in flight
// and walks them round-robin through the three phases of a right-hand-side evaluation:
//   phase 0  stage value -> first layer on the CUDA cores -> fp16 planes -> issue the hidden layer's MMAs
//   phase 1  (per hidden layer) TMEM -> bias/ReLU -> fp16 planes -> issue the next layer's MMAs
//   phase 2  TMEM -> coefficients -> stencil dot products -> equation -> stage derivative
// While slot A's MMAs run on the tensor pipe the team's threads are in slot B's CUDA-core phase, so the
// mbarrier wait at the top of phases 1 and 2 normally returns at once.
//
// A tcgen05.mma issue blocks the issuing thread once the tensor pipe's queue is full, i.e. for about as long
// as the MMAs take.  A warp that both computes and issues therefore serialises its CUDA-core work with every
// burst, and its team waits for it at the next barrier.  So the MMAs are issued by one extra warp that does
// nothing else: it polls the slots' "planes stored" barriers and issues whichever layer is ready.
//
// Shared-memory LOADS stall for as long as MMAs stream their operands from shared memory (stores, shuffles,
// tcgen05.ld, global loads and barriers do not: scripts/tc_overlap.py).  So the steady-state loop issues no
// LDS at all: shared memory holds only what the tensor pipe reads (activation planes, filter planes);
// everything threads exchange among themselves (stage row + halo, row maxima, flux, forcing amplitudes)
// goes through a small per-CTA global scratch that stays in L1/L2, per-thread state that outlives a phase
// (float64 solution, stage derivatives, bounds) sits in slot-indexed local arrays, and the Runge-Kutta
// tableau is a kernel parameter (constant bank).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSlots = 2;
// MMA issuer warps.  One warp's instruction stream (descriptor arithmetic, uniform-register moves, the
// tcgen05.mma itself) sustains one MMA per ~62 clk, the tensor pipe takes one per ~42 clk at these shapes: the
// single issuer, not the pipe, bounded the MMA stream (measured with the teams switched off: 10.5 ms for the
// C2 launch against 7.1 ms of MMA time).  Two issuers on different scheduler partitions each serve half of
// the slots; their MMAs interleave in the pipe, every tcgen05.commit tracks its own warp's MMAs.
constexpr int kMaxIssuers = 4;     // P.tc_issuers of them are launched (2 by default, 4 for the many short layers of N = 128)
__global__ void __launch_bounds__(512 + 32 * kMaxIssuers, 1) tc_row_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W,
                                                        const __grid_constant__ Tableau tab) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P.N, R = P.tc_teams, SLOTS = P.tc_slots, tiles = N / 128;
  const int team_warps = N / 32;
  const bool is_alloc_warp = warp == 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + P.tc_off_slot);
  float* blob = reinterpret_cast<float*>(smem_raw + P.off_blob);
  const uint32_t plane_bytes = (uint32_t)(N + 4) * 16u;
  const int NL = P.tc_nlast;                         // 16 or 32 columns for the last layer
  const int hidden_tc_layers = P.nlayers - 2;        // layers between the first and the last
  const int TS = R * SLOTS;                          // rows in flight per CTA

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    for (int t = 0; t < TS; ++t) mbar_init(&bars[1 + t], (uint32_t)N);      // "planes stored": every thread of the team arrives
    for (int t = 0; t < TS * tiles; ++t) mbar_init(&bars[1 + TS + t], 1);    // "tile's MMAs done": tcgen05.commit arrives
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)P.blob_floats * 4u;
    mbar_expect_tx(&bars[0], bytes);
    bulk_copy_g2s(blob, P.blob, bytes, &bars[0]);
  }
  if (is_alloc_warp) tmem_alloc(tmem_slot, 512);      // TS rows x tiles x 64 columns (main | cross)
  mbar_wait_guarded(&bars[0], 0);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_teams = gridDim.x * R;
  const int nsteps = (W.op == OP_INTEGRATE) ? W.nsteps : 1;
  const int nstages = (W.op == OP_INTEGRATE) ? tab.stages : 1;
  const uint32_t smem_s = smem_u32(dyn_smem);

  if (warp >= R * team_warps) {
    // ---------------- issuer warp ----------------
    // The tensor pipe's queue is only a few MMAs deep, so the pipe drains (and pays its ~1000 clk start-up
    // latency again) unless the next layer's first MMA is queued within ~150 clk of the previous layer's last.
    // Hence the polling is lean: lane l owns slot l's bookkeeping in registers and polls its "planes stored"
    // barrier, one ballot finds the ready slots, and the oldest-served-first rotation picks one.
    const uint32_t blob_s = smem_s + (uint32_t)P.off_blob;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int my_remaining = 0, my_layer = 0;
    uint32_t my_parity = 0;
    const int issuer = warp - R * team_warps;          // serves `served` consecutive slots
    const int served = (TS + P.tc_issuers - 1) / P.tc_issuers;
    if (lane >= issuer * served && lane < (issuer + 1) * served && lane < TS) {
      const int first = blockIdx.x * R + lane / SLOTS + (lane % SLOTS) * total_teams;      // this slot's first row
      const int stride = SLOTS * total_teams;
      const int rows = first < W.batch ? (W.batch - first + stride - 1) / stride : 0;
      my_remaining = rows * nsteps * nstages * (hidden_tc_layers + 1);
    }
    const uint32_t b_hid_s = blob_s + (uint32_t)P.tc_bhid_off * 4u, b_last_s = blob_s + (uint32_t)P.tc_blast_off * 4u;
    const uint32_t team0_s = smem_s + (uint32_t)P.tc_off_team0;
    long long start = 0;
    uint32_t idle = 0;
    int next = 0;                                     // rotation: the slot after the one served last goes first
#ifdef DDD1D_TRACE
    long long* const tr = (P.tc_trace && blockIdx.x == 0 && lane == 0) ? P.tc_trace + 4 * kTraceCap : nullptr;
    int trn = 0;
#endif
    while (true) {
      // (debug bit 6: the MMA stream alone -- every request counts as made, the teams sit the launch out)
      const bool ready = my_remaining > 0 && ((P.tc_debug & 64) || mbar_test(&bars[1 + lane], my_parity));    // (acquire)
      const uint32_t mask = __ballot_sync(0xffffffffu, ready);
      if (mask == 0u) {
        if (__ballot_sync(0xffffffffu, my_remaining > 0) == 0u) break;
        if ((++idle & 0xfffu) == 0) {
          const long long now = clock64();
          if (start == 0) start = now;
          else if (now - start > kSpinCycles) asm volatile("trap;");
        }
        if (P.tc_debug & 4) __nanosleep(32);
        continue;
      }
      idle = 0;
      start = 0;
      const uint32_t rot = (mask >> next) | (mask << ((32 - next) & 31));       // bit i = slot (next + i) % 32
      const int ts = (next + __ffs((int)rot) - 1) & 31;
      next = ts + 1 == TS ? 0 : ts + 1;
      const int layer_idx = __shfl_sync(0xffffffffu, my_layer, ts);
      trace_ev(tr, trn, ts * 16 + 1);
      fence_after();
      const uint32_t slot_s = team0_s + (uint32_t)ts * (uint32_t)P.tc_team_stride;
      const uint32_t act_hi_s = slot_s + (uint32_t)P.tc_t_act_hi, act_lo_s = slot_s + (uint32_t)P.tc_t_act_lo;
      const uint32_t d_col0 = tmem_u + (uint32_t)(ts * tiles * 64);
      const bool last = layer_idx == hidden_tc_layers;
      if (!(P.tc_debug & 1)) {          // (bit 0: timing experiment without MMAs)
        const uint32_t b_s = last ? b_last_s : b_hid_s + (uint32_t)(layer_idx * P.tc_bhid_stride) * 4u;
        const int nb = last ? NL : 32;
        for (int m = 0; m < tiles; ++m) {       // one commit per tile: its threads start their epilogue early
          issue_layer<true, false>(act_hi_s, act_lo_s, plane_bytes, b_s, 2u * (uint32_t)nb * 16u, m,
                                   d_col0 + (uint32_t)m * 64u, nb);
          if (elect_one()) mma_commit(&bars[1 + TS + ts * tiles + m]);
          __syncwarp();
        }
      } else {
        for (int m = 0; m < tiles; ++m)
          if (elect_one()) mma_commit(&bars[1 + TS + ts * tiles + m]);
        __syncwarp();
      }
      trace_ev(tr, trn, ts * 16 + 2);
      if (lane == ts) {
        my_parity ^= 1u;
        my_layer = last ? 0 : layer_idx + 1;
        my_remaining -= 1;
      }
    }
  } else {
  // ---------------- row teams ----------------
  const int team = warp / team_warps;
  const int x = tid - team * N;                      // this thread's grid point
  const int tile = x >> 7;
  const int warp_in_team = __shfl_sync(0xffffffffu, warp - team * team_warps, 0);
  const bool edge = warp_in_team == 0 || warp_in_team == team_warps - 1;
  const bool cons = eq_conservative(P.eq);
  const bool forced_eq = eq_forced(P.eq) && P.P > 0;
  const bool forced = forced_eq && (W.op == OP_RHS || W.op == OP_INTEGRATE);
  const float cross_scale = 1.f / kLoScale;
  uint32_t done_parity = 0;
#ifdef DDD1D_TRACE
  long long* const tr = (P.tc_trace && blockIdx.x == 0 && (x == 0 || x == N - 1))
                            ? P.tc_trace + (team + (x == 0 ? 0 : 2)) * kTraceCap : nullptr;
  int trn = 0;
#endif

  // ---- per-slot views -------------------------------------------------------------------------
  struct SlotView {
    unsigned char *act_hi, *act_lo;      // shared memory (written with STS, read by the tensor pipe)
    float *ust, *unr, *flux, *fs;        // global scratch
    uint32_t* umax_w;
    uint64_t *req, *done, *done_nb;
    uint32_t taddr;
  };
  // MMAs complete tile by tile.  A tile's MMAs also read two positions of each neighbouring tile (and the
  // periodic halo copies), so the first / last warp of a tile must see the neighbouring tile's MMAs complete
  // as well before it overwrites its planes; the inner warps only depend on their own tile.
  const int warp_in_tile = warp_in_team & 3;
  const int nb_tile = tiles == 1 ? -1
                      : warp_in_tile == 0 ? (tile + tiles - 1) % tiles
                      : warp_in_tile == 3 ? (tile + 1) % tiles : -1;
  // The stage row (raw, normalised) and its warp maxima are double-buffered on the stage parity: phase 2
  // of stage s reads them while a faster warp may already be writing stage s+1 (phase 0).
  uint32_t stage_par = 0;
  auto view = [&](int sl) {
    const int ts = team * SLOTS + sl;
    unsigned char* tb = smem_raw + P.tc_off_team0 + (size_t)ts * P.tc_team_stride;
    SlotView v;
    v.act_hi = tb + P.tc_t_act_hi;
    v.act_lo = tb + P.tc_t_act_lo;
    float* sc = P.tc_scratch + ((size_t)blockIdx.x * TS + ts) * P.tc_sc_stride;
    v.ust = sc + stage_par * (uint32_t)(2 * (N + 2 * kHalo + 2));
    v.unr = v.ust + (N + 2 * kHalo + 2);             // the same row divided by sigma
    v.umax_w = reinterpret_cast<uint32_t*>(sc + P.tc_sc_umax);
    v.flux = sc + P.tc_sc_flux;
    v.fs = sc + P.tc_sc_fs;
    v.req = &bars[1 + ts];
    v.done = &bars[1 + TS + ts * tiles + tile];
    v.done_nb = nb_tile >= 0 ? &bars[1 + TS + ts * tiles + nb_tile] : nullptr;
    v.taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((ts * tiles + tile) * 64);
    return v;
  };
  // The fp16 plane scales rest on a bound `umax_s` on the row's max |u / sigma|.  It is fixed when a row
  // starts (twice the row's maximum) and every stage only VERIFIES it, as a vote carried by the barrier the
  // stage needs anyway; the slow path -- row maximum by atomics performed in L2, read past L1 -- runs for a
  // row's first stage and again if a row ever outgrows its bound.  (Reading the maximum every stage put an
  // L2 round trip on the critical path of every right-hand side.)
  auto recalibrate = [&](const SlotView& v, float usn) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(usn)));
    if (lane == 0) atomicMax(v.umax_w, wmax);
    team_sync(team, N);
    const float m = __uint_as_float(__ldcg(v.umax_w));
    team_sync(team, N);
    if (x == 0) *v.umax_w = 0u;                       // next use is at least one barrier away
    return 2.f * m;
  };

  // per-thread state that outlives a phase, indexed by slot (local memory: L1, never LDS)
  double y_s[kMaxSlots];
  float k_s[kMaxSlots][kMaxStages];
  float bound_s[kMaxSlots];
  float umax_s[kMaxSlots];

  const int g = blockIdx.x * R + team;
  for (int row0 = g; row0 < W.batch && !(P.tc_debug & 64); row0 += SLOTS * total_teams) {
    int nslots = 0;
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int row = row0 + sl * total_teams;
      if (row >= W.batch) break;
      ++nslots;
      const SlotView v = view(sl);
      y_s[sl] = W.u64 ? W.u64[(size_t)row * N + x] : (double)__ldg(W.u + (size_t)row * N + x);
      if (x == 0) {
        *reinterpret_cast<unsigned int*>(v.fs + kFsWords) = 0xffffffffu;
        *v.umax_w = 0u;
      }
      umax_s[sl] = -1.f;                             // no bound yet: the first stage calibrates
    }
    team_sync(team, N);
    int save_idx = 0;
    for (int step = 0; step < nsteps; ++step) {
      const double t0 = W.t0 + (double)step * W.dt;
      for (int s = 0; s < nstages; ++s) {
        // ================= phase 0 =================
#pragma unroll 1
        for (int sl = 0; sl < nslots; ++sl) {
          const SlotView v = view(sl);
          trace_ev(tr, trn, sl * 16 + 1);
          const int sample = W.sample_offset + row0 + sl * total_teams;
          // ---- stage value, rounded to float32 (integrate.py:57-60,71) ----
          double accd = 0.0;
#pragma unroll
          for (int j = 0; j < kMaxStages; ++j)
            if (j < s && tab.a[s][j] != 0.0) accd += tab.a[s][j] * (double)k_s[sl][j];
          const double y = y_s[sl];
          const float us = (float)(s == 0 ? y : y + W.dt * accd);
          const float usn = __fdiv_rn(us, P.sigma);            // model.py:450-451
          v.ust[x + kHalo] = us;
          v.unr[x + kHalo] = usn;
          if (edge) {
            if (x < kHalo) { v.ust[x + kHalo + N] = us; v.unr[x + kHalo + N] = usn; }
            if (x >= N - kHalo) { v.ust[x + kHalo - N] = us; v.unr[x + kHalo - N] = usn; }
          }
          if (forced && s == 0 && step == 0 && warp_in_team < nstages) {
            // the first step's amplitudes; later steps get theirs one step ahead, after the planes are stored
            const int sq = warp_in_team;
            const float ts = (float)(W.op == OP_INTEGRATE ? t0 + tab.c[sq] * W.dt : W.t0);
            forcing_amplitudes(P, v.fs + sq * kFsStride, sample, ts, lane);
          }
          trace_ev(tr, trn, sl * 16 + 12);
          float umax = umax_s[sl];
          bool keep = team_sync_all(team, N, fabsf(usn) <= umax);    // team-uniform; NaN rows fail every stage
          if (keep && s == 0 && (step & 15) == 15)                   // now and then: has the row decayed far below
            keep = !team_sync_all(team, N, fabsf(usn) < umax * (1.f / 256.f));   // its bound (lo planes would thin out)?
          if (!keep) {
            umax = recalibrate(v, usn);
            umax_s[sl] = umax;
          }
          trace_ev(tr, trn, sl * 16 + 2);
          const float bound1 = fmaf(P.tc_w1abs, umax, P.tc_b1abs);   // |h1| <= |b1| + sum|W1| * max|u/sigma|
          bound_s[sl] = bound1;
          trace_ev(tr, trn, sl * 16 + 13);
          const float s_act = scale_for(bound1);

          // ---- first layer 1 -> 32 on the CUDA cores, split and written as A planes ----
          float un[kTaps];
#pragma unroll
          for (int k = 0; k < kTaps; ++k) un[k] = v.unr[x + k + 1];
#pragma unroll
          for (int c8 = 0; c8 < kChunks / 2; ++c8) {
            float h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = P.tc_b1[8 * c8 + i];
#pragma unroll
            for (int k = 0; k < kTaps; ++k)        // filters are constant-bank operands of the FFMAs
#pragma unroll
              for (int i = 0; i < 8; ++i) h[i] = fmaf(un[k], P.tc_w1[k * kF + 8 * c8 + i], h[i]);
            // hidden activations are ReLU on this engine (other nonlinearities use the FFMA engine)
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = fmaxf(h[i], 0.f) * s_act;
            store_split_f16(v.act_hi + (size_t)c8 * plane_bytes, v.act_lo + (size_t)c8 * plane_bytes, x, N, edge, h);
          }
          fence_async_smem();
          mbar_arrive(v.req);
          trace_ev(tr, trn, sl * 16 + 3);
          if (forced && s == 0 && W.op == OP_INTEGRATE && step + 1 < nsteps && warp_in_team < nstages) {
            // Forcing amplitudes of the NEXT step, off the critical path (the MMAs just requested are running).
            // Warp sq prepares stage sq: one forcing term per lane, mode amplitudes by warp sums.  Three sets
            // rotate: set (step + 1) % 3 was last read in step - 2, and every warp that gets here has passed a
            // barrier of step `step`, which no warp reaches before it has finished step - 1.
            const int sq = warp_in_team;
            const float ts = (float)(W.t0 + (double)(step + 1) * W.dt + tab.c[sq] * W.dt);
            forcing_amplitudes(P, v.fs + (((step + 1) % kFsBuffers) * kMaxStages + sq) * kFsStride, sample, ts, lane);
          }
        }

        // ================= phase 1: hidden layers on the tensor pipe =================
        for (int l = 0; l < hidden_tc_layers; ++l) {
#pragma unroll 1
          for (int sl = 0; sl < nslots; ++sl) {
            const SlotView v = view(sl);
            trace_ev(tr, trn, sl * 16 + 4);
            mbar_wait_guarded(v.done, done_parity);
            trace_ev(tr, trn, sl * 16 + 5);
            fence_after();
            float acc[32];
            tmem_pair16(v.taddr, v.taddr + 32, acc, cross_scale);
            tmem_pair16(v.taddr + 16, v.taddr + 48, acc + 16, cross_scale);
            fence_before();
            // accumulators carry (activation scale x filter scale); the next planes get their own scale
            const float bound1 = bound_s[sl];
            const float inv = pow2_inverse(scale_for(bound1)) * P.tc_inv_sw_hid;
            const float s_act = scale_for(fmaf(P.tc_whabs, bound1, P.tc_bhabs));   // |h2| <= |b2| + sum|W2| max|h1|
            if (v.done_nb) mbar_wait_guarded(v.done_nb, done_parity);
            trace_ev(tr, trn, sl * 16 + 6);
#pragma unroll
            for (int c8 = 0; c8 < kChunks / 2; ++c8) {
              float h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                h[i] = fmaxf(fmaf(acc[8 * c8 + i], inv, P.tc_bh[8 * c8 + i]), 0.f) * s_act;
              store_split_f16(v.act_hi + (size_t)c8 * plane_bytes, v.act_lo + (size_t)c8 * plane_bytes, x, N, edge, h);
            }
            fence_async_smem();
            mbar_arrive(v.req);
            trace_ev(tr, trn, sl * 16 + 7);
          }
          done_parity ^= 1u;
        }

        // ================= phase 2: coefficients, derivatives, equation =================
#pragma unroll 1
        for (int sl = 0; sl < nslots; ++sl) {
          const SlotView v = view(sl);
          const int row = row0 + sl * total_teams;
          float u7[kWin];
#pragma unroll
          for (int j = 0; j < kWin; ++j) u7[j] = v.ust[x + j];
          const float bound1 = bound_s[sl];
          const float s_last = hidden_tc_layers > 0 ? scale_for(fmaf(P.tc_whabs, bound1, P.tc_bhabs)) : scale_for(bound1);
          const float inv_last = pow2_inverse(s_last) * P.tc_inv_sw_last;
          trace_ev(tr, trn, sl * 16 + 8);
          mbar_wait_guarded(v.done, done_parity);
          trace_ev(tr, trn, sl * 16 + 9);
          if (v.done_nb) mbar_wait_guarded(v.done_nb, done_parity);     // before the next stage rewrites the planes
          trace_ev(tr, trn, sl * 16 + 10);
          fence_after();
          float dv[kMaxD];
          if (NL == 16) last_epilogue<16>(P, W, v.taddr, u7, row, x, dv, cross_scale, inv_last);
          else last_epilogue<32>(P, W, v.taddr, u7, row, x, dv, cross_scale, inv_last);
          if (W.op == OP_COEF || W.op == OP_DERIV) continue;
          trace_ev(tr, trn, sl * 16 + 14);
          float r = equation_point(P.eq, u7[kHalo], dv, P.eta);
          if (cons) {
            v.flux[x] = r;
            team_sync(team, N);
            const float fwd = v.flux[x + 1 == N ? 0 : x + 1];
            r = -__fmul_rn(P.inv_dx, __fsub_rn(fwd, r));
          }
          if (forced) {
            // all loads issued together (unrolled, predicated): one L1 latency instead of 2M in a chain
            const float* amp = v.fs + ((step % kFsBuffers) * kMaxStages + s) * kFsStride;
            const float* basis = P.fbasis + x;
            float f = 0.f;
#pragma unroll
            for (int m = 0; m < kMaxModes; ++m)
              if (m < P.M) {
                f = fmaf(amp[m], __ldg(basis + (size_t)m * N), f);
                f = fmaf(amp[P.M + m], __ldg(basis + (size_t)(P.M + m) * N), f);
              }
            r = __fadd_rn(r, f);
          }
          if (W.op == OP_RHS) {
            if (W.out64) W.out64[(size_t)row * N + x] = (double)r;
            else W.out[(size_t)row * N + x] = r;
          } else {
            k_s[sl][s] = r;
          }
          trace_ev(tr, trn, sl * 16 + 11);
        }
        done_parity ^= 1u;
        stage_par ^= 1u;
      }
      if (W.op != OP_INTEGRATE) continue;
      const bool save = ((step + 1) % W.save_every) == 0;
      for (int sl = 0; sl < nslots; ++sl) {
        const SlotView v = view(sl);
        const int row = row0 + sl * total_teams;
        double accd = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxStages; ++j)
          if (j < tab.stages && tab.b[j] != 0.0) accd += tab.b[j] * (double)k_s[sl][j];
        const double y = y_s[sl] + W.dt * accd;
        y_s[sl] = y;
        if (!isfinite(y))       // first step at which the row left the finite range (rare, so an atomic is fine)
          atomicMin(reinterpret_cast<unsigned int*>(v.fs + kFsWords), (unsigned int)step);
        if (save) W.snaps[((size_t)save_idx * W.batch + row) * N + x] = (float)y;
      }
      if (save) ++save_idx;
    }
    if (W.op == OP_INTEGRATE && W.first_bad) {
      team_sync(team, N);
      for (int sl = 0; sl < nslots; ++sl) {
        const SlotView v = view(sl);
        const unsigned int fb = *reinterpret_cast<unsigned int*>(v.fs + kFsWords);
        if (x == 0) W.first_bad[row0 + sl * total_teams] = (fb == 0xffffffffu) ? -1 : (int)fb;
      }
    }
    team_sync(team, N);       // the next rows reuse the slot regions
  }
  }   // row teams
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
    issue_layer<false, true>(smem_u32(a_hi), smem_u32(a_lo), plane_bytes, smem_u32(b_cat), 2u * (uint32_t)nout * 16u, 0, base_u,
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

template <int KIND, int N1, int N2, int BROWS, int ALT, int DOFF2 = 128>
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
              if (KIND) mma_bf16_split(base_u + dsel + (uint32_t)DOFF2, a10 + ao, b0 + bo, desc_hi, id2, 1u);
              else mma_tf32_split(base_u + dsel + (uint32_t)DOFF2, a10 + ao, b0 + bo, desc_hi, id2, 1u);
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

// ------------------------------------------------------------------------------------------------
// Overlap experiment: warp 0 streams fp16 MMAs (M128 N64 K16 + M128 N32 K16 per step, operands in shared
// memory, the production hidden-layer step) while warps 4..7 run a CUDA-core workload:
//   work 0 nothing, 1 FFMA chain, 2 STS.128 + LDS.128, 3 packed fp16 conversions, 4 tcgen05.ld, 5 SHFL,
//        6 LDG (L1-resident), 7 LDS.128 only, 8 STS.128 only, 9 mbarrier arrive + wait, 10 bar.sync,
//        11 STS + fence.proxy.async, 12 tcgen05 fences, 13 plane store + fence + mbarrier round
// mode bit 0 = run the MMAs, bits 1.. = work.  cycles[2*b] = MMA stream, cycles[2*b+1] = CUDA stream.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) tc_overlap_kernel(int reps, int mode, int iters,
                                                            long long* __restrict__ cycles, float* __restrict__ sink) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t plane_bytes = 260u * 16u, b_plane_bytes = 64u * 16u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  unsigned char* a_0 = smem_raw + 128;
  unsigned char* a_1 = a_0 + 4 * plane_bytes;
  unsigned char* b_cat = a_1 + 4 * plane_bytes;
  unsigned char* scratch = b_cat + kTaps * 4 * b_plane_bytes;       // 4 warps x 32 lanes x 16 B x 4
  const uint32_t init_words = (2 * 4 * plane_bytes + kTaps * 4 * b_plane_bytes + 8192) / 4;
  for (uint32_t i = tid; i < init_words; i += blockDim.x) reinterpret_cast<uint32_t*>(a_0)[i] = 0x3c003c00u + (i & 63u);
  if (tid == 0) {
    mbar_init(bar, 1);
    for (int w = 0; w < 4; ++w) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 32) + w, 32);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  const bool run_mma = mode & 1;
  const int work = mode >> 1;
  if (warp == 0 && run_mma) {
    const uint32_t base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t plane16 = plane_bytes >> 4, bplane16 = b_plane_bytes >> 4;
    const uint32_t a00 = ((smem_u32(a_0) >> 4) & 0x3FFFu) | (plane16 << 16);
    const uint32_t a10 = ((smem_u32(a_1) >> 4) & 0x3FFFu) | (plane16 << 16);
    const uint32_t b0 = ((smem_u32(b_cat) >> 4) & 0x3FFFu) | (bplane16 << 16);
    const uint32_t id1 = instr_desc_f16(128, 64), id2 = instr_desc_f16(128, 32);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t ao = (uint32_t)(2 * kb) * plane16 + (uint32_t)k;
            const uint32_t bo = (uint32_t)(k * 4 + 2 * kb) * bplane16;
            mma_f16_split(base_u + 256u, a00 + ao, b0 + bo, desc_hi, id1, 1u);
            mma_f16_split(base_u + 288u, a10 + ao, b0 + bo, desc_hi, id2, 1u);
          }
        }
      }
    }
    __syncwarp();
    if (elect_one()) mma_commit(bar);
    __syncwarp();
    mbar_wait_guarded(bar, 0);
    if (tid == 0) cycles[2 * blockIdx.x] = clock64() - t0;
  }
  if (warp >= 4 && work > 0) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (float)(lane + i);
    uint4* mine = reinterpret_cast<uint4*>(scratch) + (warp - 4) * 128 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (work == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
      } else if (work == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v = mine[j * 32];
          v.x += (uint32_t)it;
          mine[j * 32] = v;
        }
      } else if (work == 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            uint32_t hi, lo;
            split_half2(acc[i], acc[i + 1], hi, lo);
            acc[i] += __uint_as_float(hi & 0x3fffffu);
            acc[i + 1] += __uint_as_float(lo & 0x3fffffu);
          }
      } else if (work == 4) {
        uint32_t r[16];
        tmem_ld16_issue(taddr + (uint32_t)((it & 7) * 16), r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(r[i]) + __uint_as_float(r[i + 8]);
      } else if (work == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], 1);
      } else if (work == 6) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += __ldg(sink + 256 + ((it + j * 32 + lane) & 1023));
      } else if (work == 7) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = mine[j * 32];
          acc[j] += __uint_as_float(v.x & 0x3fffffu);
        }
      } else if (work == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mine[j * 32] = make_uint4((uint32_t)it, 0u, 0u, 0u);
      } else if (work == 9) {
        // mbarrier round: every lane arrives on the warp's own barrier, then waits for the phase
        uint64_t* wb = reinterpret_cast<uint64_t*>(smem_raw + 32) + (warp - 4);
        mbar_arrive(wb);
        mbar_wait_guarded(wb, (uint32_t)(it & 1));
      } else if (work == 10) {
        asm volatile("bar.sync %0, 128;" ::"r"(1) : "memory");     // named barrier among warps 4..7
      } else if (work == 11) {
        mine[0] = make_uint4((uint32_t)it, 0u, 0u, 0u);
        fence_async_smem();                                        // generic -> async proxy fence after a store
      } else if (work == 12) {
        fence_before();
        fence_after();
      } else {
        // plane store as the row kernel does it: 8 STS.128 into a [chunk][pos] plane + fence + mbarrier arrive
        uint4* plane = reinterpret_cast<uint4*>(scratch);
#pragma unroll
        for (int j = 0; j < 4; ++j) plane[j * 128 + (warp - 4) * 32 + lane] = make_uint4((uint32_t)it, 1u, 2u, 3u);
        fence_async_smem();
        uint64_t* wb = reinterpret_cast<uint64_t*>(smem_raw + 32) + (warp - 4);
        mbar_arrive(wb);
        mbar_wait_guarded(wb, (uint32_t)(it & 1));
      }
    }
    const long long t1 = clock64();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += acc[i];
    if (sum == 1.2345f) sink[tid] = sum;
    if (tid == 128) cycles[2 * blockIdx.x + 1] = t1 - t0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace ddd1d
