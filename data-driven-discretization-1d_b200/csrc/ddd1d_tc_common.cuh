// tcgen05 / TMEM / mbarrier building blocks shared by the tensor-core row kernel (ddd1d_tc.cuh) and the
// test-only laboratory kernels (ddd1d_debug.cu), sm_100a.
#pragma once
#include <cuda_fp16.h>

#include "ddd1d_device.cuh"

namespace ddd1d {
namespace tc {

constexpr int kF = 32;             // hidden width this path is built for
constexpr int kTaps = 5;
constexpr int kChunks = kF / 4;    // 16-byte chunks along ci (TF32 planes; fp16 planes have kChunks / 2)
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

}  // namespace tc
}  // namespace ddd1d
