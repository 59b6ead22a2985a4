// Laboratory kernels of the tensor engine: descriptor probe, MMA issue-rate and overlap microbenchmarks.
// TEST-ONLY: built into libddd1d_debug.so (include/ddd1d_debug.h), never into the product library.
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/ddd1d.h"
#include "../../include/ddd1d_debug.h"
#include "ddd1d_tc_common.cuh"

namespace ddd1d {
namespace tc {

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
// Probe of the TMEM-resident A operand: tcgen05.cp (128x256b, 4x256b), tcgen05.shift.down, tcgen05.mma with A in
// TMEM (+ .ashift).  Dumps raw TMEM contents after every step so that the host can read off the semantics.
//   out  uint32 [10][128][16]
// Planes: raw 16-bit pattern  p | ((chunk * 8 + j) << 8)  at (chunk, position p, half j); a second plane set holds
// small integers as fp16 for the MMA checks: A2[p][k] = ((p * 3 + k) % 7) - 3,  B[n][k] = ((n + 2 * k) % 5) - 2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_cp(uint32_t taddr, uint64_t desc, int shape) {
  if (shape == 0) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
  else asm volatile("tcgen05.cp.cta_group::1.4x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tc_shift(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void mma_f16_ta(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                           bool ashift) {
  if (ashift)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.ashift [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(160, 1) tc_shift_probe_kernel(uint32_t* __restrict__ out) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int P = 160;
  constexpr uint32_t plane = P * 16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  unsigned short* raw = reinterpret_cast<unsigned short*>(smem_raw + 128);            // [2][P][8]
  __half* a2 = reinterpret_cast<__half*>(smem_raw + 128 + 2 * plane);                 // [2][P][8]
  __half* b = reinterpret_cast<__half*>(smem_raw + 128 + 4 * plane);                  // [2][16][8]
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(slot, 64);
  for (int i = tid; i < 2 * P * 8; i += blockDim.x) {
    const int c = i / (P * 8), p = (i / 8) % P, j = i % 8, k = c * 8 + j;
    raw[i] = (unsigned short)(p | (k << 8));
    a2[i] = __float2half((float)(((p * 3 + k) % 7) - 3));
  }
  for (int i = tid; i < 2 * 16 * 8; i += blockDim.x) {
    const int c = i / 128, n = (i / 8) % 16, j = i % 8, k = c * 8 + j;
    b[i] = __float2half((float)(((n + 2 * k) % 5) - 2));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t raw_s = smem_u32(raw), a2_s = smem_u32(a2), b_s = smem_u32(b);
  const uint64_t bdesc = smem_desc(b_s, 256, 128);
  const uint32_t idesc = instr_desc_f16(128, 16);
  uint32_t parity = 0;
  // one step: the issuer runs `what`, commits; everybody waits and dumps columns [col, col + 16)
  auto dump = [&](int step, uint32_t col) {
    mbar_wait_guarded(bar, parity);
    parity ^= 1u;
    fence_after();
    if (warp < 4) {
      uint32_t r[16];
      tmem_ld16_issue(tmem_base + ((uint32_t)(warp * 32) << 16) + col, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) out[((size_t)step * 128 + tid) * 16 + i] = r[i];
      fence_before();
    }
    __syncthreads();
    fence_after();
  };
  const bool issuer = warp == 4 && elect_one();
  // 0: 128x256b copy of positions 2..129
  if (issuer) { tc_cp(tmem_base, smem_desc(raw_s + 2 * 16, plane, 128), 0); mma_commit(bar); }
  dump(0, 0);
  // 1: one shift
  if (issuer) { tc_shift(tmem_base); mma_commit(bar); }
  dump(1, 0);
  // 2: 4x256b copy from position 100
  if (issuer) { tc_cp(tmem_base, smem_desc(raw_s + 100 * 16, plane, 128), 1); mma_commit(bar); }
  dump(2, 0);
  // 3: shift, then 4x256b (position 120) back to back (pipeline order)
  if (issuer) { tc_shift(tmem_base); tc_cp(tmem_base, smem_desc(raw_s + 120 * 16, plane, 128), 1); mma_commit(bar); }
  dump(3, 0);
  // 4: 4x256b with lane 31 (and 16, at columns 8..15 so that both show) in the address
  if (issuer) {
    tc_cp(tmem_base + (31u << 16), smem_desc(raw_s + 140 * 16, plane, 128), 1);
    tc_cp(tmem_base + (16u << 16) + 8, smem_desc(raw_s + 150 * 16, plane, 128), 1);
    mma_commit(bar);
  }
  dump(4, 0);
  // 5: fresh copy of the fp16 integers, plain MMA with A in TMEM -> D at column 16
  if (issuer) {
    tc_cp(tmem_base, smem_desc(a2_s, plane, 128), 0);
    mma_f16_ta(tmem_base + 16, tmem_base, bdesc, idesc, 0u, false);
    mma_commit(bar);
  }
  dump(5, 16);
  // 6: MMA.ashift -> D at column 32; 7: plain MMA afterwards -> D at column 48; 8: the A columns afterwards
  if (issuer) {
    mma_f16_ta(tmem_base + 32, tmem_base, bdesc, idesc, 0u, true);
    mma_f16_ta(tmem_base + 48, tmem_base, bdesc, idesc, 0u, false);
    mma_commit(bar);
  }
  dump(6, 32);
  if (issuer) mma_commit(bar);
  dump(7, 48);
  if (issuer) mma_commit(bar);
  dump(8, 0);
  // 9: explicit shift then MMA back to back (pipeline order) -> D at column 16
  if (issuer) {
    tc_shift(tmem_base);
    mma_f16_ta(tmem_base + 16, tmem_base, bdesc, idesc, 0u, false);
    mma_commit(bar);
  }
  dump(9, 16);
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 64);
}


// WAR stress: does a tcgen05.cp that overwrites an A operand in TMEM wait for earlier MMAs that still read it?
// `chain` MMAs (N = 256, accumulating) read A = rows of set 0, then a cp overwrites A with set 1 and one more
// MMA (N = 16) reads it into other columns.  out[0][lane][0..15] = first 16 columns of the chain's D (expected:
// chain * A0 B^T), out[1] = the second MMA's D (expected A1 B^T).  Repeated `rounds` times, mismatches counted.
__global__ void __launch_bounds__(160, 1) tc_war_probe_kernel(uint32_t* __restrict__ out, int chain, int rounds) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int P = 128;
  constexpr uint32_t plane = P * 16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  __half* a0 = reinterpret_cast<__half*>(smem_raw + 128);                    // [2][P][8]
  __half* a1 = reinterpret_cast<__half*>(smem_raw + 128 + 2 * plane);
  __half* b = reinterpret_cast<__half*>(smem_raw + 128 + 4 * plane);          // [2][256][8]
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(slot, 512);
  for (int i = tid; i < 2 * P * 8; i += blockDim.x) {
    const int c = i / (P * 8), p = (i / 8) % P, j = i % 8, k = c * 8 + j;
    a0[i] = __float2half((float)(((p * 3 + k) % 7) - 3));
    a1[i] = __float2half((float)(((p * 5 + 2 * k) % 9) - 4));
  }
  for (int i = tid; i < 2 * 256 * 8; i += blockDim.x) {
    const int c = i / 2048, n = (i / 8) % 256, j = i % 8, k = c * 8 + j;
    b[i] = __float2half((float)(((n + 2 * k) % 5) - 2));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  const uint64_t bdesc = smem_desc(smem_u32(b), 256 * 16, 128);
  const uint32_t idesc_big = instr_desc_f16(128, 256), idesc_small = instr_desc_f16(128, 16);
  const bool issuer = warp == 4 && elect_one();
  uint32_t parity = 0;
  uint32_t bad = 0;
  for (int r = 0; r < rounds; ++r) {
    if (issuer) {
      tc_cp(tmem_base + 480, smem_desc(smem_u32(a0), plane, 128), 0);
      for (int i = 0; i < chain; ++i) mma_f16_ta(tmem_base, tmem_base + 480, bdesc, idesc_big, i ? 1u : 0u, false);
      tc_cp(tmem_base + 480, smem_desc(smem_u32(a1), plane, 128), 0);          // WAR on the A columns
      mma_f16_ta(tmem_base + 256, tmem_base + 480, bdesc, idesc_small, 0u, false);
      mma_commit(bar);
    }
    mbar_wait_guarded(bar, parity);
    parity ^= 1u;
    fence_after();
    if (warp < 4) {
      uint32_t d0[16], d1[16];
      tmem_ld16_issue(tmem_base + ((uint32_t)(warp * 32) << 16), d0);
      tmem_ld16_issue(tmem_base + ((uint32_t)(warp * 32) << 16) + 256, d1);
      tmem_wait_ld();
      for (int n = 0; n < 16; ++n) {
        float e0 = 0.f, e1 = 0.f;
        for (int k = 0; k < 16; ++k) {
          const float bb = (float)(((n + 2 * k) % 5) - 2);
          e0 += (float)(((tid * 3 + k) % 7) - 3) * bb;
          e1 += (float)(((tid * 5 + 2 * k) % 9) - 4) * bb;
        }
        if (__uint_as_float(d0[n]) != e0 * (float)chain) ++bad;
        if (__uint_as_float(d1[n]) != e1) ++bad;
        if (r == rounds - 1) { out[tid * 16 + n] = d0[n]; out[128 * 16 + tid * 16 + n] = d1[n]; }
      }
      fence_before();
    }
    __syncthreads();
    fence_after();
  }
  if (warp < 4) atomicAdd(out + 2 * 128 * 16, bad);
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}


// Rate of the TMEM-resident-A tile-layer pattern: per repetition one 128-position tile of one layer =
//   cross pass: 4 x cp.128x256b (hi kb0, hi kb1, lo kb0, lo kb1), 5 taps x 4 MMAs (N = NB, .ashift) + 4 x cp.4x256b
//   main pass : 2 x cp.128x256b (hi again), 5 taps x 2 MMAs + 2 x cp.4x256b
// all into ONE accumulator block of NB columns (cross terms first).  `issuers` warps run the pattern concurrently
// on their own staging / accumulator columns.  flags bit 0: issue the patches; bit 1: .ashift.
template <int NB>
__global__ void __launch_bounds__(128, 1) tc_ta_rate_kernel(int reps, int issuers, int flags, long long* __restrict__ cycles) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t plane = 260u * 16u, bplane = 2u * NB * 16u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
  unsigned char* a_hi = smem_raw + 128;                         // [4 chunks][260][16 B]
  unsigned char* a_lo = a_hi + 4 * plane;
  unsigned char* bw = a_lo + 4 * plane;                         // [5 taps * 4 chunks][Wh NB | Wl NB][16 B]
  unsigned char* patch = bw + kTaps * 4 * bplane;               // [64 B] blocks
  const uint32_t words = (8 * plane + kTaps * 4 * bplane + 4096) / 4;
  for (uint32_t i = tid; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(a_hi)[i] = 0x3c003c00u + (i & 63u);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  if (warp < issuers) {
    const uint32_t base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t D = base_u + (uint32_t)warp * 64u, S = base_u + 256u + (uint32_t)warp * 32u;
    const uint32_t ahi = smem_u32(a_hi), alo = smem_u32(a_lo), b0 = smem_u32(bw), pt = smem_u32(patch);
    const uint32_t idesc = instr_desc_f16(128, NB);
    const bool do_patch = flags & 1, ash = flags & 2;
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
        tc_cp(S + 0, smem_desc(ahi, plane, 128), 0);
        tc_cp(S + 8, smem_desc(ahi + 2 * plane, plane, 128), 0);
        tc_cp(S + 16, smem_desc(alo, plane, 128), 0);
        tc_cp(S + 24, smem_desc(alo + 2 * plane, plane, 128), 0);
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
          const bool sh = ash && k < kTaps - 1;
          const uint64_t wh0 = smem_desc(b0 + (k * 4 + 0) * bplane, bplane, 128), wh1 = smem_desc(b0 + (k * 4 + 2) * bplane, bplane, 128);
          const uint64_t wl0 = smem_desc(b0 + (k * 4 + 0) * bplane + NB * 16, bplane, 128),
                         wl1 = smem_desc(b0 + (k * 4 + 2) * bplane + NB * 16, bplane, 128);
          mma_f16_ta(D, S + 0, wl0, idesc, k ? 1u : 0u, sh);
          mma_f16_ta(D, S + 8, wl1, idesc, 1u, sh);
          mma_f16_ta(D, S + 16, wh0, idesc, 1u, sh);
          mma_f16_ta(D, S + 24, wh1, idesc, 1u, sh);
          if (do_patch && k < kTaps - 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tc_cp(S + 8 * q + (31u << 16), smem_desc(pt + (k * 4 + q) * 128, 64, 128), 1);
          }
        }
        tc_cp(S + 0, smem_desc(ahi, plane, 128), 0);
        tc_cp(S + 8, smem_desc(ahi + 2 * plane, plane, 128), 0);
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
          const bool sh = ash && k < kTaps - 1;
          const uint64_t wh0 = smem_desc(b0 + (k * 4 + 0) * bplane, bplane, 128), wh1 = smem_desc(b0 + (k * 4 + 2) * bplane, bplane, 128);
          mma_f16_ta(D, S + 0, wh0, idesc, 1u, sh);
          mma_f16_ta(D, S + 8, wh1, idesc, 1u, sh);
          if (do_patch && k < kTaps - 1) {
#pragma unroll
            for (int q = 0; q < 2; ++q) tc_cp(S + 8 * q + (31u << 16), smem_desc(pt + (16 + k * 2 + q) * 128, 64, 128), 1);
          }
        }
      }
    }
    __syncwarp();
    if (elect_one()) mma_commit(&bars[warp]);
    __syncwarp();
    mbar_wait_guarded(&bars[warp], 0);
    if ((tid & 31) == 0) cycles[blockIdx.x * 4 + warp] = clock64() - t0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// Collector-buffer reuse: clocks per MMA of pairs of fp16 MMAs (M = 128, K = 16, width N) that share an operand.
//   mode 0: plain pair, different A tiles, same B           (the production order would re-fetch B)
//   mode 1: tcgen05.mma.ws pair, B kept in collector b0     (fill, then use with the other A tile)
//   mode 2: plain pair on the same A tile, .collector::a::fill then ::lastuse, different B
//   mode 3: plain pair on the same A tile without the qualifiers
template <int N>
__global__ void __launch_bounds__(128, 1) tc_reuse_rate_kernel(int reps, int mode, long long* __restrict__ cycles) {
  unsigned char* const smem_raw = dyn_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t plane = 260u * 16u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
  unsigned char* a0 = smem_raw + 128;                  // two tiles: [2 chunks][260][16 B] each
  unsigned char* a1 = a0 + 2 * plane;
  unsigned char* bw = a1 + 2 * plane;                  // two B tiles [2 chunks][N rows][16 B]
  for (uint32_t i = tid; i < (4 * plane + 4 * N * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a0)[i] = 0x3c003c00u + (i & 63u);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *slot;
  if (warp == 0) {
    const uint32_t base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t A0 = smem_desc(smem_u32(a0), plane, 128), A1 = smem_desc(smem_u32(a1), plane, 128);
    const uint64_t B0 = smem_desc(smem_u32(bw), N * 16, 128), B1 = smem_desc(smem_u32(bw) + 2 * N * 16, N * 16, 128);
    const uint32_t idesc = instr_desc_f16(128, N);
    const uint32_t D0 = base_u, D1 = base_u + 256;
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D0), "l"(A0), "l"(B0), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D1), "l"(A1), "l"(B0), "r"(idesc), "r"(1u) : "memory");
        } else if (mode == 1) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(D0), "l"(A0), "l"(B0), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(D1), "l"(A1), "l"(B0), "r"(idesc), "r"(1u) : "memory");
        } else if (mode == 2) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(D0), "l"(A0), "l"(B0), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(D1), "l"(A0), "l"(B1), "r"(idesc), "r"(1u) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D0), "l"(A0), "l"(B0), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D1), "l"(A0), "l"(B1), "r"(idesc), "r"(1u) : "memory");
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

using namespace ddd1d;

namespace {

thread_local std::string g_error;

int fail(const char* what) {
  g_error = what;
  fprintf(stderr, "ddd1d_debug: %s\n", what);
  return DDD1D_EINVAL;
}

#define CUDA_TRY(h, expr)                                                        \
  do {                                                                           \
    cudaError_t e_ = (expr);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "ddd1d_debug: %s failed: %s\n", #expr, cudaGetErrorString(e_)); \
      return DDD1D_ECUDA;                                                        \
    }                                                                            \
  } while (0)

template <int KIND, int N1, int N2, int BROWS, int ALT, int DOFF2 = 128>
int run_rate(int reps, int blocks, long long* d) {
  const int smem = 128 + 2 * tc::kChunks * 516 * 16 + tc::kTaps * tc::kChunks * BROWS * 16;
  CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_rate_kernel<KIND, N1, N2, BROWS, ALT, DOFF2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::tc_rate_kernel<KIND, N1, N2, BROWS, ALT, DOFF2><<<blocks, 128, smem>>>(reps, d);
  CUDA_TRY(nullptr, cudaGetLastError());
  return DDD1D_OK;
}
template <int N>
int run_reuse_rate(int reps, int mode, long long* d) {
  const int smem = 128 + 4 * 260 * 16 + 4 * N * 16 + 256;
  CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_reuse_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::tc_reuse_rate_kernel<N><<<1, 128, smem>>>(reps, mode, d);
  CUDA_TRY(nullptr, cudaGetLastError());
  return DDD1D_OK;
}

}  // namespace

extern "C" {

// variant: index into a fixed table of compile-time MMA patterns (scripts/tc_rate.py lists them)
int ddd1d_debug_tc_rate(int device, int variant, int reps, int blocks, long long* cycles_host) {
  if (reps < 1 || blocks < 1 || !cycles_host) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  long long* d = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&d, (size_t)blocks * sizeof(long long)));
  int rc = DDD1D_EINVAL;
  switch (variant) {
    case 0: rc = run_rate<0, 16, 0, 16, 1>(reps, blocks, d); break;
    case 1: rc = run_rate<0, 32, 0, 32, 1>(reps, blocks, d); break;
    case 2: rc = run_rate<0, 64, 0, 64, 1>(reps, blocks, d); break;
    case 3: rc = run_rate<0, 128, 0, 128, 1>(reps, blocks, d); break;
    case 4: rc = run_rate<0, 32, 0, 64, 1>(reps, blocks, d); break;    // N=32 out of a 64-row plane
    case 5: rc = run_rate<0, 32, 0, 32, 0>(reps, blocks, d); break;    // single accumulator
    case 6: rc = run_rate<0, 64, 32, 64, 1>(reps, blocks, d); break;   // production hidden layer
    case 7: rc = run_rate<0, 32, 16, 32, 1>(reps, blocks, d); break;   // production last layer
    case 8: rc = run_rate<0, 32, 32, 32, 1>(reps, blocks, d); break;
    case 9: rc = run_rate<1, 32, 0, 32, 1>(reps, blocks, d); break;    // bf16 K=16
    case 10: rc = run_rate<1, 64, 0, 64, 1>(reps, blocks, d); break;
    case 11: rc = run_rate<1, 96, 0, 96, 1>(reps, blocks, d); break;
    case 12: rc = run_rate<1, 96, 64, 96, 1>(reps, blocks, d); break;  // bf16x3 first two of a step
    case 13: rc = run_rate<1, 128, 0, 128, 1>(reps, blocks, d); break;
    case 14: rc = run_rate<1, 64, 32, 64, 0, 128>(reps, blocks, d); break;  // f16 hidden step, separate accumulators
    case 15: rc = run_rate<1, 64, 32, 64, 0, 32>(reps, blocks, d); break;   // f16 hidden step, production D overlap
    case 16: rc = run_rate<1, 32, 16, 32, 0, 16>(reps, blocks, d); break;   // f16 last step, production D overlap
    case 17: rc = run_rate<1, 96, 0, 96, 0>(reps, blocks, d); break;        // one MMA per step, N = 96
    default: return fail("unknown variant");
  }
  if (rc) return rc;
  CUDA_TRY(nullptr, cudaMemcpy(cycles_host, d, (size_t)blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  return DDD1D_OK;
}

int ddd1d_debug_tc_overlap(int device, int mode, int reps, int iters, int blocks, long long* cycles_host) {
  if (reps < 1 || iters < 1 || blocks < 1 || !cycles_host) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  long long* d = nullptr;
  float* sink = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&d, (size_t)blocks * 2 * sizeof(long long)));
  CUDA_TRY(nullptr, cudaMemset(d, 0, (size_t)blocks * 2 * sizeof(long long)));
  CUDA_TRY(nullptr, cudaMalloc(&sink, (256 + 1024) * sizeof(float)));
  CUDA_TRY(nullptr, cudaMemset(sink, 0, (256 + 1024) * sizeof(float)));
  const int smem = 128 + 2 * 4 * 260 * 16 + tc::kTaps * 4 * 64 * 16 + 8192;
  CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::tc_overlap_kernel<<<blocks, 256, smem>>>(reps, mode, iters, d, sink);
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaMemcpy(cycles_host, d, (size_t)blocks * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  CUDA_TRY(nullptr, cudaFree(sink));
  return DDD1D_OK;
}

int ddd1d_debug_tc_shift_probe(int device, unsigned int* out_host) {
  if (!out_host) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  unsigned int* d = nullptr;
  const size_t n = (size_t)10 * 128 * 16;
  CUDA_TRY(nullptr, cudaMalloc(&d, n * sizeof(unsigned int)));
  CUDA_TRY(nullptr, cudaMemset(d, 0xee, n * sizeof(unsigned int)));
  const int smem = 128 + 4 * 160 * 16 + 2 * 16 * 16 + 256;
  tc::tc_shift_probe_kernel<<<1, 160, smem>>>(d);
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaDeviceSynchronize());
  CUDA_TRY(nullptr, cudaMemcpy(out_host, d, n * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  return DDD1D_OK;
}

int ddd1d_debug_tc_war_probe(int device, int chain, int rounds, unsigned int* out_host) {
  if (!out_host || chain < 1 || rounds < 1) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  unsigned int* d = nullptr;
  const size_t n = (size_t)2 * 128 * 16 + 1;
  CUDA_TRY(nullptr, cudaMalloc(&d, n * sizeof(unsigned int)));
  CUDA_TRY(nullptr, cudaMemset(d, 0, n * sizeof(unsigned int)));
  const int smem = 128 + 4 * 128 * 16 + 2 * 256 * 16 + 256;
  tc::tc_war_probe_kernel<<<1, 160, smem>>>(d, chain, rounds);
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaDeviceSynchronize());
  CUDA_TRY(nullptr, cudaMemcpy(out_host, d, n * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  return DDD1D_OK;
}

int ddd1d_debug_tc_ta_rate(int device, int nb, int reps, int issuers, int flags, int blocks, long long* cycles_host) {
  if (reps < 1 || blocks < 1 || issuers < 1 || issuers > 4 || !cycles_host || (nb != 16 && nb != 32)) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  long long* d = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&d, (size_t)blocks * 4 * sizeof(long long)));
  CUDA_TRY(nullptr, cudaMemset(d, 0, (size_t)blocks * 4 * sizeof(long long)));
  const int smem = 128 + 8 * 260 * 16 + tc::kTaps * 4 * 2 * nb * 16 + 4096 + 256;
  if (nb == 32) {
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_ta_rate_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::tc_ta_rate_kernel<32><<<blocks, 128, smem>>>(reps, issuers, flags, d);
  } else {
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_ta_rate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::tc_ta_rate_kernel<16><<<blocks, 128, smem>>>(reps, issuers, flags, d);
  }
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaDeviceSynchronize());
  CUDA_TRY(nullptr, cudaMemcpy(cycles_host, d, (size_t)blocks * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  return DDD1D_OK;
}

int ddd1d_debug_tc_reuse_rate(int device, int n, int mode, int reps, long long* cycles_host) {
  if (reps < 1 || mode < 0 || mode > 3 || !cycles_host) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  long long* d = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&d, sizeof(long long)));
  int rc = DDD1D_EINVAL;
  if (n == 32) rc = run_reuse_rate<32>(reps, mode, d);
  else if (n == 64) rc = run_reuse_rate<64>(reps, mode, d);
  else if (n == 128) rc = run_reuse_rate<128>(reps, mode, d);
  else return fail("n must be 32, 64 or 128");
  if (rc) return rc;
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  CUDA_TRY(nullptr, cudaMemcpy(cycles_host, d, sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(nullptr, cudaFree(d));
  return DDD1D_OK;
}

int ddd1d_debug_tc_probe(int device, const float* x, const float* w_cat, float* out, int nout, void* stream) {
  if (!x || !w_cat || !out || (nout != 16 && nout != 32)) return fail("bad argument");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  const int smem = 128 + 2 * tc::kChunks * 132 * 16 + 2 * tc::kTaps * tc::kChunks * nout * 16;
  CUDA_TRY(nullptr, cudaFuncSetAttribute(tc::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::tc_probe_kernel<<<1, 160, smem, static_cast<cudaStream_t>(stream)>>>(x, w_cat, out, nout);
  CUDA_TRY(nullptr, cudaGetLastError());
  return DDD1D_OK;
}

}  // extern "C"
