// CTA-per-row fused WENO5 integrator with the row in registers (float32 WENO, BASELINE config 5:
// Godunov Burgers, N = 2048), sm_100a.
//
// row_kernel<MODE_WENO> keeps the stage row, the stage derivatives and the flux in shared memory and walks
// the points with a strided loop: 576 thread instructions per point-stage, of which 35 % are IEEE divisions
// and 9 % the forcing basis loads (profiles/r01/row_kernel_weno_c5_ncu_full.json).  Here a thread owns PPT = 4
// CONSECUTIVE points for the whole launch -- the float64-equivalent solution as an unevaluated float pair (no
// FP64 instruction or float64 conversion in the loop), float32 stage derivatives and the stage values in
// registers -- so that
//   * the smoothness indicators (weno.py:46-57) are computed once per stencil centre and shared by the right
//     reconstruction of point p and the left reconstruction of point p + 1 (the reference computes them twice);
//   * 1 / (eps + beta)^2 is one MUFU.RCP (1 ulp) per indicator, shared by both sides, the weight normalisation
//     one reciprocal per side and the /3, /6 of weno.py:82-88 multiplications (3 + 2 reciprocals per point instead
//     of 22 IEEE divisions; every factor within 1-2 ulp of the reference's quotient, far inside the 2e-5 gate of
//     the float32 WENO path, tests/test_gpu_parity.py);
//   * the flux is evaluated at PPT + 1 points per thread, so the flux difference needs no second exchange;
//   * halos (3 left, 4 right) come from the neighbouring lanes by shuffles; only the first / last lane of a warp
//     goes through shared memory (one block barrier per stage, double buffered);
//   * the forcing basis of the thread's 4 points is 6 LDG.128 per stage from L1; the mode amplitudes are
//     computed by warp 0 (one forcing term per lane) ahead of the same barrier: one sincosf per STEP rotated to
//     the later stages by angle addition, fixed-point warp sums on REDUX's integer adder (see ddd1d_warp.cuh),
//     so that warp 0 is not the warp everybody waits for.
// Reference: weno.py:43-123, integrate.WENODifferentiator (integrate.py:124-140), model.py:81-97 (float32
// form), equations.py:341-370 (Godunov flux), :196-227 (forcing), integrate.odeint's Bogacki-Shampine steps.
#pragma once
#include "ddd1d_device.cuh"

namespace ddd1d {

struct WenoBeta { float i0, i1, i2; };      // 1 / (eps + beta_k)^2 of one stencil centre

// MUFU.RCP: reciprocal to 1 ulp, no slow path (arguments here are positive and normal for any finite row)
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// weno.py:46-57 + the reciprocal squares of weno.py:60-66, centred on c
__device__ __forceinline__ WenoBeta weno_inverse_beta(float a, float b, float c, float d, float e) {
  const float eps = 1e-6f;
  const float q0 = a - 4.f * b + 3.f * c, r0 = a - 2.f * b + c;
  const float q1 = b - d, r1 = b - 2.f * c + d;
  const float q2 = 3.f * c - 4.f * d + e, r2 = c - 2.f * d + e;
  const float b0 = 0.25f * (q0 * q0) + (13.f / 12.f) * (r0 * r0);
  const float b1 = 0.25f * (q1 * q1) + (13.f / 12.f) * (r1 * r1);
  const float b2 = 0.25f * (q2 * q2) + (13.f / 12.f) * (r2 * r2);
  WenoBeta o;
  o.i0 = rcp_fast((eps + b0) * (eps + b0));
  o.i1 = rcp_fast((eps + b1) * (eps + b1));
  o.i2 = rcp_fast((eps + b2) * (eps + b2));
  return o;
}

// left-biased reconstruction (weno.py:76-97) from the five points a..e centred on c, indicators `bt` of c
__device__ __forceinline__ float weno_left(const WenoBeta& bt, float a, float b, float c, float d, float e) {
  const float a0 = 0.1f * bt.i0, a1 = 0.6f * bt.i1, a2 = 0.3f * bt.i2;
  const float rs = rcp_fast(a0 + a1 + a2);
  const float w0 = a0 * rs, w1 = a1 * rs, w2 = a2 * rs;
  constexpr float k3 = 1.f / 3.f, k6 = 1.f / 6.f;      // (the reference divides; x * (1/6) is within 1 ulp of x / 6)
  return (w0 * k3) * a + (-(7.f * w0 + w1) * k6) * b + ((11.f * w0 + 5.f * w1 + 2.f * w2) * k6) * c +
         ((2.f * w1 + 5.f * w2) * k6) * d + (-w2 * k6) * e;
}

// right-biased reconstruction at the cell LEFT of the centre (weno.py:100-123): points a..e centred on c
__device__ __forceinline__ float weno_right(const WenoBeta& bt, float a, float b, float c, float d, float e) {
  const float a0 = 0.3f * bt.i0, a1 = 0.6f * bt.i1, a2 = 0.1f * bt.i2;
  const float rs = rcp_fast(a0 + a1 + a2);
  const float w2 = a0 * rs, w1 = a1 * rs, w0 = a2 * rs;    // omega2, omega1, omega0 of weno.py:106-108
  constexpr float k3 = 1.f / 3.f, k6 = 1.f / 6.f;
  return (-w2 * k6) * a + ((5.f * w2 + 2.f * w1) * k6) * b + ((2.f * w2 + 5.f * w1 + 11.f * w0) * k6) * c +
         (-(w1 + 7.f * w0) * k6) * d + (w0 * k3) * e;
}

constexpr int kWenoPpt = 4;              // consecutive points per thread
constexpr int kWenoHaloL = 3, kWenoHaloR = 4;

struct WenoShared {
  float edge[2][32][2][4];               // [stage parity][warp][first lane's points 0..3 | last lane's points 1..3]
  float amps[2][2 * kMaxModes];          // [stage parity][sine sums 0..7 | cosine sums 0..7]
  float rot[kMaxStages][32][2];          // [stage][lane of warp 0]: cos, sin of w c_s dt (rotation to stage s)
  float term[32][4];                     // [lane of warp 0]: fixed-point amplitude, its copy signed by k, mode index
  unsigned int first_bad;
};

// EQ: EQ_BURGERS_GOD for the instantiation specialised for BASELINE config 5 (folds equation_point's switch, an
// indirect branch per point and stage), -1 = read P.eq
#ifndef DDD1D_WENO_MIN_BLOCKS
#define DDD1D_WENO_MIN_BLOCKS 2          // CTAs per SM the register budget is set for (A/B builds: 1 = 128 registers)
#endif
template <int EQ>
__global__ void __launch_bounds__(512, DDD1D_WENO_MIN_BLOCKS) weno_block_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W,
                                                            const __grid_constant__ Tableau tab) {
  constexpr int PPT = kWenoPpt;
  __shared__ WenoShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = P.N;                       // == PPT * blockDim.x
  const int p0 = tid * PPT;
  const int eq = EQ >= 0 ? EQ : P.eq;
  const bool forced = eq_forced(eq) && P.P > 0;
  const float4* const basis4 = reinterpret_cast<const float4*>(P.fbasis + p0);

  // window-form stencils of the non-WENO derivative channels (d >= 2), in registers
  // (Godunov Burgers has one: u_x.  Its instantiation drops the second row, 7 registers and 7 FFMAs per point.)
  constexpr int NCF = EQ == EQ_BURGERS_GOD ? 1 : 2;
  float cf[NCF][kWin];
#pragma unroll
  for (int d = 0; d < NCF; ++d)
#pragma unroll
    for (int j = 0; j < kWin; ++j) cf[d][j] = d + 2 < P.D ? __ldg(P.blob + P.st_off + (d + 2) * kWinPad + j) : 0.f;

  uint32_t par = 0;
  for (int row = blockIdx.x; row < W.batch; row += gridDim.x) {
    const int sample = W.sample_offset + row;
    const ForcingTerm fterm = warp == 0 ? load_forcing_term(P, sample, lane) : ForcingTerm{0.f, 0.f, 0.f, 0.f};
    // warp 0: term amplitude in fixed-point units (2^25 / largest amplitude of the row), cosine copy signed by k,
    // mode |k| of this lane's term, cos / sin of w c_s dt (the rotation from the step's angle to stage s)
    float inv_scale = 0.f;
    if (forced && warp == 0) {
      const bool on = lane < P.P;
      const unsigned int amax_bits = __reduce_max_sync(0xffffffffu, on ? __float_as_uint(fabsf(fterm.a)) : 0u);
      const int eb = (int)(amax_bits >> 23) + ((amax_bits & 0x7fffffu) ? 1 : 0);
      const int es = min(max(279 - eb, 1), 254);                 // 2^(25 - e), 2^e >= the largest |a|
      const float scale = amax_bits == 0u ? 1.f : __uint_as_float((unsigned int)es << 23);
      inv_scale = __uint_as_float((254u << 23) - __float_as_uint(scale));
      const float a_s = on ? fterm.a * scale : 0.f;
      *reinterpret_cast<float4*>(sh.term[lane]) =
          make_float4(a_s, fterm.k < 0.f ? -a_s : a_s, __int_as_float(on ? (int)fabsf(fterm.k) - 1 : -1), 0.f);
      for (int q = 0; q < tab.stages; ++q) {       // (read back by the same lane only)
        float rs, rc;
        sincosf(fterm.w * (float)(tab.c[q] * W.dt), &rs, &rc);
        sh.rot[q][lane][0] = rc;
        sh.rot[q][lane][1] = rs;
      }
    }
    float yh[PPT], yl[PPT];                        // solution = yh + yl, yh = float(yh + yl)
    float k0[PPT], k1[PPT], k2[PPT], k3[PPT];      // stage derivatives (static indexing keeps them in registers)
    {
      const float4 v = __ldg(reinterpret_cast<const float4*>(W.u + (size_t)row * N + p0));
      yh[0] = v.x; yh[1] = v.y; yh[2] = v.z; yh[3] = v.w;
      yl[0] = yl[1] = yl[2] = yl[3] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) k0[i] = k1[i] = k2[i] = k3[i] = 0.f;
    if (tid == 0) sh.first_bad = 0xffffffffu;
    int first_bad = -1;
    int save_idx = 0, until_save = W.save_every;       // a down-counter: no division per step

    for (int step = 0; step < W.nsteps; ++step) {
      float sn0 = 0.f, cs0 = 0.f;
      if (forced && warp == 0) sincosf(fmaf(fterm.w, (float)(W.t0 + (double)step * W.dt), fterm.phi), &sn0, &cs0);
#pragma unroll 1
      for (int s = 0; s < tab.stages; ++s) {
        // ---- stage values, rounded to float32 (integrate.py:57-60,71); a[s][j] = 0 for j >= s ----
        float E[PPT + kWenoHaloL + kWenoHaloR];        // E[j] = u[p0 + j - 3]
        // (the increment in float32, a_j = float(dt * a[s][j]), added low part first: as the tensor engine does)
        const float a0 = s > 0 ? W.adt[s][0] : 0.f, a1 = s > 1 ? W.adt[s][1] : 0.f, a2 = s > 2 ? W.adt[s][2] : 0.f;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          const float inc = fmaf(a2, k2[i], fmaf(a1, k1[i], a0 * k0[i]));
          E[kWenoHaloL + i] = s == 0 ? yh[i] : yh[i] + (yl[i] + inc);
        }
        // ---- halo exchange: shuffles inside the warp, shared memory across warps ----
        if (lane == 0) *reinterpret_cast<float4*>(sh.edge[par][warp][0]) = make_float4(E[3], E[4], E[5], E[6]);
        if (lane == 31) *reinterpret_cast<float4*>(sh.edge[par][warp][1]) = make_float4(E[4], E[5], E[6], 0.f);
        if (forced && warp == 0) {
          // mode amplitudes of this stage (equations.py:214-219): the step's sine / cosine rotated by w c_s dt,
          // one forcing term per lane, exact fixed-point warp sums
          const float2 rot = *reinterpret_cast<const float2*>(sh.rot[s][lane]);
          const float rc = rot.x, rs = rot.y;
          const float sn = fmaf(sn0, rc, cs0 * rs), cs = fmaf(cs0, rc, -(sn0 * rs));
          const float4 term = *reinterpret_cast<const float4*>(sh.term[lane]);
          const int my_mode = __float_as_int(term.z);
          const int is = __float2int_rn(term.x * sn), ic = __float2int_rn(term.y * cs);
          float mine_s = 0.f, mine_c = 0.f;
#pragma unroll
          for (int m = 0; m < kMaxModes; ++m) {
            if (m >= P.M) break;
            const float a = (float)__reduce_add_sync(0xffffffffu, my_mode == m ? is : 0) * inv_scale;
            const float b = (float)__reduce_add_sync(0xffffffffu, my_mode == m ? ic : 0) * inv_scale;
            if (lane == m) { mine_s = a; mine_c = b; }
          }
          if (lane < kMaxModes) {
            sh.amps[par][lane] = mine_s;
            sh.amps[par][kMaxModes + lane] = mine_c;
          }
        }
        {
          // from the lane below: its points 1..3; from the lane above: its points 0..3
          const float l1 = __shfl_up_sync(0xffffffffu, E[4], 1), l2 = __shfl_up_sync(0xffffffffu, E[5], 1),
                      l3 = __shfl_up_sync(0xffffffffu, E[6], 1);
          const float r0 = __shfl_down_sync(0xffffffffu, E[3], 1), r1 = __shfl_down_sync(0xffffffffu, E[4], 1),
                      r2 = __shfl_down_sync(0xffffffffu, E[5], 1), r3 = __shfl_down_sync(0xffffffffu, E[6], 1);
          E[0] = l1; E[1] = l2; E[2] = l3;
          E[7] = r0; E[8] = r1; E[9] = r2; E[10] = r3;
        }
        __syncthreads();
        if (lane == 0) {
          const float4 v = *reinterpret_cast<const float4*>(sh.edge[par][warp == 0 ? nwarps - 1 : warp - 1][1]);
          E[0] = v.x; E[1] = v.y; E[2] = v.z;
        }
        if (lane == 31) {
          const float4 v = *reinterpret_cast<const float4*>(sh.edge[par][warp + 1 == nwarps ? 0 : warp + 1][0]);
          E[7] = v.x; E[8] = v.y; E[9] = v.z; E[10] = v.w;
        }
        // ---- WENO5 reconstructions and the flux at points p0 .. p0 + PPT ----
        // inverse smoothness indicators of the centres p0 - 1 .. p0 + PPT (E index 2 .. PPT + 3)
        WenoBeta bt[PPT + 2];
#pragma unroll
        for (int c = 0; c < PPT + 2; ++c) bt[c] = weno_inverse_beta(E[c], E[c + 1], E[c + 2], E[c + 3], E[c + 4]);
        float flux[PPT + 1];
#pragma unroll
        for (int i = 0; i <= PPT; ++i) {
          // point p = p0 + i: u_minus = left reconstruction centred on p - 1, u_plus = right one centred on p
          // (weno.py:92-97,118-123 and the roll by +1 of integrate.py:137-138)
          float dv[kMaxD];
          dv[0] = weno_left(bt[i], E[i], E[i + 1], E[i + 2], E[i + 3], E[i + 4]);
          dv[1] = weno_right(bt[i + 1], E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5]);
          dv[3] = 0.f;
#pragma unroll
          for (int d = 0; d < NCF; ++d) {
            float acc = 0.f;                       // constant stencil rows (model.py:99-109, 536-548)
#pragma unroll
            for (int j = 0; j < kWin; ++j) acc = fmaf(cf[d][j], E[i + j], acc);
            dv[2 + d] = acc;
          }
          flux[i] = equation_point(eq, E[kWenoHaloL + i], dv, P.eta);
        }
        // ---- y_t = -(1/dx) (flux[x+1] - flux[x]) + forcing (equations.py:305-320, 276-277) ----
        float f[PPT] = {0.f, 0.f, 0.f, 0.f};
        if (forced) {
#pragma unroll
          for (int m = 0; m < kMaxModes; ++m) {
            if (m >= P.M) break;
            const float as = sh.amps[par][m], ac = sh.amps[par][kMaxModes + m];
            const float4 bc = __ldg(basis4 + (size_t)m * (N / 4)), bs = __ldg(basis4 + (size_t)(P.M + m) * (N / 4));
            f[0] = fmaf(as, bc.x, f[0]); f[0] = fmaf(ac, bs.x, f[0]);
            f[1] = fmaf(as, bc.y, f[1]); f[1] = fmaf(ac, bs.y, f[1]);
            f[2] = fmaf(as, bc.z, f[2]); f[2] = fmaf(ac, bs.z, f[2]);
            f[3] = fmaf(as, bc.w, f[3]); f[3] = fmaf(ac, bs.w, f[3]);
          }
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          float r = -__fmul_rn(P.inv_dx, __fsub_rn(flux[i + 1], flux[i]));
          if (forced) r = __fadd_rn(r, f[i]);
          if (s == 0) k0[i] = r;
          else if (s == 1) k1[i] = r;
          else if (s == 2) k2[i] = r;
          else k3[i] = r;
        }
        par ^= 1u;
      }
      // ---- end of the step: y += dt * sum b k in float-float (the float64 sum of the reference to ~2^-48) ----
      const bool save = --until_save == 0;
      float out[PPT];
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const float kk[kMaxStages] = {k0[i], k1[i], k2[i], k3[i]};
        float ih = W.bdt_hi[0] * kk[0];
        float il = fmaf(W.bdt_hi[0], kk[0], -ih) + W.bdt_lo[0] * kk[0];
#pragma unroll
        for (int j = 1; j < kMaxStages; ++j) {
          if (j >= tab.stages) break;
          const float ph = W.bdt_hi[j] * kk[j];
          const float pl = fmaf(W.bdt_hi[j], kk[j], -ph) + W.bdt_lo[j] * kk[j];
          float sh_, er;
          warp_two_sum(ih, ph, sh_, er);
          ih = sh_;
          il += er + pl;
        }
        float nh, er;
        warp_two_sum(yh[i], ih, nh, er);
        const float nl = yl[i] + (er + il);
        const float yn = nh + nl;
        yl[i] = W.state_f32 ? 0.f : nl - (yn - nh);     // float32 carry (tf odeint_fixed, model.py:138-159)
        yh[i] = yn;
        if (first_bad < 0 && !isfinite(yn)) first_bad = step;
        out[i] = yn;
      }
      if (save) {
        *reinterpret_cast<float4*>(W.snaps + ((size_t)save_idx * W.batch + row) * N + p0) =
            make_float4(out[0], out[1], out[2], out[3]);
        ++save_idx;
        until_save = W.save_every;
      }
    }
    if (W.first_bad) {
      if (first_bad >= 0) atomicMin(&sh.first_bad, (unsigned int)first_bad);
      __syncthreads();
      if (tid == 0) W.first_bad[row] = sh.first_bad == 0xffffffffu ? -1 : (int)sh.first_bad;
    }
    __syncthreads();       // the next row reuses the shared words
  }
}

}  // namespace ddd1d
