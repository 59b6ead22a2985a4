// Warp-per-row fused integrator for the fixed-stencil and WENO5 modes (no conv net), sm_100a.
//
// The CTA-per-row kernel (row_kernel) needs five block barriers and a shared-memory round trip per
// Runge-Kutta stage; for the short rows of the baseline configurations (N = 64 in BASELINE config 1 and its
// batched twin) that is nearly all it does.  Here ONE WARP owns a row for the whole launch:
//   * lane l holds the PPL = N / 32 consecutive points l*PPL .. l*PPL + PPL-1: float64 solution, float32
//     stage derivatives and the stage row all live in registers;
//   * the three halo points on either side come from the neighbouring lanes by shuffles (the row is periodic
//     and exactly one warp wide, so lane -1 is lane 31);
//   * forcing: lane q evaluates term q of the sample's RandomForcing (one sincosf per stage), the amplitude of
//     every spatial mode is a warp sum (xor shuffles), and the spatial basis is read from L1;
//   * the flux difference of the conservative forms needs the flux of the next point: one more shuffle.
// No shared memory, no block barriers.  Arithmetic follows row_kernel operation by operation (same helpers:
// equation_point, weno_pair, the stage / update formulas in float64), so both satisfy the same tolerances.
//
// Reference: model.baseline_space_derivatives (model.py:59-112), polynomials.reconstruct
// (polynomials.py:280-303), equations.*.equation_of_motion, RandomForcing (equations.py:196-227),
// weno.py:43-123, integrate.odeint's Bogacki-Shampine steps (integrate.py:143-169).
#pragma once
#include "ddd1d_device.cuh"

namespace ddd1d {

// MM: compile-time bound on the number of forcing modes (4 covers the reference's k_max = 3)
template <int PPL, bool WENO, int MM>
__global__ void __launch_bounds__(256, PPL <= 2 ? 4 : PPL == 4 ? 2 : 1) warp_row_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W,
                                                       const __grid_constant__ Tableau tab) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const int N = P.N;                       // == 32 * PPL
  const bool cons = eq_conservative(P.eq);
  const bool forced = eq_forced(P.eq) && P.P > 0;

  // window-form stencils of the derivative channels (ddd1d_set_stencils), in registers
  float cf[kMaxD][kWin];
#pragma unroll
  for (int d = 0; d < kMaxD; ++d)
#pragma unroll
    for (int j = 0; j < kWin; ++j) cf[d][j] = d < P.D ? __ldg(P.blob + P.st_off + d * kWinPad + j) : 0.f;

  for (int row = gwarp; row < W.batch; row += total_warps) {
    const int sample = W.sample_offset + row;
    const ForcingTerm fterm = load_forcing_term(P, sample, lane);
    double y[PPL];
    float k[kMaxStages][PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i)
      y[i] = W.u64 ? W.u64[(size_t)row * N + lane * PPL + i] : (double)__ldg(W.u + (size_t)row * N + lane * PPL + i);
    bool bad_seen = false;
    int first_bad = -1;
    int save_idx = 0;

    for (int step = 0; step < W.nsteps; ++step) {
      const double t = W.t0 + (double)step * W.dt;
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) {
        if (s >= tab.stages) break;
        // ---- stage row y + dt * sum_j a[s][j] k_j, rounded to float32 (integrate.py:57-60,71) ----
        float e[PPL + 2 * kHalo];            // e[kHalo + i] = u[point i of this lane]
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < kMaxStages; ++j)
            if (j < s && tab.a[s][j] != 0.0) acc += tab.a[s][j] * (double)k[j][i];
          e[kHalo + i] = (float)(s == 0 ? y[i] : y[i] + W.dt * acc);
        }
        // ---- periodic halo from the neighbouring lanes ----
#pragma unroll
        for (int h = 1; h <= kHalo; ++h) {
          const int dl = (h + PPL - 1) / PPL;                  // lanes to the left
          const int il = (PPL - (h % PPL)) % PPL;              // its point index
          e[kHalo - h] = __shfl_sync(0xffffffffu, e[kHalo + il], (lane - dl) & 31);
          const int dr = (PPL - 1 + h) / PPL;                  // lanes to the right
          const int ir = (PPL - 1 + h) % PPL;
          e[kHalo + PPL - 1 + h] = __shfl_sync(0xffffffffu, e[kHalo + ir], (lane + dr) & 31);
        }
        // ---- forcing amplitudes of this stage (equations.py:214-219) ----
        float amp[2 * MM];
        if (forced) {
          const float ts = (float)(t + tab.c[s] * W.dt);
          float sn, cs;
          sincosf(fmaf(fterm.w, ts, fterm.phi), &sn, &cs);
          const bool on = lane < P.P;
          const float a_sin = on ? fterm.a * sn : 0.f;
          const float a_cos = on ? (fterm.k < 0.f ? -fterm.a : fterm.a) * cs : 0.f;
          const float ka = fabsf(fterm.k);
#pragma unroll
          for (int m = 0; m < MM; ++m) {
            amp[m] = amp[MM + m] = 0.f;
            if (m >= P.M) continue;
            float a = ka == (float)(m + 1) ? a_sin : 0.f, b = ka == (float)(m + 1) ? a_cos : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            amp[m] = a;
            amp[MM + m] = b;
          }
        }
        // ---- derivatives, equation of motion ----
        float r[PPL];
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          float dv[kMaxD];
#pragma unroll
          for (int d = 0; d < kMaxD; ++d) {
            float acc = 0.f;                 // einsum('bxdi,bxi->bxd') with constant rows (model.py:536-548)
            if (d < P.D && !(WENO && d < 2)) {      // (warp-uniform; WENO overwrites channels 0 and 1 below)
#pragma unroll
              for (int j = 0; j < kWin; ++j) acc = fmaf(cf[d][j], e[i + j], acc);
            }
            dv[d] = acc;
          }
          if (WENO) {                        // u_minus / u_plus replaced by WENO5 (integrate.py:134-138)
            float um, up;
            weno_pair<float>(&e[kHalo + i], um, up);
            dv[0] = um;
            dv[1] = up;
          }
          r[i] = equation_point(P.eq, e[kHalo + i], dv, P.eta);
        }
        if (cons) {
          // y_t = -(1/dx) (flux[x+1] - flux[x])  (equations.py:305-320)
          const float next_lane_first = __shfl_sync(0xffffffffu, r[0], (lane + 1) & 31);
#pragma unroll
          for (int i = 0; i < PPL; ++i) {
            const float fwd = i + 1 < PPL ? r[i + 1] : next_lane_first;
            r[i] = -__fmul_rn(P.inv_dx, __fsub_rn(fwd, r[i]));
          }
        }
        if (forced) {
#pragma unroll
          for (int i = 0; i < PPL; ++i) {
            const float* basis = P.fbasis + lane * PPL + i;
            float f = 0.f;
#pragma unroll
            for (int m = 0; m < MM; ++m)
              if (m < P.M) {
                f = fmaf(amp[m], __ldg(basis + (size_t)m * N), f);
                f = fmaf(amp[MM + m], __ldg(basis + (size_t)(P.M + m) * N), f);
              }
            r[i] = __fadd_rn(r[i], f);
          }
        }
#pragma unroll
        for (int i = 0; i < PPL; ++i) k[s][i] = r[i];
      }
      // ---- end of the step: float64 update, snapshot ----
      const bool save = ((step + 1) % W.save_every) == 0;
      float* snap = save ? W.snaps + ((size_t)save_idx * W.batch + row) * N + lane * PPL : nullptr;
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxStages; ++j)
          if (j < tab.stages && tab.b[j] != 0.0) acc += tab.b[j] * (double)k[j][i];
        double yn = y[i] + W.dt * acc;
        if (W.state_f32) yn = (double)(float)yn;     // float32 carry (tf odeint_fixed, model.py:138-159)
        y[i] = yn;
        if (!bad_seen && !isfinite(yn)) { bad_seen = true; first_bad = step; }
        if (save) snap[i] = (float)yn;
      }
      if (save) ++save_idx;
    }
    if (W.first_bad) {
      unsigned int key = first_bad < 0 ? 0xffffffffu : (unsigned int)first_bad;
      key = __reduce_min_sync(0xffffffffu, key);
      if (lane == 0) W.first_bad[row] = key == 0xffffffffu ? -1 : (int)key;
    }
  }
}

}  // namespace ddd1d
