// Warp-per-row fused integrator for the fixed-stencil and WENO5 modes (no conv net), sm_100a.
//
// The CTA-per-row kernel (row_kernel) needs five block barriers and a shared-memory round trip per
// Runge-Kutta stage; for the short rows of the baseline configurations (N = 64 in BASELINE config 1 and its
// batched twin) that is nearly all it does.  Here ONE WARP owns a row for the whole launch:
//   * lane l holds the PPL = N / 32 consecutive points l*PPL .. l*PPL + PPL-1: the float64-equivalent solution
//     as an unevaluated float pair (no FP64 instruction and no float64 conversion in the loop: conversions
//     run at 16 lanes per clock per SM), float32 stage derivatives and the stage row all live in registers;
//   * the HALO (1..3) points the stencils actually reach on either side come from the neighbouring lanes by
//     shuffles (the row is periodic and exactly one warp wide, so lane -1 is lane 31); the host picks HALO
//     from the stencil table, so first-order rows do 3 taps and 2 shuffles instead of 7 and 6;
//   * forcing: lane q owns term q of the sample's RandomForcing.  ONE sincosf per Runge-Kutta STEP; the later
//     stages rotate it by the per-term constant angle w c_s dt (angle addition, 4 FMAs).  The amplitude of
//     every spatial mode is a warp sum done by the integer adder of REDUX (redux.sync.add.s32): the terms are
//     scaled to 2^25 / (largest amplitude of the row) and rounded, so the sum is exact in fixed point and its
//     error (<= 2^-26 of the largest amplitude per term) is below that of a float32 summation -- 3
//     instructions per mode instead of 10 shuffles and adds.  Short rows keep their forcing basis in registers;
//   * the flux difference of the conservative forms needs the flux of the next point: one more shuffle.
// No shared memory, no block barriers.  The arithmetic of a right-hand side follows row_kernel operation by
// operation (same helpers: equation_point, weno_pair), so both satisfy the same tolerances.
//
// Reference: model.baseline_space_derivatives (model.py:59-112), polynomials.reconstruct
// (polynomials.py:280-303), equations.*.equation_of_motion, RandomForcing (equations.py:196-227),
// weno.py:43-123, integrate.odeint's Bogacki-Shampine steps (integrate.py:143-169).
#pragma once
#include "ddd1d_device.cuh"

namespace ddd1d {

constexpr int warp_rows_min_blocks(int ppl) { return ppl <= 2 ? 3 : ppl == 4 ? 2 : 1; }

// MM: compile-time bound on the number of forcing modes (3 is the reference's k_max)
// HALO: points the stencil table reaches on either side of a point (WENO: 3)
// EQ: the equation (EQ_* of ddd1d_device.cuh) when the instantiation is specialised for it, -1 = read P.eq.  The
//     Burgers forms the baselines are quoted on get their own instantiation: equation_point's switch is an
//     indirect branch per point and stage (17 % of the warp samples of the c1b launch before it was folded).
template <int PPL, bool WENO, int MM, int HALO, int EQ>
__global__ void __launch_bounds__(256, warp_rows_min_blocks(PPL))
warp_row_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W, const __grid_constant__ Tableau tab) {
  static_assert(!WENO || HALO == kHalo, "WENO5 reads three points on either side");
  constexpr int N = 32 * PPL;                         // == P.N
  constexpr int TAPS = 2 * HALO + 1;
  constexpr bool KEEP_BASIS = PPL * 2 * MM <= 24;     // the lane's forcing basis stays in registers
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const int eq = EQ >= 0 ? EQ : P.eq;
  const bool cons = eq_conservative(eq);
  const bool forced = eq_forced(eq) && P.P > 0;
  const int nstages = tab.stages;

  // window-form stencils of the derivative channels (ddd1d_set_stencils), offsets -HALO..+HALO, in registers
  float cf[kMaxD][TAPS];
#pragma unroll
  for (int d = 0; d < kMaxD; ++d)
#pragma unroll
    for (int j = 0; j < TAPS; ++j) cf[d][j] = d < P.D ? __ldg(P.blob + P.st_off + d * kWinPad + (kHalo - HALO) + j) : 0.f;

  for (int row = gwarp; row < W.batch; row += total_warps) {
    const int sample = W.sample_offset + row;
    // ---- per-row forcing constants (equations.py:196-219) ----
    const ForcingTerm fterm = load_forcing_term(P, sample, lane);
    float a_s = 0.f, a_c = 0.f, inv_scale = 0.f;        // term amplitude in fixed-point units, cosine copy signed by k
    // the spatial basis of this lane's points over the row's fixed-point scale:
    // [sine-amplitude factors 0..MM) | cosine-amplitude factors 0..MM)]
    float sbasis[KEEP_BASIS ? 2 * MM : 1][KEEP_BASIS ? PPL : 1];
    float rot_c[kMaxStages], rot_s[kMaxStages];         // cos / sin of w c_s dt
    unsigned int in_mode[MM];                           // all ones where this lane's term belongs to mode m + 1
#pragma unroll
    for (int m = 0; m < MM; ++m) in_mode[m] = 0u;
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) { rot_c[s] = 1.f; rot_s[s] = 0.f; }
    if (forced) {
      const bool on = lane < P.P;
      const unsigned int amax_bits = __reduce_max_sync(0xffffffffu, on ? __float_as_uint(fabsf(fterm.a)) : 0u);
      // 2^e >= largest |a| > 2^(e-1): terms scaled by 2^(25 - e) stay below 2^25, 32 of them below 2^30
      const int eb = (int)(amax_bits >> 23) + ((amax_bits & 0x7fffffu) ? 1 : 0);      // biased exponent of 2^e
      const int es = min(max(25 + 127 - (eb - 127), 1), 254);
      const float scale = amax_bits == 0u ? 1.f : __uint_as_float((unsigned int)es << 23);
      inv_scale = __uint_as_float((254u << 23) - __float_as_uint(scale));             // exact: powers of two
      a_s = on ? fterm.a * scale : 0.f;
      a_c = fterm.k < 0.f ? -a_s : a_s;
      const float ka = fabsf(fterm.k);
#pragma unroll
      for (int m = 0; m < MM; ++m) in_mode[m] = (on && m < P.M && ka == (float)(m + 1)) ? 0xffffffffu : 0u;
#pragma unroll
      for (int s = 1; s < kMaxStages; ++s)
        if (s < nstages) sincosf(fterm.w * (float)(tab.c[s] * W.dt), &rot_s[s], &rot_c[s]);
    }
    if constexpr (KEEP_BASIS) {
#pragma unroll
      for (int m = 0; m < 2 * MM; ++m)
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          const int mode = m < MM ? m : m - MM;
          const bool have = forced && mode < P.M;
          sbasis[m][i] = have ? __ldg(P.fbasis + (size_t)(m < MM ? mode : P.M + mode) * N + lane * PPL + i) * inv_scale : 0.f;
        }
    }

    float yh[PPL], yl[PPL];                              // solution = yh + yl, yh = float(yh + yl)
    float k[kMaxStages][PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if (W.u64) {
        const double v = W.u64[(size_t)row * N + lane * PPL + i];
        yh[i] = (float)v;
        yl[i] = (float)(v - (double)yh[i]);
      } else {
        yh[i] = __ldg(W.u + (size_t)row * N + lane * PPL + i);
        yl[i] = 0.f;
      }
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) k[s][i] = 0.f;
    }
    int first_bad = -1;
    int save_idx = 0;
    int until_save = W.save_every;

    for (int step = 0; step < W.nsteps; ++step) {
      float sn0 = 0.f, cs0 = 0.f;
      if (forced) {
        const float t = (float)(W.t0 + (double)step * W.dt);
        sincosf(fmaf(fterm.w, t, fterm.phi), &sn0, &cs0);
      }
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) {
        if (s >= nstages) break;
        // ---- stage row y + dt * sum_j a[s][j] k_j, rounded to float32 (integrate.py:57-60,71): the increment in
        //      float32, added low part first ----
        float e[PPL + 2 * HALO];             // e[HALO + i] = u[point i of this lane]
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          float inc = 0.f;
#pragma unroll
          for (int j = 0; j < kMaxStages; ++j)
            if (j < s) inc = fmaf(W.adt[s][j], k[j][i], inc);
          e[HALO + i] = s == 0 ? yh[i] : yh[i] + (yl[i] + inc);
        }
        // ---- periodic halo from the neighbouring lanes ----
#pragma unroll
        for (int h = 1; h <= HALO; ++h) {
          const int dl = (h + PPL - 1) / PPL;                  // lanes to the left
          const int il = (PPL - (h % PPL)) % PPL;              // its point index
          e[HALO - h] = __shfl_sync(0xffffffffu, e[HALO + il], (lane - dl) & 31);
          const int dr = (PPL - 1 + h) / PPL;                  // lanes to the right
          const int ir = (PPL - 1 + h) % PPL;
          e[HALO + PPL - 1 + h] = __shfl_sync(0xffffffffu, e[HALO + ir], (lane + dr) & 31);
        }
        // ---- forcing amplitudes of this stage: rotate the step's sine / cosine, fixed-point warp sums ----
        float amp[2 * MM];
        if (forced) {
          const float sn = s == 0 ? sn0 : fmaf(sn0, rot_c[s], cs0 * rot_s[s]);
          const float cs = s == 0 ? cs0 : fmaf(cs0, rot_c[s], -(sn0 * rot_s[s]));
          const int is = __float2int_rn(a_s * sn), ic = __float2int_rn(a_c * cs);
          // (modes beyond P.M have an empty mask: their sums are zero, no branch needed)
#pragma unroll
          for (int m = 0; m < MM; ++m) {
            amp[m] = (float)__reduce_add_sync(0xffffffffu, (int)((unsigned int)is & in_mode[m]));
            amp[MM + m] = (float)__reduce_add_sync(0xffffffffu, (int)((unsigned int)ic & in_mode[m]));
            if constexpr (!KEEP_BASIS) { amp[m] *= inv_scale; amp[MM + m] *= inv_scale; }
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
            if (!(WENO && d < 2)) {          // (rows beyond P.D are zero; WENO overwrites channels 0 and 1 below)
#pragma unroll
              for (int j = 0; j < TAPS; ++j) acc = fmaf(cf[d][j], e[i + j], acc);
            }
            dv[d] = acc;
          }
          if (WENO) {                        // u_minus / u_plus replaced by WENO5 (integrate.py:134-138)
            float um, up;
            weno_pair<float>(&e[HALO + i], um, up);
            dv[0] = um;
            dv[1] = up;
          }
          r[i] = equation_point(eq, e[HALO + i], dv, P.eta);
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
            float f = 0.f;
#pragma unroll
            for (int m = 0; m < MM; ++m) {
              float bs, bc;
              if constexpr (KEEP_BASIS) {
                bs = sbasis[m][i];
                bc = sbasis[MM + m][i];
              } else {
                bs = m < P.M ? __ldg(P.fbasis + (size_t)m * N + lane * PPL + i) : 0.f;
                bc = m < P.M ? __ldg(P.fbasis + (size_t)(P.M + m) * N + lane * PPL + i) : 0.f;
              }
              f = fmaf(amp[m], bs, f);
              f = fmaf(amp[MM + m], bc, f);
            }
            r[i] = __fadd_rn(r[i], f);
          }
        }
#pragma unroll
        for (int i = 0; i < PPL; ++i) k[s][i] = r[i];
      }
      // ---- end of the step: y += dt * sum b k in float-float (the float64 sum of the reference to ~2^-48):
      //      exact products of the float-float constants dt * b_j, highs summed with TwoSum ----
      const bool save = --until_save == 0;
      float* snap = save ? W.snaps + ((size_t)save_idx * W.batch + row) * N + lane * PPL : nullptr;
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        float ih = W.bdt_hi[0] * k[0][i];
        float il = fmaf(W.bdt_hi[0], k[0][i], -ih) + W.bdt_lo[0] * k[0][i];
#pragma unroll
        for (int j = 1; j < kMaxStages; ++j) {
          if (j >= nstages) break;
          const float ph = W.bdt_hi[j] * k[j][i];
          const float pl = fmaf(W.bdt_hi[j], k[j][i], -ph) + W.bdt_lo[j] * k[j][i];
          float sh, er;
          warp_two_sum(ih, ph, sh, er);
          ih = sh;
          il += er + pl;
        }
        float nh, er;
        warp_two_sum(yh[i], ih, nh, er);
        const float nl = yl[i] + (er + il);
        const float y = nh + nl;                       // renormalise: y = float(yh + yl)
        yl[i] = W.state_f32 ? 0.f : nl - (y - nh);     // (float32 carry: tf odeint_fixed, model.py:138-159)
        yh[i] = y;
        if (first_bad < 0 && !isfinite(y)) first_bad = step;
        if (save) snap[i] = y;
      }
      if (save) { ++save_idx; until_save = W.save_every; }
    }
    if (W.first_bad) {
      unsigned int key = first_bad < 0 ? 0xffffffffu : (unsigned int)first_bad;
      key = __reduce_min_sync(0xffffffffu, key);
      if (lane == 0) W.first_bad[row] = key == 0xffffffffu ? -1 : (int)key;
    }
  }
}

}  // namespace ddd1d
