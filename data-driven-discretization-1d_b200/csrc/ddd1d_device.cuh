// Device code of the fused 1-D PDE row integrator (sm_100a).
//
// One CTA owns one solution row u[N] at a time and keeps it on chip for the whole
// call: the periodic halo lives in shared memory, the conv-net filters and the
// polynomial-accuracy tables are staged once per CTA with a TMA bulk copy
// (cp.async.bulk -> UBLKCP), activations ping-pong between two shared buffers,
// and the Runge-Kutta state (float64 y, float32 stage slopes) never leaves the
// SM between snapshots.  Reference semantics are cited per function
// (paths relative to /root/reference/pde_superresolution/).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ddd1d {

constexpr int kMaxD = 4;        // derivative channels (DDD1D_MAX_DERIVATIVES)
constexpr int kWin = 7;         // stencil window of the kernels' common case, offsets -3..+3
constexpr int kWinPad = 8;      // row pitch of the 7-slot window tables
constexpr int kWinWide = 11;    // the window of the C ABI (DDD1D_WINDOW), offsets -5..+5: 9- and 10-point coefficient grids
constexpr int kWinWidePad = 12; // row pitch of 11-slot tables
constexpr int kHalo = 3;        // halo of a 7-slot window
constexpr int kRowHalo = 5;     // halo of the stage row in row_kernel's shared memory (covers the 11-slot window)
constexpr int kMaxLayers = 6;
constexpr int kMaxStages = 4;
constexpr int kMaxModes = 8;
constexpr int kMaxForcing = 32;  // forcing terms per sample (the reference uses 20 / 10)

enum Op { OP_RHS = 0, OP_COEF = 1, OP_DERIV = 2, OP_INTEGRATE = 3, OP_ADAPTIVE = 4 };
enum Mode { MODE_STENCIL = 0, MODE_LEARNED = 1, MODE_WENO = 2 };
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU6 = 2, ACT_TANH = 3, ACT_SOFTPLUS = 4, ACT_ELU = 5 };
enum Proj { PROJ_NULLSPACE = 0, PROJ_RAW = 1, PROJ_RAW_UNBIASED = 2, PROJ_DERIVS = 3, PROJ_TIME = 4, PROJ_FLUX = 5 };
// eq code = equation * 3 + variant
enum Eq {
  EQ_BURGERS = 0, EQ_BURGERS_CONS = 1, EQ_BURGERS_GOD = 2,
  EQ_KDV = 3, EQ_KDV_CONS = 4, EQ_KDV_GOD = 5,
  EQ_KS = 6, EQ_KS_CONS = 7, EQ_KS_GOD = 8
};

struct LayerPlan {
  int cin, cout, cout_pad;  // cout_pad: multiple of 8, zero-filled
  int w_off, b_off;         // float offsets into the constant blob: W[(ci*K+k)*cout_pad+co], b[cout_pad]
  int act;
  int cg, pbt;              // fast path tile plan: channels per warp task, 128-position blocks per task
};

struct Params {
  int eq, mode, N, D, S, wshift, weno_real;
  int win;                  // 7, or 11 when a coefficient grid reaches beyond offsets +-3 (learned mode, FFMA engine only)
  float sigma, eta, inv_dx;
  double inv_dx_d;          // 1 / dx in float64 (float64 WENO path)
  // conv net
  int nlayers, K, kleft, fast_conv, pitch;
  LayerPlan layer[kMaxLayers];
  int C, projection;
  int cstart[kMaxD + 1];
  int ns_off, st_off;       // blob offsets: nullspace [C][8]; stencil / bias window [kMaxD][8]
  const float* blob;
  int blob_floats;          // multiple of 4
  // forcing (equations.py:196-219), decomposed into 2M spatial modes
  int P, M, fcap;
  const float* fparams;     // [fcap][4][P]: a, omega, phi, signed k
  const float* fbasis;      // [2M][N]: resampled cos(2 pi m x/L), sin(2 pi m x/L), m = 1..M
  const double* fparams64;  // float64 copies for the float64 WENO path (integrate.py:124-140)
  const double* fbasis64;
  // shared memory carve-up (byte offsets)
  int off_bar, off_blob, off_ust, off_ydbl, off_ynew, off_red, off_k, off_flux, off_fs, off_act0, off_act1;
  int off_ustd, off_kd, off_fluxd, off_fsd;   // float64 twins (float64 WENO path only)
  int smem_bytes;
  int use_bulk_copy;
};

struct Work {
  int op, batch, sample_offset;
  int integrating;     // op == OP_INTEGRATE as its own field (tensor engine: a test of `op` in the step loop joins the
                       // per-call ops' tests in one jump table; this one is a plain predicate)
  const float* u;      // [batch][N] (or null when u64 is given)
  const double* u64;
  float* out;          // rhs [batch][N] | coef [batch][N][D][S] | deriv [batch][N][D]
  double* out64;
  double t0, dt;
  int nsteps, save_every, scheme;
  // dt * tableau as float32 / float-float constants, filled by the host per launch (the tensor engine carries its
  // float64-equivalent state as an unevaluated float pair: conversions to and from float64 run at 16 lanes/clk/SM)
  float adt[kMaxStages][kMaxStages];     // float(dt * a[s][j])
  float bdt_hi[kMaxStages], bdt_lo[kMaxStages];     // dt * b[j] = hi + lo
  int state_f32;       // carry the solution in float32 between steps (tf odeint_fixed, model.py:138-159) instead of float64 (SciPy)
  float* snaps;        // [nsteps/save_every][batch][N]
  int* first_bad;      // [batch] or null
  // OP_ADAPTIVE (scipy RK23 twin): output times, tolerances, float64 record
  const double* times; // device [ntimes], increasing; times[0] is the start time
  int ntimes;
  double rtol, atol, max_step;
  double* y_out;       // [ntimes][batch][N], NaN where the solver gave up
  int* nfev;           // [batch]
  int* status;         // [batch]: 0 reached the end, -1 step size underflow
};

// a + b = s + e exactly (Knuth's TwoSum): the float-pair state of the fused integrators
__device__ __forceinline__ void warp_two_sum(float a, float b, float& s, float& e) {
  s = a + b;
  const float bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}

// The one dynamic shared-memory window of every kernel in this library.  Device functions that
// are not inlined re-derive their pointers from this symbol so loads stay LDS/STS.
extern __shared__ __align__(128) unsigned char dyn_smem[];

// ---------------------------------------------------------------------------------
// Runge-Kutta tableaus.  RK3 is Bogacki-Shampine as in scipy/integrate/_ivp/rk.py
// (the scheme integrate.odeint runs, integrate.py:154-155); midpoint is what
// tf.contrib.integrate.odeint_fixed(method='midpoint') applies (model.py:156-157).
// ---------------------------------------------------------------------------------
struct Tableau {
  int stages;
  double c[kMaxStages];
  double a[kMaxStages][kMaxStages];
  double b[kMaxStages];
};

__host__ __device__ __forceinline__ Tableau make_tableau(int scheme) {
  Tableau t;
#pragma unroll
  for (int i = 0; i < kMaxStages; ++i) {
    t.c[i] = 0.0; t.b[i] = 0.0;
#pragma unroll
    for (int j = 0; j < kMaxStages; ++j) t.a[i][j] = 0.0;
  }
  if (scheme == 0) {          // RK3 Bogacki-Shampine
    t.stages = 3;
    t.c[1] = 0.5; t.c[2] = 0.75;
    t.a[1][0] = 0.5; t.a[2][1] = 0.75;
    t.b[0] = 2.0 / 9.0; t.b[1] = 1.0 / 3.0; t.b[2] = 4.0 / 9.0;
  } else if (scheme == 1) {   // explicit midpoint
    t.stages = 2;
    t.c[1] = 0.5; t.a[1][0] = 0.5; t.b[1] = 1.0;
  } else if (scheme == 2) {   // forward Euler
    t.stages = 1;
    t.b[0] = 1.0;
  } else {                    // classic RK4
    t.stages = 4;
    t.c[1] = 0.5; t.c[2] = 0.5; t.c[3] = 1.0;
    t.a[1][0] = 0.5; t.a[2][1] = 0.5; t.a[3][2] = 1.0;
    t.b[0] = 1.0 / 6.0; t.b[1] = 1.0 / 3.0; t.b[2] = 1.0 / 3.0; t.b[3] = 1.0 / 6.0;
  }
  return t;
}

// ---------------------------------------------------------------------------------
// TMA bulk copy + mbarrier helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// ---------------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int wrap(int x, int n) {
  x %= n;
  return x < 0 ? x + n : x;
}

static __device__ __noinline__ float activate_rare(float x, int act) {
  switch (act) {   // model.py:411-417
    case ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
    case ACT_TANH: return tanhf(x);
    case ACT_SOFTPLUS: return (x > 0.f ? x : 0.f) + log1pf(expf(-fabsf(x)));
    case ACT_ELU: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}

__device__ __forceinline__ float activate(float x, int act) {
  if (act == ACT_NONE) return x;
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  return activate_rare(x, act);   // out of line: keeps the conv epilogues small
}

// equations.py:341-349
template <typename T>
__device__ __forceinline__ T godunov_flux(T um, T up) {
  T m2 = um * um, p2 = up * up;
  T lo = m2 < p2 ? m2 : p2, hi = m2 < p2 ? p2 : m2;
  return T(0.5) * (um <= up ? lo : hi);
}

// weno.py:43-123 evaluated for one output point.  v points at u[p] inside a row
// with >= 3 valid halo points on both sides.  Returns
//   um = roll(reconstruct_left(u), 1)[p]  = left reconstruction at cell p-1 (u[p-3..p+1])
//   up = roll(reconstruct_right(u), 1)[p] = right reconstruction at cell p-1 (u[p-2..p+2])
// as consumed by integrate.py:137-138 / model.py:83-87.
template <typename T>
__device__ __forceinline__ void weno_beta(T a, T b, T c, T d, T e, T& b0, T& b1, T& b2) {
  // smoothness indicators centred on c  (weno.py:46-57)
  T q0 = a - T(4) * b + T(3) * c, r0 = a - T(2) * b + c;
  T q1 = b - d, r1 = b - T(2) * c + d;
  T q2 = T(3) * c - T(4) * d + e, r2 = c - T(2) * d + e;
  b0 = T(0.25) * (q0 * q0) + T(13.0 / 12.0) * (r0 * r0);
  b1 = T(0.25) * (q1 * q1) + T(13.0 / 12.0) * (r1 * r1);
  b2 = T(0.25) * (q2 * q2) + T(13.0 / 12.0) * (r2 * r2);
}

template <typename T>
__device__ __forceinline__ void weno_pair(const T* v, T& um, T& up) {
  const T eps = T(1e-6);
  {  // left-biased, centre i = p-1: points u[i-2..i+2] = v[-3..1]   (weno.py:76-97)
    T a = v[-3], b = v[-2], c = v[-1], d = v[0], e = v[1];
    T b0, b1, b2;
    weno_beta(a, b, c, d, e, b0, b1, b2);
    T a0 = T(0.1) / ((eps + b0) * (eps + b0));
    T a1 = T(0.6) / ((eps + b1) * (eps + b1));
    T a2 = T(0.3) / ((eps + b2) * (eps + b2));
    T s = a0 + a1 + a2;
    T w0 = a0 / s, w1 = a1 / s, w2 = a2 / s;
    um = (w0 / T(3)) * a + (-(T(7) * w0 + w1) / T(6)) * b +
         ((T(11) * w0 + T(5) * w1 + T(2) * w2) / T(6)) * c + ((T(2) * w1 + T(5) * w2) / T(6)) * d +
         (-w2 / T(6)) * e;
  }
  {  // right-biased at cell i = p-1: indicators centred on i+1 = p (weno.py:100-123)
    T a = v[-2], b = v[-1], c = v[0], d = v[1], e = v[2];
    T b0, b1, b2;
    weno_beta(a, b, c, d, e, b0, b1, b2);
    T a0 = T(0.3) / ((eps + b0) * (eps + b0));
    T a1 = T(0.6) / ((eps + b1) * (eps + b1));
    T a2 = T(0.1) / ((eps + b2) * (eps + b2));
    T s = a0 + a1 + a2;
    T w2 = a0 / s, w1 = a1 / s, w0 = a2 / s;   // omega2, omega1, omega0 of weno.py:106-108
    up = (-w2 / T(6)) * a + ((T(5) * w2 + T(2) * w1) / T(6)) * b +
         ((T(2) * w2 + T(5) * w1 + T(11) * w0) / T(6)) * c + (-(w1 + T(7) * w0) / T(6)) * d +
         (w0 / T(3)) * e;
  }
}

// ---------------------------------------------------------------------------------
// Periodic conv layer, fast path (K == 5, N % 4 == 0):
//   out[co][x] = act(b[co] + sum_{k,ci} W[k][ci][co] * in[ci][(x + k - 2) mod N])
// (layers.py:103-137 with center=True: 2 wrapped points each side, layers.py:76-79).
// Buffers are [channel][pitch] with position x stored at index x + 2.
// A warp task = CG output channels x PBT blocks of 128 positions; a lane owns 4
// consecutive positions per block, so activations load as conflict-free LDS.128
// and the weights as warp-broadcast LDS.128.
// ---------------------------------------------------------------------------------
template <int CG, int PBT>
__device__ __forceinline__ void conv5_task(const float* __restrict__ in, float* __restrict__ out,
                                           const float* __restrict__ wgt, const float* __restrict__ bias,
                                           int cin, int cout_pad, int cout, int co0, int pb0, int N,
                                           int pitch, int act, int lane) {
  float acc[PBT][4][CG];
  int p0[PBT];
  bool live[PBT];
#pragma unroll
  for (int b = 0; b < PBT; ++b) {
    p0[b] = (pb0 + b) * 128 + lane * 4;
    live[b] = p0[b] < N;
    if (!live[b]) p0[b] = 0;   // keep loads in bounds; results discarded
#pragma unroll
    for (int c = 0; c < CG; ++c) {
      float bv = bias[co0 + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[b][j][c] = bv;
    }
  }
  const float* wrow = wgt + co0;
  for (int ci = 0; ci < cin; ++ci) {
    float a[PBT][8];
#pragma unroll
    for (int b = 0; b < PBT; ++b) {
      const float4* src = reinterpret_cast<const float4*>(in + ci * pitch + p0[b]);
      float4 lo = src[0], hi = src[1];
      a[b][0] = lo.x; a[b][1] = lo.y; a[b][2] = lo.z; a[b][3] = lo.w;
      a[b][4] = hi.x; a[b][5] = hi.y; a[b][6] = hi.z; a[b][7] = hi.w;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float w[CG];
      const float4* wsrc = reinterpret_cast<const float4*>(wrow + (ci * 5 + k) * cout_pad);
#pragma unroll
      for (int q = 0; q < CG / 4; ++q) {
        float4 t = wsrc[q];
        w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int b = 0; b < PBT; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < CG; ++c) acc[b][j][c] = fmaf(a[b][j + k], w[c], acc[b][j][c]);
    }
  }
#pragma unroll
  for (int b = 0; b < PBT; ++b) {
    if (!live[b]) continue;
#pragma unroll
    for (int c = 0; c < CG; ++c) {
      if (co0 + c >= cout) continue;
      float r0 = activate(acc[b][0][c], act), r1 = activate(acc[b][1][c], act);
      float r2 = activate(acc[b][2][c], act), r3 = activate(acc[b][3][c], act);
      float* dst = out + (co0 + c) * pitch + p0[b] + 2;
      reinterpret_cast<float2*>(dst)[0] = make_float2(r0, r1);
      reinterpret_cast<float2*>(dst)[1] = make_float2(r2, r3);
      if (p0[b] == 0) {            // positions 0,1 also feed the right halo
        float* h = out + (co0 + c) * pitch + N + 2;
        h[0] = r0; h[1] = r1;
      }
      if (p0[b] + 4 == N) {        // positions N-2,N-1 also feed the left halo
        float* h = out + (co0 + c) * pitch;
        h[0] = r2; h[1] = r3;
      }
    }
  }
}

template <int CG, int PBT>
__device__ __forceinline__ void conv5_layer(const LayerPlan& L, const float* blob, const float* in,
                                            float* out, int N, int pitch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ncg = (L.cout + CG - 1) / CG;
  const int npb = (N + 127) / 128;
  const int npt = (npb + PBT - 1) / PBT;
  const int ntasks = ncg * npt;
  for (int task = warp; task < ntasks; task += nwarps) {
    int cgi = task % ncg, pt = task / ncg;
    conv5_task<CG, PBT>(in, out, blob + L.w_off, blob + L.b_off, L.cin, L.cout_pad, L.cout, cgi * CG,
                        pt * PBT, N, pitch, L.act, lane);
  }
}

// Generic periodic conv (any K, any N): one output element per thread iteration.
__device__ __forceinline__ void conv_generic_layer(const LayerPlan& L, const float* blob, const float* in,
                                                   float* out, int N, int pitch, int K, int kleft) {
  const float* wgt = blob + L.w_off;
  const float* bias = blob + L.b_off;
  for (int idx = threadIdx.x; idx < N * L.cout; idx += blockDim.x) {
    int p = idx % N, co = idx / N;
    float acc = bias[co];
    for (int ci = 0; ci < L.cin; ++ci) {
      const float* row = in + ci * pitch + p;
      for (int k = 0; k < K; ++k) acc = fmaf(row[k], wgt[(ci * K + k) * L.cout_pad + co], acc);
    }
    out[co * pitch + p + kleft] = activate(acc, L.act);
  }
}

// fill the wrapped halo of a [channels][pitch] buffer (generic path)
__device__ __forceinline__ void fill_halo(float* buf, int channels, int N, int pitch, int K, int kleft) {
  const int kright = K - 1 - kleft;
  const int per = kleft + kright;
  for (int idx = threadIdx.x; idx < channels * per; idx += blockDim.x) {
    int c = idx / per, h = idx % per;
    float* row = buf + c * pitch;
    if (h < kleft) {
      row[h] = row[kleft + wrap(h - kleft, N)];
    } else {
      int r = h - kleft;
      row[kleft + N + r] = row[kleft + wrap(r, N)];
    }
  }
}

// ---------------------------------------------------------------------------------
// One right-hand-side evaluation for the row resident in shared memory.
// ---------------------------------------------------------------------------------
struct Smem {
  uint64_t* bar;
  float* blob;
  float* ust;     // raw stage row, index x + kRowHalo, halo filled
  double* ydbl;   // float64 state
  double* ynew;   // float64 trial state (adaptive stepping)
  double* red;    // reduction scratch [34]
  float* k;       // [kMaxStages][N] stage slopes
  float* flux;    // [N + 1]
  float* fs;      // [2*kMaxModes] mode amplitudes | [2*kMaxForcing] per-term sin/cos parts | [kMaxForcing] |k|
  float* act0;
  float* act1;
  double* ustd;   // float64 stage row (halo 3), float64 WENO path
  double* kd;     // float64 stage slopes
  double* fluxd;
  double* fsd;    // float64 forcing scratch, same layout as fs
};

__device__ __forceinline__ Smem carve(const Params& P, unsigned char* base) {
  Smem s;
  s.bar = reinterpret_cast<uint64_t*>(base + P.off_bar);
  s.blob = reinterpret_cast<float*>(base + P.off_blob);
  s.ust = reinterpret_cast<float*>(base + P.off_ust);
  s.ydbl = reinterpret_cast<double*>(base + P.off_ydbl);
  s.ynew = reinterpret_cast<double*>(base + P.off_ynew);
  s.red = reinterpret_cast<double*>(base + P.off_red);
  s.k = reinterpret_cast<float*>(base + P.off_k);
  s.flux = reinterpret_cast<float*>(base + P.off_flux);
  s.fs = reinterpret_cast<float*>(base + P.off_fs);
  s.act0 = reinterpret_cast<float*>(base + P.off_act0);
  s.act1 = reinterpret_cast<float*>(base + P.off_act1);
  s.ustd = reinterpret_cast<double*>(base + P.off_ustd);
  s.kd = reinterpret_cast<double*>(base + P.off_kd);
  s.fluxd = reinterpret_cast<double*>(base + P.off_fluxd);
  s.fsd = reinterpret_cast<double*>(base + P.off_fsd);
  return s;
}

template <typename KT> __device__ __forceinline__ KT* slopes(const Smem& S);
template <> __device__ __forceinline__ float* slopes<float>(const Smem& S) { return S.k; }
template <> __device__ __forceinline__ double* slopes<double>(const Smem& S) { return S.kd; }

// stage row <- float(value) with its periodic halo; value(p) is evaluated per thread.  In learned
// mode the same pass writes the net input u / standard_deviation (model.py:450-451) as channel 0 of
// act0, so the RHS evaluation starts at the first conv layer without another barrier.
template <typename F>
__device__ __forceinline__ void write_stage_row(const Params& P, const Smem& S, F value) {
  const int N = P.N;
  const bool learned = P.mode == MODE_LEARNED;
  const int kl = P.kleft, kr = P.K - 1 - P.kleft;
  const bool f64 = P.mode == MODE_WENO && P.weno_real == 1;
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    const double vd = value(p);
    const float v = (float)vd;         // the float32 placeholder feed (integrate.py:57-60,71)
    S.ust[p + kRowHalo] = v;
    // wrapped copies (also correct when N < kRowHalo: every halo slot is assigned by some p)
    for (int q = p - N; q >= -kRowHalo; q -= N) S.ust[q + kRowHalo] = v;
    for (int q = p + N; q < N + kRowHalo; q += N) S.ust[q + kRowHalo] = v;
    if (f64) {                         // the float64 row the NumPy WENO code sees (integrate.py:137-138)
      S.ustd[p + kRowHalo] = vd;
      for (int q = p - N; q >= -kRowHalo; q -= N) S.ustd[q + kRowHalo] = vd;
      for (int q = p + N; q < N + kRowHalo; q += N) S.ustd[q + kRowHalo] = vd;
    }
    if (learned) {
      const float vn = __fdiv_rn(v, P.sigma);
      S.act0[p + kl] = vn;
      for (int q = p - N; q >= -kl; q -= N) S.act0[q + kl] = vn;
      for (int q = p + N; q < N + kr; q += N) S.act0[q + kl] = vn;
    }
  }
}

// Forcing at time t for one sample:
//   F(x,t) = sum_p a_p sin(w_p t + 2 pi k_p x / L + phi_p)           (equations.py:214-219)
//          = sum_m [sum_{|k_p|=m} a_p sin(w_p t + phi_p)] cos(2 pi m x/L)
//                + [sum_{|k_p|=m} sgn(k_p) a_p cos(w_p t + phi_p)] sin(2 pi m x/L)
// Term p lives in the registers of thread p for the whole row (ForcingTerm); each RHS evaluation
// costs one sincosf on P threads (forcing_terms), a barrier the caller already has, and a
// P-term sum on 2M threads (forcing_reduce).
struct ForcingTerm { float a, w, phi, k; };

__device__ __forceinline__ ForcingTerm load_forcing_term(const Params& P, int sample, int q) {
  ForcingTerm f = {0.f, 0.f, 0.f, 0.f};
  if (P.P > 0 && q < P.P) {
    const float* fp = P.fparams + (size_t)sample * 4 * P.P;
    f.a = __ldg(fp + q); f.w = __ldg(fp + P.P + q); f.phi = __ldg(fp + 2 * P.P + q); f.k = __ldg(fp + 3 * P.P + q);
  }
  return f;
}

// fs layout: [0, 2*kMaxModes) amplitudes, then 2*kMaxForcing per-term parts, then kMaxForcing |k|
__device__ __forceinline__ void forcing_terms(const Params& P, float* fs, const ForcingTerm& f, int q, float t) {
  if (q < P.P) {
    float sn, cs;
    sincosf(fmaf(f.w, t, f.phi), &sn, &cs);
    float* parts = fs + 2 * kMaxModes;
    parts[q] = f.a * sn;
    parts[kMaxForcing + q] = (f.k < 0.f ? -f.a : f.a) * cs;
    parts[2 * kMaxForcing + q] = fabsf(f.k);
  }
}

__device__ __forceinline__ void forcing_reduce(const Params& P, float* fs, int j) {
  if (j < 2 * P.M) {
    const float* parts = fs + 2 * kMaxModes;
    const float m = (float)((j < P.M ? j : j - P.M) + 1);
    const float* src = parts + (j < P.M ? 0 : kMaxForcing);
    float acc = 0.f;
    for (int q = 0; q < P.P; ++q)
      if (parts[2 * kMaxForcing + q] == m) acc += src[q];
    fs[j] = acc;
  }
}

// Warp-wide version: lane q holds term q (terms beyond 32 in further rounds); the amplitude of mode m is the
// warp sum of the terms with |k| == m.  Writes fs[0..M) (sine) and fs[M..2M) (cosine).
__device__ __forceinline__ void forcing_amplitudes(const Params& P, float* fs, int sample, float t, int lane) {
  float ps[kMaxModes], pc[kMaxModes];
#pragma unroll
  for (int m = 0; m < kMaxModes; ++m) ps[m] = pc[m] = 0.f;
  for (int q = lane; q < P.P; q += 32) {
    const ForcingTerm f = load_forcing_term(P, sample, q);
    float sn, cs;
    sincosf(fmaf(f.w, t, f.phi), &sn, &cs);
    const float a_sin = f.a * sn, a_cos = (f.k < 0.f ? -f.a : f.a) * cs, ka = fabsf(f.k);
#pragma unroll
    for (int m = 0; m < kMaxModes; ++m)
      if (ka == (float)(m + 1)) { ps[m] += a_sin; pc[m] += a_cos; }
  }
#pragma unroll
  for (int m = 0; m < kMaxModes; ++m) {
    if (m >= P.M) break;
    float a = ps[m], b = pc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { fs[m] = a; fs[P.M + m] = b; }
  }
}

__device__ __forceinline__ float forcing_at(const Params& P, const Smem& S, int p) {
  float f = 0.f;
  for (int m = 0; m < P.M; ++m) {
    f = fmaf(S.fs[m], __ldg(P.fbasis + (size_t)m * P.N + p), f);
    f = fmaf(S.fs[P.M + m], __ldg(P.fbasis + (size_t)(P.M + m) * P.N + p), f);
  }
  return f;
}

// Spatial derivatives at point p.  dv[d] for d < D; optionally exports the
// coefficient rows (OP_COEF).
// WIN = 7 (offsets -3..+3, the common case) or 11 (hparams.coefficient_grid_min_size = 9: 9 centred or 10 staggered points).
template <int MODE, int WIN = kWin>
__device__ __forceinline__ void point_derivatives(const Params& P, const Smem& S, const float* net, int p,
                                                  float (&dv)[kMaxD], float* gcoef) {
  constexpr int kWin = WIN, kWinPad = WIN == 7 ? ddd1d::kWinPad : kWinWidePad;      // (shadow the 7-slot constants)
  const float* up = S.ust + p + (kRowHalo - WIN / 2);   // up[j] = u[p + j - WIN / 2]
  float u7[kWin];
#pragma unroll
  for (int j = 0; j < kWin; ++j) u7[j] = up[j];
  const float* st = S.blob + P.st_off;
#pragma unroll
  for (int d = 0; d < kMaxD; ++d) {
    dv[d] = 0.f;
    if (d >= P.D) continue;
    if (MODE == MODE_WENO && d < 2) continue;      // u_minus / u_plus come from WENO5 below, not from stencils
    float cf[kWin];
    if (MODE == MODE_LEARNED && P.projection >= PROJ_DERIVS) {
      // model_target='space_derivatives': the net's channels ARE the derivatives (model.py:571-576)
      dv[d] = (P.projection == PROJ_DERIVS) ? net[d * P.pitch + p + P.kleft] : 0.f;
      continue;
    }
    if (MODE == MODE_LEARNED) {
      if (P.projection == PROJ_NULLSPACE) {
        // coef = bias + z @ nullspace (polynomials.py:266-277), window-aligned on the host
#pragma unroll
        for (int j = 0; j < kWin; ++j) cf[j] = 0.f;
        const float* ns = S.blob + P.ns_off;
        for (int c = P.cstart[d]; c < P.cstart[d + 1]; ++c) {
          float z = net[c * P.pitch + p + P.kleft];
#pragma unroll
          for (int j = 0; j < kWin; ++j) cf[j] = fmaf(z, ns[c * kWinPad + j], cf[j]);
        }
#pragma unroll
        for (int j = 0; j < kWin; ++j) cf[j] += st[d * kWinPad + j];
      } else {
        // polynomial_accuracy_order == 0: raw reshape [D][S] (model.py:460-475)
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < kWin; ++j) {
          int i = j - P.wshift;
          cf[j] = (i >= 0 && i < P.S) ? net[(d * P.S + i) * P.pitch + p + P.kleft] : 0.f;
          mean += cf[j];
        }
        if (P.projection == PROJ_RAW_UNBIASED) {
          mean /= (float)P.S;
#pragma unroll
          for (int j = 0; j < kWin; ++j) {
            int i = j - P.wshift;
            if (i >= 0 && i < P.S) cf[j] -= mean;
          }
        }
      }
      if (gcoef) {
#pragma unroll
        for (int j = 0; j < kWin; ++j) {
          int i = j - P.wshift;
          if (i >= 0 && i < P.S) gcoef[d * P.S + i] = cf[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kWin; ++j) cf[j] = st[d * kWinPad + j];
    }
    // einsum('bxdi,bxi->bxd') on the un-normalised inputs (model.py:536-548)
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kWin; ++j) s = fmaf(cf[j], u7[j], s);
    dv[d] = s;
  }
  if (MODE == MODE_WENO) {
    // u_minus / u_plus replaced by WENO5 (integrate.py:134-138)
    float um, upv;
    weno_pair<float>(S.ust + p + kRowHalo, um, upv);
    dv[0] = um;
    dv[1] = upv;
  }
}

// equation_of_motion (equations.py:269-274,331-338,360-370,410-415,450-457,468-478,
// 518-524,559-567,576-587).  Returns y_t for the non-conservative forms and the
// flux at x - 1/2 for the conservative ones.  Products are kept unfused so the
// float32 rounding follows the reference's op-by-op graph.
__device__ __forceinline__ float equation_point(int eq, float u, const float (&dv)[kMaxD], float eta) {
  switch (eq) {
    case EQ_BURGERS: return __fsub_rn(__fmul_rn(eta, dv[1]), __fmul_rn(u, dv[0]));
    case EQ_KDV: return __fsub_rn(__fmul_rn(__fmul_rn(-6.f, u), dv[0]), dv[1]);
    case EQ_KS: return __fsub_rn(__fsub_rn(__fmul_rn(-u, dv[0]), dv[2]), dv[1]);
    case EQ_BURGERS_CONS: return __fsub_rn(__fmul_rn(0.5f, __fmul_rn(dv[0], dv[0])), __fmul_rn(eta, dv[1]));
    case EQ_KDV_CONS: return __fadd_rn(__fmul_rn(3.f, __fmul_rn(dv[0], dv[0])), dv[1]);
    case EQ_KS_CONS:
      return __fadd_rn(__fadd_rn(__fmul_rn(0.5f, __fmul_rn(dv[0], dv[0])), dv[2]), dv[1]);
    case EQ_BURGERS_GOD: return __fsub_rn(godunov_flux<float>(dv[0], dv[1]), __fmul_rn(eta, dv[2]));
    case EQ_KDV_GOD: return __fadd_rn(__fmul_rn(6.f, godunov_flux<float>(dv[0], dv[1])), dv[2]);
    default: /* EQ_KS_GOD */
      return __fadd_rn(__fadd_rn(dv[3], dv[2]), godunov_flux<float>(dv[0], dv[1]));
  }
}

__device__ __forceinline__ bool eq_conservative(int eq) { return (eq % 3) != 0; }
__device__ __forceinline__ bool eq_forced(int eq) { return eq < 3; }   // Burgers family, equations.py:276-277

// Cooperative RHS evaluation.  Preconditions: S.ust holds the stage row (halo
// filled) and a __syncthreads() has made it visible.  Postcondition: S.k[kslot][p],
// p < N, holds dy/dt (float32) and is visible to all threads.
template <int MODE>
__device__ __noinline__ void row_rhs(const Params& P, const ForcingTerm& fterm, float t, int kslot, int op,
                                     float* gout_row) {
  const Smem S = carve(P, dyn_smem);
  float* kout = S.k + kslot * P.N;
  const int N = P.N;
  const float* net = nullptr;
  const bool forced = eq_forced(P.eq) && P.P > 0 && op != OP_COEF && op != OP_DERIV;
  if (forced) forcing_terms(P, S.fs, fterm, threadIdx.x, t);

  if (MODE == MODE_LEARNED) {
    // act0 channel 0 already holds inputs / standard_deviation (write_stage_row)
    float* in = S.act0;
    float* out = S.act1;
    for (int l = 0; l < P.nlayers; ++l) {
      const LayerPlan& L = P.layer[l];
      if (P.fast_conv) {
        if (L.cg == 8 && L.pbt == 2) conv5_layer<8, 2>(L, S.blob, in, out, N, P.pitch);
        else if (L.cg == 8) conv5_layer<8, 1>(L, S.blob, in, out, N, P.pitch);
        else if (L.pbt == 2) conv5_layer<4, 2>(L, S.blob, in, out, N, P.pitch);
        else conv5_layer<4, 1>(L, S.blob, in, out, N, P.pitch);
        __syncthreads();
        // the forcing terms written on entry are visible now; the sums become visible at the next barrier
        if (forced && l == 0) {
          forcing_reduce(P, S.fs, threadIdx.x);
          if (P.nlayers == 1) __syncthreads();
        }
      } else {
        conv_generic_layer(L, S.blob, in, out, N, P.pitch, P.K, P.kleft);
        __syncthreads();
        if (forced && l == 0) {
          forcing_reduce(P, S.fs, threadIdx.x);
          if (P.nlayers == 1) __syncthreads();
        }
        if (l + 1 < P.nlayers) {
          fill_halo(out, L.cout, N, P.pitch, P.K, P.kleft);
          __syncthreads();
        }
      }
      float* tmp = in; in = out; out = tmp;
    }
    net = in;
  } else if (forced) {
    __syncthreads();
    forcing_reduce(P, S.fs, threadIdx.x);
    __syncthreads();
  }

  const bool direct_flux = MODE == MODE_LEARNED && P.projection == PROJ_FLUX;
  const bool cons = eq_conservative(P.eq) && !(MODE == MODE_LEARNED && P.projection == PROJ_TIME);
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    float dv[kMaxD];
    float* gc = (op == OP_COEF) ? gout_row + (size_t)p * P.D * P.S : nullptr;
    if (MODE == MODE_LEARNED && P.win == kWinWide) point_derivatives<MODE, kWinWide>(P, S, net, p, dv, gc);
    else point_derivatives<MODE>(P, S, net, p, dv, gc);
    if (op == OP_DERIV) {
#pragma unroll
      for (int d = 0; d < kMaxD; ++d)
        if (d < P.D) gout_row[(size_t)p * P.D + d] = dv[d];
    }
    if (op == OP_COEF || op == OP_DERIV) continue;
    if (MODE == MODE_LEARNED && P.projection == PROJ_TIME) {
      // model_target='time_derivative': the single channel is dy/dt (model.py:603-606)
      float r = net[p + P.kleft];
      if (forced) r = __fadd_rn(r, forcing_at(P, S, p));
      kout[p] = r;
      continue;
    }
    if (MODE == MODE_LEARNED && P.projection == PROJ_FLUX) {
      S.flux[p] = net[p + P.kleft];      // model_target='flux' (model.py:609-615)
      continue;
    }
    float r = equation_point(P.eq, S.ust[p + kRowHalo], dv, P.eta);
    if (cons) {
      S.flux[p] = r;
    } else {
      if (forced) r = __fadd_rn(r, forcing_at(P, S, p));
      kout[p] = r;
    }
  }
  if (op == OP_COEF || op == OP_DERIV) return;
  if (cons || direct_flux) {
    __syncthreads();
    // y_t = -(1/dx) (flux[x+1] - flux[x])  (equations.py:305-320); predict_flux_directly returns
    // +staggered_first_derivative(flux) without the minus sign (model.py:615)
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      float fwd = S.flux[p + 1 == N ? 0 : p + 1];
      float r = __fmul_rn(P.inv_dx, __fsub_rn(fwd, S.flux[p]));
      if (!direct_flux) r = -r;
      if (forced) r = __fadd_rn(r, forcing_at(P, S, p));
      kout[p] = r;
    }
  }
  __syncthreads();
}



// float64 twin of the WENO right-hand side (WENODifferentiator.__call__, integrate.py:133-140):
// u_minus / u_plus reconstructed from the float64 row, the remaining derivatives from the float32
// stencils (a TF float32 graph in the reference, integrate.py:104-105,134), Godunov flux, flux
// difference and forcing in float64 (NumPy semantics: python-scalar * float32 array stays float32).
static __device__ __noinline__ void row_rhs_weno_f64(const Params& P, int sample, double t, int kslot) {
  const Smem S = carve(P, dyn_smem);
  const int N = P.N;
  double* kout = S.kd + kslot * N;
  const bool forced = eq_forced(P.eq) && P.P > 0;
  if (forced && threadIdx.x < P.P) {
    const double* fp = P.fparams64 + (size_t)sample * 4 * P.P;
    const int q = threadIdx.x;
    double sn, cs;
    sincos(fp[P.P + q] * t + fp[2 * P.P + q], &sn, &cs);
    double* parts = S.fsd + 2 * kMaxModes;
    const double a = fp[q], kk = fp[3 * P.P + q];
    parts[q] = a * sn;
    parts[kMaxForcing + q] = (kk < 0.0 ? -a : a) * cs;
    parts[2 * kMaxForcing + q] = fabs(kk);
  }
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    float dv[kMaxD];
    point_derivatives<MODE_STENCIL>(P, S, nullptr, p, dv, nullptr);     // float32 stencil channels
    double um, up;
    weno_pair<double>(S.ustd + p + kRowHalo, um, up);
    const double g = godunov_flux<double>(um, up);
    double flux;
    if (P.eq == EQ_BURGERS_GOD) flux = g - (double)__fmul_rn(P.eta, dv[2]);
    else if (P.eq == EQ_KDV_GOD) flux = 6.0 * g + (double)dv[2];
    else flux = (double)__fadd_rn(dv[3], dv[2]) + g;
    S.fluxd[p] = flux;
  }
  __syncthreads();
  if (forced && threadIdx.x < 2 * P.M) {
    const int j = threadIdx.x;
    const double* parts = S.fsd + 2 * kMaxModes;
    const double m = (double)((j < P.M ? j : j - P.M) + 1);
    const double* src = parts + (j < P.M ? 0 : kMaxForcing);
    double acc = 0.0;
    for (int q = 0; q < P.P; ++q)
      if (parts[2 * kMaxForcing + q] == m) acc += src[q];
    S.fsd[j] = acc;
  }
  if (forced) __syncthreads();
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    const double fwd = S.fluxd[p + 1 == N ? 0 : p + 1];
    double r = -(P.inv_dx_d * (fwd - S.fluxd[p]));
    if (forced) {
      double f = 0.0;
      for (int m = 0; m < P.M; ++m) {
        f += S.fsd[m] * P.fbasis64[(size_t)m * N + p];
        f += S.fsd[P.M + m] * P.fbasis64[(size_t)(P.M + m) * N + p];
      }
      r += f;
    }
    kout[p] = r;
  }
  __syncthreads();
}

// One RHS evaluation into slope slot `kslot`: float32 graph (KT = float) or the float64 WENO twin.
template <int MODE, typename KT>
__device__ __forceinline__ void eval_rhs(const Params& P, const ForcingTerm& fterm, int sample, double t, int kslot) {
  if (sizeof(KT) == sizeof(double)) row_rhs_weno_f64(P, sample, t, kslot);
  else row_rhs<MODE>(P, fterm, (float)t, kslot, OP_RHS, nullptr);
}

// ---------------------------------------------------------------------------------
// Adaptive Bogacki-Shampine 3(2): a per-row twin of scipy.integrate.solve_ivp(method='RK23',
// rtol, atol, max_step, t_eval) as driven by integrate.odeint (integrate.py:143-169), following
// scipy/integrate/_ivp/{rk.py,base.py,common.py,ivp.py}: select_initial_step, the step-size
// controller (SAFETY 0.9, MIN_FACTOR 0.2, MAX_FACTOR 10, exponent -1/3, RMS norm), FSAL, and the
// cubic dense output at t_eval.  float64 state and slopes-as-float64, float32 right-hand side.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(const Smem& S, double v) {
  // deterministic: warp shuffles then a fixed-order sum over the warp partials, broadcast via shared
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) S.red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += S.red[w];
    S.red[32] = t;
  }
  __syncthreads();
  return S.red[32];
}

template <int MODE, typename KT>
__device__ void row_adaptive(const Params& P, const Smem& S, const Work& W, int row, int sample,
                             const ForcingTerm& fterm) {
  const int N = P.N;
  const double t_start = W.times[0], t_bound = W.times[W.ntimes - 1];
  const double rtol = W.rtol, atol = W.atol, max_step = W.max_step;
  const double inv_sqrt_n = 1.0 / sqrt((double)N);
  KT* K0 = slopes<KT>(S);
  KT* K1 = K0 + N;
  KT* K2 = K0 + 2 * N;
  KT* K3 = K0 + 3 * N;
  int nfev = 0, status = 0, next_out = 0;
  double t = t_start;

  // f0 = fun(t0, y0)
  write_stage_row(P, S, [&](int p) { return S.ydbl[p]; });
  __syncthreads();
  eval_rhs<MODE, KT>(P, fterm, sample, t, 0);
  nfev++;

  // select_initial_step (common.py)
  double h_abs;
  {
    const double interval = fabs(t_bound - t_start);
    double s0 = 0.0, s1 = 0.0;
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      const double sc = atol + fabs(S.ydbl[p]) * rtol;
      const double a = S.ydbl[p] / sc, b = (double)K0[p] / sc;
      s0 += a * a;
      s1 += b * b;
    }
    const double d0 = sqrt(block_sum(S, s0)) * inv_sqrt_n;
    const double d1 = sqrt(block_sum(S, s1)) * inv_sqrt_n;
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = fmin(h0, interval);
    write_stage_row(P, S, [&](int p) { return S.ydbl[p] + h0 * (double)K0[p]; });
    __syncthreads();
    eval_rhs<MODE, KT>(P, fterm, sample, t + h0, 1);
    nfev++;
    double s2 = 0.0;
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      const double sc = atol + fabs(S.ydbl[p]) * rtol;
      const double a = ((double)K1[p] - (double)K0[p]) / sc;
      s2 += a * a;
    }
    const double d2 = sqrt(block_sum(S, s2)) * inv_sqrt_n / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 3.0);
    h_abs = fmin(fmin(100.0 * h0, h1), fmin(interval, max_step));
  }

  // t_eval points at (or before) the start are the initial state itself (ivp.py: side='right')
  while (next_out < W.ntimes && W.times[next_out] <= t) {
    double* dst = W.y_out + ((size_t)next_out * W.batch + row) * N;
    for (int p = threadIdx.x; p < N; p += blockDim.x) dst[p] = S.ydbl[p];
    ++next_out;
  }

  while (t < t_bound) {
    const double min_step = 10.0 * fabs(nextafter(t, CUDART_INF) - t);
    if (h_abs > max_step) h_abs = max_step;
    else if (h_abs < min_step) h_abs = min_step;
    bool accepted = false, rejected = false;
    double h = 0.0, t_new = t;
    while (!accepted) {
      if (h_abs < min_step) { status = -1; break; }
      h = h_abs;
      t_new = t + h;
      if (t_new - t_bound > 0.0) t_new = t_bound;
      h = t_new - t;
      h_abs = fabs(h);
      // rk_step (rk.py): K0 = f (FSAL), two inner stages, y_new, f_new
      write_stage_row(P, S, [&](int p) { return S.ydbl[p] + (0.5 * (double)K0[p]) * h; });
      __syncthreads();
      eval_rhs<MODE, KT>(P, fterm, sample, t + 0.5 * h, 1);
      write_stage_row(P, S, [&](int p) { return S.ydbl[p] + (0.75 * (double)K1[p]) * h; });
      __syncthreads();
      eval_rhs<MODE, KT>(P, fterm, sample, t + 0.75 * h, 2);
      for (int p = threadIdx.x; p < N; p += blockDim.x)
        S.ynew[p] = S.ydbl[p] + h * ((2.0 / 9.0) * (double)K0[p] + (1.0 / 3.0) * (double)K1[p] +
                                     (4.0 / 9.0) * (double)K2[p]);
      write_stage_row(P, S, [&](int p) { return S.ynew[p]; });
      __syncthreads();
      eval_rhs<MODE, KT>(P, fterm, sample, t + h, 3);
      nfev += 3;
      double se = 0.0;
      for (int p = threadIdx.x; p < N; p += blockDim.x) {
        const double sc = atol + fmax(fabs(S.ydbl[p]), fabs(S.ynew[p])) * rtol;
        const double e = ((5.0 / 72.0) * (double)K0[p] + (-1.0 / 12.0) * (double)K1[p] +
                          (-1.0 / 9.0) * (double)K2[p] + (1.0 / 8.0) * (double)K3[p]) * h / sc;
        se += e * e;
      }
      const double err = sqrt(block_sum(S, se)) * inv_sqrt_n;
      if (err < 1.0) {
        double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -1.0 / 3.0));
        if (rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        accepted = true;
      } else {
        // NaN error norms land here too: python's max(0.2, nan) is 0.2
        const double shrink = 0.9 * pow(err, -1.0 / 3.0);
        h_abs *= (shrink > 0.2) ? shrink : 0.2;
        rejected = true;
      }
    }
    if (!accepted) break;
    // dense output (RkDenseOutput with P of RK23) for every requested time in (t, t_new]
    while (next_out < W.ntimes && W.times[next_out] <= t_new) {
      const double xq = (W.times[next_out] - t) / h;
      double* dst = W.y_out + ((size_t)next_out * W.batch + row) * N;
      for (int p = threadIdx.x; p < N; p += blockDim.x) {
        const double k0 = K0[p], k1 = K1[p], k2 = K2[p], k3 = K3[p];
        const double q0 = k0;
        const double q1 = (-4.0 / 3.0) * k0 + k1 + (4.0 / 3.0) * k2 - k3;
        const double q2 = (5.0 / 9.0) * k0 + (-2.0 / 3.0) * k1 + (-8.0 / 9.0) * k2 + k3;
        dst[p] = S.ydbl[p] + h * (q0 * xq + q1 * (xq * xq) + q2 * (xq * xq * xq));
      }
      ++next_out;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      S.ydbl[p] = S.ynew[p];
      K0[p] = K3[p];          // FSAL
    }
    t = t_new;
    __syncthreads();
  }
  // the reference pads what the solver did not reach with NaN (integrate.py:161-167)
  for (; next_out < W.ntimes; ++next_out) {
    double* dst = W.y_out + ((size_t)next_out * W.batch + row) * N;
    for (int p = threadIdx.x; p < N; p += blockDim.x) dst[p] = CUDART_NAN;
  }
  if (threadIdx.x == 0) {
    if (W.nfev) W.nfev[row] = nfev;
    if (W.status) W.status[row] = status;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------
// The persistent row kernel
// ---------------------------------------------------------------------------------
template <int MODE, typename KT>
__global__ void __launch_bounds__(MODE == MODE_LEARNED ? 512 : 1024, 1)
    row_kernel(const __grid_constant__ Params P, const __grid_constant__ Work W) {
  const Smem S = carve(P, dyn_smem);
  const int N = P.N;
  KT* const K = slopes<KT>(S);

  // ---- stage the constant blob (filters, biases, window tables) once per CTA ----
  if (P.blob_floats > 0) {
    if (P.use_bulk_copy) {
      if (threadIdx.x == 0) {
        mbar_init(S.bar, 1);
        mbar_fence_init();
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)P.blob_floats * 4u;
        mbar_expect_tx(S.bar, bytes);
        bulk_copy_g2s(S.blob, P.blob, bytes, S.bar);
      }
      mbar_wait(S.bar, 0);
    } else {
      const float4* src = reinterpret_cast<const float4*>(P.blob);
      float4* dst = reinterpret_cast<float4*>(S.blob);
      for (int i = threadIdx.x; i < P.blob_floats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
  }
  __syncthreads();

  const Tableau tab = make_tableau(W.scheme);

  for (int row = blockIdx.x; row < W.batch; row += gridDim.x) {
    const int sample = W.sample_offset + row;
    const ForcingTerm fterm = load_forcing_term(P, sample, threadIdx.x);
    // ---- load the row: float64 master copy + float32 stage row ----
    if (W.u64) {
      for (int p = threadIdx.x; p < N; p += blockDim.x) S.ydbl[p] = W.u64[(size_t)row * N + p];
    } else if ((N & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(W.u + (size_t)row * N);
      for (int i = threadIdx.x; i < N / 4; i += blockDim.x) {
        float4 v = __ldg(src + i);
        S.ydbl[4 * i] = v.x; S.ydbl[4 * i + 1] = v.y; S.ydbl[4 * i + 2] = v.z; S.ydbl[4 * i + 3] = v.w;
      }
    } else {
      for (int p = threadIdx.x; p < N; p += blockDim.x) S.ydbl[p] = W.u[(size_t)row * N + p];
    }
    __syncthreads();

    if (W.op == OP_ADAPTIVE) {
      row_adaptive<MODE, KT>(P, S, W, row, sample, fterm);
      continue;
    }
    if (W.op != OP_INTEGRATE) {
      write_stage_row(P, S, [&](int p) { return S.ydbl[p]; });
      __syncthreads();
      float* gout = nullptr;
      if (W.op == OP_COEF) gout = W.out + (size_t)row * N * P.D * P.S;
      if (W.op == OP_DERIV) gout = W.out + (size_t)row * N * P.D;
      if (W.op == OP_RHS) eval_rhs<MODE, KT>(P, fterm, sample, W.t0, 0);
      else row_rhs<MODE>(P, fterm, (float)W.t0, 0, W.op, gout);
      if (W.op == OP_RHS) {
        if (W.out64) {
          for (int p = threadIdx.x; p < N; p += blockDim.x) W.out64[(size_t)row * N + p] = (double)K[p];
        } else {
          for (int p = threadIdx.x; p < N; p += blockDim.x) W.out[(size_t)row * N + p] = (float)K[p];
        }
      }
      __syncthreads();
      continue;
    }

    // ---- fused fixed-step explicit Runge-Kutta, float64 state / float32 slopes ----
    int first_bad = -1;
    int save_idx = 0;
    for (int step = 0; step < W.nsteps; ++step) {
      const double t = W.t0 + (double)step * W.dt;
      for (int s = 0; s < tab.stages; ++s) {
        // stage input y + dt * sum_j a[s][j] k_j, rounded to float32 as the
        // reference's float32 placeholder feed does (integrate.py:57-60,71)
        write_stage_row(P, S, [&](int p) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < kMaxStages; ++j)
            if (j < s && tab.a[s][j] != 0.0) acc += tab.a[s][j] * (double)K[j * N + p];
          return s == 0 ? S.ydbl[p] : S.ydbl[p] + W.dt * acc;
        });
        __syncthreads();
        eval_rhs<MODE, KT>(P, fterm, sample, t + tab.c[s] * W.dt, s);
      }
      const bool save = ((step + 1) % W.save_every) == 0;
      float* snap = save ? W.snaps + ((size_t)save_idx * W.batch + row) * N : nullptr;
      for (int p = threadIdx.x; p < N; p += blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxStages; ++j)
          if (j < tab.stages && tab.b[j] != 0.0) acc += tab.b[j] * (double)K[j * N + p];
        double yn = S.ydbl[p] + W.dt * acc;
        if (W.state_f32) yn = (double)(float)yn;
        S.ydbl[p] = yn;
        if (first_bad < 0 && !isfinite(yn)) first_bad = step;
        if (save) snap[p] = (float)yn;
      }
      if (save) ++save_idx;
      __syncthreads();
    }
    if (W.first_bad) {
      // min over the CTA of the first non-finite step (-1 = none)
      unsigned int key = first_bad < 0 ? 0xffffffffu : (unsigned int)first_bad;
      unsigned int* slot = reinterpret_cast<unsigned int*>(S.red);
      if (threadIdx.x == 0) *slot = 0xffffffffu;
      __syncthreads();
      atomicMin(slot, key);
      __syncthreads();
      if (threadIdx.x == 0) W.first_bad[row] = (*slot == 0xffffffffu) ? -1 : (int)*slot;
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------
// Stand-alone WENO5 reconstruction (weno.reconstruct_left / reconstruct_right)
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void weno_kernel(const T* __restrict__ u, T* __restrict__ left, T* __restrict__ right, int batch,
                            int N) {
  T* row = reinterpret_cast<T*>(dyn_smem);   // index x + 3, halo 3 + 4
  for (int r = blockIdx.x; r < batch; r += gridDim.x) {
    for (int q = threadIdx.x; q < N + 7; q += blockDim.x) row[q] = u[(size_t)r * N + wrap(q - 3, N)];
    __syncthreads();
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      // reconstruct_*(u)[p] = value at p + 1/2 = the (p+1)-indexed u_minus / u_plus
      T um, up;
      weno_pair<T>(row + p + 1 + 3, um, up);
      left[(size_t)r * N + p] = um;
      right[(size_t)r * N + p] = up;
    }
    __syncthreads();
  }
}

}  // namespace ddd1d
