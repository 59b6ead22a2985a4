// C ABI of libddd1d (see include/ddd1d.h): handle management, packing of the
// reference's TF-layout weights / NumPy tables into the on-chip constant blob,
// launch planning and the kernel launches.  No torch types, no C++ exceptions
// across the boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ddd1d.h"
#include "ddd1d_device.cuh"
#include "ddd1d_tc.cuh"
#include "ddd1d_tc_host.h"
#include "ddd1d_warp.cuh"
#include "ddd1d_weno.cuh"

using namespace ddd1d;

#ifndef DDD1D_AUTO_PREFERS_TENSOR
#define DDD1D_AUTO_PREFERS_TENSOR 1
#endif

namespace {

thread_local std::string g_error;

struct HostLayer {
  bool set = false;
  int k = 0, cin = 0, cout = 0;
  std::vector<float> kernel, bias;
};

}  // namespace

struct ddd1d_handle {
  ddd1d_config cfg;
  Params P;
  std::string error;
  bool dirty = true;
  bool have_stencils = false, have_projection = false;
  const void* warp_kernel = nullptr;   // the warp_row_kernel instantiation last launched, and its resident CTAs per SM
  int warp_occ = 0;
  std::vector<double> stencils;      // [D][7]
  std::vector<double> nullspace;     // [C][7]
  std::vector<int> input_sizes;      // [D]
  HostLayer layers[kMaxLayers];
  float* d_blob = nullptr;
  float* d_fparams = nullptr;
  float* d_fbasis = nullptr;
  double* d_fparams64 = nullptr;
  double* d_fbasis64 = nullptr;
  int forcing_batch = 0, forcing_P = 0, forcing_M = 0;
  int threads = 0, blocks_per_sm = 0, num_sms = 0, weno_block_occ = 0;
  long long launches = 0;
  // tensor-core engine
  tc::TcParams Ptc;
  tc::TcEntry tc_entry;
  bool tc_ok = false;
  int tc_engine = DDD1D_ENGINE_TENSOR;
  std::string tc_why;
  float* d_blob_tc = nullptr;
  float* d_scratch_tc = nullptr;   // tensor engine: per-CTA exchange scratch (stage rows, maxima, flux, forcing)
  // staging for the *_host entry points
  void* d_stage_in = nullptr;
  void* d_stage_out = nullptr;
  int* d_stage_bad = nullptr;
  double* d_times = nullptr;
  size_t times_cap = 0;
  size_t stage_in_bytes = 0, stage_out_bytes = 0, stage_bad_bytes = 0;
};

namespace {

int fail(ddd1d_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->error = buf;
  g_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(h, DDD1D_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),      \
                  __FILE__, __LINE__);                                                         \
  } while (0)

int expected_derivatives(int equation, int variant) {
  // equations.py DERIVATIVE_NAMES of the nine classes
  static const int table[3][3] = {{2, 2, 3}, {2, 2, 3}, {3, 3, 4}};
  return table[equation][variant];
}

int align_up(int x, int a) { return (x + a - 1) / a * a; }

// Host tables arrive in the ABI's 11-slot window (offsets -5..+5); a handle is "wide" when any of them reaches
// beyond the central 7 slots (a 9- or 10-point coefficient grid), else every kernel works on those 7.
bool tables_are_wide(const ddd1d_handle* h);
constexpr int kWinHost = kWinWide;
constexpr int kCentre = (kWinWide - kWin) / 2;      // first host slot of the 7-slot window

double plan_cost(int cout, int N, int nwarps, int cg, int pbt) {
  int ncg = (cout + cg - 1) / cg;
  int npb = (N + 127) / 128;
  if (pbt == 2 && npb < 2) return 1e30;
  int ntasks = ncg * ((npb + pbt - 1) / pbt);
  int rounds = (ntasks + nwarps - 1) / nwarps;
  // FFMAs per input channel + ~4 issue slots per shared-memory load
  double per_ci = 20.0 * pbt * cg + 4.0 * (2.0 * pbt + 1.25 * cg);
  return rounds * per_ci;
}



bool tables_are_wide(const ddd1d_handle* h) {
  auto outer = [](const std::vector<double>& t) {
    for (size_t r = 0; r + kWinHost <= t.size(); r += kWinHost)
      for (int j = 0; j < kCentre; ++j)
        if (t[r + j] != 0.0 || t[r + kWinHost - 1 - j] != 0.0) return true;
    return false;
  };
  const ddd1d_config& c = h->cfg;
  if (c.mode == DDD1D_MODE_LEARNED && c.projection <= DDD1D_PROJ_RAW_UNBIASED && c.stencil_size > kWin) return true;
  return (h->have_stencils && outer(h->stencils)) || (h->have_projection && outer(h->nullspace));
}

int engine_request(const ddd1d_handle* h) {
  int e = h->cfg.engine;
  const char* env = getenv("DDD1D_ENGINE");
  if (e == DDD1D_ENGINE_AUTO && env) {
    if (!strcmp(env, "ffma")) e = DDD1D_ENGINE_FFMA;
    if (!strcmp(env, "tensor")) e = DDD1D_ENGINE_TENSOR;
    if (!strcmp(env, "tensor_f16x2")) e = DDD1D_ENGINE_TENSOR_F16X2;
    if (!strcmp(env, "tensor_f16")) e = DDD1D_ENGINE_TENSOR_F16;
  }
  return e;
}

bool wants_tensor(int e) {
  return e == DDD1D_ENGINE_TENSOR || e == DDD1D_ENGINE_TENSOR_F16X2 || e == DDD1D_ENGINE_TENSOR_F16;
}

// Build the tensor-core engine's operand blob and constant-bank tables (ddd1d_tc.cuh) when the net has the
// shape that kernel is written for.  Not being eligible is not an error unless the caller forced a
// DDD1D_ENGINE_TENSOR* engine.
int finalize_tc(ddd1d_handle* h) {
  const ddd1d_config& c = h->cfg;
  h->tc_ok = false;
  h->tc_why.clear();
  const int N = c.num_points, D = c.num_derivatives;
  const int want = engine_request(h);
  int tiles = 0, rpt = 1;
  switch (N) {
    case 128: tiles = 1; break;
    case 256: tiles = 2; break;
    case 512: tiles = 4; break;
    case 64: tiles = 1; rpt = 2; break;      // two rows per 128-position tile
    case 32: tiles = 1; rpt = 4; break;      // four
    default: break;
  }
  if (c.mode != DDD1D_MODE_LEARNED) h->tc_why = "not a learned-coefficient handle";
  else if (c.kernel_size != 5 || c.filter_size != tc::kF) h->tc_why = "needs kernel_size 5 and filter_size 32";
  else if (c.num_layers < 2 || c.num_layers > 3) h->tc_why = "needs 2 or 3 conv layers";
  else if (c.activation != DDD1D_ACT_RELU) h->tc_why = "needs the relu nonlinearity";
  else if (c.projection > DDD1D_PROJ_RAW_UNBIASED) h->tc_why = "only model_target='coefficients'";
  else if (tiles == 0) h->tc_why = "needs num_points in {32, 64, 128, 256, 512}";
  else if (tables_are_wide(h)) h->tc_why = "coefficient grids beyond 7 points run on the FFMA engine";
  else if (h->forcing_batch > 0 && h->forcing_P > 0 && h->forcing_M > kMaxModes) h->tc_why = "too many forcing modes";
  if (!h->tc_why.empty() || want == DDD1D_ENGINE_FFMA) {
    if (wants_tensor(want))
      return fail(h, DDD1D_EUNSUPPORTED, "tensor engine unavailable: %s", h->tc_why.c_str());
    return DDD1D_OK;
  }
  const int prec = want == DDD1D_ENGINE_TENSOR_F16 ? 1 : want == DDD1D_ENGINE_TENSOR_F16X2 ? 2 : 3;
  const int L = c.num_layers;
  const int NL = (D * kWin <= 16) ? 16 : 32;
  const int K = 5, F = tc::kF;
  tc::TcEntry entry;
  // DDD1D_TC_SLOTS=3: three rows per team on the TMEM block pool where it is compiled (N = 256, NL = 16)
  const bool pool = getenv("DDD1D_TC_SLOTS") && atoi(getenv("DDD1D_TC_SLOTS")) == 3 && tiles == 2 && rpt == 1 &&
                    tc::lookup_t2_pool(NL, prec, &entry);
  if (!pool && !tc::lookup(tiles, rpt, NL, prec, &entry)) {
    h->tc_why = "this (num_points, precision) combination is not compiled";
    if (wants_tensor(want)) return fail(h, DDD1D_EUNSUPPORTED, "tensor engine unavailable: %s", h->tc_why.c_str());
    return DDD1D_OK;
  }
  tc::TcParams T;
  memset(&T, 0, sizeof(T));
  // linear map net channel -> window coefficient (projection folded into the last layer)
  const int Q = D * kWin;
  std::vector<double> pm((size_t)c.net_outputs * Q, 0.0), pbias(Q, 0.0);
  if (c.projection == DDD1D_PROJ_NULLSPACE) {
    int ch = 0;
    for (int d = 0; d < D; ++d)
      for (int i = 0; i < h->input_sizes[d]; ++i, ++ch)
        for (int j = 0; j < kWin; ++j) pm[(size_t)ch * Q + d * kWin + j] = h->nullspace[(size_t)ch * kWinHost + kCentre + j];
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < kWin; ++j) pbias[d * kWin + j] = h->stencils[d * kWinHost + kCentre + j];
  } else {
    const int S = c.stencil_size, ws = 3 - S / 2;
    for (int d = 0; d < D; ++d)
      for (int i = 0; i < S; ++i)
        for (int i2 = 0; i2 < S; ++i2) {
          double v = (i == i2) ? 1.0 : 0.0;
          if (c.projection == DDD1D_PROJ_RAW_UNBIASED) v -= 1.0 / S;
          pm[(size_t)(d * S + i2) * Q + d * kWin + i + ws] = v;
        }
  }
  // Operand format of the tensor layers: fp16 planes, value * scale = hi + lo (~22 significant bits per
  // operand, K = 16 per MMA); scales are powers of two, so they are exact.
  std::vector<float> blob((size_t)(entry.blob_hidden_bytes + entry.blob_last_bytes) / 4, 0.f);
  const int planes = 4;                   // 16-byte chunk planes per tap (8 input channels each)
  const int cpp = 8;
  auto pow2_scale = [](double maxabs) {                       // largest 2^e with maxabs * 2^e < 2^14
    if (!(maxabs > 0.0)) return 1.0;
    int e;
    std::frexp(maxabs, &e);
    return std::ldexp(1.0, std::min(14 - e, 60));
  };
  // writes one filter value into a [Wh | Wl] plane pair (rows = 2 * nb)
  auto put_weight = [&](float* cat, int nb, int k, int ci, int col, double w, double sw) {
    const size_t plane = (size_t)(k * planes + ci / cpp) * (2 * nb) * 4;      // in floats (16 B per row)
    __half* hp = reinterpret_cast<__half*>(cat + plane);
    const float v = (float)(w * sw);
    const __half hi = __float2half_rn(v);
    hp[(size_t)col * 8 + (ci % 8)] = hi;
    hp[(size_t)(nb + col) * 8 + (ci % 8)] = __float2half_rn((v - __half2float(hi)) * tc::kLoScale);
  };
  const HostLayer& l0 = h->layers[0];
  for (int k = 0; k < K; ++k)
    for (int co = 0; co < F; ++co) T.w1[k * F + co] = l0.kernel[(size_t)k * F + co];
  T.inv_sw_hid = T.inv_sw_last = 1.f;
  for (int co = 0; co < F; ++co) {
    T.b1[co] = l0.bias[co];
    double a = 0.0;
    for (int k = 0; k < K; ++k) a += std::fabs((double)l0.kernel[(size_t)k * F + co]);
    T.w1abs = std::max(T.w1abs, (float)(a * (1.0 + 1e-6)));
    T.b1abs = std::max(T.b1abs, std::fabs(l0.bias[co]));
  }
  T.nhid = L - 2;
  if (T.nhid == 1) {
    const HostLayer& hl = h->layers[1];
    double wmax = 0.0;
    for (size_t i = 0; i < (size_t)K * F * F; ++i) wmax = std::max(wmax, std::fabs((double)hl.kernel[i]));
    const double sw = pow2_scale(wmax);
    T.inv_sw_hid = (float)(1.0 / sw);
    for (int co = 0; co < F; ++co) {
      T.bh[co] = hl.bias[co];
      double a = 0.0;
      for (int k = 0; k < K; ++k)
        for (int ci = 0; ci < F; ++ci) a += std::fabs((double)hl.kernel[((size_t)k * F + ci) * F + co]);
      T.whabs = std::max(T.whabs, (float)(a * (1.0 + 1e-6)));
      T.bhabs = std::max(T.bhabs, std::fabs(hl.bias[co]));
    }
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < F; ++ci)
        for (int co = 0; co < F; ++co) put_weight(blob.data(), F, k, ci, co, hl.kernel[((size_t)k * F + ci) * F + co], sw);
  }
  // last layer with the projection folded in: W'[k][ci][q] = sum_c W[k][ci][c] pm[c][q]
  {
    const HostLayer& ll = h->layers[L - 1];
    std::vector<double> folded((size_t)K * F * Q, 0.0);
    double wmax = 0.0;
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < F; ++ci)
        for (int q = 0; q < Q; ++q) {
          double acc = 0.0;
          for (int ch = 0; ch < c.net_outputs; ++ch)
            acc += (double)ll.kernel[((size_t)k * F + ci) * c.net_outputs + ch] * pm[(size_t)ch * Q + q];
          folded[((size_t)k * F + ci) * Q + q] = acc;
          wmax = std::max(wmax, std::fabs(acc));
        }
    const double sw = pow2_scale(wmax);
    T.inv_sw_last = (float)(1.0 / sw);
    float* cat = blob.data() + entry.blob_hidden_bytes / 4;
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < F; ++ci)
        for (int q = 0; q < Q; ++q) put_weight(cat, NL, k, ci, q, folded[((size_t)k * F + ci) * Q + q], sw);
    for (int q = 0; q < Q; ++q) {
      double acc = pbias[q];
      for (int ch = 0; ch < c.net_outputs; ++ch) acc += (double)ll.bias[ch] * pm[(size_t)ch * Q + q];
      T.bl[q] = (float)acc;
    }
  }
  T.eq = h->P.eq; T.D = D; T.S = c.stencil_size; T.wshift = 3 - (c.stencil_size / 2);
  T.plain_burgers = T.eq == EQ_BURGERS; T.plain_kdv = T.eq == EQ_KDV; T.plain_ks = T.eq == EQ_KS;
  T.M = h->P.M; T.P = h->P.P; T.fcap = h->P.fcap;
  T.sigma = h->P.sigma; T.eta = h->P.eta; T.inv_dx = h->P.inv_dx;
  T.fparams = h->P.fparams; T.fbasis = h->P.fbasis;
  T.debug = getenv("DDD1D_TC_DEBUG") ? atoi(getenv("DDD1D_TC_DEBUG")) : 0;
  CUDA_TRY(h, cudaSetDevice(c.device));
  if (h->d_blob_tc) CUDA_TRY(h, cudaFree(h->d_blob_tc));
  h->d_blob_tc = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->d_blob_tc, blob.size() * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(h->d_blob_tc, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  T.blob = h->d_blob_tc;
  if (h->d_scratch_tc) CUDA_TRY(h, cudaFree(h->d_scratch_tc));
  h->d_scratch_tc = nullptr;
  const size_t scratch_floats = (size_t)h->num_sms * entry.slots_per_cta * entry.sc_stride;
  CUDA_TRY(h, cudaMalloc(&h->d_scratch_tc, scratch_floats * sizeof(float)));
  CUDA_TRY(h, cudaMemset(h->d_scratch_tc, 0, scratch_floats * sizeof(float)));
  T.scratch = h->d_scratch_tc;
  CUDA_TRY(h, cudaFuncSetAttribute(entry.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, entry.smem_bytes));
  int occ = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, entry.kernel, entry.threads, entry.smem_bytes));
  if (occ < 1) {
    h->tc_why = "tensor kernel does not fit on an SM";
    if (wants_tensor(want)) return fail(h, DDD1D_EUNSUPPORTED, "%s", h->tc_why.c_str());
    return DDD1D_OK;
  }
  h->tc_entry = entry;
  h->Ptc = T;
  h->tc_ok = true;
  h->tc_engine = prec == 1 ? DDD1D_ENGINE_TENSOR_F16 : prec == 2 ? DDD1D_ENGINE_TENSOR_F16X2 : DDD1D_ENGINE_TENSOR;
  return DDD1D_OK;
}

// CTAs of a tensor-engine launch: every team takes two slots per round
int tc_grid(const ddd1d_handle* h, int batch) {
  const tc::TcEntry& e = h->tc_entry;
  const int units = (batch + e.rows_per_slot - 1) / e.rows_per_slot;
  return std::max(1, std::min((units + e.teams - 1) / e.teams, h->num_sms));
}

bool use_tc(const ddd1d_handle* h) {
  if (!h->tc_ok) return false;
  const int want = engine_request(h);
  if (wants_tensor(want)) return true;
  if (want == DDD1D_ENGINE_FFMA) return false;
  return DDD1D_AUTO_PREFERS_TENSOR != 0;
}

// Build P (plans, blob, shared-memory carve-up) from the host-side description.
int finalize(ddd1d_handle* h) {
  if (!h->dirty) return DDD1D_OK;
  const ddd1d_config& c = h->cfg;
  Params& P = h->P;
  const int N = c.num_points;
  const int D = c.num_derivatives;
  if (!h->have_stencils && !(c.mode == DDD1D_MODE_LEARNED && c.projection != DDD1D_PROJ_NULLSPACE))
    return fail(h, DDD1D_ESTATE, "ddd1d_set_stencils has not been called");

  // window of this handle's kernels: the central 7 slots unless a table reaches offsets +-4
  const bool wide = tables_are_wide(h);
  if (wide && c.mode != DDD1D_MODE_LEARNED)
    return fail(h, DDD1D_EUNSUPPORTED, "fixed stencils wider than 7 points are not built (offsets beyond +-3)");
  const int win = wide ? kWinWide : kWin, wpitch = wide ? kWinWidePad : kWinPad, first = wide ? 0 : kCentre;
  P.win = win;
  std::vector<float> blob;
  P.nlayers = 0;
  P.fast_conv = 0;
  P.pitch = 0;
  int chan_max = 0;
  if (c.mode == DDD1D_MODE_LEARNED) {
    const int K = c.kernel_size;
    P.nlayers = c.num_layers;
    P.K = K;
    P.kleft = K / 2;                 // ceil((K-1)/2), layers.py:77
    P.fast_conv = (K == 5 && (N % 4) == 0 && N >= 4) ? 1 : 0;
    if (getenv("DDD1D_FORCE_GENERIC_CONV")) P.fast_conv = 0;
    P.pitch = align_up(N + K - 1, 4);
    // one thread per position (4x8 register tiles, 16 warps/SM at N=256: +12 % over 8x8 tiles on 8 warps)
    h->threads = P.fast_conv ? std::min(512, std::max(128, align_up(N, 128))) : 256;
    if (P.fast_conv && getenv("DDD1D_FFMA_THREADS")) h->threads = std::max(64, std::min(512, atoi(getenv("DDD1D_FFMA_THREADS")) / 32 * 32));
    const int nwarps = h->threads / 32;
    int cin = 1;
    for (int l = 0; l < c.num_layers; ++l) {
      const HostLayer& hl = h->layers[l];
      const int want_out = (l == c.num_layers - 1) ? c.net_outputs : c.filter_size;
      if (!hl.set) return fail(h, DDD1D_ESTATE, "ddd1d_set_layer(%d) has not been called", l);
      if (hl.k != K || hl.cin != cin || hl.cout != want_out)
        return fail(h, DDD1D_EINVAL, "layer %d has shape [%d,%d,%d], expected [%d,%d,%d]", l, hl.k,
                    hl.cin, hl.cout, K, cin, want_out);
      LayerPlan& L = P.layer[l];
      L.cin = cin;
      L.cout = hl.cout;
      L.cout_pad = align_up(hl.cout, 8);
      L.act = (l == c.num_layers - 1) ? DDD1D_ACT_NONE : c.activation;
      L.w_off = (int)blob.size();
      blob.resize(blob.size() + (size_t)cin * K * L.cout_pad, 0.f);
      // TF kernel [k][ci][co] -> W[(ci*K + k)*cout_pad + co]
      for (int k = 0; k < K; ++k)
        for (int ci = 0; ci < cin; ++ci)
          for (int co = 0; co < hl.cout; ++co)
            blob[L.w_off + ((size_t)ci * K + k) * L.cout_pad + co] =
                hl.kernel[((size_t)k * cin + ci) * hl.cout + co];
      L.b_off = (int)blob.size();
      blob.resize(blob.size() + L.cout_pad, 0.f);
      for (int co = 0; co < hl.cout; ++co) blob[L.b_off + co] = hl.bias[co];
      // tile plan
      double best = 1e30;
      L.cg = 8; L.pbt = 1;
      const int cgs[2] = {8, 4}, pbts[2] = {2, 1};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          double cost = plan_cost(L.cout, N, nwarps, cgs[a], pbts[b]);
          if (cost < best) { best = cost; L.cg = cgs[a]; L.pbt = pbts[b]; }
        }
      chan_max = std::max(chan_max, std::max(cin, L.cout_pad));
      cin = hl.cout;
    }
    P.C = c.net_outputs;
    P.projection = c.projection;
    P.S = c.stencil_size;
    P.wshift = win / 2 - (c.stencil_size / 2);   // win/2 - ceil((S-1)/2)
    if (c.projection == DDD1D_PROJ_NULLSPACE) {
      if (!h->have_projection) return fail(h, DDD1D_ESTATE, "ddd1d_set_projection has not been called");
      int total = 0;
      P.cstart[0] = 0;
      for (int d = 0; d < D; ++d) {
        total += h->input_sizes[d];
        P.cstart[d + 1] = total;
      }
      for (int d = D; d < kMaxD; ++d) P.cstart[d + 1] = total;
      if (total != c.net_outputs)
        return fail(h, DDD1D_EINVAL, "sum(input_sizes)=%d but net_outputs=%d", total, c.net_outputs);
      P.ns_off = (int)blob.size();
      blob.resize(blob.size() + (size_t)c.net_outputs * wpitch, 0.f);
      for (int ch = 0; ch < c.net_outputs; ++ch)
        for (int j = 0; j < win; ++j)
          blob[P.ns_off + ch * wpitch + j] = (float)h->nullspace[(size_t)ch * kWinHost + first + j];
    } else {
      const int want_out = c.projection == DDD1D_PROJ_DERIVATIVES ? D
                           : c.projection >= DDD1D_PROJ_TIME_DERIVATIVE ? 1 : D * c.stencil_size;
      if (c.net_outputs != want_out)
        return fail(h, DDD1D_EINVAL, "projection %d needs net_outputs == %d, got %d", c.projection, want_out,
                    c.net_outputs);
      P.ns_off = 0;
      for (int d = 0; d <= kMaxD; ++d) P.cstart[d] = 0;
    }
  } else {
    h->threads = std::min(512, std::max(64, align_up(N, 32)));
    P.K = 1; P.kleft = 0; P.C = 0; P.projection = 0; P.S = 0; P.wshift = 0; P.ns_off = 0;
    for (int d = 0; d <= kMaxD; ++d) P.cstart[d] = 0;
  }
  // stencil / bias window table [kMaxD][8]
  P.st_off = (int)blob.size();
  blob.resize(blob.size() + kMaxD * wpitch, 0.f);
  if (h->have_stencils)
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < win; ++j) blob[P.st_off + d * wpitch + j] = (float)h->stencils[d * kWinHost + first + j];
  blob.resize(align_up((int)blob.size(), 4), 0.f);

  P.eq = c.equation * 3 + c.variant;
  P.mode = c.mode;
  P.N = N;
  P.D = D;
  P.weno_real = c.weno_real;
  P.sigma = (float)c.standard_deviation;
  P.eta = (float)c.eta;
  P.inv_dx = (float)(1.0 / c.dx);
  P.inv_dx_d = 1.0 / c.dx;
  P.blob_floats = (int)blob.size();

  // shared-memory carve-up
  int off = 0;
  P.off_bar = off; off += 16;
  P.off_blob = off; off += P.blob_floats * 4;
  P.off_ust = off; off += align_up((N + 2 * kRowHalo) * 4, 16);
  P.off_ydbl = off; off += align_up(N * 8, 16);
  P.off_ynew = off; off += align_up(N * 8, 16);
  P.off_red = off; off += 34 * 8;
  P.off_k = off; off += align_up(kMaxStages * N * 4, 16);
  P.off_flux = off; off += align_up((N + 1) * 4, 16);
  P.off_fs = off; off += (2 * kMaxModes + 3 * kMaxForcing) * 4;
  P.off_ustd = P.off_kd = P.off_fluxd = P.off_fsd = off;
  if (c.mode == DDD1D_MODE_WENO && c.weno_real == DDD1D_REAL_F64) {
    off = align_up(off, 16);
    P.off_ustd = off; off += align_up((N + 2 * kRowHalo) * 8, 16);
    P.off_kd = off; off += kMaxStages * N * 8;
    P.off_fluxd = off; off += align_up((N + 1) * 8, 16);
    P.off_fsd = off; off += (2 * kMaxModes + 3 * kMaxForcing) * 8;
  }
  P.off_act0 = off;
  P.off_act1 = off;
  if (c.mode == DDD1D_MODE_LEARNED) {
    int bytes = align_up(chan_max * P.pitch * 4, 16);
    P.off_act0 = off; off += bytes;
    P.off_act1 = off; off += bytes;
  }
  P.smem_bytes = off;
  if (P.smem_bytes > 227 * 1024)
    return fail(h, DDD1D_EUNSUPPORTED,
                "row of %d points with this net needs %d bytes of shared memory (limit 232448)", N,
                P.smem_bytes);
  P.use_bulk_copy = getenv("DDD1D_NO_BULK_COPY") ? 0 : 1;

  CUDA_TRY(h, cudaSetDevice(c.device));
  if (h->d_blob) CUDA_TRY(h, cudaFree(h->d_blob));
  CUDA_TRY(h, cudaMalloc(&h->d_blob, blob.size() * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(h->d_blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  P.blob = h->d_blob;
  P.fparams = h->d_fparams;
  P.fbasis = h->d_fbasis;
  P.fparams64 = h->d_fparams64;
  P.fbasis64 = h->d_fbasis64;
  P.P = h->forcing_batch > 0 ? h->forcing_P : 0;
  P.M = h->forcing_M;
  P.fcap = h->forcing_batch;

  // launch shape
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, c.device));
  h->num_sms = prop.multiProcessorCount;
  const bool weno64 = c.mode == DDD1D_MODE_WENO && c.weno_real == DDD1D_REAL_F64;
  const void* fn = c.mode == DDD1D_MODE_LEARNED  ? (const void*)row_kernel<MODE_LEARNED, float>
                   : weno64                      ? (const void*)row_kernel<MODE_WENO, double>
                   : c.mode == DDD1D_MODE_WENO   ? (const void*)row_kernel<MODE_WENO, float>
                                                 : (const void*)row_kernel<MODE_STENCIL, float>;
  CUDA_TRY(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
  int occ = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, h->threads, P.smem_bytes));
  if (occ < 1) return fail(h, DDD1D_EUNSUPPORTED, "kernel does not fit on an SM (%d threads, %d B smem)",
                           h->threads, P.smem_bytes);
  h->blocks_per_sm = occ;
  int rc_tc = finalize_tc(h);
  if (rc_tc) return rc_tc;
  h->dirty = false;
  return DDD1D_OK;
}

constexpr int kWarpRowsPerBlock = 8;      // warps (= rows in flight) per CTA of warp_row_kernel

// The warp-per-row kernel takes the fused fixed-step integration of the modes without a conv net when a row
// is exactly 1, 2, 4 or 8 points per lane.  DDD1D_NO_WARP_ROWS=1 keeps the CTA-per-row kernel (A/B, tests).
bool use_warp_rows(const ddd1d_handle* h, int op) {
  const ddd1d_config& c = h->cfg;
  if (op != OP_INTEGRATE || c.mode == DDD1D_MODE_LEARNED) return false;
  if (c.mode == DDD1D_MODE_WENO && c.weno_real != DDD1D_REAL_F32) return false;
  const int n = c.num_points;
  if (n % 32 != 0 || (n != 32 && n != 64 && n != 128 && n != 256)) return false;
  if (h->P.P > 32) return false;            // one forcing term per lane
  return getenv("DDD1D_NO_WARP_ROWS") == nullptr;
}

// The warp_row_kernel instantiation of a handle (points per lane x WENO x forcing modes x stencil reach) and,
// cached on the handle, how many of its CTAs are resident per SM: the grid is exactly the resident CTAs
// (persistent warps), so that no CTA waits for a slot.
const void* pick_warp_kernel(ddd1d_handle* h) {
  const ddd1d_config& c = h->cfg;
  const Params& P = h->P;
  const bool weno = c.mode == DDD1D_MODE_WENO;
  const bool few = P.M <= 3;                 // forcing modes (the reference's k_max = 3)
  // how far the stencil table reaches from a point (WENO5 reads three points either side)
  int halo = 1;
  if (weno) halo = kHalo;
  else
    for (int d = 0; d < c.num_derivatives; ++d)
      for (int j = 0; j < kWin; ++j)
        if (h->stencils[(size_t)d * kWinHost + kCentre + j] != 0.0) halo = std::max(halo, std::abs(j - kHalo));
  const void* kernel = nullptr;
  // (the Burgers forms of the baseline configurations get instantiations with the equation folded in)
#define DDD1D_WARP_PICK_EQ(PPL, EQ)                                                                   \
  do {                                                                                                \
    if (weno) kernel = few ? (const void*)warp_row_kernel<PPL, true, 3, 3, EQ> : (const void*)warp_row_kernel<PPL, true, kMaxModes, 3, EQ>;   \
    else if (halo == 1) kernel = few ? (const void*)warp_row_kernel<PPL, false, 3, 1, EQ> : (const void*)warp_row_kernel<PPL, false, kMaxModes, 1, EQ>;   \
    else if (halo == 2) kernel = few ? (const void*)warp_row_kernel<PPL, false, 3, 2, EQ> : (const void*)warp_row_kernel<PPL, false, kMaxModes, 2, EQ>;   \
    else kernel = few ? (const void*)warp_row_kernel<PPL, false, 3, 3, EQ> : (const void*)warp_row_kernel<PPL, false, kMaxModes, 3, EQ>;   \
  } while (0)
#define DDD1D_WARP_PICK(PPL)                                                                          \
  do {                                                                                                \
    if (P.eq == EQ_BURGERS) DDD1D_WARP_PICK_EQ(PPL, EQ_BURGERS);                                      \
    else if (P.eq == EQ_BURGERS_GOD) DDD1D_WARP_PICK_EQ(PPL, EQ_BURGERS_GOD);                         \
    else DDD1D_WARP_PICK_EQ(PPL, -1);                                                                 \
  } while (0)
  switch (c.num_points / 32) {
    case 1: DDD1D_WARP_PICK(1); break;
    case 2: DDD1D_WARP_PICK(2); break;
    case 4: DDD1D_WARP_PICK(4); break;
    default: DDD1D_WARP_PICK(8); break;
  }
#undef DDD1D_WARP_PICK_EQ
#undef DDD1D_WARP_PICK
  // persistent warps: exactly the CTAs that are resident at once, so that no CTA waits for a slot
  if (h->warp_kernel != kernel) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * kWarpRowsPerBlock, 0) != cudaSuccess || occ < 1) occ = 1;
    h->warp_kernel = kernel;
    h->warp_occ = occ;
  }
  return kernel;
}

int warp_rows_grid(ddd1d_handle* h, int batch) {
  pick_warp_kernel(h);
  return std::max(1, std::min((batch + kWarpRowsPerBlock - 1) / kWarpRowsPerBlock, h->num_sms * h->warp_occ));
}

// The register-resident CTA-per-row WENO5 integrator (ddd1d_weno.cuh) takes the fused fixed-step integration of
// float32 WENO rows of 4 points per thread: N a multiple of 128 up to 2048 (BASELINE config 5).
bool use_weno_block(const ddd1d_handle* h, int op) {
  const ddd1d_config& c = h->cfg;
  if (op != OP_INTEGRATE || c.mode != DDD1D_MODE_WENO || c.weno_real != DDD1D_REAL_F32) return false;
  const int n = c.num_points;
  if (n % 128 != 0 || n <= 256 || n > 2048) return false;
  if (h->P.P > 32) return false;            // one forcing term per lane of warp 0
  return getenv("DDD1D_NO_WENO_BLOCK") == nullptr;
}

int weno_block_grid(ddd1d_handle* h, int batch) {
  if (h->weno_block_occ == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, weno_block_kernel<-1>, h->cfg.num_points / kWenoPpt, 0) != cudaSuccess || occ < 1)
      occ = 1;
    h->weno_block_occ = occ;
  }
  return std::max(1, std::min(batch, h->num_sms * h->weno_block_occ));
}

int launch(ddd1d_handle* h, Work& W, void* stream) {
  int rc = finalize(h);
  if (rc) return rc;
  if (W.batch <= 0) return DDD1D_OK;
  const ddd1d_config& c = h->cfg;
  const Params& P = h->P;
  if (c.equation == DDD1D_BURGERS && P.P > 0 && (W.op == OP_RHS || W.op == OP_INTEGRATE || W.op == OP_ADAPTIVE) &&
      (W.sample_offset < 0 || W.sample_offset + W.batch > P.fcap))
    return fail(h, DDD1D_EINVAL, "samples [%d, %d) exceed the %d forcing rows set", W.sample_offset,
                W.sample_offset + W.batch, P.fcap);
  CUDA_TRY(h, cudaSetDevice(c.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // dt * tableau as float32 / float-float constants (the float-pair state of the tensor and warp-row kernels)
    const Tableau tab = make_tableau(W.scheme);
    for (int s = 0; s < kMaxStages; ++s) {
      for (int j = 0; j < kMaxStages; ++j) W.adt[s][j] = (float)(W.dt * tab.a[s][j]);
      const double b = W.dt * tab.b[s];
      W.bdt_hi[s] = (float)b;
      W.bdt_lo[s] = (float)(b - (double)W.bdt_hi[s]);
    }
  }
  if (use_tc(h) && W.op != OP_ADAPTIVE) {
    W.integrating = W.op == OP_INTEGRATE;
    h->tc_entry.launch(h->Ptc, W, make_tableau(W.scheme), tc_grid(h, W.batch), st);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return DDD1D_OK;
  }
  if (use_warp_rows(h, W.op)) {
    // one warp per row, everything in registers (ddd1d_warp.cuh): the fixed-step integrator of the
    // fixed-stencil / float32 WENO modes for rows of 32 * {1, 2, 4, 8} points
    const Tableau tab = make_tableau(W.scheme);
    const void* kernel = pick_warp_kernel(h);
    const int blocks = warp_rows_grid(h, W.batch);
    void* args[] = {(void*)&P, (void*)&W, (void*)&tab};
    CUDA_TRY(h, cudaLaunchKernel(kernel, dim3(blocks), dim3(32 * kWarpRowsPerBlock), args, 0, st));
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return DDD1D_OK;
  }
  if (use_weno_block(h, W.op) && W.u) {
    if (P.eq == EQ_BURGERS_GOD)
      weno_block_kernel<EQ_BURGERS_GOD><<<weno_block_grid(h, W.batch), c.num_points / kWenoPpt, 0, st>>>(P, W, make_tableau(W.scheme));
    else
      weno_block_kernel<-1><<<weno_block_grid(h, W.batch), c.num_points / kWenoPpt, 0, st>>>(P, W, make_tableau(W.scheme));
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return DDD1D_OK;
  }
  const int grid = std::min(W.batch, h->num_sms * h->blocks_per_sm);
  if (c.mode == DDD1D_MODE_LEARNED)
    row_kernel<MODE_LEARNED, float><<<grid, h->threads, P.smem_bytes, st>>>(P, W);
  else if (c.mode == DDD1D_MODE_WENO && c.weno_real == DDD1D_REAL_F64)
    row_kernel<MODE_WENO, double><<<grid, h->threads, P.smem_bytes, st>>>(P, W);
  else if (c.mode == DDD1D_MODE_WENO)
    row_kernel<MODE_WENO, float><<<grid, h->threads, P.smem_bytes, st>>>(P, W);
  else
    row_kernel<MODE_STENCIL, float><<<grid, h->threads, P.smem_bytes, st>>>(P, W);
  CUDA_TRY(h, cudaGetLastError());
  h->launches += 1;
  return DDD1D_OK;
}

Work blank_work() {
  Work W;
  memset(&W, 0, sizeof(W));
  W.save_every = 1;
  return W;
}

int ensure_stage(ddd1d_handle* h, void** buf, size_t* have, size_t need) {
  if (*have >= need) return DDD1D_OK;
  if (*buf) CUDA_TRY(h, cudaFree(*buf));
  *buf = nullptr;
  *have = 0;
  CUDA_TRY(h, cudaMalloc(buf, need));
  *have = need;
  return DDD1D_OK;
}

}  // namespace

namespace ddd1d {
namespace tc {
bool lookup(int tiles, int rpt, int nl, int prec, TcEntry* out) {
  if (rpt == 1 && tiles == 1) return lookup_t1(nl, prec, out);
  if (rpt == 1 && tiles == 2) return lookup_t2(nl, prec, out);
  if (rpt == 1 && tiles == 4) return lookup_t4(nl, prec, out);
  if (tiles == 1 && rpt == 2) return lookup_p2(nl, prec, out);
  if (tiles == 1 && rpt == 4) return lookup_p4(nl, prec, out);
  return false;
}
}  // namespace tc
}  // namespace ddd1d

extern "C" {

int ddd1d_version(void) { return DDD1D_VERSION; }

const char* ddd1d_last_error(const ddd1d_handle* handle) {
  return handle ? handle->error.c_str() : g_error.c_str();
}

int ddd1d_create(const ddd1d_config* config, ddd1d_handle** out) {
  if (!config || !out) return fail(nullptr, DDD1D_EINVAL, "null argument");
  *out = nullptr;
  if (config->struct_bytes != (int)sizeof(ddd1d_config))
    return fail(nullptr, DDD1D_EINVAL, "ddd1d_config is %d bytes, library expects %d",
                config->struct_bytes, (int)sizeof(ddd1d_config));
  const ddd1d_config& c = *config;
  if (c.equation < 0 || c.equation > 2 || c.variant < 0 || c.variant > 2)
    return fail(nullptr, DDD1D_EINVAL, "unknown equation/variant %d/%d", c.equation, c.variant);
  if (c.mode < 0 || c.mode > 2) return fail(nullptr, DDD1D_EINVAL, "unknown mode %d", c.mode);
  if (c.num_points < 1) return fail(nullptr, DDD1D_EINVAL, "num_points must be positive");
  if (c.num_derivatives != expected_derivatives(c.equation, c.variant))
    return fail(nullptr, DDD1D_EINVAL, "equation %d/%d has %d derivative channels, got %d", c.equation,
                c.variant, expected_derivatives(c.equation, c.variant), c.num_derivatives);
  if (!(c.dx > 0)) return fail(nullptr, DDD1D_EINVAL, "dx must be positive");
  if (c.mode == DDD1D_MODE_WENO && c.variant != DDD1D_GODUNOV)
    return fail(nullptr, DDD1D_EINVAL, "WENO mode needs a Godunov-flux equation (integrate.py:320-321)");
  if (c.weno_real != DDD1D_REAL_F32 && c.weno_real != DDD1D_REAL_F64)
    return fail(nullptr, DDD1D_EINVAL, "unknown weno_real %d", c.weno_real);
  if (c.mode == DDD1D_MODE_LEARNED) {
    if (c.num_layers < 1 || c.num_layers > kMaxLayers)
      return fail(nullptr, c.num_layers == 0 ? DDD1D_EUNSUPPORTED : DDD1D_EINVAL,
                  "num_layers=%d outside 1..%d", c.num_layers, kMaxLayers);
    if (c.kernel_size < 1 || c.kernel_size > 15)
      return fail(nullptr, DDD1D_EUNSUPPORTED, "kernel_size=%d outside 1..15", c.kernel_size);
    if (c.filter_size < 1 || c.net_outputs < 1) return fail(nullptr, DDD1D_EINVAL, "bad net widths");
    if (c.activation < 0 || c.activation > DDD1D_ACT_ELU)
      return fail(nullptr, DDD1D_EINVAL, "unknown activation %d", c.activation);
    if (c.projection < 0 || c.projection > DDD1D_PROJ_FLUX)
      return fail(nullptr, DDD1D_EINVAL, "unknown projection %d", c.projection);
    if (c.projection <= DDD1D_PROJ_RAW_UNBIASED && (c.stencil_size < 1 || c.stencil_size > kWinHost))
      return fail(nullptr, DDD1D_EUNSUPPORTED,
                  "coefficient grid of %d points exceeds the %d-point window", c.stencil_size, kWinHost);
    if (!(c.standard_deviation > 0)) return fail(nullptr, DDD1D_EINVAL, "standard_deviation must be positive");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || c.device < 0 || c.device >= ndev)
    return fail(nullptr, DDD1D_ECUDA, "CUDA device %d unavailable (%s, %d devices)", c.device,
                cudaGetErrorString(e), ndev);
  ddd1d_handle* h = new ddd1d_handle();
  h->cfg = c;
  memset(&h->P, 0, sizeof(h->P));
  *out = h;
  return DDD1D_OK;
}

int ddd1d_destroy(ddd1d_handle* h) {
  if (!h) return DDD1D_OK;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->d_blob);
  cudaFree(h->d_blob_tc);
  cudaFree(h->d_scratch_tc);
  cudaFree(h->d_fparams);
  cudaFree(h->d_fbasis);
  cudaFree(h->d_fparams64);
  cudaFree(h->d_fbasis64);
  cudaFree(h->d_stage_in);
  cudaFree(h->d_stage_out);
  cudaFree(h->d_stage_bad);
  cudaFree(h->d_times);
  delete h;
  return DDD1D_OK;
}

int ddd1d_set_stencils(ddd1d_handle* h, const double* w) {
  if (!h || !w) return fail(h, DDD1D_EINVAL, "null argument");
  h->stencils.assign(w, w + (size_t)h->cfg.num_derivatives * kWinHost);
  h->have_stencils = true;
  h->dirty = true;
  return DDD1D_OK;
}

int ddd1d_set_layer(ddd1d_handle* h, int layer, const float* kernel, const float* bias, int kernel_size,
                    int cin, int cout) {
  if (!h || !kernel || !bias) return fail(h, DDD1D_EINVAL, "null argument");
  if (h->cfg.mode != DDD1D_MODE_LEARNED) return fail(h, DDD1D_EINVAL, "handle has no conv net");
  if (layer < 0 || layer >= h->cfg.num_layers)
    return fail(h, DDD1D_EINVAL, "layer %d outside 0..%d", layer, h->cfg.num_layers - 1);
  if (kernel_size < 1 || cin < 1 || cout < 1) return fail(h, DDD1D_EINVAL, "bad layer shape");
  HostLayer& L = h->layers[layer];
  L.set = true;
  L.k = kernel_size; L.cin = cin; L.cout = cout;
  L.kernel.assign(kernel, kernel + (size_t)kernel_size * cin * cout);
  L.bias.assign(bias, bias + cout);
  h->dirty = true;
  return DDD1D_OK;
}

int ddd1d_set_projection(ddd1d_handle* h, const double* ns, const int* input_sizes) {
  if (!h || !ns || !input_sizes) return fail(h, DDD1D_EINVAL, "null argument");
  if (h->cfg.mode != DDD1D_MODE_LEARNED) return fail(h, DDD1D_EINVAL, "handle has no conv net");
  int total = 0;
  for (int d = 0; d < h->cfg.num_derivatives; ++d) {
    if (input_sizes[d] < 0) return fail(h, DDD1D_EINVAL, "negative input size");
    total += input_sizes[d];
  }
  if (total != h->cfg.net_outputs)
    return fail(h, DDD1D_EINVAL, "sum(input_sizes)=%d but net_outputs=%d", total, h->cfg.net_outputs);
  h->nullspace.assign(ns, ns + (size_t)total * kWinHost);
  h->input_sizes.assign(input_sizes, input_sizes + h->cfg.num_derivatives);
  h->have_projection = true;
  h->dirty = true;
  return DDD1D_OK;
}

int ddd1d_set_forcing(ddd1d_handle* h, const double* a, const double* omega, const double* k,
                      const double* phi, int batch, int nparams, int resample_factor, int mean_resample,
                      double period) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (h->d_fparams) { CUDA_TRY(h, cudaFree(h->d_fparams)); h->d_fparams = nullptr; }
  if (h->d_fbasis) { CUDA_TRY(h, cudaFree(h->d_fbasis)); h->d_fbasis = nullptr; }
  if (h->d_fparams64) { CUDA_TRY(h, cudaFree(h->d_fparams64)); h->d_fparams64 = nullptr; }
  if (h->d_fbasis64) { CUDA_TRY(h, cudaFree(h->d_fbasis64)); h->d_fbasis64 = nullptr; }
  h->forcing_batch = 0; h->forcing_P = 0; h->forcing_M = 0;
  h->dirty = true;
  if (batch == 0) return DDD1D_OK;
  if (!a || !omega || !k || !phi || batch < 0 || nparams < 1 || resample_factor < 1 || !(period > 0))
    return fail(h, DDD1D_EINVAL, "bad forcing arguments");
  const int N = h->cfg.num_points, P = nparams;
  int M = 0;
  for (size_t i = 0; i < (size_t)batch * P; ++i) {
    double kk = k[i];
    if (kk != std::floor(kk)) return fail(h, DDD1D_EINVAL, "forcing wavenumbers must be integers");
    M = std::max(M, (int)std::fabs(kk));
  }
  if (M > kMaxModes) return fail(h, DDD1D_EUNSUPPORTED, "|k| up to %d supported, got %d", kMaxModes, M);
  if (nparams > kMaxForcing)
    return fail(h, DDD1D_EUNSUPPORTED, "up to %d forcing terms per sample supported, got %d", kMaxForcing, nparams);
  if (M == 0) M = 1;
  std::vector<float> fp((size_t)batch * 4 * P);
  std::vector<double> fp64((size_t)batch * 4 * P);
  for (int b = 0; b < batch; ++b)
    for (int q = 0; q < P; ++q) {
      size_t s = (size_t)b * P + q, base = (size_t)b * 4 * P;
      fp[base + q] = (float)(fp64[base + q] = a[s]);
      fp[base + P + q] = (float)(fp64[base + P + q] = omega[s]);
      fp[base + 2 * P + q] = (float)(fp64[base + 2 * P + q] = phi[s]);
      fp[base + 3 * P + q] = (float)(fp64[base + 3 * P + q] = k[s]);
    }
  // basis on the reference grid, resampled like Grid.resample (equations.py:65-68)
  std::vector<float> basis((size_t)2 * M * N);
  std::vector<double> basis64((size_t)2 * M * N);
  const int Nref = N * resample_factor;
  const double two_pi = 6.283185307179586476925286766559;
  for (int m = 1; m <= M; ++m)
    for (int x = 0; x < N; ++x) {
      double cs = 0.0, sn = 0.0;
      const int count = mean_resample ? resample_factor : 1;
      for (int r = 0; r < count; ++r) {
        double xr = period / Nref * (double)(x * resample_factor + r);
        double th = two_pi * m * xr / period;
        cs += std::cos(th);
        sn += std::sin(th);
      }
      basis[(size_t)(m - 1) * N + x] = (float)(basis64[(size_t)(m - 1) * N + x] = cs / count);
      basis[(size_t)(M + m - 1) * N + x] = (float)(basis64[(size_t)(M + m - 1) * N + x] = sn / count);
    }
  CUDA_TRY(h, cudaMalloc(&h->d_fparams, fp.size() * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(h->d_fparams, fp.data(), fp.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMalloc(&h->d_fbasis, basis.size() * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(h->d_fbasis, basis.data(), basis.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMalloc(&h->d_fparams64, fp64.size() * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(h->d_fparams64, fp64.data(), fp64.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMalloc(&h->d_fbasis64, basis64.size() * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(h->d_fbasis64, basis64.data(), basis64.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->forcing_batch = batch; h->forcing_P = P; h->forcing_M = M;
  return DDD1D_OK;
}

int ddd1d_rhs(ddd1d_handle* h, double t, const float* u, float* dudt, int batch, int sample_offset,
              void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (batch == 0) return DDD1D_OK;
  if (!u || !dudt) return fail(h, DDD1D_EINVAL, "null argument");
  Work W = blank_work();
  W.op = OP_RHS; W.batch = batch; W.sample_offset = sample_offset; W.u = u; W.out = dudt; W.t0 = t;
  return launch(h, W, stream);
}

int ddd1d_rhs_f64(ddd1d_handle* h, double t, const double* u, double* dudt, int batch, int sample_offset,
                  void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (batch == 0) return DDD1D_OK;
  if (!u || !dudt) return fail(h, DDD1D_EINVAL, "null argument");
  Work W = blank_work();
  W.op = OP_RHS; W.batch = batch; W.sample_offset = sample_offset; W.u64 = u; W.out64 = dudt; W.t0 = t;
  return launch(h, W, stream);
}

int ddd1d_coefficients(ddd1d_handle* h, const float* u, float* coefficients, int batch, void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (batch == 0) return DDD1D_OK;
  if (!u || !coefficients) return fail(h, DDD1D_EINVAL, "null argument");
  if (h->cfg.mode != DDD1D_MODE_LEARNED) return fail(h, DDD1D_EINVAL, "handle has no conv net");
  if (h->cfg.projection > DDD1D_PROJ_RAW_UNBIASED)
    return fail(h, DDD1D_EINVAL, "this model_target predicts no coefficients");
  Work W = blank_work();
  W.op = OP_COEF; W.batch = batch; W.u = u; W.out = coefficients;
  return launch(h, W, stream);
}

int ddd1d_space_derivatives(ddd1d_handle* h, const float* u, float* derivatives, int batch, void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (batch == 0) return DDD1D_OK;
  if (!u || !derivatives) return fail(h, DDD1D_EINVAL, "null argument");
  if (h->cfg.mode == DDD1D_MODE_LEARNED && h->cfg.projection > DDD1D_PROJ_DERIVATIVES)
    return fail(h, DDD1D_EINVAL, "this model_target predicts no space derivatives");
  Work W = blank_work();
  W.op = OP_DERIV; W.batch = batch; W.u = u; W.out = derivatives;
  return launch(h, W, stream);
}

int ddd1d_integrate(ddd1d_handle* h, double t0, double dt, int num_steps, int save_every, int scheme,
                    const float* u0, float* snapshots, int* first_bad_step, int batch, int sample_offset,
                    void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (num_steps < 0 || save_every < 1) return fail(h, DDD1D_EINVAL, "bad step counts");
  if (batch == 0) return DDD1D_OK;
  if (!u0 || (!snapshots && num_steps / save_every > 0)) return fail(h, DDD1D_EINVAL, "null argument");
  const int state_f32 = (scheme & DDD1D_STATE_F32) != 0;
  scheme &= ~DDD1D_STATE_F32;
  if (scheme < 0 || scheme > DDD1D_RK4) return fail(h, DDD1D_EINVAL, "unknown scheme %d", scheme);
  Work W = blank_work();
  W.state_f32 = state_f32;
  W.op = OP_INTEGRATE; W.batch = batch; W.sample_offset = sample_offset; W.u = u0; W.snaps = snapshots;
  W.first_bad = first_bad_step; W.t0 = t0; W.dt = dt; W.nsteps = num_steps; W.save_every = save_every;
  W.scheme = scheme;
  return launch(h, W, stream);
}

int ddd1d_integrate_adaptive(ddd1d_handle* h, const double* times, int num_times, double rtol, double atol,
                             double max_step, const float* u0, const double* u0_f64, double* y_out, int* nfev,
                             int* status, int batch, int sample_offset, void* stream) {
  if (!h) return fail(h, DDD1D_EINVAL, "null handle");
  if (!times || num_times < 2) return fail(h, DDD1D_EINVAL, "need at least two output times");
  for (int i = 1; i < num_times; ++i)
    if (!(times[i] > times[i - 1])) return fail(h, DDD1D_EINVAL, "times must be strictly increasing");
  if (!(rtol > 0) || !(atol >= 0) || !(max_step > 0)) return fail(h, DDD1D_EINVAL, "bad tolerances");
  if (batch == 0) return DDD1D_OK;
  if ((!u0 && !u0_f64) || !y_out) return fail(h, DDD1D_EINVAL, "null argument");
  int rc = finalize(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (h->times_cap < (size_t)num_times) {
    if (h->d_times) CUDA_TRY(h, cudaFree(h->d_times));
    h->d_times = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->d_times, (size_t)num_times * sizeof(double)));
    h->times_cap = (size_t)num_times;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_times, times, (size_t)num_times * sizeof(double), cudaMemcpyHostToDevice, st));
  Work W = blank_work();
  W.op = OP_ADAPTIVE; W.batch = batch; W.sample_offset = sample_offset; W.u = u0; W.u64 = u0_f64;
  W.times = h->d_times; W.ntimes = num_times; W.rtol = rtol; W.atol = atol; W.max_step = max_step;
  W.y_out = y_out; W.nfev = nfev; W.status = status;
  return launch(h, W, stream);
}

int ddd1d_rhs_host(ddd1d_handle* h, double t, const double* u, double* dudt, int batch, int sample_offset) {
  if (!h || !u || !dudt) return fail(h, DDD1D_EINVAL, "null argument");
  int rc = finalize(h);
  if (rc) return rc;
  const size_t bytes = (size_t)batch * h->cfg.num_points * sizeof(double);
  if ((rc = ensure_stage(h, &h->d_stage_in, &h->stage_in_bytes, bytes))) return rc;
  if ((rc = ensure_stage(h, &h->d_stage_out, &h->stage_out_bytes, bytes))) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage_in, u, bytes, cudaMemcpyHostToDevice, 0));
  rc = ddd1d_rhs_f64(h, t, (const double*)h->d_stage_in, (double*)h->d_stage_out, batch, sample_offset,
                     nullptr);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(dudt, h->d_stage_out, bytes, cudaMemcpyDeviceToHost, 0));
  CUDA_TRY(h, cudaStreamSynchronize(0));
  return DDD1D_OK;
}

int ddd1d_integrate_host(ddd1d_handle* h, double t0, double dt, int num_steps, int save_every, int scheme,
                         const float* u0, float* snapshots, int* first_bad_step, int batch,
                         int sample_offset) {
  if (!h || !u0 || !snapshots) return fail(h, DDD1D_EINVAL, "null argument");
  if (save_every < 1) return fail(h, DDD1D_EINVAL, "bad step counts");
  int rc = finalize(h);
  if (rc) return rc;
  const size_t row = (size_t)batch * h->cfg.num_points * sizeof(float);
  const size_t nsave = (size_t)(num_steps / save_every);
  if ((rc = ensure_stage(h, &h->d_stage_in, &h->stage_in_bytes, row))) return rc;
  if ((rc = ensure_stage(h, &h->d_stage_out, &h->stage_out_bytes, std::max<size_t>(row * nsave, 16)))) return rc;
  void* bad = nullptr;
  if (first_bad_step) {
    if ((rc = ensure_stage(h, (void**)&h->d_stage_bad, &h->stage_bad_bytes, (size_t)batch * sizeof(int)))) return rc;
    bad = h->d_stage_bad;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage_in, u0, row, cudaMemcpyHostToDevice, 0));
  rc = ddd1d_integrate(h, t0, dt, num_steps, save_every, scheme, (const float*)h->d_stage_in,
                       (float*)h->d_stage_out, (int*)bad, batch, sample_offset, nullptr);
  if (rc) return rc;
  if (nsave) CUDA_TRY(h, cudaMemcpyAsync(snapshots, h->d_stage_out, row * nsave, cudaMemcpyDeviceToHost, 0));
  if (first_bad_step)
    CUDA_TRY(h, cudaMemcpyAsync(first_bad_step, bad, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, 0));
  CUDA_TRY(h, cudaStreamSynchronize(0));
  return DDD1D_OK;
}

int ddd1d_weno_reconstruct(int device, int real, const void* u, void* left, void* right, int batch,
                           int num_points, void* stream) {
  if (!u || !left || !right || batch < 0 || num_points < 1)
    return fail(nullptr, DDD1D_EINVAL, "bad argument");
  if (batch == 0) return DDD1D_OK;
  CUDA_TRY(nullptr, cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = std::min(512, std::max(64, align_up(num_points, 32)));
  const int grid = std::min(batch, 148 * 8);
  if (real == DDD1D_REAL_F64) {
    size_t smem = (size_t)(num_points + 7) * sizeof(double);
    if (smem > 48 * 1024)
      CUDA_TRY(nullptr, cudaFuncSetAttribute(weno_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weno_kernel<double><<<grid, threads, smem, st>>>((const double*)u, (double*)left, (double*)right, batch, num_points);
  } else if (real == DDD1D_REAL_F32) {
    size_t smem = (size_t)(num_points + 7) * sizeof(float);
    if (smem > 48 * 1024)
      CUDA_TRY(nullptr, cudaFuncSetAttribute(weno_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weno_kernel<float><<<grid, threads, smem, st>>>((const float*)u, (float*)left, (float*)right, batch, num_points);
  } else {
    return fail(nullptr, DDD1D_EINVAL, "unknown real type %d", real);
  }
  CUDA_TRY(nullptr, cudaGetLastError());
  return DDD1D_OK;
}

long long ddd1d_launch_count(const ddd1d_handle* h) { return h ? h->launches : 0; }

int ddd1d_engine(const ddd1d_handle* handle) {
  ddd1d_handle* h = const_cast<ddd1d_handle*>(handle);
  if (!h) return fail(nullptr, DDD1D_EINVAL, "null handle");
  int rc = finalize(h);
  if (rc) return rc;
  return use_tc(h) ? h->tc_engine : DDD1D_ENGINE_FFMA;
}

int ddd1d_launch_shape(const ddd1d_handle* handle, int batch, int* grid, int* block, int* shared_bytes) {
  ddd1d_handle* h = const_cast<ddd1d_handle*>(handle);
  if (!h) return fail(nullptr, DDD1D_EINVAL, "null handle");
  int rc = finalize(h);
  if (rc) return rc;
  if (use_tc(h)) {
    if (grid) *grid = tc_grid(h, batch);
    if (block) *block = h->tc_entry.threads;
    if (shared_bytes) *shared_bytes = h->tc_entry.smem_bytes;
    return DDD1D_OK;
  }
  if (use_weno_block(h, OP_INTEGRATE)) {
    if (grid) *grid = weno_block_grid(h, batch);
    if (block) *block = h->cfg.num_points / kWenoPpt;
    if (shared_bytes) *shared_bytes = (int)sizeof(WenoShared);
    return DDD1D_OK;
  }
  if (use_warp_rows(h, OP_INTEGRATE)) {     // (the shape of ddd1d_integrate launches)
    if (grid) *grid = warp_rows_grid(h, batch);
    if (block) *block = 256;
    if (shared_bytes) *shared_bytes = 0;
    return DDD1D_OK;
  }
  if (grid) *grid = std::min(batch, h->num_sms * h->blocks_per_sm);
  if (block) *block = h->threads;
  if (shared_bytes) *shared_bytes = h->P.smem_bytes;
  return DDD1D_OK;
}

}  // extern "C"
