/*
 * ddd1d.h -- C ABI of the B200-native 1-D PDE time integrator with learned
 * finite-difference coefficients (drop-in for the hot path of
 * google/data-driven-discretization-1d, package `pde_superresolution`).
 *
 * The reference has no FFI: its boundary is a Python call surface that ends in
 * `sess.run` (integrate.py:70-71,101-102).  Each entry point below names the
 * reference call it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.  All citations are relative to
 * /root/reference/pde_superresolution/.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ types, no torch types;
 *   - every function returns 0 on success or a negative DDD1D_E* code; the text of
 *     the last failure on a handle is available from ddd1d_last_error();
 *   - "device" pointers are CUDA device pointers owned by the caller; "host"
 *     pointers are ordinary host memory.  `stream` is a cudaStream_t passed as
 *     void* (NULL = the legacy default stream).  Device entry points are
 *     asynchronous and stream ordered; *_host entry points copy in and out and
 *     return after the result is in host memory;
 *   - numerical blow-up is data, not an error: rows turn NaN and
 *     first_bad_step[] records when (the reference NaN-pads, integrate.py:161-167);
 *   - one handle per (device, stream); handles are independent and re-entrant.
 *
 * Stencil window: every derivative channel is an 11-wide window over offsets
 * -5..+5 relative to the output point.  A stencil of `s` points applied with the
 * reference's centred periodic padding (layers.py:76-79: ceil((s-1)/2) points on
 * the left) occupies window slots 5-ceil((s-1)/2) ... ; unused slots are zero.
 * Tables that use only the central 7 slots (every stencil of up to 7 points: the
 * reference's defaults) run on 7-slot kernels; the coefficient grids of
 * hparams.coefficient_grid_min_size = 9 (9 centred or 10 staggered points,
 * model.py:445-448, training_test.py:56) run on the FFMA engine.
 */
#ifndef DDD1D_H_
#define DDD1D_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DDD1D_VERSION 2
#define DDD1D_MAX_DERIVATIVES 4
#define DDD1D_WINDOW 11
#define DDD1D_MAX_LAYERS 6
#define DDD1D_MAX_FORCING_MODES 8

/* error codes */
#define DDD1D_OK 0
#define DDD1D_EINVAL (-1)      /* bad argument (the reference's ValueError) */
#define DDD1D_EUNSUPPORTED (-2)/* valid for the reference, not built here (size / hparam) */
#define DDD1D_ECUDA (-3)       /* a CUDA call failed; see ddd1d_last_error */
#define DDD1D_ESTATE (-4)      /* call order: weights / stencils / forcing not set */

/* equations.py:590-606 registries: EQUATION_TYPES / CONSERVATIVE_ / FLUX_ */
enum { DDD1D_BURGERS = 0, DDD1D_KDV = 1, DDD1D_KS = 2 };
enum { DDD1D_PLAIN = 0, DDD1D_CONSERVATIVE = 1, DDD1D_GODUNOV = 2 };

/* how spatial derivatives are produced */
enum {
  DDD1D_MODE_STENCIL = 0, /* fixed coefficients: model.baseline_space_derivatives, model.py:99-109 */
  DDD1D_MODE_LEARNED = 1, /* conv net -> coefficients: model.predict_coefficients, model.py:420-513 */
  DDD1D_MODE_WENO = 2     /* u_minus/u_plus from WENO5 (weno.py), the rest fixed: integrate.py:124-140 */
};

/* model.py:411-417 */
enum { DDD1D_ACT_NONE = 0, DDD1D_ACT_RELU = 1, DDD1D_ACT_RELU6 = 2, DDD1D_ACT_TANH = 3,
       DDD1D_ACT_SOFTPLUS = 4, DDD1D_ACT_ELU = 5 };

/* how the last conv's channels become coefficients (model.py:460-511) */
enum {
  DDD1D_PROJ_NULLSPACE = 0,   /* coef_d = bias_d + net[slice_d] @ nullspace_d  (polynomials.py:266-277) */
  DDD1D_PROJ_RAW = 1,         /* polynomial_accuracy_order == 0: coef = reshape(net, [D, S]) */
  DDD1D_PROJ_RAW_UNBIASED = 2,/* ... minus the mean over the stencil (ensure_unbiased_coefficients) */
  /* the other hparams.model_target values (model.py:551-640): no stencil is applied */
  DDD1D_PROJ_DERIVATIVES = 3, /* 'space_derivatives': net_outputs == D, channels are the derivatives */
  DDD1D_PROJ_TIME_DERIVATIVE = 4, /* 'time_derivative': net_outputs == 1, the channel is dy/dt */
  DDD1D_PROJ_FLUX = 5         /* 'flux': net_outputs == 1, dy/dt = staggered_first_derivative(channel) */
};

/* explicit Runge-Kutta schemes for the fused fixed-step integrator */
enum {
  DDD1D_RK3_BS = 0,   /* Bogacki-Shampine 3-stage: what scipy RK23 executes when pinned at max_step (integrate.py:154-155) */
  DDD1D_MIDPOINT = 1, /* tf.contrib.integrate.odeint_fixed(method='midpoint'), model.py:138-159 */
  DDD1D_EULER = 2,
  DDD1D_RK4 = 3
};
/* OR into `scheme`: round the carried solution to float32 after every step, as the reference's TF-graph unroll
 * does (tf.contrib.integrate.odeint_fixed on a float32 tensor, model.py:138-159).  Without it the state is
 * carried in float64 between steps, as SciPy carries it (integrate.py:154; the RHS is float32 either way). */
#define DDD1D_STATE_F32 0x100

/* conv-stack engine (learned mode).  AUTO picks the tcgen05 kernel in its FP32-faithful form (TENSOR) when
 * the net is the shape it is built for (kernel_size 5, filter_size 32, relu, 2 or 3 layers, N in
 * {32, 64, 128, 256, 512}) and the FP32-FFMA kernel otherwise.  FFMA and TENSOR satisfy the same float32
 * tolerances: TENSOR splits every operand into fp16 hi + lo (22 bits) and accumulates in FP32.  The two
 * cheaper tensor forms are opt-in only, with their own (looser) accuracy, measured over the configured
 * 10 000-step horizons in profiles/r02/tc_trajectory_error.json:
 *   TENSOR_F16X2  activations rounded to fp16 (11 bits), filters at 22 bits: one MMA per K step, no lo planes;
 *   TENSOR_F16    plain fp16 operands, FP32 accumulate. */
enum { DDD1D_ENGINE_AUTO = 0, DDD1D_ENGINE_FFMA = 1, DDD1D_ENGINE_TENSOR = 2, DDD1D_ENGINE_TENSOR_F16X2 = 3,
       DDD1D_ENGINE_TENSOR_F16 = 4 };

/* arithmetic of the WENO reconstruction: float32 (TF path, model.py:81-87) or
 * float64 (NumPy path of WENODifferentiator, integrate.py:137-138) */
enum { DDD1D_REAL_F32 = 0, DDD1D_REAL_F64 = 1 };

typedef struct ddd1d_handle ddd1d_handle;

typedef struct ddd1d_config {
  int struct_bytes;        /* sizeof(ddd1d_config), for forward compatibility */
  int device;              /* CUDA device ordinal */
  int equation;            /* DDD1D_BURGERS / KDV / KS */
  int variant;             /* DDD1D_PLAIN / CONSERVATIVE / GODUNOV */
  int mode;                /* DDD1D_MODE_* */
  int num_points;          /* N = equation.grid.solution_num_points */
  int num_derivatives;     /* D = len(equation.DERIVATIVE_ORDERS), must match (equation, variant) */
  int weno_real;           /* DDD1D_REAL_* (MODE_WENO only) */
  double dx;               /* equation.grid.solution_dx */
  double eta;              /* Burgers viscosity, equations.py:244 */
  double standard_deviation; /* input normalisation, model.py:450-451 */
  /* --- learned mode only (hparams, training.py:133-141) --- */
  int num_layers;          /* conv layers including the last, >= 1 */
  int filter_size;         /* F */
  int kernel_size;         /* K */
  int activation;          /* DDD1D_ACT_* of the hidden layers */
  int net_outputs;         /* C = channels of the last conv */
  int stencil_size;        /* S = size of the coefficient grid (6 staggered, 7 centred; up to 11) */
  int projection;          /* DDD1D_PROJ_* */
  int engine;              /* DDD1D_ENGINE_*: which kernel evaluates the conv stack */
} ddd1d_config;

/* Build a handle.  Replaces graph construction in SavedModelDifferentiator /
 * PolynomialDifferentiator / WENODifferentiator.__init__ (integrate.py:51-68,77-99,127-131).
 * Fails with DDD1D_EINVAL on inconsistent shapes (the reference's ValueError at
 * model.py:53-56) and DDD1D_EUNSUPPORTED when the row does not fit on chip. */
int ddd1d_create(const ddd1d_config* config, ddd1d_handle** out);
int ddd1d_destroy(ddd1d_handle* handle);
/* Never NULL; "" when no error has occurred.  handle may be NULL (global/create errors). */
const char* ddd1d_last_error(const ddd1d_handle* handle);
int ddd1d_version(void);

/* Fixed coefficients per derivative channel in window form, host double [D][DDD1D_WINDOW].
 * STENCIL mode: polynomials.coefficients() of each derivative (polynomials.py:280-303);
 * LEARNED mode with PROJ_NULLSPACE: PolynomialAccuracyLayer.bias (polynomials.py:237-239);
 * WENO mode: rows of u_minus/u_plus are ignored. */
int ddd1d_set_stencils(ddd1d_handle* handle, const double* window_coefficients);

/* One conv layer, TF variable layout: kernel [K][cin][cout], bias [cout], host float32.
 * Replaces tf.train.Saver.restore of predict_coefficients/conv1d[_i]/{kernel,bias}
 * (integrate.py:66-68, model.py:442). */
int ddd1d_set_layer(ddd1d_handle* handle, int layer, const float* kernel, const float* bias,
                    int kernel_size, int cin, int cout);

/* Null-space rows in window form, host double [C][DDD1D_WINDOW], and the number of net
 * channels feeding each derivative, int [D] (PolynomialAccuracyLayer.nullspace /
 * input_size, polynomials.py:246-262; model.py:504-511). */
int ddd1d_set_projection(ddd1d_handle* handle, const double* window_nullspace, const int* input_sizes);

/* Per-sample forcing parameters of equations.RandomForcing (equations.py:196-219),
 * host double [batch][nparams] each; k holds signed integer wavenumbers.  The term
 * is evaluated on the reference grid (num_points*resample_factor) and resampled
 * (mean when mean_resample != 0, else subsampled) exactly like grid.resample
 * (equations.py:65-68,215-219).  Passing batch == 0 disables forcing. */
int ddd1d_set_forcing(ddd1d_handle* handle, const double* a, const double* omega, const double* k,
                      const double* phi, int batch, int nparams, int resample_factor,
                      int mean_resample, double period);

/* dy/dt for `batch` rows.  Replaces Differentiator.__call__ -> sess.run
 * (integrate.py:70-71,101-102,133-140).  `sample_offset` selects which forcing
 * rows the batch uses.  u, dudt: device float32 [batch][N]. */
int ddd1d_rhs(ddd1d_handle* handle, double t, const float* u, float* dudt, int batch,
              int sample_offset, void* stream);
/* Same with float64 rows (the dtype SciPy hands to the Differentiator). */
int ddd1d_rhs_f64(ddd1d_handle* handle, double t, const double* u, double* dudt, int batch,
                  int sample_offset, void* stream);

/* model.predict_coefficients (model.py:420-513): device float32 coef [batch][N][D][S]. */
int ddd1d_coefficients(ddd1d_handle* handle, const float* u, float* coefficients, int batch,
                       void* stream);

/* model.predict_space_derivatives / baseline_space_derivatives (model.py:59-112,579-600):
 * device float32 derivatives [batch][N][D]. */
int ddd1d_space_derivatives(ddd1d_handle* handle, const float* u, float* derivatives, int batch,
                            void* stream);

/* The fused persistent integrator: `num_steps` explicit RK steps of size dt from t0
 * with every row resident on chip; a snapshot every `save_every` steps.
 * Replaces odeint's step loop (integrate.py:154-155) for a whole batch, and
 * model.integrate_ode (model.py:138-159) with DDD1D_MIDPOINT.
 *   u0              device float32 [batch][N]
 *   snapshots       device float32 [num_steps/save_every][batch][N]
 *   first_bad_step  device int32 [batch] or NULL: first step whose result was
 *                   non-finite, or -1 */
int ddd1d_integrate(ddd1d_handle* handle, double t0, double dt, int num_steps, int save_every,
                    int scheme, const float* u0, float* snapshots, int* first_bad_step,
                    int batch, int sample_offset, void* stream);

/* Adaptive twin of integrate.odeint for a whole batch: every row runs SciPy's RK23
 * (Bogacki-Shampine 3(2) with the scipy.integrate.solve_ivp step-size controller,
 * FSAL, cubic dense output at `times`) on the device, rows independent
 * (integrate.py:143-169 -> scipy/integrate/_ivp/rk.py).
 *   times    HOST float64 [num_times], strictly increasing; times[0] is the start
 *   u0 / u0_f64  device float32 or float64 [batch][N] (exactly one non-NULL)
 *   y_out    device float64 [num_times][batch][N]; samples the solver did not reach
 *            are NaN (the reference's NaN padding, integrate.py:161-167)
 *   nfev     device int32 [batch] or NULL: right-hand-side evaluations, as sol.nfev
 *   status   device int32 [batch] or NULL: 0 = reached times[-1], -1 = step size underflow
 * The conv stack runs on the FFMA engine for this entry point. */
int ddd1d_integrate_adaptive(ddd1d_handle* handle, const double* times, int num_times, double rtol,
                             double atol, double max_step, const float* u0, const double* u0_f64,
                             double* y_out, int* nfev, int* status, int batch, int sample_offset,
                             void* stream);

/* Host-buffer forms (copies inside): what a caller without device memory uses. */
int ddd1d_rhs_host(ddd1d_handle* handle, double t, const double* u, double* dudt, int batch,
                   int sample_offset);
int ddd1d_integrate_host(ddd1d_handle* handle, double t0, double dt, int num_steps, int save_every,
                         int scheme, const float* u0, float* snapshots, int* first_bad_step,
                         int batch, int sample_offset);

/* weno.reconstruct_left / reconstruct_right (weno.py:92-123) on device rows
 * [batch][N]; real = DDD1D_REAL_F32 (float*) or DDD1D_REAL_F64 (double*). */
int ddd1d_weno_reconstruct(int device, int real, const void* u, void* left, void* right,
                           int batch, int num_points, void* stream);

/* Introspection for benchmarks: kernels launched so far through this handle, and
 * the launch shape the integrator uses (grid, block, dynamic shared bytes). */
long long ddd1d_launch_count(const ddd1d_handle* handle);
int ddd1d_launch_shape(const ddd1d_handle* handle, int batch, int* grid, int* block, int* shared_bytes);
/* DDD1D_ENGINE_FFMA or one of DDD1D_ENGINE_TENSOR*: the engine the next launch will use. */
int ddd1d_engine(const ddd1d_handle* handle);

#ifdef __cplusplus
}
#endif
#endif  /* DDD1D_H_ */
