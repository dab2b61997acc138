/*
 * bridge_b200.h -- C ABI of the B200-native hot path of mschauer/Bridge.jl.
 *
 * One shared library (libbridge_b200.so, hand-written CUDA for sm_100a) replaces
 * the reference's single-threaded Julia loops for:
 *   sample!(W, Wiener)                         src/wiener.jl:24-58
 *   solve!(EulerMaruyama(), Y, u, W, P)        src/euler.jl:135-152, src/sde!.jl:21-53
 *   GuidedBridge / PartialBridge / PartialBridgeνH constructors (backward ODEs)
 *                                              src/guip.jl:172-180, src/partialbridge.jl:1-22,
 *                                              src/partialbridgenuH.jl:1-55,84-103, src/lyap.jl:2-6
 *   solve!(Euler(), Y, u, W, P°)               src/euler.jl:246-268
 *   llikelihood(LeftRule(), X, P°; skip)       src/partialbridgenuH.jl:171-189, src/guip.jl:429-446,
 *                                              src/partialbridge.jl:67-87
 *   innovations!(EulerMaruyama(), W, Y, P)     src/euler.jl:357-376
 *   the pCN / Metropolis-Hastings path update  test/partialbridgenuH.jl:155-198 (script code)
 *
 * The reference has no FFI of its own (it is pure Julia; extension is by multiple
 * dispatch).  These entry points are what a `ccall` shim adds methods for; the
 * shim itself is julia/BridgeB200.jl and INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - All floating point data is IEEE double.  Indices are 0-based in this ABI
 *     (the Julia shim converts).  Small matrices are ROW-major.
 *   - Host path layout is the reference's: one path = N consecutive states, a
 *     state = d consecutive doubles (Vector{SVector{d,Float64}} / VSamplePath's
 *     d x N column-major Matrix have exactly this byte layout).  An ensemble of
 *     np chains x S segments is [np][S][N][d].
 *   - Every function returns 0 (BB_OK) or a negative bb_status; no exception or
 *     longjmp crosses the ABI.  bb_strerror maps codes to the reference's own
 *     error strings where the reference has one.
 *   - Pointers are caller-owned unless returned by a *_create function.
 *   - One CUDA stream per bb_ctx; calls on one ctx are not re-entrant; different
 *     contexts are independent (one ctx per GPU / per host task).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point
 *     returns BB_ERR_NODEVICE.
 */
#ifndef BRIDGE_B200_H
#define BRIDGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 1
#define BB_MAXD 4   /* largest state dimension of the narrow registry models (and of bb_theta_spec) */
#define BB_MAXD_WIDE 16 /* state dimension of the wide model (LANDMARKS); constructors accept d <= 16 */
#define BB_NPAR 32  /* doubles in a model parameter block */

/* ------------------------------------------------------------------ status */
typedef enum {
  BB_OK = 0,
  BB_ERR_LENGTH = -1,      /* "Y and W differ in length."            src/euler.jl:137,251,361 */
  BB_ERR_TIMEAXIS = -2,    /* "Time axis mismatch between bridge P and driving W."  src/euler.jl:248 */
  BB_ERR_STARTPOINT = -3,  /* "Starting point has wrong length."     src/sde!.jl:30 */
  BB_ERR_DIM = -4,         /* DimensionMismatch("length(tt) != size(yy, 2)")  src/types.jl:127 */
  BB_ERR_ASSERT_M = -5,    /* AssertionError: m == length(v)         src/partialbridgenuH.jl:3 */
  BB_ERR_MODEL = -6,       /* unknown model id / model-ensemble dimension mismatch */
  BB_ERR_ARG = -7,         /* NULL pointer, negative size, bad enum */
  BB_ERR_CUDA = -8,        /* a CUDA runtime call failed; see bb_last_cuda_error */
  BB_ERR_NOMEM = -9,       /* device allocation failed */
  BB_ERR_NODEVICE = -10,   /* no CUDA device: the library has no CPU path */
  BB_ERR_UNSUPPORTED = -11,/* combination not instantiated (model x guide x dims) */
  BB_ERR_SINGULAR = -12,   /* singular matrix in a backward solve / update */
  BB_ERR_STALE = -13,      /* X holds rejected proposals; bb_ens_refresh_x recomputes the current paths */
  BB_ERR_COMM = -14,       /* an NCCL call failed; see bb_comm_last_error */
  BB_ERR_USERSRC = -15     /* the CUDA C source of a user-defined model does not compile; see bb_user_model_log */
} bb_status;

const char* bb_strerror(int status);
const char* bb_last_cuda_error(void);
int bb_abi_version(void);

/* ------------------------------------------------------------------ models
 * Target processes: the reference's user-defined b(t,x,P), sigma(t,x,P) are Julia
 * closures and cannot run on the device, so the device has a closed registry.
 * par[] layouts:
 *   WIENER    d=d'∈{1,2,3}; b=0, sigma=I                    src/wiener.jl:143-167
 *   OU        d=d'=1; par={beta, sigma}; b=-beta*x           docs/src/manual.md:44-46
 *   LINPRO    d=d'∈{1,2,3}; par={B[d*d], mu[d], sigma[d*d]}; b=B(x-mu)   src/linpro.jl:78-87
 *   FHN_DIAG  d=d'=2; par={eps,s,gamma,beta,sigma1,sigma2}   src/Models.jl:18-19
 *   FHN_HYPO  d=2,d'=1; par={eps,s,gamma,beta,sigma}; sigma=(0,sigma)'
 *                                       project_partialbridge/partialbridge_fitzhugh.jl:44-45
 *   INTDIFF   d=2,d'=1; par={gamma}; b=(x2, -(x2+sin x2)+1/2), sigma=(0,gamma)'
 *                                       test/partialbridge.jl:25-27
 *   NCLAR3    d=3,d'=1; par={alpha,omega,sigma}; b=(x2,x3,-alpha sin(omega x3))
 *                                       project_partialbridge/partialbridge_nclar.jl:58-60
 *   LORENZ    d=d'=3; par={th1,th2,th3, s1,s2,s3}; sigma=diag(s)  src/Models.jl:38-55, test/euler.jl:49-50
 *   BOLUS     d=d'=2; par={alpha,beta,lambda,mu,sigma1,sigma2}; b=(alpha dose(t) - (lambda+beta)x1 + mu x2,
 *             lambda x1 - mu x2), dose(t) = 2(t/2)/(1+(t/2)^2), sigma = sigma1 I   project_partialbridge/
 *             partialbridge_bolus3.jl:38-51,73.  The drift depends on t: this model runs on the per-chain-parameter
 *             path only (bb_theta_*, which is where the reference uses it); the shared-table kernels are autonomous.
 *   LANDMARKS d=16,d'=8; par={a, sigma, lambda}: n = 4 landmarks in the plane, state (q1,p1,...,q4,p4) with
 *             q_i, p_i in R^2 (the flattened Vector{Point} of the script), Gaussian kernel
 *             k(x) = exp(-|x|^2/(2a))/(2 pi a), drift  dq_i = 1/2 sum_j p_j k(q_i-q_j),
 *             dp_i = sum_j [-lambda/2 p_j + <p_i,p_j>(q_i-q_j)/(2a)] k(q_i-q_j), noise sigma dW on the momenta
 *                                       project_partialbridge/partialbridge_landmarks.jl:47,86-101,111-118
 *             (BASELINE config 5.  The reference script is an unfinished draft that does not run -- SURVEY 8d --
 *             so parity for this model is against the oracle restatement of these lines only.)
 *             "Wide" model: one segment-table kind (BB_GUIDE_NUH with a constant auxiliary drift), no second-pass
 *             llikelihood / innovations, plain thread-per-chain kernels (csrc/bb_wide.cuh).
 */
typedef enum {
  BB_MODEL_WIENER = 0,
  BB_MODEL_OU = 1,
  BB_MODEL_LINPRO = 2,
  BB_MODEL_FHN_DIAG = 3,
  BB_MODEL_FHN_HYPO = 4,
  BB_MODEL_INTDIFF = 5,
  BB_MODEL_NCLAR3 = 6,
  BB_MODEL_LORENZ = 7,
  BB_MODEL_LANDMARKS = 8,
  BB_MODEL_BOLUS = 9,
  BB_MODEL_USER = 10,  /* drift and diffusion given as CUDA C source, compiled at run time: bb_user_model_create */
  BB_MODEL_COUNT = 11
} bb_model_id;

typedef struct {
  int32_t id;       /* bb_model_id */
  int32_t d;        /* state dimension */
  int32_t dprime;   /* dimension of the driving Wiener process */
  int32_t reserved; /* BB_MODEL_USER: bb_user_model_handle(); 0 otherwise */
  double par[BB_NPAR];
} bb_model;

/* ------------------------------------------------------------------ auxiliary process
 * The linear auxiliary process  dX~ = (B~(t) X~ + beta~(t)) dt + sigma~(t) dW  of a guided
 * proposal (Bridge.B(t,Pt), Bridge.β(t,Pt), Bridge.a(t,Pt); src/linpro.jl:78-86,
 * project_partialbridge/partialbridge_fitzhugh.jl:99-116).  It is passed as VALUES, never
 * as code: either one constant triple, or the values at the three Ralston stage times of
 * every grid interval (the shim evaluates the user's closures there).
 *   interval i (0 <= i < N-1) is [tt[i], tt[i+1]]; the backward step over it starts at
 *   t = tt[i+1] with h = tt[i]-tt[i+1] < 0 and evaluates at t, t+h/2, t+3h/4
 *   (src/ode.jl:44-49,92-95)  ->  stage k of interval i is entry 3*i+k.
 */
typedef struct {
  int32_t d;
  int32_t is_const;     /* 1: single value each; 0: staged arrays */
  const double* B;      /* [d*d]  or [(N-1)*3][d*d] */
  const double* beta;   /* [d]    or [(N-1)*3][d]   */
  const double* a;      /* [d*d]  or [(N-1)*3][d*d] */
  const double* a_left; /* staged only: a~(tt[i]) per interval, [(N-1)][d*d] (Lyapunov step, src/lyap.jl:5) */
} bb_aux;

/* ------------------------------------------------------------------ opaque handles */
typedef struct bb_ctx bb_ctx;     /* device + stream */
typedef struct bb_ens bb_ens;     /* device-resident ensemble of chains */
typedef struct bb_guide bb_guide; /* device-resident guiding tables of one segment */

/* ------------------------------------------------------------------ context */
int bb_ctx_create(int device, bb_ctx** out);
int bb_ctx_destroy(bb_ctx* ctx);
int bb_ctx_synchronize(bb_ctx* ctx);
/* adopt an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
int bb_ctx_set_stream(bb_ctx* ctx, void* cuda_stream);
void* bb_ctx_get_stream(bb_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t bb_ctx_launch_count(bb_ctx* ctx);
/* elapsed device time (ms) of the most recent compute call, measured with CUDA
 * events on the context's stream; valid after bb_ctx_synchronize */
int bb_ctx_set_timing(bb_ctx* ctx, int enabled);
double bb_ctx_last_kernel_ms(bb_ctx* ctx);
/* Rounding order of the shared-table proposal constructors (bb_update_nuHC, bb_gpupdate_*, bb_backward_*, d <= 3):
 *   BB_ARITH_REFERENCE (default)  plain products and sums, nothing fused: the operations the reference performs
 *                                 (Julia does not contract a*b+c; StaticArrays' unrolled products), so the tables are
 *                                 bit-identical to the CPU restatement in reference arithmetic (liboracle_ref.so).
 *                                 The backward recursions are ill-conditioned at the scripts' settings (Σ = 1e-10,
 *                                 ϵ = 1e-3): fused and unfused tables differ by 4e-8 relative, which moves ll by 2e-6.
 *   BB_ARITH_FUSED                explicit fused multiply-adds, the rounding order of the per-chain-parameter
 *                                 kernels (bb_theta_*), where the fp64 pipe is the limit; with it the shared-table
 *                                 constructors reproduce the per-chain tables bit for bit. */
enum { BB_ARITH_REFERENCE = 0, BB_ARITH_FUSED = 1 };
int bb_ctx_set_arith(bb_ctx* ctx, int arith);
/* Which kernel runs bb_pcn_step with X° stored for scalar-noise models: one thread per chain, or the warp-specialised
 * kernel (noise warps + dynamics warps, two threads per chain), or its variant with two chains per dynamics thread
 * (d <= 2).  Results are identical bit for bit; AUTO (default) takes the warp-specialised kernel for small ensembles
 * (fewer than two 128-chain CTAs per SM), where it is 5-8 % faster; the variant is never chosen automatically. */
enum { BB_PCN_AUTO = 0, BB_PCN_ONE_THREAD = 1, BB_PCN_WARP_SPECIALISED = 2, BB_PCN_WARP_SPECIALISED_2 = 3 /* two chains per dynamics thread */ };
int bb_ctx_set_pcn_kernel(bb_ctx* ctx, int mode);

/* ------------------------------------------------------------------ user-defined target processes
 * The reference's extension point is Julia dispatch: a user defines Bridge.b(t, x, P), Bridge.σ(t, x, P) for an own struct
 * (src/types.jl:23,32-33, src/Bridge.jl:105-106; project_partialbridge/partialbridge_fitzhugh.jl:44-46).  Julia closures
 * cannot run on the device, so the equivalent here is CUDA C source compiled at run time (NVRTC, bound by dlopen; the
 * kernel headers are embedded in the library): the path kernel (sample!, solve!, guided solve! + llikelihood, pCN, the
 * second-pass llikelihood, StochasticHeun) is instantiated for the user's model exactly as for a registry model --
 * same TMA / cp.async staging, same arithmetic contract (-fmad=false: only explicit fma() fuses).
 *   drift_src   statements of   void b(const double* par, const double* x, double* o)   -- o[0..d-1] = b(x; par);
 *               autonomous (no t), double precision, e.g. the FitzHugh-Nagumo drift of partialbridge_fitzhugh.jl:44:
 *                 "double u = x[0] - x[1]; u = fma(-(x[0]*x[0]), x[0], u); o[0] = (u + par[1]) * (1.0/par[0]);"
 *                 "o[1] = fma(par[2], x[0], -x[1]) + par[3];"
 *   col[i]      column of the driving Wiener process that enters component i, or -1 (each row of σ has at most one
 *               non-zero: scalar, UniformScaling, SDiagonal or column-vector σ, as every registry model but LinPro)
 *   sigma_src[i] expression in par[] for that entry of σ (ignored where col[i] < 0); constant in x
 * par[] is the bb_model's parameter block, passed with every call (so parameters change without recompiling).
 * bb_model { id = BB_MODEL_USER, d, dprime, reserved = bb_user_model_handle(um), par } then works wherever a registry
 * model does, except innovations!, the per-chain-parameter path (bb_theta_*) and bb_pcn_step_host.  Kernels are
 * compiled on first use of a (guide kind, mode) combination, ~1-2 s each; bb_user_model_log returns NVRTC's log. */
typedef struct bb_user_model bb_user_model;
int bb_user_model_create(bb_ctx* ctx, int32_t d, int32_t dprime, const char* drift_src, const int32_t* col,
                         const char* const* sigma_src, bb_user_model** out);
/* compile-only check (no device needed): builds the kernel of the given combination (gk: 0 or bb_guide_kind; gm: rows of L
 * for LMMU; auxm: 1 constant / 0 tabulated / 2 non-constdiff auxiliary; rng: 0 read W, 1 pCN, 2 sample, 3 pCN without X°,
 * 10 llikelihood, 12 StochasticHeun) and returns the cubin size (> 0) or a negative status; NVRTC's log -> log */
int bb_user_source_check(int32_t d, int32_t dprime, const char* drift_src, const int32_t* col,
                         const char* const* sigma_src, int32_t gk, int32_t gm, int32_t auxm, int32_t rng, char* log,
                         int32_t log_size);
int32_t bb_user_model_handle(bb_user_model* um);
const char* bb_user_model_log(bb_user_model* um);
int bb_user_model_destroy(bb_user_model* um);


/* ------------------------------------------------------------------ ensemble
 * P chains, each a concatenation of S segments of N grid points (N-1 Euler steps);
 * segment s+1 starts at the end point of segment s (bolus3.jl:187-192).  S=1 is the
 * plain ensemble of independent paths.  Replaces P*S reference SamplePath pairs (W, X)
 * (src/types.jl:71-76) and, with BB_ENS_DOUBLE_BUFFER, their proposal copies (Wo, Xo)
 * (test/partialbridgenuH.jl:168-170).
 */
enum {
  BB_ENS_DOUBLE_BUFFER = 1u, /* allocate the proposal buffer of W (needed by bb_pcn_step) */
  BB_ENS_NO_X = 2u           /* do not allocate X (paths are never stored) */
};
int bb_ens_create(bb_ctx* ctx, int64_t P, int32_t S, int32_t N, int32_t d, int32_t dprime,
                  uint32_t flags, bb_ens** out);
int bb_ens_destroy(bb_ens* ens);
/* chain ids used for the random streams are  chain_offset + local index  (multi-GPU sharding) */
int bb_ens_set_chain_offset(bb_ens* ens, int64_t chain_offset);
/* time grid of segment seg; n must equal N (else BB_ERR_LENGTH) */
int bb_ens_set_grid(bb_ens* ens, int32_t seg, const double* tt, int32_t n);
int bb_ens_get_grid(bb_ens* ens, int32_t seg, double* tt, int32_t n);
/* starting points: u is [d] (broadcast != 0) or [P][d] */
int bb_ens_set_start(bb_ens* ens, const double* u, int32_t n_u, int32_t broadcast);

enum { BB_W = 0, BB_X = 1 };       /* which array */
enum { BB_CUR = 0, BB_PROP = 1 };  /* current state of each chain, or its last proposal */
/* host layout [np][S][N][k], k = dprime (W) or d (X); chains p0 .. p0+np-1 */
int bb_ens_upload(bb_ens* ens, int what, int which, int64_t p0, int64_t np, const double* host);
int bb_ens_download(bb_ens* ens, int what, int which, int64_t p0, int64_t np, double* host);

enum {
  BB_F_LL = 0,       /* double[P]  log-likelihood of the current path (sum over segments) */
  BB_F_LL_PROP = 1,  /* double[P]  log-likelihood of the last proposal */
  BB_F_LOGU = 2,     /* double[P]  log(U) drawn for the last accept test */
  BB_F_XEND = 3,     /* double[P][d] end point yy[N] of the last solve (current) */
  BB_F_XEND_PROP = 4 /* double[P][d] end point of the last proposal */
};
int bb_ens_get_f64(bb_ens* ens, int field, int64_t p0, int64_t np, double* host);
int bb_ens_set_ll(bb_ens* ens, int64_t p0, int64_t np, const double* host);
/* uint8[P]: 1 if the last bb_pcn_step accepted the chain's proposal */
int bb_ens_get_accepted(bb_ens* ens, int64_t p0, int64_t np, uint8_t* host);
/* number of accepted proposals since creation / last reset, summed over this ensemble's chains */
int bb_ens_get_acc(bb_ens* ens, int64_t* acc);
int bb_ens_reset_acc(bb_ens* ens);
/* device address of the int64 acceptance counter (for an NCCL all-reduce issued by the host) */
void* bb_ens_acc_device_ptr(bb_ens* ens);
int64_t bb_ens_bytes(bb_ens* ens); /* device bytes held */

/* ------------------------------------------------------------------ a4: sample!(W, Wiener{T}())
 * W_cur[row][0] is kept (y1 = W.yy[1], src/wiener.jl:50-51); W[j] = W[j-1] + sqrt(tt[j]-tt[j-1]) xi.
 * xi ~ N(0,1): Philox4x32-10 keyed by seed, counter (pair index, stream, global row id), Box-Muller
 * in double.  `stream` separates independent draws with the same seed (e.g. MCMC iteration).
 */
int bb_wiener_sample(bb_ens* ens, uint64_t seed, uint32_t stream);

/* ------------------------------------------------------------------ a5/a6: solve!(EulerMaruyama(), X, u, W, P)
 * X_cur <- Euler-Maruyama solution driven by W_cur from the ensemble's starting points; segment
 * s+1 continues from the end point of segment s.  src/euler.jl:135-152.
 */
int bb_euler(bb_ens* ens, const bb_model* model);
/* The reference's other SDE schemes for a plain (unguided) target, S = 1:
 *   BB_SCHEME_EULER         solve!(EulerMaruyama(), ...)                          src/euler.jl:135-152  (= bb_euler)
 *   BB_SCHEME_STRATONOVICH  solve!(StratonovichEuler(), ...)  src/euler.jl:68-88: y + b dt + ½(σ(y^E) + σ(y)) dw.  Every
 *                           registry model has a constant σ, for which ½(σ + σ) = σ exactly: the Euler-Maruyama kernel.
 *   BB_SCHEME_SRK           solve!(StochasticRungeKutta(), ...) src/euler.jl:330-356 (scalar processes): its correction
 *                           ½(σ(ups) - σ(y))(dw² - δ)/√δ is exactly 0 for constant σ: the Euler-Maruyama kernel, d = 1.
 *   BB_SCHEME_HEUN          solve!(StochasticHeun(), ...)     src/euler.jl:178-198: drift by Heun's rule; as in the
 *                           reference the loop stops at N-2 and yy[N] is left untouched.
 * Mdb (src/euler.jl:308-327) needs a proposal process with its own time axis: bb_guided_mdb runs it on a guided proposal;
 * through this entry point (a plain target) it is BB_ERR_UNSUPPORTED. */
enum { BB_SCHEME_EULER = 0, BB_SCHEME_STRATONOVICH = 1, BB_SCHEME_HEUN = 2, BB_SCHEME_SRK = 3, BB_SCHEME_MDB = 4 };
int bb_solve_scheme(bb_ens* ens, const bb_model* model, int32_t scheme);
/* solve!(Mdb(), Y, u, W, P°)  src/euler.jl:308-327 for a guided proposal P° (the scheme needs P.tt and the indexed drift
 * _b((i,t), x, P), which GuidedBridge / PartialBridge / PartialBridgeνH provide), S = 1: the guided Euler step with the noise
 * scaled by sqrt((tt[end] - tt[i+1]) / (tt[end] - tt[i])).  X_cur <- the path, BB_F_XEND <- yy[N]. */
int bb_guided_mdb(bb_ens* ens, const bb_model* model, bb_guide* const* guides);
/* fused sample! + solve! (W and X are both written, W is never read) */
int bb_sample_euler(bb_ens* ens, const bb_model* model, uint64_t seed, uint32_t stream);

/* ------------------------------------------------------------------ guiding tables
 * Values on the time grid of one segment, as the reference's proposal structs hold them
 * (src/partialbridgenuH.jl:122-130, src/guip.jl:165-170, src/partialbridge.jl:33-41), plus
 * the auxiliary drift B~(tt[i]), beta~(tt[i]) that llikelihood evaluates (b~ = B~ x + beta~).
 */
typedef enum {
  BB_GUIDE_NUH = 1,  /* PartialBridgeνH: r = H[i](nu[i]-x)                 partialbridgenuH.jl:157-161 */
  BB_GUIDE_HV = 2,   /* GuidedBridge:    r = H♢[i] \ (V[i]-x)               guip.jl:192-194 */
  BB_GUIDE_LMMU = 3  /* PartialBridge:   r = L[i]'M[i](v-mu[i]-L[i]x)       partialbridge.jl:53-58 */
} bb_guide_kind;

/*  kind   A                 b               Mm            v
 *  NUH    H   [N][d][d]     nu [N][d]       NULL          NULL
 *  HV     H♢  [N][d][d]     V  [N][d]       NULL          NULL
 *  LMMU   L   [N][m][d]     mu [N][m]       M [N][m][m]   v [m]
 *  Bt, betat: [d*d],[d] if aux_const != 0, else [N][d*d],[N][d].
 */
int bb_guide_create(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                    const double* A, const double* b, const double* Mm, const double* v,
                    const double* Bt, const double* betat, int32_t aux_const, bb_guide** out);
/* The same for a pair with a(t,x,Target) != a~(t) (constdiff(P°) == false): Adiff = a - a~ on the grid ([d*d] if
 * adiff_const != 0, else [N][d*d]) adds the terms  -1/2 tr((a-a~)H) dt + 1/2 r'(a-a~)r dt  of
 * src/partialbridge.jl:79-84 to the log-likelihood (H = H[i], inv(H♢[i]) or L[i]'M[i]L[i]). */
int bb_guide_create_ncd(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                        const double* A, const double* b, const double* Mm, const double* v, const double* Bt,
                        const double* betat, int32_t aux_const, const double* Adiff, int32_t adiff_const,
                        bb_guide** out);
int bb_guide_destroy(bb_guide* g);

/* ------------------------------------------------------------------ a7-a10: backward ODEs (constructors)
 * All run on the device (one thread per system; they are O(N d^3), once per segment).
 */
/* updateνH⁺C(L, Σ, v, ϵ)  src/partialbridgenuH.jl:1-17.  L [m][d], Sigma [m][m], v [m]. */
int bb_update_nuHC(bb_ctx* ctx, int32_t d, int32_t m, const double* L, const double* Sigma,
                   const double* v, double eps, double* nu, double* Hplus, double* C);
/* observation update of (ν, H⁺) between segments (partialbridge_bolus3.jl:128-137):
 * Z = I - H⁺L'(Σ + LH⁺L')⁻¹L;  ν <- Z H⁺L'Σ⁻¹v + Zν;  H⁺ <- Z H⁺   (in place) */
int bb_gpupdate_nuH(bb_ctx* ctx, int32_t d, int32_t m, double* nu, double* Hplus,
                    const double* L, const double* Sigma, const double* v);
/* Bridge.gpupdate(H♢, V, L, Σ, v)  src/guip.jl:221-243 (in place; handles H♢ = Inf diag) */
int bb_gpupdate_HV(bb_ctx* ctx, int32_t d, int32_t m, double* Hdia, double* V,
                   const double* L, const double* Sigma, const double* v);

enum { BB_ODE_R3 = 0, BB_ODE_LYAP = 1 };
/* partialbridgeodeνH!(R3()/Lyap(), tt, νt, Ht, Pt, (ν, H⁺, C))  src/partialbridgenuH.jl:21-55,86-103.
 * in : nu_end [d], Hplus_end [d][d], C0
 * out: nu [N][d], H [N][d][d];  nu_left [d], Hplus_left [d][d], C  at tt[0] (for chaining). */
int bb_backward_nuH(bb_ctx* ctx, int32_t method, int32_t N, int32_t d, const double* tt,
                    const bb_aux* aux, const double* nu_end, const double* Hplus_end, double C0,
                    double* nu, double* H, double* nu_left, double* Hplus_left, double* C);
/* partialbridgeodeHνH!(R3(), tt, Ft, Ht, Pt, (F, H, C))  src/partialbridgenuH.jl:64-81 */
int bb_backward_FH(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const bb_aux* aux,
                   const double* F_end, const double* H_end, double C0,
                   double* F, double* H, double* C);
/* gpHinv!/gpV! of the GuidedBridge constructor  src/guip.jl:172-180, src/gode.jl:2-3,13,21.
 * in: v [d], hdia_end [d][d] (NULL = zero);  out: Hdia [N][d][d], V [N][d] */
int bb_backward_HV(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const bb_aux* aux,
                   const double* v, const double* hdia_end, double* Hdia, double* V);
/* partialbridgeode!(R3(), t, L, Σ, Lt, Mt, μt, P)  src/partialbridge.jl:1-22.
 * out: Lt [N][m][d], Mt [N][m][m] (= inv(M⁺)), mut [N][m] */
int bb_backward_LMmu(bb_ctx* ctx, int32_t N, int32_t d, int32_t m, const double* tt,
                     const bb_aux* aux, const double* L, const double* Sigma,
                     double* Lt, double* Mt, double* mut);

/* The backward pass of a CHAIN of S PartialBridgeνH segments in ONE launch (the script loop of
 * partialbridge_bolus3.jl:162-180: ν = 0, H⁺ = I/ϵ right of the last observation; gpupdate with v[S-1]; for s = S-1 .. 0:
 * partialbridgeνH over tt[s] -- `method` BB_ODE_LYAP as the script, or BB_ODE_R3 -- and, for s > 0, gpupdate with v[s-1]).
 * The tables are written in place on the device, where the path kernels read them: guides[s] is updated if it exists
 * (same N, d, kind NUH, constant auxiliary drift) or created if NULL.  Nothing returns to the host but the left-end
 * values -- this is what a sampler that re-proposes parameters calls between two bb_pcn_step launches.
 * aux[s]: the constant auxiliary process of segment s; v [S][m]; tt [S][N].  d <= 3.
 * bb_guide_download_nuH reads ν[i], H[i] (i < N-1; the terminal values are not kept: NaN) back from a guide. */
int bb_guides_chain_nuH(bb_ctx* ctx, int32_t method, int32_t S, int32_t N, int32_t d, int32_t m, const double* tt,
                        const bb_aux* aux, const double* L, const double* Sigma, const double* v, double eps,
                        bb_guide** guides, double* nu_left, double* Hplus_left, double* C);
int bb_guide_download_nuH(bb_guide* g, double* nu, double* H);

/* lptilde: log of the auxiliary process' transition density at the left end of a proposal (the p~ of every importance
 * weight exp(ll) p~/p, test/guip.jl:245-274).
 *   bb_lptilde_nuH   lptilde(x, P::PartialBridgeνH) = -1/2 (x'H[1]x - 2x'H[1]ν[1]) - C, the formula the reference TESTS
 *                    (test/partialbridgenuH.jl:124; src/partialbridgenuH.jl:169 itself has a typo and does not run).
 *   bb_lptilde_HV    lptilde(P::GuidedBridge, u) = logpdfnormal(V[1] - u, H♢[1]) - traceB(tt, Pt)   src/guip.jl:203-206,
 *                    src/gaussian.jl:66-75; traceB = forward Ralston integral of tr B~(t) (src/ode.jl:178-184): trB holds
 *                    tr B~ at the stage times tt[i], tt[i]+h/2, tt[i]+3h/4 of every interval ([(N-1)*3]) or, with
 *                    trB_const != 0, the one constant value. */
int bb_lptilde_nuH(bb_ctx* ctx, int32_t d, const double* nu0, const double* H0, double C, const double* x, double* out);
int bb_lptilde_HV(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const double* trB, int32_t trB_const,
                  const double* V0, const double* Hdia0, const double* u, double* out);

/* ------------------------------------------------------------------ a11-a13: guided Euler + Girsanov ll
 * solve!(Euler(), X, u, W, P°) fused with llikelihood(LeftRule(), X, P°; skip):
 * X_cur <- guided Euler path driven by W_cur, ll_cur <- sum over segments of the log-likelihood,
 * xend <- yy[N] of the last segment.  guides: one table per segment.  src/euler.jl:246-268.
 */
enum { BB_RUN_STORE_X = 1u, BB_RUN_NO_LL = 2u, BB_RUN_SKIP_REJECTED = 4u /* bb_pcn_step_host only, see there */ };
int bb_guided_euler_ll(bb_ens* ens, const bb_model* model, bb_guide* const* guides,
                       int32_t skip, uint32_t flags);
/* llikelihood(LeftRule(), X, P°; skip) on the stored X_cur -> ll_cur  (second-pass form) */
int bb_llikelihood(bb_ens* ens, const bb_model* model, bb_guide* const* guides, int32_t skip);
/* innovations!(EulerMaruyama(), W, X, P): W_cur <- sigma^{-1}-increments of X_cur  src/euler.jl:357-376.
 * guides == NULL: unguided drift.  Needs an invertible sigma (d' = d). */
int bb_innovations(bb_ens* ens, const bb_model* model, bb_guide* const* guides);

/* ------------------------------------------------------------------ a15: one pCN / MH iteration for every chain
 * W2 ~ Wiener;  W° = rho W + sqrt(1-rho^2) W2 (cumulative values);  X° = guided Euler(W°);
 * ll° = llikelihood(X°);  accept iff log(U) <= ll° - ll;  on accept the chain's buffers swap
 * roles (no copy), ll <- ll°, acc += 1.   test/partialbridgenuH.jl:176-191.
 * All S segments of a chain are updated jointly with one accept (bolus3.jl:300-355 with
 * ind = all segments and a fixed starting point).  flags: BB_RUN_STORE_X keeps X°.
 */
int bb_pcn_step(bb_ens* ens, const bb_model* model, bb_guide* const* guides, double rho,
                uint64_t seed, uint32_t iter, int32_t skip, uint32_t flags);
/* The same iteration on HOST buffers, for callers that keep W, X in host memory as the reference loop does
 * (test/partialbridgenuH.jl:168-191): W_host [P][S][N][d'] (current W of every chain) goes up; the proposal
 * Wo_host [P][S][N][d'], Xo_host [P][S][N][d] (may be NULL), llo_host [P] (may be NULL) and accepted_host [P]
 * (may be NULL) come back.  Slabs of chains are pipelined over three streams so that both PCIe directions and
 * the GPU are busy at once; pinned host memory is needed for that overlap (pageable memory works, serially).
 * The ensemble's device state is updated exactly as by bb_pcn_step.
 * flags | BB_RUN_SKIP_REJECTED: the reference loop reads Wo / Xo only to swap them in on accept (X, Xo = Xo, X;
 * W, Wo = Wo, W, test/partialbridgenuH.jl:184-186) and overwrites them in the next iteration otherwise, so the rows of
 * chains that REJECT need not cross the link: with the flag their rows of Wo_host / Xo_host are left as they were.  When
 * Wo_host / Xo_host are page-locked, device-mapped memory (cudaHostAlloc / cudaHostRegister under unified addressing) the
 * accepted rows are written straight into them by a kernel, without staging, and the device-to-host traffic drops by the
 * rejection rate; with other memory the flag is accepted and everything is copied as without it.  On that direct path
 * Wo_host may be W_host itself: the array is then updated in place for the chains that accept -- the loop's swap of W and
 * Wo -- and always holds the current W of every chain (likewise Xo_host the current X once it has been initialised);
 * aliasing without the direct path is BB_ERR_ARG. */
int bb_pcn_step_host(bb_ens* ens, const bb_model* model, bb_guide* const* guides, double rho, uint64_t seed,
                     uint32_t iter, int32_t skip, uint32_t flags, const double* W_host, double* Wo_host,
                     double* Xo_host, double* llo_host, uint8_t* accepted_host);
/* X is kept ONCE per chain and always holds the path of the chain's last proposal X° (the reference's Xo);
 * for a chain whose proposal was rejected the reference's X (current path) is a pure function of the
 * chain's current W, and this call recomputes it in place for exactly those chains (solve!(Euler(), X, x0,
 * W, P°) again).  bb_ens_download(BB_X, BB_CUR) returns BB_ERR_STALE until it has been called; BB_X/BB_PROP
 * downloads need no refresh.  No-op if nothing is stale. */
int bb_ens_refresh_x(bb_ens* ens, const bb_model* model, bb_guide* const* guides);

/* ------------------------------------------------------------------ online statistics
 * mcstart / mcnext! / mcstats (src/mclog.jl:22-56, 88-93) for an ensemble: first and second moments of the chains'
 * CURRENT paths per grid point, pooled over the P chains and over calls (n = calls x P samples).  bb_ens_mc_update
 * returns BB_ERR_STALE while X holds rejected proposals (bb_ens_refresh_x first).
 * bb_ens_mc_stats: mean [S][N][d], cov = m2/(n-1) [S][N][d][d]; either pointer may be NULL. */
int bb_ens_mc_reset(bb_ens* ens);
int bb_ens_mc_update(bb_ens* ens);
int bb_ens_mc_stats(bb_ens* ens, double* mean, double* cov, int64_t* n);
/* The same with the reference's own semantics -- one state (m, m2, k) PER CHAIN, over the recorded iterations, as the
 * scripts keep it: `mcstate = [mcnext!(mcstate[i], XX[i].yy) for i in ...]`, partialbridge_fitzhugh.jl:169-189:
 *   bb_ens_chain_mc_reset   mcstart (src/mclog.jl:22-24): m = 0, m2 = 0, k = 0.  The state takes (1 + d) times the memory
 *                           of X and is allocated on first use.
 *   bb_ens_chain_mc_update  mcnext! (src/mclog.jl:47-56) for every chain with its CURRENT path, Welford's update in the
 *                           reference's operation order (delta = x - m; m += delta/(k+1); m2 += delta (x - m)'): bit-identical
 *                           to the reference recurrence, deterministic.  BB_ERR_STALE as for bb_ens_mc_update.  d <= 3.
 *   bb_ens_chain_mc_stats   mcstats (src/mclog.jl:88-93) of chains p0 .. p0+np-1: mean [np][S][N][d], cov = m2/(k-1)
 *                           [np][S][N][d][d] (either may be NULL); *k = number of updates.
 *   bb_ens_chain_mc_band    mcband (src/mclog.jl:75-85): m -/+ Q sqrt(diag(m2) (1/(k-1))), Q = sqrt(2.) erfinv(0.95);
 *                           lower, upper [np][S][N][d]. */
#define BB_MCBAND_Q 1.9599639845400538 /* sqrt(2.) * erfinv(0.95) evaluated in double precision */
int bb_ens_chain_mc_reset(bb_ens* ens);
int bb_ens_chain_mc_update(bb_ens* ens);
int bb_ens_chain_mc_stats(bb_ens* ens, int64_t p0, int64_t np, double* mean, double* cov, int64_t* k);
int bb_ens_chain_mc_band(bb_ens* ens, int64_t p0, int64_t np, double* lower, double* upper);

/* ------------------------------------------------------------------ per-chain parameters
 * Every chain p carries its own parameter vector θ_p of the target model (the leading BB_NTHETA entries of the
 * model's par[] block) and therefore its own guiding tables.  This is the parameter-update branch of the script
 * loop  project_partialbridge/partialbridge_bolus3.jl:248-365  (`updateparams == true`) for P independent chains:
 *
 *   θ° = θ + rw_sd .* ξ                                    propose(σ, P)                         bolus3.jl:239-242
 *   (ν, H⁺) right of the last observation: ν = 0, H⁺ = I/ϵ, then gpupdate with v[S-1]             :162-165
 *   for s = S-1 .. 0:  partialbridgeνH(tt_s, P°, Pt°, ν, H⁺)  (Lyapunov backward step,           :276-283,
 *                      src/partialbridgenuH.jl:86-103,148-155, src/lyap.jl:2-6);
 *                      if s > 0: gpupdate(ν, H⁺, Σ, L, v[s-1])                                    :284-291, :128-137
 *   W° = W (innovations held fixed)                                                               :306
 *   X° = solve!(Euler(), x0, W, Q°) over all segments;  ll° = Σ_s llikelihood(LeftRule(), X°_s, Q°_s)   :324-333
 *   diffll = logpdfnormal(x0 - ν°(0), symmetrize(H⁺°(0))) - logpdfnormal(x0 - ν(0), symmetrize(H⁺(0)))   :319
 *          + ll° - ll  + Σ_s (t_s,end - t_s,0) (tr B~°_s - tr B~_s) + logπ(θ°) - logπ(θ)          :331-336
 *   accept iff log(U) <= diffll:  θ <- θ°, tables, X, ll follow                                   :340-355
 *
 * The guiding tables of ALL chains are built on the device (one thread per chain) into a table array
 * T [S][N][d+d*d][P] (ν[i], H[i] per grid point, chain-minor: warp accesses are contiguous), so that no
 * θ-proposal round-trips to the host.  The auxiliary process follows from (θ, v_s) by a closed registry
 * (bb_aux_kind), for the same reason the target models do.  The script's joint random walk on the starting point
 * (:311-318) is available through start_sd / start_dir of bb_theta_spec; its blocked segment updates (klow..kup) are
 * bb_theta_block_step below.
 */
#define BB_NTHETA 8
typedef enum {
  BB_AUX_FHN_MATCHING = 1,      /* B~ = [1/ϵ -1/ϵ; γ -1], β~ = (s/ϵ - v³/ϵ, β)           partialbridge_fitzhugh.jl:106-108 */
  BB_AUX_FHN_LINEARISED_END = 2,/* B~ = [1/ϵ - 3v²/ϵ  -1/ϵ; γ -1], β~ = (s/ϵ + 2v³/ϵ, β)  partialbridge_fitzhugh.jl:98-100 */
  BB_AUX_BOLUS = 3              /* DiffusionAux: B~ = [-λ-β μ; λ -μ], β~(t) = (α dose(t), 0), a~ = diag(σ1², σ2²)  bolus3.jl:54-71
                                   (time dependent: evaluated on the device at the Ralston stage times) */
} bb_aux_kind;
enum { BB_PRIOR_FLAT = 0, BB_PRIOR_GAMMA = 1 /* Gamma(shape a, scale b): logπ of bolus3.jl:237 */ };
typedef struct {
  int32_t m;                         /* rows of L */
  int32_t aux_kind;                  /* bb_aux_kind.  Every registry pair declares constdiff = true (bolus3.jl:54,70;
                                        partialbridge_fitzhugh.jl:46,113), so no a != a~ terms enter ll -- also for
                                        BB_AUX_BOLUS, whose a~ = diag(σ1², σ2²) differs from a = σ1² I, as in the reference */
  double L[BB_MAXD * BB_MAXD];       /* m x d, row-major */
  double Sigma[BB_MAXD * BB_MAXD];   /* m x m */
  double eps;                        /* H⁺ = I/eps right of the last observation */
  double v[16][BB_MAXD];             /* v[s][0..m-1]: observation at the right end of segment s */
  int32_t prior_kind[BB_NTHETA];
  double prior_a[BB_NTHETA], prior_b[BB_NTHETA];
  /* joint update of the starting point in a parameter step (bolus3.jl:311-318): with probability 1/2
   * x0° = x0 + start_sd * u * start_dir, u ~ N(0,1) (the script: 0.1 * (u, -u)); start_sd = 0 switches it off.
   * Random numbers: u is normal 3 of the proposal quad (so at most 3 parameters are updated), the coin is bit 0 of
   * word 0 of Philox counter (0xFFFFFFFC, iter, chain). */
  double start_sd;
  double start_dir[BB_MAXD];
} bb_theta_spec;

/* allocate θ (all chains start at model->par), the table array T and the per-chain left-end values */
int bb_theta_attach(bb_ens* ens, const bb_model* model, const bb_theta_spec* spec);
/* theta: [np][BB_NTHETA] */
int bb_theta_set(bb_ens* ens, int64_t p0, int64_t np, const double* theta);
int bb_theta_get(bb_ens* ens, int which /* BB_CUR | BB_PROP */, int64_t p0, int64_t np, double* theta);
/* starting points of the chains (current) or of the last parameter proposal: x0 [np][d] */
int bb_theta_get_start(bb_ens* ens, int which, int64_t p0, int64_t np, double* x0);
/* backward pass for the CURRENT θ of every chain -> T, left-end values */
int bb_theta_guides(bb_ens* ens);
/* left-end values of the last backward pass: out [np][d + d*d + 4] = ν(0), H⁺(0), C,
 * logpdfnormal(x0 - ν(0), H⁺(0)), Σ_s (t_s,end - t_s,0) tr B~_s, logπ(θ) */
int bb_theta_get_left(bb_ens* ens, int which, int64_t p0, int64_t np, double* out);
/* tables of one chain as the reference holds them: nu [S][N][d], H [S][N][d][d] */
int bb_theta_get_tables(bb_ens* ens, int64_t p, double* nu, double* H);
/* solve!(Euler(), X, x0, W, Q_p) + llikelihood with each chain's own θ and tables (initialisation, bolus3.jl:186-192) */
int bb_theta_guided_euler_ll(bb_ens* ens, int32_t skip, uint32_t flags);
/* one pCN iteration (as bb_pcn_step) with each chain's own θ and tables */
int bb_theta_pcn_step(bb_ens* ens, double rho, uint64_t seed, uint32_t iter, int32_t skip, uint32_t flags);
/* one parameter-update MH iteration; rw_sd [BB_NTHETA] (0 = parameter not updated, at most 4 non-zero).
 * Random numbers: Philox counter (0xFFFFFFFE, iter, chain) for ξ, (0xFFFFFFFD, iter, chain) for U. */
int bb_theta_param_step(bb_ens* ens, const double* rw_sd, uint64_t seed, uint32_t iter, int32_t skip,
                        uint32_t flags);
/* bb_ens_refresh_x for chains with their own tables: X <- current path of the chains whose last proposal (pCN or
 * parameter) was rejected */
int bb_theta_refresh_x(bb_ens* ens);
/* Blocked path update: the `updateparams == false` branch of the same loop (bolus3.jl:258-275, :300-355) for ONE block
 * of segments s_lo .. s_hi-1 (0-based; the script's ind = (kup-1):-1:klow with klow = s_lo+1, kup = s_hi+1), every
 * chain on the same block -- the host draws the block sequence (segnum_update = sample(1:obsnum-klow), :266) and calls
 * once per block:
 *   right end (:272-275): the block ends the chain (s_hi == S): ν = 0, H⁺ = I/ϵ, gpupdate with v[S-1] (νrightmost,
 *             Hrightmost⁺, :162-167); else ν = XX[s_hi-1].yy[end] -- the CHAIN'S OWN current path -- and
 *             H⁺ = Hzero⁺ = hzero·I (:234, 0.1 in the script);
 *   tables    for s = s_hi-1 .. s_lo: partialbridgeνH on the segment, gpupdate with v[s-1] unless s == s_lo (:281-296);
 *   noise     W°[s] = ρ W[s] + sqrt(1-ρ²) W2, W2 ~ Wiener from noise row chain·S + s (:304-305; as bb_pcn_step);
 *   start     s_lo == 0: x0° = x0 + (start_sd u) start_dir, ALWAYS proposed in this branch (:316-318; u = normal 3 of
 *             Philox counter (0xFFFFFFFE, iter, chain); start_sd = 0: x0° = x0) with the start term
 *             logpdfnormal(x0° - ν(0), symmetrize(H⁺(0))) - logpdfnormal(x0 - ν(0), ·) (:319);
 *             s_lo > 0: both passes start at the current path's value at the block's left end, XX[s_lo-1].yy[end].
 *             (The script reads XXtemp[i-1] / XXᵒ[i-1] there, :321-322 -- buffers of the PREVIOUS block's proposal
 *             that, after its swap or rejection, no longer hold the current path; a chain of segments that starts
 *             where the current path is, is what the comment on :260 describes and what is built here.  Likewise
 *             `ind = 1:1` for a single-segment block (:268) names segment 1 whatever klow is; the caller passes the
 *             block it means.)
 *   passes    XXtemp = solve!(Euler(), ·, W, Q) -- the CURRENT noise re-run under the block's guide -- and
 *             XXᵒ = solve!(Euler(), ·, W°, Q) side by side (:324-325);
 *   diffll    start term + Σ_{s = s_hi-1 .. s_lo} (ll°[s] - ll_temp[s]) in that order (:331-333);
 *   accept    iff log(U) <= diffll (U as bb_pcn_step): W°, X° of the block's segments (and x0°) replace the current
 *             ones, acc += 1 (:340-353); otherwise nothing changes (XX keeps the path it had, as in the script).
 * Needs a double-buffered ensemble with X stored (BB_ENS_DOUBLE_BUFFER, no BB_ENS_NO_X); allocates a proposal copy of X
 * on first use.  Block updates keep no running ll (the script keeps none in this branch); the next whole-path step
 * (bb_theta_pcn_step / bb_theta_param_step) re-establishes it with one forward pass under the full backward chain, which
 * also re-simulates X from the chains' W under that chain's guide. */
int bb_theta_block_step(bb_ens* ens, int32_t s_lo, int32_t s_hi, double rho, double hzero, uint64_t seed,
                        uint32_t iter, int32_t skip);
/* numbers of the last block update: out [np][5 + 2 S] = logpdfnormal(x0 - ν(0)), logpdfnormal(x0° - ν(0)) (0, 0 when
 * s_lo > 0), Σ ll_temp, Σ ll°, diffll, then (ll_temp[s], ll°[s]) for s = 0 .. S-1 (entries outside the block are stale) */
int bb_theta_get_block(bb_ens* ens, int64_t p0, int64_t np, double* out);
/* accepted parameter proposals since attach, summed over this ensemble's chains */
int bb_theta_get_acc(bb_ens* ens, int64_t* acc);
void* bb_theta_acc_device_ptr(bb_ens* ens);

/* ------------------------------------------------------------------ multi-GPU
 * Chains are independent: rank r owns a contiguous block of GLOBAL chain ids (bb_ens_set_chain_offset; the random
 * streams are keyed by global chain id, so results do not depend on the number of GPUs) and all segments of a chain stay
 * on one GPU.  The only exchange is the acceptance statistic -- the reference's `acc += 1` (test/partialbridgenuH.jl:189)
 * -- one all-reduce(sum) of an int64.  It runs on the communicator's own stream behind an event: the compute stream gets
 * an 8-byte snapshot copy and nothing else, so the next iteration's kernel is never delayed.
 * One process (or host task) per GPU; NCCL is bound at run time (dlopen of libnccl.so.2, BB_NCCL_LIB overrides).
 *   rank 0: bb_comm_unique_id(id), ship the 128 bytes to the other ranks by any means (a file, MPI, torch.distributed);
 *   all ranks: bb_comm_create(ctx, nranks, rank, id, &comm)   [collective: ncclCommInitRank]
 *   or, for a host that already owns an ncclComm_t: bb_comm_adopt. */
#define BB_NCCL_ID_BYTES 128
typedef struct bb_comm bb_comm;
int bb_comm_unique_id(uint8_t* id /* [BB_NCCL_ID_BYTES] */);
int bb_comm_create(bb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id, bb_comm** out);
int bb_comm_adopt(bb_ctx* ctx, void* nccl_comm /* ncclComm_t */, int32_t nranks, int32_t rank, bb_comm** out);
int bb_comm_destroy(bb_comm* comm);
int bb_comm_rank(bb_comm* comm);
int bb_comm_size(bb_comm* comm);
const char* bb_comm_last_error(void);
/* start the all-reduce of the ensemble's acceptance counter as it stands after the calls issued so far (asynchronous) */
int bb_allreduce_acc(bb_ens* ens, bb_comm* comm);
/* the same for the counter of accepted parameter proposals (bb_theta_param_step) */
int bb_allreduce_theta_acc(bb_ens* ens, bb_comm* comm);
/* global sum delivered by the most recent bb_allreduce_*acc on this communicator (waits for that all-reduce only) */
int bb_comm_get_acc(bb_comm* comm, int64_t* acc_global);
int bb_comm_synchronize(bb_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* BRIDGE_B200_H */
