/*
 * p2de_b200.h — C ABI of libp2de_b200.so
 *
 * Drop-in boundary for ONE hot path of yiminllin/P2DE.jl: the per-SSP-RK-stage DG
 * right-hand side + limiter that `SSP33!` calls three times per time step.
 * Reference interfaces replaced (paths relative to the reference checkout):
 *
 *   rhs!(state, solver, state_param, time_param) -> dt   src/dg/rhs/rhs.jl:5-13
 *   rhs!(::LowOrderPositivity | ::FluxDiffRHS | ::LimitedDG, ...)  src/dg/rhs/rhs.jl:25-55
 *   apply_rhs_limiter!(::ZhangShuLimiter | ::SubcellLimiter, ...)  src/dg/limiter/limiter.jl:8-56
 *   SSP33!(state, solver, state_param)                  src/timestepping/SSPRK33.jl:1-62
 *   check_conservation(state, solver)                   src/dg/utils.jl:1-12
 *
 * The host (Julia via `ccall`, or the Python mirror in p2de_b200/) keeps building the
 * reference's own `Param`, operators and `BCData` (src/dg/init.jl:44-228) and hands the
 * arrays over VERBATIM: every matrix is column-major Float64 exactly as Julia stores
 * it, every index is Int64 and 1-based.  `Array{SVector{Nc,Float64},2}(n,K)` is a dense
 * `double[K][n][Nc]` (component fastest), see src/common/types/State.jl:1-26.
 *
 * Conventions
 *   - plain C, no C++/torch types; every entry point returns int32 (0 = ok, <0 = error)
 *     and never throws or aborts; `p2de_last_error` gives the message.
 *   - the library owns all device memory behind the opaque handle; host arrays passed
 *     in are copied during the call and may be freed afterwards.
 *   - one host thread drives one handle; calls are synchronous unless suffixed _async.
 *   - there is NO CPU fallback: without a CUDA device `p2de_create` fails with
 *     P2DE_ERR_CUDA.
 */
#ifndef P2DE_B200_H
#define P2DE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2DE_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------ */
enum {
  P2DE_OK = 0,
  P2DE_ERR_ARG = -1,         /* bad argument / inconsistent sizes                        */
  P2DE_ERR_UNSUPPORTED = -2, /* valid reference option that this build has no kernel for */
  P2DE_ERR_CUDA = -3,        /* CUDA runtime error (message has the cudaError string)    */
  P2DE_ERR_NCCL = -4,        /* NCCL error                                               */
  P2DE_ERR_STATE = -5        /* call order (e.g. rhs before set_state)                   */
};

/* ---- option enums: one per singleton type of src/common/types/Solver.jl ----------- */
enum { P2DE_BASIS_LOBATTO = 0, P2DE_BASIS_GAUSS = 1 };                 /* Solver.jl:89-91   */
enum { P2DE_RHS_LOW_ORDER_POSITIVITY = 0,                               /* Solver.jl:1-13    */
       P2DE_RHS_FLUX_DIFF = 1,
       P2DE_RHS_LIMITED_DG = 2 };
enum { P2DE_VOLFLUX_CHANDRASHEKAR = 0, P2DE_VOLFLUX_CENTRAL = 1 };       /* Solver.jl:15-17   */
enum { P2DE_SURFFLUX_CHANDRASHEKAR_PROJECTED = 0,                        /* Solver.jl:19-22   */
       P2DE_SURFFLUX_LF_NODAL = 1,
       P2DE_SURFFLUX_LF_PROJECTED = 2 };
enum { P2DE_PROJLIM_NONE = 0, P2DE_PROJLIM_NODEWISE = 1 };               /* Solver.jl:38-41   */
enum { P2DE_LIMITER_NONE = 0, P2DE_LIMITER_ZHANGSHU = 1, P2DE_LIMITER_SUBCELL = 2 }; /* :43-87 */
enum { P2DE_BOUND_POSITIVITY = 0,                                        /* Solver.jl:47-63   */
       P2DE_BOUND_POS_MIN_ENTROPY = 1,
       P2DE_BOUND_POS_RELAXED_MIN_ENTROPY = 2,
       P2DE_BOUND_POS_CELL_ENTROPY = 3,
       P2DE_BOUND_POS_RELAXED_CELL_ENTROPY = 4,
       P2DE_BOUND_TVD = 5,
       P2DE_BOUND_TVD_MIN_ENTROPY = 6,
       P2DE_BOUND_TVD_RELAXED_MIN_ENTROPY = 7,
       P2DE_BOUND_TVD_CELL_ENTROPY = 8,
       P2DE_BOUND_TVD_RELAXED_CELL_ENTROPY = 9 };
enum { P2DE_SHOCKCAPTURE_NONE = 0, P2DE_SHOCKCAPTURE_HENNEMANN = 1 };    /* Solver.jl:65-74   */

/* ---- fields readable with p2de_get_field (src/common/types/State.jl:1-26) --------- */
enum {
  P2DE_FIELD_UQ = 0,          /* [Nc,Nq,K]                                               */
  P2DE_FIELD_RHSU = 1,        /* [Nc,Nq,K]                                               */
  P2DE_FIELD_RHSH = 2,        /* [Nc,Nq,K]  (kept only when cfg.keep_diagnostics != 0)   */
  P2DE_FIELD_RHSL = 3,        /* [Nc,Nq,K]  (kept only when cfg.keep_diagnostics != 0)   */
  P2DE_FIELD_L = 4,           /* [K,Ns]     Zhang-Shu coefficient per element and stage  */
  P2DE_FIELD_L_LOCAL = 5,     /* [Nq+N1D,Nd,K,Ns] subcell coefficients (1D: first Nq+1 used) */
  P2DE_FIELD_THETA = 6,       /* [K,Ns]                                                  */
  P2DE_FIELD_THETA_LOCAL = 7, /* [Nfp,K,Ns]                                              */
  P2DE_FIELD_RESW = 8         /* [Nc,Nq,K]  the other state buffer (scratch of SSP33!; content after a step is unspecified) */
};

/* ---- reductions (p2de_reduce) ------------------------------------------------------ */
enum {
  P2DE_REDUCE_CONSERVATION = 0, /* sum_c sum_{i,k} J wq Uq[c]   src/dg/utils.jl:1-12     */
  P2DE_REDUCE_MIN_RHO = 1,      /* min_{i,k} rho                                           */
  P2DE_REDUCE_MIN_RHOE = 2      /* min_{i,k} rho*e  (rhoe_ufun, compressible_Navier_Stokes.jl:70-78) */
};

/* ---- Param (src/common/types/Solver.jl:151-170) flattened to a POD ----------------- */
typedef struct p2de_config {
  int32_t abi_version;      /* = P2DE_ABI_VERSION                                        */
  int32_t dim;              /* 1 | 2 (Dim1/Dim2)                                         */
  int32_t N;                /* polynomial degree, 1..4 on the GPU path                   */
  int32_t basis;            /* P2DE_BASIS_*                                              */
  int64_t K;                /* elements owned by this handle (= Kx*Ky_local)             */
  int32_t Kx, Ky;           /* structured extents of the local block; Ky = 1 in 1D       */
  int32_t Nq, Nfp, Nh, Np;  /* SizeData (Solver.jl:173-183); Np == Nq (collocation)      */
  int32_t rhs_type;         /* P2DE_RHS_*                                                */
  int32_t vol_flux;         /* P2DE_VOLFLUX_*                                            */
  int32_t surf_flux_low;    /* P2DE_SURFFLUX_LF_NODAL | _LF_PROJECTED                    */
  int32_t surf_flux_high;   /* P2DE_SURFFLUX_LF_PROJECTED | _CHANDRASHEKAR_PROJECTED     */
  int32_t proj_limiter;     /* P2DE_PROJLIM_*                                            */
  int32_t limiter;          /* P2DE_LIMITER_*                                            */
  int32_t bound;            /* P2DE_BOUND_*                                              */
  int32_t shockcapture;     /* P2DE_SHOCKCAPTURE_*                                       */
  int32_t keep_diagnostics; /* !=0: also keep rhsH/rhsL (costs 64 B per node per stage)  */
  int32_t device;           /* CUDA device ordinal; -1 = current device                  */
  int32_t lgl_projection_roundtrip; /* Lobatto face states u_tilde_f = u(v(U_node)) (rhs.jl:84-94).
                               0 (default): use U_node itself - Vf is a 0/1 gather so the entropy
                               projection is the identity up to rounding (~1e-16; the reference's
                               own TODO at rhs.jl:107).  1: evaluate the log/pow/exp round trip
                               like the reference does.                                     */
  int32_t _reserved;
  double hennemann_a, hennemann_c; /* HennemannShockCapture(a, c)  Solver.jl:68-74       */
  double bound_beta;        /* PositivityAndRelaxedCellEntropyBound(beta)                */
  double gamma;             /* CompressibleIdealGas.gamma                                */
  double POSTOL, ZEROTOL;   /* GlobalConstant                                            */
  double zeta, eta;         /* LimitingParameter                                         */
  double CFL, dt0, t0, T;   /* TimesteppingParameter                                     */
} p2de_config;

/* ---- Operators (src/common/types/Solver.jl:192-212), built in src/dg/init.jl:133-228 */
typedef struct p2de_operators {
  const double *Srsh_db[2]; /* [Nh,Nh] each, hybridized SBP 2*S  (init.jl:148-155)        */
  const double *Srs0[2];    /* [Nq,Nq] each, dense copy of the sparse low-order S (init.jl:276-322) */
  const double *Brs[2];     /* diag of Br, Bs: [Nfp] each (init.jl:151)                    */
  const double *Vf;         /* [Nfp,Nq]                                                    */
  const double *Vf_low;     /* [Nfp,Nq]  (init.jl:178-185)                                 */
  const double *MinvVhT;    /* [Np,Nh]   (init.jl:171)                                     */
  const double *MinvVfT;    /* [Np,Nfp]  (init.jl:172)                                     */
  const double *VDM_inv;    /* [Np,Nq]   inv(VDM), used by the Hennemann indicator         */
  const double *wq;         /* [Nq]                                                        */
  const int64_t *fq2q;      /* [Nfp] 1-based (init.jl:209-213)                             */
} p2de_operators;

/* ---- GeomData (Solver.jl:185-190).  The reference only builds uniform meshes
 *      (init.jl:137 "Assume uniform mesh"); arrays may be passed verbatim (they are
 *      checked to be constant) or omitted (NULL) with `uniform` set.                  */
typedef struct p2de_geometry {
  const double *J;      /* [Nq,K] md.J  or NULL                                            */
  const double *Jq;     /* [Nq,K]       or NULL                                            */
  const double *GJh[4]; /* rxJh,sxJh,ryJh,syJh: [Nh,K] each (1D: only [0]) or NULL         */
  int32_t uniform;      /* !=0: use the constants below instead of the arrays              */
  int32_t _pad;
  double J_const;       /* hx/2 (1D) | hx*hy/4 (2D)                                        */
  double GJ_const[4];   /* rxJ,sxJ,ryJ,syJ = hy/2,0,0,hx/2 (2D) | 1 (1D)                   */
} p2de_geometry;

/* ---- BCData (src/common/types/StateParam.jl:1-7) ----------------------------------- */
typedef struct p2de_bcdata {
  const int64_t *mapP; /* [Nfp,K] 1-based linear index into [Nfp,K]; NULL = structured     */
  int32_t periodic_x;  /* used when mapP == NULL: wrap in x / y, else self (make_periodic)  */
  int32_t periodic_y;
  int64_t nI;          /* inflow (Dirichlet) face nodes                                    */
  const int64_t *mapI; /* [nI] 1-based linear indices into [Nfp,K]                         */
  const double *Ival;  /* [Nc,nI]                                                          */
  int64_t nO;          /* outflow (copy-out) face nodes                                    */
  const int64_t *mapO; /* [nO]                                                             */
} p2de_bcdata;

typedef struct p2de_handle p2de_handle;

/* Build the device-side solver+state for one GPU.  Replaces the allocation half of
 * initialize_DG (src/dg/init.jl:44-60); the operators are the caller's.               */
int32_t p2de_create(const p2de_config *cfg, const p2de_operators *ops,
                    const p2de_geometry *geom, const p2de_bcdata *bc, p2de_handle **out);
int32_t p2de_destroy(p2de_handle *h);

/* Message of the last error on this handle (or of the last failed p2de_create when
 * h == NULL).  The pointer stays valid until the next call on the same handle.        */
const char *p2de_last_error(const p2de_handle *h);

/* Launch all work of this handle on `cuda_stream` (a cudaStream_t; NULL = default).   */
int32_t p2de_set_stream(p2de_handle *h, void *cuda_stream);

/* state.preallocation.Uq  <->  host `double[K][Nq][Nc]` (host->device / device->host) */
int32_t p2de_set_state(p2de_handle *h, const double *Uq_host);
int32_t p2de_get_state(p2de_handle *h, double *Uq_host);
/* Same, without the final stream synchronisation (pinned host memory expected).       */
int32_t p2de_set_state_async(p2de_handle *h, const double *Uq_host);
int32_t p2de_get_state_async(p2de_handle *h, double *Uq_host);
int32_t p2de_synchronize(p2de_handle *h);

/* rhs!(state, solver, state_param, TimeParam(t, dt, nstage)) -> dt  (rhs.jl:5-55).
 * Fills rhsU (and L / L_local[..., nstage]); nstage is 1-based like the reference.
 * At nstage == 1 `*dt_out` is the CFL-limited step (low_order_graph_viscosity.jl:222-243),
 * otherwise `dt` unchanged.                                                            */
int32_t p2de_rhs(p2de_handle *h, double t, double dt, int32_t nstage, double *dt_out);

/* Copy one observable field to the host; `n` = number of doubles `dst` can hold.      */
int32_t p2de_get_field(p2de_handle *h, int32_t field, double *dst_host, int64_t n);

/* One iteration of the `while t < T` loop of SSP33! (SSPRK33.jl:28-40): three fused
 * stages, state stays on the device.  `*dt_out` = the step actually taken.            */
int32_t p2de_ssp33_step(p2de_handle *h, double t, double *dt_out);
/* Same without host synchronisation: the step size stays on the device (read it later
 * with p2de_last_dt).  Lets a caller enqueue many steps back to back.                 */
int32_t p2de_ssp33_step_async(p2de_handle *h, double t);
int32_t p2de_last_dt(p2de_handle *h, double *dt_out);

/* The whole time loop of SSP33! from t0 = *t_inout up to cfg.T or max_steps steps.
 * dthist (may be NULL) receives up to max_steps step sizes.                            */
int32_t p2de_ssp33_run(p2de_handle *h, double *t_inout, int64_t max_steps,
                       int64_t *steps_out, double *dthist);

int32_t p2de_reduce(p2de_handle *h, int32_t what, double *out);

/* calculate_error (src/dg/postprocess.jl:1-28) on the device: `exact_host` = the exact solution's conserved
 * variables at the nodes, double[K][Nq][Nc] (evaluated by the caller's exact_sol callback); `out` receives
 * double[6][Nc] = sum wJ |ex-U|, sum wJ |ex-U|^2, max |ex-U|, sum wJ |ex|, sum wJ |ex|^2, max |ex| per component
 * (wJ = wq[i] * Jq).  The relative norms of :29-37 are three divisions away.                              */
int32_t p2de_calculate_error(p2de_handle *h, const double *exact_host, double *out);

/* DataHistory of SSP33! (src/timestepping/SSPRK33.jl:41-55, src/common/types/State.jl): p2de_ssp33_run keeps a
 * copy of Uq on the device whenever the reference would push one to Uhist (step counter i % output_interval == 0
 * or |t - T| < 1e-10), in a ring of `slots` states (the newest `slots` survive).  slots = 0 frees the ring.
 * p2de_snapshot_get copies snapshot `index` (0-based, in the order they were taken) and its time / step counter. */
int32_t p2de_snapshot_ring(p2de_handle *h, int32_t slots, int64_t output_interval);
int64_t p2de_snapshot_count(const p2de_handle *h);
int32_t p2de_snapshot_get(p2de_handle *h, int64_t index, double *Uq_host, double *t_out, int64_t *step_out);

/* ---- multi-GPU: element rows are partitioned in y, one handle (one process) per GPU.
 * `unique_id` is the 128-byte ncclUniqueId produced by p2de_comm_unique_id on rank 0
 * and broadcast by the host.  After this call p2de_rhs / p2de_ssp33_step exchange the
 * face halo with ranks +-1 and all-reduce(min) the CFL dt.                             */
int32_t p2de_comm_unique_id(uint8_t id_out[128]);
int32_t p2de_comm_init(p2de_handle *h, int32_t rank, int32_t nranks, const uint8_t unique_id[128]);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch of the two
 * hot kernels is bracketed by a CUDA event pair on the handle's stream.  kernel_id 0 =
 * stage_kernel, 1 = update_kernel.  p2de_profile(h, x) also clears the records.              */
int32_t p2de_profile(p2de_handle *h, int32_t enable);
int32_t p2de_profile_get(p2de_handle *h, int32_t kernel_id, double *total_ms, int64_t *launches);

/* Debug / evidence counters of the 2D FAST stage kernel (tests assert that the compile-time INTERIOR and deferred-
 * combine instantiations the benchmark times really ran; bench.py reports how much data-dependent work a workload
 * skips).  enable != 0 starts counting from zero, 0 stops; `out` (may be NULL) receives the current values:       */
enum {
  P2DE_DBG_CTA_GENERAL = 0,    /* CTAs that ran the general instantiation (boundary conditions, partial batches) */
  P2DE_DBG_CTA_INTERIOR = 1,   /* CTAs that ran the INTERIOR instantiation                                        */
  P2DE_DBG_CTA_DEFER = 2,      /* CTAs (of either kind) that formed the stage-1 SSP combine while loading         */
  P2DE_DBG_ELEM_LOGS = 3,      /* elements whose log(rho), log(beta) were evaluated (logmean off its series branch possible) */
  P2DE_DBG_ELEM = 4,           /* elements processed                                                              */
  P2DE_DBG_LINES = 5,          /* grid lines that ran the subcell limiter                                         */
  P2DE_DBG_LINES_NOT_EASY = 6, /* ... of which some coefficient needed the exact evaluation                       */
  P2DE_DBG_LIMITER_SLOW = 7,   /* evaluations that solved the quadratic (limiter_utils.jl:52-76)                  */
  P2DE_DBG_COUNT = 8
};
int32_t p2de_debug_counters(p2de_handle *h, int32_t enable, uint64_t out[P2DE_DBG_COUNT]);

/* Number of kernels this handle has launched so far (bench.py reports it).            */
int64_t p2de_kernel_launch_count(const p2de_handle *h);
/* Device pointer of Uq (for zero-copy wrapping by the host, e.g. torch.from_dlpack).  */
void *p2de_device_state_ptr(p2de_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* P2DE_B200_H */
