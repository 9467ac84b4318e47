// kernels2d.cuh — the per-stage hot path on 2D uniform quad meshes with LGL collocation.
//
// Two kernels per SSP-RK stage:
//
//   stage_kernel   = entropy projection + low-order graph-viscosity RHS + CFL dt +
//                    entropy-stable flux-differencing RHS + element-local part of the limiter
//                    (reference: rhs.jl:41-55, low_order_graph_viscosity.jl:4-243,
//                    flux_differencing.jl:4-361, limiter/subcell.jl:163-349, limiter/zhangshu.jl:4-45)
//   update_kernel  = interface symmetrisation of the subcell coefficients, limited reassembly
//                    and the SSP stage combine (subcell.jl:418-456,841-924, SSPRK33.jl:31-39)
//
// Thread mapping ("line threads"): an element with N1D x N1D nodes is handled by 2*N1D threads;
// thread (d, line) owns one grid line of nodes along axis d (d=0: x-line `line` = fixed j,
// d=1: y-line = fixed i).  On a Cartesian mesh every operator of the scheme couples nodes of one
// line only (tensor-product SBP; S_x acts along x-lines, S_y along y-lines), the two end nodes of
// a line are exactly its two face nodes, and the subcell limiter's running sums / coefficients
// are per line as well.  So a line thread keeps its N1D nodal states in registers and does the
// volume pairs, the two surface fluxes, the f_bar prefix sums and the limiting coefficients of
// its line without talking to anyone; x- and y-lines meet through shared memory only for
// rhsL = rhs_x + rhs_y (needed for u^L), the CFL sum and the final rhsU.
// Direction is data, not control flow: both kinds of lines run the same instruction stream.
//
// stage_kernel here is the GENERIC variant: every option of the reference (flux choices, Gauss collocation with the
// projected face states of gauss.cuh, Hennemann shock capturing, all ten subcell bounds) as run-time values of
// StageArgs, in the reference's operation order; CFG = 1 instantiates it with the options of the reference's shipped 2D
// examples fixed at compile time.  The default Lobatto configuration has its own kernel (stage_fast.cuh).
#pragma once
#include "physics.cuh"

namespace p2de {

#ifndef P2DE_STAGE_MIN_BLOCKS
#define P2DE_STAGE_MIN_BLOCKS 3   // generic kernel: 3 CTAs/SM at 168 registers beat 2 at 255 (S-KH-gauss 6.8 -> 6.1 ms); shared memory allows no fourth
#endif

enum { MODE_SUBCELL = 0, MODE_ZHANGSHU = 1, MODE_LOW = 2, MODE_HIGH = 3 };

template <int N1D>
struct Tables2D {
  // ---- the part stage_kernel_fast uses comes first: it copies only these FAST_BYTES to shared memory
  double SHt[2][N1D][N1D][N1D]; // physical hybridized S as [d][a][b][line] = GJ_dd * Srsh_db[d][node(a), node(b)]: bank-conflict free when the lanes of a warp differ in `line`
  double S0[2][N1D][N1D];       // physical low-order S0 of pair (a+1, a): [d][line][a]
  double Bf[2][N1D][2];         // physical signed boundary weight at the line ends [d][line][end]
  double wq[N1D * N1D];
  double rwJl[2][N1D][N1D];     // 1 / (Jq * wq) per grid line: [d][line][a]
  // ---- generic kernels only
  double SH[2][N1D][N1D][N1D];  // the same S as [d][line][a][b]
  double rwJ[N1D * N1D];        // 1 / (Jq * wq)
  double minv[N1D * N1D];       // MinvVhT[i, i]
  double minvf[4 * N1D];        // MinvVfT[fq2q[f], f]
  int fq2q[4 * N1D];            // 0-based
  // Gauss collocation (face nodes are not volume nodes): per line and line end e
  double VfL[2][N1D][2][N1D];   // Vf[f, node(a)]: extrapolation weights of the line's nodes to its end face node
  double SHf[2][N1D][2][N1D];   // physical hybridized S, face row x volume column: GJ_dd * Srsh_db[d][Nq + f, node(a)]
  static constexpr int FAST_BYTES = (int)(sizeof(double) * (2 * N1D * N1D * N1D + 2 * N1D * N1D + 2 * N1D * 2 + N1D * N1D + 2 * N1D * N1D));
};

struct MeshTopo {
  long long K;
  int Kx, Ky, periodic_x, periodic_y;
  int ghost_lo, ghost_hi;            // multi-GPU y-stripes: element row -1 / row Ky is a halo row owned by rank -/+ 1
  const int *mapP32;                 // generic mode: [K][Nfp] 0-based linear index; nullptr = structured
  const int *bcflag;                 // generic mode: 0 none, >0 inflow (index+1 into Ival), -1 outflow
  const double *Ival;                // generic mode: [nI][4]
  const unsigned char *bc_type[4];   // structured mode, per side L,R,B,T: [len][N1D] 0 none,1 inflow,2 outflow
  const double *bc_val[4];           // structured mode: [len][N1D][4]
};

struct StageArgs {
  const double *Uq;
  double *rhsL, *dF, *lpre;          // MODE_SUBCELL scratch (generic kernel: rhsL, dF, lpre; FAST: rpre, dFend, lpre)
  double *rpre, *dFend;
  double *rhsU;                      // other modes
  double *Lout;                      // MODE_ZHANGSHU: L[:, nstage]
  double *rhsH_diag, *rhsL_diag;     // optional full fields (keep_diagnostics)
  unsigned long long *dt_bits;       // stage 1: atomicMin target, pre-set to the cap
  const double *dt_dev;              // limiter dt read from the device when use_dt_dev
  double dt_host;
  int use_dt_dev, nstage;
  double gamma, ZEROTOL, POSTOL, zeta, CFL, Jq, blend;
  int vol_flux, surf_low, surf_high; // P2DE_VOLFLUX_*, P2DE_SURFFLUX_*
  int roundtrip;                     // evaluate u(v(U)) at LGL face nodes instead of using U
  double half_inv_gm1;               // 1 / (2 (gamma - 1))
  // generic kernel only: shock capturing and minimum-entropy bounds
  int hennemann;                     // HennemannShockCapture: blending factor from the modal indicator
  int entropy_bound;                 // 0 none, 1 PositivityAndMinEntropyBound, 2 PositivityAndRelaxedMinEntropyBound
  int N;                             // polynomial degree
  double hen_a, hen_c;
  const double *VDM_inv;             // [Np, Nq] column-major (device)
  const double *smin_dev;            // global minimum of s_modified at t0 (device scalar)
  // FAST subcell kernel, stages 2 and 3: write a*resW + b*(Uq + dt*rpre) instead of rpre
  int fuse;
  double fuse_a, fuse_b;
  const double *fuse_resW;
  const double *tab_dev;             // Tables2D<N1D> in global memory (same bytes as the kernel parameter)
  // generic kernel, Gauss collocation: entropy-projected face states u_tilde_f [K][Nfp][4] and the
  // projection-limiting parameters theta_local [K][Nfp] of this stage, both written by gauss_project_kernel
  int gauss;
  const double *utf, *theta_local;
  // generic kernel, remaining subcell bounds
  int tvd;                           // TVD*Bound: rho in [min, max] of the low-order update over the low-order stencil
  const double *rhsLpre;             // [K][Nq][4] low-order rhs of ALL elements (a MODE_LOW pre-pass): the stencil crosses faces
  int cell_entropy;                  // 0 none, 1 *CellEntropyBound, 2 *RelaxedCellEntropyBound(beta)
  double bound_beta;
  // subcell family, stage 2 of the direct schedule: the stage input is Uq + (dt / dt_host) (defer_add - Uq), defer_add = stage 1's W (the stage-1 SSP combine
  // U1 = U^n + dt rhsU, SSPRK33.jl:31-33, formed on the fly instead of by a separate pass over the mesh)
  const double *defer_add;
  double inv_cap;                    // 1 / dt_host
  int wform;                         // stage 1 of p2de_ssp33_step (subcell family): write W = U + dt_host rhsU instead of rhsU (stage_subcell.cuh: KIND_S1)
  int rowblocks;                     // FAST kernel: > 0 = 2D grid (rowblocks x rows), Kx = rowblocks * EPB; 0 = 1D grid over batches
  int row0, row_stride, nrows;       // 2D grid: the launch covers element rows row0 + i * row_stride, i < nrows (nrows = 0: all Ky rows)
  double *fstar;                     // Gauss + cell entropy: [K][Nfp][2][4] normal components of fstar_H, fstar_L (State.jl:11-12)
  unsigned long long *dbg;           // nullptr, or the counters of p2de_debug_counters (P2DE_DBG_*): which instantiations ran, data-dependent shortcuts taken
};

// counters behind p2de_debug_counters (include/p2de_b200.h)
enum { DBG_CTA_GENERAL = 0, DBG_CTA_INTERIOR = 1, DBG_CTA_DEFER = 2, DBG_ELEM_LOGS = 3, DBG_ELEM = 4, DBG_LINES = 5,
       DBG_LINES_NOT_EASY = 6, DBG_LIMITER_SLOW = 7, DBG_COUNT = 8 };

struct UpdateArgs {
  const double *rhsL, *dF, *lpre;    // MODE_SUBCELL inputs
  const double *rpre, *dFend;        // FAST stage kernel's scratch (update_kernel_fast)
  const double *rhsU_in;             // other modes
  double *Llocal_out;                // symmetrised L_local[:, :, :, nstage] or nullptr
  double *rhsU_out;                  // or nullptr
  const double *Uq_in, *resW;        // stage combine: Uq_out = a*resW + b*(Uq_in + dt*rhsU); Uq_out nullptr = skip
  double *Uq_out;
  double a, b;
  const double *dt_dev;
  double dt_host;
  int use_dt_dev;
  int rotated;                       // dF of y-lines is stored in the rotated frame (FAST stage kernel)
  int pre_updated;                   // rpre already holds the SSP combine of the un-corrected rhs (StageArgs.fuse)
  double Jq;
  // Gauss + cell-entropy bounds: enforce_ES_subcell_interface! (subcell.jl:718-805) needs the numerical fluxes of both
  // sides of a face and the entropy variables / potentials of the face nodes' volume nodes (from Uq_in)
  const double *fstar;               // [K][Nfp][2][4] or nullptr
  double gamma;
};

struct Nbr { long long kP; int fP; int bc; const double *ival; };

template <int N1D>
P2DE_DEV Nbr neighbor(const MeshTopo &M, long long k, int ix, int iy, int f) {
  constexpr int Nfp = 4 * N1D;
  Nbr nb;
  nb.bc = 0; nb.ival = nullptr;
  if (M.mapP32) {
    int m = M.mapP32[k * Nfp + f];
    nb.kP = m / Nfp; nb.fP = m % Nfp;
    int fl = M.bcflag ? M.bcflag[k * Nfp + f] : 0;
    if (fl > 0) { nb.bc = 1; nb.ival = M.Ival + 4ll * (fl - 1); }
    else if (fl < 0) nb.bc = 2;
    return nb;
  }
  int F = f / N1D, a = f % N1D;
  int jx = ix + (F == 0 ? -1 : F == 1 ? 1 : 0), jy = iy + (F == 2 ? -1 : F == 3 ? 1 : 0);
  bool out = jx < 0 || jx >= M.Kx || jy < 0 || jy >= M.Ky;
  if (out && ((F == 2 && M.ghost_lo) || (F == 3 && M.ghost_hi))) {
    nb.kP = jx + (long long)jy * M.Kx; nb.fP = (F ^ 1) * N1D + a;   // jy = -1 or Ky: halo row
    return nb;
  }
  if (out) {
    bool wrap = F < 2 ? M.periodic_x : M.periodic_y;
    if (wrap) { jx = (jx + M.Kx) % M.Kx; jy = (jy + M.Ky) % M.Ky; nb.kP = jx + (long long)jy * M.Kx; nb.fP = (F ^ 1) * N1D + a; }
    else { nb.kP = k; nb.fP = f; }
    if (M.bc_type[F]) {
      int pos = F < 2 ? iy : ix;
      int t = M.bc_type[F][pos * N1D + a];
      if (t) { nb.bc = t; nb.ival = M.bc_val[F] + 4ll * (pos * N1D + a); }
    }
  } else { nb.kP = jx + (long long)jy * M.Kx; nb.fP = (F ^ 1) * N1D + a; }
  return nb;
}

P2DE_DEV Cons2 load_cons(const double *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = q[0], b = q[1];
  Cons2 U; U.rho = a.x; U.m1 = a.y; U.m2 = b.x; U.E = b.y;
  return U;
}
P2DE_DEV void store4(double *p, const double v[4]) {
  double2 *q = reinterpret_cast<double2 *>(p);
  q[0] = make_double2(v[0], v[1]); q[1] = make_double2(v[2], v[3]);
}
P2DE_DEV void cons_arr(const Cons2 &U, double v[4]) { v[0] = U.rho; v[1] = U.m1; v[2] = U.m2; v[3] = U.E; }

#ifndef P2DE_FIND_ALPHA_BISECT
#define P2DE_FIND_ALPHA_BISECT 0   // 1: the reference's 50-step bisection (validation aid)
#endif
// find_alpha, low_order_graph_viscosity.jl:299-327: the smallest alpha with rho(alpha u - u~) > eps and
// rho e(alpha u - u~) > eps.  The reference brackets it by doubling and bisects 50 times (an interval of 2^-50 alpha_0);
// what the bisection converges to is known in closed form.  With beta = alpha - 1 and delta = u~ - u (formed first, so
// that nothing cancels when u~ is close to u), alpha u - u~ = beta u - delta and
//   rho > eps                   <=>  beta > (delta_rho + eps) / rho
//   E rho - |m|^2/2 > eps rho   <=>  A beta^2 + B beta + C > 0,   A = E rho - |m|^2/2  (> 0),
//                                    B = -(E delta_rho + delta_E rho) + m . delta_m - eps rho,
//                                    C = delta_E delta_rho - |delta_m|^2/2 + eps delta_rho,
// so alpha = 1 + max(density bound, larger root), the root taken in its cancellation-free form.  Agrees with the
// bisection to its own evaluation noise (1e-14 typical, the predicate is evaluated on differences of size eps); alpha
// only enters the CFL dt (lambda_B_CFL :287-293), tested to 1e-12.
template <class C>
P2DE_DEV double find_alpha_closed(double eps, double rho, double E, double msq, double drho, double dE, double mdm, double dmsq) {
  double beta = (drho + eps) / rho;
  const double Aq = E * rho - 0.5 * msq;
  const double Bq = -(E * drho + dE * rho) + mdm - eps * rho;
  const double Cq = dE * drho - 0.5 * dmsq + eps * drho;
  const double disc = Bq * Bq - 4.0 * Aq * Cq;
  if (disc >= 0.0) {
    const double sq = sqrt(disc);
    const double root = Bq <= 0.0 ? (sq - Bq) / (2.0 * Aq) : -2.0 * Cq / (Bq + sq);
    beta = fmax(beta, root);
  }
  return fmax(1.0 + beta, 8.881784197001252e-16);   // (the bisection never returns less than 2^-50)
}
P2DE_DEV double find_alpha(double POSTOL, const Cons2 &ui, const Cons2 &ut) {
  if (!P2DE_FIND_ALPHA_BISECT) {
    const double dr = ut.rho - ui.rho, d1 = ut.m1 - ui.m1, d2 = ut.m2 - ui.m2, dE = ut.E - ui.E;
    return find_alpha_closed<void>(POSTOL, ui.rho, ui.E, ui.m1 * ui.m1 + ui.m2 * ui.m2, dr, dE, ui.m1 * d1 + ui.m2 * d2, d1 * d1 + d2 * d2);
  }
  double alphaL = 0.0, alphaR = 1.0;
  Cons2 s;
  auto sub = [&](double al) { s.rho = al * ui.rho - ut.rho; s.m1 = al * ui.m1 - ut.m1; s.m2 = al * ui.m2 - ut.m2; s.E = al * ui.E - ut.E; };
  // (rhoe_ufun with a Newton reciprocal: 51 IEEE divisions per face node otherwise; guarded by rho > POSTOL > 0)
  auto ok = [&]() { return s.rho > POSTOL && s.E - 0.5 * (s.m1 * s.m1 + s.m2 * s.m2) * rcp_fast(s.rho) > POSTOL; };
  sub(alphaR);
  while (!ok() && alphaR < 1e300) { alphaR = 2 * alphaR; sub(alphaR); }
  for (int it = 0; it < 50; ++it) {
    double alphaM = (alphaL + alphaR) / 2;
    sub(alphaM);
    if (ok()) alphaR = alphaM; else alphaL = alphaM;
  }
  return alphaR;
}

template <int N1D, int MODE>
constexpr int stage_smem_doubles_per_elem() {
  constexpr int Nq = N1D * N1D;
  return 12 * Nq + 8 * Nq + ((MODE == MODE_SUBCELL) ? 0 : 8 * Nq) + 6 * Nq + N1D + 4 * Nq + 3 * 2 * N1D;
}
// extra shared memory of the TVD / cell-entropy bounds (allocated only when the bound is on): rhoL [S], its ghosts
// [2][S], and per direction dvdf, dv.f_bar_L and the interior coefficients [2][Nq - N1D] each
template <int N1D>
constexpr int stage_smem_extra_doubles_per_elem(bool tvd, bool cell, bool slim = false) {
  return (tvd ? 3 * N1D * N1D : 0) + (cell ? 6 * N1D * (N1D - 1) : 0) + (slim ? 8 * N1D * N1D : 0);
}

// s_modified_ufun, compressible_Navier_Stokes.jl:80-85
P2DE_DEV double s_modified(double gamma, const Cons2 &U) { return rhoe2(U) * pow(U.rho, -gamma); }

// bisection, src/math/nonlinear_solvers.jl:3-20, on f(l) = s_modified(U + l P) >= Lphi - POSTOL
// (limiting_param_bound_phi, limiter_utils.jl:42-50)
__device__ __noinline__ double limiting_param_phi(double gamma, double POSTOL, const Cons2 &U, const double Pv[4], double Lphi, double lpos) {
  auto f = [&](double l) {
    Cons2 w; w.rho = U.rho + l * Pv[0]; w.m1 = U.m1 + l * Pv[1]; w.m2 = U.m2 + l * Pv[2]; w.E = U.E + l * Pv[3];
    return s_modified(gamma, w) >= Lphi - POSTOL;
  };
  if (f(lpos)) return lpos;
  double x_valid = 0.0, x_invalid = lpos;
  for (int iter = 0; iter <= 20; ++iter) {
    double x_new = 0.5 * (x_valid + x_invalid);
    if (f(x_new)) x_valid = x_new; else x_invalid = x_new;
  }
  return x_valid;
}

// limiting_param_bound_rho_rhoe, limiter_utils.jl:26-40, with a finite upper density bound (TVD bounds) and
// Urhoe = Inf; the caller's min(L_local, .) with L_local = 1 is folded in
__device__ __noinline__ double limiting_param_rho_bounds(double ZEROTOL, const Cons2 &U, double c, const double Pv[4], double Lrho,
                                                         double Urho, double Lrhoe) {
  double a, b;
  quad_coeff_ab(U, Pv, Lrhoe, a, b);
  double l = 1.0;
  if (U.rho + Pv[0] < Lrho) l = jl_max((Lrho - U.rho) / Pv[0], 0.0);
  if (U.rho + Pv[0] > Urho) l = jl_min(l, jl_max((Urho - U.rho) / Pv[0], 0.0));
  l = jl_min(l, rhoe_quadratic_roots(ZEROTOL, a, b, c));
  return jl_min(l, 1.0);
}

// v_ufun(::Dim2), compressible_Navier_Stokes.jl:134-144
P2DE_DEV void v_ufun2(double gamma, double gm1, const Cons2 &U, double V[4]) {
  double p = pfun2(gm1, U);
  double s = log(p / pow(U.rho, gamma));                 // sfun :64-68
  V[0] = (gamma + 1 - s) - gm1 * U.E / p;
  V[1] = U.m1 * gm1 / p; V[2] = U.m2 * gm1 / p; V[3] = -U.rho * gm1 / p;
}

// Base.isless on Float64 (NaN largest, -0.0 < 0.0), for the descending (value, index) order of
// sort!(dvdf_order_k, rev=true), subcell.jl:607,673,703
P2DE_DEV bool jl_isless(double a, double b) {
  if (a != a) return false;
  if (b != b) return true;
  if (a == b) return signbit(a) && !signbit(b);
  return a < b;
}

// enforce_ES_subcell_volume!, subcell.jl:612-707, one direction of one element, serial.  dv / dvfL / Lc hold the
// NE = N1D (N1D - 1) interior subcell faces of the direction in the reference's dvdf index order
// (x: (si - 1) + sj (N1D - 1); y: si + (sj - 1) N1D); `ysum` = the y estimate is accumulated si outer, sj inner (:641-646).
template <int N1D>
__device__ __noinline__ void es_volume_greedy(const double *dv, const double *dvfL, double *Lc, bool ysum, double sBpsi,
                                              int relaxed, double beta, double epsk, double ZEROTOL) {
  constexpr int NE = N1D * (N1D - 1);
  double sdvfL = 0.0, sum_poslim = 0.0;
  for (int i = 0; i < NE; ++i) {
    const int idx = ysum ? (i / (N1D - 1)) + (i % (N1D - 1)) * N1D : i;
    sdvfL += dvfL[idx];
    sum_poslim += Lc[idx] * dv[idx];
  }
  const double rhs = relaxed ? (1 - beta * epsk) * (sBpsi - sdvfL) : sBpsi - sdvfL;   // rhs_es :709-716
  const double tol = jl_max(0.0, sdvfL - sBpsi);
  if (!(sum_poslim - rhs > tol)) return;
  unsigned used = 0;          // NE <= 20 entries: selection in descending (value, index) order instead of a sort
  int taken[NE];
  int curr = 0;
  double lhs = sum_poslim;
  while (lhs > rhs + tol && curr < NE) {
    int best = -1;
    for (int i = 0; i < NE; ++i) {
      if ((used >> i) & 1u) continue;
      if (best < 0 || jl_isless(dv[best], dv[i]) || (!jl_isless(dv[i], dv[best]) && i > best)) best = i;
    }
    if (dv[best] < ZEROTOL) break;
    used |= 1u << best;
    lhs = lhs - Lc[best] * dv[best];
    taken[curr++] = best;
  }
  for (int i = 0; i < curr; ++i) {
    const int idx = taken[i];
    const double l_new = (i == curr - 1) ? jl_max((rhs + tol - lhs) / dv[idx], 0.0) : 0.0;
    Lc[idx] = jl_min(Lc[idx], l_new);
  }
}

// FAST = default flux configuration (Chandrashekar volume flux, Lax-Friedrichs surface fluxes on
// nodal/projected values, identity LGL projection): hoisted reciprocals, merged low/high surface
// flux, three-division two-point flux.  !FAST = every other option, reference operation order.
// CFG = 0: every option is a run-time value of StageArgs.  CFG = 1: the configuration of the reference's shipped 2D
// examples (examples/2D/kelvin-helmholtz.jl:44-55: Gauss collocation, Chandrashekar volume flux, Lax-Friedrichs on projected
// values for both surface fluxes, PositivityBound, no shock capturing) with the options fixed at compile time: the same
// arithmetic, but the branches and the code of the other options do not exist (the run-time version of that
// configuration no longer fits the instruction cache: 18 % instruction-fetch stalls).
template <int N1D, int MODE, int EPB, int CFG>
__global__ void __launch_bounds__(EPB * 2 * N1D, P2DE_STAGE_MIN_BLOCKS)
stage_kernel(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M,
             const __grid_constant__ Tables2D<N1D> Tc) {
  constexpr int Nq = N1D * N1D, TPE = 2 * N1D, NF = N1D + 1, NFLD = 12;
  constexpr bool FAST = false;   // (the default Lobatto configuration has its own kernel, stage_fast.cuh)
  constexpr bool DG = CFG == 1;
  const int o_gauss = DG ? 1 : A.gauss, o_roundtrip = DG ? 0 : A.roundtrip, o_vol_flux = DG ? P2DE_VOLFLUX_CHANDRASHEKAR : A.vol_flux;
  const int o_surf_low = DG ? P2DE_SURFFLUX_LF_PROJECTED : A.surf_low, o_surf_high = DG ? P2DE_SURFFLUX_LF_PROJECTED : A.surf_high;
  const int o_hennemann = DG ? 0 : A.hennemann, o_entropy_bound = DG ? 0 : A.entropy_bound, o_tvd = DG ? 0 : A.tvd;
  const int o_cell_entropy = DG ? 0 : A.cell_entropy;
  double *const o_fstar = DG ? nullptr : A.fstar;
  constexpr bool DO_LOW = MODE != MODE_HIGH, DO_HIGH = MODE != MODE_LOW;
  constexpr int TBLC = (sizeof(Tables2D<N1D>) + 7) / 8;          // doubles to copy
  constexpr int TBL = ((sizeof(Tables2D<N1D>) + 15) / 16) * 2;   // keeps what follows 16-byte aligned
  constexpr int S = EPB * Nq;
  extern __shared__ double sm[];
  Tables2D<N1D> &T = *reinterpret_cast<Tables2D<N1D> *>(sm);
  double *nodes = sm + TBL;                       // [NFLD][EPB*Nq]
  double *partsL = nodes + NFLD * EPB * Nq;       // [EPB*Nq][2][4]
  double *partsH = partsL + 8 * EPB * Nq;         // [EPB*Nq][2][4]   (not MODE_SUBCELL)
  double *lamp = partsH + ((MODE == MODE_SUBCELL) ? 0 : 8 * EPB * Nq);  // [EPB*Nq][6]
  double *lmin = lamp + 6 * EPB * Nq;             // [EPB][N1D]
  double *smod = lmin + EPB * N1D;                // [S] s_modified at the nodes
  double *lbnd = smod + S;                        // [S] indicator rho*p, later the lower bound on s_modified
  double *ghst = lbnd + S;                        // [2][S] s_modified across the x / y face of boundary nodes
  double *ered = ghst + 2 * S;                    // [EPB][TPE][3] modal energy partial sums
  // optional arrays (stage_smem_extra_doubles_per_elem)
  constexpr int NE = N1D * (N1D - 1);
  double *rhoLs = ered + EPB * TPE * 3;           // [S] density of the low-order update (TVD bounds)
  double *ghR = rhoLs + S;                        // [2][S] the same across the x / y face of boundary nodes
  double *esD = ghR + 2 * S - (o_tvd ? 0 : 3 * S);   // [EPB][2][NE] dvdf (cell-entropy bounds)
  double *esF = esD + EPB * 2 * NE;               // [EPB][2][NE] dv . f_bar_L
  double *esL = esF + EPB * 2 * NE;               // [EPB][2][NE] interior coefficients
  const bool need_ind = o_hennemann || o_entropy_bound == 2 || o_cell_entropy == 2;
  // CFG = 1 writes the slim scratch of the FAST family (un-symmetrised limited rhs + end-face dF, consumed by
  // update_kernel_fast) instead of rhsL + every subcell face's dF: 84 instead of 132 B/node out, 84 instead of 196 B/node
  // into the update kernel.  (No TVD / cell-entropy arrays in that configuration: the shares take their place.)
  constexpr bool SLIM = DG && MODE == MODE_SUBCELL;
  double *shr = nodes + 4 * S;                    // [S][2][4] the two directions' shares of the limited rhs: node fields 4..11 are
                                                  // only read in the line phase, two barriers earlier (no extra shared memory)

  const int tid = threadIdx.x, el = tid / TPE, ln = tid % TPE, d = ln / N1D, line = ln % N1D;
  const long long k = (long long)blockIdx.x * EPB + el;
  const bool active = k < M.K;
  const int nbase = el * Nq;
  const double gamma = A.gamma, gm1 = A.gamma - 1.0;

  {
    // from global memory (same bytes as the kernel parameter): an indexed read of the parameter is a lane-serialised
    // constant-bank access (6 % of this kernel's stall samples before)
    const double *src = A.tab_dev;
    for (int i = tid; i < TBLC; i += EPB * TPE) sm[i] = src[i];
  }
  // ---- node phase: primitives, logs, axis wavespeeds of every volume node (once per node)
  //      calculate_primitive_variables! flux_differencing.jl:39-52; wavespeed_estimate :48-62
  if (active) {
    for (int node = ln; node < Nq; node += TPE) {
      Cons2 U = load_cons(A.Uq + (k * Nq + node) * 4);
      double *o = nodes + nbase + node;
      o[0 * S] = U.rho; o[1 * S] = U.m1; o[2 * S] = U.m2; o[3 * S] = U.E;
      double p, beta;
      if (FAST) {
        double rinv = 1.0 / U.rho;
        p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
        o[4 * S] = U.m1 * rinv; o[5 * S] = U.m2 * rinv;
        if (DO_LOW) { o[10 * S] = wavespeed_fast(gamma, gm1, rinv, U.m1, U.E); o[11 * S] = wavespeed_fast(gamma, gm1, rinv, U.m2, U.E); }
      } else if (o_gauss) {   // fast reciprocal / square root (physics.cuh: "_fd"), see fS_dir_fd
        double rinv = rcp_fast(U.rho);
        p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
        o[4 * S] = U.m1 * rinv; o[5 * S] = U.m2 * rinv;
        if (DO_LOW) { o[10 * S] = wavespeed_dir_fd(gamma, gm1, U, 0); o[11 * S] = wavespeed_dir_fd(gamma, gm1, U, 1); }
      } else {
        p = pfun2(gm1, U);
        o[4 * S] = U.m1 / U.rho; o[5 * S] = U.m2 / U.rho;
        if (DO_LOW) { o[10 * S] = wavespeed_dir(gamma, gm1, U, 0); o[11 * S] = wavespeed_dir(gamma, gm1, U, 1); }
      }
      beta = (!FAST && o_gauss) ? div_fast(U.rho, 2 * p) : U.rho / (2 * p);
      o[6 * S] = p; o[7 * S] = beta;
      if (DO_HIGH) { o[8 * S] = log(U.rho); o[9 * S] = log(beta); }
      if (o_entropy_bound) smod[nbase + node] = s_modified(gamma, U);
      if (need_ind) lbnd[nbase + node] = U.rho * p;   // indicator, shock_capture.jl:82-94
    }
  }
  __syncthreads();

  // ---- modal smoothness indicator (shock_capture.jl:47-80), blending factor (:111-132) and
  //      smoothness factor of the relaxed bound (subcell.jl:932-956); per element, every thread
  double blend = A.blend, epsk = o_entropy_bound == 1 ? 1.0 : 0.0;
  if (need_ind) {
    double eN = 0.0, eNm1 = 0.0, etot = 0.0;
    if (active)
      for (int m = ln; m < Nq; m += TPE) {
        double coef = 0.0;
        for (int j = 0; j < Nq; ++j) coef += A.VDM_inv[m + j * Nq] * lbnd[nbase + j];
        double e = coef * coef;
        int mi = m % N1D, mj = m / N1D;
        if (mi == N1D - 1 || mj == N1D - 1) eN += e;
        if (mi == N1D - 2 || mj == N1D - 2) eNm1 += e;
        etot += e;
      }
    ered[(el * TPE + ln) * 3 + 0] = eN; ered[(el * TPE + ln) * 3 + 1] = eNm1; ered[(el * TPE + ln) * 3 + 2] = etot;
    __syncthreads();
    eN = eNm1 = etot = 0.0;
    for (int t2 = 0; t2 < TPE; ++t2) { eN += ered[(el * TPE + t2) * 3]; eNm1 += ered[(el * TPE + t2) * 3 + 1]; etot += ered[(el * TPE + t2) * 3 + 2]; }
    const double sigma = jl_max(eN / etot, eNm1 / etot);
    if (o_hennemann) {
      const double TN = A.hen_a * pow(10.0, -A.hen_c * pow((double)(A.N + 1), 0.25));
      const double s_factor = log((1 - 0.0001) / 0.0001);
      const double al = 1.0 / (1.0 + exp(-s_factor / TN * (sigma - TN)));
      blend = jl_max(jl_min(1.0 - al, 1.0), 0.5);
    }
    if (o_entropy_bound == 2 || o_cell_entropy == 2) {
      const double kappa = 1.0, s0 = log10(pow((double)A.N, -4.0)), sk = log10(sigma);
      epsk = sk < s0 - kappa ? 0.0 : (sk > s0 + kappa ? 1.0 : 0.5 - 0.5 * sin(3.141592653589793 * (sk - s0) / (2 * kappa)));
    }
  }

  // ---- line phase
  Cons2 U[N1D];
  double GL[N1D][4], GH[N1D][4];          // wJ * rhsxy_d of this line's nodes (low / high order)
  double BFL[2][4], BFH[2][4];
  double lamPair[N1D], lamFace[2];
  double wJ[N1D], rwJ[N1D];
  const double dtl = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;   // dt the limiter sees (rhs.jl:46,52)
  if (active) {
    double uu[N1D], vv[N1D], pp[N1D];
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      int node = d == 0 ? a + line * N1D : line + a * N1D;
      const double *o = nodes + nbase + node;
      U[a].rho = o[0 * S]; U[a].m1 = o[1 * S]; U[a].m2 = o[2 * S]; U[a].E = o[3 * S];
      uu[a] = o[4 * S]; vv[a] = o[5 * S]; pp[a] = o[6 * S];
      wJ[a] = A.Jq * T.wq[node]; rwJ[a] = T.rwJ[node];
#pragma unroll
      for (int c = 0; c < 4; ++c) { GL[a][c] = 0.0; GH[a][c] = 0.0; }
    }
    const int ix = (int)(k % M.Kx), iy = (int)(k / M.Kx);
    Nbr nb[2];
    Cons2 Unb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      nb[e] = neighbor<N1D>(M, k, ix, iy, (2 * d + e) * N1D + line);
      Unb[e] = load_cons(A.Uq + (nb[e].kP * Nq + T.fq2q[nb[e].fP]) * 4);
    }
    if (o_entropy_bound) {   // low_order_stencil across the element boundary (limiter_utils.jl:184-231)
      ghst[d * S + nbase + (d == 0 ? 0 + line * N1D : line)] = s_modified(gamma, Unb[0]);
      ghst[d * S + nbase + (d == 0 ? (N1D - 1) + line * N1D : line + (N1D - 1) * N1D)] = s_modified(gamma, Unb[1]);
    }
    if (MODE == MODE_SUBCELL && o_tvd) {   // rhoL = rho + dt rhsL[1] of the stencil node across the face (subcell.jl:119-141)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ae = e ? N1D - 1 : 0;
        ghR[d * S + nbase + (d == 0 ? ae + line * N1D : line + ae * N1D)] =
            Unb[e].rho + dtl * A.rhsLpre[(nb[e].kP * Nq + T.fq2q[nb[e].fP]) * 4];
      }
    }
    double fl[N1D][4];                      // nodal flux along d
#pragma unroll
    for (int a = 0; a < N1D; ++a) flux_dir(U[a], uu[a], vv[a], pp[a], d, fl[a]);

    if (FAST) {
      // ---- both surface fluxes at the two line ends.  With the identity LGL projection the
      //      low-order LF flux on nodal values (low_order_graph_viscosity.jl:175-204) and the
      //      high-order LF flux on projected values (flux_differencing.jl:90-151,223-272) are
      //      the same numbers; on inflow/outflow faces the high-order one drops the dissipation.
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ae = e ? N1D - 1 : 0;
        const int node = d == 0 ? ae + line * N1D : line + ae * N1D;
        double B = T.Bf[d][line][e], nn = fabs(B);
        double rinvP = 1.0 / Unb[e].rho;
        double wsP = wavespeed_fast(gamma, gm1, rinvP, d == 0 ? Unb[e].m1 : Unb[e].m2, Unb[e].E);
        double wsM = DO_LOW ? nodes[(10 + d) * S + nbase + node] : wavespeed_fast(gamma, gm1, 1.0 / U[ae].rho, d == 0 ? U[ae].m1 : U[ae].m2, U[ae].E);
        double lamB = 0.5 * nn * jl_max(wsM, wsP);
        Cons2 uP = Unb[e];
        if (nb[e].bc) {
          uP = nb[e].bc == 1 ? load_cons(nb[e].ival) : U[ae];
          rinvP = 1.0 / uP.rho;
        }
        double fP[4], up[4], uf[4];
        flux_dir(uP, uP.m1 * rinvP, uP.m2 * rinvP, gm1 * (uP.E - 0.5 * (uP.m1 * uP.m1 + uP.m2 * uP.m2) * rinvP), d, fP);
        cons_arr(uP, up); cons_arr(U[ae], uf);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double bfs = B * (0.5 * (fl[ae][c] + fP[c]));
          double lf = lamB * (up[c] - uf[c]);
          BFL[e][c] = bfs - lf;
          BFH[e][c] = nb[e].bc ? bfs : bfs - lf;
        }
        lamFace[e] = lamB;
      }
    }

    Cons2 Ut[2], Utnb[2];   // entropy-projected face states (rhs.jl:84-94), mine and the neighbour's
    if (!FAST) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        Ut[e] = U[e ? N1D - 1 : 0];
        Utnb[e] = Unb[e];
        if (o_gauss) {   // rhs.jl:113-133: u(v_tilde) at the face nodes, mine and my neighbour's
          Ut[e] = load_cons(A.utf + (k * (4 * N1D) + (2 * d + e) * N1D + line) * 4);
          Utnb[e] = load_cons(A.utf + (nb[e].kP * (4 * N1D) + nb[e].fP) * 4);
        } else if (o_roundtrip) {
          Ut[e] = entropy_roundtrip(gamma, gm1, Ut[e]);
          Utnb[e] = entropy_roundtrip(gamma, gm1, Utnb[e]);
        }
      }
    }

    if (DO_LOW) {
      // ---- low-order graph-viscosity volume terms along this line, low_order_graph_viscosity.jl:139-173
      double ws[N1D];
#pragma unroll
      for (int a = 0; a < N1D; ++a) ws[a] = nodes[(10 + d) * S + nbase + (d == 0 ? a + line * N1D : line + a * N1D)];
#pragma unroll
      for (int a = 0; a < N1D - 1; ++a) {
        const int i = a + 1, j = a;
        double Sv = T.S0[d][line][a], nn = fabs(Sv);
        double lam = nn * jl_max(ws[i], ws[j]);
        lamPair[a] = lam;
        double ui[4], uj[4];
        cons_arr(U[i], ui); cons_arr(U[j], uj);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double F = 0.5 * (fl[i][c] + fl[j][c]);
          double SF = 2.0 * Sv * F - lam * (uj[c] - ui[c]);
          GL[i][c] -= SF; GL[j][c] += SF;          // GL = -Q0F1
        }
      }
      lamPair[N1D - 1] = 0.0;
      if (!FAST) {
        // surface, :175-204 (reference operation order; LaxFriedrichsOnProjectedVal supported)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ae = e ? N1D - 1 : 0;
          double B = T.Bf[d][line][e], nn = fabs(B);
          bool proj = o_surf_low == P2DE_SURFFLUX_LF_PROJECTED;
          Cons2 Uf = proj ? Ut[e] : U[ae];
          Cons2 UfP = proj ? Utnb[e] : Unb[e];
          double wsM = proj ? (o_gauss ? wavespeed_dir_fd(gamma, gm1, Uf, d) : wavespeed_dir(gamma, gm1, Uf, d)) : ws[ae];
          double wsP = o_gauss ? wavespeed_dir_fd(gamma, gm1, UfP, d) : wavespeed_dir(gamma, gm1, UfP, d);
          double lamB = 0.5 * nn * jl_max(wsM, wsP);
          Cons2 uP = UfP;
          if (nb[e].bc == 1) uP = load_cons(nb[e].ival);
          else if (nb[e].bc == 2) uP = U[ae];
          double fM[4], fP[4], uf[4], up[4];
          if (proj) { if (o_gauss) flux_dir_fd(gm1, Uf, d, fM); else flux_dir(gm1, Uf, d, fM); }
          else {
#pragma unroll
            for (int c = 0; c < 4; ++c) fM[c] = fl[ae][c];
          }
          if (o_gauss) flux_dir_fd(gm1, uP, d, fP); else flux_dir(gm1, uP, d, fP);
          cons_arr(Uf, uf); cons_arr(uP, up);
#pragma unroll
          for (int c = 0; c < 4; ++c) BFL[e][c] = B * (0.5 * (fM[c] + fP[c])) - lamB * (up[c] - uf[c]);
          if (DG) {
            // both surface fluxes are Lax-Friedrichs on the projected values: BF_H (flux_differencing.jl:223-272) is the
            // same expression on the same data, except that LFc = 0 on inflow/outflow faces (:116-151)
#pragma unroll
            for (int c = 0; c < 4; ++c) BFH[e][c] = nb[e].bc ? B * (0.5 * (fM[c] + fP[c])) : BFL[e][c];
          }
          if (MODE == MODE_SUBCELL && o_fstar) {   // fstar_L = f* - lf / B (apply_LF_dissipation_to_fstar, rhs_utils.jl:93-102)
            double *fo = o_fstar + ((k * (4 * N1D) + (2 * d + e) * N1D + line) * 2 + 1) * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) fo[c] = 0.5 * (fM[c] + fP[c]) - (lamB * (up[c] - uf[c])) / B;
          }
          lamFace[e] = lamB;
          if (proj && A.nstage == 1) {   // lambda_B_CFL(::LaxFriedrichsOnProjectedVal), :287-291
            double alpha = find_alpha(A.POSTOL, U[ae], Uf);
            lamFace[e] = alpha * lamB + 0.5 * nn * wsM;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { GL[0][c] -= BFL[0][c]; GL[N1D - 1][c] -= BFL[1][c]; }
    }

    if (DO_HIGH) {
      // ---- flux differencing along this line, flux_differencing.jl:164-211 (pairs j<i, j outer).
      //      GH = -(QF1 + B F*) (LGL: M^-1 Vh^T is a scaled 0/1 gather; the hybridized face-volume
      //      pairs cancel identically and are not formed)
      Prim2 q[N1D];
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double *o = nodes + nbase + (d == 0 ? a + line * N1D : line + a * N1D);
        q[a].rho = U[a].rho; q[a].u = uu[a]; q[a].v = vv[a];
        q[a].beta = o[7 * S]; q[a].rholog = o[8 * S]; q[a].betalog = o[9 * S];
      }
      if (FAST || o_vol_flux == P2DE_VOLFLUX_CHANDRASHEKAR) {
#pragma unroll
        for (int j = 0; j < N1D; ++j)
#pragma unroll
          for (int i = j + 1; i < N1D; ++i) {
            double F[4];
            if (FAST) fS_fast(A.half_inv_gm1, q[i], q[j], d, F);
            else if (o_gauss) fS_dir_fd(gm1, q[i], q[j], d, F);
            else fS_dir(gm1, q[i], q[j], d, F);
            double Sv = T.SH[d][line][i][j];
#pragma unroll
            for (int c = 0; c < 4; ++c) { double Sf = Sv * F[c]; GH[i][c] -= Sf; GH[j][c] += Sf; }
          }
      } else {   // CentralFlux, flux_differencing.jl:217-221
#pragma unroll
        for (int j = 0; j < N1D; ++j)
#pragma unroll
          for (int i = j + 1; i < N1D; ++i) {
            double Sv = T.SH[d][line][i][j];
#pragma unroll
            for (int c = 0; c < 4; ++c) { double Sf = Sv * (0.5 * (fl[i][c] + fl[j][c])); GH[i][c] -= Sf; GH[j][c] += Sf; }
          }
      }
      double GHf[2][4];   // -QF1 at the two hybridized face rows of this line (Gauss)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int c = 0; c < 4; ++c) GHf[e][c] = 0.0;
      if (!FAST && o_gauss) {
        // hybridized face-volume pairs of Srsh_db = [Q - Q^T, E^T B; -B E, 0] (init.jl:148-155): face row i > volume
        // column j, QF1[i] += S_ij fS(u_i, u_j), QF1[j] -= the same (flux_differencing.jl:164-211)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          Prim2 pf = prim_of_fd(gm1, Ut[e]);
          double ff[4];
          flux_dir(gm1, Ut[e], d, ff);
#pragma unroll
          for (int a = 0; a < N1D; ++a) {
            double F[4];
            if (o_vol_flux == P2DE_VOLFLUX_CHANDRASHEKAR) fS_dir_fd(gm1, pf, q[a], d, F);
            else {
#pragma unroll
              for (int c = 0; c < 4; ++c) F[c] = 0.5 * (ff[c] + fl[a][c]);
            }
            double Sv = T.SHf[d][line][e][a];
#pragma unroll
            for (int c = 0; c < 4; ++c) { double Sf = Sv * F[c]; GHf[e][c] -= Sf; GH[a][c] += Sf; }
          }
        }
      }
      if (!FAST && !(DG && DO_LOW)) {
        // surface, flux_differencing.jl:90-151,223-272
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ae = e ? N1D - 1 : 0;
          double B = T.Bf[d][line][e], nn = fabs(B);
          double LFc = o_gauss ? 0.5 * nn * jl_max(wavespeed_dir_fd(gamma, gm1, Ut[e], d), wavespeed_dir_fd(gamma, gm1, Utnb[e], d))
                               : 0.5 * nn * jl_max(wavespeed_dir(gamma, gm1, Ut[e], d), wavespeed_dir(gamma, gm1, Utnb[e], d));
          Cons2 uP = Utnb[e];
          if (nb[e].bc == 1) { uP = load_cons(nb[e].ival); LFc = 0.0; }
          else if (nb[e].bc == 2) { uP = U[ae]; LFc = 0.0; }
          double fs[4];
          if (o_surf_high == P2DE_SURFFLUX_CHANDRASHEKAR_PROJECTED) {
            fS_dir(gm1, prim_of(gm1, Ut[e]), prim_of(gm1, uP), d, fs);
          } else {
            double fM[4], fP[4];
            flux_dir(gm1, Ut[e], d, fM); flux_dir(gm1, uP, d, fP);
#pragma unroll
            for (int c = 0; c < 4; ++c) fs[c] = 0.5 * (fM[c] + fP[c]);
          }
          double uf[4], up[4];
          cons_arr(Ut[e], uf); cons_arr(uP, up);
#pragma unroll
          for (int c = 0; c < 4; ++c) BFH[e][c] = B * fs[c] - LFc * (up[c] - uf[c]);
          if (MODE == MODE_SUBCELL && o_fstar) {   // fstar_H, flux_differencing.jl:263-270
            double *fo = o_fstar + ((k * (4 * N1D) + (2 * d + e) * N1D + line) * 2 + 0) * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) fo[c] = fs[c] - (LFc * (up[c] - uf[c])) / B;
          }
        }
      }
      if (!FAST && o_gauss) {
        // assemble_rhs! (flux_differencing.jl:274-361): M^-1 Vh^T = (1/wq) [I Vf_new^T], Vf_new = theta Vf + (1 - theta) Vf_low
        // per face node (:288-319); the 1/(wq J) factor is applied below with rwJ
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ae = e ? N1D - 1 : 0;
          const double th = A.theta_local ? A.theta_local[k * (4 * N1D) + (2 * d + e) * N1D + line] : 1.0;
#pragma unroll
          for (int a = 0; a < N1D; ++a) {
            const double w = th * T.VfL[d][line][e][a] + (1 - th) * (a == ae ? 1.0 : 0.0);
#pragma unroll
            for (int c = 0; c < 4; ++c) GH[a][c] += w * (GHf[e][c] - BFH[e][c]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { GH[0][c] -= BFH[0][c]; GH[N1D - 1][c] -= BFH[1][c]; }
      }
    }

    // ---- publish this line's share of rhsL / rhsH / lambda for the node-wise combination
    //      (scale_low_order_rhs_by_mass! :206-220, assemble_rhs! flux_differencing.jl:331-361)
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      int node = d == 0 ? a + line * N1D : line + a * N1D;
      if (DO_LOW) {
        double r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = GL[a][c] * rwJ[a];
        store4(partsL + (nbase + node) * 8 + d * 4, r);
      }
      if (MODE != MODE_SUBCELL && DO_HIGH) {
        double r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = GH[a][c] * rwJ[a];
        store4(partsH + (nbase + node) * 8 + d * 4, r);
      }
      if (DO_LOW && A.nstage == 1) {
        double *lp = lamp + (nbase + node) * 6 + d * 3;
        lp[0] = a > 0 ? lamPair[a - 1] : 0.0;
        lp[1] = lamPair[a];
        lp[2] = a == 0 ? lamFace[0] : (a == N1D - 1 ? lamFace[1] : 0.0);
      }
    }
  }
  __syncthreads();

  // ---- lower bound on s_modified: stencil minimum relaxed towards the global minimum
  //      (initialize_lower_bound!, subcell.jl:55-75)
  if (o_entropy_bound) {
    if (active)
      for (int node = ln; node < Nq; node += TPE) {
        const int i = node % N1D, j = node / N1D;
        double lb = smod[nbase + node];
        lb = jl_min(lb, i > 0 ? smod[nbase + node - 1] : ghst[0 * S + nbase + node]);
        lb = jl_min(lb, i < N1D - 1 ? smod[nbase + node + 1] : ghst[0 * S + nbase + node]);
        lb = jl_min(lb, j > 0 ? smod[nbase + node - N1D] : ghst[1 * S + nbase + node]);
        lb = jl_min(lb, j < N1D - 1 ? smod[nbase + node + N1D] : ghst[1 * S + nbase + node]);
        lbnd[nbase + node] = epsk * lb + (1 - epsk) * (*A.smin_dev);
      }
    __syncthreads();
  }

  // ---- density of the low-order update at the element's own nodes (initialize_TVD_bounds!, subcell.jl:119-127)
  if (MODE == MODE_SUBCELL && o_tvd) {
    if (active)
      for (int node = ln; node < Nq; node += TPE) {
        const double *pl = partsL + (nbase + node) * 8;
        rhoLs[nbase + node] = nodes[0 * S + nbase + node] + dtl * (pl[0] + pl[4]);
      }
    __syncthreads();
  }

  // ---- CFL: dt = min_i CFL * 0.5 * wJ_i / lambda_i, low_order_graph_viscosity.jl:222-281
  if (DO_LOW && A.nstage == 1) {
    double dtloc = INFINITY;
    if (active && d == 0) {
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        int node = a + line * N1D;
        const double *lp = lamp + (nbase + node) * 6;
        double li = 0.0;
        li += lp[3]; li += lp[0]; li += lp[1]; li += lp[4];   // partners ascending: y-, x-, x+, y+
        li += lp[2]; li += lp[5];                             // q2fq: vertical face first, then horizontal
        dtloc = jl_min(dtloc, A.CFL * 0.5 * wJ[a] / li);
      }
    }
#pragma unroll
    {
      const unsigned wmask = __activemask();   // the CTA's last warp may be partial
      for (int off = 16; off > 0; off >>= 1) {
        double other = __shfl_xor_sync(wmask, dtloc, off);
        if ((wmask >> ((tid & 31) ^ off)) & 1u) dtloc = jl_min(dtloc, other);
      }
    }
    if ((tid & 31) == 0) dt_publish(A.dt_bits, dtloc);
  }

  if (!active && MODE != MODE_ZHANGSHU && !(MODE == MODE_SUBCELL && (o_cell_entropy || SLIM))) return;

  if (MODE == MODE_SUBCELL) {
    // ---- subcell limiter, element-local part: f_bar prefix sums (subcell.jl:163-206) and the
    //      limiting coefficients of this line's N1D+1 subcell faces (subcell.jl:248-349)
    // (inactive threads of a partial batch only get here with the cell-entropy bounds, whose barriers every thread of
    //  the CTA has to reach at the same place)
    double dFv[NF][4];
    double lv[NF];
    if (active) {
#pragma unroll
    for (int c = 0; c < 4; ++c) dFv[0][c] = BFH[0][c] - BFL[0][c];
#pragma unroll
    for (int s = 1; s < NF; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) dFv[s][c] = dFv[s - 1][c] + (GH[s - 1][c] - GL[s - 1][c]);
    Cons2 uL[N1D];
    double Lrho[N1D], Lrhoe[N1D], c0[N1D], Urho[N1D];
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      int node = d == 0 ? a + line * N1D : line + a * N1D;
      const double *pl = partsL + (nbase + node) * 8;
      double r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) r[c] = pl[c] + pl[4 + c];
      uL[a].rho = U[a].rho + dtl * r[0]; uL[a].m1 = U[a].m1 + dtl * r[1];
      uL[a].m2 = U[a].m2 + dtl * r[2]; uL[a].E = U[a].E + dtl * r[3];
      Lrho[a] = A.zeta * uL[a].rho; Lrhoe[a] = A.zeta * rhoe2(uL[a]);
      Urho[a] = INFINITY;
      if (o_tvd) {   // stencil min / max of rhoL (subcell.jl:129-141, low_order_stencil limiter_utils.jl:222-231; rho_bound :352-361)
        const int i = node % N1D, j = node / N1D;
        double lb = rhoLs[nbase + node], ub = lb, v;
        v = i > 0 ? rhoLs[nbase + node - 1] : ghR[0 * S + nbase + node]; lb = jl_min(lb, v); ub = jl_max(ub, v);
        v = i < N1D - 1 ? rhoLs[nbase + node + 1] : ghR[0 * S + nbase + node]; lb = jl_min(lb, v); ub = jl_max(ub, v);
        v = j > 0 ? rhoLs[nbase + node - N1D] : ghR[1 * S + nbase + node]; lb = jl_min(lb, v); ub = jl_max(ub, v);
        v = j < N1D - 1 ? rhoLs[nbase + node + N1D] : ghR[1 * S + nbase + node]; lb = jl_min(lb, v); ub = jl_max(ub, v);
        Lrho[a] = lb; Urho[a] = ub;
      }
      c0[a] = quad_coeff_c(uL[a], Lrhoe[a]);
      if (d == 0) {
        if (!SLIM) store4(A.rhsL + (k * Nq + node) * 4, r);
        if (A.rhsL_diag) store4(A.rhsL_diag + (k * Nq + node) * 4, r);
      }
    }
#pragma unroll
    for (int s = 0; s < NF; ++s) {
      double l = 1.0;
      if (s < N1D) {   // node to the right/top of the face: P = -4 dt (fH - fL) / wJ (subcell.jl:300,328)
        double Pv[4], kk = -4 * dtl * rwJ[s];
#pragma unroll
        for (int c = 0; c < 4; ++c) Pv[c] = kk * dFv[s][c];
        double lp = o_tvd ? limiting_param_rho_bounds(A.ZEROTOL, uL[s], c0[s], Pv, Lrho[s], Urho[s], Lrhoe[s])
                          : limiting_param_pos(A.ZEROTOL, uL[s], c0[s], Pv, Lrho[s], Lrhoe[s]);
        if (o_entropy_bound) lp = limiting_param_phi(gamma, A.POSTOL, uL[s], Pv, lbnd[nbase + (d == 0 ? s + line * N1D : line + s * N1D)], lp);
        l = jl_min(l, lp);
      }
      if (s >= 1) {    // node to the left/bottom: P = +4 dt (fH - fL) / wJ (subcell.jl:312,340)
        double Pv[4], kk = 4 * dtl * rwJ[s - 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) Pv[c] = kk * dFv[s][c];
        double lp = o_tvd ? limiting_param_rho_bounds(A.ZEROTOL, uL[s - 1], c0[s - 1], Pv, Lrho[s - 1], Urho[s - 1], Lrhoe[s - 1])
                          : limiting_param_pos(A.ZEROTOL, uL[s - 1], c0[s - 1], Pv, Lrho[s - 1], Lrhoe[s - 1]);
        if (o_entropy_bound) lp = limiting_param_phi(gamma, A.POSTOL, uL[s - 1], Pv, lbnd[nbase + (d == 0 ? (s - 1) + line * N1D : line + (s - 1) * N1D)], lp);
        l = jl_min(l, lp);
      }
      lv[s] = jl_min(l, blend);
    }
    }   // active
    if (o_cell_entropy) {
      // enforce_ES_subcell! on the element's interior subcell faces (subcell.jl:462-707; on Lobatto nodes the interface
      // part is a no-op, :714-716, on Gauss nodes it is done by update_kernel): dvdf = (v_{s-1} - v_s) . (f_bar_H - f_bar_L),
      // dv . f_bar_L per face by the line threads, the greedy update by one thread per element and direction
      if (active) {
      double vprev[4], vcur[4], fL[4];
      v_ufun2(gamma, gm1, U[0], vprev);
#pragma unroll
      for (int c = 0; c < 4; ++c) fL[c] = BFL[0][c];
#pragma unroll
      for (int s = 1; s < N1D; ++s) {
        v_ufun2(gamma, gm1, U[s], vcur);
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          fL[c] += GL[s - 1][c];
          const double dv = vprev[c] - vcur[c];
          a1 += dv * dFv[s][c]; a2 += dv * fL[c];
          vprev[c] = vcur[c];
        }
        const int i = (el * 2 + d) * NE + (d == 0 ? (s - 1) + line * (N1D - 1) : line + (s - 1) * N1D);
        esD[i] = a1; esF[i] = a2; esL[i] = lv[s];
      }
      }   // active
      __syncthreads();
      if (active && line == 0) {
        double sB = 0.0;   // sum_Bpsi[k][d], subcell.jl:519-528: psi = (gamma - 1) (rho u, rho v) at the face nodes' volume nodes
#pragma unroll
        for (int e = 0; e < 2; ++e)
          for (int l2 = 0; l2 < N1D; ++l2) {
            const int ae = e ? N1D - 1 : 0, node = d == 0 ? ae + l2 * N1D : l2 + ae * N1D;
            sB += T.Bf[d][l2][e] * (gm1 * nodes[(1 + d) * S + nbase + node]);
          }
        es_volume_greedy<N1D>(esD + (el * 2 + d) * NE, esF + (el * 2 + d) * NE, esL + (el * 2 + d) * NE, d == 1, sB,
                              o_cell_entropy == 2, A.bound_beta, epsk, A.ZEROTOL);
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int s = 1; s < N1D; ++s) lv[s] = esL[(el * 2 + d) * NE + (d == 0 ? (s - 1) + line * (N1D - 1) : line + (s - 1) * N1D)];
      }
    }
    if (SLIM) {
      if (active) {
        // this line's share of the un-symmetrised limited rhs: rhsxyL_d + (l_{a+1} dF_{a+1} - l_a dF_a) / wJ (subcell.jl:841-924)
#pragma unroll
        for (int a = 0; a < N1D; ++a) {
          const int node = d == 0 ? a + line * N1D : line + a * N1D;
          double t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) t[c] = GL[a][c] * rwJ[a] + (lv[a + 1] * dFv[a + 1][c] - lv[a] * dFv[a][c]) * rwJ[a];
          store4(shr + ((nbase + node) * 2 + d) * 4, t);
        }
        store4(A.dFend + (k * (4 * N1D) + (2 * d + 0) * N1D + line) * 4, dFv[0]);
        store4(A.dFend + (k * (4 * N1D) + (2 * d + 1) * N1D + line) * 4, dFv[N1D]);
        double *ldst = A.lpre + (k * 2 + d) * (N1D * NF);
#pragma unroll
        for (int s = 0; s < NF; ++s) ldst[d == 0 ? s + line * NF : line + s * N1D] = lv[s];
        if (A.rhsH_diag) {
#pragma unroll
          for (int a = 0; a < N1D; ++a) {
            int node = d == 0 ? a + line * N1D : line + a * N1D;
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(A.rhsH_diag + (k * Nq + node) * 4 + c, GH[a][c] * rwJ[a]);
          }
        }
      }
      __syncthreads();
      if (active && d == 0) {
#pragma unroll
        for (int a = 0; a < N1D; ++a) {
          const int node = a + line * N1D;
          const double *x = shr + (nbase + node) * 8;
          double r[4] = {x[0] + x[4], x[1] + x[5], x[2] + x[6], x[3] + x[7]};
          store4(A.rpre + (k * Nq + node) * 4, r);
        }
      }
      return;
    }
    if (!active) return;
    double *dst = A.dF + ((k * 2 + d) * N1D + line) * (NF * 4);
#pragma unroll
    for (int s = 0; s < NF; ++s) store4(dst + s * 4, dFv[s]);
    double *ldst = A.lpre + (k * 2 + d) * (N1D * NF);
#pragma unroll
    for (int s = 0; s < NF; ++s) ldst[d == 0 ? s + line * NF : line + s * N1D] = lv[s];
    if (A.rhsH_diag) {   // diagnostics: each line adds its share of rhsH (rhsH_diag is zeroed by the host)
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        int node = d == 0 ? a + line * N1D : line + a * N1D;
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(A.rhsH_diag + (k * Nq + node) * 4 + c, GH[a][c] * rwJ[a]);
      }
    }
    return;
  }

  // ---- element-local limiters / no limiter: produce rhsU here
  double rL[N1D][4], rH[N1D][4];
  double lline = 1.0;
  if (active && d == 0) {
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      int node = a + line * N1D;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        rL[a][c] = DO_LOW ? partsL[(nbase + node) * 8 + c] + partsL[(nbase + node) * 8 + 4 + c] : 0.0;
        rH[a][c] = DO_HIGH ? partsH[(nbase + node) * 8 + c] + partsH[(nbase + node) * 8 + 4 + c] : 0.0;
      }
      if (MODE == MODE_ZHANGSHU) {   // zhangshu.jl:4-45
        Cons2 uL;
        uL.rho = U[a].rho + dtl * rL[a][0]; uL.m1 = U[a].m1 + dtl * rL[a][1];
        uL.m2 = U[a].m2 + dtl * rL[a][2]; uL.E = U[a].E + dtl * rL[a][3];
        double Pv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) Pv[c] = dtl * (rH[a][c] - rL[a][c]);
        double Lrhoe = A.zeta * rhoe2(uL);
        lline = jl_min(lline, limiting_param_pos(A.ZEROTOL, uL, quad_coeff_c(uL, Lrhoe), Pv, A.zeta * uL.rho, Lrhoe));
      }
    }
    if (MODE == MODE_ZHANGSHU) lmin[el * N1D + line] = lline;
  }
  double l = 1.0;
  if (MODE == MODE_ZHANGSHU) {
    __syncthreads();
    if (!active) return;
#pragma unroll
    for (int j = 0; j < N1D; ++j) l = jl_min(l, lmin[el * N1D + j]);
    if (ln == 0) A.Lout[k] = l;
    l = jl_min(l, blend);
  }
  if (d == 0) {
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      int node = a + line * N1D;
      double r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        r[c] = MODE == MODE_ZHANGSHU ? (1 - l) * rL[a][c] + l * rH[a][c] : (MODE == MODE_LOW ? rL[a][c] : rH[a][c]);
      store4(A.rhsU + (k * Nq + node) * 4, r);
      if (A.rhsL_diag && DO_LOW) store4(A.rhsL_diag + (k * Nq + node) * 4, rL[a]);
      if (A.rhsH_diag && DO_HIGH) store4(A.rhsH_diag + (k * Nq + node) * 4, rH[a]);
    }
  }
}

// subcell index of the neighbour's interface coefficient (limiter_utils.jl:123-181)
template <int N1D>
P2DE_DEV int lidx_of_face(int fP) {
  constexpr int NF = N1D + 1;
  int F = fP / N1D, a = fP % N1D;
  if (F < 2) return (F == 0 ? 0 : N1D) + a * NF;   // x block: si + sj*(N1D+1)
  return a + (F == 2 ? 0 : N1D) * N1D;             // y block: si + sj*N1D
}

// solve_l_es_interface!, subcell.jl:797-805: bisection(l -> l dvfH + (1 - l) dvfL <= dpsi, 0, l0) (nonlinear_solvers.jl:3-20)
__device__ __noinline__ double es_interface_bisection(double l0, double dvfH, double dvfL, double dpsi) {
  auto f = [&](double l) { return l * dvfH + (1 - l) * dvfL <= dpsi; };
  if (f(l0)) return l0;
  double x_valid = 0.0, x_invalid = l0;
  for (int iter = 0; iter <= 20; ++iter) {
    const double x_new = 0.5 * (x_valid + x_invalid);
    if (f(x_new)) x_valid = x_new; else x_invalid = x_new;
  }
  return x_valid;
}

template <int N1D, int MODE, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D)
update_kernel(const __grid_constant__ UpdateArgs A, const __grid_constant__ MeshTopo M,
              const __grid_constant__ Tables2D<N1D> Tc) {
  constexpr int Nq = N1D * N1D, TPE = 2 * N1D, NF = N1D + 1;
  __shared__ double cy[EPB * Nq * 4];
  __shared__ double s_rwJ[Nq];
  const int tid = threadIdx.x, el = tid / TPE, ln = tid % TPE, d = ln / N1D, line = ln % N1D;
  const long long k = (long long)blockIdx.x * EPB + el;
  const bool active = k < M.K;
  const int nbase = el * Nq;
  double cx[N1D][4];
  if (MODE == MODE_SUBCELL) {
    if (tid < Nq) s_rwJ[tid] = Tc.rwJ[tid];
    __syncthreads();
    if (active) {
      // symmetrize_limiting_parameters!, subcell.jl:418-456, as a pure gather (min is idempotent)
      double lv[NF];
      const double *lsrc = A.lpre + (k * 2 + d) * (N1D * NF);
#pragma unroll
      for (int s = 0; s < NF; ++s) lv[s] = lsrc[d == 0 ? s + line * NF : line + s * N1D];
      const int ix = (int)(k % M.Kx), iy = (int)(k / M.Kx);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int f = (2 * d + e) * N1D + line;
        Nbr nb = neighbor<N1D>(M, k, ix, iy, f);
        double lP = A.lpre[(nb.kP * 2 + d) * (N1D * NF) + lidx_of_face<N1D>(nb.fP)];
        double lsym = jl_min(lv[e ? N1D : 0], lP);
        if (A.fstar && nb.kP != k) {
          // enforce_ES_subcell_interface!(::Dim2, ::GaussCollocation), subcell.jl:718-805: each side replaces its
          // coefficient by bisection(l -> l dv.f*_H + (1 - l) dv.f*_L <= dpsi, 0, min(l, l_P)) with ITS OWN fluxes and
          // dv = v_f - v_fP, dpsi = psi_f - psi_fP, then symmetrize takes the minimum.  The reference does this in a
          // loop over k that reads the partner's coefficient as it is at that moment; in element order the element with
          // the lower index goes first and the other one starts from its result (B(l) <= l), which is what both
          // sides evaluate here.
          const int ae = e ? N1D - 1 : 0, node = d == 0 ? ae + line * N1D : line + ae * N1D;
          const double gm1 = A.gamma - 1.0;
          const Cons2 uM = load_cons(A.Uq_in + (k * Nq + node) * 4), uPn = load_cons(A.Uq_in + (nb.kP * Nq + Tc.fq2q[nb.fP]) * 4);
          double vM[4], vP[4], dv[4];
          v_ufun2(A.gamma, gm1, uM, vM); v_ufun2(A.gamma, gm1, uPn, vP);
#pragma unroll
          for (int c = 0; c < 4; ++c) dv[c] = vM[c] - vP[c];
          const double dpsi = gm1 * (d == 0 ? uM.m1 : uM.m2) - gm1 * (d == 0 ? uPn.m1 : uPn.m2);
          const double *fm = A.fstar + (k * (4 * N1D) + f) * 8, *fp = A.fstar + (nb.kP * (4 * N1D) + nb.fP) * 8;
          double hM = 0.0, lM = 0.0, hP = 0.0, lPd = 0.0;
#pragma unroll
          for (int c = 0; c < 4; ++c) { hM += dv[c] * fm[c]; lM += dv[c] * fm[4 + c]; hP += (-dv[c]) * fp[c]; lPd += (-dv[c]) * fp[4 + c]; }
          const bool mine_first = k < nb.kP;
          double l = es_interface_bisection(lsym, mine_first ? hM : hP, mine_first ? lM : lPd, mine_first ? dpsi : -dpsi);
          lsym = es_interface_bisection(l, mine_first ? hP : hM, mine_first ? lPd : lM, mine_first ? -dpsi : dpsi);
        }
        lv[e ? N1D : 0] = lsym;
      }
      if (A.Llocal_out) {
        double *ldst = A.Llocal_out + (k * 2 + d) * (N1D * NF);
#pragma unroll
        for (int s = 0; s < NF; ++s) ldst[d == 0 ? s + line * NF : line + s * N1D] = lv[s];
      }
      // accumulate_f_bar_limited! + apply_subcell_limiter!, subcell.jl:841-924, in the form
      // rhsU = rhsL + sum_d (l_{s+1} dF_{s+1} - l_s dF_s) / wJ   (f_lim = f_L + l (f_H - f_L))
      const double *dsrc = A.dF + ((k * 2 + d) * N1D + line) * (NF * 4);
      double g[NF][4];
#pragma unroll
      for (int s = 0; s < NF; ++s) {
        Cons2 t = load_cons(dsrc + s * 4);
        g[s][0] = lv[s] * t.rho; g[s][1] = lv[s] * t.m1; g[s][2] = lv[s] * t.m2; g[s][3] = lv[s] * t.E;
      }
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        int node = d == 0 ? a + line * N1D : line + a * N1D;
        double rw = s_rwJ[node];
#pragma unroll
        for (int c = 0; c < 4; ++c) cx[a][c] = (g[a + 1][c] - g[a][c]) * rw;
        if (d == 1) store4(cy + (nbase + node) * 4, cx[a]);
      }
    }
    __syncthreads();
  }
  if (!active || d != 0) return;
  const double dt = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;
#pragma unroll
  for (int a = 0; a < N1D; ++a) {
    int node = a + line * N1D;
    long long off = (k * Nq + node) * 4;
    double r[4];
    if (MODE == MODE_SUBCELL) {
      Cons2 rl = load_cons(A.rhsL + off);
      const double *y = cy + (nbase + node) * 4;
      // FAST stage kernel: the y-lines' dF (hence cy) is in their rotated frame (momenta swapped)
      const int s1 = A.rotated ? 2 : 1, s2 = A.rotated ? 1 : 2;
      r[0] = rl.rho + (cx[a][0] + y[0]); r[1] = rl.m1 + (cx[a][1] + y[s1]);
      r[2] = rl.m2 + (cx[a][2] + y[s2]); r[3] = rl.E + (cx[a][3] + y[3]);
    } else {
      Cons2 ru = load_cons(A.rhsU_in + off);
      cons_arr(ru, r);
    }
    if (A.rhsU_out) store4(A.rhsU_out + off, r);
    if (A.Uq_out) {   // SSPRK33.jl:31-39
      Cons2 u = load_cons(A.Uq_in + off);
      double un[4], uo[4];
      cons_arr(u, uo);
      if (A.b == 1.0 && A.a == 0.0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) un[c] = uo[c] + dt * r[c];
      } else {
        Cons2 w = load_cons(A.resW + off);
        double wv[4];
        cons_arr(w, wv);
#pragma unroll
        for (int c = 0; c < 4; ++c) un[c] = A.a * wv[c] + A.b * (uo[c] + dt * r[c]);
      }
      store4(A.Uq_out + off, un);
    }
  }
}

}  // namespace p2de
