// stage_subcell.cuh — the stage kernel of the default configuration with the subcell positivity limiter (the path
// p2de_ssp33_step runs three times per step): LGL nodes, Chandrashekar volume flux, Lax-Friedrichs surface fluxes,
// PositivityBound.  Same scheme, thread mapping ("line threads", direction-homogeneous warps, rotated frame, swizzled
// shared memory) and results as stage_fast.cuh's MODE_SUBCELL body, which this file replaces; reorganised around what the
// round-2 profile showed to bind that body: not the FP64 pipe (47 %) but the L1 / shared-memory data pipe (83 % of its
// wavefront rate; every 8-byte access of a warp costs two 128-byte wavefronts).
//
//   * Only the conserved state goes through shared memory (4 fields per node instead of 12).  Velocities, pressure,
//     beta, the axis wavespeed (and, where needed, the logs) are recomputed in registers by the line thread that uses
//     them: +22 FP64 instructions per thread against 80 fewer wavefronts per warp.
//   * Values a line thread computed once stay in registers for the flux-differencing pairs (rho, un, ut, beta) instead of
//     being re-read.
//   * Stages 2 and 3 (the stage's dt is known, KIND_S23): the two directions exchange A_d = U/2 + dt rhsxyL_d instead of
//     rhsxyL_d, so the limiter's u^L = U + dt rhsL is A_x + A_y and the new state is a resW + b (t_x + t_y) with
//     t_d = A_d + dt (limited flux differences)_d / wJ: neither the limiter nor the output phase re-reads U.
//   * The "is any pair off logmean's series branch" vote is taken by the line threads themselves (high words of rho and
//     beta against the element's first node), warp-uniform, so quiet warps never touch a log.
//
// Reference: low_order_graph_viscosity.jl:4-243, flux_differencing.jl:4-361, subcell.jl:163-349,418-456,841-924,
// SSPRK33.jl:31-39 (the same lines as stage_fast.cuh; see there and DESIGN.md for the exact-zero end faces).
#pragma once
#include "stage_fast.cuh"

namespace p2de {

// How much of the stage's role is a compile-time fact:
//   KIND_RT   run-time flags of StageArgs (p2de_rhs: any stage index, optional diagnostics, optional fused combine);
//             exchanges plain low-order shares and writes rhsU (or the fused combine)
//   KIND_S1   stage 1 of p2de_ssp33_step: CFL reduction; the limiter's dt is the cap (rhs.jl:46,52), the step's dt does not
//             exist before the whole grid has finished, so the kernel writes W = U + cap rhsU and whoever forms
//             U1 = U + dt rhsU (SSPRK33.jl:31-33) takes the convex combination U + (dt / cap) (W - U)
//   KIND_S23  stages 2 and 3: the SSP combine a resW + b (U + dt rhsU) is fused, no CFL reduction
enum { KIND_RT = 0, KIND_S1 = 1, KIND_S23 = 2 };

// U + theta (W - U) at one node: the stage-1 combine from stage 1's W = U + cap rhsU, theta = dt / cap in (0, 1]
P2DE_DEV Cons2 load_cons_plus(const double *pu, const double *pw, double theta) {
  Cons2 U = load_cons(pu);
  const Cons2 W = load_cons(pw);
  U.rho = fma(theta, W.rho - U.rho, U.rho); U.m1 = fma(theta, W.m1 - U.m1, U.m1);
  U.m2 = fma(theta, W.m2 - U.m2, U.m2); U.E = fma(theta, W.E - U.E, U.E);
  return U;
}

#ifndef P2DE_SUB_MIN_BLOCKS
#define P2DE_SUB_MIN_BLOCKS 4
#endif
#ifndef P2DE_SUB_SMOOTH_PAIRS
#define P2DE_SUB_SMOOTH_PAIRS 1   // warps whose elements vary by less than ~9 %: log-free two-point flux (series of the log mean)
#endif
#ifndef P2DE_SUB_MIN_BLOCKS_S23
#define P2DE_SUB_MIN_BLOCKS_S23 P2DE_SUB_MIN_BLOCKS   // stages 2, 3 (their shared memory allows a fifth CTA per SM)
#endif
#ifndef P2DE_SUB_MIN_BLOCKS5
#define P2DE_SUB_MIN_BLOCKS5 3
#endif

// doubles of shared memory per element besides the table prefix: U [4][Nq], the low-order shares [2][2][Nq] double2,
// the limited shares [2][2][Nq] double2, the CFL lambda sums [2][Nq] / L_local staging [2 N1D (N1D+1)]
// The kernels of the direct schedule (A-form exchange) do not need U after the line phase: their L_local staging lives in
// the dead U block, and stage 1's lambda sums in the not yet written limited-share block
template <int N1D>
__host__ __device__ constexpr int subcell_smem_doubles_per_elem(bool aform) {
  return N1D * N1D * (4 + 8 + 8) + (aform ? 0 : fast_lamp_per_elem<N1D>());
}
static_assert(2 * 4 * 5 <= 4 * 16 && 2 * 5 * 6 <= 4 * 25 && 2 * 3 * 4 <= 4 * 9 && 2 * 2 * 3 <= 4 * 4, "L_local staging fits the U block");

// log for the line threads of a non-quiet warp (operands are positive normals: rho, beta)
P2DE_DEV double log_pos(double x) { return log(x); }

template <int N1D, int EPB, bool INTERIOR, bool DEFER, int KIND>
__device__ __forceinline__ void stage_subcell_impl(const StageArgs &A, const MeshTopo &M, const Tables2D<N1D> &Tc, const long long kb) {
  constexpr int Nq = N1D * N1D, NF = N1D + 1, HALF = EPB * N1D, NT = 2 * HALF;
  constexpr int TBL = fast_table_doubles<N1D>();
  constexpr int S = EPB * Nq;
  constexpr int NL = 2 * N1D * NF;
  // A-form exchange (see the header): the limiter's dt multiplies everything the kernel writes
  constexpr bool AFORM = KIND == KIND_S23 || KIND == KIND_S1;
  constexpr bool DIAG = KIND == KIND_RT;
  extern __shared__ double sm[];
  Tables2D<N1D> &T = *reinterpret_cast<Tables2D<N1D> *>(sm);
  double *nodes = sm + TBL;                                            // [4][S] rho, m1, m2, E at swizzled positions
  double2 *partsL = reinterpret_cast<double2 *>(nodes + 4 * S);        // [d][half][S] low-order shares (or A_d)
  double2 *tbuf = partsL + 4 * S;                                      // [d][half][S] limited shares
  // [2][S] CFL sums and [EPB][NL] L_local staging.  AFORM: the sums sit in the limited-share block (written only after the
  // CFL block's barrier) and the staging in the U block (dead after the line phase)
  double *lamp = AFORM ? reinterpret_cast<double *>(tbuf) : reinterpret_cast<double *>(tbuf + 4 * S);
  double *lstage = AFORM ? nodes : lamp;

  const int tid = threadIdx.x;
  const int d = tid / HALF, rr = tid % HALF, el = rr / N1D, line = rr % N1D;
  const long long k = kb + el;
  const bool active = INTERIOR || k < M.K;
  const bool full = INTERIOR || kb + EPB <= M.K;
  const double gamma = A.gamma, gm1 = A.gamma - 1.0;
  const double *Ubase = A.Uq + kb * (Nq * 4);
  const bool nst1 = KIND == KIND_S1 || (KIND == KIND_RT && A.nstage == 1);
  const bool fuse = KIND == KIND_S23 || (KIND == KIND_RT && A.fuse != 0);
  if (A.dbg && tid == 0) {   // p2de_debug_counters: which instantiation this CTA runs
    atomicAdd(A.dbg + (INTERIOR ? DBG_CTA_INTERIOR : DBG_CTA_GENERAL), 1ull);
    if (DEFER) atomicAdd(A.dbg + DBG_CTA_DEFER, 1ull);
  }
  if (P2DE_FAST_PREFETCH) {   // the batch one wave of resident CTAs ahead, into L2 (stage_fast.cuh)
    constexpr int AHEAD = 148 * (N1D == 5 ? P2DE_SUB_MIN_BLOCKS5 : P2DE_SUB_MIN_BLOCKS * (16 / (EPB < 16 ? EPB : 16)));
    const long long kp = kb + (long long)AHEAD * EPB;
    constexpr int LINES = EPB * Nq * 32 / 128;
    if (kp + EPB <= M.K) {
      if (tid < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.Uq + kp * (Nq * 4) + tid * 16));
      else if (DEFER && tid < 2 * LINES)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.defer_add + kp * (Nq * 4) + (tid - LINES) * 16));
      else if (fuse && tid < 2 * LINES)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.fuse_resW + kp * (Nq * 4) + (tid - LINES) * 16));
    }
  }
  // table prefix: loads issued now, stored to shared memory after the state loads have been issued too
  constexpr int TF2 = (Tables2D<N1D>::FAST_BYTES + 15) / 16, NTL = (TF2 + NT - 1) / NT;
  double2 treg[NTL];
#pragma unroll
  for (int it = 0; it < NTL; ++it) {
    const int i = tid + it * NT;
    if (i < TF2) treg[it] = reinterpret_cast<const double2 *>(A.tab_dev)[i];
  }
  // dt the limiter sees (rhs.jl:46,52).  Only the deferred combine needs it while loading; otherwise its (uniform) load is
  // issued after the first barrier, where its latency hides behind the line phase instead of stalling the CTA's start
  double dtl = A.dt_host;
  if (DEFER) dtl = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;
  const double theta = DEFER ? dtl * A.inv_cap : 1.0;   // stage 1 wrote W = U + cap rhsU with cap = dt_host (the same in all stages)
  // ---- the two neighbour face nodes of this line (one 32-byte node each)
  Nbr nb[2];
  Cons2 UnbC[2];
  bool nb_in_batch[2] = {false, false};
#pragma unroll
  for (int e = 0; e < 2; ++e) { UnbC[e].rho = 1.0; UnbC[e].m1 = 0.0; UnbC[e].m2 = 0.0; UnbC[e].E = 1.0; nb[e].bc = 0; nb[e].ival = nullptr; nb[e].kP = 0; nb[e].fP = 0; }
  if (INTERIOR) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      // partner node: d=0: (N1D-1, line) of k-1 / (0, line) of k+1;  d=1: (line, N1D-1) of k-Kx / (line, 0) of k+Kx
      const int node = d == 0 ? (e ? 0 : N1D - 1) + line * N1D : line + (e ? 0 : N1D - 1) * N1D;
      const int dk = d == 0 ? (e ? 1 : -1) : (e ? M.Kx : -M.Kx);
      nb_in_batch[e] = d == 0 && (e ? el + 1 < EPB : el > 0);   // read from shared memory after the barrier
      if (!nb_in_batch[e]) {
        const long long off = ((long long)(el + dk) * Nq + node) * 4;
        UnbC[e] = DEFER ? load_cons_plus(Ubase + off, A.defer_add + kb * (Nq * 4) + off, theta) : load_cons(Ubase + off);
      }
    }
  } else if (active) {
    int ix, iy;
    if (M.K < 0x7fffffffll) { iy = (int)((unsigned)k / (unsigned)M.Kx); ix = (int)((unsigned)k - (unsigned)iy * (unsigned)M.Kx); }
    else { ix = (int)(k % M.Kx); iy = (int)(k / M.Kx); }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      nb[e] = neighbor<N1D>(M, k, ix, iy, (2 * d + e) * N1D + line);
      const long long noff = (nb[e].kP * Nq + Tc.fq2q[nb[e].fP]) * 4;
      UnbC[e] = DEFER ? load_cons_plus(A.Uq + noff, A.defer_add + noff, theta) : load_cons(A.Uq + noff);
    }
  }
  // ---- the batch's states: flat coalesced loads, stored at their swizzled positions
  constexpr int NITN = (S + NT - 1) / NT;
  Cons2 Uraw[NITN];
#pragma unroll
  for (int it = 0; it < NITN; ++it) {
    const int n = tid + it * NT;
    Uraw[it].rho = 1.0; Uraw[it].m1 = 0.0; Uraw[it].m2 = 0.0; Uraw[it].E = 1.0;   // partial batch: harmless dummy state
    if (n < S && (full || kb + n / Nq < M.K)) {
      Uraw[it] = DEFER ? load_cons_plus(Ubase + n * 4, A.defer_add + kb * (Nq * 4) + n * 4, theta) : load_cons(Ubase + n * 4);
      if (fuse && !DEFER)   // the flat output phase of this same thread reads resW here: pull it into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.fuse_resW + kb * (Nq * 4) + n * 4));
    }
  }
#pragma unroll
  for (int it = 0; it < NTL; ++it) {
    const int i = tid + it * NT;
    if (i < TF2) reinterpret_cast<double2 *>(sm)[i] = treg[it];
  }
#pragma unroll
  for (int it = 0; it < NITN; ++it) {
    const int n = tid + it * NT;
    if (n < S) {
      const int e2 = n / Nq, node = n % Nq;
      double *o = nodes + node_pos<N1D>(e2, node % N1D, node / N1D);
      o[0 * S] = Uraw[it].rho; o[1 * S] = Uraw[it].m1; o[2 * S] = Uraw[it].m2; o[3 * S] = Uraw[it].E;
    }
  }
  __syncthreads();
  if (!DEFER) dtl = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;

  // ---- line phase (all threads; threads of a partial batch's missing elements work on the dummy state and store nothing)
  int pos[N1D];
#pragma unroll
  for (int a = 0; a < N1D; ++a) pos[a] = line_pos<N1D>(el, d, line, a);
  const double *rwJ = T.rwJl[d][line];   // 1 / (Jq wq) of this line's nodes
  double G[N1D][4];                      // wJ (rhsxyH - rhsxyL) along this line
  double dF0[4] = {0.0, 0.0, 0.0, 0.0};
  PrimR q[N1D];
  {
    // Streaming over the line's nodes: node a is derived, its pair with node a-1 is formed, node a-1 is then final and
    // published.  Only a two-node window of (U, flux, wavespeed, low-order sum) is live besides the outputs q and G.
    ConsR Unb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (nb_in_batch[e]) {   // x-neighbour inside this batch
        const double *o = nodes + node_pos<N1D>(e ? el + 1 : el - 1, e ? 0 : N1D - 1, line);
        UnbC[e].rho = o[0 * S]; UnbC[e].m1 = o[1 * S]; UnbC[e].m2 = o[2 * S]; UnbC[e].E = o[3 * S];
      }
      Unb[e].rho = UnbC[e].rho; Unb[e].mn = d ? UnbC[e].m2 : UnbC[e].m1;
      Unb[e].mt = d ? UnbC[e].m1 : UnbC[e].m2; Unb[e].E = UnbC[e].E;
    }
    // primitives, beta, axis wavespeed and flux of one node (pfun :24-28, betafun :42-45, wavespeed_estimate :58-62,
    // fluxes :175-194), in registers
    auto derive = [&](int a, ConsR &U, double fl[4], double &ws) {
      const double *o = nodes + pos[a];
      U.rho = o[0 * S]; U.mn = o[(1 + d) * S]; U.mt = o[(2 - d) * S]; U.E = o[3 * S];
      const double rinv = rcp_fast(U.rho);
      const double un = U.mn * rinv, ut = U.mt * rinv;
      const double hn = 0.5 * (U.mn * U.mn) * rinv;
      const double p = gm1 * (U.E - fma(0.5 * U.mt, ut, hn));
      q[a].rho = U.rho; q[a].un = un; q[a].ut = ut; q[a].beta = 0.5 * U.rho * rcp_fast(p);
      ws = fabs(un) + sqrt_newton(gamma * (gm1 * (U.E - hn)) * rinv);
      flux_rot(U, un, ut, p, fl);
    };
    // surface flux at line end e against the neighbour state (low_order_graph_viscosity.jl:168-204,
    // flux_differencing.jl:90-151,223-272): returns -BF_L's contribution to the low-order sum in Fc
    auto face = [&](int e, const ConsR &U, const double fl[4], double ws, double Fc[4], double &lamB_out, double Gadj[4]) {
      const double B = T.Bf[d][line][e], nn = fabs(B), hB = 0.5 * B;
      double rinvP = rcp_fast(Unb[e].rho);
      const double wsP = wavespeed_rot(gamma, gm1, rinvP, Unb[e].mn, Unb[e].E);
      const double lamB = 0.5 * nn * jl_max(ws, wsP);   // (Julia's max: a NaN wavespeed reaches the CFL dt)
      ConsR uP = Unb[e];
      const int bce = INTERIOR ? 0 : nb[e].bc;
      if (bce) {
        if (bce == 1) { const double *p = nb[e].ival; uP.rho = p[0]; uP.mn = p[1 + d]; uP.mt = p[2 - d]; uP.E = p[3]; }
        else uP = U;
        rinvP = rcp_fast(uP.rho);
      }
      double fP[4];
      flux_rot(uP, uP.mn * rinvP, uP.mt * rinvP, gm1 * (uP.E - 0.5 * (uP.mn * uP.mn + uP.mt * uP.mt) * rinvP), fP);
      const double up[4] = {uP.rho, uP.mn, uP.mt, uP.E}, uf[4] = {U.rho, U.mn, U.mt, U.E};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double lf = lamB * (up[c] - uf[c]);
        Fc[c] = fma(-hB, fl[c] + fP[c], lf);           // - BF_L = -(B (f + f_P) / 2 - lf)
        Gadj[c] = bce ? lf : 0.0;                      // BF_H - BF_L: zero unless the face carries a boundary condition
      }
      lamB_out = lamB;
    };
    // node a is final: its G starts as the volume part of -GL (the surface terms cancel identically on interior faces,
    // identity projection, and leave the LF term on inflow/outflow faces), its low-order share is published
    // (scale_low_order_rhs_by_mass! :206-220; AFORM: A_d = U/2 + dt rhsxyL_d, so that A_x + A_y = u^L, subcell.jl:269)
    auto publish = [&](int a, const ConsR &U, const double GLv[4], const double Fc[4], const double Gadj[4], bool has_face, double lam_sum) {
      double tot[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        G[a][c] = has_face ? -GLv[c] - Gadj[c] : -GLv[c];
        tot[c] = has_face ? GLv[c] + Fc[c] : GLv[c];
      }
      if (AFORM) {
        const double w = dtl * rwJ[a];
        partsL[(d * 2 + 0) * S + pos[a]] = make_double2(fma(w, tot[0], 0.5 * U.rho), fma(w, tot[1], 0.5 * U.mn));
        partsL[(d * 2 + 1) * S + pos[a]] = make_double2(fma(w, tot[2], 0.5 * U.mt), fma(w, tot[3], 0.5 * U.E));
      } else {
        partsL[(d * 2 + 0) * S + pos[a]] = make_double2(tot[0] * rwJ[a], tot[1] * rwJ[a]);
        partsL[(d * 2 + 1) * S + pos[a]] = make_double2(tot[2] * rwJ[a], tot[3] * rwJ[a]);
      }
      if (nst1) lamp[d * S + pos[a]] = lam_sum;   // this direction's share of lambda_i (:222-281): its volume pairs and its face
    };
    ConsR Uc, Up;
    double flc[4], flp[4], wsc, wsp, GLc[4], GLp[4], Fc0[4], Gadj0[4], lamF0, lam_prev = 0.0;
    derive(0, Uc, flc, wsc);
    face(0, Uc, flc, wsc, Fc0, lamF0, Gadj0);
#pragma unroll
    for (int c = 0; c < 4; ++c) { GLc[c] = 0.0; dF0[c] = Gadj0[c]; }   // dF0 = BF_H - BF_L on the seed face
#pragma unroll
    for (int a = 1; a < N1D; ++a) {
      Up = Uc; wsp = wsc;
#pragma unroll
      for (int c = 0; c < 4; ++c) { flp[c] = flc[c]; GLp[c] = GLc[c]; GLc[c] = 0.0; }
      derive(a, Uc, flc, wsc);
      // low-order graph-viscosity pair (a, a-1), low_order_graph_viscosity.jl:139-166
      const double Sv = T.S0[d][line][a - 1];
      const double lam = fabs(Sv) * jl_max(wsc, wsp);
      const double ui[4] = {Uc.rho, Uc.mn, Uc.mt, Uc.E}, uj[4] = {Up.rho, Up.mn, Up.mt, Up.E};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double SF = Sv * (flc[c] + flp[c]) - lam * (uj[c] - ui[c]);   // 2 Sv (f_i + f_j)/2: the scalings by 2 are exact
        GLc[c] -= SF; GLp[c] += SF;          // GL = -Q0F1
      }
      publish(a - 1, Up, GLp, Fc0, Gadj0, a == 1, (lam_prev + lam) + (a == 1 ? lamF0 : 0.0));
      lam_prev = lam;
    }
    double Fc1[4], Gadj1[4], lamF1;
    face(1, Uc, flc, wsc, Fc1, lamF1, Gadj1);
    publish(N1D - 1, Uc, GLc, Fc1, Gadj1, true, (lam_prev + 0.0) + lamF1);
  }
  // ---- flux differencing along this line, flux_differencing.jl:164-211 (pairs j<i, j outer).
  // Which form of the two-point flux the line's pairs need is decided from the line's own nodes (its pairs are all it
  // evaluates): if rho and beta each stay within 32 units of 2^-20 (high words) of the line's first node they vary by less
  // than 6.2e-5 relative and no pair can leave logmean's series branch (|f| < 1e-4, :307-321): "quiet", fS_rot_quiet; within
  // 0x6000 units any two nodes differ by less than 9.4 %, so |z| = |da| / (a_L + a_R) < 0.05 for every pair and the series
  // of fS_rot_smooth applies: "smooth".  Neither evaluates a log.  The choice is made per warp (all its lanes vote), so
  // the three variants never diverge inside a warp.
  bool quiet = false, smooth = false;
  {
    const int rr0 = __double2hiint(q[0].rho), rb0 = __double2hiint(q[0].beta);
    bool far = false, far2 = false;
#pragma unroll
    for (int a = 1; a < N1D; ++a) {
      const int er = __double2hiint(q[a].rho) - rr0, eb = __double2hiint(q[a].beta) - rb0;
      far = far | ((unsigned)(er + 32) > 64u) | ((unsigned)(eb + 32) > 64u);
      far2 = far2 | ((unsigned)(er + 0x6000) > 0xC000u) | ((unsigned)(eb + 0x6000) > 0xC000u);
    }
    const unsigned am = __activemask();   // (the CTA's last warp may be partial when 2 N1D EPB is not a multiple of 32)
    quiet = P2DE_FAST_QUIET_PAIRS && __ballot_sync(am, far) == 0u;
    smooth = P2DE_SUB_SMOOTH_PAIRS && !quiet && __ballot_sync(am, far2) == 0u;
    if (A.dbg && d == 0 && line == 0 && active) {
      atomicAdd(A.dbg + DBG_ELEM, 1ull);
      if (!quiet && !smooth) atomicAdd(A.dbg + DBG_ELEM_LOGS, 1ull);
    }
  }
  if (quiet) {
    PairLoop<N1D, 0, 1>::run([&](auto jc, auto ic) {
      constexpr int j = decltype(jc)::value, i = decltype(ic)::value;
      double F[4];
      fS_rot_quiet(A.half_inv_gm1, q[i], q[j], F);
      const double Sv = T.SHt[d][i][j][line];
#pragma unroll
      for (int c = 0; c < 4; ++c) { G[i][c] = fma(-Sv, F[c], G[i][c]); G[j][c] = fma(Sv, F[c], G[j][c]); }
    });
  } else if (smooth) {
    PairLoop<N1D, 0, 1>::run([&](auto jc, auto ic) {
      constexpr int j = decltype(jc)::value, i = decltype(ic)::value;
      double F[4];
      fS_rot_smooth(A.half_inv_gm1, q[i], q[j], F);
      const double Sv = T.SHt[d][i][j][line];
#pragma unroll
      for (int c = 0; c < 4; ++c) { G[i][c] = fma(-Sv, F[c], G[i][c]); G[j][c] = fma(Sv, F[c], G[j][c]); }
    });
  } else {
#pragma unroll
    for (int a = 0; a < N1D; ++a) { q[a].rholog = log_pos(q[a].rho); q[a].betalog = log_pos(q[a].beta); }
    PairLoop<N1D, 0, 1>::run([&](auto jc, auto ic) {
      constexpr int j = decltype(jc)::value, i = decltype(ic)::value;
      double F[4];
      fS_rot(A.half_inv_gm1, q[i], q[j], F);
      const double Sv = T.SHt[d][i][j][line];
#pragma unroll
      for (int c = 0; c < 4; ++c) { G[i][c] = fma(-Sv, F[c], G[i][c]); G[j][c] = fma(Sv, F[c], G[j][c]); }
    });
  }
  __syncthreads();

  // ---- CFL: dt = min_i CFL * 0.5 * wJ_i / lambda_i, low_order_graph_viscosity.jl:222-281
  if (nst1) {
    double dtloc = INFINITY;
    if (active && d == 0) {
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double li = lamp[0 * S + pos[a]] + lamp[1 * S + pos[a]];
        // (Newton reciprocal instead of the IEEE division routine: <= 2 ulp, dt is tested to 1e-13; li > 0 for positive states)
        const double ci = A.CFL * 0.5 * (A.Jq * T.wq[a + line * N1D]);
        dtloc = jl_min(dtloc, li > 0.0 ? ci * rcp_fast(li) : ci / li);
      }
    }
    {
      const unsigned wmask = __activemask();   // the CTA's last warp may be partial
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double other = __shfl_xor_sync(wmask, dtloc, off);
        if ((wmask >> ((tid & 31) ^ off)) & 1u) dtloc = jl_min(dtloc, other);
      }
    }
    if ((tid & 31) == 0) dt_publish(A.dt_bits, dtloc);
    __syncthreads();   // CFL block done with lamp (it becomes the L_local staging)
  }

  if (active) {
    // ---- f_bar_H - f_bar_L by prefix sum (subcell.jl:163-206) and the limiting coefficients of this line's N1D+1
    //      subcell faces (subcell.jl:248-349).  End faces: exact zero unless the face carries an inflow/outflow
    //      condition (stage_fast.cuh "End faces"), so their coefficients are 1 from both sides.
    double dFv[NF][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) dFv[0][c] = INTERIOR ? 0.0 : dF0[c];
#pragma unroll
    for (int s = 1; s < NF; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) dFv[s][c] = (INTERIOR && s == 1) ? G[0][c] : dFv[s - 1][c] + G[s - 1][c];
    const bool bc0 = !INTERIOR && nb[0].bc != 0, bc1 = !INTERIOR && nb[1].bc != 0;
    if (!bc1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) dFv[N1D][c] = 0.0;
    }
    double lv[NF];
#pragma unroll
    for (int s = 0; s < NF; ++s) lv[s] = 1.0;
    // u^L = Uq + dt rhsL of node a in this line's rotated frame (the other direction's share is in ITS rotated frame:
    // momentum components swap); r[] = rhsL where it is formed (not AFORM)
    auto low_state = [&](int pa, double r[4]) -> Cons2 {
      const double2 m0 = partsL[(d * 2 + 0) * S + pa], m1 = partsL[(d * 2 + 1) * S + pa];
      const double2 o0 = partsL[((1 - d) * 2 + 0) * S + pa], o1 = partsL[((1 - d) * 2 + 1) * S + pa];
      r[0] = m0.x + o0.x; r[1] = m0.y + o1.x; r[2] = m1.x + o0.y; r[3] = m1.y + o1.y;
      Cons2 uL;
      if (AFORM) { uL.rho = r[0]; uL.m1 = r[1]; uL.m2 = r[2]; uL.E = r[3]; }
      else {
        const double *o = nodes + pa;
        uL.rho = o[0 * S] + dtl * r[0]; uL.m1 = o[(1 + d) * S] + dtl * r[1];
        uL.m2 = o[(2 - d) * S] + dtl * r[2]; uL.E = o[3 * S] + dtl * r[3];
      }
      return uL;
    };
    // First pass, branch-free and division-free: is EVERY coefficient of this line certainly 1?  (stage_fast.cuh has the
    // derivation: 2 rho_L q(1) > margin and rho' - zeta rho_L > margin, rho e concave, so no root of the reference's
    // quadratic in (0, 1]; everything else goes to the exact evaluation below.)
    bool all_easy = true;
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      double r[4];
      const Cons2 uL = low_state(pos[a], r);
      const double kk = 4 * dtl * rwJ[a];     // P = -/+ 4 dt (fH - fL) / wJ, subcell.jl:300,312,328,340
      const double r2L = 2.0 * uL.rho, eL = r2L * uL.E;
      const double cL = fma(-uL.m2, uL.m2, fma(-uL.m1, uL.m1, eL));   // 2 rho rho e of u^L
      const double zc = A.zeta * cL, tolq = (1e-9 * eL) * uL.rho, tolr = 1e-9 * uL.rho;
      all_easy = all_easy & (cL > 1e-9 * eL) & (uL.rho > 0.0);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? (a > 0 || bc0) : (a < N1D - 1 || bc1)) {
          const double ks = side ? kk : -kk;
          const double *dFs = dFv[a + side];
          const double rp = fma(ks, dFs[0], uL.rho), m1p = fma(ks, dFs[1], uL.m1), m2p = fma(ks, dFs[2], uL.m2), Ep = fma(ks, dFs[3], uL.E);
          const double t1 = fma(r2L, Ep, -zc), msq = fma(m2p, m2p, m1p * m1p);
          const double q1 = fma(-msq, uL.rho, rp * t1);
          all_easy = all_easy & (fma(-A.zeta, uL.rho, rp) > tolr) & (q1 > tolq);
        }
      }
      // this line's limited share, written here on the assumption that every coefficient of the line is 1 (then the
      // increment l_{a+1} dF_{a+1} - l_a dF_a is the prefix sum's own G[a]); each direction adds half of the common part
      // r (AFORM: u^L, otherwise rhsL), so t_x + t_y = r + (increment_x + increment_y) w.  Rewritten below otherwise.
      {
        const double w = AFORM ? dtl * rwJ[a] : rwJ[a];
        tbuf[(d * 2 + 0) * S + pos[a]] = make_double2(fma(G[a][0], w, 0.5 * r[0]), fma(G[a][1], w, 0.5 * r[1]));
        tbuf[(d * 2 + 1) * S + pos[a]] = make_double2(fma(G[a][2], w, 0.5 * r[2]), fma(G[a][3], w, 0.5 * r[3]));
      }
      if (DIAG) {
        if (d == 0 && A.rhsL_diag) store4(A.rhsL_diag + (k * Nq + a + line * N1D) * 4, r);
        if (A.rhsH_diag) {   // diagnostics: rhsxyH_d = rhsxyL_d + G / wJ; each line adds its share (buffer pre-zeroed)
          const int node = d == 0 ? a + line * N1D : line + a * N1D;
          double *hd = A.rhsH_diag + (k * Nq + node) * 4;
          const double2 m0 = partsL[(d * 2 + 0) * S + pos[a]], m1 = partsL[(d * 2 + 1) * S + pos[a]];
          atomicAdd(hd + 0, m0.x + G[a][0] * rwJ[a]); atomicAdd(hd + 1 + d, m0.y + G[a][1] * rwJ[a]);
          atomicAdd(hd + 2 - d, m1.x + G[a][2] * rwJ[a]); atomicAdd(hd + 3, m1.y + G[a][3] * rwJ[a]);
        }
      }
    }
    if (A.dbg) { atomicAdd(A.dbg + DBG_LINES, 1ull); if (!all_easy) atomicAdd(A.dbg + DBG_LINES_NOT_EASY, 1ull); }
    if (!all_easy) {
      // exact evaluation (the reference's formulas: quadratic coefficients, root selection), node by node
      bool one = true;
#pragma unroll 1
      for (int a = 0; a < N1D; ++a) {
        double r[4];
        const Cons2 uL = low_state(line_pos<N1D>(el, d, line, a), r);
        // rhoe_ufun (:75-78) with a Newton reciprocal; c = E rho - |m|^2/2 - rho Lrhoe = (1 - zeta) rho rhoe
        const double rhoeL = uL.E - 0.5 * (uL.m1 * uL.m1 + uL.m2 * uL.m2) * rcp_fast(uL.rho);
        const double Lrho = A.zeta * uL.rho, Lrhoe = A.zeta * rhoeL;
        const double c0 = (1.0 - A.zeta) * uL.rho * rhoeL;
        const double kk = 4 * dtl * rwJ[a];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          if (side == 0 ? (a > 0 || bc0) : (a < N1D - 1 || bc1)) {
            double Pv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) Pv[c] = (side ? kk : -kk) * dFv_at(dFv, a + side, c);
            double qa, qb;
            quad_coeff_ab(uL, Pv, Lrhoe, qa, qb);
            // (lv <= 1 throughout, so the common result 1.0 of limiting_param_pos needs no min)
            if (!limiting_param_pos_easy(uL.rho, Pv[0], Lrho, qa, qb, c0)) {
              if (A.dbg) atomicAdd(A.dbg + DBG_LIMITER_SLOW, 1ull);
              const double lnew = limiting_param_pos_slow(A.ZEROTOL, uL.rho, Pv[0], Lrho, qa, qb, c0);
              lv_set_min(lv, a + side, lnew);
              one = one & (lnew >= 1.0);
            }
          }
        }
      }
      all_easy = one;   // a line the margin sent here may still have all coefficients 1
    }
    // this line's share of the un-symmetrised limited rhs (subcell.jl:841-924 with the line's own coefficients), in the
    // line's rotated frame:  t_d = r / 2 + (l_{a+1} dF_{a+1} - l_a dF_a) w,  w = 1 / wJ (AFORM: dt / wJ).  Only lines with a
    // coefficient below 1 get here; all others wrote their share in the first pass.
    if (!all_easy) {
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        double r[4];
        low_state(pos[a], r);
        const double w = AFORM ? dtl * rwJ[a] : rwJ[a];
        double inc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // (INTERIOR: dF is exactly zero on the two end faces, the products are dropped at compile time)
          const double hi = (INTERIOR && a == N1D - 1) ? 0.0 : lv[a + 1] * dFv[a + 1][c];
          const double lo = (INTERIOR && a == 0) ? 0.0 : lv[a] * dFv[a][c];
          inc[c] = hi - lo;
        }
        tbuf[(d * 2 + 0) * S + pos[a]] = make_double2(fma(inc[0], w, 0.5 * r[0]), fma(inc[1], w, 0.5 * r[1]));
        tbuf[(d * 2 + 1) * S + pos[a]] = make_double2(fma(inc[2], w, 0.5 * r[2]), fma(inc[3], w, 0.5 * r[3]));
      }
    }
#pragma unroll
    for (int s = 0; s < NF; ++s) lstage[el * NL + d * (N1D * NF) + (d == 0 ? s + line * NF : line + s * N1D)] = lv[s];
  }   // active
  // ---- flat, coalesced output phase: x share + y share (y share un-rotated), lpre
  constexpr int NIT = (S + NT - 1) / NT;
  double2 wres[NIT][2];
  {
    const double *rw = A.fuse_resW + kb * (Nq * 4);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {   // resW of this thread's nodes: in flight across the barrier
      const int n = tid + it * NT;
      wres[it][0] = make_double2(0.0, 0.0); wres[it][1] = make_double2(0.0, 0.0);
      if (fuse && n < S && (full || kb + n / Nq < M.K)) {
        const double2 *qq = reinterpret_cast<const double2 *>(rw + n * 4);
        wres[it][0] = qq[0]; wres[it][1] = qq[1];
      }
    }
  }
  __syncthreads();
  double *out = A.rpre + kb * (Nq * 4);
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = tid + it * NT;
    const int e2 = n / Nq, node = n % Nq;
    if (n < S && (full || kb + e2 < M.K)) {
      const int p2 = node_pos<N1D>(e2, node % N1D, node / N1D);
      const double2 x0 = tbuf[0 * S + p2], x1 = tbuf[1 * S + p2], y0 = tbuf[2 * S + p2], y1 = tbuf[3 * S + p2];
      double r[4] = {x0.x + y0.x, x0.y + y1.x, x1.x + y0.y, x1.y + y1.y};
      if (KIND == KIND_S1) {
        // t_x + t_y = W = U + cap rhsU is what stage 1 leaves behind (see KIND_S1)
      } else if (AFORM) {   // t_x + t_y = U + dt rhsU: the SSP combine (SSPRK33.jl:34-39) is one FMA per component
        r[0] = fma(A.fuse_a, wres[it][0].x, A.fuse_b * r[0]); r[1] = fma(A.fuse_a, wres[it][0].y, A.fuse_b * r[1]);
        r[2] = fma(A.fuse_a, wres[it][1].x, A.fuse_b * r[2]); r[3] = fma(A.fuse_a, wres[it][1].y, A.fuse_b * r[3]);
      } else if (KIND == KIND_RT && A.wform) {   // stage 1 in W form with diagnostics on: W = U + cap rhsU, cap = dtl
        r[0] = fma(dtl, r[0], nodes[0 * S + p2]); r[1] = fma(dtl, r[1], nodes[1 * S + p2]);
        r[2] = fma(dtl, r[2], nodes[2 * S + p2]); r[3] = fma(dtl, r[3], nodes[3 * S + p2]);
      } else if (fuse) {    // run-time version (p2de_rhs-side schedules): the combine of the un-corrected rhs
        r[0] = A.fuse_a * wres[it][0].x + A.fuse_b * (nodes[0 * S + p2] + dtl * r[0]);
        r[1] = A.fuse_a * wres[it][0].y + A.fuse_b * (nodes[1 * S + p2] + dtl * r[1]);
        r[2] = A.fuse_a * wres[it][1].x + A.fuse_b * (nodes[2 * S + p2] + dtl * r[2]);
        r[3] = A.fuse_a * wres[it][1].y + A.fuse_b * (nodes[3 * S + p2] + dtl * r[3]);
      }
      store4(out + n * 4, r);
    }
  }
  double *lout = A.lpre + kb * NL;
  if (full && (EPB * NL) % 2 == 0) {   // 16-byte copies, all loads first
    const double2 *ls2 = reinterpret_cast<const double2 *>(lstage);
    double2 *lo2 = reinterpret_cast<double2 *>(lout);
    constexpr int NC = EPB * NL / 2, NITL = (NC + NT - 1) / NT;
    double2 lreg[NITL];
#pragma unroll
    for (int it = 0; it < NITL; ++it) if (tid + it * NT < NC) lreg[it] = ls2[tid + it * NT];
#pragma unroll
    for (int it = 0; it < NITL; ++it) if (tid + it * NT < NC) lo2[tid + it * NT] = lreg[it];
  } else {
    for (int n = tid; n < EPB * NL; n += NT)
      if (kb + n / NL < M.K) lout[n] = lstage[n];
  }
}

#ifdef P2DE_SUB_MAXNREG   // experiment: register cap instead of a minimum CTA count (9 CTAs of 64 threads per SM at 112 registers)
#define P2DE_SUB_BOUNDS(KIND_) __maxnreg__(P2DE_SUB_MAXNREG)
#else
#define P2DE_SUB_BOUNDS(KIND_) __launch_bounds__(EPB * 2 * N1D, (N1D == 5 ? P2DE_SUB_MIN_BLOCKS5 : (KIND_ != KIND_RT ? P2DE_SUB_MIN_BLOCKS_S23 : P2DE_SUB_MIN_BLOCKS) * (16 / (EPB < 16 ? EPB : 16))))
#endif
#define P2DE_SUBCELL_KERNEL(NAME, DEFER_, KIND_)                                                                        \
  template <int N1D, int EPB>                                                                                           \
  __global__ void P2DE_SUB_BOUNDS(KIND_)                                                                                \
  NAME(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M, const __grid_constant__ Tables2D<N1D> Tc) { \
    bool interior;                                                                                                      \
    const long long kb = fast_batch<EPB>(A, M, interior);                                                               \
    if (interior) stage_subcell_impl<N1D, EPB, true, DEFER_, KIND_>(A, M, Tc, kb);                                      \
    else stage_subcell_impl<N1D, EPB, false, DEFER_, KIND_>(A, M, Tc, kb);                                              \
  }
// Four kernels (each its own __global__ function: further copies of the body inside one kernel made ptxas' register
// allocation for the other copies worse, measured 3 %):
P2DE_SUBCELL_KERNEL(stage_subcell_rt, false, KIND_RT)        // p2de_rhs and the testing schedules: run-time flags, diagnostics
P2DE_SUBCELL_KERNEL(stage_subcell_s1, false, KIND_S1)        // stage 1: CFL reduction, writes W = U + cap rhsU
P2DE_SUBCELL_KERNEL(stage_subcell_s2, true, KIND_S23)        // stage 2: forms U1 = U^n + dt rhsU while loading, writes U2
P2DE_SUBCELL_KERNEL(stage_subcell_s3, false, KIND_S23)       // stage 3: writes U^{n+1}
#undef P2DE_SUBCELL_KERNEL

}  // namespace p2de
