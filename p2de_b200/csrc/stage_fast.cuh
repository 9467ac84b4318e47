// stage_fast.cuh — stage kernel for the default flux configuration (FAST variant, see
// kernels2d.cuh for the scheme and the line-thread mapping).  Same results as the generic
// stage_kernel within a few ulp; differences are purely organisational:
//
//   * warps are direction-homogeneous (first half of the CTA = x-lines, second half = y-lines) and
//     every line works in a frame rotated to its own axis (normal momentum first), so the axis is
//     an address offset instead of per-pair selects;
//   * node fields live in shared memory in a swizzled order that is bank-conflict free for both
//     x-line and y-line warps (N1D = 4);
//   * 1/x and sqrt are MUFU seed + Newton steps without the IEEE special-case branches, which
//     lets ptxas interleave the independent node-pair chains (operands are positive normals);
//   * with the identity LGL projection the low- and high-order surface fluxes coincide, so an
//     interior face contributes nothing to f_H - f_L; only inflow/outflow faces do;
//   * phases are ordered so that at most one 4-node accumulator set is live at a time;
//   * that exact zero also makes the interface coefficients 1 on both sides of every interior face, so the
//     limiter's symmetrisation is the identity and no second kernel follows (see "End faces" below);
//   * the body is compiled per CTA kind: INTERIOR batches (strictly inside a structured mesh) have no
//     boundary-condition code at all; DEFER (its own kernel) forms the stage-1 SSP combine while loading;
//   * an L2 prefetch of the batch one wave of resident CTAs ahead, table loads overlapped with the state loads,
//     a branch-free first pass of the limiter over the whole line (profiles/README.md has each step's measurement).
#pragma once
#include <type_traits>
#include "kernels2d.cuh"

namespace p2de {

#ifndef P2DE_FAST_MIN_BLOCKS
#define P2DE_FAST_MIN_BLOCKS 4
#endif
#ifndef P2DE_FAST_LIMITER_TWO_PASS
#define P2DE_FAST_LIMITER_TWO_PASS 1
#endif
#ifndef P2DE_FAST_PREFETCH
#define P2DE_FAST_PREFETCH 1
#endif
#ifndef P2DE_FAST_MIN_BLOCKS5
#define P2DE_FAST_MIN_BLOCKS5 3   // N=4 (N1D=5): 168 registers, no spills
#endif
// A/B switches of this round's restructurings (profiles/README.md has each step's measurement)
#ifndef P2DE_FAST_SURE_LIMITER
#define P2DE_FAST_SURE_LIMITER 1   // division-free sufficient test "every coefficient of the line is 1" before the exact evaluation
#endif
#ifndef P2DE_FAST_QUIET_PAIRS
#define P2DE_FAST_QUIET_PAIRS 1    // warps whose elements cannot leave logmean's series branch: two-point flux without logs / selects
#endif
#ifndef P2DE_FAST_KINDS
#define P2DE_FAST_KINDS 1          // stage role (stage 1 / stages 2,3) known at compile time in the direct schedule's kernels
#endif

// How much of the stage's role is a compile-time fact (MODE_SUBCELL; the other modes use KIND_RT):
//   KIND_RT   run-time flags of StageArgs (p2de_rhs: any stage index, optional diagnostics, optional fused combine)
//   KIND_S1   stage 1 of the direct schedule: CFL reduction, writes rhsU, no diagnostics
//   KIND_S23  stages 2 and 3 of the direct schedule: SSP combine fused, no CFL reduction, no diagnostics
enum { KIND_RT = 0, KIND_S1 = 1, KIND_S23 = 2 };

// constants of logmean's series branch (:315-317) and its reciprocal: read as constant-bank operands
__constant__ double kSeries[5] = {-0.2, 0.0512, 0.026038857142857, 0.2, 0.0912};

// U + dt * R at one node (the deferred stage-1 combine, StageArgs.defer_add)
P2DE_DEV Cons2 load_cons_plus(const double *pu, const double *pr, double dt) {
  Cons2 U = load_cons(pu);
  const Cons2 R = load_cons(pr);
  U.rho = fma(dt, R.rho, U.rho); U.m1 = fma(dt, R.m1, U.m1); U.m2 = fma(dt, R.m2, U.m2); U.E = fma(dt, R.E, U.E);
  return U;
}

// state in the frame of one axis: (rho, normal momentum, tangential momentum, E)
struct ConsR { double rho, mn, mt, E; };
struct PrimR { double rho, un, ut, beta, rholog, betalog; };

// sqrt: MUFU.RSQ64H seed (>= 20 bits) + two coupled Newton steps; relative error ~1e-23 before the final rounding, i.e.
// within 1 ulp (the residual correction of sqrt_fast only decides the last bit)
P2DE_DEV double sqrt_newton(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  return fma(g, r, g);
}
P2DE_DEV double wavespeed_rot(double gamma, double gm1, double rinv, double mn, double E) {
  double p = gm1 * (E - 0.5 * (mn * mn) * rinv);
  return fabs(mn * rinv) + sqrt_newton(gamma * p * rinv);
}
// normal flux of fluxes(::Dim2) (:175-194) in the rotated frame
P2DE_DEV void flux_rot(const ConsR &U, double un, double ut, double p, double f[4]) {
  f[0] = U.mn; f[1] = U.mn * un + p; f[2] = U.mn * ut; f[3] = un * (U.E + p);   // (rho un ut = mn ut up to one rounding)
}
// three independent quotients n_i / a_i, written in lock step so that the three MUFU + Newton
// chains overlap in the instruction stream (ptxas keeps source order for straight-line code)
P2DE_DEV void div3_fast(double n0, double a0, double n1, double a1, double n2, double a2,
                        double &q0, double &q1, double &q2) {
  double x0, x1, x2;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a0));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x1) : "d"(a1));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x2) : "d"(a2));
  double e0 = fma(-a0, x0, 1.0), e1 = fma(-a1, x1, 1.0), e2 = fma(-a2, x2, 1.0);
  double t0 = fma(e0, e0, e0), t1 = fma(e1, e1, e1), t2 = fma(e2, e2, e2);
  x0 = fma(x0, t0, x0); x1 = fma(x1, t1, x1); x2 = fma(x2, t2, x2);
  q0 = n0 * x0; q1 = n1 * x1; q2 = n2 * x2;   // <= ~1.5 ulp each; no residual correction needed at 1e-12
}

// fS (:220-249) along the line's own axis, three reciprocals (see fS_fast in physics.cuh)
P2DE_DEV void fS_rot(double half_inv_gm1, const PrimR &L, const PrimR &R, double F[4]) {
  double da = R.rho - L.rho, aavg = 0.5 * (R.rho + L.rho);
  bool ser = fabs(da) < 1e-4 * fabs(aavg);
  double db = R.beta - L.beta, bavg = 0.5 * (R.beta + L.beta);
  bool serb = fabs(db) < 1e-4 * fabs(bavg);
  double q, qb, pa;
  div3_fast(da, ser ? aavg : (R.rholog - L.rholog),
            serb ? 1.0 : (R.betalog - L.betalog), serb ? bavg : db,
            aavg, L.beta + R.beta, q, qb, pa);
  double v = q * q;
  double rholog = ser ? aavg * (1 + v * (kSeries[0] - v * (kSeries[1] - v * kSeries[2]))) : q;
  double fb = db * qb, vb = fb * fb;
  double inv_betalog = serb ? qb * (1 + vb * (kSeries[3] + vb * kSeries[4])) : qb;
  double unavg = 0.5 * (L.un + R.un), utavg = 0.5 * (L.ut + R.ut);
  double unorm = L.un * R.un + L.ut * R.ut;
  double f4aux = rholog * inv_betalog * half_inv_gm1 + pa + 0.5 * rholog * unorm;
  double F1 = rholog * unavg;
  F[0] = F1; F[1] = F1 * unavg + pa; F[2] = F1 * utavg; F[3] = f4aux * unavg;
}

// The same flux for a pair of an element in which rho and beta each vary by less than 6.2e-5 relative (the LAZY_LOGS vote
// of the node phase): both logmeans are on their series branch, where with f = da/aavg, v = f^2 < 4e-9
//   logmean(rho)      = aavg (1 - v/5 - O(v^2))        (the dropped terms are < 1e-18 relative)
//   1 / logmean(beta) = (1 + v_b/5 + O(v_b^2)) / bavg
// so f is only needed to ~1e-5 relative: the raw MUFU reciprocal (>= 20 bits) without Newton steps, and 1/bavg is the
// reciprocal pa = aavg / (beta_L + beta_R) needs anyway.  One refined reciprocal, no logs, no selects.
P2DE_DEV void fS_rot_quiet(double half_inv_gm1, const PrimR &L, const PrimR &R, double F[4]) {
  const double sa = R.rho + L.rho, da = R.rho - L.rho, sb = R.beta + L.beta, db = R.beta - L.beta;
  double xa;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(xa) : "d"(sa));
  const double y = rcp_fast(sb);                       // 1 / (beta_L + beta_R) = 1 / (2 bavg)
  const double fa = da * xa, fb = db * y;              // f / 2 of rho (approximate) and of beta
  const double rholog = sa * fma(-0.4, fa * fa, 0.5);  // aavg (1 - f^2 / 5),  f = 2 fa, aavg = sa / 2
  const double inv_betalog = y * fma(1.6, fb * fb, 2.0);   // (1 + f_b^2 / 5) / bavg
  const double pa = 0.5 * (sa * y);
  const double unavg = 0.5 * (L.un + R.un), utavg = 0.5 * (L.ut + R.ut);
  const double unorm = L.un * R.un + L.ut * R.ut;
  const double f4aux = fma(0.5 * rholog, unorm, fma(rholog * inv_betalog, half_inv_gm1, pa));
  const double F1 = rholog * unavg;
  F[0] = F1; F[1] = fma(F1, unavg, pa); F[2] = F1 * utavg; F[3] = f4aux * unavg;
}

// run-time indexed access to small per-line arrays inside the (rolled, rarely executed) exact limiter loop: a select chain
// keeps the arrays in registers
template <int NF>
P2DE_DEV double dFv_at(const double (&dFv)[NF][4], int s, int c) {
  double v = dFv[0][c];
#pragma unroll
  for (int t = 1; t < NF; ++t) v = (s == t) ? dFv[t][c] : v;
  return v;
}
template <int NF>
P2DE_DEV void lv_set_min(double (&lv)[NF], int s, double l) {
#pragma unroll
  for (int t = 0; t < NF; ++t) lv[t] = (s == t) ? jl_min(lv[t], l) : lv[t];
}

// compile-time loop over the node pairs (j, i), j < i, j outer: indices are constants for every N1D, so the
// per-node arrays of the caller stay in registers (a `#pragma unroll` nest is not unrolled at N1D = 5)
template <int N1D, int J, int I>
struct PairLoop {
  template <class F>
  static P2DE_DEV void run(F &&f) {
    f(std::integral_constant<int, J>{}, std::integral_constant<int, I>{});
    if constexpr (I + 1 < N1D) PairLoop<N1D, J, I + 1>::run(f);
    else if constexpr (J + 2 < N1D) PairLoop<N1D, J + 1, J + 2>::run(f);
  }
};

// shared-memory position of node (i, j) of CTA-local element el
template <int N1D>
__host__ __device__ __forceinline__ int node_pos(int el, int i, int j) {
  if (N1D == 4) return el * 16 + 4 * ((j + el) & 3) + ((i + j) & 3);
  return el * (N1D * N1D) + i + j * N1D;
}

// position of the a-th node of grid line `line` along axis d (d = 0: node (a, line); d = 1: node (line, a)); i + j = a + line
// for both directions, so only the swizzled row depends on d
template <int N1D>
__device__ __forceinline__ int line_pos(int el, int d, int line, int a) {
  if (N1D == 4) return el * 16 + 4 * (((d ? a : line) + el) & 3) + ((a + line) & 3);
  return el * (N1D * N1D) + (d ? line + a * N1D : a + line * N1D);
}

// doubles of shared memory per element, besides the tables: 12 node fields, rhsxyL shares (partsL),
// rhsxyH shares (partsH, not MODE_SUBCELL), the CFL lambda sums [2][Nq] that are later reused as the
// L_local staging [2*N1D*(N1D+1)], and lmin
template <int N1D>
__host__ __device__ constexpr int fast_lamp_per_elem() {
  return 2 * N1D * N1D > 2 * N1D * (N1D + 1) ? 2 * N1D * N1D : 2 * N1D * (N1D + 1);
}
// the FAST kernels keep only the table prefix they use in shared memory (rounded up to 16-byte words)
template <int N1D>
__host__ __device__ constexpr int fast_table_doubles() { return ((Tables2D<N1D>::FAST_BYTES + 15) / 16) * 2; }
template <int N1D, int MODE>
constexpr int fast_smem_doubles_per_elem() {
  constexpr int Nq = N1D * N1D;
  return 12 * Nq + 8 * Nq + ((MODE == MODE_SUBCELL) ? 0 : 8 * Nq) + fast_lamp_per_elem<N1D>() + N1D + 1;   // + 1: the element's "quiet" flag
}


// INTERIOR = the batch lies strictly inside a structured mesh (CTA-uniform, decided by the kernel below): the neighbours
// are k -+ 1 / k -+ Kx and no face carries a boundary condition, so the boundary-condition branches, the seed
// f_bar_H - f_bar_L of the prefix sums and the two end-face limiter evaluations of every line vanish at compile time
// (fewer live registers: the generic version spills the boundary flags across the whole kernel).
template <int N1D, int MODE, int EPB, bool INTERIOR, bool DEFER, int KIND = KIND_RT>
__device__ __forceinline__ void stage_fast_impl(const StageArgs &A, const MeshTopo &M, const Tables2D<N1D> &Tc, const long long kb) {
  constexpr int Nq = N1D * N1D, NF = N1D + 1, NFLD = 12, HALF = EPB * N1D, NT = 2 * HALF;
  constexpr bool DO_LOW = MODE != MODE_HIGH, DO_HIGH = MODE != MODE_LOW;
  constexpr int TBLC = (sizeof(Tables2D<N1D>) + 7) / 8;
  constexpr int TBL = fast_table_doubles<N1D>();
  constexpr int S = EPB * Nq;
  extern __shared__ double sm[];
  Tables2D<N1D> &T = *reinterpret_cast<Tables2D<N1D> *>(sm);
  double *nodes = sm + TBL;                       // [NFLD][S]   swizzled node positions
  double2 *partsL = reinterpret_cast<double2 *>(nodes + NFLD * S);   // [d][half][S]
  double2 *partsH = partsL + 4 * S;                                   // [d][half][S] (not MODE_SUBCELL)
  double *lamp = reinterpret_cast<double *>(partsH + ((MODE == MODE_SUBCELL) ? 0 : 4 * S));  // [2][S] / lstage
  double *lmin = lamp + EPB * fast_lamp_per_elem<N1D>();   // [EPB][N1D]
  int *needlog = reinterpret_cast<int *>(lmin + EPB * N1D);   // [EPB] (LAZY_LOGS): some pair of the element may leave logmean's series branch
  static_assert(KIND == KIND_RT || MODE == MODE_SUBCELL, "stage kinds exist for the subcell limiter's direct schedule only");
  const bool nst1 = KIND == KIND_S1 || (KIND == KIND_RT && A.nstage == 1);      // CFL reduction in this launch
  const bool fuse = KIND == KIND_S23 || (KIND == KIND_RT && MODE == MODE_SUBCELL && A.fuse != 0);   // SSP combine fused into the output phase
  constexpr bool DIAG = KIND == KIND_RT;                                         // rhsL / rhsH diagnostics possible

  const int tid = threadIdx.x;
  const int d = tid / HALF, rr = tid % HALF, el = rr / N1D, line = rr % N1D;
  const long long k = kb + el;                            // kb = first element of this CTA's batch
  const bool active = INTERIOR || k < M.K;
  const bool full = INTERIOR || kb + EPB <= M.K;          // no partial batch: skip the per-element guards
  const double gamma = A.gamma, gm1 = A.gamma - 1.0;
  const double *Ubase = A.Uq + kb * (Nq * 4);             // this batch's states; 32-bit offsets from here on
  if (A.dbg && tid == 0) {   // p2de_debug_counters: which instantiation this CTA runs
    atomicAdd(A.dbg + (INTERIOR ? DBG_CTA_INTERIOR : DBG_CTA_GENERAL), 1ull);
    if (DEFER) atomicAdd(A.dbg + DBG_CTA_DEFER, 1ull);
  }

  if (P2DE_FAST_PREFETCH) {
    // pull the states of the batch one full wave of resident CTAs ahead into L2: that batch starts on some SM about
    // when this one ends, and its first instruction is a wait on exactly these lines (measured: -2.7 % on S-DMR;
    // a persistent-CTA loop over batches was measured too and loses 13 % to the spills of its loop-carried state)
    constexpr int AHEAD = 148 * (N1D == 5 ? P2DE_FAST_MIN_BLOCKS5 : P2DE_FAST_MIN_BLOCKS);
    const long long kp = kb + (long long)AHEAD * EPB;
    constexpr int LINES = EPB * Nq * 32 / 128;
    if (kp + EPB <= M.K) {
      if (tid < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.Uq + kp * (Nq * 4) + tid * 16));
      else if (DEFER && tid < 2 * LINES)      // (resW is Uq itself in that stage)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.defer_add + kp * (Nq * 4) + (tid - LINES) * 16));
      else if (fuse && tid < 2 * LINES)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.fuse_resW + kp * (Nq * 4) + (tid - LINES) * 16));
    }
  }
  // tables: coalesced copy from global memory (an indexed read of the kernel parameter would be a lane-serialised
  // constant-bank access).  Only the loads are issued here; the stores to shared memory come after the state loads
  // below have been issued too, so that the two round trips overlap instead of following each other.
  static_assert(Tables2D<N1D>::FAST_BYTES + 8 <= (int)sizeof(Tables2D<N1D>), "the copy is rounded up to 16-byte words");
  constexpr int TF2 = (Tables2D<N1D>::FAST_BYTES + 15) / 16, NTL = (TF2 + NT - 1) / NT;
  double2 treg[NTL];
#pragma unroll
  for (int it = 0; it < NTL; ++it) {
    const int i = tid + it * NT;
    if (i < TF2) treg[it] = reinterpret_cast<const double2 *>(A.tab_dev)[i];
  }
  const double dtl = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;   // dt the limiter sees (rhs.jl:46,52)
  // ---- the two neighbour face nodes of this line: issue the loads now (one 32-byte node each, two
  //      16-byte loads) so that their latency is covered by the node phase
  Nbr nb[2];
  Cons2 UnbC[2];
  int nbpos[2] = {-1, -1};
  {
    // Structured mesh, batch strictly inside the domain (the common case, CTA-uniform): the neighbours
    // are k -+ 1 / k -+ Kx, no boundary condition, and the partner face node follows from the LGL face
    // map that p2de_create verified (build_tables: fq2q).
    if (INTERIOR) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        nb[e].bc = 0; nb[e].ival = nullptr; nb[e].kP = 0; nb[e].fP = 0;
        // partner node: d=0: (N1D-1, line) of k-1 / (0, line) of k+1;  d=1: (line, N1D-1) of k-Kx / (line, 0) of k+Kx
        const int node = d == 0 ? (e ? 0 : N1D - 1) + line * N1D : line + (e ? 0 : N1D - 1) * N1D;
        const int dk = d == 0 ? (e ? 1 : -1) : (e ? M.Kx : -M.Kx);
        // an x-neighbour inside this batch is read from shared memory after the node phase (nbpos >= 0)
        const bool in_batch = d == 0 && (e ? el + 1 < EPB : el > 0);
        nbpos[e] = -1;
        if (in_batch) nbpos[e] = 0;
        else if (DEFER)
          UnbC[e] = load_cons_plus(Ubase + ((long long)(el + dk) * Nq + node) * 4, A.defer_add + kb * (Nq * 4) + ((long long)(el + dk) * Nq + node) * 4, dtl);
        else UnbC[e] = load_cons(Ubase + ((long long)(el + dk) * Nq + node) * 4);
      }
    } else if (active) {
      int ix, iy;
      if (M.K < 0x7fffffffll) { iy = (int)((unsigned)k / (unsigned)M.Kx); ix = (int)((unsigned)k - (unsigned)iy * (unsigned)M.Kx); }
      else { ix = (int)(k % M.Kx); iy = (int)(k / M.Kx); }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        nb[e] = neighbor<N1D>(M, k, ix, iy, (2 * d + e) * N1D + line);
        const long long noff = (nb[e].kP * Nq + Tc.fq2q[nb[e].fP]) * 4;
        UnbC[e] = (DEFER) ? load_cons_plus(A.Uq + noff, A.defer_add + noff, dtl) : load_cons(A.Uq + noff);
      }
    }
  }
  // ---- node phase (every volume node once): primitives, logs, axis wavespeeds
  constexpr bool LAZY_LOGS = DO_HIGH && Nq == 16 && NT % 32 == 0 && S % NT == 0;
  constexpr int NITN = (S + NT - 1) / NT;
  Cons2 Uraw[NITN];
#pragma unroll
  for (int it = 0; it < NITN; ++it) {   // all of this thread's loads are in flight before anything waits
    const int n = tid + it * NT;
    Uraw[it].rho = 1.0; Uraw[it].m1 = 0.0; Uraw[it].m2 = 0.0; Uraw[it].E = 1.0;
    if (n < S && (full || kb + n / Nq < M.K)) {
      Uraw[it] = (DEFER) ? load_cons_plus(Ubase + n * 4, A.defer_add + kb * (Nq * 4) + n * 4, dtl)
                                                       : load_cons(Ubase + n * 4);
      if (fuse && !DEFER)   // the flat output phase of this same thread reads resW here: pull it into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.fuse_resW + kb * (Nq * 4) + n * 4));
    }
  }
#pragma unroll
  for (int it = 0; it < NTL; ++it) {
    const int i = tid + it * NT;
    if (i < TF2) reinterpret_cast<double2 *>(sm)[i] = treg[it];
  }
  __syncthreads();   // tables are in shared memory
#pragma unroll
  for (int it = 0; it < NITN; ++it) {
    const int n = tid + it * NT;
    const int e2 = n / Nq, node = n % Nq;
    const bool valid = n < S && (full || kb + e2 < M.K);
    if (valid || (LAZY_LOGS && n < S)) {
      const Cons2 U = Uraw[it];
      double *o = nodes + node_pos<N1D>(e2, node % N1D, node / N1D);
      double rinv = rcp_fast(U.rho);
      double p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
      double beta = 0.5 * U.rho * rcp_fast(p);
      o[0 * S] = U.rho; o[1 * S] = U.m1; o[2 * S] = U.m2; o[3 * S] = U.E;
      o[4 * S] = U.m1 * rinv; o[5 * S] = U.m2 * rinv; o[6 * S] = p; o[7 * S] = beta;
      if (DO_HIGH) {
        // log(rho), log(beta) are only read by the non-series branch of logmean (:307-321), i.e. by a
        // node pair with |da| >= 1e-4 |aavg|.  If rho and beta each vary by less than 6.2e-5
        // relative over the whole element (compared on the high words of the doubles: 65 units of
        // 2^-20 at most), no pair of this element can take that branch and the logs are never looked at.
        bool need = true;
        if (LAZY_LOGS) {
          // every node within 32 units of the element's first node => spread <= 64 units
          const int hr = __double2hiint(U.rho), hb = __double2hiint(beta);
          const int lead = tid & 16;                     // first of the 16 lanes holding this element's nodes
          const int er = hr - __shfl_sync(0xffffffffu, hr, lead), eb = hb - __shfl_sync(0xffffffffu, hb, lead);
          const bool far = (unsigned)(er + 32) > 64u || (unsigned)(eb + 32) > 64u;
          need = ((__ballot_sync(0xffffffffu, far) >> lead) & 0xFFFFu) != 0u;
          if (P2DE_FAST_QUIET_PAIRS && (tid & 15) == 0) needlog[e2] = need ? 1 : 0;
          if (A.dbg && (tid & 15) == 0 && valid) { atomicAdd(A.dbg + DBG_ELEM, 1ull); if (need) atomicAdd(A.dbg + DBG_ELEM_LOGS, 1ull); }
        }
        if (need) { o[8 * S] = log(U.rho); o[9 * S] = log(beta); }
      }
      o[10 * S] = wavespeed_rot(gamma, gm1, rinv, U.m1, U.E);
      o[11 * S] = wavespeed_rot(gamma, gm1, rinv, U.m2, U.E);
    }
  }
  __syncthreads();

  // ---- line phase.  G = wJ * (rhsxyH - rhsxyL) along this line (SUBCELL) or the two parts.
  int pos[N1D];
  double G[N1D][4];
  double dF0[4];
#pragma unroll
  for (int a = 0; a < N1D; ++a) pos[a] = line_pos<N1D>(el, d, line, a);
  const double *rwJ = T.rwJl[d][line];   // 1 / (Jq wq) of this line's nodes
  // warp-uniform (all lanes vote, also those of a partial batch): no element of this warp can leave the series branch of
  // logmean (constant and smooth regions), so its pairs take fS_rot_quiet
  bool quiet = false;
  if (P2DE_FAST_QUIET_PAIRS && LAZY_LOGS) quiet = __all_sync(0xffffffffu, needlog[el] == 0);
  if (active) {
    ConsR Unb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (nbpos[e] >= 0) {   // x-neighbour of this batch: its state is in `nodes`
        const int eln = e ? el + 1 : el - 1;
        const double *o = nodes + node_pos<N1D>(eln, e ? 0 : N1D - 1, line);
        UnbC[e].rho = o[0 * S]; UnbC[e].m1 = o[1 * S]; UnbC[e].m2 = o[2 * S]; UnbC[e].E = o[3 * S];
      }
      Unb[e].rho = UnbC[e].rho; Unb[e].mn = d ? UnbC[e].m2 : UnbC[e].m1;
      Unb[e].mt = d ? UnbC[e].m1 : UnbC[e].m2; Unb[e].E = UnbC[e].E;
    }
    double lamPair[N1D], lamFace[2];
    {
      // ---- low-order graph-viscosity terms + both surface fluxes (low_order_graph_viscosity.jl:139-204,
      //      flux_differencing.jl:90-151,223-272)
      ConsR U[N1D];
      double fl[N1D][4], ws[N1D], GL[N1D][4];
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double *o = nodes + pos[a];
        U[a].rho = o[0 * S]; U[a].mn = o[(1 + d) * S]; U[a].mt = o[(2 - d) * S]; U[a].E = o[3 * S];
        flux_rot(U[a], o[(4 + d) * S], o[(5 - d) * S], o[6 * S], fl[a]);
        ws[a] = o[(10 + d) * S];
#pragma unroll
        for (int c = 0; c < 4; ++c) GL[a][c] = 0.0;
      }
      if (DO_LOW) {
#pragma unroll
        for (int a = 0; a < N1D - 1; ++a) {
          const int i = a + 1, j = a;
          double Sv = T.S0[d][line][a];
          double lam = fabs(Sv) * fmax(ws[i], ws[j]);
          lamPair[a] = lam;
          const double ui[4] = {U[i].rho, U[i].mn, U[i].mt, U[i].E}, uj[4] = {U[j].rho, U[j].mn, U[j].mt, U[j].E};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double SF = Sv * (fl[i][c] + fl[j][c]) - lam * (uj[c] - ui[c]);   // 2 Sv (f_i + f_j)/2: the scalings by 2 are exact
            GL[i][c] -= SF; GL[j][c] += SF;          // GL = -Q0F1
          }
        }
        lamPair[N1D - 1] = 0.0;
      }
      double BFH[2][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) dF0[c] = 0.0;
      if (MODE == MODE_SUBCELL) {
        // G = wJ (rhsxyH - rhsxyL) starts as Q0F1 - (BF_H - BF_L): the volume part of -GL; the surface terms
        // cancel identically on interior faces (identity projection) and leave the LF term on inflow/outflow faces
#pragma unroll
        for (int a = 0; a < N1D; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) G[a][c] = -GL[a][c];
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ae = e ? N1D - 1 : 0;
        double B = T.Bf[d][line][e], nn = fabs(B);
        double rinvP = rcp_fast(Unb[e].rho);
        double wsP = wavespeed_rot(gamma, gm1, rinvP, Unb[e].mn, Unb[e].E);
        double lamB = 0.5 * nn * fmax(ws[ae], wsP);
        ConsR uP = Unb[e];
        const int bce = INTERIOR ? 0 : nb[e].bc;
        if (bce) {
          if (bce == 1) { const double *p = nb[e].ival; uP.rho = p[0]; uP.mn = p[1 + d]; uP.mt = p[2 - d]; uP.E = p[3]; }
          else uP = U[ae];
          rinvP = rcp_fast(uP.rho);
        }
        double fP[4];
        flux_rot(uP, uP.mn * rinvP, uP.mt * rinvP, gm1 * (uP.E - 0.5 * (uP.mn * uP.mn + uP.mt * uP.mt) * rinvP), fP);
        const double up[4] = {uP.rho, uP.mn, uP.mt, uP.E}, uf[4] = {U[ae].rho, U[ae].mn, U[ae].mt, U[ae].E};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double bfs = B * (0.5 * (fl[ae][c] + fP[c]));
          double lf = lamB * (up[c] - uf[c]);
          GL[ae][c] -= bfs - lf;                         // - BF_L
          if (MODE == MODE_SUBCELL) { if (bce) G[ae][c] -= lf; }
          else BFH[e][c] = bce ? bfs : bfs - lf;         // BF_H (LFc = 0 on inflow/outflow faces)
          if (e == 0 && bce) dF0[c] = lf;                // BF_H - BF_L on the seed face
        }
        lamFace[e] = lamB;
      }
      // publish rhsxyL_d = GL / wJ (scale_low_order_rhs_by_mass! :206-220) and the CFL lambdas
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        if (DO_LOW) {
          partsL[(d * 2 + 0) * S + pos[a]] = make_double2(GL[a][0] * rwJ[a], GL[a][1] * rwJ[a]);
          partsL[(d * 2 + 1) * S + pos[a]] = make_double2(GL[a][2] * rwJ[a], GL[a][3] * rwJ[a]);
          if (nst1)   // this direction's share of lambda_i (:222-281): its two volume pairs and its face
            lamp[d * S + pos[a]] = ((a > 0 ? lamPair[a - 1] : 0.0) + lamPair[a]) + (a == 0 ? lamFace[0] : (a == N1D - 1 ? lamFace[1] : 0.0));
        }
        // the other modes keep wJ rhsxyH by itself: G starts as -BF_H, the volume pairs below add the rest
        if (MODE != MODE_SUBCELL) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double bh = a == 0 ? BFH[0][c] : 0.0;
            if (a == N1D - 1) bh += BFH[1][c];
            G[a][c] = -bh;
          }
        }
      }
    }
    if (DO_HIGH) {
      // ---- flux differencing along this line, flux_differencing.jl:164-211 (pairs j<i, j outer)
      PrimR q[N1D];
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double *o = nodes + pos[a];
        q[a].rho = o[0 * S]; q[a].un = o[(4 + d) * S]; q[a].ut = o[(5 - d) * S];
        q[a].beta = o[7 * S];
      }
      if (quiet) {
        PairLoop<N1D, 0, 1>::run([&](auto jc, auto ic) {
          constexpr int j = decltype(jc)::value, i = decltype(ic)::value;
          double F[4];
          fS_rot_quiet(A.half_inv_gm1, q[i], q[j], F);
          double Sv = T.SHt[d][i][j][line];
#pragma unroll
          for (int c = 0; c < 4; ++c) { double Sf = Sv * F[c]; G[i][c] -= Sf; G[j][c] += Sf; }
        });
      } else {
#pragma unroll
        for (int a = 0; a < N1D; ++a) {
          const double *o = nodes + pos[a];
          q[a].rholog = o[8 * S]; q[a].betalog = o[9 * S];
        }
        PairLoop<N1D, 0, 1>::run([&](auto jc, auto ic) {
          constexpr int j = decltype(jc)::value, i = decltype(ic)::value;
          double F[4];
          fS_rot(A.half_inv_gm1, q[i], q[j], F);
          double Sv = T.SHt[d][i][j][line];
#pragma unroll
          for (int c = 0; c < 4; ++c) { double Sf = Sv * F[c]; G[i][c] -= Sf; G[j][c] += Sf; }
        });
      }
    }
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      if (MODE != MODE_SUBCELL && DO_HIGH) {
        partsH[(d * 2 + 0) * S + pos[a]] = make_double2(G[a][0] * rwJ[a], G[a][1] * rwJ[a]);
        partsH[(d * 2 + 1) * S + pos[a]] = make_double2(G[a][2] * rwJ[a], G[a][3] * rwJ[a]);
      }
    }
  }
  __syncthreads();

  // ---- CFL: dt = min_i CFL * 0.5 * wJ_i / lambda_i, low_order_graph_viscosity.jl:222-281
  if (DO_LOW && nst1) {
    double dtloc = INFINITY;
    if (active && d == 0) {
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double li = lamp[0 * S + pos[a]] + lamp[1 * S + pos[a]];
        dtloc = jl_min(dtloc, A.CFL * 0.5 * (A.Jq * T.wq[a + line * N1D]) / li);
      }
    }
    {
      const unsigned wmask = __activemask();   // the CTA's last warp may be partial
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        double other = __shfl_xor_sync(wmask, dtloc, off);
        if ((wmask >> ((tid & 31) ^ off)) & 1u) dtloc = jl_min(dtloc, other);
      }
    }
    if ((tid & 31) == 0) dt_publish(A.dt_bits, dtloc);
  }

  if (!active && MODE != MODE_ZHANGSHU && MODE != MODE_SUBCELL) return;

  if (MODE == MODE_SUBCELL) {
    // shared-memory staging of this kernel's outputs (regions that are dead by now: node fields
    // 4..11 are only read before the barrier above, lamp only by the CFL block)
    double2 *tbuf = reinterpret_cast<double2 *>(nodes + 4 * S);   // [d][half][S]
    double *lstage = lamp;                                        // [EPB][2*N1D*NF], L_local layout
    if (nst1) __syncthreads();   // CFL block done with lamp
    if (active) {
    // ---- f_bar_H - f_bar_L by prefix sum (subcell.jl:163-206) and the limiting coefficients of
    //      this line's N1D+1 subcell faces (subcell.jl:248-349)
    double dFv[NF][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) dFv[0][c] = INTERIOR ? 0.0 : dF0[c];
#pragma unroll
    for (int s = 1; s < NF; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) dFv[s][c] = (INTERIOR && s == 1) ? G[0][c] : dFv[s - 1][c] + G[s - 1][c];
    // End faces.  f_bar_H - f_bar_L on an element face is BF_H - BF_L there, which with the identity projection is
    // exactly zero unless the face carries an inflow/outflow condition (the seed dF0 above; at the far end the prefix
    // sum returns to it up to rounding, ~1e-16 |flux|).  The exact zero is used: the face's P is then zero, its
    // coefficient is 1 from both sides without evaluating limiting_param, and the interface symmetrisation
    // (subcell.jl:418-456) is the identity (boundary faces are their own partners), so no second kernel is needed.
    const bool bc0 = !INTERIOR && nb[0].bc != 0, bc1 = !INTERIOR && nb[1].bc != 0;
    if (!bc1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) dFv[N1D][c] = 0.0;
    }
    // node by node: u^L = Uq + dt rhsL (subcell.jl:269), its bounds, and the two subcell faces
    // next to it (P = -/+ 4 dt (fH - fL) / wJ, subcell.jl:300,312,328,340)
    double lv[NF];
#pragma unroll
    for (int s = 0; s < NF; ++s) lv[s] = 1.0;
    // First pass, branch-free and division-free: is EVERY coefficient of this line certainly 1?  With u' = u^L + P,
    //   rho(u') >= zeta rho(u^L)                  <=>  rho' - zeta rho_L >= 0
    //   rho e(u') >= zeta rho e(u^L)              <=>  rho' (2 rho_L E' - zeta c_L) - |m'|^2 rho_L >= 0,  c_L = 2 rho_L E_L - |m_L|^2
    // (the second line is 2 rho_L q(1) of the reference's quadratic q(l) = a l^2 + b l + c, limiter_utils.jl:85-90).
    // rho e is concave in the conserved variables, so q > 0 on all of [0, 1] when q(0) = c > 0 and q(1) > 0: no root in
    // (0, 1], and limiting_param_bound_rho_rhoe returns min(., 1) = 1.  Both tests carry a margin of 1e-9 relative to the
    // size of their terms (rounding moves either evaluation by ~1e-16), so a line that passes has all coefficients 1 in
    // the reference's evaluation as well; everything else -- a sliver of |q(1)| < 1e-9 |terms| around the limiter's
    // activation and the genuinely limited faces -- goes to the exact evaluation below.
    bool all_easy = true;
    constexpr bool SURE = P2DE_FAST_SURE_LIMITER != 0;
    static_assert(P2DE_FAST_LIMITER_TWO_PASS, "the single-pass limiter was removed");
    {
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const double *o = nodes + pos[a];
        double2 m0 = partsL[(d * 2 + 0) * S + pos[a]], m1 = partsL[(d * 2 + 1) * S + pos[a]];
        double2 o0 = partsL[((1 - d) * 2 + 0) * S + pos[a]], o1 = partsL[((1 - d) * 2 + 1) * S + pos[a]];
        // the other direction's share is in ITS rotated frame: momentum components swap
        double r0 = m0.x + o0.x, r1 = m0.y + o1.x, r2 = m1.x + o0.y, r3 = m1.y + o1.y;
        Cons2 uL;   // u^L = Uq + dt rhsL (subcell.jl:269)
        uL.rho = o[0 * S] + dtl * r0; uL.m1 = o[(1 + d) * S] + dtl * r1;
        uL.m2 = o[(2 - d) * S] + dtl * r2; uL.E = o[3 * S] + dtl * r3;
        const double kk = 4 * dtl * rwJ[a];     // P = -/+ 4 dt (fH - fL) / wJ, subcell.jl:300,312,328,340
        if (SURE) {
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
        } else {
          const double rhoeL = uL.E - 0.5 * (uL.m1 * uL.m1 + uL.m2 * uL.m2) * rcp_fast(uL.rho);
          const double Lrho = A.zeta * uL.rho, Lrhoe = A.zeta * rhoeL;
          const double c0 = (1.0 - A.zeta) * uL.rho * rhoeL;
          double Pm[4], Pp[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) { Pm[c] = -kk * dFv[a][c]; Pp[c] = kk * dFv[a + 1][c]; }
          if (a > 0 || bc0) {
            double qa, qb;
            quad_coeff_ab(uL, Pm, Lrhoe, qa, qb);
            all_easy = all_easy & limiting_param_pos_easy(uL.rho, Pm[0], Lrho, qa, qb, c0);
          }
          if (a < N1D - 1 || bc1) {
            double qa, qb;
            quad_coeff_ab(uL, Pp, Lrhoe, qa, qb);
            all_easy = all_easy & limiting_param_pos_easy(uL.rho, Pp[0], Lrho, qa, qb, c0);
          }
        }
        if (DIAG) {
          if (d == 0 && A.rhsL_diag) {
            const int node = a + line * N1D;
            double r[4] = {r0, r1, r2, r3};
            store4(A.rhsL_diag + (k * Nq + node) * 4, r);
          }
          if (A.rhsH_diag) {   // diagnostics: rhsxyH_d = rhsxyL_d + G / wJ; each line adds its share (buffer pre-zeroed)
            const int node = d == 0 ? a + line * N1D : line + a * N1D;
            double *hd = A.rhsH_diag + (k * Nq + node) * 4;
            atomicAdd(hd + 0, m0.x + G[a][0] * rwJ[a]); atomicAdd(hd + 1 + d, m0.y + G[a][1] * rwJ[a]);
            atomicAdd(hd + 2 - d, m1.x + G[a][2] * rwJ[a]); atomicAdd(hd + 3, m1.y + G[a][3] * rwJ[a]);
          }
        }
      }
    }
    if (A.dbg) { atomicAdd(A.dbg + DBG_LINES, 1ull); if (!all_easy) atomicAdd(A.dbg + DBG_LINES_NOT_EASY, 1ull); }
    if (!all_easy) {
      // exact evaluation (the reference's formulas: quadratic coefficients, root selection), node by node
      bool one = true;
#pragma unroll 1
      for (int a = 0; a < N1D; ++a) {
        const int pa = line_pos<N1D>(el, d, line, a);   // (= pos[a]; a is a run-time index in this rolled loop)
        const double *o = nodes + pa;
        double2 m0 = partsL[(d * 2 + 0) * S + pa], m1 = partsL[(d * 2 + 1) * S + pa];
        double2 o0 = partsL[((1 - d) * 2 + 0) * S + pa], o1 = partsL[((1 - d) * 2 + 1) * S + pa];
        double r0 = m0.x + o0.x, r1 = m0.y + o1.x, r2 = m1.x + o0.y, r3 = m1.y + o1.y;
        Cons2 uL;
        uL.rho = o[0 * S] + dtl * r0; uL.m1 = o[(1 + d) * S] + dtl * r1;
        uL.m2 = o[(2 - d) * S] + dtl * r2; uL.E = o[3 * S] + dtl * r3;
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
    }   // !all_easy
    // (update_blending_factor! = 1 without shock capturing, shock_capture.jl:111-114: the FAST path has none, so the
    //  min with A.blend is the identity)
    // this line's share of the un-symmetrised limited rhs (subcell.jl:841-924 with the line's own
    // coefficients): t_d = rhsxyL_d + (l_{a+1} dF_{a+1} - l_a dF_a) / wJ, in the line's rotated frame
    if (all_easy) {
      // every coefficient of the line is 1: l_{a+1} dF_{a+1} - l_a dF_a is the prefix sum's own increment G[a],
      // i.e. the share is rhsxyH_d
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        double2 m0 = partsL[(d * 2 + 0) * S + pos[a]], m1 = partsL[(d * 2 + 1) * S + pos[a]];
        tbuf[(d * 2 + 0) * S + pos[a]] = make_double2(m0.x + G[a][0] * rwJ[a], m0.y + G[a][1] * rwJ[a]);
        tbuf[(d * 2 + 1) * S + pos[a]] = make_double2(m1.x + G[a][2] * rwJ[a], m1.y + G[a][3] * rwJ[a]);
      }
    } else {
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      double2 m0 = partsL[(d * 2 + 0) * S + pos[a]], m1 = partsL[(d * 2 + 1) * S + pos[a]];
      // (INTERIOR: dF is exactly zero on the two end faces, the products are dropped at compile time)
      double hi[4], lo[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        hi[c] = (INTERIOR && a == N1D - 1) ? 0.0 : lv[a + 1] * dFv[a + 1][c];
        lo[c] = (INTERIOR && a == 0) ? 0.0 : lv[a] * dFv[a][c];
      }
      tbuf[(d * 2 + 0) * S + pos[a]] = make_double2(m0.x + (hi[0] - lo[0]) * rwJ[a], m0.y + (hi[1] - lo[1]) * rwJ[a]);
      tbuf[(d * 2 + 1) * S + pos[a]] = make_double2(m1.x + (hi[2] - lo[2]) * rwJ[a], m1.y + (hi[3] - lo[3]) * rwJ[a]);
    }
    }   // !all_easy
#pragma unroll
    for (int s = 0; s < NF; ++s) lstage[el * (2 * N1D * NF) + d * (N1D * NF) + (d == 0 ? s + line * NF : line + s * N1D)] = lv[s];
    }   // active
    // ---- flat, coalesced output phase: rpre = x share + y share (y share un-rotated), lpre
    constexpr int NIT = (S + NT - 1) / NT;
    constexpr int NL = 2 * N1D * NF;
    double2 wres[NIT][2];
    {
      const double *rw = A.fuse_resW + kb * (Nq * 4);
#pragma unroll
      for (int it = 0; it < NIT; ++it) {   // resW of this thread's nodes: in flight across the barrier
        const int n = tid + it * NT;
        wres[it][0] = make_double2(0.0, 0.0); wres[it][1] = make_double2(0.0, 0.0);
        if (fuse && n < S && (full || kb + n / Nq < M.K)) {
          const double2 *q = reinterpret_cast<const double2 *>(rw + n * 4);
          wres[it][0] = q[0]; wres[it][1] = q[1];
        }
      }
    }
    __syncthreads();
    double *out = A.rpre + kb * (Nq * 4);
    if (NIT > 2) {   // N1D = 5: node by node (three nodes' worth of staging registers spills)
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int n = tid + it * NT;
        const int e2 = n / Nq, node = n % Nq;
        if (n < S && (full || kb + e2 < M.K)) {
          const int p2 = node_pos<N1D>(e2, node % N1D, node / N1D);
          double2 x0 = tbuf[0 * S + p2], x1 = tbuf[1 * S + p2], y0 = tbuf[2 * S + p2], y1 = tbuf[3 * S + p2];
          double r[4] = {x0.x + y0.x, x0.y + y1.x, x1.x + y0.y, x1.y + y1.y};
          if (fuse) {   // stages 2, 3: dt is known, so the SSP combine (SSPRK33.jl:34-39) of the un-corrected rhs is done here
            r[0] = A.fuse_a * wres[it][0].x + A.fuse_b * (nodes[0 * S + p2] + dtl * r[0]);
            r[1] = A.fuse_a * wres[it][0].y + A.fuse_b * (nodes[1 * S + p2] + dtl * r[1]);
            r[2] = A.fuse_a * wres[it][1].x + A.fuse_b * (nodes[2 * S + p2] + dtl * r[2]);
            r[3] = A.fuse_a * wres[it][1].y + A.fuse_b * (nodes[3 * S + p2] + dtl * r[3]);
          }
          store4(out + n * 4, r);
        }
      }
    } else {
    // all shared-memory loads of this thread's NIT nodes first, then the arithmetic and the stores (the loads of
    // one node used to wait behind the stores of the previous one)
    int p2v[NIT];
    bool okv[NIT];
    double2 xv[NIT][4];
    double uo[NIT][4];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int n = tid + it * NT;
      const int e2 = n / Nq, node = n % Nq;
      okv[it] = n < S && (full || kb + e2 < M.K);
      // swizzled position: arithmetic for N1D = 4 (no dependent table look-up), table otherwise
      p2v[it] = n < S ? node_pos<N1D>(e2, node % N1D, node / N1D) : 0;
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it)
      if (okv[it]) {
        const int p2 = p2v[it];
        xv[it][0] = tbuf[0 * S + p2]; xv[it][1] = tbuf[1 * S + p2]; xv[it][2] = tbuf[2 * S + p2]; xv[it][3] = tbuf[3 * S + p2];
        if (fuse) {
#pragma unroll
          for (int c = 0; c < 4; ++c) uo[it][c] = nodes[c * S + p2];
        }
      }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int n = tid + it * NT;
      if (okv[it]) {
        const double2 x0 = xv[it][0], x1 = xv[it][1], y0 = xv[it][2], y1 = xv[it][3];
        double r[4] = {x0.x + y0.x, x0.y + y1.x, x1.x + y0.y, x1.y + y1.y};
        if (fuse) {   // stages 2, 3: dt is known, so the SSP combine (SSPRK33.jl:34-39) of the un-corrected rhs is done here
          r[0] = A.fuse_a * wres[it][0].x + A.fuse_b * (uo[it][0] + dtl * r[0]);
          r[1] = A.fuse_a * wres[it][0].y + A.fuse_b * (uo[it][1] + dtl * r[1]);
          r[2] = A.fuse_a * wres[it][1].x + A.fuse_b * (uo[it][2] + dtl * r[2]);
          r[3] = A.fuse_a * wres[it][1].y + A.fuse_b * (uo[it][3] + dtl * r[3]);
        }
        store4(out + n * 4, r);
      }
    }
    }   // NIT <= 2
    double *lout = A.lpre + kb * NL;
    if (full && (EPB * NL) % 2 == 0) {   // 16-byte copies
      const double2 *ls2 = reinterpret_cast<const double2 *>(lstage);
      double2 *lo2 = reinterpret_cast<double2 *>(lout);
      for (int n = tid; n < EPB * NL / 2; n += NT) lo2[n] = ls2[n];
    } else {
      for (int n = tid; n < EPB * NL; n += NT)
        if (kb + n / NL < M.K) lout[n] = lstage[n];
    }
    return;
  }

  // ---- element-local limiters / no limiter: x-line threads produce rhsU
  double rL[N1D][4], rH[N1D][4];
  double lline = 1.0;
  if (active && d == 0) {
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      const double *o = nodes + pos[a];
#pragma unroll
      for (int c = 0; c < 4; ++c) { rL[a][c] = 0.0; rH[a][c] = 0.0; }
      if (DO_LOW) {
        double2 m0 = partsL[0 * S + pos[a]], m1 = partsL[1 * S + pos[a]], o0 = partsL[2 * S + pos[a]], o1 = partsL[3 * S + pos[a]];
        rL[a][0] = m0.x + o0.x; rL[a][1] = m0.y + o1.x; rL[a][2] = m1.x + o0.y; rL[a][3] = m1.y + o1.y;
      }
      if (DO_HIGH) {
        double2 m0 = partsH[0 * S + pos[a]], m1 = partsH[1 * S + pos[a]], o0 = partsH[2 * S + pos[a]], o1 = partsH[3 * S + pos[a]];
        rH[a][0] = m0.x + o0.x; rH[a][1] = m0.y + o1.x; rH[a][2] = m1.x + o0.y; rH[a][3] = m1.y + o1.y;
      }
      if (MODE == MODE_ZHANGSHU) {   // zhangshu.jl:4-45
        Cons2 uL;
        uL.rho = o[0 * S] + dtl * rL[a][0]; uL.m1 = o[1 * S] + dtl * rL[a][1];
        uL.m2 = o[2 * S] + dtl * rL[a][2]; uL.E = o[3 * S] + dtl * rL[a][3];
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
    if (d == 0 && line == 0) A.Lout[k] = l;
    l = jl_min(l, A.blend);
  }
  if (d == 0) {
#pragma unroll
    for (int a = 0; a < N1D; ++a) {
      const int node = a + line * N1D;
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

// first element of this CTA's batch and whether the batch lies strictly inside a structured mesh
template <int EPB>
P2DE_DEV long long fast_batch(const StageArgs &A, const MeshTopo &M, bool &interior) {
  long long kb;
  interior = false;
  if (A.rowblocks) {   // structured mesh with Kx a multiple of EPB: 2D grid (batch in row, element row), no division
    kb = ((long long)blockIdx.y * A.rowblocks + blockIdx.x) * EPB;
    interior = blockIdx.x > 0u && blockIdx.x + 1u < (unsigned)A.rowblocks && blockIdx.y > 0u && blockIdx.y + 1u < gridDim.y;
  } else {
    kb = (long long)blockIdx.x * EPB;
    if (!M.mapP32 && kb + EPB <= M.K && M.K < 0x7fffffffll) {
      const unsigned iy0 = (unsigned)kb / (unsigned)M.Kx, ix0 = (unsigned)kb - iy0 * (unsigned)M.Kx;
      interior = ix0 > 0u && ix0 + EPB < (unsigned)M.Kx && iy0 > 0u && iy0 + 1u < (unsigned)M.Ky;
    }
  }
  return kb;
}

template <int N1D, int MODE, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D, (N1D == 5 ? P2DE_FAST_MIN_BLOCKS5 : P2DE_FAST_MIN_BLOCKS))
stage_kernel_fast(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M,
                  const __grid_constant__ Tables2D<N1D> Tc) {
  bool interior;
  const long long kb = fast_batch<EPB>(A, M, interior);
  if (interior) stage_fast_impl<N1D, MODE, EPB, true, false>(A, M, Tc, kb);
  else stage_fast_impl<N1D, MODE, EPB, false, false>(A, M, Tc, kb);
}

// The three kernels of the direct schedule (p2de_ssp33_step, subcell limiter): the stage's role is a compile-time fact
// (KIND), so the CFL block, the fused combine and the diagnostics are present or absent without run-time flags.  Each is
// its own __global__ function: as further copies of the body inside one kernel they made ptxas' register allocation for
// the other copies worse (measured, 3 %).
template <int N1D, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D, (N1D == 5 ? P2DE_FAST_MIN_BLOCKS5 : P2DE_FAST_MIN_BLOCKS))
stage_kernel_fast_s1(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M,
                     const __grid_constant__ Tables2D<N1D> Tc) {
  bool interior;
  const long long kb = fast_batch<EPB>(A, M, interior);
  if (interior) stage_fast_impl<N1D, MODE_SUBCELL, EPB, true, false, KIND_S1>(A, M, Tc, kb);
  else stage_fast_impl<N1D, MODE_SUBCELL, EPB, false, false, KIND_S1>(A, M, Tc, kb);
}
template <int N1D, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D, (N1D == 5 ? P2DE_FAST_MIN_BLOCKS5 : P2DE_FAST_MIN_BLOCKS))
stage_kernel_fast_s3(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M,
                     const __grid_constant__ Tables2D<N1D> Tc) {
  bool interior;
  const long long kb = fast_batch<EPB>(A, M, interior);
  if (interior) stage_fast_impl<N1D, MODE_SUBCELL, EPB, true, false, KIND_S23>(A, M, Tc, kb);
  else stage_fast_impl<N1D, MODE_SUBCELL, EPB, false, false, KIND_S23>(A, M, Tc, kb);
}
// stage 2 of the direct schedule: the stage input is Uq + dt * defer_add (the stage-1 combine formed while loading)
template <int N1D, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D, (N1D == 5 ? P2DE_FAST_MIN_BLOCKS5 : P2DE_FAST_MIN_BLOCKS))
stage_kernel_fast_defer(const __grid_constant__ StageArgs A, const __grid_constant__ MeshTopo M,
                        const __grid_constant__ Tables2D<N1D> Tc) {
  bool interior;
  const long long kb = fast_batch<EPB>(A, M, interior);
  if (interior) stage_fast_impl<N1D, MODE_SUBCELL, EPB, true, true, P2DE_FAST_KINDS ? KIND_S23 : KIND_RT>(A, M, Tc, kb);
  else stage_fast_impl<N1D, MODE_SUBCELL, EPB, false, true, P2DE_FAST_KINDS ? KIND_S23 : KIND_RT>(A, M, Tc, kb);
}

// update kernel for the FAST stage kernel's scratch (rpre, dFend, lpre): interface symmetrisation
// of the coefficients (subcell.jl:418-456) as a correction of the un-symmetrised limited rhs,
//   rhsU = rpre -+ (min(l, l_P) - l) dF_end / wJ   on the nodes of the element boundary,
// then the SSP stage combine (SSPRK33.jl:31-39).  16 threads per element: one per face node for
// the corrections, then one per volume node with fully coalesced 32-byte accesses.
template <int N1D, int EPB>
__global__ void __launch_bounds__(EPB * 16)
update_kernel_fast(const __grid_constant__ UpdateArgs A, const __grid_constant__ MeshTopo M,
                   const __grid_constant__ Tables2D<N1D> Tc) {
  constexpr int Nq = N1D * N1D, Nfp = 4 * N1D, NF = N1D + 1, NL = 2 * N1D * NF, TPE = 16;
  __shared__ double corr[EPB * Nfp * 4];
  __shared__ double s_rwJ[Nq];
  __shared__ int s_fq2q[Nfp];
  const int tid = threadIdx.x, el = tid / TPE, tl = tid % TPE;
  const long long k = (long long)blockIdx.x * EPB + el;
  const bool active = k < M.K;
  if (tid < Nq) s_rwJ[tid] = Tc.rwJ[tid];
  if (tid < Nfp) s_fq2q[tid] = Tc.fq2q[tid];
  __syncthreads();
  if (active) {
    const int ix = (int)(k % M.Kx), iy = (int)(k / M.Kx);
    for (int f = tl; f < Nfp; f += TPE) {
      const int F = f / N1D, line = f % N1D, d = F >> 1, e = F & 1;
      const int s = e ? N1D : 0;
      const int lidx = d * (N1D * NF) + (d == 0 ? s + line * NF : line + s * N1D);
      const double lv = A.lpre[k * NL + lidx];
      Nbr nb = neighbor<N1D>(M, k, ix, iy, f);
      const double lP = A.lpre[nb.kP * NL + d * (N1D * NF) + lidx_of_face<N1D>(nb.fP)];
      const double lsym = jl_min(lv, lP);
      double *c = corr + (el * Nfp + f) * 4;
      if (lsym != lv) {   // the neighbour limits this face harder than this element did
        const double w = (e ? (lsym - lv) : -(lsym - lv)) * s_rwJ[s_fq2q[f]];
        Cons2 t = load_cons(A.dFend + (k * Nfp + f) * 4);    // rotated frame of axis d
        const int dd = A.rotated ? d : 0;                    // (the generic stage kernel's dFend is not rotated)
        c[0] = w * t.rho; c[1 + dd] = w * t.m1; c[2 - dd] = w * t.m2; c[3] = w * t.E;
      } else { c[0] = 0.0; c[1] = 0.0; c[2] = 0.0; c[3] = 0.0; }   // dFend is not even read
      if (A.Llocal_out) A.Llocal_out[k * NL + lidx] = lsym;
    }
  }
  __syncthreads();
  if (!active) return;
  const double dt = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;
  if (A.Llocal_out)   // interior subcell faces: the stage kernel's coefficients are final
    for (int n = tl; n < NL; n += TPE) {
      const int dd = n / (N1D * NF), r = n % (N1D * NF);
      const int s = dd == 0 ? r % NF : r / N1D;
      if (s != 0 && s != N1D) A.Llocal_out[k * NL + n] = A.lpre[k * NL + n];
    }
  for (int node = tl; node < Nq; node += TPE) {
    const int i = node % N1D, j = node / N1D;
    const long long off = (k * Nq + node) * 4;
    Cons2 rp = load_cons(A.rpre + off);
    double r[4] = {rp.rho, rp.m1, rp.m2, rp.E};
    const double *cb = corr + el * Nfp * 4;
    if (i == 0) { const double *c = cb + (0 * N1D + j) * 4; r[0] += c[0]; r[1] += c[1]; r[2] += c[2]; r[3] += c[3]; }
    if (i == N1D - 1) { const double *c = cb + (1 * N1D + j) * 4; r[0] += c[0]; r[1] += c[1]; r[2] += c[2]; r[3] += c[3]; }
    if (j == 0) { const double *c = cb + (2 * N1D + i) * 4; r[0] += c[0]; r[1] += c[1]; r[2] += c[2]; r[3] += c[3]; }
    if (j == N1D - 1) { const double *c = cb + (3 * N1D + i) * 4; r[0] += c[0]; r[1] += c[1]; r[2] += c[2]; r[3] += c[3]; }
    if (A.pre_updated) {   // rpre already holds a*resW + b*(Uq + dt*rhs_uncorrected): add b*dt*correction
      const double *cb2 = corr + el * Nfp * 4;
      double un[4] = {rp.rho, rp.m1, rp.m2, rp.E};
      const double bd = A.b * dt;
      if (i == 0) { const double *c = cb2 + (0 * N1D + j) * 4; un[0] += bd * c[0]; un[1] += bd * c[1]; un[2] += bd * c[2]; un[3] += bd * c[3]; }
      if (i == N1D - 1) { const double *c = cb2 + (1 * N1D + j) * 4; un[0] += bd * c[0]; un[1] += bd * c[1]; un[2] += bd * c[2]; un[3] += bd * c[3]; }
      if (j == 0) { const double *c = cb2 + (2 * N1D + i) * 4; un[0] += bd * c[0]; un[1] += bd * c[1]; un[2] += bd * c[2]; un[3] += bd * c[3]; }
      if (j == N1D - 1) { const double *c = cb2 + (3 * N1D + i) * 4; un[0] += bd * c[0]; un[1] += bd * c[1]; un[2] += bd * c[2]; un[3] += bd * c[3]; }
      store4(A.Uq_out + off, un);
      continue;
    }
    if (A.rhsU_out) store4(A.rhsU_out + off, r);
    if (A.Uq_out) {
      Cons2 u = load_cons(A.Uq_in + off);
      double un[4], uo[4] = {u.rho, u.m1, u.m2, u.E};
      if (A.b == 1.0 && A.a == 0.0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) un[c] = uo[c] + dt * r[c];
      } else {
        Cons2 w = load_cons(A.resW + off);
        const double wv[4] = {w.rho, w.m1, w.m2, w.E};
#pragma unroll
        for (int c = 0; c < 4; ++c) un[c] = A.a * wv[c] + A.b * (uo[c] + dt * r[c]);
      }
      store4(A.Uq_out + off, un);
    }
  }
}

}  // namespace p2de
