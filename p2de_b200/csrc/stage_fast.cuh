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
//     boundary-condition code at all;
//   * an L2 prefetch of the batch one wave of resident CTAs ahead, table loads overlapped with the state loads,
//     a branch-free first pass of the limiter over the whole line (profiles/README.md has each step's measurement).
#pragma once
#include <type_traits>
#include "kernels2d.cuh"

namespace p2de {

#ifndef P2DE_FAST_MIN_BLOCKS
#define P2DE_FAST_MIN_BLOCKS 4
#endif
#ifndef P2DE_FAST_PREFETCH
#define P2DE_FAST_PREFETCH 1
#endif
#ifndef P2DE_FAST_MIN_BLOCKS5
#define P2DE_FAST_MIN_BLOCKS5 3   // N=4 (N1D=5): 168 registers, no spills
#endif
// A/B switches of this round's restructurings (profiles/README.md has each step's measurement)
#ifndef P2DE_FAST_QUIET_PAIRS
#define P2DE_FAST_QUIET_PAIRS 1    // warps whose elements cannot leave logmean's series branch: two-point flux without logs / selects
#endif

// (this file keeps the element-local limiters and the unlimited right-hand sides -- MODE_ZHANGSHU, MODE_LOW, MODE_HIGH -- and
//  the pointwise helpers; the subcell limiter's kernels are in stage_subcell.cuh)

// constants of logmean's series branch (:315-317) and its reciprocal: read as constant-bank operands
__constant__ double kSeries[5] = {-0.2, 0.0512, 0.026038857142857, 0.2, 0.0912};

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

// The same flux for a pair of an element in which rho and beta each vary by less than ~9 % (the "smooth" vote of the line
// phase): no logs.  With z = da / (a_L + a_R) (= f / 2 of logmean, :307-321) and t = z^2 <= 3e-3,
//   log(a_R / a_L) = 2 atanh(z)   =>   logmean = aavg / P(t),   P(t) = 1 + t/3 + t^2/5 + ... + t^6/13   (truncation < 1e-16)
// for the pairs on the reference's log branch (|f| >= 1e-4); pairs with |f| < 1e-4 keep the reference's own series
// (:315-317; it is Winters' polytropic mean, not the expansion of the log mean, so the two branches are not interchangeable).
// The series is accurate to ~4e-16, the reference's -da / (log a_L - log a_R) loses ~1e-16 / |f| to cancellation (3e-12 at
// |f| = 1e-4); the difference between the two is that noise.
P2DE_DEV double logmean_series_P(double t) {
  return fma(t, fma(t, fma(t, fma(t, fma(t, fma(t, 1.0 / 13.0, 1.0 / 11.0), 1.0 / 9.0), 1.0 / 7.0), 1.0 / 5.0), 1.0 / 3.0), 1.0);
}
P2DE_DEV void fS_rot_smooth(double half_inv_gm1, const PrimR &L, const PrimR &R, double F[4]) {
  const double sa = R.rho + L.rho, da = R.rho - L.rho, sb = R.beta + L.beta, db = R.beta - L.beta;
  const double xa = rcp_fast(sa), y = rcp_fast(sb);    // 1 / (2 aavg), 1 / (2 bavg)
  const double za = da * xa, zb = db * y, ta = za * za, tb = zb * zb;
  const bool ser = ta < 2.5e-9, serb = tb < 2.5e-9;    // |f| < 1e-4  <=>  z^2 < 2.5e-9
  // rho: aavg * (reference series in v = 4 t)  or  aavg / P(t)
  const double va = 4.0 * ta;
  const double refa = fma(va, kSeries[0] - va * (kSeries[1] - va * kSeries[2]), 1.0);
  const double rholog = (0.5 * sa) * (ser ? refa : rcp_fast(logmean_series_P(ta)));
  // 1 / logmean(beta) = (reference series  or  P(t)) / bavg
  const double vb = 4.0 * tb;
  const double refb = fma(vb, fma(vb, kSeries[4], kSeries[3]), 1.0);
  const double inv_betalog = (2.0 * y) * (serb ? refb : logmean_series_P(tb));
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
// rhsxyH shares (partsH), the CFL lambda sums [2][Nq] that are later reused as the
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
  return 12 * Nq + 8 * Nq + 8 * Nq + fast_lamp_per_elem<N1D>() + N1D + 1;   // + 1: the element's "quiet" flag
}


// INTERIOR = the batch lies strictly inside a structured mesh (CTA-uniform, decided by the kernel below): the neighbours
// are k -+ 1 / k -+ Kx and no face carries a boundary condition, so the boundary-condition branches, the seed
// f_bar_H - f_bar_L of the prefix sums and the two end-face limiter evaluations of every line vanish at compile time
// (fewer live registers: the generic version spills the boundary flags across the whole kernel).
template <int N1D, int MODE, int EPB, bool INTERIOR>
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
  double2 *partsH = partsL + 4 * S;                                   // [d][half][S]
  double *lamp = reinterpret_cast<double *>(partsH + 4 * S);          // [2][S]
  double *lmin = lamp + EPB * fast_lamp_per_elem<N1D>();   // [EPB][N1D]
  int *needlog = reinterpret_cast<int *>(lmin + EPB * N1D);   // [EPB] (LAZY_LOGS): some pair of the element may leave logmean's series branch
  static_assert(MODE != MODE_SUBCELL, "the subcell limiter has its own kernel family (stage_subcell.cuh)");
  const bool nst1 = A.nstage == 1;      // CFL reduction in this launch

  const int tid = threadIdx.x;
  const int d = tid / HALF, rr = tid % HALF, el = rr / N1D, line = rr % N1D;
  const long long k = kb + el;                            // kb = first element of this CTA's batch
  const bool active = INTERIOR || k < M.K;
  const bool full = INTERIOR || kb + EPB <= M.K;          // no partial batch: skip the per-element guards
  const double gamma = A.gamma, gm1 = A.gamma - 1.0;
  const double *Ubase = A.Uq + kb * (Nq * 4);             // this batch's states; 32-bit offsets from here on
  if (A.dbg && tid == 0) {   // p2de_debug_counters: which instantiation this CTA runs
    atomicAdd(A.dbg + (INTERIOR ? DBG_CTA_INTERIOR : DBG_CTA_GENERAL), 1ull);
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
        UnbC[e] = load_cons(A.Uq + noff);
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
      Uraw[it] = load_cons(Ubase + n * 4);
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
          double lam = fabs(Sv) * jl_max(ws[i], ws[j]);
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
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ae = e ? N1D - 1 : 0;
        double B = T.Bf[d][line][e], nn = fabs(B);
        double rinvP = rcp_fast(Unb[e].rho);
        double wsP = wavespeed_rot(gamma, gm1, rinvP, Unb[e].mn, Unb[e].E);
        double lamB = 0.5 * nn * jl_max(ws[ae], wsP);
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
          BFH[e][c] = bce ? bfs : bfs - lf;              // BF_H (LFc = 0 on inflow/outflow faces)
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
        // G = wJ rhsxyH starts as -BF_H, the volume pairs below add the rest
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double bh = a == 0 ? BFH[0][c] : 0.0;
          if (a == N1D - 1) bh += BFH[1][c];
          G[a][c] = -bh;
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
      if (DO_HIGH) {
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

  if (!active && MODE != MODE_ZHANGSHU) return;

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
    // (a launch may cover a subset of the element rows: row = row0 + blockIdx.y * row_stride, see run_step_overlapped)
    const unsigned row = (unsigned)A.row0 + blockIdx.y * (unsigned)A.row_stride;
    kb = ((long long)row * A.rowblocks + blockIdx.x) * EPB;
    // a stripe's first / last row is interior too when its neighbour row is a halo row (stored right before / after the
    // owned rows, no boundary condition on the cut)
    interior = blockIdx.x > 0u && blockIdx.x + 1u < (unsigned)A.rowblocks &&
               (row > 0u || M.ghost_lo) && (row + 1u < (unsigned)M.Ky || M.ghost_hi);
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
  if (interior) stage_fast_impl<N1D, MODE, EPB, true>(A, M, Tc, kb);
  else stage_fast_impl<N1D, MODE, EPB, false>(A, M, Tc, kb);
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
