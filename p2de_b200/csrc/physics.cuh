// physics.cuh — pointwise compressible-Euler device functions (FP64, CUDA cores).
//
// Two families:
//   * "reference-order" functions (pfun2, wavespeed_dir, flux_dir, fS_dir, entropy_roundtrip,
//     limiting_param_pos): same formulas in the same operation order as the reference's
//     src/math/compressible_Navier_Stokes.jl and src/dg/limiter/limiter_utils.jl.  Used by the
//     generic kernel variant (non-default flux options).
//   * "_fast" functions: algebraically identical, but with the divisions hoisted into per-node
//     reciprocals, the two divisions of each logmean merged into one, and the limiter's
//     quadratic solved only when a root in (0, 1] is possible.  Differences from the
//     reference-order results are a few ulp (tests hold both to 1e-12 against the CPU oracle).
// IEEE division and sqrt everywhere (no -use_fast_math); Inf/NaN behaviour of the limiter's
// root selection is the reference's.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace p2de {

#define P2DE_DEV __device__ __forceinline__

struct Cons2 { double rho, m1, m2, E; };

// Julia's min/max propagate NaN (oracle deviation D3): fmin/fmax do not, so spell it out.
// (a < b || a is NaN) ? a : b   -> NaN if either is NaN, else the smaller one
P2DE_DEV double jl_min(double a, double b) { return (a < b || a != a) ? a : b; }
P2DE_DEV double jl_max(double a, double b) { return (a > b || a != a) ? a : b; }

// CFL dt reduction target (low_order_graph_viscosity.jl:222-243): word 0 = minimum over the positive candidates (raw
// bits of positive doubles are order preserving, so atomicMin works on them); word 1 = 1.0 as long as every candidate
// was a positive number and 0.0 once some warp saw a NaN or non-positive one.  Julia's `min` propagates NaN, dt then
// becomes NaN and `while t < T` (SSPRK33.jl:28) ends; dt_read gives every consumer the same NaN.
P2DE_DEV void dt_publish(unsigned long long *dt_bits, double dtloc) {
  if (dtloc > 0.0) { if (dtloc < INFINITY) atomicMin(dt_bits, (unsigned long long)__double_as_longlong(dtloc)); }
  else atomicExch(dt_bits + 1, 0ull);
}
P2DE_DEV double dt_read(const double *dt_dev) {
  const double2 v = *reinterpret_cast<const double2 *>(dt_dev);
  return v.y == 1.0 ? v.x : __longlong_as_double(0x7ff8000000000000ll);
}

// pfun, compressible_Navier_Stokes.jl:24-28
P2DE_DEV double pfun2(double gm1, const Cons2 &U) {
  return gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) / U.rho);
}
// rhoe_ufun, :75-78
P2DE_DEV double rhoe2(const Cons2 &U) { return U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) / U.rho; }
// pfun(::Dim1), :18-22 on the normal-projected state (wavespeed_estimate(::Dim2), :58-62)
P2DE_DEV double wavespeed_n(double gamma, double gm1, double rho, double mn, double E) {
  double p = gm1 * (E - 0.5 * (mn * mn) / rho);
  return fabs(mn / rho) + sqrt(gamma * p / rho);
}
// wavespeed of U along axis d (unit normal +-e_d; the sign drops out)
P2DE_DEV double wavespeed_dir(double gamma, double gm1, const Cons2 &U, int d) {
  return wavespeed_n(gamma, gm1, U.rho, d == 0 ? U.m1 : U.m2, U.E);
}
// same with rinv = 1/rho supplied
P2DE_DEV double wavespeed_fast(double gamma, double gm1, double rinv, double mn, double E) {
  double p = gm1 * (E - 0.5 * (mn * mn) * rinv);
  return fabs(mn * rinv) + sqrt(gamma * p * rinv);
}

// flux component along axis d, fluxes(::Dim2) :175-194, given u = m1/rho, v = m2/rho, p
P2DE_DEV void flux_dir(const Cons2 &U, double u, double v, double p, int d, double f[4]) {
  double rhouv = U.rho * u * v, Ep = U.E + p;
  if (d == 0) { f[0] = U.m1; f[1] = U.m1 * u + p; f[2] = rhouv; f[3] = u * Ep; }
  else        { f[0] = U.m2; f[1] = rhouv; f[2] = U.m2 * v + p; f[3] = v * Ep; }
}
P2DE_DEV void flux_dir(double gm1, const Cons2 &U, int d, double f[4]) {
  double p = pfun2(gm1, U);
  flux_dir(U, U.m1 / U.rho, U.m2 / U.rho, p, d, f);
}

// v_ufun(::Dim2) :134-144 followed by u_vfun(::Dim2) :155-163 (entropy-projection round trip of a
// collocated LGL face node, rhs.jl:84-94 with Vf a 0/1 row).
__device__ __noinline__ Cons2 entropy_roundtrip(double gamma, double gm1, Cons2 U) {
  double p = pfun2(gm1, U);
  double s = log(p / pow(U.rho, gamma));                 // sfun :64-68
  double v1 = (gamma + 1 - s) - gm1 * U.E / p;
  double vu = U.m1 * gm1 / p, vv = U.m2 * gm1 / p, vE = -U.rho * gm1 / p;
  double q = vu * vu + vv * vv;
  double sv = gamma - v1 + q / (2 * vE);                 // s_vfun :93-97
  double rhoeV = pow(gm1 / pow(-vE, gamma), 1 / gm1) * exp(-sv / gm1);  // rhoe_vfun :99-104
  Cons2 W;
  W.rho = -rhoeV * vE; W.m1 = rhoeV * vu; W.m2 = rhoeV * vv;
  W.E = rhoeV * (1 - q / (2 * vE));
  return W;
}

// reciprocal: MUFU.RCP64H seed (>= 20 bits) + one third-order step x (1 + e + e^2), e = 1 - a x
// (error ~ e^3 < 2^-60: three dependent FMAs instead of the four of two Newton steps)
P2DE_DEV double rcp_fast(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  double t = fma(e, e, e);
  return fma(x, t, x);
}
// n / a with one residual correction (last-bit accurate for normal operands)
P2DE_DEV double div_fast(double n, double a) {
  double x = rcp_fast(a);
  double q = n * x;
  double r = fma(-a, q, n);
  return fma(r, x, q);
}
// sqrt: MUFU.RSQ64H seed + two coupled Newton steps + residual correction
P2DE_DEV double sqrt_fast(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  double dd = fma(-g, g, a);
  return fma(dd, h, g);
}

// ---- Gauss collocation in the generic kernel: the same reference formulas with div_fast / sqrt_fast instead of the IEEE
//      division and square-root routines (14 flux-differencing pairs per line x 6 divisions each dominate that kernel)
P2DE_DEV double logmean_fd(double aL, double aR, double logL, double logR) {
  double da = aR - aL, aavg = 0.5 * (aR + aL);
  double f = div_fast(da, aavg), v = f * f;
  if (fabs(f) < 1e-4) return aavg * (1 + v * (-0.2 - v * (0.0512 - v * 0.026038857142857)));
  return -div_fast(da, logL - logR);
}
// logmean :307-321
P2DE_DEV double logmean(double aL, double aR, double logL, double logR) {
  double da = aR - aL, aavg = 0.5 * (aR + aL);
  double f = da / aavg, v = f * f;
  if (fabs(f) < 1e-4) return aavg * (1 + v * (-0.2 - v * (0.0512 - v * 0.026038857142857)));
  return -da / (logL - logR);
}

struct Prim2 { double rho, u, v, beta, rholog, betalog; };

// fS(::Dim2) :220-249, only the component along axis d (the reference evaluates both and
// multiplies the other by an exactly-zero metric term on Cartesian meshes).
P2DE_DEV void fS_dir(double gm1, const Prim2 &L, const Prim2 &R, int d, double F[4]) {
  double rholog = logmean(L.rho, R.rho, L.rholog, R.rholog);
  double betalog = logmean(L.beta, R.beta, L.betalog, R.betalog);
  double rhoavg = 0.5 * (L.rho + R.rho), uavg = 0.5 * (L.u + R.u), vavg = 0.5 * (L.v + R.v);
  double unorm = L.u * R.u + L.v * R.v;
  double pa = rhoavg / (L.beta + R.beta);
  double f4aux = rholog / (2 * gm1 * betalog) + pa + 0.5 * rholog * unorm;
  double FxS1 = rholog * uavg, FxS3 = FxS1 * vavg;
  if (d == 0) { F[0] = FxS1; F[1] = FxS1 * uavg + pa; F[2] = FxS3; F[3] = f4aux * uavg; }
  else { double FyS1 = rholog * vavg; F[0] = FyS1; F[1] = FxS3; F[2] = FyS1 * vavg + pa; F[3] = f4aux * vavg; }
}

P2DE_DEV void fS_dir_fd(double gm1, const Prim2 &L, const Prim2 &R, int d, double F[4]) {
  double rholog = logmean_fd(L.rho, R.rho, L.rholog, R.rholog);
  double betalog = logmean_fd(L.beta, R.beta, L.betalog, R.betalog);
  double rhoavg = 0.5 * (L.rho + R.rho), uavg = 0.5 * (L.u + R.u), vavg = 0.5 * (L.v + R.v);
  double unorm = L.u * R.u + L.v * R.v;
  double pa = div_fast(rhoavg, L.beta + R.beta);
  double f4aux = div_fast(rholog, 2 * gm1 * betalog) + pa + 0.5 * rholog * unorm;
  double FxS1 = rholog * uavg, FxS3 = FxS1 * vavg;
  if (d == 0) { F[0] = FxS1; F[1] = FxS1 * uavg + pa; F[2] = FxS3; F[3] = f4aux * uavg; }
  else { double FyS1 = rholog * vavg; F[0] = FyS1; F[1] = FxS3; F[2] = FyS1 * vavg + pa; F[3] = f4aux * vavg; }
}

// Same flux with three divisions instead of six:
//   logmean(rho)      = |f| < 1e-4 ? aavg * P((da/aavg)^2) : da / (logR - logL)      (one quotient)
//   1 / logmean(beta) = |f| < 1e-4 ? (1/bavg) * (1 + 0.2 v + 0.0912 v^2)  : (logR - logL) / db
//                       (1/P(v) to O(v^3) ~ 1e-24 since v < 1e-8)
//   pa                = rhoavg / (betaL + betaR)
// `half_inv_gm1` = 1 / (2 (gamma - 1)).
P2DE_DEV void fS_fast(double half_inv_gm1, const Prim2 &L, const Prim2 &R, int d, double F[4]) {
  double da = R.rho - L.rho, aavg = 0.5 * (R.rho + L.rho);
  bool ser = fabs(da) < 1e-4 * fabs(aavg);
  double q = da / (ser ? aavg : (R.rholog - L.rholog));
  double v = q * q;
  double rholog = ser ? aavg * (1 + v * (-0.2 - v * (0.0512 - v * 0.026038857142857))) : q;
  double db = R.beta - L.beta, bavg = 0.5 * (R.beta + L.beta);
  bool serb = fabs(db) < 1e-4 * fabs(bavg);
  double qb = (serb ? 1.0 : (R.betalog - L.betalog)) / (serb ? bavg : db);
  double fb = db * qb, vb = fb * fb;
  double inv_betalog = serb ? qb * (1 + vb * (0.2 + vb * 0.0912)) : qb;
  double rhoavg = aavg, uavg = 0.5 * (L.u + R.u), vavg = 0.5 * (L.v + R.v);
  double unorm = L.u * R.u + L.v * R.v;
  double pa = rhoavg / (L.beta + R.beta);
  double f4aux = rholog * inv_betalog * half_inv_gm1 + pa + 0.5 * rholog * unorm;
  double FxS1 = rholog * uavg, FxS3 = FxS1 * vavg;
  if (d == 0) { F[0] = FxS1; F[1] = FxS1 * uavg + pa; F[2] = FxS3; F[3] = f4aux * uavg; }
  else { double FyS1 = rholog * vavg; F[0] = FyS1; F[1] = FxS3; F[2] = FyS1 * vavg + pa; F[3] = f4aux * vavg; }
}

P2DE_DEV Prim2 prim_of(double gm1, const Cons2 &U) {
  Prim2 q;
  double p = pfun2(gm1, U);
  q.rho = U.rho; q.u = U.m1 / U.rho; q.v = U.m2 / U.rho;
  q.beta = U.rho / (2 * p);                              // betafun :42-45
  q.rholog = log(U.rho); q.betalog = log(q.beta);
  return q;
}

P2DE_DEV Prim2 prim_of_fd(double gm1, const Cons2 &U) {
  Prim2 q;
  double rinv = rcp_fast(U.rho);
  double p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
  q.rho = U.rho; q.u = U.m1 * rinv; q.v = U.m2 * rinv;
  q.beta = div_fast(U.rho, 2 * p);
  q.rholog = log(U.rho); q.betalog = log(q.beta);
  return q;
}
P2DE_DEV void flux_dir_fd(double gm1, const Cons2 &U, int d, double f[4]) {
  double rinv = rcp_fast(U.rho);
  double p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
  flux_dir(U, U.m1 * rinv, U.m2 * rinv, p, d, f);
}
P2DE_DEV double wavespeed_dir_fd(double gamma, double gm1, const Cons2 &U, int d) {
  double rinv = rcp_fast(U.rho), mn = d == 0 ? U.m1 : U.m2;
  double p = gm1 * (U.E - 0.5 * (mn * mn) * rinv);
  return fabs(mn * rinv) + sqrt_fast(gamma * p * rinv);
}

// rhoe_quadratic_coefficients(::Dim2), src/dg/limiter/limiter_utils.jl:85-90
P2DE_DEV void quad_coeff_ab(const Cons2 &U, const double Pv[4], double Lrhoe, double &a, double &b) {
  a = Pv[0] * Pv[3] - 1.0 / 2.0 * (Pv[1] * Pv[1] + Pv[2] * Pv[2]);
  b = U.E * Pv[0] + U.rho * Pv[3] - U.m1 * Pv[1] - U.m2 * Pv[2] - Pv[0] * Lrhoe;
}
P2DE_DEV double quad_coeff_c(const Cons2 &U, double Lrhoe) {
  return U.E * U.rho - 1.0 / 2.0 * (U.m1 * U.m1 + U.m2 * U.m2) - U.rho * Lrhoe;
}
// root selection of rhoe_quadratic_solve, limiter_utils.jl:52-76
__device__ __noinline__ double rhoe_quadratic_roots(double ZEROTOL, double a, double b, double c) {
  double l = 1.0;
  double disc = b * b - 4 * a * c;
  if (disc >= 0) {
    double sq = sqrt(disc);
    double r1 = (-b + sq) / (2 * a), r2 = (-b - sq) / (2 * a);
    if (r1 > ZEROTOL && r2 > ZEROTOL) l = jl_min(r1, r2);
    else if (r1 > ZEROTOL && r2 < -ZEROTOL) l = r1;
    else if (r2 > ZEROTOL && r1 < -ZEROTOL) l = r2;
  }
  return l;
}
// limiting_param_bound_rho_rhoe, limiter_utils.jl:26-40 with Urho = Urhoe = Inf (positivity
// bounds): min(l_rho, quad(Lrhoe), quad(Inf) == 1.0).  `c` = quad_coeff_c(U, Lrhoe).
// The quadratic q(l) = a l^2 + b l + c has no root in (0, 1] when q(0) > 0, q(1) > 0 and its
// vertex is not an interior minimum; the reference then returns either 1 or a root > 1, which
// the trailing min(., 1.0) turns into 1, so the sqrt and the two divisions are skipped.
__device__ __noinline__ double limiting_param_pos_slow(double ZEROTOL, double rho, double P0, double Lrho, double a, double b, double c) {
  double l = 1.0;
  if (rho + P0 < Lrho) l = jl_max((Lrho - rho) / P0, 0.0);
  bool no_root = (c > 0.0) && (a + b + c > 0.0) && !(a > 0.0 && b < 0.0 && -b < 2.0 * a);
  if (!no_root) l = jl_min(l, rhoe_quadratic_roots(ZEROTOL, a, b, c));
  return jl_min(l, 1.0);
}
// common case: density bound inactive and no root in (0, 1]  ->  1 (no min, no division)
P2DE_DEV bool limiting_param_pos_easy(double rho, double P0, double Lrho, double a, double b, double c) {
  return !(rho + P0 < Lrho) && (c > 0.0) && (a + b + c > 0.0) && !(a > 0.0 && b < 0.0 && -b < 2.0 * a);
}
P2DE_DEV double limiting_param_pos(double ZEROTOL, const Cons2 &U, double c, const double Pv[4], double Lrho, double Lrhoe) {
  double a, b;
  quad_coeff_ab(U, Pv, Lrhoe, a, b);
  if (limiting_param_pos_easy(U.rho, Pv[0], Lrho, a, b, c)) return 1.0;
  return limiting_param_pos_slow(ZEROTOL, U.rho, Pv[0], Lrho, a, b, c);
}

}  // namespace p2de
