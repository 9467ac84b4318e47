// physics.cuh — pointwise compressible-Euler device functions (FP64, CUDA cores).
//
// Same formulas, same operation order as the reference's src/math/compressible_Navier_Stokes.jl
// (line numbers cited per function) so that the only differences from the CPU oracle are FMA
// contraction and libm last-bit differences (log/exp/pow); IEEE division and sqrt are kept
// (no -use_fast_math, no reciprocal tricks that would change Inf/NaN outcomes in the limiter).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace p2de {

#define P2DE_DEV __device__ __forceinline__

struct Cons2 { double rho, m1, m2, E; };

// Julia's min/max propagate NaN (oracle deviation D3): fmin/fmax do not, so spell it out.
P2DE_DEV double jl_min(double a, double b) { return (a != a || b != b) ? (a + b) : (a < b ? a : b); }
P2DE_DEV double jl_max(double a, double b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }

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
P2DE_DEV Cons2 entropy_roundtrip(double gamma, double gm1, const Cons2 &U) {
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

P2DE_DEV Prim2 prim_of(double gm1, const Cons2 &U) {
  Prim2 q;
  double p = pfun2(gm1, U);
  q.rho = U.rho; q.u = U.m1 / U.rho; q.v = U.m2 / U.rho;
  q.beta = U.rho / (2 * p);                              // betafun :42-45
  q.rholog = log(U.rho); q.betalog = log(q.beta);
  return q;
}

// rhoe_quadratic_solve, src/dg/limiter/limiter_utils.jl:52-90 (Dim2 coefficients :85-90)
P2DE_DEV double rhoe_quadratic_solve(double ZEROTOL, const Cons2 &U, const double Pv[4], double Lrhoe) {
  double a = Pv[0] * Pv[3] - 1.0 / 2.0 * (Pv[1] * Pv[1] + Pv[2] * Pv[2]);
  double b = U.E * Pv[0] + U.rho * Pv[3] - U.m1 * Pv[1] - U.m2 * Pv[2] - Pv[0] * Lrhoe;
  double c = U.E * U.rho - 1.0 / 2.0 * (U.m1 * U.m1 + U.m2 * U.m2) - U.rho * Lrhoe;
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
// limiting_param_bound_rho_rhoe, limiter_utils.jl:26-40 with Urho = Urhoe = Inf (positivity bounds)
P2DE_DEV double limiting_param_pos(double ZEROTOL, const Cons2 &U, const double Pv[4], double Lrho, double Lrhoe) {
  double l = 1.0;
  if (U.rho + Pv[0] < Lrho) l = jl_max((Lrho - U.rho) / Pv[0], 0.0);
  // min(l, quad(Lrhoe), quad(Urhoe = Inf) == 1.0)
  return jl_min(jl_min(l, rhoe_quadratic_solve(ZEROTOL, U, Pv, Lrhoe)), 1.0);
}

}  // namespace p2de
