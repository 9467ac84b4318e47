// gauss.cuh — entropy projection to the face nodes for 2D Gauss collocation.
//
// On Gauss nodes the face nodes are not volume nodes: the face state the fluxes see is
//   u_tilde_f = u( Vf_theta v(Uq) ),   Vf_theta = theta Vf + (1 - theta) Vf_low        (rhs.jl:59-133)
// with theta = 1 (NoEntropyProjectionLimiter) or the largest theta in [0, 1] for which the
// projected face state stays within bounds of the unlimited extrapolation, found per face node by
// the 21-step bisection of NodewiseScaledExtrapolation (src/dg/filter.jl:6-130,
// src/math/nonlinear_solvers.jl:3-20).
//
// Vf acts along grid lines (tensor-product Gauss), so the same "line thread" mapping as the stage
// kernel is used: thread (d, line) of an element owns the two face nodes at the ends of its line.
// Outputs: utf[K][Nfp][4] (consumed by stage_kernel for this element AND its neighbours, which is
// why this is its own kernel: a grid-wide dependency), theta_local[K][Nfp], theta[K].
#pragma once
#include "kernels2d.cuh"

namespace p2de {

struct ProjArgs {
  const double *Uq;
  double *utf;            // [K][Nfp][4]
  double *theta_local;    // [K][Nfp] of this stage
  double *theta;          // [K] of this stage
  double gamma, POSTOL, zeta, eta;
  int nodewise;
};

// v_ufun2 (v_ufun(::Dim2), compressible_Navier_Stokes.jl:134-144) lives in kernels2d.cuh
// u_vfun(::Dim2), :155-163 with s_vfun :93-97 and rhoe_vfun :99-104
P2DE_DEV Cons2 u_vfun2(double gamma, double gm1, const double v[4]) {
  double q = v[1] * v[1] + v[2] * v[2];
  double sv = gamma - v[0] + q / (2 * v[3]);
  double rhoeV = pow(gm1 / pow(-v[3], gamma), 1 / gm1) * exp(-sv / gm1);
  Cons2 W;
  W.rho = -rhoeV * v[3]; W.m1 = rhoeV * v[1]; W.m2 = rhoeV * v[2];
  W.E = rhoeV * (1 - q / (2 * v[3]));
  return W;
}

// The same two maps with the powers folded into one logarithm / one exponential
//   s = log(p / rho^gamma) = log p - gamma log rho,
//   rhoe(v) = ((gamma-1) / (-v4)^gamma)^(1/(gamma-1)) exp(-s/(gamma-1)) = exp((log(gamma-1) - gamma log(-v4) - s) / (gamma-1))
// and Newton reciprocals: a few 1e-16 relative away from the reference's evaluation order (pow is ~2 ulp itself),
// at a third of its cost; the projection kernel is nothing but these two functions.
P2DE_DEV void v_ufun2_fd(double gamma, double gm1, const Cons2 &U, double v[4]) {
  const double rinv = rcp_fast(U.rho);
  const double p = gm1 * (U.E - 0.5 * (U.m1 * U.m1 + U.m2 * U.m2) * rinv);
  const double s = log(p) - gamma * log(U.rho);
  const double g = gm1 * rcp_fast(p);
  v[0] = (gamma + 1 - s) - g * U.E;
  v[1] = U.m1 * g; v[2] = U.m2 * g; v[3] = -U.rho * g;
}
P2DE_DEV Cons2 u_vfun2_fd(double gamma, double gm1, double log_gm1, const double v[4]) {
  const double q = v[1] * v[1] + v[2] * v[2];
  const double h = 0.5 * q * rcp_fast(v[3]);          // q / (2 v4)
  const double sv = gamma - v[0] + h;
  const double rhoeV = exp((log_gm1 - gamma * log(-v[3]) - sv) * rcp_fast(gm1));
  Cons2 W;
  W.rho = -rhoeV * v[3]; W.m1 = rhoeV * v[1]; W.m2 = rhoeV * v[2];
  W.E = rhoeV * (1 - h);
  return W;
}

template <int N1D, int EPB>
__global__ void __launch_bounds__(EPB * 2 * N1D)
gauss_project_kernel(const __grid_constant__ ProjArgs A, const __grid_constant__ MeshTopo M,
                     const __grid_constant__ Tables2D<N1D> Tc) {
  constexpr int Nq = N1D * N1D, Nfp = 4 * N1D, TPE = 2 * N1D, S = EPB * Nq;
  __shared__ double su[S * 4], sv[S * 4];
  __shared__ double sth[EPB * Nfp];
  const int tid = threadIdx.x, el = tid / TPE, ln = tid % TPE, d = ln / N1D, line = ln % N1D;
  const long long k = (long long)blockIdx.x * EPB + el;
  const bool active = k < M.K;
  const double gamma = A.gamma, gm1 = A.gamma - 1.0, log_gm1 = log(gm1);
  if (active)
    for (int node = ln; node < Nq; node += TPE) {
      Cons2 U = load_cons(A.Uq + (k * Nq + node) * 4);
      double v[4];
      v_ufun2_fd(gamma, gm1, U, v);
      double *pu = su + (el * Nq + node) * 4, *pv = sv + (el * Nq + node) * 4;
      pu[0] = U.rho; pu[1] = U.m1; pu[2] = U.m2; pu[3] = U.E;
      pv[0] = v[0]; pv[1] = v[1]; pv[2] = v[2]; pv[3] = v[3];
    }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int f = (2 * d + e) * N1D + line, ae = e ? N1D - 1 : 0;
      double w[N1D], Uf[4] = {0, 0, 0, 0}, VUf[4] = {0, 0, 0, 0};
      const double *vn[N1D];
#pragma unroll
      for (int a = 0; a < N1D; ++a) {
        const int node = d == 0 ? a + line * N1D : line + a * N1D;
        w[a] = Tc.VfL[d][line][e][a];
        vn[a] = sv + (el * Nq + node) * 4;
        const double *un = su + (el * Nq + node) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) { Uf[c] += w[a] * un[c]; VUf[c] += w[a] * vn[a][c]; }   // calc_face_values! filter.jl:26-41
      }
      Cons2 UfC; UfC.rho = Uf[0]; UfC.m1 = Uf[1]; UfC.m2 = Uf[2]; UfC.E = Uf[3];
      const double rhoef = rhoe2(UfC);
      // v_tilde_f(theta) = sum_j (theta Vf + (1 - theta) Vf_low)[f, j] vq[j]  (filter.jl:119-122, rhs.jl:88-93)
      auto vtilde = [&](double th, double vt[4]) {
        vt[0] = vt[1] = vt[2] = vt[3] = 0.0;
#pragma unroll
        for (int a = 0; a < N1D; ++a) {
          const double wa = th * w[a] + (1 - th) * (a == ae ? 1.0 : 0.0);
#pragma unroll
          for (int c = 0; c < 4; ++c) vt[c] += wa * vn[a][c];
        }
      };
      // update_and_check_bound_limited_entropyproj_var_on_face_node! :100-130, check_bound_on_face_node :84-98
      // (returns u(v_tilde_f(theta)) too: the accepted theta's state is the kernel's output)
      auto ok = [&](double th, Cons2 &ut) {
        double vt[4];
        vtilde(th, vt);
        ut = u_vfun2_fd(gamma, gm1, log_gm1, vt);
        if (!(vt[3] < -A.POSTOL)) return false;
        const double rhoe = rhoe2(ut);
        return vt[3] < jl_min(A.zeta * VUf[3], -A.POSTOL) && ut.rho > jl_max((1 - A.eta) * Uf[0], A.POSTOL) &&
               ut.rho < (1 + A.eta) * Uf[0] && rhoe > jl_max((1 - A.eta) * rhoef, A.POSTOL) && rhoe < (1 + A.eta) * rhoef;
      };
      double th = 1.0;
      Cons2 ut;
      const bool ok1 = ok(1.0, ut);           // NoEntropyProjectionLimiter: theta = 1 whatever the test says
      if (A.nodewise && !ok1) {   // bisection(f, 0.0, 1.0), nonlinear_solvers.jl:3-20
        double xv = 0.0, xi = 1.0;
        Cons2 tmp;
        for (int it = 0; it <= 20; ++it) {
          const double xn = 0.5 * (xv + xi);
          if (ok(xn, tmp)) xv = xn; else xi = xn;
        }
        th = xv;
        double vt[4];
        vtilde(th, vt);
        ut = u_vfun2_fd(gamma, gm1, log_gm1, vt);
      }
      double r[4] = {ut.rho, ut.m1, ut.m2, ut.E};
      store4(A.utf + (k * Nfp + f) * 4, r);
      if (A.theta_local) A.theta_local[k * Nfp + f] = th;
      sth[el * Nfp + f] = th;
    }
  }
  __syncthreads();
  if (active && ln == 0 && A.theta) {   // theta[k] = mean of its face nodes (filter.jl:57)
    double s = 0.0;
    for (int f = 0; f < Nfp; ++f) s += sth[el * Nfp + f];
    A.theta[k] = s / Nfp;
  }
}

}  // namespace p2de
