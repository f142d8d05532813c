// P1 (boundary projection, boundary_projection.py:33-57) for a safe state set whose generator is WIDE (BalanceQuadrotor:
// G_s 6 x 24, m = 2).  The containment encoding enc(Z(c_f + B a, G_w) in Z(c_s, G_s)) (zonotope.py:131-141) cannot be
// reduced to a fixed polygon in action space (its feasible offsets form a 6-D polytope with ~10^4 facets), so the program
// is solved per environment in its lifted form:
//
//   min_a 1/2 |a - a0|^2   s.t.  -1 <= a <= 1,  a in the safe action polygon,
//        exists beta, Gamma:  G beta = c_s - c_f - B a,  G Gamma = G_w,  |beta_i| + sum_j |Gamma_ij| <= 1  for every generator i
//
// The equalities are eliminated with constants computed once at registration (generators of equal direction merged):
//   beta = p - Q a + N y,   Gamma_j = Gamma_p,j + N z_j      (p = G^+ (c_s - c_f), Q = G^+ B per environment; N = null(G))
// and every row budget becomes 2^(1+q) linear rows  s0 beta_i + s1 Gamma_i1 + s2 Gamma_i2 <= 1  (signs s = +-1).  Unknowns
// x = (a, y, z_1, z_2): 2 + 3 (p' - n) = 38 for the quadrotor, 164 rows.  Primal-dual interior point in DOUBLE per thread
// (dense normal equations + Cholesky in local memory), then one Newton step of the same KKT system with the environment-
// dependent row parts (Q, p, a0) in the dual-number type: implicit-function derivative of a_s w.r.t. (a, c_f, B).
// The solve only runs for environments whose gate asks for the safeguarded action.
#pragma once
#include "sgbpo_safeguard2.cuh"

namespace sgbpo {

constexpr int LF_MAXV = 2 + 3 * MAXNULL;
constexpr int LF_MAXR = 4 + MAXHP + 8 * MAXGEN;

// lower-triangular packed symmetric matrix helpers
SG_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }      // i >= j

SG_HD bool chol_packed(double* K, int n) {
  for (int j = 0; j < n; ++j) {
    double* Kj = K + tri(j, 0);
    const double orig = Kj[j];
    double d = orig;
    for (int p = 0; p < j; ++p) d -= Kj[p] * Kj[p];
    if (!(d == d)) return false;
    if (!(d > 1e-12 * orig)) d = 1e60;        // direction left free by the active rows: frozen (see chol6)
    d = sqrt(d);
    Kj[j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double* Ki = K + tri(i, 0);
      double v = Ki[j];
      for (int p = 0; p < j; ++p) v -= Ki[p] * Kj[p];
      Ki[j] = v * inv;
    }
  }
  return true;
}
template <class W> SG_HD void chol_packed_solve(const double* L, int n, W* x) {
  for (int i = 0; i < n; ++i) {
    const double* Li = L + tri(i, 0);
    W v = x[i];
    for (int p = 0; p < i; ++p) v = v - W(Li[p]) * x[p];
    x[i] = v / W(Li[i]);
  }
  for (int i = n - 1; i >= 0; --i) {
    W v = x[i];
    for (int p = i + 1; p < n; ++p) v = v - W(L[tri(p, 0) + i]) * x[p];
    x[i] = v / W(L[tri(i, 0) + i]);
  }
}

// per-environment data of the lifted program in scalar type W
// The two leading unknowns u parametrise the action,  a = base + M u :  P1: a = u (base 0, M = I), objective
// 1/2 |u - a0|^2;  P4 (max t with start + t d safe): u = (t, unused), base = start, M = [d 0], objective -t, extra row t >= 0.
template <class W> struct LiftedEnv {
  W p[MAXGEN];            // beta = p - Q u + N y  (already in terms of u)
  W Q[MAXGEN][2];
  W base[2], M[2][2];
  double h[2];            // objective 1/2 sum_j h_j u_j^2 + lin_j u_j
  W lin[2];
  int extra;              // 1: one more row  -u_0 <= 0
};

// Row layout.  Rows 0..3: the action box; then n_act rows of the safe action polygon (coefficients on the action only);
// then, for generator i and sign combination c = (s0, s1, s2), the budget row
//     s0 beta_i + s1 Gamma_i1 + s2 Gamma_i2 <= 1,   beta_i = p_i - Q_i a + N_i y,   Gamma_ij = Gam_p,ij + N_i z_j .
// Every budget row of one generator has the coefficient vector (-s0 Q_i, s0 N_i, s1 N_i, s2 N_i): the normal-equation
// matrix  sum_r w_r coef_r coef_r^T  is a sum over generators of 4 x 4 sign-weighted sums times the outer products of
// (Q_i, N_i) -- ~8x fewer operations than row-by-row accumulation, and the row sweeps need 4 dot products per generator.
template <class W, class T>
SG_HD void lifted_small_row(int r, const LiftedEnv<W>& E, const SafeConsts<T>& k, const ActRows<T>* dyn, int n_act,
                            W& c0, W& c1, W& g) {
  double n0, n1, h;
  if (r < 4) {
    const double sg = (r & 1) ? -1.0 : 1.0;
    n0 = (r >> 1) ? 0.0 : sg; n1 = (r >> 1) ? sg : 0.0; h = 1.0;
  } else if (r < 4 + n_act) {
    r -= 4;
    if (dyn && dyn->count > 0) { n0 = (double)dyn->n[r][0]; n1 = (double)dyn->n[r][1]; h = (double)dyn->h[r]; }
    else { n0 = (double)k.act_hp[r][0]; n1 = (double)k.act_hp[r][1]; h = (double)k.act_hp[r][2]; }
  } else {
    c0 = W(-1.0); c1 = W(0.0); g = W(0.0);
    return;
  }
  c0 = W(n0) * E.M[0][0] + W(n1) * E.M[1][0];
  c1 = W(n0) * E.M[0][1] + W(n1) * E.M[1][1];
  g = W(h) - W(n0) * E.base[0] - W(n1) * E.base[1];
}

// (beta_i, Gamma_i1, Gamma_i2) at x; with affine = false only the part linear in x (for a step direction)
template <class W, class T>
SG_HD void lifted_gen_vals(int i, const LiftedEnv<W>& E, const SafeConsts<T>& k, int nn, const double* x, bool affine, W* b) {
  double dy = 0.0, d1 = 0.0, d2 = 0.0;
  for (int j = 0; j < nn; ++j) {
    const double nij = (double)k.Nmat[i][j];
    dy += nij * x[2 + j];
    if (k.n_wcol > 0) d1 += nij * x[2 + nn + j];
    if (k.n_wcol > 1) d2 += nij * x[2 + 2 * nn + j];
  }
  b[0] = W(dy) - E.Q[i][0] * W(x[0]) - E.Q[i][1] * W(x[1]);
  b[1] = W(d1); b[2] = W(d2);
  if (affine) {
    b[0] = b[0] + E.p[i];
    if (k.n_wcol > 0) b[1] = b[1] + W((double)k.Gam_p[i][0]);
    if (k.n_wcol > 1) b[2] = b[2] + W((double)k.Gam_p[i][1]);
  }
}

// Normal equations K (packed lower triangle, primal parts) and right-hand side of one Newton step of the perturbed KKT
// system at (x, s, lam).  target >= 0: complementarity target of the centring step; target < 0: every row keeps its own
// lam_r s_r (the implicit-function step after convergence).  Returns the largest primal residual.
template <class W, class T>
SG_HD double lifted_assemble(const LiftedEnv<W>& E, const SafeConsts<T>& k, const ActRows<T>* dyn, int n_act, int nn, int nv,
                             const double* x, const double* s, const double* lam, double target, double* K, W* rhs) {
  for (int q = 0; q < nv * (nv + 1) / 2; ++q) K[q] = 0.0;
  for (int j = 0; j < nv; ++j) rhs[j] = W(0.0);
  rhs[0] = W(-E.h[0] * x[0]) - E.lin[0]; rhs[1] = W(-E.h[1] * x[1]) - E.lin[1];          // -(H x + lin)
  double k00 = E.h[0], k10 = 0.0, k11 = E.h[1], rp_max = 0.0;
  const int nsmall = 4 + n_act + E.extra, combos = 1 << (1 + k.n_wcol);
  for (int r = 0; r < nsmall; ++r) {
    W c0, c1, g;
    lifted_small_row<W, T>(r, E, k, dyn, n_act, c0, c1, g);
    const W rp = c0 * W(x[0]) + c1 * W(x[1]) + W(s[r]) - g;
    rp_max = fmax(rp_max, fabs(primal(rp)));
    const double w = lam[r] / s[r], p0 = primal(c0), p1 = primal(c1);
    const W t = W(lam[r]) + (W(lam[r]) * rp - W(target < 0.0 ? 0.0 : lam[r] * s[r] - target)) / W(s[r]);
    rhs[0] = rhs[0] - c0 * t; rhs[1] = rhs[1] - c1 * t;
    k00 += w * p0 * p0; k10 += w * p1 * p0; k11 += w * p1 * p1;
  }
  const int oy = 2, o1 = 2 + nn, o2 = 2 + 2 * nn;
  for (int i = 0; i < k.n_gen; ++i) {
    W b[3];
    lifted_gen_vals<W, T>(i, E, k, nn, x, true, b);
    double sw = 0.0, m01 = 0.0, m02 = 0.0, m12 = 0.0;
    W t0 = W(0.0), t1 = W(0.0), t2 = W(0.0);
    for (int c = 0; c < combos; ++c) {
      const int r = nsmall + i * combos + c;
      const double s0 = (c & 1) ? -1.0 : 1.0, s1 = (c & 2) ? -1.0 : 1.0, s2 = (c & 4) ? -1.0 : 1.0;
      W rp = W(s0) * b[0] + W(s[r] - 1.0);
      if (k.n_wcol > 0) rp = rp + W(s1) * b[1];
      if (k.n_wcol > 1) rp = rp + W(s2) * b[2];
      rp_max = fmax(rp_max, fabs(primal(rp)));
      const double w = lam[r] / s[r];
      const W t = W(lam[r]) + (W(lam[r]) * rp - W(target < 0.0 ? 0.0 : lam[r] * s[r] - target)) / W(s[r]);
      sw += w; m01 += w * s0 * s1; m02 += w * s0 * s2; m12 += w * s1 * s2;
      t0 = t0 + W(s0) * t; t1 = t1 + W(s1) * t; t2 = t2 + W(s2) * t;
    }
    const double q0 = primal(E.Q[i][0]), q1 = primal(E.Q[i][1]);
    rhs[0] = rhs[0] + E.Q[i][0] * t0; rhs[1] = rhs[1] + E.Q[i][1] * t0;
    k00 += sw * q0 * q0; k10 += sw * q1 * q0; k11 += sw * q1 * q1;
    double nd[MAXNULL];                       // row i of the null-space basis, converted once
    for (int j = 0; j < nn; ++j) nd[j] = (double)k.Nmat[i][j];
    for (int j = 0; j < nn; ++j) {
      const double nj = nd[j];
      rhs[oy + j] = rhs[oy + j] - W(nj) * t0;
      double* Ky = K + tri(oy + j, 0);
      Ky[0] -= sw * nj * q0; Ky[1] -= sw * nj * q1;
      const double a = sw * nj;
      for (int l = 0; l <= j; ++l) Ky[oy + l] += a * nd[l];
      if (k.n_wcol > 0) {
        rhs[o1 + j] = rhs[o1 + j] - W(nj) * t1;
        double* K1 = K + tri(o1 + j, 0);
        K1[0] -= m01 * nj * q0; K1[1] -= m01 * nj * q1;
        const double a1 = m01 * nj;
        for (int l = 0; l < nn; ++l) K1[oy + l] += a1 * nd[l];
      }
      if (k.n_wcol > 1) {
        rhs[o2 + j] = rhs[o2 + j] - W(nj) * t2;
        double* K2 = K + tri(o2 + j, 0);
        K2[0] -= m02 * nj * q0; K2[1] -= m02 * nj * q1;
        const double a2 = m02 * nj, a12 = m12 * nj;
        for (int l = 0; l < nn; ++l) {
          K2[oy + l] += a2 * nd[l];
          K2[o1 + l] += a12 * nd[l];
        }
      }
    }
  }
  K[tri(0, 0)] = k00; K[tri(1, 0)] = k10; K[tri(1, 1)] = k11;
  // the diagonal blocks of the three lifted groups are equal (every sign squares to one)
  for (int j = 0; j < nn; ++j)
    for (int l = 0; l <= j; ++l) {
      const double v = K[tri(oy + j, oy + l)];
      if (k.n_wcol > 0) K[tri(o1 + j, o1 + l)] = v;
      if (k.n_wcol > 1) K[tri(o2 + j, o2 + l)] = v;
    }
  return rp_max;
}

// p, Q in terms of the action (base 0, M = I); the callers re-parametrise
template <class TASK, class W, class S, class T>
SG_HD void lifted_env(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, LiftedEnv<W>& E) {
  constexpr int NS = TASK::NS;
  for (int i = 0; i < k.n_gen; ++i) {
    W p = W(0.0), q0 = W(0.0), q1 = W(0.0);
    for (int j = 0; j < NS; ++j) {
      const W gp = W((double)k.Gpinv[i][j]);
      p = p + gp * (W((double)k.c_s[j]) - widen(c_f[j]));
      q0 = q0 + gp * widen(B[j][0]);
      q1 = q1 + gp * widen(B[j][1]);
    }
    E.p[i] = p; E.Q[i][0] = q0; E.Q[i][1] = q1;
  }
  E.base[0] = W(0.0); E.base[1] = W(0.0);
  E.M[0][0] = W(1.0); E.M[0][1] = W(0.0); E.M[1][0] = W(0.0); E.M[1][1] = W(1.0);
  E.h[0] = 1.0; E.h[1] = 1.0; E.lin[0] = W(0.0); E.lin[1] = W(0.0);
  E.extra = 0;
}
template <class W> SG_HD void lifted_set_p1(LiftedEnv<W>& E, const W* a0) {
  E.lin[0] = W(0.0) - a0[0]; E.lin[1] = W(0.0) - a0[1];
}
template <class W, class T> SG_HD void lifted_set_p4(LiftedEnv<W>& E, const SafeConsts<T>& k, const W* start, const W* d) {
  for (int i = 0; i < k.n_gen; ++i) {
    E.p[i] = E.p[i] - E.Q[i][0] * start[0] - E.Q[i][1] * start[1];
    E.Q[i][0] = E.Q[i][0] * d[0] + E.Q[i][1] * d[1];
    E.Q[i][1] = W(0.0);
  }
  E.base[0] = start[0]; E.base[1] = start[1];
  E.M[0][0] = d[0]; E.M[1][0] = d[1]; E.M[0][1] = W(0.0); E.M[1][1] = W(0.0);
  E.h[0] = 0.0; E.h[1] = 1.0; E.lin[0] = W(-1.0); E.lin[1] = W(0.0);
  E.extra = 1;
}

// Interior point on the program described by E (primal values), then the implicit-function Newton step with ES (the same
// program with the environment-dependent parts in the dual-number type).  out: the two leading unknowns.
// x0: start of the iteration.  returns 0: failed, 1: solved, 2: solved and the optimum is the point `ref` (within 1e-7;
// no Newton step taken, `out` untouched)
template <class SW, class T>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int lifted_solve(const LiftedEnv<double>& E, const LiftedEnv<SW>& ES, const SafeConsts<T>& k, const ActRows<T>* dyn,
                 const double* x0, const double* ref, SW* out) {
  const int nn = k.n_null, nv = 2 + nn * (1 + k.n_wcol);
  const int n_act = k.action_constrained ? ((dyn && dyn->count > 0) ? dyn->count : k.n_act_hp) : 0;
  const int nsmall = 4 + n_act + E.extra, combos = 1 << (1 + k.n_wcol);
  const int nr = nsmall + k.n_gen * combos;
  bool ok = false;
  double x[LF_MAXV], s[LF_MAXR], lam[LF_MAXR], K[LF_MAXV * (LF_MAXV + 1) / 2], rhsv[LF_MAXV];
  double mu = 1.0;
  int stalls = 0;
  double alpha_prev = 0.0;
  for (int j = 0; j < nv; ++j) x[j] = 0.0;
  x[0] = x0[0]; x[1] = x0[1];
  for (int r = 0; r < nsmall; ++r) {
    double c0, c1, g;
    lifted_small_row<double, T>(r, E, k, dyn, n_act, c0, c1, g);
    s[r] = fmax(g - c0 * x[0] - c1 * x[1], 0.1);
    lam[r] = 1.0 / s[r];
  }
  for (int i = 0; i < k.n_gen; ++i) {
    double b[3];
    lifted_gen_vals<double, T>(i, E, k, nn, x, true, b);
    for (int c = 0; c < combos; ++c) {
      const int r = nsmall + i * combos + c;
      const double v = ((c & 1) ? -b[0] : b[0]) + ((c & 2) ? -b[1] : b[1]) + ((c & 4) ? -b[2] : b[2]);
      s[r] = fmax(1.0 - v, 0.1);
      lam[r] = 1.0 / s[r];
    }
  }
  for (int it = 0; it < 80; ++it) {
    double gap = 0.0;
    for (int r = 0; r < nr; ++r) gap += lam[r] * s[r];
    mu = gap / nr;
    // centring parameter: a tenth of the current gap, a hundredth after a (nearly) full step (13.5 -> 11 iterations on
    // the fixtures; overridable for experiments)
#ifndef SGBPO_LIFTED_SIGMA
#define SGBPO_LIFTED_SIGMA(ap) ((ap) > 0.95 ? 0.01 : 0.1)
#endif
    const double target = fmax(SGBPO_LIFTED_SIGMA(alpha_prev) * mu, 5e-11);
    const double rp_max = lifted_assemble<double, T>(E, k, dyn, n_act, nn, nv, x, s, lam, target, K, rhsv);
    if (!chol_packed(K, nv)) break;
    chol_packed_solve<double>(K, nv, rhsv);                  // rhsv := dx
    // only the leading unknowns are unique at the optimum (the lifted variables have a face of optimal values): the
    // stopping test looks at their part of the step
    double alpha = 1.0;
    const double dx_max = fmax(fabs(rhsv[0]), fabs(rhsv[1]));
    for (int pass = 0; pass < 2; ++pass) {                  // pass 0: step length; pass 1: slack and multiplier update
      for (int r = 0; r < nsmall; ++r) {
        double c0, c1, g;
        lifted_small_row<double, T>(r, E, k, dyn, n_act, c0, c1, g);
        const double rp = c0 * x[0] + c1 * x[1] + s[r] - g;
        const double ds = -rp - (c0 * rhsv[0] + c1 * rhsv[1]);
        const double dl = (-(lam[r] * s[r] - target) - lam[r] * ds) / s[r];
        if (pass == 0) {
          if (ds < 0.0) alpha = fmin(alpha, -0.995 * s[r] / ds);
          if (dl < 0.0) alpha = fmin(alpha, -0.995 * lam[r] / dl);
        } else { s[r] += alpha * ds; lam[r] += alpha * dl; }
      }
      for (int i = 0; i < k.n_gen; ++i) {
        double b[3], d[3];
        lifted_gen_vals<double, T>(i, E, k, nn, x, true, b);
        lifted_gen_vals<double, T>(i, E, k, nn, rhsv, false, d);
        for (int c = 0; c < combos; ++c) {
          const int r = nsmall + i * combos + c;
          const double s0 = (c & 1) ? -1.0 : 1.0, s1 = (c & 2) ? -1.0 : 1.0, s2 = (c & 4) ? -1.0 : 1.0;
          const double rp = s0 * b[0] + s1 * b[1] + s2 * b[2] + s[r] - 1.0;
          const double ds = -rp - (s0 * d[0] + s1 * d[1] + s2 * d[2]);
          const double dl = (-(lam[r] * s[r] - target) - lam[r] * ds) / s[r];
          if (pass == 0) {
            if (ds < 0.0) alpha = fmin(alpha, -0.995 * s[r] / ds);
            if (dl < 0.0) alpha = fmin(alpha, -0.995 * lam[r] / dl);
          } else { s[r] += alpha * ds; lam[r] += alpha * dl; }
        }
      }
    }
    for (int j = 0; j < nv; ++j) x[j] += alpha * rhsv[j];
    alpha_prev = alpha;
#ifdef SGBPO_LIFTED_DEBUG
    printf("it %d mu %.3e rp %.3e alpha %.3e dx %.3e x %.6f %.6f\n", it, mu, rp_max, alpha, dx_max, x[0], x[1]);
#endif
    if (rp_max < 1e-9 && mu < 2e-10 && alpha * dx_max < 1e-9) {
      ok = true;
#ifdef SGBPO_LIFTED_COUNT
      printf("ITERS %d\n", it + 1);
#endif
      break;
    }
    // infeasible program: the step length collapses while the primal residual stays (observed: alpha < 1e-4 from the
    // ~10th iteration on, feasible programs never go below 0.05) -- stop instead of spending all iterations
    if (alpha < 1e-4 && rp_max > 1e-7) { if (++stalls >= 3) break; } else stalls = 0;
  }
  if (!ok) return 0;
  if (fabs(x[0] - ref[0]) < 1e-7 && fabs(x[1] - ref[1]) < 1e-7) return 2;
  SW rhsS[LF_MAXV];
  lifted_assemble<SW, T>(ES, k, dyn, n_act, nn, nv, x, s, lam, -1.0, K, rhsS);      // row target := lam_r s_r
  if (!chol_packed(K, nv)) return 0;
  chol_packed_solve<SW>(K, nv, rhsS);
  out[0] = SW(x[0]) + rhsS[0];
  out[1] = SW(x[1]) + rhsS[1];
  return 1;
}

template <class TASK, class S, class T>
SG_HD void lifted_primal_inputs(const S* c_f, const S (*B)[MAXM], T* cf_p, T (*B_p)[MAXM]) {
  for (int i = 0; i < TASK::NS; ++i) { cf_p[i] = primal(c_f[i]); B_p[i][0] = primal(B[i][0]); B_p[i][1] = primal(B[i][1]); }
}

// P1: the safe action closest to a
template <class TASK, class S, class T>
SG_HD bool project_lifted_m2(const S* a, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, const ActRows<T>* dyn,
                             S* out) {
  using SW = typename Widen<S>::type;
  T cf_p[MAXN];
  T B_p[MAXN][MAXM];
  lifted_primal_inputs<TASK, S, T>(c_f, B, cf_p, B_p);
  LiftedEnv<double> E;
  LiftedEnv<SW> ES;
  lifted_env<TASK, double, T, T>(cf_p, B_p, k, E);
  lifted_env<TASK, SW, S, T>(c_f, B, k, ES);
  const double a0[2] = {(double)primal(a[0]), (double)primal(a[1])};
  const SW a0S[2] = {widen(a[0]), widen(a[1])};
  lifted_set_p1<double>(E, a0);
  lifted_set_p1<SW>(ES, a0S);
  // start inside the box; if the optimum turns out to be the action itself (already safe: the barrier centre sits ~1e-8
  // inside) return it exactly -- identity map, identity Jacobian, what the exact projection is there
  const double x0[2] = {fmin(fmax(a0[0], -0.9), 0.9), fmin(fmax(a0[1], -0.9), 0.9)};
  SW res[2];
  const int st = lifted_solve<SW, T>(E, ES, k, dyn, x0, a0, res);
  if (st == 0) { out[0] = C<S>(NAN); out[1] = C<S>(NAN); return false; }
  if (st == 2) {
    out[0] = a[0]; out[1] = a[1];
    return true;
  }
  narrow(res[0], out[0]);
  narrow(res[1], out[1]);
  return true;
}

// P4 (ray_mask.py:275-334): the largest t >= 0 with  start + t d  feasible, action-safe and state-safe
template <class TASK, class S, class T>
SG_HD bool ray_exit_lifted_m2(const S* start, const S* d, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k,
                              const ActRows<T>* dyn, S& t_out) {
  using SW = typename Widen<S>::type;
  T cf_p[MAXN];
  T B_p[MAXN][MAXM];
  lifted_primal_inputs<TASK, S, T>(c_f, B, cf_p, B_p);
  LiftedEnv<double> E;
  LiftedEnv<SW> ES;
  lifted_env<TASK, double, T, T>(cf_p, B_p, k, E);
  lifted_env<TASK, SW, S, T>(c_f, B, k, ES);
  const double st0[2] = {(double)primal(start[0]), (double)primal(start[1])}, d0[2] = {(double)primal(d[0]), (double)primal(d[1])};
  const SW stS[2] = {widen(start[0]), widen(start[1])}, dS[2] = {widen(d[0]), widen(d[1])};
  lifted_set_p4<double, T>(E, k, st0, d0);
  lifted_set_p4<SW, T>(ES, k, stS, dS);
  const double x0[2] = {0.0, 0.0}, never[2] = {-1.0, 0.0};            // t >= 0: the solve always returns its own point
  SW res[2];
  const int st = lifted_solve<SW, T>(E, ES, k, dyn, x0, never, res);
  if (st != 1) { t_out = C<S>(NAN); return false; }
  narrow(res[0], t_out);
  if (primal(t_out) < T(0)) t_out = C<S>(0.0);
  return true;
}

// Orthogonal ray mask (ray_mask.py:268-341) on the lifted program: P1 gives the boundary point, P4 the chord through the
// safe region along the projection direction; the centre is the chord's midpoint (same steps as safeguard_m2's branch).
template <class TASK, class S, class T>
SG_HD bool ray_mask_orthogonal_lifted_m2(const S* a, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k,
                                         const ActRows<T>* dyn, S* out) {
  S start[2], center[2], safe_dist, feas_dist;
  bool feasible = project_lifted_m2<TASK>(a, c_f, B, k, dyn, start);
  if (!feasible) { out[0] = C<S>(NAN); out[1] = C<S>(NAN); return false; }
  const S gx = start[0] - a[0], gy = start[1] - a[1];
  const S gap = norm2(gx, gy);
  if (primal(gap) < T(1e-8)) {
    center[0] = start[0]; center[1] = start[1]; safe_dist = C<S>(1.0); feas_dist = C<S>(1.0);
  } else {
    S dir[2] = {gx / gap, gy / gap};
    S t;
    feasible = ray_exit_lifted_m2<TASK>(start, dir, c_f, B, k, dyn, t);
    if (!feasible) { out[0] = C<S>(NAN); out[1] = C<S>(NAN); return false; }
    safe_dist = t * C<S>(0.5);
    center[0] = start[0] + dir[0] * safe_dist;
    center[1] = start[1] + dir[1] * safe_dist;
    S nd[2] = {-dir[0], -dir[1]};
    feas_dist = unit_box_dist_m2(center, nd);
  }
  const S dx = a[0] - center[0], dy = a[1] - center[1];
  const S dist = norm2(dx, dy);
  if (primal(dist) < T(1e-8)) { out[0] = center[0]; out[1] = center[1]; return true; }
  const S scale = radial_map(dist, safe_dist, feas_dist, k) / (dist + C<S>(1e-8));
  out[0] = center[0] + dx * scale;
  out[1] = center[1] + dy * scale;
  return true;
}

}  // namespace sgbpo
