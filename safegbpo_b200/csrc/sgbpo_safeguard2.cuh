// m = 2 safeguard maps and the containment verdicts.
//
// For 2-D actions every program of safeguards/{boundary_projection,ray_mask}.py becomes geometry on a
// convex polygon  P = { a : n_r . a <= h_r }  in action space whose rows are affine in (c_f, B):
//   feasibility box (4 rows), safe-action zonotope polygon (2 p_a rows, registered as half-planes),
//   state block, model 0:  |(G_s^-1 (c_s - c_f - B a))_i| <= rb_i  (2 n rows).
// P1 = Euclidean projection onto P (active set of <= 2 rows), P3/P4 = ray/polygon intersection.
#pragma once
#include "sgbpo_safeguard.cuh"

namespace sgbpo {

// per-environment safe state set of this step in the form the m = 2 row builder consumes
template <class T> struct StateDyn {
  T Ginv[MAXN * MAXN];   // G_s^-1 (row-major n x n)
  T cs_hat[MAXN];        // G_s^-1 c_s
  T rb[MAXN];            // 1 - row sums of |G_s^-1 G_w|
  T center[MAXN];        // c_s
  T gen[MAXN * MAXN];    // G_s (row-major), for the public safe_state_set() accessor
  bool degenerate;       // the reference's fallback blocks (or an infeasible support-point program): no usable set
};

// per-environment safe action set computed this step (NavigateSeeker, P5): rows n . a <= h, constants for autograd
template <class T> struct ActRows {
  T n[4][2], h[4];
  int count = 0;                          // 0: use the registered constant set
  const StateDyn<T>* sdyn = nullptr;      // per-environment safe STATE set of this step (NavigateQuadrotor, P6-P8; state model SM_DYNAMIC)
};

// P5, envs/navigate_seeker.py:415-477: largest (geo-mean) rectangle Z(0, U diag(len)) of actions whose one-step
// displacement keeps the seeker inside the state box and outside every obstacle.  Two variables: the maximum of
// len0 * len1 over a down-closed polygon lies on an edge midpoint or a vertex -- enumerated exactly.
template <class T>
SG_HD bool seeker_action_set(const T* s, const T* aux, const SafeConsts<T>& k, T (*U)[2], T* len) {
  U[0][0] = T(1); U[0][1] = T(0); U[1][0] = T(0); U[1][1] = T(1);
  T wall[2];
  for (int d = 0; d < 2; ++d) wall[d] = sg_abs(s[d]) - k.box_max[d];
  T min_dist = sg_sqrt(wall[0] * wall[0] + wall[1] * wall[1]);
  T dirs[SEEKER_MAXK][2], dist[SEEKER_MAXK];
  for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
    const T cx = aux[2 + 3 * i], cy = aux[3 + 3 * i], r = aux[4 + 3 * i];
    const T dx = cx - s[0], dy = cy - s[1];
    const T nrm = sg_sqrt(dx * dx + dy * dy);
    dirs[i][0] = dx / nrm; dirs[i][1] = dy / nrm;
    dist[i] = nrm - r;
    if (dist[i] < min_dist) {                 // generators: towards the closest obstacle and along its tangent
      U[0][0] = dirs[i][0]; U[1][0] = dirs[i][1];
      U[0][1] = dirs[i][1]; U[1][1] = -dirs[i][0];
      min_dist = dist[i];
    }
  }
  // rows  al * len0 + be * len1 <= ga  (all coefficients >= 0)
  T al[6 + SEEKER_MAXK], be[6 + SEEKER_MAXK], ga[6 + SEEKER_MAXK];
  int n = 0;
  for (int d = 0; d < 2; ++d) {               // state feasibility of s -+ |U diag(len)| 1
    al[n] = sg_abs(U[d][0]); be[n] = sg_abs(U[d][1]); ga[n] = s[d] - k.box_min[d]; ++n;
    al[n] = sg_abs(U[d][0]); be[n] = sg_abs(U[d][1]); ga[n] = k.box_max[d] - s[d]; ++n;
  }
  al[n] = T(1); be[n] = T(0); ga[n] = T(1); ++n;   // len <= 1 (action feasibility)
  al[n] = T(0); be[n] = T(1); ga[n] = T(1); ++n;
  for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
    al[n] = sg_abs(dirs[i][0] * U[0][0] + dirs[i][1] * U[1][0]);
    be[n] = sg_abs(dirs[i][0] * U[0][1] + dirs[i][1] * U[1][1]);
    ga[n] = dist[i]; ++n;
  }
  T best = T(-1), b0 = T(0), b1 = T(0);
  const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-11);
  auto feasible = [&](T x, T y) {
    if (x < T(0) || y < T(0)) return false;
    for (int r = 0; r < n; ++r)
      if (al[r] * x + be[r] * y > ga[r] + tol * (T(1) + sg_abs(ga[r]))) return false;
    return true;
  };
  for (int r = 0; r < n; ++r) {
    if (al[r] > T(0) && be[r] > T(0)) {       // maximiser of the product on this edge
      const T x = ga[r] / (T(2) * al[r]), y = ga[r] / (T(2) * be[r]);
      if (x * y > best && feasible(x, y)) { best = x * y; b0 = x; b1 = y; }
    }
    for (int q = r + 1; q < n; ++q) {
      const T det = al[r] * be[q] - al[q] * be[r];
      if (sg_abs(det) < T(1e-14)) continue;
      const T x = (ga[r] * be[q] - ga[q] * be[r]) / det, y = (al[r] * ga[q] - al[q] * ga[r]) / det;
      if (x * y > best && feasible(x, y)) { best = x * y; b0 = x; b1 = y; }
    }
  }
  len[0] = b0; len[1] = b1;
  return best > T(0);
}

template <class T> SG_HD void seeker_rows(const T (*U)[2], const T* len, ActRows<T>& R) {
  R.count = 4;
  for (int j = 0; j < 2; ++j) {               // |u_j . a| <= len_j  (U orthonormal)
    R.n[2 * j][0] = U[0][j]; R.n[2 * j][1] = U[1][j]; R.h[2 * j] = len[j];
    R.n[2 * j + 1][0] = -U[0][j]; R.n[2 * j + 1][1] = -U[1][j]; R.h[2 * j + 1] = len[j];
  }
}

// rows  n0 a0 + n1 a1 <= h   (S-typed, streamed; `idx` is a stable row number)
template <class TASK, class S, class T, class F>
SG_HD void for_each_row_m2(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, const ActRows<T>* dyn, F&& emit) {
  constexpr int NS = TASK::NS;
  int idx = 0;
  emit(C<S>(1.0), C<S>(0.0), C<S>(1.0), idx++);
  emit(C<S>(-1.0), C<S>(0.0), C<S>(1.0), idx++);
  emit(C<S>(0.0), C<S>(1.0), C<S>(1.0), idx++);
  emit(C<S>(0.0), C<S>(-1.0), C<S>(1.0), idx++);
  if (k.action_constrained) {
    if (dyn && dyn->count > 0) {
      for (int r = 0; r < dyn->count; ++r) emit(S(dyn->n[r][0]), S(dyn->n[r][1]), S(dyn->h[r]), idx++);
    } else {
      for (int r = 0; r < k.n_act_hp; ++r) emit(S(k.act_hp[r][0]), S(k.act_hp[r][1]), S(k.act_hp[r][2]), idx++);
    }
  }
  const bool dynamic_state = k.state_model == SM_DYNAMIC && dyn && dyn->sdyn;
  if (k.state_constrained && (k.state_model == SM_INVERTIBLE || dynamic_state)) {
    // constants registered once, or this step's per-environment set (built under no_grad in the reference: constants for autograd)
    const T* Ginv = dynamic_state ? dyn->sdyn->Ginv : k.Ginv;
    const T* cs_hat = dynamic_state ? dyn->sdyn->cs_hat : k.cs_hat;
    const T* rb = dynamic_state ? dyn->sdyn->rb : k.rb;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S e = S(cs_hat[i]), b0 = S(T(0)), b1 = S(T(0));
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const S gi = S(Ginv[i * NS + j]);
        e = e - gi * c_f[j];
        b0 = b0 + gi * B[j][0];
        b1 = b1 + gi * B[j][1];
      }
      emit(-b0, -b1, S(rb[i]) - e, idx++);
      emit(b0, b1, S(rb[i]) + e, idx++);
    }
  }
}

template <class T> struct Rows2 {
  T n0[MAXROWS], n1[MAXROWS], h[MAXROWS];
  int n;
};

template <class TASK, class S, class T>
SG_HD void primal_rows_m2(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, const ActRows<T>* dyn, Rows2<T>& R) {
  T cf_p[MAXN];
  T B_p[MAXN][MAXM];
#pragma unroll
  for (int i = 0; i < TASK::NS; ++i) { cf_p[i] = primal(c_f[i]); B_p[i][0] = primal(B[i][0]); B_p[i][1] = primal(B[i][1]); }
  R.n = 0;
  for_each_row_m2<TASK, T, T>(cf_p, B_p, k, dyn, [&](T a0, T a1, T hh, int) {
    if (R.n < MAXROWS) { R.n0[R.n] = a0; R.n1[R.n] = a1; R.h[R.n] = hh; ++R.n; }
  });
}

template <class T> SG_HD bool rows_feasible(const Rows2<T>& R, T x, T y, T tol) {
  for (int r = 0; r < R.n; ++r)
    if (R.n0[r] * x + R.n1[r] * y > R.h[r] + tol * (T(1) + sg_abs(R.h[r]))) return false;
  return true;
}

// P1 for m = 2: projection onto the polygon.  Active set on primal values, closed form in S.
template <class TASK, class S, class T>
SG_HD bool project_m2(const S* a, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, const ActRows<T>* dyn, S* out,
                      Rows2<T>& R) {
  primal_rows_m2<TASK>(c_f, B, k, dyn, R);
  const T ax = primal(a[0]), ay = primal(a[1]);
  const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-10);
  int act0 = -1, act1 = -1;
  bool found = false;
  if (rows_feasible(R, ax, ay, T(0))) {
    found = true;
  } else {
    T best = T(INFINITY);
    for (int r = 0; r < R.n; ++r) {            // single active row
      const T nn = R.n0[r] * R.n0[r] + R.n1[r] * R.n1[r];
      if (!(nn > T(0))) continue;
      const T viol = R.n0[r] * ax + R.n1[r] * ay - R.h[r];
      if (!(viol > T(0))) continue;
      const T px = ax - R.n0[r] * viol / nn, py = ay - R.n1[r] * viol / nn;
      const T d2 = viol * viol / nn;
      if (d2 < best && rows_feasible(R, px, py, tol)) { best = d2; act0 = r; act1 = -1; found = true; }
    }
    for (int r = 0; r < R.n; ++r)              // vertices
      for (int q = r + 1; q < R.n; ++q) {
        const T det = R.n0[r] * R.n1[q] - R.n0[q] * R.n1[r];
        if (sg_abs(det) < T(1e-12)) continue;
        const T px = (R.h[r] * R.n1[q] - R.h[q] * R.n1[r]) / det;
        const T py = (R.n0[r] * R.h[q] - R.n0[q] * R.h[r]) / det;
        const T d2 = (px - ax) * (px - ax) + (py - ay) * (py - ay);
        if (d2 < best * (T(1) - tol) && rows_feasible(R, px, py, tol)) { best = d2; act0 = r; act1 = q; found = true; }
      }
  }
  if (!found) { out[0] = C<S>(NAN); out[1] = C<S>(NAN); return false; }
  if (act0 < 0) { out[0] = a[0]; out[1] = a[1]; return true; }
  S r0[3], r1[3];
  for (int q = 0; q < 3; ++q) r0[q] = r1[q] = C<S>(0.0);
  for_each_row_m2<TASK, S, T>(c_f, B, k, dyn, [&](const S& n0, const S& n1, const S& hh, int idx) {
    if (idx == act0) { r0[0] = n0; r0[1] = n1; r0[2] = hh; }
    if (idx == act1) { r1[0] = n0; r1[1] = n1; r1[2] = hh; }
  });
  if (act1 < 0) {
    const S nn = r0[0] * r0[0] + r0[1] * r0[1];
    const S viol = (r0[0] * a[0] + r0[1] * a[1] - r0[2]) / nn;
    out[0] = a[0] - r0[0] * viol;
    out[1] = a[1] - r0[1] * viol;
  } else {
    const S det = r0[0] * r1[1] - r1[0] * r0[1];
    out[0] = (r0[2] * r1[1] - r1[2] * r0[1]) / det;
    out[1] = (r0[0] * r1[2] - r1[0] * r0[2]) / det;
  }
  return true;
}

// max t >= 0 with  p + t d  inside the polygon: the binding row is found on primal values, t is S-typed
template <class TASK, class S, class T>
SG_HD S ray_exit_m2(const S* p, const S* d, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k,
                    const ActRows<T>* dyn, const Rows2<T>& R) {
  const T px = primal(p[0]), py = primal(p[1]), dx = primal(d[0]), dy = primal(d[1]);
  T best = T(INFINITY);
  int arg = -1;
  for (int r = 0; r < R.n; ++r) {
    const T nd = R.n0[r] * dx + R.n1[r] * dy;
    if (!(nd > T(1e-12))) continue;
    T t = (R.h[r] - R.n0[r] * px - R.n1[r] * py) / nd;
    if (t < T(0)) t = T(0);
    if (t < best) { best = t; arg = r; }
  }
  if (arg < 0) return C<S>(INFINITY);
  S out = C<S>(0.0);
  for_each_row_m2<TASK, S, T>(c_f, B, k, dyn, [&](const S& n0, const S& n1, const S& hh, int idx) {
    if (idx == arg) out = (hh - n0 * p[0] - n1 * p[1]) / (n0 * d[0] + n1 * d[1]);
  });
  if (primal(out) < T(0)) out = C<S>(0.0);
  return out;
}

template <class S> SG_HD S norm2(const S& x, const S& y) {
  using T = typename Traits<S>::base;
  const S q = x * x + y * y;
  if (primal(q) == T(0)) return C<S>(0.0);     // torch vector_norm: zero subgradient at the origin
  return sg_sqrt(q);
}

template <class S> SG_HD S unit_box_dist_m2(const S* c, const S* d) {   // ray_mask.py:243-266
  using T = typename Traits<S>::base;
  S best = C<S>(INFINITY);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (primal(d[q]) == T(0)) continue;
    const S hi = (C<S>(1.0) - c[q]) / d[q], lo = (C<S>(-1.0) - c[q]) / d[q];
    if (!(primal(hi) < T(0)) && primal(hi) < primal(best)) best = hi;
    if (!(primal(lo) < T(0)) && primal(lo) < primal(best)) best = lo;
  }
  return best;
}

// ================================================================================================
// P2 for m = 2 (ray_mask.py:145-191) with an invertible safe-state generator (ManageHousehold):
//   max geo_mean(l),  l in R^4_+,  centre c in R^2,  G = D diag(l)  (D: 2 x 4 random unit columns)
//   rows  alpha_r . c + omega_r . l <= gamma_r,  omega_r >= 0:
//     box    +-c_k + sum_j |D_kj| l_j <= 1                                    (ray_mask.py:157-158)
//     state  +-(e_i - b_i . c) + sum_j |(G_s^-1 B0 D)_ij| l_j <= rb_i          (enc(Z(c+, [B0 G, G_w]) in Z(c_s, G_s)))
// A strictly concave maximisation (sum of logs) over 6 variables and 4 + 2n rows: primal-dual interior point in
// DOUBLE whatever the I/O type (fixed small dense algebra, 6 x 6 Cholesky in registers / local memory), then one
// Newton step of the perturbed KKT system with the rows evaluated in the dual-number type: its derivative part is
// the implicit-function derivative of the optimum w.r.t. (c_f, B), its primal part a last correction.
// ================================================================================================
constexpr int P2_NV = 6, P2_NR = 4 + 2 * MAXN;

template <class W> struct P2Rows { W a[P2_NR][2]; double om[P2_NR][4]; W g[P2_NR]; int n; };

template <class TASK, class SW, class S, class T>
SG_HD void p2_rows_m2(const S* c_f, const S (*B)[MAXM], const T* dirs, const SafeConsts<T>& k, P2Rows<SW>& R) {
  constexpr int NS = TASK::NS;
  int r = 0;
  for (int q = 0; q < 2; ++q)
    for (int sgn = 0; sgn < 2; ++sgn) {
      R.a[r][0] = SW(q == 0 ? (sgn ? -1.0 : 1.0) : 0.0);
      R.a[r][1] = SW(q == 1 ? (sgn ? -1.0 : 1.0) : 0.0);
      for (int j = 0; j < 4; ++j) R.om[r][j] = fabs((double)dirs[q * 4 + j]);
      R.g[r] = SW(1.0);
      ++r;
    }
  for (int i = 0; i < NS; ++i) {
    SW e = SW((double)k.cs_hat[i]), b0 = SW(0.0), b1 = SW(0.0);
    for (int j = 0; j < NS; ++j) {
      const double gi = (double)k.Ginv[i * NS + j];
      e = e - SW(gi) * widen(c_f[j]);
      b0 = b0 + SW(gi) * widen(B[j][0]);
      b1 = b1 + SW(gi) * widen(B[j][1]);
    }
    double w[4];
    for (int j = 0; j < 4; ++j)
      w[j] = fabs((double)k.B0hat[i * MAXM + 0] * (double)dirs[j] + (double)k.B0hat[i * MAXM + 1] * (double)dirs[4 + j]);
    // (e - b.c) + w.l <= rb   and   -(e - b.c) + w.l <= rb
    R.a[r][0] = -b0; R.a[r][1] = -b1; R.g[r] = SW((double)k.rb[i]) - e;
    for (int j = 0; j < 4; ++j) R.om[r][j] = w[j];
    ++r;
    R.a[r][0] = b0; R.a[r][1] = b1; R.g[r] = SW((double)k.rb[i]) + e;
    for (int j = 0; j < 4; ++j) R.om[r][j] = w[j];
    ++r;
  }
  R.n = r;
}

// Cholesky of a 6 x 6 SPD matrix (lower factor in place).  A pivot that cancelled to (almost) nothing is a direction
// the active rows leave free (the optimal centre of P2 is not unique there): it is frozen (huge pivot -> zero step)
// instead of dividing by rounding noise.  Returns false only for NaNs.
SG_HD bool chol6(double (*K)[P2_NV]) {
  for (int j = 0; j < P2_NV; ++j) {
    const double orig = K[j][j];
    double d = orig;
    for (int p = 0; p < j; ++p) d -= K[j][p] * K[j][p];
    if (!(d == d)) return false;
    // (threshold: rounding noise of the cancellation is ~1e-16 orig; genuine pivots are the log-objective curvature
    //  1 / l^2 >= 1 against barrier entries ~1e10 at the final mu, i.e. >= ~1e-10 orig)
    if (!(d > 1e-12 * orig)) d = 1e60;
    d = sqrt(d);
    K[j][j] = d;
    for (int i = j + 1; i < P2_NV; ++i) {
      double v = K[i][j];
      for (int p = 0; p < j; ++p) v -= K[i][p] * K[j][p];
      K[i][j] = v / d;
    }
  }
  return true;
}
template <class W> SG_HD void chol6_solve(const double (*L)[P2_NV], W* x) {      // x := K^-1 x
  for (int i = 0; i < P2_NV; ++i) {
    W v = x[i];
    for (int p = 0; p < i; ++p) v = v - W(L[i][p]) * x[p];
    x[i] = v / W(L[i][i]);
  }
  for (int i = P2_NV - 1; i >= 0; --i) {
    W v = x[i];
    for (int p = i + 1; p < P2_NV; ++p) v = v - W(L[p][i]) * x[p];
    x[i] = v / W(L[i][i]);
  }
}

template <class TASK, class S, class T>
SG_HD bool p2_expand_m2(const S* c_f, const S (*B)[MAXM], const T* dirs, const SafeConsts<T>& k, S* c_out, S* len_out) {
  using SW = typename Widen<S>::type;
  // ---- primal interior point in double ------------------------------------------------------------------
  P2Rows<double> R;
  {
    T cf_p[MAXN];
    T B_p[MAXN][MAXM];
    for (int i = 0; i < TASK::NS; ++i) { cf_p[i] = primal(c_f[i]); B_p[i][0] = primal(B[i][0]); B_p[i][1] = primal(B[i][1]); }
    p2_rows_m2<TASK, double, T, T>(cf_p, B_p, dirs, k, R);
  }
  const int nr = R.n;
  double x[P2_NV] = {0.0, 0.0, 0.05, 0.05, 0.05, 0.05};
  double s[P2_NR], lam[P2_NR];
  for (int r = 0; r < nr; ++r) {
    double ax = R.a[r][0] * x[0] + R.a[r][1] * x[1];
    for (int j = 0; j < 4; ++j) ax += R.om[r][j] * x[2 + j];
    s[r] = R.g[r] - ax;
    if (s[r] < 0.1) s[r] = 0.1;
    lam[r] = 1.0 / s[r];
  }
  double L[P2_NV][P2_NV];
  double mu = 1.0;
  bool ok = false;
  for (int it = 0; it < 60; ++it) {
    double rp[P2_NR], rd[P2_NV] = {0.0, 0.0, -1.0 / x[2], -1.0 / x[3], -1.0 / x[4], -1.0 / x[5]};
    double gap = 0.0, rp_max = 0.0;
    for (int r = 0; r < nr; ++r) {
      double ax = R.a[r][0] * x[0] + R.a[r][1] * x[1];
      for (int j = 0; j < 4; ++j) ax += R.om[r][j] * x[2 + j];
      rp[r] = ax + s[r] - R.g[r];
      rp_max = fmax(rp_max, fabs(rp[r]));
      rd[0] += lam[r] * R.a[r][0]; rd[1] += lam[r] * R.a[r][1];
      for (int j = 0; j < 4; ++j) rd[2 + j] += lam[r] * R.om[r][j];
      gap += lam[r] * s[r];
    }
    mu = gap / nr;
    double rd_max = 0.0;
    for (int i = 0; i < P2_NV; ++i) rd_max = fmax(rd_max, fabs(rd[i]));
    // converged on the central path at mu ~ 1e-10: the optimum is reached to ~1e-10 and the KKT matrix of the
    // derivative step below is still well inside double precision (entries ~ 1 / mu)
    if (rp_max < 1e-9 && rd_max < 1e-6 && mu < 2e-10) { ok = true; break; }
    const double target = fmax(0.1 * mu, 5e-11);
    // K = H + A^T diag(lam / s) A ,  rhs = -rd - A^T [(lam rp - (lam s - target)) / s]
    double rhs[P2_NV];
    for (int i = 0; i < P2_NV; ++i) { rhs[i] = -rd[i]; for (int j = 0; j < P2_NV; ++j) L[i][j] = 0.0; }
    for (int j = 0; j < 4; ++j) L[2 + j][2 + j] = 1.0 / (x[2 + j] * x[2 + j]);
    for (int r = 0; r < nr; ++r) {
      double row[P2_NV] = {R.a[r][0], R.a[r][1], R.om[r][0], R.om[r][1], R.om[r][2], R.om[r][3]};
      const double w = lam[r] / s[r];
      const double t = (lam[r] * rp[r] - (lam[r] * s[r] - target)) / s[r];
      for (int i = 0; i < P2_NV; ++i) {
        rhs[i] -= row[i] * t;
        for (int j = 0; j <= i; ++j) L[i][j] += w * row[i] * row[j];
      }
    }
    if (!chol6(L)) break;
    chol6_solve<double>(L, rhs);                 // rhs := dx
    double alpha = 1.0, ds[P2_NR], dl[P2_NR];
    for (int j = 0; j < 4; ++j)
      if (rhs[2 + j] < 0.0) alpha = fmin(alpha, -0.995 * x[2 + j] / rhs[2 + j]);
    for (int r = 0; r < nr; ++r) {
      double adx = R.a[r][0] * rhs[0] + R.a[r][1] * rhs[1];
      for (int j = 0; j < 4; ++j) adx += R.om[r][j] * rhs[2 + j];
      ds[r] = -rp[r] - adx;
      dl[r] = (-(lam[r] * s[r] - target) - lam[r] * ds[r]) / s[r];
      if (ds[r] < 0.0) alpha = fmin(alpha, -0.995 * s[r] / ds[r]);
      if (dl[r] < 0.0) alpha = fmin(alpha, -0.995 * lam[r] / dl[r]);
    }
    for (int i = 0; i < P2_NV; ++i) x[i] += alpha * rhs[i];
    for (int r = 0; r < nr; ++r) { s[r] += alpha * ds[r]; lam[r] += alpha * dl[r]; }
  }
  if (!ok) {
    c_out[0] = C<S>(NAN); c_out[1] = C<S>(NAN);
    for (int j = 0; j < 4; ++j) len_out[j] = C<S>(NAN);
    return false;
  }
  // ---- one Newton step of the same system with S-typed rows (implicit-function derivative + last correction) ---
  P2Rows<SW> RS;
  p2_rows_m2<TASK, SW, S, T>(c_f, B, dirs, k, RS);
  SW rhs[P2_NV];
  for (int i = 0; i < P2_NV; ++i) { rhs[i] = SW(0.0); for (int j = 0; j < P2_NV; ++j) L[i][j] = 0.0; }
  for (int j = 0; j < 4; ++j) { rhs[2 + j] = SW(1.0 / x[2 + j]); L[2 + j][2 + j] = 1.0 / (x[2 + j] * x[2 + j]); }
  for (int r = 0; r < nr; ++r) {
    const SW rowS[P2_NV] = {RS.a[r][0], RS.a[r][1], SW(RS.om[r][0]), SW(RS.om[r][1]), SW(RS.om[r][2]), SW(RS.om[r][3])};
    const double row[P2_NV] = {R.a[r][0], R.a[r][1], R.om[r][0], R.om[r][1], R.om[r][2], R.om[r][3]};
    SW ax = rowS[0] * SW(x[0]) + rowS[1] * SW(x[1]);
    for (int j = 0; j < 4; ++j) ax = ax + rowS[2 + j] * SW(x[2 + j]);
    const SW rp = ax + SW(s[r]) - RS.g[r];
    const double w = lam[r] / s[r];
    // complementarity target of row r := its current product lam_r s_r, so that the system being differentiated is
    // satisfied EXACTLY at the converged point (a residual lam s - mu != 0 would leak into the derivative part)
    const SW t = (SW(lam[r]) * rp) / SW(s[r]);
    for (int i = 0; i < P2_NV; ++i) {
      rhs[i] = rhs[i] - rowS[i] * SW(lam[r]) - rowS[i] * t;     // -(rd + A^T t):  rd = grad g + A^T lam
      for (int j = 0; j <= i; ++j) L[i][j] += w * row[i] * row[j];
    }
  }
  if (!chol6(L)) { c_out[0] = C<S>(NAN); c_out[1] = C<S>(NAN); for (int j = 0; j < 4; ++j) len_out[j] = C<S>(NAN); return false; }
  chol6_solve<SW>(L, rhs);
  for (int q = 0; q < 2; ++q) narrow(SW(x[q]) + rhs[q], c_out[q]);
  for (int j = 0; j < 4; ++j) narrow(SW(x[2 + j]) + rhs[2 + j], len_out[j]);
  return true;
}

// P3 (ray_mask.py:215-240) on the 2-D zonotope Z(c, D diag(l)) along `dir`:  t* = min_e  h(n_e) / |n_e . dir|,
// n_e = perp(g_e), h(n) = sum_j |n . g_j|
template <class S, class T>
SG_HD S ray_dist_zonotope_m2(const S* dir, const T* dirs, const S* len) {
  S g[4][2];
  for (int j = 0; j < 4; ++j) { g[j][0] = S(dirs[j]) * len[j]; g[j][1] = S(dirs[4 + j]) * len[j]; }
  T best = T(INFINITY);
  int arg = -1;
  for (int e = 0; e < 4; ++e) {
    const T nx = -primal(g[e][1]), ny = primal(g[e][0]);
    const T nd = sg_abs(nx * primal(dir[0]) + ny * primal(dir[1]));
    if (!(nd > T(1e-30))) continue;
    T h = T(0);
    for (int j = 0; j < 4; ++j) h += sg_abs(nx * primal(g[j][0]) + ny * primal(g[j][1]));
    if (h / nd < best) { best = h / nd; arg = e; }
  }
  if (arg < 0) return C<S>(INFINITY);
  const S nx = -g[arg][1], ny = g[arg][0];
  S h = C<S>(0.0);
  for (int j = 0; j < 4; ++j) h = h + sg_abs(nx * g[j][0] + ny * g[j][1]);
  return h / sg_abs(nx * dir[0] + ny * dir[1]);
}

// safeguard(a) for m = 2.  `dirs`: the 2 x 4 random unit directions of RM-zonotopic (ray_mask.py:184-185).
template <class TASK, class S, class T>
SG_HD bool safeguard_m2(const S* a, const S* c_f, const S (*B)[MAXM], const T* dirs, const SafeConsts<T>& k,
                        const ActRows<T>* dyn, S* out) {
  Rows2<T> R;
  if (k.state_constrained && k.state_model == SM_UNREDUCED) {
    out[0] = C<S>(NAN); out[1] = C<S>(NAN);
    return false;
  }
  if (k.sg_kind == SG_BOUNDARY_PROJECTION) return project_m2<TASK>(a, c_f, B, k, dyn, out, R);
  S center[2], safe_dist, feas_dist;
  bool feasible = true;
  if (k.state_constrained && k.rm_zonotopic) {
    // zonotopic approximation: P2 (four free lengths) then P3 along the ray from the centre (ray_mask.py:63-89)
    if (k.state_model != SM_INVERTIBLE || k.action_constrained || !dirs) {      // (incl. SM_DYNAMIC: P2 on a per-step set is not built)
      out[0] = C<S>(NAN); out[1] = C<S>(NAN);
      return false;
    }
    S len[4];
    feasible = p2_expand_m2<TASK>(c_f, B, dirs, k, center, len);
    const S dx = a[0] - center[0], dy = a[1] - center[1];
    const S nrm = norm2(dx, dy);
    S dir[2] = {dx / (nrm + C<S>(1e-8)), dy / (nrm + C<S>(1e-8))};
    safe_dist = ray_dist_zonotope_m2<S, T>(dir, dirs, len);
    feas_dist = unit_box_dist_m2(center, dir);
  } else if (k.state_constrained) {
    S start[2];
    feasible = project_m2<TASK>(a, c_f, B, k, dyn, start, R);
    const S gx = start[0] - a[0], gy = start[1] - a[1];
    const S gap = norm2(gx, gy);
    if (primal(gap) < T(1e-8)) {
      center[0] = start[0]; center[1] = start[1]; safe_dist = C<S>(1.0); feas_dist = C<S>(1.0);
    } else {
      S dir[2] = {gx / gap, gy / gap};
      const S t = ray_exit_m2<TASK>(start, dir, c_f, B, k, dyn, R);
      safe_dist = t * C<S>(0.5);
      center[0] = start[0] + dir[0] * safe_dist;
      center[1] = start[1] + dir[1] * safe_dist;
      S nd[2] = {-dir[0], -dir[1]};
      feas_dist = unit_box_dist_m2(center, nd);
    }
  } else {
    // action-constrained only (ray_mask.py:77-78): centre of the safe action set, P3 along the ray
    primal_rows_m2<TASK>(c_f, B, k, dyn, R);
    center[0] = S(k.c_a[0]); center[1] = S(k.c_a[1]);
    const S dx = a[0] - center[0], dy = a[1] - center[1];
    const S nrm = norm2(dx, dy);
    S dir[2] = {dx / (nrm + C<S>(1e-8)), dy / (nrm + C<S>(1e-8))};
    // P3 ignores the feasibility box: only the zonotope rows bound t (rows 4.. are the zonotope)
    Rows2<T> Z;
    Z.n = 0;
    for (int r = 4; r < R.n; ++r) { Z.n0[Z.n] = R.n0[r]; Z.n1[Z.n] = R.n1[r]; Z.h[Z.n] = R.h[r]; ++Z.n; }
    const T px = primal(center[0]), py = primal(center[1]), ddx = primal(dir[0]), ddy = primal(dir[1]);
    T best = T(INFINITY);
    int arg = -1;
    for (int r = 0; r < Z.n; ++r) {
      const T nd = Z.n0[r] * ddx + Z.n1[r] * ddy;
      if (!(nd > T(1e-12))) continue;
      const T t = (Z.h[r] - Z.n0[r] * px - Z.n1[r] * py) / nd;
      if (t < best) { best = t; arg = r; }
    }
    safe_dist = C<S>(INFINITY);
    if (arg >= 0) {
      const S n0 = S(Z.n0[arg]), n1 = S(Z.n1[arg]), hh = S(Z.h[arg]);
      safe_dist = (hh - n0 * center[0] - n1 * center[1]) / (n0 * dir[0] + n1 * dir[1]);
    }
    feas_dist = unit_box_dist_m2(center, dir);
  }
  const S dx = a[0] - center[0], dy = a[1] - center[1];
  const S dist = norm2(dx, dy);
  if (primal(dist) < T(1e-8)) { out[0] = center[0]; out[1] = center[1]; return feasible; }
  const S scale = radial_map(dist, safe_dist, feas_dist, k) / (dist + C<S>(1e-8));
  out[0] = center[0] + dx * scale;
  out[1] = center[1] + dy * scale;
  return feasible;
}

// ---- verdicts (sets/zonotope.py:143-239 restricted to the registered models) ----------------------
// a projected action sits ON the boundary up to rounding, so the verdicts use a relative slack well
// above rounding and well below any real violation
template <class T> SG_HD T verdict_tol();
template <> SG_HD float verdict_tol<float>() { return 1e-4f; }
template <> SG_HD double verdict_tol<double>() { return 1e-9; }
template <class TASK, class T> SG_HD bool action_in_set(const T* a, const SafeConsts<T>& k, const ActRows<T>* dyn = nullptr) {
  const T tol = verdict_tol<T>();
  if (dyn && dyn->count > 0) {
    for (int r = 0; r < dyn->count; ++r)
      if (dyn->n[r][0] * a[0] + dyn->n[r][1] * a[1] > dyn->h[r] + tol * (T(1) + sg_abs(dyn->h[r]))) return false;
    return true;
  }
  if (TASK::NA == 1) return a[0] >= k.a_lo[0] - tol * (T(1) + sg_abs(k.a_lo[0])) && a[0] <= k.a_hi[0] + tol * (T(1) + sg_abs(k.a_hi[0]));
  for (int r = 0; r < k.n_act_hp; ++r)
    if (k.act_hp[r][0] * a[0] + k.act_hp[r][1] * a[1] > k.act_hp[r][2] + tol * (T(1) + sg_abs(k.act_hp[r][2]))) return false;
  return true;
}

template <class TASK, class T>
SG_HD bool next_state_in_set(const T* a, const T* c_f, const T (*B)[MAXM], const SafeConsts<T>& k, const ActRows<T>* dyn = nullptr) {
  constexpr int NS = TASK::NS, NA = TASK::NA;
  const T tol = verdict_tol<T>();
  if (k.state_model == SM_UNREDUCED || k.state_model == SM_LIFTED) return false;      // verdict not available for these models
  const bool dynamic_state = k.state_model == SM_DYNAMIC && dyn && dyn->sdyn;
  if (k.state_model == SM_DYNAMIC && (!dynamic_state || dyn->sdyn->degenerate)) return false;
  if (k.state_model == SM_INVERTIBLE || dynamic_state) {
    const T* Ginv = dynamic_state ? dyn->sdyn->Ginv : k.Ginv;
    const T* cs_hat = dynamic_state ? dyn->sdyn->cs_hat : k.cs_hat;
    const T* rb = dynamic_state ? dyn->sdyn->rb : k.rb;
    for (int i = 0; i < NS; ++i) {
      T e = cs_hat[i];
      for (int j = 0; j < NS; ++j) {
        T nx = c_f[j];
        for (int q = 0; q < NA; ++q) nx += B[j][q] * a[q];
        e -= Ginv[i * NS + j] * nx;
      }
      if (sg_abs(e) > rb[i] + tol * (T(1) + sg_abs(rb[i]))) return false;
    }
    return true;
  }
  for (int r = 0; r < k.n_k_hp; ++r) {
    T e = T(0);
    for (int j = 0; j < (NS < 2 ? NS : 2); ++j) {
      T nx = c_f[j];
      for (int q = 0; q < NA; ++q) nx += B[j][q] * a[q];
      e += k.k_hp[r][j] * (k.c_s[j] - nx);
    }
    if (e > T(1) + tol) return false;
  }
  return true;
}

}  // namespace sgbpo
