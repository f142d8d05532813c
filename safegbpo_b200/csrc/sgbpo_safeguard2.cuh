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

// per-environment safe action set computed this step (NavigateSeeker, P5): rows n . a <= h, constants for autograd
template <class T> struct ActRows {
  T n[4][2], h[4];
  int count;        // 0: use the registered constant set
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
  if (k.state_constrained && k.state_model == SM_INVERTIBLE) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S e = S(k.cs_hat[i]), b0 = S(T(0)), b1 = S(T(0));
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const S gi = S(k.Ginv[i * NS + j]);
        e = e - gi * c_f[j];
        b0 = b0 + gi * B[j][0];
        b1 = b1 + gi * B[j][1];
      }
      emit(-b0, -b1, S(k.rb[i]) - e, idx++);
      emit(b0, b1, S(k.rb[i]) + e, idx++);
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
    // P2 with four free lengths (geo-mean cone program): not reduced to closed form yet.
    (void)dirs;
    out[0] = C<S>(NAN); out[1] = C<S>(NAN);
    return false;
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
SG_HD bool next_state_in_set(const T* a, const T* c_f, const T (*B)[MAXM], const SafeConsts<T>& k) {
  constexpr int NS = TASK::NS, NA = TASK::NA;
  const T tol = verdict_tol<T>();
  if (k.state_model == SM_UNREDUCED) return false;      // verdict not available for this model
  if (k.state_model == SM_INVERTIBLE) {
    for (int i = 0; i < NS; ++i) {
      T e = k.cs_hat[i];
      for (int j = 0; j < NS; ++j) {
        T nx = c_f[j];
        for (int q = 0; q < NA; ++q) nx += B[j][q] * a[q];
        e -= k.Ginv[i * NS + j] * nx;
      }
      if (sg_abs(e) > k.rb[i] + tol * (T(1) + sg_abs(k.rb[i]))) return false;
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
