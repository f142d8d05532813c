// NavigateQuadrotor's per-step safe state set (reference: src/envs/navigate_quadrotor.py:436-779), forward only (the
// reference builds it under torch.no_grad()).  Three kinds of two-variable programs, each solved exactly in registers:
//
//   P6  compute_support_point (:542-582)   max (d^T G) beta  s.t.  box_min <= c + G beta <= box_max, |beta|_inf <= 1
//       an LP over a polygon of <= 16 half-planes: the optimum is a vertex (enumerated); used by feasible_center (:491-517)
//       and collision_free_center (:519-540).
//   P7  compute_position_generator (:585-654) and
//   P8  compute_velocity_generator (:657-779): max geo_mean(len) over a down-closed polygon  al len0 + be len1 <= ga with
//       non-negative coefficients -- the maximiser of len0 * len1 is an edge midpoint or a vertex (same enumeration as
//       NavigateSeeker's P5, sgbpo_safeguard2.cuh::seeker_action_set).
//
// The resulting zonotope Z(c_s, [pos | vel | roll]) has a block generator (positions rows 0-1, velocities rows 3-4, the two
// constant roll generators on rows 2 and 5), so G_s^-1 is two 2 x 2 orthonormal-times-diagonal inverses and two scalars and
// the containment rows of the safeguard are those of the invertible model with per-environment constants (StateDyn).
// Where the reference falls back to its all-rows 1e-2 generator blocks (unsafe position / velocity: the generator becomes
// rank deficient and its programs infeasible) the set is reported as degenerate instead.
#pragma once
#include "sgbpo_safeguard2.cuh"

namespace sgbpo {

// max of x * y over { x, y >= 0, al_r x + be_r y <= ga_r } (al, be >= 0): edge midpoints and vertices
template <class T> SG_HD bool max_product_2d(const T* al, const T* be, const T* ga, int n, T& bx, T& by) {
  T best = T(-1);
  bx = T(0); by = T(0);
  const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-11);
  auto feasible = [&](T x, T y) {
    if (x < T(0) || y < T(0)) return false;
    for (int r = 0; r < n; ++r)
      if (al[r] * x + be[r] * y > ga[r] + tol * (T(1) + sg_abs(ga[r]))) return false;
    return true;
  };
  for (int r = 0; r < n; ++r) {
    if (al[r] > T(0) && be[r] > T(0)) {
      const T x = ga[r] / (T(2) * al[r]), y = ga[r] / (T(2) * be[r]);
      if (x * y > best && feasible(x, y)) { best = x * y; bx = x; by = y; }
    }
    for (int q = r + 1; q < n; ++q) {
      const T det = al[r] * be[q] - al[q] * be[r];
      if (sg_abs(det) < T(1e-14)) continue;
      const T x = (ga[r] * be[q] - ga[q] * be[r]) / det, y = (al[r] * ga[q] - al[q] * ga[r]) / det;
      if (x * y > best && feasible(x, y)) { best = x * y; bx = x; by = y; }
    }
  }
  return best > T(0);
}

// P6: support point of Z(c, G) (G: 6 x 2) in direction d inside the state box.  Returns false when no beta is feasible.
template <class T>
SG_HD bool navquad_support_point(const T* c, const T (*G)[MAXM], const T* d, const SafeConsts<T>& k, T* out) {
  T n0[16], n1[16], h[16];
  int n = 0;
  for (int i = 0; i < 6; ++i) {
    n0[n] = G[i][0]; n1[n] = G[i][1]; h[n] = k.box_max[i] - c[i]; ++n;        //  (c + G beta)_i <= max_i
    n0[n] = -G[i][0]; n1[n] = -G[i][1]; h[n] = c[i] - k.box_min[i]; ++n;      // -(c + G beta)_i <= -min_i
  }
  n0[n] = T(1); n1[n] = T(0); h[n] = T(1); ++n;
  n0[n] = T(-1); n1[n] = T(0); h[n] = T(1); ++n;
  n0[n] = T(0); n1[n] = T(1); h[n] = T(1); ++n;
  n0[n] = T(0); n1[n] = T(-1); h[n] = T(1); ++n;
  T o0 = T(0), o1 = T(0);
  for (int i = 0; i < 6; ++i) { o0 += d[i] * G[i][0]; o1 += d[i] * G[i][1]; }
  const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-10);
  auto feasible = [&](T x, T y) {
    for (int r = 0; r < n; ++r)
      if (n0[r] * x + n1[r] * y > h[r] + tol * (T(1) + sg_abs(h[r]))) return false;
    return true;
  };
  bool found = false;
  T best = T(0), bx = T(0), by = T(0);
  for (int r = 0; r < n; ++r)
    for (int q = r + 1; q < n; ++q) {
      const T det = n0[r] * n1[q] - n0[q] * n1[r];
      if (sg_abs(det) < T(1e-12)) continue;
      const T x = (h[r] * n1[q] - h[q] * n1[r]) / det, y = (n0[r] * h[q] - n0[q] * h[r]) / det;
      if (!feasible(x, y)) continue;
      const T v = o0 * x + o1 * y;
      if (!found || v > best) { found = true; best = v; bx = x; by = y; }
    }
  if (!found) return false;
  for (int i = 0; i < 6; ++i) out[i] = c[i] + G[i][0] * bx + G[i][1] * by;
  return true;
}

template <class T> SG_HD int navquad_last_hit(const T* p, const T* aux, const SafeConsts<T>& k) {      // collision_check + obstacle_mask
  int hit = -1;
  for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
    const T dx = p[0] - aux[2 + 3 * i], dy = p[1] - aux[3 + 3 * i];
    if (sg_sqrt(dx * dx + dy * dy) <= aux[4 + 3 * i]) hit = i;
  }
  return hit;
}

// safe_state_set (:436-460).  c_f, B: the one-step reachable set Z(c_f, B I_2); G_w (n x 2) = k.Gam_p rows 0..5.
template <class T>
SG_HD void navquad_safe_state_set(const T* c_f, const T (*B)[MAXM], const T* aux, const SafeConsts<T>& k, StateDyn<T>& out) {
  constexpr T DT = T(0.05);
  const T PI = T(3.14159265358979323846);
  const T max_acc[2] = {T((6.0 / 7.0 + 9.81) * 0.19509032201612825), T(6.0 / 7.0)};      // (T_mag + g) sin(pi/16), T_mag (:83-86)
  out.degenerate = false;
  // ---- feasible_center (:491-517) --------------------------------------------------------------------------------------
  bool low[6], high[6], inside = true;
  for (int i = 0; i < 6; ++i) {
    low[i] = k.box_min[i] > c_f[i];
    high[i] = k.box_max[i] < c_f[i];
    inside = inside && !low[i] && !high[i];
  }
  low[2] = low[2] || low[3]; high[2] = high[2] || high[3];
  low[5] = low[5] || low[3]; high[5] = high[5] || high[3];
  T center[6];
  for (int i = 0; i < 6; ++i) center[i] = c_f[i];
  if (!inside) {
    T dir[6];
    for (int i = 0; i < 6; ++i) dir[i] = high[i] ? -k.box_max[i] : (low[i] ? -k.box_min[i] : T(1));
    if (!navquad_support_point<T>(c_f, B, dir, k, center)) out.degenerate = true;
  }
  // ---- collision_free_center (:519-540) ---------------------------------------------------------------------------------
  {
    const int hit = navquad_last_hit<T>(center, aux, k);
    if (hit >= 0 && !out.degenerate) {
      T dir[6] = {center[0] - aux[2 + 3 * hit], center[1] - aux[3 + 3 * hit], T(0), T(0), T(0), T(0)};
      T shifted[6];
      if (navquad_support_point<T>(c_f, B, dir, k, shifted)) {
        for (int i = 0; i < 6; ++i) center[i] = shifted[i];
      } else {
        out.degenerate = true;
      }
    }
  }
  // ---- safe_center masks (:478-483) ---------------------------------------------------------------------------------------
  const bool safe_pos = !(k.box_min[0] > c_f[0] || k.box_min[1] > c_f[1] || k.box_max[0] < c_f[0] || k.box_max[1] < c_f[1] ||
                          navquad_last_hit<T>(center, aux, k) >= 0);
  const bool safe_vel = !(k.box_min[3] > c_f[3] || k.box_min[4] > c_f[4] || k.box_max[3] < c_f[3] || k.box_max[4] < c_f[4]);
  if (!safe_pos || !safe_vel) out.degenerate = true;          // the reference's all-rows 1e-2 fallback blocks (:447-452)
  // ---- P7: position generators (:585-654) -----------------------------------------------------------------------------------
  T Up[2][2] = {{T(1), T(0)}, {T(0), T(1)}}, lp[2] = {T(1e-2), T(1e-2)};
  {
    const T w0 = sg_abs(center[0]) - k.box_max[0], w1 = sg_abs(center[1]) - k.box_max[1];
    T min_dist = sg_sqrt(w0 * w0 + w1 * w1);
    T dirs[SEEKER_MAXK][2], dist[SEEKER_MAXK];
    for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
      const T dx = aux[2 + 3 * i] - center[0], dy = aux[3 + 3 * i] - center[1];
      const T nrm = sg_sqrt(dx * dx + dy * dy);
      dirs[i][0] = dx / nrm; dirs[i][1] = dy / nrm;
      dist[i] = nrm - aux[4 + 3 * i];
      if (dist[i] < min_dist) {
        Up[0][0] = dirs[i][0]; Up[1][0] = dirs[i][1];
        Up[0][1] = dirs[i][1]; Up[1][1] = -dirs[i][0];
        min_dist = dist[i];
      }
    }
    T al[4 + SEEKER_MAXK], be[4 + SEEKER_MAXK], ga[4 + SEEKER_MAXK];
    int n = 0;
    for (int d = 0; d < 2; ++d) {
      al[n] = sg_abs(Up[d][0]); be[n] = sg_abs(Up[d][1]); ga[n] = center[d] - k.box_min[d]; ++n;
      al[n] = sg_abs(Up[d][0]); be[n] = sg_abs(Up[d][1]); ga[n] = k.box_max[d] - center[d]; ++n;
    }
    for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
      al[n] = sg_abs(dirs[i][0] * Up[0][0] + dirs[i][1] * Up[1][0]);
      be[n] = sg_abs(dirs[i][0] * Up[0][1] + dirs[i][1] * Up[1][1]);
      ga[n] = dist[i]; ++n;
    }
    T x, y;
    if (max_product_2d<T>(al, be, ga, n, x, y)) { lp[0] = x; lp[1] = y; }      // else: the 1e-2 fallback length (:650-651)
  }
  // ---- P8: velocity generators (:657-779) -----------------------------------------------------------------------------------
  T Uv[2][2] = {{T(1), T(0)}, {T(0), T(1)}}, lv[2] = {T(1e-2), T(1e-2)};
  {
    T lower_d[2], upper_d[2], lower_w[2], upper_w[2];
    for (int d = 0; d < 2; ++d) {
      const T hw = sg_abs(B[d][0]) + sg_abs(B[d][1]);                          // reachable_set.box() half width (:706-708)
      lower_d[d] = sg_abs(k.box_min[d] - (c_f[d] - hw));
      upper_d[d] = sg_abs(k.box_max[d] - (c_f[d] + hw));
    }
    lower_d[0] += T(5) * DT * center[3];
    upper_d[0] -= T(5) * DT * center[3];
    for (int d = 0; d < 2; ++d) {
      if (lower_d[d] < T(0)) lower_d[d] = T(1e-2);
      if (upper_d[d] < T(0)) upper_d[d] = T(1e-2);
      lower_w[d] = -sg_sqrt(lower_d[d] * T(2) * max_acc[d]) - center[3 + d];
      if (lower_w[d] > T(0)) lower_w[d] = T(0);
      upper_w[d] = sg_sqrt(upper_d[d] * T(2) * max_acc[d]) - center[3 + d];
      if (upper_w[d] < T(0)) upper_w[d] = T(0);
    }
    const T v0 = sg_abs(center[3]) - k.box_max[3], v1 = sg_abs(center[4]) - k.box_max[4];
    T min_dist = sg_sqrt(v0 * v0 + v1 * v1);
    const T ld = sg_abs(lower_w[0]) < sg_abs(lower_w[1]) ? sg_abs(lower_w[0]) : sg_abs(lower_w[1]);
    if (ld < min_dist) min_dist = ld;
    const T ud = sg_abs(upper_w[0]) < sg_abs(upper_w[1]) ? sg_abs(upper_w[0]) : sg_abs(upper_w[1]);
    if (ud < min_dist) min_dist = ud;
    T dirs[SEEKER_MAXK][2], vb[SEEKER_MAXK];
    for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
      const T dx = aux[2 + 3 * i] - center[0], dy = aux[3 + 3 * i] - center[1];
      const T nrm = sg_sqrt(dx * dx + dy * dy);
      dirs[i][0] = dx / nrm; dirs[i][1] = dy / nrm;
      const T speed = center[3] * dirs[i][0] + center[4] * dirs[i][1];
      const T acc = sg_abs(T(2) * dirs[i][0] * max_acc[0] + T(2) * dirs[i][1] * max_acc[1]);
      // first pass (:727-745): the distance to the obstacle CENTRE picks the generator directions
      T d1 = nrm - T(5) * DT * speed;
      if (d1 < T(0)) d1 = T(1e-2);
      T b1 = sg_sqrt(d1 * acc) - speed;
      if (b1 < T(0)) b1 = T(0);
      if (b1 < min_dist) {
        Uv[0][0] = dirs[i][0]; Uv[1][0] = dirs[i][1];
        Uv[0][1] = dirs[i][1]; Uv[1][1] = -dirs[i][0];
        min_dist = b1;
      }
      // second pass (:748-764): the distance to the obstacle SURFACE bounds the speed towards it
      T d2 = nrm - aux[4 + 3 * i] - T(5) * DT * speed;
      if (d2 < T(0)) d2 = T(1e-2);
      vb[i] = sg_sqrt(d2 * acc) - speed;
      if (vb[i] < T(0)) vb[i] = T(0);
    }
    T al[4 + SEEKER_MAXK], be[4 + SEEKER_MAXK], ga[4 + SEEKER_MAXK];
    int n = 0;
    for (int d = 0; d < 2; ++d) {
      al[n] = sg_abs(Uv[d][0]); be[n] = sg_abs(Uv[d][1]); ga[n] = -lower_w[d]; ++n;
      al[n] = sg_abs(Uv[d][0]); be[n] = sg_abs(Uv[d][1]); ga[n] = upper_w[d]; ++n;
    }
    for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
      // as shipped: the parameter named "unscaled_generator_times_direction" receives the bare direction (:765)
      al[n] = sg_abs(dirs[i][0]); be[n] = sg_abs(dirs[i][1]); ga[n] = vb[i]; ++n;
    }
    T x, y;
    if (max_product_2d<T>(al, be, ga, n, x, y)) { lv[0] = x; lv[1] = y; }
  }
  // ---- assemble: G_s = [pos | vel | roll], its inverse, the noise budgets ---------------------------------------------------
  for (int i = 0; i < 36; ++i) { out.gen[i] = T(0); out.Ginv[i] = T(0); }
  const T roll_g = PI / T(12), rolld_g = PI / T(2);
  for (int r = 0; r < 2; ++r)
    for (int j = 0; j < 2; ++j) {
      out.gen[r * 6 + j] = Up[r][j] * lp[j];                   // columns 0, 1: positions (rows 0, 1)
      out.gen[(3 + r) * 6 + 2 + j] = Uv[r][j] * lv[j];         // columns 2, 3: velocities (rows 3, 4)
      out.Ginv[j * 6 + r] = Up[r][j] / lp[j];                  // (U diag(len))^-1 = diag(1 / len) U^T
      out.Ginv[(2 + j) * 6 + 3 + r] = Uv[r][j] / lv[j];
    }
  out.gen[2 * 6 + 4] = roll_g;                                  // column 4: roll (row 2), column 5: roll rate (row 5)
  out.gen[5 * 6 + 5] = rolld_g;
  out.Ginv[4 * 6 + 2] = T(1) / roll_g;
  out.Ginv[5 * 6 + 5] = T(1) / rolld_g;
  for (int i = 0; i < 6; ++i) {
    out.center[i] = center[i];
    T e = T(0), g = T(0);
    for (int j = 0; j < 6; ++j) {
      e += out.Ginv[i * 6 + j] * center[j];
      // (see below for g)
    }
    for (int q = 0; q < 2; ++q) {
      T t = T(0);
      for (int j = 0; j < 6; ++j) t += out.Ginv[i * 6 + j] * k.Gam_p[j][q];
      g += sg_abs(t);
    }
    out.cs_hat[i] = e;
    out.rb[i] = T(1) - g;
  }
}

}  // namespace sgbpo
