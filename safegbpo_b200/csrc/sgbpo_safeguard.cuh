// Safeguard maps (gate, boundary projection P1, ray mask P2/P3/P4 + radial map), scalar-generic.
//
// The reference solves one conic program per environment per step on the host
// (safeguards/boundary_projection.py:33-57, safeguards/ray_mask.py:135-341, constraint blocks from
// safeguards/interfaces/safeguard.py:98-208 and sets/zonotope.py:82-141).  Here each program is
// reduced ONCE, at constant-registration time, to a polytope in ACTION space whose rows are affine in
// the per-step linearisation (c_f, B); the per-environment solve is then exact arithmetic in registers:
//
//   state block   enc(Z(c_f + B a, G_w) in Z(c_s, G_s))
//     model 0 (G_s invertible: cartpoles, household, seeker-free envs):  Gamma = G_s^-1 G_w is forced,
//       row budgets rb_i = 1 - sum_j |Gamma_ij|,  rows  |(G_s^-1 (c_s - c_f - B a))_i| <= rb_i.
//     model 1 (n = 2, G_s 2 x p: pendulum): the feasible d = c_s - c_f - B a form a FIXED convex polygon
//       K (projection of the lifted polytope), registered as half-planes  hk . d <= 1.
//   action block  pt(a in Z(c_a, G_a)):  m = 1 -> interval;  m = 2 -> zonotope polygon half-planes.
//   feasibility   -1 <= a <= 1.
//
// Differentiation: every map is evaluated with the active set chosen on primal values and the final
// closed form evaluated in the scalar type S, so Dual<> instantiations yield the implicit-function
// derivative of the optimal solution (what cvxpylayers' backward returns away from degeneracy).
#pragma once
#include "sgbpo_env.cuh"

namespace sgbpo {

constexpr int MAXHP = 48;      // half-planes for polygon models
constexpr int MAXROWS = 64;    // rows of a per-env action-space polytope
constexpr int MAXGEN = 20;     // merged generators of a wide safe-state zonotope (lifted model)
constexpr int MAXNULL = 14;    // dimension of their null space

enum SgKind : int { SG_NONE = 0, SG_BOUNDARY_PROJECTION = 1, SG_RAY_MASK = 2 };
enum GateMode : int { GATE_REFERENCE = 0, GATE_INTENDED = 1, GATE_ALWAYS = 2 };
// SM_UNREDUCED: safe state set whose generator is neither invertible nor 2-D (BalanceQuadrotor's 6 x 24 RCI set): the
// linearisation, the gate and the verdict on the action set run, but the program itself has no register-level
// reduction yet -- wherever the gate asks for the safeguarded action the step reports FL_UNREDUCED (+ NaN action)
// SM_DYNAMIC: the safe state set is rebuilt every step per environment (NavigateQuadrotor, sgbpo_navquad.cuh); its generator is
// block-invertible, so the rows are those of SM_INVERTIBLE with per-environment constants
enum StateModel : int { SM_INVERTIBLE = 0, SM_POLYGON = 1, SM_UNREDUCED = 2, SM_LIFTED = 3, SM_DYNAMIC = 4 };

// status / verdict bits written per environment per step
enum Flag : int {
  FL_PROJECTABLE = 1,        // state_set.intersects(reachable_set)           (safeguard.py:65)
  FL_INFEASIBLE = 2,         // the program of this step has an empty feasible set (reference: solver error)
  FL_INTERVENED = 4,         // all action dims differ (isclose) from the policy action (safeguard.py:79)
  FL_NONFINITE = 8,          // NaN/Inf safe action (reference raises ValueError, safeguard.py:68-75)
  FL_GUARD_USED = 16,        // the safeguarded action (not the raw one) was executed
  FL_ACTION_IN_SET = 32,     // verdict: executed action inside the safe action set (where one exists)
  FL_NEXT_IN_SET = 64,       // verdict: linearised next-state zonotope inside the safe state set
  FL_CLAMPED = 128,          // execute_action clamped the state to the feasible box
  FL_COLLIDED = 256,         // Navigate*: the free step entered an obstacle (collision operator applied)
  FL_UNREDUCED = 512,        // the safeguarded action was needed but this program has no GPU reduction (SM_UNREDUCED)
  FL_REACHED = 1024          // NavigateQuadrotor: the goal was reached at this step for the first time (reward bonus paid)
};

template <class T> struct SafeConsts {
  int sg_kind, gate_mode;
  int rm_linear, rm_zonotopic, rm_passthrough;
  int action_constrained, state_constrained, state_model;
  int disable_clamp, n_obstacles;
  int act_set_mode, infeasible_raw;
  T box_min[MAXN], box_max[MAXN];      // feasible state box: gate (box.py:139-154) and clamp
  T noise_center[MAXN];
  // action block
  T a_lo[MAXM], a_hi[MAXM];            // m = 1: the 1-D zonotope [c_a - sum|G_a|, c_a + sum|G_a|]
  int n_act_hp;                        // m = 2: zonotope polygon, rows n . a <= h
  T act_hp[MAXHP][3];
  T c_a[MAXM];
  // state block, model 0
  T Ginv[MAXN * MAXN];                 // G_s^-1
  T cs_hat[MAXN];                      // G_s^-1 c_s
  T rb[MAXN];                          // row budgets 1 - sum_j |(G_s^-1 G_w)_ij|
  T B0hat_abs[MAXN * MAXM];            // |G_s^-1 B0| (Q4: init-time action matrix of env 0), for P2
  // state block, model 1 (n = 2): K = {d : hp . d <= 1};  Q = {(d, L): hq . (d, L) <= 1} for P2
  int n_k_hp, n_q_hp;
  T k_hp[MAXHP][2];
  T q_hp[MAXHP][3];
  T c_s[MAXN];
  T B0hat[MAXN * MAXM];                // G_s^-1 B0 (signed), for P2 with m = 2
  // state block, model 3 (lifted): beta = Gpinv d + Nmat y, Gamma_j = Gam_p[:, j] + Nmat z_j
  int n_gen, n_null, n_wcol, reserved2;
  T Gpinv[MAXGEN][MAXN];
  T Nmat[MAXGEN][MAXNULL];
  T Gam_p[MAXGEN][2];
};

// ---- gate: AxisAlignedBox.intersects(Zonotope(c_f, B I_m))  (box.py:139-154, simulator.py:240-243) --
template <class TASK, class T>
SG_HD bool gate_projectable(const T* c_f, const T (*B)[MAXM], const SafeConsts<T>& k) {
  constexpr int NS = TASK::NS, NA = TASK::NA;
  bool ok = true;
  T ctr[NS], half[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const T bc = T(0.5) * (k.box_min[i] + k.box_max[i]);
    half[i] = T(0.5) * (k.box_max[i] - k.box_min[i]);
    ctr[i] = c_f[i] - bc;
    T sup = T(0);
#pragma unroll
    for (int a = 0; a < NA; ++a) sup += sg_abs(B[i][a]);
    ok = ok && (sg_abs(ctr[i]) <= half[i] + sup);
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {       // axes = normalised generators b / |b|:  |ctr . b| / |b| <= |b| + sum_i |half_i b_i| / |b|,
    // evaluated multiplied through by |b| (no square root, no divisions); a zero column is NaN -> false in the reference
    T nrm2 = T(0), proj = T(0), sup = T(0);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      nrm2 += B[i][a] * B[i][a];
      proj += ctr[i] * B[i][a];
      sup += sg_abs(half[i] * B[i][a]);
    }
    ok = ok && nrm2 > T(0) && (sg_abs(proj) <= nrm2 + sup);
  }
  return ok;
}

// ================================================================================================
// m = 1: every program reduces to operations on one interval [lo, hi]
// ================================================================================================
template <class S> struct Interval { S lo, hi; bool feasible; };

// rows of the state block in the 1-D action:  alpha * a <= gamma  (S-typed, streamed)
template <class TASK, class S, class T, class F>
SG_HD void for_each_state_row_m1(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, F&& emit) {
  constexpr int NS = TASK::NS;
  if (k.state_model == SM_INVERTIBLE) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S e = S(k.cs_hat[i]), b = S(T(0));
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        e = e - S(k.Ginv[i * NS + j]) * c_f[j];
        b = b + S(k.Ginv[i * NS + j]) * B[j][0];
      }
      // |e - b a| <= rb   <=>   -b a <= rb - e   and   b a <= rb + e
      emit(-b, S(k.rb[i]) - e, i);
      emit(b, S(k.rb[i]) + e, i);
    }
  } else {
    // polygon K in d = c_s - c_f - B a :  hp . d <= 1   <=>   -(hp . B) a <= 1 - hp . (c_s - c_f)
    for (int r = 0; r < k.n_k_hp; ++r) {
      S e = S(T(0)), b = S(T(0));
#pragma unroll
      for (int j = 0; j < (NS < 2 ? NS : 2); ++j) {       // polygon model is registered for n = 2 only
        e = e + S(k.k_hp[r][j]) * (S(k.c_s[j]) - c_f[j]);
        b = b + S(k.k_hp[r][j]) * B[j][0];
      }
      emit(-b, C<S>(1.0) - e, r);
    }
  }
}

template <class TASK, class S, class T>
SG_HD Interval<S> safe_interval_m1(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k) {
  Interval<S> iv;
  iv.lo = C<S>(-1.0);
  iv.hi = C<S>(1.0);
  iv.feasible = true;
  if (k.action_constrained) {
    iv.lo = sg_max(iv.lo, S(k.a_lo[0]));
    iv.hi = sg_min(iv.hi, S(k.a_hi[0]));
  }
  if (k.state_constrained && k.state_model == SM_INVERTIBLE) {
    // |e_i - b_i a| <= rb_i per row (for_each_state_row_m1's two rows), with ONE reciprocal per row and no dependence between
    // the rows' divisions:  b > 0: -(rb - e)/b <= a <= (rb + e)/b;  b < 0: (rb + e)/b <= a <= -(rb - e)/b
    constexpr int NS = TASK::NS;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S e = S(k.cs_hat[i]), b = S(T(0));
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        e = e - S(k.Ginv[i * NS + j]) * c_f[j];
        b = b + S(k.Ginv[i * NS + j]) * B[j][0];
      }
      const T bp = primal(b);
      if (bp != T(0)) {
        const S r = C<S>(1.0) / b;
        const S q1 = (S(k.rb[i]) + e) * r, q2 = -((S(k.rb[i]) - e) * r);
        iv.hi = sg_min(iv.hi, bp > T(0) ? q1 : q2);
        iv.lo = sg_max(iv.lo, bp > T(0) ? q2 : q1);
      } else if (primal(S(k.rb[i]) - e) < T(0) || primal(S(k.rb[i]) + e) < T(0)) {
        iv.feasible = false;
      }
    }
  } else if (k.state_constrained) {
    for_each_state_row_m1<TASK>(c_f, B, k, [&](const S& alpha, const S& gamma, int) {
      const T al = primal(alpha);
      if (al > T(0)) iv.hi = sg_min(iv.hi, gamma / alpha);
      else if (al < T(0)) iv.lo = sg_max(iv.lo, gamma / alpha);
      else if (primal(gamma) < T(0)) iv.feasible = false;
    });
  }
  if (primal(iv.lo) > primal(iv.hi)) iv.feasible = false;
  return iv;
}

// P1 (boundary_projection.py:33-57) for m = 1
template <class S> SG_HD S project_interval(const S& a, const Interval<S>& iv) {
  return sg_min(sg_max(a, iv.lo), iv.hi);
}

// exact distance from c to the boundary of [-1, 1] along d (ray_mask.py:243-266), m = 1
template <class S> SG_HD S unit_box_dist_m1(const S& c, const S& d) {
  using T = typename Traits<S>::base;
  const T dp = primal(d);
  if (dp == T(0)) return C<S>(INFINITY);
  const S hi = (C<S>(1.0) - c) / d, lo = (C<S>(-1.0) - c) / d;
  const bool hi_ok = !(primal(hi) < T(0)), lo_ok = !(primal(lo) < T(0));
  if (hi_ok && lo_ok) return sg_min(hi, lo);
  if (hi_ok) return hi;
  if (lo_ok) return lo;
  return C<S>(INFINITY);
}

// P2 for m = 1 (ray_mask.py:145-191): max geo_mean(l1, l2) == max L = l1 + l2 with l1 = l2 (the program is
// symmetric in the two collinear generators), a 2-variable LP in (c, L):  alpha c + beta L <= gamma.
template <class S> struct Lp2Row { S alpha, beta, gamma; };

template <class TASK, class S, class T, class F>
SG_HD void for_each_p2_row_m1(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, F&& emit) {
  constexpr int NS = TASK::NS;
  int idx = 0;
  emit(C<S>(1.0), C<S>(1.0), C<S>(1.0), idx++);      // c + L <= 1
  emit(C<S>(-1.0), C<S>(1.0), C<S>(1.0), idx++);     // -c + L <= 1
  if (k.action_constrained) {                        // 1-D zonotope containment is exact: |c - c_a| + L <= r_a
    emit(C<S>(1.0), C<S>(1.0), S(k.a_hi[0]), idx++);
    emit(C<S>(-1.0), C<S>(1.0), -S(k.a_lo[0]), idx++);
  }
  if (k.state_constrained) {
    if (k.state_model == SM_INVERTIBLE) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        S e = S(k.cs_hat[i]), b = S(T(0));
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          e = e - S(k.Ginv[i * NS + j]) * c_f[j];
          b = b + S(k.Ginv[i * NS + j]) * B[j][0];
        }
        const S w = S(k.B0hat_abs[i * MAXM + 0]);
        emit(-b, w, S(k.rb[i]) - e, idx++);          //  (e - b c) + w L <= rb
        emit(b, w, S(k.rb[i]) + e, idx++);           // -(e - b c) + w L <= rb
      }
    } else {
      for (int r = 0; r < k.n_q_hp; ++r) {           // hq . (c_s - c_f - B c, L) <= 1
        S e = S(T(0)), b = S(T(0));
#pragma unroll
        for (int j = 0; j < (NS < 2 ? NS : 2); ++j) {
          e = e + S(k.q_hp[r][j]) * (S(k.c_s[j]) - c_f[j]);
          b = b + S(k.q_hp[r][j]) * B[j][0];
        }
        emit(-b, S(k.q_hp[r][2]), C<S>(1.0) - e, idx++);
      }
    }
  }
}

template <class TASK, class S, class T>
SG_HD bool p2_expand_m1(const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, S& c_out, S& L_out) {
  // pass 1 on primal values: incremental 2-variable LP (Seidel).  Rows 0, 1 bound the problem (L <= 1 - |c|);
  // every further row that cuts off the current optimum moves it onto that row's boundary, where a 1-D LP over
  // the earlier rows finds the new optimum.  Rows are re-generated by the emitter (a few FMAs each) instead of
  // being staged in local memory; the cost is O(n * (1 + #rows that ever cut)), not the O(n^3) of enumerating pairs.
  T cf_p[MAXN];
  T B_p[MAXN][MAXM];
#pragma unroll
  for (int i = 0; i < TASK::NS; ++i) { cf_p[i] = primal(c_f[i]); B_p[i][0] = primal(B[i][0]); }
  const T vtol = sizeof(T) == 4 ? T(1e-6) : T(1e-11);
  T best_c = T(0), best_L = T(1);
  int bj = 0, bk = 1;
  bool feasible = true;
  for (int start = 2; feasible;) {
    // first row at or after `start` violated by the current optimum
    int vi = -1;
    T va = T(0), vb = T(0), vg = T(0);
    for_each_p2_row_m1<TASK, T, T>(cf_p, B_p, k, [&](T al, T be, T ga, int idx) {
      if (vi >= 0 || idx < start) return;
      if (al * best_c + be * best_L > ga + vtol * (T(1) + sg_abs(ga))) { vi = idx; va = al; vb = be; vg = ga; }
    });
    if (vi < 0) break;
    // 1-D LP on the boundary of row vi against rows 0 .. vi-1
    if (vb > T(0)) {
      // L = (vg - va c) / vb ; rows r:  (al - be va / vb) c <= ga - be vg / vb
      T lo = T(-INFINITY), hi = T(INFINITY);
      int rlo = -1, rhi = -1;
      for_each_p2_row_m1<TASK, T, T>(cf_p, B_p, k, [&](T al, T be, T ga, int idx) {
        if (idx >= vi) return;
        const T coef = al - be * va / vb, rhs = ga - be * vg / vb;
        if (coef > T(1e-30)) { const T v = rhs / coef; if (v < hi) { hi = v; rhi = idx; } }
        else if (coef < T(-1e-30)) { const T v = rhs / coef; if (v > lo) { lo = v; rlo = idx; } }
        else if (rhs < -vtol * (T(1) + sg_abs(ga))) feasible = false;
      });
      if (!feasible || lo > hi + vtol * (T(1) + sg_abs(hi))) { feasible = false; break; }
      // maximise L = (vg - va c) / vb:  slope -va / vb
      const bool take_lo = va > T(0) || (va == T(0));
      if (take_lo && rlo >= 0) { best_c = lo; bk = rlo; }
      else if (rhi >= 0) { best_c = hi; bk = rhi; }
      else if (rlo >= 0) { best_c = lo; bk = rlo; }
      else { feasible = false; break; }
      best_L = (vg - va * best_c) / vb;
      bj = vi;
    } else {
      // vertical boundary c = vg / va: L is limited by the earlier rows with a positive L coefficient
      if (sg_abs(va) < T(1e-30)) { feasible = false; break; }
      best_c = vg / va;
      T lim = T(INFINITY);
      int rl = -1;
      for_each_p2_row_m1<TASK, T, T>(cf_p, B_p, k, [&](T al, T be, T ga, int idx) {
        if (idx >= vi) return;
        if (be > T(0)) { const T v = (ga - al * best_c) / be; if (v < lim) { lim = v; rl = idx; } }
        else if (al * best_c > ga + vtol * (T(1) + sg_abs(ga))) feasible = false;
      });
      if (!feasible || rl < 0) { feasible = false; break; }
      best_L = lim; bj = vi; bk = rl;
    }
    start = vi + 1;
  }
  if (!feasible) { bj = -1; }
  if (bj < 0 || best_L < T(0)) { c_out = C<S>(NAN); L_out = C<S>(NAN); return false; }
  // pass 2: S-typed solve on the two active rows
  Lp2Row<S> r0, r1;
  r0.alpha = r0.beta = r0.gamma = r1.alpha = r1.beta = r1.gamma = C<S>(0.0);
  for_each_p2_row_m1<TASK, S, T>(c_f, B, k, [&](const S& al, const S& be, const S& ga, int idx) {
    if (idx == bj) { r0.alpha = al; r0.beta = be; r0.gamma = ga; }
    if (idx == bk) { r1.alpha = al; r1.beta = be; r1.gamma = ga; }
  });
  const S det = r0.alpha * r1.beta - r1.alpha * r0.beta;
  c_out = (r0.gamma * r1.beta - r1.gamma * r0.beta) / det;
  L_out = (r0.alpha * r1.gamma - r1.alpha * r0.gamma) / det;
  (void)best_c;
  return true;
}

// radial map (ray_mask.py:80-89, 108-113), as shipped (Q2)
template <class S, class T>
SG_HD S radial_map(const S& dist, const S& safe_dist, const S& feas_dist, const SafeConsts<T>& k) {
  if (k.rm_linear) return dist / safe_dist;
  return sg_tanh(dist / safe_dist) / sg_tanh(feas_dist / safe_dist);
}

// full safeguard(a) for m = 1.  Returns false when the program is infeasible.
template <class TASK, class S, class T>
SG_HD bool safeguard_m1(const S& a, const S* c_f, const S (*B)[MAXM], const SafeConsts<T>& k, S& out) {
  const Interval<S> iv = (k.sg_kind == SG_BOUNDARY_PROJECTION || !k.rm_zonotopic || !k.state_constrained)
                             ? safe_interval_m1<TASK>(c_f, B, k)
                             : Interval<S>{C<S>(0.0), C<S>(0.0), true};
  if (k.sg_kind == SG_BOUNDARY_PROJECTION) {
    // an empty interval has no projection: NaN like the m = 2 path (and like the reference's solver failure), never a
    // finite action that violates a constraint
    out = iv.feasible ? project_interval(a, iv) : C<S>(NAN);
    return iv.feasible;
  }
  // ---- ray mask ----
  S center, safe_dist, feas_dist;
  bool feasible = true;
  if (k.state_constrained && k.rm_zonotopic) {
    S L;
    feasible = p2_expand_m1<TASK>(c_f, B, k, center, L);
    // P3 (ray_mask.py:215-240) on Z(c, [+-l1, +-l2]) along dir = (a-c)/(|a-c|+1e-8):  t* = L / |dir|
    const S diff = a - center;
    const S nrm = sg_abs(diff);
    const S dir = diff / (nrm + C<S>(1e-8));
    safe_dist = L / sg_abs(dir);
    feas_dist = unit_box_dist_m1(center, dir);
  } else if (k.state_constrained) {
    // orthogonal approximation (ray_mask.py:268-341): start = P1(a); chord from start along (start - a)
    feasible = iv.feasible;
    const S start = project_interval(a, iv);
    const S gap = sg_abs(start - a);
    const bool safe = primal(gap) < T(1e-8);
    if (safe) {
      center = start; safe_dist = C<S>(1.0); feas_dist = C<S>(1.0);
    } else {
      const bool up = primal(start) > primal(a);
      const S dir = up ? C<S>(1.0) : C<S>(-1.0);
      const S t = up ? (iv.hi - start) : (start - iv.lo);      // P4: max t with start + t dir in [lo, hi]
      safe_dist = t * C<S>(0.5);
      center = start + dir * safe_dist;
      feas_dist = unit_box_dist_m1(center, -dir);
    }
  } else {
    // action-constrained only: centre of the safe action set, P3 on G_a  (ray_mask.py:77-78)
    center = S(k.c_a[0]);
    const S diff = a - center;
    const S dir = diff / (sg_abs(diff) + C<S>(1e-8));
    safe_dist = (S(k.a_hi[0]) - S(k.a_lo[0])) * C<S>(0.5) / sg_abs(dir);
    feas_dist = unit_box_dist_m1(center, dir);
  }
  const S diff = a - center;
  const S dist = sg_abs(diff);
  if (primal(dist) < T(1e-8)) { out = center; return feasible; }
  const S dir = diff / (dist + C<S>(1e-8));
  out = center + dir * radial_map(dist, safe_dist, feas_dist, k);
  return feasible;
}

}  // namespace sgbpo
