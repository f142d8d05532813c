// One safeguarded environment step for ONE environment, scalar-generic:
//   Safeguard.actions (safeguards/interfaces/safeguard.py:61-80)  ->  Simulator.step
//   (envs/simulators/interfaces/simulator.py:82-108): linearise, gate, safeguard map, execute action
//   (dynamics + clamp), reset-all substitution (Q6), observation, reward.
#pragma once
#include "sgbpo_safeguard.cuh"
#include "sgbpo_safeguard2.cuh"
#include "sgbpo_lifted.cuh"
#include "sgbpo_navquad.cuh"

namespace sgbpo {

template <class T> SG_HD bool is_close(T a, T b) {   // torch.isclose defaults: rtol 1e-5, atol 1e-8
  return sg_abs(a - b) <= T(1e-8) + T(1e-5) * sg_abs(b);
}
template <class T> SG_HD bool finite_val(T x) { return (x - x) == T(0); }

// Navigate* environments: elastic collision with the (last) obstacle the free step entered
// (envs/navigate_seeker.py:251-296; the remaining-time factor (1 - t) * DT is kept as shipped).
template <class TASK, class S, class T>
SG_HD void collision_operator(const S* s_prev, S* s_next, const T* aux, const SafeConsts<T>& k, int& flags) {
  int hit = -1;
  for (int i = 0; i < k.n_obstacles && i < SEEKER_MAXK; ++i) {
    const T dx = primal(s_next[0]) - aux[2 + 3 * i], dy = primal(s_next[1]) - aux[3 + 3 * i];
    if (sg_sqrt(dx * dx + dy * dy) <= aux[4 + 3 * i]) hit = i;
  }
  if (hit < 0) return;
  flags |= FL_COLLIDED;
  const S cx = S(aux[2 + 3 * hit]), cy = S(aux[3 + 3 * hit]), rad = S(aux[4 + 3 * hit]);
  const S tx = s_prev[0] - cx, ty = s_prev[1] - cy;
  const S vx = s_next[0] - s_prev[0], vy = s_next[1] - s_prev[1];
  const S qa = vx * vx + vy * vy;
  const S qb = C<S>(2.0) * (tx * vx + ty * vy);
  const S qc = tx * tx + ty * ty - rad * rad;
  const S t = (-qb - sg_sqrt(qb * qb - C<S>(4.0) * qa * qc)) / (C<S>(2.0) * qa);
  const S ix = s_prev[0] + t * vx, iy = s_prev[1] + t * vy;
  S nx = ix - cx, ny = iy - cy;
  const S nn = sg_sqrt(nx * nx + ny * ny);
  nx = nx / nn; ny = ny / nn;
  const S vdn = vx * nx + vy * ny;
  const S rx = vx - C<S>(2.0) * vdn * nx, ry = vy - C<S>(2.0) * vdn * ny;
  const S rem = (C<S>(1.0) - t) * C<S>(0.1);
  s_next[0] = ix + rx * rem;
  s_next[1] = iy + ry * rem;
}

// NavigateQuadrotor: elastic collision of the free step with the (last) obstacle it entered (navigate_quadrotor.py:275-328).
// Position and velocity are reflected; roll and roll rate of the collided environments keep their PRE-step values, as shipped.
template <class S, class T>
SG_HD void collision_operator_quad(const S* s_prev, S* s_next, const T* aux, int hit) {
  const S cx = S(aux[2 + 3 * hit]), cy = S(aux[3 + 3 * hit]), rad = S(aux[4 + 3 * hit]);
  const S tx = s_prev[0] - cx, ty = s_prev[1] - cy;
  const S fvx = s_next[3], fvy = s_next[4];
  const S vx = fvx * C<S>(0.05), vy = fvy * C<S>(0.05);
  const S qa = vx * vx + vy * vy;
  const S qb = C<S>(2.0) * (tx * vx + ty * vy);
  const S qc = tx * tx + ty * ty - rad * rad;
  const S t = (-qb - sg_sqrt(qb * qb - C<S>(4.0) * qa * qc)) / (C<S>(2.0) * qa);
  const S ix = s_prev[0] + t * vx, iy = s_prev[1] + t * vy;
  S nx = ix - cx, ny = iy - cy;
  const S nn = sg_sqrt(nx * nx + ny * ny);
  nx = nx / nn; ny = ny / nn;
  const S vdn = fvx * nx + fvy * ny;
  const S rx = fvx - C<S>(2.0) * vdn * nx, ry = fvy - C<S>(2.0) * vdn * ny;
  const S rem = (C<S>(1.0) - t) * C<S>(0.05);
  s_next[0] = ix + rx * rem;
  s_next[1] = iy + ry * rem;
  s_next[2] = s_prev[2];
  s_next[3] = rx;
  s_next[4] = ry;
  s_next[5] = s_prev[5];
}

template <class TASK, class S, class T>
SG_HD void safeguarded_step(const S* s, const S* a, const T* w, const T* aux, const T* dirs,
                            const StepShared<T>* sh, const SafeConsts<T>& k, bool reset_now,
                            const T* reset_state, S* s_next, S* a_exec, S* obs, S& reward, int& flags) {
  constexpr int NS = TASK::NS, NA = TASK::NA;
  flags = 0;
#pragma unroll
  for (int q = 0; q < NA; ++q) a_exec[q] = a[q];
  if (k.sg_kind != SG_NONE) {
    S c_f[NS];
    S B[NS][MAXM];
    linearise_cB<TASK>(s, k.noise_center, c_f, B);
    T cfp[NS];
    T Bp[NS][MAXM];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      cfp[i] = primal(c_f[i]);
#pragma unroll
      for (int q = 0; q < NA; ++q) Bp[i][q] = primal(B[i][q]);
    }
    const bool projectable = gate_projectable<TASK>(cfp, Bp, k);
    const bool use_guard = k.gate_mode == GATE_REFERENCE ? !projectable
                                                         : (k.gate_mode == GATE_INTENDED ? projectable : true);
    S g[NA];
    bool feasible;
    ActRows<T> dyn;
    dyn.count = 0;
    StateDyn<T> sdyn;                             // (only NavigateQuadrotor touches it)
    if constexpr (TASK::NAUX == NAVQUAD_NAUX) {
      if (k.state_constrained && k.state_model == SM_DYNAMIC) {      // this step's safe state set (P6-P8, no_grad in the reference)
        navquad_safe_state_set<T>(cfp, Bp, aux, k, sdyn);
        dyn.sdyn = &sdyn;
      }
    }
    if constexpr (NA == 1) {
      feasible = safeguard_m1<TASK>(a[0], c_f, B, k, g[0]);
    } else if (TASK::LIFTED_OK && k.state_constrained && k.state_model == SM_LIFTED) {
      // wide safe-state generator: the program is an iterative per-environment solve (sgbpo_lifted.cuh), run only
      // where the gate asks for its solution (the reference evaluates it everywhere and discards it, safeguard.py:67)
      if (!use_guard) {
#pragma unroll
        for (int q = 0; q < NA; ++q) g[q] = a[q];
        feasible = true;
      } else if (k.sg_kind == SG_BOUNDARY_PROJECTION) {
        if constexpr (TASK::LIFTED_OK) feasible = project_lifted_m2<TASK>(a, c_f, B, k, &dyn, g);
        else feasible = false;
      } else if (k.sg_kind == SG_RAY_MASK && !k.rm_zonotopic) {
        if constexpr (TASK::LIFTED_OK) feasible = ray_mask_orthogonal_lifted_m2<TASK>(a, c_f, B, k, &dyn, g);
        else feasible = false;
      } else {
#pragma unroll
        for (int q = 0; q < NA; ++q) g[q] = C<S>(NAN);
        feasible = false;
      }
    } else {
      bool set_ok = true;
      if (k.act_set_mode == 1) {                 // NavigateSeeker: the safe action set is rebuilt every step (P5)
        T sp[NS], U[2][2], len[2];
#pragma unroll
        for (int i = 0; i < NS; ++i) sp[i] = primal(s[i]);
        set_ok = seeker_action_set<T>(sp, aux, k, U, len);
        seeker_rows<T>(U, len, dyn);
      }
      if (dyn.sdyn && dyn.sdyn->degenerate) {      // no usable safe state set at this step: nothing to project onto
        feasible = false;
#pragma unroll
        for (int q = 0; q < NA; ++q) g[q] = C<S>(NAN);
      } else {
        feasible = safeguard_m2<TASK>(a, c_f, B, dirs, k, &dyn, g) && set_ok;
      }
    }
    if (projectable) flags |= FL_PROJECTABLE;
    if (use_guard) {
      flags |= FL_GUARD_USED;
      if (!feasible) flags |= FL_INFEASIBLE;
      if (k.state_constrained && k.state_model == SM_UNREDUCED) flags |= FL_UNREDUCED;
      const bool straight_through = (k.sg_kind == SG_RAY_MASK) && k.rm_passthrough;   // Q11
      if (feasible || !k.infeasible_raw) {       // (infeasible_raw: the raw action stays in place where no safe action exists)
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          if (straight_through) a_exec[q] = a[q] + S(primal(g[q]) - primal(a[q]));
          else a_exec[q] = g[q];
        }
      }
    }
    int differ = 0;
    bool fin = true;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      differ += is_close(primal(a_exec[q]), primal(a[q])) ? 0 : 1;
      fin = fin && finite_val(primal(a_exec[q]));
    }
    if (differ == NA) flags |= FL_INTERVENED;
    if (!fin) flags |= FL_NONFINITE;
    // verdicts on the executed action
    T ap[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) ap[q] = primal(a_exec[q]);
    if (k.action_constrained && action_in_set<TASK>(ap, k, &dyn)) flags |= FL_ACTION_IN_SET;
    if (k.state_constrained && next_state_in_set<TASK>(ap, cfp, Bp, k, &dyn)) flags |= FL_NEXT_IN_SET;
  }
  S wS[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) wS[i] = S(w[i]);
  TASK::Sim::f(s, a_exec, wS, s_next);
  bool quad_hit = false;
  if constexpr (TASK::NAUX == NAVQUAD_NAUX) {
    if (k.n_obstacles > 0) {
      T fp[2] = {primal(s_next[0]), primal(s_next[1])};
      const int hit = navquad_last_hit<T>(fp, aux, k);
      if (hit >= 0) { quad_hit = true; flags |= FL_COLLIDED; collision_operator_quad<S, T>(s, s_next, aux, hit); }
    }
  } else if constexpr (TASK::NAUX > 2) {
    if (k.n_obstacles > 0) collision_operator<TASK>(s, s_next, aux, k, flags);
  }
  if (TASK::CLAMP && !k.disable_clamp) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const T p = primal(s_next[i]);
      if (p < k.box_min[i]) { s_next[i] = S(k.box_min[i]); flags |= FL_CLAMPED; }
      else if (p > k.box_max[i]) { s_next[i] = S(k.box_max[i]); flags |= FL_CLAMPED; }
    }
  }
  if (reset_now) {
#pragma unroll
    for (int i = 0; i < NS; ++i) s_next[i] = S(reset_state[i]);
  }
  TASK::observe(s_next, aux, sh, obs);
  reward = TASK::reward(s_next, a_exec, aux, sh);
  if constexpr (TASK::NAUX == NAVQUAD_NAUX) {
    if (quad_hit) reward = reward - C<S>(10.0);                   // COLLISION_PENALTY * collided.any() (navigate_quadrotor.py:236)
    const T dx = aux[0] - primal(s_next[0]), dz = aux[1] - primal(s_next[1]);
    if (aux[NAVQUAD_NAUX - 1] == T(0) && sg_sqrt(dx * dx + dz * dz) < T(0.1)) flags |= FL_REACHED;
  }
}

}  // namespace sgbpo
