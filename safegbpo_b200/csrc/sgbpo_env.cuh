// Per-environment maps of the safeguarded rollout, written ONCE as scalar-generic templates:
// instantiated with float/double for the forward kernels and with (nested) Dual<> for the reverse
// pass.  Also compiles as plain C++ so the same code is unit-tested on a GPU-less host
// (tests/hostmath) -- test infrastructure only, never a product path.
//
// Reference behaviour restated here (file:line under /root/reference/src):
//   dynamics      envs/simulators/{pendulum.py:179-190, cartpole.py:160-180, quadrotor.py:204-221,
//                 seeker.py:185, household.py:197-213}
//   observation   pendulum.py:92-96, cartpole.py:83-88, quadrotor.py:132, seeker.py:113, household.py:109-113
//   reward        balance_pendulum.py:81-84, balance_cartpole.py:77-82, swingup_cartpole.py:78-83,
//                 balance_quadrotor.py:90-101, manage_household.py:57-69, navigate_seeker.py:212
//   execute       envs/simulators/interfaces/simulator.py:174-182 (clamp), household.py:168-177 (no clamp)
//   linearise     simulator.py:206-230      reachable set: simulator.py:233-243
#pragma once
#include "sgbpo_dual.cuh"

namespace sgbpo {

constexpr int MAXN = 6;   // state dim
constexpr int MAXM = 2;   // action dim
constexpr int MAXO = 24;  // observation dim (household 23)

enum EnvId : int {
  ENV_BALANCE_PENDULUM = 0,
  ENV_BALANCE_CARTPOLE = 1,
  ENV_SWINGUP_CARTPOLE = 2,
  ENV_BALANCE_QUADROTOR = 3,
  ENV_MANAGE_HOUSEHOLD = 4,
  ENV_NAVIGATE_SEEKER = 5,
  ENV_NAVIGATE_QUADROTOR = 6,
  ENV_COUNT = 7
};

// ---- simulators --------------------------------------------------------------------------------
struct PendulumSim {
  static constexpr int NS = 2, NA = 1;
  template <class S> SG_HD static void f(const S* s, const S* a, const S* w, S* o) {
    const S acc = sg_sin(s[0]) * C<S>(1.5 * 9.81 / 1.0) + (a[0] * C<S>(30.0)) * C<S>(3.0) + w[1];
    o[1] = s[1] + C<S>(0.05) * acc;
    o[0] = wrap_angle(s[0] + C<S>(0.05) * o[1]);
  }
};

struct CartPoleSim {
  static constexpr int NS = 4, NA = 1;
  template <class S> SG_HD static void f(const S* s, const S* a, const S* w, S* o) {
    S sn, cs;
    sg_sincos(s[1], sn, cs);
    const S tmp = (a[0] * C<S>(10.0) + C<S>(0.1 * 0.5) * (s[3] * s[3]) * sn) / C<S>(1.1);
    const S thdd = C<S>(2.0) * ((C<S>(9.8) * sn + tmp * cs) / (C<S>(4.0 / 3.0) - C<S>(0.1 / 1.1) * (cs * cs)) + w[0]);
    const S xdd = tmp - C<S>(0.1 * 0.5 / 1.1) * cs * thdd;
    o[2] = s[2] + C<S>(0.02) * xdd;
    o[3] = s[3] + C<S>(0.02) * thdd;
    o[0] = s[0] + C<S>(0.02) * o[2];
    o[1] = wrap_angle(s[1] + C<S>(0.02) * o[3]);
  }
};

struct QuadrotorSim {
  static constexpr int NS = 6, NA = 2;
  template <class S> SG_HD static void f(const S* s, const S* a, const S* w, S* o) {
    const S thrust = a[0] * C<S>(6.0 / 7.0) + C<S>(9.81);
    const S roll_cmd = a[1] * C<S>(3.14159265358979323846 / 12.0);
    S sr, cr;
    sg_sincos(s[2], sr, cr);
    const S xdd = thrust * sr + w[3];
    const S zdd = thrust * cr - C<S>(9.81) + w[4];
    const S rdd = roll_cmd * C<S>(55.0) - C<S>(70.0) * s[2] - C<S>(17.0) * s[5];
    o[3] = s[3] + C<S>(0.05) * xdd;
    o[4] = s[4] + C<S>(0.05) * zdd;
    o[5] = s[5] + C<S>(0.05) * rdd;
    o[0] = s[0] + C<S>(0.05) * o[3];
    o[1] = s[1] + C<S>(0.05) * o[4];
    o[2] = wrap_angle(s[2] + C<S>(0.05) * o[5]);
  }
};

struct SeekerSim {
  static constexpr int NS = 2, NA = 2;
  template <class S> SG_HD static void f(const S* s, const S* a, const S* w, S* o) {
    o[0] = s[0] + C<S>(0.1) * a[0] + w[0];
    o[1] = s[1] + C<S>(0.1) * a[1] + w[1];
  }
};

struct HouseholdSim {
  static constexpr int NS = 3, NA = 2;
  template <class S> SG_HD static void f(const S* s, const S* a, const S* w, S* o) {
    constexpr double h_fh = 1.1, h_out = 0.26, tau = 240.0, c_w = 1.1625, cop = 3.01;
    constexpr double c0 = (h_fh + h_out) / (h_out * tau), c1 = h_fh / (h_out * tau), c2 = h_fh / c_w, c4 = cop / c_w;
    o[0] = s[0] + a[0] * C<S>(2.0);
    o[1] = s[1] + (-(C<S>(c0) * s[1]) + C<S>(c1) * s[2] + w[1] / C<S>(tau));
    o[2] = s[2] + (C<S>(c2) * s[1] - C<S>(c2) * s[2] + C<S>(c4) * (a[1] * C<S>(2.5)));
  }
};

// ---- per-step shared scalars (household time series, same for every environment) ---------------
template <class T> struct StepShared {
  T series[20];   // household obs tail: load[t:t+5], pv[t:t+5], t_out[t:t+5], price[t:t+5] at the POST-step index
  T load_t, pv_t, price_t;   // reward terms at the post-step index (manage_household.py:60-66)
};

// ---- task environments ---------------------------------------------------------------------------
// aux: per-environment constants that are not part of the state (quadrotor / seeker goal).
template <int ENV> struct Task;

template <> struct Task<ENV_BALANCE_PENDULUM> {
  using Sim = PendulumSim;
  static constexpr int NS = 2, NA = 1, NO = 3, NAUX = 0;
  static constexpr bool CLAMP = true;
  static constexpr bool LIFTED_OK = false;   // wide safe-state generator -> per-env interior point (sgbpo_lifted.cuh)
  template <class S, class T> SG_HD static void observe(const S* s, const T*, const StepShared<T>*, S* o) {
    sg_sincos(s[0], o[0], o[1]); o[2] = s[1];
  }
  template <class S, class T> SG_HD static S reward(const S* s, const S* a, const T*, const StepShared<T>*) {
    return -(s[0] * s[0]) - C<S>(0.1) * (s[1] * s[1]) - C<S>(0.001) * (a[0] * a[0]);
  }
};

template <int ENV> struct CartPoleTask {
  using Sim = CartPoleSim;
  static constexpr int NS = 4, NA = 1, NO = 5, NAUX = 0;
  static constexpr bool CLAMP = true;
  static constexpr bool LIFTED_OK = false;   // wide safe-state generator -> per-env interior point (sgbpo_lifted.cuh)
  template <class S, class T> SG_HD static void observe(const S* s, const T*, const StepShared<T>*, S* o) {
    o[0] = s[0]; sg_sincos(s[1], o[1], o[2]); o[3] = s[2]; o[4] = s[3];
  }
  template <class S, class T> SG_HD static S reward(const S* s, const S* a, const T*, const StepShared<T>*) {
    return -(C<S>(0.05) * (s[0] * s[0])) - C<S>(0.1) * (s[2] * s[2]) - (s[1] * s[1]) - C<S>(0.1) * (s[3] * s[3])
           - C<S>(0.0) * (a[0] * a[0]);
  }
};
template <> struct Task<ENV_BALANCE_CARTPOLE> : CartPoleTask<ENV_BALANCE_CARTPOLE> {};
template <> struct Task<ENV_SWINGUP_CARTPOLE> : CartPoleTask<ENV_SWINGUP_CARTPOLE> {};

template <> struct Task<ENV_BALANCE_QUADROTOR> {
  using Sim = QuadrotorSim;
  static constexpr int NS = 6, NA = 2, NO = 8, NAUX = 2;   // aux = goal (x, z)
  static constexpr bool CLAMP = true;
  static constexpr bool LIFTED_OK = true;   // wide safe-state generator -> per-env interior point (sgbpo_lifted.cuh)
  template <class S, class T> SG_HD static void observe(const S* s, const T* aux, const StepShared<T>*, S* o) {
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = s[i];
    o[6] = S(aux[0]); o[7] = S(aux[1]);
  }
  template <class S, class T> SG_HD static S reward(const S* s, const S* a, const T* aux, const StepShared<T>*) {
    const S thrust = a[0] * C<S>(6.0 / 7.0) + C<S>(9.81);
    const S roll_cmd = a[1] * C<S>(3.14159265358979323846 / 12.0);
    const S dx = s[0] - S(aux[0]), dz = s[1] - S(aux[1]);
    const S q2 = dx * dx + dz * dz;
    // At distance exactly 0 the state IS the goal, which only happens on the reset step (goal = reset position,
    // quadrotor.py:119): the reference's post-reset state carries no autograd graph there, so torch.sqrt's infinite slope is
    // never evaluated.  Forward-mode duals would form 0 * inf = NaN; the distance is a constant 0 instead.
    const S dist = primal(q2) == T(0) ? C<S>(0.0) : sg_sqrt(q2);
    return -(C<S>(2.5) * dist) - C<S>(0.1) * (s[2] * s[2]) - C<S>(0.1) * (s[3] * s[3]) - C<S>(0.1) * (s[4] * s[4])
           - C<S>(0.1) * (s[5] * s[5]) - C<S>(0.002) * (thrust * thrust) - C<S>(0.001) * (roll_cmd * roll_cmd);
  }
};

template <> struct Task<ENV_MANAGE_HOUSEHOLD> {
  using Sim = HouseholdSim;
  static constexpr int NS = 3, NA = 2, NO = 23, NAUX = 0;
  static constexpr bool CLAMP = false;   // Q7
  static constexpr bool LIFTED_OK = false;   // wide safe-state generator -> per-env interior point (sgbpo_lifted.cuh)
  template <class S, class T> SG_HD static void observe(const S* s, const T*, const StepShared<T>* sh, S* o) {
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
#pragma unroll
    for (int i = 0; i < 20; ++i) o[3 + i] = S(sh->series[i]);
  }
  template <class S, class T> SG_HD static S reward(const S* s, const S* a, const T*, const StepShared<T>* sh) {
    const S p_total = a[0] + a[1] + S(sh->load_t) + S(sh->pv_t);
    const S cost = primal(p_total) >= T(0) ? p_total * S(sh->price_t) : p_total * C<S>(0.08);
    const S dt = s[1] - C<S>(21.0);
    return -cost - (dt * dt) * C<S>(100.0);
  }
};

constexpr int SEEKER_MAXK = 3;   // obstacles carried in aux: goal (2) then (cx, cy, r) per obstacle

template <> struct Task<ENV_NAVIGATE_SEEKER> {
  using Sim = SeekerSim;
  // observation written by the kernel: (state, goal); the constant obstacle columns are appended by the host
  static constexpr int NS = 2, NA = 2, NO = 4, NAUX = 2 + 3 * SEEKER_MAXK;
  static constexpr bool CLAMP = true;
  static constexpr bool LIFTED_OK = false;   // wide safe-state generator -> per-env interior point (sgbpo_lifted.cuh)
  template <class S, class T> SG_HD static void observe(const S* s, const T* aux, const StepShared<T>*, S* o) {
    o[0] = s[0]; o[1] = s[1]; o[2] = S(aux[0]); o[3] = S(aux[1]);
  }
  template <class S, class T> SG_HD static S reward(const S* s, const S*, const T* aux, const StepShared<T>*) {
    const S dx = S(aux[0]) - s[0], dz = S(aux[1]) - s[1];
    const S q = dx * dx + dz * dz;
    if (primal(q) == T(0)) return C<S>(0.0);
    return -(C<S>(2.0) * sg_sqrt(q));
  }
};

// NavigateQuadrotor (envs/navigate_quadrotor.py): quadrotor dynamics, goal + up to SEEKER_MAXK circular obstacles.
// aux = goal (2), (cx, cz, r) per obstacle (9), initial goal distance (1), reached flag 0/1 (1).  The reward is stateful in
// the reference (`reached`, navigate_quadrotor.py:229-231): the kernel reads the flag from aux and reports a fresh arrival
// through the flag word (FL_REACHED), the host keeps the flag.
constexpr int NAVQUAD_NAUX = 2 + 3 * SEEKER_MAXK + 2;
template <> struct Task<ENV_NAVIGATE_QUADROTOR> {
  using Sim = QuadrotorSim;
  static constexpr int NS = 6, NA = 2, NO = 8, NAUX = NAVQUAD_NAUX;
  static constexpr bool CLAMP = true;
  static constexpr bool LIFTED_OK = false;
  template <class S, class T> SG_HD static void observe(const S* s, const T* aux, const StepShared<T>*, S* o) {
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = s[i];
    o[6] = S(aux[0]); o[7] = S(aux[1]);
  }
  // GOAL_REWARD * just_reached - goal_distance / initial_goal_distance  (the collision penalty is added by the step, which
  // knows whether the free step entered an obstacle: navigate_quadrotor.py:218-236)
  template <class S, class T> SG_HD static S reward(const S* s, const S*, const T* aux, const StepShared<T>*) {
    const S dx = S(aux[0]) - s[0], dz = S(aux[1]) - s[1];
    const S q = dx * dx + dz * dz;
    const S dist = primal(q) == T(0) ? C<S>(0.0) : sg_sqrt(q);
    const bool just_reached = aux[NAVQUAD_NAUX - 1] == T(0) && primal(dist) < T(0.1);
    return (just_reached ? C<S>(100.0) : C<S>(0.0)) - dist / S(aux[NAVQUAD_NAUX - 2]);
  }
};

// ---- linearisation around (s, a = 0, w = noise centre): c_f and B = df/da by an inner dual ------
template <class TASK, class S, class T>
SG_HD void linearise_cB(const S* s, const T* noise_center, S* c_f, S (*B)[MAXM]) {
  constexpr int NS = TASK::NS, NA = TASK::NA;
  using D = Dual<S, NA>;
  D si[NS], ai[NA], wi[NS], out[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) { si[i] = D::lift(s[i]); wi[i] = D::lift(S(noise_center[i])); }
#pragma unroll
  for (int k = 0; k < NA; ++k) ai[k] = D::seed(S(T(0)), k);
  TASK::Sim::f(si, ai, wi, out);
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    c_f[i] = out[i].v;
#pragma unroll
    for (int k = 0; k < NA; ++k) B[i][k] = out[i].d[k];
  }
}

// full (c, A, B, E) for the public ``linear_dynamics()`` API
template <class TASK, class S, class T>
SG_HD void linearise_full(const S* s, const T* noise_center, S* c_f, S* A, S* Bm, S* E) {
  constexpr int NS = TASK::NS, NA = TASK::NA, K = 2 * NS + NA;
  using D = Dual<S, K>;
  D si[NS], ai[NA], wi[NS], out[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) { si[i] = D::seed(s[i], i); wi[i] = D::seed(S(noise_center[i]), NS + NA + i); }
#pragma unroll
  for (int k = 0; k < NA; ++k) ai[k] = D::seed(S(T(0)), NS + k);
  TASK::Sim::f(si, ai, wi, out);
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    c_f[i] = out[i].v;
#pragma unroll
    for (int j = 0; j < NS; ++j) { A[i * NS + j] = out[i].d[j]; E[i * NS + j] = out[i].d[NS + NA + j]; }
#pragma unroll
    for (int k = 0; k < NA; ++k) Bm[i * NA + k] = out[i].d[NS + k];
  }
}

}  // namespace sgbpo
