// TEST INFRASTRUCTURE ONLY: compiles the scalar-generic per-environment math of
// safegbpo_b200/csrc/*.cuh with the HOST compiler so the device code's arithmetic (dynamics, dual-number
// Jacobians, gate, projection / ray-mask maps) can be checked against the oracle on a GPU-less machine.
// Never loaded by the product package.
#include <cstdint>
#include <cstring>
#include "../../include/sgbpo.h"
#include "../../safegbpo_b200/csrc/sgbpo_step.cuh"

using namespace sgbpo;

template <class T> static SafeConsts<T> conv(const sgbpo_consts& c) {
  SafeConsts<T> k;
  memset(&k, 0, sizeof(k));
  k.sg_kind = c.sg_kind; k.gate_mode = c.gate_mode;
  k.rm_linear = c.rm_linear; k.rm_zonotopic = c.rm_zonotopic; k.rm_passthrough = c.rm_passthrough;
  k.action_constrained = c.action_constrained; k.state_constrained = c.state_constrained; k.state_model = c.state_model; k.disable_clamp = c.disable_clamp; k.n_obstacles = c.n_obstacles; k.act_set_mode = c.act_set_mode;
  k.infeasible_raw = c.infeasible_raw;
  for (int i = 0; i < MAXN; ++i) {
    k.box_min[i] = T(c.box_min[i]); k.box_max[i] = T(c.box_max[i]); k.noise_center[i] = T(c.noise_center[i]);
    k.cs_hat[i] = T(c.cs_hat[i]); k.rb[i] = T(c.rb[i]); k.c_s[i] = T(c.c_s[i]);
  }
  for (int i = 0; i < MAXN * MAXN; ++i) k.Ginv[i] = T(c.Ginv[i]);
  for (int i = 0; i < MAXN * MAXM; ++i) { k.B0hat_abs[i] = T(c.B0hat_abs[i]); k.B0hat[i] = T(c.B0hat[i]); }
  k.n_gen = c.n_gen; k.n_null = c.n_null; k.n_wcol = c.n_wcol;
  for (int i = 0; i < MAXGEN; ++i) {
    for (int j = 0; j < MAXN; ++j) k.Gpinv[i][j] = T(c.Gpinv[i][j]);
    for (int j = 0; j < MAXNULL; ++j) k.Nmat[i][j] = T(c.Nmat[i][j]);
    for (int j = 0; j < 2; ++j) k.Gam_p[i][j] = T(c.Gam_p[i][j]);
  }
  for (int i = 0; i < MAXM; ++i) { k.a_lo[i] = T(c.a_lo[i]); k.a_hi[i] = T(c.a_hi[i]); k.c_a[i] = T(c.c_a[i]); }
  k.n_act_hp = c.n_act_hp; k.n_k_hp = c.n_k_hp; k.n_q_hp = c.n_q_hp;
  for (int r = 0; r < MAXHP; ++r) {
    for (int j = 0; j < 3; ++j) { k.act_hp[r][j] = T(c.act_hp[r][j]); k.q_hp[r][j] = T(c.q_hp[r][j]); }
    for (int j = 0; j < 2; ++j) k.k_hp[r][j] = T(c.k_hp[r][j]);
  }
  return k;
}
template <class T> static StepShared<T> convs(const sgbpo_step_shared* s) {
  StepShared<T> o;
  memset(&o, 0, sizeof(o));
  if (s) { for (int i = 0; i < 20; ++i) o.series[i] = T(s->series[i]); o.load_t = T(s->load_t); o.pv_t = T(s->pv_t); o.price_t = T(s->price_t); }
  return o;
}

template <int ENV, class T>
static void run(int64_t N, const sgbpo_consts* c, const sgbpo_step_shared* shp, const T* s, const T* a, const T* w,
                const T* aux, const T* dirs, const T* reset_state, T* s_next, T* a_exec, T* obs, T* reward,
                int32_t* flags, const T* g_s_next, const T* g_a_exec, const T* g_obs, const T* g_reward, T* gs, T* ga) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX, K = NS + NA;
  const SafeConsts<T> k = conv<T>(*c);
  const StepShared<T> sh = convs<T>(shp);
  for (int64_t i = 0; i < N; ++i) {
    const T* auxp = NAUX ? aux + i * NAUX : nullptr;
    const T* dp = dirs ? dirs + i * NA * 2 * NA : nullptr;
    const T* rs = reset_state ? reset_state + i * NS : nullptr;
    if (s_next) {
      T sn[NS], ae[NA], ob[NO], rew; int fl;
      safeguarded_step<TASK, T, T>(s + i * NS, a + i * NA, w + i * NS, auxp, dp, &sh, k, rs != nullptr, rs, sn, ae, ob, rew, fl);
      for (int q = 0; q < NS; ++q) s_next[i * NS + q] = sn[q];
      for (int q = 0; q < NA; ++q) a_exec[i * NA + q] = ae[q];
      for (int q = 0; q < NO; ++q) obs[i * NO + q] = ob[q];
      reward[i] = rew; flags[i] = fl;
    }
    if (gs) {
      using D = Dual<T, K>;
      D sd[NS], ad[NA], sn[NS], ae[NA], ob[NO], rew; int fl;
      for (int q = 0; q < NS; ++q) sd[q] = D::seed(s[i * NS + q], q);
      for (int q = 0; q < NA; ++q) ad[q] = D::seed(a[i * NA + q], NS + q);
      safeguarded_step<TASK, D, T>(sd, ad, w + i * NS, auxp, dp, &sh, k, rs != nullptr, rs, sn, ae, ob, rew, fl);
      T acc[K];
      for (int d = 0; d < K; ++d) acc[d] = 0;
      for (int d = 0; d < K; ++d) {
        if (g_s_next) for (int q = 0; q < NS; ++q) acc[d] += g_s_next[i * NS + q] * sn[q].d[d];
        if (g_a_exec) for (int q = 0; q < NA; ++q) acc[d] += g_a_exec[i * NA + q] * ae[q].d[d];
        if (g_obs) for (int q = 0; q < NO; ++q) acc[d] += g_obs[i * NO + q] * ob[q].d[d];
        if (g_reward) acc[d] += g_reward[i] * rew.d[d];
      }
      for (int q = 0; q < NS; ++q) gs[i * NS + q] = acc[q];
      for (int q = 0; q < NA; ++q) ga[i * NA + q] = acc[NS + q];
    }
  }
}

template <int ENV, class T>
static void lin(int64_t N, const sgbpo_consts* c, const T* s, T* cc, T* A, T* B, T* E) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA;
  const SafeConsts<T> k = conv<T>(*c);
  for (int64_t i = 0; i < N; ++i)
    linearise_full<TASK, T, T>(s + i * NS, k.noise_center, cc + i * NS, A + i * NS * NS, B + i * NS * NA, E + i * NS * NS);
}

#define DISPATCH(FN, ...)                                                        \
  switch (env_id) {                                                              \
    case 0: FN<0, T>(__VA_ARGS__); break; case 1: FN<1, T>(__VA_ARGS__); break;  \
    case 2: FN<2, T>(__VA_ARGS__); break; case 3: FN<3, T>(__VA_ARGS__); break;  \
    case 4: FN<4, T>(__VA_ARGS__); break; case 5: FN<5, T>(__VA_ARGS__); break;  \
    case 6: FN<6, T>(__VA_ARGS__); break;                                        \
    default: return -1; }

template <class T>
static int step_t(int env_id, int64_t N, const sgbpo_consts* c, const sgbpo_step_shared* sh, const void* s, const void* a,
                  const void* w, const void* aux, const void* dirs, const void* rs, void* s_next, void* a_exec, void* obs,
                  void* reward, int32_t* flags, const void* g1, const void* g2, const void* g3, const void* g4, void* gs, void* ga) {
  DISPATCH(run, N, c, sh, (const T*)s, (const T*)a, (const T*)w, (const T*)aux, (const T*)dirs, (const T*)rs, (T*)s_next,
           (T*)a_exec, (T*)obs, (T*)reward, flags, (const T*)g1, (const T*)g2, (const T*)g3, (const T*)g4, (T*)gs, (T*)ga)
  return 0;
}
template <class T>
static int lin_t(int env_id, int64_t N, const sgbpo_consts* c, const void* s, void* cc, void* A, void* B, void* E) {
  DISPATCH(lin, N, c, (const T*)s, (T*)cc, (T*)A, (T*)B, (T*)E)
  return 0;
}

template <class T>
static void navquad_set_t(int64_t N, const sgbpo_consts* c, const T* s, const T* aux, T* center, T* gen, int32_t* status) {
  using TASK = Task<ENV_NAVIGATE_QUADROTOR>;
  const SafeConsts<T> k = conv<T>(*c);
  for (int64_t i = 0; i < N; ++i) {
    T cf[6], B[6][MAXM];
    linearise_cB<TASK, T, T>(s + i * 6, k.noise_center, cf, B);
    StateDyn<T> sd;
    navquad_safe_state_set<T>(cf, B, aux + i * NAVQUAD_NAUX, k, sd);
    for (int q = 0; q < 6; ++q) center[i * 6 + q] = sd.center[q];
    for (int q = 0; q < 36; ++q) gen[i * 36 + q] = sd.gen[q];
    status[i] = sd.degenerate ? 1 : 0;
  }
}

extern "C" {
int hm_navquad_set(int dtype, int64_t N, const sgbpo_consts* c, const void* s, const void* aux, void* center, void* gen, int32_t* status) {
  if (dtype) navquad_set_t<double>(N, c, (const double*)s, (const double*)aux, (double*)center, (double*)gen, status);
  else navquad_set_t<float>(N, c, (const float*)s, (const float*)aux, (float*)center, (float*)gen, status);
  return 0;
}
int hm_step(int env_id, int dtype, int64_t N, const sgbpo_consts* c, const sgbpo_step_shared* sh, const void* s,
            const void* a, const void* w, const void* aux, const void* dirs, const void* rs, void* s_next, void* a_exec,
            void* obs, void* reward, int32_t* flags, const void* g_s_next, const void* g_a_exec, const void* g_obs,
            const void* g_reward, void* gs, void* ga) {
  return dtype ? step_t<double>(env_id, N, c, sh, s, a, w, aux, dirs, rs, s_next, a_exec, obs, reward, flags, g_s_next, g_a_exec, g_obs, g_reward, gs, ga)
               : step_t<float>(env_id, N, c, sh, s, a, w, aux, dirs, rs, s_next, a_exec, obs, reward, flags, g_s_next, g_a_exec, g_obs, g_reward, gs, ga);
}
// wrap_angle (sgbpo_dual.cuh): value and derivative (one dual direction) of n angles
int hm_wrap_angle(int dtype, int64_t n, const void* x, void* y, void* dy) {
  for (int64_t i = 0; i < n; ++i) {
    if (dtype) {
      Dual<double, 1> d = Dual<double, 1>::seed(((const double*)x)[i], 0);
      const Dual<double, 1> r = wrap_angle(d);
      ((double*)y)[i] = r.v; ((double*)dy)[i] = r.d[0];
    } else {
      Dual<float, 1> d = Dual<float, 1>::seed(((const float*)x)[i], 0);
      const Dual<float, 1> r = wrap_angle(d);
      ((float*)y)[i] = r.v; ((float*)dy)[i] = r.d[0];
    }
  }
  return 0;
}
int hm_linearise(int env_id, int dtype, int64_t N, const sgbpo_consts* c, const void* s, void* cc, void* A, void* B, void* E) {
  return dtype ? lin_t<double>(env_id, N, c, s, cc, A, B, E) : lin_t<float>(env_id, N, c, s, cc, A, B, E);
}
}
