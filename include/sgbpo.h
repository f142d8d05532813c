/* sgbpo.h -- C ABI of the B200-native safeguarded-rollout library (libsgbpo.so).
 *
 * The reference (TimWalter/SafeGBPO) has no FFI: its boundary is the Python class contract
 * (SURVEY.md §8b).  These entry points are what that contract's hot methods bind to; each one names the
 * reference code it replaces (paths relative to /root/reference/src).  Conventions: plain pointers and
 * sizes; every data pointer is a DEVICE pointer owned by the caller (torch); nothing is allocated
 * inside; `stream` is a cudaStream_t passed as void*; the return value is 0 on success or a negative
 * SGBPO_E_* code, with a message retrievable from sgbpo_last_error(); no exception crosses the ABI.
 * dtype: 0 = float32, 1 = float64 (the reference runs float64 by default, main.py:14).
 */
#ifndef SGBPO_H
#define SGBPO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGBPO_MAXN 6
#define SGBPO_MAXM 2
#define SGBPO_MAXHP 48

enum { SGBPO_F32 = 0, SGBPO_F64 = 1 };
enum { SGBPO_OK = 0, SGBPO_E_ARG = -1, SGBPO_E_CUDA = -2, SGBPO_E_UNSUPPORTED = -3 };

/* environment ids (envs/*.py) */
enum {
  SGBPO_ENV_BALANCE_PENDULUM = 0, SGBPO_ENV_BALANCE_CARTPOLE = 1, SGBPO_ENV_SWINGUP_CARTPOLE = 2,
  SGBPO_ENV_BALANCE_QUADROTOR = 3, SGBPO_ENV_MANAGE_HOUSEHOLD = 4, SGBPO_ENV_NAVIGATE_SEEKER = 5
};
/* safeguard kinds (safeguards/*.py), gate modes (safeguard.py:63-67) and per-env flag bits */
enum { SGBPO_SG_NONE = 0, SGBPO_SG_BOUNDARY_PROJECTION = 1, SGBPO_SG_RAY_MASK = 2 };
enum { SGBPO_GATE_REFERENCE = 0, SGBPO_GATE_INTENDED = 1, SGBPO_GATE_ALWAYS = 2 };
enum {
  SGBPO_FL_PROJECTABLE = 1, SGBPO_FL_INFEASIBLE = 2, SGBPO_FL_INTERVENED = 4, SGBPO_FL_NONFINITE = 8,
  SGBPO_FL_GUARD_USED = 16, SGBPO_FL_ACTION_IN_SET = 32, SGBPO_FL_NEXT_IN_SET = 64, SGBPO_FL_CLAMPED = 128
};

/* Constant set data, registered once per wrapper (replaces the numpy constants cvxpy bakes into the
 * problems at first call: safeguard.py:111-122, ray_mask.py:157-171).  All doubles, host memory. */
typedef struct sgbpo_consts {
  int32_t sg_kind, gate_mode;
  int32_t rm_linear, rm_zonotopic, rm_passthrough;
  int32_t action_constrained, state_constrained, state_model; /* 0: G_s invertible, 1: polygon (n = 2) */
  int32_t disable_clamp, reserved0;                          /* 1: raw dynamics (Simulator.dynamics), no box clamp */
  double box_min[SGBPO_MAXN], box_max[SGBPO_MAXN];           /* feasible state box */
  double noise_center[SGBPO_MAXN];
  double a_lo[SGBPO_MAXM], a_hi[SGBPO_MAXM];                 /* m = 1 safe action interval */
  int32_t n_act_hp;                                          /* m = 2 safe action polygon rows n.a <= h */
  double act_hp[SGBPO_MAXHP][3];
  double c_a[SGBPO_MAXM];
  double Ginv[SGBPO_MAXN * SGBPO_MAXN];                      /* G_s^-1, row-major n x n */
  double cs_hat[SGBPO_MAXN];                                 /* G_s^-1 c_s */
  double rb[SGBPO_MAXN];                                     /* 1 - row sums of |G_s^-1 G_w| */
  double B0hat_abs[SGBPO_MAXN * SGBPO_MAXM];                 /* |G_s^-1 B0| */
  int32_t n_k_hp, n_q_hp;
  double k_hp[SGBPO_MAXHP][2];                               /* polygon K: hp . d <= 1 */
  double q_hp[SGBPO_MAXHP][3];                               /* polytope Q: hq . (d, L) <= 1 */
  double c_s[SGBPO_MAXN];
} sgbpo_consts;

/* per-step scalars shared by all environments (household series; household.py:109-113,
 * manage_household.py:60-66).  Host memory, doubles. */
typedef struct sgbpo_step_shared {
  double series[20];
  double load_t, pv_t, price_t;
} sgbpo_step_shared;

const char* sgbpo_version(void);
const char* sgbpo_last_error(void);
/* static shape table: writes n (state), m (action), o (observation), naux for env_id; 0 on success */
int sgbpo_env_dims(int env_id, int* n, int* m, int* o, int* naux);

/* K1+K2+K3+K4..K9 fused: Safeguard.actions (safeguards/interfaces/safeguard.py:61-80) followed by
 * Simulator.step (envs/simulators/interfaces/simulator.py:82-108) for N environments.
 *   s (N,n) a (N,m) w (N,n) noise sample incl. centre; aux (N,naux) or NULL; dirs (N,m,2m) or NULL;
 *   reset_state (N,n) or NULL: when non-NULL the post-step state is REPLACED by it (Q6 reset-all);
 *   out: s_next (N,n) a_exec (N,m) obs (N,o) reward (N) flags (N) int32. */
int sgbpo_step_fwd(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const sgbpo_step_shared* shared,
                   const void* s, const void* a, const void* w, const void* aux, const void* dirs,
                   const void* reset_state, void* s_next, void* a_exec, void* obs, void* reward,
                   int32_t* flags, void* stream);

/* reverse pass of sgbpo_step_fwd (what autograd does today through vmap'd dynamics, clamp, where,
 * _CvxpyLayerFn.backward and the jacrev double-backward: SURVEY §3.3).  Recomputes the step with dual
 * numbers from (s, a, w).  Cotangents g_* may be NULL (= zero).  Writes gs (N,n), ga (N,m). */
int sgbpo_step_bwd(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const sgbpo_step_shared* shared,
                   const void* s, const void* a, const void* w, const void* aux, const void* dirs,
                   const void* reset_state, const void* g_s_next, const void* g_a_exec, const void* g_obs,
                   const void* g_reward, void* gs, void* ga, void* stream);

/* K2: Simulator.linear_dynamics (simulator.py:206-230): c (N,n), A (N,n,n), B (N,n,m), E (N,n,n). */
int sgbpo_linearise(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const void* s,
                    void* c, void* A, void* B, void* E, void* stream);

#ifdef __cplusplus
}
#endif
#endif
