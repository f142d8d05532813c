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
#define SGBPO_MAXGEN 20
#define SGBPO_MAXNULL 14

enum { SGBPO_F32 = 0, SGBPO_F64 = 1 };
enum { SGBPO_OK = 0, SGBPO_E_ARG = -1, SGBPO_E_CUDA = -2, SGBPO_E_UNSUPPORTED = -3 };

/* environment ids (envs/*.py) */
enum {
  SGBPO_ENV_BALANCE_PENDULUM = 0, SGBPO_ENV_BALANCE_CARTPOLE = 1, SGBPO_ENV_SWINGUP_CARTPOLE = 2,
  SGBPO_ENV_BALANCE_QUADROTOR = 3, SGBPO_ENV_MANAGE_HOUSEHOLD = 4, SGBPO_ENV_NAVIGATE_SEEKER = 5,
  SGBPO_ENV_NAVIGATE_QUADROTOR = 6
};
/* safeguard kinds (safeguards/*.py), gate modes (safeguard.py:63-67) and per-env flag bits */
enum { SGBPO_SG_NONE = 0, SGBPO_SG_BOUNDARY_PROJECTION = 1, SGBPO_SG_RAY_MASK = 2 };
enum { SGBPO_GATE_REFERENCE = 0, SGBPO_GATE_INTENDED = 1, SGBPO_GATE_ALWAYS = 2 };
enum {
  SGBPO_FL_PROJECTABLE = 1, SGBPO_FL_INFEASIBLE = 2, SGBPO_FL_INTERVENED = 4, SGBPO_FL_NONFINITE = 8,
  SGBPO_FL_GUARD_USED = 16, SGBPO_FL_ACTION_IN_SET = 32, SGBPO_FL_NEXT_IN_SET = 64, SGBPO_FL_CLAMPED = 128,
  SGBPO_FL_COLLIDED = 256, SGBPO_FL_UNREDUCED = 512, SGBPO_FL_REACHED = 1024
};

/* Constant set data, registered once per wrapper (replaces the numpy constants cvxpy bakes into the
 * problems at first call: safeguard.py:111-122, ray_mask.py:157-171).  All doubles, host memory. */
typedef struct sgbpo_consts {
  int32_t sg_kind, gate_mode;
  int32_t rm_linear, rm_zonotopic, rm_passthrough;
  int32_t action_constrained, state_constrained, state_model; /* 0: G_s invertible, 1: polygon (n = 2), 2: unreduced (gate + raw action only), 3: lifted (per-env interior point), 4: rebuilt per step (NavigateQuadrotor; G_w in Gam_p) */
  int32_t disable_clamp, n_obstacles;                        /* 1: raw dynamics (no clamp); obstacles in aux (Navigate*) */
  int32_t act_set_mode;                                      /* 0: constant safe action set; 1: NavigateSeeker per-step set (P5) */
  int32_t infeasible_raw;                                    /* action executed where the guarded program is infeasible (always flagged
                                                              * FL_INFEASIBLE): 0 = NaN (the reference's solver raises there), 1 = the raw
                                                              * policy action, identity gradient (non-strict rollouts stay finite) */
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
  double B0hat[SGBPO_MAXN * SGBPO_MAXM];                     /* G_s^-1 B0 (signed) */
  /* state block, model 3 (G_s wide, e.g. BalanceQuadrotor 6 x 24): the containment encoding is kept in its lifted
   * form and solved per environment (interior point).  Generators with equal direction merged, zero ones dropped:
   * beta = Gpinv d + Nmat y,  Gamma_j = Gam_p[:, j] + Nmat z_j  parametrise  G beta = d,  G Gamma = G_w. */
  int32_t n_gen, n_null, n_wcol, reserved2;
  double Gpinv[SGBPO_MAXGEN][SGBPO_MAXN];                    /* pinv(G)            (n_gen x n) */
  double Nmat[SGBPO_MAXGEN][SGBPO_MAXNULL];                  /* null space of G    (n_gen x n_null), orthonormal columns */
  double Gam_p[SGBPO_MAXGEN][2];                             /* pinv(G) G_w        (n_gen x n_wcol) */
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

/* K11: NavigateSeekerEnv.safe_action_set / compute_generator (envs/navigate_seeker.py:398-477, program P5):
 * generator (N,2,2) of the per-environment safe action zonotope Z(0, U diag(len)); aux (N, 2 + 3*3) holds
 * goal then (cx, cy, r) per obstacle; status (N) int32: 0 ok, 1 degenerate/infeasible. */
int sgbpo_seeker_action_set(int dtype, int64_t N, const sgbpo_consts* consts, const void* s, const void* aux,
                            void* generator, int32_t* status, void* stream);

/* NavigateQuadrotorEnv.safe_state_set (envs/navigate_quadrotor.py:436-779, programs P6-P8, no_grad): the per-environment
 * safe state zonotope of the CURRENT state: center (N,6), generator (N,6,6) = [position | velocity | roll] blocks;
 * aux (N, 2 + 3*3 + 2): goal, (cx, cz, r) per obstacle, initial goal distance, reached flag; status (N) int32: 0 ok,
 * 1 degenerate (the reference's all-rows fallback blocks / an infeasible support-point program). */
int sgbpo_navquad_safe_state_set(int dtype, int64_t N, const sgbpo_consts* consts, const void* s, const void* aux,
                                 void* center, void* generator, int32_t* status, void* stream);

/* ---- whole-horizon fused rollout (SHAC.collect_trajectories + reverse pass, shac.py:96-151) ------------ */

/* MLP parameters on the DEVICE in the kernels' layout (policy.py:50-60 / value_function.py:38-48:
 * Linear -> ELU -> LayerNorm per hidden layer, then Linear).  All hidden layers share one width. */
typedef struct sgbpo_mlp {
  int32_t in_dim, width, n_hidden, out_dim;
  /* parameters exactly as torch stores them (no repacking): Linear.weight of layer l is [out_l][in_l]
   * row-major; l = 0..n_hidden-1 are the hidden layers, l = n_hidden is the output layer */
  const void* weight[5];
  const void* bias[5];
  const void* ln_weight[4];  /* LayerNorm weight / bias of hidden layer l: [width] */
  const void* ln_bias[4];
  const void* log_std;       /* [out_dim] policy only, else NULL */
} sgbpo_mlp;

/* Rollout buffers, time-major like CoupledBuffer's (T, N, .) tensors (coupled_buffer.py:79-92). */
typedef struct sgbpo_rollout {
  int64_t N;
  int32_t H;
  int32_t engine;            /* bit 0, forward: 0 = FMA register tiles (float32 / float64, widths 32-128), 1 = tcgen05 kernel;
                              * bit 2, forward: warp-MMA kernel (mma.sync on register fragments, no block-wide barrier in the time
                              * loop; float32 policies [in -> 128 -> 128 -> m] of the first four environments; exclusive with bit 0);
                              * bit 1, reverse pass: 0 = sequential FMA kernel, 1 = parallel tensor-core reverse pass
                              * (row-batched policy backward + per-environment scan).  The tensor-core engines cover float32
                              * policies [in -> 128 -> 128 -> m]: sgbpo_rollout_engine_supported.  Both forward engines
                              * write the same activation stash, so any combination is valid. */
  const void* eps;           /* [H][N][m]   policy draws (Normal.rsample, policy.py:84-85) */
  const void* noise;         /* [H][N][n]   noise_set.sample() per step (simulator.py:181) */
  const void* dirs;          /* [H][N][m][2m] ray-mask directions (ray_mask.py:184-185) or NULL */
  const void* aux0;          /* [N][naux] or NULL */
  const void* reset_states;  /* [R][N][n] states drawn by reset() inside the horizon (Q6) or NULL */
  const int32_t* reset_idx;  /* [H] device: -1 or r: s_{t+1} := reset_states[r] */
  const int32_t* aux_src;    /* [H] device: -1 -> aux0, r -> reset_states[r][:, :naux] in effect after step t */
  const void* shared;        /* [H][23] per-step shared scalars (layout of sgbpo_step_shared in dtype) or NULL */
  void* states;              /* [H+1][N][n]  row 0 = s_0 (input) */
  void* obs;                 /* [H+2][N][o]  row 0 = obs_0 (input) */
  void* actions;             /* [H+1][N][m] */
  void* exec_actions;        /* [H][N][m] */
  void* rewards;             /* [H+2][N] */
  int32_t* flags;            /* [H][N] */
  /* activation stash, written by sgbpo_rollout_fwd and read by sgbpo_rollout_bwd (policies with two hidden
   * layers; all four NULL = forward only).  Recomputing the policy forward in the reverse pass would cost two
   * of its three GEMMs; 1 KB per environment step of HBM traffic is far cheaper on this path. */
  void* stash_h1;            /* [H][N][width]  LayerNorm output of hidden layer 1 = operand of Linear2.weight.grad */
  void* stash_e1;            /* [H][N][width]  ELU output of hidden layer 1 */
  void* stash_e2;            /* [H][N][width]  ELU output of hidden layer 2.  With reverse engine 1 (engine bit 1) both ELU stashes
                              * are stored in 128-row tiles of 16-byte column chunks instead -- element (row = t N + env, col) at
                              * ((row / 128 * width / 4 + col / 4) * 128 + row % 128) * 4 + col % 4 -- so that the row-batched
                              * reverse kernels read them with contiguous 512-byte requests and one 64 KB TMA bulk copy per
                              * tile; the buffers must then hold ceil(H N / 128) * 128 rows. */
  void* stash_stats;         /* [H][N][4]      LayerNorm mean1, rstd1, mean2, rstd2 */
  const void* reset_aux;     /* [R][N][naux] aux (goal, obstacles) in effect after reset r, or NULL: reset_states[r][:, :naux]
                              * (forward engine 1 and sgbpo_rollout_jac) */
  /* reverse engine 1 only (stash_h1 is not used by it and may be NULL) */
  void* jac;                 /* [H][N][sgbpo_rollout_jac_stride(env)] step Jacobians d(s', obs', r)/d(s, a): written by
                              * sgbpo_rollout_jac, read by sgbpo_rollout_bwd */
  void* pol_jac;             /* [H][N][sgbpo_rollout_pol_jac_stride(env)] scratch: policy Jacobians d mu / d obs per row */
  void* dmu;                 /* [H][N][m] scratch: d loss / d mu_t from the scan */
} sgbpo_rollout;

typedef struct sgbpo_rollout_grads {
  const void* grw;             /* [H]  d loss / d rewards[t] */
  const void* gobs_term;       /* [K][N][o] d(terminal value terms)/d obs rows, or NULL */
  const int32_t* term_of_row;  /* [H+2] device: -1 or k */
  void* dz2_out;               /* [H][N][width]  Linear2.weight.grad = dz2^T stash_h1 (one GEMM by the caller) */
  /* parameter gradients, accumulated (caller zero-fills), in torch's layouts:
   * g_w_in [width][in_dim], g_w_out [out_dim][width], g_b_hid/g_gamma/g_beta [2][width], g_b_out/g_log_std [out_dim] */
  void* g_w_in; void* g_w_out; void* g_b_hid; void* g_gamma; void* g_beta; void* g_b_out; void* g_log_std;
  void* g_w_hid;               /* reverse engine 1: [width][width] Linear2.weight.grad accumulated in-kernel (dz2_out unused);
                                * 16-byte aligned (flushed with vector reductions) */
} sgbpo_rollout_grads;

/* forward over the whole horizon: policy MLP + safeguarded step per environment per t (one launch) */
int sgbpo_rollout_fwd(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_mlp* policy,
                      const sgbpo_rollout* r, void* stream);
/* reverse pass over the whole horizon (replaces autograd's H-long chain) */
int sgbpo_rollout_bwd(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_mlp* policy,
                      const sgbpo_rollout* r, const sgbpo_rollout_grads* g, void* stream);
/* engine 1, between forward and reverse pass: the dual-number Jacobian of every safeguarded step (t, environment)
 * w.r.t. (s_t, a_t) from the stored (s_t, a_t, w_t) -- what autograd's per-step backward nodes hold, but
 * evaluated for all H x N steps in ONE fully parallel launch instead of on the sequential reverse chain. */
int sgbpo_rollout_jac(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_rollout* r, void* stream);
int sgbpo_rollout_jac_stride(int env_id);     /* floats per (t, environment) record of sgbpo_rollout.jac; < 0: unknown env */
int sgbpo_rollout_pol_jac_stride(int env_id); /* same for sgbpo_rollout.pol_jac */
/* 1 when engine 1 covers (env_id, dtype, policy), else 0 */
int sgbpo_rollout_engine_supported(int env_id, int dtype, const sgbpo_mlp* policy);
/* batched MLP forward over M rows (target critic over the buffer's observation rows, shac.py:111) */
int sgbpo_mlp_rows(int dtype, int64_t M, const sgbpo_mlp* net, const void* x, void* out, void* stream);

/* input gradient of a scalar-output MLP over M rows: gx[row] = row_weight[row] * d net(x[row]) / d x[row]  ([M][in_dim]).
 * The terminal-value terms of the SHAC loss (shac.py:137-147) need it as the cotangent of the terminal observation rows. */
int sgbpo_mlp_rows_vjp(int dtype, int64_t M, const sgbpo_mlp* net, const void* x, const void* row_weight, void* gx, void* stream);

/* ---- policy update tail (shac.py:148-149: clip_grad_norm_(policy, max_norm) + torch.optim.Adam.step()) ---- */
#define SGBPO_OPT_MAX_TENSORS 24
typedef struct sgbpo_optim {
  int32_t n_tensors, reserved;
  void* param[SGBPO_OPT_MAX_TENSORS];       /* parameter tensors (updated in place) */
  void* grad[SGBPO_OPT_MAX_TENSORS];        /* their gradients (scaled in place by the clip coefficient, like torch) */
  void* exp_avg[SGBPO_OPT_MAX_TENSORS];     /* torch.optim.Adam state tensors, updated in place */
  void* exp_avg_sq[SGBPO_OPT_MAX_TENSORS];
  void* step[SGBPO_OPT_MAX_TENSORS];        /* 0-dim device step counters (capturable=True), dtype step_dtype */
  int64_t numel[SGBPO_OPT_MAX_TENSORS];
  double lr, beta1, beta2, eps, max_norm;
} sgbpo_optim;
/* one launch: global gradient norm -> clip -> Adam (no weight decay / amsgrad); total_norm_out: 1 element of dtype or NULL */
int sgbpo_clip_adam(int dtype, int step_dtype, const sgbpo_optim* o, void* total_norm_out, void* stream);

/* ---- critic targets (shac.py:187-241, calculate_estimated_values): backward TD-lambda scan per environment ----
 * rewards / values: [H+2][N] (time-major, like CoupledBuffer); terminals: [H+2][N] bytes (torch.bool storage);
 * est: [H][N] targets, valid: [H][N] bytes = 1 where row t of environment e is not terminal (a target exists). */
int sgbpo_td_lambda(int dtype, int64_t N, int H, const void* rewards, const void* values, const uint8_t* terminals,
                    double gamma, double td_weight, void* est, uint8_t* valid, void* stream);

/* ---- critic fit (shac.py:154-184): hidden layer tail  h = LayerNorm(ELU(z + bias))  on (M, width) batches, one
 * pass forward and one backward.  fwd keeps the ELU output e (M, width) and (mean, rstd) per row in stats (M, 2);
 * bwd returns dz (gradient w.r.t. z) and ACCUMULATES the column sums d gamma, d beta, d bias (caller zero-fills). */
int sgbpo_bias_elu_ln_fwd(int dtype, int64_t M, int width, const void* z, const void* bias, const void* gamma,
                          const void* beta, void* h, void* e_out, void* stats, void* stream);
int sgbpo_bias_elu_ln_bwd(int dtype, int64_t M, int width, const void* gy, const void* e, const void* stats,
                          const void* gamma, void* dz, void* dgamma, void* dbeta, void* dbias, void* stream);

/* ---- episode scalars of collect_trajectories / update_policy (shac.py:119-120, 131-142; safeguard.py:68-80), one
 * single-CTA launch each.  stats: scal[0] = sum of the n_rewards reward entries, scal[1] = 1 if any flag word has a bit of
 * bad_mask, *n_interv = number of flag words with a bit of interv_mask, *bad = scal[1] as int (both optional).
 * loss: -sum_t disc[t] * sum_n (term_int[t] ? values[t][n] : rewards[t][n]) / norm[0] over `rows` buffer rows. */
int sgbpo_episode_stats(int dtype, int64_t n_rewards, const void* rewards, int64_t n_flags, const int32_t* flags,
                        int interv_mask, int bad_mask, double* scal, int64_t* n_interv, int32_t* bad, void* stream);
int sgbpo_shac_loss(int dtype, int rows, int64_t N, const void* disc, const int32_t* term_int, const void* values,
                    const void* rewards, const void* norm, void* loss, void* stream);

/* ---- random inputs of one episode in one launch (fused_rollout.py, rng_order = "batched"): eps [H][N][m] ~ N(0,1)
 * (policy.py:84-85), noise [H][N][n] = lo + span * u (box.py:73-74), dirs [H][N][m][2m] unit columns (ray_mask.py:184-185) or
 * NULL, reset_states [n_resets][N][n] = reset_center + reset_gen (2u - 1) (simulator.py:73 + the task's reset distribution).
 * Counter-based Philox4x32-10 keyed by (seed, *episode, element): the same numbers whatever the launch shape. */
#define SGBPO_MAX_RESET_GENS 24
typedef struct sgbpo_episode_inputs_args {
  int64_t N;
  int32_t H, n, m, n_resets, n_reset_gens, want_dirs;
  uint64_t seed;
  const void* episode;       /* device int64: episode counter */
  void* eps; void* noise; void* dirs; void* reset_states;
  double noise_lo[SGBPO_MAXN], noise_span[SGBPO_MAXN];
  double reset_center[SGBPO_MAXN], reset_gen[SGBPO_MAXN][SGBPO_MAX_RESET_GENS];
} sgbpo_episode_inputs_args;
int sgbpo_episode_inputs(int dtype, const sgbpo_episode_inputs_args* args, void* stream);

/* ---- containment verdicts per environment (replaces the host LPs of src/sets/zonotope.py:143-239 -- programs P9 -- and
 * src/sets/box.py:90-102).  The outer set is passed as the K facets  h_k . d <= 1  of the verdict program's feasible region in the
 * offset d = point - outer_center (safegbpo_b200/sets/verdicts.py enumerates them once per outer set on the host):
 *   point in zonotope            value_i = max_k h_k . d_i                       = the LP value  min |gamma|_inf  (zonotope gauge)
 *   zonotope (fixed generators)  value_i = max_k h_k . d_i  (facets of the LP's feasible region)   value <= 1 <=> LP value <= 1
 *   zonotope (per-env generators, inner_gen != NULL, N x n x q row-major)
 *                                value_i = max_k [h_k . d_i + sum_j |h_k . g_ij|]   exact containment in the outer polytope
 * verdict_i = value_i <= 1 + tol (NaN -> 0).  outer_center is n values, or N x n when center_per_env != 0. */
typedef struct sgbpo_contains_args {
  int64_t N;
  int32_t n, K, q, center_per_env;
  const void* facets;        /* K x n */
  const void* outer_center;
  const void* points;        /* N x n: the point, or the centre of the inner zonotope */
  const void* inner_gen;     /* NULL or N x n x q */
  void* value;               /* N (may be NULL) */
  int32_t* verdict;          /* N (may be NULL) */
  double tol;
} sgbpo_contains_args;
int sgbpo_contains(int dtype, const sgbpo_contains_args* args, void* stream);

/* ---- tensor-core operand-view self-test (diagnostic; tests/test_gpu_tc_mma.py): X, Z, D are 128 x 128 float32 device
 * matrices; mode 0: D = X Z^T, 1: D = X Z, 2: D = X^T Z, each through the fp16 hi/lo split tcgen05 path the rollout
 * kernels use (K-major / MN-major shared-memory views of one stored operand). */
int sgbpo_tc_selftest(int mode, const void* x, const void* z, void* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
