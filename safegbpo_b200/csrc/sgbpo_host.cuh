// Host-side conversion of the C-ABI structs (include/sgbpo.h, doubles) into the kernels' typed constants.
#pragma once
#include <cstring>
#include "../../include/sgbpo.h"
#include "sgbpo_safeguard.cuh"

namespace sgbpo {

template <class T> inline SafeConsts<T> convert_consts(const sgbpo_consts& c) {
  SafeConsts<T> k;
  memset(&k, 0, sizeof(k));
  k.sg_kind = c.sg_kind; k.gate_mode = c.gate_mode;
  k.rm_linear = c.rm_linear; k.rm_zonotopic = c.rm_zonotopic; k.rm_passthrough = c.rm_passthrough;
  k.action_constrained = c.action_constrained; k.state_constrained = c.state_constrained;
  k.state_model = c.state_model; k.disable_clamp = c.disable_clamp;
  k.n_obstacles = c.n_obstacles; k.act_set_mode = c.act_set_mode; k.infeasible_raw = c.infeasible_raw;
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

template <class T> inline StepShared<T> convert_shared(const sgbpo_step_shared* s) {
  StepShared<T> o;
  memset(&o, 0, sizeof(o));
  if (s) {
    for (int i = 0; i < 20; ++i) o.series[i] = T(s->series[i]);
    o.load_t = T(s->load_t); o.pv_t = T(s->pv_t); o.price_t = T(s->price_t);
  }
  return o;
}


}  // namespace sgbpo
