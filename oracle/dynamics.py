"""Oracle simulators: float64 torch-CPU restatement of the reference's five simulators and the
seven task environments built on them.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every ``f_*`` is the *unbatched* transition the reference vmaps
(src/envs/simulators/interfaces/simulator.py:54); linearisation goes through
``torch.func.jacrev`` + ``vmap`` as at simulator.py:221-230, so the CUDA path's closed-form /
dual-number Jacobians are checked against an independent mechanism.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor

_ASSETS = Path(__file__).resolve().parent.parent / "safegbpo_b200" / "assets" / "assets.npz"


def _wrap(angle: Tensor) -> Tensor:
    # every simulator wraps its angle this way (pendulum.py:189, cartpole.py:179, quadrotor.py:220)
    return torch.atan2(torch.sin(angle), torch.cos(angle))


# ----------------------------------------------------------------------------------------------
# unbatched dynamics  f(s, a, w) -> s'
# ----------------------------------------------------------------------------------------------
def f_pendulum(s: Tensor, a: Tensor, w: Tensor) -> Tensor:
    """src/envs/simulators/pendulum.py:179-190 (DT .05, g 9.81, L 1, M 1, torque 30)."""
    dt, g, length, mass, tau = 0.05, 9.81, 1.0, 1.0, 30.0
    th, thd = s[0:1], s[1:2]
    acc = torch.sin(th) * 1.5 * g / length + (tau * a) * 3.0 / mass / length ** 2 + w[1:2]
    thd = thd + dt * acc
    th = _wrap(th + dt * thd)
    return torch.cat([th, thd])


def f_cartpole(s: Tensor, a: Tensor, w: Tensor) -> Tensor:
    """src/envs/simulators/cartpole.py:160-180.  Note (Q5): reads w[0], the noise set excites w[3]."""
    dt, g, half_len, m_cart, m_pole, f_mag = 0.02, 9.8, 0.5, 1.0, 0.1, 10.0
    x, th, xd, thd = s[0:1], s[1:2], s[2:3], s[3:4]
    total = m_pole + m_cart
    tmp = (f_mag * a + m_pole * half_len * thd ** 2 * torch.sin(th)) / total
    thdd = (1.0 / half_len) * ((g * torch.sin(th) + tmp * torch.cos(th))
                               / (4.0 / 3.0 - m_pole / total * torch.cos(th) ** 2) + w[0:1])
    xdd = tmp - m_pole * half_len / total * torch.cos(th) * thdd
    xd = xd + dt * xdd
    thd = thd + dt * thdd
    x = x + dt * xd
    th = _wrap(th + dt * thd)
    return torch.cat([x, th, xd, thd])


def f_quadrotor(s: Tensor, a: Tensor, w: Tensor) -> Tensor:
    """src/envs/simulators/quadrotor.py:204-221."""
    dt, g, k0, k1, k2 = 0.05, 9.81, 70.0, 17.0, 55.0
    t_mag, r_mag = 6.0 / 7.0, math.pi / 12
    x, z, roll, xd, zd, rolld = (s[i:i + 1] for i in range(6))
    thrust = a[0:1] * t_mag + g
    roll_cmd = a[1:2] * r_mag
    xdd = thrust * torch.sin(roll) + w[3:4]
    zdd = thrust * torch.cos(roll) - g + w[4:5]
    rolldd = roll_cmd * k2 - k0 * roll - k1 * rolld
    xd = xd + dt * xdd
    zd = zd + dt * zdd
    rolld = rolld + dt * rolldd
    x = x + dt * xd
    z = z + dt * zd
    roll = _wrap(roll + dt * rolld)
    return torch.cat([x, z, roll, xd, zd, rolld])


def f_seeker(s: Tensor, a: Tensor, w: Tensor) -> Tensor:
    """src/envs/simulators/seeker.py:185."""
    return s + 0.1 * a + w


def f_household(s: Tensor, a: Tensor, w: Tensor) -> Tensor:
    """src/envs/simulators/household.py:197-213 (DT 1; HEAT_PUMP_POWER_OFFSET unused, Q7)."""
    h_fh, h_out, tau, c_w, cop = 1.1, 0.26, 240.0, 1.1625, 3.01
    c0, c1, c2, c4 = (h_fh + h_out) / (h_out * tau), h_fh / (h_out * tau), h_fh / c_w, cop / c_w
    soc, t_in, t_ret = s[0:1], s[1:2], s[2:3]
    p_bat = a[0:1] * 2.0
    p_hp = a[1:2] * 2.5
    c3 = w[1:2] / tau
    soc = soc + p_bat
    t_in_new = t_in + (-c0 * t_in + c1 * t_ret + c3)
    t_ret_new = t_ret + (c2 * t_in - c2 * t_ret + c4 * p_hp)
    return torch.cat([soc, t_in_new, t_ret_new])


# ----------------------------------------------------------------------------------------------
# specification table
# ----------------------------------------------------------------------------------------------
@dataclass
class SimSpec:
    name: str
    n: int
    m: int
    f: Callable
    state_center: list
    state_half: list
    noise_center: list
    noise_half: list
    clamp: bool = True      # Household.execute_action does not clamp (household.py:168-177, Q7)


PI = math.pi
SIMS = {
    "pendulum": SimSpec("pendulum", 2, 1, f_pendulum, [0, 0], [PI, 8.0], [0, 0], [0.0, 0.1]),
    "cartpole": SimSpec("cartpole", 4, 1, f_cartpole, [0] * 4, [10.0, PI, 10.0, 10.0], [0] * 4, [0, 0, 0, 0.1]),
    "quadrotor": SimSpec("quadrotor", 6, 2, f_quadrotor, [0] * 6, [8.0, 8.0, PI / 12, 0.8, 1.0, PI / 2],
                         [0] * 6, [0, 0, 0, 0.1, 0.1, 0]),
    "seeker": SimSpec("seeker", 2, 2, f_seeker, [0, 0], [8.0, 8.0], [0, 0], [0.05, 0.05]),
    "household": SimSpec("household", 3, 2, f_household, [5.0, 21.0, 55.0], [5.0, 3.0, 45.0],
                         [0.0, 15.0, 0.0], [0.0, 25.0, 0.0], clamp=False),
}

ENV_TO_SIM = {
    "BalancePendulum": "pendulum", "BalanceCartPole": "cartpole", "SwingUpCartPole": "cartpole",
    "BalanceQuadrotor": "quadrotor", "NavigateQuadrotor": "quadrotor", "NavigateSeeker": "seeker",
    "ManageHousehold": "household",
}
RCI_ENVS = {"BalancePendulum": "pendulum", "BalanceCartPole": "cartpole", "BalanceQuadrotor": "quadrotor"}
HOUSEHOLD_START = 7800            # household.py:57
BUYING_DAY = [0.2] * 6 + [0.4] * 5 + [0.2] * 6 + [0.5] * 5 + [0.2] * 2   # household.py:63-66


def load_assets() -> dict:
    return {k: torch.from_numpy(v) for k, v in np.load(_ASSETS).items()}


# ----------------------------------------------------------------------------------------------
# the stateful oracle environment
# ----------------------------------------------------------------------------------------------
class OracleEnv:
    """Restates Simulator (simulator.py:13-282) + one task env.  float64, CPU.

    RNG: draws come from ``rng(kind, shape)`` (default: torch.rand on the global generator) in the
    reference's call order, so the same seed reproduces the reference's stream.
    """

    def __init__(self, env_name: str, num_envs: int, num_steps: int, rng: Optional[Callable] = None):
        self.env_name = env_name
        self.sim = SIMS[ENV_TO_SIM[env_name]]
        self.num_envs, self.num_steps = num_envs, num_steps
        n, m = self.sim.n, self.sim.m
        self.state_dim, self.action_dim = n, m
        f64 = torch.float64
        self.state_center = torch.tensor(self.sim.state_center, dtype=f64)
        self.state_half = torch.tensor(self.sim.state_half, dtype=f64)
        self.noise_center = torch.tensor(self.sim.noise_center, dtype=f64)
        self.noise_half = torch.tensor(self.sim.noise_half, dtype=f64)
        self.state_min, self.state_max = self.state_center - self.state_half, self.state_center + self.state_half
        self.rng = rng or (lambda kind, shape: torch.rand(shape, dtype=f64) if kind == "rand"
                           else torch.randn(shape, dtype=f64))
        self.state = torch.zeros(num_envs, n, dtype=f64)   # reference: torch.empty (simulator.py:50, Q4)
        self.steps = 0
        self.scheduled_reset = True                        # uniform over envs (truncation only)
        self.dynamics = torch.vmap(self.sim.f)
        self.goal = torch.zeros(num_envs, 2, dtype=f64)
        assets = load_assets()
        self.ctrl = self.rci = None
        if env_name in RCI_ENVS:                           # rci_env.py:20-35
            p = RCI_ENVS[env_name]
            self.ctrl = (assets[f"{p}_ctrl_center"], assets[f"{p}_ctrl_generators"])
            self.rci = (assets[f"{p}_rci_center"], assets[f"{p}_rci_generators"])
        self.obstacles = []                                 # NavigateSeeker: list of (center (N,2), radius (N))
        if self.sim.name == "household":                   # household.py:59-66, 95-98
            s0 = HOUSEHOLD_START
            self.load = assets["load_data"][s0:]
            self.pv = assets["pv_data"][s0:]
            self.t_out = assets["outside_temperature_data"][s0:]
            self.price = torch.tensor(BUYING_DAY * 366, dtype=f64)[s0:]

    # -- sets --------------------------------------------------------------------------------
    @property
    def action_constrained(self) -> bool:
        return self.env_name in RCI_ENVS or self.env_name == "NavigateSeeker"

    @property
    def state_constrained(self) -> bool:
        return self.env_name != "NavigateSeeker"

    def safe_action_set(self):
        """(center (m,), generator (m,p)) -- rci_env.py:38-45.  NavigateSeeker: per-environment
        (center (N,m), generator (N,m,2)) from program P5 (navigate_seeker.py:398-477, no_grad)."""
        if self.env_name == "NavigateSeeker":
            from .programs import p5_seeker_generator
            with torch.no_grad():
                gens = [p5_seeker_generator(self.state[i].detach(), [(c[i], r[i]) for c, r in self.obstacles],
                                            self.state_min, self.state_max) for i in range(self.num_envs)]
            return torch.zeros(self.num_envs, 2, dtype=torch.float64), torch.stack(gens)
        return self.ctrl

    def safe_state_set(self):
        """rci_env.py:48-56; swingup_cartpole.py:86-96; manage_household.py:72-79 (feasible box)."""
        if self.rci is not None:
            return self.rci
        return self.state_center, torch.diag(self.state_half)

    def sample_box(self, center, half):
        # Box.sample: center + sum(G * (2u-1)), u ~ rand(N,1,dim)  (box.py:73-74)
        u = self.rng("rand", (self.num_envs, 1, center.shape[0]))
        return center + (torch.diag(half) * (2 * u - 1)).sum(dim=2)

    def sample_zonotope(self, center, gen):
        # Zonotope.sample (zonotope.py:77-80)
        u = self.rng("rand", (self.num_envs, 1, gen.shape[1]))
        return center + (gen * (2 * u - 1)).sum(dim=2)

    # -- episode logic -----------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None):
        """simulator.py:57-79 then the task override (e.g. balance_cartpole.py:60-63)."""
        if seed is not None:
            torch.manual_seed(seed)
        self.state = self.sample_box(self.state_center, self.state_half)
        self.steps = 0
        self.scheduled_reset = False
        if self.rci is not None:
            self.state = self.sample_zonotope(*self.rci)
        if self.env_name == "SwingUpCartPole":              # swingup_cartpole.py:62-63
            self.state = self.rng("rand", (self.num_envs, 4)) * 0.1 - 0.05
            self.state[:, 1] = math.pi
        if self.sim.name in ("quadrotor", "seeker"):        # quadrotor.py:119, seeker.py:100
            self.goal = self.state[:, 0:2].clone()
        return self.observation()

    def noise_sample(self):
        w = self.sample_box(self.noise_center, self.noise_half)
        if self.sim.name == "household":                    # household.py:175-176
            w = w.clone()
            w[:, 1] = self.t_out[self.steps]
        return w

    def execute_action(self, action: Tensor, noise: Optional[Tensor] = None):
        """simulator.py:174-182; household.py:168-177."""
        w = self.noise_sample() if noise is None else noise
        self.last_noise = w
        prev = self.state
        self.state = self.dynamics(self.state, action, w)
        if self.obstacles:
            self.state = self._collision_operator(prev, self.state)
        if self.sim.clamp:
            self.state = torch.clamp(self.state, self.state_min, self.state_max)

    def _collision_operator(self, prev: Tensor, free: Tensor) -> Tensor:
        """navigate_seeker.py:233-296: elastic bounce off the LAST obstacle the free state entered."""
        N = self.num_envs
        collided = torch.stack([torch.linalg.vector_norm(free - c, dim=1) <= r for c, r in self.obstacles], dim=1)
        self.collided = collided
        mask = collided.any(dim=1)
        if not bool(mask.any()):
            return free
        which = torch.zeros(N, dtype=torch.long)
        for j in range(len(self.obstacles)):
            which = torch.where(collided[:, j], torch.full_like(which, j), which)
        centers = torch.stack([c for c, _ in self.obstacles])[which, torch.arange(N)]
        radii = torch.stack([r for _, r in self.obstacles])[which, torch.arange(N)].unsqueeze(1)
        safe_mask = mask.unsqueeze(1)
        to_start = prev - centers
        vel = free - prev
        a = (vel * vel).sum(1, keepdim=True)
        b = 2 * (to_start * vel).sum(1, keepdim=True)
        c = (to_start * to_start).sum(1, keepdim=True) - radii ** 2
        disc = torch.where(safe_mask, b ** 2 - 4 * a * c, torch.ones_like(a))
        a_safe = torch.where(safe_mask, a, torch.ones_like(a))
        t = (-b - torch.sqrt(disc)) / (2 * a_safe)
        inter = prev + t * vel
        normal = inter - centers
        nn = torch.where(safe_mask, torch.linalg.vector_norm(normal, dim=1, keepdim=True), torch.ones_like(a))
        normal = normal / nn
        vdn = (vel * normal).sum(1, keepdim=True)
        refl = vel - 2 * vdn * normal
        bounced = inter + refl * ((1 - t) * 0.1)
        return torch.where(safe_mask, bounced, free)

    def step(self, action: Tensor, noise: Optional[Tensor] = None):
        """simulator.py:82-108 (Q6: reset-all, reward on the post-reset state)."""
        self.execute_action(action, noise)
        self.steps += 1
        if self.scheduled_reset:
            self.reset()
        truncated = self.steps >= self.num_steps
        self.scheduled_reset = truncated
        term = torch.zeros(self.num_envs, dtype=torch.bool)
        trunc = torch.full((self.num_envs,), truncated, dtype=torch.bool)
        return self.observation(), self.reward(action), term, trunc

    def observation(self) -> Tensor:
        s = self.state
        if self.sim.name == "pendulum":      # pendulum.py:92-96
            return torch.cat([torch.sin(s[:, 0:1]), torch.cos(s[:, 0:1]), s[:, 1:2]], dim=1)
        if self.sim.name == "cartpole":      # cartpole.py:83-88
            return torch.cat([s[:, 0:1], torch.sin(s[:, 1:2]), torch.cos(s[:, 1:2]), s[:, 2:4]], dim=1)
        if self.sim.name in ("quadrotor", "seeker"):   # quadrotor.py:132, seeker.py:113, navigate_seeker.py:189-197
            extra = [torch.cat([c, r.unsqueeze(1)], dim=1) for c, r in self.obstacles]
            return torch.cat([s, self.goal] + extra, dim=1)
        t = self.steps                        # household.py:109-113
        series = torch.cat([self.load[t:t + 5], self.pv[t:t + 5], self.t_out[t:t + 5], self.price[t:t + 5]])
        return torch.cat([s, series.unsqueeze(0).expand(self.num_envs, -1)], dim=1)

    def reward(self, a: Tensor) -> Tensor:
        s = self.state
        name = self.env_name
        if name == "BalancePendulum":        # balance_pendulum.py:81-84
            return -1.0 * s[:, 0] ** 2 - 0.1 * s[:, 1] ** 2 - 0.001 * a[:, 0] ** 2
        if name in ("BalanceCartPole", "SwingUpCartPole"):   # balance_cartpole.py:77-82
            return (-0.05 * s[:, 0] ** 2 - 0.1 * s[:, 2] ** 2 - 1.0 * s[:, 1] ** 2
                    - 0.1 * s[:, 3] ** 2 - 0.0 * a[:, 0] ** 2)
        if name == "BalanceQuadrotor":       # balance_quadrotor.py:90-101
            thrust = a[:, 0] * (6.0 / 7.0) + 9.81
            roll_cmd = a[:, 1] * (math.pi / 12)
            dist = torch.sqrt((s[:, 0] - self.goal[:, 0]) ** 2 + (s[:, 1] - self.goal[:, 1]) ** 2)
            return (-2.5 * dist - 0.1 * s[:, 2] ** 2 - 0.1 * s[:, 3] ** 2 - 0.1 * s[:, 4] ** 2
                    - 0.1 * s[:, 5] ** 2 - 0.002 * thrust ** 2 - 0.001 * roll_cmd ** 2)
        if name == "ManageHousehold":        # manage_household.py:57-69 (raw a, not scaled powers)
            t = self.steps
            p_total = a[:, 0] + a[:, 1] + self.load[t] + self.pv[t]
            cost = torch.where(p_total >= 0, p_total * self.price[t], p_total * 0.08)
            return -cost - (s[:, 1] - 21.0) ** 2 * 100.0
        if name == "NavigateSeeker":         # navigate_seeker.py:212
            return -2.0 * torch.linalg.vector_norm(self.goal - s, dim=1)
        raise NotImplementedError(name)

    # -- linearisation (simulator.py:206-243) -----------------------------------------------------
    def linear_dynamics(self):
        n, m, N = self.state_dim, self.action_dim, self.num_envs
        a0 = torch.zeros(N, m, dtype=torch.float64)
        w0 = self.noise_center.expand(N, n)
        const = self.dynamics(self.state, a0, w0)
        jac = torch.vmap(torch.func.jacrev(self.sim.f, argnums=(0, 1, 2)))(self.state, a0, w0)
        return const, jac[0], jac[1], jac[2]

    def reachable_set(self):
        const, _, b_mat, _ = self.linear_dynamics()
        eye = torch.eye(self.action_dim, dtype=torch.float64).expand(self.num_envs, -1, -1)
        return const, torch.bmm(b_mat, eye)

    def cut_computation_graph(self):
        self.state = self.state.detach()
