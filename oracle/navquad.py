"""Oracle restatement of NavigateQuadrotorEnv's per-step safe state set and step pieces (float64, CPU).
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/envs/navigate_quadrotor.py statement by statement:
  safe_state_set :436-460, safe_center :462-484, feasible_center :486-517, collision_free_center :519-540,
  compute_support_point (P6) :542-582, compute_position_generator (P7) :585-654, compute_velocity_generator (P8) :657-779,
  collision_check / collision_operator :258-328, reward :218-236.
The three programs are solved GENERICALLY here (scipy HiGHS for the LP, SLSQP + active-set polish for the two geo-mean
programs), not with the kernels' vertex / edge-midpoint enumeration: the formulation is shared, the solver is independent.
Parity unpinned at the solver boundary like P1-P5 (cvxpy / Clarabel are not installable); where an optimum is not unique
(an LP objective parallel to an edge) the tests compare the objective value, not the point.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog, minimize

DT = 0.05
GRAVITY, THRUST_MAG = 9.81, 6.0 / 7.0
BOX_MAX = np.array([8.0, 8.0, np.pi / 12, 0.8, 1.0, np.pi / 2])
BOX_MIN = -BOX_MAX
MAX_ACC = np.array([(THRUST_MAG + GRAVITY) * np.sin(np.pi / 16), THRUST_MAG])          # :83-86
ROLL_GEN = np.zeros((6, 2))
ROLL_GEN[2, 0], ROLL_GEN[5, 1] = np.pi / 12, np.pi / 2                                     # :97-99


class Infeasible(Exception):
    pass


def collision_check(pos: np.ndarray, obstacles) -> np.ndarray:
    return np.array([np.linalg.norm(pos - c) <= r for c, r in obstacles], dtype=bool)


def support_point(center: np.ndarray, gen: np.ndarray, direction: np.ndarray):
    """P6 (:542-582): beta* = argmax (direction^T G) beta  s.t. the support point stays in the state box, |beta|_inf <= 1."""
    obj = direction @ gen
    A = np.vstack([gen, -gen])
    b = np.concatenate([BOX_MAX - center, center - BOX_MIN])
    res = linprog(-obj, A_ub=A, b_ub=b, bounds=[(-1.0, 1.0)] * gen.shape[1], method="highs")
    if res.status != 0:
        raise Infeasible("support point program")
    return center + gen @ res.x, float(obj @ res.x)


def max_geo_mean(A: np.ndarray, b: np.ndarray):
    """argmax geo_mean(len), len >= 0, A len <= b (A >= 0 elementwise); None when infeasible / zero optimum."""
    if np.any(b < 0):
        return None
    scale = np.maximum(A.sum(1), 1e-12)
    x0 = np.full(2, 0.25 * float(np.min(b / scale)))
    if not np.all(x0 > 0):
        return None
    res = minimize(lambda x: -np.sum(np.log(x)), x0, jac=lambda x: -1.0 / x,
                   constraints=[{"type": "ineq", "fun": lambda x: b - A @ x, "jac": lambda x: -A}],
                   bounds=[(1e-12, None)] * 2, method="SLSQP", options={"maxiter": 500, "ftol": 1e-16})
    x = res.x
    act = np.where(b - A @ x <= 1e-7 * (1 + np.abs(b)))[0]
    if len(act) >= 2:
        sol = np.linalg.lstsq(A[act], b[act], rcond=None)[0]
        if np.all(sol > 0) and np.all(A @ sol <= b + 1e-9 * (1 + np.abs(b))):
            x = sol
    elif len(act) == 1 and np.all(A[act[0]] > 0):
        x = b[act[0]] / (2 * A[act[0]])
    return x


def feasible_center(c_f: np.ndarray, gen: np.ndarray):
    too_low, too_high = BOX_MIN > c_f, BOX_MAX < c_f
    inside = not (too_low.any() or too_high.any())                 # state_set.contains(center)  (before the coupling below)
    too_low, too_high = too_low.copy(), too_high.copy()
    too_low[2] |= too_low[3]; too_high[2] |= too_high[3]
    too_low[5] |= too_low[3]; too_high[5] |= too_high[3]
    direction = np.ones(6)
    direction[too_low] = -BOX_MIN[too_low]
    direction[too_high] = -BOX_MAX[too_high]
    if inside:
        return c_f.copy()
    return support_point(c_f, gen, direction)[0]


def collision_free_center(center: np.ndarray, c_f: np.ndarray, gen: np.ndarray, obstacles):
    hits = collision_check(center[:2], obstacles)
    if not hits.any():
        return center
    which = int(np.nonzero(hits)[0][-1])                               # obstacle_mask: the last colliding obstacle
    direction = np.concatenate([center[:2] - obstacles[which][0], np.zeros(4)])
    return support_point(c_f, gen, direction)[0]


def position_generator(center: np.ndarray, obstacles):
    """P7 (:585-654)."""
    c2 = center[:2]
    U = np.eye(2)
    min_distance = np.linalg.norm(np.abs(c2) - BOX_MAX[:2])
    for oc, r in obstacles:
        direction = oc - c2
        distance = np.linalg.norm(direction)
        to_obs = direction / distance
        distance -= r
        if distance < min_distance:
            U = np.stack([to_obs, np.array([to_obs[1], -to_obs[0]])], axis=1)
            min_distance = distance
    A, b = [], []
    for d in range(2):
        A.append(np.abs(U[d])); b.append(c2[d] - BOX_MIN[d])
        A.append(np.abs(U[d])); b.append(BOX_MAX[d] - c2[d])
    for oc, r in obstacles:
        direction = oc - c2
        distance = np.linalg.norm(direction)
        A.append(np.abs((direction / distance) @ U)); b.append(distance - r)
    length = max_geo_mean(np.array(A), np.array(b))
    if length is None:
        length = np.full(2, 1e-2)                                      # :650-651
    gen = np.zeros((6, 2))
    gen[:2] = U * length
    return gen


def velocity_generator(center: np.ndarray, c_f: np.ndarray, reach_gen: np.ndarray, obstacles):
    """P8 (:657-779), as shipped (incl. the bare-direction parameter of the obstacle rows)."""
    half = np.abs(reach_gen[:2]).sum(axis=1)                           # reachable_set.box() half widths of x, z
    lower_distance = np.abs(BOX_MIN[:2] - (c_f[:2] - half))
    upper_distance = np.abs(BOX_MAX[:2] - (c_f[:2] + half))
    lower_distance[0] += 5 * DT * center[3]
    upper_distance[0] -= 5 * DT * center[3]
    lower_distance[lower_distance < 0] = 1e-2
    upper_distance[upper_distance < 0] = 1e-2
    lower_wall = -np.sqrt(lower_distance * 2 * MAX_ACC) - center[3:5]
    lower_wall[lower_wall > 0] = 0.0
    upper_wall = np.sqrt(upper_distance * 2 * MAX_ACC) - center[3:5]
    upper_wall[upper_wall < 0] = 0.0
    U = np.eye(2)
    min_distance = np.linalg.norm(np.abs(center[3:5]) - BOX_MAX[3:5])
    min_distance = min(min_distance, np.abs(lower_wall).min(), np.abs(upper_wall).min())
    for oc, r in obstacles:
        direction = oc - center[:2]
        distance = np.linalg.norm(direction)
        direction = direction / distance
        speed = center[3:5] @ direction
        distance -= 5 * DT * speed
        if distance < 0:
            distance = 1e-2
        vel_bound = max(np.sqrt(distance * abs(np.sum(2 * direction * MAX_ACC))) - speed, 0.0)
        if vel_bound < min_distance:
            U = np.stack([direction, np.array([direction[1], -direction[0]])], axis=1)
            min_distance = vel_bound
    A, b = [], []
    for d in range(2):
        A.append(np.abs(U[d])); b.append(-lower_wall[d])
        A.append(np.abs(U[d])); b.append(upper_wall[d])
    for oc, r in obstacles:
        direction = oc - center[:2]
        distance = np.linalg.norm(direction)
        direction = direction / distance
        distance -= r
        speed = center[3:5] @ direction
        distance -= 5 * DT * speed
        if distance < 0:
            distance = 1e-2
        vel_bound = max(np.sqrt(distance * abs(np.sum(2 * direction * MAX_ACC))) - speed, 0.0)
        A.append(np.abs(direction)); b.append(vel_bound)
    length = max_geo_mean(np.array(A), np.array(b))
    if length is None:
        length = np.full(2, 1e-2)
    gen = np.zeros((6, 2))
    gen[3:5] = U * length
    return gen


def safe_state_set(c_f: np.ndarray, reach_gen: np.ndarray, obstacles):
    """(:436-460) for ONE environment: (center (6,), generator (6, 6), safe_position, safe_velocity)."""
    center = feasible_center(c_f, reach_gen)
    center = collision_free_center(center, c_f, reach_gen, obstacles)
    safe_position = not ((BOX_MIN[:2] > c_f[:2]).any() or (BOX_MAX[:2] < c_f[:2]).any() or collision_check(center[:2], obstacles).any())
    safe_velocity = not ((BOX_MIN[3:5] > c_f[3:5]).any() or (BOX_MAX[3:5] < c_f[3:5]).any())
    pos = position_generator(center, obstacles) if safe_position else np.full((6, 2), 1e-2)
    vel = velocity_generator(center, c_f, reach_gen, obstacles) if safe_velocity else np.full((6, 2), 1e-2)
    return center, np.concatenate([pos, vel, ROLL_GEN], axis=1), safe_position, safe_velocity


def collision_operator(prev: np.ndarray, free: np.ndarray, obstacles):
    """(:258-328) for ONE environment; returns (state, collided)."""
    hits = collision_check(free[:2], obstacles)
    if not hits.any():
        return free.copy(), False
    oc, r = obstacles[int(np.nonzero(hits)[0][-1])]
    pos, free_vel = prev[:2], free[3:5]
    to_start = pos - oc
    direction = free_vel * DT
    a, b, c = direction @ direction, 2 * to_start @ direction, to_start @ to_start - r ** 2
    t = (-b - np.sqrt(b * b - 4 * a * c)) / (2 * a)
    inter = pos + t * direction
    normal = (inter - oc) / np.linalg.norm(inter - oc)
    refl = free_vel - 2 * (free_vel @ normal) * normal
    out = np.array([*(inter + refl * (1 - t) * DT), prev[2], *refl, prev[5]])
    return out, True


def reward(state: np.ndarray, goal: np.ndarray, d0: float, reached: bool, collided: bool):
    """(:218-236); returns (reward, reached_after)."""
    dist = np.linalg.norm(goal - state[:2])
    just = (not reached) and dist < 0.1
    return 100.0 * just - dist / d0 - 10.0 * collided, reached or just
