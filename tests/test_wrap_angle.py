"""The simulators wrap angles with atan2(sin x, cos x) (reference: pendulum.py / cartpole.py / quadrotor.py dynamics).  The kernels
evaluate the same map as x - 2 pi rint(x / 2 pi) with a two-term 2 pi (csrc/sgbpo_dual.cuh::wrap_angle): exact for |x| <= pi, within
rounding of the atan2 form elsewhere, derivative 1 -- two range reductions and an atan2 less per dynamics evaluation."""
import numpy as np

import hostmath_util as hm


def _ref(x):
    return np.arctan2(np.sin(x), np.cos(x))


def test_identity_inside_the_principal_interval():
    for dtype in (np.float64, np.float32):
        x = np.linspace(-np.pi, np.pi, 20001).astype(dtype)[1:-1]
        y, dy = hm.wrap_angle(x, dtype)
        assert np.array_equal(y, x) and np.all(dy == 1)


def test_matches_atan2_of_sin_cos():
    rs = np.random.default_rng(0)
    x = np.concatenate([rs.uniform(-40, 40, 200000), np.pi + rs.uniform(-1e-6, 1e-6, 1000), -np.pi + rs.uniform(-1e-6, 1e-6, 1000),
                        3 * np.pi + rs.uniform(-1e-3, 1e-3, 1000)])
    y, dy = hm.wrap_angle(x, np.float64)
    ref = _ref(x)
    # (at the cut +-pi both forms are the same angle: compare on the circle)
    err = np.abs(np.angle(np.exp(1j * (y - ref))))
    assert err.max() < 1e-14 and np.all(np.abs(y) <= np.pi + 1e-15) and np.all(dy == 1)
    away = np.abs(np.abs(ref) - np.pi) > 1e-9
    assert np.abs(y[away] - ref[away]).max() < 1e-14
    x32 = x.astype(np.float32)
    y32, _ = hm.wrap_angle(x32, np.float32)
    ref32 = _ref(x32.astype(np.float64))
    err32 = np.abs(np.angle(np.exp(1j * (y32.astype(np.float64) - ref32))))
    assert err32.max() < 4e-6                      # float32: one rounding of a value of magnitude <= 40
