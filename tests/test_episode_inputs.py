"""The one-launch episode-input generator (csrc/sgbpo_inputs.cu): Philox4x32-10 keyed by (seed, element, purpose, episode).

CPU: the numpy restatement of the generator used below is pinned on the published Random123 known-answer vectors of
philox4x32-10.  GPU: the kernel's noise samples and reset states equal the restatement element by element (so the stream is a
documented function of (seed, episode, t, environment), independent of the launch shape), the policy draws are standard
normal, the ray-mask directions are unit columns, and two launches with the same counter are bit-identical."""
import ctypes as C

import numpy as np
import pytest
import torch

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """ctr: (..., 4) uint32 counters; key: two ints.  Returns (..., 4) uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def test_philox_known_answers():
    """Random123 kat_vectors: philox4x32 10 rounds."""
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        got = philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want


def uniforms(seed: int, elements: np.ndarray, purpose: int, episode: int, count: int, dtype) -> np.ndarray:
    """The kernel's Uniforms<T> stream: `count` uniforms per element (sgbpo_inputs.cu)."""
    per_call = 4 if dtype == np.float32 else 2
    calls = -(-count // per_call)
    out = []
    for call in range(calls):
        ctr = np.stack([elements & 0xFFFFFFFF, elements >> 32, np.full_like(elements, call | (purpose << 24)),
                        np.full_like(elements, episode & 0xFFFFFFFF)], axis=-1).astype(np.uint32)
        bits = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        if dtype == np.float32:
            out.append(((bits >> 8).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0))
        else:
            b = bits.astype(np.uint64)
            hi, lo = b[..., 0::2], b[..., 1::2]
            out.append((((hi << np.uint64(21)) ^ (lo >> np.uint64(11))).astype(np.float64) + 0.5) / 9007199254740992.0)
    return np.concatenate(out, axis=-1)[..., :count]


def _launch(dtype, N, H, n, m, n_resets, gens, want_dirs, seed, episode, lo, span, center, G):
    from safegbpo_b200 import _lib
    dev = torch.device("cuda:0")
    z = lambda *s: torch.full(s, float("nan"), dtype=dtype, device=dev)
    eps, noise, dirs, rs = z(H, N, m), z(H, N, n), z(H, N, m, 2 * m), z(max(n_resets, 1), N, n)
    ctr = torch.tensor([episode], dtype=torch.int64, device=dev)
    a = _lib.EpisodeInputs()
    a.N, a.H, a.n, a.m, a.n_resets, a.n_reset_gens, a.want_dirs, a.seed = N, H, n, m, n_resets, gens, int(want_dirs), seed
    a.episode, a.eps, a.noise, a.dirs, a.reset_states = ctr.data_ptr(), eps.data_ptr(), noise.data_ptr(), dirs.data_ptr(), rs.data_ptr()
    for i in range(n):
        a.noise_lo[i], a.noise_span[i], a.reset_center[i] = lo[i], span[i], center[i]
        for j in range(gens):
            a.reset_gen[i][j] = G[i][j]
    _lib.check(_lib.lib().sgbpo_episode_inputs(_lib.dtype_code(dtype), C.byref(a), _lib.stream_ptr()), "sgbpo_episode_inputs")
    torch.cuda.synchronize()
    return eps.cpu().numpy(), noise.cpu().numpy(), dirs.cpu().numpy(), rs.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_kernel_equals_restatement(dtype):
    npdt = np.float32 if dtype == torch.float32 else np.float64
    N, H, n, m, R, P = 777, 9, 4, 2, 3, 6
    rs_ = np.random.default_rng(5)
    lo, span, center = rs_.normal(size=n), rs_.uniform(0.1, 1.0, size=n), rs_.normal(size=n)
    G = rs_.normal(size=(n, P))
    seed, episode = 0x1234567890ABCDEF, 41
    eps, noise, dirs, rs = _launch(dtype, N, H, n, m, R, P, True, seed, episode, lo, span, center, G)
    elements = np.arange(H * N, dtype=np.uint64)
    tol = dict(rtol=3e-6, atol=3e-6) if dtype == torch.float32 else dict(rtol=1e-14, atol=1e-14)
    u = uniforms(seed, elements, 2, episode, n, npdt).astype(np.float64)
    np.testing.assert_allclose(noise.reshape(H * N, n), lo + span * u, **tol)
    u = uniforms(seed, elements[:R * N], 4, episode, P, npdt).astype(np.float64)
    np.testing.assert_allclose(rs.reshape(R * N, n), center + (2 * u - 1) @ G.T, **tol)
    u = uniforms(seed, elements, 1, episode, 2, npdt).astype(np.float64)
    r, ang = np.sqrt(-2 * np.log(u[:, 0])), 2 * np.pi * u[:, 1]
    etol = dict(rtol=2e-5, atol=2e-5) if dtype == torch.float32 else dict(rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(eps.reshape(H * N, m), np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1), **etol)
    u = uniforms(seed, elements, 3, episode, 2 * m * m, npdt).astype(np.float64).reshape(H * N, m, 2 * m)
    d = 2 * u - 1
    np.testing.assert_allclose(dirs.reshape(H * N, m, 2 * m), d / np.linalg.norm(d, axis=1, keepdims=True), **etol)


@pytest.mark.gpu
def test_distributions_and_counters():
    N, H, n, m = 4096, 32, 4, 1
    lo, span, center, G = [-0.1, 0, -1, 2], [0.2, 0, 2, 0.5], [0, 3.0, 0, 0], np.diag([0.05, 0.0, 0.05, 0.05])
    args = (torch.float32, N, H, n, m, 2, 4, True, 7, 3, lo, span, center, G)
    eps, noise, dirs, rs = _launch(*args)
    assert np.isfinite(eps).all() and abs(eps.mean()) < 0.01 and abs(eps.std() - 1) < 0.01
    assert abs((eps ** 4).mean() - 3.0) < 0.1                                   # kurtosis of a normal
    assert (noise >= np.array(lo, dtype=np.float32)).all() and (noise <= np.array(lo, dtype=np.float32) + np.array(span, dtype=np.float32)).all()
    assert abs(noise[..., 2].mean()) < 0.01 and abs(noise[..., 2].var() - 4 / 12) < 0.01
    assert (noise[..., 1] == 0).all()
    assert np.allclose(np.abs(dirs), 1.0)                                       # m = 1: the unit "columns" are +-1
    assert (rs[..., 1] == 3.0).all() and (np.abs(rs[..., [0, 2, 3]]) <= 0.05 + 1e-7).all()
    again = _launch(*args)
    for a, b in zip((eps, noise, dirs, rs), again):
        assert np.array_equal(a, b)
    other = _launch(*(args[:9] + (4,) + args[10:]))                              # next episode: a different stream
    assert not np.array_equal(other[0], eps) and abs(np.corrcoef(other[0].ravel(), eps.ravel())[0, 1]) < 0.01
    # steps of one environment are uncorrelated, environments of one step are uncorrelated
    assert abs(np.corrcoef(eps[0, :, 0], eps[1, :, 0])[0, 1]) < 0.05
    assert abs(np.corrcoef(eps[:, 0, 0], eps[:, 1, 0])[0, 1]) < 0.5


@pytest.mark.gpu
def test_argument_errors():
    from safegbpo_b200 import _lib
    a = _lib.EpisodeInputs()
    assert _lib.lib().sgbpo_episode_inputs(0, C.byref(a), None) != 0
    assert b"null" in _lib.lib().sgbpo_last_error()
