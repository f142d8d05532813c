"""Device containment verdicts (reference: src/sets/zonotope.py:143-239, programs P9; src/sets/box.py:90-102).

The reference answers `outer.contains(other)` with one host LP per batch element.  For an outer zonotope shared by all
environments (the constant safe sets of the tasks) the LP's feasible region in the offset between the centres is a fixed polytope;
its facets are enumerated ONCE here (init-time host work, like consts.build_consts) and the per-environment verdicts are then one
kernel launch (csrc/sgbpo_contains.cu, C ABI sgbpo_contains).

    v = DeviceVerdict(outer_center, outer_generator)                   # point program: facets of the zonotope
    v = DeviceVerdict(outer_center, outer_generator, inner_generator)   # zonotope program with fixed inner generators
    value, inside = v(points)                                           # (N,), (N,) bool
    value, inside = v(centers, inner_generators)                        # per-environment inner generators (exact containment)
"""
from __future__ import annotations

import ctypes as C
from itertools import combinations
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from .. import _lib


def zonotope_facets(G: np.ndarray, tol: float = 1e-11) -> np.ndarray:
    """Facets  h . x <= 1  of the zonotope Z(0, G) (full-dimensional): every facet normal is orthogonal to n - 1 generators."""
    n = G.shape[0]
    G = G[:, np.linalg.norm(G, axis=0) > tol]
    if np.linalg.matrix_rank(G) < n:
        raise ValueError("zonotope_facets: the zonotope is not full-dimensional")
    if n == 1:
        return np.array([[1.0], [-1.0]]) / np.abs(G).sum()
    rows, seen = [], set()
    for idx in combinations(range(G.shape[1]), n - 1):
        sub = G[:, idx]
        u, s, vt = np.linalg.svd(sub.T, full_matrices=True)
        if s.size < n - 1 or s[-1] < tol * max(1.0, s[0]):
            continue                                   # the n - 1 generators do not span a hyperplane
        nrm = vt[-1]
        nrm = nrm / np.linalg.norm(nrm)
        k = int(np.argmax(np.abs(nrm)))
        nrm = nrm * np.sign(nrm[k])
        key = tuple(np.round(nrm, 9))
        if key in seen:
            continue
        seen.add(key)
        h = np.abs(nrm @ G).sum()
        rows += [nrm / h, -nrm / h]
    return np.array(rows)


class DeviceVerdict:
    def __init__(self, outer_center: Tensor, outer_generator: Tensor, inner_generator: Optional[Tensor] = None, tol: float = 0.0):
        """outer_center (n,) or (N, n) [per-environment centres, shared generators]; outer_generator (n, p); inner_generator
        (n, q) = generators every inner zonotope shares (the zonotope program), or None (the point program)."""
        G = outer_generator.detach().double().cpu().numpy()
        if inner_generator is None:
            facets = zonotope_facets(G)
        else:
            from ..consts import merge_parallel_generators, polytope_facets
            facets = polytope_facets(merge_parallel_generators(G), inner_generator.detach().double().cpu().numpy())
        self.center = outer_center.detach().contiguous()
        self.facets = torch.as_tensor(facets, dtype=self.center.dtype, device=self.center.device).contiguous()
        self.tol = float(tol)

    def __call__(self, points: Tensor, inner_generators: Optional[Tensor] = None) -> tuple[Tensor, Tensor]:
        if points.device.type != "cuda":
            raise _lib.SgbpoError("DeviceVerdict needs CUDA tensors (there is no CPU fallback)")
        pts = points.detach().to(self.center.dtype).contiguous()
        N, n = pts.shape
        gen = None if inner_generators is None else inner_generators.detach().to(pts.dtype).contiguous()
        value = torch.empty(N, dtype=pts.dtype, device=pts.device)
        verdict = torch.empty(N, dtype=torch.int32, device=pts.device)
        a = _lib.ContainsArgs()
        a.N, a.n, a.K, a.q = N, n, self.facets.shape[0], 0 if gen is None else gen.shape[2]
        a.center_per_env = int(self.center.dim() == 2)
        a.facets, a.outer_center, a.points = self.facets.data_ptr(), self.center.data_ptr(), pts.data_ptr()
        a.inner_gen = None if gen is None else gen.data_ptr()
        a.value, a.verdict, a.tol = value.data_ptr(), verdict.data_ptr(), self.tol
        _lib.check(_lib.lib().sgbpo_contains(_lib.dtype_code(pts.dtype), C.byref(a), _lib.stream_ptr()), "sgbpo_contains")
        return value, verdict != 0
