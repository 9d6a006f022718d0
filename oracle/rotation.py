"""Rotation oracles (TEST INFRASTRUCTURE, see oracle/__init__.py).

`haar_rotation_qr`        restates scipy 1.18.1 `stats.special_ortho_group.rvs(N)`
                          (scipy/stats/_multivariate.py:4384-4390 + 4535-4543), the call the
                          reference makes at optex.py:149.  scipy is un-vendored and un-pinned by
                          the reference; pinned here bit-exactly against the installed scipy by
                          tests/test_rotation_oracle.py.
`haar_rotation_householder`  restates the reference's own `impl="torch"` branch
                          (optex.py:151-164, Stewart's Householder construction) with the Gaussian
                          draws INJECTED, in float64.  This is the algorithm the on-device
                          generator (`csrc/rotation.cu`) implements in fp32.
"""
from __future__ import annotations

import numpy as np


def haar_rotation_qr(n: int, random_state) -> np.ndarray:
    """float64 [n, n] SO(n) sample; `random_state` is a numpy RandomState/Generator or int seed."""
    if isinstance(random_state, (int, np.integer)):
        random_state = np.random.RandomState(int(random_state))
    z = random_state.normal(size=(n, n))
    q, r = np.linalg.qr(z)
    d = r.diagonal()
    q *= (d / abs(d))[np.newaxis, :]          # column signs: make R's diagonal positive
    det = np.linalg.det(q)
    if n:
        q[0, :] /= det                        # flip first row if det == -1
    return q


def householder_vectors(gauss: np.ndarray):
    """Normalised reflection vectors v_n (aligned to columns n..N-1) and signs D, from
    Gaussians `gauss[n, n:]` (rows 0..N-2 used).  optex.py:154-161."""
    g = np.asarray(gauss, dtype=np.float64)
    n_dim = g.shape[1]
    v = np.zeros((n_dim - 1, n_dim))
    d = np.empty(n_dim)
    for n in range(n_dim - 1):
        x = g[n, n:].copy()
        norm2 = x @ x
        x0 = x[0]
        d[n] = 1.0 if x0 >= 0 else -1.0       # sign(sign(x0) + 0.5)
        x[0] += d[n] * np.sqrt(norm2)
        x /= np.sqrt((norm2 - x0 * x0 + x[0] * x[0]) / 2.0)
        v[n, n:] = x
    d[-1] = (-1.0) ** (n_dim - 1) * np.prod(d[:-1])
    return v, d


def haar_rotation_householder(gauss: np.ndarray) -> np.ndarray:
    """optex.py:151-164 with injected normals: H <- H (I - v v^T) for n = 0..N-2, rows scaled by D."""
    v, d = householder_vectors(gauss)
    n_dim = v.shape[1]
    h = np.eye(n_dim)
    for n in range(n_dim - 1):
        x = v[n, n:]
        h[:, n:] -= np.outer(h[:, n:] @ x, x)
    return d[:, None] * h
