"""Seeded synthetic inputs of the shapes BASELINE.json names.

All generators return Fortran-ordered float64 arrays (the reference is only ever exercised
with ``matrix_layout=102``, column-major: ``test.ipynb`` cell 2).
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps


def gaussian(m: int, n: int, seed: int = 0) -> np.ndarray:
    """Dense i.i.d. N(0,1) matrix (configs C1, C3, C4)."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.standard_normal((m, n)))


def graded(n: int, r: int | None = None, seed: int = 0, m: int | None = None) -> np.ndarray:
    """Graded-spectrum rank-deficient matrix, the generator of the reference's only test
    (``test.ipynb`` cell 3): ``X = U diag(sv) V^T`` with ``sv_i = 2^(1-i) + 1e-17`` (i = 1..)
    and ``sv[:r] += 0.01 (r - i)``; U, V = Q-factors of seeded Gaussians.  Since the r-th
    term adds 0 the true numerical rank is r-1 (config C2: n = 4096, r = 2048)."""
    m = n if m is None else m
    k = min(m, n)
    r = k // 2 if r is None else r
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, m)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    i = np.arange(1, k + 1, dtype=np.float64)
    sv = 2.0 ** (1.0 - i) + 1e-17
    sv[:r] += 0.01 * (r - i[:r])
    X = (U[:, :k] * sv) @ V[:, :k].T
    return np.asfortranarray(X)


def graded_tall(m: int, n: int, r: int | None = None, seed: int = 0) -> np.ndarray:
    """Tall graded-spectrum rank-deficient matrix: the ``graded`` recipe with a THIN left factor (Q of an m x n
    Gaussian, so m can be 10^5 without an m x m QR).  Exercises the DM early stop inside the blocked tall panel."""
    r = n // 2 if r is None else r
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    i = np.arange(1, n + 1, dtype=np.float64)
    sv = 2.0 ** (1.0 - i) + 1e-17
    sv[:r] += 0.01 * (r - i[:r])
    return np.asfortranarray((U * sv) @ V.T)


def kahan(n: int, theta: float = 1.25, perturb: float = 0.0, seed: int | None = None) -> np.ndarray:
    """Kahan matrix ``K = diag(s^0..s^(n-1)) (I - c triu(ones,1))``, c = cos(theta), s = sin(theta)
    (config C5; not in the reference — defined in SURVEY.md §8d).  ``perturb`` scales the
    customary diagonal perturbation ``perturb * eps * (n, n-1, .., 1)`` that keeps column
    pivoting from being triggered by rounding; with ``seed`` the perturbation is additionally
    multiplied by seeded U(0.5,1.5) factors so that matrices of a batch differ."""
    c, s = np.cos(theta), np.sin(theta)
    K = np.eye(n) - c * np.triu(np.ones((n, n)), 1)
    K = (s ** np.arange(n))[:, None] * K
    if perturb:
        d = perturb * EPS * np.arange(n, 0, -1, dtype=np.float64)
        if seed is not None:
            d = d * np.random.default_rng(seed).uniform(0.5, 1.5, n)
        K[np.diag_indices(n)] += d
    return np.asfortranarray(K)


def planted(m: int, n: int, seed: int = 0, eps: float = 1e-9) -> np.ndarray:
    """Gaussian columns among which every third one is (almost) the sum of its two left neighbours: pairwise cosines stay
    below 0.9 (0.707), so Deviation Maximisation may select such triples into one block — inside the block the third column
    then loses all but `eps` of its norm, which is what the panel's early stop (and the grouped panels' accuracy guard) exist for."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    for c in range(2, n, 3):
        A[:, c] = A[:, c - 1] + A[:, c - 2] + eps * rng.standard_normal(m)
    return np.asfortranarray(A)


def flops(m: int, n: int, r: int) -> float:
    """Algorithmic FLOPs of a rank-r Householder QR of an m x n matrix:
    ``4mnr - 2(m+n)r^2 + (4/3)r^3`` (= ``2mn^2 - (2/3)n^3`` for r = n <= m)."""
    m, n, r = float(m), float(n), float(r)
    return 4.0 * m * n * r - 2.0 * (m + n) * r * r + (4.0 / 3.0) * r ** 3
