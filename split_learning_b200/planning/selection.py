"""Device selection threshold from a 2-component Gaussian mixture on log(speed).

Reference src/Selection.py:4-48: fit a 2-GMM on log performance, threshold = intersection
of the two weighted Gaussians that lies between the means (closest to the midpoint), else
the midpoint; return exp(threshold).  ``gmm_1d`` is a dependency-free EM (1-D, full
covariance == scalar variance) with deterministic restarts standing in for sklearn's
``GaussianMixture(n_components=2, n_init=9, random_state=0)``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def gmm_1d(x: np.ndarray, n_init: int = 9, iters: int = 200, seed: int = 0, reg: float = 1e-6
           ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (means[2], variances[2], weights[2]) of the best-likelihood fit."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    rng = np.random.RandomState(seed)
    best = (-np.inf, None)
    xs = np.sort(x)
    for trial in range(n_init):
        if trial == 0:      # split at the largest gap: a strong deterministic start
            gaps = np.diff(xs)
            cut = int(np.argmax(gaps)) + 1 if n > 1 else 1
            mu = np.array([xs[:cut].mean(), xs[cut:].mean() if cut < n else xs[-1]])
        else:
            mu = rng.choice(x, 2, replace=n < 2)
        var = np.full(2, max(x.var(), reg))
        w = np.full(2, 0.5)
        ll_old = -np.inf
        for _ in range(iters):
            logp = -0.5 * ((x[:, None] - mu[None]) ** 2 / var[None] + np.log(2 * np.pi * var[None])) + np.log(w[None])
            m = logp.max(1, keepdims=True)
            lse = m + np.log(np.exp(logp - m).sum(1, keepdims=True))
            r = np.exp(logp - lse)
            ll = float(lse.sum())
            nk = r.sum(0) + 1e-300
            mu = (r * x[:, None]).sum(0) / nk
            var = (r * (x[:, None] - mu[None]) ** 2).sum(0) / nk + reg
            w = nk / n
            if abs(ll - ll_old) < 1e-10:
                break
            ll_old = ll
        if ll > best[0]:
            best = (ll, (mu.copy(), var.copy(), w.copy()))
    return best[1]


def _threshold_from_params(mu, var, w) -> float:
    order = np.argsort(mu)
    mu, var, w = mu[order], var[order], w[order]
    a = var[0] - var[1]
    b = 2 * (var[1] * mu[0] - var[0] * mu[1])
    c = var[0] * mu[1] ** 2 - var[1] * mu[0] ** 2 + 2 * var[0] * var[1] * np.log((var[1] * w[0]) / (var[0] * w[1]))
    mid = float(np.mean(mu))
    if np.isclose(a, 0):
        if np.isclose(b, 0):
            return mid
        root = -c / b
        return float(root) if mu[0] < root < mu[1] else mid
    roots = np.roots([a, b, c])
    real = roots[np.isreal(roots)].real
    cand = real[(real > mu[0]) & (real < mu[1])]
    if cand.size:
        return float(cand[np.argmin(np.abs(cand - mid))])
    return mid


def auto_threshold(performance, n_init: int = 9, backend: str = "auto") -> float:
    perf = np.asarray(performance, dtype=float)
    if perf.size <= 1:
        return 0.0
    x = np.log(perf)
    params = None
    if backend in ("auto", "sklearn"):
        try:
            from sklearn.mixture import GaussianMixture
            gm = GaussianMixture(n_components=2, n_init=n_init, covariance_type="full", random_state=0).fit(x.reshape(-1, 1))
            params = (gm.means_.flatten(), gm.covariances_.reshape(-1), gm.weights_)
        except Exception:
            if backend == "sklearn":
                raise
    if params is None:
        params = gmm_1d(x, n_init=n_init)
    return float(np.exp(_threshold_from_params(*map(np.asarray, params))))
