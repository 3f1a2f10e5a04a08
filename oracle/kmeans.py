"""CPU restatement of the KMeans the reference's multi-vehicle split calls (TEST INFRASTRUCTURE ONLY).

"mvp" = /root/reference/multi_vehicle_planner.py:186-209 — ``KMeans(n_clusters=V, random_state=42).fit_predict``.
scikit-learn is a third-party dependency of the reference (absent from /root/reference, unpinned there; 1.9.0 in the
container that made the fixtures).  Restated from its published algorithm: data centred on its mean, k-means++ seeding
with 2 + log(k) greedy local trials drawn from ``numpy.random.RandomState``, Lloyd iterations (E step = first minimum of
||c||^2 - 2 x.c, M step = member means, empty cluster <- the point farthest from its centre), stop on unchanged labels
or centre shift <= 1e-4 * mean per-axis variance, one more E step without strict convergence.
Pinned by tests/golden/multi_vehicle.npz: labels the UNMODIFIED reference method produced with the real sklearn
(tests/golden/make_multi_vehicle_golden.py).
"""
from __future__ import annotations

import numpy as np


def _sq(A, X, x_sq):
    d = (A * A).sum(axis=1)[:, None] - 2.0 * (A @ X.T) + x_sq[None, :]
    return np.maximum(d, 0.0)


def seeds(Xc: np.ndarray, k: int, rs: np.random.RandomState) -> np.ndarray:
    n = len(Xc)
    x_sq = (Xc * Xc).sum(axis=1)
    trials = 2 + int(np.log(k))
    idx = [int(rs.choice(n, p=np.full(n, 1.0) / float(n)))]
    closest = _sq(Xc[idx[0]][None, :], Xc, x_sq)[0]
    pot = closest.sum()
    for _ in range(1, k):
        r = rs.uniform(size=trials) * pot
        cand = np.minimum(np.searchsorted(np.cumsum(closest), r), n - 1)
        d = np.minimum(closest[None, :], _sq(Xc[cand], Xc, x_sq))
        pots = d.sum(axis=1)
        b = int(np.argmin(pots))
        pot, closest = pots[b], d[b]
        idx.append(int(cand[b]))
    return np.asarray(idx)


def lloyd(Xc: np.ndarray, centers: np.ndarray, max_iter: int = 300, tol: float = 1e-4):
    """-> (labels, centres, iterations) on mean-centred data."""
    n, k = len(Xc), len(centers)
    c = centers.copy()
    thr = tol * np.mean(np.var(Xc, axis=0))
    labels = np.full(n, -1)
    strict = False
    it = 0

    def e_step(c):
        return np.argmin((c * c).sum(axis=1)[None, :] - 2.0 * (Xc @ c.T), axis=1)

    for it in range(1, max_iter + 1):
        new = e_step(c)
        sums = np.zeros((k, 2))
        np.add.at(sums, new, Xc)
        cnt = np.bincount(new, minlength=k).astype(float)
        empty = np.nonzero(cnt == 0)[0]
        if len(empty):
            d = ((Xc - c[new]) ** 2).sum(axis=1)
            far = np.argsort(-d, kind="stable")[:len(empty)]
            for j, i in zip(empty, far):
                sums[new[i]] -= Xc[i]
                cnt[new[i]] -= 1
                sums[j], cnt[j] = Xc[i], 1.0
        cn = np.where(cnt[:, None] > 0, sums / np.maximum(cnt, 1)[:, None], sums)
        shift = (np.sqrt(((cn - c) ** 2).sum(axis=1)) ** 2).sum()
        same = np.array_equal(new, labels)
        labels, c = new, cn
        if same:
            strict = True
            break
        if shift <= thr:
            break
    if not strict:
        labels = e_step(c)
    return labels.astype(np.int32), c, it


def kmeans_labels(points, k: int, random_state: int = 42):
    X = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    mean = X.mean(axis=0)
    Xc = X - mean
    labels, c, it = lloyd(Xc, Xc[seeds(Xc, k, np.random.RandomState(random_state))])
    return labels, c + mean, it


def cluster_fields(fields_data: dict, num_vehicles: int):
    """mvp:186-209: fields grouped by KMeans label, in the order of the dict."""
    ids = list(fields_data.keys())
    labels, _, _ = kmeans_labels([fields_data[f]['centroid'] for f in ids], num_vehicles)
    clusters = [[] for _ in range(num_vehicles)]
    for f, l in zip(ids, labels):
        clusters[int(l)].append(f)
    return clusters
