"""CPU restatement of the build-defined 2-opt solver (TEST INFRASTRUCTURE ONLY).

The reference calls ``TSPSolver.solve(distance_matrix)`` (multi_field_planner.py:176-177, multi_vehicle_planner.py:131-132)
from a module it does not ship, so there is NO reference algorithm to pin: "parity unpinned".  This file states the
algorithm the product implements on the device, in numpy, with the same floating-point expression and the same
tie-breaks, so that device and CPU tours must be IDENTICAL:
  1. nearest-neighbour tour from node 0 (first minimum among the unvisited nodes);
  2. best-improvement 2-opt on the closed tour: for all tour-edge pairs (i, i+1), (j, j+1) with j >= i + 2 (and not the
     cyclically identical pair i = 0, j = n - 1): delta = (D[a,c] + D[b,d]) - (D[a,b] + D[c,d]); apply the lowest delta
     if it is < -1e-9 (ties: lowest i, then lowest j) by reversing tour[i+1 .. j]; repeat."""
from __future__ import annotations

import numpy as np


def nearest_neighbour(D: np.ndarray):
    n = len(D)
    tour, used = [0], np.zeros(n, dtype=bool)
    used[0] = True
    for _ in range(1, n):
        d = np.where(used, np.inf, D[tour[-1]])
        k = int(np.argmin(d))            # first minimum
        tour.append(k)
        used[k] = True
    return tour


def two_opt(D: np.ndarray, max_iter: int = 0):
    """-> (tour, closed length summed left to right, moves)."""
    D = np.asarray(D, dtype=np.float64)
    n = len(D)
    if n == 0:
        return [], 0.0, 0
    tour = np.array(nearest_neighbour(D), dtype=np.int64)
    max_iter = max_iter or 100 * max(n, 1)
    I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ok = (J >= I + 2) & ~((I == 0) & (J == n - 1))
    it = 0
    while it < max_iter and ok.any():
        a, b = tour[I], tour[np.minimum(I + 1, n - 1)]
        c, d = tour[J], tour[(J + 1) % n]
        delta = np.where(ok, (D[a, c] + D[b, d]) - (D[a, b] + D[c, d]), np.inf)
        q = int(np.argmin(delta))        # row-major first minimum = lowest (i, j)
        if not (delta.flat[q] < -1e-9):
            break
        i, j = divmod(q, n)
        tour[i + 1:j + 1] = tour[i + 1:j + 1][::-1].copy()
        it += 1
    length = 0.0
    for k in range(n):
        length += D[tour[k], tour[(k + 1) % n]]
    return tour.tolist(), float(length), it
