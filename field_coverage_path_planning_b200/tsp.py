"""``TSPSolver`` — the 2-opt solver the reference imports from ``multi_field_planner_v37`` (multi_field_planner.py:176,
multi_vehicle_planner.py:131) but does not ship.  BUILD-DEFINED, parity unpinned: there is no
reference source, only the call ``TSPSolver.solve(distance_matrix) -> route`` (node indices, node 0 = depot).

Algorithm (restated in oracle/tsp.py, deterministic): nearest-neighbour tour from node 0, then best-improvement 2-opt on
the closed tour until no move shortens it by more than 1e-9 — on the device, one CTA per problem, batched
(``fcpp_tsp_two_opt``).  The distance matrix must be symmetric (the reference's are Euclidean)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from .batch import _dev


def two_opt_batch(matrices: Sequence[np.ndarray], max_iter: int = 0, device=None):
    """[(route list, closed-tour length, moves)] for every distance matrix, all problems in ONE launch."""
    dev = _dev(device)
    h = _lib.handle(dev.index)
    P = len(matrices)
    if P == 0:
        return []
    mats = [np.ascontiguousarray(m, dtype=np.float64) for m in matrices]
    for m in mats:
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("distance_matrix must be [n, n]")
    ns = [m.shape[0] for m in mats]
    mat_start = np.concatenate([[0], np.cumsum([n * n for n in ns])]).astype(np.int64)
    node_start = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    with torch.cuda.device(dev):
        D = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats]) if mat_start[-1] else np.zeros(1)).to(dev)
        ms = torch.from_numpy(mat_start).to(dev)
        nst = torch.from_numpy(node_start).to(dev)
        tours = torch.empty(max(int(node_start[-1]), 1), dtype=torch.int32, device=dev)
        lengths = torch.empty(P, dtype=torch.float64, device=dev)
        iters = torch.empty(P, dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        mi = int(max_iter) if max_iter > 0 else 100 * max(max(ns), 1)
        h.check(h.lib.fcpp_tsp_two_opt(h.h, P, ms.data_ptr(), D.data_ptr(), nst.data_ptr(), int(max(ns)), tours.data_ptr(),
                                       lengths.data_ptr(), iters.data_ptr(), mi, st))
        t, ln, it = tours.cpu().numpy(), lengths.cpu().numpy(), iters.cpu().numpy()
    return [(t[node_start[p]:node_start[p + 1]].tolist(), float(ln[p]), int(it[p])) for p in range(P)]


class TSPSolver:
    """``TSPSolver.solve(distance_matrix)`` as the reference calls it (a static call on the class)."""

    @staticmethod
    def solve(distance_matrix, device=None) -> List[int]:
        return two_opt_batch([np.asarray(distance_matrix, dtype=np.float64)], device=device)[0][0]


def tsp_solver_class():
    """The class behind ``from multi_field_planner_v37 import TSPSolver``: a module of that name supplied by the caller
    (what the reference expects to find) wins; otherwise this package's device 2-opt."""
    try:
        from multi_field_planner_v37 import TSPSolver as theirs
        return theirs
    except ModuleNotFoundError:
        return TSPSolver
