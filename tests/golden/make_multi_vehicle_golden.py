"""Golden fixtures for the multi-vehicle split, produced by EXECUTING THE UNMODIFIED
REFERENCE /root/reference/multi_vehicle_planner.py with the REAL scikit-learn of this container.

    python tests/golden/make_multi_vehicle_golden.py        # writes tests/golden/multi_vehicle.npz

matplotlib (imported at module level for `visualize`, which is not on the path) comes from oracle/shapely_stub.py's
stand-ins; sklearn.cluster.KMeans is the real one (scikit-learn version stored in the fixture).  Stored, per
scenario (seeded centroid clouds: uniform, blobs, collinear, duplicates; 2..8 vehicles):
  * the labels of mvp:186-209 `_cluster_fields` (as cluster lists) — the pin of oracle/kmeans.py and of the
    device Lloyd kernel,
  * for the GA scenarios the whole `plan(..., use_genetic=True)` of mvp:65-184: per-vehicle field ids, work
    distance, and the transfer distances of 12 runs of the reference's own GA (it draws from the unseeded global
    `random`: a distribution, not a value),
  * the distance matrix of mvp:229-259 of the first vehicle.
`plan(..., use_genetic=False)` raises ModuleNotFoundError in the reference (mvp:131 imports a module that is not in
its tree); the fixture records that too.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_mvp():
    from oracle import shapely_stub
    shapely_stub.install()            # matplotlib stand-in (mvp:17-18); sklearn stays the real package
    for name in ("genetic_algorithm_solver", "multi_vehicle_planner"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["multi_vehicle_planner"]


def cloud(kind: str, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(0, 5000, size=(n, 2))
    if kind == "blobs":
        c = rng.uniform(0, 8000, size=(5, 2))
        return c[rng.integers(0, 5, n)] + rng.normal(0, 300, size=(n, 2))
    if kind == "line":
        t = rng.uniform(0, 10000, n)
        return np.stack([t, 0.3 * t + rng.normal(0, 5, n)], axis=1)
    if kind == "dups":
        base = rng.uniform(0, 3000, size=(max(n // 3, 2), 2))
        return base[rng.integers(0, len(base), n)]
    raise ValueError(kind)


SCENARIOS = [("uniform", 12, 2, 1), ("uniform", 60, 3, 2), ("uniform", 200, 4, 3), ("uniform", 1000, 8, 4),
             ("blobs", 90, 5, 5), ("blobs", 300, 3, 6), ("blobs", 64, 6, 7), ("line", 80, 4, 8), ("dups", 45, 3, 9),
             ("uniform", 7, 7, 10), ("blobs", 2500, 8, 11), ("uniform", 150, 2, 12)]
GA_SCENARIOS = [("uniform", 120, 3, 21), ("blobs", 160, 4, 22)]


class Veh:
    working_width = 3.2


def main():
    import sklearn
    mvp = load_reference_mvp()
    out = {}
    meta = {"sklearn": sklearn.__version__, "scenarios": [], "ga_scenarios": []}
    for kind, n, v, seed in SCENARIOS:
        pts = cloud(kind, n, seed)
        fd = {f"F{i:04d}": {"centroid": (float(p[0]), float(p[1])), "area": 1.0} for i, p in enumerate(pts)}
        with contextlib.redirect_stdout(io.StringIO()):
            clusters = mvp.MultiVehiclePlanner(v)._cluster_fields(fd, (0.0, 0.0))
        labels = np.full(n, -1, dtype=np.int32)
        for j, c in enumerate(clusters):
            for f in c:
                labels[int(f[1:])] = j
        key = f"{kind}_{n}_{v}"
        out[key + "_pts"], out[key + "_labels"] = pts, labels
        meta["scenarios"].append(key)
    for kind, n, v, seed in GA_SCENARIOS:
        pts = cloud(kind, n, seed)
        rng = np.random.default_rng(seed + 100)
        area = rng.uniform(2e4, 2e5, n)
        fd = {f"F{i:04d}": {"centroid": (float(p[0]), float(p[1])), "area": float(a)} for i, (p, a) in enumerate(zip(pts, area))}
        depot = (100.0, 100.0)
        runs = []
        for rep in range(12):          # the reference GA is unseeded: 12 runs give its distribution per vehicle
            random.seed(1000 * seed + rep)
            with contextlib.redirect_stdout(io.StringIO()):
                planner = mvp.MultiVehiclePlanner(v)
                route = planner.plan(fd, depot, Veh(), use_genetic=True)
                D0 = planner._build_distance_matrix(route.vehicle_routes[0].field_ids, fd, depot)
            runs.append([vr.total_transfer_distance for vr in route.vehicle_routes])
        key = f"ga_{kind}_{n}_{v}"
        out[key + "_pts"], out[key + "_area"] = pts, area
        out[key + "_D0"] = D0
        out[key + "_labels"] = np.array([next(vr.vehicle_id for vr in route.vehicle_routes if f"F{i:04d}" in vr.field_ids)
                                         for i in range(n)], dtype=np.int32)
        out[key + "_transfer"] = np.array(runs)          # [12 runs, vehicles]; the other values are of the last run
        out[key + "_work"] = np.array([vr.total_work_distance for vr in route.vehicle_routes])
        out[key + "_time"] = np.array([vr.work_time for vr in route.vehicle_routes])
        out[key + "_totals"] = np.array([route.total_transfer_distance, route.total_work_distance, route.total_distance,
                                         route.max_work_time, route.load_balance_ratio])
        meta["ga_scenarios"].append(key)
    # the default path of the reference: the 2-opt module it imports does not exist
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mvp.MultiVehiclePlanner(2).plan({f"F{i}": {"centroid": (float(i), 0.0), "area": 1.0} for i in range(6)},
                                            (0.0, 0.0), Veh(), use_genetic=False)
        meta["default_path_error"] = None
    except ModuleNotFoundError as e:
        meta["default_path_error"] = str(e)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "multi_vehicle.npz"), **out)
    print("wrote multi_vehicle.npz:", meta["sklearn"], len(meta["scenarios"]), "cluster scenarios,",
          len(meta["ga_scenarios"]), "GA scenarios; default path:", meta["default_path_error"])


if __name__ == "__main__":
    main()
