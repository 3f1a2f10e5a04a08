"""Extended parity sweep (run by hand on a GPU box; not collected by pytest): many more candidates than
test_gpu_fullsize.py samples, drawn from inside the full-size BASELINE batches, compared field by field with the CPU
oracle.  Prints one JSON line with the number of candidates compared and the worst differences observed.

    python tests/parity_sweep.py [--c2 768] [--c3 512] [--c5 48] > profiles/r2_parity_sweep.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import field_coverage_path_planning_b200 as fc  # noqa: E402
from benchmarks import workloads as wl  # noqa: E402
from test_gpu_fullsize import oracle_many, sample_indices  # noqa: E402

INT_KEYS = ("n_passes", "n_loops", "n_main", "n_head", "n_accel_viol", "n_boundary_viol", "n_obstacle_viol", "corner_g",
            "cov_total", "cov_cells")
FP_KEYS = ("len_main", "len_head", "time_main", "time_head", "time_main_pre", "time_head_pre", "max_curvature",
           "max_lateral_accel", "max_jump")


def compare(res, w, idx, paths):
    recs = oracle_many([(w.oracle_args(i), True, paths) for i in idx])
    out = {"candidates": len(idx), "int_mismatches": 0, "status_mismatches": 0, "worst_rel": 0.0, "worst_point_m": 0.0,
           "worst_speed_kmh": 0.0}
    for i, o in zip(idx, recs):
        s = res.summary[i]
        if int(s["status"]) != o["status"]:
            out["status_mismatches"] += 1
            continue
        if o["status"]:
            continue
        for k in INT_KEYS:
            out["int_mismatches"] += int(s[k]) != int(o[k])
        out["int_mismatches"] += list(map(int, s["corner_before"])) != o["corner_before"]
        out["int_mismatches"] += list(map(int, s["corner_after"])) != o["corner_after"]
        for k in FP_KEYS:
            d = abs(float(s[k]) - o[k]) / max(abs(o[k]), 1e-12) if o[k] != 0 else abs(float(s[k]))
            out["worst_rel"] = max(out["worst_rel"], d)
        if paths:
            p, v, _ = res.path(i)
            out["worst_point_m"] = max(out["worst_point_m"], float(np.abs(p - o["path"]).max()))
            out["worst_speed_kmh"] = max(out["worst_speed_kmh"], float(np.abs(v - o["speeds"]).max()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c2", type=int, default=768)
    ap.add_argument("--c3", type=int, default=512)
    ap.add_argument("--c5", type=int, default=48)
    a = ap.parse_args()
    veh = fc.VehicleParams()
    report = {}
    w = wl.c2(1)
    res = fc.plan_batch(w.fields, veh, w.axes, obstacles=w.obstacles, outputs="paths", grid_h=w.grid_h)
    report["c2 (4096 candidates, paths)"] = compare(res, w, sample_indices(w.n_cand, a.c2, 11), True)
    w = wl.c3(1, 256)
    res = fc.plan_batch(w.fields, veh, w.axes, outputs="summary", grid_h=w.grid_h)
    report["c3 (256 fields x 180 headings)"] = compare(res, w, sample_indices(w.n_cand, a.c3, 12), False)
    w = wl.c5(1)
    res = fc.plan_batch(w.fields, veh, w.axes, outputs="summary", grid_h=w.grid_h)
    report["c5 (8192 candidates, h = 0.05 m)"] = compare(res, w, sample_indices(w.n_cand, a.c5, 13), False)
    report["criteria"] = "integers exact; FP64 sums relative; points / speeds absolute (north_star: 1e-4 m, 3.6e-4 km/h)"
    print(json.dumps(report))
    bad = sum(r["int_mismatches"] + r["status_mismatches"] for r in report.values() if isinstance(r, dict))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
