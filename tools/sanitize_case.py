"""Smallest run that touches every kernel of libfcpp.so once — for compute-sanitizer
(memcheck / racecheck), one tool per gpurun call (B200_PROFILING.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import field_coverage_path_planning_b200 as fc  # noqa: E402

RECT = [(0, 0), (120, 0), (120, 60), (0, 60)]
PARA = [(1000, 2000), (1150, 2030), (1170, 2110), (1020, 2080)]
OBST = [[(40, 20), (50, 20), (50, 30), (40, 30)], [(10, 8), (30, 8), (20, 14)]]
cand = fc.make_candidates(2, headings=[0.0, 0.4], radii=[5.0, 13.0], start_corners=[0, 3])
res = fc.plan_batch([RECT, PARA], fc.VehicleParams(), cand, obstacles=[OBST, []], outputs="paths", grid_h=0.05)
print("statuses", np.unique(res.summary["status"]), "cov", int(res.summary["cov_cells"].sum()))
res = fc.plan_batch([RECT], fc.VehicleParams(), fc.make_candidates(1, radii=[28.0]), grid_h=0.1)   # multi-tile corner windows
print("big R status", res.summary["status"], res.summary["corner_after"])
p = fc.TwoLayerPathPlannerV37(fc.VehicleParams(), field_length=120, field_width=60, start_point=(100, 50), end_point=(5, 5))
r = p.plan_complete_coverage()
allp = np.vstack([r["main_work"]["path"], r["headland"]["path"]])
alls = np.concatenate([r["main_work"]["speeds"], r["headland"]["speeds"]])
print(p.verify_curvature_constraints(allp, alls)["accel_violations"], p.verify_all_corners_coverage(r["headland"])["avg_improvement"])
print(p.verify_corner_coverage_grid_based((8, 8), 0, allp[-60:-40], allp[-40:-20])["cells_after"])
rng = np.random.default_rng(0)
D = rng.uniform(1, 10, (33, 33))
pop = np.array([rng.permutation(33) for _ in range(200)], dtype=np.int32)
print(float(fc.tour_lengths(D, pop).sum()))
