"""Oracle for ONE candidate of the batch path (TEST INFRASTRUCTURE ONLY): everything a
``fcpp_summary`` record holds, computed with the CPU restatement (ref_planner.py, raster.py).

This is also the unit of work of bench.py's CPU baseline ("port"): one call = one full candidate
plan (generation + speed planning + kinematic/geofence validation + coverage), i.e. what a user
of the reference does with plan_complete_coverage() + verify_curvature_constraints() +
verify_all_corners_coverage() (mlp3:1663-1668, test/test_multi-layer_planner_v3.py:41-46).
"""
from __future__ import annotations

from dataclasses import replace
from typing import Dict, Optional, Sequence

import numpy as np

from . import raster, ref_planner as rp

STATUS_INSET_EMPTY = 1
STATUS_LOOP_SKIPPED = 2


def evaluate_candidate(field_vertices, vehicle: rp.VehicleParams, R: Optional[float] = None,
                       heading: Optional[float] = None, start_corner: Optional[int] = None,
                       obstacles: Sequence = (), grid_h: float = 0.1, coverage: bool = True,
                       keep_paths: bool = False, turn_model: str = "arc", clothoid_share: float = 0.5) -> Dict:
    veh = replace(vehicle, min_turn_radius=float(R)) if R is not None else vehicle
    fs = rp.setup_field(veh, field_vertices=[tuple(map(float, v)) for v in field_vertices],
                        obstacles=[list(map(tuple, o)) for o in obstacles], turn_model=turn_model,
                        clothoid_share=clothoid_share)
    out: Dict = {"status": 0}
    try:
        res = rp.plan_complete_coverage(fs, heading=heading, start_corner=start_corner)
    except rp.PlanError as e:
        out["status"] = STATUS_INSET_EMPTY if "无法定义主作业区域" in str(e) else STATUS_LOOP_SKIPPED
        return out
    mp, hp = res["main_work"]["path"], res["headland"]["path"]
    ms, hs = res["main_work"]["speeds"], res["headland"]["speeds"]
    allp = np.vstack([mp, hp])
    alls = np.concatenate([ms, hs])
    pre = res["_info"]["speeds_pre"]
    cc = rp.verify_curvature_constraints(allp, alls, veh)
    out.update(
        n_passes=res["_info"]["P"], n_loops=res["_info"]["K"], n_main=len(mp), n_head=len(hp),
        len_main=rp.path_length(mp), len_head=rp.path_length(hp),
        time_main=rp.work_time(mp, ms), time_head=rp.work_time(hp, hs),
        time_main_pre=rp.work_time(mp, pre[:len(mp)]), time_head_pre=rp.work_time(hp, pre[len(mp):]),
        n_accel_viol=cc["accel_violations"], max_curvature=cc["max_curvature"],
        max_lateral_accel=cc["max_lateral_accel"], max_jump=cc["max_jump"],
        n_boundary_viol=rp.boundary_violations(allp, fs.field_vertices),
        n_obstacle_viol=rp.obstacle_violations(allp, fs.obstacles, veh.working_width),
    )
    if coverage:
        cells, g = raster.corner_coverage(fs)
        total, cov = raster.band_coverage(fs, hp, grid_h)
        out.update(corner_g=g, corner_before=[c[0] for c in cells], corner_after=[c[1] for c in cells],
                   cov_total=total, cov_cells=cov)
    if keep_paths:
        out.update(path=allp, speeds=alls)
    return out


def candidate_cost(o: Dict, kind: str = "length") -> float:
    if o["status"]:
        return float("inf")
    return o["len_main"] + o["len_head"] if kind == "length" else o["time_main"] + o["time_head"]
