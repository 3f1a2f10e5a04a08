"""Generate the golden fixtures in this directory by EXECUTING THE UNMODIFIED REFERENCE
(/root/reference/multi_layer_planner_v3.py) through the Shapely/matplotlib stand-ins of
oracle/shapely_stub.py (decisions D1/D2; GEOS itself is not available, so GEOS-dependent
values — ring order, buffer discretisation, exact coverage_rate — are "parity unpinned").

Run from the repo root, in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/ref_*.npz + ga_tours.npz

Scenarios = the reference's own test scripts and __main__ demo:
  test/test_v37_complete.py:142-161 (three rectangles with start points)   -> BASELINE config 1
  test/test_v351_start_end_points.py:13-46 (start/end point, 14 km/h headland)
  test/test_multi-layer_planner_v3.py:30-46 (500x200 default + corner-grid verification)
  multi_layer_planner_v3.py:1618-1640 (obstacle scenarios)                  -> BASELINE config 2
plus the tilted-rectangle and parallelogram probes of SURVEY.md App. C.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shapely_stub  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tilted(L, Wd, deg, ox=100.0, oy=0.0):
    a = np.deg2rad(deg)
    return [(float(x * np.cos(a) - y * np.sin(a) + ox), float(x * np.sin(a) + y * np.cos(a) + oy))
            for x, y in [(0, 0), (L, 0), (L, Wd), (0, Wd)]]


OBST2 = [[(200, 80), (250, 80), (250, 120), (200, 120)], [(350, 140), (380, 140), (380, 170), (350, 170)]]

# name -> (VehicleParams kwargs, planner kwargs, run corner-grid verification?)
SCENARIOS = {
    "rect500": (dict(), dict(field_length=500, field_width=200), True),
    "rect500_se": (dict(max_headland_speed_kmh=14.0),
                   dict(field_length=500, field_width=200, start_point=(10, 10), end_point=(490, 190)), False),
    "v37_small": (dict(), dict(field_length=100, field_width=80, start_point=(90, 70)), True),
    "v37_medium": (dict(), dict(field_length=500, field_width=200, start_point=(50, 180)), False),
    "v37_large": (dict(), dict(field_length=3500, field_width=320, start_point=(3400, 300)), False),
    "start_rb": (dict(), dict(field_length=500, field_width=200, start_point=(490, 10)), False),
    "obst2": (dict(), dict(field_length=500, field_width=200, obstacles=OBST2), False),
    "obst_xj": (dict(), dict(field_length=3500, field_width=320,
                             obstacles=[[(1500, 100), (1600, 100), (1600, 200), (1500, 200)]]), False),
    "para": (dict(), dict(field_vertices=[(0, 0), (500, 0), (580, 200), (80, 200)]), False),
    "tilt20": (dict(), dict(field_vertices=tilted(500, 200, 20), start_point=(300, 200)), False),
    "tilt20_obst": (dict(), dict(field_vertices=tilted(500, 200, 20),
                                 obstacles=[tilted(40, 30, 20, ox=100 + 150, oy=120)]), False),
    "r72_knife": (dict(min_turn_radius=7.2), dict(field_length=500, field_width=200), False),
    "r96_knife": (dict(min_turn_radius=9.6), dict(field_length=300, field_width=150), True),
}


def run_one(m, vkw, pkw, corners):
    base = dict(working_width=3.2, min_turn_radius=8.0, max_work_speed_kmh=9.0, max_headland_speed_kmh=15.0)
    base.update(vkw)
    veh = m.VehicleParams(**base)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        p = m.TwoLayerPathPlannerV37(veh, **pkw)
        r = p.plan_complete_coverage()
        allp = np.vstack([r['main_work']['path'], r['headland']['path']])
        alls = np.concatenate([r['main_work']['speeds'], r['headland']['speeds']])
        cc = p.verify_curvature_constraints(allp, alls)
        cov = p.verify_all_corners_coverage(r['headland']) if corners else None
    out = {
        "main_path": r['main_work']['path'], "main_speeds": r['main_work']['speeds'],
        "head_path": r['headland']['path'], "head_speeds": r['headland']['speeds'],
        "main_stats": np.array([r['main_work']['stats'][k] for k in ('path_length_km', 'time_hours', 'avg_speed_kmh')]),
        "head_stats": np.array([r['headland']['stats'][k] for k in ('path_length_km', 'time_hours', 'avg_speed_kmh')]),
        "coverage_rate_sampled": np.float64(r['headland']['stats']['coverage_rate']),  # numeric, GEOS excluded
        "curv": np.array([cc['max_curvature'], cc['max_lateral_accel'], cc['accel_violations'],
                          cc['accel_violation_rate'], cc['max_jump'], float(cc['pass'])]),
        "approach": r['approach_path'] if r['approach_path'] is not None else np.zeros((0, 2)),
        "departure": r['departure_path'] if r['departure_path'] is not None else np.zeros((0, 2)),
        "meta": np.array(json.dumps(dict(vehicle=base, planner={k: v for k, v in pkw.items()},
                                         field_shape=p.field_shape, corner_angles=list(map(float, p.corner_angles)),
                                         field_length=float(p.field_length), field_width=float(p.field_width),
                                         pattern=p.main_work_pattern), ensure_ascii=False)),
    }
    if cov is not None:
        # float D2 counts (the integer oracle may differ by a cell where a lattice point sits on
        # the buffer boundary — SURVEY.md App. C)
        out["corner_cells_float"] = np.array([[int(round(c['coverage_before'] * c['grid'].size / 100)),
                                               int(c['grid'].sum())] for c in cov['corners']])
        out["corner_avg"] = np.array([cov['avg_coverage_before'], cov['avg_coverage_after'], cov['avg_improvement']])
    return out


def make_ga(path="/root/reference/genetic_algorithm_solver.py"):
    """Tour lengths of genetic_algorithm_solver.py:_calculate_distance on seeded inputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_ga", path)
    ga = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ga)
    solver = ga.GeneticAlgorithmSolver()
    rng = np.random.default_rng(42)
    n = 41
    pos = np.vstack([[100.0, 100.0], rng.uniform(0, 5000, size=(n - 1, 2))])
    D = np.zeros((n, n))
    for i in range(n):           # multi_field_planner.py:263-288 layout: node 0 = depot
        for j in range(n):
            if i != j:
                D[i, j] = np.linalg.norm(pos[i] - pos[j])
    prng = np.random.default_rng(7)
    pop = np.array([prng.permutation(n) for _ in range(64)], dtype=np.int32)
    d = np.array([solver._calculate_distance(list(map(int, r)), D) for r in pop])
    f = np.array([solver._calculate_fitness(list(map(int, r)), D) for r in pop])
    np.savez_compressed(os.path.join(HERE, "ga_tours.npz"), pos=pos, D=D, pop=pop, dist=d, fitness=f)


def main():
    m = shapely_stub.load_reference()
    if m is None:
        raise SystemExit("/root/reference not present: fixtures can only be regenerated in the build container")
    only = set(sys.argv[1:])
    for name, (vkw, pkw, corners) in SCENARIOS.items():
        if only and name not in only:
            continue
        out = run_one(m, vkw, pkw, corners)
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
        print(name, out["main_path"].shape, out["head_path"].shape, flush=True)
    if not only or "ga" in only:
        make_ga()
        print("ga_tours")


if __name__ == "__main__":
    main()
