"""Golden fixtures for the multi-field glue, produced by EXECUTING THE UNMODIFIED
REFERENCE /root/reference/multi_field_planner.py through oracle/shapely_stub.py.

    python tests/golden/make_multi_field_golden.py        # writes tests/golden/multi_field.npz

The reference module cannot be imported as shipped: it asks multi_layer_planner_v3
for ``TwoLayerPathPlannerV36``, a name that file does not define.  The loader below supplies the
alias V36 -> V37 in the imported module object (no source is modified); shapely / matplotlib come
from the stubs (decisions D1/D2).  Stored: the distance matrix of mfp:263-288, the best connection of
mfp:290-320 for EVERY ordered node pair, centroid / area / entry points of mfp:105-151 and the
area / W work-distance estimate of mfp:213-216 for 12 seeded quadrilateral fields + depot.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_mfp():
    from oracle import shapely_stub
    shapely_stub.install()
    mlp3 = shapely_stub.load_reference()
    mlp3.TwoLayerPathPlannerV36 = mlp3.TwoLayerPathPlannerV37      # the missing alias (F3)
    sys.modules["multi_layer_planner_v3"] = mlp3
    for name in ("genetic_algorithm_solver", "multi_vehicle_planner", "multi_field_planner"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["multi_field_planner"], mlp3


def fields(seed=5, F=12):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(F):
        L, Wd = rng.uniform(150, 600), rng.uniform(80, 300)
        sx = rng.uniform(-0.3, 0.3) * Wd
        phi = rng.uniform(0, np.pi) if k % 3 else 0.0
        ox, oy = rng.uniform(0, 5000, size=2)
        q = np.array([(0, 0), (L, 0), (L + sx, Wd), (sx, Wd)], dtype=np.float64)
        c, s = np.cos(phi), np.sin(phi)
        q = q @ np.array([[c, s], [-s, c]]) + (ox, oy)
        out.append({"id": f"F{k:02d}", "vertices": [tuple(map(float, v)) for v in q]})
    return out


def main():
    mfp, mlp3 = load_reference_mfp()
    defs = fields()
    depot = (100.0, 100.0)
    with contextlib.redirect_stdout(io.StringIO()):
        p = mfp.MultiFieldPlannerV38(defs, depot, mlp3.VehicleParams(), num_vehicles=1, optimization_method="genetic")
        D, node_ids = p._calculate_distance_matrix()
        n = len(node_ids)
        Cm = np.zeros((n, n))
        fpt = np.zeros((n, n, 2))
        tpt = np.zeros((n, n, 2))
        for a, ia in enumerate(node_ids):
            for b, ib in enumerate(node_ids):
                c = p._find_best_connection(ia, ib)
                Cm[a, b] = c.distance
                fpt[a, b] = c.from_point
                tpt[a, b] = c.to_point
        route = p.optimize_sequence()
    ids = node_ids[1:]
    np.savez_compressed(
        os.path.join(HERE, "multi_field.npz"),
        verts=np.array([d["vertices"] for d in defs]), depot=np.array(depot), D=D, C=Cm, from_pt=fpt, to_pt=tpt,
        centroid=np.array([p.fields[f].centroid for f in ids]), area=np.array([p.fields[f].area for f in ids]),
        entry_dir=np.array([[e[1] for e in p.fields[f].entry_points] for f in ids]),
        work_estimate=np.array([p.fields[f].area / p.vehicle_params.working_width for f in ids]),
        route_total_work=np.float64(route.total_work_distance),
        route_n_connections=np.int64(len(route.connections)))
    print("multi_field.npz:", D.shape, "work", route.total_work_distance, "transfer", route.total_transfer_distance)


if __name__ == "__main__":
    main()
