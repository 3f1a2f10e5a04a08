"""Multi-field glue: distance matrix, best-connection matrix and the
MultiFieldPlannerV38 scheduler against the unmodified reference (tests/golden/multi_field.npz, made by
tests/golden/make_multi_field_golden.py) and the oracle restatement."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "multi_field.npz")


def _nodes(z):
    return [[z["depot"]]] + [list(v) for v in z["verts"]]


# ------------------------------------------------------------------------------------- CPU
def test_oracle_reproduces_reference_multi_field():
    from oracle import multi_field as om
    z = np.load(GOLD)
    assert np.array_equal(om.distance_matrix(z["depot"], z["centroid"]), z["D"])
    Cm, fi, ti = om.connection_matrix(z["verts"], z["depot"])
    assert np.array_equal(Cm, z["C"])
    nodes = _nodes(z)
    n = len(nodes)
    for a in range(n):
        for b in range(n):
            assert np.array_equal(nodes[a][fi[a, b]], z["from_pt"][a, b])
            assert np.array_equal(nodes[b][ti[a, b]], z["to_pt"][a, b])
    assert np.array_equal(np.array([om.entry_directions(v) for v in z["verts"]]), z["entry_dir"])


def test_host_field_records_match_reference():
    """centroid / area of the drop-in's field polygon == the reference's (Shapely stub, D1)."""
    from field_coverage_path_planning_b200 import _geometry as G
    z = np.load(GOLD)
    for k, v in enumerate(z["verts"]):
        q = G.QuadPolygon([tuple(x) for x in v])
        assert abs(q.area - z["area"][k]) <= 1e-9 * z["area"][k]
        np.testing.assert_allclose(q.centroid.coords[0], z["centroid"][k], rtol=0, atol=1e-8)


def test_multi_field_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import __graft_entry__ as g
    g.build()
    import field_coverage_path_planning_b200 as fc
    z = np.load(GOLD)
    with pytest.raises(fc.FcppError):
        fc.distance_matrix(np.vstack([z["depot"], z["centroid"]]))
    with pytest.raises(fc.FcppError):
        fc.connection_matrix(z["verts"], z["depot"])


# ------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def fc():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g
    g.build()
    import field_coverage_path_planning_b200 as fc
    return fc


@pytest.mark.gpu
def test_distance_and_connection_matrices_vs_reference_golden(fc):
    z = np.load(GOLD)
    D = fc.distance_matrix(np.vstack([z["depot"], z["centroid"]]))
    # np.linalg.norm of a 2-vector is sqrt(dot(x, x)) (BLAS, possibly fused): <= 1 ulp from sqrt(dx*dx+dy*dy)
    np.testing.assert_allclose(D, z["D"], rtol=4e-16, atol=0)
    assert (np.diag(D) == 0).all() and np.array_equal(D, D.T)
    Cm, fi, ti = fc.connection_matrix(z["verts"], z["depot"])
    np.testing.assert_allclose(Cm, z["C"], rtol=4e-16, atol=0)
    nodes = _nodes(z)
    for a in range(len(nodes)):
        for b in range(len(nodes)):
            assert np.array_equal(nodes[a][fi[a, b]], z["from_pt"][a, b]), (a, b)
            assert np.array_equal(nodes[b][ti[a, b]], z["to_pt"][a, b]), (a, b)


@pytest.mark.gpu
def test_matrices_vs_oracle_at_config4_size(fc):
    """200 fields + depot (BASELINE config 4), seeded quads; ties (shared vertices) resolve to the
    reference's first minimum."""
    from oracle import multi_field as om
    rng = np.random.default_rng(42)
    cen = rng.uniform(0, 5000, size=(200, 2))
    pos = np.vstack([[100.0, 100.0], cen])
    D = fc.distance_matrix(pos)
    np.testing.assert_allclose(D, om.distance_matrix(pos[0], pos[1:]), rtol=4e-16, atol=0)
    half = rng.uniform(20, 150, size=(200, 2))
    verts = np.stack([np.stack([c + (-h[0], -h[1]), c + (h[0], -h[1]), c + (h[0], h[1]), c + (-h[0], h[1])])
                      for c, h in zip(cen, half)])
    verts[7] = verts[3]                     # identical fields: every pair distance ties -> index 0/0 first
    verts[11, 0] = verts[12, 2]             # a shared vertex: distance exactly 0
    Cm, fi, ti = fc.connection_matrix(verts[:40], (100.0, 100.0))
    Co, fo, to = om.connection_matrix(verts[:40], (100.0, 100.0))
    np.testing.assert_allclose(Cm, Co, rtol=4e-16, atol=0)
    assert np.array_equal(fi, fo) and np.array_equal(ti, to)
    assert Cm[12, 13] == 0.0 and Cm[4, 8] == 0.0
    Cm200, _, _ = fc.connection_matrix(verts, (100.0, 100.0))
    assert Cm200.shape == (201, 201) and np.array_equal(Cm200[:41, :41], Cm)
    # empty / single
    C0, _, _ = fc.connection_matrix(np.zeros((0, 4, 2)), (1.0, 2.0))
    assert C0.shape == (1, 1) and C0[0, 0] == 0.0
    assert fc.distance_matrix(np.zeros((1, 2))).shape == (1, 1)


@pytest.mark.gpu
def test_multi_field_planner_v38_drop_in(fc):
    """Same constructor / records / totals as mfp:63-233; the GA order is random in the reference, so
    the sequence is checked structurally and the totals against their definitions."""
    z = np.load(GOLD)
    defs = [{"id": f"F{k:02d}", "vertices": [tuple(map(float, p)) for p in v]} for k, v in enumerate(z["verts"])]
    p = fc.MultiFieldPlannerV38(defs, tuple(z["depot"]), fc.VehicleParams(), num_vehicles=1,
                                optimization_method="genetic", seed=3)
    D, node_ids = p._calculate_distance_matrix()
    assert node_ids == ["depot"] + [d["id"] for d in defs]
    np.testing.assert_allclose(D, z["D"], rtol=1e-12)
    for k, d in enumerate(defs):
        f = p.fields[d["id"]]
        np.testing.assert_allclose(f.centroid, z["centroid"][k], atol=1e-8)
        assert abs(f.area - z["area"][k]) <= 1e-9 * z["area"][k]
        np.testing.assert_allclose(np.array([e[1] for e in f.entry_points]), z["entry_dir"][k], atol=1e-12)
    c = p._find_best_connection("F03", "F07")
    assert c.distance == pytest.approx(z["C"][4, 8], rel=1e-15)
    assert np.array_equal(c.from_point, z["from_pt"][4, 8]) and np.array_equal(c.to_point, z["to_pt"][4, 8])
    route = p.optimize_sequence()
    assert sorted(route.field_sequence) == sorted(d["id"] for d in defs)
    assert len(route.connections) == int(z["route_n_connections"]) == len(defs) + 1
    assert route.connections[0].from_field == "depot" and route.connections[-1].to_field == "depot"
    assert route.total_work_distance == pytest.approx(float(z["route_total_work"]), rel=1e-12)   # area / W, mfp:213-216
    assert route.total_transfer_distance == pytest.approx(sum(c.distance for c in route.connections), rel=1e-15)
    assert route.total_distance == pytest.approx(route.total_transfer_distance + route.total_work_distance)
    assert route.optimization_method == "genetic" and route.optimization_stats["method"] == "genetic"
    # the GA found a decent order: transfer no worse than 1.5x the reference's run on the same instance
    idx = [0] + [1 + int(f[1:]) for f in route.field_sequence]
    tour = sum(z["D"][a, b] for a, b in zip(idx, idx[1:] + idx[:1]))
    rnd = np.mean([sum(z["D"][a, b] for a, b in zip(r, np.roll(r, -1)))
                   for r in (np.random.default_rng(s).permutation(13) for s in range(200))])
    assert tour < 0.75 * rnd
    # "auto" picks 2-opt below 50 fields (mfp:153-162); the reference then imports a module it does not ship (mfp:176,
    # ModuleNotFoundError there) — here the device 2-opt of tsp.py answers (build-defined, == oracle/tsp.py)
    from oracle import tsp as otsp
    p2 = fc.MultiFieldPlannerV38(defs, (0, 0), fc.VehicleParams(), optimization_method="auto")
    assert p2.optimization_method == "2opt"
    r2 = p2.optimize_sequence()
    D2, ids2 = p2._calculate_distance_matrix()
    want, want_len, _ = otsp.two_opt(D2)
    assert r2.field_sequence == [ids2[i] for i in want if ids2[i] != "depot"]
    assert r2.optimization_stats == {'method': '2opt'} and sorted(r2.field_sequence) == sorted(ids2[1:])
    with pytest.raises(ValueError):
        fc.MultiFieldPlannerV38(defs, (0, 0), fc.VehicleParams(), num_vehicles=2).optimize_sequence()
    import multi_field_planner
    assert multi_field_planner.MultiFieldPlannerV38 is fc.MultiFieldPlannerV38


@pytest.mark.gpu
def test_planned_work_distance_closes_the_loop(fc):
    """work_distance='planned': per-field best plan length from plan_batch (4 start corners, argmin)
    == the oracle's best over the same candidates."""
    from oracle import batch as ob, ref_planner as rp
    z = np.load(GOLD)
    defs = [{"id": f"F{k:02d}", "vertices": [tuple(map(float, p)) for p in v]} for k, v in enumerate(z["verts"][:4])]
    p = fc.MultiFieldPlannerV38(defs, tuple(z["depot"]), fc.VehicleParams(), optimization_method="genetic", seed=1,
                                work_distance="planned")
    planned = p.planned_work_lengths()
    for d in defs:
        best = min((ob.evaluate_candidate(d["vertices"], rp.VehicleParams(), R=8.0, start_corner=c, coverage=False)
                    for c in range(4)), key=lambda o: o["len_main"] + o["len_head"])
        assert planned[d["id"]]["length"] == pytest.approx(best["len_main"] + best["len_head"], abs=1e-6)
    route = p.optimize_sequence()
    assert route.total_work_distance == pytest.approx(sum(v["length"] for v in planned.values()), rel=1e-12)
