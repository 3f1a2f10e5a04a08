#!/usr/bin/env python
"""Measure every BASELINE.json config on one GPU (CUDA events, inputs resident) next to the CPU
oracle on a bounded subsample.  Not the driver's bench (that is ../bench.py, config 2); this fills
the per-config table of BASELINE.md / profiles/.

    python benchmarks/configs.py [--configs c1,c2,c3,c4,c5] [--out gpurun_out/configs.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RECT = [(0.0, 0.0), (500.0, 0.0), (500.0, 200.0), (0.0, 200.0)]
OBST2 = [[(200, 80), (250, 80), (250, 120), (200, 120)], [(350, 140), (380, 140), (380, 170), (350, 170)]]


def c3_fields(F=4096, seed=1234):
    """seeded tilted parallelograms, CCW from the lower-left vertex."""
    rng = np.random.default_rng(seed)
    L, Wd = rng.uniform(200, 800, F), rng.uniform(100, 400, F)
    sx, phi = rng.uniform(-0.4, 0.4, F) * Wd, rng.uniform(0, np.pi, F)
    org = rng.uniform(0, 5000, (F, 2))
    q = np.stack([np.zeros((F, 2)), np.stack([L, np.zeros(F)], 1), np.stack([L + sx, Wd], 1), np.stack([sx, Wd], 1)], 1)
    c, s = np.cos(phi)[:, None], np.sin(phi)[:, None]
    x = q[:, :, 0] * c - q[:, :, 1] * s + org[:, :1]
    y = q[:, :, 0] * s + q[:, :, 1] * c + org[:, 1:]
    return np.stack([x, y], axis=2)


def time_gpu(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def cpu_rate(items, fn):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(fn, items[:cores])
        t0 = time.perf_counter()
        pool.map(fn, items, chunksize=1)
        dt = time.perf_counter() - t0
    return len(items) / dt, cores


def _cpu_c(args):
    from oracle import batch as ob, ref_planner as rp
    verts, R, heading, corner, obst, h = args
    ob.evaluate_candidate(verts, rp.VehicleParams(), R=R, heading=heading, start_corner=corner, obstacles=obst, grid_h=h)
    return 0


def batch_case(fc, name, fields, cand, obstacles, grid_h, outputs, cpu_n):
    import torch
    from field_coverage_path_planning_b200.batch import BatchBuffers, DeviceBatch, prepare_batch, run_device_batch
    dev = torch.device("cuda", 0)
    veh = fc.VehicleParams()
    pb = prepare_batch(fields, veh, cand, obstacles, None, grid_h, True)
    db = DeviceBatch(pb, dev)
    first = run_device_batch(db, outputs)
    assert (first.summary["status"] == 0).all(), np.unique(first.summary["status"])
    total = int(first.offsets[-1]) if outputs == "paths" else 0
    bufs = BatchBuffers(dev, pb.n_cand, pb.n_fields, total)
    ms = time_gpu(lambda: run_device_batch(db, outputs, buffers=bufs, fetch=False))
    # per-kernel CUDA-event times of one more step (layout+scan, plan, coverage)
    import ctypes as C
    from field_coverage_path_planning_b200 import _lib
    h = _lib.handle(0)
    h.check(h.lib.fcpp_set_profiling(h.h, 1))
    run_device_batch(db, outputs, buffers=bufs, fetch=False)
    torch.cuda.synchronize()
    ms3 = (C.c_float * 3)()
    h.check(h.lib.fcpp_kernel_times(h.h, C.byref(ms3)))
    h.check(h.lib.fcpp_set_profiling(h.h, 0))
    t0 = time.perf_counter()
    for _ in range(3):
        fc.plan_batch(fields, veh, cand, obstacles=obstacles, outputs=outputs, grid_h=grid_h)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 3
    B = pb.n_cand
    idx = np.linspace(0, B - 1, cpu_n).astype(int)
    items = [(np.asarray(fields)[int(cand["field_id"][i])].tolist(), float(cand["R"][i]) if "R" in cand else None,
              float(cand["heading"][i]) if "heading" in cand else None,
              int(cand["start_corner"][i]) if "start_corner" in cand else None,
              (obstacles[int(cand["field_id"][i])] if obstacles else ()), grid_h) for i in idx]
    cr, cores = cpu_rate(items, _cpu_c)
    s = first.summary
    return {"config": name, "candidates": B, "outputs": outputs, "grid_h": grid_h,
            "points_per_plan_mean": float((s["n_main"] + s["n_head"]).mean()),
            "band_cells_mean": float(s["cov_total"].mean()),
            "gpu_ms": ms, "kernel_ms": {"layout": ms3[0], "plan": ms3[1], "cover": ms3[2]},
            "gpu_plans_per_s": B / (ms / 1e3), "e2e_s": e2e, "e2e_plans_per_s": B / e2e,
            "cpu_plans_per_s": cr, "cpu_cores": cores, "cpu_sample": cpu_n,
            "n_boundary_viol_mean": float(s["n_boundary_viol"].mean()),
            "coverage_rate_mean": float((s["cov_cells"] / np.maximum(s["cov_total"], 1)).mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--c3-fields", type=int, default=4096)
    ap.add_argument("--c5-cands", type=int, default=8192, help="per-GPU shard of the 65536 candidates")
    args = ap.parse_args()
    import torch
    import field_coverage_path_planning_b200 as fc
    from oracle import raster
    raster.build()
    want = args.configs.split(",")
    out = []
    if "c1" in want:   # single plan() latency (README.md:196 publishes 0.046 s)
        for (L, Wd, sp) in ((100, 80, (90, 70)), (500, 200, (50, 180)), (3500, 320, (3400, 300))):
            p = fc.TwoLayerPathPlannerV37(fc.VehicleParams(3.2, 8.0, 9.0, 15.0), field_length=L, field_width=Wd,
                                          start_point=sp)
            p.plan_complete_coverage()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                r = p.plan_complete_coverage()
            dt = (time.perf_counter() - t0) / 20
            out.append({"config": f"c1 single plan {L}x{Wd}", "latency_ms": dt * 1e3,
                        "n_main": len(r["main_work"]["path"]), "n_head": len(r["headland"]["path"]),
                        "published_reference_s": 0.046 if L == 500 else None})
            print(out[-1], flush=True)
    if "c2" in want:
        cand = fc.make_candidates(1, radii=np.linspace(5.0, 12.0, 1024), start_corners=[0, 1, 2, 3])
        out.append(batch_case(fc, "c2 500x200 + 2 obstacles, 4 corners x 1024 radii", [RECT], cand, [OBST2], 0.1,
                              "paths", 64))
        print(out[-1], flush=True)
    if "c3" in want:
        F = args.c3_fields
        fields = c3_fields(F)
        cand = fc.make_candidates(F, headings=np.deg2rad(np.arange(180.0)))
        out.append(batch_case(fc, f"c3 {F} parallelograms x 180 headings (summary-only, argmin per field)", fields,
                              cand, None, 0.1, "summary", 64))
        print(out[-1], flush=True)
    if "c4" in want:
        rng = np.random.default_rng(42)
        pos = np.vstack([[100.0, 100.0], rng.uniform(0, 5000, size=(200, 2))])
        D = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
        prng = np.random.default_rng(7)
        pop = np.array([prng.permutation(201) for _ in range(8192)], dtype=np.int32)
        dD, dP = torch.from_numpy(D).cuda(), torch.from_numpy(pop).cuda()
        ms = time_gpu(lambda: fc.tour_lengths(dD, dP), reps=20)
        t0 = time.perf_counter()
        for _ in range(5):
            fc.tour_lengths(D, pop)
        e2e = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        ref = np.array([sum(D[r[i], r[(i + 1) % 201]] for i in range(201)) for r in pop[:256]])  # ga:174-181 loop
        cpu = 256 / (time.perf_counter() - t0)
        assert np.array_equal(ref, fc.tour_lengths(D, pop[:256]))
        out.append({"config": "c4 GA fitness, 201 nodes, population 8192", "gpu_ms": ms,
                    "gpu_tours_per_s": 8192 / (ms / 1e3), "e2e_tours_per_s": 8192 / e2e,
                    "cpu_tours_per_s_1core_python_loop": cpu})
        print(out[-1], flush=True)
        # whole generations on the device (N1): selection + OX + mutation + elitism + fitness + tracking
        from field_coverage_path_planning_b200 import ga
        G = 200
        cfg = fc.GAConfig(population_size=8192, max_generations=G, convergence_threshold=10 ** 9)
        ga.ga_solve_device(cfg, dD, seed=1, initial_population=dP)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        route, stats, hist = ga.ga_solve_device(cfg, dD, seed=1, initial_population=dP)
        dt = time.perf_counter() - t0
        # the reference's generation on one core: its own operators on a 512-individual slice, scaled
        cpu_gen = None
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ref_ga", "/root/reference/genetic_algorithm_solver.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sol = mod.GeneticAlgorithmSolver(mod.GAConfig(population_size=512))
            P = [list(map(int, r)) for r in pop[:512]]
            t1 = time.perf_counter()
            fit = [sol._calculate_fitness(r, D) for r in P]
            new = sol._elitism(P, sol._mutation(sol._crossover(sol._selection(P, fit))), fit)
            cpu_gen = (time.perf_counter() - t1) * (8192 / 512)
        except Exception:       # the GPU box has no /root/reference
            pass
        out.append({"config": "c4 GA whole generations on the device, 201 nodes, population 8192", "generations": G,
                    "wall_s": dt, "generations_per_s": G / dt, "ms_per_generation": 1e3 * dt / G,
                    "best_distance": stats["best_distance"], "first_best": float(1 / hist[0, 0]),
                    "reference_cpu_s_per_generation_1core_scaled_from_512": cpu_gen})
        print(out[-1], flush=True)
    if "c5" in want:
        big = [(0.0, 0.0), (2000.0, 0.0), (2000.0, 1000.0), (0.0, 1000.0)]
        n = args.c5_cands // 4
        cand = fc.make_candidates(1, radii=np.linspace(5.0, 12.0, 16384)[:n], start_corners=[0, 1, 2, 3])
        out.append(batch_case(fc, f"c5 2000x1000, h=0.05, {4 * n} candidates (one GPU's shard of 65536)", [big], cand,
                              None, 0.05, "summary", 16))
        print(out[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
