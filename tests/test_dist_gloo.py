"""Multi-GPU host logic on CPU: world_size-2 gloo process group exercising the candidate
sharding and the exact per-field argmin reduction of field_coverage_path_planning_b200/dist.py
(SURVEY.md §8(e)).  The per-rank 'summaries' are produced by the CPU oracle — the CUDA compute
itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

RECT = [(0, 0), (500, 0), (500, 200), (0, 200)]
SMALL = [(0, 0), (100, 0), (100, 80), (0, 80)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import field_coverage_path_planning_b200 as fc
    from field_coverage_path_planning_b200 import _lib, dist as fdist
    from oracle import batch as ob, ref_planner as rp

    fields = [RECT, SMALL, [(0, 0), (12, 0), (12, 9), (0, 9)]]            # field 2 has no valid candidate
    cand = fc.make_candidates(3, radii=[5.0, 8.0, 8.0, 11.0], start_corners=[0, 2])
    n = len(cand["field_id"])
    local, lo = fdist.shard_candidates(cand, world, rank)
    hi = lo + len(local["field_id"])
    assert (lo, hi) == fdist.shard_range(n, world, rank)
    # local "device" results from the oracle
    summ = np.zeros(hi - lo, dtype=_lib.SUMMARY_DTYPE)
    for i in range(hi - lo):
        o = ob.evaluate_candidate(fields[int(local["field_id"][i])], rp.VehicleParams(), R=local["R"][i],
                                  start_corner=int(local["start_corner"][i]), coverage=False)
        summ["status"][i] = o["status"]
        if not o["status"]:
            summ["len_main"][i], summ["len_head"][i] = o["len_main"], o["len_head"]
            summ["n_main"][i] = o["n_main"]
    F = 3
    cost = np.full(F, np.inf)
    best = np.full(F, -1, dtype=np.int64)
    for i in range(hi - lo):
        f = int(local["field_id"][i])
        c = summ["len_main"][i] + summ["len_head"][i]
        if summ["status"][i] == 0 and c < cost[f]:
            cost[f], best[f] = c, lo + i
    t_cost, t_best = torch.from_numpy(cost.copy()), torch.from_numpy(best.copy())
    fdist.reduce_best(t_cost, t_best)
    win = fdist.gather_winner_records(torch.from_numpy(summ.view(np.uint8).reshape(-1).copy()), t_best, lo, hi)
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.concatenate([t_cost.numpy(), t_best.numpy().astype(np.float64),
                            win.numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)["n_main"].astype(np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_argmin_world2(tmp_path):
    import __graft_entry__ as g
    g.build()
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)                                         # every rank holds the global result
    # single-process reference of the same reduction
    import field_coverage_path_planning_b200 as fc
    from oracle import batch as ob, ref_planner as rp
    fields = [RECT, SMALL, [(0, 0), (12, 0), (12, 9), (0, 9)]]
    cand = fc.make_candidates(3, radii=[5.0, 8.0, 8.0, 11.0], start_corners=[0, 2])
    cost = np.full(3, np.inf)
    best = np.full(3, -1.0)
    nmain = np.zeros(3)
    for i in range(len(cand["field_id"])):
        f = int(cand["field_id"][i])
        o = ob.evaluate_candidate(fields[f], rp.VehicleParams(), R=cand["R"][i], start_corner=int(cand["start_corner"][i]),
                                  coverage=False)
        c = ob.candidate_cost(o)
        if c < cost[f]:                                                   # strict: ties keep the lowest index
            cost[f], best[f], nmain[f] = c, i, o["n_main"]
    assert np.array_equal(r0[:3], cost) and np.array_equal(r0[3:6], best) and np.array_equal(r0[6:], nmain)
    assert best[2] == -1 and np.isinf(cost[2])
    # the duplicated radius 8.0 produces exact ties inside field 0/1: the lowest index must win
    assert best[0] == np.argmin([ob.candidate_cost(ob.evaluate_candidate(RECT, rp.VehicleParams(), R=cand["R"][i],
                                 start_corner=int(cand["start_corner"][i]), coverage=False)) for i in range(8)])


def test_shard_range_partition():
    from field_coverage_path_planning_b200 import dist as fdist
    for n in (0, 1, 7, 4096, 65536, 737280):
        for w in (1, 2, 3, 8):
            r = [fdist.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
