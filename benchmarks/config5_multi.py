#!/usr/bin/env python
"""BASELINE config 5: one 2000 m x 1000 m field at a 0.05 m coverage grid, 65 536 candidates
(4 start corners x 16 384 radii) sharded over the GPUs of one box, per-field argmin through NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P benchmarks/config5_multi.py [--cands 65536] [--reps 3]

Every rank passes the GLOBAL candidate set to the public API (plan_batch(distributed=True)); the
ranks plan their contiguous shards and exchange only the per-field best.  Rank 0 prints one JSON line
with wall-clock plans/s (synchronised, max over ranks) and the properties checked on the result."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cands", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import field_coverage_path_planning_b200 as fc
    big = [(0.0, 0.0), (2000.0, 0.0), (2000.0, 1000.0), (0.0, 1000.0)]
    cand = fc.make_candidates(1, radii=np.linspace(5.0, 12.0, args.cands // 4), start_corners=[0, 1, 2, 3])

    def run():
        return fc.plan_batch([big], fc.VehicleParams(), cand, outputs="summary", grid_h=0.05, device=dev,
                             distributed=True)
    res = run()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        res = run()
    torch.cuda.synchronize(dev)
    dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / args.reps], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # properties that do not depend on the size: every rank agrees on the winner; the winner's record is
    # the best of the local summaries of its owner; coverage counts are sane
    lo, hi = res.extras["shard"]
    s = res.summary
    cost = s["len_main"] + s["len_head"]
    ok = s["status"] == 0
    local_best = float(np.min(cost[ok])) if ok.any() else np.inf
    glob = torch.tensor([local_best], dtype=torch.float64, device=dev)
    dist.all_reduce(glob, op=dist.ReduceOp.MIN)
    win = res.extras["winner_summary"][0]
    assert float(res.best_cost[0]) == float(glob.item()), (res.best_cost[0], glob.item())
    assert float(win["len_main"] + win["len_head"]) == float(res.best_cost[0])
    if lo <= int(res.best_cand[0]) < hi:
        assert s[int(res.best_cand[0]) - lo].tobytes() == win.tobytes()
    assert (s["cov_cells"][ok] <= s["cov_total"][ok]).all() and (s["n_boundary_viol"][ok] == 0).all()
    if rank == 0:
        print(json.dumps({"config": "c5 2000x1000 m, h=0.05 m, NCCL argmin", "n_gpus": world, "candidates": len(cand["R"]),
                          "wall_ms": 1e3 * float(dt.item()), "plans_per_s_end_to_end": len(cand["R"]) / float(dt.item()),
                          "best_candidate": int(res.best_cand[0]), "best_cost_m": float(res.best_cost[0]),
                          "winner_R": float(cand["R"][int(res.best_cand[0])]),
                          "winner_coverage": float(win["cov_cells"] / max(int(win["cov_total"]), 1)),
                          "band_cells_winner": int(win["cov_total"]), "points_winner": int(win["n_main"] + win["n_head"])}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
