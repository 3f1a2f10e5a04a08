"""Host-side trace of the pipelined public call (two batches in flight): time spent in submit and in result()
per step, and the device time between the batches' completion events.   python tools/pipeline_trace.py [c2|c3|c5]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import field_coverage_path_planning_b200 as fc  # noqa: E402
from benchmarks import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = wl.WORKLOADS[name](1)
dev = torch.device("cuda", 0)
veh = fc.VehicleParams()


def submit():
    return fc.plan_batch(w.fields, veh, w.axes, obstacles=w.obstacles, outputs=w.outputs, grid_h=w.grid_h, device=dev,
                         winners=True, wait=False)


pend = submit()
for _ in range(6):
    nxt = submit()
    pend.result()
    pend = nxt
pend.result()
torch.cuda.synchronize()
rows = []
t_start = time.perf_counter()
pend = submit()
for k in range(10):
    t0 = time.perf_counter()
    nxt = submit()
    t1 = time.perf_counter()
    r = pend.result()
    t2 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    pend = nxt
pend.result()
torch.cuda.synchronize()
tot = 1e3 * (time.perf_counter() - t_start)
for k, (a, b) in enumerate(rows):
    print(f"step {k}: submit {a:7.3f} ms   result {b:7.3f} ms")
print(f"{name}: {tot / 11:.3f} ms per step over 11 steps -> {w.n_cand / (tot / 11) / 1e3:.3f} M plans/s")
# where result() spends its time
import cProfile
import pstats
pr = cProfile.Profile()
pend = submit()
pr.enable()
for k in range(6):
    nxt = submit()
    r = pend.result()
    pend = nxt
pr.disable()
pend.result()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
