"""Where the end-to-end time of plan_batch(host numpy) goes (host set-up, H2D, kernels, D2H)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import field_coverage_path_planning_b200 as fc
from field_coverage_path_planning_b200.batch import prepare_batch, DeviceBatch, run_device_batch

R, c = bench.global_candidates(1)
cands = {"field_id": np.zeros(len(R), dtype=np.int32), "R": R, "start_corner": c}
veh = fc.VehicleParams()
dev = torch.device("cuda", 0)
def T(f, n=20):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
t, pb = T(lambda: prepare_batch([bench.RECT], veh, cands, [bench.OBST2], None, 0.1, True)); print("prepare_batch ms", t)
t, db = T(lambda: DeviceBatch(pb, dev)); print("DeviceBatch (H2D) ms", t)
t, res = T(lambda: run_device_batch(db, "paths")); print("run_device_batch(paths) ms", t)
t, res = T(lambda: run_device_batch(db, "summary")); print("run_device_batch(summary) ms", t)
t, res = T(lambda: fc.plan_batch([bench.RECT], veh, cands, obstacles=[bench.OBST2], outputs="paths", device=dev)); print("plan_batch(paths) ms", t)
t, res = T(lambda: fc.plan_batch([bench.RECT], veh, cands, obstacles=[bench.OBST2], outputs="summary", device=dev)); print("plan_batch(summary) ms", t)

# ---- finer: the pieces of run_device_batch(paths) with a synchronisation after each ----
import ctypes as C
from field_coverage_path_planning_b200 import _lib
from field_coverage_path_planning_b200.batch import BatchBuffers
h = _lib.handle(0); L = h.lib
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
B, F = db.pb.n_cand, db.pb.n_fields
def piece(name, f, n=20):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f(); torch.cuda.synchronize()
    print(f"  {name}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms")
d_off = torch.empty(B + 1, dtype=torch.int64, device=dev)
db.c.max_points_hint = 0; db.c.max_head_points_hint = 0
piece("fcpp_layout (sync readback inside)", lambda: h.check(L.fcpp_layout(h.h, C.byref(db.c), None, d_off.data_ptr(), stream)))
piece("d_off.cpu()", lambda: d_off.cpu().numpy())
total = int(d_off.cpu().numpy()[-1])
piece("BatchBuffers alloc", lambda: BatchBuffers(dev, B, F, total, False))
bufs = BatchBuffers(dev, B, F, total, False); bufs.d_off = d_off
out = _lib.Outputs(); out.summary = bufs.d_sum.data_ptr(); out.offsets = d_off.data_ptr(); out.path_xy = bufs.d_path.data_ptr(); out.speeds_kmh = bufs.d_spd.data_ptr()
def planb():
    h.check(L.fcpp_layout(h.h, C.byref(db.c), None, d_off.data_ptr(), stream))
    h.check(L.fcpp_plan_batch(h.h, C.byref(db.c), C.byref(out), stream))
piece("layout + plan_batch kernels", planb)
piece("argmin", lambda: h.check(L.fcpp_field_argmin(h.h, bufs.d_sum.data_ptr(), db.t["cand_field"].data_ptr(), B, F, 0, 0, bufs.d_cost.data_ptr(), bufs.d_best.data_ptr(), stream)))
piece("d_sum.cpu()", lambda: bufs.d_sum.cpu())
db.c.max_points_hint = int(L.fcpp_last_max_points(h.h)); db.c.max_head_points_hint = int(L.fcpp_last_max_head_points(h.h))
piece("layout + plan_batch kernels (hinted, async)", planb)
