import sys, numpy as np
sys.path.insert(0, '/root/repo')
import field_coverage_path_planning_b200 as fc
RECT = [(0, 0), (500, 0), (500, 200), (0, 200)]
cand = fc.make_candidates(1, radii=np.linspace(5.0, 12.0, 1024), start_corners=[0, 1, 2, 3])
res = fc.plan_batch([RECT], fc.VehicleParams(), cand)
import torch; torch.cuda.synchronize()
print(res.summary["cov_cells"][:4])
