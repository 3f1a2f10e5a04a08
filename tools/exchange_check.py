"""torchrun -N check: fcpp_field_argmin_exchange (peer-memory exchange + merge in one kernel) against the
NCCL all-gather + merge path, many back-to-back calls (double buffering), several field counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from field_coverage_path_planning_b200 import dist as fdist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for F in (1, 7, 4096):
    ex = fdist._PeerExchange.get(dev, F, None)
    if rank == 0:
        print("F", F, "peer exchange available:", ex.ok, flush=True)
    g = torch.Generator(device="cpu").manual_seed(100 * F + rank)
    for it in range(40):
        cost = torch.randint(0, 5, (F,), generator=g).double()          # many ties across ranks
        cand = torch.randint(-1, 1000, (F,), generator=g).long() * world + rank
        cand[cand < 0] = -1
        cb = torch.empty(2 * F, dtype=torch.int64, device=dev)
        c1, k1 = cb[:F].view(torch.float64), cb[F:]
        c1.copy_(cost); k1.copy_(cand)
        c2, k2 = cost.to(dev), cand.to(dev)
        fdist.reduce_best(c1, k1, peer=True)
        fdist.reduce_best(c2, k2, peer=False)
        if not (torch.equal(c1, c2) and torch.equal(k1, k2)):
            ok = False
            print("MISMATCH rank", rank, "F", F, "it", it, flush=True)
            break
# timing: 200 back-to-back exchanges
for peer in (True, False):
    cb = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fdist.reduce_best(cb[:1].view(torch.float64), cb[1:], peer=peer)
    torch.cuda.synchronize()
    if rank == 0:
        print("peer" if peer else "nccl", f"{(time.perf_counter() - t0) / 200 * 1e6:.1f} us per reduce_best", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("EXCHANGE CHECK", "OK" if t.item() else "FAILED", flush=True)
dist.destroy_process_group()
