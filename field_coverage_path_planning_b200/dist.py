"""Multi-GPU: one process per GPU (torch.distributed), candidates sharded in contiguous ranges,
NO data-path collective — the only exchange is the final per-field argmin.

The reduction is exact and deterministic (ties go to the lowest global candidate index):
    1+2. CUDA: ONE all-gather of every rank's (best cost, best candidate) words and the library's
       merge kernel (fcpp_field_argmin_merge).  CPU/gloo (tests): all_reduce(MIN) of the cost, then
       all_reduce(MIN) of the candidate index where the local best equals the global minimum;
    3. the winner's 176-byte summary record is contributed by its owner (fcpp_winner_records) and summed
       as int32 words (every other rank contributes zeros), i.e. an all-gather of one record per field.
Works on NCCL (CUDA tensors) and gloo (CPU tensors — used by the world_size-2 CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

I64_MAX = torch.iinfo(torch.int64).max


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n candidates for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_candidates(cands: Dict[str, np.ndarray], world: int, rank: int) -> Tuple[Dict[str, np.ndarray], int]:
    """The rank's contiguous shard of a candidate set + its first global index.  Factored sets
    (batch.candidate_axes) keep their axes and get a ``range`` — nothing is sliced or copied."""
    from .batch import axes_count, is_axes
    if is_axes(cands):
        first = int(cands.get("range", (0, 0))[0])
        lo, hi = shard_range(axes_count(cands), world, rank)
        return dict(cands, range=(first + lo, first + hi)), first + lo
    n = len(cands["field_id"])
    lo, hi = shard_range(n, world, rank)
    return {k: v[lo:hi] for k, v in cands.items()}, lo


class _PeerExchange:
    """Symmetric-memory buffers for fcpp_field_argmin_exchange: per rank [2][world][2 F] int64 slots + a
    flag array, mapped into every process of the group (torch symmetric memory over NVLink P2P).  One
    instance per (device, F, group); set-up is a collective.  ``ok`` is False — on EVERY rank — when
    symmetric memory is unavailable on any of them, and reduce_best falls back to NCCL."""
    _cache: Dict[tuple, "_PeerExchange"] = {}

    def __init__(self, dev: torch.device, F: int, group):
        import ctypes as C
        self.ok = False
        self.epoch = 0
        world = dist.get_world_size(group)
        good = 0
        try:
            import torch.distributed._symmetric_memory as symm
            if world <= 16:
                words = 2 * world * 2 * F
                self.nbytes = words * 8 + 256
                self.t = symm.empty(self.nbytes, dtype=torch.uint8, device=dev)
                self.t.zero_()
                self.hdl = symm.rendezvous(self.t, dist.group.WORLD if group is None else group)
                ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                self.bufs = (C.c_uint64 * world)(*ptrs)
                self.flags = (C.c_uint64 * world)(*[p + words * 8 for p in ptrs])
                self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
                good = 1
        except Exception:           # no symmetric memory in this build / on this box
            good = 0
        agree = torch.tensor([good], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)     # also orders the zeroing before any use
        torch.cuda.synchronize(dev)
        self.ok = bool(agree.item())

    @classmethod
    def get(cls, dev: torch.device, F: int, group) -> "_PeerExchange":
        key = (dev.index, F, id(group))
        ex = cls._cache.get(key)
        if ex is None:
            ex = cls._cache[key] = cls(dev, F, group)
        return ex


def reduce_best(best_cost: torch.Tensor, best_cand: torch.Tensor, group=None, peer: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global per-field (cost, candidate) from the local ones.  ``best_cand`` holds GLOBAL
    candidate indices (-1 = this rank has no valid candidate for the field).  In place.

    CUDA tensors: one NCCL all-gather of the (cost, candidate) words + the library's merge kernel
    (fcpp_field_argmin_merge).  ``peer=True`` / FCPP_PEER_EXCHANGE=1 (opt-in): ONE kernel that exchanges
    the words over peer memory and merges them (fcpp_field_argmin_exchange) — bit-identical, 16 vs 42 us
    per isolated call at N=2, but measured slower inside the 8-GPU bench loop, hence not
    the default.  CPU tensors (gloo tests): the same rule with torch ops."""
    if best_cost.is_cuda:
        import ctypes as C
        F = best_cost.numel()
        world = dist.get_world_size(group)
        dev = best_cost.device
        use_peer = (os.environ.get("FCPP_PEER_EXCHANGE", "0") == "1") if peer is None else peer
        if (use_peer and world > 1 and best_cost.dtype == torch.float64 and best_cand.dtype == torch.int64
                and best_cost.is_contiguous() and best_cand.is_contiguous()):
            ex = _PeerExchange.get(dev, F, group)
            if ex.ok:
                ex.epoch += 1
                h = _lib.handle(dev.index)
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                h.check(h.lib.fcpp_field_argmin_exchange(h.h, ex.world, ex.rank, F, ex.epoch, ex.bufs, ex.flags,
                                                         best_cost.data_ptr(), best_cand.data_ptr(), st))
                return best_cost, best_cand
        adjacent = (best_cost.dtype == torch.float64 and best_cand.dtype == torch.int64 and best_cost.is_contiguous()
                    and best_cand.is_contiguous() and best_cand.data_ptr() == best_cost.data_ptr() + 8 * F
                    and best_cost.untyped_storage().data_ptr() == best_cand.untyped_storage().data_ptr())
        if adjacent:   # BatchBuffers lays them out back to back
            send = torch.empty(0, dtype=torch.int64, device=dev).set_(
                best_cost.untyped_storage(), best_cost.storage_offset(), (2 * F,))
        else:
            send = torch.cat([best_cost.view(torch.int64), best_cand])
        gathered = torch.empty(world * 2 * F, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, send, group=group)
        h = _lib.handle(dev.index)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_field_argmin_merge(h.h, gathered.data_ptr(), world, F, best_cost.data_ptr(),
                                              best_cand.data_ptr(), st))
        return best_cost, best_cand
    local_cost = best_cost.clone()
    dist.all_reduce(best_cost, op=dist.ReduceOp.MIN, group=group)
    mine = (local_cost == best_cost) & (best_cand >= 0)
    key = torch.where(mine, best_cand, torch.full_like(best_cand, I64_MAX))
    dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    best_cand.copy_(torch.where(key == I64_MAX, torch.full_like(key, -1), key))
    return best_cost, best_cand


def gather_winner_records(summary_bytes: torch.Tensor, best_cand: torch.Tensor, lo: int, hi: int,
                          group=None, count_mask: int = 0) -> torch.Tensor:
    """[F, 176] uint8: the summary record of every field's global winner.
    ``summary_bytes`` = this rank's records as a flat uint8 tensor ((hi-lo)*176 bytes).
    CUDA: one library kernel writes the records this rank owns (zeros elsewhere), one all-reduce sums them;
    CPU tensors (gloo tests): the same with torch ops.
    ``count_mask`` != 0 (CUDA): the returned tensor is flat, F * 176 + 16 bytes — behind the records rides one
    int32 word, the number of candidates OF ALL RANKS whose status has a bit of the mask (fcpp_status_count; the
    same all-reduce sums it): every rank learns in the same read-back whether any rank's launch sizes were too
    small, so that all of them repeat the batch together."""
    rec = _lib.SUMMARY_DTYPE.itemsize
    F = best_cand.numel()
    if best_cand.is_cuda:
        import ctypes as C
        dev = best_cand.device
        out = torch.empty(F * rec + (16 if count_mask else 0), dtype=torch.uint8, device=dev)
        h = _lib.handle(dev.index)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_winner_records(h.h, summary_bytes.data_ptr() if hi > lo else None, lo, hi,
                                          best_cand.data_ptr(), F, out.data_ptr(), st))
        if count_mask:
            out[F * rec + 4:].zero_()
            h.check(h.lib.fcpp_status_count(h.h, summary_bytes.data_ptr() if hi > lo else None, hi - lo, count_mask,
                                            out.data_ptr() + F * rec, st))
        dist.all_reduce(out.view(torch.int32), op=dist.ReduceOp.SUM, group=group)
        return out if count_mask else out.view(F, rec)
    own = (best_cand >= lo) & (best_cand < hi)
    idx = torch.where(own, best_cand - lo, torch.zeros_like(best_cand))
    rows = summary_bytes.view(-1, rec)[idx] if hi > lo else torch.zeros((F, rec), dtype=torch.uint8,
                                                                        device=best_cand.device)
    rows = torch.where(own[:, None], rows, torch.zeros_like(rows)).contiguous()
    words = rows.view(torch.int32)
    dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words.view(torch.uint8).view(F, rec)


class PendingShardedBatch:
    """A sharded batch in flight on every rank (``plan_batch(..., distributed=True, wait=False)``).  ``result()`` is
    COLLECTIVE in one case only: when some rank's remembered launch sizes did not fit its shard (every rank sees
    that in the flag word that rides on the winners' records), all ranks repeat the batch together."""

    def __init__(self, db, pf, hw, outputs, args, lo, hi, speculated):
        self.db, self.pf, self.hw, self.outputs, self.args = db, pf, hw, outputs, args
        self.lo, self.hi, self.speculated = lo, hi, speculated
        self._res = None

    def __del__(self):
        try:        # an abandoned batch: wait for its read-back before the buffers go back to their pools
            if self._res is None and self.pf is not None:
                self.pf.event.synchronize()
        except Exception:
            pass

    def result(self):
        from .batch import _Hints, _finish_fetch, _side_stream, fetch_winner_paths
        if self._res is not None:
            return self._res
        db, pf, outputs = self.db, self.pf, self.outputs
        want_curvature, cost, winners = self.args
        pb = db.pb
        F = pb.n_fields
        rec = _lib.SUMMARY_DTYPE.itemsize
        res = _finish_fetch(pf)                                    # ONE synchronisation
        hv = self.hw[1]
        if self.speculated and int(hv[F * rec:F * rec + 4].view(np.int32)[0]) != 0:
            # a rank's remembered sizes were too small (or a candidate is genuinely too large): every rank sees the
            # same word and launches the batch again, sized from its own layout pass
            _Hints.drop(_Hints.key(db, outputs, want_curvature))
            self._res = _submit_sharded(db, outputs, want_curvature, cost, winners, self.lo, self.hi, False).result()
            return self._res
        res.extras["winner_summary"] = hv[:F * rec].view(_lib.SUMMARY_DTYPE).reshape(-1)[:F]
        res.extras["shard"] = (self.lo, self.hi)
        res.extras["h2d_bytes"] = pb.h2d_bytes()
        res.extras["d2h_bytes"] = (len(res.summary) * rec + 16 * F + F * rec + 16
                                   + ((pb.n_cand + 1) * 8 if outputs == "paths" else 0))
        res.extras["speculative"] = bool(pf.buffers.speculative)
        if winners:
            with torch.cuda.stream(_side_stream(db.dev)):      # not behind later batches in flight
                res.extras["d2h_bytes"] += fetch_winner_paths(db, res, outputs, slot=db.slot + 1)
        self._res = res
        return res


def _submit_sharded(db, outputs, want_curvature, cost, winners, lo, hi, speculate):
    """Kernels of this rank's shard, the two collectives and the read-back, all enqueued on the current stream."""
    from .batch import _SIZE_FLAGS, _ResultPool, _Staging, _enqueue_fetch, _launch_device_batch
    dev = db.dev
    F = db.pb.n_fields
    rec = _lib.SUMMARY_DTYPE.itemsize
    bufs, offsets = _launch_device_batch(db, outputs, want_curvature, cost, lo, None, speculate=speculate)
    reduce_best(bufs.d_cost, bufs.d_best)                   # in place: d_cost / d_best now hold the GLOBAL result
    # every rank passes the mask whether or not IT speculated: the all-reduce must have the same shape everywhere
    win = gather_winner_records(bufs.d_sum, bufs.d_best, lo, hi, count_mask=_SIZE_FLAGS)
    with torch.cuda.device(dev):                            # the winners' records ride on the same read-back
        nb = F * rec + 16
        pooled = _ResultPool.get(nb)
        if pooled is None:
            t = torch.empty(nb, dtype=torch.uint8).pin_memory()
            pooled = (t, t.numpy())
        pooled[0][:nb].copy_(win, non_blocking=True)
    pf = _enqueue_fetch(db, bufs, outputs, offsets, True, lo)
    # (the numpy view of a pooled buffer must stay referenced: the pool hands a buffer out again once its view is gone)
    return PendingShardedBatch(db, pf, pooled, outputs, (want_curvature, cost, winners), lo, hi, bool(speculate))


def plan_batch_sharded(fields, vehicle, candidates, obstacles, start_points, outputs, grid_h, coverage, cost,
                       device, want_curvature, turn_model="arc", clothoid_share=0.5, winners=False, wait=True):
    """plan_batch over all ranks of the default process group (see batch.plan_batch): every rank plans its
    contiguous shard; ONE all-gather + merge kernel gives every rank the per-field argmin, one kernel + one
    all-reduce every winner's summary record; the local summaries, the merged argmin and the winners' records
    come back to the host with ONE synchronisation.  Launch sizes are remembered from earlier batches of the same
    shape (no layout read-back); ``wait=False`` returns a ``PendingShardedBatch``."""
    from .batch import DeviceBatch, _dev, axes_count, is_axes, prepare_batch
    if not dist.is_initialized():
        raise RuntimeError("distributed=True needs torch.distributed.init_process_group (backend 'nccl')")
    world, rank = dist.get_world_size(), dist.get_rank()
    fv = np.asarray(fields, dtype=np.float64).reshape(-1, 4, 2)
    if candidates is None:
        candidates = {"field_id": np.arange(len(fv), dtype=np.int32)}
    n = axes_count(candidates) if is_axes(candidates) else len(candidates["field_id"])
    local, lo = shard_candidates(candidates, world, rank)
    hi = lo + (axes_count(local) if is_axes(local) else len(local["field_id"]))
    sp = None if start_points is None else np.asarray(start_points, dtype=np.float64).reshape(n, 2)[lo:hi]
    dev = _dev(device)
    pb = prepare_batch(fv, vehicle, local, obstacles, sp, grid_h, coverage, turn_model, clothoid_share)
    db = DeviceBatch(pb, dev)
    pend = _submit_sharded(db, outputs, want_curvature, cost, winners, lo, hi, True)
    return pend if not wait else pend.result()
