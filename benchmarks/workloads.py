"""The BASELINE.json batch configurations as data: shared by bench.py, the
per-config runner (configs.py) and the full-size parity tests, so that what is timed and what is
parity-checked are the same candidate sets.

    c2  500 m x 200 m field + the two obstacles of mlp3:1629-1632, 4 start corners x 1024 radii
        linspace(5, 12) per GPU, h = 0.1 m, paths materialised
    c3  F seeded tilted parallelograms x 180 headings (1 degree steps), h = 0.1 m, summary only,
        argmin per field
    c5  2000 m x 1000 m field, h = 0.05 m, 4 start corners x radii linspace(5, 12, 16384)
        (65 536 candidates over 8 GPUs = 8192 per GPU), summary only
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

RECT = [(0.0, 0.0), (500.0, 0.0), (500.0, 200.0), (0.0, 200.0)]
OBST2 = [[(200, 80), (250, 80), (250, 120), (200, 120)], [(350, 140), (380, 140), (380, 170), (350, 170)]]
BIG = [(0.0, 0.0), (2000.0, 0.0), (2000.0, 1000.0), (0.0, 1000.0)]
CORNERS = [0, 1, 2, 3]


@dataclass
class Workload:
    name: str
    text: str
    fields: np.ndarray                  # [F, 4, 2]
    cands: Dict[str, np.ndarray]        # the WHOLE job (all GPUs); shard with dist.shard_candidates
    obstacles: Optional[List]
    grid_h: float
    outputs: str                        # 'paths' | 'summary'
    axes: Optional[Dict[str, object]] = None   # the same candidates, same order, in factored form (batch.candidate_axes)

    @property
    def n_cand(self) -> int:
        return len(self.cands["field_id"])

    def oracle_args(self, i: int):
        """(verts, R, heading, start_corner, obstacles, grid_h) of global candidate i for oracle.batch."""
        f = int(self.cands["field_id"][i])
        return (np.asarray(self.fields)[f].tolist(),
                float(self.cands["R"][i]) if "R" in self.cands else None,
                float(self.cands["heading"][i]) if "heading" in self.cands else None,
                int(self.cands["start_corner"][i]) if "start_corner" in self.cands else None,
                (self.obstacles[f] if self.obstacles else ()), self.grid_h)


def _radius_corner(radii: np.ndarray, n_gpus: int):
    """radius-major x start corner, enumerated shard-major (shard r = radii[r::N]): every GPU's contiguous
    shard spans the whole radius range — plan cost grows with R (more headland loops, larger corner
    windows), and radius-sorted contiguous shards would leave the last rank ~25 % more work."""
    radii = np.concatenate([radii[r::n_gpus] for r in range(n_gpus)])
    R = np.repeat(radii, len(CORNERS))
    c = np.tile(np.asarray(CORNERS, dtype=np.int32), len(radii))
    return {"field_id": np.zeros(len(R), dtype=np.int32), "R": R, "start_corner": c}, \
        {"axes": True, "n_fields": 1, "R": radii, "start_corner": np.asarray(CORNERS, dtype=np.int32)}


def c2(n_gpus: int = 1, radii_per_gpu: int = 1024) -> Workload:
    cands, axes = _radius_corner(np.linspace(5.0, 12.0, radii_per_gpu * n_gpus), n_gpus)
    return Workload("c2", f"config2: 500x200 m field + 2 obstacles, {4 * radii_per_gpu} candidates/GPU "
                    f"(4 start corners x {radii_per_gpu} radii 5..12 m), h=0.1 m",
                    np.asarray([RECT]), cands, [OBST2], 0.1, "paths", axes)


def c3_fields(F: int = 4096, seed: int = 1234) -> np.ndarray:
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


def c3(n_gpus: int = 1, fields_per_gpu: int = 4096, n_headings: int = 180) -> Workload:
    F = fields_per_gpu * n_gpus
    heads = np.deg2rad(np.arange(float(n_headings)) * (180.0 / n_headings))
    fid = np.repeat(np.arange(F, dtype=np.int32), n_headings)
    cands = {"field_id": fid, "heading": np.tile(heads, F)}
    return Workload("c3", f"config3: {fields_per_gpu} tilted parallelograms/GPU x {n_headings} headings, h=0.1 m, "
                    "summary only, argmin per field", c3_fields(F), cands, None, 0.1, "summary",
                    {"axes": True, "n_fields": F, "heading": heads})


def c5(n_gpus: int = 1, cands_per_gpu: int = 8192) -> Workload:
    n_r = cands_per_gpu // 4 * n_gpus
    # the radii of the 65 536-candidate job are linspace(5, 12, 16384); fewer GPUs take the same range coarser
    cands, axes = _radius_corner(np.linspace(5.0, 12.0, n_r), n_gpus)
    return Workload("c5", f"config5: 2000x1000 m field, h=0.05 m, {cands_per_gpu} candidates/GPU "
                    f"(4 start corners x {cands_per_gpu // 4} radii 5..12 m), summary only",
                    np.asarray([BIG]), cands, None, 0.05, "summary", axes)


WORKLOADS = {"c2": c2, "c3": c3, "c5": c5}
