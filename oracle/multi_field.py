"""CPU restatement of the reference's multi-field glue (TEST INFRASTRUCTURE ONLY).

"mfp" = /root/reference/multi_field_planner.py.  Pinned by tests/golden/multi_field.npz, made by
executing the unmodified reference module (tests/golden/make_multi_field_golden.py).
"""
from __future__ import annotations

import numpy as np


def distance_matrix(depot, centroids):
    """mfp:263-288: nodes = depot + field centroids, Euclidean distance, zero diagonal."""
    pos = [np.asarray(depot, dtype=np.float64)] + [np.asarray(c, dtype=np.float64) for c in centroids]
    n = len(pos)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = np.linalg.norm(pos[i] - pos[j])
    return D


def best_connection(from_pts, to_pts):
    """mfp:290-320: first strict minimum over from points (outer loop) x to points (inner loop).
    Returns (distance, from_index, to_index)."""
    best, bi, bj = float("inf"), None, None
    for i, fp in enumerate(from_pts):
        for j, tp in enumerate(to_pts):
            d = np.linalg.norm(np.asarray(fp, dtype=np.float64) - np.asarray(tp, dtype=np.float64))
            if d < best:
                best, bi, bj = d, i, j
    return best, bi, bj


def connection_matrix(verts, depot):
    """best_connection for every ordered pair of nodes (node 0 = depot, node f+1 = field f whose
    vertices are both its exit and its entry points, mfp:123-141)."""
    verts = np.asarray(verts, dtype=np.float64)
    nodes = [[np.asarray(depot, dtype=np.float64)]] + [list(v) for v in verts]
    n = len(nodes)
    Cm, fi, ti = np.zeros((n, n)), np.zeros((n, n), dtype=np.int64), np.zeros((n, n), dtype=np.int64)
    for a in range(n):
        for b in range(n):
            Cm[a, b], fi[a, b], ti[a, b] = best_connection(nodes[a], nodes[b])
    return Cm, fi, ti


def entry_directions(verts):
    """mfp:123-141: unit bisector of the incoming and outgoing edge directions at every vertex
    (the incoming direction when the bisector is shorter than 0.1)."""
    out = []
    n = len(verts)
    for i, v in enumerate(verts):
        v_in = np.array(v) - np.array(verts[i - 1])
        v_in = v_in / np.linalg.norm(v_in)
        v_out = np.array(verts[(i + 1) % n]) - np.array(v)
        v_out = v_out / np.linalg.norm(v_out)
        v_avg = (v_in + v_out) / 2
        out.append(v_avg / np.linalg.norm(v_avg) if np.linalg.norm(v_avg) > 0.1 else v_in)
    return np.asarray(out)
