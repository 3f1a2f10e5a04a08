"""Python face of the integer rasterisation oracle (TEST INFRASTRUCTURE ONLY).

Wraps oracle/raster_oracle.c (built by oracle/Makefile into oracle/_build/libfcpo.so) and states
how the reference's two coverage measures are turned into integer problems (decision D5,
SURVEY.md §8(c)):

  q(x) = rint(x * 1e4)                      1e-4 m fixed point, round-half-even, one FP64 multiply
  corner windows (A10, mlp3:1426-1510)      lattice POINTS q(origin) + (i*H, j*H), i,j in [0,g),
                                            g = int(2R/0.1) (mlp3:1457), H = q(0.1) = 1000
  headland band  (A11, mlp3:1357-1371)      lattice of cell CENTRES anchored at the field bbox
                                            minimum: q(min) + H/2 + (i*H, j*H); band = centre in
                                            field and not in the R-inset; h = 0.1 m default
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import geom, ref_planner as rp

UNIT = 1e4
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "libfcpo.so")
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.fcpo_raster.restype = ctypes.c_int64
        L.fcpo_raster.argtypes = [i64p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        L.fcpo_band.restype = ctypes.c_int
        L.fcpo_band.argtypes = [i64p, ctypes.c_int, i64p, ctypes.c_int, i64p, ctypes.c_int, ctypes.c_int64,
                                ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                ctypes.c_int64, i64p]
        L.fcpo_tour_lengths.restype = None
        L.fcpo_tour_lengths.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_void_p]
        _LIB = L
    return _LIB


def q(x):
    """Snap metres to the 1e-4 m integer lattice (normative: one FP64 multiply, rint)."""
    return np.rint(np.asarray(x, dtype=np.float64) * UNIT).astype(np.int64)


def _i64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def raster_window(path, r, origin, h, nx, ny, bits=None):
    """OR the W/2-buffer cover of ``path`` into a lattice-point window; returns (count, bits)."""
    pts = np.ascontiguousarray(q(path).reshape(-1, 2))
    if bits is None:
        bits = np.zeros((nx * ny + 7) // 8, dtype=np.uint8)
    cnt = lib().fcpo_raster(_i64p(pts), len(pts), int(q(r)), int(q(origin[0])), int(q(origin[1])),
                            int(q(h)), nx, ny, bits.ctypes.data)
    return int(cnt), bits


def corner_coverage(fs: rp.FieldSetup):
    """A10: [(cells_before, cells_after)] for the 4 verification corners + g (mlp3:1512-1578)."""
    R, W = fs.vehicle.min_turn_radius, fs.vehicle.working_width
    g = int(2 * R / rp.GRID_RESOLUTION)
    out = []
    for (cx, cy), ci, arc, rev in rp.verification_corner_paths(fs):
        ox = cx if ci in (0, 3) else cx - 2 * R     # mlp3:1461-1468
        oy = cy if ci in (0, 1) else cy - 2 * R
        before, bits = raster_window(arc, W / 2, (ox, oy), rp.GRID_RESOLUTION, g, g)
        after = before
        if rev is not None and len(rev) > 0:
            after, bits = raster_window(rev, W / 2, (ox, oy), rp.GRID_RESOLUTION, g, g, bits)
        out.append((before, after))
    return out, g


def band_dims(field_vertices, h):
    b = geom.bounds(field_vertices)
    H = int(q(h))
    X0, Y0 = int(q(b[0])), int(q(b[1]))
    nx = -((X0 - int(q(b[2]))) // H)   # ceil div
    ny = -((Y0 - int(q(b[3]))) // H)
    return X0, Y0, H, nx, ny


def band_coverage(fs: rp.FieldSetup, head_path, h=0.1):
    """A11 as integer counts: (band cells, covered band cells)."""
    W = fs.vehicle.working_width
    fv = np.ascontiguousarray(q(np.asarray(fs.field_vertices, dtype=np.float64)))
    main = geom.inset_convex(fs.field_vertices, fs.headland_width)
    if main is not None and abs(geom.signed_area(main)) < 1.0:
        main = None
    X0, Y0, H, nx, ny = band_dims(fs.field_vertices, h)
    assert H % 2 == 0
    pts = np.ascontiguousarray(q(head_path).reshape(-1, 2))
    out = np.zeros(2, dtype=np.int64)
    if main is not None:
        mv = np.ascontiguousarray(q(np.asarray(main, dtype=np.float64)))
        mp, nm = _i64p(mv), 4
    else:
        mp, nm = None, 0
    rc = lib().fcpo_band(_i64p(fv), len(fv), mp, nm, _i64p(pts), len(pts), int(q(W / 2)),
                         X0 + H // 2, Y0 + H // 2, H, nx, ny, _i64p(out))
    assert rc == 0
    return int(out[0]), int(out[1])


def tour_lengths(D, pop):
    """A12: closed-tour lengths, sequential FP64 sums (genetic_algorithm_solver.py:174-181)."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    pop = np.ascontiguousarray(pop, dtype=np.int32)
    out = np.empty(len(pop), dtype=np.float64)
    lib().fcpo_tour_lengths(D.ctypes.data, D.shape[0], pop.ctypes.data, len(pop), out.ctypes.data)
    return out
