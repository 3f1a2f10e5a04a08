"""CPU model of the coverage kernel's ZONED band evaluation (TEST INFRASTRUCTURE ONLY).

The CUDA kernel (csrc/fcpp_cover.cu: setup_entries, band_zoned) does not rasterise the whole
headland band when the field's straights are axis-aligned chains; it relies on four claims.  This
module restates them in plain Python/numpy integers so that `tests/test_oracle_golden.py` can check
them on the CPU against the brute-force oracle (raster_oracle.c: fcpo_band), independently of the
CUDA code:

  C1  a chain of same-direction axis-aligned segments IS one capsule, and that capsule is exactly
      its two end discs plus the lattice rectangle between them (columns |iH - x| < r on the rows
      ay <= jH <= by for a vertical chain, rows and columns swapped for a horizontal one);
  C2  an entry whose bounding box grown by r lies in the closed R-inset covers no band cell;
  C3  every other ("general") entry lies inside the bounding box of its field quadrant's entries
      (a ZONE), so outside the zones a band cell can only be covered by a rectangle;
  C4  between breakpoints (rows where a rectangle / zone starts or ends, rows of the quads'
      vertices) the per-row count is constant whenever the four window boundaries agree on the
      run's first and last row (and an interval empty at both ends only counts when the run is
      outside the quad's y-range).

`band_zoned(fs, head_path, h)` returns (band cells, covered cells) computed THAT way — the general
entries inside the zones are rasterised with the same exact C predicate as the oracle — or None
when the zones overlap (the kernel then falls back to rasterising the whole band).
"""
from __future__ import annotations

import numpy as np

from . import geom, raster


def _axis_class(p, q):
    dx, dy = int(q[0] - p[0]), int(q[1] - p[1])
    if dy == 0 and dx != 0:
        return 1 if dx > 0 else 2
    if dx == 0 and dy != 0:
        return 3 if dy > 0 else 4
    return 0


def _in_quad_mask(quad, xs, ys):
    """closed containment of the lattice points (xs[i], ys[j]) in a convex CCW integer quad -> [ny, nx]"""
    X, Y = np.meshgrid(xs.astype(np.int64), ys.astype(np.int64))   # |coordinates| < 2^30: products fit int64
    m = np.ones(X.shape, dtype=bool)
    for k in range(4):
        ax, ay = int(quad[k][0]), int(quad[k][1])
        bx, by = int(quad[(k + 1) % 4][0]), int(quad[(k + 1) % 4][1])
        cr = (bx - ax) * (Y - ay) - (by - ay) * (X - ax)
        m &= cr >= 0
    return m


def _row_interval(mask_row):
    idx = np.flatnonzero(mask_row)
    return (int(idx[0]), int(idx[-1])) if len(idx) else (0, -1)


def band_zoned(fs, head_path, h=0.1):
    W = fs.vehicle.working_width
    r = int(raster.q(W / 2))
    X0, Y0, H, nx, ny = raster.band_dims(fs.field_vertices, h)
    Xc0, Yc0 = X0 + H // 2, Y0 + H // 2
    org = np.array([Xc0, Yc0], dtype=np.int64)
    pts = raster.q(np.asarray(head_path, dtype=np.float64)).reshape(-1, 2) - org
    fq = raster.q(np.asarray(fs.field_vertices, dtype=np.float64)) - org
    main = geom.inset_convex(fs.field_vertices, fs.headland_width)
    assert main is not None
    mq = raster.q(np.asarray(main, dtype=np.float64)) - org
    xs, ys = np.arange(nx, dtype=np.int64) * H, np.arange(ny, dtype=np.int64) * H
    in_f, in_m = _in_quad_mask(fq, xs, ys), _in_quad_mask(mq, xs, ys)
    band = in_f & ~in_m

    # ---- C1: chains -> rectangles + end discs; everything else stays a general entry ----
    n = len(pts) - 1
    rects, general = [], []            # rects: (ia, ib, ja, jb); general: (p, q) lattice points
    e = 0
    while e < n:
        c = _axis_class(pts[e], pts[e + 1])
        j = e + 1
        if c:
            while j < n and _axis_class(pts[j], pts[j + 1]) == c:
                j += 1
        p, qq = pts[e], pts[j]
        if c and j > e + 1:
            if c >= 3:
                x, ylo, yhi = int(p[0]), int(min(p[1], qq[1])), int(max(p[1], qq[1]))
                rects.append(((x - r) // H + 1, -((-(x + r)) // H) - 1, -((-ylo) // H), yhi // H))
                general += [((x, ylo), (x, ylo)), ((x, yhi), (x, yhi))]
            else:
                y, xlo, xhi = int(p[1]), int(min(p[0], qq[0])), int(max(p[0], qq[0]))
                rects.append((-((-xlo) // H), xhi // H, (y - r) // H + 1, -((-(y + r)) // H) - 1))
                general += [((xlo, y), (xlo, y)), ((xhi, y), (xhi, y))]
        else:
            general.append(((int(p[0]), int(p[1])), (int(qq[0]), int(qq[1]))))   # one segment or a merged 1-chain
        e = j
    if not rects:
        return None

    # ---- C2: entries inside the closed R-inset are dropped ----
    def inside_main(x, y):
        return all((int(mq[(k + 1) % 4][0]) - int(mq[k][0])) * (y - int(mq[k][1])) -
                   (int(mq[(k + 1) % 4][1]) - int(mq[k][1])) * (x - int(mq[k][0])) >= 0 for k in range(4))
    live = []
    for (a, b) in general:
        x0, x1 = min(a[0], b[0]) - r, max(a[0], b[0]) + r
        y0, y1 = min(a[1], b[1]) - r, max(a[1], b[1]) + r
        if not (inside_main(x0, y0) and inside_main(x1, y0) and inside_main(x1, y1) and inside_main(x0, y1)):
            live.append((a, b))

    # ---- C3: zones = per-quadrant bounding boxes of the live general entries ----
    zones = {}
    midx2, midy2 = (nx - 1) * H, (ny - 1) * H
    for (a, b) in live:
        ylo, yhi = min(a[1], b[1]), max(a[1], b[1])
        jlo, jhi = max((ylo - r) // H + 1, 0), min(-((-(yhi + r)) // H) - 1, ny - 1)
        cl, ch = max((min(a[0], b[0]) - r) // H, 0), min(-((-(max(a[0], b[0]) + r)) // H), nx - 1)
        if jlo > jhi or cl > ch:
            continue
        z = (2 if (a[1] + b[1]) >= midy2 else 0) | (1 if (a[0] + b[0]) >= midx2 else 0)
        zb = zones.get(z)
        zones[z] = (cl, ch, jlo, jhi) if zb is None else (min(zb[0], cl), max(zb[1], ch), min(zb[2], jlo), max(zb[3], jhi))
    zl = list(zones.values())
    for i in range(len(zl)):
        for k in range(i + 1, len(zl)):
            a, b = zl[i], zl[k]
            if a[0] <= b[1] and b[0] <= a[1] and a[2] <= b[3] and b[2] <= a[3]:
                return None                                 # overlapping zones: the kernel falls back
    zmask = np.zeros((ny, nx), dtype=bool)
    for (cl, ch, jlo, jhi) in zl:
        zmask[jlo:jhi + 1, cl:ch + 1] = True

    # inside the zones: rectangles + general entries, exact C predicate per cell
    rmask = np.zeros((ny, nx), dtype=bool)
    for (ia, ib, ja, jb) in rects:
        ia, ib, ja, jb = max(ia, 0), min(ib, nx - 1), max(ja, 0), min(jb, ny - 1)
        if ia <= ib and ja <= jb:
            rmask[ja:jb + 1, ia:ib + 1] = True
    bits = np.zeros((nx * ny + 7) // 8, dtype=np.uint8)
    L = raster.lib()
    for (a, b) in live:
        seg = np.ascontiguousarray(np.array([a, b], dtype=np.int64) + org)
        L.fcpo_raster(raster._i64p(seg), 2, r, Xc0, Yc0, H, nx, ny, bits.ctypes.data)
    gmask = np.unpackbits(bits, bitorder="little")[:nx * ny].reshape(ny, nx).astype(bool)
    assert not (gmask & band & ~zmask).any(), "C3 violated: a general entry covers a band cell outside the zones"
    cov_zones = int(((gmask | rmask) & band & zmask).sum())

    # ---- C4: outside the zones row by row vs run by run ----
    rows_cov = ((rmask & band & ~zmask).sum(axis=1)).astype(np.int64)
    rows_tot = band.sum(axis=1).astype(np.int64)
    raw = [(_row_interval(in_f[j]), _row_interval(in_m[j])) for j in range(ny)]
    bps = {0, ny}
    for quad in (fq, mq):
        for k in range(4):
            bps.add(min(max(-((-int(quad[k][1])) // H), 0), ny))
    for (ia, ib, ja, jb) in rects:
        bps.update((min(max(ja, 0), ny), min(max(jb + 1, 0), ny)))
    for (cl, ch, jlo, jhi) in zl:
        bps.update((jlo, jhi + 1))
    bps = sorted(bps)

    def outside(quad, lo, hi):
        ys_ = [int(v[1]) for v in quad]
        return hi * H < min(ys_) or lo * H > max(ys_)
    run_cov = run_tot = 0
    for lo, nxt in zip(bps[:-1], bps[1:]):
        hi = nxt - 1
        if lo > hi:
            continue
        same = raw[lo] == raw[hi]
        if hi > lo and raw[lo][0][0] > raw[lo][0][1]:
            same = same and outside(fq, lo, hi)
        if hi > lo and raw[lo][1][0] > raw[lo][1][1]:
            same = same and outside(mq, lo, hi)
        if same:       # counted once, multiplied
            run_cov += int(rows_cov[lo]) * (hi - lo + 1)
            run_tot += int(rows_tot[lo]) * (hi - lo + 1)
        else:          # slanted boundaries: row by row
            run_cov += int(rows_cov[lo:hi + 1].sum())
            run_tot += int(rows_tot[lo:hi + 1].sum())
    assert run_cov == int(rows_cov.sum()) and run_tot == int(rows_tot.sum()), "C4 violated: a run is not constant"
    return run_tot, cov_zones + run_cov
