"""CPU restatement of the reference's single-field two-layer planner hot path
(TEST INFRASTRUCTURE ONLY — never imported by the product package).

Follows ``/root/reference/multi_layer_planner_v3.py`` ("mlp3") function by function; every
function cites the lines it restates.  All quirks of SURVEY.md App. A are reproduced, not fixed.
Shapely/GEOS calls are replaced by oracle/geom.py under decisions D1/D2 (parity with real GEOS
is UNPINNED: Shapely cannot be installed here; see oracle/README.md).

Pinning: tests/test_oracle_golden.py checks this file against fixtures produced by executing
the unmodified reference through oracle/shapely_stub.py (tests/golden/make_golden.py), and
against the prose known-answers of the reference's README / changelog.

Two batch extensions that the reference does not have are restated here so that the CUDA
batch path has something to be compared with:
  * ``heading``      replaces the return value of ``_calculate_rotation_angle`` (mlp3:244-263);
  * ``start_corner`` c replaces ``start_point``: ``start_corner_index=c`` (mlp3:397-399),
                     ``reverse_order = c in (2,3)``, ``start_from_right = c in (1,2)`` (mlp3:650-658).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import geom

Pt = Tuple[float, float]

# hard-coded constants of the reference (SURVEY.md §5 "Config / flags")
UTURN_POINTS = 20          # mlp3:807
CORNER_ARC_POINTS = 15     # mlp3:1046, :1589
STRAIGHT_POINTS = 20       # mlp3:990
REVERSE_SPACING = 0.5      # mlp3:1214
REVERSE_MIN_POINTS = 10    # mlp3:1214
REVERSE_CAP_FACTOR = 3.0   # mlp3:1279
REVERSE_SPEED_KMH = 2.5    # mlp3:1080
GRID_RESOLUTION = 0.1      # mlp3:1452
KAPPA_EPS = 1e-6           # mlp3:496
ZERO_LEN = 1e-6            # mlp3:526, :560, :576
ROT_THRESHOLD = 0.01       # mlp3:686, :709
MIN_SPEED_MS = 0.1         # mlp3:1308
APPROACH_POINTS = 50       # mlp3:1317
GAP_AREA_GATE = 0.1        # mlp3:1070
GEOFENCE_EPS = 1e-9        # D3


@dataclass
class VehicleParams:
    """mlp3:29-39."""
    working_width: float = 3.2
    min_turn_radius: float = 8.0
    max_work_speed_kmh: float = 9.0
    max_headland_speed_kmh: float = 15.0
    headland_turn_speed_kmh: float = 4.0
    max_lateral_accel: float = 2.0
    max_longitudinal_accel: float = 1.5
    safety_factor: float = 0.85


class PlanError(ValueError):
    """Per-candidate geometric failure (mlp3:597-598 ValueError, or the shape error that
    ``np.vstack`` raises after a skipped loop, mlp3:967-969 + :939)."""


# ------------------------------------------------------------------------------------------
# A2  field setup (mlp3:109-385)
# ------------------------------------------------------------------------------------------
@dataclass
class FieldSetup:
    vehicle: VehicleParams
    field_vertices: List[Pt]
    field_length: float
    field_width: float
    field_shape: str
    corner_angles: List[float]
    headland_width: float
    main_work_pattern: str
    obstacles: List[List[Pt]]
    start_point: Optional[Pt]
    end_point: Optional[Pt]
    # build-defined opt-in: "arc" = the reference's sampled circular arcs,
    # "clothoid" = clothoid-arc-clothoid turns (oracle/clothoid.py); parity unpinned for the latter
    turn_model: str = "arc"
    clothoid_share: float = 0.5


def corner_angle(vertices: Sequence[Pt], i: int) -> float:
    """mlp3:165-192."""
    n = len(vertices)
    prev = vertices[(i - 1) % n]
    curr = vertices[i]
    nxt = vertices[(i + 1) % n]
    v1 = np.array([prev[0] - curr[0], prev[1] - curr[1]])
    v2 = np.array([nxt[0] - curr[0], nxt[1] - curr[1]])
    cos_angle = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
    return float(np.degrees(np.arccos(np.clip(cos_angle, -1.0, 1.0))))


def is_parallelogram(vertices: Sequence[Pt]) -> bool:
    """mlp3:194-222."""
    if len(vertices) != 4:
        return False
    edges = []
    for i in range(4):
        a = vertices[i]
        b = vertices[(i + 1) % 4]
        edges.append(np.array([b[0] - a[0], b[1] - a[1]]))

    def par(u, v, tol=0.01):
        return abs(u[0] * v[1] - u[1] * v[0]) < tol * (np.linalg.norm(u) * np.linalg.norm(v))

    return bool(par(edges[0], edges[2]) and par(edges[1], edges[3]))


def validate_point(p: Optional[Pt], field_length: float, field_width: float) -> Optional[Pt]:
    """mlp3:322-343 — closed bbox-EXTENT test anchored at the origin (Q11)."""
    if p is None:
        return None
    x, y = p
    if not (0 <= x <= field_length and 0 <= y <= field_width):
        return None
    return (x, y)


def setup_field(vehicle: VehicleParams, field_length=None, field_width=None, field_vertices=None,
                obstacles=None, start_point=None, end_point=None, turn_model="arc",
                clothoid_share=0.5) -> FieldSetup:
    """mlp3:63-135 (constructor + _process_field_input)."""
    if field_vertices is not None:
        verts = [tuple(v) for v in field_vertices]
        b = geom.bounds(verts)
        fl = b[2] - b[0]
        fw = b[3] - b[1]
    elif field_length is not None and field_width is not None:
        fl, fw = field_length, field_width
        verts = [(0, 0), (field_length, 0), (field_length, field_width), (0, field_width)]
    else:
        raise ValueError("必须提供 field_vertices 或 (field_length, field_width)")
    n = len(verts)
    angles = [corner_angle(verts, i) for i in range(n)]
    if n != 4:
        shape = "other"
    elif all(abs(a - 90) < 1.0 for a in angles):
        shape = "rectangle"
    elif is_parallelogram(verts):
        shape = "parallelogram"
    else:
        shape = "other"
    aspect = fl / fw  # mlp3:312-320 (label only)
    pattern = "Ω型跨行" if aspect < 1.5 else "U型往复"
    return FieldSetup(vehicle, verts, fl, fw, shape, angles, vehicle.min_turn_radius, pattern,
                      [list(map(tuple, o)) for o in (obstacles or [])],
                      validate_point(start_point, fl, fw), validate_point(end_point, fl, fw),
                      turn_model, clothoid_share)


def select_best_start_corner(fs: FieldSetup, parking: Pt) -> int:
    """mlp3:345-385 — candidates at (hw/2, hw/2)…, first minimum wins (Q11)."""
    w = fs.headland_width
    cands = [(w / 2, w / 2), (fs.field_length - w / 2, w / 2),
             (fs.field_length - w / 2, fs.field_width - w / 2), (w / 2, fs.field_width - w / 2)]
    d = [float(np.sqrt((x - parking[0]) ** 2 + (y - parking[1]) ** 2)) for x, y in cands]
    return int(min(range(4), key=lambda i: d[i]))


def should_apply_reverse_filling(fs: FieldSetup, corner_index: int) -> bool:
    """mlp3:224-242."""
    return fs.corner_angles[corner_index] >= 60


# ------------------------------------------------------------------------------------------
# A3/A4  layer 1 (mlp3:591-830)
# ------------------------------------------------------------------------------------------
def work_area_centroid(main_boundary: List[Pt], obstacles: List[List[Pt]], W: float) -> Pt:
    """Centroid of ``main_boundary.difference(unary_union(buffer(obs, W/2)))`` (mlp3:601-609)
    under D2, assuming the buffered obstacles are disjoint and inside the work area (Q2)."""
    a = abs(geom.signed_area(main_boundary))
    c = geom.centroid(main_boundary)
    mx, my = a * c[0], a * c[1]
    for o in obstacles:
        o = list(o)
        if geom.signed_area(o) < 0:
            o = o[::-1]
        ha, hx, hy = geom.round_buffer_moments(o, W / 2)
        a -= ha
        mx -= hx
        my -= hy
    return (mx / a, my / a)


def uturn_tables():
    """cos/sin of linspace(0, pi, 20) (mlp3:807-808) and linspace(0, pi/2, 15) (mlp3:1046-1047).
    The Python host of the product computes the same tables with the same numpy calls and hands
    them to the CUDA library, so device arcs are bit-identical to numpy's on the same machine."""
    a20 = np.linspace(0, np.pi, UTURN_POINTS)
    a15 = np.linspace(0, np.pi / 2, CORNER_ARC_POINTS)
    return np.cos(a20), np.sin(a20), np.cos(a15), np.sin(a15)


def omega_skip(R: float, W: float) -> int:
    """Rows skipped by a connecting turn of the Ω pattern: the smallest s with s W >= 2 R (a plain half circle of
    radius s W / 2 >= R joins the two rows), at least 1."""
    return max(1, int(np.ceil(2.0 * R / W - 1e-9)))


def omega_order(P: int, s: int) -> List[int]:
    """Visit order of the P rows in the Ω (skip-row) pattern: blocks of 2 s rows, inside a block the rows of its
    lower half and of its upper half alternate (0, h, 1, h + 1, ...; h = ceil(m / 2) for a block of m rows), so
    consecutive rows are h or h - 1 apart; the blocks follow each other."""
    order, base = [], 0
    while base < P:
        m = min(2 * s, P - base)
        h = (m + 1) // 2
        order.extend(base + k // 2 + (h if k % 2 else 0) for k in range(m))
        base += m
    return order


def omega_turn_local(d_abs: float, R: float, n: int = 20):
    """The connecting turn of the Ω pattern between two rows |d| apart, in its own frame (u = outward along the swath,
    v = towards the next row; starts at (0, 0) heading +u, ends at (0, |d|) heading -u), n samples equally spaced in
    arc length, first and last ON the two swath ends.
      |d| >= 2 R: a half circle of radius |d| / 2.
      |d| <  2 R: the Ω (bulb) turn — three arcs of radius R: right by alpha, left by pi + 2 alpha, right by alpha,
                  cos(alpha) = (R + |d| / 2) / (2 R)."""
    if d_abs >= 2.0 * R:
        rho = d_abs / 2.0
        a = np.linspace(0, np.pi, n)
        return rho * np.sin(a), rho - rho * np.cos(a)
    xc = np.sqrt(4.0 * R * R - (R + d_abs / 2.0) ** 2)
    alpha = np.arctan2(xc, R + d_abs / 2.0)
    total = np.pi + 4.0 * alpha
    u = np.empty(n)
    v_ = np.empty(n)
    c2 = (2.0 * R * np.sin(alpha), -R + 2.0 * R * np.cos(alpha))
    for i in range(n):
        phi = total if i == n - 1 else i * (total / (n - 1))
        if phi <= alpha:
            u[i], v_[i] = R * np.sin(phi), -R + R * np.cos(phi)
        elif phi <= np.pi + 3.0 * alpha:
            h = -alpha + (phi - alpha)
            u[i], v_[i] = c2[0] + R * np.sin(h), c2[1] - R * np.cos(h)
        else:
            h = np.pi + alpha - (phi - (np.pi + 3.0 * alpha))
            u[i], v_[i] = -R * np.sin(h), d_abs + R + R * np.cos(h)
    return u, v_


def omega_pattern_in_rotated_space(bnds, v: VehicleParams, reverse_order: bool, start_from_right: bool):
    """Ω-type skip-row main work (BUILD-DEFINED, parity unpinned: the reference only returns the LABEL 'Ω型跨行',
    mlp3:312-320, and always generates the U pattern).  Same rows, same swath ends (mlp3:736-751), same sample counts
    (2 + 20 per pass, mlp3:761-830) as the U pattern; the rows are visited in ``omega_order`` and every turn really
    connects the two swath ends (``omega_turn_local``), bulging outwards beyond the swath end."""
    min_x, min_y, max_x, max_y = bnds
    R, W = v.min_turn_radius, v.working_width
    line_start_x, line_end_x = min_x + R, max_x - R
    P = int((max_y - min_y) / W) + 1
    rows = omega_order(P, omega_skip(R, W))
    if reverse_order:
        rows = [P - 1 - r for r in rows]
    segs, speeds = [], []
    for idx, i in enumerate(rows):
        y = min_y + i * W
        go_left = (idx % 2 == 0) if start_from_right else (idx % 2 == 1)
        line = np.array([[line_end_x, y], [line_start_x, y]]) if go_left else np.array([[line_start_x, y], [line_end_x, y]])
        segs.append(line)
        speeds.extend([v.max_work_speed_kmh] * 2)
        if idx < P - 1:
            d = (min_y + rows[idx + 1] * W) - y
            u, vv = omega_turn_local(abs(d), R, UTURN_POINTS)
            dir_x = -1.0 if go_left else 1.0
            sg = 1.0 if d >= 0 else -1.0
            segs.append(np.column_stack([line[-1][0] + dir_x * u, y + sg * vv]))
            speeds.extend([v.headland_turn_speed_kmh] * UTURN_POINTS)
    return np.vstack(segs), np.array(speeds, dtype=np.float64), P


def u_pattern_in_rotated_space(bnds, v: VehicleParams, reverse_order: bool, start_from_right: bool,
                               turn_model: str = "arc", clothoid_share: float = 0.5):
    """mlp3:720-830 (swath layout + 20-pt half-circle 'turns', Q3/Q4).  With
    turn_model='clothoid' the 20 turn samples are a clothoid-arc-clothoid U-turn that leaves the
    swath end TANGENTIALLY towards the next swath (build-defined, A16)."""
    min_x, min_y, max_x, max_y = bnds
    R = v.min_turn_radius
    line_start_x = min_x + R
    line_end_x = max_x - R
    num_passes = int((max_y - min_y) / v.working_width) + 1
    order = list(range(num_passes - 1, -1, -1)) if reverse_order else list(range(num_passes))
    angles = np.linspace(0, np.pi, UTURN_POINTS)
    segs = []
    speeds: List[float] = []
    for idx, i in enumerate(order):
        y = min_y + i * v.working_width
        go_left = (idx % 2 == 0) if start_from_right else (idx % 2 == 1)
        if go_left:
            line = np.array([[line_end_x, y], [line_start_x, y]])
        else:
            line = np.array([[line_start_x, y], [line_end_x, y]])
        segs.append(line)
        speeds.extend([v.max_work_speed_kmh] * 2)
        if idx < num_passes - 1 and turn_model == "clothoid":
            from .clothoid import cac_unit
            xi, eta = cac_unit(np.pi, UTURN_POINTS, clothoid_share)
            dir_x = -1.0 if go_left else 1.0
            dir_y = -1.0 if reverse_order else 1.0
            arc_x = line[-1][0] + dir_x * (R * xi)
            arc_y = y + dir_y * (R * eta)
            segs.append(np.column_stack([arc_x, arc_y]))
            speeds.extend([v.headland_turn_speed_kmh] * UTURN_POINTS)
        elif idx < num_passes - 1:
            turn_right = not go_left
            if turn_right:
                arc_x = max_x - R * np.cos(angles)
                arc_y = y + R * np.sin(angles)
            else:
                arc_x = min_x + R * np.cos(angles)
                arc_y = y + R * np.sin(angles)
            segs.append(np.column_stack([arc_x, arc_y]))
            speeds.extend([v.headland_turn_speed_kmh] * UTURN_POINTS)
    return np.vstack(segs), np.array(speeds, dtype=np.float64), num_passes


def plan_main_work(fs: FieldSetup, heading: Optional[float] = None,
                   start_corner: Optional[int] = None):
    """mlp3:591-629 + :670-718.  Returns (path, speeds, info)."""
    v = fs.vehicle
    main_boundary = geom.inset_convex(fs.field_vertices, fs.headland_width)
    if main_boundary is None or abs(geom.signed_area(main_boundary)) < 1.0:
        raise PlanError(f"田头宽度{fs.headland_width}m过大，无法定义主作业区域")
    if heading is None:
        v0, v1 = fs.field_vertices[0], fs.field_vertices[1]
        angle = float(np.arctan2(v1[1] - v0[1], v1[0] - v0[0]))  # mlp3:244-263
    else:
        angle = float(heading)
    rotated = abs(angle) > ROT_THRESHOLD
    center = (0.0, 0.0)
    sp = fs.start_point
    if rotated:
        if fs.obstacles:
            center = work_area_centroid(main_boundary, fs.obstacles, v.working_width)
        else:
            center = geom.centroid(main_boundary)
        cn, sn = float(np.cos(-angle)), float(np.sin(-angle))
        rverts = [geom.rotate_point(p, cn, sn, center) for p in main_boundary]
        if sp is not None:
            sp = geom.rotate_point(sp, cn, sn, center)
    else:
        rverts = main_boundary
    bnds = geom.bounds(rverts)
    # mlp3:631-668
    reverse_order = False
    start_from_right = False
    if start_corner is not None:
        reverse_order = start_corner in (2, 3)
        start_from_right = start_corner in (1, 2)
    elif sp is not None:
        if sp[1] > (bnds[1] + bnds[3]) / 2:
            reverse_order = True
        if sp[0] > (bnds[0] + bnds[2]) / 2:
            start_from_right = True
    if fs.turn_model == "omega":
        path, speeds, P = omega_pattern_in_rotated_space(bnds, v, reverse_order, start_from_right)
    else:
        path, speeds, P = u_pattern_in_rotated_space(bnds, v, reverse_order, start_from_right, fs.turn_model,
                                                     fs.clothoid_share)
    if rotated:
        cp, sp_ = float(np.cos(angle)), float(np.sin(angle))
        x = path[:, 0] - center[0]
        y = path[:, 1] - center[1]
        xn = x * cp - y * sp_
        yn = x * sp_ + y * cp
        path = np.column_stack([xn + center[0], yn + center[1]])
    info = dict(P=P, angle=angle, rotated=rotated, center=center, bounds=bnds,
                reverse_order=reverse_order, start_from_right=start_from_right,
                main_boundary=main_boundary)
    return path, speeds, info


# ------------------------------------------------------------------------------------------
# A5/A6  layer 2 (mlp3:860-1288, :1580-1608)
# ------------------------------------------------------------------------------------------
def corner_turn_arc(corner: Pt, corner_index: int, R: float, num_points: int = CORNER_ARC_POINTS,
                    turn_model: str = "arc", clothoid_share: float = 0.5):
    """mlp3:1580-1608 (also :1046-1062 and, with 30 points, :1124-1140).  turn_model='clothoid':
    same start pose and turning sense, clothoid-arc-clothoid quarter turn (A16)."""
    x, y = corner
    if turn_model == "clothoid":
        from .clothoid import cac_unit
        xi, eta = cac_unit(np.pi / 2, num_points, clothoid_share)
        if corner_index == 0:
            return np.column_stack([x + R * eta, y + R * xi])
        if corner_index == 1:
            return np.column_stack([x - R * xi, y + R * eta])
        if corner_index == 2:
            return np.column_stack([x - R * eta, y - R * xi])
        return np.column_stack([x + R * xi, y - R * eta])
    a = np.linspace(0, np.pi / 2, num_points)
    if corner_index == 0:
        ax = x + R * (1 - np.cos(a)); ay = y + R * np.sin(a)
    elif corner_index == 1:
        ax = x - R * np.sin(a); ay = y + R * (1 - np.cos(a))
    elif corner_index == 2:
        ax = x - R * (1 - np.cos(a)); ay = y - R * np.sin(a)
    else:
        ax = x + R * np.sin(a); ay = y - R * (1 - np.cos(a))
    return np.column_stack([ax, ay])


def corner_gap_area_lower_bound(R: float, W: float) -> float:
    """Lower bound of ``square.difference(buffer(arc, W/2)).area`` (mlp3:1086-1152): the 2R×2R
    square minus the whole area of the round buffer of a quarter arc of length <= pi/2·R."""
    r = W / 2
    return 4 * R * R - (0.5 * math.pi * R * W + math.pi * r * r)


def corner_gap_area(corner: Pt, corner_index: int, R: float, W: float, h: float = 0.01) -> float:
    """Numeric value of the gap area (D2), only needed when the lower bound is not decisive."""
    x, y = corner
    ox = x if corner_index in (0, 3) else x - 2 * R
    oy = y if corner_index in (0, 1) else y - 2 * R
    arc = corner_turn_arc(corner, corner_index, R, 30)
    n = int(round(2 * R / h))
    xs = ox + (np.arange(n) + 0.5) * h
    X, Y = np.meshgrid(xs, oy + (np.arange(n) + 0.5) * h)
    cov = np.zeros(X.shape, dtype=bool)
    for k in range(len(arc) - 1):
        cov |= geom.dist2_point_segment(X, Y, arc[k, 0], arc[k, 1], arc[k + 1, 0], arc[k + 1, 1]) < (W / 2) ** 2
    return float(np.count_nonzero(~cov)) * h * h


def gap_gate(corner: Pt, corner_index: int, R: float, W: float) -> bool:
    """``gap is not None and gap.area > 0.1`` (mlp3:1070, :1551)."""
    if corner_gap_area_lower_bound(R, W) > GAP_AREA_GATE:
        return True
    return corner_gap_area(corner, corner_index, R, W) > GAP_AREA_GATE


def distance_to_boundary(fs: FieldSetup, start: np.ndarray, direction: np.ndarray) -> float:
    """mlp3:1220-1288 — ray to the bbox lines x=0, x=field_length, y=0, y=field_width (Q10)."""
    x, y = float(start[0]), float(start[1])
    dx, dy = float(direction[0]), float(direction[1])
    ds = []
    if abs(dx) > 1e-6:
        t = (0 - x) / dx
        if t > 0:
            ds.append(t)
        t = (fs.field_length - x) / dx
        if t > 0:
            ds.append(t)
    if abs(dy) > 1e-6:
        t = (0 - y) / dy
        if t > 0:
            ds.append(t)
        t = (fs.field_width - y) / dy
        if t > 0:
            ds.append(t)
    if not ds:
        return 2.0 * fs.vehicle.min_turn_radius
    return min(min(ds), REVERSE_CAP_FACTOR * fs.vehicle.min_turn_radius)


def optimal_reverse_path(fs: FieldSetup, turn_end: np.ndarray, turn_second_last: np.ndarray):
    """mlp3:1154-1218 (chord direction, Q10).  The centroid fallback (:1195-1206) is reachable
    only when the last two arc samples coincide (R < ~1e-5 m) and is restated as the fixed
    direction (-1, 0)."""
    tang = turn_end - turn_second_last
    nrm = float(np.sqrt(tang[0] * tang[0] + tang[1] * tang[1]))
    if nrm > 1e-6:
        rdir = -tang / nrm
    else:
        rdir = np.array([-1.0, 0.0])
    length = distance_to_boundary(fs, turn_end, rdir)
    n = max(REVERSE_MIN_POINTS, int(length / REVERSE_SPACING))
    t = np.linspace(0, length, n)
    return turn_end + t[:, np.newaxis] * rdir, length


def headland_loop(fs: FieldSetup, offset: float, loop_index: int, start_corner_index: int):
    """mlp3:943-1011."""
    v = fs.vehicle
    R, W = v.min_turn_radius, v.working_width
    corners = geom.inset_convex(fs.field_vertices, offset)
    if corners is None or abs(geom.signed_area(corners)) < 1.0:
        raise PlanError("headland loop skipped (mlp3:967-969) -> vstack shape error (mlp3:939)")
    segs = [np.array([corners[start_corner_index]])]
    speeds: List[float] = [v.max_headland_speed_kmh]
    for i in range(4):
        ci = (start_corner_index + i) % 4
        ni = (start_corner_index + i + 1) % 4
        cur, nxt = corners[ci], corners[ni]
        sx = np.linspace(cur[0], nxt[0], STRAIGHT_POINTS)
        sy = np.linspace(cur[1], nxt[1], STRAIGHT_POINTS)
        segs.append(np.column_stack([sx, sy]))
        speeds.extend([v.max_headland_speed_kmh] * STRAIGHT_POINTS)
        if i < 3:
            arc = corner_turn_arc(nxt, ni, R, CORNER_ARC_POINTS, fs.turn_model, fs.clothoid_share)
            tsp = [v.headland_turn_speed_kmh] * CORNER_ARC_POINTS
            if loop_index == 0 and should_apply_reverse_filling(fs, ni) and gap_gate(nxt, ni, R, W):
                rev, _ = optimal_reverse_path(fs, arc[-1], arc[-2])
                arc = np.vstack([arc, rev])
                tsp.extend([REVERSE_SPEED_KMH] * len(rev))
            segs.append(arc)
            speeds.extend(tsp)
    return np.vstack(segs), speeds


def plan_headland(fs: FieldSetup, start_corner_index: int = 0):
    """mlp3:860-941 (path + speeds; coverage_rate is computed by oracle/raster.py)."""
    v = fs.vehicle
    K = math.ceil(fs.headland_width / v.working_width)
    paths, speeds = [], []
    for k in range(K):
        offset = v.working_width / 2 + k * v.working_width
        p, s = headland_loop(fs, offset, k, start_corner_index)
        paths.append(p)
        speeds.extend(s)
    return np.vstack(paths), np.array(speeds, dtype=np.float64), K


# ------------------------------------------------------------------------------------------
# A7  speed planning (mlp3:467-589)
# ------------------------------------------------------------------------------------------
def curvatures(path: np.ndarray) -> np.ndarray:
    """mlp3:513-536 for i = 1..N-2 (vectorised; element-wise identical operations)."""
    d = np.diff(path, axis=0)
    dx1, dy1 = d[:-1, 0], d[:-1, 1]
    dx2, dy2 = d[1:, 0], d[1:, 1]
    ds1 = np.sqrt(dx1 * dx1 + dy1 * dy1)
    ds2 = np.sqrt(dx2 * dx2 + dy2 * dy2)
    th1 = np.arctan2(dy1, dx1)
    th2 = np.arctan2(dy2, dx2)
    dth = th2 - th1
    dth = np.arctan2(np.sin(dth), np.cos(dth))
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.abs(2 * dth / (ds1 + ds2))
    return np.where((ds1 < ZERO_LEN) | (ds2 < ZERO_LEN), 0.0, k)


def speed_plan(path: np.ndarray, speeds: np.ndarray, v: VehicleParams) -> np.ndarray:
    """mlp3:467-589, km/h in and out, literal three passes (Q6, Q7)."""
    n = len(path)
    if n < 3:
        return speeds
    out = np.array(speeds, dtype=np.float64)
    kap = curvatures(path)
    has = kap > KAPPA_EPS
    with np.errstate(divide="ignore"):
        vmax = np.sqrt(v.max_lateral_accel / np.where(has, kap, 1.0)) * v.safety_factor * 3.6
    idx = np.nonzero(has & (out[1:-1] > vmax))[0] + 1
    out[idx] = vmax[idx - 1]
    d = np.diff(path, axis=0)
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    a = v.max_longitudinal_accel
    s = out.tolist()
    dl = dist.tolist()
    for i in range(1, n):  # mlp3:558-571
        ds = dl[i - 1]
        if ds < ZERO_LEN:
            continue
        v1 = s[i - 1] / 3.6
        vm = math.sqrt(v1 * v1 + 2 * a * ds) * 3.6
        if s[i] > vm:
            s[i] = vm
    for i in range(n - 2, -1, -1):  # mlp3:574-587
        ds = dl[i]
        if ds < ZERO_LEN:
            continue
        v2 = s[i + 1] / 3.6
        vm = math.sqrt(v2 * v2 + 2 * a * ds) * 3.6
        if s[i] > vm:
            s[i] = vm
    return np.array(s, dtype=np.float64)


# ------------------------------------------------------------------------------------------
# A13  metrics (mlp3:1290-1311)
# ------------------------------------------------------------------------------------------
def path_length(path: np.ndarray) -> float:
    if len(path) < 2:
        return 0.0
    d = np.diff(path, axis=0)
    return float(np.sum(np.sqrt(np.sum(d ** 2, axis=1))))


def work_time(path: np.ndarray, speeds: np.ndarray) -> float:
    if len(path) < 2 or len(speeds) == 0:
        return 0.0
    d = np.diff(path, axis=0)
    dist = np.sqrt(np.sum(d ** 2, axis=1))
    avg = np.maximum((speeds[:-1] + speeds[1:]) / 2 / 3.6, MIN_SPEED_MS)
    return float(np.sum(dist / avg))


# ------------------------------------------------------------------------------------------
# A8/A9  validation
# ------------------------------------------------------------------------------------------
def verify_curvature_constraints(path: np.ndarray, speeds: np.ndarray, v: VehicleParams) -> Dict:
    """mlp3:1373-1424."""
    if len(path) < 3:
        return {'max_curvature': 0, 'violations': 0, 'pass': True}
    kap = curvatures(path)
    a_lat = (speeds[1:-1] / 3.6) ** 2 * kap
    viol = int(np.sum(a_lat > v.max_lateral_accel))
    rate = viol / len(a_lat) * 100 if len(a_lat) else 0
    max_jump = float(np.max(np.abs(np.diff(kap)))) if len(kap) > 1 else 0
    return {'max_curvature': float(np.max(kap)), 'max_lateral_accel': float(np.max(a_lat)),
            'max_allowed_accel': v.max_lateral_accel, 'accel_violations': viol,
            'accel_violation_rate': rate, 'max_jump': max_jump, 'pass': rate < 5}


def boundary_violations(path: np.ndarray, field_vertices: Sequence[Pt]) -> int:
    """D3 geofence: #points outside the convex CCW field by more than 1e-9 m.
    Test per edge: cross(e, p - v) < -eps * |e|  (normative operation order)."""
    px, py = path[:, 0], path[:, 1]
    out = np.zeros(len(path), dtype=bool)
    n = len(field_vertices)
    for k in range(n):
        ax, ay = field_vertices[k]
        bx, by = field_vertices[(k + 1) % n]
        ex, ey = bx - ax, by - ay
        ln = math.sqrt(ex * ex + ey * ey)
        cr = ex * (py - ay) - ey * (px - ax)
        out |= cr < -GEOFENCE_EPS * ln
    return int(np.count_nonzero(out))


def obstacle_violations(path: np.ndarray, obstacles: Sequence[Sequence[Pt]], W: float) -> int:
    """D3: #points strictly inside any obstacle buffered by W/2 under D2:
    inside the polygon (even-odd crossing) OR dist² to an edge < (W/2)²."""
    if not obstacles:
        return 0
    px, py = path[:, 0], path[:, 1]
    r = W / 2
    r2 = r * r
    hit = np.zeros(len(path), dtype=bool)
    for o in obstacles:
        o = np.asarray(o, dtype=np.float64)
        hit |= geom.point_in_polygon_crossing(px, py, o)
        for k in range(len(o)):
            a, b = o[k], o[(k + 1) % len(o)]
            hit |= geom.dist2_point_segment(px, py, a[0], a[1], b[0], b[1]) < r2
    return int(np.count_nonzero(hit))


# ------------------------------------------------------------------------------------------
# A10 inputs: the four verification corners (mlp3:1531-1554)
# ------------------------------------------------------------------------------------------
def verification_corner_paths(fs: FieldSetup):
    """For c = 0..3: (corner, 15-pt arc, reverse path or None) exactly as
    verify_all_corners_coverage builds them (Q15) — NOT the actual headland path."""
    hw = fs.headland_width
    R, W = fs.vehicle.min_turn_radius, fs.vehicle.working_width
    data = [(hw, hw, 0), (fs.field_length - hw, hw, 1),
            (fs.field_length - hw, fs.field_width - hw, 2), (hw, fs.field_width - hw, 3)]
    out = []
    for cx, cy, ci in data:
        arc = corner_turn_arc((cx, cy), ci, R, CORNER_ARC_POINTS, fs.turn_model, fs.clothoid_share)
        rev = None
        if gap_gate((cx, cy), ci, R, W):
            rev, _ = optimal_reverse_path(fs, arc[-1], arc[-2])
        out.append(((cx, cy), ci, arc, rev))
    return out


# ------------------------------------------------------------------------------------------
# A15  orchestration (mlp3:387-465)
# ------------------------------------------------------------------------------------------
def plan_complete_coverage(fs: FieldSetup, heading: Optional[float] = None,
                           start_corner: Optional[int] = None) -> Dict:
    v = fs.vehicle
    sci = 0
    if start_corner is not None:
        sci = int(start_corner)
    elif fs.start_point:
        sci = select_best_start_corner(fs, fs.start_point)
    mp, ms, info = plan_main_work(fs, heading, start_corner)
    hp, hs, K = plan_headland(fs, sci)
    main_len, head_len = path_length(mp), path_length(hp)
    t_main_pre, t_head_pre = work_time(mp, ms), work_time(hp, hs)
    all_path = np.vstack([mp, hp])
    adj = speed_plan(all_path, np.concatenate([ms, hs]), v)
    nm = len(mp)
    ms2, hs2 = adj[:nm], adj[nm:]
    res = {
        'main_work': {'path': mp, 'speeds': ms2, 'pattern': fs.main_work_pattern,
                      'stats': {'path_length_km': main_len / 1000,
                                'time_hours': work_time(mp, ms2) / 3600,
                                'avg_speed_kmh': (main_len / 1000) / (t_main_pre / 3600) if t_main_pre > 0 else 0}},
        'headland': {'path': hp, 'speeds': hs2,
                     'stats': {'path_length_km': head_len / 1000,
                               'time_hours': work_time(hp, hs2) / 3600,
                               'avg_speed_kmh': (head_len / 1000) / (t_head_pre / 3600) if t_head_pre > 0 else 0}},
        'approach_path': None, 'departure_path': None,
        'version': 'V3.5.1',
        'features': ['真正两层', '切线倒车', '网格验证', '强制降速', '智能起点'],
        '_info': dict(info, K=K, start_corner_index=sci, speeds_pre=np.concatenate([ms, hs])),
    }
    if fs.start_point:  # mlp3:437-441, :1313-1333 (targets headland.path[0], Q12)
        e = hp[0]
        res['approach_path'] = np.column_stack([np.linspace(fs.start_point[0], e[0], APPROACH_POINTS),
                                                np.linspace(fs.start_point[1], e[1], APPROACH_POINTS)])
    if fs.end_point:    # mlp3:443-447, :1335-1355
        s = hp[-1]
        res['departure_path'] = np.column_stack([np.linspace(s[0], fs.end_point[0], APPROACH_POINTS),
                                                 np.linspace(s[1], fs.end_point[1], APPROACH_POINTS)])
    return res
