"""Multi-field glue on the device behind the reference's scheduler interface.

"mfp" = multi_field_planner.py of the reference.  ``MultiFieldPlannerV38`` keeps the constructor,
attributes, ``optimize_sequence()`` and the ``FieldData`` / ``Connection`` / ``OptimizedRoute``
records of mfp:29-61, :63-233, with the numeric steps on the GPU:

  mfp:263-288  _calculate_distance_matrix   -> ``distance_matrix``   (fcpp_distance_matrix)
  mfp:290-320  _find_best_connection        -> ``connection_matrix`` (fcpp_connection_matrix: every
               ordered node pair at once; a route's connections are look-ups)
  mfp:184-191  GeneticAlgorithmSolver.solve -> the device GA (ga.py)
  mfp:213-216  total work distance          -> the reference's estimate area / W, or — closing the
               loop of BASELINE config 4 ("fitness = transit + plan length") — the real best-plan
               length of every field from ``plan_batch`` (four start corners per field, per-field
               argmin) with ``work_distance="planned"``

``optimization_method="2opt"`` (the reference's choice for fewer than 50 fields, mfp:153-162) imports ``TSPSolver`` from
``multi_field_planner_v37`` (mfp:176), a module the reference does not ship — there it fails with ModuleNotFoundError.
Here the same import is attempted first (a caller's own module wins) and otherwise the build-defined device 2-opt of
``tsp.py`` is used (no reference source, parity unpinned).  ``optimize_multi_vehicle`` hands over to
``multi_vehicle.py`` (KMeans split).  The matplotlib visualisations are out of scope.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .batch import _dev, make_candidates, plan_batch
from .ga import GAConfig, GeneticAlgorithmSolver
from .planner import TwoLayerPathPlannerV36
from .vehicle import VehicleParams


def distance_matrix(positions, device=None):
    """mfp:263-288: ``D[i, j] = ||pos_i - pos_j||`` (0 on the diagonal) for ``positions`` [n, 2]
    (row 0 = depot, then the field centroids).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor."""
    on_dev = torch.is_tensor(positions)
    dev = _dev(device if device is not None else (positions.device if on_dev else None))
    h = _lib.handle(dev.index)
    P = positions if on_dev else torch.from_numpy(np.ascontiguousarray(positions, dtype=np.float64))
    P = P.to(dev, dtype=torch.float64).contiguous()
    if P.ndim != 2 or P.shape[1] != 2:
        raise ValueError("positions must be [n, 2]")
    n = P.shape[0]
    with torch.cuda.device(dev):
        D = torch.empty((n, n), dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_distance_matrix(h.h, P.data_ptr(), n, D.data_ptr(), st))
    return D if on_dev else D.cpu().numpy()


def connection_matrix(field_vertices, depot, device=None):
    """mfp:290-320 for every ordered pair of nodes (node 0 = depot, node f+1 = field f):
    returns (C [F+1, F+1] shortest exit-vertex -> entry-vertex distances, from_index, to_index)."""
    on_dev = torch.is_tensor(field_vertices)
    dev = _dev(device if device is not None else (field_vertices.device if on_dev else None))
    h = _lib.handle(dev.index)
    V = field_vertices if on_dev else torch.from_numpy(np.ascontiguousarray(field_vertices, dtype=np.float64))
    V = V.to(dev, dtype=torch.float64).contiguous()
    if V.ndim != 3 or V.shape[1:] != (4, 2):
        raise ValueError("field_vertices must be [F, 4, 2]")
    F = V.shape[0]
    with torch.cuda.device(dev):
        Cm = torch.empty((F + 1, F + 1), dtype=torch.float64, device=dev)
        arg = torch.empty((F + 1, F + 1), dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_connection_matrix(h.h, V.data_ptr() if F else None, F, float(depot[0]), float(depot[1]),
                                             Cm.data_ptr(), arg.data_ptr(), st))
    if on_dev:
        return Cm, arg // 4, arg % 4
    a = arg.cpu().numpy()
    return Cm.cpu().numpy(), a // 4, a % 4


@dataclass
class FieldData:
    """mfp:29-38."""
    id: str
    vertices: np.ndarray
    planner: TwoLayerPathPlannerV36 = None
    centroid: Tuple[float, float] = None
    area: float = None
    entry_points: List[Tuple[np.ndarray, np.ndarray]] = None
    exit_points: List[Tuple[np.ndarray, np.ndarray]] = None


@dataclass
class Connection:
    """mfp:41-48."""
    from_field: str
    to_field: str
    from_point: np.ndarray
    to_point: np.ndarray
    distance: float


@dataclass
class OptimizedRoute:
    """mfp:51-60."""
    field_sequence: List[str]
    connections: List[Connection]
    total_transfer_distance: float
    total_work_distance: float
    total_distance: float
    optimization_method: str
    optimization_stats: dict = None


class MultiFieldPlannerV38:
    """mfp:63-320 with the numeric steps on the GPU."""

    def __init__(self, fields_definitions: List[dict], depot_point: Tuple[float, float],
                 vehicle_params: VehicleParams, num_vehicles: int = 1, optimization_method: str = "auto", *,
                 device=None, seed: Optional[int] = None, verbose: bool = False,
                 work_distance: str = "estimate"):
        if work_distance not in ("estimate", "planned"):
            raise ValueError("work_distance must be 'estimate' (mfp:213-216) or 'planned' (plan_batch)")
        self.depot = np.array(depot_point)
        self.vehicle_params = vehicle_params
        self.num_vehicles = num_vehicles
        self.optimization_method = optimization_method
        self.fields: Dict[str, FieldData] = {}
        self.verbose = verbose
        self.work_distance = work_distance
        self._device = device
        self._seed = seed
        self._conn = None
        self._prepare_fields(fields_definitions)
        if self.optimization_method == "auto":
            self.optimization_method = self._select_optimization_method()

    def _prepare_fields(self, fields_definitions: List[dict]):
        """mfp:105-151: one planner per field (centroid, area) and, per vertex, the entry / exit point with the
        bisector of the incoming and outgoing edge directions — all vertices of a field at once."""
        for spec in fields_definitions:
            planner = TwoLayerPathPlannerV36(vehicle_params=self.vehicle_params, field_vertices=spec['vertices'],
                                             device=self._device)
            P = np.asarray(planner.field_vertices, dtype=np.float64)            # [n, 2]
            into = P - np.roll(P, 1, axis=0)                                     # edge arriving at vertex i
            away = np.roll(P, -1, axis=0) - P                                    # edge leaving it
            into /= np.hypot(into[:, 0], into[:, 1])[:, None]
            away /= np.hypot(away[:, 0], away[:, 1])[:, None]
            mid = (into + away) / 2
            mlen = np.hypot(mid[:, 0], mid[:, 1])
            keep = mlen > 0.1                                                    # a U-turn vertex keeps the incoming edge
            heading = np.where(keep[:, None], mid / np.where(keep, mlen, 1.0)[:, None], into)
            gates = [(P[i].copy(), heading[i].copy()) for i in range(len(P))]
            poly = planner.field_polygon
            self.fields[spec['id']] = FieldData(id=spec['id'], vertices=np.array(spec['vertices']), planner=planner,
                                                centroid=poly.centroid.coords[0], area=poly.area,
                                                entry_points=gates, exit_points=list(gates))

    def _select_optimization_method(self) -> str:
        """mfp:153-162."""
        return "2opt" if len(self.fields) < 50 else "genetic"

    # ---- mfp:263-288 -------------------------------------------------------------------------
    def _calculate_distance_matrix(self) -> Tuple[np.ndarray, List[str]]:
        field_ids = list(self.fields.keys())
        node_ids = ["depot"] + field_ids
        pos = np.vstack([self.depot.astype(np.float64)] +
                        [np.asarray(self.fields[f].centroid, dtype=np.float64) for f in field_ids])
        return distance_matrix(pos, device=self._device), node_ids

    # ---- mfp:290-320 -------------------------------------------------------------------------
    def _connection_tables(self):
        if self._conn is None:
            ids = list(self.fields.keys())
            quads = [f for f in ids if len(self.fields[f].vertices) == 4]
            if len(quads) != len(ids):
                raise ValueError("only 4-vertex fields are supported (the planner's corner logic, mlp3:983-1007)")
            V = np.stack([np.asarray(self.fields[f].vertices, dtype=np.float64) for f in ids]) if ids else \
                np.zeros((0, 4, 2))
            Cm, fi, ti = connection_matrix(V, self.depot, device=self._device)
            self._conn = (Cm, fi, ti, {f: k + 1 for k, f in enumerate(ids)})
        return self._conn

    def _find_best_connection(self, from_id: str, to_id: str) -> Connection:
        Cm, fi, ti, index = self._connection_tables()
        a = 0 if from_id == "depot" else index[from_id]
        b = 0 if to_id == "depot" else index[to_id]
        fp = self.depot if a == 0 else self.fields[from_id].exit_points[int(fi[a, b])][0]
        tp = self.depot if b == 0 else self.fields[to_id].entry_points[int(ti[a, b])][0]
        return Connection(from_field=from_id, to_field=to_id, from_point=fp, to_point=tp, distance=float(Cm[a, b]))

    # ---- work distance ------------------------------------------------------------------------
    def planned_work_lengths(self) -> Dict[str, dict]:
        """Best plan of every field over the four start corners (plan_batch + per-field argmin):
        {field_id: {'length': len_main + len_head in m, 'start_corner': c, 'time_s': ...}}."""
        ids = list(self.fields.keys())
        fields = [[tuple(map(float, v)) for v in self.fields[f].vertices] for f in ids]
        cand = make_candidates(len(ids), start_corners=[0, 1, 2, 3])
        res = plan_batch(fields, self.vehicle_params, cand, outputs="summary", device=self._device)
        out = {}
        for k, f in enumerate(ids):
            b = int(res.best_cand[k])
            if b < 0:
                raise ValueError(f"field {f!r}: no valid plan (headland wider than the field, mlp3:597)")
            s = res.summary[b]
            out[f] = {'length': float(res.best_cost[k]), 'start_corner': int(cand["start_corner"][b]),
                      'time_s': float(s["time_main"] + s["time_head"])}
        return out

    # ---- mfp:164-233 --------------------------------------------------------------------------
    def optimize_sequence(self) -> OptimizedRoute:
        if self.num_vehicles > 1:
            raise ValueError("多机协同请使用 optimize_multi_vehicle() 方法")
        distance_matrix_, node_ids = self._calculate_distance_matrix()
        if self.optimization_method == "2opt":
            from .tsp import tsp_solver_class
            optimal_route_indices = tsp_solver_class().solve(distance_matrix_)      # mfp:176-177
            stats = {'method': '2opt'}
        else:
            config = GAConfig(population_size=min(200, len(self.fields) * 4), max_generations=500,
                              convergence_threshold=50)
            solver = GeneticAlgorithmSolver(config, seed=self._seed, device=self._device)
            optimal_route_indices, stats = solver.solve(distance_matrix_, verbose=self.verbose)
            stats['method'] = 'genetic'
        return self._route_of(optimal_route_indices, node_ids, stats)

    def _route_of(self, order, node_ids, stats) -> OptimizedRoute:
        """mfp:193-233: the fields in tour order, the best connection of every hop (depot -> first ... last -> depot:
        look-ups into the connection tables) and the totals."""
        sequence = [node_ids[k] for k in order if node_ids[k] != "depot"]
        stops = ["depot"] + sequence + ["depot"]
        hops = [self._find_best_connection(a, b) for a, b in zip(stops[:-1], stops[1:])]
        transfer = 0
        for hop in hops:                       # left to right, as the reference accumulates (mfp:199-210)
            transfer += hop.distance
        if self.work_distance == "planned":
            stats['planned'] = planned = self.planned_work_lengths()
            work = sum(planned[f]['length'] for f in sequence)
        else:
            width = self.vehicle_params.working_width
            work = sum(self.fields[f].area / width for f in sequence)
        return OptimizedRoute(field_sequence=sequence, connections=hops, total_transfer_distance=transfer,
                              total_work_distance=work, total_distance=transfer + work,
                              optimization_method=self.optimization_method, optimization_stats=stats)

    def optimize_multi_vehicle(self):
        """mfp:235-261: field data (centroid, area, vertices) handed to MultiVehiclePlanner.plan — KMeans split on
        the device, one GA per vehicle (multi_vehicle.py)."""
        if self.num_vehicles == 1:
            raise ValueError("单机优化请使用 optimize_sequence() 方法")
        from .multi_vehicle import MultiVehiclePlanner
        fields_data = {fid: {'centroid': f.centroid, 'area': f.area, 'vertices': f.vertices}
                       for fid, f in self.fields.items()}
        mvp = MultiVehiclePlanner(num_vehicles=self.num_vehicles, optimization_method=self.optimization_method,
                                  device=self._device, seed=self._seed, verbose=self.verbose)
        return mvp.plan(fields_data, tuple(self.depot), self.vehicle_params, self.optimization_method == "genetic")
