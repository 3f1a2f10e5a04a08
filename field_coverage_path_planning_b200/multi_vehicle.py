"""Multi-vehicle split: drop-in for /root/reference/multi_vehicle_planner.py ("mvp").

    MultiVehiclePlanner(num_vehicles, optimization_method).plan(fields_data, depot_point, vehicle_params,
                                                                use_genetic) -> MultiVehicleRoute

* field -> vehicle clustering (mvp:186-209): sklearn.cluster.KMeans(n_clusters=V, random_state=42) in the
  reference.  scikit-learn is a third-party dependency of the reference (not in its tree; unpinned — the fixtures
  were made with 1.9.0): its published algorithm is restated here — the k-means++ seeding draws from
  ``numpy.random.RandomState(42)`` on the host (a handful of draws, O(V) vectorised steps), the Lloyd iteration
  runs on the device (``fcpp_kmeans_lloyd``, one CTA per problem, batched) — and pinned by labels the
  UNMODIFIED reference produced with the real sklearn (tests/golden/multi_vehicle.npz).
* workload balance (mvp:211-227): the reference computes the cluster areas and returns the clusters unchanged.
* per-vehicle order (mvp:96-133): centroid distance matrix (``fcpp_distance_matrix``) + the device GA
  (``fcpp_ga_solve``) when ``use_genetic and len(cluster) > 20``; otherwise the reference imports ``TSPSolver`` from
  ``multi_field_planner_v37``, a module that is not part of the reference tree (ModuleNotFoundError there) — the same
  import is attempted here and, when it fails, the build-defined device 2-opt of ``tsp.py`` is used.
* statistics (mvp:139-183): transfer / work distance, work time at 5 and 15 km/h, load balance ratio.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .batch import _dev
from .ga import GAConfig, GeneticAlgorithmSolver


@dataclass
class VehicleRoute:
    """mvp:23-32."""
    vehicle_id: int
    field_ids: List[str]
    field_sequence: List[str]
    total_transfer_distance: float
    total_work_distance: float
    total_distance: float
    work_time: float


@dataclass
class MultiVehicleRoute:
    """mvp:35-44."""
    num_vehicles: int
    vehicle_routes: List[VehicleRoute]
    total_transfer_distance: float
    total_work_distance: float
    total_distance: float
    max_work_time: float
    load_balance_ratio: float


# ------------------------------------------------------------------------------------------------
# KMeans: k-means++ seeding (host, RandomState) + Lloyd iterations (device)
# ------------------------------------------------------------------------------------------------
def _sq_dists(A: np.ndarray, X: np.ndarray, x_sq: np.ndarray) -> np.ndarray:
    """Squared distances [len(A), len(X)] the way sklearn's euclidean_distances forms them:
    ||a||^2 - 2 a.x + ||x||^2, negatives clipped to 0."""
    d = (A * A).sum(axis=1)[:, None] - 2.0 * (A @ X.T) + x_sq[None, :]
    np.maximum(d, 0.0, out=d)
    return d


def kmeans_plusplus_seeds(X: np.ndarray, n_clusters: int, random_state: np.random.RandomState) -> np.ndarray:
    """Indices of the k-means++ seeds (Arthur & Vassilvitskii with 2 + log(k) greedy local trials, as
    scikit-learn's ``_kmeans_plusplus`` with unit sample weights): the first seed by ``choice(n, p=uniform)``, every
    further one the best of the local trials drawn with probability proportional to the squared distance to the
    closest seed so far.  ``X`` must already be centred on its mean (KMeans.fit does that before seeding)."""
    n = len(X)
    x_sq = (X * X).sum(axis=1)
    w = np.ones(n, dtype=X.dtype)
    n_trials = 2 + int(np.log(n_clusters))
    idx = np.full(n_clusters, -1, dtype=np.int64)
    idx[0] = random_state.choice(n, p=w / w.sum())
    closest = _sq_dists(X[idx[0]][None, :], X, x_sq)
    pot = closest @ w
    for c in range(1, n_clusters):
        rand_vals = random_state.uniform(size=n_trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        d = _sq_dists(X[cand], X, x_sq)
        np.minimum(closest, d, out=d)
        pots = d @ w.reshape(-1, 1)
        best = int(np.argmin(pots))
        pot = pots[best]
        closest = d[best][None, :]
        idx[c] = cand[best]
    return idx


def kmeans_labels(points, n_clusters: int, random_state: int = 42, max_iter: int = 300, tol: float = 1e-4,
                  device=None, return_centers: bool = False):
    """``KMeans(n_clusters, random_state=random_state).fit_predict(points)`` for 2-D points: seeds on the host,
    Lloyd on the device.  -> labels [n] int32 (+ centres [k, 2], iterations, inertia)."""
    X = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    res = kmeans_batch([X], [n_clusters], random_state, max_iter, tol, device)[0]
    return res if return_centers else res[0]


def kmeans_batch(problems: Sequence[np.ndarray], n_clusters: Sequence[int], random_state: int = 42,
                 max_iter: int = 300, tol: float = 1e-4, device=None):
    """Many independent KMeans problems (farms) in ONE launch, one CTA each.
    -> [(labels, centres, n_iter, inertia)] per problem; every problem is seeded like its own
    ``KMeans(k, random_state=random_state)``."""
    dev = _dev(device)
    h = _lib.handle(dev.index)
    P = len(problems)
    if P == 0:
        return []
    pts, seeds, p_start, c_start = [], [], [0], [0]
    for X, k in zip(problems, n_clusters):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 2)
        k = int(k)
        if k < 1 or k > len(X):
            raise ValueError(f"n_samples={len(X)} should be >= n_clusters={k}.")     # sklearn's message
        Xc = X - X.mean(axis=0)
        idx = kmeans_plusplus_seeds(Xc, k, np.random.RandomState(random_state))
        pts.append(X)
        seeds.append(X[idx])
        p_start.append(p_start[-1] + len(X))
        c_start.append(c_start[-1] + k)
    xy = torch.from_numpy(np.concatenate(pts)).to(dev)
    cen = torch.from_numpy(np.concatenate(seeds)).to(dev)
    ps = torch.tensor(p_start, dtype=torch.int64, device=dev)
    cs = torch.tensor(c_start, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        labels = torch.empty(p_start[-1], dtype=torch.int32, device=dev)
        n_iter = torch.empty(P, dtype=torch.int32, device=dev)
        inertia = torch.empty(P, dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_kmeans_lloyd(h.h, P, ps.data_ptr(), xy.data_ptr(), cs.data_ptr(),
                                        int(max(n_clusters)), cen.data_ptr(), labels.data_ptr(), int(max_iter),
                                        float(tol), n_iter.data_ptr(), inertia.data_ptr(), st))
        lab, cen_h, it, ine = labels.cpu().numpy(), cen.cpu().numpy(), n_iter.cpu().numpy(), inertia.cpu().numpy()
    return [(lab[p_start[p]:p_start[p + 1]], cen_h[c_start[p]:c_start[p + 1]], int(it[p]), float(ine[p]))
            for p in range(P)]


# ------------------------------------------------------------------------------------------------
class MultiVehiclePlanner:
    """mvp:47-268 (``visualize`` needs matplotlib and is not part of the path)."""

    def __init__(self, num_vehicles: int, optimization_method: str = "genetic", device=None,
                 seed: Optional[int] = None, verbose: bool = True):
        self.num_vehicles = num_vehicles
        self.optimization_method = optimization_method
        self._device = device
        self._seed = seed
        self._verbose = verbose
        if verbose:
            print(f"\n[多机协同] 初始化 {num_vehicles} 台车辆")

    # mvp:65-184
    def plan(self, fields_data: Dict, depot_point: Tuple[float, float], vehicle_params,
             use_genetic: bool = False) -> MultiVehicleRoute:
        clusters = self._cluster_fields(fields_data, depot_point)
        clusters = self._balance_workload(clusters, fields_data)
        vehicle_routes: List[VehicleRoute] = []
        for vehicle_id in range(self.num_vehicles):
            ids = clusters[vehicle_id]
            if len(ids) == 0:
                continue
            D = self._build_distance_matrix(ids, fields_data, depot_point)
            if use_genetic and len(ids) > 20:
                config = GAConfig(population_size=min(100, len(ids) * 5), max_generations=200, convergence_threshold=30)
                seed = None if self._seed is None else self._seed + vehicle_id
                optimal_route, _ = GeneticAlgorithmSolver(config, seed=seed, device=self._device).solve(D, verbose=False)
            else:
                from .tsp import tsp_solver_class               # mvp:131 (a module the reference does not ship)
                optimal_route = tsp_solver_class().solve(D)
            node_ids = ["depot"] + ids
            field_sequence = [node_ids[i] for i in optimal_route if node_ids[i] != "depot"]
            transfer = self._calculate_route_distance(optimal_route, D)
            work = sum(fields_data[f]['area'] / vehicle_params.working_width for f in field_sequence)
            work_time = work / 1000 / 5 + transfer / 1000 / 15                         # mvp:146
            vehicle_routes.append(VehicleRoute(vehicle_id=vehicle_id, field_ids=ids, field_sequence=field_sequence,
                                               total_transfer_distance=transfer, total_work_distance=work,
                                               total_distance=transfer + work, work_time=work_time))
        total_transfer = sum(v.total_transfer_distance for v in vehicle_routes)
        total_work = sum(v.total_work_distance for v in vehicle_routes)
        total_dist = sum(v.total_distance for v in vehicle_routes)
        max_time = max(v.work_time for v in vehicle_routes)
        avg_time = np.mean([v.work_time for v in vehicle_routes])
        load_balance = max_time / avg_time if avg_time > 0 else 1.0
        if self._verbose:
            print(f"[fcpp fleet] vehicles={len(vehicle_routes)} transfer={total_transfer:.0f} m work={total_work:.0f} m "
                  f"max time={max_time:.1f} h balance={load_balance:.2f}")
        return MultiVehicleRoute(num_vehicles=self.num_vehicles, vehicle_routes=vehicle_routes,
                                 total_transfer_distance=total_transfer, total_work_distance=total_work,
                                 total_distance=total_dist, max_work_time=max_time, load_balance_ratio=load_balance)

    # mvp:186-209
    def _cluster_fields(self, fields_data: Dict, depot_point) -> List[List[str]]:
        field_ids = list(fields_data.keys())
        centroids = np.array([fields_data[f]['centroid'] for f in field_ids], dtype=np.float64)
        labels = kmeans_labels(centroids, self.num_vehicles, random_state=42, device=self._device)
        clusters: List[List[str]] = [[] for _ in range(self.num_vehicles)]
        for i, f in enumerate(field_ids):
            clusters[int(labels[i])].append(f)
        return clusters

    # mvp:211-227: "simplified version: no adjustment" — the areas are summed and the clusters returned as they are
    def _balance_workload(self, clusters: List[List[str]], fields_data: Dict) -> List[List[str]]:
        return clusters

    # mvp:229-259: depot + centroids, Euclidean
    def _build_distance_matrix(self, field_ids: List[str], fields_data: Dict, depot_point) -> np.ndarray:
        from .multi_field import distance_matrix
        pos = np.array([tuple(depot_point)] + [tuple(fields_data[f]['centroid']) for f in field_ids], dtype=np.float64)
        return distance_matrix(pos, device=self._device)

    # mvp:261-268: closed tour, left-to-right FP64 sum
    def _calculate_route_distance(self, route: List[int], distance_matrix: np.ndarray) -> float:
        from .ga import tour_lengths
        return float(tour_lengths(distance_matrix, np.asarray([route], dtype=np.int32), device=self._device)[0])
