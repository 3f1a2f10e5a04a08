"""Batched GA tour-length fitness (A12) and a drop-in ``GeneticAlgorithmSolver``.

``tour_lengths`` is the hot path named by north_star: the whole population's closed-tour
lengths in one kernel launch (genetic_algorithm_solver.py:168-181, "ga").  The FP64 sum runs
left to right per tour exactly like the reference loop, so lengths are bit-identical.

``GeneticAlgorithmSolver`` keeps the reference's interface (``GAConfig``, ``solve(distance_matrix,
verbose) -> (route, stats)``, ``best_fitness_history`` / ``avg_fitness_history``).  By default
(``operators="device"``) the WHOLE solve loop runs on the GPU through ``fcpp_ga_solve``
(SURVEY.md §8(f) N1): tournament selection (ga:183-196), OX crossover (ga:198-242), swap mutation
(ga:244-252), elitism that overwrites the LAST ``elite_size`` children (ga:254-268), fitness,
best tracking and the stop after ``convergence_threshold`` stagnant generations (ga:113-116), the
final rotation to node 0 (ga:119-120) — two generations per CUDA graph, the host only polls the
convergence flag.  The reference draws from Python's unseeded global ``random``; here a Philox
counter-based generator keyed by ``seed`` makes runs reproducible.  Whole-run parity is
statistical; operator parity is exact given the same decisions (``ga_generation(...,
return_trace=True)`` exposes them; tests replay them through the reference's operators).
``operators="host"`` keeps the numpy restatement of the operators with only the fitness on the GPU.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .batch import _dev


def tour_lengths(distance_matrix, population, device=None, return_fitness: bool = False,
                 distributed: bool = False):
    """Closed-tour lengths of ``population`` [pop, n] (permutations of 0..n-1) under
    ``distance_matrix`` [n, n] (FP64; layout of multi_field_planner.py:263-288, node 0 = depot).

    Accepts numpy arrays (copied to the device and back) or CUDA tensors (results stay on the
    device).  ``return_fitness`` adds 1/(d + 1e-6) (ga:172).  ``distributed=True`` shards the
    population over the ranks of the torch.distributed job and all-gathers the lengths."""
    dev = _dev(device if device is not None else (population.device if torch.is_tensor(population) else None))
    h = _lib.handle(dev.index)
    on_dev = torch.is_tensor(population)
    D = distance_matrix if torch.is_tensor(distance_matrix) else torch.from_numpy(
        np.ascontiguousarray(distance_matrix, dtype=np.float64))
    P = population if on_dev else torch.from_numpy(np.ascontiguousarray(population, dtype=np.int32))
    D = D.to(dev, dtype=torch.float64).contiguous()
    P = P.to(dev, dtype=torch.int32).contiguous()
    n = D.shape[0]
    if D.shape != (n, n) or P.ndim != 2 or P.shape[1] != n:
        raise ValueError("distance_matrix must be [n, n] and population [pop, n]")
    pop = P.shape[0]
    lo, hi = 0, pop
    if distributed:
        import torch.distributed as dist
        ws, rk = dist.get_world_size(), dist.get_rank()
        per = (pop + ws - 1) // ws
        lo, hi = min(pop, rk * per), min(pop, (rk + 1) * per)
    with torch.cuda.device(dev):
        out = torch.empty(max(hi - lo, 1), dtype=torch.float64, device=dev)
        fit = torch.empty(max(hi - lo, 1), dtype=torch.float64, device=dev) if return_fitness else None
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        Ps = P[lo:hi]
        h.check(h.lib.fcpp_tour_lengths(h.h, D.data_ptr(), n, Ps.data_ptr() if hi > lo else None, hi - lo,
                                        out.data_ptr(), fit.data_ptr() if fit is not None else None, st))
        out = out[:hi - lo]
        if fit is not None:
            fit = fit[:hi - lo]
        if distributed:
            import torch.distributed as dist
            ws = dist.get_world_size()
            per = (pop + ws - 1) // ws
            pad = torch.zeros(per, dtype=torch.float64, device=dev)
            pad[:hi - lo] = out
            allv = torch.empty(per * ws, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allv, pad)
            out = allv[:pop]
            if fit is not None:
                fit = 1.0 / (out + 1e-6)
    if on_dev:
        return (out, fit) if return_fitness else out
    if return_fitness:
        return out.cpu().numpy(), fit.cpu().numpy()
    return out.cpu().numpy()


def _cfg_c(config, seed: int, check_every: int = 0) -> "_lib.GAConfigC":
    return _lib.GAConfigC(int(config.population_size), int(config.max_generations), float(config.crossover_rate),
                          float(config.mutation_rate), int(config.elite_size), int(config.tournament_size),
                          int(config.convergence_threshold), int(check_every), int(seed) & (2 ** 64 - 1))


def ga_init_population(config, n: int, seed: int = 0, device=None) -> torch.Tensor:
    """ga:137-166 on the device: 2*(population_size//2) individuals [pop, n] int32 (CUDA tensor)."""
    dev = _dev(device)
    h = _lib.handle(dev.index)
    m = 2 * (int(config.population_size) // 2)
    with torch.cuda.device(dev):
        pop = torch.empty((m, n), dtype=torch.int32, device=dev)
        cfg = _cfg_c(config, seed)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_ga_init_population(h.h, C.byref(cfg), n, pop.data_ptr(), st))
    return pop


def ga_generation(config, population, fitness, generation: int = 0, seed: int = 0, device=None,
                  return_trace: bool = False):
    """One generation ga:78-88 on the device: selection + OX crossover + mutation + elitism.

    ``population`` [m, n] int32 and ``fitness`` [m] float64 (numpy or CUDA tensors) -> new
    population (CUDA tensor, ``fcpp_ga_next_size`` rows).  ``return_trace`` adds the int32
    [(m+1)//2, 48] decision trace (layout in include/fcpp.h)."""
    dev = _dev(device if device is not None else (population.device if torch.is_tensor(population) else None))
    h = _lib.handle(dev.index)
    P = population if torch.is_tensor(population) else torch.from_numpy(np.ascontiguousarray(population, dtype=np.int32))
    F = fitness if torch.is_tensor(fitness) else torch.from_numpy(np.ascontiguousarray(fitness, dtype=np.float64))
    P = P.to(dev, dtype=torch.int32).contiguous()
    F = F.to(dev, dtype=torch.float64).contiguous()
    m, n = P.shape
    if F.shape != (m,):
        raise ValueError("fitness must be [population]")
    cfg = _cfg_c(config, seed)
    m_out = int(h.lib.fcpp_ga_next_size(C.byref(cfg), m))
    with torch.cuda.device(dev):
        out = torch.empty((m_out, n), dtype=torch.int32, device=dev)
        trace = torch.zeros(((m + 1) // 2, _lib.GA_TRACE_INTS), dtype=torch.int32, device=dev) if return_trace else None
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_ga_generation(h.h, C.byref(cfg), int(generation), n, P.data_ptr(), F.data_ptr(), m,
                                         out.data_ptr(), trace.data_ptr() if trace is not None else None, st))
    return (out, trace) if return_trace else out


def ga_solve_device(config, distance_matrix, seed: int = 0, initial_population=None, device=None,
                    check_every: int = 0):
    """``solve()`` ga:44-135 entirely on the device.  Returns (route list, stats dict, history [G, 2])."""
    dev = _dev(device)
    h = _lib.handle(dev.index)
    D = distance_matrix if torch.is_tensor(distance_matrix) else torch.from_numpy(
        np.ascontiguousarray(distance_matrix, dtype=np.float64))
    D = D.to(dev, dtype=torch.float64).contiguous()
    n = D.shape[0]
    if D.shape != (n, n):
        raise ValueError("distance_matrix must be [n, n]")
    P0 = None
    if initial_population is not None:
        P0 = initial_population if torch.is_tensor(initial_population) else torch.from_numpy(
            np.ascontiguousarray(initial_population, dtype=np.int32))
        P0 = P0.to(dev, dtype=torch.int32).contiguous()
        if P0.shape != (int(config.population_size), n):
            raise ValueError("initial_population must be [population_size, n]")
    cfg = _cfg_c(config, seed, check_every)
    res = _lib.GAResultC()
    G = max(int(config.max_generations), 0)
    with torch.cuda.device(dev):
        route = torch.empty(n, dtype=torch.int32, device=dev)
        hist = torch.zeros((max(G, 1), 2), dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        h.check(h.lib.fcpp_ga_solve(h.h, C.byref(cfg), D.data_ptr(), n, P0.data_ptr() if P0 is not None else None,
                                    route.data_ptr(), hist.data_ptr(), C.byref(res), st))
        route_h = route.cpu().tolist()
        hist_h = hist[:res.generations].cpu().numpy()
    stats = {'generations': int(res.generations), 'best_distance': float(res.best_distance),
             'best_fitness': float(res.best_fitness), 'convergence_gen': int(res.convergence_gen),
             'final_population': int(res.final_population)}
    return route_h, stats, hist_h


@dataclass
class GAConfig:
    """ga:20-29."""
    population_size: int = 200
    max_generations: int = 500
    crossover_rate: float = 0.85
    mutation_rate: float = 0.02
    elite_size: int = 20
    tournament_size: int = 5
    convergence_threshold: int = 50


class GeneticAlgorithmSolver:
    """Permutation GA for the multi-field TSP ordering (ga:32-268) with GPU fitness."""

    def __init__(self, config: GAConfig = None, seed: Optional[int] = None, device=None):
        self.config = config or GAConfig()
        self.best_fitness_history: List[float] = []
        self.avg_fitness_history: List[float] = []
        # the reference is unseeded (global `random`): without a seed every solver draws a fresh one
        self.seed = int(seed) if seed is not None else int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        self._device = device

    # -- fitness on the device (the hot path, ga:168-181) ----------------------------------------
    def _calculate_distance(self, route, distance_matrix) -> float:
        return float(tour_lengths(distance_matrix, np.asarray([route], dtype=np.int32), device=self._device)[0])

    def _calculate_fitness(self, route, distance_matrix) -> float:
        return 1.0 / (self._calculate_distance(route, distance_matrix) + 1e-6)

    def solve(self, distance_matrix: np.ndarray, verbose: bool = True) -> Tuple[List[int], dict]:
        """ga:44-135."""
        cfg = self.config
        t0 = time.time()
        n = len(distance_matrix)
        # selection, OX crossover, swap mutation, elitism, fitness and best tracking all run on the device
        # (fcpp_ga_solve); there is no host operator path
        route, stats, hist = ga_solve_device(cfg, distance_matrix, seed=self.seed, device=self._device)
        self.best_fitness_history = hist[:, 0].tolist()
        self.avg_fitness_history = hist[:, 1].tolist()
        stats['time'] = time.time() - t0
        if verbose:
            print(f"[fcpp GA] nodes={n} generations={stats['generations']} best={stats['best_distance']:.1f} m "
                  f"time={stats['time']:.2f} s")
        return route, stats
