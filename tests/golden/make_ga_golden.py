"""Golden fixtures for the GA evolution operators, produced by EXECUTING THE UNMODIFIED REFERENCE
(/root/reference/genetic_algorithm_solver.py).

    python tests/golden/make_ga_golden.py      # writes tests/golden/ga_ops_*.npz, ga_solve_stats.npz

1. Operator fixtures (exact): the reference's ``_selection``, ``_crossover``, ``_mutation`` and
   ``_elitism`` (ga:183-268) are run with the module's ``random`` replaced by a SCRIPTED object that
   returns pre-drawn decisions in the order the reference asks for them.  Stored: old population,
   fitness, the decisions in the device's trace layout (include/fcpp.h), the resulting population.
2. Run statistics (statistical parity): the unmodified ``solve()`` (ga:44-135) with
   ``random.seed(s)`` for 24 seeds on a fixed 25-node instance: best distance and generations.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import random as pyrandom
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/genetic_algorithm_solver.py"

TRACE_INTS = 48


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_ga", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Scripted:
    """Stands in for the `random` module inside the reference: hands out scripted decisions."""

    def __init__(self, samples, floats):
        self.samples, self.floats = list(samples), list(floats)

    def sample(self, population, k):
        s = self.samples.pop(0)
        assert len(s) == k and all(0 <= x < len(population) for x in s), (s, k, len(population))
        return list(s)

    def random(self):
        return self.floats.pop(0)


def make_trace(rng, m_in, n, k, cross_p, mut_p):
    pairs = (m_in + 1) // 2
    tr = np.zeros((pairs, TRACE_INTS), dtype=np.int32)
    slot_draws = [rng.permutation(m_in)[:k] for _ in range(m_in)]
    for p in range(pairs):
        sa, sb = 2 * p, (2 * p + 1 if 2 * p + 1 < m_in else 0)
        tr[p, 11] = k
        tr[p, 12:12 + k] = slot_draws[sa]
        tr[p, 28:28 + k] = slot_draws[sb]
        if rng.random() < cross_p and n >= 2:
            a, b = sorted(rng.choice(n, size=2, replace=False).tolist())
            tr[p, 2:5] = (1, a, b)
        for c in range(2):
            if rng.random() < mut_p and n >= 2:
                i, j = rng.choice(n, size=2, replace=False).tolist()
                tr[p, 5 + 3 * c:8 + 3 * c] = (1, i, j)
    return tr, slot_draws


def run_reference_generation(mod, pop, fit, tr, slot_draws, cfg_kwargs):
    m_in = len(pop)
    cfg = mod.GAConfig(**cfg_kwargs)
    solver = mod.GeneticAlgorithmSolver(cfg)
    # --- _selection: one random.sample(range(m), k) per slot (ga:186-196) ---
    mod.random = Scripted([d.tolist() for d in slot_draws], [])
    selected = solver._selection([list(map(int, r)) for r in pop], list(map(float, fit)))
    # --- _crossover: random.random() per pair, then random.sample(range(n), 2) when crossing (ga:207, :219) ---
    floats, samples = [], []
    for p in range(len(tr)):
        floats.append(0.0 if tr[p, 2] else 1.0)
        if tr[p, 2]:
            samples.append([int(tr[p, 4]), int(tr[p, 3])])     # unsorted on purpose: the reference sorts
    mod.random = Scripted(samples, floats)
    offspring = solver._crossover(selected)
    assert not mod.random.samples and not mod.random.floats
    # --- _mutation: random.random() per child, random.sample(range(n), 2) when hit (ga:246-250) ---
    floats, samples = [], []
    for c in range(len(offspring)):
        t = tr[c // 2]
        o = 5 + 3 * (c & 1)
        floats.append(0.0 if t[o] else 1.0)
        if t[o]:
            samples.append([int(t[o + 1]), int(t[o + 2])])
    mod.random = Scripted(samples, floats)
    offspring = solver._mutation(offspring)
    assert not mod.random.samples and not mod.random.floats
    new = solver._elitism([list(map(int, r)) for r in pop], offspring, list(map(float, fit)))
    mod.random = pyrandom
    return np.asarray(new, dtype=np.int32), np.asarray(selected, dtype=np.int32)


CASES = {
    # name: (m_in, n, k, elite, crossover prob of the script, mutation prob of the script, duplicate rows?)
    "even": (12, 9, 5, 2, 0.8, 0.3, False),
    "odd": (11, 7, 3, 3, 0.7, 0.5, False),
    "elite0": (8, 10, 4, 0, 0.9, 0.2, False),
    "ties": (40, 33, 5, 6, 0.85, 0.1, True),
    "n2": (6, 2, 2, 1, 1.0, 1.0, False),
    "big": (64, 201, 5, 20, 0.85, 0.02, False),
}


def main():
    mod = load_reference()
    from oracle import ga_ops
    for name, (m_in, n, k, elite, cp, mp, dup) in CASES.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        pop = np.stack([rng.permutation(n) for _ in range(m_in)]).astype(np.int32)
        if dup:
            pop[5] = pop[3]
            pop[17] = pop[3]
            pop[30] = pop[29]
        xy = rng.uniform(0, 1000, size=(n, 2))
        D = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
        fit = np.asarray([1.0 / (ga_ops.tour_length(r, D) + 1e-6) for r in pop])
        tr, slot_draws = make_trace(rng, m_in, n, k, cp, mp)
        # winners (trace [0], [1]) by the reference's rule: first max in draw order (ga:190-194)
        for p in range(len(tr)):
            tr[p, 0] = ga_ops.tournament_winner(fit, tr[p, 12:12 + k])
            tr[p, 1] = ga_ops.tournament_winner(fit, tr[p, 28:28 + k])
        cfg = dict(population_size=m_in, elite_size=elite, tournament_size=k, crossover_rate=0.5, mutation_rate=0.5)
        with contextlib.redirect_stdout(io.StringIO()):
            new, selected = run_reference_generation(mod, pop, fit, tr, slot_draws, cfg)
        # the reference's selection must agree with the recorded winners
        for s in range(m_in):
            w = tr[s // 2, s & 1]
            assert np.array_equal(selected[s], pop[w]), (name, s)
        np.savez_compressed(os.path.join(HERE, f"ga_ops_{name}.npz"), pop=pop, fit=fit, D=D, trace=tr, new_pop=new,
                            elite_size=elite, tournament_size=k)
        print(f"ga_ops_{name}: {m_in}x{n} -> {new.shape}")

    # ---- run statistics of the unmodified solve() ----
    rng = np.random.default_rng(2024)
    xy = rng.uniform(0, 5000, size=(25, 2))
    D = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    cfg_kwargs = dict(population_size=60, max_generations=120, crossover_rate=0.85, mutation_rate=0.02,
                      elite_size=6, tournament_size=5, convergence_threshold=40)
    best, gens = [], []
    for s in range(24):
        pyrandom.seed(s)
        solver = mod.GeneticAlgorithmSolver(mod.GAConfig(**cfg_kwargs))
        with contextlib.redirect_stdout(io.StringIO()):
            route, stats = solver.solve(D, verbose=False)
        assert route[0] == 0 and sorted(route) == list(range(25))
        best.append(stats["best_distance"])
        gens.append(stats["generations"])
    # random-tour baseline for scale
    rnd = [ga_ops.tour_length(rng.permutation(25), D) for _ in range(2000)]
    np.savez_compressed(os.path.join(HERE, "ga_solve_stats.npz"), D=D, best=np.asarray(best), gens=np.asarray(gens),
                        random_mean=np.mean(rnd), **{f"cfg_{k}": v for k, v in cfg_kwargs.items()})
    print("ga_solve_stats: best mean %.1f std %.1f, gens mean %.1f, random tours %.1f" %
          (np.mean(best), np.std(best), np.mean(gens), np.mean(rnd)))


if __name__ == "__main__":
    main()
