"""CPU restatement of the reference's GA evolution operators (TEST INFRASTRUCTURE ONLY).

"ga" = /root/reference/genetic_algorithm_solver.py.  The reference draws every random decision
from Python's global ``random``; here the decisions are EXPLICIT inputs (the trace layout of
include/fcpp.h, FCPP_GA_TRACE_INTS per offspring pair), so that one generation becomes a pure
function  (population, fitness, decisions) -> population  that can be compared bit for bit with
the device kernels.  Pinned against the unmodified reference operators executed with a scripted
``random`` module: tests/golden/make_ga_golden.py -> tests/golden/ga_ops_*.npz.

Trace row of pair p (int32[48]):
  [0] parent A = winner of tournament slot 2p          [1] parent B = winner of slot 2p+1
      (slot 0 again when 2p+1 == len(population), ga:205)
  [2] crossed  [3] a  [4] b   (a < b: sorted(random.sample(range(n), 2)), ga:219)
  [5] child-1 mutated [6] i [7] j   [8] child-2 mutated [9] i [10] j   (ga:246-250)
  [11] k = tournament size   [12:12+k] draws of slot A   [28:28+k] draws of slot B (ga:189)
"""
from __future__ import annotations

import numpy as np

TRACE_INTS = 48
MAX_TOURNAMENT = 16


def next_size(m_in: int, elite_size: int) -> int:
    """len(new_population[:-elite] + elites), ga:262-266, python slicing semantics included."""
    children = 2 * ((m_in + 1) // 2)
    if elite_size <= 0:            # new[:-0] == [] and argsort[-0:] == everything
        return m_in
    return max(children - elite_size, 0) + min(elite_size, m_in)


def tournament_winner(fit, draws):
    """ga:190-194: np.argmax of the drawn fitness values = FIRST maximum in draw order."""
    best, best_f = None, None
    for x in draws:
        f = fit[x]
        if best is None or f > best_f:
            best, best_f = int(x), f
    return best


def ox_child(seg_parent, fill_parent, a, b):
    """ga:214-242 for one child: child[a:b] = seg_parent[a:b]; the remaining genes in the order
    fill_parent[b:] + fill_parent[:b] go to positions b, b+1, ..., wrapping to 0."""
    n = len(seg_parent)
    child = np.full(n, -1, dtype=np.int64)
    child[a:b] = seg_parent[a:b]
    used = np.zeros(n, dtype=bool)
    used[seg_parent[a:b]] = True
    pos = b
    for g in list(fill_parent[b:]) + list(fill_parent[:b]):
        if not used[g]:
            if pos >= n:
                pos = 0
            child[pos] = g
            pos += 1
    return child


def elites_ascending(fit, e_take):
    """ga:259-260 ``np.argsort(old_fitness)[-elite:]`` with ties resolved as a STABLE sort (numpy's
    default quicksort leaves the order of equal keys unspecified; the device uses the stable rule)."""
    order = np.argsort(np.asarray(fit), kind="stable")
    return order[len(order) - e_take:]


def replay_generation(pop, fit, trace, elite_size):
    """One generation ga:78-88 = _selection + _crossover + _mutation + _elitism with the random
    decisions taken from ``trace``.  Returns the new population as an int array."""
    pop = np.asarray(pop, dtype=np.int64)
    fit = np.asarray(fit, dtype=np.float64)
    m_in, n = pop.shape
    pairs = (m_in + 1) // 2
    children = []
    for p in range(pairs):
        t = trace[p]
        k = int(t[11])
        wa = tournament_winner(fit, t[12:12 + k])
        wb = tournament_winner(fit, t[28:28 + k])
        assert wa == t[0] and wb == t[1], (p, wa, wb, t[:2])
        pa, pb = pop[wa], pop[wb]
        if t[2]:
            a, b = int(t[3]), int(t[4])
            c1, c2 = ox_child(pa, pb, a, b), ox_child(pb, pa, a, b)
        else:
            c1, c2 = pa.copy(), pb.copy()
        children += [c1, c2]
    for c, child in enumerate(children):          # ga:246-250
        t = trace[c // 2]
        o = 5 + 3 * (c & 1)
        if t[o]:
            i, j = int(t[o + 1]), int(t[o + 2])
            child[i], child[j] = child[j], child[i]
    if elite_size <= 0:
        keep, e_take = 0, m_in
    else:
        keep, e_take = max(len(children) - elite_size, 0), min(elite_size, m_in)
    el = elites_ascending(fit, e_take)
    out = children[:keep] + [pop[i].copy() for i in el]
    return np.asarray(out, dtype=np.int32).reshape(len(out), n)


def check_trace(trace, m_in, n, k):
    """Structural validity of a device trace: draws distinct and in range, cut points ordered,
    mutation positions distinct; the odd one out reuses slot 0 (ga:205)."""
    pairs = (m_in + 1) // 2
    for p in range(pairs):
        t = trace[p]
        assert t[11] == k
        for o in (12, 28):
            d = t[o:o + k]
            assert len(set(d.tolist())) == k and d.min() >= 0 and d.max() < m_in, (p, d)
        if t[2]:
            assert 0 <= t[3] < t[4] < n, (p, t[3], t[4])
        for o in (5, 8):
            if t[o]:
                assert t[o + 1] != t[o + 2] and 0 <= t[o + 1] < n and 0 <= t[o + 2] < n
    if m_in & 1:
        assert np.array_equal(trace[pairs - 1][28:28 + k], trace[0][12:12 + k])


def tour_length(route, D):
    """ga:174-181 sequential FP64 sum of the closed tour."""
    s = 0.0
    n = len(route)
    for i in range(n):
        s += D[route[i], route[(i + 1) % n]]
    return s
