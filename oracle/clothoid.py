"""Clothoid (Euler-spiral) turn model — oracle side (TEST INFRASTRUCTURE ONLY).

The reference only *describes* Clothoid turns (README.md:105-113) and sketches the formula in a
roadmap (doc/两层路径规划器 - 深度优化和改进路线图.md:31-41: ``C, S = fresnel(s*sqrt(c/pi));
x = C*sqrt(pi/c); y = S*sqrt(pi/c)``); no planner code uses it.  The
build-defined opt-in ``turn_model="clothoid"`` is restated here with ``scipy.special.fresnel``;
**parity with the reference is unpinned** (there is nothing to compare with).

Turn = line -> clothoid -> arc -> clothoid -> line with total deflection PHI, maximum curvature
1/R on the arc, and a share ``lam`` of the deflection spent on the two clothoids (alpha = lam*PHI/2
each).  Everything scales with R, so the turn is described by unit-radius local coordinates
(xi along the entry heading, eta to the turning side) sampled at n equal arc-length steps —
the reference's sample counts (20 per U-turn, 15 per corner) are kept so the layout is unchanged.
"""
import numpy as np
from scipy.special import fresnel


def cac_unit(phi: float, n: int, lam: float):
    """(xi, eta) [n] of the unit-radius clothoid-arc-clothoid turn."""
    alpha = lam * phi / 2
    Lc = 2 * alpha
    La = phi - 2 * alpha
    Lt = 2 * Lc + La
    s = np.linspace(0, Lt, n)
    if Lc <= 0:
        return np.sin(s), 1 - np.cos(s)
    a = np.sqrt(np.pi * Lc)
    S1, C1 = fresnel(Lc / a)
    p1x, p1y = a * C1, a * S1
    xi = np.empty(n)
    eta = np.empty(n)
    for i, si in enumerate(s):
        if si <= Lc:
            S, C = fresnel(si / a)
            xi[i], eta[i] = a * C, a * S
        elif si <= Lc + La:
            f = si - Lc
            xi[i] = p1x - np.sin(alpha) + np.sin(alpha + f)
            eta[i] = p1y + np.cos(alpha) - np.cos(alpha + f)
        else:
            p2x = p1x - np.sin(alpha) + np.sin(alpha + La)
            p2y = p1y + np.cos(alpha) - np.cos(alpha + La)
            cb, sb = -np.cos(phi), -np.sin(phi)          # rotation by phi + pi
            m1x, m1y = a * C1, -a * S1
            ex = p2x - (cb * m1x - sb * m1y)
            ey = p2y - (sb * m1x + cb * m1y)
            S, C = fresnel((Lt - si) / a)
            mx, my = a * C, -a * S
            xi[i] = ex + (cb * mx - sb * my)
            eta[i] = ey + (sb * mx + cb * my)
    return xi, eta
