"""Oracle: ``lap.lapjv(cost, extend_cost=True, cost_limit=t)`` restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: ``lap``
(PyPI ``lap``, gatagat/lap; no version pinned by the reference, absent from the
image) is called at tracker/matching.py:34.  Published algorithm (lap 0.4
``_lapjv.pyx``): when ``extend_cost`` or a finite ``cost_limit`` is given, build

    ext = full((N+M, N+M), cost_limit / 2)      # cost.max()+1 without a limit
    ext[N:, M:] = 0 ;  ext[:N, :M] = cost

solve the square LAP on ``ext`` exactly (Jonker-Volgenant), then
``x[x >= M] = -1``, ``y[y >= N] = -1`` and truncate to N / M entries.
Equivalent objective: minimise  sum_matched (c_ij - cost_limit).

The exact solver used here is ``scipy.optimize.linear_sum_assignment`` (the CPU
baseline BASELINE.json names).  When the optimum is unique every exact solver
returns the same x / y; ``is_unique`` below tells the parity tests when indices
(and not only the objective) must agree.
"""
import itertools

import numpy as np
from scipy.optimize import linear_sum_assignment


def extend(cost, cost_limit):
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    if np.isfinite(cost_limit):
        ext = np.full((n + m, n + m), cost_limit / 2.0, dtype=np.float64)
    else:
        ext = np.full((n + m, n + m), cost.max() + 1.0, dtype=np.float64)
    ext[n:, m:] = 0.0
    ext[:n, :m] = cost
    return ext


def lapjv(cost, extend_cost=True, cost_limit=np.inf):
    """Returns (opt, x, y) with lap's conventions: x[i] = column of row i or -1."""
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    if not (extend_cost or np.isfinite(cost_limit)):
        if n != m:
            raise ValueError("Square cost array expected. Pass extend_cost=True for non-square input.")
        r, c = linear_sum_assignment(cost)
        x = np.empty(n, dtype=np.int64); x[r] = c
        y = np.empty(n, dtype=np.int64); y[c] = r
        return float(cost[r, c].sum()), x, y
    ext = extend(cost, cost_limit)
    r, c = linear_sum_assignment(ext)
    xx = np.empty(n + m, dtype=np.int64); xx[r] = c
    yy = np.empty(n + m, dtype=np.int64); yy[c] = r
    x = xx[:n].copy(); y = yy[:m].copy()
    x[x >= m] = -1
    y[y >= n] = -1
    opt = float(cost[np.nonzero(x >= 0)[0], x[x >= 0]].sum()) if n and m else 0.0
    return opt, x, y


def linear_assignment(cost_matrix, thresh):
    """tracker/matching.py:30-41 (post-processing of x, y into matches / unmatched)."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return (np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])),
                tuple(range(cost_matrix.shape[1])))
    _, x, y = lapjv(cost_matrix, extend_cost=True, cost_limit=thresh)
    matches = [[ix, mx] for ix, mx in enumerate(x) if mx >= 0]
    unmatched_a = np.where(x < 0)[0]
    unmatched_b = np.where(y < 0)[0]
    return np.asarray(matches), unmatched_a, unmatched_b


def objective(cost, x, thresh):
    """sum over matched pairs of (c_ij - thresh): the quantity every exact solver minimises."""
    cost = np.asarray(cost, dtype=np.float64)
    x = np.asarray(x)
    rows = np.nonzero(x >= 0)[0]
    return float((cost[rows, x[rows]] - thresh).sum())


def brute_force(cost, thresh):
    """Exhaustive optimum of sum_matched (c - thresh) for tiny problems (tests only)."""
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    best, best_x, ties = 0.0, -np.ones(n, dtype=np.int64), 1
    cols = list(range(m)) + [-1] * n
    seen = set()
    for perm in itertools.permutations(cols, n):
        if perm in seen:
            continue
        seen.add(perm)
        val = sum(cost[i, j] - thresh for i, j in enumerate(perm) if j >= 0)
        if val < best - 1e-12:
            best, best_x, ties = val, np.array(perm, dtype=np.int64), 1
        elif abs(val - best) <= 1e-12 and not np.array_equal(best_x, np.array(perm)):
            ties += 1
    return best, best_x, ties == 1


def is_unique(cost, x, thresh, eps=1e-9):
    """Heuristic uniqueness check: perturb-and-resolve.  True when every matched pair survives
    a re-solve in which it is made slightly more expensive only by being forbidden -> the
    optimum strictly worsens by more than eps."""
    cost = np.asarray(cost, dtype=np.float64)
    base = objective(cost, x, thresh)
    rows = np.nonzero(np.asarray(x) >= 0)[0]
    for i in rows:
        c2 = cost.copy()
        c2[i, x[i]] = thresh + 1.0
        _, x2, _ = lapjv(c2, True, thresh)
        if objective(c2, x2, thresh) <= base + eps:
            return False
    # unmatched rows: forcing a match must cost strictly more
    return True
