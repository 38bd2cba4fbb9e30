"""
Zero-sum matrix-game solver used by `environment.tree.Tree.generate`.

The reference delegates this to the third-party `pygambit==16.0.2`
(reference requirements.txt:3; call sites environment/tree.py:199-234:
`enummixed_solve`, fallback `lcp_solve`, then a sort that puts pure
equilibria first).  pygambit is not vendored with the reference and is not
installable here, so this module restates the published algorithm it relies
on for two-player zero-sum games: Shapley-Snow support enumeration.  Every
extreme pair of optimal strategies of a matrix game is the solution of a
square sub-system

    sum_{i in I} x_i M[i, j] = v   (j in J),   sum_i x_i = 1
    sum_{j in J} M[i, j] y_j = v   (i in I),   sum_j y_j = 1

with |I| == |J|, x, y >= 0 and no profitable deviation outside the supports.
The game value v is unique, so `root_value_tensor`, `value_tensor` and
`expected_value_tensor` do not depend on which solver produced them; only
`solution_tensor` can differ on degenerate matrices (several equilibria).
There we enumerate supports by increasing size, which reproduces the
behaviour of reference tree.py:227-231 (its sort key is *negative* for pure
strategies, so pure equilibria are taken first).

This is host-side, one-off tree construction logic - not on the hot path.
"""

from functools import lru_cache
from itertools import combinations

import numpy as np


@lru_cache(maxsize=None)
def _supports(rows: int, cols: int):
    """All (I, J) support pairs with |I| == |J|, smallest supports first."""
    out = []
    for k in range(1, min(rows, cols) + 1):
        for I in combinations(range(rows), k):
            for J in combinations(range(cols), k):
                out.append((np.array(I), np.array(J)))
    return out


def _solve_support(M, I, J):
    k = len(I)
    sub = M[np.ix_(I, J)]
    # unknowns (x_I, v):  sub^T x - v = 0 ; sum x = 1
    lhs = np.zeros((k + 1, k + 1))
    rhs = np.zeros(k + 1)
    rhs[k] = 1.0
    lhs[:k, :k] = sub.T
    lhs[:k, k] = -1.0
    lhs[k, :k] = 1.0
    x = np.linalg.solve(lhs, rhs)
    lhs[:k, :k] = sub
    y = np.linalg.solve(lhs, rhs)
    return x[:k], x[k], y[:k], y[k]


def solve_zero_sum(M: np.ndarray, tol: float = 1e-9):
    """
    M: (rows, cols) payoff matrix of the row (maximising) player.
    Returns (x, y, value): row strategy, column strategy, game value (float64).
    """
    M = np.asarray(M, dtype=np.float64)
    rows, cols = M.shape
    scale = max(1.0, float(np.abs(M).max()))
    eps = tol * scale
    for I, J in _supports(rows, cols):
        if len(I) == 1:
            i, j = I[0], J[0]
            v = M[i, j]
            # saddle point: max of its column, min of its row
            if M[:, j].max() <= v + eps and M[i, :].min() >= v - eps:
                x = np.zeros(rows)
                y = np.zeros(cols)
                x[i] = 1.0
                y[j] = 1.0
                return x, y, float(v)
            continue
        try:
            xs, vx, ys, vy = _solve_support(M, I, J)
        except np.linalg.LinAlgError:
            continue
        if xs.min() < -tol or ys.min() < -tol or abs(vx - vy) > eps:
            continue
        x = np.zeros(rows)
        y = np.zeros(cols)
        x[I] = np.clip(xs, 0.0, None)
        y[J] = np.clip(ys, 0.0, None)
        x /= x.sum()
        y /= y.sum()
        if (x @ M).min() >= vx - 10 * eps and (M @ y).max() <= vx + 10 * eps:
            return x, y, float(vx)
    return _solve_lp(M)


def _solve_lp(M):
    """Fallback (ill-conditioned input): the two LPs of the minimax theorem."""
    from scipy.optimize import linprog

    rows, cols = M.shape

    def one_side(P):
        # maximise v  s.t.  P^T x >= v, sum x = 1, x >= 0
        n, m = P.shape
        c = np.zeros(n + 1)
        c[n] = -1.0
        a_ub = np.hstack([-P.T, np.ones((m, 1))])
        a_eq = np.zeros((1, n + 1))
        a_eq[0, :n] = 1.0
        res = linprog(c, A_ub=a_ub, b_ub=np.zeros(m), A_eq=a_eq, b_eq=[1.0],
                      bounds=[(0, None)] * n + [(None, None)], method="highs")
        if not res.success:
            raise Exception(f"Game matrix not solved: {M.tolist()}")
        return res.x[:n], res.x[n]

    x, v = one_side(M)
    y, _ = one_side(-M.T)
    return x, y, float(v)
