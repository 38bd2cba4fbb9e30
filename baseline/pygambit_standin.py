"""
Stand-in for the third-party `pygambit==16.0.2` (reference requirements.txt:3), which is neither vendored with the
reference nor installable offline, so that the UNMODIFIED reference under baseline/_ref imports (tree.py:5) and, if
asked to, generates trees.  Independent of this repository's package: it implements only what reference
environment/tree.py:199-234 touches - `Decimal`, `Game.from_arrays(A, -A)`, `nash.enummixed_solve(g, rational=False)`,
`nash.lcp_solve(...)` returning [[*x, *y]] - with the two linear programs of a zero-sum matrix game (scipy `highs`).
Installed as baseline/_ref/_standin/pygambit.py by baseline/install_reference.py; used by `bench.py --impl reference`.
"""
from decimal import Decimal  # noqa: F401  (reference uses pygambit.Decimal)

import numpy as np
from scipy.optimize import linprog


def _maximin(a):
    """Mixed strategy x maximising min_j (x^T a)_j."""
    m, n = a.shape
    # variables (x_0..x_{m-1}, v): minimise -v  s.t.  v - (x^T a)_j <= 0,  sum x = 1,  x >= 0
    c = np.zeros(m + 1)
    c[-1] = -1.0
    a_ub = np.hstack([-a.T, np.ones((n, 1))])
    res = linprog(c, A_ub=a_ub, b_ub=np.zeros(n), A_eq=np.array([[1.0] * m + [0.0]]), b_eq=[1.0],
                  bounds=[(0, None)] * m + [(None, None)], method="highs")
    x = np.clip(res.x[:m], 0.0, None)
    return x / x.sum()


class Game:
    def __init__(self, a):
        self.a = np.array(a, dtype=np.float64)

    @classmethod
    def from_arrays(cls, a, b):
        return cls(a)


class nash:
    @staticmethod
    def enummixed_solve(g, rational=False):
        return [[*_maximin(g.a), *_maximin(-g.a.T)]]

    lcp_solve = enummixed_solve
