"""
Stand-in for the third-party `pygambit==16.0.2` (reference requirements.txt:3),
which is neither vendored with the reference nor installable offline.  It
implements only what reference environment/tree.py:199-234 touches, and is used
ONLY by tests/golden/make_golden.py to import and run the unmodified reference
from /root/reference when generating golden vectors.  Never imported by the
product or by tests at run time.
"""
import os
import sys
from decimal import Decimal  # noqa: F401  (reference uses pygambit.Decimal)

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "r-nad_b200", "util"))
from matrix_game import solve_zero_sum  # noqa: E402

sys.path.pop(0)


class Game:
    def __init__(self, a):
        self.a = np.array(a, dtype=np.float64)

    @classmethod
    def from_arrays(cls, a, b):
        return cls(a)


class nash:
    @staticmethod
    def enummixed_solve(g, rational=False):
        x, y, _ = solve_zero_sum(g.a)
        return [[*x, *y]]

    lcp_solve = enummixed_solve
