"""Stand-in for `bezier` (absent): Curve(nodes, degree).evaluate_multi(s) = Bernstein evaluation.
Needed by reference isegm/engine/trainer.py:7,1139-1141 (prompt simulators used as host fixtures)."""
import numpy as np
from math import comb


class Curve:
    def __init__(self, nodes, degree):
        self.nodes = np.asarray(nodes, dtype=np.float64)
        self.degree = degree

    def evaluate_multi(self, s):
        s = np.asarray(s, dtype=np.float64)
        n = self.degree
        out = np.zeros((self.nodes.shape[0], s.shape[0]))
        for i in range(n + 1):
            out += np.outer(self.nodes[:, i], comb(n, i) * (s ** i) * ((1 - s) ** (n - i)))
        return out
