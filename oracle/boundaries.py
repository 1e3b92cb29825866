"""Exercise boundaries with the reference's `exerciseBoundary(self, t)` signature (Amer_GBM :1495-1499) that are NOT
of the demo notebook's square-root family - used by oracle/make_golden.py to run the reference and by the tests to
run the same callables through `multilevel_mc_sdes_b200.Amer_GBM` (host-tabulated boundary).  Test infrastructure."""
import math


def linear_boundary(self, t):
    return self.K * (0.75 + 0.2 * t / self.T)


def power_boundary(self, t):
    return max(self.K * (1.0 - 0.6 * math.pow(self.T - t, 0.4)), 0.0)


BOUNDARIES = {"linear": linear_boundary, "power": power_boundary}
