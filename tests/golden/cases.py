"""Case list shared by make_golden.py (generation, dev container only) and the tests.

Each case: name -> dict(gen=(generator, kwargs) | special, thres, nb, stop_mode, store_input).
Inputs are regenerated from the seed at test time unless ``store_input`` (generators that go
through LAPACK, whose low-order bits may differ between CPUs) — those are saved in the .npz.
"""
import numpy as np

from qrdm_b200 import generators as g

DEFAULT = dict(thres=(0.9, 0.15), nb=64, stop_mode=0, store_input=False, exact=False)


def _nan_case():
    A = g.gaussian(200, 200, 0)
    A[100, 150] = np.nan
    return A


def _inf_case():
    A = g.gaussian(40, 30, 0)
    A[3, 4] = np.inf
    return A


def _rank1():
    u = g.gaussian(60, 1, 5)
    v = g.gaussian(1, 40, 6)
    return np.asfortranarray(u @ v)


CASES = {
    # SURVEY.md §8(c) seeded checks
    "gauss200": dict(make=lambda: g.gaussian(200, 200, 0)),
    "gauss500x150": dict(make=lambda: g.gaussian(500, 150, 1)),
    "gauss120x300": dict(make=lambda: g.gaussian(120, 300, 2)),
    "gauss200_d06_t05_nb16": dict(make=lambda: g.gaussian(200, 200, 0), thres=(0.6, 0.5), nb=16),
    "gauss300_nb8": dict(make=lambda: g.gaussian(300, 300, 3), nb=8),
    "gauss257x131_nb32": dict(make=lambda: g.gaussian(257, 131, 4), nb=32),
    "gauss64x1": dict(make=lambda: g.gaussian(64, 1, 0)),
    "gauss1x50": dict(make=lambda: g.gaussian(1, 50, 0)),
    "gauss2x2": dict(make=lambda: g.gaussian(2, 2, 7)),
    "kahan96": dict(make=lambda: g.kahan(96), exact=True),
    "kahan96_perturbed": dict(make=lambda: g.kahan(96, perturb=1e3, seed=3), exact=True),
    "kahan64_theta12_nb4": dict(make=lambda: g.kahan(64, theta=1.2), nb=4, exact=True),
    "zeros16": dict(make=lambda: np.zeros((16, 16), order="F"), exact=True),
    "eye16": dict(make=lambda: np.asfortranarray(np.eye(16)), exact=True),
    "rank1_60x40": dict(make=_rank1),
    # the reference's only test input family (test.ipynb cell 3); noise tail graded by invariants
    "graded128": dict(make=lambda: g.graded(128, seed=0), store_input=True),
    "graded128_stop1": dict(make=lambda: g.graded(128, seed=0), stop_mode=1, store_input=True),
    "graded128_stop2": dict(make=lambda: g.graded(128, seed=0), stop_mode=2, store_input=True),
    "graded128_stop3": dict(make=lambda: g.graded(128, seed=0), stop_mode=3, thres=(0.9, 0.15, 1e-6),
                            store_input=True),
    "graded96x160": dict(make=lambda: g.graded(160, seed=4, m=96), store_input=True),
    # error paths
    "nan_in_trailing": dict(make=_nan_case),
    "inf_in_panel": dict(make=_inf_case),
}

for _k, _v in CASES.items():
    for _d, _dv in DEFAULT.items():
        _v.setdefault(_d, _dv)
