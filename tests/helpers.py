"""Shared helpers for the parity tests: load a seeded case + its golden vectors, build the
oracle object for it."""
import os

import numpy as np

from cases import CASES, input_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_cache = {}


def load_case(name):
    if name not in _cache:
        case = CASES[name]()
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        assert str(g["digest"]) == input_digest(case), "seeded inputs no longer match the golden fixture"
        _cache[name] = (case, g)
    return _cache[name]


def make_oracle(case):
    from oracle.getdist_oracle import OracleSamples

    return OracleSamples(case["samples"], case["weights"], names=case["names"], ranges=case["ranges"],
                         sampler=case.get("sampler", "uncorrelated"), settings=case["settings"],
                         loglikes=case.get("loglikes") if case.get("meanlikes") else None)
