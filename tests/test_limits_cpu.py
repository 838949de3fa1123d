"""Marginalised 1D limits, host logic on the CPU: Density1D.getLimits and limits.marge_limits against
limits produced by the unmodified reference (tests/golden/limits.npz, make_golden.py: run_limits).  The densities come
from the oracle (pinned to the reference at 1e-10), the order statistics from its exact weighted quantiles."""
import os

import numpy as np
import pytest

from cases import input_digest
from helpers import GOLDEN, load_case, make_oracle

TAGS = {0: "two", 1: ">", 2: "<", 3: "none"}


@pytest.mark.parametrize("name", ["mix3", "bounded", "likes"])
def test_marge_limits_match_reference(name):
    from getdist_b200.densities import Density1D
    from getdist_b200.limits import limit_fractions, marge_limits
    from oracle.getdist_oracle import weighted_quantiles

    g = np.load(os.path.join(GOLDEN, "limits.npz"))
    case, _ = load_case(name)
    assert str(g[name + "/digest"]) == input_digest(case)
    o = make_oracle(case)
    contours, mft = g[name + "/contours"], g[name + "/max_frac_twotail"]
    keys = limit_fractions(contours)
    for j in range(o.n):
        d = o.density_1d(j)
        par = o.pars[j]
        dens = Density1D(d.x, d.P.copy(), view_ranges=d.view_ranges)
        fr = np.array([(1 - lf) if up else lf for lf, up in keys])
        table = dict(zip(keys, weighted_quantiles(o.samples[:, j], o.weights, fr)))
        lims = marge_limits(dens, par, contours, mft, lambda lf, up: table[(lf, up)])
        ref, tags = g["%s/%d/limits" % (name, j)], g["%s/%d/tags" % (name, j)]
        assert [l.limitTag() for l in lims] == [TAGS[int(t)] for t in tags], (name, j)
        got = np.array([[l.lower, l.upper] for l in lims], dtype=np.float64)
        np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9 * par.err, err_msg=str((name, j)))


def test_get_limits_scalar_and_list():
    from getdist_b200.densities import Density1D

    x = np.linspace(-5, 5, 201)
    d = Density1D(x, np.exp(-x * x / 2))
    lo, hi, bot, top = d.getLimits(0.6827)
    assert not bot and not top and abs(lo + 1) < 2e-3 and abs(hi - 1) < 2e-3
    both = d.getLimits(np.array([0.6827, 0.9545]))  # an ndarray gives a list (a Python list gives the first, as in the reference)
    assert len(both) == 2 and abs(both[1][1] - 2) < 5e-3
    cut = Density1D(x[100:], np.exp(-x[100:] ** 2 / 2))  # half Gaussian: no lower limit
    assert cut.getLimits(0.68)[2] and not cut.getLimits(0.68)[3]
