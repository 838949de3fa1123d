"""INTEGRATION.md shows the subclass a reference maintainer would add.  Where the unmodified reference is importable
(the build container: /root/reference; not on the GPU box), the stub is extracted from the document and executed
against the test-double context: the reference's own MCSamples object then serves get1DDensityGridData, getMeans and
getCov through this package's host mirror, and the results must agree with the reference's own methods."""
import os
import re
import sys

import numpy as np
import pytest

from helpers import load_case
from test_host_mirror_cpu import fake_ctx  # noqa: F401
from test_hostsim import hs  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "getdist")), reason="the reference tree is not present on this machine")
def test_integration_stub_runs_against_the_reference(fake_ctx):  # noqa: F811
    if REF not in sys.path:
        sys.path.insert(0, REF)
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n# getdist/gpu_backend.py.*?\n(.*?)```", text, re.S).group(1)
    ns = {}
    exec(compile(code, "INTEGRATION.md:gpu_backend", "exec"), ns)
    Stub = ns["MCSamples"]
    case, _ = load_case("bounded")
    kw = dict(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"], sampler="uncorrelated")
    import getdist

    ref = getdist.MCSamples(**kw)
    mc = Stub(**kw)
    np.testing.assert_allclose(mc.getMeans(), ref.getMeans(), rtol=1e-12)
    np.testing.assert_allclose(mc.getCov(), ref.getCov(), rtol=1e-10, atol=1e-14)
    for name in case["names"][:3]:
        a, b = mc.get1DDensityGridData(name), ref.get1DDensityGridData(name)
        assert type(a) is type(b)  # the reference's own Density1D
        assert np.max(np.abs(a.P - b.P)) < 1e-7
        np.testing.assert_allclose(a.view_ranges, b.view_ranges, rtol=1e-12)
