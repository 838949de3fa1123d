"""Statistics pass on its own (k_stats_fused + k_stats_merge): python tools/stats_probe.py [N] [P] [nchains]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from getdist_b200 import _abi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(3)
X = rng.standard_normal((N, P)) * 10.0 ** rng.uniform(-3, 2, P) + 100.0
w = rng.exponential(1.0, N)
ctx = _abi.Context(0)
offs = None if nch == 1 else np.linspace(0, N, nch + 1).astype(np.int64)
ctx.set_samples(X, w, offs)
m = ctx.moments()
ref_mean = np.average(X, axis=0, weights=w)
ref_cov = np.cov(X.T, aweights=w, ddof=0)
sc = np.sqrt(np.outer(np.diag(ref_cov), np.diag(ref_cov)))
out = {"N": N, "P": P, "nchains": nch, "mean_err": float(np.max(np.abs(m["means"] - ref_mean) / np.sqrt(np.diag(ref_cov)))),
       "cov_err": float(np.max(np.abs(m["cov"] - ref_cov) / sc)), "min_exact": bool(np.array_equal(m["xmin"], X.min(0))),
       "max_exact": bool(np.array_equal(m["xmax"], X.max(0)))}
ctx.set_kernel_timing(True)
t0 = time.perf_counter()
for _ in range(5):
    ctx.moments_recompute()
out["wall_ms"] = (time.perf_counter() - t0) / 5 * 1e3
ks = ctx.kernel_stats()["k_stats_fused"]
out["kernel_ms"] = ks["ms"] / 5
out["tflops"] = ks["flops"] / (ks["ms"] * 1e-3) / 1e12
out["hbm_gbs"] = ks["bytes"] / (ks["ms"] * 1e-3) / 1e9
print(json.dumps(out))
