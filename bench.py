#!/usr/bin/env python
"""bench.py -- headline benchmark: 1D+2D densities/sec for the full triangle of N=1e7 weighted samples x P=64
parameters (BASELINE.json configs[1], "C2": correlated Gaussian, fine_bins=2048 / 256^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n N --p P (debug sizes)]

A "step" is one pass of the hot path over the resident sample store: exact weighted quantiles for all P
parameters, P 1D densities and P(P-1)/2 2D densities (histograms, bandwidths, convolutions, corrections,
normalisation).  Means/covariance are computed at upload (as the reference does at construction) and are part of
the end-to-end figure only.

  value : densities/s, samples resident in HBM, results left on the device; CUDA events on the library stream.
  e2e   : densities/s through the public API (MCSamples(samples=...) + prefetch_triangle()) from PINNED HOST
          buffers: H2D upload, moments, quantiles, densities, D2H of every grid inside the timed region.
  roofline    : the dominant kernel by CUDA-event time (2D histogram pass): the bytes its design must move (3 B per
                pair-sample + the byte pre-binning) over its event duration, against MEASURED_PEAKS.json hbm_gbs; the
                reference-shaped N*24 B per pair figure is kept beside it as `standalone_equiv`.
  hist1d      : the north-star "histogram-pass HBM GB/s": N*(P+1)*8 B over the 1D sweep's event duration.
  cpu_baseline: the oracle (numpy/scipy restatement of the reference, pinned to it by goldens) timed on the host
                on a bounded sample of the same workload; also used as a parity check of the GPU result.

--impl reference times that same CPU path alone (rank 0 only) and prints the reference-arm line.
Multi-GPU (torchrun): every rank holds the full sample store, the list of densities is partitioned across ranks,
result grids are all-gathered with NCCL (torch.distributed); total work is fixed => "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "densities/s"
SETTINGS = {"fine_bins": 2048, "fine_bins_2D": 256}
PARITY_PAIRS = 64  # full-size parity leg: all P 1D densities + this many seeded pairs against the CPU implementation

WORKLOADS = {
    # BASELINE.json configs[1] (the headline), its rho = 0.95 variant (variable grids 384..960 through the
    # shear / scaled-grid branches) and configs[4] (P = 256 triangle)
    "c2": dict(rho=0.85, p=64, n=10_000_000, tag="C2 correlated Gaussian (AR1 rho=0.85)"),
    "c2r95": dict(rho=0.95, p=64, n=10_000_000, tag="C2 variant: correlated Gaussian (AR1 rho=0.95), variable 2D grids"),
    "c5": dict(rho=0.85, p=256, n=10_000_000, tag="C5 correlated Gaussian (AR1 rho=0.85), P=256 triangle"),
}


def metric_name(args):
    if args.workload == "c2" and args.n == 10_000_000 and args.p == 64:
        return "1D+2D densities/sec (full triangle, N=1e7 x P=64, fine_bins=2048 / 256^2)"
    return "1D+2D densities/sec (full triangle, N=%d x P=%d, fine_bins=%d / %d^2)" % (args.n, args.p, SETTINGS["fine_bins"], SETTINGS["fine_bins_2D"])


def workload_name(args):
    N, P = args.n, args.p
    return ("%s N=%d P=%d, Exp(1) weights, fine_bins=%d, fine_bins_2D=%d, full triangle: %d 1D + %d 2D densities" % (
        WORKLOADS[args.workload]["tag"], N, P, SETTINGS["fine_bins"], SETTINGS["fine_bins_2D"], P, P * (P - 1) // 2))


def gen_workload(args, out_X=None, out_w=None):
    return gen_c2(args.n, args.p, out_X, out_w, rho=WORKLOADS[args.workload]["rho"])


def gen_c2(N, P, out_X=None, out_w=None, rho=0.85, seed=1234):
    """AR(1) correlated Gaussian, scales 10^U(-4,2), offsets sigma*U(-150,150), Exp(1) weights."""
    rng = np.random.default_rng(seed)
    R = rho ** np.abs(np.subtract.outer(np.arange(P), np.arange(P)))
    L = np.linalg.cholesky(R)
    sig = 10.0 ** rng.uniform(-4, 2, P)
    mu = sig * rng.uniform(-150, 150, P)
    X = np.empty((N, P)) if out_X is None else out_X
    chunk = 1 << 19
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        Z = rng.standard_normal((r1 - r0, P))
        np.multiply(Z.dot(L.T), sig, out=X[r0:r1])
        X[r0:r1] += mu
    w = np.empty(N) if out_w is None else out_w
    w[:] = np.random.default_rng(seed + 1).exponential(1.0, N)
    return X, w


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled during the timed region (B200_PROFILING.md): NVML queries
    from a thread of this process every 50 ms.  (A polling `nvidia-smi -lms` child was measured to stall kernel launches
    for tens of milliseconds per poll on some boxes and inflated single steps by up to 2x; it remains the fallback when
    the NVML bindings are missing.)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []  # (sm_mhz, max_mhz, reason bits)
        self.nvml = None
        self.first = 0
        self._stop = False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((mhz, self.max_mhz, bits))
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """start of the timed region: only samples from here on are reported"""
        self.first = len(self.samples) if self.nvml else len(self.lines)

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml:
            self._stop = True
            nv = self.nvml
            masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                     nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            sel = self.samples[self.first:]
            reasons = sorted({nm for (_, _, bits) in sel for nm, m in zip(names, masks) if bits & m})
            return {"sm_mhz": float(np.median([x[0] for x in sel])) if sel else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(sel), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation of the path on the host cores.  The UNMODIFIED reference (getdist 1.7.7,
# offline install under baseline/_ref: it travels to the GPU box with the snapshot) when it is importable, else the
# oracle port (numpy/scipy restatement pinned to the reference by the goldens).  Used three ways, never on the product
# path: the `--impl reference` arm, the cpu_baseline figure, and the checker of the full-size parity leg.
# ---------------------------------------------------------------------------------------------------------------
_CPU = {}


def cpu_backend_kind():
    if "kind" not in _CPU:
        ref = os.path.join(ROOT, "baseline", "_ref")
        kind = "port"
        if os.path.isdir(os.path.join(ref, "getdist")):
            try:
                if ref not in sys.path:
                    sys.path.insert(0, ref)
                import getdist  # noqa: F401

                kind = "reference"
            except Exception:
                kind = "port"
        _CPU["kind"] = kind
    return _CPU["kind"]


class _CpuObject:
    """one analysis object of the CPU implementation on a few columns; .d1(k) / .d2(k, l) return density grids"""

    def __init__(self, cols, w, names):
        import contextlib
        import io

        X = np.ascontiguousarray(np.stack(cols, axis=1))
        if cpu_backend_kind() == "reference":
            from getdist import MCSamples as RefSamples

            with contextlib.redirect_stdout(io.StringIO()):
                self.o = RefSamples(samples=X, weights=w, names=names, sampler="uncorrelated", settings=dict(SETTINGS))
            self.d1 = lambda k: self.o.get1DDensityGridData(k).P
            self.d2 = lambda k, l: self.o.get2DDensityGridData(k, l, get_density=True).P
        else:
            from oracle.getdist_oracle import OracleSamples

            self.o = OracleSamples(X, w, names=names, sampler="uncorrelated", settings=SETTINGS)
            self.d1 = lambda k: self.o.density_1d(k).P
            self.d2 = lambda k, l: self.o.density_2d(k, l).P


def _cpu_columns():
    """(P, N) column store + weights of the worker: memory-mapped files written by the parent (spawned workers) or
    the parent's arrays inherited by fork (reference arm: no CUDA in that process)"""
    if "Xt" not in _CPU:
        _CPU["Xt"] = np.load(_CPU["xt_path"], mmap_mode="r")
        _CPU["w"] = np.load(_CPU["w_path"], mmap_mode="r")
    return _CPU["Xt"], _CPU["w"]


def _cpu_init(xt_path, w_path, kind):
    _CPU.update(xt_path=xt_path, w_path=w_path, kind=kind)
    if kind == "reference":
        ref = os.path.join(ROOT, "baseline", "_ref")
        if ref not in sys.path:
            sys.path.insert(0, ref)


def _cpu_parity_task(task):
    """('1d', j) or ('2d', jx, jy): the density grid of the CPU implementation and the seconds it took (the object
    construction -- moments of the columns -- included, as the GPU step includes its quantile/range stage)"""
    Xt, w = _cpu_columns()
    t0 = time.perf_counter()
    if task[0] == "1d":
        o = _CpuObject([np.asarray(Xt[task[1]])], np.asarray(w), ["p%d" % task[1]])
        P = o.d1(0)
    else:
        o = _CpuObject([np.asarray(Xt[task[1]]), np.asarray(Xt[task[2]])], np.asarray(w), ["p%d" % task[1], "p%d" % task[2]])
        P = o.d2(0, 1)
    return task, np.asarray(P), time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _ref_task(k):
    """one worker task of the reference arm: a 1D and a 2D (shear branch) density on columns (2k, 2k+1)"""
    obj = _CPU.get("mine")
    if obj is None:
        # one object per WORKER PROCESS (its own pair of columns), whichever tasks the pool hands it
        import multiprocessing as mp

        ident = getattr(mp.current_process(), "_identity", None) or (k + 1,)
        wk = ident[0] - 1
        X, w = _CPU["X"], _CPU["w"]
        cols = [(2 * wk) % X.shape[1], (2 * wk + 1) % X.shape[1]]
        obj = _CPU["mine"] = _CpuObject([X[:, c] for c in cols], w, ["p%d" % c for c in cols])
    obj.d1(0)
    obj.d2(0, 1)
    return 2


def run_reference(args):
    """CPU arm: the reference's own implementation (see cpu_backend_kind) on the box's host cores.  The path is
    single-threaded per density and densities are independent, so every core runs its own (1D + 2D) pair of densities."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    N, P = args.n, args.p
    kind = cpu_backend_kind()
    X, w = gen_workload(args)
    ncore = host_cores()
    nw = max(1, min(ncore, P // 2, 32))
    _CPU.update(X=X, w=w)
    pool = mp.get_context("fork").Pool(nw) if nw > 1 else None

    def step():
        if pool is None:
            return _ref_task(0)
        return sum(pool.map(_ref_task, range(nw), chunksize=1))

    for _ in range(max(1, args.warmup)):  # first call builds the per-worker objects (construction is not timed)
        step()
    t0 = time.perf_counter()
    nd = 0
    for _ in range(args.steps):
        nd += step()
    dt = time.perf_counter() - t0
    if pool is not None:
        pool.close()
    val = nd / dt
    impl = "unmodified getdist 1.7.7 (baseline/_ref)" if kind == "reference" else "oracle port of the reference"
    sample = ("%s; per step: %d worker processes x (1 x get1DDensityGridData + 1 x get2DDensityGridData, shear branch) on distinct "
              "column pairs at full N=%d, fine_bins %d/%d^2" % (impl, nw, N, SETTINGS["fine_bins"], SETTINGS["fine_bins_2D"]))
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nw, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_parity(X, w, tasks, tmpdir=None):
    """The CPU implementation on `tasks` (('1d', j) / ('2d', jx, jy)) at full N, one task per worker process on all host
    cores.  The workers are spawned (no CUDA state inherited) and map the columns from two .npy files.  Returns
    ({task: grid}, wall seconds, cores used, summed task seconds)."""
    import multiprocessing as mp
    import shutil
    import tempfile

    need = X.shape[0] * X.shape[1] * 8 + X.shape[0] * 8
    base = None
    for cand in ([tmpdir] if tmpdir else []) + ["/dev/shm", tempfile.gettempdir()]:
        try:
            if shutil.disk_usage(cand).free > need + (1 << 30):
                base = cand
                break
        except Exception:
            continue
    if base is None:
        return None
    d = tempfile.mkdtemp(prefix="gdk_parity_", dir=base)
    try:
        xt_path, w_path = os.path.join(d, "xt.npy"), os.path.join(d, "w.npy")
        cols = sorted({c for t in tasks for c in t[1:]})
        xt = np.lib.format.open_memmap(xt_path, mode="w+", dtype=np.float64, shape=(X.shape[1], X.shape[0]))
        for c in cols:
            xt[c] = X[:, c]
        xt.flush()
        del xt
        np.save(w_path, np.asarray(w))
        nw = max(1, min(host_cores(), len(tasks), 64))
        t0 = time.perf_counter()
        with mp.get_context("spawn").Pool(nw, initializer=_cpu_init, initargs=(xt_path, w_path, cpu_backend_kind())) as pool:
            t1 = time.perf_counter()  # interpreter start-up of the workers is not CPU work of the path
            res = pool.map(_cpu_parity_task, tasks, chunksize=1)
            wall = time.perf_counter() - t1
        return {t: P for t, P, _ in res}, wall, nw, float(sum(dt for _, _, dt in res)), t1 - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    pg = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)

    from getdist_b200 import MCSamples, _abi
    from getdist_b200.parallel import PeerGroup, bind_to_gpu_numa, prefetch_triangle_group

    numa_cpus = None
    if world > 1:
        pg = PeerGroup(dist, rank, world, use_p2p=not args.nccl_gather)
        if not args.no_numa:
            numa_cpus = bind_to_gpu_numa(dev)  # before the pinned input buffers are allocated

    N, P = args.n, args.p
    # pinned host inputs (e2e path copies from here every step)
    X, xh = _abi.pinned_empty((N, P))
    w, wh = _abi.pinned_empty((N,))
    t0 = time.perf_counter()
    gen_workload(args, X, w)
    t_gen = time.perf_counter() - t0
    names = ["p%d" % i for i in range(P)]

    mc = MCSamples(samples=X, weights=w, names=names, sampler="uncorrelated", settings=SETTINGS, device=dev, process_group=pg)
    measured = mc._ctx.measure_peaks() if rank == 0 else None
    idx, pairs = mc.triangle_pairs()
    F, G = SETTINGS["fine_bins"], SETTINGS["fine_bins_2D"]
    ndens_total = len(idx) + len(pairs)
    d1 = d2 = None
    if world == 1:
        mc._ensure_param_ranges(idx)
        fb = mc._specs_2d_batch(pairs, {})["fine_bins"].astype(np.int64)
        total2 = int((fb * fb).sum())
        d1 = torch.zeros((len(idx), F), dtype=torch.float64, device="cuda")
        d2 = torch.zeros((total2,), dtype=torch.float64, device="cuda")
        grid_sizes = sorted({int(g) for g in fb})
    phases = {}
    host_log = []
    group_timings = []

    def step_resident():
        """one pass of the hot path over the resident samples; the results stay on the device (on every rank)"""
        th0 = time.perf_counter()
        mc.invalidate_density_caches()
        mc._ctx.timer_start()
        if world > 1:
            prefetch_triangle_group(mc, pg, idx, to_host=False)
            ph = mc._ctx.phase_ms()
            for k in ("hist1d", "kde1d", "quantiles", "hist2d", "shear", "xform2d", "bw2d", "conv2d"):
                phases[k] = ph[k]
            group_timings.append(dict(pg.timings))
        else:
            mc._densities_1d(idx, _device_ptr=d1.data_ptr())
            ph = mc._ctx.phase_ms()
            phases["hist1d"], phases["kde1d"], phases["quantiles"] = ph["hist1d"], ph["kde1d"], ph["quantiles"]
            mc._densities_2d(pairs, _device_ptr=d2.data_ptr())  # with the contour levels, as the public call
            ph = mc._ctx.phase_ms()
            for k in ("hist2d", "shear", "xform2d", "bw2d", "conv2d"):
                phases[k] = ph[k]
        th1 = time.perf_counter()
        ms = mc._ctx.timer_stop_ms()
        wl = mc._ctx.wall_ms()
        # host wall clock of the step and of the library calls inside it: the rest is the Python planner
        host_log.append({"wall_ms": round((th1 - th0) * 1e3, 2), "lib_1d_ms": round(wl["call_1d"], 2),
                         "lib_2d_ms": round(wl["call_2d"], 2), "lib_quant_ms": round(wl["call_quantiles"], 2)})
        return ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler (one nvidia-smi process polling every 100 ms) is started BEFORE the warm-up: its start-up
    # stalls kernel launches for tens of milliseconds; samples taken before the timed region are discarded below
    clocks = ClockSampler(dev)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_resident()
    # settle: on a cold box single steps still came out 10-40 % slow right after the W warm-up steps (clock / power
    # state, first-touch of the 25 GB record scratch); up to 6 more UNTIMED steps until two in a row agree within 2 %
    extra_warm, prev = 0, None
    while extra_warm < 6:
        cur = step_resident()
        extra_warm += 1
        if world > 1:  # the steps contain collectives: every rank must run the same number of them
            if extra_warm == 2:
                break
            continue
        if prev is not None and abs(cur - prev) <= 0.02 * prev:
            break
        prev = cur
    barrier()
    if rank == 0:
        clocks.mark()
    l0 = mc._ctx.launch_count()
    total_ms = 0.0
    step_ms = []
    mc._ctx.set_kernel_timing(True)  # event pairs around the tagged kernels of the timed steps (roofline figures)
    for _ in range(args.steps):
        step_ms.append(step_resident())
        total_ms += step_ms[-1]
    barrier()
    clk = clocks.stop() if rank == 0 else None
    kstats = mc._ctx.kernel_stats()
    mc._ctx.set_kernel_timing(False)
    launches = mc._ctx.launch_count() - l0
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = ndens_total / (ms_per_step * 1e-3)

    # ---------------- the statistics pass on its own (fused one-sweep moments; rides behind the upload otherwise) ----
    stats_line = None
    if world == 1:
        mc._ctx.set_kernel_timing(True)
        mc._ctx.timer_start()
        for _ in range(3):
            mc._ctx.moments_recompute()
        pass_ms = mc._ctx.timer_stop_ms() / 3.0
        ks = mc._ctx.kernel_stats().get("k_stats_fused")
        mc._ctx.set_kernel_timing(False)
        if ks:
            stats_line = {"ms": ks["ms"] / 3.0, "bytes": ks["bytes"] / 3.0, "flops": ks["flops"] / 3.0, "pass_ms": pass_ms}

    # ---------------- end to end through the public API, pinned host -> host results ----------------
    e2e_steps = max(1, min(args.steps, 3))
    checksum = 0.0
    e2e_parts = {}
    holder = {}

    def step_e2e():
        """what a user runs: new samples into the object (H2D + statistics), then every density of the triangle plot
        with its contour levels, grids back on the host"""
        t0 = time.perf_counter()
        m = holder.get("m")
        if m is None:
            m = holder["m"] = MCSamples(samples=X, weights=w, names=names, sampler="uncorrelated", settings=SETTINGS, device=dev,
                                        process_group=pg)
        else:
            m.setSamples(X, w)          # new samples -> device copy invalidated (chains.py:262-308, 310-323)
            m.updateBaseStatistics()    # H2D upload of the pinned host arrays; the fused moments ride behind the chunks
        t1 = time.perf_counter()
        e2e_parts["upload_ms_events"] = m._ctx.phase_ms()["upload"]
        a, b = m.prefetch_triangle() if world == 1 else m.prefetch_triangle(root=0)  # N ranks: the grids end on rank 0's host
        t2 = time.perf_counter()
        s = float(a[0].P[F // 2]) + float(b[0].P[b[0].P.shape[0] // 2, b[0].P.shape[1] // 2]) if a else 0.0
        e2e_parts.update(upload_and_moments_s=t1 - t0, prefetch_triangle_s=t2 - t1)
        e2e_parts["phases_ms"] = {k: round(v, 2) for k, v in m._ctx.phase_ms().items() if v and v > 0}
        if getattr(m, "last_prefetch_ms", None):
            e2e_parts["prefetch_ms"] = {k: (round(v, 2) if v is not None else None) for k, v in m.last_prefetch_ms.items()}
        if world > 1:
            e2e_parts["group"] = {k: round(v, 2) for k, v in pg.timings.items()}
            e2e_parts["upload_group"] = {k: round(v, 2) for k, v in getattr(pg, "upload_timings", {}).items()}
        holder["d"] = (a, b)
        return s, ndens_total

    if world > 1:
        dist.barrier()
    mc._ctx.close()
    del mc  # free the resident copy before timing fresh uploads
    import gc

    gc.collect()
    e2e_s = float("nan")
    if not args.no_e2e:
        step_e2e()  # warm: context, device buffers
        step_e2e()  # warm: the second pooled pinned result buffer (the previous step's grids are still referenced)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            s, _nd = step_e2e()
            checksum += s
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_diag = None
    if not args.no_e2e and world == 1:
        # one more (untimed) public call with the per-kernel event timers on: what the e2e path adds to the resident step
        m = holder["m"]
        m.invalidate_density_caches()
        m._ctx.set_kernel_timing(True)
        m._ctx.timer_start()
        m.prefetch_triangle()
        dev_ms = m._ctx.timer_stop_ms()
        ks = m._ctx.kernel_stats()
        m._ctx.set_kernel_timing(False)
        e2e_diag = {"prefetch_device_ms": round(dev_ms, 2), "phases_ms": {k: round(v, 2) for k, v in m._ctx.phase_ms().items() if v and v > 0},
                    "k_contours2d_ms": round(ks.get("k_contours2d", {}).get("ms", 0.0), 3), "host_ms": dict(m.last_prefetch_ms)}
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_val = ndens_total / e2e_s
    rows_mine = N if world == 1 or not pg.p2p else (pg.row_range(N)[1] - pg.row_range(N)[0])
    h2d = rows_mine * P * 8 + N * 8
    d2h = None
    if "d" in holder and rank == 0:
        d2h = int(sum(d.P.size for d in holder["d"][0]) + sum(d.P.size for d in holder["d"][1])) * 8

    # ---------------- multi-GPU: the gathered grids against a single-GPU computation (rank 0) ----------------
    rng = np.random.default_rng(99)
    sel = sorted(rng.choice(len(pairs), size=min(PARITY_PAIRS, len(pairs)), replace=False).tolist())
    multi = None
    if world > 1 and "d" in holder and not args.no_parity:
        if rank == 0:
            m1 = MCSamples(samples=X, weights=w, names=names, sampler="uncorrelated", settings=SETTINGS, device=dev)
            s1 = m1._densities_1d(idx)
            s2 = m1._densities_2d([pairs[k] for k in sel])
            g1, g2 = holder["d"]
            e1 = max(float(np.max(np.abs(g1[i].P - s1[i].P))) for i in range(len(idx)))
            e2 = max(float(np.max(np.abs(g2[k].P - s2[n].P))) for n, k in enumerate(sel))
            lv = max(float(np.max(np.abs(np.array(g2[k]._gdk["levels"][1]) - np.array(s2[n]._gdk["levels"][1])))) for n, k in enumerate(sel))
            multi = {"n_1d": len(idx), "n_2d": len(sel), "max_abs_diff_1d": e1, "max_abs_diff_2d": e2, "max_abs_diff_levels": lv,
                     "bit_identical": bool(e1 == 0.0 and e2 == 0.0),
                     "moments_bit_identical": bool(np.array_equal(m1.getCov(), holder["m"].getCov()) and np.array_equal(m1.getMeans(), holder["m"].getMeans())),
                     "what": "every 1D grid and %d seeded 2D grids gathered over %d ranks vs the same densities from a single-GPU object" % (len(sel), world)}
            m1._ctx.close()
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- rooflines: every tagged kernel against the peak that bounds it ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp64_peak = float(measured["fp64_tflops"])
    atom_peak = float(measured["smem_updates_per_s"])
    # Per-kernel figures: the library brackets the tagged kernels of the timed steps with CUDA-event pairs on its
    # stream and counts the ALGORITHMIC bytes / flops / histogram updates of every launch (gdk_kernel_stat; DESIGN.md s4
    # states the per-unit figures).  achieved = work per launch / average launch time.  `traffic` = dram bytes per launch
    # from the ncu --set full capture of exactly this workload (profiles/); other sizes have no capture -> null.
    c2 = (args.workload == "c2" and N == 10_000_000 and P == 64 and world == 1)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r4u_ncu_full_summary.csv (ncu --set full, this workload)
    ncu_traffic = {"k_bin8c": 5.96e9, "k_bucket_records": 13.49e9, "k_hist2d_records": 13.30e9, "k_hist1d_tma": 5.83e9,
                   "k_qhist": 6.14e9, "k_shear_minmax_tma": 9.45e9, "k_shear_hist_w": 8.70e9,  # average over the step's three launches (25.04 + 0.82 + 0.24 GB)
                    "k_bw2d": 27.37e9,
                   "k_xform_rows": 4.17e9, "k_xform_cols": 5.26e9, "k_conv2d<0>": 0.36e9, "k_conv2d<1>": 0.79e9, "k_contours2d": 1.49e9}
    from getdist_b200.parallel import partition_triangle

    n_my_pairs = len(partition_triangle(idx, pairs, rank, world)[1])
    # histogram updates (one 64-bit fixed-point shared-memory add each): every pair of this rank bins all N rows once
    # (k_hist2d_records); the sheared re-binning does the same for the shear-branch pairs (3 "flops" per pair-sample)
    updates = {"k_hist2d_records": lambda st: float(N) * n_my_pairs * args.steps, "k_shear_hist_w": lambda st: st["flops"] / 3.0}
    fp64_kernels = ("k_conv2d<0>", "k_conv2d<1>", "k_stats_fused", "k_xform_rows", "k_xform_cols")
    kernels = []
    for nm, st in kstats.items():
        if st["ms"] <= 0:
            continue
        per_launch_ms = st["ms"] / st["launches"]
        ent = {"kernel": nm, "launches_per_step": st["launches"] / args.steps, "avg_launch_ms": round(per_launch_ms, 4),
               "ms_per_step": round(st["ms"] / args.steps, 4)}
        gbs = st["bytes"] / (st["ms"] * 1e-3) / 1e9
        ent.update({"hbm_GBs": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "algorithmic_bytes_per_launch": st["bytes"] / st["launches"],
                    "traffic": ncu_traffic.get(nm) if c2 else None})
        if c2 and nm in ncu_traffic and st["bytes"] > 0:
            ent["traffic_over_algorithmic"] = round(ncu_traffic[nm] / (st["bytes"] / st["launches"]), 2)
        if nm in fp64_kernels and st["flops"] > 0:
            tf = st["flops"] / (st["ms"] * 1e-3) / 1e12
            ent.update({"bound": "fp64", "achieved": round(tf, 3), "peak": round(fp64_peak, 2), "unit": "TFLOP/s", "frac": round(tf / fp64_peak, 4),
                        "algorithmic_flops_per_launch": st["flops"] / st["launches"]})
        elif nm in updates:
            ups = updates[nm](st) / (st["ms"] * 1e-3)
            ent.update({"bound": "smem_atomics", "achieved": round(ups / 1e9, 2), "peak": round(atom_peak / 1e9, 2), "unit": "Gupdates/s",
                        "frac": round(ups / atom_peak, 4), "updates_per_launch": updates[nm](st) / st["launches"]})
        else:
            ent.update({"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4)})
        kernels.append(ent)
    kernels.sort(key=lambda e: -e["ms_per_step"])
    dom = max(phases, key=lambda k: phases.get(k, 0) or 0) if phases else None
    roof = None
    hbm_side = [e for e in kernels if e["bound"] in ("hbm", "smem_atomics")]
    if hbm_side:
        top = hbm_side[0]  # the data-path kernel with the largest share of the step
        gbs = top["hbm_GBs"]
        roof = {"kernel": top["kernel"], "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                "traffic": top["traffic"], "peak_source": peak_src, "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
                "avg_launch_ms": top["avg_launch_ms"], "launches_per_step": top["launches_per_step"],
                "share_of_step": round(top["ms_per_step"] / ms_per_step, 4), "dominant_phase_by_time": dom,
                "limiter": ({"bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"]}
                            if top["bound"] != "hbm" else None),
                "note": "dominant data-path kernel by CUDA-event time; `kernels` lists every tagged kernel with the peak that bounds it "
                        "(measured in-tree at bench start: `measured_peaks`)"}
    k1 = kstats.get("k_hist1d_tma")
    hist1d = None
    if k1 and k1["ms"] > 0:
        a1 = k1["bytes"] / (k1["ms"] * 1e-3) / 1e9
        hist1d = {"kernel": "k_hist1d_tma", "bound": "hbm", "achieved": round(a1, 1), "peak": peak, "unit": "GB/s", "frac": round(a1 / peak, 4),
                  "algorithmic_bytes": k1["bytes"] / k1["launches"], "kernel_ms": k1["ms"] / k1["launches"], "traffic": 5.83e9 if c2 else None}
    stats_pass = None
    if stats_line:
        gbs = stats_line["bytes"] / (stats_line["ms"] * 1e-3) / 1e9
        tf = stats_line["flops"] / (stats_line["ms"] * 1e-3) / 1e12
        stats_pass = {"kernel": "k_stats_fused", "ms": round(stats_line["ms"], 4), "pass_ms": round(stats_line["pass_ms"], 3), "hbm_GBs": round(gbs, 1), "hbm_frac": round(gbs / peak, 4),
                      "fp64_TFLOPs": round(tf, 3), "fp64_frac": round(tf / fp64_peak, 4),
                      "note": "one sweep: sum w, means, min/max and the centred P x P block per chain; bytes = N (P+1) 8, flops = 2 N 64^2 per tile; "
                              "pass_ms = the whole statistics pass on the device clock (sweep + segment merge + record copy + host merge); "
                              "in the end-to-end path the sweep rides behind the upload chunks"}

    # ---------------- CPU implementation on a bounded sample + full-size parity against it ----------------
    cpu = None
    parity = None
    if not args.no_cpu and "d" in holder:
        g1, g2 = holder["d"]
        st2 = np.array([d._gdk["status"] for d in g2], dtype=np.int64)
        amise = np.flatnonzero(st2 & (64 | 128))
        psel = list(sel)
        if len(amise) and not any(k in set(amise.tolist()) for k in psel):
            psel[-1] = int(amise[0])  # at least one pair whose bandwidth the AMISE optimiser decided
        tasks = [("1d", j) for j in idx] + [("2d",) + tuple(pairs[k]) for k in psel]
        if args.parity_1d is not None:
            tasks = [("1d", j) for j in idx[: args.parity_1d]] + [("2d",) + tuple(pairs[k]) for k in psel[: args.parity_1d]]
        got = cpu_parity(X, w, tasks)
        if got is not None:
            grids, wall, nw, cpu_s, spawn_s = got
            e1, e2p, e2t = 0.0, 0.0, 0.0
            n_t = 0
            for tsk in tasks:
                if tsk[0] == "1d":
                    e1 = max(e1, float(np.max(np.abs(g1[idx.index(tsk[1])].P - grids[tsk]))))
                else:
                    k = pairs.index((tsk[1], tsk[2]))
                    err = float(np.max(np.abs(g2[k].P - grids[tsk])))
                    if st2[k] & (64 | 128):
                        e2t, n_t = max(e2t, err), n_t + 1
                    else:
                        e2p = max(e2p, err)
            n2 = sum(1 for tsk in tasks if tsk[0] == "2d")
            parity = {"checker": "unmodified getdist 1.7.7 (baseline/_ref)" if cpu_backend_kind() == "reference" else "oracle port",
                      "n_1d": len(tasks) - n2, "n_2d": n2, "n_2d_amise_decided": n_t,
                      "max_abs_dP_1d": e1, "max_abs_dP_2d": e2p, "max_abs_dP_2d_amise_decided": e2t if n_t else None,
                      "tolerance": 1e-6, "tolerance_amise_decided": 1e-5,
                      "pass": bool(e1 < 1e-6 and e2p < 1e-6 and e2t < 1e-5),
                      "triangle_status": {"pairs": len(pairs), "amise_corr_accepted": int(np.count_nonzero(st2 & 64)),
                                          "amise_full_accepted": int(np.count_nonzero(st2 & 128)),
                                          "amise_full_refused_like_the_reference": int(np.count_nonzero((st2 & 2048) != 0)),
                                          "bandwidth_fallback": int(np.count_nonzero(st2 & 1)),
                                          "fallback_t": int(np.count_nonzero(st2 & 32))}}
            cpu = {"value": len(tasks) / wall, "unit": UNIT, "cores": nw, "kind": cpu_backend_kind(),
                   "sample": "%d 1D + %d 2D densities of this workload at full N=%d, one per worker process on %d of %d host cores: %.1f s wall "
                             "(%.1f core-seconds; worker start-up %.1f s not counted)" % (len(tasks) - n2, n2, N, nw, host_cores(), wall, cpu_s, spawn_s)}

    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "partition": ("single GPU" if world == 1 else
                                 "densities split across %d ranks (1D round-robin, 2D by anchor blocks); every rank uploads 1/%d of the rows "
                                 "over PCIe and stores them into the peers' column stores over NVLink; finished grids are stored into every "
                                 "rank's gathered window (transport: %s)" % (world, world, pg.transport)),
                   "l2": "inputs (%.1f GB) are larger than L2; no flush needed between steps" % ((N * P * 8 + N * 8) / 1e9),
                   "datagen_s": round(t_gen, 2), "extra_untimed_settle_steps": extra_warm,
                   "rank0_cpu_affinity": (None if numa_cpus is None else "%d CPUs next to GPU %d (NVML ideal affinity)" % (len(numa_cpus), dev))},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "s_per_step": e2e_s, "checksum": checksum, "parts": e2e_parts,
                "path": "MCSamples.setSamples(pinned host) + updateBaseStatistics() [H2D, fused moments behind the chunks] -> "
                        "prefetch_triangle() [quantiles, 1D + 2D batches with contour levels] -> host grids (pooled pinned buffers)",
                "bytes_are": ("h2d: rank 0's row block + all weights (every rank uploads its own block); d2h: all grids, to rank 0's "
                              "host (prefetch_triangle(root=0))") if world > 1 else "total", "diagnostic_call": e2e_diag},
        "gpu_launches": int(launches), "phases_ms": {k: round(v, 3) for k, v in phases.items()}, "step_ms": [round(x, 3) for x in step_ms],
        "host_ms": host_log[-len(step_ms):], "group_ms": group_timings[-1] if group_timings else None,
        "clocks": clk, "roofline": roof, "measured_peaks": measured, "kernels": kernels, "hist1d": hist1d, "stats_pass": stats_pass,
        "cpu_baseline": cpu, "parity_check": parity, "multi_gpu_parity": multi,
    }
    print(json.dumps(line))
    _abi.free_pinned(xh)
    _abi.free_pinned(wh)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--p", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / full-size parity leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU vs single-GPU comparison")
    ap.add_argument("--parity-1d", type=int, default=None, help="debug: limit the CPU parity leg to this many 1D and 2D densities")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the ranks to the CPUs next to their GPUs")
    ap.add_argument("--nccl-gather", action="store_true", help="force the fallback transport (NCCL all-gather after the batch)")
    args = ap.parse_args()
    if args.n is None:
        args.n = WORKLOADS[args.workload]["n"]
    if args.p is None:
        args.p = WORKLOADS[args.workload]["p"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
