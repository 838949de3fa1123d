"""ctypes binding of libgdk.so (include/gdk.h).  The library is hand-written CUDA for sm_100a; there is no
CPU fallback: if the shared library is missing or no CUDA device is present, construction fails loudly."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgdk.so")

GDK_OUT_DEVICE = 1
GDK_BW_ONLY = 2
GDK_OUT_PEERS = 4
GDK_WIN_G1, GDK_WIN_G2, GDK_WIN_X, GDK_WIN_STATS = 0, 1, 2, 3
GDK_ROW_BLOCK = 65536

ST_BW_FALLBACK = 1
ST_BW_FAILED_NONE = 2
ST_USED_BRENT = 4
ST_SMALL_SMOOTH = 8
ST_ZERO_MAX = 16
ST_FALLBACK_T = 32
ST_AMISE_CORR = 64
ST_AMISE_FULL = 128
ST_BIAS_NEG = 256
ST_NONFINITE = 512
ST_CONTOUR_RANGE = 1024
ST_AMISE_ABORT = 2048

BW2D_FIXED, BW2D_PLAIN, BW2D_SHEAR, BW2D_RULE = 0, 1, 2, 3

PHASES = ["hist1d", "kde1d", "hist2d", "shear", "xform2d", "bw2d", "conv2d", "moments", "quantiles", "upload"]


class Spec1D(C.Structure):
    _fields_ = [("param", C.c_int32), ("fine_bins", C.c_int32), ("binmin", C.c_double), ("binmax", C.c_double),
                ("range_min", C.c_double), ("range_max", C.c_double), ("param_min", C.c_double),
                ("param_max", C.c_double), ("sigma_range", C.c_double), ("err", C.c_double), ("neff", C.c_double),
                ("smooth_scale_1D", C.c_double), ("width", C.c_double), ("boundary_correction_order", C.c_int32),
                ("mult_bias_correction_order", C.c_int32), ("has_limits_bot", C.c_int32), ("has_limits_top", C.c_int32),
                ("periodic", C.c_int32), ("pad", C.c_int32)]


class Result1D(C.Structure):
    _fields_ = [("kde_h", C.c_double), ("h_raw", C.c_double), ("smooth_1D", C.c_double), ("winw", C.c_int32),
                ("status", C.c_uint32), ("n_feval", C.c_int32), ("pad", C.c_int32)]


class Spec2D(C.Structure):
    _fields_ = [("px", C.c_int32), ("py", C.c_int32), ("fine_bins", C.c_int32), ("base_fine_bins", C.c_int32),
                ("xbinmin", C.c_double), ("xbinmax", C.c_double), ("ybinmin", C.c_double), ("ybinmax", C.c_double),
                ("x_sigma_range", C.c_double), ("y_sigma_range", C.c_double), ("x_err", C.c_double), ("y_err", C.c_double),
                ("neff", C.c_double), ("corr", C.c_double), ("kernel_corr", C.c_double), ("max_corr_2D", C.c_double),
                ("rx_fixed", C.c_double), ("ry_fixed", C.c_double), ("smooth_scale_2D", C.c_double),
                ("bw_mode", C.c_int32), ("boundary_correction_order", C.c_int32),
                ("mult_bias_correction_order", C.c_int32),
                ("x_has_bot", C.c_int32), ("x_has_top", C.c_int32), ("y_has_bot", C.c_int32), ("y_has_top", C.c_int32),
                ("shear_i", C.c_int32), ("shear_j", C.c_int32), ("shear_swapped", C.c_int32),
                ("r0", C.c_double), ("r1", C.c_double), ("S00", C.c_double), ("S10", C.c_double), ("S11", C.c_double),
                ("p1_min", C.c_double), ("p1_max", C.c_double),
                ("n_contours", C.c_int32), ("anchor_hint", C.c_int32), ("contours", C.c_double * 4),
                ("x_periodic", C.c_int32), ("y_periodic", C.c_int32)]


class Result2D(C.Structure):
    _fields_ = [("hx", C.c_double), ("hy", C.c_double), ("c", C.c_double), ("rx", C.c_double), ("ry", C.c_double),
                ("t_star", C.c_double), ("winw", C.c_int32), ("status", C.c_uint32), ("n_brent", C.c_int32),
                ("pad", C.c_int32), ("levels", C.c_double * 4)]


RES2D_DTYPE = np.dtype([("hx", "f8"), ("hy", "f8"), ("c", "f8"), ("rx", "f8"), ("ry", "f8"), ("t_star", "f8"), ("winw", "i4"),
                        ("status", "u4"), ("n_brent", "i4"), ("pad", "i4"), ("levels", "f8", (4,))])
assert RES2D_DTYPE.itemsize == C.sizeof(Result2D)
SPEC2D_DTYPE = np.dtype(Spec2D)  # the structured-array layout of gdk_spec2d (converted once: the planner builds one per batch)


def results2d_columns(res):
    """{field: list} of a batch of Result2D records (a ctypes array from the library, or any sequence of records)"""
    if isinstance(res, dict):  # already columns (the multi-GPU path builds them from the gathered table)
        return res
    if isinstance(res, C.Array):
        a = np.frombuffer(res, dtype=RES2D_DTYPE)
        return {k: a[k].tolist() for k in RES2D_DTYPE.names if k != "pad"}
    cols = {k: [getattr(r, k) for r in res] for k in RES2D_DTYPE.names if k not in ("pad", "levels")}
    cols["levels"] = [list(r.levels) for r in res]
    return cols


class LagJob(C.Structure):
    _fields_ = [("param", C.c_int32), ("mode", C.c_int32), ("k0", C.c_int64), ("nk", C.c_int32), ("pad", C.c_int32),
                ("mean", C.c_double), ("inv4s2", C.c_double)]


class GdkError(RuntimeError):
    pass


_lib = None


def load():
    """Load libgdk.so (once).  Raises GdkError if it was not built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdkError("libgdk.so not found at %s -- build it with `python -m getdist_b200.build` "
                       "(needs nvcc; there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    pd = C.POINTER(C.c_double)
    lib.gdk_abi_version.restype = i32
    lib.gdk_create.argtypes = [i32, C.POINTER(vp)]
    lib.gdk_create.restype = i32
    lib.gdk_destroy.argtypes = [vp]
    lib.gdk_destroy.restype = None
    lib.gdk_last_error.argtypes = [vp]
    lib.gdk_last_error.restype = C.c_char_p
    lib.gdk_alloc_pinned.argtypes = [u64, C.POINTER(vp)]
    lib.gdk_alloc_pinned.restype = i32
    lib.gdk_free_pinned.argtypes = [vp]
    lib.gdk_free_pinned.restype = i32
    lib.gdk_host_register.argtypes = [vp, u64]
    lib.gdk_host_register.restype = i32
    lib.gdk_host_unregister.argtypes = [vp]
    lib.gdk_host_unregister.restype = i32
    lib.gdk_launch_count.argtypes = [vp]
    lib.gdk_launch_count.restype = i64
    lib.gdk_timer_start.argtypes = [vp]
    lib.gdk_timer_start.restype = i32
    lib.gdk_timer_stop_ms.argtypes = [vp]
    lib.gdk_timer_stop_ms.restype = dbl
    lib.gdk_phase_ms.argtypes = [vp, i32]
    lib.gdk_phase_ms.restype = dbl
    lib.gdk_set_kernel_timing.argtypes = [vp, i32]
    lib.gdk_set_kernel_timing.restype = i32
    lib.gdk_kernel_stat.argtypes = [vp, i32, i32]
    lib.gdk_kernel_stat.restype = dbl
    lib.gdk_set_samples.argtypes = [vp, vp, i64, i32, i64, i64, vp, vp, i32]
    lib.gdk_set_samples.restype = i32
    lib.gdk_moments.argtypes = [vp] + [vp] * 9
    lib.gdk_moments.restype = i32
    lib.gdk_moments_recompute.argtypes = [vp]
    lib.gdk_moments_recompute.restype = i32
    lib.gdk_peer_init.argtypes = [vp, i32, i32]
    lib.gdk_peer_init.restype = i32
    lib.gdk_peer_targets.argtypes = [vp, u32]
    lib.gdk_peer_targets.restype = i32
    lib.gdk_window_export.argtypes = [vp, i32, u64, vp, vp]
    lib.gdk_window_export.restype = i32
    lib.gdk_window_import.argtypes = [vp, i32, i32, vp]
    lib.gdk_window_import.restype = i32
    lib.gdk_window_read.argtypes = [vp, i32, u64, u64, vp]
    lib.gdk_window_read.restype = i32
    lib.gdk_stream_sync.argtypes = [vp]
    lib.gdk_stream_sync.restype = i32
    lib.gdk_samples_prepare.argtypes = [vp, i64, i32, vp, i32]
    lib.gdk_samples_prepare.restype = i32
    lib.gdk_samples_upload.argtypes = [vp, vp, i64, i64, vp, i64, i64]
    lib.gdk_samples_upload.restype = i32
    lib.gdk_samples_finish.argtypes = [vp]
    lib.gdk_samples_finish.restype = i32
    lib.gdk_weighted_quantiles.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.gdk_weighted_quantiles.restype = i32
    lib.gdk_density1d_batch.argtypes = [vp, i32, vp, vp, i64, vp, u32]
    lib.gdk_density1d_batch.restype = i32
    lib.gdk_density2d_batch.argtypes = [vp, i32, vp, vp, vp, vp, u32]
    lib.gdk_density2d_batch.restype = i32
    lib.gdk_set_loglikes.argtypes = [vp, vp, i64, vp]
    lib.gdk_set_loglikes.restype = i32
    lib.gdk_density1d_likes_batch.argtypes = [vp, i32, vp, vp, vp, i64, vp, u32]
    lib.gdk_density1d_likes_batch.restype = i32
    lib.gdk_density2d_likes_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, u32]
    lib.gdk_density2d_likes_batch.restype = i32
    lib.gdk_density2d_masked_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, u32]
    lib.gdk_density2d_masked_batch.restype = i32
    lib.gdk_lag_sums.argtypes = [vp, i32, vp, vp]
    lib.gdk_lag_sums.restype = i32
    lib.gdk_hist1d_batch.argtypes = [vp, i32, vp, vp, i64]
    lib.gdk_hist1d_batch.restype = i32
    lib.gdk_hist2d_batch.argtypes = [vp, i32, vp, vp, vp]
    lib.gdk_hist2d_batch.restype = i32
    lib.gdk_measure_peaks.argtypes = [vp, vp]
    lib.gdk_measure_peaks.restype = i32
    lib.gdk_weighted_quantiles_range.argtypes = [vp, vp, i32, vp, i32, i64, i64, vp]
    lib.gdk_weighted_quantiles_range.restype = i32
    lib.gdk_weight_fraction_rows.argtypes = [vp, vp, i32, vp]
    lib.gdk_weight_fraction_rows.restype = i32
    lib.gdk_histnd.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
    lib.gdk_histnd.restype = i32
    if lib.gdk_abi_version() != 6:
        raise GdkError("libgdk.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One library context = one device + one resident sample store."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.gdk_create(int(device), C.byref(h))
        if rc != 0:
            raise GdkError("gdk_create(device=%d) failed (code %d): no usable CUDA device -- this package has no CPU "
                           "fallback" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.gdk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise GdkError("%s failed (code %d): %s" % (what, rc, self.lib.gdk_last_error(self.h).decode()))

    def set_samples(self, X, w=None, chain_offsets=None, group=None):
        X = np.asarray(X)
        if X.dtype != np.float64 or X.ndim != 2:
            raise GdkError("samples must be a 2D float64 array")
        N, P = X.shape
        rs, cs = X.strides[0] // 8, X.strides[1] // 8
        if not ((cs == 1 and rs >= P) or (rs == 1 and cs >= N)):
            X = np.ascontiguousarray(X)
            rs, cs = P, 1
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64)
        co = None
        nch = 0
        if chain_offsets is not None:
            co = np.ascontiguousarray(chain_offsets, dtype=np.int64)
            nch = co.size - 1
        self._keep = (X, w, co)
        if group is None or group.world == 1:
            self._ck(self.lib.gdk_set_samples(self.h, _ptr(X), N, P, rs, cs, _ptr(w), _ptr(co), nch), "gdk_set_samples")
        else:
            # sharded upload: this rank copies its row block over PCIe and stores it into every peer's column store over
            # NVLink; the peers do the same with theirs (include/gdk.h, "multi-GPU")
            import time as _t

            t0 = _t.perf_counter()
            self.peer_init(group.rank, group.world)
            self._ck(self.lib.gdk_samples_prepare(self.h, N, P, _ptr(co), nch), "gdk_samples_prepare")
            ld = (N + 63) & ~63
            nblk = self.stat_block_count(N, co)
            group.map_window(self, GDK_WIN_X, ld * P * 8)
            group.map_window(self, GDK_WIN_STATS, nblk * (3 * P + 1 + P * P) * 8)
            r0, r1 = group.row_range(N)
            t1 = _t.perf_counter()
            group.barrier()  # nobody is still reading the previous contents of the stores
            t2 = _t.perf_counter()
            self._ck(self.lib.gdk_samples_upload(self.h, _ptr(X), rs, cs, _ptr(w), r0, r1), "gdk_samples_upload")
            t3 = _t.perf_counter()
            group.barrier()  # every rank's rows and statistics records have landed
            t4 = _t.perf_counter()
            self._ck(self.lib.gdk_samples_finish(self.h), "gdk_samples_finish")
            t5 = _t.perf_counter()
            group.upload_timings = dict(prepare_map_ms=(t1 - t0) * 1e3, barrier1_ms=(t2 - t1) * 1e3, upload_ms=(t3 - t2) * 1e3,
                                        barrier2_ms=(t4 - t3) * 1e3, finish_ms=(t5 - t4) * 1e3)
        self.N, self.P, self.nchains = N, P, max(nch, 1)

    @staticmethod
    def stat_block_count(N, chain_offsets=None):
        """number of statistics blocks (GDK_ROW_BLOCK rows, cut at chain boundaries) of a data set"""
        co = [0, N] if chain_offsets is None else [int(c) for c in chain_offsets]
        n = 0
        for a, b in zip(co[:-1], co[1:]):
            n += (b - 1) // GDK_ROW_BLOCK - a // GDK_ROW_BLOCK + 1
        return n

    # -- multi-GPU windows (include/gdk.h) ------------------------------------------------------------------------
    def peer_init(self, rank, nranks):
        self._ck(self.lib.gdk_peer_init(self.h, rank, nranks), "gdk_peer_init")

    def peer_targets(self, mask):
        self._ck(self.lib.gdk_peer_targets(self.h, int(mask) & 0xFFFFFFFF), "gdk_peer_targets")

    def window_export(self, window, nbytes):
        """(device address, 64-byte IPC handle) of a window of at least nbytes"""
        handle = C.create_string_buffer(64)
        addr = C.c_uint64()
        self._ck(self.lib.gdk_window_export(self.h, window, int(nbytes), C.cast(handle, C.c_void_p), C.cast(C.byref(addr), C.c_void_p)),
                 "gdk_window_export")
        return int(addr.value), handle.raw

    def window_import(self, window, peer, handle):
        buf = C.create_string_buffer(handle, 64)
        self._ck(self.lib.gdk_window_import(self.h, window, peer, C.cast(buf, C.c_void_p)), "gdk_window_import")

    def window_read(self, window, offset, out, sync=True):
        """device -> host copy of out.nbytes bytes of a window starting at byte `offset`; sync=False returns with the
        copy in flight on the library stream (stream_sync() before the data is read)"""
        self._ck(self.lib.gdk_window_read(self.h, window, int(offset), int(out.nbytes) | (0 if sync else 1 << 63), _ptr(out)),
                 "gdk_window_read")
        return out

    def stream_sync(self):
        self._ck(self.lib.gdk_stream_sync(self.h), "gdk_stream_sync")

    def moments_recompute(self):
        self._ck(self.lib.gdk_moments_recompute(self.h), "gdk_moments_recompute")

    def set_loglikes(self, loglikes):
        """uploads the log-likelihoods, builds the mean-likelihood weights on the device, returns mean_loglike"""
        if loglikes is None:
            self._ck(self.lib.gdk_set_loglikes(self.h, None, 0, None), "gdk_set_loglikes")
            return None
        ll = np.ascontiguousarray(loglikes, dtype=np.float64)
        out = C.c_double()
        self._ck(self.lib.gdk_set_loglikes(self.h, _ptr(ll), ll.size, C.cast(C.byref(out), C.c_void_p)), "gdk_set_loglikes")
        return float(out.value)

    def moments(self):
        P, nch = self.P, self.nchains
        out = dict(means=np.empty(P), vars=np.empty(P), cov=np.empty((P, P)), scalars=np.empty(8), xmin=np.empty(P),
                   xmax=np.empty(P), chain_means=np.empty((nch, P)), chain_covs=np.empty((nch, P, P)),
                   chain_norms=np.empty(nch))
        self._ck(self.lib.gdk_moments(self.h, *[_ptr(out[k]) for k in
                                                ("means", "vars", "cov", "scalars", "xmin", "xmax", "chain_means",
                                                 "chain_covs", "chain_norms")]), "gdk_moments")
        return out

    def weighted_quantiles(self, params, fracs):
        params = np.ascontiguousarray(params, dtype=np.int32)
        fracs = np.ascontiguousarray(fracs, dtype=np.float64)
        out = np.empty((params.size, fracs.size))
        self._ck(self.lib.gdk_weighted_quantiles(self.h, _ptr(params), params.size, _ptr(fracs), fracs.size, _ptr(out)),
                 "gdk_weighted_quantiles")
        return out

    def weighted_quantiles_range(self, params, fracs, start, end):
        """order statistics of the rows [start, end) (confidence(..., start=, end=), chains.py:793-838)"""
        params = np.ascontiguousarray(params, dtype=np.int32)
        fracs = np.ascontiguousarray(fracs, dtype=np.float64)
        out = np.empty((params.size, fracs.size))
        self._ck(self.lib.gdk_weighted_quantiles_range(self.h, _ptr(params), params.size, _ptr(fracs), fracs.size, int(start), int(end),
                                                       _ptr(out)), "gdk_weighted_quantiles_range")
        return out

    def weight_fraction_rows(self, fracs):
        """np.searchsorted(np.cumsum(weights), fracs * sum(weights)) (getFractionIndices, mcsamples.py:668-680)"""
        fracs = np.ascontiguousarray(fracs, dtype=np.float64)
        out = np.empty(fracs.size, dtype=np.int64)
        self._ck(self.lib.gdk_weight_fraction_rows(self.h, _ptr(fracs), fracs.size, _ptr(out)), "gdk_weight_fraction_rows")
        return out

    def histnd(self, params, nbins, binmin, binmax, which=0):
        """raw ND histogram, returned with the reference's shape nbins[::-1] (C order: axis 0 of `params` is the fastest)"""
        params = np.ascontiguousarray(params, dtype=np.int32)
        nb = np.ascontiguousarray(nbins, dtype=np.int32)
        lo = np.ascontiguousarray(binmin, dtype=np.float64)
        hi = np.ascontiguousarray(binmax, dtype=np.float64)
        out = np.empty(int(np.prod(nb.astype(np.int64))))
        self._ck(self.lib.gdk_histnd(self.h, params.size, _ptr(params), _ptr(nb), _ptr(lo), _ptr(hi), int(which), _ptr(out)), "gdk_histnd")
        return out.reshape(tuple(int(n) for n in nb[::-1]))

    def density1d_batch(self, specs, out=None, device_ptr=None, likes=False, stride=None, peers=False):
        n = len(specs)
        arr = (Spec1D * n)(*specs)
        stride = max(s.fine_bins for s in specs) if stride is None else int(stride)
        res = (Result1D * n)()
        if likes:  # get1DDensityGridData(meanlikes=True): (P, likes, results)
            P, L = np.empty((n, stride)), np.empty((n, stride))
            self._ck(self.lib.gdk_density1d_likes_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(P), _ptr(L), stride,
                                                        C.cast(res, C.c_void_p), 0), "gdk_density1d_likes_batch")
            return P, L, list(res)
        if device_ptr is not None:
            self._ck(self.lib.gdk_density1d_batch(self.h, n, C.cast(arr, C.c_void_p), C.c_void_p(device_ptr), stride,
                                                  C.cast(res, C.c_void_p), GDK_OUT_PEERS if peers else GDK_OUT_DEVICE),
                     "gdk_density1d_batch")
            return None, list(res)
        P = np.empty((n, stride)) if out is None else out
        self._ck(self.lib.gdk_density1d_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(P), stride, C.cast(res, C.c_void_p), 0),
                 "gdk_density1d_batch")
        return P, list(res)

    def hist1d_batch(self, specs):
        n = len(specs)
        arr = (Spec1D * n)(*specs)
        stride = max(s.fine_bins for s in specs)
        out = np.empty((n, stride))
        self._ck(self.lib.gdk_hist1d_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(out), stride), "gdk_hist1d_batch")
        return out

    def density2d_batch(self, specs, out=None, device_ptr=None, likes=False, offsets=None, peers=False):
        n = len(specs)
        if isinstance(specs, np.ndarray):  # structured array with the gdk_spec2d layout (vectorised planner)
            assert specs.dtype.itemsize == C.sizeof(Spec2D) and specs.flags.c_contiguous
            arr = C.c_void_p(specs.ctypes.data)
            fb = specs["fine_bins"].astype(np.int64)
            sizes = fb * fb
        else:
            arr = (Spec2D * n)(*specs)
            sizes = np.array([s.fine_bins * s.fine_bins for s in specs], dtype=np.int64)
        if offsets is None:
            offsets = np.zeros(n, dtype=np.int64)
            offsets[1:] = np.cumsum(sizes)[:-1]
        else:  # the caller's layout (gathered multi-GPU results): element offsets from device_ptr / from `out`
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            assert offsets.size == n and (device_ptr is not None or out is not None)
        total = int(sizes.sum())
        res = (Result2D * n)()
        if likes:  # get2DDensityGridData(meanlikes=True): (P, likes, offsets, results)
            out, lout = np.empty(total), np.empty(total)
            self._ck(self.lib.gdk_density2d_likes_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(out), _ptr(lout), _ptr(offsets),
                                                        C.cast(res, C.c_void_p), 0), "gdk_density2d_likes_batch")
            return out, lout, offsets, list(res)
        if device_ptr is not None:
            self._ck(self.lib.gdk_density2d_batch(self.h, n, C.cast(arr, C.c_void_p), C.c_void_p(device_ptr), _ptr(offsets),
                                                  C.cast(res, C.c_void_p), GDK_OUT_PEERS if peers else GDK_OUT_DEVICE),
                     "gdk_density2d_batch")
            return None, offsets, res  # the ctypes array of records (read column-wise by the callers)
        if out is None:
            out = result_buffer(total)
        self._ck(self.lib.gdk_density2d_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(out), _ptr(offsets),
                                              C.cast(res, C.c_void_p), 0), "gdk_density2d_batch")
        return out, offsets, res  # the ctypes array of records (indexable like a list; _finish_2d reads it column-wise)

    def bandwidth2d_batch(self, specs):
        """bandwidth stage only (GDK_BW_ONLY): results with rx, ry, c, winw, status; no grids"""
        n = len(specs)
        assert isinstance(specs, np.ndarray) and specs.dtype.itemsize == C.sizeof(Spec2D) and specs.flags.c_contiguous
        offsets = np.zeros(n, dtype=np.int64)
        res = (Result2D * n)()
        self._ck(self.lib.gdk_density2d_batch(self.h, n, C.c_void_p(specs.ctypes.data), None, _ptr(offsets),
                                              C.cast(res, C.c_void_p), GDK_BW_ONLY), "gdk_density2d_batch(GDK_BW_ONLY)")
        return list(res)

    def density2d_masked_batch(self, specs, masks, mask_w, likes=False):
        """specs: structured array (bandwidths fixed); masks: list of (G+2w, G+2w) float64 arrays"""
        n = len(specs)
        assert isinstance(specs, np.ndarray) and specs.dtype.itemsize == C.sizeof(Spec2D) and specs.flags.c_contiguous
        fb = specs["fine_bins"].astype(np.int64)
        sizes = fb * fb
        offsets = np.zeros(n, dtype=np.int64)
        offsets[1:] = np.cumsum(sizes)[:-1]
        msizes = np.array([m.size for m in masks], dtype=np.int64)
        moff = np.zeros(n, dtype=np.int64)
        moff[1:] = np.cumsum(msizes)[:-1]
        mbuf = np.concatenate([np.ascontiguousarray(m, dtype=np.float64).ravel() for m in masks])
        mw = np.ascontiguousarray(mask_w, dtype=np.int32)
        out = np.empty(int(sizes.sum()))
        lout = np.empty(int(sizes.sum())) if likes else None
        res = (Result2D * n)()
        self._ck(self.lib.gdk_density2d_masked_batch(self.h, n, C.c_void_p(specs.ctypes.data), _ptr(mbuf), _ptr(moff), _ptr(mw),
                                                     _ptr(out), _ptr(lout), _ptr(offsets), C.cast(res, C.c_void_p), 0),
                 "gdk_density2d_masked_batch")
        return out, lout, offsets, list(res)

    def hist2d_batch(self, specs):
        n = len(specs)
        arr = (Spec2D * n)(*specs)
        sizes = np.array([s.fine_bins * s.fine_bins for s in specs], dtype=np.int64)
        offsets = np.zeros(n, dtype=np.int64)
        offsets[1:] = np.cumsum(sizes)[:-1]
        out = np.empty(int(sizes.sum()))
        self._ck(self.lib.gdk_hist2d_batch(self.h, n, C.cast(arr, C.c_void_p), _ptr(out), _ptr(offsets)), "gdk_hist2d_batch")
        return out, offsets

    def lag_sums(self, jobs):
        """jobs: list of (param, mode, k0, nk, mean, inv4s2) -> list of arrays (one per job, nk values)"""
        n = len(jobs)
        arr = (LagJob * n)(*[LagJob(int(p), int(m), int(k0), int(nk), 0, float(mean), float(i4)) for (p, m, k0, nk, mean, i4) in jobs])
        tot = sum(int(j[3]) for j in jobs)
        out = np.empty(tot)
        self._ck(self.lib.gdk_lag_sums(self.h, n, C.cast(arr, C.c_void_p), _ptr(out)), "gdk_lag_sums")
        res, o = [], 0
        for j in jobs:
            res.append(out[o: o + int(j[3])])
            o += int(j[3])
        return res

    def measure_peaks(self):
        """in-tree microbenchmarks: the roofline denominators of the kernels that are not HBM-bound"""
        out = np.zeros(8)
        self._ck(self.lib.gdk_measure_peaks(self.h, _ptr(out)), "gdk_measure_peaks")
        return dict(fp64_tflops=out[0], smem_updates_per_s=out[1], smem_updates_random_per_s=out[2], l2_red_per_s=out[3],
                    hbm_read_gbs=out[4])

    def timer_start(self):
        self._ck(self.lib.gdk_timer_start(self.h), "gdk_timer_start")

    def timer_stop_ms(self):
        return float(self.lib.gdk_timer_stop_ms(self.h))

    def phase_ms(self):
        return {nm: self.lib.gdk_phase_ms(self.h, i) for i, nm in enumerate(PHASES)}

    KERNEL_SLOTS = ("k_bin8c", "k_bucket_records", "k_hist2d_records", "k_shear_minmax_tma", "k_shear_hist_w",
                    "k_conv2d<0>", "k_conv2d<1>", "k_hist1d_tma", "k_kde1d", "k_qhist", "k_qscan+k_qgather+k_qselect",
                    "k_col_sums", "k_cov_tiles", "k_xform_rows", "k_xform_cols", "k_bw2d", "k_contours2d", "k_stats_fused")

    def set_kernel_timing(self, on):
        self._ck(self.lib.gdk_set_kernel_timing(self.h, 1 if on else 0), "gdk_set_kernel_timing")

    def kernel_stats(self):
        """{kernel: dict(ms, launches, bytes, flops)} accumulated since set_kernel_timing(True)"""
        out = {}
        for i, nm in enumerate(self.KERNEL_SLOTS):
            n = int(self.lib.gdk_kernel_stat(self.h, i, 1))
            if n > 0:
                out[nm] = dict(ms=self.lib.gdk_kernel_stat(self.h, i, 0), launches=n,
                               bytes=self.lib.gdk_kernel_stat(self.h, i, 2), flops=self.lib.gdk_kernel_stat(self.h, i, 3))
        return out

    def wall_ms(self):
        """host wall clock spent inside the last 1D batch / 2D batch / quantile call"""
        return {nm: self.lib.gdk_phase_ms(self.h, 10 + i) for i, nm in enumerate(("call_1d", "call_2d", "call_quantiles"))}

    def launch_count(self):
        return int(self.lib.gdk_launch_count(self.h))


class PinnedPool:
    """Caching allocator for page-locked host result buffers.  cudaHostAlloc of a gigabyte costs as much as computing
    the grids that fill it, so blocks are allocated once and recycled: an array handed out here returns its block to the
    pool when the last view of it is garbage collected (weakref.finalize on the base ndarray).  Results of the batched
    calls larger than `threshold` bytes are placed in such blocks: device -> host copies then run at full PCIe rate and
    asynchronously behind the kernels of the next group."""

    threshold = 32 << 20

    def __init__(self):
        self.free = []  # (bytes, pointer)
        self.outstanding = 0

    def empty(self, n, dtype=np.float64):
        import weakref

        need = int(n) * np.dtype(dtype).itemsize
        need = max((need + (2 << 20) - 1) & ~((2 << 20) - 1), 2 << 20)
        best = None
        for k, (sz, _) in enumerate(self.free):
            if sz >= need and (best is None or sz < self.free[best][0]) and sz <= 2 * need:
                best = k
        if best is not None:
            sz, ptr = self.free.pop(best)
        else:
            lib = load()
            p = C.c_void_p()
            if lib.gdk_alloc_pinned(need, C.byref(p)) != 0:
                for _, q in self.free:  # release the cache and retry once
                    lib.gdk_free_pinned(C.c_void_p(q))
                self.free = []
                if lib.gdk_alloc_pinned(need, C.byref(p)) != 0:
                    raise GdkError("pinned allocation of %d bytes failed" % need)
            sz, ptr = need, p.value
        buf = (C.c_char * sz).from_address(ptr)
        base = np.frombuffer(buf, dtype=np.uint8)
        self.outstanding += 1
        weakref.finalize(base, self._release, sz, ptr)
        return base[: int(n) * np.dtype(dtype).itemsize].view(dtype)

    def _release(self, sz, ptr):
        self.outstanding -= 1
        self.free.append((sz, ptr))

    def trim(self):
        lib = load()
        for _, q in self.free:
            lib.gdk_free_pinned(C.c_void_p(q))
        self.free = []


pinned_pool = PinnedPool()


def result_buffer(n):
    """host buffer for n float64 results: pinned (pooled) when large, plain numpy otherwise"""
    return pinned_pool.empty(n) if n * 8 >= PinnedPool.threshold else np.empty(n)


def host_register(arr):
    """page-lock the memory of a numpy array this process owns or maps (full-rate, asynchronous device copies)"""
    if load().gdk_host_register(C.c_void_p(arr.ctypes.data), int(arr.nbytes)) != 0:
        raise GdkError("cudaHostRegister of %d bytes failed" % arr.nbytes)


def host_unregister(arr):
    load().gdk_host_unregister(C.c_void_p(arr.ctypes.data))


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (full-rate H2D for the e2e bench path)."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    if lib.gdk_alloc_pinned(n, C.byref(p)) != 0:
        raise GdkError("pinned allocation of %d bytes failed" % n)
    buf = (C.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def free_pinned(p):
    load().gdk_free_pinned(p)
