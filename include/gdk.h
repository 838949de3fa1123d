/*
 * gdk.h -- C-ABI of the B200-native GetDist density kernel library (libgdk.so).
 *
 * The reference (cmbant/getdist 1.7.7) is pure Python and has NO FFI for this path:
 * the boundary it offers is the Python method surface of MCSamples / Chains / WeightedSamples.
 * Each entry point below names the reference methods (file:line under /root/reference/getdist/)
 * whose N-sized / grid-sized arithmetic it replaces.  The host-side mirror of those methods
 * (getdist_b200/mcsamples.py) is the only caller; it binds these symbols with ctypes
 * (getdist_b200/_abi.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int32: 0 = OK, negative = error (text via gdk_last_error()).
 *   - per-density soft failures (bandwidth optimiser failed -> reference fallback applied) are
 *     reported in the result structs' `status` bit mask, never as a function error.
 *   - plain pointers and sizes only; all host pointers are BORROWED for the duration of the call.
 *   - output buffers are caller-allocated.  `P_out` may be a host pointer or a device pointer
 *     (flag GDK_OUT_DEVICE) so that a caller holding a torch tensor can all-gather it with NCCL.
 *   - calls on one context must be serialised by the caller (the reference is single-threaded).
 *   - all arithmetic on the path is float64 (+ exact 64-bit fixed-point weight accumulation).
 */
#ifndef GDK_H
#define GDK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gdk_ctx gdk_ctx;

#define GDK_ABI_VERSION 6

/* error codes */
#define GDK_OK 0
#define GDK_ERR_CUDA (-1)
#define GDK_ERR_ARG (-2)
#define GDK_ERR_STATE (-3)
#define GDK_ERR_NOMEM (-4)
#define GDK_ERR_UNSUPPORTED (-5)

/* flags for the batch calls */
#define GDK_OUT_DEVICE 1u /* P_out is a device pointer */
#define GDK_OUT_PEERS 4u  /* P_out lies in this context's gathered result window (GDK_WIN_G1 for the 1D call, the base of
                             GDK_WIN_G2 for the 2D call): results stay on the device AND are stored at the same offset of
                             every peer's window over NVLink (CUDA IPC peer memory), group by group behind the
                             convolutions of the next group.  A single-rank context treats it as GDK_OUT_DEVICE.   */
#define GDK_BW_ONLY 2u    /* 2D: stop after the bandwidth stage -- only `res` (rx, ry, c, winw, status) is written; used by
                             the mask_function path, whose prior mask needs the kernel half-width first          */

/* per-density status bits (result structs) */
#define GDK_ST_BW_FALLBACK 1u    /* 1D: ISJ root failed or too small -> rule-of-thumb (mcsamples.py:1258-1268)   */
                                 /* 2D: optimiser failed -> fallback widths (mcsamples.py:1336-1345)            */
#define GDK_ST_BW_FAILED_NONE 2u /* 1D: gaussian_kde_bandwidth_binned returned None (kde_bandwidth.py:133-135)   */
#define GDK_ST_USED_BRENT 4u     /* 1D: second-root guard replaced the fsolve root (kde_bandwidth.py:124-131)    */
#define GDK_ST_SMALL_SMOOTH 8u   /* smoothing scale < 2 bins: host logs the reference's warning (:1579, :1860)   */
#define GDK_ST_ZERO_MAX 16u      /* max(P)==0: host raises DensitiesError (densities.py:85-86)                   */
#define GDK_ST_FALLBACK_T 32u    /* 2D: fallback_t replaced t* (kde_bandwidth.py:163-173)                        */
#define GDK_ST_AMISE_CORR 64u    /* 2D: fixed-correlation AMISE minimum accepted (kde_bandwidth.py:275-288)      */
#define GDK_ST_AMISE_FULL 128u   /* 2D: 3-parameter AMISE minimum accepted (kde_bandwidth.py:292-304)            */
#define GDK_ST_BIAS_NEG 256u     /* 2D: AMISE bias term negative at the closed-form h (reference raises)         */
#define GDK_ST_NONFINITE 512u    /* a non-finite intermediate was met; treated as optimiser failure              */
#define GDK_ST_CONTOUR_RANGE 1024u /* 2D: a contour level lies outside the plotted range (densities.py:50-51)     */
#define GDK_ST_AMISE_ABORT 2048u  /* 2D: 3-parameter AMISE search not run: bias negative at the correlation bound,
                                     where the reference's TNC run ends in its bare except (kde_bandwidth.py:292-304) */

/* 2D bandwidth branch (mcsamples.py:1347-1409) */
#define GDK_BW2D_FIXED 0 /* smooth_scale_2D >= 0: rx, ry given in bins by the host       */
#define GDK_BW2D_PLAIN 1 /* KernelOptimizer2D on the pair's own histogram                */
#define GDK_BW2D_SHEAR 2 /* shear branch: re-bin (p1, r0*xi + r1*xj), optimise, de-rotate */
#define GDK_BW2D_RULE 3  /* rule of thumb: sigma_range / N_eff^(1/6)                     */

/* ---------------------------------------------------------------------------------------------
 * context
 * ------------------------------------------------------------------------------------------- */
int32_t gdk_abi_version(void);
int32_t gdk_create(int32_t device, gdk_ctx** out);
void gdk_destroy(gdk_ctx* ctx);
const char* gdk_last_error(gdk_ctx* ctx);
/* pinned host staging buffers for callers that want full-rate H2D (bench e2e path) */
int32_t gdk_alloc_pinned(uint64_t bytes, void** out);
int32_t gdk_free_pinned(void* p);
/* page-lock / release host memory the caller owns (e.g. a shared-memory segment that the ranks of a node all map: each
 * rank's batch call then copies its grids straight into the one host buffer the plotting process reads)           */
int32_t gdk_host_register(void* p, uint64_t bytes);
int32_t gdk_host_unregister(void* p);
/* number of kernel launches issued by this context since creation (bench `gpu_launches`) */
int64_t gdk_launch_count(gdk_ctx* ctx);
/* CUDA-event stopwatch on the library's stream: start records an event, stop records another, waits for it
 * and returns the elapsed milliseconds (host gaps between the enclosed calls included).                   */
int32_t gdk_timer_start(gdk_ctx* ctx);
double gdk_timer_stop_ms(gdk_ctx* ctx);
/* CUDA-event time (ms) of the tagged phase of the most recent batch call:
 * 0 = 1D histogram sweep, 1 = 1D grid stage, 2 = 2D histogram pass, 3 = 2D shear re-bin,
 * 4 = 2D transforms, 5 = 2D bandwidth, 6 = 2D convolution stage, 7 = moments, 8 = quantiles, 9 = upload */
double gdk_phase_ms(gdk_ctx* ctx, int32_t phase);
/* Per-kernel CUDA-event timing for the bench's roofline figures.  gdk_set_kernel_timing(ctx, 1) resets the counters
 * and makes the batch calls bracket the launches of the tagged kernels with event pairs (on the library stream);
 * gdk_kernel_stat(ctx, slot, what): what 0 = summed event time (ms), 1 = launches, 2 = algorithmic bytes,
 * 3 = algorithmic flops, accumulated since the reset.  Slots: 0 k_bin8c, 1 k_bucket_records, 2 k_hist2d_records,
 * 3 k_shear_minmax_tiled, 4 k_shear_hist, 5 k_conv2d<0>, 6 k_conv2d<1>.                                          */
int32_t gdk_set_kernel_timing(gdk_ctx* ctx, int32_t on);
double gdk_kernel_stat(gdk_ctx* ctx, int32_t slot, int32_t what);

/* In-tree microbenchmarks (a few ms each) for the roofline denominators that are not HBM bandwidth:
 * out[0] FP64 FMA TFLOP/s; out[1] 64-bit fixed-point shared-memory histogram updates/s (ATOMS + carry + RED), lanes on
 * distinct banks; out[2] the same on random bins of a 96 x 96 window; out[3] L2 REDG.ADD.64 reductions/s at random
 * addresses of a 32 MB region; out[4] HBM read GB/s of a 16-byte streaming sweep; out[5..7] reserved (0).        */
int32_t gdk_measure_peaks(gdk_ctx* ctx, double* out);

/* ---------------------------------------------------------------------------------------------
 * multi-GPU: one process (one context) per GPU on one node -- SURVEY.md s8b/s8e.  The path shards by independent
 * densities, so there is no data-path collective; what is exchanged moves through WINDOWS: device buffers of a context
 * that every peer maps with CUDA IPC and writes with plain stores / copy engines over NVLink:
 *   GDK_WIN_G1 / GDK_WIN_G2   gathered 1D / 2D result grids (every rank ends up with every density)
 *   GDK_WIN_X                 the column store (a rank uploads 1/nranks of the rows over PCIe, the peers push the rest)
 *   GDK_WIN_STATS             stat-block records of the fused statistics sweep (moments of rows uploaded elsewhere)
 * The caller's rendezvous (torch.distributed in getdist_b200/parallel.py) carries the 64-byte handles and the barriers:
 *   gdk_peer_init -> [gdk_samples_prepare] -> gdk_window_export on every rank -> all-gather of the handles ->
 *   gdk_window_import per peer -> barrier -> the call that writes the window -> barrier -> read.
 * ------------------------------------------------------------------------------------------- */
#define GDK_WIN_G1 0
#define GDK_WIN_G2 1
#define GDK_WIN_X 2
#define GDK_WIN_STATS 3
int32_t gdk_peer_init(gdk_ctx* ctx, int32_t rank, int32_t nranks);
/* ranks (bit r = rank r) whose windows receive this rank's result grids in GDK_OUT_PEERS calls; default: every rank
 * (replicated results).  One bit set = gather to that rank only (the process that plots).                          */
int32_t gdk_peer_targets(gdk_ctx* ctx, uint32_t rank_mask);
/* make window `window` at least `bytes` large (G1/G2; X and STATS are sized by gdk_samples_prepare), return its device
 * address and the 64-byte CUDA IPC handle of its allocation.  The address changes when the window had to grow: the
 * handles must then be exchanged and imported again.                                                              */
int32_t gdk_window_export(gdk_ctx* ctx, int32_t window, uint64_t bytes, void* handle64, uint64_t* device_address);
int32_t gdk_window_import(gdk_ctx* ctx, int32_t window, int32_t peer, const void* handle64);
/* device -> host copy of a byte range of a window (after the barrier that follows the writers).  With the top bit of
 * `bytes` set the call returns with the copy in flight on the library stream; gdk_stream_sync waits for it.       */
int32_t gdk_window_read(gdk_ctx* ctx, int32_t window, uint64_t offset, uint64_t bytes, void* host_out);
int32_t gdk_stream_sync(gdk_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * data residency -- replaces WeightedSamples.setSamples / Chains.makeSingle state
 * (chains.py:262-308, 1488-1503) as far as the device copy is concerned.
 *   X: N x P float64, element (n, j) at X[n*row_stride + j*col_stride] (strides in elements);
 *      stored on the device column-major (one contiguous N-vector per parameter).
 *   w: N weights or NULL (unit weights).  Must be >= 0.
 *   chain_offsets: nchains+1 row offsets (chains.py:1497) or NULL (single chain).
 * ------------------------------------------------------------------------------------------- */
int32_t gdk_set_samples(gdk_ctx* ctx, const double* X, int64_t N, int32_t P, int64_t row_stride,
                        int64_t col_stride, const double* w, const int64_t* chain_offsets, int32_t nchains);
/* The same in three steps (gdk_set_samples = prepare + upload(0, N) + finish), for the sharded upload of a multi-rank
 * group: prepare allocates the store and lays out the statistics; upload copies rows [row_begin, row_end) of X (X
 * points at row 0 of the full matrix; the range must be cut at multiples of GDK_ROW_BLOCK rows) and ALL the weights,
 * transposes them into the column store, runs the fused statistics sweep on them chunk by chunk behind the copies, and
 * pushes rows and statistics records into every imported peer window; finish (after the group's barrier) derives the
 * fixed-point weights and merges the statistics records of all row blocks -- in an order fixed by the data layout, so
 * means / covariances are bit-identical on 1 and on N ranks.                                                      */
#define GDK_ROW_BLOCK 65536
int32_t gdk_samples_prepare(gdk_ctx* ctx, int64_t N, int32_t P, const int64_t* chain_offsets, int32_t nchains);
int32_t gdk_samples_upload(gdk_ctx* ctx, const double* X, int64_t row_stride, int64_t col_stride, const double* w,
                           int64_t row_begin, int64_t row_end);
int32_t gdk_samples_finish(gdk_ctx* ctx);

/* log-likelihoods of the stored rows -- replaces the mean_loglike dot product of setMeans (chains.py:380-381)
 * and the N-sized mean-likelihood weights  weights * exp(mean_loglike - loglikes)  of the `meanlikes` option
 * (mcsamples.py:1556-1561, 1829-1831), built on the device (float64, then the same 64-bit fixed point as the
 * sample weights).  loglikes: n = N doubles (borrowed) or NULL to clear.  mean_loglike_out (optional) receives
 * sum(w * loglikes) / sum(w).  Call after gdk_set_samples.                                                    */
int32_t gdk_set_loglikes(gdk_ctx* ctx, const double* loglikes, int64_t n, double* mean_loglike_out);

/* ---------------------------------------------------------------------------------------------
 * weighted moments -- replaces setMeans/getMeans (chains.py:373-398), getVars (:400-412),
 * cov/_setCov/getCov (:709-733, 339-361), the w statistics of updateBaseStatistics
 * (:1340-1352, mcsamples.py:552-562), and the per-chain means/covs of
 * getGelmanRubinEigenvalues (:1446-1474).  All outputs optional (NULL to skip).
 *   scalars[8]: sum w, sum w^2, max w, #outliers(w > mult_max), N, min w, 0, 0
 *   chain_covs: nchains x P x P, centred on the CHAIN mean, normalised by the chain's sum w.
 * ------------------------------------------------------------------------------------------- */
int32_t gdk_moments(gdk_ctx* ctx, double* means, double* vars, double* cov, double* scalars, double* xmin,
                    double* xmax, double* chain_means, double* chain_covs, double* chain_norms);
/* run the fused statistics sweep again over the resident store (normally it rides behind the upload chunks): the
 * stats pass on its own, for measurements                                                                          */
int32_t gdk_moments_recompute(gdk_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * exact weighted order statistics -- replaces initParamConfidenceData + confidence
 * (chains.py:793-838): value of the first sample, in sorted order, whose inclusive cumulative
 * weight reaches frac * sum(w); clamped to the last sample.  out[np * nf].
 * ------------------------------------------------------------------------------------------- */
int32_t gdk_weighted_quantiles(gdk_ctx* ctx, const int32_t* params, int32_t np, const double* fracs,
                               int32_t nf, double* out);

/* Same on the row range [row_begin, row_end): confidence(paramVec, limfrac, start=, end=) (chains.py:793-838), as the
 * split tests of getConvergeTests issue it (mcsamples.py:1013-1031).  Fractions refer to the weight of the range.   */
int32_t gdk_weighted_quantiles_range(gdk_ctx* ctx, const int32_t* params, int32_t np, const double* fracs,
                                     int32_t nf, int64_t row_begin, int64_t row_end, double* out);

/* getFractionIndices (mcsamples.py:668-680): rows_out[i] = np.searchsorted(np.cumsum(weights), fracs[i] * sum(w)),
 * the first row whose inclusive cumulative weight reaches the target (exact fixed-point sums).                  */
int32_t gdk_weight_fraction_rows(gdk_ctx* ctx, const double* fracs, int32_t nf, int64_t* rows_out);

/* raw (unsmoothed) ND histogram -- the N-sized part of getRawNDDensityGridData (mcsamples.py:2098-2166):
 * _binSamples per axis (:1486-1498) + _makeNDhist (:2065-2079).  out: prod(nbins) doubles, axis 0 fastest (the
 * reference reshapes to nbins[::-1], C order).  which: 0 = sample weights, 1 = mean-likelihood weights
 * weights*exp(mean_loglike-loglikes) (:2155-2159), 2 = profile likelihood max(exp(-bestfit-loglike)) per bin
 * (:2163-2169).  which != 0 needs gdk_set_loglikes.  ndim <= 8.                                                     */
int32_t gdk_histnd(gdk_ctx* ctx, int32_t ndim, const int32_t* params, const int32_t* nbins, const double* binmin,
                   const double* binmax, int32_t which, double* out);

/* ---------------------------------------------------------------------------------------------
 * 1D densities -- replaces the body of get1DDensityGridData (mcsamples.py:1517-1686) after
 * _initParam: _binSamples + bincount (:1486-1498, 1554), getAutoBandwidth1D (:1237-1283) with
 * gaussian_kde_bandwidth_binned (kde_bandwidth.py:102-135), Kernel1D (:129-135), convolve1D
 * (convolve.py:196-202), boundary (:1600-1647) and multiplicative bias correction (:1649-1666),
 * normalize('max') (densities.py:71-92).
 * ------------------------------------------------------------------------------------------- */
typedef struct gdk_spec1d {
    int32_t param;                     /* column index                                            */
    int32_t fine_bins;                 /* F                                                       */
    double binmin, binmax;             /* grid geometry from _binSamples                          */
    double range_min, range_max;       /* par.range_min/max after _initParam                      */
    double param_min, param_max;       /* sample min/max                                          */
    double sigma_range, err;           /* par.sigma_range, par.err (std dev)                      */
    double neff;                       /* N_eff (par.N_eff_kde)                                   */
    double smooth_scale_1D;            /* <=0 auto, <1 in units of err, else in units of width    */
    double width;                      /* paramrange/(num_bins-1)                                 */
    int32_t boundary_correction_order; /* -1 off, 0, 1, 2                                         */
    int32_t mult_bias_correction_order;
    int32_t has_limits_bot, has_limits_top;
    int32_t periodic, pad;             /* par.periodic: circular convolution (convolve.py:326-367), no boundary /
                                          normaliser correction (mcsamples.py:1588-1666)            */
} gdk_spec1d;

typedef struct gdk_result1d {
    double kde_h;     /* par.kde_h (fraction of bin range, after the small-h fallback)            */
    double h_raw;     /* ISJ root before the fallback test (NaN if None)                          */
    double smooth_1D; /* kernel std dev in bins, after clipping                                   */
    int32_t winw;
    uint32_t status;
    int32_t n_feval; /* fixed-point evaluations used by the root finder                           */
    int32_t pad;
} gdk_result1d;

/* P_out: n x max(fine_bins) doubles, density i at P_out + i * stride (stride in doubles). */
int32_t gdk_density1d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* P_out, int64_t stride,
                            gdk_result1d* res, uint32_t flags);

/* Same, plus the mean likelihoods of get1DDensityGridData(meanlikes=True) (mcsamples.py:1556-1561, 1597-1598,
 * 1672-1684; default shade_likes_is_mean_loglikes = False): likes_out laid out like P_out (NULL: plain call).
 * Needs gdk_set_loglikes.                                                                                   */
int32_t gdk_density1d_likes_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* P_out, double* likes_out,
                                  int64_t stride, gdk_result1d* res, uint32_t flags);

/* ---------------------------------------------------------------------------------------------
 * 2D densities -- replaces the body of get2DDensityGridData (mcsamples.py:1748-1990) after
 * _initParamRanges: _binSamples x2 + _make2Dhist (:1821-1827, 1724-1728), getAutoBandwidth2D
 * (:1285-1419) with KernelOptimizer2D (kde_bandwidth.py:146-309) and kde.bin_samples (:76-87),
 * the kernel build (:1857-1867), convolve2D (convolve.py:205-212, 405-436), boundary and bias
 * corrections (:1905-1976, masks :1688-1712), normalize('max').
 * Grids are stored [y][x] (mcsamples.py:1725).
 * ------------------------------------------------------------------------------------------- */
typedef struct gdk_spec2d {
    int32_t px, py;                    /* column indices of the x and y parameters                */
    int32_t fine_bins;                 /* G for this pair (may be scaled up, :1812-1819)          */
    int32_t base_fine_bins;            /* G used for the shear re-binning (:1373-1375)            */
    double xbinmin, xbinmax, ybinmin, ybinmax;
    double x_sigma_range, y_sigma_range;
    double x_err, y_err;
    double neff;
    double corr;                       /* actual correlation handed to getAutoBandwidth2D          */
    double kernel_corr;                /* correlation used for the kernel when bw_mode == FIXED    */
    double max_corr_2D;
    double rx_fixed, ry_fixed;         /* bw_mode FIXED: smoothing in bins                        */
    double smooth_scale_2D;            /* multiplies the auto bandwidth (abs value), :1848         */
    int32_t bw_mode;                   /* GDK_BW2D_*                                              */
    int32_t boundary_correction_order; /* -1 off, 0, 1                                            */
    int32_t mult_bias_correction_order;
    int32_t x_has_bot, x_has_top, y_has_bot, y_has_top;
    /* shear branch (host computes the 2x2 Cholesky algebra, :1364-1369) */
    int32_t shear_i, shear_j;          /* p1 = X[:, i];  p2 = r0*X[:, i] + r1*X[:, j]             */
    int32_t shear_swapped;             /* pary.has_limits: (i, j) = (py, px) and hx<->hy at the end */
    double r0, r1;
    double S00, S10, S11;              /* S * ichol[0,0], lower triangular                        */
    double p1_min, p1_max;             /* range of p1 (imin/imax or sample min/max +- 10%)        */
    /* contour levels of the normalised grid (getContourLevels, densities.py:19-56; requested by
     * get2DDensityGridData(get_density=False), mcsamples.py:1994-2002): probability fractions, 0..4 of them */
    int32_t n_contours;
    int32_t anchor_hint;               /* bucket-sorted sweep: 0 = library chooses which parameter's bins order the rows of this
                                          pair, 1 = px, 2 = py (multi-GPU partitions keep whole anchors on one rank)        */
    double contours[4];
    int32_t x_periodic, y_periodic;    /* periodic axes: convolve2D_periodic (convolve.py:215-323), masks only on the
                                          non-periodic axes (mcsamples.py:1688-1712, 1874-1976)      */
} gdk_spec2d;

typedef struct gdk_result2d {
    double hx, hy, c;   /* bandwidth matrix in parameter units as returned by getAutoBandwidth2D  */
    double rx, ry;      /* smoothing in bins                                                      */
    double t_star;      /* fixed point (NaN if not run)                                           */
    int32_t winw;
    uint32_t status;
    int32_t n_brent;    /* fixed-point evaluations used by Brent                                  */
    int32_t pad;
    double levels[4];   /* density levels enclosing spec.contours[] of the probability            */
} gdk_result2d;

/* P_out: densities packed back to back; density i (G_i x G_i doubles, [y][x]) at P_out + offsets[i]. */
int32_t gdk_density2d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* P_out,
                            const int64_t* offsets, gdk_result2d* res, uint32_t flags);

/* Same, plus the mean likelihoods of get2DDensityGridData(meanlikes=True) (mcsamples.py:1829-1831, 1886-1901,
 * 2004-2006): likes_out laid out like P_out (NULL: plain call), max-normalised.  Needs gdk_set_loglikes.     */
int32_t gdk_density2d_likes_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* P_out, double* likes_out,
                                  const int64_t* offsets, gdk_result2d* res, uint32_t flags);

/* Same, with a user prior mask per pair (mask_function of get2DDensityGridData, mcsamples.py:1909-1919, 1973-1979):
 * pair i's mask is (G_i + 2 w_i)^2 doubles at masks + mask_offsets[i] (mask_offsets[i] < 0: none), built by the caller
 * for the kernel half-width mask_w[i] that a GDK_BW_ONLY call returned; the specs must then carry that bandwidth as
 * GDK_BW2D_FIXED (rx_fixed, ry_fixed, kernel_corr) so that the half-width is reproduced (checked).  Masked pixels
 * (mask < 1e-8) are zero in P_out.  likes_out optional as above.  Periodic axes are not supported with a mask.  */
int32_t gdk_density2d_masked_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, const double* masks,
                                   const int64_t* mask_offsets, const int32_t* mask_w, double* P_out, double* likes_out,
                                   const int64_t* offsets, gdk_result2d* res, uint32_t flags);

/* ---------------------------------------------------------------------------------------------
 * lagged sums over the stored rows -- the N-sized arithmetic of the MCMC effective-sample estimate
 * getEffectiveSamplesGaussianKDE / getCorrelationLength / getAutocorrelation (chains.py:423-466,
 * 477-574; autoConvolve convolve.py:458-478, evaluated as direct lag products instead of a size-2N FFT):
 *   mode 0: out[k] = sum_{i < N-k} d_i d_{i+k},  d = (x - mean) * w                (auto-covariance)
 *   mode 1: out[k] = sum_{i < N-k} exp(-(x_i - x_{i+k})^2 * inv4s2) w_i w_{i+k}    (kernel-weighted pairs)
 * for the consecutive lags k = k0 .. k0+nk-1 (nk <= 16) of one stored column.  out is packed job after job.
 * ------------------------------------------------------------------------------------------- */
typedef struct gdk_lagjob {
    int32_t param, mode;
    int64_t k0;
    int32_t nk, pad;
    double mean;   /* mode 0 */
    double inv4s2; /* mode 1: 1 / (4 kernel_std^2) */
} gdk_lagjob;
int32_t gdk_lag_sums(gdk_ctx* ctx, int32_t njobs, const gdk_lagjob* jobs, double* out);

/* raw histograms (test hooks; also the unit the `hist HBM GB/s` roofline figure is measured on):
 * weighted fine-grid histograms exactly as _binSamples + bincount build them.                   */
int32_t gdk_hist1d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* bins_out, int64_t stride);
int32_t gdk_hist2d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* bins_out,
                         const int64_t* offsets);

#ifdef __cplusplus
}
#endif
#endif /* GDK_H */
