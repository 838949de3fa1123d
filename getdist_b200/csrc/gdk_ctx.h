// gdk_ctx.h -- the library context: resident sample store, cached statistics, grow-only device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

#include <chrono>

#include <map>
#include <string>
#include <vector>

#include "kde1d_core.cuh"
#include "kde2d_core.cuh"
#include "kernels_1d.cuh"
#include "kernels_quant.cuh"
#include "kernels_stats.cuh"

#define GDK_NPHASE 10
#define GDK_MAX_RANKS 16
#define GDK_NWINDOW 4
enum {
    GDK_PH_HIST1D = 0,
    GDK_PH_KDE1D = 1,
    GDK_PH_HIST2D = 2,
    GDK_PH_SHEAR = 3,
    GDK_PH_XFORM2D = 4,
    GDK_PH_BW2D = 5,
    GDK_PH_CONV2D = 6,
    GDK_PH_MOMENTS = 7,
    GDK_PH_QUANT = 8,
    GDK_PH_UPLOAD = 9
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        cap = n;
        return 0;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

struct Kde1dTablesHost {
    cplx* tw = nullptr;      // n-th roots (pow2 sizes)
    cplx* tw4 = nullptr;     // first n of the 4n-th roots (pow2 sizes)
    double* cos4 = nullptr;  // cos(2 pi j / 4n), j < 4n (other sizes)
    cplx* twn = nullptr;     // n-th roots (other sizes, direct DFT)
};

struct gdk_ctx {
    int device = 0, num_sms = 148, max_smem = 48 * 1024;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev0[GDK_NPHASE], ev1[GDK_NPHASE], tm0 = nullptr, tm1 = nullptr;
    double phase_acc[GDK_NPHASE];
    int phase_valid[GDK_NPHASE];
    std::string err;
    int64_t launches = 0;
    bool ktiming = false;
    struct KEv {
        cudaEvent_t a, b;
        int slot;
    };
    std::vector<KEv> kev;
    int kev_used = 0;
    double kbytes[24] = {0}, kflops[24] = {0};
    int klaunches[24] = {0};
    double wall_ms[4] = {0, 0, 0, 0};  // host wall clock of the last 1D / 2D batch call, quantile call (gdk_phase_ms 10..12)
    // sample store
    int64_t N = 0, ld = 0;
    int P = 0, nchains = 1;
    std::vector<int64_t> chain_off;
    bool unit_weights = true;
    DevBuf<double> dX, dW, stage[2];
    DevBuf<unsigned long long> dWq;
    int wshift = 0;
    double wscale = 1.0;
    unsigned long long wq_total = 0;
    double sum_w = 0, sum_w2 = 0, max_w = 0, min_w = 0, n_outliers = 0;
    // mean-likelihood weights (gdk_set_loglikes): w * exp(mean_loglike - loglike) as float64 and fixed point
    bool have_loglikes = false;
    double mean_loglike = 0, wlscale = 1.0;
    DevBuf<double> dLL, dLW;
    DevBuf<unsigned long long> dWlq, gbins_l, gbins2l;
    DevBuf<double> f2l, umask;
    // cached moments
    bool have_moments = false;
    double norm = 0;
    std::vector<double> means, cov, xmin, xmax, chain_means, chain_norm, chain_S;
    // constants / tables
    IsjConsts isj;
    std::map<int, Kde1dTablesHost> tables;
    // grow-only work buffers
    DevBuf<Seg> segs, segs2;
    DevBuf<double> scratch, dS, dmeans, fbuf, gwork, f2, a2buf, affbuf, work2d;
    DevBuf<int2> dtiles;
    DevBuf<int> iscratch, qparams;
    DevBuf<QSlot> qslots;
    DevBuf<unsigned long long> qhist, gbins, gbins2, gbins_rot;
    DevBuf<QCand> qcand;
    DevBuf<Hist1dJob> jobs1d;
    DevBuf<gdk_spec1d> specs1d;
    DevBuf<gdk_result1d> res1d;
    DevBuf<Kde1dTables> tabs1d;
    DevBuf<unsigned char> bytes2d, bytes2d_b, bytes2d_c, bytes2d_d, bytes2d_e, bytes2d_res, bytes2d_mx, bytes_arena, qbase;
    Kde2dConsts k2d;
    DevBuf<unsigned char> ix8;
    bool cluster_ok = false, use_bands = false, use_hot = true, use_sorted = true, shear_sorted = false;
    int64_t sorted_min_n = 1 << 15;
    int shear_np = 3;        // max jobs per group of k_shear_hist_w (GDK_SHEAR_NP: 1..6; fewer jobs = larger windows)
    int bw2d_threads = 256;  // CTA size of k_bw2d (GDK_BW2D_THREADS: 128 or 256)
    DevBuf<unsigned char> recs;         // bucket-sorted sweep: 32-byte records [job][position]
    DevBuf<unsigned long long> recw;    // ... and their fixed-point weights
    DevBuf<unsigned> bucket;            // bucket counts / starts / write cursors
    DevBuf<unsigned char> bytes2d_s, bytes2d_sh1, bytes2d_sh2;
    DevBuf<unsigned> bucket2;           // sheared sweep: anchor bucket counts / starts / cursors
    // per-context (= per-device) record of opted-in dynamic shared-memory sizes
    bool q_attr_set = false;
    size_t h1_tma_smem = 0, h1_smem = 0, kde1d_smem = 0;
    DevBuf<cplx> cwork2d;
    std::vector<cudaEvent_t> pipe_events;  // group-finished events of the 2D output pipeline
    // fused one-sweep statistics (kernels_stats.cuh): stat blocks of the whole data set in row order, their records on
    // the device / host, and which of them hold valid data (computed here during the upload, or pushed by a peer)
    struct StatBlock {
        int64_t r0, r1;
        int chain, seg0, nseg;
    };
    std::vector<StatBlock> sblocks;
    std::vector<Seg> ssegs_h;
    std::vector<char> sblock_done;
    DevBuf<double> dblock, spart[2];
    DevBuf<Seg> ssegs;
    DevBuf<int2> sblkseg, stiles;
    DevBuf<int> sblkout;
    int stT = 0, stNtile = 0;
    // multi-GPU (one process per GPU on one node): windows of this context mapped into the peers with CUDA IPC
    int rank = 0, nranks = 1;
    uint32_t push_mask = 0xffffffffu;  // ranks that receive the result grids of GDK_OUT_PEERS calls (gdk_peer_targets)
    int64_t row_begin = 0, row_end = 0;  // rows this rank uploads itself (the rest arrives from the peers over NVLink)
    void* peer_ptr[GDK_NWINDOW][GDK_MAX_RANKS] = {{nullptr}};
    DevBuf<double> win[2];  // GDK_WIN_G1, GDK_WIN_G2: gathered result grids
    DevBuf<unsigned char> bytes_push;
    // free device memory as last queried (cudaMemGetInfo takes a driver-wide lock; asked once per sample store, the work
    // buffers only grow afterwards)
    size_t mem_free = 0;
    size_t query_free() {
        if (!mem_free) {
            size_t t = 0;
            if (cudaMemGetInfo(&mem_free, &t) != cudaSuccess) mem_free = (size_t)8 << 30;
        }
        return mem_free;
    }
};

// per-kernel CUDA-event timing + algorithmic work counters for the bench's roofline figures (gdk_set_kernel_timing)
enum {
    GDK_K_BIN8C = 0,
    GDK_K_BUCKET_RECORDS = 1,
    GDK_K_HIST2D_RECORDS = 2,
    GDK_K_SHEAR_MINMAX = 3,
    GDK_K_SHEAR_HIST = 4,
    GDK_K_CONV2D_0 = 5,
    GDK_K_CONV2D_1 = 6,
    GDK_K_HIST1D = 7,
    GDK_K_KDE1D = 8,
    GDK_K_QHIST = 9,
    GDK_K_QFINISH = 10,  // k_qscan + k_qgather + k_qselect
    GDK_K_COL_SUMS = 11,
    GDK_K_COV_TILES = 12,
    GDK_K_XFORM_ROWS = 13,
    GDK_K_XFORM_COLS = 14,
    GDK_K_BW2D = 15,
    GDK_K_CONTOURS2D = 16,
    GDK_K_STATS_FUSED = 17,
    GDK_K_NSLOT = 24
};
struct KernelTimer {  // records an event pair around one launch when timing is on
    gdk_ctx* ctx;
    int idx = -1;
    KernelTimer(gdk_ctx* c, int slot, double bytes, double flops);
    ~KernelTimer();
};

struct PhaseTimer {
    gdk_ctx* ctx = nullptr;
    int phase = 0;
    void begin(gdk_ctx* c, int ph);
    void resume(gdk_ctx* c, int ph) {  // continue a phase begun earlier in the same call: end() moves its stop event
        ctx = c;
        phase = ph;
    }
    void end();
};

struct WallTimer {
    gdk_ctx* ctx;
    int slot;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~WallTimer() {
        if (ctx) ctx->wall_ms[slot] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

int gdk_fail(gdk_ctx* c, int code, const char* fmt, ...);
int gdk_compute_moments(gdk_ctx* ctx);
std::vector<Seg> gdk_make_segments(const gdk_ctx* c, int64_t seglen);
int gdk_upload_segs(gdk_ctx* ctx, const std::vector<Seg>& v, DevBuf<Seg>& buf);
const Kde1dTablesHost* gdk_tables_for(gdk_ctx* ctx, int n);
