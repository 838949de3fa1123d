// gdk.cu -- host side of libgdk.so: context, resident sample store, and the C-ABI entry points declared in
// include/gdk.h.  No torch types, no Python: plain CUDA runtime.  All arithmetic runs in the kernels of
// kernels_*.cuh / kde*_core.cuh; the host code only plans launches and does O(P^2) bookkeeping.
#include <cuda_runtime.h>
#include <chrono>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/gdk.h"
#include "host_tables.h"
#include "kernels_1d.cuh"
#include "kernels_quant.cuh"
#include "kernels_stats.cuh"
#include "gdk_ctx.h"
#include "gdk_2d.cuh"
#include "kernels_peak.cuh"

// -------------------------------------------------------------------------------------------------
// helpers
// -------------------------------------------------------------------------------------------------
int gdk_fail(gdk_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return gdk_fail(ctx, GDK_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                            cudaGetErrorString(e_));                                                     \
    } while (0)

std::vector<Seg> gdk_make_segments(const gdk_ctx* c, int64_t seglen) {
    std::vector<Seg> v;
    seglen = std::max<int64_t>(2, seglen & ~int64_t(1));
    for (int ch = 0; ch < c->nchains; ch++) {
        const int64_t a = c->chain_off[ch], b = c->chain_off[ch + 1];
        // interior cuts on even rows so that 16-byte vector loads stay aligned
        int64_t r = a;
        while (r < b) {
            int64_t e = std::min(b, ((r + seglen) & ~int64_t(1)));
            if (e <= r) e = b;
            v.push_back(Seg{r, e, ch, 0});
            r = e;
        }
    }
    return v;
}

int gdk_upload_segs(gdk_ctx* ctx, const std::vector<Seg>& v, DevBuf<Seg>& buf) {
    int rc = buf.ensure(v.size());
    if (rc) return gdk_fail(ctx, GDK_ERR_NOMEM, "segment buffer");
    CK(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

const Kde1dTablesHost* gdk_tables_for(gdk_ctx* ctx, int n) {
    auto it = ctx->tables.find(n);
    if (it != ctx->tables.end()) return &it->second;
    Kde1dTablesHost t{};
    if (is_pow2(n) && n >= 2) {
        std::vector<cplx> tw(n), tw4(n);
        gdk_fill_roots(tw.data(), n, n);
        gdk_fill_roots(tw4.data(), 4 * n, n);
        if (cudaMalloc(&t.tw, n * sizeof(cplx)) != cudaSuccess) return nullptr;
        if (cudaMalloc(&t.tw4, n * sizeof(cplx)) != cudaSuccess) return nullptr;
        cudaMemcpy(t.tw, tw.data(), n * sizeof(cplx), cudaMemcpyHostToDevice);
        cudaMemcpy(t.tw4, tw4.data(), n * sizeof(cplx), cudaMemcpyHostToDevice);
    } else {
        std::vector<double> c4(4 * (size_t)n);
        gdk_fill_cos(c4.data(), 4 * n);
        std::vector<cplx> tw(n);
        gdk_fill_roots(tw.data(), n, n);
        if (cudaMalloc(&t.cos4, 4 * (size_t)n * sizeof(double)) != cudaSuccess) return nullptr;
        if (cudaMalloc(&t.twn, n * sizeof(cplx)) != cudaSuccess) return nullptr;
        cudaMemcpy(t.cos4, c4.data(), 4 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(t.twn, tw.data(), n * sizeof(cplx), cudaMemcpyHostToDevice);
    }
    auto r = ctx->tables.emplace(n, t);
    return &r.first->second;
}

KernelTimer::KernelTimer(gdk_ctx* c, int slot, double bytes, double flops) : ctx(c) {
    if (!c->ktiming) return;
    if (c->kev_used == (int)c->kev.size()) {
        gdk_ctx::KEv e{};
        if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
        c->kev.push_back(e);
    }
    idx = c->kev_used++;
    c->kev[idx].slot = slot;
    c->kbytes[slot] += bytes;
    c->kflops[slot] += flops;
    c->klaunches[slot]++;
    cudaEventRecord(c->kev[idx].a, c->stream);
}
KernelTimer::~KernelTimer() {
    if (idx >= 0) cudaEventRecord(ctx->kev[idx].b, ctx->stream);
}

extern "C" int32_t gdk_set_kernel_timing(gdk_ctx* ctx, int32_t on) {
    if (!ctx) return GDK_ERR_ARG;
    ctx->ktiming = on != 0;
    ctx->kev_used = 0;
    while (on && ctx->kev.size() < 1024) {  // event pairs are created up front, not inside the timed steps
        gdk_ctx::KEv e{};
        if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) break;
        ctx->kev.push_back(e);
    }
    for (int i = 0; i < GDK_K_NSLOT; i++) {
        ctx->kbytes[i] = ctx->kflops[i] = 0;
        ctx->klaunches[i] = 0;
    }
    return GDK_OK;
}

extern "C" double gdk_kernel_stat(gdk_ctx* ctx, int32_t slot, int32_t what) {
    if (!ctx || slot < 0 || slot >= GDK_K_NSLOT) return -1.0;
    if (what == 1) return (double)ctx->klaunches[slot];
    if (what == 2) return ctx->kbytes[slot];
    if (what == 3) return ctx->kflops[slot];
    double ms = 0;
    for (int i = 0; i < ctx->kev_used; i++)
        if (ctx->kev[i].slot == slot) {
            float t = 0;
            if (cudaEventSynchronize(ctx->kev[i].b) != cudaSuccess || cudaEventElapsedTime(&t, ctx->kev[i].a, ctx->kev[i].b) != cudaSuccess)
                return -1.0;
            ms += t;
        }
    return ms;
}

void PhaseTimer::begin(gdk_ctx* c, int ph) {
    ctx = c;
    phase = ph;
    cudaEventRecord(c->ev0[ph], c->stream);
}
void PhaseTimer::end() {
    cudaEventRecord(ctx->ev1[phase], ctx->stream);
    ctx->phase_valid[phase] = 1;
}

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
extern "C" int32_t gdk_abi_version(void) { return GDK_ABI_VERSION; }

extern "C" int32_t gdk_create(int32_t device, gdk_ctx** out) {
    if (!out) return GDK_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GDK_ERR_CUDA;
    if (device < 0 || device >= ndev) return GDK_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return GDK_ERR_CUDA;
    gdk_ctx* ctx = new gdk_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem = (int)prop.sharedMemPerBlockOptin;
    {
        int cl = 0;
        cudaDeviceGetAttribute(&cl, cudaDevAttrClusterLaunch, device);
        ctx->cluster_ok = cl != 0 && prop.major >= 9;
        const char* eb = getenv("GDK_BANDS");
        ctx->use_bands = eb && eb[0] == '1';
        const char* eh = getenv("GDK_HOT");
        ctx->use_hot = !(eh && eh[0] == '0');
        const char* es = getenv("GDK_SORTED");
        ctx->use_sorted = !(es && es[0] == '0');
        const char* ess = getenv("GDK_SHEAR_SORTED");
        ctx->shear_sorted = ess && ess[0] == '1';
        const char* ebt = getenv("GDK_BW2D_THREADS");
        if (ebt && (atoi(ebt) == 128 || atoi(ebt) == 256)) ctx->bw2d_threads = atoi(ebt);
        const char* enp = getenv("GDK_SHEAR_NP");
        if (enp && atoi(enp) >= 1 && atoi(enp) <= 6) ctx->shear_np = atoi(enp);
        const char* em = getenv("GDK_SORTED_MIN_N");
        if (em) ctx->sorted_min_n = atoll(em);
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return GDK_ERR_CUDA;
    }
    for (int i = 0; i < GDK_NPHASE; i++) {
        cudaEventCreate(&ctx->ev0[i]);
        cudaEventCreate(&ctx->ev1[i]);
        ctx->phase_valid[i] = 0;
    }
    gdk_fill_isj_consts(&ctx->isj);
    gdk_fill_kde2d_consts(&ctx->k2d);
    *out = ctx;
    return GDK_OK;
}

extern "C" void gdk_destroy(gdk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    for (auto& kv : ctx->tables) {
        cudaFree(kv.second.tw);
        cudaFree(kv.second.tw4);
        cudaFree(kv.second.cos4);
        cudaFree(kv.second.twn);
    }
    for (int i = 0; i < GDK_NPHASE; i++) {
        cudaEventDestroy(ctx->ev0[i]);
        cudaEventDestroy(ctx->ev1[i]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    for (int w = 0; w < GDK_NWINDOW; w++)
        for (int p = 0; p < ctx->nranks; p++)
            if (p != ctx->rank && ctx->peer_ptr[w][p]) cudaIpcCloseMemHandle(ctx->peer_ptr[w][p]);
    delete ctx;  // DevBuf destructors free device memory
}

extern "C" const char* gdk_last_error(gdk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int32_t gdk_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return GDK_ERR_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? GDK_OK : GDK_ERR_NOMEM;
}
extern "C" int32_t gdk_host_register(void* p, uint64_t bytes) {
    if (!p || !bytes) return GDK_ERR_ARG;
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return GDK_ERR_CUDA;
    }
    return GDK_OK;
}
extern "C" int32_t gdk_host_unregister(void* p) { return (p && cudaHostUnregister(p) == cudaSuccess) ? GDK_OK : GDK_ERR_CUDA; }
extern "C" int32_t gdk_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? GDK_OK : GDK_ERR_CUDA; }
extern "C" int64_t gdk_launch_count(gdk_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int32_t gdk_timer_start(gdk_ctx* ctx) {
    if (!ctx) return GDK_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->tm0) {
        cudaEventCreate(&ctx->tm0);
        cudaEventCreate(&ctx->tm1);
    }
    return cudaEventRecord(ctx->tm0, ctx->stream) == cudaSuccess ? GDK_OK : GDK_ERR_CUDA;
}
extern "C" double gdk_timer_stop_ms(gdk_ctx* ctx) {
    if (!ctx || !ctx->tm0) return -1.0;
    float ms = 0;
    if (cudaEventRecord(ctx->tm1, ctx->stream) != cudaSuccess || cudaEventSynchronize(ctx->tm1) != cudaSuccess ||
        cudaEventElapsedTime(&ms, ctx->tm0, ctx->tm1) != cudaSuccess)
        return -1.0;
    return (double)ms;
}
extern "C" double gdk_phase_ms(gdk_ctx* ctx, int32_t phase) {
    if (ctx && phase >= 10 && phase < 14) return ctx->wall_ms[phase - 10];
    if (!ctx || phase < 0 || phase >= GDK_NPHASE || !ctx->phase_valid[phase]) return -1.0;
    float ms = 0;
    if (cudaEventSynchronize(ctx->ev1[phase]) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->ev0[phase], ctx->ev1[phase]) != cudaSuccess) return -1.0;
    return (double)ms;
}

// -------------------------------------------------------------------------------------------------
// data residency
// -------------------------------------------------------------------------------------------------
// Stat blocks (kernels_stats.cuh): ST_BLOCK-row blocks cut at chain boundaries, each divided into segments at absolute
// multiples of ST_SEG rows.  The layout depends on N and the chain offsets only.
static int stats_layout(gdk_ctx* ctx) {
    ctx->sblocks.clear();
    ctx->ssegs_h.clear();
    for (int ch = 0; ch < ctx->nchains; ch++) {
        const int64_t a = ctx->chain_off[ch], b = ctx->chain_off[ch + 1];
        int64_t r = a;
        while (r < b) {
            const int64_t e = std::min<int64_t>(b, (r / ST_BLOCK + 1) * ST_BLOCK);
            gdk_ctx::StatBlock blk{r, e, ch, (int)ctx->ssegs_h.size(), 0};
            int64_t q = r;
            while (q < e) {
                const int64_t qe = std::min<int64_t>(e, (q / ST_SEG + 1) * ST_SEG);
                ctx->ssegs_h.push_back(Seg{q, qe, ch, 0});
                q = qe;
            }
            blk.nseg = (int)ctx->ssegs_h.size() - blk.seg0;
            ctx->sblocks.push_back(blk);
            r = e;
        }
    }
    ctx->sblock_done.assign(ctx->sblocks.size(), 0);
    const int P = ctx->P;
    const int T = (P + ST_T - 1) / ST_T;
    std::vector<int2> tiles;
    for (int x = 0; x < T; x++)
        for (int y = x; y < T; y++) tiles.push_back(int2{x, y});
    ctx->stT = T;
    ctx->stNtile = (int)tiles.size();
    std::vector<int2> blkseg(ctx->sblocks.size());
    std::vector<int> blkout(ctx->sblocks.size());
    for (size_t i = 0; i < ctx->sblocks.size(); i++) {
        blkseg[i] = int2{ctx->sblocks[i].seg0, ctx->sblocks[i].nseg};
        blkout[i] = (int)i;
    }
    if (ctx->ssegs.ensure(ctx->ssegs_h.size()) || ctx->sblkseg.ensure(blkseg.size()) || ctx->sblkout.ensure(blkout.size()) ||
        ctx->stiles.ensure(tiles.size()) || ctx->dblock.ensure(ctx->sblocks.size() * (size_t)st_block_stride(P)))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "statistics buffers");
    CK(cudaMemcpyAsync(ctx->ssegs.p, ctx->ssegs_h.data(), ctx->ssegs_h.size() * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->sblkseg.p, blkseg.data(), blkseg.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->sblkout.p, blkout.data(), blkout.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->stiles.p, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// fused statistics of the stat blocks that lie inside rows [ra, rb), on stream st with partial buffer k
static int stats_rows(gdk_ctx* ctx, int64_t ra, int64_t rb, cudaStream_t st, int k) {
    int b0 = -1, b1 = -1;
    for (size_t i = 0; i < ctx->sblocks.size(); i++)
        if (ctx->sblocks[i].r0 >= ra && ctx->sblocks[i].r1 <= rb) {
            if (b0 < 0) b0 = (int)i;
            b1 = (int)i + 1;
        }
    if (b0 < 0) return 0;
    const int P = ctx->P, T = ctx->stT, nt = ctx->stNtile;
    const int s0 = ctx->sblocks[b0].seg0, s1 = ctx->sblocks[b1 - 1].seg0 + ctx->sblocks[b1 - 1].nseg;
    if (ctx->spart[k].ensure((size_t)(s1 - s0) * (size_t)st_part_stride(T, nt))) return gdk_fail(ctx, GDK_ERR_NOMEM, "statistics partials");
    double rows = 0;
    for (int i = b0; i < b1; i++) rows += (double)(ctx->sblocks[i].r1 - ctx->sblocks[i].r0);
    {
        KernelTimer kt(ctx, GDK_K_STATS_FUSED, rows * (P + 1) * 8.0, rows * 2.0 * nt * ST_T * ST_T);
        dim3 g((unsigned)(s1 - s0), (unsigned)nt);
        static bool attr = false;  // per process; every context of a process sits on a B200
        if (!attr) {
            CK(cudaFuncSetAttribute(k_stats_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
            attr = true;
        }
        k_stats_fused<<<g, 128, ST_SMEM, st>>>(ctx->dX.p, ctx->ld, ctx->dW.p, ctx->ssegs.p + s0, P, T, nt, ctx->stiles.p, ctx->spart[k].p);
    }
    const int slices = std::max(1, std::min(16, (P * P + 255) / 256));
    k_stats_merge<<<dim3((unsigned)(b1 - b0), (unsigned)slices), 256, 3 * P * sizeof(double), st>>>(ctx->spart[k].p, ctx->ssegs.p, ctx->sblkseg.p + b0, ctx->sblkout.p + b0, s0,
                                                                ctx->dX.p, ctx->ld, P, T, nt, ctx->dblock.p);
    ctx->launches += 2;
    for (int i = b0; i < b1; i++) ctx->sblock_done[i] = 1;
    return 0;
}

extern "C" int32_t gdk_peer_init(gdk_ctx* ctx, int32_t rank, int32_t nranks) {
    if (!ctx) return GDK_ERR_ARG;
    if (nranks < 1 || nranks > GDK_MAX_RANKS || rank < 0 || rank >= nranks) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_peer_init: rank %d of %d", rank, nranks);
    ctx->rank = rank;
    ctx->nranks = nranks;
    return GDK_OK;
}

extern "C" int32_t gdk_peer_targets(gdk_ctx* ctx, uint32_t rank_mask) {
    if (!ctx) return GDK_ERR_ARG;
    ctx->push_mask = rank_mask;
    return GDK_OK;
}

static int window_buffer(gdk_ctx* ctx, int window, uint64_t bytes, void** ptr) {
    switch (window) {
        case GDK_WIN_G1:
        case GDK_WIN_G2:
            if (ctx->win[window].ensure((size_t)(bytes + 7) / 8)) return gdk_fail(ctx, GDK_ERR_NOMEM, "result window (%llu bytes)", (unsigned long long)bytes);
            *ptr = ctx->win[window].p;
            return 0;
        case GDK_WIN_X:
            if (!ctx->dX.p || ctx->dX.cap * 8 < bytes) return gdk_fail(ctx, GDK_ERR_STATE, "sample window: call gdk_samples_prepare first");
            *ptr = ctx->dX.p;
            return 0;
        case GDK_WIN_STATS:
            if (!ctx->dblock.p || ctx->dblock.cap * 8 < bytes) return gdk_fail(ctx, GDK_ERR_STATE, "statistics window: call gdk_samples_prepare first");
            *ptr = ctx->dblock.p;
            return 0;
    }
    return gdk_fail(ctx, GDK_ERR_ARG, "unknown window %d", window);
}

extern "C" int32_t gdk_window_export(gdk_ctx* ctx, int32_t window, uint64_t bytes, void* handle64, uint64_t* dptr) {
    if (!ctx || !handle64 || !dptr) return GDK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    int rc = window_buffer(ctx, window, bytes, &p);
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *dptr = (uint64_t)(uintptr_t)p;
    ctx->peer_ptr[window][ctx->rank] = p;
    return GDK_OK;
}

extern "C" int32_t gdk_window_import(gdk_ctx* ctx, int32_t window, int32_t peer, const void* handle64) {
    if (!ctx || !handle64 || window < 0 || window >= GDK_NWINDOW || peer < 0 || peer >= ctx->nranks) return GDK_ERR_ARG;
    if (peer == ctx->rank) return GDK_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->peer_ptr[window][peer]) {
        cudaIpcCloseMemHandle(ctx->peer_ptr[window][peer]);
        ctx->peer_ptr[window][peer] = nullptr;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gdk_fail(ctx, GDK_ERR_CUDA, "cudaIpcOpenMemHandle(window %d, peer %d): %s", window, peer, cudaGetErrorString(e));
    }
    ctx->peer_ptr[window][peer] = p;
    return GDK_OK;
}

extern "C" int32_t gdk_stream_sync(gdk_ctx* ctx) {
    if (!ctx) return GDK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return GDK_OK;
}

extern "C" int32_t gdk_window_read(gdk_ctx* ctx, int32_t window, uint64_t offset, uint64_t bytes, void* host_out) {
    if (!ctx || !host_out) return GDK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool async = (bytes >> 63) != 0;  // top bit of `bytes`: leave the copy in flight (gdk_stream_sync)
    bytes &= ~(1ull << 63);
    void* p = nullptr;
    int rc = window_buffer(ctx, window, offset + bytes, &p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(host_out, (const char*)p + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!async) CK(cudaStreamSynchronize(ctx->stream));
    return GDK_OK;
}

extern "C" int32_t gdk_samples_prepare(gdk_ctx* ctx, int64_t N, int32_t P, const int64_t* chain_offsets, int32_t nchains) {
    if (!ctx) return GDK_ERR_ARG;
    if (N <= 0 || P <= 0) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_samples_prepare: bad shape N=%lld P=%d", (long long)N, P);
    CK(cudaSetDevice(ctx->device));
    ctx->have_moments = false;
    ctx->have_loglikes = false;
    if (ctx->N != N || ctx->P != P) ctx->mem_free = 0;  // a different store: ask the driver again
    ctx->N = N;
    ctx->P = P;
    ctx->ld = (N + 63) & ~int64_t(63);
    ctx->chain_off.clear();
    if (chain_offsets && nchains > 0) {
        if (chain_offsets[0] != 0 || chain_offsets[nchains] != N) return gdk_fail(ctx, GDK_ERR_ARG, "chain_offsets must span [0, N]");
        for (int i = 0; i <= nchains; i++) {
            if (i && chain_offsets[i] <= chain_offsets[i - 1]) return gdk_fail(ctx, GDK_ERR_ARG, "empty chain");
            ctx->chain_off.push_back(chain_offsets[i]);
        }
        ctx->nchains = nchains;
    } else {
        ctx->chain_off = {0, N};
        ctx->nchains = 1;
    }
    if (ctx->dX.ensure((size_t)ctx->ld * P) || ctx->dW.ensure((size_t)ctx->ld) || ctx->dWq.ensure((size_t)ctx->ld))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "sample store (%lld x %d)", (long long)N, P);
    return stats_layout(ctx);
}

// Rows [row_begin, row_end) of X (and all of w) from the host: chunked H2D into a row-major staging buffer, transposed
// into the column store on the device while the next chunk is copied (two staging buffers, two streams).  Behind every
// chunk, on the same stream: the fused statistics of its stat blocks, and -- with peers -- the push of its rows into
// every peer's column store over NVLink.  X points at row 0 of the full matrix.
extern "C" int32_t gdk_samples_upload(gdk_ctx* ctx, const double* X, int64_t row_stride, int64_t col_stride, const double* w,
                                      int64_t row_begin, int64_t row_end) {
    if (!ctx) return GDK_ERR_ARG;
    const int64_t N = ctx->N;
    const int P = ctx->P;
    if (!X || N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "gdk_samples_upload: call gdk_samples_prepare first");
    if (row_begin < 0 || row_end > N || row_begin > row_end || (row_begin % ST_BLOCK) || (row_end != N && (row_end % ST_BLOCK)))
        return gdk_fail(ctx, GDK_ERR_ARG, "row range [%lld, %lld) must be cut at multiples of %d rows", (long long)row_begin, (long long)row_end, ST_BLOCK);
    CK(cudaSetDevice(ctx->device));
    ctx->row_begin = row_begin;
    ctx->row_end = row_end;
    PhaseTimer pt;
    pt.begin(ctx, GDK_PH_UPLOAD);
    // weights first: the statistics of the first chunk need them
    ctx->unit_weights = (w == nullptr);
    if (w) {
        CK(cudaMemcpyAsync(ctx->dW.p, w, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        k_fill<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->dW.p, N, 1.0);
        ctx->launches++;
    }
    const int nb = ctx->num_sms * 4;
    if (ctx->scratch.ensure((size_t)nb * 4 + 16)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
    k_wstats<<<nb, 256, 0, ctx->stream>>>(ctx->dW.p, N, ctx->scratch.p);
    ctx->launches++;
    std::vector<double> part((size_t)nb * 4);
    CK(cudaMemcpyAsync(part.data(), ctx->scratch.p, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    const int64_t rows_all = row_end - row_begin;
    if (col_stride == 1 && row_stride >= P) {
        // chunk = a whole number of stat blocks, about 128 MB
        int64_t chunk = std::max<int64_t>(1, (int64_t)(128ll << 20) / ((int64_t)P * 8) / ST_BLOCK) * ST_BLOCK;
        const int64_t stage_rows = std::min(chunk, std::max<int64_t>(rows_all, 1));
        if (ctx->stage[0].ensure((size_t)stage_rows * P) || ctx->stage[1].ensure((size_t)stage_rows * P))
            return gdk_fail(ctx, GDK_ERR_NOMEM, "staging buffers");
        cudaEvent_t done[2];
        CK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
        cudaStream_t st[2] = {ctx->stream, ctx->stream2};
        CK(cudaEventRecord(done[0], ctx->stream));  // stream2 must not start before the weights are in place
        CK(cudaStreamWaitEvent(ctx->stream2, done[0], 0));
        int k = 0;
        for (int64_t r0 = row_begin; r0 < row_end; r0 += chunk, k ^= 1) {
            const int64_t rows = std::min(chunk, row_end - r0);
            if (row_stride == P)
                CK(cudaMemcpyAsync(ctx->stage[k].p, X + r0 * row_stride, (size_t)rows * P * 8, cudaMemcpyHostToDevice, st[k]));
            else
                CK(cudaMemcpy2DAsync(ctx->stage[k].p, (size_t)P * 8, X + r0 * row_stride, (size_t)row_stride * 8, (size_t)P * 8,
                                     (size_t)rows, cudaMemcpyHostToDevice, st[k]));
            dim3 g((unsigned)((rows + 31) / 32), (unsigned)((P + 31) / 32));
            k_transpose_in<<<g, 256, 0, st[k]>>>(ctx->stage[k].p, rows, P, ctx->dX.p, ctx->ld, r0);
            ctx->launches++;
            for (int p = 0; p < ctx->nranks; p++) {
                if (p == ctx->rank) continue;
                double* px = (double*)ctx->peer_ptr[GDK_WIN_X][p];
                if (!px) return gdk_fail(ctx, GDK_ERR_STATE, "sample window of peer %d is not mapped", p);
                CK(cudaMemcpy2DAsync(px + r0, (size_t)ctx->ld * 8, ctx->dX.p + r0, (size_t)ctx->ld * 8, (size_t)rows * 8, (size_t)P,
                                     cudaMemcpyDeviceToDevice, st[k]));
            }
            int rc = stats_rows(ctx, r0, r0 + rows, st[k], k);
            if (rc) return rc;
        }
        CK(cudaEventRecord(done[1], ctx->stream2));
        CK(cudaStreamWaitEvent(ctx->stream, done[1], 0));
        cudaEventDestroy(done[0]);
        cudaEventDestroy(done[1]);
    } else if (row_stride == 1 && col_stride >= N) {
        if (rows_all > 0) {
            CK(cudaMemcpy2DAsync(ctx->dX.p + row_begin, (size_t)ctx->ld * 8, X + row_begin, (size_t)col_stride * 8, (size_t)rows_all * 8,
                                 (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
            for (int p = 0; p < ctx->nranks; p++) {
                if (p == ctx->rank) continue;
                double* px = (double*)ctx->peer_ptr[GDK_WIN_X][p];
                if (!px) return gdk_fail(ctx, GDK_ERR_STATE, "sample window of peer %d is not mapped", p);
                CK(cudaMemcpy2DAsync(px + row_begin, (size_t)ctx->ld * 8, ctx->dX.p + row_begin, (size_t)ctx->ld * 8, (size_t)rows_all * 8,
                                     (size_t)P, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            const int64_t step = (int64_t)16 * ST_BLOCK;
            for (int64_t r0 = row_begin; r0 < row_end; r0 += step) {
                int rc = stats_rows(ctx, r0, std::min(row_end, r0 + step), ctx->stream, 0);
                if (rc) return rc;
            }
        }
    } else {
        return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "samples must be C- or F-contiguous along one axis (strides %lld, %lld)",
                        (long long)row_stride, (long long)col_stride);
    }
    // this rank's stat-block records to the peers
    if (ctx->nranks > 1) {
        int b0 = -1, b1 = -1;
        for (size_t i = 0; i < ctx->sblocks.size(); i++)
            if (ctx->sblocks[i].r0 >= row_begin && ctx->sblocks[i].r1 <= row_end) {
                if (b0 < 0) b0 = (int)i;
                b1 = (int)i + 1;
            }
        const size_t bs = (size_t)st_block_stride(P);
        for (int p = 0; p < ctx->nranks && b0 >= 0; p++) {
            if (p == ctx->rank) continue;
            double* pb = (double*)ctx->peer_ptr[GDK_WIN_STATS][p];
            if (!pb) return gdk_fail(ctx, GDK_ERR_STATE, "statistics window of peer %d is not mapped", p);
            CK(cudaMemcpyAsync(pb + (size_t)b0 * bs, ctx->dblock.p + (size_t)b0 * bs, (size_t)(b1 - b0) * bs * 8, cudaMemcpyDeviceToDevice,
                               ctx->stream));
        }
    }
    pt.end();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    // weight statistics (every rank holds all the weights)
    double sw = 0, sw2 = 0, mx = -INFINITY, mn = INFINITY;
    for (int i = 0; i < nb; i++) {
        sw += part[i * 4 + 0];
        sw2 += part[i * 4 + 1];
        mx = std::max(mx, part[i * 4 + 2]);
        mn = std::min(mn, part[i * 4 + 3]);
    }
    if (!(mn >= 0) || !(sw > 0) || !std::isfinite(sw))
        return gdk_fail(ctx, GDK_ERR_ARG, "weights must be finite, >= 0 and not all zero (min %g, sum %g)", mn, sw);
    ctx->sum_w = sw;
    ctx->sum_w2 = sw2;
    ctx->max_w = mx;
    ctx->min_w = mn;
    return GDK_OK;
}

static int combine_moments(gdk_ctx* ctx);

// After every rank's rows (and stat-block records) have arrived: fixed-point weights, outlier count, and the merge of
// the stat blocks into per-chain and total moments.
extern "C" int32_t gdk_samples_finish(gdk_ctx* ctx) {
    if (!ctx) return GDK_ERR_ARG;
    const int64_t N = ctx->N;
    if (N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "gdk_samples_finish: no samples");
    CK(cudaSetDevice(ctx->device));
    const int nb = ctx->num_sms * 4;
    // scale 2^k with sum(w)*2^k < 2^61 (headroom for rounding of N terms)
    int e = 0;
    frexp(ctx->sum_w, &e);  // sw = m * 2^e, m in [0.5, 1)
    ctx->wshift = 61 - e;
    ctx->wscale = ldexp(1.0, ctx->wshift);
    const double mean_mult = ctx->sum_w / (double)N;
    const double mult_max = (mean_mult * (double)N) / (double)std::min<int64_t>(N / 2, 500);  // mcsamples.py:559
    if (ctx->scratch.ensure((size_t)nb * 4 + 16)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->scratch.p);
    CK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    k_make_wq<<<nb, 256, 0, ctx->stream>>>(ctx->dW.p, N, ctx->wscale, mult_max, ctx->dWq.p, acc);
    ctx->launches++;
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->nranks > 1)  // the peers' records are in place (the caller's barrier): all blocks are valid
        for (auto& d : ctx->sblock_done) d = 1;
    bool all = true;
    for (char d : ctx->sblock_done) all = all && d;
    int rc = all ? combine_moments(ctx) : 0;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->wq_total = h[0];
    ctx->n_outliers = (double)h[1];
    return rc;
}

extern "C" int32_t gdk_set_samples(gdk_ctx* ctx, const double* X, int64_t N, int32_t P, int64_t row_stride,
                                   int64_t col_stride, const double* w, const int64_t* chain_offsets, int32_t nchains) {
    if (!ctx) return GDK_ERR_ARG;
    if (!X) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_set_samples: null samples");
    if (ctx->nranks > 1) return gdk_fail(ctx, GDK_ERR_STATE, "gdk_set_samples on a multi-rank context: use prepare / upload / finish");
    int rc = gdk_samples_prepare(ctx, N, P, chain_offsets, nchains);
    if (rc) return rc;
    rc = gdk_samples_upload(ctx, X, row_stride, col_stride, w, 0, N);
    if (rc) return rc;
    return gdk_samples_finish(ctx);
}

extern "C" int32_t gdk_set_loglikes(gdk_ctx* ctx, const double* loglikes, int64_t n, double* mean_loglike_out) {
    if (!ctx) return GDK_ERR_ARG;
    if (!loglikes) {
        ctx->have_loglikes = false;
        return GDK_OK;
    }
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "gdk_set_loglikes: no samples set");
    if (n != ctx->N) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_set_loglikes: %lld values for %lld rows", (long long)n, (long long)ctx->N);
    CK(cudaSetDevice(ctx->device));
    const int64_t N = ctx->N;
    if (ctx->dLL.ensure((size_t)ctx->ld) || ctx->dLW.ensure((size_t)ctx->ld) || ctx->dWlq.ensure((size_t)ctx->ld))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "log-likelihood buffers");
    CK(cudaMemcpyAsync(ctx->dLL.p, loglikes, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
    const int nb = ctx->num_sms * 4;
    if (ctx->scratch.ensure((size_t)nb * 4 + 16)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
    // mean_loglike = sum(w * loglike) / sum(w)   (chains.py:380-381)
    k_wll_partial<<<nb, 256, 0, ctx->stream>>>(ctx->dW.p, ctx->dLL.p, N, ctx->scratch.p);
    std::vector<double> part((size_t)nb * 4);
    CK(cudaMemcpyAsync(part.data(), ctx->scratch.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double swl = 0;
    for (int i = 0; i < nb; i++) swl += part[i];
    if (!std::isfinite(swl)) return gdk_fail(ctx, GDK_ERR_ARG, "log-likelihoods must be finite (weighted sum %g)", swl);
    ctx->mean_loglike = swl / ctx->sum_w;
    // lw = w * exp(mean_loglike - loglike), then the same fixed-point construction as for the weights
    k_like_weights<<<nb, 256, 0, ctx->stream>>>(ctx->dW.p, ctx->dLL.p, ctx->mean_loglike, N, ctx->dLW.p);
    k_wstats<<<nb, 256, 0, ctx->stream>>>(ctx->dLW.p, N, ctx->scratch.p);
    CK(cudaMemcpyAsync(part.data(), ctx->scratch.p, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double sl = 0, mn = INFINITY;
    for (int i = 0; i < nb; i++) {
        sl += part[i * 4 + 0];
        mn = std::min(mn, part[i * 4 + 3]);
    }
    if (!(mn >= 0) || !(sl > 0) || !std::isfinite(sl))
        return gdk_fail(ctx, GDK_ERR_ARG, "mean-likelihood weights w*exp(mean_loglike - loglike) are not finite (sum %g)", sl);
    int e = 0;
    frexp(sl, &e);
    ctx->wlscale = ldexp(1.0, 61 - e);
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->scratch.p);
    CK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    k_make_wq<<<nb, 256, 0, ctx->stream>>>(ctx->dLW.p, N, ctx->wlscale, INFINITY, ctx->dWlq.p, acc);
    ctx->launches += 4;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->have_loglikes = true;
    if (mean_loglike_out) *mean_loglike_out = ctx->mean_loglike;
    return GDK_OK;
}

// -------------------------------------------------------------------------------------------------
// moments
// -------------------------------------------------------------------------------------------------
// Host merge of the stat-block records (row order, per chain) into per-chain means / centred second moments, then the
// totals; the order is fixed by the data layout, so the result is the same on 1 and on N ranks.
static int combine_moments(gdk_ctx* ctx) {
    const int P = ctx->P, nch = ctx->nchains;
    const size_t bs = (size_t)st_block_stride(P), nblk = ctx->sblocks.size();
    std::vector<double> rec(nblk * bs);
    CK(cudaMemcpyAsync(rec.data(), ctx->dblock.p, rec.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->chain_means.assign((size_t)nch * P, 0.0);
    ctx->chain_norm.assign(nch, 0.0);
    ctx->chain_S.assign((size_t)nch * P * P, 0.0);
    ctx->xmin.assign(P, INFINITY);
    ctx->xmax.assign(P, -INFINITY);
    std::vector<double> D(P);
    for (size_t b = 0; b < nblk; b++) {
        const double* r = &rec[b * bs];
        const int ch = ctx->sblocks[b].chain;
        for (int j = 0; j < P; j++) {
            ctx->xmin[j] = std::min(ctx->xmin[j], r[1 + P + j]);
            ctx->xmax[j] = std::max(ctx->xmax[j], r[1 + 2 * P + j]);
        }
        const double Ab = r[0];
        if (!(Ab > 0)) continue;
        double& A = ctx->chain_norm[ch];
        double* m = &ctx->chain_means[(size_t)ch * P];
        double* S = &ctx->chain_S[(size_t)ch * P * P];
        const double* Sb = r + 3 * P + 1;
        if (!(A > 0)) {
            for (int j = 0; j < P; j++) m[j] = r[1 + j];
            for (size_t e = 0; e < (size_t)P * P; e++) S[e] = Sb[e];
            A = Ab;
            continue;
        }
        const double W = A + Ab, f = A * Ab / W;
        for (int j = 0; j < P; j++) D[j] = r[1 + j] - m[j];
        for (int i = 0; i < P; i++) {
            const double fi = f * D[i];
            for (int j = 0; j < P; j++) S[(size_t)i * P + j] += Sb[(size_t)i * P + j] + fi * D[j];
        }
        for (int j = 0; j < P; j++) m[j] += D[j] * (Ab / W);
        A = W;
    }
    ctx->means.assign(P, 0.0);
    double norm = 0;
    for (int ch = 0; ch < nch; ch++) norm += ctx->chain_norm[ch];
    ctx->norm = norm;
    for (int j = 0; j < P; j++) {
        // total mean as the weighted mean of the chain means about the first chain's (no cancellation)
        const double ref = ctx->chain_means[j];
        double t = 0;
        for (int ch = 0; ch < nch; ch++) t += ctx->chain_norm[ch] * (ctx->chain_means[(size_t)ch * P + j] - ref);
        ctx->means[j] = ref + t / norm;
    }
    // global covariance: sum_c (S_c + W_c d_c d_c^T) / W with d_c = m_c - m  (no cancellation)
    ctx->cov.assign((size_t)P * P, 0.0);
    for (int i = 0; i < P; i++)
        for (int j = i; j < P; j++) {
            double t = 0;
            for (int ch = 0; ch < nch; ch++) {
                const double di = ctx->chain_means[(size_t)ch * P + i] - ctx->means[i];
                const double dj = ctx->chain_means[(size_t)ch * P + j] - ctx->means[j];
                t += ctx->chain_S[((size_t)ch * P + i) * P + j] + ctx->chain_norm[ch] * di * dj;
            }
            ctx->cov[(size_t)i * P + j] = ctx->cov[(size_t)j * P + i] = t / ctx->norm;
        }
    ctx->have_moments = true;
    return 0;
}

static int compute_moments(gdk_ctx* ctx) {
    if (ctx->have_moments) return 0;
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK(cudaSetDevice(ctx->device));
    PhaseTimer pt;
    pt.begin(ctx, GDK_PH_MOMENTS);
    // stat blocks that were not produced during the upload: one fused sweep over their rows
    const int64_t step = (int64_t)16 * ST_BLOCK;
    for (int64_t r0 = 0; r0 < ctx->N; r0 += step) {
        const int64_t r1 = std::min(ctx->N, r0 + step);
        bool need = false;
        for (size_t i = 0; i < ctx->sblocks.size(); i++)
            if (ctx->sblocks[i].r0 >= r0 && ctx->sblocks[i].r1 <= r1 && !ctx->sblock_done[i]) need = true;
        if (!need) continue;
        int rc = stats_rows(ctx, r0, r1, ctx->stream, 0);
        if (rc) return rc;
    }
    pt.end();
    return combine_moments(ctx);
}

// re-run the fused statistics sweep over the resident store (bench: the stats pass on its own)
extern "C" int32_t gdk_moments_recompute(gdk_ctx* ctx) {
    if (!ctx) return GDK_ERR_ARG;
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    for (auto& d : ctx->sblock_done) d = 0;
    ctx->have_moments = false;
    return compute_moments(ctx);
}

int gdk_compute_moments(gdk_ctx* ctx) { return compute_moments(ctx); }

extern "C" int32_t gdk_moments(gdk_ctx* ctx, double* means, double* vars, double* cov, double* scalars, double* xmin,
                               double* xmax, double* chain_means, double* chain_covs, double* chain_norms) {
    if (!ctx) return GDK_ERR_ARG;
    int rc = compute_moments(ctx);
    if (rc) return rc;
    const int P = ctx->P, nch = ctx->nchains;
    if (means) memcpy(means, ctx->means.data(), P * 8);
    if (vars)
        for (int j = 0; j < P; j++) vars[j] = ctx->cov[(size_t)j * P + j];
    if (cov) memcpy(cov, ctx->cov.data(), (size_t)P * P * 8);
    if (scalars) {
        scalars[0] = ctx->norm;
        scalars[1] = ctx->sum_w2;
        scalars[2] = ctx->max_w;
        scalars[3] = ctx->n_outliers;
        scalars[4] = (double)ctx->N;
        scalars[5] = ctx->min_w;
        scalars[6] = 0;
        scalars[7] = 0;
    }
    if (xmin) memcpy(xmin, ctx->xmin.data(), P * 8);
    if (xmax) memcpy(xmax, ctx->xmax.data(), P * 8);
    if (chain_means) memcpy(chain_means, ctx->chain_means.data(), (size_t)nch * P * 8);
    if (chain_norms) memcpy(chain_norms, ctx->chain_norm.data(), nch * 8);
    if (chain_covs)
        for (int ch = 0; ch < nch; ch++)
            for (size_t e = 0; e < (size_t)P * P; e++)
                chain_covs[(size_t)ch * P * P + e] = ctx->chain_S[(size_t)ch * P * P + e] / ctx->chain_norm[ch];
    return GDK_OK;
}

// -------------------------------------------------------------------------------------------------
// weighted quantiles
// -------------------------------------------------------------------------------------------------
// segments of the row range [a, b) (interior cuts on even rows), chain index 0
static std::vector<Seg> make_segments_range(int64_t a, int64_t b, int64_t seglen) {
    std::vector<Seg> v;
    seglen = std::max<int64_t>(2, seglen & ~int64_t(1));
    int64_t r = a;
    while (r < b) {
        int64_t e = std::min(b, ((r + seglen) & ~int64_t(1)));
        if (e <= r) e = b;
        v.push_back(Seg{r, e, 0, 0});
        r = e;
    }
    return v;
}

// total fixed-point weight of the rows [a, b)
static int range_weight(gdk_ctx* ctx, int64_t a, int64_t b, unsigned long long* out) {
    if (ctx->scratch.ensure(16)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->scratch.p);
    CK(cudaMemsetAsync(acc, 0, 8, ctx->stream));
    k_sum_u64<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->dWq.p + a, b - a, acc);
    ctx->launches++;
    CK(cudaMemcpyAsync(out, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int quantiles_impl(gdk_ctx* ctx, const int32_t* params, int32_t np, const double* fracs, int32_t nf, int64_t row_begin,
                          int64_t row_end, double* out) {
    if (!ctx) return GDK_ERR_ARG;
    WallTimer wt{ctx, 2};
    if (!params || !fracs || !out || np <= 0 || nf <= 0 || nf > QMAXF)
        return gdk_fail(ctx, GDK_ERR_ARG, "gdk_weighted_quantiles: need 1 <= nf <= %d", QMAXF);
    int rc = compute_moments(ctx);  // column min / max
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int P = ctx->P;
    const bool ranged = !(row_begin == 0 && row_end == ctx->N);
    if (row_begin < 0 || row_end > ctx->N || row_begin >= row_end)
        return gdk_fail(ctx, GDK_ERR_ARG, "gdk_weighted_quantiles_range: bad row range [%lld, %lld)", (long long)row_begin, (long long)row_end);
    unsigned long long wq_total = ctx->wq_total;
    if (ranged) {
        rc = range_weight(ctx, row_begin, row_end, &wq_total);
        if (rc) return rc;
        if (wq_total == 0) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_weighted_quantiles_range: the rows [%lld, %lld) carry no weight", (long long)row_begin, (long long)row_end);
    }
    for (int i = 0; i < np; i++)
        if (params[i] < 0 || params[i] >= P) return gdk_fail(ctx, GDK_ERR_ARG, "parameter index %d out of range", params[i]);
    const int B1_LOG2 = 13, B2_LOG2 = 10;
    const int B1 = 1 << B1_LOG2, B2 = 1 << B2_LOG2;
    std::vector<QSlot> slots((size_t)np * QMAXF);
    std::vector<QBase> qbase(np);
    memset(slots.data(), 0, slots.size() * sizeof(QSlot));
    for (int i = 0; i < np; i++) {
        const unsigned long long klo = f64_to_key(ctx->xmin[params[i]]), khi = f64_to_key(ctx->xmax[params[i]]);
        qbase[i] = QBase{klo, q_shift_for(khi - klo, B1_LOG2), B1};
        for (int s = 0; s < QMAXF; s++) {
            QSlot& q = slots[(size_t)i * QMAXF + s];
            q.klo = klo;
            q.khi = khi;
            q.below = 0;
            q.state = 1;  // unused slots stay "resolved"
            q.value = 0;
            if (s < nf) {
                double t = ceil(fracs[s] * (double)wq_total);
                if (!(t >= 1.0)) t = 1.0;
                if (t > (double)wq_total) t = (double)wq_total;
                q.target = (unsigned long long)t;
                if (q.target > wq_total) q.target = wq_total;
                if (klo == khi) {
                    q.value = key_to_f64(klo);
                } else {
                    q.state = 0;
                    q.shift = q_shift_for(khi - klo, B1_LOG2);
                }
            }
        }
    }
    PhaseTimer pt;
    pt.begin(ctx, GDK_PH_QUANT);
    const size_t nslots = slots.size();
    if (ctx->qslots.ensure(nslots) || ctx->qparams.ensure(np) || ctx->qhist.ensure(nslots * (size_t)std::max(B1 / QMAXF + 1, B2)) ||
        ctx->qhist.ensure((size_t)np * QMAXF * B2) || ctx->qcand.ensure(nslots * QCAP) || ctx->iscratch.ensure(4))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "quantile buffers");
    // pass-1 histogram lives at ghist[(p*QMAXF + 0) * B1 ...]; make sure the buffer covers it
    if (ctx->qhist.ensure((size_t)np * QMAXF * std::max(B1, B2))) return gdk_fail(ctx, GDK_ERR_NOMEM, "quantile histograms");
    CK(cudaMemcpyAsync(ctx->qslots.p, slots.data(), nslots * sizeof(QSlot), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->qparams.p, params, np * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->qbase.ensure((size_t)np * sizeof(QBase))) return gdk_fail(ctx, GDK_ERR_NOMEM, "quantile buffers");
    QBase* dqbase = reinterpret_cast<QBase*>(ctx->qbase.p);
    CK(cudaMemcpyAsync(dqbase, qbase.data(), (size_t)np * sizeof(QBase), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t want = std::max<int64_t>(1, (int64_t)ctx->num_sms * 6 / np);
    const int64_t seglen = std::max<int64_t>(1 << 15, (row_end - row_begin + want - 1) / want);
    std::vector<Seg> segs = ranged ? make_segments_range(row_begin, row_end, seglen) : gdk_make_segments(ctx, seglen);
    rc = gdk_upload_segs(ctx, segs, ctx->segs);
    if (rc) return rc;
    dim3 g((unsigned)segs.size(), (unsigned)np);
    if (!ctx->q_attr_set) {
        CK(cudaFuncSetAttribute(k_qhist, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * QMAXF * B2 * 4 + B1 * 4));
        CK(cudaFuncSetAttribute(k_qselect, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * QCAP * 8));
        ctx->q_attr_set = true;
    }
    // pass 1: shared histogram per parameter
    CK(cudaMemsetAsync(ctx->qhist.p, 0, (size_t)np * QMAXF * B1 * 8, ctx->stream));
    const double qrows = (double)(row_end - row_begin);
    {
        KernelTimer kt(ctx, GDK_K_QHIST, qrows * (np + 1) * 8.0, 0);
        k_qhist<<<g, 1024, 2 * B1 * 4, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->dWq.p, ctx->segs.p, ctx->qparams.p, ctx->qslots.p, nf, B1, 1,
                                                     ctx->qhist.p, dqbase);
    }
    k_qscan<<<np, 32 * QMAXF, 0, ctx->stream>>>(ctx->qslots.p, nf, B1, 1, 0, B2_LOG2, ctx->qhist.p);
    ctx->launches += 2;
    int next_state = 2;
    for (int iter = 0; iter < 12; iter++) {
        // refinement pass: one histogram of B2 bins per refining slot
        CK(cudaMemsetAsync(ctx->qhist.p, 0, (size_t)np * QMAXF * B2 * 8, ctx->stream));
        {
            KernelTimer kt(ctx, GDK_K_QHIST, qrows * np * 8.0, 0);
            k_qhist<<<g, 1024, 2 * nf * B2 * 4 + B1 * 4, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->dWq.p, ctx->segs.p, ctx->qparams.p,
                                                                       ctx->qslots.p, nf, B2, 0, ctx->qhist.p, dqbase);
        }
        {
            KernelTimer kt(ctx, GDK_K_QFINISH, qrows * np * 8.0, 0);
            k_qscan<<<np, 32 * QMAXF, 0, ctx->stream>>>(ctx->qslots.p, nf, B2, 0, next_state, B2_LOG2, ctx->qhist.p);
            CK(cudaMemsetAsync(ctx->iscratch.p, 0, sizeof(int), ctx->stream));
            k_qgather<<<g, 256, B1 * 4, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->dWq.p, ctx->segs.p, ctx->qparams.p, ctx->qslots.p, nf,
                                                       ctx->qcand.p, dqbase);
            k_qselect<<<(unsigned)nslots, 512, 2 * QCAP * 8, ctx->stream>>>(ctx->qslots.p, nf, ctx->qcand.p, B2_LOG2, ctx->iscratch.p);
        }
        ctx->launches += 4;
        int nover = 0;
        CK(cudaMemcpyAsync(&nover, ctx->iscratch.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (nover == 0) break;
    }
    CK(cudaMemcpyAsync(slots.data(), ctx->qslots.p, nslots * sizeof(QSlot), cudaMemcpyDeviceToHost, ctx->stream));
    pt.end();
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < np; i++)
        for (int s = 0; s < nf; s++) {
            const QSlot& q = slots[(size_t)i * QMAXF + s];
            if (q.state != 1) return gdk_fail(ctx, GDK_ERR_STATE, "quantile selection did not converge (param %d, frac %g)", params[i], fracs[s]);
            out[(size_t)i * nf + s] = q.value;
        }
    return GDK_OK;
}

extern "C" int32_t gdk_weighted_quantiles(gdk_ctx* ctx, const int32_t* params, int32_t np, const double* fracs, int32_t nf,
                                          double* out) {
    if (!ctx) return GDK_ERR_ARG;
    return quantiles_impl(ctx, params, np, fracs, nf, 0, ctx->N, out);
}

extern "C" int32_t gdk_weighted_quantiles_range(gdk_ctx* ctx, const int32_t* params, int32_t np, const double* fracs, int32_t nf,
                                                int64_t row_begin, int64_t row_end, double* out) {
    if (!ctx) return GDK_ERR_ARG;
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    return quantiles_impl(ctx, params, np, fracs, nf, row_begin, row_end, out);
}

// getFractionIndices (mcsamples.py:668-680): np.searchsorted(np.cumsum(weights), fracs * norm) -- the first row whose
// inclusive cumulative weight reaches the target (exact fixed-point sums).  Two levels: per-chunk totals, host prefix over
// the chunks, then one CTA per target scans its chunk.
extern "C" int32_t gdk_weight_fraction_rows(gdk_ctx* ctx, const double* fracs, int32_t nf, int64_t* rows_out) {
    if (!ctx) return GDK_ERR_ARG;
    if (!fracs || !rows_out || nf <= 0) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_weight_fraction_rows: bad arguments");
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK(cudaSetDevice(ctx->device));
    const int64_t N = ctx->N;
    const int nchunk = (int)((N + FRC_CHUNK - 1) / FRC_CHUNK);
    if (ctx->qhist.ensure((size_t)nchunk + 2 * (size_t)nf + 8)) return gdk_fail(ctx, GDK_ERR_NOMEM, "fraction buffers");
    unsigned long long* dsum = ctx->qhist.p;
    k_chunk_sums_u64<<<nchunk, 256, 0, ctx->stream>>>(ctx->dWq.p, N, dsum);
    ctx->launches++;
    std::vector<unsigned long long> sums(nchunk);
    CK(cudaMemcpyAsync(sums.data(), dsum, (size_t)nchunk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned long long> job(2 * (size_t)nf);  // (chunk, remaining target inside the chunk)
    std::vector<int> direct(nf, 0);
    for (int i = 0; i < nf; i++) {
        double t = ceil(fracs[i] * (double)ctx->wq_total);
        if (!(t > 0)) t = 0;
        unsigned long long target = t >= (double)ctx->wq_total ? ctx->wq_total : (unsigned long long)t;
        unsigned long long run = 0;
        int c = 0;
        while (c < nchunk - 1 && run + sums[c] < target) run += sums[c++];
        job[2 * i] = (unsigned long long)c;
        job[2 * i + 1] = target - run;
    }
    unsigned long long* djob = dsum + nchunk;
    long long* drow = reinterpret_cast<long long*>(djob + 2 * (size_t)nf);
    CK(cudaMemcpyAsync(djob, job.data(), job.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_fraction_rows<<<nf, 256, 0, ctx->stream>>>(ctx->dWq.p, N, djob, drow);
    ctx->launches++;
    std::vector<long long> rows(nf);
    CK(cudaMemcpyAsync(rows.data(), drow, (size_t)nf * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    for (int i = 0; i < nf; i++) rows_out[i] = rows[i];
    return GDK_OK;
}

// raw ND histogram: _binSamples per axis + _makeNDhist (mcsamples.py:1486-1498, 2065-2079); axis 0 is the fastest index
extern "C" int32_t gdk_histnd(gdk_ctx* ctx, int32_t ndim, const int32_t* params, const int32_t* nbins, const double* binmin,
                              const double* binmax, int32_t which, double* out) {
    if (!ctx) return GDK_ERR_ARG;
    if (ndim < 1 || ndim > HND_MAXD || !params || !nbins || !binmin || !binmax || !out || which < 0 || which > 2)
        return gdk_fail(ctx, GDK_ERR_ARG, "gdk_histnd: need 1 <= ndim <= %d", HND_MAXD);
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    if (which != 0 && !ctx->have_loglikes) return gdk_fail(ctx, GDK_ERR_STATE, "gdk_histnd: mean / max likelihoods need gdk_set_loglikes");
    CK(cudaSetDevice(ctx->device));
    HistNdJob jb{};
    jb.ndim = ndim;
    size_t total = 1;
    for (int d = 0; d < ndim; d++) {
        if (params[d] < 0 || params[d] >= ctx->P) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_histnd: parameter %d out of range", params[d]);
        if (nbins[d] < 2 || !(binmax[d] > binmin[d])) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_histnd: bad geometry for axis %d", d);
        jb.param[d] = params[d];
        jb.n[d] = nbins[d];
        jb.stride[d] = (long long)total;
        jb.binmin[d] = binmin[d];
        jb.fw[d] = (binmax[d] - binmin[d]) / (nbins[d] - 1);
        jb.inv[d] = 1.0 / jb.fw[d];
        total *= (size_t)nbins[d];
        if (total > ((size_t)1 << 28)) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "gdk_histnd: more than 2^28 bins");
    }
    if (ctx->gbins.ensure(total)) return gdk_fail(ctx, GDK_ERR_NOMEM, "ND histogram");
    CK(cudaMemsetAsync(ctx->gbins.p, 0, total * 8, ctx->stream));
    double shift = 0;
    if (which == 2) {  // profile likelihood: max over the bin of exp(-bestfit - loglike), bestfit = max(-loglike)
        const int nb = ctx->num_sms * 4;
        if (ctx->scratch.ensure((size_t)nb * 4 + 16)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
        k_wstats<<<nb, 256, 0, ctx->stream>>>(ctx->dLL.p, ctx->N, ctx->scratch.p);
        ctx->launches++;
        std::vector<double> part((size_t)nb * 4);
        CK(cudaMemcpyAsync(part.data(), ctx->scratch.p, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        double mn = INFINITY;
        for (int i = 0; i < nb; i++) mn = std::min(mn, part[i * 4 + 3]);
        shift = mn;  // exp(-bestfit - loglike) = exp(min(loglike) - loglike)
    }
    const size_t smem = total * 8 <= (size_t)ctx->max_smem - 4096 && which != 2 ? total * 8 : 0;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_histnd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (int64_t)ctx->num_sms * 2;
    std::vector<Seg> segs = gdk_make_segments(ctx, std::max<int64_t>(1 << 14, (ctx->N + want - 1) / want));
    int rc = gdk_upload_segs(ctx, segs, ctx->segs);
    if (rc) return rc;
    k_histnd<<<(unsigned)segs.size(), 512, smem, ctx->stream>>>(ctx->dX.p, ctx->ld, which == 1 ? ctx->dWlq.p : ctx->dWq.p,
                                                              which == 2 ? ctx->dLL.p : nullptr, shift, ctx->segs.p, jb, (int)total,
                                                              smem ? 1 : 0, ctx->gbins.p);
    ctx->launches++;
    CK(cudaGetLastError());
    if (ctx->fbuf.ensure(total)) return gdk_fail(ctx, GDK_ERR_NOMEM, "ND histogram output");
    if (which == 2) {
        CK(cudaMemcpyAsync(out, ctx->gbins.p, total * 8, cudaMemcpyDeviceToHost, ctx->stream));  // bit patterns of doubles
    } else {
        k_bins_to_f64<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(ctx->gbins.p, ctx->fbuf.p, (int64_t)total,
                                                                1.0 / (which == 1 ? ctx->wlscale : ctx->wscale));
        ctx->launches++;
        CK(cudaMemcpyAsync(out, ctx->fbuf.p, total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return GDK_OK;
}

// -------------------------------------------------------------------------------------------------
// 1D densities
// -------------------------------------------------------------------------------------------------
static int run_hist1d(gdk_ctx* ctx, int n, const gdk_spec1d* specs, int64_t* gstride_out, const unsigned long long* wq = nullptr,
                      DevBuf<unsigned long long>* outbuf = nullptr) {
    if (!wq) wq = ctx->dWq.p;
    DevBuf<unsigned long long>& gb = outbuf ? *outbuf : ctx->gbins;
    int maxF = 0;
    std::vector<Hist1dJob> jobs(n);
    for (int i = 0; i < n; i++) {
        const gdk_spec1d& s = specs[i];
        if (s.param < 0 || s.param >= ctx->P) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: parameter %d out of range", i, s.param);
        if (s.fine_bins < 8 || s.fine_bins > (1 << 16)) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: fine_bins %d unsupported", i, s.fine_bins);
        if (!(s.binmax > s.binmin)) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: empty bin range", i);
        maxF = std::max(maxF, s.fine_bins);
        const double fw = (s.binmax - s.binmin) / (s.fine_bins - 1);
        jobs[i] = Hist1dJob{s.param, s.fine_bins, s.binmin, fw, 1.0 / fw};
    }
    const int64_t gstride = (maxF + 15) & ~15;
    *gstride_out = gstride;
    if (gb.ensure((size_t)n * gstride) || ctx->jobs1d.ensure(n)) return gdk_fail(ctx, GDK_ERR_NOMEM, "1D histogram buffers");
    if (maxF * 8 > ctx->max_smem) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "fine_bins %d exceeds shared memory", maxF);
    CK(cudaMemsetAsync(gb.p, 0, (size_t)n * gstride * 8, ctx->stream));
    CK(cudaMemcpyAsync(ctx->jobs1d.p, jobs.data(), n * sizeof(Hist1dJob), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t want = std::max<int64_t>(1, (int64_t)ctx->num_sms * 16 / n);
    const int64_t seglen = std::max<int64_t>(1 << 15, (ctx->N + want - 1) / want);
    std::vector<Seg> segs = gdk_make_segments(ctx, seglen);
    int rc = gdk_upload_segs(ctx, segs, ctx->segs);
    if (rc) return rc;
    // TMA-pipelined kernel when every segment starts on an even row (16-byte aligned bulk copies) and the bins fit
    bool aligned = true;
    for (const Seg& sgm : segs) aligned = aligned && ((sgm.r0 & 1) == 0);
    const size_t smem_tma = (size_t)2 * H1_STAGES * H1_CHUNK * 8 + (size_t)maxF * 8;
    const bool use_tma = aligned && maxF <= 32768 && smem_tma <= (size_t)ctx->max_smem - 2048;
    PhaseTimer pt;
    dim3 g((unsigned)segs.size(), (unsigned)n);
    if (use_tma) {
        std::vector<Hist1dJobT> jt(n);
        for (int i = 0; i < n; i++) {
            int bits = 0;
            while ((1 << bits) < specs[i].fine_bins + 2) bits++;
            const int sh = 31 - bits;
            jt[i] = Hist1dJobT{jobs[i].param, jobs[i].F, sh, 0, jobs[i].binmin, jobs[i].fine_width, jobs[i].inv_width, ldexp(1.0, sh)};
        }
        if (ctx->bytes2d.ensure((size_t)n * sizeof(Hist1dJobT))) return gdk_fail(ctx, GDK_ERR_NOMEM, "1D job table");
        Hist1dJobT* djt = reinterpret_cast<Hist1dJobT*>(ctx->bytes2d.p);
        CK(cudaMemcpyAsync(djt, jt.data(), (size_t)n * sizeof(Hist1dJobT), cudaMemcpyHostToDevice, ctx->stream));
        if (smem_tma > ctx->h1_tma_smem) {
            CK(cudaFuncSetAttribute(k_hist1d_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_tma, 48 * 1024)));
            ctx->h1_tma_smem = std::max<size_t>(smem_tma, 48 * 1024);
        }
        pt.begin(ctx, GDK_PH_HIST1D);
        KernelTimer kt(ctx, GDK_K_HIST1D, (double)ctx->N * (n + 1) * 8.0, 0);
        k_hist1d_tma<<<g, 256, smem_tma, ctx->stream>>>(ctx->dX.p, ctx->ld, wq, ctx->segs.p, djt, gb.p, gstride);
    } else {
        if ((size_t)maxF * 8 > ctx->h1_smem) {
            CK(cudaFuncSetAttribute(k_hist1d, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(maxF * 8, 48 * 1024)));
            ctx->h1_smem = (size_t)std::max(maxF * 8, 48 * 1024);
        }
        pt.begin(ctx, GDK_PH_HIST1D);
        KernelTimer kt(ctx, GDK_K_HIST1D, (double)ctx->N * (n + 1) * 8.0, 0);
        k_hist1d<<<g, 256, maxF * 8, ctx->stream>>>(ctx->dX.p, ctx->ld, wq, ctx->segs.p, ctx->jobs1d.p, gb.p, gstride);
    }
    ctx->launches++;
    pt.end();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int32_t gdk_hist1d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* bins_out, int64_t stride) {
    if (!ctx) return GDK_ERR_ARG;
    if (n <= 0 || !specs || !bins_out) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_hist1d_batch: bad arguments");
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK(cudaSetDevice(ctx->device));
    int64_t gstride = 0;
    int rc = run_hist1d(ctx, n, specs, &gstride);
    if (rc) return rc;
    if (ctx->fbuf.ensure((size_t)n * gstride)) return gdk_fail(ctx, GDK_ERR_NOMEM, "histogram output");
    k_bins_to_f64<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(ctx->gbins.p, ctx->fbuf.p, (int64_t)n * gstride, 1.0 / ctx->wscale);
    ctx->launches++;
    for (int i = 0; i < n; i++)
        CK(cudaMemcpyAsync(bins_out + (int64_t)i * stride, ctx->fbuf.p + (int64_t)i * gstride, (size_t)specs[i].fine_bins * 8,
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GDK_OK;
}

extern "C" int32_t gdk_density1d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* P_out, int64_t stride,
                                       gdk_result1d* res, uint32_t flags) {
    return gdk_density1d_likes_batch(ctx, n, specs, P_out, nullptr, stride, res, flags);
}

extern "C" int32_t gdk_density1d_likes_batch(gdk_ctx* ctx, int32_t n, const gdk_spec1d* specs, double* P_out, double* likes_out,
                                             int64_t stride, gdk_result1d* res, uint32_t flags) {
    if (!ctx) return GDK_ERR_ARG;
    WallTimer wt{ctx, 0};
    const bool likes = likes_out != nullptr;
    if (likes && !ctx->have_loglikes) return gdk_fail(ctx, GDK_ERR_STATE, "meanlikes needs gdk_set_loglikes");
    if (n <= 0 || !specs || !P_out || !res) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_density1d_batch: bad arguments");
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < n; i++) {
        const gdk_spec1d& s = specs[i];
        if (s.boundary_correction_order > 2) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: boundary_correction_order must be <= 2", i);
        if (s.smooth_scale_1D <= 0 && !(s.neff > 0)) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: N_eff must be > 0", i);
        if (stride < s.fine_bins) return gdk_fail(ctx, GDK_ERR_ARG, "spec %d: output stride too small", i);
    }
    int64_t gstride = 0;
    int rc = 0;
    if (likes) {  // second weighted histogram with the mean-likelihood weights (mcsamples.py:1561), same bin indices
        rc = run_hist1d(ctx, n, specs, &gstride, ctx->dWlq.p, &ctx->gbins_l);
        if (rc) return rc;
    }
    rc = run_hist1d(ctx, n, specs, &gstride);
    if (rc) return rc;
    // twiddle / cosine tables per density
    std::vector<Kde1dTables> tabs(n);
    int maxF = 0;
    for (int i = 0; i < n; i++) {
        const Kde1dTablesHost* t = gdk_tables_for(ctx, specs[i].fine_bins);
        if (!t) return gdk_fail(ctx, GDK_ERR_NOMEM, "transform tables");
        tabs[i] = Kde1dTables{t->tw, t->tw4, t->cos4};
        maxF = std::max(maxF, specs[i].fine_bins);
    }
    const int nwork = likes ? 11 : 9;  // work arrays of F doubles per density
    const size_t smem_need = (size_t)nwork * maxF * 8;
    const int use_smem = smem_need <= (size_t)ctx->max_smem - 1024;
    if (ctx->specs1d.ensure(n) || ctx->res1d.ensure(n) || ctx->tabs1d.ensure(n) || ctx->fbuf.ensure((size_t)n * gstride * (likes ? 2 : 1)))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "1D density buffers");
    if (!use_smem && ctx->gwork.ensure((size_t)n * nwork * maxF)) return gdk_fail(ctx, GDK_ERR_NOMEM, "1D workspace");
    CK(cudaMemcpyAsync(ctx->specs1d.p, specs, n * sizeof(gdk_spec1d), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->tabs1d.p, tabs.data(), n * sizeof(Kde1dTables), cudaMemcpyHostToDevice, ctx->stream));
    if (use_smem && smem_need > ctx->kde1d_smem) {
        CK(cudaFuncSetAttribute(k_kde1d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_need, 48 * 1024)));
        ctx->kde1d_smem = std::max<size_t>(smem_need, 48 * 1024);
    }
    const bool peers_out = (flags & GDK_OUT_PEERS) != 0 && ctx->nranks > 1;
    const bool dev_out = (flags & (GDK_OUT_DEVICE | GDK_OUT_PEERS)) != 0;
    if (peers_out && (likes || P_out < ctx->win[GDK_WIN_G1].p || P_out + (size_t)n * stride > ctx->win[GDK_WIN_G1].p + ctx->win[GDK_WIN_G1].cap))
        return gdk_fail(ctx, GDK_ERR_ARG, "GDK_OUT_PEERS: P_out must lie inside the GDK_WIN_G1 window");
    double* dP = dev_out ? P_out : ctx->fbuf.p;
    double* dL = !likes ? nullptr : (dev_out ? likes_out : ctx->fbuf.p + (size_t)n * gstride);
    const int64_t pstride = dev_out ? stride : gstride;
    PhaseTimer pt;
    pt.begin(ctx, GDK_PH_KDE1D);
    {
        KernelTimer kt(ctx, GDK_K_KDE1D, (double)n * maxF * 16.0, 0);
        k_kde1d<<<n, 256, use_smem ? smem_need : 0, ctx->stream>>>(ctx->specs1d.p, ctx->gbins.p, gstride, 1.0 / ctx->wscale, ctx->isj,
                                                                   ctx->tabs1d.p, dP, pstride, ctx->res1d.p, ctx->gwork.p, use_smem,
                                                                   likes ? ctx->gbins_l.p : nullptr, 1.0 / ctx->wlscale, dL);
    }
    ctx->launches++;
    pt.end();
    CK(cudaGetLastError());
    if (peers_out) {  // this rank's rows of the gathered 1D window into every peer's copy (NVLink, copy engine)
        const size_t off = (size_t)(P_out - ctx->win[GDK_WIN_G1].p);
        for (int p = 0; p < ctx->nranks; p++) {
            if (p == ctx->rank || !((ctx->push_mask >> p) & 1u)) continue;
            double* pw = (double*)ctx->peer_ptr[GDK_WIN_G1][p];
            if (!pw) return gdk_fail(ctx, GDK_ERR_STATE, "result window of peer %d is not mapped", p);
            CK(cudaMemcpyAsync(pw + off, P_out, (size_t)n * stride * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (!dev_out)
        for (int i = 0; i < n; i++)
            CK(cudaMemcpyAsync(P_out + (int64_t)i * stride, ctx->fbuf.p + (int64_t)i * gstride, (size_t)specs[i].fine_bins * 8,
                               cudaMemcpyDeviceToHost, ctx->stream));
    if (!dev_out && likes)
        for (int i = 0; i < n; i++)
            CK(cudaMemcpyAsync(likes_out + (int64_t)i * stride, dL + (int64_t)i * gstride, (size_t)specs[i].fine_bins * 8,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(res, ctx->res1d.p, n * sizeof(gdk_result1d), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GDK_OK;
}

// -------------------------------------------------------------------------------------------------
// lagged sums
// -------------------------------------------------------------------------------------------------
extern "C" int32_t gdk_lag_sums(gdk_ctx* ctx, int32_t njobs, const gdk_lagjob* jobs, double* out) {
    if (!ctx) return GDK_ERR_ARG;
    if (njobs <= 0 || !jobs || !out) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_lag_sums: bad arguments");
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK(cudaSetDevice(ctx->device));
    std::vector<LagJob> lj(njobs);
    for (int i = 0; i < njobs; i++) {
        const gdk_lagjob& j = jobs[i];
        if (j.param < 0 || j.param >= ctx->P || j.nk < 1 || j.nk > LAG_MAXK || j.k0 < 0 || (j.mode != 0 && j.mode != 1))
            return gdk_fail(ctx, GDK_ERR_ARG, "gdk_lag_sums: bad job %d", i);
        lj[i] = LagJob{j.param, j.mode, (long long)j.k0, j.nk, 0, j.mean, j.inv4s2};
    }
    const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->N + 8191) / 8192, (int64_t)ctx->num_sms * 8 / njobs + 1));
    const int64_t chunk = (ctx->N + nchunk - 1) / nchunk;
    if (ctx->bytes2d.ensure((size_t)njobs * sizeof(LagJob)) || ctx->scratch.ensure((size_t)njobs * nchunk * LAG_MAXK))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "lag-sum buffers");
    LagJob* dj = reinterpret_cast<LagJob*>(ctx->bytes2d.p);
    CK(cudaMemcpyAsync(dj, lj.data(), (size_t)njobs * sizeof(LagJob), cudaMemcpyHostToDevice, ctx->stream));
    dim3 g((unsigned)nchunk, (unsigned)njobs);
    k_lag_sums<<<g, 256, 0, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->dW.p, ctx->N, chunk, dj, ctx->scratch.p);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<double> part((size_t)njobs * nchunk * LAG_MAXK);
    CK(cudaMemcpyAsync(part.data(), ctx->scratch.p, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t o = 0;
    for (int i = 0; i < njobs; i++)
        for (int k = 0; k < jobs[i].nk; k++) {
            double t = 0;
            for (int c = 0; c < nchunk; c++) t += part[((size_t)i * nchunk + c) * LAG_MAXK + k];
            out[o++] = t;
        }
    return GDK_OK;
}

// -------------------------------------------------------------------------------------------------
// measured peaks (roofline denominators that MEASURED_PEAKS.json does not hold)
// -------------------------------------------------------------------------------------------------
extern "C" int32_t gdk_measure_peaks(gdk_ctx* ctx, double* out) {
    if (!ctx || !out) return GDK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto timed = [&](auto&& launch, int reps) -> double {  // best of `reps` after one warm-up launch, in ms
        launch();
        double best = 1e30;
        for (int r = 0; r < reps; r++) {
            cudaEventRecord(e0, ctx->stream);
            launch();
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, (double)ms);
        }
        return best;
    };
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (ctx->scratch.ensure((size_t)ctx->num_sms * 8 * 512 + 64)) return gdk_fail(ctx, GDK_ERR_NOMEM, "scratch");
    {  // FP64 FMA
        const int grid = ctx->num_sms * 4, iters = 8192;
        const double ms = timed([&] { k_peak_fp64<<<grid, 512, 0, ctx->stream>>>(ctx->scratch.p, iters, 1.0000001, 0.9999999); }, 3);
        out[0] = 2.0 * 8 * iters * (double)grid * 512 / (ms * 1e-3) / 1e12;  // TFLOP/s
    }
    {  // shared-memory 64-bit fixed-point updates (ATOMS + carry + RED)
        const int grid = ctx->num_sms, iters = 2048;
        const size_t smem = (size_t)2 * 96 * 96 * 4;
        CK(cudaFuncSetAttribute(k_peak_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unsigned long long* o = reinterpret_cast<unsigned long long*>(ctx->scratch.p);
        for (int mode = 0; mode < 2; mode++) {
            const double ms = timed([&] { k_peak_atoms<<<grid, 512, smem, ctx->stream>>>(o, iters, mode); }, 3);
            out[1 + mode] = 4.0 * iters * (double)grid * 512 / (ms * 1e-3);  // updates/s
        }
    }
    {  // L2 reductions at random addresses of a 32 MB region
        const size_t nb = (size_t)4 << 20;
        if (ctx->gbins2.ensure(nb)) return gdk_fail(ctx, GDK_ERR_NOMEM, "peak buffer");
        CK(cudaMemsetAsync(ctx->gbins2.p, 0, nb * 8, ctx->stream));
        const int grid = ctx->num_sms * 16, iters = 1024;
        const double ms = timed([&] { k_peak_l2red<<<grid, 256, 0, ctx->stream>>>(ctx->gbins2.p, (unsigned)(nb - 1), iters); }, 3);
        out[3] = (double)iters * grid * 256 / (ms * 1e-3);  // reductions/s
    }
    {  // HBM read sweep over 2 GB
        const size_t n = (size_t)1 << 28;  // doubles
        double* buf = nullptr;
        if (cudaMalloc(&buf, n * 8) == cudaSuccess) {
            CK(cudaMemsetAsync(buf, 0, n * 8, ctx->stream));
            const double ms = timed([&] { k_peak_read<<<ctx->num_sms * 8, 512, 0, ctx->stream>>>(buf, (int64_t)(n / 2), ctx->scratch.p); }, 3);
            out[4] = (double)n * 8 / (ms * 1e-3) / 1e9;  // GB/s
            cudaFree(buf);
        } else {
            cudaGetLastError();
        }
    }
    ctx->launches += 20;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    return GDK_OK;
}
