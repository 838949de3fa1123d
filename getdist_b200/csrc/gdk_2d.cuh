// gdk_2d.cuh -- host orchestration of the 2D density path (gdk_density2d_batch / gdk_hist2d_batch).
// Included by gdk.cu.  Per chunk of pairs: histograms (tiled global reductions) -> sheared re-binning ->
// transforms -> bandwidth -> window/mask maps/convolutions/corrections -> max-normalised output.
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "gdk_ctx.h"
#include "host_tables.h"
#include "kernels_2d.cuh"

#define CK2(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return gdk_fail(ctx, GDK_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                            cudaGetErrorString(e_));                                                       \
    } while (0)

struct Arena {
    unsigned char* base = nullptr;
    size_t cap = 0, used = 0;
    void* take(size_t bytes) {
        const size_t a = (used + 255) & ~size_t(255);
        if (a + bytes > cap) return nullptr;
        used = a + bytes;
        return base + a;
    }
};

static size_t bytes_per_pair_estimate(const gdk_spec2d& s, bool likes = false) {
    const size_t g = (size_t)s.fine_bins * s.fine_bins * 8, gb = (size_t)s.base_fine_bins * s.base_fine_bins * 8;
    const size_t w = 256;  // generous window half-width guess for the T scratch
    return g * (likes ? 18 : 13) + gb * 6 + (2 * w + 1) * (size_t)s.fine_bins * 4 * 8 + (2 * w + 1) * (2 * w + 1) * 8;
}

template <class T>
static int upload_vec(gdk_ctx* ctx, const std::vector<T>& v, DevBuf<unsigned char>& buf, T** out) {
    if (buf.ensure(std::max<size_t>(v.size() * sizeof(T), 256))) return gdk_fail(ctx, GDK_ERR_NOMEM, "job table");
    if (!v.empty()) CK2(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = reinterpret_cast<T*>(buf.p);
    return 0;
}

static int density2d_chunk(gdk_ctx* ctx, int n, const gdk_spec2d* specs, double* P_out, const int64_t* offsets,
                           gdk_result2d* res, uint32_t flags, bool hist_only, double* likes_out = nullptr,
                           const double* masks = nullptr, const int64_t* mask_offsets = nullptr, const int32_t* mask_w = nullptr) {
    const bool peers_out = (flags & GDK_OUT_PEERS) != 0 && ctx->nranks > 1;
    const bool dev_out = (flags & (GDK_OUT_DEVICE | GDK_OUT_PEERS)) != 0;
    const bool likes = likes_out != nullptr && !hist_only;
    PeerTable ptab{};
    if (peers_out) {
        if (P_out != ctx->win[GDK_WIN_G2].p || likes) return gdk_fail(ctx, GDK_ERR_ARG, "GDK_OUT_PEERS: P_out must be the GDK_WIN_G2 window");
        for (int p = 0; p < ctx->nranks; p++) {
            if (p == ctx->rank || !((ctx->push_mask >> p) & 1u)) continue;
            if (!ctx->peer_ptr[GDK_WIN_G2][p]) return gdk_fail(ctx, GDK_ERR_STATE, "result window of peer %d is not mapped", p);
            ptab.base[ptab.n++] = (double*)ctx->peer_ptr[GDK_WIN_G2][p];
        }
    }
    // ---------------- layout of the pair grids ----------------
    std::vector<long long> goff(n);
    size_t gtot = 0;
    for (int i = 0; i < n; i++) {
        goff[i] = (long long)gtot;
        gtot += (size_t)specs[i].fine_bins * specs[i].fine_bins;
    }
    std::vector<int> shear_of(n, -1);
    std::vector<ShearJob> sjobs;
    size_t rtot = 0;
    if (!hist_only)
        for (int i = 0; i < n; i++)
            if (specs[i].bw_mode == GDK_BW2D_SHEAR) {
                const gdk_spec2d& s = specs[i];
                ShearJob j{};
                j.pi = s.shear_i;
                j.pj = s.shear_j;
                j.Gb = s.base_fine_bins;
                j.r0 = s.r0;
                j.r1 = s.r1;
                j.p1_min = s.p1_min;
                j.dx1 = (s.p1_max - s.p1_min) / (s.base_fine_bins - 1);
                j.inv1 = 1.0 / j.dx1;
                j.off = (long long)rtot;
                rtot += (size_t)j.Gb * j.Gb;
                shear_of[i] = (int)sjobs.size();
                sjobs.push_back(j);
            }
    if (ctx->gbins2.ensure(gtot) || ctx->gbins_rot.ensure(std::max<size_t>(rtot, 1)))
        return gdk_fail(ctx, GDK_ERR_NOMEM, "2D histogram grids (%zu MB)", (gtot + rtot) * 8 >> 20);
    CK2(cudaMemsetAsync(ctx->gbins2.p, 0, gtot * 8, ctx->stream));
    if (rtot) CK2(cudaMemsetAsync(ctx->gbins_rot.p, 0, rtot * 8, ctx->stream));

    // ---------------- shared-memory privatised path for the 256 x 256 grids ----------------
    std::vector<char> in_bands(n, 0);
    std::vector<int> band_pairs, hot_pairs;
    bool sorted = false;
    {
        // default for 256^2 grids: hot-window privatisation (k_bin8 + k_hist2d_hot); GDK_HOT=0 falls back to REDG tiles
        const size_t hot_smem = (size_t)4 * 2 * HW * HW * 4;
        // default: bucket-sorted sweep (k_bin8c + k_bucket_scatter + k_hist2d_sorted); GDK_SORTED=0 -> hot windows
        if (ctx->use_sorted && !ctx->use_bands && ctx->N >= ctx->sorted_min_n && ctx->N < (int64_t)0xfffffff0u &&
            (size_t)2 * 256 * 32 * 4 + 4096 <= (size_t)ctx->max_smem) {
            for (int i = 0; i < n; i++)
                if (specs[i].fine_bins == 256) {
                    in_bands[i] = 1;
                    hot_pairs.push_back(i);
                }
            sorted = !hot_pairs.empty();
        } else if (ctx->use_hot && !ctx->use_bands && ctx->N >= (1 << 17) && hot_smem + 2048 <= (size_t)ctx->max_smem)
            for (int i = 0; i < n; i++)
                if (specs[i].fine_bins == 256) {
                    in_bands[i] = 1;  // excluded from the REDG tiles
                    hot_pairs.push_back(i);
                }
    }
    {
        // Opt-in (GDK_BANDS=1 at context creation).  Measured on B200 at C2 (profiles/README.md, r1h): 218 ms vs
        // 103 ms for the tiled REDG kernel -- with 128 KB of exact 64-bit bins per CTA the multicast ring is only
        // 80 KB deep, so the stream is bound by the issue->consume->remote-arrive round trip, not by atomics.
        const size_t band_smem = (size_t)2 * 64 * 256 * 4 + (size_t)HB_STAGES * HB_CHUNK * 10;
        if (ctx->use_bands && ctx->cluster_ok && ctx->N >= (1 << 17) && band_smem + 1024 <= (size_t)ctx->max_smem)
            for (int i = 0; i < n; i++)
                if (specs[i].fine_bins == 256) {
                    in_bands[i] = 1;
                    band_pairs.push_back(i);
                }
    }
    // ---------------- histogram tiles, grouped by grid size ----------------
    std::vector<Tile2d> tiles;
    {
        std::map<int, std::vector<int>> byG;
        for (int i = 0; i < n; i++)
            if (!in_bands[i]) byG[specs[i].fine_bins].push_back(i);
        for (auto& kv : byG) {
            const int G = kv.first;
            std::vector<int> A, B;
            for (int i : kv.second) {
                A.push_back(specs[i].px);
                B.push_back(specs[i].py);
            }
            std::sort(A.begin(), A.end());
            A.erase(std::unique(A.begin(), A.end()), A.end());
            std::sort(B.begin(), B.end());
            B.erase(std::unique(B.begin(), B.end()), B.end());
            std::map<int, int> ia, ib;
            for (size_t k = 0; k < A.size(); k++) ia[A[k]] = (int)k;
            for (size_t k = 0; k < B.size(); k++) ib[B[k]] = (int)k;
            std::map<std::pair<int, int>, int> tix;
            for (int i : kv.second) {
                const gdk_spec2d& s = specs[i];
                const int a = ia[s.px], b = ib[s.py];
                const std::pair<int, int> key(a / HT, b / HT);
                auto it = tix.find(key);
                if (it == tix.end()) {
                    Tile2d t{};
                    t.G = G;
                    for (int x = 0; x < HT; x++)
                        for (int y = 0; y < HT; y++) t.off[x][y] = -1;
                    it = tix.emplace(key, (int)tiles.size()).first;
                    tiles.push_back(t);
                }
                Tile2d& t = tiles[it->second];
                const int la = a % HT, lb = b % HT;
                if (t.off[la][lb] >= 0) {
                    // the same (px, py, G) requested twice in one batch: give it its own 1x1 tile
                    Tile2d d{};
                    d.G = G;
                    for (int x = 0; x < HT; x++)
                        for (int y = 0; y < HT; y++) d.off[x][y] = -1;
                    d.na = d.nb = 1;
                    d.pa[0] = s.px;
                    d.pb[0] = s.py;
                    d.amin[0] = s.xbinmin;
                    d.afw[0] = (s.xbinmax - s.xbinmin) / (G - 1);
                    d.ainv[0] = 1.0 / d.afw[0];
                    d.bmin[0] = s.ybinmin;
                    d.bfw[0] = (s.ybinmax - s.ybinmin) / (G - 1);
                    d.binv[0] = 1.0 / d.bfw[0];
                    d.off[0][0] = goff[i];
                    tiles.push_back(d);
                    continue;
                }
                t.na = std::max(t.na, la + 1);
                t.nb = std::max(t.nb, lb + 1);
                t.pa[la] = s.px;
                t.amin[la] = s.xbinmin;
                t.afw[la] = (s.xbinmax - s.xbinmin) / (G - 1);
                t.ainv[la] = 1.0 / t.afw[la];
                t.pb[lb] = s.py;
                t.bmin[lb] = s.ybinmin;
                t.bfw[lb] = (s.ybinmax - s.ybinmin) / (G - 1);
                t.binv[lb] = 1.0 / t.bfw[lb];
                t.off[la][lb] = goff[i];
            }
        }
        // slots of a tile that were never assigned keep pa/pb = 0 with all offsets -1: harmless loads
        for (Tile2d& t : tiles) {
            for (int k = 0; k < t.na; k++)
                if (t.afw[k] == 0) {
                    t.afw[k] = t.ainv[k] = 1;
                }
            for (int k = 0; k < t.nb; k++)
                if (t.bfw[k] == 0) {
                    t.bfw[k] = t.binv[k] = 1;
                }
        }
    }
    Tile2d* dtiles = nullptr;
    int rc = upload_vec(ctx, tiles, ctx->bytes2d, &dtiles);
    if (rc) return rc;
    const int ntiles = (int)tiles.size();
    // the histogram pass over all pairs of the chunk with fixed-point weights WQ into the grids GR
    auto fill_hist = [&](const unsigned long long* WQ, unsigned long long* GR) -> int {
        PhaseTimer pt;
        pt.begin(ctx, GDK_PH_HIST2D);
        const std::vector<int>& b8pairs = hot_pairs.empty() ? band_pairs : hot_pairs;
        if (!b8pairs.empty()) {
            // byte bin indices per parameter (geometry of a parameter is the same in every 256^2 pair)
            std::map<int, int> slot;
            std::vector<Bin8Job> bj;
            auto add = [&](int p, double lo, double hi) {
                if (slot.count(p)) return;
                const double fw = (hi - lo) / 255.0;
                slot[p] = (int)bj.size();
                bj.push_back(Bin8Job{p, 0, lo, fw, 1.0 / fw});
            };
            {
                // slots in ascending parameter order: the circular partner windows of the bucket-sorted sweep are then
                // contiguous byte ranges of the row-major tile whatever order the pairs arrive in
                std::map<int, std::pair<double, double>> first;
                for (int i : b8pairs) {
                    first.emplace(specs[i].px, std::make_pair(specs[i].xbinmin, specs[i].xbinmax));
                    first.emplace(specs[i].py, std::make_pair(specs[i].ybinmin, specs[i].ybinmax));
                }
                for (const auto& kv : first) add(kv.first, kv.second.first, kv.second.second);
            }
            // a parameter must have ONE geometry across the batch; otherwise route the odd pairs through the tiles
            bool consistent = true;
            for (int i : b8pairs) {
                const Bin8Job& a = bj[slot[specs[i].px]];
                const Bin8Job& b = bj[slot[specs[i].py]];
                if (a.binmin != specs[i].xbinmin || a.fw != (specs[i].xbinmax - specs[i].xbinmin) / 255.0 ||
                    b.binmin != specs[i].ybinmin || b.fw != (specs[i].ybinmax - specs[i].ybinmin) / 255.0)
                    consistent = false;
            }
            if (!consistent) return gdk_fail(ctx, GDK_ERR_ARG, "inconsistent grid geometry for one parameter within a batch");
            const int np8 = (int)bj.size();
            if (ctx->ix8.ensure((size_t)np8 * ctx->ld)) return gdk_fail(ctx, GDK_ERR_NOMEM, "byte bin indices");
            Bin8Job* dbj = nullptr;
            rc = upload_vec(ctx, bj, ctx->bytes2d_b, &dbj);
            if (rc) return rc;
            const int64_t want8 = std::max<int64_t>(1, (int64_t)ctx->num_sms * 16 / np8);
            std::vector<Seg> segs8 = gdk_make_segments(ctx, std::max<int64_t>(1 << 14, (ctx->N + want8 - 1) / want8));
            rc = gdk_upload_segs(ctx, segs8, ctx->segs);
            if (rc) return rc;
            dim3 g8((unsigned)segs8.size(), (unsigned)np8);
            if (sorted) {
                // ---- bucket-sorted sweep: byte bins + bucket counts, then per batch of <= 64 (anchor, 32 partners)
                // jobs: counting-sort scatter of 32-byte records + weights, and the conflict-free sweep over them ----
                int pitch = (np8 + 36 + 3) & ~3;
                if (((pitch >> 2) & 1) == 0) pitch += 4;  // odd number of words per row: conflict-free row-strided loads
                int rows = 1024;
                auto rec_smem = [&](int r, int nj) { return (size_t)r * pitch + (size_t)nj * 256 * 8 + (size_t)r * 44 + 2048; };
                while (rows > 32 && rec_smem(rows, SRT_MAXJOBS) + 1024 > (size_t)ctx->max_smem) rows >>= 1;
                if (rec_smem(rows, SRT_MAXJOBS) + 1024 > (size_t)ctx->max_smem)
                    return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "too many parameters (%d) in one 2D batch for the bucket-sorted sweep", np8);
                // circular rule: the anchor of a pair is the parameter from which the other one is at most np8/2
                // slots ahead (mod np8); every anchor gets <= np8/2 partners, 32 per job
                struct Partner { int col, sb, sc; long long off; };
                std::vector<std::vector<Partner>> plist(np8);
                for (int i : hot_pairs) {
                    const int sx = slot[specs[i].px], sy = slot[specs[i].py];
                    const int d = ((sy - sx) % np8 + np8) % np8;
                    bool anchor_x = (2 * d < np8) || (2 * d == np8 && sx < sy) || sx == sy;
                    if (specs[i].anchor_hint == 1) anchor_x = true;
                    if (specs[i].anchor_hint == 2) anchor_x = false;
                    if (anchor_x)
                        plist[sx].push_back(Partner{sy, 256, 1, goff[i]});  // rows of bucket c fill column c: [iy][c]
                    else
                        plist[sy].push_back(Partner{sx, 1, 256, goff[i]});  // rows of bucket c fill row c: [c][ix]
                }
                std::vector<SortJob> sj;
                for (int a = 0; a < np8; a++) {
                    // partners in circular order after the anchor, so that a full triangle gives contiguous windows
                    std::stable_sort(plist[a].begin(), plist[a].end(), [&](const Partner& x, const Partner& y) {
                        return ((x.col - a - 1 + 2 * np8) % np8) < ((y.col - a - 1 + 2 * np8) % np8);
                    });
                    for (size_t k0 = 0; k0 < plist[a].size(); k0 += 32) {
                        SortJob j{};
                        j.slot = a;
                        j.nl = (int)std::min<size_t>(32, plist[a].size() - k0);
                        j.lg = 0;
                        while ((1 << j.lg) < j.nl) j.lg++;
                        for (int l = 0; l < 32; l++) j.pcol[l] = a;
                        for (int l = 0; l < j.nl; l++) {
                            const Partner& q = plist[a][k0 + l];
                            j.pcol[l] = q.col;
                            j.sb[l] = q.sb;
                            j.sc[l] = q.sc;
                            j.off[l] = q.off;
                        }
                        j.c0 = j.pcol[0];
                        for (int l = 0; l < j.nl; l++)
                            if (j.pcol[l] != (j.c0 + l) % np8 || j.c0 + l >= np8 + std::min(32, np8)) j.c0 = -1;
                        sj.push_back(j);
                    }
                }
                // jobs of one launch share their lane layout: order by lanes per row (stable: anchors stay grouped)
                std::stable_sort(sj.begin(), sj.end(), [](const SortJob& x, const SortJob& y) { return x.lg > y.lg; });
                const int njobs = (int)sj.size();
                const int64_t pld = (ctx->N + 31) & ~int64_t(31);
                // jobs per batch: <= SRT_MAXJOBS, and the record scratch (40 B per row and job) within a third of the
                // memory that is free or already held by it (very large N: smaller batches instead of an allocation failure)
                int maxjobs = SRT_MAXJOBS;
                {
                    const size_t have = ctx->recs.cap + ctx->recw.cap * 8;
                    const size_t fit = have >= (size_t)SRT_MAXJOBS * pld * 40 ? (size_t)SRT_MAXJOBS
                                                                              : ((ctx->query_free() + have) / 3) / ((size_t)pld * 40);
                    maxjobs = (int)std::max<size_t>(1, std::min<size_t>((size_t)SRT_MAXJOBS, fit));
                }
                const int nbatch_max = std::min(njobs, maxjobs);
                if (ctx->bucket.ensure((size_t)np8 * (256 + 257) + (size_t)njobs * 256) || ctx->recs.ensure((size_t)nbatch_max * pld * 32) ||
                    ctx->recw.ensure((size_t)nbatch_max * pld))
                    return gdk_fail(ctx, GDK_ERR_NOMEM, "bucket-sorted sweep work space (%zu MB)", ((size_t)nbatch_max * pld * 40) >> 20);
                unsigned* counts = ctx->bucket.p;
                unsigned* start = counts + (size_t)np8 * 256;
                unsigned* cursor = start + (size_t)np8 * 257;
                CK2(cudaMemsetAsync(counts, 0, (size_t)np8 * 256 * 4, ctx->stream));
                {
                    KernelTimer kt(ctx, GDK_K_BIN8C, (double)ctx->N * np8 * 9.0, 0);
                    k_bin8c<<<g8, 256, 0, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->segs.p, dbj, ctx->ix8.p, ctx->ld, counts);
                }
                k_bucket_scan<<<np8, 256, 0, ctx->stream>>>(counts, start);
                ctx->launches += 2;
                SortJob* dsj = nullptr;
                rc = upload_vec(ctx, sj, ctx->bytes2d_s, &dsj);
                if (rc) return rc;
                k_cursor_init<<<njobs, 256, 0, ctx->stream>>>(dsj, start, cursor);
                ctx->launches++;
                const int chunk = 16384;
                const size_t srt_smem = (size_t)2 * 256 * 32 * 4;
                CK2(cudaFuncSetAttribute(k_bucket_records, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_smem(rows, SRT_MAXJOBS)));
                for (int b0 = 0; b0 < njobs; b0 += maxjobs) {
                    const int nb = std::min(maxjobs, njobs - b0);
                    {
                        KernelTimer kt(ctx, GDK_K_BUCKET_RECORDS, (double)ctx->N * (np8 + 8.0 + 40.0 * nb), 0);
                        k_bucket_records<<<(unsigned)((ctx->N + rows - 1) / rows), 1024, rec_smem(rows, nb), ctx->stream>>>(
                            ctx->ix8.p, ctx->ld, np8, pitch, rows, ctx->N, WQ, dsj + b0, nb, cursor + (size_t)b0 * 256,
                            reinterpret_cast<uint4*>(ctx->recs.p), ctx->recw.p, pld);
                    }
                    ctx->launches++;
                    for (int j0 = 0; j0 < nb;) {  // sub-ranges of equal lane layout
                        int j1 = j0;
                        while (j1 < nb && sj[b0 + j1].lg == sj[b0 + j0].lg) j1++;
                        dim3 gh((unsigned)((ctx->N + chunk - 1) / chunk), (unsigned)(j1 - j0));
#define GDK_LAUNCH_REC(LG)                                                                                                   \
    case LG:                                                                                                                 \
        CK2(cudaFuncSetAttribute(k_hist2d_records<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)srt_smem));        \
        k_hist2d_records<LG><<<gh, SRT_THREADS, srt_smem, ctx->stream>>>(dsj + b0, j0, ctx->recs.p, ctx->recw.p, pld, start, \
                                                                         GR, chunk, ctx->N);                      \
        break;
                        {
                            KernelTimer kt(ctx, GDK_K_HIST2D_RECORDS, (double)ctx->N * 40.0 * (j1 - j0), 0);
                            switch (sj[b0 + j0].lg) {
                                GDK_LAUNCH_REC(0)
                                GDK_LAUNCH_REC(1)
                                GDK_LAUNCH_REC(2)
                                GDK_LAUNCH_REC(3)
                                GDK_LAUNCH_REC(4)
                                GDK_LAUNCH_REC(5)
                            }
                        }
#undef GDK_LAUNCH_REC
                        ctx->launches++;
                        j0 = j1;
                    }
                }
            } else {
            k_bin8<<<g8, 256, 0, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->segs.p, dbj, ctx->ix8.p, ctx->ld);
            ctx->launches++;
            }
            if (sorted) {
            } else if (!hot_pairs.empty()) {
                // 2 x 2 tiles over (x parameter, y parameter); window origin from the weighted mean of each parameter
                std::vector<int> A, B;
                for (int i : hot_pairs) {
                    A.push_back(specs[i].px);
                    B.push_back(specs[i].py);
                }
                std::sort(A.begin(), A.end());
                A.erase(std::unique(A.begin(), A.end()), A.end());
                std::sort(B.begin(), B.end());
                B.erase(std::unique(B.begin(), B.end()), B.end());
                std::map<int, int> ia, ib;
                for (size_t k = 0; k < A.size(); k++) ia[A[k]] = (int)k;
                for (size_t k = 0; k < B.size(); k++) ib[B[k]] = (int)k;
                auto origin = [&](int p) {
                    const Bin8Job& j = bj[slot[p]];
                    int c = (int)floor((ctx->means[p] - j.binmin) / j.fw + 0.5) - HW / 2;
                    return std::max(0, std::min(256 - HW, c));
                };
                std::vector<HotTile> ht;
                std::map<std::pair<int, int>, int> tix;
                for (int i : hot_pairs) {
                    const gdk_spec2d& sp = specs[i];
                    const int a = ia[sp.px], b = ib[sp.py];
                    const std::pair<int, int> key(a / 2, b / 2);
                    auto it = tix.find(key);
                    if (it == tix.end()) {
                        HotTile t{};
                        for (int x = 0; x < 2; x++)
                            for (int y = 0; y < 2; y++) t.off[x][y] = -1;
                        t.ia[0] = t.ia[1] = t.ib[0] = t.ib[1] = ctx->ix8.p;
                        it = tix.emplace(key, (int)ht.size()).first;
                        ht.push_back(t);
                    }
                    HotTile* t = &ht[it->second];
                    const int la = a % 2, lb = b % 2;
                    if (t->off[la][lb] >= 0) {  // duplicate request: own tile
                        HotTile d{};
                        for (int x = 0; x < 2; x++)
                            for (int y = 0; y < 2; y++) d.off[x][y] = -1;
                        d.ia[0] = d.ia[1] = d.ib[0] = d.ib[1] = ctx->ix8.p;
                        ht.push_back(d);
                        t = &ht.back();
                        t->na = t->nb = 0;
                        t->off[0][0] = goff[i];
                        t->na = t->nb = 1;
                        t->ia[0] = ctx->ix8.p + (size_t)slot[sp.px] * ctx->ld;
                        t->ib[0] = ctx->ix8.p + (size_t)slot[sp.py] * ctx->ld;
                        t->ax0[0] = origin(sp.px);
                        t->by0[0] = origin(sp.py);
                        continue;
                    }
                    t->na = std::max(t->na, la + 1);
                    t->nb = std::max(t->nb, lb + 1);
                    t->ia[la] = ctx->ix8.p + (size_t)slot[sp.px] * ctx->ld;
                    t->ib[lb] = ctx->ix8.p + (size_t)slot[sp.py] * ctx->ld;
                    t->ax0[la] = origin(sp.px);
                    t->by0[lb] = origin(sp.py);
                    t->off[la][lb] = goff[i];
                }
                HotTile* dht = nullptr;
                rc = upload_vec(ctx, ht, ctx->bytes2d_d, &dht);
                if (rc) return rc;
                // long segments amortise the window flush (4 x 4096 reductions per CTA); segments are the fast index
                const int64_t seglen = (std::max<int64_t>(1 << 16, (ctx->N + 255) / 256) + 3) & ~int64_t(3);
                std::vector<Seg> segsh = gdk_make_segments(ctx, seglen);
                rc = gdk_upload_segs(ctx, segsh, ctx->segs);
                if (rc) return rc;
                const size_t hot_smem = (size_t)4 * 2 * HW * HW * 4;
                CK2(cudaFuncSetAttribute(k_hist2d_hot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_smem));
                dim3 gh((unsigned)segsh.size(), (unsigned)ht.size());
                k_hist2d_hot<<<gh, 1024, hot_smem, ctx->stream>>>(dht, WQ, ctx->segs.p, GR);
                ctx->launches++;
            } else {
            std::vector<BandJob> jobs(band_pairs.size());
            for (size_t k = 0; k < band_pairs.size(); k++) {
                const int i = band_pairs[k];
                jobs[k] = BandJob{ctx->ix8.p + (size_t)slot[specs[i].px] * ctx->ld, ctx->ix8.p + (size_t)slot[specs[i].py] * ctx->ld,
                                  GR + goff[i]};
            }
            BandJob* dband = nullptr;
            rc = upload_vec(ctx, jobs, ctx->bytes2d_d, &dband);
            if (rc) return rc;
            const size_t band_smem = (size_t)2 * 64 * 256 * 4 + (size_t)HB_STAGES * HB_CHUNK * 10;
            CK2(cudaFuncSetAttribute(k_hist2d_bands, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)band_smem));
            k_hist2d_bands<<<(unsigned)(4 * jobs.size()), HB_THREADS, band_smem, ctx->stream>>>(dband, WQ, ctx->N);
            ctx->launches++;
            }
        }
        if (ntiles) {
            // Row segments are the FAST grid index: all CTAs resident at one time (~5 per SM) then work on one or
            // two tiles, whose <= 64 grids (32 MB at 256^2) stay L2-resident for the reductions.  With few segments
            // per tile the resident CTAs span many tiles and every REDG misses L2 (profiles/r1a: 437 GB of DRAM
            // traffic for 2e10 updates).
            const int64_t want = (int64_t)ctx->num_sms * 6;
            const int64_t seglen = std::max<int64_t>(1 << 12, (ctx->N + want - 1) / want);
            std::vector<Seg> segs = gdk_make_segments(ctx, seglen);
            rc = gdk_upload_segs(ctx, segs, ctx->segs);
            if (rc) return rc;
            dim3 g((unsigned)segs.size(), (unsigned)ntiles);
            k_hist2d_tiles<<<g, 256, 0, ctx->stream>>>(ctx->dX.p, ctx->ld, WQ, ctx->segs.p, dtiles, GR);
            ctx->launches++;
        }
        pt.end();
        CK2(cudaGetLastError());
        return 0;
    };
    if (likes) {  // second histogram with the mean-likelihood weights (mcsamples.py:1829-1831), same bin indices
        if (ctx->gbins2l.ensure(gtot)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D mean-likelihood grids");
        CK2(cudaMemsetAsync(ctx->gbins2l.p, 0, gtot * 8, ctx->stream));
        rc = fill_hist(ctx->dWlq.p, ctx->gbins2l.p);
        if (rc) return rc;
        k_u64_to_f64_inplace<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->gbins2l.p, (int64_t)gtot, 1.0 / ctx->wlscale);
        ctx->launches++;
    }
    rc = fill_hist(ctx->dWq.p, ctx->gbins2.p);
    if (rc) return rc;
    // ---------------- sheared re-binning ----------------
    ShearJob* dsj = nullptr;
    ShearGeom* dgeom = nullptr;
    const int nshear = (int)sjobs.size();
    if (nshear) {
        rc = upload_vec(ctx, sjobs, ctx->bytes2d_b, &dsj);
        if (rc) return rc;
        std::vector<int> ord(nshear);
        for (int i = 0; i < nshear; i++) ord[i] = i;
        // default: column-tiled min/max pass (k_shear_minmax_tiled) + hot-window re-binning (k_shear_hist);
        // GDK_SHEAR_SORTED=1: bucket-sorted records for the re-binning as well (measured slower at C2, profiles/r2e)
        const bool shear_tiled = ctx->use_sorted;
        bool shear_sorted = shear_tiled && ctx->shear_sorted && ctx->N >= ctx->sorted_min_n && ctx->N < (int64_t)0xfffffff0u;
        for (int i = 0; i < nshear; i++) shear_sorted = shear_sorted && sjobs[i].Gb == 256;
        bool geom_done = false;
        auto keyless = [&](int a, int b) {
            const ShearJob &A = sjobs[a], &B = sjobs[b];
            if (A.pi != B.pi) return A.pi < B.pi;
            if (A.Gb != B.Gb) return A.Gb < B.Gb;
            if (A.p1_min != B.p1_min) return A.p1_min < B.p1_min;
            if (A.dx1 != B.dx1) return A.dx1 < B.dx1;
            return A.pj < B.pj;
        };
        std::sort(ord.begin(), ord.end(), keyless);
        if (shear_tiled) {
            // ---- column-tiled passes (kernels_2d.cuh, "bucket-sorted sweep for the sheared re-binning") ----
            // anchors = (p1 column, p1 geometry) with <= SHR_MAXCOLS-1 partners each, in column order
            struct Anchor { int pi; double p1_min, dx1, inv1; std::vector<int> jobs; };
            std::vector<Anchor> anchors;
            for (int k = 0; k < nshear; k++) {
                const ShearJob& j = sjobs[ord[k]];
                bool fresh = anchors.empty();
                if (!fresh) {
                    const Anchor& a = anchors.back();
                    fresh = (int)a.jobs.size() == SHR_MAXCOLS - 1 || a.pi != j.pi || a.p1_min != j.p1_min || a.dx1 != j.dx1;
                }
                if (fresh) anchors.push_back(Anchor{j.pi, j.p1_min, j.dx1, j.inv1, {}});
                anchors.back().jobs.push_back(ord[k]);
            }
            // batches of consecutive anchors whose columns fit one shared-memory tile
            std::vector<ShearBatch> batches;
            std::vector<ShearRecJob> rjobs;
            std::vector<ShearPairRef> prefs;
            std::vector<SortJob> sjs;
            {
                std::vector<int> cols;
                auto col_of = [&](int p) {
                    for (size_t c = 0; c < cols.size(); c++)
                        if (cols[c] == p) return (int)c;
                    return -1;
                };
                ShearBatch cur{};
                auto close = [&]() {
                    if (!cur.njobs) return;
                    cur.ncols = (int)cols.size();
                    for (size_t c = 0; c < cols.size(); c++) cur.cols[c] = cols[c];
                    batches.push_back(cur);
                    cols.clear();
                };
                for (size_t ai = 0; ai < anchors.size(); ai++) {
                    const Anchor& a = anchors[ai];
                    std::vector<int> add;
                    auto want = [&](int p) {
                        if (col_of(p) < 0 && std::find(add.begin(), add.end(), p) == add.end()) add.push_back(p);
                    };
                    want(a.pi);
                    for (int jb : a.jobs) want(sjobs[jb].pj);
                    const int nu = ((int)a.jobs.size() + 3) / 4;  // units of <= 4 pairs (k_shear_minmax_tma)
                    if (cur.njobs && (cols.size() + add.size() > (size_t)SHR_MAXCOLS || cur.njobs == SHR_MAXJOBS ||
                                      cur.npairs + (int)a.jobs.size() > 16 * SHR_MAXPW || cur.nunits + nu > SHR_MAXUNITS)) {
                        close();
                        add.clear();
                        want(a.pi);
                        for (int jb : a.jobs) want(sjobs[jb].pj);
                    }
                    if (!cur.njobs || cols.empty()) {
                        cur = ShearBatch{};
                        cur.job0 = (int)rjobs.size();
                        cur.pair0 = (int)prefs.size();
                    }
                    for (int p : add) cols.push_back(p);
                    ShearRecJob rj{};
                    rj.acol = col_of(a.pi);
                    rj.np = (int)a.jobs.size();
                    rj.pair0 = (int)prefs.size() - cur.pair0;
                    for (int u = 0; u < nu; u++) {  // balanced split: 9 partners -> 3 + 3 + 3
                        const int cnt = (int)a.jobs.size(), b0 = cnt * u / nu, b1 = cnt * (u + 1) / nu;
                        cur.ufirst[cur.nunits] = (short)(rj.pair0 + b0);
                        cur.unp[cur.nunits] = (unsigned char)(b1 - b0);
                        cur.uacol[cur.nunits] = (unsigned char)rj.acol;
                        cur.nunits++;
                    }
                    rj.p1_min = a.p1_min;
                    rj.dx1 = a.dx1;
                    rj.inv1s = a.inv1 * 1048576.0;
                    SortJob sj{};
                    sj.slot = (int)rjobs.size();
                    sj.nl = rj.np;
                    while ((1 << sj.lg) < sj.nl) sj.lg++;
                    sj.c0 = -1;
                    for (int l = 0; l < rj.np; l++) {
                        const ShearJob& j = sjobs[a.jobs[l]];
                        prefs.push_back(ShearPairRef{rj.acol, col_of(j.pj), a.jobs[l], cur.njobs, j.r0, j.r1});
                        sj.sb[l] = 256;  // rot grid [b2][b1]
                        sj.sc[l] = 1;
                        sj.off[l] = j.off;
                    }
                    rjobs.push_back(rj);
                    sjs.push_back(sj);
                    cur.njobs++;
                    cur.npairs += rj.np;
                }
                close();
            }
            const int nanch = (int)rjobs.size(), nbatch = (int)batches.size();
            ShearBatch* dbat = nullptr;
            ShearRecJob* drj = nullptr;
            ShearPairRef* dpr = nullptr;
            SortJob* dsortj = nullptr;
            rc = upload_vec(ctx, batches, ctx->bytes2d_c, &dbat);
            if (rc) return rc;
            rc = upload_vec(ctx, rjobs, ctx->bytes2d_sh1, &drj);
            if (rc) return rc;
            rc = upload_vec(ctx, prefs, ctx->bytes2d_sh2, &dpr);
            if (rc) return rc;
            rc = upload_vec(ctx, sjs, ctx->bytes2d_s, &dsortj);
            if (rc) return rc;
            const int64_t want = std::max<int64_t>(1, (int64_t)ctx->num_sms * 8 / nbatch);
            int64_t seglen = std::max<int64_t>(16 * SHR_ROWS, (ctx->N + want - 1) / want);
            seglen = (seglen + SHR_ROWS - 1) / SHR_ROWS * SHR_ROWS;
            std::vector<Seg> segs = gdk_make_segments(ctx, seglen);
            rc = gdk_upload_segs(ctx, segs, ctx->segs);
            if (rc) return rc;
            const int nseg = (int)segs.size();
            const int64_t pld = (ctx->N + 31) & ~int64_t(31);
            int maxjobs = 0;
            size_t mm_smem = 0, rec_smem = 0;
            for (const ShearBatch& b : batches) {
                maxjobs = std::max(maxjobs, b.njobs);
                mm_smem = std::max(mm_smem, (size_t)b.ncols * SHR_ROWS * 8 + (size_t)b.npairs * sizeof(ShearPairRef) + (size_t)b.njobs * 1024);
                rec_smem = std::max(rec_smem, (size_t)SHR_ROWS * 44 + 2048 + (size_t)b.njobs * (2048 + SHR_ROWS) + 32 * sizeof(ShearLaneParam) +
                                                  (size_t)b.ncols * SHR_ROWS * 8);
            }
            if (ctx->scratch.ensure((size_t)nshear * nseg * 2 + (size_t)nshear * 4 + 8) || ctx->bucket2.ensure((size_t)nanch * (256 + 257 + 256)) ||
                (shear_sorted && (ctx->recs.ensure((size_t)maxjobs * pld * 32) || ctx->recw.ensure((size_t)maxjobs * pld))))
                return gdk_fail(ctx, GDK_ERR_NOMEM, "sheared re-binning work space");
            double* part = ctx->scratch.p;
            dgeom = reinterpret_cast<ShearGeom*>(ctx->scratch.p + (((size_t)nshear * nseg * 2 + 3) & ~size_t(3)));
            unsigned* counts = ctx->bucket2.p;
            unsigned* start = counts + (size_t)nanch * 256;
            unsigned* cursor = start + (size_t)nanch * 257;
            CK2(cudaFuncSetAttribute(k_shear_minmax_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(mm_smem, 48 << 10)));
            CK2(cudaFuncSetAttribute(k_shear_records, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(rec_smem, 48 << 10)));
            PhaseTimer pt;
            pt.begin(ctx, GDK_PH_SHEAR);
            CK2(cudaMemsetAsync(counts, 0, (size_t)nanch * 256 * 4, ctx->stream));
            dim3 gm((unsigned)nseg, (unsigned)nbatch);
            {
                double colsum = 0;
                size_t tma_smem = 0;
                for (const ShearBatch& b : batches) {
                    colsum += b.ncols;
                    tma_smem = std::max(tma_smem, (size_t)SHM_STAGES * b.ncols * SHR_ROWS * 8 + (size_t)b.npairs * sizeof(ShearPairRef));
                }
                bool aligned = !shear_sorted && tma_smem + 2048 <= (size_t)ctx->max_smem;  // bulk copies need 16-byte aligned tiles
                for (const Seg& sgm : segs) aligned = aligned && ((sgm.r0 & 1) == 0);
                KernelTimer kt(ctx, GDK_K_SHEAR_MINMAX, (double)ctx->N * colsum * 8.0, (double)ctx->N * nshear * 3.0);
                if (aligned) {
                    CK2(cudaFuncSetAttribute(k_shear_minmax_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(tma_smem, 48 << 10)));
                    k_shear_minmax_tma<<<gm, 544, tma_smem, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->segs.p, nseg, dbat, dpr, part);
                } else {
                    k_shear_minmax_tiled<<<gm, SHR_ROWS, mm_smem, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->segs.p, nseg, dbat, drj, dpr, part,
                                                                                 shear_sorted ? counts : nullptr);
                }
            }
            k_shear_geom<<<(nshear + 127) / 128, 128, 0, ctx->stream>>>(part, nseg, nshear, dsj, dgeom);
            ctx->launches += 2;
            geom_done = true;
            if (shear_sorted) {
                k_bucket_scan<<<nanch, 256, 0, ctx->stream>>>(counts, start);
                k_cursor_init<<<nanch, 256, 0, ctx->stream>>>(dsortj, start, cursor);
                ctx->launches += 2;
            }
            const int chunk = 16384;
            const size_t srt_smem = (size_t)2 * 256 * 32 * 4;
            for (int b = 0; b < nbatch && shear_sorted; b++) {
                const ShearBatch& B = batches[b];
                k_shear_records<<<(unsigned)((ctx->N + SHR_ROWS - 1) / SHR_ROWS), SHR_ROWS, rec_smem, ctx->stream>>>(
                    ctx->dX.p, ctx->ld, ctx->N, ctx->dWq.p, dbat, b, drj, dpr, dgeom, cursor + (size_t)B.job0 * 256,
                    reinterpret_cast<uint4*>(ctx->recs.p), ctx->recw.p, pld);
                ctx->launches++;
                for (int j0 = 0; j0 < B.njobs;) {  // sub-ranges of equal lane layout
                    int j1 = j0;
                    while (j1 < B.njobs && sjs[B.job0 + j1].lg == sjs[B.job0 + j0].lg) j1++;
                    dim3 gh((unsigned)((ctx->N + chunk - 1) / chunk), (unsigned)(j1 - j0));
#define GDK_LAUNCH_REC(LG)                                                                                                       \
    case LG:                                                                                                                     \
        CK2(cudaFuncSetAttribute(k_hist2d_records<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)srt_smem));            \
        k_hist2d_records<LG><<<gh, SRT_THREADS, srt_smem, ctx->stream>>>(dsortj + B.job0, j0, ctx->recs.p, ctx->recw.p, pld, start, \
                                                                         ctx->gbins_rot.p, chunk, ctx->N);                       \
        break;
                    switch (sjs[B.job0 + j0].lg) {
                        GDK_LAUNCH_REC(0)
                        GDK_LAUNCH_REC(1)
                        GDK_LAUNCH_REC(2)
                        GDK_LAUNCH_REC(3)
                        GDK_LAUNCH_REC(4)
                        GDK_LAUNCH_REC(5)
                    }
#undef GDK_LAUNCH_REC
                    ctx->launches++;
                    j0 = j1;
                }
            }
            pt.end();
            CK2(cudaGetLastError());
        }
        if (!shear_sorted) {
        // groups of jobs sharing (p1 column, p1 geometry, grid size): x_i is read once per row for the group.  An anchor's
        // partners are split into balanced groups of <= shear_np jobs (9 partners -> 3 + 3 + 3): the fewer jobs a group
        // has, the larger the shared-memory window each of them gets (k_shear_hist_w)
        std::vector<ShearGroup> sgroups;
        const int npmax = std::max(1, std::min(SG, ctx->shear_np));
        for (int k = 0; k < nshear;) {
            const ShearJob& j0 = sjobs[ord[k]];
            int e = k;
            while (e < nshear && sjobs[ord[e]].pi == j0.pi && sjobs[ord[e]].Gb == j0.Gb && sjobs[ord[e]].p1_min == j0.p1_min &&
                   sjobs[ord[e]].dx1 == j0.dx1)
                e++;
            const int cnt = e - k, ng = (cnt + npmax - 1) / npmax;
            for (int gi = 0; gi < ng; gi++) {
                const int b0 = k + (int)((long long)cnt * gi / ng), b1 = k + (int)((long long)cnt * (gi + 1) / ng);
                ShearGroup g{};
                g.pi = j0.pi;
                g.Gb = j0.Gb;
                g.p1_min = j0.p1_min;
                g.dx1 = j0.dx1;
                g.inv1 = j0.inv1;
                g.mean1 = ctx->means[j0.pi];
                for (int q = b0; q < b1; q++) {
                    const ShearJob& j = sjobs[ord[q]];
                    g.mean2[g.nj] = j.r0 * ctx->means[j.pi] + j.r1 * ctx->means[j.pj];
                    g.pj[g.nj] = j.pj;
                    g.job[g.nj] = ord[q];
                    g.r0[g.nj] = j.r0;
                    g.r1[g.nj] = j.r1;
                    g.off[g.nj] = j.off;
                    g.nj++;
                }
                sgroups.push_back(g);
            }
            k = e;
        }
        // one launch per group size: order by size (stable: neighbouring anchors stay neighbours -> shared columns hit L2)
        std::stable_sort(sgroups.begin(), sgroups.end(), [](const ShearGroup& a, const ShearGroup& b) { return a.nj > b.nj; });
        ShearGroup* dsg = nullptr;
        rc = upload_vec(ctx, sgroups, ctx->bytes2d_c, &dsg);
        if (rc) return rc;
        const int ngroups = (int)sgroups.size();
        PhaseTimer pt;
        if (!geom_done) {
            const int64_t want = std::max<int64_t>((int64_t)ctx->num_sms * 8 / ngroups + 1, std::min<int64_t>((int64_t)ctx->num_sms * 6, 64));
            const int64_t seglen = std::max<int64_t>(1 << 16, (ctx->N + want - 1) / want);
            std::vector<Seg> segs = gdk_make_segments(ctx, seglen);
            rc = gdk_upload_segs(ctx, segs, ctx->segs);
            if (rc) return rc;
            const int nseg = (int)segs.size();
            if (ctx->scratch.ensure((size_t)nshear * nseg * 2 + (size_t)nshear * 4 + 8)) return gdk_fail(ctx, GDK_ERR_NOMEM, "shear scratch");
            double* part = ctx->scratch.p;
            dgeom = reinterpret_cast<ShearGeom*>(ctx->scratch.p + (((size_t)nshear * nseg * 2 + 3) & ~size_t(3)));
            pt.begin(ctx, GDK_PH_SHEAR);
            dim3 g((unsigned)nseg, (unsigned)ngroups);
            k_shear_minmax<<<g, 256, 0, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->segs.p, nseg, dsg, part);
            k_shear_geom<<<(nshear + 127) / 128, 128, 0, ctx->stream>>>(part, nseg, nshear, dsj, dgeom);
            ctx->launches += 2;
        } else {
            pt.resume(ctx, GDK_PH_SHEAR);
        }
        // long segments amortise the window flush (<= NP * W^2 reductions per CTA)
        const int64_t wantseg = std::max<int64_t>(1, std::min<int64_t>(ctx->N / (1 << 18), 4096));
        std::vector<Seg> hsegs = gdk_make_segments(ctx, (ctx->N + wantseg - 1) / wantseg);
        rc = gdk_upload_segs(ctx, hsegs, ctx->segs2);
        if (rc) return rc;
        Seg* dhsegs = ctx->segs2.p;
        const int nhseg = (int)hsegs.size();
        for (int g0 = 0; g0 < ngroups;) {
            int g1 = g0;
            while (g1 < ngroups && sgroups[g1].nj == sgroups[g0].nj) g1++;
            const int nj = sgroups[g0].nj;
            // algorithmic bytes of the launch: the MINIMUM sweep -- every distinct column its groups touch once, plus the
            // weights (not what the grouping happens to re-read: groups share columns, the rest comes out of L2)
            std::vector<char> seen(ctx->P, 0);
            double ncols = 1.0;
            for (int q = g0; q < g1; q++) {
                if (!seen[sgroups[q].pi]) seen[sgroups[q].pi] = 1, ncols += 1.0;
                for (int k = 0; k < sgroups[q].nj; k++)
                    if (!seen[sgroups[q].pj[k]]) seen[sgroups[q].pj[k]] = 1, ncols += 1.0;
            }
            KernelTimer kt(ctx, GDK_K_SHEAR_HIST, (double)ctx->N * ncols * 8.0, (double)ctx->N * (g1 - g0) * nj * 3.0);
            dim3 g((unsigned)(g1 - g0), (unsigned)nhseg);
#define GDK_LAUNCH_SHW(NP_, W_)                                                                                               \
    case NP_: {                                                                                                               \
        const size_t sm_ = (size_t)NP_ * 2 * (W_ * W_ + 4) * 4;                                                               \
        CK2(cudaFuncSetAttribute(k_shear_hist_w<NP_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_));            \
        k_shear_hist_w<NP_, W_><<<g, 512, sm_, ctx->stream>>>(ctx->dX.p, ctx->ld, ctx->dWq.p, dhsegs, dsg + g0, dgeom,        \
                                                               ctx->gbins_rot.p);                                            \
    } break;
            switch (nj) {
                GDK_LAUNCH_SHW(1, 128)
                GDK_LAUNCH_SHW(2, 112)
                GDK_LAUNCH_SHW(3, 96)
                GDK_LAUNCH_SHW(4, 80)
                GDK_LAUNCH_SHW(5, 72)
                GDK_LAUNCH_SHW(6, 64)
            }
#undef GDK_LAUNCH_SHW
            ctx->launches++;
            g0 = g1;
        }
        pt.end();
        CK2(cudaGetLastError());
        }
    }
    // fixed point -> float64, in place
    k_u64_to_f64_inplace<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->gbins2.p, (int64_t)gtot, 1.0 / ctx->wscale);
    if (rtot) k_u64_to_f64_inplace<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->gbins_rot.p, (int64_t)rtot, 1.0 / ctx->wscale);
    ctx->launches += rtot ? 2 : 1;
    const double* H = reinterpret_cast<const double*>(ctx->gbins2.p);
    const double* Hrot = reinterpret_cast<const double*>(ctx->gbins_rot.p);
    if (hist_only) {
        for (int i = 0; i < n; i++)
            CK2(cudaMemcpyAsync(P_out + offsets[i], H + goff[i], (size_t)specs[i].fine_bins * specs[i].fine_bins * 8,
                                cudaMemcpyDeviceToHost, ctx->stream));
        CK2(cudaStreamSynchronize(ctx->stream));
        return 0;
    }

    // ---------------- arena for everything grid-sized below ----------------
    size_t need = 0;
    for (int i = 0; i < n; i++) need += bytes_per_pair_estimate(specs[i], likes);
    need += (size_t)64 << 20;
    if (ctx->bytes_arena.ensure(need)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D work arena (%zu MB)", need >> 20);
    Arena ar{ctx->bytes_arena.p, ctx->bytes_arena.cap, 0};
    auto take_d = [&](size_t cnt) { return reinterpret_cast<double*>(ar.take(cnt * 8)); };

    // ---------------- transforms ----------------
    std::vector<XformJob> xjobs;
    std::vector<Bw2dJob> bjobs(n);
    int Gmax_opt = 0;
    for (int i = 0; i < n; i++) {
        const gdk_spec2d& s = specs[i];
        bjobs[i] = Bw2dJob{nullptr, nullptr, 0, -1};
        if (s.bw_mode != GDK_BW2D_PLAIN && s.bw_mode != GDK_BW2D_SHEAR) continue;
        const bool shear = s.bw_mode == GDK_BW2D_SHEAR;
        const int G = shear ? s.base_fine_bins : s.fine_bins;
        const bool has_limits = s.x_has_bot || s.x_has_top || s.y_has_bot || s.y_has_top;
        const Kde1dTablesHost* t = gdk_tables_for(ctx, G);
        if (!t) return gdk_fail(ctx, GDK_ERR_NOMEM, "transform tables");
        XformJob x{};
        x.src = shear ? Hrot + sjobs[shear_of[i]].off : H + goff[i];
        x.G = G;
        x.a2 = take_d((size_t)G * G);
        x.aFFT = has_limits ? nullptr : take_d((size_t)G * G);
        x.tw = t->tw;
        x.tw4 = t->tw4;
        x.cos4 = t->cos4;
        x.twn = t->twn;
        if (!x.a2 || (!has_limits && !x.aFFT)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D arena exhausted (a2)");
        bjobs[i] = Bw2dJob{x.a2, x.aFFT, G, shear ? shear_of[i] : -1};
        Gmax_opt = std::max(Gmax_opt, G);
        xjobs.push_back(x);
    }
    const size_t mark = ar.used;
    if (!xjobs.empty()) {
        // scratch that is only alive during the transforms (released afterwards)
        for (XformJob& x : xjobs) {
            x.tmpD = take_d((size_t)x.G * x.G);
            x.tmpC = x.aFFT ? reinterpret_cast<cplx*>(ar.take((size_t)x.G * x.G * 16)) : nullptr;
            if (!x.tmpD || (x.aFFT && !x.tmpC)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D arena exhausted (transform scratch)");
        }
        // launches are per grid size (shared-memory footprint and grid shape depend on G)
        std::sort(xjobs.begin(), xjobs.end(), [](const XformJob& a, const XformJob& b) { return a.G < b.G; });
        XformJob* dx = nullptr;
        rc = upload_vec(ctx, xjobs, ctx->bytes2d_c, &dx);
        if (rc) return rc;
        PhaseTimer pt;
        pt.begin(ctx, GDK_PH_XFORM2D);
        k_grid_totals<<<(unsigned)xjobs.size(), 256, 0, ctx->stream>>>(dx);
        ctx->launches++;
        size_t b = 0;
        while (b < xjobs.size()) {
            size_t e = b;
            while (e < xjobs.size() && xjobs[e].G == xjobs[b].G) e++;
            const int G = xjobs[b].G;
            const int lines = std::max(1, std::min(8, (int)((size_t)(ctx->max_smem - 2048) / ((size_t)G * 48))));
            const size_t smem_r = (size_t)lines * G * 40, smem_c = (size_t)lines * G * 48;
            if (smem_c > (size_t)ctx->max_smem) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "grid size %d too large for the transform kernels", G);
            CK2(cudaFuncSetAttribute(k_xform_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_r, 48 << 10)));
            CK2(cudaFuncSetAttribute(k_xform_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_c, 48 << 10)));
            dim3 g((unsigned)((G + lines - 1) / lines), (unsigned)(e - b));
            {
                int nfft = 0;
                for (size_t q = b; q < e; q++) nfft += xjobs[q].aFFT ? 1 : 0;
                const double G2 = (double)G * G, lg = log2((double)G);
                // rows: read the histogram, write the real (+ complex) half-transformed lines; cols: read them, write a2 (+ aFFT)
                KernelTimer kt(ctx, GDK_K_XFORM_ROWS, ((double)(e - b) * 16.0 + nfft * 16.0) * G2, ((double)(e - b) + nfft) * 5.0 * G2 * lg);
                k_xform_rows<<<g, 256, smem_r, ctx->stream>>>(dx + b, lines);
            }
            {
                int nfft = 0;
                for (size_t q = b; q < e; q++) nfft += xjobs[q].aFFT ? 1 : 0;
                const double G2 = (double)G * G, lg = log2((double)G);
                KernelTimer kt(ctx, GDK_K_XFORM_COLS, ((double)(e - b) * 16.0 + nfft * 24.0) * G2, ((double)(e - b) + nfft) * 5.0 * G2 * lg);
                k_xform_cols<<<g, 256, smem_c, ctx->stream>>>(dx + b, lines);
            }
            ctx->launches += 2;
            b = e;
        }
        pt.end();
        CK2(cudaGetLastError());
    }
    // ---------------- bandwidth ----------------
    gdk_spec2d* dspecs = nullptr;
    Bw2dJob* dbj = nullptr;
    {
        std::vector<gdk_spec2d> sv(specs, specs + n);
        rc = upload_vec(ctx, sv, ctx->bytes2d_d, &dspecs);
        if (rc) return rc;
        rc = upload_vec(ctx, bjobs, ctx->bytes2d_e, &dbj);
        if (rc) return rc;
    }
    if (ctx->bytes2d_res.ensure((size_t)n * sizeof(gdk_result2d))) return gdk_fail(ctx, GDK_ERR_NOMEM, "result buffer");
    gdk_result2d* dres = reinterpret_cast<gdk_result2d*>(ctx->bytes2d_res.p);
    {
        const int Gm = std::max(Gmax_opt, 8);
        size_t smem = std::max<size_t>((size_t)2 * PSI_MAXE * Gm * 8, 1024);
        if (smem > (size_t)ctx->max_smem) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "optimiser grid too large");
        // row ring of the a2 sweeps: stages of about 16 KB, as long as two CTAs still fit an SM
        int ring_rows = std::max(1, std::min(8, 2048 / Gm));
        if (smem + (size_t)COOP_RING_STAGES * ring_rows * Gm * 8 > (size_t)100 << 10) ring_rows = 0;
        smem += (size_t)COOP_RING_STAGES * ring_rows * Gm * 8;
        CK2(cudaFuncSetAttribute(k_bw2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 << 10)));
        PhaseTimer pt;
        pt.begin(ctx, GDK_PH_BW2D);
        {
            double abytes = 0;  // a2 (+ |fft2|^2) of every optimised pair read once
            for (int i = 0; i < n; i++)
                if (bjobs[i].a2) abytes += (double)bjobs[i].G * bjobs[i].G * 8.0 * (bjobs[i].aFFT ? 2 : 1);
            KernelTimer kt(ctx, GDK_K_BW2D, abytes, 0);
            k_bw2d<<<n, ctx->bw2d_threads, smem, ctx->stream>>>(dspecs, dbj, dgeom, ctx->k2d, dres, Gm, ring_rows);
        }
        ctx->launches++;
        pt.end();
        CK2(cudaGetLastError());
    }
    CK2(cudaMemcpyAsync(res, dres, (size_t)n * sizeof(gdk_result2d), cudaMemcpyDeviceToHost, ctx->stream));
    CK2(cudaStreamSynchronize(ctx->stream));
    if (flags & GDK_BW_ONLY) return 0;
    // user prior masks (mask_function): checked against the kernel half-width, then made resident
    std::vector<long long> umoff(n, -1);
    if (masks && mask_offsets) {
        size_t utot = 0;
        for (int i = 0; i < n; i++) {
            if (mask_offsets[i] < 0) continue;
            if (!mask_w || mask_w[i] != res[i].winw)
                return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: the mask was built for half-width %d but the kernel half-width is %d", i,
                                mask_w ? mask_w[i] : -1, res[i].winw);
            if (specs[i].x_periodic || specs[i].y_periodic)
                return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "pair %d: mask_function with periodic parameters", i);
            umoff[i] = (long long)utot;
            const size_t gp = (size_t)specs[i].fine_bins + 2 * (size_t)res[i].winw;
            utot += gp * gp;
        }
        if (ctx->umask.ensure(std::max<size_t>(utot, 1))) return gdk_fail(ctx, GDK_ERR_NOMEM, "prior masks");
        for (int i = 0; i < n; i++) {
            if (umoff[i] < 0) continue;
            const size_t gp = (size_t)specs[i].fine_bins + 2 * (size_t)res[i].winw;
            CK2(cudaMemcpyAsync(ctx->umask.p + umoff[i], masks + mask_offsets[i], gp * gp * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
    }

    // ---------------- convolution stage ----------------
    ar.used = mark;  // transform scratch is dead
    std::vector<ConvJob> cj(n);
    int wmax_all = 0;
    if (ctx->bytes2d_mx.ensure((size_t)n * 16 * 8)) return gdk_fail(ctx, GDK_ERR_NOMEM, "max buffer");
    CK2(cudaMemsetAsync(ctx->bytes2d_mx.p, 0, (size_t)n * 128, ctx->stream));
    int max_mbc = 0;
    bool any_periodic = false;
    for (int i = 0; i < n; i++) {
        const gdk_spec2d& s = specs[i];
        const int G = s.fine_bins, w = res[i].winw, K = 2 * w + 1;
        if (w > 384) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "pair %d: kernel half-width %d exceeds the supported 384 bins", i, w);
        if (s.mult_bias_correction_order > 6) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "mult_bias_correction_order > 6");
        const bool has_prior = s.x_has_bot || s.x_has_top || s.y_has_bot || s.y_has_top || umoff[i] >= 0;  // mcsamples.py:1794
        ConvJob& c = cj[i];
        c = ConvJob{};
        c.umask = umoff[i] >= 0 ? ctx->umask.p + umoff[i] : nullptr;
        c.hist = H + goff[i];
        c.G = G;
        c.w = w;
        c.rx = res[i].rx;
        c.ry = res[i].ry;
        c.c = res[i].c;
        if (s.bw_mode == GDK_BW2D_FIXED) c.c = s.kernel_corr;
        c.xper = s.x_periodic ? 1 : 0;
        c.yper = s.y_periodic ? 1 : 0;
        const bool both_per = c.xper && c.yper;  // mcsamples.py:1921, 1963: both blocks are skipped
        c.bounded = (has_prior && s.boundary_correction_order >= 0 && !both_per) ? 1 : 0;
        c.bco = s.boundary_correction_order;
        c.mbc = both_per ? 0 : s.mult_bias_correction_order;
        any_periodic = any_periodic || c.xper || c.yper;
        c.nc = std::max(0, std::min(4, s.n_contours));
        for (int k = 0; k < 4; k++) c.contours[k] = s.contours[k];
        c.xb = s.x_has_bot;
        c.xt = s.x_has_top;
        c.yb = s.y_has_bot;
        c.yt = s.y_has_top;
        c.Wk = take_d((size_t)K * K);
        c.P = take_d((size_t)G * G);
        c.Pn = c.mbc ? take_d((size_t)G * G) : c.P;
        c.a00b = c.mbc ? take_d((size_t)G * G) : nullptr;
        c.box = c.mbc ? take_d((size_t)G * G) : nullptr;
        c.T = ((c.mbc || c.bounded) && !c.umask) ? take_d((size_t)K * G * 4) : nullptr;
        if (c.bounded) {
            c.maps = take_d((size_t)G * G * 6);
            if (c.bco == 1) {
                c.xP = take_d((size_t)G * G);
                c.yP = take_d((size_t)G * G);
            }
        }
        c.mx = reinterpret_cast<unsigned long long*>(ctx->bytes2d_mx.p) + (size_t)i * 16;
        if (likes) {
            c.lhist = reinterpret_cast<const double*>(ctx->gbins2l.p) + goff[i];
            c.lmbc = s.mult_bias_correction_order;
            c.lP = take_d((size_t)G * G);
            c.lP2 = take_d((size_t)G * G);
            c.lbox = c.lmbc ? take_d((size_t)G * G) : nullptr;
            if (!c.lP || !c.lP2 || (c.lmbc && !c.lbox)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D arena exhausted (mean likelihoods, pair %d)", i);
        }
        if (!c.Wk || !c.P || !c.Pn || (c.mbc && (!c.a00b || (!c.T && !c.umask) || !c.box)) ||
            (c.bounded && (!c.maps || (!c.T && !c.umask) || (c.bco == 1 && (!c.xP || !c.yP)))))
            return gdk_fail(ctx, GDK_ERR_NOMEM, "2D arena exhausted (convolution stage, pair %d)", i);
        wmax_all = std::max(wmax_all, w);
        max_mbc = std::max(max_mbc, c.mbc);
    }
    // sort jobs by window size so that each launch group uses a tight shared-memory footprint
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cj[a].w < cj[b].w; });
    std::vector<ConvJob> cjs(n);
    // offs_sorted: where the caller wants density k (elements from P_out: any layout, e.g. the gathered layout of a
    // multi-rank call); devo_sorted: where it sits on the device -- the same for device output, packed in launch order
    // in the staging buffer for host output
    std::vector<long long> offs_sorted(n), devo_sorted(n);
    size_t otot = 0;
    for (int k = 0; k < n; k++) {
        cjs[k] = cj[order[k]];
        offs_sorted[k] = offsets[order[k]];
        devo_sorted[k] = dev_out ? offs_sorted[k] : (long long)otot;
        otot += (size_t)cjs[k].G * cjs[k].G;
    }
    ConvJob* dcj = nullptr;
    rc = upload_vec(ctx, cjs, ctx->bytes2d_c, &dcj);
    if (rc) return rc;
    long long* doffs = nullptr;
    rc = upload_vec(ctx, devo_sorted, ctx->bytes2d_e, &doffs);
    if (rc) return rc;
    int* dcnt = nullptr;
    if (peers_out) {
        std::vector<int> cnts(n);
        for (int k = 0; k < n; k++) cnts[k] = cjs[k].G * cjs[k].G;
        rc = upload_vec(ctx, cnts, ctx->bytes_push, &dcnt);
        if (rc) return rc;
    }
    // result structs in sorted order for the finalize kernel's status bit
    std::vector<gdk_result2d> res_sorted(n);
    for (int k = 0; k < n; k++) res_sorted[k] = res[order[k]];
    CK2(cudaMemcpyAsync(dres, res_sorted.data(), (size_t)n * sizeof(gdk_result2d), cudaMemcpyHostToDevice, ctx->stream));

    PhaseTimer pt;
    pt.begin(ctx, GDK_PH_CONV2D);
    k_build_kernel2d<<<n, 256, 0, ctx->stream>>>(dcj);
    ctx->launches++;
    // groups of jobs with similar window size
    struct Grp {
        int b, e, wmax, Gmax;
    };
    std::vector<Grp> groups;
    {
        int b = 0;
        while (b < n) {
            int lim = 32;
            while (lim < cjs[b].w) lim *= 2;
            int e = b, gm = 0, wm = 0;
            while (e < n && cjs[e].w <= lim && e - b < 384) {  // <= 384 jobs: the finished grids of a group are copied
                                                                // back (second stream) while the next group computes
                gm = std::max(gm, cjs[e].G);
                wm = std::max(wm, cjs[e].w);
                e++;
            }
            groups.push_back(Grp{b, e, wm, gm});
            b = e;
        }
    }
    // kernel rows per shared-memory chunk: as many as fit a per-mode budget (fewer refills of the input tile)
    auto conv_smem = [](int wmax, int kc) { return ((size_t)(CV_TY + kc - 1) * 8 * cv_q(wmax) + (size_t)kc * cv_kp(wmax)) * 8; };
    auto pick_kc = [&](int wmax, size_t budget) {
        int kc = 2 * wmax + 1;
        while (kc > CV_KC && conv_smem(wmax, kc) > budget) kc--;
        return std::max(kc, std::min(CV_KC, 2 * wmax + 1));
    };
    const size_t budget0 = std::min<size_t>(100 << 10, (size_t)ctx->max_smem - 4096), budget1 = std::min<size_t>(54 << 10, (size_t)ctx->max_smem - 4096);
    size_t smem_max = 0;
    for (const Grp& g : groups)
        smem_max = std::max(smem_max, std::max(conv_smem(g.wmax, pick_kc(g.wmax, budget0)), conv_smem(g.wmax, pick_kc(g.wmax, budget1))));
    if (smem_max > (size_t)ctx->max_smem) return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "window too large for the convolution kernel");
    CK2(cudaFuncSetAttribute(k_conv2d<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 48 << 10)));
    CK2(cudaFuncSetAttribute(k_conv2d<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 48 << 10)));
    CK2(cudaFuncSetAttribute(k_conv2d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 48 << 10)));
    if (likes) {
        CK2(cudaFuncSetAttribute(k_conv2d<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 48 << 10)));
        CK2(cudaFuncSetAttribute(k_conv2d<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 48 << 10)));
    }
    CK2(cudaFuncSetAttribute(k_contours2d, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    // output
    double* dout = nullptr;
    if (dev_out) {
        dout = P_out;
    } else {
        if (ctx->f2.ensure(otot)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D output buffer");
        dout = ctx->f2.p;
    }
    double* dlout = nullptr;
    if (likes) {
        if (dev_out) {
            dlout = likes_out;
        } else {
            if (ctx->f2l.ensure(otot)) return gdk_fail(ctx, GDK_ERR_NOMEM, "2D mean-likelihood output buffer");
            dlout = ctx->f2l.p;
        }
    }
    bool any_contours = false;
    for (int i = 0; i < n; i++) any_contours = any_contours || specs[i].n_contours > 0;
    while (ctx->pipe_events.size() < groups.size()) {
        cudaEvent_t e;
        CK2(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    int gidx = 0;
    for (const Grp& g : groups) {
        const int nj = g.e - g.b;
        const int K = 2 * g.wmax + 1;
        // mask tables / maps (jobs without bias correction and without boundary correction skip internally)
        dim3 gt((unsigned)K, (unsigned)nj);
        k_mask_T<<<gt, 256, 0, ctx->stream>>>(dcj + g.b);
        dim3 gm((unsigned)((g.Gmax + 255) / 256), (unsigned)nj);
        k_mask_maps<<<gm, 256, 0, ctx->stream>>>(dcj + g.b);
        if (masks) {
            dim3 gu((unsigned)((g.Gmax * g.Gmax + 255) / 256), (unsigned)nj);
            k_umask_maps<<<gu, 256, 0, ctx->stream>>>(dcj + g.b);
            ctx->launches++;
        }
        const int tiles = ((g.Gmax + CV_TX - 1) / CV_TX) * ((g.Gmax + CV_TY - 1) / CV_TY);
        dim3 gc((unsigned)tiles, (unsigned)nj);
        const int kc0 = pick_kc(g.wmax, budget0), kc1 = pick_kc(g.wmax, budget1);
        double cflops0 = 0, cflops1 = 0, cbytes = 0;
        for (int k = g.b; k < g.e; k++) {
            const double K2 = (2.0 * cjs[k].w + 1) * (2.0 * cjs[k].w + 1), G2 = (double)cjs[k].G * cjs[k].G;
            cflops0 += 2.0 * G2 * K2 * ((cjs[k].bounded && cjs[k].bco == 1) ? 3.0 : 1.0);
            cflops1 += 2.0 * G2 * K2;
            cbytes += 16.0 * G2;
        }
        {
            bool any_mom = false;
            for (int k = g.b; k < g.e; k++) any_mom = any_mom || (cjs[k].bounded && cjs[k].bco == 1);
            KernelTimer kt(ctx, GDK_K_CONV2D_0, cbytes, cflops0);
            if (any_mom)
                k_conv2d<0><<<gc, 256, conv_smem(g.wmax, kc0), ctx->stream>>>(dcj + g.b, 0, g.wmax, kc0);
            else
                k_conv2d<4><<<gc, 256, conv_smem(g.wmax, kc1), ctx->stream>>>(dcj + g.b, 0, g.wmax, kc1);
        }
        dim3 gcirc((unsigned)((g.Gmax * g.Gmax + 255) / 256), (unsigned)nj);
        if (any_periodic) {
            k_conv2d_circ<0><<<gcirc, 256, 0, ctx->stream>>>(dcj + g.b, 0);
            ctx->launches++;
        }
        dim3 gb(64, (unsigned)nj);
        if (likes) {
            // mean likelihoods (mcsamples.py:1886-1898): likes (*) Win, optional bias step, ratio to the raw bins2D
            k_conv2d<2><<<gc, 256, conv_smem(g.wmax, kc1), ctx->stream>>>(dcj + g.b, 0, g.wmax, kc1);
            if (any_periodic) k_conv2d_circ<2><<<gcirc, 256, 0, ctx->stream>>>(dcj + g.b, 0);
            k_likes2d<0><<<gb, 256, 0, ctx->stream>>>(dcj + g.b);
            k_conv2d<3><<<gc, 256, conv_smem(g.wmax, kc1), ctx->stream>>>(dcj + g.b, 0, g.wmax, kc1);
            if (any_periodic) k_conv2d_circ<3><<<gcirc, 256, 0, ctx->stream>>>(dcj + g.b, 0);
            k_likes2d<1><<<gb, 256, 0, ctx->stream>>>(dcj + g.b);
            ctx->launches += any_periodic ? 6 : 4;
        }
        k_boundary2d<<<gb, 256, 0, ctx->stream>>>(dcj + g.b);
        ctx->launches += 4;
        for (int it = 0; it < max_mbc; it++) {
            k_make_box<<<gb, 256, 0, ctx->stream>>>(dcj + g.b, it);
            {
                KernelTimer kt(ctx, GDK_K_CONV2D_1, cbytes, cflops1);
                k_conv2d<1><<<gc, 256, conv_smem(g.wmax, kc1), ctx->stream>>>(dcj + g.b, it, g.wmax, kc1);
            }
            ctx->launches += 2;
            if (any_periodic) {
                k_conv2d_circ<1><<<gcirc, 256, 0, ctx->stream>>>(dcj + g.b, it);
                ctx->launches++;
            }
        }
        // normalised output of this group; its device->host copies run on the second stream behind an event
        dim3 gf(64, (unsigned)nj);
        if (masks) {
            k_umask_zero<<<gf, 256, 0, ctx->stream>>>(dcj + g.b);
            ctx->launches++;
        }
        k_finalize2d<<<gf, 256, 0, ctx->stream>>>(dcj + g.b, dout, doffs + g.b, dres + g.b);
        ctx->launches++;
        if (likes) {
            k_likes_finalize2d<<<gf, 256, 0, ctx->stream>>>(dcj + g.b, dlout, doffs + g.b);
            ctx->launches++;
        }
        if (any_contours) {
            {
                KernelTimer kt(ctx, GDK_K_CONTOURS2D, cbytes / 2, 0);  // every grid read once
                k_contours2d<<<nj, 512, CT_SMEM, ctx->stream>>>(dcj + g.b, dout, doffs + g.b, dres + g.b);
            }
            ctx->launches++;
        }
        if (peers_out) {  // this group's grids into every peer's gathered window, behind the next group's convolutions
            cudaEvent_t ev = ctx->pipe_events[gidx];
            CK2(cudaEventRecord(ev, ctx->stream));
            CK2(cudaStreamWaitEvent(ctx->stream2, ev, 0));
            if (ptab.n > 0) {
                k_push_peers<<<dim3((unsigned)nj, (unsigned)ptab.n), 256, 0, ctx->stream2>>>(dout, doffs + g.b, dcnt + g.b, 0, ptab);
                ctx->launches++;
            }
        }
        if (!dev_out) {
            cudaEvent_t ev = ctx->pipe_events[gidx];
            CK2(cudaEventRecord(ev, ctx->stream));
            CK2(cudaStreamWaitEvent(ctx->stream2, ev, 0));
            for (int k = g.b; k < g.e; k++) {
                const size_t cnt = (size_t)cjs[k].G * cjs[k].G * 8;
                CK2(cudaMemcpyAsync(P_out + offs_sorted[k], dout + devo_sorted[k], cnt, cudaMemcpyDeviceToHost, ctx->stream2));
                if (likes) CK2(cudaMemcpyAsync(likes_out + offs_sorted[k], dlout + devo_sorted[k], cnt, cudaMemcpyDeviceToHost, ctx->stream2));
            }
        }
        gidx++;
    }
    pt.end();
    CK2(cudaGetLastError());
    CK2(cudaMemcpyAsync(res_sorted.data(), dres, (size_t)n * sizeof(gdk_result2d), cudaMemcpyDeviceToHost, ctx->stream));
    if (!dev_out || peers_out) CK2(cudaStreamSynchronize(ctx->stream2));
    CK2(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n; k++) {
        res[order[k]].status = res_sorted[k].status;
        for (int c = 0; c < 4; c++) res[order[k]].levels[c] = res_sorted[k].levels[c];
    }
    return 0;
}

static int density2d_impl(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* P_out, const int64_t* offsets,
                          gdk_result2d* res, uint32_t flags, bool hist_only, double* likes_out = nullptr,
                          const double* masks = nullptr, const int64_t* mask_offsets = nullptr, const int32_t* mask_w = nullptr) {
    if (!ctx) return GDK_ERR_ARG;
    WallTimer wt{ctx, 1};
    if (likes_out && !ctx->have_loglikes) return gdk_fail(ctx, GDK_ERR_STATE, "meanlikes needs gdk_set_loglikes");
    if (n <= 0 || !specs || (!P_out && !(flags & GDK_BW_ONLY)) || !offsets || (!hist_only && !res))
        return gdk_fail(ctx, GDK_ERR_ARG, "2D batch: bad arguments");
    if (ctx->N <= 0) return gdk_fail(ctx, GDK_ERR_STATE, "no samples set");
    CK2(cudaSetDevice(ctx->device));
    for (int i = 0; i < n; i++) {
        const gdk_spec2d& s = specs[i];
        if (s.px < 0 || s.px >= ctx->P || s.py < 0 || s.py >= ctx->P) return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: parameter out of range", i);
        if (s.fine_bins < 8 || s.fine_bins > 4096 || s.base_fine_bins < 8 || s.base_fine_bins > 4096)
            return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: fine_bins_2D %d unsupported", i, s.fine_bins);
        if (!(s.xbinmax > s.xbinmin) || !(s.ybinmax > s.ybinmin)) return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: empty bin range", i);
        if (hist_only) continue;
        if (s.bw_mode < 0 || s.bw_mode > 3) return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: bad bw_mode", i);
        if (s.bw_mode != GDK_BW2D_FIXED && !(s.neff > 0)) return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: N_eff must be > 0", i);
        if (s.boundary_correction_order > 1) return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: boundary_correction_order must be <= 1", i);
        if (s.bw_mode == GDK_BW2D_SHEAR && (s.shear_i < 0 || s.shear_i >= ctx->P || s.shear_j < 0 || s.shear_j >= ctx->P))
            return gdk_fail(ctx, GDK_ERR_ARG, "pair %d: bad shear parameters", i);
    }
    if (ctx->use_hot) {  // the hot-window origins come from the weighted means
        int rcm = gdk_compute_moments(ctx);
        if (rcm) return rcm;
    }
    // chunks bounded by a memory budget
    const size_t freeb = ctx->query_free();
    const size_t budget = std::max<size_t>((size_t)1 << 30, std::min<size_t>((size_t)32 << 30, (freeb + ctx->bytes_arena.cap) / 2));
    int b = 0;
    while (b < n) {
        size_t acc = 0;
        int e = b;
        while (e < n) {
            const size_t add = bytes_per_pair_estimate(specs[e], likes_out != nullptr);
            if (e > b && acc + add > budget) break;
            acc += add;
            e++;
        }
        int rc = density2d_chunk(ctx, e - b, specs + b, P_out, offsets + b, hist_only ? nullptr : res + b, flags, hist_only, likes_out,
                                 masks, mask_offsets ? mask_offsets + b : nullptr, mask_w ? mask_w + b : nullptr);
        if (rc) return rc;
        b = e;
    }
    return GDK_OK;
}

extern "C" int32_t gdk_density2d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* P_out, const int64_t* offsets,
                                       gdk_result2d* res, uint32_t flags) {
    return density2d_impl(ctx, n, specs, P_out, offsets, res, flags, false);
}

extern "C" int32_t gdk_density2d_likes_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* P_out, double* likes_out,
                                             const int64_t* offsets, gdk_result2d* res, uint32_t flags) {
    return density2d_impl(ctx, n, specs, P_out, offsets, res, flags, false, likes_out);
}

extern "C" int32_t gdk_density2d_masked_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, const double* masks,
                                              const int64_t* mask_offsets, const int32_t* mask_w, double* P_out, double* likes_out,
                                              const int64_t* offsets, gdk_result2d* res, uint32_t flags) {
    if (ctx && (!masks || !mask_offsets || !mask_w)) return gdk_fail(ctx, GDK_ERR_ARG, "gdk_density2d_masked_batch: masks, offsets and half-widths are required");
    return density2d_impl(ctx, n, specs, P_out, offsets, res, flags, false, likes_out, masks, mask_offsets, mask_w);
}

extern "C" int32_t gdk_hist2d_batch(gdk_ctx* ctx, int32_t n, const gdk_spec2d* specs, double* bins_out, const int64_t* offsets) {
    return density2d_impl(ctx, n, specs, bins_out, offsets, nullptr, 0, true);
}
