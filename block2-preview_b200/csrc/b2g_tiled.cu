// b2g_tiled.cu — the two-phase DMMA replay of the H.C pair list (see b2g_tiled.cuh).
#include "b2g_tiled.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <numeric>
#include <tuple>

namespace b2g {

// ----------------------------------------------------------------------------
// Shared pipelined main loop.  Src provides:
//   int  steps() const          number of BK-deep stages of this unit
//   void issue(double*, double*) cp.async the next stage into (As, Bs) and advance
// ----------------------------------------------------------------------------
template <class Cfg, bool A_KC, bool B_KC, bool FULL, class Src>
__device__ __forceinline__ void mainloop_impl(Src &src, double *smem, double (&acc)[Cfg::MI][Cfg::NI][2], int wm0,
                                              int wn0, int mi_n, int ni_n) {
    double *As = smem, *Bs = smem + Cfg::STAGES * Cfg::A_STAGE;
    const int total = src.steps();
#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; s++) {
        if (s < total)
            src.issue(As + s * Cfg::A_STAGE, Bs + s * Cfg::B_STAGE);
        cp_async_commit();
    }
    for (int step = 0; step < total; step++) {
        cp_async_wait<Cfg::STAGES - 2>();
        __syncthreads();
        const int nxt = step + Cfg::STAGES - 1;
        if (nxt < total) {
            const int st = nxt % Cfg::STAGES;
            src.issue(As + st * Cfg::A_STAGE, Bs + st * Cfg::B_STAGE);
        }
        cp_async_commit();
        const int cur = step % Cfg::STAGES;
        compute_stage<Cfg, A_KC, B_KC, FULL>(As + cur * Cfg::A_STAGE, Bs + cur * Cfg::B_STAGE, acc, wm0, wn0, mi_n,
                                             ni_n);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// The guard-free body is chosen per CTA (block-uniform: every warp of an interior tile is full).
template <class Cfg, bool A_KC, bool B_KC, class Src>
__device__ __forceinline__ void mainloop(Src &src, double *smem, double (&acc)[Cfg::MI][Cfg::NI][2], int wm0, int wn0,
                                         int mi_n, int ni_n, bool cta_full) {
    if (cta_full)
        mainloop_impl<Cfg, A_KC, B_KC, true>(src, smem, acc, wm0, wn0, mi_n, ni_n);
    else
        mainloop_impl<Cfg, A_KC, B_KC, false>(src, smem, acc, wm0, wn0, mi_n, ni_n);
}

template <class Cfg> __device__ __forceinline__ void warp_origin(int &wm0, int &wn0) {
    const int warp = threadIdx.x >> 5;
    wm0 = (warp / Cfg::WN) * Cfg::WTM;
    wn0 = (warp % Cfg::WN) * Cfg::WTN;
}

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ------------------------------ phase 1 -------------------------------------
template <class Cfg, bool B_KC> struct P1Src {
    const double *a, *b;
    int lda, ldb, m_valid, n_valid, k_left, nsteps;
    __device__ int steps() const { return nsteps; }
    __device__ void issue(double *As, double *Bs) {
        load_tile<Cfg::BM, Cfg::THREADS, true, Cfg::BK>(As, a, lda, m_valid, k_left);
        load_tile<Cfg::BN, Cfg::THREADS, B_KC, Cfg::BK>(Bs, b, ldb, n_valid, k_left);
        a += Cfg::BK;
        b += B_KC ? Cfg::BK : (size_t)Cfg::BK * ldb;
        k_left -= Cfg::BK;
    }
};

template <class Cfg, bool B_KC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
phase1_kernel(const P1Pair *__restrict__ pairs, const Unit *__restrict__ units, int n_units,
              unsigned int *__restrict__ counter, const double *__restrict__ c, double *__restrict__ wbuf) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_unit;
    int wm0, wn0;
    warp_origin<Cfg>(wm0, wn0);
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    // units are claimed one ahead: the atomic for the next unit is in flight while this one runs
    int u = blockIdx.x;
    while (u < n_units) {
        if (threadIdx.x == 0)
            s_unit = (int)(atomicAdd(counter, 1u) + gridDim.x);
        const Unit un = units[u];
        const P1Pair p = pairs[un.idx];
        const int row0 = un.row0, col0 = un.col0;
        P1Src<Cfg, B_KC> src;
        src.a = c + p.a_off + (size_t)row0 * p.lda;
        src.b = B_KC ? p.b0 + (size_t)col0 * p.ldb : p.b0 + col0;
        src.lda = p.lda, src.ldb = p.ldb;
        src.m_valid = p.m0 - row0, src.n_valid = p.n0 - col0;
        src.k_left = p.k0, src.nsteps = (p.k0 + Cfg::BK - 1) / Cfg::BK;
        const int mi_n = clampi((src.m_valid - wm0 + 7) / 8, 0, Cfg::MI);
        const int ni_n = clampi((src.n_valid - wn0 + 7) / 8, 0, Cfg::NI);
        double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++)
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        mainloop<Cfg, true, B_KC>(src, smem, acc, wm0, wn0, mi_n, ni_n, src.m_valid >= Cfg::BM && src.n_valid >= Cfg::BN);
        double *w = wbuf + p.w_off;
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++) {
                const int r = row0 + wm0 + mi * 8 + lr, cc = col0 + wn0 + ni * 8 + lc * 2;
                if (mi < mi_n && ni < ni_n && r < p.m0) {
                    if (cc < p.n0)
                        w[(size_t)r * p.n0 + cc] = p.alpha * acc[mi][ni][0];
                    if (cc + 1 < p.n0)
                        w[(size_t)r * p.n0 + cc + 1] = p.alpha * acc[mi][ni][1];
                }
            }
        u = s_unit; // written before the barriers of the main loop
        __syncthreads();
    }
}

// ------------------------------ phase 2 -------------------------------------
template <class Cfg, bool A_KC> struct P2Src {
    const P2Seg *seg, *seg_end;
    const double *wbuf;
    const double *a, *b;
    P2Seg nxt; // descriptor of the following segment, fetched one segment ahead
    int lda, n0, row0, col0, m_valid, n_valid, k_left, nsteps;
    __device__ int steps() const { return nsteps; }
    __device__ void use(const P2Seg &s) {
        lda = s.lda;
        a = A_KC ? s.a1 + (size_t)row0 * lda : s.a1 + row0;
        b = wbuf + s.w_off + col0;
        k_left = s.klen;
        if (seg + 1 < seg_end)
            nxt = seg[1];
    }
    __device__ void open() { use(*seg); }
    __device__ void issue(double *As, double *Bs) {
        load_tile<Cfg::BM, Cfg::THREADS, A_KC, Cfg::BK>(As, a, lda, m_valid, k_left);
        load_tile<Cfg::BN, Cfg::THREADS, false, Cfg::BK>(Bs, b, n0, n_valid, k_left);
        k_left -= Cfg::BK;
        if (k_left > 0) {
            a += A_KC ? Cfg::BK : (size_t)Cfg::BK * lda;
            b += (size_t)Cfg::BK * n0;
        } else if (++seg < seg_end)
            use(nxt);
    }
};

template <class Cfg, bool A_KC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
phase2_kernel(const P2Window *__restrict__ wins, const P2Seg *__restrict__ segs, const Unit *__restrict__ units,
              int n_units, unsigned int *__restrict__ counter, const double *__restrict__ wbuf,
              double *__restrict__ v, double scale, double *__restrict__ pbuf) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_unit;
    int wm0, wn0;
    warp_origin<Cfg>(wm0, wn0);
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    // units are claimed one ahead: the atomic for the next unit is in flight while this one runs
    int u = blockIdx.x;
    while (u < n_units) {
        if (threadIdx.x == 0)
            s_unit = (int)(atomicAdd(counter, 1u) + gridDim.x);
        const Unit un = units[u];
        const P2Window win = wins[un.idx];
        P2Src<Cfg, A_KC> src;
        src.seg = segs + un.seg_begin, src.seg_end = segs + un.seg_end, src.wbuf = wbuf;
        src.n0 = win.n0, src.row0 = un.row0, src.col0 = un.col0;
        src.m_valid = win.m1 - src.row0, src.n_valid = win.n0 - src.col0;
        int ns = 0;
        for (const P2Seg *s = src.seg; s < src.seg_end; s++)
            ns += (s->klen + Cfg::BK - 1) / Cfg::BK;
        src.nsteps = ns;
        src.open();
        const int mi_n = clampi((src.m_valid - wm0 + 7) / 8, 0, Cfg::MI);
        const int ni_n = clampi((src.n_valid - wn0 + 7) / 8, 0, Cfg::NI);
        double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++)
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        mainloop<Cfg, A_KC, false>(src, smem, acc, wm0, wn0, mi_n, ni_n, src.m_valid >= Cfg::BM && src.n_valid >= Cfg::BN);
        if (un.poff >= 0) {
            // deterministic mode: the K-chunk partial goes to its own tile slot; reduce_kernel sums
            // the slots of a sigma tile in chunk order
            double *part = pbuf + un.poff;
#pragma unroll
            for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ni++) {
                    const int r = wm0 + mi * 8 + lr, cc = wn0 + ni * 8 + lc * 2;
                    if (mi < mi_n && ni < ni_n) {
                        part[r * Cfg::BN + cc] = acc[mi][ni][0];
                        part[r * Cfg::BN + cc + 1] = acc[mi][ni][1];
                    }
                }
        } else {
            double *out = v + win.c_off;
#pragma unroll
            for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ni++) {
                    const int r = src.row0 + wm0 + mi * 8 + lr, cc = src.col0 + wn0 + ni * 8 + lc * 2;
                    if (mi < mi_n && ni < ni_n && r < win.m1) {
                        if (cc < win.n0)
                            atomicAdd(out + (size_t)r * win.ldc + cc, scale * acc[mi][ni][0]);
                        if (cc + 1 < win.n0)
                            atomicAdd(out + (size_t)r * win.ldc + cc + 1, scale * acc[mi][ni][1]);
                    }
                }
        }
        u = s_unit; // written before the barriers of the main loop
        __syncthreads();
    }
}

// ------------------------------ W pre-sum -----------------------------------
// sum_p A1 * W_p = A1 * (sum_p W_p) for pairs that share the sigma window and the operator
// block A1: their W are added (memory-bound, in place into the first one) so phase 2 multiplies
// by A1 once.
__global__ void __launch_bounds__(256) wsum_kernel(const SumTask *__restrict__ tasks, int n_tasks,
                                                   double *__restrict__ wbuf) {
    for (int t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const SumTask k = tasks[t];
        double *d = wbuf + k.dst + k.start;
        for (int64_t e = threadIdx.x; e < k.count; e += 256) {
            double s = d[e];
            for (int j = 0; j < k.nsrc; j++)
                s += wbuf[k.src[j] + k.start + e];
            d[e] = s;
        }
    }
}

// ------------------------------ sigma reduce --------------------------------
// Deterministic mode: sigma tile += scale * sum over its K-chunk partials, in chunk order; every
// sigma element is written by exactly one thread, so the matvec is bit-reproducible run to run.
constexpr int RED_SPLIT = 8; // CTAs per sigma tile
__global__ void __launch_bounds__(256)
reduce_kernel(const OutTile *__restrict__ tiles, int n_tiles, const int64_t *__restrict__ part_off,
              const P2Window *__restrict__ wins, const double *__restrict__ pbuf, double *__restrict__ v,
              double scale) {
    for (int job = blockIdx.x; job < n_tiles * RED_SPLIT; job += gridDim.x) {
        const OutTile ot = tiles[job / RED_SPLIT];
        const int part = job % RED_SPLIT;
        const P2Window win = wins[ot.win];
        const int rows = min(ot.bm, win.m1 - ot.row0), cols = min(ot.bn, win.n0 - ot.col0);
        const int total = rows * ot.bn, chunk = (total + RED_SPLIT - 1) / RED_SPLIT;
        const int e_end = min(total, (part + 1) * chunk);
        double *out = v + win.c_off + (size_t)ot.row0 * win.ldc + ot.col0;
        for (int e = part * chunk + threadIdx.x; e < e_end; e += 256) {
            const int r = e / ot.bn, c = e - r * ot.bn;
            if (c < cols) {
                double s = 0.0;
                int p = ot.part_begin;
                for (; p + 4 <= ot.part_end; p += 4) { // fixed summation order, four loads in flight
                    const double x0 = pbuf[part_off[p] + e], x1 = pbuf[part_off[p + 1] + e],
                                 x2 = pbuf[part_off[p + 2] + e], x3 = pbuf[part_off[p + 3] + e];
                    s = (((s + x0) + x1) + x2) + x3;
                }
                for (; p < ot.part_end; p++)
                    s += pbuf[part_off[p] + e];
                out[(size_t)r * win.ldc + c] += scale * s;
            }
        }
    }
}

// ----------------------------------------------------------------------------
// Host side: tile configurations, unit lists, launches
// ----------------------------------------------------------------------------
// Tile configurations.  An output matrix is cut into 128-row tiles plus one 64-row strip for the
// remainder, and into 64-column tiles plus 16- or 8-column strips for the remainder, so padding
// waste stays at the 8-element granularity of the DMMA blocks.
using Cfg0 = TileCfg<128, 64, 4, 2, 3, 2>; // 8 warps, 32x32 warp tiles (bulk of large sectors)
using Cfg1 = TileCfg<64, 64, 2, 2, 3, 3>; // 4 warps, 32x32
using Cfg2 = TileCfg<128, 16, 8, 1, 4>; // 8 warps, 16x16 (column remainders)
using Cfg3 = TileCfg<64, 16, 4, 1, 4>;  // 4 warps, 16x16
using Cfg4 = TileCfg<128, 8, 8, 1, 4>;  // 8 warps, 16x8  (skinny sigma windows, n0 <= 8)
using Cfg5 = TileCfg<64, 8, 4, 1, 4>;   // 4 warps, 16x8
constexpr int NCFG = 6;

struct CfgInfo {
    int bm, bn, threads, smem;
};
static const CfgInfo kCfg[NCFG] = {{Cfg0::BM, Cfg0::BN, Cfg0::THREADS, Cfg0::SMEM_BYTES},
                                   {Cfg1::BM, Cfg1::BN, Cfg1::THREADS, Cfg1::SMEM_BYTES},
                                   {Cfg2::BM, Cfg2::BN, Cfg2::THREADS, Cfg2::SMEM_BYTES},
                                   {Cfg3::BM, Cfg3::BN, Cfg3::THREADS, Cfg3::SMEM_BYTES},
                                   {Cfg4::BM, Cfg4::BN, Cfg4::THREADS, Cfg4::SMEM_BYTES},
                                   {Cfg5::BM, Cfg5::BN, Cfg5::THREADS, Cfg5::SMEM_BYTES}};

static inline int cfg_of(int bm, int bn) { return (bn == 64 ? 0 : bn == 16 ? 2 : 4) + (bm == 128 ? 0 : 1); }

struct Strip {
    int origin, tile; // first element, tile extent class
};
// rows: 128-row tiles, then a 64-row strip when the remainder fits one
static std::vector<Strip> split_rows(int m) {
    std::vector<Strip> out;
    int r = 0;
    while (m - r > 64) {
        out.push_back(Strip{r, 128});
        r += 128;
    }
    if (m - r > 0)
        out.push_back(Strip{r, 64});
    return out;
}
// columns: 64-column tiles, then 16- / 8-column strips for the remainder
static std::vector<Strip> split_cols(int n) {
    std::vector<Strip> out;
    int c = 0;
    while (n - c > 32) {
        out.push_back(Strip{c, 64});
        c += 64;
    }
    while (n - c > 8) {
        out.push_back(Strip{c, 16});
        c += 16;
    }
    if (n - c > 0)
        out.push_back(Strip{c, 8});
    return out;
}

struct LaunchGroup { // one kernel launch: units of one (phase, cfg, layout)
    int phase, cfg, layout;
    Unit *d_units = nullptr;
    int n_units = 0;
    double flops = 0; // useful 2*m*n*k FLOPs of the launch (valid rows x cols only)
};

struct TiledPlan {
    b2g_context *ctx = nullptr;
    SumTask *d_sum = nullptr;
    int n_sum = 0;
    double sum_bytes = 0;
    // deterministic sigma accumulation
    OutTile *d_tiles = nullptr;
    int64_t *d_part_off = nullptr;
    double *d_pbuf = nullptr;
    int n_tiles = 0;
    // sigma windows may overlap (a block and its sub-windows): tiles are ordered by the colour of
    // their window in the interval-overlap graph and each colour is reduced by its own launch
    std::vector<std::pair<int, int>> tile_ranges;
    size_t pbuf_doubles = 0;
    P1Pair *d_p1 = nullptr;
    P2Window *d_win = nullptr;
    P2Seg *d_seg = nullptr;
    double *d_wbuf = nullptr;
    unsigned int *d_counters = nullptr;
    size_t wbuf_doubles = 0;
    std::vector<LaunchGroup> groups;
    std::vector<void *> to_free;
};

template <class Cfg, bool L> static int launch_p1(const LaunchGroup &g, const TiledPlan &tp, b2g_context *ctx,
                                                  cudaStream_t stream, unsigned int *counter, const double *c) {
    auto kern = phase1_kernel<Cfg, L>;
    // the dynamic shared memory limit is a per-device function attribute: one bit per device ordinal
    static std::atomic<uint64_t> attr_mask{0};
    const uint64_t bit = (uint64_t)1 << (ctx->device & 63);
    if (!(attr_mask.load(std::memory_order_acquire) & bit)) {
        B2G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_mask.fetch_or(bit, std::memory_order_release);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM_BYTES);
    const int grid = std::min(g.n_units, ctx->sm_count * std::max(per_sm, 1));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tp.d_p1, g.d_units, g.n_units, counter, c, tp.d_wbuf);
    return 0;
}
template <class Cfg, bool L> static int launch_p2(const LaunchGroup &g, const TiledPlan &tp, b2g_context *ctx,
                                                  cudaStream_t stream, unsigned int *counter, double *v,
                                                  double scale) {
    auto kern = phase2_kernel<Cfg, L>;
    // the dynamic shared memory limit is a per-device function attribute: one bit per device ordinal
    static std::atomic<uint64_t> attr_mask{0};
    const uint64_t bit = (uint64_t)1 << (ctx->device & 63);
    if (!(attr_mask.load(std::memory_order_acquire) & bit)) {
        B2G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_mask.fetch_or(bit, std::memory_order_release);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM_BYTES);
    const int grid = std::min(g.n_units, ctx->sm_count * std::max(per_sm, 1));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tp.d_win, tp.d_seg, g.d_units, g.n_units, counter,
                                                                tp.d_wbuf, v, scale, tp.d_pbuf);
    return 0;
}

} // namespace b2g

using namespace b2g;

void b2g_tiled_destroy(void *h) {
    TiledPlan *tp = (TiledPlan *)h;
    if (!tp)
        return;
    for (void *p : tp->to_free)
        b2g_dfree(tp->ctx, p);
    delete tp;
}

// Build the two-phase plan from the (device-pointer) pair list of the generic plan.
int b2g_tiled_build(b2g_plan *p) {
    b2g_context *ctx = p->ctx;
    const std::vector<B2GPair> &hp = p->h_pairs;
    const size_t n = hp.size();
    TiledPlan *tp = new TiledPlan();
    tp->ctx = ctx;
    p->tiled = tp;
    if (n == 0)
        return 0;
    const char *env_kc = getenv("B2G_KCHUNK");
    const int64_t kchunk = env_kc ? atoll(env_kc) : 2048;

    // ---- phase 1 descriptors + W workspace layout
    std::vector<P1Pair> p1(n);
    size_t woff = 0;
    for (size_t i = 0; i < n; i++) {
        const B2GPair &q = hp[i];
        P1Pair &d = p1[i];
        d.b0 = q.b0, d.w_off = (int64_t)woff, d.alpha = q.alpha0 * q.alpha1;
        d.a_off = q.a0_off, d.lda = q.lda0, d.ldb = q.ldb0;
        d.m0 = q.m0, d.n0 = q.n0, d.k0 = q.k0, d.tb0 = (q.flags & B2G_F_TB0) ? 1 : 0;
        woff += (size_t)q.m0 * q.n0;
    }
    tp->wbuf_doubles = woff;

    // ---- windows: pairs that accumulate into the same sigma window
    std::map<std::tuple<int64_t, int, int, int>, int> wid;
    std::vector<P2Window> wins;
    std::vector<std::vector<size_t>> wpairs[2]; // [layout][window] -> pair indices
    for (size_t i = 0; i < n; i++) {
        const B2GPair &q = hp[i];
        auto key = std::make_tuple(q.c1_off, q.m1, q.n0, q.ldc1);
        auto it = wid.find(key);
        int w;
        if (it == wid.end()) {
            w = (int)wins.size();
            wid[key] = w;
            wins.push_back(P2Window{q.c1_off, q.ldc1, q.m1, q.n0, 0});
            wpairs[0].emplace_back(), wpairs[1].emplace_back();
        } else
            w = it->second;
        wpairs[(q.flags & B2G_F_TA1) ? 1 : 0][w].push_back(i);
    }

    // ---- units
    struct HostUnit {
        Unit u;
        double cost, flops;
    };
    std::map<std::tuple<int, int, int>, std::vector<HostUnit>> groups; // (phase, cfg, layout)
    for (size_t i = 0; i < n; i++) {
        const B2GPair &q = hp[i];
        if (q.m0 == 0 || q.n0 == 0)
            continue;
        for (const Strip &rs : split_rows(q.m0))
            for (const Strip &cs : split_cols(q.n0)) {
                const int c = cfg_of(rs.tile, cs.tile);
                groups[std::make_tuple(1, c, p1[i].tb0)].push_back(
                    HostUnit{Unit{(int)i, rs.origin, cs.origin, 0, 0, 0, -1}, (double)rs.tile * cs.tile * (q.k0 + 32),
                             2.0 * std::min(rs.tile, q.m0 - rs.origin) * std::min(cs.tile, q.n0 - cs.origin) * q.k0});
            }
    }
    // K-chunk: 2048 for the big lists; shorter when the list is small so that phase 2 still
    // spreads over the whole chip (few sigma windows, each with a long chain of short segments)
    int64_t kchunk_eff = kchunk;
    if (!env_kc) {
        double tile_k = 0;
        for (int lay = 0; lay < 2; lay++)
            for (size_t w = 0; w < wins.size(); w++) {
                double ks = 0;
                for (size_t idx : wpairs[lay][w])
                    ks += hp[idx].m0;
                tile_k += ks * (double)split_rows(wins[w].m1).size() * (double)split_cols(wins[w].n0).size();
            }
        const double want_units = 24.0 * ctx->sm_count;
        kchunk_eff = (int64_t)std::min<double>(2048.0, std::max(128.0, tile_k / want_units));
        kchunk_eff = (kchunk_eff + 15) / 16 * 16;
    }
    std::vector<P2Seg> segs;
    std::vector<SumTask> sums;
    const char *env_merge = getenv("B2G_NO_WSUM");
    const bool merge_w = !(env_merge && env_merge[0] == '1');
    const int64_t sum_chunk = 32768;
    for (int lay = 0; lay < 2; lay++)
        for (size_t w = 0; w < wins.size(); w++) {
            auto &lst = wpairs[lay][w];
            if (lst.empty() || wins[w].m1 == 0 || wins[w].n0 == 0)
                continue;
            // neighbours share the operator block (pre-summed below; L2 reuse across K-chunks)
            std::stable_sort(lst.begin(), lst.end(), [&hp](size_t x, size_t y) {
                if (hp[x].a1 != hp[y].a1)
                    return hp[x].a1 < hp[y].a1;
                if (hp[x].lda1 != hp[y].lda1)
                    return hp[x].lda1 < hp[y].lda1;
                return hp[x].m0 < hp[y].m0;
            });
            const std::vector<Strip> rsv = split_rows(wins[w].m1), csv = split_cols(wins[w].n0);
            size_t s0 = segs.size();
            int64_t ksum = 0;
            auto flush = [&](size_t s1) {
                if (s1 == s0)
                    return;
                for (const Strip &rs : rsv)
                    for (const Strip &cs : csv)
                        groups[std::make_tuple(2, cfg_of(rs.tile, cs.tile), lay)].push_back(
                            HostUnit{Unit{(int)w, rs.origin, cs.origin, (int)s0, (int)s1, 0, -1},
                                     (double)rs.tile * cs.tile * (double)(ksum + 32),
                                     2.0 * std::min(rs.tile, wins[w].m1 - rs.origin) *
                                         std::min(cs.tile, wins[w].n0 - cs.origin) * (double)ksum});
                s0 = s1, ksum = 0;
            };
            for (size_t z = 0; z < lst.size();) {
                const B2GPair &q = hp[lst[z]];
                // run of pairs with the same operator block: one segment, W summed beforehand
                size_t z1 = z + 1;
                while (merge_w && z1 < lst.size() && z1 - z < 4 && hp[lst[z1]].a1 == q.a1 && hp[lst[z1]].lda1 == q.lda1 &&
                       hp[lst[z1]].m0 == q.m0)
                    z1++;
                if (q.m0 != 0) {
                    const int64_t total = (int64_t)q.m0 * q.n0;
                    for (size_t y = z + 1; y < z1; y += 3) { // up to 3 sources per task
                        SumTask st{};
                        st.dst = p1[lst[z]].w_off;
                        st.nsrc = (int)std::min<size_t>(3, z1 - y);
                        for (int j = 0; j < st.nsrc; j++)
                            st.src[j] = p1[lst[y + j]].w_off;
                        for (int64_t s = 0; s < total; s += sum_chunk) {
                            st.start = s, st.count = std::min<int64_t>(sum_chunk, total - s);
                            sums.push_back(st);
                            tp->sum_bytes += 8.0 * st.count * (st.nsrc + 2);
                        }
                    }
                    segs.push_back(P2Seg{q.a1, p1[lst[z]].w_off, q.lda1, q.m0});
                    ksum += q.m0;
                    if (ksum >= kchunk_eff)
                        flush(segs.size());
                }
                z = z1;
            }
            flush(segs.size());
        }

    // ---- upload
    auto upload = [&](const void *src, size_t bytes, void **dst) -> int {
        if (b2g_dmalloc(ctx, dst, bytes))
            return 1;
        tp->to_free.push_back(*dst);
        if (bytes)
            B2G_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };
    if (upload(p1.data(), p1.size() * sizeof(P1Pair), (void **)&tp->d_p1))
        return 1;
    if (upload(wins.data(), wins.size() * sizeof(P2Window), (void **)&tp->d_win))
        return 1;
    if (upload(segs.data(), segs.size() * sizeof(P2Seg), (void **)&tp->d_seg))
        return 1;
    // at most 4 pairs are merged per run, so every destination range belongs to exactly one task
    tp->n_sum = (int)sums.size();
    if (upload(sums.data(), sums.size() * sizeof(SumTask), (void **)&tp->d_sum))
        return 1;
    if (b2g_dmalloc(ctx, (void **)&tp->d_wbuf, std::max<size_t>(tp->wbuf_doubles, 2) * sizeof(double)))
        return 1;
    tp->to_free.push_back(tp->d_wbuf);
    // deterministic sigma accumulation: one partial slot per phase-2 unit, grouped by sigma tile
    // (a tile collects partials from both operand layouts), slots in K-chunk order
    const char *env_atomic = getenv("B2G_ATOMIC_SIGMA");
    std::vector<OutTile> tiles;
    std::vector<int64_t> part_off;
    if (!(env_atomic && env_atomic[0] == '1')) {
        std::map<std::tuple<int, int, int>, std::vector<std::pair<int, HostUnit *>>> by_tile; // (win,row0,col0)
        for (auto &kv : groups)
            if (std::get<0>(kv.first) == 2)
                for (HostUnit &hu : kv.second)
                    by_tile[std::make_tuple(hu.u.idx, hu.u.row0, hu.u.col0)].push_back(
                        std::make_pair(std::get<1>(kv.first), &hu));
        // colour the windows so that windows sharing sigma elements never share a launch
        std::vector<int> colour(wins.size(), 0);
        {
            std::vector<int> order(wins.size());
            std::iota(order.begin(), order.end(), 0);
            auto lo = [&wins](int w) { return (int64_t)wins[w].c_off; };
            auto hi = [&wins](int w) {
                return (int64_t)wins[w].c_off + (int64_t)(wins[w].m1 - 1) * wins[w].ldc + wins[w].n0;
            };
            std::sort(order.begin(), order.end(), [&](int x, int y) { return lo(x) < lo(y); });
            std::vector<std::pair<int64_t, int>> active; // (hi, colour)
            for (int w : order) {
                std::vector<char> used(active.size() + 1, 0);
                std::vector<std::pair<int64_t, int>> keep_a;
                for (auto &a : active)
                    if (a.first > lo(w)) {
                        keep_a.push_back(a);
                        if (a.second < (int)used.size())
                            used[a.second] = 1;
                    }
                int c = 0;
                while (c < (int)used.size() && used[c])
                    c++;
                colour[w] = c;
                keep_a.push_back(std::make_pair(hi(w), c));
                active.swap(keep_a);
            }
        }
        std::vector<std::pair<int, std::tuple<int, int, int>>> tile_order;
        for (auto &kv : by_tile)
            tile_order.push_back(std::make_pair(colour[std::get<0>(kv.first)], kv.first));
        std::stable_sort(tile_order.begin(), tile_order.end(),
                         [](const std::pair<int, std::tuple<int, int, int>> &x,
                            const std::pair<int, std::tuple<int, int, int>> &y) { return x.first < y.first; });
        size_t poff = 0;
        for (auto &to : tile_order) {
            auto kvit = by_tile.find(to.second);
            auto &kv = *kvit;
            if (tp->tile_ranges.empty() || to.first != (int)tp->tile_ranges.size() - 1) {
                while ((int)tp->tile_ranges.size() <= to.first)
                    tp->tile_ranges.push_back(std::make_pair((int)tiles.size(), (int)tiles.size()));
            }
            auto &lst = kv.second;
            std::stable_sort(lst.begin(), lst.end(),
                             [](const std::pair<int, HostUnit *> &x, const std::pair<int, HostUnit *> &y) {
                                 return x.second->u.seg_begin < y.second->u.seg_begin;
                             });
            const int c = lst[0].first;
            OutTile ot{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first), kCfg[c].bm, kCfg[c].bn,
                       (int)part_off.size(), 0, 0};
            for (auto &pr : lst) {
                pr.second->u.poff = (int64_t)poff;
                part_off.push_back((int64_t)poff);
                poff += (size_t)kCfg[c].bm * kCfg[c].bn;
            }
            ot.part_end = (int)part_off.size();
            tiles.push_back(ot);
            tp->tile_ranges[to.first].second = (int)tiles.size();
        }
        tp->pbuf_doubles = poff;
        tp->n_tiles = (int)tiles.size();
        if (upload(tiles.data(), tiles.size() * sizeof(OutTile), (void **)&tp->d_tiles))
            return 1;
        if (upload(part_off.data(), part_off.size() * sizeof(int64_t), (void **)&tp->d_part_off))
            return 1;
        if (b2g_dmalloc(ctx, (void **)&tp->d_pbuf, std::max<size_t>(poff, 2) * sizeof(double)))
            return 1;
        tp->to_free.push_back(tp->d_pbuf);
    }
    std::vector<std::vector<Unit>> keep; // host copies must outlive the async copies
    for (auto &kv : groups) {
        auto &hu = kv.second;
        if (hu.empty())
            continue;
        std::stable_sort(hu.begin(), hu.end(), [](const HostUnit &x, const HostUnit &y) { return x.cost > y.cost; });
        keep.emplace_back(hu.size());
        for (size_t i = 0; i < hu.size(); i++)
            keep.back()[i] = hu[i].u;
        LaunchGroup g;
        g.phase = std::get<0>(kv.first), g.cfg = std::get<1>(kv.first), g.layout = std::get<2>(kv.first);
        g.n_units = (int)hu.size();
        for (const HostUnit &x : hu)
            g.flops += x.flops;
        if (upload(keep.back().data(), hu.size() * sizeof(Unit), (void **)&g.d_units))
            return 1;
        tp->groups.push_back(g);
    }
    std::stable_sort(tp->groups.begin(), tp->groups.end(),
                     [](const LaunchGroup &a, const LaunchGroup &b) { return a.phase < b.phase; });
    if (b2g_dmalloc(ctx, (void **)&tp->d_counters, sizeof(unsigned int) * 64))
        return 1;
    tp->to_free.push_back(tp->d_counters);
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    p->stats.launches = (int64_t)tp->groups.size() + 1;
    p->stats.n_large = (int64_t)n, p->stats.n_small = 0;
    return 0;
}

int b2g_tiled_launch(b2g_plan *p, const double *c_dev, double *v_dev, double scale, b2g_kernel_stat *stats,
                     int cap, int *count) {
    TiledPlan *tp = (TiledPlan *)p->tiled;
    b2g_context *ctx = p->ctx;
    if (count)
        *count = 0;
    if (!tp || tp->groups.empty())
        return 0;
    B2G_CUDA(cudaMemsetAsync(tp->d_counters, 0, sizeof(unsigned int) * 64, ctx->stream));
    struct Rec {
        std::string name;
        double flops;
        int64_t units;
        cudaEvent_t e0, e1;
    };
    std::vector<Rec> recs;
    auto begin = [&](const std::string &name, double flops, int64_t units) -> int {
        if (!stats)
            return 0;
        Rec r{name, flops, units, nullptr, nullptr};
        B2G_CUDA(cudaEventCreate(&r.e0));
        B2G_CUDA(cudaEventCreate(&r.e1));
        B2G_CUDA(cudaEventRecord(r.e0, ctx->stream));
        recs.push_back(r);
        return 0;
    };
    auto end = [&]() -> int {
        if (stats)
            B2G_CUDA(cudaEventRecord(recs.back().e1, ctx->stream));
        return 0;
    };
    // Without profiling, the launches of one phase are forked onto side streams and joined before
    // the next phase; with profiling everything stays on the context stream (timed one by one).
    // Forking pays when no single launch fills the chip (measured: up to ~2x on the C2 / H10 lists,
    // -5 % on the 2 TFLOP Cr2 list where the big persistent kernels then compete), hence the bound.
    const bool fork_all = stats == nullptr && 2.0 * (double)p->stats.nflop_mnk < 3e11;
    // Large lists, experiment (B2G_FORK_SMALL=<flops>): only the minor launches (edge-strip and 64-row
    // configurations, a few % of the FLOPs each) go to the low-priority side streams, to fill the tails of
    // the big persistent launches that stay on the context stream.  Measured on the Cr2 M=4000 list: 81.5 /
    // 83.6 ms with the threshold at 50 / 15 GFLOP against 80.4 ms in sequence, so it is off by default.
    static const double fork_small_below = getenv("B2G_FORK_SMALL") ? atof(getenv("B2G_FORK_SMALL")) : 0.0;
    const bool fork_small = stats == nullptr && !fork_all && fork_small_below > 0.0;
    const bool fork = fork_all || fork_small;
    int n_forked = 0;
    auto join = [&]() -> int {
        for (int i = 0; i < std::min(n_forked, (int)b2g_context::N_SIDE); i++)
            B2G_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->side_done[i], 0));
        n_forked = 0;
        return 0;
    };
    if (fork)
        B2G_CUDA(cudaEventRecord(ctx->fork_ev, ctx->stream));
    int gi = 0;
    bool summed = false;
    for (const LaunchGroup &g : tp->groups) {
        if (g.phase == 2 && !summed) {
            summed = true;
            if (fork) {
                if (join())
                    return 1;
            }
            if (tp->n_sum > 0) {
                if (begin("wsum", 0.0, tp->n_sum))
                    return 1;
                wsum_kernel<<<std::min(tp->n_sum, ctx->sm_count * 8), 256, 0, ctx->stream>>>(tp->d_sum, tp->n_sum,
                                                                                             tp->d_wbuf);
                ctx->launches++;
                if (end())
                    return 1;
            }
            if (fork)
                B2G_CUDA(cudaEventRecord(ctx->fork_ev, ctx->stream));
        }
        cudaStream_t gs = ctx->stream;
        const bool side = fork_all || (fork_small && g.flops < fork_small_below);
        if (side) {
            gs = ctx->side[n_forked % b2g_context::N_SIDE];
            if (n_forked < b2g_context::N_SIDE)
                B2G_CUDA(cudaStreamWaitEvent(gs, ctx->fork_ev, 0));
        }
        unsigned int *counter = tp->d_counters + gi++;
        int rc = 0;
        char nm[64];
        snprintf(nm, sizeof(nm), "phase%d_%dx%d_%s", g.phase, kCfg[g.cfg].bm, kCfg[g.cfg].bn,
                 g.phase == 1 ? (g.layout ? "Bt" : "Bn") : (g.layout ? "At" : "An"));
        if (begin(nm, g.flops, g.n_units))
            return 1;
#define B2G_DISPATCH(PH, CFG, LAY, CALL)                                                \
    if (g.phase == PH && g.cfg == CFG && g.layout == LAY)                               \
        rc = CALL;
        B2G_DISPATCH(1, 0, 0, (launch_p1<Cfg0, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 0, 1, (launch_p1<Cfg0, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 1, 0, (launch_p1<Cfg1, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 1, 1, (launch_p1<Cfg1, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 2, 0, (launch_p1<Cfg2, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 2, 1, (launch_p1<Cfg2, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 3, 0, (launch_p1<Cfg3, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 3, 1, (launch_p1<Cfg3, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 4, 0, (launch_p1<Cfg4, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 4, 1, (launch_p1<Cfg4, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 5, 0, (launch_p1<Cfg5, false>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(1, 5, 1, (launch_p1<Cfg5, true>(g, *tp, ctx, gs, counter, c_dev)))
        B2G_DISPATCH(2, 0, 0, (launch_p2<Cfg0, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 0, 1, (launch_p2<Cfg0, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 1, 0, (launch_p2<Cfg1, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 1, 1, (launch_p2<Cfg1, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 2, 0, (launch_p2<Cfg2, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 2, 1, (launch_p2<Cfg2, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 3, 0, (launch_p2<Cfg3, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 3, 1, (launch_p2<Cfg3, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 4, 0, (launch_p2<Cfg4, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 4, 1, (launch_p2<Cfg4, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 5, 0, (launch_p2<Cfg5, true>(g, *tp, ctx, gs, counter, v_dev, scale)))
        B2G_DISPATCH(2, 5, 1, (launch_p2<Cfg5, false>(g, *tp, ctx, gs, counter, v_dev, scale)))
#undef B2G_DISPATCH
        if (rc)
            return rc;
        ctx->launches++;
        if (end())
            return 1;
        if (side) {
            B2G_CUDA(cudaEventRecord(ctx->side_done[n_forked % b2g_context::N_SIDE], gs));
            n_forked++;
        }
    }
    if (fork && join())
        return 1;
    if (tp->n_tiles > 0) {
        if (begin("sigma_reduce", 0.0, tp->n_tiles))
            return 1;
        for (const auto &rg : tp->tile_ranges) {
            const int nt = rg.second - rg.first;
            if (nt <= 0)
                continue;
            reduce_kernel<<<std::min(nt * RED_SPLIT, ctx->sm_count * 16), 256, 0, ctx->stream>>>(
                tp->d_tiles + rg.first, nt, tp->d_part_off, tp->d_win, tp->d_pbuf, v_dev, scale);
            ctx->launches++;
        }
        if (end())
            return 1;
    }
    B2G_CUDA(cudaGetLastError());
    if (stats) {
        B2G_CUDA(cudaStreamSynchronize(ctx->stream));
        int n = 0;
        for (size_t i = 0; i < recs.size(); i++) {
            float ms = 0;
            B2G_CUDA(cudaEventElapsedTime(&ms, recs[i].e0, recs[i].e1));
            if (n < cap) {
                snprintf(stats[n].name, sizeof(stats[n].name), "%s", recs[i].name.c_str());
                stats[n].flops = recs[i].flops, stats[n].ms = ms, stats[n].units = recs[i].units;
                n++;
            }
            cudaEventDestroy(recs[i].e0), cudaEventDestroy(recs[i].e1);
        }
        if (count)
            *count = n;
    }
    return 0;
}
