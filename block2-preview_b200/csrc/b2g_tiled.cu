// b2g_tiled.cu — the two-phase DMMA replay of the H.C pair list (see b2g_tiled.cuh).
#include "b2g_tiled.cuh"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <map>
#include <numeric>
#include <tuple>
#include <unordered_map>

namespace b2g {

// ----------------------------------------------------------------------------
// Shared pipelined main loop.  Src provides:
//   int    steps() const            number of BK-deep stages of this unit
//   double issue(double*, double*)  cp.async the next stage into (As, Bs), advance, and return the factor
//                                   of the K-segment the stage belongs to (phase 1; phase 2 returns 1)
// ----------------------------------------------------------------------------
template <class Cfg, bool A_KC, bool B_KC, bool FULL, bool SCALE, class Src>
__device__ __forceinline__ void mainloop_impl(Src &src, double *smem, double *s_alpha,
                                              double (&acc)[Cfg::MI][Cfg::NI][2], int wm0, int wn0, int mi_n,
                                              int ni_n) {
    double *As = smem, *Bs = smem + Cfg::STAGES * Cfg::A_STAGE;
    const int total = src.steps();
#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; s++) {
        if (s < total) {
            const double al = src.issue(As + s * Cfg::A_STAGE, Bs + s * Cfg::B_STAGE);
            if (SCALE && threadIdx.x == 0)
                s_alpha[s] = al;
        }
        cp_async_commit();
    }
    for (int step = 0; step < total; step++) {
        cp_async_wait<Cfg::STAGES - 2>();
        __syncthreads();
        const int nxt = step + Cfg::STAGES - 1;
        if (nxt < total) {
            const int st = nxt % Cfg::STAGES;
            const double al = src.issue(As + st * Cfg::A_STAGE, Bs + st * Cfg::B_STAGE);
            if (SCALE && threadIdx.x == 0)
                s_alpha[st] = al; // slot st was last read before the barrier above
        }
        cp_async_commit();
        const int cur = step % Cfg::STAGES;
        compute_stage<Cfg, A_KC, B_KC, FULL, SCALE>(As + cur * Cfg::A_STAGE, Bs + cur * Cfg::B_STAGE, acc, wm0, wn0,
                                                    mi_n, ni_n, SCALE ? s_alpha[cur] : 1.0);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// The guard-free body is chosen per CTA (block-uniform: every warp of an interior tile is full).
template <class Cfg, bool A_KC, bool B_KC, bool SCALE, class Src>
__device__ __forceinline__ void mainloop(Src &src, double *smem, double *s_alpha, double (&acc)[Cfg::MI][Cfg::NI][2],
                                         int wm0, int wn0, int mi_n, int ni_n, bool cta_full) {
    if (cta_full)
        mainloop_impl<Cfg, A_KC, B_KC, true, SCALE>(src, smem, s_alpha, acc, wm0, wn0, mi_n, ni_n);
    else
        mainloop_impl<Cfg, A_KC, B_KC, false, SCALE>(src, smem, s_alpha, acc, wm0, wn0, mi_n, ni_n);
}

template <class Cfg> __device__ __forceinline__ void warp_origin(int &wm0, int &wn0) {
    const int warp = threadIdx.x >> 5;
    wm0 = (warp / Cfg::WN) * Cfg::WTM;
    wn0 = (warp % Cfg::WN) * Cfg::WTN;
}

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ------------------------------ phase 1 -------------------------------------
// Programmatic dependent launch: the launches of one phase write disjoint outputs, so the next launch of the chain
// may fill the SMs as the CTAs of this one run out of units (no idle tail between 30+ persistent launches).
// Every CTA lets the dependents go at once and waits for its predecessor only when it is about to exit, which
// makes grid completion transitive along the chain; the first launch of phase 2 (wait_first) waits up front,
// because it reads the W panels of all phase-1 launches.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_enter(int wait_first) {
    if (wait_first)
        asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_exit() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__global__ void set_args_kernel(MatvecArgs *a, const double *c, double *v, double scale) {
    a->c = c, a->v = v, a->scale = scale;
}

template <class Cfg, bool B_KC> struct P1Src {
    const P1Seg *seg, *seg_end;
    const double *c;
    const double *a, *b;
    P1Seg nxt; // descriptor of the following segment, fetched one segment ahead
    double alpha;
    int lda, ldb, row0, col0, m_valid, n_valid, k_left, nsteps;
    __device__ int steps() const { return nsteps; }
    __device__ void use(const P1Seg &s) {
        lda = s.lda, ldb = s.ldb, alpha = s.alpha;
        a = c + s.a_off + (size_t)row0 * lda;
        b = B_KC ? s.b0 + (size_t)col0 * ldb : s.b0 + col0;
        k_left = s.k0;
        if (seg + 1 < seg_end)
            nxt = seg[1];
    }
    __device__ void open() { use(*seg); }
    __device__ double issue(double *As, double *Bs) {
        const double al = alpha;
        load_tile<Cfg::BM, Cfg::THREADS, true, Cfg::BK>(As, a, lda, m_valid, k_left);
        load_tile<Cfg::BN, Cfg::THREADS, B_KC, Cfg::BK>(Bs, b, ldb, n_valid, k_left);
        k_left -= Cfg::BK;
        if (k_left > 0) {
            a += Cfg::BK;
            b += B_KC ? Cfg::BK : (size_t)Cfg::BK * ldb;
        } else if (++seg < seg_end)
            use(nxt);
        return al;
    }
};

template <class Cfg, bool B_KC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
phase1_kernel(const P1Group *__restrict__ groups, const P1Seg *__restrict__ segs, const Unit *__restrict__ units,
              int n_units, unsigned int *__restrict__ counter, const double *__restrict__ c,
              double *__restrict__ wbuf, int wait_first, const MatvecArgs *__restrict__ ind) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_unit;
    __shared__ double s_alpha[Cfg::STAGES];
    if (ind) // replay of a captured graph: the wavefunction of this call
        c = ind->c;
    pdl_enter(wait_first);
    int wm0, wn0;
    warp_origin<Cfg>(wm0, wn0);
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    // Units are claimed one ahead (the atomic for the next unit is in flight while this one runs), the first one
    // included: under programmatic dependent launch the CTAs of a grid start as slots free up, and a CTA that
    // starts late must not sit on one of the most expensive units (the list is in cost order).
    if (threadIdx.x == 0)
        s_unit = (int)atomicAdd(counter, 1u);
    __syncthreads();
    int u = s_unit;
    __syncthreads();
    while (u < n_units) {
        if (threadIdx.x == 0)
            s_unit = (int)atomicAdd(counter, 1u);
        const Unit un = units[u];
        const P1Group g = groups[un.idx];
        const int row0 = un.row0, col0 = un.col0;
        P1Src<Cfg, B_KC> src;
        src.seg = segs + g.seg_begin, src.seg_end = segs + g.seg_end, src.c = c;
        src.row0 = row0, src.col0 = col0;
        src.m_valid = g.m0 - row0, src.n_valid = g.n0 - col0;
        src.nsteps = un.pad; // BK-deep stages of the unit (counted by the host: no dependent loads here)
        src.open();
        const int mi_n = clampi((src.m_valid - wm0 + 7) / 8, 0, Cfg::MI);
        const int ni_n = clampi((src.n_valid - wn0 + 7) / 8, 0, Cfg::NI);
        double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++)
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        mainloop<Cfg, true, B_KC, true>(src, smem, s_alpha, acc, wm0, wn0, mi_n, ni_n,
                                        src.m_valid >= Cfg::BM && src.n_valid >= Cfg::BN);
        double *w = wbuf + g.w_off;
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++) {
                const int r = row0 + wm0 + mi * 8 + lr, cc = col0 + wn0 + ni * 8 + lc * 2;
                if (mi < mi_n && ni < ni_n && r < g.m0) {
                    if (cc < g.n0)
                        w[(size_t)r * g.wld + cc] = acc[mi][ni][0];
                    if (cc + 1 < g.n0)
                        w[(size_t)r * g.wld + cc + 1] = acc[mi][ni][1];
                }
            }
        u = s_unit; // written before the barriers of the main loop
        __syncthreads();
    }
    pdl_exit();
}

// ------------------------------ phase 2 -------------------------------------
template <class Cfg, bool A_KC> struct P2Src {
    const P2Seg *seg, *seg_end;
    const double *wbuf;
    const double *a, *b;
    P2Seg nxt; // descriptor of the following segment, fetched one segment ahead
    int lda, wld, row0, col0, m_valid, n_valid, n_lo, n_hi, k_left, nsteps;
    __device__ int steps() const { return nsteps; }
    __device__ void use(const P2Seg &s) {
        lda = s.lda, wld = s.wld;
        a = A_KC ? s.a1 + (size_t)row0 * lda : s.a1 + row0;
        // tile column j is panel column col0 + j; the W panel of this segment holds [col_lo, col_hi) only
        b = wbuf + s.w_off + (col0 - s.col_lo);
        n_lo = max(0, s.col_lo - col0), n_hi = min(n_valid, s.col_hi - col0);
        k_left = s.klen;
        if (seg + 1 < seg_end)
            nxt = seg[1];
    }
    __device__ void open() { use(*seg); }
    __device__ double issue(double *As, double *Bs) {
        load_tile<Cfg::BM, Cfg::THREADS, A_KC, Cfg::BK>(As, a, lda, m_valid, k_left);
        load_tile<Cfg::BN, Cfg::THREADS, false, Cfg::BK>(Bs, b, wld, n_hi, k_left, n_lo);
        k_left -= Cfg::BK;
        if (k_left > 0) {
            a += A_KC ? Cfg::BK : (size_t)Cfg::BK * lda;
            b += (size_t)Cfg::BK * wld;
        } else if (++seg < seg_end)
            use(nxt);
        return 1.0;
    }
};

template <class Cfg, bool A_KC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
phase2_kernel(const P2Window *__restrict__ wins, const P2Seg *__restrict__ segs, const Unit *__restrict__ units,
              int n_units, unsigned int *__restrict__ counter, const double *__restrict__ wbuf,
              double *__restrict__ v, double scale, double *__restrict__ pbuf, int wait_first,
              const MatvecArgs *__restrict__ ind) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_unit;
    if (ind)
        v = ind->v, scale = ind->scale;
    pdl_enter(wait_first);
    int wm0, wn0;
    warp_origin<Cfg>(wm0, wn0);
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    // Units are claimed one ahead (the atomic for the next unit is in flight while this one runs), the first one
    // included: under programmatic dependent launch the CTAs of a grid start as slots free up, and a CTA that
    // starts late must not sit on one of the most expensive units (the list is in cost order).
    if (threadIdx.x == 0)
        s_unit = (int)atomicAdd(counter, 1u);
    __syncthreads();
    int u = s_unit;
    __syncthreads();
    while (u < n_units) {
        if (threadIdx.x == 0)
            s_unit = (int)atomicAdd(counter, 1u);
        const Unit un = units[u];
        const P2Window win = wins[un.idx];
        P2Src<Cfg, A_KC> src;
        src.seg = segs + un.seg_begin, src.seg_end = segs + un.seg_end, src.wbuf = wbuf;
        src.row0 = un.row0, src.col0 = un.col0;
        src.m_valid = win.m1 - src.row0, src.n_valid = win.n0 - src.col0;
        src.nsteps = un.pad; // BK-deep stages of the unit (counted by the host: no dependent loads here)
        src.open();
        const int mi_n = clampi((src.m_valid - wm0 + 7) / 8, 0, Cfg::MI);
        const int ni_n = clampi((src.n_valid - wn0 + 7) / 8, 0, Cfg::NI);
        double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ni++)
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        mainloop<Cfg, A_KC, false, false>(src, smem, nullptr, acc, wm0, wn0, mi_n, ni_n,
                                          src.m_valid >= Cfg::BM && src.n_valid >= Cfg::BN);
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
    pdl_exit();
}

// ------------------------------ sigma reduce --------------------------------
// Deterministic mode: sigma tile += scale * sum over its K-chunk partials, in chunk order; every
// sigma element is written by exactly one thread, so the matvec is bit-reproducible run to run.
constexpr int RED_SPLIT = 8; // CTAs per sigma tile
__global__ void __launch_bounds__(256)
reduce_kernel(const OutTile *__restrict__ tiles, int n_tiles, const int64_t *__restrict__ part_off,
              const P2Window *__restrict__ wins, const double *__restrict__ pbuf, double *__restrict__ v,
              double scale, const MatvecArgs *__restrict__ ind) {
    if (ind)
        v = ind->v, scale = ind->scale;
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
// remainder, and into 64-column tiles plus 16- or 8-column strips for the remainder; a remainder of
// 1..8 columns widens the last 64-column tile to 72 instead of opening an 8-column strip that would
// stream the whole A operand again for a ninth of the work.
using Cfg0 = TileCfg<128, 64, 4, 2, 3, 2>; // 8 warps, 32x32 warp tiles (bulk of large sectors)
using Cfg1 = TileCfg<64, 64, 2, 2, 3, 3>; // 4 warps, 32x32
using Cfg2 = TileCfg<128, 16, 8, 1, 4>; // 8 warps, 16x16 (column remainders)
using Cfg3 = TileCfg<64, 16, 4, 1, 4>;  // 4 warps, 16x16
using Cfg4 = TileCfg<128, 8, 8, 1, 4>;  // 8 warps, 16x8  (narrow panels, width <= 8)
using Cfg5 = TileCfg<64, 8, 4, 1, 4>;   // 4 warps, 16x8
using Cfg6 = TileCfg<128, 72, 8, 1, 3, 2>; // 8 warps, 16x72 (64 + a remainder of 1..8 columns)
using Cfg7 = TileCfg<64, 72, 4, 1, 3, 3>;  // 4 warps, 16x72
constexpr int NCFG = 8;

struct CfgInfo {
    int bm, bn, threads, smem;
};
static const CfgInfo kCfg[NCFG] = {{Cfg0::BM, Cfg0::BN, Cfg0::THREADS, Cfg0::SMEM_BYTES},
                                   {Cfg1::BM, Cfg1::BN, Cfg1::THREADS, Cfg1::SMEM_BYTES},
                                   {Cfg2::BM, Cfg2::BN, Cfg2::THREADS, Cfg2::SMEM_BYTES},
                                   {Cfg3::BM, Cfg3::BN, Cfg3::THREADS, Cfg3::SMEM_BYTES},
                                   {Cfg4::BM, Cfg4::BN, Cfg4::THREADS, Cfg4::SMEM_BYTES},
                                   {Cfg5::BM, Cfg5::BN, Cfg5::THREADS, Cfg5::SMEM_BYTES},
                                   {Cfg6::BM, Cfg6::BN, Cfg6::THREADS, Cfg6::SMEM_BYTES},
                                   {Cfg7::BM, Cfg7::BN, Cfg7::THREADS, Cfg7::SMEM_BYTES}};

static inline int cfg_of(int bm, int bn) {
    return (bn == 64 ? 0 : bn == 16 ? 2 : bn == 8 ? 4 : 6) + (bm == 128 ? 0 : 1);
}

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
// columns: 64-column tiles (the last one 72 wide when 1..8 columns would be left over), then 16- / 8-column
// strips for what remains
static std::vector<Strip> split_cols(int n) {
    static const bool wide = getenv("B2G_NO_TILE72") == nullptr;
    std::vector<Strip> out;
    int c = 0;
    while (n - c > 32) {
        const int left = n - c - 64; // columns after a 64-wide tile here
        const bool last64 = left <= 32; // no further 64-wide tile follows
        if (wide && last64 && ((left >= 1 && left <= 8) || (left >= 17 && left <= 24))) {
            out.push_back(Strip{c, 72});
            c += 72;
        } else {
            out.push_back(Strip{c, 64});
            c += 64;
        }
    }
    while (n - c > 8) {
        out.push_back(Strip{c, 16});
        c += 16;
    }
    if (n - c > 0)
        out.push_back(Strip{c, 8});
    return out;
}

struct LaunchGroup { // one kernel launch: units of one (slab, phase, cfg, layout)
    int slab = 0;
    int phase, cfg, layout;
    Unit *d_units = nullptr;
    int n_units = 0;
    double flops = 0; // useful 2*m*n*k FLOPs of the launch (valid rows x cols only)
};

struct TiledPlan {
    b2g_context *ctx = nullptr;
    // deterministic sigma accumulation
    OutTile *d_tiles = nullptr;
    int64_t *d_part_off = nullptr;
    double *d_pbuf = nullptr;
    int n_tiles = 0;
    // row panels may overlap in sigma (a block and its row sub-windows): tiles are ordered by the colour of
    // their panel in the interval-overlap graph and each colour is reduced by its own launch
    std::vector<std::pair<int, int>> tile_ranges;
    size_t pbuf_doubles = 0;
    P1Group *d_p1g = nullptr;
    P1Seg *d_p1s = nullptr;
    P2Window *d_win = nullptr;
    P2Seg *d_seg = nullptr;
    double *d_wbuf = nullptr;
    unsigned int *d_counters = nullptr;
    size_t wbuf_doubles = 0;
    // W slabs: the row panels are cut into groups whose W panels fit a bounded workspace; a slab runs its phase 1,
    // then its phase 2, and the next slab reuses the same workspace (zeroed again: slab_doubles[s] of it)
    std::vector<size_t> slab_doubles;
    int n_counters = 0;
    // small lists (launch-bound): the whole matvec - counters, forked phase-1 launches, join, phase-2 launches,
    // join, sigma reduce - is captured once into a CUDA graph and replayed with one launch; c, sigma and scale of
    // a replay come through d_args
    MatvecArgs *d_args = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_state = 0; // 0 = first call runs eagerly (function attributes, occupancy), 1 = capture next, 2 = replay, -1 = off
    int graph_kernels = 0;
    std::vector<LaunchGroup> groups;
    std::vector<void *> to_free;
};

// the dynamic shared memory limit is a per-device function attribute: one bit per device ordinal
template <class K> static int raise_smem_limit(K kern, int bytes, int device, std::atomic<uint64_t> &mask) {
    const uint64_t bit = (uint64_t)1 << (device & 63);
    if (!(mask.load(std::memory_order_acquire) & bit)) {
        B2G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        mask.fetch_or(bit, std::memory_order_release);
    }
    return 0;
}

// pdl: this launch follows another launch of the same chain on the same stream and may overlap its tail
template <class... Params, class... Args>
static int launch_ex(void (*kern)(Params...), int grid, int threads, size_t smem, cudaStream_t stream, bool pdl,
                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdl ? 1 : 0;
    B2G_CUDA(cudaLaunchKernelEx(&cfg, kern, Params(args)...));
    return 0;
}

template <class Cfg, bool L> static int launch_p1(const LaunchGroup &g, const TiledPlan &tp, b2g_context *ctx,
                                                  cudaStream_t stream, unsigned int *counter, const double *c,
                                                  bool pdl, int wait_first, const MatvecArgs *ind) {
    auto kern = phase1_kernel<Cfg, L>;
    static std::atomic<uint64_t> attr_mask{0};
    if (raise_smem_limit(kern, Cfg::SMEM_BYTES, ctx->device, attr_mask))
        return 1;
    static std::atomic<int> per_sm_cached{0}; // same on every B200 of the box
    int per_sm = per_sm_cached.load(std::memory_order_relaxed);
    if (per_sm == 0) {
        per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM_BYTES);
        per_sm_cached.store(per_sm, std::memory_order_relaxed);
    }
    const int grid = std::min(g.n_units, ctx->sm_count * std::max(per_sm, 1));
    return launch_ex(kern, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, pdl, tp.d_p1g, tp.d_p1s, g.d_units, g.n_units,
                     counter, c, tp.d_wbuf, wait_first, ind);
}
template <class Cfg, bool L> static int launch_p2(const LaunchGroup &g, const TiledPlan &tp, b2g_context *ctx,
                                                  cudaStream_t stream, unsigned int *counter, double *v,
                                                  double scale, bool pdl, int wait_first, const MatvecArgs *ind) {
    auto kern = phase2_kernel<Cfg, L>;
    static std::atomic<uint64_t> attr_mask{0};
    if (raise_smem_limit(kern, Cfg::SMEM_BYTES, ctx->device, attr_mask))
        return 1;
    static std::atomic<int> per_sm_cached{0}; // same on every B200 of the box
    int per_sm = per_sm_cached.load(std::memory_order_relaxed);
    if (per_sm == 0) {
        per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM_BYTES);
        per_sm_cached.store(per_sm, std::memory_order_relaxed);
    }
    const int grid = std::min(g.n_units, ctx->sm_count * std::max(per_sm, 1));
    return launch_ex(kern, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, pdl, tp.d_win, tp.d_seg, g.d_units, g.n_units,
                     counter, tp.d_wbuf, v, scale, tp.d_pbuf, wait_first, ind);
}

} // namespace b2g

using namespace b2g;

namespace {
struct Key4 {
    uint64_t a, b, c, d;
    bool operator==(const Key4 &o) const { return a == o.a && b == o.b && c == o.c && d == o.d; }
};
struct Key4Hash {
    size_t operator()(const Key4 &k) const {
        uint64_t h = k.a * 0x9E3779B97F4A7C15ull;
        h = (h ^ (h >> 29)) + k.b * 0xBF58476D1CE4E5B9ull;
        h = (h ^ (h >> 31)) + k.c * 0x94D049BB133111EBull;
        h = (h ^ (h >> 27)) + k.d * 0xD6E8FEB86659FD93ull;
        return (size_t)(h ^ (h >> 32));
    }
};
// Key4 -> int, open addressing with linear probing (the regrouping does a few lookups per pair of lists with
// 10^5 pairs, once per site: node-based maps were a third of the plan construction time)
struct FlatMap4 {
    std::vector<Key4> keys;
    std::vector<int> vals; // -1 = empty
    size_t mask = 0;
    explicit FlatMap4(size_t expected) {
        size_t cap = 16;
        while (cap < 2 * expected + 2)
            cap <<= 1;
        keys.resize(cap), vals.assign(cap, -1), mask = cap - 1;
    }
    // value of key, inserting `fresh` when absent; second = inserted
    std::pair<int, bool> get_or_insert(const Key4 &k, int fresh) {
        size_t i = Key4Hash()(k) & mask;
        while (vals[i] >= 0) {
            if (keys[i] == k)
                return std::make_pair(vals[i], false);
            i = (i + 1) & mask;
        }
        keys[i] = k, vals[i] = fresh;
        return std::make_pair(fresh, true);
    }
};
} // namespace

void b2g_tiled_destroy(void *h) {
    TiledPlan *tp = (TiledPlan *)h;
    if (!tp)
        return;
    if (tp->graph_exec)
        cudaGraphExecDestroy(tp->graph_exec);
    for (void *p : tp->to_free)
        if (tp->ctx)
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
    static const bool merge_cols = getenv("B2G_NO_PANELS") == nullptr; // A/B switch: one panel per window
    static const bool verbose = getenv("B2G_VERBOSE") != nullptr;
    auto tstart = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (verbose || b2g_prof_enabled()) {
            auto now = std::chrono::steady_clock::now();
            if (verbose)
                fprintf(stderr, "[b2g] tiled_build %-14s %8.3f ms\n", what,
                        std::chrono::duration<double, std::milli>(now - tstart).count());
            b2g_prof_record((std::string("tiled_build.") + what).c_str(),
                            std::chrono::duration<double>(now - tstart).count());
            tstart = now;
        }
    };

    // ---- 1. sigma blocks: connected sets of overlapping windows; a window is (row, column) placed in its block
    struct Win {
        int64_t c_off;
        int32_t m1, n0, ldc;
        int64_t lo, hi;
        int block, row, col, panel, layer;
    };
    FlatMap4 wid(n);
    std::vector<Win> wv;
    std::vector<int> pair_win(n, -1);
    for (size_t i = 0; i < n; i++) {
        const B2GPair &q = hp[i];
        if (q.m0 == 0 || q.n0 == 0 || q.m1 == 0)
            continue;
        const Key4 key{(uint64_t)q.c1_off, (uint64_t)(uint32_t)q.m1, (uint64_t)(uint32_t)q.n0, (uint64_t)(uint32_t)q.ldc1};
        const std::pair<int, bool> it = wid.get_or_insert(key, (int)wv.size());
        if (it.second)
            wv.push_back(Win{q.c1_off, q.m1, q.n0, q.ldc1, q.c1_off,
                             q.c1_off + (int64_t)(q.m1 - 1) * q.ldc1 + q.n0, -1, 0, 0, -1, 0});
        pair_win[i] = it.first;
    }
    {
        std::vector<int> order(wv.size());
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&wv](int x, int y) { return wv[x].lo < wv[y].lo; });
        int nb = 0;
        int64_t cur_hi = 0, cur_lo = 0;
        int cur_ld = 0;
        bool uniform = true;
        std::vector<int> members;
        auto close = [&]() {
            // a block whose windows share one pitch is a matrix: (row, col) of every window inside it
            for (int w : members) {
                wv[w].block = nb;
                if (uniform && merge_cols) {
                    wv[w].row = (int)((wv[w].lo - cur_lo) / cur_ld), wv[w].col = (int)((wv[w].lo - cur_lo) % cur_ld);
                    if (wv[w].col + wv[w].n0 > cur_ld) // wraps around the pitch: not a sub-matrix
                        wv[w].row = -1;
                } else
                    wv[w].row = -1;
            }
            members.clear();
            nb++;
        };
        for (int w : order) {
            if (!members.empty() && wv[w].lo >= cur_hi)
                close();
            if (members.empty())
                cur_lo = wv[w].lo, cur_hi = wv[w].hi, cur_ld = wv[w].ldc, uniform = true;
            else
                cur_hi = std::max(cur_hi, wv[w].hi), uniform = uniform && wv[w].ldc == cur_ld;
            members.push_back(w);
        }
        if (!members.empty())
            close();
    }
    lap("windows");
    // ---- 2. row panels: windows of one block with the same row range (merged along the columns);
    //         windows that overlap in columns inside a panel go to different layers (their W never share a slot)
    struct Panel {
        int64_t c_off; // sigma offset of (row 0, column col_lo)
        int32_t m1, ldc, col_lo, col_hi;
        std::vector<int> wins;
    };
    std::vector<Panel> panels;
    {
        std::map<std::tuple<int, int, int>, int> pid; // (block, row, m1)
        for (size_t w = 0; w < wv.size(); w++) {
            Win &x = wv[w];
            if (x.row < 0) { // irregular: a panel of its own
                x.panel = (int)panels.size(), x.col = 0;
                panels.push_back(Panel{x.c_off, x.m1, x.ldc, 0, x.n0, {(int)w}});
                continue;
            }
            auto key = std::make_tuple(x.block, x.row, x.m1);
            auto it = pid.find(key);
            if (it == pid.end()) {
                it = pid.emplace(key, (int)panels.size()).first;
                panels.push_back(Panel{x.c_off - x.col, x.m1, x.ldc, x.col, x.col + x.n0, {}});
            }
            Panel &pn = panels[it->second];
            pn.col_lo = std::min(pn.col_lo, x.col), pn.col_hi = std::max(pn.col_hi, x.col + x.n0);
            pn.wins.push_back((int)w);
            x.panel = it->second;
        }
        for (Panel &pn : panels) {
            if (pn.wins.size() > 1 || (pn.wins.size() == 1 && wv[pn.wins[0]].row >= 0)) {
                // columns relative to the panel origin; layers by interval colouring over the column ranges
                std::sort(pn.wins.begin(), pn.wins.end(), [&wv](int x, int y) { return wv[x].col < wv[y].col; });
                std::vector<int> layer_end; // last column used per layer
                for (int w : pn.wins) {
                    int l = 0;
                    while (l < (int)layer_end.size() && layer_end[l] > wv[w].col)
                        l++;
                    if (l == (int)layer_end.size())
                        layer_end.push_back(0);
                    layer_end[l] = wv[w].col + wv[w].n0;
                    wv[w].layer = l;
                }
                pn.c_off += pn.col_lo;
                for (int w : pn.wins)
                    wv[w].col -= pn.col_lo;
                pn.col_hi -= pn.col_lo, pn.col_lo = 0;
            }
        }
    }
    lap("panels");
    // ---- 3. phase-2 segments = (panel, layer, A1 block, layout); phase-1 groups = (segment, window, tb0)
    struct HostSeg {
        int panel, layer, ta1, lda1, m0;
        const double *a1;
        int64_t w_off;
        int wld, col_lo, col_hi;
    };
    FlatMap4 sid(n);
    std::vector<HostSeg> hsegs;
    // (segment, window) -> group; a (segment, window) fed through both operand layouts of B0 needs two W slots:
    // the second layout gets a twin segment (same A1, K longer by m0) so that no W element is written twice
    struct HostGroup {
        int seg, win, tb0;
        int count, first; // pairs of the group: pair_of[first .. first + count)
        int64_t ksum;     // sum of k0 over the pairs
        int nsteps;       // BK-deep stages (every tile configuration has BK = 16)
    };
    std::vector<int> pair_group(n, -1);
    FlatMap4 gid(2 * n);
    std::vector<HostGroup> hgroups;
    auto seg_of = [&](int panel, int layer, const B2GPair &q, int variant) -> int {
        const int ta1 = (q.flags & B2G_F_TA1) ? 1 : 0;
        const Key4 sk{(uint64_t)(uintptr_t)q.a1, ((uint64_t)(uint32_t)panel << 32) | (uint32_t)q.lda1,
                      ((uint64_t)(uint32_t)q.m0 << 32) | (uint32_t)layer, (uint64_t)(ta1 | (variant << 1))};
        const std::pair<int, bool> r = sid.get_or_insert(sk, (int)hsegs.size());
        if (r.second)
            hsegs.push_back(HostSeg{panel, layer, ta1, q.lda1, q.m0, q.a1, 0, 0, INT32_MAX, 0});
        return r.first;
    };
    for (size_t i = 0; i < n; i++) {
        if (pair_win[i] < 0)
            continue;
        const B2GPair &q = hp[i];
        const Win &x = wv[pair_win[i]];
        const int tb0 = (q.flags & B2G_F_TB0) ? 1 : 0;
        int seg = seg_of(x.panel, x.layer, q, 0);
        const Key4 gk{(uint64_t)(uint32_t)seg, (uint64_t)(uint32_t)pair_win[i], 0, 0};
        std::pair<int, bool> r = gid.get_or_insert(gk, (int)hgroups.size());
        int g = r.first;
        if (r.second)
            hgroups.push_back(HostGroup{seg, pair_win[i], tb0, 0, 0, 0, 0});
        else if (hgroups[g].tb0 != tb0) { // the other layout: same window, twin segment
            seg = seg_of(x.panel, x.layer, q, 1);
            const Key4 gk2{(uint64_t)(uint32_t)seg, (uint64_t)(uint32_t)pair_win[i], 0, 0};
            r = gid.get_or_insert(gk2, (int)hgroups.size());
            g = r.first;
            if (r.second)
                hgroups.push_back(HostGroup{seg, pair_win[i], tb0, 0, 0, 0, 0});
        }
        HostSeg &hs = hsegs[hgroups[g].seg];
        hs.col_lo = std::min(hs.col_lo, x.col), hs.col_hi = std::max(hs.col_hi, x.col + x.n0);
        hgroups[g].count++, hgroups[g].ksum += q.k0, hgroups[g].nsteps += (q.k0 + 15) / 16;
        pair_group[i] = g;
    }
    // pairs by group, in list order (= ascending a0_off: neighbours read the same wavefunction window)
    std::vector<int> pair_of(n);
    {
        int acc = 0;
        for (HostGroup &hg : hgroups)
            hg.first = acc, acc += hg.count, hg.count = 0;
        for (size_t i = 0; i < n; i++)
            if (pair_group[i] >= 0) {
                HostGroup &hg = hgroups[pair_group[i]];
                pair_of[hg.first + hg.count++] = (int)i;
            }
    }
    lap("groups.hash");
    // W panels: m0 x wld per segment, wld = the column span its groups cover (even, so that rows stay 16-byte
    // aligned); columns of the span no group writes stay zero from the one memset below
    // The workspace is bounded (B2G_WCAP_GB, default 24 GB): at the NC/CN switch site of Cr2 M=4000 the W panels
    // of one H_eff are 75+ GB next to 96 GB of operators.  Row panels go to slabs in panel order; all segments of
    // a panel share its slab.
    const char *env_cap = getenv("B2G_WCAP_GB");
    const size_t wcap = (size_t)((env_cap ? atof(env_cap) : 24.0) * 1e9 / 8.0);
    std::vector<size_t> panel_w(panels.size(), 0);
    for (HostSeg &hs : hsegs) {
        hs.col_lo &= ~1;
        hs.wld = ((hs.col_hi - hs.col_lo) + 1) & ~1;
        panel_w[hs.panel] += (size_t)hs.m0 * hs.wld;
    }
    std::vector<int> panel_slab(panels.size(), 0);
    {
        size_t acc = 0;
        int slab = 0;
        for (size_t k = 0; k < panels.size(); k++) {
            if (acc != 0 && acc + panel_w[k] > wcap)
                slab++, acc = 0;
            panel_slab[k] = slab, acc += panel_w[k];
        }
        tp->slab_doubles.assign((size_t)slab + 1, 0);
    }
    for (HostSeg &hs : hsegs) {
        size_t &cur = tp->slab_doubles[panel_slab[hs.panel]];
        hs.w_off = (int64_t)cur;
        cur += (size_t)hs.m0 * hs.wld;
    }
    size_t woff = 0;
    for (size_t v : tp->slab_doubles)
        woff = std::max(woff, v);
    tp->wbuf_doubles = woff;
    const int n_slabs = (int)tp->slab_doubles.size();
    std::vector<P1Group> p1g;
    std::vector<P1Seg> p1s;
    p1g.reserve(hgroups.size()), p1s.reserve(n);
    for (HostGroup &hg : hgroups) {
        const HostSeg &hs = hsegs[hg.seg];
        const Win &x = wv[hg.win];
        P1Group g;
        g.w_off = hs.w_off + (x.col - hs.col_lo), g.wld = hs.wld, g.m0 = hs.m0, g.n0 = x.n0;
        g.seg_begin = (int)p1s.size();
        // neighbours read the same wavefunction window (L2 reuse of c inside the group)
        if (hg.count > 1)
            std::stable_sort(pair_of.begin() + hg.first, pair_of.begin() + hg.first + hg.count,
                             [&hp](int a, int b) { return hp[a].a0_off < hp[b].a0_off; });
        for (int z = 0; z < hg.count; z++) {
            const B2GPair &q = hp[pair_of[hg.first + z]];
            p1s.push_back(P1Seg{q.b0, q.a0_off, q.alpha0 * q.alpha1, q.lda0, q.ldb0, q.k0, 0});
        }
        g.seg_end = (int)p1s.size(), g.pad = 0;
        p1g.push_back(g);
    }

    lap("groups");
    // ---- 4. units
    struct HostUnit {
        Unit u;
        double cost, flops;
    };
    // (phase, cfg, layout) -> units; flat table (a map lookup per unit costs more than the unit itself)
    constexpr int N_CFG = 8;
    std::vector<std::vector<HostUnit>> unit_tab((size_t)n_slabs * 2 * N_CFG * 2);
    auto units_of = [&unit_tab](int slab, int phase, int cfg, int layout) -> std::vector<HostUnit> & {
        return unit_tab[(((size_t)slab * 2 + (phase - 1)) * N_CFG + cfg) * 2 + layout];
    };
    for (size_t gi = 0; gi < hgroups.size(); gi++) {
        const HostGroup &hg = hgroups[gi];
        const P1Group &g = p1g[gi];
        const int64_t ksum = hg.ksum;
        const int nsteps = hg.nsteps;
        const std::vector<Strip> rsv1 = split_rows(g.m0), csv1 = split_cols(g.n0);
        for (const Strip &rs : rsv1)
            for (const Strip &cs : csv1) {
                const int c = cfg_of(rs.tile, cs.tile);
                units_of(panel_slab[hsegs[hg.seg].panel], 1, c, hg.tb0).push_back(HostUnit{
                    Unit{(int)gi, rs.origin, cs.origin, 0, 0, nsteps, -1},
                    (double)rs.tile * cs.tile * (double)(ksum + 32 * (int64_t)hg.count),
                    2.0 * std::min(rs.tile, g.m0 - rs.origin) * std::min(cs.tile, g.n0 - cs.origin) * (double)ksum});
            }
    }
    lap("units.p1");
    std::vector<P2Window> wins(panels.size());
    for (size_t k = 0; k < panels.size(); k++)
        wins[k] = P2Window{panels[k].c_off, panels[k].ldc, panels[k].m1, panels[k].col_hi - panels[k].col_lo, 0};
    // segments of a panel, per layout, neighbours share the operator block (L2 reuse across K-chunks)
    std::vector<std::vector<int>> pseg[2];
    pseg[0].resize(panels.size()), pseg[1].resize(panels.size());
    for (size_t k = 0; k < hsegs.size(); k++)
        if (hsegs[k].col_hi > hsegs[k].col_lo)
            pseg[hsegs[k].ta1][hsegs[k].panel].push_back((int)k);
    for (int lay = 0; lay < 2; lay++)
        for (auto &lst : pseg[lay])
            std::sort(lst.begin(), lst.end(), [&hsegs](int x, int y) { return hsegs[x].a1 < hsegs[y].a1; });
    // K-chunk: 2048 for the big lists; shorter when the list is small so that phase 2 still
    // spreads over the whole chip (few panels, each with a long chain of short segments)
    int64_t kchunk_eff = kchunk;
    if (!env_kc) {
        double tile_k = 0;
        for (int lay = 0; lay < 2; lay++)
            for (size_t w = 0; w < panels.size(); w++) {
                double ks = 0;
                for (int k : pseg[lay][w])
                    ks += hsegs[k].m0;
                tile_k += ks * (double)split_rows(wins[w].m1).size() * (double)split_cols(wins[w].n0).size();
            }
        const double want_units = 24.0 * (ctx ? ctx->sm_count : 148);
        kchunk_eff = (int64_t)std::min<double>(2048.0, std::max(128.0, tile_k / want_units));
        kchunk_eff = (kchunk_eff + 15) / 16 * 16;
    }
    lap("units.kchunk");
    std::vector<P2Seg> segs;
    for (int lay = 0; lay < 2; lay++)
        for (size_t w = 0; w < panels.size(); w++) {
            const std::vector<int> &lst = pseg[lay][w];
            if (lst.empty() || wins[w].m1 == 0 || wins[w].n0 == 0)
                continue;
            const std::vector<Strip> rsv = split_rows(wins[w].m1), csv = split_cols(wins[w].n0);
            // one filtered segment list per column tile: a tile skips the operator blocks whose W panel has
            // nothing in its columns (block-sparse B)
            for (const Strip &cs : csv) {
                const int t_lo = cs.origin, t_hi = std::min(cs.origin + cs.tile, wins[w].n0);
                size_t s0 = segs.size();
                int64_t ksum = 0;
                int nsteps = 0;
                auto flush = [&](size_t s1) {
                    if (s1 == s0)
                        return;
                    for (const Strip &rs : rsv)
                        units_of(panel_slab[w], 2, cfg_of(rs.tile, cs.tile), lay).push_back(
                            HostUnit{Unit{(int)w, rs.origin, cs.origin, (int)s0, (int)s1, nsteps, -1},
                                     (double)rs.tile * cs.tile * (double)(ksum + 32),
                                     2.0 * std::min(rs.tile, wins[w].m1 - rs.origin) * (t_hi - t_lo) * (double)ksum});
                    s0 = s1, ksum = 0, nsteps = 0;
                };
                for (int k : lst) {
                    const HostSeg &hs = hsegs[k];
                    if (hs.col_hi <= t_lo || hs.col_lo >= t_hi || hs.m0 == 0)
                        continue;
                    segs.push_back(P2Seg{hs.a1, hs.w_off, hs.lda1, hs.m0, hs.wld, hs.col_lo, hs.col_hi, 0});
                    ksum += hs.m0, nsteps += (hs.m0 + 15) / 16;
                    if (ksum >= kchunk_eff)
                        flush(segs.size());
                }
                flush(segs.size());
            }
        }

    // non-empty (slab, phase, cfg, layout), in launch order
    std::map<std::tuple<int, int, int, int>, std::vector<HostUnit>> groups;
    for (int sl = 0; sl < n_slabs; sl++)
        for (int ph = 1; ph <= 2; ph++)
            for (int c = 0; c < N_CFG; c++)
                for (int lay = 0; lay < 2; lay++)
                    if (!units_of(sl, ph, c, lay).empty())
                        groups[std::make_tuple(sl, ph, c, lay)].swap(units_of(sl, ph, c, lay));
    lap("units");
    // ---- 5. sigma tiles, partial slots, unit order (host)
    // deterministic sigma accumulation: one partial slot per phase-2 unit, grouped by sigma tile
    // (a tile collects partials from both operand layouts), slots in K-chunk order
    const char *env_atomic = getenv("B2G_ATOMIC_SIGMA");
    std::vector<OutTile> tiles;
    std::vector<int64_t> part_off;
    if (!(env_atomic && env_atomic[0] == '1')) {
        // colour the panels so that panels sharing sigma elements never share a launch
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
        // phase-2 units by (colour of the panel, panel, row0, col0), the partials of a tile in K-chunk order
        // (= segment order; the segment ranges of the two operand layouts are disjoint)
        struct TileRef {
            int colour, panel, row0, col0, seg_begin, cfg;
            uint32_t seq;
            HostUnit *hu;
        };
        std::vector<TileRef> refs;
        {
            size_t n2 = 0;
            for (auto &kv : groups)
                if (std::get<1>(kv.first) == 2)
                    n2 += kv.second.size();
            refs.reserve(n2);
            for (auto &kv : groups)
                if (std::get<1>(kv.first) == 2)
                    for (HostUnit &hu : kv.second)
                        refs.push_back(TileRef{colour[hu.u.idx], hu.u.idx, hu.u.row0, hu.u.col0, hu.u.seg_begin,
                                               std::get<2>(kv.first), (uint32_t)refs.size(), &hu});
        }
        std::sort(refs.begin(), refs.end(), [](const TileRef &x, const TileRef &y) {
            if (x.colour != y.colour)
                return x.colour < y.colour;
            if (x.panel != y.panel)
                return x.panel < y.panel;
            if (x.row0 != y.row0)
                return x.row0 < y.row0;
            if (x.col0 != y.col0)
                return x.col0 < y.col0;
            if (x.seg_begin != y.seg_begin)
                return x.seg_begin < y.seg_begin;
            return x.seq < y.seq;
        });
        size_t poff = 0;
        for (size_t a = 0; a < refs.size();) {
            size_t b = a + 1;
            while (b < refs.size() && refs[b].panel == refs[a].panel && refs[b].row0 == refs[a].row0 &&
                   refs[b].col0 == refs[a].col0)
                b++;
            const int col = refs[a].colour;
            while ((int)tp->tile_ranges.size() <= col)
                tp->tile_ranges.push_back(std::make_pair((int)tiles.size(), (int)tiles.size()));
            const int c = refs[a].cfg;
            OutTile ot{refs[a].panel, refs[a].row0, refs[a].col0, kCfg[c].bm, kCfg[c].bn, (int)part_off.size(), 0, 0};
            for (size_t z = a; z < b; z++) {
                refs[z].hu->u.poff = (int64_t)poff;
                part_off.push_back((int64_t)poff);
                poff += (size_t)kCfg[c].bm * kCfg[c].bn;
            }
            ot.part_end = (int)part_off.size();
            tiles.push_back(ot);
            tp->tile_ranges[col].second = (int)tiles.size();
            a = b;
        }
        tp->pbuf_doubles = poff;
        p->stats.workspace_doubles = (int64_t)(tp->wbuf_doubles + tp->pbuf_doubles);
        tp->n_tiles = (int)tiles.size();
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
        g.slab = std::get<0>(kv.first), g.phase = std::get<1>(kv.first), g.cfg = std::get<2>(kv.first);
        g.layout = std::get<3>(kv.first);
        g.n_units = (int)hu.size();
        for (const HostUnit &x : hu)
            g.flops += x.flops;
        tp->groups.push_back(g);
    }
    lap("tiles");
    if (ctx == nullptr) { // b2g_debug_tiled_plan: the host-side regrouping alone (no device)
        int64_t nu = 0;
        for (auto &kv : groups)
            nu += (int64_t)kv.second.size();
        p->stats.launches = (int64_t)tp->groups.size(), p->stats.n_large = nu;
        // fingerprint of everything the device would get (planner changes must not change it): FNV-1a over the
        // unit lists with their partial-slot offsets, the sigma tiles and the slot table
        uint64_t h = 1469598103934665603ull;
        auto mix = [&h](const void *ptr, size_t bytes) {
            const unsigned char *q = (const unsigned char *)ptr;
            for (size_t i = 0; i < bytes; i++)
                h = (h ^ q[i]) * 1099511628211ull;
        };
        for (auto &k : keep)
            for (const Unit &u : k) { // field by field: the struct has padding
                mix(&u.idx, 4), mix(&u.row0, 4), mix(&u.col0, 4), mix(&u.seg_begin, 4), mix(&u.seg_end, 4);
                mix(&u.pad, 4), mix(&u.poff, 8);
            }
        for (const OutTile &t : tiles)
            mix(&t, sizeof(OutTile));
        mix(part_off.data(), part_off.size() * sizeof(int64_t));
        for (const P2Seg &g : segs)
            mix(&g.w_off, 8), mix(&g.klen, 4), mix(&g.col_lo, 4), mix(&g.col_hi, 4);
        p->stats.arenas = (int64_t)(h >> 1); // debug entry only: reported as the fingerprint
        return 0;
    }
    // ---- 6. upload
    auto upload = [&](const void *src, size_t bytes, void **dst) -> int {
        if (b2g_dmalloc(ctx, dst, bytes))
            return 1;
        tp->to_free.push_back(*dst);
        if (bytes)
            B2G_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };
    if (upload(p1g.data(), p1g.size() * sizeof(P1Group), (void **)&tp->d_p1g))
        return 1;
    if (upload(p1s.data(), p1s.size() * sizeof(P1Seg), (void **)&tp->d_p1s))
        return 1;
    if (upload(wins.data(), wins.size() * sizeof(P2Window), (void **)&tp->d_win))
        return 1;
    if (upload(segs.data(), segs.size() * sizeof(P2Seg), (void **)&tp->d_seg))
        return 1;
    if (b2g_dmalloc(ctx, (void **)&tp->d_wbuf, std::max<size_t>(tp->wbuf_doubles, 2) * sizeof(double)))
        return 1;
    tp->to_free.push_back(tp->d_wbuf);
    // phase 1 overwrites the column range of every group on every matvec; what no group covers must read as zero
    B2G_CUDA(cudaMemsetAsync(tp->d_wbuf, 0, std::max<size_t>(tp->wbuf_doubles, 2) * sizeof(double), ctx->stream));
    if (!tiles.empty() || tp->pbuf_doubles != 0 || !part_off.empty()) {
        if (upload(tiles.data(), tiles.size() * sizeof(OutTile), (void **)&tp->d_tiles))
            return 1;
        if (upload(part_off.data(), part_off.size() * sizeof(int64_t), (void **)&tp->d_part_off))
            return 1;
        if (b2g_dmalloc(ctx, (void **)&tp->d_pbuf, std::max<size_t>(tp->pbuf_doubles, 2) * sizeof(double)))
            return 1;
        tp->to_free.push_back(tp->d_pbuf);
    }
    for (size_t gi = 0; gi < tp->groups.size(); gi++) // keep[gi] belongs to groups[gi] (not yet reordered)
        if (upload(keep[gi].data(), keep[gi].size() * sizeof(Unit), (void **)&tp->groups[gi].d_units))
            return 1;
    std::stable_sort(tp->groups.begin(), tp->groups.end(), [](const LaunchGroup &a, const LaunchGroup &b) {
        return a.slab != b.slab ? a.slab < b.slab : a.phase < b.phase;
    });
    if (b2g_dmalloc(ctx, (void **)&tp->d_args, sizeof(MatvecArgs)))
        return 1;
    tp->to_free.push_back(tp->d_args);
    tp->n_counters = std::max<int>(64, (int)tp->groups.size());
    if (b2g_dmalloc(ctx, (void **)&tp->d_counters, sizeof(unsigned int) * tp->n_counters))
        return 1;
    tp->to_free.push_back(tp->d_counters);
    lap("tiles+upload");
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    lap("sync");
    p->stats.launches = (int64_t)tp->groups.size() + (int64_t)tp->tile_ranges.size();
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
    // ---- CUDA graph replay of small lists (B2G_NO_GRAPH: off)
    static const bool graphs_on = getenv("B2G_NO_GRAPH") == nullptr;
    // below 50 GFLOP per matvec (N2, H10, C2, chain ends): there the launches, not the kernels, set the pace
    const bool small_list = 2.0 * (double)p->stats.nflop_mnk < 5e10 && tp->slab_doubles.size() <= 1;
    const MatvecArgs *ind = nullptr;
    bool capturing = false;
    if (graphs_on && small_list && stats == nullptr && tp->graph_state >= 0) {
        if (tp->graph_state == 2) {
            set_args_kernel<<<1, 1, 0, ctx->stream>>>(tp->d_args, c_dev, v_dev, scale);
            B2G_CUDA(cudaGraphLaunch(tp->graph_exec, ctx->stream));
            ctx->launches += tp->graph_kernels + 1;
            return 0;
        }
        if (tp->graph_state == 1) {
            set_args_kernel<<<1, 1, 0, ctx->stream>>>(tp->d_args, c_dev, v_dev, scale);
            ctx->launches++;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess)
                capturing = true, ind = tp->d_args;
            else
                cudaGetLastError(), tp->graph_state = -1;
        } else
            tp->graph_state = 1; // this call runs eagerly
    }
    const int64_t launches_before = ctx->launches;
    // 0: the graph was instantiated and launched; 1: capture failed - the caller runs the launches eagerly
    auto finish_capture = [&]() -> int {
        capturing = false;
        cudaGraph_t graph = nullptr;
        if (cudaStreamEndCapture(ctx->stream, &graph) != cudaSuccess || graph == nullptr) {
            cudaGetLastError();
            if (graph)
                cudaGraphDestroy(graph);
            tp->graph_state = -1;
            return 1;
        }
        const cudaError_t e = cudaGraphInstantiate(&tp->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            tp->graph_exec = nullptr, tp->graph_state = -1;
            return 1;
        }
        tp->graph_kernels = (int)(ctx->launches - launches_before);
        if (cudaGraphLaunch(tp->graph_exec, ctx->stream) != cudaSuccess) { // the captured work has not run yet
            cudaGetLastError();
            cudaGraphExecDestroy(tp->graph_exec);
            tp->graph_exec = nullptr, tp->graph_state = -1;
            return 1;
        }
        tp->graph_state = 2;
        return 0;
    };
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
    auto body = [&]() -> int { // every launch of one matvec (eager, or recorded into the graph being captured)
    B2G_CUDA(cudaMemsetAsync(tp->d_counters, 0, sizeof(unsigned int) * tp->n_counters, ctx->stream));
    // Without profiling, the launches of one phase are forked onto side streams and joined before
    // the next phase; with profiling everything stays on the context stream (timed one by one).
    // Forking pays when no single launch fills the chip (measured: up to ~2x on the C2 / H10 lists,
    // -5 % on the 2 TFLOP Cr2 list where the big persistent kernels then compete), hence the bound.
    const bool multi_slab = tp->slab_doubles.size() > 1;
    const bool fork_all = stats == nullptr && 2.0 * (double)p->stats.nflop_mnk < 3e11 && !multi_slab;
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
    int gi = 0, cur_slab = -1;
    bool phase2_open = false;
    // Measured on the Cr2 M=4000 list: 81.8 - 82.7 ms with the chain against 73.7 - 74.1 ms without (also with the
    // first unit claimed dynamically), so it is an experiment switch (B2G_PDL=1), off by default.
    static const bool use_pdl = getenv("B2G_PDL") != nullptr;
    bool chain_open = false, first_p2 = true;
    for (const LaunchGroup &g : tp->groups) {
        if (g.slab != cur_slab) { // next slab: same workspace; what its groups do not cover must read as zero again
            cur_slab = g.slab, phase2_open = false, first_p2 = true, chain_open = false;
            if (multi_slab)
                B2G_CUDA(cudaMemsetAsync(tp->d_wbuf, 0, tp->slab_doubles[g.slab] * sizeof(double), ctx->stream));
        }
        if (g.phase == 2 && !phase2_open) { // every W panel is complete before the first phase-2 launch
            phase2_open = true;
            if (fork) {
                if (join())
                    return 1;
                B2G_CUDA(cudaEventRecord(ctx->fork_ev, ctx->stream));
            }
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
        // chained launches on the context stream (no events, no side streams in between) overlap their tails
        const bool chain = use_pdl && !fork && stats == nullptr;
        const bool pdl = chain && chain_open;
        const int wait_first = (g.phase == 2 && first_p2) ? 1 : 0;
        if (g.phase == 2)
            first_p2 = false;
        chain_open = chain;
        char nm[64];
        snprintf(nm, sizeof(nm), "phase%d_%dx%d_%s", g.phase, kCfg[g.cfg].bm, kCfg[g.cfg].bn,
                 g.phase == 1 ? (g.layout ? "Bt" : "Bn") : (g.layout ? "At" : "An"));
        if (begin(nm, g.flops, g.n_units))
            return 1;
#define B2G_DISPATCH(PH, CFG, LAY, CALL)                                                \
    if (g.phase == PH && g.cfg == CFG && g.layout == LAY)                               \
        rc = CALL;
        B2G_DISPATCH(1, 0, 0, (launch_p1<Cfg0, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 0, 1, (launch_p1<Cfg0, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 1, 0, (launch_p1<Cfg1, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 1, 1, (launch_p1<Cfg1, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 2, 0, (launch_p1<Cfg2, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 2, 1, (launch_p1<Cfg2, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 3, 0, (launch_p1<Cfg3, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 3, 1, (launch_p1<Cfg3, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 4, 0, (launch_p1<Cfg4, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 4, 1, (launch_p1<Cfg4, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 5, 0, (launch_p1<Cfg5, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 5, 1, (launch_p1<Cfg5, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 6, 0, (launch_p1<Cfg6, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 6, 1, (launch_p1<Cfg6, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 7, 0, (launch_p1<Cfg7, false>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(1, 7, 1, (launch_p1<Cfg7, true>(g, *tp, ctx, gs, counter, c_dev, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 0, 0, (launch_p2<Cfg0, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 0, 1, (launch_p2<Cfg0, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 1, 0, (launch_p2<Cfg1, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 1, 1, (launch_p2<Cfg1, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 2, 0, (launch_p2<Cfg2, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 2, 1, (launch_p2<Cfg2, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 3, 0, (launch_p2<Cfg3, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 3, 1, (launch_p2<Cfg3, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 4, 0, (launch_p2<Cfg4, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 4, 1, (launch_p2<Cfg4, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 5, 0, (launch_p2<Cfg5, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 5, 1, (launch_p2<Cfg5, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 6, 0, (launch_p2<Cfg6, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 6, 1, (launch_p2<Cfg6, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 7, 0, (launch_p2<Cfg7, true>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
        B2G_DISPATCH(2, 7, 1, (launch_p2<Cfg7, false>(g, *tp, ctx, gs, counter, v_dev, scale, pdl, wait_first, ind)))
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
                tp->d_tiles + rg.first, nt, tp->d_part_off, tp->d_win, tp->d_pbuf, v_dev, scale, ind);
            ctx->launches++;
        }
        if (end())
            return 1;
    }
    return 0;
    };
    int body_rc = body();
    if (capturing) {
        const int64_t recorded = ctx->launches - launches_before;
        bool replayed = false;
        if (body_rc == 0)
            replayed = finish_capture() == 0;
        else { // a launch failed while recording: drop the capture
            cudaGraph_t graph = nullptr;
            cudaStreamEndCapture(ctx->stream, &graph);
            if (graph)
                cudaGraphDestroy(graph);
            cudaGetLastError();
            capturing = false, tp->graph_state = -1;
        }
        if (!replayed) { // no graph: this matvec runs eagerly, and so do the following ones
            ctx->launches -= recorded;
            ind = nullptr;
            body_rc = body();
        }
    }
    if (body_rc)
        return body_rc;
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
