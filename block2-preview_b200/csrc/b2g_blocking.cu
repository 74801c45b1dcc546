// b2g_blocking.cu — executor of the blocking lists (sm_100a).
//
// What is executed: the single-batch GEMM list TensorFunctions::left_contract / right_contract
// record in SeqTypes::Auto through OperatorFunctions::tensor_product and
// AdvancedGEMM<double>::tensor_product (block2 src/core/tensor_functions.hpp:2842-2885, 2941-2984,
// src/core/operator_functions.hpp:672-711, src/core/batch_gemm.hpp:433-503): every term
// a[x] (x) b[y] of a blocked operator becomes k = 1 "rows-as-AXPY" GEMM groups that all accumulate
// into the same output block.  The reference resolves those write conflicts with work arrays and a
// post-batch reduction (BatchGEMMSeq::prepare / auto_perform, batch_gemm.hpp:1222-1530); here the
// list is regrouped by OUTPUT window instead: each window element is owned by one thread, which
// walks the contributions of that window in list order with the running value in a register and
// writes it once.  No atomics, no work arrays, bit-reproducible, and each source element is read
// exactly once - the kernel is HBM-bound (algorithmic bytes = 8 * (sources + outputs)).
//
// Entries are held in an element-wise canonical form that covers the GEMM (any k), the AXPY rows
// and the 2-D windows the rows of one group fold into:
//     C(i, j) = beta * C(i, j) + alpha * sum_kk A[i*sa_i + j*sa_j + kk*sa_k] * B[i*sb_i + j*sb_j + kk*sb_k]
// Nothing of the reference is compiled into this file.
#include "b2g_internal.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace {

struct BlkEntry { // 64 bytes
    const double *a, *b;
    double alpha, beta;
    int32_t sa_i, sa_j, sa_k, sb_i, sb_j, sb_k;
    int32_t k;
    int32_t nd; // != 0: the window is addressed flat (one contiguous vector) and this entry's own
                // logical width is nd: i = e / nd, j = e % nd
};
static_assert(sizeof(BlkEntry) == 64, "BlkEntry layout");

struct BlkUnit { // one warp: `len` consecutive logical elements of one output window
    double *dst;
    int32_t e0, len, n, ldc, first, count;
};
static_assert(sizeof(BlkUnit) == 32, "BlkUnit layout");

struct BlkSerial { // entry of an irregularly overlapping component, with its own window
    BlkEntry e;
    double *dst;
    int32_t m, n, ldc, comp;
};

constexpr int UNIT_ELEMS = 2048; // elements per warp unit
constexpr int PER_LANE = 4;      // elements per lane and sub-chunk
constexpr int BLK_THREADS = 256;

__device__ __forceinline__ double blk_apply(const BlkEntry &E, double acc, int64_t i, int64_t j) {
    if (E.nd) {
        const int64_t e = i;
        i = e / E.nd, j = e - i * E.nd;
    }
    const double base = E.beta == 1.0 ? acc : (E.beta == 0.0 ? 0.0 : E.beta * acc);
    if (E.alpha == 0.0 || E.k == 0) // BLAS: A and B are not referenced
        return base;
    const double *ap = E.a + i * E.sa_i + j * E.sa_j;
    const double *bp = E.b + i * E.sb_i + j * E.sb_j;
    double s;
    if (E.k == 1)
        s = __ldg(ap) * __ldg(bp);
    else {
        s = 0.0;
        for (int kk = 0; kk < E.k; kk++)
            s = fma(__ldg(ap + (int64_t)kk * E.sa_k), __ldg(bp + (int64_t)kk * E.sb_k), s);
    }
    return fma(E.alpha, s, base);
}

// AXPY windows (every contribution is k = 1 with one scalar B): one warp per unit, persistent
// grid-stride over units.  A unit is up to UNIT_ELEMS consecutive logical elements of one output window;
// the warp walks it in sub-chunks of 128 elements (4 per lane, running values in registers across the
// contributions), so the unit / entry descriptors are fetched once per UNIT_ELEMS elements and the
// loads of a sub-chunk are independent.  Index arithmetic is 32-bit, and the division by the window
// width is only done for windows / sources that are not linear in the element index.
// dst_zero: outputs start from 0 (not read).
__device__ __forceinline__ int64_t blk_src_index(const BlkEntry &E, int32_t e, int32_t n) {
    // e: logical element of the window; n: window width (1 = flat)
    if (E.nd) { // flat window, this entry's own width
        const uint32_t i = (uint32_t)e / (uint32_t)E.nd, j = (uint32_t)e - i * (uint32_t)E.nd;
        return (int64_t)i * E.sa_i + (int64_t)j * E.sa_j;
    }
    if (n == 1)
        return (int64_t)e * E.sa_i;
    const uint32_t i = (uint32_t)e / (uint32_t)n, j = (uint32_t)e - i * (uint32_t)n;
    return (int64_t)i * E.sa_i + (int64_t)j * E.sa_j;
}

constexpr int ROW_MIN = 128;   // windows at least this wide are cut into row-aligned units
constexpr int STREAM_PER = 8;  // elements per lane and step
constexpr int RING_STAGES = 4; // cp.async ring depth of the streaming kernel
constexpr size_t RING_BYTES = (size_t)(BLK_THREADS / 32) * RING_STAGES * STREAM_PER * 32 * sizeof(double);
constexpr int ACC_STAGES_DEFAULT = 3;     // ring depth of the accumulating kernel
constexpr int ACC_SLOT = STREAM_PER + 3;  // per lane and stage: 8 source values, b, alpha, beta
constexpr size_t acc_ring_bytes(int stages) { return (size_t)(BLK_THREADS / 32) * stages * ACC_SLOT * 32 * sizeof(double); }

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}

// Accumulating kernel: windows with several contributions, narrow 2-D windows, transposed sources.
// One warp per unit.  The unit is walked as a flat sequence of steps (sub-chunk of 256 elements x
// contribution); the source values of step q + ACC_STAGES - 1 are copied into the warp's cp.async ring
// while step q is folded into the running values (8 per lane, in registers from the first to the last
// contribution of a sub-chunk), so the loads of different contributions overlap.  The descriptor of the
// next contribution is fetched one step ahead.  dst_zero: outputs start from 0 (not read).
template <int ACC_STAGES>
__global__ void __launch_bounds__(BLK_THREADS, 2)
b2g_blocking_kernel(const BlkUnit *__restrict__ units, int64_t nunits, const BlkEntry *__restrict__ entries,
                    int dst_zero) {
    extern __shared__ double ring[];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * BLK_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * BLK_THREADS) >> 5;
    double *my = ring + ((size_t)(threadIdx.x >> 5) * ACC_STAGES * ACC_SLOT) * 32 + lane;
    for (int64_t u = warp; u < nunits; u += nwarps) {
        const BlkUnit U = units[u];
        const bool rowu = U.n >= ROW_MIN; // unit lies inside one row of a wide window
        const bool lin = rowu || U.n == 1; // destination (and untransposed sources) linear in the position
        uint32_t i0 = 0, j0 = 0;
        if (rowu)
            i0 = (uint32_t)U.e0 / (uint32_t)U.n, j0 = (uint32_t)U.e0 - i0 * (uint32_t)U.n;
        double *__restrict__ dptr = U.dst + (rowu ? (int64_t)i0 * U.ldc + j0 : (lin ? (int64_t)U.e0 * U.ldc : 0));
        const int64_t dstep = rowu ? 1 : U.ldc;
        auto dst_index = [&](int l) -> int64_t {
            if (lin)
                return l * dstep;
            const uint32_t e = (uint32_t)(U.e0 + l), i = e / (uint32_t)U.n, j = e - i * (uint32_t)U.n;
            return (int64_t)i * U.ldc + j;
        };
        const int nchunks = (U.len + 32 * STREAM_PER - 1) / (32 * STREAM_PER);
        const int steps = nchunks * U.count;
        BlkEntry Ep = entries[U.first]; // descriptor of the next step to issue
        int qi = 0, ci = 0, ti = 0;     // issue cursor: step, sub-chunk, contribution
        auto issue = [&]() {
            if (qi < steps) {
                const BlkEntry E = Ep;
                const int tn = ti + 1 == U.count ? 0 : ti + 1;
                if (qi + 1 < steps && U.count > 1)
                    Ep = entries[U.first + tn]; // in flight until the next issue
                double *slot = my + (size_t)(qi % ACC_STAGES) * ACC_SLOT * 32;
                slot[STREAM_PER * 32 + 32] = E.alpha, slot[STREAM_PER * 32 + 64] = E.beta;
                if (E.alpha != 0.0) {
                    cp_async8(slot + STREAM_PER * 32, E.b);
                    const bool slin = lin && E.nd == 0;
                    const int64_t sbase = rowu ? (int64_t)i0 * E.sa_i + (int64_t)j0 * E.sa_j : (int64_t)U.e0 * E.sa_i;
                    const int64_t sstep = rowu ? E.sa_j : E.sa_i;
#pragma unroll
                    for (int r = 0; r < STREAM_PER; r++) {
                        const int l = ci * 32 * STREAM_PER + r * 32 + lane;
                        if (l < U.len)
                            cp_async8(slot + r * 32, E.a + (slin ? sbase + l * sstep : blk_src_index(E, U.e0 + l, U.n)));
                    }
                }
                qi++, ti = tn, ci += tn == 0 ? 1 : 0;
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
#pragma unroll
        for (int p = 0; p < ACC_STAGES - 1; p++)
            issue();
        double acc[STREAM_PER];
        int tc = 0, cc = 0;
        for (int q = 0; q < steps; q++) {
            issue();
            asm volatile("cp.async.wait_group %0;\n" ::"n"(ACC_STAGES - 1) : "memory");
            const double *slot = my + (size_t)(q % ACC_STAGES) * ACC_SLOT * 32;
            if (tc == 0) {
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++) {
                    const int l = cc * 32 * STREAM_PER + r * 32 + lane;
                    acc[r] = (!dst_zero && l < U.len) ? dptr[dst_index(l)] : 0.0;
                }
            }
            const double alpha = slot[STREAM_PER * 32 + 32], beta = slot[STREAM_PER * 32 + 64];
            if (alpha != 0.0) {
                const double f = alpha * slot[STREAM_PER * 32];
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++) {
                    const int l = cc * 32 * STREAM_PER + r * 32 + lane;
                    if (l < U.len)
                        acc[r] = fma(f, slot[r * 32], beta == 1.0 ? acc[r] : (beta == 0.0 ? 0.0 : beta * acc[r]));
                }
            } else {
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++)
                    acc[r] = beta == 1.0 ? acc[r] : (beta == 0.0 ? 0.0 : beta * acc[r]);
            }
            if (++tc == U.count) {
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++) {
                    const int l = cc * 32 * STREAM_PER + r * 32 + lane;
                    if (l < U.len)
                        dptr[dst_index(l)] = acc[r];
                }
                tc = 0, cc++;
            }
        }
    }
}

// Streaming units: windows (or window rows) with ONE contribution whose addresses are linear in the
// position inside the unit - 80-90 % of the bytes of a blocking step.  The descriptor is self-contained
// (pre-offset pointers), the next unit's descriptor and scalar are fetched while the current unit
// streams, and the source goes through the per-thread cp.async ring.
struct StreamUnit { // 64 bytes: `rows` rows of `len` elements each (row pitches drow / srow)
    double *dst;
    const double *src;
    const double *b;
    double alpha;
    int32_t len, dstep, sstep, rows;
    int32_t drow, srow;
    int64_t pad2;
};
static_assert(sizeof(StreamUnit) == 64, "StreamUnit layout");

// One CTA per unit.  A unit (rows x len) is cut into pieces of up to PIECE_CHUNKS chunks of one row (a chunk =
// 32 lanes x STREAM_PER elements); the warps of the CTA take the pieces in turn (piece w, w + 8, ...), so the
// pieces in flight on the chip are neighbours in source and destination - as with one descriptor per row or
// per 2048 elements - at a fraction of the descriptors, and the cp.async ring of a warp keeps running from one
// piece into the next.
constexpr int PIECE_CHUNKS = 8;
__global__ void __launch_bounds__(BLK_THREADS, 3)
b2g_blocking_stream_kernel(const StreamUnit *__restrict__ units, int64_t nunits) {
    extern __shared__ double ring[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int NW = BLK_THREADS / 32;
    if (blockIdx.x >= nunits)
        return;
    double *my = ring + ((size_t)wib * RING_STAGES * STREAM_PER) * 32 + lane;
    StreamUnit U = units[blockIdx.x];
    double bval = __ldg(U.b);
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        const bool more = u + gridDim.x < nunits;
        StreamUnit N = U;
        if (more)
            N = units[u + gridDim.x]; // in flight while this unit streams
        const double f = U.alpha * bval;
        const int64_t sstep = U.sstep, dstep = U.dstep;
        const int len = U.len;
        const int ncr = (len + 32 * STREAM_PER - 1) / (32 * STREAM_PER); // chunks per row
        const int ngr = (ncr + PIECE_CHUNKS - 1) / PIECE_CHUNKS;         // pieces per row
        const int npieces = U.rows * ngr;
        // the warp that takes piece 0 rotates with the unit: the first (npieces mod 8) warps get one piece more,
        // and with a fixed start they would be the same warps in every unit
        const int w0 = (wib + NW - (int)(u & (NW - 1))) & (NW - 1);
        // issue side: piece ip, chunks [ic, ic_end) of its row, row base isrc
        int ip = w0, ic = 0, ic_end = 0, qi = 0;
        const double *__restrict__ isrc = U.src;
        auto open_issue = [&]() {
            if (ip < npieces) {
                const int r = ip / ngr, g = ip - r * ngr;
                isrc = U.src + (int64_t)r * U.srow;
                ic = g * PIECE_CHUNKS, ic_end = min(ncr, ic + PIECE_CHUNKS);
            }
        };
        open_issue();
        auto issue = [&]() {
            if (ip < npieces) {
                double *slot = my + (size_t)(qi % RING_STAGES) * STREAM_PER * 32;
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++) {
                    const int l = ic * 32 * STREAM_PER + r * 32 + lane;
                    if (l < len)
                        cp_async8(slot + r * 32, isrc + l * sstep);
                }
                qi++;
                if (++ic == ic_end)
                    ip += NW, open_issue();
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
#pragma unroll
        for (int c = 0; c < RING_STAGES - 1; c++)
            issue();
        // write side: the same walk, RING_STAGES - 1 steps behind
        int q = 0;
        for (int wp = w0; wp < npieces; wp += NW) {
            const int r = wp / ngr, g = wp - r * ngr;
            double *__restrict__ dptr = U.dst + (int64_t)r * U.drow;
            const int wc_end = min(ncr, (g + 1) * PIECE_CHUNKS);
            for (int wc = g * PIECE_CHUNKS; wc < wc_end; wc++, q++) {
                issue();
                asm volatile("cp.async.wait_group %0;\n" ::"n"(RING_STAGES - 1) : "memory");
                const double *slot = my + (size_t)(q % RING_STAGES) * STREAM_PER * 32;
#pragma unroll
                for (int rr = 0; rr < STREAM_PER; rr++) {
                    const int l = wc * 32 * STREAM_PER + rr * 32 + lane;
                    if (l < len)
                        dptr[l * dstep] = f * slot[rr * 32];
                }
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        U = N;
        if (more)
            bval = __ldg(U.b); // N arrived long ago; one dependent load per unit
    }
}

// Multi-source streaming units: 2..MULTI_MAX contributions, all linear - e.g. the two-term sums that
// make up half of the bytes of an H_eff blocking step.  Like the streaming kernel (self-contained
// descriptor, per-thread cp.async ring with RING_STAGES - 1 steps in flight), a step being one
// (256-element sub-chunk, contribution) pair; the running values stay in registers across the
// contributions of a sub-chunk, in list order.
constexpr int MULTI_MAX = 4;
struct MultiUnit { // 160 bytes
    double *dst;
    const double *src[MULTI_MAX];
    const double *b[MULTI_MAX];
    double alpha[MULTI_MAX];
    double beta[MULTI_MAX];
    int32_t sstep[MULTI_MAX];
    int32_t srow[MULTI_MAX]; // row pitch of every source
    int32_t len, dstep, count, rows; // `rows` rows of `len` elements each
    int32_t drow, pad;               // row pitch of the window
};
static_assert(sizeof(MultiUnit) == 8 + 4 * 8 * MULTI_MAX + 8 * MULTI_MAX + 24, "MultiUnit layout");

__global__ void __launch_bounds__(BLK_THREADS, 3)
b2g_blocking_multi_kernel(const MultiUnit *__restrict__ units, int64_t nunits, int dst_zero) {
    extern __shared__ double ring[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int NW = BLK_THREADS / 32;
    double *my = ring + ((size_t)wib * RING_STAGES * STREAM_PER) * 32 + lane;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) { // one CTA per unit, warps take its pieces in turn
        const MultiUnit &G = units[u];
        const int count = G.count, len = G.len;
        const int64_t dstep = G.dstep;
        const double *s0[MULTI_MAX]; // unit origin of every source
        const double *sp[MULTI_MAX]; // row of the piece being issued
        double f[MULTI_MAX], bt[MULTI_MAX];
        int64_t ss[MULTI_MAX], sr[MULTI_MAX];
#pragma unroll
        for (int t = 0; t < MULTI_MAX; t++) {
            const bool on = t < count;
            s0[t] = on ? G.src[t] : nullptr, sp[t] = s0[t], ss[t] = on ? G.sstep[t] : 0, sr[t] = on ? G.srow[t] : 0;
            f[t] = on ? G.alpha[t] * __ldg(G.b[t]) : 0.0, bt[t] = on ? G.beta[t] : 1.0;
        }
        const int ncr = (len + 32 * STREAM_PER - 1) / (32 * STREAM_PER); // chunks per row
        const int ngr = (ncr + PIECE_CHUNKS - 1) / PIECE_CHUNKS;         // pieces per row
        const int npieces = G.rows * ngr;
        const int w0 = (wib + NW - (int)(u & (NW - 1))) & (NW - 1); // rotates with the unit (see the stream kernel)
        // issue side: piece ip, chunk ci of [.., ci_end) of its row, contribution ti
        int ip = w0, ci = 0, ci_end = 0, ti = 0, qi = 0;
        auto open_issue = [&]() {
            if (ip < npieces) {
                const int r = ip / ngr, g = ip - r * ngr;
#pragma unroll
                for (int t = 0; t < MULTI_MAX; t++)
                    sp[t] = s0[t] + (int64_t)r * sr[t];
                ci = g * PIECE_CHUNKS, ci_end = min(ncr, ci + PIECE_CHUNKS);
            }
        };
        open_issue();
        auto issue = [&]() {
            if (ip < npieces) {
                double *slot = my + (size_t)(qi % RING_STAGES) * STREAM_PER * 32;
                const double *sx = ti == 0 ? sp[0] : (ti == 1 ? sp[1] : (ti == 2 ? sp[2] : sp[3]));
                const int64_t st = ti == 0 ? ss[0] : (ti == 1 ? ss[1] : (ti == 2 ? ss[2] : ss[3]));
#pragma unroll
                for (int r = 0; r < STREAM_PER; r++) {
                    const int l = ci * 32 * STREAM_PER + r * 32 + lane;
                    if (l < len)
                        cp_async8(slot + r * 32, sx + l * st);
                }
                qi++;
                if (++ti == count) {
                    ti = 0;
                    if (++ci == ci_end)
                        ip += NW, open_issue();
                }
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
#pragma unroll
        for (int p = 0; p < RING_STAGES - 1; p++)
            issue();
        double acc[STREAM_PER];
        int q = 0;
        for (int wp = w0; wp < npieces; wp += NW) { // write side: the same walk, RING_STAGES - 1 steps behind
            const int r = wp / ngr, g = wp - r * ngr;
            double *__restrict__ dptr = G.dst + (int64_t)r * G.drow;
            const int cc_end = min(ncr, (g + 1) * PIECE_CHUNKS);
            for (int cc = g * PIECE_CHUNKS; cc < cc_end; cc++)
                for (int tc = 0; tc < count; tc++, q++) {
                    issue();
                    asm volatile("cp.async.wait_group %0;\n" ::"n"(RING_STAGES - 1) : "memory");
                    const double *slot = my + (size_t)(q % RING_STAGES) * STREAM_PER * 32;
                    if (tc == 0) {
#pragma unroll
                        for (int rr = 0; rr < STREAM_PER; rr++) {
                            const int l = cc * 32 * STREAM_PER + rr * 32 + lane;
                            acc[rr] = (!dst_zero && l < len) ? dptr[l * dstep] : 0.0;
                        }
                    }
                    const double ft = tc == 0 ? f[0] : (tc == 1 ? f[1] : (tc == 2 ? f[2] : f[3]));
                    const double be = tc == 0 ? bt[0] : (tc == 1 ? bt[1] : (tc == 2 ? bt[2] : bt[3]));
#pragma unroll
                    for (int rr = 0; rr < STREAM_PER; rr++) {
                        const int l = cc * 32 * STREAM_PER + rr * 32 + lane;
                        if (l < len)
                            acc[rr] = fma(ft, slot[rr * 32], be == 1.0 ? acc[rr] : (be == 0.0 ? 0.0 : be * acc[rr]));
                    }
                    if (tc == count - 1) {
#pragma unroll
                        for (int rr = 0; rr < STREAM_PER; rr++) {
                            const int l = cc * 32 * STREAM_PER + rr * 32 + lane;
                            if (l < len)
                                dptr[l * dstep] = acc[rr];
                        }
                    }
                }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
}

// Tile units: 2-D windows with several contributions of which some are TRANSPOSED sources
// (C(i, j) += f * A(j, i): contiguous along i).  Row-shaped units would read such a source with a stride
// of one pitch per lane; here one warp owns a TILE_R x TILE_C tile of the window, every contribution is
// copied into a warp-shared padded tile in shared memory along ITS OWN contiguous direction (cp.async
// ring, one contribution per step) and folded into the running values (TILE_R per lane, registers) in
// list order.  The tile is written once, row by row.
constexpr int TILE_R = 16, TILE_C = 32, TILE_LD = TILE_C + 1, TILE_STAGES = 3;
constexpr int TILE_STRIP = 32; // column tiles per unit (a CTA of 8 warps works on one unit)
constexpr int TILE_SLOT = TILE_R * TILE_LD + 4; // tile + (b, alpha, beta, pad) of the step
constexpr size_t TILE_RING_BYTES = (size_t)(BLK_THREADS / 32) * TILE_STAGES * TILE_SLOT * sizeof(double);
struct TileUnit { // 40 bytes: `ntiles` tiles, row-major over the tile rows of the window that start at row i0
    double *dst;  // window origin
    int32_t i0, j0, m, n, ldc, first, count, ntiles;
};
static_assert(sizeof(TileUnit) == 40, "TileUnit layout");

__global__ void __launch_bounds__(BLK_THREADS, 2)
b2g_blocking_tile_kernel(const TileUnit *__restrict__ units, int64_t nunits, const BlkEntry *__restrict__ entries,
                         int dst_zero) {
    extern __shared__ double ring[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int NW = BLK_THREADS / 32;
    double *my = ring + (size_t)wib * TILE_STAGES * TILE_SLOT;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) // one CTA per unit, warps take its tiles in turn
      for (int tt = (wib + NW - (int)(u & (NW - 1))) & (NW - 1), ntt = units[u].ntiles; tt < ntt; tt += NW) {
        TileUnit U = units[u];
        const int tpr = (U.n + TILE_C - 1) / TILE_C, trow = tt / tpr; // tiles per tile row
        U.i0 += trow * TILE_R, U.j0 = (tt - trow * tpr) * TILE_C;
        const int rows = min(TILE_R, U.m - U.i0), cols = min(TILE_C, U.n - U.j0);
        double *__restrict__ dptr = U.dst + (int64_t)U.i0 * U.ldc + U.j0;
        BlkEntry Ep = entries[U.first];
        int qi = 0;
        auto issue = [&]() {
            if (qi < U.count) {
                const BlkEntry E = Ep;
                if (qi + 1 < U.count)
                    Ep = entries[U.first + qi + 1]; // in flight until the next issue
                double *slot = my + (size_t)(qi % TILE_STAGES) * TILE_SLOT;
                if (lane == 0) {
                    slot[TILE_R * TILE_LD + 1] = E.alpha, slot[TILE_R * TILE_LD + 2] = E.beta;
                    if (E.alpha != 0.0)
                        cp_async8(slot + TILE_R * TILE_LD, E.b);
                }
                if (E.alpha != 0.0) {
                    const double *base = E.a + (int64_t)U.i0 * E.sa_i + (int64_t)U.j0 * E.sa_j;
                    if (E.sa_i == 1 && E.sa_j != 1) { // transposed source: consecutive i are contiguous
                        const int i = lane & (TILE_R - 1), jh = lane / TILE_R;
#pragma unroll
                        for (int r = 0; r < TILE_C / (32 / TILE_R); r++) {
                            const int j = r * (32 / TILE_R) + jh;
                            if (i < rows && j < cols)
                                cp_async8(slot + i * TILE_LD + j, base + i + (int64_t)j * E.sa_j);
                        }
                    } else { // consecutive j are contiguous (or a general stride)
#pragma unroll
                        for (int r = 0; r < TILE_R; r++)
                            if (r < rows && lane < cols)
                                cp_async8(slot + r * TILE_LD + lane, base + (int64_t)r * E.sa_i + (int64_t)lane * E.sa_j);
                    }
                }
                qi++;
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
#pragma unroll
        for (int p = 0; p < TILE_STAGES - 1; p++)
            issue();
        double acc[TILE_R];
#pragma unroll
        for (int r = 0; r < TILE_R; r++)
            acc[r] = (!dst_zero && r < rows && lane < cols) ? dptr[(int64_t)r * U.ldc + lane] : 0.0;
        for (int q = 0; q < U.count; q++) {
            issue();
            asm volatile("cp.async.wait_group %0;\n" ::"n"(TILE_STAGES - 1) : "memory");
            __syncwarp(); // the copies of every lane of the warp have landed
            const double *slot = my + (size_t)(q % TILE_STAGES) * TILE_SLOT;
            const double alpha = slot[TILE_R * TILE_LD + 1], beta = slot[TILE_R * TILE_LD + 2];
            if (alpha != 0.0) {
                const double f = alpha * slot[TILE_R * TILE_LD];
#pragma unroll
                for (int r = 0; r < TILE_R; r++)
                    if (r < rows && lane < cols)
                        acc[r] = fma(f, slot[r * TILE_LD + lane], beta == 1.0 ? acc[r] : (beta == 0.0 ? 0.0 : beta * acc[r]));
            } else {
#pragma unroll
                for (int r = 0; r < TILE_R; r++)
                    acc[r] = beta == 1.0 ? acc[r] : (beta == 0.0 ? 0.0 : beta * acc[r]);
            }
            __syncwarp(); // all lanes are done with this stage before a later step overwrites it
        }
#pragma unroll
        for (int r = 0; r < TILE_R; r++)
            if (r < rows && lane < cols)
                dptr[(int64_t)r * U.ldc + lane] = acc[r];
    }
}

// Windows with a general contribution (k > 1, or B varying over the window): one thread per element.
__global__ void __launch_bounds__(BLK_THREADS)
b2g_blocking_general_kernel(const BlkUnit *__restrict__ units, int64_t nunits, const BlkEntry *__restrict__ entries,
                            int dst_zero) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * BLK_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * BLK_THREADS) >> 5;
    for (int64_t u = warp; u < nunits; u += nwarps) {
        const BlkUnit U = units[u];
        for (int l = lane; l < U.len; l += 32) {
            const int64_t e = U.e0 + l;
            const int64_t i = e / U.n, j = e - i * U.n;
            double *p = U.dst + i * U.ldc + j;
            double acc = dst_zero ? 0.0 : *p;
            for (int t = 0; t < U.count; t++)
                acc = blk_apply(entries[U.first + t], acc, i, j);
            *p = acc;
        }
    }
}

// Irregular components (windows that overlap without being identical): one CTA per component, the
// entries strictly one after the other in list order.
__global__ void __launch_bounds__(BLK_THREADS)
b2g_blocking_serial_kernel(const BlkSerial *__restrict__ se, const int64_t *__restrict__ comp_first, int ncomp) {
    for (int c = blockIdx.x; c < ncomp; c += gridDim.x) {
        for (int64_t t = comp_first[c]; t < comp_first[c + 1]; t++) {
            const BlkSerial S = se[t];
            const int64_t total = (int64_t)S.m * S.n;
            for (int64_t e = threadIdx.x; e < total; e += BLK_THREADS) {
                const int64_t i = e / S.n, j = e - i * S.n;
                double *p = S.dst + i * S.ldc + j;
                *p = blk_apply(S.e, *p, i, j);
            }
            __threadfence_block();
            __syncthreads();
        }
    }
}

struct HostEntry {
    BlkEntry e;
    double *dst;
    int32_t m, n, ldc;
    int64_t order; // position in the recorded list
};

inline bool is_t(int32_t t) { return t == B2G_TRANS || t == 1; }
inline bool ok_t(int32_t t) { return t == B2G_TRANS || t == B2G_NOTRANS || t == 0 || t == 1; }

inline size_t window_extent(int32_t m, int32_t n, int32_t ldc) {
    return (size_t)(m - 1) * (size_t)ldc + (size_t)n;
}

// element-overlap test of two row-major windows: exact for equal pitches, conservative otherwise
struct Win {
    uintptr_t lo;
    int64_t rows, cols, ld;
};
inline Win win_of(const HostEntry &h) {
    if (h.n == 1 && h.ldc == 1) // contiguous column vector == one row
        return Win{(uintptr_t)h.dst, 1, h.m, h.m};
    return Win{(uintptr_t)h.dst, h.m, h.n, h.ldc};
}
bool windows_overlap(const HostEntry &hx, const HostEntry &hy) {
    Win x = win_of(hx), y = win_of(hy);
    const uintptr_t xh = x.lo + ((x.rows - 1) * x.ld + x.cols) * 8, yh = y.lo + ((y.rows - 1) * y.ld + y.cols) * 8;
    if (xh <= y.lo || yh <= x.lo)
        return false;
    if (x.rows == 1) // the pitch of a single row is irrelevant
        x.ld = y.ld;
    if (y.rows == 1)
        y.ld = x.ld;
    if (x.ld != y.ld)
        return true;
    const int64_t ld = x.ld;
    const Win &lo = x.lo <= y.lo ? x : y, &hi = x.lo <= y.lo ? y : x;
    if (lo.cols > ld || hi.cols > ld)
        return true;
    const int64_t d = (int64_t)((hi.lo - lo.lo) / 8);
    const int64_t r = d / ld, col = d % ld;
    if (col + hi.cols > ld)
        return true; // wraps around the pitch
    if (col >= lo.cols)
        return false; // disjoint column ranges on every shared row
    return r < lo.rows; // shared columns: overlap iff the row ranges intersect
}

struct UF {
    std::vector<int> p;
    explicit UF(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) {
        while (p[x] != x)
            x = p[x] = p[p[x]];
        return x;
    }
    void unite(int a, int b) { p[find(a)] = find(b); }
};

} // namespace

// Back end shared by the list and the term entry points: entries -> output-window clusters -> warp
// units -> kernels, with the operand mirroring of the host operand space around it.
static int execute_entries(b2g_context *ctx, std::vector<HostEntry> &he, int operand_space, int flags,
                           b2g_blocking_stats &st, b2g_blocking_stats *stats,
                           std::chrono::steady_clock::time_point t_begin, const char *who) {
    const bool dst_zero = (flags & B2G_DST_ZERO) != 0;
    st.merged = (int64_t)he.size();
    b2g_prof_record("blocking.parse", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    double prof_t = B2GProfScope::now();
    auto prof_lap = [&prof_t](const char *label) {
        if (b2g_prof_enabled()) {
            const double now = B2GProfScope::now();
            b2g_prof_record(label, now - prof_t);
            prof_t = now;
        }
    };
    if (he.empty()) {
        if (stats)
            *stats = st;
        return 0;
    }
    for (const HostEntry &h : he)
        if (h.e.alpha != 0.0 && h.e.k > 0) { // distinct source elements of the entry
            const int64_t na = (int64_t)(h.e.sa_i ? h.m : 1) * (h.e.sa_j ? h.n : 1) * h.e.k;
            const int64_t nb = (int64_t)(h.e.sb_i ? h.m : 1) * (h.e.sb_j ? h.n : 1) * h.e.k;
            st.bytes_in += 8 * (na + nb);
        }
    // host operand space: the address ranges to mirror, from the entries' own (unflattened) shapes
    std::vector<B2GRange> in_rg, out_rg;
    if (operand_space == B2G_OPERANDS_HOST) {
        auto add = [](std::vector<B2GRange> &v, const void *ptr, size_t ext) {
            if (ext)
                v.push_back(B2GRange{(uintptr_t)ptr, (uintptr_t)ptr + ext * sizeof(double), 0});
        };
        // device-resident blocks (b2g_resident_map): sources are read in place, windows of resident
        // output blocks are written in place and not copied back.  Those operands switch to device
        // addresses here (unified addressing: a device address never lies inside a host range, which is
        // what the translation of the remaining operands below relies on).
        const bool have_map = ctx && !ctx->rmap.empty();
        for (size_t z = 0; z < he.size(); z++) {
            HostEntry &h = he[z];
            const size_t e_out = window_extent(h.m, h.n, h.ldc);
            if (double *d = have_map ? b2g_map_lookup(ctx, h.dst, e_out * 8) : nullptr)
                h.dst = d;
            else
                add(out_rg, h.dst, e_out);
            if (h.e.alpha == 0.0 || h.e.k == 0)
                continue;
            const size_t e_a = (size_t)(h.m - 1) * h.e.sa_i + (size_t)(h.n - 1) * h.e.sa_j + (size_t)(h.e.k - 1) * h.e.sa_k + 1;
            const size_t e_b = (size_t)(h.m - 1) * h.e.sb_i + (size_t)(h.n - 1) * h.e.sb_j + (size_t)(h.e.k - 1) * h.e.sb_k + 1;
            if (double *d = have_map ? b2g_map_lookup(ctx, h.e.a, e_a * 8) : nullptr)
                h.e.a = d, ctx->resident_hit_bytes += (int64_t)e_a * 8;
            else
                add(in_rg, h.e.a, e_a);
            if (double *d = have_map ? b2g_map_lookup(ctx, h.e.b, e_b * 8) : nullptr)
                h.e.b = d;
            else
                add(in_rg, h.e.b, e_b);
        }
    }
    // dense windows (whole rows of their block, or a single row) are addressed as one contiguous vector
    // whatever logical shape the contributing entry has: a term and its transposed partner, or the
    // recorder's whole-block AXPY (a.n == c.n branch), then share one cluster
    for (HostEntry &h : he)
        if (h.n > 1 && (h.ldc == h.n || h.m == 1) && (int64_t)h.m * h.n < INT32_MAX) {
            h.e.nd = h.n;
            h.m = h.m * h.n, h.n = 1, h.ldc = 1;
            // sources that are themselves linear in the flat element index need no (i, j) split
            if ((int64_t)h.e.sa_i == (int64_t)h.e.nd * h.e.sa_j && (int64_t)h.e.sb_i == (int64_t)h.e.nd * h.e.sb_j) {
                h.e.sa_i = h.e.sa_j, h.e.sb_i = h.e.sb_j;
                h.e.sa_j = h.e.sb_j = 0, h.e.nd = 0;
            }
        }

    // ---- 2. clusters = identical output windows, members in list order
    std::vector<int> idx(he.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&he](int x, int y) {
        const HostEntry &p = he[x], &q = he[y];
        if (p.dst != q.dst)
            return p.dst < q.dst;
        if (p.m != q.m)
            return p.m < q.m;
        if (p.n != q.n)
            return p.n < q.n;
        if (p.ldc != q.ldc)
            return p.ldc < q.ldc;
        return p.order < q.order;
    });
    struct Cluster {
        int first, count; // into idx
    };
    std::vector<Cluster> cl;
    for (size_t i = 0; i < idx.size();) {
        size_t j = i + 1;
        const HostEntry &p = he[idx[i]];
        while (j < idx.size() && he[idx[j]].dst == p.dst && he[idx[j]].m == p.m && he[idx[j]].n == p.n &&
               he[idx[j]].ldc == p.ldc)
            j++;
        cl.push_back(Cluster{(int)i, (int)(j - i)});
        i = j;
    }
    st.clusters = (int64_t)cl.size();

    // ---- 3. windows that overlap without being identical -> serial components (sweep over sorted starts)
    UF uf(cl.size());
    std::vector<char> irregular(cl.size(), 0);
    {
        std::vector<int> active;
        for (size_t ci = 0; ci < cl.size(); ci++) {
            const HostEntry &w = he[idx[cl[ci].first]];
            const uintptr_t lo = (uintptr_t)w.dst;
            size_t keep = 0;
            for (size_t q = 0; q < active.size(); q++) {
                const HostEntry &o = he[idx[cl[active[q]].first]];
                if ((uintptr_t)o.dst + window_extent(o.m, o.n, o.ldc) * 8 > lo)
                    active[keep++] = active[q];
            }
            active.resize(keep);
            for (int q : active)
                if (windows_overlap(he[idx[cl[q].first]], w)) {
                    irregular[q] = irregular[ci] = 1;
                    uf.unite(q, (int)ci);
                }
            active.push_back((int)ci);
        }
    }

    // ---- 4. device descriptors
    std::vector<BlkEntry> dev_entries;
    std::vector<BlkUnit> units, gunits; // AXPY windows / windows with a general contribution
    std::vector<TileUnit> tunits;       // AXPY windows with transposed sources among several contributions
    // windows whose 1..MULTI_MAX contributions are all linear (the bulk of a blocking step): rows [i0, i0 + rows)
    // x columns [j0, j0 + len) per unit, turned into self-contained stream / multi descriptors once the operands
    // have their device addresses.  One unit spans several rows, so the descriptor count follows the bytes of
    // the step (M^2), not the number of window rows times the number of terms.
    struct LinUnit {
        double *dst;
        int32_t i0, rows, j0, len, n, ldc, first, count;
    };
    std::vector<LinUnit> lunits;
    int64_t lin_target = UNIT_ELEMS; // elements x contributions per linear unit: ~32 units per warp of the grid
    {
        int64_t tot = 0;
        for (size_t ci = 0; ci < cl.size(); ci++)
            if (!irregular[ci])
                tot += (int64_t)he[idx[cl[ci].first]].m * he[idx[cl[ci].first]].n * cl[ci].count;
        const int64_t warps = (int64_t)(ctx ? ctx->sm_count : 148) * 3 * (BLK_THREADS / 32);
        // upper bound of a unit (B2G_BLK_LINMAX).  Measured on the Cr2 M=4000 H_eff list: 9.2 / 9.0 / 6.6 / 4.8 ms with
        // 2048 / 4096 / 8192 / 16384 elements - a CTA pays the latency of its descriptor and of the first loads of
        // its warps once per unit, so small units lose more than the shorter L2 reuse distance gains
        static const int64_t lin_max = getenv("B2G_BLK_LINMAX") ? atoll(getenv("B2G_BLK_LINMAX")) : 16384;
        lin_target = std::min<int64_t>(lin_max, std::max<int64_t>(UNIT_ELEMS, tot / (warps * 32) / 256 * 256));
    }
    static const bool tile_on = getenv("B2G_BLK_NOTILE") == nullptr;
    dev_entries.reserve(he.size());
    std::vector<BlkSerial> serial;
    for (size_t ci = 0; ci < cl.size(); ci++) {
        const HostEntry &w = he[idx[cl[ci].first]];
        if (irregular[ci]) {
            for (int q = 0; q < cl[ci].count; q++) {
                const HostEntry &h = he[idx[cl[ci].first + q]];
                serial.push_back(BlkSerial{h.e, h.dst, h.m, h.n, h.ldc, uf.find((int)ci)});
            }
            continue;
        }
        const int first = (int)dev_entries.size();
        bool axpy = true;
        for (int q = 0; q < cl[ci].count; q++) {
            const BlkEntry &e = he[idx[cl[ci].first + q]].e;
            axpy = axpy && (e.alpha == 0.0 || (e.k == 1 && e.sb_i == 0 && e.sb_j == 0));
            dev_entries.push_back(e);
        }
        const int64_t total = (int64_t)w.m * w.n;
        std::vector<BlkUnit> &dstu = axpy ? units : gunits;
        if (total >= INT32_MAX) {
            b2g_set_error(std::string(who) + ": output window of 2^31 or more elements");
            return 1;
        }
        // unit size by cost (elements x contributions) so that a window with many contributions is
        // spread over many warps; wide 2-D windows are cut into row-aligned units (linear addressing)
        static const int64_t cap_min = getenv("B2G_BLK_CAPMIN") ? atoll(getenv("B2G_BLK_CAPMIN")) : 128;
        static const int64_t cap_num = getenv("B2G_BLK_CAPNUM") ? atoll(getenv("B2G_BLK_CAPNUM")) : 4 * UNIT_ELEMS;
        const int64_t cap = std::max<int64_t>(cap_min, std::min<int64_t>(UNIT_ELEMS, cap_num / cl[ci].count) / 128 * 128);
        bool transposed_src = false;
        // every contribution linear in the window element, read (alpha != 0) and k = 1
        bool lin = axpy && cl[ci].count <= MULTI_MAX && (w.n >= ROW_MIN || w.n == 1);
        for (int q = 0; q < cl[ci].count; q++) {
            const BlkEntry &e = dev_entries[first + q];
            transposed_src = transposed_src || (e.alpha != 0.0 && e.sa_j != 1);
            lin = lin && e.nd == 0 && e.alpha != 0.0 && e.k == 1;
        }
        if (cl[ci].count == 1) // a single contribution overwrites the window
            lin = lin && (dst_zero || dev_entries[first].beta == 0.0);
        // tile units: several contributions with transposed sources among them, and every 2-D window of
        // 32..127 columns (too narrow for row units: the tile kernel needs no per-element division and reads
        // transposed sources along their contiguous direction)
        static const bool tile_narrow = getenv("B2G_BLK_NOTILE_NARROW") == nullptr;
        if (axpy && tile_on && w.n >= TILE_C && w.m >= 2 &&
            ((transposed_src && cl[ci].count >= 2) || (tile_narrow && w.n < ROW_MIN))) {
            // a unit = whole tile rows, about TILE_STRIP tiles (cost: tiles x contributions)
            const int64_t tpr = (w.n + TILE_C - 1) / TILE_C;
            const int64_t want = std::max<int64_t>(8, std::min<int64_t>(TILE_STRIP, 4 * TILE_STRIP / cl[ci].count));
            const int64_t trows = std::max<int64_t>(1, want / tpr); // tile rows per unit
            for (int64_t i0 = 0; i0 < w.m; i0 += TILE_R * trows) {
                const int64_t nrows = std::min<int64_t>(trows, (w.m - i0 + TILE_R - 1) / TILE_R);
                tunits.push_back(TileUnit{w.dst, (int32_t)i0, 0, w.m, w.n, w.ldc, first, cl[ci].count,
                                          (int32_t)(nrows * tpr)});
            }
        } else if (lin) {
            // elements per unit: the warps of a CTA share a unit piece by piece (b2g_blocking_stream_kernel)
            // (not divided by the number of contributions: a multi-source unit needs at least one piece per warp
            // to keep the rings of all 8 warps busy - ncu showed the multi kernel at 0.86 TB/s with 2048-element units)
            static const bool per_div = getenv("B2G_BLK_PERDIV") != nullptr; // A/B: the old rule
            const int64_t per = std::max<int64_t>(2048, lin_target / (per_div ? cl[ci].count : 1) / 256 * 256);
            if (w.n == 1) { // a column (or a dense window addressed flat): one "row" of elements ldc apart
                for (int64_t e0 = 0; e0 < total; e0 += per)
                    lunits.push_back(LinUnit{w.dst, 0, 1, (int32_t)e0, (int32_t)std::min<int64_t>(per, total - e0), w.n,
                                             w.ldc, first, cl[ci].count});
            } else { // whole rows
                const int64_t rows = std::max<int64_t>(1, per / w.n);
                for (int64_t i0 = 0; i0 < w.m; i0 += rows)
                    lunits.push_back(LinUnit{w.dst, (int32_t)i0, (int32_t)std::min<int64_t>(rows, w.m - i0), 0, w.n, w.n,
                                             w.ldc, first, cl[ci].count});
            }
        } else if (!axpy || w.n < ROW_MIN) {
            for (int64_t e0 = 0; e0 < total; e0 += cap)
                dstu.push_back(BlkUnit{w.dst, (int32_t)e0, (int32_t)std::min<int64_t>(cap, total - e0), w.n, w.ldc, first,
                                       cl[ci].count});
        } else {
            for (int64_t i = 0; i < w.m; i++)
                for (int64_t j0 = 0; j0 < w.n; j0 += cap)
                    dstu.push_back(BlkUnit{w.dst, (int32_t)(i * w.n + j0), (int32_t)std::min<int64_t>(cap, w.n - j0), w.n,
                                           w.ldc, first, cl[ci].count});
        }
        st.bytes_out += total * 8;
    }
    // serial entries: by component, then list order (order kept in a side array: pad is 31-bit only)
    std::vector<int64_t> comp_first;
    if (!serial.empty()) {
        std::vector<int64_t> sorder;
        sorder.reserve(serial.size());
        for (size_t ci = 0; ci < cl.size(); ci++)
            if (irregular[ci])
                for (int q = 0; q < cl[ci].count; q++)
                    sorder.push_back(he[idx[cl[ci].first + q]].order);
        std::vector<int> sidx(serial.size());
        std::iota(sidx.begin(), sidx.end(), 0);
        std::sort(sidx.begin(), sidx.end(), [&](int x, int y) {
            if (serial[x].comp != serial[y].comp)
                return serial[x].comp < serial[y].comp;
            return sorder[x] < sorder[y];
        });
        std::vector<BlkSerial> tmp(serial.size());
        for (size_t i = 0; i < sidx.size(); i++)
            tmp[i] = serial[sidx[i]];
        serial.swap(tmp);
        for (size_t i = 0; i < serial.size(); i++)
            if (i == 0 || serial[i].comp != serial[i - 1].comp)
                comp_first.push_back((int64_t)i);
        comp_first.push_back((int64_t)serial.size());
        for (const BlkSerial &s : serial)
            st.bytes_out += (int64_t)s.m * s.n * 8;
    }
    st.units = (int64_t)(units.size() + gunits.size() + tunits.size()); // before the streaming split
    st.serial_entries = (int64_t)serial.size();
    if (dst_zero && operand_space == B2G_OPERANDS_DEVICE && !serial.empty()) {
        b2g_set_error("" + std::string(who) + ": B2G_DST_ZERO with device operands needs regular (identical or disjoint) output windows");
        return 1;
    }
    st.plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    prof_lap("blocking.regroup");
    if (flags & B2G_PLAN_ONLY) { // the regrouping alone (host work, no device): counts for tests and tools
        st.units = (int64_t)(units.size() + gunits.size() + tunits.size() + lunits.size());
        if (getenv("B2G_VERBOSE"))
            fprintf(stderr, "[b2g] blocking plan: %zu accumulate units, %zu general, %zu tile, %zu linear, %zu serial\n",
                    units.size(), gunits.size(), tunits.size(), lunits.size(), serial.size());
        if (stats)
            *stats = st;
        return 0;
    }

    // ---- 5. operands
    double *d_in = nullptr, *d_out = nullptr;
    BlkEntry *d_entries = nullptr;
    BlkUnit *d_units = nullptr, *d_gunits = nullptr;
    StreamUnit *d_sunits = nullptr;
    MultiUnit *d_munits = nullptr;
    TileUnit *d_tunits = nullptr;
    BlkSerial *d_serial = nullptr;
    int64_t *d_comp = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    auto cleanup = [&]() {
        b2g_dfree(ctx, d_in), b2g_dfree(ctx, d_out), b2g_dfree(ctx, d_entries), b2g_dfree(ctx, d_units), b2g_dfree(ctx, d_gunits), b2g_dfree(ctx, d_sunits), b2g_dfree(ctx, d_munits), b2g_dfree(ctx, d_tunits);
        b2g_dfree(ctx, d_serial), b2g_dfree(ctx, d_comp);
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
    };
    auto fail = [&](const std::string &msg) {
        if (!msg.empty())
            b2g_set_error(msg);
        cudaStreamSynchronize(ctx->stream);
        cleanup();
        return 1;
    };
    auto t_up = std::chrono::steady_clock::now();
    if (operand_space == B2G_OPERANDS_HOST) {
        size_t in_total = 0, out_total = 0;
        b2g_merge_ranges(in_rg, in_total), b2g_merge_ranges(out_rg, out_total);
        for (const B2GRange &o : out_rg) // an output block must not also be an input of the same list
            if (!in_rg.empty()) {
                const B2GRange &r = b2g_locate_range(in_rg, o.lo);
                const B2GRange *nx = (&r + 1 < in_rg.data() + in_rg.size()) ? &r + 1 : nullptr;
                if ((r.lo < o.hi && o.lo < r.hi) || (nx && nx->lo < o.hi && o.lo < nx->hi))
                    return fail("" + std::string(who) + ": an output block aliases an input block of the same list");
            }
        if ((in_total && b2g_dmalloc(ctx, (void **)&d_in, in_total * sizeof(double))) ||
            (out_total && b2g_dmalloc(ctx, (void **)&d_out, out_total * sizeof(double))))
            return fail("");
        if (in_total && b2g_mirror_ranges(ctx, in_rg, d_in))
            return fail("");
        if (out_total) {
            if (dst_zero) {
                if (cudaMemsetAsync(d_out, 0, out_total * sizeof(double), ctx->stream) != cudaSuccess)
                    return fail("" + std::string(who) + ": memset failed");
            } else if (b2g_mirror_ranges(ctx, out_rg, d_out))
                return fail("");
        }
        // host addresses -> mirror; operands that already are device addresses (resident blocks) stay
        auto xin = [&](const double *ptr) -> const double * {
            if (in_rg.empty())
                return ptr;
            const B2GRange &r = b2g_locate_range(in_rg, (uintptr_t)ptr);
            return r.lo <= (uintptr_t)ptr && (uintptr_t)ptr < r.hi ? d_in + r.dev_off + ((uintptr_t)ptr - r.lo) / sizeof(double)
                                                                  : ptr;
        };
        auto xout = [&](double *ptr) -> double * {
            if (out_rg.empty())
                return ptr;
            const B2GRange &r = b2g_locate_range(out_rg, (uintptr_t)ptr);
            return r.lo <= (uintptr_t)ptr && (uintptr_t)ptr < r.hi ? d_out + r.dev_off + ((uintptr_t)ptr - r.lo) / sizeof(double)
                                                                  : ptr;
        };
        for (BlkEntry &e : dev_entries)
            if (e.alpha != 0.0 && e.k > 0)
                e.a = xin(e.a), e.b = xin(e.b);
        for (BlkUnit &u : units)
            u.dst = xout(u.dst);
        for (BlkUnit &u : gunits)
            u.dst = xout(u.dst);
        for (TileUnit &u : tunits)
            u.dst = xout(u.dst);
        for (LinUnit &u : lunits)
            u.dst = xout(u.dst);
        for (BlkSerial &s : serial) {
            if (s.e.alpha != 0.0 && s.e.k > 0)
                s.e.a = xin(s.e.a), s.e.b = xin(s.e.b);
            s.dst = xout(s.dst);
        }
    }
    prof_lap("blocking.mirror+translate");
    // linear units become self-contained stream (one contribution) / multi (2..MULTI_MAX) descriptors
    std::vector<StreamUnit> sunits;
    std::vector<MultiUnit> munits;
    for (const LinUnit &u : lunits) {
        const bool flat = u.n == 1;
        const int64_t doff = flat ? (int64_t)u.j0 * u.ldc : (int64_t)u.i0 * u.ldc + u.j0;
        auto soff = [&](const BlkEntry &E) {
            return flat ? (int64_t)u.j0 * E.sa_i : (int64_t)u.i0 * E.sa_i + (int64_t)u.j0 * E.sa_j;
        };
        if (u.count == 1) {
            const BlkEntry &E = dev_entries[u.first];
            StreamUnit su;
            su.dst = u.dst + doff, su.src = E.a + soff(E), su.b = E.b, su.alpha = E.alpha;
            su.len = u.len, su.dstep = flat ? u.ldc : 1, su.sstep = flat ? E.sa_i : E.sa_j;
            su.rows = u.rows, su.drow = flat ? 0 : u.ldc, su.srow = flat ? 0 : E.sa_i, su.pad2 = 0;
            sunits.push_back(su);
        } else {
            MultiUnit mu;
            memset(&mu, 0, sizeof(mu));
            mu.dst = u.dst + doff;
            for (int t = 0; t < u.count; t++) {
                const BlkEntry &Et = dev_entries[u.first + t];
                mu.src[t] = Et.a + soff(Et), mu.b[t] = Et.b, mu.alpha[t] = Et.alpha, mu.beta[t] = Et.beta;
                mu.sstep[t] = flat ? Et.sa_i : Et.sa_j, mu.srow[t] = flat ? 0 : Et.sa_i;
            }
            mu.len = u.len, mu.dstep = flat ? u.ldc : 1, mu.count = u.count, mu.rows = u.rows;
            mu.drow = flat ? 0 : u.ldc;
            munits.push_back(mu);
        }
    }
    // Source-major order: the same environment block feeds several windows (a source is read by ~2 terms on
    // average in an H_eff blocking step), and a kernel works through its unit array front to back with all
    // warps of the grid, so units that read the same source run at the same time and the second reader finds
    // the block in L2.  Outputs are written once whatever the order.  (B2G_BLK_NOSORT: list order, for A/B.)
    static const bool source_major = getenv("B2G_BLK_NOSORT") == nullptr;
    if (source_major) {
        auto reorder = [](auto &vec, auto key_of) {
            typedef typename std::remove_reference<decltype(vec)>::type Vec;
            const size_t n = vec.size();
            if (n < 2)
                return;
            std::vector<std::pair<uintptr_t, uint32_t>> key(n);
            for (size_t i = 0; i < n; i++)
                key[i] = std::make_pair((uintptr_t)key_of(vec[i]), (uint32_t)i);
            std::sort(key.begin(), key.end());
            Vec out(n);
            for (size_t i = 0; i < n; i++)
                out[i] = vec[key[i].second];
            vec.swap(out);
        };
        reorder(sunits, [](const StreamUnit &u) { return u.src; });
        reorder(munits, [](const MultiUnit &u) { return u.src[0]; });
        reorder(tunits, [&dev_entries](const TileUnit &u) {
            const BlkEntry &e = dev_entries[u.first];
            return e.a + (int64_t)u.i0 * e.sa_i + (int64_t)u.j0 * e.sa_j;
        });
        reorder(units, [&dev_entries](const BlkUnit &u) {
            const BlkEntry &e = dev_entries[u.first];
            const int w = e.nd ? e.nd : u.n; // logical width of the addressing: i = e / w, j = e % w
            return e.a + (w > 1 ? (int64_t)(u.e0 / w) * e.sa_i + (int64_t)(u.e0 % w) * e.sa_j : (int64_t)u.e0 * e.sa_i);
        });
    }
    prof_lap("blocking.stream_units");
    // descriptors (pageable -> device; synchronised below before the vectors die)
    if (!dev_entries.empty()) {
        if (b2g_dmalloc(ctx, (void **)&d_entries, dev_entries.size() * sizeof(BlkEntry)) ||
            b2g_dmalloc(ctx, (void **)&d_units, units.size() * sizeof(BlkUnit)) ||
            b2g_dmalloc(ctx, (void **)&d_gunits, gunits.size() * sizeof(BlkUnit)))
            return fail("");
        if (cudaMemcpyAsync(d_entries, dev_entries.data(), dev_entries.size() * sizeof(BlkEntry),
                            cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_units, units.data(), units.size() * sizeof(BlkUnit), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_gunits, gunits.data(), gunits.size() * sizeof(BlkUnit), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess)
            return fail("" + std::string(who) + ": descriptor upload failed");
    }
    if (!sunits.empty()) {
        if (b2g_dmalloc(ctx, (void **)&d_sunits, sunits.size() * sizeof(StreamUnit)))
            return fail("");
        if (cudaMemcpyAsync(d_sunits, sunits.data(), sunits.size() * sizeof(StreamUnit), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess)
            return fail(std::string(who) + ": descriptor upload failed");
    }
    if (!tunits.empty()) {
        if (b2g_dmalloc(ctx, (void **)&d_tunits, tunits.size() * sizeof(TileUnit)))
            return fail("");
        if (cudaMemcpyAsync(d_tunits, tunits.data(), tunits.size() * sizeof(TileUnit), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess)
            return fail(std::string(who) + ": descriptor upload failed");
    }
    if (!munits.empty()) {
        if (b2g_dmalloc(ctx, (void **)&d_munits, munits.size() * sizeof(MultiUnit)))
            return fail("");
        if (cudaMemcpyAsync(d_munits, munits.data(), munits.size() * sizeof(MultiUnit), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess)
            return fail(std::string(who) + ": descriptor upload failed");
    }
    if (!serial.empty()) {
        if (b2g_dmalloc(ctx, (void **)&d_serial, serial.size() * sizeof(BlkSerial)) ||
            b2g_dmalloc(ctx, (void **)&d_comp, comp_first.size() * sizeof(int64_t)))
            return fail("");
        if (cudaMemcpyAsync(d_serial, serial.data(), serial.size() * sizeof(BlkSerial), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_comp, comp_first.data(), comp_first.size() * sizeof(int64_t), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess)
            return fail("" + std::string(who) + ": descriptor upload failed");
    }
    prof_lap("blocking.desc_upload_issue");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return fail("" + std::string(who) + ": operand upload failed");
    prof_lap("blocking.desc_upload_sync");
    st.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_up).count();

    // ---- 6. kernels
    if (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess)
        return fail("" + std::string(who) + ": event creation failed");
    cudaEventRecord(ev0, ctx->stream);
    static const int acc_stages = getenv("B2G_BLK_STAGES") ? atoi(getenv("B2G_BLK_STAGES")) : ACC_STAGES_DEFAULT;
    if (!ctx->blocking_attr_set) {
        if (cudaFuncSetAttribute(b2g_blocking_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)acc_ring_bytes(3)) != cudaSuccess ||
            cudaFuncSetAttribute(b2g_blocking_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)acc_ring_bytes(4)) != cudaSuccess ||
            cudaFuncSetAttribute(b2g_blocking_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)acc_ring_bytes(5)) != cudaSuccess ||
            cudaFuncSetAttribute(b2g_blocking_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)RING_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(b2g_blocking_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)RING_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(b2g_blocking_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)TILE_RING_BYTES) != cudaSuccess)
            return fail(std::string(who) + ": cudaFuncSetAttribute failed");
        ctx->blocking_attr_set = true;
    }
    if (!sunits.empty()) {
        const int grid = (int)std::min<int64_t>((int64_t)sunits.size(), (int64_t)ctx->sm_count * 3); // 3 CTAs per SM
        b2g_blocking_stream_kernel<<<grid, BLK_THREADS, RING_BYTES, ctx->stream>>>(d_sunits, (int64_t)sunits.size());
        ctx->launches++, st.launches++;
    }
    if (!munits.empty()) {
        const int grid = (int)std::min<int64_t>((int64_t)munits.size(), (int64_t)ctx->sm_count * 3);
        b2g_blocking_multi_kernel<<<grid, BLK_THREADS, RING_BYTES, ctx->stream>>>(d_munits, (int64_t)munits.size(),
                                                                                  dst_zero ? 1 : 0);
        ctx->launches++, st.launches++;
    }
    if (!tunits.empty()) {
        const int grid = (int)std::min<int64_t>((int64_t)tunits.size(), (int64_t)ctx->sm_count * 2);
        b2g_blocking_tile_kernel<<<grid, BLK_THREADS, TILE_RING_BYTES, ctx->stream>>>(d_tunits, (int64_t)tunits.size(),
                                                                                      d_entries, dst_zero ? 1 : 0);
        ctx->launches++, st.launches++;
    }
    cudaEvent_t evs = nullptr;
    const bool verbose = getenv("B2G_VERBOSE") != nullptr;
    if (verbose) {
        cudaEventCreate(&evs);
        cudaEventRecord(evs, ctx->stream);
    }
    if (!units.empty()) {
        const int64_t want = ((int64_t)units.size() * 32 + BLK_THREADS - 1) / BLK_THREADS;
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 2); // 2 resident CTAs per SM
        if (acc_stages == 5)
            b2g_blocking_kernel<5><<<grid, BLK_THREADS, acc_ring_bytes(5), ctx->stream>>>(d_units, (int64_t)units.size(),
                                                                                         d_entries, dst_zero ? 1 : 0);
        else if (acc_stages == 4)
            b2g_blocking_kernel<4><<<grid, BLK_THREADS, acc_ring_bytes(4), ctx->stream>>>(d_units, (int64_t)units.size(),
                                                                                         d_entries, dst_zero ? 1 : 0);
        else
            b2g_blocking_kernel<3><<<grid, BLK_THREADS, acc_ring_bytes(3), ctx->stream>>>(d_units, (int64_t)units.size(),
                                                                                         d_entries, dst_zero ? 1 : 0);
        ctx->launches++, st.launches++;
    }
    if (!gunits.empty()) {
        const int64_t want = ((int64_t)gunits.size() * 32 + BLK_THREADS - 1) / BLK_THREADS;
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 8);
        b2g_blocking_general_kernel<<<grid, BLK_THREADS, 0, ctx->stream>>>(d_gunits, (int64_t)gunits.size(), d_entries,
                                                                           dst_zero ? 1 : 0);
        ctx->launches++, st.launches++;
    }
    if (!serial.empty()) {
        const int ncomp = (int)comp_first.size() - 1;
        b2g_blocking_serial_kernel<<<std::min(ncomp, ctx->sm_count * 4), BLK_THREADS, 0, ctx->stream>>>(d_serial, d_comp,
                                                                                                       ncomp);
        ctx->launches++, st.launches++;
    }
    cudaEventRecord(ev1, ctx->stream);
    if (cudaGetLastError() != cudaSuccess || cudaEventSynchronize(ev1) != cudaSuccess)
        return fail(std::string("" + std::string(who) + ": kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms = ms;
    prof_lap("blocking.kernels+sync");
    b2g_prof_record("blocking.kernels_gpu", ms * 1e-3);
    if (verbose) {
        float ms_s = 0;
        cudaEventElapsedTime(&ms_s, ev0, evs);
        size_t se = 0, re = 0;
        for (const StreamUnit &u : sunits)
            se += (size_t)u.len;
        for (const BlkUnit &u : units)
            re += (size_t)u.len;
        fprintf(stderr, "[b2g] blocking: tile %zu units | multi %zu units | stream %zu units %zu elements %.3f ms | regular %zu units %zu elements, general %zu units, "
                        "serial %zu entries %.3f ms\n",
                tunits.size(), munits.size(), sunits.size(), se, ms_s, units.size(), re, gunits.size(), serial.size(), ms - ms_s);
        cudaEventDestroy(evs);
    }

    // ---- 7. results (windows of resident output blocks were written in place and stay on the device)
    if (operand_space == B2G_OPERANDS_HOST && !out_rg.empty()) {
        auto t_dn = std::chrono::steady_clock::now();
        if (b2g_download_ranges(ctx, out_rg, d_out, dst_zero))
            return fail("");
        st.download_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_dn).count();
    }
    cleanup();
    prof_lap("blocking.download+cleanup");
    if (stats)
        *stats = st;
    return 0;
}

extern "C" int b2g_batch_execute(b2g_context *ctx, int64_t group_count, const int32_t *ta, const int32_t *tb,
                                 const int32_t *m, const int32_t *n, const int32_t *k, const double *alpha,
                                 const double *const *a, const int32_t *lda, const double *const *b,
                                 const int32_t *ldb, const double *beta, double *const *c, const int32_t *ldc,
                                 const int32_t *group_size, int operand_space, int flags,
                                 b2g_blocking_stats *stats) {
    if (!ctx && !(flags & B2G_PLAN_ONLY)) {
        b2g_set_error("b2g_batch_execute: null context");
        return 1;
    }
    if (group_count > 0 && (!ta || !tb || !m || !n || !k || !alpha || !a || !lda || !b || !ldb || !beta || !c || !ldc ||
                            !group_size)) {
        b2g_set_error("b2g_batch_execute: null argument");
        return 1;
    }
    if (operand_space != B2G_OPERANDS_HOST && operand_space != B2G_OPERANDS_DEVICE) {
        b2g_set_error("b2g_batch_execute: unknown operand space");
        return 1;
    }
    if (!(flags & B2G_PLAN_ONLY))
        B2G_CUDA(cudaSetDevice(ctx->device));
    b2g_blocking_stats st;
    memset(&st, 0, sizeof(st));
    auto t_begin = std::chrono::steady_clock::now();
    // ---- 1. expand the groups, fold constant-stride AXPY rows of one group into 2-D windows
    std::vector<HostEntry> he;
    int64_t z = 0, order = 0;
    for (int64_t g = 0; g < group_count; g++) {
        if (!ok_t(ta[g]) || !ok_t(tb[g])) {
            b2g_set_error("b2g_batch_execute: group " + std::to_string(g) + ": transpose flag is not N/T");
            return 1;
        }
        const bool tA = is_t(ta[g]), tB = is_t(tb[g]);
        const int32_t gm = m[g], gn = n[g], gk = k[g], gs = group_size[g];
        if (gs < 0 || gk < 0) {
            b2g_set_error("b2g_batch_execute: negative group size or k");
            return 1;
        }
        if (gm <= 0 || gn <= 0) {
            z += gs;
            continue;
        }
        if (ldc[g] < gn || (gk > 0 && (lda[g] < (tA ? gm : gk) || ldb[g] < (tB ? gk : gn)))) {
            b2g_set_error("b2g_batch_execute: group " + std::to_string(g) + ": leading dimension too small");
            return 1;
        }
        st.entries += gs;
        st.nflop_mnk += (int64_t)gm * gn * gk * gs;
        const bool foldable = gk == 1 && gn == 1 && ldc[g] == 1;
        for (int32_t q = 0; q < gs;) {
            HostEntry h;
            h.e.a = a[z + q], h.e.b = b[z + q], h.dst = c[z + q];
            h.e.alpha = alpha[g], h.e.beta = beta[g];
            h.e.sa_i = tA ? 1 : lda[g], h.e.sa_k = tA ? lda[g] : 1, h.e.sa_j = 0;
            h.e.sb_j = tB ? ldb[g] : 1, h.e.sb_k = tB ? 1 : ldb[g], h.e.sb_i = 0;
            h.e.k = gk, h.e.nd = 0;
            h.m = gm, h.n = gn, h.ldc = ldc[g];
            h.order = order;
            int32_t run = 1;
            if (foldable && q + 1 < gs) {
                const int64_t da = a[z + q + 1] - a[z + q], db = b[z + q + 1] - b[z + q],
                              dc = c[z + q + 1] - c[z + q];
                if (dc >= gm && dc < INT32_MAX && da >= 0 && da < INT32_MAX && db >= 0 && db < INT32_MAX) {
                    while (q + run < gs && a[z + q + run] - a[z + q + run - 1] == da &&
                           b[z + q + run] - b[z + q + run - 1] == db && c[z + q + run] - c[z + q + run - 1] == dc)
                        run++;
                    if (run > 1) { // rows r = 0..run-1, columns = the gm elements of one row
                        h.e.sa_j = h.e.sa_i, h.e.sa_i = (int32_t)da;
                        h.e.sb_i = (int32_t)db, h.e.sb_j = 0;
                        h.m = run, h.n = gm, h.ldc = (int32_t)dc;
                    }
                }
            }
            he.push_back(h);
            order += run, q += run;
        }
        z += gs;
    }
    return execute_entries(ctx, he, operand_space, flags, st, stats, t_begin, "b2g_batch_execute");
}

// One GMatrixFunctions::tensor_product call (block2 src/core/matrix_functions.hpp:1269-1397, recorded
// form AdvancedGEMM<double>::tensor_product, src/core/batch_gemm.hpp:433-503):
//     C[(i*bm + k), (j*bn + l)] += scale * op(A)(i, j) * op(B)(k, l),   C = c (stride already added), pitch cn
// as whole 2-D windows instead of one GEMM per row.
extern "C" int b2g_tensor_product_execute(b2g_context *ctx, int64_t count, const b2g_tp_term *terms,
                                          int operand_space, int flags, b2g_blocking_stats *stats) {
    if ((!ctx && !(flags & B2G_PLAN_ONLY)) || (count > 0 && !terms)) {
        b2g_set_error("b2g_tensor_product_execute: null argument");
        return 1;
    }
    if (operand_space != B2G_OPERANDS_HOST && operand_space != B2G_OPERANDS_DEVICE) {
        b2g_set_error("b2g_tensor_product_execute: unknown operand space");
        return 1;
    }
    if (!(flags & B2G_PLAN_ONLY))
        B2G_CUDA(cudaSetDevice(ctx->device));
    b2g_blocking_stats st;
    memset(&st, 0, sizeof(st));
    auto t_begin = std::chrono::steady_clock::now();
    std::vector<HostEntry> he;
    he.reserve((size_t)count);
    int64_t order = 0;
    for (int64_t t = 0; t < count; t++) {
        const b2g_tp_term &q = terms[t];
        if (q.am <= 0 || q.an <= 0 || q.bm <= 0 || q.bn <= 0)
            continue;
        const bool ca = q.conja != 0, cb = q.conjb != 0;
        // shape of op(A), op(B)
        const int32_t ar = ca ? q.an : q.am, ac = ca ? q.am : q.an, br = cb ? q.bn : q.bm, bc = cb ? q.bm : q.bn;
        if ((int64_t)ac * bc > q.cn) {
            b2g_set_error("b2g_tensor_product_execute: term " + std::to_string(t) + ": window wider than the pitch of c");
            return 1;
        }
        st.entries += 1;
        st.nflop_mnk += (int64_t)q.am * q.an * q.bm * q.bn;
        HostEntry h;
        h.e.alpha = q.scale, h.e.beta = 1.0, h.e.k = 1, h.e.nd = 0;
        h.e.sa_k = h.e.sb_k = 0;
        if (q.bm == 1 && q.bn == 1) { // C(i, j) += scale * b * op(A)(i, j)
            h.e.a = q.a, h.e.b = q.b;
            h.e.sa_i = ca ? 1 : q.an, h.e.sa_j = ca ? q.an : 1, h.e.sb_i = h.e.sb_j = 0;
            h.dst = q.c, h.m = ar, h.n = ac, h.ldc = q.cn, h.order = order++;
            he.push_back(h);
        } else if (q.am == 1 && q.an == 1) { // C(k, l) += scale * a * op(B)(k, l)
            h.e.a = q.b, h.e.b = q.a;
            h.e.sa_i = cb ? 1 : q.bn, h.e.sa_j = cb ? q.bn : 1, h.e.sb_i = h.e.sb_j = 0;
            h.dst = q.c, h.m = br, h.n = bc, h.ldc = q.cn, h.order = order++;
            he.push_back(h);
        } else { // Kronecker product: one op(B) window per element of op(A)
            for (int32_t i = 0; i < ar; i++)
                for (int32_t j = 0; j < ac; j++) {
                    HostEntry g = h;
                    g.e.a = q.b, g.e.b = q.a + (ca ? (int64_t)j * q.an + i : (int64_t)i * q.an + j);
                    g.e.sa_i = cb ? 1 : q.bn, g.e.sa_j = cb ? q.bn : 1, g.e.sb_i = g.e.sb_j = 0;
                    g.dst = q.c + ((int64_t)i * br) * q.cn + (int64_t)j * bc;
                    g.m = br, g.n = bc, g.ldc = q.cn, g.order = order++;
                    he.push_back(g);
                }
        }
    }
    return execute_entries(ctx, he, operand_space, flags, st, stats, t_begin, "b2g_tensor_product_execute");
}
