// b2g_tiled.cuh — FP64 tensor-core (DMMA m8n8k4) tile engine of the H.C replay.
//
// The recorded pair list (block2 src/core/batch_gemm.hpp:564-575) is regrouped into two grid-wide phases of
// long-K GEMMs that share one tile engine:
//
//   phase 1  (one GEMM per GROUP = pairs that feed the same sigma window through the same operator block A1)
//            W_g = sum_p alpha_p * c[a0_p ..] * op(B0_p)           K-segments = the pairs of the group
//            written once into the column range of its window inside the W panel of (row panel, A1)
//   phase 2  (one GEMM per ROW PANEL = the sigma windows that share a row range of one sigma block)
//            sigma[rows, all columns of the block] += scale * [A1_s ...] * [W_s; ...]
//            K-segments = the distinct operator blocks A1 of the panel; W_s is one m0 x (block width) panel
//            that holds the W_g of every column window side by side (zero where a window has no term with
//            A1_s), so A1 is streamed once for all column windows of the row panel instead of once per window
//            and the narrow windows (n0 = 1..8) ride in the tiles of their neighbours.
//
// A "unit" is one CTA tile (BM x BN) times one K-chunk; units are sorted by cost and
// claimed through an atomic counter by a persistent grid.
//
// Tile engine: cp.async (LDGSTS) multi-stage ring -> padded shared-memory tiles -> DMMA
// fragments with conflict-free 64-bit LDS -> FP64 accumulators in registers.  Operand
// blocks have arbitrary 8-byte alignment and odd leading dimensions (sector dimensions),
// which rules out TMA tensor maps (16-byte strides) on the reference's layout.
#pragma once
#include "b2g_internal.h"

namespace b2g {

// Shared-memory tile pitches: a K-contiguous tile [row][k] has pitch BK + 4, an MN-contiguous tile
// [k][row] has pitch rows + 4; both are = 4 mod 16, which makes the 8x4 / 4x8 DMMA fragment reads of
// a half-warp hit 16 distinct bank pairs.  BK = K depth of one pipeline stage (16 or 32).

struct P1Seg {           // phase 1: one K-segment (= one pair)
    const double *b0;    // operator block (device)
    int64_t a_off;       // window offset inside c
    double alpha;        // alpha0 * alpha1 of the pair
    int32_t lda, ldb;
    int32_t k0, pad;
};

struct P1Group {         // phase 1: one output W_g (m0 x n0, leading dimension wld inside its W panel)
    int64_t w_off;       // element offset of W_g(0, 0) in the workspace
    int32_t wld, m0, n0;
    int32_t seg_begin, seg_end, pad;
};

struct P2Window {        // phase 2: one row panel of a sigma block
    int64_t c_off;       // offset of (row 0, column 0) of the panel inside sigma
    int32_t ldc, m1, n0, pad; // n0 = panel width in columns
};

struct P2Seg {           // phase 2: one K-segment (= one distinct operator block A1 of the panel)
    const double *a1;    // operator block (device)
    int64_t w_off;       // element offset of the W panel (m0 x wld) in the workspace
    int32_t lda, klen;   // leading dimension of a1, K length (= m0)
    int32_t wld;         // leading dimension of the W panel
    int32_t col_lo, col_hi, pad; // panel columns the W panel covers: [col_lo, col_hi), zero outside
};

struct MatvecArgs {      // c, sigma and scale of a replay, read by the kernels of a captured CUDA graph
    const double *c;
    double *v;
    double scale, pad;
};

struct Unit {            // one CTA-tile x K-chunk
    int32_t idx;         // phase 1: pair index; phase 2: window index
    int32_t row0, col0;  // origin of the tile inside the output matrix (elements)
    int32_t seg_begin, seg_end; // phase 2: segment range (phase 1: unused)
    int32_t pad;         // number of BK-deep stages of the unit
    int64_t poff;        // phase 2, deterministic mode: element offset of this unit's partial tile; else -1
};

struct OutTile {         // phase 2, deterministic mode: one sigma tile and its K-chunk partials
    int32_t win, row0, col0, bm, bn;
    int32_t part_begin, part_end; // range in the partial-offset array, in K-chunk order
    int32_t pad;
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    // ignore-src form: zeros are written and the source is not read when the element is clipped
    asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.ca.shared.global [%0], [%1], 8, p;\n}\n" ::"r"(s),
                 "l"(gmem), "r"((unsigned)valid));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Tile configuration: CTA tile BM x BN, WM x WN warps, each warp (BM/WM) x (BN/WN).
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_ = 1, int BK_ = 16> struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr int BK = BK_, KC_LD = BK_ + 4;
    static constexpr int MINB = MINB_; // resident CTAs per SM the register budget is capped for
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int WTM = BM / WM, WTN = BN / WN; // warp tile
    static constexpr int MI = WTM / 8, NI = WTN / 8;   // 8x8 DMMA blocks per warp
    static constexpr int A_STAGE = (BM * KC_LD > BK * (BM + 4)) ? BM * KC_LD : BK * (BM + 4);
    static constexpr int B_STAGE = (BN * KC_LD > BK * (BN + 4)) ? BN * KC_LD : BK * (BN + 4);
    static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);
};

// Issue the cp.async copies of one BK-deep stage of a ROWS-row operand tile.
//   KC = true : global operand is [row][k] (k contiguous, pitch ld)  -> smem [row][KC_LD]
//   KC = false: global operand is [k][row] (row contiguous, pitch ld) -> smem [k][ROWS + 4]
// rows_valid / k_valid clip the tile (zero fill outside).
// rows_lo..rows_valid / k_valid clip the tile: clipped elements are written as zeros and their (possibly out of
// range) source addresses are never read (ignore-src form of cp.async).
template <int ROWS, int THREADS, bool KC, int BK>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int ld, int rows_valid, int k_valid,
                                          int rows_lo = 0) {
    const int tid = threadIdx.x;
    if (KC) {
        constexpr int RPP = THREADS / BK; // rows per pass
        const int k = tid % BK, r0 = tid / BK;
        const bool kv = k < k_valid;
#pragma unroll
        for (int i = 0; i < (ROWS + RPP - 1) / RPP; i++) {
            const int r = r0 + i * RPP;
            if (ROWS % RPP == 0 || r < ROWS) {
                const bool ok = kv && r < rows_valid && r >= rows_lo;
                cp_async8(smem + r * (BK + 4) + k, g + (size_t)r * ld + k, ok);
            }
        }
    } else {
        constexpr int LDS_ = ROWS + 4;
        if (THREADS >= ROWS) {
            constexpr int KPP = THREADS / ROWS; // k rows per pass
            const int r = tid % ROWS, k0 = tid / ROWS;
            const bool rv = r < rows_valid && r >= rows_lo;
            if (THREADS % ROWS == 0 || k0 < KPP) {
#pragma unroll
                for (int i = 0; i < (BK + KPP - 1) / KPP; i++) {
                    const int k = k0 + i * KPP;
                    if (BK % KPP == 0 || k < BK) {
                        const bool ok = rv && k < k_valid;
                        cp_async8(smem + k * LDS_ + r, g + (size_t)k * ld + r, ok);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < BK; k++)
#pragma unroll
                for (int r = tid; r < ROWS; r += THREADS) {
                    const bool ok = r < rows_valid && r >= rows_lo && k < k_valid;
                    cp_async8(smem + k * LDS_ + r, g + (size_t)k * ld + r, ok);
                }
        }
    }
}

// One BK-deep stage of DMMAs for this warp. mi_n / ni_n: number of 8-row / 8-col blocks of the
// warp tile that intersect the valid output (warp-uniform), so edge tiles skip dead blocks.
// FULL = true: the warp tile lies entirely inside the output (no guards, no reconvergence
// points around the DMMAs - the common case); FULL = false: edge tiles.
// SCALE = true: the A fragments are multiplied by `alpha` (the factor of the K-segment this stage belongs to;
// phase 1 sums pairs with different factors into one accumulator).
template <class Cfg, bool A_KC, bool B_KC, bool FULL, bool SCALE>
__device__ __forceinline__ void compute_stage(const double *As, const double *Bs, double (&acc)[Cfg::MI][Cfg::NI][2],
                                              int wm0, int wn0, int mi_n, int ni_n, double alpha) {
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    const double *ap = A_KC ? As + (wm0 + lr) * Cfg::KC_LD + lc : As + lc * (Cfg::BM + 4) + wm0 + lr;
    const double *bp = B_KC ? Bs + (wn0 + lr) * Cfg::KC_LD + lc : Bs + lc * (Cfg::BN + 4) + wn0 + lr;
    // B fragments are fetched NC at a time: wide warp tiles (NI = 9) would otherwise hold 18 registers of them
    constexpr int NC = Cfg::NI > 4 ? 3 : Cfg::NI;
#pragma unroll
    for (int kk = 0; kk < Cfg::BK / 4; kk++) {
        double a[Cfg::MI];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
            if (FULL || mi < mi_n) {
                a[mi] = A_KC ? ap[mi * 8 * Cfg::KC_LD + kk * 4] : ap[kk * 4 * (Cfg::BM + 4) + mi * 8];
                if (SCALE)
                    a[mi] *= alpha;
            }
#pragma unroll
        for (int n0 = 0; n0 < Cfg::NI; n0 += NC) {
            double b[NC];
#pragma unroll
            for (int j = 0; j < NC; j++)
                if (n0 + j < Cfg::NI && (FULL || n0 + j < ni_n))
                    b[j] = B_KC ? bp[(n0 + j) * 8 * Cfg::KC_LD + kk * 4] : bp[kk * 4 * (Cfg::BN + 4) + (n0 + j) * 8];
#pragma unroll
            for (int mi = 0; mi < Cfg::MI; mi++)
                if (FULL || mi < mi_n) {
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        if (n0 + j < Cfg::NI && (FULL || n0 + j < ni_n))
                            dmma884(acc[mi][n0 + j][0], acc[mi][n0 + j][1], a[mi], b[j]);
                }
        }
    }
}

} // namespace b2g
