// b2g_tiled.cuh — FP64 tensor-core (DMMA m8n8k4) tile engine of the H.C replay.
//
// The recorded pair list (block2 src/core/batch_gemm.hpp:564-575) is executed in two
// grid-wide phases that share one tile engine:
//
//   phase 1  (one GEMM per pair)      W_p = (alpha0*alpha1) * c[a0_off..] * op(B0_p)
//                                     written once into a workspace in HBM/L2
//   phase 2  (one GEMM per window)    sigma[window] += scale * [A1_p ...] * [W_p; ...]
//                                     every pair writing the same sigma window becomes a
//                                     K-segment of ONE long-K GEMM, so sigma is touched once
//                                     per K-chunk instead of once per pair
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

struct P1Pair {          // phase 1: one pair
    const double *b0;    // operator block (device)
    int64_t w_off;       // element offset of W_p in the workspace (row-major m0 x n0, ld = n0)
    double alpha;        // alpha0 * alpha1
    int64_t a_off;       // window offset inside c
    int32_t lda, ldb;
    int32_t m0, n0, k0;
    int32_t tb0;
};

struct P2Window {        // phase 2: one sigma window
    int64_t c_off;
    int32_t ldc, m1, n0, pad;
};

struct P2Seg {           // phase 2: one K-segment (= one pair)
    const double *a1;    // operator block (device)
    int64_t w_off;       // element offset of W_p in the workspace
    int32_t lda, klen;   // leading dimension of a1, K length (= m0 of the pair)
};

struct SumTask {         // W[dst] += sum_j W[src_j]  over `count` doubles (pairs sharing sigma window and A1)
    int64_t dst, src[3];
    int64_t start, count; // element range of this work unit inside the blocks
    int32_t nsrc, pad;
};

struct Unit {            // one CTA-tile x K-chunk
    int32_t idx;         // phase 1: pair index; phase 2: window index
    int32_t row0, col0;  // origin of the tile inside the output matrix (elements)
    int32_t seg_begin, seg_end; // phase 2: segment range (phase 1: unused)
    int32_t pad;
    int64_t poff;        // phase 2, deterministic mode: element offset of this unit's partial tile; else -1
};

struct OutTile {         // phase 2, deterministic mode: one sigma tile and its K-chunk partials
    int32_t win, row0, col0, bm, bn;
    int32_t part_begin, part_end; // range in the partial-offset array, in K-chunk order
    int32_t pad;
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
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
template <int ROWS, int THREADS, bool KC, int BK>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int ld, int rows_valid, int k_valid) {
    const int tid = threadIdx.x;
    if (KC) {
        constexpr int RPP = THREADS / BK; // rows per pass
        const int k = tid % BK, r0 = tid / BK;
        const bool kv = k < k_valid;
#pragma unroll
        for (int i = 0; i < (ROWS + RPP - 1) / RPP; i++) {
            const int r = r0 + i * RPP;
            if (ROWS % RPP == 0 || r < ROWS) {
                const bool ok = kv && r < rows_valid;
                cp_async8(smem + r * (BK + 4) + k, ok ? g + (size_t)r * ld + k : g, ok);
            }
        }
    } else {
        constexpr int LDS_ = ROWS + 4;
        if (THREADS >= ROWS) {
            constexpr int KPP = THREADS / ROWS; // k rows per pass
            const int r = tid % ROWS, k0 = tid / ROWS;
            const bool rv = r < rows_valid;
#pragma unroll
            for (int i = 0; i < (BK + KPP - 1) / KPP; i++) {
                const int k = k0 + i * KPP;
                if (BK % KPP == 0 || k < BK) {
                    const bool ok = rv && k < k_valid;
                    cp_async8(smem + k * LDS_ + r, ok ? g + (size_t)k * ld + r : g, ok);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < BK; k++)
#pragma unroll
                for (int r = tid; r < ROWS; r += THREADS) {
                    const bool ok = r < rows_valid && k < k_valid;
                    cp_async8(smem + k * LDS_ + r, ok ? g + (size_t)k * ld + r : g, ok);
                }
        }
    }
}

// One BK-deep stage of DMMAs for this warp. mi_n / ni_n: number of 8-row / 8-col blocks of the
// warp tile that intersect the valid output (warp-uniform), so edge tiles skip dead blocks.
// FULL = true: the warp tile lies entirely inside the output (no guards, no reconvergence
// points around the DMMAs - the common case); FULL = false: edge tiles.
template <class Cfg, bool A_KC, bool B_KC, bool FULL>
__device__ __forceinline__ void compute_stage(const double *As, const double *Bs, double (&acc)[Cfg::MI][Cfg::NI][2],
                                              int wm0, int wn0, int mi_n, int ni_n) {
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    const double *ap = A_KC ? As + (wm0 + lr) * Cfg::KC_LD + lc : As + lc * (Cfg::BM + 4) + wm0 + lr;
    const double *bp = B_KC ? Bs + (wn0 + lr) * Cfg::KC_LD + lc : Bs + lc * (Cfg::BN + 4) + wn0 + lr;
#pragma unroll
    for (int kk = 0; kk < Cfg::BK / 4; kk++) {
        double a[Cfg::MI], b[Cfg::NI];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
            if (FULL || mi < mi_n)
                a[mi] = A_KC ? ap[mi * 8 * Cfg::KC_LD + kk * 4] : ap[kk * 4 * (Cfg::BM + 4) + mi * 8];
#pragma unroll
        for (int ni = 0; ni < Cfg::NI; ni++)
            if (FULL || ni < ni_n)
                b[ni] = B_KC ? bp[ni * 8 * Cfg::KC_LD + kk * 4] : bp[kk * 4 * (Cfg::BN + 4) + ni * 8];
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; mi++)
            if (FULL || mi < mi_n) {
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ni++)
                    if (FULL || ni < ni_n)
                        dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
    }
}

} // namespace b2g
