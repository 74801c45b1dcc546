// b2g_internal.h — internal types of libb2g.so (not part of the C ABI).
#pragma once
#include "../../include/b2g.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <functional>
#include <string>
#include <utility>
#include <vector>

// One GEMM pair of the H.C replay list, device layout (88 bytes).
//   W[m0 x n0]        = alpha0 * op(c + a0_off)[m0 x k0] * op(b0)[k0 x n0]
//   v + c1_off [m1 x n0] += alpha1 * op(a1)[m1 x m0] * W          (k1 == m0, n1 == n0)
struct B2GPair {
    const double *b0; // operator block of GEMM 0 (device)
    const double *a1; // operator block of GEMM 1 (device)
    double alpha0, alpha1;
    int64_t a0_off, c1_off; // offsets into c / sigma (|psi| < 2^31 for H.C, effective_hamiltonian.hpp:413;
                            // 64-bit because the rotation lists address whole operator arenas)
    int32_t m0, n0, k0, m1;
    int32_t lda0, ldb0, lda1, ldc1;
    uint32_t flags; // bit0 ta0, bit1 tb0, bit2 ta1
    uint32_t pad;
};
static_assert(sizeof(B2GPair) == 88, "B2GPair layout");

#define B2G_F_TA0 1u
#define B2G_F_TB0 2u
#define B2G_F_TA1 4u

// host address ranges mirrored into one device allocation (b2g_core.cu)
struct B2GRange {
    uintptr_t lo, hi; // host bytes [lo, hi)
    size_t dev_off;   // doubles from the device base
};
// Device-resident operands (b2g_resident_map): host byte range [lo, hi) currently lives at dev.
struct B2GMapEntry {
    uintptr_t lo, hi;
    double *dev;
};

struct b2g_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int64_t launches = 0;
    // pinned staging for the host-buffer entry points
    double *h_stage = nullptr;
    size_t h_stage_doubles = 0;
    double *d_c = nullptr, *d_v = nullptr;
    size_t d_cv_doubles = 0;
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // double-buffered pinned staging of the operand mirror (host pageable -> HBM)
    // side streams: the per-configuration launches of one phase run concurrently (small lists
    // do not fill the chip with any single launch)
    static constexpr int N_SIDE = 8;
    cudaStream_t side[N_SIDE] = {};
    cudaEvent_t side_done[N_SIDE] = {};
    cudaEvent_t fork_ev = nullptr;
    std::vector<B2GMapEntry> rmap; // sorted by lo, disjoint (b2g_resident_map)
    int64_t resident_hits = 0, resident_hit_bytes = 0, mirrored_bytes = 0;
    bool blocking_attr_set = false; // dynamic shared memory limits of the blocking kernels raised on this device
    void *eig_pool = nullptr; // b2g_eig.cu: streams / cuSOLVER handles of b2g_syevd (created at first use)
    void *h_up[2] = {nullptr, nullptr};
    cudaEvent_t up_done[2] = {nullptr, nullptr};
    size_t up_bytes = 0;
    int up_threads = 8;
};

struct b2g_plan {
    b2g_context *ctx = nullptr;
    int64_t npairs = 0, csize = 0, vsize = 0, max_work = 0;
    b2g_plan_stats stats{};
    double *d_operands = nullptr; // mirrored operator arenas (operand_space == HOST)
    B2GPair *d_pairs = nullptr;   // sorted by kernel class, then by output window
    int64_t n_generic = 0;        // pairs [0, n_generic) run through the generic kernel
    double *d_work = nullptr;     // spill space for W of pairs too large for shared memory
    size_t work_doubles = 0;
    std::vector<B2GPair> h_pairs; // host copy (debug / stats)
    void *tiled = nullptr;        // b2g::TiledPlan (two-phase DMMA path); null -> generic kernel only
};

void b2g_set_error(const std::string &msg);
void b2g_eig_destroy(b2g_context *ctx);
// host-side wall-clock profile (B2G_PROF): RAII section that adds to a label
struct B2GProfScope {
    const char *label;
    bool on;
    double t0;
    static double now();
    explicit B2GProfScope(const char *label) : label(label), on(b2g_prof_enabled() != 0), t0(on ? now() : 0.0) {}
    ~B2GProfScope() {
        if (on)
            b2g_prof_record(label, now() - t0);
    }
};
#define B2G_PROF_CAT2(a, b) a##b
#define B2G_PROF_CAT(a, b) B2G_PROF_CAT2(a, b)
#define B2G_PROF_SCOPE(label) B2GProfScope B2G_PROF_CAT(b2g_prof_scope_, __LINE__)(label)
// stream-ordered pool allocations (cached by the driver mempool across plans / Davidson calls)
int b2g_dmalloc(b2g_context *ctx, void **ptr, size_t bytes);
void b2g_dfree(b2g_context *ctx, void *ptr);
// pageable host -> device copy through pinned staging filled by several host threads
int b2g_upload(b2g_context *ctx, void *dst, const void *src, size_t bytes);
// fn(lo, hi) over [0, n) in at most nt contiguous chunks on host threads
void b2g_parallel_chunks(size_t n, int nt, const std::function<void(size_t, size_t)> &fn);
#define B2G_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            b2g_set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));          \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

// sort + merge touching ranges, assign 16-byte aligned device offsets; total = doubles needed
void b2g_merge_ranges(std::vector<B2GRange> &rg, size_t &total);
// range that holds ptr (rg merged and sorted)
const B2GRange &b2g_locate_range(const std::vector<B2GRange> &rg, uintptr_t ptr);
inline double *b2g_translate(const std::vector<B2GRange> &rg, double *dev_base, const void *host) {
    const B2GRange &r = b2g_locate_range(rg, (uintptr_t)host);
    return dev_base + r.dev_off + ((uintptr_t)host - r.lo) / sizeof(double);
}
// host ranges -> device (pinned staging, small neighbours packed into one DMA); asynchronous
int b2g_mirror_ranges(b2g_context *ctx, const std::vector<B2GRange> &rg, double *dev_base);
// device address of the host range [ptr, ptr + bytes) if it lies inside one mapped range, else nullptr
double *b2g_map_lookup(b2g_context *ctx, const void *ptr, size_t bytes);
// device -> host ranges through pinned staging; add = true: host += device, else host = device. Synchronous.
int b2g_download_ranges(b2g_context *ctx, const std::vector<B2GRange> &rg, const double *dev_base, bool add);

// kernels (b2g_kernels.cu)
int b2g_launch_matvec(b2g_plan *plan, const double *c_dev, double *v_dev, double scale);
// two-phase DMMA path (b2g_tiled.cu)
int b2g_tiled_build(b2g_plan *plan);
int b2g_tiled_launch(b2g_plan *plan, const double *c_dev, double *v_dev, double scale,
                     b2g_kernel_stat *stats = nullptr, int cap = 0, int *count = nullptr);
void b2g_tiled_destroy(void *tiled);
