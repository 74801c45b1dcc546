// b2g_core.cu — context, plan construction and the host entry points of libb2g.so.
//
// The plan is the device-side form of what EffectiveHamiltonian::precompute()
// records into BatchGEMMSeq (block2 src/dmrg/effective_hamiltonian.hpp:226-246,
// src/core/batch_gemm.hpp:893-902, 952-1022): a flat list of GEMM pairs whose
// wavefunction operands are null-based offsets and whose operator operands are
// raw pointers.  Nothing of the reference is compiled into this library.
#include "b2g_internal.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

static thread_local std::string g_err;
void b2g_set_error(const std::string &msg) { g_err = msg; }

extern "C" const char *b2g_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ host profile (B2G_PROF)
namespace {
struct ProfEntry {
    double seconds = 0;
    int64_t calls = 0;
};
std::mutex g_prof_mutex;
std::map<std::string, ProfEntry> g_prof;
} // namespace
double B2GProfScope::now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
extern "C" int b2g_prof_enabled(void) {
    static const int on = getenv("B2G_PROF") != nullptr ? 1 : 0;
    return on;
}
extern "C" void b2g_prof_record(const char *label, double seconds) {
    if (!b2g_prof_enabled() || !label)
        return;
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    ProfEntry &e = g_prof[label];
    e.seconds += seconds, e.calls++;
}
extern "C" int b2g_prof_dump(const char *path) {
    if (!b2g_prof_enabled())
        return 0;
    FILE *f = (path == nullptr || strcmp(path, "-") == 0) ? stderr : fopen(path, "w");
    if (!f) {
        b2g_set_error(std::string("b2g_prof_dump: cannot open ") + path);
        return 1;
    }
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    fprintf(f, "{");
    bool first = true;
    for (auto &kv : g_prof) {
        fprintf(f, "%s\n \"%s\": {\"seconds\": %.6f, \"calls\": %lld}", first ? "" : ",", kv.first.c_str(),
                kv.second.seconds, (long long)kv.second.calls);
        first = false;
    }
    fprintf(f, "\n}\n");
    if (f != stderr)
        fclose(f);
    return 0;
}

extern "C" int b2g_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

extern "C" int b2g_context_create(int device, b2g_context **out) {
    if (!out) {
        b2g_set_error("b2g_context_create: null out");
        return 1;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        b2g_set_error(std::string("b2g_context_create: no CUDA device (") +
                      (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                      "); libb2g has no CPU fallback");
        return 2;
    }
    B2G_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2G_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        b2g_set_error("b2g_context_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                      ", libb2g is built for sm_100a only");
        return 3;
    }
    b2g_context *ctx = new b2g_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    B2G_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    // keep freed blocks in the stream-ordered pool: one plan per site and one Davidson call per
    // site would otherwise pay cudaMalloc/cudaFree every time
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        // Freed blocks stay in the pool (one plan and one Davidson call per site would otherwise pay the driver's
        // map / unmap every time: measured 5 s of a 37 s Cr2 M=1000 run with a 32 GB threshold).  The price is
        // fragmentation at large M (Cr2 M=4000: 132 GB reserved for ~100 GB in use); b2g_dmalloc hands the unused
        // part back and retries when an allocation fails.  B2G_POOL_KEEP_GB sets a finite threshold.
        const char *keep = getenv("B2G_POOL_KEEP_GB");
        uint64_t thr = keep ? (uint64_t)(atof(keep) * 1e9) : UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ctx->up_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    int prio_lo = 0, prio_hi = 0; // side streams at the lowest priority: their CTAs fill the tails of the
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi); // launches on the context stream
    for (int i = 0; i < b2g_context::N_SIDE; i++) {
        B2G_CUDA(cudaStreamCreateWithPriority(&ctx->side[i], cudaStreamNonBlocking, prio_lo));
        B2G_CUDA(cudaEventCreateWithFlags(&ctx->side_done[i], cudaEventDisableTiming));
    }
    B2G_CUDA(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    *out = ctx;
    return 0;
}

static void trim_pool(b2g_context *ctx) {
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess)
        cudaMemPoolTrimTo(pool, 0);
}
int b2g_dmalloc(b2g_context *ctx, void **ptr, size_t bytes) {
    const size_t want = std::max<size_t>(bytes, 16);
    cudaError_t e = cudaMallocAsync(ptr, want, ctx->stream);
    if (e == cudaErrorMemoryAllocation) { // unused blocks of the pool that do not fit: give them back, try again
        cudaGetLastError();
        trim_pool(ctx);
        e = cudaMallocAsync(ptr, want, ctx->stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        size_t f = 0, t = 0;
        cudaMemGetInfo(&f, &t);
        b2g_set_error("device allocation of " + std::to_string(want) + " bytes failed (" + cudaGetErrorString(e) +
                      "; out of memory: " + std::to_string(f >> 20) + " MiB free of " + std::to_string(t >> 20) + ")");
        *ptr = nullptr;
        return 1;
    }
    return 0;
}
extern "C" int b2g_mem_trim(b2g_context *ctx) {
    if (!ctx) {
        b2g_set_error("b2g_mem_trim: null context");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    trim_pool(ctx);
    return 0;
}
void b2g_dfree(b2g_context *ctx, void *ptr) {
    if (ptr)
        cudaFreeAsync(ptr, ctx->stream);
}

static constexpr size_t B2G_UP_CHUNK = (size_t)32 << 20;

static int ensure_upload_buffers(b2g_context *ctx) {
    if (!ctx->h_up[0]) {
        for (int i = 0; i < 2; i++) {
            B2G_CUDA(cudaMallocHost(&ctx->h_up[i], B2G_UP_CHUNK));
            B2G_CUDA(cudaEventCreateWithFlags(&ctx->up_done[i], cudaEventDisableTiming));
        }
        ctx->up_bytes = B2G_UP_CHUNK;
    }
    return 0;
}

// Packs many small host ranges that are neighbours in the device arena into one pinned
// staging slice and ships the slice with a single DMA (a site has thousands of operator
// blocks; one cudaMemcpyAsync per block from pageable memory costs ~20 us each).
struct MirrorWriter {
    b2g_context *ctx;
    char *dst_base;
    size_t slice_begin = 0, fill = 0;
    int buf = 0;
    struct Piece {
        size_t off; // inside the slice
        const void *src;
        size_t bytes;
    };
    std::vector<Piece> pieces;
    // copy the recorded pieces into the staging buffer (several host threads for a large slice), one DMA
    int flush() {
        if (fill) {
            B2G_CUDA(cudaEventSynchronize(ctx->up_done[buf])); // last DMA out of this buffer has finished
            char *stage = (char *)ctx->h_up[buf];
            const int nt = fill > ((size_t)4 << 20) ? std::min<int>(ctx->up_threads, (int)pieces.size()) : 1;
            if (nt <= 1) {
                for (const Piece &pc : pieces)
                    memcpy(stage + pc.off, pc.src, pc.bytes);
            } else { // contiguous groups of pieces with about equal byte counts
                std::vector<std::thread> th;
                const size_t per = (fill + nt - 1) / nt;
                size_t i = 0;
                for (int t = 0; t < nt && i < pieces.size(); t++) {
                    size_t j = i, acc = 0;
                    while (j < pieces.size() && (acc < per || t == nt - 1))
                        acc += pieces[j++].bytes;
                    const Piece *pp = pieces.data();
                    auto work = [stage, pp, i, j]() {
                        for (size_t q = i; q < j; q++)
                            memcpy(stage + pp[q].off, pp[q].src, pp[q].bytes);
                    };
                    if (j < pieces.size())
                        th.emplace_back(work);
                    else
                        work();
                    i = j;
                }
                for (auto &x : th)
                    x.join();
            }
            B2G_CUDA(cudaMemcpyAsync(dst_base + slice_begin, stage, fill, cudaMemcpyHostToDevice, ctx->stream));
            B2G_CUDA(cudaEventRecord(ctx->up_done[buf], ctx->stream));
            buf ^= 1;
            fill = 0;
            pieces.clear();
        }
        return 0;
    }
    int add(size_t dev_off, const void *src, size_t bytes) {
        if (bytes >= B2G_UP_CHUNK / 4) {
            if (flush())
                return 1;
            return b2g_upload(ctx, dst_base + dev_off, src, bytes);
        }
        if (fill && (dev_off < slice_begin || dev_off - slice_begin + bytes > B2G_UP_CHUNK))
            if (flush())
                return 1;
        if (!fill)
            slice_begin = dev_off;
        pieces.push_back(Piece{dev_off - slice_begin, src, bytes});
        fill = dev_off - slice_begin + bytes;
        return 0;
    }
};

// staging slice per host thread: 4 KB aligned, nt * slice >= len
static inline size_t upload_slice(size_t len, int nt) {
    return ((len + (size_t)nt - 1) / (size_t)nt + 4095) & ~(size_t)4095;
}
// test hook (no device): the [lo, hi) byte ranges the nt staging threads of b2g_upload copy for a chunk of len bytes
extern "C" int b2g_debug_upload_slices(int64_t len, int nt, int64_t *lo, int64_t *hi) {
    if (len < 0 || nt < 1 || !lo || !hi)
        return 1;
    const size_t slice = upload_slice((size_t)len, nt);
    for (int t = 0; t < nt; t++)
        lo[t] = (int64_t)std::min((size_t)len, slice * t), hi[t] = (int64_t)std::min((size_t)len, slice * (t + 1));
    return 0;
}

int b2g_upload(b2g_context *ctx, void *dst, const void *src, size_t bytes) {
    constexpr size_t CH = B2G_UP_CHUNK;
    if (ensure_upload_buffers(ctx))
        return 1;
    B2G_CUDA(cudaEventSynchronize(ctx->up_done[0]));
    B2G_CUDA(cudaEventSynchronize(ctx->up_done[1]));
    int buf = 0;
    for (size_t off = 0; off < bytes; off += CH, buf ^= 1) {
        const size_t len = std::min(CH, bytes - off);
        B2G_CUDA(cudaEventSynchronize(ctx->up_done[buf])); // previous DMA out of this buffer finished
        const int nt = ctx->up_threads;
        const size_t slice = upload_slice(len, nt);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) {
            const size_t lo = std::min(len, slice * t), hi = std::min(len, slice * (t + 1));
            if (hi > lo)
                th.emplace_back([=]() { memcpy((char *)ctx->h_up[buf] + lo, (const char *)src + off + lo, hi - lo); });
        }
        memcpy(ctx->h_up[buf], (const char *)src + off, std::min(len, slice));
        for (auto &x : th)
            x.join();
        B2G_CUDA(cudaMemcpyAsync((char *)dst + off, ctx->h_up[buf], len, cudaMemcpyHostToDevice, ctx->stream));
        B2G_CUDA(cudaEventRecord(ctx->up_done[buf], ctx->stream));
    }
    return 0;
}

extern "C" int b2g_context_destroy(b2g_context *ctx) {
    if (!ctx)
        return 0;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm)
        b2g_comm_destroy(ctx);
    b2g_eig_destroy(ctx);
    if (ctx->h_stage)
        cudaFreeHost(ctx->h_stage);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_up[i])
            cudaFreeHost(ctx->h_up[i]);
        if (ctx->up_done[i])
            cudaEventDestroy(ctx->up_done[i]);
    }
    if (ctx->d_c)
        cudaFree(ctx->d_c);
    if (ctx->d_v)
        cudaFree(ctx->d_v);
    for (int i = 0; i < b2g_context::N_SIDE; i++) {
        if (ctx->side[i])
            cudaStreamDestroy(ctx->side[i]);
        if (ctx->side_done[i])
            cudaEventDestroy(ctx->side_done[i]);
    }
    if (ctx->fork_ev)
        cudaEventDestroy(ctx->fork_ev);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

extern "C" int64_t b2g_context_launches(const b2g_context *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *b2g_context_stream(const b2g_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int b2g_context_synchronize(b2g_context *ctx) {
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// stream-ordered pool allocations (ordered on the context stream like every other piece of work)
extern "C" int b2g_malloc(b2g_context *ctx, size_t bytes, void **dev) {
    if (!ctx || !dev) {
        b2g_set_error("b2g_malloc: null argument");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    return b2g_dmalloc(ctx, dev, bytes);
}
extern "C" int b2g_free(b2g_context *ctx, void *dev) {
    if (!ctx) {
        b2g_set_error("b2g_free: null context");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    b2g_dfree(ctx, dev);
    return 0;
}
extern "C" int b2g_mem_info(b2g_context *ctx, int64_t *free_bytes, int64_t *total_bytes) {
    if (!ctx) {
        b2g_set_error("b2g_mem_info: null context");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    B2G_CUDA(cudaMemGetInfo(&f, &t));
    // memory parked in the stream-ordered pool is reusable by b2g_malloc
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
        uint64_t reserved = 0, used = 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        f += (size_t)(reserved - used);
    }
    if (free_bytes)
        *free_bytes = (int64_t)f;
    if (total_bytes)
        *total_bytes = (int64_t)t;
    return 0;
}
extern "C" int b2g_memcpy_h2d(b2g_context *ctx, void *dev, const void *host, size_t bytes) {
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int b2g_memcpy_d2h(b2g_context *ctx, void *host, const void *dev, size_t bytes) {
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int b2g_memset_zero(b2g_context *ctx, void *dev, size_t bytes) {
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaMemsetAsync(dev, 0, bytes, ctx->stream));
    return 0;
}

// ------------------------------------------------------------------ plan

static inline bool is_trans(int32_t t) { return t == B2G_TRANS || t == 1; }
static inline bool valid_trans(int32_t t) { return t == B2G_TRANS || t == B2G_NOTRANS || t == 0 || t == 1; }

// elements spanned by a row-major rows x cols matrix with leading dimension ld
static inline size_t extent(int rows, int cols, int ld) {
    return (rows <= 0 || cols <= 0) ? 0 : (size_t)(rows - 1) * (size_t)ld + (size_t)cols;
}

struct Range {
    uintptr_t lo, hi;
    size_t dev_off; // doubles
};

extern "C" int b2g_plan_create(b2g_context *ctx, const b2g_batch *b0, const b2g_batch *b1, int64_t max_work,
                               int64_t csize, int64_t vsize, int operand_space, b2g_plan **out) {
    if (!ctx || !b0 || !b1 || !out) {
        b2g_set_error("b2g_plan_create: null argument");
        return 1;
    }
    if (b0->count != b1->count) {
        b2g_set_error("b2g_plan_create: batch[0] and batch[1] differ in length (gp[i] must be 1, acidxs empty)");
        return 1;
    }
    if (csize >= ((int64_t)1 << 31) || vsize >= ((int64_t)1 << 31)) {
        b2g_set_error("b2g_plan_create: wavefunction larger than 2^31 doubles");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    const auto t_enter = std::chrono::steady_clock::now();
    const int64_t n = b0->count;
    const char *force = getenv("B2G_FORCE_GENERIC");
    const bool force_generic = force && force[0] == '1';
    b2g_plan *p = new b2g_plan();
    p->ctx = ctx, p->npairs = n, p->csize = csize, p->vsize = vsize, p->max_work = max_work;
    std::vector<B2GPair> &hp = p->h_pairs;
    hp.resize((size_t)n);
    std::vector<Range> rg;
    rg.reserve((size_t)2 * n);
    int64_t nflop = 0;
    size_t resident_doubles = 0; // per-pair extents read in place from resident blocks (with repeats)
    for (int64_t i = 0; i < n; i++) {
        if (!valid_trans(b0->ta[i]) || !valid_trans(b0->tb[i]) || !valid_trans(b1->ta[i]) ||
            !valid_trans(b1->tb[i])) {
            b2g_set_error("b2g_plan_create: pair " + std::to_string(i) + ": transpose flag is not N/T");
            delete p;
            return 1;
        }
        const bool ta0 = is_trans(b0->ta[i]), tb0 = is_trans(b0->tb[i]), ta1 = is_trans(b1->ta[i]),
                   tb1 = is_trans(b1->tb[i]);
        const int m0 = b0->m[i], n0 = b0->n[i], k0 = b0->k[i], m1 = b1->m[i], n1 = b1->n[i], k1 = b1->k[i];
        // the chained form recorded by AdvancedGEMM<real>::rotate / BatchGEMMSeq::three_rotate:
        // GEMM 1 consumes the contiguous work matrix of GEMM 0 as its untransposed B operand
        if ((const void *)b0->c[i] != (const void *)b1->b[i] || tb1 || k1 != m0 || n1 != n0 ||
            b0->ldc[i] != n0 || b1->ldb[i] != n0 || b0->beta[i] != 0.0 || b1->beta[i] != 1.0) {
            b2g_set_error("b2g_plan_create: pair " + std::to_string(i) +
                          " is not a chained W = A0*B0, C1 += A1*W pair");
            delete p;
            return 1;
        }
        // rotate / three_rotate never record a transposed wavefunction operand (batch_gemm.hpp:893-1022)
        // and the tile engine has no A-transposed phase 1: refuse instead of computing the wrong product
        if (ta0 && !force_generic) {
            b2g_set_error("b2g_plan_create: pair " + std::to_string(i) +
                          ": transposed wavefunction operand (ta of batch[0]) is not part of the H.C replay list");
            delete p;
            return 1;
        }
        const uintptr_t a0 = (uintptr_t)b0->a[i], c1 = (uintptr_t)b1->c[i];
        if (a0 % sizeof(double) || c1 % sizeof(double)) {
            b2g_set_error("b2g_plan_create: misaligned wavefunction offset");
            delete p;
            return 1;
        }
        const size_t a0_off = a0 / sizeof(double), c1_off = c1 / sizeof(double);
        const size_t ea = extent(ta0 ? k0 : m0, ta0 ? m0 : k0, b0->lda[i]);
        const size_t ec = extent(m1, n1, b1->ldc[i]);
        if (a0_off + ea > (size_t)csize || c1_off + ec > (size_t)vsize) {
            b2g_set_error("b2g_plan_create: pair " + std::to_string(i) +
                          ": wavefunction window outside [0, size) - operands must be recorded null-based");
            delete p;
            return 1;
        }
        B2GPair &q = hp[(size_t)i];
        q.b0 = b0->b[i], q.a1 = b1->a[i];
        q.alpha0 = b0->alpha[i], q.alpha1 = b1->alpha[i];
        q.a0_off = (int64_t)a0_off, q.c1_off = (int64_t)c1_off;
        q.m0 = m0, q.n0 = n0, q.k0 = k0, q.m1 = m1;
        q.lda0 = b0->lda[i], q.ldb0 = b0->ldb[i], q.lda1 = b1->lda[i], q.ldc1 = b1->ldc[i];
        q.flags = (ta0 ? B2G_F_TA0 : 0) | (tb0 ? B2G_F_TB0 : 0) | (ta1 ? B2G_F_TA1 : 0);
        q.pad = 0;
        nflop += (int64_t)m0 * n0 * k0 + (int64_t)m1 * n1 * k1;
        const size_t eb0 = extent(tb0 ? n0 : k0, tb0 ? k0 : n0, q.ldb0);
        const size_t ea1 = extent(ta1 ? k1 : m1, ta1 ? m1 : k1, q.lda1);
        const uintptr_t pb = (uintptr_t)q.b0, pa = (uintptr_t)q.a1;
        // device-resident blocks (b2g_resident_map) are read in place; the rest is mirrored below
        if (operand_space == B2G_OPERANDS_HOST && eb0)
            if (double *d = b2g_map_lookup(ctx, q.b0, eb0 * sizeof(double)))
                q.b0 = d, q.pad |= 1u, resident_doubles += eb0;
        if (operand_space == B2G_OPERANDS_HOST && ea1)
            if (double *d = b2g_map_lookup(ctx, q.a1, ea1 * sizeof(double)))
                q.a1 = d, q.pad |= 2u, resident_doubles += ea1;
        if (eb0 && !(q.pad & 1u))
            rg.push_back(Range{pb, pb + eb0 * sizeof(double), 0});
        if (ea1 && !(q.pad & 2u))
            rg.push_back(Range{pa, pa + ea1 * sizeof(double), 0});
    }
    ctx->resident_hit_bytes += (int64_t)(resident_doubles * sizeof(double));
    const bool verbose = getenv("B2G_VERBOSE") != nullptr;
    auto tstart = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (verbose || b2g_prof_enabled()) {
            auto now = std::chrono::steady_clock::now();
            if (verbose)
                fprintf(stderr, "[b2g] plan_create %-14s %8.3f ms\n", what,
                        std::chrono::duration<double, std::milli>(now - tstart).count());
            b2g_prof_record((std::string("plan_create.") + what).c_str(),
                            std::chrono::duration<double>(now - tstart).count());
            tstart = now;
        }
    };
    b2g_prof_record("plan_create.scan pairs",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enter).count());
    tstart = std::chrono::steady_clock::now();
    // merge the referenced operator ranges into arenas
    std::sort(rg.begin(), rg.end(), [](const Range &x, const Range &y) { return x.lo < y.lo; });
    std::vector<Range> ar;
    for (const Range &r : rg) {
        if (!ar.empty() && r.lo <= ar.back().hi)
            ar.back().hi = std::max(ar.back().hi, r.hi);
        else
            ar.push_back(r);
    }
    size_t total = 0;
    for (Range &r : ar) {
        r.dev_off = total;
        total += (r.hi - r.lo) / sizeof(double);
        total = (total + 1) & ~(size_t)1; // keep every arena 16-byte aligned relative to its host alignment
    }
    p->stats.pairs = n, p->stats.csize = csize, p->stats.vsize = vsize, p->stats.nflop_mnk = nflop;
    p->stats.arenas = (int64_t)ar.size();
    size_t op_doubles = 0;
    for (const Range &r : ar)
        op_doubles += (r.hi - r.lo) / sizeof(double);
    p->stats.operand_doubles = (int64_t)op_doubles;
    p->stats.mirrored_doubles = operand_space == B2G_OPERANDS_HOST ? (int64_t)total : 0;

    auto t0 = std::chrono::steady_clock::now();
    if (operand_space == B2G_OPERANDS_HOST && total > 0) {
        if (b2g_dmalloc(ctx, (void **)&p->d_operands, total * sizeof(double)) != 0) {
            b2g_set_error("b2g_plan_create: cudaMalloc of " + std::to_string(total * 8) + " operand bytes failed");
            b2g_plan_destroy(p);
            return 1;
        }
        {
            std::vector<B2GRange> brg(ar.size());
            for (size_t i = 0; i < ar.size(); i++)
                brg[i] = B2GRange{ar[i].lo, ar[i].hi, ar[i].dev_off};
            if (b2g_mirror_ranges(ctx, brg, p->d_operands)) {
                b2g_plan_destroy(p);
                return 1;
            }
        }
        auto locate = [&ar](uintptr_t ptr) -> const Range & {
            size_t lo = 0, hi = ar.size();
            while (hi - lo > 1) {
                size_t mid = (lo + hi) / 2;
                if (ar[mid].lo <= ptr)
                    lo = mid;
                else
                    hi = mid;
            }
            return ar[lo];
        };
        for (B2GPair &q : hp) {
            if (!(q.pad & 1u)) {
                const Range &rb = locate((uintptr_t)q.b0);
                q.b0 = p->d_operands + rb.dev_off + ((uintptr_t)q.b0 - rb.lo) / sizeof(double);
            }
            if (!(q.pad & 2u)) {
                const Range &ra = locate((uintptr_t)q.a1);
                q.a1 = p->d_operands + ra.dev_off + ((uintptr_t)q.a1 - ra.lo) / sizeof(double);
            }
        }
    }
    for (B2GPair &q : hp)
        q.pad = 0;
    lap("mirror issue");
    // order: pairs writing the same sigma window become neighbours (locality of the accumulation)
    std::stable_sort(hp.begin(), hp.end(), [](const B2GPair &x, const B2GPair &y) {
        if (x.c1_off != y.c1_off)
            return x.c1_off < y.c1_off;
        return x.a0_off < y.a0_off;
    });
    if (n > 0) {
        if (b2g_dmalloc(ctx, (void **)&p->d_pairs, (size_t)n * sizeof(B2GPair)) ||
            cudaMemcpyAsync(p->d_pairs, hp.data(), (size_t)n * sizeof(B2GPair), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess) {
            b2g_set_error("b2g_plan_create: pair list upload failed");
            b2g_plan_destroy(p);
            return 1;
        }
    }
    lap("sort pairs");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        b2g_set_error("b2g_plan_create: operand mirror failed");
        b2g_plan_destroy(p);
        return 1;
    }
    lap("mirror sync");
    p->stats.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    p->n_generic = n;
    p->stats.n_small = n, p->stats.n_large = 0;
    p->stats.launches = n > 0 ? 1 : 0;
    // default route: two-phase DMMA tile engine; B2G_FORCE_GENERIC=1 keeps the one-CTA-per-pair
    // kernel (the correctness anchor the tiled path is tested against)
    if (!force_generic && n > 0) {
        if (b2g_tiled_build(p)) {
            b2g_plan_destroy(p);
            return 1;
        }
        lap("tiled build");
    }
    *out = p;
    return 0;
}

// Host-side regrouping of a chained pair list into the two-phase tile plan, without a device (timing and
// tests of the planner): seconds = wall time of the regrouping, units = CTA work units, launches = kernel launches.
extern "C" int b2g_debug_tiled_plan(const b2g_batch *b0, const b2g_batch *b1, double *seconds, int64_t *units,
                                    int64_t *launches, int64_t *fingerprint) {
    if (!b0 || !b1 || b0->count != b1->count) {
        b2g_set_error("b2g_debug_tiled_plan: bad argument");
        return 1;
    }
    b2g_plan *p = new b2g_plan();
    p->ctx = nullptr, p->npairs = b0->count;
    p->h_pairs.resize((size_t)b0->count);
    for (int64_t i = 0; i < b0->count; i++) {
        B2GPair &q = p->h_pairs[(size_t)i];
        q.b0 = b0->b[i], q.a1 = b1->a[i];
        q.alpha0 = b0->alpha[i], q.alpha1 = b1->alpha[i];
        q.a0_off = (int64_t)((uintptr_t)b0->a[i] / sizeof(double)), q.c1_off = (int64_t)((uintptr_t)b1->c[i] / sizeof(double));
        q.m0 = b0->m[i], q.n0 = b0->n[i], q.k0 = b0->k[i], q.m1 = b1->m[i];
        q.lda0 = b0->lda[i], q.ldb0 = b0->ldb[i], q.lda1 = b1->lda[i], q.ldc1 = b1->ldc[i];
        q.flags = (is_trans(b0->tb[i]) ? B2G_F_TB0 : 0) | (is_trans(b1->ta[i]) ? B2G_F_TA1 : 0);
        q.pad = 0;
    }
    std::stable_sort(p->h_pairs.begin(), p->h_pairs.end(), [](const B2GPair &x, const B2GPair &y) {
        if (x.c1_off != y.c1_off)
            return x.c1_off < y.c1_off;
        return x.a0_off < y.a0_off;
    });
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = b2g_tiled_build(p);
    if (seconds)
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (units)
        *units = p->stats.n_large;
    if (launches)
        *launches = p->stats.launches;
    if (fingerprint)
        *fingerprint = p->stats.arenas; // the host-only build leaves the fingerprint of the plan here
    b2g_tiled_destroy(p->tiled);
    delete p;
    return rc;
}

extern "C" int b2g_plan_destroy(b2g_plan *p) {
    if (!p)
        return 0;
    B2G_PROF_SCOPE("plan_destroy");
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    b2g_dfree(p->ctx, p->d_operands);
    b2g_dfree(p->ctx, p->d_pairs);
    b2g_dfree(p->ctx, p->d_work);
    b2g_tiled_destroy(p->tiled);
    delete p;
    return 0;
}

extern "C" int b2g_plan_get_stats(const b2g_plan *p, b2g_plan_stats *out) {
    if (!p || !out) {
        b2g_set_error("b2g_plan_get_stats: null argument");
        return 1;
    }
    *out = p->stats;
    return 0;
}

extern "C" int b2g_seq_matvec_dev(b2g_plan *p, const double *c_dev, double *v_dev, double scale) {
    if (!p) {
        b2g_set_error("b2g_seq_matvec_dev: null plan");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(p->ctx->device));
    return b2g_launch_matvec(p, c_dev, v_dev, scale);
}

extern "C" int b2g_plan_profile(b2g_plan *p, const double *c_dev, double *v_dev, double scale,
                                b2g_kernel_stat *out, int capacity, int *count) {
    if (!p || !out || !count) {
        b2g_set_error("b2g_plan_profile: null argument");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(p->ctx->device));
    if (!p->tiled) {
        b2g_set_error("b2g_plan_profile: plan runs the generic kernel only");
        return 1;
    }
    return b2g_tiled_launch(p, c_dev, v_dev, scale, out, capacity, count);
}

// fn(lo, hi) over [0, n) cut into at most nt contiguous chunks, one host thread each (the caller runs chunk 0)
void b2g_parallel_chunks(size_t n, int nt, const std::function<void(size_t, size_t)> &fn) {
    nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nt, 1), n / ((size_t)1 << 15)));
    const size_t slice = (n + (size_t)nt - 1) / (size_t)nt;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) {
        const size_t lo = std::min(n, slice * t), hi = std::min(n, slice * (t + 1));
        if (hi > lo)
            th.emplace_back(fn, lo, hi);
    }
    if (n)
        fn(0, std::min(n, slice));
    for (auto &x : th)
        x.join();
}

static int ensure_staging(b2g_context *ctx, size_t csize, size_t vsize) {
    const size_t need = std::max(csize, vsize);
    if (ctx->h_stage_doubles < need) {
        if (ctx->h_stage)
            cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr, ctx->h_stage_doubles = 0;
        B2G_CUDA(cudaMallocHost(&ctx->h_stage, need * sizeof(double)));
        ctx->h_stage_doubles = need;
    }
    if (ctx->d_cv_doubles < need) {
        if (ctx->d_c)
            cudaFree(ctx->d_c);
        if (ctx->d_v)
            cudaFree(ctx->d_v);
        ctx->d_c = ctx->d_v = nullptr, ctx->d_cv_doubles = 0;
        B2G_CUDA(cudaMalloc(&ctx->d_c, need * sizeof(double)));
        B2G_CUDA(cudaMalloc(&ctx->d_v, need * sizeof(double)));
        ctx->d_cv_doubles = need;
    }
    return 0;
}

// Host-buffer drop-in for BatchGEMMSeq::operator()(c, v, scale): the reference accumulates
// into v (beta = 1 on every second GEMM), so the device result is added to the caller's v.
extern "C" int b2g_seq_matvec(b2g_plan *p, const double *c_host, double *v_host, double scale) {
    if (!p || !c_host || !v_host) {
        b2g_set_error("b2g_seq_matvec: null argument");
        return 1;
    }
    b2g_context *ctx = p->ctx;
    B2G_CUDA(cudaSetDevice(ctx->device));
    if (ensure_staging(ctx, (size_t)p->csize, (size_t)p->vsize))
        return 1;
    // pageable c -> pinned staging and sigma += staging, both cut over the upload threads (19 MB each at
    // M = 4000: 8 ms of an 88 ms step on one host thread)
    b2g_parallel_chunks((size_t)p->csize, ctx->up_threads, [&](size_t lo, size_t hi) {
        memcpy(ctx->h_stage + lo, c_host + lo, (hi - lo) * sizeof(double));
    });
    B2G_CUDA(cudaMemcpyAsync(ctx->d_c, ctx->h_stage, (size_t)p->csize * sizeof(double), cudaMemcpyHostToDevice,
                             ctx->stream));
    B2G_CUDA(cudaMemsetAsync(ctx->d_v, 0, (size_t)p->vsize * sizeof(double), ctx->stream));
    if (b2g_launch_matvec(p, ctx->d_c, ctx->d_v, scale))
        return 1;
    if (ctx->nccl_comm && b2g_allreduce_sum(ctx, ctx->d_v, p->vsize))
        return 1;
    B2G_CUDA(cudaMemcpyAsync(ctx->h_stage, ctx->d_v, (size_t)p->vsize * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    const double *hs = ctx->h_stage;
    b2g_parallel_chunks((size_t)p->vsize, ctx->up_threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++)
            v_host[i] += hs[i];
    });
    return 0;
}

// ------------------------------------------------------------------ rotation lists

extern "C" int b2g_pairs_execute(b2g_context *ctx, const b2g_batch *b0, const b2g_batch *b1, int64_t max_work,
                                 b2g_plan_stats *stats) {
    if (!ctx || !b0 || !b1) {
        b2g_set_error("b2g_pairs_execute: null argument");
        return 1;
    }
    if (b0->count != b1->count) {
        b2g_set_error("b2g_pairs_execute: batch[0] and batch[1] differ in length");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = b0->count;
    if (n == 0)
        return 0;
    auto t0 = std::chrono::steady_clock::now();
    double prof_t = B2GProfScope::now();
    auto prof_lap = [&prof_t](const char *label) {
        if (b2g_prof_enabled()) {
            const double now = B2GProfScope::now();
            b2g_prof_record(label, now - prof_t);
            prof_t = now;
        }
    };
    b2g_plan *p = new b2g_plan();
    p->ctx = ctx, p->npairs = n, p->max_work = max_work;
    std::vector<B2GPair> &hp = p->h_pairs;
    hp.resize((size_t)n);
    std::vector<Range> in_rg, out_rg;
    in_rg.reserve((size_t)3 * n), out_rg.reserve((size_t)n);
    int64_t nflop = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!valid_trans(b0->ta[i]) || !valid_trans(b0->tb[i]) || !valid_trans(b1->ta[i]) ||
            !valid_trans(b1->tb[i])) {
            b2g_set_error("b2g_pairs_execute: transpose flag is not N/T");
            delete p;
            return 1;
        }
        const bool ta0 = is_trans(b0->ta[i]), tb0 = is_trans(b0->tb[i]), ta1 = is_trans(b1->ta[i]),
                   tb1 = is_trans(b1->tb[i]);
        const int m0 = b0->m[i], n0 = b0->n[i], k0 = b0->k[i], m1 = b1->m[i], n1 = b1->n[i], k1 = b1->k[i];
        if ((const void *)b0->c[i] != (const void *)b1->b[i] || ta0 || tb1 || k1 != m0 || n1 != n0 ||
            b0->ldc[i] != n0 || b1->ldb[i] != n0 || b0->beta[i] != 0.0 || b1->beta[i] != 1.0) {
            b2g_set_error("b2g_pairs_execute: pair " + std::to_string(i) +
                          " is not a chained W = A0*B0, C1 += A1*W pair");
            delete p;
            return 1;
        }
        B2GPair &q = hp[(size_t)i];
        q.b0 = b0->b[i], q.a1 = b1->a[i];
        q.alpha0 = b0->alpha[i], q.alpha1 = b1->alpha[i];
        q.a0_off = (int64_t)(uintptr_t)b0->a[i], q.c1_off = (int64_t)(uintptr_t)b1->c[i]; // host addresses for now
        q.m0 = m0, q.n0 = n0, q.k0 = k0, q.m1 = m1;
        q.lda0 = b0->lda[i], q.ldb0 = b0->ldb[i], q.lda1 = b1->lda[i], q.ldc1 = b1->ldc[i];
        q.flags = (tb0 ? B2G_F_TB0 : 0) | (ta1 ? B2G_F_TA1 : 0);
        q.pad = 0;
        nflop += (int64_t)m0 * n0 * k0 + (int64_t)m1 * n1 * k1;
        auto add = [](std::vector<Range> &v, const void *ptr, size_t ext) {
            if (ext)
                v.push_back(Range{(uintptr_t)ptr, (uintptr_t)ptr + ext * sizeof(double), 0});
        };
        // device-resident blocks (b2g_resident_map): inputs are read in place, outputs written in place and
        // not copied back; q.pad remembers which operands already are device addresses
        const size_t e_a0 = extent(m0, k0, q.lda0), e_b0 = extent(tb0 ? n0 : k0, tb0 ? k0 : n0, q.ldb0),
                     e_a1 = extent(ta1 ? k1 : m1, ta1 ? m1 : k1, q.lda1), e_c1 = extent(m1, n1, q.ldc1);
        if (double *d = e_a0 ? b2g_map_lookup(ctx, b0->a[i], e_a0 * 8) : nullptr)
            q.a0_off = (int64_t)(uintptr_t)d, q.pad |= 4u, ctx->resident_hit_bytes += (int64_t)e_a0 * 8;
        else
            add(in_rg, b0->a[i], e_a0);
        if (double *d = e_b0 ? b2g_map_lookup(ctx, q.b0, e_b0 * 8) : nullptr)
            q.b0 = d, q.pad |= 1u;
        else
            add(in_rg, q.b0, e_b0);
        if (double *d = e_a1 ? b2g_map_lookup(ctx, q.a1, e_a1 * 8) : nullptr)
            q.a1 = d, q.pad |= 2u;
        else
            add(in_rg, q.a1, e_a1);
        if (double *d = e_c1 ? b2g_map_lookup(ctx, b1->c[i], e_c1 * 8) : nullptr)
            q.c1_off = (int64_t)(uintptr_t)d, q.pad |= 8u;
        else
            add(out_rg, b1->c[i], e_c1);
    }
    auto merge = [](std::vector<Range> &rg, size_t &total) {
        std::sort(rg.begin(), rg.end(), [](const Range &x, const Range &y) { return x.lo < y.lo; });
        std::vector<Range> ar;
        for (const Range &r : rg) {
            if (!ar.empty() && r.lo <= ar.back().hi)
                ar.back().hi = std::max(ar.back().hi, r.hi);
            else
                ar.push_back(r);
        }
        total = 0;
        for (Range &r : ar) {
            r.dev_off = total;
            total += (r.hi - r.lo) / sizeof(double);
            total = (total + 1) & ~(size_t)1;
        }
        rg.swap(ar);
    };
    prof_lap("pairs.scan");
    size_t in_total = 0, out_total = 0;
    merge(in_rg, in_total), merge(out_rg, out_total);
    auto locate = [](const std::vector<Range> &ar, uintptr_t ptr) -> const Range & {
        size_t lo = 0, hi = ar.size();
        while (hi - lo > 1) {
            size_t mid = (lo + hi) / 2;
            if (ar[mid].lo <= ptr)
                lo = mid;
            else
                hi = mid;
        }
        return ar[lo];
    };
    // an output block must not also be an input of the same list
    if (!in_rg.empty())
        for (const Range &o : out_rg) {
            const Range &r = locate(in_rg, o.lo);
            const Range *nx = (&r + 1 < in_rg.data() + in_rg.size()) ? &r + 1 : nullptr;
            if ((r.lo <= o.lo && o.lo < r.hi) || (o.lo <= r.lo && r.lo < o.hi) ||
                (nx && nx->lo < o.hi && nx->lo >= o.lo)) {
                b2g_set_error("b2g_pairs_execute: an output block aliases an input block");
                delete p;
                return 1;
            }
        }
    double *d_in = nullptr, *d_out = nullptr;
    int rc = 0;
    auto fail = [&](const std::string &msg) {
        if (!msg.empty())
            b2g_set_error(msg);
        b2g_dfree(ctx, d_in), b2g_dfree(ctx, d_out);
        b2g_plan_destroy(p);
        return 1;
    };
    if (b2g_dmalloc(ctx, (void **)&d_in, in_total * sizeof(double)) ||
        b2g_dmalloc(ctx, (void **)&d_out, out_total * sizeof(double)))
        return fail("");
    if (ensure_upload_buffers(ctx))
        return fail("");
    if (out_total && cudaMemsetAsync(d_out, 0, out_total * sizeof(double), ctx->stream) != cudaSuccess)
        return fail("b2g_pairs_execute: memset failed");
    if (!in_rg.empty()) {
        std::vector<B2GRange> brg(in_rg.size());
        for (size_t i = 0; i < in_rg.size(); i++)
            brg[i] = B2GRange{in_rg[i].lo, in_rg[i].hi, in_rg[i].dev_off};
        if (b2g_mirror_ranges(ctx, brg, d_in))
            return fail("");
    }
    // wavefunction-side operands are addressed as offsets from d_in / d_out (negative or beyond the
    // allocation for resident blocks: plain device address arithmetic)
    for (B2GPair &q : hp) {
        if (q.pad & 4u)
            q.a0_off = (int64_t)((double *)(uintptr_t)q.a0_off - d_in);
        else {
            const Range &ra0 = locate(in_rg, (uintptr_t)q.a0_off);
            q.a0_off = (int64_t)(ra0.dev_off + ((uintptr_t)q.a0_off - ra0.lo) / sizeof(double));
        }
        if (!(q.pad & 1u)) {
            const Range &rb = locate(in_rg, (uintptr_t)q.b0);
            q.b0 = d_in + rb.dev_off + ((uintptr_t)q.b0 - rb.lo) / sizeof(double);
        }
        if (!(q.pad & 2u)) {
            const Range &ra = locate(in_rg, (uintptr_t)q.a1);
            q.a1 = d_in + ra.dev_off + ((uintptr_t)q.a1 - ra.lo) / sizeof(double);
        }
        if (q.pad & 8u)
            q.c1_off = (int64_t)((double *)(uintptr_t)q.c1_off - d_out);
        else {
            const Range &rc1 = locate(out_rg, (uintptr_t)q.c1_off);
            q.c1_off = (int64_t)(rc1.dev_off + ((uintptr_t)q.c1_off - rc1.lo) / sizeof(double));
        }
        q.pad = 0;
    }
    p->csize = (int64_t)in_total, p->vsize = (int64_t)out_total;
    const double t_up = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    prof_lap("pairs.mirror+translate");
    if (b2g_tiled_build(p))
        return fail("");
    prof_lap("pairs.tiled_build");
    rc = b2g_tiled_launch(p, d_in, d_out, 1.0);
    if (rc)
        return fail("");
    prof_lap("pairs.launch");
    // results: device -> pinned staging -> += into the host blocks (beta = 1 of the recorded GEMMs)
    {
        const size_t CH = B2G_UP_CHUNK / sizeof(double);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            return fail("b2g_pairs_execute: kernel execution failed");
        prof_lap("pairs.kernels_sync");
        for (const Range &r : out_rg) {
            const size_t len = (r.hi - r.lo) / sizeof(double);
            double *host = (double *)r.lo;
            for (size_t off = 0; off < len; off += CH) {
                const size_t m = std::min(CH, len - off);
                if (cudaMemcpyAsync(ctx->h_up[0], d_out + r.dev_off + off, m * sizeof(double), cudaMemcpyDeviceToHost,
                                    ctx->stream) != cudaSuccess ||
                    cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                    return fail("b2g_pairs_execute: result download failed");
                const double *src = (const double *)ctx->h_up[0];
                const int nt = m > ((size_t)1 << 16) ? ctx->up_threads : 1;
                const size_t slice = (m + nt - 1) / nt;
                std::vector<std::thread> th;
                for (int t = 1; t < nt; t++) {
                    const size_t lo = std::min(m, slice * t), hi = std::min(m, slice * (t + 1));
                    if (hi > lo)
                        th.emplace_back([=]() {
                            for (size_t j = lo; j < hi; j++)
                                host[off + j] += src[j];
                        });
                }
                for (size_t j = 0; j < std::min(m, slice); j++)
                    host[off + j] += src[j];
                for (auto &x : th)
                    x.join();
            }
        }
    }
    if (stats) {
        *stats = p->stats;
        stats->pairs = n, stats->nflop_mnk = nflop, stats->operand_doubles = (int64_t)in_total;
        stats->csize = (int64_t)in_total, stats->vsize = (int64_t)out_total, stats->upload_seconds = t_up;
    }
    prof_lap("pairs.download");
    b2g_dfree(ctx, d_in), b2g_dfree(ctx, d_out);
    b2g_plan_destroy(p);
    prof_lap("pairs.destroy");
    return 0;
}

// ------------------------------------------------------------------ shared range helpers

void b2g_merge_ranges(std::vector<B2GRange> &rg, size_t &total) {
    std::sort(rg.begin(), rg.end(), [](const B2GRange &x, const B2GRange &y) { return x.lo < y.lo; });
    std::vector<B2GRange> ar;
    for (const B2GRange &r : rg) {
        if (!ar.empty() && r.lo <= ar.back().hi)
            ar.back().hi = std::max(ar.back().hi, r.hi);
        else
            ar.push_back(r);
    }
    total = 0;
    for (B2GRange &r : ar) {
        r.dev_off = total;
        total += (r.hi - r.lo) / sizeof(double);
        total = (total + 1) & ~(size_t)1;
    }
    rg.swap(ar);
}

const B2GRange &b2g_locate_range(const std::vector<B2GRange> &ar, uintptr_t ptr) {
    size_t lo = 0, hi = ar.size();
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) / 2;
        if (ar[mid].lo <= ptr)
            lo = mid;
        else
            hi = mid;
    }
    return ar[lo];
}

// ------------------------------------------------------------------ device-resident operands

double *b2g_map_lookup(b2g_context *ctx, const void *ptr, size_t bytes) {
    const std::vector<B2GMapEntry> &m = ctx->rmap;
    if (m.empty())
        return nullptr;
    const uintptr_t lo = (uintptr_t)ptr, hi = lo + bytes;
    size_t a = 0, b = m.size();
    while (b - a > 1) {
        const size_t mid = (a + b) / 2;
        if (m[mid].lo <= lo)
            a = mid;
        else
            b = mid;
    }
    if (m[a].lo <= lo && hi <= m[a].hi)
        return m[a].dev + (lo - m[a].lo) / sizeof(double);
    return nullptr;
}

int b2g_mirror_ranges(b2g_context *ctx, const std::vector<B2GRange> &rg, double *dev_base) {
    if (ensure_upload_buffers(ctx))
        return 1;
    MirrorWriter mw{ctx, (char *)dev_base};
    B2G_CUDA(cudaEventSynchronize(ctx->up_done[0]));
    B2G_CUDA(cudaEventSynchronize(ctx->up_done[1]));
    for (const B2GRange &r : rg) {
        if (mw.add(r.dev_off * sizeof(double), (const void *)r.lo, r.hi - r.lo))
            return 1;
        ctx->mirrored_bytes += (int64_t)(r.hi - r.lo);
    }
    return mw.flush();
}

extern "C" int b2g_resident_map(b2g_context *ctx, int64_t count, const double *const *host, const int64_t *doubles,
                                double *const *dev) {
    if (!ctx || (count > 0 && (!host || !doubles || !dev))) {
        b2g_set_error("b2g_resident_map: null argument");
        return 1;
    }
    std::vector<B2GMapEntry> m;
    m.reserve((size_t)std::max<int64_t>(count, 0));
    for (int64_t i = 0; i < count; i++)
        if (host[i] && dev[i] && doubles[i] > 0)
            m.push_back(B2GMapEntry{(uintptr_t)host[i], (uintptr_t)host[i] + (uintptr_t)doubles[i] * sizeof(double),
                                    dev[i]});
    std::sort(m.begin(), m.end(), [](const B2GMapEntry &x, const B2GMapEntry &y) { return x.lo < y.lo; });
    for (size_t i = 1; i < m.size(); i++)
        if (m[i].lo < m[i - 1].hi) {
            b2g_set_error("b2g_resident_map: host ranges overlap (two live blocks cannot share host addresses)");
            return 1;
        }
    ctx->rmap.swap(m);
    return 0;
}

extern "C" int b2g_resident_stats(const b2g_context *ctx, int64_t *bytes_hit, int64_t *bytes_mirrored) {
    if (!ctx) {
        b2g_set_error("b2g_resident_stats: null context");
        return 1;
    }
    if (bytes_hit)
        *bytes_hit = ctx->resident_hit_bytes;
    if (bytes_mirrored)
        *bytes_mirrored = ctx->mirrored_bytes;
    return 0;
}

// Batched block transfers between host blocks and their device shadows, ordered on the context stream.
// Registered (pinned) host memory is copied directly; pageable memory goes through the pinned staging ring.
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int b2g_download(b2g_context *ctx, int64_t count, double *const *host, const double *const *dev,
                            const int64_t *doubles) {
    if (!ctx || (count > 0 && (!host || !dev || !doubles))) {
        b2g_set_error("b2g_download: null argument");
        return 1;
    }
    B2G_PROF_SCOPE("download_blocks");
    B2G_CUDA(cudaSetDevice(ctx->device));
    std::vector<B2GRange> staged; // pageable destinations: packed through the staging ring
    std::vector<const double *> staged_dev;
    for (int64_t i = 0; i < count; i++) {
        if (doubles[i] <= 0)
            continue;
        if (host_is_pinned(host[i]))
            B2G_CUDA(cudaMemcpyAsync(host[i], dev[i], (size_t)doubles[i] * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
        else {
            staged.push_back(B2GRange{(uintptr_t)host[i], (uintptr_t)host[i] + (size_t)doubles[i] * sizeof(double), 0});
            staged_dev.push_back(dev[i]);
        }
    }
    // pageable blocks one by one (each a contiguous device stretch): base = the block itself
    for (size_t i = 0; i < staged.size(); i++) {
        std::vector<B2GRange> one(1, staged[i]);
        if (b2g_download_ranges(ctx, one, staged_dev[i], false))
            return 1;
    }
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int b2g_upload_blocks(b2g_context *ctx, int64_t count, double *const *dev, const double *const *host,
                                 const int64_t *doubles) {
    if (!ctx || (count > 0 && (!host || !dev || !doubles))) {
        b2g_set_error("b2g_upload_blocks: null argument");
        return 1;
    }
    B2G_PROF_SCOPE("upload_blocks");
    B2G_CUDA(cudaSetDevice(ctx->device));
    if (ensure_upload_buffers(ctx))
        return 1;
    for (int64_t i = 0; i < count; i++) {
        if (doubles[i] <= 0)
            continue;
        const size_t bytes = (size_t)doubles[i] * sizeof(double);
        if (host_is_pinned(host[i]))
            B2G_CUDA(cudaMemcpyAsync(dev[i], host[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
        else if (bytes >= B2G_UP_CHUNK / 4) {
            if (b2g_upload(ctx, dev[i], host[i], bytes))
                return 1;
        } else { // small pageable block: one staged slice
            MirrorWriter mw{ctx, (char *)dev[i]};
            if (mw.add(0, host[i], bytes) || mw.flush())
                return 1;
        }
        ctx->mirrored_bytes += (int64_t)bytes;
    }
    B2G_CUDA(cudaStreamSynchronize(ctx->stream)); // the staging buffers and the host blocks are free again
    return 0;
}

extern "C" int b2g_host_register(b2g_context *ctx, void *ptr, size_t bytes) {
    if (!ctx || !ptr) {
        b2g_set_error("b2g_host_register: null argument");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}
extern "C" int b2g_host_unregister(b2g_context *ctx, void *ptr) {
    if (!ctx || !ptr) {
        b2g_set_error("b2g_host_unregister: null argument");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_CUDA(cudaHostUnregister(ptr));
    return 0;
}

int b2g_download_ranges(b2g_context *ctx, const std::vector<B2GRange> &rg, const double *dev_base, bool add) {
    if (ensure_upload_buffers(ctx))
        return 1;
    B2G_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t CH = B2G_UP_CHUNK / sizeof(double);
    // spans: contiguous device stretches of at most one staging buffer, each a list of host pieces
    struct Piece {
        double *host;
        size_t off, len; // offset inside the span, doubles
    };
    struct Span {
        size_t dev_lo, len;
        size_t p0, p1; // pieces [p0, p1)
    };
    std::vector<Piece> pieces;
    std::vector<Span> spans;
    for (size_t i = 0; i < rg.size();) {
        const size_t len_i = (rg[i].hi - rg[i].lo) / sizeof(double);
        if (len_i > CH) { // one large range: cut
            for (size_t off = 0; off < len_i; off += CH) {
                const size_t m = std::min(CH, len_i - off);
                spans.push_back(Span{rg[i].dev_off + off, m, pieces.size(), pieces.size() + 1});
                pieces.push_back(Piece{(double *)rg[i].lo + off, 0, m});
            }
            i++;
            continue;
        }
        const size_t span_lo = rg[i].dev_off;
        size_t j = i, span_hi = span_lo;
        const size_t p0 = pieces.size();
        while (j < rg.size()) {
            const size_t l = (rg[j].hi - rg[j].lo) / sizeof(double), e = rg[j].dev_off + l;
            if (e - span_lo > CH)
                break;
            pieces.push_back(Piece{(double *)rg[j].lo, rg[j].dev_off - span_lo, l});
            span_hi = e, j++;
        }
        spans.push_back(Span{span_lo, span_hi - span_lo, p0, pieces.size()});
        i = j;
    }
    auto issue = [&](size_t k) -> int {
        const int buf = (int)(k & 1);
        B2G_CUDA(cudaMemcpyAsync(ctx->h_up[buf], dev_base + spans[k].dev_lo, spans[k].len * sizeof(double),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        B2G_CUDA(cudaEventRecord(ctx->up_done[buf], ctx->stream));
        return 0;
    };
    if (!spans.empty() && issue(0))
        return 1;
    for (size_t k = 0; k < spans.size(); k++) {
        const int buf = (int)(k & 1);
        B2G_CUDA(cudaEventSynchronize(ctx->up_done[buf]));
        if (k + 1 < spans.size() && issue(k + 1)) // the next span travels while this one is folded into the host
            return 1;
        const double *stage = (const double *)ctx->h_up[buf];
        const Span &sp = spans[k];
        auto fold = [&](size_t lo, size_t hi) { // element range [lo, hi) of the span
            for (size_t q = sp.p0; q < sp.p1; q++) {
                const Piece &pc = pieces[q];
                const size_t a0 = std::max(lo, pc.off), a1 = std::min(hi, pc.off + pc.len);
                if (a0 >= a1)
                    continue;
                double *h = pc.host + (a0 - pc.off);
                const double *sv = stage + a0;
                if (add)
                    for (size_t x = 0; x < a1 - a0; x++)
                        h[x] += sv[x];
                else
                    memcpy(h, sv, (a1 - a0) * sizeof(double));
            }
        };
        const int nt = sp.len > ((size_t)1 << 16) ? ctx->up_threads : 1;
        if (nt == 1)
            fold(0, sp.len);
        else {
            const size_t slice = (sp.len + nt - 1) / nt;
            std::vector<std::thread> th;
            for (int t = 1; t < nt; t++) {
                const size_t lo = std::min(sp.len, slice * t), hi = std::min(sp.len, slice * (t + 1));
                if (hi > lo)
                    th.emplace_back(fold, lo, hi);
            }
            fold(0, std::min(sp.len, slice));
            for (auto &x : th)
                x.join();
        }
    }
    return 0;
}
