// b2g_eig.cu — dense symmetric eigenproblems of the density-matrix split on the device.
//
// LIBRARY-BACKED, and said so: this is cuSOLVER's cusolverDnDsyevd, not a kernel of this repository.  The split of
// the two-site wavefunction (MovingEnvironment::split_density_matrix -> truncate_density_matrix,
// dmrg/moving_environment.hpp:3716-3790, 4250) diagonalises every quantum-number block of the density matrix with
// LAPACK dsyev, one block per OpenMP thread; at Cr2 M=4000 with density-matrix noise the blocks are dense and up to a
// few thousand rows, and this step is 2/3 of the sweep once H.C, blocking and rotation run on the GPU.  It is not on
// the north-star path (H.C, blocking, Davidson); the entry point exists so that the reference-side binding can route
// the dsyev calls of that step to the device (host/b2g_adapter.hpp, opt-in).
//
// cuSOLVER is loaded with dlopen at first use: libb2g.so itself does not depend on it, and when it is missing the
// call fails loudly and the binding stays on the reference's CPU routine.
#include "b2g_internal.h"
#include <condition_variable>
#include <cstring>
#include <cusolverDn.h>
#include <dlfcn.h>
#include <mutex>

namespace {

struct EigSlot { // one concurrent caller: own stream, handle and buffers
    cudaStream_t stream = nullptr;
    cusolverDnHandle_t handle = nullptr;
    double *d_a = nullptr, *d_w = nullptr, *d_work = nullptr, *h_pin = nullptr;
    int *d_info = nullptr;
    size_t a_cap = 0, work_cap = 0, pin_cap = 0;
    bool busy = false;
};

struct EigPool {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<EigSlot *> slots;
    int max_slots = 8;
    void *lib = nullptr;
    bool tried = false;
    cusolverStatus_t (*create)(cusolverDnHandle_t *) = nullptr;
    cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*buffer_size)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double *, int,
                                    const double *, int *) = nullptr;
    cusolverStatus_t (*syevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double *, int, double *,
                              double *, int, int *) = nullptr;
};

std::mutex g_pools_mu;

EigPool *pool_of(b2g_context *ctx) {
    std::lock_guard<std::mutex> lk(g_pools_mu);
    if (ctx->eig_pool == nullptr) {
        EigPool *p = new EigPool();
        p->max_slots = std::max(1, std::min(16, ctx->up_threads));
        ctx->eig_pool = p;
    }
    return (EigPool *)ctx->eig_pool;
}

bool load_cusolver(EigPool *p) {
    if (p->tried)
        return p->syevd != nullptr;
    p->tried = true;
    for (const char *name : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"}) {
        p->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (p->lib)
            break;
    }
    if (!p->lib)
        return false;
    p->create = (decltype(p->create))dlsym(p->lib, "cusolverDnCreate");
    p->destroy = (decltype(p->destroy))dlsym(p->lib, "cusolverDnDestroy");
    p->set_stream = (decltype(p->set_stream))dlsym(p->lib, "cusolverDnSetStream");
    p->buffer_size = (decltype(p->buffer_size))dlsym(p->lib, "cusolverDnDsyevd_bufferSize");
    p->syevd = (decltype(p->syevd))dlsym(p->lib, "cusolverDnDsyevd");
    if (!p->create || !p->set_stream || !p->buffer_size || !p->syevd) {
        p->syevd = nullptr;
        return false;
    }
    return true;
}

struct SlotLease {
    EigPool *pool;
    EigSlot *slot;
    ~SlotLease() {
        if (slot) {
            std::lock_guard<std::mutex> lk(pool->mu);
            slot->busy = false;
            pool->cv.notify_one();
        }
    }
};

} // namespace

void b2g_eig_destroy(b2g_context *ctx) {
    EigPool *p = (EigPool *)ctx->eig_pool;
    if (!p)
        return;
    for (EigSlot *s : p->slots) {
        if (s->handle && p->destroy)
            p->destroy(s->handle);
        cudaFree(s->d_a), cudaFree(s->d_w), cudaFree(s->d_work), cudaFree(s->d_info);
        cudaFreeHost(s->h_pin);
        if (s->stream)
            cudaStreamDestroy(s->stream);
        delete s;
    }
    delete p;
    ctx->eig_pool = nullptr;
}

// a (n x n, leading dimension lda, symmetric; row- or column-major alike) is overwritten by its eigenvectors the way
// LAPACK dsyev("V", "U") leaves them (vector k in column k of the column-major view = row k of the row-major view,
// the reference's convention, core/matrix_functions.hpp:1672-1688); w receives the eigenvalues in ascending order.
// Host buffers; thread-safe (callers beyond the number of slots wait); synchronous.
extern "C" int b2g_syevd(b2g_context *ctx, int n, double *a_host, int lda, double *w_host) {
    if (!ctx || !a_host || !w_host || n <= 0 || lda < n) {
        b2g_set_error("b2g_syevd: bad argument");
        return 1;
    }
    B2G_PROF_SCOPE("syevd");
    EigPool *pool = pool_of(ctx);
    SlotLease lease{pool, nullptr};
    {
        std::unique_lock<std::mutex> lk(pool->mu);
        if (!load_cusolver(pool)) {
            b2g_set_error("b2g_syevd: cuSOLVER (libcusolver.so) could not be loaded");
            return 2;
        }
        for (;;) {
            for (EigSlot *s : pool->slots)
                if (!s->busy) {
                    lease.slot = s;
                    break;
                }
            if (lease.slot)
                break;
            if ((int)pool->slots.size() < pool->max_slots) {
                lease.slot = new EigSlot();
                pool->slots.push_back(lease.slot);
                break;
            }
            pool->cv.wait(lk);
        }
        lease.slot->busy = true;
    }
    EigSlot &s = *lease.slot;
    B2G_CUDA(cudaSetDevice(ctx->device));
    if (!s.stream) {
        B2G_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (pool->create(&s.handle) != CUSOLVER_STATUS_SUCCESS || pool->set_stream(s.handle, s.stream) != CUSOLVER_STATUS_SUCCESS) {
            b2g_set_error("b2g_syevd: cusolverDnCreate failed");
            return 1;
        }
        B2G_CUDA(cudaMalloc(&s.d_info, sizeof(int)));
    }
    const size_t nn = (size_t)n * n;
    if (nn > s.a_cap) {
        cudaFree(s.d_a), cudaFree(s.d_w);
        s.d_a = s.d_w = nullptr, s.a_cap = 0;
        B2G_CUDA(cudaMalloc(&s.d_a, nn * sizeof(double)));
        B2G_CUDA(cudaMalloc(&s.d_w, (size_t)n * sizeof(double)));
        s.a_cap = nn;
    }
    if (nn + n + 2 > s.pin_cap) {
        cudaFreeHost(s.h_pin);
        s.h_pin = nullptr, s.pin_cap = 0;
        B2G_CUDA(cudaMallocHost(&s.h_pin, (nn + n + 2) * sizeof(double)));
        s.pin_cap = nn + n + 2;
    }
    int lwork = 0;
    if (pool->buffer_size(s.handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, s.d_a, n, s.d_w, &lwork) !=
        CUSOLVER_STATUS_SUCCESS) {
        b2g_set_error("b2g_syevd: cusolverDnDsyevd_bufferSize failed");
        return 1;
    }
    if ((size_t)lwork > s.work_cap) {
        cudaFree(s.d_work);
        s.d_work = nullptr, s.work_cap = 0;
        B2G_CUDA(cudaMalloc(&s.d_work, (size_t)std::max(lwork, 1) * sizeof(double)));
        s.work_cap = (size_t)lwork;
    }
    if (lda == n)
        memcpy(s.h_pin, a_host, nn * sizeof(double));
    else
        for (int i = 0; i < n; i++)
            memcpy(s.h_pin + (size_t)i * n, a_host + (size_t)i * lda, (size_t)n * sizeof(double));
    B2G_CUDA(cudaMemcpyAsync(s.d_a, s.h_pin, nn * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    const cusolverStatus_t st = pool->syevd(s.handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, s.d_a, n,
                                            s.d_w, s.d_work, lwork, s.d_info);
    if (st != CUSOLVER_STATUS_SUCCESS) {
        cudaStreamSynchronize(s.stream);
        b2g_set_error("b2g_syevd: cusolverDnDsyevd failed with status " + std::to_string((int)st));
        return 1;
    }
    B2G_CUDA(cudaMemcpyAsync(s.h_pin, s.d_a, nn * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    B2G_CUDA(cudaMemcpyAsync(s.h_pin + nn, s.d_w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    B2G_CUDA(cudaMemcpyAsync(s.h_pin + nn + n, s.d_info, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    B2G_CUDA(cudaStreamSynchronize(s.stream));
    int info = 0;
    memcpy(&info, s.h_pin + nn + n, sizeof(int));
    if (info != 0) {
        b2g_set_error("b2g_syevd: cusolverDnDsyevd did not converge (info = " + std::to_string(info) + ")");
        return 3;
    }
    if (lda == n)
        memcpy(a_host, s.h_pin, nn * sizeof(double));
    else
        for (int i = 0; i < n; i++)
            memcpy(a_host + (size_t)i * lda, s.h_pin + (size_t)i * n, (size_t)n * sizeof(double));
    memcpy(w_host, s.h_pin + nn, (size_t)n * sizeof(double));
    return 0;
}
