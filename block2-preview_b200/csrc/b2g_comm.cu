// b2g_comm.cu — NCCL plumbing for the multi-GPU H.C path: one process per GPU,
// each rank owns a slice of the MPO terms (the reference's ParallelRuleQC split,
// block2 src/dmrg/qc_parallel_rule.hpp:44-80) and the partial sigma vectors are
// summed in place, on the context stream, right behind the matvec kernels.
// Replaces MPICommunicator::allreduce_sum(double*, size_t)
// (src/core/parallel_mpi.hpp:300-309) as used by ParallelTensorFunctions::operator()
// (src/core/parallel_tensor_functions.hpp:51-55).
//
// NCCL is bound at run time (dlopen) so the library loads on boxes without it and
// shares the copy a host process (e.g. torch) may already have mapped.
#include "b2g_internal.h"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle)
        return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle)
            break;
    }
    if (!g_nccl.handle) {
        b2g_set_error(std::string("b2g_comm: cannot dlopen libnccl.so.2: ") + dlerror());
        return 1;
    }
#define B2G_SYM(field, name)                                                            \
    g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.handle, name);                  \
    if (!g_nccl.field) {                                                                \
        b2g_set_error(std::string("b2g_comm: missing symbol ") + name);                 \
        return 1;                                                                       \
    }
    B2G_SYM(GetUniqueId, "ncclGetUniqueId")
    B2G_SYM(CommInitRank, "ncclCommInitRank")
    B2G_SYM(CommDestroy, "ncclCommDestroy")
    B2G_SYM(AllReduce, "ncclAllReduce")
    B2G_SYM(GetErrorString, "ncclGetErrorString")
#undef B2G_SYM
    return 0;
}
} // namespace

#define B2G_NCCL(expr)                                                                  \
    do {                                                                                \
        ncclResult_t r__ = (expr);                                                      \
        if (r__ != ncclSuccess) {                                                       \
            b2g_set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(r__));      \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in the C ABI");

extern "C" int b2g_comm_unique_id(void *id128) {
    if (load_nccl())
        return 1;
    B2G_NCCL(g_nccl.GetUniqueId((ncclUniqueId *)id128));
    return 0;
}

extern "C" int b2g_comm_init(b2g_context *ctx, int nranks, int rank, const void *id128) {
    if (!ctx || !id128) {
        b2g_set_error("b2g_comm_init: null argument");
        return 1;
    }
    if (load_nccl())
        return 1;
    B2G_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    B2G_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = (void *)comm, ctx->nranks = nranks, ctx->rank = rank;
    return 0;
}

extern "C" int b2g_comm_destroy(b2g_context *ctx) {
    if (ctx && ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr, ctx->nranks = 1, ctx->rank = 0;
    }
    return 0;
}

extern "C" int b2g_allreduce_sum(b2g_context *ctx, double *dev, int64_t count) {
    if (!ctx || !ctx->nccl_comm) {
        b2g_set_error("b2g_allreduce_sum: communicator not initialised (call b2g_comm_init)");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    B2G_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm,
                              ctx->stream));
    return 0;
}
