/* include/b2g.h — C ABI of libb2g.so, the sm_100a CUDA executor for block2's
 * Davidson H.C hot path.  Plain pointers and sizes only; no torch, no
 * reference types.  Every entry point returns 0 on success and a non-zero
 * code on failure (b2g_last_error() holds the message); nothing here falls
 * back to the CPU.
 *
 * Reference interfaces replaced (paths under block2 src/):
 *   b2g_batch            <- BatchGEMM<double> SoA arrays      core/batch_gemm.hpp:237-247
 *   b2g_plan_create      <- EffectiveHamiltonian::precompute  dmrg/effective_hamiltonian.hpp:226-246
 *                           (consumes seq->batch[0], seq->batch[1], seq->max_work)
 *   b2g_seq_matvec       <- BatchGEMMSeq<double>::operator()(c, v, scale), Tasked branch
 *                           core/batch_gemm.hpp:1570-1691 (= TensorFunctions::operator(),
 *                           core/tensor_functions.hpp:59-62)
 *   b2g_plan_destroy     <- EffectiveHamiltonian::post_precompute  :247-253
 *   b2g_pairs_execute    <- OperatorFunctions::tensor_rotate lists executed by BatchGEMMSeq::auto_perform /
 *                           simple_perform (left_rotate / right_rotate, core/tensor_functions.hpp:2365-2403)
 *   b2g_dgemm_batch      <- cblas_xgemm_batch / BatchGEMM::perform  core/batch_gemm.hpp:81-111, 339-357
 *   b2g_batch_execute    <- BatchGEMMSeq::auto_perform / simple_perform on a single-batch (batch[1]-only) list
 *                           core/batch_gemm.hpp:1417-1530: the blocking lists TensorFunctions::left_contract /
 *                           right_contract record (core/tensor_functions.hpp:2842-2885, 2941-2984) through
 *                           OperatorFunctions::tensor_product (core/operator_functions.hpp:672-711) and
 *                           AdvancedGEMM::tensor_product (core/batch_gemm.hpp:433-503), plus iadd / iscale /
 *                           tensor_product_diagonal entries (:1110-1135, 327-336, 506-511)
 *   b2g_tensor_product_execute
 *                        <- the same blocking step one level up: the GMatrixFunctions::tensor_product calls
 *                           (core/matrix_functions.hpp:1269-1397) OperatorFunctions::tensor_product makes per
 *                           connection-info entry (core/operator_functions.hpp:672-711)
 *   b2g_resident_map / b2g_download / b2g_upload_blocks / b2g_host_register
 *                        <- the partition traffic of MovingEnvironment::left_contract_rotate / move_to
 *                           (dmrg/moving_environment.hpp:226-430, 1542-1575) through DataFrame stack 1
 *                           (core/allocator.hpp:617-660): environments stay in HBM
 *   b2g_davidson         <- IterativeMatrixFunctions<double>::davidson (k = 1, Normal type,
 *                           Olsen preconditioner)  core/iterative_matrix_functions.hpp:864-1173, 93-108
 *   b2g_comm_* / b2g_allreduce_sum
 *                        <- MPICommunicator::allreduce_sum(double*, size_t)  core/parallel_mpi.hpp:300-309
 *                           as used by ParallelTensorFunctions::operator()  core/parallel_tensor_functions.hpp:51-55
 */
#ifndef B2G_H
#define B2G_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_NOTRANS 111 /* CblasNoTrans; 0 is accepted too */
#define B2G_TRANS 112   /* CblasTrans;   1 is accepted too */

#define B2G_OPERANDS_HOST 0   /* a/b operator pointers are host addresses: mirrored to HBM by the plan */
#define B2G_OPERANDS_DEVICE 1 /* a/b operator pointers already are device addresses (resident environments) */

typedef struct b2g_context b2g_context;
typedef struct b2g_plan b2g_plan;

/* One recorded GEMM list, entry i = one row-major GEMM
 *   C_i[m x n] = alpha_i * op(A_i) * op(B_i) + beta_i * C_i,   lda >= (ta ? m : k), ldb >= (tb ? k : n), ldc >= n
 * exactly as BatchGEMM<double> stores it with gp[i] == 1. */
typedef struct b2g_batch {
    int64_t count;
    const int32_t *ta, *tb;
    const int32_t *m, *n, *k;
    const int32_t *lda, *ldb, *ldc;
    const double *alpha, *beta;
    const double *const *a;
    const double *const *b;
    double *const *c;
} b2g_batch;

typedef struct b2g_plan_stats {
    int64_t pairs;          /* GEMM pairs per matvec */
    int64_t csize, vsize;   /* |c|, |sigma| in doubles */
    int64_t nflop_mnk;      /* sum m*n*k over both GEMMs (reference units, batch_gemm.hpp:307) */
    int64_t operand_doubles;/* distinct operator doubles referenced (mirrored or resident) */
    int64_t arenas;         /* merged contiguous operand ranges */
    int64_t launches;       /* kernel launches per matvec */
    int64_t n_small, n_large; /* pairs executed by the generic one-CTA-per-pair kernel / by the DMMA tile engine */
    double upload_seconds;  /* host->device mirror time of the operands */
    int64_t mirrored_doubles;  /* operator doubles copied into the plan's own arena (0: all read in place) */
    int64_t workspace_doubles; /* W panels + partial sigma tiles of the tile engine */
} b2g_plan_stats;

/* per-launch record of one profiled matvec (b2g_plan_profile) */
typedef struct b2g_kernel_stat {
    char name[64];   /* e.g. phase2_128x64_An */
    double flops;    /* useful 2*m*n*k FLOPs of the launch */
    double ms;       /* CUDA-event duration on the context stream */
    int64_t units;   /* CTA tile x K-chunk work units */
} b2g_kernel_stat;

const char *b2g_last_error(void);
int b2g_device_count(void);
int b2g_context_create(int device, b2g_context **ctx);
int b2g_context_destroy(b2g_context *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t b2g_context_launches(const b2g_context *ctx);
/* the CUDA stream all work of this context is ordered on (cudaStream_t) */
void *b2g_context_stream(const b2g_context *ctx);
int b2g_context_synchronize(b2g_context *ctx);

/* Host-side wall-clock profile of the library and its binding (measurement only; enabled by the environment
 * variable B2G_PROF): seconds and calls accumulated per label.  b2g_prof_record adds to a label (the
 * reference-side binding uses it for its own sections), b2g_prof_dump writes one JSON object to `path`
 * (NULL or "-": stderr) and returns 0; both are no-ops when the profile is off. */
/* Host-side regrouping of a chained pair list into the two-phase tile plan, without a device (planner timing and
 * tests): work units, launches, and a fingerprint (FNV-1a) of everything the device would receive - unit lists with
 * their partial-slot offsets, sigma tiles, slot table, W offsets. */
int b2g_debug_tiled_plan(const b2g_batch *batch0, const b2g_batch *batch1, double *seconds, int64_t *units,
                         int64_t *launches, int64_t *fingerprint);
int b2g_prof_enabled(void);
void b2g_prof_record(const char *label, double seconds);
int b2g_prof_dump(const char *path);

/* Build the replay plan of one H_eff from the two recorded lists.
 * Pair i:  W = alpha0 * op(c + a0_i) * op(B0_i);   sigma + c1_i += alpha1 * scale * op(A1_i) * W
 * batch0->a[i] and batch1->c[i] are null-based offsets into c / sigma (the reference records
 * them with cmat->data = vmat->data = 0), batch0->c[i] == batch1->b[i] is the work slot. */
int b2g_plan_create(b2g_context *ctx, const b2g_batch *batch0, const b2g_batch *batch1,
                    int64_t max_work, int64_t csize, int64_t vsize, int operand_space,
                    b2g_plan **plan);
int b2g_plan_destroy(b2g_plan *plan);
int b2g_plan_get_stats(const b2g_plan *plan, b2g_plan_stats *out);

/* sigma += scale * H.c, host buffers (drop-in for BatchGEMMSeq::operator()). Synchronous. */
int b2g_seq_matvec(b2g_plan *plan, const double *c_host, double *v_host, double scale);
/* same with device-resident c and sigma; asynchronous on the context stream */
int b2g_seq_matvec_dev(b2g_plan *plan, const double *c_dev, double *v_dev, double scale);

/* One matvec with CUDA events between the kernel launches (measurement only; synchronous). */
int b2g_plan_profile(b2g_plan *plan, const double *c_dev, double *v_dev, double scale,
                     b2g_kernel_stat *out, int capacity, int *count);

/* Execute once a recorded chained-pair list whose operands are ALL host pointers - the list
 * OperatorFunctions::tensor_rotate records through BatchGEMMSeq::rotate when an environment block
 * is renormalised (core/operator_functions.hpp:175-210, core/tensor_functions.hpp:2365-2403):
 *     W_i = alpha0 * op(A0_i) * op(B0_i);      C1_i += alpha1 * op(A1_i) * W_i
 * A0/B0/A1 ranges are mirrored to HBM, the products run through the same tile engine as the H.C
 * replay, and the results are added into the host C1 blocks (beta = 1).  Synchronous.
 * stats (optional): pairs, nflop_mnk, operand_doubles (inputs), upload_seconds. */
int b2g_pairs_execute(b2g_context *ctx, const b2g_batch *batch0, const b2g_batch *batch1,
                      int64_t max_work, b2g_plan_stats *stats);

/* Grouped GEMM list with the cblas_dgemm_batch signature (device pointers), asynchronous. */
int b2g_dgemm_batch(b2g_context *ctx, int64_t group_count, const int32_t *ta, const int32_t *tb,
                    const int32_t *m, const int32_t *n, const int32_t *k, const double *alpha,
                    const double *const *a, const int32_t *lda, const double *const *b,
                    const int32_t *ldb, const double *beta, double *const *c, const int32_t *ldc,
                    const int32_t *group_size);

/* per-call record of b2g_batch_execute */
typedef struct b2g_blocking_stats {
    int64_t entries;        /* GEMMs of the list after group expansion */
    int64_t merged;         /* entries after constant-stride rows were folded into 2-D windows */
    int64_t clusters;       /* distinct output windows (each written once, contributions summed in registers) */
    int64_t units;          /* warp work units */
    int64_t serial_entries; /* entries of irregularly overlapping windows (executed in list order by one CTA) */
    int64_t nflop_mnk;      /* sum m*n*k (reference units) */
    int64_t bytes_in;       /* 8 * source elements read (algorithmic) */
    int64_t bytes_out;      /* 8 * destination elements written (algorithmic) */
    int64_t launches;
    double kernel_ms;       /* CUDA events around the kernels on the context stream */
    double upload_seconds, download_seconds, plan_seconds;
} b2g_blocking_stats;

#define B2G_PLAN_ONLY 8 /* regroup the list (clusters, units, serial components, algorithmic bytes) and return the
                           stats without touching a device; ctx may be NULL */
#define B2G_DST_ZERO 1 /* caller guarantees every output block is zero on entry (freshly allocate()d operators):
                          outputs are not uploaded, the device result is added into the host blocks */

/* Execute a recorded single-batch GEMM list (cblas_dgemm_batch group signature, exactly the arrays of
 * BatchGEMM<double>: per-group parameters, per-entry pointers) whose entries may write the SAME output
 * blocks - the conflict-carrying list BatchGEMMSeq::auto_perform resolves with work arrays and a
 * post-batch reduction.  Entry semantics: C = alpha * op(A) * op(B) + beta * C, applied per output
 * element IN LIST ORDER (the order simple_perform would execute them), each output element by one thread:
 * deterministic, no atomics.  operand_space = B2G_OPERANDS_HOST: all pointers are host addresses, inputs
 * are mirrored, results copied back (synchronous).  B2G_OPERANDS_DEVICE: all pointers are device
 * addresses, asynchronous on the context stream apart from the plan upload. */
int b2g_batch_execute(b2g_context *ctx, int64_t group_count, const int32_t *ta, const int32_t *tb,
                      const int32_t *m, const int32_t *n, const int32_t *k, const double *alpha,
                      const double *const *a, const int32_t *lda, const double *const *b,
                      const int32_t *ldb, const double *beta, double *const *c, const int32_t *ldc,
                      const int32_t *group_size, int operand_space, int flags, b2g_blocking_stats *stats);

/* One eager tensor-product call of the blocking step, GMatrixFunctions<double>::tensor_product(a, conja, b,
 * conjb, c, scale, stride) (core/matrix_functions.hpp:1269-1397) as emitted per connection-info entry by
 * OperatorFunctions::tensor_product (core/operator_functions.hpp:672-711):
 *     C[(i*bm' + k), (j*bn' + l)] += scale * op(A)(i, j) * op(B)(k, l)
 * a: am x an row-major block, b: bm x bn row-major block, op = transpose when conj != 0,
 * c: address of the window's first element (block base + stride), cn: pitch of the output block. */
typedef struct b2g_tp_term {
    const double *a;
    const double *b;
    double *c;
    int32_t am, an, bm, bn, cn;
    int32_t conja, conjb, reserved;
    double scale;
} b2g_tp_term;

/* Execute a list of tensor-product terms; terms that write the same window are applied in list order by
 * the thread that owns the output element (same back end and flags as b2g_batch_execute).  This is the
 * compact form of the blocking list: one descriptor per (a-block, b-block) pair instead of one GEMM per row. */
int b2g_tensor_product_execute(b2g_context *ctx, int64_t count, const b2g_tp_term *terms, int operand_space,
                               int flags, b2g_blocking_stats *stats);

/* Device-resident operands (environments stay in HBM between left/right_contract, left/right_rotate and
 * the H.C plan of the next site).  The host binding owns device buffers (b2g_malloc) that shadow operator blocks
 * and tells the library where they are: a table of host ranges whose content currently lives at the given
 * device addresses.  Every entry point that takes HOST operand pointers (b2g_plan_create, b2g_pairs_execute,
 * b2g_batch_execute, b2g_tensor_product_execute) consults the table:
 *   - an input operand that lies inside a mapped range is read in place from HBM (no upload, no copy);
 *   - an output window that lies inside a mapped range is written in place in HBM and NOT copied back to the
 *     host (with B2G_DST_ZERO the caller has zeroed the device buffer; the host range is never touched, it may
 *     be address space without memory behind it).
 * Ranges must not overlap.  The table stays in force until the next b2g_resident_map call (count = 0 clears it).
 * It replaces the per-partition host<->device traffic of MovingEnvironment::left_contract_rotate / move_to
 * (dmrg/moving_environment.hpp:226-430, 1542-1575), whose data live in DataFrame stack 1 (core/allocator.hpp:617-660). */
int b2g_resident_map(b2g_context *ctx, int64_t count, const double *const *host, const int64_t *doubles,
                     double *const *dev);
/* host->device bytes avoided by resident inputs so far / bytes mirrored from the host so far */
int b2g_resident_stats(const b2g_context *ctx, int64_t *bytes_hit, int64_t *bytes_mirrored);
/* Block transfers between host blocks and their device shadows, ordered on the context stream; synchronous.
 * Registered (b2g_host_register) host memory is copied by DMA directly, pageable memory through pinned staging. */
int b2g_download(b2g_context *ctx, int64_t count, double *const *host, const double *const *dev,
                 const int64_t *doubles);
int b2g_upload_blocks(b2g_context *ctx, int64_t count, double *const *dev, const double *const *host,
                      const int64_t *doubles);
/* page-lock an existing host allocation (e.g. the DataFrame stacks, core/allocator.hpp:536) */
int b2g_host_register(b2g_context *ctx, void *ptr, size_t bytes);
int b2g_host_unregister(b2g_context *ctx, void *ptr);

/* Davidson ground state with device-resident vectors; H applied through the plan.
 * ket_host: in = initial guess, out = eigenvector.  diag_host: H_eff diagonal.
 * Returns the eigenvalue (without const_e) and the number of matvecs, like
 * EffectiveHamiltonian::eigs (effective_hamiltonian.hpp:480-566). */
int b2g_davidson(b2g_plan *plan, const double *diag_host, double *ket_host, double conv_thrd,
                 double rel_conv_thrd, int max_iter, int soft_max_iter, int deflation_min_size,
                 int deflation_max_size, double *eigenvalue, int *ndav);

/* Multi-GPU: one process per GPU, sigma summed with an NCCL all-reduce on the context stream. */
int b2g_comm_unique_id(void *id128);               /* rank 0: fill a 128-byte ncclUniqueId */
int b2g_comm_init(b2g_context *ctx, int nranks, int rank, const void *id128);
int b2g_comm_destroy(b2g_context *ctx);
int b2g_allreduce_sum(b2g_context *ctx, double *dev, int64_t count); /* in place, async */

/* device memory helpers for hosts without a CUDA runtime binding of their own; b2g_malloc / b2g_free are
 * stream-ordered pool allocations on the context stream (cheap enough for one buffer per operator tensor) */
int b2g_malloc(b2g_context *ctx, size_t bytes, void **dev);
int b2g_free(b2g_context *ctx, void *dev);
/* free (including memory parked in the allocation pool) and total device memory */
int b2g_mem_info(b2g_context *ctx, int64_t *free_bytes, int64_t *total_bytes);
/* Dense symmetric eigenproblem on the device, LIBRARY-BACKED (cuSOLVER cusolverDnDsyevd, loaded with dlopen at
 * first use): a (n x n, leading dimension lda, host) is overwritten by its eigenvectors exactly as LAPACK
 * dsyev("V", "U") leaves them, w receives the eigenvalues in ascending order.  Thread-safe, synchronous.  Replaces
 * the dsyev calls of MovingEnvironment::truncate_density_matrix (dmrg/moving_environment.hpp:3716-3790) when the
 * binding routes them here (not on the north-star path).  Non-zero: a is unchanged. */
int b2g_syevd(b2g_context *ctx, int n, double *a_host, int lda, double *w_host);
/* synchronise and return every unused block of the stream-ordered pool to the driver (after evicting shadows) */
int b2g_mem_trim(b2g_context *ctx);
int b2g_memcpy_h2d(b2g_context *ctx, void *dev, const void *host, size_t bytes);
int b2g_memcpy_d2h(b2g_context *ctx, void *host, const void *dev, size_t bytes);
int b2g_memset_zero(b2g_context *ctx, void *dev, size_t bytes);

/* test hook, no device: byte ranges [lo[t], hi[t]) the nt staging threads copy for an upload chunk of len bytes */
int b2g_debug_upload_slices(int64_t len, int nt, int64_t *lo, int64_t *hi);

#ifdef __cplusplus
}
#endif
#endif /* B2G_H */
