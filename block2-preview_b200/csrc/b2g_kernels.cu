// b2g_kernels.cu — device kernels of the H.C replay (sm_100a).
//
// Semantics being executed (block2 src/core/batch_gemm.hpp:1634-1643, one pair):
//   W             = alpha0 * op(c + a0_off) * op(b0)
//   v + c1_off   += alpha1 * scale * op(a1) * W
#include "b2g_internal.h"

// ---------------------------------------------------------------------------
// Generic pair kernel: one CTA per pair, any shape, any leading dimension.
// W lives in shared memory when it fits, otherwise in a per-CTA global slot.
// Kept as the correctness anchor for the tiled kernels and as the route for
// shapes they do not cover.
// ---------------------------------------------------------------------------
constexpr int GEN_THREADS = 256;
constexpr int GEN_SMEM_W = 4096; // doubles of W held in shared memory

__global__ void __launch_bounds__(GEN_THREADS)
b2g_pair_generic_kernel(const B2GPair *__restrict__ pairs, int64_t npairs, const double *__restrict__ c,
                        double *__restrict__ v, double scale, double *__restrict__ work, int64_t work_stride) {
    __shared__ double w_s[GEN_SMEM_W];
    for (int64_t ip = blockIdx.x; ip < npairs; ip += gridDim.x) {
        const B2GPair q = pairs[ip];
        const int m0 = q.m0, n0 = q.n0, k0 = q.k0, m1 = q.m1;
        const bool ta0 = q.flags & B2G_F_TA0, tb0 = q.flags & B2G_F_TB0, ta1 = q.flags & B2G_F_TA1;
        const double *a0 = c + q.a0_off;
        double *w = (m0 * n0 <= GEN_SMEM_W) ? w_s : work + (int64_t)blockIdx.x * work_stride;
        for (int idx = threadIdx.x; idx < m0 * n0; idx += GEN_THREADS) {
            const int i = idx / n0, j = idx - i * n0;
            double s = 0.0;
            for (int l = 0; l < k0; l++) {
                const double av = ta0 ? a0[(size_t)l * q.lda0 + i] : a0[(size_t)i * q.lda0 + l];
                const double bv = tb0 ? q.b0[(size_t)j * q.ldb0 + l] : q.b0[(size_t)l * q.ldb0 + j];
                s = fma(av, bv, s);
            }
            w[idx] = q.alpha0 * s;
        }
        __syncthreads();
        const double al = q.alpha1 * scale;
        double *out = v + q.c1_off;
        for (int idx = threadIdx.x; idx < m1 * n0; idx += GEN_THREADS) {
            const int i = idx / n0, j = idx - i * n0;
            double s = 0.0;
            for (int l = 0; l < m0; l++) {
                const double av = ta1 ? q.a1[(size_t)l * q.lda1 + i] : q.a1[(size_t)i * q.lda1 + l];
                s = fma(av, w[(size_t)l * n0 + j], s);
            }
            atomicAdd(&out[(size_t)i * q.ldc1 + j], al * s);
        }
        __syncthreads();
    }
}

int b2g_launch_matvec(b2g_plan *p, const double *c_dev, double *v_dev, double scale) {
    b2g_context *ctx = p->ctx;
    if (p->npairs == 0)
        return 0;
    if (p->tiled)
        return b2g_tiled_launch(p, c_dev, v_dev, scale);
    if (p->n_generic > 0) {
        const int grid = (int)std::min<int64_t>(p->n_generic, (int64_t)ctx->sm_count * 8);
        const size_t need = (p->max_work > GEN_SMEM_W) ? (size_t)p->max_work * grid : 0;
        if (need > p->work_doubles) {
            b2g_dfree(ctx, p->d_work);
            p->d_work = nullptr, p->work_doubles = 0;
            if (b2g_dmalloc(ctx, (void **)&p->d_work, need * sizeof(double)))
                return 1;
            p->work_doubles = need;
        }
        b2g_pair_generic_kernel<<<grid, GEN_THREADS, 0, ctx->stream>>>(p->d_pairs, p->n_generic, c_dev, v_dev, scale,
                                                                       p->d_work, p->max_work);
        ctx->launches++;
        B2G_CUDA(cudaGetLastError());
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Grouped GEMM list (cblas_dgemm_batch signature, block2 batch_gemm.hpp:81-111):
// the blocking / rotation lists (tensor_product rows-as-AXPY, tensor_rotate pairs).
// One CTA per GEMM, generic shapes; members of one call must not alias in C
// (the reference guarantees that per simple_perform / prepare() batch).
// ---------------------------------------------------------------------------
struct B2GGemm {
    const double *a, *b;
    double *c;
    double alpha, beta;
    int32_t m, n, k, lda, ldb, ldc;
    uint32_t flags, pad;
};

__global__ void __launch_bounds__(256)
b2g_gemm_list_kernel(const B2GGemm *__restrict__ g, int64_t count) {
    for (int64_t ig = blockIdx.x; ig < count; ig += gridDim.x) {
        const B2GGemm q = g[ig];
        const bool ta = q.flags & 1u, tb = q.flags & 2u;
        for (int64_t idx = threadIdx.x; idx < (int64_t)q.m * q.n; idx += blockDim.x) {
            const int i = (int)(idx / q.n), j = (int)(idx - (int64_t)i * q.n);
            double s = 0.0;
            for (int l = 0; l < q.k; l++) {
                const double av = ta ? q.a[(size_t)l * q.lda + i] : q.a[(size_t)i * q.lda + l];
                const double bv = tb ? q.b[(size_t)j * q.ldb + l] : q.b[(size_t)l * q.ldb + j];
                s = fma(av, bv, s);
            }
            double *cp = q.c + (size_t)i * q.ldc + j;
            *cp = q.beta == 0.0 ? q.alpha * s : fma(q.beta, *cp, q.alpha * s);
        }
    }
}

extern "C" int b2g_dgemm_batch(b2g_context *ctx, int64_t group_count, const int32_t *ta, const int32_t *tb,
                               const int32_t *m, const int32_t *n, const int32_t *k, const double *alpha,
                               const double *const *a, const int32_t *lda, const double *const *b,
                               const int32_t *ldb, const double *beta, double *const *c, const int32_t *ldc,
                               const int32_t *group_size) {
    if (!ctx) {
        b2g_set_error("b2g_dgemm_batch: null context");
        return 1;
    }
    B2G_CUDA(cudaSetDevice(ctx->device));
    std::vector<B2GGemm> h;
    int64_t z = 0;
    for (int64_t g = 0; g < group_count; g++)
        for (int32_t j = 0; j < group_size[g]; j++, z++) {
            B2GGemm q;
            q.a = a[z], q.b = b[z], q.c = c[z];
            q.alpha = alpha[g], q.beta = beta[g];
            q.m = m[g], q.n = n[g], q.k = k[g], q.lda = lda[g], q.ldb = ldb[g], q.ldc = ldc[g];
            q.flags = ((ta[g] == B2G_TRANS || ta[g] == 1) ? 1u : 0u) | ((tb[g] == B2G_TRANS || tb[g] == 1) ? 2u : 0u);
            q.pad = 0;
            if (q.m > 0 && q.n > 0)
                h.push_back(q);
        }
    if (h.empty())
        return 0;
    B2GGemm *d = nullptr;
    B2G_CUDA(cudaMallocAsync(&d, h.size() * sizeof(B2GGemm), ctx->stream));
    B2G_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(B2GGemm), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (int)std::min<int64_t>((int64_t)h.size(), (int64_t)ctx->sm_count * 8);
    b2g_gemm_list_kernel<<<grid, 256, 0, ctx->stream>>>(d, (int64_t)h.size());
    ctx->launches++;
    B2G_CUDA(cudaGetLastError());
    B2G_CUDA(cudaFreeAsync(d, ctx->stream));
    B2G_CUDA(cudaStreamSynchronize(ctx->stream)); // h (pageable) must outlive the copy
    return 0;
}
