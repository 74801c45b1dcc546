// b2g_davidson.cu — device-resident Davidson for the lowest eigenpair of H_eff.
//
// Same iteration as IterativeMatrixFunctions<double>::davidson for k = 1,
// DavidsonTypes::Normal (block2 src/core/iterative_matrix_functions.hpp:864-1173):
//   * H is applied only to the basis vectors added since the last iteration (:971-977)
//   * subspace matrix alpha(i,j) = <b_i|sigma_j>, j <= i, diagonalised (m <= 50) on the
//     host (:999-1003); basis and sigma are both rotated into the Ritz basis (:1005-1026)
//   * residual q = sigma_0 - theta b_0 (:1072-1073), convergence on |q.q| <
//     conv_thrd + theta^2 rel^2 (:1098-1100)
//   * Olsen preconditioner with the |theta - aa_i| > 1e-12 guard (:93-108)
//   * sequential re-orthogonalisation of q against all b_j, normalise, append (:1139-1146)
//   * deflation when m reaches deflation_max_size: m = msig = deflation_min_size (:1104-1107)
// The vectors never leave HBM; per iteration only the m x m matrix and |q|^2 cross PCIe.
#include "b2g_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_BLOCKS = 592; // 4 x 148 SMs

// out[0] = sum_e x[e] * y[e]; deterministic: fixed partition, partials reduced in index
// order by whichever block finishes last.
__global__ void __launch_bounds__(RED_THREADS)
dot_kernel(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *__restrict__ partials,
           unsigned int *__restrict__ counter, double *__restrict__ out) {
    __shared__ double sh[RED_THREADS];
    __shared__ bool last;
    double s = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; e < n; e += (int64_t)gridDim.x * RED_THREADS)
        s = fma(x[e], y[e], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS)
            t += partials[i];
        sh[threadIdx.x] = t;
        __syncthreads();
        for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off)
                sh[threadIdx.x] += sh[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            out[0] = sh[0];
            *counter = 0;
        }
    }
}

// y += (sign * num[0] / den[0]) * x   (den == nullptr: divide by 1)
__global__ void axpy_dev_kernel(double *__restrict__ y, const double *__restrict__ x, int64_t n,
                                const double *__restrict__ num, const double *__restrict__ den, double sign) {
    const double a = sign * num[0] / (den ? den[0] : 1.0);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        y[e] = fma(a, x[e], y[e]);
}

// x *= 1 / sqrt(nrm2[0])
__global__ void scale_rsqrt_kernel(double *__restrict__ x, int64_t n, const double *__restrict__ nrm2) {
    const double a = 1.0 / sqrt(nrm2[0]);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        x[e] *= a;
}

// q = sigma - theta * b
__global__ void residual_kernel(double *__restrict__ q, const double *__restrict__ sigma,
                                const double *__restrict__ b, double theta, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        q[e] = fma(-theta, b[e], sigma[e]);
}

// Olsen, first half: t = b; where |theta - aa| > 1e-12: t /= (theta - aa), q /= (theta - aa)
__global__ void olsen_divide_kernel(double *__restrict__ q, double *__restrict__ t, const double *__restrict__ b,
                                    const double *__restrict__ aa, double theta, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double d = theta - aa[e];
        double tv = b[e], qv = q[e];
        if (fabs(d) > 1E-12) {
            tv /= d;
            qv /= d;
        }
        t[e] = tv, q[e] = qv;
    }
}

// Lower triangle of G(i,j) = <B_i|S_j> for i,j < m: partial sums per block, then reduced.
constexpr int GRAM_CHUNK = 128;
__global__ void __launch_bounds__(256)
gram_partial_kernel(const double *__restrict__ B, const double *__restrict__ S, int m, int64_t n, int64_t ld,
                    double *__restrict__ partials) {
    extern __shared__ double sm[];
    double *bs = sm, *ss = sm + (size_t)m * GRAM_CHUNK;
    const int npair = m * (m + 1) / 2;
    // each thread owns pairs tid, tid + 256, ...
    double acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
        acc[r] = 0.0;
    for (int64_t e0 = (int64_t)blockIdx.x * GRAM_CHUNK; e0 < n; e0 += (int64_t)gridDim.x * GRAM_CHUNK) {
        const int len = (int)min((int64_t)GRAM_CHUNK, n - e0);
        for (int idx = threadIdx.x; idx < m * GRAM_CHUNK; idx += 256) {
            const int i = idx / GRAM_CHUNK, e = idx - i * GRAM_CHUNK;
            bs[idx] = e < len ? B[(size_t)i * ld + e0 + e] : 0.0;
            ss[idx] = e < len ? S[(size_t)i * ld + e0 + e] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int pr = threadIdx.x + r * 256;
            if (pr < npair) {
                // unrank (i, j), j <= i
                int i = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > pr) i--;
                while ((i + 1) * (i + 2) / 2 <= pr) i++;
                const int j = pr - i * (i + 1) / 2;
                const double *bi = bs + (size_t)i * GRAM_CHUNK, *sj = ss + (size_t)j * GRAM_CHUNK;
                double s = acc[r];
                for (int e = 0; e < GRAM_CHUNK; e++)
                    s = fma(bi[e], sj[e], s);
                acc[r] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int pr = threadIdx.x + r * 256;
        if (pr < npair)
            partials[(size_t)blockIdx.x * npair + pr] = acc[r];
    }
}

__global__ void gram_reduce_kernel(const double *__restrict__ partials, int npair, int nblocks,
                                   double *__restrict__ out) {
    const int pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr < npair) {
        double s = 0.0;
        for (int b = 0; b < nblocks; b++)
            s += partials[(size_t)b * npair + pr];
        out[pr] = s;
    }
}

// In-place X_j <- sum_i rot(j, i) X_i for the m vectors stored with leading dimension ld.
constexpr int ROT_MAX = 64;
__global__ void __launch_bounds__(128)
rotate_kernel(double *__restrict__ X, int m, int64_t n, int64_t ld, const double *__restrict__ rot) {
    extern __shared__ double r_s[];
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x)
        r_s[idx] = rot[idx];
    __syncthreads();
    double x[ROT_MAX];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int i = 0; i < ROT_MAX; i++)
            if (i < m)
                x[i] = X[(size_t)i * ld + e];
        for (int j = 0; j < m; j++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < ROT_MAX; i++)
                if (i < m)
                    s = fma(r_s[j * m + i], x[i], s);
            X[(size_t)j * ld + e] = s;
        }
    }
}

// Cyclic Jacobi for a symmetric m x m matrix (row-major, full). On exit w ascending,
// v(:,k) (column k of row-major v) the k-th eigenvector. m <= 64.
void jacobi_eigh(int m, std::vector<double> &a, std::vector<double> &w, std::vector<double> &v) {
    v.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++)
        v[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++)
                (i == j ? dia : off) += a[(size_t)i * m + j] * a[(size_t)i * m + j];
        if (off <= 1e-32 * (dia + off) || off == 0.0)
            break;
        for (int p = 0; p < m - 1; p++)
            for (int q = p + 1; q < m; q++) {
                const double apq = a[(size_t)p * m + q];
                if (apq == 0.0)
                    continue;
                const double app = a[(size_t)p * m + p], aqq = a[(size_t)q * m + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; k++) {
                    const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
                    a[(size_t)k * m + p] = c * akp - s * akq;
                    a[(size_t)k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; k++) {
                    const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
                    a[(size_t)p * m + k] = c * apk - s * aqk;
                    a[(size_t)q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; k++) {
                    const double vkp = v[(size_t)k * m + p], vkq = v[(size_t)k * m + q];
                    v[(size_t)k * m + p] = c * vkp - s * vkq;
                    v[(size_t)k * m + q] = s * vkp + c * vkq;
                }
            }
    }
    w.resize(m);
    std::vector<int> idx(m);
    for (int i = 0; i < m; i++)
        w[i] = a[(size_t)i * m + i], idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&w](int x, int y) { return w[x] < w[y]; });
    std::vector<double> w2(m), v2((size_t)m * m);
    for (int k = 0; k < m; k++) {
        w2[k] = w[idx[k]];
        for (int i = 0; i < m; i++)
            v2[(size_t)i * m + k] = v[(size_t)i * m + idx[k]];
    }
    w.swap(w2), v.swap(v2);
}

struct DevBuf {
    void *p = nullptr;
    b2g_context *ctx = nullptr;
    ~DevBuf() { b2g_dfree(ctx, p); }
};

} // namespace

extern "C" int b2g_davidson(b2g_plan *plan, const double *diag_host, double *ket_host, double conv_thrd,
                            double rel_conv_thrd, int max_iter, int soft_max_iter, int deflation_min_size,
                            int deflation_max_size, double *eigenvalue, int *ndav) {
    if (!plan || !diag_host || !ket_host || !eigenvalue || !ndav) {
        b2g_set_error("b2g_davidson: null argument");
        return 1;
    }
    if (plan->csize != plan->vsize) {
        b2g_set_error("b2g_davidson: H_eff must be square");
        return 1;
    }
    b2g_context *ctx = plan->ctx;
    B2G_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = plan->csize;
    const bool prof = b2g_prof_enabled() != 0;
    double prof_t = B2GProfScope::now();
    auto prof_lap = [&prof_t, prof](const char *label) {
        if (prof) {
            const double now = B2GProfScope::now();
            b2g_prof_record(label, now - prof_t);
            prof_t = now;
        }
    };
    std::vector<cudaEvent_t> prof_ev; // profile: pairs of events around every matvec (GPU time of the H.c products)
    const int k = 1;
    if (deflation_min_size < k)
        deflation_min_size = k;
    if (deflation_max_size < k + k / 2)
        deflation_max_size = k + k / 2;
    if (deflation_max_size > 62) { // 62*63/2 pairs fit the 8 x 256 Gram accumulators
        b2g_set_error("b2g_davidson: deflation_max_size > 62 not supported");
        return 1;
    }
    const int M = deflation_max_size;
    const int64_t ld = (n + 1) & ~(int64_t)1;
    DevBuf d_bs, d_ss, d_q, d_t, d_aa, d_scal, d_part, d_cnt, d_gpart, d_gram, d_rot;
    d_bs.ctx = ctx; if (b2g_dmalloc(ctx, &d_bs.p, sizeof(double) * ld * M)) return 1;
    d_ss.ctx = ctx; if (b2g_dmalloc(ctx, &d_ss.p, sizeof(double) * ld * M)) return 1;
    d_q.ctx = ctx; if (b2g_dmalloc(ctx, &d_q.p, sizeof(double) * ld)) return 1;
    d_t.ctx = ctx; if (b2g_dmalloc(ctx, &d_t.p, sizeof(double) * ld)) return 1;
    d_aa.ctx = ctx; if (b2g_dmalloc(ctx, &d_aa.p, sizeof(double) * ld)) return 1;
    d_scal.ctx = ctx; if (b2g_dmalloc(ctx, &d_scal.p, sizeof(double) * 8)) return 1;
    d_part.ctx = ctx; if (b2g_dmalloc(ctx, &d_part.p, sizeof(double) * RED_BLOCKS)) return 1;
    d_cnt.ctx = ctx; if (b2g_dmalloc(ctx, &d_cnt.p, sizeof(unsigned int))) return 1;
    const int gram_blocks = ctx->sm_count * 2;
    d_gpart.ctx = ctx; if (b2g_dmalloc(ctx, &d_gpart.p, sizeof(double) * gram_blocks * (M * (M + 1) / 2))) return 1;
    d_gram.ctx = ctx; if (b2g_dmalloc(ctx, &d_gram.p, sizeof(double) * M * M)) return 1;
    d_rot.ctx = ctx; if (b2g_dmalloc(ctx, &d_rot.p, sizeof(double) * M * M)) return 1;
    double *bs = (double *)d_bs.p, *ss = (double *)d_ss.p, *q = (double *)d_q.p, *t = (double *)d_t.p,
           *aa = (double *)d_aa.p, *scal = (double *)d_scal.p, *part = (double *)d_part.p;
    unsigned int *cnt = (unsigned int *)d_cnt.p;
    B2G_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), st));
    B2G_CUDA(cudaMemcpyAsync(aa, diag_host, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    B2G_CUDA(cudaMemcpyAsync(bs, ket_host, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    B2G_CUDA(cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * M * GRAM_CHUNK * sizeof(double))));
    const int ew_grid = ctx->sm_count * 8, ew_thr = 256;
    auto dot = [&](const double *x, const double *y, double *out) {
        dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(x, y, n, part, cnt, out);
        ctx->launches++;
    };
    prof_lap("davidson.alloc+h2d_issue");
    // normalise the initial guess (:945-955)
    dot(bs, bs, scal + 0);
    double h_scal[8];
    B2G_CUDA(cudaMemcpyAsync(h_scal, scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    B2G_CUDA(cudaStreamSynchronize(st));
    if (!(fabs(h_scal[0]) >= 1E-14)) {
        b2g_set_error("b2g_davidson: initial guess has zero norm");
        return 1;
    }
    scale_rsqrt_kernel<<<ew_grid, ew_thr, 0, st>>>(bs, n, scal + 0);
    ctx->launches++;
    prof_lap("davidson.init_sync");

    int m = k, msig = 0, xiter = 0, ck = 0;
    double theta = 0.0, qq = 0.0;
    std::vector<double> h_gram, w, vmat, h_rot;
    bool converged = false;
    while (xiter < max_iter && (soft_max_iter == -1 || xiter < soft_max_iter)) {
        xiter++;
        for (int i = msig; i < m; i++, msig++) {
            B2G_CUDA(cudaMemsetAsync(ss + (size_t)i * ld, 0, sizeof(double) * n, st));
            if (prof) {
                cudaEvent_t e0, e1;
                B2G_CUDA(cudaEventCreate(&e0));
                B2G_CUDA(cudaEventCreate(&e1));
                prof_ev.push_back(e0), prof_ev.push_back(e1);
                B2G_CUDA(cudaEventRecord(e0, st));
            }
            if (b2g_launch_matvec(plan, bs + (size_t)i * ld, ss + (size_t)i * ld, 1.0))
                return 1;
            if (prof)
                B2G_CUDA(cudaEventRecord(prof_ev.back(), st));
            if (ctx->nccl_comm && b2g_allreduce_sum(ctx, ss + (size_t)i * ld, n))
                return 1;
        }
        // subspace matrix, lower triangle
        const int npair = m * (m + 1) / 2;
        gram_partial_kernel<<<gram_blocks, 256, 2 * m * GRAM_CHUNK * sizeof(double), st>>>(bs, ss, m, n, ld,
                                                                                            (double *)d_gpart.p);
        gram_reduce_kernel<<<(npair + 127) / 128, 128, 0, st>>>((double *)d_gpart.p, npair, gram_blocks,
                                                                 (double *)d_gram.p);
        ctx->launches += 2;
        h_gram.resize(npair);
        B2G_CUDA(cudaMemcpyAsync(h_gram.data(), d_gram.p, sizeof(double) * npair, cudaMemcpyDeviceToHost, st));
        B2G_CUDA(cudaStreamSynchronize(st));
        std::vector<double> amat((size_t)m * m);
        for (int i = 0; i < m; i++)
            for (int j = 0; j <= i; j++)
                amat[(size_t)i * m + j] = amat[(size_t)j * m + i] = h_gram[i * (i + 1) / 2 + j];
        jacobi_eigh(m, amat, w, vmat);
        // rot(j, i) = component i of eigenvector j ("alpha row/column is diff from python", :1004)
        h_rot.resize((size_t)m * m);
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++)
                h_rot[(size_t)j * m + i] = vmat[(size_t)i * m + j];
        B2G_CUDA(cudaMemcpyAsync(d_rot.p, h_rot.data(), sizeof(double) * m * m, cudaMemcpyHostToDevice, st));
        if (m > 1) {
            rotate_kernel<<<ctx->sm_count * 4, 128, m * m * sizeof(double), st>>>(ss, m, n, ld, (double *)d_rot.p);
            rotate_kernel<<<ctx->sm_count * 4, 128, m * m * sizeof(double), st>>>(bs, m, n, ld, (double *)d_rot.p);
            ctx->launches += 2;
        } else if (h_rot[0] < 0) { // 1 x 1 "eigenvector" is +-1; keep the sign convention harmless
            h_rot[0] = 1.0;
        }
        theta = w[0];
        // residual of the lowest Ritz pair
        residual_kernel<<<ew_grid, ew_thr, 0, st>>>(q, ss, bs, theta, n);
        ctx->launches++;
        dot(q, q, scal + 1);
        B2G_CUDA(cudaMemcpyAsync(h_scal, scal + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        // Olsen preconditioner: q = Kinv q - (b, Kinv q) / (b, Kinv b) Kinv b
        olsen_divide_kernel<<<ew_grid, ew_thr, 0, st>>>(q, t, bs, aa, theta, n);
        ctx->launches++;
        dot(bs, q, scal + 2);
        dot(bs, t, scal + 3);
        axpy_dev_kernel<<<ew_grid, ew_thr, 0, st>>>(q, t, n, scal + 2, scal + 3, -1.0);
        ctx->launches++;
        B2G_CUDA(cudaStreamSynchronize(st));
        qq = h_scal[0];
        if (fabs(qq) < conv_thrd + fabs(theta) * fabs(theta) * rel_conv_thrd * rel_conv_thrd && m >= k) {
            ck++;
            converged = true;
            break;
        }
        if (m >= deflation_max_size)
            m = msig = deflation_min_size;
        for (int j = 0; j < m; j++) {
            dot(bs + (size_t)j * ld, q, scal + 4);
            axpy_dev_kernel<<<ew_grid, ew_thr, 0, st>>>(q, bs + (size_t)j * ld, n, scal + 4, nullptr, -1.0);
            ctx->launches++;
        }
        dot(q, q, scal + 5);
        scale_rsqrt_kernel<<<ew_grid, ew_thr, 0, st>>>(q, n, scal + 5);
        ctx->launches++;
        B2G_CUDA(cudaMemcpyAsync(bs + (size_t)m * ld, q, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        m++;
        if (xiter == soft_max_iter)
            break;
    }
    if (!converged && xiter == max_iter) {
        b2g_set_error("b2g_davidson: not converged within max_iter");
        return 4;
    }
    prof_lap("davidson.iterations");
    B2G_CUDA(cudaMemcpyAsync(ket_host, bs, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    B2G_CUDA(cudaStreamSynchronize(st));
    B2G_CUDA(cudaGetLastError());
    prof_lap("davidson.result_d2h");
    if (prof) {
        double gpu_s = 0, first_s = 0;
        for (size_t i = 0; i + 1 < prof_ev.size(); i += 2) {
            float ms = 0;
            cudaEventElapsedTime(&ms, prof_ev[i], prof_ev[i + 1]);
            gpu_s += ms * 1e-3;
            if (i == 0)
                first_s = ms * 1e-3;
            cudaEventDestroy(prof_ev[i]), cudaEventDestroy(prof_ev[i + 1]);
        }
        b2g_prof_record("davidson.matvec_gpu", gpu_s);
        b2g_prof_record("davidson.matvec_gpu_first", first_s);
        fprintf(stderr,
                "[b2g] davidson n=%lld pairs=%lld iters=%d matvec_gpu=%.1f ms (first %.1f ms) gflop/matvec=%.1f -> %.2f TFLOP/s\n",
                (long long)n, (long long)plan->npairs, xiter, gpu_s * 1e3, first_s * 1e3,
                2e-9 * (double)plan->stats.nflop_mnk,
                gpu_s > 0 ? 2e-12 * (double)plan->stats.nflop_mnk * (double)(prof_ev.size() / 2) / gpu_s : 0.0);
    }
    *eigenvalue = theta;
    *ndav = xiter;
    return 0;
}
