// tools/fp64_probe.cu — measures the FP64 issue ceilings of the device this runs on:
// DFMA (CUDA cores) and DMMA (mma.sync f64, shapes m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16).
// Prints one JSON line; bench.py / DESIGN.md use it as the FP64 roofline denominator
// next to a cuBLAS DGEMM timed from Python.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double *out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma884(double *out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma1688(double *out, int iters) {
    double c[NACC][4];
    double a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-3 + i;
    b[0] = 1.0 + threadIdx.x * 1e-4, b[1] = 0.5;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double *out, int iters) {
    double c[NACC][4];
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    f();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    for (int warps : {4, 8, 16, 32}) {
        int bpsm = 1, thr = warps * 32; if (thr > 1024) { bpsm = thr / 1024; thr = 1024; }
        int grid = sms * bpsm;
        float ms = time_ms([&] { k_dfma<<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 16 * iters * (double)grid * thr;
        printf(", \"dfma_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        int thr = warps * 32, grid = sms;
        float ms = time_ms([&] { k_dmma884<8><<<grid, thr>>>(out, iters); });
        double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * warps;
        printf(", \"dmma884_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<4><<<grid, thr>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 8 * 4 * iters * (double)grid * warps;
        printf(", \"dmma1688_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<4><<<grid, thr>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)grid * warps;
        printf(", \"dmma16816_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
    }
    CK(cudaGetLastError());
    printf("}\n");
    return 0;
}
