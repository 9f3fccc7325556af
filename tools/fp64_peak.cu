// fp64_peak.cu -- measures the FP64 FMA issue rate of the GPU (the "issue roofline"
// SURVEY 8d asks for; it is not in MEASURED_PEAKS.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu && ./fp64_peak
// Three variants: DFMA with three register operands, DFMA with one constant-bank
// operand (what the FIR uses), DADD.
#include <cuda_runtime.h>
#include <cstdio>

__constant__ double cc[16];

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, double a, double b, int iters) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a + i + threadIdx.x * 1e-3;
    double m0 = b, m1 = b * 1.0001, m2 = b * 1.0002, m3 = b * 1.0003;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) x[i] = fma(x[i], (i & 1) ? m0 : m1, (i & 2) ? m2 : m3);
            if (MODE == 1) x[i] = fma(cc[i], m0, x[i]);
            if (MODE == 2) x[i] = x[i] + m0;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char *name, double *out) {
    const int blocks = 148 * 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 1.0, 0.999, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, 1.0, 0.999, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * iters * 16;
    double rate = ops / (best * 1e-3);
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"Gop_per_s\": %.1f}\n", name, best, rate / 1e9);
    return rate;
}

// DFMA (constant-bank operand, 24 independent accumulators like the FIR) at a given number of warps per SM
// sub-partition: how many FIR warps does it take to saturate the FP64 pipe?
__global__ void __launch_bounds__(1024) ksweep(double *out, double b, int iters) {
    double x[24];
#pragma unroll
    for (int i = 0; i < 24; i++) x[i] = 1.0 + i + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 24; i++) x[i] = fma(cc[i & 15], b, x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static void sweep(double *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int wps : {1, 2, 3, 4, 6, 8}) {   // warps per sub-partition: one CTA per SM of 4*wps warps
        const int threads = 128 * wps, iters = 4096;
        ksweep<<<148, threads>>>(out, 0.999, 64);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            ksweep<<<148, threads>>>(out, 0.999, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double rate = (double)148 * threads * iters * 24 / (best * 1e-3);
        printf("{\"variant\": \"dfma_const_%d_warps_per_subpartition\", \"ms\": %.4f, \"Gop_per_s\": %.1f}\n", wps, best,
               rate / 1e9);
    }
}

int main() {
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 256);
    double h[16];
    for (int i = 0; i < 16; i++) h[i] = 0.5 + 0.01 * i;
    cudaMemcpyToSymbol(cc, h, sizeof h);
    run<0>("dfma_rrr", out);
    run<1>("dfma_const_operand", out);
    run<2>("dadd", out);
    sweep(out);
    return 0;
}
