// reconstruct.cu -- reconstruct_signal (src/reconstruction.jl:1-9) and
// unroll_mlseq (src/extraction.jl:4-13): pure gathers, HBM-bound
// (2 B read + 8 B write per sample).  The per-state sum
// Y[i] = sum_j mu[states[j, x[i]], j] is folded on the host once per call into
// a state-mean table with the reference's own accumulation order, so the
// kernel is one table lookup per sample.
#include <algorithm>

#include "engines.h"

namespace hmm {

constexpr int REC_VEC = 8;  // samples per thread: one 16-byte load of x

template <bool SMEM_TABLE>
__global__ void __launch_bounds__(256)
    reconstruct_kernel(const int16_t *__restrict__ x, int64_t T, const double *__restrict__ m, int nstates,
                       double *__restrict__ Y, int *__restrict__ err) {
    extern __shared__ double tab[];
    if (SMEM_TABLE) {
        for (int i = threadIdx.x; i < nstates; i += blockDim.x) tab[i] = m[i];
        __syncthreads();
    }
    const double *t = SMEM_TABLE ? tab : m;
    const int64_t nvec = T / REC_VEC;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    int bad = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        int16_t xs[REC_VEC];
        if (aligned) {
            *reinterpret_cast<int4 *>(xs) = __ldg(reinterpret_cast<const int4 *>(x) + v);
        } else {
#pragma unroll
            for (int k = 0; k < REC_VEC; k++) xs[k] = x[v * REC_VEC + k];
        }
        double ys[REC_VEC];
#pragma unroll
        for (int k = 0; k < REC_VEC; k++) {
            int s = xs[k] - 1;
            if (s < 0 || s >= nstates) {
                bad = 1;
                s = 0;
            }
            ys[k] = SMEM_TABLE ? t[s] : __ldg(t + s);
        }
        if (aligned) {
#pragma unroll
            for (int k = 0; k < REC_VEC; k += 2)
                *reinterpret_cast<double2 *>(Y + v * REC_VEC + k) = make_double2(ys[k], ys[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < REC_VEC; k++) Y[v * REC_VEC + k] = ys[k];
        }
    }
    // tail
    for (int64_t i = nvec * REC_VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T;
         i += (int64_t)gridDim.x * blockDim.x) {
        int s = x[i] - 1;
        if (s < 0 || s >= nstates) {
            bad = 1;
            s = 0;
        }
        Y[i] = t[s];
    }
    if (bad) atomicExch(err, 1);
}

__global__ void __launch_bounds__(256)
    unroll_kernel(const int16_t *__restrict__ x, int64_t T, const int16_t *__restrict__ states, int N, int nstates,
                  int16_t *__restrict__ out, int *__restrict__ err) {
    int bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        int s = x[i] - 1;
        if (s < 0 || s >= nstates) {
            bad = 1;
            s = 0;
        }
        for (int j = 0; j < N; j++) out[j + (size_t)N * i] = __ldg(states + j + (size_t)N * s);
    }
    if (bad) atomicExch(err, 1);
}

static int grid_for(int64_t work_items, int threads) {
    int64_t g = (work_items + threads - 1) / threads;
    const int64_t cap = 148 * 8;  // a few resident CTAs per SM, grid-stride beyond
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static void check_flag(int *flag_dev, cudaStream_t st, const char *what) {
    int h = 0;
    HMM_CUDA(cudaMemcpyAsync(&h, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaStreamSynchronize(st));
    if (h) fail(HMM_EINVAL, "%s: state index outside 1..nstates", what);
}

void reconstruct_run(const int16_t *x_dev, int64_t T, const std::vector<double> &m, double *Y_dev, cudaStream_t st) {
    Workspace &ws = workspace();
    const int ns = (int)m.size();
    char *buf = (char *)ws.get(Workspace::MODEL, sizeof(double) * ns + 64);
    double *m_dev = (double *)buf;
    int *flag = (int *)(buf + sizeof(double) * ns + 16 - (sizeof(double) * ns) % 16);
    HMM_CUDA(cudaMemcpyAsync(m_dev, m.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    int grid = grid_for(T / REC_VEC + 1, 256);
    if (sizeof(double) * ns <= 48 * 1024)
        reconstruct_kernel<true><<<grid, 256, sizeof(double) * ns, st>>>(x_dev, T, m_dev, ns, Y_dev, flag);
    else
        reconstruct_kernel<false><<<grid, 256, 0, st>>>(x_dev, T, m_dev, ns, Y_dev, flag);
    HMM_CUDA(cudaGetLastError());
    check_flag(flag, st, "reconstruct_signal");
}

void unroll_run(const int16_t *x_dev, int64_t T, const int16_t *states_host, int N, int nstates, int16_t *out_dev,
                cudaStream_t st) {
    Workspace &ws = workspace();
    size_t sb = sizeof(int16_t) * (size_t)N * nstates;
    char *buf = (char *)ws.get(Workspace::MODEL, sb + 64);
    int *flag = (int *)(buf + ((sb + 15) & ~size_t(15)));
    HMM_CUDA(cudaMemcpyAsync(buf, states_host, sb, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    unroll_kernel<<<grid_for(T, 256), 256, 0, st>>>(x_dev, T, (const int16_t *)buf, N, nstates, out_dev, flag);
    HMM_CUDA(cudaGetLastError());
    check_flag(flag, st, "unroll_mlseq");
}

}  // namespace hmm


// ---------------------------------------------------------------------------
// Roofline denominators measured on the device the caller is using (bench.py reports them live instead of quoting
// constants): the FP64 fused-multiply-add issue rate (DFMA with one constant-bank operand, the form the FIR uses)
// and the device-memory copy bandwidth.
// ---------------------------------------------------------------------------
namespace hmm {
__constant__ double peak_cc[16];

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, double b, int iters) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = 1.0 + i + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(peak_cc[i], b, x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) copy_peak_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

void measure_peaks(double *gdfma_per_s, double *copy_gb_per_s, cudaStream_t st) {
    int dev = 0, sms = 148;
    HMM_CUDA(cudaGetDevice(&dev));
    HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 4096;
    const size_t nvec = (size_t)32 << 20;  // 512 MB read + 512 MB written per copy: far beyond L2
    double *buf = nullptr;
    HMM_CUDA(cudaMalloc((void **)&buf, 2 * nvec * sizeof(double2)));
    double h[16];
    for (int i = 0; i < 16; i++) h[i] = 1.0 + 1e-9 * i;
    HMM_CUDA(cudaMemcpyToSymbolAsync(peak_cc, h, sizeof h, 0, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemsetAsync(buf, 0, 2 * nvec * sizeof(double2), st));
    Timer t(st);
    fp64_peak_kernel<<<blocks, 256, 0, st>>>(buf, 0.999, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        t.start();
        fp64_peak_kernel<<<blocks, 256, 0, st>>>(buf, 0.999, iters);
        t.stop();
        best = std::min(best, t.ms());
    }
    *gdfma_per_s = (double)blocks * 256 * iters * 16 / (best * 1e-3) / 1e9;
    best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        t.start();
        copy_peak_kernel<<<sms * 16, 256, 0, st>>>((const double2 *)buf, (double2 *)buf + nvec, nvec);
        t.stop();
        best = std::min(best, t.ms());
    }
    *copy_gb_per_s = 2.0 * nvec * sizeof(double2) / (best * 1e-3) / 1e9;
    HMM_CUDA(cudaGetLastError());
    cudaFree(buf);
}
}  // namespace hmm
