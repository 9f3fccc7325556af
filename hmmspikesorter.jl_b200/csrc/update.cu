// update.cu -- update(alpha, beta, lA, mu, sigma, x) on DENSE alpha/beta
// (src/baumwelch.jl:205-309) for any StateMatrix.  This is the reference's
// stand-alone `update` entry point and the M-step of the generic (overlap-model)
// E/M path; ring models use the fused engine in ring_em.cu instead.
//
// Parallel over time: every CTA owns a contiguous range of samples, one thread
// per state; per sample two block-wide log-sum-exp reductions give
//   g_t = LSE_j(alpha+beta)              (:216-224, gamma = alpha+beta-g_t)
//   q_t = LSE_e(alpha[src,t]+lp+beta[dst,t+1]+b_dst(x[t+1]))   (:242-249)
// and the statistics sum e^gamma, sum e^gamma x, sum e^gamma x^2 per state and the
// xi sums for the transitions leaving state 1 are accumulated in shared memory,
// then written as per-CTA partials that the host adds in a fixed order.
#include <cmath>

#include "engines.h"

namespace hmm {

struct UpdParams {
    const double *alpha, *beta, *x;
    int64_t T;
    int ns, nt;
    const double *m;      // [ns] state means (old mu)
    const int *in_ptr;    // [ns+1]
    const int *in_src;    // [nt]
    const double *in_lp;  // [nt]
    double c_emit, two_s2;
    double *part;         // [nblk][4*ns + 2]: S0, S1, S2, xi(per dst state), bb, unused
    double *pp;           // [ns] gamma[:,0]
    int64_t per_block;
};

// LSE over the block of one value per thread (inactive threads pass -inf).
__device__ __forceinline__ double block_lse(double v, double *wm, double *wsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double m = v;
    for (int d = 16; d >= 1; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    double s = (v == -INFINITY) ? 0.0 : exp(v - m);
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
        wm[warp] = m;
        wsum[warp] = s;
    }
    __syncthreads();
    double M = -INFINITY;
    for (int w = 0; w < nw; w++) M = fmax(M, wm[w]);
    double S = 0.0;
    for (int w = 0; w < nw; w++)
        if (wm[w] != -INFINITY) S += wsum[w] * exp(wm[w] - M);
    __syncthreads();
    return M + log(S);
}

__global__ void dense_update_kernel(UpdParams p) {
    extern __shared__ __align__(16) double sm[];
    const int ns = p.ns;
    double *acol = sm;            // alpha column t
    double *S0 = acol + ns, *S1 = S0 + ns, *S2 = S1 + ns, *XI = S2 + ns;
    double *wm = XI + ns, *wsum = wm + 32;
    for (int j = threadIdx.x; j < 4 * ns; j += blockDim.x) S0[j] = 0.0;
    double bb = 0.0;
    const int64_t ta = (int64_t)blockIdx.x * p.per_block;
    int64_t tb = ta + p.per_block;
    if (tb > p.T) tb = p.T;
    __syncthreads();
    for (int64_t t = ta; t < tb; t++) {
        const double xt = p.x[t];
        const bool has_next = t + 1 < p.T;
        const double xn = has_next ? p.x[t + 1] : 0.0;
        // ---- g_t and gamma ----
        double loc = -INFINITY;
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            double a = p.alpha[(size_t)t * ns + j];
            acol[j] = a;
            double s = a + p.beta[(size_t)t * ns + j];
            loc = (loc == -INFINITY) ? s : ((s == -INFINITY) ? loc : fmax(loc, s) + log1p(exp(-fabs(loc - s))));
        }
        const double g = block_lse(loc, wm, wsum);  // also makes acol visible
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            double gam = acol[j] + p.beta[(size_t)t * ns + j] - g;
            double eg = exp(gam);
            S0[j] += eg;
            S1[j] += xt * eg;
            S2[j] += xt * xt * eg;
            if (t == 0) p.pp[j] = gam;
            if (j == 0 && has_next) bb += eg;
        }
        if (!has_next) continue;
        // ---- q_t over all transitions, xi for the ones leaving state 1 ----
        double locq = -INFINITY;
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            double dd = xn - p.m[j];
            double bj = p.c_emit - (dd * dd) / p.two_s2;
            double bn = p.beta[(size_t)(t + 1) * ns + j];
            for (int e = p.in_ptr[j]; e < p.in_ptr[j + 1]; e++) {
                double v = ((acol[p.in_src[e]] + p.in_lp[e]) + bn) + bj;
                locq = (locq == -INFINITY) ? v : ((v == -INFINITY) ? locq : fmax(locq, v) + log1p(exp(-fabs(locq - v))));
            }
        }
        const double q = block_lse(locq, wm, wsum);
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            for (int e = p.in_ptr[j]; e < p.in_ptr[j + 1]; e++)
                if (p.in_src[e] == 0) {
                    double dd = xn - p.m[j];
                    double bj = p.c_emit - (dd * dd) / p.two_s2;
                    double v = ((acol[0] + p.in_lp[e]) + p.beta[(size_t)(t + 1) * ns + j]) + bj;
                    XI[j] += exp(v - q);
                }
        }
        __syncthreads();
    }
    __syncthreads();
    double *dst = p.part + (size_t)blockIdx.x * (4 * ns + 2);
    for (int j = threadIdx.x; j < 4 * ns; j += blockDim.x) dst[j] = S0[j];
    if (threadIdx.x == 0) dst[4 * ns] = bb;
}

void dense_update_run(const double *alpha_dev, const double *beta_dev, const double *x_dev, int64_t T,
                      const HostModel &M, const int16_t *states, EmResult &out, cudaStream_t st) {
    Workspace &ws = workspace();
    const int ns = M.nstates, nt = (int)M.ntrans, N = M.N, K = M.K;
    int nblk = 148 * 4;
    int64_t per_block = (T + nblk - 1) / nblk;
    if (per_block < 8) per_block = 8;
    nblk = (int)((T + per_block - 1) / per_block);
    const size_t pstride = 4 * (size_t)ns + 2;
    size_t off = 0;
    auto carve = [&](size_t b) {
        size_t r = off;
        off += (b + 255) & ~size_t(255);
        return r;
    };
    size_t o_m = carve(sizeof(double) * ns), o_lp = carve(sizeof(double) * nt), o_ptr = carve(sizeof(int) * (ns + 1)),
           o_src = carve(sizeof(int) * nt), o_pp = carve(sizeof(double) * ns),
           o_part = carve(sizeof(double) * pstride * nblk);
    char *base = (char *)ws.get(Workspace::STATS, off);
    HMM_CUDA(cudaMemcpyAsync(base + o_m, M.m.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemcpyAsync(base + o_lp, M.in_lp.data(), sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemcpyAsync(base + o_ptr, M.in_ptr.data(), sizeof(int) * (ns + 1), cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemcpyAsync(base + o_src, M.in_src.data(), sizeof(int) * nt, cudaMemcpyHostToDevice, st));
    UpdParams p{};
    p.alpha = alpha_dev;
    p.beta = beta_dev;
    p.x = x_dev;
    p.T = T;
    p.ns = ns;
    p.nt = nt;
    p.m = (const double *)(base + o_m);
    p.in_lp = (const double *)(base + o_lp);
    p.in_ptr = (const int *)(base + o_ptr);
    p.in_src = (const int *)(base + o_src);
    const double LOG2PI = 0.9189385332046727;
    p.c_emit = (-LOG2PI) - M.lsig;
    p.two_s2 = 2 * (M.sigma * M.sigma);
    p.part = (double *)(base + o_part);
    p.pp = (double *)(base + o_pp);
    p.per_block = per_block;
    int threads = ((ns + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    if (threads < 64) threads = 64;
    size_t smb = sizeof(double) * (5 * (size_t)ns + 64);
    if (smb > 227 * 1024) fail(HMM_EUNSUPPORTED, "model too large for the dense update kernel");
    HMM_CUDA(cudaFuncSetAttribute(dense_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    dense_update_kernel<<<nblk, threads, smb, st>>>(p);
    HMM_CUDA(cudaGetLastError());
    std::vector<double> part(pstride * nblk), pp(ns), alast(ns);
    HMM_CUDA(cudaMemcpyAsync(part.data(), p.part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaMemcpyAsync(pp.data(), p.pp, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaMemcpyAsync(alast.data(), alpha_dev + (size_t)(T - 1) * ns, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaStreamSynchronize(st));
    {  // log-likelihood = LSE_j alpha[j,T], accumulated in state order with logsumexpl (src/utils.jl:24-32)
        double g = -INFINITY;
        for (int j = 0; j < ns; j++) {
            const double v = alast[j];
            if (g == -INFINITY) g = v;
            else if (v != -INFINITY) g = g > v ? g + std::log1p(std::exp(v - g)) : v + std::log1p(std::exp(g - v));
        }
        out.loglik = g;
    }
    // fixed-order host reduction of the per-CTA partials
    std::vector<double> S(4 * (size_t)ns, 0.0);
    double bb = 0.0;
    for (int b = 0; b < nblk; b++) {
        const double *q = part.data() + (size_t)b * pstride;
        for (size_t j = 0; j < 4 * (size_t)ns; j++) S[j] += q[j];
        bb += q[4 * ns];
    }
    const double *S0 = S.data(), *S1 = S0 + ns, *S2 = S1 + ns, *XI = S2 + ns;
    out.pp = pp;
    // new noise->active log-probabilities: xb[2:end] (src/baumwelch.jl:254-265), list order of
    // the transitions leaving state 1
    out.lp.clear();
    {
        // the out-edges of state 1 in list order (CSR by source keeps it): entry 0 is xb[1]
        // (normally noise -> noise), the rest are xb[2:end]
        const double lbb = std::log(bb);
        for (int e = M.out_ptr[0] + 1; e < M.out_ptr[1]; e++) out.lp.push_back(std::log(XI[M.out_dst[e]]) - lbb);
    }
    out.mu.assign((size_t)K * N, 0.0);  // fill!(mu, 0.0), :268
    std::vector<double> gg((size_t)K * N, 0.0);
    for (int j = 0; j < ns; j++) {
        int cnt = 0, l1 = -1;
        for (int l = 0; l < N; l++)
            if (states[l + (size_t)N * j] >= 2) {
                cnt++;
                l1 = l;
            }
        if (cnt == 1) {  // sidx: exactly one active neuron, :269
            int ss = states[l1 + (size_t)N * j] - 1;
            out.mu[ss + (size_t)K * l1] += S1[j];
            gg[ss + (size_t)K * l1] += S0[j];
        }
    }
    for (int l = 0; l < N; l++)
        for (int j = 1; j < K; j++) out.mu[j + (size_t)K * l] /= gg[j + (size_t)K * l];  // :283-287
    std::vector<double> mnew;
    state_means(states, N, K, ns, out.mu.data(), mnew);  // :288-293
    double x2 = 0.0, qq = 0.0;
    for (int j = 0; j < ns; j++) {
        // sum_t e^gamma (x_t - m_j)^2 = S2 - 2 m S1 + m^2 S0
        x2 += S2[j] - 2.0 * mnew[j] * S1[j] + mnew[j] * mnew[j] * S0[j];
        qq += S0[j];
    }
    out.sigma = std::sqrt(x2 / qq);  // :306-307
}

}  // namespace hmm
