// faithful.cu -- the "faithful" engine: sequential-in-time kernels that follow
// the reference's operation order exactly, for ANY StateMatrix (ring or
// overlap).  One CTA per channel, one thread per state (or a strided loop when
// nstates > blockDim).  Used for: short sequences, the trellis-on-request form
// (T1/T2 of src/viterbi.jl:52-53), overlap models, the dense forward/backward
// API, the exact prologue of the ring engine, and as the bit-exact reference
// mode.  All FP64 arithmetic goes through the _rn intrinsics so that nvcc can
// never contract a multiply-add the reference rounds twice.
#include <cfloat>

#include "engines.h"
#include "faithful_dev.cuh"

namespace hmm {

// --------------------------------------------------------------------------
// device model blob (one per channel, fixed layout for a given topology)
// --------------------------------------------------------------------------
FaithfulLayout faithful_layout(int nstates, int64_t ntrans) {
    FaithfulLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 15) & ~size_t(15);
        return r;
    };
    L.scal = take(8 * sizeof(double));
    L.m = take(sizeof(double) * nstates);
    L.in_lp = take(sizeof(double) * ntrans);
    L.out_lp = take(sizeof(double) * ntrans);
    L.in_ptr = take(sizeof(int) * (nstates + 1));
    L.in_src = take(sizeof(int) * ntrans);
    L.out_ptr = take(sizeof(int) * (nstates + 1));
    L.out_dst = take(sizeof(int) * ntrans);
    L.dec_slot = take(sizeof(int) * nstates);
    L.static_pred = take(sizeof(int) * nstates);
    L.bytes = o;
    return L;
}

void faithful_pack(const HostModel &M, const FaithfulLayout &L, char *dst) {
    double *sc = (double *)(dst + L.scal);
    const double LOG2PI = 0.9189385332046727;  // src/utils.jl:1
    sc[0] = M.sigma;
    sc[1] = M.lsig;
    sc[2] = (-LOG2PI) - M.lsig;          // -log2pi - l_sigma, src/utils.jl:3-4
    sc[3] = 2 * (M.sigma * M.sigma);      // 2*sigma2
    memcpy(dst + L.m, M.m.data(), sizeof(double) * M.nstates);
    memcpy(dst + L.in_lp, M.in_lp.data(), sizeof(double) * M.ntrans);
    memcpy(dst + L.out_lp, M.out_lp.data(), sizeof(double) * M.ntrans);
    memcpy(dst + L.in_ptr, M.in_ptr.data(), sizeof(int) * (M.nstates + 1));
    memcpy(dst + L.in_src, M.in_src.data(), sizeof(int) * M.ntrans);
    memcpy(dst + L.out_ptr, M.out_ptr.data(), sizeof(int) * (M.nstates + 1));
    memcpy(dst + L.out_dst, M.out_dst.data(), sizeof(int) * M.ntrans);
    memcpy(dst + L.dec_slot, M.dec_slot.data(), sizeof(int) * M.nstates);
    memcpy(dst + L.static_pred, M.static_pred.data(), sizeof(int) * M.nstates);
}


// src/utils.jl:24-32
__device__ __forceinline__ double logsumexpl_dev(double xp, double yp) {
    if (xp > yp) return __dadd_rn(xp, log1p(exp(__dsub_rn(yp, xp))));
    return __dadd_rn(yp, log1p(exp(__dsub_rn(xp, yp))));
}

constexpr int YTILE = 512;

struct FaithSmem {
    double *col0, *col1, *m, *lp, *ytile;
    int *ptr, *idx, *dec;
};

__device__ __forceinline__ FaithSmem carve(char *base, int ns, int nt) {
    FaithSmem s;
    double *d = (double *)base;
    s.col0 = d; d += ns;
    s.col1 = d; d += ns;
    s.m = d; d += ns;
    s.lp = d; d += nt;
    s.ytile = d; d += YTILE;
    int *i = (int *)d;
    s.ptr = i; i += ns + 1;
    s.idx = i; i += nt;
    s.dec = i;
    return s;
}

size_t faithful_smem_bytes(int ns, int64_t nt) {
    return sizeof(double) * (3 * (size_t)ns + nt + YTILE) + sizeof(int) * ((size_t)ns + 1 + nt + ns) + 16;
}

// --------------------------------------------------------------------------
// Viterbi forward, src/viterbi.jl:55-88.  Writes, per step, the backpointers of
// the multi-predecessor ("decision") states only; optionally the dense T1/T2
// columns for steps < trellis_cols.
// --------------------------------------------------------------------------
__global__ void faithful_viterbi_fwd(const double *__restrict__ y, int64_t T, int64_t y_stride, const char *blob,
                                     size_t blob_stride, FaithfulLayout L, int ns, int nt, int ndec,
                                     int16_t *__restrict__ decbp, int16_t *__restrict__ xlast,
                                     double *__restrict__ T1_out, int16_t *__restrict__ T2_out, int64_t trellis_cols,
                                     double *__restrict__ final_col) {
    extern __shared__ __align__(16) char smem_raw[];
    const int c = blockIdx.x;
    const char *mb = blob + (size_t)c * blob_stride;
    const double *sc = (const double *)(mb + L.scal);
    const double c_emit = sc[2], two_s2 = sc[3];
    FaithSmem S = carve(smem_raw, ns, nt);
    {
        const double *gm = (const double *)(mb + L.m), *glp = (const double *)(mb + L.in_lp);
        const int *gp = (const int *)(mb + L.in_ptr), *gs = (const int *)(mb + L.in_src),
                  *gd = (const int *)(mb + L.dec_slot);
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            S.m[i] = gm[i];
            S.dec[i] = gd[i];
        }
        for (int i = threadIdx.x; i <= ns; i += blockDim.x) S.ptr[i] = gp[i];
        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
            S.lp[i] = glp[i];
            S.idx[i] = gs[i];
        }
    }
    y += (size_t)c * y_stride;
    decbp += (size_t)c * (size_t)T * ndec;
    if (T1_out) T1_out += (size_t)c * (size_t)trellis_cols * ns;
    if (T2_out) T2_out += (size_t)c * (size_t)trellis_cols * ns;
    __syncthreads();
    double *prev = S.col0, *cur = S.col1;
    // column 1 (src/viterbi.jl:55-63): emissions only, noise forced to 0
    const double y0 = y[0];
    for (int j = threadIdx.x; j < ns; j += blockDim.x) {
        double v = emit_rn(y0, S.m[j], c_emit, two_s2);
        if (j == 0) v = 0.0;
        prev[j] = v;
        if (T1_out && trellis_cols > 0) T1_out[j] = v;
        if (T2_out && trellis_cols > 0) T2_out[j] = 1;
    }
    __syncthreads();
    for (int64_t t0 = 1; t0 < T; t0 += YTILE) {
        int n = (int)((T - t0 < YTILE) ? (T - t0) : YTILE);
        for (int k = threadIdx.x; k < n; k += blockDim.x) S.ytile[k] = y[t0 + k];
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const int64_t t = t0 + k;
            const double yv = S.ytile[k];
            for (int j = threadIdx.x; j < ns; j += blockDim.x) {
                double q = emit_rn(yv, S.m[j], c_emit, two_s2);
                double best = -INFINITY;
                int bp = 0;  // T2 default 1 (src/viterbi.jl:53)
                const int e1 = S.ptr[j + 1];
                for (int e = S.ptr[j]; e < e1; e++) {
                    int k2 = S.idx[e];
                    double tt = __dadd_rn(prev[k2], S.lp[e]);
                    if (tt > best) {  // strict: first candidate in list order wins ties
                        best = tt;
                        bp = k2;
                    }
                }
                double v = __dadd_rn(best, q);
                cur[j] = v;
                int slot = S.dec[j];
                if (slot >= 0) decbp[(size_t)t * ndec + slot] = (int16_t)bp;
                if (t < trellis_cols) {
                    if (T1_out) T1_out[(size_t)t * ns + j] = v;
                    if (T2_out) T2_out[(size_t)t * ns + j] = (int16_t)(bp + 1);
                }
            }
            __syncthreads();
            double *tmp = prev;
            prev = cur;
            cur = tmp;
        }
    }
    // x[T] = argmax(T1[:,T]) -- first maximum (src/viterbi.jl:90)
    if (final_col)
        for (int j = threadIdx.x; j < ns; j += blockDim.x) final_col[(size_t)c * ns + j] = prev[j];
    if (threadIdx.x == 0 && xlast) {
        int best = 0;
        double bv = prev[0];
        for (int j = 1; j < ns; j++)
            if (prev[j] > bv) {
                bv = prev[j];
                best = j;
            }
        xlast[c] = (int16_t)best;
    }
}

// --------------------------------------------------------------------------
// Backtrack, src/viterbi.jl:93-95: x[i-1] = T2[x[i], i]
// --------------------------------------------------------------------------
__global__ void faithful_backtrack(int64_t T, const char *blob, size_t blob_stride, FaithfulLayout L, int ns, int ndec,
                                   const int16_t *__restrict__ decbp, const int16_t *__restrict__ xlast,
                                   int16_t *__restrict__ x, int64_t x_stride, int tile) {
    extern __shared__ __align__(16) char smem_raw[];
    const int c = blockIdx.x;
    const char *mb = blob + (size_t)c * blob_stride;
    const int *gd = (const int *)(mb + L.dec_slot), *gsp = (const int *)(mb + L.static_pred);
    int16_t *bpt = (int16_t *)smem_raw;             // [tile * ndec]
    int16_t *xt = bpt + (size_t)tile * ndec;         // [tile]
    int16_t *dec = xt + tile;                        // [ns]
    int16_t *sp = dec + ns;                          // [ns]
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        dec[i] = (int16_t)gd[i];
        sp[i] = (int16_t)gsp[i];
    }
    decbp += (size_t)c * (size_t)T * ndec;
    x += (size_t)c * x_stride;
    __shared__ int cur_s;
    if (threadIdx.x == 0) {
        cur_s = xlast[c];
        x[T - 1] = (int16_t)(cur_s + 1);
    }
    __syncthreads();
    // steps t in [a, b), a >= 1; resolves x[a-1 .. b-2]
    for (int64_t b = T; b > 1; b -= tile) {
        int64_t a = b - tile;
        if (a < 1) a = 1;
        int n = (int)(b - a);
        for (size_t i = threadIdx.x; i < (size_t)n * ndec; i += blockDim.x) bpt[i] = decbp[(size_t)a * ndec + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int cur = cur_s;
            for (int k = n - 1; k >= 0; k--) {
                int slot = dec[cur];
                cur = slot >= 0 ? bpt[(size_t)k * ndec + slot] : sp[cur];
                xt[k] = (int16_t)(cur + 1);
            }
            cur_s = cur;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) x[a - 1 + k] = xt[k];
        __syncthreads();
    }
}

// --------------------------------------------------------------------------
// ll = sum_{i=T..2} T1[x[i], i] (src/viterbi.jl:92-96).  The winning
// candidate is stored verbatim by the forward sweep, so along the decoded path
// T1[x_i, i] = (T1[x_{i-1}, i-1] + lp(x_{i-1} -> x_i)) + q_i(x_i) with the
// reference's own rounding; it is re-accumulated here sequentially, then summed
// in the reference's (descending) order.
// --------------------------------------------------------------------------
constexpr int PTILE = 2048;

__global__ void faithful_path_score(const double *__restrict__ y, int64_t T, int64_t y_stride, const char *blob,
                                    size_t blob_stride, FaithfulLayout L, int ns, int nt,
                                    const int16_t *__restrict__ x, int64_t x_stride, double *__restrict__ pscore,
                                    double *__restrict__ ll_out) {
    extern __shared__ __align__(16) char smem_raw[];
    const int c = blockIdx.x;
    const char *mb = blob + (size_t)c * blob_stride;
    const double *sc = (const double *)(mb + L.scal);
    const double c_emit = sc[2], two_s2 = sc[3];
    const double *gm = (const double *)(mb + L.m), *glp = (const double *)(mb + L.in_lp);
    const int *gp = (const int *)(mb + L.in_ptr), *gs = (const int *)(mb + L.in_src);
    double *w = (double *)smem_raw;  // [PTILE] transition weight
    double *q = w + PTILE;           // [PTILE] emission
    double *ps = q + PTILE;          // [PTILE]
    y += (size_t)c * y_stride;
    x += (size_t)c * x_stride;
    pscore += (size_t)c * (size_t)T;
    __shared__ double carry;
    if (threadIdx.x == 0) {
        int x0 = x[0] - 1;
        carry = x0 == 0 ? 0.0 : emit_rn(y[0], gm[x0], c_emit, two_s2);
        pscore[0] = carry;
    }
    __syncthreads();
    for (int64_t t0 = 1; t0 < T; t0 += PTILE) {
        int n = (int)((T - t0 < PTILE) ? (T - t0) : PTILE);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            int64_t t = t0 + k;
            int d = x[t] - 1, s = x[t - 1] - 1;
            double lp = __longlong_as_double(0x7ff8000000000000LL);
            for (int e = gp[d]; e < gp[d + 1]; e++)
                if (gs[e] == s) {
                    lp = glp[e];
                    break;
                }
            w[k] = lp;
            q[k] = emit_rn(y[t], gm[d], c_emit, two_s2);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = carry;
            for (int k = 0; k < n; k++) {
                s = __dadd_rn(__dadd_rn(s, w[k]), q[k]);
                ps[k] = s;
            }
            carry = s;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) pscore[t0 + k] = ps[k];
        __syncthreads();
    }
    // descending sum i = T .. 2
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int64_t b = T; b > 1; b -= PTILE) {
        int64_t a = b - PTILE;
        if (a < 1) a = 1;
        int n = (int)(b - a);
        for (int k = threadIdx.x; k < n; k += blockDim.x) ps[k] = pscore[a + k];
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = carry;
            for (int k = n - 1; k >= 0; k--) s = __dadd_rn(s, ps[k]);
            carry = s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ll_out[c] = carry;
}

// --------------------------------------------------------------------------
// forward / backward with dense output, src/baumwelch.jl:25-51, 73-98
// --------------------------------------------------------------------------
template <bool BACKWARD>
__global__ void faithful_fb(const double *__restrict__ V, int64_t T, const char *blob, FaithfulLayout L, int ns, int nt,
                            double *__restrict__ out) {
    extern __shared__ __align__(16) char smem_raw[];
    const char *mb = blob;
    const double *sc = (const double *)(mb + L.scal);
    const double c_emit = sc[2], two_s2 = sc[3];
    FaithSmem S = carve(smem_raw, ns, nt);
    {
        const double *gm = (const double *)(mb + L.m);
        const double *glp = (const double *)(mb + (BACKWARD ? L.out_lp : L.in_lp));
        const int *gp = (const int *)(mb + (BACKWARD ? L.out_ptr : L.in_ptr));
        const int *gs = (const int *)(mb + (BACKWARD ? L.out_dst : L.in_src));
        for (int i = threadIdx.x; i < ns; i += blockDim.x) S.m[i] = gm[i];
        for (int i = threadIdx.x; i <= ns; i += blockDim.x) S.ptr[i] = gp[i];
        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
            S.lp[i] = glp[i];
            S.idx[i] = gs[i];
        }
    }
    __syncthreads();
    double *prev = S.col0, *cur = S.col1;
    if (!BACKWARD) {
        const double v0 = V[0];
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            double v = emit_rn(v0, S.m[j], c_emit, two_s2);  // :36 (pi overwritten)
            prev[j] = v;
            out[j] = v;
        }
    } else {
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            prev[j] = 0.0;  // :80
            out[(size_t)(T - 1) * ns + j] = 0.0;
        }
    }
    __syncthreads();
    for (int64_t base = 1; base < T; base += YTILE) {
        int n = (int)((T - base < YTILE) ? (T - base) : YTILE);
        // forward consumes V[base + k]; backward step k computes column T-1-(base+k) from V[T-(base+k)]
        for (int k = threadIdx.x; k < n; k += blockDim.x) S.ytile[k] = BACKWARD ? V[T - (base + k)] : V[base + k];
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const double v = S.ytile[k];
            const int64_t col = BACKWARD ? (T - 1 - (base + k)) : (base + k);
            for (int j = threadIdx.x; j < ns; j += blockDim.x) {
                double a = -INFINITY;
                const int e1 = S.ptr[j + 1];
                double bj = 0;
                if (!BACKWARD) bj = emit_rn(v, S.m[j], c_emit, two_s2);
                for (int e = S.ptr[j]; e < e1; e++) {
                    int k2 = S.idx[e];
                    if (BACKWARD) bj = emit_rn(v, S.m[k2], c_emit, two_s2);
                    a = logsumexpl_dev(a, __dadd_rn(__dadd_rn(prev[k2], S.lp[e]), bj));  // :47 / :94
                }
                cur[j] = a;
                out[(size_t)col * ns + j] = a;
            }
            __syncthreads();
            double *tmp = prev;
            prev = cur;
            cur = tmp;
        }
    }
}

// --------------------------------------------------------------------------
// host drivers
// --------------------------------------------------------------------------
static int pick_threads(int ns) {
    int t = ((ns + 31) / 32) * 32;
    if (t > 1024) t = 1024;
    if (t < 32) t = 32;
    return t;
}

void faithful_viterbi_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const FaithfulLayout &L,
                          const char *blob_dev, const HostModel &M0, int16_t *x_dev, int64_t x_stride,
                          double *ll_dev /*[C] or null*/, double *T1_dev, int16_t *T2_dev, int64_t trellis_cols,
                          bool forward_only, double *final_col_dev, cudaStream_t st, hmm_info *info) {
    Workspace &ws = workspace();
    const int ns = M0.nstates, nt = (int)M0.ntrans, ndec = M0.ndec > 0 ? M0.ndec : 1;
    int16_t *decbp = (int16_t *)ws.get(Workspace::DEC, sizeof(int16_t) * (size_t)C * T * ndec);
    int16_t *xlast = (int16_t *)ws.get(Workspace::MISC, sizeof(int16_t) * C + 64);
    size_t sm = faithful_smem_bytes(ns, nt);
    if (sm > 227 * 1024) fail(HMM_EUNSUPPORTED, "model too large for the faithful engine (%zu B shared memory)", sm);
    HMM_CUDA(cudaFuncSetAttribute(faithful_viterbi_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    faithful_viterbi_fwd<<<C, pick_threads(ns), sm, st>>>(y_dev, T, y_stride, blob_dev, L.bytes, L, ns, nt, ndec, decbp,
                                                          xlast, T1_dev, T2_dev, trellis_cols, final_col_dev);
    HMM_CUDA(cudaGetLastError());
    if (info) info->kernel_launches++;
    if (forward_only) return;
    int tile = 4096;
    while (tile > 64 && (size_t)tile * (ndec + 1) * 2 + (size_t)ns * 4 > 96 * 1024) tile /= 2;
    size_t sm2 = (size_t)tile * (ndec + 1) * 2 + (size_t)ns * 4 + 16;
    HMM_CUDA(cudaFuncSetAttribute(faithful_backtrack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    faithful_backtrack<<<C, 256, sm2, st>>>(T, blob_dev, L.bytes, L, ns, ndec, decbp, xlast, x_dev, x_stride, tile);
    HMM_CUDA(cudaGetLastError());
    if (info) info->kernel_launches++;
    if (ll_dev) path_score_run(y_dev, T, y_stride, C, L, blob_dev, M0, x_dev, x_stride, ll_dev, st, info);
}

void path_score_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const FaithfulLayout &L,
                    const char *blob_dev, const HostModel &M0, const int16_t *x_dev, int64_t x_stride, double *ll_dev,
                    cudaStream_t st, hmm_info *info) {
    Workspace &ws = workspace();
    double *ps = (double *)ws.get(Workspace::PSCORE, sizeof(double) * (size_t)C * T);
    size_t sm = sizeof(double) * 3 * PTILE;
    HMM_CUDA(cudaFuncSetAttribute(faithful_path_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    faithful_path_score<<<C, 256, sm, st>>>(y_dev, T, y_stride, blob_dev, L.bytes, L, M0.nstates, (int)M0.ntrans, x_dev,
                                            x_stride, ps, ll_dev);
    HMM_CUDA(cudaGetLastError());
    if (info) info->kernel_launches++;
}

void faithful_fb_run(bool backward, const double *V_dev, int64_t T, const FaithfulLayout &L, const char *blob_dev,
                     const HostModel &M, double *out_dev, cudaStream_t st) {
    const int ns = M.nstates, nt = (int)M.ntrans;
    size_t sm = faithful_smem_bytes(ns, nt);
    if (sm > 227 * 1024) fail(HMM_EUNSUPPORTED, "model too large for the faithful engine (%zu B shared memory)", sm);
    if (backward) {
        HMM_CUDA(cudaFuncSetAttribute(faithful_fb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        faithful_fb<true><<<1, pick_threads(ns), sm, st>>>(V_dev, T, blob_dev, L, ns, nt, out_dev);
    } else {
        HMM_CUDA(cudaFuncSetAttribute(faithful_fb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        faithful_fb<false><<<1, pick_threads(ns), sm, st>>>(V_dev, T, blob_dev, L, ns, nt, out_dev);
    }
    HMM_CUDA(cudaGetLastError());
}

}  // namespace hmm
