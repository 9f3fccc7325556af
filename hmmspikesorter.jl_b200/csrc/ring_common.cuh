// ring_common.cuh -- device model and helpers shared by the ring engines.
//
// The non-overlap StateMatrix (src/types.jl:71-77,94-113) is one noise state
// plus N chains of L = K-1 states; inside a chain every state has exactly one
// predecessor, so a chain is a delay line.  Writing emissions relative to the
// noise emission makes them LINEAR in the sample:
//     q_t(i,s) - q_t(noise) = a[i][s] * y_t + b[i][s]
// and the score of running through chain i entered at t0 becomes
//     F_i(t0) = sum_{r<L} a[i][r] * y[t0+r] + Bc[i]          (an FIR filter)
// The per-step recursion then only involves the N+1 decision states
// (noise and the N chain heads):
//     Tail_t(i)  = P_{t-L+1}(i) + F_i(t-L+1)
//     G_t        = max_first{ G_{t-1}, Tail_{t-1}(j) + eG[j] }
//     P_t(i)     = max_first{ G_{t-1} + eH[i], Tail_{t-1}(j) + eT[j][i] (j != i) }
// with all scores normalised by the all-noise path.  (max -> sum of
// exponentials for forward/backward.)  See DESIGN.md section 3.
#pragma once
#include "engines.h"

namespace hmm {

constexpr int RING_MAX_N = 7;   // (N+1) decisions x 4 bits must fit one u32 per step
constexpr int RING_MAX_L = 96;   // K <= 97
constexpr int RING_Q = 128;      // ring-buffer length (>= L + 32), power of two

// Per-channel ring model, a flat array of doubles on the device.
struct RingLayout {
    int N, L, LP, NP;            // LP = L rounded up to 8; NP = N rounded up to even
    int A, BW, B0, BWsuf, Bc, eG, eH, eT, scal, cL, xG, xH, xT, hot, total;  // offsets in doubles
};
// scal[]: 0 w_nn, 1 c_emit, 2 two_s2, 3 m0, 4 sigma
__host__ __device__ inline RingLayout ring_layout(int N, int L) {
    RingLayout R;
    R.N = N;
    R.L = L;
    R.LP = (L + 7) & ~7;
    R.NP = (N + 1) & ~1;
    int o = 0;
    // hot prefix (copied to shared memory by the recursion kernels)
    R.A = o;  o += R.LP * R.NP;   // A[r*NP + i]
    R.Bc = o; o += R.NP;
    R.eG = o; o += R.NP;
    R.eH = o; o += R.NP;
    R.eT = o; o += N * R.NP;      // eT[j*NP + i]
    R.scal = o; o += 8;
    R.cL = o; o += R.NP;          // liveness margin per neuron (ring_viterbi.cu)
    R.xG = o; o += R.NP;          // exp(eG), exp(eH), exp(eT): linear-domain weights (ring_em.cu)
    R.xH = o; o += R.NP;
    R.xT = o; o += N * R.NP;
    R.hot = o;
    // cold part (boundary handling only; read from global memory)
    R.BW = o; o += R.LP * R.NP;   // BW[r*NP + i] = b[i][r] + (r>0 ? w_c[i][r-1]-w_nn : 0)
    R.B0 = o; o += R.LP * R.NP;   // B0[r*NP + i] = b[i][r]
    R.BWsuf = o; o += (R.LP + 1) * R.NP;  // BWsuf[k*NP + i] = sum_{r>=k} BW[r][i]
    R.total = o;
    return R;
}

void ring_pack(const HostModel &M, const RingLayout &R, double *dst);

// Warp index broadcast from lane 0: the value is the same as threadIdx.x >> 5, but the compiler now KNOWS it is
// warp-uniform, so the chunk index, the role and every loop bound derived from it can live in uniform
// registers and loop control runs on the uniform datapath (e.g. UR-indexed LDCU in a rolled FIR tap loop).
__device__ __forceinline__ int warp_index_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Geometry of the register-blocked FIR: each lane produces R consecutive
// outputs per neuron; a "super-window" is 32*R samples.  Tiles are stored
// transposed (sample e -> column e/R, row e%R) with row strides chosen so that
// both the transposing writes and the lane-per-sample reads are conflict-free.
template <int R>
struct FirGeom {
    static constexpr int SW = 32 * R;                       // samples per super-window
    static constexpr int LOGR = (R == 8) ? 3 : 2;
    static constexpr int FS = 32 + 16 / R;                  // F tile row stride (34 | 36)
    static constexpr int YS = (R == 8) ? 50 : 68;           // y tile row stride
    static constexpr int YTILE = R * YS;                    // doubles
    static constexpr int FTILE = R * FS;                    // doubles per neuron
};

// Stage y[b, b + SW + LP) into the transposed tile with cp.async (zero beyond T),
// then run the register-blocked FIR: lane computes F_i(b + R*lane + j), j < R,
// F_i(t0) = Bc[i] + sum_r A[r][i] * y[t0 + r], and leaves the results in `fbuf`
// (element t0-b = R*l + j of neuron i at fbuf[i*FTILE + j*FS + l]).
// `fbuf` may alias `ytile`.  A is the shared-memory copy [LP][NP] (zero padded).

// FIR coefficients passed BY VALUE as a kernel parameter: they live in the
// constant bank, and with the tap loop fully unrolled every DFMA takes its
// coefficient as a c[0x0][imm] operand.  A DFMA with three distinct 64-bit
// register operands is limited by register-file read bandwidth to one per 3
// cycles per SM sub-partition; with a constant operand it issues at the FP64
// pipe rate (one per 2 cycles) and the coefficient LDS traffic disappears.
template <int N, int LPC, typename S = double>
struct FirCoef {
    using scalar = S;           // double, or float in FP32 mode (hmm_set_precision / hmm_viterbi_f32)
    S a[(LPC > 0 ? LPC : 1) * N];  // a[r*N + i]
};

// Issue the asynchronous staging of y[b, b + need) into a transposed tile (zero beyond T) and
// commit it as one cp.async group; the caller waits for the group before the FIR reads it.
template <int R>
__device__ __forceinline__ void fir_stage(const double *__restrict__ y, int64_t T, int64_t b, int need, double *ytile,
                                          int lane) {
    using G = FirGeom<R>;
    if (b + need <= T) {
        // interior: element k = lane + 32 m goes to row lane % R, column lane / R + (32 / R) m -- one pointer pair
        // per lane and immediate offsets, no per-element bounds test
        const double *src = y + b + lane;
        double *dst = ytile + (lane & (R - 1)) * G::YS + (lane >> G::LOGR);
        constexpr int MAXM = (G::SW + RING_MAX_L + 31) / 32;
#pragma unroll
        for (int m = 0; m < MAXM; m++)
            if (lane + 32 * m < need) cp_async8(dst + (32 / R) * m, src + 32 * m);
    } else {
        for (int k = lane; k < need; k += 32) {
            int64_t g = b + k;
            double *dst = ytile + (k & (R - 1)) * G::YS + (k >> G::LOGR);
            if (g < T)
                cp_async8(dst, y + g);
            else
                *dst = 0.0;
        }
    }
    cp_async_commit();
}
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// Path-score piece 1 (ring_viterbi.cu ll_assemble): the FIR's initial register window holds exactly the super-window's
// own samples, R per lane (element e = R lane + j): accumulate sum (w0 - e) (y_e - m0)^2 over the first nvalid of them.
// The dozen extra FP64 operations per lane sit in the issue gaps of the DFMA stream that follows.
template <int R>
__device__ __forceinline__ void ll_noise_from_window(const double (&w)[R], double *nacc, double m0, double w0, int lane,
                                                     int nvalid) {
    double acc = *nacc;
    const int e0 = R * lane;
    double wg = w0 - (double)e0;
#pragma unroll
    for (int j = 0; j < R; j++) {
        const double dd = w[j] - m0;
        if (e0 + j < nvalid) acc = fma(wg, dd * dd, acc);
        wg -= 1.0;
    }
    *nacc = acc;
}

// I0, I1: the neurons [I0, I1) this call computes (the FIR of one super-window can be split between the producer
// and the consumer warp of a slot; both read the same y tile and write disjoint planes of the F tile).
// FP32 mode (typename S = float): the samples and the coefficients are rounded to float when they are loaded and the
// multiply-accumulates run in FP32 (twice the FP64 rate); everything downstream of F -- the recursion, the
// normalisation, the boundary vectors, ll -- stays FP64.  The rounding is a pure function of (y, model, t0), the
// same in every chunk and in the repair kernel, so speculation / verification work unchanged.
template <int N, int R, int I0 = 0, int I1 = N, typename S = double>
__device__ __forceinline__ void fir_compute(const double *A, const double *Bc, int LP, const double *ytile,
                                            double *fbuf, int lane, double *nacc = nullptr, double m0 = 0.0,
                                            double w0 = 0.0, int nvalid = 0) {
    using G = FirGeom<R>;
    constexpr int NP = (N + 1) & ~1;
    __builtin_assume(__isShared(ytile));
    __builtin_assume(__isShared(fbuf));
    __builtin_assume(__isShared(A));
    S acc[N][R];
#pragma unroll
    for (int i = I0; i < I1; i++)
#pragma unroll
        for (int j = 0; j < R; j++) acc[i][j] = (S)Bc[i];
    S w[R];
    {
        double wd[R];
#pragma unroll
        for (int j = 0; j < R; j++) wd[j] = ytile[j * G::YS + lane];  // elements R*lane + j
        if (nacc) ll_noise_from_window<R>(wd, nacc, m0, w0, lane, nvalid);
#pragma unroll
        for (int j = 0; j < R; j++) w[j] = (S)wd[j];
    }
    auto load_coef = [&](int r, S *dst) {
        const double2 *src = reinterpret_cast<const double2 *>(A + r * NP);
#pragma unroll
        for (int i2 = 0; i2 < NP / 2; i2++) {
            double2 v = src[i2];
            if (2 * i2 < N) dst[2 * i2] = (S)v.x;
            if (2 * i2 + 1 < N) dst[2 * i2 + 1] = (S)v.y;
        }
    };
    S a0[N], a1[N];
    load_coef(0, a0);
    for (int r0 = 0; r0 < LP; r0 += R) {
        const int col = lane + 1 + (r0 >> G::LOGR);
#pragma unroll
        for (int u = 0; u < R; u += 2) {
            // coefficients are fetched one tap ahead into the other register set
            load_coef(r0 + u + 1, a1);
            S ynew = (S)ytile[u * G::YS + col];  // element R*lane + (r0+u) + R
#pragma unroll
            for (int j = 0; j < R; j++) {
                const S yv = w[(u + j) % R];
#pragma unroll
                for (int i = I0; i < I1; i++) acc[i][j] = fma(a0[i], yv, acc[i][j]);
            }
            w[u] = ynew;
            load_coef(r0 + u + 2 < LP ? r0 + u + 2 : LP - 1, a0);
            ynew = (S)ytile[(u + 1) * G::YS + col];
#pragma unroll
            for (int j = 0; j < R; j++) {
                const S yv = w[(u + 1 + j) % R];
#pragma unroll
                for (int i = I0; i < I1; i++) acc[i][j] = fma(a1[i], yv, acc[i][j]);
            }
            w[u + 1] = ynew;
        }
    }
    __syncwarp();  // every lane is done with the y tile before F overwrites it
#pragma unroll
    for (int i = I0; i < I1; i++)
#pragma unroll
        for (int j = 0; j < R; j++) fbuf[i * G::FTILE + j * G::FS + lane] = (double)acc[i][j];
    __syncwarp();
}

template <int N, int R, typename S = double>
__device__ __forceinline__ void fir_superwindow(const double *__restrict__ y, int64_t T, int64_t b,
                                                const double *A, const double *Bc, int LP, double *ytile,
                                                double *fbuf, int lane) {
    fir_stage<R>(y, T, b, FirGeom<R>::SW + LP, ytile, lane);
    cp_async_wait_all();
    __syncwarp();
    fir_compute<N, R, 0, N, S>(A, Bc, LP, ytile, fbuf, lane);
}


#define HMM_PRAGMA_(x) _Pragma(#x)
#define HMM_UNROLL_N(n) HMM_PRAGMA_(unroll n)

template <int N, int R, int LPC, int I0 = 0, int I1 = N, typename S = double>
__device__ __forceinline__ void fir_compute_c(const FirCoef<N, LPC, S> &coef, const double *Bc, const double *ytile,
                                              double *fbuf, int lane, double *nacc = nullptr, double m0 = 0.0,
                                              double w0 = 0.0, int nvalid = 0) {
    using G = FirGeom<R>;
    __builtin_assume(__isShared(ytile));
    __builtin_assume(__isShared(fbuf));
    S acc[N][R];
#pragma unroll
    for (int i = I0; i < I1; i++)
#pragma unroll
        for (int j = 0; j < R; j++) acc[i][j] = (S)Bc[i];
    S w[R];
    {
        double wd[R];
#pragma unroll
        for (int j = 0; j < R; j++) wd[j] = ytile[j * G::YS + lane];
        if (nacc) ll_noise_from_window<R>(wd, nacc, m0, w0, lane, nvalid);
#pragma unroll
        for (int j = 0; j < R; j++) w[j] = (S)wd[j];
    }
    // Fully unrolled over exactly LPC = L taps, so that every coefficient is a compile-time constant-bank offset
    // (the compiler keeps them in uniform registers: LDCU.128 + DFMA R, R, UR, R -- no register-file or
    // shared-memory traffic for them).  Measured at C2 (forward kernel, 18 M samples): full unroll 0.36 ms
    // (25 KB of code: ncu shows a stall_no_instruction sample at every 128-byte line).  -DHMM_FIR_ROLLED rolls
    // the loop in groups of 8 taps (3.7 KB body, no instruction-fetch stalls): with the warp index known to be
    // uniform (warp_index_uniform) the compiler indexes the coefficients through uniform registers
    // (LDCU.64 c[0x0][UR+imm]) and the kernel takes 0.38 ms; before that it used register-indexed LDC and took
    // 0.45 ms.  Shared-memory coefficients: 0.52 ms.
    // -DHMM_FIR_GROUPS=g: g groups of R taps per loop iteration (a partly rolled loop: g*3.7 KB body).
#if defined(HMM_FIR_GROUPS)
    HMM_UNROLL_N(HMM_FIR_GROUPS)
#elif defined(HMM_FIR_ROLLED)
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int r0 = 0; r0 < LPC; r0 += R) {
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int r = r0 + u;
            if (r < LPC) {
                const S ynew = (S)ytile[u * G::YS + lane + 1 + (r0 >> G::LOGR)];
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const S yv = w[(u + j) % R];
#pragma unroll
                    for (int i = I0; i < I1; i++) acc[i][j] = fma(coef.a[r * N + i], yv, acc[i][j]);
                }
                w[u] = ynew;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = I0; i < I1; i++)
#pragma unroll
        for (int j = 0; j < R; j++) fbuf[i * G::FTILE + j * G::FS + lane] = (double)acc[i][j];
    __syncwarp();
}

template <int N, int R, int LPC, typename S = double>
__device__ __forceinline__ void fir_superwindow_c(const double *__restrict__ y, int64_t T, int64_t b,
                                                  const FirCoef<N, LPC, S> &coef, const double *Bc, double *ytile,
                                                  double *fbuf, int lane) {
    fir_stage<R>(y, T, b, FirGeom<R>::SW + LPC, ytile, lane);
    cp_async_wait_all();
    __syncwarp();
    fir_compute_c<N, R, LPC, 0, N, S>(coef, Bc, ytile, fbuf, lane);
}

// ---- mbarrier helpers (producer / consumer hand-off between warps of one CTA) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        " .reg .pred P1;\n"
        "MBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        " @P1 bra MBAR_DONE;\n"
        " bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}

// F value of step (local index tl in the super-window) for the lane-per-step mapping
template <int R>
__device__ __forceinline__ int fbuf_index(int tl) {
    using G = FirGeom<R>;
    return (tl & (R - 1)) * G::FS + (tl >> G::LOGR);
}

// log(exp(a) + exp(b)); exact shortcut when the smaller term is below half an ulp
// of the larger (log1p(exp(d)) < 2^-54 for d < -37.5), which is also what the
// reference's logsumexpl (src/utils.jl:24-32) rounds to.
__device__ __forceinline__ double lse2(double a, double b) {
    double hi = a > b ? a : b, lo = a > b ? b : a;
    if (lo == -INFINITY) return hi;
    double d = lo - hi;
    if (d > -37.5) hi += log1p(exp(d));
    return hi;
}

}  // namespace hmm
