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
    int A, BW, Bc, eG, eH, eT, scal, total;  // offsets in doubles
};
// scal[]: 0 w_nn, 1 c_emit, 2 two_s2, 3 m0, 4 sigma
__host__ __device__ inline RingLayout ring_layout(int N, int L) {
    RingLayout R;
    R.N = N;
    R.L = L;
    R.LP = (L + 7) & ~7;
    R.NP = (N + 1) & ~1;
    int o = 0;
    R.A = o;  o += R.LP * R.NP;   // A[r*NP + i]
    R.BW = o; o += R.LP * R.NP;   // BW[r*NP + i] = b[i][r] + (r>0 ? w_c[i][r-1]-w_nn : 0)
    R.Bc = o; o += R.NP;
    R.eG = o; o += R.NP;
    R.eH = o; o += R.NP;
    R.eT = o; o += N * R.NP;      // eT[j*NP + i]
    R.scal = o; o += 8;
    R.total = o;
    return R;
}

void ring_pack(const HostModel &M, const RingLayout &R, double *dst);

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

}  // namespace hmm
