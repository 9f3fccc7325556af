// ring_em.cu -- one fused Baum-Welch E/M step for non-overlap ring models
// (forward + backward + update of src/baumwelch.jl:25-51, 73-98, 205-309,
// 362-370) without ever materialising alpha, beta or gamma (nstates x T).
//
// Because a chain is a delay line, being in chain state (i, s) at time t is
// the same event as "chain i was entered at t0 = t - s + 1", so
//     gamma_t(i, s) = pi_i(t - s + 1)
// and the whole E-step reduces to, per sample, the noise posterior, one
// entry posterior per neuron, and the FIR scores F_i(t0) (ring_common.cuh):
//   forward : lg_t  = LSE(lg_{t-1}, ltail_{t-1}(j) + lA_j)
//             lp_t(i)= LSE(lg_{t-1} + lH_i, ltail_{t-1}(j) + lC_ji)
//             ltail_t(i) = lp_{t-L+1}(i) + F_i(t-L+1)
//   backward: lh_t  = LSE(lh_{t+1}, lH_i + lr_{t+1}(i))
//             le_t(i)= LSE(lA_i + lh_{t+1}, lC_ij + lr_{t+1}(j))
//             lr_t(i)= F_i(t) + le_{t+L-1}(i)
// all in the log domain, normalised by the all-noise path, 32 time steps per
// warp pass with a warp-shuffle log-sum-exp scan for the noise state.
// M-step statistics: S0 = sum pi, S1[s] = sum pi(t0) y[t0+s] (a sparse
// correlation), xi counts for lA, sum y^2 -- fused reductions (pass 3).
// Time is cut into chunks with speculative warm-up, verified and repaired
// exactly like the Viterbi engine (ring_viterbi.cu).
#include <cmath>
#include <cstdlib>
#include <limits>

#include "ring_common.cuh"

namespace hmm {

struct EmParams {
    const double *y;
    int64_t T;
    const double *model;  // one channel
    RingLayout RL;
    int64_t Lc, W;
    int nchunks, ns;
    double *Fg, *LG, *LQ, *LH, *LE;  // per-step arrays: Fg/LQ/LE are [N][T]
    double *LQneg;                    // [N][L]: virtual chains already running at t=0 (entered at -r0)
    double *SBf, *EBf, *SBb, *EBb;    // boundary vectors [nchunks][bvec]
    int bvec;
    int *flag_f, *flag_b;
    double *kappa, *lambda;           // per-chunk log offsets (forward / backward)
    double *lS;                       // [1] log of the normalised total likelihood
    int *counters;                    // [0] fwd repaired [1] bwd repaired [2] any fwd flag [3] any bwd flag
    double *dk;                       // [2][nchunks] boundary differences in scan order (em_check)
    double *part;                     // [nblk][PSTRIDE] statistic partials
    double *tot;                      // [PSTRIDE] their column sums (em_reduce)
    int nblk, pstride;
    double *out;                      // finalize output
    int dbg;                          // debug switch (HMMCUDA_EM_DBG): 1 = log-domain live windows
    // time-sharded E-step (hmm_emshard_*): the statistics are accumulated over the local steps [st_lo, st_hi) only
    // (this shard's main span; the ghost chunks on either side only make the posteriors there exact)
    int64_t st_lo, st_hi;
};

enum { EM_INIT = 0, EM_SPEC = 1, EM_EXACT = 2 };
constexpr int S1_LAGS = 96;  // == RING_MAX_L


template <int N, int R>
struct EmWarpSmem {
    using G = FirGeom<R>;
    static constexpr int TILE = (G::YTILE > N * G::FTILE) ? G::YTILE : N * G::FTILE;  // em_fir: y tile / F tile
    static constexpr int DOUBLES = N * RING_Q;                                        // recursions: the ring
};

// Inclusive log-sum-exp scan over the lanes of a warp (lane 0 first).
__device__ __forceinline__ double lse_scan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = shfl_up_d(v, d);
        if (lane >= d) v = lse2(v, o);
    }
    return v;
}

// ---------------------------------------------------------------------------
// forward, one chunk per warp
// ---------------------------------------------------------------------------
template <int N, int R>
__device__ void em_fwd_chunk(const EmParams &p, int c, int kind, const double *mdl, double *ws) {
    using G = FirGeom<R>;
    constexpr int NP = (N + 1) & ~1;
    const int lane = threadIdx.x & 31;
    const RingLayout &RL = p.RL;
    const int L = RL.L;
    const double NEG = -INFINITY;
    double *ring = ws;
    const double *lA = mdl + RL.eG, *lH = mdl + RL.eH, *lC = mdl + RL.eT;
    // cF[j] = max(lA_j, max_i(lC_ji - lH_i)): how much a tail of neuron j can gain on the noise term
    double cF[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        double c = lA[j];
#pragma unroll
        for (int i = 0; i < N; i++)
            if (i != j) c = fmax(c, lC[j * NP + i] - lH[i]);
        cF[j] = c;
    }
    // linear-domain copies of the weights, and cM[j] = the largest weight a tail of neuron j is multiplied by
    const double *xA = mdl + RL.xG, *xH = mdl + RL.xH, *xC = mdl + RL.xT;
    double cM[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        double c = lA[j];
#pragma unroll
        for (int i = 0; i < N; i++)
            if (i != j) c = fmax(c, lC[j * NP + i]);
        cM[j] = c;
    }
    const double *cold = p.model;
    const double *y = p.y;
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    const bool last = (c == p.nchunks - 1);
    if (last || e > T) e = T;
    int64_t base0, tau_first;
    double lgprev;
    for (int k = lane; k < N * RING_Q; k += 32) ring[k] = NEG;
    __syncwarp();
    if (kind == EM_INIT) {
        base0 = 0;
        tau_first = 1;
        lgprev = 0.0;  // alpha_1(noise) / itself
        // chains already running at t = 0: state (i, r0+1) <-> entered at tau0 = -r0 with weight 1
        const double *BW = cold + RL.BW, *B0 = cold + RL.B0;
        for (int idx = lane; idx < N * L; idx += 32) {
            int i = idx / L, r0 = idx % L;
            double f = 0.0;
            for (int r = r0; r < L; r++) f += fma(cold[RL.A + r * NP + i], y[r - r0], BW[r * NP + i]);
            if (r0 >= 1) f -= BW[r0 * NP + i] - B0[r0 * NP + i];  // no chain transition INTO the first sample
            ring[i * RING_Q + ((-r0) & (RING_Q - 1))] = f;
            p.LQneg[i * L + r0] = f;
        }
    } else if (kind == EM_SPEC) {
        base0 = s - p.W;
        if (base0 < 0) base0 = 0;
        tau_first = base0;
        lgprev = 0.0;
    } else {
        base0 = s;
        tau_first = s;
        const double *eb = p.EBf + (size_t)(c - 1) * p.bvec;
        lgprev = eb[0];
        for (int k = lane; k < L; k += 32) {
            int64_t t0 = s - L + k;
            for (int j = 0; j < N; j++) ring[j * RING_Q + (int)(t0 & (RING_Q - 1))] = eb[1 + j * L + k];
        }
    }
    __syncwarp();
    const int Wd = L < 32 ? L : 32;
    const int nsub = (32 + Wd - 1) / Wd;
    const int mysub = lane / Wd;
    const int tf_rel = (int)(tau_first - base0), e_rel = (int)(e - base0), s_rel = (int)(s - base0);

    // F_i(tau) comes from the FIR pass (em_fir; end-of-recording terms already dropped).  The loads of the
    // next window are issued before the current one is processed.
    // per-neuron row pointers at base0 (steps are 32-bit offsets from it below)
    const double *fgp[N];
    double *lqp[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        fgp[i] = p.Fg + (size_t)i * T + base0;
        lqp[i] = p.LQ + (size_t)i * T + base0;
    }
    double *lgp = p.LG + base0;
    const int T_rel = (int)((T - base0) < (int64_t)1 << 30 ? (T - base0) : (int64_t)1 << 30);
    double Fnx[N];
#pragma unroll
    for (int i = 0; i < N; i++) Fnx[i] = lane < T_rel ? fgp[i][lane] : 0.0;
    // dead-window constants, valid while lg == lgq (lg moves only where a chain ends)
    double lgq = NAN, lhq[N], thq[N];
#pragma unroll
    for (int i = 0; i < N; i++) lhq[i] = thq[i] = 0.0;
    for (int64_t b = base0; b < e; b += G::SW) {
        if (kind == EM_SPEC && b == s) {
            double *sb = p.SBf + (size_t)c * p.bvec;
            if (lane == 0) sb[0] = lgprev;
            for (int k = lane; k < L; k += 32) {
                int64_t t0 = s - L + k;
                for (int j = 0; j < N; j++) sb[1 + j * L + k] = ring[j * RING_Q + (int)(t0 & (RING_Q - 1))];
            }
        }
        const int b_rel = (int)(b - base0);
        for (int wdw = 0; wdw < R; wdw++) {
            const int t0_rel = b_rel + 32 * wdw;
            if (t0_rel >= e_rel) break;
            const int t_rel = t0_rel + lane;
            const int64_t tau = base0 + t_rel;
            double Fv[N];
#pragma unroll
            for (int i = 0; i < N; i++) {
                Fv[i] = Fnx[i];
                Fnx[i] = (t0_rel + 32 < e_rel && t_rel + 32 < T_rel) ? fgp[i][t_rel + 32] : 0.0;
            }
            const bool in_range = t_rel >= tf_rel && t_rel < e_rel;
            const int slot_w = t_rel & (RING_Q - 1);
            const int slot_r = (t_rel - L) & (RING_Q - 1);
            if (kind == EM_INIT && tau == 0) {
                p.LG[0] = 0.0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    p.LQ[(size_t)i * T] = Fv[i];
                    ring[i * RING_Q] = Fv[i];
                }
            }
            __syncwarp();
            for (int sub = 0; sub < nsub; sub++) {
                const bool active = in_range && (mysub == sub);
                double lt[N];
#pragma unroll
                for (int j = 0; j < N; j++) lt[j] = active ? ring[j * RING_Q + slot_r] : NEG;
                // Dead-window fast path: if every arriving tail is more than e^-43 below the noise
                // term in every sum it enters, each lse2 below would take its "smaller term is under
                // half an ulp" shortcut (32 lanes x N terms of e^-43 still sum to < e^-37.5), so
                // lg stays and lp_i = lg + lH_i exactly -- no transcendental, no scan.
                {
                    if (lgq != lgprev) {
                        lgq = lgprev;
#pragma unroll
                        for (int i = 0; i < N; i++) {
                            lhq[i] = lgprev + lH[i];
                            thq[i] = (lgprev - 43.0) - cF[i];  // alive <=> lt_j + cF_j > lg - 43
                        }
                    }
                    bool alive = false;
#pragma unroll
                    for (int j = 0; j < N; j++) alive = alive || (lt[j] > thq[j]);
                    if (!__any_sync(0xffffffffu, active && alive)) {
                        if (active) {
#pragma unroll
                            for (int i = 0; i < N; i++) {
                                const double lq = lhq[i] + Fv[i];
                                ring[i * RING_Q + slot_w] = lq;
                                if (t0_rel >= s_rel) lqp[i][t_rel] = lq;
                            }
                            if (t0_rel >= s_rel) lgp[t_rel] = lgprev;
                        }
                        __syncwarp();
                        continue;
                    }
                }
                // Live window.  The sums are formed in the linear domain, relative to the largest term of
                // the window (ref), so that a step costs N exp + 1 log instead of ~N^2 + 5 log-sum-exps:
                //   g_t = g_{t-1} + sum_j u_j(t) xA_j   (a prefix sum over the lanes),   lg_t = ref + log g_t
                double mx = NEG;
#pragma unroll
                for (int j = 0; j < N; j++) mx = fmax(mx, lt[j] + cM[j]);
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
                const double ref = fmax(lgprev, mx);
                double lg, lgm1;
                if (ref - lgprev < 600.0 && !(p.dbg & 1)) {
                    double u[N];
#pragma unroll
                    for (int j = 0; j < N; j++) u[j] = exp(lt[j] - ref);  // exp(-inf) = 0 for idle lanes
                    const double g0 = exp(lgprev - ref);
                    double cs = 0.0;
#pragma unroll
                    for (int j = 0; j < N; j++) cs = fma(u[j], xA[j], cs);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double o = shfl_up_d(cs, d);
                        if (lane >= d) cs += o;
                    }
                    const double g = g0 + cs;
                    lg = ref + log(g);
                    double gm1 = shfl_up_d(g, 1);
                    lgm1 = shfl_up_d(lg, 1);
                    if (lane == 0) {
                        gm1 = g0;
                        lgm1 = lgprev;
                    }
                    if (active) {
#pragma unroll
                        for (int i = 0; i < N; i++) {
                            double cross = 0.0;
#pragma unroll
                            for (int j = 0; j < N; j++)
                                if (j != i) cross = fma(u[j], xC[j * NP + i], cross);
                            const double nz = gm1 * xH[i];
                            double lp = lgm1 + lH[i];
                            if (cross > nz * 0x1p-60) lp = ref + log(nz + cross);  // a tail feeds head i directly
                            const double lq = lp + Fv[i];
                            ring[i * RING_Q + slot_w] = lq;
                            if (t0_rel >= s_rel) lqp[i][t_rel] = lq;
                        }
                        if (t0_rel >= s_rel) lgp[t_rel] = lg;
                    }
                } else {
                    // a single window spans more than e^600: stay in the log domain
                    double lX = NEG;
#pragma unroll
                    for (int j = 0; j < N; j++) lX = lse2(lX, lt[j] + lA[j]);
                    const double lCs = lse_scan(lX, lane);
                    lg = lse2(lgprev, lCs);
                    lgm1 = shfl_up_d(lg, 1);
                    if (lane == 0) lgm1 = lgprev;
                    if (active) {
#pragma unroll
                        for (int i = 0; i < N; i++) {
                            double lp = lgm1 + lH[i];
#pragma unroll
                            for (int j = 0; j < N; j++)
                                if (j != i) lp = lse2(lp, lt[j] + lC[j * NP + i]);
                            const double lq = lp + Fv[i];
                            ring[i * RING_Q + slot_w] = lq;
                            if (t0_rel >= s_rel) lqp[i][t_rel] = lq;
                        }
                        if (t0_rel >= s_rel) lgp[t_rel] = lg;
                    }
                }
                lgprev = shfl_d(lg, 31);
                __syncwarp();
            }
        }
    }
    if (!last) {
        double *eb = p.EBf + (size_t)c * p.bvec;
        if (lane == 0) eb[0] = lgprev;
        for (int k = lane; k < L; k += 32) {
            int64_t t0 = e - L + k;
            for (int j = 0; j < N; j++) eb[1 + j * L + k] = ring[j * RING_Q + (int)(t0 & (RING_Q - 1))];
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// FIR pass: F_i(t0) = Bc_i + sum_r a[i][r] y[t0 + r] for every t0, one super-window per warp, stored to
// Fg[N][T].  Chains that would run past the end of the recording have the terms beyond T-1 dropped here.
// LPC > 0: coefficients as constant-bank operands (ring_common.cuh), else from shared memory.
// ---------------------------------------------------------------------------
template <int N, int R, int LPC>
__global__ void __launch_bounds__(128) em_fir(EmParams p, const __grid_constant__ FirCoef<N, LPC> coef) {
    using G = FirGeom<R>;
    constexpr int NP = (N + 1) & ~1;
    extern __shared__ __align__(16) double smem_d[];
    double *mdl = smem_d;
    for (int k = threadIdx.x; k < p.RL.hot; k += blockDim.x) mdl[k] = p.model[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = warp_index_uniform();
    const RingLayout &RL = p.RL;
    const int L = RL.L;
    const int64_t T = p.T;
    const int64_t b = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * G::SW;
    if (b >= T) return;
    double *tile = smem_d + ((RL.hot + 1) & ~1) + (size_t)warp * EmWarpSmem<N, R>::TILE;
    if constexpr (LPC > 0)
        fir_superwindow_c<N, R, LPC>(p.y, T, b, coef, mdl + RL.Bc, tile, tile, lane);
    else
        fir_superwindow<N, R>(p.y, T, b, mdl + RL.A, mdl + RL.Bc, RL.LP, tile, tile, lane);
    const double *BWsuf = p.model + RL.BWsuf;
#pragma unroll
    for (int wdw = 0; wdw < R; wdw++) {
        const int tl = 32 * wdw + lane;
        const int64_t tau = b + tl;
        if (tau < T) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                double f = tile[i * G::FTILE + fbuf_index<R>(tl)];
                if (tau > T - L) f -= BWsuf[(int)(T - tau) * NP + i];
                p.Fg[(size_t)i * T + tau] = f;
            }
        }
    }
}

template <int N, int R>
__global__ void __launch_bounds__(128, 4) em_forward(EmParams p) {
    extern __shared__ __align__(16) double smem_d[];
    double *mdl = smem_d;
    for (int k = threadIdx.x; k < p.RL.hot; k += blockDim.x) mdl[k] = p.model[k];
    __syncthreads();
    const int warp = warp_index_uniform();
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= p.nchunks) return;
    double *ws = smem_d + ((p.RL.hot + 1) & ~1) + (size_t)warp * EmWarpSmem<N, R>::DOUBLES;
    em_fwd_chunk<N, R>(p, c, c == 0 ? EM_INIT : EM_SPEC, mdl, ws);
}

// Two boundary vectors describe the same distribution iff they differ by a
// constant (every finite entry is compared, see below).
__device__ __forceinline__ bool em_boundary_matches(const double *sb, const double *eb, int n, int lane) {
    double ms = -INFINITY, me = -INFINITY;
    for (int k = lane; k < n; k += 32) {
        ms = fmax(ms, sb[k]);
        me = fmax(me, eb[k]);
    }
    for (int d = 16; d >= 1; d >>= 1) {
        ms = fmax(ms, __shfl_xor_sync(0xffffffffu, ms, d));
        me = fmax(me, __shfl_xor_sync(0xffffffffu, me, d));
    }
    bool bad = false;
    for (int k = lane; k < n; k += 32) {
        double a = sb[k] - ms, b = eb[k] - me;
        // EVERY finite entry counts, however far below the maximum: the future multiplies a pending chain's entry by
        // its own emission product, which at high SNR is e^(+1000s) -- an entry 745 below the largest one NOW can
        // carry the whole mass a few steps later.  (An earlier version skipped entries more than 745 below the
        // maximum in both vectors and accepted a boundary whose noise score had not converged: found by
        // tools/fuzz_parity.py, N=6 K=81 sigma=0.31 with template amplitudes of 12 sigma.)
        if (a == -INFINITY && b == -INFINITY) continue;
        if (!(fabs(a - b) <= 1e-11 + 1e-13 * fabs(b))) bad = true;
    }
    return !__any_sync(0xffffffffu, bad);
}

// dirs: bit 0 = forward boundaries, bit 1 = backward boundaries (blockIdx.y selects the direction)
__global__ void em_check(EmParams p, int dirs) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int backward = blockIdx.y;
    if (!((dirs >> backward) & 1)) return;
    const int n = p.nchunks;
    if (gw >= n) return;
    // also leaves the difference of the two normalisations in scan order for em_fixup (valid when nothing is
    // repaired) and raises counters[2 + direction] when any boundary failed
    if (!backward) {
        if (gw == 0) {
            if (lane == 0) p.dk[0] = 0.0;
            return;
        }
        const double *sb = p.SBf + (size_t)gw * p.bvec, *eb = p.EBf + (size_t)(gw - 1) * p.bvec;
        bool ok = em_boundary_matches(sb, eb, p.bvec, lane);
        if (lane == 0) {
            p.flag_f[gw] = ok ? 0 : 1;
            p.dk[gw] = eb[0] - sb[0];
            if (!ok) atomicOr(&p.counters[2], 1);
        }
    } else {
        if (gw == n - 1) {
            if (lane == 0) p.dk[n] = 0.0;  // r = n-1-c = 0
            return;
        }
        const double *sb = p.SBb + (size_t)gw * p.bvec, *eb = p.EBb + (size_t)(gw + 1) * p.bvec;
        bool ok = em_boundary_matches(sb, eb, p.bvec, lane);
        if (lane == 0) {
            p.flag_b[gw] = ok ? 0 : 1;
            p.dk[n + (n - 1 - gw)] = eb[0] - sb[0];
            if (!ok) atomicOr(&p.counters[3], 1);
        }
    }
}

// ---------------------------------------------------------------------------
// backward, one chunk per warp
// ---------------------------------------------------------------------------
template <int N>
__device__ void em_bwd_chunk(const EmParams &p, int c, int kind, const double *mdl, double *ring) {
    constexpr int NP = (N + 1) & ~1;
    const int lane = threadIdx.x & 31;
    const RingLayout &RL = p.RL;
    const int L = RL.L;
    const double *lA = mdl + RL.eG, *lH = mdl + RL.eH, *lC = mdl + RL.eT;
    // cB[j] = max(lH_j, max_i(lC_ij - lA_i)): how much entering chain j can gain on the noise term
    double cB[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        double c2 = lH[j];
#pragma unroll
        for (int i = 0; i < N; i++)
            if (i != j) c2 = fmax(c2, lC[i * NP + j] - lA[i]);
        cB[j] = c2;
    }
    const double *xA = mdl + RL.xG, *xH = mdl + RL.xH, *xC = mdl + RL.xT;
    double cM[N];  // the largest weight an entry into chain j is multiplied by
#pragma unroll
    for (int j = 0; j < N; j++) {
        double c2 = lH[j];
#pragma unroll
        for (int i = 0; i < N; i++)
            if (i != j) c2 = fmax(c2, lC[i * NP + j]);
        cM[j] = c2;
    }
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    const bool last = (c == p.nchunks - 1);
    if (last || e > T) e = T;
    // state is known (or assumed) at time `hi`; steps hi-1 .. s are computed
    int64_t hi;
    double lhprev = 0.0;
    for (int k = lane; k < N * RING_Q; k += 32) ring[k] = 0.0;  // beyond-the-end chains contribute e = 1
    __syncwarp();
    bool true_end = false;
    if (kind == EM_EXACT) {
        hi = e;  // boundary vector of chunk c+1 lives at time e
        const double *eb = p.EBb + (size_t)(c + 1) * p.bvec;
        lhprev = eb[0];
        for (int k = lane; k < L; k += 32)
            for (int j = 0; j < N; j++) ring[j * RING_Q + (int)((e + k) & (RING_Q - 1))] = eb[1 + j * L + k];
    } else {
        hi = e + p.W;
        if (last || hi >= T - 1) {
            hi = T - 1;
            true_end = true;
        }
        if (true_end && e == T && lane == 0) {  // beta_T = 0 (src/baumwelch.jl:80)
            p.LH[T - 1] = 0.0;
            for (int i = 0; i < N; i++) p.LE[(size_t)i * T + T - 1] = 0.0;
        }
    }
    __syncwarp();
    const int Wd = L < 32 ? L : 32;
    const int nsub = (32 + Wd - 1) / Wd;
    const int mysub = lane / Wd;
    // windows of 32 steps, descending; lane 0 is the latest step of the window
    // per-neuron row pointers (F shifted by one step: the recursion at t uses F(t+1))
    const double *fgp[N];
    double *lep[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        fgp[i] = p.Fg + (size_t)i * T + 1;
        lep[i] = p.LE + (size_t)i * T;
    }
    double Fnx[N];  // F of the next (earlier) window, loaded one window ahead
    {
        const int64_t t = ((hi - 1) | 31) - lane;
#pragma unroll
        for (int i = 0; i < N; i++) Fnx[i] = (t <= hi - 1 && t >= s) ? fgp[i][t] : 0.0;
    }
    // dead-window constants, valid while lh == lhq
    double lhq = NAN, leq[N], thq[N];
#pragma unroll
    for (int i = 0; i < N; i++) leq[i] = thq[i] = 0.0;
    for (int64_t wtop = ((hi - 1) | 31); wtop >= s; wtop -= 32) {
        const int64_t t = wtop - lane;
        const bool in_range = t <= hi - 1 && t >= s;
        double Fn[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            Fn[i] = Fnx[i];
            const int64_t tn = t - 32;
            Fnx[i] = (tn <= hi - 1 && tn >= s) ? fgp[i][tn] : 0.0;
        }
        // F of the warm-up region belongs to the next chunk and is complete: the forward pass has finished
        for (int sub = 0; sub < nsub; sub++) {
            const bool active = in_range && (mysub == sub);
            double lr[N];
#pragma unroll
            for (int j = 0; j < N; j++)
                lr[j] = active ? Fn[j] + ring[j * RING_Q + (int)((t + L) & (RING_Q - 1))] : -INFINITY;
            // Dead-window fast path (see em_fwd_chunk): every entry term is more than e^-43 below
            // the noise continuation, so lh stays and le_i = lA_i + lh exactly.
            {
                if (lhq != lhprev) {
                    lhq = lhprev;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        leq[i] = lA[i] + lhprev;
                        thq[i] = (lhprev - 43.0) - cB[i];  // alive <=> lr_j + cB_j > lh - 43
                    }
                }
                bool alive = false;
#pragma unroll
                for (int j = 0; j < N; j++) alive = alive || (lr[j] > thq[j]);
                if (!__any_sync(0xffffffffu, active && alive)) {
                    if (active) {
#pragma unroll
                        for (int i = 0; i < N; i++) {
                            ring[i * RING_Q + (int)(t & (RING_Q - 1))] = leq[i];
                            if (t < e) lep[i][t] = leq[i];
                        }
                        if (t < e) p.LH[t] = lhprev;
                    }
                    __syncwarp();
                    continue;
                }
            }
            // Live window: linear domain relative to the window's largest term (see em_fwd_chunk).
            double mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < N; j++) mx = fmax(mx, lr[j] + cM[j]);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            const double ref = fmax(lhprev, mx);
            double lh, lhp1;
            if (ref - lhprev < 600.0 && !(p.dbg & 1)) {
                double v[N];
#pragma unroll
                for (int j = 0; j < N; j++) v[j] = exp(lr[j] - ref);
                const double h0 = exp(lhprev - ref);
                double cs = 0.0;
#pragma unroll
                for (int j = 0; j < N; j++) cs = fma(v[j], xH[j], cs);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double o = shfl_up_d(cs, d);
                    if (lane >= d) cs += o;
                }
                const double h = h0 + cs;
                lh = ref + log(h);
                double hp1 = shfl_up_d(h, 1);
                lhp1 = shfl_up_d(lh, 1);  // lh_{t+1}
                if (lane == 0) {
                    hp1 = h0;
                    lhp1 = lhprev;
                }
                if (active) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        double cross = 0.0;
#pragma unroll
                        for (int j = 0; j < N; j++)
                            if (j != i) cross = fma(v[j], xC[i * NP + j], cross);
                        const double nz = hp1 * xA[i];
                        double le = lA[i] + lhp1;
                        if (cross > nz * 0x1p-60) le = ref + log(nz + cross);
                        ring[i * RING_Q + (int)(t & (RING_Q - 1))] = le;
                        if (t < e) lep[i][t] = le;
                    }
                    if (t < e) p.LH[t] = lh;
                }
            } else {
                double lY = -INFINITY;
#pragma unroll
                for (int j = 0; j < N; j++) lY = lse2(lY, lH[j] + lr[j]);
                const double sc = lse_scan(lY, lane);
                lh = lse2(lhprev, sc);
                lhp1 = shfl_up_d(lh, 1);  // lh_{t+1}
                if (lane == 0) lhp1 = lhprev;
                if (active) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        double le = lA[i] + lhp1;
#pragma unroll
                        for (int j = 0; j < N; j++)
                            if (j != i) le = lse2(le, lC[i * NP + j] + lr[j]);
                        ring[i * RING_Q + (int)(t & (RING_Q - 1))] = le;
                        if (t < e) lep[i][t] = le;
                    }
                    if (t < e) p.LH[t] = lh;
                }
            }
            lhprev = shfl_d(lh, 31);
            __syncwarp();
        }
        // speculative boundary vector at time e (after the window whose lowest step is e)
        if (kind == EM_SPEC && !last && wtop - 31 == e) {
            double *sb = p.SBb + (size_t)c * p.bvec;
            if (lane == 0) sb[0] = lhprev;
            for (int k = lane; k < L; k += 32)
                for (int j = 0; j < N; j++) sb[1 + j * L + k] = ring[j * RING_Q + (int)((e + k) & (RING_Q - 1))];
            __syncwarp();
        }
    }
    if (c > 0) {  // true boundary vector at time s for chunk c-1
        double *eb = p.EBb + (size_t)c * p.bvec;
        if (lane == 0) eb[0] = lhprev;
        for (int k = lane; k < L; k += 32)
            for (int j = 0; j < N; j++) eb[1 + j * L + k] = ring[j * RING_Q + (int)((s + k) & (RING_Q - 1))];
    }
    __syncwarp();
}

template <int N>
__global__ void __launch_bounds__(128, 4) em_backward(EmParams p) {
    extern __shared__ __align__(16) double smem_d[];
    double *mdl = smem_d;
    for (int k = threadIdx.x; k < p.RL.hot; k += blockDim.x) mdl[k] = p.model[k];
    __syncthreads();
    const int warp = warp_index_uniform();
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= p.nchunks) return;
    double *ring = smem_d + ((p.RL.hot + 1) & ~1) + (size_t)warp * N * RING_Q;
    em_bwd_chunk<N>(p, c, EM_SPEC, mdl, ring);
}

// ---------------------------------------------------------------------------
// sequential repair + per-chunk log offsets (one warp)
// ---------------------------------------------------------------------------
// Block-wide inclusive scan of one double per thread over the first 256 threads, fixed order.  (Threads
// beyond the first 256 may take part in the barriers; their result is unused.)
__device__ __forceinline__ double block_scan_256(double v, double *wsum) {
    const int lane = threadIdx.x & 31, warp = warp_index_uniform();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = shfl_up_d(v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31 && warp < 8) wsum[warp] = v;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp && w < 8; w++) base += wsum[w];
    __syncthreads();
    return v + base;
}

constexpr int EM_SCAN_SMEM = 4096;  // chunk differences scanned in shared memory up to this many chunks

// One launch per E/M step, one CTA per direction (block 0 forward, block 1 backward):
//  1. sequential repair by warp 0: re-run flagged chunks from the true boundary vector; a re-run changes the
//     chunk's end vector, so its successor is re-checked against it;
//  2. per-chunk log offsets  kappa_c = sum_{k<=c} (EBf[k-1].lg - SBf[k].lg),
//                            lambda_c = sum_{k>=c} (EBb[k+1].lh - SBb[k].lh),
//     and lS = kappa_last + LSE(alpha-hat at T-1).
template <int N, int R>
__global__ void __launch_bounds__(1024) em_fixup(EmParams p, int dirs) {
    extern __shared__ __align__(16) double smem_d[];
    __shared__ double wsum[8];
    const int backward = blockIdx.x;
    if (!((dirs >> backward) & 1)) return;
    const int n = p.nchunks;
    double *mdl = smem_d;
    double *ws = mdl + ((p.RL.hot + 1) & ~1);
    double *dsm = ws + N * RING_Q;  // [EM_SCAN_SMEM]
    const int any = p.counters[2 + backward];  // raised by em_check when some boundary failed
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int repaired = 0;
        if (any) {
            for (int k = lane; k < p.RL.hot; k += 32) mdl[k] = p.model[k];
            __syncwarp();
            bool prev = false;
            if (!backward) {
                for (int c = 1; c < n; c++) {
                    bool need = p.flag_f[c] != 0;
                    if (!need && prev)
                        need = !em_boundary_matches(p.SBf + (size_t)c * p.bvec, p.EBf + (size_t)(c - 1) * p.bvec, p.bvec, lane);
                    if (need) {
                        em_fwd_chunk<N, R>(p, c, EM_EXACT, mdl, ws);
                        // an exactly restarted chunk continues chunk c-1's normalisation
                        if (lane == 0) p.SBf[(size_t)c * p.bvec] = p.EBf[(size_t)(c - 1) * p.bvec];
                        __threadfence();
                        __syncwarp();
                        repaired++;
                    }
                    prev = need;
                }
            } else {
                for (int c = n - 2; c >= 0; c--) {
                    bool need = p.flag_b[c] != 0;
                    if (!need && prev)
                        need = !em_boundary_matches(p.SBb + (size_t)c * p.bvec, p.EBb + (size_t)(c + 1) * p.bvec, p.bvec, lane);
                    if (need) {
                        em_bwd_chunk<N>(p, c, EM_EXACT, mdl, ws);
                        if (lane == 0) p.SBb[(size_t)c * p.bvec] = p.EBb[(size_t)(c + 1) * p.bvec];
                        __threadfence();
                        __syncwarp();
                        repaired++;
                    }
                    prev = need;
                }
            }
        }
        if (lane == 0) p.counters[backward] = repaired;
    }
    __syncthreads();
    // ---- offsets: gather the per-chunk differences with independent loads (the boundary vectors are bvec
    // doubles apart), then scan them in a fixed order on the first 256 threads ----
    const int per = (n + 255) / 256;  // contiguous chunks per thread
    const int c0 = threadIdx.x < 256 ? threadIdx.x * per : n;
    double *res = backward ? p.lambda : p.kappa;
    double *d = n <= EM_SCAN_SMEM ? dsm : res;  // staging of the differences (in place when too many for smem)
    if (!any) {  // nothing was repaired: em_check's differences stand
        for (int r = threadIdx.x; r < n; r += blockDim.x) d[r] = p.dk[(size_t)backward * n + r];
    } else if (!backward) {
        for (int c = threadIdx.x; c < n; c += blockDim.x)
            d[c] = c >= 1 ? p.EBf[(size_t)(c - 1) * p.bvec] - p.SBf[(size_t)c * p.bvec] : 0.0;
    } else {  // reversed order: r = n-1-c
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            const int c = n - 1 - r;
            d[r] = c < n - 1 ? p.EBb[(size_t)(c + 1) * p.bvec] - p.SBb[(size_t)c * p.bvec] : 0.0;
        }
    }
    // meanwhile warp 8: log-sum-exp of alpha-hat at T-1 (all loads first, then one max / exp / sum pass)
    __shared__ double ls_tail;
    if (!backward && (threadIdx.x >> 5) == 8) {
        const int lane = threadIdx.x & 31, L = p.RL.L;
        const int64_t T = p.T;
        constexpr int PER = (N * RING_MAX_L + 31) / 32;
        double v[PER];
        double m = lane == 0 ? p.LG[T - 1] : -INFINITY;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int idx = lane + 32 * q;
            v[q] = -INFINITY;
            if (idx < N * L) {
                const int i = idx / L, k = idx % L;  // t0 = T - L + k
                v[q] = p.LQ[(size_t)i * T + (T - L + k)];
            }
        }
        const double v0 = m;
#pragma unroll
        for (int q = 0; q < PER; q++) m = fmax(m, v[q]);
        for (int d2 = 16; d2 >= 1; d2 >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d2));
        double sum = v0 == -INFINITY ? 0.0 : exp(v0 - m);
#pragma unroll
        for (int q = 0; q < PER; q++)
            if (v[q] != -INFINITY) sum += exp(v[q] - m);
        for (int d2 = 16; d2 >= 1; d2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d2);
        if (lane == 0) ls_tail = m + log(sum);
    }
    __syncthreads();
    double loc = 0.0;
    for (int r = c0; r < c0 + per && r < n; r++) loc += d[r];
    const double incl = block_scan_256(loc, wsum);
    double run = incl - loc;
    if (!backward) {
        for (int r = c0; r < c0 + per && r < n; r++) {
            run += d[r];
            res[r] = run;
        }
    } else if (d == res) {
        // in place and reversed: lambda[c] lives at index n-1-r, which another thread may still read as d[..]
        for (int r = c0; r < c0 + per && r < n; r++) {
            run += d[r];
            d[r] = run;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < n / 2; r += blockDim.x) {
            const double a2 = d[r], b2 = d[n - 1 - r];
            d[r] = b2;
            d[n - 1 - r] = a2;
        }
    } else {
        for (int r = c0; r < c0 + per && r < n; r++) {
            run += d[r];
            res[n - 1 - r] = run;
        }
    }
    if (!backward) {
        __syncthreads();
        if (threadIdx.x == 0) p.lS[0] = ls_tail + p.kappa[n - 1];
    }
}

// ---------------------------------------------------------------------------
// pass 3: posteriors and M-step sufficient statistics (fused reductions)
// partial layout: [0] sum gamma0 (t <= T-2)  [1] sum y  [2] sum y^2  [3] sum gamma0 (all t)
//                 [4 .. 4+N) xi_i   [4+N .. 4+2N) S0tot_i   then S1[i][lag], lag < S1_LAGS
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) em_stats(EmParams p) {
    extern __shared__ __align__(16) double smem_d[];
    constexpr int WPB = 4;
    const int lane = threadIdx.x & 31, warp = warp_index_uniform();
    const RingLayout &RL = p.RL;
    const int L = RL.L;
    const int64_t T = p.T;
    const double lS = p.lS[0];
    const double *lH = p.model + RL.eH;
    double *ysm = smem_d + warp * (160 + N * 32);  // y[t0 .. t0+32+L) then pi[i][32]
    double *pism = ysm + 160;
    double a_g0 = 0, a_y = 0, a_y2 = 0, a_g0all = 0;
    double a_xi[N], a_s0[N], a_s1[N][3];
#pragma unroll
    for (int i = 0; i < N; i++) {
        a_xi[i] = 0;
        a_s0[i] = 0;
        a_s1[i][0] = a_s1[i][1] = a_s1[i][2] = 0;
    }
    const int64_t nwin = (T + 31) / 32;
    const int64_t gw = (int64_t)blockIdx.x * WPB + warp, nw = (int64_t)gridDim.x * WPB;
    // Everything one window reads from global memory.  The loads of window w + nw are issued before window w
    // is processed (the pass is otherwise bound by one exposed memory round trip per window).
    struct WinLoads {
        double kap, lam, lame, lamx, vLG, vLH, vLQ[N], vLEe[N], vLEx[N], vFn[N], yk[4];
    };
    auto issue = [&](int64_t w, WinLoads &o) {
        const int64_t t = w * 32 + lane;
        // chunk indices without per-lane 64-bit divisions: Lc is a multiple of 256, so the whole
        // window lies in one chunk; t+L-1 and t+L are at most one chunk further (L < Lc)
        int c = (int)((w * 32) / p.Lc);
        if (c >= p.nchunks) c = p.nchunks - 1;
        const int64_t cend = (c == p.nchunks - 1) ? T : (int64_t)(c + 1) * p.Lc;
        const int64_t tc = t < T ? t : T - 1;
        const int64_t te = tc + L - 1;  // the chain entered at t ends here
        const int64_t tx = tc + L;      // for xi: the chain entered at t+1 ends here
        const int ce = (te >= cend) ? c + 1 : c, cx = (tx >= cend) ? c + 1 : c;
        const bool e_in = te <= T - 1, x_in = tx <= T - 1, has_next = tc <= T - 2;
        o.kap = p.kappa[c];
        o.lam = p.lambda[c];
        o.lame = p.lambda[e_in ? ce : c];
        o.lamx = p.lambda[x_in ? cx : c];
        o.vLG = p.LG[tc];
        o.vLH = p.LH[tc];
#pragma unroll
        for (int i = 0; i < N; i++) {
            o.vLQ[i] = p.LQ[(size_t)i * T + tc];
            o.vLEe[i] = p.LE[(size_t)i * T + (e_in ? te : T - 1)];
            o.vLEx[i] = p.LE[(size_t)i * T + (x_in ? tx : T - 1)];
            o.vFn[i] = p.Fg[(size_t)i * T + (has_next ? tc + 1 : T - 1)];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = lane + 32 * q;
            const int64_t g = w * 32 + k;
            o.yk[q] = (k < 32 + L && g < T) ? p.y[g] : 0.0;
        }
    };
    WinLoads cur, nxt;
    if (gw < nwin) issue(gw, cur);
    for (int64_t w = gw; w < nwin; w += nw) {
        if (w + nw < nwin) issue(w + nw, nxt);
        const int64_t t = w * 32 + lane;
        const bool ok = t < T && t >= p.st_lo && t < p.st_hi;
        const int64_t tc = t < T ? t : T - 1;
        const bool e_in = tc + L - 1 <= T - 1, x_in = tc + L <= T - 1, has_next = tc <= T - 2;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (lane + 32 * q < 32 + L) ysm[lane + 32 * q] = cur.yk[q];
        const double yv = cur.yk[0];
        double pi[N];
        double pmax = 0.0;
        if (ok) {
            // exp underflows to zero below -745.2: skip it there (a nearly silent neuron's statistics are sums
            // of very small terms whose RELATIVE accuracy matters -- lp = log(xi / gamma0) -- so nothing larger
            // may be dropped)
            const double g0 = exp(cur.vLG + cur.kap + cur.vLH + cur.lam - lS);
            a_g0all += g0;
            if (has_next) a_g0 += g0;
            a_y += yv;
            a_y2 += yv * yv;
#pragma unroll
            for (int i = 0; i < N; i++) {
                const double le = e_in ? cur.vLEe[i] + cur.lame : 0.0;
                const double ap = cur.vLQ[i] + cur.kap + le - lS;
                pi[i] = ap > -745.5 ? exp(ap) : 0.0;
                // S0tot counts the chains that run their whole length inside the recording; the last L-1 entry times
                // are added per phase by the finalize kernel (sums of positive terms only: subtracting them from a
                // total instead cancels catastrophically when a neuron's mass sits at the very end of the recording)
                if (e_in) a_s0[i] += pi[i];
                pmax = fmax(pmax, pi[i]);
                if (has_next) {
                    const double lex = x_in ? cur.vLEx[i] + cur.lamx : 0.0;
                    const double ax = cur.vLG + cur.kap + lH[i] + cur.vFn[i] + lex - lS;
                    if (ax > -745.5) a_xi[i] += exp(ax);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; i++) pi[i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < N; i++) pism[i * 32 + lane] = pi[i];
        __syncwarp();
        // S1[i][lag] += pi_i(t0) * y[t0 + lag]; entries below 1e-30 cannot change a double sum >= ~1
        unsigned sig = __ballot_sync(0xffffffffu, pmax > 1e-30);
        while (sig) {
            const int k = __ffs(sig) - 1;
            sig &= sig - 1;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int lag = lane + 32 * q;
                if (lag < L) {
                    const double yy = ysm[k + lag];
#pragma unroll
                    for (int i = 0; i < N; i++) a_s1[i][q] = fma(pism[i * 32 + k], yy, a_s1[i][q]);
                }
            }
        }
        __syncwarp();
        cur = nxt;
    }
    // block reduction (fixed order -> deterministic)
    __syncthreads();
    double *red = smem_d;  // reuse: [WPB][pstride]
    auto wsum = [&](double v) {
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        return v;
    };
    double r_g0 = wsum(a_g0), r_y = wsum(a_y), r_y2 = wsum(a_y2), r_g0all = wsum(a_g0all);
    double *mine = red + (size_t)warp * p.pstride;
    if (lane == 0) {
        mine[0] = r_g0;
        mine[1] = r_y;
        mine[2] = r_y2;
        mine[3] = r_g0all;
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
        double rx = wsum(a_xi[i]), rs = wsum(a_s0[i]);
        if (lane == 0) {
            mine[4 + i] = rx;
            mine[4 + N + i] = rs;
        }
#pragma unroll
        for (int q = 0; q < 3; q++) mine[4 + 2 * N + i * S1_LAGS + lane + 32 * q] = a_s1[i][q];
    }
    __syncthreads();
    double *dst = p.part + (size_t)blockIdx.x * p.pstride;
    for (int k = threadIdx.x; k < p.pstride; k += blockDim.x) {
        double v = 0.0;
        for (int w2 = 0; w2 < WPB; w2++) v += red[(size_t)w2 * p.pstride + k];
        dst[k] = v;
    }
}

// Column sums of the per-CTA statistic partials in a fixed order: 32 columns x 32 row groups per CTA.
__global__ void __launch_bounds__(1024) em_reduce(EmParams p) {
    __shared__ double red[32][33];
    const int col = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + col;
    const int rows = (p.nblk + 31) / 32;
    double v = 0.0;
    if (k < p.pstride) {
        const int b1 = (g + 1) * rows < p.nblk ? (g + 1) * rows : p.nblk;
        for (int b = g * rows; b < b1; b++) v += p.part[(size_t)b * p.pstride + k];
    }
    red[g][col] = v;
    __syncthreads();
    if (g == 0 && k < p.pstride) {
        double t = 0.0;
        for (int q = 0; q < 32; q++) t += red[q][col];
        p.tot[k] = t;
    }
}

// out layout: [0] sigma [1] loglik [2..2+N) lp  then mu [K*N] then pp [ns]
template <int N>
__global__ void __launch_bounds__(1024) em_finalize(EmParams p) {
    extern __shared__ __align__(16) double sm[];  // tot[pstride] then S0[N][S1_LAGS]
    const RingLayout &RL = p.RL;
    const int L = RL.L, K = L + 1;
    const int64_t T = p.T;
    double *tot = sm;
    double *S0 = sm + p.pstride;
    for (int k = threadIdx.x; k < p.pstride; k += blockDim.x) tot[k] = p.tot[k];
    __syncthreads();
    const double lS = p.lS[0];
    const double lam0 = p.lambda[0];
    const double kapl = p.kappa[p.nchunks - 1];
    double *S1 = tot + 4 + 2 * N;
    // boundary posteriors, computed once in parallel:
    //   piv[i][r0]: chains already running at t = 0 (entered at -r0, r0 = 1..L-1)
    //   pie[i][k] : chains entered at t0 = T-1-k (k = 0..L-2), which run past the end
    double *piv = S0 + (size_t)N * S1_LAGS, *pie = piv + (size_t)N * S1_LAGS;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, r = idx % L;
        double v = 0.0, w2 = 0.0;
        if (r >= 1) v = exp(p.LQneg[i * L + r] + p.LE[(size_t)i * T + (L - 1 - r)] + lam0 - lS);
        if (r <= L - 2) w2 = exp(p.LQ[(size_t)i * T + (T - 1 - r)] + kapl - lS);
        piv[i * S1_LAGS + r] = v;
        pie[i * S1_LAGS + r] = w2;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, sph = idx % L;  // 0-based phase: template row sph+1
        double s0 = tot[4 + N + i], s1 = S1[i * S1_LAGS + sph];
        // of the chains entered in the last L-1 samples (t0 = T-1-k), those with k >= sph reach phase sph
        for (int k = sph; k <= L - 2; k++) s0 += pie[i * S1_LAGS + k];
        // chains already running at t = 0 reach phases >= r0; at phase sph they sit on y[sph - r0]
        for (int r0 = 1; r0 <= sph; r0++) {
            const double pi = piv[i * S1_LAGS + r0];
            s0 += pi;
            s1 = fma(pi, p.y[sph - r0], s1);
        }
        S0[i * S1_LAGS + sph] = s0;
        S1[i * S1_LAGS + sph] = s1;
    }
    __syncthreads();
    double *out = p.out;
    double *mu = out + 2 + N, *pp = mu + (size_t)K * N;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, sph = idx % L;
        mu[(sph + 1) + (size_t)K * i] = S1[i * S1_LAGS + sph] / S0[i * S1_LAGS + sph];  // src/baumwelch.jl:283-287
        // pp = gamma[:,1] (log), src/baumwelch.jl:263
        pp[1 + i * L + sph] = p.LQneg[i * L + sph] + p.LE[(size_t)i * T + (L - 1 - sph)] + lam0 - lS;
    }
    if (threadIdx.x < N) mu[(size_t)K * threadIdx.x] = 0.0;  // row 1 stays 0 (src/baumwelch.jl:268)
    // S1^2 / S0 per template sample (the divisions in parallel; summed in a fixed order below)
    double *qterm = S0 + (size_t)N * S1_LAGS;  // piv is no longer needed
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, sph = idx % L;
        const double s1 = S1[i * S1_LAGS + sph];
        qterm[i * S1_LAGS + sph] = s1 * s1 / S0[i * S1_LAGS + sph];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pp[0] = p.LG[0] + p.LH[0] + lam0 - lS;
        double q = 0.0;
        for (int i = 0; i < N; i++)
            for (int sph = 0; sph < L; sph++) q += qterm[i * S1_LAGS + sph];
        // sigma^2 = sum_t sum_j gamma (y - m_j_new)^2 / sum gamma, with m_noise_new = 0 and
        // m_(i,s)_new = S1/S0  =>  (sum y^2 - sum S1^2/S0) / T      (src/baumwelch.jl:288-307)
        out[0] = sqrt((tot[2] - q) / (double)T);
        const double *sc = p.model + RL.scal;
        const double w_nn = sc[0], c_emit = sc[1], two_s2 = sc[2], m0 = sc[3];
        const double ssq = tot[2] - 2.0 * m0 * tot[1] + (double)T * m0 * m0;
        out[1] = (double)T * c_emit - ssq / two_s2 + (double)(T - 1) * w_nn + lS;
        for (int i = 0; i < N; i++) out[2 + i] = log(tot[4 + i]) - log(tot[0]);  // xb[2:end], :254-265
        double *cnt_out = pp + p.ns;
        cnt_out[0] = (double)p.counters[0];
        cnt_out[1] = (double)p.counters[1];
    }
}

// ---------------------------------------------------------------------------
// Time-sharded E/M step (contiguous spans of the recording on different GPUs).  Every shard runs the whole
// E-step on its span PLUS one ghost chunk on either side as a stand-alone problem: sum-product messages forget their
// start exponentially, so inside the main span the local posteriors are the global ones -- which is VERIFIED, not
// assumed: the forward / backward boundary vectors two neighbours hold for the same instant must agree up to a constant
// (em_shard_boundaries; compared by the caller after one all-gather).  The sufficient statistics are accumulated over the
// main span only (em_stats with [st_lo, st_hi)), extended by the end-of-recording corrections on the first / last shard
// and by gamma[:,1] on the first (em_shard_pack), summed over the shards by ONE all-reduce, and finalised identically on
// every rank (em_shard_finalize).
//   xvec layout: tot[pstride] | S0adj[N][S1_LAGS] | S1adj[N][S1_LAGS] | pp[ns] | 2 spare
// ---------------------------------------------------------------------------
__host__ __device__ inline int em_xvec_len(int N, int ns) { return (4 + 2 * N + N * S1_LAGS) + 2 * N * S1_LAGS + ns + 2; }

template <int N>
__global__ void __launch_bounds__(1024) em_shard_pack(EmParams p, int first, int last, double *xvec) {
    extern __shared__ __align__(16) double sm[];
    const RingLayout &RL = p.RL;
    const int L = RL.L;
    const int64_t T = p.T;
    const int nx = em_xvec_len(N, p.ns);
    for (int k = threadIdx.x; k < nx; k += blockDim.x) xvec[k] = k < p.pstride ? p.tot[k] : 0.0;
    const double lS = p.lS[0], lam0 = p.lambda[0], kapl = p.kappa[p.nchunks - 1];
    double *piv = sm, *pie = sm + (size_t)N * S1_LAGS;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, r = idx % L;
        double v = 0.0, w2 = 0.0;
        if (first && r >= 1) v = exp(p.LQneg[i * L + r] + p.LE[(size_t)i * T + (L - 1 - r)] + lam0 - lS);  // running at t = 0
        if (last && r <= L - 2) w2 = exp(p.LQ[(size_t)i * T + (T - 1 - r)] + kapl - lS);                    // run past the end
        piv[i * S1_LAGS + r] = v;
        pie[i * S1_LAGS + r] = w2;
    }
    __syncthreads();
    double *S0adj = xvec + p.pstride, *S1adj = S0adj + (size_t)N * S1_LAGS, *pp = S1adj + (size_t)N * S1_LAGS;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, sph = idx % L;
        double a0 = 0.0, a1 = 0.0;
        for (int k = sph; k <= L - 2; k++) a0 += pie[i * S1_LAGS + k];  // (zero unless this is the last shard)
        for (int r0 = 1; r0 <= sph; r0++) {
            const double pi = piv[i * S1_LAGS + r0];
            a0 += pi;
            if (first) a1 = fma(pi, p.y[sph - r0], a1);
        }
        S0adj[i * S1_LAGS + sph] = a0;
        S1adj[i * S1_LAGS + sph] = a1;
        if (first) pp[1 + i * L + sph] = p.LQneg[i * L + sph] + p.LE[(size_t)i * T + (L - 1 - sph)] + lam0 - lS;
    }
    if (threadIdx.x == 0 && first) pp[0] = p.LG[0] + p.LH[0] + lam0 - lS;
}

// The four boundary vectors of a shard in its own (chunk-0 consistent) normalisation: forward at main_begin and at
// main_end, backward at main_begin and at main_end; a vector that does not exist (no ghost on that side) is zero.
__global__ void em_shard_boundaries(EmParams p, int c_mb, int c_me, double *out /*[4][bvec]*/) {
    const int n = p.bvec;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        out[k] = c_mb >= 1 ? p.EBf[(size_t)(c_mb - 1) * n + k] + p.kappa[c_mb - 1] : 0.0;
        out[n + k] = (c_me >= 1 && c_me < p.nchunks) ? p.EBf[(size_t)(c_me - 1) * n + k] + p.kappa[c_me - 1] : 0.0;
        out[2 * n + k] = c_mb >= 1 ? p.EBb[(size_t)c_mb * n + k] + p.lambda[c_mb] : 0.0;
        out[3 * n + k] = c_me < p.nchunks ? p.EBb[(size_t)c_me * n + k] + p.lambda[c_me] : 0.0;
    }
    if (threadIdx.x == 0) out[4 * n] = p.lS[0];
}

// out layout as em_finalize: [0] sigma [1] loglik [2..2+N) lp  then mu [K*N] then pp [ns]
template <int N>
__global__ void __launch_bounds__(1024)
    em_shard_finalize(const double *xsum, int L, int ns, int pstride, double T_glob, double lS_glob, double w_nn,
                      double c_emit, double two_s2, double m0, double *out) {
    extern __shared__ __align__(16) double sm[];
    const int K = L + 1;
    const double *tot = xsum, *S1t = xsum + 4 + 2 * N, *S0adj = xsum + pstride, *S1adj = S0adj + (size_t)N * S1_LAGS,
                 *ppx = S1adj + (size_t)N * S1_LAGS;
    double *qterm = sm;
    double *mu = out + 2 + N, *pp = mu + (size_t)K * N;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        const int i = idx / L, sph = idx % L;
        const double s0 = tot[4 + N + i] + S0adj[i * S1_LAGS + sph];
        const double s1 = S1t[i * S1_LAGS + sph] + S1adj[i * S1_LAGS + sph];
        mu[(sph + 1) + (size_t)K * i] = s1 / s0;  // src/baumwelch.jl:283-287
        qterm[idx] = s1 * s1 / s0;
    }
    for (int k = threadIdx.x; k < ns; k += blockDim.x) pp[k] = ppx[k];  // gamma[:,1], src/baumwelch.jl:263
    if (threadIdx.x < N) mu[(size_t)K * threadIdx.x] = 0.0;             // row 1 stays 0 (src/baumwelch.jl:268)
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int k = 0; k < N * L; k++) q += qterm[k];
        out[0] = sqrt((tot[2] - q) / T_glob);  // src/baumwelch.jl:288-307
        const double ssq = tot[2] - 2.0 * m0 * tot[1] + T_glob * m0 * m0;
        out[1] = T_glob * c_emit - ssq / two_s2 + (T_glob - 1.0) * w_nn + lS_glob;
        for (int i = 0; i < N; i++) out[2 + i] = log(tot[4 + i]) - log(tot[0]);  // xb[2:end], :254-265
    }
}

// ---------------------------------------------------------------------------
// Dense alpha / beta on request (forward / backward of src/baumwelch.jl:25-51, 73-98 as
// stand-alone calls): materialised from the semi-Markov quantities.
//   alpha[noise, t]  = Z_t + lg_t                     Z_t = sum_{tau<=t} q_tau(noise) + t * w_nn
//   alpha[(i,s), t]  = Z_t + lp_{t-s}(i) + sum_{r<=s} (a[i][r] y[t-s+r] + bw[i][r])
//   beta[noise, t]   = (Z_{T-1} - Z_t) + lh_t
//   beta[(i,s), t]   = (Z_{T-1} - Z_t) + le_{t-s+L-1}(i) + sum_{r>s} (a[i][r] y[t-s+r] + bw[i][r])
// One thread per chain entry time t0 walks its chain and writes one column entry per step.
// ---------------------------------------------------------------------------
constexpr int ZS_ITEMS = 16;  // samples per thread in the Z prefix scan (256 threads -> 4096 per block)

__global__ void __launch_bounds__(256) fb_zscan(EmParams p, double *Zs, double *bsum) {
    __shared__ double wsum[8];
    const double *sc = p.model + p.RL.scal;
    const double w_nn = sc[0], c_emit = sc[1], two_s2 = sc[2], m0 = sc[3];
    const int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * ZS_ITEMS;
    double v[ZS_ITEMS];
    double loc = 0.0;
#pragma unroll
    for (int k = 0; k < ZS_ITEMS; k++) {
        const int64_t t = base + k;
        double inc = 0.0;
        if (t < p.T) {
            const double dd = p.y[t] - m0;
            inc = (c_emit - (dd * dd) / two_s2) + (t > 0 ? w_nn : 0.0);
        }
        loc += inc;
        v[k] = loc;
    }
    const double incl = block_scan_256(loc, wsum);
    const double off = incl - loc;
#pragma unroll
    for (int k = 0; k < ZS_ITEMS; k++)
        if (base + k < p.T) Zs[base + k] = v[k] + off;
    if (threadIdx.x == 255) bsum[blockIdx.x] = incl;
}

__global__ void fb_zscan_blocks(double *bsum, int nb) {  // exclusive scan of the block totals, in place
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0;
        for (int b = 0; b < nb; b++) {
            double v = bsum[b];
            bsum[b] = a;
            a += v;
        }
    }
}

__device__ __forceinline__ double zval(const double *Zs, const double *boff, int64_t t) {
    return Zs[t] + boff[t / (256 * ZS_ITEMS)];
}

template <int N, bool BETA>
__global__ void __launch_bounds__(128) fb_dense(EmParams p, const double *Zs, const double *boff, double *out) {
    const RingLayout &RL = p.RL;
    const int L = RL.L, NP = RL.NP, ns = p.ns;
    const int64_t T = p.T;
    const double *A = p.model + RL.A, *BW = p.model + RL.BW, *B0 = p.model + RL.B0;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - (L - 1);
    if (t0 > T - 1) return;
    const double ZT = zval(Zs, boff, T - 1);
    auto chunk_of = [&](int64_t t) {
        int c = (int)(t / p.Lc);
        return c >= p.nchunks ? p.nchunks - 1 : c;
    };
    if (t0 >= 0) {  // noise row
        const int c = chunk_of(t0);
        const double z = zval(Zs, boff, t0);
        out[(size_t)ns * t0] = BETA ? (ZT - z) + p.LH[t0] + p.lambda[c] : z + p.LG[t0] + p.kappa[c];
    }
    const int r_lo = t0 < 0 ? (int)(-t0) : 0;                        // first phase with a sample
    const int r_hi = (t0 + L - 1 <= T - 1) ? L - 1 : (int)(T - 1 - t0);  // last phase with a sample
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        if (!BETA) {
            double acc = t0 >= 0 ? (p.LQ[(size_t)i * T + t0] - p.Fg[(size_t)i * T + t0]) + p.kappa[chunk_of(t0)] : 0.0;
            for (int r = r_lo; r <= r_hi; r++) {
                const int64_t t = t0 + r;
                // a chain already running at t = 0 has no transition INTO its first sample
                const double bw = (t0 < 0 && r == r_lo) ? B0[r * NP + i] : BW[r * NP + i];
                acc += fma(A[r * NP + i], p.y[t], bw);
                out[(size_t)ns * t + 1 + i * L + r] = zval(Zs, boff, t) + acc;
            }
        } else {
            const int64_t te = t0 + L - 1;
            double acc = te <= T - 1 ? p.LE[(size_t)i * T + te] + p.lambda[chunk_of(te)] : 0.0;
            for (int r = r_hi; r >= r_lo; r--) {
                const int64_t t = t0 + r;
                out[(size_t)ns * t + 1 + i * L + r] = (ZT - zval(Zs, boff, t)) + acc;
                acc += fma(A[r * NP + i], p.y[t], BW[r * NP + i]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// cudaFuncSetAttribute costs a few microseconds and the E/M step is launch-latency sensitive: set once per host thread
// and per (device, value)
#define EM_SMEM_ATTR(bytes, ...)                                                                                     \
    do {                                                                                                             \
        static thread_local size_t set_ = 0;                                                                         \
        static thread_local int dev_ = -1;                                                                           \
        int cur_ = 0;                                                                                                \
        HMM_CUDA(cudaGetDevice(&cur_));                                                                              \
        if (set_ != (size_t)(bytes) || dev_ != cur_) {                                                               \
            HMM_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            set_ = (size_t)(bytes);                                                                                  \
            dev_ = cur_;                                                                                             \
        }                                                                                                            \
    } while (0)

// mode: 0 = full E/M step; 1 = forward quantities only; 2 = forward + backward quantities
// Warps of the forward and of the backward kernel that are co-resident on one SM (the smaller of the two):
// the chunk count is matched to it so that each pass runs as exactly one wave.
template <int N, int R>
static int em_warps_per_sm(const RingLayout &RL) {
    constexpr int WPB = 4;
    // (asked on every E/M step: the answer depends on the model shape only; one host thread = one device)
    static thread_local int c_L = -1, c_dev = -1, c_val = 0;
    int dev = 0;
    HMM_CUDA(cudaGetDevice(&dev));
    if (c_L == RL.L && c_dev == dev) return c_val;
    const size_t mdl_d = (RL.hot + 1) & ~1;
    const size_t sm_fwd = sizeof(double) * (mdl_d + (size_t)WPB * EmWarpSmem<N, R>::DOUBLES);
    const size_t sm_bwd = sizeof(double) * (mdl_d + (size_t)WPB * N * RING_Q);
    EM_SMEM_ATTR(sm_fwd, em_forward<N, R>);
    int nf = 0, nb = 0;
    HMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nf, em_forward<N, R>, 32 * WPB, sm_fwd));
    HMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, em_backward<N>, 32 * WPB, sm_bwd));
    const int n = nf < nb ? nf : nb;
    c_L = RL.L;
    c_dev = dev;
    c_val = (n > 0 ? n : 1) * WPB;
    return c_val;
}

template <int N, int R, int LPC>
static void em_fir_launch(EmParams &p, const double *hmdl, cudaStream_t st) {
    constexpr int WPB = 4;
    const size_t sm = sizeof(double) * (((p.RL.hot + 1) & ~1) + (size_t)WPB * EmWarpSmem<N, R>::TILE);
    EM_SMEM_ATTR(sm, em_fir<N, R, LPC>);
    FirCoef<N, LPC> coef{};
    if (LPC > 0)
        for (int r = 0; r < LPC; r++)
            for (int i = 0; i < N; i++) coef.a[r * N + i] = hmdl[p.RL.A + r * p.RL.NP + i];
    const int64_t nsw = (p.T + 32 * R - 1) / (32 * R);
    em_fir<N, R, LPC><<<(unsigned)((nsw + WPB - 1) / WPB), 32 * WPB, sm, st>>>(p, coef);
}

// Optional per-stage device timing (HMMCUDA_EM_TIMING=1): events between the launches, printed after the step.
struct EmStageTimes {
    static constexpr int MAXE = 10;
    cudaEvent_t ev[MAXE] = {};
    const char *name[MAXE] = {};
    int n = 0;
    bool on = false;
    void mark(const char *what, cudaStream_t st) {
        if (!on || n >= MAXE) return;
        if (!ev[n]) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        name[n++] = what;
    }
    void report() {
        if (!on || n < 2) return;
        cudaEventSynchronize(ev[n - 1]);
        for (int k = 1; k < n; k++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
            fprintf(stderr, "%s %.1f us%s", name[k], 1e3 * ms, k == n - 1 ? "\n" : " | ");
        }
        n = 0;
    }
};
static thread_local EmStageTimes g_em_times;

template <int N, int R>
static void em_launch(EmParams &p, const double *hmdl, cudaStream_t st, hmm_info *info, Timer &ttop, int mode,
                      double *alpha_out, double *beta_out, double *Zs, double *bsum, const EmShardOpts *sh) {
    NvtxRange nvtx_step(sh ? "hmm.em.shard_estep" : mode == 0 ? "hmm.em.step" : "hmm.em.dense_forward_backward");
    EmStageTimes &tm = g_em_times;
    tm.on = getenv("HMMCUDA_EM_TIMING") != nullptr;
    tm.n = 0;
    tm.mark("start", st);
    constexpr int WPB = 4;
    const size_t mdl_d = (p.RL.hot + 1) & ~1;
    const size_t sm_fwd = sizeof(double) * (mdl_d + (size_t)WPB * EmWarpSmem<N, R>::DOUBLES);
    const size_t sm_bwd = sizeof(double) * (mdl_d + (size_t)WPB * N * RING_Q);
    const size_t sm_stats = sizeof(double) * std::max<size_t>((size_t)WPB * (160 + N * 32), (size_t)WPB * p.pstride);
    const size_t sm_fin = sizeof(double) * ((size_t)p.pstride + std::max<size_t>(3 * (size_t)N * S1_LAGS, 4 * (size_t)p.pstride));
    EM_SMEM_ATTR(sm_fwd, em_forward<N, R>);
    EM_SMEM_ATTR(sm_stats, em_stats<N>);
    const int gridc = (p.nchunks + WPB - 1) / WPB;
    const int gchk = (p.nchunks * 32 + 127) / 128;
    const int dirs = mode != 1 ? 3 : 1;  // bit 0 forward, bit 1 backward
    const size_t sm_rep = sizeof(double) * (mdl_d + (size_t)N * RING_Q + EM_SCAN_SMEM);
    EM_SMEM_ATTR(sm_rep, em_fixup<N, R>);
    // FIR pass first: F_i(t0) for every sample, consumed by both recursions and by the statistics pass
    if (N >= 3 && N <= 5 && p.RL.L == 59)
        em_fir_launch<N, R, (N >= 3 && N <= 5) ? 59 : 0>(p, hmdl, st);
    else if (N >= 3 && N <= 5 && p.RL.L == 47)
        em_fir_launch<N, R, (N >= 3 && N <= 5) ? 47 : 0>(p, hmdl, st);
    else
        em_fir_launch<N, R, 0>(p, hmdl, st);
    tm.mark("fir", st);
    ttop.start();
    em_forward<N, R><<<gridc, 32 * WPB, sm_fwd, st>>>(p);
    ttop.stop();
    tm.mark("forward", st);
    // (the backward pass needs only the F scores, not the forward results: measured on a second stream it did
    // not overlap -- each pass already fills every SM with one wave of chunks -- so it simply follows)
    if (mode != 1) em_backward<N><<<gridc, 32 * WPB, sm_bwd, st>>>(p);
    tm.mark("backward", st);
    em_check<<<dim3(gchk, 2), 128, 0, st>>>(p, dirs);
    tm.mark("check", st);
    em_fixup<N, R><<<2, 1024, sm_rep, st>>>(p, dirs);
    tm.mark("fixup", st);
    if (info) info->kernel_launches += (mode != 1 ? 5 : 4);
    if (mode == 0 && sh) {  // time shard: statistics of the main span, packed for the all-reduce; boundary vectors
        em_stats<N><<<p.nblk, 32 * WPB, sm_stats, st>>>(p);
        em_reduce<<<(p.pstride + 31) / 32, 1024, 0, st>>>(p);
        const size_t sm_pack = sizeof(double) * 2 * (size_t)N * S1_LAGS;
        em_shard_pack<N><<<1, 1024, sm_pack, st>>>(p, sh->first ? 1 : 0, sh->last ? 1 : 0, sh->xvec_dev);
        em_shard_boundaries<<<1, 256, 0, st>>>(p, (int)(sh->st_lo / p.Lc), sh->last ? p.nchunks : (int)(sh->st_hi / p.Lc), sh->bnd_dev);
        if (info) info->kernel_launches += 4;
    } else if (mode == 0) {
        em_stats<N><<<p.nblk, 32 * WPB, sm_stats, st>>>(p);
        tm.mark("stats", st);
        EM_SMEM_ATTR(sm_fin, em_finalize<N>);
        em_reduce<<<(p.pstride + 31) / 32, 1024, 0, st>>>(p);
        tm.mark("reduce", st);
        em_finalize<N><<<1, 1024, sm_fin, st>>>(p);
        tm.mark("finalize", st);
        if (info) info->kernel_launches += 3;
    } else {
        const int nb = (int)((p.T + 256 * ZS_ITEMS - 1) / (256 * ZS_ITEMS));
        fb_zscan<<<nb, 256, 0, st>>>(p, Zs, bsum);
        fb_zscan_blocks<<<1, 32, 0, st>>>(bsum, nb);
        const int64_t nthreads = p.T + p.RL.L - 1;
        const int g = (int)((nthreads + 127) / 128);
        if (alpha_out) fb_dense<N, false><<<g, 128, 0, st>>>(p, Zs, bsum, alpha_out);
        if (beta_out) fb_dense<N, true><<<g, 128, 0, st>>>(p, Zs, bsum, beta_out);
    }
    HMM_CUDA(cudaGetLastError());
}

static void ring_em_core(const double *X_dev, int64_t T, const HostModel &M, EmResult *outp, cudaStream_t st,
                        hmm_info *info, int mode, double *alpha_out, double *beta_out, const EmShardOpts *sh = nullptr) {
    Workspace &ws = workspace();
    const int N = M.N, L = M.K - 1, K = M.K, ns = M.nstates;
    const int R = (N <= 4) ? 8 : 4, SW = 32 * R;
    RingLayout RL = ring_layout(N, L);
    // The E-step is latency-bound per warp (log-sum-exp chains), so short chunks and a short
    // warm-up pay: every boundary is verified (and repaired if needed) anyway.
    int64_t W = sh ? sh->W : (ring_config().warmup > 0 ? ring_config().warmup : 256);
    W = ((W + SW - 1) / SW) * SW;
    if (W < ((L + 32 + SW - 1) / SW) * SW) W = ((L + 32 + SW - 1) / SW) * SW;
    int64_t Lc = sh ? sh->Lc : ring_config().chunk_len;
    if (sh && (Lc % SW || Lc < W || sh->st_lo % Lc || (!sh->last && sh->st_hi % Lc)))
        fail(HMM_EINVAL, "time shard: chunk_len must be a multiple of %d and the main span chunk aligned", SW);
    if (Lc <= 0) {
        int dev = 0, sms = 148;
        HMM_CUDA(cudaGetDevice(&dev));
        HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int wps = 16;
        switch (N) {
            case 1: wps = em_warps_per_sm<1, 8>(RL); break;
            case 2: wps = em_warps_per_sm<2, 8>(RL); break;
            case 3: wps = em_warps_per_sm<3, 8>(RL); break;
            case 4: wps = em_warps_per_sm<4, 8>(RL); break;
            case 5: wps = em_warps_per_sm<5, 4>(RL); break;
            case 6: wps = em_warps_per_sm<6, 4>(RL); break;
            case 7: wps = em_warps_per_sm<7, 4>(RL); break;
            default: break;
        }
        Lc = (T + (int64_t)sms * wps - 1) / ((int64_t)sms * wps);
        if (Lc < 3 * W) Lc = 3 * W;
    }
    Lc = ((Lc + SW - 1) / SW) * SW;
    if (Lc < W) Lc = W;
    if (Lc < 256) Lc = 256;
    int nchunks = (int)((T + Lc - 1) / Lc);
    if (nchunks > 1 && T - (int64_t)(nchunks - 1) * Lc < RING_Q) nchunks--;

    double *hmdl = (double *)ws.pinned(0, sizeof(double) * RL.total);  // pinned: the upload below is truly async
    ring_pack(M, RL, hmdl);
    const int bvec = 1 + N * L;
    const int pstride = 4 + 2 * N + N * S1_LAGS;
    const int nblk = 148 * 3;  // one resident wave of the statistics pass (165 registers: 3 CTAs per SM)
    const int nout = 2 + N + K * N + ns + 2;  // ... + the two repair counters
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t r = off;
        off += (bytes + 255) & ~size_t(255);
        return r;
    };
    size_t o_model = carve(sizeof(double) * RL.total);
    size_t o_neg = carve(sizeof(double) * N * L);
    size_t o_b = carve(sizeof(double) * 4 * (size_t)nchunks * bvec);
    size_t o_flag = carve(sizeof(int) * 2 * (size_t)nchunks);
    size_t o_kl = carve(sizeof(double) * 2 * (size_t)nchunks);
    size_t o_dk = carve(sizeof(double) * 2 * (size_t)nchunks);
    size_t o_ls = carve(sizeof(double) * 2);
    size_t o_cnt = carve(sizeof(int) * 4);
    size_t o_part = carve(sizeof(double) * (size_t)nblk * pstride);
    size_t o_tot = carve(sizeof(double) * pstride);
    size_t o_out = carve(sizeof(double) * nout);
    char *base = (char *)ws.get(Workspace::CHUNKS, off);
    double *steps = (double *)ws.get(Workspace::FWDQ, sizeof(double) * (size_t)T * (3 * N + 2));
    HMM_CUDA(cudaMemcpyAsync(base + o_model, hmdl, sizeof(double) * RL.total, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemsetAsync(base + o_cnt, 0, sizeof(int) * 4, st));

    EmParams p{};
    p.y = X_dev;
    p.T = T;
    p.model = (const double *)(base + o_model);
    p.RL = RL;
    p.Lc = Lc;
    p.W = W;
    p.nchunks = nchunks;
    p.ns = ns;
    p.Fg = steps;
    p.LQ = steps + (size_t)N * T;
    p.LE = steps + (size_t)2 * N * T;
    p.LG = steps + (size_t)3 * N * T;
    p.LH = p.LG + T;
    p.LQneg = (double *)(base + o_neg);
    p.SBf = (double *)(base + o_b);
    p.EBf = p.SBf + (size_t)nchunks * bvec;
    p.SBb = p.EBf + (size_t)nchunks * bvec;
    p.EBb = p.SBb + (size_t)nchunks * bvec;
    p.bvec = bvec;
    p.flag_f = (int *)(base + o_flag);
    p.flag_b = p.flag_f + nchunks;
    p.kappa = (double *)(base + o_kl);
    p.lambda = p.kappa + nchunks;
    p.dk = (double *)(base + o_dk);
    p.lS = (double *)(base + o_ls);
    p.counters = (int *)(base + o_cnt);
    p.part = (double *)(base + o_part);
    p.tot = (double *)(base + o_tot);
    p.nblk = nblk;
    p.pstride = pstride;
    // em_finalize writes its few hundred doubles straight into mapped pinned host memory: no device-to-host
    // copy, one stream synchronisation per E/M step
    void *out_dev = nullptr;
    double *out_host = (double *)ws.pinned(1, sizeof(double) * nout, &out_dev);
    p.out = (double *)out_dev;
    p.dbg = getenv("HMMCUDA_EM_DBG") ? atoi(getenv("HMMCUDA_EM_DBG")) : 0;
    p.st_lo = sh ? sh->st_lo : 0;
    p.st_hi = sh ? sh->st_hi : T;
    (void)o_out;

    double *Zs = nullptr, *bsum = nullptr;
    if (mode != 0) {
        const size_t nb = (size_t)((T + 256 * ZS_ITEMS - 1) / (256 * ZS_ITEMS));
        Zs = (double *)ws.get(Workspace::PSCORE, sizeof(double) * ((size_t)T + nb + 8));
        bsum = Zs + T;
    }
    Timer ttop(st);
    switch (N) {
        case 1: em_launch<1, 8>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 2: em_launch<2, 8>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 3: em_launch<3, 8>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 4: em_launch<4, 8>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 5: em_launch<5, 4>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 6: em_launch<6, 4>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        case 7: em_launch<7, 4>(p, hmdl, st, info, ttop, mode, alpha_out, beta_out, Zs, bsum, sh); break;
        default: fail(HMM_EUNSUPPORTED, "ring E/M engine supports 1..%d neurons", RING_MAX_N);
    }
    if (mode != 0) {
        HMM_CUDA(cudaStreamSynchronize(st));
        return;
    }
    if (sh) {  // asynchronous: the caller synchronises after its collectives
        if (info) info->n_chunks = nchunks;
        return;
    }
    EmResult &out = *outp;
    HMM_CUDA(cudaStreamSynchronize(st));
    g_em_times.report();
    const double *h = out_host;
    const int cnt[2] = {(int)h[nout - 2], (int)h[nout - 1]};
    out.sigma = h[0];
    out.loglik = h[1];
    out.lp.assign(h + 2, h + 2 + N);
    out.mu.assign(h + 2 + N, h + 2 + N + (size_t)K * N);
    out.pp.assign(h + 2 + N + (size_t)K * N, h + 2 + N + (size_t)K * N + ns);
    if (info) {
        info->n_chunks = nchunks;
        info->fwd_repaired = cnt[0];
        info->bwd_repaired = cnt[1];
        info->top_kernel_ms = ttop.ms();
    }
}

void ring_em_run(const double *X_dev, int64_t T, const HostModel &M, EmResult &out, cudaStream_t st, hmm_info *info) {
    ring_em_core(X_dev, T, M, &out, st, info, 0, nullptr, nullptr);
}

// ---- time-sharded E/M step (see the kernels above) ----
// Chunk length / warm-up the E-step would choose for T_local samples on one GPU (one chunk per resident warp).
void ring_em_default_chunking(int N, int K, int64_t T_local, int64_t *Lc_out, int64_t *W_out) {
    const int L = K - 1;
    const int R = (N <= 4) ? 8 : 4, SW = 32 * R;
    RingLayout RL = ring_layout(N, L);
    int64_t W = 256;
    W = ((W + SW - 1) / SW) * SW;
    if (W < ((L + 32 + SW - 1) / SW) * SW) W = ((L + 32 + SW - 1) / SW) * SW;
    int dev = 0, sms = 148;
    HMM_CUDA(cudaGetDevice(&dev));
    HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int wps = 16;
    switch (N) {
        case 1: wps = em_warps_per_sm<1, 8>(RL); break;
        case 2: wps = em_warps_per_sm<2, 8>(RL); break;
        case 3: wps = em_warps_per_sm<3, 8>(RL); break;
        case 4: wps = em_warps_per_sm<4, 8>(RL); break;
        case 5: wps = em_warps_per_sm<5, 4>(RL); break;
        case 6: wps = em_warps_per_sm<6, 4>(RL); break;
        case 7: wps = em_warps_per_sm<7, 4>(RL); break;
        default: break;
    }
    int64_t Lc = (T_local + (int64_t)sms * wps - 1) / ((int64_t)sms * wps);
    if (Lc < 3 * W) Lc = 3 * W;
    Lc = ((Lc + 255) / 256) * 256;
    *Lc_out = Lc;
    *W_out = W;
}
int ring_em_xvec_len(int N, int nstates) { return em_xvec_len(N, nstates); }
int ring_em_bnd_len(int N, int K) { return 4 * (1 + N * (K - 1)) + 1; }

void ring_em_shard_estep(const double *X_dev, int64_t T_local, const HostModel &M, const EmShardOpts &sh, cudaStream_t st,
                         hmm_info *info) {
    ring_em_core(X_dev, T_local, M, nullptr, st, info, 0, nullptr, nullptr, &sh);
}

void ring_em_shard_mstep(const double *xsum_dev, const HostModel &M, int64_t T_glob, double lS_glob, EmResult &out,
                         cudaStream_t st) {
    const int N = M.N, L = M.K - 1, K = M.K, ns = M.nstates;
    const int pstride = 4 + 2 * N + N * S1_LAGS;
    const int nout = 2 + N + K * N + ns + 2;
    Workspace &ws = workspace();
    void *out_dev = nullptr;
    double *out_host = (double *)ws.pinned(1, sizeof(double) * nout, &out_dev);
    const double LOG2PI = 0.9189385332046727;
    const double w_nn = M.ring.w_nn, c_emit = (-LOG2PI) - M.lsig, two_s2 = 2 * (M.sigma * M.sigma), m0 = M.m[0];
    const size_t smf = sizeof(double) * (size_t)N * S1_LAGS;
#define HMM_EMSH(NN)                                                                                                \
    case NN:                                                                                                        \
        em_shard_finalize<NN><<<1, 1024, smf, st>>>(xsum_dev, L, ns, pstride, (double)T_glob, lS_glob, w_nn, c_emit, \
                                                    two_s2, m0, (double *)out_dev);                                 \
        break;
    switch (N) {
        HMM_EMSH(1) HMM_EMSH(2) HMM_EMSH(3) HMM_EMSH(4) HMM_EMSH(5) HMM_EMSH(6) HMM_EMSH(7)
        default: fail(HMM_EUNSUPPORTED, "ring E/M engine supports 1..%d neurons", RING_MAX_N);
    }
#undef HMM_EMSH
    HMM_CUDA(cudaGetLastError());
    HMM_CUDA(cudaStreamSynchronize(st));
    const double *h = out_host;
    out.sigma = h[0];
    out.loglik = h[1];
    out.lp.assign(h + 2, h + 2 + N);
    out.mu.assign(h + 2 + N, h + 2 + N + (size_t)K * N);
    out.pp.assign(h + 2 + N + (size_t)K * N, h + 2 + N + (size_t)K * N + ns);
}

void ring_fb_dense_run(const double *X_dev, int64_t T, const HostModel &M, double *alpha_dev, double *beta_dev,
                       cudaStream_t st) {
    ring_em_core(X_dev, T, M, nullptr, st, nullptr, beta_dev ? 2 : 1, alpha_dev, beta_dev);
}

}  // namespace hmm
