// ring_viterbi.cu -- time-parallel exact Viterbi decode for non-overlap ring
// models (src/viterbi.jl:44-98 semantics).
//
// A recording is cut into chunks; ONE WARP owns one chunk and runs, without
// any block-level synchronisation,
//   (1) a register-blocked FP64 FIR (matched filter) F_i(t0) over a
//       super-window of 32*R samples staged through shared memory with
//       cp.async, then
//   (2) the max-plus recursion over the N+1 decision states, 32 time steps at
//       a time: one lane per step, a warp-shuffle prefix-max for the noise
//       state, backpointers packed 4 bits per decision state (one u32 per
//       step) plus one "noise entered from a tail" bit-mask word per 32 steps.
// Chunks other than the first start SPECULATIVELY from an empty state W samples
// early; afterwards every chunk boundary is VERIFIED (the speculative boundary
// vector must equal the true one up to a constant) and any chunk that fails
// is re-run from the true boundary vector, so the decode is exact, not
// approximate.  The traceback is parallel in the same way (speculative
// look-ahead + verification + repair).  The first L+1 samples are decoded by
// the faithful engine in the reference's exact arithmetic, which reproduces
// the structural ties at t=2.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <memory>
#include <vector>

#include "faithful_dev.cuh"
#include "ring_common.cuh"

namespace hmm {

RingConfig &ring_config() {
    static RingConfig c;
    return c;
}

bool ring_supported(const HostModel &M, int64_t T) {
    return M.is_ring && M.N <= RING_MAX_N && (M.K - 1) <= RING_MAX_L && (M.K - 1) >= 2 && T >= 2048;
}

void ring_pack(const HostModel &M, const RingLayout &R, double *dst) {
    const RingParams &P = M.ring;
    const int N = R.N, L = R.L;
    for (int i = 0; i < R.total; i++) dst[i] = 0.0;
    const double m0 = M.m[0];
    const double s2 = M.sigma * M.sigma;
    for (int i = 0; i < N; i++) {
        double bc = 0.0;
        for (int r = 0; r < L; r++) {
            double mi = M.m[1 + i * L + r];
            double a = (mi - m0) / s2;
            double b = (m0 - mi) * (m0 + mi) / (2 * s2);
            double bw = b + (r > 0 ? P.w_c[(size_t)i * (L - 1) + r - 1] - P.w_nn : 0.0);
            dst[R.A + r * R.NP + i] = a;
            dst[R.BW + r * R.NP + i] = bw;
            dst[R.B0 + r * R.NP + i] = b;
            bc += bw;
        }
        {
            double suf = 0.0;
            for (int r = L - 1; r >= 0; r--) {
                suf += dst[R.BW + r * R.NP + i];
                dst[R.BWsuf + r * R.NP + i] = suf;
            }
        }
        dst[R.Bc + i] = bc;
        dst[R.eG + i] = P.w_tn[i] - P.w_nn;
        dst[R.eH + i] = P.w_nh[i] - P.w_nn;
        for (int j = 0; j < N; j++)
            dst[R.eT + j * R.NP + i] = (i == j) ? -std::numeric_limits<double>::infinity()
                                                : P.w_th[(size_t)j * N + i] - P.w_nn;
    }
    const double LOG2PI = 0.9189385332046727;
    dst[R.scal + 0] = P.w_nn;
    dst[R.scal + 1] = (-LOG2PI) - M.lsig;
    dst[R.scal + 2] = 2 * s2;
    dst[R.scal + 3] = m0;
    dst[R.scal + 4] = M.sigma;
    // bounds used by the decision fast path: max finite tail->head weight, min noise->head weight
    double eTmax = -std::numeric_limits<double>::infinity(), eHmin = std::numeric_limits<double>::infinity();
    for (int i = 0; i < N; i++) {
        eHmin = std::min(eHmin, dst[R.eH + i]);
        for (int j = 0; j < N; j++)
            if (i != j) eTmax = std::max(eTmax, dst[R.eT + j * R.NP + i]);
    }
    dst[R.scal + 5] = eTmax;  // -inf when N == 1
    dst[R.scal + 6] = eHmin;
    // Liveness margin: a pending chain score Q_j created at t0 can influence a later decision
    // only if Q_j + cL[j] > G_t0 (G is non-decreasing): it must beat the noise path either into
    // the noise state (eG) or into some head (eT[j][i] - eH[i]).  1e-9 covers all rounding.
    for (int j = 0; j < N; j++) {
        double c = dst[R.eG + j];
        for (int i = 0; i < N; i++)
            if (i != j) c = std::max(c, dst[R.eT + j * R.NP + i] - dst[R.eH + i]);
        dst[R.cL + j] = c + 1e-9;
    }
    for (int i = 0; i < N; i++) {
        dst[R.xG + i] = std::exp(dst[R.eG + i]);
        dst[R.xH + i] = std::exp(dst[R.eH + i]);
        for (int j = 0; j < N; j++) dst[R.xT + j * R.NP + i] = (i == j) ? 0.0 : std::exp(dst[R.eT + j * R.NP + i]);
    }
}

// ---------------------------------------------------------------------------
struct VitParams {
    const double *y;       // [T x C]
    int64_t T, y_stride;
    const double *model;   // [C x RL.total]
    RingLayout RL;
    int64_t Lc, W;         // chunk length, warm-up / look-ahead (multiples of the super-window)
    int nchunks;           // per channel
    int64_t Lc_t;          // traceback chunk length (Lc / tfac): the traceback is latency-bound per chunk, so it
    int nchunks_t, tfac;   // uses shorter chunks than the forward pass; boundaries stay aligned
    int ns;                // nstates
    uint32_t *dec;         // [C x T]      packed decisions
    uint32_t *nzmask;      // [C x ceil(T/32)]
    double *SB, *EB;       // [C x nchunks x bvec]  boundary vectors (start: speculative, end: true)
    int bvec;              // 1 + N*L
    const char *fblob;     // faithful model blobs (prologue in the reference's arithmetic)
    size_t fblob_stride;
    FaithfulLayout FL;
    double *T1pro;         // [C x ns x (L+1)] prologue trellis, written by chunk 0's warp
    double *Pfin;          // [C x N x RING_Q]  P of the last steps (final chunk)
    double *Gfin;          // [C]
    int *fwd_flag;         // [C x nchunks] boundary mismatch flags
    int *counters;         // [C x 4]: 0 fwd repaired, 1 trace repaired
    // trace
    int16_t *T2pro;        // [C x (L+1) x 8] prologue backpointers of the DECISION states (noise, heads), 1-based source
    int16_t *xend;         // [C] final state (0-based), written by the last chunk's warp
    int16_t *x;            // [T x C]
    int64_t x_stride;
    long long *own_start, *look_end;  // [C x nchunks] encoded states
    int *tr_flag;
    int64_t x_lo, x_hi;    // x is written only for local steps in [x_lo, x_hi) (ghost chunks of a shard are not)
    int first_prologue;    // chunk 0 starts from the reference's initial condition (else: ghost chunk of a time shard)
    int last_true_end;     // the sequence really ends at T (else: ghost chunk; traceback starts speculatively)
    int dbg_flag_every;    // > 0: the boundary checks also flag every k-th chunk (tests force the repair paths with it)
    unsigned *sync_cnt;    // [C x 4] arrival counters of the fused check+repair kernels ("last CTA continues"), self-resetting
    double *res_host;      // [C x 4] device alias of mapped pinned host memory: ll, chunks repaired (forward, traceback); nullable
    int ch0;               // channel offset of this launch (long recordings: one channel per launch out of a C-channel plan)
    // ll = sum_t T1[x_t, t] assembled from pieces the decode produces anyway (see ll_assemble):
    double *ll_noise;      // [C x nchunks x 2] per forward chunk (and FIR producer warp of its slot): sum over its main range of (Tg - g) (y_g - m0)^2
    double *ll_spike;      // [C x nchunks_t] per traceback chunk: sum over its steps of (Tg - g) * (normalised increment)
    double *ll_out;        // [C]
    int64_t ll_lo, ll_hi;  // local sample range the sum runs over (chunk aligned; a shard's main span)
    int64_t t_off, T_glob; // global time of local step t is t + t_off; the weights are T_glob - (t + t_off)
    int ll_with_p0;        // add (T_glob - 1) * T1[x_1, 1] (the first shard / the whole recording)
    int want_ll;
};

enum StartKind { START_PROLOGUE = 0, START_SPEC = 1, START_EXACT = 2 };

template <int N, int R>
struct WarpSmem {
    using G = FirGeom<R>;
    // the F tile is written only after the FIR has consumed the y tile, so the two
    // share storage; the prologue's Z scratch (chunk 0 only) lives behind the ring
    static constexpr int TILE = (G::YTILE > N * G::FTILE) ? G::YTILE : N * G::FTILE;
    static constexpr int DOUBLES = TILE + N * RING_Q + 104;
};

// Shared (per CTA) copy of the model: A interleaved [r][NP], then the DP constants.
template <int N>
struct CtaModel {
    static constexpr int NP = (N + 1) & ~1;
};

// ---------------------------------------------------------------------------
// The first L+1 columns of the trellis in the reference's exact arithmetic
// (src/viterbi.jl:55-88).  This is where the reference's initialisation creates
// exact ties, so the decisions must come from bit-identical scores:
// every value below is ((T1[k,t-1] + lp) + q_t(j)) with separately rounded
// additions and the candidates of a state scanned in list order with a strict >.
//
// The ring structure makes the 60 columns almost fully parallel.  Inside a chain
// every state has one predecessor, so the chain that is at phase s0 at column 0
// simply runs along a diagonal of the trellis until it reaches its tail at column
// L - s0; chains entered at column >= 1 reach their tails after column L, i.e.
// outside the prologue.  Hence
//   A. N*L independent diagonals (one thread each, <= L-1 dependent add pairs) give
//      every tail score of columns 0..L-1;
//   B. per (column, decision state) the best tail candidate in list order -- fully
//      parallel -- and then ONE warp runs the recursion of the N+1 decision states
//      (noise and the chain heads: four dependent operations per column).
// Only the decision states' scores and backpointers are needed afterwards (chunk 0
// converts the head scores into ring entries; the traceback follows T2pro, which is
// static for chain-interior states).  One CTA per channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ring_vit_prologue(VitParams p, int q_in_smem) {
    extern __shared__ __align__(16) double psm[];
    const int ch = blockIdx.x + p.ch0;
    const int ns = p.ns, N = p.RL.N, L = p.RL.L, cols = L + 1, ND = N + 1;
    const char *mb = p.fblob + (size_t)ch * p.fblob_stride;
    const double *sc = (const double *)(mb + p.FL.scal);
    const double c_emit = sc[2], two_s2 = sc[3];
    const double *gm = (const double *)(mb + p.FL.m), *glp = (const double *)(mb + p.FL.in_lp);
    const int *gp = (const int *)(mb + p.FL.in_ptr), *gs = (const int *)(mb + p.FL.in_src);
    const double *y = p.y + (size_t)ch * p.y_stride;
    double *t1 = p.T1pro + (size_t)ch * ns * cols;
    int16_t *t2 = p.T2pro + (size_t)ch * cols * 8;
    double *wc = psm;                       // [ns]   weight of the single in-edge of a chain-interior state
    double *tailv = wc + ns;                // [L][N] tail_i at column t
    double *bT = tailv + (size_t)L * N;     // [cols][ND] best tail candidate of a decision state at column t
    int *aT = (int *)(bT + (size_t)cols * ND);                       // its source state
    double *q = q_in_smem ? (double *)(aT + (((size_t)cols * ND + 1) & ~(size_t)1)) : t1;  // [cols][ns] emissions
    const int tid = threadIdx.x, nth = blockDim.x;
    // in-edges of the decision states (at most N + 1 <= 8 each), staged once: the phases below are short and
    // latency-bound, a chain of dependent global loads per edge would dominate them
    __shared__ double d_lp[8][8];
    __shared__ int d_src[8][8], d_deg[8];
    if (tid < ND * 8) {
        const int d = tid >> 3, k = tid & 7;
        const int sd = d == 0 ? 0 : 1 + (d - 1) * L;
        const int e0 = gp[sd], deg = gp[sd + 1] - e0;
        if (k == 0) d_deg[d] = deg < 8 ? deg : 8;
        if (k < deg) {
            d_src[d][k] = gs[e0 + k];
            d_lp[d][k] = glp[e0 + k];
        }
    }
    // ---- all emissions (the FP64 division is the long-latency part: fully parallel) ----
    {
        const int tpc = nth / ns;  // columns processed concurrently when there is a thread per state
        if (tpc >= 1) {
            const int tq = tid / ns, j = tid - tq * ns;
            if (tq < tpc) {
                const double mj = gm[j];
                for (int t = tq; t < cols; t += tpc) q[(size_t)t * ns + j] = emit_rn(y[t], mj, c_emit, two_s2);
            }
        } else {
            for (int idx = tid; idx < cols * ns; idx += nth) q[idx] = emit_rn(y[idx / ns], gm[idx % ns], c_emit, two_s2);
        }
    }
    for (int j = tid; j < ns; j += nth) wc[j] = glp[gp[j]];
    // (backpointers: a chain-interior state j comes from j - 1, nothing to store; the decision states' are written in B2)
    __syncthreads();
    // ---- A: the chains already running at column 0 ----
    for (int k = tid; k < N * L; k += nth) {
        const int i = k / L, s0 = k - i * L + 1;  // phase at column 0; state index 1 + k
        int j = 1 + k;
        double v = q[j];                           // T1[j, 1] of :55-62
#pragma unroll 4
        for (int t = 1; s0 + t <= L; t++) {
            j++;
            v = __dadd_rn(__dadd_rn(v, wc[j]), q[(size_t)t * ns + j]);
        }
        tailv[(size_t)(L - s0) * N + i] = v;
    }
    __syncthreads();
    // ---- B1: best tail candidate per (column, decision state), list order, strict > ----
    for (int idx = tid; idx < L * ND; idx += nth) {
        const int t = 1 + idx / ND, d = idx % ND;
        double best = -INFINITY;
        int arg = 0;
        for (int e = 0; e < d_deg[d]; e++) {
            const int src = d_src[d][e];
            if (src == 0) continue;  // the noise candidate depends on the recursion: phase B2
            const double v = __dadd_rn(tailv[(size_t)(t - 1) * N + (src - 1) / L], d_lp[d][e]);
            if (v > best) {
                best = v;
                arg = src;
            }
        }
        bT[t * ND + d] = best;
        aT[t * ND + d] = arg;
    }
    __syncthreads();
    // ---- B2: the recursion of the decision states, one warp, one lane per state ----
    if (tid < 32) {
        const int d = tid;
        const bool act = d < ND;
        const int sd = (!act || d == 0) ? 0 : 1 + (d - 1) * L;
        // the noise candidate is the first of every list (lowest source index) and therefore keeps ties
        const double wn = act ? d_lp[d][0] : 0.0;
        double noise = 0.0;  // T1[1,1] = 0, :63
        if (act) t1[sd] = d == 0 ? 0.0 : q[sd];
        for (int t = 1; t <= L; t++) {
            double v = 0.0;
            if (act) {
                const double cn = __dadd_rn(noise, wn), bt = bT[t * ND + d];
                const bool tail = bt > cn;
                v = __dadd_rn(tail ? bt : cn, q[(size_t)t * ns + sd]);
                t1[(size_t)t * ns + sd] = v;
                t2[t * 8 + d] = (int16_t)((tail ? aT[t * ND + d] : 0) + 1);
            }
            noise = __shfl_sync(0xffffffffu, v, 0);
        }
    }
}

// Final state: x[T] = argmax_j T1[j, T] (first maximum, src/viterbi.jl:90), from the last
// chunk's G and P ring plus partial chain sums.  One CTA per channel, one thread per state.
template <int N>
__device__ void final_state(const VitParams &p, int ch) {  // called by all 256 threads of a CTA
    const RingLayout &RL = p.RL;
    const int L = RL.L, NP = RL.NP;
    const double *mdl = p.model + (size_t)ch * RL.total;
    const double *A = mdl + RL.A, *BW = mdl + RL.BW;
    const double *y = p.y + (size_t)ch * p.y_stride;
    const int64_t T = p.T;
    __shared__ double sA[RING_MAX_L * RING_MAX_N], sB[RING_MAX_L * RING_MAX_N], sy[RING_MAX_L];
    for (int k = threadIdx.x; k < N * L; k += blockDim.x) {
        const int i = k / L, r = k % L;
        sA[k] = A[r * NP + i];
        sB[k] = BW[r * NP + i];
    }
    for (int k = threadIdx.x; k < L; k += blockDim.x) sy[k] = y[T - L + k];
    __syncthreads();
    // candidate per state j = 1 + i*L + (sph-1): entered at t0 = T - sph
    double best = -INFINITY;
    int bj = 0x7fffffff;
    for (int idx = threadIdx.x; idx < N * L; idx += blockDim.x) {
        int i = idx / L, sph = idx % L + 1;
        int64_t t0 = T - sph;
        double v = p.Pfin[((size_t)ch * N + i) * RING_Q + (int)(t0 & (RING_Q - 1))];
        for (int r = 0; r < sph; r++) v += fma(sA[i * L + r], sy[L - sph + r], sB[i * L + r]);
        int j = 1 + idx;
        if (v > best || (v == best && j < bj)) {
            best = v;
            bj = j;
        }
    }
    if (threadIdx.x == 0) {  // noise is state 0: wins ties against everything
        const double g = p.Gfin[ch];
        if (g >= best) {
            best = g;
            bj = 0;
        }
    }
    __shared__ double sb[256];
    __shared__ int sj[256];
    sb[threadIdx.x] = best;
    sj[threadIdx.x] = bj;
    __syncthreads();
    for (int k = 128; k >= 1; k >>= 1) {
        if (threadIdx.x < k) {
            double ob = sb[threadIdx.x + k];
            int oj = sj[threadIdx.x + k];
            if (ob > sb[threadIdx.x] || (ob == sb[threadIdx.x] && oj < sj[threadIdx.x])) {
                sb[threadIdx.x] = ob;
                sj[threadIdx.x] = oj;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) p.xend[ch] = (int16_t)sj[0];
}

// ---------------------------------------------------------------------------
// One chunk, one warp.
// ---------------------------------------------------------------------------
// ROLE_BOTH: one warp does the FIR and the recursion of its chunk in turn (repair kernel).
// ROLE_FIR / ROLE_DP: a producer warp runs the FIR one super-window ahead into a double-buffered
// F tile while a consumer warp of the same CTA runs the recursion; they hand the tiles over with
// mbarriers, so the FP64-pipe-bound FIR and the latency-bound recursion overlap.
enum { ROLE_BOTH = 0, ROLE_FIR = 1, ROLE_DP = 2 };

template <int N, int R, int FW = 1>
struct SlotSmem {  // one chunk slot of the warp-specialised kernel, in doubles (FW = FIR producer warps of the slot)
    using G = FirGeom<R>;
    static constexpr int YT = 0;                                   // 2 y tiles per producer
    static constexpr int FT = 2 * FW * G::YTILE;                   // 2 F tiles
    static constexpr int RING = FT + 2 * N * G::FTILE;             // ring + prologue scratch
    static constexpr int BAR = RING + N * RING_Q + 104;            // full[2], empty[2], yready[2], yfree[2] mbarriers
    static constexpr int DOUBLES = BAR + 8;
};

// -DHMM_WS2 selects the dual-producer forward kernel (ring_vit_forward_ws2: four FIR warps and two recursions per SM
// sub-partition, registers re-divided with setmaxnreg).  Measured on B200 it is SLOWER than the paired single-producer
// kernel (N=3: 330 vs 302 us, N=4: 371 vs 326, N=5: 645 vs 513 at 18 M samples), so it is off by default.
#ifndef HMM_WS2
#define HMM_NO_WS2
#endif
#ifndef HMM_LL_NOISE_MODE
#define HMM_LL_NOISE_MODE 1  // 1: path-score piece 1 by a pass over the staged y tile (measured 1 % faster than 0: from the FIR register window)
#endif
// sum over the first min(n, SW) samples of a staged (transposed) y tile of w(e) (y_e - m0)^2, w(e) = w0 - e
template <int R>
__device__ __forceinline__ double ll_noise_tile(const double *ytile, double acc, double m0, double w0, int lane, int n) {
    using G = FirGeom<R>;
    const double *src = ytile + (lane & (R - 1)) * G::YS + (lane >> G::LOGR);
    double w = w0 - (double)lane;
#pragma unroll
    for (int m = 0; m < R; m++) {  // element e = lane + 32 m
        const double dd = src[(32 / R) * m] - m0;
        if (lane + 32 * m < n) acc = fma(w, dd * dd, acc);
        w -= 32.0;
    }
    return acc;
}

// Number of super-windows the forward pass of chunk c covers (0 if there is no such chunk).
__device__ __forceinline__ int chunk_superwindows(const VitParams &p, int c, int SW) {
    if (c >= p.nchunks) return 0;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    if (c == p.nchunks - 1 || e > p.T) e = p.T;
    int64_t base0 = (c == 0 && p.first_prologue) ? 0 : s - p.W;
    if (base0 < 0) base0 = 0;
    return (int)((e - base0 + SW - 1) / SW);
}

// The two FIR producers that share an SM sub-partition meet at a named barrier once per super-window.
__device__ __forceinline__ void pair_sync(int bar_id) {
    asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
}

// NC: neurons whose FIR the CONSUMER warp computes itself (warp-specialised roles only).  A single warp cannot keep
// the FP64 pipe busy (back-to-back DFMAs of one warp issue at half the pipe rate), and a consumer spends a third of its
// time waiting for its producer: with the FIR of a super-window split N - NC : NC, four warps per SM sub-partition feed
// the pipe instead of two, and producer and consumer finish a super-window at about the same time.
// FW: FIR producer warps per slot.  FW = 2: producer `half` (0 / 1) computes the even / odd super-windows into F tile
// `half`, from its own pair of y tiles -- four FIR warps per SM sub-partition keep the FP64 pipe busy where two leave
// it idle a quarter of the time (back-to-back DFMAs of one warp issue at half the pipe rate).
template <int N, int R, int LPC, int ROLE, int NC = 0, int FW = 1, typename S = double>
__device__ void vit_process_chunk(const VitParams &p, const FirCoef<N, LPC, S> &coef, int ch, int c, int kind,
                                  const double *mdl /*smem model*/, double *ws /*per-warp or per-slot smem*/,
                                  int pair_bar = 0 /*named barrier shared with the partner producer, 0 = none*/,
                                  int pair_nsw = 0 /*iterations of the longer chunk of the pair*/, int half = 0) {
    using SS = SlotSmem<N, R, FW>;
    using G = FirGeom<R>;
    constexpr int NP = (N + 1) & ~1;
    const int lane = threadIdx.x & 31;
    const RingLayout &RL = p.RL;
    const int L = RL.L, LP = RL.LP;
    const double NEG = -INFINITY;
    __builtin_assume(__isShared(ws));
    __builtin_assume(__isShared(mdl));
    double *ytile = ws;
    double *fbuf = ws;  // ROLE_BOTH: aliases ytile (see WarpSmem); specialised roles: set per super-window
    double *ring = ws + (ROLE == ROLE_BOTH ? WarpSmem<N, R>::TILE : SS::RING);
    double *zs = ring + N * RING_Q;  // [L+1] <= 97 doubles
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(ws + SS::BAR), *bar_empty = bar_full + 2;
    uint64_t *bar_yready = bar_full + 4, *bar_yfree = bar_full + 6;
    (void)bar_full;
    (void)bar_empty;
    (void)bar_yready;
    (void)bar_yfree;
    const double *A = mdl + RL.A;
    const double *Bc = mdl + RL.Bc, *eG = mdl + RL.eG, *eH = mdl + RL.eH, *eT = mdl + RL.eT;
    const double eTmax = mdl[RL.scal + 5], eHmin = mdl[RL.scal + 6];
    const double *cL = mdl + RL.cL;
    const double *y = p.y + (size_t)ch * p.y_stride;
    const int64_t T = p.T;
    if (ROLE == ROLE_FIR && c >= p.nchunks) {  // no chunk of its own: keep the partner's barrier company
        for (int k = 0; k < pair_nsw; k++) pair_sync(pair_bar);
        return;
    }
    const int64_t s = (int64_t)c * p.Lc;                 // main range [s, e)
    int64_t e = s + p.Lc;
    const bool last = (c == p.nchunks - 1);
    if (last || e > T) e = T;
    int64_t base0;                                       // first super-window
    int64_t tau_first;                                   // first DP step
    double Gprev;
    if (ROLE != ROLE_FIR)
        for (int k = lane; k < N * RING_Q; k += 32) ring[k] = NEG;
    if (kind == START_PROLOGUE) {
        base0 = 0;
        tau_first = L + 1;
        Gprev = 0;  // set after the first FIR
    } else if (kind == START_SPEC) {
        base0 = s - p.W;
        if (base0 < 0) base0 = 0;
        tau_first = base0;
        Gprev = 0.0;
    } else {
        base0 = s;
        tau_first = s;
        const double *eb = p.EB + ((size_t)ch * p.nchunks + (c - 1)) * p.bvec;
        // SB[c] always holds the vector this chunk's CURRENT decisions were started from, so that every later
        // check (a second verification round after a neighbour's boundary arrived, the shard judge) compares
        // what the chunk was really computed from, not a stale speculative vector.
        double *sb = p.SB + ((size_t)ch * p.nchunks + c) * p.bvec;
        Gprev = eb[0];
        if (ROLE != ROLE_FIR) {
            if (lane == 0) sb[0] = Gprev;
            for (int k = lane; k < L; k += 32) {
                int64_t t0 = s - L + k;
                for (int j = 0; j < N; j++) {
                    const double v = eb[1 + j * L + k];
                    ring[j * RING_Q + (int)(t0 & (RING_Q - 1))] = v;
                    sb[1 + j * L + k] = v;
                }
            }
        }
    }
    __syncwarp();
    if (ROLE == ROLE_FIR) {
        // ---- producer: stage tile k+1 while the FIR of tile k runs; hand F tiles to the consumer ----
        const int need = G::SW + (LPC > 0 ? LPC : LP);
        double *yt[2] = {ws + SS::YT + (2 * half) * G::YTILE, ws + SS::YT + (2 * half + 1) * G::YTILE};
        double *ft[2] = {ws + SS::FT, ws + SS::FT + N * G::FTILE};
        const int64_t bstep = (int64_t)FW * G::SW, bfirst = base0 + (int64_t)half * G::SW;
        if (bfirst < e) fir_stage<R>(y, T, bfirst, need, yt[0], lane);
        int k = 0;  // this producer's iteration; it computes super-window FW * k + half
        // path-score piece 1 (see ll_assemble): sum over the chunk's main range of (Tg - g) (y_g - m0)^2, from the y
        // tiles this warp stages anyway -- the recording is not read a second time for ll
        double nacc = 0.0;
        const double m0n = mdl[RL.scal + 3];
        const double wg0 = (double)(p.T_glob - p.t_off - s);  // weight of local step s
#ifdef HMM_PHASE_TIMING
        long long pt_stage = 0, pt_wait = 0, pt_fir = 0;
#endif
        for (int64_t b = bfirst; b < e; b += bstep, k++) {
            const int ybuf = k & 1;                              // own y tile pair
            const int buf = FW == 1 ? (k & 1) : half;            // F tile and its barriers
            const unsigned fpar = FW == 1 ? ((k >> 1) & 1) : (k & 1);
            const bool more = b + bstep < e;
#ifdef HMM_PHASE_TIMING
            const long long q0 = clock64();
#endif
            if (more) {
                // the consumer reads the y tiles too (its share of the FIR): tile buf^1 last held super-window k-1
                if (NC > 0 && k >= 1) mbar_wait(bar_yfree + (buf ^ 1), ((k - 1) >> 1) & 1);
                fir_stage<R>(y, T, b + bstep, need, yt[ybuf ^ 1], lane);
            }
            if (more)
                cp_async_wait_but_one();
            else
                cp_async_wait_all();
            __syncwarp();
            if (NC > 0 && lane == 0) mbar_arrive(bar_yready + buf);  // y tile k has landed

#ifdef HMM_PHASE_TIMING
            const long long q1 = clock64();
#endif
            // Both producers of an SM sub-partition start their FIR together: the warp scheduler otherwise lets one
            // of them run ahead at nearly full rate while the other crawls (measured: 4.5 k vs 8.5 k cycles per
            // super-window), and with one chunk per slot the kernel lasts as long as its slowest slot.
            if (pair_bar) pair_sync(pair_bar);
            mbar_wait(bar_empty + buf, fpar ^ 1);  // the consumer is done with this F tile
#ifdef HMM_PHASE_TIMING
            const long long q2 = clock64();
#endif
            // (own samples of a main-range super-window also feed the path score, from the FIR's register window)
#if HMM_LL_NOISE_MODE == 1
            if (b >= s) nacc = ll_noise_tile<R>(yt[ybuf], nacc, m0n, wg0 - (double)(b - s), lane, (int)(e - b > G::SW ? G::SW : e - b));
            double *llp = nullptr;
#elif HMM_LL_NOISE_MODE == 2
            double *llp = nullptr;  // (timing experiments only: ll is wrong)
#else
            double *llp = b >= s ? &nacc : nullptr;
#endif
            const double w0 = wg0 - (double)(b - s);
            const int64_t left = e - b;
            const int nval = left > G::SW ? G::SW : (int)left;
            if constexpr (LPC > 0)
                fir_compute_c<N, R, LPC, 0, N - NC, S>(coef, Bc, yt[ybuf], ft[buf], lane, llp, m0n, w0, nval);
            else
                fir_compute<N, R, 0, N - NC, S>(A, Bc, LP, yt[ybuf], ft[buf], lane, llp, m0n, w0, nval);
            if (lane == 0) mbar_arrive(bar_full + buf);   // fir_compute ends with __syncwarp()
#ifdef HMM_PHASE_TIMING
            const long long q3 = clock64();
            pt_stage += q1 - q0;
            pt_wait += q2 - q1;
            pt_fir += q3 - q2;
#endif
        }
#ifdef HMM_PHASE_TIMING
        if (lane == 0 && kind == START_SPEC && (c % 293) == 1)
            printf("producer %d: %d super-windows, staging %lld, wait-for-empty %lld, FIR %lld cyc/sw\n", c, k,
                   pt_stage / (k ? k : 1), pt_wait / (k ? k : 1), pt_fir / (k ? k : 1));
#endif
        {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) nacc += __shfl_xor_sync(0xffffffffu, nacc, d);
            if (lane == 0) p.ll_noise[((size_t)ch * p.nchunks + c) * 2 + half] = nacc;
        }
        if (pair_bar)
            for (; k < pair_nsw; k++) pair_sync(pair_bar);  // the partner's chunk is longer
        return;
    }
    // decision words / mask words of this chunk's first super-window: indexed with 32-bit relative steps below
    uint32_t *decb = p.dec + (size_t)ch * T + base0;
    uint32_t *nzb = p.nzmask + (size_t)ch * ((T + 31) / 32) + (base0 >> 5);
    const int Wd = L < 32 ? L : 32;
    const int nsub = (32 + Wd - 1) / Wd;
    const int mysub = lane / Wd;
    // window-invariant constants in registers (the compiler cannot hoist them past the ring stores)
    double eHr[N], cLr[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        eHr[i] = eH[i];
        cLr[i] = cL[i];
    }
    // "live" masks of the last four 32-step windows (bit = a chain score created at that step
    // may still win a decision when it arrives L steps later); see the quiet-window fast path
    uint32_t lq1, lq2, lq3, lq4;
    lq1 = lq2 = lq3 = lq4 = (kind == START_SPEC) ? 0u : 0xffffffffu;
    double Gq = NAN, ghq[N], thq[N];  // quiet-window constants, valid for G == Gq
#pragma unroll
    for (int i = 0; i < N; i++) ghq[i] = thq[i] = 0.0;
    // (compile-time for the constant-coefficient variants, where L = LPC: the selects below fold away)
    const int Lq = LPC > 0 ? LPC : L;
    const int qq = (Lq >> 5) + ((Lq & 31) ? 1 : 0);  // windows back to the first word holding step t-L
    const int lsh = (32 - (Lq & 31)) & 31;
    const int tf_rel = (int)(tau_first - base0);  // steps are tracked relative to base0 in 32 bits
    const int e_rel = (int)(e - base0);
    const int s_rel = (int)(s - base0);

#ifdef HMM_PHASE_TIMING
    long long tm_fir = 0, tm_dp = 0, tm_n = 0;
#endif
    int swk = 0;
    for (int64_t b = base0; b < e; b += G::SW, swk++) {
#ifdef HMM_PHASE_TIMING
        const long long tm0 = clock64();
#endif
        if (ROLE == ROLE_DP) {
            const int buf = swk & 1;
            fbuf = ws + SS::FT + buf * N * G::FTILE;
            if constexpr (NC > 0) {
                // ---- consumer's share of the FIR: the last NC neurons, from the y tile the producer staged ----
                const double *ytk = ws + SS::YT + buf * G::YTILE;
                mbar_wait(bar_yready + buf, (swk >> 1) & 1);
                if constexpr (LPC > 0)
                    fir_compute_c<N, R, LPC, N - NC, N, S>(coef, Bc, ytk, fbuf, lane);
                else
                    fir_compute<N, R, N - NC, N, S>(A, Bc, LP, ytk, fbuf, lane);
                if (lane == 0) mbar_arrive(bar_yfree + buf);
            }
            // ---- consumer: wait for the producer's F planes of this super-window ----
            mbar_wait(bar_full + buf, (swk >> 1) & 1);
        } else {
            // ---- stage y and run the FIR: F_i(b + t) for the whole super-window ----
            if constexpr (LPC > 0)
                fir_superwindow_c<N, R, LPC, S>(y, T, b, coef, Bc, ytile, fbuf, lane);
            else
                fir_superwindow<N, R, S>(y, T, b, A, Bc, LP, ytile, fbuf, lane);
        }
#ifdef HMM_PHASE_TIMING
        const long long tm1 = clock64();
#endif
        // ---- chunk 0: convert the faithful prologue (columns 0..L) into ring state ----
        if (kind == START_PROLOGUE && b == 0) {
            const double *sc = mdl + RL.scal;
            const double w_nn = sc[0], c_emit = sc[1], two_s2 = sc[2], m0 = sc[3];
            const double *t1 = p.T1pro + (size_t)ch * p.ns * (L + 1);
            if (lane == 0) {
                double z = 0.0;
                zs[0] = 0.0;
                for (int t = 1; t <= L; t++) {
                    double dd = y[t] - m0;
                    z += w_nn + (c_emit - (dd * dd) / two_s2);
                    zs[t] = z;
                }
            }
            __syncwarp();
            const double *BW = p.model + (size_t)ch * RL.total + RL.BW;  // cold part: global memory
            for (int t0 = 1 + lane; t0 <= L; t0 += 32) {
                const double yv = y[t0];
#pragma unroll
                for (int i = 0; i < N; i++) {
                    double uh = t1[(size_t)t0 * p.ns + 1 + i * L] - zs[t0];          // U(head_i) at t0
                    double pu = uh - fma(A[i], yv, BW[i]);                             // minus d_t0(head_i)
                    double f = fbuf[i * G::FTILE + (t0 & (R - 1)) * G::FS + (t0 >> G::LOGR)];
                    ring[i * RING_Q + (t0 & (RING_Q - 1))] = pu + f;
                }
            }
            Gprev = t1[(size_t)L * p.ns] - zs[L];
            __syncwarp();
        }
        // ---- boundary dump: speculative start vector at s ----
        if (kind == START_SPEC && b == s) {
            double *sb = p.SB + ((size_t)ch * p.nchunks + c) * p.bvec;
            if (lane == 0) sb[0] = Gprev;
            for (int k = lane; k < L; k += 32) {
                int64_t t0 = s - L + k;
                for (int j = 0; j < N; j++) sb[1 + j * L + k] = ring[j * RING_Q + (int)(t0 & (RING_Q - 1))];
            }
        }
        // ---- max-plus recursion over the super-window, 32 steps per window ----
        const int b_rel = (int)(b - base0);
        // Per-super-window invariants, hoisted by hand: the quiet-window path below is the consumer's inner loop
        // (80 % of the windows) and the compiler otherwise rebuilds every 64-bit address from the kernel
        // parameters in each window (110 instructions per quiet window, most of them address arithmetic).
        const bool regular = Lq >= 32 && b_rel >= tf_rel && b_rel + G::SW <= e_rel;  // 8 full windows of recursion steps
        const double *fp0 = fbuf + (lane & (R - 1)) * G::FS + (lane >> G::LOGR);       // F of (window 0, this lane's step)
        double *rp0 = ring + lane;                                                      // ring slot of that step
        uint32_t *dq = decb + b_rel + lane;
        uint32_t *nq = nzb + (b_rel >> 5);
        double *pfin = p.Pfin + (size_t)ch * N * RING_Q + lane;
        for (int wdw = 0; wdw < R; wdw++) {
            const int t0_rel = b_rel + 32 * wdw;
            if (!regular && t0_rel + 32 <= tf_rel) {  // before the first recursion step: pre-loaded entries only
                lq4 = lq3; lq3 = lq2; lq2 = lq1;
                lq1 = (kind == START_SPEC) ? 0u : 0xffffffffu;
                if (t0_rel >= s_rel) {  // (prologue columns: the traceback takes them from T2pro)
                    decb[t0_rel + lane] = 0u;
                    if (lane == 0) nzb[t0_rel >> 5] = 0u;
                }
                continue;
            }
            if (!regular && t0_rel >= e_rel) break;
            const int t_rel = t0_rel + lane;
            // ring slots: base0 and the super-window are multiples of RING_Q, so the slot of step (window, lane) is
            // (32 window mod RING_Q) + lane
            const int slot0 = (32 * wdw) & (RING_Q - 1);
            const int slot_w = slot0 + lane;
            double Fv[N];
#pragma unroll
            for (int i = 0; i < N; i++) Fv[i] = fp0[i * G::FTILE + (32 / R) * wdw];
            // Quiet-window fast path: if none of the pending chain scores that arrive in this
            // window was live when it was created, no tail can win any decision here: G stays,
            // every head is entered from noise and all backpointers are 0 -- the tails are not
            // even read.
            if (regular || (L >= 32 && t0_rel >= tf_rel && t0_rel + 32 <= e_rel)) {
                const uint32_t lo = qq == 1 ? lq1 : qq == 2 ? lq2 : qq == 3 ? lq3 : lq4;
                const uint32_t hi = qq == 1 ? 0u : qq == 2 ? lq1 : qq == 3 ? lq2 : lq3;
                if (__funnelshift_r(lo, hi, lsh) == 0u) {
                    if (Gq != Gprev) {  // G moves only where a chain ends: G + eH and the liveness thresholds
                        Gq = Gprev;     // are recomputed then, not once per window
                        const double marg = 1e-13 * fabs(Gprev);
#pragma unroll
                        for (int i = 0; i < N; i++) {
                            ghq[i] = Gprev + eHr[i];
                            // live <=> q + cL + marg > G with q = ghq + F, i.e. F > (G - cL - marg) - ghq: tested on F
                            // itself, so that the vote does not wait for the addition below (FP64 latency is long
                            // and this chain is the consumer's critical path); the margins dwarf the rounding
                            thq[i] = ((Gprev - cLr[i]) - marg) - ghq[i];
                        }
                    }
                    bool live = false;
#pragma unroll
                    for (int i = 0; i < N; i++) live = live | (Fv[i] > thq[i]);
#pragma unroll
                    for (int i = 0; i < N; i++) rp0[i * RING_Q + slot0] = ghq[i] + Fv[i];
                    if (last) {
#pragma unroll
                        for (int i = 0; i < N; i++) pfin[i * RING_Q + slot0] = ghq[i];
                    }
                    lq4 = lq3; lq3 = lq2; lq2 = lq1;
                    lq1 = __ballot_sync(0xffffffffu, live);
                    if (t0_rel >= s_rel) {  // all-noise decisions: one coalesced 128-byte store
                        dq[32 * wdw] = 0u;
                        if (lane == 0) nq[wdw] = 0u;
                    }
                    __syncwarp();
                    continue;
                }
            }
            const bool in_range = t_rel >= tf_rel && t_rel < e_rel;
            const int slot_r = (t_rel - L) & (RING_Q - 1);
            unsigned nz = 0;
            uint32_t myword = 0;
            unsigned lvw = 0;
            for (int sub = 0; sub < nsub; sub++) {
                const bool active = in_range && (mysub == sub);
                double tails[N];
#pragma unroll
                for (int j = 0; j < N; j++) tails[j] = active ? ring[j * RING_Q + slot_r] : NEG;
                double X = NEG;
#pragma unroll
                for (int j = 0; j < N; j++) {
                    double v = tails[j] + eG[j];
                    X = (v > X) ? v : X;
                }
                // Fast path 1: the noise score can only change inside this window if some
                // X_t exceeds the window-start G (G is non-decreasing) -- true only where a
                // chain has just ended.  Otherwise G is constant and no scan is needed.
                double Gincl = Gprev, Gexcl = Gprev;
                uint32_t word = 0;
                const unsigned any_end = __ballot_sync(0xffffffffu, active && X > Gprev);
                if (any_end) {
                    int jx = 0;
                    double Xi = NEG;
#pragma unroll
                    for (int j = 0; j < N; j++) {
                        double v = tails[j] + eG[j];
                        if (v > Xi) {  // strict: first maximum in candidate order
                            Xi = v;
                            jx = j + 1;
                        }
                    }
                    double M = X;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        double o = shfl_up_d(M, d);
                        if (lane >= d && o > M) M = o;
                    }
                    Gincl = (M > Gprev) ? M : Gprev;
                    Gexcl = shfl_up_d(Gincl, 1);
                    if (lane == 0) Gexcl = Gprev;
                    word = (X > Gexcl) ? (uint32_t)jx : 0u;  // noise (first candidate) keeps ties
                }
                // Fast path 2: a tail -> head candidate can win only if
                // max_j tail_j + max eT > G + min eH (rounding is monotone, so this bound is exact).
                double tmax = tails[0];
#pragma unroll
                for (int j = 1; j < N; j++) tmax = (tails[j] > tmax) ? tails[j] : tmax;
                const unsigned any_th = __ballot_sync(0xffffffffu, active && (tmax + eTmax > Gexcl + eHmin));
                double Pv[N];
                if (any_th) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        double best = Gexcl + eH[i];
                        int k = 0;
#pragma unroll
                        for (int j = 0; j < N; j++) {
                            if (j == i) continue;
                            double v = tails[j] + eT[j * NP + i];
                            if (v > best) {
                                best = v;
                                k = j + 1;
                            }
                        }
                        word |= (uint32_t)k << (4 * (i + 1));
                        Pv[i] = best;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < N; i++) Pv[i] = Gexcl + eH[i];
                }
                bool live = false;
                if (active) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        const double q = Pv[i] + Fv[i];
                        ring[i * RING_Q + slot_w] = q;
                        if (last) p.Pfin[((size_t)ch * N + i) * RING_Q + slot_w] = Pv[i];
                        live = live || (q + cLr[i] + 1e-13 * fabs(Gexcl) > Gexcl);
                    }
                }
                // entries pre-loaded before the first recursion step (prologue) stay live
                if (sub == 0 && kind != START_SPEC && t_rel < tf_rel) live = true;
                lvw |= __ballot_sync(0xffffffffu, live);
                if (active) myword = word;
                if (any_end) {
                    nz |= __ballot_sync(0xffffffffu, active && (word & 15u) != 0);
                    Gprev = shfl_d(Gincl, 31);
                }
                __syncwarp();
            }
            lq4 = lq3; lq3 = lq2; lq2 = lq1;
            lq1 = lvw;
            if (t0_rel >= s_rel) {  // main range only (warm-up decisions belong to the previous chunk)
                if (t_rel < e_rel) decb[t_rel] = myword;
                if (lane == 0) nzb[t0_rel >> 5] = nz;
            }
        }
        if (ROLE == ROLE_DP) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + (swk & 1));
        }
#ifdef HMM_PHASE_TIMING
        {
            const long long tm2 = clock64();
            tm_fir += tm1 - tm0;
            tm_dp += tm2 - tm1;
            tm_n++;
        }
#endif
    }
#ifdef HMM_PHASE_TIMING
    if (lane == 0 && kind == START_SPEC && (c % 293) == 1)
        printf("chunk %d: %lld super-windows, FIR+staging %lld cyc/sw, recursion %lld cyc/sw\n", c, tm_n,
               tm_fir / (tm_n ? tm_n : 1), tm_dp / (tm_n ? tm_n : 1));
#endif
    // ---- end of chunk: true boundary vector for the next chunk, or final state ----
    if (!last) {
        double *eb = p.EB + ((size_t)ch * p.nchunks + c) * p.bvec;
        if (lane == 0) eb[0] = Gprev;
        for (int k = lane; k < L; k += 32) {
            int64_t t0 = e - L + k;
            for (int j = 0; j < N; j++) eb[1 + j * L + k] = ring[j * RING_Q + (int)(t0 & (RING_Q - 1))];
        }
    } else if (lane == 0) {
        p.Gfin[ch] = Gprev;
    }
    __syncwarp();
}

template <int N, int R>
__device__ void load_model_smem(const VitParams &p, int ch, double *mdl) {
    const double *g = p.model + (size_t)ch * p.RL.total;
    for (int k = threadIdx.x; k < p.RL.hot; k += blockDim.x) mdl[k] = g[k];
    __syncthreads();
}

// Warp-specialised forward kernel: SLOTS chunk slots per CTA, each a {FIR producer, recursion consumer} warp pair.
// SLOTS = 8: one CTA of 16 warps per SM -- the producers of slots s and s + 4 run on the same SM sub-partition and
// advance in step (pair_sync); SLOTS = 4: two CTAs of 8 warps per SM (models whose slots are too large for 8).
template <int N>
struct ConsumerFirShare {  // neurons whose FIR the consumer warp computes (see vit_process_chunk)
#ifdef HMM_CONSUMER_FIR
    static constexpr int value = HMM_CONSUMER_FIR < N ? HMM_CONSUMER_FIR : N - 1;
#else
    static constexpr int value = 0;  // measured at N = 3, 4, 5: the consumer is co-critical, any share slows the kernel down
#endif
};

template <int N, int R, int LPC, int SLOTS, typename S>
__global__ void __launch_bounds__(SLOTS * 64, SLOTS == 8 ? 1 : 2)
    ring_vit_forward_ws(const __grid_constant__ VitParams p, const __grid_constant__ FirCoef<N, LPC, S> coef) {
    extern __shared__ __align__(16) double smem_d[];
    using G = FirGeom<R>;
    const int ch = blockIdx.y + p.ch0;
    double *mdl = smem_d;
    const int warp = warp_index_uniform(), slot = warp & (SLOTS - 1);
    double *ws = smem_d + ((p.RL.hot + 1) & ~1) + (size_t)slot * SlotSmem<N, R>::DOUBLES;
    if ((threadIdx.x & 31) == 0 && warp < SLOTS) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(ws + SlotSmem<N, R>::BAR);
        for (int k = 0; k < 8; k++) mbar_init(bars + k, 1);
    }
    load_model_smem<N, R>(p, ch, mdl);  // ends with __syncthreads(): barriers initialised, model staged
    const int c = blockIdx.x * SLOTS + slot;
    const int kind = (c == 0 && p.first_prologue) ? START_PROLOGUE : START_SPEC;
    if (warp < SLOTS) {
        int pair_bar = 0, pair_nsw = 0;
        if (SLOTS == 8) {
            pair_bar = 1 + (slot & 3);
            const int a = chunk_superwindows(p, c, G::SW), b = chunk_superwindows(p, blockIdx.x * SLOTS + (slot ^ 4), G::SW);
            pair_nsw = a > b ? a : b;
        }
        vit_process_chunk<N, R, LPC, ROLE_FIR, ConsumerFirShare<N>::value, 1, S>(p, coef, ch, c, kind, mdl, ws, pair_bar, pair_nsw);
    } else if (c < p.nchunks)
        vit_process_chunk<N, R, LPC, ROLE_DP, ConsumerFirShare<N>::value, 1, S>(p, coef, ch, c, kind, mdl, ws);
}

// Dual-producer variant (N <= 5, R = 4): ONE CTA of 24 warps per SM -- 8 chunk slots x {2 FIR producers, 1 recursion
// consumer}.  Warps 0..15 are producers (slot = w & 7, half = w >> 3), warps 16..23 consumers; the slots s and s + 4 live
// on the same SM sub-partition, so each sub-partition runs four FIR warps and two recursions.  Registers are
// re-divided between the roles after launch (setmaxnreg, per warpgroup): the FIR needs ~60, the recursion ~125,
// and 24 x 32 x 80 is all a CTA can be launched with.
// setmaxnreg draws from the CTA's own pool: what the 16 producer warps give back ((80 - P) x 512) must cover what the
// 8 consumer warps ask for ((C - 80) x 256), i.e. C <= 240 - 2 P -- otherwise the consumers wait forever.
template <int N>
struct Ws2Regs {
    static constexpr int producer = N <= 4 ? 56 : 64;  // FIR: 4 N accumulators x 2 + window + addressing
    static constexpr int consumer = N <= 4 ? 128 : 112;
    static_assert(consumer <= 240 - 2 * producer, "consumers would starve");
};
template <int N, int LPC, typename S>
__global__ void __launch_bounds__(768, 1)
    ring_vit_forward_ws2(const __grid_constant__ VitParams p, const __grid_constant__ FirCoef<N, LPC, S> coef) {
    extern __shared__ __align__(16) double smem_d[];
    constexpr int R = 4, SLOTS = 8;
    using G = FirGeom<R>;
    using SS = SlotSmem<N, R, 2>;
    const int ch = blockIdx.y + p.ch0;
    double *mdl = smem_d;
    const int warp = warp_index_uniform();
    const bool producer = warp < 2 * SLOTS;
    const int slot = producer ? (warp & (SLOTS - 1)) : warp - 2 * SLOTS;
    double *ws = smem_d + ((p.RL.hot + 1) & ~1) + (size_t)slot * SS::DOUBLES;
    if ((threadIdx.x & 31) == 0 && warp < SLOTS) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(ws + SS::BAR);
        for (int k = 0; k < 8; k++) mbar_init(bars + k, 1);
    }
    load_model_smem<N, R>(p, ch, mdl);  // ends with __syncthreads(): barriers initialised, model staged
    const int c = blockIdx.x * SLOTS + slot;
    const int kind = (c == 0 && p.first_prologue) ? START_PROLOGUE : START_SPEC;
    if (producer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Ws2Regs<N>::producer));
        const int half = warp >> 3;
        const int a = chunk_superwindows(p, c, G::SW), b = chunk_superwindows(p, blockIdx.x * SLOTS + (slot ^ 4), G::SW);
        const int ia = a > half ? (a - half + 1) / 2 : 0, ib = b > half ? (b - half + 1) / 2 : 0;
        vit_process_chunk<N, R, LPC, ROLE_FIR, 0, 2, S>(p, coef, ch, c, kind, mdl, ws, 1 + (slot & 3) + 4 * half,
                                                     ia > ib ? ia : ib, half);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Ws2Regs<N>::consumer));
        if (c < p.nchunks) vit_process_chunk<N, R, LPC, ROLE_DP, 0, 2, S>(p, coef, ch, c, kind, mdl, ws);
    }
}

// Boundary check: speculative start vector of chunk c vs true end vector of c-1
// must agree up to an additive constant.
__device__ __forceinline__ bool boundary_matches(const double *sb, const double *eb, int n, int lane) {
    const double s0 = sb[0], e0 = eb[0];
    bool bad = false;
    for (int k = lane; k < n; k += 32) {
        double a = sb[k], b = eb[k];
        bool ia = isinf(a), ib = isinf(b);
        if (ia || ib) {
            if (ia != ib) bad = true;
            continue;
        }
        // Accept only what rounding explains: both vectors carry the same path scores up to a constant, and the
        // part of a score that is not common with its own G went through a handful of additions at the
        // magnitude of the normalised scores, so the two differences agree to a few ulp of that magnitude
        // (45 ulp allowed).  A chunk accepted here made every decision from scores within `tol` of the true
        // ones, i.e. its decisions are the true ones unless a margin is below 2 tol (~1e-10 at |score| 1e4) --
        // far below the 1e-9 |score| the input screen guarantees.
        double da = a - s0, db = b - e0;
        double mag = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(s0), fabs(e0)));
        double tol = 1e-13 + 1e-14 * mag;
        if (!(fabs(da - db) <= tol)) bad = true;
    }
    return !__any_sync(0xffffffffu, bad);
}

// Time-sharded decode, one-collective protocol: every rank checks every shard boundary of the all-gathered
// summaries (layout: api.cu shard_summary_kernel).  One warp per boundary r | r+1, then a fixed-order sum of ll.
__global__ void vshard_judge_kernel(const double *g, int n, int bvec, double *out) {
    __shared__ int bad[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int len = 2 * bvec + 4;
    if (lane == 0) bad[w] = 0;
    __syncwarp();
    for (int r = w; r < n - 1; r += (blockDim.x >> 5)) {
        const double *left = g + (size_t)r * len, *right = g + (size_t)(r + 1) * len;
        const bool okf = boundary_matches(right + bvec, left, bvec, lane);          // start vector of r+1 vs end vector of r
        const bool okt = left[2 * bvec + 1] == right[2 * bvec + 0];                  // state assumed by r vs state of r+1
        if (lane == 0 && !(okf && okt)) bad[w]++;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ll = 0.0;
        int nb = 0;
        for (int r = 0; r < n; r++) ll += g[(size_t)r * len + 2 * bvec + 2];
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) nb += bad[k];
        out[0] = ll;
        out[1] = (double)nb;
    }
}

// ---- peer-memory exchange (time-sharded decode, one process or one rank per GPU) --------------------------
// Every rank owns an "exchange block": [2][world] arrival flags (u64) followed by [2][world][slen] summaries,
// double-buffered by decode parity.  After its local decode a rank stores its summary straight into EVERY
// peer's block over NVLink (peer-mapped or CUDA-IPC pointers) and then raises its flag there with a system-
// scope release; the judge kernel of each rank spins on the `world` flags of its own block, checks every shard
// boundary and leaves [total ll, inconsistent boundaries] in mapped pinned host memory.  No NCCL call, no
// host round trip between the decode and the verdict.  A rank cannot be two decodes ahead of a peer (its
// judge needs that peer's summary of the current decode), so two buffers suffice.
__device__ __forceinline__ size_t xchg_data_off(int world) { return ((size_t)2 * world * sizeof(unsigned long long) + 255) & ~size_t(255); }

__global__ void __launch_bounds__(256)
    vshard_exchange_kernel(const double *eb_last, const double *sb_first, const long long *own_first,
                           const long long *own_ghost, const double *ll, long long shift, int bvec,
                           char *const *peers /*[world] exchange blocks*/, int rank, int world,
                           const unsigned long long *epoch) {
    const int slen = 2 * bvec + 4;
    const unsigned long long e = *epoch;
    const int par = (int)(e & 1);
    auto glob = [&](const long long *q) {
        if (!q) return -2.0;
        const long long v = *q;
        return (double)(v >= 0 ? v + shift : v);  // < 2^53: exact
    };
    for (int q = 0; q < world; q++) {
        double *dst = reinterpret_cast<double *>(peers[q] + xchg_data_off(world)) + ((size_t)par * world + rank) * slen;
        for (int k = threadIdx.x; k < bvec; k += blockDim.x) {
            dst[k] = eb_last ? eb_last[k] : 0.0;
            dst[bvec + k] = sb_first ? sb_first[k] : 0.0;
        }
        if (threadIdx.x == 0) {
            dst[2 * bvec + 0] = glob(own_first);
            dst[2 * bvec + 1] = glob(own_ghost);
            dst[2 * bvec + 2] = ll[0];
            dst[2 * bvec + 3] = 0.0;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(peers[threadIdx.x]) + (size_t)par * world + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(flag), "l"(e + 1) : "memory");
    }
}

__global__ void __launch_bounds__(256)
    vshard_judge_p2p_kernel(char *own_block, int world, int bvec, unsigned long long *epoch, double *out /*mapped host*/,
                            long long timeout_ns) {
    __shared__ int bad[8], s_timeout;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long e = *epoch;
    const int par = (int)(e & 1);
    const int slen = 2 * bvec + 4;
    if (threadIdx.x == 0) s_timeout = 0;
    if (lane == 0) bad[w] = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(own_block) + (size_t)par * world + threadIdx.x;
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(flag) : "memory");
            if (v == e + 1) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if ((long long)(t1 - t0) > timeout_ns) {  // a peer never arrived (crashed rank, missing launch): report
                s_timeout = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const double *g = reinterpret_cast<const double *>(own_block + xchg_data_off(world)) + (size_t)par * world * slen;
    if (!s_timeout)
        for (int r = w; r < world - 1; r += (blockDim.x >> 5)) {
            const double *left = g + (size_t)r * slen, *right = g + (size_t)(r + 1) * slen;
            const bool okf = boundary_matches(right + bvec, left, bvec, lane);
            const bool okt = left[2 * bvec + 1] == right[2 * bvec + 0];
            if (lane == 0 && !(okf && okt)) bad[w]++;
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ll = 0.0;
        int nb = 0;
        for (int r = 0; r < world; r++) ll += g[(size_t)r * slen + 2 * bvec + 2];
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) nb += bad[k];
        out[0] = ll;
        out[1] = s_timeout ? -1.0 : (double)nb;
        *epoch = e + 1;
    }
}

void vshard_exchange_run(const double *eb_last, const double *sb_first, const long long *own_first,
                         const long long *own_ghost, const double *ll, long long shift, int bvec, char *const *peers_dev,
                         int rank, int world, const unsigned long long *epoch_dev, cudaStream_t st) {
    vshard_exchange_kernel<<<1, 256, 0, st>>>(eb_last, sb_first, own_first, own_ghost, ll, shift, bvec, peers_dev, rank,
                                              world, epoch_dev);
    HMM_CUDA(cudaGetLastError());
}
void vshard_judge_p2p_run(char *own_block, int world, int bvec, unsigned long long *epoch_dev, double *out_mapped,
                          cudaStream_t st) {
    vshard_judge_p2p_kernel<<<1, 256, 0, st>>>(own_block, world, bvec, epoch_dev, out_mapped, 2000000000LL);
    HMM_CUDA(cudaGetLastError());
}
size_t vshard_exchange_block_bytes(int world, int bvec) {
    return (((size_t)2 * world * sizeof(unsigned long long) + 255) & ~size_t(255)) + sizeof(double) * 2 * (size_t)world * (2 * bvec + 4);
}

void vshard_judge_run(const double *gathered_dev, int n_ranks, int bvec, double *out_dev, cudaStream_t st) {
    vshard_judge_kernel<<<1, 256, 0, st>>>(gathered_dev, n_ranks, bvec, out_dev);
    HMM_CUDA(cudaGetLastError());
}

// "Last CTA continues": every CTA of a (channel) row arrives at a counter; the one that arrives last sees all the
// others' global writes (fence + atomic) and carries on alone.  The counter is re-armed for the next run.
__device__ __forceinline__ bool last_cta_of_row(unsigned *cnt, int *s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(cnt, 1u);
        *s_flag = (prev == gridDim.x - 1);
        if (*s_flag) *cnt = 0u;
    }
    __syncthreads();
    if (*s_flag) __threadfence();
    return *s_flag != 0;
}

// Same, and the CTAs also tell the last one whether ANY of them saw a failed check (bit 16 and up of the counter
// count the CTAs that did), so that it does not have to read the flags back.  gridDim.x < 65536.
__device__ __forceinline__ bool last_cta_of_row_any(unsigned *cnt, int *s_flag, bool bad, int *any_out) {
    const int bad_cta = __syncthreads_or(bad ? 1 : 0);
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(cnt, 1u + (bad_cta ? 0x10000u : 0u));
        const int last = (prev & 0xffffu) == gridDim.x - 1;
        s_flag[0] = last;
        s_flag[1] = ((prev >> 16) != 0u || bad_cta) ? 1 : 0;
        if (last) *cnt = 0u;
    }
    __syncthreads();
    if (s_flag[0]) __threadfence();
    *any_out = s_flag[1];
    return s_flag[0] != 0;
}

// Forward verification in ONE launch: (1) every warp checks one chunk boundary -- the speculative start vector of
// chunk c against the true end vector of chunk c-1; (2) the last CTA to finish re-runs the flagged chunks from the
// true vectors, sequentially (a re-run changes EB[c], so chunk c+1 is re-checked against it), and (3) computes
// the final state x[T] = argmax_j T1[j,T] when the sequence really ends in this plan.
template <int N, int R, typename S>
__global__ void __launch_bounds__(256) ring_vit_verify_fwd(VitParams p) {
    extern __shared__ __align__(16) double smem_d[];
    __shared__ int s_flag[2];
    const int ch = blockIdx.y + p.ch0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *flag = p.fwd_flag + (size_t)ch * p.nchunks;
    bool bad = false;
    {
        const int c = blockIdx.x * 8 + warp;
        if (c >= 1 && c < p.nchunks) {
            const double *sb = p.SB + ((size_t)ch * p.nchunks + c) * p.bvec;
            const double *eb = p.EB + ((size_t)ch * p.nchunks + c - 1) * p.bvec;
            bool ok = boundary_matches(sb, eb, p.bvec, lane);
            if (p.dbg_flag_every > 0 && c % p.dbg_flag_every == 0) ok = false;  // HMMCUDA_DEBUG_FLAG_EVERY: force the repair path
            if (lane == 0) flag[c] = ok ? 0 : 1;
            bad = !ok;
        }
    }
    int any = 0;
    if (!last_cta_of_row_any(p.sync_cnt + ch * 4 + 0, s_flag, bad, &any)) return;
    int repaired = 0;
    if (any) {
        double *mdl = smem_d;
        load_model_smem<N, R>(p, ch, mdl);
        if (warp == 0) {
            double *ws = smem_d + ((p.RL.hot + 1) & ~1);
            bool prev_rerun = false;
            for (int c = 1; c < p.nchunks; c++) {
                bool need = __ldcg(flag + c) != 0;
                if (!need && prev_rerun) {
                    const double *sb = p.SB + ((size_t)ch * p.nchunks + c) * p.bvec;
                    const double *eb = p.EB + ((size_t)ch * p.nchunks + c - 1) * p.bvec;
                    need = !boundary_matches(sb, eb, p.bvec, lane);
                }
                if (need) {
                    vit_process_chunk<N, R, 0, ROLE_BOTH, 0, 1, S>(p, FirCoef<N, 0, S>{}, ch, c, START_EXACT, mdl, ws);
                    __threadfence();
                    repaired++;
                }
                prev_rerun = need;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // the counter always describes the LAST verification
        p.counters[ch * 4 + 0] = repaired;
        if (p.res_host) p.res_host[ch * 4 + 1] = (double)repaired;
    }
    if (p.last_true_end) final_state<N>(p, ch);
}

// ---------------------------------------------------------------------------
// Traceback.  State encoding: -1 = noise, else t0 * 8 + neuron (chain entered at t0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long enc_spike(int64_t t0, int i) { return (long long)t0 * 8 + i; }

// Per-warp shared workspace of the traceback: one tile of 8192 steps.
constexpr int TR_TILE_W = 256;  // mask words per tile
constexpr int TR_CAP = 512;     // pre-fetched decision entries per tile
constexpr int TR_WARP_U32 = TR_TILE_W /*mw*/ + TR_TILE_W /*pre*/ + 2 * TR_CAP /*ent*/ + TR_CAP / 2 /*pos, u16*/;

template <int N>
__device__ void trace_chunk(const VitParams &p, int ch, int c, int64_t tau_hi, long long st, bool record_look,
                            uint32_t *tws /*per-warp, TR_WARP_U32 words*/) {
    const int lane = threadIdx.x & 31;
    const int L = p.RL.L;
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc_t;
    int64_t e = s + p.Lc_t;
    if (c == p.nchunks_t - 1 || e > T) e = T;
    const int64_t lo = (c == 0 && p.first_prologue) ? (int64_t)(L + 1) : s;
    // Everything below is in 32-bit coordinates relative to the chunk start s (a multiple of 256):
    // the walk covers [lo, tau_hi], at most Lc_t + W steps, and a chain entered before lo lies at most L below 0.
    const uint32_t *dec_s = p.dec + (size_t)ch * T + s;
    const uint32_t *nzm_s = p.nzmask + (size_t)ch * ((T + 31) / 32) + (s >> 5);
    int16_t *x = p.x + (size_t)ch * p.x_stride;
    int16_t *xs = x + s;
    const int e_r = (int)(e - s), lo_r = (int)(lo - s);
    const int dmin_r = s >= L ? -(1 << 30) : (int)(L - s);  // steps below it have no decision word L steps earlier
    int xlo_r, xhi_r;                                        // the part of [s, e) this plan writes x for
    {
        int64_t a = p.x_lo - s, b = p.x_hi - s;
        xlo_r = (int)(a < 0 ? 0 : (a > e_r ? e_r : a));
        xhi_r = (int)(b < 0 ? 0 : (b > e_r ? e_r : b));
    }
    uint32_t *mw = tws, *pre = mw + TR_TILE_W, *ent = pre + TR_TILE_W;
    uint16_t *pos = reinterpret_cast<uint16_t *>(ent + 2 * TR_CAP);
    long long own = -2, look = -2;
    int cur = (int)(tau_hi - s);
    int tile_wlo = 0, tile_whi = -1;  // invalid
    const int wlo = lo_r >> 5;
    // path-score piece 2 (see ll_assemble): sum over this chunk's own steps of (Tg - g) * inc', inc' being the
    // increment of the path score relative to the all-noise path -- non-zero only inside spikes (A y + BW per
    // step), where a chain is entered (eH or eT) and where noise is re-entered from a tail (eG).  The y values
    // of a spike are loaded when it is painted and consumed one spike later, so the walk never waits for them.
    const RingLayout &RLt = p.RL;
    const double *mdl_g = p.model + (size_t)ch * RLt.total;
    const double *gA = mdl_g + RLt.A, *gBW = mdl_g + RLt.BW, *geG = mdl_g + RLt.eG, *geH = mdl_g + RLt.eH,
                 *geT = mdl_g + RLt.eT;
    const int NPt = RLt.NP;
    const double *ys = p.y + (size_t)ch * p.y_stride + s;
    const double wgs = (double)(p.T_glob - p.t_off - s);  // weight of the chunk's first step
    double sacc = 0.0;
    double pd_y0 = 0.0, pd_a0 = 0.0, pd_c0 = 0.0, pd_w0 = 0.0, pd_y1 = 0.0, pd_a1 = 0.0, pd_c1 = 0.0, pd_w1 = 0.0;
    auto flush_pending = [&]() {
        sacc = fma(pd_w0, fma(pd_a0, pd_y0, pd_c0), sacc);
        sacc = fma(pd_w1, fma(pd_a1, pd_y1, pd_c1), sacc);
        pd_w0 = pd_w1 = 0.0;
    };
    auto add_point = [&](int t, double v) {  // (one lane) a transition weight at own step t
        if (lane == 0 && t >= xlo_r && t < xhi_r) sacc = fma(wgs - (double)t, v, sacc);
    };
    // ---- noise everywhere first (16-byte stores); the spikes are painted over it ----
    if (xhi_r > xlo_r) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(xs + xlo_r);
        int head = (int)(((16 - (addr & 15)) & 15) >> 1);
        if (head > xhi_r - xlo_r) head = xhi_r - xlo_r;
        if (lane < head) xs[xlo_r + lane] = 1;
        const int a2 = xlo_r + head;
        const int nvec = (xhi_r - a2) >> 3;
        uint4 *vp = reinterpret_cast<uint4 *>(xs + a2);
        const uint4 ones = make_uint4(0x00010001u, 0x00010001u, 0x00010001u, 0x00010001u);
        for (int k = lane; k < nvec; k += 32) vp[k] = ones;
        const int a3 = a2 + 8 * nvec;
        if (a3 + lane < xhi_r) xs[a3 + lane] = 1;
    }
    __syncwarp();
    // segment [a, b] is decoded as `state`: remember what the chunk boundaries see
    auto see = [&](int a, int b, long long state) {
        if (record_look && e_r <= b && e_r >= a) look = state;
        if (0 <= b && 0 >= a) own = state;
    };
    auto emit_spike = [&](int a, int b, long long state, int t0_r) {
        see(a, b, state);
        const int wa = a < xlo_r ? xlo_r : a, wb = b < xhi_r - 1 ? b : xhi_r - 1;
        const int ni = (int)(state & 7);
        const int base = 2 + ni * L - t0_r;
        if (p.want_ll) flush_pending();
        int t = wa + lane;
        if (t <= wb) {
            xs[t] = (int16_t)(base + t);
            if (p.want_ll) {
                const int r = t - t0_r;  // phase index 0 .. L-1
                pd_y0 = __ldg(ys + t);
                pd_a0 = __ldg(gA + r * NPt + ni);
                pd_c0 = __ldg(gBW + r * NPt + ni);
                pd_w0 = wgs - (double)t;
            }
            t += 32;
            if (t <= wb) {
                xs[t] = (int16_t)(base + t);
                if (p.want_ll) {
                    const int r = t - t0_r;
                    pd_y1 = __ldg(ys + t);
                    pd_a1 = __ldg(gA + r * NPt + ni);
                    pd_c1 = __ldg(gBW + r * NPt + ni);
                    pd_w1 = wgs - (double)t;
                }
                for (t += 32; t <= wb; t += 32) {  // (chains longer than 64 steps: K > 65)
                    xs[t] = (int16_t)(base + t);
                    const int r = t - t0_r;
                    if (p.want_ll)
                        sacc = fma(wgs - (double)t, fma(__ldg(gA + r * NPt + ni), __ldg(ys + t), __ldg(gBW + r * NPt + ni)), sacc);
                }
            }
        }
    };
    while (cur >= lo_r) {
        if (st < 0) {
            // ---- make sure the tile holding `cur` is staged: mask words, ranks, decision entries ----
            const int wc = cur >> 5;
            if (wc > tile_whi || wc < tile_wlo) {
                tile_whi = wc;
                tile_wlo = wc - (TR_TILE_W - 1);
                if (tile_wlo < wlo) tile_wlo = wlo;
                const int nW = tile_whi - tile_wlo + 1;
                __syncwarp();
                for (int k = lane; k < TR_TILE_W; k += 32) {
                    uint32_t word = k < nW ? nzm_s[tile_wlo + k] : 0u;
                    if (tile_wlo + k == wlo) word &= ~((1u << (lo_r & 31)) - 1u);
                    mw[k] = word;
                }
                __syncwarp();
                int cnt = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) cnt += __popc(mw[8 * lane + q]);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += o;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                int r = incl - cnt;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    uint32_t w = mw[8 * lane + q];
                    pre[8 * lane + q] = (uint32_t)r;
                    while (w) {
                        const int bpos = __ffs(w) - 1;
                        w &= w - 1;
                        if (r < TR_CAP) pos[r] = (uint16_t)((8 * lane + q) * 32 + bpos);
                        r++;
                    }
                }
                __syncwarp();
                const int n = total < TR_CAP ? total : TR_CAP;
                for (int q = lane; q < n; q += 32) {
                    const int tp = (tile_wlo << 5) + pos[q];
                    ent[2 * q] = dec_s[tp];
                    ent[2 * q + 1] = tp >= dmin_r ? dec_s[tp - L] : 0u;
                }
                __syncwarp();
            }
            // ---- latest step tp <= cur in this tile where noise was entered from a tail ----
            int tp = -(1 << 30);
            for (int whi = wc; whi >= tile_wlo; whi -= 32) {
                const int wi = whi - lane;
                uint32_t word = 0;
                if (wi >= tile_wlo) {
                    word = mw[wi - tile_wlo];
                    if (wi == wc) {
                        const int hb = cur & 31;
                        if (hb < 31) word &= (2u << hb) - 1u;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, word != 0);
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    const uint32_t wsel = __shfl_sync(0xffffffffu, word, src);
                    tp = ((whi - src) << 5) + (31 - __clz(wsel));
                    break;
                }
            }
            if (tp == -(1 << 30)) {  // noise all the way down to the bottom of the tile
                const int bottom = (tile_wlo << 5) < lo_r ? lo_r : (tile_wlo << 5);
                see(bottom, cur, -1);
                cur = bottom - 1;
                continue;  // reaches lo -> loop ends; otherwise the next tile is staged
            }
            see(tp, cur, -1);  // (tp, cur] and tp itself are noise
            uint32_t d1, d2;
            {
                const int widx = (tp >> 5) - tile_wlo;
                const int rank = (int)pre[widx] + __popc(mw[widx] & ((1u << (tp & 31)) - 1u));
                if (rank < TR_CAP) {
                    d1 = ent[2 * rank];
                    d2 = ent[2 * rank + 1];
                } else {
                    d1 = dec_s[tp];
                    d2 = tp >= dmin_r ? dec_s[tp - L] : 0u;
                }
            }
            const int j = (int)(d1 & 15u);  // 1..N: noise at tp was entered from tail_j at tp-1
            const int t0 = tp - L;          // that chain was entered at t0 and occupies [t0, tp-1]
            const long long sp = enc_spike(s + t0, j - 1);
            if (p.want_ll) add_point(tp, __ldg(geG + (j - 1)));
            emit_spike(t0 < lo_r ? lo_r : t0, tp - 1, sp, t0);
            if (t0 < lo_r) {
                st = sp;
                cur = lo_r - 1;
                break;
            }
            const int k = (int)((d2 >> (4 * j)) & 15u);
            if (p.want_ll) add_point(t0, k == 0 ? __ldg(geH + (j - 1)) : __ldg(geT + (k - 1) * NPt + (j - 1)));
            cur = t0 - 1;
            st = (k == 0) ? -1 : enc_spike(s + t0 - L, k - 1);
        } else {
            // chain state at cur (start state, or chains entered directly from another chain's tail)
            const int i = (int)(st & 7);
            const int t0 = (int)((st >> 3) - s);
            emit_spike(t0 < lo_r ? lo_r : t0, cur, st, t0);
            if (t0 < lo_r) {
                cur = lo_r - 1;
                break;
            }
            const int k = (int)((dec_s[t0] >> (4 * (i + 1))) & 15u);
            if (p.want_ll) add_point(t0, k == 0 ? __ldg(geH + i) : __ldg(geT + (k - 1) * NPt + i));
            cur = t0 - 1;
            st = (k == 0) ? -1 : enc_spike(s + t0 - L, k - 1);
        }
    }
    if (p.want_ll) flush_pending();
    if (c == 0 && p.first_prologue) {
        // steps L .. 0 from the faithful prologue's backpointers (reference arithmetic)
        // (inside a chain the predecessor of state j is j - 1; the decision states' backpointers come from the
        // prologue kernel's compact table, staged into this warp's -- by now idle -- tile workspace)
        int curj = (st < 0) ? 0 : (1 + (int)(st & 7) * L + (int)(L - (st >> 3)));
        int16_t *t2d = reinterpret_cast<int16_t *>(tws);
        const int16_t *g2 = p.T2pro + (size_t)ch * (L + 1) * 8;
        __syncwarp();
        for (int k = lane; k < (L + 1) * 8; k += 32) t2d[k] = g2[k];
        __syncwarp();
        int16_t *xpro = t2d + (L + 1) * 8;  // states of columns 0..L (0-based), for the parallel ll terms below
        if (lane == 0) {
            for (int t = L; t >= 1; t--) {
                x[t] = (int16_t)(curj + 1);
                xpro[t] = (int16_t)curj;
                const int ph = curj == 0 ? 0 : (curj - 1) % L;  // 0: noise or a chain head
                const int d = curj == 0 ? 0 : 1 + (curj - 1) / L;
                curj = ph == 0 ? t2d[t * 8 + d] - 1 : curj - 1;
            }
            x[0] = (int16_t)(curj + 1);
            xpro[0] = (int16_t)curj;
        }
        __syncwarp();
        if (p.want_ll)
            for (int t = 1 + lane; t <= L; t += 32) {  // columns 1..L: one lane per column
                const int cj = xpro[t], pj = xpro[t - 1];
                double inc = 0.0;
                if (cj == 0) {
                    if (pj != 0) inc = __ldg(geG + (pj - 1) / L);
                } else {
                    const int ni = (cj - 1) / L, ph = (cj - 1) % L;
                    inc = fma(__ldg(gA + ph * NPt + ni), __ldg(ys + t), __ldg(gBW + ph * NPt + ni));
                    if (ph == 0) inc += pj == 0 ? __ldg(geH + ni) : __ldg(geT + ((pj - 1) / L) * NPt + ni);
                }
                sacc = fma(wgs - (double)t, inc, sacc);
            }
        own = 0;  // nothing precedes chunk 0
    }
    if (p.want_ll) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, d);
        if (lane == 0) p.ll_spike[(size_t)ch * p.nchunks_t + c] = sacc;
    }
    if (lane == 0) {
        p.own_start[(size_t)ch * p.nchunks_t + c] = own;
        if (record_look) p.look_end[(size_t)ch * p.nchunks_t + c] = look;
    }
}

// ll = sum_{t >= 1} T1[x_t, t] = (Tg - 1) T1[x_0, 0] + sum_{g >= 1} (Tg - g) inc_g, and with the increment split into the
// all-noise part and the rest, inc_g = (w_nn + c_emit - (y_g - m0)^2 / (2 sigma^2)) + inc'_g:
//   ll = (Tg - 1) p0 + (w_nn + c_emit) sum (Tg - g) - (1 / 2 sigma^2) sum (Tg - g) (y_g - m0)^2 + sum (Tg - g) inc'_g
// The second sum is piece 1 (forward producers, per forward chunk), the third piece 2 (traceback, per traceback
// chunk); both are added here in chunk order by one warp, so the result does not depend on scheduling.
// Strided sum of v[lo .. hi) by the whole CTA, eight independent loads in flight per thread (the pieces are a few
// thousand doubles straight from L2: the latency of dependent round trips is all this costs); fixed order.
__device__ __forceinline__ double cta_strided_sum(const double *v, int lo, int hi) {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; q++) acc[q] = 0.0;
    const int nthr = blockDim.x;
    for (int c = lo + threadIdx.x; c < hi; c += 8 * nthr) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = c + q * nthr;
            acc[q] += k < hi ? __ldcg(v + k) : 0.0;
        }
    }
    return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

__device__ void ll_assemble(const VitParams &p, int ch) {  // called by every thread of the CTA (<= 1024 threads)
    __shared__ double s_part[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    const int64_t lo = p.ll_lo, hi = p.ll_hi;
    const int cf0 = (int)(lo / p.Lc), cf1 = hi >= p.T ? p.nchunks : (int)(hi / p.Lc);
    const int ct0 = (int)(lo / p.Lc_t), ct1 = hi >= p.T ? p.nchunks_t : (int)(hi / p.Lc_t);
    double a = cta_strided_sum(p.ll_noise + (size_t)ch * p.nchunks * 2, 2 * cf0, 2 * cf1);
    double b = cta_strided_sum(p.ll_spike + (size_t)ch * p.nchunks_t, ct0, ct1);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    if (lane == 0) {
        s_part[0][warp] = a;
        s_part[1][warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = 0.0;
        for (int w = 0; w < nwarps; w++) {
            a += s_part[0][w];
            b += s_part[1][w];
        }
        const double *sc = p.model + (size_t)ch * p.RL.total + p.RL.scal;  // 0 w_nn, 1 c_emit, 2 two_s2, 3 m0
        const double *y = p.y + (size_t)ch * p.y_stride;
        int64_t g0 = lo + p.t_off, g1 = hi + p.t_off - 1;  // global steps [g0, g1]
        if (g0 == 0) {  // step 0 carries no increment: take its (y - m0)^2 term out of piece 1 again
            const double dd = y[0] - sc[3];
            a -= (double)p.T_glob * (dd * dd);
            g0 = 1;
        }
        const double n = (double)(g1 - g0 + 1);
        const double sumw = n * (double)p.T_glob - 0.5 * n * (double)(g0 + g1);  // sum_{g0..g1} (Tg - g)
        double ll = (sc[0] + sc[1]) * sumw - a / sc[2] + b;
        if (p.ll_with_p0) {
            const int x0 = p.x[(size_t)ch * p.x_stride] - 1;
            if (x0 != 0) {  // T1[j, 1] = emission for j != noise (:55-62), 0 for the noise state (:63)
                const RingLayout &RL = p.RL;
                const int ni = (x0 - 1) / RL.L, ph = (x0 - 1) % RL.L;
                const double *mdl = p.model + (size_t)ch * RL.total;
                const double dd0 = y[0] - sc[3];
                const double qn = sc[1] - (dd0 * dd0) / sc[2];
                // emission of state x0 = noise emission + (a y + b) with the plain b (no transition weight)
                ll += (double)(p.T_glob - 1) * (qn + fma(mdl[RL.A + ph * RL.NP + ni], y[0], mdl[RL.B0 + ph * RL.NP + ni]));
            }
        }
        p.ll_out[ch] = ll;
        if (p.res_host) p.res_host[ch * 4 + 0] = ll;
    }
}

__device__ __forceinline__ long long state_from_xend(int j, int64_t T, int L) {
    if (j == 0) return -1;
    int i = (j - 1) / L, sph = (j - 1) % L + 1;
    return enc_spike(T - sph, i);
}

#ifndef HMM_TRACE_MIN_CTAS
#define HMM_TRACE_MIN_CTAS 1
#endif
template <int N>
__global__ void __launch_bounds__(128, HMM_TRACE_MIN_CTAS) ring_vit_trace(VitParams p) {
    extern __shared__ __align__(16) uint32_t trsm[];
    const int ch = blockIdx.y + p.ch0;
    const int warp = warp_index_uniform();
    uint32_t *tws = trsm + (size_t)warp * TR_WARP_U32;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    const int L = p.RL.L;
    if (c >= p.nchunks_t) return;
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc_t;
    int64_t e = s + p.Lc_t;
    const bool last = (c == p.nchunks_t - 1);
    if (last || e > T) e = T;
    int64_t tau_hi = e + p.W - 1;
    long long st = -1;  // speculative: noise at tau_hi
    if (last || tau_hi >= T - 1) {
        tau_hi = T - 1;
        if (p.last_true_end) st = state_from_xend(p.xend[ch], T, L);  // else: ghost chunk, speculative noise
    }
    trace_chunk<N>(p, ch, c, tau_hi, st, !last, tws);
}

// Traceback verification in ONE launch: the state a chunk assumed at its end (after a look-ahead of W steps) must
// be the state its right neighbour really starts in; the last CTA re-walks the chunks that fail, right to left.
template <int N>
__global__ void __launch_bounds__(128) ring_vit_verify_trace(VitParams p) {
    extern __shared__ __align__(16) uint32_t trsm[];
    __shared__ int s_flag[2];
    const int ch = blockIdx.y + p.ch0;
    const size_t o = (size_t)ch * p.nchunks_t;
    bool bad = false;
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c < p.nchunks_t - 1) {
            bad = p.look_end[o + c] != p.own_start[o + c + 1];
            if (p.dbg_flag_every > 0 && c % p.dbg_flag_every == 0) bad = true;
            p.tr_flag[o + c] = bad ? 1 : 0;
        }
    }
    int any = 0;
    if (!last_cta_of_row_any(p.sync_cnt + ch * 4 + 1, s_flag, bad, &any)) return;
    int repaired = 0;
    if (any && threadIdx.x < 32) {
        uint32_t *tws = trsm;
        bool next_changed = false;
        for (int c = p.nchunks_t - 2; c >= 0; c--) {
            bool need = __ldcg(p.tr_flag + o + c) != 0;
            if (!need && next_changed) need = __ldcg(p.look_end + o + c) != __ldcg(p.own_start + o + c + 1);
            if (need) {
                // true state at time e_c is the (final) start state of chunk c+1
                int64_t e = (int64_t)(c + 1) * p.Lc_t;
                long long before = __ldcg(p.own_start + o + c);
                trace_chunk<N>(p, ch, c, e, __ldcg(p.own_start + o + c + 1), false, tws);
                __threadfence();
                __syncwarp();
                next_changed = (__ldcg(p.own_start + o + c) != before);
                repaired++;
            } else
                next_changed = false;
        }
    }
    if (threadIdx.x == 0) {
        p.counters[ch * 4 + 1] = repaired;
        if (p.res_host) p.res_host[ch * 4 + 2] = (double)repaired;
    }
    if (p.want_ll) {  // the whole CTA: a repair may have rewritten a chunk's piece
        __threadfence();
        __syncthreads();
        ll_assemble(p, ch);
    }
}

// ---------------------------------------------------------------------------
// ll = sum_{i=T..2} T1[x[i], i]   (src/viterbi.jl:92-96), parallel form:
// T1 along the decoded path is p_t = p_0 + sum_{u<=t} inc_u with
// inc_u = lp(x_{u-1} -> x_u) + q_u(x_u), hence  ll = (T-1) p_0 + sum_u (T-u) inc_u.
// ---------------------------------------------------------------------------
// One launch: every CTA accumulates its share with a grid-stride loop over groups of 8 steps; the last CTA to finish
// adds the partial sums in index order (so the result does not depend on which CTA happens to be last) and writes ll.
// A group that lies inside a noise run (x = 1 throughout, 73 % of the samples at config 2) takes a path without any
// table look-up; the arithmetic per step -- fma(T - t, lp + q, acc) -- is the same on both paths.
__global__ void __launch_bounds__(256)
    ring_path_ll(const double *__restrict__ y, int64_t T, int64_t y_stride, const char *blob, size_t blob_stride,
                 FaithfulLayout L, int ns, int nt, const int16_t *__restrict__ x, int64_t x_stride,
                 double *__restrict__ partial /*[C x gridDim.x]*/, int64_t t_lo, int64_t t_hi, int64_t t_off,
                 int64_t T_glob, unsigned *sync_cnt /*[C x 4], slot 2*/, double *__restrict__ ll_out, int with_p0,
                 double *res_host, int ch0, int tables_in_global) {
    // steps t in [t_lo, t_hi) of this (local) buffer; global time = t + t_off, weights use T_glob
    extern __shared__ __align__(16) char llsm[];
    __shared__ int s_flag;
    const int ch = blockIdx.y + ch0;
    const char *mb = blob + (size_t)ch * blob_stride;
    const double *sc = (const double *)(mb + L.scal);
    const double c_emit = sc[2], inv2s2 = 1.0 / sc[3];
    // models of many thousand states (generic engine): the tables stay in the model blob (L2), no single-predecessor
    // shortcut table
    const double *sm_m, *sm_lp, *sm_lp1 = nullptr;
    const int *sm_ptr, *sm_src, *sm_pred = nullptr;
    if (tables_in_global) {
        sm_m = (const double *)(mb + L.m);
        sm_lp = (const double *)(mb + L.in_lp);
        sm_ptr = (const int *)(mb + L.in_ptr);
        sm_src = (const int *)(mb + L.in_src);
    } else {
        double *w_m = (double *)llsm, *w_lp = w_m + ns, *w_lp1 = w_lp + nt;
        int *w_ptr = (int *)(w_lp1 + ns), *w_src = w_ptr + ns + 1, *w_pred = w_src + nt;
        sm_m = w_m; sm_lp = w_lp; sm_lp1 = w_lp1; sm_ptr = w_ptr; sm_src = w_src; sm_pred = w_pred;
        const double *gm = (const double *)(mb + L.m), *glp = (const double *)(mb + L.in_lp);
        const int *gp = (const int *)(mb + L.in_ptr), *gs = (const int *)(mb + L.in_src);
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            w_m[i] = gm[i];
            // states with exactly one predecessor (chain interiors: nearly every non-noise step of a path): one look-up
            const int e0 = gp[i], e1 = gp[i + 1];
            w_pred[i] = e1 - e0 == 1 ? gs[e0] : -1;
            w_lp1[i] = e1 - e0 == 1 ? glp[e0] : 0.0;
        }
        for (int i = threadIdx.x; i <= ns; i += blockDim.x) w_ptr[i] = gp[i];
        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
            w_lp[i] = glp[i];
            w_src[i] = gs[i];
        }
    }
    __syncthreads();
    const double *yc = y + (size_t)ch * y_stride;
    const int16_t *xc = x + (size_t)ch * x_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(xc) & 15) == 0) && ((reinterpret_cast<uintptr_t>(yc) & 15) == 0);
    // noise -> noise: the first in-edge of state 0 (lists are sorted by source); NaN if the model has none
    const double m0 = sm_m[0];
    const double w_nn = (sm_ptr[1] > sm_ptr[0] && sm_src[sm_ptr[0]] == 0) ? sm_lp[sm_ptr[0]]
                                                                          : __longlong_as_double(0x7ff8000000000000LL);
    double acc = 0.0;
    const int64_t ngroups = (T + 7) / 8;  // group g covers steps [8g, 8g+8)
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t0 = g * 8;
        int16_t xs[8];
        double ys[8];
        if (aligned && t0 + 8 <= T) {
            *reinterpret_cast<int4 *>(xs) = __ldg(reinterpret_cast<const int4 *>(xc + t0));
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                double2 v = __ldg(reinterpret_cast<const double2 *>(yc + t0 + k));
                ys[k] = v.x;
                ys[k + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                xs[k] = t0 + k < T ? xc[t0 + k] : (int16_t)1;
                ys[k] = t0 + k < T ? yc[t0 + k] : 0.0;
            }
        }
        int s = t0 > 0 ? xc[t0 - 1] - 1 : 0;
        bool quiet = s == 0 && t0 >= 1 && t0 + 8 <= T && t0 >= t_lo && t0 + 8 <= t_hi && t0 + t_off >= 1;
#pragma unroll
        for (int k = 0; k < 8; k++) quiet = quiet && xs[k] == 1;
        if (quiet) {
            const double w0 = (double)(T_glob - (t0 + t_off));
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dd = ys[k] - m0;
                const double q = c_emit - (dd * dd) * inv2s2;
                acc = fma(w0 - (double)k, w_nn + q, acc);
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int64_t t = t0 + k;
            const int d = xs[k] - 1;
            if (t >= 1 && t < T && t >= t_lo && t < t_hi && t + t_off >= 1) {
                double lp;
                if (sm_pred && sm_pred[d] == s)
                    lp = sm_lp1[d];
                else {
                    lp = __longlong_as_double(0x7ff8000000000000LL);
                    for (int e = sm_ptr[d]; e < sm_ptr[d + 1]; e++)
                        if (sm_src[e] == s) {
                            lp = sm_lp[e];
                            break;
                        }
                }
                const double dd = ys[k] - sm_m[d];
                const double q = c_emit - (dd * dd) * inv2s2;
                acc = fma((double)(T_glob - (t + t_off)), lp + q, acc);
            }
            s = d;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k >= 1; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)ch * gridDim.x + blockIdx.x] = red[0];
    if (!last_cta_of_row(sync_cnt + ch * 4 + 2, &s_flag)) return;
    double a2 = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) a2 += __ldcg(partial + (size_t)ch * gridDim.x + k);
    red[threadIdx.x] = a2;
    __syncthreads();
    for (int k = 128; k >= 1; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int x0 = xc[0] - 1;
        const double dd = yc[0] - sm_m[x0];
        const double p0 = x0 == 0 ? 0.0 : sc[2] - (dd * dd) / sc[3];
        const double ll = (with_p0 ? (double)(T_glob - 1) * p0 : 0.0) + red[0];
        ll_out[ch] = ll;
        if (res_host) res_host[ch * 4 + 0] = ll;
    }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
// Chunk slots per CTA: 8 (one 16-warp CTA per SM, paired producers) whenever 8 slots fit the 227 KB of shared memory.
template <int N, int R>
constexpr int fwd_slots() {
    return (sizeof(double) * ((size_t)8 * SlotSmem<N, R>::DOUBLES + 1024) <= 220 * 1024) ? 8 : 4;
}
// The dual-producer kernel (ring_vit_forward_ws2) serves the R = 4 geometries whose eight slots fit shared memory.
template <int N, int R>
constexpr bool use_ws2() {
#ifdef HMM_NO_WS2
    return false;
#else
    return R == 4 && sizeof(double) * ((size_t)8 * SlotSmem<N, 4, 2>::DOUBLES + 1024) <= 222 * 1024;
#endif
}
template <int N, int R, int LPC>
static size_t fwd_smem_bytes(const RingLayout &RL) {
    if (use_ws2<N, R>()) return sizeof(double) * (((RL.hot + 1) & ~1) + (size_t)8 * SlotSmem<N, 4, 2>::DOUBLES);
    return sizeof(double) * (((RL.hot + 1) & ~1) + (size_t)fwd_slots<N, R>() * SlotSmem<N, R>::DOUBLES);
}

// Resident warps per SM of the forward kernel (sets the one-wave chunk count).
template <int N, int R, int LPC, typename S>
static int fwd_warps_per_sm(const RingLayout &RL) {
    const size_t sm_fwd = fwd_smem_bytes<N, R, LPC>(RL);
    if constexpr (use_ws2<N, R>()) {
        HMM_CUDA(cudaFuncSetAttribute(ring_vit_forward_ws2<N, LPC, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
        return 8;  // one CTA of 8 slots per SM
    } else {
        constexpr int SLOTS = fwd_slots<N, R>();
        HMM_CUDA(cudaFuncSetAttribute(ring_vit_forward_ws<N, R, LPC, SLOTS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
        int nb = 0;
        HMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ring_vit_forward_ws<N, R, LPC, SLOTS, S>, SLOTS * 64, sm_fwd));
        return (nb > 0 ? nb : 1) * SLOTS;  // chunk slots (producer/consumer warp pairs) per SM
    }
}

static size_t prologue_smem(const VitParams &p, int *q_in_smem) {
    const size_t ns = p.ns, N = p.RL.N, L = p.RL.L, cols = L + 1, ND = N + 1;
    const size_t small = sizeof(double) * (ns + L * N + cols * ND) + sizeof(int) * ((cols * ND + 1) & ~(size_t)1);
    const size_t withq = small + sizeof(double) * cols * ns;
    *q_in_smem = withq <= 200 * 1024 ? 1 : 0;  // else the emissions live in the (otherwise unused) T1pro scratch
    return *q_in_smem ? withq : small;
}
static size_t trace_smem(const VitParams &p, int warps) { return sizeof(uint32_t) * (size_t)warps * TR_WARP_U32; }

// Kernel attributes are set once per plan (not per launch: the launches may be captured into a CUDA graph).
template <int N, int R, int LPC, typename S>
static void stage_prepare(const VitParams &p) {
    const size_t sm_fwd = fwd_smem_bytes<N, R, LPC>(p.RL);
    if constexpr (use_ws2<N, R>())
        HMM_CUDA(cudaFuncSetAttribute(ring_vit_forward_ws2<N, LPC, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
    else
        HMM_CUDA(cudaFuncSetAttribute(ring_vit_forward_ws<N, R, LPC, fwd_slots<N, R>(), S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
    int qs = 0;
    const size_t sm_pro = prologue_smem(p, &qs);
    if (sm_pro > 227 * 1024) fail(HMM_EUNSUPPORTED, "prologue does not fit shared memory");
    HMM_CUDA(cudaFuncSetAttribute(ring_vit_prologue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pro));
    const size_t sm_rep = sizeof(double) * (((p.RL.hot + 1) & ~1) + WarpSmem<N, R>::DOUBLES);
    HMM_CUDA(cudaFuncSetAttribute(ring_vit_verify_fwd<N, R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_rep));
    HMM_CUDA(cudaFuncSetAttribute(ring_vit_trace<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trace_smem(p, 4)));
    HMM_CUDA(cudaFuncSetAttribute(ring_vit_verify_trace<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trace_smem(p, 1)));
}

template <int N, int R, int LPC, typename S>
static void stage_forward(VitParams &p, const double *hmodel /*host ring model of channel 0*/, int C, cudaStream_t st,
                          Timer *ttop) {
    constexpr int WPB = use_ws2<N, R>() ? 8 : fwd_slots<N, R>();
    const size_t sm_fwd = fwd_smem_bytes<N, R, LPC>(p.RL);
    FirCoef<N, LPC, S> coef{};
    if (LPC > 0)
        for (int r = 0; r < LPC; r++)
            for (int i = 0; i < N; i++) coef.a[r * N + i] = (S)hmodel[p.RL.A + r * p.RL.NP + i];
    dim3 gridc((p.nchunks + WPB - 1) / WPB, C);
    if (p.first_prologue) {
        int qs = 0;
        const size_t sm_pro = prologue_smem(p, &qs);
        int nth = ((p.RL.N * p.RL.L + 31) / 32) * 32;
        nth = nth < 256 ? 256 : (nth > 1024 ? 1024 : nth);
        if (qs) nth = 1024;  // (with the emissions in shared memory every thread helps computing them)
        ring_vit_prologue<<<C, nth, sm_pro, st>>>(p, qs);
    }
    if (ttop) ttop->start();
    if constexpr (use_ws2<N, R>())
        ring_vit_forward_ws2<N, LPC, S><<<gridc, 768, sm_fwd, st>>>(p, coef);
    else
        ring_vit_forward_ws<N, R, LPC, WPB, S><<<gridc, WPB * 64, sm_fwd, st>>>(p, coef);
    if (ttop) ttop->stop();
    HMM_CUDA(cudaGetLastError());
}

template <int N, int R, int LPC, typename S>
static void stage_verify_fwd(VitParams &p, int C, cudaStream_t st) {
    const size_t sm_rep = sizeof(double) * (((p.RL.hot + 1) & ~1) + WarpSmem<N, R>::DOUBLES);
    ring_vit_verify_fwd<N, R, S><<<dim3((p.nchunks + 7) / 8, C), 256, sm_rep, st>>>(p);
    HMM_CUDA(cudaGetLastError());
}

template <int N, int R, int LPC, typename S>
static void stage_trace(VitParams &p, int C, cudaStream_t st) {
    constexpr int WPB = 4;
    dim3 gridc((p.nchunks_t + WPB - 1) / WPB, C);
    ring_vit_trace<N><<<gridc, 32 * WPB, trace_smem(p, WPB), st>>>(p);
    HMM_CUDA(cudaGetLastError());
}

template <int N, int R, int LPC, typename S>
static void stage_verify_trace(VitParams &p, int C, cudaStream_t st) {
    ring_vit_verify_trace<N><<<dim3((p.nchunks_t + 127) / 128, C), 128, trace_smem(p, 1), st>>>(p);
    HMM_CUDA(cudaGetLastError());
}

// (N, LP) -> kernel variant.  LPC > 0: FIR coefficients as constant-bank operands
// (single channel per launch; K = 48 and K = 60 models); LPC = 0: generic, coefficients
// in shared memory (any K <= 97, any number of channels per launch).
// Traceback warps that are resident on the whole device at once (one chunk per warp).
template <int N>
static int trace_slots_dev() {
    static thread_local int c_dev = -1, c_val = 0;
    int dev = 0;
    HMM_CUDA(cudaGetDevice(&dev));
    if (dev == c_dev) return c_val;
    int sms = 148, nb = 1;
    const size_t smb = sizeof(uint32_t) * 4 * TR_WARP_U32;
    HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HMM_CUDA(cudaFuncSetAttribute(ring_vit_trace<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    HMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ring_vit_trace<N>, 128, smb));
    c_dev = dev;
    c_val = sms * (nb > 0 ? nb : 1) * 4;
    return c_val;
}

struct VitVariant {
    int (*trace_slots)();
    int (*warps_per_sm)(const RingLayout &);
    void (*prepare)(const VitParams &);
    void (*forward)(VitParams &, const double *, int, cudaStream_t, Timer *);
    void (*verify_fwd)(VitParams &, int, cudaStream_t);
    void (*trace)(VitParams &, int, cudaStream_t);
    void (*verify_trace)(VitParams &, int, cudaStream_t);
};
template <int N, int R, int LPC, typename S>
static VitVariant make_variant_s() {
    return VitVariant{&trace_slots_dev<N>, &fwd_warps_per_sm<N, R, LPC, S>, &stage_prepare<N, R, LPC, S>, &stage_forward<N, R, LPC, S>,
                      &stage_verify_fwd<N, R, LPC, S>, &stage_trace<N, R, LPC, S>, &stage_verify_trace<N, R, LPC, S>};
}
static thread_local bool t_pick_f32 = false;  // set by pick_variant for the helpers below
template <int N, int R, int LPC>
static VitVariant make_variant() {
    return t_pick_f32 ? make_variant_s<N, R, LPC, float>() : make_variant_s<N, R, LPC, double>();
}
template <int N, int R>
static VitVariant pick_lp(int L, bool const_ok) {
    if (const_ok && L == 59) return make_variant<N, R, 59>();  // K = 60 templates
    if (const_ok && L == 47) return make_variant<N, R, 47>();  // K = 48 templates
    return make_variant<N, R, 0>();
}
static VitVariant pick_variant(int N, int L, bool const_ok, bool f32 = false) {
    t_pick_f32 = f32;
    switch (N) {
#ifdef HMM_NO_WS2
        case 1: return make_variant<1, 8, 0>();
        case 2: return make_variant<2, 8, 0>();
        case 3: return pick_lp<3, 8>(L, const_ok);
#else
        case 1: return make_variant<1, 4, 0>();
        case 2: return make_variant<2, 4, 0>();
        case 3: return pick_lp<3, 4>(L, const_ok);
#endif
        case 4: return pick_lp<4, 4>(L, const_ok);  // R = 4: eight chunk slots fit one SM's shared memory (R = 8: four)
        case 5: return pick_lp<5, 4>(L, const_ok);
        case 6: return make_variant<6, 4, 0>();
        case 7: return make_variant<7, 4, 0>();
    }
    fail(HMM_EUNSUPPORTED, "ring engine supports 1..%d neurons", RING_MAX_N);
}

// ---------------------------------------------------------------------------
// VitPlan: buffers + staged launches of one decode (whole recording, or one time
// shard of it with ghost chunks on either side)
// ---------------------------------------------------------------------------
struct VitPlan::Impl {
    VitVariant variant;
};

VitPlan::VitPlan() : impl(new Impl), p_(new VitParams{}) {}
VitPlan::~VitPlan() {
    for (void *q : owned) cudaFree(q);
    delete impl;
    delete p_;
}

void *VitPlan::alloc(int slot, size_t bytes) {
    if (arena_base) {  // bump allocation out of a caller-provided device buffer
        size_t o = (arena_used + 255) & ~size_t(255);
        if (o + bytes > arena_cap) fail(HMM_ENOMEM, "decode arena too small (%zu + %zu > %zu)", o, bytes, arena_cap);
        arena_used = o + bytes;
        return arena_base + o;
    }
    if (!own_memory) return workspace().get((Workspace::Slot)slot, bytes);
    // HMMCUDA_DEBUG_GUARD=1 (compute-sanitizer is not available on every pool): every buffer of the plan sits between
    // two 4 KB guard zones filled with a pattern; check_guards() verifies them after a run, so an out-of-bounds write of
    // any kernel of the decode is caught by the tests that run with the switch on.
    const bool guard = getenv("HMMCUDA_DEBUG_GUARD") && atoi(getenv("HMMCUDA_DEBUG_GUARD"));
    const size_t G = guard ? 4096 : 0;
    const size_t body = ((bytes ? bytes : 16) + 255) & ~size_t(255);
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, body + 2 * G);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(HMM_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    owned.push_back(q);
    if (guard) {
        HMM_CUDA(cudaMemset(q, 0xA5, G));
        HMM_CUDA(cudaMemset((char *)q + G + body, 0xA5, G));
        guards.push_back((char *)q);
        guards.push_back((char *)q + G + body);
    }
    return (char *)q + G;
}

void VitPlan::check_guards(cudaStream_t st) {
    if (guards.empty()) return;
    HMM_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned char> h(4096);
    for (size_t k = 0; k < guards.size(); k++) {
        HMM_CUDA(cudaMemcpy(h.data(), guards[k], h.size(), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < h.size(); b++)
            if (h[b] != 0xA5)
                fail(HMM_ECUDA, "guard zone %zu (%s buffer %zu) overwritten at byte %zu: a kernel wrote out of bounds", k,
                     (k & 1) ? "after" : "before", k / 2, b);
    }
}

int ring_default_chunking(const HostModel &M0, int64_t T_total, int C, int n_gpus, int64_t *Lc_out, int64_t *W_out) {
    const int N = M0.N, L = M0.K - 1;
#ifdef HMM_NO_WS2
    const int R = (N <= 3) ? 8 : 4, SW = 32 * R;
#else
    const int R = 4, SW = 32 * R;
#endif
    RingLayout RL = ring_layout(N, L);
    const bool no_const = getenv("HMMCUDA_NO_CONST_FIR") && atoi(getenv("HMMCUDA_NO_CONST_FIR")) != 0;
    const VitVariant variant = pick_variant(N, RL.L, C == 1 && !no_const, ring_config().precision == 1);
    int64_t W = ring_config().warmup > 0 ? ring_config().warmup : 512;
    W = ((W + SW - 1) / SW) * SW;
    if (W < ((L + 32 + SW - 1) / SW) * SW) W = ((L + 32 + SW - 1) / SW) * SW;
    // HMMCUDA_DEBUG_WARMUP=n (tests only): a warm-up BELOW the look-back, down to 0 -- speculative starts then really
    // are wrong wherever a spike straddles a boundary, and only verification + repair make the decode exact
    if (const char *e = getenv("HMMCUDA_DEBUG_WARMUP")) W = (std::max<int64_t>(0, atoll(e)) / SW) * SW;
    int64_t Lc = ring_config().chunk_len;
    if (Lc <= 0)
        if (const char *e = getenv("HMMCUDA_CHUNK")) Lc = std::max<int64_t>(0, atoll(e));  // chunk length without a call
    if (Lc <= 0) {
        // (queried once per model shape: this runs on every decode call)
        static thread_local int c_dev = -1, c_sms = 0, c_N = 0, c_L = 0, c_const = -1, c_wps = 0;
        int dev = 0;
        HMM_CUDA(cudaGetDevice(&dev));
        const int is_const = (C == 1 && !no_const) ? 1 : 0;
        if (dev != c_dev || N != c_N || L != c_L || is_const != c_const) {
            HMM_CUDA(cudaDeviceGetAttribute(&c_sms, cudaDevAttrMultiProcessorCount, dev));
            c_wps = variant.warps_per_sm(RL);
            c_dev = dev; c_N = N; c_L = L; c_const = is_const;
        }
        // one chunk per resident warp: a single, full wave over every GPU
        const int64_t target_warps = (int64_t)c_sms * c_wps * (n_gpus > 0 ? n_gpus : 1);
        int64_t per_channel = (target_warps + C - 1) / C;
        Lc = (T_total + per_channel - 1) / per_channel;
        if (Lc < 4 * W) Lc = 4 * W;
    }
    Lc = ((Lc + SW - 1) / SW) * SW;
    if (Lc < W) Lc = W;
    *Lc_out = Lc;
    *W_out = W;
    return SW;
}

void VitPlan::build(const double *y_dev, int64_t T, int64_t y_stride, int C_, const std::vector<HostModel> &models,
                    const FaithfulLayout &FL_, const char *blob_dev_, int16_t *x_dev, int64_t x_stride, int64_t Lc,
                    int64_t W, bool first_prologue, bool last_true_end, cudaStream_t st) {
    VitParams &p = *p_;
    C = C_;
    FL = FL_;
    blob_dev = blob_dev_;
    M0 = models[0];
    const int N = M0.N, L = M0.K - 1, ns = M0.nstates;
    RingLayout RL = ring_layout(N, L);
    const bool no_const = getenv("HMMCUDA_NO_CONST_FIR") && atoi(getenv("HMMCUDA_NO_CONST_FIR")) != 0;
    per_channel = C > 1 && T >= 262144 && !no_const;
    impl->variant = pick_variant(N, RL.L, (C == 1 || per_channel) && !no_const, ring_config().precision == 1);
    int nchunks = (int)((T + Lc - 1) / Lc);
    // the last chunk must be long enough to hold the final look-back of L steps
    if (nchunks > 1 && T - (int64_t)(nchunks - 1) * Lc < RING_Q) nchunks--;

    // Traceback chunks: a divisor of the forward chunk that is a multiple of 32 steps (mask-word alignment) and
    // at least max(1024, 2 W) -- the longest one of at most 4096 steps, else the shortest admissible one -- the walk is latency-bound per warp, so it
    // wants several times more chunks than the forward pass has.
    // Among the admissible lengths of at most 8192 steps (one staging tile) the one whose chunk count fills whole
    // waves of resident warps best wins: estimated time = waves x (length + look-ahead).  (At config 2 a third of the
    // forward chunk -- 3 516 chunks, one wave of 3 552 warps -- beats a quarter -- 4 688 chunks, 1.3 waves -- by 17 us.)
    int tfac = 1;
    {
        const int64_t slots = std::max<int64_t>(1, (int64_t)impl->variant.trace_slots());
        const int64_t launches_C = per_channel ? 1 : C_;  // channels per traceback launch
        double best = 0.0;
        bool have = false;
        int f_last = 1;
        for (int64_t f = 1; f <= 64; f++) {  // f ascending = sub-chunk length descending
            if (Lc % f) continue;
            const int64_t lt = Lc / f;
            if (f > 1 && (lt < 1024 || lt < 2 * W)) break;
            if (lt % 32) continue;
            f_last = (int)f;
            if (lt > 8192) continue;
            const int64_t n_t = launches_C * ((T + lt - 1) / lt);
            const double cost = (double)((n_t + slots - 1) / slots) * (double)(lt + W);
            if (!have || cost < best * 0.98) {  // prefer the longer chunk unless a finer split is clearly better
                best = cost;
                tfac = (int)f;
                have = true;
            }
        }
        if (!have) tfac = f_last;  // nothing of at most 8192 steps divides the forward chunk: the shortest admissible one
    }
    const int64_t Lc_t = Lc / tfac;
    const int nchunks_t = (int)((T + Lc_t - 1) / Lc_t);
    const int64_t pcols = L + 1;
    double *T1pro = (double *)alloc(Workspace::PROLOG, sizeof(double) * (size_t)C * ns * pcols + sizeof(int16_t) * (size_t)C * pcols * 8 + 64);
    int16_t *T2pro = (int16_t *)(T1pro + (size_t)C * ns * pcols);
    hmdl.assign((size_t)C * RL.total, 0.0);
    for (int c = 0; c < C; c++) ring_pack(models[c], RL, hmdl.data() + (size_t)c * RL.total);
    const int bvec = 1 + N * L;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t r = off;
        off += (bytes + 255) & ~size_t(255);
        return r;
    };
    size_t o_model = carve(sizeof(double) * hmdl.size());
    size_t o_sb = carve(sizeof(double) * (size_t)C * nchunks * bvec);
    size_t o_eb = carve(sizeof(double) * (size_t)C * nchunks * bvec);
    size_t o_pfin = carve(sizeof(double) * (size_t)C * N * RING_Q);
    size_t o_gfin = carve(sizeof(double) * C);
    size_t o_flag = carve(sizeof(int) * (size_t)C * nchunks);
    size_t o_trflag = carve(sizeof(int) * (size_t)C * nchunks_t);
    size_t o_cnt = carve(sizeof(int) * (size_t)C * 4);
    size_t o_own = carve(sizeof(long long) * (size_t)C * nchunks_t);
    size_t o_look = carve(sizeof(long long) * (size_t)C * nchunks_t);
    size_t o_xend = carve(sizeof(int16_t) * C);
    size_t o_part = carve(sizeof(double) * (size_t)C * 1024);
    size_t o_sync = carve(sizeof(unsigned) * (size_t)C * 4);
    size_t o_ll = carve(sizeof(double) * (size_t)C);
    size_t o_lln = carve(sizeof(double) * (size_t)C * nchunks * 2);  // one partial per FIR producer warp of a slot
    size_t o_lls = carve(sizeof(double) * (size_t)C * nchunks_t);
    char *base = (char *)alloc(Workspace::CHUNKS, off);
    // pageable source: the copy is staged before cudaMemcpyAsync returns, and hmdl outlives it anyway
    HMM_CUDA(cudaMemcpyAsync(base + o_model, hmdl.data(), sizeof(double) * hmdl.size(), cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaMemsetAsync(base + o_sync, 0, sizeof(unsigned) * (size_t)C * 4, st));  // self-resetting afterwards
    HMM_CUDA(cudaMemsetAsync(base + o_cnt, 0, sizeof(int) * (size_t)C * 4, st));
    HMM_CUDA(cudaMemsetAsync(base + o_lln, 0, sizeof(double) * (size_t)C * nchunks * 2, st));

    p = VitParams{};
    p.y = y_dev;
    p.T = T;
    p.y_stride = y_stride;
    p.model = (const double *)(base + o_model);
    p.RL = RL;
    p.Lc = Lc;
    p.W = W;
    p.nchunks = nchunks;
    p.Lc_t = Lc_t;
    p.nchunks_t = nchunks_t;
    p.tfac = tfac;
    p.ns = ns;
    p.dec = (uint32_t *)alloc(Workspace::DEC, sizeof(uint32_t) * (size_t)C * T);
    p.nzmask = (uint32_t *)alloc(Workspace::MASK, sizeof(uint32_t) * (size_t)C * ((T + 31) / 32));
    p.SB = (double *)(base + o_sb);
    p.EB = (double *)(base + o_eb);
    p.bvec = bvec;
    p.fblob = blob_dev;
    p.fblob_stride = FL.bytes;
    p.FL = FL;
    p.T1pro = T1pro;
    p.Pfin = (double *)(base + o_pfin);
    p.Gfin = (double *)(base + o_gfin);
    p.fwd_flag = (int *)(base + o_flag);
    p.counters = (int *)(base + o_cnt);
    p.T2pro = T2pro;
    p.xend = (int16_t *)(base + o_xend);
    p.x = x_dev;
    p.x_stride = x_stride;
    p.own_start = (long long *)(base + o_own);
    p.look_end = (long long *)(base + o_look);
    p.tr_flag = (int *)(base + o_trflag);
    p.x_lo = 0;
    p.x_hi = T;
    p.first_prologue = first_prologue ? 1 : 0;
    p.last_true_end = last_true_end ? 1 : 0;
    p.dbg_flag_every = getenv("HMMCUDA_DEBUG_FLAG_EVERY") ? atoi(getenv("HMMCUDA_DEBUG_FLAG_EVERY")) : 0;
    p.sync_cnt = (unsigned *)(base + o_sync);
    p.res_host = nullptr;
    p.ch0 = 0;
    part = (double *)(base + o_part);
    ll_dev_ = (double *)(base + o_ll);
    p.ll_noise = (double *)(base + o_lln);
    p.ll_spike = (double *)(base + o_lls);
    p.ll_out = ll_dev_;
    p.ll_lo = 0;
    p.ll_hi = T;
    p.t_off = 0;
    p.T_glob = T;
    p.ll_with_p0 = first_prologue ? 1 : 0;
    p.want_ll = 1;
    impl->variant.prepare(p);
}

void VitPlan::forward(cudaStream_t st, Timer *ttop) {
    // A plan can be run repeatedly without clearing anything: the decision words and non-zero masks are written
    // for every step of every chunk's main range, the last chunk rewrites every Pfin slot the final state reads,
    // the repair counters are overwritten by every verification and the arrival counters re-arm themselves.
    impl->variant.forward(*p_, hmdl.data() + (size_t)p_->ch0 * p_->RL.total, launch_C(), st, ttop);
}
void VitPlan::verify_fwd(cudaStream_t st) { impl->variant.verify_fwd(*p_, launch_C(), st); }
void VitPlan::trace(cudaStream_t st) { impl->variant.trace(*p_, launch_C(), st); }
void VitPlan::verify_trace(cudaStream_t st) { impl->variant.verify_trace(*p_, launch_C(), st); }
int VitPlan::launch_C() const { return per_channel ? 1 : C; }

static size_t ll_smem(const HostModel &M0) {
    return sizeof(double) * (2 * (size_t)M0.nstates + M0.ntrans) + sizeof(int) * (2 * (size_t)M0.nstates + 1 + M0.ntrans) + 16;
}

// The path score is assembled by verify_trace from the pieces the forward pass and the traceback leave behind
// (ll_assemble); this only hands it over.  The range must be the one configured with set_ll_range.
void VitPlan::path_ll(cudaStream_t st, double *ll_dev, int64_t t_lo, int64_t t_hi, int64_t t_off, int64_t T_glob,
                      bool with_p0) {
    VitParams &p = *p_;
    if (t_lo != p.ll_lo || t_hi != p.ll_hi || t_off != p.t_off || T_glob != p.T_glob || (with_p0 ? 1 : 0) != p.ll_with_p0 ||
        !p.want_ll)
        fail(HMM_EINVAL, "path_ll: range differs from the plan's configured ll range");
    if (ll_dev && ll_dev != ll_dev_)
        HMM_CUDA(cudaMemcpyAsync(ll_dev, ll_dev_, sizeof(double) * (size_t)C, cudaMemcpyDeviceToDevice, st));
}
void VitPlan::set_ll_range(int64_t lo, int64_t hi, int64_t t_off, int64_t T_glob, bool with_p0, bool want) {
    VitParams &p = *p_;
    p.ll_lo = lo; p.ll_hi = hi; p.t_off = t_off; p.T_glob = T_glob; p.ll_with_p0 = with_p0 ? 1 : 0; p.want_ll = want ? 1 : 0;
}

// Stand-alone path score of a decoded x (host-pointer pipeline: one pass over the whole recording at the end).
// `scratch`: >= 592 doubles followed by 4 zero-initialised unsigned arrival counters.
void ring_path_ll_run(const double *y_dev, int64_t T, const FaithfulLayout &FL, const char *blob_dev, const HostModel &M0,
                      const int16_t *x_dev, double *ll_dev, double *scratch, cudaStream_t st) {
    const int nparts = 592;
    unsigned *cnt = reinterpret_cast<unsigned *>(scratch + nparts);
    HMM_CUDA(cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned), st));
    const bool in_global = ll_smem(M0) > 200 * 1024;  // (the CLI's overlap models: 10 000+ states)
    const size_t smb = in_global ? 16 : ll_smem(M0);
    if (smb > 48 * 1024)  // overlap models (thousands of states) through the generic engine
        HMM_CUDA(cudaFuncSetAttribute(ring_path_ll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    ring_path_ll<<<dim3(nparts, 1), 256, smb, st>>>(y_dev, T, T, blob_dev, FL.bytes, FL, M0.nstates, (int)M0.ntrans, x_dev, T,
                                                    scratch, 0, T, 0, T, cnt, ll_dev, 1, nullptr, 0, in_global ? 1 : 0);
    HMM_CUDA(cudaGetLastError());
}

void VitPlan::read_counters(cudaStream_t st, int *fwd_rep, int *bwd_rep) {
    VitParams &p = *p_;
    std::vector<int> cnt((size_t)C * 4);
    HMM_CUDA(cudaMemcpyAsync(cnt.data(), p.counters, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaStreamSynchronize(st));
    *fwd_rep = *bwd_rep = 0;
    for (int c = 0; c < C; c++) {
        *fwd_rep += cnt[c * 4 + 0];
        *bwd_rep += cnt[c * 4 + 1];
    }
}

void VitPlan::reset_counters(cudaStream_t st) {
    HMM_CUDA(cudaMemsetAsync(p_->counters, 0, sizeof(int) * (size_t)C * 4, st));
}

void VitPlan::use_arena(char *base, size_t cap) {
    arena_base = base;
    arena_cap = cap;
    arena_used = 0;
}
size_t VitPlan::arena_bytes_used() const { return arena_used; }
void VitPlan::set_x_window(int64_t lo, int64_t hi) {
    p_->x_lo = lo;
    p_->x_hi = hi;
}
void VitPlan::set_result_sink(double *res_dev_alias) { p_->res_host = res_dev_alias; }
void VitPlan::retarget(const double *y_dev, int16_t *x_dev) {
    p_->y = y_dev;
    p_->x = x_dev;
}

int *VitPlan::counters_ptr() { return p_->counters; }
int VitPlan::nchunks() const { return p_->nchunks; }
int VitPlan::bvec() const { return p_->bvec; }
double *VitPlan::eb_ptr(int chunk) { return p_->EB + (size_t)chunk * p_->bvec; }
double *VitPlan::sb_ptr(int chunk) { return p_->SB + (size_t)chunk * p_->bvec; }
long long *VitPlan::own_start_ptr(int chunk) { return p_->own_start + (size_t)chunk * p_->tfac; }

// The whole decode of a plan: five launches (per channel, when the channels are long enough to fill the GPU on
// their own: the FIR then takes its coefficients from the constant bank), nothing else.
void VitPlan::run_all(cudaStream_t st, bool want_ll, Timer *ttop) {
    p_->want_ll = want_ll ? 1 : 0;
    const int nrun = per_channel ? C : 1;
    for (int k = 0; k < nrun; k++) {
        p_->ch0 = per_channel ? k : 0;
        {
            NvtxRange r("hmm.viterbi.forward");
            forward(st, k == 0 ? ttop : nullptr);
        }
        {
            NvtxRange r("hmm.viterbi.verify_forward");
            verify_fwd(st);
        }
        {
            NvtxRange r("hmm.viterbi.traceback");
            trace(st);
        }
        {
            NvtxRange r("hmm.viterbi.verify_traceback+ll");
            verify_trace(st);  // (assembles ll as well)
        }
    }
    p_->ch0 = 0;
}

// ---------------------------------------------------------------------------
// Cached decode programs.  Analysing and packing the model, carving the buffers, a dozen launches and two
// synchronisations cost ~60 us of host time per decode of a 0.5 ms step; a caller that decodes the same buffers
// with the same model again (every benchmark step, every chunk group of a probe, a Julia loop over trials)
// re-launches ONE CUDA graph instead and synchronises once.  Key = everything the launches depend on.  The
// kernels leave ll and the repair counts in mapped pinned memory, so there is no device -> host copy either.
// ---------------------------------------------------------------------------
namespace {

struct ProgKey {
    int dev;
    const double *y;
    int16_t *x;
    int64_t T, y_stride, x_stride, Lc, W;
    int C, want_ll, dbg, prec;
    uint64_t model_id;
    bool operator==(const ProgKey &o) const {
        return dev == o.dev && y == o.y && x == o.x && T == o.T && y_stride == o.y_stride && x_stride == o.x_stride &&
               Lc == o.Lc && W == o.W && C == o.C && want_ll == o.want_ll && dbg == o.dbg && prec == o.prec &&
               model_id == o.model_id;
    }
};

struct RingProgram {
    ProgKey key;
    VitPlan plan;
    cudaGraphExec_t exec = nullptr;
    double *res_h = nullptr;  // mapped pinned: [C x 4] ll, forward repairs, traceback repairs
    int runs = 0;
    uint64_t epoch = 0;
    ~RingProgram() {
        if (exec) cudaGraphExecDestroy(exec);
        if (res_h) cudaFreeHost(res_h);
    }
};

// per host thread, like the workspace (the multi-device dispatcher runs one worker thread per GPU)
thread_local std::vector<std::unique_ptr<RingProgram>> g_programs;
thread_local uint64_t g_epoch = 0;
constexpr size_t MAX_PROGRAMS = 40;

}  // namespace

void ring_new_epoch() { g_epoch++; }

void ring_drop_programs() { g_programs.clear(); }

void ring_collect(std::vector<RingPending> &pend) {
    for (auto &q : pend) {
        int f = 0, b = 0;
        for (int c = 0; c < q.C; c++) {
            if (q.ll_host) q.ll_host[c] = q.res_h[c * 4 + 0];
            f += (int)q.res_h[c * 4 + 1];
            b += (int)q.res_h[c * 4 + 2];
        }
        if (q.info) {
            q.info->fwd_repaired += f;
            q.info->bwd_repaired += b;
        }
    }
    pend.clear();
}

void ring_viterbi_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const std::vector<HostModel> &models,
                      const FaithfulLayout &FL, const char *blob_dev, uint64_t model_id, int16_t *x_dev,
                      int64_t x_stride, double *ll_host, cudaStream_t st, hmm_info *info,
                      std::vector<RingPending> *defer) {
    // Long recordings: one channel per LAUNCH (each already fills the GPU, and the FIR then takes its coefficients
    // from the constant bank) out of one C-channel plan -- and one CUDA graph for all of them.
    const bool no_const = getenv("HMMCUDA_NO_CONST_FIR") && atoi(getenv("HMMCUDA_NO_CONST_FIR")) != 0;
    const bool per_channel = C > 1 && T >= 262144 && !no_const;
    std::vector<RingPending> local;
    std::vector<RingPending> *pend = defer ? defer : &local;
    int64_t Lc = 0, W = 0;
    ring_default_chunking(models[0], T, per_channel ? 1 : C, 1, &Lc, &W);
    ProgKey key{};
    HMM_CUDA(cudaGetDevice(&key.dev));
    key.y = y_dev; key.x = x_dev; key.T = T; key.y_stride = y_stride; key.x_stride = x_stride; key.Lc = Lc; key.W = W;
    key.C = C; key.want_ll = ll_host ? 1 : 0; key.model_id = model_id;
    key.dbg = getenv("HMMCUDA_DEBUG_FLAG_EVERY") ? atoi(getenv("HMMCUDA_DEBUG_FLAG_EVERY")) : 0;
    key.prec = ring_config().precision;
    const bool profiling = ring_config().profile != 0;
    const bool cacheable = model_id != 0 && !(getenv("HMMCUDA_NO_GRAPH") && atoi(getenv("HMMCUDA_NO_GRAPH")));
    RingProgram *prog = nullptr;
    if (cacheable)
        for (auto &q : g_programs)
            if (q->key == key) prog = q.get();
    std::unique_ptr<RingProgram> fresh;
    if (!prog) {
        fresh.reset(new RingProgram);
        fresh->key = key;
        fresh->plan.own_memory = cacheable;  // a one-off plan lives in the thread's grow-only workspace (no cudaMalloc)
        HMM_CUDA(cudaHostAlloc((void **)&fresh->res_h, sizeof(double) * 4 * (size_t)C, cudaHostAllocMapped));
        memset(fresh->res_h, 0, sizeof(double) * 4 * (size_t)C);
        double *res_d = nullptr;
        HMM_CUDA(cudaHostGetDevicePointer((void **)&res_d, fresh->res_h, 0));
        fresh->plan.build(y_dev, T, y_stride, C, models, FL, blob_dev, x_dev, x_stride, Lc, W, true, true, st);
        fresh->plan.set_result_sink(res_d);
        prog = fresh.get();
        if (cacheable) {
            if (g_programs.size() >= MAX_PROGRAMS) {  // drop the least recently used program of an EARLIER call
                size_t victim = g_programs.size();
                for (size_t k = 0; k < g_programs.size(); k++)
                    if (g_programs[k]->epoch < g_epoch && (victim == g_programs.size() || g_programs[k]->epoch < g_programs[victim]->epoch))
                        victim = k;
                if (victim < g_programs.size()) g_programs.erase(g_programs.begin() + victim);
            }
            g_programs.push_back(std::move(fresh));
        }
    }
    prog->epoch = g_epoch;
    prog->runs++;
    const bool want_ll = ll_host != nullptr;
    bool launched = false;
    if (cacheable && !profiling && prog->runs >= 2) {
        if (!prog->exec) {  // second use of this program: capture its launches once
            cudaGraph_t graph = nullptr;
            HMM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                prog->plan.run_all(st, want_ll, nullptr);
            } catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            HMM_CUDA(cudaStreamEndCapture(st, &graph));
            cudaError_t e = cudaGraphInstantiate(&prog->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                prog->exec = nullptr;
                fail(HMM_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
            }
        }
        HMM_CUDA(cudaGraphLaunch(prog->exec, st));
        launched = true;
    }
    Timer *ttop = nullptr;
    std::unique_ptr<Timer> tt;
    if (!launched) {
        if (profiling && info) {
            tt.reset(new Timer(st));
            ttop = tt.get();
        }
        prog->plan.run_all(st, want_ll, ttop);
    }
    if (info) {
        info->kernel_launches += (int64_t)5 * (per_channel ? C : 1);
        info->n_chunks = prog->plan.nchunks();
    }
    pend->push_back(RingPending{prog->res_h, C, ll_host, info});
    prog->plan.check_guards(st);  // (HMMCUDA_DEBUG_GUARD only; synchronises)
    if (!defer || fresh || ttop) {
        // an uncached program dies with this call, and a profiled run reads its event timer: synchronise here
        HMM_CUDA(cudaStreamSynchronize(st));
        if (ttop && info) info->top_kernel_ms = ttop->ms();
        if (!defer) ring_collect(local);
        else if (fresh) ring_collect(*defer);
    }
}

}  // namespace hmm
