/*
 * hmm_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A literal, single-threaded, Float64 restatement of the HMM inference hot
 * path of grero/HMMSpikeSorter.jl, in the reference's own operation order.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  The product
 * (libhmmcuda.so) never links, loads or calls it.
 *
 * PARITY PINNING STATUS: the reference is pure Julia and Julia is not
 * installed in the build container nor on the GPU box, so the reference
 * itself cannot be executed here.  The reference's tests hold no golden
 * vectors for T1/T2/x/alpha/beta/gamma/ll/mu/sigma/lp.
 * This oracle is pinned to what the reference's tests DO fix
 * (test/runtests.jl:36-42 "Unroll" state layout; :55 template energy
 * 100.66411692920131) plus brute-force path enumeration, an independent
 * dense formulation and invariants (tests/test_oracle_*.py).  Last-ulp
 * behaviour of Julia Base log/exp/log1p versus glibc's is "parity unpinned".
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile)
 * so that no FMA contraction or reassociation changes the reference's
 * rounding sequence.
 *
 * All indices crossing this API are 1-based and all matrices column-major,
 * exactly as the Julia arrays they restate.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ENOMEM 1
#define ORC_EARG 2

typedef struct {
    int64_t src; /* 1-based */
    int64_t dst; /* 1-based */
    double lp;
} orc_trans; /* == Julia Tuple{Int64,Int64,Float64}, src/types.jl:3 */

/* src/utils.jl:1  const log2pi = 0.5*log(2*pi) */
static const double ORC_LOG2PI = 0.9189385332046727;

/* src/utils.jl:3  funcl(x, mu, sigma) -- recomputes log(sigma) per call. */
static inline double funcl3(double x, double mu, double sigma) {
    double s2 = sigma * sigma;
    double dd = x - mu;
    return (-ORC_LOG2PI - log(sigma)) - (dd * dd) / (2 * s2);
}

/* src/utils.jl:4  funcl(x, mu, sigma, lsigma) */
static inline double funcl4(double x, double mu, double sigma, double lsig) {
    double s2 = sigma * sigma;
    double dd = x - mu;
    return (-ORC_LOG2PI - lsig) - (dd * dd) / (2 * s2);
}

/* src/utils.jl:24-32  logsumexpl(xp, yp) */
static inline double logsumexpl(double xp, double yp) {
    double z;
    if (xp > yp)
        z = xp + log1p(exp(yp - xp));
    else
        z = yp + log1p(exp(xp - yp));
    return z;
}

double orc_funcl3(double x, double mu, double sigma) { return funcl3(x, mu, sigma); }
double orc_funcl4(double x, double mu, double sigma, double ls) { return funcl4(x, mu, sigma, ls); }
double orc_logsumexpl(double a, double b) { return logsumexpl(a, b); }

/* ------------------------------------------------------------------ */
/* StateMatrix construction, src/types.jl:65-127,148-151              */
/* ------------------------------------------------------------------ */

/* src/types.jl:66-70  number of columns generate_states allocates. */
int64_t orc_nstates(int64_t N, int64_t K, int allow_overlaps) {
    int64_t n = 1 + N * (K - 1);
    if (allow_overlaps) n += (N * (N - 1) * (K - 1) * (K - 1)) / 2;
    return n;
}

/* src/types.jl:65-92  generate_states: 0-based phases, [N x nstates] col-major. */
int orc_generate_states(int64_t N, int64_t K, int allow_overlaps, int16_t *states /* zero-filled by us */) {
    int64_t ns = orc_nstates(N, K, allow_overlaps);
    memset(states, 0, sizeof(int16_t) * (size_t)(N * ns));
    int64_t k = 1; /* 0-based column; Julia k = 2 */
    for (int64_t i = 0; i < N; i++)
        for (int64_t k1 = 1; k1 <= K - 1; k1++) {
            states[i + N * k] = (int16_t)k1;
            k++;
        }
    if (allow_overlaps)
        for (int64_t i = 0; i < N - 1; i++)
            for (int64_t j = i + 1; j < N; j++)
                for (int64_t k1 = 1; k1 <= K - 1; k1++)
                    for (int64_t k2 = 1; k2 <= K - 1; k2++) {
                        states[i + N * k] = (int16_t)k1;
                        states[j + N * k] = (int16_t)k2;
                        k++;
                    }
    return ORC_OK;
}

/* src/types.jl:94-113  isvalid_transition on 0-based phases; j1, j2 0-based
 * columns.  lpz = log1p(-exp(sum(lp))) is a pure function of lp and is
 * passed in (the reference recomputes the identical value on every call). */
static double isvalid_transition(const int16_t *states, int64_t N, int64_t K, const double *lp, double lpz,
                                 int64_t j1, int64_t j2) {
    double lpt = 0.0;
    for (int64_t i = 0; i < N; i++) {
        int s1 = states[i + N * j1];
        int s2 = states[i + N * j2];
        double lpi = lp[i];
        if (s1 == 0 && s2 == 0) {
            lpt += lpz;
        } else if (s1 == 0 && s2 == 1) {
            lpt += lpi;
        } else if ((s2 - s1 == 1) || (s1 == K - 1 && s2 == 0)) {
            lpt += 0.0;
        } else {
            lpt = -INFINITY;
            break;
        }
    }
    return lpt;
}

/* sum(lp): Julia's mapreduce is a plain left-to-right loop for the array
 * lengths used here (N <= a handful), src/types.jl:96. */
double orc_lpz(const double *lp, int64_t N) {
    double s = 0.0;
    if (N > 0) {
        s = lp[0];
        for (int64_t i = 1; i < N; i++) s += lp[i];
    }
    return log1p(-exp(s));
}

/* src/types.jl:115-127  get_valid_transitions: row-major scan (src outer,
 * dst inner) -> list sorted by (src, dst).  Returns count; if tr == NULL
 * only counts. */
int64_t orc_get_valid_transitions(const int16_t *states0, int64_t N, int64_t nstates, int64_t K, const double *lp,
                                  int64_t nlp, orc_trans *tr, int64_t cap) {
    /* sum(lp) runs over the WHOLE vector (nlp >= N entries: update() hands xb[2:end], which for overlap
     * models has one entry per transition out of the silent state, src/baumwelch.jl:226,265), while
     * isvalid_transition only indexes lp[i], i < N. */
    double lpz = orc_lpz(lp, nlp < N ? N : nlp);
    int64_t n = 0;
    for (int64_t i = 0; i < nstates; i++)
        for (int64_t j = 0; j < nstates; j++) {
            double aa = isvalid_transition(states0, N, K, lp, lpz, i, j);
            if (isfinite(aa)) {
                if (tr) {
                    if (n >= cap) return -1;
                    tr[n].src = i + 1;
                    tr[n].dst = j + 1;
                    tr[n].lp = aa;
                }
                n++;
            }
        }
    return n;
}

/* ------------------------------------------------------------------ */
/* helpers                                                            */
/* ------------------------------------------------------------------ */

/* state mean m[j] = sum_{l=1..N} mu[states[l,j], l], starting from 0.0
 * (src/viterbi.jl:58-60,68-71; src/baumwelch.jl:32-35,82-86,211-215). */
static void state_means(const int16_t *states1, int64_t N, int64_t nstates, const double *mu, int64_t K, double *m) {
    for (int64_t j = 0; j < nstates; j++) {
        double s = 0.0;
        for (int64_t l = 0; l < N; l++) s += mu[(states1[l + N * j] - 1) + K * l];
        m[j] = s;
    }
}

void orc_state_means(const int16_t *states1, int64_t N, int64_t nstates, const double *mu, int64_t K, double *m) {
    state_means(states1, N, nstates, mu, K, m);
}

/* ------------------------------------------------------------------ */
/* Viterbi, src/viterbi.jl:44-98                                      */
/* ------------------------------------------------------------------ */

/* One forward column: cur <- step(prev) exactly as src/viterbi.jl:66-87.
 * t2 may be NULL. */
static void viterbi_column(const double *prev, double *cur, int16_t *t2, double yi, const double *m, double sigma,
                           double lsig, int64_t nstates, const orc_trans *tr, int64_t ntrans, double *q) {
    for (int64_t j = 0; j < nstates; j++) {
        q[j] = funcl4(yi, m[j], sigma, lsig);
        cur[j] = -INFINITY; /* :52 fill(-Inf) */
        if (t2) t2[j] = 1;  /* :53 ones(Int16) */
    }
    for (int64_t e = 0; e < ntrans; e++) {
        int64_t k = tr[e].src - 1, j = tr[e].dst - 1;
        double t = prev[k] + tr[e].lp;
        if (t > cur[j]) {
            cur[j] = t;
            if (t2) t2[j] = (int16_t)(k + 1);
        }
    }
    for (int64_t j = 0; j < nstates; j++) cur[j] += q[j];
}

/*
 * viterbi(y, lA::StateMatrix, mu, sigma) -> (x, ll)   src/viterbi.jl:44-98
 * T1_out / T2_out (nullable) receive the dense [nstates x T] trellis of
 * :52-53.  Without them the trellis is rebuilt block by block from stored
 * checkpoint columns (identical arithmetic => bit-identical x and ll) so
 * that long sequences fit in host memory.
 */
int orc_viterbi(const double *y, int64_t T, const int16_t *states1, int64_t N, int64_t K, int64_t nstates,
                const orc_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out, double *ll_out,
                int16_t *T2_out, double *T1_out) {
    if (T < 1 || nstates < 1 || nstates > 32767) return ORC_EARG;
    /* Steps i = 1..T-1 (0-based columns) are grouped in blocks of B; the
     * column preceding each block is kept as a checkpoint. */
    const int64_t B = 4096;
    int64_t nsteps = T - 1;
    int64_t nblk = (nsteps + B - 1) / B;
    double lsig = log(sigma); /* :47 */
    double *m = malloc(sizeof(double) * nstates);
    double *q = malloc(sizeof(double) * nstates);
    double *ckpt = malloc(sizeof(double) * nstates * (nblk ? nblk : 1));
    double *b1 = malloc(sizeof(double) * nstates * (B + 1));
    int16_t *b2 = malloc(sizeof(int16_t) * nstates * (B + 1));
    double *colA = malloc(sizeof(double) * nstates), *colB = malloc(sizeof(double) * nstates);
    if (!m || !q || !ckpt || !b1 || !b2 || !colA || !colB) return ORC_ENOMEM;
    state_means(states1, N, nstates, mu, K, m);

    /* pass 1: forward; remember the column that precedes every block */
    for (int64_t j = 0; j < nstates; j++) colA[j] = funcl4(y[0], m[j], sigma, lsig); /* :55-62 */
    colA[0] = 0; /* :63 */
    if (T1_out) memcpy(T1_out, colA, sizeof(double) * nstates);
    if (T2_out)
        for (int64_t j = 0; j < nstates; j++) T2_out[j] = 1;
    double *prev = colA, *cur = colB;
    for (int64_t i = 1; i < T; i++) {
        if ((i - 1) % B == 0) memcpy(ckpt + nstates * ((i - 1) / B), prev, sizeof(double) * nstates);
        int16_t *t2 = T2_out ? T2_out + nstates * i : NULL;
        viterbi_column(prev, cur, t2, y[i], m, sigma, lsig, nstates, tr, ntrans, q);
        if (T1_out) memcpy(T1_out + nstates * i, cur, sizeof(double) * nstates);
        double *tmp = prev;
        prev = cur;
        cur = tmp;
    }
    /* :90 argmax(T1[:,end]) -- first maximum */
    int64_t best = 0;
    for (int64_t j = 1; j < nstates; j++)
        if (prev[j] > prev[best]) best = j;
    x_out[T - 1] = (int16_t)(best + 1);

    /* pass 2: blocks in reverse; rebuild the block trellis with identical
     * arithmetic, then backtrack and accumulate ll in the order of :92-96 */
    double ll = 0.0;
    for (int64_t b = nblk - 1; b >= 0; b--) {
        int64_t i0 = 1 + b * B;                          /* first column of the block */
        int64_t i1 = i0 + B < T ? i0 + B : T;            /* one past the last */
        memcpy(b1, ckpt + nstates * b, sizeof(double) * nstates); /* column i0-1 */
        for (int64_t i = i0; i < i1; i++)
            viterbi_column(b1 + nstates * (i - i0), b1 + nstates * (i - i0 + 1), b2 + nstates * (i - i0 + 1), y[i], m,
                           sigma, lsig, nstates, tr, ntrans, q);
        for (int64_t i = i1 - 1; i >= i0; i--) {
            int64_t xi = x_out[i] - 1;
            x_out[i - 1] = b2[xi + nstates * (i - i0 + 1)];
            ll += b1[xi + nstates * (i - i0 + 1)];
        }
    }
    *ll_out = ll;
    free(m); free(q); free(ckpt); free(b1); free(b2); free(colA); free(colB);
    return ORC_OK;
}

/*
 * Near-tie screen (SURVEY 8d "inputs screened for decision margins"): one forward sweep in the reference's
 * own arithmetic (src/viterbi.jl:65-88) that records, for every destination state with two or more finite
 * candidates at every column i >= from_col (0-based; the structural exact ties of SURVEY H3 sit at column 1),
 * the margin best - second best of `T1[k,i-1] + lp`, and the margin of the final argmax (:90).
 * out[0] = smallest absolute margin, out[1] = smallest margin relative to |T1[j,i]|,
 * out[2] = number of decisions with relative margin < rel_a, out[3] = same for rel_b,
 * out[4] = column of the smallest relative margin, out[5] = number of decisions examined.
 * A decision whose margin is below a few dozen ulp of the score is decided by the reference's own rounding
 * noise; bit-exact agreement of x is only meaningful for inputs without such decisions.
 */
int orc_viterbi_screen(const double *y, int64_t T, const int16_t *states1, int64_t N, int64_t K, int64_t nstates,
                       const orc_trans *tr, int64_t ntrans, const double *mu, double sigma, int64_t from_col,
                       double rel_a, double rel_b, double *out /* [6] */) {
    if (T < 1 || nstates < 1) return ORC_EARG;
    double lsig = log(sigma);
    double *m = malloc(sizeof(double) * nstates), *q = malloc(sizeof(double) * nstates);
    double *colA = malloc(sizeof(double) * nstates), *colB = malloc(sizeof(double) * nstates);
    double *sec = malloc(sizeof(double) * nstates);
    if (!m || !q || !colA || !colB || !sec) return ORC_ENOMEM;
    state_means(states1, N, nstates, mu, K, m);
    for (int64_t j = 0; j < nstates; j++) colA[j] = funcl4(y[0], m[j], sigma, lsig);
    colA[0] = 0;
    double min_abs = INFINITY, min_rel = INFINITY, na = 0, nb = 0, argcol = -1, nexam = 0;
    double *prev = colA, *cur = colB;
    for (int64_t i = 1; i < T; i++) {
        for (int64_t j = 0; j < nstates; j++) {
            q[j] = funcl4(y[i], m[j], sigma, lsig);
            cur[j] = -INFINITY;
            sec[j] = -INFINITY;
        }
        for (int64_t e = 0; e < ntrans; e++) {
            int64_t k = tr[e].src - 1, j = tr[e].dst - 1;
            double t = prev[k] + tr[e].lp;
            if (t > cur[j]) {
                sec[j] = cur[j];
                cur[j] = t;
            } else if (t > sec[j])
                sec[j] = t;
        }
        for (int64_t j = 0; j < nstates; j++) {
            const double best = cur[j];
            cur[j] += q[j];
            if (i >= from_col && isfinite(sec[j]) && isfinite(best)) {
                double mg = best - sec[j], sc = fabs(cur[j]);
                double rel = sc > 0 ? mg / sc : INFINITY;
                nexam += 1;
                if (mg < min_abs) min_abs = mg;
                if (rel < min_rel) {
                    min_rel = rel;
                    argcol = (double)i;
                }
                if (rel < rel_a) na += 1;
                if (rel < rel_b) nb += 1;
            }
        }
        double *tmp = prev;
        prev = cur;
        cur = tmp;
    }
    {   /* final argmax, :90 */
        double b1 = -INFINITY, b2 = -INFINITY;
        for (int64_t j = 0; j < nstates; j++) {
            if (prev[j] > b1) {
                b2 = b1;
                b1 = prev[j];
            } else if (prev[j] > b2)
                b2 = prev[j];
        }
        if (isfinite(b2) && T - 1 >= from_col) {
            double mg = b1 - b2, rel = fabs(b1) > 0 ? mg / fabs(b1) : INFINITY;
            nexam += 1;
            if (mg < min_abs) min_abs = mg;
            if (rel < min_rel) {
                min_rel = rel;
                argcol = (double)(T - 1);
            }
            if (rel < rel_a) na += 1;
            if (rel < rel_b) nb += 1;
        }
    }
    out[0] = min_abs; out[1] = min_rel; out[2] = na; out[3] = nb; out[4] = argcol; out[5] = nexam;
    free(m); free(q); free(colA); free(colB); free(sec);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* forward / backward, src/baumwelch.jl:25-51, 73-98                  */
/* ------------------------------------------------------------------ */

int orc_forward(const double *V, int64_t T, const int16_t *states1, int64_t N, int64_t K, int64_t nstates,
                const orc_trans *tr, int64_t ntrans, const double *mu, double sigma, double *a /* [nstates x T] */) {
    double *m = malloc(sizeof(double) * nstates);
    if (!m) return ORC_ENOMEM;
    state_means(states1, N, nstates, mu, K, m);
    for (int64_t i = 0; i < nstates * T; i++) a[i] = -INFINITY; /* :28 */
    for (int64_t i = 0; i < nstates; i++) a[i] = funcl3(V[0], m[i], sigma); /* :30-37 (pi overwritten) */
    for (int64_t i = 1; i < T; i++) {
        double v = V[i];
        double *ai = a + nstates * i;
        const double *ap = a + nstates * (i - 1);
        for (int64_t e = 0; e < ntrans; e++) {
            int64_t k = tr[e].src - 1, j = tr[e].dst - 1;
            double b = funcl3(v, m[j], sigma);
            ai[j] = logsumexpl(ai[j], ap[k] + tr[e].lp + b); /* :47 */
        }
    }
    free(m);
    return ORC_OK;
}

int orc_backward(const double *V, int64_t T, const int16_t *states1, int64_t N, int64_t K, int64_t nstates,
                 const orc_trans *tr, int64_t ntrans, const double *mu, double sigma, double *a /* [nstates x T] */) {
    double *m = malloc(sizeof(double) * nstates);
    if (!m) return ORC_ENOMEM;
    state_means(states1, N, nstates, mu, K, m);
    for (int64_t i = 0; i < nstates * T; i++) a[i] = -INFINITY; /* :79 */
    for (int64_t j = 0; j < nstates; j++) a[j + nstates * (T - 1)] = 0.0; /* :80 */
    for (int64_t i = T - 2; i >= 0; i--) {
        double v = V[i + 1];
        double *ai = a + nstates * i;
        const double *an = a + nstates * (i + 1);
        for (int64_t e = 0; e < ntrans; e++) {
            int64_t j = tr[e].src - 1, k = tr[e].dst - 1;
            double b = funcl3(v, m[k], sigma);
            ai[j] = logsumexpl(ai[j], an[k] + tr[e].lp + b); /* :94 */
        }
    }
    free(m);
    return ORC_OK;
}

/* log-likelihood of the data = LSE_j alpha[j,T]; not returned by the
 * reference -- extra output used for parity checks. */
double orc_loglik_from_alpha(const double *a, int64_t T, int64_t nstates) {
    double g = -INFINITY;
    for (int64_t j = 0; j < nstates; j++) g = logsumexpl(g, a[j + nstates * (T - 1)]);
    return g;
}

/* ------------------------------------------------------------------ */
/* update, src/baumwelch.jl:205-309                                   */
/* ------------------------------------------------------------------ */

/*
 * Outputs: lp_out[nxi-1] = xb[2:end] (:264-265; for non-overlap models
 * nxi-1 == N), pp_out[nstates] = gamma[:,1] (:263), mu updated IN PLACE
 * (:268-287), *sigma_out (:306-307).  nxi = number of
 * transitions leaving state 1.  The caller rebuilds the StateMatrix from
 * lp_out with orc_get_valid_transitions, as :265 does.
 * gamma_out (nullable) receives gamma [nstates x T] for invariant tests.
 */
int orc_update(const double *alpha, const double *beta, int64_t T, const int16_t *states1, int64_t N, int64_t K,
               int64_t nstates, const orc_trans *tr, int64_t ntrans, double *mu, double sigma, const double *x,
               double *lp_out, int64_t lp_cap, double *pp_out, double *sigma_out, double *gamma_out) {
    double *gf = gamma_out ? gamma_out : malloc(sizeof(double) * nstates * T);
    double *m = malloc(sizeof(double) * nstates);
    if (!gf || !m) return ORC_ENOMEM;
    state_means(states1, N, nstates, mu, K, m); /* :210-215 */
    for (int64_t t = 0; t < T; t++) {           /* :216-224 */
        double g = -INFINITY;
        for (int64_t j = 0; j < nstates; j++) g = logsumexpl(g, alpha[j + nstates * t] + beta[j + nstates * t]);
        for (int64_t j = 0; j < nstates; j++)
            gf[j + nstates * t] = alpha[j + nstates * t] + beta[j + nstates * t] - g;
    }
    /* :226 tidx = findall(q->q[1]==1, transitions) */
    int64_t nxi = 0;
    for (int64_t e = 0; e < ntrans; e++)
        if (tr[e].src == 1) nxi++;
    if (nxi - 1 > lp_cap) { free(m); if (!gamma_out) free(gf); return ORC_EARG; }
    int64_t *tidx = malloc(sizeof(int64_t) * (nxi ? nxi : 1));
    double *xi = malloc(sizeof(double) * (nxi ? nxi : 1) * (T > 1 ? T - 1 : 1));
    double *xx = malloc(sizeof(double) * (nxi ? nxi : 1));
    if (!tidx || !xi || !xx) return ORC_ENOMEM;
    {
        int64_t c = 0;
        for (int64_t e = 0; e < ntrans; e++)
            if (tr[e].src == 1) tidx[c++] = e;
    }
    for (int64_t t = 0; t < T - 1; t++) { /* :229-253 */
        double _x = x[t + 1];
        for (int64_t i = 0; i < nxi; i++) {
            int64_t j = tr[tidx[i]].dst - 1;
            double lp = tr[tidx[i]].lp;
            double bb = funcl3(_x, m[j], sigma);
            xi[i + nxi * t] = alpha[0 + nstates * t] + lp + beta[j + nstates * (t + 1)] + bb; /* :240 */
        }
        double q = -INFINITY;
        for (int64_t e = 0; e < ntrans; e++) {
            int64_t i = tr[e].src - 1, j = tr[e].dst - 1;
            double bb = funcl3(_x, m[j], sigma);
            q = logsumexpl(q, alpha[i + nstates * t] + tr[e].lp + beta[j + nstates * (t + 1)] + bb); /* :248 */
        }
        for (int64_t i = 0; i < nxi; i++) xi[i + nxi * t] -= q;
    }
    double bb = -INFINITY; /* :254-261 */
    for (int64_t i = 0; i < nxi; i++) xx[i] = -INFINITY;
    for (int64_t t = 0; t < T - 1; t++) {
        bb = logsumexpl(bb, gf[0 + nstates * t]);
        for (int64_t j = 0; j < nxi; j++) xx[j] = logsumexpl(xx[j], xi[j + nxi * t]);
    }
    for (int64_t j = 0; j < nstates; j++) pp_out[j] = gf[j]; /* :263 */
    for (int64_t j = 1; j < nxi; j++) lp_out[j - 1] = xx[j] - bb; /* :264-265 xb[2:end] */

    /* :266-287 template update, in place */
    double *gg = calloc((size_t)(K * N), sizeof(double));
    int64_t *sidx = malloc(sizeof(int64_t) * nstates);
    if (!gg || !sidx) return ORC_ENOMEM;
    for (int64_t i = 0; i < K * N; i++) mu[i] = 0.0; /* :268 fill!(mu, 0.0) */
    int64_t ns1 = 0;
    for (int64_t j = 0; j < nstates; j++) { /* :269 */
        int c = 0;
        for (int64_t l = 0; l < N; l++) c += states1[l + N * j] >= 2;
        if (c == 1) sidx[ns1++] = j;
    }
    for (int64_t t = 0; t < T; t++) { /* :270-282 */
        double _x = x[t];
        for (int64_t s = 0; s < ns1; s++) {
            int64_t j = sidx[s];
            double eg = exp(gf[j + nstates * t]);
            for (int64_t l = 0; l < N; l++) {
                int ss = states1[l + N * j];
                if (ss > 1) {
                    mu[(ss - 1) + K * l] += _x * eg;
                    gg[(ss - 1) + K * l] += eg;
                }
            }
        }
    }
    for (int64_t l = 0; l < N; l++) /* :283-287 */
        for (int64_t j = 1; j < K; j++) mu[j + K * l] /= gg[j + K * l];
    state_means(states1, N, nstates, mu, K, m); /* :288-293 */
    double x2 = 0.0, qq = 0.0;                  /* :295-305 */
    for (int64_t t = 0; t < T; t++)
        for (int64_t j = 0; j < nstates; j++) {
            double _x = x[t];
            double eg = exp(gf[j + nstates * t]);
            double d = _x - m[j];
            x2 += d * d * eg;
            qq += eg;
        }
    double s2 = x2 / qq; /* :306 */
    *sigma_out = sqrt(s2);
    free(gg); free(sidx); free(tidx); free(xi); free(xx); free(m);
    if (!gamma_out) free(gf);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* one E/M step, src/baumwelch.jl:362-370                             */
/* ------------------------------------------------------------------ */
int orc_em_step(const double *X, int64_t T, const int16_t *states1, int64_t N, int64_t K, int64_t nstates,
                const orc_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                int64_t lp_cap, double *pp_out, double *loglik_out) {
    double *a = malloc(sizeof(double) * nstates * T);
    double *b = malloc(sizeof(double) * nstates * T);
    if (!a || !b) { free(a); free(b); return ORC_ENOMEM; }
    int rc = orc_forward(X, T, states1, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, a);
    if (!rc) rc = orc_backward(X, T, states1, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, b);
    if (!rc && loglik_out) *loglik_out = orc_loglik_from_alpha(a, T, nstates);
    double snew = 0.0;
    if (!rc)
        rc = orc_update(a, b, T, states1, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, X, lp_out, lp_cap, pp_out,
                        &snew, NULL);
    if (!rc) *sigma_inout = snew;
    free(a); free(b);
    return rc;
}

/* ------------------------------------------------------------------ */
/* reconstruct_signal, src/reconstruction.jl:1-9                      */
/* ------------------------------------------------------------------ */
int orc_reconstruct(const int16_t *x, int64_t T, const int16_t *states1, int64_t N, int64_t nstates, const double *mu,
                    int64_t K, double *Y) {
    for (int64_t i = 0; i < T; i++) {
        double s = 0.0; /* zeros(Float64) then += in j order */
        int64_t xi = x[i] - 1;
        if (xi < 0 || xi >= nstates) return ORC_EARG;
        for (int64_t j = 0; j < N; j++) s += mu[(states1[j + N * xi] - 1) + K * j];
        Y[i] = s;
    }
    return ORC_OK;
}

/* unroll_mlseq, src/extraction.jl:4-13 : out [N x T] */
int orc_unroll_mlseq(const int16_t *x, int64_t T, const int16_t *states1, int64_t N, int64_t nstates, int16_t *out) {
    for (int64_t i = 0; i < T; i++) {
        int64_t mi = x[i] - 1;
        if (mi < 0 || mi >= nstates) return ORC_EARG;
        for (int64_t j = 0; j < N; j++) out[j + N * i] = states1[j + N * mi];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* chunked decode, src/fit.jl:11-42 (with the dead gc() call dropped) */
/* ------------------------------------------------------------------ */
int orc_fit_chunked(const double *X, int64_t n, int64_t chunksize, const int16_t *states1, int64_t N, int64_t K,
                    int64_t nstates, const orc_trans *tr, int64_t ntrans, const double *mu, double sigma,
                    int16_t *ml_seq, double *ll_out) {
    int64_t i = 1, j = 1;
    double ll = 0.0;
    for (int64_t t = 0; t < n; t++) ml_seq[t] = 1;
    int16_t *x = malloc(sizeof(int16_t) * (size_t)(chunksize > n ? n : chunksize));
    if (!x) return ORC_ENOMEM;
    while (j < n) {
        j = (i + chunksize - 1 < n) ? i + chunksize - 1 : n;
        int64_t k = j - i + 1, l = 1;
        double _ll;
        int rc = orc_viterbi(X + (i - 1), k, states1, N, K, nstates, tr, ntrans, mu, sigma, x, &_ll, NULL, NULL);
        if (rc) { free(x); return rc; }
        if (i > 1)
            while (l <= k && x[l - 1] > 1) l++;
        if (j < n)
            while (k >= 1 && x[k - 1] > 1) { j--; k--; }
        for (int64_t q = l; q <= k; q++) ml_seq[(i + l - 1) + (q - l) - 1] = x[q - 1];
        ll += _ll;
        if (j <= i) break; /* the reference would loop forever here */
        i = j;
    }
    *ll_out = ll;
    free(x);
    return ORC_OK;
}
