// generic_parallel.cu -- time-parallel exact Viterbi decode for ANY StateMatrix (overlap models, N > 7, K > 97):
// the models the ring engine cannot take, among them the reference's own `Viterbi` testset and CLI model
// (allow_overlaps = true, 3 600 states; test/runtests.jl:24, src/hmmsort.jl:54).
//
// Same idea as the ring engine, on the plain per-state recursion of src/viterbi.jl:65-88 (CSR by destination, candidates
// in list order, strict >): the recording is cut into chunks, ONE CTA per chunk, one thread per state.  Chunks other than
// the first start SPECULATIVELY from a flat score vector W steps early; afterwards every boundary is VERIFIED -- the
// speculative column at the chunk start must equal the previous chunk's true end column up to an additive constant --
// and a chunk that fails is re-run from the true column by a sequential repair pass.  The traceback is parallel the same
// way (speculative look-ahead from the noise state, boundary states compared, repair right to left).  The first chunk
// starts from the reference's initialisation (:55-63), so its scores -- and the structural exact ties at the second
// sample -- are the reference's bit for bit; a chunk re-run from a true column continues that arithmetic.
// ll is computed from (x, y) by ring_path_ll (its CSR form is generic), i.e. to 1e-9 relative, not bit-exact: callers
// that need the reference's own rounding of ll use HMM_MODE_FAITHFUL.
#include <algorithm>
#include <cmath>

#include "engines.h"
#include "faithful_dev.cuh"

namespace hmm {

namespace {

struct GenParams {
    const double *y;
    int64_t T, Lc, W;
    int ns, nt, ndec, nchunks;
    int16_t *decbp;        // [T x ndec] backpointers of the multi-predecessor states (0-based source)
    double *SB, *EB;       // [nchunks x ns] speculative start column (at s-1) / true end column (at e-1)
    int *flag;             // [nchunks] forward boundary mismatch
    int16_t *xlast;        // argmax of the last column
    int *own_start, *look_end, *tflag;  // traceback boundary states
    int *counters;         // [0] forward chunks repaired [1] traceback chunks repaired
    int16_t *x;
    double *res_host;      // mapped pinned: [1] fwd repaired [2] traceback repaired (nullable)
    int dbg_flag_every;
    // Models of many thousand states (the CLI's overlap models: 10 621 states at N=3, 21 123 at N=4, K=60) do not
    // fit shared memory: `place` says which arrays live there (bit set) and which are read from the model blob /
    // a per-CTA global scratch (L2-resident) instead.
    int place;             // GP_COLS | GP_M | GP_DEC | GP_PTR | GP_LP | GP_IDX
    double *colg;          // [gridDim.x x 2 x ns] score columns when GP_COLS is clear
};
enum { GP_COLS = 1, GP_M = 2, GP_DEC = 4, GP_PTR = 8, GP_LP = 16, GP_IDX = 32, GP_ALL = 63 };

struct GenSmem {  // generic pointers: shared memory or global, per `place`
    double *col0, *col1, *ytile;
    const double *m, *lp;
    const int *ptr, *idx, *dec;
};
constexpr int GYT = 256;

__host__ __device__ inline size_t gen_smem_bytes(int ns, int64_t nt, int place) {
    size_t d = GYT, i = 0;
    if (place & GP_COLS) d += 2 * (size_t)ns;
    if (place & GP_M) d += ns;
    if (place & GP_LP) d += nt;
    if (place & GP_DEC) i += ns;
    if (place & GP_PTR) i += (size_t)ns + 1;
    if (place & GP_IDX) i += nt;
    return sizeof(double) * d + sizeof(int) * i + 16;
}

// Carves the CTA's shared memory and copies the arrays placed there; the others point into the model blob.
__device__ __forceinline__ GenSmem gen_setup(const GenParams &p, char *base, const char *mb, const FaithfulLayout &L) {
    const int ns = p.ns, nt = p.nt, place = p.place;
    const double *gm = (const double *)(mb + L.m), *glp = (const double *)(mb + L.in_lp);
    const int *gp = (const int *)(mb + L.in_ptr), *gs = (const int *)(mb + L.in_src), *gd = (const int *)(mb + L.dec_slot);
    GenSmem s;
    double *d = (double *)base;
    s.ytile = d; d += GYT;
    if (place & GP_COLS) {
        s.col0 = d; d += ns;
        s.col1 = d; d += ns;
    } else {
        s.col0 = p.colg + (size_t)blockIdx.x * 2 * ns;
        s.col1 = s.col0 + ns;
    }
    double *sm_m = nullptr, *sm_lp = nullptr;
    if (place & GP_M) { sm_m = d; d += ns; }
    if (place & GP_LP) { sm_lp = d; d += nt; }
    int *i = (int *)d, *sm_dec = nullptr, *sm_ptr = nullptr, *sm_idx = nullptr;
    if (place & GP_DEC) { sm_dec = i; i += ns; }
    if (place & GP_PTR) { sm_ptr = i; i += ns + 1; }
    if (place & GP_IDX) { sm_idx = i; i += nt; }
    for (int k = threadIdx.x; k < ns; k += blockDim.x) {
        if (sm_m) sm_m[k] = gm[k];
        if (sm_dec) sm_dec[k] = gd[k];
    }
    if (sm_ptr)
        for (int k = threadIdx.x; k <= ns; k += blockDim.x) sm_ptr[k] = gp[k];
    for (int k = threadIdx.x; k < nt; k += blockDim.x) {
        if (sm_lp) sm_lp[k] = glp[k];
        if (sm_idx) sm_idx[k] = gs[k];
    }
    s.m = sm_m ? sm_m : gm;
    s.lp = sm_lp ? sm_lp : glp;
    s.dec = sm_dec ? sm_dec : gd;
    s.ptr = sm_ptr ? sm_ptr : gp;
    s.idx = sm_idx ? sm_idx : gs;
    __syncthreads();
    return s;
}

enum { GEN_SPEC = 0, GEN_EXACT = 1 };

// One chunk, the whole CTA.  Returns with the chunk's decisions, SB (speculative runs) and EB written.
__device__ void gen_run_chunk(const GenParams &p, const GenSmem &S, double c_emit, double two_s2, int c, int kind) {
    const int ns = p.ns;
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    if (c == p.nchunks - 1 || e > T) e = T;
    double *prev = S.col0, *cur = S.col1;
    int64_t t_first;
    const bool true_start = c == 0 || (kind == GEN_SPEC && s - p.W < 1);
    if (true_start) {  // column 1 of src/viterbi.jl:55-63: emissions, noise forced to 0
        const double y0 = p.y[0];
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            double v = emit_rn(y0, S.m[j], c_emit, two_s2);
            if (j == 0) v = 0.0;
            prev[j] = v;
        }
        t_first = 1;
    } else if (kind == GEN_EXACT) {
        const double *eb = p.EB + (size_t)(c - 1) * ns;
        double *sb = p.SB + (size_t)c * ns;
        for (int j = threadIdx.x; j < ns; j += blockDim.x) {
            const double v = eb[j];
            prev[j] = v;
            sb[j] = v;  // the start column on record is the one this run was really started from
        }
        t_first = s;
    } else {  // speculative: a flat column W steps before the chunk
        for (int j = threadIdx.x; j < ns; j += blockDim.x) prev[j] = 0.0;
        t_first = s - p.W;
    }
    __syncthreads();
    if (c > 0 && true_start && s - 1 == 0) {  // (W reaches back to the very first column: it is the start column)
        double *sb = p.SB + (size_t)c * ns;
        for (int j = threadIdx.x; j < ns; j += blockDim.x) sb[j] = prev[j];
    }
    for (int64_t t0 = t_first; t0 < e; t0 += GYT) {
        const int n = (int)((e - t0 < GYT) ? (e - t0) : GYT);
        for (int k = threadIdx.x; k < n; k += blockDim.x) S.ytile[k] = p.y[t0 + k];
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const int64_t t = t0 + k;
            const double yv = S.ytile[k];
            const bool keep = t >= s;  // warm-up decisions belong to the previous chunk
            for (int j = threadIdx.x; j < ns; j += blockDim.x) {
                const double q = emit_rn(yv, S.m[j], c_emit, two_s2);
                double best = -INFINITY;
                int bp = 0;
                const int e1 = S.ptr[j + 1];
                for (int ed = S.ptr[j]; ed < e1; ed++) {
                    const int k2 = S.idx[ed];
                    const double tt = __dadd_rn(prev[k2], S.lp[ed]);
                    if (tt > best) {  // strict: first candidate in list order wins ties
                        best = tt;
                        bp = k2;
                    }
                }
                const double v = __dadd_rn(best, q);
                cur[j] = v;
                const int slot = S.dec[j];
                if (keep && slot >= 0) p.decbp[(size_t)t * p.ndec + slot] = (int16_t)bp;
                if (kind == GEN_SPEC && c > 0 && t == s - 1) p.SB[(size_t)c * ns + j] = v;
                if (t == e - 1) p.EB[(size_t)c * ns + j] = v;
            }
            __syncthreads();
            double *tmp = prev;
            prev = cur;
            cur = tmp;
        }
    }
    if (c == p.nchunks - 1 && threadIdx.x == 0) {  // x[T] = argmax(T1[:,T]), first maximum (src/viterbi.jl:90)
        int best = 0;
        double bv = prev[0];
        for (int j = 1; j < ns; j++)
            if (prev[j] > bv) {
                bv = prev[j];
                best = j;
            }
        p.xlast[0] = (int16_t)best;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) gen_vit_forward(GenParams p, const char *blob, FaithfulLayout L) {
    extern __shared__ __align__(16) char smem_raw[];
    const double *sc = (const double *)(blob + L.scal);
    GenSmem S = gen_setup(p, smem_raw, blob, L);
    gen_run_chunk(p, S, sc[2], sc[3], blockIdx.x, GEN_SPEC);
}

// two columns describe the same scores iff they differ by a constant (to a few dozen ulp of their magnitude)
__device__ bool gen_columns_match(const double *sb, const double *eb, int ns) {
    bool bad = false;
    const double d0 = sb[0] - eb[0];
    for (int j = threadIdx.x; j < ns; j += blockDim.x) {
        const double a = sb[j], b = eb[j];
        const bool ia = isinf(a), ib = isinf(b);
        if (ia || ib) {
            if (ia != ib) bad = true;
            continue;
        }
        const double tol = 1e-12 + 64 * 2.220446049250313e-16 * fmax(fabs(a), fabs(b));
        if (!(fabs((a - b) - d0) <= tol)) bad = true;
    }
    return !__syncthreads_or(bad ? 1 : 0);
}

// Verification of the forward pass in one launch: CTA c checks boundary c | c-1; the last CTA to finish repairs.
__global__ void __launch_bounds__(1024) gen_vit_verify_fwd(GenParams p, const char *blob, FaithfulLayout L, unsigned *arrive) {
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ int s_last;
    const int c = blockIdx.x;
    if (c >= 1) {
        bool ok = gen_columns_match(p.SB + (size_t)c * p.ns, p.EB + (size_t)(c - 1) * p.ns, p.ns);
        if (p.dbg_flag_every > 0 && c % p.dbg_flag_every == 0) ok = false;
        if (threadIdx.x == 0) p.flag[c] = ok ? 0 : 1;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(arrive, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *arrive = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int any = 0;
    for (int k = 1 + threadIdx.x; k < p.nchunks; k += blockDim.x) any |= __ldcg(p.flag + k);
    any = __syncthreads_or(any);
    int repaired = 0;
    if (any) {
        const double *sc = (const double *)(blob + L.scal);
        GenSmem S = gen_setup(p, smem_raw, blob, L);
        bool prev_rerun = false;
        for (int k = 1; k < p.nchunks; k++) {
            bool need = __ldcg(p.flag + k) != 0;
            if (!need && prev_rerun) need = !gen_columns_match(p.SB + (size_t)k * p.ns, p.EB + (size_t)(k - 1) * p.ns, p.ns);
            if (need) {
                gen_run_chunk(p, S, sc[2], sc[3], k, GEN_EXACT);
                __threadfence();
                repaired++;
            }
            prev_rerun = need;
        }
    }
    if (threadIdx.x == 0) {
        p.counters[0] = repaired;
        if (p.res_host) p.res_host[1] = (double)repaired;
    }
}

// Traceback of one chunk: from time t_hi (state `cur`) down to the chunk start s; writes x[s .. min(e, t_hi+1)).
__device__ void gen_trace_chunk(const GenParams &p, int c, int64_t t_hi, int cur, bool record_look, int16_t *bpt, int16_t *xt,
                                const int16_t *dec, const int16_t *sp, int tile) {
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    if (c == p.nchunks - 1 || e > T) e = T;
    __shared__ int cur_s, look_s;
    if (threadIdx.x == 0) {
        cur_s = cur;
        look_s = -2;
        if (t_hi < e) p.x[t_hi] = (int16_t)(cur + 1);
        if (t_hi == e) look_s = cur;
    }
    __syncthreads();
    if (tile == 0) {
        // models with many multi-predecessor states (rows of decbp too long to stage): one thread follows the
        // backpointers straight from global memory -- inside a chain the predecessor is static, no load
        if (threadIdx.x == 0) {
            int cs = cur_s;
            for (int64_t t = t_hi; t > s; t--) {  // step t -> state at t - 1
                const int slot = dec[cs];
                cs = slot >= 0 ? p.decbp[(size_t)t * p.ndec + slot] : sp[cs];
                if (t - 1 < e) p.x[t - 1] = (int16_t)(cs + 1);
                if (t - 1 == e) look_s = cs;
            }
            cur_s = cs;
        }
        __syncthreads();
    }
    // steps t in (s, t_hi], processed in tiles [a, b]: the backpointer of step t gives the state at t - 1
    for (int64_t b = t_hi; tile > 0 && b > s; b -= tile) {
        int64_t a = b - tile + 1;
        if (a < s + 1) a = s + 1;
        const int n = (int)(b - a + 1);
        for (size_t i = threadIdx.x; i < (size_t)n * p.ndec; i += blockDim.x) bpt[i] = p.decbp[(size_t)a * p.ndec + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int cs = cur_s;
            for (int k = n - 1; k >= 0; k--) {  // step t = a + k  ->  state at t - 1
                const int slot = dec[cs];
                cs = slot >= 0 ? bpt[(size_t)k * p.ndec + slot] : sp[cs];
                xt[k] = (int16_t)(cs + 1);
                if (a + k - 1 == e) look_s = cs;
            }
            cur_s = cs;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int64_t tx = a + k - 1;
            if (tx < e) p.x[tx] = xt[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.own_start[c] = cur_s;  // state at time s
        if (record_look) p.look_end[c] = look_s;
    }
    __syncthreads();
}

__device__ void gen_trace_tables(const char *blob, const FaithfulLayout &L, int ns, int16_t *dec, int16_t *sp) {
    const int *gd = (const int *)(blob + L.dec_slot), *gsp = (const int *)(blob + L.static_pred);
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        dec[i] = (int16_t)gd[i];
        sp[i] = (int16_t)gsp[i];
    }
    __syncthreads();
}

__global__ void gen_vit_trace(GenParams p, const char *blob, FaithfulLayout L, int tile) {
    extern __shared__ __align__(16) char smem_raw[];
    int16_t *bpt = (int16_t *)smem_raw, *xt = bpt + (size_t)tile * p.ndec, *dec = xt + tile, *sp = dec + p.ns;
    gen_trace_tables(blob, L, p.ns, dec, sp);  // (dec_slot < ndec <= nstates <= 32767: both tables fit int16)
    const int c = blockIdx.x;
    const int64_t T = p.T;
    const int64_t s = (int64_t)c * p.Lc;
    int64_t e = s + p.Lc;
    const bool last = c == p.nchunks - 1;
    if (last || e > T) e = T;
    int64_t t_hi = e + p.W;
    int cur = 0;  // speculative: the noise state W steps past the chunk
    if (last || t_hi >= T - 1) {
        t_hi = T - 1;
        cur = p.xlast[0];
    }
    gen_trace_chunk(p, c, t_hi, cur, !last, bpt, xt, dec, sp, tile);
}

__global__ void gen_vit_verify_trace(GenParams p, const char *blob, FaithfulLayout L, int tile, unsigned *arrive) {
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ int s_last;
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c < p.nchunks - 1) {
            bool bad = p.look_end[c] != p.own_start[c + 1];
            if (p.dbg_flag_every > 0 && c % p.dbg_flag_every == 0) bad = true;
            p.tflag[c] = bad ? 1 : 0;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(arrive + 1, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) arrive[1] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int any = 0;
    for (int k = threadIdx.x; k < p.nchunks - 1; k += blockDim.x) any |= __ldcg(p.tflag + k);
    any = __syncthreads_or(any);
    int repaired = 0;
    if (any) {
        int16_t *bpt = (int16_t *)smem_raw, *xt = bpt + (size_t)tile * p.ndec, *dec = xt + tile, *sp = dec + p.ns;
        gen_trace_tables(blob, L, p.ns, dec, sp);
        bool next_changed = false;
        for (int c = p.nchunks - 2; c >= 0; c--) {
            bool need = __ldcg(p.tflag + c) != 0;
            if (!need && next_changed) need = __ldcg(p.look_end + c) != __ldcg(p.own_start + c + 1);
            if (need) {
                const int before = __ldcg(p.own_start + c);
                const int64_t e = (int64_t)(c + 1) * p.Lc;
                gen_trace_chunk(p, c, e, __ldcg(p.own_start + c + 1), false, bpt, xt, dec, sp, tile);
                __threadfence();
                __syncthreads();
                next_changed = __ldcg(p.own_start + c) != before;
                repaired++;
            } else
                next_changed = false;
        }
    }
    if (threadIdx.x == 0) {
        p.counters[1] = repaired;
        if (p.res_host) p.res_host[2] = (double)repaired;
    }
}

int gen_threads(int ns) {
    int t = 32;
    while (t < ns && t < 1024) t <<= 1;
    return t;
}

}  // namespace

// Which arrays go to shared memory: in order of how often a step touches them, as long as they fit.
static int gen_placement(int ns, int64_t nt, size_t budget) {
    int place = 0;
    const int order[] = {GP_COLS, GP_M, GP_DEC, GP_PTR, GP_LP, GP_IDX};
    for (int bit : order)
        if (gen_smem_bytes(ns, nt, place | bit) <= budget) place |= bit;
    return place;
}

bool generic_parallel_supported(const HostModel &M, int64_t T) {
    (void)T;  // any length: a short sequence is simply one chunk
    return M.nstates <= 32767;
}
// models the sequential per-state engine cannot hold in shared memory (> ~5 000 states) go here at any length
bool generic_parallel_preferred(const HostModel &M, int64_t T) {
    return M.nstates <= 32767 && (T >= 4096 || faithful_smem_bytes(M.nstates, M.ntrans) > 227 * 1024);
}

void generic_parallel_viterbi_run(const double *y_dev, int64_t T, const FaithfulLayout &L, const char *blob_dev,
                                  const HostModel &M0, int16_t *x_dev, double *ll_host, cudaStream_t st, hmm_info *info) {
    Workspace &ws = workspace();
    const int ns = M0.nstates, nt = (int)M0.ntrans, ndec = M0.ndec > 0 ? M0.ndec : 1;
    // HMMCUDA_DEBUG_GEN_SMEM_KB: a smaller shared-memory budget (tests exercise every placement on small models)
    const char *dbg_kb = getenv("HMMCUDA_DEBUG_GEN_SMEM_KB");
    const size_t budget = dbg_kb && atoi(dbg_kb) > 0 ? std::min<size_t>(220, (size_t)atoi(dbg_kb)) * 1024 : 220 * 1024;
    const int place = gen_placement(ns, nt, budget);
    const size_t sm = gen_smem_bytes(ns, nt, place);
    const int nth = gen_threads(ns);
    HMM_CUDA(cudaFuncSetAttribute(gen_vit_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    HMM_CUDA(cudaFuncSetAttribute(gen_vit_verify_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int dev = 0, sms = 148, occ = 1;
    HMM_CUDA(cudaGetDevice(&dev));
    HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gen_vit_forward, nth, sm));
    if (occ < 1) occ = 1;
    // speculative warm-up / look-ahead: a few chain lengths (paths coalesce once every survivor has passed a stretch
    // of noise); every boundary is verified, so this only trades repairs against redundant steps
    int64_t W = ring_config().warmup > 0 ? ring_config().warmup : std::max<int64_t>(512, 4 * (int64_t)M0.K);
    int64_t Lc = ring_config().chunk_len;
    if (Lc <= 0) {
        Lc = (T + (int64_t)sms * occ - 1) / ((int64_t)sms * occ);  // one wave of chunks
        if (Lc < 2 * W) Lc = 2 * W;
    }
    if (Lc < 64) Lc = 64;
    int nchunks = (int)((T + Lc - 1) / Lc);
    if (nchunks > 1 && T - (int64_t)(nchunks - 1) * Lc < 2) nchunks--;
    if (nchunks < 1) nchunks = 1;
    // traceback: rows of decbp staged in tiles when they are short enough, else followed straight from global memory
    int tile = 2048;
    while (tile >= 64 && (size_t)tile * (ndec + 1) * 2 + (size_t)ns * 4 > 96 * 1024) tile /= 2;
    if (tile < 64 || (getenv("HMMCUDA_DEBUG_GEN_DIRECT_TRACE") && atoi(getenv("HMMCUDA_DEBUG_GEN_DIRECT_TRACE")))) tile = 0;
    const size_t sm_tr = (size_t)tile * (ndec + 1) * 2 + (size_t)ns * 4 + 16;
    HMM_CUDA(cudaFuncSetAttribute(gen_vit_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tr));
    HMM_CUDA(cudaFuncSetAttribute(gen_vit_verify_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tr));

    GenParams p{};
    p.y = y_dev; p.T = T; p.Lc = Lc; p.W = W; p.ns = ns; p.nt = nt; p.ndec = ndec; p.nchunks = nchunks;
    p.decbp = (int16_t *)ws.get(Workspace::DEC, sizeof(int16_t) * (size_t)T * ndec);
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t r = off;
        off += (bytes + 255) & ~size_t(255);
        return r;
    };
    const size_t o_sb = carve(sizeof(double) * (size_t)nchunks * ns), o_eb = carve(sizeof(double) * (size_t)nchunks * ns);
    const size_t o_flag = carve(sizeof(int) * (size_t)nchunks * 4), o_cnt = carve(sizeof(int) * 8);
    const size_t o_arr = carve(sizeof(unsigned) * 8), o_xl = carve(64), o_part = carve(sizeof(double) * 1024);
    const size_t o_col = carve((place & GP_COLS) ? 0 : sizeof(double) * 2 * (size_t)ns * nchunks);
    char *base = (char *)ws.get(Workspace::CHUNKS, off);
    p.place = place;
    p.colg = (double *)(base + o_col);
    p.SB = (double *)(base + o_sb);
    p.EB = (double *)(base + o_eb);
    p.flag = (int *)(base + o_flag);
    p.own_start = p.flag + nchunks;
    p.look_end = p.own_start + nchunks;
    p.tflag = p.look_end + nchunks;
    p.counters = (int *)(base + o_cnt);
    unsigned *arrive = (unsigned *)(base + o_arr);
    p.xlast = (int16_t *)(base + o_xl);
    p.x = x_dev;
    p.dbg_flag_every = getenv("HMMCUDA_DEBUG_FLAG_EVERY") ? atoi(getenv("HMMCUDA_DEBUG_FLAG_EVERY")) : 0;
    void *res_dev = nullptr;
    double *res_h = (double *)ws.pinned(3, sizeof(double) * 4, &res_dev);
    res_h[0] = res_h[1] = res_h[2] = 0.0;
    p.res_host = (double *)res_dev;
    HMM_CUDA(cudaMemsetAsync(arrive, 0, sizeof(unsigned) * 8, st));
    HMM_CUDA(cudaMemsetAsync(p.counters, 0, sizeof(int) * 8, st));
    {
        NvtxRange r("hmm.viterbi.generic_parallel");
        gen_vit_forward<<<nchunks, nth, sm, st>>>(p, blob_dev, L);
        gen_vit_verify_fwd<<<nchunks, nth, sm, st>>>(p, blob_dev, L, arrive);
        gen_vit_trace<<<nchunks, 256, sm_tr, st>>>(p, blob_dev, L, tile);
        gen_vit_verify_trace<<<(nchunks + 255) / 256, 256, sm_tr, st>>>(p, blob_dev, L, tile, arrive);
        HMM_CUDA(cudaGetLastError());
    }
    double *ll_dev = (double *)(base + o_part);
    if (ll_host) ring_path_ll_run(y_dev, T, L, blob_dev, M0, x_dev, ll_dev, ll_dev + 8, st);
    if (ll_host) HMM_CUDA(cudaMemcpyAsync(ll_host, ll_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
    HMM_CUDA(cudaStreamSynchronize(st));
    if (info) {
        info->kernel_launches += ll_host ? 5 : 4;
        info->n_chunks = nchunks;
        info->fwd_repaired += (int)res_h[1];
        info->bwd_repaired += (int)res_h[2];
    }
}

}  // namespace hmm
