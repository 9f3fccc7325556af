// engines.h -- internal interfaces between the C-ABI layer (api.cu) and the
// kernel translation units.
#pragma once
#include <cstring>

#include "common.h"

namespace hmm {

// ---- faithful engine (faithful.cu) ----------------------------------------
struct FaithfulLayout {  // byte offsets inside one channel's device model blob
    size_t scal, m, in_lp, out_lp, in_ptr, in_src, out_ptr, out_dst, dec_slot, static_pred, bytes;
};
FaithfulLayout faithful_layout(int nstates, int64_t ntrans);
void faithful_pack(const HostModel &M, const FaithfulLayout &L, char *dst);
size_t faithful_smem_bytes(int ns, int64_t nt);

// Decode C channels; y_dev [T x C] with column stride y_stride.  When
// forward_only, stops after the forward sweep (used for the ring prologue).
void faithful_viterbi_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const FaithfulLayout &L,
                          const char *blob_dev, const HostModel &M0, int16_t *x_dev, int64_t x_stride, double *ll_dev,
                          double *T1_dev, int16_t *T2_dev, int64_t trellis_cols, bool forward_only,
                          double *final_col_dev, cudaStream_t st, hmm_info *info);
void path_score_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const FaithfulLayout &L,
                    const char *blob_dev, const HostModel &M0, const int16_t *x_dev, int64_t x_stride, double *ll_dev,
                    cudaStream_t st, hmm_info *info);
void faithful_fb_run(bool backward, const double *V_dev, int64_t T, const FaithfulLayout &L, const char *blob_dev,
                     const HostModel &M, double *out_dev, cudaStream_t st);

// ---- time-parallel per-state engine for any StateMatrix (generic_parallel.cu) ----
bool generic_parallel_supported(const HostModel &M, int64_t T);
bool generic_parallel_preferred(const HostModel &M, int64_t T);  // auto mode: long, or too large for the sequential engine
void generic_parallel_viterbi_run(const double *y_dev, int64_t T, const FaithfulLayout &L, const char *blob_dev,
                                  const HostModel &M0, int16_t *x_dev, double *ll_host, cudaStream_t st, hmm_info *info);

// ---- ring engine, Viterbi (ring_viterbi.cu) --------------------------------
struct RingConfig {
    int64_t chunk_len = 0;  // 0 = auto
    int64_t warmup = 0;     // 0 = default
    int profile = 0;        // hmm_set_profiling: eager launches with per-kernel event timers instead of a CUDA graph
    int precision = 0;      // 0: FP64 throughout; 1: FP32 mode (the FIR of the ring decode in FP32; hmm_set_precision)
};
RingConfig &ring_config();
bool ring_supported(const HostModel &M, int64_t T);
// Results of a decode that has been launched but not yet synchronised with: ll and the repair counts sit in
// mapped pinned memory (res_h: [C x 4] doubles) and are handed to the caller by ring_collect after the
// stream has been synchronised.
struct RingPending {
    const double *res_h;
    int C;
    double *ll_host;  // nullable
    hmm_info *info;   // nullable
};
void ring_collect(std::vector<RingPending> &pend);
void ring_new_epoch();      // called once per C-ABI entry: programs used by the current call are never evicted
void ring_drop_programs();  // frees every cached decode program (hmm_release_workspace)
// models[C] must share topology (N, K).  Leaves x in x_dev; ll (nullable, HOST pointer) via path scoring.
// model_id != 0 identifies (models, blob_dev) for the program cache: a repeated decode of the same buffers with the
// same model re-launches one CUDA graph.  Without `defer` the call synchronises the stream once and delivers ll /
// info; with it, the caller synchronises and calls ring_collect (many channels per call).
void ring_viterbi_run(const double *y_dev, int64_t T, int64_t y_stride, int C, const std::vector<HostModel> &models,
                      const FaithfulLayout &L, const char *blob_dev, uint64_t model_id, int16_t *x_dev,
                      int64_t x_stride, double *ll_host, cudaStream_t st, hmm_info *info,
                      std::vector<RingPending> *defer);

// Staged decode (used directly by the time-shard API in api.cu).
struct VitParams;
class VitPlan {
  public:
    VitPlan();
    ~VitPlan();
    VitPlan(const VitPlan &) = delete;
    VitPlan &operator=(const VitPlan &) = delete;
    bool own_memory = false;  // false: thread-local grow-only workspace; true: cudaMalloc owned by the plan
    void build(const double *y_dev, int64_t T, int64_t y_stride, int C, const std::vector<HostModel> &models,
               const FaithfulLayout &FL, const char *blob_dev, int16_t *x_dev, int64_t x_stride, int64_t Lc, int64_t W,
               bool first_prologue, bool last_true_end, cudaStream_t st);
    void forward(cudaStream_t st, Timer *ttop);
    void verify_fwd(cudaStream_t st);    // boundary check + repair (+ final state when the sequence really ends)
    void trace(cudaStream_t st);
    void verify_trace(cudaStream_t st);  // boundary check + repair
    void path_ll(cudaStream_t st, double *ll_dev, int64_t t_lo, int64_t t_hi, int64_t t_off, int64_t T_glob, bool with_p0);
    // local sample range [lo, hi) the path score covers (chunk aligned), global time offset and length, whether the
    // t = 0 term belongs to it; verify_trace leaves the score in ll_dev()
    void set_ll_range(int64_t lo, int64_t hi, int64_t t_off, int64_t T_glob, bool with_p0, bool want = true);
    void run_all(cudaStream_t st, bool want_ll, Timer *ttop);  // the whole decode: six launches
    void set_result_sink(double *res_dev_alias);  // [C x 4] mapped pinned memory the kernels leave ll / repair counts in
    void retarget(const double *y_dev, int16_t *x_dev);
    double *ll_dev() { return ll_dev_; }
    void check_guards(cudaStream_t st);  // HMMCUDA_DEBUG_GUARD: fails if any kernel wrote outside a plan buffer
    void read_counters(cudaStream_t st, int *fwd_rep, int *bwd_rep);
    void reset_counters(cudaStream_t st);
    int nchunks() const;
    int bvec() const;
    double *eb_ptr(int chunk);              // device pointer: true end-boundary vector of `chunk` (channel 0)
    long long *own_start_ptr(int chunk);    // device pointer: traceback state at the start of `chunk`
    void *alloc(int slot, size_t bytes);
    void use_arena(char *base, size_t cap);  // allocate everything out of this device buffer
    size_t arena_bytes_used() const;
    void set_x_window(int64_t lo, int64_t hi);  // local steps whose x may be written
    int *counters_ptr();
    double *sb_ptr(int chunk);  // speculative start vector of a chunk

  private:
    struct Impl;
    Impl *impl;
    VitParams *p_;  // owned; defined in ring_viterbi.cu
    std::vector<void *> owned;
    std::vector<char *> guards;  // HMMCUDA_DEBUG_GUARD: guard zones around every owned buffer
    std::vector<double> hmdl;
    HostModel M0;
    FaithfulLayout FL;
    const char *blob_dev = nullptr;
    double *part = nullptr, *ll_dev_ = nullptr;
    int C = 1;
    bool per_channel = false;  // long recordings: every channel is launched on its own (see run_all)
    int launch_C() const;
    char *arena_base = nullptr;
    size_t arena_cap = 0, arena_used = 0;
};
void vshard_judge_run(const double *gathered_dev, int n_ranks, int bvec, double *out_dev, cudaStream_t st);
// peer-memory exchange of the shard summaries (see ring_viterbi.cu)
size_t vshard_exchange_block_bytes(int world, int bvec);
void vshard_exchange_run(const double *eb_last, const double *sb_first, const long long *own_first,
                         const long long *own_ghost, const double *ll, long long shift, int bvec, char *const *peers_dev,
                         int rank, int world, const unsigned long long *epoch_dev, cudaStream_t st);
void vshard_judge_p2p_run(char *own_block, int world, int bvec, unsigned long long *epoch_dev, double *out_mapped,
                          cudaStream_t st);
void ring_path_ll_run(const double *y_dev, int64_t T, const FaithfulLayout &FL, const char *blob_dev, const HostModel &M0,
                      const int16_t *x_dev, double *ll_dev, double *part, cudaStream_t st);
// Default chunk length / warm-up for a recording of T_total samples decoded on n_gpus GPUs.
int ring_default_chunking(const HostModel &M0, int64_t T_total, int C, int n_gpus, int64_t *Lc_out, int64_t *W_out);

// ---- ring engine, E/M step (ring_em.cu) ------------------------------------
struct EmResult {
    std::vector<double> lp;  // [N]
    std::vector<double> pp;  // [nstates] = gamma[:,1] (log)
    std::vector<double> mu;  // [K x N]
    double sigma = 0, loglik = 0;
};
void ring_em_run(const double *X_dev, int64_t T, const HostModel &M, EmResult &out, cudaStream_t st, hmm_info *info);
// Time-sharded E/M step: the E-step of one shard (local samples incl. one ghost chunk per side; statistics over the
// local steps [st_lo, st_hi) only) leaves its statistics vector (ring_em_xvec_len doubles) and its four boundary
// vectors + local lS (ring_em_bnd_len doubles) in device buffers, asynchronously; the M-step consumes the summed vector.
struct EmShardOpts {
    int64_t Lc, W, st_lo, st_hi;
    bool first, last;
    double *xvec_dev, *bnd_dev;
};
void ring_em_default_chunking(int N, int K, int64_t T_local, int64_t *Lc_out, int64_t *W_out);
int ring_em_xvec_len(int N, int nstates);
int ring_em_bnd_len(int N, int K);
void ring_em_shard_estep(const double *X_dev, int64_t T_local, const HostModel &M, const EmShardOpts &sh, cudaStream_t st,
                         hmm_info *info);
void ring_em_shard_mstep(const double *xsum_dev, const HostModel &M, int64_t T_glob, double lS_glob, EmResult &out,
                         cudaStream_t st);
// dense alpha and/or beta [nstates x T] (device pointers, nullable) from the semi-Markov engine
void ring_fb_dense_run(const double *X_dev, int64_t T, const HostModel &M, double *alpha_dev, double *beta_dev,
                       cudaStream_t st);

// ---- generic E/M pieces (update.cu) ----------------------------------------
// update() on dense alpha/beta (device pointers), src/baumwelch.jl:205-309
void dense_update_run(const double *alpha_dev, const double *beta_dev, const double *x_dev, int64_t T,
                      const HostModel &M, const int16_t *states_host, EmResult &out, cudaStream_t st);

// roofline denominators measured live (reconstruct.cu)
void measure_peaks(double *gdfma_per_s, double *copy_gb_per_s, cudaStream_t st);

// ---- reconstruct (reconstruct.cu) ------------------------------------------
void reconstruct_run(const int16_t *x_dev, int64_t T, const std::vector<double> &m, double *Y_dev, cudaStream_t st);
void unroll_run(const int16_t *x_dev, int64_t T, const int16_t *states_host, int N, int nstates, int16_t *out_dev,
                cudaStream_t st);

}  // namespace hmm
