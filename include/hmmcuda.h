/*
 * hmmcuda.h -- C ABI of libhmmcuda.so: the B200-native HMM inference hot path
 * of grero/HMMSpikeSorter.jl (Viterbi decode, Baum-Welch E/M step,
 * reconstruct_signal) behind the reference's own function signatures.
 *
 * The reference is pure Julia and has no FFI seam; these entry points are
 * what a `ccall` shim replacing the Julia methods binds.
 * Every array is caller-owned, column-major, with 1-based state indices,
 * exactly as the Julia objects lay them out:
 *
 *   states  Int16  [N x nstates]   StateMatrix.states       (src/types.jl:2,150)
 *   tr      24-byte records {Int64 src, Int64 dst, Float64 lp}, sorted by
 *           (src, dst)             StateMatrix.transitions  (src/types.jl:3,115-127)
 *   mu      Float64 [K x N]        templates, row 1 = silent (src/types.jl:17)
 *   N = neurons, K = states per ring incl. the silent one   (src/types.jl:150)
 *
 * All functions return 0 on success, else an HMM_E* code; hmm_last_error()
 * returns a thread-local message.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with HMM_ENODEV.
 */
#ifndef HMMCUDA_H
#define HMMCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMM_OK 0
#define HMM_EINVAL 1       /* shape / index / ordering violation  -> Julia ArgumentError */
#define HMM_ECUDA 2        /* CUDA runtime error                  -> Julia ErrorException */
#define HMM_ENOMEM 3       /* host or device allocation failed */
#define HMM_ENODEV 4       /* no CUDA device */
#define HMM_EUNSUPPORTED 5 /* valid request outside what this build implements */

/* == Julia Tuple{Int64,Int64,Float64}; src/types.jl:3 */
typedef struct hmm_trans {
    int64_t src; /* 1-based */
    int64_t dst; /* 1-based */
    double lp;
} hmm_trans;

/* Decode engines (hmm_viterbi_ex_f64 `mode`). */
#define HMM_MODE_AUTO 0     /* ring engine for non-overlap ring models, generic time-parallel engine for other long decodes, else faithful */
#define HMM_MODE_FAITHFUL 1 /* sequential kernel in the reference's exact operation order (any StateMatrix) */
#define HMM_MODE_RING 2     /* time-parallel ring kernels; HMM_EUNSUPPORTED if the model is not ring-structured */
#define HMM_MODE_GENERIC 3  /* time-parallel per-state kernels for any StateMatrix (overlap models, N > 7); T >= 4096 */

/* Diagnostics of one decode / E-M call (all optional outputs). */
typedef struct hmm_info {
    int32_t engine;           /* HMM_MODE_FAITHFUL, _RING or _GENERIC actually used */
    int32_t n_chunks;         /* time chunks per channel (ring engine) */
    int32_t fwd_repaired;     /* chunks whose speculative forward start failed verification and were re-run */
    int32_t bwd_repaired;     /* same for the backward / traceback pass */
    int64_t kernel_launches;  /* kernels launched by this call */
    double device_ms;         /* CUDA-event time of the device work (H2D/D2H included for host-pointer entry points) */
    double kernel_ms;         /* CUDA-event time of the kernels only */
    double top_kernel_ms;     /* time of the dominant kernel (ring forward / E-step forward) */
} hmm_info;

/* ---- library / device ---------------------------------------------------- */
int hmm_version(void);              /* major*10000 + minor*100 + patch */
const char *hmm_last_error(void);   /* thread-local, never NULL */
int hmm_device_count(void);         /* number of CUDA devices, 0 if none */
int hmm_set_device(int device);     /* device used by subsequent calls from this thread */
int hmm_get_device(void);
/* Devices the host-pointer decode entry points spread their work over, INSIDE the library (one worker thread and
 * stream set per device): the channels of hmm_viterbi_batch_f64 in contiguous blocks, one long recording of
 * hmm_viterbi_f64 (>= 8 M samples, ring model) as time shards with peer-memory boundary exchange -- so a Julia caller
 * of viterbi / the batch call uses the whole box (the reference sorts one channel per process, src/hmmsort.jl:79-83).
 * n <= 1 or NULL restores the single-device behaviour (the device of hmm_set_device).  The environment variable
 * HMMCUDA_DEVICES=0,1,2,... does the same without a call.  Results are identical to the single-device decode. */
int hmm_set_devices(const int *devices, int n);
int hmm_get_devices(int *devices_out, int cap); /* returns the number of selected devices (0: single-device mode) */
/* Run the calling thread's subsequent work on `cuda_stream` (a cudaStream_t owned by the caller, e.g.
 * torch.cuda.current_stream().cuda_stream) so that it is stream-ordered with the caller's own kernels
 * and NCCL collectives; NULL restores the library's private stream (so the legacy default stream, whose
 * handle is 0, cannot be selected: use a created stream). */
int hmm_set_stream(void *cuda_stream);

/* Tunables of the ring engine (0 keeps the default): chunk length and
 * speculative warm-up / look-ahead, in samples (rounded to multiples of 256). */
int hmm_set_ring_params(int64_t chunk_len, int64_t warmup);

/* Frees the calling thread's grow-only device workspace and pinned staging buffers (they are otherwise kept
 * for the next call; a pageable-host decode of T samples leaves 10 T bytes of pinned staging behind). */
int hmm_release_workspace(void);

/* Repeated decodes of the same device buffers with the same model re-launch one cached CUDA graph (no per-kernel
 * timing possible).  hmm_set_profiling(1) makes the calling process launch eagerly with CUDA-event timers instead, so
 * that hmm_info.top_kernel_ms is filled (benchmarks / roofline reports); 0 restores the default. */
int hmm_set_profiling(int on);

/* ---- Viterbi ------------------------------------------------------------- */
/*
 * viterbi(y, lA::StateMatrix, mu, sigma) -> (x, ll)       src/viterbi.jl:44-98
 * T2_out / T1_out: nullable [nstates x T]; when given, the dense trellis of
 * src/viterbi.jl:52-53 is materialised (the `(x, T2, T1)` form of
 * README.md:34).  Host pointers, pageable or pinned (hmm_host_alloc): long ring-model
 * decodes overlap the upload, the decode and the download segment by segment, pageable
 * buffers going through the library's own pinned staging; the result is the same decode.
 */
int hmm_viterbi_f64(const double *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                    double *ll_out, int16_t *T2_out, double *T1_out);

/* Same, with engine selection and diagnostics. */
int hmm_viterbi_ex_f64(const double *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                       double *ll_out, int16_t *T2_out, double *T1_out, int32_t mode, hmm_info *info);

/*
 * Batched decode of C independent channels (BASELINE config 4; the reference
 * sorts one channel per process, src/hmmsort.jl:79-83).  y is [T x C]
 * column-major; every per-model array carries a leading channel stride:
 * states [N x nstates x C] (or one shared copy if states_shared != 0),
 * tr [ntrans x C], mu [K x N x C], sigma [C]; outputs x [T x C], ll [C].
 * All channels share N, K, nstates, ntrans (same topology, own weights).
 */
int hmm_viterbi_batch_f64(const double *y, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                          int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                          const double *mu, const double *sigma, int16_t *x_out, double *ll_out, int32_t mode,
                          hmm_info *info);

/* Device-resident variant: y_dev [T x C] and x_dev [T x C] are DEVICE
 * pointers on the current device (model arrays stay host pointers; they are
 * tiny).  ll_out is a host pointer (nullable).  Used when the recording is
 * already in HBM (bench `value`, multi-GPU orchestration). */
int hmm_viterbi_dev_f64(const double *y_dev, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                        int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                        const double *mu, const double *sigma, int16_t *x_dev, double *ll_out, int32_t mode,
                        hmm_info *info);

/* Roofline denominators measured on the current device: the FP64 FMA issue rate in GDFMA/s (DFMA with a constant-bank
 * operand, the form the matched-filter FIR uses) and the device-memory copy bandwidth in GB/s (read + written bytes of a
 * 512 MB copy).  Takes ~30 ms; benchmarks report their fractions against these instead of quoting constants. */
int hmm_measure_peaks(double *fp64_gdfma_per_s, double *copy_gb_per_s);

/* ---- FP32 mode (BASELINE north_star: "T1 and log-likelihoods within ... 1e-4 in FP32 mode") ------------------
 * The reference is Float64 only (src/viterbi.jl:44), so this mode is defined by the build: the recording may be
 * Float32 (half the PCIe / HBM bytes per sample; widened exactly on the device) and the matched-filter FIR of the
 * ring decode -- 99 % of its arithmetic -- runs in FP32 with Float32-rounded templates, at twice the FP64 rate.
 * The recursion, the per-chunk normalisation, the boundary verification and ll stay FP64, which is the periodic
 * FP64 renormalisation unnormalised FP32 scores would need.  T1 / ll agree with the FP64 decode of the
 * same values to <= 1e-4 relative; x is NOT promised bit-exact: it can differ where a decision margin is below the
 * FIR's FP32 rounding (~1e-4 on scores of O(100)).  hmm_set_precision(HMM_PREC_F32) (or HMMCUDA_PRECISION=f32)
 * makes the *_f64 decode entry points compute that way too; the *_f32 entry points always do.  Models other than
 * non-overlap ring models (faithful engine) and the Baum-Welch entry points compute in FP64 in either mode. */
#define HMM_PREC_F64 0
#define HMM_PREC_F32 1
int hmm_set_precision(int32_t precision);
int hmm_get_precision(void);
int hmm_viterbi_f32(const float *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                    double *ll_out);
int hmm_viterbi_ex_f32(const float *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                       double *ll_out, int32_t mode, hmm_info *info);
int hmm_viterbi_batch_f32(const float *y, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                          int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                          const double *mu, const double *sigma, int16_t *x_out, double *ll_out, int32_t mode,
                          hmm_info *info);
int hmm_viterbi_dev_f32(const float *y_dev, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                        int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                        const double *mu, const double *sigma, int16_t *x_dev, double *ll_out, int32_t mode,
                        hmm_info *info);

/* ---- time-sharded decode of ONE long recording across GPUs (BASELINE config 5) ------------- */
/*
 * Every rank owns a contiguous span [main_begin, main_end) of the recording (multiples of
 * chunk_len, except the global end) and holds the samples [local_begin, local_end) with one
 * "ghost" chunk on either side: local_begin = main_begin - chunk_len (0 on the first rank),
 * local_end = min(T_global, main_end + chunk_len).  The ghost chunks are decoded speculatively
 * and thrown away; the shard boundaries are then verified exactly like the chunk boundaries
 * inside one GPU, with two tiny messages per neighbour pair (exchanged by the caller, e.g. over
 * NCCL): the forward boundary vector (hmm_vshard_bvec() doubles) travelling right and the
 * traceback state (one int64) travelling left.  Protocol (all ranks, same order):
 *   create -> forward -> [send fwd boundary to r+1 / set from r-1] -> fwd_verify
 *          (repeat exchange + fwd_verify while any rank reports repairs)
 *          -> trace -> [send trace boundary to r-1 / set from r+1] -> trace_verify (repeat likewise)
 *          -> finish (x of the main span, partial ll to be summed over ranks) -> destroy
 * y_local is a DEVICE pointer unless y_is_host != 0 (then it is copied once).  Boundary buffers
 * are device pointers when *_is_device != 0 (then get/set are asynchronous on the library's
 * stream, see hmm_set_stream), host pointers otherwise.  fwd_verify / trace_verify with a NULL
 * n_repaired do not synchronise; hmm_vshard_repairs reads both counters afterwards.  A handle can
 * be re-run (forward ... finish) any number of times.  Single channel, ring models only.
 */
typedef struct hmm_vshard hmm_vshard;
int hmm_vshard_chunking(int64_t T_global, int32_t n_ranks, int32_t N, int32_t K, int64_t *chunk_len_out,
                        int64_t *warmup_out);
int hmm_vshard_create(const double *y_local, int32_t y_is_host, int64_t local_begin, int64_t local_end,
                      int64_t main_begin, int64_t main_end, int64_t T_global, int64_t chunk_len, int64_t warmup,
                      const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *mu, double sigma, hmm_vshard **out);
int hmm_vshard_bvec(const hmm_vshard *h);
/* new samples for a shard that was created from a host pointer (same spans): asynchronous upload into its buffer */
int hmm_vshard_set_y(hmm_vshard *h, const double *y_local_host);
int hmm_vshard_forward(hmm_vshard *h);
int hmm_vshard_fwd_boundary_get(hmm_vshard *h, double *out, int32_t out_is_device);        /* at main_end  -> rank r+1 */
int hmm_vshard_fwd_boundary_set(hmm_vshard *h, const double *in, int32_t in_is_device);   /* at main_begin <- rank r-1 */
int hmm_vshard_fwd_verify(hmm_vshard *h, int32_t *n_repaired);
int hmm_vshard_trace(hmm_vshard *h);
int hmm_vshard_trace_boundary_get(hmm_vshard *h, int64_t *out, int32_t out_is_device);      /* at main_begin -> rank r-1 */
int hmm_vshard_trace_boundary_set(hmm_vshard *h, const int64_t *in, int32_t in_is_device); /* at main_end   <- rank r+1 */
int hmm_vshard_trace_verify(hmm_vshard *h, int32_t *n_repaired);
int hmm_vshard_finish(hmm_vshard *h, int16_t *x_main_out, int32_t x_is_device, double *ll_partial_out);
/* finish + repair counters with a single synchronisation */
int hmm_vshard_finish_ex(hmm_vshard *h, int16_t *x_main_out, int32_t x_is_device, double *ll_partial_out,
                         int32_t *fwd_repaired, int32_t *trace_repaired);
/* One-collective protocol (no host synchronisation until the verdict is read).  After forward ->
 * fwd_verify -> trace -> trace_verify, run purely locally on the ghost chunks' results:
 *   summary_dev : x of the main span into x_main_dev (may be NULL) and, into summary_dev
 *                 (hmm_vshard_summary_len() doubles): the true forward vector at main_end, the speculative one
 *                 this shard started from at main_begin, the traceback states at main_begin (own) and at
 *                 main_end (assumed), and the partial ll -- all on the library's stream;
 *   [caller: ONE all-gather of the summaries over the ranks, rank order]
 *   judge_dev   : every rank checks EVERY shard boundary of the gathered summaries (same arithmetic as the
 *                 chunk boundaries inside one GPU), so all ranks reach the same verdict without another
 *                 collective: out_dev[0] = total ll, out_dev[1] = number of inconsistent shard boundaries.
 * A non-zero verdict means some ghost chunk guessed wrong: fall back to the exchange / verify rounds above. */
int hmm_vshard_summary_len(const hmm_vshard *h);
int hmm_vshard_summary_dev(hmm_vshard *h, int16_t *x_main_dev, double *summary_dev);
int hmm_vshard_judge_dev(hmm_vshard *h, const double *gathered_dev, int32_t n_ranks, double *out_dev);
/* Peer-memory protocol: the same summaries, but every rank STORES its summary straight into every peer's exchange block
 * over NVLink (CUDA-IPC pointers between processes, peer access inside one process) and raises a flag there; the judge
 * kernel spins on the flags of its own block.  No collective call and no host round trip between decode and verdict;
 * the local decode + exchange is one CUDA graph from the second call on.
 *   p2p_init   : allocates this rank's exchange block; ipc_handle_out (64 bytes, nullable) for other processes,
 *                block_ptr_out (nullable) for shards driven from the same process;
 *   [caller: all-gather the handles (or pointers) once, rank order]
 *   p2p_attach : opens the peers' blocks (ipc_handles: world x 64 bytes) or takes their pointers (block_ptrs[world]);
 *   p2p_launch : local decode, x of the main span into x_main_dev (nullable), summary to all peers -- asynchronous;
 *   p2p_finish : judge + the decode's single synchronisation: total ll and the number of inconsistent shard
 *                boundaries (non-zero: fall back to the exchange / verify rounds above).  Every rank must call
 *                launch and finish the same number of times. */
int hmm_vshard_p2p_init(hmm_vshard *h, int32_t rank, int32_t world, void *ipc_handle_out, void **block_ptr_out);
int hmm_vshard_p2p_attach(hmm_vshard *h, const void *ipc_handles, void *const *block_ptrs);
int hmm_vshard_p2p_launch(hmm_vshard *h, int16_t *x_main_dev);
int hmm_vshard_p2p_finish(hmm_vshard *h, double *ll_total_out, int32_t *bad_out);
int hmm_vshard_repairs(hmm_vshard *h, int32_t *fwd_repaired, int32_t *trace_repaired); /* of the last fwd_verify / trace_verify */
int hmm_vshard_destroy(hmm_vshard *h);

/* ---- Baum-Welch ---------------------------------------------------------- */
/* forward(V, lA, mu, sigma) -> alpha [nstates x T]          src/baumwelch.jl:25-51 */
int hmm_forward_f64(const double *V, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, double *alpha_out);

/* backward(V, lA, mu, sigma) -> beta [nstates x T]          src/baumwelch.jl:73-98 */
int hmm_backward_f64(const double *V, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                     const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, double *beta_out);

/*
 * update(alpha, beta, lA, mu, sigma, x) -> (lA_new, mu, sigma)   src/baumwelch.jl:205-309
 * mu_inout is overwritten in place like the reference's fill!(mu, 0.0)
 * (src/baumwelch.jl:268); *sigma_inout receives the new sigma.  The new
 * StateMatrix is rebuilt by the caller (the unchanged Julia constructor,
 * src/types.jl:148-151) from lp_out [nxi-1] (= xb[2:end], :264-265, nxi =
 * number of transitions out of state 1; == N for non-overlap models) and
 * pp_out [nstates] (= gamma[:,1], :263).
 */
int hmm_update_f64(const double *alpha, const double *beta, int64_t T, const int16_t *states, int32_t N, int32_t K,
                   int32_t nstates, const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout,
                   const double *x, double *lp_out, double *pp_out);

/*
 * One fused E/M step: train_model(X, lA, mu0, sigma0) of src/baumwelch.jl:362-370
 * without alpha/beta/gamma crossing the boundary.  loglik_out (nullable)
 * receives log p(X | model) = LSE_j alpha[j,T] (not returned by the
 * reference).
 */
int hmm_em_step_f64(const double *X, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                    double *pp_out, double *loglik_out);

int hmm_em_step_ex_f64(const double *X, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                       double *pp_out, double *loglik_out, int32_t mode, hmm_info *info);

/*
 * Device-resident training context: keeps X in HBM across the E/M
 * iterations of train_model's outer loop (src/baumwelch.jl:324-354), so each
 * iteration moves only the model (KBs) across PCIe.  The host loop (callback,
 * yield, merge/prune) stays with the caller.
 */
typedef struct hmm_train_ctx hmm_train_ctx;
int hmm_train_create(const double *X, int64_t T, hmm_train_ctx **ctx_out);          /* X: host pointer, copied once */
int hmm_train_create_dev(const double *X_dev, int64_t T, hmm_train_ctx **ctx_out);  /* X_dev: device pointer, borrowed */
int hmm_train_em_step(hmm_train_ctx *ctx, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                      const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                      double *pp_out, double *loglik_out, hmm_info *info);
/*
 * The E/M loop of train_model (src/baumwelch.jl:325-335) in one call, for callers that pass no callback: nsteps
 * iterations, each one's lp feeding the next one's transition weights inside the library (the weight part of the
 * StateMatrix rebuild, src/baumwelch.jl:265 / src/types.jl:94-127; the set of finite transitions depends on the state
 * layout only).  tr_inout's weights, mu_inout and sigma_inout are updated in place; lp_out [nlp] (nlp = transitions out
 * of state 1 minus one), pp_out [nstates] and loglik_out [nsteps, nullable] as in hmm_em_step_f64.  *steps_done <
 * nsteps when a weight stopped being finite (degenerate lp): the caller rebuilds its StateMatrix and decides.
 */
int hmm_train_run(hmm_train_ctx *ctx, const int16_t *states, int32_t N, int32_t K, int32_t nstates, hmm_trans *tr_inout,
                  int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out, int32_t nlp, double *pp_out,
                  double *loglik_out, int32_t nsteps, int32_t *steps_done, hmm_info *info);
int hmm_train_destroy(hmm_train_ctx *ctx);

/*
 * Host-only model management helper (no device needed): the WEIGHTS of an unchanged set of transitions for a new lp
 * vector -- what `StateMatrix(states, pp, K, lp; ...)` (src/types.jl:94-127, called by `update` at
 * src/baumwelch.jl:265) changes when the state layout stays the same, without its O(nstates^2 N) scan.  Per
 * transition and neuron the term is lpz = log1p(-exp(sum(lp))) (silent -> silent), lp[i] (silent -> first phase) or 0,
 * added in neuron order.  *all_finite = 0 (and tr_inout untouched) when a weight would not be finite: the set of
 * transitions changes and the caller has to run the constructor.  hmm_train_run uses the same routine between steps.
 */
int hmm_transition_weights(const int16_t *states, int32_t N, int32_t nstates, hmm_trans *tr_inout, int64_t ntrans,
                           const double *lp, int32_t nlp, int32_t *all_finite);

/* ---- I/O front-end (src/hmmsort.jl:36-104: data file -> Float64 -> decode -> unrolled sequence) ------------------
 * Decodes C channels of a raw recording FILE without the samples ever passing through a caller array: the file is read
 * block by block into pinned staging by a reader thread, every block crosses PCIe while the next is read, one kernel
 * picks the requested channels out of it and widens the samples (value = raw * scale) into the Float64 [T x C] layout,
 * and all channels are decoded from HBM.  The file holds n_file_channels channels of T samples from byte_offset on, either
 * interleaved ([T x n_file_channels], sample-major, the usual acquisition layout) or channel-major; a contiguous
 * (uncompressed) HDF5 dataset -- what the reference memory-maps, src/hmmsort.jl:72-76 -- is such a block at its data
 * offset.  channels[C]: 0-based indices into the file's channels; the model arrays as in hmm_viterbi_batch_f64. */
#define HMM_RAW_F64 0
#define HMM_RAW_F32 1
#define HMM_RAW_I16 2
int hmm_viterbi_rawfile(const char *path, int64_t byte_offset, int32_t sample_dtype, int32_t n_file_channels,
                        int32_t interleaved, double scale, int64_t T, int32_t C, const int32_t *channels,
                        const int16_t *states, int32_t states_shared, int32_t N, int32_t K, int32_t nstates,
                        const hmm_trans *tr, int64_t ntrans, const double *mu, const double *sigma, int16_t *x_out,
                        double *ll_out, int32_t mode, hmm_info *info);

/* ---- time-sharded Baum-Welch: one recording over several GPUs ------------------------------------
 * Every rank owns a contiguous span [main_begin, main_end) of X (chunk aligned) and holds [local_begin, local_end) with at
 * least one ghost chunk on either side.  One E/M iteration of src/baumwelch.jl:362-370 over all ranks:
 *   estep  (every rank, asynchronous on the library's stream): forward, backward and the sufficient statistics of the
 *          rank's main span -> stats_dev (hmm_emshard_stats_len doubles: sum gamma, sum gamma y, sum y^2, the xi sums,
 *          S1[i][s] = sum pi_i(t0) y[t0+s], end-of-recording corrections on the first / last rank, gamma[:,1] on the
 *          first) and boundary_dev (hmm_emshard_boundary_len doubles: the forward and backward boundary vectors at
 *          main_begin and at main_end in the rank's own normalisation, then its local log-likelihood term);
 *   [caller: ALL-REDUCE (sum) of stats over the ranks; ALL-GATHER of boundary.  Neighbours' vectors for the same instant
 *    must agree up to a constant (1e-11 relative): that verifies that one ghost chunk made the main-span posteriors
 *    exact; the constants chain the ranks' normalisations: lS_global = lS_last + sum_r (fwd_r-1@end - fwd_r@begin)]
 *   mstep  (every rank, identical result): new mu (in place, src/baumwelch.jl:268), sigma, lp = xb[2:end], pp = gamma[:,1],
 *          log-likelihood.
 * hmmspikesorter.jl_b200/timeshard.py (EmSharded) drives this over torch.distributed (NCCL) or over shards on one GPU. */
typedef struct hmm_emshard hmm_emshard;
int hmm_emshard_create(const double *X_local, int32_t x_is_host, int64_t local_begin, int64_t local_end,
                       int64_t main_begin, int64_t main_end, int64_t T_global, int64_t chunk_len, int64_t warmup,
                       hmm_emshard **out);
/* chunk length / warm-up the E-step of a rank's share would choose on its own (one chunk per resident warp) */
int hmm_emshard_chunking(int64_t T_global, int32_t n_ranks, int32_t N, int32_t K, int64_t *chunk_len_out,
                         int64_t *warmup_out);
int hmm_emshard_stats_len(int32_t N, int32_t nstates);
int hmm_emshard_boundary_len(int32_t N, int32_t K);
int hmm_emshard_estep(hmm_emshard *h, const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *mu, double sigma, double *stats_dev, double *boundary_dev);
int hmm_emshard_mstep(hmm_emshard *h, const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *stats_sum_dev, double lS_global, double *mu_inout, double *sigma_inout,
                      double *lp_out, double *pp_out, double *loglik_out);
int hmm_emshard_destroy(hmm_emshard *h);

/* ---- reconstruction ------------------------------------------------------ */
/* reconstruct_signal(x, lA, mu, sigma) -> Y [T]             src/reconstruction.jl:1-9 */
int hmm_reconstruct_f64(const int16_t *x, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                        const double *mu, int32_t K, double *Y_out);
/* device-pointer variant: x_dev, Y_dev on the current device */
int hmm_reconstruct_dev_f64(const int16_t *x_dev, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                            const double *mu, int32_t K, double *Y_dev);
/* unroll_mlseq(mlseq, state_matrix) -> Int16 [N x T]        src/extraction.jl:4-13 */
int hmm_unroll_mlseq_i16(const int16_t *x, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                         int16_t *out);

/* ---- pinned host memory for callers that want full PCIe rate ------------- */
int hmm_host_alloc(void **ptr_out, uint64_t bytes);
int hmm_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* HMMCUDA_H */
