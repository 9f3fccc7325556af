// common.h -- error handling, device buffers and the analysed host model shared
// by every translation unit of libhmmcuda.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hmmcuda.h"

namespace hmm {

struct Error {
    int code;
    std::string msg;
};

[[noreturn]] inline void fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error{code, buf};
}

#define HMM_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            ::hmm::fail(e__ == cudaErrorMemoryAllocation ? HMM_ENOMEM : HMM_ECUDA, "%s:%d: %s -> %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e__));                       \
    } while (0)

// Grow-only device workspace: one slot per logical buffer, kept across calls so
// that steady-state calls do no cudaMalloc/cudaFree (both synchronise the device).
class Workspace {
  public:
    enum Slot {
        Y = 0, X, MODEL, DEC, MASK, BOUND_S, BOUND_E, CHUNKS, TRACE, T1, T2, PSCORE, MISC, FWDQ, FWDF, STATS, PROLOG,
        ALPHA, BETA, SCRATCH, NSLOTS
    };
    void *get(Slot s, size_t bytes);
    // Grow-only pinned, device-mapped host buffers (six slots): staging for per-call uploads / pageable callers and a
    // landing zone kernels can write results into directly.  *dev_ptr receives the device-side alias.
    void *pinned(int slot, size_t bytes, void **dev_ptr = nullptr);
    void release();
    ~Workspace() { release(); }

  private:
    void bind_device();
    void *ptr_[NSLOTS] = {};
    size_t cap_[NSLOTS] = {};
    void *hptr_[6] = {};
    size_t hcap_[6] = {};
    int dev_ = -1;
};

Workspace &workspace();      // per host thread
cudaStream_t main_stream();  // per host thread, non-blocking stream
cudaStream_t copy_stream();  // host -> device copies of the pipelined decode
cudaStream_t out_stream();   // device -> host copies of the pipelined decode
void set_external_stream(cudaStream_t s, bool use);  // per host thread
cudaEvent_t sync_event(int k);  // k = 0..3: per host thread, timing disabled (fork/join between the streams)

// ---------------------------------------------------------------------------
// Host-side analysis of one StateMatrix + templates (one channel).
// ---------------------------------------------------------------------------
struct RingParams {
    int N = 0, L = 0;  // neurons, chain length K-1
    double w_nn = 0;   // noise -> noise
    std::vector<double> w_nh;  // [N]    noise -> head_i
    std::vector<double> w_tn;  // [N]    tail_i -> noise
    std::vector<double> w_th;  // [N*N]  tail_j -> head_i at [j*N+i], -inf on the diagonal
    std::vector<double> w_c;   // [N*(L-1)] (i,s)->(i,s+1) at [i*(L-1)+s-1]
};

struct HostModel {
    int N = 0, K = 0, nstates = 0;
    int64_t ntrans = 0;
    double sigma = 0, lsig = 0;
    std::vector<double> m;  // state means, src/viterbi.jl:68-71 order
    // CSR by destination, list order preserved inside each row (tie-break order)
    std::vector<int> in_ptr, in_src;
    std::vector<double> in_lp;
    // CSR by source, list order preserved (backward recursion order)
    std::vector<int> out_ptr, out_dst;
    std::vector<double> out_lp;
    std::vector<int> dec_slot;     // [nstates] slot among multi-predecessor states, else -1
    std::vector<int> static_pred;  // [nstates] the single predecessor (0 if none), 0-based
    int ndec = 0;
    std::vector<int> xi_edge;      // list indices of transitions leaving state 1, list order
    bool is_ring = false;
    RingParams ring;
};

// Validates shapes / indices (HMM_EINVAL on violation) and fills `out`.
void analyse_model(const int16_t *states, int N, int K, int nstates, const hmm_trans *tr, int64_t ntrans,
                   const double *mu, double sigma, HostModel &out);

void state_means(const int16_t *states, int N, int K, int nstates, const double *mu, std::vector<double> &m);

// NVTX range around a stage (visible in Nsight Systems / Compute timelines; a no-op without a profiler attached:
// NVTX v3 is header-only and binds lazily).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Timer {  // CUDA-event stopwatch on a stream
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, s); }
    void stop() { cudaEventRecord(b, s); }
    float ms() {
        float t = 0;
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&t, a, b);
        return t;
    }
};

}  // namespace hmm
