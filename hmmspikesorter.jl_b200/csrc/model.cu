// model.cu -- host-side validation and analysis of the StateMatrix arrays that
// cross the C ABI (src/types.jl:1-9), workspace and stream management.
#include <cmath>
#include <cstring>
#include <limits>
#include <map>

#include "common.h"

namespace hmm {

// ---------------------------------------------------------------------------
void Workspace::bind_device() {  // buffers belong to one device: switching devices drops them
    int dev = 0;
    HMM_CUDA(cudaGetDevice(&dev));
    if (dev != dev_) {
        release();
        dev_ = dev;
    }
}

void *Workspace::get(Slot s, size_t bytes) {
    bind_device();
    if (bytes == 0) bytes = 16;
    if (cap_[s] < bytes) {
        if (ptr_[s]) {
            HMM_CUDA(cudaFree(ptr_[s]));
            ptr_[s] = nullptr;
            cap_[s] = 0;
        }
        size_t want = bytes + bytes / 8 + 256;  // a little slack so growing sizes do not thrash
        cudaError_t e = cudaMalloc(&ptr_[s], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&ptr_[s], want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            fail(HMM_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        cap_[s] = want;
    }
    return ptr_[s];
}

void *Workspace::pinned(int slot, size_t bytes, void **dev_ptr) {
    bind_device();
    if (bytes == 0) bytes = 16;
    if (hcap_[slot] < bytes) {
        if (hptr_[slot]) {
            HMM_CUDA(cudaFreeHost(hptr_[slot]));
            hptr_[slot] = nullptr;
            hcap_[slot] = 0;
        }
        const size_t want = bytes + bytes / 4 + 4096;
        HMM_CUDA(cudaHostAlloc(&hptr_[slot], want, cudaHostAllocMapped | cudaHostAllocPortable));
        hcap_[slot] = want;
    }
    if (dev_ptr) HMM_CUDA(cudaHostGetDevicePointer(dev_ptr, hptr_[slot], 0));
    return hptr_[slot];
}

void Workspace::release() {
    for (int i = 0; i < NSLOTS; i++) {
        if (ptr_[i]) cudaFree(ptr_[i]);
        ptr_[i] = nullptr;
        cap_[i] = 0;
    }
    for (int i = 0; i < 6; i++) {
        if (hptr_[i]) cudaFreeHost(hptr_[i]);
        hptr_[i] = nullptr;
        hcap_[i] = 0;
    }
}

Workspace &workspace() {
    static thread_local Workspace ws;
    return ws;
}

struct Streams {
    cudaStream_t main = nullptr, copy = nullptr, out = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int dev = -1;
};
static thread_local cudaStream_t g_external = nullptr;
static thread_local bool g_use_external = false;
void set_external_stream(cudaStream_t s, bool use) {
    g_external = s;
    g_use_external = use;
}
static Streams &streams() {
    static thread_local Streams st;
    int dev = 0;
    HMM_CUDA(cudaGetDevice(&dev));
    if (st.dev != dev) {
        st.main = st.copy = st.out = nullptr;  // streams of another device are simply abandoned
        for (auto &e : st.ev) e = nullptr;
        st.dev = dev;
    }
    if (!st.main) {
        HMM_CUDA(cudaStreamCreateWithFlags(&st.main, cudaStreamNonBlocking));
        HMM_CUDA(cudaStreamCreateWithFlags(&st.copy, cudaStreamNonBlocking));
        HMM_CUDA(cudaStreamCreateWithFlags(&st.out, cudaStreamNonBlocking));
    }
    return st;
}
cudaStream_t main_stream() { return g_use_external ? g_external : streams().main; }
cudaStream_t copy_stream() { return streams().copy; }
cudaStream_t out_stream() { return streams().out; }
cudaEvent_t sync_event(int k) {
    Streams &st = streams();
    if (!st.ev[k]) HMM_CUDA(cudaEventCreateWithFlags(&st.ev[k], cudaEventDisableTiming));
    return st.ev[k];
}

// ---------------------------------------------------------------------------
// m[j] = sum_{l=1..N} mu[states[l,j], l] accumulated from 0.0 in neuron order
// (src/viterbi.jl:68-71, src/baumwelch.jl:32-35, src/reconstruction.jl:3-7).
void state_means(const int16_t *states, int N, int K, int nstates, const double *mu, std::vector<double> &m) {
    m.resize(nstates);
    for (int j = 0; j < nstates; j++) {
        double s = 0.0;
        for (int l = 0; l < N; l++) s += mu[(states[l + (size_t)N * j] - 1) + (size_t)K * l];
        m[j] = s;
    }
}

static bool detect_ring(const int16_t *states, const hmm_trans *tr, int64_t ntrans, HostModel &M) {
    const int N = M.N, K = M.K, L = K - 1, ns = M.nstates;
    if (N < 1 || L < 2 || ns != 1 + N * L) return false;
    if (ntrans != (int64_t)(N + 1) + (int64_t)N * N + (int64_t)N * (L - 1)) return false;
    // state layout of generate_states(N, K, false), src/types.jl:71-77
    for (int l = 0; l < N; l++)
        if (states[l] != 1) return false;
    for (int i = 0; i < N; i++)
        for (int s = 1; s <= L; s++) {
            int j = 1 + i * L + (s - 1);
            for (int l = 0; l < N; l++)
                if (states[l + (size_t)N * j] != (l == i ? s + 1 : 1)) return false;
        }
    auto head = [&](int i) { return 1 + i * L; };      // 0-based
    const double NEG = -std::numeric_limits<double>::infinity();
    RingParams &R = M.ring;
    R.N = N;
    R.L = L;
    R.w_nh.assign(N, NEG);
    R.w_tn.assign(N, NEG);
    R.w_th.assign((size_t)N * N, NEG);
    R.w_c.assign((size_t)N * (L - 1), NEG);
    R.w_nn = NEG;
    std::vector<int> phase_of(ns, 0), neuron_of(ns, -1);
    for (int i = 0; i < N; i++)
        for (int s = 1; s <= L; s++) {
            phase_of[head(i) + s - 1] = s;
            neuron_of[head(i) + s - 1] = i;
        }
    for (int64_t e = 0; e < ntrans; e++) {
        int s = (int)tr[e].src - 1, d = (int)tr[e].dst - 1;
        double w = tr[e].lp;
        if (!std::isfinite(w)) return false;
        double *slot = nullptr;
        if (s == 0 && d == 0)
            slot = &R.w_nn;
        else if (s == 0) {
            if (phase_of[d] != 1) return false;
            slot = &R.w_nh[neuron_of[d]];
        } else if (phase_of[s] == L) {  // tail
            int j = neuron_of[s];
            if (d == 0)
                slot = &R.w_tn[j];
            else {
                if (phase_of[d] != 1 || neuron_of[d] == j) return false;
                slot = &R.w_th[(size_t)j * N + neuron_of[d]];
            }
        } else {  // chain interior
            if (d != s + 1) return false;
            slot = &R.w_c[(size_t)neuron_of[s] * (L - 1) + phase_of[s] - 1];
        }
        if (*slot != NEG) return false;  // duplicate edge
        *slot = w;
    }
    // every ring edge present?
    if (R.w_nn == NEG) return false;
    for (int i = 0; i < N; i++) {
        if (R.w_nh[i] == NEG || R.w_tn[i] == NEG) return false;
        for (int s = 0; s < L - 1; s++)
            if (R.w_c[(size_t)i * (L - 1) + s] == NEG) return false;
        for (int j = 0; j < N; j++)
            if (i != j && R.w_th[(size_t)j * N + i] == NEG) return false;
    }
    // candidate order at the decision states must be [noise, tail_1..tail_N]
    // (ascending source) -- the order the ring kernels break ties in.
    for (int d = 0; d < ns; d++)
        for (int e = M.in_ptr[d] + 1; e < M.in_ptr[d + 1]; e++)
            if (M.in_src[e] <= M.in_src[e - 1]) return false;
    return true;
}

void analyse_model(const int16_t *states, int N, int K, int nstates, const hmm_trans *tr, int64_t ntrans,
                   const double *mu, double sigma, HostModel &M) {
    if (!states || !tr || !mu) fail(HMM_EINVAL, "null model array");
    if (N < 1 || K < 1 || nstates < 1) fail(HMM_EINVAL, "N, K, nstates must be positive (got %d, %d, %d)", N, K, nstates);
    if (nstates > 32767) fail(HMM_EINVAL, "nstates=%d exceeds the Int16 range of the state sequence", nstates);
    if (ntrans < 0 || ntrans > (int64_t)1 << 30) fail(HMM_EINVAL, "bad transition count %lld", (long long)ntrans);
    if (!(sigma > 0) || !std::isfinite(sigma)) fail(HMM_EINVAL, "sigma must be positive and finite");
    for (size_t i = 0; i < (size_t)N * nstates; i++)
        if (states[i] < 1 || states[i] > K)
            fail(HMM_EINVAL, "states[%zu]=%d outside 1..K=%d", i, (int)states[i], K);
    for (int64_t e = 0; e < ntrans; e++)
        if (tr[e].src < 1 || tr[e].src > nstates || tr[e].dst < 1 || tr[e].dst > nstates)
            fail(HMM_EINVAL, "transition %lld = (%lld -> %lld) outside 1..nstates=%d", (long long)e,
                 (long long)tr[e].src, (long long)tr[e].dst, nstates);
    M.N = N;
    M.K = K;
    M.nstates = nstates;
    M.ntrans = ntrans;
    M.sigma = sigma;
    M.lsig = std::log(sigma);  // src/viterbi.jl:47
    state_means(states, N, K, nstates, mu, M.m);

    // CSR by destination / by source, stable in list order
    M.in_ptr.assign(nstates + 1, 0);
    M.out_ptr.assign(nstates + 1, 0);
    for (int64_t e = 0; e < ntrans; e++) {
        M.in_ptr[tr[e].dst]++;
        M.out_ptr[tr[e].src]++;
    }
    for (int j = 0; j < nstates; j++) {
        M.in_ptr[j + 1] += M.in_ptr[j];
        M.out_ptr[j + 1] += M.out_ptr[j];
    }
    M.in_src.resize(ntrans);
    M.in_lp.resize(ntrans);
    M.out_dst.resize(ntrans);
    M.out_lp.resize(ntrans);
    {
        std::vector<int> fi(M.in_ptr.begin(), M.in_ptr.end() - 1), fo(M.out_ptr.begin(), M.out_ptr.end() - 1);
        for (int64_t e = 0; e < ntrans; e++) {
            int s = (int)tr[e].src - 1, d = (int)tr[e].dst - 1;
            M.in_src[fi[d]] = s;
            M.in_lp[fi[d]++] = tr[e].lp;
            M.out_dst[fo[s]] = d;
            M.out_lp[fo[s]++] = tr[e].lp;
        }
    }
    M.dec_slot.assign(nstates, -1);
    M.static_pred.assign(nstates, 0);
    M.ndec = 0;
    for (int j = 0; j < nstates; j++) {
        int deg = M.in_ptr[j + 1] - M.in_ptr[j];
        if (deg >= 2)
            M.dec_slot[j] = M.ndec++;
        else if (deg == 1)
            M.static_pred[j] = M.in_src[M.in_ptr[j]];
    }
    M.xi_edge.clear();
    for (int64_t e = 0; e < ntrans; e++)
        if (tr[e].src == 1) M.xi_edge.push_back((int)e);
    M.is_ring = detect_ring(states, tr, ntrans, M);
}

}  // namespace hmm
