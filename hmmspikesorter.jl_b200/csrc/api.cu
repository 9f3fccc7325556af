// api.cu -- the extern "C" boundary of libhmmcuda.so (include/hmmcuda.h).
// Plain pointers and sizes in, status codes out; every exception is caught
// here and turned into a code + thread-local message.  There is no CPU path:
// without a device every compute entry point returns HMM_ENODEV.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>

#include "engines.h"

using namespace hmm;

static constexpr int64_t RING_Q_MIN = 128;  // the ring engine's look-back length

static thread_local std::string g_err = "";
static std::mutex g_entry;  // the library serialises concurrent entry (SURVEY 8b "Threading")

static int set_err(int code, const std::string &m) {
    g_err = m;
    return code;
}

static void begin_call();  // advances the cache epochs (entries used by the running call are never evicted)

// Worker threads of the multi-device dispatcher call the public entry points on behalf of a caller that already
// holds the entry lock: they do not take it again.
static thread_local bool t_worker = false;

template <class F>
static int guarded(F &&f) {
    std::unique_lock<std::mutex> lk(g_entry, std::defer_lock);
    if (!t_worker) lk.lock();
    try {
        g_err.clear();
        begin_call();
        f();
        return HMM_OK;
    } catch (const Error &e) {
        cudaGetLastError();
        return set_err(e.code, e.msg);
    } catch (const std::bad_alloc &) {
        return set_err(HMM_ENOMEM, "host allocation failed");
    } catch (const std::exception &e) {
        return set_err(HMM_ECUDA, e.what());
    }
}

static void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        fail(HMM_ENODEV, "no CUDA device available (libhmmcuda has no CPU fallback): %s",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
}

// Copies `bytes` from a host pointer of unknown kind (pageable or pinned) to the device.
static void h2d(void *dst, const void *src, size_t bytes, cudaStream_t st) {
    HMM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
}
static void d2h(void *dst, const void *src, size_t bytes, cudaStream_t st) {
    HMM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
}

static void drop_caches();
static bool g_precision_from_env = false;
// HMMCUDA_PRECISION=f32 selects FP32 mode without a call (read once, at the first decode)
static void precision_from_env_once() {
    if (g_precision_from_env) return;
    g_precision_from_env = true;
    const char *e = getenv("HMMCUDA_PRECISION");
    if (e && (!strcmp(e, "f32") || !strcmp(e, "F32") || !strcmp(e, "float32"))) ring_config().precision = HMM_PREC_F32;
}

extern "C" {

int hmm_version(void) { return 200; /* 0.2.0 */ }
const char *hmm_last_error(void) { return g_err.c_str(); }

int hmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hmm_set_device(int device) {
    return guarded([&] {
        require_device();
        HMM_CUDA(cudaSetDevice(device));
    });
}

int hmm_get_device(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return d;
}

int hmm_set_stream(void *cuda_stream) {
    set_external_stream((cudaStream_t)cuda_stream, cuda_stream != nullptr);
    return HMM_OK;
}

int hmm_set_ring_params(int64_t chunk_len, int64_t warmup) {
    return guarded([&] {  // under the entry lock: a decode on another thread never sees half an update
        if (chunk_len < 0 || warmup < 0) fail(HMM_EINVAL, "negative ring parameter");
        ring_config().chunk_len = chunk_len;
        ring_config().warmup = warmup;
    });
}

int hmm_release_workspace(void) {
    return guarded([&] {
        cudaDeviceSynchronize();
        drop_caches();
        workspace().release();
    });
}

int hmm_set_profiling(int on) {
    return guarded([&] { ring_config().profile = on ? 1 : 0; });
}

int hmm_measure_peaks(double *fp64_gdfma_per_s, double *copy_gb_per_s) {
    return guarded([&] {
        if (!fp64_gdfma_per_s || !copy_gb_per_s) fail(HMM_EINVAL, "null argument");
        require_device();
        measure_peaks(fp64_gdfma_per_s, copy_gb_per_s, main_stream());
    });
}

int hmm_set_precision(int32_t precision) {
    return guarded([&] {
        if (precision != HMM_PREC_F64 && precision != HMM_PREC_F32) fail(HMM_EINVAL, "precision must be HMM_PREC_F64 or HMM_PREC_F32");
        g_precision_from_env = true;  // an explicit call wins over HMMCUDA_PRECISION
        ring_config().precision = precision;
    });
}
int hmm_get_precision(void) { return ring_config().precision; }

int hmm_host_alloc(void **ptr_out, uint64_t bytes) {
    return guarded([&] {
        require_device();
        if (!ptr_out) fail(HMM_EINVAL, "null ptr_out");
        HMM_CUDA(cudaHostAlloc(ptr_out, bytes ? bytes : 1, cudaHostAllocDefault));
    });
}
int hmm_host_free(void *ptr) {
    return guarded([&] {
        if (ptr) HMM_CUDA(cudaFreeHost(ptr));
    });
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Multi-device dispatch inside the library (hmm_set_devices): one persistent worker thread per selected GPU, each
// with its own thread-local streams, workspace and caches.  A host-pointer call made by the application is split
// by the calling thread (which holds the entry lock) and the pieces run on the workers through the same public
// entry points.
// ---------------------------------------------------------------------------
namespace {

class DeviceWorker {
  public:
    explicit DeviceWorker(int dev) : dev_(dev), th_([this] { loop(); }) {}
    ~DeviceWorker() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }
    int device() const { return dev_; }
    // runs f on the worker thread (device selected); the future carries (status, message)
    std::future<std::pair<int, std::string>> run(std::function<int()> f) {
        auto task = std::make_shared<std::packaged_task<std::pair<int, std::string>()>>([f] {
            const int rc = f();
            return std::make_pair(rc, rc ? std::string(hmm_last_error()) : std::string());
        });
        auto fut = task->get_future();
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back([task] { (*task)(); });
        }
        cv_.notify_one();
        return fut;
    }

  private:
    void loop() {
        t_worker = true;
        cudaSetDevice(dev_);
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    int dev_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false;
    std::thread th_;
};

std::vector<std::unique_ptr<DeviceWorker>> g_workers;  // under the entry lock
bool g_devices_from_env = false;

void drop_shardset();

void set_devices_locked(const int *devs, int n) {
    drop_shardset();
    g_workers.clear();
    if (!devs || n <= 1) return;  // zero or one device: the ordinary single-device path on the current device
    int have = 0;
    HMM_CUDA(cudaGetDeviceCount(&have));
    for (int k = 0; k < n; k++) {
        if (devs[k] < 0 || devs[k] >= have) fail(HMM_EINVAL, "device %d does not exist (%d visible)", devs[k], have);
        // (HMMCUDA_DEBUG_ALLOW_DUP_DEVICES=1: tests drive the dispatcher with several workers on one GPU)
        const bool dup_ok = getenv("HMMCUDA_DEBUG_ALLOW_DUP_DEVICES") && atoi(getenv("HMMCUDA_DEBUG_ALLOW_DUP_DEVICES"));
        for (int j = 0; j < k && !dup_ok; j++)
            if (devs[j] == devs[k]) fail(HMM_EINVAL, "device %d listed twice", devs[k]);
    }
    for (int k = 0; k < n; k++) g_workers.emplace_back(new DeviceWorker(devs[k]));
}

// HMMCUDA_DEVICES=0,1,2,3 selects the devices without a call (read once, at the first decode)
void devices_from_env_once() {
    if (g_devices_from_env) return;
    g_devices_from_env = true;
    const char *e = getenv("HMMCUDA_DEVICES");
    if (!e || !*e || !g_workers.empty()) return;
    std::vector<int> d;
    for (const char *q = e; *q;) {
        char *end = nullptr;
        long v = strtol(q, &end, 10);
        if (end == q) break;
        d.push_back((int)v);
        q = *end == ',' ? end + 1 : end;
    }
    set_devices_locked(d.data(), (int)d.size());
}

// waits for every piece; the first failure becomes the call's failure
void join_all(std::vector<std::future<std::pair<int, std::string>>> &futs) {
    int rc = HMM_OK;
    std::string msg;
    for (auto &f : futs) {
        auto r = f.get();
        if (r.first != HMM_OK && rc == HMM_OK) {
            rc = r.first;
            msg = r.second;
        }
    }
    futs.clear();
    if (rc != HMM_OK) throw Error{rc, msg};
}

}  // namespace

extern "C" {

int hmm_set_devices(const int *devices, int n) {
    return guarded([&] {
        require_device();
        if (n < 0) fail(HMM_EINVAL, "negative device count");
        g_devices_from_env = true;  // an explicit call wins over HMMCUDA_DEVICES
        set_devices_locked(devices, n);
    });
}

int hmm_get_devices(int *devices_out, int cap) {
    std::lock_guard<std::mutex> lk(g_entry);
    const int n = (int)g_workers.size();
    for (int k = 0; k < n && k < cap; k++) devices_out[k] = g_workers[k]->device();
    return n;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Viterbi
// ---------------------------------------------------------------------------
namespace {

struct BatchModels {
    std::vector<HostModel> models;
    FaithfulLayout layout;
    char *blob_dev = nullptr;  // owned (cudaMalloc): stable for as long as the entry lives, graphs may refer to it
    uint64_t id = 0;           // identity for the decode-program cache
    uint64_t hash = 0, epoch = 0;
    int dev = 0, C = 0, shared = 0, N = 0, K = 0, nstates = 0;
    int64_t ntrans = 0;
    std::vector<char> raw;     // copy of the caller's arrays: a hit is confirmed byte for byte
    ~BatchModels() {
        if (blob_dev) cudaFree(blob_dev);
    }
};

// ---- model cache -------------------------------------------------------------------------------------------
// Validating a StateMatrix, building its CSR forms, packing and uploading them costs tens of microseconds per
// channel; callers decode many recordings (or one recording many times) with the same model.  Entries are keyed by
// the bytes of the arrays that crossed the ABI (hash first, then compared byte for byte).
thread_local std::vector<std::unique_ptr<BatchModels>> g_models;  // per host thread, like the workspace
thread_local uint64_t g_model_seq = 0, g_call_epoch = 0;
constexpr size_t MAX_MODELS = 48;

inline uint64_t mix64(uint64_t h, const void *data, size_t n) {
    const unsigned char *q = (const unsigned char *)data;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, q + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    uint64_t w = 0;
    if (i < n) memcpy(&w, q + i, n - i);
    h = (h ^ w ^ (uint64_t)n) * 0xBF58476D1CE4E5B9ull;
    return h ^ (h >> 32);
}

BatchModels &get_models(int C, const int16_t *states, int states_shared, int N, int K, int nstates, const hmm_trans *tr,
                        int64_t ntrans, const double *mu, const double *sigma, cudaStream_t st) {
    if (!states || !tr || !mu || !sigma) fail(HMM_EINVAL, "null model array");
    if (C < 1 || N < 1 || K < 1 || nstates < 1 || ntrans < 1) fail(HMM_EINVAL, "bad model sizes");
    int dev = 0;
    HMM_CUDA(cudaGetDevice(&dev));
    const size_t b_st = sizeof(int16_t) * (size_t)N * nstates * (states_shared ? 1 : C);
    const size_t b_tr = sizeof(hmm_trans) * (size_t)ntrans * C, b_mu = sizeof(double) * (size_t)K * N * C;
    const size_t b_sg = sizeof(double) * (size_t)C;
    uint64_t h = 0x1234567ull + (uint64_t)dev;
    h = mix64(h, states, b_st);
    h = mix64(h, tr, b_tr);
    h = mix64(h, mu, b_mu);
    h = mix64(h, sigma, b_sg);
    for (auto &q : g_models) {
        if (q->hash != h || q->dev != dev || q->C != C || q->shared != states_shared || q->N != N || q->K != K ||
            q->nstates != nstates || q->ntrans != ntrans || q->raw.size() != b_st + b_tr + b_mu + b_sg)
            continue;
        const char *r = q->raw.data();
        if (memcmp(r, states, b_st) || memcmp(r + b_st, tr, b_tr) || memcmp(r + b_st + b_tr, mu, b_mu) ||
            memcmp(r + b_st + b_tr + b_mu, sigma, b_sg))
            continue;
        q->epoch = g_call_epoch;
        return *q;
    }
    std::unique_ptr<BatchModels> B(new BatchModels);
    B->models.resize(C);
    for (int c = 0; c < C; c++) {
        const int16_t *stc = states_shared ? states : states + (size_t)c * N * nstates;
        analyse_model(stc, N, K, nstates, tr + (size_t)c * ntrans, ntrans, mu + (size_t)c * K * N, sigma[c],
                      B->models[c]);
        if (c > 0 && (B->models[c].in_ptr != B->models[0].in_ptr || B->models[c].in_src != B->models[0].in_src))
            fail(HMM_EINVAL, "channel %d has a different transition topology than channel 0", c);
    }
    B->layout = faithful_layout(nstates, ntrans);
    // packed into a pinned staging buffer (slot 2) so that the upload is asynchronous; the stream is synchronised
    // before the staging buffer can be reused (every entry point synchronises before it returns)
    const size_t bytes = (size_t)C * B->layout.bytes;
    char *host = (char *)workspace().pinned(2, bytes);
    memset(host, 0, bytes);
    for (int c = 0; c < C; c++) faithful_pack(B->models[c], B->layout, host + (size_t)c * B->layout.bytes);
    cudaError_t e = cudaMalloc((void **)&B->blob_dev, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        B->blob_dev = nullptr;
        fail(HMM_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    HMM_CUDA(cudaMemcpyAsync(B->blob_dev, host, bytes, cudaMemcpyHostToDevice, st));
    HMM_CUDA(cudaStreamSynchronize(st));  // once per model: the staging buffer is free again
    B->raw.resize(b_st + b_tr + b_mu + b_sg);
    memcpy(B->raw.data(), states, b_st);
    memcpy(B->raw.data() + b_st, tr, b_tr);
    memcpy(B->raw.data() + b_st + b_tr, mu, b_mu);
    memcpy(B->raw.data() + b_st + b_tr + b_mu, sigma, b_sg);
    B->hash = h; B->dev = dev; B->C = C; B->shared = states_shared; B->N = N; B->K = K; B->nstates = nstates;
    B->ntrans = ntrans;
    B->id = ++g_model_seq;
    B->epoch = g_call_epoch;
    if (g_models.size() >= MAX_MODELS) {  // least recently used entry of an earlier call (none of its work is in flight)
        size_t victim = g_models.size();
        for (size_t k = 0; k < g_models.size(); k++)
            if (g_models[k]->epoch < g_call_epoch && (victim == g_models.size() || g_models[k]->epoch < g_models[victim]->epoch))
                victim = k;
        if (victim < g_models.size()) {
            // decode programs refer to the model's device blob: they go with it
            ring_drop_programs();
            g_models.erase(g_models.begin() + victim);
        }
    }
    g_models.push_back(std::move(B));
    return *g_models.back();
}

}  // namespace

static void begin_call() {
    g_call_epoch++;
    ring_new_epoch();
}
static void drop_caches() {
    ring_drop_programs();
    g_models.clear();
}

namespace {

// Core decode on device-resident y / x.
void viterbi_core(const double *y_dev, int64_t T, int C, BatchModels &B, int16_t *x_dev, double *ll_host,
                  double *T1_dev, int16_t *T2_dev, int mode, cudaStream_t st, hmm_info *info) {
    const HostModel &M0 = B.models[0];
    bool all_ring = true;
    for (auto &m : B.models) all_ring = all_ring && m.is_ring;
    bool want_trellis = T1_dev || T2_dev;
    int engine;
    if (mode == HMM_MODE_FAITHFUL || want_trellis)  // (models beyond ~5 000 states: HMM_EUNSUPPORTED there)
        engine = HMM_MODE_FAITHFUL;
    else if (mode == HMM_MODE_RING) {
        if (!all_ring) fail(HMM_EUNSUPPORTED, "HMM_MODE_RING requested but the model is not a non-overlap ring model");
        if (!ring_supported(M0, T)) fail(HMM_EUNSUPPORTED, "HMM_MODE_RING: sequence too short or K/N outside the ring engine's range");
        engine = HMM_MODE_RING;
    } else if (mode == HMM_MODE_GENERIC) {
        if (!generic_parallel_supported(M0, T)) fail(HMM_EUNSUPPORTED, "HMM_MODE_GENERIC: more than 32767 states");
        engine = HMM_MODE_GENERIC;
    } else
        // every ring model the time-parallel engine supports goes to it (T >= 2048): even a 20 000-sample decode
        // (config 1) is two orders of magnitude faster there than in the sequential per-state engine, and as exact;
        // any other StateMatrix (overlap models, N > 7) takes the time-parallel per-state engine when it is long enough
        engine = (all_ring && ring_supported(M0, T)) ? HMM_MODE_RING
                 : generic_parallel_preferred(M0, T) ? HMM_MODE_GENERIC
                                                      : HMM_MODE_FAITHFUL;
    if (info) info->engine = engine;
    if (engine == HMM_MODE_GENERIC) {
        for (int ch = 0; ch < C; ch++)
            generic_parallel_viterbi_run(y_dev + (size_t)ch * T, T, B.layout, B.blob_dev + (size_t)ch * B.layout.bytes,
                                         B.models[ch < (int)B.models.size() ? ch : 0], x_dev + (size_t)ch * T,
                                         ll_host ? ll_host + ch : nullptr, st, info);
        return;
    }
    if (engine == HMM_MODE_FAITHFUL) {
        double *ll_dev = nullptr;
        if (ll_host) ll_dev = (double *)workspace().get(Workspace::SCRATCH, sizeof(double) * C);
        faithful_viterbi_run(y_dev, T, T, C, B.layout, B.blob_dev, M0, x_dev, T, ll_dev, T1_dev, T2_dev,
                             want_trellis ? T : 0, false, nullptr, st, info);
        if (ll_host) d2h(ll_host, ll_dev, sizeof(double) * C, st);
    } else  // synchronises the stream once and leaves ll / the repair counts in ll_host / info
        ring_viterbi_run(y_dev, T, T, C, B.models, B.layout, B.blob_dev, B.id, x_dev, T, ll_host, st, info, nullptr);
}

// traceback state codes carry a chain entry time (code = 8*t0 + neuron, -1 = noise)
__global__ void shift_state_kernel(const long long *src, long long *dst, long long delta) {
    long long v = *src;
    *dst = v >= 0 ? v + delta : v;
}

// ---------------------------------------------------------------------------
// Host-pointer decode as a software pipeline on one GPU: the recording is cut into
// segments (time shards with ghost chunks, exactly the hmm_vshard_* mechanism); segment
// k is decoded while segment k+1 is still crossing PCIe, and its part of x goes back
// while later segments are decoded.  Boundaries are verified left to right (forward)
// and against the right neighbour (traceback), so the result is the exact decode.
// ---------------------------------------------------------------------------
struct PipeSeg {
    int64_t lb, le, mb, me;  // local span [lb, le) incl. ghosts, main span [mb, me)
    int c_main0, c_main1;
    std::unique_ptr<VitPlan> plan;
    cudaEvent_t ev_copy = nullptr, ev_x = nullptr, ev_xd = nullptr;  // y arrived / x final on device / x on host
};

// ---- pageable callers (a Julia Array, a plain numpy array) -------------------------------------------------
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at ~10 GB/s and does not
// overlap with anything.  The pipelined decode therefore copies a pageable recording into its own pinned staging
// buffer with a few host threads, block by block in time order, while earlier blocks are already on their way to
// the device; x travels back through pinned staging the same way.
static bool host_is_pinned(const void *q) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, q) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int staging_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    int n = hc ? (int)(hc / 2) : 4;
    return n < 2 ? 2 : (n > 8 ? 8 : n);
}

struct HostStager {  // parallel, in-order memcpy of src[0, n) bytes into dst, in blocks
    std::vector<std::thread> th;
    std::unique_ptr<std::atomic<unsigned char>[]> done;
    std::atomic<int64_t> next{0};
    int64_t nb = 0, blk = 0, n = 0;
    const char *src = nullptr;
    char *dst = nullptr;
    void start(const void *s, void *d, int64_t bytes, int64_t block_bytes, int nthreads) {
        src = (const char *)s;
        dst = (char *)d;
        n = bytes;
        blk = block_bytes;
        nb = (bytes + blk - 1) / blk;
        done.reset(new std::atomic<unsigned char>[(size_t)(nb > 0 ? nb : 1)]);
        for (int64_t j = 0; j < nb; j++) done[(size_t)j].store(0, std::memory_order_relaxed);
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([this] {
                for (;;) {
                    const int64_t j = next.fetch_add(1);
                    if (j >= nb) return;
                    const int64_t o = j * blk, len = std::min<int64_t>(blk, n - o);
                    memcpy(dst + o, src + o, (size_t)len);
                    done[(size_t)j].store(1, std::memory_order_release);
                }
            });
    }
    void wait_bytes(int64_t upto) {  // until [0, upto) has been copied
        const int64_t jb = std::min<int64_t>(nb, (upto + blk - 1) / blk);
        for (int64_t j = 0; j < jb; j++)
            while (!done[(size_t)j].load(std::memory_order_acquire)) std::this_thread::yield();
    }
    void join() {
        for (auto &t : th)
            if (t.joinable()) t.join();
        th.clear();
    }
    ~HostStager() { join(); }
};

bool pipeline_wanted(const HostModel &M0, int64_t T, int C) {
    if (getenv("HMMCUDA_NO_PIPELINE") && atoi(getenv("HMMCUDA_NO_PIPELINE"))) return false;
    return C == 1 && T >= (int64_t)1 << 22 && ring_config().chunk_len == 0 && ring_config().warmup == 0 && M0.is_ring;
}

void viterbi_host_pipelined(const double *y, int64_t T, BatchModels &B, int16_t *x_out, double *ll_out, hmm_info *info) {
    Workspace &ws = workspace();
    cudaStream_t sc = main_stream(), sh = copy_stream(), sd = out_stream();
    const HostModel &M0 = B.models[0];
    const int L = M0.K - 1;
    int S = 8;
    if (const char *e = getenv("HMMCUDA_PIPE_SEGMENTS")) S = std::max(2, atoi(e));
    int dev = 0, sms = 148;
    HMM_CUDA(cudaGetDevice(&dev));
    HMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one wave of chunks per segment; short warm-up (every boundary is verified anyway)
    const int64_t W = 256;
    int64_t Lc = (T / S + (int64_t)sms * 16 - 1) / ((int64_t)sms * 16);
    Lc = ((Lc + 255) / 256) * 256;
    if (Lc < 1024) Lc = 1024;
    if (Lc < ((L + 32 + 255) / 256) * 256 + W) Lc = ((L + 32 + 255) / 256) * 256 + W;
    const int64_t nchunks_tot = (T + Lc - 1) / Lc;
    if (nchunks_tot < 2 * S) S = (int)std::max<int64_t>(1, nchunks_tot / 2);
    std::vector<PipeSeg> seg(S);
    {
        int64_t q = nchunks_tot / S, r = nchunks_tot % S, c0 = 0;
        for (int k = 0; k < S; k++) {
            int64_t c1 = c0 + q + (k < r ? 1 : 0);
            seg[k].mb = c0 * Lc;
            seg[k].me = std::min<int64_t>(T, c1 * Lc);
            seg[k].lb = std::max<int64_t>(0, seg[k].mb - Lc);
            seg[k].le = std::min<int64_t>(T, seg[k].me + Lc);
            c0 = c1;
        }
        // a last main span shorter than the engine's look-back joins the previous segment
        if (S > 1 && T - seg[S - 1].mb < 2 * RING_Q_MIN) {
            seg[S - 2].me = T;
            seg[S - 2].le = T;
            seg.pop_back();
            S--;
        }
    }
    double *y_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T);
    int16_t *x_dev = (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * (size_t)T);
    const bool no_staging = getenv("HMMCUDA_NO_STAGING") && atoi(getenv("HMMCUDA_NO_STAGING"));
    const bool y_pageable = !no_staging && !host_is_pinned(y), x_pageable = !no_staging && !host_is_pinned(x_out);
    const double *y_src = y;
    int16_t *x_dst = x_out;
    HostStager stager;
    std::thread drainer;
    std::atomic<int> x_issued{0}, drain_abort{0};
    if (y_pageable) {
        double *ys = (double *)ws.pinned(4, sizeof(double) * (size_t)T);
        stager.start(y, ys, (int64_t)sizeof(double) * T, (int64_t)1 << 20, staging_threads());
        y_src = ys;
    }
    if (x_pageable) x_dst = (int16_t *)ws.pinned(5, sizeof(int16_t) * (size_t)T);
    size_t arena_cap = 1 << 20;
    const int bvec = 1 + M0.N * L;
    for (auto &g : seg) {
        const size_t Tl = (size_t)(g.le - g.lb), nch = (size_t)((Tl + Lc - 1) / Lc);
        arena_cap += Tl * 4 + Tl / 8 + 64 + nch * (2 * 8 * (size_t)bvec + 64) + (size_t)M0.nstates * (L + 1) * 10 + (256 << 10);
    }
    arena_cap += arena_cap / 8;
    char *arena = (char *)ws.get(Workspace::TRACE, arena_cap);
    double *ll_dev = (double *)ws.get(Workspace::SCRATCH, sizeof(double) * 1024);  // [0] ll, [8..600) partials, then counters
    size_t arena_off = 0;
    for (auto &g : seg) {
        HMM_CUDA(cudaEventCreateWithFlags(&g.ev_copy, cudaEventDisableTiming));
        HMM_CUDA(cudaEventCreateWithFlags(&g.ev_x, cudaEventDisableTiming));
        HMM_CUDA(cudaEventCreateWithFlags(&g.ev_xd, cudaEventDisableTiming));
    }
    auto cleanup = [&] {
        for (auto &g : seg) {
            if (g.ev_copy) cudaEventDestroy(g.ev_copy);
            if (g.ev_x) cudaEventDestroy(g.ev_x);
            if (g.ev_xd) cudaEventDestroy(g.ev_xd);
        }
    };
    try {
        // copy boundaries are shifted right by one chunk so that segment k's ghost arrives with it
        auto copy_end = [&](int k) { return k == S - 1 ? T : std::min<int64_t>(T, seg[k].me + Lc); };
        auto issue_copy = [&](int k) {
            const int64_t a = k == 0 ? 0 : copy_end(k - 1), b = copy_end(k);
            if (y_pageable) stager.wait_bytes((int64_t)sizeof(double) * b);
            if (b > a) HMM_CUDA(cudaMemcpyAsync(y_dev + a, y_src + a, sizeof(double) * (size_t)(b - a), cudaMemcpyHostToDevice, sh));
            HMM_CUDA(cudaEventRecord(seg[k].ev_copy, sh));
        };
        auto issue_x = [&](int k) {
            HMM_CUDA(cudaStreamWaitEvent(sd, seg[k].ev_x, 0));
            HMM_CUDA(cudaMemcpyAsync(x_dst + seg[k].mb, x_dev + seg[k].mb, sizeof(int16_t) * (size_t)(seg[k].me - seg[k].mb),
                                     cudaMemcpyDeviceToHost, sd));
            HMM_CUDA(cudaEventRecord(seg[k].ev_xd, sd));
            x_issued.store(k + 1, std::memory_order_release);
        };
        auto link_trace = [&](int k) {  // own_start of segment k's first main chunk -> segment k-1's right ghost
            shift_state_kernel<<<1, 1, 0, sc>>>(seg[k].plan->own_start_ptr(seg[k].c_main0),
                                                seg[k - 1].plan->own_start_ptr(seg[k - 1].c_main1),
                                                8 * (long long)(seg[k].lb - seg[k - 1].lb));
            HMM_CUDA(cudaGetLastError());
        };
        if (x_pageable) {
            // a drainer thread moves each segment's x from pinned staging to the caller's array as soon as it landed
            int dev_id = 0;
            HMM_CUDA(cudaGetDevice(&dev_id));
            drainer = std::thread([&, dev_id] {
                cudaSetDevice(dev_id);
                for (int k = 0; k < S; k++) {
                    while (x_issued.load(std::memory_order_acquire) <= k) {
                        if (drain_abort.load(std::memory_order_acquire)) return;
                        std::this_thread::yield();
                    }
                    if (cudaEventSynchronize(seg[k].ev_xd) != cudaSuccess) return;
                    memcpy(x_out + seg[k].mb, x_dst + seg[k].mb, sizeof(int16_t) * (size_t)(seg[k].me - seg[k].mb));
                }
            });
        }
        issue_copy(0);
        for (int k = 0; k < S; k++) {
            PipeSeg &g = seg[k];
            const bool first = k == 0, last = k == S - 1;
            g.plan.reset(new VitPlan);
            g.plan->use_arena(arena + arena_off, arena_cap - arena_off);
            const int64_t Tl = g.le - g.lb;
            g.plan->build(y_dev + g.lb, Tl, Tl, 1, B.models, B.layout, B.blob_dev, x_dev + g.lb, Tl, Lc, W, first, last, sc);
            arena_off += (g.plan->arena_bytes_used() + 255) & ~size_t(255);
            g.plan->set_x_window(g.mb - g.lb, g.me - g.lb);
            g.c_main0 = (int)((g.mb - g.lb) / Lc);
            g.c_main1 = last ? g.plan->nchunks() : (int)((g.me - g.lb) / Lc);
            HMM_CUDA(cudaStreamWaitEvent(sc, g.ev_copy, 0));
            g.plan->forward(sc, nullptr);
            if (!first)  // the neighbour's true end vector replaces the left ghost chunk's speculative one
                HMM_CUDA(cudaMemcpyAsync(g.plan->eb_ptr(g.c_main0 - 1), seg[k - 1].plan->eb_ptr(seg[k - 1].c_main1 - 1),
                                         sizeof(double) * bvec, cudaMemcpyDeviceToDevice, sc));
            g.plan->verify_fwd(sc);
            g.plan->trace(sc);
            if (!first) {
                link_trace(k);
                seg[k - 1].plan->verify_trace(sc);
                HMM_CUDA(cudaEventRecord(seg[k - 1].ev_x, sc));
            }
            if (info) info->kernel_launches += first ? 8 : 9;
            if (!last) issue_copy(k + 1);
            if (!first) issue_x(k - 1);
        }
        seg[S - 1].plan->verify_trace(sc);
        HMM_CUDA(cudaEventRecord(seg[S - 1].ev_x, sc));
        issue_x(S - 1);
        // traceback repairs inside a segment can, very rarely, move its own start state after the left
        // neighbour was already verified against it: detect and redo the chain right to left
        int tot_f = 0, tot_b = 0, nch = 0;
        for (int k = 0; k < S; k++) {
            int f = 0, b = 0;
            seg[k].plan->read_counters(sc, &f, &b);
            tot_f += f;
            tot_b += b;
            nch += seg[k].c_main1 - seg[k].c_main0;
        }
        if (tot_b > 0) {
            HMM_CUDA(cudaStreamSynchronize(sd));
            for (int k = S - 1; k >= 1; k--) {
                link_trace(k);
                seg[k - 1].plan->verify_trace(sc);
            }
            HMM_CUDA(cudaMemcpyAsync(x_dst, x_dev, sizeof(int16_t) * (size_t)T, cudaMemcpyDeviceToHost, sc));
        }
        if (ll_out) {
            ring_path_ll_run(y_dev, T, B.layout, B.blob_dev, M0, x_dev, ll_dev, ll_dev + 8, sc);
            HMM_CUDA(cudaMemcpyAsync(ll_out, ll_dev, sizeof(double), cudaMemcpyDeviceToHost, sc));
            if (info) info->kernel_launches += 2;
        }
        HMM_CUDA(cudaStreamSynchronize(sc));
        HMM_CUDA(cudaStreamSynchronize(sd));
        HMM_CUDA(cudaStreamSynchronize(sh));
        stager.join();
        if (drainer.joinable()) drainer.join();
        if (x_pageable && tot_b > 0) {  // the traceback was redone after the segments had been drained: copy x again
            HostStager back;
            back.start(x_dst, x_out, (int64_t)sizeof(int16_t) * T, (int64_t)4 << 20, staging_threads());
            back.join();
        }
        if (info) {
            info->engine = HMM_MODE_RING;
            info->n_chunks = nch;
            info->fwd_repaired = tot_f;
            info->bwd_repaired = tot_b;
        }
    } catch (...) {
        drain_abort.store(1);
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(sd);
        cudaStreamSynchronize(sh);
        if (drainer.joinable()) drainer.join();
        stager.join();
        cleanup();
        throw;
    }
    cleanup();
}

// ---- multi-device: a batch of channels, contiguous blocks of channels per device --------------------------------
void viterbi_batch_multidev(const double *y, int64_t T, int C, const int16_t *states, int states_shared, int N, int K,
                            int nstates, const hmm_trans *tr, int64_t ntrans, const double *mu, const double *sigma,
                            int16_t *x_out, double *ll_out, int mode, hmm_info *info) {
    const int nd = (int)std::min<size_t>(g_workers.size(), (size_t)C);
    std::vector<hmm_info> infos((size_t)nd);
    std::vector<std::future<std::pair<int, std::string>>> futs;
    for (int k = 0; k < nd; k++) {
        const int c0 = (int)((int64_t)C * k / nd), c1 = (int)((int64_t)C * (k + 1) / nd);
        if (c1 <= c0) continue;
        hmm_info *ik = &infos[(size_t)k];
        futs.push_back(g_workers[(size_t)k]->run([=] {
            return hmm_viterbi_batch_f64(y + (size_t)c0 * T, T, c1 - c0, states_shared ? states : states + (size_t)c0 * N * nstates,
                                         states_shared, N, K, nstates, tr + (size_t)c0 * ntrans, ntrans,
                                         mu + (size_t)c0 * K * N, sigma + c0, x_out + (size_t)c0 * T,
                                         ll_out ? ll_out + c0 : nullptr, mode, ik);
        }));
    }
    join_all(futs);
    if (info) {
        memset(info, 0, sizeof *info);
        for (auto &q : infos) {
            info->engine = q.engine;
            info->n_chunks += q.n_chunks;
            info->fwd_repaired += q.fwd_repaired;
            info->bwd_repaired += q.bwd_repaired;
            info->kernel_launches += q.kernel_launches;
            info->device_ms = std::max(info->device_ms, q.device_ms);
        }
    }
}

// ---- multi-device: ONE long recording as time shards, one per device, peer-memory boundary exchange -----------
struct ShardSpan {
    int64_t lb, le, mb, me;
};
std::vector<ShardSpan> plan_shards(int64_t T, int n, int64_t Lc, int64_t W) {
    int64_t nchunks = (T + Lc - 1) / Lc;
    const int64_t tail = T - (nchunks - 1) * Lc;
    if (nchunks > 1 && tail < std::min<int64_t>(W + 128, Lc)) nchunks--;  // a short tail joins the chunk before it
    if (n > nchunks / 2) n = (int)std::max<int64_t>(1, nchunks / 2);      // at least two chunks per shard
    std::vector<ShardSpan> out((size_t)n);
    const int64_t q = nchunks / n, r = nchunks % n;
    int64_t c0 = 0;
    for (int k = 0; k < n; k++) {
        const int64_t c1 = c0 + q + (k < r ? 1 : 0);
        out[(size_t)k].mb = c0 * Lc;
        out[(size_t)k].me = k == n - 1 ? T : c1 * Lc;
        out[(size_t)k].lb = std::max<int64_t>(0, out[(size_t)k].mb - Lc);
        out[(size_t)k].le = std::min<int64_t>(T, out[(size_t)k].me + Lc);
        c0 = c1;
    }
    return out;
}

// The shard handles (one per device: buffers, decode plan, CUDA graph, opened peer blocks) are kept between calls and
// re-used when the next recording has the same length and model: setting them up costs ~200 ms, a decode ~1 ms.
struct ShardSet {
    std::vector<hmm_vshard *> sh;
    std::vector<ShardSpan> spans;
    std::vector<int> devs;
    int64_t T = 0, Lc = 0, W = 0;
    uint64_t model_hash = 0;
};
std::unique_ptr<ShardSet> g_shardset;  // under the entry lock

void drop_shardset() {
    if (!g_shardset) return;
    std::vector<std::future<std::pair<int, std::string>>> f2;
    for (size_t k = 0; k < g_shardset->sh.size() && k < g_workers.size(); k++)
        if (g_shardset->sh[k]) {
            hmm_vshard *h = g_shardset->sh[k];
            f2.push_back(g_workers[k]->run([h] { return hmm_vshard_destroy(h); }));
        }
    for (auto &f : f2) f.get();
    g_shardset.reset();
}

bool viterbi_timeshard_multidev(const double *y, int64_t T, const int16_t *states, int N, int K, int nstates,
                                const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                                double *ll_out, hmm_info *info) {
    uint64_t mh = 0x51ed27ull;
    mh = mix64(mh, states, sizeof(int16_t) * (size_t)N * nstates);
    mh = mix64(mh, tr, sizeof(hmm_trans) * (size_t)ntrans);
    mh = mix64(mh, mu, sizeof(double) * (size_t)K * N);
    mh = mix64(mh, &sigma, sizeof sigma);
    std::vector<int> devs;
    for (auto &w : g_workers) devs.push_back(w->device());
    std::vector<std::future<std::pair<int, std::string>>> futs;
    const bool hit = g_shardset && g_shardset->T == T && g_shardset->model_hash == mh && g_shardset->devs == devs;
    if (!hit) {
        drop_shardset();
        int64_t Lc = 0, W = 0;
        {
            HostModel M;
            M.N = N;
            M.K = K;
            M.nstates = nstates;
            M.is_ring = true;
            ring_default_chunking(M, T, 1, (int)g_workers.size(), &Lc, &W);
            W = 256;  // short chunks per GPU: a short speculative warm-up (every boundary is verified anyway)
        }
        std::unique_ptr<ShardSet> S(new ShardSet);
        S->spans = plan_shards(T, (int)g_workers.size(), Lc, W);
        const int n = (int)S->spans.size();
        if (n < 2) return false;
        S->sh.assign((size_t)n, nullptr);
        S->devs = devs; S->T = T; S->Lc = Lc; S->W = W; S->model_hash = mh;
        std::vector<void *> blocks((size_t)n, nullptr);
        g_shardset = std::move(S);
        ShardSet &G = *g_shardset;
        try {
            for (int k = 0; k < n; k++) {
                const ShardSpan sp = G.spans[(size_t)k];
                hmm_vshard **hk = &G.sh[(size_t)k];
                void **bk = &blocks[(size_t)k];
                futs.push_back(g_workers[(size_t)k]->run([=] {
                    int rc = hmm_vshard_create(y + sp.lb, 1, sp.lb, sp.le, sp.mb, sp.me, T, Lc, W, states, N, K, nstates, tr,
                                               ntrans, mu, sigma, hk);
                    if (rc) return rc;
                    return hmm_vshard_p2p_init(*hk, k, n, nullptr, bk);
                }));
            }
            join_all(futs);
            for (int k = 0; k < n; k++) {
                hmm_vshard *h = G.sh[(size_t)k];
                void *const *bp = blocks.data();
                futs.push_back(g_workers[(size_t)k]->run([=] { return hmm_vshard_p2p_attach(h, nullptr, bp); }));
            }
            join_all(futs);
        } catch (...) {
            drop_shardset();
            throw;
        }
    }
    ShardSet &G = *g_shardset;
    const int n = (int)G.sh.size();
    std::vector<double> ll((size_t)n, 0.0);
    std::vector<int32_t> bad((size_t)n, 0);
    try {
        for (int k = 0; k < n; k++) {
            hmm_vshard *h = G.sh[(size_t)k];
            const ShardSpan sp = G.spans[(size_t)k];
            futs.push_back(g_workers[(size_t)k]->run([=] {
                if (hit) {  // (a fresh handle already holds this recording)
                    int rc = hmm_vshard_set_y(h, y + sp.lb);
                    if (rc) return rc;
                }
                return hmm_vshard_p2p_launch(h, nullptr);
            }));
        }
        join_all(futs);  // every shard has launched its decode and its summary stores: the judges cannot wait in vain
        for (int k = 0; k < n; k++) {
            hmm_vshard *h = G.sh[(size_t)k];
            const ShardSpan sp = G.spans[(size_t)k];
            double *lk = &ll[(size_t)k];
            int32_t *bk = &bad[(size_t)k];
            futs.push_back(g_workers[(size_t)k]->run([=] {
                int rc = hmm_vshard_p2p_finish(h, lk, bk);
                if (rc) return rc;
                return hmm_vshard_finish(h, x_out + sp.mb, 0, nullptr);  // x of the main span -> the caller's array
            }));
        }
        join_all(futs);
    } catch (...) {
        drop_shardset();
        throw;
    }
    for (int k = 0; k < n; k++)
        if (bad[(size_t)k] != 0) return false;  // a ghost chunk guessed wrong (not seen on real data): exact single-GPU decode
    if (ll_out) *ll_out = ll[0];
    if (info) {
        memset(info, 0, sizeof *info);
        info->engine = HMM_MODE_RING;
        info->n_chunks = (int32_t)((T + G.Lc - 1) / G.Lc);
        info->kernel_launches = 8 * (int64_t)n;
    }
    return true;
}

int viterbi_host(const double *y, int64_t T, int C, const int16_t *states, int states_shared, int N, int K,
                 int nstates, const hmm_trans *tr, int64_t ntrans, const double *mu, const double *sigma,
                 int16_t *x_out, double *ll_out, int16_t *T2_out, double *T1_out, int mode, hmm_info *info) {
    return guarded([&] {
        if (info) memset(info, 0, sizeof *info);
        if (!y || !x_out || !sigma) fail(HMM_EINVAL, "null y / x_out / sigma");
        if (T < 1 || C < 1) fail(HMM_EINVAL, "T and C must be positive");
        if ((T1_out || T2_out) && C != 1) fail(HMM_EINVAL, "trellis output is single-channel");
        require_device();
        precision_from_env_once();
        // ---- several devices selected (hmm_set_devices / HMMCUDA_DEVICES): split the call over them ----
        if (!t_worker) {
            devices_from_env_once();
            if (g_workers.size() > 1 && !T1_out && !T2_out && mode != HMM_MODE_FAITHFUL && mode != HMM_MODE_GENERIC) {
                if (C > 1) {
                    viterbi_batch_multidev(y, T, C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, x_out,
                                           ll_out, mode, info);
                    return;
                }
                if (T >= ((int64_t)1 << 23)) {
                    HostModel M;
                    analyse_model(states, N, K, nstates, tr, ntrans, mu, sigma[0], M);
                    if (M.is_ring && ring_supported(M, T) &&
                        viterbi_timeshard_multidev(y, T, states, N, K, nstates, tr, ntrans, mu, sigma[0], x_out, ll_out, info))
                        return;
                }
            }
        }
        cudaStream_t st = main_stream();
        Workspace &ws = workspace();
        Timer tall(st);
        tall.start();
        BatchModels &B = get_models(C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, st);
        if (!T1_out && !T2_out && mode != HMM_MODE_FAITHFUL && mode != HMM_MODE_GENERIC && ring_supported(B.models[0], T) &&
            pipeline_wanted(B.models[0], T, C)) {
            viterbi_host_pipelined(y, T, B, x_out, ll_out, info);
            tall.stop();
            if (info) info->device_ms = tall.ms();
            return;
        }
        double *y_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T * C);
        int16_t *x_dev = (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * (size_t)T * C);
        {
            // Many long channels (config 4): pipeline by channel -- channel c is decoded while channel
            // c+1 crosses PCIe and channel c-1's x travels back.
            bool all_ring = true;
            for (auto &m : B.models) all_ring = all_ring && m.is_ring;
            const bool no_pipe = getenv("HMMCUDA_NO_PIPELINE") && atoi(getenv("HMMCUDA_NO_PIPELINE"));
            if (C > 1 && !T1_out && !T2_out && mode != HMM_MODE_FAITHFUL && mode != HMM_MODE_GENERIC && all_ring && ring_supported(B.models[0], T) &&
                T >= 262144 && !no_pipe) {
                cudaStream_t sh = copy_stream(), sd = out_stream();
                std::vector<cudaEvent_t> evc(C), evx(C);
                for (int c = 0; c < C; c++) {
                    HMM_CUDA(cudaEventCreateWithFlags(&evc[c], cudaEventDisableTiming));
                    HMM_CUDA(cudaEventCreateWithFlags(&evx[c], cudaEventDisableTiming));
                }
                std::vector<RingPending> pend;
                auto copy_in = [&](int c) {
                    HMM_CUDA(cudaMemcpyAsync(y_dev + (size_t)c * T, y + (size_t)c * T, sizeof(double) * (size_t)T,
                                             cudaMemcpyHostToDevice, sh));
                    HMM_CUDA(cudaEventRecord(evc[c], sh));
                };
                auto copy_out = [&](int c) {
                    HMM_CUDA(cudaStreamWaitEvent(sd, evx[c], 0));
                    HMM_CUDA(cudaMemcpyAsync(x_out + (size_t)c * T, x_dev + (size_t)c * T, sizeof(int16_t) * (size_t)T,
                                             cudaMemcpyDeviceToHost, sd));
                };
                try {
                    copy_in(0);
                    for (int c = 0; c < C; c++) {
                        HMM_CUDA(cudaStreamWaitEvent(st, evc[c], 0));
                        std::vector<HostModel> one(1, B.models[c]);
                        ring_viterbi_run(y_dev + (size_t)c * T, T, T, 1, one, B.layout, B.blob_dev + (size_t)c * B.layout.bytes,
                                         B.id, x_dev + (size_t)c * T, T, ll_out ? ll_out + c : nullptr, st, nullptr, &pend);
                        HMM_CUDA(cudaEventRecord(evx[c], st));
                        if (c + 1 < C) copy_in(c + 1);
                        if (c >= 1) copy_out(c - 1);
                    }
                    copy_out(C - 1);
                    tall.stop();
                    HMM_CUDA(cudaStreamSynchronize(st));
                    HMM_CUDA(cudaStreamSynchronize(sd));
                    HMM_CUDA(cudaStreamSynchronize(sh));
                    ring_collect(pend);
                } catch (...) {
                    cudaDeviceSynchronize();
                    for (int c = 0; c < C; c++) {
                        cudaEventDestroy(evc[c]);
                        cudaEventDestroy(evx[c]);
                    }
                    throw;
                }
                for (int c = 0; c < C; c++) {
                    cudaEventDestroy(evc[c]);
                    cudaEventDestroy(evx[c]);
                }
                if (info) {
                    info->engine = HMM_MODE_RING;
                    info->device_ms = tall.ms();
                    info->kernel_launches = (int64_t)C * 5;
                }
                return;
            }
        }
        double *T1_dev = T1_out ? (double *)ws.get(Workspace::T1, sizeof(double) * (size_t)T * nstates) : nullptr;
        int16_t *T2_dev = T2_out ? (int16_t *)ws.get(Workspace::T2, sizeof(int16_t) * (size_t)T * nstates) : nullptr;
        h2d(y_dev, y, sizeof(double) * (size_t)T * C, st);
        Timer tk(st);
        tk.start();
        viterbi_core(y_dev, T, C, B, x_dev, ll_out, T1_dev, T2_dev, mode, st, info);
        tk.stop();
        d2h(x_out, x_dev, sizeof(int16_t) * (size_t)T * C, st);
        if (T1_out) d2h(T1_out, T1_dev, sizeof(double) * (size_t)T * nstates, st);
        if (T2_out) d2h(T2_out, T2_dev, sizeof(int16_t) * (size_t)T * nstates, st);
        tall.stop();
        HMM_CUDA(cudaStreamSynchronize(st));
        if (info) {
            info->device_ms = tall.ms();
            info->kernel_ms = tk.ms();
        }
    });
}

}  // namespace

extern "C" {

int hmm_viterbi_f64(const double *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                    double *ll_out, int16_t *T2_out, double *T1_out) {
    return viterbi_host(y, T, 1, states, 1, N, K, nstates, tr, ntrans, mu, &sigma, x_out, ll_out, T2_out, T1_out,
                        HMM_MODE_AUTO, nullptr);
}

int hmm_viterbi_ex_f64(const double *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                       double *ll_out, int16_t *T2_out, double *T1_out, int32_t mode, hmm_info *info) {
    return viterbi_host(y, T, 1, states, 1, N, K, nstates, tr, ntrans, mu, &sigma, x_out, ll_out, T2_out, T1_out,
                        mode, info);
}

int hmm_viterbi_batch_f64(const double *y, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                          int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                          const double *mu, const double *sigma, int16_t *x_out, double *ll_out, int32_t mode,
                          hmm_info *info) {
    return viterbi_host(y, T, C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, x_out, ll_out, nullptr,
                        nullptr, mode, info);
}

int hmm_viterbi_dev_f64(const double *y_dev, int64_t T, int32_t C, const int16_t *states, int32_t states_shared,
                        int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans,
                        const double *mu, const double *sigma, int16_t *x_dev, double *ll_out, int32_t mode,
                        hmm_info *info) {
    return guarded([&] {
        if (info) memset(info, 0, sizeof *info);
        if (!y_dev || !x_dev || !sigma) fail(HMM_EINVAL, "null y_dev / x_dev / sigma");
        if (T < 1 || C < 1) fail(HMM_EINVAL, "T and C must be positive");
        require_device();
        precision_from_env_once();
        cudaStream_t st = main_stream();
        BatchModels &B = get_models(C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, st);
        Timer tk(st);
        tk.start();
        viterbi_core(y_dev, T, C, B, x_dev, ll_out, nullptr, nullptr, mode, st, info);
        tk.stop();
        HMM_CUDA(cudaStreamSynchronize(st));
        if (info) info->device_ms = info->kernel_ms = tk.ms();
    });
}

// ---------------------------------------------------------------------------
// FP32 mode: Float32 recordings (half the PCIe / HBM bytes per sample) and the FIR of the ring decode in FP32.
// The recording is widened to FP64 on the device (exact), so every downstream kernel is the FP64 one; what
// differs from hmm_viterbi_f64 on the same values is the rounding of the FIR (~1e-4 absolute on scores of O(100)):
// T1 / ll agree to <= 1e-4 relative, x can differ where a decision margin is below that (rate reported by tests).
// ---------------------------------------------------------------------------
__global__ void widen_f32_kernel(const float *__restrict__ src, double *__restrict__ dst, int64_t n) {
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 4 <= n && ((reinterpret_cast<uintptr_t>(src + i4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst + i4) & 15) == 0)) {
        const float4 v = *reinterpret_cast<const float4 *>(src + i4);
        *reinterpret_cast<double2 *>(dst + i4) = make_double2((double)v.x, (double)v.y);
        *reinterpret_cast<double2 *>(dst + i4 + 2) = make_double2((double)v.z, (double)v.w);
    } else
        for (int64_t i = i4; i < n && i < i4 + 4; i++) dst[i] = (double)src[i];
}

struct PrecisionScope {  // FP32 mode for the duration of one call
    int saved;
    PrecisionScope() : saved(ring_config().precision) { ring_config().precision = HMM_PREC_F32; }
    ~PrecisionScope() { ring_config().precision = saved; }
};

static void viterbi_f32_body(const float *y, bool y_on_device, int64_t T, int C, const int16_t *states, int states_shared,
                             int N, int K, int nstates, const hmm_trans *tr, int64_t ntrans, const double *mu,
                             const double *sigma, int16_t *x_out, bool x_on_device, double *ll_out, int mode,
                             hmm_info *info) {
    if (info) memset(info, 0, sizeof *info);
    if (!y || !x_out || !sigma) fail(HMM_EINVAL, "null y / x_out / sigma");
    if (T < 1 || C < 1) fail(HMM_EINVAL, "T and C must be positive");
    require_device();
    PrecisionScope fp32;
    cudaStream_t st = main_stream();
    Workspace &ws = workspace();
    Timer tall(st);
    tall.start();
    BatchModels &B = get_models(C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, st);
    const size_t n = (size_t)T * C;
    const float *src = y;
    if (!y_on_device) {
        float *yf = (float *)ws.get(Workspace::FWDF, sizeof(float) * n);
        h2d(yf, y, sizeof(float) * n, st);
        src = yf;
    }
    double *y_dev = (double *)ws.get(Workspace::Y, sizeof(double) * n);
    widen_f32_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(src, y_dev, (int64_t)n);
    HMM_CUDA(cudaGetLastError());
    int16_t *x_dev = x_on_device ? x_out : (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * n);
    viterbi_core(y_dev, T, C, B, x_dev, ll_out, nullptr, nullptr, mode, st, info);
    if (!x_on_device) d2h(x_out, x_dev, sizeof(int16_t) * n, st);
    tall.stop();
    HMM_CUDA(cudaStreamSynchronize(st));
    if (info) {
        info->device_ms = tall.ms();
        info->kernel_launches += 1;
    }
}

int hmm_viterbi_ex_f32(const float *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out,
                       double *ll_out, int32_t mode, hmm_info *info) {
    return guarded([&] {
        viterbi_f32_body(y, false, T, 1, states, 1, N, K, nstates, tr, ntrans, mu, &sigma, x_out, false, ll_out, mode, info);
    });
}
int hmm_viterbi_f32(const float *y, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, int16_t *x_out, double *ll_out) {
    return hmm_viterbi_ex_f32(y, T, states, N, K, nstates, tr, ntrans, mu, sigma, x_out, ll_out, HMM_MODE_AUTO, nullptr);
}
int hmm_viterbi_batch_f32(const float *y, int64_t T, int32_t C, const int16_t *states, int32_t states_shared, int32_t N,
                          int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans, const double *mu,
                          const double *sigma, int16_t *x_out, double *ll_out, int32_t mode, hmm_info *info) {
    return guarded([&] {
        viterbi_f32_body(y, false, T, C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, x_out, false, ll_out,
                         mode, info);
    });
}
int hmm_viterbi_dev_f32(const float *y_dev, int64_t T, int32_t C, const int16_t *states, int32_t states_shared, int32_t N,
                        int32_t K, int32_t nstates, const hmm_trans *tr, int64_t ntrans, const double *mu,
                        const double *sigma, int16_t *x_dev, double *ll_out, int32_t mode, hmm_info *info) {
    return guarded([&] {
        viterbi_f32_body(y_dev, true, T, C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, x_dev, true, ll_out,
                         mode, info);
    });
}

// ---------------------------------------------------------------------------
// time-sharded decode of one recording (config 5)
// ---------------------------------------------------------------------------
struct hmm_vshard {
    VitPlan plan;
    std::vector<HostModel> models;
    FaithfulLayout FL;
    char *blob_dev = nullptr;
    double *y_dev = nullptr;
    bool y_owned = false;
    int16_t *x_loc = nullptr;
    double *ll_dev = nullptr;
    int64_t local_begin = 0, local_end = 0, main_begin = 0, main_end = 0, T_global = 0, Lc = 0;
    bool first = false, last = false;
    int c_main0 = 0, c_main1 = 0;  // local chunk index range of the main span [c_main0, c_main1)
    int device = 0;
    int last_fwd_rep = 0, last_bwd_rep = 0;
    std::vector<void *> owned;
    // peer-memory exchange (hmm_vshard_p2p_*)
    int rank = -1, world = 0;
    char *xblock = nullptr;                  // this rank's exchange block (flags + summaries, double-buffered)
    char **peers_dev = nullptr;              // device array [world]: every rank's exchange block
    std::vector<void *> ipc_opened;          // peers' blocks opened through CUDA IPC (closed on destroy)
    unsigned long long *epoch_dev = nullptr; // decode counter (parity selects the buffer)
    double *out_h = nullptr, *out_d = nullptr;  // mapped pinned: [total ll, inconsistent boundaries]
    cudaGraphExec_t p2p_exec = nullptr;
    int16_t *p2p_x = nullptr;
    int p2p_runs = 0;
    ~hmm_vshard() {
        if (p2p_exec) cudaGraphExecDestroy(p2p_exec);
        for (void *q : ipc_opened) cudaIpcCloseMemHandle(q);
        if (out_h) cudaFreeHost(out_h);
        for (void *q : owned) cudaFree(q);
    }
};

static void *shard_alloc(hmm_vshard *h, size_t bytes) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(HMM_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    h->owned.push_back(q);
    return q;
}

int hmm_vshard_chunking(int64_t T_global, int32_t n_ranks, int32_t N, int32_t K, int64_t *chunk_len_out,
                        int64_t *warmup_out) {
    return guarded([&] {
        if (!chunk_len_out || !warmup_out || T_global < 1 || n_ranks < 1) fail(HMM_EINVAL, "bad arguments");
        require_device();
        HostModel M;
        M.N = N;
        M.K = K;
        M.nstates = 1 + N * (K - 1);
        M.is_ring = true;
        if (!ring_supported(M, T_global)) fail(HMM_EUNSUPPORTED, "model/sequence outside the ring engine's range");
        ring_default_chunking(M, T_global, 1, n_ranks, chunk_len_out, warmup_out);
    });
}

int hmm_vshard_create(const double *y_local, int32_t y_is_host, int64_t local_begin, int64_t local_end,
                      int64_t main_begin, int64_t main_end, int64_t T_global, int64_t chunk_len, int64_t warmup,
                      const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *mu, double sigma, hmm_vshard **out) {
    return guarded([&] {
        if (!y_local || !out) fail(HMM_EINVAL, "null argument");
        if (!(0 <= local_begin && local_begin <= main_begin && main_begin < main_end && main_end <= local_end &&
              local_end <= T_global))
            fail(HMM_EINVAL, "inconsistent shard spans");
        if (chunk_len < 256 || warmup < 0 || warmup > chunk_len) fail(HMM_EINVAL, "bad chunk_len / warmup");
        const bool first = main_begin == 0, last = main_end == T_global;
        if ((main_begin - local_begin) % chunk_len || (!last && (main_end - local_begin) % chunk_len))
            fail(HMM_EINVAL, "shard spans must be multiples of chunk_len from local_begin");
        if (first != (local_begin == 0)) fail(HMM_EINVAL, "only the first shard may start at sample 0");
        if (!first && main_begin - local_begin != chunk_len) fail(HMM_EINVAL, "left ghost must be exactly one chunk");
        if (!last && local_end - main_end < warmup + 128) fail(HMM_EINVAL, "right ghost shorter than the look-ahead");
        require_device();
        std::unique_ptr<hmm_vshard> h(new hmm_vshard);
        HMM_CUDA(cudaGetDevice(&h->device));
        cudaStream_t st = main_stream();
        h->models.resize(1);
        analyse_model(states, N, K, nstates, tr, ntrans, mu, sigma, h->models[0]);
        const int64_t Tl = local_end - local_begin;
        if (!h->models[0].is_ring || !ring_supported(h->models[0], Tl))
            fail(HMM_EUNSUPPORTED, "time sharding needs a non-overlap ring model within the ring engine's range");
        h->FL = faithful_layout(nstates, ntrans);
        std::vector<char> hostblob(h->FL.bytes, 0);
        faithful_pack(h->models[0], h->FL, hostblob.data());
        h->blob_dev = (char *)shard_alloc(h.get(), hostblob.size());
        HMM_CUDA(cudaMemcpyAsync(h->blob_dev, hostblob.data(), hostblob.size(), cudaMemcpyHostToDevice, st));
        HMM_CUDA(cudaStreamSynchronize(st));
        if (y_is_host) {
            h->y_dev = (double *)shard_alloc(h.get(), sizeof(double) * (size_t)Tl);
            h->y_owned = true;
            HMM_CUDA(cudaMemcpyAsync(h->y_dev, y_local, sizeof(double) * (size_t)Tl, cudaMemcpyHostToDevice, st));
        } else
            h->y_dev = const_cast<double *>(y_local);
        h->x_loc = (int16_t *)shard_alloc(h.get(), sizeof(int16_t) * (size_t)Tl);
        h->ll_dev = (double *)shard_alloc(h.get(), sizeof(double));
        h->local_begin = local_begin;
        h->local_end = local_end;
        h->main_begin = main_begin;
        h->main_end = main_end;
        h->T_global = T_global;
        h->Lc = chunk_len;
        h->first = first;
        h->last = last;
        h->plan.own_memory = true;
        h->plan.build(h->y_dev, Tl, Tl, 1, h->models, h->FL, h->blob_dev, h->x_loc, Tl, chunk_len, warmup, first, last,
                      st);
        h->plan.set_ll_range(main_begin - local_begin, main_end - local_begin, local_begin, T_global, first);
        h->c_main0 = (int)((main_begin - local_begin) / chunk_len);
        h->c_main1 = last ? h->plan.nchunks() : (int)((main_end - local_begin) / chunk_len);
        if (h->c_main1 > h->plan.nchunks()) h->c_main1 = h->plan.nchunks();
        if (!last && h->c_main1 >= h->plan.nchunks()) fail(HMM_EINVAL, "right ghost chunk missing");
        HMM_CUDA(cudaStreamSynchronize(st));
        *out = h.release();
    });
}

int hmm_vshard_bvec(const hmm_vshard *h) { return h ? h->plan.bvec() : 0; }

int hmm_vshard_set_y(hmm_vshard *h, const double *y_local_host) {
    return guarded([&] {
        if (!h || !y_local_host) fail(HMM_EINVAL, "null argument");
        if (!h->y_owned) fail(HMM_EINVAL, "the shard was created on a caller-owned device buffer: write the new samples there");
        HMM_CUDA(cudaSetDevice(h->device));
        HMM_CUDA(cudaMemcpyAsync(h->y_dev, y_local_host, sizeof(double) * (size_t)(h->local_end - h->local_begin),
                                 cudaMemcpyHostToDevice, main_stream()));
    });
}

static void shard_dev(hmm_vshard *h) {
    if (!h) fail(HMM_EINVAL, "null shard");
    HMM_CUDA(cudaSetDevice(h->device));
}

int hmm_vshard_forward(hmm_vshard *h) {
    return guarded([&] {
        shard_dev(h);
        h->plan.forward(main_stream(), nullptr);
    });
}

int hmm_vshard_fwd_boundary_get(hmm_vshard *h, double *out, int32_t out_is_device) {
    return guarded([&] {
        shard_dev(h);
        if (!out) fail(HMM_EINVAL, "null out");
        if (h->last) fail(HMM_EINVAL, "the last shard has no right neighbour");
        cudaStream_t st = main_stream();
        HMM_CUDA(cudaMemcpyAsync(out, h->plan.eb_ptr(h->c_main1 - 1), sizeof(double) * h->plan.bvec(),
                                 out_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (!out_is_device) HMM_CUDA(cudaStreamSynchronize(st));
    });
}

int hmm_vshard_fwd_boundary_set(hmm_vshard *h, const double *in, int32_t in_is_device) {
    return guarded([&] {
        shard_dev(h);
        if (!in) fail(HMM_EINVAL, "null in");
        if (h->first) fail(HMM_EINVAL, "the first shard has no left neighbour");
        cudaStream_t st = main_stream();
        // the left ghost chunk's own (speculative) end vector is replaced by the neighbour's true one
        HMM_CUDA(cudaMemcpyAsync(h->plan.eb_ptr(h->c_main0 - 1), in, sizeof(double) * h->plan.bvec(),
                                 in_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        if (!in_is_device) HMM_CUDA(cudaStreamSynchronize(st));
    });
}

int hmm_vshard_fwd_verify(hmm_vshard *h, int32_t *n_repaired) {
    return guarded([&] {
        shard_dev(h);
        cudaStream_t st = main_stream();
        if (n_repaired) h->plan.reset_counters(st);
        h->plan.verify_fwd(st);
        if (n_repaired) {
            int f = 0, b = 0;
            h->plan.read_counters(st, &f, &b);
            *n_repaired = f;
        }
    });
}

int hmm_vshard_trace(hmm_vshard *h) {
    return guarded([&] {
        shard_dev(h);
        h->plan.trace(main_stream());
    });
}

int hmm_vshard_trace_boundary_get(hmm_vshard *h, int64_t *out, int32_t out_is_device) {
    return guarded([&] {
        shard_dev(h);
        if (!out) fail(HMM_EINVAL, "null out");
        if (h->first) fail(HMM_EINVAL, "the first shard has no left neighbour");
        cudaStream_t st = main_stream();
        if (out_is_device) {  // chain entry time: local -> global, on the device, asynchronous
            shift_state_kernel<<<1, 1, 0, st>>>(h->plan.own_start_ptr(h->c_main0), (long long *)out,
                                                8 * (long long)h->local_begin);
            HMM_CUDA(cudaGetLastError());
            return;
        }
        long long v = 0;
        HMM_CUDA(cudaMemcpyAsync(&v, h->plan.own_start_ptr(h->c_main0), sizeof v, cudaMemcpyDeviceToHost, st));
        HMM_CUDA(cudaStreamSynchronize(st));
        if (v >= 0) v += 8 * (long long)h->local_begin;
        *out = (int64_t)v;
    });
}

int hmm_vshard_trace_boundary_set(hmm_vshard *h, const int64_t *in, int32_t in_is_device) {
    return guarded([&] {
        shard_dev(h);
        if (!in) fail(HMM_EINVAL, "null in");
        if (h->last) fail(HMM_EINVAL, "the last shard has no right neighbour");
        // traceback state at main_end = start of the right ghost chunk; global -> local entry time
        cudaStream_t st = main_stream();
        if (in_is_device) {
            shift_state_kernel<<<1, 1, 0, st>>>((const long long *)in, h->plan.own_start_ptr(h->c_main1),
                                                -8 * (long long)h->local_begin);
            HMM_CUDA(cudaGetLastError());
            return;
        }
        long long v = (long long)*in;
        if (v >= 0) v -= 8 * (long long)h->local_begin;
        HMM_CUDA(cudaMemcpyAsync(h->plan.own_start_ptr(h->c_main1), &v, sizeof v, cudaMemcpyHostToDevice, st));
        HMM_CUDA(cudaStreamSynchronize(st));
    });
}

int hmm_vshard_trace_verify(hmm_vshard *h, int32_t *n_repaired) {
    return guarded([&] {
        shard_dev(h);
        cudaStream_t st = main_stream();
        if (n_repaired) h->plan.reset_counters(st);
        h->plan.verify_trace(st);
        if (n_repaired) {
            int f = 0, b = 0;
            h->plan.read_counters(st, &f, &b);
            *n_repaired = b;
        }
    });
}

int hmm_vshard_finish(hmm_vshard *h, int16_t *x_main_out, int32_t x_is_device, double *ll_partial_out) {
    return hmm_vshard_finish_ex(h, x_main_out, x_is_device, ll_partial_out, nullptr, nullptr);
}

int hmm_vshard_finish_ex(hmm_vshard *h, int16_t *x_main_out, int32_t x_is_device, double *ll_partial_out,
                         int32_t *fwd_repaired, int32_t *trace_repaired) {
    return guarded([&] {
        shard_dev(h);
        if (!x_main_out) fail(HMM_EINVAL, "null x_main_out");
        cudaStream_t st = main_stream();
        const int64_t lo = h->main_begin - h->local_begin, hi = h->main_end - h->local_begin;
        if (ll_partial_out) {
            // sum over the main span of (T_global - t) * (lp + q) with global t, plus the t = 0 term on the first shard
            h->plan.path_ll(st, h->ll_dev, lo, hi, h->local_begin, h->T_global, h->first);
            HMM_CUDA(cudaMemcpyAsync(ll_partial_out, h->ll_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        HMM_CUDA(cudaMemcpyAsync(x_main_out, h->x_loc + lo, sizeof(int16_t) * (size_t)(hi - lo),
                                 x_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        int cnt[4] = {0, 0, 0, 0};
        if (fwd_repaired || trace_repaired)
            HMM_CUDA(cudaMemcpyAsync(cnt, h->plan.counters_ptr(), sizeof cnt, cudaMemcpyDeviceToHost, st));
        HMM_CUDA(cudaStreamSynchronize(st));
        if (fwd_repaired) *fwd_repaired = cnt[0];
        if (trace_repaired) *trace_repaired = cnt[1];
    });
}

// summary layout (doubles): [0, bvec) forward vector at main_end | [bvec, 2 bvec) speculative start vector at
// main_begin | trace state at main_begin (own, global time) | trace state assumed at main_end | partial ll | 0
__global__ void shard_summary_kernel(const double *eb_last, const double *sb_first, const long long *own_first,
                                     const long long *own_ghost, const double *ll, long long shift, int bvec,
                                     double *out) {
    for (int k = threadIdx.x; k < bvec; k += blockDim.x) {
        out[k] = eb_last ? eb_last[k] : 0.0;
        out[bvec + k] = sb_first ? sb_first[k] : 0.0;
    }
    if (threadIdx.x == 0) {
        auto glob = [&](const long long *q) {
            if (!q) return -2.0;
            const long long v = *q;
            return (double)(v >= 0 ? v + shift : v);  // < 2^53: exact
        };
        out[2 * bvec + 0] = glob(own_first);
        out[2 * bvec + 1] = glob(own_ghost);
        out[2 * bvec + 2] = ll[0];
        out[2 * bvec + 3] = 0.0;
    }
}

int hmm_vshard_summary_len(const hmm_vshard *h) { return h ? 2 * h->plan.bvec() + 4 : 0; }

int hmm_vshard_summary_dev(hmm_vshard *h, int16_t *x_main_dev, double *summary_dev) {
    return guarded([&] {
        shard_dev(h);
        if (!summary_dev) fail(HMM_EINVAL, "null summary_dev");
        cudaStream_t st = main_stream();
        const int64_t lo = h->main_begin - h->local_begin, hi = h->main_end - h->local_begin;
        h->plan.path_ll(st, h->ll_dev, lo, hi, h->local_begin, h->T_global, h->first);
        shard_summary_kernel<<<1, 256, 0, st>>>(h->last ? nullptr : h->plan.eb_ptr(h->c_main1 - 1),
                                                h->first ? nullptr : h->plan.sb_ptr(h->c_main0),
                                                h->first ? nullptr : h->plan.own_start_ptr(h->c_main0),
                                                h->last ? nullptr : h->plan.own_start_ptr(h->c_main1), h->ll_dev,
                                                8 * (long long)h->local_begin, h->plan.bvec(), summary_dev);
        if (x_main_dev)
            HMM_CUDA(cudaMemcpyAsync(x_main_dev, h->x_loc + lo, sizeof(int16_t) * (size_t)(hi - lo),
                                     cudaMemcpyDeviceToDevice, st));
        HMM_CUDA(cudaGetLastError());
    });
}

int hmm_vshard_judge_dev(hmm_vshard *h, const double *gathered_dev, int32_t n_ranks, double *out_dev) {
    return guarded([&] {
        shard_dev(h);
        if (!gathered_dev || !out_dev || n_ranks < 1) fail(HMM_EINVAL, "bad judge arguments");
        vshard_judge_run(gathered_dev, n_ranks, h->plan.bvec(), out_dev, main_stream());
    });
}

// ---- peer-memory protocol: summaries travel by direct stores into every peer's exchange block ---------------
int hmm_vshard_p2p_init(hmm_vshard *h, int32_t rank, int32_t world, void *ipc_handle_out, void **block_ptr_out) {
    return guarded([&] {
        shard_dev(h);
        if (rank < 0 || world < 1 || rank >= world || world > 64) fail(HMM_EINVAL, "bad rank / world");
        if (h->xblock) fail(HMM_EINVAL, "p2p already initialised for this shard");
        cudaStream_t st = main_stream();
        const size_t bytes = vshard_exchange_block_bytes(world, h->plan.bvec());
        h->xblock = (char *)shard_alloc(h, bytes);
        h->peers_dev = (char **)shard_alloc(h, sizeof(char *) * (size_t)world);
        h->epoch_dev = (unsigned long long *)shard_alloc(h, sizeof(unsigned long long));
        HMM_CUDA(cudaMemsetAsync(h->xblock, 0, bytes, st));
        HMM_CUDA(cudaMemsetAsync(h->epoch_dev, 0, sizeof(unsigned long long), st));
        HMM_CUDA(cudaHostAlloc((void **)&h->out_h, 2 * sizeof(double), cudaHostAllocMapped));
        HMM_CUDA(cudaHostGetDevicePointer((void **)&h->out_d, h->out_h, 0));
        h->out_h[0] = h->out_h[1] = 0.0;
        HMM_CUDA(cudaStreamSynchronize(st));
        h->rank = rank;
        h->world = world;
        if (ipc_handle_out) {
            cudaIpcMemHandle_t hd;
            HMM_CUDA(cudaIpcGetMemHandle(&hd, h->xblock));
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
            memcpy(ipc_handle_out, &hd, sizeof hd);
        }
        if (block_ptr_out) *block_ptr_out = h->xblock;
    });
}

int hmm_vshard_p2p_attach(hmm_vshard *h, const void *ipc_handles, void *const *block_ptrs) {
    return guarded([&] {
        shard_dev(h);
        if (!h->xblock) fail(HMM_EINVAL, "hmm_vshard_p2p_init first");
        if (!ipc_handles && !block_ptrs) fail(HMM_EINVAL, "need IPC handles (other processes) or block pointers (this process)");
        std::vector<char *> peers((size_t)h->world, nullptr);
        for (int r = 0; r < h->world; r++) {
            if (r == h->rank) {
                peers[r] = h->xblock;
            } else if (block_ptrs) {
                peers[r] = (char *)block_ptrs[r];
                // same process, possibly another device: make its memory reachable from this one
                cudaPointerAttributes a;
                HMM_CUDA(cudaPointerGetAttributes(&a, peers[r]));
                if (a.device != h->device) {
                    int can = 0;
                    HMM_CUDA(cudaDeviceCanAccessPeer(&can, h->device, a.device));
                    if (!can) fail(HMM_EUNSUPPORTED, "device %d cannot access device %d's memory", h->device, a.device);
                    cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) HMM_CUDA(e);
                    cudaGetLastError();
                }
            } else {
                cudaIpcMemHandle_t hd;
                memcpy(&hd, (const char *)ipc_handles + (size_t)r * sizeof hd, sizeof hd);
                void *q = nullptr;
                HMM_CUDA(cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess));
                h->ipc_opened.push_back(q);
                peers[r] = (char *)q;
            }
        }
        cudaStream_t st = main_stream();
        HMM_CUDA(cudaMemcpyAsync(h->peers_dev, peers.data(), sizeof(char *) * peers.size(), cudaMemcpyHostToDevice, st));
        HMM_CUDA(cudaStreamSynchronize(st));
    });
}

static void p2p_local_sequence(hmm_vshard *h, int16_t *x_main_dev, cudaStream_t st) {
    const int64_t lo = h->main_begin - h->local_begin, hi = h->main_end - h->local_begin;
    h->plan.forward(st, nullptr);
    h->plan.verify_fwd(st);
    h->plan.trace(st);
    h->plan.verify_trace(st);
    h->plan.path_ll(st, h->ll_dev, lo, hi, h->local_begin, h->T_global, h->first);
    vshard_exchange_run(h->last ? nullptr : h->plan.eb_ptr(h->c_main1 - 1), h->first ? nullptr : h->plan.sb_ptr(h->c_main0),
                        h->first ? nullptr : h->plan.own_start_ptr(h->c_main0),
                        h->last ? nullptr : h->plan.own_start_ptr(h->c_main1), h->ll_dev, 8 * (long long)h->local_begin,
                        h->plan.bvec(), h->peers_dev, h->rank, h->world, h->epoch_dev, st);
    if (x_main_dev)
        HMM_CUDA(cudaMemcpyAsync(x_main_dev, h->x_loc + lo, sizeof(int16_t) * (size_t)(hi - lo), cudaMemcpyDeviceToDevice, st));
}

int hmm_vshard_p2p_launch(hmm_vshard *h, int16_t *x_main_dev) {
    return guarded([&] {
        shard_dev(h);
        if (!h->peers_dev || h->rank < 0) fail(HMM_EINVAL, "hmm_vshard_p2p_init / attach first");
        cudaStream_t st = main_stream();
        const bool no_graph = getenv("HMMCUDA_NO_GRAPH") && atoi(getenv("HMMCUDA_NO_GRAPH"));
        h->p2p_runs++;
        if (!no_graph && h->p2p_runs >= 2) {
            if (h->p2p_exec && h->p2p_x != x_main_dev) {
                cudaGraphExecDestroy(h->p2p_exec);
                h->p2p_exec = nullptr;
            }
            if (!h->p2p_exec) {
                cudaGraph_t graph = nullptr;
                HMM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                try {
                    p2p_local_sequence(h, x_main_dev, st);
                } catch (...) {
                    cudaStreamEndCapture(st, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                HMM_CUDA(cudaStreamEndCapture(st, &graph));
                cudaError_t e = cudaGraphInstantiate(&h->p2p_exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) {
                    h->p2p_exec = nullptr;
                    fail(HMM_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
                }
                h->p2p_x = x_main_dev;
            }
            HMM_CUDA(cudaGraphLaunch(h->p2p_exec, st));
        } else
            p2p_local_sequence(h, x_main_dev, st);
    });
}

int hmm_vshard_p2p_finish(hmm_vshard *h, double *ll_total_out, int32_t *bad_out) {
    return guarded([&] {
        shard_dev(h);
        if (!h->peers_dev) fail(HMM_EINVAL, "hmm_vshard_p2p_init / attach first");
        cudaStream_t st = main_stream();
        vshard_judge_p2p_run(h->xblock, h->world, h->plan.bvec(), h->epoch_dev, h->out_d, st);
        HMM_CUDA(cudaStreamSynchronize(st));
        if (h->out_h[1] < 0) fail(HMM_ECUDA, "time-sharded decode: a peer's boundary summary never arrived (2 s)");
        h->plan.check_guards(st);
        if (ll_total_out) *ll_total_out = h->out_h[0];
        if (bad_out) *bad_out = (int32_t)h->out_h[1];
    });
}

int hmm_vshard_repairs(hmm_vshard *h, int32_t *fwd_repaired, int32_t *trace_repaired) {
    return guarded([&] {
        shard_dev(h);
        int f = 0, b = 0;
        h->plan.read_counters(main_stream(), &f, &b);
        if (fwd_repaired) *fwd_repaired = f;
        if (trace_repaired) *trace_repaired = b;
    });
}

int hmm_vshard_destroy(hmm_vshard *h) {
    return guarded([&] {
        if (!h) return;
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        delete h;
    });
}

// ---------------------------------------------------------------------------
// forward / backward (dense output)
// ---------------------------------------------------------------------------
static int fb_host(bool backward, const double *V, int64_t T, const int16_t *states, int32_t N, int32_t K,
                   int32_t nstates, const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, double *out) {
    return guarded([&] {
        if (!V || !out) fail(HMM_EINVAL, "null V / output");
        if (T < 1) fail(HMM_EINVAL, "T must be positive");
        require_device();
        cudaStream_t st = main_stream();
        Workspace &ws = workspace();
        BatchModels &B = get_models(1, states, 1, N, K, nstates, tr, ntrans, mu, &sigma, st);
        double *V_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T);
        double *o_dev = (double *)ws.get(backward ? Workspace::BETA : Workspace::ALPHA,
                                         sizeof(double) * (size_t)T * nstates);
        h2d(V_dev, V, sizeof(double) * (size_t)T, st);
        const bool no_ring = getenv("HMMCUDA_DENSE_FB_FAITHFUL") && atoi(getenv("HMMCUDA_DENSE_FB_FAITHFUL"));
        if (B.models[0].is_ring && ring_supported(B.models[0], T) && !no_ring)
            ring_fb_dense_run(V_dev, T, B.models[0], backward ? nullptr : o_dev, backward ? o_dev : nullptr, st);
        else
            faithful_fb_run(backward, V_dev, T, B.layout, B.blob_dev, B.models[0], o_dev, st);
        d2h(out, o_dev, sizeof(double) * (size_t)T * nstates, st);
        HMM_CUDA(cudaStreamSynchronize(st));
    });
}

int hmm_forward_f64(const double *V, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, double *alpha_out) {
    return fb_host(false, V, T, states, N, K, nstates, tr, ntrans, mu, sigma, alpha_out);
}
int hmm_backward_f64(const double *V, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                     const hmm_trans *tr, int64_t ntrans, const double *mu, double sigma, double *beta_out) {
    return fb_host(true, V, T, states, N, K, nstates, tr, ntrans, mu, sigma, beta_out);
}

// ---------------------------------------------------------------------------
// update / E-M step
// ---------------------------------------------------------------------------
static void export_em(const EmResult &r, const HostModel &M, double *mu_inout, double *sigma_inout, double *lp_out,
                      double *pp_out, double *loglik_out) {
    memcpy(mu_inout, r.mu.data(), sizeof(double) * (size_t)M.K * M.N);  // in place, src/baumwelch.jl:268
    *sigma_inout = r.sigma;
    if (lp_out) memcpy(lp_out, r.lp.data(), sizeof(double) * r.lp.size());
    if (pp_out) memcpy(pp_out, r.pp.data(), sizeof(double) * M.nstates);
    if (loglik_out) *loglik_out = r.loglik;
}

int hmm_update_f64(const double *alpha, const double *beta, int64_t T, const int16_t *states, int32_t N, int32_t K,
                   int32_t nstates, const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout,
                   const double *x, double *lp_out, double *pp_out) {
    return guarded([&] {
        if (!alpha || !beta || !x || !mu_inout || !sigma_inout) fail(HMM_EINVAL, "null argument");
        if (T < 2) fail(HMM_EINVAL, "T must be at least 2");
        require_device();
        cudaStream_t st = main_stream();
        Workspace &ws = workspace();
        HostModel M;
        analyse_model(states, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, M);
        size_t nb = sizeof(double) * (size_t)T * nstates;
        double *a_dev = (double *)ws.get(Workspace::ALPHA, nb);
        double *b_dev = (double *)ws.get(Workspace::BETA, nb);
        double *x_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T);
        h2d(a_dev, alpha, nb, st);
        h2d(b_dev, beta, nb, st);
        h2d(x_dev, x, sizeof(double) * (size_t)T, st);
        EmResult r;
        dense_update_run(a_dev, b_dev, x_dev, T, M, states, r, st);
        export_em(r, M, mu_inout, sigma_inout, lp_out, pp_out, nullptr);
    });
}

static void em_step_dev(const double *X_dev, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                        const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                        double *pp_out, double *loglik_out, int mode, cudaStream_t st, hmm_info *info) {
    HostModel M;
    analyse_model(states, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, M);
    bool ring = M.is_ring && ring_supported(M, T);
    if (mode == HMM_MODE_RING && !ring)
        fail(HMM_EUNSUPPORTED, "HMM_MODE_RING requested but the model/sequence is outside the ring engine's range");
    if (mode == HMM_MODE_FAITHFUL) ring = false;
    EmResult r;
    if (ring) {
        if (info) info->engine = HMM_MODE_RING;
        ring_em_run(X_dev, T, M, r, st, info);
    } else {
        // generic path: dense alpha/beta on the device (never cross the boundary), then update
        if (info) info->engine = HMM_MODE_FAITHFUL;
        Workspace &ws = workspace();
        size_t nb = sizeof(double) * (size_t)T * nstates;
        double *a_dev = (double *)ws.get(Workspace::ALPHA, nb);
        double *b_dev = (double *)ws.get(Workspace::BETA, nb);
        FaithfulLayout L = faithful_layout(nstates, ntrans);
        std::vector<char> host(L.bytes, 0);
        faithful_pack(M, L, host.data());
        char *blob = (char *)ws.get(Workspace::MODEL, host.size());
        HMM_CUDA(cudaMemcpyAsync(blob, host.data(), host.size(), cudaMemcpyHostToDevice, st));
        HMM_CUDA(cudaStreamSynchronize(st));
        faithful_fb_run(false, X_dev, T, L, blob, M, a_dev, st);
        faithful_fb_run(true, X_dev, T, L, blob, M, b_dev, st);
        dense_update_run(a_dev, b_dev, X_dev, T, M, states, r, st);
        if (info) info->kernel_launches += 4;
    }
    export_em(r, M, mu_inout, sigma_inout, lp_out, pp_out, loglik_out);
}

int hmm_em_step_ex_f64(const double *X, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                       const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                       double *pp_out, double *loglik_out, int32_t mode, hmm_info *info) {
    return guarded([&] {
        if (info) memset(info, 0, sizeof *info);
        if (!X || !mu_inout || !sigma_inout) fail(HMM_EINVAL, "null argument");
        if (T < 2) fail(HMM_EINVAL, "T must be at least 2");
        require_device();
        cudaStream_t st = main_stream();
        Timer tall(st);
        tall.start();
        double *X_dev = (double *)workspace().get(Workspace::Y, sizeof(double) * (size_t)T);
        h2d(X_dev, X, sizeof(double) * (size_t)T, st);
        em_step_dev(X_dev, T, states, N, K, nstates, tr, ntrans, mu_inout, sigma_inout, lp_out, pp_out, loglik_out,
                    mode, st, info);
        tall.stop();
        HMM_CUDA(cudaStreamSynchronize(st));
        if (info) info->device_ms = tall.ms();
    });
}

int hmm_em_step_f64(const double *X, int64_t T, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                    const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                    double *pp_out, double *loglik_out) {
    return hmm_em_step_ex_f64(X, T, states, N, K, nstates, tr, ntrans, mu_inout, sigma_inout, lp_out, pp_out,
                              loglik_out, HMM_MODE_AUTO, nullptr);
}

struct hmm_train_ctx {
    double *X_dev = nullptr;
    int64_t T = 0;
    bool owned = false;
    int device = 0;
};

int hmm_train_create(const double *X, int64_t T, hmm_train_ctx **ctx_out) {
    return guarded([&] {
        if (!X || !ctx_out || T < 2) fail(HMM_EINVAL, "bad arguments");
        require_device();
        hmm_train_ctx *c = new hmm_train_ctx;
        c->T = T;
        c->owned = true;
        HMM_CUDA(cudaGetDevice(&c->device));
        cudaError_t e = cudaMalloc(&c->X_dev, sizeof(double) * (size_t)T);
        if (e != cudaSuccess) {
            delete c;
            cudaGetLastError();
            fail(HMM_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e));
        }
        cudaStream_t st = main_stream();
        h2d(c->X_dev, X, sizeof(double) * (size_t)T, st);
        HMM_CUDA(cudaStreamSynchronize(st));
        *ctx_out = c;
    });
}

int hmm_train_create_dev(const double *X_dev, int64_t T, hmm_train_ctx **ctx_out) {
    return guarded([&] {
        if (!X_dev || !ctx_out || T < 2) fail(HMM_EINVAL, "bad arguments");
        require_device();
        hmm_train_ctx *c = new hmm_train_ctx;
        c->T = T;
        c->X_dev = const_cast<double *>(X_dev);
        HMM_CUDA(cudaGetDevice(&c->device));
        *ctx_out = c;
    });
}

int hmm_train_em_step(hmm_train_ctx *ctx, const int16_t *states, int32_t N, int32_t K, int32_t nstates,
                      const hmm_trans *tr, int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out,
                      double *pp_out, double *loglik_out, hmm_info *info) {
    return guarded([&] {
        if (info) memset(info, 0, sizeof *info);
        if (!ctx || !mu_inout || !sigma_inout) fail(HMM_EINVAL, "null argument");
        require_device();
        cudaStream_t st = main_stream();
        Timer tall(st);
        tall.start();
        em_step_dev(ctx->X_dev, ctx->T, states, N, K, nstates, tr, ntrans, mu_inout, sigma_inout, lp_out, pp_out,
                    loglik_out, HMM_MODE_AUTO, st, info);
        tall.stop();
        HMM_CUDA(cudaStreamSynchronize(st));
        if (info) info->device_ms = tall.ms();
    });
}

// Weights of an unchanged set of transitions from a new lp vector (src/types.jl:94-127): per transition and neuron the
// term is lpz (silent -> silent), lp[i] (silent -> first phase) or 0 (advance / return), added in neuron order like
// `lpt += ...`; lpz = log1p(-exp(sum(lp))) over the WHOLE vector (left-to-right sum).  Host arithmetic only.
static void transition_term_codes(const int16_t *states, int N, int nstates, const hmm_trans *tr, int64_t ntrans,
                                  std::vector<unsigned char> &code) {
    code.resize((size_t)ntrans * N);
    for (int64_t e = 0; e < ntrans; e++) {
        const int64_t a = tr[e].src - 1, b = tr[e].dst - 1;
        if (a < 0 || a >= nstates || b < 0 || b >= nstates) fail(HMM_EINVAL, "transition %lld out of range", (long long)e);
        for (int i = 0; i < N; i++) {
            const int s1 = states[(size_t)a * N + i], s2 = states[(size_t)b * N + i];
            code[(size_t)e * N + i] = (s1 == 1 && s2 == 1) ? 0 : (s1 == 1 && s2 == 2) ? 1 : 2;  // phases are 1-based, 1 = silent
        }
    }
}
// false (and tr untouched) if a weight is not finite: the set of finite transitions would change
static bool transition_weights_from_lp(const std::vector<unsigned char> &code, int N, const double *lp, int nlp,
                                       hmm_trans *tr, int64_t ntrans, std::vector<double> &tmp) {
    double sum = 0.0;
    for (int k = 0; k < nlp; k++) sum = k == 0 ? lp[0] : sum + lp[k];
    const double lpz = log1p(-exp(sum));
    if (!std::isfinite(lpz)) return false;
    tmp.resize((size_t)ntrans);
    for (int64_t e = 0; e < ntrans; e++) {
        double w = 0.0;
        for (int i = 0; i < N; i++) {
            const unsigned char c = code[(size_t)e * N + i];
            w += c == 0 ? lpz : c == 1 ? lp[i] : 0.0;
        }
        if (!std::isfinite(w)) return false;
        tmp[(size_t)e] = w;
    }
    for (int64_t e = 0; e < ntrans; e++) tr[e].lp = tmp[(size_t)e];
    return true;
}

int hmm_transition_weights(const int16_t *states, int32_t N, int32_t nstates, hmm_trans *tr_inout, int64_t ntrans,
                           const double *lp, int32_t nlp, int32_t *all_finite) {
    return guarded([&] {
        if (!states || !tr_inout || !lp || nlp < N || N < 1) fail(HMM_EINVAL, "null argument or lp shorter than N");
        std::vector<unsigned char> code;
        std::vector<double> tmp;
        transition_term_codes(states, N, nstates, tr_inout, ntrans, code);
        const bool ok = transition_weights_from_lp(code, N, lp, nlp, tr_inout, ntrans, tmp);
        if (all_finite) *all_finite = ok ? 1 : 0;
    });
}

// The E/M loop of src/baumwelch.jl:325-335 in one call (no callback between the steps): each step's lp goes
// straight into the next step's transition weights -- the StateMatrix rebuild of src/baumwelch.jl:265 /
// src/types.jl:94-127 restricted to what it can change, the WEIGHTS: which transitions are finite depends on the
// state layout only.  Per transition and neuron the term is lpz (silent -> silent), lp[i] (silent -> first phase)
// or 0 (advance / return), added in neuron order like `lpt += ...`; lpz = log1p(-exp(sum(lp))) over the whole vector.
// Stops early, reporting the steps done, when a weight stops being finite (the set of transitions would change:
// the caller rebuilds its StateMatrix and decides).
int hmm_train_run(hmm_train_ctx *ctx, const int16_t *states, int32_t N, int32_t K, int32_t nstates, hmm_trans *tr_inout,
                  int64_t ntrans, double *mu_inout, double *sigma_inout, double *lp_out, int32_t nlp, double *pp_out,
                  double *loglik_out, int32_t nsteps, int32_t *steps_done, hmm_info *info) {
    return guarded([&] {
        if (info) memset(info, 0, sizeof *info);
        if (steps_done) *steps_done = 0;
        if (!ctx || !states || !tr_inout || !mu_inout || !sigma_inout || !lp_out || !pp_out) fail(HMM_EINVAL, "null argument");
        if (nsteps < 0 || nlp < N) fail(HMM_EINVAL, "nsteps must be >= 0 and lp_out must hold at least N entries");
        require_device();
        cudaStream_t st = main_stream();
        std::vector<unsigned char> code;  // term codes per (transition, neuron): 0 -> lpz, 1 -> lp[i], 2 -> 0.0
        std::vector<double> wtmp;
        transition_term_codes(states, N, nstates, tr_inout, ntrans, code);
        hmm_info acc{}, one{};
        double dev_ms = 0, ker_ms = 0;
        Timer tall(st);  // (one pair of events for the whole loop: creating them costs microseconds per step)
        for (int it = 0; it < nsteps; it++) {
            double ll = 0;
            tall.start();
            em_step_dev(ctx->X_dev, ctx->T, states, N, K, nstates, tr_inout, ntrans, mu_inout, sigma_inout, lp_out, pp_out,
                        &ll, HMM_MODE_AUTO, st, &one);
            tall.stop();
            HMM_CUDA(cudaStreamSynchronize(st));
            dev_ms += tall.ms();
            ker_ms += one.top_kernel_ms;
            acc.engine = one.engine;
            acc.n_chunks = one.n_chunks;
            acc.fwd_repaired += one.fwd_repaired;
            acc.bwd_repaired += one.bwd_repaired;
            acc.kernel_launches += one.kernel_launches;
            memset(&one, 0, sizeof one);
            if (loglik_out) loglik_out[it] = ll;
            if (steps_done) *steps_done = it + 1;
            if (!transition_weights_from_lp(code, N, lp_out, nlp, tr_inout, ntrans, wtmp)) break;
        }
        if (info) {
            *info = acc;
            info->device_ms = dev_ms;
            info->top_kernel_ms = ker_ms;
        }
    });
}

int hmm_train_destroy(hmm_train_ctx *ctx) {
    return guarded([&] {
        if (!ctx) return;
        if (ctx->owned && ctx->X_dev) cudaFree(ctx->X_dev);
        delete ctx;
    });
}

// ---------------------------------------------------------------------------
// I/O front-end: raw recording file -> pinned staging -> HBM -> decode (src/hmmsort.jl:36-104 reads an HDF5 file,
// converts the samples to Float64 and decodes one channel per run; a contiguous HDF5 dataset is a raw block at a byte
// offset, which is what this entry point takes).  The file is read block by block by a reader thread into two pinned
// buffers, each block crosses PCIe while the next one is being read, and one kernel picks the requested channels out
// of the (interleaved or channel-major) block and widens int16 / float32 / float64 samples to the Float64 [T x C]
// layout the decoders take.  Then all channels are decoded from HBM and only x travels back.
// ---------------------------------------------------------------------------
}  // extern "C"

template <typename RAW>
__global__ void __launch_bounds__(256)
    raw_pick_kernel(const RAW *__restrict__ blk, int64_t t0, int64_t nt, int nfile, int interleaved, int64_t T_file,
                    const int *__restrict__ chans, int C, double scale, double *__restrict__ y /*[T x C]*/, int64_t T) {
    // one thread per (sample of the block, requested channel); consecutive threads walk consecutive samples
    const int64_t n = nt * C;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(k / nt);
        const int64_t t = k - (int64_t)c * nt;
        const int fc = chans[c];
        // interleaved: the block holds samples [t0, t0 + nt) of every file channel; channel-major: the block is a
        // slab [nfile x nt] gathered by the reader (row fc, column t)
        const RAW v = interleaved ? blk[t * nfile + fc] : blk[(int64_t)fc * nt + t];
        y[(size_t)c * T + t0 + t] = (double)v * scale;
    }
}

static void rawfile_decode_body(const char *path, int64_t byte_offset, int dtype, int nfile, int interleaved, double scale,
                                int64_t T, int C, const int32_t *channels, const int16_t *states, int states_shared, int N,
                                int K, int nstates, const hmm_trans *tr, int64_t ntrans, const double *mu,
                                const double *sigma, int16_t *x_out, double *ll_out, int mode, hmm_info *info) {
    if (info) memset(info, 0, sizeof *info);
    if (!path || !channels || !x_out || !sigma) fail(HMM_EINVAL, "null argument");
    if (T < 1 || C < 1 || nfile < 1) fail(HMM_EINVAL, "T, C and the file's channel count must be positive");
    const size_t esz = dtype == HMM_RAW_I16 ? 2 : dtype == HMM_RAW_F32 ? 4 : dtype == HMM_RAW_F64 ? 8 : 0;
    if (!esz) fail(HMM_EINVAL, "sample_dtype must be HMM_RAW_I16, HMM_RAW_F32 or HMM_RAW_F64");
    for (int c = 0; c < C; c++)
        if (channels[c] < 0 || channels[c] >= nfile) fail(HMM_EINVAL, "channel index %d outside the file's %d channels", channels[c], nfile);
    require_device();
    const int fd = open(path, O_RDONLY);
    if (fd < 0) fail(HMM_EINVAL, "cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (int64_t)sb.st_size < byte_offset + (int64_t)esz * nfile * T) {
        close(fd);
        fail(HMM_EINVAL, "%s is shorter than offset + %lld samples x %d channels", path, (long long)T, nfile);
    }
    cudaStream_t st = main_stream(), sh = copy_stream();
    Workspace &ws = workspace();
    double *y_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T * C);
    int16_t *x_dev = (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * (size_t)T * C);
    int *ch_dev = (int *)ws.get(Workspace::MISC, sizeof(int) * (size_t)C);
    HMM_CUDA(cudaMemcpyAsync(ch_dev, channels, sizeof(int) * (size_t)C, cudaMemcpyHostToDevice, st));
    // block = nt samples of every file channel, ~32 MB
    int64_t nt = std::max<int64_t>(4096, ((int64_t)32 << 20) / (int64_t)(esz * nfile));
    nt = std::min<int64_t>(nt, T);
    const size_t bbytes = (size_t)nt * nfile * esz;
    char *pin = (char *)ws.pinned(4, 2 * bbytes);
    char *blk_dev = (char *)ws.get(Workspace::FWDF, 2 * bbytes);
    cudaEvent_t ev_free[2], ev_copied[2];
    for (int k = 0; k < 2; k++) {
        HMM_CUDA(cudaEventCreateWithFlags(&ev_free[k], cudaEventDisableTiming));
        HMM_CUDA(cudaEventCreateWithFlags(&ev_copied[k], cudaEventDisableTiming));
    }
    const int64_t nblk = (T + nt - 1) / nt;
    // reader thread: block b into pinned buffer b & 1 as soon as that buffer's previous upload has left it
    std::mutex m;
    std::condition_variable cv;
    int64_t filled = 0, released = 2;  // blocks read so far / blocks whose pinned buffer may be refilled (b < released)
    bool io_error = false;
    std::thread reader([&] {
        for (int64_t b = 0; b < nblk; b++) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return b < released; });
            }
            const int64_t t0 = b * nt, n = std::min<int64_t>(nt, T - t0);
            char *dst = pin + (size_t)(b & 1) * bbytes;
            bool ok = true;
            if (interleaved) {
                const size_t want = (size_t)n * nfile * esz;
                ok = pread(fd, dst, want, byte_offset + (int64_t)t0 * nfile * (int64_t)esz) == (ssize_t)want;
            } else {  // channel-major file: gather the requested time range of every channel into a [nfile x n] slab
                for (int fc = 0; fc < nfile && ok; fc++)
                    ok = pread(fd, dst + (size_t)fc * n * esz, (size_t)n * esz,
                               byte_offset + ((int64_t)fc * T + t0) * (int64_t)esz) == (ssize_t)((size_t)n * esz);
            }
            std::lock_guard<std::mutex> lk(m);
            if (!ok) io_error = true;
            filled = b + 1;
            cv.notify_all();
            if (!ok) return;
        }
    });
    auto finish_reader = [&] {
        {
            std::lock_guard<std::mutex> lk(m);
            released = nblk + 2;
        }
        cv.notify_all();
        if (reader.joinable()) reader.join();
        close(fd);
        for (int k = 0; k < 2; k++) {
            cudaEventDestroy(ev_free[k]);
            cudaEventDestroy(ev_copied[k]);
        }
    };
    try {
        Timer tall(st);
        tall.start();
        for (int64_t b = 0; b < nblk; b++) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return filled > b || io_error; });
                if (io_error) fail(HMM_EINVAL, "read error on %s", path);
            }
            const int k = (int)(b & 1);
            const int64_t t0 = b * nt, n = std::min<int64_t>(nt, T - t0);
            if (b >= 2) HMM_CUDA(cudaStreamWaitEvent(sh, ev_free[k], 0));  // the pick kernel of block b-2 is done with blk_dev[k]
            HMM_CUDA(cudaMemcpyAsync(blk_dev + (size_t)k * bbytes, pin + (size_t)k * bbytes, (size_t)n * nfile * esz,
                                     cudaMemcpyHostToDevice, sh));
            HMM_CUDA(cudaEventRecord(ev_copied[k], sh));
            HMM_CUDA(cudaStreamWaitEvent(st, ev_copied[k], 0));
            const unsigned grid = (unsigned)std::min<int64_t>(148 * 16, (n * C + 255) / 256);
            const void *bd = blk_dev + (size_t)k * bbytes;
            if (dtype == HMM_RAW_I16)
                raw_pick_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t *)bd, t0, n, nfile, interleaved, T, ch_dev, C, scale, y_dev, T);
            else if (dtype == HMM_RAW_F32)
                raw_pick_kernel<float><<<grid, 256, 0, st>>>((const float *)bd, t0, n, nfile, interleaved, T, ch_dev, C, scale, y_dev, T);
            else
                raw_pick_kernel<double><<<grid, 256, 0, st>>>((const double *)bd, t0, n, nfile, interleaved, T, ch_dev, C, scale, y_dev, T);
            HMM_CUDA(cudaGetLastError());
            HMM_CUDA(cudaEventRecord(ev_free[k], st));
            // the pinned buffer may be refilled once its upload has completed
            HMM_CUDA(cudaEventSynchronize(ev_copied[k]));
            {
                std::lock_guard<std::mutex> lk(m);
                released = b + 3;
            }
            cv.notify_all();
        }
        BatchModels &B = get_models(C, states, states_shared, N, K, nstates, tr, ntrans, mu, sigma, st);
        viterbi_core(y_dev, T, C, B, x_dev, ll_out, nullptr, nullptr, mode, st, info);
        d2h(x_out, x_dev, sizeof(int16_t) * (size_t)T * C, st);
        tall.stop();
        HMM_CUDA(cudaStreamSynchronize(st));
        if (info) {
            info->device_ms = tall.ms();
            info->kernel_launches += nblk;
        }
    } catch (...) {
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(sh);
        finish_reader();
        throw;
    }
    finish_reader();
}

extern "C" {

int hmm_viterbi_rawfile(const char *path, int64_t byte_offset, int32_t sample_dtype, int32_t n_file_channels,
                        int32_t interleaved, double scale, int64_t T, int32_t C, const int32_t *channels,
                        const int16_t *states, int32_t states_shared, int32_t N, int32_t K, int32_t nstates,
                        const hmm_trans *tr, int64_t ntrans, const double *mu, const double *sigma, int16_t *x_out,
                        double *ll_out, int32_t mode, hmm_info *info) {
    return guarded([&] {
        rawfile_decode_body(path, byte_offset, sample_dtype, n_file_channels, interleaved, scale, T, C, channels, states,
                            states_shared, N, K, nstates, tr, ntrans, mu, sigma, x_out, ll_out, mode, info);
    });
}

// ---------------------------------------------------------------------------
// time-sharded Baum-Welch (one recording over several GPUs)
// ---------------------------------------------------------------------------
struct hmm_emshard {
    double *X_dev = nullptr;
    bool owned = false;
    int64_t local_begin = 0, local_end = 0, main_begin = 0, main_end = 0, T_global = 0, Lc = 0, W = 0;
    int device = 0;
};

int hmm_emshard_create(const double *X_local, int32_t x_is_host, int64_t local_begin, int64_t local_end,
                       int64_t main_begin, int64_t main_end, int64_t T_global, int64_t chunk_len, int64_t warmup,
                       hmm_emshard **out) {
    return guarded([&] {
        if (!X_local || !out) fail(HMM_EINVAL, "null argument");
        if (!(0 <= local_begin && local_begin <= main_begin && main_begin < main_end && main_end <= local_end &&
              local_end <= T_global))
            fail(HMM_EINVAL, "inconsistent shard spans");
        if (chunk_len < 256 || chunk_len % 256 || warmup < 0 || warmup > chunk_len) fail(HMM_EINVAL, "bad chunk_len / warmup");
        const bool first = main_begin == 0, last = main_end == T_global;
        if ((main_begin - local_begin) % chunk_len || (!last && (main_end - local_begin) % chunk_len))
            fail(HMM_EINVAL, "shard spans must be multiples of chunk_len from local_begin");
        if (first != (local_begin == 0)) fail(HMM_EINVAL, "only the first shard may start at sample 0");
        if (!first && main_begin - local_begin < chunk_len) fail(HMM_EINVAL, "left ghost must be at least one chunk");
        // a right ghost clipped by the end of the recording is complete whatever its length (the backward pass starts
        // from the true end there); it only has to be a chunk of its own for the kernels (>= 256 samples)
        if (!last && local_end - main_end < chunk_len && !(local_end == T_global && local_end - main_end >= 256))
            fail(HMM_EINVAL, "right ghost must be at least one chunk (or reach the end of the recording)");
        require_device();
        std::unique_ptr<hmm_emshard> h(new hmm_emshard);
        HMM_CUDA(cudaGetDevice(&h->device));
        const int64_t Tl = local_end - local_begin;
        if (x_is_host) {
            cudaStream_t st = main_stream();
            cudaError_t e = cudaMalloc((void **)&h->X_dev, sizeof(double) * (size_t)Tl);
            if (e != cudaSuccess) {
                cudaGetLastError();
                fail(HMM_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e));
            }
            h->owned = true;
            h2d(h->X_dev, X_local, sizeof(double) * (size_t)Tl, st);
            HMM_CUDA(cudaStreamSynchronize(st));
        } else
            h->X_dev = const_cast<double *>(X_local);
        h->local_begin = local_begin; h->local_end = local_end; h->main_begin = main_begin; h->main_end = main_end;
        h->T_global = T_global; h->Lc = chunk_len; h->W = warmup;
        *out = h.release();
    });
}

int hmm_emshard_chunking(int64_t T_global, int32_t n_ranks, int32_t N, int32_t K, int64_t *chunk_len_out,
                         int64_t *warmup_out) {
    return guarded([&] {
        if (!chunk_len_out || !warmup_out || T_global < 1 || n_ranks < 1 || N < 1 || N > 7 || K < 3) fail(HMM_EINVAL, "bad arguments");
        require_device();
        ring_em_default_chunking(N, K, (T_global + n_ranks - 1) / n_ranks, chunk_len_out, warmup_out);
    });
}
int hmm_emshard_stats_len(int32_t N, int32_t nstates) { return ring_em_xvec_len(N, nstates); }
int hmm_emshard_boundary_len(int32_t N, int32_t K) { return ring_em_bnd_len(N, K); }

int hmm_emshard_estep(hmm_emshard *h, const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *mu, double sigma, double *stats_dev, double *boundary_dev) {
    return guarded([&] {
        if (!h || !stats_dev || !boundary_dev) fail(HMM_EINVAL, "null argument");
        HMM_CUDA(cudaSetDevice(h->device));
        HostModel M;
        analyse_model(states, N, K, nstates, tr, ntrans, mu, sigma, M);
        const int64_t Tl = h->local_end - h->local_begin;
        if (!M.is_ring || !ring_supported(M, Tl)) fail(HMM_EUNSUPPORTED, "time-sharded E/M needs a non-overlap ring model");
        EmShardOpts o{h->Lc, h->W, h->main_begin - h->local_begin, h->main_end - h->local_begin, h->main_begin == 0,
                      h->main_end == h->T_global, stats_dev, boundary_dev};
        ring_em_shard_estep(h->X_dev, Tl, M, o, main_stream(), nullptr);
    });
}

int hmm_emshard_mstep(hmm_emshard *h, const int16_t *states, int32_t N, int32_t K, int32_t nstates, const hmm_trans *tr,
                      int64_t ntrans, const double *stats_sum_dev, double lS_global, double *mu_inout, double *sigma_inout,
                      double *lp_out, double *pp_out, double *loglik_out) {
    return guarded([&] {
        if (!h || !stats_sum_dev || !mu_inout || !sigma_inout) fail(HMM_EINVAL, "null argument");
        HMM_CUDA(cudaSetDevice(h->device));
        HostModel M;
        analyse_model(states, N, K, nstates, tr, ntrans, mu_inout, *sigma_inout, M);
        EmResult r;
        ring_em_shard_mstep(stats_sum_dev, M, h->T_global, lS_global, r, main_stream());
        export_em(r, M, mu_inout, sigma_inout, lp_out, pp_out, loglik_out);
    });
}

int hmm_emshard_destroy(hmm_emshard *h) {
    return guarded([&] {
        if (!h) return;
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        if (h->owned && h->X_dev) cudaFree(h->X_dev);
        delete h;
    });
}

// ---------------------------------------------------------------------------
// reconstruct / unroll
// ---------------------------------------------------------------------------
static void check_model_small(const int16_t *states, int N, int K, int nstates, const double *mu) {
    if (!states || !mu) fail(HMM_EINVAL, "null model array");
    if (N < 1 || K < 1 || nstates < 1 || nstates > 32767) fail(HMM_EINVAL, "bad N/K/nstates");
    for (size_t i = 0; i < (size_t)N * nstates; i++)
        if (states[i] < 1 || states[i] > K) fail(HMM_EINVAL, "states[%zu]=%d outside 1..K=%d", i, (int)states[i], K);
}

int hmm_reconstruct_f64(const int16_t *x, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                        const double *mu, int32_t K, double *Y_out) {
    return guarded([&] {
        if (!x || !Y_out) fail(HMM_EINVAL, "null x / Y_out");
        if (T < 0) fail(HMM_EINVAL, "negative T");
        check_model_small(states, N, K, nstates, mu);
        require_device();
        if (T == 0) return;
        cudaStream_t st = main_stream();
        Workspace &ws = workspace();
        std::vector<double> m;
        state_means(states, N, K, nstates, mu, m);
        int16_t *x_dev = (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * (size_t)T);
        double *Y_dev = (double *)ws.get(Workspace::Y, sizeof(double) * (size_t)T);
        h2d(x_dev, x, sizeof(int16_t) * (size_t)T, st);
        reconstruct_run(x_dev, T, m, Y_dev, st);
        d2h(Y_out, Y_dev, sizeof(double) * (size_t)T, st);
        HMM_CUDA(cudaStreamSynchronize(st));
    });
}

int hmm_reconstruct_dev_f64(const int16_t *x_dev, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                            const double *mu, int32_t K, double *Y_dev) {
    return guarded([&] {
        if (!x_dev || !Y_dev) fail(HMM_EINVAL, "null x_dev / Y_dev");
        if (T < 0) fail(HMM_EINVAL, "negative T");
        check_model_small(states, N, K, nstates, mu);
        require_device();
        if (T == 0) return;
        std::vector<double> m;
        state_means(states, N, K, nstates, mu, m);
        reconstruct_run(x_dev, T, m, Y_dev, main_stream());
    });
}

int hmm_unroll_mlseq_i16(const int16_t *x, int64_t T, const int16_t *states, int32_t N, int32_t nstates,
                         int16_t *out) {
    return guarded([&] {
        if (!x || !out || !states) fail(HMM_EINVAL, "null argument");
        if (T < 0 || N < 1 || nstates < 1) fail(HMM_EINVAL, "bad sizes");
        require_device();
        if (T == 0) return;
        cudaStream_t st = main_stream();
        Workspace &ws = workspace();
        int16_t *x_dev = (int16_t *)ws.get(Workspace::X, sizeof(int16_t) * (size_t)T);
        int16_t *o_dev = (int16_t *)ws.get(Workspace::T2, sizeof(int16_t) * (size_t)T * N);
        h2d(x_dev, x, sizeof(int16_t) * (size_t)T, st);
        unroll_run(x_dev, T, states, N, nstates, o_dev, st);
        d2h(out, o_dev, sizeof(int16_t) * (size_t)T * N, st);
        HMM_CUDA(cudaStreamSynchronize(st));
    });
}

}  // extern "C"
