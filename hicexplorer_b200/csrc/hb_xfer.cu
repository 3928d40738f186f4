// Host <-> device transfers for caller-owned PAGEABLE buffers (numpy arrays behind the Python drop-ins).
// cudaMemcpy on pageable memory goes through a single-threaded driver staging path (measured here: 3.3 GB/s for a
// fresh 11.5 GB destination, i.e. 3.4 s to hand the corrected matrix back -- 25x the whole ICE solve).  These
// helpers run HB_XFER_WORKERS host threads, each owning two pinned staging buffers and a stream: a worker copies
// (memcpy) one chunk while its other buffer is in flight over PCIe, and the workers' first-touch page faults and
// memcpys proceed in parallel.  Pinned / registered caller buffers (bench.py's e2e leg) take the direct DMA path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include "hb_internal.cuh"

#include <cstdlib>

// measured on the B200 boxes (tools/d2h_probe.py, 2.2 GB into a fresh numpy array): 6 workers x 16 MB 22 GB/s, 12 x 4 MB
// 33 GB/s, 16 x 4 MB 34 GB/s; into pages that exist already 34 / 40 / 42 GB/s.  The ranks of a node share the cores.
#define HB_XFER_WORKERS_DEFAULT 12
#define HB_XFER_CHUNK_DEFAULT ((size_t)4 << 20)

// tunables (environment, read once): HB_XFER_WORKERS = host threads, HB_XFER_CHUNK_MB = staging chunk size
static int hb_xfer_workers() {
    static const int v = [] {
        const char* e = getenv("HB_XFER_WORKERS");
        int w = e ? atoi(e) : 0;
        if (w <= 0) {
            int hw = (int)std::thread::hardware_concurrency();
            if (hw <= 0) hw = 8;
            const char* lw = getenv("LOCAL_WORLD_SIZE");       // torchrun
            const int ranks = lw && atoi(lw) > 0 ? atoi(lw) : 1;
            w = hw / ranks;
            if (w > HB_XFER_WORKERS_DEFAULT) w = HB_XFER_WORKERS_DEFAULT;
            if (w < 2) w = 2;
        }
        return w < 1 ? 1 : (w > 64 ? 64 : w);
    }();
    return v;
}
static size_t hb_xfer_chunk() {
    static const size_t v = [] {
        const char* e = getenv("HB_XFER_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return mb >= 1 && mb <= 1024 ? (size_t)mb << 20 : HB_XFER_CHUNK_DEFAULT;
    }();
    return v;
}

static bool hb_is_device_accessible_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Staging resources live for the life of the process (pinning memory costs milliseconds per buffer, more than a
// small transfer itself); one transfer at a time uses them.
struct HbXferWorker {
    int device = -1;
    size_t chunk = 0;
    cudaStream_t st = nullptr;
    unsigned char* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
};
static std::mutex g_xfer_mutex;
static HbXferWorker g_xfer_worker[64];

static void hb_xfer_teardown(HbXferWorker& w) {
    if (w.device >= 0) cudaSetDevice(w.device);
    for (int b = 0; b < 2; ++b) {
        if (w.done[b]) cudaEventDestroy(w.done[b]);
        if (w.stage[b]) cudaFreeHost(w.stage[b]);
        w.done[b] = nullptr;
        w.stage[b] = nullptr;
    }
    if (w.st) cudaStreamDestroy(w.st);
    w.st = nullptr;
    w.device = -1;
    w.chunk = 0;
}

// A worker is marked ready (device / chunk recorded) only after EVERY allocation succeeded; a partial failure
// (e.g. a transient pinned-memory shortage) releases what was taken, so the next transfer rebuilds the worker
// instead of copying into a null staging buffer.
static cudaError_t hb_xfer_prepare(HbXferWorker& w, int device, size_t chunk) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (w.device == device && w.chunk == chunk && w.st != nullptr) return cudaSuccess;
    hb_xfer_teardown(w);  // another device or chunk size (or a half-built worker): rebuild
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaHostAlloc((void**)&w.stage[b], chunk, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w.done[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        w.device = device;   // so that teardown selects the right device
        hb_xfer_teardown(w);
        cudaGetLastError();
        return e;
    }
    w.device = device;
    w.chunk = chunk;
    return cudaSuccess;
}

// dir = 0: host -> device, 1: device -> host
static int hb_xfer(void* dst, const void* src, size_t bytes, int dir, int device) {
    if (bytes == 0) return HB_OK;
    const void* host = dir == 0 ? src : dst;
    if (bytes < ((size_t)8 << 20)) {
        HB_CUDA(cudaMemcpy(dst, src, bytes, dir == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
        return HB_OK;
    }
    if (hb_is_device_accessible_host(host)) {
        // direct DMA, on a stream of its own: a synchronous cudaMemcpy would go through the null stream and stall the API
        // calls of every other host thread (the narrow ingest's workers) for as long as it runs
        cudaStream_t st = nullptr;
        HB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        // in pieces, at most two of them queued: the copy engine serves its queue in order, so a copy queued whole would
        // hold back every chunk the other streams submit while it runs
        const size_t piece = (size_t)16 << 20;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaError_t e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        size_t k = 0;
        for (size_t off = 0; off < bytes && e == cudaSuccess; off += piece, ++k) {
            if (k >= 2) e = cudaEventSynchronize(ev[k & 1]);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync((unsigned char*)dst + off, (const unsigned char*)src + off, std::min(piece, bytes - off),
                                    dir == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        for (int b = 0; b < 2; ++b)
            if (ev[b]) cudaEventDestroy(ev[b]);
        cudaStreamDestroy(st);
        HB_CUDA(e);
        return HB_OK;
    }
    std::lock_guard<std::mutex> guard(g_xfer_mutex);
    const size_t HB_XFER_CHUNK = hb_xfer_chunk();
    const size_t nchunks = (bytes + HB_XFER_CHUNK - 1) / HB_XFER_CHUNK;
    const int workers = (int)std::min<size_t>((size_t)hb_xfer_workers(), nchunks);
    std::vector<cudaError_t> err((size_t)workers, cudaSuccess);
    auto work = [&](int t) {
        HbXferWorker& w = g_xfer_worker[t];
        cudaError_t e = hb_xfer_prepare(w, device, HB_XFER_CHUNK);
        cudaStream_t st = w.st;
        unsigned char** stage = w.stage;
        cudaEvent_t* done = w.done;
        size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
        int slot = 0;
        for (size_t c = (size_t)t; c < nchunks && e == cudaSuccess; c += (size_t)workers, slot ^= 1) {
            const size_t off = c * HB_XFER_CHUNK;
            const size_t len = std::min(HB_XFER_CHUNK, bytes - off);
            if (pending_len[slot]) {  // the buffer is reused: its previous transfer must be complete
                e = cudaEventSynchronize(done[slot]);
                if (e == cudaSuccess && dir == 1)
                    memcpy((unsigned char*)dst + pending_off[slot], stage[slot], pending_len[slot]);
                pending_len[slot] = 0;
                if (e != cudaSuccess) break;
            }
            if (dir == 0) {
                memcpy(stage[slot], (const unsigned char*)src + off, len);
                e = cudaMemcpyAsync((unsigned char*)dst + off, stage[slot], len, cudaMemcpyHostToDevice, st);
            } else {
                e = cudaMemcpyAsync(stage[slot], (const unsigned char*)src + off, len, cudaMemcpyDeviceToHost, st);
            }
            if (e == cudaSuccess) e = cudaEventRecord(done[slot], st);
            pending_off[slot] = off;
            pending_len[slot] = len;
        }
        for (int b = 0; b < 2; ++b) {
            if (pending_len[b]) {   // always drain: the staging buffers outlive this call
                const cudaError_t e2 = cudaEventSynchronize(done[b]);
                if (e == cudaSuccess) e = e2;
                if (e == cudaSuccess && dir == 1) memcpy((unsigned char*)dst + pending_off[b], stage[b], pending_len[b]);
            }
        }
        err[(size_t)t] = e;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < workers; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < workers; ++t)
        if (err[(size_t)t] != cudaSuccess) return hb_cuda_fail(err[(size_t)t], "staged transfer", __FILE__, __LINE__);
    return HB_OK;
}

int hb_copy_h2d(void* d_dst, const void* h_src, size_t bytes, int device) { return hb_xfer(d_dst, h_src, bytes, 0, device); }
int hb_copy_d2h(void* h_dst, const void* d_src, size_t bytes, int device) { return hb_xfer(h_dst, d_src, bytes, 1, device); }


// ---- narrow ingest: fp64 counts leave the host as uint16 ------------------------------------------------------
// hb_csr_create's value upload when the values are raw Hi-C counts (integers in [0, 65535]).  Two ways lead from the
// caller's fp64 array to the device, and they use different resources:
//   narrow  host threads convert chunk after chunk to uint16 in pinned staging (hb_hostpack.cpp; lossless or refused),
//           2 bytes per entry cross PCIe, the device writes the fp64 store (d_data) and keeps the uint16 stream (d_pack);
//   direct  the chunk is copied as it is, 8 bytes per entry, no host core involved (pinned / registered sources only).
// Which one is faster depends on how many cores this rank has against how fast its link is (one rank with 16 threads:
// narrow, 0.19 s instead of 0.33 s for 1.5e9 entries; 8 ranks sharing 32 threads: direct, 0.10 s instead of 0.135 s),
// so both run at once on the same array: the workers take chunks from the FRONT, a feeder thread takes chunks from the
// BACK and sends them directly once the column ids are through and as long as no worker finds its staging buffer still
// waiting for the link (= the cores are the bottleneck, not the bus), and they meet wherever the machine puts the balance.  The directly copied tail is converted to uint16 on the
// device afterwards (one pass over HBM), so the store ends up in the same two layouts either way.
// A value that does not fit ends the narrow side only: what it has delivered is valid fp64 already, the feeder takes
// the rest, and *narrowed = 0 tells the caller that no uint16 copy exists.  *direct_entries = entries sent as fp64.
extern "C" int hb_host_pack_u16(const double* src, uint16_t* dst, size_t n);

// HB_NARROW_INGEST=0 switches the narrow side off (everything goes the direct / staged way, as in round 1); read at
// every call so that A/B runs can toggle it inside one process.
static bool hb_narrow_ingest_enabled() {
    const char* e = getenv("HB_NARROW_INGEST");
    return !(e && e[0] == '0');
}
// HB_INGEST_DIRECT=1 / 0 forces the feeder on / off.  Unset: on when this rank converts with at most 8 threads.  With 16
// threads on one GPU the converters alone keep the link busy and the feeder's 8-byte entries only compete with them for
// host memory bandwidth (measured: 0.175 s without, 0.175-0.21 s with); with 8 / 4 / 2 threads it is what makes the
// upload 0.127 / 0.149 / 0.167 s instead of 0.152 / 0.235 / 0.40 s (9.6 GB of values, tools/ingest_probe.py).
static bool hb_ingest_feeder_enabled(int workers) {
    const char* e = getenv("HB_INGEST_DIRECT");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return workers <= 8;
}
static int hb_ingest_workers() {
    const char* e = getenv("HB_INGEST_WORKERS");
    if (e && atoi(e) > 0) return atoi(e) > 64 ? 64 : atoi(e);
    static const int v = [] {
        int hw = (int)std::thread::hardware_concurrency();
        if (hw <= 0) hw = 8;
        const char* lw = getenv("LOCAL_WORLD_SIZE");          // torchrun: the ranks of this node share the cores
        const int ranks = lw && atoi(lw) > 0 ? atoi(lw) : 1;
        int w = hw / ranks;
        return w < 2 ? 2 : (w > 16 ? 16 : w);
    }();
    return v;
}
static size_t hb_ingest_chunk_entries() {
    static const size_t v = [] {
        const char* e = getenv("HB_INGEST_CHUNK");             // entries per chunk (tests use a small one)
        const long n = e ? atol(e) : 0;
        return n >= 64 ? (size_t)n : ((size_t)1 << 20);   // 8 MB of fp64 read, 2 MB sent; pinning the staging costs ~0.8 ms/MB once
    }();
    return v;
}

__global__ void __launch_bounds__(256) hb_expand_u16_kernel(const uint16_t* __restrict__ src, double* __restrict__ dst, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (double)src[i];
}

// the directly copied entries: fp64 -> uint16 on the device, *bad != 0 if one of them is not a count <= 65535
__global__ void __launch_bounds__(256)
hb_narrow_on_device_kernel(const double* __restrict__ src, uint16_t* __restrict__ dst, int64_t n, unsigned int* bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool b = false;
    for (; i < n; i += stride) {
        const double v = src[i];
        const unsigned int u = (v >= 0.0 && v <= 65535.0) ? (unsigned int)v : 0u;
        b |= __double_as_longlong((double)u) != __double_as_longlong(v);   // fractions, -0.0, NaN, out of range
        dst[i] = (uint16_t)u;
    }
    if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

struct HbIngestWorker {
    int device = -1;
    size_t chunk = 0;
    cudaStream_t st = nullptr;
    uint16_t* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
};
static std::mutex g_ingest_mutex;
static HbIngestWorker g_ingest_worker[64];

static void hb_ingest_teardown(HbIngestWorker& w) {
    if (w.device >= 0) cudaSetDevice(w.device);
    for (int b = 0; b < 2; ++b) {
        if (w.done[b]) cudaEventDestroy(w.done[b]);
        if (w.stage[b]) cudaFreeHost(w.stage[b]);
        w.done[b] = nullptr;
        w.stage[b] = nullptr;
    }
    if (w.st) cudaStreamDestroy(w.st);
    w.st = nullptr;
    w.device = -1;
    w.chunk = 0;
}

static cudaError_t hb_ingest_prepare(HbIngestWorker& w, int device, size_t chunk) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (w.device == device && w.chunk == chunk && w.st != nullptr) return cudaSuccess;
    hb_ingest_teardown(w);
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaHostAlloc((void**)&w.stage[b], chunk * sizeof(uint16_t), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w.done[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        w.device = device;
        hb_ingest_teardown(w);
        cudaGetLastError();
        return e;
    }
    w.device = device;
    w.chunk = chunk;
    return cudaSuccess;
}

int hb_ingest_values(double* d_data, uint16_t* d_pack, const double* h_src, int64_t nnz, int device, int* narrowed,
                     int64_t* direct_entries, const std::atomic<bool>* other_upload_done) {
    *narrowed = 0;
    *direct_entries = -1;          // -1: nothing was uploaded here, the caller copies the values itself
    if (nnz <= 0 || !hb_narrow_ingest_enabled()) return HB_OK;
    {
        // a first look before any thread or staging buffer is set up: corrected / normalised matrices fail here
        uint16_t probe[256];
        const size_t np = (size_t)(nnz < 256 ? nnz : 256);
        if (!hb_host_pack_u16(h_src, probe, np) || !hb_host_pack_u16(h_src + (nnz - (int64_t)np), probe, np)) return HB_OK;
    }
    std::lock_guard<std::mutex> guard(g_ingest_mutex);
    const size_t chunk = hb_ingest_chunk_entries();
    const int64_t nchunks = (int64_t)(((size_t)nnz + chunk - 1) / chunk);
    const int workers = (int)std::min<int64_t>((int64_t)hb_ingest_workers(), nchunks);
    const bool feeder_on = hb_ingest_feeder_enabled(hb_ingest_workers()) && hb_is_device_accessible_host(h_src) && nchunks >= 4;
    std::vector<cudaError_t> err((size_t)workers + 1, cudaSuccess);
    // chunks [lo, hi) are unclaimed: the workers take lo++, the feeder takes --hi
    std::mutex range_mutex;
    int64_t lo = 0, hi = nchunks;
    auto take_front = [&]() -> int64_t {
        std::lock_guard<std::mutex> g(range_mutex);
        return lo < hi ? lo++ : -1;
    };
    auto take_back = [&]() -> int64_t {
        std::lock_guard<std::mutex> g(range_mutex);
        return lo < hi ? --hi : -1;
    };
    std::atomic<bool> refused(false);
    std::atomic<int> workers_running(workers);
    std::atomic<int> link_waiters(0);                    // workers whose staging buffer is still waiting for the link
    std::atomic<long long> convert_us(0), blocked_us(0); // what the workers have spent converting / waiting for the link
    std::vector<int64_t> redo;                           // chunks a worker had claimed when it met a value that does not fit
    std::mutex redo_mutex;
    std::atomic<uint64_t> launches(0);
    static const bool timing = getenv("HB_INGEST_TIMING") != nullptr;
    std::vector<double> t_prep((size_t)workers, 0.0), t_pack((size_t)workers, 0.0), t_wait((size_t)workers, 0.0),
        t_issue((size_t)workers, 0.0);
    std::vector<int64_t> n_chunks_w((size_t)workers, 0);
    int64_t n_direct_chunks = 0;
    long long held[3] = {0, 0, 0};   // feeder polls that found: column ids still going / workers mostly waiting / a worker waiting
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    auto work = [&](int t) {
        HbIngestWorker& w = g_ingest_worker[t];
        double t0 = now();
        cudaError_t e = hb_ingest_prepare(w, device, chunk);
        t_prep[(size_t)t] = now() - t0;
        bool used[2] = {false, false};
        int slot = 0;
        while (e == cudaSuccess && !refused.load(std::memory_order_relaxed)) {
            const int64_t c = take_front();
            if (c < 0) break;
            const size_t off = (size_t)c * chunk;
            const size_t len = std::min(chunk, (size_t)nnz - off);
            t0 = now();
            if (used[slot] && cudaEventQuery(w.done[slot]) == cudaErrorNotReady) {
                // the staging buffer is reused and its copy has not gone out yet: the link is the bottleneck right now --
                // the feeder holds back while any worker is in this state
                link_waiters.fetch_add(1, std::memory_order_relaxed);
                e = cudaEventSynchronize(w.done[slot]);
                link_waiters.fetch_sub(1, std::memory_order_relaxed);
            }
            const double dt_wait = now() - t0;
            t_wait[(size_t)t] += dt_wait;
            if (e != cudaSuccess) break;
            t0 = now();
            const bool ok = hb_host_pack_u16(h_src + off, w.stage[slot], len) != 0;
            const double dt_pack = now() - t0;
            t_pack[(size_t)t] += dt_pack;
            blocked_us.fetch_add((long long)(dt_wait * 1e6), std::memory_order_relaxed);
            convert_us.fetch_add((long long)(dt_pack * 1e6), std::memory_order_relaxed);
            if (!ok) {
                refused.store(true, std::memory_order_relaxed);
                std::lock_guard<std::mutex> g(redo_mutex);
                redo.push_back(c);
                break;
            }
            t0 = now();
            e = cudaMemcpyAsync(d_pack + off, w.stage[slot], len * sizeof(uint16_t), cudaMemcpyHostToDevice, w.st);
            if (e != cudaSuccess) break;
            hb_expand_u16_kernel<<<HB_NUM_SMS * 2, 256, 0, w.st>>>(d_pack + off, d_data + off, (int64_t)len);
            launches.fetch_add(1, std::memory_order_relaxed);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaEventRecord(w.done[slot], w.st);
            used[slot] = true;
            slot ^= 1;
            ++n_chunks_w[(size_t)t];
            t_issue[(size_t)t] += now() - t0;
        }
        if (w.st) {   // always drain: the staging buffers outlive this call
            const cudaError_t e2 = cudaStreamSynchronize(w.st);
            if (e == cudaSuccess) e = e2;
        }
        err[(size_t)t] = e;
        workers_running.fetch_sub(1, std::memory_order_release);
    };
    // the feeder: whole chunks straight from the caller's (pinned) array, at most two in flight, and -- while the narrow
    // side is still producing -- only when none of its copies is waiting for the link
    auto feed = [&]() {
        cudaError_t e = cudaSetDevice(device);
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};   // blocking-sync events: the feeder must not spin on a core the workers need
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b)
            e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming | cudaEventBlockingSync);
        int64_t k = 0;
        double win_t0 = now();
        long long win_conv = 0, win_blk = 0;
        bool link_bound = false;
        while (e == cudaSuccess) {
            const bool narrow_active = workers_running.load(std::memory_order_acquire) > 0 && !refused.load(std::memory_order_relaxed);
            // the link has room for 8-byte entries only when nothing else is queueing for it: the column ids (their copy
            // always has pieces in flight until it is done) and the narrow copies
            const bool ids_pending = other_upload_done != nullptr && !other_upload_done->load(std::memory_order_acquire);
            // ... and only while the workers spend their time converting, not waiting: when more than a tenth of the last
            // few milliseconds went into waiting for staging buffers, the link is the bottleneck and every 8-byte entry
            // makes it worse.  (Windows start once the column ids are through: behind them everybody waits.)
            if (ids_pending) {
                win_t0 = now();
                win_conv = convert_us.load(std::memory_order_relaxed);
                win_blk = blocked_us.load(std::memory_order_relaxed);
            } else if (now() - win_t0 > 3e-3) {
                const long long cv = convert_us.load(std::memory_order_relaxed), bl = blocked_us.load(std::memory_order_relaxed);
                link_bound = 10 * (bl - win_blk) > (cv - win_conv);
                win_t0 = now();
                win_conv = cv;
                win_blk = bl;
            }
            if (narrow_active && (ids_pending || link_bound || link_waiters.load(std::memory_order_relaxed) > 0)) {
                ++held[ids_pending ? 0 : (link_bound ? 1 : 2)];
                std::this_thread::sleep_for(std::chrono::microseconds(50));
                continue;
            }
            const int64_t c = take_back();
            if (c < 0) {
                if (workers_running.load(std::memory_order_acquire) == 0) break;   // every chunk is claimed and delivered
                std::this_thread::sleep_for(std::chrono::microseconds(50));
                continue;
            }
            const size_t off = (size_t)c * chunk;
            const size_t len = std::min(chunk, (size_t)nnz - off);
            if (k >= 2) e = cudaEventSynchronize(ev[k & 1]);       // two in flight keep the link busy, no more: it is shared
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_data + off, h_src + off, len * sizeof(double), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], st);
            ++k;
            ++n_direct_chunks;
        }
        if (st) {
            const cudaError_t e2 = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = e2;
        }
        for (int b = 0; b < 2; ++b)
            if (ev[b]) cudaEventDestroy(ev[b]);
        if (st) cudaStreamDestroy(st);
        err[(size_t)workers] = e;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < workers; ++t) pool.emplace_back(work, t);
    std::thread feeder;
    if (feeder_on) feeder = std::thread(feed);
    work(0);
    for (auto& th : pool) th.join();
    if (feeder.joinable()) feeder.join();
    g_hb_launches.fetch_add(launches.load());
    for (int t = 0; t <= workers; ++t)
        if (err[(size_t)t] != cudaSuccess) return hb_cuda_fail(err[(size_t)t], "narrow ingest", __FILE__, __LINE__);
    // what nobody delivered: the chunk(s) at which a worker met a value that does not fit, and -- without a feeder, or if
    // the feeder stopped early -- everything still unclaimed
    int64_t direct = n_direct_chunks;   // in chunks, converted to entries below
    std::vector<int64_t> rest(redo);
    for (int64_t c : redo) {
        const size_t off = (size_t)c * chunk;
        const size_t len = std::min(chunk, (size_t)nnz - off);
        if (hb_copy_h2d(d_data + off, h_src + off, len * sizeof(double), device) != HB_OK) return HB_ERR_CUDA;
        ++direct;
    }
    if (lo < hi) {   // one contiguous stretch
        const size_t off = (size_t)lo * chunk;
        const size_t end = std::min((size_t)hi * chunk, (size_t)nnz);
        if (hb_copy_h2d(d_data + off, h_src + off, (end - off) * sizeof(double), device) != HB_OK) return HB_ERR_CUDA;
        for (int64_t c = lo; c < hi; ++c) rest.push_back(c);
        direct += hi - lo;
    }
    // entries that went over the bus as fp64: whole chunks, the last chunk of the array may be short
    int64_t direct_n = direct * (int64_t)chunk;
    {
        const bool last_is_direct = hi < nchunks || std::find(rest.begin(), rest.end(), nchunks - 1) != rest.end();
        if (last_is_direct && direct > 0) direct_n -= (int64_t)nchunks * (int64_t)chunk - nnz;
    }
    *direct_entries = direct_n;
    bool ok = !refused.load();
    if (ok && hi < nchunks) {
        // the feeder's share is the tail [hi * chunk, nnz): make its uint16 copy on the device
        const size_t off = (size_t)hi * chunk;
        unsigned int* d_bad = nullptr;
        unsigned int bad = 0;
        HB_CUDA(hb_dmalloc(&d_bad, sizeof(unsigned int)));
        cudaError_t e = cudaMemset(d_bad, 0, sizeof(unsigned int));
        if (e == cudaSuccess) {
            hb_narrow_on_device_kernel<<<HB_NUM_SMS * 8, 256>>>(d_data + off, d_pack + off, (int64_t)((size_t)nnz - off), d_bad);
            g_hb_launches.fetch_add(1);
            e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
        }
        hb_dfree(d_bad);
        HB_CUDA(e);
        ok = bad == 0;
    }
    *narrowed = ok ? 1 : 0;
    if (timing) {
        auto mx = [](const std::vector<double>& v) { return *std::max_element(v.begin(), v.end()); };
        int64_t narrow_chunks = 0;
        for (int64_t c : n_chunks_w) narrow_chunks += c;
        fprintf(stderr, "[hb ingest] %lld entries, %lld chunks: %lld narrow (%d workers), %lld direct%s: total %.1f ms | per worker "
                        "(max) prepare %.1f, convert %.1f, wait for staging %.1f, issue %.1f ms | all workers: converting %.1f, "
                        "waiting %.1f ms | feeder held back %lld / %lld / %lld times (ids / link-bound / waiter)\n",
                (long long)nnz, (long long)nchunks, (long long)narrow_chunks, workers, (long long)direct,
                ok ? "" : " (some value is not a small count: no uint16 copy)", 1e3 * (now() - t_begin), 1e3 * mx(t_prep),
                1e3 * mx(t_pack), 1e3 * mx(t_wait), 1e3 * mx(t_issue), 1e-3 * (double)convert_us.load(), 1e-3 * (double)blocked_us.load(),
                held[0], held[1], held[2]);
    }
    return HB_OK;
}
