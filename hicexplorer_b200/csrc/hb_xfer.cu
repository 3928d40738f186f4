// Host <-> device transfers for caller-owned PAGEABLE buffers (numpy arrays behind the Python drop-ins).
// cudaMemcpy on pageable memory goes through a single-threaded driver staging path (measured here: 3.3 GB/s for a
// fresh 11.5 GB destination, i.e. 3.4 s to hand the corrected matrix back -- 25x the whole ICE solve).  These
// helpers run HB_XFER_WORKERS host threads, each owning two pinned staging buffers and a stream: a worker copies
// (memcpy) one chunk while its other buffer is in flight over PCIe, and the workers' first-touch page faults and
// memcpys proceed in parallel.  Pinned / registered caller buffers (bench.py's e2e leg) take the direct DMA path.
#include <algorithm>
#include <cstring>
#include <thread>

#include "hb_internal.cuh"

#define HB_XFER_WORKERS 6
#define HB_XFER_CHUNK ((size_t)16 << 20)

static bool hb_is_device_accessible_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// dir = 0: host -> device, 1: device -> host
static int hb_xfer(void* dst, const void* src, size_t bytes, int dir, int device) {
    if (bytes == 0) return HB_OK;
    const void* host = dir == 0 ? src : dst;
    if (bytes < ((size_t)8 << 20) || hb_is_device_accessible_host(host)) {
        HB_CUDA(cudaMemcpy(dst, src, bytes, dir == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
        return HB_OK;
    }
    const size_t nchunks = (bytes + HB_XFER_CHUNK - 1) / HB_XFER_CHUNK;
    const int workers = (int)std::min<size_t>(HB_XFER_WORKERS, nchunks);
    std::vector<cudaError_t> err((size_t)workers, cudaSuccess);
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        cudaStream_t st = nullptr;
        unsigned char* stage[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaHostAlloc((void**)&stage[b], HB_XFER_CHUNK, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
        }
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
            if (pending_len[b] && e == cudaSuccess) {
                e = cudaEventSynchronize(done[b]);
                if (e == cudaSuccess && dir == 1) memcpy((unsigned char*)dst + pending_off[b], stage[b], pending_len[b]);
            }
        }
        for (int b = 0; b < 2; ++b) {
            if (done[b]) cudaEventDestroy(done[b]);
            if (stage[b]) cudaFreeHost(stage[b]);
        }
        if (st) cudaStreamDestroy(st);
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
