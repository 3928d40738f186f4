// Handle management of the device-resident CSR store (include/hicbalance.h).
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <thread>

#include "hb_internal.cuh"

std::atomic<uint64_t> g_hb_launches{0};
static thread_local char g_err[512] = "";

void hb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hb_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    hb_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return HB_ERR_NO_DEVICE;
    return HB_ERR_CUDA;
}

int hb_use_device(const hb_csr* h) {
    HB_CUDA(cudaSetDevice(h->device));
    return HB_OK;
}

int hb_spmv_grid(int64_t n_rows) {
    // persistent grid: a multiple of the SM count; shrink for tiny matrices
    int64_t warps_needed = (n_rows + HB_ROWS_PER_GRAB - 1) / HB_ROWS_PER_GRAB;
    int64_t ctas_needed = (warps_needed + (HB_SPMV_THREADS / 32) - 1) / (HB_SPMV_THREADS / 32);
    int64_t full = (int64_t)HB_NUM_SMS * HB_SPMV_CTAS_PER_SM;
    if (ctas_needed >= full) return (int)full;
    return (int)(ctas_needed < 1 ? 1 : ctas_needed);
}

// ---- device memory --------------------------------------------------------------------------------------------------
// One place for every allocation of the library.  A stream-ordered pool (cudaMallocAsync, release threshold raised) was
// tried here in round 2: it saves ~1.5 ms per per-chromosome handle, but its first use costs ~125 ms per process and
// growing / trimming it by tens of gigabytes is slower than mapping the stores directly (bin deletion on the 10 kb
// matrix: 150-190 ms instead of 24 ms) -- so this stays plain cudaMalloc / cudaFree.
cudaError_t hb_pool_alloc(void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 1); }
cudaError_t hb_pool_free(void* p) { return p != nullptr ? cudaFree(p) : cudaSuccess; }

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }
const char* hb_last_error(void) { return g_err; }
uint64_t hb_launch_count(void) { return g_hb_launches.load(); }

int hb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hb_pinned_alloc(void** out, uint64_t bytes) {
    HB_REQUIRE(out != nullptr, "null output pointer");
    HB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return HB_OK;
}

int hb_pinned_free(void* p) {
    if (p) HB_CUDA(cudaFreeHost(p));
    return HB_OK;
}

}  // extern "C"

__global__ void hb_rebase_indptr_kernel(int64_t* indptr, int64_t n1, int64_t base) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n1) indptr[i] -= base;
}

__global__ void hb_narrow_indices_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // a value that does not fit becomes -1 and is rejected by the validation pass
    for (; i < n; i += stride) out[i] = (in[i] >= 0 && in[i] < 2147483647LL) ? (int32_t)in[i] : -1;
}

// values that arrive in the file's own type (counts are int32 / int64 in .cool files, float32 after hicNormalize) are
// widened to the fp64 store on the device: no fp64 copy on the host, fewer bytes over PCIe
template <typename T>
__global__ void hb_widen_values_kernel(const T* __restrict__ in, double* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (double)in[i];
}

// The kernels of the path trust the store: an out-of-range column id would be an out-of-bounds gather.  One streaming
// pass over the column ids (4 B per entry) checks what the contract in hicbalance.h states:
// flags bit 0: indptr decreases; bit 1: a column id outside [0, n_cols); bit 2: columns not strictly ascending in a row
__global__ void __launch_bounds__(256)
hb_validate_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t n_rows, int64_t n_cols,
                   int64_t first, int64_t last, unsigned int* flags) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned int bad = 0;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t s = indptr[row], e = indptr[row + 1];
        if (e < s || s < first || e > last) {
            bad |= 1u;
            continue;
        }
        for (int64_t k = s + lane; k < e; k += 32) {
            const int c = idx[k];
            if (c < 0 || c >= n_cols) bad |= 2u;
            if (k > s && idx[k - 1] >= c) bad |= 4u;
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicOr(flags, bad);
}

static int hb_new_handle(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz, int device) {
    HB_REQUIRE(out != nullptr, "null output handle");
    HB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && row0 >= 0, "negative size");
    HB_REQUIRE(row0 + n_rows <= n_cols, "row block exceeds the matrix");
    HB_REQUIRE(n_cols < 2147483647LL, "bin count must fit int32 column ids");
    int ndev = hb_device_count();
    if (ndev <= 0) {
        hb_set_error("no CUDA device visible: libhicbalance has no CPU fallback");
        return HB_ERR_NO_DEVICE;
    }
    HB_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    HB_CUDA(cudaSetDevice(device));
    hb_csr* h = new hb_csr();
    h->device = device;
    h->n_rows = n_rows;
    h->n_cols = n_cols;
    h->row0 = row0;
    h->nnz = nnz;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return hb_cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    h->nseg = 1;
    h->seg_off = {0, n_cols};
    *out = h;
    return HB_OK;
}

extern "C" {

int hb_csr_create(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz, const int64_t* indptr,
                  const void* indices, int indices_i64, const double* data, int device) {
    return hb_csr_create_typed(out, n_rows, n_cols, row0, nnz, indptr, indices, indices_i64, data, HB_VALUES_F64, device);
}

int hb_csr_create_typed(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz, const int64_t* indptr,
                        const void* indices, int indices_i64, const void* data_any, int value_kind, int device) {
    HB_RANGE("hb_csr_create_typed");
    HB_REQUIRE(value_kind >= HB_VALUES_F64 && value_kind <= HB_VALUES_I64, "unknown value type");
    const double* data = static_cast<const double*>(data_any);
    HB_REQUIRE(indptr != nullptr, "null indptr");
    HB_REQUIRE(nnz == 0 || (indices != nullptr && data != nullptr), "null indices/data");
    HB_REQUIRE(indptr[n_rows] - indptr[0] == nnz, "indptr does not span nnz entries");
    int rc = hb_new_handle(out, n_rows, n_cols, row0, nnz, device);
    if (rc != HB_OK) return rc;
    hb_csr* h = *out;
    h->owns_csr = true;
    cudaStream_t s = h->stream;
    std::thread idx_thread;   // uploads the column ids beside the values; joined before any exit path frees the handle
#define HB_CREATE_CUDA(expr)                                               \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) {                                           \
            if (idx_thread.joinable()) idx_thread.join();                  \
            hb_csr_destroy(h);                                             \
            *out = nullptr;                                                \
            return hb_cuda_fail(_e, #expr, __FILE__, __LINE__);            \
        }                                                                  \
    } while (0)
    static const bool create_timing = getenv("HB_INGEST_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_stamp[6] = {now(), 0, 0, 0, 0, 0};   // start | allocations | values | column ids joined | validated | tiles
    double t_idx = 0.0;
    HB_CREATE_CUDA(hb_dmalloc(&h->d_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    const int64_t phase = indptr[0] & 3;
    h->phase = phase;
    HB_CREATE_CUDA(hb_dmalloc(&h->d_indices, sizeof(int32_t) * (size_t)(nnz + phase + 4) + 16));
    HB_CREATE_CUDA(hb_dmalloc(&h->d_data, sizeof(double) * (size_t)(nnz + phase + 4) + 16));
    // zero the alignment slots in front of the block.  Only those: the entries themselves arrive through hb_copy_h2d,
    // which runs on other streams (the handle's stream is non-blocking), so a memset that overlapped them could land
    // AFTER the copy and wipe the first entries.  The stream is drained before the copies start for the same reason.
    if (phase > 0) {
        HB_CREATE_CUDA(cudaMemsetAsync(h->d_indices, 0, sizeof(int32_t) * (size_t)phase, s));
        HB_CREATE_CUDA(cudaMemsetAsync(h->d_data, 0, sizeof(double) * (size_t)phase, s));
    }
    HB_CREATE_CUDA(cudaMemcpyAsync(h->d_indptr, indptr, sizeof(int64_t) * (size_t)(n_rows + 1),
                                   cudaMemcpyHostToDevice, s));
    if (indptr[0] != phase) {
        int64_t n1 = n_rows + 1;
        hb_rebase_indptr_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, s>>>(h->d_indptr, n1, indptr[0] - phase);
        g_hb_launches.fetch_add(1);
    }
    HB_CREATE_CUDA(cudaStreamSynchronize(s));
    t_stamp[1] = t_stamp[2] = t_stamp[3] = now();
    if (nnz > 0) {
        // int32 column ids go over the bus from a second host thread while this one deals with the values
        int idx_rc = HB_OK;
        std::atomic<bool> idx_done(indices_i64 != 0);   // int64 ids are narrowed after the values, nothing runs beside them
        auto fail_create = [&]() {
            if (idx_thread.joinable()) idx_thread.join();
            hb_csr_destroy(h);
            *out = nullptr;
            return (int)HB_ERR_CUDA;
        };
        if (!indices_i64)
            idx_thread = std::thread([&] {
                const double t0 = now();
                idx_rc = cudaSetDevice(device) == cudaSuccess
                             ? hb_copy_h2d(h->d_indices + phase, (const int32_t*)indices + indptr[0], sizeof(int32_t) * (size_t)nnz, device)
                             : HB_ERR_CUDA;
                t_idx = now() - t0;
                idx_done.store(true, std::memory_order_release);
            });
        h->h2d_bytes = (int64_t)sizeof(int64_t) * (n_rows + 1) + (int64_t)(indices_i64 ? 8 : 4) * nnz;
        if (value_kind == HB_VALUES_F64) {
            // raw counts: 2 bytes per entry cross the bus and the device keeps both forms (hb_ingest_values)
            int narrowed = 0;
            uint16_t* d_pack = nullptr;
            HB_CREATE_CUDA(hb_dmalloc(&d_pack, sizeof(uint16_t) * (size_t)(nnz + phase + 8) + 16));
            h->d_pack = d_pack;
            HB_CREATE_CUDA(cudaMemsetAsync(d_pack, 0, sizeof(uint16_t) * 8, s));
            HB_CREATE_CUDA(cudaStreamSynchronize(s));
            int64_t direct_entries = -1;
            if (hb_ingest_values(h->d_data + phase, d_pack + phase, data + indptr[0], nnz, device, &narrowed, &direct_entries,
                                 &idx_done) != HB_OK) {
                return fail_create();
            }
            if (narrowed) h->pack_kind = 1;
            else {
                hb_dfree(h->d_pack);
                h->d_pack = nullptr;
            }
            if (direct_entries < 0) {
                // not raw counts (or HB_NARROW_INGEST=0): the plain fp64 copy
                if (hb_copy_h2d(h->d_data + phase, data + indptr[0], sizeof(double) * (size_t)nnz, device) != HB_OK) {
                    return fail_create();
                }
                h->h2d_bytes += (int64_t)sizeof(double) * nnz;
            } else {
                h->h2d_bytes += (int64_t)sizeof(double) * direct_entries + (int64_t)sizeof(uint16_t) * (nnz - direct_entries);
            }
        } else {
            h->h2d_bytes += (int64_t)(value_kind == HB_VALUES_I64 ? 8 : 4) * nnz;
            const size_t elt = value_kind == HB_VALUES_I64 ? 8 : 4;
            const unsigned char* src = static_cast<const unsigned char*>(data_any) + elt * (size_t)indptr[0];
            const int64_t chunk = (int64_t)1 << 26;
            void* d_tmp = nullptr;
            HbScratch staging;   // released on every exit path
            HB_CREATE_CUDA(staging.alloc(&d_tmp, elt * (size_t)(nnz < chunk ? nnz : chunk)));
            for (int64_t o = 0; o < nnz; o += chunk) {
                const int64_t m = nnz - o < chunk ? nnz - o : chunk;
                if (hb_copy_h2d(d_tmp, src + elt * (size_t)o, elt * (size_t)m, device) != HB_OK) {
                    return fail_create();
                }
                double* dst = h->d_data + phase + o;
                if (value_kind == HB_VALUES_F32)
                    hb_widen_values_kernel<float><<<HB_NUM_SMS * 8, 256, 0, s>>>((const float*)d_tmp, dst, m);
                else if (value_kind == HB_VALUES_I32)
                    hb_widen_values_kernel<int32_t><<<HB_NUM_SMS * 8, 256, 0, s>>>((const int32_t*)d_tmp, dst, m);
                else
                    hb_widen_values_kernel<long long><<<HB_NUM_SMS * 8, 256, 0, s>>>((const long long*)d_tmp, dst, m);
                g_hb_launches.fetch_add(1);
                HB_CREATE_CUDA(cudaStreamSynchronize(s));   // the staging buffer is reused by the next chunk
            }
        }
        t_stamp[2] = now();
        if (indices_i64) {
            // krbalancing's constructor takes int64 indices (hicCorrectMatrix.py:744-747): narrow on the device
            const int64_t* src = (const int64_t*)indices + indptr[0];
            const int64_t chunk = (int64_t)1 << 26;
            int64_t* d_tmp = nullptr;
            HbScratch staging;   // released on every exit path
            HB_CREATE_CUDA(staging.alloc(&d_tmp, sizeof(int64_t) * (size_t)(nnz < chunk ? nnz : chunk)));
            for (int64_t o = 0; o < nnz; o += chunk) {
                int64_t m = nnz - o < chunk ? nnz - o : chunk;
                HB_CREATE_CUDA(cudaMemcpyAsync(d_tmp, src + o, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, s));
                hb_narrow_indices_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(d_tmp, h->d_indices + phase + o, m);
                g_hb_launches.fetch_add(1);
                HB_CREATE_CUDA(cudaStreamSynchronize(s));   // the staging buffer is reused by the next chunk
            }
        } else {
            idx_thread.join();
            if (idx_rc != HB_OK) return fail_create();
        }
        t_stamp[3] = now();
    }
    HB_CREATE_CUDA(cudaStreamSynchronize(s));
    {
        // reject a malformed block before any kernel gathers through its column ids
        unsigned int* d_flags = nullptr;
        unsigned int flags = 0;
        HB_CREATE_CUDA(hb_dmalloc(&d_flags, sizeof(unsigned int)));
        cudaError_t ve = cudaMemsetAsync(d_flags, 0, sizeof(unsigned int), s);
        if (ve == cudaSuccess && n_rows > 0) {
            hb_validate_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, n_rows, n_cols, phase, phase + nnz, d_flags);
            g_hb_launches.fetch_add(1);
        }
        if (ve == cudaSuccess) ve = cudaMemcpyAsync(&flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, s);
        if (ve == cudaSuccess) ve = cudaStreamSynchronize(s);
        hb_dfree(d_flags);
        HB_CREATE_CUDA(ve);
        if (flags) {
            hb_csr_destroy(h);
            *out = nullptr;
            hb_set_error("hb_csr_create: malformed CSR block:%s%s%s", (flags & 1u) ? " indptr is not monotone;" : "",
                         (flags & 2u) ? " a column id lies outside [0, n_cols);" : "",
                         (flags & 4u) ? " column ids are not strictly ascending inside a row (sort and sum duplicates first);" : "");
            return HB_ERR_INVALID_ARG;
        }
    }
#undef HB_CREATE_CUDA
    t_stamp[4] = now();
    rc = hb_build_tiles(h);
    if (rc != HB_OK) {
        hb_csr_destroy(h);
        *out = nullptr;
    }
    t_stamp[5] = now();
    if (create_timing)
        fprintf(stderr, "[hb create] %lld entries: allocations %.1f ms | values %.1f | wait for column ids %.1f (their copy took %.1f) | "
                        "validation %.1f | tiles %.1f | total %.1f ms\n",
                (long long)nnz, 1e3 * (t_stamp[1] - t_stamp[0]), 1e3 * (t_stamp[2] - t_stamp[1]), 1e3 * (t_stamp[3] - t_stamp[2]),
                1e3 * t_idx, 1e3 * (t_stamp[4] - t_stamp[3]), 1e3 * (t_stamp[5] - t_stamp[4]), 1e3 * (t_stamp[5] - t_stamp[0]));
    return rc;
}

int hb_dev_csr_wrap(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz, const int64_t* d_indptr,
                    const int32_t* d_indices, const double* d_data, int device) {
    HB_REQUIRE(d_indptr != nullptr, "null indptr");
    HB_REQUIRE(((uintptr_t)d_indices % 16) == 0 && ((uintptr_t)d_data % 16) == 0,
               "indices/data must be 16-byte aligned");
    int rc = hb_new_handle(out, n_rows, n_cols, row0, nnz, device);
    if (rc != HB_OK) return rc;
    hb_csr* h = *out;
    h->owns_csr = false;
    h->d_indptr = const_cast<int64_t*>(d_indptr);
    h->d_indices = const_cast<int32_t*>(d_indices);
    h->d_data = const_cast<double*>(d_data);
    rc = hb_build_tiles(h);
    if (rc != HB_OK) {
        hb_csr_destroy(h);
        *out = nullptr;
    }
    return rc;
}

int hb_csr_destroy(hb_csr* h) {
    if (!h) return HB_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    hb_free_state(h);
    if (h->owns_csr) {
        hb_dfree(h->d_indptr);
        hb_dfree(h->d_indices);
        hb_dfree(h->d_data);
    }
    hb_dfree(h->d_tile_slab);
    hb_dfree(h->d_pack);
    hb_peer_release(h);
    if (h->loop_event) cudaEventDestroy(h->loop_event);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return HB_OK;
}

int hb_csr_ingest_info(const hb_csr* h, int* pack_kind, int64_t* h2d_bytes) {
    HB_REQUIRE(h != nullptr, "null handle");
    if (pack_kind) *pack_kind = h->pack_kind;
    if (h2d_bytes) *h2d_bytes = h->h2d_bytes;
    return HB_OK;
}

int hb_csr_shape(const hb_csr* h, int64_t* n_rows, int64_t* n_cols, int64_t* row0, int64_t* nnz) {
    HB_REQUIRE(h != nullptr, "null handle");
    if (n_rows) *n_rows = h->n_rows;
    if (n_cols) *n_cols = h->n_cols;
    if (row0) *row0 = h->row0;
    if (nnz) *nnz = h->nnz;
    return HB_OK;
}

int hb_csr_download(hb_csr* h, int64_t* indptr, int32_t* indices, double* data) {
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    HB_CUDA(cudaStreamSynchronize(h->stream));
    if (indptr) {
        HB_CUDA(cudaMemcpy(indptr, h->d_indptr, sizeof(int64_t) * (size_t)(h->n_rows + 1), cudaMemcpyDeviceToHost));
        if (h->phase)
            for (int64_t i = 0; i <= h->n_rows; ++i) indptr[i] -= h->phase;
    }
    if (indices && h->nnz) {
        rc = hb_copy_d2h(indices, h->d_indices + h->phase, sizeof(int32_t) * (size_t)h->nnz, h->device);
        if (rc) return rc;
    }
    if (data && h->nnz) {
        rc = hb_copy_d2h(data, h->d_data + h->phase, sizeof(double) * (size_t)h->nnz, h->device);
        if (rc) return rc;
    }
    return HB_OK;
}

int hb_dev_csr_pointers(hb_csr* h, const int64_t** d_indptr, const int32_t** d_indices, const double** d_data) {
    HB_REQUIRE(h != nullptr, "null handle");
    if (d_indptr) *d_indptr = h->d_indptr;
    if (d_indices) *d_indices = h->d_indices + h->phase;  // entry 0 of the block; d_indptr[0] == phase
    if (d_data) *d_data = h->d_data + h->phase;
    return HB_OK;
}

int hb_sync(hb_csr* h) {
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    HB_CUDA(cudaStreamSynchronize(h->stream));
    return HB_OK;
}

int hb_set_comm(hb_csr* h, int rank, int world, hb_allreduce_fn fn, void* ctx) {
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank/world");
    HB_REQUIRE(world == 1 || fn != nullptr, "a collective callback is required when world > 1");
    h->rank = rank;
    h->world = world;
    h->allreduce = fn;
    h->comm_ctx = ctx;
    return HB_OK;
}

// ---- peer regions: process-wide registry -------------------------------------------------------------------------
static std::mutex g_peer_mutex;
static std::vector<HbPeerCtx*> g_peer_ctxs;

static void hb_peer_ctx_free(HbPeerCtx* c) {
    cudaSetDevice(c->device);
    for (int q = 0; q < HB_MAX_PEERS; ++q)
        if (c->base[q] != nullptr && c->ipc[q]) cudaIpcCloseMemHandle(c->base[q]);
    cudaFree(c->d_sym);
    delete c;
}

int hb_peer_export(hb_csr* h, int rank, int world, void* ipc_handle_out) {
    HB_RANGE("hb_peer_export");
    HB_REQUIRE(h != nullptr && ipc_handle_out != nullptr, "bad arguments");
    HB_REQUIRE(world >= 2 && world <= HB_MAX_PEERS && rank >= 0 && rank < world, "peer exchange needs 2..8 ranks");
    int rc = hb_use_device(h);
    if (rc) return rc;
    hb_peer_release(h);
    std::lock_guard<std::mutex> guard(g_peer_mutex);
    // every rank sees the same sequence of requests, so every rank takes the same decision here
    HbPeerCtx* c = nullptr;
    for (HbPeerCtx* k : g_peer_ctxs)
        if (k->device == h->device && k->rank == rank && k->world == world && k->users == 0 && k->capacity_cols >= h->n_cols) {
            c = k;
            break;
        }
    if (c == nullptr) {
        // retire idle regions of this rank that are too small, then build a new one
        for (size_t i = 0; i < g_peer_ctxs.size();) {
            HbPeerCtx* k = g_peer_ctxs[i];
            if (k->device == h->device && k->rank == rank && k->world == world && k->users == 0) {
                hb_peer_ctx_free(k);
                g_peer_ctxs.erase(g_peer_ctxs.begin() + (long)i);
            } else {
                ++i;
            }
        }
        c = new HbPeerCtx();
        c->device = h->device;
        c->rank = rank;
        c->world = world;
        c->capacity_cols = h->n_cols;
        c->buf_bytes = ((sizeof(double) * (size_t)(h->n_cols + 64)) + 255) & ~(size_t)255;
        c->flag_off = 2 * c->buf_bytes;
        c->total = 2 * c->buf_bytes + 2 * 512 + 256;
        cudaError_t e = cudaMalloc(&c->d_sym, c->total);
        if (e == cudaSuccess) e = cudaMemset(c->d_sym, 0, c->total);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();   // cudaMemset is asynchronous and the work streams are non-blocking
        if (e != cudaSuccess) {
            cudaFree(c->d_sym);
            delete c;
            return hb_cuda_fail(e, "peer region", __FILE__, __LINE__);
        }
        c->d_done = (unsigned int*)((unsigned char*)c->d_sym + 2 * c->buf_bytes + 2 * 512);
        cudaIpcMemHandle_t hd;
        e = cudaIpcGetMemHandle(&hd, c->d_sym);
        if (e != cudaSuccess) {
            cudaFree(c->d_sym);
            delete c;
            return hb_cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
        }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        memcpy(c->handles + 64 * rank, &hd, 64);
        g_peer_ctxs.push_back(c);
    }
    memcpy(ipc_handle_out, c->handles + 64 * rank, 64);
    c->users = 1;
    h->pc = c;
    h->rank = rank;
    h->world = world;
    return HB_OK;
}

int hb_peer_attach(hb_csr* h, const void* all_handles) {
    HB_RANGE("hb_peer_attach");
    HB_REQUIRE(h != nullptr && all_handles != nullptr && h->pc != nullptr, "call hb_peer_export first");
    int rc = hb_use_device(h);
    if (rc) return rc;
    HbPeerCtx* c = h->pc;
    bool same = c->attached;
    for (int q = 0; same && q < c->world; ++q)
        if (q != c->rank && memcmp(c->handles + 64 * q, (const unsigned char*)all_handles + 64 * q, 64) != 0) same = false;
    if (!same) {
        for (int q = 0; q < c->world; ++q) {
            if (c->base[q] != nullptr && c->ipc[q]) cudaIpcCloseMemHandle(c->base[q]);
            c->base[q] = nullptr;
            c->ipc[q] = false;
        }
        c->attached = false;
        for (int q = 0; q < c->world; ++q) {
            if (q == c->rank) {
                c->base[q] = c->d_sym;
                continue;
            }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const unsigned char*)all_handles + 64 * q, 64);
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                hb_peer_release(h);
                return hb_cuda_fail(e, "cudaIpcOpenMemHandle (peer access between the GPUs is required)", __FILE__, __LINE__);
            }
            c->base[q] = p;
            c->ipc[q] = true;
            memcpy(c->handles + 64 * q, &hd, 64);
        }
        c->attached = true;
    }
    h->peer = true;          // the SpMV exchange goes over peer stores; a registered callback stays for the other collectives
    return HB_OK;
}

int hb_peer_attach_inprocess(hb_csr* h, hb_csr* const* ranks) {
    HB_REQUIRE(h != nullptr && ranks != nullptr && h->pc != nullptr, "call hb_peer_export first");
    int rc = hb_use_device(h);
    if (rc) return rc;
    // Ranks of one process may share a GPU, where an allocation made by one rank (cudaMalloc synchronises the device)
    // would wait for another rank's persistent kernel -- which in turn waits for this rank's flags.  Everything the
    // solvers need is therefore reserved now, before any of them runs.
    rc = hb_kr_reserve(h);
    if (rc) return rc;
    HbPeerCtx* c = h->pc;
    for (int q = 0; q < h->world; ++q) {
        hb_csr* o = ranks[q];
        HB_REQUIRE(o != nullptr && o->pc != nullptr && o->world == h->world && o->rank == q && o->pc->buf_bytes == c->buf_bytes,
                   "every rank must have called hb_peer_export with the same world and matrix size");
        if (o->device != h->device) {
            int can = 0;
            HB_CUDA(cudaDeviceCanAccessPeer(&can, h->device, o->device));
            HB_REQUIRE(can, "no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return hb_cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        }
        if (c->base[q] != nullptr && c->ipc[q]) cudaIpcCloseMemHandle(c->base[q]);
        c->base[q] = o->pc->d_sym;
        c->ipc[q] = false;
    }
    c->attached = false;   // raw pointers of this process: never matched against IPC handles later
    h->peer = true;
    return HB_OK;
}

int hb_peer_active(const hb_csr* h) { return (h != nullptr && h->peer) ? 1 : 0; }

int hb_csr_keep_rows(hb_csr* h, int64_t row_begin, int64_t row_end) {
    HB_RANGE("hb_csr_keep_rows");
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_csr_keep_rows starts from the full matrix");
    HB_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= h->n_rows, "bad row range");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);
    const int64_t nb = row_end - row_begin;
    int64_t ends[2] = {0, 0};
    HB_CUDA(cudaMemcpyAsync(&ends[0], h->d_indptr + row_begin, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(&ends[1], h->d_indptr + row_end, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    const int64_t nnz_b = ends[1] - ends[0];
    int64_t* d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    cudaError_t e = hb_dmalloc(&d_ptr, sizeof(int64_t) * (size_t)(nb + 1));
    if (e == cudaSuccess) e = hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(nnz_b + 4) + 16);
    if (e == cudaSuccess) e = hb_dmalloc(&d_val, sizeof(double) * (size_t)(nnz_b + 4) + 16);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_ptr, h->d_indptr + row_begin, sizeof(int64_t) * (size_t)(nb + 1), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && ends[0] != 0) {
        hb_rebase_indptr_kernel<<<(unsigned)((nb + 1 + 255) / 256), 256, 0, s>>>(d_ptr, nb + 1, ends[0]);
        g_hb_launches.fetch_add(1);
    }
    if (e == cudaSuccess && nnz_b > 0) {
        e = cudaMemcpyAsync(d_idx, h->d_indices + ends[0], sizeof(int32_t) * (size_t)nnz_b, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_val, h->d_data + ends[0], sizeof(double) * (size_t)nnz_b, cudaMemcpyDeviceToDevice, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        hb_dfree(d_ptr);
        hb_dfree(d_idx);
        hb_dfree(d_val);
        return hb_cuda_fail(e, "keep rows", __FILE__, __LINE__);
    }
    if (h->owns_csr) {
        hb_dfree(h->d_indptr);
        hb_dfree(h->d_indices);
        hb_dfree(h->d_data);
    }
    h->owns_csr = true;
    h->d_indptr = d_ptr;
    h->d_indices = d_idx;
    h->d_data = d_val;
    h->phase = 0;
    h->row0 = row_begin;
    h->n_rows = nb;
    h->nnz = nnz_b;
    return HB_OK;
}

int hb_set_segments(hb_csr* h, int nseg, const int64_t* seg_offsets) {
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(nseg >= 1 && nseg <= HB_MAX_SEG && seg_offsets != nullptr, "bad segment count");
    HB_REQUIRE(seg_offsets[0] == 0 && seg_offsets[nseg] == h->n_cols, "segments must cover [0, n_cols)");
    for (int s = 0; s < nseg; ++s) HB_REQUIRE(seg_offsets[s] <= seg_offsets[s + 1], "segment offsets must ascend");
    int rc = hb_use_device(h);
    if (rc) return rc;
    HB_CUDA(cudaStreamSynchronize(h->stream));
    hb_free_state(h);
    h->nseg = nseg;
    h->seg_off.assign(seg_offsets, seg_offsets + nseg + 1);
    return hb_build_tiles(h);
}

}  // extern "C"

void hb_peer_release(hb_csr* h) {
    if (h->loop_pending) hb_loop_settle(h);
    if (h->pc != nullptr) {
        // the region stays in the registry (mapped on every peer) for the next handle of this rank
        std::lock_guard<std::mutex> guard(g_peer_mutex);
        h->pc->users = 0;
        h->pc = nullptr;
    }
    h->peer = false;
}

// Fixed reduction tiles: HB_TILE_ROWS consecutive GLOBAL rows, never crossing a segment boundary.
// Sums over rows are always formed tile by tile in this order, so the result does not depend on
// how many GPUs share the matrix.  All tile/segment tables live in ONE device slab (allocation count
// matters: --perchr creates and destroys a handle per chromosome).
static inline size_t hb_align256(size_t b) { return (b + 255) & ~(size_t)255; }

int hb_build_tiles(hb_csr* h) {
    int rc = hb_use_device(h);
    if (rc) return rc;
    std::vector<int64_t> row0;
    std::vector<int32_t> len, seg, seg_tile0(h->nseg + 1, 0);
    for (int s = 0; s < h->nseg; ++s) {
        seg_tile0[s] = (int32_t)row0.size();
        for (int64_t r = h->seg_off[s]; r < h->seg_off[s + 1]; r += HB_TILE_ROWS) {
            int64_t e = r + HB_TILE_ROWS < h->seg_off[s + 1] ? r + HB_TILE_ROWS : h->seg_off[s + 1];
            row0.push_back(r);
            len.push_back((int32_t)(e - r));
            seg.push_back(s);
        }
    }
    seg_tile0[h->nseg] = (int32_t)row0.size();
    h->n_tiles = (int)row0.size();
    hb_dfree(h->d_tile_slab);
    h->d_tile_slab = nullptr;
    const size_t nt = h->n_tiles > 0 ? (size_t)h->n_tiles : 1, ns1 = (size_t)h->nseg + 1;
    const size_t o_seg = 0;
    const size_t o_row0 = o_seg + hb_align256(sizeof(int64_t) * ns1);
    const size_t o_len = o_row0 + hb_align256(sizeof(int64_t) * nt);
    const size_t o_tseg = o_len + hb_align256(sizeof(int32_t) * nt);
    const size_t o_st0 = o_tseg + hb_align256(sizeof(int32_t) * nt);
    const size_t total = o_st0 + hb_align256(sizeof(int32_t) * ns1);
    std::vector<unsigned char> host(total, 0);
    memcpy(host.data() + o_seg, h->seg_off.data(), sizeof(int64_t) * ns1);
    if (h->n_tiles) {
        memcpy(host.data() + o_row0, row0.data(), sizeof(int64_t) * nt);
        memcpy(host.data() + o_len, len.data(), sizeof(int32_t) * nt);
        memcpy(host.data() + o_tseg, seg.data(), sizeof(int32_t) * nt);
    }
    memcpy(host.data() + o_st0, seg_tile0.data(), sizeof(int32_t) * ns1);
    HB_CUDA(hb_dmalloc(&h->d_tile_slab, total));
    HB_CUDA(cudaMemcpy(h->d_tile_slab, host.data(), total, cudaMemcpyHostToDevice));
    unsigned char* base = (unsigned char*)h->d_tile_slab;
    h->d_seg_off = (int64_t*)(base + o_seg);
    h->d_tile_row0 = (int64_t*)(base + o_row0);
    h->d_tile_len = (int32_t*)(base + o_len);
    h->d_tile_seg = (int32_t*)(base + o_tseg);
    h->d_seg_tile0 = (int32_t*)(base + o_st0);
    return HB_OK;
}

void hb_free_state(hb_csr* h) {
    if (h->loop_pending) hb_loop_settle(h);
    hb_dfree(h->d_state_slab);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    hb_dfree(h->d_kr_buf);
    hb_dfree(h->d_dense);
    h->d_dense = nullptr;
    h->dense_m = 0;
    h->d_state_slab = nullptr;
    h->h_pinned = nullptr;
    h->d_kr_buf = nullptr;
    h->kr_buf_words = 0;
    h->h_kr_sc = nullptr;
    h->d_bias = h->d_r = h->d_exch = h->d_tile_sum = h->d_tile_cnt = h->d_seg_mean = h->d_seg_corr = h->d_seg_dev = nullptr;
    h->d_seg_devbits = nullptr;
    h->d_seg_done = h->d_seg_iters = h->d_status = nullptr;
    h->d_counters = nullptr;
    h->d_row_counter = nullptr;
    h->d_scalar_bits = nullptr;
    h->h_status = nullptr;
    h->d_loop_devbits = h->d_loop_wmax = nullptr;
    h->d_loop_err = nullptr;
    h->h_loop = h->h_loop_dev = nullptr;
    h->state_ready = false;
}

// A persistent solver kernel reports how many exchanges it performed (it may stop early on convergence); the host
// folds that into the handle's sequence number before the next exchange is issued.
int hb_loop_settle(hb_csr* h) {
    if (!h->loop_pending) return HB_OK;
    HB_CUDA(cudaEventSynchronize(h->loop_event));
    if (h->pc != nullptr) h->pc->seq += (unsigned long long)(unsigned int)((volatile int32_t*)h->h_loop)[HB_LP_EXCHANGES];
    h->loop_pending = false;
    return HB_OK;
}

// Grid of a cooperative (grid-synchronising) persistent kernel: what hb_spmv_grid asks for, capped by what is
// co-resident on this device for THIS kernel.
int hb_coop_grid(const void* kernel, int64_t n_rows, int device, int dyn_smem) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, HB_SPMV_THREADS, (size_t)dyn_smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms < 1) {
        cudaGetLastError();
        sms = HB_NUM_SMS;
    }
    if (per_sm > HB_SPMV_CTAS_PER_SM) per_sm = HB_SPMV_CTAS_PER_SM;
    const int want = hb_spmv_grid(n_rows);
    const int cap = per_sm * sms;
    return want < cap ? want : cap;
}

// replicated vectors + per-segment scalars: one device slab, one pinned block
int hb_ensure_state(hb_csr* h) {
    if (h->state_ready) return HB_OK;
    int rc = hb_use_device(h);
    if (rc) return rc;
    const size_t n = (size_t)(h->n_cols > 0 ? h->n_cols : 1);
    const size_t nt = (size_t)(h->n_tiles > 0 ? h->n_tiles : 1);
    const size_t ns = (size_t)h->nseg;
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        size_t o = off;
        off += hb_align256(bytes);
        return o;
    };
    const size_t o_bias = take(sizeof(double) * n), o_r = take(sizeof(double) * n), o_exch = take(sizeof(double) * (n + 64));
    const size_t o_tsum = take(sizeof(double) * nt), o_tcnt = take(sizeof(double) * nt);
    const size_t o_mean = take(sizeof(double) * ns), o_corr = take(sizeof(double) * ns), o_dev = take(sizeof(double) * ns);
    const size_t o_devb = take(sizeof(unsigned long long) * ns);
    const size_t o_ldevb = take(sizeof(unsigned long long) * 2 * ns);
    const size_t o_done = take(sizeof(int32_t) * ns), o_iters = take(sizeof(int32_t) * ns);
    // status (16 B) | counters (32 B) | row counters (16 B) | scalar bits (64 B) | loop wmax (16 B) | loop err (8 B)
    const size_t o_small = take(256);
    const size_t total = off;
    HB_CUDA(hb_dmalloc(&h->d_state_slab, total));
    unsigned char* b = (unsigned char*)h->d_state_slab;
    h->d_bias = (double*)(b + o_bias);
    h->d_r = (double*)(b + o_r);
    h->d_exch = (double*)(b + o_exch);
    h->d_tile_sum = (double*)(b + o_tsum);
    h->d_tile_cnt = (double*)(b + o_tcnt);
    h->d_seg_mean = (double*)(b + o_mean);
    h->d_seg_corr = (double*)(b + o_corr);
    h->d_seg_dev = (double*)(b + o_dev);
    h->d_seg_devbits = (unsigned long long*)(b + o_devb);
    h->d_seg_done = (int32_t*)(b + o_done);
    h->d_seg_iters = (int32_t*)(b + o_iters);
    h->d_status = (int32_t*)(b + o_small);
    h->d_counters = (unsigned int*)(b + o_small + 32);
    h->d_row_counter = (unsigned long long*)(b + o_small + 64);
    h->d_scalar_bits = (unsigned long long*)(b + o_small + 128);
    h->d_loop_wmax = (unsigned long long*)(b + o_small + 192);
    h->d_loop_err = (double*)(b + o_small + 208);
    h->d_loop_devbits = (unsigned long long*)(b + o_ldevb);
    HB_CUDA(cudaMemset(b + o_small, 0, 256));
    HB_CUDA(cudaDeviceSynchronize());   // cudaMemset is asynchronous and the work streams are non-blocking: finish it here
    HB_CUDA(cudaHostAlloc(&h->h_pinned, 1024, cudaHostAllocMapped));
    memset(h->h_pinned, 0, 1024);
    h->h_status = (int32_t*)h->h_pinned;                              // 16 slots x HB_ST_WORDS x 4 B = 256 B
    h->h_kr_sc = (double*)((unsigned char*)h->h_pinned + 512);       // 16 doubles
    h->h_loop = (int32_t*)((unsigned char*)h->h_pinned + 768);       // HB_LOOP_WORDS x 4 B
    {
        void* dp = nullptr;
        HB_CUDA(cudaHostGetDevicePointer(&dp, h->h_loop, 0));
        h->h_loop_dev = (int32_t*)dp;
    }
    if (h->loop_event == nullptr) HB_CUDA(cudaEventCreateWithFlags(&h->loop_event, cudaEventDisableTiming));
    h->state_ready = true;
    return HB_OK;
}
