// Internal declarations shared by the libhicbalance translation units.
// Public contract: include/hicbalance.h.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "hicbalance.h"

#define HB_NUM_SMS 148          // B200: 2 dies x 74 SMs
#ifndef HB_SPMV_THREADS
#define HB_SPMV_THREADS 256     // 8 warps per CTA
#endif
#ifndef HB_SPMV_CTAS_PER_SM
#define HB_SPMV_CTAS_PER_SM 3   // 24 resident warps per SM, <= 80 regs/thread (deep unroll beats occupancy here)
#endif
#ifndef HB_ROWS_PER_GRAB
#define HB_ROWS_PER_GRAB 1      // rows a warp claims per atomic on the row counter
#endif
#ifndef HB_SPMV_UNROLL
#define HB_SPMV_UNROLL 16       // independent entries per lane per step of the row dot product
#endif
#define HB_TILE_ROWS 1024       // fixed reduction tile (rows) -> rank-count independent sums
#define HB_MAX_SEG 65536        // chromosomes / contigs treated as independent blocks (draft assemblies have thousands)
#define HB_MAX_PEERS 8          // peer-store exchange: ranks of one NVLink domain

extern std::atomic<uint64_t> g_hb_launches;
void hb_set_error(const char* fmt, ...);
int hb_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define HB_CUDA(expr)                                                      \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return hb_cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define HB_LAUNCH_CHECK()                                                  \
    do {                                                                   \
        g_hb_launches.fetch_add(1, std::memory_order_relaxed);             \
        cudaError_t _e = cudaGetLastError();                               \
        if (_e != cudaSuccess) return hb_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define HB_REQUIRE(cond, msg)                                              \
    do {                                                                   \
        if (!(cond)) {                                                     \
            hb_set_error("%s (%s)", msg, #cond);                           \
            return HB_ERR_INVALID_ARG;                                     \
        }                                                                  \
    } while (0)

// status words kept on the device (int32 each)
enum { HB_ST_ALL_DONE = 0, HB_ST_OVERFLOW = 1, HB_ST_ITER = 2, HB_ST_ACTIVE = 3, HB_ST_WORDS = 4 };
// pinned progress words of the persistent solver kernels: the first four mirror the status words above
enum { HB_LP_EXCHANGES = 4, HB_LP_OUTER = 5, HB_LP_MATVECS = 6, HB_LP_FLAGS = 7, HB_LOOP_WORDS = 16 };

// Symmetric region of one rank: 2 x (exchange vector | arrival flags) + the ticket counter, exported through CUDA IPC and
// mapped by every peer.  Kept in a process-wide registry keyed by (device, rank, world): attaching costs an allocation,
// a device synchronisation and world-1 cudaIpcOpenMemHandle calls the first time only.  The exchange sequence number
// belongs to the region (flags are monotone), not to the handle.
struct HbPeerCtx {
    int device = -1, world = 0, rank = 0;
    void* d_sym = nullptr;                          // this rank's region (cudaMalloc)
    size_t buf_bytes = 0, flag_off = 0, total = 0;  // layout: buf p at p*buf_bytes ; flags p at flag_off + p*512
    int64_t capacity_cols = 0;
    void* base[HB_MAX_PEERS] = {nullptr};           // every rank's region mapped here (own entry = d_sym)
    bool ipc[HB_MAX_PEERS] = {false};               // mapping came from cudaIpcOpenMemHandle (to be closed)
    unsigned char handles[HB_MAX_PEERS * 64] = {0}; // the IPC handles the mappings were opened from
    bool attached = false;
    unsigned long long seq = 0;                     // exchange sequence number (monotonic)
    unsigned int* d_done = nullptr;                 // ticket counter of the publishing SpMV
    int users = 0;                                  // handles currently bound to the region
};

struct hb_csr {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_csr = true;
    int64_t n_rows = 0, n_cols = 0, row0 = 0, nnz = 0;
    // entries are stored from slot `phase` (= global entry offset of the block mod 4) so that the 16-byte alignment
    // pattern of every row -- and with it the summation order -- is the same however the matrix is sharded
    int64_t phase = 0;
    int64_t* d_indptr = nullptr;
    int32_t* d_indices = nullptr;
    double* d_data = nullptr;
    // optional lossless narrow copy of the values for the streaming kernels (hb_csr_pack_values):
    // 0 = none, 1 = uint16 (integer counts <= 65535), 2 = float (exactly representable).  Same entry indexing as d_data.
    int pack_kind = 0;
    void* d_pack = nullptr;

    // sharded runs: the one collective of the path is delegated to the host (torch.distributed / NCCL)
    int rank = 0, world = 1;
    hb_allreduce_fn allreduce = nullptr;
    void* comm_ctx = nullptr;

    // peer-store exchange (hb_peer_export / hb_peer_attach): the symmetric region lives in a process-wide context that
    // outlives the handle, so a second handle of the same rank (next matrix, next call) re-uses the mapped regions
    bool peer = false;
    struct HbPeerCtx* pc = nullptr;

    // segments (independent diagonal blocks) and fixed reduction tiles over GLOBAL rows
    int nseg = 1;
    std::vector<int64_t> seg_off;      // nseg+1
    void* d_tile_slab = nullptr;       // one allocation holding the five tables below
    int64_t* d_seg_off = nullptr;      // nseg+1
    int n_tiles = 0;
    int64_t* d_tile_row0 = nullptr;    // n_tiles
    int32_t* d_tile_len = nullptr;     // n_tiles
    int32_t* d_tile_seg = nullptr;     // n_tiles
    int32_t* d_seg_tile0 = nullptr;    // nseg+1

    // replicated n_cols-vectors and per-segment scalars (lazily allocated)
    bool state_ready = false;
    void* d_state_slab = nullptr;      // one allocation holding every device pointer below
    void* h_pinned = nullptr;          // one pinned block: h_status | h_kr_sc
    double* d_bias = nullptr;          // n_cols
    double* d_r = nullptr;             // n_cols
    double* d_exch = nullptr;          // n_cols + 64 (single-GPU exchange buffer)
    double* d_tile_sum = nullptr;      // n_tiles
    double* d_tile_cnt = nullptr;      // n_tiles
    double* d_seg_mean = nullptr;      // nseg
    double* d_seg_corr = nullptr;      // nseg (final bias normalisation)
    unsigned long long* d_seg_devbits = nullptr;  // nseg
    double* d_seg_dev = nullptr;       // nseg (last deviation)
    int32_t* d_seg_done = nullptr;     // nseg
    int32_t* d_seg_iters = nullptr;    // nseg
    int32_t* d_status = nullptr;       // HB_ST_WORDS
    unsigned int* d_counters = nullptr;          // [0] reduce tiles done, [1] update tiles done
    unsigned long long* d_row_counter = nullptr; // dynamic row scheduler (two counters, alternating)
    int spmv_parity = 0;
    unsigned long long* d_scalar_bits = nullptr; // 8 scratch words
    int32_t* h_status = nullptr;       // pinned, 16 slots x HB_ST_WORDS
    // persistent solver kernels (hb_ice_loop_kernel / hb_kr_solve_kernel): one cooperative launch runs the whole loop
    unsigned long long* d_loop_devbits = nullptr;  // 2 x nseg: deviation maxima, ping-pong by iteration parity
    unsigned long long* d_loop_wmax = nullptr;     // 2 words: running maximum of the rescaled values, ping-pong
    double* d_loop_err = nullptr;                  // != 0: a peer did not publish in time
    int32_t* h_loop = nullptr;         // pinned + mapped, HB_LOOP_WORDS: progress written by the kernel itself
    int32_t* h_loop_dev = nullptr;     // device alias of h_loop
    bool loop_pending = false;         // a loop kernel is in flight whose exchange count is not yet added to peer_seq
    cudaEvent_t loop_event = nullptr;
    double* d_kr_buf = nullptr;        // Knight-Ruiz workspace (8 n-vectors + tile partials + scalars), kept across calls
    size_t kr_buf_words = 0;
    double* h_kr_sc = nullptr;         // pinned mirror of the KR scalars
    int64_t h2d_bytes = 0;             // bytes hb_csr_create* moved over the bus for this store
    double* d_dense = nullptr;         // hb_block_correlation: the dense block between its two calls
    int64_t dense_m = 0;
};

// Every device allocation of the library goes through these two (hb_core.cu); the peer regions alone call cudaMalloc
// directly (CUDA IPC).
cudaError_t hb_pool_alloc(void** p, size_t bytes);
cudaError_t hb_pool_free(void* p);
template <typename T>
static inline cudaError_t hb_dmalloc(T** p, size_t bytes) {
    void* q = nullptr;
    const cudaError_t e = hb_pool_alloc(&q, bytes);
    *p = static_cast<T*>(q);
    return e;
}
static inline cudaError_t hb_dfree(void* p) { return hb_pool_free(p); }

// Device scratch that is released on every exit path of a host function (early error returns included).
struct HbScratch {
    std::vector<void*> ptrs;
    template <typename T>
    cudaError_t alloc(T** p, size_t bytes) {
        void* q = nullptr;
        const cudaError_t e = hb_dmalloc(&q, bytes ? bytes : 1);
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = static_cast<T*>(q);
        return e;
    }
    ~HbScratch() {
        for (void* q : ptrs) hb_dfree(q);
    }
    HbScratch() = default;
    HbScratch(const HbScratch&) = delete;
    HbScratch& operator=(const HbScratch&) = delete;
};

// NVTX range over a phase of the path (tracing): shows up in Nsight Systems / Compute timelines; costs
// one function-pointer test when no profiler is attached (header-only NVTX v3).
struct HbRange {
    explicit HbRange(const char* name) { nvtxRangePushA(name); }
    ~HbRange() { nvtxRangePop(); }
    HbRange(const HbRange&) = delete;
    HbRange& operator=(const HbRange&) = delete;
};
#define HB_RANGE(name) HbRange hb_range_guard_(name)

int hb_build_tiles(hb_csr* h);
int hb_ensure_state(hb_csr* h);
void hb_free_state(hb_csr* h);
int hb_use_device(const hb_csr* h);
int hb_spmv_grid(int64_t n_rows);
void hb_peer_release(hb_csr* h);
int hb_kr_reserve(hb_csr* h);                        // allocate the Knight-Ruiz workspace of the handle (idempotent)
int hb_loop_settle(hb_csr* h);                       // fold a finished loop kernel's exchange count into peer_seq
int hb_coop_grid(const void* kernel, int64_t n_rows, int device, int dyn_smem = 0);   // co-resident persistent grid for a cooperative launch
void hb_drop_pack(hb_csr* h);
// multi-threaded staged copies for pageable caller buffers (hb_xfer.cu); synchronous
int hb_copy_h2d(void* d_dst, const void* h_src, size_t bytes, int device);
int hb_copy_d2h(void* h_dst, const void* d_src, size_t bytes, int device);
// value upload of hb_csr_create for raw counts (hb_xfer.cu): uint16 over the bus where the host cores keep up, fp64 where
// they do not.  *direct_entries < 0: nothing was uploaded (the caller copies); otherwise d_data is complete, *direct_entries
// of its entries crossed the bus as fp64, and *narrowed says whether d_pack holds the uint16 copy of all of them
int hb_ingest_values(double* d_data, uint16_t* d_pack, const double* h_src, int64_t nnz, int device, int* narrowed,
                     int64_t* direct_entries, const std::atomic<bool>* other_upload_done);
// exchange buffer p (0/1) of this rank's symmetric region and its flag array
static inline double* hb_sym_buf(const hb_csr* h, int p) { return (double*)((unsigned char*)h->pc->d_sym + (size_t)p * h->pc->buf_bytes); }
static inline unsigned long long* hb_sym_flags(const hb_csr* h, int p) {
    return (unsigned long long*)((unsigned char*)h->pc->d_sym + h->pc->flag_off + (size_t)p * 512);
}

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ int4 hb_ldg_stream_i4(const int32_t* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double2 hb_ldg_stream_d2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
#ifndef HB_STREAM_HINT
#define HB_STREAM_HINT 0
#endif
#if HB_STREAM_HINT == 1
#define HB_STREAM_QUAL "ld.global.nc.L1::no_allocate.L2::128B"
#elif HB_STREAM_HINT == 2
#define HB_STREAM_QUAL "ld.global.nc.L1::no_allocate.L2::256B"
#elif HB_STREAM_HINT == 3
#define HB_STREAM_QUAL "ld.global.nc.L1::no_allocate.L2::evict_first"
#else
#define HB_STREAM_QUAL "ld.global.nc.L1::no_allocate"
#endif
__device__ __forceinline__ int hb_ldg_stream_i1(const int32_t* p) {
    int r;
    asm volatile(HB_STREAM_QUAL ".s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double hb_ldg_stream_d1(const double* p) {
    double r;
    asm volatile(HB_STREAM_QUAL ".f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double hb_ldg_stream_val(const double* p) { return hb_ldg_stream_d1(p); }
__device__ __forceinline__ double hb_ldg_stream_val(const float* p) {
    float r;
    asm volatile(HB_STREAM_QUAL ".f32 %0, [%1];" : "=f"(r) : "l"(p));
    return (double)r;
}
__device__ __forceinline__ double hb_ldg_stream_val(const uint16_t* p) {
    unsigned short r;
    asm volatile(HB_STREAM_QUAL ".u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (double)r;
}
__device__ __forceinline__ double hb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double hb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long hb_warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
