// The one streaming kernel of the path: a row-block SpMV  t_i = sum_j A_ij u_j  with a fused
// per-row epilogue.  ICE (one iteration == one SpMV with the ORIGINAL matrix) and
// Knight-Ruiz (every CG step == one SpMV) both instantiate it.
//
// Roofline: HBM.  Algorithmic bytes per launch = 12 B per stored entry (int32 column id + fp64
// value, each read exactly once) + ~24 B per row.  The gathered vector u (n*8 B: 1-25 MB) stays
// in L2/L1; the matrix streams through with L1::no_allocate so it does not evict u.
//
// Mapping: persistent grid (148 SMs x 3 CTAs x 8 warps, 80 registers/thread); a warp owns a row,
// lanes read it lane-strided (coalesced 128 B / 256 B per instruction, 16 entries per lane in
// flight per step), partial sums are combined with __shfl_xor.  Rows are claimed from a global
// counter, the next claim being issued before the current row is processed, so long and short
// rows (Hi-C rows range from a handful to tens of thousands of entries) balance out.
// The constants were chosen by the sweep recorded in profiles/r1_spmv_variant_sweep.txt.
#pragma once
#include <cstring>

#include "hb_internal.cuh"

// Row dot product, lane-strided: lane L takes entries s+L, s+L+32, ...  Every load instruction of the warp covers
// one contiguous run (128 B of column ids, 256 B of values) and -- the point -- so does every GATHER when the row
// has a run of consecutive columns (the near-diagonal band of a Hi-C row): 2 cache lines per gather instruction.
// The 4-entries-per-lane mapping with 128-bit loads needs fewer instructions but spreads each gather over 8 lines
// and ran the L1TEX pipe at 91 % (profiles/README.md): 2.93 ms vs 2.59 ms per launch on config 3.
// HB_SPMV_UNROLL independent entries per lane are in flight per step (3 x 16 loads); the tail runs in
// predicated groups of 4 so short rows still overlap their loads.  There is no alignment requirement,
// so the summation order does not depend on where a row starts.
template <bool WANT_MAX, typename VT>
__device__ __forceinline__ void hb_row_dot(const int32_t* __restrict__ idx, const VT* __restrict__ val,
                                           const double* __restrict__ u, int64_t s, int64_t e, int lane,
                                           double& sum_out, double& max_out) {
    constexpr int U = HB_SPMV_UNROLL;
    double acc0 = 0.0, acc1 = 0.0;
    double mx = -INFINITY;
    int64_t k = s + lane;
    for (; k + 32 * (U - 1) < e; k += 32 * U) {
        int c[U];
        double v[U], g[U];
#pragma unroll
        for (int j = 0; j < U; ++j) c[j] = hb_ldg_stream_i1(idx + k + 32 * j);
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = hb_ldg_stream_val(val + k + 32 * j);
#pragma unroll
        for (int j = 0; j < U; ++j) g[j] = __ldg(u + c[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (WANT_MAX) {
                const double p = v[j] * g[j];
                if (j & 1) acc1 += p; else acc0 += p;
                mx = fmax(mx, p);
            } else {
                if (j & 1) acc1 = fma(v[j], g[j], acc1); else acc0 = fma(v[j], g[j], acc0);
            }
        }
    }
    for (; k < e; k += 32 * 4) {  // tail (< 32*U entries left): predicated groups of 4 keep loads overlapped
        int c[4];
        double v[4], g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = (k + 32 * j < e) ? hb_ldg_stream_i1(idx + k + 32 * j) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (k + 32 * j < e) ? hb_ldg_stream_val(val + k + 32 * j) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = (k + 32 * j < e) ? __ldg(u + c[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k + 32 * j < e) {
                const double p = v[j] * g[j];
                if (j & 1) acc1 += p; else acc0 += p;
                if (WANT_MAX) mx = fmax(mx, p);
            }
        }
    }
    sum_out = hb_warp_sum(acc0 + acc1);
    if (WANT_MAX) max_out = hb_warp_max(mx);
}

// Where a row's result goes.  Single GPU / host-collective mode: world == 1 and out[0] is the local vector.
// Peer mode (hb_peer_attach): out[q] is rank q's replica of the exchange vector, mapped into this process through
// CUDA IPC; lane q of the warp stores the row's value straight into it over NVLink, so the transfer of a row
// overlaps the streaming of the next rows and no collective is launched.  When its last warp has finished, the
// kernel publishes the per-rank maximum and an arrival flag (sequence number) in every replica; consumers wait on
// the flags with hb_peer_wait_kernel.  Ordering: data stores -> __threadfence_system -> atomic ticket; the warp
// that draws the last ticket -> __threadfence_system -> flag stores (the "last block" pattern at system scope).
struct HbPeerOut {
    double* out[HB_MAX_PEERS];
    unsigned long long* flag[HB_MAX_PEERS];  // flag[q][rank] = seq   (NULL in non-peer mode)
    unsigned int* done;                      // local ticket counter
    unsigned long long seq;
    int64_t n;                               // slot n + rank of every replica receives this rank's maximum
    int world, rank, publish;
    __device__ __forceinline__ void store(int64_t g, double v, int lane) const {
        if (lane < world) out[lane][g] = v;
    }
};

__device__ __forceinline__ void hb_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long hb_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp: wait until every rank's arrival flag has reached `seq` (with a ~4 s guard so that a failed peer cannot
// hang the GPU: on timeout the status words / err flag are set and every later kernel of the loop returns at once).
static __global__ void hb_peer_wait_kernel(const unsigned long long* __restrict__ flags, unsigned long long seq, int world,
                                    int32_t* status, double* err_flag) {
    const int lane = threadIdx.x;
    if (status != nullptr && status[HB_ST_ALL_DONE] != 0) return;  // converged: the producers returned early too
    bool ok = true;
    if (lane < world) {
        const long long t0 = clock64();
        while (hb_ld_acquire_sys(flags + lane) < seq) {
            if (clock64() - t0 > 8000000000LL) {
                ok = false;
                break;
            }
            __nanosleep(200);
        }
    }
    if (!__all_sync(0xffffffffu, ok) && lane == 0) {
        if (status != nullptr) {
            status[HB_ST_OVERFLOW] = 2;  // 2 = peer exchange timed out
            status[HB_ST_ALL_DONE] = 1;
        }
        if (err_flag != nullptr) *err_flag = 1.0;
    }
}

// Epi must provide:
//   __device__ bool active(int64_t global_row) const;                                  // skip rows of finished segments
//   __device__ void row(int64_t global_row, double t, double m, int lane, peer);       // all lanes, t and m warp-uniform
//   __device__ void warp_done();                                                       // lane 0 only, once per warp
//   __device__ double local_max() const;                                               // last warp: value for slot n+rank
template <class Epi, bool WANT_MAX, typename VT>
__global__ void __launch_bounds__(HB_SPMV_THREADS, HB_SPMV_CTAS_PER_SM)
hb_spmv_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const VT* __restrict__ val,
               int64_t n_rows, int64_t row0, const double* __restrict__ u, unsigned long long* counters, int parity,
               const int32_t* __restrict__ status, Epi epi, HbPeerOut peer) {
    if (status != nullptr && status[HB_ST_ALL_DONE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (HB_SPMV_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (HB_SPMV_THREADS / 32);
    unsigned long long* counter = counters + parity;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[parity ^ 1] = 0ull;  // arm the next launch's counter

    int64_t batch = gwarp * HB_ROWS_PER_GRAB;
    while (batch < n_rows) {
        // claim the next batch now; its latency hides behind this batch's rows
        unsigned long long nxt = 0;
        if (lane == 0) nxt = atomicAdd(counter, (unsigned long long)HB_ROWS_PER_GRAB);
        int64_t bounds = 0;
        if (lane <= HB_ROWS_PER_GRAB && batch + lane <= n_rows) bounds = indptr[batch + lane];
#pragma unroll
        for (int q = 0; q < HB_ROWS_PER_GRAB; ++q) {
            const int64_t s = __shfl_sync(0xffffffffu, bounds, q);
            const int64_t e = __shfl_sync(0xffffffffu, bounds, q + 1);
            const int64_t row = batch + q;
            if (row < n_rows && epi.active(row0 + row)) {
                double t, m = 0.0;
                hb_row_dot<WANT_MAX, VT>(idx, val, u, s, e, lane, t, m);
                epi.row(row0 + row, t, m, lane, peer);
            }
        }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        batch = nwarps * HB_ROWS_PER_GRAB + (int64_t)nxt;
    }
    if (lane == 0) epi.warp_done();
    if (peer.publish) {
        int last = 0;
        if (lane == 0) {
            __threadfence_system();
            last = (atomicAdd(peer.done, 1u) == (unsigned int)(nwarps - 1)) ? 1 : 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (lane == 0) *peer.done = 0u;
            __threadfence_system();
            if (lane < peer.world) {
                if (WANT_MAX) peer.out[lane][peer.n + peer.rank] = epi.local_max();
                __threadfence_system();
                hb_st_release_sys(peer.flag[lane] + peer.rank, peer.seq);
            }
        }
    }
}

// Destination of the next SpMV on this handle: the local vector, or (peer mode) buffer seq&1 of every replica.
static inline HbPeerOut hb_next_out(hb_csr* h, double* local_out) {
    HbPeerOut po;
    memset(&po, 0, sizeof(po));
    po.n = h->n_cols;
    po.rank = h->rank;
    if (!h->peer) {
        po.world = 1;
        po.out[0] = local_out;
        po.publish = 0;
        return po;
    }
    const unsigned long long seq = ++h->peer_seq;
    const int p = (int)(seq & 1ull);
    po.world = h->world;
    po.publish = 1;
    po.seq = seq;
    po.done = h->d_done;
    for (int q = 0; q < h->world; ++q) {
        unsigned char* base = (unsigned char*)h->peer_base[q];
        po.out[q] = (double*)(base + (size_t)p * h->sym_buf_bytes);
        po.flag[q] = (unsigned long long*)(base + h->sym_flag_off + (size_t)p * 512);
    }
    return po;
}

// Launch the SpMV over the handle's block with the narrowest lossless value array it has
// (fp64 = the contract layout, 12 B/entry; uint16 / float copies made by hb_csr_pack_values: 6 / 8 B/entry).
// The values are widened to fp64 before the first multiply, so results are bit-identical across layouts.
template <class Epi, bool WANT_MAX>
static inline void hb_launch_spmv(hb_csr* h, const double* u, const int32_t* status, const Epi& epi, const HbPeerOut& po,
                                  cudaStream_t s) {
    const int p = h->spmv_parity;  // the two row counters alternate between consecutive launches on this handle
    h->spmv_parity ^= 1;
    const int grid = hb_spmv_grid(h->n_rows);
    if (h->pack_kind == 1)
        hb_spmv_kernel<Epi, WANT_MAX, uint16_t><<<grid, HB_SPMV_THREADS, 0, s>>>(
            h->d_indptr, h->d_indices, (const uint16_t*)h->d_pack, h->n_rows, h->row0, u, h->d_row_counter, p, status, epi, po);
    else if (h->pack_kind == 2)
        hb_spmv_kernel<Epi, WANT_MAX, float><<<grid, HB_SPMV_THREADS, 0, s>>>(
            h->d_indptr, h->d_indices, (const float*)h->d_pack, h->n_rows, h->row0, u, h->d_row_counter, p, status, epi, po);
    else
        hb_spmv_kernel<Epi, WANT_MAX, double><<<grid, HB_SPMV_THREADS, 0, s>>>(
            h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, u, h->d_row_counter, p, status, epi, po);
}

// out[g] = a_g * (t + eps * u_g) + b_g * c_g      (Knight-Ruiz: w = x o (A (x o p)) + v o p ; v = x o (A x))
struct HbAffineEpi {
    const double* __restrict__ u;
    const double* __restrict__ a;
    const double* __restrict__ b;
    const double* __restrict__ c;
    double eps;
    __device__ __forceinline__ bool active(int64_t) const { return true; }
    __device__ __forceinline__ void row(int64_t g, double t, double, int lane, const HbPeerOut& peer) const {
        double v = t + eps * u[g];
        if (a != nullptr) v *= a[g];
        if (b != nullptr && c != nullptr) v += b[g] * c[g];
        peer.store(g, v, lane);
    }
    __device__ __forceinline__ void warp_done() const {}
    __device__ __forceinline__ double local_max() const { return 0.0; }
};
