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
#include <cstdlib>
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
// COHERENT selects how the gathered vector is read: false = non-coherent read-only path (__ldg; the vector is constant
// for the whole launch), true = ordinary L1-cached loads that the acquire side of a grid barrier invalidates (the
// persistent solver kernels rewrite the vector between iterations of one launch).
#ifndef HB_LOOP_GATHER
#define HB_LOOP_GATHER 0   // 0: weak ld.global (shipped) ; tuning experiments only: 1 = non-coherent, 2 = weak + L1::evict_last
#endif
template <bool COHERENT>
__device__ __forceinline__ double hb_ld_gather(const double* p) {
    if (COHERENT && HB_LOOP_GATHER != 1) {
        double r;
#if HB_LOOP_GATHER == 2
        asm volatile("ld.global.L1::evict_last.f64 %0, [%1];" : "=d"(r) : "l"(p));
#else
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p));   // weak, default caching: LDG.E.64
#endif
        return r;
    }
    return __ldg(p);
}

template <bool WANT_MAX, typename VT, bool COHERENT = false>
__device__ __forceinline__ void hb_row_dot(const int32_t* __restrict__ idx, const VT* __restrict__ val,
                                           const double* u, int64_t s, int64_t e, int lane,
                                           double& sum_out, double& max_out) {
    double acc0 = 0.0, acc1 = 0.0;
    double mx = -INFINITY;
    int64_t k = s + lane;
    constexpr int U = HB_SPMV_UNROLL;
    for (; k + 32 * (U - 1) < e; k += 32 * U) {
        int c[U];
        double v[U], g[U];
#pragma unroll
        for (int j = 0; j < U; ++j) c[j] = hb_ldg_stream_i1(idx + k + 32 * j);
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = hb_ldg_stream_val(val + k + 32 * j);
#pragma unroll
        for (int j = 0; j < U; ++j) g[j] = hb_ld_gather<COHERENT>(u + c[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (WANT_MAX) {
                const double p = __dmul_rn(v[j], g[j]);   // the product is also the value under the 1e100 guard: no contraction
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
        for (int j = 0; j < 4; ++j) g[j] = (k + 32 * j < e) ? hb_ld_gather<COHERENT>(u + c[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k + 32 * j < e) {
                if (WANT_MAX) {
                    const double p = __dmul_rn(v[j], g[j]);
                    if (j & 1) acc1 += p; else acc0 += p;
                    mx = fmax(mx, p);
                } else {
                    if (j & 1) acc1 = fma(v[j], g[j], acc1); else acc0 = fma(v[j], g[j], acc0);
                }
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
//   struct Pre; __device__ Pre pre(int64_t global_row) const;                          // per-row operands, loaded early
//   __device__ void row(int64_t global_row, double t, double m, int lane, peer, pre);  // all lanes, t and m warp-uniform
//   __device__ void warp_done();                                                       // lane 0 only, once per warp
//   __device__ double local_max() const;                                               // last warp: value for slot n+rank
//
// hb_spmv_rows: the row loop shared by the stand-alone kernel below and by the persistent solver kernels
// (hb_ice_loop_kernel, hb_kr_solve_kernel).  `counter` must be 0 on entry; whoever calls it re-arms the OTHER counter.
template <class Epi, bool WANT_MAX, typename VT, bool COHERENT = false, class Out = HbPeerOut>
__device__ __forceinline__ void hb_spmv_rows(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                             const VT* __restrict__ val, int64_t n_rows, int64_t row0,
                                             const double* u, unsigned long long* counter, Epi& epi,
                                             const Out& peer) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (HB_SPMV_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (HB_SPMV_THREADS / 32);
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
                // the epilogue's own operands (r_i; x_i, u_i, v_i, p_i) are requested BEFORE the row is streamed: at the
                // end of the row their latency would sit, un-overlapped, in front of the next row's first loads
                const typename Epi::Pre pre = epi.pre(row0 + row);
                double t, m = 0.0;
                hb_row_dot<WANT_MAX, VT, COHERENT>(idx, val, u, s, e, lane, t, m);
                epi.row(row0 + row, t, m, lane, peer, pre);
            }
        }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        batch = nwarps * HB_ROWS_PER_GRAB + (int64_t)nxt;
    }
    if (lane == 0) epi.warp_done();
}

template <class Epi, bool WANT_MAX, typename VT>
__global__ void __launch_bounds__(HB_SPMV_THREADS, HB_SPMV_CTAS_PER_SM)
hb_spmv_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const VT* __restrict__ val,
               int64_t n_rows, int64_t row0, const double* __restrict__ u, unsigned long long* counters, int parity,
               const int32_t* __restrict__ status, Epi epi, HbPeerOut peer) {
    if (status != nullptr && status[HB_ST_ALL_DONE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (HB_SPMV_THREADS / 32);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[parity ^ 1] = 0ull;  // arm the next launch's counter
    hb_spmv_rows<Epi, WANT_MAX, VT>(indptr, idx, val, n_rows, row0, u, counters + parity, epi, peer);
    if (peer.publish) {
        // Release chain (PTX memory model, system scope): every lane that stored into a peer replica fences ITS OWN
        // stores, the warp barrier orders those fences before lane 0's ticket, the ticket chain (gpu-scope atomics
        // bracketed by fences) reaches the last warp's lane 0, whose fence + warp barrier precede the st.release.sys
        // of every flag.  A consumer that ld.acquire.sys-es a flag therefore observes every row of this rank.
        if (lane < peer.world) __threadfence_system();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence();
            last = (atomicAdd(peer.done, 1u) == (unsigned int)(nwarps - 1)) ? 1 : 0;
            __threadfence_system();
        }
        last = __shfl_sync(0xffffffffu, last, 0);   // also a warp barrier: lanes 1.. are ordered after lane 0's fence
        if (last) {
            if (lane == 0) *peer.done = 0u;
            __syncwarp();
            if (lane < peer.world) {
                if (WANT_MAX) peer.out[lane][peer.n + peer.rank] = epi.local_max();
                hb_st_release_sys(peer.flag[lane] + peer.rank, peer.seq);   // release = fence.sys + store
            }
        }
    }
}

// ---- persistent solver kernels (hb_ice_loop_kernel, hb_kr_solve_kernel): exchange state for a whole launch ----------
struct HbLoopPeer {
    double* buf[2][HB_MAX_PEERS];               // exchange buffer p of every rank's replica (own entry = local memory)
    unsigned long long* flag[2][HB_MAX_PEERS];  // arrival flags p of every replica; slot [rank] is written by `rank`
    unsigned long long seq0;                    // exchanges performed on this handle before this launch
    int world, rank;
};

// Destination of a row inside a persistent kernel: the local vector (one GPU) or buffer pb of every rank's replica,
// read straight from the kernel parameters (`peer` points into the __grid_constant__ argument block).
struct HbLoopOut {
    const HbLoopPeer* peer;
    double* local;
    int pb, world;
    __device__ __forceinline__ void store(int64_t g, double v, int lane) const {
        if (world == 1) {
            if (lane == 0) local[g] = v;
        } else if (lane < world) {
            peer->buf[pb][lane][g] = v;
        }
    }
};

// every rank's arrival flag of exchange `seq` (one warp per CTA polls; ~4 s guard so a dead peer cannot hang the GPU)
__device__ __forceinline__ void hb_loop_wait_flags(const unsigned long long* flags, unsigned long long seq, int world,
                                                   double* err) {
    if ((threadIdx.x >> 5) == 0) {
        const int lane = threadIdx.x & 31;
        if (lane < world) {
            const long long t0 = clock64();
            while (hb_ld_acquire_sys(flags + lane) < seq) {
                if (clock64() - t0 > 8000000000LL) {
                    *err = 1.0;
                    break;
                }
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
}


// Exchange `seq` of a persistent kernel, after its rows were stored and the grid has synchronised: block 0 releases
// this rank's arrival flag in every replica (and ships `slot_value` into slot n + rank), then every CTA waits for all
// ranks' flags in the local replica.  Each storing lane has fenced its own peer stores before the grid barrier.
__device__ __forceinline__ void hb_loop_publish_and_wait(const HbLoopPeer& peer, int pb, unsigned long long seq,
                                                         int64_t n_cols, double slot_value, double* err) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && warp == 0 && lane < peer.world) {
        peer.buf[pb][lane][n_cols + peer.rank] = slot_value;
        hb_st_release_sys(peer.flag[pb][lane] + peer.rank, seq);
    }
    hb_loop_wait_flags(peer.flag[pb][peer.rank], seq, peer.world, err);
}

// Fill the launch-constant exchange description of a handle in peer mode (world = 1 otherwise).
static inline void hb_fill_loop_peer(const hb_csr* h, HbLoopPeer* lp) {
    memset(lp, 0, sizeof(*lp));
    lp->world = 1;
    lp->rank = h->rank;
    lp->seq0 = h->pc != nullptr ? h->pc->seq : 0ull;
    if (!h->peer) return;
    lp->world = h->world;
    for (int p = 0; p < 2; ++p)
        for (int q = 0; q < h->world; ++q) {
            unsigned char* base = (unsigned char*)h->pc->base[q];
            lp->buf[p][q] = (double*)(base + (size_t)p * h->pc->buf_bytes);
            lp->flag[p][q] = (unsigned long long*)(base + h->pc->flag_off + (size_t)p * 512);
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
    hb_loop_settle(h);   // a persistent kernel still in flight owns part of the sequence: wait for its count
    const unsigned long long seq = ++h->pc->seq;
    const int p = (int)(seq & 1ull);
    po.world = h->world;
    po.publish = 1;
    po.seq = seq;
    po.done = h->pc->d_done;
    for (int q = 0; q < h->world; ++q) {
        unsigned char* base = (unsigned char*)h->pc->base[q];
        po.out[q] = (double*)(base + (size_t)p * h->pc->buf_bytes);
        po.flag[q] = (unsigned long long*)(base + h->pc->flag_off + (size_t)p * 512);
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
    struct Pre {
        double u, a, b, c;
    };
    __device__ __forceinline__ bool active(int64_t) const { return true; }
    __device__ __forceinline__ Pre pre(int64_t g) const {
        Pre q;
        q.u = u[g];
        q.a = a != nullptr ? a[g] : 1.0;
        q.b = (b != nullptr && c != nullptr) ? b[g] : 0.0;
        q.c = (b != nullptr && c != nullptr) ? c[g] : 0.0;
        return q;
    }
    template <class Out>
    __device__ __forceinline__ void row(int64_t g, double t, double, int lane, const Out& peer, const Pre& q) const {
        double v = fma(eps, q.u, t);
        if (a != nullptr) v *= q.a;
        if (b != nullptr && c != nullptr) v = fma(q.b, q.c, v);
        peer.store(g, v, lane);
    }
    __device__ __forceinline__ void warp_done() const {}
    __device__ __forceinline__ double local_max() const { return 0.0; }
};
