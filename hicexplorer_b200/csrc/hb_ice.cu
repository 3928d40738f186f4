// ICE iterative correction on the device CSR store.
// Replaces hicexplorer/iterativeCorrection.py:10-86 of the reference (see include/hicbalance.h).
//
// Formulation.  The reference rescales W in place every iteration (lines 56-57):
//   W_k = A o (r_k r_k^T),  r_k = prod_{m<=k} 1/s_m .
// Its next marginals (line 42) are therefore s_{k+1} = r_k o (A r_k): ONE SpMV with the original
// matrix per iteration, 12 B per stored entry, nothing written back.  The rescaled values exist
// only in registers, where their maximum is taken for the 1e100 guard (line 58).  The matrix is
// materialised once, after convergence (hb_dev_ice_emit; lines 79-80).
//
// Per iteration: hb_spmv_kernel<IceEpi>  ->  hb_ice_reduce_kernel  ->  hb_ice_update_kernel.
// Between the first and the second, a sharded run sum-allreduces the exchange vector (each rank
// has written only its own rows, the rest is zero); the two O(n) kernels then run identically on
// every rank.  Convergence lives on the device (per-segment flags), the host only polls.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "hb_spmv.cuh"

namespace cg = cooperative_groups;

struct HbIceEpi {
    const double* __restrict__ r;
    unsigned long long* wmax_slot;  // bits of a non-negative double (slot n + rank of the local exchange vector)
    const int64_t* __restrict__ seg_off;
    const int32_t* __restrict__ seg_done;
    int nseg;
    double wmax;  // per-thread running maximum (lane 0)

    __device__ __forceinline__ bool active(int64_t g) const {
        if (nseg <= 1) return true;
        int lo = 0, hi = nseg;  // largest s with seg_off[s] <= g
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= g) lo = mid; else hi = mid;
        }
        return seg_done[lo] == 0;
    }
    struct Pre {
        double r;
    };
    __device__ __forceinline__ Pre pre(int64_t g) const {
        Pre q;
        q.r = r[g];
        return q;
    }
    template <class Out>
    __device__ __forceinline__ void row(int64_t g, double t, double m, int lane, const Out& peer, const Pre& q) {
        peer.store(g, q.r * t, lane);
        wmax = fmax(wmax, q.r * m);  // NaN-ignoring, like np.any(W.data > 1e100)
    }
    __device__ __forceinline__ double local_max() const {
        return __longlong_as_double((long long)atomicAdd(wmax_slot, 0ull));
    }
    __device__ __forceinline__ void warp_done() const {
        if (wmax > 0.0) atomicMax(wmax_slot, (unsigned long long)__double_as_longlong(wmax));
    }
};

// Block-wide fixed-order sum of two values; result valid in thread 0.
__device__ __forceinline__ void hb_block_sum2(double& a, double& b, double* smem /* 2*32 */) {
    a = hb_warp_sum(a);
    b = hb_warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) {
        smem[w] = a;
        smem[32 + w] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < nw; ++i) {
            sa += smem[i];
            sb += smem[32 + i];
        }
        a = sa;
        b = sb;
    }
}

// tile_sum/tile_cnt = sum / count of the non-zero entries of vec in each fixed tile; the last
// block to finish combines the tiles of every segment (in tile order) into seg_out = sum / count.
// Used for the marginal mean (iterativeCorrection.py:43-44) and the final bias mean (line 77).
__global__ void __launch_bounds__(256)
hb_ice_reduce_kernel(const double* __restrict__ vec, const int64_t* __restrict__ tile_row0,
                     const int32_t* __restrict__ tile_len, const int32_t* __restrict__ tile_seg,
                     const int32_t* __restrict__ seg_tile0, int nseg, const int32_t* __restrict__ seg_done,
                     double* __restrict__ tile_sum, double* __restrict__ tile_cnt, double* __restrict__ seg_out,
                     unsigned int* counter, int32_t* status, const double* __restrict__ wmax_slots, int world) {
    __shared__ double smem[64];
    __shared__ int s_last;
    if (status != nullptr && status[HB_ST_ALL_DONE] != 0) return;
    const int t = blockIdx.x;
    const int seg = tile_seg[t];
    double s = 0.0, c = 0.0;
    if (seg_done == nullptr || seg_done[seg] == 0) {
        const int64_t r0 = tile_row0[t];
        const int len = tile_len[t];
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const double v = vec[r0 + i];
            if (v != 0.0) {
                s += v;
                c += 1.0;
            }
        }
    }
    hb_block_sum2(s, c, smem);
    if (threadIdx.x == 0) {
        tile_sum[t] = s;
        tile_cnt[t] = c;
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int sg = 0; sg < nseg; ++sg) {
        double a = 0.0, b = 0.0;
        for (int i = seg_tile0[sg] + threadIdx.x; i < seg_tile0[sg + 1]; i += blockDim.x) {
            a += __ldcg(tile_sum + i);
            b += __ldcg(tile_cnt + i);
        }
        hb_block_sum2(a, b, smem);
        if (threadIdx.x == 0) seg_out[sg] = a / b;
    }
    if (threadIdx.x == 0) {
        *counter = 0u;
        if (wmax_slots != nullptr && status != nullptr) {
            double wm = 0.0;
            for (int w = 0; w < world; ++w) wm = fmax(wm, wmax_slots[w]);
            if (wm > 1e100) {  // iterativeCorrection.py:58
                status[HB_ST_OVERFLOW] = 1;
                status[HB_ST_ALL_DONE] = 1;
            }
        }
    }
}

// s = s_raw / mean ; bias *= s ; deviation = max|s - 1| ; r *= 1/s   (iterativeCorrection.py:44-57)
__global__ void __launch_bounds__(256)
hb_ice_update_kernel(const double* __restrict__ exch, const int64_t* __restrict__ tile_row0,
                     const int32_t* __restrict__ tile_len, const int32_t* __restrict__ tile_seg, int nseg,
                     const double* __restrict__ seg_mean, double* __restrict__ bias, double* __restrict__ r,
                     unsigned long long* seg_devbits, double* seg_dev, int32_t* seg_done, int32_t* seg_iters,
                     unsigned int* counter, int32_t* status, double tol) {
    __shared__ unsigned long long smax[32];
    __shared__ int s_last;
    if (status[HB_ST_ALL_DONE] != 0) return;
    const int t = blockIdx.x;
    const int seg = tile_seg[t];
    if (seg_done[seg] == 0) {
        const double mean = seg_mean[seg];
        const int64_t r0 = tile_row0[t];
        const int len = tile_len[t];
        unsigned long long dmax = 0ull;
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const int64_t g = r0 + i;
            const double sraw = exch[g];
            const double s = sraw / mean;
            bias[g] *= s;
            const unsigned long long d = (unsigned long long)__double_as_longlong(fabs(s - 1.0));
            dmax = d > dmax ? d : dmax;  // NaN bits sort above +inf: a NaN deviation never converges, like numpy
            r[g] = (sraw == 0.0) ? 0.0 : r[g] * (1.0 / s);
        }
        dmax = hb_warp_max_u64(dmax);
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = dmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < (int)(blockDim.x >> 5); ++i) dmax = smax[i] > dmax ? smax[i] : dmax;
            atomicMax(seg_devbits + seg, dmax);
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ int s_active;
    if (threadIdx.x == 0) s_active = 0;
    __syncthreads();
    int my_active = 0;
    for (int sg = threadIdx.x; sg < nseg; sg += blockDim.x) {
        if (seg_done[sg] == 0) {
            const double dev = __longlong_as_double((long long)__ldcg(seg_devbits + sg));
            seg_dev[sg] = dev;
            seg_devbits[sg] = 0ull;
            seg_iters[sg] += 1;
            if (dev < tol) seg_done[sg] = 1;  // strict '<', tested after the rescale (line 72)
            else my_active += 1;
        }
    }
    if (my_active) atomicAdd(&s_active, my_active);
    __syncthreads();
    if (threadIdx.x == 0) {
        *counter = 0u;
        status[HB_ST_ITER] += 1;
        status[HB_ST_ACTIVE] = s_active;
        if (s_active == 0) status[HB_ST_ALL_DONE] = 1;
    }
}

__global__ void hb_ice_begin_kernel(double* bias, double* r, int64_t n, unsigned long long* seg_devbits, double* seg_dev,
                                    int32_t* seg_done, int32_t* seg_iters, double* seg_corr, int nseg, int32_t* status,
                                    unsigned int* counters, unsigned long long* row_counters) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < n; i += stride) {
        bias[i] = 1.0;
        r[i] = 1.0;
    }
    for (int64_t i = i0; i < nseg; i += stride) {
        seg_devbits[i] = 0ull;
        seg_dev[i] = __longlong_as_double(0x7ff8000000000000LL);
        seg_done[i] = 0;
        seg_iters[i] = 0;
        seg_corr[i] = 1.0;
    }
    if (i0 == 0) {
        for (int k = 0; k < HB_ST_WORDS; ++k) status[k] = 0;
        status[HB_ST_ACTIVE] = nseg;
        counters[0] = counters[1] = 0u;
        row_counters[0] = row_counters[1] = 0ull;
    }
}

// bias /= corr  (iterativeCorrection.py:78)
__global__ void __launch_bounds__(256)
hb_ice_scale_bias_kernel(double* __restrict__ bias, const int64_t* __restrict__ tile_row0,
                         const int32_t* __restrict__ tile_len, const int32_t* __restrict__ tile_seg,
                         const double* __restrict__ seg_corr) {
    const int t = blockIdx.x;
    const double corr = seg_corr[tile_seg[t]];
    const int64_t r0 = tile_row0[t];
    for (int i = threadIdx.x; i < tile_len[t]; i += blockDim.x) bias[r0 + i] = bias[r0 + i] / corr;
}

// W_ij = ((A_ij * r_i) * r_j) * corr * corr  (lines 56-57, 79); running maxima of the value before
// and after the corr^2 factor for the two overflow guards (lines 58, 80).
__global__ void __launch_bounds__(256)
hb_ice_emit_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                   int64_t n_rows, int64_t row0, const double* __restrict__ r, const int64_t* __restrict__ seg_off,
                   const double* __restrict__ seg_corr, int nseg, int64_t phase, double* __restrict__ out,
                   unsigned long long* max_bits /* [0] before corr, [1] after */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double m0 = 0.0, m1 = 0.0;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int lo = 0, hi = nseg;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= g) lo = mid; else hi = mid;
        }
        const double corr = seg_corr[lo];
        const double ri = r[g];
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t k = s + lane; k < e; k += 32) {
            const double w = (hb_ldg_stream_d1(val + k) * ri) * __ldg(r + hb_ldg_stream_i1(idx + k));
            const double f = w * corr * corr;
            out[k - phase] = f;
            m0 = fmax(m0, w);
            m1 = fmax(m1, f);
        }
    }
    m0 = hb_warp_max(m0);
    m1 = hb_warp_max(m1);
    if (lane == 0) {
        if (m0 > 0.0) atomicMax(max_bits + 0, (unsigned long long)__double_as_longlong(m0));
        if (m1 > 0.0) atomicMax(max_bits + 1, (unsigned long long)__double_as_longlong(m1));
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The whole ICE loop in ONE cooperative launch (hb_ice_loop_kernel).  A persistent grid of 148 x 3 CTAs runs, per
// iteration:   A  the SpMV over the block's rows (same row loop as hb_spmv_kernel; in peer mode every row goes
//                 straight into every rank's replica over NVLink)
//              -- grid barrier; peer mode: publish this rank's arrival flag, wait for everybody's
//              B  sum / count of the non-zero marginals per fixed tile            (iterativeCorrection.py:43-44)
//              -- grid barrier
//              C  mean per segment (every CTA folds the tile partials in tile order -> bit-identical everywhere),
//                 s = s_raw / mean ; bias *= s ; deviation = max|s - 1| ; r *= 1/s (:44-57); 1e100 guard (:58)
//              -- grid barrier
//              D  every CTA evaluates the strict '<' convergence test (:72) per segment from the same maxima
// so an iteration costs three grid barriers instead of three kernel launches, a memset, a status copy and an event,
// and the host never synchronises inside the loop (progress is mirrored into pinned memory for whoever wants to
// look).  Arithmetic and summation order are those of the step-level kernels above: results are bit-identical.
// Cross-CTA data is read with ld.cg; the gathered vector r is read with plain (coherent, L1-cached) loads -- the
// acquire side of a grid barrier invalidates L1 -- never through the non-coherent path, since r changes every
// iteration inside this kernel.
struct HbIceLoopArgs {
    const int64_t* indptr;
    const int32_t* idx;
    const void* val;
    int64_t n_rows, row0, n_cols;
    const int64_t* tile_row0;
    const int32_t* tile_len;
    const int32_t* tile_seg;
    const int32_t* seg_tile0;
    const int64_t* seg_off;
    int n_tiles, nseg;
    double* bias;
    double* r;
    double* exch;  // single GPU: the marginals (n_cols + 64)
    double* tile_sum;
    double* tile_cnt;
    unsigned long long* devbits;  // 2 x nseg, zero on entry
    unsigned long long* wmax;     // 2 words, zero on entry
    double* seg_dev;
    int32_t* seg_done;
    int32_t* seg_iters;
    int32_t* status;
    int32_t* h_loop;  // pinned, device alias
    unsigned long long* row_counters;
    double* err;
    int iters, row_parity;
    double tol;
    unsigned long long* timing;   // debug (HB_LOOP_TIMING=1): block 0 stamps %globaltimer at the phase boundaries
    HbLoopPeer peer;
};

__device__ __forceinline__ unsigned long long hb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define HB_STAMP(slot)                                                                              \
    do {                                                                                            \
        if (a.timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && it < 32) a.timing[it * 8 + (slot)] = hb_globaltimer(); \
    } while (0)

struct HbIceLoopEpi {
    const double* r;  // written by this kernel between iterations: no __restrict__, no non-coherent loads
    unsigned long long* wmax_slot;
    const int64_t* __restrict__ seg_off;
    const int32_t* seg_done;
    int nseg;
    double wmax;
    __device__ __forceinline__ bool active(int64_t g) const {
        if (nseg <= 1) return true;
        int lo = 0, hi = nseg;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= g) lo = mid; else hi = mid;
        }
        return __ldcg(seg_done + lo) == 0;
    }
    struct Pre {
        double r;
    };
    __device__ __forceinline__ Pre pre(int64_t g) const {
        Pre q;
        q.r = hb_ld_gather<true>(r + g);   // coherent, L1-cached: the row gathers its own diagonal neighbourhood anyway
        return q;
    }
    template <class Out>
    __device__ __forceinline__ void row(int64_t g, double t, double m, int lane, const Out& peer, const Pre& q) {
        peer.store(g, q.r * t, lane);
        wmax = fmax(wmax, q.r * m);
    }
    __device__ __forceinline__ double local_max() const { return 0.0; }
    __device__ __forceinline__ void warp_done() const {
        if (wmax > 0.0) atomicMax(wmax_slot, (unsigned long long)__double_as_longlong(wmax));
    }
};

template <typename VT>
__global__ void __launch_bounds__(HB_SPMV_THREADS, HB_SPMV_CTAS_PER_SM)
hb_ice_loop_kernel(const __grid_constant__ HbIceLoopArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double smem[64];
    __shared__ unsigned long long smax[HB_SPMV_THREADS / 32];
    __shared__ double s_mean;
    __shared__ int s_active;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool peer = a.peer.world > 1;
    const int world = peer ? a.peer.world : 1;
    // status is stable here: it was last written before the previous launch on this stream ended
    if (a.status[HB_ST_ALL_DONE] != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.h_loop[HB_LP_EXCHANGES] = 0;
        return;
    }
    const int it_base = a.status[HB_ST_ITER];
    int executed = 0;
    for (int it = 0; it < a.iters; ++it) {
        const int par = (a.row_parity + it) & 1;
        const int cur = it & 1;
        const unsigned long long seq = a.peer.seq0 + 1ull + (unsigned long long)it;
        const int pb = (int)(seq & 1ull);
        double* exch = peer ? a.peer.buf[pb][a.peer.rank] : a.exch;
        // ---- A: marginals of the block's rows
        HB_STAMP(0);
        if (blockIdx.x == 0 && threadIdx.x == 0) a.row_counters[par ^ 1] = 0ull;
        {
            HbIceLoopEpi epi;
            epi.r = a.r;
            epi.wmax_slot = a.wmax + cur;
            epi.seg_off = a.seg_off;
            epi.seg_done = a.seg_done;
            epi.nseg = a.nseg;
            epi.wmax = 0.0;
            HbLoopOut po;
            po.peer = &a.peer;
            po.local = exch;
            po.pb = pb;
            po.world = world;
            hb_spmv_rows<HbIceLoopEpi, true, VT, true, HbLoopOut>(a.indptr, a.idx, (const VT*)a.val, a.n_rows, a.row0, a.r,
                                                                  a.row_counters + par, epi, po);
        }
        if (peer && lane < world) __threadfence_system();   // each storing lane releases its own peer stores
        HB_STAMP(1);
        grid.sync();
        HB_STAMP(2);
        if (peer)
            hb_loop_publish_and_wait(a.peer, pb, seq, a.n_cols, __longlong_as_double((long long)__ldcg(a.wmax + cur)), a.err);
        // ---- B: per-tile sum / count of the non-zero marginals
        HB_STAMP(3);
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            double s = 0.0, c = 0.0;
            if (__ldcg(a.seg_done + a.tile_seg[t]) == 0) {
                const int64_t r0 = a.tile_row0[t];
                const int len = a.tile_len[t];
                for (int i = threadIdx.x; i < len; i += blockDim.x) {
                    const double v = __ldcg(exch + r0 + i);
                    if (v != 0.0) {
                        s += v;
                        c += 1.0;
                    }
                }
            }
            hb_block_sum2(s, c, smem);
            if (threadIdx.x == 0) {
                a.tile_sum[t] = s;
                a.tile_cnt[t] = c;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.wmax[cur ^ 1] = 0ull;   // next iteration's accumulator
        HB_STAMP(4);
        grid.sync();
        HB_STAMP(5);
        if (peer && __ldcg(a.err) != 0.0) {   // consistent: every writer of err arrived at the barrier before
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                a.status[HB_ST_OVERFLOW] = 2;
                a.status[HB_ST_ALL_DONE] = 1;
                a.h_loop[HB_ST_OVERFLOW] = 2;
                a.h_loop[HB_ST_ALL_DONE] = 1;
            }
            executed = it + 1;
            break;
        }
        // ---- C: normalise, accumulate the bias, deviation, next scaling vector
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double wm = 0.0;
            if (peer) {
                for (int q = 0; q < world; ++q) wm = fmax(wm, __ldcg(exch + a.n_cols + q));
            } else {
                wm = __longlong_as_double((long long)__ldcg(a.wmax + cur));
            }
            if (wm > 1e100) a.status[HB_ST_OVERFLOW] = 1;   // iterativeCorrection.py:58
        }
        int cached_seg = -1;
        double mean = 0.0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            const int seg = a.tile_seg[t];
            if (__ldcg(a.seg_done + seg) != 0) continue;
            if (seg != cached_seg) {
                double sa = 0.0, sb = 0.0;
                for (int i = a.seg_tile0[seg] + threadIdx.x; i < a.seg_tile0[seg + 1]; i += blockDim.x) {
                    sa += __ldcg(a.tile_sum + i);
                    sb += __ldcg(a.tile_cnt + i);
                }
                hb_block_sum2(sa, sb, smem);
                if (threadIdx.x == 0) s_mean = sa / sb;
                __syncthreads();
                mean = s_mean;
                cached_seg = seg;
            }
            const int64_t r0 = a.tile_row0[t];
            const int len = a.tile_len[t];
            unsigned long long dmax = 0ull;
            for (int i = threadIdx.x; i < len; i += blockDim.x) {
                const int64_t g = r0 + i;
                const double sraw = __ldcg(exch + g);
                const double sv = sraw / mean;
                a.bias[g] *= sv;
                const unsigned long long d = (unsigned long long)__double_as_longlong(fabs(sv - 1.0));
                dmax = d > dmax ? d : dmax;
                a.r[g] = (sraw == 0.0) ? 0.0 : a.r[g] * (1.0 / sv);
            }
            dmax = hb_warp_max_u64(dmax);
            __syncthreads();
            if (lane == 0) smax[warp] = dmax;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int i = 1; i < (int)(blockDim.x >> 5); ++i) dmax = smax[i] > dmax ? smax[i] : dmax;
                atomicMax(a.devbits + (size_t)cur * a.nseg + seg, dmax);
            }
        }
        HB_STAMP(6);
        grid.sync();
        HB_STAMP(7);
        // ---- D: convergence per segment; every CTA computes the same answers from the same maxima
        if (threadIdx.x == 0) s_active = 0;
        __syncthreads();
        int my_active = 0;
        for (int sg = threadIdx.x; sg < a.nseg; sg += blockDim.x) {
            if (__ldcg(a.seg_done + sg) == 0) {
                const double dev = __longlong_as_double((long long)__ldcg(a.devbits + (size_t)cur * a.nseg + sg));
                if (dev < a.tol) {   // strict '<', tested after the rescale (iterativeCorrection.py:72)
                    // idempotent writes: every CTA stores the same values
                    a.seg_dev[sg] = dev;
                    a.seg_iters[sg] = it_base + it + 1;
                    a.seg_done[sg] = 1;
                } else {
                    a.seg_dev[sg] = dev;
                    a.seg_iters[sg] = it_base + it + 1;
                    my_active += 1;
                }
            }   // else: finished earlier, or another CTA has just recorded its convergence -- not active either way
            a.devbits[(size_t)(cur ^ 1) * a.nseg + sg] = 0ull;
        }
        if (my_active) atomicAdd(&s_active, my_active);
        __syncthreads();
        const int active = s_active;
        const int overflow = __ldcg(a.status + HB_ST_OVERFLOW);
        executed = it + 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.status[HB_ST_ITER] = it_base + it + 1;
            a.status[HB_ST_ACTIVE] = active;
            a.h_loop[HB_ST_ITER] = it_base + it + 1;
            a.h_loop[HB_ST_ACTIVE] = active;
            if (active == 0 || overflow != 0) {
                a.status[HB_ST_ALL_DONE] = 1;
                a.h_loop[HB_ST_OVERFLOW] = overflow;
                a.h_loop[HB_ST_ALL_DONE] = 1;
            }
        }
        if (active == 0 || overflow != 0) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.row_counters[0] = a.row_counters[1] = 0ull;   // both armed for whichever launch comes next
        a.h_loop[HB_LP_EXCHANGES] = executed;
        __threadfence_system();
    }
}

static cudaStream_t pick_stream(hb_csr* h, void* stream) { return stream ? (cudaStream_t)stream : h->stream; }

extern "C" {

int hb_dev_ice_begin(hb_csr* h, void* stream) {
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    rc = hb_loop_settle(h);
    if (rc) return rc;
    for (int i = 0; i < HB_LOOP_WORDS; ++i) ((volatile int32_t*)h->h_loop)[i] = 0;
    cudaStream_t s = pick_stream(h, stream);
    int64_t n = h->n_cols;
    int grid = (int)((n + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > HB_NUM_SMS * 8) grid = HB_NUM_SMS * 8;
    hb_ice_begin_kernel<<<grid, 256, 0, s>>>(h->d_bias, h->d_r, n, h->d_seg_devbits, h->d_seg_dev, h->d_seg_done,
                                             h->d_seg_iters, h->d_seg_corr, h->nseg, h->d_status, h->d_counters,
                                             h->d_row_counter);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

int hb_dev_ice_spmv(hb_csr* h, double* d_exch, int rank, int world, void* stream) {
    HB_REQUIRE(h != nullptr && h->state_ready, "call hb_dev_ice_begin first");
    HB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank/world");
    cudaStream_t s = pick_stream(h, stream);
    HbPeerOut po;
    if (h->peer) {
        // peer mode: rows go straight into every rank's replica (buffer seq & 1); only the local maximum slot is reset
        HB_REQUIRE(world == h->world && rank == h->rank, "rank/world differ from hb_peer_export");
        po = hb_next_out(h, nullptr);
        d_exch = hb_sym_buf(h, (int)(po.seq & 1ull));
        HB_CUDA(cudaMemsetAsync(d_exch + h->n_cols + rank, 0, sizeof(double), s));
    } else {
        if (d_exch == nullptr) d_exch = h->d_exch;
        if (world > 1)
            HB_CUDA(cudaMemsetAsync(d_exch, 0, sizeof(double) * (size_t)(h->n_cols + world), s));
        else
            HB_CUDA(cudaMemsetAsync(d_exch + h->n_cols, 0, sizeof(double), s));
        if (h->n_rows == 0) return HB_OK;
        po = hb_next_out(h, d_exch);
    }
    HbIceEpi epi;
    epi.r = h->d_r;
    epi.wmax_slot = reinterpret_cast<unsigned long long*>(d_exch + h->n_cols + rank);
    epi.seg_off = h->d_seg_off;
    epi.seg_done = h->d_seg_done;
    epi.nseg = h->nseg;
    epi.wmax = 0.0;
    hb_launch_spmv<HbIceEpi, true>(h, h->d_r, h->d_status, epi, po, s);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

int hb_dev_ice_update(hb_csr* h, const double* d_exch, int world, double tol, void* stream) {
    HB_REQUIRE(h != nullptr && h->state_ready, "call hb_dev_ice_begin first");
    cudaStream_t s = pick_stream(h, stream);
    if (h->peer) {
        // wait for every rank's arrival flag of the exchange just issued, then read the local replica
        const int p = (int)(h->pc->seq & 1ull);
        d_exch = hb_sym_buf(h, p);
        hb_peer_wait_kernel<<<1, 32, 0, s>>>(hb_sym_flags(h, p), h->pc->seq, h->world, h->d_status, nullptr);
        HB_LAUNCH_CHECK();
    } else if (d_exch == nullptr) {
        d_exch = h->d_exch;
    }
    if (h->n_tiles == 0) return HB_OK;
    hb_ice_reduce_kernel<<<h->n_tiles, 256, 0, s>>>(d_exch, h->d_tile_row0, h->d_tile_len, h->d_tile_seg,
                                                     h->d_seg_tile0, h->nseg, h->d_seg_done, h->d_tile_sum,
                                                     h->d_tile_cnt, h->d_seg_mean, h->d_counters + 0, h->d_status,
                                                     d_exch + h->n_cols, world);
    HB_LAUNCH_CHECK();
    hb_ice_update_kernel<<<h->n_tiles, 256, 0, s>>>(d_exch, h->d_tile_row0, h->d_tile_len, h->d_tile_seg, h->nseg,
                                                     h->d_seg_mean, h->d_bias, h->d_r, h->d_seg_devbits, h->d_seg_dev,
                                                     h->d_seg_done, h->d_seg_iters, h->d_counters + 1, h->d_status, tol);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// Enqueue ONE cooperative launch that runs up to `iters` ICE iterations (fewer if every segment converges, a value
// exceeds 1e100 or a peer times out).  Asynchronous; the state continues from wherever hb_dev_ice_begin / earlier
// launches left it.  Needs the whole exchange on the device: one GPU, or peer mode.
static bool hb_persistent_enabled() {
    const char* e = getenv("HB_NO_PERSISTENT");
    return !(e != nullptr && e[0] == '1');
}

int hb_dev_ice_loop(hb_csr* h, int iters, double tol, void* stream) {
    HB_RANGE("hb_dev_ice_loop");
    HB_REQUIRE(h != nullptr && h->state_ready, "call hb_dev_ice_begin first");
    HB_REQUIRE(iters >= 0, "negative iteration count");
    HB_REQUIRE(h->world == 1 || h->peer, "hb_dev_ice_loop needs the exchange on the device (one GPU or hb_peer_attach)");
    if (iters == 0) return HB_OK;
    int rc = hb_loop_settle(h);
    if (rc) return rc;
    cudaStream_t s = pick_stream(h, stream);
    HbIceLoopArgs a;
    memset(&a, 0, sizeof(a));
    a.indptr = h->d_indptr;
    a.idx = h->d_indices;
    a.val = h->pack_kind ? h->d_pack : (const void*)h->d_data;
    a.n_rows = h->n_rows;
    a.row0 = h->row0;
    a.n_cols = h->n_cols;
    a.tile_row0 = h->d_tile_row0;
    a.tile_len = h->d_tile_len;
    a.tile_seg = h->d_tile_seg;
    a.seg_tile0 = h->d_seg_tile0;
    a.seg_off = h->d_seg_off;
    a.n_tiles = h->n_tiles;
    a.nseg = h->nseg;
    a.bias = h->d_bias;
    a.r = h->d_r;
    a.exch = h->d_exch;
    a.tile_sum = h->d_tile_sum;
    a.tile_cnt = h->d_tile_cnt;
    a.devbits = h->d_loop_devbits;
    a.wmax = h->d_loop_wmax;
    a.seg_dev = h->d_seg_dev;
    a.seg_done = h->d_seg_done;
    a.seg_iters = h->d_seg_iters;
    a.status = h->d_status;
    a.h_loop = h->h_loop_dev;
    a.row_counters = h->d_row_counter;
    a.err = h->d_loop_err;
    a.iters = iters;
    a.row_parity = h->spmv_parity;
    a.tol = tol;
    hb_fill_loop_peer(h, &a.peer);
    HB_CUDA(cudaMemsetAsync(h->d_loop_devbits, 0, sizeof(unsigned long long) * 2 * (size_t)h->nseg, s));
    HB_CUDA(cudaMemsetAsync(h->d_loop_wmax, 0, 24, s));   // wmax[2] | err
    const void* fn = h->pack_kind == 1   ? (const void*)hb_ice_loop_kernel<uint16_t>
                     : h->pack_kind == 2 ? (const void*)hb_ice_loop_kernel<float>
                                         : (const void*)hb_ice_loop_kernel<double>;
    const int grid = hb_coop_grid(fn, h->n_rows, h->device);
    const bool timing = getenv("HB_LOOP_TIMING") != nullptr;
    unsigned long long* d_timing = nullptr;
    if (timing) {
        HB_CUDA(hb_dmalloc(&d_timing, sizeof(unsigned long long) * 256));
        HB_CUDA(cudaMemsetAsync(d_timing, 0, sizeof(unsigned long long) * 256, s));
        a.timing = d_timing;
    }
    void* params[] = {(void*)&a};
    HB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(HB_SPMV_THREADS), params, 0, s));
    g_hb_launches.fetch_add(1, std::memory_order_relaxed);
    if (h->peer) {
        HB_CUDA(cudaEventRecord(h->loop_event, s));
        h->loop_pending = true;
    }
    if (timing) {   // debug aid: where an iteration spends its time, as seen by block 0 (ns)
        unsigned long long t[256];
        HB_CUDA(cudaStreamSynchronize(s));
        HB_CUDA(cudaMemcpy(t, d_timing, sizeof(t), cudaMemcpyDeviceToHost));
        hb_dfree(d_timing);
        const int m = iters < 32 ? iters : 32;
        double acc[8] = {0};
        int cnt = 0;
        for (int it = 1; it + 1 < m; ++it) {   // skip the first and the last iteration of the launch
            if (t[it * 8 + 7] == 0 || t[(it + 1) * 8] == 0) break;
            for (int k = 0; k < 7; ++k) acc[k] += (double)(t[it * 8 + k + 1] - t[it * 8 + k]);
            acc[7] += (double)(t[(it + 1) * 8] - t[it * 8 + 7]);
            ++cnt;
        }
        if (cnt > 0)
            fprintf(stderr,
                    "[hb loop timing, block 0, us, mean of %d iterations, grid %d] rows %.1f | barrier1 %.1f | publish+wait %.1f | tile sums %.1f | "
                    "barrier2 %.1f | update %.1f | barrier3 %.1f | convergence %.1f\n",
                    cnt, grid, acc[0] / cnt / 1e3, acc[1] / cnt / 1e3, acc[2] / cnt / 1e3, acc[3] / cnt / 1e3, acc[4] / cnt / 1e3,
                    acc[5] / cnt / 1e3, acc[6] / cnt / 1e3, acc[7] / cnt / 1e3);
    }
    return HB_OK;
}

int hb_loop_progress(hb_csr* h, int32_t* words16) {
    HB_REQUIRE(h != nullptr && h->state_ready && words16 != nullptr, "bad arguments");
    for (int i = 0; i < HB_LOOP_WORDS; ++i) words16[i] = ((volatile int32_t*)h->h_loop)[i];
    return HB_OK;
}

int hb_dev_ice_status(hb_csr* h, int32_t* pinned_status4, void* stream) {
    HB_REQUIRE(h != nullptr && h->state_ready && pinned_status4 != nullptr, "bad arguments");
    HB_CUDA(cudaMemcpyAsync(pinned_status4, h->d_status, sizeof(int32_t) * HB_ST_WORDS, cudaMemcpyDeviceToHost,
                            pick_stream(h, stream)));
    return HB_OK;
}

int hb_dev_ice_finish(hb_csr* h, void* stream) {
    HB_RANGE("hb_dev_ice_finish");
    HB_REQUIRE(h != nullptr && h->state_ready, "call hb_dev_ice_begin first");
    cudaStream_t s = pick_stream(h, stream);
    if (h->n_tiles == 0) return HB_OK;
    hb_ice_reduce_kernel<<<h->n_tiles, 256, 0, s>>>(h->d_bias, h->d_tile_row0, h->d_tile_len, h->d_tile_seg,
                                                     h->d_seg_tile0, h->nseg, nullptr, h->d_tile_sum, h->d_tile_cnt,
                                                     h->d_seg_corr, h->d_counters + 0, nullptr, nullptr, 0);
    HB_LAUNCH_CHECK();
    hb_ice_scale_bias_kernel<<<h->n_tiles, 256, 0, s>>>(h->d_bias, h->d_tile_row0, h->d_tile_len, h->d_tile_seg,
                                                         h->d_seg_corr);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

int hb_dev_ice_emit(hb_csr* h, double* d_out, double* d_max, void* stream) {
    HB_REQUIRE(h != nullptr && h->state_ready && d_out != nullptr, "bad arguments");
    cudaStream_t s = pick_stream(h, stream);
    HB_CUDA(cudaMemsetAsync(h->d_scalar_bits, 0, sizeof(unsigned long long) * 2, s));
    if (h->n_rows > 0 && h->nnz > 0) {
        hb_ice_emit_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                          h->d_r, h->d_seg_off, h->d_seg_corr, h->nseg, h->phase, d_out,
                                                          h->d_scalar_bits);
        HB_LAUNCH_CHECK();
    }
    if (d_max)
        HB_CUDA(cudaMemcpyAsync(d_max, h->d_scalar_bits, sizeof(double) * 2, cudaMemcpyDeviceToDevice, s));
    return HB_OK;
}

int hb_dev_ice_vectors(hb_csr* h, double** d_bias, double** d_r) {
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    if (d_bias) *d_bias = h->d_bias;
    if (d_r) *d_r = h->d_r;
    return HB_OK;
}

int hb_ice_results(hb_csr* h, double* bias_out, int32_t* iters_out, double* dev_out) {
    HB_REQUIRE(h != nullptr && h->state_ready, "no ICE state");
    int rc = hb_use_device(h);
    if (rc) return rc;
    HB_CUDA(cudaStreamSynchronize(h->stream));
    if (bias_out && h->n_cols)
        HB_CUDA(cudaMemcpy(bias_out, h->d_bias, sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToHost));
    if (iters_out) HB_CUDA(cudaMemcpy(iters_out, h->d_seg_iters, sizeof(int32_t) * h->nseg, cudaMemcpyDeviceToHost));
    if (dev_out) HB_CUDA(cudaMemcpy(dev_out, h->d_seg_dev, sizeof(double) * h->nseg, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_ice_run(hb_csr* h, int max_iter, double tol, double* bias_out, int32_t* iters_out, double* dev_out) {
    HB_RANGE("hb_ice_run");
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(h->world > 1 || (h->row0 == 0 && h->n_rows == h->n_cols),
               "hb_ice_run needs the full matrix on one device unless hb_set_comm was called");
    HB_REQUIRE(max_iter >= 0, "negative iteration count");
    int rc = hb_use_device(h);
    if (rc) return rc;
    rc = hb_dev_ice_begin(h, nullptr);
    if (rc) return rc;
    if ((h->world == 1 || h->peer) && hb_persistent_enabled()) {
        // the whole loop is one cooperative launch; the host waits once
        rc = hb_dev_ice_loop(h, max_iter, tol, nullptr);
        if (rc != HB_OK) return rc;
        HB_CUDA(cudaStreamSynchronize(h->stream));
        rc = hb_loop_settle(h);
        if (rc != HB_OK) return rc;
    } else {
        // step-level path (host collective between the SpMV and the update, or HB_NO_PERSISTENT=1)
        const int LAG = 4, SLOTS = 16;
        cudaEvent_t ev[SLOTS];
        for (int i = 0; i < SLOTS; ++i) HB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        bool stop = false;
        for (int k = 0; k < max_iter && !stop; ++k) {
            rc = hb_dev_ice_spmv(h, nullptr, h->rank, h->world, nullptr);
            if (rc == HB_OK && h->world > 1 && !h->peer) {
                // the one exchange of the iteration: marginals + per-rank maxima, summed over the ranks
                if (h->allreduce == nullptr ||
                    h->allreduce(h->comm_ctx, h->d_exch, h->n_cols + h->world, (void*)h->stream) != 0) {
                    hb_set_error("collective callback failed");
                    rc = HB_ERR_CUDA;
                }
            }
            if (rc == HB_OK) rc = hb_dev_ice_update(h, nullptr, h->world, tol, nullptr);
            if (rc == HB_OK) rc = hb_dev_ice_status(h, h->h_status + (k % SLOTS) * HB_ST_WORDS, nullptr);
            if (rc != HB_OK) break;
            cudaEventRecord(ev[k % SLOTS], h->stream);
            if (k >= LAG) {
                // look at the status of iteration k-LAG: the device queue stays LAG iterations deep and
                // iterations enqueued after convergence are no-ops (every kernel returns on ALL_DONE)
                int j = (k - LAG) % SLOTS;
                cudaEventSynchronize(ev[j]);
                if (h->h_status[j * HB_ST_WORDS + HB_ST_ALL_DONE]) stop = true;
            }
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        for (int i = 0; i < SLOTS; ++i) cudaEventDestroy(ev[i]);
        if (rc != HB_OK) return rc;
        if (e != cudaSuccess) return hb_cuda_fail(e, "ICE loop", __FILE__, __LINE__);
    }
    int32_t st[HB_ST_WORDS];
    HB_CUDA(cudaMemcpy(st, h->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[HB_ST_OVERFLOW] == 2) {
        hb_set_error("peer exchange timed out: a rank did not publish its rows");
        return HB_ERR_CUDA;
    }
    if (st[HB_ST_OVERFLOW]) {
        hb_set_error("matrix correction is producing extremely large values (> 1e100)");
        return HB_ERR_OVERFLOW_ITER;
    }
    rc = hb_dev_ice_finish(h, nullptr);
    if (rc) return rc;
    return hb_ice_results(h, bias_out, iters_out, dev_out);
}

int hb_ice_scaled_values(hb_csr* h, double* data_out, double* max_out) {
    HB_RANGE("hb_ice_scaled_values");
    HB_REQUIRE(h != nullptr && h->state_ready, "run ICE first");
    HB_REQUIRE(data_out != nullptr || h->nnz == 0, "null output");
    int rc = hb_use_device(h);
    if (rc) return rc;
    double* d_out = nullptr;
    HB_CUDA(hb_dmalloc(&d_out, sizeof(double) * (size_t)(h->nnz > 0 ? h->nnz : 1)));
    rc = hb_dev_ice_emit(h, d_out, nullptr, nullptr);
    double mx[2] = {0.0, 0.0};
    cudaError_t e = cudaSuccess;
    if (rc == HB_OK) {
        e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess && h->nnz && hb_copy_d2h(data_out, d_out, sizeof(double) * (size_t)h->nnz, h->device) != HB_OK)
            rc = HB_ERR_CUDA;
        if (e == cudaSuccess) e = cudaMemcpy(mx, h->d_scalar_bits, sizeof(mx), cudaMemcpyDeviceToHost);
    }
    hb_dfree(d_out);
    if (rc) return rc;
    if (e != cudaSuccess) return hb_cuda_fail(e, "emit", __FILE__, __LINE__);
    if (max_out) *max_out = mx[1];
    if (mx[0] > 1e100) {
        hb_set_error("matrix correction is producing extremely large values (> 1e100)");
        return HB_ERR_OVERFLOW_ITER;
    }
    if (mx[1] > 1e10) {
        hb_set_error("matrix correction produced extremely large values (> 1e10)");
        return HB_ERR_OVERFLOW_FINAL;
    }
    return HB_OK;
}

}  // extern "C"
