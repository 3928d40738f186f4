// The one streaming kernel of the path: a row-block SpMV  t_i = sum_j A_ij u_j  with a fused
// per-row epilogue.  ICE (one iteration == one SpMV with the ORIGINAL matrix) and
// Knight-Ruiz (every CG step == one SpMV) both instantiate it.
//
// Roofline: HBM.  Algorithmic bytes per launch = 12 B per stored entry (int32 column id + fp64
// value, each read exactly once) + ~24 B per row.  The gathered vector u (n*8 B: 1-25 MB) stays
// in L2/L1; the matrix streams through with L1::no_allocate so it does not evict u.
//
// Mapping: persistent grid (148 SMs x 4 CTAs x 8 warps); a warp owns a row, lanes read the row
// with 128-bit loads (4 column ids / 2 values per load, 8 entries per lane in flight per step),
// partial sums are combined with __shfl_xor.  Rows are claimed HB_ROWS_PER_GRAB at a time from a
// global counter, the next claim being issued before the current rows are processed, so long and
// short rows (Hi-C rows range from a handful to tens of thousands of entries) balance out.
#pragma once
#include "hb_internal.cuh"

template <bool WANT_MAX>
__device__ __forceinline__ void hb_row_dot(const int32_t* __restrict__ idx, const double* __restrict__ val,
                                           const double* __restrict__ u, int64_t s, int64_t e, int lane,
                                           double& sum_out, double& max_out) {
    double acc0 = 0.0, acc1 = 0.0;
    double mx = -INFINITY;
    // head: up to 3 entries until the entry index is a multiple of 4 (=> 16-byte aligned ids and values)
    int64_t a = (s + 3) & ~(int64_t)3;
    if (a > e) a = e;
    {
        int64_t k = s + lane;
        if (k < a) {
            double p = hb_ldg_stream_d1(val + k) * __ldg(u + hb_ldg_stream_i1(idx + k));
            acc0 = p;
            if (WANT_MAX) mx = p;
        }
    }
    const int64_t ng = (e - a) >> 2;
    const int32_t* idx_a = idx + a;
    const double* val_a = val + a;
    int64_t g = lane;
    for (; g + 32 < ng; g += 64) {
        const int4 c0 = hb_ldg_stream_i4(idx_a + 4 * g);
        const int4 c1 = hb_ldg_stream_i4(idx_a + 4 * (g + 32));
        const double2 v00 = hb_ldg_stream_d2(val_a + 4 * g);
        const double2 v01 = hb_ldg_stream_d2(val_a + 4 * g + 2);
        const double2 v10 = hb_ldg_stream_d2(val_a + 4 * (g + 32));
        const double2 v11 = hb_ldg_stream_d2(val_a + 4 * (g + 32) + 2);
        const double u0 = __ldg(u + c0.x), u1 = __ldg(u + c0.y), u2 = __ldg(u + c0.z), u3 = __ldg(u + c0.w);
        const double u4 = __ldg(u + c1.x), u5 = __ldg(u + c1.y), u6 = __ldg(u + c1.z), u7 = __ldg(u + c1.w);
        if (WANT_MAX) {
            const double p0 = v00.x * u0, p1 = v00.y * u1, p2 = v01.x * u2, p3 = v01.y * u3;
            const double p4 = v10.x * u4, p5 = v10.y * u5, p6 = v11.x * u6, p7 = v11.y * u7;
            acc0 += p0; acc0 += p1; acc0 += p2; acc0 += p3;
            acc1 += p4; acc1 += p5; acc1 += p6; acc1 += p7;
            mx = fmax(mx, fmax(fmax(fmax(p0, p1), fmax(p2, p3)), fmax(fmax(p4, p5), fmax(p6, p7))));
        } else {
            acc0 = fma(v00.x, u0, acc0); acc0 = fma(v00.y, u1, acc0);
            acc0 = fma(v01.x, u2, acc0); acc0 = fma(v01.y, u3, acc0);
            acc1 = fma(v10.x, u4, acc1); acc1 = fma(v10.y, u5, acc1);
            acc1 = fma(v11.x, u6, acc1); acc1 = fma(v11.y, u7, acc1);
        }
    }
    if (g < ng) {
        const int4 c0 = hb_ldg_stream_i4(idx_a + 4 * g);
        const double2 v00 = hb_ldg_stream_d2(val_a + 4 * g);
        const double2 v01 = hb_ldg_stream_d2(val_a + 4 * g + 2);
        const double u0 = __ldg(u + c0.x), u1 = __ldg(u + c0.y), u2 = __ldg(u + c0.z), u3 = __ldg(u + c0.w);
        if (WANT_MAX) {
            const double p0 = v00.x * u0, p1 = v00.y * u1, p2 = v01.x * u2, p3 = v01.y * u3;
            acc0 += p0; acc0 += p1; acc0 += p2; acc0 += p3;
            mx = fmax(mx, fmax(fmax(p0, p1), fmax(p2, p3)));
        } else {
            acc0 = fma(v00.x, u0, acc0); acc0 = fma(v00.y, u1, acc0);
            acc0 = fma(v01.x, u2, acc0); acc0 = fma(v01.y, u3, acc0);
        }
    }
    {
        int64_t k = a + 4 * ng + lane;
        if (k < e) {
            double p = hb_ldg_stream_d1(val + k) * __ldg(u + hb_ldg_stream_i1(idx + k));
            acc1 += p;
            if (WANT_MAX) mx = fmax(mx, p);
        }
    }
    sum_out = hb_warp_sum(acc0 + acc1);
    if (WANT_MAX) max_out = hb_warp_max(mx);
}

// Epi must provide:
//   __device__ bool active(int64_t global_row) const;                 // skip rows of finished segments
//   __device__ void row(int64_t global_row, double t, double m);      // lane 0 only
//   __device__ void warp_done();                                      // lane 0 only, once per warp
template <class Epi, bool WANT_MAX>
__global__ void __launch_bounds__(HB_SPMV_THREADS, HB_SPMV_CTAS_PER_SM)
hb_spmv_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
               int64_t n_rows, int64_t row0, const double* __restrict__ u, unsigned long long* counters, int parity,
               const int32_t* __restrict__ status, Epi epi) {
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
                hb_row_dot<WANT_MAX>(idx, val, u, s, e, lane, t, m);
                if (lane == 0) epi.row(row0 + row, t, m);
            }
        }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        batch = nwarps * HB_ROWS_PER_GRAB + (int64_t)nxt;
    }
    if (lane == 0) epi.warp_done();
}

// out[g] = a_g * (t + eps * u_g) + b_g * c_g      (Knight-Ruiz: w = x o (A (x o p)) + v o p ; v = x o (A x))
struct HbAffineEpi {
    const double* __restrict__ u;
    const double* __restrict__ a;
    const double* __restrict__ b;
    const double* __restrict__ c;
    double* __restrict__ out;
    double eps;
    __device__ __forceinline__ bool active(int64_t) const { return true; }
    __device__ __forceinline__ void row(int64_t g, double t, double) const {
        double v = t + eps * u[g];
        if (a != nullptr) v *= a[g];
        if (b != nullptr && c != nullptr) v += b[g] * c[g];
        out[g] = v;
    }
    __device__ __forceinline__ void warp_done() const {}
};
