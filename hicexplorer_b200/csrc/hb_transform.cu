// Observed / expected transformation of hicTransform on the resident store.
// Replaces the per-entry Python work of hicexplorer/utilities.py:293-338 (expected_interactions_in_distance,
// expected_interactions_non_zero), :356-444 (expected_interactions: an O(n_distances x nnz) mask loop in the reference)
// and :479-600 (obs_exp_matrix_lieberman / _non_zero / obs_exp_matrix), as driven by hicexplorer/hicTransform.py:162-215.
//
//   hb_distance_stats   S[d] = sum of the stored non-zero values at genomic distance d = |row - col| and C[d] = their
//                       number, per segment (a chromosome block under --perChromosome; slot seg_off[s] + d).
//                       One streaming pass; fp64 atomics (exact for counts; for already-corrected float matrices the
//                       sums agree with numpy's to rounding and may differ in the last bits between runs).
//   hb_obs_exp_apply    value <- observed / expected[index(d)] with the reference's index rule (d, or ceil(d / 2) --
//                       sic, utilities.py:489, 581), its arithmetic type (float32 observed over float64 expected) and its
//                       final cast (back to an integer input type by truncation, or rounded to float32), NaN/Inf -> 0.
// The O(n) step in between (sums -> expected values) stays on the host in numpy, in the reference's own expressions.
#include "hb_internal.cuh"

__device__ __forceinline__ int hb_seg_of(const int64_t* __restrict__ seg_off, int nseg, int64_t g) {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
hb_distance_stats_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                         int64_t n_rows, int64_t row0, const int64_t* __restrict__ seg_off, int nseg, double* __restrict__ S,
                         unsigned long long* __restrict__ C) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int64_t lo = 0, hi = INT64_MAX;
        if (nseg > 1) {
            const int s = hb_seg_of(seg_off, nseg, g);
            lo = seg_off[s];
            hi = seg_off[s + 1];
        }
        const int64_t e = indptr[row + 1];
        for (int64_t k = indptr[row] + lane; k < e; k += 32 * 4) {
            int c[4];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = (k + 32 * j < e) ? hb_ldg_stream_i1(idx + k + 32 * j) : -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (k + 32 * j < e) ? hb_ldg_stream_d1(val + k + 32 * j) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // `row, col = matrix.nonzero()`: explicit zeros do not take part; entries outside the row's own
                // segment do not exist in a chromosome's submatrix
                if (c[j] < 0 || v[j] == 0.0 || c[j] < lo || c[j] >= hi) continue;
                const int64_t d = g > c[j] ? g - c[j] : c[j] - g;
                atomicAdd(S + lo + d, v[j]);
                atomicAdd(C + lo + d, 1ull);
            }
        }
    }
}

// index_mode 0: expected[d]          (obs_exp_non_zero, utilities.py:531)
//            1: expected[ceil(d/2)]  (obs_exp, obs_exp_lieberman: :489, :581)
// out_mode   0: float64 quotient kept                                   (float64 input matrices)
//            1: quotient truncated toward zero to an integer            (`.astype(data_type)` of an integer input)
//            2: quotient rounded to float32                             (obs_exp_non_zero: stored into a float32 array;
//                                                                        also a float32 input matrix)
// ligation (obs_exp_non_zero --ligation_factor): expected *= row_sum[row] * row_sum[col] / seg_total   (:533-534)
__global__ void __launch_bounds__(256)
hb_obs_exp_apply_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, double* __restrict__ val,
                        int64_t n_rows, int64_t row0, const int64_t* __restrict__ seg_off, int nseg,
                        const double* __restrict__ expected, int index_mode, int out_mode, const double* __restrict__ row_sum,
                        const double* __restrict__ seg_total) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int s = 0;
        int64_t lo = 0, hi = INT64_MAX;
        if (nseg > 1) {
            s = hb_seg_of(seg_off, nseg, g);
            lo = seg_off[s];
            hi = seg_off[s + 1];
        }
        const double rs_g = row_sum ? row_sum[g] : 0.0;
        const double tot = row_sum ? seg_total[s] : 0.0;
        const int64_t e = indptr[row + 1];
        for (int64_t k = indptr[row] + lane; k < e; k += 32) {
            const int c = idx[k];
            const double v = val[k];
            if (c < lo || c >= hi) {     // not part of any chromosome submatrix: absent from the output
                val[k] = 0.0;
                continue;
            }
            if (v == 0.0) continue;
            const int64_t d = g > c ? g - c : c - g;
            double ex = expected[lo + (index_mode ? (d + 1) / 2 : d)];
            if (row_sum) ex = __dmul_rn(ex, __ddiv_rn(__dmul_rn(rs_g, row_sum[c]), tot));
            double q = __ddiv_rn((double)(float)v, ex);       // observed in float32, expected in float64
            if (isnan(q) || isinf(q)) q = 0.0;
            if (out_mode == 1) q = trunc(q);
            else if (out_mode == 2) {
                q = (double)(float)q;
                if (isinf(q)) q = 0.0;
            }
            val[k] = q;
        }
    }
}

extern "C" {

int hb_distance_stats(hb_csr* h, double* sums_out, int64_t* counts_out) {
    HB_REQUIRE(h != nullptr && sums_out != nullptr, "bad arguments");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_distance_stats needs the full matrix on one device");
    int rc = hb_ensure_state(h);   // segment table on the device
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const size_t n = (size_t)(h->n_cols > 0 ? h->n_cols : 1);
    double* d_S = nullptr;
    HB_CUDA(cudaMalloc(&d_S, (sizeof(double) + sizeof(unsigned long long)) * n));
    unsigned long long* d_C = reinterpret_cast<unsigned long long*>(d_S + n);
    cudaError_t e = cudaMemsetAsync(d_S, 0, (sizeof(double) + sizeof(unsigned long long)) * n, s);
    if (e == cudaSuccess && h->n_rows > 0 && h->nnz > 0) {
        hb_distance_stats_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                                h->d_seg_off, h->nseg, d_S, d_C);
        g_hb_launches.fetch_add(1);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && h->n_cols > 0) e = cudaMemcpy(sums_out, d_S, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h->n_cols > 0 && counts_out)
        e = cudaMemcpy(counts_out, d_C, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost);
    cudaFree(d_S);
    if (e != cudaSuccess) return hb_cuda_fail(e, "distance statistics", __FILE__, __LINE__);
    return HB_OK;
}

int hb_obs_exp_apply(hb_csr* h, const double* expected, int index_mode, int out_mode, const double* row_sum,
                     const double* seg_total) {
    HB_REQUIRE(h != nullptr && expected != nullptr, "bad arguments");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_obs_exp_apply needs the full matrix on one device");
    HB_REQUIRE(index_mode >= 0 && index_mode <= 1 && out_mode >= 0 && out_mode <= 2, "bad mode");
    HB_REQUIRE((row_sum == nullptr) == (seg_total == nullptr), "row_sum and seg_total go together");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);
    const size_t n = (size_t)(h->n_cols > 0 ? h->n_cols : 1);
    double* d_buf = nullptr;
    HB_CUDA(cudaMalloc(&d_buf, sizeof(double) * (2 * n + (size_t)h->nseg)));
    double* d_rs = row_sum ? d_buf + n : nullptr;
    double* d_tot = row_sum ? d_buf + 2 * n : nullptr;
    cudaError_t e = cudaMemcpyAsync(d_buf, expected, sizeof(double) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && row_sum) e = cudaMemcpyAsync(d_rs, row_sum, sizeof(double) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && row_sum) e = cudaMemcpyAsync(d_tot, seg_total, sizeof(double) * (size_t)h->nseg, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && h->n_rows > 0 && h->nnz > 0) {
        hb_obs_exp_apply_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                               h->d_seg_off, h->nseg, d_buf, index_mode, out_mode, d_rs, d_tot);
        g_hb_launches.fetch_add(1);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_buf);
    if (e != cudaSuccess) return hb_cuda_fail(e, "obs/exp", __FILE__, __LINE__);
    return HB_OK;
}

}  // extern "C"
