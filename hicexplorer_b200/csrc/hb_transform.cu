// Observed / expected transformation of hicTransform on the resident store.
// Replaces the per-entry Python work of hicexplorer/utilities.py:293-338 (expected_interactions_in_distance,
// expected_interactions_non_zero), :356-444 (expected_interactions: an O(n_distances x nnz) mask loop in the reference)
// and :479-600 (obs_exp_matrix_lieberman / _non_zero / obs_exp_matrix), as driven by hicexplorer/hicTransform.py:162-215.
//
//   hb_distance_stats   S[d] = sum of the stored non-zero values at genomic distance d = |row - col| and C[d] = their
//                       number, per segment (a chromosome block under --perChromosome; slot seg_off[s] + d).
//                       Many rows add into the same distance, so the sums go through atomics -- on INTEGERS: every value
//                       is split against the largest magnitude of the matrix into two 38-bit fixed-point limbs
//                       (v = k1 2^(e-38) + k2 2^(e-76) + <2^(e-76)), the limbs are accumulated with 64-bit integer
//                       atomics (associative: the result does not depend on the order, the run or the sharding) and
//                       converted back at the end.  Exact for counts, within 2^-76 of the largest value per term
//                       otherwise -- tighter than a floating-point sum in any order.
//   hb_obs_exp_apply    value <- observed / expected[index(d)] with the reference's index rule (d, or ceil(d / 2) --
//                       sic, utilities.py:489, 581), its arithmetic type (float32 observed over float64 expected) and its
//                       final cast (back to an integer input type by truncation, or rounded to float32), NaN/Inf -> 0.
// The O(n) step in between (sums -> expected values) stays on the host in numpy, in the reference's own expressions.
#include <cmath>
#include <cstring>

#include "hb_internal.cuh"

__device__ __forceinline__ int hb_seg_of(const int64_t* __restrict__ seg_off, int nseg, int64_t g) {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// largest finite magnitude among the stored values (bit pattern of a non-negative double orders like the integer)
__global__ void hb_absmax_kernel(const double* __restrict__ val, int64_t nnz, unsigned long long* out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long m = 0;
    for (; i < nnz; i += stride) {
        const double a = fabs(hb_ldg_stream_d1(val + i));
        if (a <= 1.7976931348623157e308) m = max(m, (unsigned long long)__double_as_longlong(a));
    }
    m = hb_warp_max_u64(m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

#define HB_FX_BITS 38   // bits per fixed-point limb: |limb| < 2^38, so 2^25 terms per distance fit a signed 64-bit sum

__global__ void __launch_bounds__(256)
hb_distance_stats_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                         int64_t n_rows, int64_t row0, const int64_t* __restrict__ seg_off, int nseg, double scale1,
                         long long* __restrict__ K1, long long* __restrict__ K2, double* __restrict__ P,
                         unsigned long long* __restrict__ C) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double scale2 = (double)(1ll << HB_FX_BITS);
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
                atomicAdd(C + lo + d, 1ull);
                if (isnan(v[j]) || isinf(v[j])) {
                    atomicAdd(P + lo + d, v[j]);          // poisons the distance like numpy's sum would (order-free)
                    continue;
                }
                const double a = v[j] * scale1;             // exact (power of two), |a| < 2^38
                const double f1 = floor(a);
                const double f2 = floor((a - f1) * scale2); // a - f1 in [0, 1) is exact; what lies below is dropped
                atomicAdd(reinterpret_cast<unsigned long long*>(K1 + lo + d), (unsigned long long)(long long)f1);
                atomicAdd(reinterpret_cast<unsigned long long*>(K2 + lo + d), (unsigned long long)(long long)f2);
            }
        }
    }
}

__global__ void hb_distance_finish_kernel(const long long* __restrict__ K1, const long long* __restrict__ K2,
                                          const double* __restrict__ P, int64_t n, int exp1, double* __restrict__ S) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) S[i] = (ldexp((double)K1[i], exp1) + ldexp((double)K2[i], exp1 - HB_FX_BITS)) + P[i];
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
    HB_RANGE("hb_distance_stats");
    HB_REQUIRE(h != nullptr && sums_out != nullptr, "bad arguments");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_distance_stats needs the full matrix on one device");
    int rc = hb_ensure_state(h);   // segment table on the device
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const size_t n = (size_t)(h->n_cols > 0 ? h->n_cols : 1);
    HbScratch scratch;
    unsigned char* d_buf = nullptr;
    // S | K1 | K2 | P | C  (8 bytes each per distance slot) + the magnitude word
    HB_CUDA(scratch.alloc(&d_buf, 8 * (5 * n + 1)));
    double* d_S = reinterpret_cast<double*>(d_buf);
    long long* d_K1 = reinterpret_cast<long long*>(d_buf + 8 * n);
    long long* d_K2 = reinterpret_cast<long long*>(d_buf + 16 * n);
    double* d_P = reinterpret_cast<double*>(d_buf + 24 * n);
    unsigned long long* d_C = reinterpret_cast<unsigned long long*>(d_buf + 32 * n);
    unsigned long long* d_max = reinterpret_cast<unsigned long long*>(d_buf + 40 * n);
    HB_CUDA(cudaMemsetAsync(d_buf, 0, 8 * (5 * n + 1), s));
    if (h->n_rows > 0 && h->nnz > 0) {
        hb_absmax_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_data + h->phase, h->nnz, d_max);
        g_hb_launches.fetch_add(1);
        unsigned long long mbits = 0;
        HB_CUDA(cudaMemcpyAsync(&mbits, d_max, sizeof(mbits), cudaMemcpyDeviceToHost, s));
        HB_CUDA(cudaStreamSynchronize(s));
        double vmax;
        memcpy(&vmax, &mbits, sizeof(vmax));
        int e = 0;
        if (vmax > 0.0) frexp(vmax, &e);      // vmax = m 2^e, m in [0.5, 1): every |v| < 2^e
        hb_distance_stats_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                                h->d_seg_off, h->nseg, ldexp(1.0, HB_FX_BITS - e), d_K1, d_K2, d_P,
                                                                d_C);
        hb_distance_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_K1, d_K2, d_P, (int64_t)n, e - HB_FX_BITS, d_S);
        g_hb_launches.fetch_add(2);
    }
    HB_CUDA(cudaStreamSynchronize(s));
    HB_CUDA(cudaGetLastError());
    if (h->n_cols > 0) HB_CUDA(cudaMemcpy(sums_out, d_S, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (h->n_cols > 0 && counts_out) HB_CUDA(cudaMemcpy(counts_out, d_C, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_obs_exp_apply(hb_csr* h, const double* expected, int index_mode, int out_mode, const double* row_sum,
                     const double* seg_total) {
    HB_RANGE("hb_obs_exp_apply");
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
    HB_CUDA(hb_dmalloc(&d_buf, sizeof(double) * (2 * n + (size_t)h->nseg)));
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
    hb_dfree(d_buf);
    if (e != cudaSuccess) return hb_cuda_fail(e, "obs/exp", __FILE__, __LINE__);
    return HB_OK;
}

}  // extern "C"
