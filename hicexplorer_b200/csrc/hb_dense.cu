// hicTransform --method covariance | pearson (hicexplorer/hicTransform.py:105-113, 213-248): np.cov / np.corrcoef of
// a DENSE diagonal block of the matrix (rows = variables, columns = observations).  The only GEMM-shaped work near this
// path, and a plain one: the block is densified in HBM, centred, and handed to cuBLAS (fp64 DGEMM; the library is
// opened at run time so that libhicbalance.so itself keeps no dependency on it); the element-wise steps around the
// product -- row means, centring, the 1/(m-1) factor, corrcoef's two divisions, clipping, NaN/Inf scrub -- and the
// extraction of the non-zero cells as CSR are small kernels here.  Order of operations follows numpy's (np.cov:
// `X -= avg[:, None]; c = dot(X, X.T); c *= 1/fact`; np.corrcoef: `c /= stddev[:, None]; c /= stddev[None, :]`, clip).
#include <dlfcn.h>

#include <mutex>

#include "hb_internal.cuh"

namespace {

typedef void* cublasHandle_t;
typedef int (*cublasCreate_t)(cublasHandle_t*);
typedef int (*cublasDestroy_t)(cublasHandle_t);
typedef int (*cublasSetStream_t)(cublasHandle_t, cudaStream_t);
typedef int (*cublasDgemm_t)(cublasHandle_t, int, int, int, int, int, const double*, const double*, int, const double*, int,
                             const double*, double*, int);

struct Cublas {
    void* lib = nullptr;
    cublasCreate_t create = nullptr;
    cublasDestroy_t destroy = nullptr;
    cublasSetStream_t set_stream = nullptr;
    cublasDgemm_t dgemm = nullptr;
};

Cublas* load_cublas() {
    static Cublas api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12",
                               "/usr/local/cuda/lib64/libcublas.so"};
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.create = (cublasCreate_t)dlsym(api.lib, "cublasCreate_v2");
            api.destroy = (cublasDestroy_t)dlsym(api.lib, "cublasDestroy_v2");
            api.set_stream = (cublasSetStream_t)dlsym(api.lib, "cublasSetStream_v2");
            api.dgemm = (cublasDgemm_t)dlsym(api.lib, "cublasDgemm_v2");
        }
    });
    return (api.lib && api.create && api.destroy && api.set_stream && api.dgemm) ? &api : nullptr;
}

// X[(r-lo)*m + (c-lo)] = A_rc for the block's stored entries (X zeroed before)
__global__ void __launch_bounds__(256)
densify_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t lo,
               int64_t hi, double* __restrict__ X) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t m = hi - lo;
    for (int64_t r = lo + gwarp; r < hi; r += nwarps) {
        for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
            const int64_t c = idx[k];
            if (c >= lo && c < hi) X[(r - lo) * m + (c - lo)] = val[k];
        }
    }
}

// one CTA per row: mean (np.average(X, axis=1)), then X[i, :] -= mean
__global__ void __launch_bounds__(256) centre_kernel(double* __restrict__ X, int64_t m) {
    __shared__ double sh[8];
    __shared__ double s_mean;
    double* row = X + (int64_t)blockIdx.x * m;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < m; j += blockDim.x) s += row[j];
    s = hb_warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        s_mean = t / (double)m;
    }
    __syncthreads();
    const double mean = s_mean;
    for (int64_t j = threadIdx.x; j < m; j += blockDim.x) row[j] -= mean;
}

// c *= 1/fact ; pearson: c /= stddev[:, None]; c /= stddev[None, :]; clip to [-1, 1]; NaN / Inf -> 0
__global__ void __launch_bounds__(256)
finish_kernel(double* __restrict__ C, const double* __restrict__ diag_scaled, int64_t m, double inv_fact, int pearson) {
    const int64_t total = m * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        double c = C[t] * inv_fact;
        if (pearson) {
            const int64_t i = t / m, j = t - i * m;
            c = c / sqrt(diag_scaled[i]);
            c = c / sqrt(diag_scaled[j]);
            if (c > 1.0) c = 1.0;             // np.clip(c, -1, 1): comparisons, so that NaN stays NaN
            else if (c < -1.0) c = -1.0;
            if (isnan(c) || isinf(c)) c = 0.0;   // convertNansToZeros / convertInfsToZeros (hicTransform.py:109-110)
        }                                        // np.cov's result is stored as it is (hicTransform.py:236, 243)
        C[t] = c;
    }
}

__global__ void diag_kernel(const double* __restrict__ C, double* __restrict__ diag_scaled, int64_t m, double inv_fact) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) diag_scaled[i] = C[i * m + i] * inv_fact;
}

__global__ void __launch_bounds__(256) count_nonzero_kernel(const double* __restrict__ C, int64_t m, int64_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = gwarp; r < m; r += nwarps) {
        long long c = 0;
        for (int64_t j = lane; j < m; j += 32) c += C[r * m + j] != 0.0 ? 1 : 0;   // NaN != 0: kept, like csr_matrix(dense)
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[r] = c;
    }
}

__global__ void __launch_bounds__(256)
fill_nonzero_kernel(const double* __restrict__ C, int64_t m, const int64_t* __restrict__ ptr, int32_t col0, int32_t* __restrict__ idx,
                    double* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = gwarp; r < m; r += nwarps) {
        int64_t w = ptr[r];
        for (int64_t base = 0; base < m; base += 32) {
            const int64_t j = base + lane;
            const double c = j < m ? C[r * m + j] : 0.0;
            const bool keep = j < m && c != 0.0;
            const unsigned int bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int64_t p = w + __popc(bal & ((1u << lane) - 1u));
                idx[p] = col0 + (int32_t)j;
                val[p] = c;
            }
            w += __popc(bal);
        }
    }
}

}  // namespace

extern "C" int hb_block_correlation(hb_csr* h, int64_t lo, int64_t hi, int pearson, int64_t* nnz_out, int64_t* indptr,
                                    int32_t* indices, double* data) {
    HB_RANGE("hb_block_correlation");
    HB_REQUIRE(h != nullptr && nnz_out != nullptr, "bad arguments");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols && h->phase == 0, "hb_block_correlation needs the full matrix on one device");
    HB_REQUIRE(lo >= 0 && lo <= hi && hi <= h->n_cols, "bad block");
    HB_REQUIRE((indptr == nullptr) == (indices == nullptr) && (indptr == nullptr) == (data == nullptr), "all outputs or none");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int64_t m = hi - lo;
    if (indptr == nullptr) {
        // first call: compute the dense block and keep it in the handle; report how many cells are non-zero
        hb_dfree(h->d_dense);
        h->d_dense = nullptr;
        h->dense_m = 0;
        *nnz_out = 0;
        if (m == 0) return HB_OK;
        size_t free_b = 0, total_b = 0;
        HB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if ((double)m * (double)m * 16.0 + (double)(64 << 20) > (double)free_b) {
            hb_set_error("a dense %lld x %lld block does not fit the device (%.1f GB needed, %.1f GB free): use --perChromosome",
                         (long long)m, (long long)m, (double)m * m * 16.0 / 1e9, (double)free_b / 1e9);
            return HB_ERR_INVALID_ARG;
        }
        Cublas* api = load_cublas();
        if (api == nullptr) {
            hb_set_error("cuBLAS (libcublas.so.12) could not be opened: %s", dlerror() ? dlerror() : "symbols missing");
            return HB_ERR_CUDA;
        }
        double *X = nullptr, *C = nullptr, *diag = nullptr;
        HbScratch scratch;
        HB_CUDA(scratch.alloc(&X, sizeof(double) * (size_t)(m * m)));
        HB_CUDA(scratch.alloc(&diag, sizeof(double) * (size_t)m));
        HB_CUDA(hb_dmalloc(&C, sizeof(double) * (size_t)(m * m)));
        h->d_dense = C;
        h->dense_m = m;
        HB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * (size_t)(m * m), s));
        densify_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, lo, hi, X);
        HB_LAUNCH_CHECK();
        centre_kernel<<<(unsigned)m, 256, 0, s>>>(X, m);
        HB_LAUNCH_CHECK();
        cublasHandle_t cb = nullptr;
        if (api->create(&cb) != 0 || api->set_stream(cb, s) != 0) {
            hb_set_error("cublasCreate failed");
            return HB_ERR_CUDA;
        }
        // row-major X is column-major X^T: C = X X^T = (X_cm)^T (X_cm); C is symmetric, so its layout does not matter
        const double one = 1.0, zero = 0.0;
        const int st = api->dgemm(cb, /*CUBLAS_OP_T*/ 1, /*CUBLAS_OP_N*/ 0, (int)m, (int)m, (int)m, &one, X, (int)m, X, (int)m, &zero, C,
                                  (int)m);
        g_hb_launches.fetch_add(1);
        cudaError_t e = cudaStreamSynchronize(s);
        api->destroy(cb);
        if (st != 0 || e != cudaSuccess) {
            hb_set_error("cublasDgemm failed (status %d, CUDA %d)", st, (int)e);
            return HB_ERR_CUDA;
        }
        const double fact = (double)m - 1.0;                    // ddof = 1
        const double inv_fact = fact > 0.0 ? 1.0 / fact : INFINITY;   // numpy: fact <= 0 -> 0.0 -> c *= inf
        diag_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(C, diag, m, inv_fact);
        HB_LAUNCH_CHECK();
        finish_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(C, diag, m, inv_fact, pearson);
        HB_LAUNCH_CHECK();
        int64_t* d_cnt = nullptr;
        HB_CUDA(scratch.alloc(&d_cnt, sizeof(int64_t) * (size_t)m));
        count_nonzero_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(C, m, d_cnt);
        HB_LAUNCH_CHECK();
        std::vector<int64_t> cnt((size_t)m);
        HB_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost, s));
        HB_CUDA(cudaStreamSynchronize(s));
        int64_t total = 0;
        for (int64_t i = 0; i < m; ++i) total += cnt[(size_t)i];
        *nnz_out = total;
        return HB_OK;
    }
    // second call: the non-zero cells as CSR (block-local rows, GLOBAL column ids), then the dense block is released
    HB_REQUIRE(h->d_dense != nullptr && h->dense_m == m, "call with NULL outputs first");
    double* C = h->d_dense;
    HbScratch scratch;
    int64_t *d_cnt = nullptr, *d_ptr = nullptr;
    HB_CUDA(scratch.alloc(&d_cnt, sizeof(int64_t) * (size_t)(m + 1)));
    HB_CUDA(scratch.alloc(&d_ptr, sizeof(int64_t) * (size_t)(m + 1)));
    count_nonzero_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(C, m, d_cnt);
    HB_LAUNCH_CHECK();
    std::vector<int64_t> cnt((size_t)m);
    HB_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    indptr[0] = 0;
    for (int64_t i = 0; i < m; ++i) indptr[i + 1] = indptr[i] + cnt[(size_t)i];
    const int64_t total = indptr[m];
    *nnz_out = total;
    if (total > 0) {
        int32_t* d_idx = nullptr;
        double* d_val = nullptr;
        HB_CUDA(scratch.alloc(&d_idx, sizeof(int32_t) * (size_t)total));
        HB_CUDA(scratch.alloc(&d_val, sizeof(double) * (size_t)total));
        HB_CUDA(cudaMemcpyAsync(d_ptr, indptr, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyHostToDevice, s));
        fill_nonzero_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(C, m, d_ptr, (int32_t)lo, d_idx, d_val);
        HB_LAUNCH_CHECK();
        HB_CUDA(cudaStreamSynchronize(s));
        if (hb_copy_d2h(indices, d_idx, sizeof(int32_t) * (size_t)total, h->device) != HB_OK) return HB_ERR_CUDA;
        if (hb_copy_d2h(data, d_val, sizeof(double) * (size_t)total, h->device) != HB_OK) return HB_ERR_CUDA;
    }
    hb_dfree(h->d_dense);
    h->d_dense = nullptr;
    h->dense_m = 0;
    return HB_OK;
}
