// Element-wise steps of the workflow around the correction, on the resident store:
//   hicNormalize  (hicexplorer/hicNormalize.py:63-153)  value statistics + the three normalisations, in the float32
//                 arithmetic the reference computes in (`data.astype(np.float32)`), NaN/Inf scrub, threshold-to-zero
//                 and eliminate_zeros.
// (hicSumMatrices' `A + B` reuses the windowed row merge of hb_merge.cu: hb_csr_add.)
// Everything streams the store once or twice; nothing here is tensor-core shaped.
#include <cmath>
#include <cstring>

#include "hb_internal.cuh"
#include "hb_scan.cuh"

// sum (fp64, per row then over rows in tile order -> deterministic) and float32 min / max of the stored values
__global__ void __launch_bounds__(256)
hb_value_stats_kernel(const int64_t* __restrict__ indptr, const double* __restrict__ val, int64_t n_rows, int scrub,
                      double* __restrict__ row_sum, unsigned int* mm /* [0] ordered min, [1] ordered max of f32 bits */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        double acc = 0.0;
        const int64_t e = indptr[row + 1];
        for (int64_t k = indptr[row] + lane; k < e; k += 32) {
            const double v = hb_ldg_stream_d1(val + k);
            acc += v;
            float f = (float)v;
            if (scrub && (isnan(f) || isinf(f))) f = 0.0f;
            mn = fminf(mn, f);   // NaN-ignoring; with scrub there are none (np.min would propagate NaN otherwise)
            mx = fmaxf(mx, f);
        }
        acc = hb_warp_sum(acc);
        if (lane == 0) row_sum[row] = acc;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        // monotone map float -> uint so that atomicMin / atomicMax order like the floats
        auto ord = [](float f) {
            const unsigned int b = __float_as_uint(f);
            return (b >> 31) ? ~b : (b | 0x80000000u);
        };
        if (mn != INFINITY) atomicMin(mm + 0, ord(mn));
        if (mx != -INFINITY) atomicMax(mm + 1, ord(mx));
    }
}

__global__ void hb_sum_rows_kernel(const double* __restrict__ row_sum, int64_t n, double* out) {
    // one block, fixed order: the same total however the rows were produced
    __shared__ double part[1024];
    double s = 0.0;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = threadIdx.x * chunk, e = (b + chunk < n) ? b + chunk : n;
    for (int64_t i = b; i < e; ++i) s += row_sum[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 1024; ++i) t += part[i];
        *out = t;
    }
}

// mode 0: (f - a) / b   norm_range   (hicNormalize.py:76-88)
// mode 1: f / a         smallest     (:104-113)
// mode 2: f * a         multiplicative (:131-139)
// mode 3: f             the matrix with the smallest sum in `smallest` mode is only cast (:105-106, 115-121)
// all in float32 with single correctly rounded operations; NaN / Inf -> 0 before and after; values below the
// threshold -> 0 (:95-96).  Results are stored back as fp64 (exact).
__global__ void hb_normalize_kernel(double* __restrict__ val, int64_t nnz, int mode, float a, float b, float thr, int scrub_first) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        float f = (float)val[i];
        if (scrub_first && (isnan(f) || isinf(f))) f = 0.0f;
        if (mode == 0) f = __fdiv_rn(__fsub_rn(f, a), b);
        else if (mode == 1) f = __fdiv_rn(f, a);
        else if (mode == 2) f = __fmul_rn(f, a);
        if (isnan(f) || isinf(f)) f = 0.0f;
        if (f < thr) f = 0.0f;
        val[i] = (double)f;
    }
}

// ---- eliminate_zeros: drop the entries whose value is exactly 0 --------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
hb_nonzero_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                  int64_t n_rows, int64_t* __restrict__ cnt, const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_idx,
                  double* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        int64_t w = FILL ? out_ptr[row] : 0;
        long long n_out = 0;
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t base = s; base < e; base += 32) {
            const int64_t k = base + lane;
            const double v = k < e ? val[k] : 0.0;
            const bool kp = k < e && v != 0.0;      // NaN != 0 is kept, like scipy
            const unsigned int bal = __ballot_sync(0xffffffffu, kp);
            if (FILL && kp) {
                const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                out_idx[pos] = idx[k];
                out_val[pos] = v;
            }
            w += __popc(bal);
            n_out += __popc(bal);
        }
        if (!FILL && lane == 0) cnt[row] = n_out;
    }
}

extern "C" {

int hb_csr_value_stats(hb_csr* h, int scrub, double* sum_out, double* min_out, double* max_out) {
    HB_RANGE("hb_csr_value_stats");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const size_t n = (size_t)(h->n_rows > 0 ? h->n_rows : 1);
    double* d_rows = nullptr;
    HB_CUDA(hb_dmalloc(&d_rows, sizeof(double) * (n + 2)));
    unsigned int* d_mm = reinterpret_cast<unsigned int*>(d_rows + n + 1);
    const unsigned int init[2] = {0xffffffffu, 0u};
    double total = 0.0;
    unsigned int mm[2] = {0xffffffffu, 0u};
    cudaError_t e = cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_rows, 0, sizeof(double) * (n + 1), s);
    if (e == cudaSuccess && h->n_rows > 0) {
        hb_value_stats_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_data, h->n_rows, scrub, d_rows, d_mm);
        hb_sum_rows_kernel<<<1, 1024, 0, s>>>(d_rows, h->n_rows, d_rows + n);
        g_hb_launches.fetch_add(2);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_rows + n, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    hb_dfree(d_rows);
    if (e != cudaSuccess) return hb_cuda_fail(e, "value statistics", __FILE__, __LINE__);
    auto unord = [](unsigned int k) {
        const unsigned int b = (k >> 31) ? (k & 0x7fffffffu) : ~k;
        float f;
        memcpy(&f, &b, sizeof(f));
        return (double)f;
    };
    if (sum_out) *sum_out = total;
    // an empty store has no minimum / maximum: NaN (np.min raises there; the caller decides)
    if (min_out) *min_out = mm[0] == 0xffffffffu && mm[1] == 0u ? NAN : unord(mm[0]);
    if (max_out) *max_out = mm[0] == 0xffffffffu && mm[1] == 0u ? NAN : unord(mm[1]);
    return HB_OK;
}

int hb_csr_normalize(hb_csr* h, int mode, double a, double b, double threshold, int scrub_first) {
    HB_RANGE("hb_csr_normalize");
    HB_REQUIRE(h != nullptr && mode >= 0 && mode <= 3, "bad arguments");
    int rc = hb_use_device(h);
    if (rc) return rc;
    hb_drop_pack(h);
    if (h->nnz > 0) {
        hb_normalize_kernel<<<HB_NUM_SMS * 8, 256, 0, h->stream>>>(h->d_data + h->phase, h->nnz, mode, (float)a, (float)b,
                                                                   (float)threshold, scrub_first);
        HB_LAUNCH_CHECK();
    }
    HB_CUDA(cudaStreamSynchronize(h->stream));
    return HB_OK;
}

int hb_csr_eliminate_zeros(hb_csr* h) {
    HB_RANGE("hb_csr_eliminate_zeros");
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(h->phase == 0, "row blocks with an alignment phase are not supported here");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);
    const int64_t n = h->n_rows;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    int64_t *d_cnt = nullptr, *d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    int64_t new_nnz = 0;
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 8;
#define EZ(expr)                                                  \
    do {                                                          \
        e = (expr);                                               \
        if (e != cudaSuccess) {                                   \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);      \
            goto fail;                                            \
        }                                                         \
    } while (0)
    EZ(hb_dmalloc(&d_cnt, sizeof(int64_t) * (nn + 1)));
    EZ(hb_dmalloc(&d_ptr, sizeof(int64_t) * (nn + 1)));
    EZ(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (nn + 1), s));
    if (n > 0) {
        hb_nonzero_kernel<false><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, n, d_cnt, nullptr, nullptr, nullptr);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n);
    g_hb_launches.fetch_add(1);
    EZ(cudaMemcpyAsync(&new_nnz, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    EZ(cudaStreamSynchronize(s));
    EZ(cudaGetLastError());
    if (new_nnz == h->nnz) {   // nothing to drop
        hb_dfree(d_cnt);
        hb_dfree(d_ptr);
        return HB_OK;
    }
    EZ(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    EZ(hb_dmalloc(&d_val, sizeof(double) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    if (n > 0 && new_nnz > 0) {
        hb_nonzero_kernel<true><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, n, nullptr, d_ptr, d_idx, d_val);
        g_hb_launches.fetch_add(1);
    }
    EZ(cudaStreamSynchronize(s));
    EZ(cudaGetLastError());
#undef EZ
    hb_dfree(d_cnt);
    if (h->owns_csr) {
        hb_dfree(h->d_indptr);
        hb_dfree(h->d_indices);
        hb_dfree(h->d_data);
    }
    h->owns_csr = true;
    h->d_indptr = d_ptr;
    h->d_indices = d_idx;
    h->d_data = d_val;
    h->nnz = new_nnz;
    return HB_OK;
fail:
    hb_dfree(d_cnt);
    hb_dfree(d_ptr);
    hb_dfree(d_idx);
    hb_dfree(d_val);
    return rc;
}

}  // extern "C"
