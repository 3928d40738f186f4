// The two ends of the file path on the device:
//   hb_csr_create_upper   what hicmatrix does when it loads a file: the file stores the UPPER triangle, the matrix in
//                         memory is full symmetric (`M + triu(M, 1).T`; call site hicCorrectMatrix.py:603-609).  Only
//                         the upper triangle crosses PCIe (half the bytes of hb_csr_create); the mirror image is built
//                         in HBM.
//   hb_ice_scaled_upper   what hicmatrix does when it saves: restoreMaskedBins (scatter the kept bins back into the
//                         original frame) + upper triangle + eliminate_zeros (hicCorrectMatrix.py:779), fused with the
//                         final ICE scaling (iterativeCorrection.py:56-57, 79) -- the corrected matrix leaves the
//                         device already in file layout, 12 B per UPPER entry.
//
// Symmetrisation = sparse transpose of the strict upper triangle.  Row i of the result is [entries (j, U_ji), j < i,
// ascending j] ++ [row i of U].  The lower parts are filled through per-row atomic cursors (arrival order is
// arbitrary) and then each row's lower part is sorted by column in shared memory (bitonic network, keys are unique),
// so the final layout -- and with it every later floating-point sum -- is deterministic.
#include <climits>
#include <cstdlib>

#include "hb_internal.cuh"
#include "hb_scan.cuh"

__global__ void __launch_bounds__(256)
hb_sym_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t n,
                    unsigned int* __restrict__ low_cnt, unsigned int* flags /* [0] = an entry below the diagonal */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t row = gwarp; row < n; row += nwarps) {
        const int64_t e = indptr[row + 1];
        for (int64_t k = indptr[row] + lane; k < e; k += 32) {
            const int c = idx[k];
            if (c < row) bad = true;
            else if (c > row) atomicAdd(low_cnt + c, 1u);
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1u);
}

__global__ void hb_sym_total_kernel(const int64_t* __restrict__ indptr, const unsigned int* __restrict__ low_cnt, int64_t n,
                                    int64_t* __restrict__ total, unsigned int* max_low) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned int m = 0;
    if (i < n) {
        total[i] = (int64_t)low_cnt[i] + (indptr[i + 1] - indptr[i]);
        m = low_cnt[i];
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_low, m);
}

__global__ void __launch_bounds__(256)
hb_sym_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                   int64_t n, const int64_t* __restrict__ out_ptr, const unsigned int* __restrict__ low_cnt,
                   unsigned int* __restrict__ cursor, int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n; row += nwarps) {
        const int64_t s = indptr[row], e = indptr[row + 1];
        const int64_t up0 = out_ptr[row] + low_cnt[row];
        for (int64_t k = s + lane; k < e; k += 32) {
            const int c = idx[k];
            const double v = val[k];
            out_idx[up0 + (k - s)] = c;        // the row's own (upper) entries, in order
            out_val[up0 + (k - s)] = v;
            if (c > row) {                      // mirror image into row c, any order for now
                const int64_t pos = out_ptr[c] + atomicAdd(cursor + c, 1u);
                out_idx[pos] = (int32_t)row;
                out_val[pos] = v;
            }
        }
    }
}

// one CTA per row whose lower part has (P/2, P] entries: bitonic sort of (column id << 32 | arrival position) in shared
// memory -- one 64-bit word per entry, half a compare-exchange per thread and step -- then the values follow their keys
// through registers (read all, barrier, write all)
#define HB_SYM_MAX_PER_THREAD 16
__global__ void hb_sym_sort_kernel(const int64_t* __restrict__ out_ptr, const unsigned int* __restrict__ low_cnt, int64_t n,
                                   int P, int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    extern __shared__ unsigned long long sk[];
    const int64_t row = blockIdx.x;
    const int L = (int)low_cnt[row];
    if (L < 2 || L > P || (P > 32 && L <= P / 2)) return;
    const int64_t base = out_ptr[row];
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        sk[i] = i < L ? (((unsigned long long)(unsigned int)out_idx[base + i] << 32) | (unsigned int)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1, sj = 31 - __clz(k >> 1); j > 0; j >>= 1, --sj) {
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int lo = ((i >> sj) << (sj + 1)) | (i & (j - 1));     // element whose bit j is clear (j = 2^sj)
                const int hi = lo + j;
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == ((lo & k) == 0)) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    double v[HB_SYM_MAX_PER_THREAD];
#pragma unroll
    for (int q = 0; q < HB_SYM_MAX_PER_THREAD; ++q) {
        const int i = threadIdx.x + q * blockDim.x;
        if (i < L) v[q] = out_val[base + (unsigned int)(sk[i] & 0xffffffffull)];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < HB_SYM_MAX_PER_THREAD; ++q) {
        const int i = threadIdx.x + q * blockDim.x;
        if (i < L) {
            out_idx[base + i] = (int32_t)(sk[i] >> 32);
            out_val[base + i] = v[q];
        }
    }
}

// [r2] The same job without a sorting network.  The keys of a row's lower part are distinct source-row numbers below the
// row's own, so the sorted position of an entry is its RANK = number of smaller keys = number of set bits below its own
// in a bitmap of the keys.  One CTA per row; every thread takes E entries (key and value) into registers, then the key
// range is walked in windows of HB_SYM_WINDOW bits: clear, set one bit per key, popcount prefix over the words (block
// scan), and every entry of the window is written straight to base + rank -- in place, since everything was read before
// the first write.  Windows without keys (the gap between the band and a far contact) cost one barrier.  O(L + span/32)
// per row instead of O(L log^2 L) compare-exchanges behind 78 barriers.
#define HB_SYM_WINDOW 65536              // bits per window: 8 KB of bitmap + 8 KB of word prefixes
#define HB_SYM_WWORDS (HB_SYM_WINDOW / 32)
template <int E, int THREADS>
__global__ void __launch_bounds__(THREADS)
hb_sym_rank_kernel(const int64_t* __restrict__ out_ptr, const unsigned int* __restrict__ low_cnt, int64_t n, int P,
                   int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    __shared__ unsigned int bm[HB_SYM_WWORDS];
    __shared__ unsigned int pre[HB_SYM_WWORDS];
    __shared__ unsigned int warp_tot[32];
    __shared__ int s_min, s_max;
    const int64_t row = blockIdx.x;
    const int L = (int)low_cnt[row];
    if (L < 2 || L > P || (P > 32 && L <= P / 2)) return;
    const int64_t base = out_ptr[row];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = THREADS >> 5;
    int key[E];
    double val[E];
    int kmin = INT_MAX, kmax = -1;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int i = tid + q * THREADS;
        key[q] = -1;
        val[q] = 0.0;
        if (i < L) {
            key[q] = out_idx[base + i];
            val[q] = out_val[base + i];
            kmin = min(kmin, key[q]);
            kmax = max(kmax, key[q]);
        }
    }
    if (tid == 0) {
        s_min = INT_MAX;
        s_max = -1;
    }
    __syncthreads();
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
        atomicMin(&s_min, kmin);
        atomicMax(&s_max, kmax);
    }
    __syncthreads();       // also: every entry of the row is in registers from here on
    kmin = s_min;
    kmax = s_max;
    unsigned int running = 0;
    for (int w0 = kmin & ~(HB_SYM_WINDOW - 1); w0 <= kmax; w0 += HB_SYM_WINDOW) {
        bool mine = false;
#pragma unroll
        for (int q = 0; q < E; ++q) mine |= key[q] >= w0 && key[q] - w0 < HB_SYM_WINDOW;
        if (!__syncthreads_or(mine)) continue;
        const int last = min(kmax - w0, HB_SYM_WINDOW - 1);
        const int nwords = (last >> 5) + 1;
        for (int w = tid; w < nwords; w += THREADS) bm[w] = 0u;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int d = key[q] - w0;
            if (key[q] >= w0 && d < HB_SYM_WINDOW) atomicOr(&bm[d >> 5], 1u << (d & 31));
        }
        __syncthreads();
        // exclusive prefix of the word popcounts: a contiguous slice of words per thread, then a block scan of the slices
        const int per = (nwords + THREADS - 1) / THREADS;
        const int wb = tid * per, we = min(wb + per, nwords);
        unsigned int mysum = 0;
        for (int w = wb; w < we; ++w) mysum += __popc(bm[w]);
        unsigned int incl = mysum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned int t = lane < nwarp ? warp_tot[lane] : 0u;
            unsigned int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            warp_tot[lane] = ti - t;                       // exclusive offsets of the warps
            if (lane == 31) pre[HB_SYM_WWORDS - 1] = ti;    // parked: the window's total (read before it is overwritten)
        }
        __syncthreads();
        const unsigned int window_total = pre[HB_SYM_WWORDS - 1];
        unsigned int acc = warp_tot[warp] + incl - mysum;
        __syncthreads();       // everyone holds window_total before the last word's prefix may replace it
        for (int w = wb; w < we; ++w) {
            pre[w] = acc;
            acc += __popc(bm[w]);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int d = key[q] - w0;
            if (key[q] >= w0 && d < HB_SYM_WINDOW) {
                const unsigned int r = running + pre[d >> 5] + __popc(bm[d >> 5] & ((1u << (d & 31)) - 1u));
                out_idx[base + r] = key[q];
                out_val[base + r] = val[q];
            }
        }
        running += window_total;
        __syncthreads();       // bm / pre are cleared again by the next window
    }
}

// rows whose lower part does not fit the shared-memory classes (more than 16384 entries: the dense rows of a coarse
// genome-wide matrix): the same network on a global scratch slice (L2 resident), one CTA of 1024 threads per row
__global__ void __launch_bounds__(1024)
hb_sym_sort_big_kernel(const int64_t* __restrict__ out_ptr, const unsigned int* __restrict__ low_cnt,
                       const int64_t* __restrict__ big_rows, const int64_t* __restrict__ big_off, unsigned long long* __restrict__ keys,
                       double* __restrict__ vals, int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    const int64_t row = big_rows[blockIdx.x];
    const int64_t L = (int64_t)low_cnt[row];
    const int64_t P = big_off[blockIdx.x + 1] - big_off[blockIdx.x];
    unsigned long long* sk = keys + big_off[blockIdx.x];
    double* sv = vals + big_off[blockIdx.x];
    const int64_t base = out_ptr[row];
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        sk[i] = i < L ? (((unsigned long long)(unsigned int)out_idx[base + i] << 32) | (unsigned long long)i) : ~0ull;
        if (i < L) sv[i] = out_val[base + i];
    }
    __syncthreads();
    for (int64_t k = 2; k <= P; k <<= 1) {
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            for (int64_t i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int64_t lo = ((i / j) * j << 1) + (i % j);
                const int64_t hi = lo + j;
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == ((lo & k) == 0)) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        out_idx[base + i] = (int32_t)(sk[i] >> 32);
        out_val[base + i] = sv[sk[i] & 0xffffffffull];
    }
}

// ---- corrected matrix in file layout -------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
hb_scaled_upper_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                       int64_t n_rows, int64_t row0, const double* __restrict__ r, const int64_t* __restrict__ seg_off,
                       const double* __restrict__ seg_corr, int nseg, const int64_t* __restrict__ orig,
                       int64_t* __restrict__ cnt_orig, const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_idx,
                       double* __restrict__ out_val, unsigned long long* max_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double m0 = 0.0, m1 = 0.0;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int lo = 0, hi = nseg;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= g) lo = mid; else hi = mid;
        }
        const double corr = seg_corr[lo];
        const double ri = r[g];
        const int64_t og = orig[g];
        if (og < 0) continue;                    // bin dropped from the output (--inflationCutoff)
        int64_t w = FILL ? out_ptr[og] : 0;
        long long n_out = 0;
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t base = s; base < e; base += 32) {
            const int64_t k = base + lane;
            int c = -1;
            double f = 0.0, wv = 0.0;
            if (k < e) {
                c = idx[k];
                wv = (val[k] * ri) * r[c];      // iterativeCorrection.py:56-57
                f = wv * corr * corr;           // :79
                m0 = fmax(m0, wv);
                m1 = fmax(m1, f);
            }
            const int64_t oc = (c >= g && f != 0.0) ? orig[c] : -1;  // upper triangle, explicit zeros eliminated
            const bool kp = oc >= 0;
            const unsigned int bal = __ballot_sync(0xffffffffu, kp);
            if (FILL && kp) {
                const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                out_idx[pos] = (int32_t)oc;
                out_val[pos] = f;
            }
            w += __popc(bal);
            n_out += __popc(bal);
        }
        if (!FILL && lane == 0) cnt_orig[og] = n_out;
    }
    if (FILL) {
        m0 = hb_warp_max(m0);
        m1 = hb_warp_max(m1);
        if (lane == 0) {
            if (m0 > 0.0) atomicMax(max_bits + 0, (unsigned long long)__double_as_longlong(m0));
            if (m1 > 0.0) atomicMax(max_bits + 1, (unsigned long long)__double_as_longlong(m1));
        }
    }
}

extern "C" {

int hb_csr_create_upper(hb_csr** out, int64_t n, int64_t nnz_upper, const int64_t* indptr, const int32_t* indices,
                        const double* data, int device) {
    return hb_csr_create_upper_typed(out, n, nnz_upper, indptr, indices, data, HB_VALUES_F64, device);
}

int hb_csr_create_upper_typed(hb_csr** out, int64_t n, int64_t nnz_upper, const int64_t* indptr, const int32_t* indices,
                              const void* data, int value_kind, int device) {
    HB_RANGE("hb_csr_create_upper_typed");
    HB_REQUIRE(out != nullptr && indptr != nullptr, "bad arguments");
    // upload the triangle with the ordinary constructor, then replace the store by its symmetric completion
    int rc = hb_csr_create_typed(out, n, n, 0, nnz_upper, indptr, indices, 0, data, value_kind, device);
    if (rc != HB_OK) return rc;
    rc = hb_csr_symmetrize_upper(*out);
    if (rc != HB_OK) {
        hb_csr_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int hb_csr_symmetrize_upper(hb_csr* h) {
    HB_RANGE("hb_csr_symmetrize_upper");
    HB_REQUIRE(h != nullptr, "null handle");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_csr_symmetrize_upper needs the full (upper-triangular) matrix");
    int rc = hb_use_device(h);
    if (rc) return rc;
    hb_drop_pack(h);
    const int64_t n = h->n_cols;
    cudaStream_t s = h->stream;
    unsigned int *d_low = nullptr, *d_cursor = nullptr, *d_flags = nullptr;
    int64_t *d_total = nullptr, *d_ptr = nullptr, *d_big_rows = nullptr, *d_big_off = nullptr;
    unsigned long long* d_big_keys = nullptr;
    double* d_big_vals = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    int64_t full_nnz = 0;
    unsigned int flags[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 8;
    const size_t nn = (size_t)(n > 0 ? n : 1);
#define SY(expr)                                                  \
    do {                                                          \
        e = (expr);                                               \
        if (e != cudaSuccess) {                                   \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);      \
            goto fail;                                            \
        }                                                         \
    } while (0)
    SY(hb_dmalloc(&d_low, sizeof(unsigned int) * nn));
    SY(hb_dmalloc(&d_cursor, sizeof(unsigned int) * nn));
    SY(hb_dmalloc(&d_flags, sizeof(unsigned int) * 2));
    SY(hb_dmalloc(&d_total, sizeof(int64_t) * (nn + 1)));
    SY(hb_dmalloc(&d_ptr, sizeof(int64_t) * (nn + 1)));
    SY(cudaMemsetAsync(d_low, 0, sizeof(unsigned int) * nn, s));
    SY(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned int) * nn, s));
    SY(cudaMemsetAsync(d_flags, 0, sizeof(unsigned int) * 2, s));
    if (n > 0) {
        hb_sym_count_kernel<<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, n, d_low, d_flags);
        g_hb_launches.fetch_add(1);
        hb_sym_total_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->d_indptr, d_low, n, d_total, d_flags + 1);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_total, d_ptr, n);
    g_hb_launches.fetch_add(1);
    SY(cudaMemcpyAsync(&full_nnz, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, s));
    SY(cudaStreamSynchronize(s));
    if (flags[0]) {
        hb_set_error("hb_csr_create_upper: the input has entries below the diagonal");
        rc = HB_ERR_INVALID_ARG;
        goto fail;
    }
    SY(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(full_nnz > 0 ? full_nnz : 1) + 16));
    SY(hb_dmalloc(&d_val, sizeof(double) * (size_t)(full_nnz > 0 ? full_nnz : 1) + 16));
    if (n > 0 && full_nnz > 0) {
        hb_sym_fill_kernel<<<grid, 256, 0, s>>>(h->d_indptr + 0, h->d_indices, h->d_data, n, d_ptr, d_low, d_cursor, d_idx, d_val);
        g_hb_launches.fetch_add(1);
        // HB_SYM_SORT=network selects the round-1 bitonic network (kept for A/B runs and as a cross-check in the tests)
        const char* sort_env = getenv("HB_SYM_SORT");
        const bool use_network = sort_env != nullptr && sort_env[0] == 'n';
        // size classes: P = 32 sorts lower parts of 2..32 entries, P > 32 those of (P/2, P] entries
        for (int P = 32; P <= 16384 && (P == 32 ? flags[1] >= 2u : flags[1] > (unsigned int)(P / 2)); P <<= 1) {
            const size_t smem = sizeof(unsigned long long) * (size_t)P;
            if (use_network && smem > 48 * 1024)
                SY(cudaFuncSetAttribute(hb_sym_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // P / threads <= HB_SYM_MAX_PER_THREAD (16384 / 1024) for every class
            const int threads = P >= 2048 ? 1024 : (P >= 64 ? P / 2 : 32);
            if (use_network) {
                hb_sym_sort_kernel<<<(unsigned)n, threads, smem, s>>>(d_ptr, d_low, n, P, d_idx, d_val);
            } else {
                // rank by bitmap: E entries per thread in registers
                switch (P) {
                    case 32: hb_sym_rank_kernel<1, 32><<<(unsigned)n, 32, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 64: hb_sym_rank_kernel<2, 32><<<(unsigned)n, 32, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 128: hb_sym_rank_kernel<2, 64><<<(unsigned)n, 64, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 256: hb_sym_rank_kernel<2, 128><<<(unsigned)n, 128, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 512: hb_sym_rank_kernel<2, 256><<<(unsigned)n, 256, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 1024: hb_sym_rank_kernel<4, 256><<<(unsigned)n, 256, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 2048: hb_sym_rank_kernel<8, 256><<<(unsigned)n, 256, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 4096: hb_sym_rank_kernel<8, 512><<<(unsigned)n, 512, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    case 8192: hb_sym_rank_kernel<8, 1024><<<(unsigned)n, 1024, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                    default: hb_sym_rank_kernel<16, 1024><<<(unsigned)n, 1024, 0, s>>>(d_ptr, d_low, n, P, d_idx, d_val); break;
                }
            }
            g_hb_launches.fetch_add(1);
        }
        if (flags[1] > 16384u) {
            // the few rows beyond the shared-memory classes: list them on the host, sort them on a global scratch
            std::vector<unsigned int> low((size_t)n);
            SY(cudaMemcpyAsync(low.data(), d_low, sizeof(unsigned int) * (size_t)n, cudaMemcpyDeviceToHost, s));
            SY(cudaStreamSynchronize(s));
            std::vector<int64_t> rows, off(1, 0);
            for (int64_t i = 0; i < n; ++i) {
                if (low[(size_t)i] <= 16384u) continue;
                int64_t P = 32768;
                while (P < (int64_t)low[(size_t)i]) P <<= 1;
                rows.push_back(i);
                off.push_back(off.back() + P);
            }
            SY(hb_dmalloc(&d_big_rows, sizeof(int64_t) * rows.size()));
            SY(hb_dmalloc(&d_big_off, sizeof(int64_t) * off.size()));
            SY(hb_dmalloc(&d_big_keys, sizeof(unsigned long long) * (size_t)off.back()));
            SY(hb_dmalloc(&d_big_vals, sizeof(double) * (size_t)off.back()));
            SY(cudaMemcpyAsync(d_big_rows, rows.data(), sizeof(int64_t) * rows.size(), cudaMemcpyHostToDevice, s));
            SY(cudaMemcpyAsync(d_big_off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, s));
            hb_sym_sort_big_kernel<<<(unsigned)rows.size(), 1024, 0, s>>>(d_ptr, d_low, d_big_rows, d_big_off, d_big_keys,
                                                                          d_big_vals, d_idx, d_val);
            g_hb_launches.fetch_add(1);
            SY(cudaStreamSynchronize(s));   // rows / off are host vectors of this scope
        }
    }
    SY(cudaStreamSynchronize(s));
    SY(cudaGetLastError());
#undef SY
    hb_dfree(d_low);
    hb_dfree(d_cursor);
    hb_dfree(d_flags);
    hb_dfree(d_total);
    hb_dfree(d_big_rows);
    hb_dfree(d_big_off);
    hb_dfree(d_big_keys);
    hb_dfree(d_big_vals);
    hb_dfree(h->d_indptr);
    hb_dfree(h->d_indices);
    hb_dfree(h->d_data);
    h->d_indptr = d_ptr;
    h->d_indices = d_idx;
    h->d_data = d_val;
    h->phase = 0;
    h->nnz = full_nnz;
    return HB_OK;
fail:
    hb_dfree(d_low);
    hb_dfree(d_cursor);
    hb_dfree(d_flags);
    hb_dfree(d_total);
    hb_dfree(d_big_rows);
    hb_dfree(d_big_off);
    hb_dfree(d_big_keys);
    hb_dfree(d_big_vals);
    hb_dfree(d_ptr);
    hb_dfree(d_idx);
    hb_dfree(d_val);
    return rc;
}

// ---- shifted copy: C[i + dr][j + dc] = A[i][j] for the targets inside the n x n frame (one term of the running-window
// merge of hicMergeMatrixBins.py:88-186, which sums (2h+1)^2 shifted copies of the upper triangle) ---------------------
__global__ void hb_shift_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t n, int dr,
                                      int dc, int64_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t out_row = gwarp; out_row < n; out_row += nwarps) {
        const int64_t src = out_row - dr;
        long long c = 0;
        if (src >= 0 && src < n) {
            for (int64_t k = indptr[src] + lane; k < indptr[src + 1]; k += 32) {
                const int64_t j = (int64_t)idx[k] + dc;
                c += (j >= 0 && j < n) ? 1 : 0;
            }
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[out_row] = c;
    }
}

__global__ void hb_shift_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                     const double* __restrict__ val, int64_t n, int dr, int dc, const int64_t* __restrict__ out_ptr,
                                     int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t out_row = gwarp; out_row < n; out_row += nwarps) {
        const int64_t src = out_row - dr;
        if (src < 0 || src >= n) continue;
        int64_t w = out_ptr[out_row];
        const int64_t s = indptr[src], e = indptr[src + 1];
        for (int64_t base = s; base < e; base += 32) {
            const int64_t k = base + lane;
            int64_t j = -1;
            double v = 0.0;
            if (k < e) {
                j = (int64_t)idx[k] + dc;
                v = val[k];
            }
            const bool kp = j >= 0 && j < n;      // columns stay ascending: a shift keeps the order
            const unsigned int bal = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                out_idx[pos] = (int32_t)j;
                out_val[pos] = v;
            }
            w += __popc(bal);
        }
    }
}

int hb_csr_shifted_copy(hb_csr** out, const hb_csr* src, int64_t dr, int64_t dc) {
    HB_REQUIRE(out != nullptr && src != nullptr, "bad arguments");
    HB_REQUIRE(src->row0 == 0 && src->n_rows == src->n_cols && src->phase == 0, "hb_csr_shifted_copy needs a full matrix");
    const int64_t n = src->n_cols;
    HB_REQUIRE(dr > -n && dr < n && dc > -n && dc < n, "shift outside the matrix");
    int rc = hb_use_device(src);
    if (rc) return rc;
    cudaStream_t s = src->stream;
    int64_t *d_cnt = nullptr, *d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    int64_t nnz = 0;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    HbScratch scratch;
    HB_CUDA(scratch.alloc(&d_cnt, sizeof(int64_t) * (nn + 1)));
    HB_CUDA(hb_dmalloc(&d_ptr, sizeof(int64_t) * (nn + 1)));
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (nn + 1), s);
    if (e == cudaSuccess && n > 0) {
        hb_shift_count_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(src->d_indptr, src->d_indices, n, (int)dr, (int)dc, d_cnt);
        g_hb_launches.fetch_add(1);
    }
    if (e == cudaSuccess) {
        hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n);
        g_hb_launches.fetch_add(1);
        e = cudaMemcpyAsync(&nnz, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1) + 16);
    if (e == cudaSuccess) e = hb_dmalloc(&d_val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1) + 16);
    if (e == cudaSuccess && n > 0 && nnz > 0) {
        hb_shift_fill_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(src->d_indptr, src->d_indices, src->d_data, n, (int)dr, (int)dc, d_ptr,
                                                            d_idx, d_val);
        g_hb_launches.fetch_add(1);
        e = cudaStreamSynchronize(s);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        hb_dfree(d_ptr);
        hb_dfree(d_idx);
        hb_dfree(d_val);
        return hb_cuda_fail(e, "shifted copy", __FILE__, __LINE__);
    }
    rc = hb_dev_csr_wrap(out, n, n, 0, nnz, d_ptr, d_idx, d_val, src->device);
    if (rc != HB_OK) {
        hb_dfree(d_ptr);
        hb_dfree(d_idx);
        hb_dfree(d_val);
        return rc;
    }
    (*out)->owns_csr = true;   // the new handle owns the arrays made here
    return HB_OK;
}

int hb_ice_scaled_upper(hb_csr* h, const int64_t* orig_ids, int64_t n_orig, int64_t* nnz_out, int64_t* indptr,
                        int32_t* indices, double* data) {
    HB_RANGE("hb_ice_scaled_upper");
    HB_REQUIRE(h != nullptr && h->state_ready && orig_ids != nullptr && nnz_out != nullptr, "run ICE first");
    HB_REQUIRE(n_orig >= h->n_cols && n_orig < 2147483647LL, "bad original frame size");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int64_t n = h->n_cols;
    for (int64_t i = 0, prev = -1; i < n; ++i) {
        if (orig_ids[i] < 0) continue;           // -1: the bin is left out of the output
        HB_REQUIRE(orig_ids[i] < n_orig && orig_ids[i] > prev,
                   "original ids must be strictly ascending and inside the original frame");
        prev = orig_ids[i];
    }
    int64_t *d_orig = nullptr, *d_cnt = nullptr, *d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    int64_t total = 0;
    double mx[2] = {0.0, 0.0};
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 8;
#define SU(expr)                                                  \
    do {                                                          \
        e = (expr);                                               \
        if (e != cudaSuccess) {                                   \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);      \
            goto done;                                            \
        }                                                         \
    } while (0)
    SU(hb_dmalloc(&d_orig, sizeof(int64_t) * (size_t)(n > 0 ? n : 1)));
    SU(hb_dmalloc(&d_cnt, sizeof(int64_t) * (size_t)(n_orig + 1)));
    SU(hb_dmalloc(&d_ptr, sizeof(int64_t) * (size_t)(n_orig + 1)));
    if (n > 0) SU(cudaMemcpyAsync(d_orig, orig_ids, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    SU(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (size_t)(n_orig + 1), s));
    if (h->n_rows > 0) {
        hb_scaled_upper_kernel<false><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, h->d_r,
                                                           h->d_seg_off, h->d_seg_corr, h->nseg, d_orig, d_cnt, nullptr,
                                                           nullptr, nullptr, nullptr);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n_orig);
    g_hb_launches.fetch_add(1);
    SU(cudaMemcpyAsync(&total, d_ptr + n_orig, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SU(cudaStreamSynchronize(s));
    *nnz_out = total;
    if (indptr != nullptr && indices != nullptr && data != nullptr) {
        SU(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
        SU(hb_dmalloc(&d_val, sizeof(double) * (size_t)(total > 0 ? total : 1)));
        SU(cudaMemsetAsync(h->d_scalar_bits, 0, sizeof(unsigned long long) * 2, s));
        if (h->n_rows > 0 && total > 0) {
            hb_scaled_upper_kernel<true><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, h->d_r,
                                                              h->d_seg_off, h->d_seg_corr, h->nseg, d_orig, nullptr, d_ptr,
                                                              d_idx, d_val, h->d_scalar_bits);
            g_hb_launches.fetch_add(1);
        }
        SU(cudaMemcpyAsync(mx, h->d_scalar_bits, sizeof(mx), cudaMemcpyDeviceToHost, s));
        SU(cudaStreamSynchronize(s));
        SU(cudaGetLastError());
        SU(cudaMemcpy(indptr, d_ptr, sizeof(int64_t) * (size_t)(n_orig + 1), cudaMemcpyDeviceToHost));
        if (total > 0) {
            rc = hb_copy_d2h(indices, d_idx, sizeof(int32_t) * (size_t)total, h->device);
            if (rc == HB_OK) rc = hb_copy_d2h(data, d_val, sizeof(double) * (size_t)total, h->device);
        }
        if (rc == HB_OK && mx[0] > 1e100) {
            hb_set_error("matrix correction is producing extremely large values (> 1e100)");
            rc = HB_ERR_OVERFLOW_ITER;
        } else if (rc == HB_OK && mx[1] > 1e10) {
            hb_set_error("matrix correction produced extremely large values (> 1e10)");
            rc = HB_ERR_OVERFLOW_FINAL;
        }
    }
#undef SU
done:
    hb_dfree(d_orig);
    hb_dfree(d_cnt);
    hb_dfree(d_ptr);
    hb_dfree(d_idx);
    hb_dfree(d_val);
    return rc;
}

}  // extern "C"
