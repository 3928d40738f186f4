// The two optional ICE-only steps around the correction:
//   --transCutoff      hicCorrectMatrix.py:695-699 -> hicmatrix `truncTrans(high)`: the `100 - high` percentile of the
//                      values stored BETWEEN chromosomes (np.percentile, linear interpolation) and the clipping of every
//                      trans value at or above it.  hb_trans_order_stats returns the exact order statistics the
//                      percentile interpolates between (8-bit radix select over the stored entries, no sort, no copy
//                      of the values); hb_trans_clip applies the clipping.
//   --inflationCutoff  hicCorrectMatrix.py:765-775: row sums of the CORRECTED matrix divided by the row sums before the
//                      correction.  hb_ice_scaled_row_sums computes the former straight from the raw store and the ICE
//                      state (the corrected matrix is never materialised for it).
#include "hb_internal.cuh"

#define HB_MAX_SLOTS 64

struct HbTransSel {
    unsigned long long prefix;  // decided high bits of the ordered key
    unsigned long long mask;    // which bits are decided
    unsigned long long rank;    // remaining 0-based rank inside the candidate set
    unsigned long long n_le;    // final pass: trans entries with key <= selected key
    unsigned long long min_gt;  // final pass: smallest key above the selected key (all ones = none)
    unsigned long long n_trans; // number of trans entries
};

// monotone map double -> uint64 (negative values below positive ones, -0.0 < +0.0 adjacent)
__device__ __forceinline__ unsigned long long hb_ordkey(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ void hb_chr_range(const int64_t* __restrict__ off, int nchr, int64_t g, int64_t& lo_col, int64_t& hi_col) {
    int lo = 0, hi = nchr;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    lo_col = off[lo];
    hi_col = off[lo + 1];
}

// MODE 0: histogram of the digit at `shift` over the trans entries whose decided bits match the prefix
// MODE 1: n_le / min_gt relative to the selected key (prefix with all bits decided)
// MODE 2: clip trans values >= limit to limit
template <int MODE>
__global__ void __launch_bounds__(256)
hb_trans_pass_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, double* __restrict__ val, int64_t n_rows,
                     int64_t row0, const int64_t* __restrict__ chr_off, int nchr, HbTransSel* st, int shift,
                     unsigned long long* __restrict__ hist, double limit, unsigned long long* n_clipped) {
    __shared__ unsigned int sh[256];
    if (MODE == 0) {
        sh[threadIdx.x] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned long long prefix = st->prefix, mask = st->mask;
    unsigned long long cnt = 0, mn = ~0ull;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        int64_t clo, chi;
        hb_chr_range(chr_off, nchr, row0 + row, clo, chi);
        const int64_t e = indptr[row + 1];
        for (int64_t k0 = indptr[row]; k0 < e; k0 += 32 * 4) {
            int c[4];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + lane + 32 * j;
                c[j] = k < e ? idx[k] : -1;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + lane + 32 * j;
                v[j] = (c[j] >= 0 && (c[j] < clo || c[j] >= chi)) ? val[k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool trans = c[j] >= 0 && (c[j] < clo || c[j] >= chi);
                if (MODE == 0) {
                    const unsigned long long key = hb_ordkey(v[j]);
                    const bool cand = trans && (key & mask) == prefix;
                    const unsigned int digit = (unsigned int)((key >> shift) & 0xffu);
                    // Hi-C counts share their high bytes: aggregate equal digits inside the warp before the atomic
                    const unsigned int peers = __match_any_sync(0xffffffffu, cand ? digit : 0x100u);
                    if (cand && lane == __ffs(peers) - 1) atomicAdd(&sh[digit], (unsigned int)__popc(peers));
                } else if (MODE == 1) {
                    if (trans) {
                        const unsigned long long key = hb_ordkey(v[j]);
                        if (key <= prefix) ++cnt;
                        else if (key < mn) mn = key;
                    }
                } else {
                    if (trans && v[j] >= limit) {
                        val[k0 + lane + 32 * j] = limit;
                        ++cnt;
                    }
                }
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
    } else {
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o);
            mn = t < mn ? t : mn;
        }
        if (lane == 0) {
            if (MODE == 1) {
                if (cnt) atomicAdd(&st->n_le, cnt);
                if (mn != ~0ull) atomicMin(&st->min_gt, mn);
            } else if (cnt) {
                atomicAdd(n_clipped, cnt);
            }
        }
    }
}

// one block: choose the digit that holds the wanted rank, narrow the candidate set, clear the histogram
__global__ void hb_trans_pick_kernel(HbTransSel* st, unsigned long long* hist, int shift, int first) {
    __shared__ unsigned long long h[256];
    h[threadIdx.x] = hist[threadIdx.x];
    hist[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long total = 0;
        for (int d = 0; d < 256; ++d) total += h[d];
        if (first) st->n_trans = total;
        unsigned long long r = st->rank;
        if (total == 0) return;
        if (r >= total) r = total - 1;
        int d = 0;
        while (d < 255 && r >= h[d]) {
            r -= h[d];
            ++d;
        }
        st->rank = r;
        st->prefix |= (unsigned long long)d << shift;
        st->mask |= 0xffull << shift;
    }
}

// sharded runs: the path's one collective sums doubles; counts below 2^53 travel exactly
__global__ void hb_u64_to_f64_kernel(const unsigned long long* __restrict__ in, double* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void hb_f64_to_u64_kernel(const double* __restrict__ in, unsigned long long* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)in[i];
}

// row sums of the corrected matrix, same per-entry arithmetic as the emission (hb_ice_emit_kernel)
__global__ void __launch_bounds__(256)
hb_scaled_row_sums_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t n_rows, int64_t row0, const double* __restrict__ r, const int64_t* __restrict__ seg_off,
                          const double* __restrict__ seg_corr, int nseg, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int lo = 0, hi = nseg;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= g) lo = mid; else hi = mid;
        }
        const double corr = seg_corr[lo];
        const double ri = r[g];
        double acc = 0.0;
        const int64_t e = indptr[row + 1];
        for (int64_t k = indptr[row] + lane; k < e; k += 32) {
            const double w = (hb_ldg_stream_d1(val + k) * ri) * __ldg(r + hb_ldg_stream_i1(idx + k));
            acc += w * corr * corr;
        }
        acc = hb_warp_sum(acc);
        if (lane == 0) out[row] = acc;
    }
}

extern "C" {

int hb_trans_order_stats(hb_csr* h, const int64_t* chr_offsets, int nchr, int64_t k, int64_t* n_trans_out, double* v_k,
                         double* v_k1) {
    HB_RANGE("hb_trans_order_stats");
    HB_REQUIRE(h != nullptr && chr_offsets != nullptr && nchr >= 1 && n_trans_out != nullptr, "bad arguments");
    const bool sharded = h->world > 1;
    HB_REQUIRE(sharded || (h->row0 == 0 && h->n_rows == h->n_cols),
               "hb_trans_order_stats needs the full matrix on one device unless hb_set_comm was called");
    HB_REQUIRE(!sharded || h->allreduce != nullptr, "sharded hb_trans_order_stats needs the allreduce callback (hb_set_comm)");
    HB_REQUIRE(chr_offsets[0] == 0 && chr_offsets[nchr] == h->n_cols, "chromosome offsets must span [0, n]");
    for (int c = 0; c < nchr; ++c) HB_REQUIRE(chr_offsets[c] <= chr_offsets[c + 1], "chromosome offsets must ascend");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    int64_t* d_off = nullptr;
    HbTransSel* d_st = nullptr;
    unsigned long long* d_hist = nullptr;
    double* d_f = nullptr;   // sharded: histogram / final slots as doubles for the collective
    HbTransSel st = {0, 0, (unsigned long long)(k > 0 ? k : 0), 0, ~0ull, 0};
    const bool want = v_k != nullptr || v_k1 != nullptr;
    HbScratch scratch;
    HB_CUDA(scratch.alloc(&d_off, sizeof(int64_t) * (size_t)(nchr + 1)));
    HB_CUDA(scratch.alloc(&d_st, sizeof(HbTransSel)));
    HB_CUDA(scratch.alloc(&d_hist, sizeof(unsigned long long) * 256));
    HB_CUDA(scratch.alloc(&d_f, sizeof(double) * 512));
    HB_CUDA(cudaMemcpyAsync(d_off, chr_offsets, sizeof(int64_t) * (size_t)(nchr + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(d_st, &st, sizeof(st), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 256, s));
    const int grid = HB_NUM_SMS * 8;
    // the first pass also counts the trans entries; a pure count (v_k == v_k1 == NULL) stops after it
    for (int pass = 0; pass < (want ? 8 : 1); ++pass) {
        const int shift = 56 - 8 * pass;
        if (h->n_rows > 0) {
            hb_trans_pass_kernel<0><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, d_off, nchr, d_st,
                                                         shift, d_hist, 0.0, nullptr);
            g_hb_launches.fetch_add(1);
        }
        if (sharded) {
            // every rank holds the histogram of its rows: sum them, then all ranks pick the same digit
            hb_u64_to_f64_kernel<<<1, 256, 0, s>>>(d_hist, d_f, 256);
            if (h->allreduce(h->comm_ctx, d_f, 256, (void*)s) != 0) {
                hb_set_error("collective callback failed");
                return HB_ERR_CUDA;
            }
            hb_f64_to_u64_kernel<<<1, 256, 0, s>>>(d_f, d_hist, 256);
            g_hb_launches.fetch_add(2);
        }
        hb_trans_pick_kernel<<<1, 256, 0, s>>>(d_st, d_hist, shift, pass == 0);
        g_hb_launches.fetch_add(1);
    }
    if (want && h->n_rows > 0) {
        hb_trans_pass_kernel<1><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, d_off, nchr, d_st, 0,
                                                     d_hist, 0.0, nullptr);
        g_hb_launches.fetch_add(1);
    }
    HB_CUDA(cudaMemcpyAsync(&st, d_st, sizeof(st), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    HB_CUDA(cudaGetLastError());
    if (want && sharded) {
        // n_le adds up; min_gt is the smallest over the ranks: each rank ships (n_le, four 16-bit limbs of min_gt) in its
        // own slot of a zeroed vector, one sum-allreduce assembles all slots
        double slots[HB_MAX_SLOTS * 5];
        HB_REQUIRE(h->world <= HB_MAX_SLOTS, "too many ranks");
        for (int i = 0; i < h->world * 5; ++i) slots[i] = 0.0;
        slots[h->rank * 5 + 0] = (double)st.n_le;
        for (int l = 0; l < 4; ++l) slots[h->rank * 5 + 1 + l] = (double)((st.min_gt >> (16 * l)) & 0xffffull);
        HB_CUDA(cudaMemcpyAsync(d_f, slots, sizeof(double) * (size_t)(h->world * 5), cudaMemcpyHostToDevice, s));
        if (h->allreduce(h->comm_ctx, d_f, h->world * 5, (void*)s) != 0) {
            hb_set_error("collective callback failed");
            return HB_ERR_CUDA;
        }
        HB_CUDA(cudaMemcpyAsync(slots, d_f, sizeof(double) * (size_t)(h->world * 5), cudaMemcpyDeviceToHost, s));
        HB_CUDA(cudaStreamSynchronize(s));
        st.n_le = 0;
        st.min_gt = ~0ull;
        for (int q = 0; q < h->world; ++q) {
            st.n_le += (unsigned long long)slots[q * 5];
            unsigned long long m = 0;
            for (int l = 0; l < 4; ++l) m |= ((unsigned long long)slots[q * 5 + 1 + l]) << (16 * l);
            if (m < st.min_gt) st.min_gt = m;
        }
    }
    *n_trans_out = (int64_t)st.n_trans;
    if (want) {
        HB_REQUIRE(k >= 0 && (unsigned long long)k < st.n_trans, "rank outside the trans entries");
        union { unsigned long long u; double d; } a, b;
        // invert the ordered key on the host (same map as hb_ordkey_inv)
        a.u = (st.prefix >> 63) ? (st.prefix & 0x7fffffffffffffffull) : ~st.prefix;
        // the (k+1)-th value equals the k-th when more than k+1 entries are <= it, else it is the next larger one
        const unsigned long long nxt = (st.n_le >= (unsigned long long)k + 2 || st.min_gt == ~0ull) ? st.prefix : st.min_gt;
        b.u = (nxt >> 63) ? (nxt & 0x7fffffffffffffffull) : ~nxt;
        if (v_k) *v_k = a.d;
        if (v_k1) *v_k1 = b.d;
    }
    return HB_OK;
}

int hb_trans_clip(hb_csr* h, const int64_t* chr_offsets, int nchr, double max_inter, int64_t* n_clipped_out) {
    HB_RANGE("hb_trans_clip");
    HB_REQUIRE(h != nullptr && chr_offsets != nullptr && nchr >= 1, "bad arguments");
    // row-local: on a row block the count is that of the block's rows
    HB_REQUIRE(chr_offsets[0] == 0 && chr_offsets[nchr] == h->n_cols, "chromosome offsets must span [0, n]");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);  // the values change
    int64_t* d_off = nullptr;
    HbTransSel* d_st = nullptr;
    unsigned long long* d_cnt = nullptr;
    unsigned long long cnt = 0;
    HbScratch scratch;
    HB_CUDA(scratch.alloc(&d_off, sizeof(int64_t) * (size_t)(nchr + 1)));
    HB_CUDA(scratch.alloc(&d_st, sizeof(HbTransSel)));
    HB_CUDA(scratch.alloc(&d_cnt, sizeof(unsigned long long)));
    HB_CUDA(cudaMemcpyAsync(d_off, chr_offsets, sizeof(int64_t) * (size_t)(nchr + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemsetAsync(d_st, 0, sizeof(HbTransSel), s));
    HB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
    if (h->n_rows > 0) {
        hb_trans_pass_kernel<2><<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, d_off, nchr,
                                                               d_st, 0, nullptr, max_inter, d_cnt);
        g_hb_launches.fetch_add(1);
    }
    HB_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    HB_CUDA(cudaGetLastError());
    if (n_clipped_out) *n_clipped_out = (int64_t)cnt;
    return HB_OK;
}

int hb_ice_scaled_row_sums(hb_csr* h, double* out) {
    HB_RANGE("hb_ice_scaled_row_sums");
    HB_REQUIRE(h != nullptr && h->state_ready && out != nullptr, "run ICE first");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    if (h->n_rows == 0) return HB_OK;
    double* d_out = nullptr;
    HB_CUDA(hb_dmalloc(&d_out, sizeof(double) * (size_t)h->n_rows));
    hb_scaled_row_sums_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, h->d_r,
                                                             h->d_seg_off, h->d_seg_corr, h->nseg, d_out);
    g_hb_launches.fetch_add(1);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * (size_t)h->n_rows, cudaMemcpyDeviceToHost);
    hb_dfree(d_out);
    if (e != cudaSuccess) return hb_cuda_fail(e, "scaled row sums", __FILE__, __LINE__);
    return HB_OK;
}

}  // extern "C"
