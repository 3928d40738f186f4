// Pre-filter stage on the device CSR store: NaN/Inf scrub, row statistics, the MAD outlier mask,
// bin deletion (row + column compaction), cis-only restriction, diagonal removal, symmetry test.
// Reference: hicexplorer/hicCorrectMatrix.py:349-391 (MAD), :548-592 (filter_by_zscore),
// :612-626 (zero bins, scrub), :648-649 (skipDiagonal), :672 (maskBins);
// hicexplorer/iterativeCorrection.py:35 (symmetry).  Contract: include/hicbalance.h.
//
// Bit-exactness of the mask: every floating-point step below is a single correctly rounded IEEE
// operation issued in the reference's order (subtract, subtract, divide, multiply); medians are
// exact order statistics (radix select) combined as (a+b)/2 like numpy.  Row sums are formed in a
// different order than scipy's, which is exact -- hence bit-identical -- for integer counts below
// 2^53 (the raw matrices this filter is applied to); see DESIGN.md.
#include "hb_internal.cuh"
#include "hb_scan.cuh"

// ---------------------------------------------------------------------------------------------
// scrub
// ---------------------------------------------------------------------------------------------
__global__ void hb_scrub_kernel(double* __restrict__ data, int64_t nnz, unsigned long long* counts) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int nn = 0, ni = 0;
    for (; i < nnz; i += stride) {
        const double v = data[i];
        if (isnan(v)) {
            data[i] = 0.0;
            ++nn;
        } else if (isinf(v)) {
            data[i] = 0.0;
            ++ni;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nn) atomicAdd(counts + 0, (unsigned long long)nn);
        if (ni) atomicAdd(counts + 1, (unsigned long long)ni);
    }
}

// ---------------------------------------------------------------------------------------------
// row statistics: warp per row; sum over the row (optionally only columns of the row's own
// segment = the chromosome block of --perchr) and the diagonal entry
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int hb_find_seg(const int64_t* __restrict__ seg_off, int nseg, int64_t g) {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
hb_row_stats_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                    int64_t n_rows, int64_t row0, const int64_t* __restrict__ seg_off, int nseg, int cis_only,
                    double* __restrict__ row_sum, double* __restrict__ diag) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        int64_t clo = 0, chi = INT64_MAX;
        if (cis_only && nseg > 1) {
            const int s = hb_find_seg(seg_off, nseg, g);
            clo = seg_off[s];
            chi = seg_off[s + 1];
        }
        double acc = 0.0, dg = 0.0;
        const int64_t s = indptr[row], e = indptr[row + 1];
        int64_t k = s + lane;
        // 8 independent (column id, value) pairs in flight per lane: this pass is a pure stream over the matrix
        for (; k + 32 * 7 < e; k += 32 * 8) {
            int c[8];
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = hb_ldg_stream_i1(idx + k + 32 * j);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = hb_ldg_stream_d1(val + k + 32 * j);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (c[j] >= clo && c[j] < chi) acc += v[j];
                if (c[j] == g) dg += v[j];
            }
        }
        for (; k < e; k += 32) {
            const int c = idx[k];
            const double v = val[k];
            if (c >= clo && c < chi) acc += v;
            if (c == g) dg += v;
        }
        acc = hb_warp_sum(acc);
        dg = hb_warp_sum(dg);
        if (lane == 0) {
            if (row_sum) row_sum[row] = acc;
            if (diag) diag[row] = dg;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// exact per-segment order statistics of non-negative doubles by 8-bit radix select
// ---------------------------------------------------------------------------------------------
struct HbSelState {          // one per segment, device resident
    unsigned long long prefix;   // decided high bits of the key
    unsigned long long mask;     // which bits are decided
    long long rank;              // remaining 0-based rank inside the candidate set
    long long count;             // number of candidates (set by the first pass)
};

// mode 0: candidates are keys > 0 (np.median(points[points > 0])); mode 1: all keys
__device__ __forceinline__ bool hb_sel_candidate(double v, int mode) { return mode == 1 || v > 0.0; }

__global__ void __launch_bounds__(256)
hb_sel_hist_kernel(const double* __restrict__ keys, const int64_t* __restrict__ tile_row0,
                   const int32_t* __restrict__ tile_len, const int32_t* __restrict__ tile_seg,
                   const HbSelState* __restrict__ st, int shift, int mode, unsigned int* __restrict__ hist /*nseg*256*/) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x;
    const int seg = tile_seg[t];
    const unsigned long long prefix = st[seg].prefix, mask = st[seg].mask;
    const int64_t r0 = tile_row0[t];
    for (int i = threadIdx.x; i < tile_len[t]; i += blockDim.x) {
        const double v = keys[r0 + i];
        if (!hb_sel_candidate(v, mode)) continue;
        const unsigned long long k = (unsigned long long)__double_as_longlong(v);
        if ((k & mask) == prefix) atomicAdd(&sh[(k >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[seg * 256 + threadIdx.x], sh[threadIdx.x]);
}

// one thread block per segment: pick the bucket holding the wanted rank, narrow the prefix
__global__ void hb_sel_pick_kernel(HbSelState* st, unsigned int* hist, int shift, int first_pass, int which /*0 lower, 1 upper middle*/) {
    const int seg = blockIdx.x;
    if (threadIdx.x != 0) return;
    unsigned int* h = hist + seg * 256;
    HbSelState s = st[seg];
    if (first_pass) {
        long long cnt = 0;
        for (int b = 0; b < 256; ++b) cnt += h[b];
        s.count = cnt;
        // numpy: even count -> elements cnt/2-1 and cnt/2 ; odd -> cnt/2 twice
        s.rank = (cnt % 2 == 0 && which == 0) ? cnt / 2 - 1 : cnt / 2;
    }
    if (s.count > 0) {
        long long r = s.rank;
        int b = 0;
        for (; b < 255; ++b) {
            if (r < (long long)h[b]) break;
            r -= h[b];
        }
        s.rank = r;
        s.prefix |= ((unsigned long long)b) << shift;
        s.mask |= 0xffull << shift;
    }
    st[seg] = s;
    for (int b = 0; b < 256; ++b) h[b] = 0;
}

__global__ void hb_sel_init_kernel(HbSelState* st, int nseg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg) {
        st[i].prefix = 0;
        st[i].mask = 0;
        st[i].rank = 0;
        st[i].count = 0;
    }
}

// median[seg] = (lower + upper) / 2, NaN for an empty candidate set (np.median of an empty array)
__global__ void hb_sel_finish_kernel(const HbSelState* lo, const HbSelState* hi, double* median, int nseg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg) {
        if (lo[i].count <= 0) {
            median[i] = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const double a = __longlong_as_double((long long)lo[i].prefix);
            const double b = __longlong_as_double((long long)hi[i].prefix);
            median[i] = (lo[i].count % 2 == 0) ? __ddiv_rn(__dadd_rn(a, b), 2.0) : b;
        }
    }
}

// points = row_sum - diag
__global__ void hb_points_kernel(const double* __restrict__ row_sum, const double* __restrict__ diag, double* __restrict__ points, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) points[i] = __dsub_rn(row_sum[i], diag[i]);
}

// diff = points - median[seg] ; absdiff = |diff|
__global__ void __launch_bounds__(256)
hb_diff_kernel(const double* __restrict__ points, const int64_t* __restrict__ tile_row0, const int32_t* __restrict__ tile_len,
               const int32_t* __restrict__ tile_seg, const double* __restrict__ median, double* __restrict__ diff,
               double* __restrict__ absdiff) {
    const int t = blockIdx.x;
    const double med = median[tile_seg[t]];
    const int64_t r0 = tile_row0[t];
    for (int i = threadIdx.x; i < tile_len[t]; i += blockDim.x) {
        const double d = __dsub_rn(points[r0 + i], med);
        diff[r0 + i] = d;
        absdiff[r0 + i] = fabs(d);
    }
}

// z = 0.6745 * (diff / mad) ; mask = (z < lo) | (z > hi)
__global__ void __launch_bounds__(256)
hb_zscore_kernel(const double* __restrict__ diff, const int64_t* __restrict__ tile_row0, const int32_t* __restrict__ tile_len,
                 const int32_t* __restrict__ tile_seg, const double* __restrict__ mad, double lo, double hi,
                 double* __restrict__ z_out, uint8_t* __restrict__ mask) {
    const int t = blockIdx.x;
    const double m = mad[tile_seg[t]];
    const int64_t r0 = tile_row0[t];
    for (int i = threadIdx.x; i < tile_len[t]; i += blockDim.x) {
        const double z = __dmul_rn(0.6745, __ddiv_rn(diff[r0 + i], m));
        if (z_out) z_out[r0 + i] = z;
        mask[r0 + i] = (z < lo || z > hi) ? 1 : 0;
    }
}

static int hb_segment_median(hb_csr* h, const double* d_keys, int mode, HbSelState* d_st /*2*nseg*/, unsigned int* d_hist,
                             double* d_median) {
    cudaStream_t s = h->stream;
    const int nseg = h->nseg;
    for (int which = 0; which < 2; ++which) {
        HbSelState* st = d_st + which * nseg;
        hb_sel_init_kernel<<<(nseg + 255) / 256, 256, 0, s>>>(st, nseg);
        HB_LAUNCH_CHECK();
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            if (h->n_tiles > 0) {
                hb_sel_hist_kernel<<<h->n_tiles, 256, 0, s>>>(d_keys, h->d_tile_row0, h->d_tile_len, h->d_tile_seg, st,
                                                              shift, mode, d_hist);
                HB_LAUNCH_CHECK();
            }
            hb_sel_pick_kernel<<<nseg, 32, 0, s>>>(st, d_hist, shift, pass == 0, which);
            HB_LAUNCH_CHECK();
        }
    }
    hb_sel_finish_kernel<<<(nseg + 255) / 256, 256, 0, s>>>(d_st, d_st + nseg, d_median, nseg);
    HB_LAUNCH_CHECK();
    return HB_OK;
}


// ---------------------------------------------------------------------------------------------
// compaction: keep entry (row, col) iff  newid[row] >= 0 && newid[col] >= 0
//                                    && (!cis || seg(row) == seg(col)) && (!drop_diag || row != col)
// ---------------------------------------------------------------------------------------------
struct HbKeepRule {
    const int32_t* __restrict__ newid;   // n_cols, -1 = bin removed (may be NULL = identity)
    const int64_t* __restrict__ seg_off;
    int nseg;
    int cis;
    int drop_diag;  // 1: drop the diagonal; 2: drop everything below the diagonal (keep the upper triangle)
    __device__ __forceinline__ bool keep(int64_t grow, int64_t clo, int64_t chi, int c) const {
        if (newid != nullptr && newid[c] < 0) return false;
        if (cis && (c < clo || c >= chi)) return false;
        if (drop_diag == 1 && c == grow) return false;
        if (drop_diag == 2 && c < grow) return false;
        return true;
    }
};

__global__ void __launch_bounds__(256)
hb_compact_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t n_rows, int64_t row0,
                        HbKeepRule rule, int64_t* __restrict__ new_count /* indexed by NEW local row */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        const int64_t nr = rule.newid ? rule.newid[g] : g;
        if (nr < 0) continue;
        int64_t clo = 0, chi = INT64_MAX;
        if (rule.cis) {
            const int s = hb_find_seg(rule.seg_off, rule.nseg, g);
            clo = rule.seg_off[s];
            chi = rule.seg_off[s + 1];
        }
        long long cnt = 0;
        const int64_t e = indptr[row + 1];
        int64_t k = indptr[row] + lane;
        for (; k + 32 * 7 < e; k += 32 * 8) {  // 8 column ids (and their renumbering look-ups) in flight per lane
            int c[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = hb_ldg_stream_i1(idx + k + 32 * j);
#pragma unroll
            for (int j = 0; j < 8; ++j) cnt += rule.keep(g, clo, chi, c[j]) ? 1 : 0;
        }
        for (; k < e; k += 32) cnt += rule.keep(g, clo, chi, idx[k]) ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) new_count[nr] = cnt;
    }
}

__global__ void __launch_bounds__(256)
hb_compact_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                       int64_t n_rows, int64_t row0, HbKeepRule rule, int64_t new_row0,
                       const int64_t* __restrict__ new_indptr, int32_t* __restrict__ new_idx, double* __restrict__ new_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        const int64_t nr = rule.newid ? rule.newid[g] : g;
        if (nr < 0) continue;
        int64_t clo = 0, chi = INT64_MAX;
        if (rule.cis) {
            const int s = hb_find_seg(rule.seg_off, rule.nseg, g);
            clo = rule.seg_off[s];
            chi = rule.seg_off[s + 1];
        }
        int64_t w = new_indptr[nr - new_row0];
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t base = s; base < e; base += 128) {  // 4 chunks of 32 entries in flight, written in order
            int c[4];
            double v[4];
            bool kp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t k = base + 32 * q + lane;
                c[q] = (k < e) ? hb_ldg_stream_i1(idx + k) : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t k = base + 32 * q + lane;
                v[q] = (k < e) ? hb_ldg_stream_d1(val + k) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t k = base + 32 * q + lane;
                kp[q] = (k < e) && rule.keep(g, clo, chi, c[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned int bal = __ballot_sync(0xffffffffu, kp[q]);
                if (kp[q]) {
                    const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                    new_idx[pos] = rule.newid ? rule.newid[c[q]] : c[q];
                    new_val[pos] = v[q];
                }
                w += __popc(bal);
            }
        }
    }
}

__global__ void hb_newid_kernel(const uint8_t* __restrict__ remove, const int64_t* __restrict__ scan, int32_t* __restrict__ newid, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) newid[i] = remove[i] ? -1 : (int32_t)scan[i];
}

__global__ void hb_keepflag_kernel(const uint8_t* __restrict__ remove, int32_t* __restrict__ keep, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) keep[i] = remove[i] ? 0 : 1;
}

// Rebuild the store with `rule`; d_newid may be NULL (no bin removal).  Full blocks only when bins are removed.
static int hb_compact_impl(hb_csr* h, const int32_t* d_newid, int64_t new_n, int64_t blk_row0, int64_t blk_rows, int cis,
                           int drop_diag) {
    cudaStream_t s = h->stream;
    HbKeepRule rule;
    rule.newid = d_newid;
    rule.seg_off = h->d_seg_off;
    rule.nseg = h->nseg;
    rule.cis = (cis && h->nseg > 1) ? 1 : 0;
    rule.drop_diag = drop_diag;
    const int64_t new_rows = d_newid ? blk_rows : h->n_rows;
    const int64_t new_row0 = d_newid ? blk_row0 : h->row0;
    int64_t* d_cnt = nullptr;
    int64_t* d_new_indptr = nullptr;
    HB_CUDA(hb_dmalloc(&d_cnt, sizeof(int64_t) * (size_t)(new_rows + 1)));
    HB_CUDA(hb_dmalloc(&d_new_indptr, sizeof(int64_t) * (size_t)(new_rows + 1)));
    HB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (size_t)(new_rows + 1), s));
    const int grid = HB_NUM_SMS * 8;
    if (h->n_rows > 0) {
        // new_count is indexed by new LOCAL row: nr - new_row0
        hb_compact_count_kernel<<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->n_rows, h->row0, rule,
                                                     d_cnt - new_row0);
        HB_LAUNCH_CHECK();
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_new_indptr, new_rows);
    HB_LAUNCH_CHECK();
    int64_t new_nnz = 0;
    HB_CUDA(cudaMemcpyAsync(&new_nnz, d_new_indptr + new_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    int32_t* d_new_idx = nullptr;
    double* d_new_val = nullptr;
    HB_CUDA(hb_dmalloc(&d_new_idx, sizeof(int32_t) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    HB_CUDA(hb_dmalloc(&d_new_val, sizeof(double) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    if (h->n_rows > 0 && new_nnz > 0) {
        hb_compact_fill_kernel<<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, rule,
                                                    new_row0, d_new_indptr, d_new_idx, d_new_val);
        HB_LAUNCH_CHECK();
    }
    HB_CUDA(cudaStreamSynchronize(s));
    hb_dfree(d_cnt);
    if (h->owns_csr) {
        hb_dfree(h->d_indptr);
        hb_dfree(h->d_indices);
        hb_dfree(h->d_data);
    }
    hb_drop_pack(h);
    h->owns_csr = true;
    h->phase = 0;
    h->d_indptr = d_new_indptr;
    h->d_indices = d_new_idx;
    h->d_data = d_new_val;
    h->nnz = new_nnz;
    if (d_newid) {
        h->n_rows = new_rows;
        h->n_cols = new_n;
        h->row0 = new_row0;
    }
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// symmetry: sum |M - M^T| and sum M   (iterativeCorrection.py:35)
// ---------------------------------------------------------------------------------------------
// Position of column `want` in the ascending column list idx[lo..hi) of one row, or -1.  Interpolation search:
// Hi-C rows are dense near the diagonal, so the first guess usually lands within a step or two (a plain binary
// search costs ~12 dependent cache misses in ANOTHER row per entry and made this test slower than the whole ICE
// solve); falls back to bisection when the guess does not shrink the range.
__device__ __forceinline__ int64_t hb_find_col(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int want) {
    hi -= 1;  // inclusive
    int guard = 0;
    while (lo <= hi) {
        const int clo = idx[lo];
        if (want < clo) return -1;
        if (want == clo) return lo;
        const int chi = idx[hi];
        if (want > chi) return -1;
        if (want == chi) return hi;
        if (hi - lo < 2) return -1;
        int64_t mid;
        if (guard < 4) {
            mid = lo + (int64_t)(((double)(want - clo) / (double)(chi - clo)) * (double)(hi - lo));
            if (mid <= lo) mid = lo + 1;
            if (mid >= hi) mid = hi - 1;
        } else {
            mid = (lo + hi) >> 1;
        }
        ++guard;
        const int cm = idx[mid];
        if (cm == want) return mid;
        if (cm < want) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// Fast exact-symmetry screen: one streaming pass, no look-ups.  Every entry (i, j, a) contributes the fingerprint
// f(i, j, bits(a)) to multiset A and f(j, i, bits(a)) to multiset B; the matrix equals its transpose iff A == B.
// The multisets are compared through two commutative 64-bit digests each (sums mod 2^64 of two differently mixed keys),
// so the accumulation order -- and the sharding: row blocks add their digests with one allreduce -- is irrelevant and a
// false "equal" needs a ~2^-64 collision.  Matrices that are symmetric by
// construction (what the driver produces, hicmatrix symmetric fill) take this path: ratio = 0 exactly, as numpy
// would compute.  Anything else falls through to the exact |M - M^T| pass below, which yields the ratio itself.
__device__ __forceinline__ unsigned long long hb_mixkey(unsigned long long z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebull;
    z ^= z >> 31;
    return z;
}

__global__ void __launch_bounds__(256)
hb_symmetry_digest_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t n_rows, int64_t row0, unsigned long long* digest /* [0..1] A ; [2..3] B */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long sa = 0, xa = 0, sb = 0, xb = 0;
    for (int64_t lrow = gwarp; lrow < n_rows; lrow += nwarps) {
        const int64_t e = indptr[lrow + 1];
        const unsigned long long row = (unsigned long long)(row0 + lrow);
        for (int64_t k = indptr[lrow] + lane; k < e; k += 32 * 4) {
            int c[4];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = (k + 32 * j < e) ? hb_ldg_stream_i1(idx + k + 32 * j) : -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (k + 32 * j < e) ? hb_ldg_stream_d1(val + k + 32 * j) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j] < 0) continue;
                const unsigned long long bits = hb_mixkey((unsigned long long)__double_as_longlong(v[j]));
                const unsigned long long ka = hb_mixkey(((row << 32) | (unsigned int)c[j]) ^ bits);
                const unsigned long long kb = hb_mixkey((((unsigned long long)(unsigned int)c[j] << 32) | row) ^ bits);
                sa += ka;
                xa += hb_mixkey(ka + 0x9E3779B97F4A7C15ull);
                sb += kb;
                xb += hb_mixkey(kb + 0x9E3779B97F4A7C15ull);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        xa += __shfl_xor_sync(0xffffffffu, xa, o);
        xb += __shfl_xor_sync(0xffffffffu, xb, o);
    }
    if (lane == 0) {
        atomicAdd(digest + 0, sa);
        atomicAdd(digest + 1, xa);
        atomicAdd(digest + 2, sb);
        atomicAdd(digest + 3, xb);
    }
}

// the path's only collective is a sum-allreduce of doubles: ship each 64-bit digest as four 16-bit limbs (limb sums
// over <= 2^30 ranks are exact in fp64) and recombine mod 2^64 on the host
__global__ void hb_digest_limbs_kernel(const unsigned long long* __restrict__ digest, double* __restrict__ limbs) {
    const int t = threadIdx.x;
    if (t < 16) limbs[t] = (double)((digest[t >> 2] >> (16 * (t & 3))) & 0xffffull);
}

// UPPER_ONLY: only entries above the diagonal look up their transposed partner and count twice (the pair (i,j),
// (j,i) contributes |a-b| on both sides); the number of matched pairs is returned so that the caller can tell
// whether every entry BELOW the diagonal has a partner too (then the one-sided sum is the complete statistic).
template <bool UPPER_ONLY>
__global__ void __launch_bounds__(256)
hb_symmetry_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                   int64_t n_rows, double* sums /* [0] sum|M-M^T|, [1] sum M */,
                   unsigned long long* counts /* [0] entries below the diagonal, [1] matched upper entries */) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double sd = 0.0, sm = 0.0;
    unsigned long long n_lower = 0, n_match = 0;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
            const int c = idx[k];
            const double a = val[k];
            sm += a;
            if (UPPER_ONLY) {
                if (c < row) {
                    ++n_lower;
                    continue;
                }
                if (c == row) continue;
            }
            const int64_t pos = hb_find_col(idx, indptr[c], indptr[c + 1], (int)row);
            const double d = pos >= 0 ? fabs(a - val[pos]) : 2.0 * fabs(a);
            if (UPPER_ONLY) {
                sd += pos >= 0 ? 2.0 * d : d;
                if (pos >= 0) ++n_match;
            } else {
                sd += d;
            }
        }
    }
    sd = hb_warp_sum(sd);
    sm = hb_warp_sum(sm);
    for (int o = 16; o > 0; o >>= 1) {
        n_lower += __shfl_xor_sync(0xffffffffu, n_lower, o);
        n_match += __shfl_xor_sync(0xffffffffu, n_match, o);
    }
    if (lane == 0) {
        atomicAdd(sums + 0, sd);
        atomicAdd(sums + 1, sm);
        if (UPPER_ONLY) {
            atomicAdd(counts + 0, n_lower);
            atomicAdd(counts + 1, n_match);
        }
    }
}

// ---- exact |M - M^T| on row blocks ------------------------------------------------------------------------------------
// The partner of entry (i, j) lives in the row block that owns row j.  Every rank hands the owner of each column range
// the entries it stores there (hb_dev_extract_cols: (row, col, value) triplets, any order); the owner looks every
// received a_ij up in its own row j (hb_dev_partner_sums) and adds |a_ij - a_ji|, or 2 |a_ij| when it stores no a_ji --
// the cell (j, i) then contributes the same amount and nobody else visits it.  Summed over all ranks this is
// sum |M - M^T| over all cells, the numerator of iterativeCorrection.py:35.  The host moves the triplets (all-to-all).
__global__ void __launch_bounds__(256)
hb_extract_cols_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                       int64_t n_rows, int64_t row0, int col_lo, int col_hi, unsigned long long* cursor, int32_t* __restrict__ rows,
                       int32_t* __restrict__ cols, double* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t e = indptr[row + 1];
        for (int64_t base = indptr[row]; base < e; base += 32) {
            const int64_t k = base + lane;
            const int c = k < e ? idx[k] : -1;
            const bool take = c >= col_lo && c < col_hi;
            const unsigned int bal = __ballot_sync(0xffffffffu, take);
            if (bal == 0) continue;
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(cursor, (unsigned long long)__popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (take && rows != nullptr) {
                const unsigned long long p = at + __popc(bal & ((1u << lane) - 1u));
                rows[p] = (int32_t)(row0 + row);
                cols[p] = c;
                vals[p] = val[k];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
hb_partner_sums_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                       int64_t n_rows, int64_t row0, int64_t n, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                       const double* __restrict__ vals, double* sum_out) {
    double sd = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = (int64_t)cols[t] - row0;   // local row that owns the partner
        const double a = vals[t];
        if (j < 0 || j >= n_rows) {
            sd += 2.0 * fabs(a);                     // (cannot happen when the host routed the triplets by owner)
            continue;
        }
        const int64_t pos = hb_find_col(idx, indptr[j], indptr[j + 1], rows[t]);
        sd += pos >= 0 ? fabs(a - val[pos]) : 2.0 * fabs(a);
    }
    sd = hb_warp_sum(sd);
    if ((threadIdx.x & 31) == 0 && sd != 0.0) atomicAdd(sum_out, sd);
}

extern "C" int hb_dev_extract_cols(hb_csr* h, int64_t col_lo, int64_t col_hi, int64_t* count_out, int32_t* d_rows, int32_t* d_cols,
                                   double* d_vals) {
    HB_REQUIRE(h != nullptr && count_out != nullptr && col_lo >= 0 && col_lo <= col_hi && col_hi <= h->n_cols, "bad arguments");
    HB_REQUIRE((d_rows == nullptr) == (d_cols == nullptr) && (d_rows == nullptr) == (d_vals == nullptr), "all three outputs or none");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    HB_CUDA(cudaMemsetAsync(h->d_scalar_bits, 0, sizeof(unsigned long long), s));
    if (h->n_rows > 0 && h->nnz > 0) {
        hb_extract_cols_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, (int)col_lo,
                                                              (int)col_hi, h->d_scalar_bits, d_rows, d_cols, d_vals);
        HB_LAUNCH_CHECK();
    }
    unsigned long long c = 0;
    HB_CUDA(cudaMemcpyAsync(&c, h->d_scalar_bits, sizeof(c), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    *count_out = (int64_t)c;
    return HB_OK;
}

extern "C" int hb_dev_partner_sums(hb_csr* h, int64_t n, const int32_t* d_rows, const int32_t* d_cols, const double* d_vals,
                                   double* sum_out) {
    HB_REQUIRE(h != nullptr && sum_out != nullptr && n >= 0, "bad arguments");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    double* d_sum = reinterpret_cast<double*>(h->d_scalar_bits + 2);
    HB_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), s));
    if (n > 0) {
        hb_partner_sums_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0, n, d_rows,
                                                              d_cols, d_vals, d_sum);
        HB_LAUNCH_CHECK();
    }
    HB_CUDA(cudaMemcpyAsync(sum_out, d_sum, sizeof(double), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// lossless narrow copy of the values for the streaming kernels
// ---------------------------------------------------------------------------------------------
__global__ void hb_pack_probe_kernel(const double* __restrict__ data, int64_t nnz, unsigned int* flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool not_u16 = false, not_f32 = false;
    for (; i < nnz; i += stride) {
        const double v = data[i];
        if (!(v >= 0.0 && v <= 65535.0 && v == floor(v))) not_u16 = true;
        if (!((double)(float)v == v)) not_f32 = true;  // also false for NaN: a NaN never packs
    }
    if (__any_sync(0xffffffffu, not_u16) && (threadIdx.x & 31) == 0) atomicOr(flags + 0, 1u);
    if (__any_sync(0xffffffffu, not_f32) && (threadIdx.x & 31) == 0) atomicOr(flags + 1, 1u);
}

template <typename VT>
__global__ void hb_pack_convert_kernel(const double* __restrict__ in, VT* __restrict__ out, int64_t nnz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) out[i] = (VT)in[i];
}

void hb_drop_pack(hb_csr* h) {
    hb_dfree(h->d_pack);
    h->d_pack = nullptr;
    h->pack_kind = 0;
}

extern "C" {

int hb_csr_pack_values(hb_csr* h, int* kind_out) {
    HB_RANGE("hb_csr_pack_values");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    if (h->pack_kind != 0 && h->d_pack != nullptr) {
        // already there: hb_csr_create's narrow ingest wrote it, or an earlier call (every operation that changes the
        // values or the structure drops the copy, so what exists is current)
        if (kind_out) *kind_out = h->pack_kind;
        return HB_OK;
    }
    hb_drop_pack(h);
    cudaStream_t s = h->stream;
    if (kind_out) *kind_out = 0;
    if (h->nnz == 0) return HB_OK;
    unsigned int* d_flags = reinterpret_cast<unsigned int*>(h->d_scalar_bits);
    HB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(unsigned int) * 2, s));
    const double* src = h->d_data + h->phase;
    hb_pack_probe_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(src, h->nnz, d_flags);
    HB_LAUNCH_CHECK();
    unsigned int flags[2] = {1, 1};
    HB_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    const int kind = !flags[0] ? 1 : (!flags[1] ? 2 : 0);
    if (kind == 0) return HB_OK;
    const size_t elt = kind == 1 ? sizeof(uint16_t) : sizeof(float);
    HB_CUDA(hb_dmalloc(&h->d_pack, elt * (size_t)(h->nnz + h->phase + 8) + 16));
    HB_CUDA(cudaMemsetAsync(h->d_pack, 0, elt * 8, s));
    if (kind == 1)
        hb_pack_convert_kernel<uint16_t><<<HB_NUM_SMS * 8, 256, 0, s>>>(src, (uint16_t*)h->d_pack + h->phase, h->nnz);
    else
        hb_pack_convert_kernel<float><<<HB_NUM_SMS * 8, 256, 0, s>>>(src, (float*)h->d_pack + h->phase, h->nnz);
    HB_LAUNCH_CHECK();
    HB_CUDA(cudaStreamSynchronize(s));
    h->pack_kind = kind;
    if (kind_out) *kind_out = kind;
    return HB_OK;
}

int hb_scrub(hb_csr* h, int64_t counts[2]) {
    HB_RANGE("hb_scrub");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    if (h->pack_kind == 1 && h->d_pack != nullptr) {
        // a current uint16 copy exists (narrow ingest, or hb_csr_pack_values): every value is a finite count, there is
        // nothing to scrub, and the copy stays valid
        if (counts) counts[0] = counts[1] = 0;
        return HB_OK;
    }
    hb_drop_pack(h);  // the values may change
    HB_CUDA(cudaMemsetAsync(h->d_scalar_bits, 0, sizeof(unsigned long long) * 2, s));
    if (h->nnz > 0) {
        hb_scrub_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_data + h->phase, h->nnz, h->d_scalar_bits);
        HB_LAUNCH_CHECK();
    }
    unsigned long long c[2] = {0, 0};
    HB_CUDA(cudaMemcpyAsync(c, h->d_scalar_bits, sizeof(c), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    if (counts) {
        counts[0] = (int64_t)c[0];
        counts[1] = (int64_t)c[1];
    }
    return HB_OK;
}

static int hb_row_stats_impl(hb_csr* h, double* row_sum, double* diag, int cis_only);

int hb_row_stats(hb_csr* h, double* row_sum, double* diag) { return hb_row_stats_impl(h, row_sum, diag, 0); }

int hb_row_stats_cis(hb_csr* h, double* row_sum, double* diag) { return hb_row_stats_impl(h, row_sum, diag, 1); }

static int hb_row_stats_impl(hb_csr* h, double* row_sum, double* diag, int cis_only) {
    HB_RANGE("hb_row_stats_impl");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    size_t n = (size_t)(h->n_rows > 0 ? h->n_rows : 1);
    double *d_rs = nullptr, *d_dg = nullptr;
    HbScratch scratch;   // freed on every return path
    // row sums and diagonals share one buffer so that a sharded run assembles both with ONE sum-allreduce
    HB_CUDA(scratch.alloc(&d_rs, sizeof(double) * 2 * n));
    d_dg = d_rs + n;
    HB_CUDA(cudaMemsetAsync(d_rs, 0, sizeof(double) * 2 * n, s));
    if (h->n_rows > 0) {
        hb_row_stats_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                           h->d_seg_off, h->nseg, cis_only, d_rs, d_dg);
        HB_LAUNCH_CHECK();
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && row_sum && h->n_rows) e = cudaMemcpy(row_sum, d_rs, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && diag && h->n_rows) e = cudaMemcpy(diag, d_dg, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return hb_cuda_fail(e, "row stats", __FILE__, __LINE__);
    return HB_OK;
}

int hb_mad_filter(hb_csr* h, double lo, double hi, uint8_t* mask_out, double* zscore_out, double* median_out,
                  double* mad_out) {
    HB_RANGE("hb_mad_filter");
    HB_REQUIRE(h != nullptr && mask_out != nullptr, "bad arguments");
    HB_REQUIRE(h->world > 1 || (h->row0 == 0 && h->n_rows == h->n_cols),
               "hb_mad_filter needs the full matrix on one device unless hb_set_comm was called");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int64_t n = h->n_cols;
    const int nseg = h->nseg;
    if (n == 0) return HB_OK;
    double *d_rs, *d_dg, *d_pts, *d_diff, *d_abs, *d_z = nullptr, *d_med, *d_mad;
    uint8_t* d_mask;
    HbSelState* d_st;
    unsigned int* d_hist;
    HbScratch scratch;   // freed on every return path
    // row sums and diagonals share one buffer so that a sharded run assembles both with ONE sum-allreduce
    HB_CUDA(scratch.alloc(&d_rs, sizeof(double) * 2 * n));
    d_dg = d_rs + n;
    HB_CUDA(cudaMemsetAsync(d_rs, 0, sizeof(double) * 2 * n, s));
    HB_CUDA(scratch.alloc(&d_pts, sizeof(double) * n));
    HB_CUDA(scratch.alloc(&d_diff, sizeof(double) * n));
    HB_CUDA(scratch.alloc(&d_abs, sizeof(double) * n));
    if (zscore_out) HB_CUDA(scratch.alloc(&d_z, sizeof(double) * n));
    HB_CUDA(scratch.alloc(&d_med, sizeof(double) * nseg));
    HB_CUDA(scratch.alloc(&d_mad, sizeof(double) * nseg));
    HB_CUDA(scratch.alloc(&d_mask, n));
    HB_CUDA(scratch.alloc(&d_st, sizeof(HbSelState) * 2 * nseg));
    HB_CUDA(scratch.alloc(&d_hist, sizeof(unsigned int) * 256 * nseg));
    HB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned int) * 256 * nseg, s));
    // with segments set, the statistics are those of each chromosome's diagonal block (hicCorrectMatrix.py:557-568)
    if (h->n_rows > 0) {
        hb_row_stats_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                           h->d_seg_off, nseg, 1, d_rs + h->row0, d_dg + h->row0);
        HB_LAUNCH_CHECK();
    }
    HB_REQUIRE(h->world == 1 || h->allreduce != nullptr, "sharded hb_mad_filter needs the allreduce callback (hb_set_comm)");
    if (h->world > 1 && h->allreduce(h->comm_ctx, d_rs, 2 * n, (void*)s) != 0) {
        hb_set_error("collective callback failed");
        return HB_ERR_CUDA;
    }
    hb_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_rs, d_dg, d_pts, n);
    HB_LAUNCH_CHECK();
    rc = hb_segment_median(h, d_pts, 0, d_st, d_hist, d_med);
    if (rc == HB_OK) {
        hb_diff_kernel<<<h->n_tiles, 256, 0, s>>>(d_pts, h->d_tile_row0, h->d_tile_len, h->d_tile_seg, d_med, d_diff, d_abs);
        g_hb_launches.fetch_add(1);
        rc = hb_segment_median(h, d_abs, 1, d_st, d_hist, d_mad);
    }
    if (rc == HB_OK) {
        hb_zscore_kernel<<<h->n_tiles, 256, 0, s>>>(d_diff, h->d_tile_row0, h->d_tile_len, h->d_tile_seg, d_mad, lo, hi, d_z, d_mask);
        g_hb_launches.fetch_add(1);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == HB_OK && e == cudaSuccess) {
        e = cudaMemcpy(mask_out, d_mask, n, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && zscore_out) e = cudaMemcpy(zscore_out, d_z, sizeof(double) * n, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && median_out) e = cudaMemcpy(median_out, d_med, sizeof(double) * nseg, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && mad_out) e = cudaMemcpy(mad_out, d_mad, sizeof(double) * nseg, cudaMemcpyDeviceToHost);
    }
    if (rc) return rc;
    if (e != cudaSuccess) return hb_cuda_fail(e, "mad filter", __FILE__, __LINE__);
    return HB_OK;
}

int hb_csr_compact(hb_csr* h, const uint8_t* remove_mask) {
    HB_RANGE("hb_csr_compact");
    HB_REQUIRE(h != nullptr && remove_mask != nullptr, "bad arguments");
    // works on a row block too: every rank passes the same global mask and keeps its own surviving rows
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int64_t n = h->n_cols;
    if (n == 0) return HB_OK;
    // new segment offsets + new size from the host mask
    std::vector<int64_t> new_off(h->nseg + 1, 0);
    int64_t kept = 0, kept_before_block = 0, kept_in_block = 0;
    {
        int sgi = 0;
        for (int64_t i = 0; i <= n; ++i) {
            while (sgi <= h->nseg && h->seg_off[sgi] == i) new_off[sgi++] = kept;
            if (i == h->row0) kept_before_block = kept;
            if (i < n && !remove_mask[i]) {
                ++kept;
                if (i >= h->row0 && i < h->row0 + h->n_rows) ++kept_in_block;
            }
        }
    }
    uint8_t* d_rm = nullptr;
    int32_t *d_keep = nullptr, *d_newid = nullptr;
    int64_t* d_scan = nullptr;
    HbScratch scratch;   // freed on every return path
    HB_CUDA(scratch.alloc(&d_rm, n));
    HB_CUDA(scratch.alloc(&d_keep, sizeof(int32_t) * n));
    HB_CUDA(scratch.alloc(&d_newid, sizeof(int32_t) * n));
    HB_CUDA(scratch.alloc(&d_scan, sizeof(int64_t) * (n + 1)));
    HB_CUDA(cudaMemcpyAsync(d_rm, remove_mask, n, cudaMemcpyHostToDevice, s));
    const unsigned g = (unsigned)((n + 255) / 256);
    hb_keepflag_kernel<<<g, 256, 0, s>>>(d_rm, d_keep, n);
    HB_LAUNCH_CHECK();
    hb_scan_kernel<int32_t><<<1, 1024, 0, s>>>(d_keep, d_scan, n);
    HB_LAUNCH_CHECK();
    hb_newid_kernel<<<g, 256, 0, s>>>(d_rm, d_scan, d_newid, n);
    HB_LAUNCH_CHECK();
    rc = hb_compact_impl(h, d_newid, kept, kept_before_block, kept_in_block, 0, 0);
    if (rc) return rc;
    hb_free_state(h);
    h->seg_off = new_off;
    return hb_build_tiles(h);
}

int hb_csr_keep_cis(hb_csr* h) {
    HB_RANGE("hb_csr_keep_cis");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    if (h->nseg <= 1) return HB_OK;
    return hb_compact_impl(h, nullptr, 0, 0, 0, 1, 0);
}

int hb_csr_keep_upper(hb_csr* h) {
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    return hb_compact_impl(h, nullptr, 0, 0, 0, 0, 2);
}

int hb_diag_zero(hb_csr* h) {
    HB_RANGE("hb_diag_zero");
    HB_REQUIRE(h != nullptr, "null handle");
    int rc = hb_use_device(h);
    if (rc) return rc;
    return hb_compact_impl(h, nullptr, 0, 0, 0, 0, 1);
}

int hb_symmetry(hb_csr* h, double* ratio_out) {
    HB_RANGE("hb_symmetry");
    HB_REQUIRE(h != nullptr && ratio_out != nullptr, "bad arguments");
    const bool sharded = h->world > 1;
    HB_REQUIRE(sharded || (h->row0 == 0 && h->n_rows == h->n_cols), "hb_symmetry needs the full matrix or hb_set_comm");
    HB_REQUIRE(!sharded || h->allreduce != nullptr, "sharded hb_symmetry needs the allreduce callback (hb_set_comm)");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    double* d_sums = nullptr;
    HB_CUDA(hb_dmalloc(&d_sums, sizeof(double) * (4 + 16)));
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(d_sums + 2);
    double sums[2] = {0, 0};
    unsigned long long counts[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(d_sums, 0, sizeof(double) * (4 + 16), s);
    if (e == cudaSuccess && (h->nnz > 0 || sharded)) {
        // screen: exactly symmetric? (one streaming pass)
        unsigned long long dg[4] = {0, 1, 2, 3};
        if (h->nnz > 0) {
            hb_symmetry_digest_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, h->row0,
                                                                     reinterpret_cast<unsigned long long*>(d_sums));
            g_hb_launches.fetch_add(1);
        }
        if (sharded) {
            double limbs[16];
            hb_digest_limbs_kernel<<<1, 32, 0, s>>>(reinterpret_cast<unsigned long long*>(d_sums), d_sums + 4);
            g_hb_launches.fetch_add(1);
            if (h->allreduce(h->comm_ctx, d_sums + 4, 16, (void*)s) != 0) {
                hb_dfree(d_sums);
                hb_set_error("hb_symmetry: the allreduce callback failed");
                return HB_ERR_CUDA;
            }
            e = cudaMemcpyAsync(limbs, d_sums + 4, sizeof(limbs), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            for (int q = 0; q < 4; ++q) {
                dg[q] = 0;
                for (int l = 0; l < 4; ++l) dg[q] += (unsigned long long)limbs[4 * q + l] << (16 * l);
            }
            hb_dfree(d_sums);
            if (e != cudaSuccess) return hb_cuda_fail(e, "symmetry", __FILE__, __LINE__);
            if (dg[0] == dg[2] && dg[1] == dg[3]) {
                *ratio_out = 0.0;
                return HB_OK;
            }
            // The exact ratio needs every entry's transposed partner, which lives in another rank's block: the host
            // moves the entries to their partners' owners and calls hb_dev_extract_cols / hb_dev_partner_sums
            // (store.py: symmetry_ratio_sharded) -- the reference accepts a relative asymmetry below 1e-10.
            hb_set_error("the matrix is not exactly symmetric (row-block screen); the exact ratio needs the block exchange");
            *ratio_out = 1.0;
            return HB_ERR_NOT_SYMMETRIC;
        }
        e = cudaMemcpyAsync(dg, d_sums, sizeof(dg), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e == cudaSuccess && dg[0] == dg[2] && dg[1] == dg[3]) {
            hb_dfree(d_sums);
            *ratio_out = 0.0;
            return HB_OK;
        }
        if (e == cudaSuccess) e = cudaMemsetAsync(d_sums, 0, sizeof(double) * 4, s);
    }
    if (e == cudaSuccess && h->nnz > 0) {
        hb_symmetry_kernel<true><<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, d_sums, d_counts);
        g_hb_launches.fetch_add(1);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && counts[0] != counts[1]) {
        // some entry below the diagonal has no partner above it: the one-sided sum missed it -> full two-sided pass
        e = cudaMemsetAsync(d_sums, 0, sizeof(double) * 4, s);
        if (e == cudaSuccess) {
            hb_symmetry_kernel<false><<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows, d_sums, d_counts);
            g_hb_launches.fetch_add(1);
            e = cudaMemcpyAsync(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost, s);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    hb_dfree(d_sums);
    if (e != cudaSuccess) return hb_cuda_fail(e, "symmetry", __FILE__, __LINE__);
    const double cells = (double)h->n_cols * (double)h->n_cols;
    const double num = sums[0] / cells;
    double den = sums[1] / cells;
    if (den < 0) den = -den;
    *ratio_out = num / (1.0 * den);
    return HB_OK;
}

}  // extern "C"
