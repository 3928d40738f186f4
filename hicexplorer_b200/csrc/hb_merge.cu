// Bin merging on the device CSR store: sum groups of consecutive bins (rows AND columns) into one bin.
// Replaces hicexplorer/reduceMatrix.py:12-232 (reduce_matrix with use_triu=True, diagonal=True) as called by
// hicexplorer/hicMergeMatrixBins.py:264 -- the step right before correction in the reference's workflow
// (SURVEY.md 8f, rank 2).  Contract: include/hicbalance.h (hb_csr_merge_bins).
//
// The reference sums the UPPER triangle into the merged cells and mirrors the result, keeping the diagonal once.
// On the full symmetric store that is: for merged row I and merged column J != I, R[I][J] = sum of A[i][j] over
// i in I, j in J (read from the rows of I); for the diagonal cell only the entries with j >= i count.
//
// Old rows are sorted by column and the bin map is monotone, so merged ids ascend along every old row; no sort and
// no atomics are needed, the result is deterministic and exact for integer counts.  Pass 1 counts the non-zero
// cells per merged row, a scan builds indptr, pass 2 writes them (kernel comment below for the mapping).
#include <climits>

#include "hb_internal.cuh"
#include "hb_scan.cuh"

#define HB_MERGE_DEPTH 4     // steps of 32 entries in flight per warp
#define HB_MERGE_WINDOW 512  // merged columns accumulated at a time per warp (4 KB of shared memory per warp)

// A warp owns merged row I.  The merged row is produced window by window (HB_MERGE_WINDOW merged columns): for each
// window the warp streams, one old row of the group after the other, the stretch of that row whose columns fall
// into the window -- 32 consecutive entries per step, coalesced -- maps them to merged ids, combines the entries
// of equal id inside the step with a segmented shuffle reduction (equal ids are adjacent: the row is sorted and the
// map monotone) and lets the first lane of each run add into a dense shared-memory accumulator.  One warp touches
// its accumulator strictly in program order (row after row, step after step), so the sums are deterministic.  The
// window is then swept in column order: non-zero cells are counted (pass 1) or written (pass 2).
//
// ADD = true turns the same machinery into C = A + B (hicSumMatrices.py:59): output row I is the "merge" of two rows,
// row I of A and row I of B (second store in indptr_b / idx_b / val_b), with the identity column map and no diagonal
// rule; cells that sum to exactly 0 are dropped, like scipy's csr_plus_csr.
template <bool FILL, bool ADD>
__global__ void __launch_bounds__(256)
hb_merge_rows_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                     const int64_t* __restrict__ indptr_b, const int32_t* __restrict__ idx_b, const double* __restrict__ val_b,
                     const int32_t* __restrict__ map, const int64_t* __restrict__ group_start,
                     const int64_t* __restrict__ group_end, int64_t n_new, int max_run,
                     int64_t* __restrict__ cnt, const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_idx,
                     double* __restrict__ out_val) {
    __shared__ double acc_all[8][HB_MERGE_WINDOW];
    const int lane = threadIdx.x & 31;
    double* acc = acc_all[threadIdx.x >> 5];
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int i = lane; i < HB_MERGE_WINDOW; i += 32) acc[i] = 0.0;
    __syncwarp();
    for (int64_t I = gwarp; I < n_new; I += nwarps) {
        const int64_t g0 = ADD ? I : group_start[I];
        const int k = ADD ? 2 : (int)(group_end[I] - g0);
        // lane r keeps the cursor of old row g0 + r (ADD: lane 0 = row I of A, lane 1 = row I of B)
        int64_t cur = 0, end = 0;
        if (ADD) {
            if (lane < 2) {
                const int64_t* ip = lane ? indptr_b : indptr;
                cur = ip[I];
                end = ip[I + 1];
            }
        } else if (lane < k) {
            cur = indptr[g0 + lane];
            end = indptr[g0 + lane + 1];
        }
        int64_t w = FILL ? out_ptr[I] : 0;
        long long n_out = 0;
        for (;;) {
            // next window = the one holding the smallest unread merged id over the k rows
            int head = INT_MAX;
            if (ADD) {
                if (cur < end) head = (lane ? idx_b : idx)[cur];
            } else if (cur < end) {
                // skip dropped columns at the cursor (map < 0) so that the head is a real id
                while (cur < end) {
                    const int J = map[idx[cur]];
                    if (J >= 0) {
                        head = J;
                        break;
                    }
                    ++cur;
                }
            }
            int m = head;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (m == INT_MAX) break;
            const int J_lo = m - (m % HB_MERGE_WINDOW);
            const int J_hi = J_lo + HB_MERGE_WINDOW;
            unsigned int touched = 0;  // 32-column chunks of the window that received something
            for (int r = 0; r < k; ++r) {
                // far (trans) entries make most windows sparse: a row whose next entry lies beyond this window is
                // skipped without touching memory
                if (__shfl_sync(0xffffffffu, head, r) >= J_hi) continue;
                int64_t c0 = __shfl_sync(0xffffffffu, cur, r);
                const int64_t e0 = __shfl_sync(0xffffffffu, end, r);
                const int64_t row = g0 + r;
                const int32_t* __restrict__ idx_r = (ADD && r) ? idx_b : idx;
                const double* __restrict__ val_r = (ADD && r) ? val_b : val;
                bool stop = false;
                while (!stop) {
                    // HB_MERGE_DEPTH steps of 32 consecutive entries are loaded before the first one is consumed: the
                    // stretch of a row inside a window is contiguous, so only the last step of a stretch is partial
                    int cc[HB_MERGE_DEPTH], JJ[HB_MERGE_DEPTH];
                    double vv[HB_MERGE_DEPTH];
#pragma unroll
                    for (int u = 0; u < HB_MERGE_DEPTH; ++u) {
                        const int64_t p = c0 + 32 * u + lane;
                        cc[u] = p < e0 ? hb_ldg_stream_i1(idx_r + p) : -1;
                    }
#pragma unroll
                    for (int u = 0; u < HB_MERGE_DEPTH; ++u) {
                        const int64_t p = c0 + 32 * u + lane;
                        vv[u] = p < e0 ? hb_ldg_stream_d1(val_r + p) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < HB_MERGE_DEPTH; ++u) {
                        int J = INT_MAX;  // sentinel: beyond the row
                        if (cc[u] >= 0) {
                            const int Jm = ADD ? cc[u] : __ldg(map + cc[u]);
                            J = Jm < 0 ? -1 : Jm;
                        }
                        JJ[u] = J;
                    }
#pragma unroll
                    for (int u = 0; u < HB_MERGE_DEPTH; ++u) {
                        if (stop) break;
                        const int J = JJ[u], c = cc[u];
                        const double v = vv[u];
                        // entries are consumed up to the first one that belongs to a later window
                        const unsigned int beyond = __ballot_sync(0xffffffffu, J >= J_hi);
                        const int take = beyond ? (__ffs(beyond) - 1) : 32;
                        const bool mine = lane < take;
                        // dropped column, or lower-triangle entry of the diagonal cell: contributes nothing
                        const bool live = mine && J >= 0 && (ADD || !(J == (int)I && c < row));
                        const int key = mine ? J : INT_MAX - 1;
                        double s = live ? v : 0.0;
                        // segmented inclusive scan from the right: after it the first lane of a run holds the run's sum.
                        // A run (equal merged ids inside one old row) is at most max_run entries long -- the size of the
                        // largest group, 1 for A + B -- so log2 steps of that suffice.
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            if (o >= max_run) break;
                            const double t = __shfl_down_sync(0xffffffffu, s, o);
                            const int kt = __shfl_down_sync(0xffffffffu, key, o);
                            if (lane + o < 32 && kt == key) s += t;
                        }
                        const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
                        const bool first = mine && J >= 0 && (lane == 0 || kprev != key);
                        if (first) acc[J - J_lo] += s;
                        touched |= __reduce_or_sync(0xffffffffu, first ? (1u << ((J - J_lo) >> 5)) : 0u);
                        __syncwarp();
                        c0 += take;
                        if (take < 32 || c0 >= e0) stop = true;
                    }
                }
                if (lane == r) cur = c0;
            }
            // sweep the touched chunks of the window in column order
            while (touched) {
                const int base = (__ffs(touched) - 1) << 5;
                touched &= touched - 1;
                const double t = acc[base + lane];
                const bool nz = t != 0.0;  // eliminate_zeros (reduceMatrix.py:230); NaN != 0 is kept, like scipy
                const unsigned int bal = __ballot_sync(0xffffffffu, nz);
                if (bal) {
                    if (FILL && nz) {
                        const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                        out_idx[pos] = J_lo + base + lane;
                        out_val[pos] = t;
                    }
                    acc[base + lane] = 0.0;
                    w += __popc(bal);
                    n_out += __popc(bal);
                }
            }
            __syncwarp();
        }
        if (!FILL && lane == 0) cnt[I] = n_out;
    }
}


extern "C" int hb_csr_merge_bins(hb_csr* h, const int32_t* bin_map, int64_t n_new) {
    HB_RANGE("hb_csr_merge_bins");
    HB_REQUIRE(h != nullptr && bin_map != nullptr, "bad arguments");
    HB_REQUIRE(h->row0 == 0 && h->n_rows == h->n_cols, "hb_csr_merge_bins needs the full matrix on one device");
    HB_REQUIRE(n_new >= 0 && n_new <= h->n_cols, "bad merged size");
    const int64_t n = h->n_cols;
    // the map must send runs of consecutive bins to 0, 1, 2, ... in order (dropped bins = -1), <= 32 bins per group
    std::vector<int64_t> g_start((size_t)(n_new > 0 ? n_new : 1), -1), g_end((size_t)(n_new > 0 ? n_new : 1), -1);
    int max_run = 1;   // size of the largest group = longest run of equal merged ids inside an old row
    {
        int64_t next = 0;
        for (int64_t i = 0; i < n; ++i) {
            const int32_t J = bin_map[i];
            if (J < 0) continue;
            HB_REQUIRE(J < n_new, "bin map exceeds the merged size");
            if (g_start[J] < 0) {
                HB_REQUIRE(J == next, "merged ids must appear in ascending order");
                g_start[J] = i;
                ++next;
            } else {
                HB_REQUIRE(g_end[J] == i, "the bins of a group must be consecutive");
            }
            g_end[J] = i + 1;
            HB_REQUIRE(g_end[J] - g_start[J] <= 32, "at most 32 bins per merged bin");
            if (g_end[J] - g_start[J] > max_run) max_run = (int)(g_end[J] - g_start[J]);
        }
        HB_REQUIRE(next == n_new, "some merged bins are empty");
    }
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);
    int64_t *d_gs = nullptr, *d_ge = nullptr, *d_cnt = nullptr, *d_ptr = nullptr;
    int32_t *d_map = nullptr, *d_idx = nullptr;
    double* d_val = nullptr;
    int64_t new_nnz = 0;
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 8;
    const size_t ng = (size_t)(n_new > 0 ? n_new : 1);
#define MG(expr)                                                  \
    do {                                                          \
        e = (expr);                                               \
        if (e != cudaSuccess) {                                   \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);      \
            goto fail;                                            \
        }                                                         \
    } while (0)
    MG(hb_dmalloc(&d_gs, sizeof(int64_t) * ng));
    MG(hb_dmalloc(&d_ge, sizeof(int64_t) * ng));
    MG(hb_dmalloc(&d_map, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
    MG(hb_dmalloc(&d_cnt, sizeof(int64_t) * (ng + 1)));
    MG(hb_dmalloc(&d_ptr, sizeof(int64_t) * (ng + 1)));
    MG(cudaMemcpyAsync(d_gs, g_start.data(), sizeof(int64_t) * ng, cudaMemcpyHostToDevice, s));
    MG(cudaMemcpyAsync(d_ge, g_end.data(), sizeof(int64_t) * ng, cudaMemcpyHostToDevice, s));
    if (n > 0) MG(cudaMemcpyAsync(d_map, bin_map, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    MG(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (ng + 1), s));
    if (n_new > 0) {
        hb_merge_rows_kernel<false, false><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, nullptr, nullptr, nullptr,
                                                                d_map, d_gs, d_ge, n_new, max_run, d_cnt, nullptr, nullptr, nullptr);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n_new);
    g_hb_launches.fetch_add(1);
    MG(cudaMemcpyAsync(&new_nnz, d_ptr + n_new, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MG(cudaStreamSynchronize(s));
    MG(cudaGetLastError());
    MG(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    MG(hb_dmalloc(&d_val, sizeof(double) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    if (n_new > 0 && new_nnz > 0) {
        hb_merge_rows_kernel<true, false><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, nullptr, nullptr, nullptr,
                                                               d_map, d_gs, d_ge, n_new, max_run, nullptr, d_ptr, d_idx, d_val);
        g_hb_launches.fetch_add(1);
    }
    MG(cudaStreamSynchronize(s));
    MG(cudaGetLastError());
#undef MG
    hb_dfree(d_gs);
    hb_dfree(d_ge);
    hb_dfree(d_map);
    hb_dfree(d_cnt);
    if (h->owns_csr) {
        hb_dfree(h->d_indptr);
        hb_dfree(h->d_indices);
        hb_dfree(h->d_data);
    }
    h->owns_csr = true;
    h->phase = 0;
    h->d_indptr = d_ptr;
    h->d_indices = d_idx;
    h->d_data = d_val;
    h->nnz = new_nnz;
    h->n_rows = h->n_cols = n_new;
    h->row0 = 0;
    hb_free_state(h);
    h->nseg = 1;
    h->seg_off = {0, n_new};
    return hb_build_tiles(h);
fail:
    hb_dfree(d_gs);
    hb_dfree(d_ge);
    hb_dfree(d_map);
    hb_dfree(d_cnt);
    hb_dfree(d_ptr);
    hb_dfree(d_idx);
    hb_dfree(d_val);
    return rc;
}


// C = A + B on two stores of the same shape (hicSumMatrices.py:59): `h` becomes the sum, `other` is left untouched
extern "C" int hb_csr_add(hb_csr* h, const hb_csr* other) {
    HB_RANGE("hb_csr_add");
    HB_REQUIRE(h != nullptr && other != nullptr && h != other, "bad arguments");
    HB_REQUIRE(h->n_rows == other->n_rows && h->n_cols == other->n_cols && h->row0 == other->row0,
               "the matrices have different shapes");
    HB_REQUIRE(h->device == other->device, "both stores must live on the same device");
    HB_REQUIRE(h->phase == 0 && other->phase == 0, "row blocks with an alignment phase are not supported here");
    int rc = hb_use_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    hb_drop_pack(h);
    const int64_t n = h->n_rows;
    int64_t *d_cnt = nullptr, *d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    int64_t new_nnz = 0;
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 7;   // 32 KB of shared memory per CTA
    const size_t nn = (size_t)(n > 0 ? n : 1);
#define AD(expr)                                                  \
    do {                                                          \
        e = (expr);                                               \
        if (e != cudaSuccess) {                                   \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);      \
            goto fail;                                            \
        }                                                         \
    } while (0)
    AD(cudaStreamSynchronize(other->stream));   // whatever produced `other` is complete before this stream reads it
    AD(hb_dmalloc(&d_cnt, sizeof(int64_t) * (nn + 1)));
    AD(hb_dmalloc(&d_ptr, sizeof(int64_t) * (nn + 1)));
    AD(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (nn + 1), s));
    if (n > 0) {
        hb_merge_rows_kernel<false, true><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, other->d_indptr,
                                                               other->d_indices, other->d_data, nullptr, nullptr, nullptr, n, 1,
                                                               d_cnt, nullptr, nullptr, nullptr);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n);
    g_hb_launches.fetch_add(1);
    AD(cudaMemcpyAsync(&new_nnz, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    AD(cudaStreamSynchronize(s));
    AD(cudaGetLastError());
    AD(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    AD(hb_dmalloc(&d_val, sizeof(double) * (size_t)(new_nnz > 0 ? new_nnz : 1) + 16));
    if (n > 0 && new_nnz > 0) {
        hb_merge_rows_kernel<true, true><<<grid, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, other->d_indptr,
                                                              other->d_indices, other->d_data, nullptr, nullptr, nullptr, n, 1,
                                                              nullptr, d_ptr, d_idx, d_val);
        g_hb_launches.fetch_add(1);
    }
    AD(cudaStreamSynchronize(s));
    AD(cudaGetLastError());
#undef AD
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
