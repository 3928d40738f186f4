// Deterministic synthetic Hi-C matrices generated directly in HBM (benchmark inputs).
// Not part of the reference: the reference has no generator, its benchmarks ran on Rao et al. data
// (docs/content/tools/hicCorrectMatrix.rst:30-44).  The matrix is a pure function of
// (seed, chromosome layout, parameters) and of the unordered pair {i, j}, so that every rank can
// generate its own row block without communication and (i,j) / (j,i) always agree.
// tests/synth_ref.py restates the same function in numpy for CPU cross-checks.
//
// All floating-point steps are single correctly rounded IEEE operations (no FMA contraction, no
// transcendental functions when gamma == 1), so CPU and GPU produce bit-identical matrices.
#include <algorithm>

#include "hb_internal.cuh"
#include "hb_scan.cuh"

#define HB_SYNTH_MAX_TRANS 32
#define HB_TAG_BASE (1ull << 40)

__host__ __device__ __forceinline__ uint64_t hb_mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t hb_hash3(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = hb_mix64(seed + 0x9E3779B97F4A7C15ull * (a + 1));
    return hb_mix64(z ^ (0xD1B54A32D192ED03ull * (b + 1)));
}
__host__ __device__ __forceinline__ double hb_u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct HbSynthDev {
    uint64_t seed;
    int64_t n;
    int32_t band;
    int32_t n_trans;
    double d0, gamma, lambda0, low_frac, empty_frac, lambda_trans;
    int64_t trans_c[HB_SYNTH_MAX_TRANS];
};

// g[i] = planted multiplicative bias of bin i; 0 marks an empty bin
__global__ void hb_synth_bins_kernel(HbSynthDev p, double* __restrict__ g) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double v;
    if (hb_u01(hb_hash3(p.seed, (uint64_t)i, HB_TAG_BASE + 0)) < p.empty_frac) v = 0.0;
    else if (hb_u01(hb_hash3(p.seed, (uint64_t)i, HB_TAG_BASE + 1)) < p.low_frac) v = 0.02;
    else {
        const double u = hb_u01(hb_hash3(p.seed, (uint64_t)i, HB_TAG_BASE + 2));
        v = __dadd_rn(0.2, __dmul_rn(4.8, __dmul_rn(__dmul_rn(u, u), u)));
    }
    g[i] = v;
}

__device__ __forceinline__ bool hb_synth_cis_exists(const HbSynthDev& p, int64_t i, int64_t j, uint64_t& hp) {
    const int64_t a = i < j ? i : j, b = i < j ? j : i;
    hp = hb_hash3(p.seed, (uint64_t)a, (uint64_t)b);
    const double d1 = (double)(b - a + 1);
    const double u = hb_u01(hp);
    if (p.gamma == 1.0) return __dmul_rn(u, d1) < p.d0;
    return u < pow(p.d0 / d1, p.gamma);
}

__device__ __forceinline__ double hb_synth_count(const HbSynthDev& p, uint64_t hp, double lam, double gi, double gj, bool i_first) {
    const double u2 = hb_u01(hb_mix64(hp + 0x9E3779B97F4A7C15ull));
    const double ga = i_first ? gi : gj, gb = i_first ? gj : gi;  // (lam * g[min]) * g[max]: symmetric
    return 1.0 + floor(__dmul_rn(__dmul_rn(__dmul_rn(lam, ga), gb), u2));
}

// chromosome [lo, hi) containing bin i
__device__ __forceinline__ void hb_synth_chrom(const int64_t* __restrict__ chrom_off, int nchrom, int64_t i, int64_t& lo, int64_t& hi) {
    int a = 0, b = nchrom;
    while (b - a > 1) {
        int m = (a + b) >> 1;
        if (chrom_off[m] <= i) a = m; else b = m;
    }
    lo = chrom_off[a];
    hi = chrom_off[a + 1];
}

// FILL == false: cnt[row] = number of entries; FILL == true: write them (ascending columns)
template <bool FILL>
__global__ void __launch_bounds__(256)
hb_synth_rows_kernel(HbSynthDev p, const int64_t* __restrict__ chrom_off, int nchrom, const double* __restrict__ g,
                     int64_t row0, int64_t n_rows, int64_t* __restrict__ cnt, const int64_t* __restrict__ indptr,
                     int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t i = row0 + row;
        const double gi = g[i];
        if (gi == 0.0) {
            if (!FILL && lane == 0) cnt[row] = 0;
            continue;
        }
        int64_t clo, chi;
        hb_synth_chrom(chrom_off, nchrom, i, clo, chi);
        const int64_t blo = (i - p.band > clo) ? i - p.band : clo;
        const int64_t bhi = (i + p.band < chi - 1) ? i + p.band : chi - 1;  // inclusive
        // far ("trans") partners j = (c_k - i) mod n, kept when outside the cis band window
        int64_t tj = -1;
        double tv = 0.0;
        if (lane < p.n_trans) {
            int64_t j = (p.trans_c[lane] - i) % p.n;
            if (j < 0) j += p.n;
            if ((j < blo || j > bhi) && g[j] != 0.0) {
                tj = j;
                if (FILL) {
                    const int64_t a = i < j ? i : j, b = i < j ? j : i;
                    tv = hb_synth_count(p, hb_hash3(p.seed, (uint64_t)a, (uint64_t)b), p.lambda_trans, gi, g[j], i < j);
                }
            }
        }
        const unsigned int t_before = __ballot_sync(0xffffffffu, tj >= 0 && tj < blo);
        const unsigned int t_after = __ballot_sync(0xffffffffu, tj >= 0 && tj > bhi);
        int64_t w = 0;
        if (FILL) {
            w = indptr[row];
            // rank of this lane's partner among the partners on the same side (ascending column order)
            int rank = 0;
            for (int m = 0; m < p.n_trans; ++m) {
                const int64_t other = __shfl_sync(0xffffffffu, tj, m);
                if (tj >= 0 && other >= 0 && other < tj && ((other < blo) == (tj < blo))) ++rank;
            }
            if (tj >= 0 && tj < blo) {
                out_idx[w + rank] = (int32_t)tj;
                out_val[w + rank] = tv;
            }
            w += __popc(t_before);
            // cis band
            for (int64_t base = blo; base <= bhi; base += 32) {
                const int64_t j = base + lane;
                bool ex = false;
                uint64_t hp = 0;
                double gj = 0.0;
                if (j <= bhi) {
                    gj = g[j];
                    if (gj != 0.0) ex = hb_synth_cis_exists(p, i, j, hp);
                }
                const unsigned int bal = __ballot_sync(0xffffffffu, ex);
                if (ex) {
                    const int64_t d = i < j ? j - i : i - j;
                    const double lam = __ddiv_rn(p.lambda0, __dadd_rn(1.0, __ddiv_rn((double)d, p.d0)));
                    const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                    out_idx[pos] = (int32_t)j;
                    out_val[pos] = hb_synth_count(p, hp, lam, gi, gj, i <= j);
                }
                w += __popc(bal);
            }
            if (tj >= 0 && tj > bhi) {
                out_idx[w + rank] = (int32_t)tj;
                out_val[w + rank] = tv;
            }
        } else {
            long long c = __popc(t_before) + __popc(t_after);
            for (int64_t base = blo; base <= bhi; base += 32) {
                const int64_t j = base + lane;
                bool ex = false;
                uint64_t hp;
                if (j <= bhi && g[j] != 0.0) ex = hb_synth_cis_exists(p, i, j, hp);
                c += __popc(__ballot_sync(0xffffffffu, ex));
            }
            if (lane == 0) cnt[row] = c;
        }
    }
}


extern "C" int hb_synth_create(hb_csr** out, const hb_synth_params* prm, const int64_t* chrom_bins, int64_t row0,
                               int64_t n_rows, int device) {
    HB_RANGE("hb_synth_create");
    HB_REQUIRE(out != nullptr && prm != nullptr && chrom_bins != nullptr, "bad arguments");
    HB_REQUIRE(prm->nchrom >= 1 && prm->nchrom <= 4096, "bad chromosome count");
    HB_REQUIRE(prm->n_trans >= 0 && prm->n_trans <= HB_SYNTH_MAX_TRANS, "n_trans must be <= 32");
    HB_REQUIRE(prm->band >= 0 && prm->d0 > 0, "bad band / d0");
    std::vector<int64_t> off(prm->nchrom + 1, 0);
    for (int c = 0; c < prm->nchrom; ++c) {
        HB_REQUIRE(chrom_bins[c] >= 0, "negative chromosome size");
        off[c + 1] = off[c] + chrom_bins[c];
    }
    const int64_t n = off[prm->nchrom];
    HB_REQUIRE(n > 0 && n < 2147483647LL && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= n, "bad row block");
    if (hb_device_count() <= 0) {
        hb_set_error("no CUDA device visible: libhicbalance has no CPU fallback");
        return HB_ERR_NO_DEVICE;
    }
    HB_CUDA(cudaSetDevice(device));
    HbSynthDev p;
    p.seed = prm->seed;
    p.n = n;
    p.band = prm->band;
    p.d0 = prm->d0;
    p.gamma = prm->gamma;
    p.lambda0 = prm->lambda0;
    p.low_frac = prm->low_frac;
    p.empty_frac = prm->empty_frac;
    p.lambda_trans = prm->lambda_trans;
    {
        std::vector<int64_t> cs;
        for (int k = 0; k < prm->n_trans; ++k) cs.push_back((int64_t)(hb_hash3(prm->seed, (uint64_t)k, HB_TAG_BASE + 3) % (uint64_t)n));
        std::sort(cs.begin(), cs.end());
        cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
        p.n_trans = (int32_t)cs.size();
        for (int k = 0; k < HB_SYNTH_MAX_TRANS; ++k) p.trans_c[k] = k < (int)cs.size() ? cs[k] : 0;
    }
    int64_t* d_off = nullptr;
    double* d_g = nullptr;
    int64_t *d_cnt = nullptr, *d_indptr = nullptr;
    int32_t* d_idx = nullptr;
    double* d_val = nullptr;
    cudaStream_t s = nullptr;
    int64_t nnz = 0;
    int rc = HB_OK;
    cudaError_t e = cudaSuccess;
    const int grid = HB_NUM_SMS * 8;
#define SY(expr)                                                   \
    do {                                                           \
        e = (expr);                                                \
        if (e != cudaSuccess) {                                    \
            rc = hb_cuda_fail(e, #expr, __FILE__, __LINE__);       \
            goto fail;                                             \
        }                                                          \
    } while (0)
    SY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    SY(hb_dmalloc(&d_off, sizeof(int64_t) * off.size()));
    SY(hb_dmalloc(&d_g, sizeof(double) * (size_t)n));
    SY(hb_dmalloc(&d_cnt, sizeof(int64_t) * (size_t)(n_rows + 1)));
    SY(hb_dmalloc(&d_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    SY(cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, s));
    hb_synth_bins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, d_g);
    g_hb_launches.fetch_add(1);
    if (n_rows > 0) {
        hb_synth_rows_kernel<false><<<grid, 256, 0, s>>>(p, d_off, prm->nchrom, d_g, row0, n_rows, d_cnt, nullptr, nullptr, nullptr);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_indptr, n_rows);
    g_hb_launches.fetch_add(1);
    SY(cudaMemcpyAsync(&nnz, d_indptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SY(cudaStreamSynchronize(s));
    SY(hb_dmalloc(&d_idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1) + 16));
    SY(hb_dmalloc(&d_val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1) + 16));
    if (n_rows > 0 && nnz > 0) {
        hb_synth_rows_kernel<true><<<grid, 256, 0, s>>>(p, d_off, prm->nchrom, d_g, row0, n_rows, nullptr, d_indptr, d_idx, d_val);
        g_hb_launches.fetch_add(1);
    }
    SY(cudaStreamSynchronize(s));
    SY(cudaGetLastError());
#undef SY
    hb_dfree(d_off);
    hb_dfree(d_g);
    hb_dfree(d_cnt);
    cudaStreamDestroy(s);
    rc = hb_dev_csr_wrap(out, n_rows, n, row0, nnz, d_indptr, d_idx, d_val, device);
    if (rc != HB_OK) {
        hb_dfree(d_indptr);
        hb_dfree(d_idx);
        hb_dfree(d_val);
        return rc;
    }
    (*out)->owns_csr = true;  // the generator's buffers now belong to the handle
    return HB_OK;
fail:
    hb_dfree(d_off);
    hb_dfree(d_g);
    hb_dfree(d_cnt);
    hb_dfree(d_indptr);
    hb_dfree(d_idx);
    hb_dfree(d_val);
    if (s) cudaStreamDestroy(s);
    return rc;
}
