// Knight-Ruiz balancing (Knight & Ruiz 2012, routine `bnewt`: inexact Newton iteration with a
// conjugate-gradient inner solve and a cone-boundary step limiter) on the device CSR store.
// Replaces the third-party C++/Eigen object `krbalancing.kr_balancing` that the reference calls at
// hicexplorer/hicCorrectMatrix.py:718-732, 744-757 (constants, the A + 1e-5*I augmentation and the
// rescale convention restated in oracle/kr_oracle.py).  Contract: include/hicbalance.h.
//
// Device mapping.  Every CG step is ONE hb_spmv_kernel launch with the Hadamard update fused into
// its row epilogue:  w = x o ((A + eps I)(x o p)) + v o p   (12 B per stored entry, HBM-bound);
// the remaining work is O(n) vector kernels whose dot products / min / max are reduced over fixed
// row tiles (deterministic, rank-count independent) into device-resident scalars.  alpha, beta
// and the step decisions are computed FROM those device scalars, so the host synchronises once
// per mat-vec (to evaluate the loop conditions), not once per vector operation.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>

#include "hb_scan.cuh"
#include "hb_spmv.cuh"

namespace cg = cooperative_groups;

enum {
    SC_RHO = 0,    // rho_km1
    SC_RHO2 = 1,   // rho_km2
    SC_PTW = 2,    // p^T w
    SC_MIN = 3,    // min(y + alpha p)
    SC_MAX = 4,    // max(y + alpha p)
    SC_GAMMA = 5,  // boundary step length
    SC_ROUT = 6,   // r^T r after the outer update
    SC_XMIN = 7,   // min(x) (breakdown check)
    SC_A = 8,      // rescale: sum of (A+eps I) o (x x^T)
    SC_B = 9,      // rescale: sum of (A+eps I)
    SC_ERR = 10,   // set by hb_peer_wait_kernel on a timeout
    SC_WORDS = 16
};

struct HbKrVec {
    double *x, *y, *p, *z, *rk, *v, *w, *u;
};

struct HbTiles {
    const int64_t* __restrict__ row0;
    const int32_t* __restrict__ len;
    int n_tiles;
};

// Block-wide (256 threads) fixed-order reduction of (sum, min, max); the result is valid in thread 0.
__device__ __forceinline__ void hb_block_red3(double& s, double& mn, double& mx, double* sh /* 24 */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    s = hb_warp_sum(s);
    mx = hb_warp_max(mx);
    mn = -hb_warp_max(-mn);
    __syncthreads();   // sh may still be read from the previous use
    if (lane == 0) {
        sh[w] = s;
        sh[8 + w] = mn;
        sh[16 + w] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            s += sh[i];
            mn = fmin(mn, sh[8 + i]);
            mx = fmax(mx, sh[16 + i]);
        }
    }
}

// Fold the per-tile partials (sum, min, max) in tile order; every thread of the block returns the result.
__device__ __forceinline__ void hb_fold3(const double* part, int n_tiles, double& s, double& mn, double& mx,
                                         double* sh /* 24 */, double* bc /* 3 */) {
    s = 0.0;
    mn = INFINITY;
    mx = -INFINITY;
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
        s += __ldcg(part + 3 * i + 0);
        mn = fmin(mn, __ldcg(part + 3 * i + 1));
        mx = fmax(mx, __ldcg(part + 3 * i + 2));
    }
    hb_block_red3(s, mn, mx, sh);
    if (threadIdx.x == 0) {
        bc[0] = s;
        bc[1] = mn;
        bc[2] = mx;
    }
    __syncthreads();
    s = bc[0];
    mn = bc[1];
    mx = bc[2];
}

// Generic tiled vector kernel (step-level path: one launch per vector operation): F::elem runs per row and may
// contribute to a sum / min / max; the last block to finish folds the per-tile partials in tile order and calls
// F::finalize (thread 0).  The persistent solver kernel below runs the same element code and the same two reduction
// helpers, so both paths give bit-identical results.
template <class F>
__global__ void __launch_bounds__(256)
hb_kr_vec_kernel(HbTiles tiles, F f, double* __restrict__ part /* 3*n_tiles */, unsigned int* counter, double* sc) {
    __shared__ double sh[24], bc[3];
    __shared__ int s_last;
    if (f.skip(sc)) return;
    f.prepare(sc);
    const int t = blockIdx.x;
    const int64_t r0 = tiles.row0[t];
    const int len = tiles.len[t];
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < len; i += blockDim.x) f.elem(r0 + i, s, mn, mx);
    if (!F::REDUCES) return;
    hb_block_red3(s, mn, mx, sh);
    if (threadIdx.x == 0) {
        part[3 * t + 0] = s;
        part[3 * t + 1] = mn;
        part[3 * t + 2] = mx;
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    hb_fold3(part, tiles.n_tiles, s, mn, mx, sh, bc);
    if (threadIdx.x == 0) {
        *counter = 0u;
        f.finalize(s, mn, mx, sc);
    }
}

struct FBase {
    static constexpr bool REDUCES = true;
    __device__ __forceinline__ bool skip(const double*) const { return false; }
    __device__ __forceinline__ void prepare(const double*) {}
};

// x = 1, y = 1
struct FInit : FBase {
    static constexpr bool REDUCES = false;
    HbKrVec q;
    __device__ __forceinline__ void elem(int64_t g, double&, double&, double&) const {
        q.x[g] = 1.0;
        q.y[g] = 1.0;
    }
    __device__ __forceinline__ void finalize(double, double, double, double*) const {}
};

// rk = 1 - v ; ROUT = RHO = rk^T rk
struct FResid : FBase {
    HbKrVec q;
    __device__ __forceinline__ void elem(int64_t g, double& s, double&, double&) const {
        const double r = 1.0 - q.v[g];
        q.rk[g] = r;
        s = fma(r, r, s);
    }
    __device__ __forceinline__ void finalize(double s, double, double, double* sc) const {
        sc[SC_ROUT] = s;
        sc[SC_RHO] = s;
    }
};

// k == 1:  Z = rk / v ; p = Z ; u = x o p ; RHO = rk^T Z
struct FFirst : FBase {
    HbKrVec q;
    __device__ __forceinline__ void elem(int64_t g, double& s, double&, double&) const {
        const double r = q.rk[g];
        const double z = r / q.v[g];
        q.z[g] = z;
        q.p[g] = z;
        q.u[g] = q.x[g] * z;
        s = fma(r, z, s);
    }
    __device__ __forceinline__ void finalize(double s, double, double, double* sc) const { sc[SC_RHO] = s; }
};

// k > 1:  beta = RHO / RHO2 ; p = Z + beta p ; u = x o p
struct FDir : FBase {
    static constexpr bool REDUCES = false;
    HbKrVec q;
    double beta;
    __device__ __forceinline__ void prepare(const double* sc) { beta = sc[SC_RHO] / sc[SC_RHO2]; }
    __device__ __forceinline__ void elem(int64_t g, double&, double&, double&) const {
        const double p = q.z[g] + __dmul_rn(beta, q.p[g]);
        q.p[g] = p;
        q.u[g] = q.x[g] * p;
    }
    __device__ __forceinline__ void finalize(double, double, double, double*) const {}
};

// PTW = p^T w
struct FDot : FBase {
    HbKrVec q;
    __device__ __forceinline__ void elem(int64_t g, double& s, double&, double&) const { s = fma(q.p[g], q.w[g], s); }
    __device__ __forceinline__ void finalize(double s, double, double, double* sc) const { sc[SC_PTW] = s; }
};

// alpha = RHO / PTW ; MIN / MAX of ynew = y + alpha p
struct FYnew : FBase {
    HbKrVec q;
    double alpha;
    __device__ __forceinline__ void prepare(const double* sc) { alpha = sc[SC_RHO] / sc[SC_PTW]; }
    __device__ __forceinline__ void elem(int64_t g, double&, double& mn, double& mx) const {
        const double yn = q.y[g] + __dmul_rn(alpha, q.p[g]);
        mn = fmin(mn, yn);
        mx = fmax(mx, yn);
    }
    __device__ __forceinline__ void finalize(double, double mn, double mx, double* sc) const {
        sc[SC_MIN] = mn;
        sc[SC_MAX] = mx;
    }
};

// regular CG step (only if no cone bound was hit): y = ynew ; rk -= alpha w ; Z = rk / v ;
// RHO2 = RHO ; RHO = rk^T Z
struct FStep : FBase {
    HbKrVec q;
    double delta, Delta, alpha;
    __device__ __forceinline__ bool skip(const double* sc) const { return sc[SC_MIN] <= delta || sc[SC_MAX] >= Delta; }
    __device__ __forceinline__ void prepare(const double* sc) { alpha = sc[SC_RHO] / sc[SC_PTW]; }
    __device__ __forceinline__ void elem(int64_t g, double& s, double&, double&) const {
        q.y[g] = q.y[g] + __dmul_rn(alpha, q.p[g]);
        const double r = q.rk[g] - __dmul_rn(alpha, q.w[g]);
        q.rk[g] = r;
        const double z = r / q.v[g];
        q.z[g] = z;
        s = fma(r, z, s);
    }
    __device__ __forceinline__ void finalize(double s, double, double, double* sc) const {
        sc[SC_RHO2] = sc[SC_RHO];
        sc[SC_RHO] = s;
    }
};

// boundary step length: lower bound  gamma = min_{ap<0} (delta - y)/ap ; upper bound  min_{ynew>Delta} (Delta - y)/ap
struct FGamma : FBase {
    HbKrVec q;
    double bound, alpha;
    int upper;
    __device__ __forceinline__ void prepare(const double* sc) { alpha = sc[SC_RHO] / sc[SC_PTW]; }
    __device__ __forceinline__ void elem(int64_t g, double&, double& mn, double&) const {
        const double ap = __dmul_rn(alpha, q.p[g]);
        const double y = q.y[g];
        const bool sel = upper ? (y + ap > bound) : (ap < 0.0);
        if (sel) mn = fmin(mn, (bound - y) / ap);
    }
    __device__ __forceinline__ void finalize(double, double mn, double, double* sc) const { sc[SC_GAMMA] = mn; }
};

// y += gamma * (alpha p)
struct FBoundStep : FBase {
    static constexpr bool REDUCES = false;
    HbKrVec q;
    double alpha, gamma;
    __device__ __forceinline__ void prepare(const double* sc) {
        alpha = sc[SC_RHO] / sc[SC_PTW];
        gamma = sc[SC_GAMMA];
    }
    __device__ __forceinline__ void elem(int64_t g, double&, double&, double&) const {
        q.y[g] = q.y[g] + __dmul_rn(gamma, __dmul_rn(alpha, q.p[g]));
    }
    __device__ __forceinline__ void finalize(double, double, double, double*) const {}
};

// x = x o y ; y = 1 ; XMIN = min(x)
struct FXUpdate : FBase {
    HbKrVec q;
    __device__ __forceinline__ void elem(int64_t g, double&, double& mn, double&) const {
        const double x = q.x[g] * q.y[g];
        q.x[g] = x;
        q.y[g] = 1.0;
        mn = fmin(mn, x);
    }
    __device__ __forceinline__ void finalize(double, double mn, double, double* sc) const { sc[SC_XMIN] = mn; }
};

// plain sums of two n-vectors (rescale factor)
struct FSum2 : FBase {
    const double* a;
    const double* b;
    int which;
    __device__ __forceinline__ void elem(int64_t g, double& s, double&, double&) const { s += which ? b[g] : a[g]; }
    __device__ __forceinline__ void finalize(double s, double, double, double* sc) const { sc[which ? SC_B : SC_A] = s; }
};

// per-row sums over the upper triangle of A + eps I, off-diagonals doubled (krbalancing rescale_norm_vector)
__global__ void __launch_bounds__(256)
hb_kr_rescale_rows_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t n_rows, int64_t row0, const double* __restrict__ x, double eps,
                          double* __restrict__ scaled, double* __restrict__ orig) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        const double xi = x[g];
        double so = 0.0, ss = 0.0, dg = 0.0;
        for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
            const int c = idx[k];
            const double a = val[k];
            if (c == g) dg += a;
            else if (c > g) {
                so += a * 2;
                ss += a * xi * x[c] * 2;
            }
        }
        so = hb_warp_sum(so);
        ss = hb_warp_sum(ss);
        dg = hb_warp_sum(dg);
        if (lane == 0) {
            const double d = dg + eps;
            orig[g] = so + d;
            scaled[g] = ss + d * xi * xi;
        }
    }
}

// upper triangle (diagonal always stored) of (A + eps I) o (x x^T)
__global__ void __launch_bounds__(256)
hb_kr_upper_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t n_rows, int64_t row0,
                         int64_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        long long c = 0;
        for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) c += idx[k] > row0 + row ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[row] = c + 1;
    }
}

__global__ void __launch_bounds__(256)
hb_kr_upper_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                        int64_t n_rows, int64_t row0, const double* __restrict__ x, double eps, const int64_t* __restrict__ out_ptr,
                        int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = gwarp; row < n_rows; row += nwarps) {
        const int64_t g = row0 + row;
        const double xi = x[g];
        int64_t w = out_ptr[row] + 1;
        double dg = 0.0;
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t base = s; base < e; base += 32) {
            const int64_t k = base + lane;
            int c = -1;
            double a = 0.0;
            if (k < e) {
                c = idx[k];
                a = val[k];
            }
            if (c == g) dg += a;
            const bool kp = c > g;
            const unsigned int bal = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
                out_idx[pos] = c;
                out_val[pos] = a * xi * x[c];
            }
            w += __popc(bal);
        }
        dg = hb_warp_sum(dg);
        if (lane == 0) {
            out_idx[out_ptr[row]] = (int32_t)g;
            out_val[out_ptr[row]] = (dg + eps) * xi * xi;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The whole Knight-Ruiz solve in ONE cooperative launch (hb_kr_solve_kernel): `bnewt`'s outer Newton loop, the inner
// conjugate-gradient loop, its cone-boundary branches and the eta update all run on the device.  Every CG step is
//     P1  [apply the previous step: y += alpha p ; rk -= alpha w]  p = rk/v + beta p ; u = x o p      -- grid barrier
//     SpMV  w = x o ((A + eps I) u) + v o p  over the block's rows, Hadamard update fused in the row epilogue
//           (peer mode: rows stored into every replica over NVLink)                                   -- grid barrier
//     P2  per-tile partials of p^T w                                                                  -- grid barrier
//     P3  alpha = rho / p^T w ; per-tile min / max of y + alpha p and the rk^T Z of the tentative step -- grid barrier
// after each barrier EVERY CTA folds the per-tile partials in tile order, so all CTAs (and all ranks) hold the same
// scalars bit for bit and take the same branches; nothing is ever read back by the host.  The element arithmetic and
// the two reduction helpers are those of the step-level kernels above -> identical results.
struct HbKrSolveArgs {
    const int64_t* indptr;
    const int32_t* idx;
    const void* val;
    int64_t n_rows, row0, n_cols;
    const int64_t* tile_row0;
    const int32_t* tile_len;
    int n_tiles;
    double *x, *y, *p, *rk, *v, *w, *u;
    double* part;   // 6 x 3 x n_tiles
    double* sc;     // SC_WORDS: results
    int32_t* h_loop;
    unsigned long long* row_counters;
    double* err;
    double tol, delta, Delta, eps;
    int max_outer, row_parity;
    HbLoopPeer peer;
};

enum { KP_FIRST = 0, KP_DOT = 1, KP_STEP = 2, KP_GAMMA = 3, KP_RESID = 4, KP_XMIN = 5, KP_ARRAYS = 6 };

// w_g = x_g (t + eps u_g) [+ v_g p_g]; all operands are rewritten inside the launch -> L2 loads
struct HbKrLoopEpi {
    const double* u;
    const double* x;
    const double* v;   // nullptr: outer residual  v = x o ((A + eps I) x)
    const double* p;
    double eps;
    struct Pre {
        double u, x, v, p;
    };
    __device__ __forceinline__ bool active(int64_t) const { return true; }
    __device__ __forceinline__ Pre pre(int64_t g) const {
        Pre q;
        q.u = hb_ld_gather<true>(u + g);
        q.x = hb_ld_gather<true>(x + g);
        q.v = v != nullptr ? hb_ld_gather<true>(v + g) : 0.0;
        q.p = v != nullptr ? hb_ld_gather<true>(p + g) : 0.0;
        return q;
    }
    template <class Out>
    __device__ __forceinline__ void row(int64_t g, double t, double, int lane, const Out& out, const Pre& q) const {
        double r = fma(eps, q.u, t);
        r *= q.x;
        if (v != nullptr) r = fma(q.v, q.p, r);
        out.store(g, r, lane);
    }
    __device__ __forceinline__ void warp_done() const {}
    __device__ __forceinline__ double local_max() const { return 0.0; }
};

template <typename VT>
__global__ void __launch_bounds__(HB_SPMV_THREADS, HB_SPMV_CTAS_PER_SM)
hb_kr_solve_kernel(const __grid_constant__ HbKrSolveArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[24], bc[3];
    const int lane = threadIdx.x & 31;
    const bool peer = a.peer.world > 1;
    const int world = peer ? a.peer.world : 1;
    const int nt = a.n_tiles;
    double* const part = a.part;
    int exchanges = 0;
    int code = 0;   // 0 ok, 1 breakdown, 2 peer timeout

    // one exchange: SpMV over the block's rows + grid barrier (+ flags); returns where the assembled vector is
    auto exchange = [&](const double* gathered, bool with_vp) -> const double* {
        const unsigned long long seq = a.peer.seq0 + 1ull + (unsigned long long)exchanges;
        const int pb = (int)(seq & 1ull);
        const int par = (a.row_parity + exchanges) & 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) a.row_counters[par ^ 1] = 0ull;
        HbKrLoopEpi epi;
        epi.u = gathered;
        epi.x = a.x;
        epi.v = with_vp ? a.v : nullptr;
        epi.p = a.p;
        epi.eps = a.eps;
        HbLoopOut po;
        po.peer = &a.peer;
        po.local = a.w;
        po.pb = pb;
        po.world = world;
        hb_spmv_rows<HbKrLoopEpi, false, VT, true, HbLoopOut>(a.indptr, a.idx, (const VT*)a.val, a.n_rows, a.row0, gathered,
                                                               a.row_counters + par, epi, po);
        if (peer && lane < world) __threadfence_system();
        grid.sync();
        if (peer) hb_loop_publish_and_wait(a.peer, pb, seq, a.n_cols, 0.0, a.err);
        ++exchanges;
        return peer ? a.peer.buf[pb][a.peer.rank] : a.w;
    };
    // per-tile pass: f(g, s, mn, mx) over this CTA's tiles, partials into array `which` (or none)
    auto tile_pass = [&](int which, auto f) {
        for (int t = blockIdx.x; t < nt; t += gridDim.x) {
            const int64_t r0 = a.tile_row0[t];
            const int len = a.tile_len[t];
            double s = 0.0, mn = INFINITY, mx = -INFINITY;
            for (int i = threadIdx.x; i < len; i += blockDim.x) f(r0 + i, s, mn, mx);
            if (which >= 0) {
                hb_block_red3(s, mn, mx, sh);
                if (threadIdx.x == 0) {
                    double* q = part + (size_t)which * 3 * nt + 3 * t;
                    q[0] = s;
                    q[1] = mn;
                    q[2] = mx;
                }
            }
        }
    };
    auto fold = [&](int which, double& s, double& mn, double& mx) {
        hb_fold3(part + (size_t)which * 3 * nt, nt, s, mn, mx, sh, bc);
    };

    const double g = 0.9, etamax = 0.1;
    double eta = etamax;
    const double stop_tol = a.tol * 0.5, rt = a.tol * a.tol;
    double rho = 0.0, rho2 = 0.0, rout = 0.0, rold = 0.0, dummy0, dummy1;
    int outer = 0, mvp = 0;

    // x = y = 1 ; v = x o ((A + eps I) x) ; rk = 1 - v ; rho = rout = rk^T rk
    tile_pass(-1, [&](int64_t i, double&, double&, double&) {
        a.x[i] = 1.0;
        a.y[i] = 1.0;
    });
    grid.sync();
    {
        const double* vin = exchange(a.x, false);
        tile_pass(KP_RESID, [&](int64_t i, double& s, double&, double&) {
            const double vv = __ldcg(vin + i);
            a.v[i] = vv;
            const double r = 1.0 - vv;
            a.rk[i] = r;
            s = fma(r, r, s);
        });
        grid.sync();
        if (peer && __ldcg(a.err) != 0.0) code = 2;
        fold(KP_RESID, rho, dummy0, dummy1);
        rout = rold = rho;
    }
    while (code == 0 && rout > rt && outer < a.max_outer) {
        ++outer;
        int k = 0;
        const double innertol = fmax(eta * eta * rout, rt);
        bool pending = false;
        double alpha = 0.0;
        const double* wv = a.w;
        while (rho > innertol) {
            ++k;
            if (k == 1) {
                // Z = rk / v ; p = Z ; u = x o p ; rho = rk^T Z
                tile_pass(KP_FIRST, [&](int64_t i, double& s, double&, double&) {
                    const double r = a.rk[i];
                    const double z = r / a.v[i];
                    a.p[i] = z;
                    a.u[i] = a.x[i] * z;
                    s = fma(r, z, s);
                });
            } else {
                // the step accepted last time, then the new direction: beta = rho / rho2 ; p = Z + beta p ; u = x o p
                const double beta = rho / rho2;
                const double al = alpha;
                const double* wl = wv;
                tile_pass(-1, [&](int64_t i, double&, double&, double&) {
                    const double pi = a.p[i];
                    a.y[i] = a.y[i] + __dmul_rn(al, pi);
                    const double r = a.rk[i] - __dmul_rn(al, __ldcg(wl + i));
                    a.rk[i] = r;
                    const double z = r / a.v[i];
                    const double pn = z + __dmul_rn(beta, pi);
                    a.p[i] = pn;
                    a.u[i] = a.x[i] * pn;
                });
                pending = false;
            }
            grid.sync();
            if (k == 1) fold(KP_FIRST, rho, dummy0, dummy1);
            // w = x o ((A + eps I)(x o p)) + v o p
            wv = exchange(a.u, true);
            {
                const double* wl = wv;
                tile_pass(KP_DOT, [&](int64_t i, double& s, double&, double&) { s = fma(a.p[i], __ldcg(wl + i), s); });
            }
            grid.sync();
            if (peer && __ldcg(a.err) != 0.0) {
                code = 2;
                break;
            }
            double ptw;
            fold(KP_DOT, ptw, dummy0, dummy1);
            alpha = rho / ptw;
            // ynew = y + alpha p : min / max ; tentative step rk - alpha w -> rk^T Z
            {
                const double al = alpha;
                const double* wl = wv;
                tile_pass(KP_STEP, [&](int64_t i, double& s, double& mn, double& mx) {
                    const double yn = a.y[i] + __dmul_rn(al, a.p[i]);
                    mn = fmin(mn, yn);
                    mx = fmax(mx, yn);
                    const double r = a.rk[i] - __dmul_rn(al, __ldcg(wl + i));
                    const double z = r / a.v[i];
                    s = fma(r, z, s);
                });
            }
            grid.sync();
            double rho_new, ymin, ymax;
            fold(KP_STEP, rho_new, ymin, ymax);
            const bool low = ymin <= a.delta;
            const bool high = !low && ymax >= a.Delta;
            if (low || high) {
                if (low && a.delta == 0.0) break;
                // boundary step: gamma = min (bound - y) / (alpha p) over the offending entries ; y += gamma alpha p
                const double bound = low ? a.delta : a.Delta;
                const double al = alpha;
                tile_pass(KP_GAMMA, [&](int64_t i, double&, double& mn, double&) {
                    const double ap = __dmul_rn(al, a.p[i]);
                    const double yy = a.y[i];
                    const bool sel = high ? (yy + ap > bound) : (ap < 0.0);
                    if (sel) mn = fmin(mn, (bound - yy) / ap);
                });
                grid.sync();
                double gamma;
                fold(KP_GAMMA, dummy0, gamma, dummy1);
                tile_pass(-1, [&](int64_t i, double&, double&, double&) {
                    a.y[i] = a.y[i] + __dmul_rn(gamma, __dmul_rn(al, a.p[i]));
                });
                break;
            }
            pending = true;   // y, rk are updated by the next pass that touches them
            rho2 = rho;
            rho = rho_new;
            if (!(rho == rho)) {
                code = 1;
                break;
            }
        }
        if (code != 0) break;
        // x = x o y ; y = 1 ; min(x)   (with the last accepted step folded in)
        {
            const double al = alpha;
            const bool pend = pending;
            tile_pass(KP_XMIN, [&](int64_t i, double&, double& mn, double&) {
                double yy = a.y[i];
                if (pend) yy = yy + __dmul_rn(al, a.p[i]);
                const double xn = a.x[i] * yy;
                a.x[i] = xn;
                a.y[i] = 1.0;
                mn = fmin(mn, xn);
            });
        }
        grid.sync();
        {
            const double* vin = exchange(a.x, false);
            tile_pass(KP_RESID, [&](int64_t i, double& s, double&, double&) {
                const double vv = __ldcg(vin + i);
                a.v[i] = vv;
                const double r = 1.0 - vv;
                a.rk[i] = r;
                s = fma(r, r, s);
            });
        }
        grid.sync();
        if (peer && __ldcg(a.err) != 0.0) {
            code = 2;
            break;
        }
        double xmin;
        fold(KP_RESID, rho, dummy0, dummy1);
        fold(KP_XMIN, dummy0, xmin, dummy1);
        rout = rho;
        mvp += k + 1;
        if (!(rout == rout) || !(xmin > 0.0)) {
            code = 1;
            break;
        }
        const double rat = rout / rold;
        rold = rout;
        const double res_norm = sqrt(rout);
        const double eta_o = eta;
        eta = g * rat;
        if (g * eta_o * eta_o > 0.1) eta = fmax(eta, g * eta_o * eta_o);
        eta = fmax(fmin(eta, etamax), stop_tol / res_norm);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.row_counters[0] = a.row_counters[1] = 0ull;
        a.sc[SC_ROUT] = rout;
        a.h_loop[HB_LP_OUTER] = outer;
        a.h_loop[HB_LP_MATVECS] = mvp;
        a.h_loop[HB_LP_FLAGS] = code;
        a.h_loop[HB_LP_EXCHANGES] = exchanges;
        __threadfence_system();
    }
}

namespace {

bool hb_kr_persistent_enabled() {
    const char* e = getenv("HB_NO_PERSISTENT");
    return !(e != nullptr && e[0] == '1');
}

struct KrCtx {
    hb_csr* h;
    cudaStream_t s;
    HbKrVec q;
    HbTiles tiles;
    double* part;
    unsigned int* counter;
    double* sc;       // device scalars
    double* h_sc;     // pinned mirror
    double* buf = nullptr;
};

template <class F>
int kr_launch(KrCtx& c, F f) {
    if (c.tiles.n_tiles == 0) return HB_OK;
    hb_kr_vec_kernel<F><<<c.tiles.n_tiles, 256, 0, c.s>>>(c.tiles, f, c.part, c.counter, c.sc);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// out = a o ((A + eps I) u) + b o c  over the block's rows, assembled over the ranks
int kr_spmv(KrCtx& c, const double* u, const double* a, const double* b, const double* cc, double* out, double eps) {
    hb_csr* h = c.h;
    if (h->world > 1 && !h->peer) HB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)h->n_cols, c.s));
    if (h->n_rows > 0 || h->peer) {
        HbAffineEpi epi;
        epi.u = u;
        epi.a = a;
        epi.b = b;
        epi.c = cc;
        epi.eps = eps;
        HbPeerOut po = hb_next_out(h, out);
        hb_launch_spmv<HbAffineEpi, false>(h, u, nullptr, epi, po, c.s);
        HB_LAUNCH_CHECK();
        if (h->peer) {
            // rows were stored into every replica over NVLink: wait for all arrival flags, then take the local copy
            const int pb = (int)(po.seq & 1ull);
            hb_peer_wait_kernel<<<1, 32, 0, c.s>>>(hb_sym_flags(h, pb), po.seq, h->world, nullptr, c.sc + SC_ERR);
            HB_LAUNCH_CHECK();
            HB_CUDA(cudaMemcpyAsync(out, hb_sym_buf(h, pb), sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToDevice, c.s));
        }
    }
    if (h->world > 1 && !h->peer) {
        if (h->allreduce == nullptr || h->allreduce(h->comm_ctx, out, h->n_cols, (void*)c.s) != 0) {
            hb_set_error("collective callback failed");
            return HB_ERR_CUDA;
        }
    }
    return HB_OK;
}

int kr_read_scalars(KrCtx& c) {
    HB_CUDA(cudaMemcpyAsync(c.h_sc, c.sc, sizeof(double) * SC_WORDS, cudaMemcpyDeviceToHost, c.s));
    HB_CUDA(cudaStreamSynchronize(c.s));
    if (c.h_sc[SC_ERR] != 0.0) {
        hb_set_error("peer exchange timed out: a rank did not publish its rows");
        return HB_ERR_CUDA;
    }
    return HB_OK;
}

}  // namespace

// Knight-Ruiz workspace of a handle: 8 n-vectors, the per-tile partials of every reduction, the scalars.
int hb_kr_reserve(hb_csr* h) {
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    const size_t n = (size_t)(h->n_cols > 0 ? h->n_cols : 1);
    const size_t words = 8 * n + 3 * KP_ARRAYS * (size_t)(h->n_tiles > 0 ? h->n_tiles : 1) + SC_WORDS;
    if (h->kr_buf_words < words) {
        hb_dfree(h->d_kr_buf);
        h->d_kr_buf = nullptr;
        h->kr_buf_words = 0;
        cudaError_t e = hb_dmalloc(&h->d_kr_buf, sizeof(double) * words);
        if (e != cudaSuccess) return hb_cuda_fail(e, "cudaMalloc KR vectors", __FILE__, __LINE__);
        h->kr_buf_words = words;
    }
    return HB_OK;
}

#define KR_TRY(expr)            \
    do {                        \
        rc = (expr);            \
        if (rc != HB_OK) goto done; \
    } while (0)

extern "C" {

int hb_dev_spmv(hb_csr* h, const double* d_u, double eps, const double* d_a, const double* d_b, const double* d_c,
                double* d_out, int zero_rest, void* stream) {
    HB_REQUIRE(h != nullptr && d_u != nullptr && d_out != nullptr, "bad arguments");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (zero_rest) HB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)h->n_cols, s));
    if (h->n_rows == 0) return HB_OK;
    HbAffineEpi epi;
    epi.u = d_u;
    epi.a = d_a;
    epi.b = d_b;
    epi.c = d_c;
    epi.eps = eps;
    HbPeerOut po;
    memset(&po, 0, sizeof(po));
    po.world = 1;
    po.out[0] = d_out;
    po.n = h->n_cols;
    hb_launch_spmv<HbAffineEpi, false>(h, d_u, nullptr, epi, po, s);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

int hb_kr_run(hb_csr* h, double tol, double delta, double Delta, double diag_eps, int max_outer, double* x_out,
              int32_t* outer_out, int32_t* matvecs_out, double* residual_out) {
    HB_RANGE("hb_kr_run");
    HB_REQUIRE(h != nullptr && x_out != nullptr, "bad arguments");
    HB_REQUIRE(h->world > 1 || (h->row0 == 0 && h->n_rows == h->n_cols),
               "hb_kr_run needs the full matrix on one device unless hb_set_comm was called");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    const int64_t n = h->n_cols;
    if (n == 0) {
        if (outer_out) *outer_out = 0;
        if (matvecs_out) *matvecs_out = 0;
        if (residual_out) *residual_out = 0.0;
        return HB_OK;
    }
    // krbalancing treats the whole matrix as one problem; per-chromosome runs are separate objects
    HB_REQUIRE(h->nseg == 1, "hb_kr_run works on a single segment (create one store per chromosome block)");
    KrCtx c;
    c.h = h;
    c.s = h->stream;
    c.tiles.row0 = h->d_tile_row0;
    c.tiles.len = h->d_tile_len;
    c.tiles.n_tiles = h->n_tiles;
    c.counter = h->d_counters + 2;
    c.h_sc = nullptr;
    c.sc = nullptr;
    c.part = nullptr;
    const double g = 0.9, etamax = 0.1;
    double eta = etamax, stop_tol = tol * 0.5, rt = tol * tol;
    double rho_km1 = 0.0, rout = 0.0, rold = 0.0;
    int outer = 0, mvp = 0;
    bool breakdown = false;
    rc = hb_kr_reserve(h);   // the workspace lives in the handle: repeated calls (and --perchr loops) pay no allocation
    if (rc) return rc;
    c.buf = h->d_kr_buf;
    c.h_sc = h->h_kr_sc;
    {
        double* b = c.buf;
        c.q.x = b; b += n;
        c.q.y = b; b += n;
        c.q.p = b; b += n;
        c.q.z = b; b += n;
        c.q.rk = b; b += n;
        c.q.v = b; b += n;
        c.q.w = b; b += n;
        c.q.u = b; b += n;
        c.part = b; b += 3 * KP_ARRAYS * h->n_tiles;
        c.sc = b;
    }
    rc = hb_loop_settle(h);
    if (rc) return rc;
    {
        cudaError_t e = cudaMemsetAsync(c.sc, 0, sizeof(double) * SC_WORDS, c.s);
        if (e == cudaSuccess) e = cudaMemsetAsync(c.counter, 0, sizeof(unsigned int), c.s);
        if (e != cudaSuccess) {
            rc = hb_cuda_fail(e, "memset", __FILE__, __LINE__);
            goto done;
        }
    }
    if ((h->world == 1 || h->peer) && hb_kr_persistent_enabled()) {
        // the whole solve is one cooperative launch; the host waits once
        HbKrSolveArgs a;
        memset(&a, 0, sizeof(a));
        a.indptr = h->d_indptr;
        a.idx = h->d_indices;
        a.val = h->pack_kind ? h->d_pack : (const void*)h->d_data;
        a.n_rows = h->n_rows;
        a.row0 = h->row0;
        a.n_cols = n;
        a.tile_row0 = h->d_tile_row0;
        a.tile_len = h->d_tile_len;
        a.n_tiles = h->n_tiles;
        a.x = c.q.x;
        a.y = c.q.y;
        a.p = c.q.p;
        a.rk = c.q.rk;
        a.v = c.q.v;
        a.w = c.q.w;
        a.u = c.q.u;
        a.part = c.part;
        a.sc = c.sc;
        a.h_loop = h->h_loop_dev;
        a.row_counters = h->d_row_counter;
        a.err = h->d_loop_err;
        a.tol = tol;
        a.delta = delta;
        a.Delta = Delta;
        a.eps = diag_eps;
        a.max_outer = max_outer;
        a.row_parity = h->spmv_parity;
        hb_fill_loop_peer(h, &a.peer);
        const void* fn = h->pack_kind == 1   ? (const void*)hb_kr_solve_kernel<uint16_t>
                         : h->pack_kind == 2 ? (const void*)hb_kr_solve_kernel<float>
                                             : (const void*)hb_kr_solve_kernel<double>;
        const int grid = hb_coop_grid(fn, h->n_rows, h->device);
        void* params[] = {(void*)&a};
        for (int i = 0; i < HB_LOOP_WORDS; ++i) ((volatile int32_t*)h->h_loop)[i] = 0;
        cudaError_t e = cudaMemsetAsync(h->d_loop_wmax, 0, 24, c.s);   // wmax[2] | err
        if (e == cudaSuccess)
            e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(HB_SPMV_THREADS), params, 0, c.s);
        g_hb_launches.fetch_add(1, std::memory_order_relaxed);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c.h_sc, c.sc, sizeof(double) * SC_WORDS, cudaMemcpyDeviceToHost, c.s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.s);
        if (e != cudaSuccess) {
            rc = hb_cuda_fail(e, "KR solve kernel", __FILE__, __LINE__);
            goto done;
        }
        const volatile int32_t* lp = (const volatile int32_t*)h->h_loop;
        if (h->pc != nullptr) h->pc->seq += (unsigned long long)(unsigned int)lp[HB_LP_EXCHANGES];
        outer = lp[HB_LP_OUTER];
        mvp = lp[HB_LP_MATVECS];
        rout = c.h_sc[SC_ROUT];
        if (lp[HB_LP_FLAGS] == 2) {
            hb_set_error("peer exchange timed out: a rank did not publish its rows");
            rc = HB_ERR_CUDA;
            goto done;
        }
        breakdown = lp[HB_LP_FLAGS] == 1;
    } else {
    {
        FInit f0;
        f0.q = c.q;
        KR_TRY(kr_launch(c, f0));
        // v = x o ((A + eps I) x) ; rk = 1 - v ; rho = rk^T rk
        KR_TRY(kr_spmv(c, c.q.x, c.q.x, nullptr, nullptr, c.q.v, diag_eps));
        FResid fr;
        fr.q = c.q;
        KR_TRY(kr_launch(c, fr));
        KR_TRY(kr_read_scalars(c));
        rho_km1 = c.h_sc[SC_RHO];
        rout = rold = rho_km1;
    }
    while (rout > rt && outer < max_outer) {
        ++outer;
        int k = 0;
        const double innertol = fmax(eta * eta * rout, rt);
        while (rho_km1 > innertol) {
            ++k;
            if (k == 1) {
                FFirst f;
                f.q = c.q;
                KR_TRY(kr_launch(c, f));
            } else {
                FDir f;
                f.q = c.q;
                f.beta = 0.0;
                KR_TRY(kr_launch(c, f));
            }
            // w = x o ((A + eps I)(x o p)) + v o p
            KR_TRY(kr_spmv(c, c.q.u, c.q.x, c.q.v, c.q.p, c.q.w, diag_eps));
            {
                FDot fd;
                fd.q = c.q;
                KR_TRY(kr_launch(c, fd));
                FYnew fy;
                fy.q = c.q;
                fy.alpha = 0.0;
                KR_TRY(kr_launch(c, fy));
                FStep fs;
                fs.q = c.q;
                fs.delta = delta;
                fs.Delta = Delta;
                fs.alpha = 0.0;
                KR_TRY(kr_launch(c, fs));
            }
            KR_TRY(kr_read_scalars(c));
            if (c.h_sc[SC_MIN] <= delta) {
                if (delta == 0.0) break;
                FGamma fg;
                fg.q = c.q;
                fg.bound = delta;
                fg.upper = 0;
                fg.alpha = 0.0;
                KR_TRY(kr_launch(c, fg));
                FBoundStep fb;
                fb.q = c.q;
                fb.alpha = fb.gamma = 0.0;
                KR_TRY(kr_launch(c, fb));
                break;
            }
            if (c.h_sc[SC_MAX] >= Delta) {
                FGamma fg;
                fg.q = c.q;
                fg.bound = Delta;
                fg.upper = 1;
                fg.alpha = 0.0;
                KR_TRY(kr_launch(c, fg));
                FBoundStep fb;
                fb.q = c.q;
                fb.alpha = fb.gamma = 0.0;
                KR_TRY(kr_launch(c, fb));
                break;
            }
            rho_km1 = c.h_sc[SC_RHO];
            if (!(rho_km1 == rho_km1)) {
                breakdown = true;
                break;
            }
        }
        if (breakdown) break;
        {
            FXUpdate fx;
            fx.q = c.q;
            KR_TRY(kr_launch(c, fx));
            KR_TRY(kr_spmv(c, c.q.x, c.q.x, nullptr, nullptr, c.q.v, diag_eps));
            FResid fr;
            fr.q = c.q;
            KR_TRY(kr_launch(c, fr));
            KR_TRY(kr_read_scalars(c));
        }
        rho_km1 = c.h_sc[SC_RHO];
        rout = c.h_sc[SC_ROUT];
        mvp += k + 1;
        if (!(rout == rout) || !(c.h_sc[SC_XMIN] > 0.0)) {
            breakdown = true;
            break;
        }
        const double rat = rout / rold;
        rold = rout;
        const double res_norm = sqrt(rout);
        const double eta_o = eta;
        eta = g * rat;
        if (g * eta_o * eta_o > 0.1) eta = fmax(eta, g * eta_o * eta_o);
        eta = fmax(fmin(eta, etamax), stop_tol / res_norm);
    }
    }   // step-level path
    {
        cudaError_t e = cudaMemcpyAsync(x_out, c.q.x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c.s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.s);
        if (e != cudaSuccess) rc = hb_cuda_fail(e, "KR result copy", __FILE__, __LINE__);
    }
    if (outer_out) *outer_out = outer;
    if (matvecs_out) *matvecs_out = mvp;
    if (residual_out) *residual_out = sqrt(rout);
    if (rc == HB_OK && breakdown) {
        hb_set_error("Knight-Ruiz breakdown (non-finite residual or non-positive scaling vector)");
        rc = HB_ERR_KR_BREAKDOWN;
    }
done:
    cudaStreamSynchronize(c.s);
    return rc;
}

int hb_kr_rescale(hb_csr* h, double diag_eps, double* x_inout, double* factor_out) {
    HB_RANGE("hb_kr_rescale");
    HB_REQUIRE(h != nullptr && x_inout != nullptr, "bad arguments");
    HB_REQUIRE(h->world > 1 || (h->row0 == 0 && h->n_rows == h->n_cols),
               "hb_kr_rescale needs the full matrix on one device unless hb_set_comm was called");
    int rc = hb_ensure_state(h);
    if (rc) return rc;
    rc = hb_use_device(h);
    if (rc) return rc;
    const int64_t n = h->n_cols;
    if (n == 0) return HB_OK;
    HB_REQUIRE(h->world == 1 || h->allreduce != nullptr, "sharded hb_kr_rescale needs the allreduce callback (hb_set_comm)");
    cudaStream_t s = h->stream;
    double* buf = nullptr;
    HB_CUDA(hb_dmalloc(&buf, sizeof(double) * (size_t)(3 * n + 3 * h->n_tiles + SC_WORDS)));
    double *d_x = buf, *d_scaled = buf + n, *d_orig = buf + 2 * n, *d_part = buf + 3 * n, *d_sc = d_part + 3 * h->n_tiles;
    double sc[SC_WORDS];
    cudaError_t e = cudaMemcpyAsync(d_x, x_inout, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->d_counters + 2, 0, sizeof(unsigned int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_scaled, 0, sizeof(double) * (size_t)(2 * n), s);
    if (e == cudaSuccess) {
        if (h->n_rows > 0) {
            hb_kr_rescale_rows_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, h->n_rows,
                                                                     h->row0, d_x, diag_eps, d_scaled, d_orig);
            g_hb_launches.fetch_add(1);
        }
        // sharded: every rank filled its rows of (scaled | orig); one sum-allreduce assembles both vectors
        if (h->world > 1 && h->allreduce(h->comm_ctx, d_scaled, 2 * n, (void*)s) != 0) {
            hb_dfree(buf);
            hb_set_error("collective callback failed");
            return HB_ERR_CUDA;
        }
        HbTiles tiles{h->d_tile_row0, h->d_tile_len, h->n_tiles};
        for (int which = 0; which < 2; ++which) {
            FSum2 f;
            f.a = d_scaled;
            f.b = d_orig;
            f.which = which;
            hb_kr_vec_kernel<FSum2><<<h->n_tiles, 256, 0, s>>>(tiles, f, d_part, h->d_counters + 2, d_sc);
            g_hb_launches.fetch_add(1);
        }
        e = cudaMemcpyAsync(sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    hb_dfree(buf);
    if (e != cudaSuccess) return hb_cuda_fail(e, "KR rescale", __FILE__, __LINE__);
    const double factor = sqrt(sc[SC_A] / sc[SC_B]);
    for (int64_t i = 0; i < n; ++i) x_inout[i] /= factor;
    if (factor_out) *factor_out = factor;
    return HB_OK;
}

int hb_kr_scaled_upper(hb_csr* h, double diag_eps, const double* x, int64_t* nnz_out, int64_t* indptr, int32_t* indices,
                       double* data) {
    HB_RANGE("hb_kr_scaled_upper");
    HB_REQUIRE(h != nullptr && x != nullptr && nnz_out != nullptr, "bad arguments");
    // row-local: on a row block the output is the block's rows (indptr[n_rows+1] block-local, global column ids)
    int rc = hb_use_device(h);
    if (rc) return rc;
    const int64_t n = h->n_rows;
    const int64_t nx = h->n_cols;
    cudaStream_t s = h->stream;
    int64_t *d_cnt = nullptr, *d_ptr = nullptr;
    HbScratch scratch;   // freed on every return path
    HB_CUDA(scratch.alloc(&d_cnt, sizeof(int64_t) * (size_t)(n + 1)));
    HB_CUDA(scratch.alloc(&d_ptr, sizeof(int64_t) * (size_t)(n + 1)));
    if (n > 0) {
        hb_kr_upper_count_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, n, h->row0, d_cnt);
        g_hb_launches.fetch_add(1);
    }
    hb_scan_kernel<int64_t><<<1, 1024, 0, s>>>(d_cnt, d_ptr, n);
    g_hb_launches.fetch_add(1);
    int64_t total = 0;
    cudaError_t e = cudaMemcpyAsync(&total, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    *nnz_out = total;
    if (e == cudaSuccess && indptr != nullptr && indices != nullptr && data != nullptr && n > 0) {
        double *d_x = nullptr, *d_val = nullptr;
        int32_t* d_idx = nullptr;
        e = scratch.alloc(&d_x, sizeof(double) * (size_t)nx);
        if (e == cudaSuccess) e = scratch.alloc(&d_val, sizeof(double) * (size_t)total);
        if (e == cudaSuccess) e = scratch.alloc(&d_idx, sizeof(int32_t) * (size_t)total);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)nx, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            hb_kr_upper_fill_kernel<<<HB_NUM_SMS * 8, 256, 0, s>>>(h->d_indptr, h->d_indices, h->d_data, n, h->row0, d_x,
                                                                   diag_eps, d_ptr, d_idx, d_val);
            g_hb_launches.fetch_add(1);
            e = cudaStreamSynchronize(s);
        }
        if (e == cudaSuccess) e = cudaMemcpy(indptr, d_ptr, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost);
        // large, caller-owned pageable buffers: multi-threaded staged copies
        if (e == cudaSuccess && total > 0 && hb_copy_d2h(indices, d_idx, sizeof(int32_t) * (size_t)total, h->device) != HB_OK)
            return HB_ERR_CUDA;
        if (e == cudaSuccess && total > 0 && hb_copy_d2h(data, d_val, sizeof(double) * (size_t)total, h->device) != HB_OK)
            return HB_ERR_CUDA;
    }
    if (e != cudaSuccess) return hb_cuda_fail(e, "KR scaled upper", __FILE__, __LINE__);
    return HB_OK;
}

}  // extern "C"
