"""numpy restatement of the device generator in hicexplorer_b200/csrc/hb_synth.cu
(test infrastructure: lets the CPU reproduce, bit for bit, the synthetic matrix a
GPU generated -- used to cross-check the generator and to feed the CPU oracle
with exactly the benchmark's input at small sizes)."""
import numpy as np
from scipy import sparse

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
TAG_BASE = np.uint64(1 << 40)


def _u64(x):
    return np.asarray(x).astype(np.uint64)


def mix64(z):
    z = _u64(z).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xbf58476d1ce4e5b9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94d049bb133111eb)
        z ^= z >> np.uint64(31)
    return z


def hash3(seed, a, b):
    with np.errstate(over="ignore"):
        z = mix64(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (_u64(a) + np.uint64(1)))
        return mix64(z ^ (np.uint64(0xD1B54A32D192ED03) * (_u64(b) + np.uint64(1))))


def u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def bin_bias(seed, n, low_frac, empty_frac):
    i = np.arange(n, dtype=np.uint64)
    empty = u01(hash3(seed, i, TAG_BASE + np.uint64(0))) < empty_frac
    low = u01(hash3(seed, i, TAG_BASE + np.uint64(1))) < low_frac
    u = u01(hash3(seed, i, TAG_BASE + np.uint64(2)))
    g = 0.2 + 4.8 * ((u * u) * u)
    g[low] = 0.02
    g[empty] = 0.0
    return g


def _count(hp, lam, ga, gb):
    with np.errstate(over="ignore"):
        u2 = u01(mix64(hp + np.uint64(0x9E3779B97F4A7C15)))
    return 1.0 + np.floor(((lam * ga) * gb) * u2)


def synth_matrix(chrom_bins, seed=0, band=2000, d0=200.0, gamma=1.0, lambda0=40.0, n_trans=8, low_frac=0.02,
                 empty_frac=0.005, lambda_trans=2.0):
    """Full symmetric CSR (float64 integer counts), identical to DeviceCSR.synthetic(...)."""
    chrom_bins = np.asarray(chrom_bins, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(chrom_bins)])
    n = int(off[-1])
    g = bin_bias(seed, n, low_frac, empty_frac)
    cs = sorted(set(int(hash3(seed, np.uint64(k), TAG_BASE + np.uint64(3))) % n for k in range(n_trans)))
    rows, cols, vals = [], [], []
    for c in range(len(chrom_bins)):
        lo, hi = int(off[c]), int(off[c + 1])
        for i in range(lo, hi):
            if g[i] == 0.0:
                continue
            blo, bhi = max(i - band, lo), min(i + band, hi - 1)
            j = np.arange(blo, bhi + 1, dtype=np.int64)
            a = np.minimum(i, j)
            b = np.maximum(i, j)
            hp = hash3(seed, a, b)
            d1 = (b - a + 1).astype(np.float64)
            u = u01(hp)
            ex = (u * d1 < d0) if gamma == 1.0 else (u < np.power(d0 / d1, gamma))
            ex &= g[j] != 0.0
            j, a, b, hp = j[ex], a[ex], b[ex], hp[ex]
            lam = lambda0 / (1.0 + (b - a).astype(np.float64) / d0)
            v = _count(hp, lam, g[a], g[b])
            # far partners
            tj = np.array([(ck - i) % n for ck in cs], dtype=np.int64)
            keep = ((tj < blo) | (tj > bhi)) & (g[tj] != 0.0) if len(tj) else np.zeros(0, bool)
            tj = np.sort(tj[keep])
            if len(tj):
                ta, tb = np.minimum(i, tj), np.maximum(i, tj)
                tv = _count(hash3(seed, ta, tb), lambda_trans, g[ta], g[tb])
            else:
                tv = np.zeros(0)
            before = tj < blo
            jj = np.concatenate([tj[before], j, tj[~before]])
            vv = np.concatenate([tv[before], v, tv[~before]])
            rows.append(np.full(len(jj), i, dtype=np.int64))
            cols.append(jj)
            vals.append(vv)
    if rows:
        r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    else:
        r = c = np.zeros(0, np.int64)
        v = np.zeros(0)
    m = sparse.csr_matrix((v, (r, c)), shape=(n, n))
    m.sort_indices()
    return m
