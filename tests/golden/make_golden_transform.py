"""Golden vectors for hicTransform's obs/exp methods (build container only):
    python tests/golden/make_golden_transform.py
-> tests/golden/transform_small.npz.  Inputs: the reference's test matrices small_test_matrix.cool (the same matrix as
`one_*` in normalize_small.npz; only its chromosome offsets are stored here) and small_test_matrix_50kb_res.h5.
Expected outputs: (1) what the reference's OWN functions (hicexplorer.hicTransform._obs_exp & co., imported unmodified
through oracle.ref_import) return on those inputs with the numpy installed here, zeros eliminated, upper triangle;
(2) the data of the reference's golden files test_data/hicTransform/*.cool|h5 (written by an older code version; the
reference's tests compare them with decimal=0)."""
import os
import sys
import warnings

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hicexplorer_b200.hicmatrix_lite import hiCMatrix  # noqa: E402
from oracle import ref_import  # noqa: E402

warnings.simplefilter("ignore")
ref_import.load()
import hicexplorer.hicTransform as ht  # noqa: E402

TD = "/root/reference/hicexplorer/test/test_data/"


def upper(m):
    m = sparse.csr_matrix(m)
    m.eliminate_zeros()
    up = sparse.triu(m, 0, format="csr")
    up.sort_indices()
    return up


def per_chromosome(fn, m, offs):
    blocks = []
    for lo, hi in zip(offs[:-1], offs[1:]):
        sub = m[lo:hi, lo:hi].tocsr()
        blocks.append(sparse.csr_matrix(fn(sub)) if sub.nnz else sub)
    return sparse.block_diag(blocks, format="csr")


out = {}
a = hiCMatrix(TD + "small_test_matrix.cool")
m = a.matrix
offs = a.chromosome_offsets()
out["offsets"] = offs
b = hiCMatrix(TD + "small_test_matrix_50kb_res.h5")
ub = upper(b.matrix)
out["b_n"] = b.matrix.shape[0]
out["b_data"], out["b_indices"], out["b_indptr"] = ub.data, ub.indices.astype(np.int32), ub.indptr.astype(np.int64)
out["b_offsets"] = b.chromosome_offsets()
out["b_chr"] = np.asarray([iv[0].encode() for iv in b.cut_intervals])
out["b_start"] = np.asarray([iv[1] for iv in b.cut_intervals], dtype=np.int64)
out["b_end"] = np.asarray([iv[2] for iv in b.cut_intervals], dtype=np.int64)
mb = b.matrix
n_b, nchr_b = int(out["b_offsets"][-1]), len(out["b_offsets"]) - 1
cases = {
    "ref_obs_exp": ht._obs_exp(m.copy()),
    "ref_obs_exp_pc": per_chromosome(ht._obs_exp, m, offs),
    "ref_non_zero_pc": per_chromosome(lambda s: ht._obs_exp_non_zero(s, False), m, offs),
    "ref_lieberman_b": per_chromosome(lambda s: ht._obs_exp_lieberman(s, n_b, nchr_b), mb, out["b_offsets"]),
    "ref_non_zero_lig_b": ht._obs_exp_non_zero(mb.copy(), True),
    "ref_non_zero_lig_pc_b": per_chromosome(lambda s: ht._obs_exp_non_zero(s, True), mb, out["b_offsets"]),
}
for key, res in cases.items():
    up = upper(res)
    out[key + "_data"], out[key + "_indices"], out[key + "_indptr"] = up.data, up.indices.astype(np.int32), up.indptr.astype(np.int64)
for key, name in [("file_obs_exp", "hicTransform/obs_exp.cool"), ("file_obs_exp_pc", "hicTransform/obs_exp_per_chromosome.cool"),
                  ("file_non_zero_pc", "hicTransform/obs_exp_non_zero_per_chromosome.cool"),
                  ("file_lieberman_b", "hicTransform/obs_exp_lieberman.h5")]:
    up = upper(hiCMatrix(TD + name).matrix)
    out[key + "_data"], out[key + "_indices"], out[key + "_indptr"] = up.data, up.indices.astype(np.int32), up.indptr.astype(np.int64)
np.savez_compressed(os.path.join(HERE, "transform_small.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items() if hasattr(v, "shape")})
