"""Golden vectors for hicTransform --method covariance | pearson (build container only):
    python tests/golden/make_golden_cov.py
-> tests/golden/transform_cov.npz.  Input: small_test_matrix_50kb_res.h5 (stored as `b_*` in transform_small.npz).
Expected outputs: (1) the reference's golden file test_data/hicTransform/covariance_perChromosome.h5 (what
test_hicTransform.py::test_covariance_per_chromosome compares with), (2) pearson per chromosome from the reference's OWN
_pearson (imported unmodified through oracle.ref_import; the reference has no golden file for it).  Each expected matrix
(upper triangle, zeros eliminated) is stored as: its indptr, a sha256 of its column indices, the sum of every row, and
every 16th stored cell -- the full 572 647-cell matrices would be 7 MB each."""
import hashlib
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
STRIDE = 16


def upper(m):
    m = sparse.csr_matrix(m).astype(np.float64)
    m.eliminate_zeros()
    up = sparse.triu(m, 0, format="csr")
    up.sort_indices()
    return up


def pack(out, key, up):
    out[key + "_indptr"] = up.indptr.astype(np.int64)
    out[key + "_indices_sha256"] = np.frombuffer(hashlib.sha256(up.indices.astype(np.int32).tobytes()).digest(), dtype=np.uint8)
    out[key + "_row_sums"] = np.asarray(up.sum(axis=1)).ravel()
    out[key + "_sample"] = up.data[::STRIDE].copy()


out = {"stride": np.int64(STRIDE)}
pack(out, "cov_pc_file", upper(hiCMatrix(TD + "hicTransform/covariance_perChromosome.h5").matrix))

b = hiCMatrix(TD + "small_test_matrix_50kb_res.h5")
offs = b.chromosome_offsets()
blocks = []
for lo, hi in zip(offs[:-1], offs[1:]):
    blocks.append(sparse.csr_matrix(ht._pearson(b.matrix[lo:hi, lo:hi].todense())))
pack(out, "pearson_pc_ref", upper(sparse.block_diag(blocks, format="csr")))
np.savez_compressed(os.path.join(HERE, "transform_cov.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
