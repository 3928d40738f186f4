"""Golden vectors for hicNormalize from the reference's own test files (build container only):
    python tests/golden/make_golden_normalize.py
-> tests/golden/normalize_small.npz: the two input matrices of test_hicNormalize.py (upper triangles as stored) and
the data/indices/indptr of the reference's golden outputs for --normalize smallest, norm_range and multiplicative."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hicexplorer_b200 import hdf5lite  # noqa: E402
from hicexplorer_b200.hicmatrix_lite import hiCMatrix  # noqa: E402

TD = "/root/reference/hicexplorer/test/test_data/hicNormalize/"


def h5(name):
    f = hdf5lite.File(TD + name)
    return f["matrix/data"].read(), f["matrix/indices"].read(), f["matrix/indptr"].read()


out = {}
f = hdf5lite.File(TD + "small_test_matrix.h5")
out["n"] = int(f["matrix/shape"].read()[0])
out["chr"] = f["intervals/chr_list"].read()
out["start"] = f["intervals/start_list"].read()
out["end"] = f["intervals/end_list"].read()
out["extra"] = f["intervals/extra_list"].read()
for key, name in [("one", "small_test_matrix.h5"), ("two", "small_test_matrix_scaled_up.h5"),
                  ("smallest_one", "smallest_one.h5"), ("smallest_two", "smallest_two.h5"),
                  ("norm_range_one", "norm_range_one.h5"), ("norm_range_two", "norm_range_two.h5")]:
    d, i, p = h5(name)
    out[key + "_data"], out[key + "_indices"], out[key + "_indptr"] = d, i.astype(np.int32), p.astype(np.int64)
by2 = hiCMatrix(TD + "small_test_matrix_scaled_by_2.cool")
up = by2.upper_for_device()
out["by2_data"], out["by2_indices"], out["by2_indptr"] = up.data, up.indices.astype(np.int32), up.indptr.astype(np.int64)
np.savez_compressed(os.path.join(HERE, "normalize_small.npz"), **out)
print({k: (v.shape, v.dtype) if hasattr(v, "shape") else v for k, v in out.items()})
