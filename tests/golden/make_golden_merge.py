"""Golden vectors for bin merging (build container only; needs /root/reference):
small_test_matrix.h5 -> hicMergeMatrixBins --numBins 5 -> test_data/hicMergeMatrixBins/result.h5.
    python tests/golden/make_golden_merge.py   ->  tests/golden/merge_small.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from hicexplorer_b200 import hdf5lite  # noqa: E402

TD = "/root/reference/hicexplorer/test/test_data/"


def read(path):
    f = hdf5lite.File(path)
    out = {"data": f["matrix/data"].read(), "indices": f["matrix/indices"].read(), "indptr": f["matrix/indptr"].read(),
           "chr": f["intervals/chr_list"].read(), "start": f["intervals/start_list"].read(),
           "end": f["intervals/end_list"].read(), "extra": f["intervals/extra_list"].read()}
    if "nan_bins" in f:
        out["nan_bins"] = f["nan_bins"].read()
    return out


src = read(TD + "small_test_matrix.h5")
res = read(TD + "hicMergeMatrixBins/result.h5")
np.savez_compressed(os.path.join(HERE, "merge_small.npz"), **{"in_" + k: v for k, v in src.items()},
                    **{"out_" + k: v for k, v in res.items()})
print(os.path.getsize(os.path.join(HERE, "merge_small.npz")), {k: v.shape for k, v in res.items()})
