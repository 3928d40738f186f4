import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from hicexplorer_b200.store import DeviceCSR
from oracle import merge_oracle
from tests.helpers import random_hic
n, k, seed = 500, 2, 0
m = random_hic(n, 0.1, seed, empty_frac=0.02)
sizes = [n // 3, n // 3, n - 2 * (n // 3)]
iv = []
for c, L in enumerate(sizes):
    iv += [("chr%d" % c, i * 100, (i + 1) * 100, 1.0) for i in range(L)]
ref, new_iv, _nan, bin_map = merge_oracle.merge_bins(m, iv, k)
bad = 0
for it in range(400):
    with DeviceCSR.from_scipy(m) as dev:
        dev.merge_bins(bin_map, len(new_iv))
        got = dev.download()
        r = dev.symmetry_ratio()
    same_ptr = np.array_equal(got.indptr, ref.indptr)
    same = same_ptr and np.array_equal(got.indices, ref.indices) and np.array_equal(got.data, ref.data)
    if not same or r != 0.0:
        bad += 1
        msg = "iter %d: same=%s same_ptr=%s ratio=%r nnz %d vs %d" % (it, same, same_ptr, r, got.nnz, ref.nnz)
        if same_ptr and not same:
            d = np.flatnonzero((got.indices != ref.indices) | (got.data != ref.data))
            rows = np.searchsorted(ref.indptr, d, side="right") - 1
            msg += " first diffs at entries %s rows %s got %s want %s" % (d[:5], rows[:5], got.data[d[:5]], ref.data[d[:5]])
        print(msg, flush=True)
print("bad runs:", bad)
