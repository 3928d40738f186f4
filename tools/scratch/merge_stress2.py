import sys, os, tempfile, pathlib
sys.path.insert(0, os.getcwd())
import numpy as np
from hicexplorer_b200.store import DeviceCSR
from oracle import merge_oracle
from tests.helpers import random_hic
from tests import test_gpu_merge as T
tmp = pathlib.Path(tempfile.mkdtemp())
T.test_merge_matrix_bins_matches_reference_golden(tmp)
n, k, seed = 500, 2, 0
m = random_hic(n, 0.1, seed, empty_frac=0.02)
sizes = [n // 3, n // 3, n - 2 * (n // 3)]
iv = []
for c, L in enumerate(sizes):
    iv += [("chr%d" % c, i * 100, (i + 1) * 100, 1.0) for i in range(L)]
ref, new_iv, _nan, bin_map = merge_oracle.merge_bins(m, iv, k)
with DeviceCSR.from_scipy(m) as dev:
    dev.merge_bins(bin_map, len(new_iv))
    got = dev.download()
    r = dev.symmetry_ratio()
    r2 = dev.symmetry_ratio()
    again = dev.download()
same = np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices) and np.array_equal(got.data, ref.data)
same2 = np.array_equal(again.indptr, ref.indptr) and np.array_equal(again.indices, ref.indices) and np.array_equal(again.data, ref.data)
sym_host = abs(got - got.T).sum()
print("RESULT same=%s same_again=%s ratio=%r ratio2=%r host_asym=%r" % (same, same2, r, r2, sym_host), flush=True)
