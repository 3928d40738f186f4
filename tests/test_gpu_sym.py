"""Device symmetrisation (upper triangle -> full store) and the corrected matrix in file layout."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR
from oracle import ice_oracle
from tests.helpers import full_from_pixels, random_hic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,density,seed", [(300, 0.2, 0), (2000, 0.05, 1), (50, 0.9, 2), (5000, 0.3, 3), (1, 1.0, 4)])
def test_from_upper_equals_full_upload(n, density, seed):
    full = random_hic(n, density, seed, empty_frac=0.02)
    up = sparse.triu(full, 0, format="csr")
    with DeviceCSR.from_upper(up) as dev:
        got = dev.download()
        assert dev.symmetry_ratio() == 0.0
    assert np.array_equal(got.indptr, full.indptr)
    assert np.array_equal(got.indices, full.indices)          # ascending columns, deterministic layout
    np.testing.assert_array_equal(got.data, full.data)


def test_from_upper_rejects_lower_entries():
    full = random_hic(100, 0.2, 5)
    with pytest.raises(ValueError, match="below the diagonal"):
        DeviceCSR.from_upper(full)


def test_from_upper_then_ice_is_bit_identical(gm12878):
    full = full_from_pixels(gm12878)
    with DeviceCSR.from_scipy(full) as a, DeviceCSR.from_upper(sparse.triu(full, 0, format="csr")) as b:
        ba, ia, _ = a.ice(20, 0.0)
        bb, ib, _ = b.ice(20, 0.0)
    np.testing.assert_array_equal(ba, bb)


def test_scaled_upper_is_the_file_layout(gm12878):
    from hicexplorer_b200.pipeline import ice_pipeline, scatter_to_frame
    full = full_from_pixels(gm12878)
    ref = ice_pipeline(full, (-1.5, 5.0), max_iter=500)
    want = sparse.triu(scatter_to_frame(ref["matrix"], ref["kept"], ref["n"]), 0, format="csr")
    want.eliminate_zeros()
    want.sort_indices()
    n = full.shape[0]
    keep = np.zeros(n, dtype=bool)
    keep[ref["kept"]] = True
    with DeviceCSR.from_scipy(full) as dev:
        dev.compact((~keep).astype(np.uint8))
        dev.scrub()
        dev.ice(500, 1e-5)
        got = dev.ice_scaled_upper(ref["kept"], n)
    assert got.shape == (n, n)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    oracle = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=500)
    np.testing.assert_allclose(got.data, oracle["upper"].data, rtol=1e-8)


def test_upload_round_trip_is_exact_every_time():
    """regression: the alignment-slot memset of hb_csr_create ran on the handle's non-blocking stream and could land
    after the (default-stream) copy of the entries, wiping the first four of them in a few percent of fresh processes"""
    m = random_hic(400, 0.1, 9)
    for _ in range(200):
        with DeviceCSR.from_scipy(m) as dev:
            got = dev.download()
        assert np.array_equal(got.indices[:8], m.indices[:8]) and np.array_equal(got.data[:8], m.data[:8])
    assert np.array_equal(got.indices, m.indices) and np.array_equal(got.data, m.data)


def test_malformed_blocks_are_rejected_at_upload():
    """the C ABI validates what it is handed (a raw ctypes caller may skip scipy's canonical form)"""
    import ctypes
    L = _lib.lib()

    def create(indptr, indices, data, n):
        ip = np.asarray(indptr, dtype=np.int64)
        ix = np.asarray(indices, dtype=np.int32)
        da = np.asarray(data, dtype=np.float64)
        h = ctypes.c_void_p()
        rc = L.hb_csr_create(ctypes.byref(h), n, n, 0, len(ix), _lib.ptr(ip), _lib.ptr(ix), 0, _lib.ptr(da), 0)
        if rc == 0:
            L.hb_csr_destroy(h)
        return rc, _lib.last_error()

    assert create([0, 2, 3, 4], [0, 2, 1, 0], [1, 2, 3, 4], 3)[0] == 0
    rc, msg = create([0, 2, 3, 4], [0, 3, 1, 0], [1, 2, 3, 4], 3)
    assert rc == _lib.HB_ERR_INVALID_ARG and "outside" in msg
    rc, msg = create([0, 2, 3, 4], [0, -1, 1, 0], [1, 2, 3, 4], 3)
    assert rc == _lib.HB_ERR_INVALID_ARG and "outside" in msg
    rc, msg = create([0, 2, 3, 4], [2, 0, 1, 0], [1, 2, 3, 4], 3)
    assert rc == _lib.HB_ERR_INVALID_ARG and "ascending" in msg
    rc, msg = create([0, 2, 2, 4], [1, 1, 0, 2], [1, 2, 3, 4], 3)           # duplicate column in a row
    assert rc == _lib.HB_ERR_INVALID_ARG and "ascending" in msg
    rc, msg = create([0, 3, 2, 4], [0, 1, 2, 0], [1, 2, 3, 4], 3)
    assert rc == _lib.HB_ERR_INVALID_ARG and "monotone" in msg


def test_from_upper_with_rows_beyond_the_shared_memory_classes():
    """dense columns (coarse genome-wide matrices): lower parts of more than 16384 entries take the global-scratch sort"""
    n = 40000
    rng = np.random.default_rng(3)
    dense_cols = np.array([n - 1, n - 7, 20000])
    rows = np.concatenate([np.arange(c + 1) for c in dense_cols])
    cols = np.concatenate([np.full(c + 1, c) for c in dense_cols])
    extra = rng.integers(0, n, size=(2, 60000))
    r = np.concatenate([rows, np.minimum(extra[0], extra[1])])
    c = np.concatenate([cols, np.maximum(extra[0], extra[1])])
    v = rng.integers(1, 100, size=len(r)).astype(np.float64)
    up = sparse.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    up.sum_duplicates()
    full = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    full.sort_indices()
    assert np.diff(sparse.tril(full, -1).tocsr().indptr).max() > 32768        # two size classes of the big-row path
    with DeviceCSR.from_upper(up) as dev:
        got = dev.download()
    assert np.array_equal(got.indptr, full.indptr) and np.array_equal(got.indices, full.indices)
    np.testing.assert_array_equal(got.data, full.data)


def test_keep_rows_equals_a_row_block_upload():
    """sharded runs build their block on the device: upper triangle -> symmetric store -> hb_csr_keep_rows"""
    full = random_hic(1500, 0.08, 17, empty_frac=0.01)
    up = sparse.triu(full, 0, format="csr")
    for a, b in [(0, 400), (400, 1100), (1100, 1500), (700, 700)]:
        with DeviceCSR.from_upper(up) as dev, DeviceCSR.from_scipy(full, row_range=(a, b)) as ref:
            dev.keep_rows(a, b)
            assert dev.dims() == ref.dims()
            got, want = dev.download(), ref.download()
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            np.testing.assert_array_equal(got.data, want.data)
            rs_got, dg_got = dev.row_stats()
            rs_want, dg_want = ref.row_stats()
            np.testing.assert_array_equal(rs_got, rs_want)
            np.testing.assert_array_equal(dg_got, dg_want)


def _sym_case(n, seed, band, per_row, far):
    """upper triangle with a band of `per_row` entries per row within `band` columns plus `far` contacts anywhere"""
    rng = np.random.default_rng(seed)
    r = np.repeat(np.arange(n), per_row)
    c = np.minimum(r + rng.integers(0, band, size=len(r)), n - 1)
    fr = rng.integers(0, n, size=far)
    fc = rng.integers(0, n, size=far)
    r = np.concatenate([r, np.minimum(fr, fc)])
    c = np.concatenate([c, np.maximum(fr, fc)])
    v = rng.integers(1, 1000, size=len(r)).astype(np.float64)
    up = sparse.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    up.sum_duplicates()
    return up


@pytest.mark.parametrize("n,band,per_row,far", [(150000, 3000, 6, 40000),       # keys of one row span several bitmap windows
                                                 (70000, 66000, 3, 0),           # a band wider than one window
                                                 (9000, 9000, 700, 0)])          # every size class up to 8192 entries
def test_rank_ordering_of_the_mirror_image(monkeypatch, n, band, per_row, far):
    """[r2] the lower parts are ordered by bitmap rank (hb_sym_rank_kernel) instead of a sorting network: same store as
    scipy's transpose, and as the round-1 network (HB_SYM_SORT=network)"""
    up = _sym_case(n, 5, band, per_row, far)
    full = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    full.sort_indices()
    with DeviceCSR.from_upper(up) as dev:
        got = dev.download()
    assert np.array_equal(got.indptr, full.indptr) and np.array_equal(got.indices, full.indices)
    np.testing.assert_array_equal(got.data, full.data)
    monkeypatch.setenv("HB_SYM_SORT", "network")
    with DeviceCSR.from_upper(up) as dev:
        ref = dev.download()
    assert np.array_equal(ref.indices, got.indices) and np.array_equal(ref.data, got.data)
