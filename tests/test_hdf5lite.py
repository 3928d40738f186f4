"""File-format decoders: the C implementation (csrc/hb_io.c -> libhbio.so) and the pure-Python one in hdf5lite.py
must agree, on a hand-built Blosc/blosclz chunk and -- in the build container -- on the reference's own fixtures."""
import os
import struct

import numpy as np
import pytest

from hicexplorer_b200 import hdf5lite

TD = "/root/reference/hicexplorer/test/test_data/"


def _blosc_chunk(stream, nbytes, typesize=1, shuffle=False):
    """one-block Blosc-1 container around a blosclz stream"""
    flags = (0x1 if shuffle else 0)            # codec 0 = blosclz
    header = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, nbytes, 16 + 4 + 4 + len(stream))
    return header + struct.pack("<i", 20) + struct.pack("<i", len(stream)) + stream


def _decode_both(chunk):
    os.environ["HDF5LITE_PURE_PYTHON"] = "1"
    hdf5lite._native = None
    py = hdf5lite._blosc_decompress(chunk)
    del os.environ["HDF5LITE_PURE_PYTHON"]
    hdf5lite._native = None
    assert hdf5lite._native_lib() is not None, "libhbio.so not built (run __graft_entry__.build())"
    c = hdf5lite._blosc_decompress(chunk)
    return py, c


def test_blosclz_literals_matches_and_overlapping_runs():
    # literal 'abcd' | match len 4 dist 4 -> 'abcd' | match len 4 dist 1 -> 'dddd' | literal 'xy'
    stream = bytes([0x03]) + b"abcd" + bytes([0x40, 0x03]) + bytes([0x40, 0x00]) + bytes([0x01]) + b"xy"
    want = b"abcdabcdddddxy"
    py, c = _decode_both(_blosc_chunk(stream, len(want)))
    assert py == want and c == want
    # long match: length field 7 -> extension byte; 3 + 6 + 10 = 19 bytes from distance 2
    stream = bytes([0x01]) + b"pq" + bytes([0xE0, 10, 0x01])
    want = b"pq" + (b"pq" * 10)[:19]
    py, c = _decode_both(_blosc_chunk(stream, len(want)))
    assert py == want and c == want


def test_blosc_shuffle_and_stored_blocks():
    vals = np.arange(40, dtype="<i4")
    shuffled = vals.view(np.uint8).reshape(-1, 4).T.tobytes()           # byte planes
    # a stream whose compressed size equals the block size is stored verbatim
    chunk = struct.pack("<BBBBIII", 2, 1, 0x1, 4, 160, 160, 16 + 4 + 4 + 160) + struct.pack("<i", 20) + \
        struct.pack("<i", 160) + shuffled
    py, c = _decode_both(chunk)
    assert py == vals.tobytes() and c == vals.tobytes()
    memcpyed = struct.pack("<BBBBIII", 2, 1, 0x2, 4, 160, 160, 176) + vals.tobytes()
    py, c = _decode_both(memcpyed)
    assert py == vals.tobytes() and c == vals.tobytes()


def test_malformed_chunk_is_rejected():
    stream = bytes([0x40, 0x09])              # back-reference before the start of the output
    with pytest.raises(ValueError):
        hdf5lite._blosc_decompress(_blosc_chunk(stream, 8))


@pytest.mark.skipif(not os.path.isdir(TD), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", ["small_test_matrix.h5", "small_test_matrix_50kb_res.h5",
                                  "hicCorrectMatrix/small_test_matrix_ICEcorrected_chrUextra_chr3LHet.h5"])
def test_native_and_python_decoders_agree_on_reference_fixtures(name):
    def read_all():
        f = hdf5lite.File(TD + name)
        return {k: f["matrix/" + k].read() for k in ("data", "indices", "indptr")}
    os.environ["HDF5LITE_PURE_PYTHON"] = "1"
    hdf5lite._native = None
    py = read_all()
    del os.environ["HDF5LITE_PURE_PYTHON"]
    hdf5lite._native = None
    c = read_all()
    for k in py:
        np.testing.assert_array_equal(py[k], c[k])


def test_overwriting_the_file_that_is_still_mapped(tmp_path, monkeypatch):
    """big datasets of an uncompressed file are read as views of the mapped file; saving to the same path must not
    pull the pages away"""
    monkeypatch.setenv("HDF5LITE_COMPRESS", "0")
    from scipy import sparse
    from hicexplorer_b200.hicmatrix_lite import hiCMatrix
    ma = hiCMatrix()
    up = sparse.random(300, 300, 0.1, random_state=np.random.default_rng(0), format="csr")
    up = sparse.triu(sparse.csr_matrix(np.round(up * 100)), 0, format="csr")
    ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    ma.cut_intervals = [("chr1", i * 10, (i + 1) * 10, 1.0) for i in range(300)]
    ma._index_intervals()
    path = str(tmp_path / "m.h5")
    ma.save(path)
    again = hiCMatrix(path)                       # arrays are views of the mapping
    assert again.upper_for_device() is not None and not again.upper_for_device().data.flags.writeable
    again.save(path)                              # same path
    third = hiCMatrix(path)
    assert (third.matrix != ma.matrix).nnz == 0 and (again.matrix != ma.matrix).nnz == 0


@pytest.mark.parametrize("ext", [".h5", ".cool"])
def test_compressed_files_mirror_the_reference_layout_and_round_trip(tmp_path, ext):
    """the writer produces what the reference's writers produce (docs/content/file-formats.rst:9-31, 64-76): chunked
    datasets with PyTables' Blosc pipeline (.h5) / cooler's shuffle + deflate (.cool), a version-1 chunk B-tree closed
    like the golden files'; both decoders (C and pure Python) read them back bit for bit"""
    from scipy import sparse
    from hicexplorer_b200.hicmatrix_lite import hiCMatrix
    rng = np.random.default_rng(1)
    n = 3000
    up = sparse.random(n, n, 0.06, random_state=rng, format="csr")     # ~270 k entries: several chunks, 2 B-tree levels in pixels
    up = sparse.triu(sparse.csr_matrix(np.round(up * 50)), 0, format="csr")
    up.eliminate_zeros()
    ma = hiCMatrix()
    ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    ma.cut_intervals = [("chr%d" % (1 + i // 1000), (i % 1000) * 10, (i % 1000 + 1) * 10, 1.0) for i in range(n)]
    ma._index_intervals()
    ma.correction_factors = rng.random(n)
    path = str(tmp_path / ("m" + ext))
    ma.save(path)
    f = hdf5lite.File(path)
    big = f["matrix/data"] if ext == ".h5" else f["pixels/count"]
    if ext == ".h5":
        assert big._filters == [(32001, (2, 2, 8, 65536, 5, 1))]              # as in the golden files
        assert f["matrix/indices"]._filters == [(32001, (2, 2, 4, 65536, 5, 1))]
        assert f["matrix/data"].attrs["CLASS"] == "CARRAY" and f.attrs["TITLE"] == "HiCExplorer matrix"
    else:
        assert big._filters == [(2, (8,)), (1, (hdf5lite._GZIP_LEVEL,))]
        assert f["bins/chrom"].enum is not None
    assert os.path.getsize(path) < 0.6 * (12 * up.nnz)                        # it does compress
    # the chunk index: fixed-size nodes for K = 32, closing key = (end of the last chunk, element size)
    lay = big._layout
    import struct as st
    (root,) = st.unpack_from("<Q", lay, 3)
    buf = f._buf
    assert buf[root:root + 4] == b"TREE" and buf[root + 4] == 1
    level, nent = buf[root + 5], st.unpack_from("<H", buf, root + 6)[0]
    nchunks = -(-big.shape[0] // 8192)
    assert (level == 0 and nent == nchunks) if nchunks <= 64 else level >= 1
    for mode in ("native", "python"):
        if mode == "python":
            os.environ["HDF5LITE_PURE_PYTHON"] = "1"
        hdf5lite._native = None
        try:
            back = hiCMatrix(path)
        finally:
            os.environ.pop("HDF5LITE_PURE_PYTHON", None)
            hdf5lite._native = None
        assert (back.matrix != ma.matrix).nnz == 0
        assert [iv[:3] for iv in back.cut_intervals] == [iv[:3] for iv in ma.cut_intervals]
        np.testing.assert_array_equal(np.asarray(back.correction_factors).ravel(), ma.correction_factors)


def test_chunk_btree_with_several_levels(tmp_path):
    """more than 64 chunks -> internal nodes; more than 4096 -> three levels"""
    for n in (8192 * 70 + 5, 8192 * 4100):
        a = (np.arange(n, dtype=np.int64) * 7919) % 1000
        w = hdf5lite.Writer()
        w.dataset("/x", a, filters="blosc", chunk_len=8192)
        path = str(tmp_path / "t.h5")
        w.save(path)
        np.testing.assert_array_equal(hdf5lite.File(path)["x"].read(), a)


def test_threaded_canonical_check_agrees_with_scipy():
    """hbio_csr_is_canonical (all host cores) replaces scipy's one-thread has_canonical_format walk at load time"""
    from scipy import sparse
    from hicexplorer_b200 import hdf5lite as h5
    if h5._native_lib() is None:
        pytest.skip("libhbio.so not built")
    rng = np.random.default_rng(0)
    m = sparse.random(3000, 3000, 0.02, random_state=rng, format="csr")
    m.sort_indices()
    assert h5.csr_is_canonical(m) is True
    swapped = m.copy()
    k = int(swapped.indptr[7])
    assert swapped.indptr[8] - k >= 2
    swapped.indices[k], swapped.indices[k + 1] = swapped.indices[k + 1], swapped.indices[k]
    dup = m.copy()
    dup.indices[k + 1] = dup.indices[k]
    beyond = m.copy()
    beyond.indices[-1] = 3000
    negative = m.copy()
    negative.indices[k] = -1
    for bad in (swapped, dup, beyond, negative):
        assert h5.csr_is_canonical(bad) is False
        fresh = sparse.csr_matrix((bad.data, bad.indices, bad.indptr), shape=bad.shape)
        assert fresh.has_canonical_format is False or bad is beyond or bad is negative
    assert h5.csr_is_canonical(sparse.csr_matrix((5, 5))) is True
    wide = sparse.csr_matrix((m.data, m.indices.astype(np.int64), m.indptr.astype(np.int64)), shape=m.shape)
    assert h5.csr_is_canonical(wide) is True
    broken_ptr = sparse.csr_matrix((m.data, m.indices, m.indptr), shape=m.shape)
    broken_ptr.indptr = m.indptr.copy()
    broken_ptr.indptr[10] = broken_ptr.indptr[11] + 5
    assert h5.csr_is_canonical(broken_ptr) is False
