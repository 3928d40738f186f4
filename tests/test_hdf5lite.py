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


def test_overwriting_the_file_that_is_still_mapped(tmp_path):
    """big datasets are read as views of the mapped file; saving to the same path must not pull the pages away"""
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
