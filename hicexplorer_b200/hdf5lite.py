"""Minimal pure-Python HDF5 reader/writer for the two on-disk layouts of the
matrix-correction path: HiCExplorer ``.h5`` (PyTables CArrays, Blosc/blosclz)
and ``.cool`` (h5py, deflate+shuffle).

Why this exists: the reference reads and writes its matrices through the
third-party ``hicmatrix`` package (PyTables / cooler / h5py; call sites
``hicexplorer/hicCorrectMatrix.py:603-609, 779``), none of which is available
on the build or GPU boxes.  The layout contract is
``docs/content/file-formats.rst:9-147`` of the reference.  Only the subset of
HDF5 those files use is implemented:

* superblock v0, object header v1 (+continuation blocks), old-style groups
  (symbol-table B-tree v1 + local heap),
* dataspace v1/v2, datatypes fixed-point / IEEE float / fixed string / enum,
* layout v3 contiguous / chunked (B-tree v1) / compact,
* filters: deflate (1), shuffle (2), Blosc (32001, blosclz/zlib codecs,
  byte-shuffle), fletcher32 (3, checksum ignored),
* attributes v1/v2/v3 with fixed-size types (variable-length strings are
  returned as ``None``).

The writer emits uncompressed contiguous datasets with PyTables-compatible
cosmetic attributes so that generic HDF5 libraries (h5py, PyTables, cooler)
can open the result.

This is host-side I/O, not part of the device hot path.
"""
import mmap
import os
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------
# Blosc-1 container / blosclz codec (decode only)
# --------------------------------------------------------------------------
def _blosclz_decompress(src, maxout):
    """FastLZ-style decoder used by Blosc's ``blosclz`` codec."""
    out = bytearray(maxout)
    ip = 0
    n = len(src)
    op = 0
    ctrl = src[ip] & 31
    ip += 1
    while True:
        if ctrl >= 32:
            length = (ctrl >> 5) - 1
            ofs = (ctrl & 31) << 8
            if length == 6:
                while True:
                    code = src[ip]
                    ip += 1
                    length += code
                    if code != 255:
                        break
            code = src[ip]
            ip += 1
            ref = op - ofs - code
            if code == 255 and ofs == (31 << 8):
                ofs = (src[ip] << 8) + src[ip + 1]
                ip += 2
                ref = op - ofs - 8191
            more = ip < n
            if more:
                ctrl = src[ip]
                ip += 1
            length += 3
            ref -= 1
            if ref < 0:
                raise ValueError("blosclz: bad back-reference")
            dist = op - ref
            if dist >= length:
                out[op:op + length] = out[ref:ref + length]
            else:
                # overlapping copy == periodic repetition of the last `dist` bytes
                pat = bytes(out[ref:op])
                reps = (length + dist - 1) // dist
                out[op:op + length] = (pat * reps)[:length]
            op += length
            if not more:
                break
        else:
            ctrl += 1
            out[op:op + ctrl] = src[ip:ip + ctrl]
            op += ctrl
            ip += ctrl
            if ip >= n:
                break
            ctrl = src[ip]
            ip += 1
    return bytes(out[:op])


def _unshuffle(buf, typesize):
    n = len(buf)
    if typesize <= 1 or n < typesize:
        return bytes(buf)
    nelem = n // typesize
    body = np.frombuffer(buf, dtype=np.uint8, count=nelem * typesize)
    body = body.reshape(typesize, nelem).T.tobytes()
    return body + bytes(buf[nelem * typesize:])


_native = None


def _native_lib():
    """libhbio.so (csrc/hb_io.c): the same decoders in C; None if it was not built."""
    global _native
    if _native is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhbio.so")
        if not os.path.exists(path) or os.environ.get("HDF5LITE_PURE_PYTHON") == "1":
            _native = False
        else:
            lib = ctypes.CDLL(path)
            lib.hbio_blosc_nbytes.restype = ctypes.c_int64
            lib.hbio_blosc_nbytes.argtypes = [ctypes.c_char_p]
            lib.hbio_blosc_blocksize.restype = ctypes.c_int64
            lib.hbio_blosc_blocksize.argtypes = [ctypes.c_char_p]
            lib.hbio_blosc_decode.restype = ctypes.c_int64
            lib.hbio_blosc_decode.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
            if hasattr(lib, "hbio_decode_chunks"):
                lib.hbio_decode_chunks.restype = ctypes.c_int64
                lib.hbio_decode_chunks.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 5 + [
                    ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
            if hasattr(lib, "hbio_encode_chunks"):
                lib.hbio_encode_bound.restype = ctypes.c_int64
                lib.hbio_encode_bound.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
                lib.hbio_encode_chunks.restype = ctypes.c_int64
                lib.hbio_encode_chunks.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p]
            if hasattr(lib, "hbio_csr_is_canonical"):
                lib.hbio_csr_is_canonical.restype = ctypes.c_int
                lib.hbio_csr_is_canonical.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
            _native = lib
    return _native or None


def csr_is_canonical(m):
    """scipy's ``has_canonical_format`` test (sorted rows, no duplicates) on all host cores; None when libhbio.so is not
    there or the arrays are of a kind it does not take (the caller then asks scipy)"""
    import os
    lib = _native_lib()
    if lib is None or not hasattr(lib, "hbio_csr_is_canonical"):
        return None
    ip, ix = m.indptr, m.indices
    if ip.dtype.itemsize not in (4, 8) or ix.dtype.itemsize not in (4, 8) or ip.dtype.kind != "i" or ix.dtype.kind != "i":
        return None
    if not (ip.flags.c_contiguous and ix.flags.c_contiguous):
        return None
    rc = lib.hbio_csr_is_canonical(ip.ctypes.data, ip.dtype.itemsize, ix.ctypes.data, ix.dtype.itemsize, m.shape[0], m.shape[1],
                                   len(ix), min(os.cpu_count() or 1, 16))
    return None if rc < 0 else bool(rc)


def _blosc_decompress(chunk):
    lib = _native_lib()
    if lib is not None:
        chunk = bytes(chunk)
        nbytes = lib.hbio_blosc_nbytes(chunk)
        out = np.empty(max(int(nbytes), 1), dtype=np.uint8)
        scratch = np.empty(max(int(lib.hbio_blosc_blocksize(chunk)), 1), dtype=np.uint8)
        got = lib.hbio_blosc_decode(chunk, len(chunk), out.ctypes.data, scratch.ctypes.data)
        if got == nbytes:
            return out[:nbytes].tobytes()
        if got == -1:
            raise ValueError("malformed Blosc chunk")
        # -2: codec the C decoder does not handle -> pure-Python path below
    version, versionlz, flags, typesize = struct.unpack_from("<BBBB", chunk, 0)
    nbytes, blocksize, cbytes = struct.unpack_from("<III", chunk, 4)
    if flags & 0x2:  # memcpyed
        return bytes(chunk[16:16 + nbytes])
    codec = flags >> 5
    doshuffle = bool(flags & 0x1)
    if flags & 0x4:
        raise NotImplementedError("blosc bit-shuffle")
    dont_split = bool(flags & 0x10)
    nblocks = (nbytes + blocksize - 1) // blocksize if blocksize else 0
    bstarts = struct.unpack_from("<%di" % nblocks, chunk, 16)
    out = bytearray()
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split and not leftover and typesize <= 16
                 and blocksize // typesize >= 128)
        nsplits = typesize if split else 1
        neblock = bsize // nsplits
        pos = bstarts[b]
        blk = bytearray()
        for _ in range(nsplits):
            (csize,) = struct.unpack_from("<i", chunk, pos)
            pos += 4
            raw = chunk[pos:pos + csize]
            pos += csize
            if csize == neblock:
                blk += raw
            elif codec == 0:
                blk += _blosclz_decompress(raw, neblock)
            elif codec == 3:
                blk += zlib.decompress(raw)
            else:
                raise NotImplementedError("blosc codec %d" % codec)
        if doshuffle:
            blk = _unshuffle(blk, typesize)
        out += blk
    return bytes(out[:nbytes])


# filter pipelines of the reference's writers: PyTables CArrays (Blosc / blosclz, level 5, byte shuffle inside Blosc) and
# cooler datasets (HDF5 shuffle + deflate); name -> (mode of hbio_encode_chunks, [(filter id, name, client data)])
def _pipeline(kind, itemsize, chunk_bytes):
    if kind == "blosc":
        return 1, [(32001, b"blosc", (2, 2, itemsize, chunk_bytes, 5, 1))]
    if kind == "gzip-shuffle":
        if itemsize > 1:
            return 2, [(2, b"shuffle", (itemsize,)), (1, b"deflate", (_GZIP_LEVEL,))]
        return 3, [(1, b"deflate", (_GZIP_LEVEL,))]
    if kind == "gzip":
        return 3, [(1, b"deflate", (_GZIP_LEVEL,))]
    raise ValueError("unknown filter pipeline %r" % (kind,))


_GZIP_LEVEL = 4   # what this writer asks zlib for (cooler's default is 6; readers do not care)


def _encode_chunks(raw, chunk_bytes, itemsize, mode):
    """Encode consecutive chunks of ``raw`` (uint8 view of the dataset; the last chunk is zero-padded to full size).
    Returns (buffer with the encoded chunks back to back, offsets, sizes).  libhbio.so does it on all host cores;
    without it the chunks are written in a form every reader accepts but that does not compress (Blosc "memcpyed"
    chunks) or through Python's zlib."""
    import os
    nbytes = int(raw.nbytes)
    n = (nbytes + chunk_bytes - 1) // chunk_bytes
    lib = _native_lib()
    if lib is not None and hasattr(lib, "hbio_encode_chunks"):
        slot = int(lib.hbio_encode_bound(chunk_bytes, itemsize, mode))
        slots = np.empty(n * slot, dtype=np.uint8)
        dst = np.empty(n * slot, dtype=np.uint8)
        offsets = np.zeros(n, dtype=np.int64)
        sizes = np.zeros(n, dtype=np.int64)
        threads = int(os.environ.get("HDF5LITE_THREADS", "0")) or min(32, os.cpu_count() or 1)
        total = lib.hbio_encode_chunks(raw.ctypes.data, nbytes, chunk_bytes, itemsize, mode, _GZIP_LEVEL, threads,
                                       slots.ctypes.data, slot, dst.ctypes.data, offsets.ctypes.data, sizes.ctypes.data)
        if total < 0:
            raise RuntimeError("chunk encoder failed")
        return dst[:total], offsets, sizes
    parts, offsets, sizes, pos = [], [], [], 0
    for c in range(n):
        blk = bytes(raw[c * chunk_bytes:(c + 1) * chunk_bytes])
        blk += b"\x00" * (chunk_bytes - len(blk))
        if mode == 1:      # Blosc container, "memcpyed" flag: header + the bytes as they are
            enc = struct.pack("<BBBBIII", 2, 1, 0x2 | (0x1 if itemsize > 1 else 0), itemsize, chunk_bytes, chunk_bytes,
                              chunk_bytes + 16) + blk
        else:
            if mode == 2 and itemsize > 1:
                blk = np.frombuffer(blk, dtype=np.uint8).reshape(-1, itemsize).T.tobytes()
            enc = zlib.compress(blk, _GZIP_LEVEL)
        parts.append(enc)
        offsets.append(pos)
        sizes.append(len(enc))
        pos += len(enc)
    return np.frombuffer(b"".join(parts), dtype=np.uint8), np.asarray(offsets, dtype=np.int64), np.asarray(sizes, dtype=np.int64)


def _filter_message(filters):
    """filter pipeline message, version 1 (what libhdf5 1.8 / PyTables / h5py write)"""
    body = struct.pack("<BB6x", 1, len(filters))
    for fid, name, cd in filters:
        nm = name + b"\x00"
        nm += b"\x00" * (_pad8(len(nm)) - len(nm))
        body += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm + struct.pack("<%dI" % len(cd), *cd)
        if len(cd) % 2:
            body += b"\x00" * 4
    return body


_CHUNK_BTREE_K = 32   # "indexed storage internal node K": superblock v0 does not store it, the library default applies


def _chunk_btree(alloc_at, first_addr, chunk_addrs, chunk_sizes, chunk_len, itemsize, n_chunks, rank=1):
    """version-1 B-tree of type 1 over the chunks of a 1-D dataset, as node blobs placed from ``first_addr`` on.
    Every node has the library's fixed size for K = 32; leaves hold up to 64 chunks; siblings are linked.
    Returns (root address, list of (address, bytes))."""
    K = _CHUNK_BTREE_K
    key_size = 8 + 8 * (rank + 1)
    node_size = 24 + (2 * K + 1) * key_size + 2 * K * 8

    def key(nbytes, offset, last_dim):
        # offsets of the chunk in elements per dimension (trailing dimensions of length 1 stay 0) + the element-size slot
        return struct.pack("<IIQ", nbytes, 0, offset) + b"\x00" * (8 * (rank - 1)) + struct.pack("<Q", last_dim)

    # level 0: (first key fields, child address) per chunk
    entries = [((int(chunk_sizes[c]), c * chunk_len), int(chunk_addrs[c])) for c in range(n_chunks)]
    end_key = key(0, n_chunks * chunk_len, itemsize)     # PyTables' files close the tree like this
    blobs = []
    addr = first_addr
    level = 0
    while True:
        groups = [entries[i:i + 2 * K] for i in range(0, len(entries), 2 * K)]
        addrs = [addr + i * node_size for i in range(len(groups))]
        nxt = []
        for gi, grp in enumerate(groups):
            left = addrs[gi - 1] if gi > 0 else _UNDEF
            right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
            blob = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right))
            for (nbytes, off), child in grp:
                blob += key(nbytes, off, 0) + struct.pack("<Q", child)
            # closing key of the node: the first key of the next node, or the end of the dataset
            if gi + 1 < len(groups):
                (nb, off), _child = groups[gi + 1][0]
                blob += key(nb, off, 0)
            else:
                blob += end_key
            blob += b"\x00" * (node_size - len(blob))
            blobs.append((addrs[gi], bytes(blob)))
            nxt.append((grp[0][0], addrs[gi]))
        addr += len(groups) * node_size
        if len(groups) == 1:
            return addrs[0], blobs, addr
        entries = nxt
        level += 1


# --------------------------------------------------------------------------
# Reader
# --------------------------------------------------------------------------
class _Dataset(object):
    def __init__(self, f, msgs):
        self._f = f
        self.attrs = {}
        self.shape = ()
        self.dtype = None
        self._layout = None
        self._filters = []
        self._enum = None
        for mtype, body in msgs:
            if mtype == 0x01:
                self.shape = _parse_dataspace(body)
            elif mtype == 0x03:
                self.dtype, self._enum, _ = _parse_datatype(body, 0)
            elif mtype == 0x08:
                self._layout = body
            elif mtype == 0x0B:
                self._filters = _parse_filters(body)
            elif mtype == 0x0C:
                k, v = _parse_attribute(body)
                self.attrs[k] = v

    @property
    def enum(self):
        """dict name -> value for enum datasets (cooler ``bins/chrom``)."""
        return self._enum

    def read(self, copy=True):
        """``copy=False``: a contiguous, unfiltered dataset comes back as a read-only view of the mapped file (no copy of
        a multi-gigabyte payload); anything else is decoded into a fresh array as usual."""
        buf = self._f._buf
        lay = self._layout
        ver = lay[0]
        if ver != 3:
            raise NotImplementedError("layout version %d" % ver)
        cls = lay[1]
        count = int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1
        itemsize = self.dtype.itemsize
        if cls == 0:  # compact
            (size,) = struct.unpack_from("<H", lay, 2)
            raw = lay[4:4 + size]
            return np.frombuffer(raw, dtype=self.dtype, count=count).reshape(self.shape).copy()
        if cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == _UNDEF:
                return np.zeros(self.shape, dtype=self.dtype)
            view = np.frombuffer(buf, dtype=self.dtype, count=count, offset=addr).reshape(self.shape)
            return view.copy() if copy else view
        if cls != 2:
            raise NotImplementedError("layout class %d" % cls)
        ndims = lay[2]
        (btree,) = struct.unpack_from("<Q", lay, 3)
        cdims = struct.unpack_from("<%dI" % ndims, lay, 11)
        chunk_shape = cdims[:-1]
        out = np.zeros(self.shape, dtype=self.dtype)
        if btree == _UNDEF or count == 0:
            return out
        chunk_bytes = int(np.prod(chunk_shape, dtype=np.int64)) * itemsize
        if len(self.shape) == 1 and self._read_chunks_native(out, btree, ndims, chunk_shape[0], chunk_bytes, itemsize):
            return out
        for size, fmask, offs, addr in self._f._iter_chunks(btree, ndims):
            raw = buf[addr:addr + size]
            for i in range(len(self._filters) - 1, -1, -1):
                if fmask & (1 << i):
                    continue
                fid, cd = self._filters[i]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    raw = _unshuffle(raw, cd[0] if cd else itemsize)
                elif fid == 3:
                    raw = raw[:-4]
                elif fid == 32001:
                    raw = _blosc_decompress(raw)
                else:
                    raise NotImplementedError("HDF5 filter %d" % fid)
            if len(raw) < chunk_bytes:
                raise ValueError("short chunk")
            arr = np.frombuffer(raw, dtype=self.dtype, count=chunk_bytes // itemsize).reshape(chunk_shape)
            sl_out = []
            sl_in = []
            for d in range(len(chunk_shape)):
                lo = offs[d]
                hi = min(lo + chunk_shape[d], self.shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = arr[tuple(sl_in)]
        return out

    def _read_chunks_native(self, out, btree, ndims, chunk_len, chunk_bytes, itemsize):
        """1-D chunked dataset: all chunks decoded by libhbio's thread pool in one call.  False = not handled here."""
        lib = _native_lib()
        buf = self._f._buf
        if lib is None or not hasattr(lib, "hbio_decode_chunks") or not isinstance(buf, mmap.mmap):
            return False
        if any(fid not in (1, 2, 3, 32001) for fid, _cd in self._filters):
            return False
        chunks = list(self._f._iter_chunks(btree, ndims))
        if not chunks:
            return True
        n = len(chunks)
        addr = np.fromiter((c[3] for c in chunks), dtype=np.int64, count=n)
        csize = np.fromiter((c[0] for c in chunks), dtype=np.int64, count=n)
        fmask = np.fromiter((c[1] for c in chunks), dtype=np.uint32, count=n)
        start = np.fromiter((c[2][0] for c in chunks), dtype=np.int64, count=n)
        stop = np.minimum(start + chunk_len, self.shape[0])
        if np.any(start >= self.shape[0]) or np.any(addr + csize > len(buf)):
            return False
        out_off = start * itemsize
        out_len = (stop - start) * itemsize
        filters = np.asarray([fid for fid, _cd in self._filters], dtype=np.int32)
        tsizes = np.asarray([(cd[0] if (fid == 2 and cd) else itemsize) for fid, cd in self._filters], dtype=np.int32)
        base = np.frombuffer(buf, dtype=np.uint8)
        import os
        threads = int(os.environ.get("HDF5LITE_THREADS", "0")) or min(16, os.cpu_count() or 1)
        view = out.view(np.uint8).reshape(-1)
        rc = lib.hbio_decode_chunks(base.ctypes.data, n, addr.ctypes.data, csize.ctypes.data, fmask.ctypes.data,
                                    out_off.ctypes.data, out_len.ctypes.data, chunk_bytes, filters.ctypes.data,
                                    tsizes.ctypes.data, len(filters), view.ctypes.data, threads)
        if rc == -1:
            raise ValueError("malformed chunk data")
        return rc == 0

    def __getitem__(self, key):
        return self.read()[key]


class _Group(object):
    def __init__(self, f, msgs):
        self._f = f
        self.attrs = {}
        self._links = {}
        for mtype, body in msgs:
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                self._links = f._read_symbol_table(btree, heap)
            elif mtype == 0x0C:
                k, v = _parse_attribute(body)
                self.attrs[k] = v

    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, _Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._open(node._links[part])
        return node


class File(_Group):
    """Read-only view of an HDF5 file: ``File(path)['matrix/data'].read()``."""

    def __init__(self, path):
        # map the file instead of reading it: only the pages that are parsed or copied out are touched (a genome-wide
        # matrix file is gigabytes; reading it whole doubled the peak memory and the load time)
        with open(path, "rb") as fh:
            try:
                self._buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except (ValueError, OSError):
                self._buf = fh.read()
        buf = self._buf
        if buf[:8] != _SIG:
            raise ValueError("%s is not an HDF5 file" % path)
        if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
            raise NotImplementedError("only HDF5 superblock v0 with 8-byte offsets is supported")
        (root_hdr,) = struct.unpack_from("<Q", buf, 56 + 8)
        self._cache = {}
        _Group.__init__(self, self, self._read_object_header(root_hdr))

    # -- low level ----------------------------------------------------------
    def _read_object_header(self, addr):
        buf = self._buf
        version = buf[addr]
        if version != 1:
            raise NotImplementedError("object header version %d" % version)
        (nmsgs,) = struct.unpack_from("<H", buf, addr + 2)
        (hsize,) = struct.unpack_from("<I", buf, addr + 8)
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, pos)
                body = buf[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    off, length = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, length))
                msgs.append((mtype, body))
        return msgs

    def _open(self, addr):
        if addr not in self._cache:
            msgs = self._read_object_header(addr)
            types = set(m[0] for m in msgs)
            if 0x11 in types:
                self._cache[addr] = _Group(self, msgs)
            elif 0x08 in types:
                self._cache[addr] = _Dataset(self, msgs)
            else:
                raise NotImplementedError("unsupported HDF5 object (new-style group?)")
        return self._cache[addr]

    def _read_symbol_table(self, btree, heap):
        buf = self._buf
        if buf[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        (data_seg,) = struct.unpack_from("<Q", buf, heap + 24)
        links = {}

        def walk(node):
            if buf[node:node + 4] == b"TREE":
                level = buf[node + 5]
                (nent,) = struct.unpack_from("<H", buf, node + 6)
                pos = node + 24
                for _ in range(nent):
                    (child,) = struct.unpack_from("<Q", buf, pos + 8)
                    pos += 16
                    walk(child)
                del level
            elif buf[node:node + 4] == b"SNOD":
                (nsym,) = struct.unpack_from("<H", buf, node + 6)
                pos = node + 8
                for _ in range(nsym):
                    name_off, hdr = struct.unpack_from("<QQ", buf, pos)
                    pos += 40
                    s = data_seg + name_off
                    e = buf.find(b"\x00", s)
                    links[buf[s:e].decode("utf-8")] = hdr
            else:
                raise ValueError("bad group node")

        walk(btree)
        return links

    def _iter_chunks(self, node, ndims):
        buf = self._buf
        if buf[node:node + 4] != b"TREE":
            raise ValueError("bad chunk B-tree node")
        level = buf[node + 5]
        (nent,) = struct.unpack_from("<H", buf, node + 6)
        pos = node + 24
        keysize = 8 + 8 * ndims
        for _ in range(nent):
            size, fmask = struct.unpack_from("<II", buf, pos)
            offs = struct.unpack_from("<%dQ" % ndims, buf, pos + 8)
            (child,) = struct.unpack_from("<Q", buf, pos + keysize)
            pos += keysize + 8
            if level == 0:
                yield size, fmask, offs, child
            else:
                for item in self._iter_chunks(child, ndims):
                    yield item


def _parse_dataspace(body):
    version, rank, flags = body[0], body[1], body[2]
    if version == 1:
        pos = 8
    elif version == 2:
        pos = 4
        if body[3] == 2:  # null dataspace
            return (0,)
    else:
        raise NotImplementedError("dataspace version %d" % version)
    return tuple(struct.unpack_from("<%dQ" % rank, body, pos))


def _parse_datatype(body, pos):
    """Returns (numpy dtype, enum-dict-or-None, bytes consumed)."""
    cv = body[pos]
    cls, version = cv & 0x0F, cv >> 4
    bits0 = body[pos + 1]
    bits = bits0 | (body[pos + 2] << 8) | (body[pos + 3] << 16)
    (size,) = struct.unpack_from("<I", body, pos + 4)
    p = pos + 8
    if cls == 0:
        order = ">" if bits0 & 1 else "<"
        kind = "i" if bits0 & 8 else "u"
        return np.dtype("%s%s%d" % (order, kind, size)), None, p + 4 - pos
    if cls == 1:
        order = ">" if bits0 & 1 else "<"
        return np.dtype("%sf%d" % (order, size)), None, p + 12 - pos
    if cls == 3:
        return np.dtype("S%d" % size), None, p - pos
    if cls == 8:
        nmemb = bits & 0xFFFF
        base, _, used = _parse_datatype(body, p)
        p += used
        names = []
        for _ in range(nmemb):
            e = body.index(b"\x00", p)
            names.append(body[p:e].decode("utf-8"))
            ln = e - p + 1
            if version < 3:
                ln = (ln + 7) // 8 * 8
            p += ln
        vals = np.frombuffer(body, dtype=base, count=nmemb, offset=p)
        p += nmemb * base.itemsize
        return base, dict(zip(names, [int(v) for v in vals])), p - pos
    if cls == 9:
        return np.dtype("V%d" % size), None, None
    if cls == 6:
        return np.dtype("V%d" % size), None, None
    raise NotImplementedError("datatype class %d" % cls)


def _parse_filters(body):
    version, nfilt = body[0], body[1]
    pos = 8 if version == 1 else 2
    out = []
    for _ in range(nfilt):
        (fid,) = struct.unpack_from("<H", body, pos)
        pos += 2
        if version == 1 or fid >= 256:
            (nlen,) = struct.unpack_from("<H", body, pos)
            pos += 2
        else:
            nlen = 0
        _flags, ncd = struct.unpack_from("<HH", body, pos)
        pos += 4
        if version == 1:
            nlen = (nlen + 7) // 8 * 8
        pos += nlen
        cd = struct.unpack_from("<%dI" % ncd, body, pos)
        pos += 4 * ncd
        if version == 1 and ncd % 2:
            pos += 4
        out.append((fid, cd))
    return out


def _pad8(n):
    return (n + 7) // 8 * 8


def _parse_attribute(body):
    version = body[0]
    nlen, tlen, slen = struct.unpack_from("<HHH", body, 2)
    pos = 8
    if version == 3:
        pos += 1
    name = body[pos:pos + nlen].split(b"\x00")[0].decode("utf-8", "replace")
    if version == 1:
        pos += _pad8(nlen)
        tbody = body[pos:pos + tlen]
        pos += _pad8(tlen)
        sbody = body[pos:pos + slen]
        pos += _pad8(slen)
    else:
        pos += nlen
        tbody = body[pos:pos + tlen]
        pos += tlen
        sbody = body[pos:pos + slen]
        pos += slen
    try:
        dt, _enum, used = _parse_datatype(tbody, 0)
        if used is None:
            return name, None
        shape = _parse_dataspace(sbody)
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        arr = np.frombuffer(body, dtype=dt, count=count, offset=pos)
        if dt.kind == "S":
            vals = [v.split(b"\x00")[0].decode("utf-8", "replace") for v in arr.tolist()]
            return name, (vals[0] if len(shape) == 0 else vals)
        if len(shape) == 0:
            return name, arr[0].item()
        return name, arr.reshape(shape).copy()
    except (NotImplementedError, ValueError, struct.error):
        return name, None


# --------------------------------------------------------------------------
# Writer (superblock v0, old-style groups, contiguous datasets)
# --------------------------------------------------------------------------
def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 8:
            return (struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8)
                    + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))
        if dt.itemsize == 4:
            return (struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4)
                    + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127))
    if dt.kind == "S":
        # null-padded ASCII fixed string
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    if dt.kind == "b":
        return _dtype_message(np.uint8)
    raise NotImplementedError("cannot write dtype %r" % dt)


def _enum_message(dt, mapping):
    dt = np.dtype(dt)
    names = list(mapping)
    body = struct.pack("<BBBBI", 0x18, len(names) & 0xFF, (len(names) >> 8) & 0xFF, 0, dt.itemsize)
    body += _dtype_message(dt)
    for nm in names:
        enc = nm.encode("utf-8") + b"\x00"
        body += enc + b"\x00" * (_pad8(len(enc)) - len(enc))
    body += np.asarray([mapping[nm] for nm in names], dtype=dt).tobytes()
    return body


def _dataspace_message(shape):
    if shape is None:  # scalar
        return struct.pack("<BBBB4x", 1, 0, 0, 0)
    rank = len(shape)
    return struct.pack("<BBBB4x", 1, rank, 0, 0) + struct.pack("<%dQ" % rank, *shape)


def _message(mtype, body, flags=0):
    body = body + b"\x00" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_message(name, value):
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        arr = np.array(value + b"\x00", dtype="S%d" % (len(value) + 1))
        shape = None
    else:
        arr = np.asarray(value)
        shape = None if arr.ndim == 0 else arr.shape
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
    nm = name.encode("utf-8") + b"\x00"
    t = _dtype_message(arr.dtype)
    s = _dataspace_message(shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(t), len(s))
    body += nm + b"\x00" * (_pad8(len(nm)) - len(nm))
    body += t + b"\x00" * (_pad8(len(t)) - len(t))
    body += s + b"\x00" * (_pad8(len(s)) - len(s))
    body += arr.tobytes()
    return _message(0x0C, body)


def _default_chunk_len(n, itemsize):
    """PyTables' CArrays of the reference's files use 64 KB chunks (8192 x 8 B, 16384 x 4 B, 5461 x 12 B); for arrays of
    hundreds of megabytes the chunk grows to 1 MiB so that the chunk index stays small"""
    target = (1 << 20) if n * itemsize >= (128 << 20) else (1 << 16)
    return max(1, target // itemsize)


def _compression_enabled():
    import os
    return os.environ.get("HDF5LITE_COMPRESS", "1") != "0"


class Writer(object):
    """Build an HDF5 file in memory: ``w = Writer(); w.group('/matrix', attrs);
    w.dataset('/matrix/data', array, attrs); w.save(path)``."""

    def __init__(self, root_attrs=None):
        self._tree = {"attrs": dict(root_attrs or {}), "children": {}, "data": None}

    def _node(self, path, create=True):
        node = self._tree
        for part in [p for p in path.split("/") if p]:
            if part not in node["children"]:
                if not create:
                    raise KeyError(path)
                node["children"][part] = {"attrs": {}, "children": {}, "data": None}
            node = node["children"][part]
        return node

    def group(self, path, attrs=None):
        node = self._node(path)
        node["attrs"].update(attrs or {})

    def dataset(self, path, array, attrs=None, enum=None, filters=None, chunk_len=None):
        """``enum`` = ordered mapping name -> integer value: store as an HDF5 enum (cooler ``bins/chrom``).
        ``filters``: None = contiguous, uncompressed; "blosc" = chunked + Blosc (blosclz, shuffle) like PyTables CArrays;
        "gzip-shuffle" / "gzip" = chunked + (shuffle +) deflate like cooler / h5py.  1-D datasets only."""
        node = self._node(path)
        node["enum"] = enum
        node["filters"] = filters
        node["chunk_len"] = chunk_len
        arr = np.ascontiguousarray(array)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        node["data"] = arr
        node["attrs"].update(attrs or {})

    def save(self, path):
        """Stream the file out in one pass (every address is known once its block has been written; only the superblock
        at offset 0 is patched at the end).  Dataset payloads go from the numpy buffer straight to the file."""
        # written beside the target and renamed at the end: an input file that is still memory-mapped (File.read with
        # copy=False) may be the very path that is being overwritten
        tmp_path = "%s.tmp%d" % (path, os.getpid())
        fd = os.open(tmp_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        pos = [96]                       # the superblock goes to offset 0 at the end
        # multi-megabyte payloads go to a small pool of writer threads right away (page-cache copies of a single write()
        # run at a few GB/s; several threads on disjoint regions scale) and overlap the encoding of the next batch
        from concurrent.futures import ThreadPoolExecutor
        writers = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
        pending = []

        def put(off, view):
            while view.nbytes:
                done = os.pwrite(fd, view, off)
                off += done
                view = view[done:]

        def alloc(blob):
            pos[0] += (-pos[0]) % 8
            addr = pos[0]
            nbytes = blob.nbytes if isinstance(blob, memoryview) else len(blob)
            if nbytes >= (32 << 20):
                step = 64 << 20
                view = blob if isinstance(blob, memoryview) else memoryview(blob)
                for o in range(0, nbytes, step):
                    pending.append(writers.submit(put, addr + o, view[o:o + step]))
            else:
                os.pwrite(fd, blob, addr)
            pos[0] += nbytes
            return addr

        def write_object_header(msgs):
            body = b"".join(msgs)
            hdr = struct.pack("<BxHII4x", 1, len(msgs), 1, len(body))
            return alloc(hdr + body)

        def write_chunked(node):
            """1-D chunked dataset: encoded chunks, then the chunk B-tree, then the object header"""
            arr = node["data"]
            itemsize = arr.dtype.itemsize
            n = int(arr.shape[0])
            chunk_len = node.get("chunk_len") or _default_chunk_len(n, itemsize)
            chunk_bytes = chunk_len * itemsize
            mode, filters = _pipeline(node["filters"], itemsize, chunk_bytes)
            raw = arr.reshape(-1).view(np.uint8) if arr.dtype.kind != "S" else np.frombuffer(arr.tobytes(), dtype=np.uint8)
            n_chunks = (n + chunk_len - 1) // chunk_len
            addrs = np.zeros(n_chunks, dtype=np.int64)
            sizes = np.zeros(n_chunks, dtype=np.int64)
            # encode and write in batches of ~256 MB of input so that the scratch stays bounded
            per_batch = max(1, (256 << 20) // chunk_bytes)
            for c0 in range(0, n_chunks, per_batch):
                c1 = min(n_chunks, c0 + per_batch)
                part = raw[c0 * chunk_bytes:min(c1 * chunk_bytes, raw.nbytes)]
                # the zero padding of the final chunk is defined relative to its own start: hand the encoder whole chunks
                enc, offs, szs = _encode_chunks(part, chunk_bytes, itemsize, mode)
                base = alloc(memoryview(enc))
                addrs[c0:c1] = base + offs
                sizes[c0:c1] = szs
            pos[0] += (-pos[0]) % 8
            rank = arr.ndim
            root, blobs, end = _chunk_btree(None, pos[0], addrs, sizes, chunk_len, itemsize, n_chunks, rank)
            for a, blob in blobs:
                os.pwrite(fd, blob, a)
            pos[0] = end
            msgs = [
                _message(0x01, _dataspace_message(arr.shape)),
                _message(0x03, _enum_message(arr.dtype, node["enum"]) if node.get("enum") else
                         _dtype_message(arr.dtype), flags=1),
                _message(0x05, struct.pack("<BBBB4x", 2, 3, 0, 1)),       # fill value as PyTables writes it
                _message(0x0B, _filter_message(filters)),
                _message(0x08, struct.pack("<BBBQ", 3, 2, rank + 1, root) +
                         struct.pack("<%dI" % (rank + 1), *([chunk_len] + [1] * (rank - 1) + [itemsize]))),
            ]
            for k, v in node["attrs"].items():
                msgs.append(_attr_message(k, v))
            return write_object_header(msgs)

        def write_dataset(node):
            arr = node["data"]
            nbytes = int(arr.nbytes)
            if (node.get("filters") and arr.ndim >= 1 and all(d == 1 for d in arr.shape[1:]) and nbytes > 0
                    and _compression_enabled()):
                return write_chunked(node)
            # a byte view of the (contiguous) array: no intermediate copy of a multi-gigabyte payload
            raw = memoryview(arr.reshape(-1).view(np.uint8)) if nbytes and arr.dtype.kind != "S" else arr.tobytes()
            daddr = alloc(raw) if nbytes else _UNDEF
            shape = arr.shape if arr.ndim else None
            msgs = [
                _message(0x01, _dataspace_message(shape)),
                _message(0x03, _enum_message(arr.dtype, node["enum"]) if node.get("enum") else
                         _dtype_message(arr.dtype), flags=1),
                _message(0x05, struct.pack("<BBBB", 2, 2, 2, 0x02)),  # fill value: never written
                _message(0x08, struct.pack("<BBQQ", 3, 1, daddr, nbytes)),
            ]
            for k, v in node["attrs"].items():
                msgs.append(_attr_message(k, v))
            return write_object_header(msgs)

        def write_group(node):
            names = sorted(node["children"], key=lambda s: s.encode("utf-8"))
            child_addr = {}
            for nm in names:
                ch = node["children"][nm]
                child_addr[nm] = write_dataset(ch) if ch["data"] is not None else write_group(ch)[0]
            # local heap: offset 0 holds the empty string
            heap_data = bytearray(b"\x00" * 8)
            name_off = {}
            for nm in names:
                name_off[nm] = len(heap_data)
                enc = nm.encode("utf-8") + b"\x00"
                heap_data += enc + b"\x00" * (_pad8(len(enc)) - len(enc))
            free_off = len(heap_data)
            heap_data += struct.pack("<QQ", 1, 16)  # free block: next=1 (none), size 16
            seg_addr = alloc(bytes(heap_data))
            heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, seg_addr))
            # symbol nodes: at most 2K = 8 entries each with leaf K = 4
            leaf_k = 4
            per = 2 * leaf_k
            snods = []
            for i in range(0, max(len(names), 1), per):
                part = names[i:i + per]
                blob = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(part)))
                for nm in part:
                    blob += struct.pack("<QQII16x", name_off[nm], child_addr[nm], 0, 0)
                blob += b"\x00" * (8 + 40 * per - len(blob))
                snods.append((alloc(bytes(blob)), name_off[part[-1]] if part else 0))
            internal_k = 16
            if len(snods) > 2 * internal_k:
                raise NotImplementedError("too many links in one group")
            blob = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF))
            blob += struct.pack("<Q", 0)
            for addr, last_off in snods:
                blob += struct.pack("<QQ", addr, last_off)
            blob += b"\x00" * (24 + 8 + 16 * 2 * internal_k - len(blob))
            btree_addr = alloc(bytes(blob))
            msgs = [_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
            for k, v in node["attrs"].items():
                msgs.append(_attr_message(k, v))
            return write_object_header(msgs), btree_addr, heap_addr

        root_hdr, root_btree, root_heap = write_group(self._tree)
        eof = pos[0]
        sb = bytearray()
        sb += _SIG
        sb += struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
        os.pwrite(fd, bytes(sb), 0)
        for fut in pending:
            fut.result()
        writers.shutdown()
        os.close(fd)
        os.replace(tmp_path, path)
