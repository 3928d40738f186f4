/*
 * Host-side decoders for the on-disk formats of the path (file I/O, not the device hot path):
 * Blosc-1 chunks with the blosclz codec + byte shuffle, as written by PyTables for HiCExplorer .h5 matrices
 * (docs/content/file-formats.rst:64-76 of the reference; container layout restated in SURVEY.md Appendix A).
 * hdf5lite.py holds the same decoders in pure Python; this file makes them usable on genome-wide files
 * (the Python blosclz loop runs at a few MB/s).  Built by _build.py with gcc into libhbio.so.
 */
#include <stdint.h>
#include <string.h>

#if defined(__GNUC__)
#define HBIO_API __attribute__((visibility("default")))
#else
#define HBIO_API
#endif

/* FastLZ-style stream used by blosclz.  Returns bytes written, or -1 on a malformed stream. */
static int64_t blosclz_decode(const uint8_t* ip, int64_t length, uint8_t* op, int64_t maxout) {
    const uint8_t* const ip_end = ip + length;
    uint8_t* const op_start = op;
    uint8_t* const op_end = op + maxout;
    if (length <= 0) return 0;
    uint32_t ctrl = (*ip++) & 31;
    for (;;) {
        if (ctrl >= 32) {
            int64_t len = (int64_t)(ctrl >> 5) - 1;
            int64_t ofs = (int64_t)(ctrl & 31) << 8;
            uint8_t code;
            if (len == 6) {
                do {
                    if (ip >= ip_end) return -1;
                    code = *ip++;
                    len += code;
                } while (code == 255);
            }
            if (ip >= ip_end) return -1;
            code = *ip++;
            const uint8_t* ref = op - ofs - code;
            if (code == 255 && ofs == (31 << 8)) {
                if (ip + 1 >= ip_end) return -1;
                ofs = ((int64_t)ip[0] << 8) + ip[1];
                ip += 2;
                ref = op - ofs - 8191;
            }
            int more = ip < ip_end;
            if (more) ctrl = *ip++;
            len += 3;
            ref -= 1;
            if (ref < op_start || op + len > op_end) return -1;
            for (int64_t i = 0; i < len; ++i) op[i] = ref[i]; /* byte-wise: overlapping copies repeat the pattern */
            op += len;
            if (!more) break;
        } else {
            ctrl++;
            if (ip + ctrl > ip_end || op + ctrl > op_end) return -1;
            memcpy(op, ip, ctrl);
            op += ctrl;
            ip += ctrl;
            if (ip >= ip_end) break;
            ctrl = *ip++;
        }
    }
    return op - op_start;
}

static void unshuffle(const uint8_t* src, uint8_t* dst, int64_t nbytes, int typesize) {
    const int64_t nelem = nbytes / typesize;
    for (int t = 0; t < typesize; ++t) {
        const uint8_t* plane = src + (int64_t)t * nelem;
        for (int64_t i = 0; i < nelem; ++i) dst[i * typesize + t] = plane[i];
    }
    memcpy(dst + nelem * typesize, src + nelem * typesize, (size_t)(nbytes - nelem * typesize));
}

/*
 * Decode one Blosc-1 chunk (16-byte header + block offsets + blocks).  `out` must hold the `nbytes` announced in
 * the header (read it with hbio_blosc_nbytes).  `scratch` must hold one block (blocksize bytes).
 * Returns nbytes, -1 malformed, -2 unsupported codec (only blosclz = 0 and stored blocks are handled here).
 */
HBIO_API int64_t hbio_blosc_nbytes(const uint8_t* chunk) {
    uint32_t nbytes;
    memcpy(&nbytes, chunk + 4, 4);
    return (int64_t)nbytes;
}

HBIO_API int64_t hbio_blosc_blocksize(const uint8_t* chunk) {
    uint32_t bs;
    memcpy(&bs, chunk + 8, 4);
    return (int64_t)bs;
}

HBIO_API int64_t hbio_blosc_decode(const uint8_t* chunk, int64_t chunk_len, uint8_t* out, uint8_t* scratch) {
    if (chunk_len < 16) return -1;
    const uint8_t flags = chunk[2];
    const int typesize = chunk[3];
    uint32_t nbytes, blocksize, cbytes;
    memcpy(&nbytes, chunk + 4, 4);
    memcpy(&blocksize, chunk + 8, 4);
    memcpy(&cbytes, chunk + 12, 4);
    if (flags & 0x2) { /* memcpyed */
        if ((int64_t)nbytes + 16 > chunk_len) return -1;
        memcpy(out, chunk + 16, nbytes);
        return nbytes;
    }
    if (flags & 0x4) return -2; /* bit shuffle */
    const int codec = flags >> 5;
    const int doshuffle = flags & 0x1;
    const int dont_split = flags & 0x10;
    if (blocksize == 0) return -1;
    const int64_t nblocks = ((int64_t)nbytes + blocksize - 1) / blocksize;
    if (16 + 4 * nblocks > chunk_len) return -1;
    for (int64_t b = 0; b < nblocks; ++b) {
        int32_t bstart;
        memcpy(&bstart, chunk + 16 + 4 * b, 4);
        int64_t bsize = blocksize;
        if ((b + 1) * (int64_t)blocksize > (int64_t)nbytes) bsize = (int64_t)nbytes - b * (int64_t)blocksize;
        const int leftover = bsize != (int64_t)blocksize;
        const int split = (!dont_split && !leftover && typesize <= 16 && (int64_t)blocksize / typesize >= 128);
        const int nsplits = split ? typesize : 1;
        const int64_t neblock = bsize / nsplits;
        int64_t pos = bstart;
        uint8_t* dst = doshuffle ? scratch : out + b * (int64_t)blocksize;
        int64_t w = 0;
        for (int sidx = 0; sidx < nsplits; ++sidx) {
            int32_t csize;
            if (pos + 4 > chunk_len) return -1;
            memcpy(&csize, chunk + pos, 4);
            pos += 4;
            if (csize < 0 || pos + csize > chunk_len) return -1;
            if (csize == neblock) {
                memcpy(dst + w, chunk + pos, (size_t)csize);
                w += csize;
            } else if (codec == 0) {
                int64_t got = blosclz_decode(chunk + pos, csize, dst + w, neblock);
                if (got != neblock) return -1;
                w += got;
            } else {
                return -2;
            }
            pos += csize;
        }
        if (doshuffle) unshuffle(scratch, out + b * (int64_t)blocksize, bsize, typesize);
    }
    return nbytes;
}

/* HDF5 shuffle filter (id 2): the inverse byte transpose over one chunk */
HBIO_API void hbio_unshuffle(const uint8_t* src, uint8_t* dst, int64_t nbytes, int typesize) {
    if (typesize <= 1) {
        memcpy(dst, src, (size_t)nbytes);
        return;
    }
    unshuffle(src, dst, nbytes, typesize);
}
