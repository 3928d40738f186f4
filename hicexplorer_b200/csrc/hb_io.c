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

/* ------------------------------------------------------------------------------------------------------------------
 * Batch decode of the chunks of ONE 1-D chunked dataset with a pool of threads: what makes a genome-wide matrix file
 * written by PyTables (Blosc) or cooler (deflate + shuffle) load in seconds -- ~140 000 chunks of 64 KB each for a
 * 10 kb human matrix, one Python-level call per chunk before.
 *
 *   file_base            start of the memory-mapped file
 *   addr / csize / fmask per chunk: file offset, stored size, filter mask (bit i set = filter i was skipped)
 *   out_off / out_len    per chunk: where its decoded bytes go in `out` and how many of them (the last chunk is clipped)
 *   chunk_bytes          decoded size of a full chunk
 *   filters / typesizes  the dataset's filter pipeline in application order (HDF5 ids 1 deflate, 2 shuffle,
 *                        3 fletcher32, 32001 blosc) and, for shuffle, the element size
 * Returns 0, -1 malformed data, -2 a codec / filter this decoder does not handle (the caller falls back to Python).
 * ------------------------------------------------------------------------------------------------------------------ */
#include <pthread.h>
#include <stdlib.h>
#include <zlib.h>

typedef struct {
    const uint8_t* file_base;
    int64_t n_chunks;
    const int64_t* addr;
    const int64_t* csize;
    const uint32_t* fmask;
    const int64_t* out_off;
    const int64_t* out_len;
    int64_t chunk_bytes;
    const int32_t* filters;
    const int32_t* typesizes;
    int nfilters;
    uint8_t* out;
    int nthreads;
    int tid;
    int64_t status;
} hbio_job;

static void* hbio_worker(void* arg) {
    hbio_job* j = (hbio_job*)arg;
    const int64_t cap = j->chunk_bytes + 64;
    uint8_t* a = (uint8_t*)malloc((size_t)cap);
    uint8_t* b = (uint8_t*)malloc((size_t)cap);
    uint8_t* scratch = (uint8_t*)malloc((size_t)cap);
    j->status = 0;
    if (!a || !b || !scratch) j->status = -1;
    for (int64_t c = j->tid; c < j->n_chunks && j->status == 0; c += j->nthreads) {
        const uint8_t* cur = j->file_base + j->addr[c];
        int64_t len = j->csize[c];
        uint8_t* bufs[2] = {a, b};
        int which = 0;
        for (int f = j->nfilters - 1; f >= 0 && j->status == 0; --f) {
            if (j->fmask[c] & (1u << f)) continue;
            uint8_t* dst = bufs[which];
            const int id = j->filters[f];
            if (id == 32001) {
                if (len < 16 || hbio_blosc_nbytes(cur) > cap || hbio_blosc_blocksize(cur) > cap) {
                    j->status = -1;
                    break;
                }
                const int64_t got = hbio_blosc_decode(cur, len, dst, scratch);
                if (got < 0) {
                    j->status = got;
                    break;
                }
                len = got;
            } else if (id == 1) {
                uLongf dlen = (uLongf)cap;
                if (uncompress(dst, &dlen, cur, (uLong)len) != Z_OK) {
                    j->status = -1;
                    break;
                }
                len = (int64_t)dlen;
            } else if (id == 2) {
                if (len > cap) {
                    j->status = -1;
                    break;
                }
                hbio_unshuffle(cur, dst, len, j->typesizes[f]);
            } else if (id == 3) {
                if (len < 4) {
                    j->status = -1;
                    break;
                }
                len -= 4; /* checksum dropped, data untouched */
                continue;
            } else {
                j->status = -2;
                break;
            }
            cur = dst;
            which ^= 1;
        }
        if (j->status == 0) {
            if (len < j->out_len[c]) j->status = -1;
            else memcpy(j->out + j->out_off[c], cur, (size_t)j->out_len[c]);
        }
    }
    free(a);
    free(b);
    free(scratch);
    return NULL;
}

HBIO_API int64_t hbio_decode_chunks(const uint8_t* file_base, int64_t n_chunks, const int64_t* addr, const int64_t* csize,
                                    const uint32_t* fmask, const int64_t* out_off, const int64_t* out_len, int64_t chunk_bytes,
                                    const int32_t* filters, const int32_t* typesizes, int nfilters, uint8_t* out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int64_t)nthreads > n_chunks) nthreads = n_chunks > 0 ? (int)n_chunks : 1;
    hbio_job jobs[64];
    pthread_t th[64];
    for (int t = 0; t < nthreads; ++t) {
        hbio_job j = {file_base, n_chunks, addr, csize, fmask, out_off, out_len, chunk_bytes, filters, typesizes, nfilters,
                      out, nthreads, t, 0};
        jobs[t] = j;
    }
    int started = 0;
    for (int t = 1; t < nthreads; ++t) {
        if (pthread_create(&th[t], NULL, hbio_worker, &jobs[t]) != 0) break;
        started = t;
    }
    int64_t status = 0;
    if (started != nthreads - 1) {
        /* could not start every thread: let the started ones finish, then decode everything here */
        for (int t = 1; t <= started; ++t) pthread_join(th[t], NULL);
        hbio_job j = {file_base, n_chunks, addr, csize, fmask, out_off, out_len, chunk_bytes, filters, typesizes, nfilters,
                      out, 1, 0, 0};
        hbio_worker(&j);
        return j.status;
    }
    hbio_worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < nthreads; ++t)
        if (jobs[t].status != 0) status = jobs[t].status;
    return status;
}

/* ==================================================================================================================
 * Encoders: what the reference's writers put on disk (docs/content/file-formats.rst:9-31, 64-76):
 *   .h5    PyTables CArrays with the Blosc filter (id 32001): blosclz codec, level 5, byte shuffle
 *   .cool  cooler datasets with shuffle (id 2) + deflate (id 1)
 * The Blosc-1 container and the blosclz stream are produced in exactly the form blosclz_decode / hbio_blosc_decode
 * above (and c-blosc) read: 16-byte header, block offsets, every full block split into `typesize` byte-plane streams
 * (each: int32 compressed size + data; size == plane size means "stored"), 32 KB blocks like PyTables' files.
 * ================================================================================================================== */
#define HBIO_HASH_LOG 13
#define HBIO_MAX_NEAR 8191            /* distances 1..8191 fit the two-byte form */
#define HBIO_MAX_FAR (65535 + 8191)   /* far form: 16-bit offset on top of 8191 */

static inline uint32_t hbio_hash3(const uint8_t* p) {
    const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    return (v * 2654435761u) >> (32 - HBIO_HASH_LOG);
}

/* blosclz (FastLZ-style) compressor.  Returns the compressed size, or 0 when the result would not be smaller than
 * `maxout` (the caller then stores the stream verbatim). */
static int64_t blosclz_encode(const uint8_t* ip, int64_t length, uint8_t* op, int64_t maxout) {
    if (length < 16 || maxout < 66) return 0;
    int32_t htab[1 << HBIO_HASH_LOG];
    for (int i = 0; i < (1 << HBIO_HASH_LOG); ++i) htab[i] = -1;
    uint8_t* const op_start = op;
    uint8_t* const op_limit = op + maxout;
    int64_t anchor = 0; /* start of the pending literals */
    int64_t i = 0;
    const int64_t last = length - 3;
#define HBIO_FLUSH_LITERALS(upto)                                        \
    do {                                                                 \
        int64_t n_ = (upto)-anchor;                                      \
        while (n_ > 0) {                                                 \
            const int64_t run_ = n_ > 32 ? 32 : n_;                      \
            if (op + 1 + run_ > op_limit) return 0;                      \
            *op++ = (uint8_t)(run_ - 1);                                 \
            memcpy(op, ip + anchor, (size_t)run_);                       \
            op += run_;                                                  \
            anchor += run_;                                              \
            n_ -= run_;                                                  \
        }                                                                \
    } while (0)
    /* the stream has to open with a literal (the decoder masks the first control byte to 5 bits) */
    i = 1;
    const int64_t probe = length / 8 > 256 ? length / 8 : 256;   /* give up early on byte planes that do not compress */
    int probed = 0;
    while (i < last) {
        if (!probed && i >= probe) {
            probed = 1;
            /* (mantissa planes of floating-point values: no matches, the stream would be stored anyway) */
            if ((op - op_start) + (i - anchor) > i - i / 32) return 0;
        }
        const uint32_t h = hbio_hash3(ip + i);
        const int32_t cand = htab[h];
        htab[h] = (int32_t)i;
        int64_t dist = cand >= 0 ? i - cand : 0;
        if (cand >= 0 && dist <= HBIO_MAX_FAR && ip[cand] == ip[i] && ip[cand + 1] == ip[i + 1] && ip[cand + 2] == ip[i + 2]) {
            int64_t len = 3;
            while (i + len < length && ip[cand + len] == ip[i + len]) ++len;
            HBIO_FLUSH_LITERALS(i);
            /* copy length = field + 3 ; field <= 5 sits in the control byte, larger ones continue in 255-steps */
            int64_t field = len - 3;
            const int64_t ofs = dist - 1;
            const int far = ofs >= HBIO_MAX_NEAR;   /* 8191 itself would read as the far marker */
            const uint32_t hi = far ? 31u : (uint32_t)(ofs >> 8);
            if (op + 8 + field / 255 > op_limit) return 0;
            if (field < 6) {
                *op++ = (uint8_t)(((field + 1) << 5) | hi);
            } else {
                *op++ = (uint8_t)((7u << 5) | hi);
                field -= 6;
                while (field >= 255) {
                    *op++ = 255;
                    field -= 255;
                }
                *op++ = (uint8_t)field;
            }
            if (far) {
                const int64_t f = ofs - HBIO_MAX_NEAR;
                *op++ = 255;
                *op++ = (uint8_t)(f >> 8);
                *op++ = (uint8_t)(f & 255);
            } else {
                *op++ = (uint8_t)(ofs & 255);
            }
            /* index a few positions inside the match so that later data can refer to them */
            const int64_t end = i + len;
            for (int64_t k = i + 1; k < end && k < last; k += (len > 64 ? 16 : 1)) htab[hbio_hash3(ip + k)] = (int32_t)k;
            i = end;
            anchor = i;
        } else {
            ++i;
        }
    }
    HBIO_FLUSH_LITERALS(length);
#undef HBIO_FLUSH_LITERALS
    const int64_t out = op - op_start;
    return out < maxout ? out : 0;
}

static void shuffle_bytes(const uint8_t* src, uint8_t* dst, int64_t nbytes, int typesize) {
    const int64_t nelem = nbytes / typesize;
    for (int t = 0; t < typesize; ++t) {
        uint8_t* plane = dst + (int64_t)t * nelem;
        for (int64_t i = 0; i < nelem; ++i) plane[i] = src[i * typesize + t];
    }
    memcpy(dst + nelem * typesize, src + nelem * typesize, (size_t)(nbytes - nelem * typesize));
}

HBIO_API int64_t hbio_blosc_bound(int64_t nbytes, int typesize) {
    const int64_t blocksize = 32768;
    const int64_t nblocks = (nbytes + blocksize - 1) / blocksize;
    return 16 + 4 * nblocks + nbytes + 4 * nblocks * (typesize > 0 ? typesize : 1) + 64;
}

/* One Blosc-1 chunk (blosclz, byte shuffle).  `scratch` holds one block.  Returns the chunk size. */
HBIO_API int64_t hbio_blosc_encode(const uint8_t* src, int64_t nbytes, int typesize, uint8_t* out, uint8_t* scratch) {
    int64_t blocksize = 32768;
    if (blocksize > nbytes) blocksize = nbytes;
    if (typesize > 1 && blocksize >= typesize) blocksize -= blocksize % typesize;
    const int doshuffle = typesize > 1;
    const int rule = typesize <= 16 && blocksize / (typesize > 0 ? typesize : 1) >= 128;
    const int64_t nblocks = nbytes > 0 ? (nbytes + blocksize - 1) / blocksize : 0;
    uint8_t flags = (uint8_t)((doshuffle ? 0x1 : 0x0) | (rule ? 0x0 : 0x10)); /* codec 0 = blosclz in bits 5-7 */
    int64_t pos = 16 + 4 * nblocks;
    int ok = nbytes >= 128;   /* tiny buffers are stored */
    for (int64_t b = 0; ok && b < nblocks; ++b) {
        int64_t bsize = blocksize;
        if ((b + 1) * blocksize > nbytes) bsize = nbytes - b * blocksize;
        const int leftover = bsize != blocksize;
        const uint8_t* blk = src + b * blocksize;
        if (doshuffle) {
            shuffle_bytes(blk, scratch, bsize, typesize);
            blk = scratch;
        }
        const int nsplits = (rule && !leftover) ? typesize : 1;
        const int64_t neblock = bsize / nsplits;
        const int32_t bstart = (int32_t)pos;
        memcpy(out + 16 + 4 * b, &bstart, 4);
        for (int s = 0; s < nsplits; ++s) {
            const uint8_t* sp = blk + (int64_t)s * neblock;
            int64_t c = blosclz_encode(sp, neblock, out + pos + 4, neblock - 1);
            if (c <= 0) {   /* not compressible: stored (size == plane size tells the decoder) */
                memcpy(out + pos + 4, sp, (size_t)neblock);
                c = neblock;
            }
            const int32_t c32 = (int32_t)c;
            memcpy(out + pos, &c32, 4);
            pos += 4 + c;
        }
        if (pos >= nbytes + 16) ok = 0;   /* no gain: fall back to a plain copy */
    }
    if (!ok) {
        flags = (uint8_t)(0x2 | (doshuffle ? 0x1 : 0x0));   /* memcpyed */
        memcpy(out + 16, src, (size_t)nbytes);
        pos = 16 + nbytes;
    }
    out[0] = 2;   /* BLOSC_VERSION_FORMAT */
    out[1] = 1;   /* blosclz format version */
    out[2] = flags;
    out[3] = (uint8_t)typesize;
    const uint32_t nb = (uint32_t)nbytes, bs = (uint32_t)blocksize, cb = (uint32_t)pos;
    memcpy(out + 4, &nb, 4);
    memcpy(out + 8, &bs, 4);
    memcpy(out + 12, &cb, 4);
    return pos;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Batch encode of consecutive chunks of one 1-D dataset on a pool of threads.
 *   src / nbytes   the dataset's bytes; chunk c covers [c * chunk_bytes, (c+1) * chunk_bytes) and is zero-padded to a
 *                  full chunk at the end of the dataset (HDF5 chunks always have the full shape)
 *   mode           1 = Blosc(blosclz + shuffle)   2 = shuffle + deflate   3 = deflate
 *   slots          n_chunks * slot_bytes scratch (slot_bytes >= hbio_encode_bound): every chunk is encoded into its slot
 *   dst            receives the encoded chunks back to back; offsets[c] / sizes[c] locate chunk c inside dst
 * Returns the total number of bytes in dst, or -1.
 * ------------------------------------------------------------------------------------------------------------------ */
HBIO_API int64_t hbio_encode_bound(int64_t chunk_bytes, int typesize, int mode) {
    if (mode == 1) return hbio_blosc_bound(chunk_bytes, typesize);
    return (int64_t)compressBound((uLong)chunk_bytes) + 64;
}

typedef struct {
    const uint8_t* src;
    int64_t nbytes, chunk_bytes, n_chunks, slot_bytes;
    int typesize, mode, level, nthreads, tid;
    uint8_t* slots;
    int64_t* sizes;
    int64_t status;
} hbio_enc_job;

static void* hbio_enc_worker(void* arg) {
    hbio_enc_job* j = (hbio_enc_job*)arg;
    uint8_t* padded = (uint8_t*)malloc((size_t)j->chunk_bytes + 64);
    uint8_t* scratch = (uint8_t*)malloc((size_t)j->chunk_bytes + 64);
    j->status = (padded && scratch) ? 0 : -1;
    for (int64_t c = j->tid; c < j->n_chunks && j->status == 0; c += j->nthreads) {
        const int64_t off = c * j->chunk_bytes;
        const uint8_t* in = j->src + off;
        if (off + j->chunk_bytes > j->nbytes) {
            const int64_t have = j->nbytes - off;
            memcpy(padded, in, (size_t)have);
            memset(padded + have, 0, (size_t)(j->chunk_bytes - have));
            in = padded;
        }
        uint8_t* out = j->slots + c * j->slot_bytes;
        if (j->mode == 1) {
            j->sizes[c] = hbio_blosc_encode(in, j->chunk_bytes, j->typesize, out, scratch);
        } else {
            const uint8_t* zin = in;
            if (j->mode == 2 && j->typesize > 1) {
                shuffle_bytes(in, scratch, j->chunk_bytes, j->typesize);
                zin = scratch;
            }
            uLongf dlen = (uLongf)j->slot_bytes;
            if (compress2(out, &dlen, zin, (uLong)j->chunk_bytes, j->level) != Z_OK) j->status = -1;
            j->sizes[c] = (int64_t)dlen;
        }
    }
    free(padded);
    free(scratch);
    return NULL;
}

HBIO_API int64_t hbio_encode_chunks(const uint8_t* src, int64_t nbytes, int64_t chunk_bytes, int typesize, int mode, int level,
                                    int nthreads, uint8_t* slots, int64_t slot_bytes, uint8_t* dst, int64_t* offsets,
                                    int64_t* sizes) {
    if (chunk_bytes <= 0 || nbytes < 0 || mode < 1 || mode > 3) return -1;
    const int64_t n_chunks = (nbytes + chunk_bytes - 1) / chunk_bytes;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int64_t)nthreads > n_chunks) nthreads = n_chunks > 0 ? (int)n_chunks : 1;
    hbio_enc_job jobs[64];
    pthread_t th[64];
    int started[64] = {0};
    for (int t = 0; t < nthreads; ++t) {
        hbio_enc_job j = {src, nbytes, chunk_bytes, n_chunks, slot_bytes, typesize, mode, level, nthreads, t, slots, sizes, 0};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) started[t] = pthread_create(&th[t], NULL, hbio_enc_worker, &jobs[t]) == 0;
    hbio_enc_worker(&jobs[0]);
    int64_t status = jobs[0].status;
    for (int t = 1; t < nthreads; ++t) {
        if (started[t]) pthread_join(th[t], NULL);
        else hbio_enc_worker(&jobs[t]);   /* could not start the thread: do its share here */
        if (jobs[t].status != 0) status = jobs[t].status;
    }
    if (status != 0) return -1;
    int64_t pos = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        offsets[c] = pos;
        memcpy(dst + pos, slots + c * slot_bytes, (size_t)sizes[c]);
        pos += sizes[c];
    }
    return pos;
}


/* ---- is a CSR canonical?  (rows sorted by column, no duplicates, columns inside [0, n_cols)) --------------------
 * scipy's own check (csr_has_canonical_format) is one thread walking the whole index array: 0.43 s for the upper
 * triangle of a genome-wide 10 kb matrix, on the critical path of every load.  Same test on all host cores.
 * Returns 1 canonical, 0 not, -1 bad arguments.  indices: int32 (idx_bytes 4) or int64 (8); indptr: int32 or int64. */
typedef struct {
    const void* indptr;
    const void* indices;
    int ptr_bytes, idx_bytes;
    int64_t n_rows, n_cols;
    int nthreads, tid;
    int ok;
} hbio_canon_job;

static void* hbio_canon_worker(void* arg) {
    hbio_canon_job* j = (hbio_canon_job*)arg;
    const int64_t r0 = j->n_rows * j->tid / j->nthreads, r1 = j->n_rows * (j->tid + 1) / j->nthreads;
    int ok = 1;
    for (int64_t r = r0; r < r1 && ok; ++r) {
        const int64_t s = j->ptr_bytes == 8 ? ((const int64_t*)j->indptr)[r] : ((const int32_t*)j->indptr)[r];
        const int64_t e = j->ptr_bytes == 8 ? ((const int64_t*)j->indptr)[r + 1] : ((const int32_t*)j->indptr)[r + 1];
        if (e < s) {
            ok = 0;
            break;
        }
        if (e == s) continue;
        if (j->idx_bytes == 4) {
            const int32_t* x = (const int32_t*)j->indices;
            int bad = x[s] < 0 || (int64_t)x[e - 1] >= j->n_cols;
            for (int64_t k = s + 1; k < e; ++k) bad |= x[k - 1] >= x[k];
            if (bad) ok = 0;
        } else {
            const int64_t* x = (const int64_t*)j->indices;
            int bad = x[s] < 0 || x[e - 1] >= j->n_cols;
            for (int64_t k = s + 1; k < e; ++k) bad |= x[k - 1] >= x[k];
            if (bad) ok = 0;
        }
    }
    j->ok = ok;
    return NULL;
}

HBIO_API int hbio_csr_is_canonical(const void* indptr, int ptr_bytes, const void* indices, int idx_bytes, int64_t n_rows,
                                   int64_t n_cols, int64_t nnz, int nthreads) {
    if ((ptr_bytes != 4 && ptr_bytes != 8) || (idx_bytes != 4 && idx_bytes != 8) || n_rows < 0 || indptr == NULL) return -1;
    if (n_rows == 0) return 1;
    const int64_t first = ptr_bytes == 8 ? ((const int64_t*)indptr)[0] : ((const int32_t*)indptr)[0];
    const int64_t last = ptr_bytes == 8 ? ((const int64_t*)indptr)[n_rows] : ((const int32_t*)indptr)[n_rows];
    if (first != 0 || last != nnz) return 0;
    if (nnz > 0 && indices == NULL) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int64_t)nthreads > n_rows) nthreads = (int)n_rows;
    /* indptr must be monotone as a whole before the workers index through it */
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t a = ptr_bytes == 8 ? ((const int64_t*)indptr)[r] : ((const int32_t*)indptr)[r];
        const int64_t b = ptr_bytes == 8 ? ((const int64_t*)indptr)[r + 1] : ((const int32_t*)indptr)[r + 1];
        if (b < a || a < 0 || b > nnz) return 0;
    }
    hbio_canon_job jobs[64];
    pthread_t th[64];
    int started[64] = {0};
    for (int t = 0; t < nthreads; ++t) {
        hbio_canon_job j = {indptr, indices, ptr_bytes, idx_bytes, n_rows, n_cols, nthreads, t, 1};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) started[t] = pthread_create(&th[t], NULL, hbio_canon_worker, &jobs[t]) == 0;
    hbio_canon_worker(&jobs[0]);
    int ok = jobs[0].ok;
    for (int t = 1; t < nthreads; ++t) {
        if (started[t]) pthread_join(th[t], NULL);
        else hbio_canon_worker(&jobs[t]);
        ok &= jobs[t].ok;
    }
    return ok;
}
