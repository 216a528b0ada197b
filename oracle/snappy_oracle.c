/*
 * snappy_oracle.c -- CPU restatement of Snappier's Snappy block path.
 * TEST INFRASTRUCTURE ONLY (see snappy_oracle.h for who may call this and for
 * the parity status: MUL-hash compress + decompress pinned by reference
 * fixtures, CRC32C-hash compress "parity unpinned").
 *
 * Written from the behaviour of the reference, not transliterated: the
 * reference's 16x-unrolled probe prefix, preload registers, 16-byte blind
 * copies and the resumable/split-input decoder state are all speed or
 * streaming devices that do not change a single output byte, so they are
 * absent here.  Citations are /root/reference/Snappier/Internal/<file>:<line>.
 */
#include "snappy_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t le32(const uint8_t *p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v; /* host is little endian (x86-64 / aarch64) */
}

/* ------------------------------------------------------------------ sizing */

int32_t orc_max_compressed_length(int32_t n) { return 32 + n + n / 6 + 1; } /* Helpers.cs:45 */
int32_t orc_get_max_compressed_length(int32_t n) {                          /* Snappy.cs:20-24 */
    return orc_max_compressed_length(n) + 5;
}

/* ------------------------------------------------------------------ varint */

int orc_varint_write(uint8_t *out, size_t cap, uint32_t v) { /* VarIntEncoding.Write.cs:5-79 */
    int need = v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5;
    if (cap < (size_t)need) return 0;
    for (int i = 0; i < need - 1; i++) {
        out[i] = (uint8_t)(v | 0x80);
        v >>= 7;
    }
    out[need - 1] = (uint8_t)v;
    return need;
}

int orc_varint_read(const uint8_t *in, size_t n, uint32_t *v, int *consumed) {
    /* VarIntEncoding.Read.cs:38-79 */
    uint32_t result = 0;
    int shift = 0;
    *consumed = 0;
    *v = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = in[i];
        uint32_t val = c & 0x7f;
        /* Helpers.LeftShiftOverflows(val, shift), Helpers.cs:66-70 */
        if (val & ~(0xffffffffu >> shift)) return ORC_INVALID_LENGTH;
        result |= val << shift;
        shift += 7;
        if (c < 128) {
            *v = result;
            *consumed = (int)i + 1;
            return ORC_OK;
        }
        if (shift >= 32) return ORC_INVALID_LENGTH;
    }
    return ORC_INCOMPLETE; /* OperationStatus.NeedMoreData */
}

/* -------------------------------------------------------------- hash table */

int orc_table_size(int n) { /* HashTable.cs:14-18,57-71 */
    if (n > 16384) return 16384;
    if (n < 256) return 256;
    int lg = 31 - __builtin_clz((unsigned)(n - 1));
    return 2 << lg;
}

static uint32_t crc_tab[256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) { /* Crc32CAlgorithm.cs:15,22-36 (first 256 entries) */
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t r = i;
        for (int k = 0; k < 8; k++) r = (r & 1) ? 0x82F63B78u ^ (r >> 1) : (r >> 1);
        crc_tab[i] = r;
    }
}

/* Architectural semantics of SSE4.2 `crc32 r32, r32` / ARM `crc32cw`
 * (what Sse42.Crc32(crc, data) and Crc32.ComputeCrc32C(crc, data) compile to):
 * four table rounds over crc ^ data, no init / final xor. */
static inline uint32_t crc32c_u32(uint32_t crc, uint32_t data) {
    uint32_t y = crc ^ data;
    for (int i = 0; i < 4; i++) y = crc_tab[y & 0xff] ^ (y >> 8);
    return y;
}

/* Table-driven form: the definition.  Exposed for the tests, which check it
 * against the hardware instruction below on every value they try. */
uint32_t orc_table_hash(uint32_t bytes, uint32_t mask, int hash_mode) { /* HashTable.cs:91-126 */
    pthread_once(&crc_once, crc_init);
    uint32_t h;
    if (hash_mode == ORC_HASH_CRC32C)
        h = crc32c_u32(bytes, mask); /* HashTable.cs:109-117 */
    else
        h = (0x1e35a7bdu * bytes) >> (31 - 14); /* HashTable.cs:120-123 */
    return h & mask;
}

/* Hot-loop form.  On x86 with SSE4.2 it issues the very instruction
 * Sse42.Crc32 compiles to, so the CPU baseline is not handicapped by table
 * lookups; elsewhere it falls back to the table (crc_init must have run). */
static inline uint32_t hash_fast(uint32_t bytes, uint32_t mask, int hash_mode) {
    uint32_t h;
    if (hash_mode == ORC_HASH_CRC32C) {
#if defined(__SSE4_2__)
        h = __builtin_ia32_crc32si(bytes, mask);
#else
        h = crc32c_u32(bytes, mask);
#endif
    } else {
        h = (0x1e35a7bdu * bytes) >> (31 - 14);
    }
    return h & mask;
}

int orc_hash_uses_hw_crc(void) {
#if defined(__SSE4_2__)
    return 1;
#else
    return 0;
#endif
}
uint32_t orc_table_hash_fast(uint32_t bytes, uint32_t mask, int hash_mode) {
    pthread_once(&crc_once, crc_init);
    return hash_fast(bytes, mask, hash_mode);
}

/* ---------------------------------------------------------------- compress */

int orc_find_match_length(const uint8_t *s1, const uint8_t *s2, const uint8_t *s2_limit) {
    /* SnappyCompressor.cs:562-688: the 8-byte compare loops and the `data`
     * side-output are speed devices; the value is the bounded common prefix. */
    int m = 0;
    while (s2_limit - (s2 + m) >= 8) {
        uint64_t a, b;
        memcpy(&a, s1 + m, 8);
        memcpy(&b, s2 + m, 8);
        if (a != b) return m + (__builtin_ctzll(a ^ b) >> 3);
        m += 8;
    }
    while (s2 + m < s2_limit && s1[m] == s2[m]) m++;
    return m;
}

static uint8_t *emit_literal(uint8_t *op, const uint8_t *lit, size_t len) {
    /* SnappyCompressor.cs:418-464 */
    uint32_t n = (uint32_t)len - 1;
    if (n < 60) {
        *op++ = (uint8_t)(n << 2);
    } else {
        int count = ((31 - __builtin_clz(n)) >> 3) + 1;
        *op++ = (uint8_t)((59 + count) << 2);
        for (int i = 0; i < count; i++) *op++ = (uint8_t)(n >> (8 * i));
    }
    memcpy(op, lit, len);
    return op + len;
}

static uint8_t *emit_copy_upto64(uint8_t *op, uint32_t offset, uint32_t len) {
    /* SnappyCompressor.cs:467-505 */
    if (len < 12 && offset < 2048) {
        *op++ = (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 8) << 5));
        *op++ = (uint8_t)offset;
    } else {
        *op++ = (uint8_t)(2 + ((len - 1) << 2));
        *op++ = (uint8_t)offset;
        *op++ = (uint8_t)(offset >> 8);
    }
    return op;
}

static uint8_t *emit_copy(uint8_t *op, uint32_t offset, uint32_t len) {
    /* SnappyCompressor.cs:507-543 */
    while (len >= 68) {
        op = emit_copy_upto64(op, offset, 64);
        len -= 64;
    }
    if (len > 64) {
        op = emit_copy_upto64(op, offset, 60);
        len -= 60;
    }
    return emit_copy_upto64(op, offset, len);
}

size_t orc_compress_fragment(const uint8_t *in, size_t n, uint8_t *out, int hash_mode) {
    /* SnappyCompressor.cs:174-415 */
    pthread_once(&crc_once, crc_init);
    uint16_t table[16384];
    int tsize = orc_table_size((int)n);
    memset(table, 0, (size_t)tsize * 2); /* HashTable.cs:52 */
    uint32_t mask = 2u * (uint32_t)(tsize - 1); /* :181 */
    uint8_t *op = out;
    size_t ip = 0;

    if (n >= 15) { /* Constants.InputMarginBytes, :190 */
        size_t ip_limit = n - 15;
        for (;;) {
            size_t next_emit = ip;
            ip += 1;
            uint32_t skip = 32; /* :227 */
            size_t cand;
            for (;;) { /* probe loop, :230-341 (unrolled prefix is behaviour-identical) */
                uint32_t x = le32(in + ip);
                uint32_t stride = skip >> 5;
                skip += stride;
                size_t nip = ip + stride;
                if (nip > ip_limit) { /* :323-327 */
                    ip = next_emit;
                    goto remainder;
                }
                uint16_t *slot = &table[hash_fast(x, mask, hash_mode) >> 1];
                cand = *slot;
                *slot = (uint16_t)ip; /* :333 (write precedes compare) */
                if (le32(in + cand) == x) break;
                ip = nip;
            }
            op = emit_literal(op, in + next_emit, ip - next_emit); /* :347 */
            for (;;) { /* emit_match, :358-398 */
                size_t base = ip;
                int m = 4 + orc_find_match_length(in + cand + 4, in + ip + 4, in + n);
                ip += (size_t)m;
                op = emit_copy(op, (uint32_t)(base - cand), (uint32_t)m);
                if (ip >= ip_limit) goto remainder; /* :381-384 */
                table[hash_fast(le32(in + ip - 1), mask, hash_mode) >> 1] =
                    (uint16_t)(ip - 1); /* :393-394 */
                uint32_t x = le32(in + ip);
                uint16_t *slot = &table[hash_fast(x, mask, hash_mode) >> 1];
                cand = *slot;
                *slot = (uint16_t)ip;
                if (le32(in + cand) != x) break; /* :398 */
            }
        }
    }
remainder:
    if (ip < n) op = emit_literal(op, in + ip, n - ip); /* :406-411 */
    return (size_t)(op - out);
}

int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written,
                 int hash_mode) {
    /* Snappy.cs:55-67 + SnappyCompressor.cs:24-83 */
    *written = 0;
    if (cap == 0) return ORC_OUTPUT_TOO_SMALL; /* Snappy.cs:57-62 */
    size_t w = (size_t)orc_varint_write(out, cap, (uint32_t)n);
    if (w == 0) return ORC_OUTPUT_TOO_SMALL;
    uint8_t *scratch = NULL;
    while (n > 0) {
        size_t frag = n < ORC_BLOCK_SIZE ? n : ORC_BLOCK_SIZE;
        size_t max_out = (size_t)orc_max_compressed_length((int32_t)frag);
        if (cap - w >= max_out) { /* :49-55 */
            w += orc_compress_fragment(in, frag, out + w, hash_mode);
        } else { /* :56-74 */
            if (!scratch) scratch = (uint8_t *)malloc((size_t)orc_max_compressed_length(ORC_BLOCK_SIZE));
            size_t c = orc_compress_fragment(in, frag, scratch, hash_mode);
            if (cap - w < c) {
                free(scratch);
                *written = 0;
                return ORC_OUTPUT_TOO_SMALL;
            }
            memcpy(out + w, scratch, c);
            w += c;
        }
        in += frag;
        n -= frag;
    }
    free(scratch);
    *written = w;
    return ORC_OK;
}

/* -------------------------------------------------------------- decompress */

int orc_uncompressed_length(const uint8_t *in, size_t n, uint32_t *len) {
    /* VarIntEncoding.Read.cs:16-24: anything but Done -> "Invalid stream length" */
    int used;
    int st = orc_varint_read(in, n, len, &used);
    if (st != ORC_OK) return ORC_INVALID_LENGTH;
    if (*len > 0x7fffffffu) return ORC_INVALID_LENGTH; /* SURVEY App. C Q4 */
    return ORC_OK;
}

int orc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written) {
    *written = 0;
    uint32_t ulen;
    int used;
    int st = orc_varint_read(in, n, &ulen, &used); /* SnappyDecompressor.cs:50-63 */
    if (st == ORC_INCOMPLETE) return ORC_INCOMPLETE; /* :57-60 then Snappy.cs:178-181 */
    if (st != ORC_OK || ulen > 0x7fffffffu) return ORC_INVALID_LENGTH;
    if (cap < ulen) return ORC_OUTPUT_TOO_SMALL;
    if (ulen == 0) return ORC_OK; /* AllDataDecompressed before any tag, :78 */

    const uint8_t *ip = in + used, *end = in + n;
    size_t op = 0, U = ulen;
    while (ip < end) { /* DecompressAllTags, :234-341 */
        uint8_t c = *ip;
        uint32_t kind = c & 3;
        /* bytes after the tag byte that belong to the tag itself (CharTable bits 11..13) */
        size_t extra = kind == 0 ? ((c >> 2) >= 60 ? (size_t)(c >> 2) - 59 : 0)
                                 : (kind == 1 ? 1 : kind == 2 ? 2 : 4);
        if ((size_t)(end - ip) < 1 + extra) break; /* RefillTag -> truncated tag, :464-483 */
        uint64_t trailer = 0;
        if ((size_t)(end - ip) >= 5) { /* one unaligned load instead of a byte loop (the reference's preload) */
            uint32_t t32;
            memcpy(&t32, ip + 1, 4);
            trailer = extra == 4 ? t32 : (t32 & ((1u << (8 * extra)) - 1u));
        } else {
            for (size_t i = 0; i < extra; i++) trailer |= (uint64_t)ip[1 + i] << (8 * i);
        }
        ip += 1 + extra;
        if (kind == 0) {
            uint64_t len = ((c >> 2) >= 60 ? trailer : (uint64_t)(c >> 2)) + 1; /* :264-288 */
            uint64_t avail = (uint64_t)(end - ip);
            /* TryFastAppend (:579-589): short literal, 16 bytes of slack on both sides */
            if (len <= 16 && avail >= 16 + 5 && U - op >= 16) {
                memcpy(out + op, ip, 16);
                op += (size_t)len;
                ip += len;
                continue;
            }
            uint64_t take = len < avail ? len : avail; /* :290-297 partial literal */
            if (take > U - op) return ORC_DATA_TOO_LONG;  /* :570-573 */
            memcpy(out + op, ip, (size_t)take);
            op += (size_t)take;
            ip += take;
            if (take < len) break;
        } else {
            uint64_t len, offset;
            if (kind == 1) { /* :316-325 via CharTable */
                len = ((c >> 2) & 7) + 4;
                offset = ((uint64_t)(c >> 5) << 8) | trailer;
            } else {
                len = (c >> 2) + 1;
                offset = trailer;
            }
            if (offset == 0 || op < offset) return ORC_INVALID_COPY_OFFSET; /* :598-601 */
            if (len > U - op) return ORC_DATA_TOO_LONG;                      /* :603-606 */
            /* CopyHelpers.cs:64-230: forward byte order (pattern replication); the wide paths are
             * speed devices with identical results */
            uint8_t *d = out + op;
            const uint8_t *s = d - offset;
            if (offset >= 8 && U - op >= len + 8) { /* 8 bytes at a time never reads unwritten bytes */
                for (uint64_t k = 0; k < len; k += 8) memcpy(d + k, s + k, 8);
            } else if (offset >= len) {
                memcpy(d, s, (size_t)len);
            } else {
                for (uint64_t k = 0; k < len; k++) d[k] = s[k];
            }
            op += (size_t)len;
        }
    }
    if (op < U) return ORC_INCOMPLETE; /* Snappy.cs:178-181 */
    *written = op;
    return ORC_OK;
}

/* ------------------------------------------------------------------ crc32c */

uint32_t orc_crc32c(uint32_t crc, const uint8_t *p, size_t n) { /* Crc32CAlgorithm.cs:46-155 */
    pthread_once(&crc_once, crc_init);
    uint32_t c = ~crc;
    for (size_t i = 0; i < n; i++) c = crc_tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return ~c;
}

uint32_t orc_crc32c_mask(uint32_t x) { /* Crc32CAlgorithm.cs:157-158 */
    return ((x >> 15) | (x << 17)) + 0xa282ead8u;
}

/* ------------------------------------------------- multi-threaded drivers */

typedef struct {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint32_t *in_len;
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint32_t *out_cap;
    uint32_t *out_len;
    int32_t *status;
    size_t lo, hi;
    int hash_mode, compress, first_bad;
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        size_t w = 0;
        int st;
        if (j->compress)
            st = orc_compress(j->in_base + j->in_off[i], j->in_len[i], j->out_base + j->out_off[i],
                              j->out_cap[i], &w, j->hash_mode);
        else
            st = orc_decompress(j->in_base + j->in_off[i], j->in_len[i],
                                j->out_base + j->out_off[i], j->out_cap[i], &w);
        j->out_len[i] = (uint32_t)w;
        j->status[i] = st;
        if (st && !j->first_bad) j->first_bad = st;
    }
    return NULL;
}

static int run_batch(job_t proto, size_t n_blocks, int threads) {
    pthread_once(&crc_once, crc_init);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n_blocks) threads = n_blocks ? (int)n_blocks : 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].lo = n_blocks * (size_t)t / (size_t)threads;
        jobs[t].hi = n_blocks * (size_t)(t + 1) / (size_t)threads;
        jobs[t].first_bad = 0;
        pthread_create(&tid[t], NULL, job_run, &jobs[t]);
    }
    int bad = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(tid[t], NULL);
        if (!bad) bad = jobs[t].first_bad;
    }
    free(tid);
    free(jobs);
    return bad;
}

int orc_compress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                       uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                       uint32_t *out_len, int32_t *status, size_t n_blocks, int hash_mode,
                       int threads) {
    job_t j = {in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, 0, 0,
               hash_mode, 1, 0};
    return run_batch(j, n_blocks, threads);
}

int orc_decompress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                         uint32_t *out_len, int32_t *status, size_t n_blocks, int threads) {
    job_t j = {in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, 0, 0, 0, 0, 0};
    return run_batch(j, n_blocks, threads);
}
