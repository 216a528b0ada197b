/*
 * snappy_oracle.h -- CPU restatement of Snappier's Snappy *block* path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (snappier_b200/,
 * include/, csrc/) may include, link or call this.  Allowed callers: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * The reference (brantburnett/Snappier, 100 % managed C#) cannot run in this
 * image (no dotnet/mono), so this is a plain-C restatement of its algorithm.
 * Parity status:
 *   - decompress, varint, CRC32C, MaxCompressedLength, FindMatchLength and
 *     compress with hash_mode = ORC_HASH_MUL are PINNED by the reference's own
 *     fixtures and KATs (tests/golden/, see tests/test_oracle_golden.py);
 *   - compress with hash_mode = ORC_HASH_CRC32C (what Snappier uses on x64 /
 *     .NET 8+) is "parity unpinned": no reference fixture holds those bytes.
 *     It differs from the pinned variant only in the one hash line.
 *
 * Reference paths are relative to /root/reference/.
 */
#ifndef SNAPPY_ORACLE_H
#define SNAPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes -- numerically identical to include/snappier_b200.h */
enum {
    ORC_OK = 0,
    ORC_OUTPUT_TOO_SMALL = 1,    /* Snappy.cs:41,158  "Output buffer is too small." */
    ORC_INVALID_LENGTH = 2,      /* VarIntEncoding.Read.cs:20 / SnappyDecompressor.cs:55 */
    ORC_INCOMPLETE = 3,          /* ThrowHelper.cs:27-28 "Incomplete Snappy block." */
    ORC_INVALID_COPY_OFFSET = 4, /* SnappyDecompressor.cs:600 */
    ORC_DATA_TOO_LONG = 5        /* SnappyDecompressor.cs:572,605 */
};

enum { ORC_HASH_CRC32C = 0, ORC_HASH_MUL = 1 };

#define ORC_BLOCK_SIZE 65536 /* Constants.cs:25-26 */

/* Helpers.cs:17-46 (without the +5 varint pad of Snappy.cs:20-24). */
int32_t orc_max_compressed_length(int32_t n);
/* Snappy.cs:20-24. */
int32_t orc_get_max_compressed_length(int32_t n);

/* VarIntEncoding.Write.cs:5-79.  Returns bytes written, 0 if cap too small. */
int orc_varint_write(uint8_t *out, size_t cap, uint32_t v);
/* VarIntEncoding.Read.cs:38-79 (slow path = the semantics).  Returns
 * ORC_OK / ORC_INCOMPLETE (need more data) / ORC_INVALID_LENGTH. */
int orc_varint_read(const uint8_t *in, size_t n, uint32_t *v, int *consumed);

/* HashTable.cs:57-71. */
int orc_table_size(int fragment_len);
/* HashTable.cs:91-126: byte offset into the u16 table (hash & mask). */
uint32_t orc_table_hash(uint32_t bytes, uint32_t mask, int hash_mode);
/* Same value through the form the hot loop uses (hardware crc32 when built
 * with SSE4.2); orc_hash_uses_hw_crc() says which. */
uint32_t orc_table_hash_fast(uint32_t bytes, uint32_t mask, int hash_mode);
int orc_hash_uses_hw_crc(void);

/* SnappyCompressor.cs:562-688.  Longest common prefix of s1.. and [s2,s2_limit). */
int orc_find_match_length(const uint8_t *s1, const uint8_t *s2, const uint8_t *s2_limit);

/* SnappyCompressor.cs:174-415.  `out` needs orc_max_compressed_length(n) bytes.
 * Returns bytes written. */
size_t orc_compress_fragment(const uint8_t *in, size_t n, uint8_t *out, int hash_mode);

/* SnappyCompressor.cs:24-83 behind Snappy.cs:55-67.  Status OK or
 * OUTPUT_TOO_SMALL (then *written = 0). */
int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written,
                 int hash_mode);

/* Snappy.cs:142-143 -> SnappyDecompressor.cs:181-182. */
int orc_uncompressed_length(const uint8_t *in, size_t n, uint32_t *len);

/* Snappy.cs:172-186 one-shot semantics over SnappyDecompressor.cs:43-92,184-347,
 * 556-611.  Decodes into `out` (cap bytes).  Strict: production beyond the
 * declared length is DATA_TOO_LONG (SURVEY App. C Q1 -- the reference bounds
 * by a pow2-rounded pool buffer instead). */
int orc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);

/* Crc32CAlgorithm.cs:41-158. */
uint32_t orc_crc32c(uint32_t crc, const uint8_t *p, size_t n);
uint32_t orc_crc32c_mask(uint32_t crc);

/* Multi-threaded batch drivers (CPU baseline for bench.py).  Blocks are
 * statically partitioned over `threads` pthreads.  Return 0 or the first
 * non-OK status seen. */
int orc_compress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                       uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                       uint32_t *out_len, int32_t *status, size_t n_blocks, int hash_mode,
                       int threads);
int orc_decompress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                         uint32_t *out_len, int32_t *status, size_t n_blocks, int threads);

#ifdef __cplusplus
}
#endif
#endif
