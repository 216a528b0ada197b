/*
 * snappier_b200.h -- C ABI of the B200-native Snappy block engine.
 *
 * This is the drop-in boundary for Snappier's block path.  Snappier itself has
 * no FFI seam (it is 100 % managed C#); the seam is the pair of internal calls
 * that `Snappy.cs` makes into `SnappyCompressor` / `SnappyDecompressor`.  Each
 * entry point below names the reference interface it replaces (paths relative
 * to /root/reference/Snappier/).  INTEGRATION.md shows the P/Invoke stub.
 *
 * Plain C: pointers, sizes, fixed-width integers.  No torch / C++ types.
 * All functions are thread-safe; per-thread CUDA resources are created lazily.
 * There is NO CPU fallback: without a usable CUDA device every compute entry
 * point returns SNP_E_NO_DEVICE.
 */
#ifndef SNAPPIER_B200_H
#define SNAPPIER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNP_ABI_VERSION 1

/* ---- per-block status (>= 0).  The C# shim maps them to the reference's
 *      exceptions exactly as listed. ------------------------------------- */
enum snp_status {
    SNP_OK = 0,
    /* Try* return false; Compress/Decompress throw
     * ArgumentException("Output buffer is too small.") -- ThrowHelper.cs:18-19 */
    SNP_OUTPUT_TOO_SMALL = 1,
    /* "Invalid stream length": InvalidDataException from GetUncompressedLength
     * (VarIntEncoding.Read.cs:20), InvalidOperationException from Decompress
     * (SnappyDecompressor.cs:53-56) */
    SNP_INVALID_LENGTH = 2,
    /* InvalidDataException("Incomplete Snappy block.") -- ThrowHelper.cs:27-28 */
    SNP_INCOMPLETE = 3,
    /* InvalidDataException("Invalid copy offset") -- SnappyDecompressor.cs:600 */
    SNP_INVALID_COPY_OFFSET = 4,
    /* InvalidDataException("Data too long") -- SnappyDecompressor.cs:572,605 */
    SNP_DATA_TOO_LONG = 5,
    /* framing format only: InvalidDataException("Unknown chunk type ..") -- SnappyStreamDecompressor.cs:182-185 */
    SNP_UNKNOWN_CHUNK_TYPE = 6,
    /* framing format only: InvalidDataException("Chunk CRC mismatch.") -- SnappyStreamDecompressor.cs:127-131,169-173 */
    SNP_CRC_MISMATCH = 7
};

/* ---- call-level errors (< 0) ------------------------------------------ */
enum snp_error {
    SNP_E_CUDA = -1,        /* a CUDA call failed; see snp_last_error() */
    SNP_E_INVALID_ARG = -2, /* null pointer / bad enum / overlap */
    SNP_E_NO_DEVICE = -3,   /* no CUDA device or kernel image not loadable */
    SNP_E_OVERLAP = -4,     /* input and output overlap -> InvalidOperationException
                               ("Input and output spans must not overlap.",
                               SnappyCompressor.cs:27-30) */
    SNP_E_NOMEM = -5,       /* a host allocation failed -> OutOfMemoryException */
    SNP_E_INTERNAL = -6     /* an unexpected C++ exception was caught at the ABI boundary; see snp_last_error() */
};

/* Hash used by the compressor's match finder (HashTable.cs:91-126).  Snappier
 * picks it per platform, so the compressed BYTES differ per platform:
 *   CRC32C : x64 with SSE4.2 / ARM64 with CRC, .NET 8+       (HashTable.cs:109-117)
 *   MUL    : netstandard2.0 / net472 / intrinsics disabled   (HashTable.cs:120-123) */
enum snp_hash_mode { SNP_HASH_CRC32C = 0, SNP_HASH_MUL = 1 };

/* Where the buffers of a batched call live. */
enum snp_mem_kind {
    SNP_MEM_HOST = 0,   /* every pointer is host memory; call is synchronous   */
    SNP_MEM_DEVICE = 1  /* every pointer is device memory of the context's GPU;
                           work is enqueued on `stream` and the call returns    */
};

#define SNP_BLOCK_SIZE 65536u /* Constants.cs:25-26 */

int snp_abi_version(void);
const char *snp_status_string(int status);
/* Thread-local text of the last SNP_E_CUDA on this thread ("" if none). */
const char *snp_last_error(void);

/* ---- sizing ------------------------------------------------------------ */
/* Helpers.MaxCompressedLength (Helpers.cs:17-46): 32 + n + n/6 + 1. */
int32_t snp_max_compressed_length(int32_t n);
/* Snappy.GetMaxCompressedLength (Snappy.cs:20-24): the above + 5. */
int32_t snp_get_max_compressed_length(int32_t n);
/* Snappy.GetUncompressedLength (Snappy.cs:142-143).  Pure host varint read. */
int snp_uncompressed_length(const uint8_t *in, size_t n, uint32_t *len);

/* ---- contexts ---------------------------------------------------------- */
/* One context = one GPU + one private stream + reusable staging buffers.  A
 * context may be used by one thread at a time (calls on it are serialised by
 * an internal mutex).  The single-call API below shares ONE lazily created
 * process-wide context per device (the current device of the calling thread):
 * device scratch does not grow with the number of calling threads.  Scratch
 * follows the largest call so far: ~2.2 x the 64 MiB pipeline chunk per slot
 * in use, plus 64 KiB of hash table per concurrently compressed 64 KiB block
 * (512 KiB for a one-block call).  No C++ exception crosses this ABI: host
 * allocation failures return SNP_E_NOMEM. */
typedef struct snp_ctx snp_ctx;
int snp_create(int device, snp_ctx **ctx);
void snp_destroy(snp_ctx *ctx);
int snp_ctx_device(const snp_ctx *ctx);
/* Number of kernel launches this context has issued (bench: gpu_launches). */
uint64_t snp_ctx_launch_count(const snp_ctx *ctx);

/* ---- single-call API: host buffers, synchronous ------------------------ */
/* Replaces SnappyCompressor.TryCompress(ReadOnlySpan<byte>, Span<byte>, out int)
 * (SnappyCompressor.cs:24-83) as called by Snappy.TryCompress (Snappy.cs:55-67).
 * Any input length < 2^32; inputs above 64 KiB are split into independent
 * 64 KiB fragments exactly like the reference (SnappyCompressor.cs:40-80).
 * SNP_OUTPUT_TOO_SMALL -> *written = 0 (SnappyCompressor.cs:63-68). */
int snp_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written,
                 uint32_t hash_mode);

/* Replaces SnappyCompressor.Compress(ReadOnlySequence<byte>, IBufferWriter<byte>) (SnappyCompressor.cs:85-144)
 * as called by Snappy.Compress(ReadOnlySequence<byte>, IBufferWriter<byte>) (Snappy.cs:82-89): the input is a
 * list of host segments.  The reference cuts fragments along the segmentation -- the next fragment is the
 * first segment's part of the next <= 64 KiB when that part is the whole fragment or >= 32 KiB, otherwise
 * the whole fragment (SnappyCompressor.cs:103-143) -- so the same bytes in different segments compress to
 * different bytes; this call reproduces that partition.  One contiguous segment == snp_compress. */
int snp_compress_sequence(const uint8_t *const *seg_ptr, const size_t *seg_len, size_t n_seg, uint8_t *out,
                          size_t cap, size_t *written, uint32_t hash_mode);

/* Replaces the one-shot use of SnappyDecompressor in Snappy.TryDecompress
 * (Snappy.cs:172-186): Decompress -> AllDataDecompressed -> Read -> EndOfFile.
 * Data errors take precedence over SNP_OUTPUT_TOO_SMALL, and on TOO_SMALL the
 * first `cap` bytes are still written (SnappyDecompressor.Read, :613-629). */
int snp_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);

/* Snappy.Decompress(ReadOnlySequence<byte>, IBufferWriter<byte>) / DecompressToMemory(ReadOnlySequence<byte>)
 * (Snappy.cs:194-212,246-261): the segments are one block split at arbitrary points; same result and statuses
 * as snp_decompress on their concatenation. */
int snp_decompress_sequence(const uint8_t *const *seg_ptr, const size_t *seg_len, size_t n_seg, uint8_t *out,
                            size_t cap, size_t *written);

/* ---- batched API: the throughput path (an extension; the reference has no
 *      batch call -- each item is one independent Snappy.Compress/Decompress) --
 *
 * Item i reads  in_base  + in_off[i]  .. + in_len[i]
 *        writes out_base + out_off[i] .. + out_cap[i]   (capacity)
 * and reports out_len[i] (bytes produced; 0 unless status[i] == SNP_OK) and
 * status[i] (enum snp_status).  A bad item never affects its neighbours.
 * Item regions must not overlap each other or the input.
 * CONTENT OF AN OUTPUT REGION BEYOND out_len[i] IS UNSPECIFIED: a batch call with
 * two or more items may write anywhere inside [out_off[i], out_off[i] + out_cap[i])
 * -- partial output of an item that is then rejected, or device scratch behind
 * the produced bytes (host mode copies whole capacity spans back before the
 * statuses are known; compress slots are copied as wide as the longest item of
 * their group).  Callers that must not see such bytes clear the regions or use
 * the single-call API, which leaves the output untouched on failure like the
 * reference (it decodes into a private buffer).
 *
 * compress: in_len[i] <= SNP_BLOCK_SIZE (one fragment per item, one warp per
 *   item); give each item snp_get_max_compressed_length(in_len[i]) of capacity
 *   to make SNP_OUTPUT_TOO_SMALL impossible.  Larger inputs: snp_compress().
 * decompress: each item is a complete block with its own varint header.
 *
 * ctx == NULL uses the process-wide default context of the calling thread's current device.  `stream` is a
 * cudaStream_t used for SNP_MEM_DEVICE only (NULL = the legacy default stream,
 * as everywhere in CUDA; the caller orders the work against its own).  Return: SNP_OK once the work is done (HOST) / enqueued (DEVICE),
 * or a negative snp_error.  Per-item failures are NOT a call failure. */
int snp_compress_batch(snp_ctx *ctx, const uint8_t *in_base, const uint64_t *in_off,
                       const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off,
                       const uint32_t *out_cap, uint32_t *out_len, int32_t *status, size_t n_items,
                       uint32_t hash_mode, int mem_kind, void *stream);

int snp_decompress_batch(snp_ctx *ctx, const uint8_t *in_base, const uint64_t *in_off,
                         const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off,
                         const uint32_t *out_cap, uint32_t *out_len, int32_t *status,
                         size_t n_items, int mem_kind, void *stream);

/* Batched Snappy.GetUncompressedLength: ulen[i] / status[i] per item. */
int snp_uncompressed_length_batch(snp_ctx *ctx, const uint8_t *in_base, const uint64_t *in_off,
                                  const uint32_t *in_len, uint32_t *ulen, int32_t *status,
                                  size_t n_items, int mem_kind, void *stream);

/* ---- framing format (the caller either side of the block path) ----
 * What `new SnappyStream(s, CompressionMode.Compress)` emits for one Write of the whole buffer
 * followed by Dispose: stream identifier, then one chunk per 64 KiB
 * [type 0x00 | 0x01][len24][masked CRC32C of the raw chunk][snappy block | raw bytes]
 * (SnappyStreamCompressor.cs:15-18,166-261).  Host buffers, synchronous; internally the stream flows through the
 * host-mode pipeline in pieces of chunks (H2D | kernels | D2H overlap), as the reference processes it chunk by chunk. */

/* 10 + n + 8 * ceil(n / 65536): every chunk falls back to raw when compression does not shrink it. */
size_t snp_frame_max_compressed_length(size_t n);
int snp_frame_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written,
                       uint32_t hash_mode);
/* Sum of the chunks' uncompressed lengths (host-only header walk).  SNP_INCOMPLETE for a
 * truncated stream, SNP_UNKNOWN_CHUNK_TYPE for reserved unskippable chunks 0x02..0x7f. */
int snp_frame_uncompressed_length(const uint8_t *in, size_t n, uint64_t *len);
/* One-shot equivalent of reading a SnappyStream to the end (SnappyStreamDecompressor.cs:38-208):
 * skippable chunks (>= 0x80, including the stream identifier, whose content the reference does
 * not validate) are skipped, every data chunk's CRC is verified.  On error *written = 0 and the
 * status is that of the first bad chunk in stream order. */
int snp_frame_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);

/* Packs a batch: copies item i's len[i] bytes from src_base + src_off[i] to dst_base + dst_off[i], where dst_off is the
 * exclusive prefix sum of len (written by the call) and *total their sum -- the batched form of what
 * Snappy.CompressToMemory returns per call (Snappy.cs:84-100: exactly the compressed bytes, no slack), and the gather(v)
 * side of block-range sharding (SnappyCompressor.cs:40-44: blocks are independent).  DEVICE pointers only (src/dst/len/
 * dst_off/total); enqueued on `stream`, asynchronous.  dst_base == NULL: offsets and total only (size query).
 * dst must not overlap src.  Calls on one context that run concurrently on different streams must not overlap in time. */
int snp_pack_batch(snp_ctx *ctx, const uint8_t *src_base, const uint64_t *src_off, const uint32_t *len, size_t n_items,
                   uint8_t *dst_base, uint64_t *dst_off, uint64_t *total, void *stream);

/* Batched SnappyCompressor.FindMatchLength (SnappyCompressor.cs:562-688): matched[i] = length of the common prefix of
 * base[s1[i]..] and base[s2[i]..s2_limit[i]) -- the compress kernels' own device function behind an entry point, so that the
 * reference's known-answer vectors (Snappier.Tests/Internal/SnappyCompressorTests.cs:10-81) run on the GPU directly.
 * DEVICE pointers only; enqueued on `stream`. */
int snp_find_match_length_batch(snp_ctx *ctx, const uint8_t *base, const uint32_t *s1, const uint32_t *s2,
                                const uint32_t *s2_limit, uint32_t *matched, size_t n_items, void *stream);

/* DIAGNOSTICS (not part of the reference's surface): measures the GPU's random-access read ceiling that bounds designs
 * which turn back-references or hash-table probes into independent DRAM accesses.  sm_count x ctas_per_sm
 * CTAs of 256 threads each issue reads_per_thread 16-byte loads at pseudo-random aligned offsets of [base, base +
 * span_bytes) (DEVICE memory, 16-byte aligned).  Time it with events on `stream`; tools/random_access_probe.py does. */
int snp_diag_random_reads(snp_ctx *ctx, const uint8_t *base, size_t span_bytes, uint32_t ctas_per_sm,
                          uint32_t reads_per_thread, uint32_t *sink, void *stream);

/* Batched Crc32CAlgorithm.Compute (+ ApplyMask when masked != 0), Crc32CAlgorithm.cs:41-44,157-158. */
int snp_crc32c_batch(snp_ctx *ctx, const uint8_t *base, const uint64_t *off, const uint32_t *len,
                     uint32_t *crc, size_t n_items, int masked, int mem_kind, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SNAPPIER_B200_H */
