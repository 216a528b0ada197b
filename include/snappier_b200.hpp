// snappier_b200.hpp -- C++ host-side mirror of Snappier's static `Snappy` facade
// (/root/reference/Snappier/Snappy.cs) above the C ABI of snappier_b200.h.
//
// The reference is compiled (managed C#) code and no .NET toolchain exists in this
// image, so the host side is written in C++: same member names, argument meaning
// and error behaviour, with .NET exceptions mapped to same-named C++ exceptions.
// csharp/SnappyNative.cs is the P/Invoke equivalent for a machine that has .NET.
// Header-only; link with -lsnappier_b200.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "snappier_b200.h"

namespace snappier {

struct ArgumentException : std::invalid_argument {
    using std::invalid_argument::invalid_argument;
};
struct InvalidDataException : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct InvalidOperationException : std::logic_error {
    using std::logic_error::logic_error;
};
struct NativeLibraryException : std::runtime_error {  // no GPU / CUDA failure: there is no CPU fallback
    using std::runtime_error::runtime_error;
};

// ReadOnlySpan<byte> / Span<byte>
struct ReadOnlySpan {
    const uint8_t *data;
    size_t size;
};
struct Span {
    uint8_t *data;
    size_t size;
};

// IMemoryOwner<byte> returned by CompressToMemory / DecompressToMemory (ByteArrayPoolMemoryOwner.cs)
class MemoryOwner {
public:
    MemoryOwner() = default;
    MemoryOwner(std::vector<uint8_t> &&buf, size_t len) : buf_(std::move(buf)), len_(len) {}
    Span Memory() { return {buf_.data(), len_}; }
    size_t Length() const { return len_; }

private:
    std::vector<uint8_t> buf_;
    size_t len_ = 0;
};

class Snappy {
public:
    // Hash variant of the match finder; Snappier on x64/.NET 8+ uses CRC32C (HashTable.cs:109-117).
    static uint32_t &HashMode() {
        static uint32_t mode = SNP_HASH_CRC32C;
        return mode;
    }

    // Snappy.cs:20-24
    static int GetMaxCompressedLength(int inputLength) { return snp_get_max_compressed_length(inputLength); }

    // Snappy.cs:55-67
    static bool TryCompress(ReadOnlySpan input, Span output, int &bytesWritten) {
        size_t w = 0;
        int st = snp_compress(input.data, input.size, output.data, output.size, &w, HashMode());
        bytesWritten = (int)w;
        if (st == SNP_OUTPUT_TOO_SMALL) return false;
        Throw(st, /*decompress=*/false);
        return true;
    }

    // Snappy.cs:37-45
    static int Compress(ReadOnlySpan input, Span output) {
        int n = 0;
        if (!TryCompress(input, output, n)) throw ArgumentException("Output buffer is too small.");
        return n;
    }

    // Snappy.cs:82-89: Compress(ReadOnlySequence<byte>, IBufferWriter<byte>).  The sequence is its list of segments,
    // the writer a byte vector that is appended to.  Fragment boundaries follow the segments exactly like
    // SnappyCompressor.cs:103-143, so the bytes depend on the segmentation.
    static void Compress(const std::vector<ReadOnlySpan> &input, std::vector<uint8_t> &output) {
        std::vector<const uint8_t *> ptr;
        std::vector<size_t> len;
        size_t total = 0;
        for (const ReadOnlySpan &s : input) ptr.push_back(s.data), len.push_back(s.size), total += s.size;
        if (total > 0xffffffffull) throw ArgumentException("input is larger than the maximum size of 4294967295 bytes.");
        const size_t at = output.size();
        const size_t cap = total + total / 6 + 64 * (input.size() + total / 32768 + 2);
        output.resize(at + cap);
        size_t w = 0;
        int st = snp_compress_sequence(ptr.data(), len.data(), ptr.size(), output.data() + at, cap, &w, HashMode());
        output.resize(at + (st == SNP_OK ? w : 0));
        Throw(st, /*decompress=*/false);
    }

    // Snappy.cs:99-113
    static MemoryOwner CompressToMemory(ReadOnlySpan input) {
        std::vector<uint8_t> buf((size_t)GetMaxCompressedLength((int)input.size));
        int n = 0;
        if (!TryCompress(input, {buf.data(), buf.size()}, n)) throw InvalidOperationException("unreachable");
        return MemoryOwner(std::move(buf), (size_t)n);
    }

    // Snappy.cs:123-132
    static std::vector<uint8_t> CompressToArray(ReadOnlySpan input) {
        MemoryOwner m = CompressToMemory(input);
        Span s = m.Memory();
        return std::vector<uint8_t>(s.data, s.data + s.size);
    }

    // Snappy.cs:142-143
    static int GetUncompressedLength(ReadOnlySpan input) {
        uint32_t len = 0;
        if (snp_uncompressed_length(input.data, input.size, &len) != SNP_OK)
            throw InvalidDataException("Invalid stream length");  // VarIntEncoding.Read.cs:20
        return (int)len;
    }

    // Snappy.cs:172-186
    static bool TryDecompress(ReadOnlySpan input, Span output, int &bytesWritten) {
        size_t w = 0;
        int st = snp_decompress(input.data, input.size, output.data, output.size, &w);
        bytesWritten = (int)w;
        if (st == SNP_OUTPUT_TOO_SMALL) return false;
        Throw(st, /*decompress=*/true);
        return true;
    }

    // Snappy.cs:153-162
    static int Decompress(ReadOnlySpan input, Span output) {
        int n = 0;
        if (!TryDecompress(input, output, n)) throw ArgumentException("Output buffer is too small.");
        return n;
    }

    // Snappy.cs:223-235
    static MemoryOwner DecompressToMemory(ReadOnlySpan input) {
        uint32_t len = 0;
        snp_uncompressed_length(input.data, input.size, &len);  // errors surface from Decompress below
        std::vector<uint8_t> buf(len ? len : 1);
        int n = 0;
        TryDecompress(input, {buf.data(), len}, n);
        return MemoryOwner(std::move(buf), (size_t)n);
    }

    // Snappy.cs:194-212 / 246-261: Decompress(ReadOnlySequence<byte>, IBufferWriter<byte>) and
    // DecompressToMemory(ReadOnlySequence<byte>): one block split into segments, appended to the writer.
    static void Decompress(const std::vector<ReadOnlySpan> &input, std::vector<uint8_t> &output) {
        std::vector<const uint8_t *> ptr;
        std::vector<size_t> len;
        uint8_t head[5];
        size_t nh = 0;
        for (const ReadOnlySpan &s : input) {
            ptr.push_back(s.data), len.push_back(s.size);
            for (size_t i = 0; i < s.size && nh < 5; i++) head[nh++] = s.data[i];
        }
        uint32_t U = 0;
        snp_uncompressed_length(head, nh, &U);  // errors surface from the decode below
        const size_t at = output.size();
        output.resize(at + (U ? U : 1));
        size_t w = 0;
        int st = snp_decompress_sequence(ptr.data(), len.data(), ptr.size(), output.data() + at, U, &w);
        output.resize(at + (st == SNP_OK ? w : 0));
        Throw(st, /*decompress=*/true);
    }

    // Snappy.cs:273-282
    static std::vector<uint8_t> DecompressToArray(ReadOnlySpan input) {
        int len = GetUncompressedLength(input);
        std::vector<uint8_t> out((size_t)len);
        Decompress(input, {out.data(), out.size()});
        return out;
    }

    // ---- framing format, one-shot (SnappyStream.cs + SnappyStreamCompressor/Decompressor.cs) ----
    // What `new SnappyStream(s, CompressionMode.Compress)` emits for Write(input) + Dispose.
    static std::vector<uint8_t> FrameCompress(ReadOnlySpan input) {
        std::vector<uint8_t> out(snp_frame_max_compressed_length(input.size));
        size_t w = 0;
        Throw(snp_frame_compress(input.data, input.size, out.data(), out.size(), &w, HashMode()), false);
        out.resize(w);
        return out;
    }
    // Reading `new SnappyStream(s, CompressionMode.Decompress)` to the end.
    static std::vector<uint8_t> FrameDecompress(ReadOnlySpan input) {
        uint64_t len = 0;
        Throw(snp_frame_uncompressed_length(input.data, input.size, &len), false);
        std::vector<uint8_t> out(len ? len : 1);
        size_t w = 0;
        Throw(snp_frame_decompress(input.data, input.size, out.data(), len, &w), false);
        out.resize(w);
        return out;
    }

private:
    static void Throw(int st, bool decompress) {
        switch (st) {
            case SNP_OK: return;
            case SNP_INVALID_LENGTH:
                if (decompress) throw InvalidOperationException("Invalid stream length");  // SnappyDecompressor.cs:53-56
                throw InvalidDataException("Invalid stream length");
            case SNP_INCOMPLETE: throw InvalidDataException("Incomplete Snappy block.");  // ThrowHelper.cs:27-28
            case SNP_INVALID_COPY_OFFSET: throw InvalidDataException("Invalid copy offset");  // SnappyDecompressor.cs:600
            case SNP_DATA_TOO_LONG: throw InvalidDataException("Data too long");  // SnappyDecompressor.cs:572,605
            case SNP_UNKNOWN_CHUNK_TYPE: throw InvalidDataException("Unknown chunk type");  // SnappyStreamDecompressor.cs:182-185
            case SNP_CRC_MISMATCH: throw InvalidDataException("Chunk CRC mismatch.");      // SnappyStreamDecompressor.cs:127-131
            case SNP_E_OVERLAP:
                throw InvalidOperationException("Input and output spans must not overlap.");  // SnappyCompressor.cs:29
            case SNP_E_INVALID_ARG: throw ArgumentException("invalid argument");
            default:
                throw NativeLibraryException(std::string(snp_status_string(st)) + ": " + snp_last_error());
        }
    }
};

}  // namespace snappier
