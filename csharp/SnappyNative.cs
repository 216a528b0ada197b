// SnappyNative.cs -- the reference-side binding a Snappier maintainer would add to route the
// block path through libsnappier_b200 (include/snappier_b200.h).  NOT compiled in this
// repository (no .NET SDK in the image); the identical ABI is exercised by ctypes
// (snappier_b200/_native.py) and by the C++ facade (include/snappier_b200.hpp).
//
// Drop-in points in the reference (paths relative to Snappier/):
//   Snappy.TryCompress      Snappy.cs:64-66   -> SnappyNative.TryCompress
//   Snappy.TryDecompress    Snappy.cs:174-185 -> SnappyNative.TryDecompress
//   Snappy.DecompressToMemory Snappy.cs:225-234 -> rent GetUncompressedLength bytes, TryDecompress
//   SnappyStreamCompressor.CompressBlock  Internal/SnappyStreamCompressor.cs:206 -> TryCompress
using System;
using System.Buffers;
using System.IO;
using System.Runtime.InteropServices;

namespace Snappier.Internal
{
    internal static unsafe partial class SnappyNative
    {
        private const string Lib = "snappier_b200";

        internal enum Status
        {
            Ok = 0, OutputTooSmall = 1, InvalidLength = 2, Incomplete = 3, InvalidCopyOffset = 4, DataTooLong = 5,
            UnknownChunkType = 6, CrcMismatch = 7,
            CudaError = -1, InvalidArgument = -2, NoDevice = -3, Overlap = -4, NoMemory = -5, Internal = -6,
        }

        internal enum HashMode : uint { Crc32C = 0, Mul = 1 }
        internal enum MemKind { Host = 0, Device = 1 }

        [LibraryImport(Lib)] private static partial int snp_get_max_compressed_length(int n);
        [LibraryImport(Lib)] private static partial int snp_uncompressed_length(byte* input, nuint n, uint* len);
        [LibraryImport(Lib)] private static partial int snp_compress(byte* input, nuint n, byte* output, nuint cap, nuint* written, uint hashMode);
        [LibraryImport(Lib)] private static partial int snp_decompress(byte* input, nuint n, byte* output, nuint cap, nuint* written);
        [LibraryImport(Lib)] private static partial int snp_compress_sequence(byte** segPtr, nuint* segLen, nuint nSeg, byte* output, nuint cap, nuint* written, uint hashMode);
        [LibraryImport(Lib)] private static partial int snp_decompress_sequence(byte** segPtr, nuint* segLen, nuint nSeg, byte* output, nuint cap, nuint* written);
        [LibraryImport(Lib)] private static partial int snp_create(int device, IntPtr* ctx);
        [LibraryImport(Lib)] private static partial void snp_destroy(IntPtr ctx);
        [LibraryImport(Lib)] private static partial int snp_compress_batch(IntPtr ctx, byte* inBase, ulong* inOff, uint* inLen,
            byte* outBase, ulong* outOff, uint* outCap, uint* outLen, int* status, nuint nItems, uint hashMode, int memKind, IntPtr stream);
        [LibraryImport(Lib)] private static partial int snp_decompress_batch(IntPtr ctx, byte* inBase, ulong* inOff, uint* inLen,
            byte* outBase, ulong* outOff, uint* outCap, uint* outLen, int* status, nuint nItems, int memKind, IntPtr stream);
        [LibraryImport(Lib)] private static partial IntPtr snp_last_error();
        // framing format (what SnappyStreamCompressor.CompressBlock / SnappyStreamDecompressor do per chunk, batched)
        [LibraryImport(Lib)] private static partial nuint snp_frame_max_compressed_length(nuint n);
        [LibraryImport(Lib)] private static partial int snp_frame_compress(byte* input, nuint n, byte* output, nuint cap, nuint* written, uint hashMode);
        [LibraryImport(Lib)] private static partial int snp_frame_uncompressed_length(byte* input, nuint n, ulong* len);
        [LibraryImport(Lib)] private static partial int snp_frame_decompress(byte* input, nuint n, byte* output, nuint cap, nuint* written);
        [LibraryImport(Lib)] private static partial int snp_crc32c_batch(IntPtr ctx, byte* inBase, ulong* off, uint* len, uint* crc, nuint nItems, int masked, int memKind, IntPtr stream);

        // The hash Snappier itself would use on this machine (HashTable.cs:103-123), so that the
        // native path emits the same bytes as the managed path it replaces.
        internal static HashMode PlatformHashMode =>
#if NET8_0_OR_GREATER
            (System.Runtime.Intrinsics.X86.Sse42.IsSupported || System.Runtime.Intrinsics.Arm.Crc32.IsSupported)
                ? HashMode.Crc32C : HashMode.Mul;
#else
            HashMode.Mul;
#endif

        internal static int GetMaxCompressedLength(int inputLength) => snp_get_max_compressed_length(inputLength);

        // Body of Snappy.TryCompress (Snappy.cs:55-67).
        internal static bool TryCompress(ReadOnlySpan<byte> input, Span<byte> output, out int bytesWritten)
        {
            fixed (byte* pin = input)
            fixed (byte* pout = output)
            {
                nuint written;
                var st = (Status)snp_compress(pin, (nuint)input.Length, pout, (nuint)output.Length, &written, (uint)PlatformHashMode);
                bytesWritten = (int)written;
                if (st == Status.OutputTooSmall) return false;
                ThrowFor(st, decompress: false);
                return true;
            }
        }

        // Body of Snappy.Compress(ReadOnlySequence<byte>, IBufferWriter<byte>) (Snappy.cs:82-89).  The native call cuts the
        // fragments along the segments exactly like SnappyCompressor.Compress(ReadOnlySequence<byte>, ..) (:103-143).
        internal static void Compress(ReadOnlySequence<byte> input, IBufferWriter<byte> output)
        {
            ArgumentNullException.ThrowIfNull(output);
            if (input.Length > uint.MaxValue)
                ThrowHelper.ThrowArgumentException($"{nameof(input)} is larger than the maximum size of {uint.MaxValue} bytes.", nameof(input));
            int nSeg = 0;
            foreach (ReadOnlyMemory<byte> _ in input) nSeg++;
            var handles = new MemoryHandle[nSeg];
            byte** ptr = stackalloc byte*[Math.Max(nSeg, 1)];
            nuint* len = stackalloc nuint[Math.Max(nSeg, 1)];
            try
            {
                int i = 0;
                foreach (ReadOnlyMemory<byte> seg in input)
                {
                    handles[i] = seg.Pin();
                    ptr[i] = (byte*)handles[i].Pointer;
                    len[i] = (nuint)seg.Length;
                    i++;
                }
                // worst case: every fragment carries its own 32 + n/6 + 1 bytes of slack (Helpers.cs:17-46)
                int cap = checked((int)(input.Length + input.Length / 6 + 64 * (nSeg + input.Length / 32768 + 2)));
                Span<byte> span = output.GetSpan(cap);
                fixed (byte* pout = span)
                {
                    nuint written;
                    var st = (Status)snp_compress_sequence(ptr, len, (nuint)nSeg, pout, (nuint)span.Length, &written, (uint)PlatformHashMode);
                    ThrowFor(st, decompress: false);
                    output.Advance((int)written);
                }
            }
            finally
            {
                foreach (MemoryHandle h in handles) h.Dispose();
            }
        }

        // Body of Snappy.GetUncompressedLength (Snappy.cs:142-143).
        internal static int GetUncompressedLength(ReadOnlySpan<byte> input)
        {
            fixed (byte* pin = input)
            {
                uint len;
                if (snp_uncompressed_length(pin, (nuint)input.Length, &len) != 0)
                    ThrowHelper.ThrowInvalidDataException("Invalid stream length"); // VarIntEncoding.Read.cs:20
                return (int)len;
            }
        }

        // Body of Snappy.TryDecompress (Snappy.cs:172-186).
        internal static bool TryDecompress(ReadOnlySpan<byte> input, Span<byte> output, out int bytesWritten)
        {
            fixed (byte* pin = input)
            fixed (byte* pout = output)
            {
                nuint written;
                var st = (Status)snp_decompress(pin, (nuint)input.Length, pout, (nuint)output.Length, &written);
                bytesWritten = (int)written;
                if (st == Status.OutputTooSmall) return false;   // decompressor.EndOfFile == false
                ThrowFor(st, decompress: true);
                return true;
            }
        }

        private static void ThrowFor(Status st, bool decompress)
        {
            switch (st)
            {
                case Status.Ok: return;
                case Status.InvalidLength:
                    if (decompress) ThrowHelper.ThrowInvalidOperationException("Invalid stream length"); // SnappyDecompressor.cs:53-56
                    ThrowHelper.ThrowInvalidDataException("Invalid stream length");
                    return;
                case Status.Incomplete: ThrowHelper.ThrowInvalidDataExceptionIncompleteSnappyBlock(); return; // ThrowHelper.cs:27-28
                case Status.InvalidCopyOffset: ThrowHelper.ThrowInvalidDataException("Invalid copy offset"); return; // SnappyDecompressor.cs:600
                case Status.DataTooLong: ThrowHelper.ThrowInvalidDataException("Data too long"); return;             // SnappyDecompressor.cs:572,605
                case Status.Overlap: ThrowHelper.ThrowInvalidOperationException("Input and output spans must not overlap."); return; // SnappyCompressor.cs:29
                case Status.UnknownChunkType: ThrowHelper.ThrowInvalidDataException("Unknown chunk type"); return;   // SnappyStreamDecompressor.cs:182-185
                case Status.CrcMismatch: ThrowHelper.ThrowInvalidDataException("Chunk CRC mismatch."); return;       // SnappyStreamDecompressor.cs:127-131
                case Status.NoMemory: throw new OutOfMemoryException("snappier_b200: host allocation failed");
                default:
                    throw new InvalidOperationException(
                        $"snappier_b200: {st}: {Marshal.PtrToStringUTF8(snp_last_error())} (there is no CPU fallback)");
            }
        }
    }
}
