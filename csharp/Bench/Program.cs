// Times the real reference (Snappier's managed C# path) on the same kind of batch bench.py uses, when a
// .NET SDK is available (it is not in this image; bench.py probes `dotnet --version` and says so).
//   dotnet run -c Release --project csharp/Bench -- <file with raw 64 KiB blocks> [threads]
// Prints one JSON line: uncompressed GB/s for Snappy.Compress and Snappy.Decompress over all blocks,
// Parallel.For over `threads` cores, plus a SHA-256 of the concatenated compressed bytes so that the
// CRC32C-hash compress parity (unpinned by the reference's fixtures, DESIGN.md section 3) can be
// confirmed against `snappier_b200` on the same input.
using System;
using System.Diagnostics;
using System.IO;
using System.Security.Cryptography;
using System.Threading.Tasks;
using Snappier;

const int Block = 65536;
byte[] raw = File.ReadAllBytes(args[0]);
int threads = args.Length > 1 ? int.Parse(args[1]) : Environment.ProcessorCount;
int n = raw.Length / Block;
var comp = new byte[n][];
var opt = new ParallelOptions { MaxDegreeOfParallelism = threads };

var sw = Stopwatch.StartNew();
Parallel.For(0, n, opt, i => comp[i] = Snappy.CompressToArray(raw.AsSpan(i * Block, Block)));
double tc = sw.Elapsed.TotalSeconds;

var back = new byte[raw.Length];
sw.Restart();
Parallel.For(0, n, opt, i => Snappy.Decompress(comp[i], back.AsSpan(i * Block, Block)));
double td = sw.Elapsed.TotalSeconds;

using var sha = SHA256.Create();
foreach (var c in comp) sha.TransformBlock(c, 0, c.Length, null, 0);
sha.TransformFinalBlock(Array.Empty<byte>(), 0, 0);
bool ok = raw.AsSpan(0, n * Block).SequenceEqual(back.AsSpan(0, n * Block));
Console.WriteLine($"{{\"impl\":\"snappier-dotnet\",\"blocks\":{n},\"threads\":{threads},\"compress_GBps\":{n * (double)Block / tc / 1e9:F3}," +
                  $"\"decompress_GBps\":{n * (double)Block / td / 1e9:F3},\"round_trip_ok\":{ok.ToString().ToLower()}," +
                  $"\"compressed_sha256\":\"{Convert.ToHexString(sha.Hash!).ToLower()}\"}}");
