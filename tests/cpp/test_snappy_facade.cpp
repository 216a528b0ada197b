// C++ test of include/snappier_b200.hpp, written to read like the reference's own
// block tests (/root/reference/Snappier.Tests/SnappyTests.cs).  Built and run by
// tests/test_gpu_cpp_facade.py on the GPU box:  ./test_snappy_facade <fixture dir>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "snappier_b200.hpp"

using namespace snappier;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)
template <class E, class F> static bool throws(F f) {
    try { f(); } catch (const E &) { return true; } catch (...) { return false; }
    return false;
}
static std::vector<uint8_t> read_file(const std::string &p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), {});
}
static ReadOnlySpan ro(const std::vector<uint8_t> &v) { return {v.data(), v.size()}; }

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    // SnappyTests.cs:20-39 CompressAndDecompressFile (one corpus file is passed as raw bytes)
    std::vector<uint8_t> input = read_file(dir + "/input.bin");
    std::vector<uint8_t> expect = read_file(dir + "/input.snappy");  // oracle bytes, CRC32C hash
    CHECK(!input.empty());
    {
        std::vector<uint8_t> compressed((size_t)Snappy::GetMaxCompressedLength((int)input.size()));
        int n = Snappy::Compress(ro(input), {compressed.data(), compressed.size()});
        CHECK((size_t)n == expect.size() && std::memcmp(compressed.data(), expect.data(), (size_t)n) == 0);
        compressed.resize((size_t)n);
        std::vector<uint8_t> out(input.size());
        int m = Snappy::Decompress(ro(compressed), {out.data(), out.size()});
        CHECK((size_t)m == input.size() && out == input);
        CHECK(Snappy::GetUncompressedLength(ro(compressed)) == (int)input.size());
        CHECK(Snappy::DecompressToArray(ro(compressed)) == input);
        MemoryOwner mo = Snappy::DecompressToMemory(ro(compressed));
        CHECK(mo.Length() == input.size());
        // SnappyTests.cs:80-118 insufficient output
        std::vector<uint8_t> small(1024);
        int bw = -1;
        CHECK(!Snappy::TryCompress(ro(input), {small.data(), small.size()}, bw) && bw == 0);
        CHECK(throws<ArgumentException>([&] { Snappy::Compress(ro(input), {small.data(), small.size()}); }));
        // SnappyTests.cs:212-242
        CHECK(throws<ArgumentException>([&] { Snappy::Decompress(ro(compressed), {small.data(), 100}); }));
        CHECK(!Snappy::TryDecompress(ro(compressed), {small.data(), 100}, bw) && bw == 100);
        // SnappyTests.cs:244-264 simple corruption
        std::vector<uint8_t> bad = Snappy::CompressToArray({(const uint8_t *)"making sure we don't crash with corrupted input", 47});
        bad[1]--; bad[3]++;
        CHECK(throws<InvalidDataException>([&] { Snappy::DecompressToArray(ro(bad)); }));
    }
    // SnappyTests.cs:122-174: the ReadOnlySequence / IBufferWriter overloads, input split into 16/32/64 KiB and
    // ragged segments.  One segment == the span overload; any segmentation round-trips; a split block decodes.
    for (size_t seg : {size_t(16384), size_t(32768), size_t(65536), size_t(1000), size_t(40000)}) {
        std::vector<ReadOnlySpan> seq;
        for (size_t o = 0; o < input.size(); o += seg) seq.push_back({input.data() + o, std::min(seg, input.size() - o)});
        std::vector<uint8_t> c;
        Snappy::Compress(seq, c);
        CHECK(Snappy::DecompressToArray(ro(c)) == input);
        std::vector<ReadOnlySpan> cseq;
        for (size_t o = 0; o < c.size(); o += 1024) cseq.push_back({c.data() + o, std::min(size_t(1024), c.size() - o)});
        std::vector<uint8_t> back;
        Snappy::Decompress(cseq, back);
        CHECK(back == input);
        if (seg == 65536) CHECK(c == expect);  // 64 KiB segments give the span overload's fragments
    }
    {
        std::vector<ReadOnlySpan> one{ro(input)};
        std::vector<uint8_t> c;
        Snappy::Compress(one, c);
        CHECK(c == expect);
    }
    // SnappyTests.cs:178-202 edge strings
    for (std::string s : {std::string(""), std::string("a"), std::string("ab"), std::string("abc"),
                          "aaaaaaa" + std::string(16, 'b') + "aaaaaabc", "aaaaaaa" + std::string(65536, 'b') + "aaaaaabc"}) {
        std::vector<uint8_t> in(s.begin(), s.end());
        std::vector<uint8_t> c = Snappy::CompressToArray(ro(in));
        CHECK(Snappy::DecompressToArray(ro(c)) == in);
    }
    // SnappyTests.cs:204-210 overlap
    std::vector<uint8_t> buf(1024);
    CHECK(throws<InvalidOperationException>([&] { Snappy::Compress(ro(buf), {buf.data() + 1023, 1}); }));
    // SnappyTests.cs:287-331 bad data fixtures
    for (int i = 1; i <= 3; i++) {
        std::vector<uint8_t> b = read_file(dir + "/baddata" + std::to_string(i) + ".snappy");
        CHECK(!b.empty());
        CHECK(throws<InvalidDataException>([&] { Snappy::DecompressToArray(ro(b)); }));
    }
    CHECK(throws<InvalidDataException>([&] { Snappy::GetUncompressedLength({nullptr, 0}); }));
    // SnappyStreamTests.cs:8-262 shape: framed round trip, golden framed stream, CRC corruption
    {
        std::vector<uint8_t> framed = Snappy::FrameCompress(ro(input));
        CHECK(framed.size() > 10 && framed[0] == 0xff && framed[4] == 's');
        CHECK(Snappy::FrameDecompress(ro(framed)) == input);
        std::vector<uint8_t> gold = read_file(dir + "/golden_framed.snappy"), gold_raw = read_file(dir + "/golden_raw.bin");
        CHECK(!gold.empty() && Snappy::FrameDecompress(ro(gold)) == gold_raw);
        Snappy::HashMode() = SNP_HASH_MUL;  // the hash Snappier uses off x64/.NET 8+: reproduces the golden bytes
        CHECK(Snappy::FrameCompress(ro(gold_raw)) == gold);
        Snappy::HashMode() = SNP_HASH_CRC32C;
        framed[14] ^= 1;
        CHECK(throws<InvalidDataException>([&] { Snappy::FrameDecompress(ro(framed)); }));
    }
    std::printf(failures ? "%d FAILURES\n" : "ALL OK\n", failures);
    return failures != 0;
}
