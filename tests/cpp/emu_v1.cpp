// emu_v1.cpp -- runs the baseline decompress block function (decompress_block_v1: the > 2 GiB / big-block path of the
// fast engines and the A/B reference) on the host SIMT emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
// Usage: emu_v1 <batch-in> <result-out> 1
//   batch-in: u32 n, then per item { u32 in_len, u32 cap, u32 in_misalign, u32 out_misalign, bytes }
//   result  : per item { i32 status, u32 written, u32 guard_ok, cap bytes of output }
#include "simt_emu.h"
#define SNP_EMU 1
#include "../../snappier_b200/csrc/snp_decompress_v1.cuh"

#include <vector>

static std::vector<uint8_t> slurp(const char *p) {
    FILE *f = fopen(p, "rb");
    if (!f) {
        perror(p);
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v(n);
    if (n && fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    FILE *f = fopen(argv[2], "wb");
    for (uint32_t i = 0; i < n; i++) {
        uint32_t h[4];
        memcpy(h, p, 16);
        p += 16;
        const uint32_t in_len = h[0], cap = h[1];
        // misaligned copies with guard bytes around the output region
        std::vector<uint8_t> ibuf(in_len + 64), obuf(cap + 192, 0xAB);
        uint8_t *in = ibuf.data() + 16 + (h[2] & 15), *out = obuf.data() + 64 + (h[3] & 15);
        memcpy(in, p, in_len);
        p += in_len;
        uint32_t w = 0xdeadbeef;
        int st = -77;
        simt::run_warp([&] {
            uint32_t ww = 0;
            int s = snp::decompress_block_v1(in, in_len, out, cap, &ww);
            if (simt::lane() == 0) {
                w = ww;
                st = s;
            }
        });
        uint32_t guard_ok = 1;
        for (uint8_t *g = obuf.data(); g < out; g++) guard_ok &= *g == 0xAB;
        for (uint8_t *g = out + cap; g < obuf.data() + obuf.size(); g++) guard_ok &= *g == 0xAB;
        fwrite(&st, 4, 1, f);
        fwrite(&w, 4, 1, f);
        fwrite(&guard_ok, 4, 1, f);
        fwrite(out, 1, cap, f);
    }
    fclose(f);
    return 0;
}
