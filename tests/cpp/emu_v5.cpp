// emu_v5.cpp -- runs the default decompress engine's block functions (v5 = sparse prefix engine + v3 dense engine,
// plus the v1 baseline) on the host SIMT emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
// Usage: emu_v5 <batch-in> <result-out> <engine: 1|3|5>   (same file formats as emu_v6)
#include "simt_emu.h"
#define SNP_EMU 1
#include "../../snappier_b200/csrc/snp_decompress_v5.cuh"

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
    const int engine = atoi(argv[3]);
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    uint32_t lut[256];
    for (int c = 0; c < 256; c++) lut[c] = snp::tag_lut3_entry(c);
    snp::WarpQueue3 *q = (snp::WarpQueue3 *)aligned_alloc(16, sizeof(snp::WarpQueue3));
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
            int s = engine == 1   ? snp::decompress_block_v1(in, in_len, out, cap, &ww)
                    : engine == 3 ? snp::decompress_block_v3(in, in_len, out, cap, &ww, lut, q)
                                  : snp::decompress_block_v5(in, in_len, out, cap, &ww, lut, q);
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
    free(q);
    return 0;
}
