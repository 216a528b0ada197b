// emu_crc.cpp -- the warp-parallel CRC32C of snp_frame.cuh on the host SIMT emulator.  TEST INFRASTRUCTURE ONLY.
// Usage: emu_crc <batch-in> <result-out>;  batch: u32 n, per item { u32 len, u32 skew, bytes };  result: per item u32 crc, u32 masked
#include "simt_emu.h"
#define SNP_EMU 1
#include "../../snappier_b200/csrc/snp_frame.cuh"

#include <vector>

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 2;
    fseek(fi, 0, SEEK_END);
    long sz = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    std::vector<uint8_t> raw(sz);
    if (fread(raw.data(), 1, sz, fi) != (size_t)sz) return 2;
    fclose(fi);
    uint32_t tab[2048];
    for (unsigned i = 0; i < 1024; i++) {  // what k_init_crc_tables builds
        const unsigned k = i >> 8, b = i & 255;
        const uint32_t t4 = snp::crc_advance_bits(b, 8 * (4 - k));
        tab[i] = t4;
        tab[1024 + i] = snp::crc_advance_bits(t4, 8 * 124);
    }
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    FILE *fo = fopen(argv[2], "wb");
    for (uint32_t i = 0; i < n; i++) {
        uint32_t h[2];
        memcpy(h, p, 8);
        p += 8;
        std::vector<uint8_t> buf(h[0] + 32);
        uint8_t *d = buf.data() + 8 + (h[1] & 7);
        memcpy(d, p, h[0]);
        p += h[0];
        uint32_t crc = 0;
        simt::run_warp([&] {
            const uint32_t c = snp::crc32c_warp(tab, d, h[0]);
            if (simt::lane() == 0) crc = c;
        });
        const uint32_t m = snp::crc32c_mask(crc);
        fwrite(&crc, 4, 1, fo);
        fwrite(&m, 4, 1, fo);
    }
    fclose(fo);
    return 0;
}
