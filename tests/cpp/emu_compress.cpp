// emu_compress.cpp -- runs the compress fragment functions (k_compress_v3's default path, its 16-bit-entry variant and
// the v1 baseline) on the host SIMT emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
// Usage: emu_compress <batch-in> <result-out> <variant: 1|3|6> <hash: 0|1> <first-width>
//   batch-in : u32 n, then per item { u32 len, bytes }      result: per item { u32 out_len, bytes }
#include "simt_emu.h"
#define SNP_EMU 1
#include "../../snappier_b200/csrc/snp_compress_v3.cuh"

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

template <int HASH>
static void run(const uint8_t *in, uint32_t n, snp::OutCursor &o, int variant, uint32_t w0, const uint16_t *lut,
                const uint32_t *sched, void *table) {
    simt::run_warp([&] {
        snp::OutCursor oc = o;  // every lane keeps its own cursor, as on the GPU
        const unsigned lane = snp::lane_id();
        uint32_t lo, hi;
        const int need = snp::varint_encode(n, &lo, &hi);  // SnappyCompressor.cs:34-38
        if ((int)lane < need) oc.put(lane, (uint8_t)(lane < 4 ? lo >> (8 * lane) : hi));
        oc.pos = need;
        if (n > 0) {
            if (variant == 1) snp::compress_fragment_v1<HASH>(in, n, oc, (uint16_t *)table, lut);
            else if (variant == 6) snp::compress_fragment_v3<HASH, false>(in, n, oc, (uint32_t *)table, lut, sched, w0);
            else snp::compress_fragment_v3<HASH, true>(in, n, oc, (uint32_t *)table, lut, sched, w0);
        }
        if (lane == 0) o.pos = oc.pos;
    });
}

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    const int variant = atoi(argv[3]), hash = atoi(argv[4]);
    const uint32_t w0 = (uint32_t)atoi(argv[5]);
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    uint16_t lut[1024];
    snp::build_crc_lut(lut, 0, 1);
    uint32_t sched[SNP_SCHED_LEN];
    {  // the data-independent probe schedule of a literal run (SnappyCompressor.cs:227,319-320)
        uint32_t skip = 32, off = 0;
        for (int k = 0; k < SNP_SCHED_LEN; k++) {
            const uint32_t stride = skip >> 5;
            sched[k] = std::min(off, 0xfffffu) | (std::min(stride, 0xfffu) << 20);
            off += stride;
            skip += stride;
        }
    }
    std::vector<uint32_t> table(16384, 0xdeadbeef);  // dirty: the fragment function must clear what it uses
    FILE *f = fopen(argv[2], "wb");
    for (uint32_t i = 0; i < n; i++) {
        uint32_t len;
        memcpy(&len, p, 4);
        p += 4;
        std::vector<uint8_t> ibuf(len + 64), obuf(len + len / 6 + 128, 0xAB);
        uint8_t *in = ibuf.data() + 16 + (i % 4);
        memcpy(in, p, len);
        p += len;
        snp::OutCursor o{obuf.data(), (uint32_t)obuf.size(), 0};
        if (hash == 0) run<SNP_HASH_CRC32C>(in, len, o, variant, w0, lut, sched, table.data());
        else run<SNP_HASH_MUL>(in, len, o, variant, w0, lut, sched, table.data());
        fwrite(&o.pos, 4, 1, f);
        fwrite(obuf.data(), 1, o.pos, f);
    }
    fclose(f);
    return 0;
}
