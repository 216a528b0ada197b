// emu_v8.cpp -- runs the lane-per-block decompress engine (snp_decompress_v8.cuh) on the host SIMT emulator: ONE warp,
// every lane with its own block, the whole batch handed out through the work counter exactly as in the kernel.
// TEST INFRASTRUCTURE ONLY (see simt_emu.h).  cp.async copies are synchronous here.
// Usage: emu_v8 <batch-in> <result-out> <instantiation: 128 | 1282 | 1284 | 1281 | 256 = ring bytes and pipeline depth>   (same file formats as emu_v5 / emu_v7)
#include "simt_emu.h"
#define SNP_EMU 1
static unsigned long g_iters, g_moves;  // printed with SNP8_EMU_STATS=1
#define SNP8_STAT(what) do { if (what) g_moves++; else if (simt::lane() == 0) g_iters++; } while (0)
#include "../../snappier_b200/csrc/snp_decompress_v8.cuh"

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

template <uint32_t IR, uint32_t ORB, int D>
static void run(size_t n, const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out_base,
                const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, int32_t *status) {
    static snp::Lane8<IR, ORB, D> lanes[32];
    memset(lanes, 0xff, sizeof(lanes));  // shared memory starts as garbage
    unsigned long long counter = 0;
    simt::run_warp([&] {
        snp::decompress_lanes_v8<IR, ORB, D>(in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n,
                                          &counter, &lanes[simt::lane()]);
    });
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int orb = atoi(argv[3]);
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    // one input arena and one output arena, every item at its own misalignment, guard bytes between the outputs
    std::vector<uint64_t> in_off(n), out_off(n);
    std::vector<uint32_t> in_len(n), out_cap(n), out_len(n, 0xdeadbeef);
    std::vector<int32_t> status(n, -77);
    std::vector<const uint8_t *> src(n);
    size_t itot = 64, otot = 64;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t h[4];
        memcpy(h, p, 16);
        p += 16;
        in_len[i] = h[0];
        out_cap[i] = h[1];
        src[i] = p;
        p += h[0];
        itot = (itot + 15) / 16 * 16 + (h[2] & 15);
        in_off[i] = itot;
        itot += h[0] + 16;
        otot = (otot + 15) / 16 * 16 + 64 + (h[3] & 15);
        out_off[i] = otot;
        otot += h[1] + 64;
    }
    std::vector<uint8_t> ibuf(itot + 64, 0xCD), obuf(otot + 64, 0xAB);
    // 16-byte aligned arena starts, so that the requested misalignments are the real ones
    uint8_t *ib = ibuf.data() + ((16 - ((uintptr_t)ibuf.data() & 15)) & 15);
    uint8_t *ob = obuf.data() + ((16 - ((uintptr_t)obuf.data() & 15)) & 15);
    for (uint32_t i = 0; i < n; i++) memcpy(ib + in_off[i], src[i], in_len[i]);
#define RUN8(IR, ORB, D) run<IR, ORB, D>(n, ib, in_off.data(), in_len.data(), ob, out_off.data(), out_cap.data(), out_len.data(), status.data())
    if (orb == 128) RUN8(128, 128, 3);       // the argument picks one of the kernel's instantiations
    else if (orb == 1282) RUN8(128, 128, 2);
    else if (orb == 1284) RUN8(128, 128, 4);
    else if (orb == 1281) RUN8(128, 128, 1);
    else RUN8(256, 256, 3);
    // guard check: every byte of the arena outside the items' [out, out + cap) regions is untouched
    std::vector<uint8_t> mine(obuf.size(), 0);
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t k = 0; k < out_cap[i]; k++) mine[(ob - obuf.data()) + out_off[i] + k] = 1;
    std::vector<uint32_t> guard(n, 1);
    for (size_t k = 0; k < obuf.size(); k++) {
        if (mine[k] || obuf[k] == 0xAB) continue;
        // attribute the stray write to the nearest item
        uint32_t best = 0;
        size_t bd = (size_t)-1;
        for (uint32_t i = 0; i < n; i++) {
            const size_t a = (ob - obuf.data()) + out_off[i];
            const size_t d = k < a ? a - k : k - a;
            if (d < bd) bd = d, best = i;
        }
        guard[best] = 0;
    }
    FILE *f = fopen(argv[2], "wb");
    for (uint32_t i = 0; i < n; i++) {
        fwrite(&status[i], 4, 1, f);
        fwrite(&out_len[i], 4, 1, f);
        fwrite(&guard[i], 4, 1, f);
        fwrite(ob + out_off[i], 1, out_cap[i], f);
    }
    fclose(f);
    if (getenv("SNP8_EMU_STATS"))
        fprintf(stderr, "emu_v8: %lu warp iterations, %lu lane moves (%.1f lanes busy per iteration)\n", g_iters, g_moves,
                g_iters ? (double)g_moves / g_iters : 0.0);
    return 0;
}
