// emu_v6.cpp -- runs the v6 decompress passes (snp_decompress_v6.cuh) on the host SIMT emulator.
//
// TEST INFRASTRUCTURE ONLY (see simt_emu.h).  Usage: emu_v6 <batch-in> <result-out>
//   batch-in : u32 n, then per item { u32 in_len, u32 cap, u32 in_skew, u32 out_skew, bytes[in_len] }
//   result   : per item { i32 status, u32 out_len, u32 guard_ok, bytes[cap] }
// Items are laid out with the requested misalignments and 64 guard bytes around every output
// region, so that a vector flush that strays outside [out, out + cap) is caught.
#include "simt_emu.h"
#define SNP_EMU 1
#include "../../snappier_b200/csrc/snp_decompress_v6.cuh"

#include <vector>

namespace snp {
int v6_emu_fallback(const uint8_t *, uint32_t, uint8_t *, uint32_t, uint32_t *written) {
    *written = 0;
    return 100;  // marker: "this block would take the v5 / v1 engine"
}
}  // namespace snp

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
    if (argc < 3) return 2;
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    std::vector<uint64_t> in_off(n), out_off(n);
    std::vector<uint32_t> in_len(n), out_cap(n), out_len(n, 0xdeadbeef);
    std::vector<int32_t> status(n, -77);
    const size_t kGuard = 64;
    size_t in_total = 64, out_total = 64;
    std::vector<const uint8_t *> src(n);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t h[4];
        memcpy(h, p, 16);
        p += 16;
        in_len[i] = h[0];
        out_cap[i] = h[1];
        src[i] = p;
        p += h[0];
        in_total = (in_total + 15) / 16 * 16 + (h[2] & 15);
        in_off[i] = in_total;
        in_total += h[0];
        out_total = (out_total + kGuard + 15) / 16 * 16 + (h[3] & 15);
        out_off[i] = out_total;
        out_total += h[1];
    }
    in_total += 64;
    out_total += kGuard + 64;
    uint8_t *in_base = (uint8_t *)aligned_alloc(64, (in_total + 63) / 64 * 64);
    uint8_t *out_base = (uint8_t *)aligned_alloc(64, (out_total + 63) / 64 * 64);
    memset(in_base, 0xEE, in_total);
    memset(out_base, 0xAB, out_total);
    for (uint32_t i = 0; i < n; i++) memcpy(in_base + in_off[i], src[i], in_len[i]);

    std::vector<uint32_t> ntags(n, 0x12345678);
    std::vector<uint2> ck((size_t)n * SNP6_CKB);
    uint32_t lut[256];
    for (int c = 0; c < 256; c++) lut[c] = snp::tag_lut3_entry(c);

    // ---- pass A ----
    unsigned long long ctr = 0;
    snp::Scan6Args sa{in_base, in_off.data(), in_len.data(), out_cap.data(), out_len.data(), status.data(),
                      0, n, &ctr, ntags.data(), ck.data()};
    std::vector<uint4> rings(32 * 3);
    simt::run_warp([&] { snp::tagscan_warp_v6(sa, lut, rings.data() + simt::lane() * 3); });

    // ---- pass B ----
    ctr = 0;
    snp::Decode6Args da{in_base, in_off.data(), in_len.data(), out_base, out_off.data(), out_cap.data(),
                        out_len.data(), status.data(), 0, n, &ctr, ntags.data(), ck.data()};
    snp::V6Smem *sm = (snp::V6Smem *)aligned_alloc(16, sizeof(snp::V6Smem));
    memset(sm, 0xCD, sizeof(*sm));
    simt::run_warp([&] { snp::decode_warp_v6(da, lut, sm); });

    if (getenv("SNP6_STATS")) {
        const snp::V6Stats &t = snp::v6_stats();
        fprintf(stderr, "v6 stats: tags %lu groups %lu subgroups %lu (fast %lu) rounds %lu trips %lu ctags %lu huge %lu slides %lu flushes %lu hops %lu stuck(kind %lu straddle %lu)\n",
                t.tags, t.groups, t.subgroups, t.fast, t.rounds, t.trips, t.ctags, t.huge, t.slides, t.flushes, t.hops, t.stuck_kind, t.stuck_straddle);
    }
    FILE *f = fopen(argv[2], "wb");
    for (uint32_t i = 0; i < n; i++) {
        // guard check: everything between this item's region and its neighbours must be untouched
        uint32_t guard_ok = 1;
        const uint64_t lo = i ? out_off[i - 1] + out_cap[i - 1] : 0;
        for (uint64_t k = lo; k < out_off[i]; k++) guard_ok &= out_base[k] == 0xAB;
        if (i + 1 == n)
            for (uint64_t k = out_off[i] + out_cap[i]; k < out_total; k++) guard_ok &= out_base[k] == 0xAB;
        // bytes of the capacity region beyond out_len must be untouched too
        if (status[i] == 0)
            for (uint64_t k = out_off[i] + out_len[i]; k < out_off[i] + out_cap[i]; k++) guard_ok &= out_base[k] == 0xAB;
        fwrite(&status[i], 4, 1, f);
        fwrite(&out_len[i], 4, 1, f);
        fwrite(&guard_ok, 4, 1, f);
        fwrite(out_base + out_off[i], 1, out_cap[i], f);
    }
    fclose(f);
    free(sm);
    free(in_base);
    free(out_base);
    return 0;
}
