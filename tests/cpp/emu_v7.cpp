// emu_v7.cpp -- runs the tag-group decompress engine's block function (snp_decompress_v7.cuh) on the host SIMT
// emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).  The TMA copies are synchronous here (snp_tma.cuh).
// Usage: emu_v7 <batch-in> <result-out> <window: 1024|2048|4096>   (same file formats as emu_v5 / emu_v6)
#include "simt_emu.h"
#define SNP_EMU 1
static unsigned long g_groups, g_group_tags, g_slow;  // fast-path coverage, printed with SNP7_EMU_STATS=1
#define SNP7_STAT(ng) do { if (simt::lane() == 0) { if (ng) { g_groups++; g_group_tags += (ng); } else g_slow++; } } while (0)
#include "../../snappier_b200/csrc/snp_decompress_v7.cuh"

#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

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

static uint8_t *g_smem;
static size_t g_ssz;
static void on_segv(int, siginfo_t *si, void *) {  // says where a lane left the warp's shared memory
    fprintf(stderr, "emu_v7: fault at shared-memory offset %ld (struct size %zu), lane %d\n",
            (long)((uint8_t *)si->si_addr - g_smem), g_ssz, simt::lane());
    void *bt[32];
    backtrace_symbols_fd(bt, backtrace(bt, 32), 2);
    _exit(139);
}

template <uint32_t W>
static int run_block(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t cap, uint32_t *w, snp::Warp7<W> *s,
                     uint32_t *phases) {
    int st = -77;
    simt::run_warp([&] {
        uint32_t ww = 0, ph = *phases;
        int r = snp::decompress_block_v7<W>(in, in_len, out, cap, &ww, s, ph);
        if (simt::lane() == 0) {
            *w = ww;
            st = r;
            *phases = ph;
        }
    });
    return st;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int window = atoi(argv[3]);
    const std::vector<uint8_t> raw = slurp(argv[1]);
    const uint8_t *p = raw.data();
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    // the warp's shared memory ends at a PROT_NONE page, so an access behind its struct faults here as it would on the
    // GPU for the last warp of a CTA; the content starts as garbage, like shared memory
    const size_t ssz = window == 1024 ? sizeof(snp::Warp7<1024>) : window == 2048 ? sizeof(snp::Warp7<2048>) : sizeof(snp::Warp7<4096>);
    const size_t page = (size_t)sysconf(_SC_PAGESIZE), span = (ssz + page - 1) / page * page;
    uint8_t *region = (uint8_t *)mmap(nullptr, span + 2 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (region == MAP_FAILED) return 2;
    mprotect(region, page, PROT_NONE);
    mprotect(region + page + span, page, PROT_NONE);
    void *smem = region + page + span - ssz;  // 16-byte aligned: sizeof is a multiple of 16
    memset(smem, 0xff, ssz);
    g_smem = (uint8_t *)smem;
    g_ssz = ssz;
    struct sigaction sa = {};
    sa.sa_sigaction = on_segv;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    static uint8_t altstack[1 << 16];
    stack_t ss = {altstack, 0, sizeof(altstack)};
    sigaltstack(&ss, nullptr);
    sigaction(SIGSEGV, &sa, nullptr);
    auto init = [&](auto *s) {  // what the kernel prologue does
        simt::run_warp([&] { snp::warp7_init(s, (unsigned)simt::lane()); });
    };
    if (window == 1024) init((snp::Warp7<1024> *)smem);
    else if (window == 2048) init((snp::Warp7<2048> *)smem);
    else init((snp::Warp7<4096> *)smem);
    uint32_t phases = 0;
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
        int st;
        if (window == 1024) st = run_block<1024>(in, in_len, out, cap, &w, (snp::Warp7<1024> *)smem, &phases);
        else if (window == 2048) st = run_block<2048>(in, in_len, out, cap, &w, (snp::Warp7<2048> *)smem, &phases);
        else st = run_block<4096>(in, in_len, out, cap, &w, (snp::Warp7<4096> *)smem, &phases);
        uint32_t guard_ok = 1;
        for (uint8_t *g = obuf.data(); g < out; g++) guard_ok &= *g == 0xAB;
        for (uint8_t *g = out + cap; g < obuf.data() + obuf.size(); g++) guard_ok &= *g == 0xAB;
        fwrite(&st, 4, 1, f);
        fwrite(&w, 4, 1, f);
        fwrite(&guard_ok, 4, 1, f);
        fwrite(out, 1, cap, f);
    }
    fclose(f);
    if (getenv("SNP7_EMU_STATS"))
        fprintf(stderr, "emu_v7: %lu groups, %.1f tags per group, %lu one-tag steps\n", g_groups,
                g_groups ? (double)g_group_tags / g_groups : 0.0, g_slow);
    return 0;
}
