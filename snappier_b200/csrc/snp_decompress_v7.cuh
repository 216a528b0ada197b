// snp_decompress_v7.cuh -- tag-group batched Snappy block decompressor (sm_100a), one warp per block.
//
// Written against the ncu picture of v3/v5 (profiles/r01_decompress_v5_ncu.md): those kernels are
// bound by warp-instruction issue (4.3 instructions per output byte; a 32-position speculative parse
// that finds ~10 tags, byte-per-lane rounds with a shared-memory queue lookup, byte loads of
// back-references from global memory).  This engine spends its instructions differently:
//
//  * INPUT: the compressed block streams through a per-warp 1 KiB shared-memory ring filled by TMA
//    bulk copies (cp.async.bulk + mbarrier, 256-byte chunks, 4 in flight).  The <= 15 unaligned head /
//    tail bytes of a block are lane copies (bulk copies need 16-byte alignment and must not read
//    outside the caller's buffer).
//  * PARSE: when a chunk lands, every lane turns 8 input bytes into 8 "advance" bytes with
//    SIMD-in-word arithmetic (bytes to the next tag if a tag started here; 0 = not decodable by the
//    fast path: literal length in trailer bytes, 4-byte-offset copy, tag running past the end of the
//    input).  Finding the tag starts -- the serial dependency chain of the format
//    (SnappyDecompressor.cs:332-333) -- is then `i += adv[i]`: one shared-memory load and one add per
//    tag, 32 tags per group, instead of decoding a tag at every byte position.
//  * GROUP: lane k decodes tag k (one lane per TAG, no idle positions), a 5-step shuffle scan of the
//    lengths gives every tag its output offset, copy offsets / lengths are validated in stream order
//    (SnappyDecompressor.cs:570-573,598-606); the group is cut in front of the first tag that is
//    special, invalid or not yet landed.
//  * COPY: output-centric rounds of 32 bytes; the owning tag of a byte is a popcount rank over the
//    round's tag-start mask, its descriptor comes from the owning lane's register by one SHFL.
//    Sources are the input ring (literals), the OUTPUT WINDOW (back-references up to W-64 bytes, both
//    shared memory) or global memory (older output); a source produced in the same round is resolved
//    by pointer doubling over SHFL -- CopyHelpers.IncrementalCopy's pattern replication
//    (CopyHelpers.cs:64-219), exact for any offset.
//  * OUTPUT: every byte is written to the per-warp output window (a ring indexed by the low bits of
//    the global output address, so alignment carries over) and flushed to HBM as aligned 16-byte
//    vectors, 512+ bytes at a time.  Literals >= 512 bytes bypass the window (vector copy from the
//    input to the output) and re-prime it.
//  * Everything the fast path cannot decode takes a warp-uniform one-tag path with the v1 semantics
//    (truncated tags, partial literals, long literals, COPY4, and every error status).
//
// Semantics: /root/reference/Snappier/Internal/SnappyDecompressor.cs:43-92,184-347,556-611
// (one-shot); identical results to v1 and oracle/snappy_oracle.c.  The block function also compiles
// against tests/cpp/simt_emu.h (tests/test_emu_v7.py).
#pragma once
#include "snp_common.cuh"
#include "snp_decompress_v1.cuh"
#include "snp_tma.cuh"

namespace snp {

#define SNP7_R 1024u       // input ring bytes per warp
#define SNP7_CH 256u       // bytes per TMA chunk / mbarrier
#define SNP7_SLOTS (SNP7_R / SNP7_CH)
#define SNP7_PAD 64u       // zeros behind the advance table: a walk that leaves the ring lap stops here
#define SNP7_LOOKAHEAD 320u
#define SNP7_DIRECT_MIN 512u  // literals at least this long are copied input -> output directly
#ifndef SNP7_STAT
#define SNP7_STAT(ng)  // tests/cpp/emu_v7.cpp counts fast-path tags per group with this hook
#endif

#define SNP7_MIRROR 64u    // ring[R, R+64) repeats ring[0, 64): a tag that starts in front of the lap end is read linearly
#define SNP7_STAGE_W 5u    // words per lane of the far-source staging area (sources of <= 16 bytes at any alignment)

template <uint32_t W>
struct alignas(16) Warp7 {
    uint8_t win[W];                      // output window: position x lives at (out address + x) mod W
    uint8_t ring[SNP7_R + SNP7_MIRROR];  // compressed input: position p lives at (in address + p) mod R
    uint8_t adv[SNP7_R + SNP7_PAD];      // advance table, same indexing as ring
    uint32_t stage[32 * SNP7_STAGE_W];   // sources of this group's far back-references, fetched once per group
    uint32_t tagpos[32];                 // shared-memory address of the advance byte of tag k of the current group
    uint64_t bar[SNP7_SLOTS];
};

// Four tag bytes -> four advances (Constants.CharTable's "tag size" column, Constants.cs:42-76, as arithmetic):
// literal with inline length: 1 + len = n6 + 2;  COPY1: 2;  COPY2: 3;  0 for trailer-length literals and COPY4.
__device__ __forceinline__ uint32_t adv4_v7(uint32_t w) {
    const uint32_t t = w & 0x03030303u;
    const uint32_t n6 = (w >> 2) & 0x3f3f3f3fu;
    const uint32_t t0 = t & 0x01010101u, t1 = (t >> 1) & 0x01010101u;
    const uint32_t nz = t0 | t1;  // 1 = copy
    const uint32_t cm = nz * 0xffu;
    const uint32_t a = ((n6 + 0x02020202u) & ~cm) | ((t + 0x01010101u) & cm);
    const uint32_t big = ((n6 + 0x04040404u) >> 6) & 0x01010101u & ~nz;  // literal, n6 >= 60
    const uint32_t sp = (big | (t0 & t1)) * 0xffu;
    return a & ~sp;
}

// Cooperative copy of a long literal, input -> output, as 16-byte vectors aligned on the destination.
__device__ __forceinline__ void copy_literal_wide7(const uint8_t *__restrict__ s, uint8_t *d, uint32_t len,
                                                   const uint8_t *in_end, unsigned lane) {
    const uint32_t h = min((uint32_t)(-(intptr_t)d) & 15u, len);  // bytes up to the first 16-byte boundary of d
    if (lane < h) d[lane] = s[lane];
    const uint8_t *sv = s + h;
    uint4 *dv = reinterpret_cast<uint4 *>(d + h);
    const uint32_t nvec = (len - h) >> 4;
    const unsigned sb = (unsigned)((uintptr_t)sv & 15);  // warp-uniform source misalignment
    const uint4 *base = reinterpret_cast<const uint4 *>(sv - sb);
    const uint4 *last = reinterpret_cast<const uint4 *>(((uintptr_t)in_end - 1) & ~(uintptr_t)15);  // last readable vector
    const unsigned wo = sb >> 2, bs = (sb & 3) * 8;
    for (uint32_t v = lane; v < nvec; v += SNP_WARP) {
        const uint4 A = base[v];
        const uint4 *pb = base + v + 1;
        const uint4 B = *(pb <= last ? pb : last);  // only used when sb != 0; clamped at the buffer end
        const uint32_t Wd[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
        uint4 r;
        switch (wo) {  // warp-uniform
            case 0: r = make_uint4(__funnelshift_r(Wd[0], Wd[1], bs), __funnelshift_r(Wd[1], Wd[2], bs),
                                   __funnelshift_r(Wd[2], Wd[3], bs), __funnelshift_r(Wd[3], Wd[4], bs)); break;
            case 1: r = make_uint4(__funnelshift_r(Wd[1], Wd[2], bs), __funnelshift_r(Wd[2], Wd[3], bs),
                                   __funnelshift_r(Wd[3], Wd[4], bs), __funnelshift_r(Wd[4], Wd[5], bs)); break;
            case 2: r = make_uint4(__funnelshift_r(Wd[2], Wd[3], bs), __funnelshift_r(Wd[3], Wd[4], bs),
                                   __funnelshift_r(Wd[4], Wd[5], bs), __funnelshift_r(Wd[5], Wd[6], bs)); break;
            default: r = make_uint4(__funnelshift_r(Wd[3], Wd[4], bs), __funnelshift_r(Wd[4], Wd[5], bs),
                                    __funnelshift_r(Wd[5], Wd[6], bs), __funnelshift_r(Wd[6], Wd[7], bs)); break;
        }
        dv[v] = r;
    }
    const uint32_t done = h + (nvec << 4);
    if (done + lane < len) d[done + lane] = s[done + lane];  // < 16 tail bytes
}


// ---- small shared-memory / shift helpers (CUDA: exact instructions; emulator: plain C) -------------------
#ifdef SNP_EMU
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t sh) { return sh >= 32 ? 0u : v << sh; }
template <class S> __device__ __forceinline__ uint32_t sm_off(const S *s, const void *p) { return (uint32_t)((const uint8_t *)p - (const uint8_t *)s); }
template <class S> __device__ __forceinline__ uint32_t sm_ld8(const S *s, uint32_t a) { return ((const uint8_t *)s)[a]; }
#else
// PTX shl clamps shift amounts above 31 to "all bits out" (C's << is undefined there)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t sh) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(sh));
    return r;
}
// 32-bit shared-window address of p (so that address arithmetic stays in one register)
template <class S> __device__ __forceinline__ uint32_t sm_off(const S *, const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <class S> __device__ __forceinline__ uint32_t sm_ld8(const S *, uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
#endif

// Once per warp, before its first block: barriers; an all-zero advance table (the walk may run through entries of
// chunks that have not landed: they must be advances a chunk once produced (<= 61) or 0, never uninitialised shared
// memory, and the pad behind the table stays 0 for good); tag slots that point into the advance table (lanes behind a
// short walk read whatever their slot held last).
template <uint32_t W>
__device__ __forceinline__ void warp7_init(Warp7<W> *s, unsigned lane) {
    if (lane < SNP7_SLOTS) mbar_init(&s->bar[lane], 1);
    s->tagpos[lane] = sm_off(s, s->adv);
    for (uint32_t k = lane; k < (SNP7_R + SNP7_PAD) / 4; k += SNP_WARP) reinterpret_cast<uint32_t *>(s->adv)[k] = 0;
}

template <uint32_t W>
__device__ __noinline__ int decompress_block_v7(const uint8_t *__restrict__ in, uint32_t n_in, uint8_t *out,
                                                uint32_t cap, uint32_t *written, Warp7<W> *s, uint32_t &phases) {
    constexpr uint32_t R = SNP7_R, CH = SNP7_CH;
    constexpr uint32_t FLUSH_AT = W / 4;                     // flush the window when this much is pending
    constexpr uint32_t GMAX = W / 2 < 1024 ? W / 2 : 1024;   // output bytes per tag group (<= 32 rounds)
    constexpr uint32_t NEAR_MAX = W - 64;                    // back-references up to this distance are served by the window
    static_assert((W & (W - 1)) == 0 && W >= 1024 && W <= 32768, "window size");
    constexpr uint32_t RING_OFF = W;                         // byte offsets inside Warp7<W>; win sits at 0
    constexpr uint32_t STAGE_OFF = W + 2 * SNP7_R + SNP7_MIRROR + SNP7_PAD;
    // descriptor flags (see GROUP below)
    constexpr uint32_t D_WIN = 0x80000000u, D_GLOBAL = 0x40000000u, D_CLOSE = 0x20000000u;
    const unsigned lane = lane_id();
    const unsigned le = (2u << lane) - 1u;  // lanes <= me
    *written = 0;
    uint32_t U, used;
    int st = varint_read(in, n_in, &U, &used);
    if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;
    if (st != SNP_OK || U > 0x7fffffffu) return SNP_INVALID_LENGTH;
    if (cap < U) return SNP_OUTPUT_TOO_SMALL;
    if (U == 0) return SNP_OK;

    // 16-byte-aligned views: stream byte p sits at aligned offset skew + p of in16 (ring index (skew + p) mod R),
    // output byte x at aligned offset ("window coordinate") oskew + x of out16 (window index (oskew + x) mod W).
    const uint32_t skew = (uint32_t)((uintptr_t)in & 15);
    const uint8_t *in16 = in - skew;
    const uint32_t A = skew + n_in;
    const uint32_t n_chunks = (A + CH - 1) / CH;
    const uint32_t oskew = (uint32_t)((uintptr_t)out & 15);
    uint8_t *out16 = out - oskew;
    uint8_t *sb = reinterpret_cast<uint8_t *>(s);

    uint32_t issued = 0, ready = 0;  // chunks [ready, issued) are in flight
    uint32_t refill_at = 0, landed_end = 0;
    uint32_t ip = used, op = 0, f = 0;  // f: output bytes < f are in global memory

    // ---- input ring ----------------------------------------------------------------------------
    auto chunk_adv = [&](uint32_t c) {
        const uint32_t base = c * CH;
        const uint32_t ri = (base & (R - 1)) + 8 * lane;
        const uint2 w = *reinterpret_cast<const uint2 *>(s->ring + ri);
        uint32_t a0 = adv4_v7(w.x), a1 = adv4_v7(w.y);
        if (base + CH + 64 > A) {  // uniform: the block ends in or right behind this chunk -- tags that run past
            const uint32_t p0 = base + 8 * lane;  // the end (or start behind it) are left to the one-tag path
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                if (p0 + j + ((a0 >> (8 * j)) & 0xff) > A) a0 &= ~(0xffu << (8 * j));
                if (p0 + 4 + j + ((a1 >> (8 * j)) & 0xff) > A) a1 &= ~(0xffu << (8 * j));
            }
        }
        *reinterpret_cast<uint2 *>(s->adv + ri) = make_uint2(a0, a1);
        if ((c & (SNP7_SLOTS - 1)) == 0 && lane < SNP7_MIRROR / 4)  // a new lap begins: repeat its head behind the ring
            reinterpret_cast<uint32_t *>(s->ring)[R / 4 + lane] = reinterpret_cast<const uint32_t *>(s->ring)[lane];
    };
    auto wait_chunk = [&](uint32_t c, bool compute) {
        const uint32_t slot = c & (SNP7_SLOTS - 1);
        uint32_t spins = 0;
        while (!mbar_try_wait(&s->bar[slot], (phases >> slot) & 1)) {
#ifndef SNP_EMU
            if (++spins > (1u << 26)) __trap();  // a lost completion must abort, never hang the GPU
#endif
        }
        (void)spins;
        phases ^= 1u << slot;
        if (compute) chunk_adv(c);
        __syncwarp();
    };
    auto issue_chunk = [&](uint32_t c) {
        const uint32_t slot = c & (SNP7_SLOTS - 1);
        const uint32_t c0 = c * CH;
        if (c != 0 && c0 + CH <= (A & ~15u)) {  // uniform: interior chunk = one 256-byte bulk copy
            if (lane == 0) {
                mbar_arrive_expect_tx(&s->bar[slot], CH);
                tma_load_1d(s->ring + slot * CH, in16 + c0, CH, &s->bar[slot]);
            }
            return;
        }
        const uint32_t blo = (skew + 15u) & ~15u, bhi = A & ~15u;  // [blo, bhi) may be bulk-copied
        const uint32_t c1 = min(c0 + CH, (A + 15u) & ~15u);
        const uint32_t b0 = max(c0, blo), b1 = min(c1, bhi);
        if (lane == 0) {
            if (b1 > b0) {
                mbar_arrive_expect_tx(&s->bar[slot], b1 - b0);
                tma_load_1d(s->ring + (b0 & (R - 1)), in16 + b0, b1 - b0, &s->bar[slot]);
            } else {
                mbar_arrive(&s->bar[slot]);
            }
        }
        // unaligned head [skew, blo) and tail [bhi, A) bytes that fall into this chunk: lane copies
        const uint32_t h0 = max(c0, skew), h1 = min(min(c1, blo), A);
        if (h1 > h0 && h0 + lane < h1) s->ring[(h0 + lane) & (R - 1)] = in16[h0 + lane];
        const uint32_t t0 = max(max(c0, bhi), blo), t1 = min(c1, A);
        if (t1 > t0 && t0 + lane < t1) s->ring[(t0 + lane) & (R - 1)] = in16[t0 + lane];
    };
    auto set_landed = [&]() { landed_end = ready >= n_chunks ? n_in : (ready ? ready * CH - skew : 0u); };
    // drop the chunks below stream position pos, prefetch as far as the slots allow
    auto refill = [&](uint32_t pos) {
        const uint32_t keep = (skew + pos) / CH;  // oldest live chunk
        if (keep > issued) {  // jumped over chunks that were never needed (long literal)
            while (ready < issued) wait_chunk(ready++, false);
            issued = ready = keep;
        }
        while (ready < keep && ready < issued) wait_chunk(ready++, false);  // a slot is re-armed only after its wait
        __syncwarp();  // every lane is done reading the slots that are about to be overwritten
        while (issued < n_chunks && issued < keep + SNP7_SLOTS) issue_chunk(issued++);
        __syncwarp();
        refill_at = (keep + 1) * CH - skew;
        set_landed();
    };
    auto ensure = [&](uint32_t pos_end) {  // chunks covering stream bytes < pos_end have landed (and have advances)
        const uint32_t need = min((skew + pos_end - 1) / CH, n_chunks - 1);
        while (ready <= need && ready < issued) wait_chunk(ready++, true);
        set_landed();
    };
    auto finish = [&]() {  // nothing may be in flight when the warp moves on to its next block
        while (ready < issued) wait_chunk(ready++, false);
    };

    // ---- output window -------------------------------------------------------------------------
    auto flush_to = [&](uint32_t e) {  // window bytes [f, e) -> global memory
        uint32_t g = oskew + f;
        const uint32_t ge = oskew + e;
        const uint32_t h = min((0u - g) & 15u, ge - g);
        if (lane < h) out16[g + lane] = s->win[(g + lane) & (W - 1)];
        g += h;
        const uint32_t nvec = (ge - g) >> 4;
        for (uint32_t v = lane; v < nvec; v += SNP_WARP)
            *reinterpret_cast<uint4 *>(out16 + g + 16 * v) = *reinterpret_cast<const uint4 *>(s->win + ((g + 16 * v) & (W - 1)));
        g += nvec << 4;
        if (g + lane < ge) out16[g + lane] = s->win[(g + lane) & (W - 1)];
        f = e;
        __syncwarp();  // later back-reference loads from global memory see these stores
    };
    auto maybe_flush = [&]() {
        if (op - f >= FLUSH_AT) flush_to(op - ((oskew + op) & 15u));
    };

    // ---- one tag at ip, warp-uniform, v1 semantics.  Returns -1 (go on), -2 (the tag stream ended) or a status.
    auto one_tag = [&]() -> int {
        const uint32_t c = in[ip];
        const uint32_t kind = c & 3;
        const uint32_t extra = kind == 0 ? ((c >> 2) >= 60 ? (c >> 2) - 59 : 0) : (kind == 1 ? 1 : kind == 2 ? 2 : 4);
        if (n_in - ip < 1 + extra) return -2;  // truncated tag (RefillTag, SnappyDecompressor.cs:464-483)
        uint32_t trailer = 0;
        for (uint32_t i = 0; i < extra; i++) trailer |= (uint32_t)in[ip + 1 + i] << (8 * i);
        ip += 1 + extra;
        if (kind == 0) {
            const uint64_t len = (uint64_t)((c >> 2) >= 60 ? trailer : (c >> 2)) + 1;  // :264-288
            const uint32_t avail = n_in - ip;
            const uint32_t take = len < avail ? (uint32_t)len : avail;  // partial literal, :290-297
            if (take > U - op) return SNP_DATA_TOO_LONG;                // :570-573
            if (take >= SNP7_DIRECT_MIN) {
                flush_to(op);
                copy_literal_wide7(in + ip, out + op, take, in + n_in, lane);
                const uint32_t keepw = min(take, NEAR_MAX);  // the window must still hold the most recent output
                for (uint32_t k = take - keepw + lane; k < take; k += SNP_WARP)
                    s->win[(oskew + op + k) & (W - 1)] = in[ip + k];
                __syncwarp();
                op += take;
                f = op;
            } else {
                for (uint32_t k = lane; k < take; k += SNP_WARP) s->win[(oskew + op + k) & (W - 1)] = in[ip + k];
                __syncwarp();
                op += take;
                maybe_flush();
            }
            ip += take;
            if (take < len) return -2;
        } else {
            uint32_t len, offset;
            if (kind == 1) {
                len = ((c >> 2) & 7) + 4;
                offset = ((c >> 5) << 8) | trailer;
            } else {
                len = (c >> 2) + 1;
                offset = trailer;
            }
            if (offset == 0 || op < offset) return SNP_INVALID_COPY_OFFSET;  // :598-601
            if (len > U - op) return SNP_DATA_TOO_LONG;                       // :603-606
            const bool pattern = offset < SNP_WARP && offset < len;
            for (uint32_t k0 = 0; k0 < len; k0 += SNP_WARP) {  // chunk j only reads bytes < op + 32j
                const uint32_t k = k0 + lane;
                if (k < len) {
                    const uint32_t sp = pattern ? op - offset + (k % offset) : op + k - offset;
                    const uint8_t v = offset <= NEAR_MAX ? s->win[(oskew + sp) & (W - 1)] : out[sp];
                    s->win[(oskew + op + k) & (W - 1)] = v;
                }
                __syncwarp();
            }
            op += len;
            maybe_flush();
        }
        return -1;
    };

    int status = SNP_OK;
    const uint32_t adv0 = sm_off(s, s->adv);
    refill(ip);
    while (ip < n_in) {
        if (ip >= refill_at) refill(ip);
        if (ip + SNP7_LOOKAHEAD > landed_end && landed_end < n_in) ensure(ip + SNP7_LOOKAHEAD);

        // ---- WALK: shared-memory address of the advance byte of each of the next 32 tags (the walk stops advancing
        //      at anything the fast path cannot take: advance 0).  One load, one add, one store per tag.
        const uint32_t a_ip = adv0 + ((skew + ip) & (R - 1));   // address of ip's advance byte
        const uint32_t a_land = a_ip + (landed_end - ip);       // address-space image of landed_end (may pass the lap end)
        const uint32_t a_lim = min(a_land, adv0 + R);
        uint32_t nwalk = 0;
        {
            uint32_t ai = a_ip;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                if (ai < a_lim) {  // uniform
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        s->tagpos[o * 8 + k] = ai;
                        ai += sm_ld8(s, ai);
                    }
                    nwalk += 8;
                }
            }
        }
        __syncwarp();

        // ---- GROUP: lane k decodes tag k.  Lanes behind the last usable tag compute garbage that nobody uses: the
        //      scan is a prefix operation and the group is cut at the first lane that is not `ok`.
        const uint32_t ak = s->tagpos[lane];
        const uint32_t a = sm_ld8(s, ak);
        const uint32_t ik = ak - adv0;  // ring index (< R for every usable tag; the mirror covers ik + 1, ik + 2)
        const uint32_t c = s->ring[ik], b1 = s->ring[ik + 1], b2 = s->ring[ik + 2];
        const uint32_t n6 = c >> 2;
        const bool is_copy = (c & 3) != 0;
        const bool is_c1 = (c & 3) == 1;
        const uint32_t off = is_c1 ? (((c >> 5) << 8) | b1) : (b1 | (b2 << 8));
        const uint32_t len = is_c1 ? (n6 & 7) + 4 : n6 + 1;
        uint32_t incl = len;
#pragma unroll
        for (int dlt = 1; dlt < SNP_WARP; dlt <<= 1) {
            const uint32_t y = __shfl_up_sync(SNP_FULL, incl, dlt);
            if (lane >= (unsigned)dlt) incl += y;
        }
        const uint32_t dst = op + (incl - len);
        // usable: landed + decodable; valid in stream order (SnappyDecompressor.cs:570-573,598-606 -- the one-tag path
        // names the error); the group stays below GMAX output bytes
        const bool ok = lane < nwalk && a != 0 && ak + a <= a_land && !(is_copy && off - 1u >= dst) && dst + len <= U &&
                        incl <= GMAX;
        const unsigned okm = __ballot_sync(SNP_FULL, ok);
        const uint32_t ng = okm == SNP_FULL ? 32u : (uint32_t)__ffs(~okm) - 1u;
        SNP7_STAT(ng);
        if (ng == 0) {
            const int r = one_tag();
            if (r == -1) continue;
            if (r != -2) status = r;
            break;
        }
        const uint32_t glen = __shfl_sync(SNP_FULL, incl, ng - 1);
        const uint32_t ip_next = ip + (__shfl_sync(SNP_FULL, ak + a, ng - 1) - a_ip);
        const bool mine = lane < ng;
        const uint32_t wxd = oskew + dst;  // window coordinate of my tag's first byte

        // far back-references (older than the window) of <= 16 bytes: fetch the aligned words around the source from
        // global memory now -- every lane's loads are in flight together, one memory latency per group -- and let the
        // rounds read them from shared memory.  (Longer far copies are read byte-wise by the rounds.)
        const bool far = mine && is_copy && off > NEAR_MAX;
        const bool staged = far && len <= 16u;
        const uint32_t gsrc = wxd - off;  // window coordinate of the source
        // (the loads are issued here and stored below, behind the descriptor arithmetic: that much of their latency is free)
        const bool any_staged = __any_sync(SNP_FULL, staged);
        const uint32_t nw = staged ? ((gsrc & 3u) + len + 3u) >> 2 : 0u;
        uint32_t w[SNP7_STAGE_W];
        if (any_staged) {
            const uint32_t *gp = reinterpret_cast<const uint32_t *>(out16 + (gsrc & ~3u));
#pragma unroll
            for (uint32_t j = 0; j < SNP7_STAGE_W; j++)
                if (j < nw) w[j] = gp[j];
        }
        // descriptor of my tag for the rounds (t = wx + descriptor; only the low 16 bits of the sum are used):
        //   literal / staged far copy: shared-memory byte (t & 0xffff), a linear address inside Warp7;
        //   D_WIN: window byte t & (W-1) (low bits = -offset);  D_CLOSE: offset < 32, may read bytes of its own round;
        //   D_GLOBAL: global byte out16[wx - (descriptor & 0xffff)] (low bits = offset).
        const bool close = mine && is_copy && off < 32u;
        uint32_t pack;
        if (!is_copy) pack = (RING_OFF + ik + 1u - wxd) & 0xffffu;
        else if (!far) pack = D_WIN | (close ? D_CLOSE : 0u) | ((0u - off) & 0xffffu);
        else if (staged) pack = (STAGE_OFF + lane * (4u * SNP7_STAGE_W) + (gsrc & 3u) - wxd) & 0xffffu;
        else pack = D_GLOBAL | off;
        const bool any_global = __any_sync(SNP_FULL, far && !staged);
        // rounds (of 32 output bytes) that hold a byte of a close copy resolve same-round sources by pointer doubling
        uint32_t slow_rounds = 0;
        if (__any_sync(SNP_FULL, close)) {
            const uint32_t r0 = (dst - op) >> 5, r1 = (dst + len - 1u - op) >> 5;
            slow_rounds = __reduce_or_sync(SNP_FULL, close ? (2u << r1) - (1u << r0) : 0u);
        }
        if (any_staged) {
#pragma unroll
            for (uint32_t j = 0; j < SNP7_STAGE_W; j++)
                if (j < nw) s->stage[lane * SNP7_STAGE_W + j] = w[j];
        }
        __syncwarp();  // the staged words are visible to every lane

        // ---- COPY: rounds of 32 output bytes, one byte per lane.  rel = start of my tag relative to the round.
        uint32_t rel = mine ? dst - op : 0xffffffffu;
        uint32_t cm1 = 0xffffffffu;         // (tags that start in front of the round) - 1
        uint32_t wx = oskew + op + lane;    // my byte of the round, window coordinate
        const uint32_t wxe = oskew + op + glen;
        auto fast_round = [&](const bool with_global) {
            const uint32_t M = __reduce_or_sync(SNP_FULL, shl_clamp(1u, rel));
            rel -= 32u;
            const uint32_t d = __shfl_sync(SNP_FULL, pack, cm1 + __popc(M & le));
            cm1 += __popc(M);
            const uint32_t t = wx + d;
            const uint32_t si = t & ((int32_t)d < 0 ? (W - 1) : 0xffffu);
            if (with_global) {
                const bool glob = wx < wxe && (d & D_GLOBAL);
                uint32_t v = sb[(wx < wxe && !glob) ? si : (wx & (W - 1))];  // lanes behind the group: self-copy (below)
                if (glob) v = out16[wx - (d & 0xffffu)];
                s->win[wx & (W - 1)] = (uint8_t)v;
            } else {
                // branch-free: lanes behind the end of the group copy a (valid) window byte onto a window position that
                // the next group overwrites -- < 32 bytes ahead of op, inside the 64 bytes of slack NEAR_MAX leaves
                s->win[wx & (W - 1)] = sb[wx < wxe ? si : (wx & (W - 1))];
            }
            wx += 32u;
            __syncwarp();
        };
        auto slow_round = [&]() {
            const uint32_t M = __reduce_or_sync(SNP_FULL, shl_clamp(1u, rel));
            rel -= 32u;
            const uint32_t d = __shfl_sync(SNP_FULL, pack, cm1 + __popc(M & le));
            cm1 += __popc(M);
            const uint32_t t = wx + d;
            const bool active = wx < wxe;
            // source: sk 0 = shared-memory byte sa, 1 = global byte sa of out16, 2 = lane sa of this round
            uint32_t sk = 0, sa = t & ((int32_t)d < 0 ? (W - 1) : 0xffffu);
            if (d & D_GLOBAL) {
                sk = 1;
                sa = wx - (d & 0xffffu);
            }
            const uint32_t o5 = (0u - d) & 0xffffu;  // the offset of a close copy
            if ((d & D_CLOSE) && o5 <= lane) {        // its source byte is produced in this round, by lane - offset
                sk = 2;
                sa = lane - o5;
            }
            if (!active) sk = 0, sa = 0;
            while (__any_sync(SNP_FULL, sk == 2)) {  // pointer doubling, <= 5 trips (CopyHelpers.IncrementalCopy's
                const uint32_t nk = __shfl_sync(SNP_FULL, sk, sa);  // pattern replication, any offset)
                const uint32_t na = __shfl_sync(SNP_FULL, sa, sa);
                if (sk == 2) {
                    sk = nk;
                    sa = na;
                }
            }
            uint32_t v = sb[sk == 0 ? sa : 0u];
            if (sk == 1 && active) v = out16[sa];
            if (active) s->win[wx & (W - 1)] = (uint8_t)v;
            wx += 32u;
            __syncwarp();
        };
        const uint32_t nr = (glen + 31u) >> 5;
        if (slow_rounds == 0) {
            if (!any_global)
                for (uint32_t n = nr; n; n--) fast_round(false);
            else
                for (uint32_t n = nr; n; n--) fast_round(true);
        } else {
            for (uint32_t n = nr; n; n--, slow_rounds >>= 1) {
                if (slow_rounds & 1u) slow_round();
                else fast_round(true);
            }
        }
        op += glen;
        ip = ip_next;
        maybe_flush();
    }
    finish();
    if (status != SNP_OK) return status;
    flush_to(op);
    if (op < U) return SNP_INCOMPLETE;  // Snappy.cs:178-181
    *written = op;
    return SNP_OK;
}

#ifndef SNP_EMU
// Persistent launch: every warp pulls the next block index from a global counter (cheap and expensive blocks balance).
template <uint32_t W, int NW, int CTAS>  // window bytes per warp, warps per CTA, CTAs per SM
__global__ void __launch_bounds__(NW * 32, CTAS)
k_decompress_v7(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_len, uint8_t *out_base, const uint64_t *__restrict__ out_off,
                const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len, int32_t *__restrict__ status,
                size_t n_items, unsigned long long *__restrict__ next_item) {
    extern __shared__ __align__(16) uint8_t smem7[];
    const unsigned lane = lane_id();
    Warp7<W> *s = reinterpret_cast<Warp7<W> *>(smem7) + threadIdx.x / SNP_WARP;
    warp7_init(s, lane);
    mbar_fence_init();
    __syncthreads();
    uint32_t phases = 0;  // per-slot mbarrier phase parity, carried across this warp's blocks
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1ull);
        item = __shfl_sync(SNP_FULL, item, 0);
        if (item >= n_items) break;
        uint32_t w = 0;
        int st;
        if (in_len[item] >= 0x7fff0000u)  // stream offsets are 32-bit with headroom here; v1 is safe to 2^32-1
            st = decompress_block_v1(in_base + in_off[item], in_len[item], out_base + out_off[item], out_cap[item], &w);
        else
            st = decompress_block_v7<W>(in_base + in_off[item], in_len[item], out_base + out_off[item], out_cap[item],
                                        &w, s, phases);
        if (lane == 0) {
            out_len[item] = w;
            status[item] = st;
        }
        __syncwarp();
    }
}
#endif  // !SNP_EMU

}  // namespace snp
