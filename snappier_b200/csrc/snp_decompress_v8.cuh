// snp_decompress_v8.cuh -- lane-per-block batched Snappy decompressor (sm_100a): the THROUGHPUT engine for large batches.
//
// Why it exists.  The warp-per-block engines (v5, v7) are bound by warp-instruction issue: 3.9-4.3 warp instructions per
// output byte (profiles/r01_decompress_v5_ncu.md, profiles/r02_decompress_v7_ncu.md), because 32 lanes cooperate on a
// format whose unit of work -- one tag, 4-12 bytes of output on text -- is far smaller than a warp.  A batch of 2^20
// independent blocks has all the parallelism the chip needs ACROSS blocks, so this engine gives every LANE its own block
// and runs the reference's scalar tag loop (SnappyDecompressor.cs:234-341) as a branch-free-ish state machine, 32 blocks
// per warp in lock step.  One warp iteration = (decode a tag, if the lane needs one) + (move up to 8 bytes of the lane's
// current literal / copy): ~100 warp instructions for ~32 x 6 output bytes, an order of magnitude fewer instructions per
// byte than one byte per lane per round.
//
// What makes lane-per-block workable on a GPU (memory side):
//   * INPUT: each lane owns a 64-byte shared-memory ring filled by 16-byte cp.async (LDGSTS) copies straight from the
//     compressed stream, one chunk per iteration, one iteration ahead of its use (cp.async.wait_group 1).  Tag bytes and
//     literal bytes are unaligned reads of that ring (2-3 LDS.32 + funnel shifts).
//   * OUTPUT: each lane owns a 256-byte shared-memory ring indexed by the low bits of the global output address.
//     Produced bytes are merged into aligned words (a one-word register accumulator + funnel shifts) and stored to the
//     ring; every completed 16-byte chunk leaves as ONE aligned LDS.128 + STG.128.  No byte stores to HBM.
//   * BACK-REFERENCES up to 240 bytes are unaligned reads of the lane's own output ring; older ones are one aligned
//     LDG.128 (+ a predicated LDG.64) of the lane's own flushed output (L2 / HBM) -- a 32-byte sector per copy tag, never
//     a byte load.  Overlapping copies (offset < 8) use CopyHelpers.IncrementalCopy's pattern doubling
//     (CopyHelpers.cs:76-160): `offset` doubles each time a whole period has been appended.
//   * LONG LITERALS (>= 256 bytes, i.e. incompressible data) are handed to the whole warp: a ballot, then all 32 lanes
//     copy that lane's literal input -> output as coalesced 16-byte vectors.
//   * The per-lane rings are contiguous with a 16-byte skew between lanes (stride 336 B = 84 words = 20 mod 32 banks),
//     so a quarter-warp's LDS.128 / cp.async land in distinct bank groups.
//
// Scattered (one line per lane) global accesses cost ~1-2 L1 wavefront cycles per lane, so they are kept to one 16-byte
// access per 16 bytes moved (refill, flush) and one per far back-reference.
//
// Scheduling: persistent CTAs; a lane that finishes its block takes the next item from a global counter (one atomicAdd
// per warp and refill event).  Blocks whose header announces more than 1 MiB are decoded by the whole warp with
// decompress_block_v1 instead (a lone lane would take forever on them).
//
// Semantics: /root/reference/Snappier/Internal/SnappyDecompressor.cs:43-92,184-347,556-611 (one-shot); identical results
// (status, length, bytes) to v1 and oracle/snappy_oracle.c.  The lane function also compiles against
// tests/cpp/simt_emu.h (tests/test_emu_v8.py).
#pragma once
#include "snp_common.cuh"
#include "snp_decompress_v1.cuh"
#include "snp_decompress_v7.cuh"  // copy_literal_wide7

namespace snp {

#ifndef SNP8_STAT
#define SNP8_STAT(what)  // tests/cpp/emu_v8.cpp counts iterations / bubbles with this hook
#endif

template <uint32_t IR, uint32_t ORB, int D>
struct alignas(16) Lane8 {
    uint8_t o[ORB];      // output ring: window coordinate x (= (out address & 15) + output position) lives at x mod ORB
    uint8_t i[IR];       // input ring: ring coordinate q (= (in address & 15) + stream position) lives at q mod IR
    uint8_t far[D][32];  // the two aligned 16-byte chunks around the source of each far back-reference in flight
    uint8_t pad[16];     // bank skew between lanes
};

#ifdef SNP_EMU
__device__ __forceinline__ void cp_async16_v8(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
    memset(smem_dst, 0, 16);
    memcpy(smem_dst, gsrc, src_bytes);
}
__device__ __forceinline__ void cp_async_commit_v8() {}
template <int N> __device__ __forceinline__ void cp_async_wait_v8() {}
#else
// 16 bytes global -> shared, asynchronous; bytes past src_bytes are zero-filled (never read from memory)
__device__ __forceinline__ void cp_async16_v8(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit_v8() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_v8() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

// Descriptor of one pipeline step: bits 0..3 = bytes to move (0 = nothing), bits 4..11 = byte index of the source inside
// its region, bits 16..27 = byte offset of the region inside Lane8, bit 31 = the region is a 32-byte far buffer (no wrap).
#define SNP8_FARBIT 0x80000000u

template <uint32_t IR, uint32_t ORB, int D>
__device__ __forceinline__ void decompress_lanes_v8(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                                                    const uint32_t *__restrict__ in_len, uint8_t *out_base,
                                                    const uint64_t *__restrict__ out_off,
                                                    const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len,
                                                    int32_t *__restrict__ status, size_t n_items,
                                                    unsigned long long *next_item, Lane8<IR, ORB, D> *me) {
    static_assert(IR == ORB && (IR & (IR - 1)) == 0 && IR >= 128 && IR <= 256, "ring sizes (one wrap mask for both)");
    static_assert(D >= 1 && D <= 4 && ORB >= 8u * D + 56u && IR >= 13u * D + 48u, "pipeline depth against the rings");
    constexpr uint32_t Q = 8;               // bytes per step
    constexpr uint32_t NEAR = ORB - 16;     // back-references up to this distance are read from the output ring
    constexpr uint32_t BULK_MIN = 256;      // literals at least this long (from a 16-byte output boundary) go to the warp
    constexpr uint32_t BIG_BLOCK = 1u << 20;
    constexpr uint32_t RM = ORB - 1;        // ring wrap mask (bytes)
    constexpr uint32_t IOFF = ORB, FOFF = ORB + IR;  // byte offsets of the input ring / far buffers inside Lane8
    enum : uint32_t { RUN = 0, DRAIN_END = 1, DRAIN_BULK = 2, BULK_READY = 3 };
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    uint8_t *const lb = reinterpret_cast<uint8_t *>(me);
    auto ldw = [&](uint32_t byte_off) -> uint32_t { return *reinterpret_cast<const uint32_t *>(lb + byte_off); };

    bool have = false;  // this lane has a block in progress
    bool more = true;   // the work counter may still hold items (warp-uniform)
    const uint8_t *in16 = nullptr;  // 16-byte-aligned views of the lane's streams
    uint8_t *out16 = nullptr;
    // PARSE side (runs D steps ahead): ring coordinate of the next input byte, window coordinate the next step writes to
    uint32_t pip = 0, iend = 0, pop = 0, oend = 0, obeg = 0;
    uint32_t prem = 0, poff = 0;       // bytes left of the current tag; its offset (0 = literal)
    uint32_t rfrom = 0;                // the output ring holds window bytes [max(rfrom, op - NEAR), op)
    uint32_t mode = RUN, drain = 0;
    uint32_t issued = 0, safe = 0;     // input chunks requested up to `issued`; bytes below `safe` have landed
    // EXECUTE side: window coordinate of the next output byte; bytes below f (a multiple of 16) are in global memory
    uint32_t op = 0, f = 0;
    uint32_t acc = 0;                  // the (op & 3) valid low bytes of the output word under construction
    // the pipeline: step descriptors; `issued` as it was when the copies of D iterations ago were committed
    uint32_t desc[D], sh[D];
#pragma unroll
    for (int k = 0; k < D; k++) desc[k] = 0, sh[k] = 0;
    size_t item = 0;

    auto drain_async = [&]() {
        cp_async_commit_v8();
        cp_async_wait_v8<0>();
    };
    auto result = [&](int st, uint32_t w) {
        out_len[item] = w;
        status[item] = st;
        have = false;
    };
    // the block failed at parse time: nothing more of it is executed (the output of a failed block is unspecified)
    auto fail = [&](int st) {
#pragma unroll
        for (int k = 0; k < D; k++) desc[k] = 0;
        result(st, 0);
    };
    // the tag stream ended and every step has been executed
    auto finalize = [&]() {
        if (op < oend) {  // Snappy.cs:178-181
            result(SNP_INCOMPLETE, 0);
            return;
        }
        for (uint32_t x = max(f, obeg); x < op; x++) out16[x] = me->o[x & RM];  // < 16 bytes (+ a short head)
        result(SNP_OK, op - obeg);
    };
    // Reads the header of `item`; true = the whole warp has to decode it (big block)
    auto init = [&]() -> bool {
        const uint8_t *in = in_base + in_off[item];
        const uint32_t n_in = in_len[item];
        uint8_t *out = out_base + out_off[item];
        uint32_t U, used;
        const int st = varint_read(in, n_in, &U, &used);
        if (n_in >= 0x7fff0000u || (st == SNP_OK && U > BIG_BLOCK && U <= 0x7fffffffu)) return true;
        if (st == SNP_INCOMPLETE) {  // SnappyDecompressor.cs:57-60
            result(SNP_INCOMPLETE, 0);
            return false;
        }
        if (st != SNP_OK || U > 0x7fffffffu) {
            result(SNP_INVALID_LENGTH, 0);
            return false;
        }
        if (out_cap[item] < U) {
            result(SNP_OUTPUT_TOO_SMALL, 0);
            return false;
        }
        if (U == 0) {  // AllDataDecompressed before any tag (SnappyDecompressor.cs:78)
            result(SNP_OK, 0);
            return false;
        }
        const uint32_t skew = (uint32_t)((uintptr_t)in & 15);
        in16 = in - skew;
        iend = skew + n_in;
        pip = skew + used;
        obeg = (uint32_t)((uintptr_t)out & 15);
        out16 = out - obeg;
        op = pop = obeg;
        oend = obeg + U;
        f = 0;
        rfrom = obeg;
        prem = 0;
        poff = 0;
        acc = 0;
        mode = RUN;
        // prime the input ring and wait for it (once per block); copies of the lane's previous block have all landed
        drain_async();
        issued = pip & ~15u;
        const uint32_t a16 = (iend + 15u) & ~15u;
        while (issued < a16 && issued - (pip & ~15u) < IR) {
            cp_async16_v8(me->i + (issued & (IR - 1)), in16 + issued, min(16u, iend - issued));
            issued += 16;
        }
        drain_async();
        safe = issued;
#pragma unroll
        for (int k = 0; k < D; k++) desc[k] = 0, sh[k] = issued;
        have = true;
        return false;
    };

    for (;;) {
#pragma unroll
        for (int k = 0; k < D; k++) {  // pipeline slot k: executes the step parsed D iterations ago, then parses a new one
            // ---- (0) rare, warp-wide: lanes without a block take the next items; long literals are copied by the warp --
            if (__any_sync(SNP_FULL, !have || mode == BULK_READY)) {
                const unsigned need = __ballot_sync(SNP_FULL, !have);
                bool big = false;
                if (need && more) {
                    const uint32_t cnt = (uint32_t)__popc(need);
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(next_item, (unsigned long long)cnt);
                    base = __shfl_sync(SNP_FULL, base, 0);
                    if (!have) {
                        const unsigned long long it = base + (unsigned)__popc(need & lt);
                        if (it < n_items) {
                            item = (size_t)it;
                            big = init();
                        }
                    }
                    if (base + cnt >= n_items) more = false;
                }
                unsigned bigm = __ballot_sync(SNP_FULL, big);
                while (bigm) {  // big blocks: all 32 lanes, v1
                    const int l = __ffs(bigm) - 1;
                    bigm &= bigm - 1;
                    const unsigned long long it = __shfl_sync(SNP_FULL, (unsigned long long)item, l);
                    uint32_t w = 0;
                    const int st = decompress_block_v1(in_base + in_off[it], in_len[it], out_base + out_off[it], out_cap[it], &w);
                    if (lane == 0) {
                        out_len[it] = w;
                        status[it] = st;
                    }
                    __syncwarp();
                }
                if (!more && __ballot_sync(SNP_FULL, have) == 0) return;
                // long literals (the lane's pipeline is empty): input -> output as coalesced 16-byte vectors
                unsigned bulkm = __ballot_sync(SNP_FULL, have && mode == BULK_READY);
                while (bulkm) {
                    const int l = __ffs(bulkm) - 1;
                    bulkm &= bulkm - 1;
                    const unsigned long long s64 = __shfl_sync(SNP_FULL, (unsigned long long)(uintptr_t)(in16 + pip), l);
                    const unsigned long long e64 = __shfl_sync(SNP_FULL, (unsigned long long)(uintptr_t)(in16 + iend), l);
                    const unsigned long long d64 = __shfl_sync(SNP_FULL, (unsigned long long)(uintptr_t)(out16 + pop), l);
                    const uint32_t blen = __shfl_sync(SNP_FULL, prem, l) & ~15u;
                    copy_literal_wide7((const uint8_t *)(uintptr_t)s64, (uint8_t *)(uintptr_t)d64, blen,
                                       (const uint8_t *)(uintptr_t)e64, lane);
                    __syncwarp();  // the owner's later back-reference loads see the other lanes' stores
                    if ((int)lane == l) {
                        drain_async();
                        pip += blen;
                        pop += blen;
                        prem -= blen;
                        op = pop;
                        f = op;       // everything below op is in global memory,
                        rfrom = op;   // none of it in the ring
                        acc = 0;
                        issued = pip & ~15u;
                        safe = issued;
#pragma unroll
                        for (int j = 0; j < D; j++) sh[j] = issued;
                        mode = RUN;
                    }
                }
            }
            SNP8_STAT(0);

            // ---- (1) the asynchronous copies committed D iterations ago (input chunk, far source of desc[k]) have landed
            cp_async_wait_v8<D - 1>();
            safe = sh[k];

            // ---- (2) EXECUTE the step parsed D iterations ago: up to 8 source bytes -> output ring ----------------------
            {
                const uint32_t d = desc[k];
                const uint32_t n = d & 15u;
                if (n) {
                    SNP8_STAT(1);
                    const uint32_t idx = (d >> 4) & 0xffu;
                    const uint32_t wm = (d & SNP8_FARBIT) ? 31u : RM;
                    const uint32_t rb = (d >> 16) & 0xfffu;
                    const uint32_t i0 = idx & ~3u;
                    const uint32_t x0 = ldw(rb + (i0 & wm)), x1 = ldw(rb + ((i0 + 4u) & wm)), x2 = ldw(rb + ((i0 + 8u) & wm));
                    const uint32_t s = (idx & 3u) * 8u;
                    const uint32_t v0 = __funnelshift_r(x0, x1, s), v1 = __funnelshift_r(x1, x2, s);
                    const uint32_t t = (op & 3u) * 8u;
                    const uint32_t w0 = acc | (v0 << t);
                    const uint32_t w1 = __funnelshift_l(v0, v1, t);
                    const uint32_t w2 = __funnelshift_l(v1, 0u, t);
                    const uint32_t j0 = op & ~3u;
                    *reinterpret_cast<uint32_t *>(lb + (j0 & RM)) = w0;         // bytes past op + n are garbage that later
                    *reinterpret_cast<uint32_t *>(lb + ((j0 + 4u) & RM)) = w1;  // steps overwrite; they alias window
                    *reinterpret_cast<uint32_t *>(lb + ((j0 + 8u) & RM)) = w2;  // positions older than NEAR reaches
                    const uint32_t tot = (op & 3u) + n;  // 1..11
                    const uint32_t keep = (1u << (8u * (tot & 3u))) - 1u;
                    acc = (tot < 4u ? w0 : tot < 8u ? w1 : w2) & keep;
                    op += n;
                }
                // a completed 16-byte chunk of the output ring goes to global memory
                if (have && (op & ~15u) > f) {
                    if (f >= obeg) {
                        *reinterpret_cast<uint4 *>(out16 + f) = *reinterpret_cast<const uint4 *>(me->o + (f & RM));
                    } else {  // the first chunk of a block whose output does not start on a 16-byte boundary
                        for (uint32_t x = obeg; x < 16u; x++) out16[x] = me->o[x];
                    }
                    f += 16;
                }
            }

            // ---- (3) PARSE the next step (SnappyDecompressor.cs:234-341; Constants.cs:42-76's table as arithmetic) -------
            uint32_t nd = 0;
            if (have) {
                if (mode != RUN) {
                    if (mode != BULK_READY) {
                        if (drain == 0) {
                            if (mode == DRAIN_END) finalize();
                            else mode = BULK_READY;
                        } else {
                            drain--;
                        }
                    }
                } else {
                    const uint32_t lim = min(safe, iend);
                    if (prem == 0) {
                        if (pip >= iend) {
                            mode = DRAIN_END;  // the stream is over: let the pipeline run empty, then finalize
                            drain = D - 1;
                        } else if ((int)(lim - pip) >= 5 || (lim == iend && lim > pip)) {
                            const uint32_t a0 = pip & ~3u;
                            const uint32_t t = __funnelshift_r(ldw(IOFF + (a0 & RM)), ldw(IOFF + ((a0 + 4u) & RM)), (pip & 3u) * 8u);
                            const uint32_t c = t & 0xffu, kind = c & 3u, n6 = c >> 2;
                            if (kind == 3u || (kind == 0 && n6 >= 60u)) {  // rare forms: COPY4, literal length in 1..4 trailer bytes
                                const uint32_t extra = kind == 3u ? 4u : n6 - 59u;
                                if (iend - pip < 1u + extra) {
                                    mode = DRAIN_END;  // truncated tag: the stream ends here (RefillTag, :464-483)
                                    drain = D - 1;
                                } else {
                                    const uint32_t a1 = (pip + 1u) & ~3u;
                                    const uint32_t t4 = __funnelshift_r(ldw(IOFF + (a1 & RM)), ldw(IOFF + ((a1 + 4u) & RM)),
                                                                        ((pip + 1u) & 3u) * 8u);
                                    const uint32_t trailer = extra == 4u ? t4 : (t4 & ~(0xffffffffu << (8u * extra)));
                                    pip += 1u + extra;
                                    if (kind == 3u) {
                                        poff = trailer;
                                        prem = n6 + 1u;
                                    } else {
                                        const uint32_t avail = iend - pip;
                                        prem = trailer >= avail ? avail : trailer + 1u;  // partial literal, :290-297
                                        poff = 0;
                                    }
                                }
                            } else if (iend - pip < 1u + kind) {
                                mode = DRAIN_END;  // truncated copy tag
                                drain = D - 1;
                            } else {  // literal of 1..60 bytes, COPY1, COPY2: tag bytes = kind + 1
                                pip += 1u + kind;
                                const uint32_t b1 = (t >> 8) & 0xffu;
                                if (kind == 0) {
                                    const uint32_t avail = iend - pip;
                                    prem = n6 >= avail ? avail : n6 + 1u;  // partial literal, :290-297
                                    poff = 0;
                                } else if (kind == 1u) {
                                    prem = (n6 & 7u) + 4u;
                                    poff = ((c >> 5) << 8) | b1;
                                } else {
                                    prem = n6 + 1u;
                                    poff = (t >> 8) & 0xffffu;
                                }
                            }
                            if (mode == RUN) {  // validation in stream order
                                if (kind != 0 && (poff == 0 || pop - obeg < poff)) fail(SNP_INVALID_COPY_OFFSET);  // :598-601
                                else if (prem > oend - pop) fail(SNP_DATA_TOO_LONG);                              // :570-573,603-606
                            }
                        }
                    }
                    if (have && mode == RUN && prem != 0) {
                        uint32_t n = min(prem, Q);
                        if (poff == 0) {
                            if (prem >= BULK_MIN && (pop & 15u) == 0) {
                                mode = DRAIN_BULK;
                                drain = D - 1;
                            } else {
                                n = min(n, lim > pip ? lim - pip : 0u);
                                nd = n ? ((IOFF << 16) | ((pip & RM) << 4) | n) : 0u;
                                pip += n;
                                pop += n;
                                prem -= n;
                            }
                        } else {
                            const uint32_t p = pop - poff;  // window coordinate of the source
                            if (poff < Q) n = min(n, poff);
                            if (poff <= NEAR && p >= rfrom) {
                                nd = ((p & RM) << 4) | n;
                            } else {  // the source is in global memory already (older than anything in flight)
                                if (p < rfrom) n = min(n, rfrom - p);  // the rest of the source sits in the ring: next step
                                const uint8_t *g = out16 + (p & ~15u);
                                cp_async16_v8(me->far[k], g, 16u);
                                if ((p & 15u) + n > 16u) cp_async16_v8(me->far[k] + 16, g + 16, 16u);
                                nd = SNP8_FARBIT | ((FOFF + 32u * k) << 16) | ((p & 15u) << 4) | n;
                            }
                            if (poff < Q && n == poff) poff <<= 1;  // a whole period appended: the pattern is twice as long
                            pop += n;
                            prem -= n;
                        }
                    }
                    // input ring: request the next chunk (bytes of literal steps still in flight stay: <= 13 per step)
                    if (have) {
                        const int keep_from = ((int)pip - 13 * D) & ~15;
                        if (issued < ((iend + 15u) & ~15u) && (int)issued + 16 - keep_from <= (int)IR) {
                            cp_async16_v8(me->i + (issued & (IR - 1)), in16 + issued, min(16u, iend - issued));
                            issued += 16;
                        }
                    }
                }
            }
            desc[k] = nd;
            cp_async_commit_v8();
            sh[k] = issued;
        }
    }
}

#ifndef SNP_EMU
template <uint32_t IR, uint32_t ORB, int D, int NT, int CTAS>  // ring bytes per lane, pipeline depth, threads per CTA, CTAs per SM
__global__ void __launch_bounds__(NT, CTAS)
k_decompress_v8(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_len, uint8_t *out_base, const uint64_t *__restrict__ out_off,
                const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len, int32_t *__restrict__ status,
                size_t n_items, unsigned long long *__restrict__ next_item) {
    extern __shared__ __align__(16) uint8_t smem8[];
    Lane8<IR, ORB, D> *me = reinterpret_cast<Lane8<IR, ORB, D> *>(smem8) + threadIdx.x;
    decompress_lanes_v8<IR, ORB, D>(in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n_items,
                                 next_item, me);
}
#endif  // !SNP_EMU

}  // namespace snp
