// snp_decompress_v6.cuh -- checkpointed two-pass batched Snappy block decompressor (sm_100a).
//
// ncu on v3/v5 (profiles/r01_decompress_v5_ncu.md) shows the one-warp-per-block kernels bound
// by warp-instruction issue: finding the tag a byte belongs to costs ~155 instructions per 32
// input bytes (speculative parse) plus ~87 per 32 output bytes (byte-per-lane rounds).  v6 removes
// both costs by giving every LANE a whole tag:
//
//   pass A  k_tagscan_v6   one THREAD per block walks the tag chain (SnappyDecompressor.cs:184-347
//           is a serial dependency chain, so a thread is the natural unit), validates every tag in
//           stream order (the block's final status is decided here) and drops a checkpoint
//           (ip, op) every 32 tag slots.  Input reaches each lane through a 32-byte register window
//           (16-byte vector loads, one vector ahead of the cursor).
//   pass B  k_decode_v6    one WARP per block (the north star's decomposition).  Per super-step
//           lane l re-parses the 32 slots behind checkpoint 32*j+l into 4-byte records in shared
//           memory (1024 tags in flight), then the warp executes the groups in stream order, ONE
//           TAG PER LANE: output offsets by a shuffle scan of the lengths, 16-byte vector loads of
//           the source (input stream, shared-memory output window, or the block's older output in
//           global memory), a register funnel to the byte offset, byte stores into a sliding
//           shared-memory OUTPUT WINDOW, which is flushed to HBM as aligned 16-byte vectors.  Tags
//           whose source is produced inside the same group wait for the frontier (multi-round
//           resolution; the first pending tag is always runnable).  Overlapping copies with
//           offset < 16 (CopyHelpers.IncrementalCopy's pattern replication) and literals > 64 bytes
//           take warp-cooperative paths.
//
// Slots: a tag is one slot {len (7 bits) | literal (1) | copy offset or literal input offset (24)};
// a literal > 64 bytes is a head slot (len 0) plus a slot holding its length << 8, never split
// across two groups (an empty pad slot is inserted when it would be).
//
// Semantics: /root/reference/Snappier/Internal/SnappyDecompressor.cs:43-92,184-347,556-611,
// identical to v1/v3/v5 and oracle/snappy_oracle.c (status precedence = stream order).
//
// The warp-level functions compile for the host-side SIMT emulator too (-DSNP_EMU, see
// tests/cpp/simt_emu.h): with no GPU in the development container that is how the parse /
// dependency / window logic is checked before a B200 run.
#pragma once
#include "snp_common.cuh"
#ifndef SNP_EMU
#include "snp_decompress_v5.cuh"  // fallback engines for blocks outside v6's envelope
#endif

namespace snp {

#define SNP6_T 32u  // slots per checkpoint group (= lanes)
#ifndef SNP6_CKB
#define SNP6_CKB 768u  // checkpoint budget per block: 24 576 slots; denser blocks fall back to v3
#endif
#define SNP6_NT_FALLBACK 0xffffffffu
#define SNP6_KEEP 768u                          // bytes of history the output window keeps across a slide
#define SNP6_SPAN 2048u                         // 32 tags x 64 bytes: most a sub-group can produce
#define SNP6_WIN (SNP6_KEEP + SNP6_SPAN + 64u)  // window bytes
#define SNP6_FLUSH 512u                         // flush the window when this many bytes are pending
#define SNP6_RLIT 0x80u
#define SNP6_VMAX 0x1000000u  // record values (copy offsets, literal input offsets) are 24-bit

struct alignas(16) V6Smem {  // pass B, per warp
    uint32_t rec[32 * 33];   // [group lane][slot] records
    uint32_t gop[32];        // output offset at the start of each group
    uint8_t win[SNP6_WIN + 16];
    uint4 pslot[32 * 3];     // parse: per-lane 32-byte input ring (48-byte pitch)
};

#ifdef SNP_EMU
struct V6Stats {
    unsigned long groups, subgroups, rounds, trips, ctags, huge, slides, tags, flushes, fast, hops, stuck_kind, stuck_straddle;
};
inline V6Stats &v6_stats() {
    static V6Stats s{};
    return s;
}
#define SNP6_STAT(f, n) do { if (lane_id() == 0) v6_stats().f += (n); } while (0)
#else
#define SNP6_STAT(f, n) do { } while (0)
#endif

// ---- explicit address-space accesses (the block functions are not inlined into the kernels, so
//      plain pointers would compile to generic LD/ST) ------------------------------------------------
#ifdef SNP_EMU
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4 *p) { return *p; }
__device__ __forceinline__ uint4 ldg_v4(const uint4 *p) { return *p; }
__device__ __forceinline__ uint4 ld_any_v4(const uint4 *p) { return *p; }
__device__ __forceinline__ void stg_v4(uint4 *p, uint4 v) { *p = v; }
// stores the low min(m, 4) bytes of r at p
__device__ __forceinline__ void st_win_bytes4(uint8_t *p, uint32_t r, uint32_t m) {
    for (uint32_t j = 0; j < 4 && j < m; j++) p[j] = (uint8_t)(r >> (8 * j));
}
#else
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint4 ldg_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
// generic load (shared or global address), ordered against the surrounding window / output accesses
__device__ __forceinline__ uint4 ld_any_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg_v4(uint4 *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_win_bytes4(uint8_t *p, uint32_t r, uint32_t m) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile(
        "{\n\t.reg .pred p0,p1,p2,p3;\n\t.reg .b32 t1,t2,t3;\n\t"
        "setp.gt.u32 p0,%2,0;\n\tsetp.gt.u32 p1,%2,1;\n\tsetp.gt.u32 p2,%2,2;\n\tsetp.gt.u32 p3,%2,3;\n\t"
        "shr.u32 t1,%1,8;\n\tshr.u32 t2,%1,16;\n\tshr.u32 t3,%1,24;\n\t"
        "@p0 st.shared.u8 [%0],%1;\n\t@p1 st.shared.u8 [%0+1],t1;\n\t"
        "@p2 st.shared.u8 [%0+2],t2;\n\t@p3 st.shared.u8 [%0+3],t3;\n\t}" ::"r"(sa), "r"(r), "r"(m) : "memory");
}
#endif

struct Tag6 {
    uint32_t hdr, take, off;
    bool is_lit, end, partial;
};

// Decodes the tag whose first four bytes are v and fifth byte is the low byte of b4;
// `left` = bytes from the tag byte to the end of the input (>= 1).
__device__ __forceinline__ Tag6 decode_tag6(uint32_t v, uint32_t b4, uint32_t ent, uint32_t left) {
    Tag6 t;
    const uint32_t trailer = __funnelshift_r(v, b4, 8);  // bytes 1..4
    t.hdr = __byte_perm(ent, 0, 0x4441);
    t.is_lit = (int32_t)ent < 0;
    const uint32_t tval = trailer & __funnelshift_rc(0xffffffffu, 0u, __byte_perm(ent, 0, 0x4442));
    uint32_t len = ent & 0xff;
    if (ent & 0x40000000u) len = max(tval + 1, tval);  // trailer-length literal, saturating
    t.off = ((ent >> 16) & 0x700u) | tval;              // copies only
    t.end = left < t.hdr;                                // truncated tag: parsing stops (:464-483)
    const uint32_t avail = left - t.hdr;
    t.partial = t.is_lit && !t.end && len > avail;  // literal cut by the end of the input (:290-297)
    t.take = t.partial ? avail : len;
    return t;
}

// ================================================================ pass A: tag scan ==========

struct Scan6Args {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint32_t *in_len;
    const uint32_t *out_cap;
    uint32_t *out_len;
    int32_t *status;
    size_t first_item, n_items;  // this wave: items [first_item, first_item + n_items)
    unsigned long long *next_item;
    uint32_t *ntags;  // per wave slot: slot count, or SNP6_NT_FALLBACK
    uint2 *ck;        // per wave slot: SNP6_CKB checkpoints (ip, op)
};

// Per-lane sequential reader of the lane's own input stream.  A 32-byte shared-memory ring holds the
// 16-byte vectors vb and vb+1 (vector v in ring[v & 1]); vector vb+2 is already in flight in registers
// (f), so walking forward never waits for memory, and the tag bytes are read with two LDS at a
// dynamic word index instead of register selects.
struct Stream6 {
    const uint4 *base;  // 16-byte aligned; stream byte p sits at byte (skew + p)
    uint32_t last_v;    // last vector index that may be read
    uint32_t vb;        // ring holds vectors vb, vb + 1
    uint4 *ring;        // this lane's two vectors in shared memory
    uint4 f;            // vector vb + 2
    __device__ __forceinline__ void open(const uint4 *b, uint32_t lv, uint4 *r) {
        base = b;
        last_v = lv;
        ring = r;
        vb = 0xfffffff0u;
    }
    // words holding stream bytes pos .. pos + 7 (pos includes the skew)
    __device__ __forceinline__ void words(uint32_t pos, uint32_t &w0, uint32_t &w1) {
        const uint32_t vq = pos >> 4;
        if (vq != vb) {
            if (vq == vb + 1) {
                ring[(vq + 1) & 1] = f;
            } else {  // first use, or a literal skipped ahead
                ring[vq & 1] = ldg_nc_v4(base + min(vq, last_v));
                ring[(vq + 1) & 1] = ldg_nc_v4(base + min(vq + 1, last_v));
            }
            f = ldg_nc_v4(base + min(vq + 2, last_v));
            vb = vq;
        }
        const volatile uint32_t *rw = (const volatile uint32_t *)ring;
        const uint32_t wi = pos >> 2;
        w0 = rw[wi & 7];
        w1 = rw[(wi + 1) & 7];
    }
};

// The persistent scan: every THREAD owns one block at a time and walks its tag chain.  The inner
// loop is one tag per lane per trip; it is left only when some lane finishes its block, so that the
// lanes of a warp fetch new blocks without waiting for each other's (long) blocks.
__device__ __forceinline__ void tagscan_warp_v6(const Scan6Args &a, const uint32_t *lut, uint4 *ring) {
    bool have = false, exhausted = false;
    size_t slot_idx = 0, item = 0;
    uint32_t ski = 0, n_in = 0, U = 0, ip = 0, op = 0, nslot = 0;
    Stream6 s;
    s.open(nullptr, 0, ring);
    for (;;) {
        while (!have && !exhausted) {
            const unsigned long long it = atomicAdd(a.next_item, 1ull);
            if (it >= a.n_items) {
                exhausted = true;
                break;
            }
            slot_idx = (size_t)it;
            item = a.first_item + slot_idx;
            const uint8_t *in = a.in_base + a.in_off[item];
            n_in = a.in_len[item];
            const uint32_t cap = a.out_cap[item];
            if (n_in >= SNP6_VMAX) {  // records hold 24-bit input offsets: v3 / v1 walker in pass B
                a.ntags[slot_idx] = SNP6_NT_FALLBACK;
                continue;
            }
            uint32_t used;
            int st = varint_read(in, n_in, &U, &used);  // SnappyDecompressor.cs:50-63
            if (st == SNP_OK && U > 0x7fffffffu) st = SNP_INVALID_LENGTH;
            if (st == SNP_OK && cap < U) st = SNP_OUTPUT_TOO_SMALL;
            if (st != SNP_OK || U == 0) {
                a.out_len[item] = 0;
                a.status[item] = st;
                a.ntags[slot_idx] = 0;
                continue;
            }
            ski = (uint32_t)((uintptr_t)in & 15);
            s.open((const uint4 *)((uintptr_t)in - ski), (ski + n_in - 1) >> 4, ring);
            ip = used;
            op = 0;
            nslot = 0;
            have = true;
        }
        if (__all_sync(SNP_FULL, !have)) break;
        bool fin = false;
        do {
            if (have) {
                int st = SNP_OK;
                bool done = false, fb = false;
                if (ip >= n_in) {
                    done = true;
                } else {
                    const uint32_t pos = ski + ip;
                    uint32_t w0, w1;
                    s.words(pos, w0, w1);
                    const unsigned sh = (pos & 3) * 8;
                    const uint32_t v = __funnelshift_r(w0, w1, sh);
                    const Tag6 t = decode_tag6(v, w1 >> sh, lut[v & 0xff], n_in - ip);
                    if (t.end) {
                        done = true;
                    } else if (!t.is_lit && t.off - 1u >= op) {  // off == 0 || off > produced (:598-601)
                        st = SNP_INVALID_COPY_OFFSET;
                        done = true;
                    } else if (t.take > U - op) {  // :570-573, :603-606
                        st = SNP_DATA_TOO_LONG;
                        done = true;
                    } else {
                        if (t.take) {
                            const bool huge = t.take > 64u;  // literals only
                            if (huge && (nslot & (SNP6_T - 1)) == SNP6_T - 1) nslot++;  // pad: head + length stay together
                            const uint32_t g = nslot / SNP6_T;
                            // outside the record format (32-bit copy offset) / denser than the checkpoint
                            // budget: the v3 engine decodes this block in pass B
                            fb = (!t.is_lit && t.off >= SNP6_VMAX) || g >= SNP6_CKB;
                            if (!fb && (nslot & (SNP6_T - 1)) == 0) a.ck[slot_idx * SNP6_CKB + g] = make_uint2(ip, op);
                            nslot += huge ? 2u : 1u;
                        }
                        op += t.take;
                        ip += t.hdr + (t.is_lit ? t.take : 0u);
                        done = t.partial;
                    }
                }
                if (fb) {
                    a.ntags[slot_idx] = SNP6_NT_FALLBACK;
                    have = false;
                    fin = true;
                } else if (done) {  // the stream ended or failed: this is the block's result
                    if (st == SNP_OK && op < U) st = SNP_INCOMPLETE;  // Snappy.cs:178-181
                    a.out_len[item] = st == SNP_OK ? op : 0u;
                    a.status[item] = st;
                    a.ntags[slot_idx] = st == SNP_OK ? nslot : 0u;
                    have = false;
                    fin = true;
                }
            }
        } while (!__any_sync(SNP_FULL, fin));
    }
}

// ================================================================ pass B: decode ============

// Cooperative global -> global copy of a long literal; >= 128 bytes go as aligned 16-byte vectors
// (source realigned with a warp-uniform funnel shift).  in_end bounds the vector reads.
__device__ __forceinline__ void copy_wide_v6(const uint8_t *s, uint8_t *d, uint32_t len, const uint8_t *in_end,
                                             unsigned lane) {
    if (len < 128) {
        for (uint32_t k = lane; k < len; k += SNP_WARP) d[k] = s[k];
        return;
    }
    const uint32_t head = (uint32_t)(-(intptr_t)d) & 15u;
    if (lane < head) d[lane] = s[lane];
    const uint8_t *sv = s + head;
    uint4 *dv = (uint4 *)(d + head);
    const uint32_t nvec = (len - head) >> 4;
    const unsigned sb = (unsigned)((uintptr_t)sv & 15);
    const uint4 *base = (const uint4 *)(sv - sb);
    const uint4 *last = (const uint4 *)(((uintptr_t)in_end - 1) & ~(uintptr_t)15);
    const unsigned ws = sb >> 2, bs = (sb & 3) * 8;
    for (uint32_t v = lane; v < nvec; v += SNP_WARP) {
        const uint4 A = ldg_nc_v4(base + v);
        const uint4 *pb = base + v + 1;
        const uint4 B = ldg_nc_v4(pb <= last ? pb : last);
        uint32_t x0 = A.x, x1 = A.y, x2 = A.z, x3 = A.w, x4 = B.x, x5 = B.y, x6 = B.z;
        if (ws & 2) x0 = x2, x1 = x3, x2 = x4, x3 = x5, x4 = x6;
        if (ws & 1) x0 = x1, x1 = x2, x2 = x3, x3 = x4, x4 = (ws & 2) ? B.w : x5;
        stg_v4(dv + v, make_uint4(__funnelshift_r(x0, x1, bs), __funnelshift_r(x1, x2, bs),
                                  __funnelshift_r(x2, x3, bs), __funnelshift_r(x3, x4, bs)));
    }
    const uint32_t done = head + (nvec << 4);
    if (done + lane < len) d[done + lane] = s[done + lane];
}

// 16 bytes starting `sh` bytes into the 32-byte pair (A, B).
__device__ __forceinline__ void funnel16(const uint4 &A, const uint4 &B, unsigned sh, uint32_t &r0, uint32_t &r1,
                                         uint32_t &r2, uint32_t &r3) {
    const unsigned bs = (sh & 3) * 8;
    uint32_t x0 = A.x, x1 = A.y, x2 = A.z, x3 = A.w, x4 = B.x, x5 = B.y;
    if (sh & 8) x0 = x2, x1 = x3, x2 = x4, x3 = x5, x4 = B.z, x5 = B.w;
    if (sh & 4) x0 = x1, x1 = x2, x2 = x3, x3 = x4, x4 = x5;
    r0 = __funnelshift_r(x0, x1, bs);
    r1 = __funnelshift_r(x1, x2, bs);
    r2 = __funnelshift_r(x2, x3, bs);
    r3 = __funnelshift_r(x3, x4, bs);
}

// Decodes one block whose tag stream pass A validated (status OK, nt slots, checkpoints ck[]).
__device__ __noinline__ void decode_block_v6(const uint8_t *in, uint32_t n_in, uint8_t *out, uint32_t U,
                                             uint32_t nt, const uint2 *ck, const uint32_t *lut, V6Smem *sm) {
    const unsigned lane = lane_id();
    // input: 16-byte aligned view
    const uint32_t ski = (uint32_t)((uintptr_t)in & 15);
    const uint4 *in_v = (const uint4 *)((uintptr_t)in - ski);
    const uint32_t in_last_v = (ski + n_in - 1) >> 4;
    // output: 16-byte aligned coordinates P = offset + sko
    const uint32_t sko = (uint32_t)((uintptr_t)out & 15);
    uint8_t *outA = out - sko;
    const uint4 *win_v = (const uint4 *)sm->win;
    uint32_t wbase = 0;      // P of win[0] (multiple of 16)
    uint32_t hstart = sko;   // window holds valid bytes for P in [hstart, produced)
    uint32_t flushed = sko;  // every byte of P < flushed is in global memory
    const uint32_t G = (nt + SNP6_T - 1) / SNP6_T;

    // writes window bytes P in [flushed, upto & ~15) to global memory as aligned vectors
    auto flush_to = [&](uint32_t upto) {
        const uint32_t fl1 = upto & ~15u;
        if (fl1 > flushed) {
            SNP6_STAT(flushes, 1);
            uint32_t fl0 = flushed & ~15u;
            if (fl0 < sko) {  // first vector of a block whose output is not 16-byte aligned (wbase == 0 here)
                if (lane >= sko && lane < 16u) outA[lane] = sm->win[lane];
                fl0 = 16;
            }
#pragma unroll 1
            for (uint32_t v = (fl0 >> 4) + lane; v < (fl1 >> 4); v += SNP_WARP)
                stg_v4((uint4 *)outA + v, win_v[v - (wbase >> 4)]);
            flushed = fl1;
        }
    };

#pragma unroll 1
    for (uint32_t g0 = 0; g0 < G; g0 += 32) {
        // ================= parse: lane l decodes the slots of group g0 + l into shared memory ===
        {
            const uint32_t my_g = g0 + lane;
            uint32_t ip = 0, op = 0, cnt = 0;
            if (my_g < G) {
                const uint2 c = ck[my_g];
                ip = c.x;
                op = c.y;
                cnt = min(SNP6_T, nt - SNP6_T * my_g);
                sm->gop[lane] = op;
            }
            uint32_t *myrec = sm->rec + lane * 33;
            Stream6 st6;
            st6.open(in_v, in_last_v, sm->pslot + lane * 3);
            const uint32_t kmax = min(SNP6_T, nt - SNP6_T * g0);  // the first group of the step is the fullest
            uint32_t k = 0;
#pragma unroll 1
            for (uint32_t iter = 0; iter < kmax; iter++) {
                if (k < cnt) {
                    const uint32_t pos = ski + ip;
                    uint32_t w0, w1;
                    st6.words(pos, w0, w1);
                    const unsigned sh = (pos & 3) * 8;
                    const uint32_t v = __funnelshift_r(w0, w1, sh);
                    const Tag6 t = decode_tag6(v, w1 >> sh, lut[v & 0xff], n_in - ip);
                    const bool huge = t.take > 64u;
                    if (huge && k == SNP6_T - 1) {
                        myrec[k] = 0;  // pad: the literal opens the next group
                        k = SNP6_T;
                    } else {
                        myrec[k] = (huge ? 0u : t.take) | (t.is_lit ? (SNP6_RLIT | ((ip + t.hdr) << 8)) : (t.off << 8));
                        if (huge) myrec[k + 1] = t.take << 8;  // low byte 0: neither a tag nor a head
                        k += huge ? 2u : 1u;
                        ip += t.hdr + (t.is_lit ? t.take : 0u);
                    }
                }
            }
        }
        __syncwarp();

        // ================= execute the groups in stream order, one slot per lane ==================
        const uint32_t nb = min(32u, G - g0);
#pragma unroll 1
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t cntb = min(SNP6_T, nt - SNP6_T * (g0 + b));
            SNP6_STAT(groups, 1);
            SNP6_STAT(tags, cntb);
            const bool valid = lane < cntb;
            const uint32_t r = valid ? sm->rec[b * 33 + lane] : 0u;
            const uint32_t gbase = sm->gop[b] + sko;
            const bool head = (r & 0xffu) == SNP6_RLIT;  // literal > 64 bytes: its length is in the next slot
            const unsigned hm = __ballot_sync(SNP_FULL, head);
            uint32_t len = r & 0x7fu;  // 0 for heads, length slots and pads
            if (hm) {
                const uint32_t nxt = __shfl_down_sync(SNP_FULL, r, 1);
                if (head) len = nxt >> 8;
            }
            const bool is_lit = (r & SNP6_RLIT) != 0;
            const uint32_t val = r >> 8;  // copy offset / literal input offset
            // output offsets: inclusive scan of the lengths
            uint32_t incl = len;
#pragma unroll
            for (int dlt = 1; dlt < SNP_WARP; dlt <<= 1) {
                const uint32_t y = __shfl_up_sync(SNP_FULL, incl, dlt);
                if (lane >= (unsigned)dlt) incl += y;
            }
            const uint32_t e = gbase + incl, d = e - len;
            unsigned hleft = hm;
            uint32_t t0 = 0;
            for (;;) {
                const uint32_t t1 = hleft ? (uint32_t)(__ffs(hleft) - 1) : cntb;
                if (t1 > t0) {
                    // ------------- sub-group [t0, t1): every tag <= 64 bytes ---------------------
                    SNP6_STAT(subgroups, 1);
                    const uint32_t ss = __shfl_sync(SNP_FULL, d, t0);
                    const uint32_t se = __shfl_sync(SNP_FULL, e, t1 - 1);
                    if (se > wbase + SNP6_WIN) {  // slide the window: keep SNP6_KEEP bytes of history
                        flush_to(ss);
                        uint32_t nbse = ss > SNP6_KEEP ? ss - SNP6_KEEP : 0u;
                        nbse = max(nbse, hstart) & ~15u;
                        if (nbse > wbase) {
                            SNP6_STAT(slides, 1);
                            const uint32_t shv = (nbse - wbase) >> 4;
                            const uint32_t nvec = (ss - nbse + 15) >> 4;
#pragma unroll 1
                            for (uint32_t v0 = 0; v0 < nvec; v0 += SNP_WARP) {
                                const uint32_t v = v0 + lane;
                                uint4 x = make_uint4(0, 0, 0, 0);
                                if (v < nvec) x = win_v[shv + v];
                                __syncwarp();
                                if (v < nvec) ((uint4 *)sm->win)[v] = x;
                                __syncwarp();
                            }
                            wbase = nbse;
                            hstart = max(hstart, nbse);
                        }
                    }
                    const bool mine = valid && lane >= t0 && lane < t1 && len != 0;
                    const uint32_t s_pos = d - val;  // copies: P of the first source byte
                    const uint32_t s_end = s_pos + len;
                    const bool dep = mine && !is_lit && s_end > ss;  // source produced inside this sub-group
                    unsigned pending = __ballot_sync(SNP_FULL, mine);
                    const unsigned depm = __ballot_sync(SNP_FULL, dep);
                    const bool fast = depm == 0;  // no in-group dependency: one round, no frontier
                    SNP6_STAT(fast, fast ? 1 : 0);
                    // Runs: consecutive copies with one offset are the pieces of ONE long match
                    // (EmitCopy splits at 64 bytes), so a piece whose source reaches into the run's own
                    // output reads the PERIODIC extension of the `val` bytes in front of the run -- it
                    // depends on the output before the run only, not on the previous piece.
                    uint32_t rd = d, pbase = 0, pend = 0xffffffffu, ph = 0;
                    bool periodic = false;
                    if (!fast) {
                        const uint32_t key = (mine && !is_lit) ? val : 0xffffffffu;
                        const uint32_t pk = __shfl_up_sync(SNP_FULL, key, 1);
                        const bool cont = mine && !is_lit && lane > t0 && pk == val;
                        const unsigned heads = __ballot_sync(SNP_FULL, !cont);
                        const unsigned hl = 31 - __clz(heads & (lanemask_lt() | (1u << lane)));
                        rd = __shfl_sync(SNP_FULL, d, hl);
                        periodic = mine && !is_lit && s_end > rd;
                        if (periodic) {
                            pbase = rd - val;
                            pend = rd;
                            ph = (d - rd) % val;
                        }
                    }
                    const bool ctype = periodic && val < 16u;  // short period: the warp replicates the pattern
                    // Source forwarding: a copy whose source lies inside ONE plain tag of this sub-group reads
                    // that tag's own source instead (the literal's input bytes, or the older output the tag
                    // copies from), so it does not have to wait for it.  Every hop follows the tag's current
                    // source, i.e. chains shorten by pointer doubling; what cannot be forwarded (a source that
                    // straddles tags) waits for the frontier as before.
                    bool src_in = is_lit;                        // the source is the input stream
                    uint32_t sp = is_lit ? ski + val : s_pos;    // first source byte (input or P coordinates)
                    bool single = fast;                          // no waiting needed: one round
                    bool exact = false;                          // readiness from needm instead of the prefix frontier
                    unsigned needm = 0;                          // lanes whose output this tag's source overlaps
                    if (!fast) {
                        // (a hop costs ~45 instructions: worth it for chains -- records, markup --, not for the one or
                        //  two dependent tags of a text group, which simply take a second round)
                        bool stuck = __popc(depm) < 4;
#pragma unroll 1
                        for (int hop = 0; hop < 3; hop++) {
                            const bool need = mine && !src_in && !periodic && !stuck && sp + len > ss;
                            if (!__any_sync(SNP_FULL, need)) break;
                            SNP6_STAT(hops, 1);
                            uint32_t j = 0;  // last lane whose tag starts at or before sp (d is non-decreasing over lanes)
#pragma unroll
                            for (uint32_t step = 16; step; step >>= 1) {
                                const uint32_t dq = __shfl_sync(SNP_FULL, d, j + step);
                                if (dq <= sp) j += step;
                            }
                            const uint32_t dj = __shfl_sync(SNP_FULL, d, j);
                            const uint32_t ej = __shfl_sync(SNP_FULL, e, j);
                            const uint32_t spj = __shfl_sync(SNP_FULL, sp, j);
                            const unsigned fj = __shfl_sync(SNP_FULL, (mine && !periodic ? 1u : 0u) | (src_in ? 2u : 0u), j);
                            if (need) {
                                if (j >= t0 && (fj & 1u) && sp >= dj && sp + len <= ej) {
                                    src_in = (fj & 2u) != 0;
                                    sp = spj + (sp - dj);
                                } else {
                                    stuck = true;
#ifdef SNP_EMU
                                    if (!(fj & 1u)) v6_stats().stuck_kind++;
                                    else v6_stats().stuck_straddle++;
#endif
                                }
                            }
                        }
                        const bool waits = mine && !src_in && (periodic ? rd > ss : sp + len > ss);
                        const unsigned wm = __ballot_sync(SNP_FULL, waits);
                        single = wm == 0;
                        // Exact readiness: the prefix frontier makes a dependent tag wait for EVERY earlier tag of the
                        // group; with several waiting tags it pays to find, once, the lanes whose output a tag really
                        // reads (the tags are contiguous in the output, so that is a lane interval).
                        if (__popc(wm) >= 3) {
                            const uint32_t lo_p = periodic ? pbase : sp;                 // first / last source byte (P)
                            const uint32_t hi_p = periodic ? rd - 1 : sp + len - 1;
                            uint32_t ja = 0, jb = 0;
#pragma unroll
                            for (uint32_t step = 16; step; step >>= 1) {
                                const uint32_t da = __shfl_sync(SNP_FULL, d, ja + step);
                                const uint32_t db = __shfl_sync(SNP_FULL, d, jb + step);
                                if (da <= lo_p) ja += step;
                                if (db <= hi_p) jb += step;
                            }
                            // lanes ja .. jb produce [lo_p, hi_p]; everything below t0 is final already
                            const unsigned upto = 0xffffffffu >> (31 - jb);
                            needm = waits ? (upto & ~((1u << ja) - 1u) & ~(1u << lane)) : 0u;
                            exact = true;
                        }
                    }
                    while (pending) {
                        SNP6_STAT(rounds, 1);
                        bool ready = mine;
                        if (exact) {
                            ready = mine && ((pending >> lane) & 1u) && (needm & pending) == 0;
                        } else if (!single) {
                            const unsigned f = __ffs(pending) - 1;
                            const uint32_t F = __shfl_sync(SNP_FULL, d, f);  // every byte below F is final
                            ready = mine && ((pending >> lane) & 1u) && (src_in || (periodic ? rd : sp + len) <= F);
                        }
                        // ---- (S) one tag per lane, <= 16 bytes per trip -----------------------------
                        {
                            uint32_t rem = (ready && !ctype) ? len : 0u;
                            uint32_t cd = d - wbase;  // window offset of the next byte to write
                            uint32_t cs = periodic ? pbase + ph : sp;
                            while (__any_sync(SNP_FULL, rem != 0)) {
                                SNP6_STAT(trips, 1);
                                const uint32_t m = min(min(rem, 16u), pend - cs);
                                uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                                if (m) {
                                    const unsigned sh = cs & 15u;
                                    // one generic pointer for the three sources: recent output lives in the
                                    // shared-memory window, older output and the input stream in global memory
                                    const uint4 *vp = src_in ? in_v + (cs >> 4)
                                                     : cs >= hstart ? win_v + ((cs - wbase) >> 4)
                                                                    : (const uint4 *)outA + (cs >> 4);
                                    const uint4 A = ld_any_v4(vp);
                                    uint4 B = make_uint4(0, 0, 0, 0);
                                    if (sh + m > 16u) B = ld_any_v4(vp + 1);
                                    funnel16(A, B, sh, r0, r1, r2, r3);
                                }
                                const uint32_t mx = __reduce_max_sync(SNP_FULL, m);
                                uint8_t *wp = sm->win + cd;
                                st_win_bytes4(wp, r0, m);
                                if (mx > 4) st_win_bytes4(wp + 4, r1, m > 4 ? m - 4 : 0u);
                                if (mx > 8) st_win_bytes4(wp + 8, r2, m > 8 ? m - 8 : 0u);
                                if (mx > 12) st_win_bytes4(wp + 12, r3, m > 12 ? m - 12 : 0u);
                                cs += m;
                                cd += m;
                                rem -= m;
                                if (cs == pend) cs = pbase;  // periodic sources wrap at the end of the period
                            }
                        }
                        // ---- (C) periods < 16 (CopyHelpers.IncrementalCopy's pattern replication)
                        unsigned cm = fast ? 0u : __ballot_sync(SNP_FULL, ready && ctype);
                        while (cm) {
                            SNP6_STAT(ctags, 1);
                            const unsigned i = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const uint32_t dd = __shfl_sync(SNP_FULL, d, i);
                            const uint32_t ll = __shfl_sync(SNP_FULL, len, i);
                            const uint32_t oo = __shfl_sync(SNP_FULL, val, i);
                            const uint32_t pb = __shfl_sync(SNP_FULL, pbase, i);  // pattern = P in [pb, pb + oo), final
                            const uint32_t p0 = __shfl_sync(SNP_FULL, ph, i);
                            const uint8_t *pat = pb >= hstart ? sm->win + (pb - wbase) : outA + pb;
                            for (uint32_t k = lane; k < ll; k += SNP_WARP) sm->win[dd - wbase + k] = pat[(p0 + k) % oo];
                        }
                        __syncwarp();  // this round's window bytes are visible to the next round / the flush
                        pending &= ~__ballot_sync(SNP_FULL, ready);
                    }
                    if (se - flushed >= SNP6_FLUSH) flush_to(se);
                }
                if (t1 >= cntb) break;
                // ------------- literal > 64 bytes: straight from the input to HBM -------------------
                {
                    SNP6_STAT(huge, 1);
                    const uint32_t dd = __shfl_sync(SNP_FULL, d, t1);
                    const uint32_t ll = __shfl_sync(SNP_FULL, len, t1);
                    const uint32_t ls = __shfl_sync(SNP_FULL, val, t1);
                    flush_to(dd);
                    if (dd > flushed) {  // < 16 pending tail bytes of the window
                        const uint32_t n = dd - flushed;
                        if (lane < n) outA[flushed + lane] = sm->win[flushed - wbase + lane];
                    }
                    copy_wide_v6(in + ls, outA + dd, ll, in + n_in, lane);
                    // re-seed the window with the literal's last bytes so that near copies stay on chip
                    const uint32_t keep = min(ll, SNP6_KEEP);
                    const uint32_t ee = dd + ll;
                    __syncwarp();
                    wbase = (ee - keep) & ~15u;
                    hstart = ee - keep;
                    flushed = ee;
                    const uint8_t *tail = in + ls + (ll - keep);
                    for (uint32_t k = lane; k < keep; k += SNP_WARP) sm->win[hstart - wbase + k] = tail[k];
                    __syncwarp();
                }
                hleft &= hleft - 1;
                t0 = t1 + 2;  // skip the head and its length slot
            }
        }
        __syncwarp();  // the next super-step's parse overwrites the records
    }
    // ---- tail: everything still in the window ------------------------------------------------------
    const uint32_t endP = U + sko;
    flush_to(endP);
    if (endP > flushed) {
        const uint32_t n = endP - flushed;
        for (uint32_t k = lane; k < n; k += SNP_WARP) outA[flushed + k] = sm->win[flushed - wbase + k];
    }
    __syncwarp();
}

struct Decode6Args {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint32_t *in_len;
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint32_t *out_cap;
    uint32_t *out_len;
    int32_t *status;
    size_t first_item, n_items;
    unsigned long long *next_item;
    const uint32_t *ntags;
    const uint2 *ck;
};

#ifdef SNP_EMU
// the emulator has no v3/v1 engines: the harness reports blocks that would take them
int v6_emu_fallback(const uint8_t *in, uint32_t n_in, uint8_t *out, uint32_t cap, uint32_t *written);
#endif

// One warp of the persistent decode kernel.
__device__ __forceinline__ void decode_warp_v6(const Decode6Args &a, const uint32_t *lut, V6Smem *sm) {
    const unsigned lane = lane_id();
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(a.next_item, 1ull);
        it = __shfl_sync(SNP_FULL, it, 0);
        if (it >= a.n_items) break;
        const size_t item = a.first_item + (size_t)it;
        const uint32_t nt = a.ntags[it];
        const uint8_t *in = a.in_base + a.in_off[item];
        const uint32_t n_in = a.in_len[item];
        uint8_t *out = a.out_base + a.out_off[item];
        if (nt == SNP6_NT_FALLBACK) {
            uint32_t w = 0;
            int st;
#ifdef SNP_EMU
            st = v6_emu_fallback(in, n_in, out, a.out_cap[item], &w);
#else
            if (n_in >= 0x7fff0000u) {
                st = decompress_block_v1(in, n_in, out, a.out_cap[item], &w);
            } else {  // v3's tag queue borrows the record area
                st = decompress_block_v3(in, n_in, out, a.out_cap[item], &w, lut, (WarpQueue3 *)sm->rec);
            }
#endif
            if (lane == 0) {
                a.out_len[item] = w;
                a.status[item] = st;
            }
            __syncwarp();
            continue;
        }
        if (nt == 0) continue;  // failed in pass A, or an empty block: nothing to produce
        decode_block_v6(in, n_in, out, a.out_len[item], nt, a.ck + (size_t)it * SNP6_CKB, lut, sm);
    }
}

#ifndef SNP_EMU

#define SNP6_SCAN_WARPS 8
#define SNP6_SCAN_CTAS 5
#define SNP6_DEC_WARPS 8
#define SNP6_DEC_CTAS 3

__global__ void __launch_bounds__(SNP6_SCAN_WARPS * 32, SNP6_SCAN_CTAS) k_tagscan_v6(Scan6Args a) {
    __shared__ uint32_t lut[256];
    __shared__ uint4 rings[SNP6_SCAN_WARPS * 32 * 3];  // 48-byte pitch per thread
    lut[threadIdx.x & 255] = tag_lut3_entry(threadIdx.x & 255);
    __syncthreads();
    tagscan_warp_v6(a, lut, rings + threadIdx.x * 3);
}

struct V6DecSmem {
    V6Smem w[SNP6_DEC_WARPS];
    uint32_t lut[256];
};
static_assert(sizeof(WarpQueue3) <= sizeof(uint32_t) * 32 * 33, "v3's queue must fit the record area");

__global__ void __launch_bounds__(SNP6_DEC_WARPS * 32, SNP6_DEC_CTAS) k_decode_v6(Decode6Args a) {
    extern __shared__ __align__(16) uint8_t v6_smem_raw[];
    V6DecSmem *s = (V6DecSmem *)v6_smem_raw;
    s->lut[threadIdx.x & 255] = tag_lut3_entry(threadIdx.x & 255);
    __syncthreads();
    decode_warp_v6(a, s->lut, &s->w[threadIdx.x / SNP_WARP]);
}

#endif  // !SNP_EMU

}  // namespace snp
