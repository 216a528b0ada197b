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
//           (ip, op) every 32 tags.  Input reaches each lane through a 128-byte shared-memory slot
//           that the warp refills with one coalesced load.
//   pass B  k_decode_v6    one WARP per block (the north star's decomposition).  Per super-step
//           lane l re-parses the 32 tags behind checkpoint 32*j+l into shared memory (1024 tags in
//           flight), then the warp executes the groups in stream order, ONE TAG PER LANE: 16-byte
//           vector loads of the source (input stream, shared-memory output window, or the block's
//           older output in global memory), a register funnel to the byte offset, byte stores into
//           a sliding shared-memory OUTPUT WINDOW, which is flushed to HBM as aligned 16-byte
//           vectors.  Tags whose source is produced inside the same group wait for the frontier
//           (multi-round resolution; the first pending tag is always runnable).  Overlapping
//           copies with offset < 16 (CopyHelpers.IncrementalCopy's pattern replication) and
//           literals > 64 bytes take warp-cooperative paths.
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

#define SNP6_T 32u  // tags per checkpoint group (= lanes)
#ifndef SNP6_CKB
#define SNP6_CKB 768u  // checkpoint budget per block: 24 576 tags; denser blocks fall back to v5
#endif
#define SNP6_NT_FALLBACK 0xffffffffu
#define SNP6_KEEP 1024u                             // bytes of history the output window keeps across a slide
#define SNP6_SPAN 2048u                             // 32 tags x 64 bytes: most a sub-group can produce
#define SNP6_WIN (SNP6_KEEP + SNP6_SPAN + 64u)      // window bytes
#define SNP6_LIT 0x80000000u
#define SNP6_SLOTW 33u  // words per pass-A input slot row (32 + 1 pad: conflict-free rows and columns)

struct alignas(16) V6Smem {  // pass B, per warp
    uint32_t dst[32 * 33];   // [group lane][tag] output offset; entry [cnt] = end of the group
    uint32_t src[32 * 33];   // literal: SNP6_LIT | input offset of the first byte; copy: offset
    uint8_t win[SNP6_WIN + 16];
};

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
    uint32_t *ntags;  // per wave slot: tag count, or SNP6_NT_FALLBACK
    uint2 *ck;        // per wave slot: SNP6_CKB checkpoints (ip, op)
};

// One warp of the persistent scan: every lane owns one block at a time and fetches the next one
// when it is done.  `slots` = 32 rows of SNP6_SLOTW words of shared memory.
__device__ __forceinline__ void tagscan_warp_v6(const Scan6Args &a, const uint32_t *lut, uint32_t *slots) {
    const unsigned lane = lane_id();
    bool have = false, exhausted = false;
    size_t slot_idx = 0, item = 0;
    const uint32_t *in_w = nullptr;
    uint32_t skew = 0, n_in = 0, last_w = 0, U = 0, ip = 0, op = 0, ntag = 0, sbyte = 0;
    bool svalid = false;
    int st = SNP_OK;
    uint32_t *myslot = slots + lane * SNP6_SLOTW;

    for (;;) {
        if (!have && !exhausted) {
            const unsigned long long it = atomicAdd(a.next_item, 1ull);
            if (it >= a.n_items) {
                exhausted = true;
            } else {
                slot_idx = (size_t)it;
                item = a.first_item + slot_idx;
                const uint8_t *in = a.in_base + a.in_off[item];
                n_in = a.in_len[item];
                const uint32_t cap = a.out_cap[item];
                if (n_in >= 0x7fff0000u) {  // 32-bit stream offsets need headroom: v1 walker in pass B
                    a.ntags[slot_idx] = SNP6_NT_FALLBACK;
                } else {
                    uint32_t used;
                    st = varint_read(in, n_in, &U, &used);  // SnappyDecompressor.cs:50-63
                    if (st == SNP_OK && U > 0x7fffffffu) st = SNP_INVALID_LENGTH;
                    if (st == SNP_OK && cap < U) st = SNP_OUTPUT_TOO_SMALL;
                    if (st != SNP_OK || U == 0) {
                        a.out_len[item] = 0;
                        a.status[item] = st;
                        a.ntags[slot_idx] = 0;
                    } else {
                        skew = (uint32_t)((uintptr_t)in & 3);
                        in_w = (const uint32_t *)((uintptr_t)in - skew);
                        last_w = (skew + n_in - 1) >> 2;
                        ip = used;
                        op = 0;
                        ntag = 0;
                        svalid = false;
                        have = true;
                    }
                }
            }
        }
        if (__all_sync(SNP_FULL, exhausted && !have)) break;

        // ---- refill: the warp loads 128 bytes for every lane whose slot ran out ------------
        const uint32_t pos0 = skew + ip;
        const bool need = have && (!svalid || pos0 - sbyte > 123u);
        unsigned m = __ballot_sync(SNP_FULL, need);
        __syncwarp();  // every lane is done reading the slots it is about to see overwritten
        while (m) {
            const unsigned b = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t w0 = __shfl_sync(SNP_FULL, pos0 >> 2, b);
            const uint32_t lw = __shfl_sync(SNP_FULL, last_w, b);
            const unsigned long long base = __shfl_sync(SNP_FULL, (unsigned long long)(uintptr_t)in_w, b);
            slots[b * SNP6_SLOTW + lane] = ((const uint32_t *)(uintptr_t)base)[min(w0 + lane, lw)];
        }
        if (need) {
            sbyte = pos0 & ~3u;
            svalid = true;
        }
        __syncwarp();

        // ---- up to 4 tags per lane from the slot ---------------------------------------------
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            if (!have) break;
            if (ip >= n_in) {
                have = false;
            } else {
                const uint32_t rel = skew + ip - sbyte;
                if (rel > 123u) break;  // next refill
                const unsigned sh = (rel & 3) * 8;
                const uint32_t w0 = myslot[rel >> 2], w1 = myslot[(rel >> 2) + 1];
                const uint32_t v = __funnelshift_r(w0, w1, sh);
                const Tag6 t = decode_tag6(v, w1 >> sh, lut[v & 0xff], n_in - ip);
                if (t.end) {
                    have = false;
                } else if (!t.is_lit && t.off - 1u >= op) {  // off == 0 || off > produced (:598-601)
                    st = SNP_INVALID_COPY_OFFSET;
                    have = false;
                } else if (t.take > U - op) {  // :570-573, :603-606
                    st = SNP_DATA_TOO_LONG;
                    have = false;
                } else {
                    if (t.take) {
                        if ((ntag & (SNP6_T - 1)) == 0) {
                            const uint32_t g = ntag / SNP6_T;
                            if (g >= SNP6_CKB) {  // denser than the checkpoint budget: pass B decodes it with v5
                                a.ntags[slot_idx] = SNP6_NT_FALLBACK;
                                have = false;
                                break;
                            }
                            a.ck[slot_idx * SNP6_CKB + g] = make_uint2(ip, op);
                        }
                        ntag++;
                    }
                    op += t.take;
                    ip += t.hdr + (t.is_lit ? t.take : 0u);
                    if (t.partial) have = false;
                }
            }
            if (!have) {  // the stream ended or failed: this is the block's result
                if (st == SNP_OK && op < U) st = SNP_INCOMPLETE;  // Snappy.cs:178-181
                a.out_len[item] = st == SNP_OK ? op : 0u;
                a.status[item] = st;
                a.ntags[slot_idx] = st == SNP_OK ? ntag : 0u;
            }
        }
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
        const uint4 A = base[v];
        const uint4 *pb = base + v + 1;
        const uint4 B = *(pb <= last ? pb : last);
        uint32_t x0 = A.x, x1 = A.y, x2 = A.z, x3 = A.w, x4 = B.x, x5 = B.y, x6 = B.z;
        if (ws & 2) x0 = x2, x1 = x3, x2 = x4, x3 = x5, x4 = x6;
        if (ws & 1) x0 = x1, x1 = x2, x2 = x3, x3 = x4, x4 = (ws & 2) ? B.w : x5;
        dv[v] = make_uint4(__funnelshift_r(x0, x1, bs), __funnelshift_r(x1, x2, bs), __funnelshift_r(x2, x3, bs),
                           __funnelshift_r(x3, x4, bs));
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

// Per-lane 32-byte register window over the lane's own part of the input stream (pass B parse).
struct InWin6 {
    const uint4 *base;  // 16-byte aligned; stream byte p sits at byte (skew + p)
    uint32_t last_v;    // last vector index that may be read
    uint32_t vb;        // vector index held in c (n holds vb + 1)
    bool valid;
    uint4 c, n;
    __device__ __forceinline__ void words(uint32_t wi, uint32_t &w0, uint32_t &w1) {
        const uint32_t vq = wi >> 2;
        if (!valid || vq != vb) {
            if (valid && vq == vb + 1) c = n;
            else c = base[min(vq, last_v)];
            n = base[min(vq + 1, last_v)];
            vb = vq;
            valid = true;
        }
        const unsigned sel = wi & 3;
        w0 = sel == 0 ? c.x : sel == 1 ? c.y : sel == 2 ? c.z : c.w;
        w1 = sel == 0 ? c.y : sel == 1 ? c.z : sel == 2 ? c.w : n.x;
    }
};

// Decodes one block whose tag stream pass A validated (status OK, nt tags, checkpoints ck[]).
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
    uint32_t wbase = 0;       // P of win[0] (multiple of 16)
    uint32_t hstart = sko;    // window holds valid bytes for P in [hstart, produced)
    uint32_t flushed = sko;   // every byte of P < flushed is in global memory
    const uint32_t G = (nt + SNP6_T - 1) / SNP6_T;

    for (uint32_t g0 = 0; g0 < G; g0 += 32) {
        // ================= parse: lane l decodes the tags of group g0 + l into shared memory ===
        {
            const uint32_t my_g = g0 + lane;
            const bool has = my_g < G;
            uint32_t ip = 0, op = 0, cnt = 0;
            if (has) {
                const uint2 c = ck[my_g];
                ip = c.x;
                op = c.y;
                cnt = min(SNP6_T, nt - SNP6_T * my_g);
            }
            InWin6 w;
            w.base = in_v;
            w.last_v = in_last_v;
            w.vb = 0;
            w.valid = false;
            const uint32_t kmax = min(SNP6_T, nt - SNP6_T * g0);  // lane 0 has the most tags
#pragma unroll 1
            for (uint32_t k = 0; k < kmax; k++) {
                if (k < cnt) {
                    const uint32_t pos = ski + ip;
                    uint32_t w0, w1;
                    w.words(pos >> 2, w0, w1);
                    const unsigned sh = (pos & 3) * 8;
                    const uint32_t v = __funnelshift_r(w0, w1, sh);
                    const Tag6 t = decode_tag6(v, w1 >> sh, lut[v & 0xff], n_in - ip);
                    sm->dst[lane * 33 + k] = op;
                    sm->src[lane * 33 + k] = t.is_lit ? (SNP6_LIT | (ip + t.hdr)) : t.off;
                    op += t.take;
                    ip += t.hdr + (t.is_lit ? t.take : 0u);
                }
            }
            if (has) sm->dst[lane * 33 + cnt] = op;
        }
        __syncwarp();

        // ================= execute the groups in stream order, one tag per lane ==================
        const uint32_t nb = min(32u, G - g0);
#pragma unroll 1
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t cntb = min(SNP6_T, nt - SNP6_T * (g0 + b));
            const bool valid = lane < cntb;
            uint32_t d = 0, e = 0, sr = 0;
            if (valid) {
                d = sm->dst[b * 33 + lane] + sko;
                e = sm->dst[b * 33 + lane + 1] + sko;
                sr = sm->src[b * 33 + lane];
            }
            const uint32_t len = e - d;
            const bool is_lit = (sr & SNP6_LIT) != 0;
            const uint32_t off = sr;                  // copies
            const uint32_t lsrc = sr & ~SNP6_LIT;     // literals: input offset
            unsigned hm = __ballot_sync(SNP_FULL, valid && len > 64u);  // literals beyond the tag-per-lane path
            uint32_t t0 = 0;
            for (;;) {
                const uint32_t t1 = hm ? (uint32_t)(__ffs(hm) - 1) : cntb;
                if (t1 > t0) {
                    // ------------- sub-group [t0, t1): every tag <= 64 bytes ---------------------
                    const uint32_t ss = __shfl_sync(SNP_FULL, d, t0);
                    const uint32_t se = __shfl_sync(SNP_FULL, e, t1 - 1);
                    if (se > wbase + SNP6_WIN) {  // slide the window: keep SNP6_KEEP bytes of history
                        uint32_t nbse = ss > SNP6_KEEP ? ss - SNP6_KEEP : 0u;
                        nbse = max(nbse, hstart) & ~15u;
                        if (nbse > wbase) {
                            const uint32_t shv = (nbse - wbase) >> 4;
                            const uint32_t nvec = (ss - nbse + 15) >> 4;
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
                    const bool mine = valid && lane >= t0 && lane < t1;
                    const uint32_t s_pos = d - off;  // copies: P of the first source byte
                    const uint32_t s_end = s_pos + len;
                    const bool ctype = mine && !is_lit && off < 16u && len > off;  // pattern replication
                    unsigned pending = (t1 >= 32 ? 0xffffffffu : ((1u << t1) - 1u)) & ~((1u << t0) - 1u);
                    while (pending) {
                        const unsigned f = __ffs(pending) - 1;
                        const uint32_t F = __shfl_sync(SNP_FULL, d, f);  // every byte below F is final
                        const bool ready =
                            mine && ((pending >> lane) & 1u) && (is_lit || s_end <= F || lane == f);
                        // ---- (S) one tag per lane, 16 bytes per trip --------------------------------
                        {
                            uint32_t rem = (ready && !ctype) ? len : 0u;
                            uint32_t cd = d;
                            uint32_t cs = is_lit ? ski + lsrc : s_pos;
                            while (__any_sync(SNP_FULL, rem != 0)) {
                                const uint32_t m = min(rem, 16u);
                                uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                                if (m) {
                                    const unsigned sh = cs & 15u;
                                    const bool need_b = sh + m > 16u;
                                    uint4 A, B;
                                    if (!is_lit && cs >= hstart) {  // recent output: shared-memory window
                                        const uint32_t vi = (cs - wbase) >> 4;
                                        A = win_v[vi];
                                        B = win_v[vi + 1];
                                    } else {  // input stream, or output older than the window
                                        const uint4 *gp = is_lit ? in_v : (const uint4 *)outA;
                                        A = gp[cs >> 4];
                                        B = A;
                                        if (need_b) B = gp[(cs >> 4) + 1];
                                    }
                                    funnel16(A, B, sh, r0, r1, r2, r3);
                                }
                                const uint32_t mx = __reduce_max_sync(SNP_FULL, m);
                                uint8_t *wp = sm->win + (cd - wbase);
                                if (mx > 0) {
                                    if (m > 0) wp[0] = (uint8_t)r0;
                                    if (m > 1) wp[1] = (uint8_t)(r0 >> 8);
                                    if (m > 2) wp[2] = (uint8_t)(r0 >> 16);
                                    if (m > 3) wp[3] = (uint8_t)(r0 >> 24);
                                }
                                if (mx > 4) {
                                    if (m > 4) wp[4] = (uint8_t)r1;
                                    if (m > 5) wp[5] = (uint8_t)(r1 >> 8);
                                    if (m > 6) wp[6] = (uint8_t)(r1 >> 16);
                                    if (m > 7) wp[7] = (uint8_t)(r1 >> 24);
                                }
                                if (mx > 8) {
                                    if (m > 8) wp[8] = (uint8_t)r2;
                                    if (m > 9) wp[9] = (uint8_t)(r2 >> 8);
                                    if (m > 10) wp[10] = (uint8_t)(r2 >> 16);
                                    if (m > 11) wp[11] = (uint8_t)(r2 >> 24);
                                }
                                if (mx > 12) {
                                    if (m > 12) wp[12] = (uint8_t)r3;
                                    if (m > 13) wp[13] = (uint8_t)(r3 >> 8);
                                    if (m > 14) wp[14] = (uint8_t)(r3 >> 16);
                                    if (m > 15) wp[15] = (uint8_t)(r3 >> 24);
                                }
                                cs += 16;
                                cd += 16;
                                rem -= m;
                            }
                        }
                        // ---- (C) overlapping copies with offset < 16: the warp replicates the pattern
                        unsigned cm = __ballot_sync(SNP_FULL, ready && ctype);
                        while (cm) {
                            const unsigned i = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const uint32_t dd = __shfl_sync(SNP_FULL, d, i);
                            const uint32_t ll = __shfl_sync(SNP_FULL, len, i);
                            const uint32_t oo = __shfl_sync(SNP_FULL, off, i);
                            const uint32_t sp = dd - oo;  // pattern = P in [sp, dd), final since dd == F
                            const uint8_t *pat = sp >= hstart ? sm->win + (sp - wbase) : outA + sp;
                            for (uint32_t k = lane; k < ll; k += SNP_WARP) sm->win[dd - wbase + k] = pat[k % oo];
                        }
                        __syncwarp();  // this round's window bytes are visible to the next round
                        pending &= ~__ballot_sync(SNP_FULL, ready);
                    }
                    // ------------- flush the finished 16-byte vectors of the window to HBM -----
                    {
                        const uint32_t fl1 = se & ~15u;
                        if (fl1 > flushed) {
                            uint32_t fl0 = flushed & ~15u;
                            if (fl0 < sko) {  // first vector of a block whose output is not 16-byte aligned
                                if (lane >= sko && lane < 16u) outA[lane] = sm->win[lane];
                                fl0 = 16;
                            }
                            for (uint32_t v = (fl0 >> 4) + lane; v < (fl1 >> 4); v += SNP_WARP)
                                ((uint4 *)outA)[v] = win_v[v - (wbase >> 4)];
                            flushed = fl1;
                        }
                        __syncwarp();
                    }
                }
                if (t1 >= cntb) break;
                // ------------- literal > 64 bytes: straight from the input to HBM -------------------
                {
                    const uint32_t dd = __shfl_sync(SNP_FULL, d, t1);
                    const uint32_t ll = __shfl_sync(SNP_FULL, len, t1);
                    const uint32_t ls = __shfl_sync(SNP_FULL, lsrc, t1);
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
                hm &= hm - 1;
                t0 = t1 + 1;
            }
        }
        __syncwarp();  // the next super-step's parse overwrites the tag records
    }
    // ---- tail: the bytes behind the last full vector ------------------------------------------
    const uint32_t endP = U + sko;
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
// the emulator has no v5/v1 engines: the harness reports blocks that would take them
int v6_emu_fallback(const uint8_t *in, uint32_t n_in, uint8_t *out, uint32_t cap, uint32_t *written);
#endif

// One warp of the persistent decode kernel.  `q` (v5's tag queue) is only used by the fallback.
__device__ __forceinline__ void decode_warp_v6(const Decode6Args &a, const uint32_t *lut, V6Smem *sm
#ifndef SNP_EMU
                                               ,
                                               WarpQueue3 *q
#endif
) {
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
            } else {
                st = decompress_block_v3(in, n_in, out, a.out_cap[item], &w, lut, q);
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
#define SNP6_SCAN_CTAS 6
#define SNP6_DEC_WARPS 8
#define SNP6_DEC_CTAS 2

__global__ void __launch_bounds__(SNP6_SCAN_WARPS * 32, SNP6_SCAN_CTAS) k_tagscan_v6(Scan6Args a) {
    __shared__ uint32_t lut[256];
    __shared__ uint32_t slots[SNP6_SCAN_WARPS][32 * SNP6_SLOTW];
    lut[threadIdx.x & 255] = tag_lut3_entry(threadIdx.x & 255);
    __syncthreads();
    tagscan_warp_v6(a, lut, slots[threadIdx.x / SNP_WARP]);
}

struct V6DecSmem {
    V6Smem w[SNP6_DEC_WARPS];
    WarpQueue3 q[SNP6_DEC_WARPS];
    uint32_t lut[256];
};

__global__ void __launch_bounds__(SNP6_DEC_WARPS * 32, SNP6_DEC_CTAS) k_decode_v6(Decode6Args a) {
    extern __shared__ __align__(16) uint8_t v6_smem_raw[];
    V6DecSmem *s = (V6DecSmem *)v6_smem_raw;
    s->lut[threadIdx.x & 255] = tag_lut3_entry(threadIdx.x & 255);
    __syncthreads();
    const unsigned w = threadIdx.x / SNP_WARP;
    decode_warp_v6(a, s->lut, &s->w[w], &s->q[w]);
}

#endif  // !SNP_EMU

}  // namespace snp
