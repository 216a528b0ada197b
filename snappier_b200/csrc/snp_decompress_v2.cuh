// snp_decompress_v2.cuh -- warp-parallel batched Snappy block decompressor (sm_100a).
//
// One warp per compressed block, but -- unlike v1 -- the lanes do not walk the tag
// stream redundantly.  Per iteration:
//
//  PARSE   every lane speculatively decodes the tag that would start at input
//          byte ip+lane (one LUT lookup replaces Constants.CharTable); the true
//          tag starts are then found by following next-pointers with one SHFL +
//          one ISETP per tag (lane 0 is a known start); a 5-step shuffle scan of
//          the lengths gives every tag its output offset; offset/length
//          validation happens here, in stream order, for all tags at once.
//  QUEUE   the window's tags are compacted into a 64-entry per-warp queue in
//          shared memory: {dst | is_copy<<31, src} with a sentinel {op}.
//  ROUND   output-centric: lane j produces output byte cur+j.  The tag owning
//          that byte is found by ranking a 32-bit mask of tag-start positions
//          (REDUX.OR + POPC); a copy whose source byte is produced in this same
//          round is resolved by pointer doubling over SHFL (<= 5 steps; this is
//          what makes short-offset / overlapping copies -- CopyHelpers.
//          IncrementalCopy's pattern fill -- exact).  Tags of >= 32 bytes take a
//          cooperative path instead (coalesced literal / back-reference copy).
//
// Semantics restated from /root/reference/Snappier/Internal/SnappyDecompressor.cs
// :43-92, :184-347, :556-611 (one-shot, whole block), identical to v1 and to
// oracle/snappy_oracle.c; the tests A/B all three.
#pragma once
#include "snp_common.cuh"

namespace snp {

// ---- tag LUT (replaces Constants.CharTable, Constants.cs:42-76) -----------------
// bits 0..7  length encoded in the tag byte (literal n<60: n+1; copies: len)
// bits 8..10 total header bytes (tag byte + trailer): 1,2,3,4,5
// bit  11    literal
// bit  12    literal whose length is in the trailer (tag 60..63)
// bits 13..18 shift that turns 0xffffffff into the trailer mask (32 - 8*trailer_bytes)
// bits 19..29 COPY1 high offset bits ((c >> 5) << 8)
__device__ __forceinline__ uint32_t tag_lut_entry(uint32_t c) {
    uint32_t kind = c & 3, n6 = c >> 2;
    uint32_t len, hdr, lit = 0, lng = 0, offhi = 0;
    if (kind == 0) {
        lit = 1;
        if (n6 < 60) {
            len = n6 + 1;
            hdr = 1;
        } else {
            len = 0;
            lng = 1;
            hdr = 1 + (n6 - 59);
        }
    } else if (kind == 1) {
        len = (n6 & 7) + 4;
        hdr = 2;
        offhi = (c >> 5) << 8;
    } else {
        len = n6 + 1;
        hdr = kind == 2 ? 3 : 5;
    }
    uint32_t shift = 32 - 8 * (hdr - 1);
    return len | (hdr << 8) | (lit << 11) | (lng << 12) | (shift << 13) | (offhi << 19);
}

struct WarpQueue {  // per-warp slice of shared memory
    uint32_t dst[64];
    uint32_t src[64];
};

// Cooperative copy of one long tag (>= 32 bytes) starting at output position cur.
__device__ __forceinline__ void copy_long_tag(const uint8_t *__restrict__ in, uint8_t *out, uint32_t cur,
                                              uint32_t len, bool is_copy, uint32_t srcw, uint32_t rel,
                                              unsigned lane) {
    if (!is_copy) {
        const uint8_t *s = in + srcw + rel;
        uint8_t *d = out + cur;
        for (uint32_t k = lane; k < len; k += SNP_WARP) d[k] = s[k];
    } else {
        const uint32_t off = srcw;
        const uint8_t *s = out + (cur - off);
        uint8_t *d = out + cur;
        if (off >= SNP_WARP) {
            for (uint32_t k0 = 0; k0 < len; k0 += SNP_WARP) {
                uint32_t k = k0 + lane;
                if (k < len) d[k] = s[k];
                if (off < len) __syncwarp();  // chunk j+1 may read what chunk j wrote
            }
        } else {
            for (uint32_t k = lane; k < len; k += SNP_WARP) d[k] = s[k % off];  // pattern fill
        }
    }
    __syncwarp();
}

// Returns the block status; *written = bytes produced on SNP_OK, else 0.
__device__ __noinline__ int decompress_block_v2(const uint8_t *__restrict__ in, uint32_t n_in, uint8_t *out,
                                                uint32_t cap, uint32_t *written, const uint32_t *lut,
                                                WarpQueue *q) {
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    *written = 0;
    uint32_t U, used;
    int st = varint_read(in, n_in, &U, &used);
    if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;
    if (st != SNP_OK || U > 0x7fffffffu) return SNP_INVALID_LENGTH;
    if (cap < U) return SNP_OUTPUT_TOO_SMALL;
    if (U == 0) return SNP_OK;

    const uintptr_t in_addr = (uintptr_t)in;
    const uintptr_t in_words_end = (in_addr + n_in + 3) & ~(uintptr_t)3;  // no word at/after this is touched

    uint32_t ip = used;       // parse position in the compressed stream
    uint32_t op = 0;          // output position after every queued tag
    uint32_t cur = 0;         // output position produced so far
    uint32_t head = 0, tail = 0;  // queue: entries head..tail-1 are tags, entry tail is the sentinel
    bool stop = false;
    if (lane == 0) q->dst[0] = 0;
    __syncwarp();

    // ---- one output-centric round (or one cooperative long-tag copy) -------------
    auto round = [&]() {
        const uint32_t e = head + 1 + lane;
        const bool exists = e <= tail;
        const uint32_t d = exists ? (q->dst[e & 63] & 0x7fffffffu) : 0xffffffffu;
        const uint32_t b = d - cur - 1;  // tag e starts at output byte cur+1+b
        const uint32_t rem = __shfl_sync(SNP_FULL, d, 0) - cur;  // bytes left in the head tag
        if (rem >= SNP_WARP) {
            const uint32_t hd = q->dst[head & 63], hs = q->src[head & 63];
            copy_long_tag(in, out, cur, rem, hd >> 31, hs, cur - (hd & 0x7fffffffu), lane);
            cur += rem;
            head += 1;
            return;
        }
        const bool inr = exists && b < SNP_WARP;
        const uint32_t M = __reduce_or_sync(SNP_FULL, inr ? (1u << b) : 0u);
        uint32_t nbytes = min(op - cur, (uint32_t)SNP_WARP);
        {  // stop in front of the first long tag; it takes the cooperative path next
            const uint32_t dn = __shfl_down_sync(SNP_FULL, d, 1);
            const bool lng = inr && (lane < 31) && (e + 1 <= tail) && (dn - d >= SNP_WARP);
            const unsigned lm = __ballot_sync(SNP_FULL, lng);
            if (lm) nbytes = min(nbytes, 1u + __shfl_sync(SNP_FULL, b, __ffs(lm) - 1));
        }
        const bool active = lane < nbytes;
        const uint32_t rank = __popc(M & ((1u << lane) - 1u));
        const uint32_t idx = (head + rank) & 63;
        const uint32_t tdw = q->dst[idx], tsrc = q->src[idx];
        const uint32_t mypos = cur + lane;
        // source descriptor: kind 0 = input byte sa, 1 = output byte sa, 2 = byte of lane sa (this round)
        uint32_t sk, sa;
        if (!(tdw >> 31)) {
            sk = 0;
            sa = tsrc + (mypos - (tdw & 0x7fffffffu));
        } else {
            const uint32_t spos = mypos - tsrc;  // tsrc = copy offset, validated at parse time
            if (spos >= cur) {
                sk = 2;
                sa = spos - cur;
            } else {
                sk = 1;
                sa = spos;
            }
        }
        while (__any_sync(SNP_FULL, active && sk == 2)) {  // pointer doubling, <= 5 trips
            const uint32_t nk = __shfl_sync(SNP_FULL, sk, sa & 31);
            const uint32_t na = __shfl_sync(SNP_FULL, sa, sa & 31);
            if (sk == 2) {
                sk = nk;
                sa = na;
            }
        }
        if (active) {
            const uint8_t *p = (sk == 0 ? in : (const uint8_t *)out) + sa;
            out[mypos] = *p;
        }
        __syncwarp();
        head += __popc(M & (0xffffffffu >> (SNP_WARP - nbytes)));
        cur += nbytes;
    };

    while (!stop && ip < n_in) {
        // ---- PARSE: speculative decode at ip+lane --------------------------------
        const uint32_t pos = ip + lane;
        const uintptr_t a = in_addr + pos;
        const uint32_t *wp = (const uint32_t *)(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        const uint32_t w0 = ((uintptr_t)wp < in_words_end) ? wp[0] : 0u;
        const uint32_t w1 = ((uintptr_t)(wp + 1) < in_words_end) ? wp[1] : 0u;
        const uint32_t v = __funnelshift_r(w0, w1, sh);
        const uint32_t c = v & 0xff;
        const uint32_t trailer = (v >> 8) | ((w1 >> sh) << 24);
        const uint32_t ent = lut[c];
        const uint32_t hdr = (ent >> 8) & 7;
        const bool is_lit = (ent >> 11) & 1;
        const uint32_t tval = trailer & __funnelshift_rc(0xffffffffu, 0u, (ent >> 13) & 63);
        uint32_t len = ent & 0xff;
        if ((ent >> 12) & 1) len = tval == 0xffffffffu ? 0xffffffffu : tval + 1;  // long literal, saturating
        const uint32_t off = (ent >> 19) | tval;  // copies only
        // classification against the end of the input (SnappyDecompressor.cs:236-297,464-483)
        const uint32_t left = pos < n_in ? n_in - pos : 0;  // bytes from the tag byte to the end
        const bool is_end = left < hdr || left == 0;          // no tag / truncated tag: parsing stops here
        uint32_t take = len;
        bool partial = false;
        if (is_lit && !is_end) {
            const uint32_t avail = left - hdr;
            if (len > avail) {
                take = avail;
                partial = true;
            }
        }
        const uint32_t adv = hdr + (is_lit ? take : 0u);  // <= n_in, no overflow
        const uint32_t nxt_true = lane + adv;
        const uint32_t nxt = (is_end || partial || nxt_true >= SNP_WARP) ? 63u : nxt_true;

        // ---- chain: which lanes are real tag starts (lane 0 is) -------------------
        bool is_start = lane == 0;
        {
            uint32_t p = 0;
#pragma unroll 1
            for (int g = 0; g < 4; g++) {  // <= 16 tags fit in 32 bytes
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    p = __shfl_sync(SNP_FULL, nxt, p & 31);  // lane 31 always holds 63: a fixed point
                    is_start |= (p == lane);
                }
                if (p >= SNP_WARP) break;
            }
        }
        const bool is_tag = is_start && !is_end && take > 0;
        const unsigned starts = __ballot_sync(SNP_FULL, is_start);
        const unsigned tags = __ballot_sync(SNP_FULL, is_tag);
        if (__ballot_sync(SNP_FULL, is_start && (is_end || partial))) stop = true;
        const uint32_t ip_next = ip + __shfl_sync(SNP_FULL, nxt_true, 31 - __clz(starts));

        // ---- output offsets: exclusive scan of the tag lengths --------------------
        // Every tag but the window's last is <= 64 bytes, so only `total` can be large.
        const uint32_t x = is_tag ? take : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int dlt = 1; dlt < SNP_WARP; dlt <<= 1) {
            const uint32_t y = __shfl_up_sync(SNP_FULL, incl, dlt);
            if (lane >= (unsigned)dlt) incl += y;
        }
        const uint32_t dst = op + (incl - x);

        // ---- validation in stream order (SnappyDecompressor.cs:570-573,598-606) ----
        int err = SNP_OK;
        if (is_tag) {
            if (!is_lit && (off == 0 || off > dst)) err = SNP_INVALID_COPY_OFFSET;
            else if (take > U - dst) err = SNP_DATA_TOO_LONG;
        }
        const unsigned errs = __ballot_sync(SNP_FULL, err != SNP_OK);
        if (errs) return __shfl_sync(SNP_FULL, err, __ffs(errs) - 1);

        // ---- QUEUE append ----------------------------------------------------------
        if (is_tag) {
            const uint32_t slot = (tail + __popc(tags & lt)) & 63;
            q->dst[slot] = dst | (is_lit ? 0u : 0x80000000u);
            q->src[slot] = is_lit ? pos + hdr : off;
        }
        tail += __popc(tags);
        op += __shfl_sync(SNP_FULL, incl, 31);
        if (lane == 0) q->dst[tail & 63] = op;  // sentinel
        __syncwarp();
        ip = ip_next;

        // ---- ROUNDS: keep < 32 bytes (hence < 32 tags) queued ----------------------
        while (op - cur >= SNP_WARP) round();
    }
    while (cur < op) round();

    if (op < U) return SNP_INCOMPLETE;  // Snappy.cs:178-181
    *written = op;
    return SNP_OK;
}

__global__ void __launch_bounds__(256)
k_decompress_v2(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_len, uint8_t *out_base,
                const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
                uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items) {
    __shared__ uint32_t lut[256];
    __shared__ WarpQueue queues[8];
    lut[threadIdx.x & 255] = tag_lut_entry(threadIdx.x & 255);
    __syncthreads();
    size_t item = (size_t)blockIdx.x * (blockDim.x / SNP_WARP) + threadIdx.x / SNP_WARP;
    if (item >= n_items) return;
    uint32_t w = 0;
    int st = decompress_block_v2(in_base + in_off[item], in_len[item], out_base + out_off[item], out_cap[item],
                                 &w, lut, &queues[threadIdx.x / SNP_WARP]);
    if (lane_id() == 0) {
        out_len[item] = w;
        status[item] = st;
    }
}

}  // namespace snp
