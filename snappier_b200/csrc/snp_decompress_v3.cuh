// snp_decompress_v3.cuh -- warp-parallel batched Snappy block decompressor (sm_100a).
//
// Same algorithm as v2 (speculative 32-position parse -> tag queue in shared
// memory -> output-centric rounds with pointer doubling), re-engineered for
// instruction and L1-wavefront count, which is what bounds this kernel (ncu:
// profiles/r01_decompress_v2.md):
//   * LUT fields sit on byte boundaries (PRMT extracts), flags in the sign bits;
//   * all stream addressing is 32-bit offsets from a 4-byte-aligned base;
//   * tag starts are found with a next-of-next table: 1 + 8 SHFL + 1 REDUX instead
//     of up to 16 dependent SHFLs;
//   * every branch condition that spans a *_sync primitive is a warp vote, so the
//     compiler emits no divergence guards (BRA.DIV/BSSY) around them.
//
// Semantics: /root/reference/Snappier/Internal/SnappyDecompressor.cs:43-92,184-347,
// 556-611 (one-shot); identical results to v1, v2 and oracle/snappy_oracle.c.
#pragma once
#include "snp_common.cuh"
#include "snp_decompress_v1.cuh"

namespace snp {

// tag_lut3_entry (the 256-entry tag table replacing Constants.CharTable) lives in snp_common.cuh

#ifndef SNP_V3_CTAS
#define SNP_V3_CTAS 8
#endif
#define SNP_QCAP 64  // queue entries per warp (power of two)

struct WarpQueue3 {
    uint32_t dst[SNP_QCAP];  // output offset of the tag | is_copy << 31   (entry `tail` = sentinel: op)
    uint32_t src[SNP_QCAP];  // literal: input offset of its first byte; copy: back-reference offset
};

__device__ __forceinline__ void copy_long_tag3(const uint8_t *__restrict__ in, uint8_t *out, uint32_t cur,
                                               uint32_t len, bool is_copy, uint32_t srcw, unsigned lane) {
    uint8_t *d = out + cur;
    if (!is_copy) {
        const uint8_t *s = in + srcw;
        for (uint32_t k = lane; k < len; k += SNP_WARP) d[k] = s[k];
    } else {
        const uint32_t off = srcw;
        const uint8_t *s = d - off;
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

__device__ __noinline__ int decompress_block_v3(const uint8_t *__restrict__ in, uint32_t n_in, uint8_t *out,
                                                uint32_t cap, uint32_t *written, const uint32_t *lut,
                                                WarpQueue3 *q) {
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    *written = 0;
    uint32_t U, used;
    int st = varint_read(in, n_in, &U, &used);
    if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;
    if (st != SNP_OK || U > 0x7fffffffu) return SNP_INVALID_LENGTH;
    if (cap < U) return SNP_OUTPUT_TOO_SMALL;
    if (U == 0) return SNP_OK;

    // 4-byte-aligned view of the input: stream byte p lives at byte (skew + p) of in_w[]
    const uint32_t skew = (uint32_t)((uintptr_t)in & 3);
    const uint32_t *in_w = (const uint32_t *)((uintptr_t)in - skew);
    const uint32_t last_w = (skew + n_in - 1) >> 2;  // no word beyond this index is touched

    uint32_t ip = used, op = 0, cur = 0, head = 0, tail = 0;
    bool stop = false;
    if (lane == 0) q->dst[0] = 0;
    __syncwarp();

    // ---- drain the queue while `want` more bytes than `keep` are queued -----------
    auto drain = [&](uint32_t keep) {
        while (op - cur > keep) {  // op, cur, keep are warp-uniform by construction
            const uint32_t e = head + 1 + lane;
            const bool exists = e <= tail;
            const uint32_t d = exists ? (q->dst[e & (SNP_QCAP - 1)] & 0x7fffffffu) : 0xffffffffu;
            const uint32_t b = d - cur - 1;  // tag e starts at output byte cur+1+b
            const uint32_t rem = __shfl_sync(SNP_FULL, d, 0) - cur;  // bytes left in the head tag
            if (rem >= SNP_WARP) {  // long tag: cooperative path
                const uint32_t hd = q->dst[head & (SNP_QCAP - 1)], hs = q->src[head & (SNP_QCAP - 1)];
                const bool isc = hd >> 31;
                copy_long_tag3(in, out, cur, rem, isc, isc ? hs : hs + (cur - hd), lane);
                cur += rem;
                head += 1;
                continue;
            }
            const bool inr = b < SNP_WARP;
            const uint32_t M = __reduce_or_sync(SNP_FULL, inr ? (1u << b) : 0u);
            uint32_t nbytes = min(op - cur, (uint32_t)SNP_WARP);
            {  // stop in front of the first long tag; it takes the cooperative path next
                const uint32_t dn = __shfl_down_sync(SNP_FULL, d, 1);
                const bool lng = inr && lane < 31 && dn != 0xffffffffu && (dn - d >= SNP_WARP);
                const unsigned lm = __ballot_sync(SNP_FULL, lng);
                if (lm) nbytes = min(nbytes, 1u + __shfl_sync(SNP_FULL, b, __ffs(lm) - 1));
            }
            const bool active = lane < nbytes;
            const uint32_t idx = (head + __popc(M & lt)) & (SNP_QCAP - 1);
            const uint32_t tdw = q->dst[idx], tsrc = q->src[idx];
            const uint32_t mypos = cur + lane;
            // source: sk 0 = input byte sa, 1 = output byte sa, 2 = byte produced by lane sa this round
            uint32_t sk, sa;
            if ((int32_t)tdw >= 0) {
                sk = 0;
                sa = tsrc + (mypos - tdw);
            } else {
                const uint32_t spos = mypos - tsrc;  // tsrc = offset, validated at parse time
                const bool internal = spos >= cur;
                sk = internal ? 2u : 1u;
                sa = internal ? spos - cur : spos;
            }
            while (__any_sync(SNP_FULL, active && sk == 2)) {  // pointer doubling, <= 5 trips
                const uint32_t nk = __shfl_sync(SNP_FULL, sk, sa);
                const uint32_t na = __shfl_sync(SNP_FULL, sa, sa);
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
        }
    };

    while (__any_sync(SNP_FULL, !stop && ip < n_in)) {
        // ---- PARSE: speculative decode of the tag that would start at ip+lane ------
        const uint32_t pos = ip + lane;
        const uint32_t bo = skew + pos;
        const uint32_t wi = bo >> 2;
        const unsigned sh = (bo & 3) * 8;
        const uint32_t w0 = in_w[min(wi, last_w)];
        const uint32_t w1 = in_w[min(wi + 1, last_w)];
        const uint32_t v = __funnelshift_r(w0, w1, sh);
        const uint32_t trailer = __funnelshift_r(v, w1 >> sh, 8);  // bytes pos+1 .. pos+4
        const uint32_t ent = lut[v & 0xff];
        const uint32_t hdr = __byte_perm(ent, 0, 0x4441);
        const bool is_lit = (int32_t)ent < 0;
        const uint32_t tval = trailer & __funnelshift_rc(0xffffffffu, 0u, __byte_perm(ent, 0, 0x4442));
        uint32_t len = ent & 0xff;
        if (ent & 0x40000000u) len = max(tval + 1, tval);  // trailer-length literal, saturating
        const uint32_t off = ((ent >> 16) & 0x700u) | tval;  // copies only
        // against the end of the input (SnappyDecompressor.cs:236-297,464-483)
        const uint32_t left = max(n_in, pos) - pos;  // bytes from the tag byte to the end (0 if past it)
        const bool is_end = left < hdr;               // nothing here / truncated tag: parsing stops
        const uint32_t avail = left - hdr;
        const bool partial = is_lit && !is_end && len > avail;
        const uint32_t take = partial ? avail : len;
        const uint32_t nxt_true = lane + hdr + (is_lit ? take : 0u);
        const uint32_t n1 = (is_end || partial || nxt_true >= SNP_WARP) ? 63u : nxt_true;

        // ---- tag starts: even-indexed tags by walking next-of-next from lane 0,
        //      odd-indexed tags are the `next` of an even one.  Lane 31 always holds 63.
        const uint32_t n2 = __shfl_sync(SNP_FULL, n1, n1);
        bool even = lane == 0;
        {
            uint32_t p = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {  // <= 16 tags fit in 32 bytes
                p = __shfl_sync(SNP_FULL, n2, p);
                even |= (p == lane);
            }
        }
        const unsigned starts = __ballot_sync(SNP_FULL, even) |
                                __reduce_or_sync(SNP_FULL, (even && n1 < SNP_WARP) ? (1u << n1) : 0u);
        const bool is_start = (starts >> lane) & 1;
        const bool is_tag = is_start && !is_end && take > 0;
        const unsigned tags = __ballot_sync(SNP_FULL, is_tag);
        stop = __any_sync(SNP_FULL, is_start && (is_end || partial));
        const uint32_t ip_next = ip + __shfl_sync(SNP_FULL, nxt_true, 31 - __clz(starts));

        // ---- output offsets: scan of the tag lengths (all but the last are <= 64) ---
        const uint32_t x = is_tag ? take : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int dlt = 1; dlt < SNP_WARP; dlt <<= 1) {
            const uint32_t y = __shfl_up_sync(SNP_FULL, incl, dlt);
            if (lane >= (unsigned)dlt) incl += y;
        }
        const uint32_t dst = op + (incl - x);

        // ---- validation in stream order (SnappyDecompressor.cs:570-573,598-606) ------
        const bool bad_off = is_tag && !is_lit && (off - 1u >= dst);  // off == 0 || off > dst
        const bool too_long = is_tag && take > U - dst;
        const unsigned errs = __ballot_sync(SNP_FULL, bad_off || too_long);
        if (errs) {
            const int err = bad_off ? SNP_INVALID_COPY_OFFSET : SNP_DATA_TOO_LONG;
            return __shfl_sync(SNP_FULL, err, __ffs(errs) - 1);
        }

        // ---- QUEUE append --------------------------------------------------------------
        if (is_tag) {
            const uint32_t slot = (tail + __popc(tags & lt)) & (SNP_QCAP - 1);
            q->dst[slot] = is_lit ? dst : (dst | 0x80000000u);
            q->src[slot] = is_lit ? pos + hdr : off;
        }
        tail += __popc(tags);
        op += __shfl_sync(SNP_FULL, incl, 31);
        if (lane == 0) q->dst[tail & (SNP_QCAP - 1)] = op;  // sentinel
        __syncwarp();
        ip = ip_next;

        drain(SNP_WARP - 1);  // keep < 32 bytes (hence < 32 tags) queued
    }
    drain(0);

    if (op < U) return SNP_INCOMPLETE;  // Snappy.cs:178-181
    *written = op;
    return SNP_OK;
}

// Persistent launch: one CTA slot per (SM x resident CTA); every warp pulls the next block
// index from a global counter, so cheap (incompressible) and expensive (text) blocks balance
// across warps instead of leaving warp slots idle until the slowest warp of a CTA retires.
#ifndef SNP_EMU
__global__ void __launch_bounds__(256, SNP_V3_CTAS)
k_decompress_v3(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_len, uint8_t *out_base,
                const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
                uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items,
                unsigned long long *__restrict__ next_item) {
    __shared__ uint32_t lut[256];
    __shared__ WarpQueue3 queues[8];
    lut[threadIdx.x & 255] = tag_lut3_entry(threadIdx.x & 255);
    __syncthreads();
    const unsigned lane = lane_id();
    WarpQueue3 *q = &queues[threadIdx.x / SNP_WARP];
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
            st = decompress_block_v3(in_base + in_off[item], in_len[item], out_base + out_off[item],
                                     out_cap[item], &w, lut, q);
        if (lane == 0) {
            out_len[item] = w;
            status[item] = st;
        }
        __syncwarp();
    }
}

#endif  // !SNP_EMU

}  // namespace snp
