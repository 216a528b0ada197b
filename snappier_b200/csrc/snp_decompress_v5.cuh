// snp_decompress_v5.cuh -- v3 plus a SPARSE-TAG PREFIX engine (sm_100a).
//
// ncu on v3 shows the kernel issue-bound, and on streams of long tags (incompressible data =
// one 64 KiB literal; record-like data = runs of 64-byte copies) almost all of that issue
// goes into the 32-position speculative parse, which finds one or two tags per window.
// v5 starts every block in a sparse mode: decode the single tag at ip (warp-uniform) and copy
// it cooperatively -- literals of >= 128 bytes as aligned 16-byte vectors -- for as long as
// tags are >= SNP_SPARSE_MIN bytes; at a run of SNP_SPARSE_PATIENCE shorter tags it hands (ip, op) to the dense
// engine of v3 (verbatim copy, kept in its own non-inlined function so that its register
// allocation is untouched) which finishes the block.  Dense blocks pay one extra tag decode.
//
// Semantics: identical to v3 / v1 / oracle (same end-of-input, partial-literal, validation
// order); A/B-tested in tests/test_gpu_parity.py.
#pragma once
#include "snp_common.cuh"
#include "snp_decompress_v1.cuh"
#include "snp_decompress_v3.cuh"

namespace snp {

#define SNP_SPARSE_MIN 16u
#define SNP_SPARSE_PATIENCE 3u  // an isolated short tag is decoded in place, a run of them is not

// Cooperative literal copy of `len` bytes; >= 128 bytes go as aligned 16-byte vectors.
__device__ __forceinline__ void copy_literal_wide(const uint8_t *__restrict__ s, uint8_t *d, uint32_t len,
                                                  const uint8_t *in_end, unsigned lane) {
    if (len >= 128) {
        const uint32_t h = (uint32_t)(-(intptr_t)d) & 15u;  // bytes up to the first 16-byte boundary of d
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
            const uint4 B = *(pb <= last ? pb : last);  // only read when sb != 0 needs it; clamped at the buffer end
            const uint32_t W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
            uint4 r;
            switch (wo) {  // warp-uniform
                case 0: r = make_uint4(__funnelshift_r(W[0], W[1], bs), __funnelshift_r(W[1], W[2], bs),
                                       __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs)); break;
                case 1: r = make_uint4(__funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs),
                                       __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs)); break;
                case 2: r = make_uint4(__funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs),
                                       __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs)); break;
                default: r = make_uint4(__funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs),
                                        __funnelshift_r(W[5], W[6], bs), __funnelshift_r(W[6], W[7], bs)); break;
            }
            dv[v] = r;
        }
        const uint32_t done = h + (nvec << 4);
        if (done + lane < len) d[done + lane] = s[done + lane];  // < 16 tail bytes
    } else {
        for (uint32_t k = lane; k < len; k += SNP_WARP) d[k] = s[k];
    }
}

struct SparseResult {
    int status;
    uint32_t ip, op;
    bool done;  // the tag stream ended inside the sparse engine
};

// Warp-uniform tag-at-a-time decode while tags are >= SNP_SPARSE_MIN bytes.
__device__ __noinline__ SparseResult sparse_run_v5(const uint8_t *__restrict__ in, uint32_t n_in, uint8_t *out,
                                                   uint32_t U, const uint32_t *lut, uint32_t ip, uint32_t op) {
    const unsigned lane = lane_id();
    const uint32_t skew = (uint32_t)((uintptr_t)in & 3);
    const uint32_t *in_w = (const uint32_t *)((uintptr_t)in - skew);
    const uint32_t last_w = (skew + n_in - 1) >> 2;
    SparseResult r{SNP_OK, ip, op, true};
    uint32_t short_run = 0;  // consecutive tags below SNP_SPARSE_MIN
    while (r.ip < n_in) {
        const uint32_t bo = skew + r.ip;
        const uint32_t wi = bo >> 2;
        const unsigned sh = (bo & 3) * 8;
        const uint32_t w0 = in_w[min(wi, last_w)];
        const uint32_t w1 = in_w[min(wi + 1, last_w)];
        const uint32_t v = __funnelshift_r(w0, w1, sh);
        const uint32_t trailer = __funnelshift_r(v, w1 >> sh, 8);
        const uint32_t ent = lut[v & 0xff];
        const uint32_t hdr = __byte_perm(ent, 0, 0x4441);
        const bool is_lit = (int32_t)ent < 0;
        const uint32_t tval = trailer & __funnelshift_rc(0xffffffffu, 0u, __byte_perm(ent, 0, 0x4442));
        uint32_t len = ent & 0xff;
        if (ent & 0x40000000u) len = max(tval + 1, tval);
        const uint32_t off = ((ent >> 16) & 0x700u) | tval;
        const uint32_t left = n_in - r.ip;
        if (left < hdr) break;  // truncated tag: parsing stops (SnappyDecompressor.cs:464-483)
        const uint32_t avail = left - hdr;
        const bool partial = is_lit && len > avail;
        const uint32_t take = partial ? avail : len;
        if (take < SNP_SPARSE_MIN && !partial) {
            if (++short_run >= SNP_SPARSE_PATIENCE) {  // dense tags from here on: hand over to the dense engine
                r.done = false;
                break;
            }
        } else {
            short_run = 0;
        }
        if (!is_lit && (off - 1u >= r.op)) {  // off == 0 || off > produced (SnappyDecompressor.cs:598-601)
            r.status = SNP_INVALID_COPY_OFFSET;
            break;
        }
        if (take > U - r.op) {  // :570-573, :603-606
            r.status = SNP_DATA_TOO_LONG;
            break;
        }
        if (is_lit) {
            copy_literal_wide(in + r.ip + hdr, out + r.op, take, in + n_in, lane);
            __syncwarp();
        } else if (take) {
            copy_long_tag3(in, out, r.op, take, true, off, lane);
        }
        r.op += take;
        r.ip += hdr + (is_lit ? take : 0u);
        if (partial) break;
    }
    return r;
}

// The dense-tag engine of v3, entered at stream position ip0 with op0 bytes already produced.
__device__ __noinline__ int decompress_dense_v5(const uint8_t *__restrict__ in, uint32_t n_in, uint8_t *out,
                                                uint32_t U, uint32_t ip0, uint32_t op0, uint32_t *written,
                                                const uint32_t *lut, WarpQueue3 *q) {
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    *written = 0;

    // 4-byte-aligned view of the input: stream byte p lives at byte (skew + p) of in_w[]
    const uint32_t skew = (uint32_t)((uintptr_t)in & 3);
    const uint32_t *in_w = (const uint32_t *)((uintptr_t)in - skew);
    const uint32_t last_w = (skew + n_in - 1) >> 2;  // no word beyond this index is touched

    uint32_t ip = ip0, op = op0, cur = op0, head = 0, tail = 0;
    bool stop = false;
    if (lane == 0) q->dst[0] = op0;
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

// One block through the sparse prefix engine and, if the tags turn dense, the dense engine.
__device__ __forceinline__ int decompress_block_v5(const uint8_t *in, uint32_t n_in, uint8_t *out, uint32_t cap,
                                                   uint32_t *written, const uint32_t *lut, WarpQueue3 *q) {
    uint32_t w = 0;
    int st;
    if (n_in >= 0x7fff0000u) {  // stream offsets are 32-bit with headroom here; v1 is safe to 2^32-1
        st = decompress_block_v1(in, n_in, out, cap, &w);
    } else {
        uint32_t U, used;
        st = varint_read(in, n_in, &U, &used);  // SnappyDecompressor.cs:50-63
        if (st == SNP_OK && U > 0x7fffffffu) st = SNP_INVALID_LENGTH;
        if (st == SNP_OK && cap < U) st = SNP_OUTPUT_TOO_SMALL;
        if (st == SNP_OK && U != 0) {
            const SparseResult r = sparse_run_v5(in, n_in, out, U, lut, used, 0);
            if (r.status != SNP_OK) {
                st = r.status;
            } else if (r.done) {
                st = r.op < U ? SNP_INCOMPLETE : SNP_OK;  // Snappy.cs:178-181
                w = st == SNP_OK ? r.op : 0;
            } else {
                st = decompress_dense_v5(in, n_in, out, U, r.ip, r.op, &w, lut, q);
            }
        }
    }
    *written = w;
    return st;
}

#ifndef SNP_EMU
__global__ void __launch_bounds__(256, SNP_V3_CTAS)
k_decompress_v5(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
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
        const int st = decompress_block_v5(in_base + in_off[item], in_len[item], out_base + out_off[item], out_cap[item],
                                           &w, lut, q);
        if (lane == 0) {
            out_len[item] = w;
            status[item] = st;
        }
        __syncwarp();
    }
}
#endif  // !SNP_EMU

}  // namespace snp
