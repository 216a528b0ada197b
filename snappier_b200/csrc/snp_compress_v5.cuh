// snp_compress_v5.cuh -- k_compress_v3's algorithm on LANE GROUPS of W = 16 lanes (sm_100a).
//
// ncu on dense-match data (text, profiles/r01_compress_width.log and DESIGN.md 4.2): the compressor is bound by
// latency, not by bandwidth or issue slots -- every match is a chain of ~4 dependent memory round trips (probe
// words, table sector from HBM, candidate bytes, match extension), 53 % issue utilisation, throughput proportional
// to the number of resident warps, and 64 warps per SM is the hardware limit.  The only parallelism left is across
// blocks, so this variant runs TWO blocks per warp: each half-warp owns one block and executes the same batch
// algorithm with 16-wide batches and half-warp collectives (independent thread scheduling lets the two halves
// follow their own control flow; their memory round trips overlap).  The table sees exactly the sequence of reads
// and writes of the sequential loop, as in k_compress_v3: the batch width is pure scheduling.
//
// Semantics: SnappyCompressor.cs:174-415 + HashTable.cs; bit-exact with k_compress_v1/v3 and the oracle
// (tests/test_gpu_parity.py A/B).  Selected with SNP_COMP_KERNEL=5.
#pragma once
#include "snp_common.cuh"
#include "snp_compress_v2.cuh"  // OutCursor, fp16, g_probe_sched, SNP_SCHED_LEN

namespace snp {

template <int W>
struct LaneGroup {
    unsigned gl;    // lane within the group
    unsigned base;  // first lane of the group
    unsigned mask;  // the group's lanes
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> base) & ((1ull << W) - 1); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, unsigned src) const { return __shfl_sync(mask, v, base + src); }
    __device__ __forceinline__ unsigned match_any(uint32_t v) const { return (__match_any_sync(mask, v) >> base) & ((1ull << W) - 1); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

template <int W>
__device__ __forceinline__ void emit_literal_g(const LaneGroup<W> &g, OutCursor &o, const uint8_t *__restrict__ lit, uint32_t len) {
    const uint32_t n = len - 1;  // SnappyCompressor.cs:418-464
    uint32_t hdr;
    if (n < 60) {
        if (g.gl == 0) o.put(o.pos, (uint8_t)(n << 2));
        hdr = 1;
    } else {
        const uint32_t count = ((31 - __clz(n)) >> 3) + 1;
        if (g.gl == 0) o.put(o.pos, (uint8_t)((59 + count) << 2));
        if (g.gl < count) o.put(o.pos + 1 + g.gl, (uint8_t)(n >> (8 * g.gl)));
        hdr = 1 + count;
    }
    const uint32_t at = o.pos + hdr;
    for (uint32_t k = g.gl; k < len; k += W) o.put(at + k, lit[k]);
    o.pos = at + len;
}

template <int W>
__device__ __forceinline__ void emit_copy_upto64_g(const LaneGroup<W> &g, OutCursor &o, uint32_t offset, uint32_t len) {
    if (len < 12 && offset < 2048) {  // SnappyCompressor.cs:467-505
        if (g.gl == 0) {
            o.put(o.pos, (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 8) << 5)));
            o.put(o.pos + 1, (uint8_t)offset);
        }
        o.pos += 2;
    } else {
        if (g.gl == 0) {
            o.put(o.pos, (uint8_t)(2 + ((len - 1) << 2)));
            o.put(o.pos + 1, (uint8_t)offset);
            o.put(o.pos + 2, (uint8_t)(offset >> 8));
        }
        o.pos += 3;
    }
}

template <int W>
__device__ __forceinline__ void emit_copy_g(const LaneGroup<W> &g, OutCursor &o, uint32_t offset, uint32_t len) {
    while (len >= 68) {  // SnappyCompressor.cs:507-543
        emit_copy_upto64_g(g, o, offset, 64);
        len -= 64;
    }
    if (len > 64) {
        emit_copy_upto64_g(g, o, offset, 60);
        len -= 60;
    }
    emit_copy_upto64_g(g, o, offset, len);
}

// SnappyCompressor.cs:562-688: bounded common prefix of in[s1..] and in[s2..n), 4 * W bytes per ballot.
template <int W>
__device__ __forceinline__ uint32_t find_match_length_g(const LaneGroup<W> &g, const uint8_t *__restrict__ in, uint32_t s1,
                                                        uint32_t s2, uint32_t n) {
    uint32_t base = 0;
    for (;;) {
        const uint32_t i = base + 4 * g.gl;
        const uint32_t q = s2 + i;
        uint32_t matched;
        if (q + 4 <= n) {
            const uint32_t x = ld_le32(in + s1 + i) ^ ld_le32(in + q);
            matched = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
        } else {
            matched = 0;
            while (q + matched < n && in[s1 + i + matched] == in[q + matched]) matched++;
        }
        const unsigned part = g.ballot(matched < 4);
        if (part) {
            const int L = __ffs(part) - 1;
            return base + 4 * L + g.shfl(matched, L);
        }
        base += 4 * W;
    }
}

template <int HASH_MODE, int W>
__device__ __noinline__ void compress_fragment_g(const LaneGroup<W> &g, const uint8_t *__restrict__ in, uint32_t n, OutCursor &o,
                                                 uint32_t *table, const uint16_t *lut, const uint32_t *sched, uint32_t w0) {
    const unsigned lt = (1u << g.gl) - 1u;
    const unsigned all = (unsigned)((1ull << W) - 1);
    const int tsize = table_size_for(n);
    const uint32_t mask = 2u * (uint32_t)(tsize - 1);
    uint32_t next_emit = 0;
    if (n >= 15) {  // Constants.InputMarginBytes, SnappyCompressor.cs:190
        {  // HashTable.cs:52: "zero" = position 0, whose bytes are in[0..3]
            const uint32_t e0 = fp16(ld_le32(in)) << 16;
            const uint4 z = make_uint4(e0, e0, e0, e0);
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
            for (int i = g.gl; i < tsize / 4; i += W) __stcg(t4 + i, z);
            g.sync();
        }
        const uint32_t ip_limit = n - 15;
        bool reprobe = false;  // lane 0 of the batch is the post-match probe at next_emit (:393-398)
        uint32_t kb = 0;       // schedule index of the first run probe in this batch
        uint32_t width = W;    // probes tried in this batch (w0 right after a match, see k_compress_v3)
        for (;;) {
            uint32_t p, nip;
            bool term = false;
            if (reprobe && g.gl == 0) {
                p = next_emit;
                nip = p;
            } else {
                const uint32_t k = kb + g.gl - (reprobe ? 1u : 0u);
                const uint32_t s = sched[min(k, (uint32_t)SNP_SCHED_LEN - 1)];
                p = next_emit + 1 + (s & 0xfffffu);
                nip = p + (s >> 20);
                term = nip > ip_limit || k >= SNP_SCHED_LEN;  // :323-327 (checked before the table is touched)
            }
            const unsigned wmask = all >> (W - width);
            const unsigned terms = g.ballot(term) & wmask;
            const unsigned live = (terms ? ((1u << (__ffs(terms) - 1)) - 1u) : all) & wmask;
            const bool is_live = (live >> g.gl) & 1;
            const uint32_t x = is_live ? ld_le32(in + p) : 0u;
            const uint32_t h = is_live ? (table_hash<HASH_MODE>(x, mask, lut) >> 1) : (0x10000u + g.gl);
            // what the sequential loop would have read from table[h]: the nearest earlier probe of this batch with
            // the same bucket supersedes the table
            const unsigned same = g.match_any(h);
            const unsigned lower = same & lt;
            const int src = lower ? 31 - __clz(lower) : (int)g.gl;
            const uint32_t p_src = g.shfl(p, src);
            const uint32_t x_src = g.shfl(x, src);
            uint32_t cand = 0;
            bool hit = false;
            if (is_live) {
                if (lower) {
                    cand = p_src;
                    hit = x_src == x;
                } else {
                    const uint32_t e = __ldcg(table + h);
                    cand = e & 0xffffu;
                    if ((e >> 16) == fp16(x)) hit = ld_le32(in + cand) == x;
                }
            }
            const unsigned hits = g.ballot(hit);
            const int f = __ffs(hits) - 1;
            const unsigned commit = hits ? (live & (all >> (W - 1 - f))) : live;
            if (((commit >> g.gl) & 1) && (same & commit & ~lt & ~(1u << g.gl)) == 0)
                __stcg(table + h, p | (fp16(x) << 16));  // the highest committed lane per bucket wins
            g.sync();
            if (!hits) {
                if (terms) break;
                kb += width - (reprobe ? 1u : 0u);
                reprobe = false;
                width = W;
                continue;
            }
            uint32_t ip = g.shfl(p, f);
            const uint32_t c = g.shfl(cand, f);
            if (ip > next_emit) emit_literal_g(g, o, in + next_emit, ip - next_emit);
            const uint32_t m = 4 + find_match_length_g(g, in, c + 4, ip + 4, n);
            emit_copy_g(g, o, ip - c, m);
            ip += m;
            next_emit = ip;
            if (ip >= ip_limit) break;  // :381-384
            if (g.gl == 0) {            // :393-394
                const uint32_t x1 = ld_le32(in + ip - 1);
                __stcg(table + (table_hash<HASH_MODE>(x1, mask, lut) >> 1), (ip - 1) | (fp16(x1) << 16));
            }
            g.sync();
            reprobe = true;
            kb = 0;
            width = min(w0, (uint32_t)W);
        }
    }
    if (next_emit < n) emit_literal_g(g, o, in + next_emit, n - next_emit);  // :406-411
}

template <int HASH_MODE, int W>
__global__ void __launch_bounds__(256, 8)
k_compress_v5(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
              const uint32_t *__restrict__ in_len, uint8_t *out_base,
              const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
              uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items, int frag_mode,
              unsigned long long *__restrict__ next_item, uint32_t *__restrict__ tables) {
    __shared__ uint16_t lut[1024];
    __shared__ uint32_t sched[SNP_SCHED_LEN];
    if (HASH_MODE == SNP_HASH_CRC32C) build_crc_lut(lut, threadIdx.x, blockDim.x);
    for (unsigned i = threadIdx.x; i < SNP_SCHED_LEN; i += blockDim.x) sched[i] = g_probe_sched[i];
    __syncthreads();
    const unsigned lane = lane_id();
    LaneGroup<W> g;
    g.gl = lane % W;
    g.base = lane - g.gl;
    g.mask = (unsigned)(((1ull << W) - 1) << g.base);
    const unsigned groups = blockDim.x / W;
    uint32_t *table = tables + ((size_t)blockIdx.x * groups + threadIdx.x / W) * 16384;
    const uint32_t w0 = min(max(((uint32_t)frag_mode >> 8) & 0xffu, 1u), (uint32_t)W);  // first-batch width (bits 8..15)
    frag_mode &= 1;

    for (;;) {
        unsigned long long item = 0;
        if (g.gl == 0) item = atomicAdd(next_item, 1ull);
        item = __shfl_sync(g.mask, item, g.base);
        if (item >= n_items) break;
        const uint8_t *in = in_base + in_off[item];
        const uint32_t n = in_len[item];
        OutCursor o{out_base + out_off[item], out_cap[item], 0};
        int st = SNP_OK;
        if (n > SNP_BLOCK_SIZE) {
            st = SNP_E_INVALID_ARG;
        } else {
            if (!frag_mode) {  // SnappyCompressor.cs:34-38
                uint32_t lo, hi;
                const int need = varint_encode(n, &lo, &hi);
                if ((int)g.gl < need) o.put(g.gl, (uint8_t)(g.gl < 4 ? lo >> (8 * g.gl) : hi));
                o.pos = need;
            }
            if (n > 0) compress_fragment_g<HASH_MODE, W>(g, in, n, o, table, lut, sched, w0);
            if (o.pos > o.cap) st = SNP_OUTPUT_TOO_SMALL;  // SnappyCompressor.cs:63-68
        }
        if (g.gl == 0) {
            out_len[item] = st == SNP_OK ? o.pos : 0;
            status[item] = st;
        }
        g.sync();
    }
}

}  // namespace snp
