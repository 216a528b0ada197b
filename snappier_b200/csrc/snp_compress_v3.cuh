// snp_compress_v3.cuh -- warp-parallel, bit-exact batched Snappy fragment compressor (sm_100a).
//
// One warp per <= 64 KiB fragment; the hash table is a per-warp slice of a device buffer served from L2.
// The reference's probe loop is strictly sequential (table state depends on probe
// order), but the probe POSITIONS of a literal run are data-independent: skip
// starts at 32, stride = skip >> 5, skip += stride (SnappyCompressor.cs:227,
// 319-320).  So the 32 lanes evaluate 32 successive probes of the run at once:
//
//   lane l: x = LE32(in[p_l]); h = hash(x)
//           candidate = position of the nearest lower lane with the same bucket
//                       (MATCH.ANY), else table[h]      -- what the sequential loop
//                       would have read after the earlier probes' writes
//           hit_l = LE32(in[candidate]) == x
//   first event (ballot) = first hit, or first probe whose successor passes
//   ip_limit (SnappyCompressor.cs:323-327); table writes are committed for the
//   lanes up to and including the hit only, highest lane per bucket winning --
//   exactly the table the sequential loop leaves behind.
//
// After a match, the reference re-probes at ip before starting the next literal
// run (SnappyCompressor.cs:393-398).  That re-probe rides as lane 0 of the next
// batch, so one batch = one dependent memory round trip per emitted copy.
//
// Bit-exactness is checked against oracle/snappy_oracle.c (both hash modes) and,
// for the MUL hash, against the reference's golden chunks (tests/).
#pragma once
#include "snp_common.cuh"
#include "snp_compress_v1.cuh"  // OutCursor, emit_literal_v1, emit_copy_v1

namespace snp {

#define SNP_SCHED_LEN 288  // probes needed to cross 64 KiB: 266

// Probe schedule of a literal run: entry k = offset of probe k from the first probe
// (bits 0..19) | stride of probe k (bits 20..31).  Data-independent.
__device__ uint32_t g_probe_sched[SNP_SCHED_LEN];

#ifndef SNP_EMU
__global__ void k_init_probe_sched() {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t skip = 32, off = 0;
        for (int k = 0; k < SNP_SCHED_LEN; k++) {
            uint32_t stride = skip >> 5;
            g_probe_sched[k] = min(off, 0xfffffu) | (min(stride, 0xfffu) << 20);
            off += stride;
            skip += stride;
        }
    }
}
#endif  // !SNP_EMU

// SnappyCompressor.cs:562-688 -- bounded common prefix of in[s1..] and in[s2..n), 128 bytes per ballot.
__device__ __forceinline__ uint32_t find_match_length_v2(const uint8_t *__restrict__ in, uint32_t s1, uint32_t s2,
                                                         uint32_t n, unsigned lane) {
    uint32_t base = 0;
    for (;;) {
        const uint32_t i = base + 4 * lane;
        const uint32_t q = s2 + i;
        uint32_t matched;  // matching bytes in this lane's 4-byte window, limited by the end of input
        if (q + 4 <= n) {
            const uint32_t x = ld_le32(in + s1 + i) ^ ld_le32(in + q);
            matched = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
        } else {
            matched = 0;
            while (q + matched < n && in[s1 + i + matched] == in[q + matched]) matched++;
        }
        const unsigned part = __ballot_sync(SNP_FULL, matched < 4);
        if (part) {
            const int L = __ffs(part) - 1;
            return base + 4 * L + __shfl_sync(SNP_FULL, matched, L);
        }
        base += 4 * SNP_WARP;
    }
}

// ---- k_compress_v3: hash tables in global memory (L2) -----------------------------------------
// One slice per warp of the launch, so the grid runs at full occupancy (a kernel with the tables in
// shared memory is capped at 7 blocks in flight per SM and was 3.5-6x slower: DESIGN.md 4.2).  Entries are
// widened to 32 bits: low 16 = position (the reference's ushort), high 16 = a fingerprint of the
// 4 bytes at that position.  The fingerprint is a pure filter -- equal bytes imply equal
// fingerprints -- that spares the candidate load (a random DRAM sector) for the ~all probes
// that cannot match; surviving candidates are still verified against the real bytes.
__device__ __forceinline__ uint32_t fp16(uint32_t x) { return (x * 0x9E3779B1u) >> 16; }

// FP = false: the reference's plain 16-bit entries (32 KiB per table, no fingerprint filter): half the table footprint
// in L2 at the price of one candidate load per live probe.
template <int HASH_MODE, bool FP = true>
__device__ __noinline__ void compress_fragment_v3(const uint8_t *__restrict__ in, uint32_t n, OutCursor &o,
                                                  uint32_t *table, const uint16_t *lut, const uint32_t *sched, uint32_t w0) {
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int tsize = table_size_for(n);
    const uint32_t mask = 2u * (uint32_t)(tsize - 1);
    uint32_t next_emit = 0;
    if (n >= 15) {  // Constants.InputMarginBytes, SnappyCompressor.cs:190
        {  // HashTable.cs:52: "zero" = position 0, whose bytes are in[0..3]
            const uint32_t e0 = FP ? fp16(ld_le32(in)) << 16 : 0u;
            uint4 z = make_uint4(e0, e0, e0, e0);
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
            for (int i = lane; i < tsize / (FP ? 4 : 8); i += SNP_WARP) __stcg(t4 + i, z);
            __syncwarp();
        }
        const uint32_t ip_limit = n - 15;
        bool reprobe = false;
        uint32_t kb = 0;
        // Probes tried per batch.  Right after a match the next hit is usually a few positions away (dense-match
        // data: text), and every probe of a batch costs a table sector from HBM whether or not the batch gets that
        // far -- ncu on text: 11 MB of DRAM reads per 64 KiB block -- so the first batch after a match is w0 wide
        // and only a miss widens it to 32.  Pure scheduling: the table sees the same sequence of reads and writes.
        // Measured (profiles/r01_compress_width.log): w0 = 16 gives +14 % on config 3, +8 % on text, +5 % on the mix.
        uint32_t width = SNP_WARP;
        for (;;) {
            uint32_t p, nip;
            bool term = false;
            if (reprobe && lane == 0) {
                p = next_emit;
                nip = p;
            } else {
                const uint32_t k = kb + lane - (reprobe ? 1u : 0u);
                const uint32_t s = sched[min(k, (uint32_t)SNP_SCHED_LEN - 1)];
                p = next_emit + 1 + (s & 0xfffffu);
                nip = p + (s >> 20);
                term = nip > ip_limit || k >= SNP_SCHED_LEN;  // :323-327
            }
            const unsigned wmask = 0xffffffffu >> (SNP_WARP - width);
            const unsigned terms = __ballot_sync(SNP_FULL, term) & wmask;
            const unsigned live = (terms ? ((1u << (__ffs(terms) - 1)) - 1u) : SNP_FULL) & wmask;
            const bool is_live = (live >> lane) & 1;
            const uint32_t x = is_live ? ld_le32(in + p) : 0u;
            const uint32_t h = is_live ? (table_hash<HASH_MODE>(x, mask, lut) >> 1) : (0x10000u + lane);
            const unsigned same = __match_any_sync(SNP_FULL, h);
            const unsigned lower = same & lt;
            const int src = lower ? 31 - __clz(lower) : (int)lane;
            const uint32_t p_src = __shfl_sync(SNP_FULL, p, src);
            const uint32_t x_src = __shfl_sync(SNP_FULL, x, src);
            uint32_t cand = 0;
            bool hit = false;
            if (is_live) {
                if (lower) {  // an earlier probe of this batch owns the bucket: its bytes are in a register
                    cand = p_src;
                    hit = x_src == x;
                } else {
                    if (FP) {
                        const uint32_t e = __ldcg(table + h);
                        cand = e & 0xffffu;
                        if ((e >> 16) == fp16(x)) hit = ld_le32(in + cand) == x;
                    } else {
                        cand = __ldcg(reinterpret_cast<const uint16_t *>(table) + h);
                        hit = ld_le32(in + cand) == x;
                    }
                }
            }
            const unsigned hits = __ballot_sync(SNP_FULL, hit);
            const int f = __ffs(hits) - 1;
            const unsigned commit = hits ? (live & (0xffffffffu >> (31 - f))) : live;
            if (((commit >> lane) & 1) && (same & commit & ~lt & ~(1u << lane)) == 0) {
                if (FP) __stcg(table + h, p | (fp16(x) << 16));
                else __stcg(reinterpret_cast<uint16_t *>(table) + h, (uint16_t)p);
            }
            __syncwarp();
            if (!hits) {
                if (terms) break;
                kb += width - (reprobe ? 1u : 0u);
                reprobe = false;
                width = SNP_WARP;
                continue;
            }
            uint32_t ip = __shfl_sync(SNP_FULL, p, f);
            const uint32_t c = __shfl_sync(SNP_FULL, cand, f);
            if (ip > next_emit) emit_literal_v1(o, in + next_emit, ip - next_emit, lane);
            const uint32_t m = 4 + find_match_length_v2(in, c + 4, ip + 4, n, lane);
            emit_copy_v1(o, ip - c, m, lane);
            ip += m;
            next_emit = ip;
            if (ip >= ip_limit) break;  // :381-384
            if (lane == 0) {            // :393-394
                const uint32_t x1 = ld_le32(in + ip - 1);
                const uint32_t h1 = table_hash<HASH_MODE>(x1, mask, lut) >> 1;
                if (FP) __stcg(table + h1, (ip - 1) | (fp16(x1) << 16));
                else __stcg(reinterpret_cast<uint16_t *>(table) + h1, (uint16_t)(ip - 1));
            }
            __syncwarp();
            reprobe = true;
            kb = 0;
            width = w0;
        }
    }
    if (next_emit < n) emit_literal_v1(o, in + next_emit, n - next_emit, lane);  // :406-411
}

#ifndef SNP_EMU
template <int HASH_MODE, int VARIANT = 3>
__global__ void __launch_bounds__(256)
k_compress_v3(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
              const uint32_t *__restrict__ in_len, uint8_t *out_base,
              const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
              uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items, int frag_mode,
              unsigned long long *__restrict__ next_item, uint32_t *__restrict__ tables) {
    __shared__ uint16_t lut[1024];
    __shared__ uint32_t sched[SNP_SCHED_LEN];
    if (HASH_MODE == SNP_HASH_CRC32C) build_crc_lut(lut, threadIdx.x, blockDim.x);
    for (unsigned i = threadIdx.x; i < SNP_SCHED_LEN; i += blockDim.x) sched[i] = g_probe_sched[i];
    __syncthreads();
    const unsigned warps = blockDim.x / SNP_WARP;
    const unsigned warp = threadIdx.x / SNP_WARP;
    const unsigned lane = lane_id();
    uint32_t *table = tables + ((size_t)blockIdx.x * warps + warp) * 16384;
    const uint32_t w0 = min(max((uint32_t)frag_mode >> 8, 1u), (uint32_t)SNP_WARP);  // first-batch width (bits 8..)
    frag_mode &= 1;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1ull);
        item = __shfl_sync(SNP_FULL, item, 0);
        if (item >= n_items) break;
        const uint8_t *in = in_base + in_off[item];
        uint32_t n = in_len[item];
        OutCursor o{out_base + out_off[item], out_cap[item], 0};
        int st = SNP_OK;
        if (n > SNP_BLOCK_SIZE) {
            st = SNP_E_INVALID_ARG;
        } else {
            if (!frag_mode) {  // SnappyCompressor.cs:34-38
                uint32_t lo, hi;
                int need = varint_encode(n, &lo, &hi);
                if ((int)lane < need) o.put(lane, (uint8_t)(lane < 4 ? lo >> (8 * lane) : hi));
                o.pos = need;
            }
            if (n > 0) {
                if (VARIANT == 6) compress_fragment_v3<HASH_MODE, false>(in, n, o, table, lut, sched, w0);
                else compress_fragment_v3<HASH_MODE>(in, n, o, table, lut, sched, w0);
            }
            if (o.pos > o.cap) st = SNP_OUTPUT_TOO_SMALL;  // SnappyCompressor.cs:63-68
        }
        if (lane == 0) {
            out_len[item] = st == SNP_OK ? o.pos : 0;
            status[item] = st;
        }
        __syncwarp();
    }
}

#endif  // !SNP_EMU

}  // namespace snp
