// snp_compress_v1.cuh -- baseline batched Snappy fragment compressor.
//
// One warp per <= 64 KiB fragment, hash table (u16[<=16384] = 32 KiB) in shared
// memory.  The probe sequence of the reference is emulated one probe at a time
// with warp-uniform control flow, which makes bit-exactness easy to see; literal
// copies and match extension are spread over the lanes.  The faster kernels are
// A/B-checked against this one (SNP_COMP_KERNEL=v1 selects it).
//
// Semantics restated from /root/reference/Snappier/Internal/SnappyCompressor.cs
// :24-83 (TryCompress), :174-415 (CompressFragment), :418-543 (emit), :562-688
// (FindMatchLength); HashTable.cs:38-71,91-126.
#pragma once
#include "snp_common.cuh"

namespace snp {

// Bounded output cursor: bytes at positions >= cap are dropped (the item then
// reports SNP_OUTPUT_TOO_SMALL), so an undersized slot can never be overrun.
struct OutCursor {
    uint8_t *base;
    uint32_t cap;
    uint32_t pos;
    __device__ __forceinline__ void put(uint32_t at, uint8_t b) const {
        if (at < cap) base[at] = b;
    }
};

// SnappyCompressor.cs:418-464 (EmitLiteralFast/Slow produce the same bytes).
__device__ __forceinline__ void emit_literal_v1(OutCursor &o, const uint8_t *__restrict__ lit,
                                                uint32_t len, unsigned lane) {
    uint32_t n = len - 1;
    uint32_t hdr;
    if (n < 60) {
        if (lane == 0) o.put(o.pos, (uint8_t)(n << 2));
        hdr = 1;
    } else {
        uint32_t count = ((31 - __clz(n)) >> 3) + 1;
        if (lane == 0) o.put(o.pos, (uint8_t)((59 + count) << 2));
        if (lane < count) o.put(o.pos + 1 + lane, (uint8_t)(n >> (8 * lane)));
        hdr = 1 + count;
    }
    uint32_t at = o.pos + hdr;
    for (uint32_t k = lane; k < len; k += SNP_WARP) o.put(at + k, lit[k]);
    o.pos = at + len;
}

// SnappyCompressor.cs:467-505: one copy element of 4..64 bytes.
__device__ __forceinline__ void emit_copy_upto64_v1(OutCursor &o, uint32_t offset, uint32_t len,
                                                    unsigned lane) {
    if (len < 12 && offset < 2048) {
        if (lane == 0) {
            o.put(o.pos, (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 8) << 5)));
            o.put(o.pos + 1, (uint8_t)offset);
        }
        o.pos += 2;
    } else {
        if (lane == 0) {
            o.put(o.pos, (uint8_t)(2 + ((len - 1) << 2)));
            o.put(o.pos + 1, (uint8_t)offset);
            o.put(o.pos + 2, (uint8_t)(offset >> 8));
        }
        o.pos += 3;
    }
}

// SnappyCompressor.cs:507-543.
__device__ __forceinline__ void emit_copy_v1(OutCursor &o, uint32_t offset, uint32_t len, unsigned lane) {
    while (len >= 68) {
        emit_copy_upto64_v1(o, offset, 64, lane);
        len -= 64;
    }
    if (len > 64) {
        emit_copy_upto64_v1(o, offset, 60, lane);
        len -= 60;
    }
    emit_copy_upto64_v1(o, offset, len, lane);
}

// SnappyCompressor.cs:562-688 -- bounded common prefix, 32 bytes per ballot.
__device__ __forceinline__ uint32_t find_match_length_v1(const uint8_t *__restrict__ in, uint32_t s1,
                                                         uint32_t s2, uint32_t n, unsigned lane) {
    uint32_t m = 0;
    for (;;) {
        uint32_t idx = m + lane;
        bool ok = (s2 + idx < n) && in[s1 + idx] == in[s2 + idx];
        unsigned bad = __ballot_sync(SNP_FULL, !ok);
        if (bad) return m + (__ffs(bad) - 1);
        m += SNP_WARP;
    }
}

// SnappyCompressor.cs:174-415.  `table` = this warp's 32 KiB of shared memory.
template <int HASH_MODE>
__device__ __noinline__ void compress_fragment_v1(const uint8_t *__restrict__ in, uint32_t n, OutCursor &o,
                                                  uint16_t *table, const uint16_t *lut) {
    const unsigned lane = lane_id();
    const int tsize = table_size_for(n);
    {  // HashTable.cs:52 -- clear only the entries this fragment can address
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *t4 = reinterpret_cast<uint4 *>(table);
        for (int i = lane; i < tsize / 8; i += SNP_WARP) t4[i] = z;
        __syncwarp();
    }
    const uint32_t mask = 2u * (uint32_t)(tsize - 1);
    uint32_t ip = 0;

    if (n >= 15) {  // Constants.InputMarginBytes, :190
        const uint32_t ip_limit = n - 15;
        for (;;) {
            uint32_t next_emit = ip;
            ip += 1;
            uint32_t skip = 32;  // :227
            uint32_t cand;
            bool hit = false;
            for (;;) {  // probe loop :230-341
                uint32_t x = ld_le32(in + ip);
                uint32_t stride = skip >> 5;
                skip += stride;
                uint32_t nip = ip + stride;
                if (nip > ip_limit) break;  // :323-327
                uint32_t h = table_hash<HASH_MODE>(x, mask, lut) >> 1;
                cand = table[h];
                __syncwarp();
                if (lane == 0) table[h] = (uint16_t)ip;  // :333 (write precedes compare)
                __syncwarp();
                if (ld_le32(in + cand) == x) {
                    hit = true;
                    break;
                }
                ip = nip;
            }
            if (!hit) {
                ip = next_emit;
                break;
            }
            emit_literal_v1(o, in + next_emit, ip - next_emit, lane);  // :347
            bool again;
            bool done = false;
            do {  // emit_match :358-398
                uint32_t base = ip;
                uint32_t m = 4 + find_match_length_v1(in, cand + 4, ip + 4, n, lane);
                ip += m;
                emit_copy_v1(o, base - cand, m, lane);
                if (ip >= ip_limit) {  // :381-384
                    done = true;
                    break;
                }
                uint32_t h1 = table_hash<HASH_MODE>(ld_le32(in + ip - 1), mask, lut) >> 1;
                uint32_t x = ld_le32(in + ip);
                uint32_t h = table_hash<HASH_MODE>(x, mask, lut) >> 1;
                __syncwarp();
                if (lane == 0) table[h1] = (uint16_t)(ip - 1);  // :393-394
                __syncwarp();
                cand = table[h];
                __syncwarp();
                if (lane == 0) table[h] = (uint16_t)ip;  // :397
                __syncwarp();
                again = ld_le32(in + cand) == x;  // :398
            } while (again);
            if (done) break;
        }
    }
    if (ip < n) emit_literal_v1(o, in + ip, n - ip, lane);  // :406-411
}

#ifndef SNP_EMU
// One item = varint(in_len) ++ CompressFragment(item)   (in_len <= 65536), i.e.
// Snappy.TryCompress of a single-fragment input (SnappyCompressor.cs:24-83).
// frag_mode != 0: no varint header (fragment of a larger input).
template <int HASH_MODE>
__global__ void __launch_bounds__(7 * SNP_WARP, 1)
k_compress_v1(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
              const uint32_t *__restrict__ in_len, uint8_t *out_base,
              const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
              uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items,
              int frag_mode) {
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned warps = blockDim.x / SNP_WARP;
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem + (size_t)warps * 32768);
    if (HASH_MODE == SNP_HASH_CRC32C) build_crc_lut(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const unsigned warp = threadIdx.x / SNP_WARP;
    const unsigned lane = lane_id();
    uint16_t *table = reinterpret_cast<uint16_t *>(smem + (size_t)warp * 32768);

    for (size_t item = (size_t)blockIdx.x * warps + warp; item < n_items; item += (size_t)gridDim.x * warps) {
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
            if (n > 0) compress_fragment_v1<HASH_MODE>(in, n, o, table, lut);
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
