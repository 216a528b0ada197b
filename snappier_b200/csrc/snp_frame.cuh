// snp_frame.cuh -- Snappy framing format on the GPU (BASELINE config 4).
//
//   k_crc32c_masked_batch   Crc32CAlgorithm.Compute + ApplyMask per chunk
//                           (/root/reference/Snappier/Internal/Crc32CAlgorithm.cs:41-158)
//   k_frame_plan            per chunk: compressed-vs-raw choice and framed size
//                           (SnappyStreamCompressor.CompressBlock, SnappyStreamCompressor.cs:194-230)
//   k_frame_emit            [type][len24][masked CRC32C LE][payload] at the scanned offsets
//                           (WriteCompressedBlockHeader / WriteUncompressedBlockHeader, :233-261)
//
// CRC32C is GF(2)-linear, so one warp checksums one chunk with every lane absorbing every
// 32nd word (one coalesced 128-byte load per step): lane state = F128(state ^ word), where
// F128 advances the CRC register by 128 bytes (slicing-by-4 tables for a 128-byte stride).
// The last full step uses the ordinary 4-byte advance F4, lanes are then aligned to the end
// of the region by multiplying with x^(32*(31-lane)) mod P, XOR-reduced, and the < 128-byte
// tail is absorbed serially.  Tables are built once per context into global memory.
#pragma once
#include "snp_common.cuh"

namespace snp {

#define SNP_CRC_POLY 0x82F63B78u  // reflected CRC-32C (Crc32CAlgorithm.cs:15)

// g_crc_tab[0..1023]   : F4 slicing tables   T4[k][b], k = 0..3 (byte k of the word)
// g_crc_tab[1024..2047]: F128 slicing tables T128[k][b]
__device__ uint32_t g_crc_tab[2048];

__device__ __forceinline__ uint32_t crc_advance_bits(uint32_t r, int bits) {
    for (int i = 0; i < bits; i++) r = (r & 1) ? SNP_CRC_POLY ^ (r >> 1) : (r >> 1);
    return r;
}

#ifndef SNP_EMU
__global__ void k_init_crc_tables() {
    // T4[k][b]: register after absorbing a word whose byte k is b (others 0) = byte b advanced by 8*(4-k) bits.
    // T128[k][b]: the same advanced by another 124 bytes.
    for (unsigned i = threadIdx.x + blockIdx.x * blockDim.x; i < 1024; i += blockDim.x * gridDim.x) {
        const unsigned k = i >> 8, b = i & 255;
        const uint32_t t4 = crc_advance_bits(b, 8 * (4 - k));
        g_crc_tab[i] = t4;
        g_crc_tab[1024 + i] = crc_advance_bits(t4, 8 * 124);
    }
}
#endif  // !SNP_EMU

// (a * b) mod P over GF(2), reflected bit order (bit 31 = x^0 ... as in zlib's multmodp)
__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & m) p ^= b;
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ SNP_CRC_POLY : b >> 1;
    }
    return p;
}

// x^(8*nbytes) mod P by square-and-multiply
__device__ __forceinline__ uint32_t crc_xpow8n(uint32_t nbytes) {
    uint32_t p = 1u << 31;       // x^0
    uint32_t sq = 1u << 23;      // x^8
    while (nbytes) {
        if (nbytes & 1) p = crc_mulmod(sq, p);
        sq = crc_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return p;
}

__device__ __forceinline__ uint32_t crc_f(const uint32_t *tab, uint32_t y) {
    return tab[y & 0xff] ^ tab[256 + ((y >> 8) & 0xff)] ^ tab[512 + ((y >> 16) & 0xff)] ^ tab[768 + (y >> 24)];
}

// Crc32C(data) computed by one warp (tab = the 2048-entry table set above, usually in shared memory).
__device__ __forceinline__ uint32_t crc32c_warp(const uint32_t *tab, const uint8_t *p, uint32_t n) {
    const unsigned lane = lane_id();
    // serial prefix up to 4-byte alignment (lane 0's register carries the ~0 init)
    uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)p & 3)) & 3));
    uint32_t state = 0xffffffffu;  // Crc32CAlgorithm.Append: crcLocal = uint.MaxValue ^ crc
    // byte-wise step = table of a byte advanced 8 bits = T4[3]
    for (uint32_t i = 0; i < head; i++) state = tab[768 + ((state ^ p[i]) & 0xff)] ^ (state >> 8);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p + head);
    const uint32_t words = (n - head) >> 2;
    const uint32_t steps = words >> 5;  // full 128-byte steps
    uint32_t s = lane == 0 ? state : 0u;
    if (steps) {
        for (uint32_t j = 0; j + 1 < steps; j++) s = crc_f(tab + 1024, s ^ w[j * 32 + lane]);
        s = crc_f(tab, s ^ w[(steps - 1) * 32 + lane]);
        // align lane `lane` to the end of the region: 4*(31-lane) more bytes
        s = crc_mulmod(crc_xpow8n(4 * (31 - lane)), s);
        for (int d = 16; d; d >>= 1) s ^= __shfl_xor_sync(SNP_FULL, s, d);
        state = s;
    }
    // serial tail: remaining words + bytes (warp-uniform work, < 128 + 3 bytes)
    for (uint32_t j = steps * 32; j < words; j++) state = crc_f(tab, state ^ w[j]);
    for (uint32_t i = head + words * 4; i < n; i++) state = tab[768 + ((state ^ p[i]) & 0xff)] ^ (state >> 8);
    return ~state;
}

// Crc32CAlgorithm.ApplyMask, Crc32CAlgorithm.cs:157-158
__device__ __forceinline__ uint32_t crc32c_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

#ifndef SNP_EMU
// One warp per item.  crc_out[i] = ApplyMask(Crc32C(data_i)).
__global__ void __launch_bounds__(256)
k_crc32c_masked_batch(const uint8_t *__restrict__ base, const uint64_t *__restrict__ off,
                      const uint32_t *__restrict__ len, uint32_t *__restrict__ crc_out, size_t n_items, int masked) {
    __shared__ uint32_t tab[2048];
    for (unsigned i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = g_crc_tab[i];
    __syncthreads();
    const unsigned lane = lane_id();
    const size_t warps = (size_t)gridDim.x * (blockDim.x / SNP_WARP);
    for (size_t item = (size_t)blockIdx.x * (blockDim.x / SNP_WARP) + threadIdx.x / SNP_WARP; item < n_items;
         item += warps) {
        uint32_t crc = crc32c_warp(tab, base + off[item], len[item]);
        if (masked) crc = crc32c_mask(crc);
        if (lane == 0) crc_out[item] = crc;
    }
}

// Per chunk: type 0 (compressed) iff the compressed form is smaller (SnappyStreamCompressor.cs:212-229);
// framed size = 8 + payload.  sizes[] feeds the exclusive scan.
__global__ void k_frame_plan(const uint32_t *__restrict__ raw_len, const uint32_t *__restrict__ comp_len,
                             uint32_t *__restrict__ sizes, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t payload = comp_len[i] < raw_len[i] ? comp_len[i] : raw_len[i];
    sizes[i] = 8 + payload;
}

// One CTA per chunk (grid-stride): header + payload at out + lead + scan[i].  lead = 10: this piece opens the stream and
// starts with the stream identifier; lead = 0: a later piece of the same stream.
__global__ void k_frame_emit(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ raw_off,
                             const uint32_t *__restrict__ raw_len, const uint8_t *__restrict__ comp, size_t pitch,
                             const uint32_t *__restrict__ comp_len, const uint32_t *__restrict__ crc,
                             const uint64_t *__restrict__ scan, uint8_t *__restrict__ out, size_t n, uint32_t lead) {
    if (lead && blockIdx.x == 0 && threadIdx.x < 10) {  // stream identifier (SnappyStreamCompressor.cs:15-18)
        const uint8_t id[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
        out[threadIdx.x] = id[threadIdx.x];
    }
    for (size_t i = blockIdx.x; i < n; i += gridDim.x) {
        const bool compressed = comp_len[i] < raw_len[i];
        const uint32_t payload = compressed ? comp_len[i] : raw_len[i];
        const uint8_t *src = compressed ? comp + i * pitch : raw + raw_off[i];
        uint8_t *d = out + lead + scan[i];
        if (threadIdx.x < 8) {
            const uint32_t size24 = payload + 4;  // + CRC
            const uint32_t c = crc[i];
            const uint8_t h[8] = {(uint8_t)(compressed ? 0x00 : 0x01), (uint8_t)size24, (uint8_t)(size24 >> 8),
                                  (uint8_t)(size24 >> 16), (uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16),
                                  (uint8_t)(c >> 24)};
            d[threadIdx.x] = h[threadIdx.x];
        }
        for (uint32_t k = threadIdx.x; k < payload; k += blockDim.x) d[8 + k] = src[k];
    }
}

#endif  // !SNP_EMU

}  // namespace snp
