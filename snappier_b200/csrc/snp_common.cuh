// snp_common.cuh -- device-side helpers shared by the Snappy block kernels (sm_100a).
#pragma once

#ifndef SNP_EMU  // tests/cpp/simt_emu.h supplies the intrinsics for the host-side warp emulator
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/snappier_b200.h"

#define SNP_WARP 32
#define SNP_FULL 0xffffffffu

namespace snp {

#ifdef SNP_EMU
__device__ __forceinline__ unsigned lane_id() { return (unsigned)simt::lane(); }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << simt::lane()) - 1u; }
#else
__device__ __forceinline__ unsigned lane_id() {
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));  // not volatile: constant per thread, the compiler may hoist / reuse it
    return l;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif

// ---- tag table (replaces Constants.CharTable, Constants.cs:42-76) -----------
// byte 0: length in the tag byte (literal n<60: n+1; copies: len)
// byte 1: header bytes (tag byte + trailer): 1..5
// byte 2: shift turning 0xffffffff into the trailer mask (32 - 8*trailer_bytes)
// bits 24..26: COPY1 offset bits 8..10;  bit 30: literal length is in the trailer;  bit 31: literal
__device__ __forceinline__ uint32_t tag_lut3_entry(uint32_t c) {
    uint32_t kind = c & 3, n6 = c >> 2;
    uint32_t len = 0, hdr, flags = 0;
    if (kind == 0) {
        flags = 0x80000000u;
        if (n6 < 60) {
            len = n6 + 1;
            hdr = 1;
        } else {
            flags |= 0x40000000u;
            hdr = 1 + (n6 - 59);
        }
    } else if (kind == 1) {
        len = (n6 & 7) + 4;
        hdr = 2;
        flags = (c >> 5) << 24;
    } else {
        len = n6 + 1;
        hdr = kind == 2 ? 3 : 5;
    }
    return len | (hdr << 8) | ((32 - 8 * (hdr - 1)) << 16) | flags;
}

// Unaligned little-endian 32-bit load from global or generic memory.  Touches only
// aligned words that contain at least one requested byte.
__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = w[0];
    uint32_t hi = sh ? w[1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// ---- varint (VarIntEncoding.Read.cs:38-79, slow path = the semantics) -------
// Uniform across the warp (every lane computes the same thing).
__device__ __forceinline__ int varint_read(const uint8_t *in, uint32_t n, uint32_t *v, uint32_t *used) {
    uint32_t result = 0;
    int shift = 0;
    *v = 0;
    *used = 0;
    for (uint32_t i = 0; i < n && i < 5; i++) {
        uint32_t c = in[i];
        uint32_t val = c & 0x7f;
        if (val & ~(0xffffffffu >> shift)) return SNP_INVALID_LENGTH;  // Helpers.cs:66-70
        result |= val << shift;
        shift += 7;
        if (c < 128) {
            *v = result;
            *used = i + 1;
            return SNP_OK;
        }
        if (shift >= 32) return SNP_INVALID_LENGTH;
    }
    return SNP_INCOMPLETE;  // OperationStatus.NeedMoreData
}

// VarIntEncoding.Write.cs:5-79.  Returns the encoded length (1..5) and the bytes
// packed little-endian in *lo (bytes 0..3) and *hi (byte 4).
__device__ __forceinline__ int varint_encode(uint32_t v, uint32_t *lo, uint32_t *hi) {
    int need = v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5;
    uint64_t acc = 0;
    for (int i = 0; i < need; i++) {
        uint32_t b = (v >> (7 * i)) & 0x7f;
        if (i != need - 1) b |= 0x80;
        acc |= (uint64_t)b << (8 * i);
    }
    *lo = (uint32_t)acc;
    *hi = (uint32_t)(acc >> 32);
    return need;
}

// HashTable.cs:57-71
__device__ __forceinline__ int table_size_for(uint32_t n) {
    if (n > 16384) return 16384;
    if (n < 256) return 256;
    return 2 << (31 - __clz(n - 1));
}

// ---- CRC32C-hash lookup tables (HashTable.cs:109-117) ------------------------
// Sse42.Crc32(bytes, mask) == F(bytes ^ mask), F = four byte rounds of the
// reflected CRC-32C polynomial 0x82F63B78 with no init / final xor.  By slicing,
// F(y) = T3[y&ff] ^ T2[(y>>8)&ff] ^ T1[(y>>16)&ff] ^ T0[y>>24].  Only bits 1..14
// of the hash survive `& mask`, so the tables keep the low 16 bits.
// Layout: lut[k*256 + i] = T_k[i] & 0xffff.  (2 KiB of shared memory.)
__device__ __forceinline__ void build_crc_lut(uint16_t *lut, unsigned tid, unsigned nthreads) {
    for (unsigned i = tid; i < 256; i += nthreads) {
        uint32_t r = i;
#pragma unroll
        for (int k = 0; k < 8; k++) r = (r & 1) ? 0x82F63B78u ^ (r >> 1) : (r >> 1);
        uint32_t t0 = r;
        lut[i] = (uint16_t)t0;
        // one more zero byte per level: T_{k+1}[i] = T_0[T_k[i] & 0xff] ^ (T_k[i] >> 8);
        // T_0 of an arbitrary byte is recomputed bitwise to avoid a sync between levels.
        uint32_t t = t0;
#pragma unroll
        for (int lvl = 1; lvl < 4; lvl++) {
            uint32_t b = t & 0xff;
#pragma unroll
            for (int k = 0; k < 8; k++) b = (b & 1) ? 0x82F63B78u ^ (b >> 1) : (b >> 1);
            t = b ^ (t >> 8);
            lut[lvl * 256 + i] = (uint16_t)t;
        }
    }
}

// Byte offset into the u16 table: hash & mask  (mask = 2*(table_size-1)).
template <int HASH_MODE>
__device__ __forceinline__ uint32_t table_hash(uint32_t bytes, uint32_t mask, const uint16_t *lut) {
    if (HASH_MODE == SNP_HASH_CRC32C) {
        uint32_t y = bytes ^ mask;
        uint32_t h = (uint32_t)lut[768 + (y & 0xff)] ^ (uint32_t)lut[512 + ((y >> 8) & 0xff)] ^
                     (uint32_t)lut[256 + ((y >> 16) & 0xff)] ^ (uint32_t)lut[y >> 24];
        return h & mask;
    } else {
        return ((0x1e35a7bdu * bytes) >> 17) & mask;  // HashTable.cs:120-123
    }
}

}  // namespace snp
