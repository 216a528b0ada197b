// snp_decompress_v1.cuh -- baseline batched Snappy block decompressor.
//
// One warp per compressed block.  Every lane walks the tag stream redundantly
// (warp-uniform control flow); the byte movement of each literal / copy is
// spread over the 32 lanes.  This is the simple, obviously-correct kernel the
// faster kernels are A/B-checked against (SNP_DECOMP_KERNEL=v1 selects it).
//
// Semantics restated from /root/reference/Snappier/Internal/SnappyDecompressor.cs
// :43-92 (length prefix), :184-347 (DecompressAllTags), :556-611 (Append /
// AppendFromSelf), one-shot whole-block only; Constants.cs:42-76 (CharTable) is
// replaced by arithmetic on the tag byte.
#pragma once
#include "snp_common.cuh"

namespace snp {

// Returns the block status; *written = bytes produced on SNP_OK, else 0.
__device__ __noinline__ int decompress_block_v1(const uint8_t *__restrict__ in, uint32_t n_in,
                                                uint8_t *out, uint32_t cap, uint32_t *written) {
    const unsigned lane = lane_id();
    *written = 0;
    uint32_t U, used;
    int st = varint_read(in, n_in, &U, &used);
    if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;  // SnappyDecompressor.cs:57-60 + Snappy.cs:178-181
    if (st != SNP_OK || U > 0x7fffffffu) return SNP_INVALID_LENGTH;
    if (cap < U) return SNP_OUTPUT_TOO_SMALL;
    if (U == 0) return SNP_OK;  // AllDataDecompressed before any tag (SnappyDecompressor.cs:78)

    uint32_t ip = used, op = 0;
    while (ip < n_in) {
        uint32_t c = in[ip];
        uint32_t kind = c & 3;
        uint32_t extra = kind == 0 ? ((c >> 2) >= 60 ? (c >> 2) - 59 : 0) : (kind == 1 ? 1 : kind == 2 ? 2 : 4);
        if (n_in - ip < 1 + extra) break;  // truncated tag (RefillTag, :464-483)
        uint32_t trailer = 0;
        for (uint32_t i = 0; i < extra; i++) trailer |= (uint32_t)in[ip + 1 + i] << (8 * i);
        ip += 1 + extra;
        if (kind == 0) {
            uint64_t len = (uint64_t)((c >> 2) >= 60 ? trailer : (c >> 2)) + 1;  // :264-288
            uint32_t avail = n_in - ip;
            uint32_t take = len < avail ? (uint32_t)len : avail;  // partial literal, :290-297
            if (take > U - op) return SNP_DATA_TOO_LONG;           // :570-573
            for (uint32_t k = lane; k < take; k += SNP_WARP) out[op + k] = in[ip + k];
            op += take;
            ip += take;
            if (take < len) break;
        } else {
            uint32_t len, offset;
            if (kind == 1) {
                len = ((c >> 2) & 7) + 4;
                offset = ((c >> 5) << 8) | trailer;
            } else {
                len = (c >> 2) + 1;
                offset = trailer;
            }
            if (offset == 0 || op < offset) return SNP_INVALID_COPY_OFFSET;  // :598-601
            if (len > U - op) return SNP_DATA_TOO_LONG;                       // :603-606
            const uint8_t *src = out + (op - offset);
            if (offset >= SNP_WARP || offset >= len) {
                // chunks of 32 in increasing order; chunk j only reads bytes < op + 32j
                for (uint32_t k0 = 0; k0 < len; k0 += SNP_WARP) {
                    uint32_t k = k0 + lane;
                    if (k < len) out[op + k] = src[k];
                    __syncwarp();
                }
            } else {
                // pattern replication (CopyHelpers.cs:76-160): every source byte is < op
                for (uint32_t k = lane; k < len; k += SNP_WARP) out[op + k] = src[k % offset];
            }
            op += len;
        }
        __syncwarp();  // order this tag's global stores before later back-reference loads
    }
    if (op < U) return SNP_INCOMPLETE;  // Snappy.cs:178-181
    *written = op;
    return SNP_OK;
}

#ifndef SNP_EMU  // kernels: device only (the block functions above also run on tests/cpp/simt_emu.h)
__global__ void __launch_bounds__(256)
k_decompress_v1(const uint8_t *__restrict__ in_base, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_len, uint8_t *out_base,
                const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
                uint32_t *__restrict__ out_len, int32_t *__restrict__ status, size_t n_items) {
    size_t item = (size_t)blockIdx.x * (blockDim.x / SNP_WARP) + threadIdx.x / SNP_WARP;
    if (item >= n_items) return;
    uint32_t w = 0;
    int st = decompress_block_v1(in_base + in_off[item], in_len[item], out_base + out_off[item],
                                 out_cap[item], &w);
    if (lane_id() == 0) {
        out_len[item] = w;
        status[item] = st;
    }
}

// Batched Snappy.GetUncompressedLength (Snappy.cs:142-143): one thread per item.
__global__ void k_uncompressed_length(const uint8_t *__restrict__ in_base,
                                      const uint64_t *__restrict__ in_off,
                                      const uint32_t *__restrict__ in_len, uint32_t *__restrict__ ulen,
                                      int32_t *__restrict__ status, size_t n_items) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint32_t v, used;
    int st = varint_read(in_base + in_off[i], in_len[i], &v, &used);
    if (st != SNP_OK || v > 0x7fffffffu) {  // VarIntEncoding.Read.cs:16-24: anything but Done
        st = SNP_INVALID_LENGTH;
        v = 0;
    }
    ulen[i] = v;
    status[i] = st;
}

#endif  // !SNP_EMU

}  // namespace snp
