// snp_engine.cu -- host side of the C ABI in include/snappier_b200.h.
//
// Contexts, staging, launches.  All arithmetic of the Snappy block path runs in
// the CUDA kernels included below; there is no CPU implementation here (the
// only host-side arithmetic is the 1..5-byte varint length prefix, which the
// reference also reads before it decides how much memory to rent,
// SnappyDecompressor.cs:110-173).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <deque>
#include <vector>

#include "snp_common.cuh"
#include "snp_compress_v1.cuh"
#include "snp_compress_v3.cuh"
#include "snp_decompress_v1.cuh"
#include "snp_decompress_v7.cuh"
#include "snp_decompress_v8.cuh"
#include "snp_frame.cuh"

namespace {

thread_local std::string g_last_error;

int cuda_fail(cudaError_t e, const char *what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at snp_engine.cu:%d: %s (%s)", what, line, cudaGetErrorName(e),
             cudaGetErrorString(e));
    g_last_error = buf;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice ||
        e == cudaErrorInvalidDevice)
        return SNP_E_NO_DEVICE;
    return SNP_E_CUDA;
}

#define CU(call)                                               \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call, __LINE__); \
    } while (0)

struct PinnedBuf {  // grow-only pinned host staging (metadata only; payloads are the caller's)
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n);
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
};

struct DevBuf {  // grow-only device scratch
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return SNP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = n;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)", __LINE__);
        cap = want;
        return SNP_OK;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

int PinnedBuf::reserve(size_t n) {
    if (n <= cap) return SNP_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, n + n / 2 + 4096);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost(meta)", __LINE__);
    cap = n + n / 2 + 4096;
    return SNP_OK;
}

}  // namespace

struct snp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::atomic<uint64_t> launches{0};
    int sm_count = 148;
    int decomp_kernel = 7;  // SNP_DECOMP_KERNEL: 7 = warp-per-block tag-group engine (TMA-staged input ring, advance-table
                            // walk, output window; the default), 8 = lane-per-block engine (the challenger: faster on
                            // batches of <= 4 KiB blocks), 1 = warp-uniform baseline (also the > 2 GiB path)
    int v7_window = 4096;   // SNP_V7_WINDOW: output window bytes per warp of k_decompress_v7: 4096 (32 warps per SM, the
                            // default) or 2048 (40 warps per SM)
    int v8_cfg = 0;         // SNP_V8_CFG: lane-per-block engine instantiation (see launch_decompress)
    int comp_kernel = 3;    // SNP_COMP_KERNEL: 3 = L2 tables with fingerprinted 32-bit entries (default), 6 = L2 tables with
                            // the reference's plain 16-bit entries (the challenger), 1 = baseline (tables in shared memory)
    DevBuf d_in, d_out, d_meta, d_tmp;
    DevBuf d_tables;           // k_compress_v3: one 64 KiB hash table slice per resident warp
    cudaEvent_t tables_done = nullptr;  // orders compress launches that arrive on different streams (shared d_tables)
    cudaStream_t tables_last_stream = nullptr;
    bool tables_used = false;
    int comp_ctas_per_sm = 8;  // SNP_COMP_CTAS_PER_SM (x 8 warps)
    int host_trace = 0;  // SNP_HOST_TRACE: print a per-chunk timeline of the host-mode pipeline (diagnostics)
    struct TraceRow { cudaEvent_t e[5]; };
    std::vector<TraceRow> trace;
    int host_early_d2h = 1;  // SNP_HOST_EARLY_D2H: enqueue the payload copy of dense decompress chunks behind the kernel
    uint64_t host_chunk_bytes = 64ull << 20;   // host-mode pipeline: bytes per chunk (SNP_HOST_CHUNK_MB)
    uint64_t host_comp_chunk_bytes = 256ull << 20;  // the same for compress calls (SNP_HOST_COMP_CHUNK_MB): raw bytes per chunk
    int comp_first_width = 16;  // SNP_COMP_FIRST_WIDTH: probes in the first batch after a match (k_compress_v3; 32 = fixed width)
    unsigned long long *d_counters = nullptr;  // pool of work counters for the persistent kernels
    unsigned counter_seq = 0;
    static constexpr int kSlots = 8;  // host-mode pipeline depth (H2D | kernel | D2H overlap)
    struct Slot {
        cudaStream_t stream = nullptr;
        cudaEvent_t meta_ready = nullptr;
        DevBuf d_in, d_out, d_meta, d_tmp;
        DevBuf d_tables;   // hash tables of this slot's compress launches (sized by the chunk's grid): the chunks of the
                           // host-mode pipeline compress concurrently, each on its slot's stream
        PinnedBuf h_meta;  // same layout as d_meta: async both ways regardless of the caller's arrays
    } slots[kSlots];
    bool attrs_set = false;
};

namespace {

int ctx_work_counter(snp_ctx *c, cudaStream_t s, unsigned long long **out);

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kCompWarps = 7;  // 7 x 32 KiB tables + 2 KiB LUT = 226 KiB <= 227 KiB
constexpr size_t kCompSmem = (size_t)kCompWarps * 32768 + 2048;

int ctx_set_attrs(snp_ctx *c) {
    if (c->attrs_set) return SNP_OK;
    CU(cudaFuncSetAttribute(snp::k_compress_v1<SNP_HASH_CRC32C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kCompSmem));
    CU(cudaFuncSetAttribute(snp::k_compress_v1<SNP_HASH_MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kCompSmem));
#define SNP7_ATTR(W, NW, CTAS)                                                                              \
    CU(cudaFuncSetAttribute(snp::k_decompress_v7<W, NW, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                            (int)(NW * sizeof(snp::Warp7<W>))))
    SNP7_ATTR(2048, 8, 5);
    SNP7_ATTR(4096, 8, 4);
#undef SNP7_ATTR
#define SNP8_ATTR(IR, ORB, D, NT, CTAS)                                                                              \
    CU(cudaFuncSetAttribute(snp::k_decompress_v8<IR, ORB, D, NT, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                            (int)(NT * sizeof(snp::Lane8<IR, ORB, D>))))
    SNP8_ATTR(128, 128, 3, 96, 6);
    SNP8_ATTR(128, 128, 3, 128, 4);
    SNP8_ATTR(128, 128, 4, 128, 4);
#undef SNP8_ATTR

    c->attrs_set = true;
    return SNP_OK;
}

// Hands out a zeroed work counter for one persistent launch on stream s (the memset is
// ordered before the launch).  Counters rotate through a pool so that launches in flight
// on different streams never share one.
constexpr unsigned kCounterPool = 4096;  // far more than the launches that can be in flight at once
int ctx_work_counter(snp_ctx *c, cudaStream_t s, unsigned long long **out) {
    if (!c->d_counters) CU(cudaMalloc((void **)&c->d_counters, kCounterPool * 32));
    unsigned long long *p = c->d_counters + 4 * (c->counter_seq++ % kCounterPool);
    CU(cudaMemsetAsync(p, 0, 8, s));
    *out = p;
    return SNP_OK;
}

int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    if (!s || !*s) return dflt;
    if (s[0] == 'v' || s[0] == 'V') s++;
    return atoi(s);
}

// ---------------------------------------------------------------- launches --

int launch_decompress(snp_ctx *c, cudaStream_t s, const uint8_t *in_base, const uint64_t *in_off,
                      const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off,
                      const uint32_t *out_cap, uint32_t *out_len, int32_t *status, size_t n) {
    if (n == 0) return SNP_OK;
    const int kernel = c->decomp_kernel;
    const int warps = 8;
    unsigned grid = (unsigned)((n + warps - 1) / warps);
    if (kernel != 8 && kernel != 1) {  // 7, the default
        int rc = ctx_set_attrs(c);
        if (rc) return rc;
        unsigned long long *ctr;
        if ((rc = ctx_work_counter(c, s, &ctr))) return rc;
#define SNP7_LAUNCH(W, NW, CTAS)                                                                                  \
    do {                                                                                                          \
        const unsigned g7 = std::min((unsigned)((n + NW - 1) / NW), (unsigned)(c->sm_count * CTAS));              \
        snp::k_decompress_v7<W, NW, CTAS><<<g7, NW * SNP_WARP, NW * sizeof(snp::Warp7<W>), s>>>(                  \
            in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, ctr);                       \
    } while (0)
        if (c->v7_window == 2048) SNP7_LAUNCH(2048, 8, 5);        // 40 warps per SM, 2 KiB windows (measured 3-4 % slower)
        else SNP7_LAUNCH(4096, 8, 4);                             // 32 warps per SM, 4 KiB windows (default)
#undef SNP7_LAUNCH
    } else if (kernel == 8) {
        int rc = ctx_set_attrs(c);
        if (rc) return rc;
        unsigned long long *ctr;
        if ((rc = ctx_work_counter(c, s, &ctr))) return rc;
#define SNP8_LAUNCH(IR, ORB, D, NT, CTAS)                                                                         \
    do {                                                                                                          \
        const unsigned g8 = std::min((unsigned)((n + NT - 1) / NT), (unsigned)(c->sm_count * CTAS));              \
        snp::k_decompress_v8<IR, ORB, D, NT, CTAS><<<g8, NT, NT * sizeof(snp::Lane8<IR, ORB, D>), s>>>(          \
            in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, ctr);                       \
    } while (0)
        if (c->v8_cfg == 1) SNP8_LAUNCH(128, 128, 3, 128, 4);       // 512 lanes per SM
        else if (c->v8_cfg == 2) SNP8_LAUNCH(128, 128, 4, 128, 4);  // deeper pipeline
        else SNP8_LAUNCH(128, 128, 3, 96, 6);                       // 576 lanes per SM, 128-byte rings, depth 3
#undef SNP8_LAUNCH
    } else if (kernel == 1)
        snp::k_decompress_v1<<<grid, warps * SNP_WARP, 0, s>>>(in_base, in_off, in_len, out_base, out_off,
                                                               out_cap, out_len, status, n);
    c->launches++;
    CU(cudaGetLastError());
    return SNP_OK;
}

int launch_compress(snp_ctx *c, cudaStream_t s, const uint8_t *in_base, const uint64_t *in_off,
                    const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                    uint32_t *out_len, int32_t *status, size_t n, uint32_t hash_mode, int frag_mode,
                    DevBuf *own_tables = nullptr) {
    if (n == 0) return SNP_OK;
    int rc = ctx_set_attrs(c);
    if (rc) return rc;
    // own_tables: a table buffer that only launches on stream s use (a pipeline slot): no cross-stream ordering needed
    DevBuf &tables = own_tables ? *own_tables : c->d_tables;
    // The L2-table kernels share ONE table buffer (a slice per resident CTA/warp), and the host-mode pipeline launches
    // consecutive chunks on different streams: a chunk's CTAs could start in the tail of the previous launch and clear a
    // slice that one of its warps is still using.  Launches on different streams are therefore chained with an event
    // (each launch fills the GPU by itself, so nothing is lost; the copies of the chunks still overlap).
    const bool shared_tables = c->comp_kernel != 1 && !own_tables;
    if (shared_tables) {
        if (!c->tables_done) CU(cudaEventCreateWithFlags(&c->tables_done, cudaEventDisableTiming));
        if (c->tables_used && c->tables_last_stream != s) CU(cudaStreamWaitEvent(s, c->tables_done, 0));
    }
    size_t ctas = (n + kCompWarps - 1) / kCompWarps;
    unsigned grid = (unsigned)(ctas < (size_t)c->sm_count ? ctas : (size_t)c->sm_count);
    if (c->comp_kernel == 1) {
        if (hash_mode == SNP_HASH_CRC32C)
            snp::k_compress_v1<SNP_HASH_CRC32C><<<grid, kCompWarps * SNP_WARP, kCompSmem, s>>>(
                in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, frag_mode);
        else
            snp::k_compress_v1<SNP_HASH_MUL><<<grid, kCompWarps * SNP_WARP, kCompSmem, s>>>(
                in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, frag_mode);
    } else {
        // hash tables in global memory (L2): occupancy not capped by shared memory
        unsigned long long *ctr;
        if ((rc = ctx_work_counter(c, s, &ctr))) return rc;
        const int wpc = 8;
        const int cps = c->comp_ctas_per_sm;
        size_t ctas3 = (n + wpc - 1) / wpc;
        unsigned grid3 = (unsigned)std::min(ctas3, (size_t)c->sm_count * cps);
        const size_t table_bytes = (size_t)grid3 * wpc * 65536;  // one 64 KiB slice per warp of THIS launch
        if ((rc = tables.reserve(table_bytes))) return rc;
#define SNP_LAUNCH_C3(H, V)                                                                                  \
    snp::k_compress_v3<H, V><<<grid3, wpc * SNP_WARP, 0, s>>>(in_base, in_off, in_len, out_base, out_off, out_cap, \
                                                              out_len, status, n,                                  \
                                                              frag_mode | (c->comp_first_width << 8), ctr,         \
                                                              (uint32_t *)tables.p)
        if (c->comp_kernel == 6) {  // plain 16-bit table entries
            if (hash_mode == SNP_HASH_CRC32C) SNP_LAUNCH_C3(SNP_HASH_CRC32C, 6);
            else SNP_LAUNCH_C3(SNP_HASH_MUL, 6);
        } else {
            if (hash_mode == SNP_HASH_CRC32C) SNP_LAUNCH_C3(SNP_HASH_CRC32C, 3);
            else SNP_LAUNCH_C3(SNP_HASH_MUL, 3);
        }
#undef SNP_LAUNCH_C3
    }
    c->launches++;
    CU(cudaGetLastError());
    if (shared_tables) {
        CU(cudaEventRecord(c->tables_done, s));
        c->tables_last_stream = s;
        c->tables_used = true;
    }
    return SNP_OK;
}

// ------------------------------------------- concat of fragment outputs ----
// The `output = output.Slice(written)` running pointer of TryCompress
// (SnappyCompressor.cs:76-79) as an exclusive scan + gather.

__global__ void k_frag_scan(const uint32_t *__restrict__ len, const int32_t *__restrict__ status,
                            uint64_t *__restrict__ off, uint64_t *__restrict__ total_and_bad, size_t n) {
    __shared__ uint64_t part[1024];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    size_t per = (n + blockDim.x - 1) / blockDim.x;
    size_t lo = (size_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint64_t sum = 0;
    for (size_t i = lo; i < hi; i++) {
        sum += len[i];
        if (status[i] != SNP_OK) bad = 1;
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (unsigned t = 0; t < blockDim.x; t++) {
            uint64_t v = part[t];
            part[t] = run;
            run += v;
        }
        total_and_bad[0] = run;
        total_and_bad[1] = (uint64_t)bad;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (size_t i = lo; i < hi; i++) {
        off[i] = run;
        run += len[i];
    }
}

__global__ void k_frag_gather(const uint8_t *__restrict__ tmp, size_t pitch, const uint32_t *__restrict__ len,
                              const uint64_t *__restrict__ off, uint8_t *__restrict__ out, size_t n) {
    for (size_t f = blockIdx.x; f < n; f += gridDim.x) {
        const uint8_t *s = tmp + f * pitch;
        uint8_t *d = out + off[f];
        uint32_t l = len[f];
        for (uint32_t k = threadIdx.x; k < l; k += blockDim.x) d[k] = s[k];
    }
}

// Batched SnappyCompressor.FindMatchLength (SnappyCompressor.cs:562-688): one warp per query, the compress kernels'
// own device function -- so that the reference's known-answer vectors run on the GPU directly.
__global__ void __launch_bounds__(256) k_find_match_length(const uint8_t *__restrict__ base, const uint32_t *__restrict__ s1,
                                                           const uint32_t *__restrict__ s2, const uint32_t *__restrict__ limit,
                                                           uint32_t *__restrict__ out, size_t n) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t i = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const uint32_t m = snp::find_match_length_v2(base, s1[i], s2[i], limit[i], lane);
        if (lane == 0) out[i] = m;
    }
}

// ---- diagnostics: the chip's random-access read ceiling ---------------------------------------------
// Every thread issues `reads` dependent-free 16-byte loads at pseudo-random 16-byte-aligned offsets of [0, span) (an LCG per
// thread; `ilp` loads in flight per thread), and folds them into a checksum so that nothing is optimised away.
__global__ void __launch_bounds__(256) k_diag_random_reads(const uint4 *__restrict__ base, unsigned long long span16,
                                                           uint32_t reads, uint32_t *__restrict__ sink) {
    unsigned long long x = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    uint32_t acc = 0;
    for (uint32_t i = 0; i < reads; i += 4) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            v[k] = __ldcg(base + (x >> 17) % span16);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345679u) sink[0] = acc;  // practically never: keeps the loads alive
}

// ---- packing a batch: slots with slack -> dense bytes (the gather(v) side of block-range sharding, SURVEY.md 8(e);
// the batched form of what Snappy.CompressToMemory returns: exactly the compressed bytes, no slack) --------------------
// Exclusive scan of len[0..n) in three launches: per-CTA sums, scan of the sums (one CTA), per-CTA offsets.
constexpr int kScanItems = 2048;  // items per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) k_pack_sums(const uint32_t *__restrict__ len, uint64_t *__restrict__ part, size_t n) {
    __shared__ uint64_t ws[8];
    const size_t base = (size_t)blockIdx.x * kScanItems;
    uint64_t sum = 0;
    for (int k = 0; k < 8; k++) {
        const size_t i = base + (size_t)k * 256 + threadIdx.x;
        if (i < n) sum += len[i];
    }
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(SNP_FULL, sum, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_pack_scan_sums(uint64_t *__restrict__ part, size_t nparts, uint64_t *__restrict__ total) {
    __shared__ uint64_t ws[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nparts; b0 += 1024) {
        const size_t i = b0 + threadIdx.x;
        const uint64_t v = i < nparts ? part[i] : 0;
        uint64_t inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(SNP_FULL, inc, d);
            if ((threadIdx.x & 31) >= (unsigned)d) inc += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t w = ws[threadIdx.x], winc = w;
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t y = __shfl_up_sync(SNP_FULL, winc, d);
                if (threadIdx.x >= (unsigned)d) winc += y;
            }
            ws[threadIdx.x] = winc - w;  // exclusive over the warps
        }
        __syncthreads();
        const uint64_t excl = carry + ws[threadIdx.x >> 5] + inc - v;
        if (i < nparts) part[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) k_pack_offsets(const uint32_t *__restrict__ len, const uint64_t *__restrict__ part,
                                                      uint64_t *__restrict__ off, size_t n) {
    __shared__ uint64_t ws[8];
    const size_t i0 = (size_t)blockIdx.x * kScanItems + (size_t)threadIdx.x * 8;
    uint32_t l[8];
    uint64_t sum = 0;
    for (int k = 0; k < 8; k++) {
        l[k] = i0 + k < n ? len[i0 + k] : 0u;
        sum += l[k];
    }
    uint64_t inc = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(SNP_FULL, inc, d);
        if ((threadIdx.x & 31) >= (unsigned)d) inc += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint64_t run = part[blockIdx.x] + inc - sum;
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) run += ws[w];
    for (int k = 0; k < 8; k++) {
        if (i0 + k < n) off[i0 + k] = run;
        run += l[k];
    }
}

// One warp per item, 16-byte vectors aligned on the destination (the sources are slots, usually 16-byte aligned).
__global__ void __launch_bounds__(256) k_pack_copy(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off,
                                                   const uint32_t *__restrict__ len, uint8_t *__restrict__ dst_base,
                                                   const uint64_t *__restrict__ dst_off, size_t n) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t i = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const uint32_t l = len[i];
        if (l == 0) continue;
        const uint8_t *sp = src_base + src_off[i];
        snp::copy_literal_wide7(sp, dst_base + dst_off[i], l, sp + l, lane);
    }
}

// --------------------------------------------------------- host-mode staging --

struct Span {
    uint64_t lo = 0, hi = 0;
};


struct MetaLayout {  // one device allocation holding every per-item array
    size_t in_off, in_len, out_off, out_cap, out_len, status, bytes;
    explicit MetaLayout(size_t n) {
        size_t p = 0;
        in_off = p, p += align_up(n * 8, 256);
        out_off = p, p += align_up(n * 8, 256);
        in_len = p, p += align_up(n * 4, 256);
        out_cap = p, p += align_up(n * 4, 256);
        out_len = p, p += align_up(n * 4, 256);
        status = p, p += align_up(n * 4, 256);
        bytes = p;
    }
};

// Runs one batched op on host buffers.  Synchronous for the caller, but internally the batch is
// cut into chunks that flow through kSlots streams, so that the H2D copy of chunk k+1, the
// kernel of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full duplex).  With pinned
// host buffers the copies are true DMA; with pageable buffers the driver stages them and the
// overlap degrades gracefully.
struct Chunk {
    size_t a = 0, b = 0;  // item range
    Span si, so;
    int slot = 0;
    bool early_d2h = false;  // the payload copy was already enqueued in phase 1
    int trace_idx = -1;
};

int chunk_phase1(snp_ctx *c, Chunk &ck, bool compress, const uint8_t *in_base, const uint64_t *in_off,
                 const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                 uint32_t hash_mode) {
    snp_ctx::Slot &sl = c->slots[ck.slot];
    cudaStream_t s = sl.stream;
    const size_t n = ck.b - ck.a;
    MetaLayout ml(n);
    int rc;
    CU(cudaStreamSynchronize(s));  // the slot's previous chunk is completely done
    if ((rc = sl.d_in.reserve(ck.si.hi - ck.si.lo + 16))) return rc;
    if ((rc = sl.d_out.reserve(ck.so.hi - ck.so.lo + 16))) return rc;
    if ((rc = sl.d_meta.reserve(ml.bytes))) return rc;
    if ((rc = sl.h_meta.reserve(ml.bytes))) return rc;
    uint8_t *dm = (uint8_t *)sl.d_meta.p;
    uint8_t *hm = (uint8_t *)sl.h_meta.p;
    {
        uint64_t *ri = (uint64_t *)(hm + ml.in_off), *ro = (uint64_t *)(hm + ml.out_off);
        for (size_t i = 0; i < n; i++) ri[i] = in_off[ck.a + i] - ck.si.lo;
        for (size_t i = 0; i < n; i++) ro[i] = out_off[ck.a + i] - ck.so.lo;
        memcpy(hm + ml.in_len, in_len + ck.a, n * 4);
        memcpy(hm + ml.out_cap, out_cap + ck.a, n * 4);
    }
    // in_off .. out_cap are contiguous in the layout: one H2D for all input metadata
    CU(cudaMemcpyAsync(dm, hm, ml.out_len, cudaMemcpyHostToDevice, s));
    cudaEvent_t tev[5] = {};
    if (c->host_trace) {
        for (auto &e : tev) CU(cudaEventCreate(&e));
        CU(cudaEventRecord(tev[0], s));
    }
    if (ck.si.hi > ck.si.lo)
        CU(cudaMemcpyAsync(sl.d_in.p, in_base + ck.si.lo, ck.si.hi - ck.si.lo, cudaMemcpyHostToDevice, s));
    if (c->host_trace) CU(cudaEventRecord(tev[1], s));
    auto *d_in_off = (const uint64_t *)(dm + ml.in_off);
    auto *d_out_off = (const uint64_t *)(dm + ml.out_off);
    auto *d_in_len = (const uint32_t *)(dm + ml.in_len);
    auto *d_out_cap = (const uint32_t *)(dm + ml.out_cap);
    auto *d_out_len = (uint32_t *)(dm + ml.out_len);
    auto *d_status = (int32_t *)(dm + ml.status);
    if (compress)
        rc = launch_compress(c, s, (const uint8_t *)sl.d_in.p, d_in_off, d_in_len, (uint8_t *)sl.d_out.p, d_out_off,
                             d_out_cap, d_out_len, d_status, n, hash_mode, 0, &sl.d_tables);
    else
        rc = launch_decompress(c, s, (const uint8_t *)sl.d_in.p, d_in_off, d_in_len, (uint8_t *)sl.d_out.p,
                               d_out_off, d_out_cap, d_out_len, d_status, n);
    if (rc) return rc;
    if (c->host_trace) CU(cudaEventRecord(tev[2], s));
    // out_len + status are contiguous: one D2H into the pinned mirror (copied to the caller in phase 2)
    CU(cudaMemcpyAsync(hm + ml.out_len, dm + ml.out_len, ml.bytes - ml.out_len, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(sl.meta_ready, s));
    // Decompress into back-to-back capacity regions (the usual dense layout): the whole output span lies inside the
    // caller's regions, so its copy can be enqueued right behind the kernel instead of after a host round trip for
    // the produced lengths (bytes of a region beyond out_len are unspecified, as in the device-mode call).
    ck.early_d2h = false;
    // A chunk of ONE item (the single-call API) waits for its status instead: a rejected block must leave the caller's
    // buffer untouched, as the reference does (it decodes into a private buffer), and only out_len bytes are copied.
    if (!compress && c->host_early_d2h && n > 1) {
        bool dense = true;
        for (size_t i = ck.a; i + 1 < ck.b && dense; i++) dense = out_off[i + 1] == out_off[i] + out_cap[i];
        if (dense && ck.so.hi > ck.so.lo) {
            if (c->host_trace) CU(cudaEventRecord(tev[3], s));
            CU(cudaMemcpyAsync(out_base + ck.so.lo, sl.d_out.p, ck.so.hi - ck.so.lo, cudaMemcpyDeviceToHost, s));
            if (c->host_trace) CU(cudaEventRecord(tev[4], s));
            ck.early_d2h = true;
        }
    }
    if (c->host_trace) {
        ck.trace_idx = (int)c->trace.size();
        c->trace.push_back({tev[0], tev[1], tev[2], tev[3], tev[4]});
    }
    return SNP_OK;
}

// Copy a chunk's results back in runs of adjacent item regions, trimmed to the bytes
// produced, so that nothing outside the callers' capacity regions is ever written.
int chunk_phase2(snp_ctx *c, const Chunk &ck, uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                 uint32_t *out_len, int32_t *status) {
    snp_ctx::Slot &sl = c->slots[ck.slot];
    CU(cudaEventSynchronize(sl.meta_ready));  // out_len / status of this chunk are on the host now
    {
        MetaLayout ml(ck.b - ck.a);
        const uint8_t *hm = (const uint8_t *)sl.h_meta.p;
        memcpy(out_len + ck.a, hm + ml.out_len, (ck.b - ck.a) * 4);
        memcpy(status + ck.a, hm + ml.status, (ck.b - ck.a) * 4);
    }
    if (ck.early_d2h) return SNP_OK;
    const bool tr = c->host_trace && ck.trace_idx >= 0;
    if (tr) CU(cudaEventRecord(c->trace[ck.trace_idx].e[3], sl.stream));
    struct TraceEnd {  // records the end of the payload copies on every exit path
        snp_ctx *c; const Chunk &ck; cudaStream_t s; bool on;
        ~TraceEnd() { if (on) cudaEventRecord(c->trace[ck.trace_idx].e[4], s); }
    } trace_end{c, ck, sl.stream, tr};
    size_t i = ck.a;
    while (i < ck.b) {
        // Equally spaced regions with slack behind the produced bytes (compress slots): one strided copy per group of
        // items, as wide as the group's longest item -- the slack (more than half of a 76 496-byte slot at ratio 0.5)
        // stays off PCIe.  A group grows while the bytes it copies beyond the items' lengths stay below 25 %.  Bytes of
        // a region between out_len and that width are unspecified (device scratch).
        if (i + 1 < ck.b && out_off[i + 1] > out_off[i] + out_len[i]) {
            const uint64_t pitch = out_off[i + 1] - out_off[i];
            size_t j = i;
            uint64_t sum = out_len[i];
            uint32_t width = out_len[i], cap_min = out_cap[i];
            while (j + 1 < ck.b && out_off[j + 1] - out_off[j] == pitch) {
                const uint32_t w2 = std::max(width, out_len[j + 1]);
                const uint64_t s2 = sum + out_len[j + 1];
                if ((uint64_t)w2 * (j + 2 - i) > s2 + s2 / 4 + 65536) break;
                width = w2, sum = s2, cap_min = std::min(cap_min, out_cap[j + 1]);
                j++;
            }
            if (j > i && pitch >= width && cap_min >= width) {
                if (width)
                    CU(cudaMemcpy2DAsync(out_base + out_off[i], pitch, (const uint8_t *)sl.d_out.p + (out_off[i] - ck.so.lo), pitch,
                                         width, j - i + 1, cudaMemcpyDeviceToHost, sl.stream));
                i = j + 1;
                continue;
            }
        }
        size_t j = i;
        while (j + 1 < ck.b && out_off[j + 1] == out_off[j] + out_cap[j]) j++;
        uint64_t lo = out_off[i], hi = out_off[j] + out_len[j];
        if (hi > lo)
            CU(cudaMemcpyAsync(out_base + lo, (const uint8_t *)sl.d_out.p + (lo - ck.so.lo), hi - lo,
                               cudaMemcpyDeviceToHost, sl.stream));
        i = j + 1;
    }
    return SNP_OK;
}

int run_host_batch(snp_ctx *c, bool compress, const uint8_t *in_base, const uint64_t *in_off,
                   const uint32_t *in_len, uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap,
                   uint32_t *out_len, int32_t *status, size_t n, uint32_t hash_mode) {
    if (n == 0) return SNP_OK;
    for (auto &sl : c->slots) {
        if (!sl.stream) CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (!sl.meta_ready) CU(cudaEventCreateWithFlags(&sl.meta_ready, cudaEventDisableTiming));
    }
    // per-chunk span target: the output span of a decompress chunk (SNP_HOST_CHUNK_MB), the input span of a compress chunk
    // (SNP_HOST_COMP_CHUNK_MB; its output span -- slots with slack -- is allowed to be proportionally larger)
    const uint64_t kChunkBytes = compress ? c->host_comp_chunk_bytes : c->host_chunk_bytes;
    constexpr size_t kChunkItems = 16384;
    // Phase 2 of a chunk (wait for its lengths / statuses, copy them to the caller, enqueue the payload copy when it was
    // not enqueued early) runs kLag chunks behind phase 1: waiting for chunk k-1's kernel before enqueueing chunk k+1
    // limited the pipeline to two chunks in flight, and a chunk's kernel is latency-bound (~1 ms: one block per warp).
    // A slot is reused kSlots chunks later, i.e. kSlots - kLag chunks after its phase 2.
    constexpr size_t kLag = snp_ctx::kSlots / 2;
    std::deque<Chunk> pend;
    int rc = SNP_OK, k = 0;
    size_t a = 0;
    while (a < n && rc == SNP_OK) {
        Chunk ck;
        ck.a = a;
        ck.slot = k++ % snp_ctx::kSlots;
        uint64_t ilo = UINT64_MAX, ihi = 0, olo = UINT64_MAX, ohi = 0;
        size_t b = a;
        // ramp up: the first chunks are small so that the first payload copy starts early (the D2H engine is the bottleneck)
        const uint64_t limit = k <= 3 ? std::max<uint64_t>(kChunkBytes >> (4 - k), 1 << 20) : kChunkBytes;
        while (b < n && b - a < kChunkItems) {
            uint64_t nilo = std::min(ilo, in_off[b]), nihi = std::max(ihi, in_off[b] + in_len[b]);
            uint64_t nolo = std::min(olo, out_off[b]), nohi = std::max(ohi, out_off[b] + out_cap[b]);
            if (b > a && (nihi - nilo > limit || (!compress && nohi - nolo > limit))) break;
            ilo = nilo, ihi = nihi, olo = nolo, ohi = nohi;
            b++;
        }
        ck.b = b;
        ck.si.lo = ilo, ck.si.hi = ihi, ck.so.lo = olo, ck.so.hi = ohi;
        rc = chunk_phase1(c, ck, compress, in_base, in_off, in_len, out_base, out_off, out_cap, hash_mode);
        pend.push_back(ck);
        if (rc == SNP_OK && pend.size() > kLag) {
            rc = chunk_phase2(c, pend.front(), out_base, out_off, out_cap, out_len, status);
            pend.pop_front();
        }
        a = b;
    }
    while (rc == SNP_OK && !pend.empty()) {
        rc = chunk_phase2(c, pend.front(), out_base, out_off, out_cap, out_len, status);
        pend.pop_front();
    }
    for (auto &sl : c->slots) {
        cudaError_t e = cudaStreamSynchronize(sl.stream);
        if (e != cudaSuccess && rc == SNP_OK) rc = cuda_fail(e, "cudaStreamSynchronize(slot)", __LINE__);
    }
    if (c->host_trace && !c->trace.empty()) {  // ms since the first chunk's start: h2d begin/end, kernel end, d2h begin/end
        cudaEvent_t t0 = c->trace[0].e[0];
        for (size_t i = 0; i < c->trace.size(); i++) {
            float v[5] = {};
            for (int j = 0; j < 5; j++)
                if (c->trace[i].e[j] && cudaEventElapsedTime(&v[j], t0, c->trace[i].e[j]) != cudaSuccess) {
                    v[j] = 0;
                    cudaGetLastError();  // an event that was never recorded (no payload copy): not an error of the call
                }
            fprintf(stderr, "chunk %3zu  h2d %7.3f..%7.3f  kernel ..%7.3f  d2h %7.3f..%7.3f\n", i, v[0], v[1], v[2], v[3], v[4]);
        }
        for (auto &r : c->trace)
            for (auto e : r.e)
                if (e) cudaEventDestroy(e);
        c->trace.clear();
    }
    return rc;
}

// The single-call API (snp_compress, snp_decompress, snp_frame_*: what sits behind Snappy.Compress / Decompress, called
// from arbitrary thread-pool threads) shares ONE context per device, created on first use and alive until the process
// exits.  Calls serialise on the context's mutex -- a call fills the GPU by itself -- so device scratch does not grow with
// the number of calling threads.  Footprint: the pipeline slots' buffers follow the largest call so far (about 2.2 x the
// chunk size per slot in use, 64 MiB chunks by default); hash tables follow the largest compress launch (64 KiB per
// concurrently compressed 64 KiB block, 512 KiB for a one-block call).
constexpr int kMaxDevices = 64;
std::mutex g_default_mu;
snp_ctx *g_default_ctx[kMaxDevices];  // never destroyed: the driver reclaims the memory at process exit

int default_ctx(snp_ctx **out) {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cuda_fail(e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount", __LINE__);
        return SNP_E_NO_DEVICE;
    }
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return SNP_E_INVALID_ARG;
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx[dev]) {
        snp_ctx *c = nullptr;
        int rc = snp_create(dev, &c);
        if (rc) return rc;
        g_default_ctx[dev] = c;
    }
    *out = g_default_ctx[dev];
    return SNP_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool overlaps(const void *a, size_t na, const void *b, size_t nb) {
    uintptr_t a0 = (uintptr_t)a, b0 = (uintptr_t)b;
    return na && nb && a0 < b0 + nb && b0 < a0 + na;
}

int host_varint_read(const uint8_t *in, size_t n, uint32_t *v, int *used) {
    // VarIntEncoding.Read.cs:38-79
    uint32_t result = 0;
    int shift = 0;
    *v = 0;
    *used = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t c = in[i], val = c & 0x7f;
        if (val & ~(0xffffffffu >> shift)) return SNP_INVALID_LENGTH;
        result |= val << shift;
        shift += 7;
        if (c < 128) {
            *v = result;
            *used = (int)i + 1;
            return SNP_OK;
        }
        if (shift >= 32) return SNP_INVALID_LENGTH;
    }
    return SNP_INCOMPLETE;
}

// varint(n) ++ concat(CompressFragment(fragment_i)): the body of SnappyCompressor.TryCompress (SnappyCompressor.cs:24-83)
// for an explicit fragment partition.  The input is given as host segments (copied back to back into device memory);
// frag_len partitions their concatenation.  Caller holds c->mu.
int compress_fragments_locked(snp_ctx *c, const uint8_t *const *seg_ptr, const size_t *seg_len, size_t n_seg, size_t n,
                              const std::vector<uint32_t> &frag_len, uint8_t *out, size_t cap, size_t *written,
                              uint32_t hash_mode) {
    cudaStream_t s = c->stream;
    int rc;
    const size_t nfrag = frag_len.size();
    const size_t pitch = (size_t)snp_max_compressed_length(SNP_BLOCK_SIZE) + 5;  // 76496, 16-byte multiple
    // Snappy's varint header, encoded on the host (VarIntEncoding.Write.cs:5-79).
    uint8_t hdr[5];
    size_t hdr_len = 0;
    {
        uint32_t v = (uint32_t)n;
        while (v >= 0x80) hdr[hdr_len++] = (uint8_t)(v | 0x80), v >>= 7;
        hdr[hdr_len++] = (uint8_t)v;
    }
    if (cap < hdr_len) return SNP_OUTPUT_TOO_SMALL;
    if (nfrag == 0) {  // empty input: the header alone ("" -> 00)
        memcpy(out, hdr, hdr_len);
        *written = hdr_len;
        return SNP_OK;
    }
    MetaLayout ml(nfrag);
    const size_t scan_bytes = align_up(nfrag * 8, 256) + 256;
    if ((rc = c->d_in.reserve(n + 16))) return rc;
    if ((rc = c->d_tmp.reserve(nfrag * pitch))) return rc;
    if ((rc = c->d_meta.reserve(ml.bytes + scan_bytes))) return rc;
    uint8_t *dm = (uint8_t *)c->d_meta.p;
    std::vector<uint64_t> off(nfrag), slot(nfrag);
    std::vector<uint32_t> capv(nfrag, (uint32_t)pitch);
    uint64_t run = 0;
    for (size_t f = 0; f < nfrag; f++) {
        off[f] = run;
        slot[f] = f * pitch;
        run += frag_len[f];
    }
    CU(cudaMemcpyAsync(dm + ml.in_off, off.data(), nfrag * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dm + ml.out_off, slot.data(), nfrag * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dm + ml.in_len, frag_len.data(), nfrag * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dm + ml.out_cap, capv.data(), nfrag * 4, cudaMemcpyHostToDevice, s));
    {
        size_t o = 0;
        for (size_t i = 0; i < n_seg; i++) {
            if (seg_len[i]) CU(cudaMemcpyAsync((uint8_t *)c->d_in.p + o, seg_ptr[i], seg_len[i], cudaMemcpyHostToDevice, s));
            o += seg_len[i];
        }
    }
    auto *d_len = (uint32_t *)(dm + ml.out_len);
    auto *d_status = (int32_t *)(dm + ml.status);
    auto *d_scan = (uint64_t *)(dm + ml.bytes);
    auto *d_total = (uint64_t *)(dm + ml.bytes + align_up(nfrag * 8, 256));
    rc = launch_compress(c, s, (const uint8_t *)c->d_in.p, (const uint64_t *)(dm + ml.in_off),
                         (const uint32_t *)(dm + ml.in_len), (uint8_t *)c->d_tmp.p,
                         (const uint64_t *)(dm + ml.out_off), (const uint32_t *)(dm + ml.out_cap), d_len, d_status,
                         nfrag, hash_mode, 1);
    if (rc) return rc;
    k_frag_scan<<<1, 1024, 0, s>>>(d_len, d_status, d_scan, d_total, nfrag);
    c->launches++;
    uint64_t total_bad[2];
    CU(cudaMemcpyAsync(total_bad, d_total, 16, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (total_bad[1]) {
        g_last_error = "internal: fragment compress reported a non-OK status";
        return SNP_E_CUDA;
    }
    size_t total = hdr_len + (size_t)total_bad[0];
    if (total > cap) return SNP_OUTPUT_TOO_SMALL;  // SnappyCompressor.cs:63-68 (bytesWritten = 0)
    if ((rc = c->d_out.reserve(total_bad[0] + 16))) return rc;
    unsigned grid = (unsigned)(nfrag < 4096 ? nfrag : 4096);
    k_frag_gather<<<grid, 256, 0, s>>>((const uint8_t *)c->d_tmp.p, pitch, d_len, d_scan, (uint8_t *)c->d_out.p,
                                       nfrag);
    c->launches++;
    CU(cudaGetLastError());
    memcpy(out, hdr, hdr_len);
    CU(cudaMemcpyAsync(out + hdr_len, c->d_out.p, total_bad[0], cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *written = total;
    return SNP_OK;
}

}  // namespace

// ------------------------------------------------------------------ C ABI ----

// No C++ exception may cross the C ABI (P/Invoke, ctypes): every entry point that can allocate is a function-try-block.
#define SNP_ABI_CATCH                                              \
    catch (const std::bad_alloc &) {                               \
        g_last_error = "out of host memory";                       \
        return SNP_E_NOMEM;                                        \
    }                                                              \
    catch (const std::exception &e) {                              \
        g_last_error = std::string("internal error: ") + e.what(); \
        return SNP_E_INTERNAL;                                     \
    }                                                              \
    catch (...) {                                                  \
        g_last_error = "internal error";                           \
        return SNP_E_INTERNAL;                                     \
    }

extern "C" {

int snp_abi_version(void) { return SNP_ABI_VERSION; }

const char *snp_status_string(int st) {
    switch (st) {
        case SNP_OK: return "OK";
        case SNP_OUTPUT_TOO_SMALL: return "Output buffer is too small.";
        case SNP_INVALID_LENGTH: return "Invalid stream length";
        case SNP_INCOMPLETE: return "Incomplete Snappy block.";
        case SNP_INVALID_COPY_OFFSET: return "Invalid copy offset";
        case SNP_DATA_TOO_LONG: return "Data too long";
        case SNP_UNKNOWN_CHUNK_TYPE: return "Unknown chunk type";
        case SNP_CRC_MISMATCH: return "Chunk CRC mismatch.";
        case SNP_E_CUDA: return "CUDA error";
        case SNP_E_INVALID_ARG: return "invalid argument";
        case SNP_E_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case SNP_E_NOMEM:
            return "out of host memory";
        case SNP_E_INTERNAL:
            return "internal error";
        case SNP_E_OVERLAP: return "Input and output spans must not overlap.";
        default: return "unknown status";
    }
}

const char *snp_last_error(void) { return g_last_error.c_str(); }

int32_t snp_max_compressed_length(int32_t n) { return 32 + n + n / 6 + 1; }
int32_t snp_get_max_compressed_length(int32_t n) { return snp_max_compressed_length(n) + 5; }

int snp_uncompressed_length(const uint8_t *in, size_t n, uint32_t *len) {
    if (!len || (!in && n)) return SNP_E_INVALID_ARG;
    int used;
    int st = host_varint_read(in, n, len, &used);
    if (st != SNP_OK || *len > 0x7fffffffu) {
        *len = 0;
        return SNP_INVALID_LENGTH;
    }
    return SNP_OK;
}

int snp_create(int device, snp_ctx **out) try {
    if (!out) return SNP_E_INVALID_ARG;
    *out = nullptr;
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cuda_fail(e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount", __LINE__);
        return SNP_E_NO_DEVICE;
    }
    if (device < 0 || device >= cnt) return SNP_E_INVALID_ARG;
    DeviceGuard g(device);
    std::unique_ptr<snp_ctx> c(new snp_ctx);
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
        g_last_error = "snappier_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                       std::to_string(prop.minor);
        return SNP_E_NO_DEVICE;
    }
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // data-independent device tables: CRC32C slicing tables, the compressor's probe schedule
    snp::k_init_crc_tables<<<4, 256, 0, c->stream>>>();
    snp::k_init_probe_sched<<<1, 32, 0, c->stream>>>();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->decomp_kernel = env_int("SNP_DECOMP_KERNEL", 7);
    c->v7_window = env_int("SNP_V7_WINDOW", 4096);
    c->v8_cfg = env_int("SNP_V8_CFG", 0);
    c->comp_kernel = env_int("SNP_COMP_KERNEL", 3);
    c->comp_ctas_per_sm = std::max(1, std::min(8, env_int("SNP_COMP_CTAS_PER_SM", 8)));
    c->comp_first_width = std::max(1, std::min(32, env_int("SNP_COMP_FIRST_WIDTH", 16)));
    c->host_chunk_bytes = (uint64_t)std::max(1, env_int("SNP_HOST_CHUNK_MB", 64)) << 20;
    c->host_comp_chunk_bytes = (uint64_t)std::max(1, env_int("SNP_HOST_COMP_CHUNK_MB", 256)) << 20;
    c->host_early_d2h = env_int("SNP_HOST_EARLY_D2H", 1);
    c->host_trace = env_int("SNP_HOST_TRACE", 0);
    *out = c.release();
    return SNP_OK;
} SNP_ABI_CATCH

void snp_destroy(snp_ctx *c) {
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->tables_done) cudaEventDestroy(c->tables_done);
    for (auto &sl : c->slots) {
        if (sl.stream) cudaStreamSynchronize(sl.stream), cudaStreamDestroy(sl.stream);
        if (sl.meta_ready) cudaEventDestroy(sl.meta_ready);
    }
    delete c;  // DevBuf destructors free the scratch on c->device
}

int snp_ctx_device(const snp_ctx *c) { return c ? c->device : -1; }
uint64_t snp_ctx_launch_count(const snp_ctx *c) { return c ? c->launches.load() : 0; }

int snp_compress_batch(snp_ctx *c, const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                       uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                       int32_t *status, size_t n, uint32_t hash_mode, int mem_kind, void *stream) try {
    if (hash_mode > SNP_HASH_MUL || (mem_kind != SNP_MEM_HOST && mem_kind != SNP_MEM_DEVICE))
        return SNP_E_INVALID_ARG;
    if (n && (!in_off || !in_len || !out_off || !out_cap || !out_len || !status || !out_base))
        return SNP_E_INVALID_ARG;
    int rc;
    if (!c && (rc = default_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (mem_kind == SNP_MEM_DEVICE)
        return launch_compress(c, (cudaStream_t)stream, in_base, in_off, in_len, out_base,
                               out_off, out_cap, out_len, status, n, hash_mode, 0);
    return run_host_batch(c, true, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n,
                          hash_mode);
} SNP_ABI_CATCH

int snp_decompress_batch(snp_ctx *c, const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                         int32_t *status, size_t n, int mem_kind, void *stream) try {
    if (mem_kind != SNP_MEM_HOST && mem_kind != SNP_MEM_DEVICE) return SNP_E_INVALID_ARG;
    if (n && (!in_base || !in_off || !in_len || !out_off || !out_cap || !out_len || !status))
        return SNP_E_INVALID_ARG;
    int rc;
    if (!c && (rc = default_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (mem_kind == SNP_MEM_DEVICE)
        return launch_decompress(c, (cudaStream_t)stream, in_base, in_off, in_len,
                                 out_base, out_off, out_cap, out_len, status, n);
    return run_host_batch(c, false, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, 0);
} SNP_ABI_CATCH

int snp_uncompressed_length_batch(snp_ctx *c, const uint8_t *in_base, const uint64_t *in_off,
                                  const uint32_t *in_len, uint32_t *ulen, int32_t *status, size_t n,
                                  int mem_kind, void *stream) try {
    if (mem_kind != SNP_MEM_HOST && mem_kind != SNP_MEM_DEVICE) return SNP_E_INVALID_ARG;
    if (n && (!in_base || !in_off || !in_len || !ulen || !status)) return SNP_E_INVALID_ARG;
    if (n == 0) return SNP_OK;
    if (mem_kind == SNP_MEM_HOST) {
        // 1..5 bytes per item: not worth a PCIe round trip (same host varint the
        // single-call path uses).
        for (size_t i = 0; i < n; i++) status[i] = snp_uncompressed_length(in_base + in_off[i], in_len[i], &ulen[i]);
        return SNP_OK;
    }
    int rc;
    if (!c && (rc = default_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    snp::k_uncompressed_length<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in_base, in_off, in_len, ulen, status, n);
    c->launches++;
    CU(cudaGetLastError());
    return SNP_OK;
} SNP_ABI_CATCH

int snp_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written, uint32_t hash_mode) try {
    if (!written || (!in && n) || (!out && cap) || hash_mode > SNP_HASH_MUL || n > 0xffffffffull)
        return SNP_E_INVALID_ARG;
    *written = 0;
    if (overlaps(in, n, out, cap)) return SNP_E_OVERLAP;  // SnappyCompressor.cs:27-30
    if (cap == 0) return SNP_OUTPUT_TOO_SMALL;            // Snappy.cs:57-62
    snp_ctx *c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStream_t s = c->stream;

    if (n <= SNP_BLOCK_SIZE) {  // one fragment: varint + fragment by one warp
        uint64_t zero = 0;
        uint32_t in_len = (uint32_t)n, out_len = 0;
        uint32_t need = (uint32_t)snp_get_max_compressed_length((int32_t)n);
        uint32_t out_cap = cap < need ? (uint32_t)cap : need;
        int32_t st = 0;
        rc = run_host_batch(c, true, in, &zero, &in_len, out, &zero, &out_cap, &out_len, &st, 1, hash_mode);
        if (rc) return rc;
        *written = st == SNP_OK ? out_len : 0;
        return st;
    }

    // SnappyCompressor.cs:40-80: independent 64 KiB fragments, then concatenation.
    const size_t nfrag = (n + SNP_BLOCK_SIZE - 1) / SNP_BLOCK_SIZE;
    std::vector<uint32_t> len(nfrag);
    for (size_t f = 0; f < nfrag; f++) {
        const size_t o = f * (size_t)SNP_BLOCK_SIZE;
        len[f] = (uint32_t)(n - o < SNP_BLOCK_SIZE ? n - o : SNP_BLOCK_SIZE);
    }
    const uint8_t *one_ptr[1] = {in};
    const size_t one_len[1] = {n};
    return compress_fragments_locked(c, one_ptr, one_len, 1, n, len, out, cap, written, hash_mode);
} SNP_ABI_CATCH

int snp_compress_sequence(const uint8_t *const *seg_ptr, const size_t *seg_len, size_t n_seg, uint8_t *out, size_t cap,
                          size_t *written, uint32_t hash_mode) try {
    if (!written || (n_seg && (!seg_ptr || !seg_len)) || (!out && cap) || hash_mode > SNP_HASH_MUL)
        return SNP_E_INVALID_ARG;
    *written = 0;
    size_t n = 0;
    for (size_t i = 0; i < n_seg; i++) {
        if (!seg_ptr[i] && seg_len[i]) return SNP_E_INVALID_ARG;
        if (overlaps(seg_ptr[i], seg_len[i], out, cap)) return SNP_E_OVERLAP;
        n += seg_len[i];
    }
    if (n > 0xffffffffull) return SNP_E_INVALID_ARG;  // SnappyCompressor.cs:88-91
    if (cap == 0) return SNP_OUTPUT_TOO_SMALL;
    // SnappyCompressor.cs:103-143: the next fragment is the first segment's part of the next <= 64 KiB when the
    // fragment is contiguous or that part is >= 32 KiB, otherwise the whole (copied) fragment -- so the fragment
    // boundaries, hence the compressed bytes, depend on the segmentation.
    std::vector<uint32_t> len;
    {
        size_t i = 0, o = 0, left = n;
        while (left) {
            while (i < n_seg && o == seg_len[i]) i++, o = 0;  // empty segments carry no bytes
            const size_t frag = left < SNP_BLOCK_SIZE ? left : SNP_BLOCK_SIZE;
            const size_t first = seg_len[i] - o < frag ? seg_len[i] - o : frag;
            const size_t take = (first == frag || first >= SNP_BLOCK_SIZE / 2) ? first : frag;
            len.push_back((uint32_t)take);
            left -= take;
            size_t adv = take;
            while (adv) {
                const size_t step = seg_len[i] - o < adv ? seg_len[i] - o : adv;
                o += step;
                adv -= step;
                if (o == seg_len[i] && adv) i++, o = 0;
            }
        }
    }
    snp_ctx *c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return compress_fragments_locked(c, seg_ptr, seg_len, n_seg, n, len, out, cap, written, hash_mode);
} SNP_ABI_CATCH

int snp_decompress_sequence(const uint8_t *const *seg_ptr, const size_t *seg_len, size_t n_seg, uint8_t *out,
                            size_t cap, size_t *written) try {
    // Snappy.Decompress(ReadOnlySequence<byte>, ..) feeds the segments to one decoder in order (Snappy.cs:194-212,
    // 246-261): the result is that of decoding their concatenation.  The batch engine needs whole blocks, so the
    // segments are joined on the host first (the resumable split-input state machine stays out of scope).
    if (!written || (n_seg && (!seg_ptr || !seg_len)) || (!out && cap)) return SNP_E_INVALID_ARG;
    *written = 0;
    size_t n = 0;
    for (size_t i = 0; i < n_seg; i++) {
        if (!seg_ptr[i] && seg_len[i]) return SNP_E_INVALID_ARG;
        n += seg_len[i];
    }
    if (n_seg == 1) return snp_decompress(seg_ptr[0], seg_len[0], out, cap, written);
    std::vector<uint8_t> joined(n);
    size_t o = 0;
    for (size_t i = 0; i < n_seg; i++) {
        if (seg_len[i]) memcpy(joined.data() + o, seg_ptr[i], seg_len[i]);
        o += seg_len[i];
    }
    return snp_decompress(joined.data(), n, out, cap, written);
} SNP_ABI_CATCH

int snp_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written) try {
    if (!written || (!in && n) || (!out && cap) || n > 0xffffffffull) return SNP_E_INVALID_ARG;
    *written = 0;
    uint32_t U;
    int used;
    int st = host_varint_read(in, n, &U, &used);  // SnappyDecompressor.cs:50-63
    if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;
    if (st != SNP_OK || U > 0x7fffffffu) return SNP_INVALID_LENGTH;
    snp_ctx *c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    uint64_t zero = 0;
    uint32_t in_len = (uint32_t)n, out_len = 0, out_cap = U;
    int32_t bst = 0;
    if (cap >= U) {
        rc = run_host_batch(c, false, in, &zero, &in_len, out, &zero, &out_cap, &out_len, &bst, 1, 0);
        if (rc) return rc;
        *written = bst == SNP_OK ? out_len : 0;
        return bst;
    }
    // Caller's buffer is smaller than the declared length.  The reference still
    // decodes the whole block into its own buffer first, so data errors win over
    // "too small", and Read() then hands back the first `cap` bytes
    // (Snappy.cs:172-186, SnappyDecompressor.cs:613-629).
    // The declared length comes from an untrusted header: a block of n bytes cannot produce more than 64 bytes per 3-byte
    // copy tag, and the pipeline only ever copies the bytes a block produced, so the private buffer is bounded by the
    // input size (a 13-byte block that announces 2 GiB costs 350 bytes, not 2 GiB); no zero fill.
    const size_t need = std::max<size_t>(std::min<uint64_t>(U, 22ull * n + 64), 1);
    std::unique_ptr<uint8_t, void (*)(void *)> full((uint8_t *)malloc(need), free);
    if (!full) {
        g_last_error = "out of host memory";
        return SNP_E_NOMEM;
    }
    rc = run_host_batch(c, false, in, &zero, &in_len, full.get(), &zero, &out_cap, &out_len, &bst, 1, 0);
    if (rc) return rc;
    if (bst != SNP_OK) return bst;
    memcpy(out, full.get(), cap);
    *written = cap;
    return SNP_OUTPUT_TOO_SMALL;
} SNP_ABI_CATCH

// ------------------------------------------------------------- framing format --

size_t snp_frame_max_compressed_length(size_t n) { return 10 + n + 8 * ((n + SNP_BLOCK_SIZE - 1) / SNP_BLOCK_SIZE); }

int snp_pack_batch(snp_ctx *c, const uint8_t *src_base, const uint64_t *src_off, const uint32_t *len, size_t n,
                   uint8_t *dst_base, uint64_t *dst_off, uint64_t *total, void *stream) try {
    if (!c || (n && (!src_base || !src_off || !len || !dst_off)) || !total) return SNP_E_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        CU(cudaMemsetAsync(total, 0, 8, s));
        return SNP_OK;
    }
    const size_t nparts = (n + kScanItems - 1) / kScanItems;
    // the partial sums live in the context's scratch: launches on different streams are chained like the compress tables
    int rc;
    if ((rc = c->d_tmp.reserve(nparts * 8))) return rc;
    uint64_t *part = (uint64_t *)c->d_tmp.p;
    k_pack_sums<<<(unsigned)nparts, 256, 0, s>>>(len, part, n);
    k_pack_scan_sums<<<1, 1024, 0, s>>>(part, nparts, total);
    k_pack_offsets<<<(unsigned)nparts, 256, 0, s>>>(len, part, dst_off, n);
    if (dst_base) {
        const unsigned grid = (unsigned)std::min((n + 7) / 8, (size_t)c->sm_count * 8);
        k_pack_copy<<<grid, 256, 0, s>>>(src_base, src_off, len, dst_base, dst_off, n);
        c->launches++;
    }
    c->launches += 3;
    CU(cudaGetLastError());
    return SNP_OK;
} SNP_ABI_CATCH

int snp_find_match_length_batch(snp_ctx *c, const uint8_t *base, const uint32_t *s1, const uint32_t *s2,
                                const uint32_t *s2_limit, uint32_t *matched, size_t n, void *stream) try {
    if (!c || (n && (!base || !s1 || !s2 || !s2_limit || !matched))) return SNP_E_INVALID_ARG;
    if (n == 0) return SNP_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    const unsigned grid = (unsigned)std::min((n + 7) / 8, (size_t)c->sm_count * 8);
    k_find_match_length<<<grid, 256, 0, (cudaStream_t)stream>>>(base, s1, s2, s2_limit, matched, n);
    c->launches++;
    CU(cudaGetLastError());
    return SNP_OK;
} SNP_ABI_CATCH

int snp_diag_random_reads(snp_ctx *c, const uint8_t *base, size_t span_bytes, uint32_t ctas_per_sm, uint32_t reads_per_thread,
                          uint32_t *sink, void *stream) try {
    if (!c || !base || span_bytes < 16 || !sink || ctas_per_sm == 0 || ctas_per_sm > 8) return SNP_E_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    k_diag_random_reads<<<(unsigned)(c->sm_count * ctas_per_sm), 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)base, (unsigned long long)(span_bytes / 16), reads_per_thread & ~3u, sink);
    CU(cudaGetLastError());
    return SNP_OK;
} SNP_ABI_CATCH

int snp_crc32c_batch(snp_ctx *c, const uint8_t *base, const uint64_t *off, const uint32_t *len, uint32_t *crc,
                     size_t n, int masked, int mem_kind, void *stream) try {
    if (mem_kind != SNP_MEM_HOST && mem_kind != SNP_MEM_DEVICE) return SNP_E_INVALID_ARG;
    if (n && (!base || !off || !len || !crc)) return SNP_E_INVALID_ARG;
    if (n == 0) return SNP_OK;
    int rc;
    if (!c && (rc = default_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if ((rc = ctx_set_attrs(c))) return rc;
    unsigned grid = (unsigned)std::min((n + 7) / 8, (size_t)c->sm_count * 8);
    if (mem_kind == SNP_MEM_DEVICE) {
        snp::k_crc32c_masked_batch<<<grid, 256, 0, (cudaStream_t)stream>>>(base, off, len, crc, n, masked);
        c->launches++;
        CU(cudaGetLastError());
        return SNP_OK;
    }
    cudaStream_t s = c->stream;
    uint64_t lo = UINT64_MAX, hi = 0;
    for (size_t i = 0; i < n; i++) lo = std::min(lo, off[i]), hi = std::max(hi, off[i] + len[i]);
    std::vector<uint64_t> rel(n);
    for (size_t i = 0; i < n; i++) rel[i] = off[i] - lo;
    size_t mo = align_up(n * 8, 256), ml = align_up(n * 4, 256);
    if ((rc = c->d_in.reserve(hi - lo + 16))) return rc;
    if ((rc = c->d_meta.reserve(mo + 2 * ml))) return rc;
    uint8_t *dm = (uint8_t *)c->d_meta.p;
    CU(cudaMemcpyAsync(dm, rel.data(), n * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dm + mo, len, n * 4, cudaMemcpyHostToDevice, s));
    if (hi > lo) CU(cudaMemcpyAsync(c->d_in.p, base + lo, hi - lo, cudaMemcpyHostToDevice, s));
    snp::k_crc32c_masked_batch<<<grid, 256, 0, s>>>((const uint8_t *)c->d_in.p, (const uint64_t *)dm,
                                                    (const uint32_t *)(dm + mo), (uint32_t *)(dm + mo + ml), n, masked);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(crc, dm + mo + ml, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SNP_OK;
} SNP_ABI_CATCH

int snp_frame_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written, uint32_t hash_mode) try {
    static const uint8_t kStreamId[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
    if (!written || (!in && n) || (!out && cap) || hash_mode > SNP_HASH_MUL) return SNP_E_INVALID_ARG;
    *written = 0;
    if (overlaps(in, n, out, cap)) return SNP_E_OVERLAP;
    if (cap < 10) return SNP_OUTPUT_TOO_SMALL;
    if (n == 0) {  // Write(empty) still emits the stream identifier (SnappyStreamCompressor.cs:40-47,148-157)
        memcpy(out, kStreamId, 10);
        *written = 10;
        return SNP_OK;
    }
    snp_ctx *c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    for (auto &sl : c->slots) {
        if (!sl.stream) CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (!sl.meta_ready) CU(cudaEventCreateWithFlags(&sl.meta_ready, cudaEventDisableTiming));
    }
    // The stream is cut into pieces of kPiece chunks that flow through the pipeline slots like the chunks of the batched
    // host-mode calls: H2D of piece k+1 | compress + CRC + frame kernels of piece k | D2H of piece k-1.  The reference
    // does the same thing one 64 KiB chunk at a time (SnappyStreamCompressor.cs:166-230).  A piece's framed bytes are
    // dense in its slot's device buffer; its position in the caller's buffer is the running total of the pieces before it.
    const size_t nch = (n + SNP_BLOCK_SIZE - 1) / SNP_BLOCK_SIZE;
    // A piece's compress kernel lasts at least as long as its slowest block (tens of ms for dense-match data: one warp per
    // block), so pieces must be large enough that the slots together hold a GPU-load of blocks: 4 x the copy-bound size.
    const size_t kPiece = std::max<size_t>(1, c->host_comp_chunk_bytes / SNP_BLOCK_SIZE);
    const size_t pitch = (size_t)snp_get_max_compressed_length(SNP_BLOCK_SIZE);
    constexpr size_t kLag = snp_ctx::kSlots / 2;
    struct Piece {
        size_t f0, f1;  // chunk range
        int slot;
        size_t total_off;  // where the piece's total lands in the slot's pinned metadata
    };
    std::deque<Piece> pend;
    size_t pos = 0;  // framed bytes handed to the caller so far
    auto phase2 = [&](const Piece &pc) -> int {
        snp_ctx::Slot &sl = c->slots[pc.slot];
        CU(cudaEventSynchronize(sl.meta_ready));
        const uint64_t *tb = (const uint64_t *)((const uint8_t *)sl.h_meta.p + pc.total_off);
        const size_t total = (size_t)tb[0] + (pc.f0 == 0 ? 10 : 0);
        if (pos + total > cap) return SNP_OUTPUT_TOO_SMALL;
        CU(cudaMemcpyAsync(out + pos, sl.d_out.p, total, cudaMemcpyDeviceToHost, sl.stream));
        pos += total;
        return SNP_OK;
    };
    int k = 0;
    for (size_t f0 = 0; f0 < nch && rc == SNP_OK; f0 += kPiece) {
        Piece pc;
        pc.f0 = f0;
        pc.f1 = std::min(nch, f0 + kPiece);
        pc.slot = k++ % snp_ctx::kSlots;
        const size_t m = pc.f1 - pc.f0;
        const size_t b0 = f0 * (size_t)SNP_BLOCK_SIZE, b1 = std::min(n, pc.f1 * (size_t)SNP_BLOCK_SIZE);
        snp_ctx::Slot &sl = c->slots[pc.slot];
        cudaStream_t s = sl.stream;
        CU(cudaStreamSynchronize(s));  // the slot's previous piece is completely done
        MetaLayout ml(m);
        const size_t a4 = align_up(m * 4, 256), a8 = align_up(m * 8, 256);
        const size_t o_crc = ml.bytes, o_sizes = o_crc + a4, o_scan = o_sizes + a4, o_total = o_scan + a8;
        pc.total_off = o_total;
        if ((rc = sl.d_in.reserve(b1 - b0 + 16))) break;
        if ((rc = sl.d_tmp.reserve(m * pitch))) break;
        if ((rc = sl.d_out.reserve(b1 - b0 + 8 * m + 32))) break;  // every chunk may fall back to raw
        if ((rc = sl.d_meta.reserve(o_total + 256))) break;
        if ((rc = sl.h_meta.reserve(o_total + 256))) break;
        uint8_t *dm = (uint8_t *)sl.d_meta.p, *hm = (uint8_t *)sl.h_meta.p;
        {
            uint64_t *ri = (uint64_t *)(hm + ml.in_off), *ro = (uint64_t *)(hm + ml.out_off);
            uint32_t *rl = (uint32_t *)(hm + ml.in_len), *rc_ = (uint32_t *)(hm + ml.out_cap);
            for (size_t i = 0; i < m; i++) {
                ri[i] = i * (uint64_t)SNP_BLOCK_SIZE;
                ro[i] = i * pitch;
                rl[i] = (uint32_t)std::min<size_t>(b1 - b0 - ri[i], SNP_BLOCK_SIZE);
                rc_[i] = (uint32_t)pitch;
            }
        }
        CU(cudaMemcpyAsync(dm, hm, ml.out_len, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(sl.d_in.p, in + b0, b1 - b0, cudaMemcpyHostToDevice, s));
        auto *d_raw_off = (const uint64_t *)(dm + ml.in_off);
        auto *d_raw_len = (const uint32_t *)(dm + ml.in_len);
        auto *d_comp_len = (uint32_t *)(dm + ml.out_len);
        auto *d_status = (int32_t *)(dm + ml.status);
        auto *d_crc = (uint32_t *)(dm + o_crc);
        auto *d_sizes = (uint32_t *)(dm + o_sizes);
        auto *d_scan = (uint64_t *)(dm + o_scan);
        auto *d_total = (uint64_t *)(dm + o_total);
        // every chunk is an independent Snappy.Compress (SnappyStreamCompressor.cs:206)
        rc = launch_compress(c, s, (const uint8_t *)sl.d_in.p, d_raw_off, d_raw_len, (uint8_t *)sl.d_tmp.p,
                             (const uint64_t *)(dm + ml.out_off), (const uint32_t *)(dm + ml.out_cap), d_comp_len, d_status, m,
                             hash_mode, 0, &sl.d_tables);
        if (rc) break;
        const unsigned grid = (unsigned)std::min((m + 7) / 8, (size_t)c->sm_count * 8);
        snp::k_crc32c_masked_batch<<<grid, 256, 0, s>>>((const uint8_t *)sl.d_in.p, d_raw_off, d_raw_len, d_crc, m, 1);
        snp::k_frame_plan<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(d_raw_len, d_comp_len, d_sizes, m);
        k_frag_scan<<<1, 1024, 0, s>>>(d_sizes, d_status, d_scan, d_total, m);
        snp::k_frame_emit<<<(unsigned)std::min<size_t>(m, 4096), 256, 0, s>>>(
            (const uint8_t *)sl.d_in.p, d_raw_off, d_raw_len, (const uint8_t *)sl.d_tmp.p, pitch, d_comp_len, d_crc, d_scan,
            (uint8_t *)sl.d_out.p, m, f0 == 0 ? 10u : 0u);
        c->launches += 4;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hm + o_total, d_total, 16, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(sl.meta_ready, s));
        pend.push_back(pc);
        if (pend.size() > kLag) {
            rc = phase2(pend.front());
            pend.pop_front();
        }
    }
    while (rc == SNP_OK && !pend.empty()) {
        rc = phase2(pend.front());
        pend.pop_front();
    }
    for (auto &sl : c->slots) {
        cudaError_t e = cudaStreamSynchronize(sl.stream);
        if (e != cudaSuccess && rc == SNP_OK) rc = cuda_fail(e, "cudaStreamSynchronize(slot)", __LINE__);
    }
    if (rc != SNP_OK) return rc;
    *written = pos;
    return SNP_OK;
} SNP_ABI_CATCH

namespace {
struct FrameChunk {
    uint8_t type;
    uint64_t body;  // offset of the payload (after the 4-byte CRC) in the framed stream
    uint32_t len;   // payload length
    uint32_t crc;   // expected masked CRC32C of the uncompressed chunk
    uint32_t ulen;  // uncompressed length
};
// Host-side chunk table (SnappyStreamDecompressor.ReadChunkHeader / ReadChunkCrc, :215-289).
int frame_scan(const uint8_t *in, size_t n, std::vector<FrameChunk> &chunks, uint64_t *total) {
    size_t i = 0;
    *total = 0;
    while (i < n) {
        if (n - i < 4) return SNP_INCOMPLETE;
        const uint8_t type = in[i];
        const size_t len = in[i + 1] | ((size_t)in[i + 2] << 8) | ((size_t)in[i + 3] << 16);
        i += 4;
        if (n - i < len) return SNP_INCOMPLETE;
        if (type == 0x00 || type == 0x01) {
            if (len < 4) return SNP_INCOMPLETE;
            FrameChunk ck;
            ck.type = type;
            ck.crc = in[i] | (in[i + 1] << 8) | (in[i + 2] << 16) | ((uint32_t)in[i + 3] << 24);
            ck.body = i + 4;
            ck.len = (uint32_t)(len - 4);
            if (type == 0x01) {
                ck.ulen = ck.len;
            } else {
                int used;
                int st = host_varint_read(in + ck.body, ck.len, &ck.ulen, &used);
                if (st == SNP_INCOMPLETE) return SNP_INCOMPLETE;
                if (st != SNP_OK || ck.ulen > 0x7fffffffu) return SNP_INVALID_LENGTH;
            }
            *total += ck.ulen;
            chunks.push_back(ck);
        } else if (type < 0x80) {
            return SNP_UNKNOWN_CHUNK_TYPE;  // :182-185
        }  // else: skippable (0x80..0xfe) or stream identifier (0xff): content not validated (:180-199)
        i += len;
    }
    return SNP_OK;
}
}  // namespace

int snp_frame_uncompressed_length(const uint8_t *in, size_t n, uint64_t *len) try {
    if (!len || (!in && n)) return SNP_E_INVALID_ARG;
    std::vector<FrameChunk> chunks;
    int st = frame_scan(in, n, chunks, len);
    if (st != SNP_OK) *len = 0;
    return st;
} SNP_ABI_CATCH

int snp_frame_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written) try {
    if (!written || (!in && n) || (!out && cap)) return SNP_E_INVALID_ARG;
    *written = 0;
    std::vector<FrameChunk> chunks;
    uint64_t total = 0;
    int st = frame_scan(in, n, chunks, &total);
    if (st != SNP_OK) return st;
    if (total > cap) return SNP_OUTPUT_TOO_SMALL;
    const size_t nch = chunks.size();
    if (nch == 0) return SNP_OK;
    snp_ctx *c;
    int rc = default_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    for (auto &sl : c->slots) {
        if (!sl.stream) CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (!sl.meta_ready) CU(cudaEventCreateWithFlags(&sl.meta_ready, cudaEventDisableTiming));
    }
    // Pieces of consecutive chunks flow through the pipeline slots: H2D of the piece's byte span | decode of its
    // compressed chunks + copies of its raw chunks + CRC32C of every chunk's output | D2H of the piece's output (enqueued
    // right behind the kernels: like the reference, which hands chunk after chunk to the caller before it meets a bad one,
    // SnappyStreamDecompressor.cs:38-208).  Statuses and CRCs are checked kLag pieces later, in stream order.
    constexpr size_t kLag = snp_ctx::kSlots / 2;
    struct Piece {
        size_t f0, f1, na;
        int slot;
        size_t o_aout, o_crc;  // offsets of the decode results / CRCs in the slot's pinned metadata
    };
    std::deque<Piece> pend;
    int verdict = SNP_OK;
    auto phase2 = [&](const Piece &pc) -> int {  // returns a call-level error; a bad chunk goes to `verdict`
        snp_ctx::Slot &sl = c->slots[pc.slot];
        CU(cudaEventSynchronize(sl.meta_ready));
        if (verdict != SNP_OK) return SNP_OK;
        const uint8_t *hm = (const uint8_t *)sl.h_meta.p;
        MetaLayout ml(pc.f1 - pc.f0);
        const int32_t *a_st = (const int32_t *)(hm + ml.status);
        const uint32_t *crc = (const uint32_t *)(hm + pc.o_crc);
        size_t ai = 0;
        for (size_t i = pc.f0; i < pc.f1; i++) {  // first bad chunk in stream order decides (the reference throws there)
            if (chunks[i].type == 0x00 && a_st[ai++] != SNP_OK) {
                verdict = a_st[ai - 1];
                return SNP_OK;
            }
            if (crc[i - pc.f0] != chunks[i].crc) {
                verdict = SNP_CRC_MISMATCH;
                return SNP_OK;
            }
        }
        return SNP_OK;
    };
    const uint64_t kSpan = c->host_chunk_bytes;
    uint64_t opos = 0;
    int k = 0;
    size_t f0 = 0;
    while (f0 < nch && rc == SNP_OK && verdict == SNP_OK) {
        Piece pc;
        pc.f0 = f0;
        size_t f1 = f0;
        uint64_t osz = 0;
        const uint64_t ilo = chunks[f0].body;
        while (f1 < nch && f1 - f0 < 16384) {
            const uint64_t ihi = chunks[f1].body + chunks[f1].len;
            if (f1 > f0 && (ihi - ilo > kSpan || osz + chunks[f1].ulen > kSpan)) break;
            osz += chunks[f1].ulen;
            f1++;
        }
        pc.f1 = f1;
        pc.slot = k++ % snp_ctx::kSlots;
        const size_t m = f1 - f0;
        const uint64_t ihi = chunks[f1 - 1].body + chunks[f1 - 1].len;
        snp_ctx::Slot &sl = c->slots[pc.slot];
        cudaStream_t s = sl.stream;
        CU(cudaStreamSynchronize(s));  // the slot's previous piece is completely done
        MetaLayout ml(m);
        const size_t a4 = align_up(m * 4, 256), a8 = align_up(m * 8, 256);
        const size_t o_boff = ml.bytes, o_blen = o_boff + a8, o_crc = o_blen + a4, o_end = o_crc + a4;
        pc.o_aout = ml.out_len;
        pc.o_crc = o_crc;
        if ((rc = sl.d_in.reserve(ihi - ilo + 16))) break;
        if ((rc = sl.d_out.reserve(osz + 16))) break;
        if ((rc = sl.d_meta.reserve(o_end))) break;
        if ((rc = sl.h_meta.reserve(o_end))) break;
        uint8_t *dm = (uint8_t *)sl.d_meta.p, *hm = (uint8_t *)sl.h_meta.p;
        size_t na = 0;
        {
            uint64_t *a_in = (uint64_t *)(hm + ml.in_off), *a_out = (uint64_t *)(hm + ml.out_off);
            uint32_t *a_len = (uint32_t *)(hm + ml.in_len), *a_cap = (uint32_t *)(hm + ml.out_cap);
            uint64_t *b_off = (uint64_t *)(hm + o_boff);
            uint32_t *b_len = (uint32_t *)(hm + o_blen);
            uint64_t op = 0;
            for (size_t i = f0; i < f1; i++) {
                b_off[i - f0] = op;
                b_len[i - f0] = chunks[i].ulen;
                if (chunks[i].type == 0x00) {
                    a_in[na] = chunks[i].body - ilo, a_len[na] = chunks[i].len;
                    a_out[na] = op, a_cap[na] = chunks[i].ulen;
                    na++;
                }
                op += chunks[i].ulen;
            }
        }
        pc.na = na;
        CU(cudaMemcpyAsync(dm, hm, ml.out_len, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dm + o_boff, hm + o_boff, a8 + a4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(sl.d_in.p, in + ilo, ihi - ilo, cudaMemcpyHostToDevice, s));
        if (na) {
            rc = launch_decompress(c, s, (const uint8_t *)sl.d_in.p, (const uint64_t *)(dm + ml.in_off),
                                   (const uint32_t *)(dm + ml.in_len), (uint8_t *)sl.d_out.p,
                                   (const uint64_t *)(dm + ml.out_off), (const uint32_t *)(dm + ml.out_cap),
                                   (uint32_t *)(dm + ml.out_len), (int32_t *)(dm + ml.status), na);
            if (rc) break;
        }
        {
            const uint64_t *b_off = (const uint64_t *)(hm + o_boff);
            for (size_t i = f0; i < f1; i++)  // type 0x01: raw copy (SnappyStreamDecompressor.cs:137-178)
                if (chunks[i].type == 0x01 && chunks[i].len)
                    CU(cudaMemcpyAsync((uint8_t *)sl.d_out.p + b_off[i - f0], (const uint8_t *)sl.d_in.p + (chunks[i].body - ilo),
                                       chunks[i].len, cudaMemcpyDeviceToDevice, s));
        }
        const unsigned grid = (unsigned)std::min((m + 7) / 8, (size_t)c->sm_count * 8);
        snp::k_crc32c_masked_batch<<<grid, 256, 0, s>>>((const uint8_t *)sl.d_out.p, (const uint64_t *)(dm + o_boff),
                                                        (const uint32_t *)(dm + o_blen), (uint32_t *)(dm + o_crc), m, 1);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hm + ml.out_len, dm + ml.out_len, ml.bytes - ml.out_len, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(hm + o_crc, dm + o_crc, a4, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(sl.meta_ready, s));
        if (osz) CU(cudaMemcpyAsync(out + opos, sl.d_out.p, osz, cudaMemcpyDeviceToHost, s));
        opos += osz;
        pend.push_back(pc);
        if (pend.size() > kLag) {
            rc = phase2(pend.front());
            pend.pop_front();
        }
        f0 = f1;
    }
    while (rc == SNP_OK && !pend.empty()) {
        rc = phase2(pend.front());
        pend.pop_front();
    }
    for (auto &sl : c->slots) {
        cudaError_t e = cudaStreamSynchronize(sl.stream);
        if (e != cudaSuccess && rc == SNP_OK) rc = cuda_fail(e, "cudaStreamSynchronize(slot)", __LINE__);
    }
    if (rc != SNP_OK) return rc;
    if (verdict != SNP_OK) return verdict;
    *written = (size_t)total;
    return SNP_OK;
} SNP_ABI_CATCH

}  // extern "C"
