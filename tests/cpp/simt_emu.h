// simt_emu.h -- a tiny single-warp SIMT emulator for host-side tests of the warp-level kernels.
//
// TEST INFRASTRUCTURE ONLY.  The device code of snappier_b200/csrc/snp_decompress_v6.cuh is
// written against CUDA's warp intrinsics; compiled with -DSNP_EMU it sees the definitions below
// instead, and every lane of ONE warp runs as a ucontext coroutine on one OS thread.  A
// collective (__shfl_sync, __ballot_sync, __syncwarp, ...) is a rendezvous of all 32 lanes; a
// lane that reaches one yields to the scheduler until everybody has arrived.  Divergent
// collectives (some lanes exit or wait at a different call) are detected as a deadlock and abort
// the test.  Memory is ordinary host memory, so the emulator checks LOGIC (parse, dependency
// rounds, window/flush bookkeeping), not the GPU memory model.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct uint2 {
    uint32_t x, y;
};
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace simt {

constexpr int kLanes = 32;
constexpr size_t kStack = 512 * 1024;

struct Warp {
    ucontext_t sched;
    ucontext_t lane[kLanes];
    bool done[kLanes];
    char *stacks = nullptr;
    int cur = 0;
    int arrived = 0;
    unsigned gen = 0;
    unsigned long progress = 0;
    uint64_t slot[2][kLanes];
    const char *where[kLanes];
    std::function<void()> body;
};

inline Warp *&active() {
    static Warp *w = nullptr;
    return w;
}

inline int lane() { return active()->cur; }

inline void yield_to_sched() {
    Warp *w = active();
    swapcontext(&w->lane[w->cur], &w->sched);
}

// Rendezvous of all 32 lanes.  Returns the generation index the rendezvous belonged to.
inline unsigned barrier(const char *what) {
    Warp *w = active();
    const unsigned g = w->gen;
    w->where[w->cur] = what;
    w->progress++;
    if (++w->arrived == kLanes) {
        w->arrived = 0;
        w->gen++;
        return g;
    }
    while (w->gen == g) yield_to_sched();
    return g;
}

inline void trampoline() {
    Warp *w = active();
    w->body();
    w->done[w->cur] = true;
    w->progress++;
    w->where[w->cur] = "exit";
    yield_to_sched();
}

// Runs `body` once per lane, as one warp.  Aborts on divergent collectives.
inline void run_warp(const std::function<void()> &body) {
    Warp w;
    w.body = body;
    w.stacks = (char *)malloc(kStack * kLanes);
    Warp *prev = active();
    active() = &w;
    for (int l = 0; l < kLanes; l++) {
        w.done[l] = false;
        w.where[l] = "start";
        getcontext(&w.lane[l]);
        w.lane[l].uc_stack.ss_sp = w.stacks + kStack * l;
        w.lane[l].uc_stack.ss_size = kStack;
        w.lane[l].uc_link = &w.sched;
        makecontext(&w.lane[l], (void (*)())trampoline, 0);
    }
    for (;;) {
        int alive = 0;
        const unsigned long p0 = w.progress;
        for (int l = 0; l < kLanes; l++) {
            if (w.done[l]) continue;
            alive++;
            w.cur = l;
            swapcontext(&w.sched, &w.lane[l]);
        }
        if (!alive) break;
        if (w.progress == p0) {
            // a whole pass in which no lane arrived anywhere new or finished: the lanes wait for
            // peers that exited or that sit in a different collective
            fprintf(stderr, "simt_emu: divergent collective / deadlock:\n");
            for (int l = 0; l < kLanes; l++) fprintf(stderr, "  lane %2d: %s\n", l, w.where[l]);
            abort();
        }
    }
    free(w.stacks);
    active() = prev;
}

template <class T>
inline T exchange(T v, unsigned src, const char *what) {
    static_assert(sizeof(T) <= 8, "exchange of <= 8 bytes");
    Warp *w = active();
    const unsigned g = w->gen;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w->slot[g & 1][w->cur] = raw;
    barrier(what);
    T r;
    memcpy(&r, &w->slot[g & 1][src & 31], sizeof(T));
    return r;
}

// Gathers one value per lane (all lanes see all values).
inline void gather(uint64_t v, uint64_t out[kLanes], const char *what) {
    Warp *w = active();
    const unsigned g = w->gen;
    w->slot[g & 1][w->cur] = v;
    barrier(what);
    memcpy(out, w->slot[g & 1], sizeof(uint64_t) * kLanes);
}

}  // namespace simt

// ---- the CUDA intrinsics the kernels use (full-mask only) -------------------------------------
#define SIMT_FULLMASK(m)                                                      \
    do {                                                                      \
        if ((m) != 0xffffffffu) {                                             \
            fprintf(stderr, "simt_emu: partial-mask collective at %s:%d\n", __FILE__, __LINE__); \
            abort();                                                          \
        }                                                                     \
    } while (0)

template <class T>
static inline T __shfl_sync(unsigned m, T v, unsigned src) {
    SIMT_FULLMASK(m);
    return simt::exchange<T>(v, src, "shfl");
}
template <class T>
static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    SIMT_FULLMASK(m);
    const unsigned l = simt::lane();
    return simt::exchange<T>(v, l >= d ? l - d : l, "shfl_up");
}
template <class T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
    SIMT_FULLMASK(m);
    const unsigned l = simt::lane();
    return simt::exchange<T>(v, l + d < 32 ? l + d : l, "shfl_down");
}
template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, unsigned x) {
    SIMT_FULLMASK(m);
    return simt::exchange<T>(v, (unsigned)simt::lane() ^ x, "shfl_xor");
}
static inline unsigned __ballot_sync(unsigned m, bool p) {
    SIMT_FULLMASK(m);
    uint64_t all[32];
    simt::gather(p ? 1 : 0, all, "ballot");
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)(all[l] & 1) << l;
    return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline void __syncwarp(unsigned m = 0xffffffffu) {
    SIMT_FULLMASK(m);
    simt::barrier("syncwarp");
}
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) {
    SIMT_FULLMASK(m);
    uint64_t all[32];
    simt::gather(v, all, "reduce_or");
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)all[l];
    return r;
}
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) {
    SIMT_FULLMASK(m);
    uint64_t all[32];
    simt::gather(v, all, "reduce_max");
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r = std::max(r, (unsigned)all[l]);
    return r;
}
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) {
    SIMT_FULLMASK(m);
    uint64_t all[32];
    simt::gather(v, all, "reduce_min");
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; l++) r = std::min(r, (unsigned)all[l]);
    return r;
}

static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31));
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> std::min(sh, 32u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (sh & 31)) >> 32);
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 7;
        r |= (uint32_t)((v >> (8 * sel)) & 0xff) << (8 * i);
    }
    return r;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
using std::max;
using std::min;

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}
static inline uint32_t __ldg(const uint32_t *p) { return *p; }
template <class T>
static inline T __ldcg(const T *p) { return *p; }
template <class T>
static inline void __stcg(T *p, T v) { *p = v; }
static inline unsigned __match_any_sync(unsigned m, uint32_t v) {
    SIMT_FULLMASK(m);
    uint64_t all[32];
    simt::gather(v, all, "match_any");
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)((uint32_t)all[l] == v) << l;
    return r;
}
