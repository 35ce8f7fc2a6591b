// trc_common.cuh -- geometry, per-unit metadata and small device helpers shared by every kernel.
//
// Vocabulary:
//   call  = one reference function call (one chunk of the batch), index j
//   unit  = the piece of a call one GPU thread codes independently: the whole call for the range-coder
//           and static-rANS codecs, one ANSBLKSIZE (4 MiB) block of the call for the adaptive rANS codecs
//           (anscdf.c:54,573-583 re-initialise model and states per block), index u = j*upc + b
//   slot  = per-unit scratch area the coder writes into before the pack kernel lays the bytes out
//           exactly as the reference does
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

typedef unsigned short cdf_t;                     // include/turborc.h:497

namespace trc {

constexpr int      PROB_BITS  = 15;               // ANS_BITS anscdf_.h:33 / RC_BITS rccdf.c:37
constexpr uint32_t PROB_TOTAL = 1u << PROB_BITS;
constexpr uint32_t PROB_MASK  = PROB_TOTAL - 1;
constexpr uint32_t ANS_L      = 1u << 15;          // ANS_LOW anscdf_.h:40-41
constexpr uint32_t ANS_BLOCK  = 1u << 22;          // ANSBLKSIZE anscdf.c:54
constexpr int      CDF_STRIDE = 257;

enum Codec { ANS4S = 0, ANS4, ANS, ANS1, RCS, RCS2, RC, RCI, RC4, RC4I, ANSW, RC8, RCI8,
             ANSU16, ANSUZ16, ANSV16, ANSVZ16, ANSV32, ANSVZ32, RCV16, RCVZ16, RCV32, RCVZ32, RCU16, RCU32,   // VLC-over-CDF integer codecs (vlc.cuh)
             NCODECS };

__host__ __device__ inline bool codec_blocked(int c) { return c == ANS4 || c == ANS || c == ANS1; }
__host__ __device__ inline bool codec_static(int c)  { return c == ANS4S || c == RCS || c == RCS2 || c == ANSW; }

struct Geom {
    size_t   total;      // bytes of uncompressed data
    size_t   chunk;      // bytes per call (last call may be shorter)
    size_t   n_calls;
    uint32_t upc;        // units per call (1 unless blocked codec and chunk > ANS_BLOCK)
    uint32_t unit_max;   // max bytes per unit
    size_t   n_units;
};

inline Geom make_geom(int codec, size_t total, size_t chunk) {
    Geom g;
    g.total = total; g.chunk = chunk;
    g.n_calls = chunk ? (total + chunk - 1) / chunk : 0;
    if (codec_blocked(codec) && chunk > ANS_BLOCK) {
        g.upc = (uint32_t)((chunk + ANS_BLOCK - 1) / ANS_BLOCK); g.unit_max = ANS_BLOCK;
    } else {
        g.upc = 1; g.unit_max = (uint32_t)(chunk < total ? chunk : total);
    }
    g.n_units = g.n_calls * g.upc;
    return g;
}

__host__ __device__ inline void call_span(const Geom &g, size_t j, size_t &start, size_t &len) {
    start = j * g.chunk;
    size_t rem = g.total - start;
    len = rem < g.chunk ? rem : g.chunk;
}
// unit u -> (call j, block b, start, len); len == 0 for the padding units of a short last call
__host__ __device__ inline void unit_span(const Geom &g, size_t u, size_t &j, uint32_t &b, size_t &start, size_t &len) {
    j = u / g.upc; b = (uint32_t)(u % g.upc);
    size_t cs, cl; call_span(g, j, cs, cl);
    size_t o = (size_t)b * g.unit_max;
    if (o >= cl) { start = cs + cl; len = 0; return; }
    start = cs + o;
    len = cl - o < g.unit_max ? cl - o : g.unit_max;
}

// What the coding kernel leaves behind for the pack kernel.
struct UnitMeta {
    uint32_t len;        // compressed bytes of the unit (a_len + b_len), or the input length when raw
    uint32_t a_off, a_len, b_off, b_len;   // byte ranges inside the slot, concatenated a then b
    uint32_t flags;      // UM_*
    uint32_t pref;       // filled by the scan kernel: byte offset of this unit inside its call's stream
    uint32_t pad;
};
enum : uint32_t {
    UM_RAW  = 1u,        // the reference call returns a raw copy (decided inside the coding kernel)
    UM_ADJ2 = 2u,        // blocked rANS: the last-coded record did not emit (mnflush guard slack)
    UM_OVF  = 4u,        // slot exhausted: the unit certainly overflows its call
    UM_QUIRK4 = 8u       // rccdf4ienc in-loop overflow: raw bytes in `out` but returns op0-out (rccdf.c:314,322)
};

// ---- unaligned-safe little-endian accessors ---------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) {
    if (((uintptr_t)p & 1) == 0) return *(const uint16_t *)p;
    return (uint32_t)p[0] | (uint32_t)p[1] << 8;
}
__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    if (((uintptr_t)p & 3) == 0) return *(const uint32_t *)p;
    if (((uintptr_t)p & 1) == 0) return (uint32_t)*(const uint16_t *)p | (uint32_t)*(const uint16_t *)(p + 2) << 16;
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
// reads past `end` return zero bytes (decoders legitimately look 2-8 bytes ahead of what they use)
__device__ __forceinline__ uint32_t ld_u16_clamped(const uint8_t *p, const uint8_t *end) {
    if (p + 2 <= end) return ld_u16(p);
    return p < end ? (uint32_t)p[0] : 0u;
}
__device__ __forceinline__ uint32_t ld_u32_clamped(const uint8_t *p, const uint8_t *end) {
    if (p + 4 <= end) return ld_u32(p);
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) if (p + i < end) v |= (uint32_t)p[i] << (8 * i);
    return v;
}
__device__ __forceinline__ void st_u16(uint8_t *p, uint32_t v) {      // p is 2-byte aligned
    *(uint16_t *)p = (uint16_t)v;
}
__device__ __forceinline__ void st_u32_a2(uint8_t *p, uint32_t v) {   // p is 2-byte aligned
    if (((uintptr_t)p & 3) == 0) { *(uint32_t *)p = v; return; }
    *(uint16_t *)p = (uint16_t)v; *(uint16_t *)(p + 2) = (uint16_t)(v >> 16);
}

// serial byte copy by one thread (raw chunks inside decode kernels; rare path)
__device__ inline void thread_copy(uint8_t *dst, const uint8_t *src, size_t n) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        size_t v = n >> 4;
        for (size_t i = 0; i < v; i++) ((uint4 *)dst)[i] = ((const uint4 *)src)[i];
        dst += v << 4; src += v << 4; n &= 15;
    }
    for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

// cooperative byte copy with arbitrary alignment of both sides: dst-aligned 32-bit stores, source words
// funnel-shifted into place.  `tid`/`nthr` = the cooperating threads.  Reads at most the aligned word that
// holds the last source byte.
__device__ inline void group_copy(uint8_t *dst, const uint8_t *src, size_t n, unsigned tid, unsigned nthr) {
    if (((((uintptr_t)dst | (uintptr_t)src) & 3) == 0) && n >= 64) {
        // both sides word aligned (every range-coder piece): 128-bit stores on the destination's 16-byte grid, source
        // words taken from two aligned 128-bit loads and rotated by the (uniform) word misalignment
        size_t headw = ((16 - ((uintptr_t)dst & 15)) & 15) >> 2;                // words until dst is 16-byte aligned
        if (tid < headw) ((uint32_t *)dst)[tid] = ((const uint32_t *)src)[tid];
        const uint32_t *s = (const uint32_t *)src + headw;
        uint4 *d16 = (uint4 *)((uint32_t *)dst + headw);
        const size_t nw = (n >> 2) - headw, nv = nw >> 2;
        const unsigned mis = (unsigned)(((uintptr_t)s & 15) >> 2);
        const uint4 *s16 = (const uint4 *)((uintptr_t)s & ~(uintptr_t)15);
        for (size_t i = tid; i < nv; i += nthr) {
            uint4 lo = s16[i], v;
            if (mis == 0) v = lo;
            else {
                uint4 hi = s16[i + 1];
                if (mis == 1) v = make_uint4(lo.y, lo.z, lo.w, hi.x);
                else if (mis == 2) v = make_uint4(lo.z, lo.w, hi.x, hi.y);
                else v = make_uint4(lo.w, hi.x, hi.y, hi.z);
            }
            d16[i] = v;
        }
        const size_t donew = headw + (nv << 2), totw = n >> 2;
        if (donew + tid < totw) ((uint32_t *)dst)[donew + tid] = ((const uint32_t *)src)[donew + tid];   // < 4 tail words
        const size_t doneb = totw << 2;
        if (tid < n - doneb) dst[doneb + tid] = src[doneb + tid];
        return;
    }
    size_t head = (4 - ((uintptr_t)dst & 3)) & 3;
    if (head > n) head = n;
    if (tid < head) dst[tid] = src[tid];
    dst += head; src += head; n -= head;
    size_t nw = n >> 2;
    unsigned sh = ((uintptr_t)src & 3) * 8;
    const uint32_t *s4 = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
    uint32_t *d4 = (uint32_t *)dst;
    if (sh == 0) { for (size_t i = tid; i < nw; i += nthr) d4[i] = s4[i]; }
    else         { for (size_t i = tid; i < nw; i += nthr) d4[i] = __funnelshift_r(s4[i], s4[i + 1], sh); }
    size_t done = nw << 2;
    if (tid < n - done) dst[done + tid] = src[done + tid];
}

// group_copy for word-aligned pieces with FOUR 16-byte chunks in flight per thread: used where only a few warps copy (the
// layout epilogue of the fused encoder), so memory-level parallelism has to come from each thread
__device__ inline void group_copy4(uint8_t *dst, const uint8_t *src, size_t n, unsigned tid, unsigned nthr) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 3) || n < 256) { group_copy(dst, src, n, tid, nthr); return; }
    const size_t headw = ((16 - ((uintptr_t)dst & 15)) & 15) >> 2;                  // words until dst is 16-byte aligned
    if (tid < headw) ((uint32_t *)dst)[tid] = ((const uint32_t *)src)[tid];
    const uint32_t *s = (const uint32_t *)src + headw;
    uint4 *d16 = (uint4 *)((uint32_t *)dst + headw);
    const size_t nw = (n >> 2) - headw, nv = nw >> 2;
    const unsigned mis = (unsigned)(((uintptr_t)s & 15) >> 2);
    const uint4 *s16 = (const uint4 *)((uintptr_t)s & ~(uintptr_t)15);
    auto rot = [&](const uint4 &lo, const uint4 &hi) -> uint4 {
        if (mis == 0) return lo;
        if (mis == 1) return make_uint4(lo.y, lo.z, lo.w, hi.x);
        if (mis == 2) return make_uint4(lo.z, lo.w, hi.x, hi.y);
        return make_uint4(lo.w, hi.x, hi.y, hi.z);
    };
    size_t i = tid;
    for (; i + 3 * (size_t)nthr < nv; i += 4 * (size_t)nthr) {
        uint4 lo[4], hi[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { lo[k] = s16[i + k * nthr]; hi[k] = mis ? s16[i + k * nthr + 1] : lo[k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) d16[i + k * nthr] = rot(lo[k], hi[k]);
    }
    for (; i < nv; i += nthr) { const uint4 lo = s16[i], hi = mis ? s16[i + 1] : lo; d16[i] = rot(lo, hi); }
    const size_t donew = headw + (nv << 2), totw = n >> 2;
    if (donew + tid < totw) ((uint32_t *)dst)[donew + tid] = ((const uint32_t *)src)[donew + tid];   // < 4 tail words
    const size_t doneb = totw << 2;
    if (tid < n - doneb) dst[doneb + tid] = src[doneb + tid];
}

}  // namespace trc
