// vnibble.cuh -- TRC_RC8 / TRC_RCI8: rccdfenc8 / rccdfdec8 and rccdfienc8 / rccdfidec8 (rccdf.c:324-389), the adaptive range
// coders over the "vnibble" byte code of cdfe8 / cdfd8 (rccdf_.h:76-96):
//     x < 13        one symbol  x                        on table 0
//     x < 13 + 32   (x-13 >> 4) + 13 on table 0, then (x-13) & 15 on table 1
//     else          15 on table 0, (x-45) >> 4 on table 1, (x-45) & 15 on table 2
// The interleaved form sends the table-1 symbols to coder 1 (stream 1 scratch at out + 4 + inlen*37/64) and everything
// else to coder 0.  One lane per call (first form: throughput for batches of many calls), built
// from the pieces of adaptive.cuh: packed 16-entry tables in shared memory (word-major / thread-minor), cdf16upd as 8
// SIMD-within-register steps, RcEnc / RcDec.
#pragma once
#include "trc_common.cuh"
#include "adaptive.cuh"

namespace trc {

constexpr int V8_NT = 128;                                       // 3 tables x 32 B x 128 lanes = 12 KB of shared memory

template <class Tab>
__device__ __forceinline__ void v8_put(RcEnc &e, Tab t, unsigned x) { uint32_t c, f; tab_enc(t, x, c, f); e.encode(c, f); }   // cdf4e rccdf_.h:28

template <class Tab>
__device__ __forceinline__ void v8_enc(RcEnc &e0, RcEnc &e1, Tab m0, Tab m1, Tab m2, unsigned x) {                             // cdfe8 rccdf_.h:80-87
    if (x < 13) v8_put(e0, m0, x);
    else if (x < 13 + 32) { x -= 13; v8_put(e0, m0, (x >> 4) + 13); v8_put(e1, m1, x & 15); }
    else { x -= 13 + 32; v8_put(e0, m0, 15); v8_put(e1, m1, x >> 4); v8_put(e0, m2, x & 15); }
}
template <class Tab>
__device__ __forceinline__ unsigned v8_dec(RcDec &d0, RcDec &d1, Tab m0, Tab m1, Tab m2) {                                     // cdfd8 rccdf_.h:89-96
    unsigned x = rc_dec_nib(m0, d0);
    if (x >= 13) {
        const unsigned y = rc_dec_nib(m1, d1);
        if (x != 15) x = ((x - 13) << 4 | y) + 13;
        else { x = rc_dec_nib(m2, d0); x = (y << 4 | x) + 13 + 32; }
    }
    return x;
}

template <int NC>
__global__ void __launch_bounds__(V8_NT)
k_rc_v8_enc(const uint8_t *__restrict__ in, Geom g, uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ uint32_t sm[3 * 8 * V8_NT];
    const size_t j = (size_t)blockIdx.x * V8_NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + j * slot_stride;
    SmTab<V8_NT> m0{sm + threadIdx.x}, m1{sm + 8 * V8_NT + threadIdx.x}, m2{sm + 16 * V8_NT + threadIdx.x};
    tab_init(m0); tab_init(m1); tab_init(m2);                    // CDF16DEC0 x3 (rccdf.c:342)
    const int64_t thr = rc_thr(n);
    UnitMeta m; m.pref = 0; m.pad = 0; m.a_off = 0; m.b_off = 0; m.b_len = 0; m.flags = 0;
    bool raw = false;
    if (NC == 1) {
        RcEnc e; e.init(slot);
        for (size_t i = 0; i < n; i++) {
            v8_enc(e, e, m0, m1, m2, ip[i]);
            if ((int64_t)e.pos >= thr) { raw = true; break; }    // OVERFLOW rccdf.c:347
        }
        if (!raw) e.flush();
        m.a_len = raw ? 0 : e.pos; m.len = raw ? (uint32_t)n : e.pos;
    } else {
        const uint32_t b1ref = 4 + (uint32_t)((n * 37) / 64);    // rccdf.c:373
        const uint32_t b1 = (b1ref + 64 + 15) & ~15u;            // our stream-1 scratch: aligned, past anything stream 0 can reach
        RcEnc e0, e1; e0.init(slot + 4); e1.init(slot + b1);
        size_t i = 0;
        const size_t n4 = n & ~(size_t)3;
        for (; i < n4 && !raw; i += 4) {                         // rccdf.c:377-382
            for (int k = 0; k < 4; k++) v8_enc(e0, e1, m0, m1, m2, ip[i + k]);
            if ((int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref) raw = true;   // OVERFLOWI rccdf.c:46
        }
        if (!raw) {
            for (; i < n; i++) v8_enc(e0, e1, m0, m1, m2, ip[i]);
            e0.flush(); e1.flush();
            *(uint32_t *)slot = e0.pos;                          // rccdf.c:387
            if ((int64_t)(4 + e0.pos + e1.pos) >= thr) raw = true;
        }
        m.a_len = raw ? 0 : 4 + e0.pos; m.b_off = b1; m.b_len = raw ? 0 : e1.pos;
        m.len = raw ? (uint32_t)n : 4 + e0.pos + e1.pos;
    }
    m.flags = raw ? UM_RAW : 0;
    meta[j] = m;
}

template <int NC>
__global__ void __launch_bounds__(V8_NT)
k_rc_v8_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    __shared__ uint32_t sm[3 * 8 * V8_NT];
    const size_t j = (size_t)blockIdx.x * V8_NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    const uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls], *stream = in + so;
    uint8_t *op = out + start;
    if (sl == n) { thread_copy(op, stream, n); return; }         // raw chunk (CCPY turborc.c:434)
    SmTab<V8_NT> m0{sm + threadIdx.x}, m1{sm + 8 * V8_NT + threadIdx.x}, m2{sm + 16 * V8_NT + threadIdx.x};
    tab_init(m0); tab_init(m1); tab_init(m2);
    if (NC == 1) {
        RcDec d; d.init(stream, gend);
        for (size_t i = 0; i < n; i++) op[i] = (uint8_t)v8_dec(d, d, m0, m1, m2);
    } else {
        const uint32_t len0 = ld_u32_clamped(stream, gend);
        const uint8_t *p1 = stream + 4 + len0;                   // rccdf.c:356
        if (p1 > gend || p1 < stream) p1 = gend;
        RcDec d0, d1; d0.init(stream + 4, gend); d1.init(p1, gend);
        for (size_t i = 0; i < n; i++) op[i] = (uint8_t)v8_dec(d0, d1, m0, m1, m2);
    }
}

}  // namespace trc
