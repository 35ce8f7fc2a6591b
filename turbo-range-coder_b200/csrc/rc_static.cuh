// rc_static.cuh -- 64-bit range coder with 32-bit output words, carry propagation, 15-bit CDFs
// (canonical RC_SIZE 64 / RC_IO 32 / RC_BITS 15 format, turborc_.h:41-82 with rccdf.c:36-37), and the
// static-CDF codecs built on it: TRC_RCS (rccdfsenc / rccdfsbdec, rccdf.c:71-98) and TRC_RCS2
// (rccdfs2enc / rccdfsb2dec, rccdf.c:125-184).  One GPU thread owns one reference call (1 or 2 coders).
#pragma once
#include "trc_common.cuh"

namespace trc {

// ---- encoder (rceinit/_rccdfenc_/_rcenorm_/_rccarry_/rceflush, turborc_.h:103-128,215) -----------------
// The reference stores each word at renormalisation and, when a later addition overflows `low`, walks
// back incrementing stored words.  Here the newest word stays pending in a register until the next one
// (and its carry decision) is known; memory is touched again only if the pending word wraps to zero.
struct RcEnc {
    uint64_t low, range;
    uint8_t *base;       // 4-byte aligned
    uint32_t pos;        // bytes put so far (== reference op - stream start)
    uint32_t pend;
    bool     carry;      // low wrapped since the last renormalisation (== reference "ilow > low")

    __device__ __forceinline__ void init(uint8_t *b) { low = 0; range = ~0ull; base = b; pos = 0; pend = 0; carry = false; }
    __device__ __noinline__ void walk_back() {           // pending word wrapped: propagate into stored words
        uint8_t *p = base + pos - 4;
        for (;;) { p -= 4; uint32_t w = *(uint32_t *)p + 1; *(uint32_t *)p = w; if (w) break; }
    }
    __device__ __forceinline__ void put(uint32_t w) {
        if (carry) { carry = false; if (++pend == 0 && pos >= 8) walk_back(); }
        if (pos) *(uint32_t *)(base + pos - 4) = pend;
        pend = w; pos += 4;
    }
    __device__ __forceinline__ void norm() {
        if ((uint32_t)(range >> 32) == 0) { put((uint32_t)(low >> 32)); low <<= 32; range <<= 32; }
    }
    __device__ __forceinline__ void add_low(uint64_t a) { uint64_t nl = low + a; carry |= nl < low; low = nl; }
    __device__ __forceinline__ void encode(uint32_t c0, uint32_t f) {
        range >>= PROB_BITS; add_low(range * c0); range *= f; norm();
    }
    __device__ inline void flush() {
        norm();
        if (range > (1ull << 33)) { add_low(1ull << 32); put((uint32_t)(low >> 32)); }
        else { add_low(1ull); put((uint32_t)(low >> 32)); put((uint32_t)low); }
        *(uint32_t *)(base + pos - 4) = pend;            // pos >= 4 here
    }
};

// OVERFLOW threshold rcutil_.h:130: op >= out + inlen*255/256 - 8, as a signed offset
__device__ __forceinline__ int64_t rc_thr(size_t n) { return (int64_t)((n * 255) / 256) - 8; }

constexpr int RC_S_NT = 128;

// cdf | freq << 16
__device__ __forceinline__ void rc_build_ctab(uint32_t *ctab, const cdf_t *c0, unsigned cdfnum, int nt) {
    for (unsigned x = threadIdx.x; x < 256; x += nt) {
        uint32_t e = 0;
        if (x < cdfnum) { uint32_t c = c0[x]; e = c | ((uint32_t)c0[x + 1] - c) << 16; }
        ctab[x] = e;
    }
}

template <int NC, bool FAST>
__device__ inline void rc_static_enc_call(const uint8_t *ip, size_t n, const uint32_t *ctab, const cdf_t *gcdf,
                                          uint8_t *slot, UnitMeta &m) {
#define TRC_CE(_x_) (FAST ? ctab[_x_] : ((uint32_t)gcdf[_x_] | ((uint32_t)gcdf[(_x_) + 1] - gcdf[_x_]) << 16))
    const int64_t thr = rc_thr(n);
    m.pref = 0; m.pad = 0; m.b_off = 0; m.b_len = 0; m.a_off = 0;
    bool raw = false;
    if (NC == 1) {
        RcEnc e; e.init(slot);
        for (size_t i = 0; i < n; i++) {
            uint32_t t = TRC_CE(ip[i]);
            e.encode(t & 0xffffu, t >> 16);
            if ((int64_t)e.pos >= thr) { raw = true; break; }                        // OVERFLOW rccdf.c:77
        }
        if (!raw) e.flush();
        m.a_len = raw ? 0 : e.pos; m.len = raw ? (uint32_t)n : e.pos; m.flags = raw ? UM_RAW : 0;
    } else {
        if (n < 4) { m.a_len = 0; m.len = (uint32_t)n; m.flags = UM_RAW; return; }   // reference is undefined here
        const uint32_t b1ref = 4 + (uint32_t)(((n - 4) * 37) / 64);                 // rccdf.c:126
        const uint32_t b1 = (b1ref + 64 + 15) & ~15u;                                // where stream 1 really lives in the slot
        RcEnc e0, e1; e0.init(slot + 4); e1.init(slot + b1);
        size_t i = 0, n2 = n & ~(size_t)1;
        for (; i < n2; i += 2) {
            uint32_t t0 = TRC_CE(ip[i]), t1 = TRC_CE(ip[i + 1]);
            e0.encode(t0 & 0xffffu, t0 >> 16);
            e1.encode(t1 & 0xffffu, t1 >> 16);
            if ((int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref) { raw = true; break; }   // OVERFLOWI rccdf.c:46,133
        }
        if (!raw) {
            if (i < n) { uint32_t t = TRC_CE(ip[i]); e0.encode(t & 0xffffu, t >> 16); }
            e0.flush(); e1.flush();
            *(uint32_t *)slot = e0.pos;                                              // rccdf.c:141
            if ((int64_t)(4 + e0.pos + e1.pos) >= thr) raw = true;                   // final OVERFLOW rccdf.c:142
        }
        m.a_len = raw ? 0 : 4 + e0.pos; m.b_off = b1; m.b_len = raw ? 0 : e1.pos;
        m.len = raw ? (uint32_t)n : 4 + e0.pos + e1.pos; m.flags = raw ? UM_RAW : 0;
    }
#undef TRC_CE
}

template <int NC>
__global__ void __launch_bounds__(RC_S_NT)
k_rc_static_enc(const uint8_t *__restrict__ in, Geom g, const cdf_t *__restrict__ cdf, unsigned cdfnum, size_t cpc,
                uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ uint32_t ctab[256];
    size_t j0 = (size_t)blockIdx.x * RC_S_NT, j = j0 + threadIdx.x;
    size_t t0 = cpc ? j0 / cpc : 0;
    rc_build_ctab(ctab, cdf + t0 * CDF_STRIDE, cdfnum, RC_S_NT);
    __syncthreads();
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    size_t t = cpc ? j / cpc : 0;
    UnitMeta m;
    if (t == t0) rc_static_enc_call<NC, true >(in + start, len, ctab, nullptr, slots + j * slot_stride, m);
    else         rc_static_enc_call<NC, false>(in + start, len, nullptr, cdf + t * CDF_STRIDE, slots + j * slot_stride, m);
    meta[j] = m;
}

// ---- decoder (rcdinit/_rccdfrange/_cdfbget/_rccdfupdate, turborc_.h:152-158,224,307-315,219-229) --------
struct RcDec {
    uint64_t range, code;
    const uint8_t *ip, *end;
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *e) {
        end = e; range = ~0ull;
        code = (uint64_t)ld_u32_clamped(p, e) << 32 | ld_u32_clamped(p + 4, e);
        ip = p + 8;
    }
    __device__ __forceinline__ void shift() { range >>= PROB_BITS; }
    template <class C> __device__ __forceinline__ unsigned bsearch(const C *cdf, unsigned cdfnum) const {
        unsigned x = 0, hi = cdfnum;
        while (x + 1 < hi) { unsigned mid = (x + hi) >> 1; if ((uint64_t)cdf[mid] * range > code) hi = mid; else x = mid; }
        return x;
    }
    __device__ __forceinline__ void update(uint32_t c0, uint32_t c1) {
        uint64_t rp = (uint64_t)c0 * range;
        range = range * c1 - rp; code -= rp;
        if ((uint32_t)(range >> 32) == 0) { range <<= 32; code = code << 32 | ld_u32_clamped(ip, end); ip += 4; }
    }
};

constexpr int RC_SD_NT = 128;

template <int NC, class C>
__device__ inline void rc_static_dec_call(const uint8_t *stream, const uint8_t *gend, uint8_t *op, size_t n,
                                          const C *cdf, unsigned cdfnum) {
    if (NC == 1) {
        RcDec d; d.init(stream, gend);
        for (size_t i = 0; i < n; i++) {
            d.shift(); unsigned x = d.bsearch(cdf, cdfnum); d.update(cdf[x], cdf[x + 1]); op[i] = (uint8_t)x;
        }
    } else {
        RcDec d0, d1;
        uint32_t len0 = ld_u32_clamped(stream, gend);
        const uint8_t *p1 = stream + 4 + len0;
        if (p1 > gend || p1 < stream) p1 = gend;                                     // garbage header: stay in bounds
        d0.init(stream + 4, gend); d1.init(p1, gend);
        size_t i = 0, n2 = n & ~(size_t)1;
        for (; i < n2; i += 2) {
            d0.shift(); d1.shift();
            unsigned x0 = d0.bsearch(cdf, cdfnum), x1 = d1.bsearch(cdf, cdfnum);
            d0.update(cdf[x0], cdf[x0 + 1]); d1.update(cdf[x1], cdf[x1 + 1]);
            if ((((uintptr_t)(op + i)) & 1) == 0) *(uint16_t *)(op + i) = (uint16_t)(x0 | x1 << 8);
            else { op[i] = (uint8_t)x0; op[i + 1] = (uint8_t)x1; }
        }
        if (i < n) { d0.shift(); unsigned x = d0.bsearch(cdf, cdfnum); d0.update(cdf[x], cdf[x + 1]); op[i] = (uint8_t)x; }
    }
}

template <int NC>
__global__ void __launch_bounds__(RC_SD_NT)
k_rc_static_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                const cdf_t *__restrict__ cdf, unsigned cdfnum, size_t cpc) {
    __shared__ uint16_t scdf[CDF_STRIDE];
    size_t j0 = (size_t)blockIdx.x * RC_SD_NT, j = j0 + threadIdx.x;
    size_t t0 = cpc ? j0 / cpc : 0;
    for (unsigned x = threadIdx.x; x <= cdfnum; x += RC_SD_NT) scdf[x] = cdf[t0 * CDF_STRIDE + x];
    __syncthreads();
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls];
    if (sl == len) { thread_copy(out + start, in + so, len); return; }
    size_t t = cpc ? j / cpc : 0;
    if (t == t0) rc_static_dec_call<NC>(in + so, gend, out + start, len, (const uint16_t *)scdf, cdfnum);
    else         rc_static_dec_call<NC>(in + so, gend, out + start, len, cdf + t * CDF_STRIDE, cdfnum);
}

}  // namespace trc
