// adaptive.cuh -- adaptive 16-symbol CDF model (cdf_.h:25-107) and the codecs built on it:
//   rANS : TRC_ANS4 (anscdf.c:87-133), TRC_ANS (anscdf.c:567-605), TRC_ANS1 order-1 (anscdf.c:607-645)
//   RC   : TRC_RC (rccdf.c:187-211), TRC_RCI (rccdf.c:213-249), TRC_RC4 / TRC_RC4I (rccdf.c:251-323)
//
// Model storage: a 16-entry table is 8 packed 32-bit words (entry 2w in the low half, 2w+1 in the high
// half; entry 16 == 32768 is implicit).  The update of all 16 entries is done as 8 SIMD-within-register
// steps that reproduce the reference's 16-bit lane arithmetic exactly (see adapt_word below).
// Order-0 tables live in shared memory, word-major / thread-minor so that every access of a warp is
// bank-conflict free whatever table each thread selects; order-1 tables (136 KB per coder) live in global
// memory (L2-resident).
#pragma once
#include "trc_common.cuh"
#include "rans_static.cuh"
#include "rc_static.cuh"

namespace trc {

constexpr uint32_t AD_MIX = 32736;   // MIXD cdf_.h:36
// CDFRATE 7, IC 10 (cdf_.h:25,35).  For entry i with value m and "greater" flag g the reference computes, in
// signed 16-bit lanes,  m += (10*i + (g ? 32736 : 0) - m) >> 7.  Table invariants (10*i <= m <= 32759,
// strictly increasing) keep d = 10*i + g*32736 - m inside [-32759, 32736], so with a +32768 bias the
// arithmetic shift becomes a logical one:  (d >> 7) == ((d + 32768) >> 7) - 256, and both halves of a
// packed word can be processed by ordinary 32-bit adds without carries crossing the halves.
__device__ __forceinline__ uint32_t adapt_bias_const(int w) {          // (10*i + 0x8000) for i = 2w, 2w+1
    return (uint32_t)(20 * w + 0x8000) | (uint32_t)(20 * w + 10 + 0x8000) << 16;
}
// le = per-half flag word: bit 15 / bit 31 set where entry <= cmp.  Returns the updated packed word.
__device__ __forceinline__ uint32_t adapt_word(uint32_t m, uint32_t le, int w) {
    uint32_t g = (~le >> 15) & 0x00010001u;                             // entries > cmp get MIXD
    uint32_t d = adapt_bias_const(w) - m + g * AD_MIX;
    uint32_t s = (d >> 7) & 0x01FF01FFu;
    return m + s - 0x01000100u;
}
__device__ __forceinline__ uint32_t adapt_le(uint32_t m, uint32_t cmp2) {   // cmp2 = cmp | cmp << 16 | 0x80008000
    return cmp2 - m;                                                    // bit15/31 set  <=>  entry <= cmp
}
__device__ __forceinline__ uint32_t adapt_init_word(int w) { return (uint32_t)(2 * w) << 11 | (uint32_t)(2 * w + 1) << 27; }

// ---- table views --------------------------------------------------------------------------------------
template <int NT> struct SmTab {       // shared memory, words interleaved across the NT threads of the CTA
    uint32_t *p;                       // &words[table*8*NT + threadIdx.x]
    __device__ __forceinline__ void load(uint32_t m[8]) const {
#pragma unroll
        for (int w = 0; w < 8; w++) m[w] = p[w * NT];
    }
    __device__ __forceinline__ void store(const uint32_t m[8]) {
#pragma unroll
        for (int w = 0; w < 8; w++) p[w * NT] = m[w];
    }
    __device__ __forceinline__ uint32_t entry(unsigned e) const {       // e in 0..16
        if (e >= 16) return PROB_TOTAL;
        return ((const uint16_t *)(p + (e >> 1) * NT))[e & 1];
    }
};
struct GmTab {                         // global memory, 32 contiguous bytes per table
    uint32_t *p;
    __device__ __forceinline__ void load(uint32_t m[8]) const {
        uint4 a = ((const uint4 *)p)[0], b = ((const uint4 *)p)[1];
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    }
    __device__ __forceinline__ void store(const uint32_t m[8]) {
        ((uint4 *)p)[0] = make_uint4(m[0], m[1], m[2], m[3]); ((uint4 *)p)[1] = make_uint4(m[4], m[5], m[6], m[7]);
    }
    __device__ __forceinline__ uint32_t entry(unsigned e) const {
        if (e >= 16) return PROB_TOTAL;
        return ((const uint16_t *)p)[e];
    }
};

template <class Tab> __device__ __forceinline__ void tab_init(Tab t) {
    uint32_t m[8];
#pragma unroll
    for (int w = 0; w < 8; w++) m[w] = adapt_init_word(w);
    t.store(m);
}
// encoder side: (cum, freq) of symbol x, then cdf16upd (cdf_.h:46-50): "greater" == entry > entry[x]
template <class Tab> __device__ __forceinline__ void tab_enc(Tab t, unsigned x, uint32_t &cum, uint32_t &freq) {
    cum = t.entry(x); freq = t.entry(x + 1) - cum;
    uint32_t m[8]; t.load(m);
    uint32_t c2 = cum * 0x00010001u | 0x80008000u;
#pragma unroll
    for (int w = 0; w < 8; w++) m[w] = adapt_word(m[w], adapt_le(m[w], c2), w);
    t.store(m);
}
// update only (RC4I applies both updates of a pair after both symbols were coded)
template <class Tab> __device__ __forceinline__ void tab_upd(Tab t, unsigned x) { uint32_t c, f; tab_enc(t, x, c, f); }
// rANS decoder side (cdf16ansdec cdf_.h:52-59): x = #entries <= r, minus one; update uses entry > r
template <class Tab> __device__ __forceinline__ unsigned tab_dec_ans(Tab t, uint32_t r, uint32_t &cum, uint32_t &freq) {
    uint32_t m[8]; t.load(m);
    uint32_t c2 = r * 0x00010001u | 0x80008000u, cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        uint32_t le = adapt_le(m[w], c2);
        cnt += __popc(le & 0x80008000u);
        m[w] = adapt_word(m[w], le, w);
    }
    unsigned x = cnt - 1;                                               // entry 0 == 0 <= r always
    cum = t.entry(x); freq = t.entry(x + 1) - cum;                      // memory still holds the old table
    t.store(m);
    return x;
}

// ======================================================================================================
// rANS adaptive encoders.  Thread per unit (a call, or one 4 MiB block of a call).
// Model pass pushes one record per nibble (mnenc4 anscdf_.h:106), coding pass pops them in reverse
// (mnflush anscdf_.h:128-138).  Records: freq | cum << 16 (the state index is implied by the position).
// ======================================================================================================
enum AMode { M_NIB = 0, M_BYTE = 1, M_O1 = 2 };

constexpr int AD_NT_BYTE = 64;        // 17 tables * 32 B * 64 threads = 34 KB shared
constexpr int AD_NT_NIB  = 128;
constexpr size_t O1_TAB_WORDS = 256 * 17 * 8;   // per coder: mbh[256][16] + mbl[256][16][16]  (anscdf.c:613-614)

__device__ __forceinline__ uint32_t rans_enc_step_rec(uint32_t s, uint32_t rec, RansWriter &w, bool &emitted) {
    uint32_t f = rec & 0xffffu, c = rec >> 16;
    emitted = s >= (f << 16);
    if (emitted) { w.put16(s); s >>= 16; }
    uint32_t q = s / f;
    return s + (q << PROB_BITS) - q * f + c;
}

// coding pass shared by the three modes.  nrec records at rec[0..nrec); NS states.
template <int NS>
__device__ inline void rans_adapt_flush(const uint32_t *rec, uint32_t nrec, uint32_t ntail, uint8_t *slot, int cap, uint32_t n, UnitMeta &m) {
    RansWriter w; w.init(slot, cap);
    uint32_t st[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) st[k] = ANS_L;
    bool em = false, ovf = false;
    uint32_t i = nrec;
    for (uint32_t k = 0; k < ntail; k++) { i--; st[0] = rans_enc_step_rec(st[0], rec[i], w, em); }   // NIB tail records (si = 0)
    while (i > 0) {
        i -= 4;
        uint4 r4 = *(const uint4 *)(rec + i);                          // groups are 16-byte aligned
        if (NS == 4) {                                                  // pushed 3,2,1,0 -> popped 0,1,2,3
            st[0] = rans_enc_step_rec(st[0], r4.w, w, em); st[1] = rans_enc_step_rec(st[1], r4.z, w, em);
            st[2 % NS] = rans_enc_step_rec(st[2 % NS], r4.y, w, em); st[3 % NS] = rans_enc_step_rec(st[3 % NS], r4.x, w, em);
        } else {                                                        // pushed 1,0,1,0 -> popped 0,1,0,1
            st[0] = rans_enc_step_rec(st[0], r4.w, w, em); st[1] = rans_enc_step_rec(st[1], r4.z, w, em);
            st[0] = rans_enc_step_rec(st[0], r4.y, w, em); st[1] = rans_enc_step_rec(st[1], r4.x, w, em);
        }
        if (w.pos < 32) { ovf = true; break; }                          // slot exhausted
    }
    w.finish_words();
#pragma unroll
    for (int k = 0; k < NS; k++) w.put32_final(st[k]);                  // ansflush
    uint32_t l = (uint32_t)(cap - w.pos);
    m.len = l; m.a_off = (uint32_t)w.pos; m.a_len = l; m.b_off = 0; m.b_len = 0;
    m.flags = (ovf ? UM_OVF : 0) | (em ? 0 : UM_ADJ2); m.pref = 0; m.pad = 0;
    (void)n;
}

template <int MODE, int NT>
__global__ void __launch_bounds__(NT)
k_rans_adapt_enc(const uint8_t *__restrict__ in, Geom g, uint8_t *__restrict__ slots, size_t slot_stride,
                 uint32_t *__restrict__ recs, size_t rec_stride, uint32_t *__restrict__ o1tabs,
                 UnitMeta *__restrict__ meta) {
    constexpr int NTAB = MODE == M_NIB ? 1 : 17;
    __shared__ uint32_t sm[(MODE == M_O1 ? 1 : NTAB * 8) * NT];
    const size_t nthreads = (size_t)gridDim.x * NT, gtid = (size_t)blockIdx.x * NT + threadIdx.x;
    for (size_t u = gtid; u < g.n_units; u += nthreads) {
        size_t j, start, len; uint32_t b;
        unit_span(g, u, j, b, start, len);
        UnitMeta m;
        if (len == 0) { m.len = 0; m.a_off = m.a_len = m.b_off = m.b_len = 0; m.flags = 0; m.pref = 0; m.pad = 0; meta[u] = m; continue; }
        const uint8_t *ip = in + start;
        uint32_t n = (uint32_t)len;
        uint32_t *rec = recs + u * rec_stride, nrec = 0, ntail = 0;
        if (MODE == M_NIB) {
            SmTab<NT> t{sm + threadIdx.x};
            tab_init(t);
            for (uint32_t i = 0; i < n; i++) {                          // anscdf.c:120-127 (state ids are positional)
                uint32_t c, f; tab_enc(t, ip[i], c, f);
                rec[nrec++] = f | c << 16;
            }
            ntail = n & 3;
        } else if (MODE == M_BYTE) {
            for (int k = 0; k < 17; k++) tab_init(SmTab<NT>{sm + k * 8 * NT + threadIdx.x});
            SmTab<NT> th{sm + threadIdx.x};
            for (uint32_t i = 0; i < n; i += 2) {                       // mnenc8x2 anscdf_.h:114-119
                uint32_t x0 = ip[i], x1 = i + 1 < n ? ip[i + 1] : 0;    // odd tail: dummy 0 (anscdf.c:581)
                uint4 r; uint32_t c, f;
                tab_enc(th, x0 >> 4, c, f); r.x = f | c << 16;
                tab_enc(SmTab<NT>{sm + (1 + (x0 >> 4)) * 8 * NT + threadIdx.x}, x0 & 15, c, f); r.y = f | c << 16;
                tab_enc(th, x1 >> 4, c, f); r.z = f | c << 16;
                tab_enc(SmTab<NT>{sm + (1 + (x1 >> 4)) * 8 * NT + threadIdx.x}, x1 & 15, c, f); r.w = f | c << 16;
                *(uint4 *)(rec + nrec) = r; nrec += 4;
            }
        } else {
            uint32_t *tb = o1tabs + gtid * O1_TAB_WORDS;                // [ctx][17][8]
            for (size_t k = 0; k < 256 * 17; k++) tab_init(GmTab{tb + k * 8});
            uint32_t cx = start > j * g.chunk ? in[start - 1] : 0;      // cx carries across blocks of a call (anscdf.c:608)
            for (uint32_t i = 0; i < n; i += 2) {                       // mnenc8x2x anscdf_.h:121-126
                uint32_t x0 = ip[i], x1 = i + 1 < n ? ip[i + 1] : 0;
                uint4 r; uint32_t c, f;
                uint32_t *t0 = tb + (size_t)cx * 17 * 8;
                tab_enc(GmTab{t0}, x0 >> 4, c, f); r.x = f | c << 16;
                tab_enc(GmTab{t0 + (1 + (x0 >> 4)) * 8}, x0 & 15, c, f); r.y = f | c << 16;
                uint32_t *t1 = tb + (size_t)x0 * 17 * 8;
                tab_enc(GmTab{t1}, x1 >> 4, c, f); r.z = f | c << 16;
                tab_enc(GmTab{t1 + (1 + (x1 >> 4)) * 8}, x1 & 15, c, f); r.w = f | c << 16;
                cx = x1;
                *(uint4 *)(rec + nrec) = r; nrec += 4;
            }
        }
        if (MODE == M_NIB) rans_adapt_flush<2>(rec, nrec, ntail, slots + u * slot_stride, (int)slot_stride, n, m);
        else               rans_adapt_flush<4>(rec, nrec, 0,     slots + u * slot_stride, (int)slot_stride, n, m);
        meta[u] = m;
    }
}

// ======================================================================================================
// rANS adaptive decoders.  Thread per call; the 4 MiB blocks of a call are sequential by format (the stream
// has no block directory: block k+1 starts where decoding block k stopped, anscdf.c:588-605).
// ======================================================================================================
template <int MODE, int NT>
__global__ void __launch_bounds__(NT)
k_rans_adapt_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                 uint32_t *__restrict__ o1tabs, unsigned flags) {
    constexpr int NTAB = MODE == M_NIB ? 1 : 17;
    __shared__ uint32_t sm[(MODE == M_O1 ? 1 : NTAB * 8) * NT];
    const size_t nthreads = (size_t)gridDim.x * NT, gtid = (size_t)blockIdx.x * NT + threadIdx.x;
    const uint8_t *gend = in + in_off[g.n_calls];
    for (size_t j = gtid; j < g.n_calls; j += nthreads) {
        size_t start, len; call_span(g, j, start, len);
        uint64_t so = in_off[j], sl = in_off[j + 1] - so;
        if (sl == len) { thread_copy(out + start, in + so, len); continue; }
        RansReader rd; rd.ip = in + so; rd.end = gend;
        uint8_t *op = out + start;
        uint32_t cx = 0;
        for (size_t pos = 0; pos < len; pos += ANS_BLOCK) {
            uint32_t n = (uint32_t)(len - pos < ANS_BLOCK ? len - pos : ANS_BLOCK);
            uint8_t *bp = op + pos;
            if (MODE == M_NIB) {
                SmTab<NT> t{sm + threadIdx.x};
                tab_init(t);
                uint32_t s0 = rd.get32(), s1 = rd.get32(), n4 = n & ~3u, i = 0;
#define TRC_AD(_s_, _dst_) { uint32_t c, f, r = _s_ & PROB_MASK; unsigned x = tab_dec_ans(t, r, c, f); \
                             _s_ = f * (_s_ >> PROB_BITS) + r - c; _s_ = rd.refill(_s_); _dst_ = (uint8_t)x; }
                for (; i < n4; i += 4) { TRC_AD(s0, bp[i]) TRC_AD(s1, bp[i + 1]) TRC_AD(s0, bp[i + 2]) TRC_AD(s1, bp[i + 3]) }   // anscdf.c:98-103
                for (; i < n; i++) { if (flags & 1u) TRC_AD(s0, bp[i]) else TRC_AD(s1, bp[i]) }                            // anscdf.c:104
#undef TRC_AD
            } else {
                uint32_t *tb = nullptr;
                if (MODE == M_BYTE) { for (int k = 0; k < 17; k++) tab_init(SmTab<NT>{sm + k * 8 * NT + threadIdx.x}); }
                else { tb = o1tabs + gtid * O1_TAB_WORDS; for (size_t k = 0; k < 256 * 17; k++) tab_init(GmTab{tb + k * 8}); }
                uint32_t s0 = rd.get32(), s1 = rd.get32(), s2 = rd.get32(), s3 = rd.get32();
                for (uint32_t i = 0; i < n; i += 2) {                   // mndec8x2 / mndec8x2x anscdf_.h:152-174
                    uint32_t c, f, r, yh, yl, x0, x1;
                    if (MODE == M_BYTE) {
                        SmTab<NT> th{sm + threadIdx.x};
                        r = s0 & PROB_MASK; yh = tab_dec_ans(th, r, c, f); s0 = f * (s0 >> PROB_BITS) + r - c;
                        r = s1 & PROB_MASK; yl = tab_dec_ans(SmTab<NT>{sm + (1 + yh) * 8 * NT + threadIdx.x}, r, c, f); s1 = f * (s1 >> PROB_BITS) + r - c;
                        x0 = yh << 4 | yl;
                        r = s2 & PROB_MASK; yh = tab_dec_ans(th, r, c, f); s2 = f * (s2 >> PROB_BITS) + r - c;
                        r = s3 & PROB_MASK; yl = tab_dec_ans(SmTab<NT>{sm + (1 + yh) * 8 * NT + threadIdx.x}, r, c, f); s3 = f * (s3 >> PROB_BITS) + r - c;
                        x1 = yh << 4 | yl;
                    } else {
                        uint32_t *t0 = tb + (size_t)cx * 17 * 8;
                        r = s0 & PROB_MASK; yh = tab_dec_ans(GmTab{t0}, r, c, f); s0 = f * (s0 >> PROB_BITS) + r - c;
                        r = s1 & PROB_MASK; yl = tab_dec_ans(GmTab{t0 + (1 + yh) * 8}, r, c, f); s1 = f * (s1 >> PROB_BITS) + r - c;
                        x0 = yh << 4 | yl;
                        uint32_t *t1 = tb + (size_t)x0 * 17 * 8;
                        r = s2 & PROB_MASK; yh = tab_dec_ans(GmTab{t1}, r, c, f); s2 = f * (s2 >> PROB_BITS) + r - c;
                        r = s3 & PROB_MASK; yl = tab_dec_ans(GmTab{t1 + (1 + yh) * 8}, r, c, f); s3 = f * (s3 >> PROB_BITS) + r - c;
                        x1 = yh << 4 | yl;
                        cx = x1;
                    }
                    s0 = rd.refill(s0); s1 = rd.refill(s1); s2 = rd.refill(s2); s3 = rd.refill(s3);
                    bp[i] = (uint8_t)x0;
                    if (i + 1 < n) bp[i + 1] = (uint8_t)x1;
                }
            }
        }
    }
}

// ======================================================================================================
// Adaptive range coders.  Thread per call, forward coding, 1 or 2 coders.
// ======================================================================================================
enum RMode { R_BYTE1 = 0, R_BYTE2 = 1, R_NIB1 = 2, R_NIB2 = 3 };

template <int RM, int NT>
__global__ void __launch_bounds__(NT)
k_rc_adapt_enc(const uint8_t *__restrict__ in, Geom g, uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    constexpr bool BYTE = RM == R_BYTE1 || RM == R_BYTE2;
    constexpr int NTAB = BYTE ? 17 : 1;
    __shared__ uint32_t sm[NTAB * 8 * NT];
    size_t j = (size_t)blockIdx.x * NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + j * slot_stride;
    for (int k = 0; k < NTAB; k++) tab_init(SmTab<NT>{sm + k * 8 * NT + threadIdx.x});
    SmTab<NT> th{sm + threadIdx.x};
    const int64_t thr = rc_thr(n);
    UnitMeta m; m.pref = 0; m.pad = 0; m.a_off = 0; m.b_off = 0; m.b_len = 0; m.flags = 0;
    bool raw = false;
    uint32_t c, f;
    if (RM == R_BYTE1 || RM == R_NIB1) {
        RcEnc e; e.init(slot);
        for (size_t i = 0; i < n; i++) {
            uint32_t x = ip[i];
            if (BYTE) {                                                 // cdf8e rccdf_.h:30-34
                tab_enc(th, x >> 4, c, f); e.encode(c, f);
                tab_enc(SmTab<NT>{sm + (1 + (x >> 4)) * 8 * NT + threadIdx.x}, x & 15, c, f); e.encode(c, f);
            } else { tab_enc(th, x, c, f); e.encode(c, f); }            // cdf4e rccdf_.h:28
            if ((int64_t)e.pos >= thr) { raw = true; break; }           // OVERFLOW rccdf.c:206,272
        }
        if (!raw) e.flush();
        m.a_len = raw ? 0 : e.pos; m.len = raw ? (uint32_t)n : e.pos; m.flags = raw ? UM_RAW : 0;
    } else {
        const uint32_t b1ref = 4 + (uint32_t)(n / 2);                   // rccdf.c:232,304
        const uint32_t b1 = (b1ref + 64 + 15) & ~15u;
        RcEnc e0, e1; e0.init(slot + 4); e1.init(slot + b1);
        bool quirk = false;
        size_t i = 0;
        if (RM == R_BYTE2) {
            size_t n4 = n & ~(size_t)3;
            for (; i < n4 && !raw; i += 4) {                            // rccdf.c:235-241
                for (int k = 0; k < 4; k++) {                           // cdf8e2 rccdf_.h:36-40
                    uint32_t x = ip[i + k];
                    tab_enc(th, x >> 4, c, f); e0.encode(c, f);
                    tab_enc(SmTab<NT>{sm + (1 + (x >> 4)) * 8 * NT + threadIdx.x}, x & 15, c, f); e1.encode(c, f);
                }
                if ((int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref) raw = true;      // OVERFLOWI rccdf.c:240
            }
            if (!raw) for (; i < n; i++) {
                uint32_t x = ip[i];
                tab_enc(th, x >> 4, c, f); e0.encode(c, f);
                tab_enc(SmTab<NT>{sm + (1 + (x >> 4)) * 8 * NT + threadIdx.x}, x & 15, c, f); e1.encode(c, f);
            }
        } else {
            size_t n2 = n & ~(size_t)1;
            for (; i < n2; i += 2) {                                    // rccdf.c:308-315
                uint32_t x0 = ip[i], x1 = ip[i + 1];
                c = th.entry(x0); f = th.entry(x0 + 1) - c; e0.encode(c, f);
                c = th.entry(x1); f = th.entry(x1 + 1) - c; e1.encode(c, f);
                tab_upd(th, x0); tab_upd(th, x1);
                if ((int64_t)b1ref + e1.pos >= thr) { raw = true; quirk = true; break; }    // OVERFLOW on op1 only (rccdf.c:314)
            }
            if (!raw && i < n) { tab_enc(th, ip[i], c, f); e0.encode(c, f); }
        }
        if (quirk) {   // reference returns op0 - out with the raw copy in out (rccdf.c:314,321-322)
            m.a_len = 0; m.len = 4 + e0.pos; m.flags = UM_RAW | UM_QUIRK4;
        } else {
            if (!raw) {
                e0.flush(); e1.flush();
                *(uint32_t *)slot = e0.pos;
                if ((int64_t)(4 + e0.pos + e1.pos) >= thr) raw = true;
            }
            m.a_len = raw ? 0 : 4 + e0.pos; m.b_off = b1; m.b_len = raw ? 0 : e1.pos;
            m.len = raw ? (uint32_t)n : 4 + e0.pos + e1.pos; m.flags = raw ? UM_RAW : 0;
        }
    }
    meta[j] = m;
}

// 16-entry linear search of the adaptive RC decoders (_cdflget16 turborc_.h:271-291): first x with
// entry[x+1] * range > code, 15 if none.  Entries come packed two per word.
__device__ __forceinline__ unsigned rc_search16(const uint32_t m[8], const RcDec &d) {
    unsigned x = 0;
#pragma unroll
    for (int e = 1; e < 16; e++) {
        uint32_t v = (e & 1) ? m[e >> 1] >> 16 : m[e >> 1] & 0xffffu;
        x += ((uint64_t)v * d.range <= d.code) ? 1u : 0u;               // monotone in e, so the count is the index
    }
    return x;
}
template <class Tab> __device__ __forceinline__ unsigned rc_dec_nib(Tab t, RcDec &d) {   // cdf4d rccdf_.h:48
    uint32_t m[8]; t.load(m);
    d.shift();
    unsigned x = rc_search16(m, d);
    uint32_t c0 = t.entry(x), c1 = t.entry(x + 1);
    d.update(c0, c1);
    uint32_t c2 = c0 * 0x00010001u | 0x80008000u;
#pragma unroll
    for (int w = 0; w < 8; w++) m[w] = adapt_word(m[w], adapt_le(m[w], c2), w);
    t.store(m);
    return x;
}

template <int RM, int NT>
__global__ void __launch_bounds__(NT)
k_rc_adapt_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    constexpr bool BYTE = RM == R_BYTE1 || RM == R_BYTE2;
    constexpr int NTAB = BYTE ? 17 : 1;
    __shared__ uint32_t sm[NTAB * 8 * NT];
    size_t j = (size_t)blockIdx.x * NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls], *stream = in + so;
    uint8_t *op = out + start;
    if (sl == n) { thread_copy(op, stream, n); return; }
    for (int k = 0; k < NTAB; k++) tab_init(SmTab<NT>{sm + k * 8 * NT + threadIdx.x});
    SmTab<NT> th{sm + threadIdx.x};
    if (RM == R_BYTE1 || RM == R_NIB1) {
        RcDec d; d.init(stream, gend);
        for (size_t i = 0; i < n; i++) {
            if (BYTE) { unsigned h = rc_dec_nib(th, d), l = rc_dec_nib(SmTab<NT>{sm + (1 + h) * 8 * NT + threadIdx.x}, d); op[i] = (uint8_t)(h << 4 | l); }
            else op[i] = (uint8_t)rc_dec_nib(th, d);
        }
    } else {
        uint32_t len0 = ld_u32_clamped(stream, gend);
        const uint8_t *p1 = stream + 4 + len0;
        if (p1 > gend || p1 < stream) p1 = gend;
        RcDec d0, d1; d0.init(stream + 4, gend); d1.init(p1, gend);
        if (RM == R_BYTE2) {
            for (size_t i = 0; i < n; i++) {                            // cdf8d2 rccdf_.h:63-73
                unsigned h = rc_dec_nib(th, d0), l = rc_dec_nib(SmTab<NT>{sm + (1 + h) * 8 * NT + threadIdx.x}, d1);
                op[i] = (uint8_t)(h << 4 | l);
            }
        } else {
            size_t i = 0, n2 = n & ~(size_t)1;
            for (; i < n2; i += 2) {                                    // rccdf.c:285-297
                uint32_t m[8]; th.load(m);
                d0.shift(); d1.shift();
                unsigned x0 = rc_search16(m, d0), x1 = rc_search16(m, d1);
                d0.update(th.entry(x0), th.entry(x0 + 1)); d1.update(th.entry(x1), th.entry(x1 + 1));
                tab_upd(th, x0); tab_upd(th, x1);
                op[i] = (uint8_t)x0; op[i + 1] = (uint8_t)x1;
            }
            if (i < n) op[i] = (uint8_t)rc_dec_nib(th, d0);
        }
    }
}

}  // namespace trc
