// vlc.cuh -- the VLC-over-CDF integer codecs: anscdf{u,uz,v,vz}{enc,dec}16, anscdf{v,vz}{enc,dec}32
// (anscdf.c:139-483) and rccdf{v,vz,u}{enc,dec}{16,32} (rccdf.c:392-632).
//
// Every 16/32-bit integer (optionally the zigzag of its delta to the previous one, rcutil_.h:144-148) goes through Turbo
// VLC (include_/vlcbit.h:24-63): values below 2^(vn+1) are their own symbol, larger ones become an exponent symbol plus mb
// mantissa bits that go to a bit stream growing DOWNWARD from the end of the output (rcutil_.h:163-192).  The symbol (< 76
// or < 136) is one or two nibbles over two adaptive 16-entry tables (cdfenc6/7 anscdf_.h:206-230, cdfe6/7 rccdf_.h:100-122),
// coded by the 2-state blocked rANS (records + mnflush per block of 4 Mi ELEMENTS) or by one range coder.
// Stream = [u32 total][entropy-coded part][bit stream]; the decoder's bit reader starts at in + total.
//
// First GPU form: one lane per call (throughput from many calls per batch).  The lane works in its slot as the reference
// works in `out`: header + entropy-coded part grow up from offset 0, the bit stream grows down from offset inlen; the pack
// kernel concatenates the two pieces, so the reference's final memmove never happens.  Only FINAL bytes of the bit stream
// are stored (the reference stores whole 64-bit words and overwrites them), and the rANS LIFO of a block runs in an aligned
// scratch half of the slot with the reference's guards evaluated on virtual offsets -- same decisions, same bytes.
#pragma once
#include "trc_common.cuh"
#include "adaptive.cuh"

namespace trc {

constexpr int VLC_NT = 64;                                       // 2 tables x 32 B x 64 lanes = 4 KB of shared memory

struct VlcParam { int w32; unsigned vn; int zz; };
__host__ __device__ inline bool codec_vlc(int c) { return c >= ANSU16 && c <= RCU32; }
__host__ __device__ inline bool codec_vlc_ans(int c) { return c >= ANSU16 && c <= ANSVZ32; }
__host__ __device__ inline VlcParam vlc_param(int c) {
    switch (c) {
    case ANSU16:  return {0, 1, 0};  case ANSUZ16: return {0, 1, 1};
    case ANSV16:  return {0, 2, 0};  case ANSVZ16: return {0, 2, 1};
    case ANSV32:  return {1, 2, 0};  case ANSVZ32: return {1, 2, 1};
    case RCV16:   return {0, 2, 0};  case RCVZ16:  return {0, 2, 1};
    case RCV32:   return {1, 2, 0};  case RCVZ32:  return {1, 2, 1};
    case RCU16:   return {0, 1, 0};  default:      return {1, 1, 0};      // RCU32
    }
}

// ---- bit IO, right to left (biteinir / bitput / bitenormr / bitflushr, bitdinir / bitdnormr / bitpeek / bitrmv) ------------
struct BitW {
    uint64_t bw; unsigned br; uint8_t *p;                        // p == the reference's pointer (its 64-bit store would cover [p, p+8))
    __device__ __forceinline__ void init(uint8_t *end) { bw = 0; br = 64; p = end - 8; }
    __device__ __forceinline__ void put(unsigned nb, uint32_t x) { br -= nb; bw |= (uint64_t)x << br; }
    __device__ __forceinline__ void final_bytes(unsigned k) { for (unsigned j = 0; j < (k >> 3); j++) p[7 - j] = (uint8_t)(bw >> (56 - 8 * j)); }
    __device__ __forceinline__ void norm() { const unsigned k = (64 - br) & ~7u; final_bytes(k); p -= k >> 3; bw <<= k; br += k; }
    __device__ __forceinline__ void flush() { const unsigned k = (64 + 7 - br) & ~7u; final_bytes(k); p -= k >> 3; p += 8; }
};
struct BitR {
    uint64_t bw; unsigned br; const uint8_t *p, *lo, *hi;        // reads outside [lo, hi) return zero (corrupt streams only)
    __device__ __forceinline__ void init(const uint8_t *end, const uint8_t *l, const uint8_t *h) { bw = 0; br = 0; p = end - 8; lo = l; hi = h; }
    __device__ __forceinline__ void norm() {
        p -= br >> 3; br &= 7;
        bw = (p >= lo && p + 8 <= hi) ? ((uint64_t)ld_u32(p) | (uint64_t)ld_u32(p + 4) << 32) : 0ull;
    }
};
__device__ __forceinline__ uint32_t vlc_put(BitW &b, unsigned vn, uint32_t x) {           // bitvrput, vb = 0 (vlcbit.h:40-48)
    if (x >= (1u << (vn + 1))) {
        const unsigned f = (31 - __clz((int)x)) - vn, expo = ((f + 1) << vn) + ((x >> f) & ((1u << vn) - 1)), mb = (expo >> vn) - 1;
        b.put(mb, x & ((1u << mb) - 1)); b.norm();
        x = expo;
    }
    return x;
}
__device__ __forceinline__ uint32_t vlc_get(BitR &b, unsigned vn, uint32_t x) {           // bitvrget (vlcbit.h:59-64)
    if (x >= (1u << (vn + 1))) {
        b.norm();
        const unsigned mb = (x >> vn) - 1;
        const uint32_t ma = (uint32_t)((b.bw << b.br) >> (64 - mb));
        x = (((1u << vn) + (x & ((1u << vn) - 1))) << mb) + ma;
        b.br += mb;
    }
    return x;
}
__device__ __forceinline__ uint32_t vlc_load(const uint8_t *in, size_t i, int w32) { return w32 ? ld_u32(in + 4 * i) : ld_u16(in + 2 * i); }
__device__ __forceinline__ void vlc_store(uint8_t *out, size_t i, int w32, uint32_t v) {
    uint8_t *q = out + (w32 ? 4 : 2) * i;
    q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8);
    if (w32) { q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); }
}
__device__ __forceinline__ uint32_t zz_enc(uint32_t cur, uint32_t prev, int w32) {
    if (w32) { const int32_t d = (int32_t)(cur - prev); return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31); }
    const int16_t d = (int16_t)(cur - prev); return (uint16_t)(((uint16_t)d << 1) ^ (uint16_t)(d >> 15));
}
__device__ __forceinline__ uint32_t zz_dec(uint32_t r, int w32) {
    if (w32) return (r >> 1) ^ (0u - (r & 1));
    const uint16_t v = (uint16_t)r; return (uint16_t)((v >> 1) ^ (uint16_t)(0u - (v & 1)));
}

// LIFO scratch of the rANS blocks: the upper part of the slot (make_plan sizes the slot as 2 x al16(inlen) + 256)
__host__ __device__ inline size_t vlc_lifo_off(size_t unit_max) { return ((unit_max + 15) & ~(size_t)15) + 64; }

// ---- encoders ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VLC_NT)
k_vlc_ans_enc(const uint8_t *__restrict__ in, Geom g, VlcParam P, uint8_t *__restrict__ slots, size_t slot_stride, uint32_t *__restrict__ recs,
              size_t rec_stride, UnitMeta *__restrict__ meta) {
    __shared__ uint32_t sm[2 * 8 * VLC_NT];
    const size_t j = (size_t)blockIdx.x * VLC_NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, nbytes; call_span(g, j, start, nbytes);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + j * slot_stride;
    const size_t lifo_off = vlc_lifo_off(g.unit_max);
    const int lcap = (int)(slot_stride - lifo_off) & ~15;        // aligned LIFO capacity (>= inlen + 64)
    uint8_t *lifo = slot + lifo_off;
    uint32_t *rec = recs + j * rec_stride;
    SmTab<VLC_NT> m0{sm + threadIdx.x}, m1{sm + 8 * VLC_NT + threadIdx.x};
    const unsigned esz = P.w32 ? 4 : 2, lim = P.vn == 1 ? 12 : 8;
    const size_t n = (nbytes + esz - 1) / esz, blk = n < ANS_BLOCK ? n : ANS_BLOCK;
    int64_t op = 4;                                              // offsets inside the slot == the reference's out
    const int64_t out_end = (int64_t)nbytes;
    BitW b; b.init(slot + out_end);
    uint32_t cx = 0;
    bool raw = false;
    for (size_t pos = 0; pos < n && !raw; pos += blk) {
        tab_init(m0); tab_init(m1);                              // CDF16DEC0 x2 per block (anscdf.c:153-154)
        const size_t cnt = n - pos < blk ? n - pos : blk;
        uint32_t nrec = 0;
        for (size_t i = 0; i < cnt; i++) {
            const uint32_t v = vlc_load(ip, pos + i, P.w32);
            uint32_t x = P.zz ? zz_enc(v, cx, P.w32) : v, c, f;
            cx = v;
            x = vlc_put(b, P.vn, x);
            if (x < lim) { tab_enc(m0, x, c, f); rec[nrec++] = f | c << 16 | 1u << 31; }                       // cdfenc6 / cdfenc7: state 1
            else { x -= lim; tab_enc(m0, (x >> 4) + lim, c, f); rec[nrec++] = f | c << 16 | 1u << 31;
                   tab_enc(m1, x & 15, c, f); rec[nrec++] = f | c << 16; }                                    //                  + state 0
        }
        // mnflush(op, bp - 8, ...) anscdf_.h:128-138: the LIFO ends at vend; guards on virtual offsets
        const int64_t vend = (int64_t)(b.p - slot) - 8;
        RansWriter w; w.init(lifo, lcap);
        uint32_t st[2] = { ANS_L, ANS_L };
        bool em;
        while (nrec) {
            const uint32_t r = rec[--nrec];
            if (vend - (lcap - w.pos) <= op + 2 + 8) { raw = true; break; }
            const unsigned si = r >> 31;
            st[si] = rans_enc_step_rec(st[si], r & 0x7fffffffu, w, em);
        }
        if (raw) break;
        w.finish_words(); w.put32_final(st[0]); w.put32_final(st[1]);
        const int64_t l = lcap - w.pos;
        if (vend - l <= op || op + l >= vend) { raw = true; break; }
        for (int64_t k = 0; k < l; k += 2) *(uint16_t *)(slot + op + k) = *(const uint16_t *)(lifo + w.pos + k);   // op, l, w.pos are even
        op += l;
    }
    UnitMeta m; m.pref = 0; m.pad = 0; m.a_off = 0; m.a_len = 0; m.b_off = 0; m.b_len = 0;
    if (!raw) {
        b.flush();
        const int64_t bo = (int64_t)(b.p - slot), l = out_end - bo;
        if (op + l >= out_end) raw = true;
        else { *(uint32_t *)slot = (uint32_t)(op + l); m.a_len = (uint32_t)op; m.b_off = (uint32_t)bo; m.b_len = (uint32_t)l; }
    }
    m.len = raw ? (uint32_t)nbytes : m.a_len + m.b_len; m.flags = raw ? UM_RAW : 0;
    meta[j] = m;
}

__global__ void __launch_bounds__(VLC_NT)
k_vlc_rc_enc(const uint8_t *__restrict__ in, Geom g, VlcParam P, uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ uint32_t sm[2 * 8 * VLC_NT];
    const size_t j = (size_t)blockIdx.x * VLC_NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, nbytes; call_span(g, j, start, nbytes);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + j * slot_stride;
    SmTab<VLC_NT> m0{sm + threadIdx.x}, m1{sm + 8 * VLC_NT + threadIdx.x};
    tab_init(m0); tab_init(m1);
    const unsigned esz = P.w32 ? 4 : 2, lim = P.vn == 1 ? 12 : 8;
    const size_t n = (nbytes + esz - 1) / esz;
    const int64_t out_end = (int64_t)nbytes;
    RcEnc e; e.init(slot + 4);
    BitW b; b.init(slot + out_end);
    uint32_t cx = 0, c, f;
    bool raw = false;
    for (size_t i = 0; i < n; i++) {
        const uint32_t v = vlc_load(ip, i, P.w32);
        uint32_t x = P.zz ? zz_enc(v, cx, P.w32) : v;
        cx = v;
        x = vlc_put(b, P.vn, x);
        if (x < lim) { tab_enc(m0, x, c, f); e.encode(c, f); }                                              // cdfe6 / cdfe7
        else { x -= lim; tab_enc(m0, (x >> 4) + lim, c, f); e.encode(c, f); tab_enc(m1, x & 15, c, f); e.encode(c, f); }
        if ((int64_t)4 + e.pos + 8 >= (int64_t)(b.p - slot)) { raw = true; break; }                         // rccdf.c:406
    }
    UnitMeta m; m.pref = 0; m.pad = 0; m.a_off = 0; m.a_len = 0; m.b_off = 0; m.b_len = 0;
    if (!raw) {
        e.flush(); b.flush();
        const int64_t op = 4 + (int64_t)e.pos, bo = (int64_t)(b.p - slot), l = out_end - bo;
        *(uint32_t *)slot = (uint32_t)(op + l);
        if (op + l >= (int64_t)rc_thr(nbytes)) raw = true;                                                  // OVERFLOW rccdf.c:409
        else { m.a_len = (uint32_t)op; m.b_off = (uint32_t)bo; m.b_len = (uint32_t)l; }
    }
    m.len = raw ? (uint32_t)nbytes : m.a_len + m.b_len; m.flags = raw ? UM_RAW : 0;
    meta[j] = m;
}

// ---- decoders ------------------------------------------------------------------------------------------------------------
template <bool RC>
__global__ void __launch_bounds__(VLC_NT)
k_vlc_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g, VlcParam P) {
    __shared__ uint32_t sm[2 * 8 * VLC_NT];
    const size_t j = (size_t)blockIdx.x * VLC_NT + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t start, nbytes; call_span(g, j, start, nbytes);
    const uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls], *stream = in + so;
    uint8_t *op = out + start;
    if (sl == nbytes) { thread_copy(op, stream, nbytes); return; }                                            // raw chunk (CCPY turborc.c:434)
    SmTab<VLC_NT> m0{sm + threadIdx.x}, m1{sm + 8 * VLC_NT + threadIdx.x};
    const unsigned esz = P.w32 ? 4 : 2, lim = P.vn == 1 ? 12 : 8;
    const size_t n = (nbytes + esz - 1) / esz, blk = n < ANS_BLOCK ? n : ANS_BLOCK;
    uint32_t total = ld_u32_clamped(stream, gend);
    if (total > sl) total = (uint32_t)sl;                                                                      // corrupt header: stay inside the chunk
    BitR b; b.init(stream + total, stream, gend);
    uint32_t cx = 0;
    if (RC) {
        tab_init(m0); tab_init(m1);
        RcDec d; d.init(stream + 4, gend);
        for (size_t i = 0; i < n; i++) {
            uint32_t x = rc_dec_nib(m0, d);                                                                    // cdfd6 / cdfd7
            if (x >= lim) { const uint32_t y = rc_dec_nib(m1, d); x = ((x - lim) << 4 | y) + lim; }
            uint32_t r = vlc_get(b, P.vn, x);
            if (P.zz) { cx += zz_dec(r, P.w32); r = cx; }
            vlc_store(op, i, P.w32, r);
        }
    } else {
        RansReader rd; rd.ip = stream + 4; rd.end = gend;
        for (size_t pos = 0; pos < n; pos += blk) {
            tab_init(m0); tab_init(m1);
            const size_t cnt = n - pos < blk ? n - pos : blk;
            uint32_t s0 = rd.get32(), s1 = rd.get32(), c, f;                                                   // mnfill(st, ip, 2)
            for (size_t i = 0; i < cnt; i++) {
                uint32_t rr = s0 & PROB_MASK, x = tab_dec_ans(m0, rr, c, f);                                   // cdfdec6 / cdfdec7: mndec4 on st[0]
                s0 = f * (s0 >> PROB_BITS) + rr - c; s0 = rd.refill(s0);
                if (x >= lim) {
                    rr = s1 & PROB_MASK; const uint32_t y = tab_dec_ans(m1, rr, c, f);
                    s1 = f * (s1 >> PROB_BITS) + rr - c; s1 = rd.refill(s1);
                    x = ((x - lim) << 4 | y) + lim;
                }
                uint32_t r = vlc_get(b, P.vn, x);
                if (P.zz) { cx += zz_dec(r, P.w32); r = cx; }
                vlc_store(op, pos + i, P.w32, r);
            }
        }
    }
}

}  // namespace trc
