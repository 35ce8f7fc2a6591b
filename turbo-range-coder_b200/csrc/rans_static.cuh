// rans_static.cuh -- static-CDF rANS, 2 interleaved states per call (TRC_ANS4S).
//
// Bit-exact with anscdf4senc / anscdf4sdec (reference anscdf.c:57-85, anscdf_.h:43-103, cdf_.h:37,61-66).
// One GPU thread owns one reference call: both rANS states live in its registers, so a warp carries
// 64 interleaved states over 32 independent calls.  The symbol table (reciprocals for the encoder, slot->
// symbol LUT for the decoder) is staged in shared memory once per CTA.
#pragma once
#include "trc_common.cuh"

namespace trc {

// ---- LIFO 16-bit word writer (descending addresses, _putc anscdf_.h:43) ------------------------------
// Two consecutive words are merged in a register and stored as one aligned 32-bit word.
struct RansWriter {
    uint8_t *base;       // slot start (16-byte aligned)
    int      pos;        // byte offset of the lowest byte written so far; starts at cap (multiple of 4)
    uint32_t acc;
    __device__ __forceinline__ void init(uint8_t *b, int cap) { base = b; pos = cap; acc = 0; }
    __device__ __forceinline__ void put16(uint32_t w) {
        pos -= 2;
        acc = __byte_perm(w, acc, 0x5410);              // acc = acc << 16 | (w & 0xffff)
        if ((pos & 2) == 0) *(uint32_t *)(base + pos) = acc;
    }
    __device__ __forceinline__ void finish_words() {    // a lone half word is still in acc
        if (pos & 2) st_u16(base + pos, acc);
    }
    __device__ __forceinline__ void put32_final(uint32_t v) { pos -= 4; st_u32_a2(base + pos, v); }   // eceflush anscdf_.h:46
};

// ---- encoder table entry: exact division by multiplication -------------------------------------------
// For 1 <= f <= 2^15 and s < 2^31:  s / f == umulhi(s, rcp) >> sh  with sh = ceil(log2 f) - 1 and
// rcp = ceil(2^(sh+32) / f); f == 1 uses rcp = 2^32-1 (q = s-1) and folds the missing step into the bias.
// (Alverson-style; verified exhaustively over f and boundary s in tests/test_host_logic.py.)
// entry = { rcp, f << 16, (2^15 - f) | sh << 16, bias }   ->   s' = s + bias + q * (2^15 - f)
__host__ __device__ inline uint4 rans_enc_entry(uint32_t cum, uint32_t f) {
    uint4 e;
    if (f < 2) {
        e.x = 0xFFFFFFFFu; e.y = f << 16; e.z = (PROB_TOTAL - f);           // sh = 0
        e.w = cum + PROB_TOTAL - 1;
    } else {
        uint32_t sh = 0;
        while (f > (1u << sh)) sh++;
        e.x = (uint32_t)(((1ull << (sh + 31)) + f - 1) / f);
        e.y = f << 16;
        e.z = (PROB_TOTAL - f) | (sh - 1) << 16;
        e.w = cum;
    }
    return e;
}

// one encode step, ece anscdf_.h:90-94
__device__ __forceinline__ uint32_t rans_enc_step(uint32_t s, const uint4 e, RansWriter &w) {
    if (s >= e.y) { w.put16(s); s >>= 16; }
    uint32_t q = __umulhi(s, e.x) >> (e.z >> 16);
    return s + e.w + q * (e.z & 0xffffu);
}
// generic step with a real division (per-call tables that are not staged in shared memory)
__device__ __forceinline__ uint32_t rans_enc_step_div(uint32_t s, uint32_t cum, uint32_t f, RansWriter &w) {
    if (s >= (f << 16)) { w.put16(s); s >>= 16; }
    uint32_t q = s / f;
    return s + (q << PROB_BITS) - q * f + cum;
}

constexpr int RANS_S_NT = 128;

template <bool FAST>
__device__ __forceinline__ uint32_t rs_step(uint32_t s, uint32_t x, const uint4 *etab, const cdf_t *gcdf, RansWriter &w) {
    if (FAST) return rans_enc_step(s, etab[x], w);
    uint32_t c = gcdf[x];
    return rans_enc_step_div(s, c, (uint32_t)gcdf[x + 1] - c, w);
}

template <bool FAST>
__device__ inline void rans_static_enc_call(const uint8_t *ip, uint32_t n, const uint4 *etab, const cdf_t *gcdf,
                                            uint8_t *slot, int cap, UnitMeta &m) {
    RansWriter w; w.init(slot, cap);
    uint32_t s0 = ANS_L, s1 = ANS_L;
    uint32_t i = n, n4 = n & ~3u;
    bool ovf = false;
    while (i > n4) { i--; s0 = rs_step<FAST>(s0, ip[i], etab, gcdf, w); }            // anscdf.c:62-64
    const bool al4 = (((uintptr_t)ip) & 3) == 0;
    while (i > 0) {                                                                   // anscdf.c:65-67
        uint32_t v;
        i -= 4;
        if (al4) v = *(const uint32_t *)(ip + i);
        else v = (uint32_t)ip[i] | (uint32_t)ip[i + 1] << 8 | (uint32_t)ip[i + 2] << 16 | (uint32_t)ip[i + 3] << 24;
        s1 = rs_step<FAST>(s1, v >> 24, etab, gcdf, w);
        s0 = rs_step<FAST>(s0, (v >> 16) & 0xff, etab, gcdf, w);
        s1 = rs_step<FAST>(s1, (v >> 8) & 0xff, etab, gcdf, w);
        s0 = rs_step<FAST>(s0, v & 0xff, etab, gcdf, w);
        if ((uint32_t)(cap - w.pos) + 8u >= n) { ovf = true; break; }                 // l >= inlen is already certain
    }
    w.finish_words();
    w.put32_final(s0); w.put32_final(s1);                                            // ansflush anscdf_.h:102
    uint32_t l = (uint32_t)(cap - w.pos);
    bool raw = ovf || l >= n;                                                         // anscdf.c:70
    m.len = raw ? n : l; m.a_off = (uint32_t)w.pos; m.a_len = raw ? 0 : l; m.b_off = 0; m.b_len = 0;
    m.flags = raw ? UM_RAW : 0; m.pref = 0; m.pad = 0;
}

__global__ void __launch_bounds__(RANS_S_NT)
k_rans_static_enc(const uint8_t *__restrict__ in, Geom g, const cdf_t *__restrict__ cdf, unsigned cdfnum, size_t cpc,
                  uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ uint4 etab[256];
    size_t j0 = (size_t)blockIdx.x * RANS_S_NT, j = j0 + threadIdx.x;
    size_t t0 = cpc ? j0 / cpc : 0;
    const cdf_t *c0 = cdf + t0 * CDF_STRIDE;
    for (unsigned x = threadIdx.x; x < 256; x += RANS_S_NT) {
        uint4 e = make_uint4(0, 0, 0, 0);
        if (x < cdfnum) { uint32_t c = c0[x]; e = rans_enc_entry(c, (uint32_t)c0[x + 1] - c); }
        etab[x] = e;
    }
    __syncthreads();
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    size_t t = cpc ? j / cpc : 0;
    UnitMeta m;
    if (t == t0) rans_static_enc_call<true >(in + start, (uint32_t)len, etab, nullptr, slots + j * slot_stride, (int)slot_stride, m);
    else         rans_static_enc_call<false>(in + start, (uint32_t)len, nullptr, cdf + t * CDF_STRIDE, slots + j * slot_stride, (int)slot_stride, m);
    meta[j] = m;
}

// ---- decoder ------------------------------------------------------------------------------------------
struct RansReader {
    const uint8_t *ip, *end;
    __device__ __forceinline__ uint32_t get32() { uint32_t v = ld_u32_clamped(ip, end); ip += 4; return v; }   // ecdini anscdf_.h:47
    __device__ __forceinline__ uint32_t refill(uint32_t s) {                                                  // ecdnorm anscdf_.h:50-73
        if (s < ANS_L) { s = s << 16 | ld_u16_clamped(ip, end); ip += 2; }
        return s;
    }
};

// LUT form: x = lut[r], (freq | cum << 16) = dtab[x]
__device__ __forceinline__ uint32_t rans_dec_step_lut(uint32_t &s, const uint8_t *lut, const uint32_t *dtab, RansReader &rd) {
    uint32_t r = s & PROB_MASK, x = lut[r], e = dtab[x];
    s = (e & 0xffffu) * (s >> PROB_BITS) + r - (e >> 16);                             // STATEUPD cdf_.h:37
    s = rd.refill(s);
    return x;
}
// generic form: binary search in a global table (x = max{ x < n : cdf[x] <= r })
__device__ __forceinline__ uint32_t rans_dec_step_bs(uint32_t &s, const cdf_t *cdf, unsigned n, RansReader &rd) {
    uint32_t r = s & PROB_MASK, x = 0, hi = n;
    while (x + 1 < hi) { uint32_t mid = (x + hi) >> 1; if (cdf[mid] <= r) x = mid; else hi = mid; }
    uint32_t c = cdf[x];
    s = ((uint32_t)cdf[x + 1] - c) * (s >> PROB_BITS) + r - c;
    s = rd.refill(s);
    return x;
}

constexpr int RANS_SD_NT = 128;

template <bool FAST>
__device__ inline void rans_static_dec_call(const uint8_t *stream, const uint8_t *gend, uint8_t *op, uint32_t n,
                                            const uint8_t *lut, const uint32_t *dtab, const cdf_t *gcdf, unsigned cdfnum,
                                            bool ref_tail) {
    RansReader rd; rd.ip = stream; rd.end = gend;
    uint32_t s0 = rd.get32(), s1 = rd.get32();                                        // mnfill anscdf_.h:176
    uint32_t o = 0, n4 = n & ~3u;
    const bool al4 = (((uintptr_t)op) & 3) == 0;
#define TRC_SD(_s_) (FAST ? rans_dec_step_lut(_s_, lut, dtab, rd) : rans_dec_step_bs(_s_, gcdf, cdfnum, rd))
    for (; o < n4; o += 4) {                                                          // anscdf.c:82
        uint32_t a = TRC_SD(s1), b = TRC_SD(s0), c = TRC_SD(s1), d = TRC_SD(s0);
        if (al4) *(uint32_t *)(op + o) = a | b << 8 | c << 16 | d << 24;
        else { op[o] = (uint8_t)a; op[o + 1] = (uint8_t)b; op[o + 2] = (uint8_t)c; op[o + 3] = (uint8_t)d; }
    }
    for (; o < n; o++) {                                                              // anscdf.c:83 (state 0 in the reference)
        uint32_t a = ref_tail ? TRC_SD(s0) : TRC_SD(s1);
        op[o] = (uint8_t)a;
    }
#undef TRC_SD
}

__global__ void __launch_bounds__(RANS_SD_NT)
k_rans_static_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                  const cdf_t *__restrict__ cdf, unsigned cdfnum, size_t cpc, unsigned flags) {
    __shared__ uint8_t  lut[PROB_TOTAL];
    __shared__ uint32_t dtab[256];
    __shared__ uint16_t scdf[CDF_STRIDE];
    size_t j0 = (size_t)blockIdx.x * RANS_SD_NT, j = j0 + threadIdx.x;
    size_t t0 = cpc ? j0 / cpc : 0;
    const cdf_t *c0 = cdf + t0 * CDF_STRIDE;
    for (unsigned x = threadIdx.x; x <= cdfnum; x += RANS_SD_NT) scdf[x] = c0[x];
    __syncthreads();
    for (unsigned x = threadIdx.x; x < cdfnum; x += RANS_SD_NT) dtab[x] = ((uint32_t)scdf[x + 1] - scdf[x]) & 0xffffu | (uint32_t)scdf[x] << 16;
    for (unsigned r = threadIdx.x; r < PROB_TOTAL; r += RANS_SD_NT) {
        unsigned x = 0, hi = cdfnum;
        while (x + 1 < hi) { unsigned mid = (x + hi) >> 1; if (scdf[mid] <= r) x = mid; else hi = mid; }
        lut[r] = (uint8_t)x;
    }
    __syncthreads();
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls];
    if (sl == len) { thread_copy(out + start, in + so, len); return; }                // raw chunk (CCPY rule turborc.c:434)
    size_t t = cpc ? j / cpc : 0;
    if (t == t0) rans_static_dec_call<true >(in + so, gend, out + start, (uint32_t)len, lut, dtab, nullptr, cdfnum, flags & 1u);
    else         rans_static_dec_call<false>(in + so, gend, out + start, (uint32_t)len, nullptr, nullptr, cdf + t * CDF_STRIDE, cdfnum, flags & 1u);
}

}  // namespace trc
