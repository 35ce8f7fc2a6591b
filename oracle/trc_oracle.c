/*
 * trc_oracle.c -- CPU restatement (plain scalar C) of the CDF entropy-coding hot path of
 * powturbo/Turbo-Range-Coder, used ONLY as the parity checker for the B200 kernels.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library (libtrc_b200.so) never
 * links or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked byte-for-byte against the compiled
 * reference (oracle/_ref/libtrcref.so, built by oracle/Makefile from the sources under
 * /root/reference) in tests/test_oracle_vs_ref.py, and against the committed fixtures in
 * tests/golden/ (generated from the compiled reference by tests/golden/make_golden.py).
 * The one function with no reference counterpart (orc_ans_sdec_n, a static rANS decoder for
 * alphabets larger than 16 symbols) is pinned only through the reference *encoder* anscdf4senc
 * (it must invert reference-produced streams) -- see its comment.
 *
 * Each function cites the reference file:line it restates.  Nothing here is copied from the
 * reference: the reference is macro/SIMD code, this is a from-scratch scalar statement of the
 * same integer arithmetic.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef uint16_t cdf_t;               /* include/turborc.h:497 */

enum {
    PROB_BITS   = 15,                 /* ANS_BITS anscdf_.h:33, RC_BITS rccdf.c:37 (non-AVX2 build) */
    PROB_TOTAL  = 1 << PROB_BITS,
    ANS_L       = 1 << 15,            /* ANS_LOW anscdf_.h:40-41 */
    ADAPT_RATE  = 7,                  /* CDFRATE cdf_.h:25 */
    ADAPT_IC    = 10,                 /* IC cdf_.h:35 */
    ADAPT_MIX   = 32736,              /* MIXD cdf_.h:36 = 0x7fff & ~31 */
    ANS_BLOCK   = 1 << 22             /* ANSBLKSIZE anscdf.c:54 */
};

static uint16_t ld16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static void st16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
static void st32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

/* ------------------------------------------------------------------------------------------
 * Adaptive 16-symbol CDF (cdf_.h:26-32 init, cdf_.h:46-50 / 87-97 update; SIMD semantics).
 * Table t[0..16], t[16] = 32768 is never touched.  The SIMD code compares *values*
 * (t[i] > t[x]); the table is strictly increasing so that equals the index compare i > x.
 * ------------------------------------------------------------------------------------------ */
static void adapt_init(cdf_t *t) { for (int j = 0; j <= 16; j++) t[j] = (cdf_t)(j << (PROB_BITS - 4)); }

static void adapt_update(cdf_t *t, unsigned x) {
    for (int i = 0; i < 16; i++) {
        int target = ADAPT_IC * i + (i > (int)x ? ADAPT_MIX : 0);
        int16_t d  = (int16_t)(target - (int)t[i]);          /* 16-bit lane arithmetic */
        t[i] = (cdf_t)(t[i] + (d >> ADAPT_RATE));            /* arithmetic shift (srai_epi16) */
    }
}

/* ------------------------------------------------------------------------------------------
 * rANS primitives.  ece/ecenorm anscdf_.h:48,90-94; ecdnorm anscdf_.h:50-73; STATEUPD cdf_.h:37
 * ------------------------------------------------------------------------------------------ */
static inline uint32_t rans_put(uint32_t st, uint32_t cum, uint32_t freq, uint8_t **ep) {
    if (st >= (freq << 16)) { *ep -= 2; st16(*ep, (uint16_t)st); st >>= 16; }
    uint32_t q = st / freq;
    return st + (q << PROB_BITS) - q * freq + cum;
}

/* static / adaptive symbol search: first i in 0..15 with t[i] > r, minus one (cdf_.h:52-66) */
static inline unsigned rans_find16(const cdf_t *t, uint32_t r) {
    unsigned i = 0;
    while (i < 16 && !(t[i] > r)) i++;
    return i - 1;
}

static inline uint32_t rans_get(uint32_t st, uint32_t lo, uint32_t hi) {       /* STATEUPD */
    return (hi - lo) * (st >> PROB_BITS) + (st & (PROB_TOTAL - 1)) - lo;
}

static inline uint32_t rans_refill(uint32_t st, const uint8_t **ip) {          /* ecdnorm */
    if (st < ANS_L) { st = (st << 16) | ld16(*ip); *ip += 2; }
    return st;
}

/* ------------------------------------------------------------------------------------------
 * S1: anscdf4senc (anscdf.c:57-73): static rANS, 2 states, whole buffer.  The reference's
 * in-loop guards compare the output cursor with the *input* pointer (a bug, SURVEY finding 6a);
 * only the final size check is meaningful and only it is restated.  Works for any alphabet the
 * caller's cdf covers (the encoder just indexes cdf[x], cdf[x+1]).
 * `out` must have room for inlen + 16 bytes *below* out+inlen being addressable is not needed:
 * we encode into a private scratch buffer and copy.
 * ------------------------------------------------------------------------------------------ */
size_t orc_anscdf4senc(const uint8_t *in, size_t inlen, uint8_t *out, const cdf_t *cdf) {
    size_t cap = 2 * inlen + 64;
    uint8_t *buf = (uint8_t *)malloc(cap), *ep = buf + cap;
    uint32_t st[2] = { ANS_L, ANS_L };
    size_t i = inlen;
    while (i > (inlen & ~(size_t)3)) { i--; st[0] = rans_put(st[0], cdf[in[i]], cdf[in[i] + 1] - cdf[in[i]], &ep); }
    while (i > 0) {
        i--; st[1] = rans_put(st[1], cdf[in[i]], cdf[in[i] + 1] - cdf[in[i]], &ep);
        i--; st[0] = rans_put(st[0], cdf[in[i]], cdf[in[i] + 1] - cdf[in[i]], &ep);
        i--; st[1] = rans_put(st[1], cdf[in[i]], cdf[in[i] + 1] - cdf[in[i]], &ep);
        i--; st[0] = rans_put(st[0], cdf[in[i]], cdf[in[i] + 1] - cdf[in[i]], &ep);
    }
    ep -= 4; st32(ep, st[0]);                                 /* ansflush anscdf_.h:102 */
    ep -= 4; st32(ep, st[1]);
    size_t l = (size_t)(buf + cap - ep);
    if (l >= inlen) { memcpy(out, in, inlen); l = inlen; }    /* anscdf.c:70 */
    else memcpy(out, ep, l);
    free(buf);
    return l;
}

/* S2: anscdf4sdec (anscdf.c:75-85).  NOTE (reference bug, verified against the compiled
 * reference): the encoder puts the inlen&3 tail bytes on encoder state 0 (= decoder state 1,
 * because ansflush/mnfill reverse the state order) but the decoder takes them from decoder
 * state 0, so the reference does not round-trip when outlen % 4 != 0.  This function restates
 * the reference decoder as it is. */
size_t orc_anscdf4sdec(const uint8_t *in, size_t outlen, uint8_t *out, const cdf_t *cdf) {
    const uint8_t *ip = in;
    uint32_t st[2];
    st[0] = ld32(ip); ip += 4; st[1] = ld32(ip); ip += 4;     /* mnfill anscdf_.h:176 */
    size_t o = 0;
#define SDEC(_s_) do { unsigned x = rans_find16(cdf, st[_s_] & (PROB_TOTAL - 1)); \
        st[_s_] = rans_get(st[_s_], cdf[x], cdf[x + 1]); st[_s_] = rans_refill(st[_s_], &ip); out[o++] = (uint8_t)x; } while (0)
    while (o < (outlen & ~(size_t)3)) { SDEC(1); SDEC(0); SDEC(1); SDEC(0); }
    while (o < outlen) SDEC(0);
#undef SDEC
    return outlen;
}

/* Static rANS decoder for an n-symbol alphabet (n <= 256): the true inverse of orc_anscdf4senc
 * when the caller's cdf has n+1 entries (tail bytes come from decoder state 1, see above).
 * NO reference counterpart exists (the reference decoder only searches 16 entries,
 * cdf_.h:61-66); it is the same arithmetic with the search widened and the tail state fixed.
 * For outlen % 4 == 0 and n <= 16 it equals orc_anscdf4sdec. */
size_t orc_ans_sdec_n(const uint8_t *in, size_t outlen, uint8_t *out, const cdf_t *cdf, unsigned n) {
    const uint8_t *ip = in;
    uint32_t st[2];
    st[0] = ld32(ip); ip += 4; st[1] = ld32(ip); ip += 4;
    size_t o = 0;
#define SDEC(_s_) do { uint32_t r = st[_s_] & (PROB_TOTAL - 1); unsigned x = 0; \
        while (x + 1 < n && cdf[x + 1] <= r) x++; \
        st[_s_] = rans_get(st[_s_], cdf[x], cdf[x + 1]); st[_s_] = rans_refill(st[_s_], &ip); out[o++] = (uint8_t)x; } while (0)
    while (o < (outlen & ~(size_t)3)) { SDEC(1); SDEC(0); SDEC(1); SDEC(0); }
    while (o < outlen) SDEC(1);
#undef SDEC
    return outlen;
}

/* ------------------------------------------------------------------------------------------
 * Adaptive rANS family (A1 anscdf4enc/dec anscdf.c:87-133, A2/A3 anscdfenc/dec anscdf.c:567-605,
 * A4 anscdf1enc/dec anscdf.c:607-645).  Model pass pushes (state id, cum, freq) records
 * (mnenc4 anscdf_.h:106); coding pass pops them in reverse (mnflush anscdf_.h:128-138).
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t *base, *top; } recstack;

static inline void model_push(recstack *s, cdf_t *t, unsigned si, unsigned x) {
    *s->top++ = (uint32_t)si << 30 | (uint32_t)t[x] << 15 | (uint32_t)(t[x + 1] - t[x]);
    adapt_update(t, x);
}

/* mnflush: returns 0 on success (block appended at *op), 1 on overflow (caller emits raw copy).
 * out_end = out + inlen of the WHOLE call (every block starts its LIFO there, anscdf_.h:130). */
static int block_flush(recstack *s, unsigned nstates, uint8_t **op, uint8_t *out_end) {
    uint32_t st[4] = { ANS_L, ANS_L, ANS_L, ANS_L };
    uint8_t *ep = out_end;
    while (s->top != s->base) {
        uint32_t rec = *--s->top;
        if (ep <= *op + 2 + 4 * nstates) return 1;
        unsigned si = rec >> 30;
        st[si] = rans_put(st[si], (rec >> 15) & 0x7fff, rec & 0x7fff, &ep);
    }
    for (unsigned i = 0; i < nstates; i++) { ep -= 4; st32(ep, st[i]); }
    if (ep <= *op) return 1;
    size_t l = (size_t)(out_end - ep);
    if (*op + l >= out_end) return 1;
    memmove(*op, ep, l); *op += l;
    return 0;
}

/* The reference's overflow exit copies from the *advanced* input pointer (SURVEY finding 6d,
 * undefined past the first block); the restatement copies the original input. */
#define ADAPT_ENC_PROLOGUE \
    const uint8_t *in0 = in; uint8_t *op = out, *out_end = out + inlen; \
    size_t blk = inlen < ANS_BLOCK ? inlen : ANS_BLOCK; \
    recstack s; s.base = (uint32_t *)malloc((blk * 2 + 16) * sizeof(uint32_t)); s.top = s.base; \
    size_t pos = 0;
#define ADAPT_ENC_RAW { memcpy(out, in0, inlen); free(s.base); return inlen; }

size_t orc_anscdf4enc(const uint8_t *in, size_t inlen, uint8_t *out) {          /* A1 */
    ADAPT_ENC_PROLOGUE
    while (pos < inlen) {
        cdf_t t[17]; adapt_init(t);
        size_t n = inlen - pos < blk ? inlen - pos : blk, i = 0;
        s.top = s.base;
        for (; i < (n & ~(size_t)3); i += 4) {
            model_push(&s, t, 1, in[pos + i]);     model_push(&s, t, 0, in[pos + i + 1]);
            model_push(&s, t, 1, in[pos + i + 2]); model_push(&s, t, 0, in[pos + i + 3]);
        }
        for (; i < n; i++) model_push(&s, t, 0, in[pos + i]);
        if (block_flush(&s, 2, &op, out_end)) ADAPT_ENC_RAW
        pos += n;
    }
    free(s.base);
    return (size_t)(op - out);
}

/* anscdf4dec has the same tail-state bug as anscdf4sdec (encoder tail on state 0 = decoder
 * state 1, decoder reads decoder state 0).  tail_state 0 restates the reference, 1 is the true
 * inverse of orc_anscdf4enc. */
static size_t ans_nib_dec(const uint8_t *in, size_t outlen, uint8_t *out, int tail_state) {
    const uint8_t *ip = in;
    size_t blk = outlen < ANS_BLOCK ? outlen : ANS_BLOCK, pos = 0;
    while (pos < outlen) {
        cdf_t t[17]; adapt_init(t);
        uint32_t st[2];
        size_t n = outlen - pos < blk ? outlen - pos : blk, i = 0;
        st[0] = ld32(ip); ip += 4; st[1] = ld32(ip); ip += 4;
#define ADEC(_s_) do { unsigned x = rans_find16(t, st[_s_] & (PROB_TOTAL - 1)); \
            st[_s_] = rans_get(st[_s_], t[x], t[x + 1]); adapt_update(t, x); \
            st[_s_] = rans_refill(st[_s_], &ip); out[pos + i++] = (uint8_t)x; } while (0)
        while (i < (n & ~(size_t)3)) { ADEC(0); ADEC(1); ADEC(0); ADEC(1); }
        if (tail_state) while (i < n) ADEC(1); else while (i < n) ADEC(0);
#undef ADEC
        pos += n;
    }
    return outlen;
}
size_t orc_anscdf4dec    (const uint8_t *in, size_t outlen, uint8_t *out) { return ans_nib_dec(in, outlen, out, 0); }
size_t orc_anscdf4dec_fix(const uint8_t *in, size_t outlen, uint8_t *out) { return ans_nib_dec(in, outlen, out, 1); }

/* byte model = high-nibble table + 16 low-nibble tables (mnenc8x2 anscdf_.h:114-119); the
 * order-1 variant selects both by the previous byte (mnenc8x2x anscdf_.h:121-126). */
static size_t ans_byte_enc(const uint8_t *in, size_t inlen, uint8_t *out, int order1) {
    ADAPT_ENC_PROLOGUE
    size_t nctx = order1 ? 256 : 1;
    cdf_t (*th)[17] = (cdf_t (*)[17])malloc(nctx * 17 * sizeof(cdf_t));
    cdf_t (*tl)[16][17] = (cdf_t (*)[16][17])malloc(nctx * 16 * 17 * sizeof(cdf_t));
    unsigned cx = 0;                                  /* NOT reset per block (anscdf.c:608) */
    while (pos < inlen) {
        for (size_t c = 0; c < nctx; c++) { adapt_init(th[c]); for (int k = 0; k < 16; k++) adapt_init(tl[c][k]); }
        size_t n = inlen - pos < blk ? inlen - pos : blk;
        s.top = s.base;
        for (size_t i = 0; i < n; i += 2) {
            unsigned x0 = in[pos + i], x1 = i + 1 < n ? in[pos + i + 1] : 0;   /* odd tail: dummy 0 (anscdf.c:581) */
            unsigned c0 = order1 ? cx : 0;
            model_push(&s, th[c0], 3, x0 >> 4); model_push(&s, tl[c0][x0 >> 4], 2, x0 & 15);
            unsigned c1 = order1 ? x0 : 0;
            model_push(&s, th[c1], 1, x1 >> 4); model_push(&s, tl[c1][x1 >> 4], 0, x1 & 15);
            cx = x1;
        }
        if (block_flush(&s, 4, &op, out_end)) { free(th); free(tl); ADAPT_ENC_RAW }
        pos += n;
    }
    free(th); free(tl); free(s.base);
    return (size_t)(op - out);
}

static size_t ans_byte_dec(const uint8_t *in, size_t outlen, uint8_t *out, int order1) {
    const uint8_t *ip = in;
    size_t blk = outlen < ANS_BLOCK ? outlen : ANS_BLOCK, pos = 0;
    size_t nctx = order1 ? 256 : 1;
    cdf_t (*th)[17] = (cdf_t (*)[17])malloc(nctx * 17 * sizeof(cdf_t));
    cdf_t (*tl)[16][17] = (cdf_t (*)[16][17])malloc(nctx * 16 * 17 * sizeof(cdf_t));
    unsigned cx = 0;
    while (pos < outlen) {
        for (size_t c = 0; c < nctx; c++) { adapt_init(th[c]); for (int k = 0; k < 16; k++) adapt_init(tl[c][k]); }
        size_t n = outlen - pos < blk ? outlen - pos : blk;
        uint32_t st[4];
        for (int k = 0; k < 4; k++) { st[k] = ld32(ip); ip += 4; }
        for (size_t i = 0; i < n; i += 2) {           /* mndec8x2 / mndec8x2x anscdf_.h:152-174 */
            unsigned c0 = order1 ? cx : 0, yh, yl, x0, x1;
            yh = rans_find16(th[c0], st[0] & 0x7fff); st[0] = rans_get(st[0], th[c0][yh], th[c0][yh + 1]); adapt_update(th[c0], yh);
            cdf_t *m = tl[c0][yh];
            yl = rans_find16(m, st[1] & 0x7fff);      st[1] = rans_get(st[1], m[yl], m[yl + 1]);           adapt_update(m, yl);
            x0 = yh << 4 | yl;
            unsigned c1 = order1 ? x0 : 0;
            yh = rans_find16(th[c1], st[2] & 0x7fff); st[2] = rans_get(st[2], th[c1][yh], th[c1][yh + 1]); adapt_update(th[c1], yh);
            m = tl[c1][yh];
            yl = rans_find16(m, st[3] & 0x7fff);      st[3] = rans_get(st[3], m[yl], m[yl + 1]);           adapt_update(m, yl);
            x1 = yh << 4 | yl;
            cx = x1;
            for (int k = 0; k < 4; k++) st[k] = rans_refill(st[k], &ip);
            out[pos + i] = (uint8_t)x0;
            if (i + 1 < n) out[pos + i + 1] = (uint8_t)x1;
        }
        pos += n;
    }
    free(th); free(tl);
    return outlen;
}

size_t orc_anscdfenc (const uint8_t *in, size_t inlen,  uint8_t *out) { return ans_byte_enc(in, inlen,  out, 0); }  /* A2 */
size_t orc_anscdfdec (const uint8_t *in, size_t outlen, uint8_t *out) { return ans_byte_dec(in, outlen, out, 0); }  /* A3 */
size_t orc_anscdf1enc(const uint8_t *in, size_t inlen,  uint8_t *out) { return ans_byte_enc(in, inlen,  out, 1); }  /* A4 */
size_t orc_anscdf1dec(const uint8_t *in, size_t outlen, uint8_t *out) { return ans_byte_dec(in, outlen, out, 1); }

/* ------------------------------------------------------------------------------------------
 * C1: cdfini (rccdf.c:50-68).  Returns inlen, or -1 where the reference would die().
 * ------------------------------------------------------------------------------------------ */
int orc_cdfini(const uint8_t *in, size_t inlen, cdf_t *cdf, unsigned cdfnum) {
    size_t cnt[256] = { 0 }, mx = 0, mxi = 0, cum = 0;
    for (size_t i = 0; i < inlen; i++) cnt[in[i]]++;
    for (size_t i = 0; i < cdfnum; i++) {
        cnt[i] = (cnt[i] << PROB_BITS) / inlen;
        if (!cnt[i]) cnt[i] = 1;
        cum += cnt[i];
        if (cnt[i] > mx) { mx = cnt[i]; mxi = i; }
    }
    cnt[mxi] -= cum - PROB_TOTAL;
    cdf[0] = 0;
    for (size_t i = 0; i < cdfnum; i++) cdf[i + 1] = (cdf_t)(cdf[i] + cnt[i]);
    for (size_t i = 0; i < cdfnum; i++) if (cdf[i] >= cdf[i + 1]) return -1;
    if (cdf[cdfnum] != PROB_TOTAL) return -1;
    return (int)inlen;
}

/* ------------------------------------------------------------------------------------------
 * Range coder, canonical 64/32/15 format (R0: turborc_.h:52-58, rccdf.c:36-37).
 * Encoder R1/R2: turborc_.h:215,105-109,103,118-128.  Decoder R3/R4/R5: turborc_.h:152-158,
 * 224,243-321.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t range, low, ilow; uint8_t *op; } rcenc;
typedef struct { uint64_t range, code; const uint8_t *ip; } rcdec;

static void rce_init(rcenc *e, uint8_t *op) { e->low = e->ilow = 0; e->range = ~(uint64_t)0; e->op = op; }

static void rce_carry(rcenc *e, uint64_t newlow) {              /* _rccarry_ */
    if (e->ilow > newlow) { uint8_t *p = e->op; for (;;) { p -= 4; uint32_t w = ld32(p) + 1; st32(p, w); if (w) break; } }
}
static void rce_norm(rcenc *e) {                                /* _rcenorm_ */
    if (e->range < ((uint64_t)1 << 32)) {
        rce_carry(e, e->low);
        st32(e->op, (uint32_t)(e->low >> 32)); e->op += 4;
        e->low <<= 32; e->range <<= 32; e->ilow = e->low;
    }
}
static void rce_put(rcenc *e, uint32_t c0, uint32_t c1) {       /* _rccdfenc */
    e->range >>= PROB_BITS; e->low += e->range * c0; e->range *= (c1 - c0);
    rce_norm(e);
}
static void rce_flush(rcenc *e) {                               /* rceflush */
    rce_norm(e);
    if (e->range > ((uint64_t)1 << 33)) {
        e->low += (uint64_t)1 << 32; rce_carry(e, e->low);
        st32(e->op, (uint32_t)(e->low >> 32)); e->op += 4;
    } else {
        e->low += 1; rce_carry(e, e->low);
        st32(e->op, (uint32_t)(e->low >> 32)); e->op += 4;
        st32(e->op, (uint32_t)e->low); e->op += 4;
    }
}
static void rcd_init(rcdec *d, const uint8_t *ip) {             /* rcdinit */
    d->range = ~(uint64_t)0;
    d->code = (uint64_t)ld32(ip) << 32 | ld32(ip + 4);
    d->ip = ip + 8;
}
static void rcd_update(rcdec *d, uint32_t c0, uint32_t c1) {    /* _rccdfupdate */
    uint64_t rp = (uint64_t)c0 * d->range;
    d->range = d->range * c1 - rp; d->code -= rp;
    if (d->range < ((uint64_t)1 << 32)) { d->range <<= 32; d->code = d->code << 32 | ld32(d->ip); d->ip += 4; }
}
static unsigned rcd_bsearch(const rcdec *d, const cdf_t *cdf, unsigned cdfnum) {   /* _cdfbget */
    unsigned x = 0, high = cdfnum;
    while (x + 1 < high) { unsigned mid = (x + high) >> 1; if ((uint64_t)cdf[mid] * d->range > d->code) high = mid; else x = mid; }
    return x;
}
static unsigned rcd_lsearch16(const rcdec *d, const cdf_t *cdf) {                  /* _cdflget16 */
    unsigned x = 0;
    while (x < 15 && !((uint64_t)cdf[x + 1] * d->range > d->code)) x++;
    return x;
}
/* raw-copy threshold, OVERFLOW rcutil_.h:130: op >= out + inlen*255/256 - 8 (pointer compare) */
static int rc_overflow(const uint8_t *op, const uint8_t *out, size_t inlen) {
    return (ptrdiff_t)(op - out) >= (ptrdiff_t)((inlen * 255) / 256) - 8;
}

/* R6 */
size_t orc_rccdfsenc(const uint8_t *in, size_t inlen, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    (void)cdfnum; rcenc e; rce_init(&e, out);
    for (size_t i = 0; i < inlen; i++) {
        rce_put(&e, cdf[in[i]], cdf[in[i] + 1]);
        if (rc_overflow(e.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    }
    rce_flush(&e);
    return (size_t)(e.op - out);
}
size_t orc_rccdfsbdec(const uint8_t *in, size_t outlen, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    rcdec d; rcd_init(&d, in);
    for (size_t i = 0; i < outlen; i++) {
        d.range >>= PROB_BITS;
        unsigned x = rcd_bsearch(&d, cdf, cdfnum);
        rcd_update(&d, cdf[x], cdf[x + 1]); out[i] = (uint8_t)x;
    }
    return outlen;
}

/* R7: rccdfs2enc (rccdf.c:125-143).  Defined for inlen >= 4 (the reference computes
 * (inlen-4)*37/64 in size_t and walks off the buffer for smaller inputs). */
size_t orc_rccdfs2enc(const uint8_t *in, size_t inlen, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    (void)cdfnum;
    if (inlen < 4) { memcpy(out, in, inlen); return inlen; }
    uint8_t *base0 = out + 4, *base1 = out + 4 + (inlen - 4) * 37 / 64;
    rcenc e0, e1; rce_init(&e0, base0); rce_init(&e1, base1);
    size_t i = 0;
    for (; i < (inlen & ~(size_t)1); i += 2) {
        rce_put(&e0, cdf[in[i]], cdf[in[i] + 1]);
        rce_put(&e1, cdf[in[i + 1]], cdf[in[i + 1] + 1]);
        if (rc_overflow(e1.op, out, inlen) || e0.op >= base1) { memcpy(out, in, inlen); return inlen; }   /* OVERFLOWI rccdf.c:46 */
    }
    for (; i < inlen; i++) rce_put(&e0, cdf[in[i]], cdf[in[i] + 1]);
    rce_flush(&e0); rce_flush(&e1);
    st32(out, (uint32_t)(e0.op - base0));
    size_t l1 = (size_t)(e1.op - base1);
    memmove(e0.op, base1, l1); e0.op += l1;
    if (rc_overflow(e0.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    return (size_t)(e0.op - out);
}
/* R8: rccdfsb2dec (rccdf.c:166-184); rccdfsl2dec (:146-164) yields identical symbols for cdfnum <= 16 */
size_t orc_rccdfsb2dec(const uint8_t *in, size_t outlen, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    rcdec d0, d1; rcd_init(&d0, in + 4); rcd_init(&d1, in + 4 + ld32(in));
    size_t i = 0;
    for (; i < (outlen & ~(size_t)1); i += 2) {
        d0.range >>= PROB_BITS; d1.range >>= PROB_BITS;
        unsigned x0 = rcd_bsearch(&d0, cdf, cdfnum), x1 = rcd_bsearch(&d1, cdf, cdfnum);
        rcd_update(&d0, cdf[x0], cdf[x0 + 1]); rcd_update(&d1, cdf[x1], cdf[x1 + 1]);
        out[i] = (uint8_t)x0; out[i + 1] = (uint8_t)x1;
    }
    for (; i < outlen; i++) {
        d0.range >>= PROB_BITS;
        unsigned x = rcd_bsearch(&d0, cdf, cdfnum);
        rcd_update(&d0, cdf[x], cdf[x + 1]); out[i] = (uint8_t)x;
    }
    return outlen;
}

/* adaptive RC: one nibble = cdfenc + cdf16upd (cdf4e rccdf_.h:28) */
static void rce_nib(rcenc *e, cdf_t *t, unsigned x) { rce_put(e, t[x], t[x + 1]); adapt_update(t, x); }
static unsigned rcd_nib(rcdec *d, cdf_t *t) {                    /* cdf4d rccdf_.h:48 */
    d->range >>= PROB_BITS;
    unsigned x = rcd_lsearch16(d, t);
    rcd_update(d, t[x], t[x + 1]); adapt_update(t, x);
    return x;
}

/* R9: rccdfenc / rccdfdec (rccdf.c:187-211) */
size_t orc_rccdfenc(const uint8_t *in, size_t inlen, uint8_t *out) {
    cdf_t th[17], tl[16][17]; adapt_init(th); for (int k = 0; k < 16; k++) adapt_init(tl[k]);
    rcenc e; rce_init(&e, out);
    for (size_t i = 0; i < inlen; i++) {
        unsigned x = in[i];
        rce_nib(&e, th, x >> 4); rce_nib(&e, tl[x >> 4], x & 15);
        if (rc_overflow(e.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    }
    rce_flush(&e);
    return (size_t)(e.op - out);
}
size_t orc_rccdfdec(const uint8_t *in, size_t outlen, uint8_t *out) {
    cdf_t th[17], tl[16][17]; adapt_init(th); for (int k = 0; k < 16; k++) adapt_init(tl[k]);
    rcdec d; rcd_init(&d, in);
    for (size_t i = 0; i < outlen; i++) { unsigned h = rcd_nib(&d, th), l = rcd_nib(&d, tl[h]); out[i] = (uint8_t)(h << 4 | l); }
    return outlen;
}

/* R10: rccdfienc / rccdfidec (rccdf.c:213-249): high nibbles -> coder 0, low nibbles -> coder 1.
 * OVERFLOWI is evaluated once per 4 bytes (rccdf.c:240), not in the tail loop. */
size_t orc_rccdfienc(const uint8_t *in, size_t inlen, uint8_t *out) {
    cdf_t th[17], tl[16][17]; adapt_init(th); for (int k = 0; k < 16; k++) adapt_init(tl[k]);
    uint8_t *base0 = out + 4, *base1 = out + 4 + inlen / 2;
    rcenc e0, e1; rce_init(&e0, base0); rce_init(&e1, base1);
    size_t i = 0;
    for (; i < (inlen & ~(size_t)3); i += 4) {
        for (int k = 0; k < 4; k++) { unsigned x = in[i + k]; rce_nib(&e0, th, x >> 4); rce_nib(&e1, tl[x >> 4], x & 15); }
        if (rc_overflow(e1.op, out, inlen) || e0.op >= base1) { memcpy(out, in, inlen); return inlen; }
    }
    for (; i < inlen; i++) { unsigned x = in[i]; rce_nib(&e0, th, x >> 4); rce_nib(&e1, tl[x >> 4], x & 15); }
    rce_flush(&e0); rce_flush(&e1);
    st32(out, (uint32_t)(e0.op - base0));
    size_t l1 = (size_t)(e1.op - base1);
    memmove(e0.op, base1, l1); e0.op += l1;
    if (rc_overflow(e0.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    return (size_t)(e0.op - out);
}
size_t orc_rccdfidec(const uint8_t *in, size_t outlen, uint8_t *out) {          /* cdf8d2 rccdf_.h:63-73 */
    cdf_t th[17], tl[16][17]; adapt_init(th); for (int k = 0; k < 16; k++) adapt_init(tl[k]);
    rcdec d0, d1; rcd_init(&d0, in + 4); rcd_init(&d1, in + 4 + ld32(in));
    for (size_t i = 0; i < outlen; i++) { unsigned h = rcd_nib(&d0, th), l = rcd_nib(&d1, tl[h]); out[i] = (uint8_t)(h << 4 | l); }
    return outlen;
}

/* V8: "vnibble" byte code, rccdfenc8 / rccdfdec8 / rccdfienc8 / rccdfidec8 (rccdf.c:324-389) over cdfe8 / cdfd8
 * (rccdf_.h:76-96): x < 13 -> one symbol on table 0; x < 45 -> (x-13 >> 4) + 13 on table 0, low nibble on table 1;
 * else 15 on table 0, (x-45) >> 4 on table 1, low nibble on table 2.  The interleaved form puts the table-1 symbols
 * on coder 1 and everything else on coder 0. */
static void rce_v8(rcenc *e0, rcenc *e1, cdf_t *m0, cdf_t *m1, cdf_t *m2, unsigned x) {
    if (x < 13) rce_nib(e0, m0, x);
    else if (x < 13 + 32) { x -= 13; rce_nib(e0, m0, (x >> 4) + 13); rce_nib(e1, m1, x & 15); }
    else { x -= 13 + 32; rce_nib(e0, m0, 15); rce_nib(e1, m1, x >> 4); rce_nib(e0, m2, x & 15); }
}
static unsigned rcd_v8(rcdec *d0, rcdec *d1, cdf_t *m0, cdf_t *m1, cdf_t *m2) {
    unsigned x = rcd_nib(d0, m0);
    if (x >= 13) {
        unsigned y = rcd_nib(d1, m1);
        if (x != 15) x = ((x - 13) << 4 | y) + 13;
        else { x = rcd_nib(d0, m2); x = (y << 4 | x) + 13 + 32; }
    }
    return x;
}
size_t orc_rccdfenc8(const uint8_t *in, size_t inlen, uint8_t *out) {                 /* rccdf.c:341-351 */
    cdf_t m0[17], m1[17], m2[17]; adapt_init(m0); adapt_init(m1); adapt_init(m2);
    rcenc e; rce_init(&e, out);
    for (size_t i = 0; i < inlen; i++) {
        rce_v8(&e, &e, m0, m1, m2, in[i]);
        if (rc_overflow(e.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    }
    rce_flush(&e);
    return (size_t)(e.op - out);
}
size_t orc_rccdfdec8(const uint8_t *in, size_t outlen, uint8_t *out) {                /* rccdf.c:324-339 */
    cdf_t m0[17], m1[17], m2[17]; adapt_init(m0); adapt_init(m1); adapt_init(m2);
    rcdec d; rcd_init(&d, in);
    for (size_t i = 0; i < outlen; i++) out[i] = (uint8_t)rcd_v8(&d, &d, m0, m1, m2);
    return outlen;
}
size_t orc_rccdfienc8(const uint8_t *in, size_t inlen, uint8_t *out) {                /* rccdf.c:371-389 */
    cdf_t m0[17], m1[17], m2[17]; adapt_init(m0); adapt_init(m1); adapt_init(m2);
    uint8_t *base0 = out + 4, *base1 = out + 4 + inlen * 37 / 64;
    rcenc e0, e1; rce_init(&e0, base0); rce_init(&e1, base1);
    size_t i = 0;
    for (; i < (inlen & ~(size_t)3); i += 4) {
        for (int k = 0; k < 4; k++) rce_v8(&e0, &e1, m0, m1, m2, in[i + k]);
        if (rc_overflow(e1.op, out, inlen) || e0.op >= base1) { memcpy(out, in, inlen); return inlen; }   /* OVERFLOWI rccdf.c:46 */
    }
    for (; i < inlen; i++) rce_v8(&e0, &e1, m0, m1, m2, in[i]);
    rce_flush(&e0); rce_flush(&e1);
    st32(out, (uint32_t)(e0.op - base0));
    size_t l1 = (size_t)(e1.op - base1);
    memmove(e0.op, base1, l1); e0.op += l1;
    if (rc_overflow(e0.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    return (size_t)(e0.op - out);
}
size_t orc_rccdfidec8(const uint8_t *in, size_t outlen, uint8_t *out) {               /* rccdf.c:354-369 */
    cdf_t m0[17], m1[17], m2[17]; adapt_init(m0); adapt_init(m1); adapt_init(m2);
    rcdec d0, d1; rcd_init(&d0, in + 4); rcd_init(&d1, in + 4 + ld32(in));
    for (size_t i = 0; i < outlen; i++) out[i] = (uint8_t)rcd_v8(&d0, &d1, m0, m1, m2);
    return outlen;
}

/* ------------------------------------------------------------------------------------------
 * VLC-over-CDF integer codecs: anscdf{u,uz,v,vz}{enc,dec}16, anscdf{v,vz}{enc,dec}32
 * (anscdf.c:139-483) and rccdf{v,vz,u}{enc,dec}{16,32} (rccdf.c:392-632).  Every 16/32-bit integer (optionally the zigzag
 * of its delta to the previous one) goes through Turbo VLC (include_/vlcbit.h:24-63): values below 2^(vn+1) are their own
 * symbol, larger ones become an exponent symbol (6 or 7 bits) plus mb mantissa bits written to a bit stream that grows
 * DOWNWARD from the end of the output (bit IO rcutil_.h:163-192).  The symbol is coded with two adaptive 16-entry tables
 * (cdfenc6 / cdfenc7 anscdf_.h:206-230, cdfe6 / cdfe7 rccdf_.h:100-122) by the 2-state blocked rANS or one range coder.
 * Stream: [u32 total length][entropy-coded part][bit stream], the decoder starts its bit reader at in + u32.
 * ------------------------------------------------------------------------------------------ */
static uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static void st64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }
typedef struct { uint64_t bw; unsigned br; uint8_t *p; } bitw;                 /* biteinir / bitput / bitenormr / bitflushr */
static void bw_init(bitw *b, uint8_t *end) { b->bw = 0; b->br = 64; b->p = end - 8; }
static void bw_put(bitw *b, unsigned nb, uint32_t x) { b->br -= nb; b->bw |= (uint64_t)x << b->br; }
static void bw_norm(bitw *b) { st64(b->p, b->bw); unsigned k = (64 - b->br) & ~7u; b->p -= k >> 3; b->bw <<= k; b->br += k; }
static void bw_flush(bitw *b) { st64(b->p, b->bw); unsigned k = (64 + 7 - b->br) & ~7u; b->p -= k >> 3; b->p += 8; }
typedef struct { uint64_t bw; unsigned br; const uint8_t *p; } bitr;           /* bitdinir / bitdnormr / bitpeek / bitrmv */
static void br_init(bitr *b, const uint8_t *end) { b->bw = 0; b->br = 0; b->p = end - 8; }
static void br_norm(bitr *b) { b->p -= b->br >> 3; b->bw = ld64(b->p); b->br &= 7; }
static unsigned bsr32(uint32_t x) { unsigned r = 0; while (x >>= 1) r++; return r; }
/* bitvrput with vb = 0 (vlcbit.h:40-48): mantissa to the bit stream, returns the symbol to entropy-code */
static unsigned vlc_put(bitw *b, unsigned vn, uint32_t x) {
    if (x >= (1u << (vn + 1))) {
        unsigned f = bsr32(x) - vn, expo = ((f + 1) << vn) + ((x >> f) & ((1u << vn) - 1)), mb = (expo >> vn) - 1;
        bw_put(b, mb, x & ((1u << mb) - 1)); bw_norm(b);
        x = expo;
    }
    return x;
}
static uint32_t vlc_get(bitr *b, unsigned vn, uint32_t x) {                    /* bitvrget vlcbit.h:59-64 */
    if (x >= (1u << (vn + 1))) {
        br_norm(b);
        unsigned mb = (x >> vn) - 1;
        uint32_t ma = (uint32_t)((b->bw << b->br) >> (64 - mb));
        x = (((1u << vn) + (x & ((1u << vn) - 1))) << mb) + ma;
        b->br += mb;
    }
    return x;
}
static uint32_t vlc_load(const uint8_t *in, size_t i, int w32) { if (w32) return ld32(in + 4 * i); return ld16(in + 2 * i); }
static uint32_t zz_enc(uint32_t cur, uint32_t prev, int w32) {                  /* zigzagenc16/32 of the difference (rcutil_.h:144-148) */
    if (w32) { int32_t d = (int32_t)(cur - prev); return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31); }
    int16_t d = (int16_t)(cur - prev); return (uint16_t)(((uint16_t)d << 1) ^ (uint16_t)(d >> 15));
}
static uint32_t zz_dec(uint32_t r, int w32) {
    if (w32) return (r >> 1) ^ (0u - (r & 1));
    uint16_t v = (uint16_t)r; return (uint16_t)((v >> 1) ^ (uint16_t)(0u - (v & 1)));
}
/* kind: vn = 1 -> 6-bit exponent, first symbol range 0..11 (cdfenc6); vn = 2 -> 7-bit exponent, 0..7 (cdfenc7) */
static size_t vlc_ans_enc(const uint8_t *in, size_t inbytes, uint8_t *out, int w32, unsigned vn, int zz) {
    const unsigned esz = w32 ? 4 : 2, lim = vn == 1 ? 12 : 8;
    size_t n = (inbytes + esz - 1) / esz, blk = n < ANS_BLOCK ? n : ANS_BLOCK, pos = 0;
    uint8_t *op = out + 4, *out_end = out + inbytes;
    recstack s; s.base = (uint32_t *)malloc((blk * 2 + 16) * sizeof(uint32_t)); s.top = s.base;
    bitw b; bw_init(&b, out_end);
    uint32_t cx = 0;
    while (pos < n) {
        cdf_t m0[17], m1[17]; adapt_init(m0); adapt_init(m1);
        size_t cnt = n - pos < blk ? n - pos : blk;
        s.top = s.base;
        for (size_t i = 0; i < cnt; i++) {
            uint32_t v = vlc_load(in, pos + i, w32), x = zz ? zz_enc(v, cx, w32) : v;
            cx = v;
            x = vlc_put(&b, vn, x);
            if (x < lim) model_push(&s, m0, 1, x);
            else { x -= lim; model_push(&s, m0, 1, (x >> 4) + lim); model_push(&s, m1, 0, x & 15); }
        }
        if (block_flush(&s, 2, &op, b.p - 8)) goto raw;                        /* mnflush(op, bp-8, ...) */
        pos += cnt;
    }
    bw_flush(&b);
    {
        size_t l = (size_t)(out_end - b.p);
        if (op + l >= out_end) goto raw;
        memmove(op, b.p, l); op += l;
        st32(out, (uint32_t)(op - out));
    }
    free(s.base);
    return (size_t)(op - out);
raw:
    memcpy(out, in, inbytes); free(s.base);
    return inbytes;
}
static size_t vlc_ans_dec(const uint8_t *in, size_t outbytes, uint8_t *out, int w32, unsigned vn, int zz) {
    const unsigned esz = w32 ? 4 : 2, lim = vn == 1 ? 12 : 8;
    size_t n = (outbytes + esz - 1) / esz, blk = n < ANS_BLOCK ? n : ANS_BLOCK, pos = 0;
    const uint8_t *ip = in + 4;
    bitr b; br_init(&b, in + ld32(in));
    uint32_t cx = 0;
    while (pos < n) {
        cdf_t m0[17], m1[17]; adapt_init(m0); adapt_init(m1);
        uint32_t st[2];
        size_t cnt = n - pos < blk ? n - pos : blk;
        st[0] = ld32(ip); ip += 4; st[1] = ld32(ip); ip += 4;                   /* mnfill(st, ip, 2) */
        for (size_t i = 0; i < cnt; i++) {
#define VDEC(_s_, _m_, _x_) do { _x_ = rans_find16(_m_, st[_s_] & (PROB_TOTAL - 1)); st[_s_] = rans_get(st[_s_], _m_[_x_], _m_[_x_ + 1]); \
                                 adapt_update(_m_, _x_); st[_s_] = rans_refill(st[_s_], &ip); } while (0)
            unsigned x, y;
            VDEC(0, m0, x);
            if (x >= lim) { VDEC(1, m1, y); x = ((x - lim) << 4 | y) + lim; }
#undef VDEC
            uint32_t r = vlc_get(&b, vn, x);
            if (zz) { cx += zz_dec(r, w32); r = cx; }
            if (w32) st32(out + 4 * (pos + i), r); else st16(out + 2 * (pos + i), (uint16_t)r);
        }
        pos += cnt;
    }
    return outbytes;
}
static size_t vlc_rc_enc(const uint8_t *in, size_t inbytes, uint8_t *out, int w32, unsigned vn, int zz) {
    const unsigned esz = w32 ? 4 : 2, lim = vn == 1 ? 12 : 8;
    size_t n = (inbytes + esz - 1) / esz;
    uint8_t *out_end = out + inbytes;
    cdf_t m0[17], m1[17]; adapt_init(m0); adapt_init(m1);
    rcenc e; rce_init(&e, out + 4);
    bitw b; bw_init(&b, out_end);
    uint32_t cx = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t v = vlc_load(in, i, w32), x = zz ? zz_enc(v, cx, w32) : v;
        x = vlc_put(&b, vn, x);
        if (x < lim) rce_nib(&e, m0, x);
        else { x -= lim; rce_nib(&e, m0, (x >> 4) + lim); rce_nib(&e, m1, x & 15); }
        if (e.op + 8 >= b.p) { memcpy(out, in, inbytes); return inbytes; }      /* rccdf.c:406 */
        cx = v;
    }
    rce_flush(&e);
    bw_flush(&b);
    size_t l = (size_t)(out_end - b.p);
    memmove(e.op, b.p, l); e.op += l;
    st32(out, (uint32_t)(e.op - out));
    if (rc_overflow(e.op, out, inbytes)) { memcpy(out, in, inbytes); return inbytes; }
    return (size_t)(e.op - out);
}
static size_t vlc_rc_dec(const uint8_t *in, size_t outbytes, uint8_t *out, int w32, unsigned vn, int zz) {
    const unsigned esz = w32 ? 4 : 2, lim = vn == 1 ? 12 : 8;
    size_t n = (outbytes + esz - 1) / esz;
    cdf_t m0[17], m1[17]; adapt_init(m0); adapt_init(m1);
    rcdec d; rcd_init(&d, in + 4);
    bitr b; br_init(&b, in + ld32(in));
    uint32_t cx = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned x = rcd_nib(&d, m0);
        if (x >= lim) { unsigned y = rcd_nib(&d, m1); x = ((x - lim) << 4 | y) + lim; }
        uint32_t r = vlc_get(&b, vn, x);
        if (zz) { cx += zz_dec(r, w32); r = cx; }
        if (w32) st32(out + 4 * i, r); else st16(out + 2 * i, (uint16_t)r);
    }
    return outbytes;
}
#define VLC_PAIR(_name_, _fn_, _w32_, _vn_, _zz_) \
    size_t orc_##_name_##enc##_w32_(const uint8_t *in, size_t n, uint8_t *out) { return _fn_##_enc(in, n, out, _w32_ == 32, _vn_, _zz_); } \
    size_t orc_##_name_##dec##_w32_(const uint8_t *in, size_t n, uint8_t *out) { return _fn_##_dec(in, n, out, _w32_ == 32, _vn_, _zz_); }
VLC_PAIR(anscdfu,  vlc_ans, 16, 1, 0)   /* anscdf.c:139-193 */
VLC_PAIR(anscdfuz, vlc_ans, 16, 1, 1)   /* anscdf.c:195-252 */
VLC_PAIR(anscdfv,  vlc_ans, 16, 2, 0)   /* anscdf.c:255-309 */
VLC_PAIR(anscdfvz, vlc_ans, 16, 2, 1)   /* anscdf.c:311-367 */
VLC_PAIR(anscdfv,  vlc_ans, 32, 2, 0)   /* anscdf.c:369-423 */
VLC_PAIR(anscdfvz, vlc_ans, 32, 2, 1)   /* anscdf.c:425-483 */
VLC_PAIR(rccdfv,   vlc_rc,  16, 2, 0)   /* rccdf.c:392-429 */
VLC_PAIR(rccdfvz,  vlc_rc,  16, 2, 1)   /* rccdf.c:432-470 */
VLC_PAIR(rccdfv,   vlc_rc,  32, 2, 0)   /* rccdf.c:473-512 */
VLC_PAIR(rccdfvz,  vlc_rc,  32, 2, 1)   /* rccdf.c:515-553 */
VLC_PAIR(rccdfu,   vlc_rc,  16, 1, 0)   /* rccdf.c:555-592 */
VLC_PAIR(rccdfu,   vlc_rc,  32, 1, 0)   /* rccdf.c:595-632 */

/* R11: nibble-alphabet adaptive RC (rccdf.c:251-323) */
size_t orc_rccdf4enc(const uint8_t *in, size_t inlen, uint8_t *out) {
    cdf_t t[17]; adapt_init(t);
    rcenc e; rce_init(&e, out);
    for (size_t i = 0; i < inlen; i++) {
        rce_nib(&e, t, in[i]);
        if (rc_overflow(e.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    }
    rce_flush(&e);
    return (size_t)(e.op - out);
}
size_t orc_rccdf4dec(const uint8_t *in, size_t outlen, uint8_t *out) {
    cdf_t t[17]; adapt_init(t);
    rcdec d; rcd_init(&d, in);
    for (size_t i = 0; i < outlen; i++) out[i] = (uint8_t)rcd_nib(&d, t);
    return outlen;
}
/* rccdf4ienc (rccdf.c:302-323): bytes alternate between the two coders; both symbols of a pair
 * are coded against the same table state, then both updates are applied (:311-314).  The
 * per-pair overflow test is the plain OVERFLOW on stream 1 only (:314). */
size_t orc_rccdf4ienc(const uint8_t *in, size_t inlen, uint8_t *out) {
    cdf_t t[17]; adapt_init(t);
    uint8_t *base0 = out + 4, *base1 = out + 4 + inlen / 2;
    rcenc e0, e1; rce_init(&e0, base0); rce_init(&e1, base1);
    size_t i = 0;
    for (; i < (inlen & ~(size_t)1); i += 2) {
        unsigned x0 = in[i], x1 = in[i + 1];
        rce_put(&e0, t[x0], t[x0 + 1]); rce_put(&e1, t[x1], t[x1 + 1]);
        adapt_update(t, x0); adapt_update(t, x1);
        /* reference quirk: OVERFLOW sets op1 but the function returns op0 - out (rccdf.c:314,321-322) */
        if (rc_overflow(e1.op, out, inlen)) { memcpy(out, in, inlen); return (size_t)(e0.op - out); }
    }
    for (; i < inlen; i++) rce_nib(&e0, t, in[i]);
    rce_flush(&e0); rce_flush(&e1);
    st32(out, (uint32_t)(e0.op - base0));
    size_t l1 = (size_t)(e1.op - base1);
    memmove(e0.op, base1, l1); e0.op += l1;
    if (rc_overflow(e0.op, out, inlen)) { memcpy(out, in, inlen); return inlen; }
    return (size_t)(e0.op - out);
}
size_t orc_rccdf4idec(const uint8_t *in, size_t outlen, uint8_t *out) {          /* rccdf.c:280-300 */
    cdf_t t[17]; adapt_init(t);
    rcdec d0, d1; rcd_init(&d0, in + 4); rcd_init(&d1, in + 4 + ld32(in));
    size_t i = 0;
    for (; i < (outlen & ~(size_t)1); i += 2) {
        d0.range >>= PROB_BITS; d1.range >>= PROB_BITS;
        unsigned x0 = rcd_lsearch16(&d0, t), x1 = rcd_lsearch16(&d1, t);
        rcd_update(&d0, t[x0], t[x0 + 1]); rcd_update(&d1, t[x1], t[x1 + 1]);
        adapt_update(t, x0); adapt_update(t, x1);
        out[i] = (uint8_t)x0; out[i + 1] = (uint8_t)x1;
    }
    for (; i < outlen; i++) out[i] = (uint8_t)rcd_nib(&d0, t);
    return outlen;
}

/* ------------------------------------------------------------------------------------------
 * TRC_ANSW: 32-way interleaved static rANS, a NEW stream format of this repository (SURVEY.md
 * section 8c "wide static byte rANS").  PARITY UNPINNED: no reference codec produces or reads
 * this layout, so this restatement is the format's specification, not a reference check; the
 * per-symbol arithmetic is the reference's (ece anscdf_.h:90-94, STATEUPD cdf_.h:37, ecdnorm
 * anscdf_.h:50-73, 15-bit CDF, 16-bit words).  Acceptance = exact round trip + size bound.
 *
 * Layout of one call: [32 x u32 states, state 0 first][u16 words ...].  Symbol i belongs to state
 * (i / 4) % 32: a 128-symbol super-group gives each state 4 consecutive symbols (one 32-bit word
 * per GPU lane).  Decoder: for every super-group, for k = 0..3: each state decodes symbol
 * 128g + 4s + k (if < n); then the states that dropped below 2^15 refill one word each, in
 * increasing state order.  The encoder runs the exact reverse and writes words downwards.
 * Raw rule: a stream that is not shorter than the input is replaced by a raw copy (length == n).
 * ------------------------------------------------------------------------------------------ */
size_t orc_answenc(const uint8_t *in, size_t n, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    (void)cdfnum;
    size_t cap = 2 * n + 256;
    uint8_t *buf = (uint8_t *)malloc(cap), *ep = buf + cap;
    uint32_t st[32];
    for (int s = 0; s < 32; s++) st[s] = ANS_L;
    size_t ng = (n + 127) / 128;
    for (size_t g = ng; g-- > 0;)
        for (int k = 3; k >= 0; k--)
            for (int s = 31; s >= 0; s--) {                    /* words of one step end up in increasing state order */
                size_t i = g * 128 + (size_t)s * 4 + (size_t)k;
                if (i >= n) continue;
                unsigned x = in[i];
                st[s] = rans_put(st[s], cdf[x], cdf[x + 1] - cdf[x], &ep);
            }
    for (int s = 31; s >= 0; s--) { ep -= 4; st32(ep, st[s]); }
    size_t l = (size_t)(buf + cap - ep);
    if (l >= n) { memcpy(out, in, n); l = n; } else memcpy(out, ep, l);
    free(buf);
    return l;
}

size_t orc_answdec(const uint8_t *in, size_t n, uint8_t *out, const cdf_t *cdf, unsigned cdfnum) {
    const uint8_t *ip = in;
    uint32_t st[32];
    for (int s = 0; s < 32; s++) { st[s] = ld32(ip); ip += 4; }
    size_t ng = (n + 127) / 128;
    for (size_t g = 0; g < ng; g++)
        for (int k = 0; k < 4; k++) {
            for (int s = 0; s < 32; s++) {
                size_t i = g * 128 + (size_t)s * 4 + (size_t)k;
                if (i >= n) continue;
                uint32_t r = st[s] & (PROB_TOTAL - 1); unsigned x = 0;
                while (x + 1 < cdfnum && cdf[x + 1] <= r) x++;
                st[s] = rans_get(st[s], cdf[x], cdf[x + 1]);
                out[i] = (uint8_t)x;
            }
            for (int s = 0; s < 32; s++) {
                size_t i = g * 128 + (size_t)s * 4 + (size_t)k;
                if (i >= n) continue;
                st[s] = rans_refill(st[s], &ip);
            }
        }
    return n;
}
