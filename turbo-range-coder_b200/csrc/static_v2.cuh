// static_v2.cuh -- throughput versions of the static-CDF codecs for regular batches (every call starts on a
// 16-byte boundary): TRC_RCS / TRC_RCS2 range coders and the TRC_ANS4S rANS coder.  Same bytes as the generic
// kernels in rans_static.cuh / rc_static.cuh (which still serve ragged tails and per-call tables), built for
// issue rate:
//   * symbol tables are built ONCE per launch by k_build_tables into global memory and pulled into shared
//     memory by one TMA bulk copy per CTA (cp.async.bulk + mbarrier -> SASS UBLKCP), instead of being
//     recomputed by every CTA;
//   * each lane (one call: both range coders / both rANS states, so two independent dependency chains per
//     lane) streams its chunk through 128-bit loads with the next 16 bytes prefetched into registers, looks
//     all 16 table entries up front, then runs the 16 coding steps back to back;
//   * the range decoder finds the symbol with one fp32 reciprocal estimate of code/range + a 32 K-entry
//     slot->symbol LUT and an exact +-1 fix-up (identical result to the reference's 8-step binary search,
//     turborc_.h:307-315), and keeps the next stream word prefetched in a register;
//   * decoded bytes leave as 128-bit stores.
#pragma once
#include "trc_common.cuh"
#include "rans_static.cuh"
#include "rc_static.cuh"

namespace trc {

// ---- table set (one per cdf table), built by k_build_tables -------------------------------------------
struct __align__(16) TableSet {
    uint4    etab[256];          // rANS encoder entries (rans_enc_entry)
    uint32_t ctab[256];          // cdf | freq << 16      (RC encoder)
    uint32_t dtab[256];          // freq | cdf << 16      (rANS decoder; RC decoder reads cdf = >>16, freq = &0xffff)
    uint8_t  lut[PROB_TOTAL];    // slot r -> max{ x < cdfnum : cdf[x] <= r }
};
static_assert(sizeof(TableSet) % 16 == 0, "bulk copies move multiples of 16 bytes");

__global__ void __launch_bounds__(1024)
k_build_tables(const cdf_t *__restrict__ cdf, unsigned cdfnum, TableSet *__restrict__ ts) {
    __shared__ uint16_t scdf[CDF_STRIDE];
    const cdf_t *c0 = cdf + (size_t)blockIdx.x * CDF_STRIDE;
    TableSet &t = ts[blockIdx.x];
    for (unsigned x = threadIdx.x; x <= cdfnum; x += blockDim.x) scdf[x] = c0[x];
    __syncthreads();
    for (unsigned x = threadIdx.x; x < 256; x += blockDim.x) {
        uint32_t c = 0, f = 0;
        if (x < cdfnum) { c = scdf[x]; f = (uint32_t)scdf[x + 1] - c; }
        t.etab[x] = x < cdfnum ? rans_enc_entry(c, f) : make_uint4(0, 0, 0, 0);
        t.ctab[x] = c | f << 16;
        t.dtab[x] = (f & 0xffffu) | c << 16;
    }
    for (unsigned r = threadIdx.x; r < PROB_TOTAL; r += blockDim.x) {
        unsigned x = 0, hi = cdfnum;
        while (x + 1 < hi) { unsigned mid = (x + hi) >> 1; if (scdf[mid] <= r) x = mid; else hi = mid; }
        t.lut[r] = (uint8_t)x;
    }
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread: init barrier, arm it with the byte count, launch the copy
__device__ __forceinline__ void tma_fetch(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    uint32_t b = smem_u32(bar), d = smem_u32(smem_dst);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void tma_wait(uint64_t *bar) {
    uint32_t b = smem_u32(bar), done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(b) : "memory");
    }
}

__device__ __forceinline__ uint4 ldg128(const uint8_t *p) { return __ldg((const uint4 *)p); }

constexpr int V2_NT = 128;

// =========================================================================================================
// Range coder encoder, NC coders per call (rccdfsenc rccdf.c:71-82 / rccdfs2enc rccdf.c:125-143)
// The reference evaluates its overflow test after every symbol (pair); the tested quantities only grow, so
// testing once per 16-byte block and once after the last symbol decides the same way.
// =========================================================================================================
template <int NC>
__global__ void __launch_bounds__(V2_NT)
k_rc_static_enc_v2(const uint8_t *__restrict__ in, Geom g, size_t n_calls, const TableSet *__restrict__ ts, size_t cpc,
                   uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ __align__(16) uint32_t ctab[256];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * V2_NT, j = j0 + threadIdx.x;
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) tma_fetch(ctab, t->ctab, sizeof ctab, &bar);
    __syncthreads();
    tma_wait(&bar);
    if (j >= n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + j * slot_stride;
    const int64_t thr = rc_thr(n);
    UnitMeta m; m.pref = 0; m.pad = 0; m.b_off = 0; m.b_len = 0; m.a_off = 0;
    bool raw = false;
    const size_t nb = n & ~(size_t)15;
    if (NC == 1) {
        RcEnc e; e.init(slot);
        uint4 cur = nb ? ldg128(ip) : make_uint4(0, 0, 0, 0);
        for (size_t i = 0; i < nb && !raw; i += 16) {
            uint4 nxt = i + 32 <= nb ? ldg128(ip + i + 16) : cur;
            const uint32_t w[4] = { cur.x, cur.y, cur.z, cur.w };
            uint32_t tt[16];
#pragma unroll
            for (int k = 0; k < 16; k++) tt[k] = ctab[(w[k >> 2] >> (8 * (k & 3))) & 0xff];
#pragma unroll
            for (int k = 0; k < 16; k++) e.encode(tt[k] & 0xffffu, tt[k] >> 16);
            raw = (int64_t)e.pos >= thr;                                              // OVERFLOW rccdf.c:77
            cur = nxt;
        }
        for (size_t i = nb; i < n && !raw; i++) { uint32_t tk = ctab[ip[i]]; e.encode(tk & 0xffffu, tk >> 16); raw = (int64_t)e.pos >= thr; }
        if (!raw) e.flush();
        m.a_len = raw ? 0 : e.pos; m.len = raw ? (uint32_t)n : e.pos; m.flags = raw ? UM_RAW : 0;
    } else {
        if (n < 4) { m.a_len = 0; m.len = (uint32_t)n; m.flags = UM_RAW; meta[j] = m; return; }
        const uint32_t b1ref = 4 + (uint32_t)(((n - 4) * 37) / 64);                  // rccdf.c:126
        const uint32_t b1 = (b1ref + 64 + 15) & ~15u;
        RcEnc e0, e1; e0.init(slot + 4); e1.init(slot + b1);
        uint4 cur = nb ? ldg128(ip) : make_uint4(0, 0, 0, 0);
        for (size_t i = 0; i < nb && !raw; i += 16) {
            uint4 nxt = i + 32 <= nb ? ldg128(ip + i + 16) : cur;
            const uint32_t w[4] = { cur.x, cur.y, cur.z, cur.w };
            uint32_t tt[16];
#pragma unroll
            for (int k = 0; k < 16; k++) tt[k] = ctab[(w[k >> 2] >> (8 * (k & 3))) & 0xff];
#pragma unroll
            for (int k = 0; k < 16; k += 2) { e0.encode(tt[k] & 0xffffu, tt[k] >> 16); e1.encode(tt[k + 1] & 0xffffu, tt[k + 1] >> 16); }
            raw = (int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref;             // OVERFLOWI rccdf.c:46,133
            cur = nxt;
        }
        size_t i = nb;
        for (; i + 2 <= n && !raw; i += 2) {
            uint32_t t0 = ctab[ip[i]], t1 = ctab[ip[i + 1]];
            e0.encode(t0 & 0xffffu, t0 >> 16); e1.encode(t1 & 0xffffu, t1 >> 16);
            raw = (int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref;
        }
        if (!raw) {
            if (i < n) { uint32_t tk = ctab[ip[i]]; e0.encode(tk & 0xffffu, tk >> 16); }
            e0.flush(); e1.flush();
            *(uint32_t *)slot = e0.pos;                                              // rccdf.c:141
            if ((int64_t)(4 + e0.pos + e1.pos) >= thr) raw = true;                   // rccdf.c:142
        }
        m.a_len = raw ? 0 : 4 + e0.pos; m.b_off = b1; m.b_len = raw ? 0 : e1.pos;
        m.len = raw ? (uint32_t)n : 4 + e0.pos + e1.pos; m.flags = raw ? UM_RAW : 0;
    }
    meta[j] = m;
}

// =========================================================================================================
// Range coder decoder
// =========================================================================================================
struct RcDec2 {
    uint64_t range, code;
    const uint8_t *ip, *lim;        // lim = last address a full 32-bit word can be read from
    uint32_t nxt;                   // next stream word, already loaded
    __device__ __forceinline__ uint32_t fetch() { uint32_t v = ip <= lim ? ld_u32(ip) : ld_u32_clamped(ip, lim + 4); ip += 4; return v; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *gend) {   // rcdinit turborc_.h:152-158
        lim = gend - 4; ip = p; range = ~0ull;
        uint32_t a = fetch(), b = fetch();
        code = (uint64_t)a << 32 | b;
        nxt = fetch();
    }
    // one symbol: range >>= 15; x = max{ x : cdf[x]*range <= code } (== _cdfbget turborc_.h:307-315); _rccdfupdate
    __device__ __forceinline__ uint32_t decode(const uint8_t *lut, const uint32_t *dtab, unsigned cdfnum) {
        range >>= PROB_BITS;
        // estimate q = code / range within +-1 (fp32: relative error < 2^-21 on a quotient < 2^16)
        float qf = __ull2float_rz(code) * __frcp_rn(__ull2float_rn(range));
        uint32_t q = (uint32_t)fminf(qf, 32767.0f);
        uint32_t x = lut[q], e = dtab[x];
        uint64_t rp = (uint64_t)(e >> 16) * range;
        if (rp > code) {                                            // estimate one too high
            x--; e = dtab[x]; rp = (uint64_t)(e >> 16) * range;
        } else if (x + 1 < cdfnum) {
            uint64_t rn = rp + (uint64_t)(e & 0xffffu) * range;      // cdf[x+1] * range
            if (rn <= code) { x++; e = dtab[x]; rp = rn; }           // estimate one too low
        }
        range *= (e & 0xffffu); code -= rp;
        if ((uint32_t)(range >> 32) == 0) {                          // _rcdnorm_ turborc_.h:111
            range <<= 32; code = code << 32 | nxt;
            nxt = fetch();
        }
        return x;
    }
};

template <int NC>
__global__ void __launch_bounds__(V2_NT)
k_rc_static_dec_v2(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                   size_t n_calls, const TableSet *__restrict__ ts, unsigned cdfnum, size_t cpc) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * V2_NT, j = j0 + threadIdx.x;
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) {
        uint32_t b = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(sizeof dtab + sizeof lut)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dtab)), "l"(t->dtab), "r"((uint32_t)sizeof dtab), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(lut)), "l"(t->lut), "r"((uint32_t)sizeof lut), "r"(b) : "memory");
    }
    __syncthreads();
    tma_wait(&bar);
    if (j >= n_calls) return;
    size_t start, n; call_span(g, j, start, n);
    uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls], *stream = in + so;
    uint8_t *op = out + start;
    if (sl == n) { thread_copy(op, stream, n); return; }
    const size_t nb = n & ~(size_t)15;
    if (NC == 1) {
        RcDec2 d; d.init(stream, gend);
        for (size_t i = 0; i < nb; i += 16) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t a = d.decode(lut, dtab, cdfnum), b = d.decode(lut, dtab, cdfnum), c = d.decode(lut, dtab, cdfnum), e = d.decode(lut, dtab, cdfnum);
                w[k] = a | b << 8 | c << 16 | e << 24;
            }
            *(uint4 *)(op + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (size_t i = nb; i < n; i++) op[i] = (uint8_t)d.decode(lut, dtab, cdfnum);
    } else {
        uint32_t len0 = ld_u32_clamped(stream, gend);
        const uint8_t *p1 = stream + 4 + len0;
        if (p1 > gend || p1 < stream) p1 = gend;
        RcDec2 d0, d1; d0.init(stream + 4, gend); d1.init(p1, gend);
        for (size_t i = 0; i < nb; i += 16) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t a = d0.decode(lut, dtab, cdfnum), b = d1.decode(lut, dtab, cdfnum), c = d0.decode(lut, dtab, cdfnum), e = d1.decode(lut, dtab, cdfnum);
                w[k] = a | b << 8 | c << 16 | e << 24;
            }
            *(uint4 *)(op + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        size_t i = nb;
        for (; i + 2 <= n; i += 2) { op[i] = (uint8_t)d0.decode(lut, dtab, cdfnum); op[i + 1] = (uint8_t)d1.decode(lut, dtab, cdfnum); }
        if (i < n) op[i] = (uint8_t)d0.decode(lut, dtab, cdfnum);
    }
}

// =========================================================================================================
// Static rANS (anscdf4senc / anscdf4sdec): both states of a call in one lane
// =========================================================================================================
__global__ void __launch_bounds__(V2_NT)
k_rans_static_enc_v2(const uint8_t *__restrict__ in, Geom g, size_t n_calls, const TableSet *__restrict__ ts, size_t cpc,
                     uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ __align__(16) uint4 etab[256];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * V2_NT, j = j0 + threadIdx.x;
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) tma_fetch(etab, t->etab, sizeof etab, &bar);
    __syncthreads();
    tma_wait(&bar);
    if (j >= n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    const uint8_t *ip = in + start;
    const uint32_t n = (uint32_t)len;
    const int cap = (int)slot_stride;
    RansWriter w; w.init(slots + j * slot_stride, cap);
    uint32_t s0 = ANS_L, s1 = ANS_L;
    uint32_t i = n;
    const uint32_t n4 = n & ~3u, n16 = n & ~15u;
    bool ovf = false;
    while (i > n4) { i--; s0 = rans_enc_step(s0, etab[ip[i]], w); }                   // tail on state 0 (anscdf.c:62-64)
    while (i > n16) {                                                                 // groups of 4 down to a 16-byte boundary
        i -= 4;
        uint32_t v = *(const uint32_t *)(ip + i);
        s1 = rans_enc_step(s1, etab[v >> 24], w); s0 = rans_enc_step(s0, etab[(v >> 16) & 0xff], w);
        s1 = rans_enc_step(s1, etab[(v >> 8) & 0xff], w); s0 = rans_enc_step(s0, etab[v & 0xff], w);
    }
    uint4 cur = i ? ldg128(ip + i - 16) : make_uint4(0, 0, 0, 0);
    while (i > 0 && !ovf) {                                                           // anscdf.c:65-67, 16 symbols per trip
        i -= 16;
        uint4 nxt = i ? ldg128(ip + i - 16) : cur;
        const uint32_t wv[4] = { cur.x, cur.y, cur.z, cur.w };
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            uint32_t v = wv[k];
            uint4 ea = etab[v >> 24], eb = etab[(v >> 16) & 0xff], ec = etab[(v >> 8) & 0xff], ed = etab[v & 0xff];
            s1 = rans_enc_step(s1, ea, w); s0 = rans_enc_step(s0, eb, w);
            s1 = rans_enc_step(s1, ec, w); s0 = rans_enc_step(s0, ed, w);
        }
        ovf = (uint32_t)(cap - w.pos) + 8u >= n;                                      // l >= inlen already certain
        cur = nxt;
    }
    w.finish_words();
    w.put32_final(s0); w.put32_final(s1);
    uint32_t l = (uint32_t)(cap - w.pos);
    bool raw = ovf || l >= n;                                                         // anscdf.c:70
    UnitMeta m;
    m.len = raw ? n : l; m.a_off = (uint32_t)w.pos; m.a_len = raw ? 0 : l; m.b_off = 0; m.b_len = 0;
    m.flags = raw ? UM_RAW : 0; m.pref = 0; m.pad = 0;
    meta[j] = m;
}

struct RansReader2 {
    const uint8_t *ip, *lim;
    uint32_t nxt;                   // next 16-bit word, already loaded
    __device__ __forceinline__ uint32_t fetch16() { uint32_t v = ip <= lim ? ld_u16(ip) : ld_u16_clamped(ip, lim + 2); ip += 2; return v; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *gend, uint32_t &s0, uint32_t &s1) {
        s0 = ld_u32_clamped(p, gend); s1 = ld_u32_clamped(p + 4, gend);
        ip = p + 8; lim = gend - 2; nxt = fetch16();
    }
    __device__ __forceinline__ uint32_t step(uint32_t &s, const uint8_t *lut, const uint32_t *dtab) {
        uint32_t r = s & PROB_MASK, x = lut[r], e = dtab[x];
        s = (e & 0xffffu) * (s >> PROB_BITS) + r - (e >> 16);                         // STATEUPD cdf_.h:37
        if (s < ANS_L) { s = s << 16 | nxt; nxt = fetch16(); }                        // ecdnorm anscdf_.h:50-73
        return x;
    }
};

__global__ void __launch_bounds__(V2_NT)
k_rans_static_dec_v2(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                     size_t n_calls, const TableSet *__restrict__ ts, size_t cpc, unsigned flags) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * V2_NT, j = j0 + threadIdx.x;
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) {
        uint32_t b = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(sizeof dtab + sizeof lut)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dtab)), "l"(t->dtab), "r"((uint32_t)sizeof dtab), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(lut)), "l"(t->lut), "r"((uint32_t)sizeof lut), "r"(b) : "memory");
    }
    __syncthreads();
    tma_wait(&bar);
    if (j >= n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    const uint8_t *gend = in + in_off[g.n_calls];
    uint8_t *op = out + start;
    const uint32_t n = (uint32_t)len;
    if (sl == len) { thread_copy(op, in + so, len); return; }
    RansReader2 rd; uint32_t s0, s1;
    rd.init(in + so, gend, s0, s1);                                                   // mnfill anscdf_.h:176
    const uint32_t n4 = n & ~3u, n16 = n & ~15u;
    uint32_t o = 0;
    for (; o < n16; o += 16) {                                                        // anscdf.c:82
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t a = rd.step(s1, lut, dtab), b = rd.step(s0, lut, dtab), c = rd.step(s1, lut, dtab), e = rd.step(s0, lut, dtab);
            wv[k] = a | b << 8 | c << 16 | e << 24;
        }
        *(uint4 *)(op + o) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (; o < n4; o += 4) {
        uint32_t a = rd.step(s1, lut, dtab), b = rd.step(s0, lut, dtab), c = rd.step(s1, lut, dtab), e = rd.step(s0, lut, dtab);
        *(uint32_t *)(op + o) = a | b << 8 | c << 16 | e << 24;
    }
    for (; o < n; o++) op[o] = (uint8_t)((flags & 1u) ? rd.step(s0, lut, dtab) : rd.step(s1, lut, dtab));   // anscdf.c:83
}

}  // namespace trc
