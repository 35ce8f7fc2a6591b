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

// grid = (tables, 1 + LUT_PARTS): y == 0 builds the three 256-entry tables, y >= 1 one slice of the slot->symbol LUT
// (decoders only: `with_lut`)
constexpr int LUT_PARTS = 8;
__global__ void __launch_bounds__(1024)
k_build_tables(const cdf_t *__restrict__ cdf, unsigned cdfnum, TableSet *__restrict__ ts, int with_lut,
               unsigned long long *__restrict__ lb_zero = nullptr, unsigned lb_n = 0) {
    __shared__ uint16_t scdf[CDF_STRIDE];
    if (lb_zero && blockIdx.x == 0 && blockIdx.y == 0) for (unsigned k = threadIdx.x; k < lb_n; k += blockDim.x) lb_zero[k] = 0;   // look-back words of k_rcs2_enc_fused
    const cdf_t *c0 = cdf + (size_t)blockIdx.x * CDF_STRIDE;
    TableSet &t = ts[blockIdx.x];
    if (blockIdx.y && !with_lut) return;
    for (unsigned x = threadIdx.x; x <= cdfnum; x += blockDim.x) scdf[x] = c0[x];
    __syncthreads();
    if (blockIdx.y == 0) {
        for (unsigned x = threadIdx.x; x < 256; x += blockDim.x) {
            uint32_t c = 0, f = 0;
            if (x < cdfnum) { c = scdf[x]; f = (uint32_t)scdf[x + 1] - c; }
            t.etab[x] = x < cdfnum ? rans_enc_entry(c, f) : make_uint4(0, 0, 0, 0);
            t.ctab[x] = c | f << 16;
            t.dtab[x] = (f & 0xffffu) | c << 16;
        }
        return;
    }
    const unsigned per = PROB_TOTAL / LUT_PARTS, r0 = (blockIdx.y - 1) * per;
    for (unsigned r = r0 + threadIdx.x; r < r0 + per; r += blockDim.x) {
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
constexpr int LPC_NT = 128;                       // lane-per-coder kernels: 64 calls per CTA by default
constexpr int LPC_MAX_NT = 512;                   // ... up to 256 calls per CTA when the batch is sized to one CTA per SM

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Same coder as RcEnc (rc_static.cuh), written so that the renormalisation is a handful of selects and one
// predicated store instead of a branch: in a warp some lane renormalises at nearly every step, so a branch
// would be taken (divergently) all the time and would also stop the two coders of a lane from overlapping.
struct RcEncV2 {
    uint64_t low, range;
    uint8_t *base;
    uint32_t pos, pend, carry;
    __device__ __forceinline__ void init(uint8_t *b) { low = 0; range = ~0ull; base = b; pos = 0; pend = 0; carry = 0; }
    __device__ __noinline__ void walk_back() {           // pending word wrapped to 0: propagate into stored words (astronomically rare)
        uint8_t *p = base + pos - 4;
        for (;;) { p -= 4; uint32_t w = *(uint32_t *)p + 1; *(uint32_t *)p = w; if (w) break; }
    }
    __device__ __forceinline__ void encode(uint32_t c0, uint32_t f) {
        range >>= PROB_BITS;
        uint64_t nl = low + range * c0;
        carry |= nl < low ? 1u : 0u;
        low = nl; range *= f;
        const bool p = (uint32_t)(range >> 32) == 0;                                  // _rcenorm_ turborc_.h:105-109
        const uint32_t np = pend + carry;
        if (p && carry && np == 0 && pos >= 8) walk_back();
        if (p && pos) *(uint32_t *)(base + pos - 4) = np;
        pend  = p ? (uint32_t)(low >> 32) : pend;
        pos   = p ? pos + 4 : pos;
        carry = p ? 0u : carry;
        low   = p ? low << 32 : low;
        range = p ? range << 32 : range;
    }
    __device__ inline void put(uint32_t w) {
        uint32_t np = pend + carry;
        if (carry && np == 0 && pos >= 8) walk_back();
        carry = 0;
        if (pos) *(uint32_t *)(base + pos - 4) = np;
        pend = w; pos += 4;
    }
    __device__ inline void flush() {                                                  // rceflush turborc_.h:118-128
        if ((uint32_t)(range >> 32) == 0) { put((uint32_t)(low >> 32)); low <<= 32; range <<= 32; }
        if (range > (1ull << 33)) { uint64_t nl = low + (1ull << 32); carry |= nl < low; low = nl; put((uint32_t)(low >> 32)); }
        else { uint64_t nl = low + 1; carry |= nl < low; low = nl; put((uint32_t)(low >> 32)); put((uint32_t)low); }
        *(uint32_t *)(base + pos - 4) = pend;
    }
};

// =========================================================================================================
// Range coder encoder, NC coders per call (rccdfsenc rccdf.c:71-82 / rccdfs2enc rccdf.c:125-143)
// The reference evaluates its overflow test after every symbol (pair); the tested quantities only grow, so
// testing once per 16-byte block and once after the last symbol decides the same way.
// =========================================================================================================
template <int NC>
__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rc_static_enc_v2(const uint8_t *__restrict__ in, Geom g, size_t n_calls, const TableSet *__restrict__ ts, size_t cpc,
                   uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ __align__(16) uint32_t ctab[256];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * blockDim.x, j = j0 + threadIdx.x;   // CTA size is chosen by the host (v2_shape)
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
        RcEncV2 e; e.init(slot);
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
        RcEncV2 e0, e1; e0.init(slot + 4); e1.init(slot + b1);
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
    const uint32_t *ip, *lim;       // 4-byte aligned stream cursor; lim = last word that may be read
    uint32_t n0, n1;                // the next two stream words, already loaded (hides the L2 latency of the refill)
    __device__ __forceinline__ uint32_t fetch() { uint32_t v = ip <= lim ? __ldg(ip) : 0u; ip++; return v; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *gend) {   // rcdinit turborc_.h:152-158
        lim = (const uint32_t *)gend - 1; ip = (const uint32_t *)p; range = ~0ull;
        uint32_t a = fetch(), b = fetch();
        code = (uint64_t)a << 32 | b;
        n0 = fetch(); n1 = fetch();
    }
    // one symbol: range >>= 15; x = max{ x : cdf[x]*range <= code } (== _cdfbget turborc_.h:307-315); _rccdfupdate
    __device__ __forceinline__ uint32_t decode(const uint8_t *lut, const uint32_t *dtab, unsigned cdfnum) {
        range >>= PROB_BITS;
        // q ~ code / range within +-1: fp32 conversions, approximate reciprocal and product each err < 2^-22
        // relative, the quotient is < 2^16, so the estimate is off by < 2^-5
        float qf = __ull2float_rz(code) * rcp_approx(__ull2float_rn(range));
        uint32_t q = min(__float2uint_rz(qf), 32767u);
        uint32_t x = lut[q], e = dtab[x];
        uint64_t rp = (uint64_t)(e >> 16) * range, fr = (uint64_t)(e & 0xffffu) * range;
        if (__builtin_expect(rp > code || (rp + fr <= code && x + 1 < cdfnum), 0)) {   // exact +-1 fix-up (rare)
            x = rp > code ? x - 1 : x + 1;
            e = dtab[x]; rp = (uint64_t)(e >> 16) * range; fr = (uint64_t)(e & 0xffffu) * range;
        }
        code -= rp; range = fr;
        const bool p = (uint32_t)(range >> 32) == 0;                                  // _rcdnorm_ turborc_.h:111
        range = p ? range << 32 : range;
        code  = p ? (code << 32 | n0) : code;
        n0 = p ? n1 : n0;
        if (p) n1 = fetch();
        return x;
    }
};

template <int NC>
__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rc_static_dec_v2(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                   size_t n_calls, const TableSet *__restrict__ ts, unsigned cdfnum, size_t cpc) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * blockDim.x, j = j0 + threadIdx.x;   // CTA size is chosen by the host (v2_shape)
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
        const uint8_t *p1 = stream + 4 + (len0 & ~3u);                               // valid streams: multiple of 4
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

// ---- hand-scheduled 32-bit-halves coders for the lane-per-coder kernels ------------------------------------
// Same arithmetic as RcEnc / RcDec (rc_static.cuh); every data-dependent event (renormalisation, carry into
// the pending word, stream refill) is a select or a predicated memory op, never a branch.
struct RcE32 {
    uint32_t rl, rh, ll, lh;        // range, low
    uint32_t pend, carry, rare;     // newest word (not stored yet), pending carry into it, "needs the slow path" flag
    uint32_t pos;                   // words put so far
    uint32_t *base;                 // word 0 of the stream; base[-1] must be writable scratch (first put stores a dummy there)
    __device__ __forceinline__ void init(uint8_t *b) { rl = rh = 0xffffffffu; ll = lh = 0; pend = carry = rare = pos = 0; base = (uint32_t *)b; }
    __device__ __forceinline__ void encode(uint32_t c0, uint32_t f) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;                    // range >>= 15
        const uint32_t tl = rl * c0, th = __umulhi(rl, c0) + rh * c0;                 // range * cdf[x]
        uint32_t cy;
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;" : "+r"(ll), "+r"(lh), "=r"(cy) : "r"(tl), "r"(th));
        carry |= cy;                                                                  // == reference "ilow > low" at the next renorm
        const uint32_t nl = rl * f, nh = __umulhi(rl, f) + rh * f;                    // range *= freq
        const bool p = nh == 0;                                                       // _rcenorm_ turborc_.h:105-109
        const uint32_t np = pend + carry;
        rare |= (p && np < carry) ? 1u : 0u;                                          // pending word wrapped: carry must walk further back
        if (p) base[(int)pos - 1] = np;
        pend = p ? lh : pend; pos += p ? 1u : 0u; carry = p ? 0u : carry;
        lh = p ? ll : lh; ll = p ? 0u : ll;
        rh = p ? nl : nh; rl = p ? 0u : nl;
    }
    // replicated-state form (warp-cooperative kernels): every lane tracks the coder, only `writer` lanes store
    __device__ __forceinline__ void encode_w(uint32_t c0, uint32_t f, bool writer) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const uint32_t tl = rl * c0, th = __umulhi(rl, c0) + rh * c0;
        uint32_t cy;
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;" : "+r"(ll), "+r"(lh), "=r"(cy) : "r"(tl), "r"(th));
        carry |= cy;
        const uint32_t nl = rl * f, nh = __umulhi(rl, f) + rh * f;
        const bool p = nh == 0;
        const uint32_t np = pend + carry;
        rare |= (p && np < carry) ? 1u : 0u;
        if (p && writer) base[(int)pos - 1] = np;
        pend = p ? lh : pend; pos += p ? 1u : 0u; carry = p ? 0u : carry;
        lh = p ? ll : lh; ll = p ? 0u : ll;
        rh = p ? nl : nh; rl = p ? 0u : nl;
    }
    __device__ __forceinline__ void put_w(uint32_t w, bool writer) {
        const uint32_t np = pend + carry;
        rare |= (np < carry) ? 1u : 0u;
        if (writer) base[(int)pos - 1] = np;
        pend = w; pos++; carry = 0;
    }
    __device__ inline void flush_w(bool writer) {
        if (rh == 0) { put_w(lh, writer); lh = ll; ll = 0; rh = rl; rl = 0; }
        if (rh > 2u || (rh == 2u && rl != 0)) { add_low(0, 1); put_w(lh, writer); }
        else { add_low(1, 0); put_w(lh, writer); put_w(ll, writer); }
        if (writer) base[(int)pos - 1] = pend;
    }
    __device__ __forceinline__ void put(uint32_t w) {                                 // flush path only
        const uint32_t np = pend + carry;
        rare |= (np < carry) ? 1u : 0u;
        base[(int)pos - 1] = np;
        pend = w; pos++; carry = 0;
    }
    __device__ __forceinline__ void add_low(uint32_t al, uint32_t ah) {
        uint32_t cy;
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;" : "+r"(ll), "+r"(lh), "=r"(cy) : "r"(al), "r"(ah));
        carry |= cy;
    }
    __device__ inline void flush() {                                                  // rceflush turborc_.h:118-128
        if (rh == 0) { put(lh); lh = ll; ll = 0; rh = rl; rl = 0; }
        if (rh > 2u || (rh == 2u && rl != 0)) { add_low(0, 1); put(lh); }             // range > 2^33
        else { add_low(1, 0); put(lh); put(ll); }
        base[(int)pos - 1] = pend;
    }
    __device__ __forceinline__ uint32_t bytes() const { return pos * 4; }
};

struct RcD32 {
    uint32_t rl, rh, cl, ch;        // range, code
    uint32_t n0, n1;                // next two stream words, already loaded
    uint32_t wi, wlim;              // index of the next word to load; last index that may be loaded
    const uint32_t *base;
    __device__ __forceinline__ uint32_t fetch() { uint32_t v = wi <= wlim ? __ldg(base + wi) : 0u; wi++; return v; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *gend) {   // rcdinit turborc_.h:152-158
        base = (const uint32_t *)p;
        long long words = (gend - p) >> 2;
        wi = 0; wlim = words > 0 ? (uint32_t)(words - 1) : 0u;
        if (words <= 0) { wi = 1; wlim = 0; }                                         // nothing readable
        rl = rh = 0xffffffffu;
        ch = fetch(); cl = fetch(); n0 = fetch(); n1 = fetch();
    }
    __device__ __forceinline__ uint32_t decode(const uint8_t *lut, const uint32_t *dtab, unsigned cdfnum) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;                    // _rccdfrange
        // q ~ code / range within +-1 (three fp32 roundings < 2^-22 relative on a quotient < 2^16)
        const float qf = __ull2float_rz((uint64_t)ch << 32 | cl) * rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl));
        // floor without a float->int conversion: (min(qf, 32767) - 0.5) + 1.5*2^23 rounds to nearest and leaves the
        // integer in the low mantissa bits (an exact-integer qf may land one low; the fix-up below covers it)
        const uint32_t q = __float_as_uint((fminf(qf, 32767.0f) - 0.5f) + 12582912.0f) & 0xffffu;
        uint32_t x = lut[q], e = dtab[x];
        uint32_t c0 = e >> 16, f = e & 0xffffu;
        uint32_t pl = rl * c0, ph = __umulhi(rl, c0) + rh * c0;                       // rp = cdf[x] * range
        uint32_t fl = rl * f, fh = __umulhi(rl, f) + rh * f;                          // fr = freq * range
        uint32_t dl, dh, bw;                                                          // d = code - rp, borrow => estimate too high
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;" : "=r"(dl), "=r"(dh), "=r"(bw) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        const bool low = dh > fh || (dh == fh && dl >= fl);                           // code - rp >= fr => estimate too low
        if (__builtin_expect(bw != 0 || low, 0)) {                                    // exact +-1 fix-up (rare)
            if (bw) x--; else if (x + 1 < cdfnum) x++;
            e = dtab[x]; c0 = e >> 16; f = e & 0xffffu;
            pl = rl * c0; ph = __umulhi(rl, c0) + rh * c0;
            fl = rl * f; fh = __umulhi(rl, f) + rh * f;
            asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        }
        const bool p = fh == 0;                                                       // _rcdnorm_ turborc_.h:111
        rh = p ? fl : fh; rl = p ? 0u : fl;
        ch = p ? dl : dh; cl = p ? n0 : dl;
        n0 = p ? n1 : n0;
        if (p) n1 = fetch();
        return x;
    }
};

// Range decoder whose stream words arrive through a per-lane ring in shared memory.
//  * The ring (RING_W words per lane, word-major / lane-minor so any per-lane index is bank-conflict free) is
//    topped up once per 8-symbol block with one aligned 16-byte global load; the loaded registers are only
//    touched by the shared-memory stores at the END of the block, so the global latency hides behind the block
//    instead of stalling the first instruction that names the register (what a register prefetch queue does).
//  * Symbols are decoded speculatively: x = lut[~code/range] from an fp32 estimate, verified with the exact
//    64-bit products the update needs anyway.  A failed check only sets a flag; the (rare) flagged block is
//    re-decoded from its saved start state by the exact binary-search path, so the hot loop has no branch.
constexpr int RING_W = 16;

struct RcDRing {
    uint32_t rl, rh, cl, ch;        // range, code
    uint32_t n0, n1;                // next two stream words (already in registers)
    uint32_t ci;                    // ring read cursor: absolute word index (relative to qbase) of the next word to move into n1
    uint32_t bad;
    __device__ __forceinline__ void step(const uint8_t *lut, const uint32_t *dtab, const uint32_t *ring, uint32_t rstride, uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;                    // _rccdfrange
        // floor(code/range) without a float->int conversion: (q - 0.5) + 1.5*2^23 rounds to nearest and leaves the integer
        // in the low mantissa bits; the mask keeps garbage streams inside the LUT
        const float qf = fmaf(__ull2float_rz((uint64_t)ch << 32 | cl), rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl)), -0.5f);
        const uint32_t q = __float_as_uint(qf + 12582912.0f) & 0x7fffu;
        const uint32_t x = lut[q], e = dtab[x];
        const uint32_t c0 = e >> 16, f = e & 0xffffu;
        const uint32_t pl = rl * c0, ph = __umulhi(rl, c0) + rh * c0;                 // rp = cdf[x] * range
        const uint32_t fl = rl * f, fh = __umulhi(rl, f) + rh * f;                    // fr = freq * range
        uint32_t dl, dh, bw;                                                          // d = code - rp (borrow => x too high)
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;" : "=r"(dl), "=r"(dh), "=r"(bw) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= bw | ((dh > fh || (dh == fh && dl >= fl)) ? 1u : 0u);                  // d >= fr => x too low
        const bool p = fh == 0;                                                       // _rcdnorm_ turborc_.h:111
        rh = p ? fl : fh; rl = p ? 0u : fl;
        ch = p ? dl : dh; cl = p ? n0 : dl;
        n0 = p ? n1 : n0;
        if (p) n1 = ring[(ci & (RING_W - 1)) * rstride];
        ci += p ? 1u : 0u;
        x_out = x;
    }
};

// exact symbol step straight from global memory (redo path and tails): binary search == _cdfbget turborc_.h:307-315
struct RcDExact {
    uint64_t range, code;
    const uint32_t *base; uint32_t wi, wlim;          // wi = index of the next word to read
    __device__ __forceinline__ uint32_t fetch() { uint32_t v = wi <= wlim ? __ldg(base + wi) : 0u; wi++; return v; }
    __device__ inline uint32_t step(const uint32_t *dtab, unsigned cdfnum) {
        range >>= PROB_BITS;
        unsigned x = 0, hi = cdfnum;
        while (x + 1 < hi) { unsigned mid = (x + hi) >> 1; if ((uint64_t)(dtab[mid] >> 16) * range > code) hi = mid; else x = mid; }
        const uint32_t e = dtab[x];
        code -= (uint64_t)(e >> 16) * range; range *= (e & 0xffffu);
        if ((uint32_t)(range >> 32) == 0) { range <<= 32; code = code << 32 | fetch(); }
        return x;
    }
    __device__ inline uint32_t step2(const uint2 *dtab, unsigned cdfnum) {          // same, table of {cdf, freq} pairs
        range >>= PROB_BITS;
        unsigned x = 0, hi = cdfnum;
        while (x + 1 < hi) { unsigned mid = (x + hi) >> 1; if ((uint64_t)dtab[mid].x * range > code) hi = mid; else x = mid; }
        const uint2 e = dtab[x];
        code -= (uint64_t)e.x * range; range *= e.y;
        if ((uint32_t)(range >> 32) == 0) { range <<= 32; code = code << 32 | fetch(); }
        return x;
    }
};

// =========================================================================================================
// TRC_RCS2, one LANE PER CODER: lanes 2r / 2r+1 of a warp own coder 0 / coder 1 of call r, so a warp carries 32
// independent range coders over 16 calls and the batch exposes twice as many warps to the schedulers as the
// lane-per-call form (the coders are dependency chains: throughput comes from warps in flight).
// The two lanes never talk inside the loop: each tests its own half of OVERFLOWI (rccdf.c:46: stream 1 against
// the size threshold, stream 0 against the start of stream 1), both halves are monotone, and they are OR-ed once
// at the end.
// =========================================================================================================

// calls_per_cta: LPC_NT/2 for big batches; for batches of about one wave the host sizes the CTAs so that every SM gets
// the same number of coders (ceil(n_calls / #SM) calls per CTA, one CTA per SM): with 2-3 CTAs per SM the fuller SMs
// would set the kernel time.
__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rcs2_enc_lpc(const uint8_t *__restrict__ in, Geom g, size_t n_calls, const TableSet *__restrict__ ts, size_t cpc,
               uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta, unsigned calls_per_cta) {
    __shared__ __align__(16) uint32_t ctab[256];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * calls_per_cta, j = j0 + (threadIdx.x >> 1);
    const unsigned c = threadIdx.x & 1;
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) tma_fetch(ctab, t->ctab, sizeof ctab, &bar);
    __syncthreads();
    tma_wait(&bar);
    const bool live = j < n_calls && (threadIdx.x >> 1) < calls_per_cta;   // dead lanes still take part in the shuffles below
    size_t start = 0, n = 0;
    if (live) call_span(g, j, start, n);
    const uint8_t *ip = in + start;
    uint8_t *slot = slots + (live ? j : 0) * slot_stride;
    const int64_t thr = rc_thr(n);
    const bool tiny = n < 4;                       // reference undefined; raw
    const uint32_t b1ref = tiny ? 4 : 4 + (uint32_t)(((n - 4) * 37) / 64);           // rccdf.c:126
    const uint32_t b1 = (b1ref + 64 + 15) & ~15u;
    RcE32 e; e.init(slot + (c ? b1 : 4));
    bool raw = tiny || !live;
    const size_t nb = n & ~(size_t)15;
    uint4 cur = (nb && !raw) ? ldg128(ip) : make_uint4(0, 0, 0, 0);
    uint4 nxt = (nb >= 32 && !raw) ? ldg128(ip + 16) : cur;
    for (size_t i = 0; i < nb && !raw; i += 16) {
        const uint4 nxt2 = i + 48 <= nb ? ldg128(ip + i + 32) : nxt;                  // two blocks ahead: covers a DRAM miss
        const uint32_t w[4] = { cur.x >> (8 * c), cur.y >> (8 * c), cur.z >> (8 * c), cur.w >> (8 * c) };
        uint32_t tt[8];
#pragma unroll
        for (int k = 0; k < 8; k++) tt[k] = ctab[(w[k >> 1] >> (16 * (k & 1))) & 0xff];
#pragma unroll
        for (int k = 0; k < 8; k++) e.encode(tt[k] & 0xffffu, tt[k] >> 16);
        raw = c ? (int64_t)b1ref + e.bytes() >= thr : 4 + e.bytes() >= b1ref;       // own half of OVERFLOWI
        cur = nxt; nxt = nxt2;
    }
    for (size_t i = nb + c; i < (n & ~(size_t)1) && !raw; i += 2) {                  // remaining full pairs
        uint32_t tk = ctab[ip[i]]; e.encode(tk & 0xffffu, tk >> 16);
        raw = c ? (int64_t)b1ref + e.bytes() >= thr : 4 + e.bytes() >= b1ref;
    }
    raw = __shfl_xor_sync(0xffffffffu, (int)raw, 1) || raw;                          // either half fired -> raw copy
    if (!raw) {
        if (c == 0 && (n & 1)) { uint32_t tk = ctab[ip[n - 1]]; e.encode(tk & 0xffffu, tk >> 16); }   // odd tail on coder 0 (rccdf.c:135-136)
        e.flush();
    }
    const uint32_t mypos = e.bytes(), other = __shfl_xor_sync(0xffffffffu, mypos, 1);
    const uint32_t rare = e.rare | __shfl_xor_sync(0xffffffffu, e.rare, 1);
    if (!live || c) return;
    const uint32_t p0 = mypos, p1 = other;
    if (!raw) {
        *(uint32_t *)slot = p0;                                                      // rccdf.c:141
        if ((int64_t)(4 + p0 + p1) >= thr) raw = true;                               // rccdf.c:142
    }
    UnitMeta m; m.pref = 0; m.pad = 0; m.a_off = 0;
    if (rare && !raw) {   // a pending word wrapped under a carry (p ~ 2^-32 per word): redo this call with the walk-back coder
        rc_static_enc_call<2, true>(ip, n, ctab, nullptr, slot, m);
        meta[j] = m;
        return;
    }
    m.a_len = raw ? 0 : 4 + p0; m.b_off = b1; m.b_len = raw ? 0 : p1;
    m.len = raw ? (uint32_t)n : 4 + p0 + p1; m.flags = raw ? UM_RAW : 0;
    meta[j] = m;
}

// look-back words of the fused encoder (rcs2_v3.cuh k_rcs2_enc3): flag in the top two bits, running byte count below
constexpr int LB_U = 24;                                        // words in flight per lane in the layout epilogue
constexpr unsigned long long LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_VAL = LB_AGG - 1;

__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rcs2_dec_lpc(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
               size_t n_calls, const TableSet *__restrict__ ts, unsigned cdfnum, size_t cpc, unsigned calls_per_cta) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    extern __shared__ uint32_t ringbuf[];                                             // RING_W * blockDim.x words
    __shared__ uint64_t bar;
    const uint32_t rstride = blockDim.x;
    const size_t j0 = (size_t)blockIdx.x * calls_per_cta, j = j0 + (threadIdx.x >> 1);
    const unsigned c = threadIdx.x & 1;
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
    const bool live = j < n_calls && (threadIdx.x >> 1) < calls_per_cta;
    size_t start = 0, n = 0;
    uint64_t so = 0, sl = 0;
    if (live) { call_span(g, j, start, n); so = in_off[j]; sl = in_off[j + 1] - so; }
    const uint8_t *gend = in + in_off[g.n_calls], *stream = in + so;
    uint8_t *op = out + start;
    const bool rawc = sl == n;                                                       // raw chunk: both lanes copy half
    if (rawc || !live) {
        if (live) { size_t h = (n / 2) & ~(size_t)15; if (c == 0) thread_copy(op, stream, h); else thread_copy(op + h, stream + h, n - h); }
        n = 0;                                                                       // still join the shuffles
    }
    uint32_t len0 = n ? ld_u32_clamped(stream, gend) : 0;
    const uint8_t *p = stream + 4 + (c ? (len0 & ~3u) : 0);                          // stream c (rccdf.c:167)
    if (p > gend || p < stream || n == 0) p = gend;
    // word addressing relative to the 16-byte aligned base below p
    const uint32_t *qbase = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)15);
    const long long wavail = ((const uint8_t *)gend - (const uint8_t *)qbase) >> 2;   // words readable from qbase (may be <= 0)
    const uint32_t wlim = wavail > 0 ? (uint32_t)(wavail - 1) : 0u;
    const bool none = wavail <= 0;
    auto gword = [&](uint32_t w) -> uint32_t { return (!none && w <= wlim) ? __ldg(qbase + w) : 0u; };
    auto gquad = [&](uint32_t w) -> uint4 {                                           // w multiple of 4
        if (!none && w + 3 <= wlim) return __ldg((const uint4 *)(qbase + w));
        return make_uint4(gword(w), gword(w + 1), gword(w + 2), gword(w + 3));
    };
    uint32_t *ring = ringbuf + threadIdx.x;
    auto ring_put = [&](uint32_t w, const uint4 &v) {                                 // w multiple of 4
        ring[((w + 0) & (RING_W - 1)) * rstride] = v.x; ring[((w + 1) & (RING_W - 1)) * rstride] = v.y;
        ring[((w + 2) & (RING_W - 1)) * rstride] = v.z; ring[((w + 3) & (RING_W - 1)) * rstride] = v.w;
    };
    RcDRing d;
    uint32_t fi;                                                                      // words [fi-RING_W, fi) are in the ring
    auto resync = [&](uint32_t w0, uint64_t range, uint64_t code) {                   // (re)start the ring at word w0 = next unread word
        fi = w0 & ~3u;
        for (int k = 0; k < 3; k++) { ring_put(fi, gquad(fi)); fi += 4; }
        d.rl = (uint32_t)range; d.rh = (uint32_t)(range >> 32); d.cl = (uint32_t)code; d.ch = (uint32_t)(code >> 32);
        d.n0 = ring[(w0 & (RING_W - 1)) * rstride]; d.n1 = ring[((w0 + 1) & (RING_W - 1)) * rstride];
        d.ci = w0 + 2; d.bad = 0;
    };
    {
        const uint32_t w0 = (uint32_t)(((uintptr_t)p & 15) >> 2);
        resync(w0 + 2, ~0ull, (uint64_t)gword(w0) << 32 | gword(w0 + 1));             // rcdinit turborc_.h:152-158
    }
    const size_t nb = n & ~(size_t)15;
    const size_t nbmax = __reduce_max_sync(0xffffffffu, (unsigned)nb);               // warp-uniform trip count for the shuffles
    for (size_t i = 0; i < nbmax; i += 16) {
        uint32_t a0 = 0, a1 = 0;
        if (i < nb) {
            const bool need = fi - d.ci <= 10;                                        // top the ring up (stores happen after the block)
            uint4 t4 = make_uint4(0, 0, 0, 0);
            if (need) t4 = gquad(fi);
            const RcDRing s0 = d;                                                     // block start state (for the redo path)
            uint32_t x;
#pragma unroll
            for (int k = 0; k < 4; k++) { d.step(lut, dtab, ring, rstride, x); a0 |= x << (8 * k); }
#pragma unroll
            for (int k = 0; k < 4; k++) { d.step(lut, dtab, ring, rstride, x); a1 |= x << (8 * k); }
            const bool dry = d.ci > fi;                                               // a read ran past the filled part of the ring
            if (need) { ring_put(fi, t4); fi += 4; }
            if (__builtin_expect(d.bad != 0 || dry, 0)) {                             // estimate missed (or ring ran dry): exact redo
                RcDExact ex;
                ex.range = (uint64_t)s0.rh << 32 | s0.rl; ex.code = (uint64_t)s0.ch << 32 | s0.cl;
                ex.base = qbase; ex.wlim = wlim; ex.wi = none ? 1u : s0.ci - 2;      // n0/n1 were words ci-2, ci-1
                a0 = a1 = 0;
                for (int k = 0; k < 4; k++) a0 |= ex.step(dtab, cdfnum) << (8 * k);
                for (int k = 0; k < 4; k++) a1 |= ex.step(dtab, cdfnum) << (8 * k);
                resync(ex.wi, ex.range, ex.code);
            }
        }
        // a0/a1 = this coder's symbols 0-3 / 4-7 of the block; interleave with the partner's
        uint32_t b0 = __shfl_xor_sync(0xffffffffu, a0, 1), b1 = __shfl_xor_sync(0xffffffffu, a1, 1);
        if (i < nb) {
            uint32_t e0 = c ? b0 : a0, o0 = c ? a0 : b0, e1 = c ? b1 : a1, o1 = c ? a1 : b1;   // even-position / odd-position symbols
            // bytes: even0 odd0 even1 odd1 | even2 odd2 even3 odd3 ...
            uint2 v = c ? make_uint2(__byte_perm(e1, o1, 0x5140), __byte_perm(e1, o1, 0x7362))
                        : make_uint2(__byte_perm(e0, o0, 0x5140), __byte_perm(e0, o0, 0x7362));
            *(uint2 *)(op + i + 8 * c) = v;
        }
    }
    // remaining full pairs, then the odd tail on coder 0 (rccdf.c:179-182): exact path
    if (n > nb) {
        RcDExact ex;
        ex.range = (uint64_t)d.rh << 32 | d.rl; ex.code = (uint64_t)d.ch << 32 | d.cl;
        ex.base = qbase; ex.wlim = wlim; ex.wi = none ? 1u : d.ci - 2;
        for (size_t i = nb + c; i < (n & ~(size_t)1); i += 2) op[i] = (uint8_t)ex.step(dtab, cdfnum);
        if (c == 0 && (n & 1)) op[n - 1] = (uint8_t)ex.step(dtab, cdfnum);
    }
}

// =========================================================================================================
// Static rANS (anscdf4senc / anscdf4sdec): both states of a call in one lane
// =========================================================================================================
// ece (anscdf_.h:90-94) with the renormalisation as selects + one predicated 32-bit store (see RansWriter)
__device__ __forceinline__ uint32_t rans_enc_step_v2(uint32_t s, const uint4 e, uint8_t *base, int &pos, uint32_t &acc) {
    const bool p = s >= e.y;
    const int np = pos - 2;
    const uint32_t nacc = __byte_perm(s, acc, 0x5410);
    if (p && !(np & 2)) *(uint32_t *)(base + np) = nacc;
    pos = p ? np : pos; acc = p ? nacc : acc; s = p ? s >> 16 : s;
    const uint32_t q = __umulhi(s, e.x) >> (e.z >> 16);
    return s + e.w + q * (e.z & 0xffffu);
}

__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rans_static_enc_v2(const uint8_t *__restrict__ in, Geom g, size_t n_calls, const TableSet *__restrict__ ts, size_t cpc,
                     uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ __align__(16) uint4 etab[256];
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * blockDim.x, j = j0 + threadIdx.x;   // CTA size is chosen by the host (v2_shape)
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
    while (i > n4) { i--; s0 = rans_enc_step_v2(s0, etab[ip[i]], w.base, w.pos, w.acc); }                   // tail on state 0 (anscdf.c:62-64)
    while (i > n16) {                                                                 // groups of 4 down to a 16-byte boundary
        i -= 4;
        uint32_t v = *(const uint32_t *)(ip + i);
        s1 = rans_enc_step_v2(s1, etab[v >> 24], w.base, w.pos, w.acc); s0 = rans_enc_step_v2(s0, etab[(v >> 16) & 0xff], w.base, w.pos, w.acc);
        s1 = rans_enc_step_v2(s1, etab[(v >> 8) & 0xff], w.base, w.pos, w.acc); s0 = rans_enc_step_v2(s0, etab[v & 0xff], w.base, w.pos, w.acc);
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
            s1 = rans_enc_step_v2(s1, ea, w.base, w.pos, w.acc); s0 = rans_enc_step_v2(s0, eb, w.base, w.pos, w.acc);
            s1 = rans_enc_step_v2(s1, ec, w.base, w.pos, w.acc); s0 = rans_enc_step_v2(s0, ed, w.base, w.pos, w.acc);
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
    const uint16_t *ip, *lim;       // 2-byte aligned stream cursor; lim = last halfword that may be read
    uint32_t n0, n1;                // the next two 16-bit words, already loaded
    __device__ __forceinline__ uint32_t fetch() { uint32_t v = ip <= lim ? (uint32_t)__ldg(ip) : 0u; ip++; return v; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *gend, uint32_t &s0, uint32_t &s1) {
        lim = (const uint16_t *)gend - 1; ip = (const uint16_t *)p;
        uint32_t a = fetch(), b = fetch(), c = fetch(), d = fetch();
        s0 = a | b << 16; s1 = c | d << 16;                                           // ecdini anscdf_.h:47
        n0 = fetch(); n1 = fetch();
    }
    __device__ __forceinline__ uint32_t step(uint32_t &s, const uint8_t *lut, const uint32_t *dtab) {
        const uint32_t r = s & PROB_MASK, x = lut[r], e = dtab[x];
        s = (e & 0xffffu) * (s >> PROB_BITS) + r - (e >> 16);                         // STATEUPD cdf_.h:37
        const bool p = s < ANS_L;                                                     // ecdnorm anscdf_.h:50-73
        s = p ? (s << 16 | n0) : s;
        n0 = p ? n1 : n0;
        if (p) n1 = fetch();
        return x;
    }
};

__global__ void __launch_bounds__(LPC_MAX_NT, 1)
k_rans_static_dec_v2(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
                     size_t n_calls, const TableSet *__restrict__ ts, size_t cpc, unsigned flags) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    extern __shared__ uint32_t ringbuf[];                                             // RING_W * blockDim.x words
    const uint32_t V2S = blockDim.x;
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * blockDim.x, j = j0 + threadIdx.x;   // CTA size is chosen by the host (v2_shape)
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
    // stream halfwords arrive through a per-lane shared-memory ring (see RcDRing): hi = index of the next halfword to
    // move into n1, relative to the 16-byte aligned base below the stream start
    const uint8_t *p = in + so;
    const uint32_t *qbase = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)15);
    const long long wavail = ((const uint8_t *)gend - (const uint8_t *)qbase) >> 2;
    const uint32_t wlim = wavail > 0 ? (uint32_t)(wavail - 1) : 0u;
    const bool none = wavail <= 0;
    const uint32_t hend = (uint32_t)(((const uint8_t *)gend - (const uint8_t *)qbase) >> 1);   // halfwords readable
    auto gword = [&](uint32_t w) -> uint32_t {
        if (!none && w <= wlim) return __ldg(qbase + w);
        return (!none && 2 * w < hend) ? (uint32_t)__ldg((const uint16_t *)qbase + 2 * w) : 0u;   // last odd halfword
    };
    auto gquad = [&](uint32_t w) -> uint4 {
        if (!none && w + 3 <= wlim) return __ldg((const uint4 *)(qbase + w));
        return make_uint4(gword(w), gword(w + 1), gword(w + 2), gword(w + 3));
    };
    uint32_t *ring = ringbuf + threadIdx.x;
    auto ring_put = [&](uint32_t w, const uint4 &v) {
        ring[((w + 0) & (RING_W - 1)) * V2S] = v.x; ring[((w + 1) & (RING_W - 1)) * V2S] = v.y;
        ring[((w + 2) & (RING_W - 1)) * V2S] = v.z; ring[((w + 3) & (RING_W - 1)) * V2S] = v.w;
    };
    auto ring_hw = [&](uint32_t h) -> uint32_t { uint32_t w = ring[((h >> 1) & (RING_W - 1)) * V2S]; return (h & 1) ? w >> 16 : w & 0xffffu; };
    uint32_t hi = (uint32_t)(((uintptr_t)p & 15) >> 1);
    uint32_t fi = 0;                                                                  // words [.., fi) are in the ring
    for (int k = 0; k < 3; k++) { ring_put(fi, gquad(fi)); fi += 4; }
    uint32_t s0 = ring_hw(hi) | ring_hw(hi + 1) << 16, s1 = ring_hw(hi + 2) | ring_hw(hi + 3) << 16;   // mnfill anscdf_.h:176
    uint32_t n0 = ring_hw(hi + 4), n1 = ring_hw(hi + 5);
    hi += 6;
#define TRC_RSTEP(_s_, _x_) { const uint32_t r_ = _s_ & PROB_MASK, x_ = lut[r_], e_ = dtab[x_]; \
        _s_ = (e_ & 0xffffu) * (_s_ >> PROB_BITS) + r_ - (e_ >> 16);                  /* STATEUPD cdf_.h:37 */ \
        const bool p_ = _s_ < ANS_L;                                                  /* ecdnorm anscdf_.h:50-73 */ \
        _s_ = p_ ? (_s_ << 16 | n0) : _s_; n0 = p_ ? n1 : n0; \
        if (p_) { const uint32_t w_ = ring[((hi >> 1) & (RING_W - 1)) * V2S]; n1 = (hi & 1) ? w_ >> 16 : w_ & 0xffffu; } \
        hi += p_ ? 1u : 0u; _x_ = x_; }
    const uint32_t n4 = n & ~3u, n8 = n & ~7u;
    uint32_t o = 0;
    for (; o < n8; o += 8) {                                                          // anscdf.c:82, 8 symbols per ring top-up
        const bool need = 2 * fi - hi <= 22;                                          // fewer than 12 words buffered (a block eats <= 4)
        uint4 t4 = make_uint4(0, 0, 0, 0);
        if (need) t4 = gquad(fi);
        uint32_t a, b, c, e, w0, w1;
        TRC_RSTEP(s1, a) TRC_RSTEP(s0, b) TRC_RSTEP(s1, c) TRC_RSTEP(s0, e)
        w0 = a | b << 8 | c << 16 | e << 24;
        TRC_RSTEP(s1, a) TRC_RSTEP(s0, b) TRC_RSTEP(s1, c) TRC_RSTEP(s0, e)
        w1 = a | b << 8 | c << 16 | e << 24;
        if (need) { ring_put(fi, t4); fi += 4; }
        *(uint2 *)(op + o) = make_uint2(w0, w1);
    }
    // tails: straight from global memory (hi-2, hi-1 are the halfwords sitting in n0, n1)
    RansReader2 rd;
    rd.lim = (const uint16_t *)gend - 1; rd.ip = (const uint16_t *)qbase + (hi - 2);
    rd.n0 = rd.fetch(); rd.n1 = rd.fetch();
    for (; o < n4; o += 4) {
        uint32_t a = rd.step(s1, lut, dtab), b = rd.step(s0, lut, dtab), c = rd.step(s1, lut, dtab), e = rd.step(s0, lut, dtab);
        *(uint32_t *)(op + o) = a | b << 8 | c << 16 | e << 24;
    }
    for (; o < n; o++) op[o] = (uint8_t)((flags & 1u) ? rd.step(s0, lut, dtab) : rd.step(s1, lut, dtab));   // anscdf.c:83
#undef TRC_RSTEP
}

}  // namespace trc
