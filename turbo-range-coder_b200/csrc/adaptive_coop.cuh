// adaptive_coop.cuh -- warp-cooperative kernels for the adaptive byte rANS codecs TRC_ANS (anscdfenc/anscdfdec,
// anscdf.c:567-605) and TRC_ANS1 (order-1, anscdf.c:607-645).  Same bytes as the lane-per-unit kernels in
// adaptive.cuh; built for the latency of ONE unit, which is what bounds these codecs: the adaptive model is a
// serial chain over the nibbles of a block, and a 64 KiB (or 4 MiB) chain run by a single lane costs ~650 cycles per
// nibble (16 dependent table-entry updates), a whole-buffer drop-in call is slower than the CPU.
//
// One WARP owns one unit.  A 16-entry CDF lives one entry per lane (cdf16upd, cdf_.h:46-50, is then ONE
// data-parallel step instead of 16 serial ones):
//   encoder model pass: lanes 0-15 update the high-nibble table while lanes 16-31 update the low-nibble table of
//     the same byte; (cum, freq) of the coded symbol come from two warp shuffles; lanes 0 and 16 push the two
//     records (mnenc4 anscdf_.h:106).  Tables sit in shared memory (17 x 16 entries for order 0, 256 x 17 x 16 =
//     136 KB for order 1: one warp per SM owns the whole table set, nothing goes to L2).
//   encoder coding pass: lanes 0-3 are the four rANS states; the words they emit in one step are compacted into
//     the reference's LIFO order with a ballot + popcount prefix (mnflush anscdf_.h:128-138).
//   decoder: the symbol search of cdf16ansdec (cdf_.h:52-59: SIMD compare + movemask + ctz) is a compare + ballot +
//     popcount across the 16 lanes; every lane keeps a copy of the four states and the stream cursor, so nothing has
//     to be broadcast.
#pragma once
#include "trc_common.cuh"
#include "adaptive.cuh"
#include "static_v2.cuh"

namespace trc {

constexpr int ADAPT_IC_ = 10;                                 // IC cdf_.h:35
constexpr size_t COOP_MIN_LANE_UNITS = 8192;                  // at least this many units: lane-per-unit kernels win
constexpr int COOP_WPB = 4;                                   // warps (units) per CTA, order 0
constexpr int O1_CTX_ENTRIES = 17 * 16;                       // entries per context: mbh[16] + mbl[16][16]
constexpr size_t O1_SMEM_BYTES = (size_t)256 * O1_CTX_ENTRIES * sizeof(uint16_t);   // 139 264

__device__ __forceinline__ int adapt_entry(int m, int i, bool greater) {      // cdf16upd on one 16-bit lane
    return m + (((ADAPT_IC_ * i + (greater ? (int)AD_MIX : 0)) - m) >> 7);
}

// ---- encoder --------------------------------------------------------------------------------------------------
template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : COOP_WPB * 32)
k_ans_byte_enc_coop(const uint8_t *__restrict__ in, Geom g, uint8_t *__restrict__ slots, size_t slot_stride,
                    uint32_t *__restrict__ recs, size_t rec_stride, UnitMeta *__restrict__ meta) {
    extern __shared__ __align__(16) uint16_t smem_tabs[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned h = lane >> 4, i = lane & 15;              // h: 0 = high-nibble table, 1 = low-nibble table
    uint16_t *T = smem_tabs + (size_t)wib * (O1 ? 256 * O1_CTX_ENTRIES : O1_CTX_ENTRIES);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    for (size_t u = gw; u < g.n_units; u += nwarps) {
        size_t j, start, len; uint32_t b;
        unit_span(g, u, j, b, start, len);
        if (len == 0) {
            if (lane == 0) { UnitMeta m; m.len = m.a_off = m.a_len = m.b_off = m.b_len = m.flags = m.pref = m.pad = 0; meta[u] = m; }
            continue;
        }
        const uint8_t *ip = in + start;
        const uint32_t n = (uint32_t)len, npairs = (n + 1) >> 1;
        uint32_t *rec = recs + u * rec_stride;
        // CDF16DEC0/1/2 (cdf_.h:26-32): every table = j << 11
        for (uint32_t k = lane; k < (O1 ? 256u * O1_CTX_ENTRIES : (uint32_t)O1_CTX_ENTRIES); k += 32) T[k] = (uint16_t)((k & 15) << 11);
        __syncwarp();
        uint32_t cx = (O1 && start > j * g.chunk) ? in[start - 1] : 0;       // cx carries across the blocks of a call (anscdf.c:608)
        // ---- model pass: one byte per step, both of its nibbles in parallel (mnenc8x2 / mnenc8x2x anscdf_.h:114-126).
        // The table a lane-half is working on stays cached in a register while consecutive bytes keep selecting it
        // (always true for the order-0 high-nibble table, frequent for the others on run-heavy data): the dependent
        // chain per byte is then shuffle + update, the shared-memory store is write-behind.
        uint16_t *cur_tab = T + (h ? 16 : 0);                                  // table currently cached in m_cache
        int m_cache = cur_tab[i];
        for (uint32_t base = 0; base < 2 * npairs; base += 32) {
            const uint32_t idx = base + lane;
            const uint32_t mine = idx < n ? ip[idx] : 0;                       // odd tail: dummy 0 byte (anscdf.c:581)
            const uint32_t cnt = 2 * npairs - base < 32 ? 2 * npairs - base : 32;
            uint32_t myrec = 0;                                                // lanes 0/16 collect 16 records each, stored coalesced below
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t x = __shfl_sync(0xffffffffu, mine, k), yh = x >> 4, xs = h ? (x & 15) : yh;
                uint16_t *tab = T + (O1 ? (size_t)cx * O1_CTX_ENTRIES : 0) + (h ? (1 + yh) * 16 : 0);
                if (tab != cur_tab) { cur_tab[i] = (uint16_t)m_cache; m_cache = tab[i]; cur_tab = tab; }   // uniform per half-warp
                const int m = m_cache;
                const int mx = __shfl_sync(0xffffffffu, m, (lane & 16) | xs);
                int mx1 = __shfl_sync(0xffffffffu, m, (lane & 16) | ((xs + 1) & 15));
                if (xs == 15) mx1 = (int)PROB_TOTAL;
                const uint32_t r = (uint32_t)(mx1 - mx) | (uint32_t)mx << 16;
                // record of byte base+k: index 2*(base+k)+h.  Lane (k&15)+16h of each half keeps it -> 128-byte stores per 16 bytes
                if (i == (k & 15)) myrec = r;
                m_cache = adapt_entry(m, (int)i, i > xs);
                if (O1) cx = x;
                if ((k & 15) == 15 || k + 1 == cnt) {                          // flush 16 (or fewer) byte-records of both halves
                    const uint32_t b0 = base + (k & ~15u);                     // first byte of this group
                    if (i <= (k & 15)) rec[2 * (b0 + i) + h] = myrec;
                }
            }
        }
        cur_tab[i] = (uint16_t)m_cache;
        __syncwarp();
        __threadfence_block();
        // ---- coding pass: lanes 0-3 = states 0-3, records popped last to first (mnflush anscdf_.h:128-138).
        // Records are fetched 32 at a time (one coalesced 128-byte load, issued one block ahead) and handed to the four
        // state lanes by shuffle, so the global-memory latency is off the per-record chain.
        uint8_t *slot = slots + u * slot_stride;
        const int cap = (int)slot_stride;
        int pos = cap;                                                         // lowest byte written so far (all lanes agree)
        uint32_t s = ANS_L;
        bool em = false, ovf = false;
        const unsigned k = lane & 3;
        const uint32_t nrec = 4 * npairs;
        int blk = (int)((nrec - 1) >> 5);                                      // 32-record block holding the last record
        uint32_t cur = (uint32_t)blk * 32 + lane < nrec ? rec[(uint32_t)blk * 32 + lane] : 0;
        uint32_t nxt = blk > 0 ? rec[(uint32_t)(blk - 1) * 32 + lane] : 0;
        for (int gi = (int)npairs - 1; gi >= 0 && !ovf; gi--) {
            if ((int)((4u * (uint32_t)gi) >> 5) != blk) { blk--; cur = nxt; nxt = blk > 0 ? rec[(uint32_t)(blk - 1) * 32 + lane] : 0; }
            const uint32_t r = __shfl_sync(0xffffffffu, cur, ((4u * (uint32_t)gi) & 31) + 3 - k);   // state k codes record 4g+3-k
            const uint32_t f = r & 0xffffu, c = r >> 16;
            const bool e = lane < 4 && s >= (f << 16);
            const unsigned bal = __ballot_sync(0xffffffffu, e) & 0xfu;
            if (e) { st_u16(slot + pos - 2 * (__popc(bal & ((1u << k) - 1)) + 1), s); s >>= 16; }
            pos -= 2 * __popc(bal);
            if (lane < 4) { const uint32_t q = s / f; s = s + (q << PROB_BITS) - q * f + c; }
            em = e;
            ovf = pos < 32;
        }
        // ansflush: st[0] highest ... st[3] lowest
        if (lane < 4) st_u32_a2(slot + pos - 4 * ((int)k + 1), s);
        pos -= 16;
        const bool em3 = __shfl_sync(0xffffffffu, (int)em, 3);                  // the last-coded record belongs to state 3
        if (lane == 0) {
            UnitMeta m;
            m.len = (uint32_t)(cap - pos); m.a_off = (uint32_t)pos; m.a_len = m.len; m.b_off = 0; m.b_len = 0;
            m.flags = (ovf ? UM_OVF : 0) | (em3 ? 0 : UM_ADJ2); m.pref = 0; m.pad = 0;
            meta[u] = m;
        }
        __syncwarp();
    }
}

// ---- decoder --------------------------------------------------------------------------------------------------
// one-table register cache: the entry a lane owns stays in `m` while consecutive symbols select the same table
struct TabCache {
    uint32_t id; int m;                                                        // id = entry offset of the cached table inside T
    __device__ __forceinline__ void select(uint16_t *T, uint32_t t, unsigned i) { if (t != id) { T[id + i] = (uint16_t)m; m = T[t + i]; id = t; } }   // warp-uniform
    __device__ __forceinline__ void flush(uint16_t *T, unsigned i) { T[id + i] = (uint16_t)m; }
};
// one nibble: cdf16ansdec (cdf_.h:52-59) + STATEUPD (cdf_.h:37); every lane returns the same x and updated state
__device__ __forceinline__ uint32_t coop_dec_nib(TabCache &c, unsigned i, unsigned lane, uint32_t &s) {
    const uint32_t r = s & PROB_MASK;
    const int m = c.m;
    const bool gt = (uint32_t)m > r;
    const unsigned bal = __ballot_sync(0xffffffffu, gt);
    const unsigned half = (bal >> (lane & 16)) & 0xffffu;                     // both half-warps hold the same table
    const unsigned x = (half ? (unsigned)__ffs((int)half) - 1 : 16u) - 1;     // first entry > r, minus one
    const int mx = __shfl_sync(0xffffffffu, m, (lane & 16) | x);
    int mx1 = __shfl_sync(0xffffffffu, m, (lane & 16) | ((x + 1) & 15));
    if (x == 15) mx1 = (int)PROB_TOTAL;
    s = (uint32_t)(mx1 - mx) * (s >> PROB_BITS) + r - (uint32_t)mx;
    c.m = adapt_entry(m, (int)i, gt);
    return x;
}

// the compressed stream as seen by a whole warp: every lane keeps the same byte window (<= 8 bytes) plus the next two
// 32-bit words already loaded, so a refill never waits on memory.  Any byte alignment of the stream start works.
struct WarpStream {
    const uint32_t *wp; const uint8_t *gend;
    uint64_t win; uint32_t have, nw0, nw1;                                     // have = valid bytes in win
    __device__ __forceinline__ uint32_t ldw(const uint32_t *p) const {
        if ((const uint8_t *)(p + 1) <= gend) return __ldg(p);
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) if ((const uint8_t *)p + k < gend) v |= (uint32_t)((const uint8_t *)p)[k] << (8 * k);
        return v;
    }
    __device__ __forceinline__ void append() { win |= (uint64_t)nw0 << (8 * have); have += 4; nw0 = nw1; nw1 = ldw(wp); wp++; }
    __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *end) {
        gend = end; wp = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
        const uint32_t skip = (uint32_t)((uintptr_t)p & 3);
        nw0 = ldw(wp); nw1 = ldw(wp + 1); wp += 2; win = 0; have = 0;
        append(); win >>= 8 * skip; have -= skip;
    }
    __device__ __forceinline__ uint32_t take16() { if (have < 2) append(); uint32_t v = (uint32_t)win & 0xffffu; win >>= 16; have -= 2; return v; }
    __device__ __forceinline__ uint32_t take32() { uint32_t a = take16(); return a | take16() << 16; }
    // address of the next unread byte (for the next block / consistency): wp counts words handed to nw*/win
    __device__ __forceinline__ const uint8_t *cursor() const { return (const uint8_t *)(wp - 2) - have; }
};

template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : COOP_WPB * 32)
k_ans_byte_dec_coop(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    extern __shared__ __align__(16) uint16_t smem_tabs[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15;
    uint16_t *T = smem_tabs + (size_t)wib * (O1 ? 256 * O1_CTX_ENTRIES : O1_CTX_ENTRIES);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint8_t *gend = in + in_off[g.n_calls];
    for (size_t j = gw; j < g.n_calls; j += nwarps) {
        size_t start, len; call_span(g, j, start, len);
        const uint64_t so = in_off[j], sl = in_off[j + 1] - so;
        uint8_t *op = out + start;
        if (sl == len) { group_copy(op, in + so, len, lane, 32); continue; }   // raw chunk (CCPY turborc.c:434)
        WarpStream ws; ws.init(in + so, gend);
        uint32_t cx = 0;                                                       // not reset per block (anscdf.c:629)
        for (size_t bp = 0; bp < len; bp += ANS_BLOCK) {
            const uint32_t n = (uint32_t)(len - bp < ANS_BLOCK ? len - bp : ANS_BLOCK), npairs = (n + 1) >> 1;
            uint8_t *bo = op + bp;
            __syncwarp();
            for (uint32_t k = lane; k < (O1 ? 256u * O1_CTX_ENTRIES : (uint32_t)O1_CTX_ENTRIES); k += 32) T[k] = (uint16_t)((k & 15) << 11);
            __syncwarp();
            uint32_t s0 = ws.take32(), s1 = ws.take32(), s2 = ws.take32(), s3 = ws.take32();   // mnfill anscdf_.h:176
            TabCache ch, cl;                                                   // high-nibble-type / low-nibble-type table caches
            ch.id = 0; ch.m = T[i]; cl.id = 16; cl.m = T[16 + i];
            for (uint32_t pi = 0; pi < npairs; pi++) {                         // mndec8x2 / mndec8x2x anscdf_.h:152-174
                const uint32_t c0 = O1 ? cx * O1_CTX_ENTRIES : 0;
                ch.select(T, c0, i);
                const uint32_t yh0 = coop_dec_nib(ch, i, lane, s0);
                cl.select(T, c0 + (1 + yh0) * 16, i);
                const uint32_t yl0 = coop_dec_nib(cl, i, lane, s1);
                const uint32_t x0 = yh0 << 4 | yl0;
                const uint32_t c1 = O1 ? x0 * O1_CTX_ENTRIES : 0;
                ch.select(T, c1, i);
                const uint32_t yh1 = coop_dec_nib(ch, i, lane, s2);
                cl.select(T, c1 + (1 + yh1) * 16, i);
                const uint32_t yl1 = coop_dec_nib(cl, i, lane, s3);
                const uint32_t x1 = yh1 << 4 | yl1;
                cx = x1;
                // ecdnorm x4 in state order (anscdf_.h:158-161); the states are replicated, so these branches are warp-uniform
                if (s0 < ANS_L) s0 = s0 << 16 | ws.take16();
                if (s1 < ANS_L) s1 = s1 << 16 | ws.take16();
                if (s2 < ANS_L) s2 = s2 << 16 | ws.take16();
                if (s3 < ANS_L) s3 = s3 << 16 | ws.take16();
                // bytes out (off the critical path): lane 0 stores the pair
                const uint32_t o = 2 * pi;
                if (lane == 0) {
                    if (o + 1 < n) {
                        if ((((uintptr_t)(bo + o)) & 1) == 0) *(uint16_t *)(bo + o) = (uint16_t)(x0 | x1 << 8);
                        else { bo[o] = (uint8_t)x0; bo[o + 1] = (uint8_t)x1; }
                    } else bo[o] = (uint8_t)x0;                                // odd tail: second byte discarded (anscdf.c:602)
                }
            }
            ch.flush(T, i); cl.flush(T, i);                                    // (tables are re-initialised for the next block anyway)
        }
    }
}

// ===============================================================================================================
// Adaptive byte range coders, warp-cooperative: TRC_RC (rccdfenc/rccdfdec rccdf.c:187-211, one coder) and TRC_RCI
// (rccdfienc/rccdfidec rccdf.c:213-249, coder 0 = high nibbles, coder 1 = low nibbles).
// Encoder: lanes 0-15 own the high-nibble table, lanes 16-31 the low-nibble table of the current byte (as in the
// rANS model pass); the coder state is replicated -- NC == 1: in all 32 lanes, which code (high, low) in sequence;
// NC == 2: coder 0 in lanes 0-15 and coder 1 in lanes 16-31, both nibbles of a byte coded in the same instructions.
// Only lanes 0 / 16 store.  Decoder: the 16-entry search of _cdflget16 (turborc_.h:271-291) is one 64-bit
// multiply-compare per lane + ballot + popcount.
// ===============================================================================================================

// scalar restatement with the carry walk-back coder, run by one lane when the fast coder flags a wrapped pending word
template <int NC>
__device__ __noinline__ void rc_byte_enc_serial(const uint8_t *ip, size_t n, uint8_t *slot, uint16_t *T, UnitMeta &m) {
    for (int k = 0; k < O1_CTX_ENTRIES; k++) T[k] = (uint16_t)((k & 15) << 11);
    const int64_t thr = rc_thr(n);
    const uint32_t b1ref = 4 + (uint32_t)(n / 2), b1 = (b1ref + 64 + 15) & ~15u;
    RcEnc e0, e1; e0.init(slot + (NC == 2 ? 4 : 16)); e1.init(slot + b1);
    bool raw = false;
    auto nib = [&](RcEnc &e, uint16_t *t, unsigned x) {
        uint32_t c = t[x], f = (x == 15 ? PROB_TOTAL : (uint32_t)t[x + 1]) - c;
        e.encode(c, f);
        for (int i = 0; i < 16; i++) t[i] = (uint16_t)adapt_entry(t[i], i, i > (int)x);
    };
    const size_t n4 = n & ~(size_t)3;
    for (size_t i = 0; i < n && !raw; i++) {
        const unsigned x = ip[i];
        nib(e0, T, x >> 4); nib(NC == 2 ? e1 : e0, T + (1 + (x >> 4)) * 16, x & 15);
        if (NC == 1) raw = (int64_t)e0.pos >= thr;
        else if (i < n4 && (i & 3) == 3) raw = (int64_t)b1ref + e1.pos >= thr || 4 + e0.pos >= b1ref;
    }
    m.pref = 0; m.pad = 0; m.b_off = 0; m.b_len = 0; m.a_off = NC == 2 ? 0 : 16;
    if (NC == 1) {
        if (!raw) e0.flush();
        m.a_len = raw ? 0 : e0.pos; m.len = raw ? (uint32_t)n : e0.pos;
    } else {
        if (!raw) { e0.flush(); e1.flush(); *(uint32_t *)slot = e0.pos; if ((int64_t)(4 + e0.pos + e1.pos) >= thr) raw = true; }
        m.a_len = raw ? 0 : 4 + e0.pos; m.b_off = b1; m.b_len = raw ? 0 : e1.pos; m.len = raw ? (uint32_t)n : 4 + e0.pos + e1.pos;
    }
    m.flags = raw ? UM_RAW : 0;
}

template <int NC>
__global__ void __launch_bounds__(COOP_WPB * 32)
k_rc_byte_enc_coop(const uint8_t *__restrict__ in, Geom g, uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta,
                   int force_redo) {
    extern __shared__ __align__(16) uint16_t smem_tabs[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
    uint16_t *T = smem_tabs + (size_t)wib * O1_CTX_ENTRIES;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    for (size_t j = gw; j < g.n_calls; j += nwarps) {
        size_t start, n; call_span(g, j, start, n);
        const uint8_t *ip = in + start;
        uint8_t *slot = slots + j * slot_stride;
        __syncwarp();
        for (uint32_t k = lane; k < (uint32_t)O1_CTX_ENTRIES; k += 32) T[k] = (uint16_t)((k & 15) << 11);
        __syncwarp();
        const int64_t thr = rc_thr(n);
        const uint32_t b1ref = 4 + (uint32_t)(n / 2), b1 = (b1ref + 64 + 15) & ~15u;   // rccdf.c:232
        RcE32 e; e.init(slot + (NC == 2 ? (h ? b1 : 4) : 16));
        const bool writer = NC == 2 ? i == 0 : lane == 0;
        uint16_t *cur_tab = T + (h ? 16 : 0);
        int m_cache = cur_tab[i];
        bool raw = false;
        const size_t n4 = n & ~(size_t)3;
        for (size_t base = 0; base < n && !raw; base += 32) {
            const uint32_t cnt = (uint32_t)(n - base < 32 ? n - base : 32);
            const uint32_t mine = lane < cnt ? ip[base + lane] : 0;
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t x = __shfl_sync(0xffffffffu, mine, k), yh = x >> 4, xs = h ? (x & 15) : yh;
                uint16_t *tab = T + (h ? (1 + yh) * 16 : 0);
                if (tab != cur_tab) { cur_tab[i] = (uint16_t)m_cache; m_cache = tab[i]; cur_tab = tab; }
                const int m = m_cache;
                const int mx = __shfl_sync(0xffffffffu, m, (lane & 16) | xs);
                int mx1 = __shfl_sync(0xffffffffu, m, (lane & 16) | ((xs + 1) & 15));
                if (xs == 15) mx1 = (int)PROB_TOTAL;
                m_cache = adapt_entry(m, (int)i, i > xs);                       // cdf16upd (cdf4e rccdf_.h:28)
                const uint32_t rec = (uint32_t)(mx1 - mx) | (uint32_t)mx << 16;
                if (NC == 2) e.encode_w(rec >> 16, rec & 0xffffu, writer);       // cdf8e2 rccdf_.h:36-40: both nibbles at once
                else {
                    const uint32_t rh = __shfl_sync(0xffffffffu, rec, 0), rl = __shfl_sync(0xffffffffu, rec, 16);
                    e.encode_w(rh >> 16, rh & 0xffffu, writer); e.encode_w(rl >> 16, rl & 0xffffu, writer);   // cdf8e rccdf_.h:30-34
                }
                if (NC == 2 && base + k + 1 == n4 && base + cnt > n4) {          // last in-loop OVERFLOWI of the reference
                    const uint32_t p0 = __shfl_sync(0xffffffffu, e.bytes(), 0), p1 = __shfl_sync(0xffffffffu, e.bytes(), 16);
                    raw = (int64_t)b1ref + p1 >= thr || 4 + p0 >= b1ref;
                    if (raw) break;
                }
            }
            // overflow tests are monotone in the cursors: once per 32 bytes decides like once per byte (NC == 1, OVERFLOW
            // rccdf.c:206).  NC == 2: OVERFLOWI (rccdf.c:240) is only evaluated inside the 4-byte loop, so blocks that lie
            // within n4 test at their end and the block that contains n4 tests exactly there (see the k loop).
            if (NC == 1) raw = (int64_t)e.bytes() >= thr;
            else if (base + cnt <= n4) {
                const uint32_t p0 = __shfl_sync(0xffffffffu, e.bytes(), 0), p1 = __shfl_sync(0xffffffffu, e.bytes(), 16);
                raw = (int64_t)b1ref + p1 >= thr || 4 + p0 >= b1ref;
            }
        }
        cur_tab[i] = (uint16_t)m_cache;
        UnitMeta m; m.pref = 0; m.pad = 0; m.b_off = 0; m.b_len = 0; m.a_off = NC == 2 ? 0 : 16;
        if (!raw) e.flush_w(writer);
        const uint32_t p0 = __shfl_sync(0xffffffffu, e.bytes(), 0), p1 = __shfl_sync(0xffffffffu, e.bytes(), 16);
        const uint32_t rare = __shfl_sync(0xffffffffu, e.rare, 0) | __shfl_sync(0xffffffffu, e.rare, 16);
        if (NC == 1) { m.a_len = raw ? 0 : p0; m.len = raw ? (uint32_t)n : p0; }
        else {
            if (!raw) { if (lane == 0) *(uint32_t *)slot = p0; if ((int64_t)(4 + p0 + p1) >= thr) raw = true; }   // rccdf.c:246
            m.a_len = raw ? 0 : 4 + p0; m.b_off = b1; m.b_len = raw ? 0 : p1; m.len = raw ? (uint32_t)n : 4 + p0 + p1;
        }
        m.flags = raw ? UM_RAW : 0;
        __syncwarp();
        if (lane == 0) {
            if ((rare && !raw) || force_redo) rc_byte_enc_serial<NC>(ip, n, slot, T, m);
            meta[j] = m;
        }
        __syncwarp();
    }
}

// decoder-side coder state, replicated in every lane
struct RcDW {
    uint64_t range, code;
    __device__ __forceinline__ void init(WarpStream &ws) { range = ~0ull; uint32_t a = ws.take32(), b = ws.take32(); code = (uint64_t)a << 32 | b; }   // rcdinit
    // one nibble against the cached table entry m (entry i of the table, i = lane & 15)
    __device__ __forceinline__ uint32_t nib(TabCache &c, unsigned i, unsigned lane, WarpStream &ws) {
        range >>= PROB_BITS;
        const int m = c.m;
        const bool le = i != 0 && (uint64_t)(uint32_t)m * range <= code;        // entries 1..15: cdf[e]*range <= code
        const unsigned bal = __ballot_sync(0xffffffffu, le);
        const unsigned x = __popc((bal >> (lane & 16)) & 0xffffu);               // monotone in e: the count is the symbol
        const uint32_t c0 = (uint32_t)__shfl_sync(0xffffffffu, m, (lane & 16) | x);
        uint32_t c1 = (uint32_t)__shfl_sync(0xffffffffu, m, (lane & 16) | ((x + 1) & 15));
        if (x == 15) c1 = PROB_TOTAL;
        const uint64_t rp = (uint64_t)c0 * range;                                // _rccdfupdate turborc_.h:219-229
        range = range * (c1 - c0); code -= rp;
        if ((uint32_t)(range >> 32) == 0) { range <<= 32; code = code << 32 | ws.take32(); }
        c.m = adapt_entry(m, (int)i, i > x);
        return x;
    }
};

template <int NC>
__global__ void __launch_bounds__(COOP_WPB * 32)
k_rc_byte_dec_coop(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    extern __shared__ __align__(16) uint16_t smem_tabs[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15;
    uint16_t *T = smem_tabs + (size_t)wib * O1_CTX_ENTRIES;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint8_t *gend = in + in_off[g.n_calls];
    for (size_t j = gw; j < g.n_calls; j += nwarps) {
        size_t start, n; call_span(g, j, start, n);
        const uint64_t so = in_off[j], sl = in_off[j + 1] - so;
        uint8_t *op = out + start;
        const uint8_t *stream = in + so;
        if (sl == n) { group_copy(op, stream, n, lane, 32); continue; }
        __syncwarp();
        for (uint32_t k = lane; k < (uint32_t)O1_CTX_ENTRIES; k += 32) T[k] = (uint16_t)((k & 15) << 11);
        __syncwarp();
        WarpStream w0, w1;
        RcDW d0, d1;
        if (NC == 1) { w0.init(stream, gend); d0.init(w0); }
        else {
            const uint32_t len0 = ld_u32_clamped(stream, gend);
            const uint8_t *p1 = stream + 4 + len0;
            if (p1 > gend || p1 < stream) p1 = gend;
            w0.init(stream + 4, gend); d0.init(w0); w1.init(p1, gend); d1.init(w1);
        }
        TabCache ch, cl;
        ch.id = 0; ch.m = T[i]; cl.id = 16; cl.m = T[16 + i];
        for (size_t k = 0; k < n; k++) {                                         // cdf8d / cdf8d2 rccdf_.h:50-73
            const uint32_t yh = d0.nib(ch, i, lane, w0);
            cl.select(T, (1 + yh) * 16, i);
            const uint32_t yl = NC == 1 ? d0.nib(cl, i, lane, w0) : d1.nib(cl, i, lane, w1);
            if (lane == 0) op[k] = (uint8_t)(yh << 4 | yl);
        }
    }
}

}  // namespace trc
