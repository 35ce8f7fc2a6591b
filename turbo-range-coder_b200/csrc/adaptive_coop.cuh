// adaptive_coop.cuh -- warp-cooperative ENCODER of the adaptive byte range coders TRC_RC / TRC_RCI (rccdfenc / rccdfienc,
// rccdf.c:201-249) for batches of fewer than COOP_MIN_LANE_UNITS calls, plus the constants shared with adaptive_v3.cuh.
// One WARP owns one call: a 16-entry CDF lives one entry per lane, so cdf16upd (cdf_.h:46-50) is ONE data-parallel step
// instead of 16 serial ones; the coder state is replicated in every lane.  (The first-generation warp-cooperative rANS
// kernels and range decoder that used to live here were superseded by adaptive_v3.cuh and removed in round 2.)
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


}  // namespace trc
