// rans_wide.cuh -- TRC_ANSW: 32-way warp-interleaved static rANS, one 32-bit state per lane, ONE stream per call.
//
// This is the layout BASELINE.json's north star describes ("32- to 128-way warp-interleaved rANS with one 32-bit
// state per lane ... warp ballot + prefix-sum compact each lane's variable-length renorm output into a dense byte
// stream").  It is a NEW format of this repository: no reference codec writes or reads it (the reference's rANS
// streams are 2- or 4-way), so its parity is UNPINNED -- oracle/trc_oracle.c (orc_answenc/orc_answdec) is the format
// specification, tests check bit-equality with that specification, exact round trip and the size bound.  The
// per-symbol arithmetic is the reference's own: ece (anscdf_.h:90-94), STATEUPD (cdf_.h:37), ecdnorm (anscdf_.h:50-73),
// 15-bit CDF, 16-bit renormalisation words.
//
// A warp owns a call.  Symbol i belongs to lane (i / 4) % 32: every 128-symbol super-group is one coalesced 32-bit
// load (encoder) / store (decoder) per lane.  In each of the 4 steps of a super-group the lanes that renormalise emit
// one 16-bit word each; ballot + popcount gives every lane its rank, so the words of a step are contiguous and in lane
// order.  The decoder never touches memory on the per-symbol chain: the next 128+ halfwords of the stream sit in three
// registers per lane (two live windows + one in flight) and a lane picks its word with two shuffles.
#pragma once
#include "trc_common.cuh"
#include "rans_static.cuh"
#include "static_v2.cuh"

namespace trc {

constexpr int ANSW_WPB = 4;                       // warps (calls) per CTA
constexpr uint32_t ANSW_HDR = 32 * 4;             // 32 final states

__global__ void __launch_bounds__(ANSW_WPB * 32)
k_answ_enc(const uint8_t *__restrict__ in, Geom g, const TableSet *__restrict__ ts, size_t cpc,
           uint8_t *__restrict__ slots, size_t slot_stride, UnitMeta *__restrict__ meta) {
    __shared__ __align__(16) uint4 etab[256];
    __shared__ uint64_t bar;
    const unsigned lane = threadIdx.x & 31;
    const size_t j0 = (size_t)blockIdx.x * ANSW_WPB, j = j0 + (threadIdx.x >> 5);
    const TableSet *t = ts + (cpc ? j0 / cpc : 0);
    if (threadIdx.x == 0) tma_fetch(etab, t->etab, sizeof etab, &bar);
    __syncthreads();
    tma_wait(&bar);
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    const uint8_t *ip = in + start;
    const uint32_t n = (uint32_t)len, ng = (n + 127) >> 7;
    uint8_t *slot = slots + j * slot_stride;
    const int cap = (int)slot_stride;
    int pos = cap;                                                             // identical in all lanes
    uint32_t s = ANS_L;
    const unsigned higher = (0xffffffffu << lane) << 1;                        // lanes above this one
    auto load_group = [&](uint32_t gi) -> uint32_t {
        const uint32_t i0 = gi * 128 + lane * 4;
        if (i0 + 4 <= n) return __ldg((const uint32_t *)(ip + i0));            // calls start 4-byte aligned (chunk_len % 4 == 0)
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) if (i0 + k < n) v |= (uint32_t)ip[i0 + k] << (8 * k);
        return v;
    };
    bool raw = false;
    uint32_t w = ng ? load_group(ng - 1) : 0;
    for (uint32_t gi = ng; gi-- > 0 && !raw;) {
        const uint32_t wn = gi ? load_group(gi - 1) : 0;                       // next (lower) super-group, in flight during this one
        const uint32_t i0 = gi * 128 + lane * 4;
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            const bool active = i0 + k < n;
            const uint4 e = etab[(w >> (8 * k)) & 0xff];
            const bool emit = active && s >= e.y;                              // ecenorm anscdf_.h:48
            const unsigned bal = __ballot_sync(0xffffffffu, emit);
            if (emit) st_u16(slot + pos - 2 * (__popc(bal & higher) + 1), s);  // words of a step: lane order, ascending addresses
            pos -= 2 * __popc(bal);
            s = emit ? s >> 16 : s;
            if (active) { const uint32_t q = __umulhi(s, e.x) >> (e.z >> 16); s = s + e.w + q * (e.z & 0xffffu); }   // ece anscdf_.h:90-94
        }
        raw = (uint32_t)(cap - pos) + ANSW_HDR >= n;                           // stream can no longer be shorter than the input
        w = wn;
    }
    pos -= (int)ANSW_HDR;                                                      // states: lane 0 lowest
    if (!raw) st_u32_a2(slot + pos + 4 * (int)lane, s);
    const uint32_t l = (uint32_t)(cap - pos);
    raw = raw || l >= n;
    if (lane == 0) {
        UnitMeta m;
        m.len = raw ? n : l; m.a_off = (uint32_t)pos; m.a_len = raw ? 0 : l; m.b_off = 0; m.b_len = 0;
        m.flags = raw ? UM_RAW : 0; m.pref = 0; m.pad = 0;
        meta[j] = m;
    }
}

__global__ void __launch_bounds__(ANSW_WPB * 32)
k_answ_dec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
           const TableSet *__restrict__ ts, size_t cpc) {
    __shared__ __align__(16) uint32_t dtab[256];
    __shared__ __align__(16) uint8_t lut[PROB_TOTAL];
    __shared__ uint64_t bar;
    const unsigned lane = threadIdx.x & 31;
    const size_t j0 = (size_t)blockIdx.x * ANSW_WPB, j = j0 + (threadIdx.x >> 5);
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
    if (j >= g.n_calls) return;
    size_t start, len; call_span(g, j, start, len);
    const uint64_t so = in_off[j], sl = in_off[j + 1] - so;
    uint8_t *op = out + start;
    const uint8_t *stream = in + so, *gend = in + in_off[g.n_calls];
    if (sl == len) { group_copy(op, stream, len, lane, 32); return; }
    const uint32_t n = (uint32_t)len, ng = (n + 127) >> 7;
    uint32_t s = ld_u32_clamped(stream + 4 * lane, gend);                      // state of lane l
    // halfword windows over the word stream: window k covers 32-bit words [32k, 32k+32) counted from wbase
    const uint8_t *wstart = stream + ANSW_HDR;
    const uint32_t *wbase = (const uint32_t *)((uintptr_t)wstart & ~(uintptr_t)3);
    uint32_t c = (uint32_t)(((uintptr_t)wstart & 2) >> 1);                     // cursor: halfword index inside window 0
    auto load_win = [&](uint32_t k) -> uint32_t {
        const uint32_t *p = wbase + 32 * k + lane;
        if ((const uint8_t *)(p + 1) <= gend) return __ldg(p);
        return ((const uint8_t *)p + 2 <= gend) ? (uint32_t)__ldg((const uint16_t *)p) : 0u;
    };
    uint32_t b0 = load_win(0), b1 = load_win(1), b2 = load_win(2), b3 = load_win(3), nextk = 4;   // b0,b1 live; b2,b3 in flight
    const unsigned lower = (1u << lane) - 1;
    for (uint32_t gi = 0; gi < ng; gi++) {
        const uint32_t i0 = gi * 128 + lane * 4;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool active = i0 + k < n;
            const uint32_t r = s & PROB_MASK, x = lut[r], e = dtab[x];
            if (active) s = (e & 0xffffu) * (s >> PROB_BITS) + r - (e >> 16);  // STATEUPD cdf_.h:37
            acc |= (active ? x : 0u) << (8 * k);
            const bool need = active && s < ANS_L;                             // ecdnorm anscdf_.h:50-73
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            const uint32_t idx = c + __popc(bal & lower);                      // my halfword: lower lanes read first
            const uint32_t v0 = __shfl_sync(0xffffffffu, b0, (idx >> 1) & 31), v1 = __shfl_sync(0xffffffffu, b1, (idx >> 1) & 31);
            const uint32_t v = idx < 64 ? v0 : v1, hw = (idx & 1) ? v >> 16 : v & 0xffffu;
            s = need ? (s << 16 | hw) : s;
            c += __popc(bal);
            if (c >= 64) { c -= 64; b0 = b1; b1 = b2; b2 = b3; b3 = load_win(nextk++); }   // warp-uniform; a new window is two rotations away from use
        }
        if (i0 + 4 <= n) *(uint32_t *)(op + i0) = acc;                          // coalesced 128 bytes per warp
        else for (int k = 0; k < 4; k++) if (i0 + k < n) op[i0 + k] = (uint8_t)(acc >> (8 * k));
    }
}

}  // namespace trc
