// adaptive_v3.cuh -- third generation of the adaptive byte rANS kernels (TRC_ANS anscdfenc/anscdfdec anscdf.c:567-605,
// TRC_ANS1 order-1 anscdf.c:607-645) for batches of few, large units (64 KiB ... 4 MiB chunks, every drop-in call).
// Same bytes as adaptive.cuh / adaptive_coop.cuh; rebuilt around what bounds these codecs on a GPU: the length of
// the dependent chain per byte and the instructions a warp has to issue per byte.
//
//   k_ans_model3  (encoder, model pass; one warp per unit: lanes 0-15 own the high-nibble table of the current byte,
//                  lanes 16-31 its low-nibble table, one CDF entry per lane)
//       * cdf16upd (cdf_.h:46-50) per entry is  m' = (127 m + 10 i + (i > x ? 32736 : 0)) >> 7  -- algebraically the
//         reference's  m += (T - m) >> 7  (floor division by 128 of a non-negative sum), i.e. ONE multiply-add and one
//         shift on the dependent chain;
//       * tables are write-through in shared memory; the table entry of byte t+2 is loaded while byte t is being
//         updated and two selects forward the fresh entries when bytes t / t+1 hit the same table, so shared-memory
//         latency is off the chain;
//       * the (freq | cum << 16) record of a nibble (mnenc4 anscdf_.h:106) is produced by the lane that owns the coded
//         symbol with one predicated 4-byte store (global stores do not order against the shared-memory table traffic,
//         so the record never sits on the chain).
//   k_ans_code3   (encoder, coding pass mnflush anscdf_.h:128-138; one LANE per (unit, rANS state): a warp runs 8
//                  units x 4 states, records popped last to first)
//       * division by the adaptive frequency is an exact multiply-high with a reciprocal taken from a 32 K-entry table
//         (128 KB, staged in shared memory by TMA bulk copies) -- no integer division on the chain;
//       * the words the four states of a unit emit in one step are ordered with ballot + popcount (LIFO order of the
//         reference); records and reciprocals are software-pipelined 16-32 steps ahead in registers.
//   k_ans_dec3    (decoder; order 0: one HALF-warp per call, two calls per warp; order 1: one warp per call, both halves
//                  replicate, 136 KB of tables in the shared memory of one SM)
//       * symbol search of cdf16ansdec (cdf_.h:52-59) = compare + ballot + popcount; every lane computes the state
//         update for ITS entry before the symbol is known and one shuffle picks the right one; tables carry the
//         reference's 17th entry so a lane reads its own and the next entry with two shared loads issued together;
//       * the stream is staged through a 64-halfword shared-memory ring per call (any byte alignment, refilled one
//         period ahead), so the four ecdnorm steps (anscdf_.h:50-73) of a byte pair are four speculative 16-bit
//         shared loads with predicated merges -- no branches in the pair loop.
#pragma once
#include "trc_common.cuh"
#include "adaptive.cuh"
#include "static_v2.cuh"
#include "adaptive_coop.cuh"

namespace trc {

constexpr unsigned FULLMASK = 0xffffffffu;

// ================================================================================================================
// encoder, model pass
// ================================================================================================================
constexpr int M3_WPB = 4;                                        // warps (units) per CTA, order 0
template <bool O1> __host__ __device__ constexpr uint32_t m3_warp_bytes() { return (O1 ? 256u : 1u) * O1_CTX_ENTRIES * 2u; }

// Pipeline registers of one lane.  Byte t is being coded; the table entries of bytes t+1 and t+2 are already on their
// way from shared memory.  `a*` are BYTE offsets of this lane's entry inside the warp's table block.
struct M3State {
    uint32_t x0, x1;         // bytes t, t+1 (pre-shifted left by 1: table offsets come out with one AND)
    uint32_t a0, a1;         // entry address of byte t / t+1
    uint32_t ap;             // entry address of byte t-1
    int m;                   // entry value for byte t (up to date)
    int m2p;                 // value written for byte t-1
    int pre1;                // entry loaded for byte t+1 (lacks the updates of bytes t-1 and t if they hit the same table)
    int mp, dnp; bool wp;    // record of byte t-1, still to be stored: entry, next entry (shuffle in flight), "this lane owns it"
};

__device__ __forceinline__ uint32_t m3_record(int m, int dn, unsigned i) {           // (next - m) | m << 16  (mnenc4 anscdf_.h:106)
    return (uint32_t)m * 65535u + (uint32_t)(i == 15 ? (int)PROB_TOTAL : dn);
}

// one byte: cdf16upd of this lane's entry, hand-over to byte t+1, and the record of the PREVIOUS byte (its shuffle was
// issued a step ago, so nothing waits on it).  x2 = byte t+2 (<< 1); rec_slot = record slot of byte t.
template <bool O1>
__device__ __forceinline__ void m3_step(M3State &s, uint32_t x2, uint8_t *Tb, uint32_t *rec_slot, unsigned i, uint32_t hmask, uint32_t xsh,
                                        int c10, int c10mix) {
    // table of byte t+2: mbh[cx] (lanes 0-15) / mbl[cx][x >> 4] (lanes 16-31), cx = byte t+1
    const uint32_t a2 = (O1 ? (s.x1 >> 1) * (uint32_t)(O1_CTX_ENTRIES * 2) : 0u) + (x2 & hmask);
    const int pre2 = *(const uint16_t *)(Tb + a2);               // issued two bytes ahead of its use
    const uint32_t xs = (s.x0 >> xsh) & 15u;
    const int dn = __shfl_down_sync(FULLMASK, s.m, 1, 16);
    if (s.wp) rec_slot[-2] = m3_record(s.mp, s.dnp, i);
    s.mp = s.m; s.dnp = dn; s.wp = i == xs;
    const int m2 = (127 * s.m + (i > xs ? c10mix : c10)) >> 7;   // cdf16upd
    *(uint16_t *)(Tb + s.a0) = (uint16_t)m2;
    const int q1 = s.a1 == s.ap ? s.m2p : s.pre1;                // byte t-1 hit the table of byte t+1
    s.m = s.a1 == s.a0 ? m2 : q1;                                // byte t did
    s.m2p = m2; s.ap = s.a0; s.a0 = s.a1; s.a1 = a2; s.pre1 = pre2; s.x0 = s.x1; s.x1 = x2;
}

template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : M3_WPB * 32)
k_ans_model3(const uint8_t *__restrict__ in, Geom g, uint32_t *__restrict__ recs, size_t rec_stride) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
    constexpr uint32_t NENT = (O1 ? 256u : 1u) * O1_CTX_ENTRIES;
    uint16_t *T = (uint16_t *)(smem_raw + (size_t)wib * m3_warp_bytes<O1>());
    uint8_t *Tb = (uint8_t *)T + (h ? 32u : 0u) + 2u * i;        // this lane's entry of table 0 (high) / table 1 (first low table)
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int c10 = ADAPT_IC_ * (int)i, c10mix = c10 + (int)AD_MIX;
    const uint32_t xsh = h ? 1u : 5u;                            // bytes are kept << 1
    const uint32_t hmask = h ? 0x1e0u : 0u;                      // (x << 1) & 0x1e0 = 32 * (x >> 4): byte offset of mbl[x >> 4]
    for (size_t u = gw; u < g.n_units; u += nwarps) {
        size_t j, start, len; uint32_t b;
        unit_span(g, u, j, b, start, len);
        if (len == 0) continue;                                   // padding unit: k_ans_code3 writes its (empty) meta
        const uint8_t *ip = in + start;
        const uint32_t n = (uint32_t)len, nb = (n + 1) & ~1u;     // odd tail: a dummy 0 byte is coded too (anscdf.c:581,621)
        uint32_t *rec = recs + u * rec_stride + h;
        __syncwarp();
        for (uint32_t k = lane; k < NENT; k += 32) T[k] = (uint16_t)((k & 15) << 11);   // CDF16DEC0/1/2 cdf_.h:26-32
        __syncwarp();
        const uint32_t cx0 = (O1 && start > j * g.chunk) ? in[start - 1] : 0;           // cx carries across blocks (anscdf.c:608)
        // input bytes: one per lane and 32-byte block, loaded three blocks ahead; the << 1 is applied a block before use so
        // that nothing touches a register a DRAM load is still filling
        auto ldb = [&](uint32_t idx) -> uint32_t { return idx < n ? (uint32_t)__ldg(ip + idx) : 0u; };
        uint32_t mine = ldb(lane) << 1, mine_n = ldb(32 + lane) << 1, mine_n2 = ldb(64 + lane);
        M3State s;
        s.x0 = __shfl_sync(FULLMASK, mine, 0); s.x1 = __shfl_sync(FULLMASK, mine, 1);
        s.a0 = (O1 ? cx0 * (uint32_t)(O1_CTX_ENTRIES * 2) : 0u) + (s.x0 & hmask);
        s.a1 = (O1 ? (s.x0 >> 1) * (uint32_t)(O1_CTX_ENTRIES * 2) : 0u) + (s.x1 & hmask);
        s.ap = 0xffffffffu; s.m2p = 0; s.mp = s.dnp = 0; s.wp = false;
        s.m = *(const uint16_t *)(Tb + s.a0);
        s.pre1 = *(const uint16_t *)(Tb + s.a1);
        for (uint32_t base = 0; base < nb; base += 32) {
            const uint32_t mine_n3 = ldb(base + 96 + lane);       // three blocks ahead: covers a DRAM miss
            const uint32_t cnt = nb - base < 32 ? nb - base : 32;
            if (cnt == 32) {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    const uint32_t x2 = k + 2 < 32 ? __shfl_sync(FULLMASK, mine, (k + 2) & 31) : __shfl_sync(FULLMASK, mine_n, (k + 2) & 31);
                    m3_step<O1>(s, x2, Tb, rec + 2 * (base + k), i, hmask, xsh, c10, c10mix);
                }
            } else {
                for (uint32_t k = 0; k < cnt; k++) {              // bytes past the end only feed table prefetches
                    const uint32_t x2 = __shfl_sync(FULLMASK, mine, (k + 2) & 31);
                    m3_step<O1>(s, x2, Tb, rec + 2 * (base + k), i, hmask, xsh, c10, c10mix);
                }
            }
            mine = mine_n; mine_n = mine_n2 << 1; mine_n2 = mine_n3;
        }
        if (s.wp) rec[2 * (nb - 1)] = m3_record(s.mp, s.dnp, i);  // the last byte's record
    }
}

// ================================================================================================================
// encoder, coding pass
// ================================================================================================================
constexpr int C3_WPB = 2;                                        // warps per CTA (they share the reciprocal table)
constexpr int C3_B = 16;                                         // steps per software-pipeline block
constexpr uint32_t C3_LUT_BYTES = PROB_TOTAL * 4;                // 128 KB

__device__ uint32_t g_rcp_lut[PROB_TOTAL];                       // [f - 1] = reciprocal of rans_enc_entry(., f), f = 1 .. 2^15

// predicated 16-bit store at p + OFF (kept as ONE predicated instruction: a branch around it costs more than the store)
__device__ __forceinline__ void st_u16_if(uint8_t *p, int off, uint32_t v, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.global.u16 [%0 + %1], %2; }" ::"l"(p), "n"(-2), "h"((uint16_t)v), "r"((uint32_t)pred) : "memory");
    (void)off;
}

__global__ void k_build_rcp() {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < PROB_TOTAL) g_rcp_lut[k] = rans_enc_entry(0, k + 1).x;
}

// Frequencies of the adaptive tables never drop below 9: adjacent entries start 2048 apart, their targets are at least
// IC = 10 apart, and  m' = floor((127 m + T) / 128)  keeps a gap >= 10 (floor(a) - floor(b) >= floor(a - b) =
// floor((127 d + 10) / 128) >= 10 for d >= 10); the top entry never exceeds 32759.  So the f == 1 special case of
// rans_enc_entry cannot occur here (tests/test_adaptive_invariants.py replays the argument numerically).
__global__ void __launch_bounds__(C3_WPB * 32)
k_ans_code3(Geom g, const uint32_t *__restrict__ recs, size_t rec_stride, uint8_t *__restrict__ slots, size_t slot_stride,
            UnitMeta *__restrict__ meta) {
    extern __shared__ __align__(16) uint32_t rcp_s[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        const uint32_t b = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(C3_LUT_BYTES) : "memory");
        for (uint32_t part = 0; part < 4; part++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(rcp_s) + part * (C3_LUT_BYTES / 4)), "l"((const uint8_t *)g_rcp_lut + part * (C3_LUT_BYTES / 4)),
                           "r"(C3_LUT_BYTES / 4), "r"(b) : "memory");
    }
    __syncthreads();
    tma_wait(&bar);
    const unsigned lane = threadIdx.x & 31, k = lane & 3, gb = lane & 28;
    const unsigned gmask = 15u << gb, lmask = ((1u << k) - 1u) << gb;       // my unit's four lanes / those of lower states
    const size_t gw = (size_t)blockIdx.x * C3_WPB + (threadIdx.x >> 5);
    const size_t u = gw * 8 + (lane >> 2);
    size_t j, start, len = 0; uint32_t blk;
    if (u < g.n_units) unit_span(g, u, j, blk, start, len);
    const bool live = len != 0;
    const uint32_t npairs = (uint32_t)((len + 1) >> 1);
    const uint32_t tmax = __reduce_max_sync(FULLMASK, npairs);
    const uint32_t lead = tmax - npairs;                         // shorter units idle FIRST, so every unit's last step is the warp's last
    const uint32_t *rec = recs + (live ? u : 0) * rec_stride;
    uint8_t *slot = slots + (live ? u : 0) * slot_stride;
    const int cap = (int)slot_stride;
    int pos = cap;                                               // lowest byte written so far (same in the 4 lanes of a unit)
    uint32_t s = ANS_L;
    bool ovf = false, pp = false;                                // pp / pbal / pword: emission of the previous step, still pending
    unsigned pbal = 0; uint32_t pword = 0;
    // step t codes record 4 (npairs-1-(t-lead)) + 3 - k on state k (pushed 3,2,1,0 per byte pair -> popped 0,1,2,3).
    // Idle steps use the record (f = 2^15, cum = 0): it never renormalises and maps every state to itself.
    auto ldrec = [&](uint32_t t) -> uint32_t { return (t >= lead && t < tmax) ? __ldg(rec + 4 * (size_t)(tmax - 1 - t) + 3 - k) : PROB_TOTAL; };
    auto ldblock = [&](uint32_t tb, uint32_t (&o)[C3_B]) {      // records of steps tb .. tb+15; whole-block case: one pointer, fixed offsets
        if (tb >= lead && tb + C3_B <= tmax) {
            const uint32_t *p = rec + 4 * (size_t)(tmax - 1 - tb) + 3 - k;
#pragma unroll
            for (int q = 0; q < C3_B; q++) o[q] = __ldg(p - 4 * q);
        } else {
#pragma unroll
            for (int q = 0; q < C3_B; q++) o[q] = ldrec(tb + q);
        }
    };
    // Records and reciprocals travel through three register blocks whose roles rotate (A = being coded, B = next, its
    // reciprocals being fetched, C = being loaded), written out three times so that no register that a load is still
    // filling is ever moved; further ahead, the record lines are pulled into L2 by prefetches.
    auto prefetch = [&](uint32_t tb) {
        if (tb >= lead && tb < tmax) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 4 * (size_t)(tmax - 1 - tb) + 3 - k));
    };
    auto lut = [&](const uint32_t (&r)[C3_B], uint32_t (&q)[C3_B]) {
#pragma unroll
        for (int x = 0; x < C3_B; x++) q[x] = rcp_s[((r[x] & 0xffffu) - 1) & PROB_MASK];
    };
    auto code_block = [&](const uint32_t (&R)[C3_B], const uint32_t (&Q)[C3_B]) {
        // a block emits at most 16 steps x 4 states x 2 bytes: one slot check per block.  An exhausted slot (the unit is
        // then certainly raw, see make_plan) restarts at the top so the stores stay inside it.
        if (pos < 32 + 8 * (C3_B + 1)) { ovf = true; pos = cap; }
#pragma unroll
        for (int q = 0; q < C3_B; q++) {
            const uint32_t f = R[q] & 0xffffu;
            const uint32_t sh = 31 - __clz((int)((f - 1) | 1));                       // ceil(log2 f) - 1
            const bool p = s >= (R[q] << 16);                                         // ecenorm anscdf_.h:48: s >= f << 16
            const unsigned bal = __ballot_sync(FULLMASK, p);
            const uint32_t word = s;
            const uint32_t s1 = p ? s >> 16 : s;
            const uint32_t qq = __umulhi(s1, Q[q]) >> sh;                             // == s1 / f
            s = s1 + (R[q] >> 16) + qq * (PROB_TOTAL - f);                            // (q << 15) + s1 % f + cum
            // the word of the PREVIOUS step leaves now: its popcount / address / store chain overlaps this step's state chain
            st_u16_if(slot + (pos - 2 * (int)__popc(pbal & lmask)), -2, pword, pp);   // state 0's word highest (LIFO of mnflush)
            pos -= 2 * (int)__popc(pbal & gmask);
            pbal = bal; pword = word; pp = p;
        }
    };
    uint32_t R0[C3_B], R1[C3_B], R2[C3_B], Q0[C3_B], Q1[C3_B], Q2[C3_B];
    ldblock(0, R0); ldblock(C3_B, R1);
    for (uint32_t tb = 2 * C3_B; tb < 8 * C3_B; tb += C3_B) prefetch(tb);
    lut(R0, Q0);
    for (uint32_t t0 = 0; t0 < tmax;) {
        ldblock(t0 + 2 * C3_B, R2); prefetch(t0 + 8 * C3_B); lut(R1, Q1); code_block(R0, Q0); t0 += C3_B;
        if (t0 >= tmax) break;
        ldblock(t0 + 2 * C3_B, R0); prefetch(t0 + 8 * C3_B); lut(R2, Q2); code_block(R1, Q1); t0 += C3_B;
        if (t0 >= tmax) break;
        ldblock(t0 + 2 * C3_B, R1); prefetch(t0 + 8 * C3_B); lut(R0, Q0); code_block(R2, Q2); t0 += C3_B;
    }
    st_u16_if(slot + (pos - 2 * (int)__popc(pbal & lmask)), -2, pword, pp);           // the last step's word
    pos -= 2 * (int)__popc(pbal & gmask);
    const bool em = pp;
    if (pos < 32) { ovf = true; pos = cap; }
    if (live) st_u32_a2(slot + pos - 4 * ((int)k + 1), s);                            // ansflush: st[0] highest ... st[3] lowest
    pos -= 16;
    const bool em3 = __shfl_sync(FULLMASK, (int)em, gb | 3);                          // the last-coded record belongs to state 3
    if (k == 0 && u < g.n_units) {
        UnitMeta m;
        m.len = m.a_off = m.a_len = m.b_off = m.b_len = m.flags = m.pref = m.pad = 0;
        if (live) {
            m.len = (uint32_t)(cap - pos); m.a_off = (uint32_t)pos; m.a_len = m.len;
            m.flags = (ovf ? UM_OVF : 0) | (em3 ? 0 : UM_ADJ2);
        }
        meta[u] = m;
    }
}

// ================================================================================================================
// decoder
// ================================================================================================================
constexpr int D3_WPB = 4;                                        // warps per CTA, order 0 (8 calls)
constexpr uint32_t D3_RING = 64;                                 // halfwords per ring; 4 more mirror the first 4
constexpr uint32_t D3_RING_BYTES = (D3_RING + 4) * 2;
// order-1 decoder tables carry a 17th entry (= 32768, never written) like the reference's cdf[17], so "the next entry" is a
// plain shared-memory read issued together with the entry itself instead of a shuffle that has to wait for it
template <bool O1> __host__ __device__ constexpr uint32_t d3_stride() { return O1 ? 17u : 16u; }              // entries per table
template <bool O1> __host__ __device__ constexpr uint32_t d3_ctx_entries() { return 17u * d3_stride<O1>(); }  // 17 tables per context
template <bool O1> __host__ __device__ constexpr uint32_t d3_smem_bytes() {
    return O1 ? 256u * d3_ctx_entries<true>() * 2u + 2u * D3_RING_BYTES : D3_WPB * 2u * (d3_ctx_entries<false>() * 2u + D3_RING_BYTES);
}

// aligned word with bytes at or past `end` read as zero (never dereferences a word that starts at or past `end`)
__device__ __forceinline__ uint32_t ldw_clamped(const uint32_t *p, const uint8_t *end) {
    if ((const uint8_t *)(p + 1) <= end) return __ldg(p);
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) if ((const uint8_t *)p + k < end) v |= (uint32_t)((const uint8_t *)p)[k] << (8 * k);
    return v;
}
// four stream bytes at any alignment
__device__ __forceinline__ uint32_t ld32_any(const uint8_t *p, const uint8_t *end) {
    const uint32_t *w = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    const unsigned sh = (unsigned)((uintptr_t)p & 3) * 8;
    const uint32_t lo = ldw_clamped(w, end), hi = sh ? ldw_clamped(w + 1, end) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// one nibble: cdf16ansdec (cdf_.h:52-59) + STATEUPD (cdf_.h:37) + cdf16upd; every lane of the unit gets the same result.
// m / nx = this lane's entry and the next one.  Returns cnt = symbol + 1 (the number of entries <= r); m becomes the updated
// entry.  hbm1 = (lane & 16) - 1.
__device__ __forceinline__ uint32_t d3_nib(uint32_t &s, int &m, int nx, unsigned hbm1, unsigned hm, int c10, int c10mix) {
    const uint32_t r = s & PROB_MASK;
    const bool le = (uint32_t)m <= r;                             // entry 0 == 0: always true
    const uint32_t cand = (uint32_t)(nx - m) * (s >> PROB_BITS) + r - (uint32_t)m;   // the new state if this lane's entry is the symbol
    const unsigned cnt = __popc(__ballot_sync(FULLMASK, le) & hm);     // entries are increasing: symbol = #(<= r) - 1
    s = __shfl_sync(FULLMASK, cand, hbm1 + cnt);
    m = (127 * m + (le ? c10 : c10mix)) >> 7;                     // entry > r  <=>  entry > cdf[x]
    return cnt;
}

template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : D3_WPB * 32)
k_ans_dec3(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15, hb = lane & 16, half = lane >> 4;
    constexpr uint32_t STR = d3_stride<O1>(), CTX = d3_ctx_entries<O1>(), NENT = (O1 ? 256u : 1u) * CTX;
    uint16_t *T, *ring;
    if (O1) { T = (uint16_t *)smem_raw; ring = (uint16_t *)(smem_raw + NENT * 2 + half * D3_RING_BYTES); }
    else { uint8_t *b = smem_raw + (size_t)(wib * 2 + half) * (NENT * 2 + D3_RING_BYTES); T = (uint16_t *)b; ring = (uint16_t *)(b + NENT * 2); }
    uint32_t *ring32 = (uint32_t *)ring;
    const unsigned hm = 0xffffu << hb, hbm1 = hb - 1;
    const int c10 = ADAPT_IC_ * (int)i, c10mix = c10 + (int)AD_MIX;
    const size_t upw = O1 ? 1 : 2;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint8_t *gend = in + in_off[g.n_calls];
    const bool writer = (O1 ? lane : i) < 4;                      // lanes that store the four bytes of two pairs
    uint16_t *Ti = T + i;                                         // this lane's entry of table 0
    for (size_t jb = gw * upw; jb < g.n_calls; jb += nwarps * upw) {
        const size_t j = jb + (O1 ? 0 : half);
        bool act = j < g.n_calls;
        size_t start = 0, len = 0; uint64_t so = 0, sl = 0;
        if (act) { call_span(g, j, start, len); so = in_off[j]; sl = in_off[j + 1] - so; }
        uint8_t *op = out + start;
        if (act && sl == len) { group_copy(op, in + so, len, O1 ? lane : i, O1 ? 32 : 16); act = false; }   // raw chunk (CCPY turborc.c:434)
        const uint32_t nblk = act ? (uint32_t)((len + ANS_BLOCK - 1) / ANS_BLOCK) : 0u;
        const uint32_t nblk_o = __shfl_xor_sync(FULLMASK, nblk, 16), nblk_max = nblk > nblk_o ? nblk : nblk_o;
        // ---- stream ring: halfwords [hf-64, hf) are in shared memory, [hf, hf+32) in flight in `pend` (4 bytes per lane)
        const uint8_t *sp = act ? in + so : gend;
        uint32_t hp = 0, hf = 32;                                 // halfwords consumed / staged
        __syncwarp();
        { const uint32_t v = ld32_any(sp + 4 * i, gend); ring32[i] = v; if (i < 2) ring32[32 + i] = v; }
        uint32_t pend = ld32_any(sp + 64 + 4 * i, gend);
        __syncwarp();
        auto refill = [&]() {
            const bool need = hf - hp <= 32;                      // then every slot about to be overwritten has been consumed
            if (__any_sync(FULLMASK, need)) {
                __syncwarp();
                if (need) {
                    const uint32_t wb = (hf >> 1) & 31;           // 0 or 16
                    ring32[wb + i] = pend;
                    if (wb == 0 && i < 2) ring32[32 + i] = pend;
                    hf += 32;
                    pend = ld32_any(sp + 2 * (size_t)hf + 4 * i, gend);
                }
                __syncwarp();
            }
        };
        uint32_t cx = 0;                                          // not reset per block (anscdf.c:629)
        for (uint32_t b = 0; b < nblk_max; b++) {
            const bool bact = act && b < nblk;
            const size_t bpos = (size_t)b * ANS_BLOCK;
            const uint32_t n = bact ? (uint32_t)(len - bpos < ANS_BLOCK ? len - bpos : ANS_BLOCK) : 0u, npairs = (n + 1) >> 1;
            const uint32_t np_o = __shfl_xor_sync(FULLMASK, npairs, 16), npmax = npairs > np_o ? npairs : np_o;
            uint8_t *bo = op + bpos;
            __syncwarp();
            for (uint32_t k = O1 ? lane : i; k < NENT; k += O1 ? 32 : 16) T[k] = (uint16_t)((k % STR) << 11);   // CDF16DEC0/1/2 cdf_.h:26-32 (order 1: entry 16 = 32768)
            __syncwarp();
            refill();
            uint32_t s0 = ANS_L, s1 = ANS_L, s2 = ANS_L, s3 = ANS_L;
            if (bact) {                                           // mnfill anscdf_.h:176: st[0..3] ascending
                auto hw = [&](uint32_t q) -> uint32_t { return ring[(hp + q) & (D3_RING - 1)]; };
                s0 = hw(0) | hw(1) << 16; s1 = hw(2) | hw(3) << 16; s2 = hw(4) | hw(5) << 16; s3 = hw(6) | hw(7) << 16;
                hp += 8;
            }
            int mh = (int)(i << 11);                              // order 0: the high-nibble table never leaves its register
            auto nib_r = [&](uint32_t &st, int &m) -> uint32_t {  // next entry by shuffle (measured faster than any hoisting of it)
                const int dn = __shfl_down_sync(FULLMASK, m, 1, 16);
                return d3_nib(st, m, i == 15 ? (int)PROB_TOTAL : dn, hbm1, hm, c10, c10mix);
            };
            auto nib_h0 = [&](uint32_t &st) -> uint32_t { return nib_r(st, mh); };
            auto nib_t = [&](uint32_t &st, uint16_t *e) -> uint32_t {   // table in shared memory
                int m = e[0];
                uint32_t c;
                if (O1) {                                         // one warp alone on its SM: latency is everything, so the next entry
                    const int nx = e[1];                          // is read together with the entry itself.  It is the neighbour
                    __syncwarp();                                 // lane's e[0]: every read precedes every write ...
                    c = d3_nib(st, m, nx, hbm1, hm, c10, c10mix);
                    e[0] = (uint16_t)m;
                    __syncwarp();                                 // ... and every write the table's next use
                } else {
                    c = nib_r(st, m);
                    e[0] = (uint16_t)m;
                }
                return c;
            };
            for (uint32_t pi = 0; pi < npmax; pi += 2) {          // two byte pairs per trip (mndec8x2 / mndec8x2x anscdf_.h:152-174)
                refill();                                         // <= 8 halfwords are consumed per trip, >= 32 are staged after this
                uint32_t w = 0;                                   // the four decoded bytes
#pragma unroll
                for (int sub = 0; sub < 2; sub++) {
                    const bool pact = pi + sub < npairs;
                    uint32_t x0, x1;
                    if (O1) {
                        uint16_t *c0 = Ti + cx * CTX;
                        const uint32_t h0 = nib_t(s0, c0);
                        const uint32_t q0 = nib_t(s1, c0 + h0 * STR);
                        x0 = h0 * 16 + q0 - 17;
                        uint16_t *c1 = Ti + x0 * CTX;
                        const uint32_t h1 = nib_t(s2, c1);
                        const uint32_t q1 = nib_t(s3, c1 + h1 * STR);
                        x1 = h1 * 16 + q1 - 17;
                        cx = x1;
                    } else {
                        const uint32_t h0 = nib_h0(s0);
                        const uint32_t q0 = nib_t(s1, Ti + h0 * STR);
                        x0 = h0 * 16 + q0 - 17;
                        const uint32_t h1 = nib_h0(s2);
                        const uint32_t q1 = nib_t(s3, Ti + h1 * STR);
                        x1 = h1 * 16 + q1 - 17;
                    }
                    // ecdnorm x4 in state order (anscdf_.h:158-161): speculative 16-bit ring reads, predicated merges
                    const uint16_t *rp = ring + (hp & (D3_RING - 1));
                    uint32_t cnt = 0;
                    { const bool p = pact && s0 < ANS_L; const uint32_t v = rp[cnt]; s0 = p ? (s0 << 16 | v) : s0; cnt += p; }
                    { const bool p = pact && s1 < ANS_L; const uint32_t v = rp[cnt]; s1 = p ? (s1 << 16 | v) : s1; cnt += p; }
                    { const bool p = pact && s2 < ANS_L; const uint32_t v = rp[cnt]; s2 = p ? (s2 << 16 | v) : s2; cnt += p; }
                    { const bool p = pact && s3 < ANS_L; const uint32_t v = rp[cnt]; s3 = p ? (s3 << 16 | v) : s3; cnt += p; }
                    hp += cnt;
                    w |= (x0 | x1 << 8) << (16 * sub);
                }
                const uint32_t o = 2 * pi + (lane & 3);
                if (writer && o < n) bo[o] = (uint8_t)(w >> (8 * (lane & 3)));   // o < n also drops the odd tail's second byte (anscdf.c:602)
            }
        }
    }
}

// ================================================================================================================
// Order-1 decoder for batches of MANY calls (TRC_ANS1, anscdf1dec anscdf.c:629-645): one HALF-warp per call like the order-0
// decoder, so an SM runs 24 calls at once instead of one.  What made k_ans_dec3<true> a one-call-per-SM kernel is the 136 KB
// of low-nibble tables mbl[256][16][17]; here only the high-nibble tables mbh[256][16] (8 KB) live in shared memory and the
// low-nibble tables live in GLOBAL memory (128 KB per resident call, L1/L2 resident: the rows a call keeps returning to stay
// in L1).  A lane only ever reads and writes ITS OWN entry of every table row, so program order alone keeps the tables
// coherent -- no fences, no warp synchronisation around the table traffic.  The row fetch (an L1 or L2 hit) sits on the
// per-byte chain of a call (measured ~860 cycles per byte and call against ~200 with the tables in shared memory; prefetching
// the 16 candidate rows of the next context into L1 did not help: the call's own stores keep evicting them), so the kernel wins
// from about five calls per SM upwards.  Same bytes as k_ans_dec3<true>.
// ================================================================================================================
constexpr int G1_WPB = 4;                                         // warps per CTA: 8 calls
constexpr uint32_t G1_MBH_BYTES = 256u * 16u * 2u;                // high-nibble tables of one call
constexpr uint32_t G1_MBL_ENTRIES = 256u * 16u * 16u;             // low-nibble tables of one call (global memory)
constexpr uint32_t G1_SMEM = G1_WPB * 2u * (G1_MBH_BYTES + D3_RING_BYTES);

__global__ void __launch_bounds__(G1_WPB * 32)
k_ans1_dec_g(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g, uint16_t *__restrict__ mbl_all) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15, hb = lane & 16, half = lane >> 4;
    uint8_t *sb = smem_raw + (size_t)(wib * 2 + half) * (G1_MBH_BYTES + D3_RING_BYTES);
    uint16_t *Th = (uint16_t *)sb, *ring = (uint16_t *)(sb + G1_MBH_BYTES);
    uint32_t *ring32 = (uint32_t *)ring;
    const unsigned hm = 0xffffu << hb, hbm1 = hb - 1;
    const int c10 = ADAPT_IC_ * (int)i, c10mix = c10 + (int)AD_MIX;
    const size_t nwarps = (size_t)gridDim.x * G1_WPB, gw = (size_t)blockIdx.x * G1_WPB + wib;
    uint16_t *G = mbl_all + (gw * 2 + half) * (size_t)G1_MBL_ENTRIES + i;     // this lane's entry of row 0
    uint16_t *Thi = Th + i;
    const uint8_t *gend = in + in_off[g.n_calls];
    const bool writer = i < 4;
    for (size_t jb = gw * 2; jb < g.n_calls; jb += nwarps * 2) {
        const size_t j = jb + half;
        bool act = j < g.n_calls;
        size_t start = 0, len = 0; uint64_t so = 0, sl = 0;
        if (act) { call_span(g, j, start, len); so = in_off[j]; sl = in_off[j + 1] - so; }
        uint8_t *op = out + start;
        if (act && sl == len) { group_copy(op, in + so, len, i, 16); act = false; }     // raw chunk (CCPY turborc.c:434)
        const uint32_t nblk = act ? (uint32_t)((len + ANS_BLOCK - 1) / ANS_BLOCK) : 0u;
        const uint32_t nblk_o = __shfl_xor_sync(FULLMASK, nblk, 16), nblk_max = nblk > nblk_o ? nblk : nblk_o;
        const uint8_t *sp = act ? in + so : gend;
        uint32_t hp = 0, hf = 32;                                 // halfwords consumed / staged (see k_ans_dec3)
        __syncwarp();
        { const uint32_t v = ld32_any(sp + 4 * i, gend); ring32[i] = v; if (i < 2) ring32[32 + i] = v; }
        uint32_t pend = ld32_any(sp + 64 + 4 * i, gend);
        __syncwarp();
        auto refill = [&]() {
            const bool need = hf - hp <= 32;
            if (__any_sync(FULLMASK, need)) {
                __syncwarp();
                if (need) {
                    const uint32_t wb = (hf >> 1) & 31;
                    ring32[wb + i] = pend;
                    if (wb == 0 && i < 2) ring32[32 + i] = pend;
                    hf += 32;
                    pend = ld32_any(sp + 2 * (size_t)hf + 4 * i, gend);
                }
                __syncwarp();
            }
        };
        uint32_t cx = 0;                                          // not reset per block (anscdf.c:629)
        for (uint32_t b = 0; b < nblk_max; b++) {
            const bool bact = act && b < nblk;
            const size_t bpos = (size_t)b * ANS_BLOCK;
            const uint32_t n = bact ? (uint32_t)(len - bpos < ANS_BLOCK ? len - bpos : ANS_BLOCK) : 0u, npairs = (n + 1) >> 1;
            const uint32_t np_o = __shfl_xor_sync(FULLMASK, npairs, 16), npmax = npairs > np_o ? npairs : np_o;
            uint8_t *bo = op + bpos;
            if (bact) {                                           // CDF16DEC1 / CDF16DEC2 cdf_.h:28-32: every table = { j << 11 }
                const uint16_t v0 = (uint16_t)(i << 11);
                for (uint32_t k = 0; k < 256u * 16u; k += 16) Thi[k] = v0;
                // low-nibble tables: 128 KB of the same 32-byte row; 16 lanes x 16 bytes per step
                const uint32_t e0 = (i & 1) * 8;                  // first entry of this lane's 16-byte piece of a row
                const uint4 pat = make_uint4(((e0 + 1) << 27) | (e0 << 11), ((e0 + 3) << 27) | ((e0 + 2) << 11), ((e0 + 5) << 27) | ((e0 + 4) << 11), ((e0 + 7) << 27) | ((e0 + 6) << 11));
                uint4 *gp = (uint4 *)(G - i) + i;
                for (uint32_t k = 0; k < G1_MBL_ENTRIES / 8; k += 16) gp[k] = pat;
            }
            __syncwarp();
            refill();
            uint32_t s0 = ANS_L, s1 = ANS_L, s2 = ANS_L, s3 = ANS_L;
            if (bact) {                                           // mnfill anscdf_.h:176
                auto hw = [&](uint32_t q) -> uint32_t { return ring[(hp + q) & (D3_RING - 1)]; };
                s0 = hw(0) | hw(1) << 16; s1 = hw(2) | hw(3) << 16; s2 = hw(4) | hw(5) << 16; s3 = hw(6) | hw(7) << 16;
                hp += 8;
            }
            auto nib_r = [&](uint32_t &st, int &m) -> uint32_t {  // next entry by shuffle inside the half-warp
                const int dn = __shfl_down_sync(FULLMASK, m, 1, 16);
                return d3_nib(st, m, i == 15 ? (int)PROB_TOTAL : dn, hbm1, hm, c10, c10mix);
            };
            auto nib_s = [&](uint32_t &st, uint16_t *e) -> uint32_t { int m = e[0]; const uint32_t c = nib_r(st, m); e[0] = (uint16_t)m; return c; };
            auto nib_g = [&](uint32_t &st, uint16_t *e) -> uint32_t {
                int m = bact ? (int)*e : (int)(i << 11);          // (an idle half keeps the shuffles company)
                const uint32_t c = nib_r(st, m);
                if (bact) *e = (uint16_t)m;
                return c;
            };
            for (uint32_t pi = 0; pi < npmax; pi += 2) {          // two byte pairs per trip (mndec8x2x anscdf_.h:164-174)
                refill();
                uint32_t w = 0;
#pragma unroll
                for (int sub = 0; sub < 2; sub++) {
                    const bool pact = pi + sub < npairs;
                    const uint32_t h0 = nib_s(s0, Thi + cx * 16);
                    const uint32_t q0 = nib_g(s1, G + ((cx * 16 + h0 - 1) << 4));
                    const uint32_t x0 = (h0 * 16 + q0 - 17) & 0xffu;
                    const uint32_t h1 = nib_s(s2, Thi + x0 * 16);
                    const uint32_t q1 = nib_g(s3, G + ((x0 * 16 + h1 - 1) << 4));
                    const uint32_t x1 = (h1 * 16 + q1 - 17) & 0xffu;
                    cx = x1;
                    const uint16_t *rp = ring + (hp & (D3_RING - 1));
                    uint32_t cnt = 0;
                    { const bool p = pact && s0 < ANS_L; const uint32_t v = rp[cnt]; s0 = p ? (s0 << 16 | v) : s0; cnt += p; }
                    { const bool p = pact && s1 < ANS_L; const uint32_t v = rp[cnt]; s1 = p ? (s1 << 16 | v) : s1; cnt += p; }
                    { const bool p = pact && s2 < ANS_L; const uint32_t v = rp[cnt]; s2 = p ? (s2 << 16 | v) : s2; cnt += p; }
                    { const bool p = pact && s3 < ANS_L; const uint32_t v = rp[cnt]; s3 = p ? (s3 << 16 | v) : s3; cnt += p; }
                    hp += cnt;
                    w |= (x0 | x1 << 8) << (16 * sub);
                }
                const uint32_t o = 2 * pi + (lane & 3);
                if (writer && o < n) bo[o] = (uint8_t)(w >> (8 * (lane & 3)));
            }
        }
    }
}

// ================================================================================================================
// adaptive byte RANGE decoders TRC_RC (rccdfdec rccdf.c:187-200) and TRC_RCI (rccdfidec rccdf.c:213-228), same mapping as
// k_ans_dec3<false>: one HALF-warp per call, one CDF entry per lane, the high-nibble table in a register.
//   * _cdflget16 (turborc_.h:271-291: first entry with cdf[e+1] * range > code) = one 47 x 15-bit multiply-compare per lane +
//     ballot + popcount;
//   * every lane prepares the coder update for ITS entry -- code - cdf[e] * range and range * (cdf[e+1] - cdf[e])
//     (_rccdfupdate turborc_.h:219-229) -- before the symbol is known; four shuffles fetch the winner's;
//   * the stream is staged through a 32-word shared-memory ring per coder (any byte alignment, refilled one period ahead);
//     the renormalisation is one speculative 32-bit shared load + selects, no branch.
// NC == 2: coder 0 decodes the high nibbles, coder 1 the low nibbles (cdf8d2 rccdf_.h:63-73), two rings.
// ================================================================================================================
constexpr uint32_t R3_RING = 32;                                 // 32-bit words per ring
template <int NC> __host__ __device__ constexpr uint32_t r3_smem_bytes() { return D3_WPB * 2u * (O1_CTX_ENTRIES * 2u + NC * R3_RING * 4u); }

struct Rc3 {                                                     // range decoder state, replicated in the 16 lanes of a call
    uint32_t rl, rh, cl, ch;                                     // range, code
    uint32_t wp, wf, pend;                                       // words consumed / staged; this lane's word of [wf, wf+16)
    uint32_t *ring; const uint8_t *sp;
};

// one nibble; m = this lane's entry (updated on return).  Returns cnt = symbol + 1.
__device__ __forceinline__ uint32_t r3_nib(Rc3 &d, int &m, unsigned i, unsigned hbm1, unsigned hm, int c10, int c10mix) {
    const int dn = __shfl_down_sync(FULLMASK, m, 1, 16);
    const uint32_t f = (uint32_t)((i == 15 ? (int)PROB_TOTAL : dn) - m), mu = (uint32_t)m;
    const uint32_t rl = __funnelshift_r(d.rl, d.rh, PROB_BITS), rh = d.rh >> PROB_BITS;     // range >>= 15
    const uint32_t pl = mu * rl, ph = __umulhi(mu, rl) + mu * rh;                           // cdf[e] * range
    const bool le = ph < d.ch || (ph == d.ch && pl <= d.cl);                                  // entry 0 == 0: always true
    uint32_t dl, dh;
    asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(d.cl), "r"(d.ch), "r"(pl), "r"(ph));
    const uint32_t nl = rl * f, nh = __umulhi(rl, f) + rh * f;                              // range * freq
    const unsigned cnt = __popc(__ballot_sync(FULLMASK, le) & hm), src = hbm1 + cnt;
    const uint32_t xdl = __shfl_sync(FULLMASK, dl, src), xdh = __shfl_sync(FULLMASK, dh, src);
    const uint32_t xnl = __shfl_sync(FULLMASK, nl, src), xnh = __shfl_sync(FULLMASK, nh, src);
    const uint32_t w = d.ring[d.wp & (R3_RING - 1)];                                          // speculative: used iff the range dropped below 2^32
    const bool p = xnh == 0;                                                                  // _rcdnorm_ turborc_.h:111
    d.rh = p ? xnl : xnh; d.rl = p ? 0u : xnl;
    d.ch = p ? xdl : xdh; d.cl = p ? w : xdl;
    d.wp += p;
    m = (127 * m + (le ? c10 : c10mix)) >> 7;                                                 // cdf16upd
    return cnt;
}

template <int NC>
__global__ void __launch_bounds__(D3_WPB * 32)
k_rc_dec3(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15, hb = lane & 16, half = lane >> 4;
    uint8_t *blk = smem_raw + (size_t)(wib * 2 + half) * (O1_CTX_ENTRIES * 2u + NC * R3_RING * 4u);
    uint16_t *T = (uint16_t *)blk, *Ti = T + i;
    uint32_t *rings = (uint32_t *)(blk + O1_CTX_ENTRIES * 2u);
    const unsigned hm = 0xffffu << hb, hbm1 = hb - 1;
    const int c10 = ADAPT_IC_ * (int)i, c10mix = c10 + (int)AD_MIX;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint8_t *gend = in + in_off[g.n_calls];
    const bool writer = i < 4;
    for (size_t jb = gw * 2; jb < g.n_calls; jb += nwarps * 2) {
        const size_t j = jb + half;
        bool act = j < g.n_calls;
        size_t start = 0, n = 0; uint64_t so = 0, sl = 0;
        if (act) { call_span(g, j, start, n); so = in_off[j]; sl = in_off[j + 1] - so; }
        uint8_t *op = out + start;
        if (act && sl == n) { group_copy(op, in + so, n, i, 16); act = false; }              // raw chunk (CCPY turborc.c:434)
        if (!act) n = 0;
        const uint32_t n32 = (uint32_t)n, n_o = __shfl_xor_sync(FULLMASK, n32, 16), nmax = n32 > n_o ? n32 : n_o;
        // ---- coders and their stream rings
        Rc3 d[NC];
        const uint8_t *stream = act ? in + so : gend;
        d[0].sp = stream + (NC == 2 ? 4 : 0);
        if (NC == 2) {
            const uint32_t len0 = ld_u32_clamped(stream, gend);                                // rccdf.c:215
            const uint8_t *p1 = stream + 4 + len0;
            d[NC - 1].sp = (p1 > gend || p1 < stream) ? gend : p1;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NC; c++) {
            d[c].ring = rings + c * R3_RING;
            d[c].ring[i] = ld32_any(d[c].sp + 4 * i, gend);
            d[c].pend = ld32_any(d[c].sp + 64 + 4 * i, gend);
            d[c].wf = 16; d[c].wp = 2;
        }
        for (uint32_t k = i; k < (uint32_t)O1_CTX_ENTRIES; k += 16) T[k] = (uint16_t)((k & 15) << 11);   // CDF16DEC0/1 (rccdf.c:188)
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NC; c++) {                                                         // rcdinit turborc_.h:152-158
            d[c].rl = d[c].rh = 0xffffffffu; d[c].ch = d[c].ring[0]; d[c].cl = d[c].ring[1];
        }
        auto refill = [&]() {
            bool need = false;
#pragma unroll
            for (int c = 0; c < NC; c++) need = need || d[c].wf - d[c].wp <= 16;               // then the slots to overwrite have been consumed
            if (__any_sync(FULLMASK, need)) {
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (d[c].wf - d[c].wp <= 16) {
                        d[c].ring[(d[c].wf & (R3_RING - 1)) + i] = d[c].pend;
                        d[c].wf += 16;
                        d[c].pend = ld32_any(d[c].sp + 4 * (size_t)d[c].wf + 4 * i, gend);
                    }
                __syncwarp();
            }
        };
        int mh = (int)(i << 11);                                                               // high-nibble table: stays in its register
        for (uint32_t k = 0; k < nmax; k += 4) {                                               // four bytes per trip: <= 8 words per coder
            refill();
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {                                                      // cdf8d / cdf8d2 rccdf_.h:50-73
                const uint32_t h = r3_nib(d[0], mh, i, hbm1, hm, c10, c10mix);
                uint16_t *e = Ti + h * 16;
                int m = *e;
                const uint32_t l = r3_nib(d[NC - 1], m, i, hbm1, hm, c10, c10mix);
                *e = (uint16_t)m;
                w |= (h * 16 + l - 17) << (8 * b);
            }
            const uint32_t o = k + (lane & 3);
            if (writer && o < n32) op[o] = (uint8_t)(w >> (8 * (lane & 3)));
        }
    }
}

}  // namespace trc
