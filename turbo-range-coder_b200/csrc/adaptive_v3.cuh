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
//       * tables are write-through in shared memory; the table of byte t+1 is loaded while byte t is being updated
//         and a select forwards the fresh entry when both bytes use the same table, so shared-memory latency is off
//         the chain;
//       * the (freq | cum << 16) record of a nibble (mnenc4 anscdf_.h:106) is produced by the lane that owns the coded
//         symbol and leaves through a 32-record staging line as one coalesced 128-byte store per 16 bytes.
//   k_ans_code3   (encoder, coding pass mnflush anscdf_.h:128-138; one LANE per (unit, rANS state): a warp runs 8
//                  units x 4 states, records popped last to first)
//       * division by the adaptive frequency is an exact multiply-high with a reciprocal taken from a 32 K-entry table
//         (128 KB, staged in shared memory by TMA bulk copies) -- no integer division on the chain;
//       * the words the four states of a unit emit in one step are ordered with ballot + popcount (LIFO order of the
//         reference); records and reciprocals are software-pipelined 16-32 steps ahead in registers.
//   k_ans_dec3    (decoder; order 0: one HALF-warp per call, two calls per warp; order 1: one warp per call, both halves
//                  replicate, 136 KB of tables in the shared memory of one SM)
//       * symbol search of cdf16ansdec (cdf_.h:52-59) = compare + ballot + popcount; every lane computes the state
//         update for ITS entry before the symbol is known and one shuffle picks the right one;
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
constexpr uint32_t M3_STAGE_WORDS = 64;                          // two staging lines of 32 records
template <bool O1> __host__ __device__ constexpr uint32_t m3_warp_bytes() { return (O1 ? 256u : 1u) * O1_CTX_ENTRIES * 2u + M3_STAGE_WORDS * 4u; }

struct M3State { uint32_t x, off; int m; };                      // current byte, entry offset of its table, this lane's entry

// one byte: record of the coded nibble, cdf16upd of this lane's entry, hand-over to the table of the next byte x_n
template <bool O1>
__device__ __forceinline__ void m3_step(M3State &s, uint32_t x_n, uint16_t *T, uint32_t *stage_slot, unsigned i, unsigned h,
                                        uint32_t xsh, int c10) {
    const uint32_t off_n = (O1 ? s.x * (uint32_t)O1_CTX_ENTRIES : 0u) + (h ? 16u + (x_n & 0xf0u) : 0u);   // mbh[cx] / mbl[cx][x_n >> 4]
    const int pre = T[off_n + i];                                 // stale only if off_n == s.off (forwarded below)
    const uint32_t xs = (s.x >> xsh) & 15u;
    const int dn = __shfl_down_sync(FULLMASK, s.m, 1, 16);
    const uint32_t f = (uint32_t)((i == 15 ? (int)PROB_TOTAL : dn) - s.m);
    if (i == xs) *stage_slot = f | (uint32_t)s.m << 16;
    const int m2 = (127 * s.m + c10 + (i > xs ? (int)AD_MIX : 0)) >> 7;
    T[s.off + i] = (uint16_t)m2;
    s.m = off_n == s.off ? m2 : pre;
    s.off = off_n; s.x = x_n;
}

template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : M3_WPB * 32)
k_ans_model3(const uint8_t *__restrict__ in, Geom g, uint32_t *__restrict__ recs, size_t rec_stride) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
    constexpr uint32_t NENT = (O1 ? 256u : 1u) * O1_CTX_ENTRIES;
    uint16_t *T = (uint16_t *)(smem_raw + (size_t)wib * m3_warp_bytes<O1>());
    uint32_t *stage = (uint32_t *)(T + NENT);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int c10 = ADAPT_IC_ * (int)i;
    const uint32_t xsh = h ? 0u : 4u;
    for (size_t u = gw; u < g.n_units; u += nwarps) {
        size_t j, start, len; uint32_t b;
        unit_span(g, u, j, b, start, len);
        if (len == 0) continue;                                   // padding unit: k_ans_code3 writes its (empty) meta
        const uint8_t *ip = in + start;
        const uint32_t n = (uint32_t)len, nb = (n + 1) & ~1u;     // odd tail: a dummy 0 byte is coded too (anscdf.c:581,621)
        uint32_t *rec = recs + u * rec_stride;
        __syncwarp();
        for (uint32_t k = lane; k < NENT; k += 32) T[k] = (uint16_t)((k & 15) << 11);   // CDF16DEC0/1/2 cdf_.h:26-32
        __syncwarp();
        const uint32_t cx0 = (O1 && start > j * g.chunk) ? in[start - 1] : 0;           // cx carries across blocks (anscdf.c:608)
        uint32_t mine = lane < n ? ip[lane] : 0;
        M3State s;
        s.x = __shfl_sync(FULLMASK, mine, 0);
        s.off = (O1 ? cx0 * (uint32_t)O1_CTX_ENTRIES : 0u) + (h ? 16u + (s.x & 0xf0u) : 0u);
        s.m = T[s.off + i];
        for (uint32_t base = 0; base < nb; base += 32) {
            const uint32_t nidx = base + 32 + lane;
            const uint32_t mine_n = nidx < n ? ip[nidx] : 0;
            const uint32_t cnt = nb - base < 32 ? nb - base : 32;
            if (cnt == 32) {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    const uint32_t x_n = k + 1 < 32 ? __shfl_sync(FULLMASK, mine, k + 1) : __shfl_sync(FULLMASK, mine_n, 0);
                    m3_step<O1>(s, x_n, T, stage + ((k >> 4) & 1) * 32 + 2 * (k & 15) + h, i, h, xsh, c10);
                    if ((k & 15) == 15) {                          // 16 bytes = 32 records: one 128-byte store
                        __syncwarp();
                        rec[2 * (base + (k & ~15)) + lane] = stage[((k >> 4) & 1) * 32 + lane];
                    }
                }
            } else {
                for (uint32_t k = 0; k < cnt; k++) {
                    const uint32_t x_n = __shfl_sync(FULLMASK, mine, (k + 1) & 31);       // past the end: any value (prefetch only)
                    m3_step<O1>(s, x_n, T, stage + ((k >> 4) & 1) * 32 + 2 * (k & 15) + h, i, h, xsh, c10);
                    if ((k & 15) == 15 || k + 1 == cnt) {
                        __syncwarp();
                        if (lane < 2 * ((k & 15) + 1)) rec[2 * (base + (k & ~15u)) + lane] = stage[((k >> 4) & 1) * 32 + lane];
                    }
                }
            }
            mine = mine_n;
        }
    }
}

// ================================================================================================================
// encoder, coding pass
// ================================================================================================================
constexpr int C3_WPB = 2;                                        // warps per CTA (they share the reciprocal table)
constexpr int C3_B = 16;                                         // steps per software-pipeline block
constexpr uint32_t C3_LUT_BYTES = PROB_TOTAL * 4;                // 128 KB

__device__ uint32_t g_rcp_lut[PROB_TOTAL];                       // [f - 1] = reciprocal of rans_enc_entry(., f), f = 1 .. 2^15

__global__ void k_build_rcp() {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < PROB_TOTAL) g_rcp_lut[k] = rans_enc_entry(0, k + 1).x;
}

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
    const size_t gw = (size_t)blockIdx.x * C3_WPB + (threadIdx.x >> 5);
    const size_t u = gw * 8 + (lane >> 2);
    size_t j, start, len = 0; uint32_t blk;
    if (u < g.n_units) unit_span(g, u, j, blk, start, len);
    const bool live = len != 0;
    const uint32_t npairs = (uint32_t)((len + 1) >> 1);
    const uint32_t tmax = __reduce_max_sync(FULLMASK, npairs);
    const uint32_t *rec = recs + (live ? u : 0) * rec_stride;
    uint8_t *slot = slots + (live ? u : 0) * slot_stride;
    const int cap = (int)slot_stride;
    int pos = cap;                                               // lowest byte written so far (same in the 4 lanes of a unit)
    uint32_t s = ANS_L;
    bool em = false, ovf = false;
    // step t codes record 4 (npairs-1-t) + 3 - k on state k (pushed 3,2,1,0 per byte pair -> popped 0,1,2,3)
    auto ldrec = [&](uint32_t t) -> uint32_t { return t < npairs ? __ldg(rec + 4 * (size_t)(npairs - 1 - t) + 3 - k) : 1u; };
    uint32_t R[C3_B], Rn[C3_B], Q[C3_B];
#pragma unroll
    for (int q = 0; q < C3_B; q++) { R[q] = ldrec(q); Rn[q] = ldrec(C3_B + q); }
#pragma unroll
    for (int q = 0; q < C3_B; q++) Q[q] = rcp_s[((R[q] & 0xffffu) - 1) & PROB_MASK];
    for (uint32_t t0 = 0; t0 < tmax; t0 += C3_B) {
        uint32_t Rn2[C3_B], Qn[C3_B];
#pragma unroll
        for (int q = 0; q < C3_B; q++) Rn2[q] = ldrec(t0 + 2 * C3_B + q);
#pragma unroll
        for (int q = 0; q < C3_B; q++) Qn[q] = rcp_s[((Rn[q] & 0xffffu) - 1) & PROB_MASK];
#pragma unroll
        for (int q = 0; q < C3_B; q++) {
            const bool act = t0 + q < npairs && !ovf;
            const uint32_t f = R[q] & 0xffffu, c = R[q] >> 16;
            const uint32_t sh = 31 - __clz((int)((f - 1) | 1));                       // ceil(log2 f) - 1 (0 for f <= 2)
            const uint32_t bias = c + (f == 1 ? PROB_TOTAL - 1 : 0);                  // rans_enc_entry's f == 1 form
            const bool p = act && s >= (f << 16);                                     // ecenorm anscdf_.h:48
            const unsigned grp = (__ballot_sync(FULLMASK, p) >> gb) & 15u;
            if (p) st_u16(slot + pos - 2 * (__popc(grp & ((1u << k) - 1)) + 1), s);   // state 0's word highest
            const uint32_t s1 = p ? s >> 16 : s;
            const uint32_t qq = __umulhi(s1, Q[q]) >> sh;                             // == s1 / f
            const uint32_t sn = s1 + bias + qq * (PROB_TOTAL - f);                    // (q << 15) + s1 % f + cum
            s = act ? sn : s;
            pos -= 2 * __popc(grp);
            em = act ? p : em;
            ovf = ovf || pos < 32;                                                    // slot exhausted
        }
#pragma unroll
        for (int q = 0; q < C3_B; q++) { R[q] = Rn[q]; Q[q] = Qn[q]; Rn[q] = Rn2[q]; }
    }
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
template <bool O1> __host__ __device__ constexpr uint32_t d3_smem_bytes() {
    return O1 ? 256u * O1_CTX_ENTRIES * 2u + 2u * D3_RING_BYTES : D3_WPB * 2u * (O1_CTX_ENTRIES * 2u + D3_RING_BYTES);
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

// one nibble: cdf16ansdec (cdf_.h:52-59) + STATEUPD (cdf_.h:37) + cdf16upd; every lane of the unit returns the same x, s
__device__ __forceinline__ uint32_t d3_nib(uint32_t &s, int &m, unsigned i, unsigned hb, unsigned hm, int c10) {
    const uint32_t r = s & PROB_MASK;
    const bool le = (uint32_t)m <= r;                             // entry 0 == 0: always true
    const int dn = __shfl_down_sync(FULLMASK, m, 1, 16);
    const uint32_t f = (uint32_t)((i == 15 ? (int)PROB_TOTAL : dn) - m);
    const uint32_t cand = f * (s >> PROB_BITS) + r - (uint32_t)m; // the new state if this lane's entry is the symbol
    const unsigned x = __popc(__ballot_sync(FULLMASK, le) & hm) - 1;   // entries are increasing: #(<= r) - 1
    s = __shfl_sync(FULLMASK, cand, hb | x);
    m = (127 * m + c10 + (le ? 0 : (int)AD_MIX)) >> 7;            // entry > r  <=>  entry > cdf[x]
    return x;
}

template <bool O1>
__global__ void __launch_bounds__(O1 ? 32 : D3_WPB * 32)
k_ans_dec3(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, i = lane & 15, hb = lane & 16, half = lane >> 4;
    constexpr uint32_t NENT = (O1 ? 256u : 1u) * O1_CTX_ENTRIES;
    uint16_t *T, *ring;
    if (O1) { T = (uint16_t *)smem_raw; ring = (uint16_t *)(smem_raw + NENT * 2 + half * D3_RING_BYTES); }
    else { uint8_t *b = smem_raw + (size_t)(wib * 2 + half) * (NENT * 2 + D3_RING_BYTES); T = (uint16_t *)b; ring = (uint16_t *)(b + NENT * 2); }
    uint32_t *ring32 = (uint32_t *)ring;
    const unsigned hm = 0xffffu << hb;
    const int c10 = ADAPT_IC_ * (int)i;
    const size_t upw = O1 ? 1 : 2;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint8_t *gend = in + in_off[g.n_calls];
    const bool writer = (O1 ? lane : i) < 2;                      // lanes that store the two bytes of a pair
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
            for (uint32_t k = O1 ? lane : i; k < NENT; k += O1 ? 32 : 16) T[k] = (uint16_t)((k & 15) << 11);
            __syncwarp();
            refill();
            uint32_t s0 = ANS_L, s1 = ANS_L, s2 = ANS_L, s3 = ANS_L;
            if (bact) {                                           // mnfill anscdf_.h:176: st[0..3] ascending
                auto hw = [&](uint32_t q) -> uint32_t { return ring[(hp + q) & (D3_RING - 1)]; };
                s0 = hw(0) | hw(1) << 16; s1 = hw(2) | hw(3) << 16; s2 = hw(4) | hw(5) << 16; s3 = hw(6) | hw(7) << 16;
                hp += 8;
            }
            int mh = (int)(i << 11);                              // order 0: the high-nibble table never leaves its register
            for (uint32_t pi = 0; pi < npmax; pi++) {             // mndec8x2 / mndec8x2x anscdf_.h:152-174
                const bool pact = pi < npairs;
                refill();
                uint32_t x0, x1;
                if (O1) {
                    const uint32_t c0 = cx * (uint32_t)O1_CTX_ENTRIES + i;
                    int m = T[c0];
                    const uint32_t yh0 = d3_nib(s0, m, i, hb, hm, c10); T[c0] = (uint16_t)m;
                    const uint32_t l0 = c0 + (1 + yh0) * 16;
                    m = T[l0];
                    const uint32_t yl0 = d3_nib(s1, m, i, hb, hm, c10); T[l0] = (uint16_t)m;
                    x0 = yh0 << 4 | yl0;
                    const uint32_t c1 = x0 * (uint32_t)O1_CTX_ENTRIES + i;
                    m = T[c1];
                    const uint32_t yh1 = d3_nib(s2, m, i, hb, hm, c10); T[c1] = (uint16_t)m;
                    const uint32_t l1 = c1 + (1 + yh1) * 16;
                    m = T[l1];
                    const uint32_t yl1 = d3_nib(s3, m, i, hb, hm, c10); T[l1] = (uint16_t)m;
                    x1 = yh1 << 4 | yl1;
                    cx = x1;
                } else {
                    const uint32_t yh0 = d3_nib(s0, mh, i, hb, hm, c10);
                    const uint32_t l0 = (1 + yh0) * 16 + i;
                    int m = T[l0];
                    const uint32_t yl0 = d3_nib(s1, m, i, hb, hm, c10); T[l0] = (uint16_t)m;
                    x0 = yh0 << 4 | yl0;
                    const uint32_t yh1 = d3_nib(s2, mh, i, hb, hm, c10);
                    const uint32_t l1 = (1 + yh1) * 16 + i;
                    m = T[l1];
                    const uint32_t yl1 = d3_nib(s3, m, i, hb, hm, c10); T[l1] = (uint16_t)m;
                    x1 = yh1 << 4 | yl1;
                }
                // ecdnorm x4 in state order (anscdf_.h:158-161): speculative 16-bit ring reads, predicated merges
                const uint16_t *rp = ring + (hp & (D3_RING - 1));
                uint32_t cnt = 0;
                { const bool p = pact && s0 < ANS_L; const uint32_t v = rp[cnt]; s0 = p ? (s0 << 16 | v) : s0; cnt += p; }
                { const bool p = pact && s1 < ANS_L; const uint32_t v = rp[cnt]; s1 = p ? (s1 << 16 | v) : s1; cnt += p; }
                { const bool p = pact && s2 < ANS_L; const uint32_t v = rp[cnt]; s2 = p ? (s2 << 16 | v) : s2; cnt += p; }
                { const bool p = pact && s3 < ANS_L; const uint32_t v = rp[cnt]; s3 = p ? (s3 << 16 | v) : s3; cnt += p; }
                hp += cnt;
                const uint32_t o = 2 * pi + (lane & 1);
                if (pact && writer && o < n) bo[o] = (uint8_t)((lane & 1) ? x1 : x0);   // odd tail: second byte discarded (anscdf.c:602)
            }
        }
    }
}

}  // namespace trc
