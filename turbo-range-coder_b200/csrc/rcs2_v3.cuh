// rcs2_v3.cuh -- TRC_RCS2 (rccdfs2enc / rccdfsb2dec, rccdf.c:125-184) headline kernels, third generation.
//
// Same mapping as static_v2.cuh's lane-per-coder kernels (lanes 2r / 2r+1 of a warp own coder 0 / coder 1 of call r),
// rebuilt around what the round-1 profiles and the micro-benchmarks in tools/ubench showed:
//
//  * ENCODE was bound by its scattered 4-byte global stores, not by arithmetic (tools/ubench/enc_step.cu: the same step
//    costs 71 cycles per warp-symbol with stores to the global slot and 38 with stores to shared memory).  Emitted words now
//    go to an 8-word ring per lane in shared memory (word-major, lane-minor: conflict-free whatever position a lane is at)
//    and leave as 128-bit stores, four words at a time.
//  * The coder step keeps `low` as a 96-bit number whose top word is the pending output word, so the carry of
//    low += range * cdf[x] (turborc_.h:215, _rccarry_ :103) runs straight into the pending word through the add-with-carry
//    chain: no carry flag, no compare, no select -- 21.5 SASS instructions per symbol instead of 41.  A carry OUT of the
//    pending word (it held 0xffffffff: the reference would walk back through stored words, p ~ 2^-32 per word) leaves a
//    ZERO word behind; every stored word passes through a running minimum, and a call that ever stored a zero word
//    (wrapped or genuine, both p ~ 2^-32 per word) is redone by the exact walk-back coder.
//  * INPUT arrives by TMA: the batch is described as a 2-D tensor [calls][chunk bytes] and every warp pulls the next
//    128 bytes of its 16 calls with ONE cp.async.bulk.tensor (16 x 128 B box, 128-byte swizzle so the lanes' 16-byte reads
//    are bank-conflict free), double buffered on two mbarriers per warp.  No lane computes a global address or holds
//    prefetch registers, DRAM sees whole 128-byte lines, and warps never synchronise with each other inside the loop.
//  * The layout epilogue (decoupled look-back over CTA tiles, copy of the pieces to their final place) is the one of
//    k_rcs2_enc_fused.
//  * One table for the batch, or one per aligned group of calls (chunks_per_cdf, a multiple of the CTA's calls): a CTA loads the
//    table of its group.  Launch shapes (one wave of one CTA per SM / waves of two CTAs per SM): e3_shape, lpc_shape in trc_b200.cu.
#pragma once
#include <cuda.h>
#include "static_v2.cuh"

namespace trc {

constexpr int      E3_MAX_NT  = 1024;                   // one CTA per SM: up to 512 calls, 32 warps (64 registers per thread = the whole register file)
constexpr int      E3_RING_W  = 8;                      // ring words per lane (<= 3 left over + <= 4 new per 8-symbol block)
constexpr uint32_t E3_RING_S  = 4096;                   // bytes between consecutive ring words of a lane (room for 1024 lanes; a power of two so that the
                                                        // ring address is ONE shift-and-add of the cursor)
constexpr uint32_t E3_LEAD    = 6;                      // blocks a warp may run ahead of the slowest warp of its CTA (see the pacing note in k_rcs2_enc3)
constexpr uint32_t E3_TILE_BYTES = 16 * 128;            // one input stage of a warp: 16 calls x 128 bytes
constexpr int      E3_STAGES = 2;
constexpr int      E3_COPY_U = 16;                     // words in flight per lane in the layout epilogue

// {cdf, freq} of every symbol as one 8-byte entry (one LDS.64 per symbol, nothing to unpack)
struct __align__(16) EncTab2 { uint2 e[256]; };

__global__ void k_build_enctab2(const cdf_t *__restrict__ cdf, unsigned cdfnum, EncTab2 *__restrict__ t) {   // one CTA per table
    const cdf_t *c0 = cdf + (size_t)blockIdx.x * CDF_STRIDE;
    for (unsigned x = threadIdx.x; x < 256; x += blockDim.x) {
        uint32_t c = 0, f = 0;
        if (x < cdfnum) { c = c0[x]; f = (uint32_t)c0[x + 1] - c; }
        t[blockIdx.x].e[x] = make_uint2(c, f);
    }
}

// ---- the coder: range (64 bit) | low (64 bit) | pending word | carries out of the pending word ------------------------
struct RcE96 {
    uint32_t rl, rh, ll, lh, pend;
    uint32_t nz;                                        // min over every word stored so far: 0 = some stored word was zero (see below)
    uint32_t k;                                         // ring cursor: (words put so far mod 8) << 29
    // pend starts at 1: the first renormalisation stores it as a scratch word in front of the stream (never part of the output),
    // no carry can reach it before that (low + range < 2^64 until the first renormalisation), and it must not look like a zero word
    __device__ __forceinline__ void init() { rl = rh = 0xffffffffu; ll = lh = 0; pend = 1; nz = 1; k = 0; }
    // one symbol (_rccdfenc_ + _rcenorm_, turborc_.h:215,105-109); ringlane = shared address of ring word 0 of this lane
    __device__ __forceinline__ void encode(uint32_t c0, uint32_t f, uint32_t ringlane) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, tl, th, a;\n\t"
            ".reg .u64 nr, tt;\n\t"
            "shf.r.wrap.b32 %0, %0, %1, 15;\n\t"         // range >>= 15
            "shr.u32 %1, %1, 15;\n\t"
            "mul.wide.u32 tt, %0, %7;\n\t"               // range * cdf[x]
            "mov.b64 {tl, th}, tt;\n\t"
            "mad.lo.u32 th, %1, %7, th;\n\t"
            "add.cc.u32 %2, %2, tl;\n\t"                 // low += ...; the carry runs into the pending word
            "addc.cc.u32 %3, %3, th;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mul.wide.u32 nr, %0, %8;\n\t"               // range *= freq
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %8, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"                  // range < 2^32: renormalise
            "mad.hi.u32 a, %6, 32768, %9;\n\t"           // ring address = lane base + (k >> 29) * 4096
            "@p st.shared.u32 [a], %4;\n\t"
            "@p add.u32 %6, %6, 0x20000000;\n\t"
            "@p mov.u32 %4, %3;\n\t"
            "@p mov.u32 %3, %2;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(pend), "+r"(nz), "+r"(k)
            : "r"(c0), "r"(f), "r"(ringlane) : "memory");
    }
};

// flush (rceflush turborc_.h:118-128) on the same state, words into the ring at sequence position `wr`
__device__ __forceinline__ void e96_put(RcE96 &e, uint32_t w, uint32_t ringlane, uint32_t &wr) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(ringlane + (wr & (E3_RING_W - 1)) * E3_RING_S), "r"(e.pend) : "memory");
    e.nz = min(e.nz, e.pend);
    e.pend = w; wr++;
}
__device__ __forceinline__ void e96_add(RcE96 &e, uint32_t al, uint32_t ah) {
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
        : "+r"(e.ll), "+r"(e.lh), "+r"(e.pend) : "r"(al), "r"(ah));
}

// ---- TMA: 2-D tile of the input (tensor = [calls][chunk bytes], box = 16 calls x 128 bytes) ---------------------------
// The input is read exactly once: its lines are marked evict-first in L2 so that the slot images written by this kernel
// (read back by its layout epilogue) stay resident instead of being pushed out to DRAM by the input stream.
__device__ __forceinline__ void tma_tile_2d(uint32_t smem_dst, const CUtensorMap *tmap, int x, int y, uint32_t bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(E3_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// dynamic shared memory of k_rcs2_enc3: [pad to 2048][table 2 KB][input tiles: warps x E3_STAGES x 2 KB][ring: 8 words x 4 KB]
__host__ __device__ inline size_t e3_smem_bytes(unsigned nthreads, bool tma) {
    return 2048 + 2048 + (tma ? (size_t)(nthreads / 32) * E3_STAGES * E3_TILE_BYTES : 0) + (size_t)E3_RING_W * E3_RING_S;
}

// phase timing probe (tools/enc_phases.py): when set, every warp records globaltimer at its phase boundaries
__device__ unsigned long long *g_e3_times = nullptr;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define E3_STAMP(k) do { if (tp && lane == 0) tp[k] = gtime(); } while (0)

// g.chunk is a multiple of 16.  The tensor map covers the FULL calls only (a shorter last call would make TMA read past the
// end of the caller's buffer): the two lanes of a short last call load their bytes directly and run the generic remainder.
template <bool TMA>
__global__ void __maxnreg__(64)
k_rcs2_enc3(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ in, Geom g, size_t n_calls,
            const EncTab2 *__restrict__ tab, uint8_t *__restrict__ slots, size_t slot_stride, unsigned calls_per_cta,
            volatile unsigned long long *__restrict__ lb, uint64_t *__restrict__ out_off, uint8_t *__restrict__ out, unsigned flags, size_t cpc) {
    __shared__ uint64_t bar;
    __shared__ uint64_t fullbar[(E3_MAX_NT / 32) * E3_STAGES];
    __shared__ uint32_t s_len[E3_MAX_NT / 2], s_alen[E3_MAX_NT / 2], s_boff[E3_MAX_NT / 2], s_blen[E3_MAX_NT / 2], s_excl[E3_MAX_NT / 2];
    __shared__ uint32_t s_prog[32];                                                        // blocks done per warp (pacing)
    constexpr int NW = E3_MAX_NT / 32;
    __shared__ uint32_t s_wsum[NW + 1];
    __shared__ unsigned long long s_base;
    __shared__ unsigned s_tile;
    extern __shared__ __align__(16) uint8_t dyn[];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long *tp = g_e3_times ? g_e3_times + ((size_t)blockIdx.x * 32 + wid) * 8 : nullptr;
    E3_STAMP(0);
    const uint32_t dyn0 = smem_u32(dyn);
    const uint32_t tb = (dyn0 + 2047u) & ~2047u;                                           // symbol table, 2 KB aligned: its address is OR-ed into the entry offsets
    uint2 *ctab = (uint2 *)(dyn + (tb - dyn0));
    const uint32_t tiles0 = tb + 2048;                                                     // 128-byte swizzle wants 1 KB aligned tiles
    const uint32_t ring0 = TMA ? tiles0 + (blockDim.x >> 5) * E3_STAGES * E3_TILE_BYTES : tiles0;
    if (threadIdx.x < 32) s_prog[threadIdx.x] = threadIdx.x < (blockDim.x >> 5) ? 0u : 0xffffffffu;
    if (threadIdx.x == 0) {
        s_tile = (unsigned)atomicAdd((unsigned long long *)(lb + gridDim.x), 1ull);     // tile index in arrival order (look-back safe)
        tma_fetch(ctab, tab[cpc ? (size_t)s_tile * calls_per_cta / cpc : 0].e, 2048, &bar);   // cpc: calls per table (a CTA never straddles two groups)
        if (TMA) {
            for (unsigned k = 0; k < (blockDim.x >> 5) * E3_STAGES; k++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fullbar[k])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const unsigned bid = s_tile;
    const size_t j0 = (size_t)bid * calls_per_cta, j = j0 + (threadIdx.x >> 1);
    const unsigned c = threadIdx.x & 1, r = threadIdx.x >> 1;
    const bool live = j < n_calls && r < calls_per_cta;
    const bool is_tail = live && j * g.chunk + g.chunk > g.total;                          // the last call of the batch when it is shorter than a chunk
    const uint32_t n = is_tail ? (uint32_t)(g.total - j * g.chunk) : (uint32_t)g.chunk;
    const uint32_t nb = n >> 4, nbmax = (uint32_t)(g.chunk >> 4);                          // 16-byte blocks of this call / of a full call (8 symbols per lane each)
    const uint32_t fb0 = smem_u32(&fullbar[wid * E3_STAGES]);
    const uint32_t mytiles = tiles0 + wid * E3_STAGES * E3_TILE_BYTES;
    const int row0 = (int)(j0 + wid * 16);                                                 // first call (tensor row) of this warp
    const uint32_t nst = ((uint32_t)g.chunk + 127) >> 7;
    if (TMA && lane == 0) {
        tma_tile_2d(mytiles, &tmap, 0, row0, fb0);
        if (nst > 1) tma_tile_2d(mytiles + E3_TILE_BYTES, &tmap, 128, row0, fb0 + 8);
    }
    tma_wait(&bar);
    E3_STAMP(1);
    uint8_t *slot = slots + (live ? j : 0) * slot_stride;
    const int64_t thr = rc_thr(n);
    const uint32_t b1ref = n < 4 ? 4 : 4 + (uint32_t)((((size_t)n - 4) * 37) / 64);       // rccdf.c:126
    const uint32_t b1 = (b1ref + 64 + 15) & ~15u;                                          // coder 1 lives at slot + b1 + 4 (word 0 of its image = scratch)
    uint4 *gq = (uint4 *)(slot + (c ? b1 : 0));                                            // image of this coder in the slot: [scratch word][stream words ...]
    const uint32_t ringlane = ring0 + threadIdx.x * 4;                                     // ring: word-major, lane-minor
    constexpr uint32_t rs = E3_RING_S;
    const uint32_t progaddr = smem_u32(s_prog);
    const uint32_t psel = c ? 0x4341u : 0x4240u;                                         // byte_perm selector: this coder's two symbols of a word
    // own half of OVERFLOWI (rccdf.c:46,133) as a word count: coder 1 fires when b1ref + 4 wr >= thr, coder 0 when 4 + 4 wr >= b1ref
    const int64_t lim64 = c ? (thr - (int64_t)b1ref + 3) >> 2 : ((int64_t)b1ref - 4 + 3) >> 2;
    const uint32_t limw = lim64 < 0 ? 0u : (uint32_t)lim64;
    RcE96 e; e.init();
    bool raw = n < 4 || !live;
    uint32_t wr = 0, dr = 0;                                                               // words put / words drained to the slot
    const uint8_t *ip = in + (live ? j : 0) * g.chunk;
    const uint32_t lrow = lane >> 1;                                                       // row of this lane inside the warp tile
    const uint32_t rowaddr = mytiles + lrow * 128, sw = (lrow & 7) << 4;
    const bool ldg = !TMA || is_tail;                                                      // the short last call is not part of the tensor: its two lanes load directly
    uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
    if (ldg && nb && !raw) nxt = ldg128(ip);
    // after every block (or single symbol, in the remainder of a short call): count the new words, test this coder's half of
    // OVERFLOWI (the tested quantity only grows, so testing less often than the reference decides the same), and let four
    // finished words (ring image positions dr .. dr+3) leave as one 128-bit store
    auto account = [&](uint32_t k0) {
        wr += (e.k - k0) >> 29;
        raw |= wr >= limw;
        if (wr - dr >= 4) {
            const uint32_t ra = ringlane + (dr & 4) * rs;
            uint4 v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.x) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.y) : "r"(ra + rs));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.z) : "r"(ra + 2 * rs));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.w) : "r"(ra + 3 * rs));
            e.nz = min(min(e.nz, min(v.x, v.y)), min(v.z, v.w));                           // zero word => walk-back case (see RcE96)
            if (!raw) { gq[dr >> 2] = v; dr += 4; } else wr = dr + (wr & 3);               // a raw lane keeps coding but stops storing (its image is never read)
        }
    };
#pragma unroll 1
    for (uint32_t st = 0; st < nst; st++) {
        if (TMA) mbar_wait(fb0 + (st & 1) * 8, (st >> 1) & 1);
        const uint32_t cmax = min(8u, nbmax - st * 8);
#pragma unroll 1
        for (uint32_t cc = 0; cc < cmax; cc++) {
            const uint32_t b = st * 8 + cc;
            if (TMA) {
                const uint32_t a = (rowaddr + (st & 1) * E3_TILE_BYTES) | ((cc << 4) ^ sw);
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(cur.x), "=r"(cur.y), "=r"(cur.z), "=r"(cur.w) : "r"(a));
            }
            if (ldg) { cur = nxt; if (b + 1 < nb && !raw) nxt = ldg128(ip + (size_t)(b + 1) * 16); }
            if (b < nb) {                                                                  // (false only for the lanes of a short last call)
                const uint32_t w[4] = { cur.x, cur.y, cur.z, cur.w };
                uint32_t tx[8], ty[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {                                              // two symbols per word: entry offsets 8 * byte, table base OR-ed / added in
                    const uint32_t d8 = __byte_perm(w[q], 0, psel) << 3;
                    uint32_t a0, a1;
                    asm("lop3.b32 %0, %1, 0xffff, %2, 0xEA;" : "=r"(a0) : "r"(d8), "r"(tb));   // (d8 & 0xffff) | tb
                    asm("mad.hi.u32 %0, %1, 65536, %2;" : "=r"(a1) : "r"(d8), "r"(tb));         // (d8 >> 16) + tb
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tx[2 * q]), "=r"(ty[2 * q]) : "r"(a0));
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tx[2 * q + 1]), "=r"(ty[2 * q + 1]) : "r"(a1));
                }
                const uint32_t k0 = e.k;
#pragma unroll
                for (int q = 0; q < 8; q++) e.encode(tx[q], ty[q], ringlane);
                account(k0);
            }
            // Pacing.  The warp scheduler is not fair (it favours some warp slots), so identical warps drift apart: without this
            // the first warp of a CTA finished its chunk after 64 us and the last after 126 us.  Every fourth block a warp
            // publishes its block count and naps while it is more than E3_LEAD blocks ahead of the slowest warp of the CTA: the
            // issue slots go to the laggards and all warps reach the layout epilogue together.  (The slowest never waits.)
            if ((b & 3) == 3) {
                if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(progaddr + wid * 4), "r"(b) : "memory");
                for (;;) {
                    uint32_t pv;
                    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(pv) : "r"(progaddr + lane * 4) : "memory");
                    if (b <= __reduce_min_sync(0xffffffffu, pv) + E3_LEAD) break;
                    __nanosleep(256);
                }
            }
        }
        if (TMA) {
            __syncwarp();
            if (lane == 0 && st + 2 < nst) tma_tile_2d(mytiles + (st & 1) * E3_TILE_BYTES, &tmap, (int)(st + 2) * 128, row0, fb0 + (st & 1) * 8);
        }
    }
    if (is_tail && !raw) {                                                                 // short last call: the pairs beyond its last 16-byte block (rccdf.c:129-134)
        for (uint32_t i = nb * 16 + c; i < (n & ~1u); i += 2) {
            const uint2 t = ctab[ip[i]];
            const uint32_t k0 = e.k;
            e.encode(t.x, t.y, ringlane);
            account(k0);
        }
    }
    if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(progaddr + wid * 4), "r"(0xffffffffu) : "memory");
    E3_STAMP(2);
    raw = __shfl_xor_sync(0xffffffffu, (int)raw, 1) || raw;                                // either half fired -> raw copy
    if (!raw) {
        if (is_tail && c == 0 && (n & 1)) {                                                // odd tail on coder 0 (rccdf.c:135-136); at most one more word
            const uint2 t = ctab[ip[n - 1]];
            const uint32_t k0 = e.k;
            e.encode(t.x, t.y, ringlane);
            wr += (e.k - k0) >> 29;
            if (wr - dr >= 4) {
                const uint32_t ra = ringlane + (dr & 4) * rs;
                uint4 v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.x) : "r"(ra));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.y) : "r"(ra + rs));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.z) : "r"(ra + 2 * rs));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.w) : "r"(ra + 3 * rs));
                e.nz = min(min(e.nz, min(v.x, v.y)), min(v.z, v.w));
                gq[dr >> 2] = v; dr += 4;
            }
        }
        // rceflush turborc_.h:118-128
        if (e.rh == 0) { e96_put(e, e.lh, ringlane, wr); e.lh = e.ll; e.ll = 0; e.rh = e.rl; e.rl = 0; }
        if (e.rh > 2u || (e.rh == 2u && e.rl != 0)) { e96_add(e, 0, 1); e96_put(e, e.lh, ringlane, wr); }   // range > 2^33
        else { e96_add(e, 1, 0); e96_put(e, e.lh, ringlane, wr); e96_put(e, e.ll, ringlane, wr); }
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(ringlane + (wr & (E3_RING_W - 1)) * rs), "r"(e.pend) : "memory");
        e.nz = min(e.nz, e.pend);
        // image positions dr .. wr are still in the ring (<= 3 + 3 + 1 words): two more 128-bit stores cover them
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t ra = ringlane + ((dr + 4 * h) & 4) * rs;
            uint4 v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.x) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.y) : "r"(ra + rs));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.z) : "r"(ra + 2 * rs));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.w) : "r"(ra + 3 * rs));
            const uint32_t p0 = dr + 4 * h;                                                // words put by the loop but not drained yet: zero check on the valid ones
            if (p0 <= wr) e.nz = min(e.nz, v.x);
            if (p0 + 1 <= wr) e.nz = min(e.nz, v.y);
            if (p0 + 2 <= wr) e.nz = min(e.nz, v.z);
            if (p0 + 3 <= wr) e.nz = min(e.nz, v.w);
            if (p0 <= wr) gq[(dr >> 2) + h] = v;
        }
    }
    const uint32_t mypos = wr * 4, other = __shfl_xor_sync(0xffffffffu, mypos, 1);         // stream bytes of this coder / of its partner
    // words still in the ring when the loop ended were checked on their way out (e96_put) or are checked here
    const uint32_t rare = (e.nz == 0 ? 1u : 0u) | __shfl_xor_sync(0xffffffffu, e.nz == 0 ? 1u : 0u, 1);
    if (c == 0 && r < E3_MAX_NT / 2) {
        UnitMeta m; m.a_off = 0; m.a_len = 0; m.b_off = b1 + 4; m.b_len = 0; m.len = 0; m.flags = 0;
        if (live) {
            const uint32_t p0 = mypos, p1 = other;
            if (!raw && (int64_t)(4 + p0 + p1) >= thr) raw = true;                         // rccdf.c:142
            if ((rare || (flags & 1u)) && !raw) {                                          // wrapped pending word (flags & 1: test hook): walk-back coder
                __shared__ uint32_t s_ctab32[256];
                // (only this lane needs the 32-bit table; built on the fly from the 8-byte entries)
                for (int x = 0; x < 256; x++) s_ctab32[x] = ctab[x].x | ctab[x].y << 16;
                rc_static_enc_call<2, true>(ip, n, s_ctab32, nullptr, slot, m);
                if (!(m.flags & UM_RAW)) {                                                 // move stream 1 to where the epilogue expects it
                    for (uint32_t k2 = m.b_len; k2 > 0; k2 -= 4) *(uint32_t *)(slot + b1 + 4 + k2 - 4) = *(uint32_t *)(slot + m.b_off + k2 - 4);
                    m.b_off = b1 + 4;
                }
            } else { m.a_len = raw ? 0 : 4 + p0; m.b_len = raw ? 0 : p1; m.len = raw ? n : 4 + p0 + p1; m.flags = raw ? UM_RAW : 0; }
        }
        s_len[r] = m.len; s_alen[r] = (m.flags & UM_RAW) ? 0xffffffffu : m.a_len; s_boff[r] = m.b_off; s_blen[r] = m.b_len;
    }
    E3_STAMP(3);
    __syncthreads();
    E3_STAMP(4);
    // ---- exclusive scan of the call lengths of this CTA (calls_per_cta <= 256: thread t scans entry t)
    uint32_t v = threadIdx.x < calls_per_cta ? s_len[threadIdx.x] : 0, inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (unsigned k = 0; k < (blockDim.x + 31) / 32; k++) { const uint32_t t = s_wsum[k]; s_wsum[k] = run; run += t; }
        s_wsum[NW] = run;
    }
    __syncthreads();
    if (threadIdx.x < calls_per_cta) s_excl[threadIdx.x] = s_wsum[wid] + inc - v;
    // ---- decoupled look-back for the CTA's base offset
    if (wid == 0) {
        const unsigned long long agg = s_wsum[NW];
        if (lane == 0) { lb[bid] = (bid == 0 ? LB_INC : LB_AGG) | agg; __threadfence(); }
        unsigned long long base = 0;
        if (bid) {
            long long idx = (long long)bid - 1;
            for (;;) {
                const long long k = idx - lane;
                unsigned long long st = LB_INC;                   // before CTA 0: inclusive prefix 0
                if (k >= 0) do { st = lb[k]; } while ((st >> 62) == 0);
                const unsigned incl = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const unsigned first = incl ? __ffs((int)incl) - 1 : 32;
                unsigned long long val = lane <= first ? (st & LB_VAL) : 0;
#pragma unroll
                for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                base += val;
                if (incl) break;
                idx -= 32;
            }
            if (lane == 0) { lb[bid] = LB_INC | (base + agg); __threadfence(); }
        }
        if (lane == 0) s_base = base;
    }
    __syncthreads();
    const unsigned long long base = s_base;
    E3_STAMP(5);
    if (threadIdx.x < calls_per_cta && j0 + threadIdx.x < n_calls) out_off[j0 + threadIdx.x] = base + s_excl[threadIdx.x];
    if (threadIdx.x == 0 && j0 + calls_per_cta >= n_calls) out_off[n_calls] = base + s_wsum[NW];
    // ---- layout: EIGHT lanes per call (four calls per warp at a time: the copy is latency-bound, so what counts is how many
    // loads each lane keeps in flight), pieces from the slot (or the input, for a raw call) to their final place
    const unsigned sub = lane & 7, grp = lane >> 3;
    for (unsigned q = wid * 4 + grp; q < calls_per_cta && j0 + q < n_calls; q += (blockDim.x >> 5) * 4) {
        uint8_t *dst = out + base + s_excl[q];
        uint8_t *sl = slots + (j0 + q) * slot_stride;
        if (s_alen[q] == 0xffffffffu) { group_copy4(dst, in + (j0 + q) * g.chunk, s_len[q], sub, 8); continue; }
        if (((uintptr_t)dst & 3) == 0) {
            const uint32_t *pa = (const uint32_t *)sl, *pb = (const uint32_t *)(sl + s_boff[q]);
            const uint32_t wa = s_alen[q] >> 2, W = wa + (s_blen[q] >> 2), len0 = s_alen[q] - 4;   // word 0 = the len0 header (rccdf.c:141)
            uint32_t *d = (uint32_t *)dst;
            for (uint32_t w0 = 0; w0 < W; w0 += E3_COPY_U * 8) {
                uint32_t vv[E3_COPY_U];
#pragma unroll
                for (int k = 0; k < E3_COPY_U; k++) { const uint32_t w = w0 + k * 8 + sub; vv[k] = w == 0 ? len0 : (w < wa ? pa[w] : (w < W ? pb[w - wa] : 0u)); }
#pragma unroll
                for (int k = 0; k < E3_COPY_U; k++) { const uint32_t w = w0 + k * 8 + sub; if (w < W) __stcs(d + w, vv[k]); }
            }
        } else {
            if (sub == 0) *(uint32_t *)sl = s_alen[q] - 4;
            __syncwarp(0xffu << (grp * 8));
            group_copy(dst, sl, s_alen[q], sub, 8); if (s_blen[q]) group_copy(dst + s_alen[q], sl + s_boff[q], s_blen[q], sub, 8);
        }
    }
    E3_STAMP(6);
}


// =========================================================================================================================
// TRC_RCS2 decoder (rccdfsb2dec, rccdf.c:166-184), lane per coder.  Same plan as k_rcs2_dec_lpc (speculative symbol from an fp32
// estimate of code / range, exact 64-bit verification, per-block redo by the reference's binary search) with a shorter step:
//   * ONE unsigned test covers both ways the estimate can be wrong: x is right  <=>  0 <= code - cdf[x]*range < freq[x]*range.
//     If x is too high the subtraction wraps to >= 2^64 - freq[x-1]*range, which is still >= freq[x]*range because
//     (freq[x-1] + freq[x]) * range <= 2^15 * range < 2^64; so `code - rp >= fr` (unsigned) flags both cases.
//   * floor() of the quotient comes out of the FMA itself (round-down mode into the 2^23 integer grid): no separate add.
//   * {cdf, freq} is one 8-byte table entry; the stream ring is addressed like the encoder's (cursor in the top bits of a
//     register: wraps for free, address = one shift-and-add); one word of look-ahead instead of two.
// =========================================================================================================================
constexpr int      D3_RING_W = 16;
constexpr uint32_t D3_RING_S = 2048;                    // bytes between consecutive ring words of a lane (512 lanes)
struct __align__(16) DecTab2 { uint2 e[256]; uint8_t lut[PROB_TOTAL]; };   // {cdf, freq} per symbol; slot -> symbol

__global__ void __launch_bounds__(1024)
k_build_dectab2(const cdf_t *__restrict__ cdf, unsigned cdfnum, DecTab2 *__restrict__ ts) {
    __shared__ uint16_t scdf[CDF_STRIDE];
    const cdf_t *c0 = cdf + (size_t)blockIdx.x * CDF_STRIDE;
    DecTab2 &t = ts[blockIdx.x];
    for (unsigned x = threadIdx.x; x <= cdfnum; x += blockDim.x) scdf[x] = c0[x];
    __syncthreads();
    if (blockIdx.y == 0) {
        for (unsigned x = threadIdx.x; x < 256; x += blockDim.x) {
            uint32_t c = 0, f = 0;
            if (x < cdfnum) { c = scdf[x]; f = (uint32_t)scdf[x + 1] - c; }
            t.e[x] = make_uint2(c, f);
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

struct RcD3 {
    uint32_t rl, rh, cl, ch;        // range, code
    uint32_t n0;                    // the next stream word (already in a register)
    uint32_t k;                     // ring cursor: (index of the word after n0, mod 16) << 28
    bool bad;
    __device__ __forceinline__ uint32_t step(const uint8_t *lut, const uint2 *dtab, uint32_t ringlane) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;                    // _rccdfrange
        float qf;                                                                     // floor(code / range) in the low mantissa bits (fix-up by verification)
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint2 e = dtab[x];
        const uint64_t rp = (uint64_t)rl * e.x, fr = (uint64_t)rl * e.y;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * e.x;       // cdf[x] * range
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * e.y;       // freq[x] * range
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);               // not (0 <= code - rp < fr): wrong symbol
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));       // (k >> 28) * 2048 + lane base
        asm volatile("{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"           // _rcdnorm_ turborc_.h:111
            "selp.u32 %0, %6, %7, p;\n\t"         // rh = p ? fl : fh
            "selp.u32 %1, 0, %6, p;\n\t"          // rl = p ? 0 : fl
            "selp.u32 %2, %8, %9, p;\n\t"         // ch = p ? dl : dh
            "selp.u32 %3, %4, %8, p;\n\t"         // cl = p ? n0 : dl
            "@p ld.shared.u32 %4, [%10];\n\t"
            "@p add.u32 %5, %5, 0x10000000;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        return x;
    }
};

// dynamic shared memory: [DecTab2: 2 KB + 32 KB][ring: 16 words x 2 KB]
constexpr size_t D3_SMEM = sizeof(DecTab2) + (size_t)D3_RING_W * D3_RING_S;

__global__ void __launch_bounds__(512, 2)
k_rcs2_dec3(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint8_t *__restrict__ out, Geom g,
            size_t n_calls, const DecTab2 *__restrict__ ts, unsigned cdfnum, size_t cpc, unsigned calls_per_cta) {
    extern __shared__ __align__(16) uint8_t dyn[];
    DecTab2 *tabs = (DecTab2 *)dyn;
    const uint2 *dtab = tabs->e;
    const uint8_t *lut = tabs->lut;
    uint32_t *ringbuf = (uint32_t *)(dyn + sizeof(DecTab2));
    __shared__ uint64_t bar;
    const size_t j0 = (size_t)blockIdx.x * calls_per_cta, j = j0 + (threadIdx.x >> 1);
    const unsigned c = threadIdx.x & 1;
    unsigned long long *tp = g_e3_times ? g_e3_times + ((size_t)blockIdx.x * 32 + (threadIdx.x >> 5)) * 8 : nullptr;   // phase probe (tools/enc_phases.py --dec)
    const unsigned lane = threadIdx.x & 31;
    E3_STAMP(0);
    if (threadIdx.x == 0) tma_fetch(tabs, ts + (cpc ? j0 / cpc : 0), (uint32_t)sizeof(DecTab2), &bar);
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
    // word addressing relative to the 16-byte aligned base below p; reads past the end of the packed buffer return zero
    const uint32_t *qbase = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)15);
    const long long wavail = ((const uint8_t *)gend - (const uint8_t *)qbase) >> 2;
    const uint32_t wlim = wavail > 0 ? (uint32_t)(wavail - 1) : 0u;
    const bool none = wavail <= 0;
    auto gword = [&](uint32_t w) -> uint32_t { return (!none && w <= wlim) ? __ldg(qbase + w) : 0u; };
    auto gquad = [&](uint32_t w) -> uint4 {                                           // w multiple of 4
        if (!none && w + 3 <= wlim) return __ldg((const uint4 *)(qbase + w));
        return make_uint4(gword(w), gword(w + 1), gword(w + 2), gword(w + 3));
    };
    uint32_t *ring = ringbuf + threadIdx.x;
    constexpr uint32_t RS = D3_RING_S / 4;
    const uint32_t ringlane = smem_u32(ring);
    auto ring_put = [&](uint32_t w, const uint4 &v) {                                 // w multiple of 4
        ring[((w + 0) & (D3_RING_W - 1)) * RS] = v.x; ring[((w + 1) & (D3_RING_W - 1)) * RS] = v.y;
        ring[((w + 2) & (D3_RING_W - 1)) * RS] = v.z; ring[((w + 3) & (D3_RING_W - 1)) * RS] = v.w;
    };
    RcD3 d;
    uint32_t fi, ci;                                                                  // words [fi-16, fi) are in the ring; ci = index of the word after n0
    auto resync = [&](uint32_t w0, uint64_t range, uint64_t code) {                   // (re)start the ring at word w0 = next unread word
        fi = w0 & ~3u;
        for (int q = 0; q < 3; q++) { ring_put(fi, gquad(fi)); fi += 4; }
        d.rl = (uint32_t)range; d.rh = (uint32_t)(range >> 32); d.cl = (uint32_t)code; d.ch = (uint32_t)(code >> 32);
        d.n0 = ring[(w0 & (D3_RING_W - 1)) * RS];
        ci = w0 + 1; d.k = ci << 28; d.bad = false;
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
            const bool need = fi - ci <= 9;                                           // top the ring up (stores happen after the block)
            uint4 t4 = make_uint4(0, 0, 0, 0);
            if (need) t4 = gquad(fi);
            const RcD3 s0 = d;                                                        // block start state (for the redo path)
#pragma unroll
            for (int q = 0; q < 4; q++) a0 |= d.step(lut, dtab, ringlane) << (8 * q);
#pragma unroll
            for (int q = 0; q < 4; q++) a1 |= d.step(lut, dtab, ringlane) << (8 * q);
            const uint32_t ci_new = ci + ((d.k - s0.k) >> 28);                        // <= 4 words per block
            const bool dry = ci_new > fi;                                             // a read ran past the filled part of the ring
            if (need) { ring_put(fi, t4); fi += 4; }
            if (__builtin_expect(d.bad || dry, 0)) {                                  // estimate missed (or ring ran dry): exact redo
                RcDExact ex;
                ex.range = (uint64_t)s0.rh << 32 | s0.rl; ex.code = (uint64_t)s0.ch << 32 | s0.cl;
                ex.base = qbase; ex.wlim = wlim; ex.wi = none ? 1u : ci - 1;         // n0 was word ci-1
                a0 = a1 = 0;
                for (int q = 0; q < 4; q++) a0 |= ex.step2(dtab, cdfnum) << (8 * q);
                for (int q = 0; q < 4; q++) a1 |= ex.step2(dtab, cdfnum) << (8 * q);
                resync(ex.wi, ex.range, ex.code);
            } else ci = ci_new;
        }
        // a0/a1 = this coder's symbols 0-3 / 4-7 of the block; interleave with the partner's
        uint32_t b0 = __shfl_xor_sync(0xffffffffu, a0, 1), b1 = __shfl_xor_sync(0xffffffffu, a1, 1);
        if (i < nb) {
            uint32_t e0 = c ? b0 : a0, o0 = c ? a0 : b0, e1 = c ? b1 : a1, o1 = c ? a1 : b1;   // even-position / odd-position symbols
            uint2 v = c ? make_uint2(__byte_perm(e1, o1, 0x5140), __byte_perm(e1, o1, 0x7362))
                        : make_uint2(__byte_perm(e0, o0, 0x5140), __byte_perm(e0, o0, 0x7362));
            *(uint2 *)(op + i + 8 * c) = v;
        }
    }
    E3_STAMP(2);
    // remaining full pairs, then the odd tail on coder 0 (rccdf.c:179-182): exact path
    if (n > nb) {
        RcDExact ex;
        ex.range = (uint64_t)d.rh << 32 | d.rl; ex.code = (uint64_t)d.ch << 32 | d.cl;
        ex.base = qbase; ex.wlim = wlim; ex.wi = none ? 1u : ci - 1;
        for (size_t i = nb + c; i < (n & ~(size_t)1); i += 2) op[i] = (uint8_t)ex.step2(dtab, cdfnum);
        if (c == 0 && (n & 1)) op[n - 1] = (uint8_t)ex.step2(dtab, cdfnum);
    }
}

}  // namespace trc
