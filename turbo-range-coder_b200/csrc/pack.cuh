// pack.cuh -- everything around the coders: per-call length/raw resolution + prefix sum (k_resolve_scan),
// the layout pass that writes each call's bytes exactly where and how the reference does (k_pack), and the
// static-table builder cdfini (rccdf.c:50-68) as histogram + normalise kernels.
#pragma once
#include "trc_common.cuh"

namespace trc {

struct CallInfo { uint32_t len; uint32_t raw; };   // raw: 0 coded, 1 raw copy of `len` input bytes

// ---- resolve + scan ------------------------------------------------------------------------------------
// Per call: combine its units.  For the blocked rANS codecs this applies the mnflush guards
// (anscdf_.h:131-136) which depend on where the block lands in the call's output: with o = bytes of the
// previous blocks, L = bytes of this block, N = inlen of the call, the reference bails out to a raw copy iff
//     o + L + (the last-coded record emitted a word ? 0 : 2)  >=  N
// (DESIGN.md derives this from the per-record guard `ep <= op + 2 + 4n` and the two post-checks).
// Then an exclusive prefix sum over the call lengths gives the packed offsets.  Single CTA.
constexpr int SCAN_NT = 1024;

// step 1 (any number of CTAs, one thread per call): call length + raw flag
__global__ void k_resolve(Geom g, int blocked, UnitMeta *__restrict__ meta, CallInfo *__restrict__ calls) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n_calls) return;
    CallInfo ci;
    if (blocked) {
        size_t cs, N; call_span(g, j, cs, N);
        uint64_t o = 0; bool raw = false;
        for (uint32_t b = 0; b < g.upc; b++) {
            UnitMeta &m = meta[j * g.upc + b];
            if (m.a_len == 0 && m.len == 0) break;                     // padding unit of a short last call
            if ((m.flags & UM_OVF) || o + m.len + ((m.flags & UM_ADJ2) ? 2 : 0) >= N) { raw = true; break; }
            m.pref = (uint32_t)o; o += m.len;
        }
        ci.raw = raw; ci.len = raw ? (uint32_t)N : (uint32_t)o;
    } else {
        const UnitMeta &m = meta[j];
        ci.raw = (m.flags & UM_RAW) ? 1 : 0; ci.len = m.len;
    }
    calls[j] = ci;
}

// step 2 (single CTA): exclusive prefix sum of the call lengths.  Each thread owns SCAN_PER consecutive calls of a
// SCAN_NT*SCAN_PER tile: its loads are issued together (one memory round trip per tile), the block-wide part is two
// shuffle scans, and its SCAN_PER offsets leave as 128-bit stores.
constexpr int SCAN_PER = 8;
__global__ void __launch_bounds__(SCAN_NT)
k_scan(size_t n, const CallInfo *__restrict__ calls, uint64_t *__restrict__ out_off) {
    __shared__ uint64_t wsum[SCAN_NT / 32];
    __shared__ uint64_t carry_s;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += (size_t)SCAN_NT * SCAN_PER) {
        const size_t j0 = base + (size_t)threadIdx.x * SCAN_PER;
        uint32_t v[SCAN_PER];
        if (j0 + SCAN_PER <= n) {
            const uint4 *p = (const uint4 *)(calls + j0);                   // 2 CallInfo per uint4 (calls is 256-byte aligned)
#pragma unroll
            for (int k = 0; k < SCAN_PER / 2; k++) { uint4 q = p[k]; v[2 * k] = q.x; v[2 * k + 1] = q.z; }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_PER; k++) v[k] = j0 + k < n ? calls[j0 + k].len : 0u;
        }
        uint64_t tot = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) tot += v[k];
        uint64_t x = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = wsum[lane], t = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
            wsum[lane] = t - w;                                             // exclusive warp offsets
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        uint64_t run = carry + wsum[wid] + x - tot;                         // exclusive offset of this thread's first call
        if (j0 + SCAN_PER <= n && ((uintptr_t)(out_off + j0) & 15) == 0) {
            uint64_t o[SCAN_PER];
#pragma unroll
            for (int k = 0; k < SCAN_PER; k++) { o[k] = run; run += v[k]; }
            ulonglong2 *d = (ulonglong2 *)(out_off + j0);
#pragma unroll
            for (int k = 0; k < SCAN_PER / 2; k++) d[k] = make_ulonglong2(o[2 * k], o[2 * k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_PER; k++) { if (j0 + k < n) out_off[j0 + k] = run; run += v[k]; }
        }
        __syncthreads();
        if (threadIdx.x == SCAN_NT - 1) carry_s = carry + wsum[wid] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_off[n] = carry_s;
}

// copies the offsets of a sub-batch into host-mapped pinned memory from the SMs: a cudaMemcpy of a few KB would queue
// on the D2H copy engine behind the multi-megabyte payload downloads of earlier sub-batches
__global__ void k_publish_u64(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst_mapped, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst_mapped[i] = src[i];
}

// ---- push: device-driven copy of a packed stream into (peer) memory ------------------------------------------
// The byte count lives in device memory (it is the encoder's out_off[n]), so no host round trip is needed to size
// the transfer.  dst may be a CUDA-IPC mapping of another GPU's buffer: the 128-bit stores then travel over
// NVLink/NVSwitch.  Completion protocol (a consumer on the destination GPU must never see a partial stream):
// every CTA fences its stores system-wide and bumps a counter in LOCAL memory; the CTA that finds the counter complete
// publishes the length, fences again, and only then writes the sequence number into the destination's flag word.  The
// consumer (k_wait_flags on the destination) spins on the flag with volatile loads: flag >= seq  =>  length and payload of
// push `seq` have landed.  A stream longer than the slot is not copied beyond `cap`: the published length keeps the TRUE size
// with its top bit set, which the consumer reports as an overflow.  A caller that can predict the size (the previous step's, say)
// lets a copy engine move that many bytes first (cudaMemcpyAsync on the same stream) and passes them as `skip`: the kernel then only
// moves the remainder and publishes -- the SMs stay with the coders.
constexpr unsigned long long PUSH_OVERFLOW = 1ull << 63;
__global__ void __launch_bounds__(256)
k_push(uint4 *__restrict__ dst, const uint4 *__restrict__ src, const uint64_t *__restrict__ d_len, size_t fixed_len, size_t cap,
       uint64_t *__restrict__ dst_len, volatile uint64_t *__restrict__ dst_flag, uint64_t seq, unsigned int *__restrict__ counter,
       const volatile uint64_t *__restrict__ ack, uint64_t ack_need, size_t skip) {
    // back-pressure: the slot may be overwritten only after the consumer acknowledged the push that used it last
    if (ack && threadIdx.x == 0) while (*ack < ack_need) __nanosleep(128);
    if (ack) __syncthreads();
    const uint64_t len = d_len ? *d_len : (uint64_t)fixed_len;
    const uint64_t clen = cap && len > cap ? cap : len;
    // `skip` bytes (a multiple of 16) were already moved by a copy engine: only what lies beyond them is copied here
    const size_t nv = (size_t)((clen + 15) >> 4), stride = (size_t)gridDim.x * blockDim.x;
    // four 16-byte chunks in flight per thread: the kernel runs on a few CTAs next to the decoder, so the bandwidth over
    // NVLink has to come from memory-level parallelism per thread, not from thread count
    size_t i = (skip >> 4) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nv; i += 4 * stride) {
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < nv; i += stride) dst[i] = src[i];
    __threadfence_system();                                           // this thread's stores are visible system-wide ...
    __syncthreads();                                                  // ... for every thread of the CTA
    if (threadIdx.x == 0) {
        const unsigned done = counter ? atomicAdd(counter, 1u) : gridDim.x - 1;
        if (done == gridDim.x - 1) {                                  // (no counter: no flag either, every CTA just repeats the length)
            if (counter) *counter = 0;                                // ready for the next push (pushes are stream-ordered)
            if (dst_len) *dst_len = len == clen ? len : (len | PUSH_OVERFLOW);
            __threadfence_system();
            if (dst_flag) *dst_flag = seq;                            // release: everything above is visible before the flag
        }
    }
}

// consumer side of the protocol: returns (kernel completes) once every flag[r] >= seq.  status[0] |= 1 when a published
// length carries the overflow bit.  One thread per flag.
__global__ void k_wait_flags(const volatile uint64_t *__restrict__ flags, const volatile uint64_t *__restrict__ lens, unsigned n, uint64_t seq,
                             unsigned int *__restrict__ status) {
    const unsigned r = threadIdx.x;                                   // one CTA, one thread per producer
    if (r < n) {
        while (flags[r] < seq) __nanosleep(64);
        __threadfence_system();                                       // acquire: payload / length reads after this see the pushed data
        if (status && lens && (lens[r] & PUSH_OVERFLOW)) atomicOr(status, 1u);
    }
}
// consumer is done with the slot set of push `seq`: producers may reuse it (k_push spins on this word over the peer mapping)
__global__ void k_ack(volatile uint64_t *__restrict__ ack, uint64_t seq) { if (threadIdx.x == 0) { __threadfence_system(); *ack = seq; } }

// ---- pack ----------------------------------------------------------------------------------------------
// grid = (n_units, segments); each CTA moves one `seg`-byte segment of one unit's output.
constexpr int    PACK_NT  = 128;
constexpr size_t PACK_SEG_MIN = 16384;

__global__ void __launch_bounds__(PACK_NT)
k_pack(const uint8_t *__restrict__ in, Geom g, const uint8_t *__restrict__ slots, size_t slot_stride,
       const UnitMeta *__restrict__ meta, const CallInfo *__restrict__ calls, const uint64_t *__restrict__ out_off,
       uint8_t *__restrict__ out, size_t PACK_SEG) {
    const size_t u = blockIdx.x;
    size_t j, start, len; uint32_t b;
    unit_span(g, u, j, b, start, len);
    if (len == 0) return;
    const CallInfo ci = calls[j];
    const size_t seg0 = (size_t)blockIdx.y * PACK_SEG;
    if (ci.raw) {                                                       // raw copy of (a prefix of) the call's input
        size_t cs = j * g.chunk, uo = start - cs;                      // this unit's share of the call
        if (uo >= ci.len) return;
        size_t n = ci.len - uo < len ? ci.len - uo : len;
        // rccdf4ienc quirk on inputs shorter than 4 bytes: the returned length (4) exceeds inlen; the reference
        // leaves those bytes untouched, we zero them so the packed stream is deterministic
        size_t ccs, N; call_span(g, j, ccs, N);
        if (ci.len > N && b == 0 && blockIdx.y == 0 && threadIdx.x == 0)
            for (size_t k = N; k < ci.len; k++) out[out_off[j] + k] = 0;
        if (seg0 >= n) return;
        size_t m = n - seg0 < PACK_SEG ? n - seg0 : PACK_SEG;
        group_copy(out + out_off[j] + uo + seg0, in + start + seg0, m, threadIdx.x, PACK_NT);
        return;
    }
    const UnitMeta m = meta[u];
    const size_t total = (size_t)m.a_len + m.b_len;
    if (seg0 >= total) return;
    size_t seg1 = seg0 + PACK_SEG < total ? seg0 + PACK_SEG : total;
    uint8_t *dst = out + out_off[j] + m.pref;
    const uint8_t *slot = slots + u * slot_stride;
    if (seg0 < m.a_len) {                                               // part of piece a
        size_t e = seg1 < m.a_len ? seg1 : m.a_len;
        group_copy(dst + seg0, slot + m.a_off + seg0, e - seg0, threadIdx.x, PACK_NT);
    }
    if (seg1 > m.a_len) {                                               // part of piece b
        size_t s = seg0 > m.a_len ? seg0 : m.a_len;
        group_copy(dst + s, slot + m.b_off + (s - m.a_len), seg1 - s, threadIdx.x, PACK_NT);
    }
}

// small units (one segment each): a warp per unit, 8 units per CTA -- the per-CTA launch cost of the general kernel
// would dominate a 3 KB copy
constexpr int PACKS_WARPS = 8;
__global__ void __launch_bounds__(PACKS_WARPS * 32)
k_pack_small(const uint8_t *__restrict__ in, Geom g, const uint8_t *__restrict__ slots, size_t slot_stride,
             const UnitMeta *__restrict__ meta, const CallInfo *__restrict__ calls, const uint64_t *__restrict__ out_off,
             uint8_t *__restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const size_t u = (size_t)blockIdx.x * PACKS_WARPS + (threadIdx.x >> 5);
    if (u >= g.n_units) return;
    size_t j, start, len; uint32_t b;
    unit_span(g, u, j, b, start, len);
    if (len == 0) return;
    const CallInfo ci = calls[j];
    if (ci.raw) {
        size_t cs = j * g.chunk, uo = start - cs;
        if (uo >= ci.len) return;
        size_t n = ci.len - uo < len ? ci.len - uo : len;
        size_t ccs, N; call_span(g, j, ccs, N);
        if (ci.len > N && b == 0 && lane == 0) for (size_t k = N; k < ci.len; k++) out[out_off[j] + k] = 0;   // rccdf4ienc quirk
        group_copy(out + out_off[j] + uo, in + start, n, lane, 32);
        return;
    }
    const UnitMeta m = meta[u];
    uint8_t *dst = out + out_off[j] + m.pref;
    const uint8_t *slot = slots + u * slot_stride;
    if (m.a_len) group_copy(dst, slot + m.a_off, m.a_len, lane, 32);
    if (m.b_len) group_copy(dst + m.a_len, slot + m.b_off, m.b_len, lane, 32);
}

// ---- cdfini --------------------------------------------------------------------------------------------
constexpr int    HIST_NT  = 256;
constexpr size_t HIST_SEG = 1 << 20;

// grid = (n_calls, segments of HIST_SEG bytes); hist = n_calls * 256 u64, zeroed by the caller
__global__ void __launch_bounds__(HIST_NT)
k_hist(const uint8_t *__restrict__ in, Geom g, unsigned long long *__restrict__ hist) {
    __shared__ unsigned int h[256];
    size_t cs, N; call_span(g, blockIdx.x, cs, N);
    size_t s0 = (size_t)blockIdx.y * HIST_SEG;
    if (s0 >= N) return;
    size_t s1 = s0 + HIST_SEG < N ? s0 + HIST_SEG : N;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t *p = in + cs;
    size_t a0 = s0, a1 = s1;
    while (a0 < a1 && ((uintptr_t)(p + a0) & 15)) a0++;                 // unaligned head
    size_t nv = (a1 - a0) >> 4;
    for (size_t i = s0 + threadIdx.x; i < a0; i += HIST_NT) atomicAdd(&h[p[i]], 1u);
    const uint4 *v = (const uint4 *)(p + a0);
    for (size_t i = threadIdx.x; i < nv; i += HIST_NT) {
        uint4 q = v[i];
        uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&h[w[k] & 0xff], 1u); atomicAdd(&h[(w[k] >> 8) & 0xff], 1u);
            atomicAdd(&h[(w[k] >> 16) & 0xff], 1u); atomicAdd(&h[w[k] >> 24], 1u);
        }
    }
    for (size_t i = a0 + (nv << 4) + threadIdx.x; i < a1; i += HIST_NT) atomicAdd(&h[p[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[(size_t)blockIdx.x * 256 + threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// one thread per call: the serial normalisation of rccdf.c:55-67
__global__ void k_cdf_finalize(Geom g, const unsigned long long *__restrict__ hist, cdf_t *__restrict__ cdf, unsigned cdfnum,
                               int *__restrict__ status) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n_calls) return;
    size_t cs, N; call_span(g, j, cs, N);
    const unsigned long long *h = hist + j * 256;
    cdf_t *c = cdf + j * CDF_STRIDE;
    unsigned long long mx = 0, cum = 0; unsigned mxi = 0;
    for (unsigned i = 0; i < cdfnum; i++) {
        unsigned long long v = (h[i] << PROB_BITS) / N;
        if (!v) v = 1;
        cum += v;
        if (v > mx) { mx = v; mxi = i; }
    }
    unsigned long long acc = 0; int bad = 0;
    c[0] = 0;
    for (unsigned i = 0; i < cdfnum; i++) {
        unsigned long long v = (h[i] << PROB_BITS) / N;
        if (!v) v = 1;
        if (i == mxi) v -= cum - PROB_TOTAL;                            // adjust max (rccdf.c:60), wraps like the reference
        unsigned prev = (cdf_t)acc;
        acc += v;
        c[i + 1] = (cdf_t)acc;
        if (prev >= (cdf_t)acc) bad = 1;                                // the reference die()s here (rccdf.c:65)
    }
    if ((cdf_t)acc != (cdf_t)PROB_TOTAL) bad = 1;                       // rccdf.c:66
    // a symbol outside the alphabet would be coded with an all-zero table entry (the reference leaves that to its caller:
    // turborc.c:535 `if(m<16)`): report it like a degenerate table
    for (unsigned i = cdfnum; i < 256; i++) if (h[i]) bad = 1;
    for (unsigned i = cdfnum + 1; i < (unsigned)CDF_STRIDE; i++) c[i] = 0;       // the unused tail of the 257-entry row is defined (it travels in the container)
    if (status) status[j] = bad ? -1 : 0;
}

}  // namespace trc
