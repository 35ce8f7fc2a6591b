// dec_step.cu -- micro-benchmark of range-coder DECODE step formulations (tools only; not part of the library).
// Same harness for every variant: 296 CTAs x 384 lanes, each lane decodes NSYM symbols from its own pseudo-random word stream
// (random code words decode to symbols distributed exactly like the model, so renormalisation frequency is realistic),
// stream words through a 16-word shared-memory ring per lane, 32 KB slot->symbol LUT + 256-entry table in shared memory.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <type_traits>
#include <cuda_runtime.h>
#ifndef T_LANES
#define T_LANES 384
#endif
constexpr int T = T_LANES;
static int g_idx = 0;
constexpr int PROB_BITS = 15;
constexpr int RING_W = 16;
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- variant 0: round-1 RcDRing::step (64-bit int->float conversions, per-step borrow/too-low flags, ring index by mask*stride)
struct D0 {
    uint32_t rl, rh, cl, ch, n0, n1, ci, bad;
    const uint8_t *lut; const uint32_t *dtab; const uint32_t *ring;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const float qf = fmaf(__ull2float_rz((uint64_t)ch << 32 | cl), rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl)), -0.5f);
        const uint32_t q = __float_as_uint(qf + 12582912.0f) & 0x7fffu;
        const uint32_t x = lut[q], e = dtab[x];
        const uint32_t c0 = e >> 16, f = e & 0xffffu;
        const uint32_t pl = rl * c0, ph = __umulhi(rl, c0) + rh * c0;
        const uint32_t fl = rl * f, fh = __umulhi(rl, f) + rh * f;
        uint32_t dl, dh, bw;
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;" : "=r"(dl), "=r"(dh), "=r"(bw) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= bw | ((dh > fh || (dh == fh && dl >= fl)) ? 1u : 0u);
        const bool p = fh == 0;
        rh = p ? fl : fh; rl = p ? 0u : fl;
        ch = p ? dl : dh; cl = p ? n0 : dl;
        n0 = p ? n1 : n0;
        if (p) n1 = ring[(ci & (RING_W - 1)) * T];
        ci += p ? 1u : 0u;
        x_out = x;
    }
};
// ---- variant 1: 32-bit conversions on the fma pipe, one unsigned test (code - rp < fr) covers both error directions,
//      {cdf, freq} as one 8-byte table entry, ring cursor = a shared-memory address advanced by add + bit-merge
constexpr uint32_t RSTRIDE = 2048;                                   // bytes between consecutive ring words of a lane (512 lanes x 4)
struct D1 {
    uint32_t rl, rh, cl, ch, n0, n1, ra; bool bad;                    // ra = byte offset of the next ring word inside dyn[]
    const uint8_t *lut; const uint2 *dtab2; const uint8_t *dyn;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const float cf = fmaf(__uint2float_rz(ch), 4294967296.0f, __uint2float_rz(cl));
        const float rf = fmaf(__uint2float_rn(rh), 4294967296.0f, __uint2float_rn(rl));
        const float qf = fmaf(cf, rcp_approx(rf), -0.5f);
        const uint32_t q = __float_as_uint(qf + 12582912.0f) & 0x7fffu;
        const uint32_t x = lut[q];
        const uint2 e = dtab2[x];
        const uint32_t c0 = e.x, f = e.y;
        const uint64_t rp = (uint64_t)rl * c0, fr = (uint64_t)rl * f;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * c0;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * f;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        const bool p = fh == 0;
        rh = p ? fl : fh; rl = p ? 0u : fl;
        ch = p ? dl : dh; cl = p ? n0 : dl;
        n0 = p ? n1 : n0;
        if (p) n1 = *(const uint32_t *)(dyn + ra);
        const uint32_t t = ra + (p ? RSTRIDE : 0u);
        ra = (ra & ~(RSTRIDE * (RING_W - 1))) | (t & (RSTRIDE * (RING_W - 1)));
        x_out = x;
    }
};
// ---- variant 2: variant 1 with a one-word look-ahead (n0 only)
struct D2 {
    uint32_t rl, rh, cl, ch, n0, ra; bool bad;
    const uint8_t *lut; const uint2 *dtab2; const uint8_t *dyn;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const float cf = fmaf(__uint2float_rz(ch), 4294967296.0f, __uint2float_rz(cl));
        const float rf = fmaf(__uint2float_rn(rh), 4294967296.0f, __uint2float_rn(rl));
        const float qf = fmaf(cf, rcp_approx(rf), -0.5f);
        const uint32_t q = __float_as_uint(qf + 12582912.0f) & 0x7fffu;
        const uint32_t x = lut[q];
        const uint2 e = dtab2[x];
        const uint32_t c0 = e.x, f = e.y;
        const uint64_t rp = (uint64_t)rl * c0, fr = (uint64_t)rl * f;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * c0;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * f;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        const bool p = fh == 0;
        rh = p ? fl : fh; rl = p ? 0u : fl;
        ch = p ? dl : dh; cl = p ? n0 : dl;
        if (p) n0 = *(const uint32_t *)(dyn + ra);
        const uint32_t t = ra + (p ? RSTRIDE : 0u);
        ra = (ra & ~(RSTRIDE * (RING_W - 1))) | (t & (RSTRIDE * (RING_W - 1)));
        x_out = x;
    }
};

// ---- variant 3: 64-bit conversions (one XU op each), magic constant folded into the FFMA, one unsigned test, {cdf,freq} 8-byte
//      entries, one-word look-ahead, ring cursor in the top bits of a register (wraps for free, address = one shift-and-add)
struct D3 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint2 *dtab2; uint32_t ringlane;      // ringlane: shared address of ring word 0 of this lane; stride 2048 B
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const float qf = fmaf(__ull2float_rz((uint64_t)ch << 32 | cl), rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl)), 12582911.5f);
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint2 e = dtab2[x];
        const uint64_t rp = (uint64_t)rl * e.x, fr = (uint64_t)rl * e.y;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * e.x;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * e.y;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));     // (k >> 28) * 2048 + lane base
        asm volatile("{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "selp.u32 %0, %6, %7, p;\n\t"       // rh = p ? fl : fh
            "selp.u32 %1, 0, %6, p;\n\t"        // rl = p ? 0 : fl
            "selp.u32 %2, %8, %9, p;\n\t"       // ch = p ? dl : dh
            "selp.u32 %3, %4, %8, p;\n\t"       // cl = p ? n0 : dl
            "@p ld.shared.u32 %4, [%10];\n\t"
            "@p add.u32 %5, %5, 0x10000000;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};

// ---- variant 4: the library's RcD3::step (floor inside the FMA: fma.rm into the 2^23 integer grid)
struct D4 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint2 *dtab2; uint32_t ringlane;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        float qf;
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint2 e = dtab2[x];
        const uint64_t rp = (uint64_t)rl * e.x, fr = (uint64_t)rl * e.y;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * e.x;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * e.y;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));
        asm volatile("{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "selp.u32 %0, %6, %7, p;\n\t"
            "selp.u32 %1, 0, %6, p;\n\t"
            "selp.u32 %2, %8, %9, p;\n\t"
            "selp.u32 %3, %4, %8, p;\n\t"
            "@p ld.shared.u32 %4, [%10];\n\t"
            "@p add.u32 %5, %5, 0x10000000;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};
// ---- variant 7: variant 4 with {cdf, freq} packed in ONE 32-bit table entry (LDS.32: half the shared-memory passes of LDS.64, two unpack instructions)
struct D7 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint32_t *dtab; const uint2 *dtab2; uint32_t ringlane;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        float qf;
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint32_t e = dtab[x], ex = e >> 16, ey = e & 0xffffu;
        const uint64_t rp = (uint64_t)rl * ex, fr = (uint64_t)rl * ey;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * ex;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * ey;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));
        asm volatile("{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "selp.u32 %0, %6, %7, p;\n\t"
            "selp.u32 %1, 0, %6, p;\n\t"
            "selp.u32 %2, %8, %9, p;\n\t"
            "selp.u32 %3, %4, %8, p;\n\t"
            "@p ld.shared.u32 %4, [%10];\n\t"
            "@p add.u32 %5, %5, 0x10000000;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};
// ---- variant 12: variant 7 with the 32-bit table replicated once per lane (entry of symbol x for lane l at word 32 x + l: always one conflict-free pass)
struct D12 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint32_t *dtab; const uint2 *dtab2; uint32_t ringlane;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        float qf;
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint32_t e = dtab[x * 32], ex = e >> 16, ey = e & 0xffffu;
        const uint64_t rp = (uint64_t)rl * ex, fr = (uint64_t)rl * ey;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * ex;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * ey;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));
        asm volatile("{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "selp.u32 %0, %6, %7, p;\n\t"
            "selp.u32 %1, 0, %6, p;\n\t"
            "selp.u32 %2, %8, %9, p;\n\t"
            "selp.u32 %3, %4, %8, p;\n\t"
            "@p ld.shared.u32 %4, [%10];\n\t"
            "@p add.u32 %5, %5, 0x10000000;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};
// ---- variant 5: variant 4 with the renormalisation selects as multiply-adds by a 0/1 word (FMA pipe instead of ALU pipe):
//      renorm <=> fh == 0, and then dh == 0 too in a valid step, so  rh = fh + pi*fl,  rl = np*fl,  ch = dh + pi*dl,  cl = np*dl + pi*n0
struct D5 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint2 *dtab2; uint32_t ringlane;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        float qf;
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint2 e = dtab2[x];
        const uint64_t rp = (uint64_t)rl * e.x, fr = (uint64_t)rl * e.y;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * e.x;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * e.y;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 np, pi, t;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "min.u32 np, %7, 1;\n\t"
            "mad.lo.u32 pi, np, -1, 1;\n\t"
            "mad.lo.u32 %0, %6, pi, %7;\n\t"       // rh = fh + pi*fl
            "mul.lo.u32 %1, %6, np;\n\t"           // rl = np*fl
            "mad.lo.u32 %2, %8, pi, %9;\n\t"       // ch = dh + pi*dl
            "mul.lo.u32 t, %8, np;\n\t"
            "mad.lo.u32 %3, %4, pi, t;\n\t"        // cl = np*dl + pi*n0
            "@p ld.shared.u32 %4, [%10];\n\t"
            "mad.lo.u32 %5, pi, 0x10000000, %5;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};
// ---- variant 6: variant 4 with only the two "free" selects moved (rh and ch: one IMAD each, no extra operand) and k by IMAD
struct D6 {
    uint32_t rl, rh, cl, ch, n0, k; bool bad;
    const uint8_t *lut; const uint2 *dtab2; uint32_t ringlane;
    __device__ __forceinline__ void step(uint32_t &x_out) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        float qf;
        asm("fma.rm.f32 %0, %1, %2, 0f4B400000;" : "=f"(qf) : "f"(__ull2float_rz((uint64_t)ch << 32 | cl)), "f"(rcp_approx(__ull2float_rn((uint64_t)rh << 32 | rl))));
        const uint32_t x = lut[__float_as_uint(qf) & 0x7fffu];
        const uint2 e = dtab2[x];
        const uint64_t rp = (uint64_t)rl * e.x, fr = (uint64_t)rl * e.y;
        const uint32_t pl = (uint32_t)rp, ph = (uint32_t)(rp >> 32) + rh * e.x;
        const uint32_t fl = (uint32_t)fr, fh = (uint32_t)(fr >> 32) + rh * e.y;
        uint32_t dl, dh;
        asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(dl), "=r"(dh) : "r"(cl), "r"(ch), "r"(pl), "r"(ph));
        bad |= ((uint64_t)dh << 32 | dl) >= ((uint64_t)fh << 32 | fl);
        uint32_t a;
        asm("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(a) : "r"(k), "r"(ringlane));
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 pi;\n\t"
            "setp.eq.u32 p, %7, 0;\n\t"
            "selp.u32 pi, 1, 0, p;\n\t"
            "mad.lo.u32 %0, %6, pi, %7;\n\t"       // rh = fh + pi*fl
            "selp.u32 %1, 0, %6, p;\n\t"           // rl = p ? 0 : fl
            "mad.lo.u32 %2, %8, pi, %9;\n\t"       // ch = dh + pi*dl
            "selp.u32 %3, %4, %8, p;\n\t"          // cl = p ? n0 : dl
            "@p ld.shared.u32 %4, [%10];\n\t"
            "mad.lo.u32 %5, pi, 0x10000000, %5;\n\t"
            "}" : "=r"(rh), "=r"(rl), "=r"(ch), "=r"(cl), "+r"(n0), "+r"(k) : "r"(fl), "r"(fh), "r"(dl), "r"(dh), "r"(a) : "memory");
        x_out = x;
    }
};

template <int V>
__global__ void __launch_bounds__(T, 2) k(const uint4 *__restrict__ in, const uint32_t *__restrict__ gdtab, const uint8_t *__restrict__ glut,
                                          uint2 *__restrict__ out, uint32_t *__restrict__ flags, int nblk) {
    extern __shared__ __align__(16) uint8_t dyn[];                   // [ring: RING_W words x 512 lanes, word-major | lut | dtab2 | dtab]
    uint32_t *ringbuf = (uint32_t *)dyn;
    uint8_t *lut = dyn + RING_W * 512 * 4;
    uint2 *dtab2 = (uint2 *)(lut + 32768);
    uint32_t *dtab = (uint32_t *)(dtab2 + 256);
    uint32_t *rep = dtab + 256;                                       // V == 12: 256 symbols x 32 lanes
    uint8_t *stage = (uint8_t *)(((uintptr_t)(dtab + 256) + 1023) & ~(uintptr_t)1023);   // V == 16: 2 KB per warp
    for (int i = threadIdx.x; i < 256; i += T) { dtab[i] = gdtab[i]; dtab2[i] = make_uint2(gdtab[i] >> 16, gdtab[i] & 0xffff); }
    for (int i = threadIdx.x; i < 32768 / 16; i += T) ((uint4 *)lut)[i] = ((const uint4 *)glut)[i];
    if (V == 12) for (int i = threadIdx.x; i < 256 * 32; i += T) rep[i] = gdtab[i >> 5];
    __syncthreads();
    const size_t gid = (size_t)blockIdx.x * T + threadIdx.x;
    const uint4 *ip = in + gid * (size_t)(nblk + 4);                // private word stream (4 words per block on average is plenty)
    uint32_t *ring = ringbuf + threadIdx.x;
    constexpr uint32_t RS = V == 0 ? T : 512;                       // lane stride in words (512: a power of two, see D1/D3)                       // lane stride in words
    auto ring_put = [&](uint32_t w, const uint4 &v) {
        ring[((w + 0) & (RING_W - 1)) * RS] = v.x; ring[((w + 1) & (RING_W - 1)) * RS] = v.y;
        ring[((w + 2) & (RING_W - 1)) * RS] = v.z; ring[((w + 3) & (RING_W - 1)) * RS] = v.w;
    };
    uint32_t fi = 0, qi = 0;
    for (int q = 0; q < 3; q++) { ring_put(fi, ip[qi++]); fi += 4; }
    uint32_t acc = 0;
    uint2 *op = out + gid * (size_t)nblk;
    if (V == 0) {
        D0 d; d.lut = lut; d.dtab = dtab; d.ring = ring;
        d.rl = d.rh = 0xffffffffu; d.ch = ring[0] >> 1; d.cl = ring[RS]; d.n0 = ring[2 * RS]; d.n1 = ring[3 * RS]; d.ci = 4; d.bad = 0;
#pragma unroll 1
        for (int b = 0; b < nblk; b++) {
            const bool need = fi - d.ci <= 10;
            uint4 t4 = make_uint4(0, 0, 0, 0);
            if (need) t4 = ip[qi];
            uint32_t a0 = 0, a1 = 0, x;
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a0 |= x << (8 * s); }
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a1 |= x << (8 * s); }
            if (need) { ring_put(fi, t4); fi += 4; qi++; }
            acc += d.bad; d.bad = 0;
            op[b] = make_uint2(a0, a1);
        }
    } else if (V >= 3) {
        typename std::conditional<V == 3, D3, typename std::conditional<V == 4, D4, typename std::conditional<V == 5, D5, typename std::conditional<V == 6, D6, typename std::conditional<V == 7 || V == 11, D7, typename std::conditional<V == 12, D12, D4>::type>::type>::type>::type>::type>::type d; d.lut = lut; d.dtab2 = dtab2; if constexpr (V == 7 || V == 11) d.dtab = dtab; if constexpr (V == 12) d.dtab = rep + (threadIdx.x & 31); d.ringlane = (uint32_t)__cvta_generic_to_shared(ring);
        d.rl = d.rh = 0xffffffffu; d.ch = ring[0] >> 1; d.cl = ring[RS]; d.n0 = ring[2 * RS]; d.bad = false; d.k = 3u << 28;
        uint32_t ci = 3, k_prev = d.k;
#pragma unroll 1
        for (int b = 0; b < nblk; b++) {
            const bool need = fi - ci <= 10;
            uint4 t4 = make_uint4(0, 0, 0, 0);
            if (need) t4 = (V == 8 || V == 10) ? in[(size_t)blockIdx.x * T * (nblk + 4) + (size_t)qi * T + threadIdx.x] : ip[qi];   // 8/10: lane-adjacent (coalesced) stream words
            uint32_t a0 = 0, a1 = 0, x;
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a0 |= x << (8 * s); }
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a1 |= x << (8 * s); }
            ci += (d.k - k_prev) >> 28; k_prev = d.k;
            if (need) { ring_put(fi, t4); fi += 4; qi++; }
            acc += d.bad ? 1u : 0u; d.bad = false;
            if (V >= 13) {                                                            // the library's geometry: lanes 2r / 2r+1 own bytes [0,8) / [8,16) of every 16-byte block of call r
                uint2 *cp = out + ((gid >> 1) * (size_t)nblk + b) * 2 + (gid & 1);
                const uint2 v = make_uint2(a0, a1);
                if (V == 13) *cp = v;
                else if (V == 14) __stcs(cp, v);
                else if (V == 15) asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" :: "l"(cp), "r"(v.x), "r"(v.y) : "memory");
                else if (V == 16) {                                                   // 8 blocks = 128 bytes per call in a swizzled warp tile, then 4 coalesced 128-bit stores per lane
                    const uint32_t lane = threadIdx.x & 31, row = lane >> 1;
                    uint8_t *tile = stage + (threadIdx.x >> 5) * 2048;
                    *(uint2 *)(tile + row * 128 + ((((uint32_t)b ^ row) & 7) << 4) + 8 * (lane & 1)) = v;
                    if ((b & 7) == 7) {
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t r2 = q * 4 + (lane >> 3), ch = lane & 7;
                            const uint4 w = *(const uint4 *)(tile + r2 * 128 + (((ch ^ r2) & 7) << 4));
                            const size_t call = (((size_t)blockIdx.x * T + (threadIdx.x & ~31u)) >> 1) + r2;
                            *(uint4 *)((uint8_t *)out + (call * (size_t)nblk + (b & ~7)) * 16 + ch * 16) = w;
                        }
                        __syncwarp();
                    }
                }
            }
            else if (V >= 9) out[((size_t)blockIdx.x * nblk + b) * T + threadIdx.x] = make_uint2(a0, a1);      // 9-12: lane-adjacent (coalesced) output
            else op[b] = make_uint2(a0, a1);
        }
    } else {
        typename std::conditional<V == 1, D1, D2>::type d;
        d.lut = lut; d.dtab2 = dtab2; d.dyn = dyn;
        const uint32_t rb = threadIdx.x * 4;
        d.rl = d.rh = 0xffffffffu; d.ch = ring[0] >> 1; d.cl = ring[RS]; d.n0 = ring[2 * RS]; d.bad = false;
        uint32_t ci0;
        if constexpr (V == 1) { d.n1 = ring[3 * RS]; d.ra = rb + 4 * RSTRIDE; ci0 = 4; } else { d.ra = rb + 3 * RSTRIDE; ci0 = 3; }
        uint32_t ci = ci0, ra_prev = d.ra;
#pragma unroll 1
        for (int b = 0; b < nblk; b++) {
            const bool need = fi - ci <= 10;
            uint4 t4 = make_uint4(0, 0, 0, 0);
            if (need) t4 = ip[qi];
            uint32_t a0 = 0, a1 = 0, x;
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a0 |= x << (8 * s); }
#pragma unroll
            for (int s = 0; s < 4; s++) { d.step(x); a1 |= x << (8 * s); }
            ci += ((d.ra - ra_prev) / RSTRIDE) & (RING_W - 1); ra_prev = d.ra;      // words consumed by this block
            if (need) { ring_put(fi, t4); fi += 4; qi++; }
            acc += d.bad ? 1u : 0u; d.bad = false;
            op[b] = make_uint2(a0, a1);
        }
    }
    flags[gid] = acc;
}

template <int V>
static void run(const char *name, const uint4 *d_in, const uint32_t *d_tab, const uint8_t *d_lut, uint2 *d_out, uint32_t *d_flags, int nblk, int ctas) {
    const int me = g_idx++;
    if (getenv("UB_ONLY") && atoi(getenv("UB_ONLY")) != me) return;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t sm = (size_t)RING_W * 512 * 4 + 32768 + 2048 + 1024 + (V == 12 ? 32768 : 0) + (V == 16 ? 1024 + (T / 32) * 2048 : 0);
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int i = 0; i < 3; i++) k<V><<<ctas, T, sm>>>(d_in, d_tab, d_lut, d_out, d_flags, nblk);
    cudaEventRecord(a);
    const int reps = getenv("UB_REPS") ? atoi(getenv("UB_REPS")) : 20;
    for (int i = 0; i < reps; i++) k<V><<<ctas, T, sm>>>(d_in, d_tab, d_lut, d_out, d_flags, nblk);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    cudaError_t e = cudaGetLastError();
    const double sym = (double)ctas * T * nblk * 8;
    int clk = 0, sms = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double cyc = ms * 1e-3 * clk * 1e3 * sms * 4 / (sym / 32);
    std::vector<uint32_t> fl((size_t)ctas * T); cudaMemcpy(fl.data(), d_flags, fl.size() * 4, cudaMemcpyDeviceToHost);
    double bad = 0; for (auto v : fl) bad += v;
    printf("%-34s %8.1f us  %6.2f Gsym/s  %6.1f cycles per warp-symbol and scheduler, flagged blocks %.3f%%  (%s)\n", name, ms * 1e3, sym / ms / 1e6, cyc,
           100.0 * bad / ((double)ctas * T * nblk), cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 110;
    const int ctas = argc > 2 ? atoi(argv[2]) : 296;
    const size_t lanes = (size_t)ctas * T;
    std::vector<uint32_t> tab(256); std::vector<uint8_t> lut(32768);
    { double s = 0; std::vector<double> p(256); for (int i = 0; i < 256; i++) { p[i] = pow(i + 1.0, -1.1); s += p[i]; }
      std::vector<uint32_t> f(256); uint32_t cum = 0; for (int i = 0; i < 256; i++) { f[i] = (uint32_t)(p[i] / s * 32768); if (!f[i]) f[i] = 1; cum += f[i]; }
      f[0] += 32768 - cum; cum = 0;
      for (int i = 0; i < 256; i++) { tab[i] = f[i] | cum << 16; for (uint32_t r = cum; r < cum + f[i]; r++) lut[r] = (uint8_t)i; cum += f[i]; } }
    std::vector<uint32_t> h(lanes * (size_t)(nblk + 4) * 4);
    { uint64_t s = 88172645463325252ull; for (auto &x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s >> 16); } }
    uint4 *d_in; uint32_t *d_tab, *d_flags; uint8_t *d_lut; uint2 *d_out;
    cudaMalloc(&d_in, h.size() * 4); cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_tab, 1024); cudaMemcpy(d_tab, tab.data(), 1024, cudaMemcpyHostToDevice);
    cudaMalloc(&d_lut, 32768); cudaMemcpy(d_lut, lut.data(), 32768, cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, lanes * (size_t)nblk * 8); cudaMalloc(&d_flags, lanes * 4);
    printf("lanes %zu, %d symbols per lane, %d CTAs x %d\n", lanes, nblk * 8, ctas, T);
    run<0>("d0 round-1 step", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<1>("d1 32-bit cvt, single test, 2-word la", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<3>("d3 64-bit cvt, folded magic, top-bit cursor", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<2>("d2 same, 1-word look-ahead", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<4>("d4 library step (fma.rm)", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<5>("d5 renorm selects as 0/1 IMADs", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<6>("d6 two selects + cursor as IMADs", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<7>("d7 library step, 32-bit table entries", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<8>("d8 = d4, coalesced stream loads", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<9>("d9 = d4, coalesced output stores", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<10>("d10 = d4, both coalesced", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<11>("d11 = d7 (32-bit entries), coalesced stores", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<12>("d12 = lane-replicated 32-bit table, coalesced stores", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<13>("d13 = d4, library output geometry (16 B per call and block)", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<14>("d14 = d13 with st.cs", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<15>("d15 = d13 with L1::no_allocate", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    run<16>("d16 = d13 staged: 128 B per call, coalesced STG.128", d_in, d_tab, d_lut, d_out, d_flags, nblk, ctas);
    return 0;
}
