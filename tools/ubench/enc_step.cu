// enc_step.cu -- micro-benchmark of range-coder ENCODE step formulations (tools only; not part of the library).
// Every variant runs the same harness: 296 CTAs x 384 lanes (2 CTAs per SM on a B200), each lane codes NSYM symbols of its own
// pseudo-random Zipf-ish byte stream through a 256-entry {cdf, freq} table in shared memory, 8 symbols per 16-byte load.
// Output: microseconds, and cycles per warp-symbol per scheduler (lower bound for the real kernel's main loop).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
constexpr int T = 384;
static int g_idx = 0;
constexpr int PROB_BITS = 15;

// ---- variant 0: the round-1 step (RcE32::encode of static_v2.cuh), stores to a global slot
struct V0 {
    uint32_t rl, rh, ll, lh, pend, carry, rare, pos; uint32_t *base;
    __device__ __forceinline__ void init(uint32_t *b, uint32_t) { rl = rh = 0xffffffffu; ll = lh = 0; pend = carry = rare = pos = 0; base = b; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        rl = __funnelshift_r(rl, rh, PROB_BITS); rh >>= PROB_BITS;
        const uint32_t tl = rl * c0, th = __umulhi(rl, c0) + rh * c0;
        uint32_t cy;
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;" : "+r"(ll), "+r"(lh), "=r"(cy) : "r"(tl), "r"(th));
        carry |= cy;
        const uint32_t nl = rl * f, nh = __umulhi(rl, f) + rh * f;
        const bool p = nh == 0;
        const uint32_t np = pend + carry;
        rare |= (p && np < carry) ? 1u : 0u;
        if (p) base[(int)pos - 1] = np;
        pend = p ? lh : pend; pos += p ? 1u : 0u; carry = p ? 0u : carry;
        lh = p ? ll : lh; ll = p ? 0u : ll;
        rh = p ? nl : nh; rl = p ? 0u : nl;
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + pos; }
    __device__ __forceinline__ void blockend(uint32_t) {}
};
// ---- variant 1: 96-bit low (carry chain runs straight into the pending word), words to a shared-memory ring
struct V1 {
    uint32_t rl, rh, ll, lh, pend, rare, sa;
    __device__ __forceinline__ void init(uint32_t *, uint32_t sa0) { rl = rh = 0xffffffffu; ll = lh = pend = rare = 0; sa = sa0; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, tl, th;\n\t"
            ".reg .u64 nr, tt;\n\t"
            "shf.r.wrap.b32 %0, %0, %1, 15;\n\t"
            "shr.u32 %1, %1, 15;\n\t"
            "mul.wide.u32 tt, %0, %7;\n\t"
            "mov.b64 {tl, th}, tt;\n\t"
            "mad.lo.u32 th, %1, %7, th;\n\t"
            "add.cc.u32 %2, %2, tl;\n\t"
            "addc.cc.u32 %3, %3, th;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mul.wide.u32 nr, %0, %8;\n\t"
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %8, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"
            "@p st.shared.u32 [%6], %4;\n\t"
            "@p add.u32 %6, %6, %9;\n\t"
            "@p mov.u32 %4, %3;\n\t"
            "@p mov.u32 %3, %2;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(pend), "+r"(rare), "+r"(sa)
            : "r"(c0), "r"(f), "n"(T * 4) : "memory");
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + sa; }
    __device__ __forceinline__ void blockend(uint32_t sa0) { sa = sa0; }
};
// ---- variant 2: 64-bit low with multiply-accumulate, carry recovered at renormalisation from "low < low at last renorm"
struct V2 {
    uint32_t rl, rh, ilh, pend, rare, sa; uint64_t low;
    __device__ __forceinline__ void init(uint32_t *, uint32_t sa0) { rl = rh = 0xffffffffu; low = 0; ilh = pend = rare = 0; sa = sa0; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        rl = __funnelshift_r(rl, rh, 15); rh >>= 15;
        low += (uint64_t)rl * c0;
        uint32_t ll = (uint32_t)low, lh = (uint32_t)(low >> 32) + rh * c0;
        const uint64_t nr = (uint64_t)rl * f;
        const uint32_t nl = (uint32_t)nr, nh = (uint32_t)(nr >> 32) + rh * f;
        const bool p = nh == 0;
        uint32_t np, t;
        asm("sub.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, 0;\n\taddc.u32 %2, %2, 0;" : "=r"(t), "=r"(np), "+r"(rare) : "r"(lh), "r"(ilh), "r"(pend));
        if (p) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa), "r"(np) : "memory");
        sa += p ? T * 4 : 0;
        pend = p ? lh : pend;
        ilh = p ? ll : ilh;
        lh = p ? ll : lh; ll = p ? 0u : ll;
        low = (uint64_t)lh << 32 | ll;
        rh = p ? nl : nh; rl = p ? 0u : nl;
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + sa; }
    __device__ __forceinline__ void blockend(uint32_t sa0) { sa = sa0; }
};
// ---- variant 3: variant 1 with the ilh carry (asm form)
struct V3 {
    uint32_t rl, rh, ll, lh, ilh, pend, rare, sa;
    __device__ __forceinline__ void init(uint32_t *, uint32_t sa0) { rl = rh = 0xffffffffu; ll = lh = ilh = pend = rare = 0; sa = sa0; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, np, t;\n\t"
            ".reg .u64 lo, nr;\n\t"
            "shf.r.wrap.b32 %0, %0, %1, 15;\n\t"
            "shr.u32 %1, %1, 15;\n\t"
            "mov.b64 lo, {%2, %3};\n\t"
            "mad.wide.u32 lo, %0, %8, lo;\n\t"
            "mov.b64 {%2, %3}, lo;\n\t"
            "mad.lo.u32 %3, %1, %8, %3;\n\t"
            "mul.wide.u32 nr, %0, %9;\n\t"
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %9, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"
            "sub.cc.u32 t, %3, %4;\n\t"
            "addc.cc.u32 np, %5, 0;\n\t"
            "addc.u32 %6, %6, 0;\n\t"
            "@p st.shared.u32 [%7], np;\n\t"
            "@p add.u32 %7, %7, %10;\n\t"
            "@p mov.u32 %5, %3;\n\t"
            "@p mov.u32 %4, %2;\n\t"
            "@p mov.u32 %3, %2;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(ilh), "+r"(pend), "+r"(rare), "+r"(sa)
            : "r"(c0), "r"(f), "n"(T * 4) : "memory");
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + sa; }
    __device__ __forceinline__ void blockend(uint32_t sa0) { sa = sa0; }
};

// ---- variant 4: variant 1's arithmetic, words stored straight to the global slot (no staging ring)
struct V4 {
    uint32_t rl, rh, ll, lh, pend, rare, pos; uint32_t *base;
    __device__ __forceinline__ void init(uint32_t *b, uint32_t) { rl = rh = 0xffffffffu; ll = lh = pend = rare = 0; pos = 0; base = b; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        uint32_t p;
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, tl, th;\n\t"
            ".reg .u64 nr, tt, ad;\n\t"
            "shf.r.wrap.b32 %0, %0, %1, 15;\n\t"
            "shr.u32 %1, %1, 15;\n\t"
            "mul.wide.u32 tt, %0, %8;\n\t"
            "mov.b64 {tl, th}, tt;\n\t"
            "mad.lo.u32 th, %1, %8, th;\n\t"
            "add.cc.u32 %2, %2, tl;\n\t"
            "addc.cc.u32 %3, %3, th;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mul.wide.u32 nr, %0, %9;\n\t"
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %9, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"
            "mad.wide.u32 ad, %6, 4, %10;\n\t"
            "@p st.global.u32 [ad], %4;\n\t"
            "@p add.u32 %6, %6, 1;\n\t"
            "@p mov.u32 %4, %3;\n\t"
            "@p mov.u32 %3, %2;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "selp.u32 %7, 1, 0, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(pend), "+r"(rare), "+r"(pos), "=r"(p)
            : "r"(c0), "r"(f), "l"(base) : "memory");
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + pos; }
    __device__ __forceinline__ void blockend(uint32_t) {}
};

// ---- variant 5: variant 1 with shifts and state moves forced onto the fma pipe (multiplier / one in registers ptxas cannot fold)
__device__ uint32_t g_M, g_ONE;
struct V5 {
    uint32_t rl, rh, ll, lh, pend, rare, sa, M, ONE;
    __device__ __forceinline__ void init(uint32_t *, uint32_t sa0) { rl = rh = 0xffffffffu; ll = lh = pend = rare = 0; sa = sa0; M = g_M; ONE = g_ONE; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, tl, th, t, z;\n\t"
            ".reg .u64 nr, tt, rr;\n\t"
            "mul.hi.u32 t, %0, %10;\n\t"
            "mov.u32 z, 0;\n\t"
            "mov.b64 rr, {t, z};\n\t"
            "mad.wide.u32 rr, %1, %10, rr;\n\t"
            "mov.b64 {%0, %1}, rr;\n\t"
            "mul.wide.u32 tt, %0, %7;\n\t"
            "mov.b64 {tl, th}, tt;\n\t"
            "mad.lo.u32 th, %1, %7, th;\n\t"
            "add.cc.u32 %2, %2, tl;\n\t"
            "addc.cc.u32 %3, %3, th;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mul.wide.u32 nr, %0, %8;\n\t"
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %8, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"
            "@p st.shared.u32 [%6], %4;\n\t"
            "@p add.u32 %6, %6, %9;\n\t"
            "@p mul.lo.u32 %4, %3, %11;\n\t"
            "@p mul.lo.u32 %3, %2, %11;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(pend), "+r"(rare), "+r"(sa)
            : "r"(c0), "r"(f), "n"(T * 4), "r"(M), "r"(ONE) : "memory");
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + sa; }
    __device__ __forceinline__ void blockend(uint32_t sa0) { sa = sa0; }
};
// ---- variant 6: variant 5 but only the shifts on the fma pipe
struct V6 {
    uint32_t rl, rh, ll, lh, pend, rare, sa, M;
    __device__ __forceinline__ void init(uint32_t *, uint32_t sa0) { rl = rh = 0xffffffffu; ll = lh = pend = rare = 0; sa = sa0; M = g_M; }
    __device__ __forceinline__ void step(uint32_t c0, uint32_t f) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 nl, nh, tl, th, t, z;\n\t"
            ".reg .u64 nr, tt, rr;\n\t"
            "mul.hi.u32 t, %0, %10;\n\t"
            "mov.u32 z, 0;\n\t"
            "mov.b64 rr, {t, z};\n\t"
            "mad.wide.u32 rr, %1, %10, rr;\n\t"
            "mov.b64 {%0, %1}, rr;\n\t"
            "mul.wide.u32 tt, %0, %7;\n\t"
            "mov.b64 {tl, th}, tt;\n\t"
            "mad.lo.u32 th, %1, %7, th;\n\t"
            "add.cc.u32 %2, %2, tl;\n\t"
            "addc.cc.u32 %3, %3, th;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mul.wide.u32 nr, %0, %8;\n\t"
            "mov.b64 {nl, nh}, nr;\n\t"
            "mad.lo.u32 nh, %1, %8, nh;\n\t"
            "setp.eq.u32 p, nh, 0;\n\t"
            "@p st.shared.u32 [%6], %4;\n\t"
            "@p add.u32 %6, %6, %9;\n\t"
            "@p mov.u32 %4, %3;\n\t"
            "@p mov.u32 %3, %2;\n\t"
            "@p mov.u32 %2, 0;\n\t"
            "selp.u32 %1, nl, nh, p;\n\t"
            "selp.u32 %0, 0, nl, p;\n\t"
            "}"
            : "+r"(rl), "+r"(rh), "+r"(ll), "+r"(lh), "+r"(pend), "+r"(rare), "+r"(sa)
            : "r"(c0), "r"(f), "n"(T * 4), "r"(M) : "memory");
    }
    __device__ __forceinline__ uint32_t fin() { return rare + pend + sa; }
    __device__ __forceinline__ void blockend(uint32_t sa0) { sa = sa0; }
};

template <class V, int CTAS_PER_SM>
__global__ void __launch_bounds__(T, CTAS_PER_SM) k(const uint4 *__restrict__ in, const uint2 *__restrict__ gtab, uint32_t *__restrict__ slots, uint32_t *__restrict__ out, int nblk) {
    __shared__ uint2 tab[256];
    __shared__ uint32_t ring[16 * T];
    if (threadIdx.x < 256) tab[threadIdx.x] = gtab[threadIdx.x];
    __syncthreads();
    const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(ring + threadIdx.x);
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(tab);
    const size_t gid = (size_t)blockIdx.x * T + threadIdx.x;
    V e; e.init(slots + gid * (size_t)(nblk * 8 + 16) + 8, sa0);
    const unsigned c = threadIdx.x & 1;
    const uint4 *ip = in + (gid >> 1) * (size_t)nblk;            // the two lanes of a call read the same 16 bytes
    uint4 cur = ip[0];
#pragma unroll 1
    for (int b = 0; b < nblk; b++) {
        const uint4 nxt = ip[b + 1 < nblk ? b + 1 : b];
        const uint32_t w[4] = { cur.x >> (8 * c), cur.y >> (8 * c), cur.z >> (8 * c), cur.w >> (8 * c) };
        uint32_t tx[8], ty[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t a = tb + (((q & 1) ? (w[q >> 1] >> 13) : (w[q >> 1] << 3)) & 0x7f8);
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tx[q]), "=r"(ty[q]) : "r"(a));
        }
#pragma unroll
        for (int q = 0; q < 8; q++) e.step(tx[q], ty[q]);
        e.blockend(sa0);
        cur = nxt;
    }
    out[gid] = e.fin() + ring[threadIdx.x];
}

template <class V, int CPS>
static void run(const char *name, const uint4 *d_in, const uint2 *d_tab, uint32_t *d_slots, uint32_t *d_out, int nblk, int ctas) {
    const int me = g_idx++;
    if (getenv("UB_ONLY") && atoi(getenv("UB_ONLY")) != me) return;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) k<V, CPS><<<ctas, T>>>(d_in, d_tab, d_slots, d_out, nblk);
    cudaEventRecord(a);
    const int reps = getenv("UB_REPS") ? atoi(getenv("UB_REPS")) : 20;
    for (int i = 0; i < reps; i++) k<V, CPS><<<ctas, T>>>(d_in, d_tab, d_slots, d_out, nblk);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    cudaError_t e = cudaGetLastError();
    const double sym = (double)ctas * T * nblk * 8;
    int clk = 0, sms = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double cyc = ms * 1e-3 * clk * 1e3 * sms * 4 / (sym / 32);
    printf("%-28s %8.1f us  %6.2f Gsym/s  %6.1f cycles per warp-symbol and scheduler  (%s)\n", name, ms * 1e3, sym / ms / 1e6, cyc, cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 110;               // 110 x 8 = 880 symbols per lane (1760-byte calls)
    const int ctas = argc > 2 ? atoi(argv[2]) : 296;
    const size_t lanes = (size_t)ctas * T;
    std::vector<uint2> tab(256);
    { double s = 0; std::vector<double> p(256); for (int i = 0; i < 256; i++) { p[i] = pow(i + 1.0, -1.1); s += p[i]; }
      uint32_t cum = 0; for (int i = 0; i < 256; i++) { uint32_t f = (uint32_t)(p[i] / s * 32768); if (!f) f = 1; tab[i] = make_uint2(cum, f); cum += f; }
      tab[0].y += 32768 - cum; cum = 0; for (int i = 0; i < 256; i++) { tab[i].x = cum; cum += tab[i].y; } }
    std::vector<uint8_t> h(lanes / 2 * nblk * 16);
    { uint64_t s = 88172645463325252ull; std::vector<double> cdf(256); double t = 0, tot = 0; for (int i = 0; i < 256; i++) tot += pow(i + 1.0, -1.1);
      for (int i = 0; i < 256; i++) { t += pow(i + 1.0, -1.1) / tot; cdf[i] = t; }
      for (auto &x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; double u = (s >> 11) * (1.0 / 9007199254740992.0); int lo = 0; while (lo < 255 && cdf[lo] < u) lo++; x = (uint8_t)lo; } }
    uint4 *d_in; uint2 *d_tab; uint32_t *d_slots, *d_out;
    cudaMalloc(&d_in, h.size()); cudaMemcpy(d_in, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_tab, 2048); cudaMemcpy(d_tab, tab.data(), 2048, cudaMemcpyHostToDevice);
    cudaMalloc(&d_slots, lanes * (size_t)(nblk * 8 + 16) * 4); cudaMalloc(&d_out, lanes * 4);
    { uint32_t m = 0x20000, one = 1; cudaMemcpyToSymbol(g_M, &m, 4); cudaMemcpyToSymbol(g_ONE, &one, 4); }
    printf("lanes %zu, %d symbols per lane, %d CTAs x %d\n", lanes, nblk * 8, ctas, T);
    run<V0, 2>("v0 round-1 step, global slot", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V1, 2>("v1 96-bit low, smem ring", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V2, 2>("v2 mad64 + ilh (C)", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V3, 2>("v3 mad64 + ilh (asm)", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V5, 2>("v5 v1 + shifts/moves on fma pipe", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V6, 2>("v6 v1 + shifts on fma pipe", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V4, 2>("v4 96-bit low, global slot", d_in, d_tab, d_slots, d_out, nblk, ctas);
    run<V4, 3>("v4, 3 CTAs per SM", d_in, d_tab, d_slots, d_out, nblk, ctas * 3 / 2);
    run<V1, 3>("v1, 3 CTAs per SM", d_in, d_tab, d_slots, d_out, nblk, ctas * 3 / 2);
    run<V2, 3>("v2, 3 CTAs per SM", d_in, d_tab, d_slots, d_out, nblk, ctas * 3 / 2);
    return 0;
}
