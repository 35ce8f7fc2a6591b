// trc_b200.cu -- C ABI (include/trc_b200.h) over the sm_100a kernels.  No CPU code path exists: every entry
// point launches kernels; failures surface as TRC_E_CUDA (batch layer) or a die()-style abort (drop-in layer).
#include "../../include/trc_b200.h"
#include "trc_common.cuh"
#include "rans_static.cuh"
#include "rc_static.cuh"
#include "adaptive.cuh"
#include "static_v2.cuh"
#include "rcs2_v3.cuh"
#include "adaptive_coop.cuh"
#include "adaptive_v3.cuh"
#include "vnibble.cuh"
#include "vlc.cuh"
#include "rans_wide.cuh"
#include "pack.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include <chrono>

using namespace trc;

static thread_local char g_err[256] = "";
static int g_dev = 0;
static const int g_force_redo = getenv("TRC_FORCE_REDO") ? 1 : 0;   // test hook: exercise the rare walk-back redo paths
static const int g_fused = getenv("TRC_FUSED") ? atoi(getenv("TRC_FUSED")) : 1;             // 0: coder, scan and pack as separate kernels (A/B runs)
static std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library (bench.py reports the delta)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return TRC_E_CUDA; } } while (0)
#define CK_LAUNCH() do { g_launches++; CK(cudaPeekAtLastError()); } while (0)

// Optional per-kernel timing of the batch calls (bench.py's roofline leg): when enabled, an event is recorded
// on the caller's stream before/after every kernel of trc_enc_batch_dev / trc_dec_batch_dev; trc_profile_read
// synchronises on them and returns the milliseconds of the most recent call.  Off by default (zero overhead).
// (state is per host thread: the events belong to the device that was current when the thread first profiled)
static std::atomic<int> g_prof{0};
static thread_local int g_prof_n = 0;
static thread_local cudaEvent_t g_pev[8];
static thread_local bool g_pev_init = false;
static void prof_mark(cudaStream_t st) {
    if (!g_prof.load(std::memory_order_relaxed)) return;
    if (!g_pev_init) { for (auto &e : g_pev) cudaEventCreate(&e); g_pev_init = true; }
    if (g_prof_n < 8) cudaEventRecord(g_pev[g_prof_n++], st);
}

// ---- per-device one-time state (several devices may be driven from one process) -----------------------------------
constexpr int MAX_DEV = 64;
static int cur_dev() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= MAX_DEV ? 0 : d; }
static int sm_count() {
    static int n_sm[MAX_DEV];
    const int d = cur_dev();
    if (!n_sm[d]) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); n_sm[d] = n > 0 ? n : 148; }
    return n_sm[d];
}
// TRC_RCS2 encoder input staging: 1 = TMA 2-D tiles (default), 0 = per-lane 128-bit loads (A/B runs)
static const int g_enc_tma = getenv("TRC_ENC_TMA") ? atoi(getenv("TRC_ENC_TMA")) : 1;
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encode() {
    static tmap_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (tmap_encode_fn)p;
    }
    return fn;
}

static int dev_attrs();
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

struct Plan {
    Geom g; int codec;
    size_t slot_stride, rec_stride, o1_threads;
    size_t off_meta, off_calls, off_slots, off_recs, off_o1, off_tabs, off_lb, total;
};

// tables needed for n calls when `cpc` consecutive calls share one (0 = one table for everything)
static inline size_t n_tables(size_t n_calls, size_t cpc) { return cpc ? (n_calls + cpc - 1) / cpc : 1; }
// the throughput kernels (static_v2.cuh) need 16-byte aligned calls and table groups that do not split a CTA
// lane-per-coder launch shape: (calls per CTA, CTAs).  Batches that fit one wave get one equally loaded CTA per SM.
static void lpc_shape(size_t n_calls, size_t cpc, unsigned &calls_per_cta, unsigned &ctas, bool two_per_sm = false, int n_sm_in = 0) {
    const int n_sm = n_sm_in > 0 ? n_sm_in : sm_count();
    static const size_t g_d3 = getenv("TRC_D3_CALLS") ? (size_t)atoi(getenv("TRC_D3_CALLS")) : LPC_MAX_NT / 2;      // A/B runs
    calls_per_cta = LPC_NT / 2;
    const size_t per_sm = (n_calls + n_sm - 1) / n_sm;
    if (two_per_sm && g_d3 >= 16 && g_d3 <= (size_t)LPC_MAX_NT / 2 && cpc == 0 && per_sm > 3 * (size_t)(LPC_MAX_NT / 2)) {
        // k_rcs2_dec3, more than one wave, one table: full waves of two equal CTAs per SM (1 GiB at 1760-byte chunks decodes at 636 GB/s
        // against 433 with 64-call CTAs, whose 34 KB tables keep all but three of them off an SM)
        const size_t slots = 2 * (size_t)n_sm, waves = (n_calls + slots * g_d3 - 1) / (slots * g_d3);
        calls_per_cta = (unsigned)((n_calls + slots * waves - 1) / (slots * waves));
        ctas = (unsigned)((n_calls + calls_per_cta - 1) / calls_per_cta);
        return;
    }
    if (two_per_sm && cpc != 0 && per_sm > 3 * (size_t)(LPC_MAX_NT / 2)) {      // a table per group of calls, more than one wave: the biggest CTA that divides a group
        calls_per_cta = cpc % 256 == 0 ? 256 : 128;                               // (v2_ok: cpc is a multiple of 128)
        ctas = (unsigned)((n_calls + calls_per_cta - 1) / calls_per_cta);
        return;
    }
    // one wave of equally loaded CTAs: k CTAs per SM (k <= 3 keeps registers and the 33 KB tables of each CTA resident)
    const size_t k = (per_sm + LPC_MAX_NT / 2 - 1) / (LPC_MAX_NT / 2);
    if (cpc == 0 && per_sm > LPC_NT / 2 && k >= 1 && k <= 3) calls_per_cta = (unsigned)((n_calls + (size_t)n_sm * k - 1) / ((size_t)n_sm * k));
    ctas = (unsigned)((n_calls + calls_per_cta - 1) / calls_per_cta);
}
// k_rcs2_enc3.  A batch of at most 512 calls per SM runs as ONE wave of one CTA per SM (its warps pace each other through shared
// memory; measured against two half-size CTAs per SM: 601 vs 576 GB/s at 384 calls per SM, 583 vs 558 at 509).  A bigger batch
// runs as waves of TWO CTAs per SM of at most 256 calls (16 warps: four per scheduler) each, so that one CTA of an SM codes while
// the other one is in its prologue / layout epilogue: 1 GiB at 1760-byte chunks encodes at 664-669 GB/s against 635 with 384-call
// CTAs and 582 with 512-call CTAs.  Up to four waves the CTAs are sized equal so that every wave is full (150 MB: 510 vs 481 GB/s).
static void e3_shape(size_t n_calls, unsigned &calls_per_cta, unsigned &ctas, int n_sm_in = 0) {
    static const size_t g_one = getenv("TRC_E3_ONE") ? (size_t)atoi(getenv("TRC_E3_ONE")) : E3_MAX_NT / 2;       // A/B runs
    static const size_t g_cap = getenv("TRC_E3_CALLS") ? (size_t)atoi(getenv("TRC_E3_CALLS")) : 256;
    const size_t n_sm = (size_t)(n_sm_in > 0 ? n_sm_in : sm_count());
    size_t per = (n_calls + n_sm - 1) / n_sm;
    if (per > (g_one < (size_t)E3_MAX_NT / 2 ? g_one : (size_t)E3_MAX_NT / 2)) {
        const size_t cap = g_cap < 16 ? 16 : g_cap > E3_MAX_NT / 2 ? E3_MAX_NT / 2 : g_cap, slots = 2 * n_sm;
        const size_t waves = (n_calls + slots * cap - 1) / (slots * cap);
        per = cap;
        if (waves <= 4) { per = (n_calls + slots * waves - 1) / (slots * waves); per = (per + 15) & ~(size_t)15; }   // whole warps
    }
    if (per < 16) per = 16;
    calls_per_cta = (unsigned)per;
    ctas = (unsigned)((n_calls + per - 1) / per);
}
// lane-per-call v2 kernels: threads per CTA (one call per thread), same balancing idea
static unsigned v2_shape(size_t n_calls, size_t cpc) {
    unsigned cpcta, ctas; lpc_shape(n_calls, cpc, cpcta, ctas);           // calls per CTA if the batch is about one wave
    if (cpc == 0 && cpcta > (unsigned)V2_NT / 2) { unsigned nt = (cpcta + 31) & ~31u; if (nt > (unsigned)V2_NT && nt <= (unsigned)LPC_MAX_NT) return nt; }
    return V2_NT;
}
static inline bool v2_ok(const void *buf, size_t chunk_len, size_t cpc) {
    return ((uintptr_t)buf & 15) == 0 && (chunk_len & 15) == 0 && (cpc == 0 || cpc % V2_NT == 0);
}

static int make_plan(int codec, size_t total_len, size_t chunk_len, Plan &p) {
    if (codec < 0 || codec >= NCODECS || total_len == 0 || chunk_len == 0) return TRC_E_ARG;
    if ((chunk_len < total_len ? chunk_len : total_len) >= (1ull << 31)) return TRC_E_ARG;   // 32-bit lengths per call
    p.codec = codec;
    p.g = make_geom(codec, total_len, chunk_len);
    const bool blocked = codec_blocked(codec);
    p.slot_stride = (blocked && p.g.upc > 1) ? al16(4 * (size_t)p.g.unit_max + 64) : al16(p.g.unit_max) + (codec == ANSW ? 448 : 192);
    if (codec_vlc(codec)) {                                      // integer codecs: whole elements per call
        const size_t esz = vlc_param(codec).w32 ? 4 : 2;
        if (total_len % esz || (chunk_len < total_len && chunk_len % esz)) return TRC_E_ARG;
        if (codec_vlc_ans(codec)) p.slot_stride = vlc_lifo_off(p.g.unit_max) + al16(p.g.unit_max) + 192;   // out image + aligned LIFO scratch
    }
    if (codec == ANSW && (chunk_len & 3)) return TRC_E_ARG;                                   // our own format: calls start 4-byte aligned
    p.rec_stride = blocked ? (((size_t)2 * p.g.unit_max + 3) & ~(size_t)3) + 16 : 0;
    if (codec_vlc_ans(codec)) {                                  // at most two records per element, one block (4 Mi elements) at a time
        size_t el = p.g.unit_max / (vlc_param(codec).w32 ? 4 : 2) + 1;
        if (el > ANS_BLOCK) el = ANS_BLOCK;
        p.rec_stride = ((2 * el + 3) & ~(size_t)3) + 16;
    }
    p.o1_threads = 0;
    // (order-1 tables live in shared memory since the warp-cooperative kernels; no global table scratch)
    size_t o = 0;
    p.off_meta = o;  o += al256(p.g.n_units * sizeof(UnitMeta));
    p.off_calls = o; o += al256(p.g.n_calls * sizeof(CallInfo));
    p.off_slots = o; o += al256(p.g.n_units * p.slot_stride);
    p.off_recs = o;  o += al256(p.g.n_units * p.rec_stride * 4);
    p.off_o1 = o;    o += al256(p.o1_threads * O1_TAB_WORDS * 4);
    // room for one TableSet per V2_NT calls is the worst case the v2 path accepts (cpc >= V2_NT)
    p.off_tabs = o;  o += codec_static(codec) ? al256(((p.g.n_calls + V2_NT - 1) / V2_NT) * sizeof(TableSet)) : 0;
    p.off_lb = o;    o += codec == RCS2 ? al256((p.g.n_calls / 16 + 3) * 8) : 0;   // look-back words of the fused encoder + tile counter
    p.total = o;
    return TRC_OK;
}

// kernels that need more than 48 KB of dynamic shared memory: the opt-in is per device
static int dev_attrs() {
    static bool done[MAX_DEV];
    const int d = cur_dev();
    if (done[d]) return TRC_OK;
    CK(cudaFuncSetAttribute(k_rcs2_enc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e3_smem_bytes(E3_MAX_NT, true)));
    CK(cudaFuncSetAttribute(k_rcs2_enc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e3_smem_bytes(E3_MAX_NT, false)));
    CK(cudaFuncSetAttribute(k_rcs2_dec3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D3_SMEM));
    // The shared-memory / L1 split is an SM-wide setting: a resident CTA of a kernel that asked for a different split keeps the
    // SM from being reconfigured, and the coder CTAs queued behind it wait.  The small helper kernels that run NEXT TO the coders
    // (peer push, flag wait, acknowledgement) use no shared memory and would pull the split towards L1: they ask for the coders'
    // split instead (k_rcs2_enc3: one 142 KB CTA, k_rcs2_dec3: two 68 KB CTAs per SM -> the 164 KB setting = 72 % of 228 KB).
    const int carve = getenv("TRC_PUSH_CARVEOUT") ? atoi(getenv("TRC_PUSH_CARVEOUT")) : 72;
    if (carve >= 0) {
        CK(cudaFuncSetAttribute(k_push, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CK(cudaFuncSetAttribute(k_wait_flags, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CK(cudaFuncSetAttribute(k_ack, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    CK(cudaFuncSetAttribute(k_rcs2_dec_lpc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RING_W * LPC_MAX_NT * sizeof(uint32_t))));
    CK(cudaFuncSetAttribute(k_rans_static_dec_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RING_W * LPC_MAX_NT * sizeof(uint32_t))));
    CK(cudaFuncSetAttribute(k_ans_model3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m3_warp_bytes<true>()));
    CK(cudaFuncSetAttribute(k_ans_code3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3_LUT_BYTES));
    CK(cudaFuncSetAttribute(k_ans_dec3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d3_smem_bytes<true>()));
    CK(cudaFuncSetAttribute(k_ans1_dec_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G1_SMEM));
    done[d] = true;
    return TRC_OK;
}

// decode-side tables come from the stream-ordered allocator; keep its pool from trimming back to the OS at every
// synchronisation (the default release threshold of 0 makes each call pay a fresh cuMemMap)
static int pool_keep() {
    static bool done[MAX_DEV];
    const int d = cur_dev();
    if (done[d]) return TRC_OK;
    cudaMemPool_t pool; CK(cudaDeviceGetDefaultMemPool(&pool, d));
    unsigned long long keep = ~0ull; CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[d] = true;
    return TRC_OK;
}

// adaptive byte rANS encoder, third generation: model pass (one warp per unit) then coding pass (one lane per state)
static int launch_ans_enc3(bool o1, const unsigned char *d_in, const Geom &g, uint8_t *slots, size_t slot_stride, uint32_t *recs,
                           size_t rec_stride, UnitMeta *meta, cudaStream_t st) {
    static bool s_rcp[MAX_DEV];
    int rc0 = dev_attrs(); if (rc0) return rc0;
    const int dev = cur_dev();
    if (!s_rcp[dev]) {                                          // once per process and device: the reciprocal table
        k_build_rcp<<<PROB_TOTAL / 256, 256, 0, st>>>();
        CK_LAUNCH();
        s_rcp[dev] = true;
    }
    if (o1) k_ans_model3<true><<<(unsigned)(g.n_units < (size_t)sm_count() ? g.n_units : (size_t)sm_count()), 32, m3_warp_bytes<true>(), st>>>(d_in, g, recs, rec_stride);
    else    k_ans_model3<false><<<(unsigned)((g.n_units + M3_WPB - 1) / M3_WPB), M3_WPB * 32, M3_WPB * m3_warp_bytes<false>(), st>>>(d_in, g, recs, rec_stride);
    CK_LAUNCH();
    const size_t groups = (g.n_units + 7) / 8;
    k_ans_code3<<<(unsigned)((groups + C3_WPB - 1) / C3_WPB), C3_WPB * 32, C3_LUT_BYTES, st>>>(g, recs, rec_stride, slots, slot_stride, meta);
    return TRC_OK;
}

// Prebuilt coding tables of the static codecs (include/trc_b200.h trc_tables_*): everything the kernels derive from a cdf
// table -- built once per table set, like the harness computes cdfini once per bench() outside its timing (turborc.c:432).
struct trc_tables {
    int dev; unsigned cdfnum; size_t n;
    TableSet *ts;        // rANS / RC encoder entries, decoder entries and slot->symbol LUT (static_v2.cuh, rans_wide.cuh)
    EncTab2 *e2;         // TRC_RCS2 encoder (rcs2_v3.cuh)
    DecTab2 *d2;         // TRC_RCS2 decoder (rcs2_v3.cuh)
};

static int enc_batch_impl(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len, const cdf_t *d_cdf, unsigned cdfnum,
                          size_t chunks_per_cdf, const trc_tables *pre, unsigned char *d_out, uint64_t *d_out_off, void *d_scratch,
                          size_t scratch_bytes, void *cuda_stream);
static int dec_batch_impl(int codec, const unsigned char *d_in, const uint64_t *d_in_off, unsigned char *d_out, size_t total_len,
                          size_t chunk_len, const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf, const trc_tables *pre,
                          unsigned flags, void *cuda_stream);

extern "C" {

int trc_tables_create_dev(const cdf_t *d_cdf, unsigned cdfnum, size_t n_tabs, void *cuda_stream, trc_tables **out) {
    if (!d_cdf || !out || cdfnum == 0 || cdfnum > 256 || n_tabs == 0) return TRC_E_ARG;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    trc_tables *t = new (std::nothrow) trc_tables();
    if (!t) return TRC_E_NOMEM;
    t->dev = cur_dev(); t->cdfnum = cdfnum; t->n = n_tabs; t->ts = nullptr; t->e2 = nullptr; t->d2 = nullptr;
    if (cudaMalloc((void **)&t->ts, n_tabs * sizeof(TableSet)) != cudaSuccess || cudaMalloc((void **)&t->e2, n_tabs * sizeof(EncTab2)) != cudaSuccess ||
        cudaMalloc((void **)&t->d2, n_tabs * sizeof(DecTab2)) != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "trc_tables_create_dev: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(t->ts); cudaFree(t->e2); cudaFree(t->d2); delete t; return TRC_E_NOMEM;
    }
    k_build_tables<<<dim3((unsigned)n_tabs, 1 + LUT_PARTS), 1024, 0, st>>>(d_cdf, cdfnum, t->ts, 1);
    k_build_enctab2<<<(unsigned)n_tabs, 256, 0, st>>>(d_cdf, cdfnum, t->e2);
    k_build_dectab2<<<dim3((unsigned)n_tabs, 1 + LUT_PARTS), 1024, 0, st>>>(d_cdf, cdfnum, t->d2);
    g_launches += 3;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "trc_tables_create_dev: %s", cudaGetErrorString(e)); cudaFree(t->ts); cudaFree(t->e2); cudaFree(t->d2); delete t; return TRC_E_CUDA; }
    *out = t;
    return TRC_OK;
}
void trc_tables_destroy(trc_tables *t) {
    if (!t) return;
    int prev = 0; cudaGetDevice(&prev);
    if (prev != t->dev) cudaSetDevice(t->dev);
    cudaFree(t->ts); cudaFree(t->e2); cudaFree(t->d2);
    if (prev != t->dev) cudaSetDevice(prev);
    delete t;
}

const char *trc_version(void) { return "trc_b200 0.2 (sm_100a)"; }
const char *trc_last_error(void) { return g_err; }
int trc_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int trc_set_device(int dev) { if (dev < 0 || dev >= MAX_DEV) return TRC_E_ARG; CK(cudaSetDevice(dev)); g_dev = dev; return TRC_OK; }
unsigned long long trc_launch_count(void) { return g_launches.load(); }
void trc_profile_enable(int on) { g_prof = on; g_prof_n = 0; }
int trc_profile_read(float *ms, int cap) {       // -> number of kernel intervals written (kernels of the last call)
    int n = g_prof_n - 1, k = 0;
    if (n <= 0) return 0;
    cudaEventSynchronize(g_pev[g_prof_n - 1]);
    for (; k < n && k < cap; k++) cudaEventElapsedTime(&ms[k], g_pev[k], g_pev[k + 1]);
    return k;
}

// Host-only check of the arithmetic the kernels rely on: for every frequency f in [1, 2^15] the table entry must
// reproduce floor(s / f) for s at all multiples-of-f boundaries reachable after renormalisation (s < f << 16).
int trc_selftest_host(void) {
    int bad = 0;
    for (uint32_t f = 1; f <= PROB_TOTAL; f++) {
        uint4 e = rans_enc_entry(123, f);
        const uint32_t smax = (uint32_t)(((uint64_t)f << 16) - 1);
        auto chk = [&](uint32_t s) {
            uint32_t q = (uint32_t)(((uint64_t)s * e.x) >> 32) >> (e.z >> 16);
            uint32_t got = s + e.w + q * (e.z & 0xffffu);
            uint32_t want = (s / f << PROB_BITS) + s % f + 123;
            if (got != want) bad++;
        };
        for (uint32_t k = 1; k <= 65535; k += (f > 64 ? 257 : 1)) {
            uint64_t b = (uint64_t)k * f;
            if (b - 1 <= smax && b >= 2) chk((uint32_t)(b - 1));   // s >= 1 always
            if (b <= smax) chk((uint32_t)b);
            if (b + 1 <= smax) chk((uint32_t)(b + 1));
        }
        chk(smax); chk(smax - (f > 1 ? f - 1 : 0)); chk(f > 1 ? f : 1); chk(ANS_L < smax ? ANS_L : smax);
    }
    Geom g = make_geom(ANS, 9000001, 9000001);
    if (g.n_calls != 1 || g.upc != 3 || g.n_units != 3) bad++;
    size_t j, st, ln; uint32_t b;
    unit_span(g, 2, j, b, st, ln); if (j != 0 || b != 2 || st != 2u * ANS_BLOCK || ln != 9000001 - 2u * ANS_BLOCK) bad++;
    g = make_geom(RCS2, 10001, 4096);
    if (g.n_calls != 3 || g.upc != 1) bad++;
    call_span(g, 2, st, ln); if (st != 8192 || ln != 1809) bad++;
    return bad;
}

size_t trc_num_chunks(size_t total_len, size_t chunk_len) { return chunk_len ? (total_len + chunk_len - 1) / chunk_len : 0; }
size_t trc_enc_bound(size_t total_len, size_t chunk_len) {
    // every call yields at most its input length, except rccdf4ienc's 4-byte answer on inputs shorter than 4 bytes
    size_t extra = (chunk_len && chunk_len < 4) ? 4 * trc_num_chunks(total_len, chunk_len) : 4;
    return total_len + extra + 64;
}
size_t trc_enc_scratch_bytes(int codec, size_t total_len, size_t chunk_len) {
    Plan p; if (make_plan(codec, total_len, chunk_len, p) != TRC_OK) return 0; return p.total + 256;
}

int trc_enc_batch_dev(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len,
                      const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf,
                      unsigned char *d_out, uint64_t *d_out_off,
                      void *d_scratch, size_t scratch_bytes, void *cuda_stream) {
    return enc_batch_impl(codec, d_in, total_len, chunk_len, d_cdf, cdfnum, chunks_per_cdf, nullptr, d_out, d_out_off, d_scratch, scratch_bytes, cuda_stream);
}
int trc_enc_batch_dev_tab(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len,
                          const trc_tables *tables, size_t chunks_per_cdf,
                          unsigned char *d_out, uint64_t *d_out_off,
                          void *d_scratch, size_t scratch_bytes, void *cuda_stream) {
    if (!tables || !codec_static(codec)) return TRC_E_ARG;
    if (tables->n < n_tables(trc_num_chunks(total_len, chunk_len), chunks_per_cdf) || tables->dev != cur_dev()) return TRC_E_ARG;
    return enc_batch_impl(codec, d_in, total_len, chunk_len, nullptr, tables->cdfnum, chunks_per_cdf, tables, d_out, d_out_off, d_scratch, scratch_bytes, cuda_stream);
}
int trc_dec_batch_dev(int codec, const unsigned char *d_in, const uint64_t *d_in_off,
                      unsigned char *d_out, size_t total_len, size_t chunk_len,
                      const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf,
                      unsigned flags, void *cuda_stream) {
    return dec_batch_impl(codec, d_in, d_in_off, d_out, total_len, chunk_len, d_cdf, cdfnum, chunks_per_cdf, nullptr, flags, cuda_stream);
}
int trc_dec_batch_dev_tab(int codec, const unsigned char *d_in, const uint64_t *d_in_off,
                          unsigned char *d_out, size_t total_len, size_t chunk_len,
                          const trc_tables *tables, size_t chunks_per_cdf, unsigned flags, void *cuda_stream) {
    if (!tables || !codec_static(codec)) return TRC_E_ARG;
    if (tables->n < n_tables(trc_num_chunks(total_len, chunk_len), chunks_per_cdf) || tables->dev != cur_dev()) return TRC_E_ARG;
    return dec_batch_impl(codec, d_in, d_in_off, d_out, total_len, chunk_len, nullptr, tables->cdfnum, chunks_per_cdf, tables, flags, cuda_stream);
}

}  // extern "C"

static int enc_batch_impl(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len, const cdf_t *d_cdf, unsigned cdfnum,
                          size_t chunks_per_cdf, const trc_tables *pre, unsigned char *d_out, uint64_t *d_out_off, void *d_scratch,
                          size_t scratch_bytes, void *cuda_stream) {
    Plan p; int rc = make_plan(codec, total_len, chunk_len, p);
    if (rc != TRC_OK) return rc;
    if (!d_in || !d_out || !d_out_off || !d_scratch) return TRC_E_ARG;
    if (codec_static(codec) && ((!d_cdf && !pre) || cdfnum == 0 || cdfnum > 256)) return TRC_E_ARG;
    uint8_t *sc = (uint8_t *)al256((size_t)d_scratch);
    if ((size_t)(sc - (uint8_t *)d_scratch) + p.total > scratch_bytes) return TRC_E_NOMEM;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    UnitMeta *meta = (UnitMeta *)(sc + p.off_meta);
    CallInfo *calls = (CallInfo *)(sc + p.off_calls);
    uint8_t *slots = sc + p.off_slots;
    uint32_t *recs = (uint32_t *)(sc + p.off_recs);
    const Geom &g = p.g;
    auto blocks = [](size_t n, int nt) { return (unsigned)((n + nt - 1) / nt); };
    g_prof_n = 0;
    const bool v2 = codec_static(codec) && v2_ok(d_in, chunk_len, chunks_per_cdf);
    if (codec == ANSW && !(((uintptr_t)d_in & 3) == 0 && (chunks_per_cdf == 0 || chunks_per_cdf % V2_NT == 0))) return TRC_E_ARG;
    TableSet *tabs = pre ? pre->ts : (TableSet *)(sc + p.off_tabs);
    const bool fused = g_fused && codec == RCS2 && v2;
    if (fused) {                 // TRC_RCS2: coder + offsets + layout in ONE kernel (rcs2_v3.cuh)
        const size_t n_full = total_len / chunk_len;                        // the tensor map covers full chunks only (see rcs2_v3.cuh)
        EncTab2 *t2 = pre ? pre->e2 : (EncTab2 *)tabs;
        unsigned cpcta = 0, ctas = 0;
        e3_shape(g.n_calls, cpcta, ctas);
        if (chunks_per_cdf) {                                               // a table per group of calls (v2_ok: a multiple of 128): CTAs of whole groups' divisors
            cpcta = chunks_per_cdf % 256 == 0 ? 256 : 128;
            ctas = (unsigned)((g.n_calls + cpcta - 1) / cpcta);
        }
        CK(cudaMemsetAsync(sc + p.off_lb, 0, (size_t)(ctas + 1) * 8, st));   // look-back words + tile counter
        if (!pre) { k_build_enctab2<<<(unsigned)n_tables(g.n_calls, chunks_per_cdf), 256, 0, st>>>(d_cdf, cdfnum, t2); CK_LAUNCH(); }
        prof_mark(st);
        const unsigned nt = (2 * cpcta + 31) & ~31u;
        const bool tma = g_enc_tma && n_full && tmap_encode() && n_full < (1ull << 31);
        CUtensorMap tm; memset(&tm, 0, sizeof tm);
        if (tma) {                                                          // input as a 2-D tensor [calls][chunk bytes], box = 16 calls x 128 bytes
            const cuuint64_t dims[2] = { (cuuint64_t)chunk_len, (cuuint64_t)n_full }, strides[1] = { (cuuint64_t)chunk_len };
            const cuuint32_t box[2] = { 128, 16 }, estr[2] = { 1, 1 };
            CUresult r = tmap_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)d_in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed (%d)", (int)r); return TRC_E_CUDA; }
        }
        const size_t smem = e3_smem_bytes(nt, tma);
        rc = dev_attrs(); if (rc) return rc;
        if (tma) k_rcs2_enc3<true><<<ctas, nt, smem, st>>>(tm, d_in, g, g.n_calls, t2, slots, p.slot_stride, cpcta,
                                                          (volatile unsigned long long *)(sc + p.off_lb), d_out_off, d_out, g_force_redo ? 1u : 0u, chunks_per_cdf);
        else     k_rcs2_enc3<false><<<ctas, nt, smem, st>>>(tm, d_in, g, g.n_calls, t2, slots, p.slot_stride, cpcta,
                                                           (volatile unsigned long long *)(sc + p.off_lb), d_out_off, d_out, g_force_redo ? 1u : 0u, chunks_per_cdf);
        CK_LAUNCH();
        prof_mark(st); prof_mark(st); prof_mark(st);
        return TRC_OK;
    }
    if ((v2 || codec == ANSW) && !pre) {   // symbol tables once per launch
        k_build_tables<<<dim3((unsigned)n_tables(g.n_calls, chunks_per_cdf), 1), 256, 0, st>>>(d_cdf, cdfnum, tabs, 0);
        CK_LAUNCH();
    }
    if (pre && !v2 && codec != ANSW) return TRC_E_ARG;                  // prebuilt tables serve the aligned (throughput) kernels only
    const unsigned v2nt = v2_shape(g.n_calls, chunks_per_cdf);
    prof_mark(st);
    switch (codec) {
    case ANS4S: if (v2) k_rans_static_enc_v2<<<blocks(g.n_calls, v2nt), v2nt, 0, st>>>(d_in, g, g.n_calls, tabs, chunks_per_cdf, slots, p.slot_stride, meta);
                else k_rans_static_enc<<<blocks(g.n_calls, RANS_S_NT), RANS_S_NT, 0, st>>>(d_in, g, d_cdf, cdfnum, chunks_per_cdf, slots, p.slot_stride, meta); break;
    case RCS:   if (v2) k_rc_static_enc_v2<1><<<blocks(g.n_calls, v2nt), v2nt, 0, st>>>(d_in, g, g.n_calls, tabs, chunks_per_cdf, slots, p.slot_stride, meta);
                else k_rc_static_enc<1><<<blocks(g.n_calls, RC_S_NT), RC_S_NT, 0, st>>>(d_in, g, d_cdf, cdfnum, chunks_per_cdf, slots, p.slot_stride, meta); break;
    case RCS2:  if (v2) { unsigned cpcta, ctas; lpc_shape(g.n_calls, chunks_per_cdf, cpcta, ctas);
                          k_rcs2_enc_lpc<<<ctas, (2 * cpcta + 31) & ~31u, 0, st>>>(d_in, g, g.n_calls, tabs, chunks_per_cdf, slots, p.slot_stride, meta, cpcta); }
                else k_rc_static_enc<2><<<blocks(g.n_calls, RC_S_NT), RC_S_NT, 0, st>>>(d_in, g, d_cdf, cdfnum, chunks_per_cdf, slots, p.slot_stride, meta); break;
    case ANSW:  k_answ_enc<<<blocks(g.n_calls, ANSW_WPB), ANSW_WPB * 32, 0, st>>>(d_in, g, tabs, chunks_per_cdf, slots, p.slot_stride, meta); break;
    case ANS4:  k_rans_adapt_enc<M_NIB, AD_NT_NIB><<<blocks(g.n_units, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, g, slots, p.slot_stride, recs, p.rec_stride, nullptr, meta); break;
    // adaptive byte rANS: many small units -> one lane per unit (throughput); few large units -> one warp per unit (latency)
    case ANS:   if (g.n_units >= COOP_MIN_LANE_UNITS) k_rans_adapt_enc<M_BYTE, AD_NT_BYTE><<<blocks(g.n_units, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, g, slots, p.slot_stride, recs, p.rec_stride, nullptr, meta);
                else { rc = launch_ans_enc3(false, d_in, g, slots, p.slot_stride, recs, p.rec_stride, meta, st); if (rc) return rc; }
                break;
    case ANS1:  rc = launch_ans_enc3(true, d_in, g, slots, p.slot_stride, recs, p.rec_stride, meta, st); if (rc) return rc; break;
    case RC:    if (g.n_calls >= COOP_MIN_LANE_UNITS) k_rc_adapt_enc<R_BYTE1, AD_NT_BYTE><<<blocks(g.n_calls, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, g, slots, p.slot_stride, meta);
                else k_rc_byte_enc_coop<1><<<blocks(g.n_calls, COOP_WPB), COOP_WPB * 32, COOP_WPB * O1_CTX_ENTRIES * 2, st>>>(d_in, g, slots, p.slot_stride, meta, g_force_redo); break;
    case RCI:   if (g.n_calls >= COOP_MIN_LANE_UNITS) k_rc_adapt_enc<R_BYTE2, AD_NT_BYTE><<<blocks(g.n_calls, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, g, slots, p.slot_stride, meta);
                else k_rc_byte_enc_coop<2><<<blocks(g.n_calls, COOP_WPB), COOP_WPB * 32, COOP_WPB * O1_CTX_ENTRIES * 2, st>>>(d_in, g, slots, p.slot_stride, meta, g_force_redo); break;
    case RC4:   k_rc_adapt_enc<R_NIB1, AD_NT_NIB><<<blocks(g.n_calls, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, g, slots, p.slot_stride, meta); break;
    case RC4I:  k_rc_adapt_enc<R_NIB2, AD_NT_NIB><<<blocks(g.n_calls, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, g, slots, p.slot_stride, meta); break;
    case RC8:   k_rc_v8_enc<1><<<blocks(g.n_calls, V8_NT), V8_NT, 0, st>>>(d_in, g, slots, p.slot_stride, meta); break;
    case RCI8:  k_rc_v8_enc<2><<<blocks(g.n_calls, V8_NT), V8_NT, 0, st>>>(d_in, g, slots, p.slot_stride, meta); break;
    default:
        if (codec_vlc_ans(codec)) k_vlc_ans_enc<<<blocks(g.n_calls, VLC_NT), VLC_NT, 0, st>>>(d_in, g, vlc_param(codec), slots, p.slot_stride, recs, p.rec_stride, meta);
        else if (codec_vlc(codec)) k_vlc_rc_enc<<<blocks(g.n_calls, VLC_NT), VLC_NT, 0, st>>>(d_in, g, vlc_param(codec), slots, p.slot_stride, meta);
        break;
    }
    CK_LAUNCH(); prof_mark(st);
    k_resolve<<<blocks(g.n_calls, 256), 256, 0, st>>>(g, codec_blocked(codec) ? 1 : 0, meta, calls);
    CK_LAUNCH();
    k_scan<<<1, SCAN_NT, 0, st>>>(g.n_calls, calls, d_out_off);
    CK_LAUNCH(); prof_mark(st);
    size_t seg = PACK_SEG_MIN;
    while ((p.slot_stride + seg - 1) / seg > 65535) seg <<= 1;
    dim3 pg((unsigned)g.n_units, (unsigned)((p.slot_stride + seg - 1) / seg));
    if (pg.y == 1) k_pack_small<<<blocks(g.n_units, PACKS_WARPS), PACKS_WARPS * 32, 0, st>>>(d_in, g, slots, p.slot_stride, meta, calls, d_out_off, d_out);
    else k_pack<<<pg, PACK_NT, 0, st>>>(d_in, g, slots, p.slot_stride, meta, calls, d_out_off, d_out, seg);
    CK_LAUNCH(); prof_mark(st);
    return TRC_OK;
}

static int dec_batch_impl(int codec, const unsigned char *d_in, const uint64_t *d_in_off, unsigned char *d_out, size_t total_len,
                          size_t chunk_len, const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf, const trc_tables *pre,
                          unsigned flags, void *cuda_stream) {
    Plan p; int rc = make_plan(codec, total_len, chunk_len, p);
    if (rc != TRC_OK) return rc;
    if (!d_in || !d_in_off || !d_out) return TRC_E_ARG;
    if (codec_static(codec) && ((!d_cdf && !pre) || cdfnum == 0 || cdfnum > 256)) return TRC_E_ARG;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Geom g = p.g; g.upc = 1; g.n_units = g.n_calls;     // decoders work per call
    auto blocks = [](size_t n, int nt) { return (unsigned)((n + nt - 1) / nt); };
    g_prof_n = 0;
    if (codec == ANSW && !((((uintptr_t)d_in | (uintptr_t)d_out) & 3) == 0 && (chunks_per_cdf == 0 || chunks_per_cdf % V2_NT == 0))) return TRC_E_ARG;
    static const int g_dec3 = getenv("TRC_DEC3") ? atoi(getenv("TRC_DEC3")) : 1;     // 0: previous generation of the TRC_RCS2 decoder (A/B runs)
    if (codec == RCS2 && g_dec3 && v2_ok(d_out, chunk_len, chunks_per_cdf) && ((uintptr_t)d_in & 15) == 0) {
        rc = dev_attrs(); if (rc) return rc;
        DecTab2 *t2 = pre ? pre->d2 : nullptr;
        if (!pre) {
            rc = pool_keep(); if (rc) return rc;
            const size_t nt = n_tables(g.n_calls, chunks_per_cdf);
            CK(cudaMallocAsync((void **)&t2, nt * sizeof(DecTab2), st));
            k_build_dectab2<<<dim3((unsigned)nt, 1 + LUT_PARTS), 1024, 0, st>>>(d_cdf, cdfnum, t2);
            g_launches++;
        }
        prof_mark(st);
        unsigned cpcta, ctas; lpc_shape(g.n_calls, chunks_per_cdf, cpcta, ctas, true);
        const unsigned nt2 = (2 * cpcta + 31) & ~31u;
        k_rcs2_dec3<<<ctas, nt2, D3_SMEM, st>>>(d_in, d_in_off, d_out, g, g.n_calls, t2, cdfnum, chunks_per_cdf, cpcta);
        g_launches++; prof_mark(st);
        cudaError_t e = cudaPeekAtLastError();
        if (!pre) cudaFreeAsync(t2, st);
        CK(e);
        return TRC_OK;
    }
    if (codec == ANSW || (codec_static(codec) && v2_ok(d_out, chunk_len, chunks_per_cdf) && ((uintptr_t)d_in & 15) == 0)) {
        TableSet *tabs = pre ? pre->ts : nullptr;
        if (!pre) {
            rc = pool_keep(); if (rc) return rc;
            const size_t nt = n_tables(g.n_calls, chunks_per_cdf);
            CK(cudaMallocAsync((void **)&tabs, nt * sizeof(TableSet), st));
            k_build_tables<<<dim3((unsigned)nt, 1 + LUT_PARTS), 1024, 0, st>>>(d_cdf, cdfnum, tabs, 1);
            g_launches++;
        }
        prof_mark(st);
        if (codec == ANSW) k_answ_dec<<<blocks(g.n_calls, ANSW_WPB), ANSW_WPB * 32, 0, st>>>(d_in, d_in_off, d_out, g, tabs, chunks_per_cdf);
        else if (codec == ANS4S) { const unsigned v2nt = v2_shape(g.n_calls, chunks_per_cdf);
            rc = dev_attrs(); if (rc) return rc;
            k_rans_static_dec_v2<<<blocks(g.n_calls, v2nt), v2nt, RING_W * v2nt * sizeof(uint32_t), st>>>(d_in, d_in_off, d_out, g, g.n_calls, tabs, chunks_per_cdf, flags); }
        else if (codec == RCS) { const unsigned v2nt = v2_shape(g.n_calls, chunks_per_cdf); k_rc_static_dec_v2<1><<<blocks(g.n_calls, v2nt), v2nt, 0, st>>>(d_in, d_in_off, d_out, g, g.n_calls, tabs, cdfnum, chunks_per_cdf); }
        else { unsigned cpcta, ctas; lpc_shape(g.n_calls, chunks_per_cdf, cpcta, ctas);
               const unsigned nt = (2 * cpcta + 31) & ~31u;
               rc = dev_attrs(); if (rc) return rc;
               k_rcs2_dec_lpc<<<ctas, nt, RING_W * nt * sizeof(uint32_t), st>>>(d_in, d_in_off, d_out, g, g.n_calls, tabs, cdfnum, chunks_per_cdf, cpcta); }
        g_launches++; prof_mark(st);
        cudaError_t e = cudaPeekAtLastError();
        if (!pre) cudaFreeAsync(tabs, st);
        CK(e);
        return TRC_OK;
    }
    if (pre) return TRC_E_ARG;                                          // prebuilt tables serve the aligned (throughput) kernels only
    prof_mark(st);
    switch (codec) {
    case ANS4S: k_rans_static_dec<<<blocks(g.n_calls, RANS_SD_NT), RANS_SD_NT, 0, st>>>(d_in, d_in_off, d_out, g, d_cdf, cdfnum, chunks_per_cdf, flags); break;
    case RCS:   k_rc_static_dec<1><<<blocks(g.n_calls, RC_SD_NT), RC_SD_NT, 0, st>>>(d_in, d_in_off, d_out, g, d_cdf, cdfnum, chunks_per_cdf); break;
    case RCS2:  k_rc_static_dec<2><<<blocks(g.n_calls, RC_SD_NT), RC_SD_NT, 0, st>>>(d_in, d_in_off, d_out, g, d_cdf, cdfnum, chunks_per_cdf); break;
    case ANS4:  k_rans_adapt_dec<M_NIB, AD_NT_NIB><<<blocks(g.n_calls, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, d_in_off, d_out, g, nullptr, flags); break;
    case ANS:   if (g.n_calls >= COOP_MIN_LANE_UNITS) k_rans_adapt_dec<M_BYTE, AD_NT_BYTE><<<blocks(g.n_calls, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, d_in_off, d_out, g, nullptr, flags);
                else k_ans_dec3<false><<<blocks(g.n_calls, 2 * D3_WPB), D3_WPB * 32, d3_smem_bytes<false>(), st>>>(d_in, d_in_off, d_out, g);
                break;
    case ANS1: {
        rc = dev_attrs(); if (rc) return rc;
        static const int g_o1g = getenv("TRC_O1_GLOBAL") ? atoi(getenv("TRC_O1_GLOBAL")) : 1;   // 0: always the one-call-per-SM kernel (A/B runs)
        if (g_o1g && g.n_calls >= 5 * (size_t)sm_count()) {               // many calls: half-warp per call, low-nibble tables in global memory
            // (one call per SM runs a byte in ~200 cycles, this kernel in ~860 but with every call at once: break-even ~4.3 calls per SM)
            rc = pool_keep(); if (rc) return rc;
            const size_t want = (g.n_calls + 2 * G1_WPB - 1) / (2 * G1_WPB), cap = (size_t)sm_count() * 3;
            const unsigned ctas = (unsigned)(want < cap ? want : cap);
            uint16_t *mbl = nullptr;
            CK(cudaMallocAsync((void **)&mbl, (size_t)ctas * 2 * G1_WPB * G1_MBL_ENTRIES * sizeof(uint16_t), st));
            k_ans1_dec_g<<<ctas, G1_WPB * 32, G1_SMEM, st>>>(d_in, d_in_off, d_out, g, mbl);
            g_launches++; prof_mark(st);
            cudaError_t e = cudaPeekAtLastError();
            cudaFreeAsync(mbl, st);
            CK(e);
            return TRC_OK;
        }
        k_ans_dec3<true><<<(unsigned)(g.n_calls < (size_t)sm_count() ? g.n_calls : (size_t)sm_count()), 32, d3_smem_bytes<true>(), st>>>(d_in, d_in_off, d_out, g);
        break;
    }
    case RC:    if (g.n_calls >= COOP_MIN_LANE_UNITS) k_rc_adapt_dec<R_BYTE1, AD_NT_BYTE><<<blocks(g.n_calls, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, d_in_off, d_out, g);
                else k_rc_dec3<1><<<blocks(g.n_calls, 2 * D3_WPB), D3_WPB * 32, r3_smem_bytes<1>(), st>>>(d_in, d_in_off, d_out, g);
                break;
    case RCI:   if (g.n_calls >= COOP_MIN_LANE_UNITS) k_rc_adapt_dec<R_BYTE2, AD_NT_BYTE><<<blocks(g.n_calls, AD_NT_BYTE), AD_NT_BYTE, 0, st>>>(d_in, d_in_off, d_out, g);
                else k_rc_dec3<2><<<blocks(g.n_calls, 2 * D3_WPB), D3_WPB * 32, r3_smem_bytes<2>(), st>>>(d_in, d_in_off, d_out, g);
                break;
    case RC4:   k_rc_adapt_dec<R_NIB1, AD_NT_NIB><<<blocks(g.n_calls, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, d_in_off, d_out, g); break;
    case RC4I:  k_rc_adapt_dec<R_NIB2, AD_NT_NIB><<<blocks(g.n_calls, AD_NT_NIB), AD_NT_NIB, 0, st>>>(d_in, d_in_off, d_out, g); break;
    case RC8:   k_rc_v8_dec<1><<<blocks(g.n_calls, V8_NT), V8_NT, 0, st>>>(d_in, d_in_off, d_out, g); break;
    case RCI8:  k_rc_v8_dec<2><<<blocks(g.n_calls, V8_NT), V8_NT, 0, st>>>(d_in, d_in_off, d_out, g); break;
    default:
        if (codec_vlc_ans(codec)) k_vlc_dec<false><<<blocks(g.n_calls, VLC_NT), VLC_NT, 0, st>>>(d_in, d_in_off, d_out, g, vlc_param(codec));
        else if (codec_vlc(codec)) k_vlc_dec<true><<<blocks(g.n_calls, VLC_NT), VLC_NT, 0, st>>>(d_in, d_in_off, d_out, g, vlc_param(codec));
        break;
    }
    CK_LAUNCH(); prof_mark(st);
    return TRC_OK;
}

extern "C" {

int trc_cdfini_batch_dev(const unsigned char *d_in, size_t total_len, size_t chunk_len,
                         cdf_t *d_cdf, unsigned cdfnum, int *d_status, void *cuda_stream) {
    if (!d_in || !d_cdf || total_len == 0 || chunk_len == 0 || cdfnum == 0 || cdfnum > 256) return TRC_E_ARG;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Geom g = make_geom(RCS, total_len, chunk_len);
    unsigned long long *hist = nullptr;
    CK(cudaMallocAsync((void **)&hist, g.n_calls * 256 * 8, st));
    CK(cudaMemsetAsync(hist, 0, g.n_calls * 256 * 8, st));
    size_t cl = chunk_len < total_len ? chunk_len : total_len;
    dim3 hg((unsigned)g.n_calls, (unsigned)((cl + HIST_SEG - 1) / HIST_SEG));
    k_hist<<<hg, HIST_NT, 0, st>>>(d_in, g, hist);
    g_launches++;
    k_cdf_finalize<<<(unsigned)((g.n_calls + 63) / 64), 64, 0, st>>>(g, hist, d_cdf, cdfnum, d_status);
    g_launches++;
    cudaError_t e = cudaPeekAtLastError();
    cudaFreeAsync(hist, st);
    CK(e);
    return TRC_OK;
}

// launch shapes of the TRC_RCS2 kernels for a batch of n_calls calls on a device with n_sm SMs (no device needed: host logic
// only; tests/test_abi.py checks its invariants).  shape[0..1] = encoder (calls per CTA, CTAs), shape[2..3] = decoder.
int trc_debug_rcs2_shapes(size_t n_calls, size_t chunks_per_cdf, int n_sm, unsigned shape[4]) {
    if (!shape || n_sm <= 0 || n_calls == 0) return TRC_E_ARG;
    e3_shape(n_calls, shape[0], shape[1], n_sm);
    lpc_shape(n_calls, chunks_per_cdf, shape[2], shape[3], true, n_sm);
    return TRC_OK;
}

// phase timing probe of k_rcs2_enc3 (tools/enc_phases.py): d_buf = ctas x 16 warps x 8 uint64, or NULL to switch it off
int trc_debug_enc_times(void *d_buf) {
    unsigned long long *p = (unsigned long long *)d_buf;
    CK(cudaMemcpyToSymbol(g_e3_times, &p, sizeof p));
    return TRC_OK;
}

// ---- peer-memory plumbing for the multi-GPU gather (shard.PeerGather) ------------------------------------------
int trc_dev_alloc(void **p, size_t bytes) { CK(cudaMalloc(p, bytes)); CK(cudaMemset(*p, 0, bytes)); return TRC_OK; }
int trc_dev_free(void *p) { CK(cudaFree(p)); return TRC_OK; }
int trc_ipc_export(void *p, unsigned char *handle64) {
    cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64); return TRC_OK;
}
int trc_ipc_open(const unsigned char *handle64, void **p) {
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess)); return TRC_OK;
}
int trc_ipc_close(void *p) { CK(cudaIpcCloseMemHandle(p)); return TRC_OK; }
int trc_memcpy_dev(void *dst, const void *src, size_t bytes, void *cuda_stream) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream)); return TRC_OK;
}
// copy *d_len bytes (or fixed_len when d_len is NULL; rounded up to 16, at most cap when cap != 0) from src to dst, then publish
// the length at dst_len and `seq` at dst_flag (see k_push for the completion protocol).  d_counter: one zeroed uint32 in LOCAL
// device memory per concurrent push (NULL only with dst_flag == NULL: length published without a completion guarantee).
// ack / ack_need (optional): do not touch dst before *ack >= ack_need (the consumer's acknowledgement of the push that used the slot last).
// skip: bytes at the front of the stream that a copy engine already moved (stream-ordered before this call); 0 = copy everything here.
int trc_push_dev(void *dst, const void *src, const uint64_t *d_len, size_t fixed_len, size_t cap, uint64_t *dst_len,
                 uint64_t *dst_flag, uint64_t seq, unsigned int *d_counter, const uint64_t *ack, uint64_t ack_need, size_t skip, void *cuda_stream) {
    if (!dst || !src || (((uintptr_t)dst | (uintptr_t)src) & 15) || (dst_flag && !d_counter) || (skip & 15) || (cap && skip > cap)) return TRC_E_ARG;
    { int rc0 = dev_attrs(); if (rc0) return rc0; }
    static const int ctas_full = getenv("TRC_PUSH_CTAS") ? atoi(getenv("TRC_PUSH_CTAS")) : 32;
    const int ctas = skip ? 4 : ctas_full;                            // a remainder behind a copy-engine transfer is small   // few CTAs: enough stores in flight for NVLink, little SM time stolen from the coders
    k_push<<<ctas, 256, 0, (cudaStream_t)cuda_stream>>>((uint4 *)dst, (const uint4 *)src, d_len, fixed_len, cap, dst_len, dst_flag, seq, d_counter, ack, ack_need, skip);
    CK_LAUNCH();
    return TRC_OK;
}
// block `cuda_stream` until flags[0..n) >= seq (consumer side); d_status (optional, zeroed): bit 0 set if a length overflowed its slot
int trc_wait_flags_dev(const uint64_t *flags, const uint64_t *lens, unsigned n, uint64_t seq, unsigned int *d_status, void *cuda_stream) {
    if (!flags || !n || n > 1024) return TRC_E_ARG;
    k_wait_flags<<<1, (n + 31) & ~31u, 0, (cudaStream_t)cuda_stream>>>(flags, lens, n, seq, d_status);
    CK_LAUNCH();
    return TRC_OK;
}
// consumer acknowledgement: the slot set of push `seq` may be reused (stream-ordered after whatever consumed it)
int trc_ack_dev(uint64_t *ack, uint64_t seq, void *cuda_stream) {
    if (!ack) return TRC_E_ARG;
    k_ack<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(ack, seq);
    CK_LAUNCH();
    return TRC_OK;
}

}  // extern "C"

// ===========================================================================================================
// Host-pointer layer: one lazily created context per process (device buffers grow, never shrink).
// ===========================================================================================================
namespace {
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int need(size_t n) {
        if (n <= cap) return TRC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t c = n + n / 8 + 4096;
        if (cudaMalloc(&p, c) != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaMalloc(%zu) failed: %s", c, cudaGetErrorString(cudaGetLastError())); return TRC_E_NOMEM; }
        cap = c; return TRC_OK;
    }
};
struct Ctx {
    std::mutex mu;
    bool init = false;
    cudaStream_t st = nullptr, s_h2d = nullptr, s_d2h = nullptr;   // compute, upload, download
    static constexpr int NCS = 8;
    cudaStream_t cs[NCS];                                            // sub-batch kernels run concurrently on these
    DevBuf sub_scratch[NCS];
    DevBuf in, out, off, cdf, scratch, status;
    cudaEvent_t ev[2][64];                                           // [uploaded | coded] per sub-batch
    uint64_t *h_off = nullptr, *h_off_dev = nullptr; size_t h_off_cap = 0;   // mapped pinned staging for sub-batch offsets
    int dev = 0;
    // every host-pointer / drop-in entry point starts here: the calling thread is switched to the context's device
    int ensure() {
        CK(cudaSetDevice(dev));
        if (init) return TRC_OK;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
        for (auto &x : cs) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        for (auto &row : ev) for (auto &e : row) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        init = true; return TRC_OK;
    }
    int need_hoff(size_t n) {
        if (n <= h_off_cap) return TRC_OK;
        if (h_off) cudaFreeHost(h_off);
        h_off = nullptr; h_off_cap = 0;
        if (cudaHostAlloc((void **)&h_off, (n + 1024) * 8, cudaHostAllocMapped) != cudaSuccess) return TRC_E_NOMEM;
        if (cudaHostGetDevicePointer((void **)&h_off_dev, h_off, 0) != cudaSuccess) return TRC_E_CUDA;
        h_off_cap = n + 1024; return TRC_OK;
    }
};
Ctx g_ctxs[MAX_DEV];                                                  // one lazily created context per device
std::mutex g_ctxs_mu;
Ctx &ctx_for(int dev) {
    if (dev < 0 || dev >= MAX_DEV) dev = 0;
    std::lock_guard<std::mutex> lk(g_ctxs_mu);
    g_ctxs[dev].dev = dev;
    return g_ctxs[dev];
}

// Sub-batching for the host-pointer calls: upload of sub-batch i+1, coding of i and download of i-1 overlap on
// three streams, so a host->host call costs about max(H2D, D2H) instead of their sum.  Sub-batches are whole
// calls, aligned to table groups, ~8 MiB each, at most 64, and only when a sub-batch still holds >= 1024 calls (the
// coders are latency-bound per call: a launch with few calls takes as long as one with thousands).
constexpr size_t SUB_BYTES = 8u << 20;
static size_t sub_calls(size_t n_calls, size_t chunk_len, size_t cpc) {
    size_t gsz = SUB_BYTES / chunk_len;
    if (gsz < 1) gsz = 1;
    if (gsz * 64 < n_calls) gsz = (n_calls + 63) / 64;
    const size_t q = cpc ? cpc : V2_NT;                              // keep table groups / CTAs whole
    gsz = (gsz + q - 1) / q * q;
    return gsz;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const bool g_trace = getenv("TRC_TRACE") != nullptr;

static int host_enc_pipelined(Ctx &c, int codec, const unsigned char *in, size_t total_len, size_t chunk_len, const cdf_t *cdf,
                              unsigned cdfnum, size_t cpc, unsigned char *out, uint64_t *out_off, size_t *out_len, size_t n, size_t gsz) {
    int rc;
    const size_t nsub = (n + gsz - 1) / gsz, sub_len = gsz * chunk_len, stride = al256(trc_enc_bound(sub_len, chunk_len));   // (rccdf4ienc answers 4 bytes on inputs shorter than 4)
    Plan p; rc = make_plan(codec, sub_len < total_len ? sub_len : total_len, chunk_len, p); if (rc) return rc;
    if ((rc = c.in.need(total_len + 64)) || (rc = c.out.need(nsub * stride)) || (rc = c.off.need((n + nsub + 1) * 8)) ||
        (rc = c.need_hoff(n + nsub + 1))) return rc;
    for (auto &sb : c.sub_scratch) if ((rc = sb.need(p.total + 256))) return rc;
    if (codec_static(codec)) {
        if (!cdf) return TRC_E_ARG;
        size_t nt = n_tables(n, cpc), bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
        if ((rc = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return rc;
        CK(cudaMemcpyAsync(c.cdf.p, cdf, bytes, cudaMemcpyHostToDevice, c.s_h2d));
    }
    const double t_start = now_ms();
    for (size_t i = 0; i < nsub; i++) {
        const size_t o = i * sub_len, len = total_len - o < sub_len ? total_len - o : sub_len, calls = (len + chunk_len - 1) / chunk_len;
        CK(cudaMemcpyAsync((uint8_t *)c.in.p + o, in + o, len, cudaMemcpyHostToDevice, c.s_h2d));
        CK(cudaEventRecord(c.ev[0][i], c.s_h2d));
        cudaStream_t cst = c.cs[i % Ctx::NCS];
        DevBuf &sb = c.sub_scratch[i % Ctx::NCS];
        CK(cudaStreamWaitEvent(cst, c.ev[0][i], 0));
        uint64_t *doff = (uint64_t *)c.off.p + i * (gsz + 1);
        rc = trc_enc_batch_dev(codec, (const unsigned char *)c.in.p + o, len, chunk_len,
                               (const cdf_t *)c.cdf.p + (cpc ? (i * gsz / cpc) * CDF_STRIDE : 0), cdfnum, cpc,
                               (unsigned char *)c.out.p + i * stride, doff, sb.p, sb.cap, cst);
        if (rc) return rc;
        k_publish_u64<<<(unsigned)((calls + 1 + 255) / 256), 256, 0, cst>>>(doff, c.h_off_dev + i * (gsz + 1), calls + 1);
        CK_LAUNCH();
        CK(cudaEventRecord(c.ev[1][i], cst));
    }
    uint64_t run = 0;
    const double t_enq = now_ms();
    if (g_trace) fprintf(stderr, "enc: enqueue took %.3f ms\n", t_enq - t_start);
    for (size_t i = 0; i < nsub; i++) {
        const size_t o = i * sub_len, len = total_len - o < sub_len ? total_len - o : sub_len, calls = (len + chunk_len - 1) / chunk_len;
        CK(cudaEventSynchronize(c.ev[1][i]));
        if (g_trace) fprintf(stderr, "  sub %zu coded at +%.3f ms\n", i, now_ms() - t_enq);
        const uint64_t *ho = c.h_off + i * (gsz + 1);
        const uint64_t tot = ho[calls];
        CK(cudaMemcpyAsync(out + run, (const uint8_t *)c.out.p + i * stride, tot, cudaMemcpyDeviceToHost, c.s_d2h));
        if (out_off) for (size_t k = 0; k < calls; k++) out_off[i * gsz + k] = run + ho[k];
        run += tot;
    }
    if (out_off) out_off[n] = run;
    CK(cudaStreamSynchronize(c.s_d2h));
    if (g_trace) fprintf(stderr, "  enc done at +%.3f ms after enqueue end (nsub %zu)\n", now_ms() - t_enq, nsub);
    if (out_len) *out_len = (size_t)run;
    return TRC_OK;
}

static int host_dec_pipelined(Ctx &c, int codec, const unsigned char *in, const uint64_t *in_off, unsigned char *out, size_t total_len,
                              size_t chunk_len, const cdf_t *cdf, unsigned cdfnum, size_t cpc, unsigned flags, size_t n, size_t gsz) {
    int rc;
    const size_t nsub = (n + gsz - 1) / gsz, sub_len = gsz * chunk_len;
    const size_t in_bytes = (size_t)in_off[n];
    if ((rc = c.in.need(in_bytes + 64)) || (rc = c.out.need(total_len + 64)) || (rc = c.off.need((n + 1) * 8))) return rc;
    CK(cudaMemcpyAsync(c.off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, c.s_h2d));
    if (codec_static(codec)) {
        if (!cdf) return TRC_E_ARG;
        size_t nt = n_tables(n, cpc), bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
        if ((rc = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return rc;
        CK(cudaMemcpyAsync(c.cdf.p, cdf, bytes, cudaMemcpyHostToDevice, c.s_h2d));
    }
    for (size_t i = 0; i < nsub; i++) {
        const size_t c0 = i * gsz, c1 = c0 + gsz < n ? c0 + gsz : n;
        const size_t o = c0 * chunk_len, len = total_len - o < sub_len ? total_len - o : sub_len;
        const size_t s0 = (size_t)in_off[c0], s1 = (size_t)in_off[c1];
        CK(cudaMemcpyAsync((uint8_t *)c.in.p + s0, in + s0, s1 - s0, cudaMemcpyHostToDevice, c.s_h2d));
        CK(cudaEventRecord(c.ev[0][i], c.s_h2d));
        cudaStream_t cst = c.cs[i % Ctx::NCS];
        CK(cudaStreamWaitEvent(cst, c.ev[0][i], 0));
        rc = trc_dec_batch_dev(codec, (const unsigned char *)c.in.p, (const uint64_t *)c.off.p + c0, (unsigned char *)c.out.p + o, len, chunk_len,
                               (const cdf_t *)c.cdf.p + (cpc ? (c0 / cpc) * CDF_STRIDE : 0), cdfnum, cpc, flags, cst);
        if (rc) return rc;
        CK(cudaEventRecord(c.ev[1][i], cst));
        CK(cudaStreamWaitEvent(c.s_d2h, c.ev[1][i], 0));
        CK(cudaMemcpyAsync(out + o, (const uint8_t *)c.out.p + o, len, cudaMemcpyDeviceToHost, c.s_d2h));
    }
    CK(cudaStreamSynchronize(c.s_d2h));
    return TRC_OK;
}

int host_enc(int dev, int codec, const unsigned char *in, size_t total_len, size_t chunk_len, const cdf_t *cdf, unsigned cdfnum,
             size_t chunks_per_cdf, unsigned char *out, uint64_t *out_off, size_t *out_len) {
    Ctx &c = ctx_for(dev);
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = c.ensure(); if (rc) return rc;
    Plan p; rc = make_plan(codec, total_len, chunk_len, p); if (rc) return rc;
    const size_t n = p.g.n_calls;
    if (p.g.upc == 1 && total_len >= 4 * SUB_BYTES && chunk_len <= SUB_BYTES) {
        const size_t gsz = sub_calls(n, chunk_len, chunks_per_cdf);
        if (gsz < n && gsz >= 1024) return host_enc_pipelined(c, codec, in, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, out, out_off, out_len, n, gsz);
    }
    if ((rc = c.in.need(total_len + 64)) || (rc = c.out.need(trc_enc_bound(total_len, chunk_len))) ||
        (rc = c.off.need((n + 1) * 8)) || (rc = c.scratch.need(p.total + 256))) return rc;
    CK(cudaMemcpyAsync(c.in.p, in, total_len, cudaMemcpyHostToDevice, c.st));
    if (codec_static(codec)) {
        if (!cdf) return TRC_E_ARG;
        size_t nt = chunks_per_cdf ? (n + chunks_per_cdf - 1) / chunks_per_cdf : 1;
        // the last table may be shorter than TRC_CDF_STRIDE entries in the caller's array
        size_t bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
        if ((rc = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return rc;
        CK(cudaMemcpyAsync(c.cdf.p, cdf, bytes, cudaMemcpyHostToDevice, c.st));
    }
    rc = trc_enc_batch_dev(codec, (const unsigned char *)c.in.p, total_len, chunk_len, (const cdf_t *)c.cdf.p, cdfnum, chunks_per_cdf,
                           (unsigned char *)c.out.p, (uint64_t *)c.off.p, c.scratch.p, c.scratch.cap, c.st);
    if (rc) return rc;
    uint64_t total = 0;
    if (out_off) {
        CK(cudaMemcpyAsync(out_off, c.off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, c.st));
        CK(cudaStreamSynchronize(c.st));
        total = out_off[n];
    } else {
        CK(cudaMemcpyAsync(&total, (uint64_t *)c.off.p + n, 8, cudaMemcpyDeviceToHost, c.st));
        CK(cudaStreamSynchronize(c.st));
    }
    CK(cudaMemcpyAsync(out, c.out.p, total, cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    if (out_len) *out_len = (size_t)total;
    return TRC_OK;
}

// in_bytes: how many bytes of `in` to ship (== in_off[n] for the batch API)
int host_dec(int dev, int codec, const unsigned char *in, const uint64_t *in_off, size_t in_bytes, unsigned char *out, size_t total_len,
             size_t chunk_len, const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, unsigned flags) {
    Ctx &c = ctx_for(dev);
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = c.ensure(); if (rc) return rc;
    Plan p; rc = make_plan(codec, total_len, chunk_len, p); if (rc) return rc;
    const size_t n = p.g.n_calls;
    if (total_len >= 4 * SUB_BYTES && chunk_len <= SUB_BYTES && in_bytes == (size_t)in_off[n]) {
        const size_t gsz = sub_calls(n, chunk_len, chunks_per_cdf);
        if (gsz < n && gsz >= 1024) return host_dec_pipelined(c, codec, in, in_off, out, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, flags, n, gsz);
    }
    if ((rc = c.in.need(in_bytes + 64)) || (rc = c.out.need(total_len + 64)) || (rc = c.off.need((n + 1) * 8))) return rc;
    CK(cudaMemcpyAsync(c.off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, c.st));
    CK(cudaMemcpyAsync(c.in.p, in, in_bytes, cudaMemcpyHostToDevice, c.st));
    if (codec_static(codec)) {
        if (!cdf) return TRC_E_ARG;
        size_t nt = chunks_per_cdf ? (n + chunks_per_cdf - 1) / chunks_per_cdf : 1;
        size_t bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
        if ((rc = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return rc;
        CK(cudaMemcpyAsync(c.cdf.p, cdf, bytes, cudaMemcpyHostToDevice, c.st));
    }
    rc = trc_dec_batch_dev(codec, (const unsigned char *)c.in.p, (const uint64_t *)c.off.p, (unsigned char *)c.out.p, total_len, chunk_len,
                           (const cdf_t *)c.cdf.p, cdfnum, chunks_per_cdf, flags, c.st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c.out.p, total_len, cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    return TRC_OK;
}


// ---- one process, several GPUs (include/trc_b200.h trc_*_batch_host_multi) -----------------------------------------------
// Chunks are independent, so the batch is cut into contiguous shards of whole chunks (whole table groups), one per device,
// block b -> device b * n_dev / n_blocks like SURVEY.md section 8e.  One host thread per device drives its own context.
// Encode: every device codes its shard, the threads meet once to turn the shard sizes into base offsets (the only exchange),
// then each device downloads its packed stream straight to its place in `out`.  Decode is the mirror: the chunk directory
// (in_off) tells every device which slice of the stream it needs; nothing is exchanged.
struct Shard { size_t c0, c1; };                                      // chunk range
static void make_shards(size_t n_chunks, size_t cpc, int n_dev, std::vector<Shard> &sh) {
    const size_t grp = cpc ? cpc : 1, n_grp = (n_chunks + grp - 1) / grp;
    sh.resize(n_dev);
    for (int r = 0; r < n_dev; r++) {
        size_t a = ((size_t)r * n_grp + n_dev - 1) / n_dev * grp, b = ((size_t)(r + 1) * n_grp + n_dev - 1) / n_dev * grp;
        sh[r].c0 = a < n_chunks ? a : n_chunks; sh[r].c1 = b < n_chunks ? b : n_chunks;
    }
}
struct SpinBarrier {                                                  // the threads of one call meet once or twice: no need for anything heavier
    std::atomic<int> count{0}, phase{0}; int n;
    explicit SpinBarrier(int n_) : n(n_) {}
    void wait() {
        const int ph = phase.load();
        if (count.fetch_add(1) + 1 == n) { count.store(0); phase.store(ph + 1); }
        else while (phase.load() == ph) std::this_thread::yield();
    }
};

static int multi_enc(int codec, const int *devs, int n_dev, const unsigned char *in, size_t total_len, size_t chunk_len, const cdf_t *cdf,
                     unsigned cdfnum, size_t cpc, unsigned char *out, uint64_t *out_off, size_t *out_len) {
    const size_t n = trc_num_chunks(total_len, chunk_len);
    std::vector<Shard> sh; make_shards(n, cpc, n_dev, sh);
    std::vector<uint64_t> size(n_dev, 0), base(n_dev + 1, 0);
    std::vector<int> rcs(n_dev, TRC_OK);
    std::vector<std::string> errs(n_dev);
    std::vector<std::vector<uint64_t>> offs(n_dev);
    SpinBarrier bar(n_dev);
    auto work = [&](int r) {
        const Shard s = sh[r];
        const size_t nc = s.c1 - s.c0, b0 = s.c0 * chunk_len, len = nc ? (s.c1 * chunk_len < total_len ? s.c1 * chunk_len : total_len) - b0 : 0;
        Ctx &c = ctx_for(devs[r]);
        std::unique_lock<std::mutex> lk(c.mu);
        int rc = TRC_OK;
        Plan p;
        auto phase_a = [&]() -> int {
            if (!nc) return TRC_OK;
            int q;
            if ((q = c.ensure()) || (q = make_plan(codec, len, chunk_len, p))) return q;
            if ((q = c.in.need(len + 64)) || (q = c.out.need(trc_enc_bound(len, chunk_len))) || (q = c.off.need((nc + 1) * 8)) || (q = c.scratch.need(p.total + 256))) return q;
            CK(cudaMemcpyAsync(c.in.p, in + b0, len, cudaMemcpyHostToDevice, c.st));
            if (codec_static(codec)) {
                if (!cdf) return TRC_E_ARG;
                const size_t nt = n_tables(nc, cpc), t0 = cpc ? s.c0 / cpc : 0, bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
                if ((q = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return q;
                CK(cudaMemcpyAsync(c.cdf.p, cdf + t0 * CDF_STRIDE, bytes, cudaMemcpyHostToDevice, c.st));
            }
            q = trc_enc_batch_dev(codec, (const unsigned char *)c.in.p, len, chunk_len, (const cdf_t *)c.cdf.p, cdfnum, cpc, (unsigned char *)c.out.p,
                                  (uint64_t *)c.off.p, c.scratch.p, c.scratch.cap, c.st);
            if (q) return q;
            offs[r].resize(nc + 1);
            CK(cudaMemcpyAsync(offs[r].data(), c.off.p, (nc + 1) * 8, cudaMemcpyDeviceToHost, c.st));
            CK(cudaStreamSynchronize(c.st));
            size[r] = offs[r][nc];
            return TRC_OK;
        };
        rc = phase_a();
        if (rc) { rcs[r] = rc; errs[r] = g_err; }
        bar.wait();                                                   // the one exchange: shard sizes -> base offsets
        uint64_t b = 0;
        for (int k = 0; k < r; k++) b += size[k];
        bool any_err = false;
        for (int k = 0; k < n_dev; k++) any_err |= rcs[k] != TRC_OK;
        if (!any_err && nc) {
            auto phase_b = [&]() -> int {
                CK(cudaMemcpyAsync(out + b, c.out.p, (size_t)size[r], cudaMemcpyDeviceToHost, c.st));
                if (out_off) for (size_t k = 0; k < nc; k++) out_off[s.c0 + k] = b + offs[r][k];
                CK(cudaStreamSynchronize(c.st));
                return TRC_OK;
            };
            rc = phase_b();
            if (rc) { rcs[r] = rc; errs[r] = g_err; }
        }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < n_dev; r++) th.emplace_back(work, r);
    work(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < n_dev; r++) if (rcs[r]) { snprintf(g_err, sizeof g_err, "device %d: %s", devs[r], errs[r].c_str()); return rcs[r]; }
    uint64_t tot = 0;
    for (int r = 0; r < n_dev; r++) tot += size[r];
    if (out_off) out_off[n] = tot;
    if (out_len) *out_len = (size_t)tot;
    return TRC_OK;
}

static int multi_dec(int codec, const int *devs, int n_dev, const unsigned char *in, const uint64_t *in_off, unsigned char *out, size_t total_len,
                     size_t chunk_len, const cdf_t *cdf, unsigned cdfnum, size_t cpc, unsigned flags) {
    const size_t n = trc_num_chunks(total_len, chunk_len);
    std::vector<Shard> sh; make_shards(n, cpc, n_dev, sh);
    std::vector<int> rcs(n_dev, TRC_OK);
    std::vector<std::string> errs(n_dev);
    auto work = [&](int r) {
        const Shard s = sh[r];
        const size_t nc = s.c1 - s.c0;
        if (!nc) return;
        const size_t b0 = s.c0 * chunk_len, len = (s.c1 * chunk_len < total_len ? s.c1 * chunk_len : total_len) - b0;
        const uint64_t s0 = in_off[s.c0], s1 = in_off[s.c1];
        Ctx &c = ctx_for(devs[r]);
        std::lock_guard<std::mutex> lk(c.mu);
        std::vector<uint64_t> roff(nc + 1);                           // this shard's directory, relative to its first byte
        for (size_t k = 0; k <= nc; k++) roff[k] = in_off[s.c0 + k] - s0;
        auto run = [&]() -> int {
            int q;
            if ((q = c.ensure())) return q;
            if ((q = c.in.need((size_t)(s1 - s0) + 64)) || (q = c.out.need(len + 64)) || (q = c.off.need((nc + 1) * 8))) return q;
            CK(cudaMemcpyAsync(c.off.p, roff.data(), (nc + 1) * 8, cudaMemcpyHostToDevice, c.st));
            CK(cudaMemcpyAsync(c.in.p, in + s0, (size_t)(s1 - s0), cudaMemcpyHostToDevice, c.st));
            CK(cudaMemsetAsync((uint8_t *)c.in.p + (s1 - s0), 0, 64, c.st));
            if (codec_static(codec)) {
                if (!cdf) return TRC_E_ARG;
                const size_t nt = n_tables(nc, cpc), t0 = cpc ? s.c0 / cpc : 0, bytes = ((nt - 1) * CDF_STRIDE + cdfnum + 1) * sizeof(cdf_t);
                if ((q = c.cdf.need(nt * CDF_STRIDE * sizeof(cdf_t)))) return q;
                CK(cudaMemcpyAsync(c.cdf.p, cdf + t0 * CDF_STRIDE, bytes, cudaMemcpyHostToDevice, c.st));
            }
            q = trc_dec_batch_dev(codec, (const unsigned char *)c.in.p, (const uint64_t *)c.off.p, (unsigned char *)c.out.p, len, chunk_len,
                                  (const cdf_t *)c.cdf.p, cdfnum, cpc, flags, c.st);
            if (q) return q;
            CK(cudaMemcpyAsync(out + b0, c.out.p, len, cudaMemcpyDeviceToHost, c.st));
            CK(cudaStreamSynchronize(c.st));
            return TRC_OK;
        };
        const int rc = run();
        if (rc) { rcs[r] = rc; errs[r] = g_err; }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < n_dev; r++) th.emplace_back(work, r);
    work(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < n_dev; r++) if (rcs[r]) { snprintf(g_err, sizeof g_err, "device %d: %s", devs[r], errs[r].c_str()); return rcs[r]; }
    return TRC_OK;
}

[[noreturn]] void die_cuda(const char *fn, int rc) {          // mirrors die() include_/conf.h:379
    fprintf(stderr, "trc_b200: %s failed (%d): %s\n", fn, rc, g_err);
    fflush(stderr);
    exit(-1);
}

// drop-in encoder: the whole buffer is one call
size_t dropin_enc(const char *fn, int codec, unsigned char *in, size_t inlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum) {
    if (inlen == 0) {
        // reference on an empty buffer: the rANS codecs write nothing; a single range coder still flushes one
        // word (low += 2^32 -> 0x00000001, turborc_.h:120-122); the 2-coder forms dereference wild pointers.
        if (codec == RCS || codec == RC || codec == RC4 || codec == RC8) { uint32_t one = 1; memcpy(out, &one, 4); return 4; }
        return 0;
    }
    size_t l = 0;
    int rc = host_enc(g_dev, codec, in, inlen, inlen, cdf, cdfnum, 0, out, nullptr, &l);
    if (rc) die_cuda(fn, rc);
    return l;
}
// drop-in decoder.  The reference decoders are never given the compressed length; a valid stream is shorter
// than outlen (otherwise the encoder returned a raw copy and the caller must not decode, turborc.c:434), and the
// encoder's contract already makes that buffer at least outlen bytes, so outlen bytes of `in` are shipped.
size_t dropin_dec(const char *fn, int codec, unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum, unsigned flags) {
    if (outlen == 0) return 0;
    uint64_t off[2] = { 0, (uint64_t)outlen + 8 };             // != outlen, so the call is never taken for raw
    {
        Ctx &c = ctx_for(g_dev);
        std::lock_guard<std::mutex> lk(c.mu);
        int rc = c.ensure(); if (rc) die_cuda(fn, rc);
        if ((rc = c.in.need(outlen + 64))) die_cuda(fn, rc);
        if (cudaMemsetAsync((uint8_t *)c.in.p + outlen, 0, 64, c.st) != cudaSuccess) die_cuda(fn, TRC_E_CUDA);
    }
    int rc = host_dec(g_dev, codec, in, off, outlen, out, outlen, outlen, cdf, cdfnum, 0, flags);
    if (rc) die_cuda(fn, rc);
    return outlen;
}
}  // namespace

extern "C" {

static int check_devs(const int *devs, int n_dev) {
    if (!devs || n_dev < 1 || n_dev > MAX_DEV) return TRC_E_ARG;
    const int have = trc_device_count();
    for (int r = 0; r < n_dev; r++) {
        if (devs[r] < 0 || devs[r] >= have || devs[r] >= MAX_DEV) return TRC_E_ARG;
        for (int k = 0; k < r; k++) if (devs[k] == devs[r]) return TRC_E_ARG;
    }
    return TRC_OK;
}
int trc_enc_batch_host_multi(int codec, const int *devs, int n_dev, const unsigned char *in, size_t total_len, size_t chunk_len,
                             const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf,
                             unsigned char *out, uint64_t *out_off, size_t *out_len) {
    if (!in || !out || !chunk_len || !total_len || codec < 0 || codec >= NCODECS) return TRC_E_ARG;
    int rc = check_devs(devs, n_dev); if (rc) return rc;
    return multi_enc(codec, devs, n_dev, in, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, out, out_off, out_len);
}
int trc_dec_batch_host_multi(int codec, const int *devs, int n_dev, const unsigned char *in, const uint64_t *in_off,
                             unsigned char *out, size_t total_len, size_t chunk_len,
                             const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, unsigned flags) {
    if (!in || !in_off || !out || !chunk_len || !total_len || codec < 0 || codec >= NCODECS) return TRC_E_ARG;
    int rc = check_devs(devs, n_dev); if (rc) return rc;
    return multi_dec(codec, devs, n_dev, in, in_off, out, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, flags);
}

int trc_enc_batch_host(int codec, const unsigned char *in, size_t total_len, size_t chunk_len,
                       const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf,
                       unsigned char *out, uint64_t *out_off, size_t *out_len) {
    if (!in || !out) return TRC_E_ARG;
    return host_enc(g_dev, codec, in, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, out, out_off, out_len);
}
int trc_dec_batch_host(int codec, const unsigned char *in, const uint64_t *in_off,
                       unsigned char *out, size_t total_len, size_t chunk_len,
                       const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, unsigned flags) {
    if (!in || !in_off || !out || !chunk_len) return TRC_E_ARG;
    size_t n = trc_num_chunks(total_len, chunk_len);
    return host_dec(g_dev, codec, in, in_off, (size_t)in_off[n], out, total_len, chunk_len, cdf, cdfnum, chunks_per_cdf, flags);
}

// ---- self-describing container ---------------------------------------------------------
// The reference's codec calls carry no lengths, tables or cdfnum: bench() keeps them on the side, and file mode wraps each
// block in a header of { block size, inlen, clen } with clen == inlen meaning "stored" (turborc.c:665-733, 1123).  The
// container keeps that idea but gathers the per-block headers into ONE directory in front of the payload, because a GPU
// decodes every block at once and needs all offsets before the first byte.  Static tables are computed on the device
// (cdfini semantics, rccdf.c:50-68) and travel in the container.  Payload = the batch layer's packed stream, i.e. every
// chunk is still byte-for-byte one reference call.
namespace {
constexpr uint32_t CT_MAGIC = 0x42435254u;                         // "TRCB"
constexpr size_t   CT_HDR = 64;
struct CtHeader {
    uint32_t magic; uint16_t version; uint8_t codec, flags;
    uint64_t total_len, chunk_len, cdf_block, n_chunks;
    uint32_t n_tables, cdfnum;
    uint64_t payload_bytes, reserved;
};
static_assert(sizeof(CtHeader) == CT_HDR, "container header is 64 bytes");
static inline size_t al8(size_t x) { return (x + 7) & ~(size_t)7; }
struct CtLayout { size_t n, ntab, cpc, off_tabs, off_dir, off_payload; unsigned cdfnum; };

int ct_layout(int codec, size_t total_len, size_t chunk_len, size_t cdf_block, CtLayout &L) {
    if (codec < 0 || codec >= NCODECS || total_len == 0 || chunk_len == 0) return TRC_E_ARG;
    L.n = (total_len + chunk_len - 1) / chunk_len;
    L.cpc = 0; L.ntab = 0; L.cdfnum = 0;
    if (codec_static(codec)) {
        if (cdf_block && cdf_block % chunk_len) return TRC_E_ARG;  // a table covers whole chunks
        L.cpc = cdf_block ? cdf_block / chunk_len : 0;
        L.ntab = n_tables(L.n, L.cpc);
        L.cdfnum = codec == ANS4S ? 16 : 256;
    }
    L.off_tabs = CT_HDR;
    L.off_dir = L.off_tabs + al8(L.ntab * CDF_STRIDE * sizeof(cdf_t));
    L.off_payload = al16(L.off_dir + L.n * sizeof(uint32_t));
    return TRC_OK;
}
}  // namespace

size_t trc_container_bound(int codec, size_t total_len, size_t chunk_len, size_t cdf_block) {
    CtLayout L; if (ct_layout(codec, total_len, chunk_len, cdf_block, L) != TRC_OK) return 0;
    return L.off_payload + trc_enc_bound(total_len, chunk_len);
}

int trc_container_info(const unsigned char *in, size_t in_len, int *codec, size_t *total_len, size_t *chunk_len, size_t *n_chunks) {
    if (!in || in_len < CT_HDR) return TRC_E_ARG;
    CtHeader h; memcpy(&h, in, CT_HDR);
    if (h.magic != CT_MAGIC || h.version != 1 || h.codec >= NCODECS || h.total_len == 0 || h.chunk_len == 0) return TRC_E_ARG;
    // untrusted sizes: bound them by what the buffer can possibly hold before any arithmetic on them (a chunk is at least one
    // byte of payload or one directory entry; lengths are 32-bit per call)
    if (h.chunk_len >= (1ull << 31) || h.total_len > (1ull << 48) || h.n_chunks > in_len / 4 || h.cdf_block > h.total_len + h.chunk_len ||
        (h.total_len + h.chunk_len - 1) / h.chunk_len != h.n_chunks) return TRC_E_ARG;
    CtLayout L; if (ct_layout(h.codec, h.total_len, h.chunk_len, h.cdf_block, L) != TRC_OK) return TRC_E_ARG;
    if (h.n_chunks != L.n || h.n_tables != L.ntab || h.cdfnum != L.cdfnum) return TRC_E_ARG;
    if (L.off_payload > in_len || h.payload_bytes > in_len - L.off_payload) return TRC_E_ARG;
    if (codec) *codec = h.codec;
    if (total_len) *total_len = (size_t)h.total_len;
    if (chunk_len) *chunk_len = (size_t)h.chunk_len;
    if (n_chunks) *n_chunks = (size_t)h.n_chunks;
    return TRC_OK;
}

int trc_compress_host(int codec, const unsigned char *in, size_t total_len, size_t chunk_len, size_t cdf_block,
                      unsigned char *out, size_t out_cap, size_t *out_len) {
    if (!in || !out) return TRC_E_ARG;
    CtLayout L; int rc = ct_layout(codec, total_len, chunk_len, cdf_block, L); if (rc) return rc;
    if (out_cap < L.off_payload + trc_enc_bound(total_len, chunk_len)) return TRC_E_NOMEM;
    Ctx &c = ctx_for(g_dev);
    std::lock_guard<std::mutex> lk(c.mu);
    if ((rc = c.ensure())) return rc;
    Plan p; if ((rc = make_plan(codec, total_len, chunk_len, p))) return rc;
    if ((rc = c.in.need(total_len + 64)) || (rc = c.out.need(trc_enc_bound(total_len, chunk_len))) || (rc = c.off.need((L.n + 1) * 8)) ||
        (rc = c.scratch.need(p.total + 256)) || (rc = c.cdf.need((L.ntab + 1) * CDF_STRIDE * sizeof(cdf_t))) || (rc = c.status.need((L.ntab + 1) * sizeof(int)))) return rc;
    CK(cudaMemcpyAsync(c.in.p, in, total_len, cudaMemcpyHostToDevice, c.st));
    if (L.ntab) {                                                  // tables on the device, one per cdf_block bytes
        rc = trc_cdfini_batch_dev((const unsigned char *)c.in.p, total_len, cdf_block ? cdf_block : total_len, (cdf_t *)c.cdf.p, L.cdfnum, (int *)c.status.p, c.st);
        if (rc) return rc;
    }
    rc = trc_enc_batch_dev(codec, (const unsigned char *)c.in.p, total_len, chunk_len, (const cdf_t *)c.cdf.p, L.cdfnum, L.cpc,
                           (unsigned char *)c.out.p, (uint64_t *)c.off.p, c.scratch.p, c.scratch.cap, c.st);
    if (rc) return rc;
    std::vector<uint64_t> off;
    std::vector<int> status;
    try { off.resize(L.n + 1); status.resize(L.ntab); } catch (const std::bad_alloc &) { return TRC_E_NOMEM; }
    CK(cudaMemcpyAsync(off.data(), c.off.p, (L.n + 1) * 8, cudaMemcpyDeviceToHost, c.st));
    if (L.ntab) {
        CK(cudaMemcpyAsync(status.data(), c.status.p, L.ntab * sizeof(int), cudaMemcpyDeviceToHost, c.st));
        CK(cudaMemcpyAsync(out + L.off_tabs, c.cdf.p, L.ntab * CDF_STRIDE * sizeof(cdf_t), cudaMemcpyDeviceToHost, c.st));
    }
    CK(cudaStreamSynchronize(c.st));
    for (int sct : status) if (sct) { snprintf(g_err, sizeof g_err, "cdfini: degenerate table (the reference would die(), rccdf.c:65)"); return TRC_E_ARG; }
    const uint64_t payload = off[L.n];
    CK(cudaMemcpyAsync(out + L.off_payload, c.out.p, payload, cudaMemcpyDeviceToHost, c.st));
    CtHeader h; memset(&h, 0, sizeof h);
    h.magic = CT_MAGIC; h.version = 1; h.codec = (uint8_t)codec; h.total_len = total_len; h.chunk_len = chunk_len; h.cdf_block = cdf_block;
    h.n_chunks = L.n; h.n_tables = (uint32_t)L.ntab; h.cdfnum = L.cdfnum; h.payload_bytes = payload;
    memcpy(out, &h, CT_HDR);
    memset(out + L.off_tabs + L.ntab * CDF_STRIDE * sizeof(cdf_t), 0, L.off_dir - (L.off_tabs + L.ntab * CDF_STRIDE * sizeof(cdf_t)));
    uint32_t *dir = (uint32_t *)(out + L.off_dir);                 // clen per chunk; == the chunk's input length -> stored raw
    for (size_t k = 0; k < L.n; k++) dir[k] = (uint32_t)(off[k + 1] - off[k]);
    memset(out + L.off_dir + L.n * 4, 0, L.off_payload - (L.off_dir + L.n * 4));
    CK(cudaStreamSynchronize(c.st));
    if (out_len) *out_len = L.off_payload + (size_t)payload;
    return TRC_OK;
}

int trc_decompress_host(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_cap, size_t *out_len) {
    int codec; size_t total_len, chunk_len, n;
    int rc = trc_container_info(in, in_len, &codec, &total_len, &chunk_len, &n); if (rc) return rc;
    if (!out || out_cap < total_len) return TRC_E_NOMEM;
    CtHeader h; memcpy(&h, in, CT_HDR);
    CtLayout L; ct_layout(codec, total_len, chunk_len, h.cdf_block, L);
    std::vector<uint64_t> off;
    try { off.resize(n + 1); } catch (const std::bad_alloc &) { return TRC_E_NOMEM; }   // never let an exception cross the C boundary
    const unsigned char *dirp = in + L.off_dir;
    off[0] = 0;
    for (size_t k = 0; k < n; k++) {
        uint32_t cl; memcpy(&cl, dirp + 4 * k, 4);
        const size_t ilen = k + 1 < n ? chunk_len : total_len - k * chunk_len;
        if (cl > ilen + 4) return TRC_E_ARG;                        // (+4: rccdf4ienc's answer on inputs shorter than 4 bytes)
        off[k + 1] = off[k] + cl;
    }
    if (off[n] != h.payload_bytes) return TRC_E_ARG;
    Ctx &c = ctx_for(g_dev);
    std::lock_guard<std::mutex> lk(c.mu);
    if ((rc = c.ensure())) return rc;
    if ((rc = c.in.need((size_t)h.payload_bytes + 64)) || (rc = c.out.need(total_len + 64)) || (rc = c.off.need((n + 1) * 8)) ||
        (rc = c.cdf.need((L.ntab + 1) * CDF_STRIDE * sizeof(cdf_t)))) return rc;
    CK(cudaMemcpyAsync(c.off.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c.st));
    CK(cudaMemcpyAsync(c.in.p, in + L.off_payload, (size_t)h.payload_bytes, cudaMemcpyHostToDevice, c.st));
    CK(cudaMemsetAsync((uint8_t *)c.in.p + h.payload_bytes, 0, 64, c.st));
    if (L.ntab) CK(cudaMemcpyAsync(c.cdf.p, in + L.off_tabs, L.ntab * CDF_STRIDE * sizeof(cdf_t), cudaMemcpyHostToDevice, c.st));
    rc = trc_dec_batch_dev(codec, (const unsigned char *)c.in.p, (const uint64_t *)c.off.p, (unsigned char *)c.out.p, total_len, chunk_len,
                           (const cdf_t *)c.cdf.p, L.cdfnum, L.cpc, 0, c.st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c.out.p, total_len, cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));                                // `off` must outlive the upload
    if (out_len) *out_len = total_len;
    return TRC_OK;
}

// ---- drop-in layer ----------------------------------------------------------------------------------------
void anscdfini(unsigned id) { (void)id; Ctx &c = ctx_for(g_dev); std::lock_guard<std::mutex> lk(c.mu); int rc = c.ensure(); if (rc) die_cuda("anscdfini", rc); }

#define ENC3(name, codec) size_t name(unsigned char *in, size_t inlen, unsigned char *out) { return dropin_enc(#name, codec, in, inlen, out, nullptr, 0); }
#define DEC3(name, codec, fl) size_t name(unsigned char *in, size_t outlen, unsigned char *out) { return dropin_dec(#name, codec, in, outlen, out, nullptr, 0, fl); }
// static nibble rANS: the drop-in contract is a 16-symbol alphabet with a cdf_t[>=17] table (anscdf.c:57)
#define ENC4S(name) size_t name(unsigned char *in, size_t inlen, unsigned char *out, cdf_t *cdf) { return dropin_enc(#name, ANS4S, in, inlen, out, cdf, 16); }
#define DEC4S(name) size_t name(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf) { return dropin_dec(#name, ANS4S, in, outlen, out, cdf, 16, TRC_F_REF_TAIL); }
ENC4S(anscdf4senc) ENC4S(anscdf4sencs) ENC4S(anscdf4sencx)
DEC4S(anscdf4sdec) DEC4S(anscdf4sdecs) DEC4S(anscdf4sdecx)
ENC3(anscdf4enc, ANS4) ENC3(anscdf4encs, ANS4) ENC3(anscdf4encx, ANS4)
DEC3(anscdf4dec, ANS4, TRC_F_REF_TAIL) DEC3(anscdf4decs, ANS4, TRC_F_REF_TAIL) DEC3(anscdf4decx, ANS4, TRC_F_REF_TAIL)
ENC3(anscdfenc, ANS) ENC3(anscdfencs, ANS) ENC3(anscdfencx, ANS)
DEC3(anscdfdec, ANS, 0) DEC3(anscdfdecs, ANS, 0) DEC3(anscdfdecx, ANS, 0)
ENC3(anscdf1enc, ANS1) ENC3(anscdf1encs, ANS1) ENC3(anscdf1encx, ANS1)
DEC3(anscdf1dec, ANS1, 0) DEC3(anscdf1decs, ANS1, 0) DEC3(anscdf1decx, ANS1, 0)
ENC3(rccdfenc, RC) DEC3(rccdfdec, RC, 0)
ENC3(rccdfienc, RCI) DEC3(rccdfidec, RCI, 0)
ENC3(rccdf4enc, RC4) DEC3(rccdf4dec, RC4, 0)
ENC3(rccdf4ienc, RC4I) DEC3(rccdf4idec, RC4I, 0)
ENC3(rccdfenc8, RC8) DEC3(rccdfdec8, RC8, 0)
ENC3(rccdfienc8, RCI8) DEC3(rccdfidec8, RCI8, 0)
// VLC-over-CDF integer codecs (anscdf.c:139-483, rccdf.c:392-632): inlen / outlen are BYTES of 16/32-bit little-endian integers
ENC3(anscdfuenc16, ANSU16) DEC3(anscdfudec16, ANSU16, 0) ENC3(anscdfuzenc16, ANSUZ16) DEC3(anscdfuzdec16, ANSUZ16, 0)
ENC3(anscdfvenc16, ANSV16) DEC3(anscdfvdec16, ANSV16, 0) ENC3(anscdfvzenc16, ANSVZ16) DEC3(anscdfvzdec16, ANSVZ16, 0)
ENC3(anscdfvenc32, ANSV32) DEC3(anscdfvdec32, ANSV32, 0) ENC3(anscdfvzenc32, ANSVZ32) DEC3(anscdfvzdec32, ANSVZ32, 0)
ENC3(rccdfvenc16, RCV16) DEC3(rccdfvdec16, RCV16, 0) ENC3(rccdfvzenc16, RCVZ16) DEC3(rccdfvzdec16, RCVZ16, 0)
ENC3(rccdfvenc32, RCV32) DEC3(rccdfvdec32, RCV32, 0) ENC3(rccdfvzenc32, RCVZ32) DEC3(rccdfvzdec32, RCVZ32, 0)
ENC3(rccdfuenc16, RCU16) DEC3(rccdfudec16, RCU16, 0) ENC3(rccdfuenc32, RCU32) DEC3(rccdfudec32, RCU32, 0)

#define ENC5(name, codec) size_t name(unsigned char *in, size_t inlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum) { return dropin_enc(#name, codec, in, inlen, out, cdf, cdfnum); }
#define DEC5(name, codec) size_t name(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum) { return dropin_dec(#name, codec, in, outlen, out, cdf, cdfnum, 0); }
ENC5(rccdfsenc, RCS) DEC5(rccdfsbdec, RCS) DEC5(rccdfsldec, RCS)          // linear and binary search find the same symbol
ENC5(rccdfs2enc, RCS2) DEC5(rccdfsb2dec, RCS2) DEC5(rccdfsl2dec, RCS2)
DEC5(rccdfsvbdec, RCS) DEC5(rccdfsvldec, RCS)     // harness id 43: the same stream decoded by division, floor(code / range) >= cdf[x]  <=>  cdf[x] * range <= code (rccdf.c:100-122)

int cdfini(unsigned char *in, size_t inlen, cdf_t *cdf, unsigned cdfnum) {
    if (inlen == 0 || cdfnum == 0 || cdfnum > 256) { fprintf(stderr, "Fatal cdf: empty input\n"); exit(-1); }
    int status = 0;
    {
        Ctx &c = ctx_for(g_dev);
        std::lock_guard<std::mutex> lk(c.mu);
        int rc = c.ensure(); if (rc) die_cuda("cdfini", rc);
        if ((rc = c.in.need(inlen + 64)) || (rc = c.cdf.need(CDF_STRIDE * sizeof(cdf_t))) || (rc = c.status.need(64))) die_cuda("cdfini", rc);
        if (cudaMemcpyAsync(c.in.p, in, inlen, cudaMemcpyHostToDevice, c.st) != cudaSuccess) die_cuda("cdfini", TRC_E_CUDA);
        rc = trc_cdfini_batch_dev((const unsigned char *)c.in.p, inlen, inlen, (cdf_t *)c.cdf.p, cdfnum, (int *)c.status.p, c.st);
        if (rc) die_cuda("cdfini", rc);
        if (cudaMemcpyAsync(cdf, c.cdf.p, (cdfnum + 1) * sizeof(cdf_t), cudaMemcpyDeviceToHost, c.st) != cudaSuccess ||
            cudaMemcpyAsync(&status, c.status.p, sizeof(int), cudaMemcpyDeviceToHost, c.st) != cudaSuccess ||
            cudaStreamSynchronize(c.st) != cudaSuccess) die_cuda("cdfini", TRC_E_CUDA);
    }
    if (status) { fprintf(stderr, "Fatal cdf\n"); fflush(stderr); exit(-1); }     // die() rccdf.c:65-66
    return (int)inlen;
}

}  // extern "C"
