/*
 * cpu_bench.c -- multi-threaded timing driver for the CPU checkers.  TEST/BENCH INFRASTRUCTURE ONLY
 * (bench.py's cpu_baseline leg and `bench.py --impl reference`).
 *
 * dlopen()s either oracle/_ref/libtrcref.so (the unmodified reference, symbol prefix "") or
 * oracle/libtrc_oracle.so (our port, prefix "orc_"), cuts the input into chunks, and lets `threads`
 * pthreads each run encoder and decoder over their share of the chunks -- the same "one reference call per
 * chunk" semantics the GPU batch API implements.  Reports wall seconds of the encode phase and of the decode
 * phase (all threads, barrier to barrier), the total compressed size, and whether every chunk round-tripped.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef size_t (*fn3)(unsigned char *, size_t, unsigned char *);
typedef size_t (*fn4)(unsigned char *, size_t, unsigned char *, uint16_t *);
typedef size_t (*fn5)(unsigned char *, size_t, unsigned char *, uint16_t *, unsigned);

typedef struct {
    int tid, threads, sig, reps;
    void *enc, *dec;
    unsigned char *in, *out, *cpy;
    size_t n, chunk, nchunks, slot;
    size_t *clen;
    uint16_t *cdf; unsigned cdfnum;
    pthread_barrier_t *bar;
    int bad;
} job;

static size_t call(void *f, int sig, unsigned char *a, size_t n, unsigned char *b, uint16_t *cdf, unsigned cdfnum) {
    if (sig == 0) return ((fn3)f)(a, n, b);
    if (sig == 1) return ((fn4)f)(a, n, b, cdf);
    return ((fn5)f)(a, n, b, cdf, cdfnum);
}

static void *worker(void *p) {
    job *j = (job *)p;
    for (int r = 0; r < j->reps; r++) {
        pthread_barrier_wait(j->bar);                       /* encode phase start */
        for (size_t c = j->tid; c < j->nchunks; c += j->threads) {
            size_t s = c * j->chunk, l = j->n - s < j->chunk ? j->n - s : j->chunk;
            j->clen[c] = call(j->enc, j->sig, j->in + s, l, j->out + c * j->slot, j->cdf, j->cdfnum);
        }
        pthread_barrier_wait(j->bar);                       /* encode end / decode start */
        for (size_t c = j->tid; c < j->nchunks; c += j->threads) {
            size_t s = c * j->chunk, l = j->n - s < j->chunk ? j->n - s : j->chunk;
            if (j->clen[c] == l) memcpy(j->cpy + s, j->out + c * j->slot, l);      /* CCPY, turborc.c:434 */
            else call(j->dec, j->sig, j->out + c * j->slot, l, j->cpy + s, j->cdf, j->cdfnum);
        }
        pthread_barrier_wait(j->bar);                       /* decode end */
    }
    return 0;
}

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* returns 0 on success; enc_s/dec_s = best-of-reps wall seconds of each phase */
int orc_cpu_bench(const char *libpath, const char *encname, const char *decname, int sig,
                  const unsigned char *in, size_t n, size_t chunk, const uint16_t *cdf, unsigned cdfnum,
                  int threads, int reps, double *enc_s, double *dec_s, size_t *total_clen, int *roundtrip_ok) {
    void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    void *enc = dlsym(h, encname), *dec = dlsym(h, decname);
    if (!enc || !dec) return -2;
    void (*ini)(unsigned) = (void (*)(unsigned))dlsym(h, "anscdfini");
    if (ini) ini(0);                                        /* the reference's lazy ISA dispatch is not thread safe */
    size_t nchunks = (n + chunk - 1) / chunk, slot = chunk + chunk / 3 + 320;   /* OSIZE = 4/3 n (turborc.c:418) + slack */
    /* one allocation, input first: the reference's anscdf4senc needs out above in (anscdf.c:63) */
    unsigned char *buf = (unsigned char *)malloc(n + 64 + nchunks * slot + n + 64);
    size_t *clen = (size_t *)calloc(nchunks, sizeof(size_t));
    if (!buf || !clen) return -3;
    unsigned char *cin = buf, *out = buf + n + 64, *cpy = out + nchunks * slot;
    memcpy(cin, in, n);
    memset(cpy, 0xA5, n);
    uint16_t tab[257] = { 0 };
    if (cdf) memcpy(tab, cdf, sizeof tab);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, 0, threads + 1);
    pthread_t *th = (pthread_t *)malloc(threads * sizeof *th);
    job *jobs = (job *)calloc(threads, sizeof *jobs);
    for (int t = 0; t < threads; t++) {
        job J = { t, threads, sig, reps, enc, dec, cin, out, cpy, n, chunk, nchunks, slot, clen, tab, cdfnum, &bar, 0 };
        jobs[t] = J;
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    double be = 1e30, bd = 1e30;
    for (int r = 0; r < reps; r++) {
        pthread_barrier_wait(&bar); double t0 = now();
        pthread_barrier_wait(&bar); double t1 = now();
        pthread_barrier_wait(&bar); double t2 = now();
        if (t1 - t0 < be) be = t1 - t0;
        if (t2 - t1 < bd) bd = t2 - t1;
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    size_t tot = 0;
    for (size_t c = 0; c < nchunks; c++) tot += clen[c];
    *enc_s = be; *dec_s = bd; *total_clen = tot; *roundtrip_ok = memcmp(cin, cpy, n) == 0;
    free(th); free(jobs); free(clen); free(buf);
    pthread_barrier_destroy(&bar);
    return 0;
}

/*
 * orc_batch_enc -- the batch semantics as the CHECKER states them: the reference encoder (or the port) called once per
 * chunk, results packed back to back, offsets in off[0..nchunks].  Multi-threaded so that full-size (100 MB) parity checks
 * and bench.py's correctness gate finish in a second.  chunks_per_cdf > 0: chunk c uses table c / chunks_per_cdf of `cdf`
 * (257 entries apart).  Returns 0, or <0 on failure (-4: packed_cap too small).
 */
typedef struct {
    int tid, threads, sig;
    void *enc;
    unsigned char *in, *slots;
    size_t n, chunk, nchunks, slot, cpc;
    size_t *clen;
    const uint16_t *cdf; unsigned cdfnum;
} bjob;

static void *bworker(void *p) {
    bjob *j = (bjob *)p;
    for (size_t c = j->tid; c < j->nchunks; c += j->threads) {
        size_t s = c * j->chunk, l = j->n - s < j->chunk ? j->n - s : j->chunk;
        uint16_t tab[257];
        if (j->cdf) memcpy(tab, j->cdf + (j->cpc ? (c / j->cpc) * 257 : 0), sizeof tab);
        j->clen[c] = call(j->enc, j->sig, j->in + s, l, j->slots + c * j->slot, j->cdf ? tab : 0, j->cdfnum);
    }
    return 0;
}

int orc_batch_enc(const char *libpath, const char *encname, int sig, const unsigned char *in, size_t n, size_t chunk,
                  const uint16_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, int threads,
                  unsigned char *packed, size_t packed_cap, uint64_t *off) {
    void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    void *enc = dlsym(h, encname);
    if (!enc) return -2;
    void (*ini)(unsigned) = (void (*)(unsigned))dlsym(h, "anscdfini");
    if (ini) ini(0);
    if (threads < 1) threads = 1;
    size_t nchunks = (n + chunk - 1) / chunk, slot = chunk + chunk / 3 + 320;
    unsigned char *buf = (unsigned char *)malloc(n + 64 + nchunks * slot);      /* input first: out above in (anscdf.c:63) */
    size_t *clen = (size_t *)calloc(nchunks ? nchunks : 1, sizeof(size_t));
    if (!buf || !clen) return -3;
    memcpy(buf, in, n);
    pthread_t *th = (pthread_t *)malloc(threads * sizeof *th);
    bjob *jobs = (bjob *)calloc(threads, sizeof *jobs);
    for (int t = 0; t < threads; t++) {
        bjob J = { t, threads, sig, enc, buf, buf + n + 64, n, chunk, nchunks, slot, chunks_per_cdf, clen, cdf, cdfnum };
        jobs[t] = J;
        pthread_create(&th[t], 0, bworker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    int rc = 0;
    size_t o = 0;
    for (size_t c = 0; c < nchunks && !rc; c++) {
        size_t s = c * chunk, l = n - s < chunk ? n - s : chunk;
        size_t r = clen[c], have = r < l ? r : l;               /* rccdf4ienc on < 4 bytes answers 4: bytes past the raw copy are 0 */
        off[c] = o;
        if (o + r > packed_cap) { rc = -4; break; }
        memcpy(packed + o, buf + n + 64 + c * slot, have);
        if (r > have) memset(packed + o + have, 0, r - have);
        o += r;
    }
    off[nchunks] = o;
    free(th); free(jobs); free(clen); free(buf);
    return rc;
}
