/*
 * dropin_demo.c -- a miniature of the reference harness's bench() (turborc.c:420-577): same call sequence,
 * same symbols, host malloc'ed buffers of OSIZE(n) = n*4/3 bytes.  It is compiled twice from this one source:
 *   dropin_demo_ref  linked against oracle/_ref/libtrcref.so   (the unmodified reference)
 *   dropin_demo_gpu  linked against libtrc_b200.so              (this repository)
 * and tests/test_dropin_link.py checks that both write identical bytes.  That is the drop-in claim at link
 * level: no source change in the caller, only the library behind the symbols.
 *
 *   usage: dropin_demo <id> <infile> <outfile>      id = 42 45 46 47 56 64 65 (turborc -e ids, turborc.c:495-536)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned short cdf_t;
int    cdfini(unsigned char *in, size_t inlen, cdf_t *cdf, unsigned cdfnum);
size_t rccdfsenc(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned), rccdfsbdec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
size_t rccdfsldec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
size_t rccdfsvbdec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned), rccdfsvldec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
size_t rccdfs2enc(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned), rccdfsb2dec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
size_t rccdfsl2dec(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
size_t rccdfenc(unsigned char *, size_t, unsigned char *), rccdfdec(unsigned char *, size_t, unsigned char *);
size_t rccdf4enc(unsigned char *, size_t, unsigned char *), rccdf4dec(unsigned char *, size_t, unsigned char *);
size_t rccdfienc(unsigned char *, size_t, unsigned char *), rccdfidec(unsigned char *, size_t, unsigned char *);
size_t rccdf4ienc(unsigned char *, size_t, unsigned char *), rccdf4idec(unsigned char *, size_t, unsigned char *);
size_t rccdfenc8(unsigned char *, size_t, unsigned char *), rccdfdec8(unsigned char *, size_t, unsigned char *);
size_t rccdfienc8(unsigned char *, size_t, unsigned char *), rccdfidec8(unsigned char *, size_t, unsigned char *);
size_t anscdfenc(unsigned char *, size_t, unsigned char *), anscdfdec(unsigned char *, size_t, unsigned char *);
size_t anscdf4enc(unsigned char *, size_t, unsigned char *), anscdf4dec(unsigned char *, size_t, unsigned char *);
size_t anscdf1enc(unsigned char *, size_t, unsigned char *), anscdf1dec(unsigned char *, size_t, unsigned char *);
size_t anscdf4senc(unsigned char *, size_t, unsigned char *, cdf_t *), anscdf4sdec(unsigned char *, size_t, unsigned char *, cdf_t *);

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s id infile outfile\n", argv[0]); return 2; }
    int id = atoi(argv[1]);
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    fseek(f, 0, SEEK_END); size_t n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    /* one allocation, in below out: the reference's anscdf4senc needs out above in (anscdf.c:63) */
    size_t on = n * 4 / 3 + 1024;
    unsigned char *in = malloc(n + 64 + 2 * on), *out = in + n + 64, *cpy = out + on;
    if (fread(in, 1, n, f) != n) return 2;
    fclose(f);
    unsigned m = 0;
    for (size_t i = 0; i < n; i++) if (in[i] > m) m = in[i];
    for (size_t i = 0; i < n; i++) cpy[i] = (unsigned char)~in[i];          /* memrcpy, turborc.c:427 */
    cdf_t cdf[0x100 + 1];
    cdfini(in, n, cdf, 0x100);                                               /* turborc.c:432 */
    size_t l = 0;
#define CCPY(dec) (l == n ? (size_t)memcpy(cpy, out, n) : (dec))
    switch (id) {
    case 42: l = rccdfsenc(in, n, out, cdf, m + 1);  CCPY(m < 16 ? rccdfsldec(out, n, cpy, cdf, m + 1) : rccdfsbdec(out, n, cpy, cdf, m + 1)); break;
    case 43: l = rccdfsenc(in, n, out, cdf, m + 1);  CCPY(m < 16 ? rccdfsvldec(out, n, cpy, cdf, m + 1) : rccdfsvbdec(out, n, cpy, cdf, m + 1)); break;   /* turborc.c:496 */
    case 45: l = rccdfs2enc(in, n, out, cdf, m + 1); CCPY(m < 16 ? rccdfsl2dec(out, n, cpy, cdf, m + 1) : rccdfsb2dec(out, n, cpy, cdf, m + 1)); break;
    case 46: if (m < 16) { l = rccdf4enc(in, n, out); CCPY(rccdf4dec(out, n, cpy)); } else { l = rccdfenc(in, n, out); CCPY(rccdfdec(out, n, cpy)); } break;
    case 47: if (m < 16) { l = rccdf4ienc(in, n, out); CCPY(rccdf4idec(out, n, cpy)); } else { l = rccdfienc(in, n, out); CCPY(rccdfidec(out, n, cpy)); } break;
    case 48: l = rccdfenc8(in, n, out);  CCPY(rccdfdec8(out, n, cpy)); break;                                      /* turborc.c:503 */
    case 49: l = rccdfienc8(in, n, out); CCPY(rccdfidec8(out, n, cpy)); break;                                     /* turborc.c:504 */
    case 56: if (m < 16) { l = anscdf4enc(in, n, out); CCPY(anscdf4dec(out, n, cpy)); } else { l = anscdfenc(in, n, out); CCPY(anscdfdec(out, n, cpy)); } break;
    case 64: l = anscdf1enc(in, n, out); CCPY(anscdf1dec(out, n, cpy)); break;
    case 65: if (m < 16) { l = anscdf4senc(in, n, out, cdf); CCPY(anscdf4sdec(out, n, cpy, cdf)); } break;
    default: fprintf(stderr, "unknown id %d\n", id); return 2;
    }
    int ok = memcmp(in, cpy, n) == 0;                                        /* memcheck, turborc.c:577 */
    f = fopen(argv[3], "wb");
    fwrite(&l, sizeof l, 1, f); fwrite(out, 1, l, f); fwrite(cdf, sizeof cdf[0], 257, f);
    fclose(f);
    printf("id %d  n %zu  l %zu  %.2f%%  roundtrip %s\n", id, n, l, n ? 100.0 * l / n : 0.0, ok ? "ok" : "ERROR");
    free(in);
    return 0;
}
