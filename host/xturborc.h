/* xturborc.h -- GPU rows for the reference's own benchmark harness.
 *
 * turborc.c:65-67 includes this file when the harness is built with -D_EXT (makefile:257-258, `make EXT=1`), and
 * turborc.c:572-574 includes xturborc.c inside bench()'s switch(id).  Together they add the ids below, which run
 * libtrc_b200.so on the SAME in/out/cpy buffers, cdf table and timing macro (TM, include_/time_.h:199) as the CPU
 * ids 42-65 -- turborc.c itself is not edited.  `turborc -e45,90 file` prints the CPU row and the GPU row side by side.
 *
 *   id   GPU codec (batch of XTRC_CHUNK-byte reference calls)          CPU id it reproduces per chunk
 *   90   TRC_RCS2  rccdfs2enc / rccdfsb2dec, 4 KiB chunks               45
 *   91   TRC_RCS   rccdfsenc  / rccdfsbdec,  4 KiB chunks               42
 *   92   TRC_ANS   anscdfenc  / anscdfdec,   64 KiB chunks              56
 *   93   TRC_RC    rccdfenc   / rccdfdec,    64 KiB chunks              46
 *   94   TRC_ANS1  anscdf1enc / anscdf1dec,  4 MiB chunks               64
 *   95   TRC_RCI   rccdfienc  / rccdfidec,   64 KiB chunks              47
 *   96   drop-in symbol rccdfs2enc / rccdfsb2dec of libtrc_b200.so (whole buffer == one call; same bytes as id 45)
 *   97   drop-in symbol anscdfenc / anscdfdec    of libtrc_b200.so (whole buffer; same bytes as id 56)
 *   98   TRC_RC8   rccdfenc8  / rccdfdec8,   4 KiB chunks               48
 *   99   TRC_RCI8  rccdfienc8 / rccdfidec8,  4 KiB chunks               49
 *   70   TRC_RCU16/32   rccdfuenc  / rccdfudec   (Turbo vlc6), 4 KiB chunks    50      (-Os2 / -Os4 select the integer width, like
 *   72   TRC_RCV16/32   rccdfvenc  / rccdfvdec   (Turbo vlc7)                  52       the reference ids; input length must be whole
 *   73   TRC_RCVZ16/32  rccdfvzenc / rccdfvzdec  (vlc7 zigzag)                 53       integers)
 *   74   TRC_ANSU16     anscdfuenc16  / anscdfudec16                           60
 *   75   TRC_ANSUZ16    anscdfuzenc16 / anscdfuzdec16                          61
 *   76   TRC_ANSV16/32  anscdfvenc / anscdfvdec                                62
 *   77   TRC_ANSVZ16/32 anscdfvzenc / anscdfvzdec                              63
 * Chunk sizes can be overridden with the environment variable TRC_CHUNK (bytes).
 */
#ifndef XTURBORC_H_
#define XTURBORC_H_
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include "trc_b200.h"

static uint64_t *xtrc_off;
static size_t    xtrc_offn;

static size_t xtrc_chunk(size_t dflt) { const char *e = getenv("TRC_CHUNK"); size_t v = e ? strtoull(e, NULL, 10) : 0; return v ? v : dflt; }

static uint64_t *xtrc_offs(size_t n, size_t chunk) {
  size_t nc = trc_num_chunks(n, chunk) + 1;
  if(nc > xtrc_offn) { xtrc_off = (uint64_t *)realloc(xtrc_off, nc * sizeof(uint64_t)); xtrc_offn = nc; if(!xtrc_off) { fprintf(stderr, "xturborc: out of memory\n"); exit(-1); } }
  return xtrc_off;
}
static void xtrc_die(const char *what, int rc) { fprintf(stderr, "xturborc: %s failed (%d): %s\n", what, rc, trc_last_error()); exit(-1); }

static size_t xtrc_enc(int codec, unsigned char *in, size_t n, size_t chunk, unsigned char *out, cdf_t *cdf, unsigned cdfnum) {
  size_t ol = 0;
  int rc = trc_enc_batch_host(codec, in, n, chunk, cdf, cdfnum, 0, out, xtrc_offs(n, chunk), &ol);
  if(rc) xtrc_die("trc_enc_batch_host", rc);
  return ol;
}
static size_t xtrc_dec(int codec, unsigned char *in, size_t n, size_t chunk, unsigned char *out, cdf_t *cdf, unsigned cdfnum) {
  int rc = trc_dec_batch_host(codec, in, xtrc_offs(n, chunk), out, n, chunk, cdf, cdfnum, 0, 0);
  if(rc) xtrc_die("trc_dec_batch_host", rc);
  return n;
}

/* the library's drop-in symbols carry the reference's names; inside this binary those names are bound to the reference's
   own objects, so the library's versions are looked up on its handle */
typedef size_t (*xtrc_f5)(unsigned char *, size_t, unsigned char *, cdf_t *, unsigned);
typedef size_t (*xtrc_f3)(unsigned char *, size_t, unsigned char *);
static void *xtrc_sym(const char *name) {
  static void *h;
  if(!h && !(h = dlopen("libtrc_b200.so", RTLD_NOW | RTLD_LOCAL))) { fprintf(stderr, "xturborc: %s\n", dlerror()); exit(-1); }
  void *f = dlsym(h, name);
  if(!f) { fprintf(stderr, "xturborc: %s\n", dlerror()); exit(-1); }
  return f;
}
#endif
