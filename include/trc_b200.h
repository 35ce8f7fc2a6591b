/*
 * trc_b200.h -- C ABI of libtrc_b200.so: the CDF entropy-coding hot path of
 * powturbo/Turbo-Range-Coder (static/adaptive-CDF rANS and range coder) on NVIDIA B200 (sm_100a).
 *
 * Two layers, both extern "C", plain pointers and sizes only:
 *
 *  1. Batch entry points (trc_*): many independent chunks per launch, one chunk == one call of the
 *     reference function named by the codec id, byte-for-byte.  This is where the GPU earns its keep:
 *     the reference coders are serial chains, the only exact parallelism is across chunks/states.
 *       - *_dev  : device pointers + a CUDA stream; nothing is copied, nothing is synchronised.
 *       - *_host : host pointers; H2D/D2H copies and a stream sync happen inside the call.
 *
 *  2. Drop-in entry points carrying the reference's own names and signatures (anscdfenc, rccdfs2enc,
 *     cdfini, ...): host pointers in, host pointers out, same return values, same raw-copy rule.
 *     Each replaces the reference symbol cited next to it.  They are thin wrappers over layer 1 with
 *     a single chunk (the whole buffer), hence latency-bound; see DESIGN.md.
 *
 * There is no CPU fallback: every entry point runs CUDA kernels and returns TRC_E_CUDA (batch layer)
 * or aborts like the reference's die() (drop-in layer, include_/conf.h:379) if no device is usable.
 */
#ifndef TRC_B200_H_
#define TRC_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned short cdf_t;                 /* reference: include/turborc.h:497 */

/* One id per reference encoder/decoder pair (SURVEY.md section 8a row in brackets). */
enum trc_codec {
    TRC_ANS4S = 0,  /* [S1/S2] anscdf4senc / anscdf4sdec  static rANS, 2 states            anscdf.c:57-85   */
    TRC_ANS4  = 1,  /* [A1]    anscdf4enc  / anscdf4dec   adaptive nibble rANS, 2 states   anscdf.c:87-133  */
    TRC_ANS   = 2,  /* [A2/A3] anscdfenc   / anscdfdec    adaptive byte rANS, 4 states     anscdf.c:567-605 */
    TRC_ANS1  = 3,  /* [A4]    anscdf1enc  / anscdf1dec   order-1 adaptive byte rANS       anscdf.c:607-645 */
    TRC_RCS   = 4,  /* [R6]    rccdfsenc   / rccdfsbdec   static range coder               rccdf.c:71-98    */
    TRC_RCS2  = 5,  /* [R7/R8] rccdfs2enc  / rccdfsb2dec  static range coder, 2 coders     rccdf.c:125-184  */
    TRC_RC    = 6,  /* [R9]    rccdfenc    / rccdfdec     adaptive byte range coder        rccdf.c:187-211  */
    TRC_RCI   = 7,  /* [R10]   rccdfienc   / rccdfidec    adaptive byte RC, 2 coders       rccdf.c:213-249  */
    TRC_RC4   = 8,  /* [R11]   rccdf4enc   / rccdf4dec    adaptive nibble RC               rccdf.c:251-278  */
    TRC_RC4I  = 9,  /* [R11]   rccdf4ienc  / rccdf4idec   adaptive nibble RC, 2 coders     rccdf.c:280-323  */
    TRC_ANSW  = 10, /* NOT a reference format: 32-way warp-interleaved static rANS, one state per lane, one stream per call
                       (the layout BASELINE's north star describes).  Parity unpinned: oracle/trc_oracle.c orc_answenc/dec is
                       the specification.  Static table like TRC_ANS4S; chunk_len must be a multiple of 4. */
    TRC_RC8   = 11, /* [f.3]   rccdfenc8   / rccdfdec8    adaptive RC over the vnibble byte code   rccdf.c:324-351  */
    TRC_RCI8  = 12, /* [f.3]   rccdfienc8  / rccdfidec8   same, 2 coders                           rccdf.c:354-389  */
    /* [f.2] VLC-over-CDF integer codecs: total_len / chunk_len are BYTES of little-endian 16- or 32-bit integers and must be
       multiples of the element size.  u = 6-bit exponent, v = 7-bit exponent, z = zigzag delta.              anscdf.c:139-483 */
    TRC_ANSU16 = 13, TRC_ANSUZ16 = 14, TRC_ANSV16 = 15, TRC_ANSVZ16 = 16, TRC_ANSV32 = 17, TRC_ANSVZ32 = 18,
    TRC_RCV16 = 19, TRC_RCVZ16 = 20, TRC_RCV32 = 21, TRC_RCVZ32 = 22, TRC_RCU16 = 23, TRC_RCU32 = 24,          /* rccdf.c:392-632 */
    TRC_NCODECS = 25
};

#define TRC_CDF_STRIDE 257                    /* entries per static table (cdf_t cdf[0x100+1], turborc.c:423) */

enum trc_status { TRC_OK = 0, TRC_E_ARG = -1, TRC_E_CUDA = -2, TRC_E_NOMEM = -3 };

/* decode flags */
#define TRC_F_REF_TAIL 1u  /* TRC_ANS4S/TRC_ANS4 only: take the (len & 3) tail symbols from decoder state 0
                              exactly as anscdf4sdec/anscdf4dec do (anscdf.c:83,104).  That is a reference bug
                              (the encoder put them on the other state) and does not round-trip; without the
                              flag the batch decoders use the state the encoder used.  Identical when every
                              chunk length is a multiple of 4. */

const char *trc_version(void);
const char *trc_last_error(void);             /* text of the last CUDA failure on this thread */
int         trc_device_count(void);
int         trc_set_device(int dev);          /* device used by *_host and drop-in calls (default 0) */
unsigned long long trc_launch_count(void);    /* kernels launched by this library so far (monotonic) */
void        trc_profile_enable(int on);       /* record CUDA events around every kernel of the batch calls */
int         trc_profile_read(float *ms, int cap); /* per-kernel milliseconds of the most recent batch call */
int         trc_selftest_host(void);          /* host-only arithmetic self-check (division-by-reciprocal table,
                                                 geometry); returns the number of failures.  Needs no GPU. */

/* ------------------------------------------------------------------------------------------------------
 * Batch layer.  The input of `total_len` bytes is cut into chunks of `chunk_len` bytes (the last one may
 * be shorter): n = trc_num_chunks(total_len, chunk_len).  Chunk c is coded exactly as one call of the
 * reference encoder on in[c*chunk_len ...]; its bytes land at out[out_off[c] .. out_off[c+1]) -- chunks are
 * packed back to back, out_off[n] is the total.  A chunk whose compressed length equals its input length
 * holds a raw copy (reference rule, include/turborc.h:47-59) and is copied back by the decoder.
 * Static codecs (TRC_ANS4S, TRC_RCS, TRC_RCS2) read `cdf`: tables of TRC_CDF_STRIDE (257) entries each, of
 * which cdfnum+1 are used (cdfnum <= 256).  chunks_per_cdf == 0 -> one table for all chunks, else chunk c
 * uses table c / chunks_per_cdf (e.g. one cdfini table per 64 MB block of 4 KiB chunks).  Other codecs
 * ignore cdf/cdfnum.
 * `out` must hold trc_enc_bound(total_len, chunk_len) bytes.
 * ------------------------------------------------------------------------------------------------------ */
size_t trc_num_chunks(size_t total_len, size_t chunk_len);
size_t trc_enc_bound(size_t total_len, size_t chunk_len);
size_t trc_enc_scratch_bytes(int codec, size_t total_len, size_t chunk_len);

int trc_enc_batch_dev(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len,
                      const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf,
                      unsigned char *d_out, uint64_t *d_out_off,
                      void *d_scratch, size_t scratch_bytes, void *cuda_stream);

int trc_dec_batch_dev(int codec, const unsigned char *d_in, const uint64_t *d_in_off,
                      unsigned char *d_out, size_t total_len, size_t chunk_len,
                      const cdf_t *d_cdf, unsigned cdfnum, size_t chunks_per_cdf,
                      unsigned flags, void *cuda_stream);

/* Prebuilt coding tables for the static codecs (TRC_ANS4S, TRC_RCS, TRC_RCS2, TRC_ANSW).  The reference harness computes its
 * cdf once per bench() and outside the timing (turborc.c:429-433); a handle does the same for everything the kernels derive
 * from the cdf (encoder / decoder entries, the 32 K-entry slot->symbol table): trc_tables_create_dev builds them once on the
 * current device from `n_tables` tables laid out TRC_CDF_STRIDE apart (asynchronously on `cuda_stream`), and the *_tab calls
 * below use them instead of rebuilding per call.  Same results as trc_enc_batch_dev / trc_dec_batch_dev byte for byte.
 * Only the aligned geometries are served (d_in / d_out 16-byte aligned, chunk_len a multiple of 16, chunks_per_cdf 0 or a
 * multiple of 128): anything else returns TRC_E_ARG -- use the plain calls there. */
typedef struct trc_tables trc_tables;
int  trc_tables_create_dev(const cdf_t *d_cdf, unsigned cdfnum, size_t n_tables, void *cuda_stream, trc_tables **out);
void trc_tables_destroy(trc_tables *t);
int trc_enc_batch_dev_tab(int codec, const unsigned char *d_in, size_t total_len, size_t chunk_len,
                          const trc_tables *tables, size_t chunks_per_cdf,
                          unsigned char *d_out, uint64_t *d_out_off,
                          void *d_scratch, size_t scratch_bytes, void *cuda_stream);
int trc_dec_batch_dev_tab(int codec, const unsigned char *d_in, const uint64_t *d_in_off,
                          unsigned char *d_out, size_t total_len, size_t chunk_len,
                          const trc_tables *tables, size_t chunks_per_cdf, unsigned flags, void *cuda_stream);

/* Host-pointer flavour: copies in, runs the device path, copies the packed result and the n+1 offsets out,
 * synchronises.  Returns TRC_OK or an error; *out_len receives out_off[n]. */
int trc_enc_batch_host(int codec, const unsigned char *in, size_t total_len, size_t chunk_len,
                       const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf,
                       unsigned char *out, uint64_t *out_off, size_t *out_len);
int trc_dec_batch_host(int codec, const unsigned char *in, const uint64_t *in_off,
                       unsigned char *out, size_t total_len, size_t chunk_len,
                       const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, unsigned flags);

/* Static tables on the device: table c = cdfini(chunk c) (reference rccdf.c:50-68 semantics), chunking as
 * above (use a large chunk_len here, e.g. the whole buffer or a 64 MB block).  d_cdf receives n tables of
 * TRC_CDF_STRIDE entries (entries past cdfnum are zero).  The call is asynchronous and returns TRC_OK once the kernels are
 * queued: whether table c is usable is reported in d_status[c] (n ints on the device, 0 = fine, -1 = degenerate table -- the
 * reference would die(), rccdf.c:65-66 -- or a byte >= cdfnum in the chunk, which the reference leaves to its caller,
 * turborc.c:535).  Pass d_status and read it before trusting the tables; NULL skips the report. */
int trc_cdfini_batch_dev(const unsigned char *d_in, size_t total_len, size_t chunk_len,
                         cdf_t *d_cdf, unsigned cdfnum, int *d_status, void *cuda_stream);

/* Several GPUs from ONE process (the reference harness calls from a single thread, turborc.c:420).
 * The batch is cut into contiguous shards of whole chunks (whole table groups when chunks_per_cdf != 0), shard r on device
 * devs[r], one host thread per device.  Encode: the devices code independently, the shard sizes are exchanged once, every
 * device downloads its packed stream to its final place in `out`: the result (bytes, offsets, length) is identical to
 * trc_enc_batch_host on one device.  Decode: the chunk directory in_off tells each device which slice of the stream to
 * fetch; every device uploads only its slice and downloads only its part of the output.  devs must be distinct. */
int trc_enc_batch_host_multi(int codec, const int *devs, int n_dev, const unsigned char *in, size_t total_len, size_t chunk_len,
                             const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf,
                             unsigned char *out, uint64_t *out_off, size_t *out_len);
int trc_dec_batch_host_multi(int codec, const int *devs, int n_dev, const unsigned char *in, const uint64_t *in_off,
                             unsigned char *out, size_t total_len, size_t chunk_len,
                             const cdf_t *cdf, unsigned cdfnum, size_t chunks_per_cdf, unsigned flags);

/* ------------------------------------------------------------------------------------------------------
 * Self-describing container.  The reference's codec calls carry no lengths, tables
 * or cdfnum -- bench() keeps them on the side (turborc.c:423-433) and file mode wraps every block in a
 * { block size, inlen, clen } header with clen == inlen meaning "stored" (turborc.c:665-733, 1123).  The
 * container gathers those per-block headers into one directory in front of the payload (a GPU decodes all
 * blocks at once), computes the static tables on the device (cdfini, rccdf.c:50-68) and ships them along:
 *
 *   [64-byte header: "TRCB", version, codec, total_len, chunk_len, cdf_block, n_chunks, n_tables, cdfnum, payload_bytes]
 *   [n_tables x 257 cdf_t]  [n_chunks x u32 compressed length (== input length: raw copy)]  [payload]
 *
 * payload == the packed stream of trc_enc_batch_*: chunk c is byte-for-byte one call of the codec's reference
 * encoder.  cdf_block: bytes of input per static table (0 = one table for the buffer; must be a multiple of
 * chunk_len); cdfnum is 256 (16 for TRC_ANS4S).  This format is this library's own: the reference has no
 * counterpart at this level, so it is round-trip tested, and its payload is checked against the oracle.
 * ------------------------------------------------------------------------------------------------------ */
size_t trc_container_bound(int codec, size_t total_len, size_t chunk_len, size_t cdf_block);
int trc_compress_host(int codec, const unsigned char *in, size_t total_len, size_t chunk_len, size_t cdf_block,
                      unsigned char *out, size_t out_cap, size_t *out_len);
int trc_decompress_host(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_cap, size_t *out_len);
/* host-only header check; any of the out pointers may be NULL */
int trc_container_info(const unsigned char *in, size_t in_len, int *codec, size_t *total_len, size_t *chunk_len, size_t *n_chunks);

/* Multi-GPU gather over peer memory (one process per GPU): the destination rank allocates a buffer, exports a
 * CUDA-IPC handle, the other ranks open it and push their packed streams straight into it with a kernel whose byte
 * count is read from device memory (the encoder's out_off[n]) -- no host round trip, no collective call. */
int trc_dev_alloc(void **p, size_t bytes);
int trc_dev_free(void *p);
int trc_ipc_export(void *p, unsigned char *handle64);
int trc_ipc_open(const unsigned char *handle64, void **p);
int trc_ipc_close(void *p);
int trc_memcpy_dev(void *dst, const void *src, size_t bytes, void *cuda_stream);
int trc_push_dev(void *dst, const void *src, const uint64_t *d_len, size_t fixed_len, size_t cap, uint64_t *dst_len,
                 uint64_t *dst_flag, uint64_t seq, unsigned int *d_counter, const uint64_t *ack, uint64_t ack_need, size_t skip, void *cuda_stream);
int trc_ack_dev(uint64_t *ack, uint64_t seq, void *cuda_stream);
int trc_wait_flags_dev(const uint64_t *flags, const uint64_t *lens, unsigned n, uint64_t seq, unsigned int *d_status, void *cuda_stream);

/* ------------------------------------------------------------------------------------------------------
 * Drop-in layer: the reference's names, signatures and return values.
 * ------------------------------------------------------------------------------------------------------ */
void   anscdfini(unsigned id);                                                          /* anscdf.c:759  */
size_t anscdf4senc(unsigned char *in, size_t inlen,  unsigned char *out, cdf_t *cdf);   /* anscdf.c:814  */
size_t anscdf4sdec(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf);   /* anscdf.c:815  */
size_t anscdf4enc (unsigned char *in, size_t inlen,  unsigned char *out);               /* anscdf.c:817  */
size_t anscdf4dec (unsigned char *in, size_t outlen, unsigned char *out);               /* anscdf.c:818  */
size_t anscdfenc  (unsigned char *in, size_t inlen,  unsigned char *out);               /* anscdf.c:820  */
size_t anscdfdec  (unsigned char *in, size_t outlen, unsigned char *out);               /* anscdf.c:821  */
size_t anscdf1enc (unsigned char *in, size_t inlen,  unsigned char *out);               /* anscdf.c:822  */
size_t anscdf1dec (unsigned char *in, size_t outlen, unsigned char *out);               /* anscdf.c:823  */
/* the harness also names the per-ISA variants directly (turborc.c:516-521); one GPU path serves all */
size_t anscdf4sencs(unsigned char *in, size_t inlen,  unsigned char *out, cdf_t *cdf);
size_t anscdf4sdecs(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf);
size_t anscdf4sencx(unsigned char *in, size_t inlen,  unsigned char *out, cdf_t *cdf);
size_t anscdf4sdecx(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf);
size_t anscdf4encs(unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdf4decs(unsigned char *in, size_t outlen, unsigned char *out);
size_t anscdf4encx(unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdf4decx(unsigned char *in, size_t outlen, unsigned char *out);
size_t anscdfencs (unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdfdecs (unsigned char *in, size_t outlen, unsigned char *out);
size_t anscdfencx (unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdfdecx (unsigned char *in, size_t outlen, unsigned char *out);
size_t anscdf1encs(unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdf1decs(unsigned char *in, size_t outlen, unsigned char *out);
size_t anscdf1encx(unsigned char *in, size_t inlen,  unsigned char *out);
size_t anscdf1decx(unsigned char *in, size_t outlen, unsigned char *out);

int    cdfini(unsigned char *in, size_t inlen, cdf_t *cdf, unsigned cdfnum);                               /* rccdf.c:50  */
size_t rccdfsenc  (unsigned char *in, size_t inlen,  unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:71  */
size_t rccdfsbdec (unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:92  */
size_t rccdfsldec (unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:84  */
size_t rccdfsvbdec(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:100 (same stream, search by division: same symbols) */
size_t rccdfsvldec(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:112 */
size_t rccdfs2enc (unsigned char *in, size_t inlen,  unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:125 */
size_t rccdfsb2dec(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:166 */
size_t rccdfsl2dec(unsigned char *in, size_t outlen, unsigned char *out, cdf_t *cdf, unsigned cdfnum);     /* rccdf.c:146 */
size_t rccdfenc   (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:201 */
size_t rccdfdec   (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:187 */
size_t rccdfienc  (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:230 */
size_t rccdfidec  (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:213 */
size_t rccdf4enc  (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:267 */
size_t rccdf4dec  (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:251 */
size_t rccdf4ienc (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:302 */
size_t rccdf4idec (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:280 */
size_t rccdfenc8  (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:341 */
size_t rccdfdec8  (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:324 */
size_t rccdfienc8 (unsigned char *in, size_t inlen,  unsigned char *out);                                  /* rccdf.c:371 */
size_t rccdfidec8 (unsigned char *in, size_t outlen, unsigned char *out);                                  /* rccdf.c:354 */
/* VLC-over-CDF integer codecs: lengths in bytes (multiples of the element size) */
size_t anscdfuenc16 (unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfudec16 (unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:139,168 */
size_t anscdfuzenc16(unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfuzdec16(unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:195,226 */
size_t anscdfvenc16 (unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfvdec16 (unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:255,284 */
size_t anscdfvzenc16(unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfvzdec16(unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:311,342 */
size_t anscdfvenc32 (unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfvdec32 (unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:369,398 */
size_t anscdfvzenc32(unsigned char *in, size_t inlen, unsigned char *out);  size_t anscdfvzdec32(unsigned char *in, size_t outlen, unsigned char *out);   /* anscdf.c:425,456 */
size_t rccdfvenc16  (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfvdec16  (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:392,413 */
size_t rccdfvzenc16 (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfvzdec16 (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:432,454 */
size_t rccdfvenc32  (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfvdec32  (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:473,495 */
size_t rccdfvzenc32 (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfvzdec32 (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:515,537 */
size_t rccdfuenc16  (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfudec16  (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:555,576 */
size_t rccdfuenc32  (unsigned char *in, size_t inlen, unsigned char *out);  size_t rccdfudec32  (unsigned char *in, size_t outlen, unsigned char *out);   /* rccdf.c:595,616 */

#ifdef __cplusplus
}
#endif
#endif /* TRC_B200_H_ */
