/* xturborc.c -- included by turborc.c:573 INSIDE bench()'s switch(id) (build with -D_EXT); see xturborc.h.
 * in, n, out, cpy, cdf, m, l are bench()'s own variables; TM is the reference's timing macro. */
    #undef  ID_LAST
    #define ID_LAST 99
    /* bench() prepares m (largest symbol) and the cdfini table only for ids 40-65 (turborc.c:429-433) */
    #define XTRC_CDF() do { for(m = i = 0; i < n; i++) if(in[i] > m) m = in[i]; cdfini(in, n, cdf, 0x100); } while(0)
    case 90: { size_t ck = xtrc_chunk(4096); XTRC_CDF();
      TM("90:gpu cdfsb  static interlv, batch      ", l = xtrc_enc(TRC_RCS2, in, n, ck, out, cdf, m+1), n, l, xtrc_dec(TRC_RCS2, out, n, ck, cpy, cdf, m+1)); } break;
    case 91: { size_t ck = xtrc_chunk(4096); XTRC_CDF();
      TM("91:gpu cdfsb  static, batch              ", l = xtrc_enc(TRC_RCS,  in, n, ck, out, cdf, m+1), n, l, xtrc_dec(TRC_RCS,  out, n, ck, cpy, cdf, m+1)); } break;
    case 92: { size_t ck = xtrc_chunk(65536);
      TM("92:gpu ans    byte adaptive, batch       ", l = xtrc_enc(TRC_ANS,  in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_ANS,  out, n, ck, cpy, NULL, 0)); } break;
    case 93: { size_t ck = xtrc_chunk(65536);
      TM("93:gpu cdf    byte adaptive, batch       ", l = xtrc_enc(TRC_RC,   in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_RC,   out, n, ck, cpy, NULL, 0)); } break;
    case 94: { size_t ck = xtrc_chunk(4194304);
      TM("94:gpu ans    o1, batch                  ", l = xtrc_enc(TRC_ANS1, in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_ANS1, out, n, ck, cpy, NULL, 0)); } break;
    case 95: { size_t ck = xtrc_chunk(65536);
      TM("95:gpu cdfi   byte adaptive interlv,batch", l = xtrc_enc(TRC_RCI,  in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_RCI,  out, n, ck, cpy, NULL, 0)); } break;
    case 98: { size_t ck = xtrc_chunk(4096);
      TM("98:gpu cdf-8  vnibble, batch             ", l = xtrc_enc(TRC_RC8,  in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_RC8,  out, n, ck, cpy, NULL, 0)); } break;
    case 99: { size_t ck = xtrc_chunk(4096);
      TM("99:gpu cdfi-8 vnibble interleaved, batch ", l = xtrc_enc(TRC_RCI8, in, n, ck, out, NULL, 0),  n, l, xtrc_dec(TRC_RCI8, out, n, ck, cpy, NULL, 0)); } break;
    /* VLC-over-CDF integer codecs (ids 50-53, 60-63 of the reference, turborc.c:505-510,526-533): z = bytes per integer (-Os2 / -Os4) */
    #define XTRC_VLC(_id_, _name_, _c16_, _c32_) case _id_: { size_t ck = xtrc_chunk(4096) & ~(size_t)3; int cd = z == 2 ? (_c16_) : z == 4 ? (_c32_) : -1; \
      if(cd < 0 || n % z) break; \
      TM(_name_, l = xtrc_enc(cd, in, n, ck, out, NULL, 0), n, l, xtrc_dec(cd, out, n, ck, cpy, NULL, 0)); } break
    XTRC_VLC(70, "70:gpu cdf    Turbo vlc6, batch          ", TRC_RCU16,   TRC_RCU32);
    XTRC_VLC(72, "72:gpu cdf    Turbo vlc7, batch          ", TRC_RCV16,   TRC_RCV32);
    XTRC_VLC(73, "73:gpu cdf    Turbo vlc7 zigzag, batch   ", TRC_RCVZ16,  TRC_RCVZ32);
    XTRC_VLC(74, "74:gpu anscdf Turbo vlc6, batch          ", TRC_ANSU16,  -1);
    XTRC_VLC(75, "75:gpu anscdf Turbo vlc6 zigzag, batch   ", TRC_ANSUZ16, -1);
    XTRC_VLC(76, "76:gpu anscdf Turbo vlc7, batch          ", TRC_ANSV16,  TRC_ANSV32);
    XTRC_VLC(77, "77:gpu anscdf Turbo vlc7 zigzag, batch   ", TRC_ANSVZ16, TRC_ANSVZ32);
    case 96: { XTRC_CDF(); xtrc_f5 e = (xtrc_f5)xtrc_sym("rccdfs2enc"), d = (xtrc_f5)xtrc_sym("rccdfsb2dec");
      TM("96:gpu cdfsb  static interlv, drop-in    ", l = e(in, n, out, cdf, m+1), n, l, CCPY:d(out, n, cpy, cdf, m+1)); } break;
    case 97: { xtrc_f3 e = (xtrc_f3)xtrc_sym("anscdfenc"), d = (xtrc_f3)xtrc_sym("anscdfdec");
      TM("97:gpu ans    byte adaptive, drop-in     ", l = e(in, n, out), n, l, CCPY:d(out, n, cpy)); } break;
