/*
 *  TEST INFRASTRUCTURE -- not product code, and not reference code.
 *
 *  Link-time stand-ins for the htslib entry points that the reference's
 *  utility/src/sequence/htsSeqFile-v1.C names.  oracle/build_ref.sh links the
 *  reference tools against these instead of building the vendored htslib
 *  (which wants bz2/lzma/curl headers this image does not have).  The parity
 *  harness only ever feeds FASTA to sqStoreCreate, which the reference opens
 *  with bufSeqFile before it would try htsSeqFile (dnaSeqFile-v1.H:255-272),
 *  so none of these is reached; each one fails the "open" cleanly if it is.
 */
#include <stddef.h>

void *hts_open(const char *fn, const char *mode)        { (void)fn; (void)mode; return NULL; }
int   hts_close(void *fp)                                { (void)fp; return 0; }
char *hts_format_description(const void *format)         { (void)format; return NULL; }
void *sam_hdr_read(void *fp)                             { (void)fp; return NULL; }
void  sam_hdr_destroy(void *h)                           { (void)h; }
void *bam_init1(void)                                    { return NULL; }
void  bam_destroy1(void *b)                              { (void)b; }
int   sam_read1(void *fp, void *h, void *b)              { (void)fp; (void)h; (void)b; return -1; }
char *bam_flag2str(int flag)                             { (void)flag; return NULL; }

/*  Named by stores/tgTig.C (BAM output of tig layouts), which ovStoreDump links for its optional -bogart filter;
 *  dumping an overlap store never reaches them.  */
int   bam_set1(void *b, ...)                             { (void)b; return -1; }
int   sam_hdr_add_line(void *h, const char *t, ...)      { (void)h; (void)t; return -1; }
int   sam_hdr_add_pg(void *h, const char *n, ...)        { (void)h; (void)n; return -1; }
void *sam_hdr_init(void)                                 { return NULL; }
int   sam_hdr_write(void *fp, const void *h)             { (void)fp; (void)h; return -1; }
long  sam_parse_cigar(const char *in, char **end, void *a, void *m) { (void)in; (void)end; (void)a; (void)m; return -1; }
int   sam_write1(void *fp, const void *h, const void *b) { (void)fp; (void)h; (void)b; return -1; }
