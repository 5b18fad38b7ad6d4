/*
 * ghmm_oracle.c — CPU restatement of the AUGUSTUS GHMM Viterbi path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared library; the product (augustus_b200/) never does.
 *
 * What it restates (reference = /root/reference, AUGUSTUS 3.5.0), column by column, pull style,
 * exactly in the reference's loop order so that ties resolve the same way:
 *   NAMGene::viterbiAndForward      src/namgene.cc:168-365   (column loop, GC-class switching)
 *   NAMGene::getViterbiPath         src/namgene.cc:432-510   (argmax x termProbs, re-derivation walk)
 *   IGenicModel::viterbi...         src/igenicmodel.cc:231-357
 *   IntronModel::viterbi...         src/intronmodel.cc:509-858, emiProbUnderModel :861-1038,
 *                                   seqProb :1046-1108, aSSProb :1116-1188, dSSProb :1195-1248
 *   ExonModel::viterbi...           src/exonmodel.cc:899-1179, endPartEmiProb :1272-1400,
 *                                   notEndPartEmiProb :1417-1859, seqProb :1925-2034,
 *                                   OpenReadingFrame :101-198
 *   ContentStairs::computeStairs    src/motif.cc:543-614
 *   State::setTruncFlag             src/gene.cc:309-321
 *   UtrModel::viterbi...            src/utrmodel.cc:796-1064 (+ endPartEmiProb :1072, notEndPartEmiProb :1167, tssProb :1761,
 *                                   computeTtsProbs :1840, SegProbs statemodel.cc:398-465, EOPList :473-520)      [--UTR=on, 71 states]
 *   NcModel::viterbi...             src/ncmodel.cc:154-359 (+ :366, :447, :702, :744)                              [--nc=on, 83 states]
 *   NAMGene::getSampledPath         src/namgene.cc:367-426, OptionsList::sample vitmatrix.cc:295-320, forward sums of every model
 * Scope: ab initio with or without softmasking (lower-case runs = nonexonpart bonus, extrinsicinfo.cc:1696-1724), no hints files,
 * no protein profile, no overlap mode — what BASELINE.json's five configurations run.  Pinned against the unmodified reference:
 * parity is NOT "unpinned" (tests/test_oracle.py, tests/golden/make_golden*.py).
 *
 * Arithmetic: the reference multiplies LLDouble probabilities; this restatement ADDS log
 * probabilities quantised to Q40 fixed point (int64, 2^-40 nat resolution), the same number
 * system the CUDA path uses, so CUDA-vs-oracle comparisons are bit-exact while
 * oracle-vs-reference comparisons are exact on the path and ~1e-9 on the score.
 *
 * History-dependent behaviour that is restated rather than "fixed": the SnippetProbs memo
 * (statemodel.cc:283-393) keeps the GC class active at first touch, see snip_get() below.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/augb200_params.h"

typedef int64_t sc_t;
#define FRAC_BITS 40
#define NEG  (-((sc_t)1 << 61))
#define NEGT (-((sc_t)1 << 60))
static inline int isneg(sc_t x) { return x <= NEGT; }
static sc_t q(double x) {
    if (!(x > -1e300)) return NEG;      /* -inf or nan */
    return (sc_t)llround(ldexp(x, FRAC_BITS));
}

/* reference StateType values (include/types.hh:492-512) */
enum { T_IGENIC = 0, T_SINGLE = 1, T_INITIAL0 = 2, T_INTERNAL0 = 5, T_TERMINAL = 8,
       T_LESSD0 = 9, T_LONGDSS0 = 10, T_EQUALD0 = 11, T_GEO0 = 12, T_LONGASS0 = 13,
       T_RSINGLE = 36, T_RINITIAL = 37, T_RINTERNAL0 = 38, T_RTERMINAL0 = 41,
       T_RLESSD0 = 44, T_RLONGDSS0 = 45, T_REQUALD0 = 46, T_RGEO0 = 47, T_RLONGASS0 = 48 };
enum { K_IGENIC, K_EXON, K_LESSD, K_LONGDSS, K_EQUALD, K_GEO, K_LONGASS, K_UTR, K_NC };
enum { U_SINGLE, U_INIT, U_INTRON, U_INTRONVAR, U_INTERNAL, U_TERM };     /* order of the utr5.. / utr3.. types, types.hh:498-499 */
enum { T_UTR5SINGLE = 24, T_UTR3SINGLE = 30, T_UTR3TERM = 35, T_RUTR5SINGLE = 59, T_RUTR3SINGLE = 65 };
enum { T_NCSINGLE = 74, T_RNCSINGLE = 80 };   /* ncsingle, ncinit, ncintron, ncintronvar, ncinternal, ncterm; then the reverse six (types.hh:508-511) */
enum { E_SINGLE, E_INITIAL, E_INTERNAL, E_TERMINAL, E_RSINGLE, E_RINITIAL, E_RINTERNAL, E_RTERMINAL };

typedef struct {
    int type, kind, fwd, frame;     /* frame = stateReadingFrames[type] (types.cc:174-188) */
    int ek;                         /* exon kind */
    int uk, u5;                     /* UTR kind (U_*), 1 = 5' UTR */
    int beginPartLen, innerPartOffset, baseOffset, innerPartEndOffset;
    int nanc, anc[16];
} StateInfo;

typedef struct {
    int S, C, k, d, dss_start, dss_end, ass_start, ass_end, ass_up, tiw, init_len, et_len;
    int max_exon_len, min_exon_length, dss_gc_allowed, GCwinsize, weighing;
    int tis_n, tis_k, assm_n, assm_k;
    int dStateLen;
    StateInfo* st;
    sc_t *init, *term, *trans;                   /* trans[c][a][s] */
    sc_t *xemi, *xinit, *xet;                    /* [c][3][1024] */
    sc_t *xpls[8];                               /* [l][c][3][4^(l+1)] */
    sc_t *iemi, *gemi;                           /* [c][1024] */
    double *gpls[8];                             /* igenic Pls, kept as double logs (ratio of sums) */
    sc_t *tis, *assm;                            /* [c][n][4^(k+1)] */
    int tis_nbins; sc_t *tis_bb, *tis_bp;        /* TRANSINITBIN: [c][nbins-1] boundaries, [c][nbins] bin probabilities (exonmodel.cc:1321-1326, 1437-1442) */
    sc_t *ld_single, *ld_initial, *ld_internal, *ld_terminal, *ld_intron;
    int n_ld_exon, n_ld_intron;
    sc_t *ass_pat, *ass_pat_non, *dss_pat, *dss_pat_non;
    sc_t startp[64]; int isstop[64];
    sc_t ochre, amber, opal, probN, log025, log3;
    double centroids[256][4]; int ncent; double wm[4][4];
    int softmask; sc_t nep_bonus;               /* softmasking: ln bonus of a nonexonpart hint of source RM (extrinsicinfo.cc:1696-1724) */
    /* UtrModel (utrmodel.cc), only with --UTR=on */
    int nc;                                     /* NcModel states present (--nc=on) */
    int utr, utr_k, tssup_k, tss_start, tss_end, tata_start, tata_end, d_tata_min, d_tata_max, tuw, dpc, boxlen, tts_spacing;
    int umax, umax3s, umax3t;
    sc_t *u5i, *u5, *u3, *tup;                   /* [c][4^(k+1)] */
    sc_t *tssm, *tsstm, *tatam, *ttsm; int tssm_n, tssm_k, tsstm_n, tsstm_k, tatam_n, tatam_k, ttsm_n, ttsm_k;
    sc_t *aataaa, log_polya, log_nopolya, log2, isstart[64];
    sc_t *uld[2][6], *utl5s, *utl3s; int n_uld[2][6], n_utl5s, n_utl3s;   /* [5'/3' = 1/0][U_*] length distributions */
} Model;

/* ------------------------------------------------------------------ blob reading */
typedef struct { char* buf; size_t n; } Blob;
static const augb200_blob_entry* bfind(const Blob* b, const char* name) {
    const augb200_blob_header* h = (const augb200_blob_header*)b->buf;
    const augb200_blob_entry* e = (const augb200_blob_entry*)(b->buf + sizeof *h);
    for (uint32_t i = 0; i < h->n_entries; i++) if (!strcmp(e[i].name, name)) return &e[i];
    fprintf(stderr, "oracle: blob entry '%s' missing\n", name); exit(3);
}
static int bhas(const Blob* b, const char* name) {
    const augb200_blob_header* h = (const augb200_blob_header*)b->buf;
    const augb200_blob_entry* e = (const augb200_blob_entry*)(b->buf + sizeof *h);
    for (uint32_t i = 0; i < h->n_entries; i++) if (!strcmp(e[i].name, name)) return 1;
    return 0;
}
static int bint(const Blob* b, const char* name) { return *(const int32_t*)(b->buf + bfind(b, name)->offset); }
static double bdbl(const Blob* b, const char* name) { return *(const double*)(b->buf + bfind(b, name)->offset); }
static const double* bdarr(const Blob* b, const char* name, size_t* n) {
    const augb200_blob_entry* e = bfind(b, name); if (n) *n = e->nbytes / 8; return (const double*)(b->buf + e->offset);
}
static const int32_t* biarr(const Blob* b, const char* name, size_t* n) {
    const augb200_blob_entry* e = bfind(b, name); if (n) *n = e->nbytes / 4; return (const int32_t*)(b->buf + e->offset);
}
static sc_t* qarr(const Blob* b, const char* name, size_t* n) {
    size_t m; const double* d = bdarr(b, name, &m);
    sc_t* r = (sc_t*)malloc(m * sizeof(sc_t));
    for (size_t i = 0; i < m; i++) r[i] = q(d[i]);
    if (n) *n = m;
    return r;
}

static void classify(StateInfo* s, const Model* m) {
    int t = s->type;
    s->fwd = (t < 36); s->frame = 0; s->uk = -1;
    if (t >= T_NCSINGLE && t < T_NCSINGLE + 12) {          /* NcModel::NcModel, ncmodel.cc:41-44: uk = position in the six nc types */
        s->kind = K_NC; s->fwd = t < T_RNCSINGLE; s->uk = (t - T_NCSINGLE) % 6; s->u5 = 0;
        return;
    }
    if (t == T_IGENIC) { s->kind = K_IGENIC; s->fwd = 1; return; }
    int e = -1;
    if (t == T_SINGLE) { e = E_SINGLE; }
    else if (t >= T_INITIAL0 && t < T_INITIAL0 + 3) { e = E_INITIAL; s->frame = t - T_INITIAL0; }
    else if (t >= T_INTERNAL0 && t < T_INTERNAL0 + 3) { e = E_INTERNAL; s->frame = t - T_INTERNAL0; }
    else if (t == T_TERMINAL) { e = E_TERMINAL; }
    else if (t == T_RSINGLE) { e = E_RSINGLE; s->frame = 2; }
    else if (t == T_RINITIAL) { e = E_RINITIAL; s->frame = 2; }
    else if (t >= T_RINTERNAL0 && t < T_RINTERNAL0 + 3) { e = E_RINTERNAL; s->frame = t - T_RINTERNAL0; }
    else if (t >= T_RTERMINAL0 && t < T_RTERMINAL0 + 3) { e = E_RTERMINAL; s->frame = t - T_RTERMINAL0; }
    if (e >= 0) {
        s->kind = K_EXON; s->ek = e;
        /* ExonModel::ExonModel, exonmodel.cc:231-279 */
        switch (e) {
        case E_SINGLE: case E_INITIAL: s->beginPartLen = 3 + m->tiw; s->innerPartOffset = 3; break;
        case E_RSINGLE: case E_RTERMINAL: s->beginPartLen = s->innerPartOffset = 3; break;
        default: s->beginPartLen = 0; s->innerPartOffset = s->fwd ? m->ass_end : m->dss_start;
        }
        if (e == E_SINGLE || e == E_TERMINAL) { s->baseOffset = 0; s->innerPartEndOffset = 3; }
        else if (e == E_RSINGLE || e == E_RINITIAL) { s->baseOffset = -m->tiw; s->innerPartEndOffset = 3; }
        else { s->baseOffset = s->innerPartEndOffset = s->fwd ? m->dss_start : m->ass_end; }
        return;
    }
    if ((t >= T_UTR5SINGLE && t <= T_UTR3TERM) || (t >= T_RUTR5SINGLE && t <= T_RUTR5SINGLE + 11)) {
        int o = t - (s->fwd ? T_UTR5SINGLE : T_RUTR5SINGLE);
        s->kind = K_UTR; s->u5 = o < 6; s->uk = o % 6;
        return;
    }
    int base = s->fwd ? T_LESSD0 : T_RLESSD0;
    int off = t - base;
    if (off < 0 || off >= 15) { fprintf(stderr, "oracle: unsupported state type %d (UTR/nc not restated)\n", t); exit(3); }
    s->frame = off / 5;
    switch (off % 5) {
    case 0: s->kind = K_LESSD; break; case 1: s->kind = K_LONGDSS; break; case 2: s->kind = K_EQUALD; break;
    case 3: s->kind = K_GEO; break; case 4: s->kind = K_LONGASS; break;
    }
}

Model* orc_model_load(const char* path) {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); return NULL; }
    Blob b; fseek(f, 0, SEEK_END); b.n = ftell(f); fseek(f, 0, SEEK_SET);
    b.buf = (char*)malloc(b.n); if (fread(b.buf, 1, b.n, f) != b.n) { fclose(f); return NULL; } fclose(f);
    if (memcmp(b.buf, AUGB200_BLOB_MAGIC, 8)) { fprintf(stderr, "oracle: bad blob magic\n"); return NULL; }
    Model* m = (Model*)calloc(1, sizeof *m);
    m->S = bint(&b, "statecount"); m->C = bint(&b, "num_gc_classes");
    m->k = bint(&b, "exon_k");
    if (bint(&b, "intron_k") != m->k || bint(&b, "igenic_k") != m->k) { fprintf(stderr, "oracle: mixed k unsupported\n"); return NULL; }
    m->d = bint(&b, "intron_d");
    m->dss_start = bint(&b, "dss_start"); m->dss_end = bint(&b, "dss_end");
    m->ass_start = bint(&b, "ass_start"); m->ass_end = bint(&b, "ass_end");
    m->ass_up = bint(&b, "ass_upwindow_size"); m->tiw = bint(&b, "trans_init_window");
    m->init_len = bint(&b, "init_coding_len"); m->et_len = bint(&b, "et_coding_len");
    m->max_exon_len = bint(&b, "max_exon_len"); m->min_exon_length = bint(&b, "min_exon_length");
    m->dss_gc_allowed = bint(&b, "dss_gc_allowed"); m->GCwinsize = bint(&b, "GCwinsize");
    m->weighing = bint(&b, "basecount_weighing_type");
    m->tis_nbins = bint(&b, "transinit_nbins");
    if (m->tis_nbins > 0) { m->tis_bp = qarr(&b, "tis_bin_probs", NULL); m->tis_bb = m->tis_nbins > 1 ? qarr(&b, "tis_bin_bounds", NULL) : NULL; }
    m->tis_n = bint(&b, "tis_motif_n"); m->tis_k = bint(&b, "tis_motif_k");
    m->assm_n = bint(&b, "ass_motif_n"); m->assm_k = bint(&b, "ass_motif_k");
    /* intronmodel.cc:519-520 */
    m->dStateLen = m->d - 2 - m->dss_end - m->ass_start - 2 - m->ass_up;
    m->init = qarr(&b, "init_probs", NULL); m->term = qarr(&b, "term_probs", NULL);
    m->trans = qarr(&b, "trans", NULL);
    m->xemi = qarr(&b, "exon_emi", NULL); m->xinit = qarr(&b, "exon_initemi", NULL); m->xet = qarr(&b, "exon_etemi", NULL);
    for (int l = 0; l <= m->k; l++) {
        char nm[32]; sprintf(nm, "exon_pls%d", l); m->xpls[l] = qarr(&b, nm, NULL);
        sprintf(nm, "igenic_pls%d", l);
        size_t n; const double* dd = bdarr(&b, nm, &n);
        m->gpls[l] = (double*)malloc(n * 8); memcpy(m->gpls[l], dd, n * 8);
    }
    m->iemi = qarr(&b, "intron_emi", NULL); m->gemi = qarr(&b, "igenic_emi", NULL);
    m->tis = qarr(&b, "tis_motif", NULL); m->assm = qarr(&b, "ass_motif", NULL);
    size_t n;
    m->ld_single = qarr(&b, "lendist_single", &n); m->n_ld_exon = (int)n;
    m->ld_initial = qarr(&b, "lendist_initial", NULL); m->ld_internal = qarr(&b, "lendist_internal", NULL);
    m->ld_terminal = qarr(&b, "lendist_terminal", NULL);
    m->ld_intron = qarr(&b, "lendist_intron", &n); m->n_ld_intron = (int)n;
    m->ass_pat = qarr(&b, "ass_pattern", NULL); m->ass_pat_non = qarr(&b, "ass_pattern_nonag", NULL);
    m->dss_pat = qarr(&b, "dss_pattern", NULL); m->dss_pat_non = qarr(&b, "dss_pattern_nongt", NULL);
    const double* sp = bdarr(&b, "start_codon_prob", NULL); const int32_t* isx = biarr(&b, "is_stop_codon", NULL);
    for (int i = 0; i < 64; i++) { m->startp[i] = q(sp[i]); m->isstop[i] = isx[i]; }
    m->ochre = q(bdbl(&b, "ochreprob")); m->amber = q(bdbl(&b, "amberprob")); m->opal = q(bdbl(&b, "opalprob"));
    m->probN = q(bdbl(&b, "probNinCoding")); m->log025 = q(log(0.25)); m->log3 = q(log(3.0));
    const double* cen = bdarr(&b, "gc_centroids", &n); m->ncent = (int)(n / 4);
    for (int i = 0; i < m->ncent; i++) for (int j = 0; j < 4; j++) m->centroids[i][j] = cen[4 * i + j];
    const double* w = bdarr(&b, "basecount_weight_matrix", NULL);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m->wm[i][j] = w[4 * i + j];
    if (bhas(&b, "softmasking")) { m->softmask = bint(&b, "softmasking"); m->nep_bonus = q(bdbl(&b, "softmask_bonus")); }   /* absent in older blobs: off */
    if (m->softmask && !bint(&b, "extrinsic_malus_all_one")) { fprintf(stderr, "oracle: extrinsic configurations with a malus are not restated\n"); return NULL; }
    m->utr = bint(&b, "utr_option_on");
    m->nc = bint(&b, "nc_option_on");
    if (m->nc && (!m->utr || m->softmask)) { fprintf(stderr, "oracle: nc states are restated for --UTR=on --softmasking=0 only\n"); return NULL; }
    if (m->utr) {
        m->utr_k = bint(&b, "utr_k"); m->tssup_k = bint(&b, "tssup_k");
        if (m->utr_k != m->k) { fprintf(stderr, "oracle: mixed k unsupported\n"); return NULL; }
        m->tss_start = bint(&b, "tss_start"); m->tss_end = bint(&b, "tss_end");
        m->tata_start = bint(&b, "tata_start"); m->tata_end = bint(&b, "tata_end");
        m->d_tata_min = bint(&b, "d_tss_tata_min"); m->d_tata_max = bint(&b, "d_tss_tata_max");
        m->tuw = bint(&b, "tss_upwindow_size"); m->dpc = bint(&b, "d_polyasig_cleavage");
        m->boxlen = bint(&b, "aataaa_boxlen"); m->tts_spacing = bint(&b, "tts_spacing");
        m->umax = bint(&b, "utr_max_exon_length"); m->umax3s = bint(&b, "utr_max3singlelength"); m->umax3t = bint(&b, "utr_max3termlength");
        m->u5i = qarr(&b, "utr5init_emi", NULL); m->u5 = qarr(&b, "utr5_emi", NULL); m->u3 = qarr(&b, "utr3_emi", NULL);
        m->tup = qarr(&b, "tssup_emi", NULL);
        m->tssm = qarr(&b, "tss_motif", NULL); m->tssm_n = bint(&b, "tss_motif_n"); m->tssm_k = bint(&b, "tss_motif_k");
        m->tsstm = qarr(&b, "tsstata_motif", NULL); m->tsstm_n = bint(&b, "tsstata_motif_n"); m->tsstm_k = bint(&b, "tsstata_motif_k");
        m->tatam = qarr(&b, "tata_motif", NULL); m->tatam_n = bint(&b, "tata_motif_n"); m->tatam_k = bint(&b, "tata_motif_k");
        m->ttsm = qarr(&b, "tts_motif", NULL); m->ttsm_n = bint(&b, "tts_motif_n"); m->ttsm_k = bint(&b, "tts_motif_k");
        m->aataaa = qarr(&b, "aataaa_probs", NULL);
        m->log_polya = q(bdbl(&b, "log_prob_polya")); m->log_nopolya = q(bdbl(&b, "log_no_polya")); m->log2 = q(log(2.0));
        const int32_t* iss = biarr(&b, "is_start_codon", NULL); for (int i = 0; i < 64; i++) m->isstart[i] = iss[i];
        static const char* nm[2][6] = { { "lendist_utr3single", "lendist_utr3initial", NULL, NULL, "lendist_utr3internal", "lendist_utr3terminal" },
                                        { "lendist_utr5single", "lendist_utr5initial", NULL, NULL, "lendist_utr5internal", "lendist_utr5terminal" } };
        for (int f = 0; f < 2; f++) for (int u = 0; u < 6; u++) if (nm[f][u]) { m->uld[f][u] = qarr(&b, nm[f][u], &n); m->n_uld[f][u] = (int)n; }
        m->utl5s = qarr(&b, "taillendist_utr5single", &n); m->n_utl5s = (int)n;
        m->utl3s = qarr(&b, "taillendist_utr3single", &n); m->n_utl3s = (int)n;
    }
    const int32_t* stt = biarr(&b, "state_type", NULL);
    const int32_t* reach = biarr(&b, "state_reachable", NULL);
    m->st = (StateInfo*)calloc(m->S, sizeof(StateInfo));
    for (int s = 0; s < m->S; s++) {
        m->st[s].type = stt[s]; classify(&m->st[s], m);
        if (!reach[s]) { fprintf(stderr, "oracle: unreachable states not restated\n"); return NULL; }
        /* StateModel::initPredecessors, statemodel.cc:41-46: ancestors in state-index order */
        for (int a = 0; a < m->S; a++) if (!isneg(m->trans[(size_t)a * m->S + s])) m->st[s].anc[m->st[s].nanc++] = a;
    }
    free(b.buf);
    return m;
}
int orc_model_statecount(const Model* m) { return m->S; }

/* ------------------------------------------------------------------ per-window context */
typedef struct SnipEnt { int len; sc_t val; struct SnipEnt* next; } SnipEnt;
typedef struct {
    const Model* m; int L; const uint8_t* c;   /* c[i] in 0..3, 4 = other */
    const int* gc;                              /* class per position */
    int cls;                                    /* class of current column */
    int *nsf, *nsr;                             /* nearestStopForward / Reverse */
    sc_t* V;                                    /* [L][S] */
    sc_t *PX[16][3], *PXR[16][3];                 /* exon content prefix sums [class][phi] (lazily built) */
    sc_t *PI[16], *PIR[16];                     /* intron content prefix sums per class (fwd k-mer / rc k-mer) */
    struct SnipEnt **snF, **snL;                /* SnippetProbs restatement: per base first/last entry, [2][L] (fwd, rc) */
    /* forward / sampling (NAMGene::viterbiAndForward with needForwardTable, getSampledPath namgene.cc:367-426) */
    int mode;                                   /* 0 Viterbi only, 1 Viterbi + forward fill, 2 sampling step */
    double* F;                                  /* [L][S] ln forward, -inf = absent */
    double lse_m, lse_s;                        /* running log-sum-exp of the forward sum of the current cell */
    struct Opt { int state, base; double lp; } *opts; int nopt, capopt;     /* OptionsList of the current sampling step */
    /* UtrModel per-sequence state */
    const char* raw;                            /* lower-case sequence (findTATA compares characters) */
    int* pmask;                                 /* softmasking: pmask[i] = number of lower-case input bases before i; [L+1] */
    int cur_gc;                                 /* NAMGene::curGCIdx */
    int walking;                                /* 1 while backtracking / sampling (algovar doBacktracking / doSampling) */
    sc_t* seg[8];                               /* SegProbs::cumProds, [L+1], NEG = 0 = not computed; [7] = NcModel::segProbs */
    sc_t *tssP[2], *ttsP[2];                    /* tssProbsPlus/Minus memo (UNSET = -1), ttsProbPlus/Minus; index 0 = plus */
    struct Eop { int *v, n, cap, it, inCache; } *eop;   /* EOPList per state */
    int eop_off;                                /* debugging: 1 = scan every endOfPred (no EOPList) */
    sc_t* assMemo[2]; int* assMemoGen[2]; int assGen, assN;     /* IntronModel::aSSProb memoF / memoR */
} Ctx;

static inline int at(const Ctx* x, int p) { return (p < 0 || p >= x->L) ? 5 : x->c[p]; }
static inline int cmpl(int b) { return b < 4 ? 3 - b : b; }
/* product of the nonexonpart bonuses over positions lo..hi (each lower-case run is one hint covering its positions; the runs are
 * disjoint): the intronpart / nonexonpart loops of igenicmodel.cc:306-316, intronmodel.cc:1011-1036, utrmodel.cc:1143-1158,1521-1545 */
static inline sc_t NB(const Ctx* x, int lo, int hi) {
    if (!x->m->softmask) return 0;
    if (lo < 0) lo = 0;
    if (hi > x->L - 1) hi = x->L - 1;
    return lo > hi ? 0 : x->m->nep_bonus * (sc_t)(x->pmask[hi + 1] - x->pmask[lo]);
}
/* Seq2Int::operator() over n bases starting at p (geneticcode.hh:166-173); -1 on invalid nucleotide */
static int s2i(const Ctx* x, int p, int n) {
    int e = 0; for (int i = 0; i < n; i++) { int b = at(x, p + i); if (b > 3) return -1; e = (e << 2) | b; } return e;
}
/* Seq2Int::rc (geneticcode.hh:174-179) */
static int s2irc(const Ctx* x, int p, int n) {
    int e = 0; for (int i = 0; i < n; i++) { int b = at(x, p + i); if (b > 3) return -1; e |= (3 - b) << (2 * i); } return e;
}
static inline int mod3(int k) { return k >= 0 ? k % 3 : (k % 3 + 3) % 3; }
static inline int is2(const Ctx* x, int p, int a, int b) { return at(x, p) == a && at(x, p + 1) == b; }
enum { A_ = 0, C_ = 1, G_ = 2, T_ = 3 };
static int onGenDSS(const Ctx* x, int p) { return is2(x, p, G_, T_) || (x->m->dss_gc_allowed && is2(x, p, G_, C_)); }
static int onGenRDSS(const Ctx* x, int p) { return is2(x, p, A_, C_) || (x->m->dss_gc_allowed && is2(x, p, G_, C_)); }
/* statemodel.hh:98-117, no hints */
static int possDSS(const Ctx* x, int pos) { return pos >= 1 && pos <= x->L - 2 && onGenDSS(x, pos); }
static int possRDSS(const Ctx* x, int pos) { return pos >= 1 && pos <= x->L - 2 && onGenRDSS(x, pos - 1); }
static int possASS(const Ctx* x, int pos) { return pos >= 1 && pos <= x->L - 2 && is2(x, pos - 1, A_, G_); }
static int possRASS(const Ctx* x, int pos) { return pos >= 1 && pos <= x->L - 2 && is2(x, pos, C_, T_); }
static int isStop(const Ctx* x, int p) { int i = s2i(x, p, 3); return i >= 0 && x->m->isstop[i]; }
static int isRCStop(const Ctx* x, int p) { int i = s2irc(x, p, 3); return i >= 0 && x->m->isstop[i]; }

/* OpenReadingFrame::OpenReadingFrame, exonmodel.cc:101-156 */
static void orf_init(Ctx* x) {
    int n = x->L;
    x->nsf = (int*)calloc(n + 3, sizeof(int)); x->nsr = (int*)calloc(n + 3, sizeof(int));
    for (int r = 0; r < 3; r++) {
        int sp = -1; for (int i = r; i <= n - 3; i += 3) { if (isStop(x, i)) sp = i; x->nsf[i] = sp; }
        sp = -1; for (int i = r; i <= n - 3; i += 3) { if (isRCStop(x, i)) sp = i; x->nsr[i] = sp; }
    }
    if (n > 5) {
        x->nsf[n - 2] = x->nsf[n - 5]; x->nsf[n - 1] = x->nsf[n - 4];
        x->nsr[n - 2] = x->nsr[n - 5]; x->nsr[n - 1] = x->nsr[n - 4];
    }
}
/* OpenReadingFrame::leftmostExonBegin, exonmodel.cc:165-198 */
static int leftmostExonBegin(const Ctx* x, int frame, int base, int forward) {
    int pos, n = x->L;
    if (forward) pos = (frame == 0 || frame == 1) ? base - frame - 3 : base - frame;
    else pos = (frame == 1 || frame == 2) ? base + frame - 5 : base - 2;
    if (pos >= n) pos -= 3 * ((pos - n + 3) / 3);
    int lmb = pos >= 0 ? (forward ? x->nsf[pos] : x->nsr[pos]) + 1 : 0;
    const Model* m = x->m;
    int max_allowed = m->max_exon_len - m->ass_up - m->ass_start - 2 - 2 - m->dss_start;
    if (lmb < base - max_allowed) lmb = base - max_allowed;
    return lmb;
}

/* ------------------------------------------------------------------ GC stairs, motif.cc:543-614 */
static int nearestClass(const Model* m, const int cnt[4]) {
    double sum = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    double r[4] = {0.25, 0.25, 0.25, 0.25};
    if (sum > 0) for (int i = 0; i < 4; i++) r[i] = cnt[i] / sum;
    double best = -1; int ret = -1;
    for (int i = 0; i < m->ncent; i++) {
        double w;
        if (m->weighing == 3) {          /* multiNormalKernelWeight, motif.cc:156-177 */
            double z[4], tmp[4] = {0, 0, 0, 0}, t = 0;
            for (int j = 0; j < 4; j++) z[j] = r[j] - m->centroids[i][j];
            for (int j = 0; j < 4; j++) for (int a = 0; a < 4; a++) tmp[j] += z[a] * m->wm[a][j];
            for (int a = 0; a < 4; a++) t += tmp[a] * z[a];
            w = 1 + 9 * exp(-t);
        } else if (m->weighing == 2) {   /* gcContentClassWeight, motif.cc:135-153 */
            double g1 = r[2] + r[1], g2 = m->centroids[i][2] + m->centroids[i][1];
            int c1 = g1 < .43 ? 0 : g1 < .51 ? 1 : g1 < .57 ? 2 : 3, c2 = g2 < .43 ? 0 : g2 < .51 ? 1 : g2 < .57 ? 2 : 3;
            w = c1 == c2;
        } else w = 1;
        if (w > best) { best = w; ret = i; }
    }
    return ret;
}
void orc_gc_stairs(const Model* m, const uint8_t* c, int n, int* idx) {
    for (int i = 0; i < n; i++) idx[i] = -1;
    int win = m->GCwinsize; if (win > n || win < 1) win = n;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < win; i++) if (c[i] < 4) cnt[c[i]]++;
    int xx = nearestClass(m, cnt);
    for (int i = 0; i <= win / 2; i++) idx[i] = xx;
    for (int i = win / 2 + 1; i <= n - (win + 1) / 2; i++) {
        int a = c[i + (win + 1) / 2 - 1], r = c[i - win / 2 - 1];
        if (a < 4) cnt[a]++;
        if (r < 4) cnt[r]--;
        idx[i] = xx = nearestClass(m, cnt);
    }
    for (int i = n - (win + 1) / 2 + 1; i < n; i++) idx[i] = xx;
    int x2 = -2, lastStep = 0, tot = 1000;
    for (int i = 0; i < n; i++) if (idx[i] != x2) {
        if (i - lastStep < tot && lastStep > 0 && idx[lastStep - 1] == idx[i])
            for (int j = lastStep; j < i; j++) idx[j] = idx[i];
        lastStep = i; x2 = idx[i];
    }
}

/* ------------------------------------------------------------------ content prefix sums */
static sc_t exon_emi1(const Ctx* x, int cls, int fwd, int f, int p) {
    const Model* m = x->m;
    int pn = fwd ? s2i(x, p - m->k, m->k + 1) : s2irc(x, p, m->k + 1);
    if (pn < 0) return m->probN;
    return m->xemi[((size_t)cls * 3 + f) << (2 * (m->k + 1)) | pn];
}
static void build_PX(Ctx* x, int cls) {
    int L = x->L;
    for (int phi = 0; phi < 3; phi++) {
        sc_t* a = (sc_t*)malloc((L + 1) * sizeof(sc_t)); sc_t* b = (sc_t*)malloc((L + 1) * sizeof(sc_t));
        /* a[p+1] = sum_{i<=p} fwd emission with frame mod3(phi+i); b likewise reverse with frame mod3(phi-i) */
        a[0] = b[0] = 0;
        for (int p = 0; p < L; p++) {
            a[p + 1] = a[p] + (p >= x->m->k ? exon_emi1(x, cls, 1, mod3(phi + p), p) : 0);
            b[p + 1] = b[p] + exon_emi1(x, cls, 0, mod3(phi - p), p);
        }
        x->PX[cls][phi] = a; x->PXR[cls][phi] = b;
    }
}
/* ExonModel::seqProb, exonmodel.cc:1925-1973: product over [left,right] of the 3-periodic content model */
static sc_t exon_seqProb(Ctx* x, int fwd, int left, int right, int frameOfRight) {
    if (left > right) return 0;
    int cls = x->cls;
    if (!x->PX[cls][0]) build_PX(x, cls);
    if (fwd) { int phi = mod3(frameOfRight - right); return x->PX[cls][phi][right + 1] - x->PX[cls][phi][left]; }
    int phi = mod3(frameOfRight + right); return x->PXR[cls][phi][right + 1] - x->PXR[cls][phi][left];
}
/* eTermSeqProb :1979-2012 and initialSeqProb :2018-2034 — short windows, summed directly */
static sc_t exon_shortProb(const Ctx* x, const sc_t* tab, int fwd, int left, int right, int frameOfRight) {
    const Model* m = x->m; sc_t s = 0;
    for (int p = right; p >= left; p--) {
        int f = fwd ? mod3(frameOfRight - right + p) : mod3(frameOfRight + right - p);
        int pn = fwd ? s2i(x, p - m->k, m->k + 1) : s2irc(x, p, m->k + 1);
        s += pn < 0 ? m->probN : tab[((size_t)x->cls * 3 + f) << (2 * (m->k + 1)) | pn];
    }
    return s;
}
static sc_t intron_emi1(const Ctx* x, int cls, int p) {          /* forward k-mer ending at p */
    const Model* m = x->m;
    if (p - m->k < 0) return m->log025;
    int pn = s2i(x, p - m->k, m->k + 1);
    return pn < 0 ? m->log025 : m->iemi[((size_t)cls << (2 * (m->k + 1))) | pn];
}
static sc_t intron_emi1r(const Ctx* x, int cls, int p) {         /* rc k-mer starting at p, statemodel.cc:298-307 */
    const Model* m = x->m;
    if (!(p >= 0 && p + m->k < x->L)) return m->log025;
    int pn = s2irc(x, p, m->k + 1);
    return pn < 0 ? m->log025 : m->iemi[((size_t)cls << (2 * (m->k + 1))) | pn];
}
static void build_PI(Ctx* x, int cls) {
    int L = x->L; sc_t* a = (sc_t*)malloc((L + 1) * sizeof(sc_t)); sc_t* b = (sc_t*)malloc((L + 1) * sizeof(sc_t));
    a[0] = b[0] = 0;
    for (int p = 0; p < L; p++) { a[p + 1] = a[p] + intron_emi1(x, cls, p); b[p + 1] = b[p] + intron_emi1r(x, cls, p); }
    x->PI[cls] = a; x->PIR[cls] = b;
}
static sc_t intron_sum(Ctx* x, int rc, int left, int right) {   /* sum over [left,right], class of current column */
    if (left > right) return 0;
    int cls = x->cls; if (!x->PI[cls]) build_PI(x, cls);
    const sc_t* P = rc ? x->PIR[cls] : x->PI[cls];
    return P[right + 1] - P[left];
}

/*
 * SnippetProbs::getSeqProb / addProb / SnippetList::getProb, statemodel.cc:283-393.
 * The memo is restated entry for entry because it is NOT a pure cache: an elementary piece keeps the
 * emission table (GC class) that was active when it was first computed (IntronModel::updateToLocalGC,
 * intronmodel.cc:495-503, swaps the table pointer but never clears the lists), so on windows with
 * more than one GC class the value of a lessD segment depends on the history of calls.
 */
static sc_t snip_elem(Ctx* x, int rc, int base, int len) {       /* getElemSeqProb :283-310, current class */
    int cls = x->cls; if (!x->PI[cls]) build_PI(x, cls);
    const sc_t* P = rc ? x->PIR[cls] : x->PI[cls];
    return P[base + 1] - P[base - len + 1];
}
static void snip_add(Ctx* x, int rc, int base, int len, sc_t p) { /* addProb :344-370 */
    SnipEnt** F = x->snF + (size_t)rc * x->L; SnipEnt** La = x->snL + (size_t)rc * x->L;
    SnipEnt* e = (SnipEnt*)malloc(sizeof *e); e->len = len; e->val = p; e->next = NULL;
    if (!F[base]) { F[base] = La[base] = e; return; }
    if (La[base]->len < len) { La[base]->next = e; La[base] = e; return; }
    SnipEnt* t = F[base];
    if (t->len > len) { e->next = t; F[base] = e; return; }
    while (t && t->next && t->next->len < len) t = t->next;
    if (t->next && t->next->len == len) { free(e); return; }      /* "tried to add snippet of same length" */
    e->next = t->next; t->next = e;
}
static sc_t snip_get(Ctx* x, int rc, int base, int len) {         /* getSeqProb :312-342 */
    if (len == 0) return 0;
    SnipEnt** F = x->snF + (size_t)rc * x->L; SnipEnt** La = x->snL + (size_t)rc * x->L;
    sc_t p;
    if (F[base]) {
        if (La[base]->len < len) {
            int ll = La[base]->len;
            p = snip_get(x, rc, base - ll, len - ll) + La[base]->val;
            snip_add(x, rc, base, len, p);
        } else {
            int partlen = len; SnipEnt* t = F[base];               /* SnippetList::getProb :379-393 */
            if (t->len > partlen) { partlen = 0; p = 0; }
            else {
                while (t->next && t->next->len < partlen) t = t->next;
                if (!t->next || t->next->len > partlen) { partlen = t->len; p = t->val; }
                else p = t->next->val;
            }
            if (partlen != len) {
                sc_t prest;
                if (partlen == 0) { prest = snip_elem(x, rc, base, len); snip_add(x, rc, base, len, prest); }
                else prest = snip_get(x, rc, base - partlen, len - partlen);
                p += prest;
            }
        }
    } else { p = snip_elem(x, rc, base, len); snip_add(x, rc, base, len, p); }
    return p;
}

/* Motif::seqProb, motif.cc:308-331 */
static sc_t motif_fwd(const Ctx* x, const sc_t* tab, int n, int k, int p) {
    sc_t s = 0; size_t w = (size_t)1 << (2 * (k + 1));
    for (int i = 0; i < n; i++) { int pn = s2i(x, p + i - k, k + 1); s += pn < 0 ? x->m->log025 : tab[((size_t)x->cls * n + i) * w + pn]; }
    return s;
}
static sc_t motif_rc(const Ctx* x, const sc_t* tab, int n, int k, int p) {
    sc_t s = 0; size_t w = (size_t)1 << (2 * (k + 1));
    for (int i = 0; i < n; i++) { int pn = s2irc(x, p + i, k + 1); s += pn < 0 ? x->m->log025 : tab[((size_t)x->cls * n + (n - 1 - i)) * w + pn]; }
    return s;
}

typedef struct { sc_t max; int state, base; } Oli;
#define VV(j, s) (x->V[(size_t)(j) * m->S + (s)])
#define TR(a, s) (m->trans[((size_t)x->cls * m->S + (a)) * m->S + (s)])
#define FF(j, s) (x->F[(size_t)(j) * m->S + (s)])
/* value used for the "is this predecessor cell non-zero" tests: the Viterbi matrix while filling, the forward matrix while
 * sampling (the reference clears the Viterbi matrix before it samples, namgene.cc:807-808) */
static inline sc_t PVAL(const Ctx* x, int col, int a) {
    const Model* m = x->m;
    if (x->mode == 2) return FF(col, a) > -1e300 ? 0 : NEG;
    return VV(col, a);
}
static inline double sc2d(sc_t v) { return ldexp((double)v, -FRAC_BITS); }
/* one (predecessor a at column col, reported end e) option with log transition*emission `te`:
 *   fill mode: fwdsum += forward[col][a] * transEmiProb           (e.g. exonmodel.cc:1094-1101, intronmodel.cc:609-617)
 *   sampling : optionslist->add(a, e, forward[col][a] * transEmiProb) if > 0 */
static void fwd_option(Ctx* x, int a, int col, int e, sc_t te) {
    const Model* m = x->m;
    if (x->mode == 0) return;
    double f = FF(col, a);
    if (!(f > -1e300) || isneg(te)) return;
    double lp = f + sc2d(te);
    if (x->mode == 1) {
        if (lp > x->lse_m) { x->lse_s = x->lse_s * exp(x->lse_m - lp) + 1.0; x->lse_m = lp; }
        else x->lse_s += exp(lp - x->lse_m);
    } else {
        if (x->nopt == x->capopt) { x->capopt = x->capopt ? 2 * x->capopt : 256; x->opts = realloc(x->opts, x->capopt * sizeof *x->opts); }
        x->opts[x->nopt].state = a; x->opts[x->nopt].base = e; x->opts[x->nopt].lp = lp; x->nopt++;
    }
}

/* ------------------------------------------------------------------ igenic, igenicmodel.cc:231-357 */
static sc_t igenic_emi(const Ctx* x, int j) {
    const Model* m = x->m;
    if (j > m->k) {
        int pn = s2i(x, j - m->k, m->k + 1);
        return pn < 0 ? m->log025 : m->gemi[((size_t)x->cls << (2 * (m->k + 1))) | pn];
    }
    int basek = s2i(x, 0, j + 1);
    if (basek < 0) return m->log025;
    const double* P = m->gpls[j] + ((size_t)x->cls << (2 * (j + 1)));
    /* quirk kept: the normaliser indexes basek/4 + i, not 4*(basek/4) + i (igenicmodel.cc:350) */
    double den = exp(P[basek / 4]) + exp(P[basek / 4 + 1]) + exp(P[basek / 4 + 2]) + exp(P[basek / 4 + 3]);
    return q(P[basek] - log(den));
}
static void igenic_eval(Ctx* x, int s, int j, Oli* o) {
    const Model* m = x->m; const StateInfo* st = &m->st[s];
    sc_t emi = igenic_emi(x, j) + NB(x, j, j);
    /* max starts at -1 (igenicmodel.cc:238): the first ancestor is recorded even when every product is 0 */
    o->base = j - 1; o->max = NEG; o->state = st->nanc ? st->anc[0] : -1;
    for (int i = 0; i < st->nanc; i++) {
        int a = st->anc[i]; sc_t pv = PVAL(x, j - 1, a);
        if (isneg(pv)) continue;
        fwd_option(x, a, j - 1, j - 1, TR(a, s) + emi);
        sc_t cur = pv + (TR(a, s) + emi);
        if (cur > o->max) { o->max = cur; o->state = a; }
    }
}

/* ------------------------------------------------------------------ splice site scores */
/* IntronModel::dSSProb, intronmodel.cc:1195-1248 */
static sc_t dSSProb(const Ctx* x, int base, int fwd) {
    const Model* m = x->m; int nonGT, idx;
    if (fwd) {
        int dsspos = base + m->dss_start;
        if (!possDSS(x, dsspos)) return NEG;
        nonGT = !is2(x, dsspos, G_, T_);
        int a = s2i(x, base, m->dss_start), b = s2i(x, dsspos + 2, m->dss_end);
        if (a < 0 || b < 0) return NEG;
        idx = (a << (2 * m->dss_end)) | b;
    } else {
        int dsspos = base + m->dss_end;
        if (!possRDSS(x, dsspos + 1)) return NEG;
        nonGT = !is2(x, dsspos, A_, C_);
        /* putReverseComplement of seq[dsspos+2, +dss_start) then of seq[base, +dss_end) */
        int a = 0, b = 0;
        for (int i = m->dss_start - 1; i >= 0; i--) { int c = at(x, dsspos + 2 + i); if (c > 3) return NEG; a = (a << 2) | (3 - c); }
        for (int i = m->dss_end - 1; i >= 0; i--) { int c = at(x, base + i); if (c > 3) return NEG; b = (b << 2) | (3 - c); }
        idx = (a << (2 * m->dss_end)) | b;
    }
    return nonGT ? m->dss_pat_non[idx] : m->dss_pat[idx];
}
/* IntronModel::aSSProb, intronmodel.cc:1116-1188 */
static sc_t aSSProb_raw(const Ctx* x, int base, int fwd, int* stored) {
    const Model* m = x->m; int nonAG, a, b; sc_t motif;
    *stored = 0;
    if (fwd) {
        int asspos = base + m->ass_up + m->ass_start;
        if (!possASS(x, asspos + 1)) return NEG;
        nonAG = !is2(x, asspos, A_, G_);
        a = s2i(x, base + m->ass_up, m->ass_start); b = s2i(x, asspos + 2, m->ass_end);
        motif = base >= m->assm_k ? motif_fwd(x, m->assm, m->assm_n, m->assm_k, base) : NEG;
    } else {
        int asspos = base + m->ass_end;
        if (!possRASS(x, asspos)) return NEG;
        nonAG = !is2(x, asspos, C_, T_);
        a = 0; b = 0;
        for (int i = m->ass_start - 1; i >= 0; i--) { int c = at(x, asspos + 2 + i); if (c > 3) { a = -1; break; } a = (a << 2) | (3 - c); }
        for (int i = m->ass_end - 1; i >= 0; i--) { int c = at(x, base + i); if (c > 3) { b = -1; break; } b = (b << 2) | (3 - c); }
        int motifstart = base + m->ass_start + m->ass_end + 2, motifend = motifstart + m->ass_up;
        motif = motifend + m->assm_k < x->L ? motif_rc(x, m->assm, m->assm_n, m->assm_k, motifstart) : m->ass_up * m->log025;
    }
    sc_t pat;
    if (a < 0 || b < 0) pat = q(log(0.001) + (m->ass_start + m->ass_end) * log(0.25));
    else { int idx = (a << (2 * m->ass_end)) | b; pat = nonAG ? m->ass_pat_non[idx] : m->ass_pat[idx]; }
    *stored = 1;
    if (isneg(motif) || isneg(pat)) return NEG;
    return motif + pat;
}
/* the memo of IntronModel::aSSProb (intronmodel.cc:1119-1136,1181-1185): values are kept per first base and strand until more
 * than 1000 forward entries have accumulated (or a reset call with base < 0); a kept value carries the GC class (motif) that
 * was active when it was computed.  Only visible with UTR states, which ask for splice sites far behind the current column. */
static sc_t aSSProb(Ctx* x, int base, int fwd) {
    if (base < 0 || x->assN > 1000) { x->assGen++; x->assN = 0; if (base < 0) return NEG; }
    int slot = base <= x->L ? base : x->L;
    sc_t* val = x->assMemo[fwd ? 0 : 1]; int* gen = x->assMemoGen[fwd ? 0 : 1];
    if (gen[slot] == x->assGen && base <= x->L) return val[slot];
    int stored; sc_t v = aSSProb_raw(x, base, fwd, &stored);
    if (stored && base <= x->L) { val[slot] = v; gen[slot] = x->assGen; if (fwd) x->assN++; }
    return v;
}

/* ------------------------------------------------------------------ introns, intronmodel.cc:509-858 */
static void intron_eval(Ctx* x, int s, int j, Oli* o) {
    const Model* m = x->m; const StateInfo* st = &m->st[s];
    o->max = NEG; o->state = -1; o->base = -1;
    int L = x->L, fwd = st->fwd;
    if (st->kind == K_LESSD) {
        int eob = fwd ? j + m->ass_up + m->ass_start + 2 : j + m->dss_end + 2;
        if (fwd ? (eob - 2 + 1 < L - 1 && !possASS(x, eob)) : (eob - 2 + 1 < L - 1 && !possRDSS(x, eob))) return;
        int lessD0 = fwd && st->frame == 0, rlessD2 = !fwd && st->frame == 2;
        int cod[3] = {4, 4, 4}, spl = !lessD0 && !rlessD2;
        if (spl && eob < L - 2) {
            if (fwd && st->frame == 1) { cod[1] = at(x, eob + 1); cod[2] = at(x, eob + 2); }
            else if (fwd && st->frame == 2) { cod[2] = at(x, eob + 1); }
            else if (!fwd && st->frame == 0) { cod[0] = cmpl(at(x, eob + 1)); }
            else if (!fwd && st->frame == 1) { cod[0] = cmpl(at(x, eob + 2)); cod[1] = cmpl(at(x, eob + 1)); }
        }
        int lme = j - m->dStateLen; if (lme < 0) lme = 0;
        for (int e = j - 1; e >= lme; e--) {
            int any = 0;
            for (int i = 0; i < st->nanc; i++) if (!isneg(PVAL(x, e, st->anc[i]))) { any = 1; break; }
            if (!any) continue;
            /* emiProbUnderModel(e+1, j), lessD branch :924-1000 */
            int begin = e + 1, bob;
            if (fwd) { bob = begin - m->dss_end - 2; if (bob >= 0 && !possDSS(x, bob)) continue; }
            else { bob = begin - (m->ass_up + m->ass_start + 2); if (bob >= 0 && !possRASS(x, bob)) continue; }
            if (spl && bob > 1) {
                int c0 = cod[0], c1 = cod[1], c2 = cod[2];
                if (fwd && st->frame == 1) c0 = at(x, bob - 1);
                else if (fwd && st->frame == 2) { c0 = at(x, bob - 2); c1 = at(x, bob - 1); }
                else if (!fwd && st->frame == 0) { c1 = cmpl(at(x, bob - 1)); c2 = cmpl(at(x, bob - 2)); }
                else if (!fwd && st->frame == 1) { c2 = cmpl(at(x, bob - 1)); }
                if (c0 < 4 && c1 < 4 && c2 < 4 && m->isstop[(c0 << 4) | (c1 << 2) | c2]) continue;
            }
            int ilen = eob - bob + 1;
            if (ilen > m->d || ilen < 0 || ilen >= m->n_ld_intron) continue;   /* > d only with hints */
            sc_t ld = m->ld_intron[ilen]; if (isneg(ld)) continue;
            sc_t emi = ld + snip_get(x, !fwd, j, j - begin + 1) + NB(x, begin, j);
            for (int i = 0; i < st->nanc; i++) {
                int a = st->anc[i]; sc_t pv = PVAL(x, e, a); if (isneg(pv)) continue;
                fwd_option(x, a, e, e, TR(a, s) + emi);
                sc_t pp = pv + (TR(a, s) + emi);
                if (pp > o->max) { o->max = pp; o->state = a; o->base = e; }
            }
        }
        return;
    }
    int eop, abort_; sc_t emi;
    int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
    switch (st->kind) {
    case K_LONGDSS:
        eop = j - dssw;
        abort_ = fwd ? (eop < 0 || !possDSS(x, j - m->dss_end - 2 + 1)) : (eop < 0 || !possRDSS(x, j - m->dss_start));
        break;
    case K_EQUALD: eop = j - m->dStateLen; abort_ = eop < 0; break;
    case K_GEO: eop = j - 1; abort_ = 0; break;
    default: /* K_LONGASS */
        eop = j - assw - m->ass_up;
        abort_ = fwd ? (eop < 0 || !possASS(x, j - m->ass_end)) : (eop < 0 || !possRASS(x, j - m->ass_up - m->ass_start - 2 + 1));
    }
    if (abort_) return;
    int any = 0;
    for (int i = 0; i < st->nanc; i++) if (!isneg(PVAL(x, eop, st->anc[i]))) { any = 1; break; }
    if (!any) return;
    /* the part of [eop+1, j] that belongs to the intron gets the nonexonpart bonus (intronBegin / intronEnd, :873-922,1011-1032) */
    int ib = eop + 1, ie = j;
    switch (st->kind) {
    case K_LONGDSS: emi = dSSProb(x, j - dssw + 1, fwd); if (fwd) ib = j - 2 - m->dss_end + 1; else ie = j - m->dss_start; break;
    case K_EQUALD: emi = intron_sum(x, 0, eop + 1, j); break;     /* forward k-mers also for requalD (:1046-1108) */
    case K_GEO: emi = intron_emi1(x, x->cls, j); break;           /* :895-915, forward k-mer also for rgeometric */
    default: emi = aSSProb(x, j - assw - m->ass_up + 1, fwd); if (fwd) ie = j - m->ass_end; else ib = eop + 1 + m->ass_end;
    }
    if (isneg(emi)) return;
    emi += NB(x, ib, ie);
    o->base = eop;
    for (int i = 0; i < st->nanc; i++) {
        int a = st->anc[i]; sc_t pv = PVAL(x, eop, a); if (isneg(pv)) continue;
        fwd_option(x, a, eop, eop, TR(a, s) + emi);
        sc_t pp = pv + (TR(a, s) + emi);
        if (pp > o->max) { o->max = pp; o->state = a; }
    }
}

/* ------------------------------------------------------------------ exons */
/* ExonModel::endPartEmiProb, exonmodel.cc:1272-1400 (no hints) */
/* BinnedMMGroup::getIndex (merkmal.cc:155-168) + avprobs: the start codon x TIS motif probability goes to the probability of its bin */
static sc_t tis_bin(const Ctx* x, sc_t p) {
    const Model* m = x->m; int nb = m->tis_nbins;
    if (nb < 1) return p;
    const sc_t* bb = m->tis_bb + (size_t)x->cls * (nb - 1); int a = 0, b = nb - 1;
    while (a < b) { int mid = (a + b) / 2; if (p < bb[mid]) b = mid; else a = mid + 1; }
    return m->tis_bp[(size_t)x->cls * nb + a];
}
static sc_t endPart(Ctx* x, const StateInfo* st, int end) {
    const Model* m = x->m; int L = x->L;
    switch (st->ek) {
    case E_SINGLE: case E_TERMINAL: {
        int sp = end - 3 + 1;
        if (sp < 0 || sp > L - 3 || !isStop(x, sp)) return NEG;
        int a = at(x, sp), b = at(x, sp + 1), c = at(x, sp + 2);
        if (a == T_ && b == A_ && c == A_) return m->ochre;
        if (a == T_ && b == A_ && c == G_) return m->amber;
        if (a == T_ && b == G_ && c == A_) return m->opal;
        fprintf(stderr, "oracle: unknown stop codon\n"); exit(3);
    }
    case E_RSINGLE: case E_RINITIAL: {
        int sp = end - m->tiw - 3 + 1;
        if (sp < 0) return NEG;
        int pn = s2irc(x, sp, 3);
        if (pn < 0 || isneg(m->startp[pn])) return NEG;
        sc_t p = m->startp[pn];
        if (sp + 3 + m->tiw - 1 + m->tis_k < L) p = tis_bin(x, p + motif_rc(x, m->tis, m->tis_n, m->tis_k, sp + 3));
        else p = (L - (sp + 3)) * m->log025;
        return p;
    }
    case E_INITIAL: case E_INTERNAL: {
        int dsspos = end + m->dss_start + 1;
        if (end == L - 1) return 0;
        if ((dsspos + 2 - 1 < L && !possDSS(x, dsspos)) || end + m->dss_start >= L ||
            leftmostExonBegin(x, st->frame - 1, end + m->dss_start, 1) >= end) return NEG;
        return 0;
    }
    default: { /* E_RTERMINAL, E_RINTERNAL */
        int asspos = end + m->ass_end + 1;
        if (end == L - 1) return 0;
        if (end + m->ass_end + 2 < L && possRASS(x, asspos)) return 0;
        return NEG;
    }
    }
}

/* ExonModel::notEndPartEmiProb, exonmodel.cc:1417-1859 (no hints) */
static sc_t notEndPart(Ctx* x, const StateInfo* st, int bos, int right, int frameOfRight) {
    const Model* m = x->m; int L = x->L, k = m->k, fwd = st->fwd, win = st->frame;
    sc_t beginPart;
    int bobe = bos - st->innerPartOffset;
    switch (st->ek) {
    case E_SINGLE: case E_INITIAL: {
        if (!(bobe >= 0 && bobe < L - 2)) return NEG;
        int pn = s2i(x, bobe, 3);
        if (pn < 0 || isneg(m->startp[pn])) return NEG;
        beginPart = m->startp[pn];
        int tis = bobe - m->tiw;
        if (tis > m->tis_k) beginPart = tis_bin(x, beginPart + motif_fwd(x, m->tis, m->tis_n, m->tis_k, tis));
        else beginPart += (sc_t)(bos - 3) * m->log025;
        break;
    }
    case E_TERMINAL: case E_INTERNAL:
        if (bos > 0) {
            if (bobe < 0 || (bobe - 2 >= 0 && !possASS(x, bobe - 1))) return NEG;
            beginPart = 0;
        } else if (bos == 0) beginPart = 0;
        else return NEG;
        break;
    case E_RSINGLE: case E_RTERMINAL: {
        if (bobe < 0) return NEG;
        int a = at(x, bobe), b = at(x, bobe + 1), c = at(x, bobe + 2);
        if (a == T_ && b == T_ && c == A_) beginPart = m->ochre;
        else if (a == C_ && b == T_ && c == A_) beginPart = m->amber;
        else if (a == T_ && b == C_ && c == A_) beginPart = m->opal;
        else return NEG;
        if (isneg(beginPart)) return NEG;
        break;
    }
    default: /* E_RINITIAL, E_RINTERNAL */
        if (bos == 0) beginPart = 0;
        else if (bobe < 0 || (bobe - 2 > 0 && !possRDSS(x, bobe - 1))) return NEG;
        else beginPart = 0;
    }
    sc_t rest;
    if (bos > right) {
        rest = -(sc_t)(bos - right - 1) * m->log025;      /* POWER4TOTHE(bos-right-1) */
    } else if (right - bos <= k) {
        int l = right - bos;
        int pn = fwd ? s2i(x, bos, l + 1) : s2irc(x, bos, l + 1);
        if (pn < 0) rest = (sc_t)(l + 1) * m->probN;
        else {
            int f = fwd ? frameOfRight : mod3(frameOfRight + right - bos);
            rest = m->xpls[l][(((size_t)x->cls * 3 + f) << (2 * (l + 1))) | pn];
        }
    } else {
        int endOfStart = bos + k - 1, beginOfInitP = right - (k - 1);
        if (k == 0) rest = 0;
        else {
            int pn = fwd ? s2i(x, bos, k) : s2irc(x, beginOfInitP, k);
            if (pn < 0) rest = (sc_t)k * m->probN;
            else {
                int f = fwd ? mod3(frameOfRight - right + endOfStart) : mod3(frameOfRight + right - beginOfInitP);
                rest = m->xpls[k - 1][(((size_t)x->cls * 3 + f) << (2 * k)) | pn];
            }
        }
        if (isneg(rest)) return NEG;
        int endOfInitial, beginOfTerm, endOfTerm, beginOfInitial;
        switch (st->ek) {
        case E_SINGLE:
            endOfInitial = endOfStart + m->init_len; if (endOfInitial > right) endOfInitial = right;
            rest += exon_shortProb(x, m->xinit, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                  + exon_seqProb(x, 1, endOfInitial + 1, right, frameOfRight);
            break;
        case E_INITIAL:
            endOfInitial = endOfStart + m->init_len;
            if (endOfInitial > right) { endOfInitial = right; beginOfTerm = right + 1; }
            else { beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfInitial) beginOfTerm = right + 1; }
            rest += exon_shortProb(x, m->xinit, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                  + exon_seqProb(x, 1, endOfInitial + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                  + exon_shortProb(x, m->xet, 1, beginOfTerm, right, frameOfRight);
            break;
        case E_INTERNAL:
            beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfStart) beginOfTerm = right + 1;
            rest += exon_seqProb(x, 1, endOfStart + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                  + exon_shortProb(x, m->xet, 1, beginOfTerm, right, frameOfRight);
            break;
        case E_TERMINAL:
            rest += exon_seqProb(x, 1, endOfStart + 1, right, frameOfRight);
            break;
        case E_RSINGLE:
            beginOfInitial = beginOfInitP - m->init_len; if (beginOfInitial < bos) beginOfInitial = bos;
            rest += exon_shortProb(x, m->xinit, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                  + exon_seqProb(x, 0, bos, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)));
            break;
        case E_RINITIAL:
            beginOfInitial = beginOfInitP - m->init_len;
            if (beginOfInitial < bos) { beginOfInitial = bos; endOfTerm = bos - 1; }
            else { endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitial) endOfTerm = bos - 1; }
            rest += exon_shortProb(x, m->xinit, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                  + exon_seqProb(x, 0, endOfTerm + 1, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)))
                  + exon_shortProb(x, m->xet, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
            break;
        case E_RINTERNAL:
            endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitP) endOfTerm = bos - 1;
            rest += exon_seqProb(x, 0, endOfTerm + 1, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                  + exon_shortProb(x, m->xet, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
            break;
        default: /* E_RTERMINAL */
            rest += exon_seqProb(x, 0, bos, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)));
        }
    }
    if (isneg(rest)) return NEG;
    int eobe = right + st->innerPartEndOffset, len = eobe - bobe + 1;
    if (len < 1 || len >= m->n_ld_exon) return NEG;
    sc_t lp;
    switch (st->ek) {
    case E_SINGLE: case E_RSINGLE: lp = len % 3 == 0 ? m->ld_single[len] : NEG; break;
    case E_INITIAL: lp = (len % 3 == win && len > 2) ? m->ld_initial[len] : NEG; break;
    case E_RINITIAL: lp = len > 2 ? m->ld_initial[len] : NEG; break;
    case E_INTERNAL: case E_RINTERNAL: lp = m->ld_internal[len]; break;
    case E_TERMINAL: lp = m->ld_terminal[len]; break;
    default: lp = mod3(2 - len) == win ? m->ld_terminal[len] : NEG;
    }
    if (isneg(lp)) return NEG;
    return beginPart + rest + (m->log3 + lp);
}

/* ExonModel::viterbiForwardAndSampling, exonmodel.cc:899-1179 */
static void exon_eval(Ctx* x, int s, int j, Oli* o) {
    const Model* m = x->m; const StateInfo* st = &m->st[s]; int L = x->L, fwd = st->fwd, win = st->frame;
    o->max = NEG; o->state = -1; o->base = -1;
    sc_t ep = endPart(x, st, j);
    int eobe = j + st->baseOffset, right = eobe - st->innerPartEndOffset;
    if (isneg(ep) || right < 0) return;
    int frameOfRight = fwd ? mod3(win - (eobe + 1) + right) : mod3(win + eobe + 1 - right);
    int eons = (st->ek == E_TERMINAL || st->ek == E_SINGLE) ? eobe - 3 : eobe;
    if (eons > L - 1) eons = L - 1;
    int feons = fwd ? mod3(win - 1 - eobe + eons) : mod3(win + 1 + eobe - eons);
    int ORFleft = leftmostExonBegin(x, feons, eons, fwd);
    int startMax = eobe + st->innerPartOffset - m->min_exon_length + 1, startMin;
    if (st->ek == E_RTERMINAL || st->ek == E_RSINGLE) startMin = startMax = ORFleft + 2;
    else {
        startMin = ORFleft <= 0 ? 0 : ORFleft + st->innerPartOffset;
        if (startMax > j + st->beginPartLen) startMax = j + st->beginPartLen;
    }
    for (int bos = startMax; bos >= startMin; bos--) {
        int eop = bos - st->beginPartLen - 1;
        sc_t nep = notEndPart(x, st, bos, right, frameOfRight);
        if (isneg(nep) || eop >= L) continue;
        int col = eop >= 0 ? eop : 0;
        int bobe = bos - st->innerPartOffset, len = eobe - bobe + 1;
        for (int i = 0; i < st->nanc; i++) {
            int a = st->anc[i]; sc_t pv = PVAL(x, col, a); if (isneg(pv)) continue;
            int pf = m->st[a].frame;
            if (st->ek == E_SINGLE || st->ek == E_RSINGLE || st->ek == E_RTERMINAL || st->ek == E_INITIAL ||
                win == mod3(fwd ? pf + len : pf - len)) {
                fwd_option(x, a, col, eop, TR(a, s) + ep + nep);
                sc_t pp = pv + (TR(a, s) + ep + nep);
                if (pp > o->max) { o->max = pp; o->base = eop; o->state = a; }
            }
        }
    }
}


/* ================================================================== UTR states, utrmodel.cc */
#define UNSET ((sc_t)1 << 62)
enum { SG_RINIT5, SG_INIT5, SG_3, SG_R3, SG_INTRON, SG_5, SG_R5 };        /* UtrModel::initSnippetProbs, utrmodel.cc:701-712 */
enum { SG_NC = 7 };                                                         /* NcModel::initSnippetProbs, ncmodel.cc:95-100: intron content, forward */
static const int seg_fwd[8] = { 0, 1, 1, 0, 1, 1, 0, 1 };
static const sc_t* seg_table(const Ctx* x, int g) {                        /* the static table the SegProbs points at: current class */
    const Model* m = x->m; size_t w = (size_t)1 << (2 * (m->k + 1));
    switch (g) {
    case SG_RINIT5: case SG_INIT5: return m->u5i + x->cls * w;
    case SG_3: case SG_R3: return m->u3 + x->cls * w;
    case SG_INTRON: case SG_NC: return m->iemi + x->cls * w;
    default: return m->u5 + x->cls * w;
    }
}
static sc_t seg_emi1(const Ctx* x, int g, int i) {                         /* one factor of SegProbs::setEmiProbs, statemodel.cc:413-432 */
    const Model* m = x->m; int pn;
    if (seg_fwd[g]) pn = i < m->k ? -1 : s2i(x, i - m->k, m->k + 1);
    else pn = i >= x->L - m->k ? -1 : s2irc(x, i, m->k + 1);
    return pn < 0 ? m->log025 : seg_table(x, g)[pn];
}
/* SegProbs::setEmiProbs, statemodel.cc:398-436 */
static void seg_set(Ctx* x, int g, int from, int to) {
    int n = x->L; sc_t* cum = x->seg[g];
    if (from < 0 || to < 0) { from = 1; to = n; }
    if (from == 1) cum[0] = x->m->log025;
    if (to > n) to = n;
    if (cum[from - 1] == NEG) cum[from - 1] = 0;                          /* "previous emiprobs not computed": set to 1 */
    for (int i = from; i <= to; i++) cum[i] = cum[i - 1] + seg_emi1(x, g, i);
}
/* SegProbs::getSeqProb, statemodel.cc:441-465 */
static sc_t seg_get(const Ctx* x, int g, int from, int to) {
    const Model* m = x->m;
    if (from == to) {
        int pn;
        if (seg_fwd[g]) { if (to < m->k) return m->log025; pn = s2i(x, to - m->k, m->k + 1); }
        else pn = s2irc(x, to, m->k + 1);
        return pn < 0 ? m->log025 : seg_table(x, g)[pn];
    }
    if (from > to) return 0;
    if (to > x->L) to = x->L;
    sc_t r = x->seg[g][to];
    return from < 1 ? r : r - x->seg[g][from - 1];
}
/* UtrModel::tssupSeqProb, utrmodel.cc:1731-1750 */
static sc_t tssup(const Ctx* x, int left, int right, int reverse) {
    const Model* m = x->m; sc_t s = 0; int k = m->tssup_k; const sc_t* tab = m->tup + ((size_t)x->cls << (2 * (k + 1)));
    for (int p = right; p >= left; p--) {
        int pn = -1;
        if (!reverse && p - k >= 0) pn = s2i(x, p - k, k + 1);
        else if (reverse && p >= 0 && p + k < x->L) pn = s2irc(x, p, k + 1);
        s += pn < 0 ? m->log025 : tab[pn];
    }
    return s;
}
static inline int rawc(const Ctx* x, int p) { return (p < 0 || p >= x->L) ? 0 : x->raw[p]; }
/* UtrModel::tssProb, utrmodel.cc:1761-1833 (no hints: extrinsicProb = 1); memo keeps the class active at first evaluation */
static sc_t tssProb(Ctx* x, int fwd, int left) {
    const Model* m = x->m;
    int right = left + m->tuw + m->tss_end - 1;
    if (right >= x->L) return NEG;
    if (left % m->tts_spacing != 0) return NEG;
    sc_t* memo = x->tssP[fwd ? 0 : 1];
    if (memo[left] != UNSET) return memo[left];
    sc_t tssMotifProb, tataMotifProb = 0, up; int maxpos = m->d_tata_max - m->d_tata_min - 1;
    if (fwd) {
        int p0 = right - m->tss_end - m->d_tata_max + 1, rel = -1;        /* findTATA, utrmodel.cc:271-285 */
        for (int pos = 0; pos <= maxpos && rel < 0; pos++)
            if (rawc(x, p0 + pos) == 't' && rawc(x, p0 + pos + 1) == 'a' && rawc(x, p0 + pos + 2) == 't' && rawc(x, p0 + pos + 3) == 'a' && rawc(x, p0 + pos + 5) == 'a') rel = pos;
        int pm = right - m->tss_end - m->tss_start + 1;
        if (rel >= 0) {
            int tatapos = p0 + rel;
            tssMotifProb = motif_fwd(x, m->tsstm, m->tsstm_n, m->tsstm_k, pm);
            tataMotifProb = motif_fwd(x, m->tatam, m->tatam_n, m->tatam_k, tatapos - m->tata_start);
            up = tssup(x, left, tatapos - m->tata_start - 1, 0) + tssup(x, tatapos + m->tata_end, right - m->tss_end - m->tss_start, 0);
        } else {
            tssMotifProb = motif_fwd(x, m->tssm, m->tssm_n, m->tssm_k, pm);
            up = tssup(x, left, right - m->tss_end - m->tss_start, 0);
        }
    } else {
        int p0 = left + m->tss_end + m->d_tata_max - 1, rel = 1;
        for (int pos = 0; pos >= -maxpos && rel > 0; pos--)
            if (rawc(x, p0 + pos) == 'a' && rawc(x, p0 + pos - 1) == 't' && rawc(x, p0 + pos - 2) == 'a' && rawc(x, p0 + pos - 3) == 't' && rawc(x, p0 + pos - 5) == 't') rel = pos;
        if (rel <= 0) {
            int tatapos = p0 + rel;
            tssMotifProb = motif_rc(x, m->tsstm, m->tsstm_n, m->tsstm_k, left);
            tataMotifProb = motif_rc(x, m->tatam, m->tatam_n, m->tatam_k, tatapos - m->tata_end + 1);
            up = tssup(x, left + m->tata_end + m->tata_start - 1, tatapos - m->tata_end, 1) + tssup(x, tatapos + m->tata_start + 1, right, 1);
        } else {
            tssMotifProb = motif_rc(x, m->tssm, m->tssm_n, m->tssm_k, left);
            up = tssup(x, left + m->tss_end + m->tss_start, right, 1);
        }
    }
    sc_t prob = (isneg(tssMotifProb) || isneg(tataMotifProb) || isneg(up)) ? NEG : tssMotifProb + tataMotifProb + up;
    memo[left] = prob;
    return prob;
}
/* UtrModel::computeTtsProbs, utrmodel.cc:1840-1912 (no hints) */
static void compute_tts(Ctx* x, int from, int to) {
    const Model* m = x->m; int L = x->L;
    if (from < 0 || to < 0) { from = 0; to = L; }
    if (to > L) to = L;
    for (int b = from; b <= to; b++) {
        int ttspos = b + m->boxlen + m->dpc - 1;
        if (ttspos >= L) x->ttsP[0][b] = NEG;
        else {
            int pn = s2i(x, b, m->boxlen);
            sc_t prob = (pn < 0 || isneg(m->aataaa[pn])) ? NEG : m->aataaa[pn] + m->log_polya;
            if (b % m->tts_spacing == 0 && isneg(prob)) prob = m->log_nopolya;
            if (!isneg(prob)) prob += motif_fwd(x, m->ttsm, m->ttsm_n, m->ttsm_k, b + m->boxlen);
            x->ttsP[0][b] = prob;
        }
        ttspos = b - m->dpc;
        if (ttspos < 0 || b + m->boxlen - 1 >= L) x->ttsP[0][b] = NEG;     /* sic: the reference zeroes the PLUS entry here (:1887) */
        else {
            int pn = s2irc(x, b, m->boxlen);
            sc_t prob = (pn < 0 || isneg(m->aataaa[pn])) ? NEG : m->aataaa[pn] + m->log_polya;
            if (b % m->tts_spacing == 0 && isneg(prob)) prob = m->log_nopolya;
            if (!isneg(prob)) prob += motif_rc(x, m->ttsm, m->ttsm_n, m->ttsm_k, ttspos);
            x->ttsP[1][b] = prob;
        }
    }
}
/* UtrModel::updateToLocalGC, utrmodel.cc:768-791; x->cls is already the new class */
static void utr_update_gc(Ctx* x, int from, int to) {
    if (!x->m->utr) return;
    for (int i = from; i < to && i >= 0; i++) x->tssP[0][i] = x->tssP[1][i] = UNSET;
    compute_tts(x, from, to);
    for (int g = 0; g < 7; g++) seg_set(x, g, from, to + 5);
    if (x->m->nc) seg_set(x, SG_NC, from, to);                                /* NcModel::updateToLocalGC, ncmodel.cc:125-127 */
}

/* EOPList, statemodel.cc:473-520.  it == n plays the role of end(); dereferencing end() of a libstdc++ std::list<int>
 * reads the low word of the size field kept in the sentinel node, which is what the reference build here does in
 * update()'s first comparison when the iterator sits at end(). */
static void eop_insert(struct Eop* e, int at, int v) {
    if (e->n == e->cap) { e->cap = e->cap ? 2 * e->cap : 64; e->v = (int*)realloc(e->v, e->cap * sizeof(int)); }
    memmove(e->v + at + 1, e->v + at, (e->n - at) * sizeof(int)); e->v[at] = v; e->n++;
}
static void eop_decrement(struct Eop* e, int* endOfPred) {
    if (e->inCache && e->it != e->n && e->v[e->it] == *endOfPred && ++e->it != e->n) *endOfPred = e->v[e->it];
    else (*endOfPred)--;
}
static void eop_update(struct Eop* e, int endOfPred) {
    if (e->n == 0) { eop_insert(e, 0, endOfPred); e->it = e->n; return; }          /* eopit stays end() */
    if (endOfPred == (e->it == e->n ? e->n : e->v[e->it])) return;                 /* case A */
    if (endOfPred >= e->v[0]) { if (endOfPred > e->v[0]) eop_insert(e, 0, endOfPred); e->it = 0; return; }   /* case B */
    if (endOfPred < e->v[e->n - 1]) { eop_insert(e, e->n, endOfPred); e->it = e->n - 1; return; }              /* case C */
    if (e->it != e->n && endOfPred < e->v[e->it]) {                                 /* case D */
        int nxt = e->it + 1;
        if (nxt != e->n && endOfPred == e->v[nxt]) { e->it = nxt; e->inCache = 1; return; }
        if (nxt == e->n || endOfPred > e->v[nxt]) { eop_insert(e, nxt, endOfPred); e->it = nxt; return; }
        while (e->it != e->n && e->v[e->it] > endOfPred) e->it++;
        eop_update(e, endOfPred);
        return;
    }
    fprintf(stderr, "oracle: EOPList state the reference would throw on (Error 2 in UtrModel::updatePossibleEOPs)\n"); exit(4);
}

/* UtrModel::getEndPositions, utrmodel.cc:1572-1643 */
static void utr_end_positions(const Ctx* x, const StateInfo* st, int end, int* boe, int* eobe) {
    const Model* m = x->m; int DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    *boe = end + 1; *eobe = end;
    if (st->uk == U_INTRON || st->uk == U_INTRONVAR) return;
    if (st->u5) {
        if (st->fwd) {
            if (st->uk == U_SINGLE || st->uk == U_TERM) { *eobe = end + m->tiw; }
            else { *boe = end - DW + 1; *eobe = end - m->dss_end - 2; }                     /* init, internal */
        } else {
            if (st->uk == U_SINGLE || st->uk == U_INIT) { *boe = end - m->tuw - m->tss_end + 1; *eobe = end - m->tuw; }
            else { *boe = end - AW - m->ass_up + 1; *eobe = end - m->ass_up - m->ass_start - 2; }   /* internal, term */
        }
    } else {
        if (st->fwd) {
            if (st->uk == U_SINGLE || st->uk == U_TERM) {
                if (end != x->L - 1) { *boe = end - m->dpc - m->boxlen + 1; *eobe = end; }
                else { *boe = x->L; *eobe = x->L - 1; }
            } else { *boe = end - DW + 1; *eobe = end - m->dss_end - 2; }
        } else {
            if (st->uk == U_INTERNAL || st->uk == U_TERM) { *boe = end - AW - m->ass_up + 1; *eobe = end - m->ass_up - m->ass_start - 2; }
            /* rutr3single, rutr3init: no end signal */
        }
    }
}
/* UtrModel::endPartEmiProb, utrmodel.cc:1072-1161 (no hints) */
static sc_t utr_endPart(Ctx* x, const StateInfo* st, int begin, int end, int eobe) {
    const Model* m = x->m; int L = x->L, DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    if (st->uk == U_INTRON) return 0;
    if (st->uk == U_INTRONVAR) return NEG;            /* only reachable through intron hints (:990-1040) */
    if (st->fwd) {
        if (st->u5 && (st->uk == U_SINGLE || st->uk == U_TERM)) {
            if (eobe + 3 <= L - 1) { int c = s2i(x, eobe + 1, 3); if (c < 0 || !m->isstart[c]) return NEG; }
            return 0;
        }
        if (st->uk == U_INIT || st->uk == U_INTERNAL) return dSSProb(x, end - DW + 1, 1);
        /* utr3single, utr3term */
        if (end == L - 1) return 0;
        if (begin < 0 || begin + m->boxlen - 1 >= L) return NEG;
        return x->ttsP[0][begin];
    }
    if (st->u5) {
        if (st->uk == U_SINGLE || st->uk == U_INIT) return tssProb(x, 0, begin);
        return aSSProb(x, end - m->ass_up - AW + 1, 0);
    }
    if (st->uk == U_SINGLE || st->uk == U_INIT) return (end + 3 > L - 1 || !isRCStop(x, end + 1)) ? NEG : 0;
    return aSSProb(x, end - m->ass_up - AW + 1, 0);
}
static sc_t uld_get(const Model* m, int u5, int uk, int len) {
    if (len < 0 || len >= m->n_uld[u5][uk]) { fprintf(stderr, "oracle: UTR length %d outside the distribution (%d,%d)\n", len, u5, uk); exit(4); }
    return m->uld[u5][uk][len];
}
static sc_t tail_get(const sc_t* t, int n, int len) {
    if (len < 0 || len >= n) { fprintf(stderr, "oracle: UTR tail length %d outside the distribution\n", len); exit(4); }
    return t[len];
}
/* UtrModel::notEndPartEmiProb, utrmodel.cc:1167-1548 (no hints: extrinsicQuot = 1) */
static sc_t utr_notEndPart(Ctx* x, const StateInfo* st, int begin, int endOfMiddle, int eobe) {
    const Model* m = x->m; int L = x->L, DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    sc_t beginPart = 0, middle = 0, lenp = 0; int bom, bobe;
    if (st->uk == U_INTRON) {                          /* :1254-1265, :1387-1399 — one base, intron content of the current class */
        for (int pos = begin; pos <= endOfMiddle; pos++) {
            int pn = pos - m->k >= 0 ? s2i(x, pos - m->k, m->k + 1) : -1;
            sc_t e = pn < 0 ? m->log025 : m->iemi[((size_t)x->cls << (2 * (m->k + 1))) | pn];
            if (st->u5) middle += e; else middle = e;
        }
        return middle + NB(x, begin, endOfMiddle);                   /* :1521-1531 */
    }
    if (st->uk == U_INTRONVAR) return NEG;
    if (st->fwd && st->u5) {
        switch (st->uk) {
        case U_SINGLE: case U_INIT:
            bom = begin + m->tuw + m->tss_end; bobe = begin + m->tuw;
            if (st->uk == U_INIT || endOfMiddle - bom + 1 >= 0) middle = seg_get(x, SG_INIT5, bom, endOfMiddle);
            else middle = m->log2 * (-(endOfMiddle - bom + 1));
            lenp = uld_get(m, 1, st->uk, eobe - bobe + 1);
            if (begin >= 0) beginPart = tssProb(x, 1, begin);
            else {
                beginPart = m->log025 * (bom - 1);
                if (begin + m->tuw == 0)
                    lenp = st->uk == U_SINGLE ? tail_get(m->utl5s, m->n_utl5s, endOfMiddle - begin + 1 + m->tiw - m->tuw)
                                              : tail_get(m->utl5s, m->n_utl5s, eobe - bobe + 1);
            }
            break;
        case U_INTERNAL:
            beginPart = aSSProb(x, begin, 1); bobe = begin + m->ass_up + m->ass_start + 2;
            if (!isneg(beginPart)) { bom = begin + m->ass_up + AW; middle = seg_get(x, SG_5, bom, endOfMiddle); lenp = uld_get(m, 1, U_INTERNAL, eobe - bobe + 1); }
            break;
        default: /* U_TERM */
            bobe = begin + m->ass_up + m->ass_start + 2;
            beginPart = bobe >= L ? NEG : aSSProb(x, begin, 1);
            if (!isneg(beginPart)) {
                bom = begin + m->ass_up + AW;
                if (endOfMiddle - bom + 1 >= 0) middle = seg_get(x, SG_5, bom, endOfMiddle);
                else middle = (-m->log025) * (-(endOfMiddle - bom + 1));
                lenp = uld_get(m, 1, U_TERM, eobe - bobe + 1);
            }
        }
    } else if (!st->fwd && st->u5) {
        switch (st->uk) {
        case U_INTERNAL: case U_INIT:
            beginPart = dSSProb(x, begin, 0); bobe = begin + m->dss_end + 2;
            if (!isneg(beginPart)) { bom = begin + DW; middle = seg_get(x, st->uk == U_INIT ? SG_RINIT5 : SG_R5, bom, endOfMiddle); lenp = uld_get(m, 1, st->uk, eobe - bobe + 1); }
            break;
        default: /* rutr5term, rutr5single */
            bom = begin; bobe = begin - m->tiw;
            if (endOfMiddle - bom + 1 >= 0) middle = seg_get(x, st->uk == U_TERM ? SG_R5 : SG_RINIT5, bom, endOfMiddle);
            else middle = (st->uk == U_TERM ? -m->log025 : m->log2) * (-(endOfMiddle - bom + 1));
            lenp = uld_get(m, 1, st->uk, eobe - bobe + 1);
        }
    } else if (st->fwd) {   /* 3' forward */
        switch (st->uk) {
        case U_SINGLE:
            bom = bobe = begin; middle = seg_get(x, SG_3, bom, endOfMiddle);
            lenp = eobe != L - 1 ? uld_get(m, 0, U_SINGLE, eobe - bobe + 1) : tail_get(m->utl3s, m->n_utl3s, eobe - bobe + 1);
            break;
        case U_INIT:
            bom = bobe = begin;
            if (endOfMiddle - bom + 1 >= 0) middle = seg_get(x, SG_3, bom, endOfMiddle);
            else middle = (-m->log025) * (-(endOfMiddle - bom + 1));
            lenp = uld_get(m, 0, U_INIT, eobe - bobe + 1);
            break;
        default: /* internal, term */
            beginPart = aSSProb(x, begin, 1); bobe = begin + m->ass_up + m->ass_start + 2;
            if (!isneg(beginPart)) {
                bom = begin + m->ass_up + AW; middle = seg_get(x, SG_3, bom, endOfMiddle);
                if (st->uk == U_INTERNAL || eobe != L - 1) lenp = uld_get(m, 0, st->uk, eobe - bobe + 1);
                else lenp = tail_get(m->utl3s, m->n_utl3s, eobe - bobe + 1);
            }
        }
    } else {                /* 3' reverse */
        switch (st->uk) {
        case U_SINGLE: case U_TERM:
            bobe = begin; bom = begin + m->boxlen + m->dpc;
            if (begin > 0) { beginPart = x->ttsP[1][begin + m->dpc]; if (st->uk == U_SINGLE) lenp = uld_get(m, 0, U_SINGLE, eobe - bobe + 1); }
            else {
                beginPart = (st->uk == U_TERM || bom > 0) ? m->log025 * (bom - 1) : 0;
                if (st->uk == U_SINGLE) lenp = tail_get(m->utl3s, m->n_utl3s, eobe - bobe + 1);
            }
            if (!isneg(beginPart)) { middle = seg_get(x, SG_R3, bom, endOfMiddle); if (st->uk == U_TERM) lenp = uld_get(m, 0, U_TERM, eobe - bobe + 1); }
            break;
        default: /* rutr3init, rutr3internal */
            beginPart = dSSProb(x, begin, 0); bobe = begin + m->dss_end + 2;
            if (!isneg(beginPart)) {
                bom = begin + DW;
                if (st->uk == U_INTERNAL || endOfMiddle - bom + 1 >= 0) middle = seg_get(x, SG_R3, bom, endOfMiddle);
                else middle = (-m->log025) * (-(endOfMiddle - bom + 1));
                lenp = uld_get(m, 0, st->uk, eobe - bobe + 1);
            }
        }
    }
    if (isneg(beginPart) || isneg(middle) || isneg(lenp)) return NEG;
    /* intron positions in front of the biological exon (utrmodel.cc:1532-1545) */
    sc_t nb = 0;
    if ((st->fwd && (st->uk == U_INTERNAL || st->uk == U_TERM)) || (!st->fwd && (st->uk == U_INTERNAL || st->uk == U_INIT))) nb = NB(x, begin, bobe - 1);
    return beginPart + middle + lenp + nb;
}
/* UtrModel::viterbiForwardAndSampling, utrmodel.cc:796-1064 */
static void utr_eval(Ctx* x, int s, int j, Oli* o) {
    const Model* m = x->m; const StateInfo* st = &m->st[s]; int L = x->L, base = j;
    int DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    o->max = NEG; o->state = -1; o->base = -1;
    int boe, eobe, lm, rm;
    utr_end_positions(x, st, base, &boe, &eobe);
    switch (st->uk) {
    case U_SINGLE:
        if (st->u5) {
            lm = base - (m->umax - m->tiw + m->tuw);
            rm = st->fwd ? base - m->tuw - m->tss_end - 1 + m->tiw + m->tss_end : base - m->tuw - 1 + m->tiw;
            if (rm > base - 1) rm = base - 1;
        } else {
            lm = base - m->umax3s;
            rm = (st->fwd && base == L - 1) ? base - 1 : base - m->dpc - m->boxlen;
        }
        break;
    case U_INIT:
        if (st->u5) { lm = base - (m->umax + 2 + m->dss_end + m->tuw); rm = base - m->tuw - m->tss_end - DW; }
        else { lm = base - (m->umax + 2 + m->dss_end); rm = base - m->dss_end - 2; }
        break;
    case U_INTERNAL:
        lm = base - (m->umax + 2 + m->dss_end + m->ass_up + m->ass_start + 2); rm = base - DW - m->ass_up - AW;
        break;
    case U_TERM:
        if (st->u5) {
            lm = base - (m->umax - m->tiw + m->ass_up + m->ass_start + 2); rm = base - m->ass_up - AW;
            if (-m->ass_up - AW + m->tiw + m->ass_end < 0) rm = base - m->ass_up - AW + m->tiw + m->ass_end;
        } else {
            lm = base - (m->umax3t + 2 + m->ass_start + m->ass_up);
            rm = (st->fwd && base == L - 1) ? base - AW - m->ass_up : base - m->dpc - m->boxlen - AW - m->ass_up;
        }
        break;
    default: lm = rm = base - 1;
    }
    sc_t ep = boe >= 0 ? utr_endPart(x, st, boe, base, eobe) : NEG;
    if (isneg(ep)) return;
    if (st->uk == U_INTRONVAR) return;
    /* intron positions handled inside an exon state behind its biological end (utrmodel.cc:1143-1158) */
    if (st->uk != U_INTRON && eobe < base && !(st->fwd && !st->u5 && (st->uk == U_SINGLE || st->uk == U_TERM)) && !(!st->fwd && st->u5 && (st->uk == U_SINGLE || st->uk == U_INIT)))
        ep += NB(x, eobe + 1, base);
    if (st->fwd && st->u5 && (st->uk == U_SINGLE || st->uk == U_INIT)) { if (lm < -m->tuw) lm = -m->tuw; }
    else if (!st->fwd && !st->u5 && (st->uk == U_SINGLE || st->uk == U_TERM)) { if (lm < -m->boxlen - m->dpc) lm = -m->boxlen - m->dpc; }
    else if (lm < 0) lm = 0;
    struct Eop* e = &x->eop[s];
    if (x->walking) e->n = 0;
    e->it = 0; e->inCache = 0;
    for (int endOfPred = rm; endOfPred >= lm; x->eop_off ? (void)endOfPred-- : eop_decrement(e, &endOfPred)) {
        int col = endOfPred > 0 ? endOfPred : 0, i0 = 0;
        while (i0 < st->nanc && isneg(PVAL(x, col, st->anc[i0]))) i0++;
        if (i0 == st->nanc) continue;
        sc_t nep = utr_notEndPart(x, st, endOfPred + 1, boe - 1, eobe);
        if (isneg(nep)) continue;
        if (!x->eop_off) eop_update(e, endOfPred);
        for (int i = i0; i < st->nanc; i++) {
            int a = st->anc[i]; sc_t pv = PVAL(x, col, a); if (isneg(pv)) continue;
            sc_t te = TR(a, s) + (nep + ep);
            fwd_option(x, a, col, endOfPred, te);
            sc_t pp = pv + te;
            if (pp > o->max) { o->max = pp; o->state = a; o->base = endOfPred; }
        }
    }
}

/* ------------------------------------------------------------------ NcModel (ncmodel.cc), --nc=on, no hints
 *
 * Transcript boundaries of non-coding genes exist only where hints put them (precomputeTxEndProbs, ncmodel.cc:744-826: without tss /
 * tts / exonpart hints all four tss / tts arrays are zero), so ab initio the states that begin or end with a transcript boundary
 * (ncsingle, ncinit, ncterm, rncsingle, rncinit, rncterm) never hold a cell beyond column 0, the hint-only ncintronvar / rncintronvar
 * neither; ncintron / rncintron (one base, self loop) and ncinternal / rncinternal (splice site to splice site) live on the initial
 * probabilities of column 0.  The evaluation order, the aSSProb calls (they feed the memo shared with IntronModel) and the EOPList of the
 * reference are kept for every state. */
static void nc_end_positions(const Ctx* x, const StateInfo* st, int end, int* boe, int* eobe) {     /* NcModel::getEndPositions :702-725 */
    const Model* m = x->m; int DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    *boe = end + 1; *eobe = end;
    if (st->fwd && (st->uk == U_INTERNAL || st->uk == U_INIT)) { *boe = end - DW + 1; *eobe = end - m->dss_end - 2; }
    else if (!st->fwd && (st->uk == U_INTERNAL || st->uk == U_TERM)) { *boe = end - AW - m->ass_up + 1; *eobe = end - m->ass_up - m->ass_start - 2; }
}
static sc_t nc_endPart(Ctx* x, const StateInfo* st, int begin, int end) {                              /* NcModel::endPartEmiProb :366-441, no hints */
    const Model* m = x->m;
    if (st->uk == U_INTRON) return 0;
    if (st->uk == U_INTRONVAR) {
        if (st->fwd) return possASS(x, end + m->ass_up + m->ass_start + 2) ? 0 : NEG;
        return possRDSS(x, end + m->dss_end + 2) ? 0 : NEG;
    }
    if (st->fwd) {
        if (st->uk == U_SINGLE || st->uk == U_TERM) return NEG;                 /* ttsProbPlus[end] = 0 without hints */
        return dSSProb(x, begin, 1);                                            /* ncinit, ncinternal */
    }
    if (st->uk == U_SINGLE || st->uk == U_INIT) return NEG;                     /* tssProbMinus[end] = 0 without hints */
    return aSSProb(x, begin, 0);                                                /* rncterm, rncinternal */
}
static sc_t nc_notEndPart(Ctx* x, const StateInfo* st, int begin, int endOfMiddle, int eobe) {        /* NcModel::notEndPartEmiProb :447-676, no hints */
    const Model* m = x->m; int DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    sc_t beginPart, middle = 0; int bobe, bom;
    if (st->uk == U_INTRON) {                                                   /* :528-539: content of the base(s), current class */
        for (int pos = begin; pos <= endOfMiddle; pos++) {
            int pn = pos - m->k >= 0 ? s2i(x, pos - m->k, m->k + 1) : -1;
            middle += pn < 0 ? m->log025 : m->iemi[((size_t)x->cls << (2 * (m->k + 1))) | pn];
        }
        return middle;
    }
    if (st->uk == U_INTRONVAR) return NEG;                                      /* (never asked: the caller handles hint introns only) */
    if (st->fwd) {
        if (st->uk == U_SINGLE || st->uk == U_INIT) return NEG;                 /* tssProbPlus[begin] = 0 */
        bobe = begin + m->ass_up + m->ass_start + 2;
        if (st->uk == U_TERM && bobe >= x->L) return NEG;
        beginPart = aSSProb(x, begin, 1);                                       /* ncinternal :471-479, ncterm :480-495 */
        if (isneg(beginPart)) return NEG;
        bom = begin + m->ass_up + AW;
        if (st->uk == U_TERM && endOfMiddle - bom + 1 < 0) middle = -m->log025 * (sc_t)(-(endOfMiddle - bom + 1));
        else middle = seg_get(x, SG_NC, bom, endOfMiddle);
    } else {
        if (st->uk == U_SINGLE || st->uk == U_TERM) return NEG;                 /* ttsProbMinus[begin] = 0 */
        beginPart = dSSProb(x, begin, 0);                                       /* rncinternal :505-513, rncinit :526-534 */
        if (isneg(beginPart)) return NEG;
        bobe = begin + m->dss_end + 2; bom = begin + DW;
        middle = seg_get(x, SG_NC, bom, endOfMiddle);
    }
    int len = eobe - bobe + 1;
    if (len < 0 || len >= m->n_ld_exon) { fprintf(stderr, "oracle: nc exon length %d outside the distribution\n", len); exit(4); }
    sc_t lp = m->ld_internal[len];                                              /* NcModel::lenDistInternal = ExonModel::lenDistInternal :146 */
    if (isneg(lp) || isneg(middle)) return NEG;
    return beginPart + middle + lp;
}
static void nc_eval(Ctx* x, int s, int j, Oli* o) {                                                    /* NcModel::viterbiForwardAndSampling :154-359 */
    const Model* m = x->m; const StateInfo* st = &m->st[s]; int base = j;
    int DW = m->dss_start + m->dss_end + 2, AW = m->ass_start + m->ass_end + 2;
    o->max = NEG; o->state = -1; o->base = -1;
    int boe, eobe, lm, rm;
    nc_end_positions(x, st, base, &boe, &eobe);
    switch (st->uk) {
    case U_SINGLE: lm = base - m->max_exon_len; rm = base - 1; break;
    case U_INIT: lm = base - (m->max_exon_len + 2 + m->dss_end); rm = base - DW; break;
    case U_INTERNAL: lm = base - (m->max_exon_len + 2 + m->dss_end + m->ass_up + m->ass_start + 2); rm = base - 2 - m->dss_end - m->ass_up - m->ass_start - 2 - 1; break;
    case U_TERM: lm = base - (m->max_exon_len + m->ass_up + m->ass_start + 2); rm = base - m->ass_up - AW; break;
    default: lm = rm = base - 1;
    }
    sc_t ep = boe >= 0 ? nc_endPart(x, st, boe, base) : NEG;
    if (isneg(ep)) return;
    if (st->uk == U_INTRONVAR) return;                                          /* only through intron hints (:286-330) */
    if (lm < 0) lm = 0;
    if (st->uk != U_INTRON) {                                                   /* no exon hints: at most 200 bases back (:219-235) */
        int minE = rm + 1;
        if (minE > lm) { lm = minE; if (lm > rm - 200) { lm = rm - 200; if (lm < 0) lm = 0; } }
    }
    struct Eop* e = &x->eop[s];
    const int use_eop = st->uk != U_INTRON && !x->eop_off;                      /* (leftMost == rightMost for the one-base introns: the list never matters) */
    if (x->walking) e->n = 0;
    e->it = 0; e->inCache = 0;
    for (int endOfPred = rm; endOfPred >= lm; use_eop ? eop_decrement(e, &endOfPred) : (void)endOfPred--) {
        int col = endOfPred > 0 ? endOfPred : 0, i0 = 0;
        while (i0 < st->nanc && isneg(PVAL(x, col, st->anc[i0]))) i0++;
        if (i0 == st->nanc) continue;
        sc_t nep = nc_notEndPart(x, st, endOfPred + 1, boe - 1, eobe);
        if (isneg(nep)) continue;
        if (use_eop) eop_update(e, endOfPred);
        for (int i = i0; i < st->nanc; i++) {
            int a = st->anc[i]; sc_t pv = PVAL(x, col, a); if (isneg(pv)) continue;
            sc_t te = TR(a, s) + (nep + ep);                                    /* (malus(intronF) = 1 without a hints file, :266-268) */
            fwd_option(x, a, col, endOfPred, te);
            sc_t pp = pv + te;
            if (pp > o->max) { o->max = pp; o->state = a; o->base = endOfPred; }
        }
    }
}

static void state_eval(Ctx* x, int s, int j, Oli* o) {
    switch (x->m->st[s].kind) {
    case K_IGENIC: igenic_eval(x, s, j, o); break;
    case K_EXON: exon_eval(x, s, j, o); break;
    case K_UTR: utr_eval(x, s, j, o); break;
    case K_NC: nc_eval(x, s, j, o); break;
    default: intron_eval(x, s, j, o);
    }
}

/* ------------------------------------------------------------------ driver */
static int trunc_flag(int type, int end, int predEnd, int L) {      /* gene.cc:309-321 */
    int utrIntron = type == 26 || type == 27 || type == 32 || type == 33 || type == 61 || type == 62 || type == 67 || type == 68;
    int isIntron = (type >= T_LESSD0 && type <= 23) || (type >= T_RLESSD0 && type <= 58) || utrIntron;
    int utrExon = ((type >= 24 && type <= 35) || (type >= 59 && type <= 70)) && !utrIntron;
    int t = 0;
    if (end == L - 1 && ((type >= 2 && type <= 7) || (type >= 38 && type <= 43) || isIntron || type == 30 || type == 35)) t |= 2;   /* TRUNC_RIGHT */
    if ((predEnd == -1 || predEnd == 0) && ((type >= 5 && type <= 8) || (type >= 37 && type <= 40) || isIntron || utrExon)) t |= 1; /* TRUNC_LEFT */
    return t;
}

/* stable descending sort of the options by probability (OptionsList::prepareSampling = list::sort with
 * operator< "larger probability first", vitmatrix.hh:760-796) */
typedef struct { double lp; int idx; } OptKey;
static int optcmp(const void* a, const void* b) {
    const OptKey* p = (const OptKey*)a; const OptKey* q = (const OptKey*)b;
    if (p->lp > q->lp) return -1;
    if (p->lp < q->lp) return 1;
    return p->idx - q->idx;
}
/* OptionsList::sample (vitmatrix.cc:295-320) on x->opts; returns the chosen index, *lognorm = ln(prob / cumprob) */
static int opt_sample(Ctx* x, double* lognorm) {
    int n = x->nopt; if (n == 0) return -1;
    double mx = -INFINITY;
    for (int i = 0; i < n; i++) if (x->opts[i].lp > mx) mx = x->opts[i].lp;
    double cum = 0;                                  /* cumprob accumulates in insertion order (OptionsList::add) */
    for (int i = 0; i < n; i++) cum += exp(x->opts[i].lp - mx);
    OptKey* k = (OptKey*)malloc(n * sizeof *k);
    for (int i = 0; i < n; i++) { k[i].lp = x->opts[i].lp; k[i].idx = i; }
    qsort(k, n, sizeof *k, optcmp);
    double z = ((double)rand() / RAND_MAX) * cum * 0.99999, cs = 0; int pick = -1;
    for (int i = 0; i < n && pick < 0; i++) { cs += exp(k[i].lp - mx); if (z < cs) pick = k[i].idx; }
    if (pick < 0) pick = k[0].idx;                  /* "Sampling Error which should not happen": first option */
    *lognorm = x->opts[pick].lp - mx - log(cum);
    free(k);
    return pick;
}

/* condense (StatePath::condenseStatePath, gene.cc:977-1000) a left-to-right list in place; returns the new length */
static int condense_path(int n, int* t, int* b, int* e, int* tr) {
    int o = 0;
    for (int i = 0; i < n; i++) {
        int coding = (t[i] >= 1 && t[i] <= 8) || (t[i] >= 36 && t[i] <= 43);
        if (o > 0 && t[o - 1] == t[i] && !coding) { e[o - 1] = e[i]; tr[o - 1] |= tr[i]; }
        else { t[o] = t[i]; b[o] = b[i]; e[o] = e[i]; tr[o] = tr[i]; o++; }
    }
    return o;
}

/*
 * Decode one window.  dna: ASCII, any case.  gc_in: per-position class or NULL (computed here).
 * Vout: optional [L][S] int64 Q40 scores (NEG = absent).  Path arrays (capacity cap) are filled in
 * left-to-right order, one entry per backtrace step exactly as NAMGene::getViterbiPath pushes them.
 * nsample > 1: the forward matrix is filled as well (Fout optional, [L][S] ln values, -inf = absent) and nsample-1 paths
 * are sampled as NAMGene::findGenes does (namgene.cc:833-871) with glibc rand() restarted at seed 1; they are returned
 * CONDENSED, concatenated in s_* (capacity scap), s_count[i] states and s_logp[i] = ln pathemiProb for sample i.
 * Returns number of Viterbi path states, or <0: -1 no feasible path, -2 stuck, -3 capacity.
 */
int orc_decode(const Model* m, const char* dna, int L, const int* gc_in, int64_t* Vout, int* gc_out,
               int cap, int* ptype, int* pbegin, int* pend, int* ptrunc, double* logp,
               int nsample, double* Fout, int scap, int* s_type, int* s_begin, int* s_end, int* s_trunc, int* s_count, double* s_logp) {
    Ctx X; memset(&X, 0, sizeof X); Ctx* x = &X;
    x->m = m; x->L = L;
    uint8_t* c = (uint8_t*)malloc(L + 8);
    int anynuc = 0;
    for (int i = 0; i < L; i++) {
        switch (dna[i]) { case 'a': case 'A': c[i] = 0; break; case 'c': case 'C': c[i] = 1; break;
                          case 'g': case 'G': c[i] = 2; break; case 't': case 'T': c[i] = 3; break; default: c[i] = 4; }
        anynuc |= c[i] < 4;
    }
    x->c = c;
    int* gc = (int*)malloc(L * sizeof(int));
    if (gc_in) memcpy(gc, gc_in, L * sizeof(int)); else orc_gc_stairs(m, c, L, gc);
    if (gc_out) memcpy(gc_out, gc, L * sizeof(int));
    x->gc = gc;
    orf_init(x);
    x->snF = (SnipEnt**)calloc((size_t)2 * L, sizeof(SnipEnt*)); x->snL = (SnipEnt**)calloc((size_t)2 * L, sizeof(SnipEnt*));
    x->V = (sc_t*)malloc((size_t)L * m->S * sizeof(sc_t));
    for (size_t i = 0; i < (size_t)L * m->S; i++) x->V[i] = NEG;      /* columns not reached yet are empty (viterbi.assign, namgene.cc:178) */
    const int fwd = nsample > 1;
    if (fwd) { x->F = (double*)malloc((size_t)L * m->S * sizeof(double)); for (size_t i = 0; i < (size_t)L * m->S; i++) x->F[i] = -INFINITY; }
    for (int s = 0; s < m->S; s++) { x->V[s] = m->init[s]; if (fwd) x->F[s] = isneg(m->init[s]) ? -INFINITY : sc2d(m->init[s]); }
    Oli o;
    x->mode = fwd ? 1 : 0;
    char* raw = (char*)malloc(L + 1);
    for (int i = 0; i < L; i++) raw[i] = (dna[i] >= 'A' && dna[i] <= 'Z') ? dna[i] + 32 : dna[i];
    raw[L] = 0; x->raw = raw; x->cur_gc = -1; x->walking = 0;
    x->pmask = (int*)malloc((size_t)(L + 1) * sizeof(int)); x->pmask[0] = 0;
    for (int i = 0; i < L; i++) x->pmask[i + 1] = x->pmask[i] + (dna[i] >= 'a' && dna[i] <= 'z');
    x->eop_off = getenv("ORC_EOP_OFF") != NULL;
    for (int g = 0; g < 2; g++) { x->assMemo[g] = (sc_t*)malloc((size_t)(L + 1) * sizeof(sc_t)); x->assMemoGen[g] = (int*)calloc(L + 1, sizeof(int)); }
    x->assGen = 1; x->assN = 0;
    if (m->utr) {
        for (int g = 0; g < (m->nc ? 8 : 7); g++) { x->seg[g] = (sc_t*)malloc((size_t)(L + 1) * sizeof(sc_t)); for (int i = 0; i <= L; i++) x->seg[g][i] = NEG; }
        for (int g = 0; g < 2; g++) {
            x->tssP[g] = (sc_t*)malloc((size_t)(L + 1) * sizeof(sc_t)); x->ttsP[g] = (sc_t*)malloc((size_t)(L + 1) * sizeof(sc_t));
            for (int i = 0; i <= L; i++) { x->tssP[g][i] = UNSET; x->ttsP[g][i] = NEG; }
        }
        x->eop = (struct Eop*)calloc(m->S, sizeof(struct Eop));
    }
    if (!anynuc) {   /* namgene.cc:205-226 */
        for (int j = 1; j < L; j++) for (int s = 0; s < m->S; s++) {
            VV(j, s) = s == 0 ? VV(j - 1, s) + m->log025 : NEG;
            if (fwd) FF(j, s) = s == 0 ? FF(j - 1, s) + sc2d(m->log025) : -INFINITY;
        }
    } else {
        for (int j = 1; j < L; j++) {
            x->cls = gc[j];
            if (gc[j] != x->cur_gc) {            /* namgene.cc:245-248: updateToLocalGCEach(idx, j, nextStep(j) - 1) */
                int nx = j + 1; while (nx < L && gc[nx] == gc[j]) nx++;
                x->cur_gc = gc[j]; utr_update_gc(x, j, nx - 1);
            }
            for (int s = 0; s < m->S; s++) {
                x->lse_m = -INFINITY; x->lse_s = 0;
                state_eval(x, s, j, &o); VV(j, s) = o.max;
                if (fwd) FF(j, s) = x->lse_s > 0 ? x->lse_m + log(x->lse_s) : -INFINITY;
            }
        }
    }
    if (Vout) memcpy(Vout, x->V, (size_t)L * m->S * sizeof(sc_t));
    if (Fout && fwd) memcpy(Fout, x->F, (size_t)L * m->S * sizeof(double));
    /* getViterbiPath, namgene.cc:432-510 */
    x->mode = 0; x->walking = 1;
    int ret = 0, state = -1; sc_t best = NEG;
    for (int s = 0; s < m->S; s++) {
        sc_t v = VV(L - 1, s); if (isneg(v) || isneg(m->term[s])) continue;
        v += m->term[s]; if (v > best) { best = v; state = s; }
    }
    if (state < 0) ret = -1;
    else {
        if (logp) *logp = ldexp((double)best, -FRAC_BITS);
        int base = L - 1, n = 0;
        /* collect right-to-left, then reverse */
        while (base > 0) {
            if (!anynuc) { o.state = 0; o.base = base - 1; }
            else {
                x->cls = gc[base];
                if (gc[base] != x->cur_gc) { x->cur_gc = gc[base]; utr_update_gc(x, 2, 1); }      /* namgene.cc:476-481 */
                state_eval(x, state, base, &o);
            }
            if (o.state < 0 || (o.base >= base && o.state == state) || o.base > base + 10) { ret = -2; break; }
            if (n >= cap) { ret = -3; break; }
            ptype[n] = m->st[state].type; pbegin[n] = o.base + 1; pend[n] = base;
            ptrunc[n] = trunc_flag(m->st[state].type, base, o.base, L);
            n++; base = o.base; state = o.state;
        }
        if (ret == 0) {
            for (int i = 0; i < n / 2; i++) {
                int t;
                t = ptype[i]; ptype[i] = ptype[n - 1 - i]; ptype[n - 1 - i] = t;
                t = pbegin[i]; pbegin[i] = pbegin[n - 1 - i]; pbegin[n - 1 - i] = t;
                t = pend[i]; pend[i] = pend[n - 1 - i]; pend[n - 1 - i] = t;
                t = ptrunc[i]; ptrunc[i] = ptrunc[n - 1 - i]; ptrunc[n - 1 - i] = t;
            }
            ret = n;
        }
    }
    /* sampling, NAMGene::getSampledPath namgene.cc:367-426; the process-wide rand() stream starts at seed 1 */
    if (fwd && ret >= 0 && s_count) {
        srand(1);
        int used = 0;
        int* tt = (int*)malloc((size_t)(L + 8) * 4 * sizeof(int)); int *tb = tt + L + 8, *te = tb + L + 8, *ttr = te + L + 8;
        for (int it = 0; it < nsample - 1; it++) {
            double lpath = 0; int n = 0, bad = 0;
            if (!anynuc) { tt[0] = m->st[0].type; tb[0] = 0; te[0] = L - 1; ttr[0] = 0; n = 1; lpath = L * log(0.25); }
            else {
                x->mode = 2; x->nopt = 0;
                for (int s = 0; s < m->S; s++) {
                    double f = FF(L - 1, s);
                    if (f > -1e300 && !isneg(m->term[s])) {
                        if (x->nopt == x->capopt) { x->capopt = x->capopt ? 2 * x->capopt : 256; x->opts = realloc(x->opts, x->capopt * sizeof *x->opts); }
                        x->opts[x->nopt].state = s; x->opts[x->nopt].base = L - 1; x->opts[x->nopt].lp = f + sc2d(m->term[s]); x->nopt++;
                    }
                }
                double ln; int pick = opt_sample(x, &ln);
                if (pick < 0) { bad = 1; }
                else {
                    int base = x->opts[pick].base, st = x->opts[pick].state; lpath += ln;
                    while (base > 0 && !bad) {
                        x->cls = gc[base]; x->nopt = 0;
                        if (gc[base] != x->cur_gc) { x->cur_gc = gc[base]; utr_update_gc(x, 2, 1); }  /* namgene.cc:401-406 */
                        state_eval(x, st, base, &o);
                        pick = opt_sample(x, &ln);
                        if (pick < 0 || n >= L + 8) { bad = 1; break; }
                        int ob = x->opts[pick].base, os = x->opts[pick].state;
                        tt[n] = m->st[st].type; tb[n] = ob + 1; te[n] = base; ttr[n] = trunc_flag(m->st[st].type, base, ob, L); n++;
                        base = ob; st = os; lpath += ln;
                    }
                }
            }
            if (bad) { s_count[it] = -1; s_logp[it] = 0; continue; }
            for (int i = 0; i < n / 2; i++) {
                int t;
                t = tt[i]; tt[i] = tt[n - 1 - i]; tt[n - 1 - i] = t; t = tb[i]; tb[i] = tb[n - 1 - i]; tb[n - 1 - i] = t;
                t = te[i]; te[i] = te[n - 1 - i]; te[n - 1 - i] = t; t = ttr[i]; ttr[i] = ttr[n - 1 - i]; ttr[n - 1 - i] = t;
            }
            n = condense_path(n, tt, tb, te, ttr);
            if (used + n > scap) { s_count[it] = -3; continue; }
            memcpy(s_type + used, tt, n * sizeof(int)); memcpy(s_begin + used, tb, n * sizeof(int));
            memcpy(s_end + used, te, n * sizeof(int)); memcpy(s_trunc + used, ttr, n * sizeof(int));
            s_count[it] = n; s_logp[it] = lpath; used += n;
        }
        free(tt);
    }
    for (int cl = 0; cl < 16; cl++) {
        for (int p = 0; p < 3; p++) { free(x->PX[cl][p]); free(x->PXR[cl][p]); }
        free(x->PI[cl]); free(x->PIR[cl]);
    }
    for (size_t i = 0; i < (size_t)2 * L; i++) { SnipEnt* t = x->snF[i]; while (t) { SnipEnt* nx = t->next; free(t); t = nx; } }
    if (m->utr) {
        for (int g = 0; g < 7; g++) free(x->seg[g]);
        for (int g = 0; g < 2; g++) { free(x->tssP[g]); free(x->ttsP[g]); }
        for (int s = 0; s < m->S; s++) free(x->eop[s].v);
        free(x->eop);
    }
    free(raw); free(x->pmask); for (int g = 0; g < 2; g++) { free(x->assMemo[g]); free(x->assMemoGen[g]); }
    free(x->snF); free(x->snL); free(x->opts); free(x->F);
    free(x->V); free(x->nsf); free(x->nsr); free(gc); free(c);
    return ret;
}

int orc_viterbi(const Model* m, const char* dna, int L, const int* gc_in, int64_t* Vout, int* gc_out,
                int cap, int* ptype, int* pbegin, int* pend, int* ptrunc, double* logp) {
    return orc_decode(m, dna, L, gc_in, Vout, gc_out, cap, ptype, pbegin, pend, ptrunc, logp, 0, NULL, 0, NULL, NULL, NULL, NULL, NULL, NULL);
}
