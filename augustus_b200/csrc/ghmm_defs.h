/*
 * ghmm_defs.h — shared definitions of the B200 GHMM decoder (host + device).
 *
 * Number system: log probabilities in Q23.40 fixed point (int64).  Sums are exact and associative,
 * so prefix-sum differences, lazily evaluated self-loop chains and warp-parallel reductions give
 * bit-identical results in any evaluation order.  SC_NEG stands for probability zero, which is a
 * structural value in the reference ("state impossible", e.g. exonmodel.cc:1083).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AUGB_HD __host__ __device__ __forceinline__
#define AUGB_D __device__ __forceinline__
#define AUGB_DN __device__ __noinline__      /* large routines: one copy in the kernel (instruction-cache footprint) */
#define AUGB_HDN __host__ __device__ __noinline__
#ifdef AUGB_VAR_EMIT_INLINE
#define AUGB_EMIT __device__ __forceinline__
#else
#define AUGB_EMIT __device__ __noinline__
#endif
#else
#define AUGB_HDN inline
#define AUGB_EMIT inline
#define AUGB_HD inline
#define AUGB_D inline
#define AUGB_DN inline
#endif

/* keep loops rolled: the sweep is instruction-fetch bound (one warp walks a long, branchy path per column), so
 * static code size matters more than loop overhead */
#if defined(__CUDACC__)
#define AUGB_ROLLED _Pragma("unroll 1")
#else
#define AUGB_ROLLED
#endif

namespace augb {

typedef int64_t sc_t;
constexpr int FRAC_BITS = 40;
constexpr sc_t SC_NEG = -((sc_t)1 << 61);
constexpr sc_t SC_NEGT = -((sc_t)1 << 60);
AUGB_HD bool isneg(sc_t x) { return x <= SC_NEGT; }
/* mod3 of types.hh:523 for |k| < 3*2^28: one unsigned remainder instead of two signed ones */
AUGB_HD int mod3(int k) { return (int)((unsigned)(k + 3 * (1 << 28)) % 3u); }

constexpr int MAXS = 96;      /* states */
constexpr int MAXC = 16;      /* GC classes (chicken and maize use 10) */
constexpr int MAXANC = 12;   /* (the intergenic state of the 83-state model has 10) */
constexpr int NCHAIN = 13;    /* igenic + 3 geometric + 3 reverse geometric + 4 UTR introns (utr5, utr3, rutr5, rutr3) + ncintron, rncintron */
constexpr int CH_UTR = 7;     /* first UTR-intron chain */
constexpr int CH_NC = 11;     /* ncintron, rncintron (NcModel, --nc=on) */
constexpr int WF_ALLN = 1 << 24;      /* (window flags live above the class-presence bits 0 .. MAXC-1) */
constexpr int WF_NOSLAB = 1 << 25;    /* the prefix-array pool ran out: decode this window again with the generous layout */

/* reference StateType values (include/types.hh:492-512) used by the kernels */
enum : int {
    T_IGENIC = 0, T_SINGLE = 1, T_INITIAL0 = 2, T_INTERNAL0 = 5, T_TERMINAL = 8,
    T_LESSD0 = 9, T_LONGDSS0 = 10, T_EQUALD0 = 11, T_GEO0 = 12, T_LONGASS0 = 13,
    T_RSINGLE = 36, T_RINITIAL = 37, T_RINTERNAL0 = 38, T_RTERMINAL0 = 41,
    T_RLESSD0 = 44, T_RLONGDSS0 = 45, T_REQUALD0 = 46, T_RGEO0 = 47, T_RLONGASS0 = 48,
    T_UTR5SINGLE = 24, T_UTR3SINGLE = 30, T_UTR3TERM = 35, T_RUTR5SINGLE = 59, T_RUTR3TERM = 70,
    T_NCSINGLE = 74, T_RNCSINGLE = 80      /* ncsingle, ncinit, ncintron, ncintronvar, ncinternal, ncterm; then the reverse six */
};
enum : int { K_IGENIC, K_EXON, K_LESSD, K_LONGDSS, K_EQUALD, K_GEO, K_LONGASS, K_UTR,
             K_DEAD /* nc states that need a transcript boundary: zero without hints, only their column-0 cell exists (ncmodel.cc:744-826) */ };
enum : int { U_SINGLE, U_INIT, U_INTRON, U_INTRONVAR, U_INTERNAL, U_TERM };      /* order of the utr5.. / utr3.. types (types.hh:498-499) */
enum : int { E_SINGLE, E_INITIAL, E_INTERNAL, E_TERMINAL, E_RSINGLE, E_RINITIAL, E_RINTERNAL, E_RTERMINAL };

struct StateDesc {
    int16_t type;
    int8_t kind, fwd, frame, ek;     /* ek: exon kind (E_*), or index into DevModel::ud for UTR states */
    int8_t uk, u5;                   /* UTR kind (U_*), 1 = 5' UTR */
    int8_t beginPartLen, innerPartOffset, baseOffset, innerPartEndOffset;   /* exonmodel.cc:231-279 */
    int8_t chain;          /* chain id if this is a self-loop chain state, else -1 */
    int8_t feeds;          /* chain id this state is an ancestor of (one-base successor), else -1 */
    int8_t nanc;
    int8_t anc[MAXANC];    /* ancestors in state-index order (statemodel.cc:41-46) */
};

/* activity mask bits, one uint16 per column, computed by the prep pass from the sequence alone */
enum : unsigned {
    MB_LONGDSS = 1u << 0, MB_LESSD = 1u << 1, MB_LONGASS = 1u << 2, MB_XDSS = 1u << 3, MB_XSTOP = 1u << 4,
    MB_RLONGASS = 1u << 5, MB_RLESSD = 1u << 6, MB_RLONGDSS = 1u << 7, MB_XRASS = 1u << 8, MB_XRSTART = 1u << 9,
    MB_SLOW = 1u << 10,     /* a GC-class boundary is near: lessD emissions go through the restated SnippetProbs memo */
    /* UTR models only.  Ends: a UTR exon state can end at this column */
    MB_U5ATG = 1u << 11,    /* utr5single / utr5term: start codon behind the translation-initiation window */
    MB_UTTS = 1u << 12,     /* utr3single / utr3term: polyA signal + cleavage window ends here (or last column) */
    MB_URTSS = 1u << 13,    /* rutr5single / rutr5init: reverse TSS window ends here */
    MB_URSTOP = 1u << 14,   /* rutr3single / rutr3init: reverse stop codon follows */
    /* Begins: a begin signal starts at this column, the predecessor chain's value at column - 1 goes to a candidate list */
    MB_TSSB = 1u << 15, MB_ASSB = 1u << 16, MB_RDSSB = 1u << 17, MB_RTTSB = 1u << 18,
    MB_UTR_ENDS = MB_U5ATG | MB_UTTS | MB_URTSS | MB_URSTOP, MB_UTR_BEGINS = MB_TSSB | MB_ASSB | MB_RDSSB | MB_RTTSB
};
typedef uint32_t mask_t;

/* prefix arrays per GC class, each (L+1) long: P[i+1] - P[l] = sum over positions l..i */
/* per-class prefix arrays; the +phi families are 3-periodic content sums (term of position p uses reading frame mod3(phi + p)
 * on the forward strand, mod3(phi - p) on the reverse strand).  XET / XIN = exon-terminal and initial content
 * (ExonModel::eTermSeqProb / initialSeqProb, exonmodel.cc:1979-2034) so that the sweep never loops over positions */
enum : int { PA_PI = 0, PA_PIR = 1, PA_PX = 2 /* +phi */, PA_PXR = 5 /* +phi */, PA_XET = 8 /* +phi */, PA_XETR = 11 /* +phi */,
             PA_XIN = 14 /* +phi */, PA_XINR = 17 /* +phi */, PA_NSCAN = 20 /* the arrays above are prefix sums */,
             PA_BEG = 20 /* per position, not a sum: begin score of a forward initial / single exon whose start codon is there */, PA_PER_CLASS = 21 };

/* how one UTR exon state scores a (predecessor end, duration) candidate: UtrModel::notEndPartEmiProb (utrmodel.cc:1167-1405) as
 * begin signal + content prefix difference + length distribution, all table driven */
enum : int { BS_NONE, BS_TSSF, BS_ASSF, BS_DSSR, BS_TTSR };           /* begin signal */
enum : int { UE_ATG, UE_DSSF, UE_TTSF, UE_TSSR, UE_ASSR, UE_RSTOP };  /* end part (utrmodel.cc:1072-1110) */
enum : int { US_INIT5 = 0, US_5 = 1, US_3 = 2, US_RINIT5 = 3, US_R5 = 4, US_R3 = 5, US_NC = 6 /* NcModel::segProbs: intron content */, NUSEG = 7 };   /* SegProbs arrays (utrmodel.cc:701-712) */
struct UtrDesc {
    int8_t list;            /* candidate list */
    int8_t begsig, endkind, seg;
    int8_t shortrule;       /* middle part shorter than nothing: 0 = plain, 1 = 2^n, 2 = 4^n (utrmodel.cc:1180,1238,1250,1273,1323,1335) */
    int8_t trunc;           /* 1 = left-truncated begins (before the window) are allowed: TSS / reverse polyA signal partly outside */
    int16_t bom_off, bobe_off;      /* beginOfMiddle, beginOfBioExon relative to begin */
    int16_t lm_off, rm_off;         /* leftMost / rightMost endOfPred = base - off */
    int16_t eobe_off, boe_off;      /* endOfBioExon = base + eobe_off, beginOfEndPart = base + boe_off (getEndPositions :1572-1643) */
    int16_t maxlen_n;       /* entries of the length table */
    int8_t ldi;             /* index into DevModel::uld */
};

struct DevModel {
    int S, C, k, d, dStateLen;
    int snip_before, snip_after;    /* memo-emulated columns around a GC-class boundary (MB_SLOW) */
    int dss_start, dss_end, ass_start, ass_end, ass_up, tiw, init_len, et_len;
    int max_exon_len, min_exon_length, dss_gc_allowed;
    int tis_n, tis_k, assm_n, assm_k, n_ld_exon, n_ld_intron;
    int GCwinsize, weighing, ncent;
    StateDesc st[MAXS];
    int8_t chain_state[NCHAIN];            /* state index of each chain (-1 if absent) */
    /* role -> state index (-1 if absent) */
    int8_t r_longdss[2][3], r_lessd[2][3], r_equald[2][3], r_longass[2][3];      /* [0]=fwd [1]=rev */
    int8_t xslot[16];          /* exon states in the slot order used by Sweep::process_column */
    int8_t r_single, r_initial[3], r_internal[3], r_terminal, r_rsingle, r_rinitial, r_rinternal[3], r_rterminal[3];
    /* tables (device pointers on the GPU, host pointers in the test emulator) */
    const sc_t *init, *term, *trans;                /* trans[(c*S + a)*S + s] */
    const sc_t *xemi, *xinit, *xet;                 /* [(c*3+f) << 10 | kmer] */
    const sc_t* xpls[5];                            /* [(c*3+f) << 2(l+1) | pat] */
    const sc_t *iemi, *gemi;                        /* [c << 10 | kmer] */
    const sc_t* gfirst;                             /* igenic emission for columns <= k: [c][off(j)+pat] */
    const sc_t *tis, *assm;                         /* [(c*n + i) << 2(k+1) | pat] */
    int tis_nbins; const sc_t *tis_bb, *tis_bp;     /* TRANSINITBIN (exonmodel.cc:1321-1326, 1437-1442): [c][nbins-1] bin boundaries, [c][nbins] bin probabilities */
    const sc_t *ld_single, *ld_initial, *ld_internal, *ld_terminal, *ld_intron;
    const sc_t *ass_pat, *ass_pat_non, *dss_pat, *dss_pat_non;
    sc_t startp[64];
    uint8_t isstop[64];
    sc_t ochre, amber, opal, probN, log025, log3, ass_invalid_pat;
    double centroids[MAXC * 4 * 4][4];
    double wm[4][4];
    /* softmasking (extrinsicinfo.cc:1696-1724): every lower-case run of the input is a nonexonpart hint; ln of its bonus */
    int softmask; sc_t nep_bonus;
    /* --temperature T (types.cc:443-448): the forward summands of the sampling pass use transEmiProb^((8 - T) / 8) (LLDouble::heated,
     * lldouble.cc:209-264); 1 = cold */
    double heat;
    /* ---- UTR states (UtrModel), only when utr != 0 ---- */
    int utr;
    int tuw, tss_start, tss_end, tata_start, tata_end, d_tata_min, d_tata_max, tssup_k, dpc, boxlen, tts_spacing;
    int umax, umax3s, umax3t;
    int tssm_n, tssm_k, tsstm_n, tsstm_k, tatam_n, tatam_k, ttsm_n, ttsm_k;
    int8_t r_utr[2][2][6];          /* [rev][5'][U_*] -> state */
    UtrDesc ud[18]; int8_t uslot[18];   /* UTR exon states: descriptors and state index per slot; 16, 17 = ncinternal, rncinternal */
    const sc_t *u5i, *u5, *u3, *tup;    /* content tables [c << 2(k+1) | kmer] */
    const sc_t *tssm, *tsstm, *tatam, *ttsm, *aataaa;
    const sc_t* uld[11]; int n_uld[11]; /* 5single,5init,5internal,5term,3single,3init,3internal,3term, tail5single, tail3single, nc internal (= coding internal) */
    sc_t log_polya, log_nopolya, log2, utr_tself;
    int nc; sc_t nc_tself;             /* NcModel states present; self transition of ncintron / rncintron */
    uint8_t isstart[64];
};

/* one DP cell of a sparse (non-chain) state that is non-zero */
struct Event {                  /* the column is implied by WinView::evstart */
    int16_t state, pred;
    int32_t predbase;
    sc_t V;
};
/* candidate list entry: a predecessor cell that later state ends look back to */
struct Cand {
    int32_t col;        /* column of the predecessor cell (endOfPred) */
    int32_t state;      /* predecessor state index */
    sc_t V;
};
/* change point of a lazily evaluated self-loop chain: V[col][chain] = tilde + A[col], entered from pred at col-1 */
struct ChainCP {
    int32_t col;
    int32_t pred;       /* -1: initial probability at column 0 */
    sc_t tilde;
};

/* forward (sum) counterpart of a chain: ln F[col][chain] = ft + A[col] from column col on, until the next entry */
/* rstay / add are filled after the forward pass (Sampler::prepare_stops): a sampling walk that stands in the chain at column col stays in
 * it iff its rand() value is <= rstay, and then adds `add` to the log-probability of the path (STOP_COMPLEX: the stop is evaluated in full) */
struct FChainCP { int32_t col; uint32_t rstay; double ft; double add; };
constexpr uint32_t STOP_COMPLEX = 0xffffffffu;

/* one (predecessor, predecessor end) option of a sampling step: OptionListItem (vitmatrix.hh:748-770) */
struct SampleOpt { double lp; int32_t ord /* insertion order in the reference's loops */, pred, eop, pad; };

/* Sampling walks of one window meet the same cells again and again (99 walks follow similar paths): the option list of a cell
 * (OptionsList of state s at column j: a function of the forward matrix only) is kept in drawing form after its first use — options in the
 * stable descending order OptionsList::prepareSampling gives them, with exp(lp - max), so that a later visit only forms the cumulative sums. */
struct OcSlot { long long key /* state << 32 | column, -1 = empty */; int32_t off, n; double mx, cum; };
struct OcOpt { double ex, lpmx /* lp - max */; int32_t pred, eop; };
constexpr int OC_SLOTS = 8192, OC_PROBE = 8, OC_MAXN = 4096;

/* per-column signal score arrays written by the prep pass (ghmm_signal.h) */
enum : int { SG_DSSF = 0, SG_DSSR = 1, SG_ASSF = 2, SG_ASSR = 3, SG_XRS = 4, NSIG = 5 };

/* SnippetProbs memo restated for the columns around GC-class boundaries (Sweep::snip_get) */
/* the emulation starts DevModel::snip_before (= dStateLen + 8) columns before a boundary and ends snip_after (= 2 * dStateLen + 16) after it */
constexpr int SNIP_RING = 2048;         /* key ring (positions), power of two > dStateLen + slack (checked when the model is built) */
struct SnipEnt { int32_t len; uint32_t next; sc_t val; };        /* next = absolute entry id, 0 = none */
struct SnipHead { uint32_t first, last; };
struct SnipFrame { int32_t base, len; int32_t add, pad; sc_t part; };
/* forward pass with sampling: a lessD content value of a memo-emulated column that is NOT the plain prefix difference (the memo mixed
 * the tables of two GC classes).  A sampling step at that column finds the value again here — the reference finds it in the memo. */
struct SnipX { int32_t col; int32_t lendir /* len << 1 | strand */; sc_t val; };

/* candidate lists of a window */
enum : int { CL_LD = 0 /* +f: longdss_f */, CL_RA = 3 /* +f: rlongass_f */, CL_LA = 6 /* +phase */, CL_RD = 9 /* +phase */, NCL_BASE = 12,
             /* UTR models: (begin-signal site, predecessor chain value) lists and the exon cells UTR states follow */
             CL_T5 = 12 /* TSS, igenic */, CL_A5 = 13 /* ASS, utr5intron */, CL_A3 = 14 /* ASS, utr3intron */,
             CL_R5I = 15 /* rDSS, rutr5intron, for rutr5init */, CL_R5N = 16 /* the same sites for rutr5internal (other content table) */,
             CL_R3 = 17 /* rDSS, rutr3intron */, CL_TR = 18 /* reverse polyA, igenic */, CL_X3 = 19 /* single, terminal cells */,
             CL_XRS = 20 /* rsingle, rinitial cells for rutr5single */, CL_XRT = 21 /* the same cells for rutr5term */,
             CL_NCA = 22 /* ASS, ncintron */, CL_NCR = 23 /* rDSS, rncintron */, NCL = 24 };
/* An entry of a UTR list holds V[col][pred] + g, g = begin-signal score - SegProbs cumulative sum just before the middle part of an exon
 * that begins at col + 1: everything of a candidate's score that does not depend on where the exon ends (forward mode keeps g beside
 * ln F in WinView::clG). */

/* per-window view of the workspace (all pointers device/global) */
struct WinView {
    int L, nclassmask, ev_cap, cl_cap, cp_cap;
    const uint8_t* code;       /* 0..3, 4 = unknown */
    const uint8_t* gc;         /* class per position */
    const mask_t* mask;
    const uint16_t *kf, *kr;   /* (k+1)-mer code ending / reverse-complement code starting at each position, 0x8000 = invalid */
    const sc_t* const* parr_c; /* -> WinOuts::slab: per GC class present [PA_PER_CLASS][L+1] (first class in the window, others from the slab pool) */
    const sc_t* sig;           /* [NSIG][L] */
    const sc_t *AIG, *AGEO;    /* chain prefix arrays, [L] */
    /* UTR models: intron emission prefix of the UTR-intron chains [L], SegProbs cumulative sums [NUSEG][L+1], TSS / TTS scores */
    const sc_t *AINT, *useg, *tssF, *tssR, *ttsF, *ttsR;     /* tss*[L] by left end of the window, tts*[L+1] by first base of the polyA box */
    const int32_t *nsf, *nsr;  /* nearestStopForward / Reverse (exonmodel.cc:101-156) */
    const int32_t* pmask;      /* softmasking models: number of lower-case input bases before position i, [L+1] */
    Event* ev; int32_t* evstart;
    Cand* cl0; ChainCP* cp0;   /* list i / chain i start at cl0 + i*cl_stride, cp0 + i*cp_stride (no pointer tables: those end up in local memory) */
    int cl_stride, cp_stride;
    /* forward pass (only when sampling is requested): ln forward value of every event / list entry, chain sums */
    double* evF; double* clF0; FChainCP* fcp0; int fcp_stride, fcp_cap;
    sc_t* clG0;                /* g of the UTR list entries (forward mode), indexed like clF */
    AUGB_HD sc_t* clG(int i) const { return clG0 + (size_t)i * cl_stride; }
    AUGB_HD double* clF(int i) const { return clF0 + (size_t)i * cl_stride; }
    AUGB_HD FChainCP* fcp(int i) const { return fcp0 + (size_t)i * fcp_stride; }
    int32_t* out_nfcp;
    SnipHead* snip_head;       /* [2][SNIP_RING] */
    SnipEnt* snip_pool;        /* [2][snip_cap] */
    SnipFrame* snip_stack;     /* [SNIP_RING] */
    int snip_cap;              /* power of two */
    SnipX* snipx; int sx_cap;  /* forward + sampling only */
    AUGB_HD Cand* cl(int i) const { return cl0 + (size_t)i * cl_stride; }
    AUGB_HD ChainCP* cp(int i) const { return cp0 + (size_t)i * cp_stride; }
    /* results */
    int32_t* out_n_ev; int32_t* out_ncp; int32_t* out_status;
    const int32_t* flags;      /* written by prep: bits 0..7 GC classes present, bit 8 = no a/c/g/t at all */
};

}  // namespace augb
