/*
 * ghmm_prep.h — per-window preparation: everything that depends on the sequence but not on the DP.
 *
 * Replaces the parts of the reference that are re-evaluated inside its duration loops:
 *   - Seq2Int k-mer lookups + content-model products (ExonModel::seqProb exonmodel.cc:1925-1973,
 *     IntronModel::seqProb intronmodel.cc:1046-1108, SnippetProbs statemodel.cc:283-310,
 *     IGenicModel::emiProbUnderModel igenicmodel.cc:299-357)  -> fixed-point prefix sums;
 *   - OpenReadingFrame's nearest-stop tables (exonmodel.cc:101-156);
 *   - ContentStairs::computeStairs (motif.cc:543-614) when the host does not pass the classes;
 *   - the sequence-only early-return tests of the state models -> activity mask.
 * This header holds the layout, the per-position formulas (host + device) and a sequential host
 * builder used by the test emulator; the CUDA kernels in ghmm_kernels.cu evaluate the same formulas in
 * parallel and scan with integer adds, so both produce identical arrays.
 */
#pragma once
#include <math.h>
#include <stddef.h>
#include "ghmm_defs.h"
#include "ghmm_seq.h"
#include "ghmm_signal.h"

namespace augb {

/* result / hand-over block of a window (WinLayout::outs) */
struct WinOuts {
    int32_t n_ev, status, path_n, path_status; int32_t ncp[NCHAIN]; int32_t pad /* window flags */; sc_t score;
    int32_t nfcp[NCHAIN]; int32_t samp_status;
    int32_t rand_used, pad2;    /* rand() draws the sampling walks of this window consumed */
    const sc_t* slab[MAXC];     /* prefix-array slab of each GC class (set by prep) */
};

/* byte layout of one window's workspace; every section is 16-byte aligned */
struct WinLayout {
    size_t code, gc, mask, kf, kr, parr, sig, aig, ageo, nsf, nsr;        /* static (prep) */
    size_t aint, useg, tssF, tssR, ttsF, ttsR; int utr, ncl, nchain;      /* UTR models only */
    size_t pmask; int softmask;                                           /* softmasking models only */
    size_t ev, evstart, cl[NCL], cp[NCHAIN], outs;           /* dynamic (sweep) */
    size_t snip_head, snip_pool, snip_stack; int snip_cap;
    size_t snipx; int sx_cap;
    size_t evF, clF, clG, fcp; int fcp_cap;                       /* forward pass (0 capacity when not requested) */
    size_t opt, sorted, optex, oc_slots, oc_pool, s_begin, s_end, s_type, s_trunc, s_count, s_logp; int opt_cap, samp_cap, nsamp, oc_cap;   /* sampling */
    size_t path_begin, path_end, path_type, path_trunc;      /* backtrace output */
    size_t total, slab;
    int ev_cap, cl_cap, cp_cap, path_cap, nslab_local;
};
AUGB_HD size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
/* Capacities of the dynamic structures.  `generous` = provable upper bounds (every column can hold at most one
 * cell per state; a candidate list gets at most one entry per column); the default sizes are ~3x what human-like
 * DNA needs (measured: 1.9 events, 0.03 list entries per base) and a window that overflows them is reported with
 * status AUGB200_ERR_CAPACITY and decoded again with the generous layout (augb200.cu: decode_batch). */
inline WinLayout make_layout(int L, int C, bool generous = false, bool forward = false, int nsamp = 0, bool utr = false, bool softmask = false) {
    WinLayout w; size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = al16(o + bytes); return r; };
    w.code = take(L); w.gc = take(L); w.mask = take((size_t)L * sizeof(mask_t)); w.kf = take((size_t)L * 2); w.kr = take((size_t)L * 2);
    /* prefix arrays: one slab (PA_PER_CLASS arrays) in the window for the first GC class present; further classes take
     * slabs from a pool shared by the batch (generous layout: all C slabs in the window) */
    w.slab = (size_t)PA_PER_CLASS * (L + 1) * sizeof(sc_t);
    w.nslab_local = generous ? C : 1;
    w.parr = take(w.slab * w.nslab_local);
    w.sig = take((size_t)NSIG * L * sizeof(sc_t));
    w.aig = take((size_t)L * sizeof(sc_t)); w.ageo = take((size_t)L * sizeof(sc_t));
    w.nsf = take((size_t)(L + 3) * 4); w.nsr = take((size_t)(L + 3) * 4);
    w.softmask = softmask ? 1 : 0; w.pmask = take(softmask ? (size_t)(L + 1) * 4 : 0);
    w.utr = utr ? 1 : 0; w.ncl = utr ? NCL : NCL_BASE; w.nchain = utr ? NCHAIN : CH_UTR;
    w.aint = take(utr ? (size_t)L * sizeof(sc_t) : 0); w.useg = take(utr ? (size_t)NUSEG * (L + 1) * sizeof(sc_t) : 0);
    w.tssF = take(utr ? (size_t)(L + 1) * sizeof(sc_t) : 0); w.tssR = take(utr ? (size_t)(L + 1) * sizeof(sc_t) : 0);     /* [L+1]: a model without a TSS window (tuw + tss_end = 0) asks for position L */
    w.ttsF = take(utr ? (size_t)(L + 1) * sizeof(sc_t) : 0); w.ttsR = take(utr ? (size_t)(L + 1) * sizeof(sc_t) : 0);
    if (generous) { w.ev_cap = 48 * L + 256; w.cl_cap = L + 64; w.cp_cap = L + 64; w.path_cap = L + 64; }
    else { w.ev_cap = (utr ? 5 : 3) * L + 256; w.cl_cap = L / 6 + 64; w.cp_cap = L / 16 + 64; w.path_cap = L / 8 + 64; }
    if ((size_t)w.ev_cap * sizeof(Event) < (size_t)4 * (L + 1) * 4) w.ev_cap = (int)(((size_t)4 * (L + 1) * 4) / sizeof(Event) + 1);   /* prep scratch */
    w.ev = take((size_t)w.ev_cap * sizeof(Event)); w.evstart = take((size_t)(L + 2) * 4);
    for (int i = 0; i < NCL; i++) w.cl[i] = take(i < w.ncl ? (size_t)w.cl_cap * sizeof(Cand) : 0);
    for (int i = 0; i < NCHAIN; i++) w.cp[i] = take(i < w.nchain ? (size_t)w.cp_cap * sizeof(ChainCP) : 0);
    w.fcp_cap = forward ? (generous ? L + 64 : L / 2 + 64) : 0;
    w.evF = take(forward ? (size_t)w.ev_cap * 8 : 0); w.clF = take(forward ? (size_t)w.ncl * w.cl_cap * 8 : 0); w.clG = take(forward && utr ? (size_t)w.ncl * w.cl_cap * 8 : 0);
    w.fcp = take((size_t)w.nchain * w.fcp_cap * sizeof(FChainCP));
    w.nsamp = forward ? nsamp : 0;
    w.opt_cap = w.nsamp ? w.cl_cap + 1024 : 0; w.samp_cap = w.nsamp ? (generous ? w.nsamp * (L / 8 + 64) : w.nsamp * (128 + L / 128)) : 0;   /* gene-dense fly DNA: ~3.5 path states per kb and sample */
    w.opt = take((size_t)w.opt_cap * sizeof(SampleOpt)); w.sorted = take((size_t)w.opt_cap * 4); w.optex = take((size_t)w.opt_cap * 8);
    w.oc_cap = w.nsamp ? (generous ? 8 * L + 65536 : L / 2 + 32768) : 0;         /* cached option lists of the walks (Sampler::step_pick) */
    w.oc_slots = take(w.nsamp ? (size_t)OC_SLOTS * sizeof(OcSlot) : 0); w.oc_pool = take((size_t)w.oc_cap * sizeof(OcOpt));
    w.s_begin = take((size_t)w.samp_cap * 4); w.s_end = take((size_t)w.samp_cap * 4); w.s_type = take(w.samp_cap); w.s_trunc = take(w.samp_cap);
    w.s_count = take((size_t)w.nsamp * 4 + 16); w.s_logp = take((size_t)w.nsamp * 8);
    w.outs = take(sizeof(WinOuts));
    w.snip_cap = generous ? 262144 : 16384;
    w.snip_head = take((size_t)2 * SNIP_RING * sizeof(SnipHead)); w.snip_pool = take((size_t)2 * w.snip_cap * sizeof(SnipEnt));
    w.snip_stack = take((size_t)SNIP_RING * sizeof(SnipFrame));
    w.sx_cap = forward ? (generous ? 8 * L + 64 : L / 4 + 256) : 0;
    w.snipx = take((size_t)w.sx_cap * sizeof(SnipX));
    w.path_begin = take((size_t)w.path_cap * 4); w.path_end = take((size_t)w.path_cap * 4);
    w.path_type = take(w.path_cap); w.path_trunc = take(w.path_cap);
    w.total = al16(o);
    return w;
}

/* per-window descriptor in device memory */
struct WinDev {
    char* base;                /* workspace of this window */
    const char* dna;           /* device copy of the ASCII window */
    const uint8_t* gc_in;      /* device copy of host-provided classes or nullptr */
    int L;
    WinLayout lay;
};

AUGB_HD WinView make_view(char* base, const WinLayout& lay, int L, int classmask) {
    WinView v; v.L = L; v.nclassmask = classmask; v.ev_cap = lay.ev_cap; v.cl_cap = lay.cl_cap; v.cp_cap = lay.cp_cap;
    v.code = (const uint8_t*)(base + lay.code); v.gc = (const uint8_t*)(base + lay.gc); v.mask = (const mask_t*)(base + lay.mask);
    v.kf = (const uint16_t*)(base + lay.kf); v.kr = (const uint16_t*)(base + lay.kr);
    v.sig = (const sc_t*)(base + lay.sig); v.AIG = (const sc_t*)(base + lay.aig); v.AGEO = (const sc_t*)(base + lay.ageo);
    v.nsf = (const int32_t*)(base + lay.nsf); v.nsr = (const int32_t*)(base + lay.nsr);
    v.pmask = (const int32_t*)(base + lay.pmask);
    v.AINT = (const sc_t*)(base + lay.aint); v.useg = (const sc_t*)(base + lay.useg);
    v.tssF = (const sc_t*)(base + lay.tssF); v.tssR = (const sc_t*)(base + lay.tssR); v.ttsF = (const sc_t*)(base + lay.ttsF); v.ttsR = (const sc_t*)(base + lay.ttsR);
    v.ev = (Event*)(base + lay.ev); v.evstart = (int32_t*)(base + lay.evstart);
    v.cl0 = (Cand*)(base + lay.cl[0]); v.cp0 = (ChainCP*)(base + lay.cp[0]);
    v.evF = (double*)(base + lay.evF); v.clF0 = (double*)(base + lay.clF); v.clG0 = (sc_t*)(base + lay.clG); v.fcp0 = (FChainCP*)(base + lay.fcp); v.fcp_cap = lay.fcp_cap; v.fcp_stride = lay.fcp_cap;
    v.snip_head = (SnipHead*)(base + lay.snip_head); v.snip_pool = (SnipEnt*)(base + lay.snip_pool); v.snip_stack = (SnipFrame*)(base + lay.snip_stack); v.snip_cap = lay.snip_cap; v.snipx = (SnipX*)(base + lay.snipx); v.sx_cap = lay.sx_cap;
    v.cl_stride = (int)((lay.cl[1] - lay.cl[0]) / sizeof(Cand)); v.cp_stride = (int)((lay.cp[1] - lay.cp[0]) / sizeof(ChainCP));
    WinOuts* o = (WinOuts*)(base + lay.outs);
    v.out_nfcp = o->nfcp;
    v.out_n_ev = &o->n_ev; v.out_status = &o->status; v.out_ncp = o->ncp; v.flags = &o->pad;
    v.parr_c = o->slab;                                          /* filled by prep */
    return v;
}

/* (k+1)-mer codes of a position: forward code of the k+1 bases ENDING at p, reverse-complement code of the
 * k+1 bases STARTING at p (Seq2Int::operator() / ::rc, geneticcode.hh:166-179); 0x8000 = not all acgt / off the end */
AUGB_HD uint16_t kmer_code_f(const Seq& s, int p, int k1) { int v = s.s2i(p - k1 + 1, k1); return v < 0 ? (uint16_t)0x8000 : (uint16_t)v; }
AUGB_HD uint16_t kmer_code_r(const Seq& s, int p, int k1) { int v = s.s2irc(p, k1); return v < 0 ? (uint16_t)0x8000 : (uint16_t)v; }

AUGB_HD uint8_t base_code(char ch) {
    switch (ch) {
    case 'a': case 'A': return 0; case 'c': case 'C': return 1; case 'g': case 'G': return 2; case 't': case 'T': return 3;
    default: return 4;
    }
}

/* nearest GC-content class of a base-count vector: ContentDecomposition::getNearestBaseCountIndex
 * (motif.cc:493-505) with BaseCount::doubleWeight (motif.cc:105-121) */
AUGB_HD int nearest_class(const DevModel* m, const int cnt[4]) {
    double sum = (double)cnt[0] + cnt[1] + cnt[2] + cnt[3];
    double r[4] = {0.25, 0.25, 0.25, 0.25};
    if (sum > 0) for (int i = 0; i < 4; i++) r[i] = cnt[i] / sum;
    double best = -1; int ret = -1;
    for (int i = 0; i < m->ncent; i++) {
        double wgt;
        if (m->weighing == 3) {
            double z[4], tmp[4] = {0, 0, 0, 0}, t = 0;
            for (int j = 0; j < 4; j++) z[j] = r[j] - m->centroids[i][j];
            for (int j = 0; j < 4; j++) for (int a = 0; a < 4; a++) tmp[j] += z[a] * m->wm[a][j];
            for (int a = 0; a < 4; a++) t += tmp[a] * z[a];
            wgt = 1 + 9 * exp(-t);
        } else if (m->weighing == 2) {
            double g1 = r[2] + r[1], g2 = m->centroids[i][2] + m->centroids[i][1];
            int c1 = g1 < .43 ? 0 : g1 < .51 ? 1 : g1 < .57 ? 2 : 3, c2 = g2 < .43 ? 0 : g2 < .51 ? 1 : g2 < .57 ? 2 : 3;
            wgt = c1 == c2;
        } else wgt = 1;
        if (wgt > best) { best = wgt; ret = i; }
    }
    return ret;
}

/* ContentStairs::computeStairs (motif.cc:543-614), sequential host version */
inline void gc_stairs_seq(const DevModel* m, const uint8_t* c, int n, uint8_t* idx) {
    int win = m->GCwinsize; if (win > n || win < 1) win = n;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < win; i++) if (c[i] < 4) cnt[c[i]]++;
    int xx = nearest_class(m, cnt);
    for (int i = 0; i <= win / 2 && i < n; i++) idx[i] = (uint8_t)xx;
    for (int i = win / 2 + 1; i <= n - (win + 1) / 2; i++) {
        int a = c[i + (win + 1) / 2 - 1], r = c[i - win / 2 - 1];
        if (a < 4) cnt[a]++;
        if (r < 4) cnt[r]--;
        idx[i] = (uint8_t)(xx = nearest_class(m, cnt));
    }
    for (int i = n - (win + 1) / 2 + 1; i < n; i++) if (i >= 0) idx[i] = (uint8_t)xx;
    int x2 = -2, lastStep = 0; const int tot = 1000;
    for (int i = 0; i < n; i++) if (idx[i] != x2) {
        if (i - lastStep < tot && lastStep > 0 && idx[lastStep - 1] == idx[i]) for (int j = lastStep; j < i; j++) idx[j] = idx[i];
        lastStep = i; x2 = idx[i];
    }
}

/* per-position addends of the scanned arrays */
/* softmasking bonus of position j (ln; 0 when the base is not masked) */
AUGB_HD sc_t nep_term(const DevModel* m, const int32_t* pmask, int j) { return (m->softmask && pmask[j + 1] != pmask[j]) ? m->nep_bonus : (sc_t)0; }
AUGB_HD sc_t aig_term(const DevModel* m, const Seq& s, const uint8_t* gc, int j, const int32_t* pmask = nullptr) {      /* j >= 1 */
    int c = gc[j], ig = m->chain_state[0];
    return m->trans[((size_t)c * m->S + ig) * m->S + ig] + igenic_emi(m, s, c, j) + nep_term(m, pmask, j);
}
AUGB_HD sc_t ageo_term(const DevModel* m, const Seq& s, const uint8_t* gc, int j, const int32_t* pmask = nullptr) {     /* j >= 1 */
    int c = gc[j]; int g = -1;
    for (int ch = 1; ch < CH_UTR; ch++) if (m->chain_state[ch] >= 0) { g = m->chain_state[ch]; break; }
    if (g < 0) return 0;
    return m->trans[((size_t)c * m->S + g) * m->S + g] + intron_emi1(m, s, c, j) + nep_term(m, pmask, j);
}
/* intron content arrays carry the softmasking bonus of their position (lessD / equalD emissions are differences of them) */
AUGB_HD sc_t parr_term(const DevModel* m, const Seq& s, int c, int which, int p, const int32_t* pmask = nullptr) {
    switch (which) {
    case PA_PI: return intron_emi1(m, s, c, p) + nep_term(m, pmask, p);
    case PA_PIR: return intron_emi1r(m, s, c, p) + nep_term(m, pmask, p);
    case PA_PX: case PA_PX + 1: case PA_PX + 2: return p >= m->k ? exon_emi1(m, s, m->xemi, c, 1, mod3(which - PA_PX + p), p) : 0;
    case PA_PXR: case PA_PXR + 1: case PA_PXR + 2: return exon_emi1(m, s, m->xemi, c, 0, mod3(which - PA_PXR - p), p);
    case PA_XET: case PA_XET + 1: case PA_XET + 2: return exon_emi1(m, s, m->xet, c, 1, mod3(which - PA_XET + p), p);
    case PA_XETR: case PA_XETR + 1: case PA_XETR + 2: return exon_emi1(m, s, m->xet, c, 0, mod3(which - PA_XETR - p), p);
    case PA_XIN: case PA_XIN + 1: case PA_XIN + 2: return exon_emi1(m, s, m->xinit, c, 1, mod3(which - PA_XIN + p), p);
    default: return exon_emi1(m, s, m->xinit, c, 0, mod3(which - PA_XINR - p), p);
    }
}

/* PA_BEG[bobe]: start codon x TIS motif, P_ls start pattern and initial content of a forward initial / single exon whose start
 * codon begins at bobe (exonmodel.cc:1427-1462, 1596-1603, 1611-1633), with the tables of class c.  The reading frame of a
 * position p of such an exon is mod3(p - bobe), whatever state and exon end ask for it; SC_NEG = no start codon here. */
AUGB_HD sc_t begin_term(const DevModel* m, const Seq& s, int c, int bobe) {
    const int k = m->k, bos = bobe + 3, L = s.L;
    if (!(bobe >= 0 && bobe < L - 2)) return SC_NEG;
    int pn = s.kmer_end(bobe + 2, 3);
    if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
    sc_t v = m->startp[pn];
    int tis = bobe - m->tiw;
    if (tis > m->tis_k) v = tis_bin(m, c, v + motif_fwd(m, s, c, m->tis, m->tis_n, m->tis_k, tis));
    else v += (sc_t)(bos - 3) * m->log025;
    const int endOfStart = bos + k - 1;
    if (k >= 1) {
        int p4 = s.kmer_end(endOfStart, k);
        v += p4 < 0 ? (sc_t)k * m->probN : m->xpls[k - 1][(((size_t)c * 3 + mod3(endOfStart - bobe)) << (2 * k)) | p4];
    }
    const int endOfInitial = endOfStart + m->init_len;
    for (int p = endOfInitial; p > endOfStart; p--) v += exon_emi1(m, s, m->xinit, c, 1, mod3(p - bobe), p);
    return v;
}

/* sequential host builder (test emulator) */
inline void prep_window_seq(const DevModel* m, const char* dna, int L, const int32_t* gc_in, char* base, const WinLayout& lay, int* classmask,
                            char* pool = nullptr, size_t pool_size = 0, size_t* pool_used = nullptr) {
    uint8_t* code = (uint8_t*)(base + lay.code); uint8_t* gc = (uint8_t*)(base + lay.gc); mask_t* mask = (mask_t*)(base + lay.mask);
    for (int i = 0; i < L; i++) code[i] = base_code(dna[i]);
    int32_t* pmask = (int32_t*)(base + lay.pmask);
    if (lay.softmask) { pmask[0] = 0; for (int i = 0; i < L; i++) pmask[i + 1] = pmask[i] + (dna[i] >= 'a' && dna[i] <= 'z'); }
    if (gc_in) for (int i = 0; i < L; i++) gc[i] = (uint8_t)gc_in[i]; else gc_stairs_seq(m, code, L, gc);
    int cm = 0; bool anynuc = false;
    for (int i = 0; i < L; i++) { cm |= 1 << gc[i]; anynuc |= code[i] < 4; }
    if (!anynuc) cm |= WF_ALLN;
    *classmask = cm;
    ((WinOuts*)(base + lay.outs))->pad = cm;
    Seq s; s.c = code; s.L = L;
    {
        uint16_t* kf = (uint16_t*)(base + lay.kf); uint16_t* kr = (uint16_t*)(base + lay.kr);
        for (int p = 0; p < L; p++) { kf[p] = kmer_code_f(s, p, m->k + 1); kr[p] = kmer_code_r(s, p, m->k + 1); }
        s.kf = kf; s.kr = kr; s.k1 = m->k + 1;
    }
    for (int j = 0; j < L; j++) mask[j] = anynuc ? (mask_t)column_mask(m, s, j) : 0;
    if (anynuc) for (int b = 1; b < L; b++) if (gc[b] != gc[b - 1])        /* GC-class boundary at b */
        for (int j = (b - m->snip_before < 1 ? 1 : b - m->snip_before); j < L && j < b + m->snip_after; j++) mask[j] |= MB_SLOW;
    {
        sc_t* sg = (sc_t*)(base + lay.sig);
        for (int which = 0; which < NSIG; which++)
            for (int j = 0; j < L; j++) sg[(size_t)which * L + j] = anynuc ? signal_term(m, s, gc[j], which, j, pmask) : SC_NEG;
    }
    if (lay.utr) {
        /* UTR models: TSS / TTS scores, SegProbs cumulative sums, intron emission prefix of the UTR-intron chains, mask bits */
        sc_t* tssF = (sc_t*)(base + lay.tssF); sc_t* tssR = (sc_t*)(base + lay.tssR); sc_t* ttsF = (sc_t*)(base + lay.ttsF); sc_t* ttsR = (sc_t*)(base + lay.ttsR);
        for (int i = 0; i <= L; i++) {
            int r = i + m->tuw + m->tss_end - 1; int c = gc[r < 0 ? 0 : r < L ? r : L - 1];
            tssF[i] = anynuc ? tss_score(m, s, c, 1, i) : SC_NEG; tssR[i] = anynuc ? tss_score(m, s, c, 0, i) : SC_NEG;
        }
        for (int b = 0; b <= L; b++) { int c = gc[b < L ? b : L - 1]; ttsF[b] = anynuc ? tts_score(m, s, c, 1, b) : SC_NEG; ttsR[b] = anynuc ? tts_score(m, s, c, 0, b) : SC_NEG; }
        sc_t* useg = (sc_t*)(base + lay.useg);
        for (int g = 0; g < NUSEG; g++) {
            sc_t* cum = useg + (size_t)g * (L + 1); cum[0] = m->log025;
            for (int i = 1; i <= L; i++) cum[i] = cum[i - 1] + useg_term(m, s, gc[i < L ? i : L - 1], g, i);
        }
        sc_t* aint = (sc_t*)(base + lay.aint); aint[0] = 0;
        for (int j = 1; j < L; j++) aint[j] = aint[j - 1] + intron_emi1(m, s, gc[j], j) + nep_term(m, pmask, j);
        const sc_t* sg = (const sc_t*)(base + lay.sig);
        if (anynuc) for (int j = 0; j < L; j++) mask[j] |= (mask_t)utr_column_mask(m, s, j, sg, tssF, tssR, ttsF, ttsR, gc[L - 1]);
    }
    WinOuts* wo = (WinOuts*)(base + lay.outs);
    int nloc = 0;
    for (int c = 0; c < m->C; c++) {
        wo->slab[c] = nullptr;
        if (!(cm >> c & 1)) continue;
        sc_t* slab;
        if (nloc < lay.nslab_local) slab = (sc_t*)(base + lay.parr + (size_t)nloc++ * lay.slab);
        else if (pool && *pool_used + lay.slab <= pool_size) { slab = (sc_t*)(pool + *pool_used); *pool_used += lay.slab; }
        else { wo->pad |= WF_NOSLAB; *classmask |= WF_NOSLAB; continue; }
        wo->slab[c] = slab;
        for (int which = 0; which < PA_NSCAN; which++) {
            sc_t* P = slab + (size_t)which * (size_t)(L + 1);
            P[0] = 0;
            for (int p = 0; p < L; p++) P[p + 1] = P[p] + parr_term(m, s, c, which, p, pmask);
        }
        { sc_t* B = slab + (size_t)PA_BEG * (size_t)(L + 1); for (int p = 0; p <= L; p++) B[p] = begin_term(m, s, c, p); }
    }
    sc_t* aig = (sc_t*)(base + lay.aig); sc_t* ageo = (sc_t*)(base + lay.ageo);
    aig[0] = ageo[0] = 0;
    for (int j = 1; j < L; j++) { aig[j] = aig[j - 1] + (anynuc ? aig_term(m, s, gc, j, pmask) : m->log025); ageo[j] = ageo[j - 1] + ageo_term(m, s, gc, j, pmask); }
    int32_t* nsf = (int32_t*)(base + lay.nsf); int32_t* nsr = (int32_t*)(base + lay.nsr);
    for (int i = 0; i < L + 3; i++) nsf[i] = nsr[i] = 0;
    for (int r = 0; r < 3; r++) {
        int sp = -1; for (int i = r; i <= L - 3; i += 3) { if (isStop(m, s, i)) sp = i; nsf[i] = sp; }
        sp = -1; for (int i = r; i <= L - 3; i += 3) { if (isRCStop(m, s, i)) sp = i; nsr[i] = sp; }
    }
    if (L > 5) { nsf[L - 2] = nsf[L - 5]; nsf[L - 1] = nsf[L - 4]; nsr[L - 2] = nsr[L - 5]; nsr[L - 1] = nsr[L - 4]; }
}

}  // namespace augb
