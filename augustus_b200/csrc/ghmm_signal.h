/*
 * ghmm_signal.h — splice-site and translation-start signal scores of a column (host + device, one thread).
 *
 * These depend on the sequence and the column's GC class only, so the prep pass tabulates them for every column
 * (WinView::sig) and the sweep just reads them:
 *   SG_DSSF / SG_DSSR : emission of longdss / rlongdss ending at column j   = IntronModel::dSSProb (intronmodel.cc:1195-1248)
 *   SG_ASSF / SG_ASSR : emission of longass / rlongass ending at column j   = IntronModel::aSSProb (intronmodel.cc:1116-1188)
 *   SG_XRS            : end part of rsingle / rinitial ending at column j   = ExonModel::endPartEmiProb, reverse start codon
 *                       x TIS motif (exonmodel.cc:1313-1332)
 */
#pragma once
#include "ghmm_defs.h"
#include "ghmm_seq.h"

namespace augb {

/* Motif::seqProb (motif.cc:308-331) */
AUGB_HD sc_t motif_fwd(const DevModel* m, const Seq& sq, int cls, const sc_t* tab, int n, int k, int p) {
    sc_t s = 0; const size_t wd = (size_t)1 << (2 * (k + 1));
    AUGB_ROLLED
    for (int i = 0; i < n; i++) { int pn = sq.kmer_end(p + i, k + 1); s += pn < 0 ? m->log025 : tab[((size_t)cls * n + i) * wd + pn]; }
    return s;
}
AUGB_HD sc_t motif_rc(const DevModel* m, const Seq& sq, int cls, const sc_t* tab, int n, int k, int p) {
    sc_t s = 0; const size_t wd = (size_t)1 << (2 * (k + 1));
    AUGB_ROLLED
    for (int i = 0; i < n; i++) { int pn = sq.kmer_rc(p + i, k + 1); s += pn < 0 ? m->log025 : tab[((size_t)cls * n + (n - 1 - i)) * wd + pn]; }
    return s;
}
/* IntronModel::dSSProb (intronmodel.cc:1195-1248) */
AUGB_HD sc_t dSSProb(const DevModel* m, const Seq& sq, int base, int fwd) {
    int nonGT, idx;
    if (fwd) {
        int dsspos = base + m->dss_start;
        if (!possDSS(m, sq, dsspos)) return SC_NEG;
        nonGT = !sq.is2(dsspos, G_, T_);
        int a = sq.kmer_end(base + m->dss_start - 1, m->dss_start), b = sq.kmer_end(dsspos + 2 + m->dss_end - 1, m->dss_end);
        if (a < 0 || b < 0) return SC_NEG;
        idx = (a << (2 * m->dss_end)) | b;
    } else {
        int dsspos = base + m->dss_end;
        if (!possRDSS(m, sq, dsspos + 1)) return SC_NEG;
        nonGT = !sq.is2(dsspos, A_, C_);
        int a = 0, b = 0;
        AUGB_ROLLED
        for (int i = m->dss_start - 1; i >= 0; i--) { int c = sq.at(dsspos + 2 + i); if (c > 3) return SC_NEG; a = (a << 2) | (3 - c); }
        AUGB_ROLLED
        for (int i = m->dss_end - 1; i >= 0; i--) { int c = sq.at(base + i); if (c > 3) return SC_NEG; b = (b << 2) | (3 - c); }
        idx = (a << (2 * m->dss_end)) | b;
    }
    return nonGT ? m->dss_pat_non[idx] : m->dss_pat[idx];
}
/* IntronModel::aSSProb (intronmodel.cc:1116-1188) */
AUGB_HD sc_t aSSProb(const DevModel* m, const Seq& sq, int cls, int base, int fwd) {
    int nonAG, a, b; sc_t motif; const int L = sq.L;
    if (fwd) {
        int asspos = base + m->ass_up + m->ass_start;
        if (!possASS(sq, asspos + 1)) return SC_NEG;
        nonAG = !sq.is2(asspos, A_, G_);
        a = sq.kmer_end(base + m->ass_up + m->ass_start - 1, m->ass_start); b = sq.kmer_end(asspos + 2 + m->ass_end - 1, m->ass_end);
        motif = base >= m->assm_k ? motif_fwd(m, sq, cls, m->assm, m->assm_n, m->assm_k, base) : SC_NEG;
    } else {
        int asspos = base + m->ass_end;
        if (!possRASS(sq, asspos)) return SC_NEG;
        nonAG = !sq.is2(asspos, C_, T_);
        a = 0; b = 0;
        AUGB_ROLLED
        for (int i = m->ass_start - 1; i >= 0; i--) { int c = sq.at(asspos + 2 + i); if (c > 3) { a = -1; break; } a = (a << 2) | (3 - c); }
        AUGB_ROLLED
        for (int i = m->ass_end - 1; i >= 0; i--) { int c = sq.at(base + i); if (c > 3) { b = -1; break; } b = (b << 2) | (3 - c); }
        int motifstart = base + m->ass_start + m->ass_end + 2, motifend = motifstart + m->ass_up;
        motif = motifend + m->assm_k < L ? motif_rc(m, sq, cls, m->assm, m->assm_n, m->assm_k, motifstart) : (sc_t)m->ass_up * m->log025;
    }
    sc_t pat;
    if (a < 0 || b < 0) pat = m->ass_invalid_pat;
    else { int idx = (a << (2 * m->ass_end)) | b; pat = nonAG ? m->ass_pat_non[idx] : m->ass_pat[idx]; }
    if (isneg(motif) || isneg(pat)) return SC_NEG;
    return motif + pat;
}
/* ExonModel::endPartEmiProb for rsingle / rinitial (exonmodel.cc:1313-1332) */
AUGB_HD sc_t rstart_endpart(const DevModel* m, const Seq& sq, int cls, int end) {
    const int L = sq.L;
    int sp = end - m->tiw - 3 + 1;
    if (sp < 0) return SC_NEG;
    int pn = sq.kmer_rc(sp, 3);
    if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
    sc_t p = m->startp[pn];
    if (sp + 3 + m->tiw - 1 + m->tis_k < L) p += motif_rc(m, sq, cls, m->tis, m->tis_n, m->tis_k, sp + 3);
    else p = (sc_t)(L - (sp + 3)) * m->log025;
    return p;
}
/* value of signal array `which` at column j */
AUGB_HD sc_t signal_term(const DevModel* m, const Seq& sq, int cls, int which, int j) {
    const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
    switch (which) {
    case SG_DSSF: return j - dssw >= 0 ? dSSProb(m, sq, j - dssw + 1, 1) : SC_NEG;
    case SG_DSSR: return j - dssw >= 0 ? dSSProb(m, sq, j - dssw + 1, 0) : SC_NEG;
    case SG_ASSF: return j - assw - m->ass_up >= 0 ? aSSProb(m, sq, cls, j - assw - m->ass_up + 1, 1) : SC_NEG;
    case SG_ASSR: return j - assw - m->ass_up >= 0 ? aSSProb(m, sq, cls, j - assw - m->ass_up + 1, 0) : SC_NEG;
    default: return rstart_endpart(m, sq, cls, j);
    }
}

}  // namespace augb
