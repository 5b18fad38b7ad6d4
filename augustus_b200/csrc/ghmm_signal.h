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
/* BinnedMMGroup::getIndex (merkmal.cc:155-168) + avprobs: with TRANSINITBIN the start codon x TIS motif probability is replaced by the
 * probability of the bin it falls into */
AUGB_HD sc_t tis_bin(const DevModel* m, int cls, sc_t p) {
    const int nb = m->tis_nbins;
    if (nb < 1) return p;
    const sc_t* bb = m->tis_bb + (size_t)cls * (nb - 1); int a = 0, b = nb - 1;
    AUGB_ROLLED
    while (a < b) { int mid = (a + b) / 2; if (p < bb[mid]) b = mid; else a = mid + 1; }
    return m->tis_bp[(size_t)cls * nb + a];
}
AUGB_HD sc_t rstart_endpart(const DevModel* m, const Seq& sq, int cls, int end) {
    const int L = sq.L;
    int sp = end - m->tiw - 3 + 1;
    if (sp < 0) return SC_NEG;
    int pn = sq.kmer_rc(sp, 3);
    if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
    sc_t p = m->startp[pn];
    if (sp + 3 + m->tiw - 1 + m->tis_k < L) p = tis_bin(m, cls, p + motif_rc(m, sq, cls, m->tis, m->tis_n, m->tis_k, sp + 3));
    else p = (sc_t)(L - (sp + 3)) * m->log025;
    return p;
}

/* ---- UTR signals (UtrModel) ---- */
/* UtrModel::tssupSeqProb (utrmodel.cc:1731-1750) */
AUGB_HD sc_t tssup_sum(const DevModel* m, const Seq& sq, int cls, int left, int right, int reverse) {
    sc_t s = 0; const int k = m->tssup_k; const sc_t* tab = m->tup + ((size_t)cls << (2 * (k + 1)));
    AUGB_ROLLED
    for (int p = right; p >= left; p--) {
        int pn = -1;
        if (!reverse && p - k >= 0) pn = sq.kmer_end(p, k + 1);
        else if (reverse && p >= 0 && p + k < sq.L) pn = sq.kmer_rc(p, k + 1);
        s += pn < 0 ? m->log025 : tab[pn];
    }
    return s;
}
/* UtrModel::tssProb (utrmodel.cc:1761-1833) without hints: TATA search (findTATA :271-285), TSS motif, TATA motif, upstream window.
 * `left` = first base of the whole TSS window; only every tts_spacing-th position can start a transcript (:1785) */
AUGB_HD sc_t tss_score(const DevModel* m, const Seq& sq, int cls, int fwd, int left) {
    const int right = left + m->tuw + m->tss_end - 1;
    if (left < 0 || right >= sq.L || left % m->tts_spacing != 0) return SC_NEG;
    const int maxpos = m->d_tata_max - m->d_tata_min - 1;
    sc_t mot, tata = 0, up;
    if (fwd) {
        const int p0 = right - m->tss_end - m->d_tata_max + 1; int rel = -1;
        AUGB_ROLLED
        for (int pos = 0; pos <= maxpos && rel < 0; pos++)
            if (sq.at(p0 + pos) == T_ && sq.at(p0 + pos + 1) == A_ && sq.at(p0 + pos + 2) == T_ && sq.at(p0 + pos + 3) == A_ && sq.at(p0 + pos + 5) == A_) rel = pos;
        const int pm = right - m->tss_end - m->tss_start + 1;
        if (rel >= 0) {
            const int tatapos = p0 + rel;
            mot = motif_fwd(m, sq, cls, m->tsstm, m->tsstm_n, m->tsstm_k, pm);
            tata = motif_fwd(m, sq, cls, m->tatam, m->tatam_n, m->tatam_k, tatapos - m->tata_start);
            up = tssup_sum(m, sq, cls, left, tatapos - m->tata_start - 1, 0) + tssup_sum(m, sq, cls, tatapos + m->tata_end, right - m->tss_end - m->tss_start, 0);
        } else {
            mot = motif_fwd(m, sq, cls, m->tssm, m->tssm_n, m->tssm_k, pm);
            up = tssup_sum(m, sq, cls, left, right - m->tss_end - m->tss_start, 0);
        }
    } else {
        const int p0 = left + m->tss_end + m->d_tata_max - 1; int rel = 1;
        AUGB_ROLLED
        for (int pos = 0; pos >= -maxpos && rel > 0; pos--)
            if (sq.at(p0 + pos) == A_ && sq.at(p0 + pos - 1) == T_ && sq.at(p0 + pos - 2) == A_ && sq.at(p0 + pos - 3) == T_ && sq.at(p0 + pos - 5) == T_) rel = pos;
        if (rel <= 0) {
            const int tatapos = p0 + rel;
            mot = motif_rc(m, sq, cls, m->tsstm, m->tsstm_n, m->tsstm_k, left);
            tata = motif_rc(m, sq, cls, m->tatam, m->tatam_n, m->tatam_k, tatapos - m->tata_end + 1);
            up = tssup_sum(m, sq, cls, left + m->tata_end + m->tata_start - 1, tatapos - m->tata_end, 1) + tssup_sum(m, sq, cls, tatapos + m->tata_start + 1, right, 1);
        } else {
            mot = motif_rc(m, sq, cls, m->tssm, m->tssm_n, m->tssm_k, left);
            up = tssup_sum(m, sq, cls, left + m->tss_end + m->tss_start, right, 1);
        }
    }
    if (isneg(mot) || isneg(tata) || isneg(up)) return SC_NEG;
    return mot + tata + up;
}
/* UtrModel::computeTtsProbs (utrmodel.cc:1840-1912) without hints; b = first base of the polyA signal box.
 * Kept: the reverse-strand bounds test zeroes the PLUS entry (:1886-1887), so ttsProbPlus is 0 for b < d_polyasig_cleavage. */
AUGB_HD sc_t tts_score(const DevModel* m, const Seq& sq, int cls, int fwd, int b) {
    const int L = sq.L;
    if (b < 0 || b > L) return SC_NEG;
    if (fwd) {
        if (b + m->boxlen + m->dpc - 1 >= L) return SC_NEG;
        if (b - m->dpc < 0 || b + m->boxlen - 1 >= L) return SC_NEG;
        int pn = sq.s2i(b, m->boxlen);
        sc_t prob = (pn < 0 || isneg(m->aataaa[pn])) ? SC_NEG : m->aataaa[pn] + m->log_polya;
        if (b % m->tts_spacing == 0 && isneg(prob)) prob = m->log_nopolya;
        if (isneg(prob)) return SC_NEG;
        return prob + motif_fwd(m, sq, cls, m->ttsm, m->ttsm_n, m->ttsm_k, b + m->boxlen);
    }
    const int ttspos = b - m->dpc;
    if (ttspos < 0 || b + m->boxlen - 1 >= L) return SC_NEG;
    int pn = sq.s2irc(b, m->boxlen);
    sc_t prob = (pn < 0 || isneg(m->aataaa[pn])) ? SC_NEG : m->aataaa[pn] + m->log_polya;
    if (b % m->tts_spacing == 0 && isneg(prob)) prob = m->log_nopolya;
    if (isneg(prob)) return SC_NEG;
    return prob + motif_rc(m, sq, cls, m->ttsm, m->ttsm_n, m->ttsm_k, ttspos);
}
/* one factor of a SegProbs cumulative product (SegProbs::setEmiProbs, statemodel.cc:413-432), position i in 1..L */
AUGB_HD sc_t useg_term(const DevModel* m, const Seq& sq, int cls, int g, int i) {
    const sc_t* tab = g == US_NC ? m->iemi : (g == US_INIT5 || g == US_RINIT5) ? m->u5i : (g == US_5 || g == US_R5) ? m->u5 : m->u3;
    int pn;
    if (g < US_RINIT5 || g == US_NC) pn = i < m->k ? -1 : sq.kmer_end(i, m->k + 1);       /* (NcModel::segProbs: forward k-mers, intron content) */
    else pn = i >= sq.L - m->k ? -1 : sq.kmer_rc(i, m->k + 1);
    return pn < 0 ? m->log025 : tab[((size_t)cls << (2 * (m->k + 1))) | pn];
}
/* UTR part of the activity mask of column j (ends of UTR exon states, begin-signal sites); tss* / tts* arrays already written */
AUGB_HD unsigned utr_column_mask(const DevModel* m, const Seq& s, int j, const sc_t* sg, const sc_t* tssF, const sc_t* tssR, const sc_t* ttsF, const sc_t* ttsR, int gc_last) {
    unsigned mb = 0; const int L = s.L;
    const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
    { int eobe = j + m->tiw; bool ok = true; if (eobe + 3 <= L - 1) { int c = s.kmer_end(eobe + 3, 3); ok = c >= 0 && m->isstart[c]; } if (ok) mb |= MB_U5ATG; }
    { int b0 = j - m->dpc - m->boxlen + 1; if (j == L - 1 || (b0 >= 0 && b0 + m->boxlen - 1 < L && !isneg(ttsF[b0]))) mb |= MB_UTTS; }
    { int b0 = j - m->tuw - m->tss_end + 1; if (b0 >= 0 && b0 <= L && !isneg(tssR[b0])) mb |= MB_URTSS; }
    if (!(j + 3 > L - 1 || !isRCStop(m, s, j + 1))) mb |= MB_URSTOP;
    /* ends that reuse the intron-state bits: one column earlier than longdss / rlongass can end (the window may start at base 0) */
    if (j == dssw - 1 && possDSS(m, s, j - m->dss_end - 2 + 1)) mb |= MB_LONGDSS;
    if (j == assw + m->ass_up - 1 && possRASS(s, j - m->ass_up - m->ass_start - 2 + 1)) mb |= MB_RLONGASS;
    if (j >= 1) {
        if (!isneg(tssF[j])) mb |= MB_TSSB;
        { int jj = j + assw + m->ass_up - 1;        /* last base of the acceptor pattern; the tabulated scores end at L - 1, but a UTR exon may begin with a
                                                      * site whose pattern runs up to ass_end - 1 bases past the window (the exon begins in its last bases; aSSProb then takes the
                                                      * invalid-pattern probability, intronmodel.cc:1116-1188) */
          if (jj < L ? !isneg(sg[(size_t)SG_ASSF * L + jj]) : (jj <= L + m->ass_end - 2 && !isneg(aSSProb(m, s, gc_last, j, 1)))) mb |= MB_ASSB; }
        { int jj = j + dssw - 1; if (jj < L && !isneg(sg[(size_t)SG_DSSR * L + jj])) mb |= MB_RDSSB; }
        if (j + m->dpc <= L && !isneg(ttsR[j + m->dpc])) mb |= MB_RTTSB;
    }
    return mb;
}

/* softmasking: ln of the nonexonpart bonuses over positions lo..hi — every lower-case run of the input is one hint (disjoint), its
 * bonus multiplies each covered intron position (intronmodel.cc:1011-1036, utrmodel.cc:1143-1158,1521-1545); pmask = prefix counts */
AUGB_HD sc_t nep_range(const DevModel* m, const int32_t* pmask, int L, int lo, int hi) {
    if (!m->softmask) return 0;
    if (lo < 0) lo = 0;
    if (hi > L - 1) hi = L - 1;
    return lo > hi ? (sc_t)0 : m->nep_bonus * (sc_t)(pmask[hi + 1] - pmask[lo]);
}
/* value of signal array `which` at column j (defined as soon as the signal window starts at base >= 0; the intron states need one more
 * column in front for their predecessor, Sweep::fixed_eval checks that).  With softmasking the bonus of the intron bases inside the signal window is part of
 * the value: the window of a (r)longdss / (r)longass state ending at j and the window a UTR exon state begins or ends with cover the
 * same intron bases (intronBegin / intronEnd of intronmodel.cc:873-922 = begin .. beginOfBioExon-1 / endOfBioExon+1 .. end of utrmodel.cc) */
AUGB_HD sc_t signal_term(const DevModel* m, const Seq& sq, int cls, int which, int j, const int32_t* pmask = nullptr) {
    const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2, L = sq.L;
    sc_t v; int lo = 0, hi = -1;
    switch (which) {
    case SG_DSSF: v = j - dssw + 1 >= 0 ? dSSProb(m, sq, j - dssw + 1, 1) : SC_NEG; lo = j - 2 - m->dss_end + 1; hi = j; break;
    case SG_DSSR: v = j - dssw + 1 >= 0 ? dSSProb(m, sq, j - dssw + 1, 0) : SC_NEG; lo = j - dssw + 1; hi = j - m->dss_start; break;
    case SG_ASSF: v = j - assw - m->ass_up + 1 >= 0 ? aSSProb(m, sq, cls, j - assw - m->ass_up + 1, 1) : SC_NEG; lo = j - assw - m->ass_up + 1; hi = j - m->ass_end; break;
    case SG_ASSR: v = j - assw - m->ass_up + 1 >= 0 ? aSSProb(m, sq, cls, j - assw - m->ass_up + 1, 0) : SC_NEG; lo = j - assw - m->ass_up + 1 + m->ass_end; hi = j; break;
    default: return rstart_endpart(m, sq, cls, j);
    }
    return isneg(v) ? v : v + nep_range(m, pmask, L, lo, hi);
}

}  // namespace augb
