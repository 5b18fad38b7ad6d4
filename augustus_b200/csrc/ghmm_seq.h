/*
 * ghmm_seq.h — sequence predicates, k-mer indexing, emissions and signal scores (host + device).
 *
 * Each function names the reference code whose result it reproduces (AUGUSTUS 3.5.0).
 */
#pragma once
#include "ghmm_defs.h"

namespace augb {

struct Seq {
    const uint8_t* c; int L;
    const uint16_t* kf = nullptr; const uint16_t* kr = nullptr; int k1 = 0;   /* optional precomputed (k+1)-mer codes */
    AUGB_HD int at(int p) const { return (p < 0 || p >= L) ? 5 : c[p]; }
    /* n-mer ENDING at q, n <= k+1 (= Seq2Int(n)(seq + q - n + 1)); -1 = invalid nucleotide / out of range */
    AUGB_HD int kmer_end(int q, int n) const {
        if (kf && n <= k1 && q >= 0 && q < L) { unsigned v = kf[q]; if (!(v & 0x8000u)) return (int)(v & ((1u << (2 * n)) - 1u)); }
        return s2i(q - n + 1, n);
    }
    /* reverse-complement n-mer STARTING at p, n <= k+1 (= Seq2Int(n).rc(seq + p)) */
    AUGB_HD int kmer_rc(int p, int n) const {
        if (kr && n <= k1 && p >= 0 && p < L) { unsigned v = kr[p]; if (!(v & 0x8000u)) return (int)(v & ((1u << (2 * n)) - 1u)); }
        return s2irc(p, n);
    }
    /* Seq2Int::operator() (geneticcode.hh:166-173); -1 = InvalidNucleotideError */
    AUGB_HDN int s2i(int p, int n) const {
        int e = 0;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) { int b = at(p + i); if (b > 3) return -1; e = (e << 2) | b; }
        return e;
    }
    /* Seq2Int::rc (geneticcode.hh:174-179) */
    AUGB_HDN int s2irc(int p, int n) const {
        int e = 0;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) { int b = at(p + i); if (b > 3) return -1; e |= (3 - b) << (2 * i); }
        return e;
    }
    AUGB_HD bool is2(int p, int a, int b) const { return at(p) == a && at(p + 1) == b; }
};
enum : int { A_ = 0, C_ = 1, G_ = 2, T_ = 3 };
AUGB_HD int cmpl(int b) { return b < 4 ? 3 - b : b; }

/* statemodel.hh:98-117 without hints; onGenDSS etc. geneticcode.hh:30-74 */
AUGB_HD bool possDSS(const DevModel* m, const Seq& s, int pos) {
    return pos >= 1 && pos <= s.L - 2 && (s.is2(pos, G_, T_) || (m->dss_gc_allowed && s.is2(pos, G_, C_)));
}
AUGB_HD bool possRDSS(const DevModel* m, const Seq& s, int pos) {
    return pos >= 1 && pos <= s.L - 2 && (s.is2(pos - 1, A_, C_) || (m->dss_gc_allowed && s.is2(pos - 1, G_, C_)));
}
AUGB_HD bool possASS(const Seq& s, int pos) { return pos >= 1 && pos <= s.L - 2 && s.is2(pos - 1, A_, G_); }
AUGB_HD bool possRASS(const Seq& s, int pos) { return pos >= 1 && pos <= s.L - 2 && s.is2(pos, C_, T_); }
AUGB_HD bool isStop(const DevModel* m, const Seq& s, int p) { int i = s.kmer_end(p + 2, 3); return i >= 0 && m->isstop[i]; }
AUGB_HD bool isRCStop(const DevModel* m, const Seq& s, int p) { int i = s.kmer_rc(p, 3); return i >= 0 && m->isstop[i]; }

/* ---- single-position emissions ---- */
/* intron content, forward k-mer ending at p (SnippetProbs::getElemSeqProb fwd, statemodel.cc:287-296;
 * IntronModel::emiProbUnderModel geometric branch, intronmodel.cc:895-915) */
AUGB_HD sc_t intron_emi1(const DevModel* m, const Seq& s, int cls, int p) {
    if (p - m->k < 0) return m->log025;
    int pn = s.kmer_end(p, m->k + 1);
    return pn < 0 ? m->log025 : m->iemi[((size_t)cls << (2 * (m->k + 1))) | pn];
}
/* intron content, reverse-complement k-mer starting at p (statemodel.cc:298-307) */
AUGB_HD sc_t intron_emi1r(const DevModel* m, const Seq& s, int cls, int p) {
    if (!(p >= 0 && p + m->k < s.L)) return m->log025;
    int pn = s.kmer_rc(p, m->k + 1);
    return pn < 0 ? m->log025 : m->iemi[((size_t)cls << (2 * (m->k + 1))) | pn];
}
/* 3-periodic exon content (ExonModel::seqProb, exonmodel.cc:1941-1966) */
AUGB_HD sc_t exon_emi1(const DevModel* m, const Seq& s, const sc_t* tab, int cls, int fwd, int f, int p) {
    int pn = fwd ? s.kmer_end(p, m->k + 1) : s.kmer_rc(p, m->k + 1);
    if (pn < 0) return m->probN;
    return tab[(((size_t)cls * 3 + f) << (2 * (m->k + 1))) | pn];
}
/* IGenicModel::emiProbUnderModel (igenicmodel.cc:299-357) for one base; the short-context quirk for
 * columns <= k is tabulated on the host (DevModel::gfirst) */
AUGB_HD int gfirst_off(int j) { return ((1 << (2 * (j + 1))) - 4) / 3; }   /* sum_{i<j} 4^(i+1) */
AUGB_HD sc_t igenic_emi(const DevModel* m, const Seq& s, int cls, int j) {
    if (j > m->k) {
        int pn = s.kmer_end(j, m->k + 1);
        return pn < 0 ? m->log025 : m->gemi[((size_t)cls << (2 * (m->k + 1))) | pn];
    }
    int basek = s.s2i(0, j + 1);
    if (basek < 0) return m->log025;
    return m->gfirst[(size_t)cls * gfirst_off(m->k + 1) + gfirst_off(j) + basek];
}

/* ---- activity mask of a column (see MB_* in ghmm_defs.h); mirrors the abort / early-return tests of
 * intronmodel.cc:553-556,690-714 and exonmodel.cc:1277-1384 that depend on the sequence only ---- */
AUGB_HD unsigned column_mask(const DevModel* m, const Seq& s, int j) {
    unsigned mb = 0; const int L = s.L;
    const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
    if (j - dssw >= 0 && possDSS(m, s, j - m->dss_end - 2 + 1)) mb |= MB_LONGDSS;
    if (j - dssw >= 0 && possRDSS(m, s, j - m->dss_start)) mb |= MB_RLONGDSS;
    if (j - assw - m->ass_up >= 0 && possASS(s, j - m->ass_end)) mb |= MB_LONGASS;
    if (j - assw - m->ass_up >= 0 && possRASS(s, j - m->ass_up - m->ass_start - 2 + 1)) mb |= MB_RLONGASS;
    { int eob = j + m->ass_up + m->ass_start + 2; if (!(eob - 2 + 1 < L - 1 && !possASS(s, eob))) mb |= MB_LESSD; }
    { int eob = j + m->dss_end + 2; if (!(eob - 2 + 1 < L - 1 && !possRDSS(m, s, eob))) mb |= MB_RLESSD; }
    { int dsspos = j + m->dss_start + 1;
      if (j == L - 1 || !((dsspos + 2 - 1 < L && !possDSS(m, s, dsspos)) || j + m->dss_start >= L)) mb |= MB_XDSS; }
    { int sp = j - 3 + 1; if (sp >= 0 && sp <= L - 3 && isStop(m, s, sp)) mb |= MB_XSTOP; }
    { int asspos = j + m->ass_end + 1; if (j == L - 1 || (j + m->ass_end + 2 < L && possRASS(s, asspos))) mb |= MB_XRASS; }
    { int sp = j - m->tiw - 3 + 1; if (sp >= 0) { int pn = s.kmer_rc(sp, 3); if (pn >= 0 && !isneg(m->startp[pn])) mb |= MB_XRSTART; } }
    return mb;
}

}  // namespace augb
