/*
 * ghmm_sweep.h — the semi-Markov Viterbi sweep of one window by one warp (host + device source).
 *
 * Replaces, for one window, the column loop of NAMGene::viterbiAndForward (namgene.cc:244-335) and
 * the StateModel::viterbiForwardAndSampling overrides it calls:
 *     ExonModel   exonmodel.cc:899-1179 (+ endPartEmiProb :1272, notEndPartEmiProb :1417)
 *     IntronModel intronmodel.cc:509-858 (+ emiProbUnderModel :861, aSSProb :1116, dSSProb :1195)
 *     IGenicModel igenicmodel.cc:231-287
 *
 * Design: the reference fills a sparse L x S matrix and finds ~9 of 47 cells non-zero
 * per column, 7 of which are one-base self-loop states.  Here
 *   - the 7 self-loop states (igenic, 6 geometric introns) are never stepped: in exact fixed point
 *     V[j] = A[j] + running-max(entries), so a chain is a short list of change points (ChainCP);
 *   - every other non-zero cell is an Event appended to a per-window log (column-indexed by
 *     evstart[]), which is also the backtrace structure;
 *   - "max over (predecessor, duration)" loops run over compact candidate lists (Cand) holding only
 *     the predecessor cells that are non-zero, lanes taking candidates in parallel and reducing with
 *     shuffle max; content emissions are differences of precomputed prefix sums;
 *   - columns whose sequence context rules out every sparse state are skipped 32 at a time using
 *     the activity mask written by the prep pass.
 * Ties: candidates are reduced with "highest score, then the option the reference's loop meets first",
 * which is what the reference's strict '>' keeps.
 */
#pragma once
#include "ghmm_defs.h"
#include "ghmm_seq.h"
#include "ghmm_warp.h"

namespace augb {

struct WarpState {
    int n_ev;
    int filled;                 /* evstart[] is valid up to this column */
    int cl_n[NCL];
    int eq_cur[6];
    int cp_n[NCHAIN];
    sc_t tilde[NCHAIN];
    sc_t pend_val[NCHAIN];
    int pend_pred[NCHAIN];
    int status;
};

struct Sweep {
    const DevModel* m; WinView w; WarpState* ws; Seq sq; int lane; int cls; int L;

    AUGB_D const sc_t* parr(int c, int which) const { return w.parr + ((size_t)c * PA_PER_CLASS + which) * (size_t)(L + 1); }
    AUGB_D sc_t TR(int a, int s) const { return m->trans[((size_t)cls * m->S + a) * m->S + s]; }

    /* ------------------------------------------------------------ chains */
    AUGB_D const sc_t* chainA(int ch) const { return ch == 0 ? w.AIG : w.AGEO; }
    /* V[e][chain]: last change point with col <= e */
    AUGB_D sc_t chain_value(int ch, int e) const {
        int n = ws->cp_n[ch];
        if (n == 0) return SC_NEG;
        const ChainCP* cp = w.cp[ch];
        int lo = 0, hi = n - 1;          /* invariant: cp[lo].col <= e (cp[0].col is the first entry column) */
        if (cp[0].col > e) return SC_NEG;
        if (cp[hi].col <= e) lo = hi;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cp[mid].col <= e) lo = mid; else hi = mid - 1; }
        return cp[lo].tilde + chainA(ch)[e];
    }
    /* V[e][a] for any state a and any finished column e */
    AUGB_D sc_t lookupV(int a, int e) const {
        int ch = m->st[a].chain;
        if (ch >= 0) return chain_value(ch, e);
        int lo = w.evstart[e], hi = w.evstart[e + 1];
        for (int i = lo; i < hi; i++) if (w.ev[i].state == a) return w.ev[i].V;
        return SC_NEG;
    }

    /* ------------------------------------------------------------ bookkeeping */
    AUGB_D void fill_evstart(int upto) {
        int f = ws->filled, n = ws->n_ev;
        for (int i = f + 1 + lane; i <= upto; i += AUGB_NLANES) w.evstart[i] = n;
        wsync();
        if (lane == 0) ws->filled = upto;
        wsync();
    }
    AUGB_D void cl_append(int list, int col, int state, sc_t V) {
        int n = ws->cl_n[list];
        if (n >= w.cl_cap) { if (lane == 0) ws->status = 8; wsync(); return; }
        if (lane == 0) { Cand c; c.col = col; c.state = state; c.V = V; w.cl[list][n] = c; ws->cl_n[list] = n + 1; }
        wsync();
    }
    /* record a non-zero cell; route it to the structures later columns look back to */
    AUGB_D void emit(int j, int s, sc_t V, int pred, int predbase) {
        int n = ws->n_ev;
        if (n >= w.ev_cap) { if (lane == 0) ws->status = 8; wsync(); return; }
        if (lane == 0) {
            Event e; e.col = j; e.state = (int16_t)s; e.pred = (int16_t)pred; e.predbase = predbase; e.pad = 0; e.V = V;
            w.ev[n] = e; ws->n_ev = n + 1;
        }
        wsync();
        const StateDesc& sd = m->st[s];
        if (sd.kind == K_LONGDSS) {
            if (sd.fwd) cl_append(CL_LD + sd.frame, j, s, V);
            else cl_append(CL_RD + mod3(sd.frame + j - 2 + 3 - m->dss_start), j, s, V);   /* phase = mod3(pf + bobe), bobe = j+1-dss_start */
        } else if (sd.kind == K_LONGASS) {
            if (sd.fwd) cl_append(CL_LA + mod3(sd.frame - (j + 1 - m->ass_end)), j, s, V);   /* phase = mod3(pf - bobe) */
            else cl_append(CL_RA + sd.frame, j, s, V);
        }
        int ch = sd.feeds;
        if (ch >= 0 && j + 1 < L) {
            /* candidate for V[j+1][chain] in tilde coordinates */
            int c1 = w.gc[j + 1], cs = m->chain_state[ch];
            sc_t t_in = m->trans[((size_t)c1 * m->S + s) * m->S + cs], t_self = m->trans[((size_t)c1 * m->S + cs) * m->S + cs];
            if (!isneg(t_in)) {
                sc_t val = V + t_in - t_self - chainA(ch)[j];
                if (lane == 0 && (val > ws->pend_val[ch] || (val == ws->pend_val[ch] && s < ws->pend_pred[ch]))) {
                    ws->pend_val[ch] = val; ws->pend_pred[ch] = s;
                }
                wsync();
            }
        }
    }
    /* end of column j: apply the pending chain entries for column j+1 (ancestors in index order, strict >) */
    AUGB_D void apply_pending(int j) {
        for (int ch = 0; ch < NCHAIN; ch++) {
            sc_t pv = ws->pend_val[ch];
            if (isneg(pv)) continue;
            int pp = ws->pend_pred[ch], self = m->chain_state[ch];
            sc_t cur = ws->tilde[ch];
            bool take = pv > cur || (pv == cur && pp < self);
            int n = ws->cp_n[ch];
            if (take && n >= w.cp_cap) { if (lane == 0) ws->status = 8; take = false; }
            wsync();
            if (lane == 0) {
                if (take) { ChainCP c; c.col = j + 1; c.pred = pp; c.tilde = pv; w.cp[ch][n] = c; ws->cp_n[ch] = n + 1; ws->tilde[ch] = pv; }
                ws->pend_val[ch] = SC_NEG; ws->pend_pred[ch] = 0x7fffffff;
            }
            wsync();
        }
    }

    /* ------------------------------------------------------------ signal scores */
    /* Motif::seqProb (motif.cc:308-331), lanes over motif positions */
    AUGB_D sc_t motif_fwd(const sc_t* tab, int n, int k, int p) const {
        sc_t s = 0; const size_t wd = (size_t)1 << (2 * (k + 1));
        for (int i = lane; i < n; i += AUGB_NLANES) { int pn = sq.s2i(p + i - k, k + 1); s += pn < 0 ? m->log025 : tab[((size_t)cls * n + i) * wd + pn]; }
        return wsum(s);
    }
    AUGB_D sc_t motif_rc(const sc_t* tab, int n, int k, int p) const {
        sc_t s = 0; const size_t wd = (size_t)1 << (2 * (k + 1));
        for (int i = lane; i < n; i += AUGB_NLANES) { int pn = sq.s2irc(p + i, k + 1); s += pn < 0 ? m->log025 : tab[((size_t)cls * n + (n - 1 - i)) * wd + pn]; }
        return wsum(s);
    }
    /* the same, evaluated by one lane alone (inside per-candidate code) */
    AUGB_D sc_t motif_fwd1(const sc_t* tab, int n, int k, int p) const {
        sc_t s = 0; const size_t wd = (size_t)1 << (2 * (k + 1));
        for (int i = 0; i < n; i++) { int pn = sq.s2i(p + i - k, k + 1); s += pn < 0 ? m->log025 : tab[((size_t)cls * n + i) * wd + pn]; }
        return s;
    }
    /* IntronModel::dSSProb (intronmodel.cc:1195-1248) */
    AUGB_D sc_t dSSProb(int base, int fwd) const {
        int nonGT, idx;
        if (fwd) {
            int dsspos = base + m->dss_start;
            if (!possDSS(m, sq, dsspos)) return SC_NEG;
            nonGT = !sq.is2(dsspos, G_, T_);
            int a = sq.s2i(base, m->dss_start), b = sq.s2i(dsspos + 2, m->dss_end);
            if (a < 0 || b < 0) return SC_NEG;
            idx = (a << (2 * m->dss_end)) | b;
        } else {
            int dsspos = base + m->dss_end;
            if (!possRDSS(m, sq, dsspos + 1)) return SC_NEG;
            nonGT = !sq.is2(dsspos, A_, C_);
            int a = 0, b = 0;
            for (int i = m->dss_start - 1; i >= 0; i--) { int c = sq.at(dsspos + 2 + i); if (c > 3) return SC_NEG; a = (a << 2) | (3 - c); }
            for (int i = m->dss_end - 1; i >= 0; i--) { int c = sq.at(base + i); if (c > 3) return SC_NEG; b = (b << 2) | (3 - c); }
            idx = (a << (2 * m->dss_end)) | b;
        }
        return nonGT ? m->dss_pat_non[idx] : m->dss_pat[idx];
    }
    /* IntronModel::aSSProb (intronmodel.cc:1116-1188); warp-cooperative (motif over lanes) */
    AUGB_D sc_t aSSProb(int base, int fwd) const {
        int nonAG, a, b; sc_t motif;
        if (fwd) {
            int asspos = base + m->ass_up + m->ass_start;
            if (!possASS(sq, asspos + 1)) return SC_NEG;
            nonAG = !sq.is2(asspos, A_, G_);
            a = sq.s2i(base + m->ass_up, m->ass_start); b = sq.s2i(asspos + 2, m->ass_end);
            motif = base >= m->assm_k ? motif_fwd(m->assm, m->assm_n, m->assm_k, base) : SC_NEG;
        } else {
            int asspos = base + m->ass_end;
            if (!possRASS(sq, asspos)) return SC_NEG;
            nonAG = !sq.is2(asspos, C_, T_);
            a = 0; b = 0;
            for (int i = m->ass_start - 1; i >= 0; i--) { int c = sq.at(asspos + 2 + i); if (c > 3) { a = -1; break; } a = (a << 2) | (3 - c); }
            for (int i = m->ass_end - 1; i >= 0; i--) { int c = sq.at(base + i); if (c > 3) { b = -1; break; } b = (b << 2) | (3 - c); }
            int motifstart = base + m->ass_start + m->ass_end + 2, motifend = motifstart + m->ass_up;
            motif = motifend + m->assm_k < L ? motif_rc(m->assm, m->assm_n, m->assm_k, motifstart) : (sc_t)m->ass_up * m->log025;
        }
        sc_t pat;
        if (a < 0 || b < 0) pat = m->ass_invalid_pat;
        else { int idx = (a << (2 * m->ass_end)) | b; pat = nonAG ? m->ass_pat_non[idx] : m->ass_pat[idx]; }
        if (isneg(motif) || isneg(pat)) return SC_NEG;
        return motif + pat;
    }

    /* ------------------------------------------------------------ ORF, exonmodel.cc:165-198 */
    AUGB_D int leftmostExonBegin(int frame, int base, int forward) const {
        int pos, n = L;
        if (forward) pos = (frame == 0 || frame == 1) ? base - frame - 3 : base - frame;
        else pos = (frame == 1 || frame == 2) ? base + frame - 5 : base - 2;
        if (pos >= n) pos -= 3 * ((pos - n + 3) / 3);
        int lmb = pos >= 0 ? (forward ? w.nsf[pos] : w.nsr[pos]) + 1 : 0;
        int max_allowed = m->max_exon_len - m->ass_up - m->ass_start - 2 - 2 - m->dss_start;
        if (lmb < base - max_allowed) lmb = base - max_allowed;
        return lmb;
    }

    /* ------------------------------------------------------------ exon emissions */
    AUGB_D sc_t exon_seqProb(int fwd, int left, int right, int frameOfRight) const {   /* exonmodel.cc:1925-1973 */
        if (left > right) return 0;
        if (fwd) { const sc_t* P = parr(cls, PA_PX + mod3(frameOfRight - right)); return P[right + 1] - P[left]; }
        const sc_t* P = parr(cls, PA_PXR + mod3(frameOfRight + right)); return P[right + 1] - P[left];
    }
    AUGB_D sc_t exon_shortProb(const sc_t* tab, int fwd, int left, int right, int frameOfRight) const {   /* :1979-2034 */
        sc_t s = 0;
        for (int p = right; p >= left; p--) {
            int f = fwd ? mod3(frameOfRight - right + p) : mod3(frameOfRight + right - p);
            s += exon_emi1(m, sq, tab, cls, fwd, f, p);
        }
        return s;
    }
    /* ExonModel::endPartEmiProb (exonmodel.cc:1272-1400), no hints; uniform across lanes */
    AUGB_D sc_t endPart(const StateDesc& st, int end) const {
        switch (st.ek) {
        case E_SINGLE: case E_TERMINAL: {
            int sp = end - 3 + 1;
            if (sp < 0 || sp > L - 3 || !isStop(m, sq, sp)) return SC_NEG;
            int a = sq.at(sp), b = sq.at(sp + 1), c = sq.at(sp + 2);
            if (a == T_ && b == A_ && c == A_) return m->ochre;
            if (a == T_ && b == A_ && c == G_) return m->amber;
            if (a == T_ && b == G_ && c == A_) return m->opal;
            return SC_NEG;
        }
        case E_RSINGLE: case E_RINITIAL: {
            int sp = end - m->tiw - 3 + 1;
            if (sp < 0) return SC_NEG;
            int pn = sq.s2irc(sp, 3);
            if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
            sc_t p = m->startp[pn];
            if (sp + 3 + m->tiw - 1 + m->tis_k < L) p += motif_rc(m->tis, m->tis_n, m->tis_k, sp + 3);
            else p = (sc_t)(L - (sp + 3)) * m->log025;
            return p;
        }
        case E_INITIAL: case E_INTERNAL: {
            int dsspos = end + m->dss_start + 1;
            if (end == L - 1) return 0;
            if ((dsspos + 2 - 1 < L && !possDSS(m, sq, dsspos)) || end + m->dss_start >= L ||
                leftmostExonBegin(st.frame - 1, end + m->dss_start, 1) >= end) return SC_NEG;
            return 0;
        }
        default: {
            int asspos = end + m->ass_end + 1;
            if (end == L - 1) return 0;
            if (end + m->ass_end + 2 < L && possRASS(sq, asspos)) return 0;
            return SC_NEG;
        }
        }
    }
    /* ExonModel::notEndPartEmiProb (exonmodel.cc:1417-1859), no hints; evaluated by one lane per candidate */
    AUGB_D sc_t notEndPart(const StateDesc& st, int bos, int right, int frameOfRight) const {
        const int k = m->k, fwd = st.fwd, win = st.frame;
        sc_t beginPart;
        int bobe = bos - st.innerPartOffset;
        switch (st.ek) {
        case E_SINGLE: case E_INITIAL: {
            if (!(bobe >= 0 && bobe < L - 2)) return SC_NEG;
            int pn = sq.s2i(bobe, 3);
            if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
            beginPart = m->startp[pn];
            int tis = bobe - m->tiw;
            if (tis > m->tis_k) beginPart += motif_fwd1(m->tis, m->tis_n, m->tis_k, tis);
            else beginPart += (sc_t)(bos - 3) * m->log025;
            break;
        }
        case E_TERMINAL: case E_INTERNAL:
            if (bos > 0) { if (bobe < 0 || (bobe - 2 >= 0 && !possASS(sq, bobe - 1))) return SC_NEG; beginPart = 0; }
            else if (bos == 0) beginPart = 0;
            else return SC_NEG;
            break;
        case E_RSINGLE: case E_RTERMINAL: {
            if (bobe < 0) return SC_NEG;
            int a = sq.at(bobe), b = sq.at(bobe + 1), c = sq.at(bobe + 2);
            if (a == T_ && b == T_ && c == A_) beginPart = m->ochre;
            else if (a == C_ && b == T_ && c == A_) beginPart = m->amber;
            else if (a == T_ && b == C_ && c == A_) beginPart = m->opal;
            else return SC_NEG;
            if (isneg(beginPart)) return SC_NEG;
            break;
        }
        default:
            if (bos == 0) beginPart = 0;
            else if (bobe < 0 || (bobe - 2 > 0 && !possRDSS(m, sq, bobe - 1))) return SC_NEG;
            else beginPart = 0;
        }
        sc_t rest;
        if (bos > right) rest = -(sc_t)(bos - right - 1) * m->log025;
        else if (right - bos <= k) {
            int l = right - bos;
            int pn = fwd ? sq.s2i(bos, l + 1) : sq.s2irc(bos, l + 1);
            if (pn < 0) rest = (sc_t)(l + 1) * m->probN;
            else { int f = fwd ? frameOfRight : mod3(frameOfRight + right - bos); rest = m->xpls[l][(((size_t)cls * 3 + f) << (2 * (l + 1))) | pn]; }
        } else {
            int endOfStart = bos + k - 1, beginOfInitP = right - (k - 1);
            if (k == 0) rest = 0;
            else {
                int pn = fwd ? sq.s2i(bos, k) : sq.s2irc(beginOfInitP, k);
                if (pn < 0) rest = (sc_t)k * m->probN;
                else {
                    int f = fwd ? mod3(frameOfRight - right + endOfStart) : mod3(frameOfRight + right - beginOfInitP);
                    rest = m->xpls[k - 1][(((size_t)cls * 3 + f) << (2 * k)) | pn];
                }
            }
            if (isneg(rest)) return SC_NEG;
            int endOfInitial, beginOfTerm, endOfTerm, beginOfInitial;
            switch (st.ek) {
            case E_SINGLE:
                endOfInitial = endOfStart + m->init_len; if (endOfInitial > right) endOfInitial = right;
                rest += exon_shortProb(m->xinit, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                      + exon_seqProb(1, endOfInitial + 1, right, frameOfRight);
                break;
            case E_INITIAL:
                endOfInitial = endOfStart + m->init_len;
                if (endOfInitial > right) { endOfInitial = right; beginOfTerm = right + 1; }
                else { beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfInitial) beginOfTerm = right + 1; }
                rest += exon_shortProb(m->xinit, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                      + exon_seqProb(1, endOfInitial + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                      + exon_shortProb(m->xet, 1, beginOfTerm, right, frameOfRight);
                break;
            case E_INTERNAL:
                beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfStart) beginOfTerm = right + 1;
                rest += exon_seqProb(1, endOfStart + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                      + exon_shortProb(m->xet, 1, beginOfTerm, right, frameOfRight);
                break;
            case E_TERMINAL:
                rest += exon_seqProb(1, endOfStart + 1, right, frameOfRight);
                break;
            case E_RSINGLE:
                beginOfInitial = beginOfInitP - m->init_len; if (beginOfInitial < bos) beginOfInitial = bos;
                rest += exon_shortProb(m->xinit, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_seqProb(0, bos, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)));
                break;
            case E_RINITIAL:
                beginOfInitial = beginOfInitP - m->init_len;
                if (beginOfInitial < bos) { beginOfInitial = bos; endOfTerm = bos - 1; }
                else { endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitial) endOfTerm = bos - 1; }
                rest += exon_shortProb(m->xinit, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_seqProb(0, endOfTerm + 1, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)))
                      + exon_shortProb(m->xet, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
                break;
            case E_RINTERNAL:
                endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitP) endOfTerm = bos - 1;
                rest += exon_seqProb(0, endOfTerm + 1, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_shortProb(m->xet, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
                break;
            default:
                rest += exon_seqProb(0, bos, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)));
            }
        }
        if (isneg(rest)) return SC_NEG;
        int eobe = right + st.innerPartEndOffset, len = eobe - bobe + 1;
        if (len < 1 || len >= m->n_ld_exon) return SC_NEG;
        sc_t lp;
        switch (st.ek) {
        case E_SINGLE: case E_RSINGLE: lp = len % 3 == 0 ? m->ld_single[len] : SC_NEG; break;
        case E_INITIAL: lp = (len % 3 == win && len > 2) ? m->ld_initial[len] : SC_NEG; break;
        case E_RINITIAL: lp = len > 2 ? m->ld_initial[len] : SC_NEG; break;
        case E_INTERNAL: case E_RINTERNAL: lp = m->ld_internal[len]; break;
        case E_TERMINAL: lp = m->ld_terminal[len]; break;
        default: lp = mod3(2 - len) == win ? m->ld_terminal[len] : SC_NEG;
        }
        if (isneg(lp)) return SC_NEG;
        return beginPart + rest + (m->log3 + lp);
    }

    /* ExonModel::viterbiForwardAndSampling (exonmodel.cc:899-1179) for state s ending at column j */
    AUGB_D void exon_eval(int s, int j) {
        const StateDesc& st = m->st[s]; const int fwd = st.fwd, win = st.frame;
        sc_t ep = endPart(st, j);
        int eobe = j + st.baseOffset, right = eobe - st.innerPartEndOffset;
        if (isneg(ep) || right < 0) return;
        int frameOfRight = fwd ? mod3(win - (eobe + 1) + right) : mod3(win + eobe + 1 - right);
        int eons = (st.ek == E_TERMINAL || st.ek == E_SINGLE) ? eobe - 3 : eobe;
        if (eons > L - 1) eons = L - 1;
        int feons = fwd ? mod3(win - 1 - eobe + eons) : mod3(win + 1 + eobe - eons);
        int ORFleft = leftmostExonBegin(feons, eons, fwd);
        int startMax = eobe + st.innerPartOffset - m->min_exon_length + 1, startMin;
        if (st.ek == E_RTERMINAL || st.ek == E_RSINGLE) startMin = startMax = ORFleft + 2;
        else {
            startMin = ORFleft <= 0 ? 0 : ORFleft + st.innerPartOffset;
            if (startMax > j + st.beginPartLen) startMax = j + st.beginPartLen;
        }
        /* per-lane best */
        sc_t best = SC_NEG; int bkey = -0x7fffffff, bpred = -1, bbase = -1;
        #define AUGB_CONSIDER(score_, bos_, a_, eop_) do { sc_t sc__ = (score_); int key__ = (bos_) * 128 + (127 - (a_)); \
            if (sc__ > best || (sc__ == best && key__ > bkey)) { best = sc__; bkey = key__; bpred = (a_); bbase = (eop_); } } while (0)
        if (st.ek == E_INTERNAL || st.ek == E_TERMINAL || st.ek == E_RINTERNAL || st.ek == E_RINITIAL) {
            /* predecessors are longass_f (fwd) / rlongdss_f (rev) cells, kept per reading-frame phase */
            int list = fwd ? CL_LA + mod3(win - eobe - 1) : CL_RD + mod3(win + eobe + 1);
            const Cand* cl = w.cl[list]; int n = ws->cl_n[list];
            int lo = startMin < 1 ? 1 : startMin;
            bool done = false;
            for (int base_i = n - 1; base_i >= 0 && !done; base_i -= AUGB_NLANES) {
                int i = base_i - lane; bool below = false;
                if (i >= 0) {
                    Cand c = cl[i]; int bos = c.col + 1;
                    if (bos < lo) below = true;
                    else if (bos <= startMax && c.col < j) {
                        sc_t nep = notEndPart(st, bos, right, frameOfRight);
                        if (!isneg(nep)) {
                            int bobe = bos - st.innerPartOffset, len = eobe - bobe + 1, pf = m->st[c.state].frame;
                            sc_t t = TR(c.state, s);
                            if (!isneg(t) && win == mod3(fwd ? pf + len : pf - len)) AUGB_CONSIDER(c.V + (t + ep + nep), bos, c.state, c.col);
                        }
                    }
                } else below = true;
                done = wballot(below) != 0;
            }
            if (startMin == 0 && lane == 0) {      /* left-truncated exon: bos = 0 reads column 0 (exonmodel.cc:1067-1068) */
                sc_t nep = notEndPart(st, 0, right, frameOfRight);
                if (!isneg(nep)) {
                    int bobe = 0 - st.innerPartOffset, len = eobe - bobe + 1;
                    for (int i = 0; i < st.nanc; i++) {
                        int a = st.anc[i]; sc_t pv = m->init[a], t = TR(a, s); int pf = m->st[a].frame;
                        if (isneg(pv) || isneg(t)) continue;
                        if (win == mod3(fwd ? pf + len : pf - len)) AUGB_CONSIDER(pv + (t + ep + nep), 0, a, -1);
                    }
                }
            }
        } else if (st.ek == E_INITIAL || st.ek == E_SINGLE) {
            /* predecessor igenic; the length must satisfy len % 3 == win (0 for single): step 3 over bos */
            int a = st.anc[0]; sc_t t = TR(a, s);
            int want = mod3(eobe + 1 - (st.ek == E_SINGLE ? 0 : win));          /* bobe mod 3 */
            int b0 = startMax; while (mod3(b0 - st.innerPartOffset) != want) b0--;
            if (!isneg(t))
                for (int bos = b0 - 3 * lane; bos >= startMin; bos -= 3 * AUGB_NLANES) {
                    int bobe = bos - 3;
                    if (bobe < 0 || bobe >= L - 2) continue;
                    int pn = sq.s2i(bobe, 3);
                    if (pn < 0 || isneg(m->startp[pn])) continue;
                    int eop = bos - st.beginPartLen - 1;
                    if (eop >= L) continue;
                    sc_t nep = notEndPart(st, bos, right, frameOfRight);
                    if (isneg(nep)) continue;
                    sc_t pv = lookupV(a, eop >= 0 ? eop : 0);
                    if (isneg(pv)) continue;
                    AUGB_CONSIDER(pv + (t + ep + nep), bos, a, eop);
                }
        } else {   /* E_RTERMINAL, E_RSINGLE: exactly one candidate */
            if (lane == 0) {
                int bos = startMin, eop = bos - st.beginPartLen - 1;
                sc_t nep = notEndPart(st, bos, right, frameOfRight);
                if (!isneg(nep) && eop < L)
                    for (int i = 0; i < st.nanc; i++) {
                        int a = st.anc[i]; sc_t t = TR(a, s); if (isneg(t)) continue;
                        sc_t pv = lookupV(a, eop >= 0 ? eop : 0); if (isneg(pv)) continue;
                        AUGB_CONSIDER(pv + (t + ep + nep), bos, a, eop);
                    }
            }
        }
        #undef AUGB_CONSIDER
        int wl = wargbest(best, bkey);
        if (wl < 0) return;
        sc_t V = wbcast64(best, wl); int pred = wbcast(bpred, wl), pbase = wbcast(bbase, wl);
        emit(j, s, V, pred, pbase);
    }

    /* ------------------------------------------------------------ intron states */
    /* fixed-length states longdss / longass (intronmodel.cc:686-786) for the three frames of one strand */
    AUGB_D void fixed_eval(int kind, int dir, int j) {
        const int fwd = !dir;
        const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
        int eop; sc_t emi;
        if (kind == K_LONGDSS) { eop = j - dssw; emi = dSSProb(j - dssw + 1, fwd); }
        else { eop = j - assw - m->ass_up; emi = aSSProb(j - assw - m->ass_up + 1, fwd); }
        if (eop < 0 || isneg(emi)) return;
        for (int f = 0; f < 3; f++) {
            int s = kind == K_LONGDSS ? m->r_longdss[dir][f] : m->r_longass[dir][f];
            if (s < 0) continue;
            const StateDesc& st = m->st[s];
            sc_t best = SC_NEG; int bpred = -1;
            for (int i = 0; i < st.nanc; i++) {
                int a = st.anc[i]; sc_t t = TR(a, s); if (isneg(t)) continue;
                sc_t pv = lookupV(a, eop); if (isneg(pv)) continue;
                sc_t pp = pv + (t + emi);
                if (pp > best) { best = pp; bpred = a; }
            }
            if (!isneg(best)) emit(j, s, best, bpred, eop);
        }
    }
    /* equalD / requalD (intronmodel.cc:695-698, 889-894): fires dStateLen columns after a longdss / rlongass cell */
    AUGB_D void equald_eval(int dir, int f, int j) {
        int s = m->r_equald[dir][f]; int list = (dir ? CL_RA : CL_LD) + f;
        int cur = ws->eq_cur[dir * 3 + f];
        Cand c = w.cl[list][cur];
        wsync();
        if (lane == 0) ws->eq_cur[dir * 3 + f] = cur + 1;
        wsync();
        if (s < 0) return;
        int eop = j - m->dStateLen;
        const sc_t* P = parr(cls, PA_PI);                 /* forward k-mers also for requalD (intronmodel.cc:1046-1108) */
        sc_t emi = P[j + 1] - P[eop + 1];
        sc_t t = TR(c.state, s);
        if (isneg(t)) return;
        emit(j, s, c.V + (t + emi), c.state, eop);
    }
    /* lessD / rlessD (intronmodel.cc:540-629, 924-1000): max over the longdss / rlongass cells of the last dStateLen columns */
    AUGB_D void lessd_eval(int dir, int j) {
        const int fwd = !dir;
        int eob = fwd ? j + m->ass_up + m->ass_start + 2 : j + m->dss_end + 2;
        int lme = j - m->dStateLen; if (lme < 0) lme = 0;
        const sc_t* P = parr(cls, fwd ? PA_PI : PA_PIR);
        for (int f = 0; f < 3; f++) {
            int s = m->r_lessd[dir][f]; if (s < 0) continue;
            int list = (dir ? CL_RA : CL_LD) + f;
            const Cand* cl = w.cl[list]; int n = ws->cl_n[list];
            bool spl = !(fwd && f == 0) && !(!fwd && f == 2);
            int cod0 = 4, cod1 = 4, cod2 = 4;
            if (spl && eob < L - 2) {
                if (fwd && f == 1) { cod1 = sq.at(eob + 1); cod2 = sq.at(eob + 2); }
                else if (fwd && f == 2) { cod2 = sq.at(eob + 1); }
                else if (!fwd && f == 0) { cod0 = cmpl(sq.at(eob + 1)); }
                else if (!fwd && f == 1) { cod0 = cmpl(sq.at(eob + 2)); cod1 = cmpl(sq.at(eob + 1)); }
            }
            sc_t best = SC_NEG; int bkey = -0x7fffffff, bpred = -1;
            bool done = false;
            for (int base_i = n - 1; base_i >= 0 && !done; base_i -= AUGB_NLANES) {
                int i = base_i - lane; bool below = false;
                if (i >= 0) {
                    Cand c = cl[i]; int e = c.col;
                    if (e < lme) below = true;
                    else if (e < j) {
                        int begin = e + 1, bob; bool ok = true;
                        if (fwd) { bob = begin - m->dss_end - 2; if (bob >= 0 && !possDSS(m, sq, bob)) ok = false; }
                        else { bob = begin - (m->ass_up + m->ass_start + 2); if (bob >= 0 && !possRASS(sq, bob)) ok = false; }
                        if (ok && spl && bob > 1) {
                            int c0 = cod0, c1 = cod1, c2 = cod2;
                            if (fwd && f == 1) c0 = sq.at(bob - 1);
                            else if (fwd && f == 2) { c0 = sq.at(bob - 2); c1 = sq.at(bob - 1); }
                            else if (!fwd && f == 0) { c1 = cmpl(sq.at(bob - 1)); c2 = cmpl(sq.at(bob - 2)); }
                            else if (!fwd && f == 1) { c2 = cmpl(sq.at(bob - 1)); }
                            if (c0 < 4 && c1 < 4 && c2 < 4 && m->isstop[(c0 << 4) | (c1 << 2) | c2]) ok = false;
                        }
                        int ilen = eob - bob + 1;
                        if (ok && !(ilen > m->d || ilen < 0 || ilen >= m->n_ld_intron)) {
                            sc_t ld = m->ld_intron[ilen], t = TR(c.state, s);
                            if (!isneg(ld) && !isneg(t)) {
                                sc_t sc = c.V + (t + (ld + (P[j + 1] - P[begin])));
                                if (sc > best || (sc == best && e > bkey)) { best = sc; bkey = e; bpred = c.state; }
                            }
                        }
                    }
                } else below = true;
                done = wballot(below) != 0;
            }
            int wl = wargbest(best, bkey);
            if (wl < 0) continue;
            emit(j, s, wbcast64(best, wl), wbcast(bpred, wl), wbcast(bkey, wl));
        }
    }

    /* ------------------------------------------------------------ one column */
    AUGB_D void process_column(int j, unsigned mb, unsigned eqbits) {
        fill_evstart(j);
        cls = w.gc[j];
        if (mb & MB_LESSD) lessd_eval(0, j);
        if (mb & MB_RLESSD) lessd_eval(1, j);
        for (int q = 0; q < 6; q++) if (eqbits & (1u << q)) equald_eval(q / 3, q % 3, j);
        if (mb & MB_LONGDSS) fixed_eval(K_LONGDSS, 0, j);
        if (mb & MB_RLONGDSS) fixed_eval(K_LONGDSS, 1, j);
        if (mb & MB_LONGASS) fixed_eval(K_LONGASS, 0, j);
        if (mb & MB_RLONGASS) fixed_eval(K_LONGASS, 1, j);
        if (mb & MB_XSTOP) {
            if (m->r_single >= 0) exon_eval(m->r_single, j);
            if (m->r_terminal >= 0) exon_eval(m->r_terminal, j);
        }
        if (mb & MB_XDSS) for (int f = 0; f < 3; f++) {
            if (m->r_initial[f] >= 0) exon_eval(m->r_initial[f], j);
            if (m->r_internal[f] >= 0) exon_eval(m->r_internal[f], j);
        }
        if (mb & MB_XRSTART) {
            if (m->r_rsingle >= 0) exon_eval(m->r_rsingle, j);
            if (m->r_rinitial >= 0) exon_eval(m->r_rinitial, j);
        }
        if (mb & MB_XRASS) for (int f = 0; f < 3; f++) {
            if (m->r_rinternal[f] >= 0) exon_eval(m->r_rinternal[f], j);
            if (m->r_rterminal[f] >= 0) exon_eval(m->r_rterminal[f], j);
        }
        apply_pending(j);
    }

    /* ------------------------------------------------------------ whole window */
    AUGB_D void run() {
        L = w.L; sq.c = w.code; sq.L = L; lane = lane_id();
        if (lane == 0) {
            ws->n_ev = 0; ws->filled = -1; ws->status = 0;
            for (int i = 0; i < NCL; i++) ws->cl_n[i] = 0;
            for (int i = 0; i < 6; i++) ws->eq_cur[i] = 0;
            for (int i = 0; i < NCHAIN; i++) { ws->cp_n[i] = 0; ws->tilde[i] = SC_NEG; ws->pend_val[i] = SC_NEG; ws->pend_pred[i] = 0x7fffffff; }
        }
        wsync();
        cls = w.gc[0];
        fill_evstart(0);
        /* column 0 = initial probabilities (NAMGene::setStatesInitialProbs, namgene.cc:144-150) */
        /* a window without a single a/c/g/t is all intergenic: viterbi[j][synch] = viterbi[j-1][synch]/4,
         * every other state erased (namgene.cc:205-226); prep wrote AIG[j] = j*log(1/4) and an empty mask */
        const bool alln = (*w.flags & WF_ALLN) != 0;
        for (int s = 0; s < m->S; s++) {
            sc_t v = m->init[s]; if (isneg(v)) continue;
            int ch = m->st[s].chain;
            if (alln && ch != 0) continue;
            if (ch >= 0) {
                if (lane == 0) { ChainCP c; c.col = 0; c.pred = -1; c.tilde = v; w.cp[ch][0] = c; ws->cp_n[ch] = 1; ws->tilde[ch] = v; }
                wsync();
            } else {
                int n = ws->n_ev;
                if (lane == 0) { Event e; e.col = 0; e.state = (int16_t)s; e.pred = -1; e.predbase = -1; e.pad = 0; e.V = v; w.ev[n] = e; ws->n_ev = n + 1; }
                wsync();
                const StateDesc& sd = m->st[s];
                if (sd.kind == K_LONGDSS && sd.fwd) cl_append(CL_LD + sd.frame, 0, s, v);
                if (sd.kind == K_LONGASS && !sd.fwd) cl_append(CL_RA + sd.frame, 0, s, v);
            }
        }
        for (int j0 = 1; j0 < L; j0 += 32) {
            /* static activity of the next 32 columns + equalD columns that fall into them */
            unsigned act = 0; unsigned mymask = 0;
#if defined(__CUDA_ARCH__)
            { int j = j0 + lane; mymask = j < L ? w.mask[j] : 0u; act = wballot(mymask != 0); }
#else
            unsigned maskbuf[32];
            for (int t = 0; t < 32; t++) { int j = j0 + t; maskbuf[t] = j < L ? w.mask[j] : 0u; if (maskbuf[t]) act |= 1u << t; }
#endif
            unsigned eqcol[6];
            for (int q = 0; q < 6; q++) {
                eqcol[q] = 0;
                if (m->r_equald[q / 3][q % 3] < 0) continue;
                int list = (q / 3 ? CL_RA : CL_LD) + q % 3; int cur = ws->eq_cur[q], n = ws->cl_n[list];
                for (int i = cur; i < n; i++) {
                    int due = w.cl[list][i].col + m->dStateLen;
                    if (due >= j0 + 32) break;
                    if (due >= j0 && due < L) { eqcol[q] |= 1u << (due - j0); act |= 1u << (due - j0); }
                }
            }
            while (act) {
                int t = wffs(act); act &= act - 1;
                int j = j0 + t;
                unsigned eqbits = 0;
                for (int q = 0; q < 6; q++) if (eqcol[q] & (1u << t)) eqbits |= 1u << q;
#if defined(__CUDA_ARCH__)
                unsigned mb = (unsigned)wbcast((int)mymask, t);
#else
                unsigned mb = maskbuf[t];
#endif
                process_column(j, mb, eqbits);
                /* cells written in this chunk can schedule equalD columns only >= dStateLen later */
            }
            if (ws->status) break;
        }
        fill_evstart(L);
        if (lane == 0) {
            *w.out_n_ev = ws->n_ev; *w.out_status = ws->status;
            for (int i = 0; i < NCHAIN; i++) w.out_ncp[i] = ws->cp_n[i];
        }
        wsync();
    }
};

}  // namespace augb
