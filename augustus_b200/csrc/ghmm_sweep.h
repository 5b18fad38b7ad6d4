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
#include "ghmm_signal.h"
#include "ghmm_warp.h"

namespace augb {

/* a cell of the current column that has not been published yet (Sweep::emit / commit_cells) */
struct PendCell { int16_t s, pred; int32_t base; sc_t V; };
constexpr int CB_CAP = 32;
struct WarpState {
    PendCell cb[CB_CAP]; double cbF[CB_CAP];
    int n_ev;
    int filled;                 /* evstart[] is valid up to this column */
    int cl_n[NCL];
    int eq_cur[6];
    int cp_n[NCHAIN];
    sc_t tilde[NCHAIN];
    sc_t pend_val[NCHAIN];
    int pend_pred[NCHAIN];
    int status;
    int any_pend;
    unsigned snip_cnt[2];       /* entries allocated so far per strand (ids start at 1) */
    int n_sx;                   /* entries of WinView::snipx */
    /* forward pass */
    int fcp_n[NCHAIN];
    double ftilde[NCHAIN];      /* ln F[j][chain] - A[j] of the current column */
    Lse pend_f[NCHAIN];         /* log-sum of the entries into column j+1 */
};

/* FWD = also fill the forward (sum) values needed by posterior sampling: the same candidates, log-sum-exp beside max
 * (exonmodel.cc:1094-1101, intronmodel.cc:609-617, igenicmodel.cc:250-257) */
/* UTR = the model has UtrModel states: compiled as a separate kernel so that the 47-state kernel keeps its code size */
template <bool FWD, bool UTR = false>
struct SweepT {
    const DevModel* m; WinView w; WarpState* ws; Seq sq; int lane; int cls; int L;
    const sc_t* trc;            /* transition matrix of the current column's GC class */
    /* sampling step (getSampledPath): when opt != nullptr the routines below list every option of state `only` instead of
     * reducing and recording a cell */
    SampleOpt* opt = nullptr; int* nopt = nullptr; int opt_cap = 0; int only = -1;
    int ncb = 0;                /* cells of the current column waiting in WarpState::cb (same value in every lane) */
    AUGB_D void push_opt(bool have, double lp, int ord, int pred, int eop) {
        unsigned b = wballot(have);
        int base = *nopt;
        if (have) {
            int k = base + wpopc(b & ((1u << lane) - 1u));
            if (k < opt_cap) { SampleOpt o; o.lp = lp; o.ord = ord; o.pred = pred; o.eop = eop; o.pad = 0; opt[k] = o; }
        }
        wsync();
        if (lane == 0) *nopt = base + wpopc(b);
        wsync();
    }

    AUGB_D const sc_t* parr(int c, int which) const { return w.parr_c[c] + (size_t)which * (size_t)(L + 1); }
    AUGB_D sc_t TR(int a, int s) const { return trc[a * m->S + s]; }
    AUGB_D void set_class(int c) { cls = c; trc = m->trans + (size_t)c * m->S * m->S; }

    /* ------------------------------------------------------------ chains */
    /* A[e] of a chain: prefix sum of (emission + self transition); the UTR-intron chains share the intron emission prefix and a
     * class-independent self transition (checked when the model is built) */
    AUGB_D sc_t chainAv(int ch, int e) const { return ch == 0 ? w.AIG[e] : (!UTR || ch < CH_UTR) ? w.AGEO[e] : w.AINT[e] + (sc_t)e * (ch < CH_NC ? m->utr_tself : m->nc_tself); }
    /* V[e][chain]: last change point with col <= e */
    AUGB_D sc_t chain_value(int ch, int e) const {
        int n = ws->cp_n[ch];
        if (n == 0) return SC_NEG;
        const ChainCP* cp = w.cp(ch);
        int lo = 0, hi = n - 1;          /* invariant: cp[lo].col <= e (cp[0].col is the first entry column) */
        if (cp[0].col > e) return SC_NEG;
        if (cp[hi].col <= e) lo = hi;
        AUGB_ROLLED
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cp[mid].col <= e) lo = mid; else hi = mid - 1; }
        return cp[lo].tilde + chainAv(ch, e);
    }
    /* V[e][a] for any state a and any finished column e */
    AUGB_D sc_t lookupV(int a, int e) const {
        int ch = m->st[a].chain;
        if (ch >= 0) return chain_value(ch, e);
        int lo = w.evstart[e], hi = w.evstart[e + 1];
        AUGB_ROLLED
        for (int i = lo; i < hi; i++) if (w.ev[i].state == a) return w.ev[i].V;
        return SC_NEG;
    }

    AUGB_D static double sc2d(sc_t v) { return (double)v * (1.0 / (double)((sc_t)1 << FRAC_BITS)); }
    /* ln of a heated transition x emission product: LLDouble::heated() raises it to (8 - temperature) / 8 before it enters a forward sum or
     * an option list (exonmodel.cc:1095, intronmodel.cc:610, igenicmodel.cc:251, utrmodel.cc:978, ncmodel.cc:274); heat = 1 leaves the bits alone */
    AUGB_D double te2d(sc_t v) const { return sc2d(v) * m->heat; }
    /* ln forward[e][a] (-1e308 = zero) */
    AUGB_D double chain_fvalue(int ch, int e) const {
        int n = ws->fcp_n[ch];
        if (n == 0) return -1e308;
        const FChainCP* cp = w.fcp(ch);
        int lo = 0, hi = n - 1;
        if (cp[0].col > e) return -1e308;
        if (cp[hi].col <= e) lo = hi;
        AUGB_ROLLED
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cp[mid].col <= e) lo = mid; else hi = mid - 1; }
        return cp[lo].ft + te2d(chainAv(ch, e));
    }
    AUGB_D double lookupF(int a, int e) const {
        int ch = m->st[a].chain;
        if (ch >= 0) return chain_fvalue(ch, e);
        int lo = w.evstart[e], hi = w.evstart[e + 1];
        AUGB_ROLLED
        for (int i = lo; i < hi; i++) if (w.ev[i].state == a) return w.evF[i];
        return -1e308;
    }

    /* ------------------------------------------------------------ bookkeeping */
    AUGB_D void fill_evstart(int upto) {
        int f = ws->filled, n = ws->n_ev;
        AUGB_ROLLED
        for (int i = f + 1 + lane; i <= upto; i += AUGB_NLANES) w.evstart[i] = n;
        wsync();
        if (lane == 0) ws->filled = upto;
        wsync();
    }
    AUGB_D void cl_append(int list, int col, int state, sc_t V, double F, sc_t G = 0) {       /* lane 0 only; caller syncs */
        int n = ws->cl_n[list];
        if (n >= w.cl_cap) { ws->status = 8; return; }
        Cand c; c.col = col; c.state = state; c.V = V + G; w.cl(list)[n] = c; ws->cl_n[list] = n + 1;
        if (FWD) { w.clF(list)[n] = F; if (UTR && list >= NCL_BASE) w.clG(list)[n] = G; }
    }
    AUGB_D const sc_t* usegp(int g) const { return w.useg + (size_t)g * (size_t)(L + 1); }
    /* a cell (j, s) whose state has a one-base transition into a self-loop chain: candidate for column j+1 of that chain,
     * in tilde coordinates.  feed_term: the chain (or -1) and what to add to V (and to ln F) */
    AUGB_D int feed_term(int j, int s, sc_t* delta) const {
        int ch = m->st[s].feeds;
        if (ch < 0 || j + 1 >= L) return -1;
        int c1 = w.gc[j + 1], cs = m->chain_state[ch];
        const sc_t* T1 = m->trans + (size_t)c1 * m->S * m->S;
        sc_t t_in = T1[s * m->S + cs], t_self = T1[cs * m->S + cs];
        if (isneg(t_in)) return -1;
        *delta = t_in - t_self - chainAv(ch, j);
        return ch;
    }
    AUGB_D void feed_merge(int ch, int s, sc_t val) {          /* one writer per chain at a time */
        if (val > ws->pend_val[ch] || (val == ws->pend_val[ch] && s < ws->pend_pred[ch])) { ws->pend_val[ch] = val; ws->pend_pred[ch] = s; }
    }
    AUGB_D void feed_chain(int j, int s, sc_t V, double F) {    /* (lane 0 only) */
        sc_t delta; const int ch = feed_term(j, s, &delta);
        if (ch < 0) return;
        feed_merge(ch, s, V + delta);
        ws->any_pend |= 1 << ch;          /* bit per chain with a pending entry */
        if (FWD) ws->pend_f[ch].add(F + te2d(delta));
    }
    /* the candidate lists later columns look a new cell (j, s) up in: returns the first list (-1: none), *l2 a second one (UTR) */
    AUGB_D int cell_lists(int j, int s, sc_t* G1, int* l2, sc_t* G2) const {
        const StateDesc& sd = m->st[s];
        *l2 = -1; *G1 = 0; *G2 = 0;
        if (sd.kind == K_LONGDSS) return sd.fwd ? CL_LD + sd.frame : CL_RD + mod3(sd.frame + j + 1 - m->dss_start);      /* phase = mod3(pf + bobe), bobe = j+1-dss_start */
        if (sd.kind == K_LONGASS) return sd.fwd ? CL_LA + mod3(sd.frame - (j + 1 - m->ass_end)) : CL_RA + sd.frame;      /* phase = mod3(pf - bobe) */
        if (UTR && sd.kind == K_EXON) {
            /* cells the 3' UTR (forward) / 5' UTR (reverse) states look back to */
            if (sd.ek == E_SINGLE || sd.ek == E_TERMINAL) { *G1 = -usegp(US_3)[j]; return CL_X3; }
            if (sd.ek == E_RSINGLE || sd.ek == E_RINITIAL) { *G1 = -usegp(US_RINIT5)[j]; *l2 = CL_XRT; *G2 = -usegp(US_R5)[j]; return CL_XRS; }
        }
        return -1;
    }
    /* A new non-zero cell.  All lanes hold the same arguments.  The cell waits in the column buffer (WarpState::cb) until
     * commit_cells publishes the cells of the column together, one lane per cell: nothing in a column reads a non-chain cell of the
     * same column (candidate loops take entries with col < j only; a predecessor end AT the current column is read from a chain). */
    AUGB_EMIT void emit(int j, int s, sc_t V, int pred, int predbase, double F = 0) {
        if (lane == 0) {
            PendCell c; c.s = (int16_t)s; c.pred = (int16_t)pred; c.base = predbase; c.V = V;
            ws->cb[ncb] = c;
            if (FWD) ws->cbF[ncb] = F;
        }
        if (++ncb == CB_CAP) commit_cells(j);
    }
    /* publish the buffered cells of column j in the order they were found: events by rank, candidate lists with one rank per list,
     * chain entries by a per-chain arg-max (ancestors in index order on ties, like feed_chain) */
    AUGB_D void commit_cells(int j) {
        if (ncb == 0) return;
        wsync();                                        /* the buffer is visible */
        AUGB_ROLLED
        for (int i0 = 0; i0 < ncb; i0 += AUGB_NLANES) {
            const int i = i0 + lane; const bool have = i < ncb;
            const unsigned below = (1u << lane) - 1u;
            const int nr = ncb - i0 < AUGB_NLANES ? ncb - i0 : AUGB_NLANES;
            const int base = ws->n_ev;
            const bool room = base + nr <= w.ev_cap, mine = have && room;
            int cs = 0; sc_t cV = 0; double cF = 0;
            int l1 = -1, l2 = -1, fch = -1; sc_t G1 = 0, G2 = 0, fdelta = 0;
            if (mine) {
                const PendCell c = ws->cb[i]; cs = c.s; cV = c.V;
                Event e; e.state = c.s; e.pred = c.pred; e.predbase = c.base; e.V = c.V;
                w.ev[base + lane] = e;
                if (FWD) { cF = ws->cbF[i]; w.evF[base + lane] = cF; }
                l1 = cell_lists(j, cs, &G1, &l2, &G2);
                fch = feed_term(j, cs, &fdelta);
            }
            wsync();                                    /* every lane has read n_ev */
            if (lane == 0) { if (room) ws->n_ev = base + nr; else ws->status = 8; }
            AUGB_ROLLED
            for (int pass = 0; pass < (UTR ? 2 : 1); pass++) {
                const int li = pass ? l2 : l1; const sc_t G = pass ? G2 : G1;
                if (wballot(li >= 0) == 0) continue;
                const unsigned same = wmatch(li >= 0 ? li : -1 - lane);      /* the cells that go to the same list take consecutive entries */
                const bool ap = li >= 0;
                int n0 = 0, r = 0, cnt = 0; bool fits = false;
                if (ap) {
                    n0 = ws->cl_n[li]; r = wpopc(same & below); cnt = wpopc(same); fits = n0 + cnt <= w.cl_cap;
                    if (!fits) ws->status = 8;
                    else {
                        Cand c; c.col = j; c.state = cs; c.V = cV + G; w.cl(li)[n0 + r] = c;
                        if (FWD) { w.clF(li)[n0 + r] = cF; if (UTR && li >= NCL_BASE) w.clG(li)[n0 + r] = G; }
                    }
                }
                wsync();                                /* every lane of a list has read its length */
                if (ap && r == 0 && fits) ws->cl_n[li] = n0 + cnt;
                wsync();
            }
            unsigned fm = wballot(fch >= 0);
            if (FWD) {
                /* the forward sums are added in the order the cells were found */
                AUGB_ROLLED
                while (fm) {
                    const int b = wffs(fm); fm &= fm - 1;
                    if (lane == b) { feed_merge(fch, cs, cV + fdelta); ws->any_pend |= 1 << fch; ws->pend_f[fch].add(cF + te2d(fdelta)); }
                    wsync();
                }
            } else {
                AUGB_ROLLED
                while (fm) {
                    const int b = wffs(fm); const int ch0 = wbcast(fch, b);
                    const bool in = fch == ch0;
                    const sc_t v = in ? cV + fdelta : SC_NEG;
                    const sc_t mx = wmax(v);
                    const int ps = wmini(in && v == mx ? cs : 0x7fffffff);
                    if (lane == b) { feed_merge(ch0, ps, mx); ws->any_pend |= 1 << ch0; }
                    fm &= ~wballot(in);
                    wsync();
                }
            }
            wsync();
        }
        ncb = 0;
    }
    /* end of column j: apply the pending chain entries for column j+1 (ancestors in index order, strict >), one lane per chain */
    AUGB_D void apply_pending(int j) {
        const unsigned pend = (unsigned)ws->any_pend;
        if (!pend) return;
        wsync();                              /* every lane has read the flags before lane 0 clears them */
        AUGB_ROLLED
        for (int ch = lane; ch < NCHAIN; ch += AUGB_NLANES) {
            if (!(pend >> ch & 1)) continue;
            sc_t pv = ws->pend_val[ch];
            if (isneg(pv)) continue;
            int pp = ws->pend_pred[ch], self = m->chain_state[ch];
            sc_t cur = ws->tilde[ch];
            bool take = pv > cur || (pv == cur && pp < self);
            int n = ws->cp_n[ch];
            if (take && n >= w.cp_cap) { ws->status = 8; take = false; }
            if (take) { ChainCP c; c.col = j + 1; c.pred = pp; c.tilde = pv; w.cp(ch)[n] = c; ws->cp_n[ch] = n + 1; ws->tilde[ch] = pv; }
            ws->pend_val[ch] = SC_NEG; ws->pend_pred[ch] = 0x7fffffff;
            if (FWD && !ws->pend_f[ch].empty()) {
                /* forward[j+1][chain] = forward[j][chain]*self + sum of the entries: one more change point */
                Lse t = ws->pend_f[ch];
                if (ws->fcp_n[ch] > 0) t.add(ws->ftilde[ch]);
                int nf = ws->fcp_n[ch];
                if (nf >= w.fcp_cap) ws->status = 8;
                else { FChainCP c; c.col = j + 1; c.rstay = STOP_COMPLEX; c.add = 0; c.ft = t.value(); w.fcp(ch)[nf] = c; ws->fcp_n[ch] = nf + 1; ws->ftilde[ch] = c.ft; }
                ws->pend_f[ch].clear();
            }
        }
        wsync();
        if (lane == 0) ws->any_pend = 0;
        wsync();
    }

    /* ------------------------------------------------------------ signal scores (tabulated by the prep pass) */
    AUGB_D sc_t sig(int which, int j) const { return w.sig[(size_t)which * L + j]; }
    AUGB_DN sc_t motif_fwd1(const sc_t* tab, int n, int k, int p) const { return motif_fwd(m, sq, cls, tab, n, k, p); }

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
    /* initial / exon-terminal content of [left, right] (ExonModel::initialSeqProb / eTermSeqProb, exonmodel.cc:1979-2034):
     * difference of the 3-periodic prefix arrays PA_XIN / PA_XET written by the prep pass (exact in fixed point) */
    AUGB_D sc_t exon_shortProb(int pa, int fwd, int left, int right, int frameOfRight) const {
        if (left > right) return 0;
        if (left < 0 || right >= L) return exon_shortProb_loop(pa == PA_XET ? m->xet : m->xinit, fwd, left, right, frameOfRight);
        const sc_t* P = fwd ? parr(cls, pa + mod3(frameOfRight - right)) : parr(cls, pa + 3 + mod3(frameOfRight + right));
        return P[right + 1] - P[left];
    }
    AUGB_DN sc_t exon_shortProb_loop(const sc_t* tab, int fwd, int left, int right, int frameOfRight) const {   /* ranges that leave the window */
        sc_t s = 0;
        AUGB_ROLLED
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
        case E_RSINGLE: case E_RINITIAL: return sig(SG_XRS, end);
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
    AUGB_DN sc_t notEndPart(const StateDesc& st, int bos, int right, int frameOfRight) const {
        const int k = m->k, fwd = st.fwd, win = st.frame;
        sc_t beginPart;
        int bobe = bos - st.innerPartOffset;
        switch (st.ek) {
        case E_SINGLE: case E_INITIAL: {
            if (!(bobe >= 0 && bobe < L - 2)) return SC_NEG;
            int pn = sq.kmer_end(bobe + 2, 3);
            if (pn < 0 || isneg(m->startp[pn])) return SC_NEG;
            beginPart = m->startp[pn];
            int tis = bobe - m->tiw;
            if (tis > m->tis_k) beginPart = tis_bin(m, cls, beginPart + motif_fwd1(m->tis, m->tis_n, m->tis_k, tis));
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
            int pn = fwd ? sq.kmer_end(bos + l, l + 1) : sq.kmer_rc(bos, l + 1);
            if (pn < 0) rest = (sc_t)(l + 1) * m->probN;
            else { int f = fwd ? frameOfRight : mod3(frameOfRight + right - bos); rest = m->xpls[l][(((size_t)cls * 3 + f) << (2 * (l + 1))) | pn]; }
        } else {
            int endOfStart = bos + k - 1, beginOfInitP = right - (k - 1);
            if (k == 0) rest = 0;
            else {
                int pn = fwd ? sq.kmer_end(bos + k - 1, k) : sq.kmer_rc(beginOfInitP, k);
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
                rest += exon_shortProb(PA_XIN, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                      + exon_seqProb(1, endOfInitial + 1, right, frameOfRight);
                break;
            case E_INITIAL:
                endOfInitial = endOfStart + m->init_len;
                if (endOfInitial > right) { endOfInitial = right; beginOfTerm = right + 1; }
                else { beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfInitial) beginOfTerm = right + 1; }
                rest += exon_shortProb(PA_XIN, 1, endOfStart + 1, endOfInitial, mod3(frameOfRight - right + endOfInitial))
                      + exon_seqProb(1, endOfInitial + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                      + exon_shortProb(PA_XET, 1, beginOfTerm, right, frameOfRight);
                break;
            case E_INTERNAL:
                beginOfTerm = right - m->et_len + 1; if (beginOfTerm <= endOfStart) beginOfTerm = right + 1;
                rest += exon_seqProb(1, endOfStart + 1, beginOfTerm - 1, mod3(frameOfRight - right + (beginOfTerm - 1)))
                      + exon_shortProb(PA_XET, 1, beginOfTerm, right, frameOfRight);
                break;
            case E_TERMINAL:
                rest += exon_seqProb(1, endOfStart + 1, right, frameOfRight);
                break;
            case E_RSINGLE:
                beginOfInitial = beginOfInitP - m->init_len; if (beginOfInitial < bos) beginOfInitial = bos;
                rest += exon_shortProb(PA_XIN, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_seqProb(0, bos, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)));
                break;
            case E_RINITIAL:
                beginOfInitial = beginOfInitP - m->init_len;
                if (beginOfInitial < bos) { beginOfInitial = bos; endOfTerm = bos - 1; }
                else { endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitial) endOfTerm = bos - 1; }
                rest += exon_shortProb(PA_XIN, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_seqProb(0, endOfTerm + 1, beginOfInitial - 1, mod3(frameOfRight + right - (beginOfInitial - 1)))
                      + exon_shortProb(PA_XET, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
                break;
            case E_RINTERNAL:
                endOfTerm = bos + m->et_len - 1; if (endOfTerm >= beginOfInitP) endOfTerm = bos - 1;
                rest += exon_seqProb(0, endOfTerm + 1, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)))
                      + exon_shortProb(PA_XET, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm));
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

    /* begin part x initial pattern x initial content of a forward initial / single exon whose start codon is at bobe,
     * when the inner sequence is long enough for all three (exonmodel.cc:1427-1462, 1596-1603, 1611-1633); df = frameOfRight - right */
    AUGB_D sc_t begin_score(int bobe, int df) const {
        /* tabulated by the prep pass (PA_BEG) for the frame every valid candidate has: position p is in frame mod3(p - bobe) */
        if (mod3(df + bobe) == 0 && bobe >= 0 && bobe <= L) return parr(cls, PA_BEG)[bobe];
        return begin_score_calc(bobe, df);
    }
    AUGB_DN sc_t begin_score_calc(int bobe, int df) const {
        const int k = m->k, bos = bobe + 3;
        int pn = sq.kmer_end(bobe + 2, 3);
        if (pn < 0 || isneg(m->startp[pn]) || !(bobe >= 0 && bobe < L - 2)) return SC_NEG;
        sc_t v = m->startp[pn];
        int tis = bobe - m->tiw;
        if (tis > m->tis_k) v = tis_bin(m, cls, v + motif_fwd(m, sq, cls, m->tis, m->tis_n, m->tis_k, tis));
        else v += (sc_t)(bos - 3) * m->log025;
        const int endOfStart = bos + k - 1;
        int p4 = sq.kmer_end(endOfStart, k);
        v += p4 < 0 ? (sc_t)k * m->probN : m->xpls[k - 1][(((size_t)cls * 3 + mod3(df + endOfStart)) << (2 * k)) | p4];
        const int endOfInitial = endOfStart + m->init_len;
        return v + exon_shortProb(PA_XIN, 1, endOfStart + 1, endOfInitial, mod3(df + endOfInitial));
    }

    /* ExonModel::viterbiForwardAndSampling (exonmodel.cc:899-1179) for state s ending at column j.
     *
     * The duration loop of the reference (beginOfStart = startMax .. startMin, exonmodel.cc:1059) becomes one
     * candidate stream, 32 candidates per warp step:
     *   internal / terminal / rinternal / rinitial : the longass_f / rlongdss_f cells of the matching reading-frame phase
     *   initial / single                           : in-frame start codons, found by scanning the ORF in steps of 3
     *   rterminal / rsingle                        : the one position behind the nearest in-frame reverse stop
     * A candidate whose inner sequence is long enough that every sub-model of notEndPartEmiProb applies in full
     * (the overwhelmingly common case) is scored by the closed form
     *     begin-signal + P_ls start pattern + content prefix difference + terminal part + 3*lenDist
     * whose end-dependent terms are computed once per exon end; shorter ones go through the general routine. */
    /* Up to three exon states of one kind (the three reading frames) are evaluated in ONE pass on the device: the warp splits into
     * groups of `gl` lanes (8, or 32 for a single state), every lane carries the state of its group in `s` (-1: no state), so the
     * end-dependent terms, the candidate stream and the scoring below are the same instruction stream for all frames; reductions
     * and the final arg-best run per group.  The host build (one lane) and the forward / sampling kernels pass gl = all lanes. */
    AUGB_D void exon_eval(int s, int j, int gl = AUGB_NLANES) {
        const int g0 = lane & ~(gl - 1), li = lane - g0;                       /* first lane of my group, my index in it */
        const unsigned gmask = (gl >= 32 ? 0xffffffffu : ((1u << gl) - 1u)) << g0;      /* (team-relative, like the ballots) */
        bool alive = s >= 0;
        if (!alive) s = m->xslot[0] >= 0 ? m->xslot[0] : 0;                    /* any valid state: the lanes of an empty group idle along */
        const StateDesc& st = m->st[s]; const int fwd = st.fwd, win = st.frame, ek = st.ek, k = m->k;
        sc_t ep = alive ? endPart(st, j) : SC_NEG;
        const int eobe = j + st.baseOffset, right = eobe - st.innerPartEndOffset;
        if (isneg(ep) || right < 0) alive = false;
        if (wballot(alive) == 0) return;
        const int frameOfRight = fwd ? mod3(win - (eobe + 1) + right) : mod3(win + eobe + 1 - right);
        int eons = (ek == E_TERMINAL || ek == E_SINGLE) ? eobe - 3 : eobe;
        if (eons > L - 1) eons = L - 1;
        const int feons = fwd ? mod3(win - 1 - eobe + eons) : mod3(win + 1 + eobe - eons);
        const int ORFleft = leftmostExonBegin(feons, eons, fwd);
        int startMax = eobe + st.innerPartOffset - m->min_exon_length + 1, startMin;
        if (ek == E_RTERMINAL || ek == E_RSINGLE) startMin = startMax = ORFleft + 2;
        else {
            startMin = ORFleft <= 0 ? 0 : ORFleft + st.innerPartOffset;
            if (startMax > j + st.beginPartLen) startMax = j + st.beginPartLen;
        }
        /* ---- end-dependent terms of the closed forms (uniform over the lanes of a group) ---- */
        const int thr = k + m->init_len + m->et_len + 2;          /* inner length from which every sub-model applies in full */
        const bool listkind = ek == E_INTERNAL || ek == E_TERMINAL || ek == E_RINTERNAL || ek == E_RINITIAL;
        const bool scankind = ek == E_INITIAL || ek == E_SINGLE;
        const sc_t* PXp = fwd ? parr(cls, PA_PX + mod3(frameOfRight - right)) : parr(cls, PA_PXR + mod3(frameOfRight + right));
        const sc_t* ldtab = (ek == E_INTERNAL || ek == E_RINTERNAL) ? m->ld_internal : ek == E_TERMINAL ? m->ld_terminal
                          : (ek == E_SINGLE || ek == E_RSINGLE) ? m->ld_single : ek == E_RTERMINAL ? m->ld_terminal : m->ld_initial;
        const int etmin = k + m->et_len - 2;                      /* the exon-terminal part applies iff right - bos > etmin */
        sc_t endc = 0; int pxhi = right + 1;                      /* closed form = cand terms + PXp[pxhi] - PXp[lo(bos)] + endc */
        if (alive && k >= 1 && right - k >= 0) {
            if (ek == E_INTERNAL || ek == E_INITIAL) {
                if (right - m->et_len + 1 >= 0) { pxhi = right - m->et_len + 1; endc = exon_shortProb(PA_XET, 1, right - m->et_len + 1, right, frameOfRight); }
            } else if (!fwd && ek != E_RTERMINAL && ek != E_RSINGLE) {
                const int beginOfInitP = right - (k - 1);
                int pn = sq.kmer_rc(beginOfInitP, k);
                endc = pn < 0 ? (sc_t)k * m->probN : m->xpls[k - 1][(((size_t)cls * 3 + mod3(frameOfRight + right - beginOfInitP)) << (2 * k)) | pn];
                pxhi = beginOfInitP;
                if (ek == E_RINITIAL && beginOfInitP - m->init_len >= 0) {
                    const int beginOfInitial = beginOfInitP - m->init_len;
                    pxhi = beginOfInitial;
                    endc += exon_shortProb(PA_XIN, 0, beginOfInitial, beginOfInitP - 1, mod3(frameOfRight + right - (beginOfInitP - 1)));
                }
            }
        }
        /* ---- candidate stream ---- */
        int list = 0, ncl = 0, scan_b0 = 0;
        const Cand* cl = nullptr;
        if (listkind) { list = fwd ? CL_LA + mod3(win - eobe - 1) : CL_RD + mod3(win + eobe + 1); cl = w.cl(list); ncl = ws->cl_n[list]; }
        else if (scankind) {
            const int want = mod3(eobe + 1 - (ek == E_SINGLE ? 0 : win));      /* bobe mod 3 such that len % 3 == win */
            scan_b0 = startMax - mod3(startMax - st.innerPartOffset - want);
        }
        const int lo = listkind ? (startMin < 1 ? 1 : startMin) : startMin;
        sc_t best = SC_NEG; int bkey = -0x7fffffff, bpred = -1, bbase = -1;
        Lse fl; fl.clear();
        /* initial / single exons: one pass over the in-frame start codons per ancestor (igenic; with UTR states the two 5' UTR
         * states that end trans_init_window bases before the start codon) */
        const int npass = UTR ? wmaxi(alive && scankind ? (int)st.nanc : 1) : 1;       /* (collective: every lane takes part) */
        AUGB_ROLLED
        for (int ai = 0; ai < npass; ai++) {
        const bool pass_on = alive && ai < (scankind ? (int)st.nanc : 1);
        const int anc0 = st.anc[pass_on ? ai : 0];
        const sc_t t0 = TR(anc0, s);
        int step = 0; bool more = true, gdone = !pass_on;
        AUGB_ROLLED
        while (more) {
            /* lane's candidate: (bos, predecessor a, its cell value pv, its column eop) */
            bool valid = false, below = gdone; int bos = 0, a = anc0, eop = 0; sc_t pv = SC_NEG, t = t0; double pf = 0;
            if (gdone) {
            } else if (listkind) {
                int i = ncl - 1 - step * gl - li;
                if (i >= 0) {
                    Cand c = cl[i]; bos = c.col + 1;
                    if (bos < lo) below = true;
                    else if (bos <= startMax && c.col < j) { valid = true; a = c.state; pv = c.V; eop = c.col; t = TR(a, s); if (FWD) pf = w.clF(list)[i]; }
                } else below = true;
            } else if (scankind) {
                bos = scan_b0 - 3 * (step * gl + li);
                if (bos < startMin) below = true;
                else {
                    int bobe = bos - 3; eop = bos - st.beginPartLen - 1;
                    if (bobe >= 0 && bobe < L - 2 && eop < j && !isneg(t0)) {
                        int pn = sq.kmer_end(bobe + 2, 3);
                        if (pn >= 0 && !isneg(m->startp[pn])) { pv = lookupV(anc0, eop >= 0 ? eop : 0); valid = !isneg(pv); if (FWD && valid) pf = lookupF(anc0, eop >= 0 ? eop : 0); }
                    }
                }
            } else {
                if (step * gl + gl >= st.nanc) below = true;              /* last step of this group */
                if (step * gl + li < st.nanc) {
                    bos = startMin; eop = bos - st.beginPartLen - 1; a = st.anc[step * gl + li]; t = TR(a, s);
                    /* the one candidate can end AT the current column (reverse stop inside the end part): the reference then reads the
                     * cell of this very column, which exists only for states evaluated earlier in the column (lower index; here the
                     * intergenic chain); behind the current column the matrix is still empty */
                    if ((eop < j || (eop == j && a < s && m->st[a].chain >= 0)) && !isneg(t)) { pv = lookupV(a, eop >= 0 ? eop : 0); valid = !isneg(pv); if (FWD && valid) pf = lookupF(a, eop >= 0 ? eop : 0); }
                }
            }
            {   /* a group stops after the step in which one of its lanes ran past the range; the warp goes on while any group does */
                const unsigned bb = wballot(below);
                if (bb & gmask) gdone = true;
                more = wballot(!gdone) != 0;
            }
            step++;
            /* score: nep = notEndPartEmiProb(bos) */
            sc_t nep = SC_NEG; bool general = false;
            if (valid && !isneg(t)) {
                const int bobe = bos - st.innerPartOffset, len = eobe - bobe + 1, pf = m->st[a].frame, il = right - bos;
                if (listkind && win != mod3(fwd ? pf + len : pf - len)) valid = false;
                else if (len < 1 || len >= m->n_ld_exon) valid = false;
                else if (listkind && eop == 0) general = true;      /* predecessor = initial probabilities: no cell has vouched for the splice site */
                else if (ek == E_INTERNAL || ek == E_TERMINAL || ek == E_RINTERNAL) {
                    /* begin part = 1: the splice site was checked by the longass / rlongdss cell (exonmodel.cc:1464-1486, 1518-1535) */
                    sc_t ld = ldtab[len], rest;
                    if (isneg(ld)) valid = false;
                    else {
                        if (il < 0) rest = -(sc_t)(-il - 1) * m->log025;                          /* begin and end part overlap (:1549-1558) */
                        else if (il <= k) {                                                        /* only a P_ls pattern (:1559-1574) */
                            int pn = fwd ? sq.kmer_end(bos + il, il + 1) : sq.kmer_rc(bos, il + 1);
                            int f = fwd ? frameOfRight : mod3(frameOfRight + il);
                            rest = pn < 0 ? (sc_t)(il + 1) * m->probN : m->xpls[il][(((size_t)cls * 3 + f) << (2 * (il + 1))) | pn];
                        } else if (fwd) {
                            int pn = sq.kmer_end(bos + k - 1, k);
                            rest = pn < 0 ? (sc_t)k * m->probN : m->xpls[k - 1][(((size_t)cls * 3 + mod3(frameOfRight - il + k - 1)) << (2 * k)) | pn];
                            if (ek == E_INTERNAL && il > etmin) rest += PXp[pxhi] - PXp[bos + k] + endc;
                            else rest += PXp[right + 1] - PXp[bos + k];
                        } else {
                            rest = endc;
                            if (il > etmin) { const int endOfTerm = bos + m->et_len - 1; rest += exon_shortProb(PA_XET, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm)) + PXp[pxhi] - PXp[endOfTerm + 1]; }
                            else rest += PXp[pxhi] - PXp[bos];
                        }
                        nep = isneg(rest) ? SC_NEG : rest + (m->log3 + ld);
                    }
                } else if (il >= thr && right - thr >= 0 && k >= 1 && (ek == E_RINITIAL || scankind)) {
                    sc_t ld = ldtab[len];
                    const bool lenok = ek == E_SINGLE ? len % 3 == 0 : ek == E_INITIAL ? (len % 3 == win && len > 2) : len > 2;
                    if (isneg(ld) || !lenok) valid = false;
                    else if (ek == E_RINITIAL) {
                        const int endOfTerm = bos + m->et_len - 1;
                        nep = endc + exon_shortProb(PA_XET, 0, bos, endOfTerm, mod3(frameOfRight + right - endOfTerm)) + (PXp[pxhi] - PXp[endOfTerm + 1]) + (m->log3 + ld);
                    } else {
                        /* start codon x TIS motif, initial pattern, initial content (:1427-1462, 1590-1636) */
                        sc_t bs = begin_score(bobe, frameOfRight - right);
                        const int endOfInitial = bos + k - 1 + m->init_len;
                        if (!isneg(bs)) nep = bs + (ek == E_INITIAL ? PXp[pxhi] - PXp[endOfInitial + 1] + endc : PXp[right + 1] - PXp[endOfInitial + 1]) + (m->log3 + ld);
                    }
                } else general = true;
            }
            if (wballot(valid && general)) { if (valid && general) nep = notEndPart(st, bos, right, frameOfRight); }
            if (FWD && opt) push_opt(valid && !isneg(nep), pf + te2d(t + ep + nep), -(bos * 128 + (127 - a)), a, eop);
            if (valid && !isneg(nep)) {
                sc_t sc = pv + (t + ep + nep); int key = bos * 128 + (127 - a);
                if (sc > best || (sc == best && key > bkey)) { best = sc; bkey = key; bpred = a; bbase = eop; }
                if (FWD && !opt) fl.add(pf + te2d(t + ep + nep));
            }
        }
        }
        if (wballot(alive && listkind && startMin == 0)) {
            /* left-truncated exon, bos = 0: the predecessor is read from column 0 = initial probabilities (exonmodel.cc:1067-1068);
             * every lane of the group walks the ancestors, its first lane keeps the result */
            const bool mine = alive && listkind && startMin == 0;
            const int len = eobe + st.innerPartOffset + 1;
            sc_t nep0 = SC_NEG; bool have = false, stop = !mine;
            const int na = wmaxi(mine ? (int)st.nanc : 0);
            AUGB_ROLLED
            for (int i = 0; i < na; i++) {
                const bool in = !stop && i < st.nanc;
                int a = st.anc[in ? i : 0]; sc_t pv = m->init[a], t = TR(a, s);
                const bool ok = in && !(isneg(pv) || isneg(t) || win != mod3(fwd ? m->st[a].frame + len : m->st[a].frame - len));
                if (wballot(ok && !have)) { if (ok && !have) { nep0 = notEndPart(st, 0, right, frameOfRight); have = true; } }
                if (ok && isneg(nep0)) stop = true;
                const bool use = ok && !stop;
                sc_t sc = pv + (t + ep + nep0); int key = 127 - a;
                if (use && li == 0 && (sc > best || (sc == best && key > bkey))) { best = sc; bkey = key; bpred = a; bbase = -1; }
                if (FWD && !opt && use && li == 0) fl.add(sc2d(pv) + te2d(t + ep + nep0));
                if (FWD && opt) push_opt(use && li == 0, sc2d(pv) + te2d(t + ep + nep0), -(127 - a), a, -1);
            }
        }
        if (FWD && opt) return;
        const int wl = gargbest(best, bkey, gl, gmask);       /* winning lane of my group, -1: none */
        if (wballot(wl >= 0) == 0) return;
        double Fv = 0;
        if (FWD) Fv = glse(fl, gl).value();                     /* forward value of my group's state */
        AUGB_ROLLED
        for (int g = 0; g < AUGB_NLANES; g += gl) {
            const int wg = wbcast(wl, g);
            if (wg >= 0) emit(j, wbcast(s, g), wbcast64(best, wg), wbcast(bpred, wg), wbcast(bbase, wg), FWD ? wbcastd(Fv, g) : 0.0);
        }
    }

    /* ------------------------------------------------------------ intron states */
    /* fixed-length states longdss / longass (intronmodel.cc:686-786) for the three frames of one strand */
    AUGB_D void fixed_eval(int kind, int dir, int j) {
        const int fwd = !dir;
        const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
        int eop; sc_t emi;
        if (kind == K_LONGDSS) { eop = j - dssw; emi = sig(fwd ? SG_DSSF : SG_DSSR, j); }
        else { eop = j - assw - m->ass_up; emi = sig(fwd ? SG_ASSF : SG_ASSR, j); }
        if (eop < 0 || isneg(emi)) return;
#if AUGB_GROUP3
        if (!FWD) {
            /* Viterbi kernels: lane = (frame, ancestor); the strict '>' of the reference keeps the first maximum in ancestor order */
            const int f = lane >> 3, i = lane & 7;
            const int s = f < 3 ? (kind == K_LONGDSS ? m->r_longdss[dir][f] : m->r_longass[dir][f]) : -1;
            sc_t pp = SC_NEG; int a = -1;
            if (s >= 0 && i < m->st[s].nanc) {
                a = m->st[s].anc[i];
                const sc_t t = TR(a, s);
                if (!isneg(t)) { const sc_t pv = lookupV(a, eop); if (!isneg(pv)) pp = pv + (t + emi); }
            }
            const int wl = gargbest(pp, -i, 8, 0xffu << (lane & ~7));
            if (wballot(wl >= 0) == 0) return;
            AUGB_ROLLED
            for (int g = 0; g < 24; g += 8) {
                const int wg = wbcast(wl, g);
                if (wg >= 0) emit(j, wbcast(s, g), wbcast64(pp, wg), wbcast(a, wg), eop, 0.0);
            }
            return;
        }
#endif
        AUGB_ROLLED
        for (int f = 0; f < 3; f++) {
            int s = kind == K_LONGDSS ? m->r_longdss[dir][f] : m->r_longass[dir][f];
            if (s < 0 || (only >= 0 && s != only)) continue;
            const StateDesc& st = m->st[s];
            sc_t best = SC_NEG; int bpred = -1; Lse fl; fl.clear();
            AUGB_ROLLED
            for (int i = 0; i < st.nanc; i++) {
                int a = st.anc[i]; sc_t t = TR(a, s); if (isneg(t)) continue;
                sc_t pv = lookupV(a, eop); if (isneg(pv)) continue;
                sc_t pp = pv + (t + emi);
                if (pp > best) { best = pp; bpred = a; }
                if (FWD && !opt) fl.add(lookupF(a, eop) + te2d(t + emi));
                if (FWD && opt) push_opt(lane == 0, lookupF(a, eop) + te2d(t + emi), i, a, eop);
            }
            if (!isneg(best) && !(FWD && opt)) emit(j, s, best, bpred, eop, FWD ? fl.value() : 0.0);
        }
    }
    /* equalD / requalD (intronmodel.cc:695-698, 889-894): fires dStateLen columns after a longdss / rlongass cell */
    AUGB_D void equald_eval(int dir, int f, int j) {
        int s = m->r_equald[dir][f]; int list = (dir ? CL_RA : CL_LD) + f;
        int cur = ws->eq_cur[dir * 3 + f];
        if (FWD && opt) {
            /* sampling step: the single option is the longdss / rlongass cell dStateLen columns back (binary search by column) */
            const Cand* cl = w.cl(list); int lo = 0, hi = ws->cl_n[list] - 1, want = j - m->dStateLen;
            AUGB_ROLLED
            while (lo < hi) { int mid = (lo + hi) >> 1; if (cl[mid].col < want) lo = mid + 1; else hi = mid; }
            if (hi < 0 || cl[lo].col != want || s < 0) return;
            const sc_t* P = parr(cls, PA_PI);
            sc_t t = TR(cl[lo].state, s);
            push_opt(lane == 0 && !isneg(t), w.clF(list)[lo] + te2d(t + (P[j + 1] - P[want + 1])), 0, cl[lo].state, want);
            return;
        }
        Cand c = w.cl(list)[cur]; double cf = FWD ? w.clF(list)[cur] : 0.0;
        wsync();
        if (lane == 0) ws->eq_cur[dir * 3 + f] = cur + 1;
        wsync();
        if (s < 0) return;
        int eop = j - m->dStateLen;
        const sc_t* P = parr(cls, PA_PI);                 /* forward k-mers also for requalD (intronmodel.cc:1046-1108) */
        sc_t emi = P[j + 1] - P[eop + 1];
        sc_t t = TR(c.state, s);
        if (isneg(t)) return;
        emit(j, s, c.V + (t + emi), c.state, eop, cf + te2d(t + emi));
    }
    /* ------------------------------------------------------------ SnippetProbs memo near GC-class boundaries
     *
     * SnippetProbs::getSeqProb (statemodel.cc:312-342) memoises products of intron emissions per (right end, length) and
     * composes longer segments from stored pieces.  IntronModel::updateToLocalGC (intronmodel.cc:495-503) swaps the emission
     * table when the GC class changes but keeps the memo, so a piece keeps the class that was active when it was first
     * computed: around a class boundary a lessD emission is a mix of both tables that depends on the order of calls.
     * Away from boundaries every piece in play has the class of the current column and the value is the plain prefix
     * difference; within [boundary - snip_before, boundary + snip_after) the memo is restated entry for entry (lane 0).
     * Starting from an empty memo snip_before (> dStateLen) columns ahead of the boundary is exact: the list of a key only depends on
     * requests coming from higher keys, all of which are replayed, and everything computed before the boundary has the old
     * class whatever its decomposition ("GC-class boundaries"). */
    AUGB_D SnipEnt* snip_ent(int rc, unsigned id) const { return w.snip_pool + (size_t)rc * w.snip_cap + (id & (unsigned)(w.snip_cap - 1)); }
    AUGB_D bool snip_stale(int rc, unsigned id) const { return id + (unsigned)w.snip_cap <= ws->snip_cnt[rc] + 1u; }
    AUGB_D sc_t snip_elem(int rc, int base, int len) const {          /* getElemSeqProb :283-310 with the current class */
        const sc_t* P = parr(cls, rc ? PA_PIR : PA_PI);
        return P[base + 1] - P[base - len + 1];
    }
    AUGB_D void snip_add(int rc, int base, int len, sc_t p) {          /* addProb :344-370 */
        SnipHead* h = w.snip_head + (size_t)rc * SNIP_RING + (base & (SNIP_RING - 1));
        unsigned id = ++ws->snip_cnt[rc];
        SnipEnt* e = snip_ent(rc, id); e->len = len; e->val = p; e->next = 0;
        if (!h->first) { h->first = h->last = id; return; }
        SnipEnt* la = snip_ent(rc, h->last);
        if (la->len < len) { la->next = id; h->last = id; return; }
        SnipEnt* t = snip_ent(rc, h->first);
        if (t->len > len) { e->next = h->first; h->first = id; return; }
        AUGB_ROLLED
        while (t->next && snip_ent(rc, t->next)->len < len) t = snip_ent(rc, t->next);
        if (t->next && snip_ent(rc, t->next)->len == len) return;       /* "tried to add snippet of same length": dropped */
        e->next = t->next; t->next = id;
    }
    AUGB_DN sc_t snip_get(int rc, int base0, int len0) {               /* getSeqProb :312-342, recursion unrolled onto a stack */
        SnipFrame* st = w.snip_stack; int sp = 0; int base = base0, len = len0; sc_t r = 0;
        AUGB_ROLLED
        for (;;) {
            if (len == 0) { r = 0; break; }
            SnipHead* h = w.snip_head + (size_t)rc * SNIP_RING + (base & (SNIP_RING - 1));
            if (!h->first) { r = snip_elem(rc, base, len); snip_add(rc, base, len, r); break; }
            if (snip_stale(rc, h->first) || snip_stale(rc, h->last)) { ws->status = 8; return SC_NEG; }
            SnipEnt* la = snip_ent(rc, h->last);
            if (la->len < len) {
                if (sp >= SNIP_RING) { ws->status = 8; return SC_NEG; }
                st[sp].base = base; st[sp].len = len; st[sp].add = 1; st[sp].part = la->val; sp++;
                base -= la->len; len -= la->len; continue;
            }
            SnipEnt* t = snip_ent(rc, h->first);                        /* SnippetList::getProb :379-393 */
            int partlen; sc_t pv;
            if (t->len > len) { partlen = 0; pv = 0; }
            else {
                AUGB_ROLLED
                while (t->next && snip_ent(rc, t->next)->len < len) t = snip_ent(rc, t->next);
                if (!t->next || snip_ent(rc, t->next)->len > len) { partlen = t->len; pv = t->val; }
                else { partlen = len; pv = snip_ent(rc, t->next)->val; }
            }
            if (partlen == len) { r = pv; break; }
            if (partlen == 0) { r = snip_elem(rc, base, len); snip_add(rc, base, len, r); break; }
            if (sp >= SNIP_RING) { ws->status = 8; return SC_NEG; }
            st[sp].base = base; st[sp].len = len; st[sp].add = 0; st[sp].part = pv; sp++;
            base -= partlen; len -= partlen;
        }
        AUGB_ROLLED
        while (sp > 0) { sp--; r += st[sp].part; if (st[sp].add) snip_add(rc, st[sp].base, st[sp].len, r); }
        return r;
    }
    /* a column inside an emulated region: forget what the key ring held for this position (and for the whole ring at the
     * first column of a region) */
    AUGB_D void snip_column_begin(int j) {
        const bool first = j == 1 || !(w.mask[j - 1] & MB_SLOW);
        if (first) {
            AUGB_ROLLED
            for (int i = lane; i < 2 * SNIP_RING; i += AUGB_NLANES) { w.snip_head[i].first = 0; w.snip_head[i].last = 0; }
        } else if (lane == 0) {
            AUGB_ROLLED
            for (int rc = 0; rc < 2; rc++) { SnipHead* h = w.snip_head + (size_t)rc * SNIP_RING + (j & (SNIP_RING - 1)); h->first = h->last = 0; }
        }
        wsync();
    }

    /* a sampling step at a memo-emulated column: the content value the forward pass got from the memo for (strand, column, length), if it
     * was not the plain prefix difference `plain` (entries are in column order) */
    AUGB_D sc_t snipx_lookup(int dir, int j, int len, sc_t plain) const {
        const int n = ws->n_sx;
        if (n == 0) return plain;
        const SnipX* x = w.snipx;
        int lo = 0, hi = n;                       /* first entry with col >= j */
        AUGB_ROLLED
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid].col < j) lo = mid + 1; else hi = mid; }
        const int key = (len << 1) | dir;
        AUGB_ROLLED
        for (int i = lo; i < n && x[i].col == j; i++) if (x[i].lendir == key) return x[i].val;
        return plain;
    }
    /* lessD / rlessD (intronmodel.cc:540-629, 924-1000): max over the longdss / rlongass cells of the last dStateLen columns */
    AUGB_D void lessd_eval(int dir, int j) {
        const int fwd = !dir;
        int eob = fwd ? j + m->ass_up + m->ass_start + 2 : j + m->dss_end + 2;
        int lme = j - m->dStateLen; if (lme < 0) lme = 0;
        const sc_t* P = parr(cls, fwd ? PA_PI : PA_PIR);
        /* (a sampling step does not run the memo again: it takes the value the forward pass got from it, snipx_lookup) */
        const bool slow = (w.mask[j] & MB_SLOW) != 0 && !(FWD && opt);
        /* Viterbi kernels away from GC-class boundaries: the three frames in one pass, 8 lanes each (as Sweep::exon_eval) */
#ifndef AUGB_LESSD_GL
#define AUGB_LESSD_GL 8
#endif
        const int gl = (AUGB_GROUP3 && !(FWD && opt) && !slow) ? AUGB_LESSD_GL : AUGB_NLANES;
        const int g0 = lane & ~(gl - 1), li = lane - g0;
        const unsigned gmask = (gl >= 32 ? 0xffffffffu : ((1u << gl) - 1u)) << g0;      /* (team-relative, like the ballots) */
        const bool grp = gl < AUGB_NLANES;            /* three frames in one pass */
        AUGB_ROLLED
        for (int f0 = 0; f0 < 3; f0 += (grp ? 3 : 1)) {
            const int f = grp ? (lane >> 3) : f0;
            int s = f < 3 ? m->r_lessd[dir][f] : -1;
            const bool alive = s >= 0 && !(only >= 0 && s != only);
            if (wballot(alive) == 0) continue;
            if (s < 0) s = 0;
            const int list = (dir ? CL_RA : CL_LD) + (f < 3 ? f : 0);
            const Cand* cl = w.cl(list); const int n = alive ? ws->cl_n[list] : 0;
            bool spl = !(fwd && f == 0) && !(!fwd && f == 2);
            int cod0 = 4, cod1 = 4, cod2 = 4;
            if (spl && eob < L - 2) {
                if (fwd && f == 1) { cod1 = sq.at(eob + 1); cod2 = sq.at(eob + 2); }
                else if (fwd && f == 2) { cod2 = sq.at(eob + 1); }
                else if (!fwd && f == 0) { cod0 = cmpl(sq.at(eob + 1)); }
                else if (!fwd && f == 1) { cod0 = cmpl(sq.at(eob + 2)); cod1 = cmpl(sq.at(eob + 1)); }
            }
            sc_t best = SC_NEG; int bkey = -0x7fffffff, bpred = -1; Lse fl; fl.clear();
            const int nl = slow ? 1 : gl;                     /* emulated columns: lane 0 walks the candidates in the reference's order */
            bool gdone = n <= 0; int step = 0;
            bool more = wballot(!gdone) != 0;
            AUGB_ROLLED
            while (more) {
                const int i = slow ? n - 1 - step : n - 1 - step * gl - li; bool below = false;
                bool okopt = false; double olp = 0; int oe = 0, opred = 0;
                if (!gdone && i >= 0 && (!slow || lane == 0)) {
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
                                /* seqProb is evaluated (and memoised) for every candidate that passes the site tests (:970-972) */
                                sc_t seq = P[j + 1] - P[begin];
                                if (slow) {
                                    const sc_t plain = seq;
                                    seq = snip_get(dir, j, j - begin + 1);
                                    if (FWD && seq != plain) {       /* keep what the memo gave for the sampling steps at this column */
                                        const int k = ws->n_sx;
                                        if (k >= w.sx_cap) ws->status = 8;
                                        else { SnipX x; x.col = j; x.lendir = ((j - begin + 1) << 1) | dir; x.val = seq; w.snipx[k] = x; ws->n_sx = k + 1; }
                                    }
                                } else if (FWD && opt && (w.mask[j] & MB_SLOW)) seq = snipx_lookup(dir, j, j - begin + 1, seq);
                                sc_t sc = c.V + (t + (ld + seq));
                                if (sc > best || (sc == best && e > bkey)) { best = sc; bkey = e; bpred = c.state; }
                                if (FWD && !opt) fl.add(w.clF(list)[i] + te2d(t + (ld + seq)));
                                if (FWD && opt) { okopt = true; olp = w.clF(list)[i] + te2d(t + (ld + seq)); oe = e; opred = c.state; }
                            }
                        }
                    }
                } else if (!gdone && i < 0) below = true;
                if (FWD && opt) push_opt(okopt, olp, -oe, opred, oe);
                /* a frame is finished when one of its lanes ran past the dStateLen range or its list is used up */
                const bool gb = slow ? wbcast((int)below, 0) != 0 : (wballot(below) & gmask) != 0;
                step++;
                if (gb || n - 1 - step * nl < 0) gdone = true;
                more = wballot(!gdone) != 0;
            }
            if (FWD && opt) continue;
            const int wl = gargbest(best, bkey, gl, gmask);
            if (wballot(wl >= 0) == 0) continue;
            double Fv = 0;
            if (FWD) Fv = glse(fl, gl).value();
            AUGB_ROLLED
            for (int g = 0; g < AUGB_NLANES; g += gl) {
                const int wg = wbcast(wl, g);
                if (wg >= 0) emit(j, wbcast(s, g), wbcast64(best, wg), wbcast(bpred, wg), wbcast(bkey, wg), FWD ? wbcastd(Fv, g) : 0.0);
            }
        }
    }

    /* ------------------------------------------------------------ UTR states (UtrModel, utrmodel.cc:796-1064)
     *
     * The 16 UTR exon states share one skeleton: end part (a signal of the column) x max over (predecessor end, duration) of
     * predecessor cell x transition x begin signal x content product x length probability.  The reference walks endOfPred from
     * rightMost down to leftMost (up to 5500 positions, skipping with its EOPList); here the positions worth visiting are
     * exactly the entries of a candidate list:
     *   - predecessor = a self-loop chain (igenic, UTR intron) and a begin signal (TSS, ASS, reverse DSS, reverse polyA):
     *     one entry per signal site, holding the chain's value just before the site (appended when the sweep passes the site);
     *   - predecessor = coding exon cells (3' UTR after terminal / single, reverse 5' UTR after rinitial / rsingle): the cells.
     * Content products are differences of the SegProbs-style cumulative sums written by the prep pass (WinView::useg).
     * UTR intron states are one-base self-loop states: four more lazily evaluated chains. */
    /* site at column j (first base of the begin signal): g = signal score - cumulative content sum at beginOfMiddle - 1 */
    AUGB_D void site_append(int list, int list2, int ch, int j, sc_t sigv, int seg, int seg2, int bom_off) {
        const int cs = m->chain_state[ch];
        if (cs < 0) return;
        const sc_t v = chain_value(ch, j - 1);
        if (isneg(v)) return;
        const double f = FWD ? chain_fvalue(ch, j - 1) : 0.0;
        if (lane == 0) {
            const int bm1 = j + bom_off - 1 > L ? L : j + bom_off - 1;       /* beginOfMiddle - 1, inside the cumulative arrays (see utr_eval) */
            cl_append(list, j - 1, cs, v, f, sigv - usegp(seg)[bm1]);
            if (list2 >= 0) cl_append(list2, j - 1, cs, v, f, sigv - usegp(seg2)[bm1]);
        }
        wsync();
    }
    AUGB_D void utr_begins(int j, unsigned mb) {
        const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
        if (mb & MB_TSSB) site_append(CL_T5, -1, 0, j, w.tssF[j], US_INIT5, 0, m->tuw + m->tss_end);
        if (mb & MB_RTTSB) site_append(CL_TR, -1, 0, j, w.ttsR[j + m->dpc], US_R3, 0, m->boxlen + m->dpc);
        if (mb & MB_ASSB) {
            const int jj = j + assw + m->ass_up - 1;          /* >= L: the pattern ends past the window (ghmm_signal.h, MB_ASSB) */
            const sc_t sv = jj < L ? sig(SG_ASSF, jj) : signal_term(m, sq, w.gc[L - 1], SG_ASSF, jj, w.pmask);      /* (with the softmasking bonus of its intron bases, like the table) */
            site_append(CL_A5, -1, CH_UTR + 0, j, sv, US_5, 0, m->ass_up + assw); site_append(CL_A3, -1, CH_UTR + 1, j, sv, US_3, 0, m->ass_up + assw);
            if (m->nc) site_append(CL_NCA, -1, CH_NC, j, sv, US_NC, 0, m->ass_up + assw);                    /* ncinternal after ncintron */
        }
        if (mb & MB_RDSSB) {
            const sc_t sv = sig(SG_DSSR, j + dssw - 1);
            site_append(CL_R5I, CL_R5N, CH_UTR + 2, j, sv, US_RINIT5, US_R5, dssw); site_append(CL_R3, -1, CH_UTR + 3, j, sv, US_R3, 0, dssw);
            if (m->nc) site_append(CL_NCR, -1, CH_NC + 1, j, sv, US_NC, 0, dssw);                              /* rncinternal after rncintron */
        }
    }
    AUGB_DN void utr_eval(int s, int j) {
        const StateDesc& st = m->st[s]; const UtrDesc& u = m->ud[st.ek];
        const int dssw = m->dss_start + m->dss_end + 2, assw = m->ass_start + m->ass_end + 2;
        const bool last = j == L - 1;
        const int eobe = j + u.eobe_off;
        int boe = j + u.boe_off, rm = j - u.rm_off, lm = j - u.lm_off;
        const sc_t* ld = m->uld[u.ldi]; int nld = m->n_uld[u.ldi];
        sc_t ep;
        switch (u.endkind) {                                      /* UtrModel::endPartEmiProb, utrmodel.cc:1072-1110 */
        case UE_ATG: { ep = 0; if (eobe + 3 <= L - 1) { int c = sq.kmer_end(eobe + 3, 3); if (c < 0 || !m->isstart[c]) ep = SC_NEG; } break; }
        case UE_DSSF: ep = boe < 0 ? SC_NEG : sig(SG_DSSF, j); break;
        case UE_TTSF:
            if (last) {                                           /* right-truncated 3' UTR: no signal, tail of the single-exon length distribution */
                ep = 0; boe = L; rm = u.begsig == BS_NONE ? j - 1 : j - assw - m->ass_up; ld = m->uld[9]; nld = m->n_uld[9];
            } else ep = (boe < 0 || boe + m->boxlen - 1 >= L) ? SC_NEG : w.ttsF[boe];
            break;
        case UE_TSSR: ep = (boe < 0 || boe > L) ? SC_NEG : w.tssR[boe]; break;
        case UE_ASSR: ep = boe < 0 ? SC_NEG : sig(SG_ASSR, j); break;
        default: ep = (j + 3 > L - 1 || !isRCStop(m, sq, j + 1)) ? SC_NEG : 0;
        }
        if (isneg(ep)) return;
        const int eom = boe - 1;                                  /* endOfMiddle */
        const sc_t* cum = w.useg + (size_t)u.seg * (size_t)(L + 1);
        const sc_t cumE = eom >= 0 ? cum[eom > L ? L : eom] : 0;
        const sc_t shortfac = u.shortrule == 1 ? m->log2 : -m->log025;
        const Cand* cl = w.cl(u.list); const int ncl = ws->cl_n[u.list];
        sc_t best = SC_NEG; int bkey = -0x7fffffff, bpred = -1, bbase = -1;
        Lse fl; fl.clear();
        int step = 0; bool more = true;
        AUGB_ROLLED
        while (more) {
            const int i = ncl - 1 - step * AUGB_NLANES - lane; step++;
            bool below = i < 0, valid = false; int a = 0, eop = 0; sc_t pv = SC_NEG, te = SC_NEG; double pf = 0;
            if (!below) {
                const Cand c = cl[i]; eop = c.col;
                if (eop < lm) below = true;
                else if (eop <= rm) {
                    /* c.V = V[eop][a] + g: the begin signal and the content sum up to the middle part are already in (site_append / emit) */
                    a = c.state;
                    const sc_t t = TR(a, s);
                    const int b = eop + 1, bom = b + u.bom_off, bobe = b + u.bobe_off, len = eobe - bobe + 1, mlen = eom - bom + 1;
                    if (!isneg(t) && len >= 0 && len < nld) {
                        const sc_t lp = ld[len];
                        sc_t mid;                                  /* content of the middle part + cum[bom - 1] */
                        if (mlen >= 0) mid = cumE;
                        else mid = cum[bom - 1 > L ? L : bom - 1] + (u.shortrule ? shortfac * (sc_t)(-mlen) : (sc_t)0);     /* (cancels the same term folded into c.V) */
                        if (!isneg(lp)) {
                            te = t + ((mid + lp) + ep); valid = true;
                            if (FWD) { const sc_t g = w.clG(u.list)[i]; pf = w.clF(u.list)[i]; pv = c.V - g; te += g; } else pv = c.V;
                        }
                    }
                }
            }
            more = wballot(below) == 0;
            if (FWD && opt) push_opt(valid, pf + te2d(te), -(eop * 128 + (127 - a)), a, eop);
            if (valid) {
                const sc_t sc = pv + te; const int key = eop * 128 + (127 - a);
                if (sc > best || (sc == best && key > bkey)) { best = sc; bkey = key; bpred = a; bbase = eop; }
                if (FWD && !opt) fl.add(pf + te2d(te));
            }
        }
        if (u.trunc && lm < 0) {
            /* the begin signal lies (partly) before the window: predecessor = column 0 = initial probabilities (utrmodel.cc:932-937,
             * 1184-1190, 1200-1204, 1301-1310, 1376-1380) */
            const int lo = u.begsig == BS_TSSF ? -m->tuw : -m->boxlen - m->dpc;
            if (lm < lo) lm = lo;
            const int top = rm < -1 ? rm : -1;
            AUGB_ROLLED
            for (int e0 = top; e0 >= lm; e0 -= AUGB_NLANES) {
                const int eop = e0 - lane; bool valid = false; sc_t te = SC_NEG;
                const int b = eop + 1, bom = b + u.bom_off, bobe = b + u.bobe_off, mlen = eom - bom + 1; int len = eobe - bobe + 1;
                if (eop >= lm) {
                    sc_t bp; const sc_t* ldt = ld; int nl = nld;
                    if (u.begsig == BS_TSSF) {
                        bp = b == 0 ? w.tssF[0] : m->log025 * (sc_t)(bom - 1);
                        if (b < 0 && b + m->tuw == 0) { ldt = m->uld[8]; nl = m->n_uld[8]; if (st.uk == U_SINGLE) len = eom - b + 1 + m->tiw - m->tuw; }
                    } else {
                        bp = (st.uk == U_TERM || bom > 0) ? m->log025 * (sc_t)(bom - 1) : 0;
                        if (st.uk == U_SINGLE) { ldt = m->uld[9]; nl = m->n_uld[9]; }
                    }
                    if (!isneg(bp) && len >= 0 && len < nl && !isneg(ldt[len])) {
                        sc_t mid;
                        if (mlen >= 0 || u.shortrule == 0) mid = bom > eom ? 0 : cumE - (bom < 1 ? 0 : cum[bom - 1]);
                        else mid = shortfac * (sc_t)(-mlen);
                        te = (bp + mid + ldt[len]) + ep; valid = true;
                    }
                }
                AUGB_ROLLED
                for (int ia = 0; ia < st.nanc; ia++) {
                    const int a = st.anc[ia]; const sc_t pv = m->init[a], t = TR(a, s);
                    const bool ok = valid && !isneg(pv) && !isneg(t);
                    if (FWD && opt) push_opt(ok, sc2d(pv) + te2d(t + te), -(eop * 128 + (127 - a)), a, eop);
                    if (ok) {
                        const sc_t sc = pv + (t + te); const int key = eop * 128 + (127 - a);
                        if (sc > best || (sc == best && key > bkey)) { best = sc; bkey = key; bpred = a; bbase = eop; }
                        if (FWD && !opt) fl.add(sc2d(pv) + te2d(t + te));
                    }
                }
            }
        }
        if (FWD && opt) return;
        const int wl = wargbest(best, bkey);
        if (wl < 0) return;
        double Fv = 0;
        if (FWD) Fv = wlse(fl).value();
        emit(j, s, wbcast64(best, wl), wbcast(bpred, wl), wbcast(bbase, wl), Fv);
    }

    /* ------------------------------------------------------------ one column */
    AUGB_D void process_column(int j, unsigned mb, unsigned eqbits) {
        fill_evstart(j);
        set_class(w.gc[j]);
        if (mb & MB_SLOW) snip_column_begin(j);
        if (UTR && (mb & MB_UTR_BEGINS)) utr_begins(j, mb);
        /* one call site per routine (the bodies are large): loop over the strands / states this column activates */
        AUGB_ROLLED
        { unsigned lm = ((mb & MB_LESSD) ? 1u : 0u) | ((mb & MB_RLESSD) ? 2u : 0u);
          AUGB_ROLLED
          while (lm) { int dir = wffs(lm); lm &= lm - 1; lessd_eval(dir, j); } }
        { unsigned qm = eqbits;
          AUGB_ROLLED
          while (qm) { int q = wffs(qm); qm &= qm - 1; equald_eval(q >= 3, q >= 3 ? q - 3 : q, j); } }
        { unsigned fm = ((mb & MB_LONGDSS) ? 1u : 0u) | ((mb & MB_RLONGDSS) ? 2u : 0u) | ((mb & MB_LONGASS) ? 4u : 0u) | ((mb & MB_RLONGASS) ? 8u : 0u);
          AUGB_ROLLED
          while (fm) { int t = wffs(fm); fm &= fm - 1; fixed_eval(t < 2 ? K_LONGDSS : K_LONGASS, t & 1, j); } }
        /* exon states whose end signal is present: 16 slots (single, terminal, 3 initial, 3 internal, rsingle, rinitial,
         * 3 rinternal, 3 rterminal) selected by the mask bits */
        unsigned slots = ((mb & MB_XSTOP) ? 0x3u : 0u) | ((mb & MB_XDSS) ? 0xfcu : 0u) | ((mb & MB_XRSTART) ? 0x300u : 0u) | ((mb & MB_XRASS) ? 0xfc00u : 0u);
        AUGB_ROLLED
        while (slots) {
            int q = wffs(slots); slots &= slots - 1;
            int xs = m->xslot[q], gl = AUGB_NLANES;
#if AUGB_GROUP3
            /* slots 2-4, 5-7, 10-12, 13-15 are the three frames of one kind (set together by the mask): one pass, 8 lanes per frame */
            if ((q == 2 || q == 5 || q == 10 || q == 13) && (slots & (3u << (q + 1))) == (3u << (q + 1))) {
                slots &= ~(3u << (q + 1));
                const int f = lane >> 3;
                xs = f < 3 ? m->xslot[q + f] : -1; gl = 8;
            } else
#endif
            if (xs < 0) continue;
            exon_eval(xs, j, gl);
        }
        if (UTR && (mb & (MB_UTR_ENDS | MB_LONGDSS | MB_RLONGASS))) {
            /* UTR exon states by end signal: slots 0-15 = utr5single, utr5init, utr5internal, utr5term, utr3single, utr3init,
             * utr3internal, utr3term, then the same for the reverse strand */
            unsigned us = ((mb & MB_U5ATG) ? 0x0009u : 0u) | ((mb & MB_LONGDSS) ? 0x0066u : 0u) | ((mb & MB_UTTS) ? 0x0090u : 0u)
                        | ((mb & MB_URTSS) ? 0x0300u : 0u) | ((mb & MB_RLONGASS) ? 0xcc00u : 0u) | ((mb & MB_URSTOP) ? 0x3000u : 0u);
            if (m->nc) us |= ((mb & MB_LONGDSS) ? 0x10000u : 0u) | ((mb & MB_RLONGASS) ? 0x20000u : 0u);      /* ncinternal, rncinternal */
            AUGB_ROLLED
            while (us) {
                int q = wffs(us); us &= us - 1;
                int xs = m->uslot[q];
                if (xs >= 0) utr_eval(xs, j);
            }
        }
        commit_cells(j);
        apply_pending(j);
    }

    /* ------------------------------------------------------------ whole window */
    AUGB_D void attach() { L = w.L; sq.c = w.code; sq.L = L; sq.kf = w.kf; sq.kr = w.kr; sq.k1 = m->k + 1; }      /* per-thread view of the window */
    /* per-window state + column 0; false: the window is not decodable with this layout (status written) */
    AUGB_D bool init_window() {
        ncb = 0;
        if (lane == 0) {
            ws->n_ev = 0; ws->filled = -1; ws->status = 0; ws->any_pend = 0; ws->snip_cnt[0] = ws->snip_cnt[1] = 0; ws->n_sx = 0;
            AUGB_ROLLED
            for (int i = 0; i < (UTR ? NCL : NCL_BASE); i++) ws->cl_n[i] = 0;
            AUGB_ROLLED
            for (int i = 0; i < 6; i++) ws->eq_cur[i] = 0;
            AUGB_ROLLED
            for (int i = 0; i < NCHAIN; i++) { ws->cp_n[i] = 0; ws->tilde[i] = SC_NEG; ws->pend_val[i] = SC_NEG; ws->pend_pred[i] = 0x7fffffff; ws->fcp_n[i] = 0; ws->ftilde[i] = -1e308; ws->pend_f[i].clear(); }
        }
        wsync();
        set_class(w.gc[0]);
        fill_evstart(0);
        /* column 0 = initial probabilities (NAMGene::setStatesInitialProbs, namgene.cc:144-150) */
        /* a window without a single a/c/g/t is all intergenic: viterbi[j][synch] = viterbi[j-1][synch]/4,
         * every other state erased (namgene.cc:205-226); prep wrote AIG[j] = j*log(1/4) and an empty mask */
        const bool alln = (*w.flags & WF_ALLN) != 0;
        if (*w.flags & WF_NOSLAB) { if (lane == 0) { *w.out_n_ev = 0; *w.out_status = 8; for (int i = 0; i < NCHAIN; i++) w.out_ncp[i] = 0; } wsync(); return false; }
        AUGB_ROLLED
        for (int s = 0; s < m->S; s++) {
            sc_t v = m->init[s]; if (isneg(v)) continue;
            int ch = m->st[s].chain;
            if (alln && ch != 0) continue;
            if (ch >= 0) {
                if (lane == 0) {
                    ChainCP c; c.col = 0; c.pred = -1; c.tilde = v; w.cp(ch)[0] = c; ws->cp_n[ch] = 1; ws->tilde[ch] = v;
                    if (FWD) { FChainCP f; f.col = 0; f.rstay = STOP_COMPLEX; f.add = 0; f.ft = sc2d(v); w.fcp(ch)[0] = f; ws->fcp_n[ch] = 1; ws->ftilde[ch] = f.ft; }
                }
                wsync();
            } else {
                int n = ws->n_ev;
                if (lane == 0) { Event e; e.state = (int16_t)s; e.pred = -1; e.predbase = -1; e.V = v; w.ev[n] = e; ws->n_ev = n + 1; if (FWD) w.evF[n] = sc2d(v); }
                wsync();
                const StateDesc& sd = m->st[s];
                if (lane == 0 && sd.kind == K_LONGDSS && sd.fwd) cl_append(CL_LD + sd.frame, 0, s, v, sc2d(v));
                if (lane == 0 && sd.kind == K_LONGASS && !sd.fwd) cl_append(CL_RA + sd.frame, 0, s, v, sc2d(v));
                /* an exon can follow a splice-site state of column 0 when the exon part of the signal is shorter than two bases (ass_end < 2,
                 * dss_start < 2 on the reverse strand): exonmodel.cc:1464-1486 then skips the site test for beginOfBioExon < 2 */
                if (lane == 0 && sd.kind == K_LONGASS && sd.fwd) cl_append(CL_LA + mod3(sd.frame - (1 - m->ass_end)), 0, s, v, sc2d(v));
                if (lane == 0 && sd.kind == K_LONGDSS && !sd.fwd) cl_append(CL_RD + mod3(sd.frame + 1 - m->dss_start), 0, s, v, sc2d(v));
                if (UTR && lane == 0 && sd.kind == K_EXON && (sd.ek == E_SINGLE || sd.ek == E_TERMINAL)) cl_append(CL_X3, 0, s, v, sc2d(v), -usegp(US_3)[0]);
                if (UTR && lane == 0 && sd.kind == K_EXON && (sd.ek == E_RSINGLE || sd.ek == E_RINITIAL)) { cl_append(CL_XRS, 0, s, v, sc2d(v), -usegp(US_RINIT5)[0]); cl_append(CL_XRT, 0, s, v, sc2d(v), -usegp(US_R5)[0]); }
                if (lane == 0 && !alln) feed_chain(0, s, v, sc2d(v));
                wsync();
            }
        }
        apply_pending(0);
        return true;
    }
    AUGB_D void run() {
        attach(); lane = lane_id();
        if (!init_window()) return;
        constexpr int CH = AUGB_SIMT ? AUGB_NLANES : 32;        /* columns per chunk: one per lane of the team */
        AUGB_ROLLED
        for (int j0 = 1; j0 < L; j0 += CH) {
            /* static activity of the next chunk of columns + equalD columns that fall into them */
            unsigned act = 0; unsigned mymask = 0;
#if AUGB_SIMT
            { int j = j0 + lane; mymask = j < L ? w.mask[j] : 0u; act = wballot(mymask != 0); }
#else
            unsigned maskbuf[32];
            AUGB_ROLLED
            for (int t = 0; t < 32; t++) { int j = j0 + t; maskbuf[t] = j < L ? w.mask[j] : 0u; if (maskbuf[t]) act |= 1u << t; }
#endif
            unsigned eqcol[6]; unsigned eqany = 0;
            AUGB_ROLLED
            for (int q = 0; q < 6; q++) {
                eqcol[q] = 0;
                if (m->r_equald[q / 3][q % 3] < 0) continue;
                int list = (q / 3 ? CL_RA : CL_LD) + q % 3; int cur = ws->eq_cur[q], n = ws->cl_n[list];
                AUGB_ROLLED
                for (int i = cur; i < n; i++) {
                    int due = w.cl(list)[i].col + m->dStateLen;
                    if (due >= j0 + CH) break;
                    if (due >= j0 && due < L) { eqcol[q] |= 1u << (due - j0); act |= 1u << (due - j0); eqany |= 1u << (due - j0); }
                }
            }
            AUGB_ROLLED
            while (act) {
                int t = wffs(act); act &= act - 1;
                int j = j0 + t;
                unsigned eqbits = 0;
                if (eqany & (1u << t)) {
                    AUGB_ROLLED
                    for (int q = 0; q < 6; q++) if (eqcol[q] & (1u << t)) eqbits |= 1u << q;
                }
#if AUGB_SIMT
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
            AUGB_ROLLED
            for (int i = 0; i < NCHAIN; i++) { w.out_ncp[i] = ws->cp_n[i]; w.out_nfcp[i] = FWD ? ws->fcp_n[i] : 0; }
        }
        wsync();
    }

};

typedef SweepT<false, false> Sweep;
typedef SweepT<true, false> SweepFwd;
typedef SweepT<false, true> SweepUtr;
typedef SweepT<true, true> SweepFwdUtr;

}  // namespace augb
