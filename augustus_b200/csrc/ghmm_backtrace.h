/*
 * ghmm_backtrace.h — path recovery from the event log and the chain change points (host + device).
 *
 * Replaces NAMGene::getViterbiPath (namgene.cc:432-510).  The reference re-runs each state's DP routine
 * in doBacktracking mode to re-derive (predecessor state, predecessor end); here the sweep stored both
 * with every non-zero cell, and a self-loop chain is walked back to its last change point in one step.
 * Output is the condensed path (StatePath::condenseStatePath, gene.cc:977-1000): runs of identical
 * non-exon states merged, truncation flags OR-ed (State::setTruncFlag, gene.cc:309-321).
 */
#pragma once
#include "ghmm_defs.h"

namespace augb {

struct PathOut {
    int cap;
    int32_t* begin; int32_t* end; uint8_t* type; uint8_t* trunc;   /* filled right-to-left, then reversed */
    int32_t* n; int32_t* status; sc_t* score;
};

AUGB_HD int trunc_flag(int type, int end, int predEnd, int L) {
    const bool utrIntron = type == 26 || type == 27 || type == 32 || type == 33 || type == 61 || type == 62 || type == 67 || type == 68;
    const bool isIntron = (type >= T_LESSD0 && type <= 23) || (type >= T_RLESSD0 && type <= 58) || utrIntron;
    const bool utrExon = ((type >= T_UTR5SINGLE && type <= T_UTR3TERM) || (type >= T_RUTR5SINGLE && type <= T_RUTR3TERM)) && !utrIntron;
    int t = 0;
    if (end == L - 1 && ((type >= 2 && type <= 7) || (type >= 38 && type <= 43) || isIntron || type == T_UTR3SINGLE || type == T_UTR3TERM)) t |= 2;
    if ((predEnd == -1 || predEnd == 0) && ((type >= 5 && type <= 8) || (type >= 37 && type <= 40) || isIntron || utrExon)) t |= 1;
    return t;
}

/* single thread per window */
AUGB_HD void backtrace_window(const DevModel* m, const WinView& w, PathOut o) {
    const int L = w.L, S = m->S;
    int ncp[NCHAIN];
    for (int i = 0; i < NCHAIN; i++) ncp[i] = w.out_ncp[i];
    if (*w.out_status) { *o.status = *w.out_status; *o.n = 0; return; }
    /* last column x termProbs, states in index order, strict > (namgene.cc:443-454) */
    sc_t best = SC_NEG; int state = -1;
    for (int s = 0; s < S; s++) {
        sc_t t = m->term[s]; if (isneg(t)) continue;
        sc_t v = SC_NEG; int ch = m->st[s].chain;
        if (ch >= 0) { if (ncp[ch] > 0) v = w.cp(ch)[ncp[ch] - 1].tilde + (ch == 0 ? w.AIG[L - 1] : ch < CH_UTR ? w.AGEO[L - 1] : w.AINT[L - 1] + (sc_t)(L - 1) * (ch < CH_NC ? m->utr_tself : m->nc_tself)); }
        else { for (int i = w.evstart[L - 1]; i < w.evstart[L]; i++) if (w.ev[i].state == s) v = w.ev[i].V; }
        if (isneg(v)) continue;
        v += t;
        if (v > best) { best = v; state = s; }
    }
    if (state < 0) { *o.status = 6; *o.n = 0; return; }
    *o.score = best;
    int base = L - 1, n = 0, guard = 0;
    while (base > 0) {
        if (++guard > 2 * L + 16) { *o.status = 7; *o.n = 0; return; }
        if (n >= o.cap) { *o.status = 8; *o.n = 0; return; }
        const StateDesc& sd = m->st[state];
        int ch = sd.chain;
        if (ch >= 0) {
            const ChainCP* cp = w.cp(ch); int lo = 0, hi = ncp[ch] - 1;
            if (hi < 0 || cp[0].col > base) { *o.status = 7; *o.n = 0; return; }
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cp[mid].col <= base) lo = mid; else hi = mid - 1; }
            int c = cp[lo].col, b = c == 0 ? 1 : c;
            int tf = 0;
            /* per-base states of the run: the rightmost one decides TRUNC_RIGHT, the leftmost (predEnd = b-1) TRUNC_LEFT */
            tf |= trunc_flag(sd.type, base, base - 1, L) & 2;
            tf |= trunc_flag(sd.type, b, b - 1, L) & 1;
            o.begin[n] = b; o.end[n] = base; o.type[n] = (uint8_t)sd.type; o.trunc[n] = (uint8_t)tf; n++;
            if (c == 0) break;
            state = cp[lo].pred; base = c - 1;
        } else {
            int lo = w.evstart[base], hi = w.evstart[base + 1], k = -1;
            for (int i = lo; i < hi; i++) if (w.ev[i].state == state) { k = i; break; }
            if (k < 0) { *o.status = 7; *o.n = 0; return; }
            int pb = w.ev[k].predbase, ps = w.ev[k].pred;
            if (ps < 0 || (pb >= base && ps == state) || pb > base + 10) { *o.status = 7; *o.n = 0; return; }
            o.begin[n] = pb + 1; o.end[n] = base; o.type[n] = (uint8_t)sd.type; o.trunc[n] = (uint8_t)trunc_flag(sd.type, base, pb, L); n++;
            state = ps; base = pb;
        }
    }
    /* merge adjacent identical non-coding-exon states (cannot happen for chains, kept for safety) and reverse */
    for (int i = 0; i < n / 2; i++) {
        int32_t tb = o.begin[i]; o.begin[i] = o.begin[n - 1 - i]; o.begin[n - 1 - i] = tb;
        int32_t te = o.end[i]; o.end[i] = o.end[n - 1 - i]; o.end[n - 1 - i] = te;
        uint8_t tt = o.type[i]; o.type[i] = o.type[n - 1 - i]; o.type[n - 1 - i] = tt;
        uint8_t tr = o.trunc[i]; o.trunc[i] = o.trunc[n - 1 - i]; o.trunc[n - 1 - i] = tr;
    }
    *o.n = n; *o.status = 0;
}

}  // namespace augb
