/*
 * ghmm_sample.h — posterior sampling of state paths from the forward values (host + device, one warp per window).
 *
 * Replaces NAMGene::getSampledPath (namgene.cc:367-426) and the doSampling branches of the state models: at every step
 * all (predecessor, predecessor end) options of the current cell are listed with probability forward[e][p] * a * emission,
 * sorted by descending probability (OptionsList::prepareSampling, a stable list::sort, vitmatrix.hh:794-796) and one is
 * drawn with  z = rand()/RAND_MAX * cumprob * 0.99999  (OptionsList::sample, vitmatrix.cc:295-320).  The draws come from the
 * caller-supplied stream of rand() values, consumed in the reference's order: one for the last column, then one per path
 * step — including every single-base step of an intergenic / geometric-intron run, which is why a run of self transitions
 * is skipped in one go by advancing the stream.
 */
#pragma once
#include "ghmm_backtrace.h"
#include "ghmm_sweep.h"

namespace augb {

struct SampleOut {
    int cap;                         /* capacity of the state arrays (all samples of the window together) */
    int32_t *begin, *end; uint8_t *type, *trunc;
    int32_t* count; double* logp;    /* per sample */
    int32_t* status;
    int32_t* rand_used = nullptr;    /* out: draws consumed (cursor at the end) */
};
struct SampleScratch { SampleOpt* opt; int opt_cap; int32_t* sorted /* 1 = option already drawn against */; double* ex /* exp(lp - max) */; int* nopt;
                       OcSlot* oc_slots = nullptr; OcOpt* oc_pool = nullptr; int oc_cap = 0; /* cached option lists (step_pick) */ };

AUGB_HD int trunc_flag_s(int type, int end, int predEnd, int L) { return trunc_flag(type, end, predEnd, L); }

template <class SW>
struct SamplerT {
    SW* sw; SampleScratch sc; const uint32_t* rng; int nrng; int cursor; int lane; int oc_used = 0;

    /* draw one option; returns its index in sc.opt (or -1), adds ln(p/cumprob) to *lp.
     * OptionsList::sample (vitmatrix.cc:295-320) walks the options in stable descending order of probability until the cumulative sum
     * passes z.  The sums are formed in the reference's order (cumprob in insertion order, the walk in sorted order); what is parallel
     * is the exponentials (one per lane and option) and the search for the next option of the sorted order (warp arg-best over the
     * options not taken yet) — the walk almost always ends after a few options, so the list is never sorted as a whole. */
    AUGB_D int pick(double* lp) {
        const int n = *sc.nopt;
        if (n > sc.opt_cap) return -2;
        if (n <= 0 || cursor >= nrng) return -1;
        double mx = -1e308;
        AUGB_ROLLED
        for (int i = lane; i < n; i += AUGB_NLANES) { double v = sc.opt[i].lp; mx = v > mx ? v : mx; }
        mx = wmaxd(mx);
        AUGB_ROLLED
        for (int i = lane; i < n; i += AUGB_NLANES) { sc.ex[i] = exp(sc.opt[i].lp - mx); sc.sorted[i] = 0; }
        wsync();
        /* cumprob accumulates in insertion order (OptionsList::add); insertion order = ascending `ord`, which is how the options sit in
         * the buffer when one lane lists them; with 32 lanes the buffer order is the same because lanes take candidates in the
         * reference's order.  Every lane forms the same sum. */
        double cum = 0;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) cum += sc.ex[i];
        const double z = ((double)rng[cursor] / 2147483647.0) * cum * 0.99999;
        int res = -1, first = -1; double cs = 0;
        AUGB_ROLLED
        for (int r = 0; r < n && res < 0; r++) {
            /* the next option of the stable descending order: highest lp, then lowest ord, then lowest index */
            double bl = -1e308; int bo = 0x7fffffff, bi = 0x7fffffff;
            AUGB_ROLLED
            for (int i = lane; i < n; i += AUGB_NLANES) {
                if (sc.sorted[i]) continue;
                const double v = sc.opt[i].lp; const int o = sc.opt[i].ord;
                if (v > bl || (v == bl && (o < bo || (o == bo && i < bi)))) { bl = v; bo = o; bi = i; }
            }
            const double ml = wmaxd(bl);
            const int mo = wmini(bl == ml && bi != 0x7fffffff ? bo : 0x7fffffff);
            const int sel = wmini(bl == ml && bo == mo ? bi : 0x7fffffff);
            if (r == 0) first = sel;
            cs += sc.ex[sel];
            if (z < cs) res = sel;
            else { if (lane == 0) sc.sorted[sel] = 1; wsync(); }
        }
        if (res < 0) res = first;
        const double add = sc.opt[res].lp - mx - log(cum);
        cursor++;
        *lp += add;
        return res;
    }

    /* OptionsList::sample on a short list held by one lane (the options of a self-loop state at a column that received entries):
     * the same arithmetic as pick() above — cumprob in insertion order, draw against the stable descending order */
    AUGB_D int pick_small(const double* lp, const int* ord, int n, uint32_t r, double* add) const {
        double mx = -1e308;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) mx = lp[i] > mx ? lp[i] : mx;
        double cum = 0;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) cum += exp(lp[i] - mx);
        const double z = ((double)r / 2147483647.0) * cum * 0.99999;
        /* the head of the sorted order is the first option that reaches the maximum (its term is exp(0) = 1): nearly every stop ends here */
        int first = 0;
        AUGB_ROLLED
        for (int i = n - 1; i >= 0; i--) if (lp[i] == mx) first = i;
        int res = -1; double cs = exp(lp[first] - mx);
        if (z < cs) res = first;
        AUGB_ROLLED
        for (int rk = 1; rk < n && res < 0; rk++) {
            /* the option of rank rk */
            int sel = -1;
            AUGB_ROLLED
            for (int i = 0; i < n && sel < 0; i++) {
                int before = 0;
                AUGB_ROLLED
                for (int k = 0; k < n; k++) before += (lp[k] > lp[i] || (lp[k] == lp[i] && (ord[k] < ord[i] || (ord[k] == ord[i] && k < i)))) ? 1 : 0;
                if (before == rk) sel = i;
            }
            cs += exp(lp[sel] - mx);
            if (z < cs) res = sel;
        }
        if (res < 0) res = first;
        *add = lp[res] - mx - log(cum);
        return res;
    }
    /* draw from a cached list: the cumulative sums are formed over the stored order exactly as OptionsList::sample forms them */
    AUGB_D void draw_cached(const OcSlot& sl, double* lp, int* pred, int* eop) {
        const OcOpt* o = sc.oc_pool + sl.off;
        const double z = ((double)rng[cursor] / 2147483647.0) * sl.cum * 0.99999;
        int res = -1; double cs = 0;
        AUGB_ROLLED
        for (int r = 0; r < sl.n && res < 0; r++) { cs += o[r].ex; if (z < cs) res = r; }
        if (res < 0) res = 0;
        *lp += o[res].lpmx - log(sl.cum);
        *pred = o[res].pred; *eop = o[res].eop;
        cursor++;
    }
    /* one step of the walk at a cell of a multi-base state: the options are listed by the state's routine the first time a walk of this
     * window stands here and kept in drawing order; 0 = drawn, -1 = no option, -2 = option buffer too small */
    AUGB_D int step_pick(int state, int base, const StateDesc& sd, double* lp, int* pred, int* eop) {
        SW& S = *sw;
        const long long key = ((long long)state << 32) | (unsigned)base;
        int slot = -1, freeslot = -1;
        if (sc.oc_slots) {
            unsigned h = ((unsigned)base * 2654435761u) ^ ((unsigned)state * 40503u);
            AUGB_ROLLED
            for (int q = 0; q < OC_PROBE && slot < 0; q++) {
                const int i = (int)((h + (unsigned)q) & (OC_SLOTS - 1));
                const long long k = sc.oc_slots[i].key;
                if (k == key) slot = i; else if (k == -1) { freeslot = i; break; }
            }
        }
        if (cursor >= nrng) return -1;
        if (slot >= 0) { draw_cached(sc.oc_slots[slot], lp, pred, eop); return 0; }
        if (lane == 0) *sc.nopt = 0;
        wsync();
        S.only = state;
        const int dir = sd.fwd ? 0 : 1;
        if (sd.kind == K_EXON) S.exon_eval(state, base);
        else if (sd.kind == K_UTR) S.utr_eval(state, base);
        else if (sd.kind == K_LESSD) S.lessd_eval(dir, base);
        else if (sd.kind == K_EQUALD) S.equald_eval(dir, sd.frame, base);
        else S.fixed_eval(sd.kind, dir, base);
        S.only = -1;
        const int n = *sc.nopt;
        if (n > sc.opt_cap) return -2;
        if (n <= 0) return -1;
        if (freeslot < 0 || n > OC_MAXN || oc_used + n > sc.oc_cap) {        /* not cached: draw as before */
            const int k = pick(lp);
            if (k < 0) return k;
            *pred = sc.opt[k].pred; *eop = sc.opt[k].eop;
            return 0;
        }
        /* maximum, exponentials, cumprob in insertion order (as pick()), then the whole stable descending order into the pool */
        double mx = -1e308;
        AUGB_ROLLED
        for (int i = lane; i < n; i += AUGB_NLANES) { double v = sc.opt[i].lp; mx = v > mx ? v : mx; }
        mx = wmaxd(mx);
        AUGB_ROLLED
        for (int i = lane; i < n; i += AUGB_NLANES) sc.ex[i] = exp(sc.opt[i].lp - mx);
        wsync();
        double cum = 0;
        AUGB_ROLLED
        for (int i = 0; i < n; i++) cum += sc.ex[i];
        OcOpt* o = sc.oc_pool + oc_used;
        AUGB_ROLLED
        for (int i = lane; i < n; i += AUGB_NLANES) {
            const double v = sc.opt[i].lp; const int od = sc.opt[i].ord; int r = 0;
            AUGB_ROLLED
            for (int k = 0; k < n; k++) { const double u = sc.opt[k].lp; r += (u > v || (u == v && (sc.opt[k].ord < od || (sc.opt[k].ord == od && k < i)))) ? 1 : 0; }
            OcOpt x; x.ex = sc.ex[i]; x.lpmx = v - mx; x.pred = sc.opt[i].pred; x.eop = sc.opt[i].eop;
            o[r] = x;
        }
        if (lane == 0) { OcSlot sl; sl.key = key; sl.off = oc_used; sl.n = n; sl.mx = mx; sl.cum = cum; sc.oc_slots[freeslot] = sl; }
        wsync();
        oc_used += n;
        draw_cached(sc.oc_slots[freeslot], lp, pred, eop);
        return 0;
    }

    /* The options of a self-loop state at column c (> 0), a column of its chain that received entries: the ancestors at c - 1 in index
     * order (igenicmodel.cc:247-261, intronmodel.cc:786-820).  Everything is local to the calling lane.  Returns the number of options. */
    AUGB_D int chain_options(int state, int c, double* lpv, int* ordv, int* prd, const FChainCP* prev) const {
        const SW& S = *sw; const DevModel* m = S.m;
        const StateDesc& sd = m->st[state];
        const sc_t* T = m->trans + (size_t)S.w.gc[c] * m->S * m->S;
        int n = 0;
        const int elo = S.w.evstart[c - 1], ehi = S.w.evstart[c];          /* the cells of column c - 1 */
        AUGB_ROLLED
        for (int i = 0; i < sd.nanc; i++) {
            const int a = sd.anc[i]; const sc_t t = T[a * m->S + state];
            if (isneg(t)) continue;
            double f = -1e308;
            if (a == state && prev) f = prev->ft + S.te2d(S.chainAv(sd.chain, c - 1));       /* the state itself: the change point before this stop (no search) */
            else if (m->st[a].chain >= 0) f = S.chain_fvalue(m->st[a].chain, c - 1);
            else {
                AUGB_ROLLED
                for (int k = elo; k < ehi; k++) if (S.w.ev[k].state == a) { f = S.w.evF[k]; break; }
            }
            if (f > -1e300) { lpv[n] = f + S.te2d(t); ordv[n] = i; prd[n] = a; n++; }
        }
        return n;
    }
    /* one step of a self-loop state at such a column: draw with the rand() value r.  Returns the chosen predecessor state, -1 if there is
     * no option.  32 lanes take 32 consecutive stops of a run. */
    AUGB_D int chain_stop(int state, int c, uint32_t r, double* add, const FChainCP* prev = nullptr) const {
        double lpv[MAXANC]; int ordv[MAXANC], prd[MAXANC];
        const int n = chain_options(state, c, lpv, ordv, prd, prev);
        if (n == 0) return -1;
        const int k = pick_small(lpv, ordv, n, r, add);
        return prd[k];
    }
    /* After the forward pass, once per window: what every walk would recompute at a stop does not depend on the walk.  If the head of the
     * sorted order is the state itself (its term is exp(0) = 1), OptionsList::sample stays in the state iff
     *     (double)r / 2147483647.0 * cumprob * 0.99999 < 1.0,
     * which is monotone in the rand() value r: the largest r that satisfies it (found with that very expression) and the ln-probability
     * of staying are stored in the change point, and a walk compares its draw with it.  Other stops stay STOP_COMPLEX. */
    AUGB_D void prepare_stops() {
        SW& S = *sw; const DevModel* m = S.m;
        AUGB_ROLLED
        for (int ch = 0; ch < NCHAIN; ch++) {
            const int state = m->chain_state[ch]; const int n = S.ws->fcp_n[ch];
            if (state < 0 || n <= 0) continue;
            FChainCP* cp = S.w.fcp(ch);
            AUGB_ROLLED
            for (int i = lane; i < n; i += AUGB_NLANES) {
                const int c = cp[i].col;
                if (c <= 0) continue;
                double lpv[MAXANC]; int ordv[MAXANC], prd[MAXANC];
                const int k = chain_options(state, c, lpv, ordv, prd, i > 0 ? &cp[i - 1] : nullptr);
                if (k <= 0) continue;
                double mx = -1e308;
                AUGB_ROLLED
                for (int q = 0; q < k; q++) mx = lpv[q] > mx ? lpv[q] : mx;
                int first = 0;
                AUGB_ROLLED
                for (int q = k - 1; q >= 0; q--) if (lpv[q] == mx) first = q;
                if (prd[first] != state) continue;
                double cum = 0;
                AUGB_ROLLED
                for (int q = 0; q < k; q++) cum += exp(lpv[q] - mx);
                /* largest r in [0, 2^31 - 1] with z(r) < 1 */
                double est = 2147483647.0 / (cum * 0.99999);
                long r0 = est >= 2147483647.0 ? 2147483647L : (long)est;
                AUGB_ROLLED
                while (r0 >= 0 && !(((double)r0 / 2147483647.0) * cum * 0.99999 < 1.0)) r0--;
                AUGB_ROLLED
                while (r0 < 2147483647L && (((double)(r0 + 1) / 2147483647.0) * cum * 0.99999 < 1.0)) r0++;
                if (r0 < 0) continue;                   /* (never stays: leave it to the full evaluation) */
                cp[i].rstay = (uint32_t)r0; cp[i].add = lpv[first] - mx - log(cum);
            }
        }
        wsync();
    }

    /* all paths of one window */
    AUGB_D void run(int nsamples, SampleOut out) {
        SW& S = *sw; const DevModel* m = S.m; const int L = S.L;
        lane = lane_id(); cursor = 0;
        S.opt = sc.opt; S.nopt = sc.nopt; S.opt_cap = sc.opt_cap;
        if (nsamples > 1) prepare_stops();
        oc_used = 0;
        if (sc.oc_slots) {
            AUGB_ROLLED
            for (int i = lane; i < OC_SLOTS; i += AUGB_NLANES) sc.oc_slots[i].key = -1;
            wsync();
        }
        int used = 0, status = 0;
        const bool alln = (*S.w.flags & WF_ALLN) != 0;
        AUGB_ROLLED
        for (int it = 0; it < nsamples && !status; it++) {
            double lp = 0; int first = used, bad = 0;
            if (alln) {
                if (used >= out.cap) { status = 8; break; }
                if (lane == 0) { out.begin[used] = 0; out.end[used] = L - 1; out.type[used] = (uint8_t)m->st[m->chain_state[0]].type; out.trunc[used] = 0; }
                used++; lp = (double)L * SW::sc2d(m->log025);
            } else {
                /* last column x termProbs (namgene.cc:385-392) */
                if (lane == 0) *sc.nopt = 0;
                wsync();
                AUGB_ROLLED
                for (int s = 0; s < m->S; s++) {
                    sc_t t = m->term[s];
                    double f = isneg(t) ? -1e308 : S.lookupF(s, L - 1);
                    S.push_opt(lane == 0 && f > -1e300, f + SW::sc2d(t), s, s, L - 1);
                }
                int k = pick(&lp);
                if (k < 0) { bad = 1; if (k == -2) status = 8; }
                int state = bad ? 0 : sc.opt[k].pred, base = L - 1;
                int run_end = -1;      /* right end of the chain run being collected (-1: none) */
                AUGB_ROLLED
                while (base > 0 && !bad) {
                    const StateDesc& sd = m->st[state];
                    S.set_class(S.w.gc[base]);
                    if (lane == 0) *sc.nopt = 0;
                    wsync();
                    int ch = sd.chain, pred, eop;
                    if (ch >= 0) {
                        if (run_end < 0) run_end = base;
                        /* last forward change point at or left of base */
                        const FChainCP* cp = S.w.fcp(ch); int lo = 0, hi = S.ws->fcp_n[ch] - 1;
                        if (hi < 0 || cp[0].col > base) { bad = 1; break; }
                        if (cp[hi].col <= base) lo = hi;
                        AUGB_ROLLED
                        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cp[mid].col <= base) lo = mid; else hi = mid - 1; }
                        /* A run of the chain: between change points every step is the self transition (one draw each, nothing to
                         * choose); at a change point column c the walk draws among the ancestors at c - 1.  The draw of the stop at
                         * column c is number cursor + (base - c) whatever happened before it, so the lanes take consecutive change
                         * points, each evaluates its stop on its own, and the first lane that does not stay in the state ends the run. */
                        pred = state; eop = base;
                        bool left = false;
                        AUGB_ROLLED
                        while (!left) {
                            const int i = lo - lane; const bool valid = i >= 0;
                            const int c = valid ? cp[i].col : -1;
                            int mypred = state; double myadd = 0; bool fail = false;
                            if (valid && c > 0) {
                                const long d = (long)cursor + (base - c);
                                if (d >= nrng) fail = true;
                                else if (cp[i].rstay != STOP_COMPLEX && rng[d] <= cp[i].rstay) myadd = cp[i].add;       /* stays (prepare_stops) */
                                else { mypred = chain_stop(state, c, rng[d], &myadd, i > 0 ? &cp[i - 1] : nullptr); if (mypred < 0) fail = true; }
                            }
                            const bool stopper = !valid || c == 0 || fail || mypred != state;
                            const unsigned sb = wballot(stopper);
                            const int f = sb ? wffs(sb) : AUGB_NLANES;
                            /* the log-probabilities of the steps add up in walk order */
                            AUGB_ROLLED
                            for (int l = 0; l < AUGB_NLANES && l <= f; l++) { const double a = wbcastd(myadd, l); lp += a; }
                            if (f > 0) { const int cp_ = wbcast(c, f - 1); cursor += (base - cp_) + 1; base = cp_ - 1; }       /* lanes < f stayed */
                            if (f == AUGB_NLANES) { lo -= AUGB_NLANES; if (base == 0) left = true; continue; }
                            const int cf = wbcast(c, f); const int vf = wbcast(valid ? 1 : 0, f), ff = wbcast(fail ? 1 : 0, f);
                            left = true;
                            if (!vf || ff) { bad = 1; }
                            else if (cf == 0) { cursor += base; base = 0; }                        /* self steps down to column 1 */
                            else { cursor += (base - cf) + 1; base = cf; pred = wbcast(mypred, f); eop = cf - 1; }
                        }
                        if (bad) break;
                        if (cursor > nrng) { bad = 1; break; }
                        if (pred == state) continue;          /* the run reached column 0 (base == 0 ends the walk) */
                        if (used >= out.cap) { status = 8; break; }
                        if (lane == 0) {
                            out.begin[used] = base; out.end[used] = run_end; out.type[used] = (uint8_t)sd.type;
                            out.trunc[used] = (uint8_t)((trunc_flag_s(sd.type, run_end, run_end - 1, L) & 2) | (trunc_flag_s(sd.type, base, base - 1, L) & 1));
                        }
                        used++; run_end = -1;
                    } else {
                        k = step_pick(state, base, sd, &lp, &pred, &eop);
                        if (k < 0) { bad = 1; if (k == -2) status = 8; break; }
                        if (used >= out.cap) { status = 8; break; }
                        if (lane == 0) { out.begin[used] = eop + 1; out.end[used] = base; out.type[used] = (uint8_t)sd.type; out.trunc[used] = (uint8_t)trunc_flag_s(sd.type, base, eop, L); }
                        used++;
                    }
                    state = pred; base = eop;
                }
                if (!bad && !status && run_end >= 0) {     /* the walk ended inside a chain run that reaches column 1 */
                    if (used >= out.cap) status = 8;
                    else {
                        const StateDesc& sd = m->st[state];
                        if (lane == 0) {
                            out.begin[used] = 1; out.end[used] = run_end; out.type[used] = (uint8_t)sd.type;
                            out.trunc[used] = (uint8_t)((trunc_flag_s(sd.type, run_end, run_end - 1, L) & 2) | (trunc_flag_s(sd.type, 1, 0, L) & 1));
                        }
                        used++;
                    }
                }
            }
            if (bad && !status) status = 7;
            if (status) break;
            /* reverse this sample's states into left-to-right order */
            wsync();
            if (lane == 0) {
                int n = used - first;
                AUGB_ROLLED
                for (int i = 0; i < n / 2; i++) {
                    int a = first + i, b = first + n - 1 - i;
                    int32_t tb = out.begin[a]; out.begin[a] = out.begin[b]; out.begin[b] = tb;
                    int32_t te = out.end[a]; out.end[a] = out.end[b]; out.end[b] = te;
                    uint8_t tt = out.type[a]; out.type[a] = out.type[b]; out.type[b] = tt;
                    uint8_t tr = out.trunc[a]; out.trunc[a] = out.trunc[b]; out.trunc[b] = tr;
                }
                out.count[it] = n; out.logp[it] = lp;
            }
            wsync();
        }
        if (lane == 0) { *out.status = status; if (out.rand_used) *out.rand_used = cursor; }
        S.opt = nullptr; S.nopt = nullptr;
        wsync();
    }
};

typedef SamplerT<SweepFwd> Sampler;
typedef SamplerT<SweepFwdUtr> SamplerUtr;

/* glibc rand() (random_r, TYPE_3 additive feedback generator: r[i] = r[i-3] + r[i-31], output r[i] >> 1; seeded by the
 * minimal-standard LCG and 310 discarded outputs) — the stream an unseeded `augustus` process draws from (seed 1).
 * The generator keeps its 34-word ring and a 64-bit position, so a caller that moves forward through the stream (the drop-in
 * accumulates the draws of all pieces and sequences of a process) only ever generates the values it has not seen yet. */
struct GlibcRand {
    uint32_t ring[34]; uint64_t pos = 0; bool seeded = false;      /* ring[(344 + pos + k) % 34] holds r of the last 34 indices */
    void seed(uint32_t sd) {
        int32_t r[34];
        r[0] = (int32_t)sd;
        for (int i = 1; i < 31; i++) {
            long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            long word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        uint32_t st[344];
        for (int i = 0; i < 31; i++) st[i] = (uint32_t)r[i];
        for (int i = 31; i < 34; i++) st[i] = st[i - 31];
        for (int i = 34; i < 344; i++) st[i] = st[i - 31] + st[i - 3];
        for (int i = 310; i < 344; i++) ring[i % 34] = st[i];
        pos = 0; seeded = true;
    }
    uint32_t next() {
        const uint64_t i = 344 + pos;
        const uint32_t v = ring[(i - 31) % 34] + ring[(i - 3) % 34];
        ring[i % 34] = v; pos++;
        return v >> 1;
    }
    /* position the generator so that the next value is number p of the stream of seed 1 */
    void seek(uint64_t p) {
        if (!seeded || p < pos) seed(1);
        while (pos < p) next();
    }
    void fill(uint32_t* out, size_t n) { for (size_t k = 0; k < n; k++) out[k] = next(); }
};
inline void glibc_rand_stream(uint32_t seed, uint32_t* out, size_t n) { GlibcRand g; g.seed(seed); g.fill(out, n); }

}  // namespace augb
