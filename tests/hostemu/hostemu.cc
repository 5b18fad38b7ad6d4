/*
 * hostemu.cc — TEST-ONLY build of the sweep / backtrace source for the host (one "lane").
 *
 * The product library (libaugb200.so) runs ghmm_sweep.h on the GPU, one warp per window.  This file
 * compiles the very same headers with g++ so that the kernel's control flow, data structures and
 * fixed-point arithmetic can be checked against the oracle in a container that has no GPU.  It is
 * not linked into, loaded by, or reachable from the product; symbols carry the hostemu_ prefix.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef AUGB_SIMT32
#include "simt32.h"          /* 32 fibers per warp: runs the lane-group paths of the device source (see the header) */
#endif
#include "../../augustus_b200/csrc/ghmm_backtrace.h"
#include "../../augustus_b200/csrc/ghmm_model.h"
#include "../../augustus_b200/csrc/ghmm_prep.h"
#include "../../augustus_b200/csrc/ghmm_sample.h"
#include "../../augustus_b200/csrc/ghmm_sweep.h"

using namespace augb;

/* run the sweep of one window: one lane, or 32 fibers that each own a copy of the per-thread state like the threads of a warp */
template <class SW>
static void run_sweep(SW& proto) {
#if defined(AUGB_SIMT32)
    simt::run([&]() { SW mine = proto; mine.run(); });
    proto.attach();
#else
    proto.run();
#endif
}

struct EmuModel { HostModel hm; };

extern "C" {

void* hostemu_model_create(const void* blob, size_t nbytes, char* errbuf, int errcap) {
    EmuModel* e = new EmuModel();
    int rc = e->hm.build(blob, nbytes);
    if (rc) { snprintf(errbuf, errcap, "%d: %s", rc, e->hm.err.c_str()); delete e; return nullptr; }
    return e;
}
void hostemu_model_destroy(void* p) { delete (EmuModel*)p; }

}  // extern "C"
/* implementation, instantiated for the 47-state and the UTR variant of the sweep */
template <class SW>
static int decode_impl(void* mp, const char* dna, int L, const int32_t* gc_in,
                   int cap, int32_t* pbegin, int32_t* pend, uint8_t* ptype, uint8_t* ptrunc, double* logp, int32_t* status,
                   int evcap, int32_t* ev_col, int32_t* ev_state, int64_t* ev_V, int32_t* n_ev_out,
                   int64_t* chainV /* [L][NCHAIN] or NULL */, uint8_t* gc_out /* [L] or NULL */) {
    EmuModel* e = (EmuModel*)mp; const DevModel* m = &e->hm.dm;
    /* as augb200_decode_batch: default capacities first, the generous layout if a structure overflowed */
    WinLayout lay; std::vector<char> buf; char* base = nullptr; WinView v; WarpState ws; SW sw; WinOuts* outs = nullptr;
    std::vector<char> pool; size_t pool_used = 0;
    for (int pass = 0; pass < 2; pass++) {
        lay = make_layout(L, m->C, pass == 1, false, 0, m->utr != 0, m->softmask != 0);
        buf.assign(lay.total + 64, 0);
        base = buf.data();
        int cm = 0;
        pool.assign((size_t)(m->C) * lay.slab + 64, 0); pool_used = 0;
        prep_window_seq(m, dna, L, gc_in, base, lay, &cm, pass == 0 ? pool.data() : nullptr, pool.size(), &pool_used);
        v = make_view(base, lay, L, cm);
        sw = SW(); sw.m = m; sw.w = v; sw.ws = &ws;
        run_sweep(sw);
        outs = (WinOuts*)(base + lay.outs);
        if (outs->status != AUGB200_ERR_CAPACITY) break;
    }
    PathOut po; po.cap = lay.path_cap;
    po.begin = (int32_t*)(base + lay.path_begin); po.end = (int32_t*)(base + lay.path_end);
    po.type = (uint8_t*)(base + lay.path_type); po.trunc = (uint8_t*)(base + lay.path_trunc);
    po.n = &outs->path_n; po.status = &outs->path_status; po.score = &outs->score;
    backtrace_window(m, v, po);
    *status = outs->path_status;
    int n = outs->path_n;
    if (n > cap) { *status = AUGB200_ERR_CAPACITY; n = 0; }
    for (int i = 0; i < n; i++) { pbegin[i] = po.begin[i]; pend[i] = po.end[i]; ptype[i] = po.type[i]; ptrunc[i] = po.trunc[i]; }
    *logp = ldexp((double)outs->score, -FRAC_BITS);
    if (n_ev_out) {
        int ne = outs->n_ev; *n_ev_out = ne;
        for (int j = 0; j < L; j++) for (int i = v.evstart[j]; i < v.evstart[j + 1] && i < evcap; i++) ev_col[i] = j;
        for (int i = 0; i < ne && i < evcap; i++) { ev_state[i] = v.ev[i].state; ev_V[i] = v.ev[i].V; }
    }
    if (chainV) {
        for (int j = 0; j < L; j++)
            for (int ch = 0; ch < NCHAIN; ch++) chainV[(size_t)j * NCHAIN + ch] = m->chain_state[ch] >= 0 ? sw.chain_value(ch, j) : SC_NEG;
    }
    if (gc_out) memcpy(gc_out, v.gc, L);
    return n;
}
/* forward fill: returns ln forward of every event (aligned with ev_col / ev_state) and of the chains per column */
template <class SW>
static int forward_impl(void* mp, const char* dna, int L, const int32_t* gc_in, int evcap, int32_t* ev_col, int32_t* ev_state, double* ev_F,
                    int32_t* n_ev_out, double* chainF /* [L][NCHAIN] */, int32_t* status) {
    EmuModel* e = (EmuModel*)mp; const DevModel* m = &e->hm.dm;
    WinLayout lay; std::vector<char> buf; char* base = nullptr; WinView v; WarpState ws; SW sw; WinOuts* outs = nullptr;
    std::vector<char> pool; size_t pool_used = 0;
    for (int pass = 0; pass < 2; pass++) {
        lay = make_layout(L, m->C, pass == 1, true, 0, m->utr != 0, m->softmask != 0);
        buf.assign(lay.total + 64, 0);
        base = buf.data();
        int cm = 0;
        pool.assign((size_t)(m->C) * lay.slab + 64, 0); pool_used = 0;
        prep_window_seq(m, dna, L, gc_in, base, lay, &cm, pass == 0 ? pool.data() : nullptr, pool.size(), &pool_used);
        v = make_view(base, lay, L, cm);
        sw = SW(); sw.m = m; sw.w = v; sw.ws = &ws;
        run_sweep(sw);
        outs = (WinOuts*)(base + lay.outs);
        if (outs->status != AUGB200_ERR_CAPACITY) break;
    }
    *status = outs->status;
    int ne = outs->n_ev; *n_ev_out = ne;
    for (int j = 0; j < L; j++) for (int i = v.evstart[j]; i < v.evstart[j + 1] && i < evcap; i++) ev_col[i] = j;
    for (int i = 0; i < ne && i < evcap; i++) { ev_state[i] = v.ev[i].state; ev_F[i] = v.evF[i]; }
    if (chainF) for (int j = 0; j < L; j++) for (int ch = 0; ch < NCHAIN; ch++) chainF[(size_t)j * NCHAIN + ch] = m->chain_state[ch] >= 0 ? sw.chain_fvalue(ch, j) : -1e308;
    return ne;
}
/* forward fill + nsamples sampled paths (condensed, concatenated) */
template <class SW>
static int sample_impl(void* mp, const char* dna, int L, const int32_t* gc_in, int nsamples, int cap,
                   int32_t* sb, int32_t* se, uint8_t* st_, uint8_t* str_, int32_t* scount, double* slogp, int32_t* status,
                   uint64_t rand_pos = 0, int32_t* rand_used = nullptr) {
    EmuModel* e = (EmuModel*)mp; const DevModel* m = &e->hm.dm;
    WinLayout lay; std::vector<char> buf; char* base = nullptr; WinView v; WarpState ws; SW sw; WinOuts* outs = nullptr;
    std::vector<char> pool; size_t pool_used = 0;
    for (int pass = 0; pass < 2; pass++) {
        lay = make_layout(L, m->C, pass == 1, true, 0, m->utr != 0, m->softmask != 0);
        buf.assign(lay.total + 64, 0);
        base = buf.data();
        int cm = 0;
        pool.assign((size_t)(m->C) * lay.slab + 64, 0); pool_used = 0;
        prep_window_seq(m, dna, L, gc_in, base, lay, &cm, pass == 0 ? pool.data() : nullptr, pool.size(), &pool_used);
        v = make_view(base, lay, L, cm);
        sw = SW(); sw.m = m; sw.w = v; sw.ws = &ws;
        run_sweep(sw);
        outs = (WinOuts*)(base + lay.outs);
        if (outs->status != AUGB200_ERR_CAPACITY) break;
    }
    if (outs->status) { *status = outs->status; return -1; }
    size_t nrng = (size_t)nsamples * (L + 2);
    /* the walks start rand_pos values into the stream of seed 1 (augb200_set_rand_position); like the library the twin keeps the
     * generator between calls and only produces the window [rand_pos, rand_pos + nrng) */
    static GlibcRand gen;
    std::vector<uint32_t> rng(nrng);
    gen.seek(rand_pos); gen.fill(rng.data(), nrng);
    std::vector<SampleOpt> opts(L + 4096); std::vector<int32_t> sorted(L + 4096); std::vector<double> optex(L + 4096); int nopt = 0;
    std::vector<OcSlot> ocs(OC_SLOTS); std::vector<OcOpt> ocp(L / 2 + 32768);
    SamplerT<SW> sp; sp.sw = &sw; sp.sc.opt = opts.data(); sp.sc.opt_cap = (int)opts.size(); sp.sc.sorted = sorted.data(); sp.sc.ex = optex.data(); sp.sc.nopt = &nopt; sp.sc.oc_slots = ocs.data(); sp.sc.oc_pool = ocp.data(); sp.sc.oc_cap = (int)ocp.size();
    sp.rng = rng.data(); sp.nrng = (int)nrng;
    SampleOut so; so.rand_used = rand_used; so.cap = cap; so.begin = sb; so.end = se; so.type = st_; so.trunc = str_; so.count = scount; so.logp = slogp; so.status = status;
#ifdef AUGB_SIMT32
    simt::run([&]() { SW mine = sw; mine.attach(); mine.lane = lane_id(); SamplerT<SW> sp2 = sp; sp2.sw = &mine; sp2.run(nsamples, so); });
#else
    sp.run(nsamples, so);
#endif
    return *status ? -1 : nsamples;
}
extern "C" {
static bool is_utr(void* mp) { return ((EmuModel*)mp)->hm.dm.utr != 0; }
/* decode one window; optional dump of all events (col, state, V) and chain values V[j][chain] */
int hostemu_decode(void* mp, const char* dna, int L, const int32_t* gc_in,
                   int cap, int32_t* pbegin, int32_t* pend, uint8_t* ptype, uint8_t* ptrunc, double* logp, int32_t* status,
                   int evcap, int32_t* ev_col, int32_t* ev_state, int64_t* ev_V, int32_t* n_ev_out, int64_t* chainV, uint8_t* gc_out) {
    return is_utr(mp) ? decode_impl<SweepUtr>(mp, dna, L, gc_in, cap, pbegin, pend, ptype, ptrunc, logp, status, evcap, ev_col, ev_state, ev_V, n_ev_out, chainV, gc_out)
                      : decode_impl<Sweep>(mp, dna, L, gc_in, cap, pbegin, pend, ptype, ptrunc, logp, status, evcap, ev_col, ev_state, ev_V, n_ev_out, chainV, gc_out);
}
int hostemu_forward(void* mp, const char* dna, int L, const int32_t* gc_in, int evcap, int32_t* ev_col, int32_t* ev_state, double* ev_F,
                    int32_t* n_ev_out, double* chainF, int32_t* status) {
    return is_utr(mp) ? forward_impl<SweepFwdUtr>(mp, dna, L, gc_in, evcap, ev_col, ev_state, ev_F, n_ev_out, chainF, status)
                      : forward_impl<SweepFwd>(mp, dna, L, gc_in, evcap, ev_col, ev_state, ev_F, n_ev_out, chainF, status);
}
/* glibc rand() stream restatement, for checking against libc */
void hostemu_rand_stream(uint32_t seed, uint32_t* out, int n) { glibc_rand_stream(seed, out, (size_t)n); }
int hostemu_sample(void* mp, const char* dna, int L, const int32_t* gc_in, int nsamples, int cap,
                   int32_t* sb, int32_t* se, uint8_t* st_, uint8_t* str_, int32_t* scount, double* slogp, int32_t* status) {
    return is_utr(mp) ? sample_impl<SweepFwdUtr>(mp, dna, L, gc_in, nsamples, cap, sb, se, st_, str_, scount, slogp, status)
                      : sample_impl<SweepFwd>(mp, dna, L, gc_in, nsamples, cap, sb, se, st_, str_, scount, slogp, status);
}
int hostemu_sample_at(void* mp, const char* dna, int L, const int32_t* gc_in, int nsamples, int cap,
                      int32_t* sb, int32_t* se, uint8_t* st_, uint8_t* str_, int32_t* scount, double* slogp, int32_t* status,
                      uint64_t rand_pos, int32_t* rand_used) {
    return is_utr(mp) ? sample_impl<SweepFwdUtr>(mp, dna, L, gc_in, nsamples, cap, sb, se, st_, str_, scount, slogp, status, rand_pos, rand_used)
                      : sample_impl<SweepFwd>(mp, dna, L, gc_in, nsamples, cap, sb, se, st_, str_, scount, slogp, status, rand_pos, rand_used);
}
int hostemu_nchain(void) { return NCHAIN; }
int hostemu_statecount(void* mp) { return ((EmuModel*)mp)->hm.dm.S; }
int hostemu_chain_state(void* mp, int ch) { return ((EmuModel*)mp)->hm.dm.chain_state[ch]; }

}  // extern "C"

/* the carried generator (GlibcRand): values [pos, pos + n) of the stream of seed 1 after an arbitrary history of seeks */
extern "C" void hostemu_rand_window(const uint64_t* seeks, int nseeks, uint64_t pos, uint32_t* out, int n) {
    GlibcRand g;
    for (int i = 0; i < nseeks; i++) { g.seek(seeks[i]); (void)g.next(); }
    g.seek(pos); g.fill(out, (size_t)n);
}
