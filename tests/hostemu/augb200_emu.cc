/*
 * augb200_emu.cc — TEST-ONLY stand-in for the part of the C ABI (include/augb200.h) that the drop-in shim (host/augshim.cc) calls,
 * on top of the host build of the kernel source (hostemu.cc, one lane).
 *
 * Purpose: the shim's own logic — model variants per initial / terminal vector, pieces of sequences longer than maxDNAPieceSize,
 * the cut search, lazy sampling and the rand() stream position, softmask case rebuild — can be checked in a container without a GPU:
 * oracle/Makefile links the reference front end + the shim against THIS file into oracle/_ref/augustus_emu, and
 * tests/test_dropin_emu.py compares its GFF with the unmodified reference binary.  It is not the product and not a fallback:
 * libaugb200.so contains none of it, and oracle/_ref/augustus_b200 (the drop-in proper) links libaugb200.so only.
 */
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/augb200.h"

extern "C" {
void* hostemu_model_create(const void* blob, size_t nbytes, char* errbuf, int errcap);
void hostemu_model_destroy(void* p);
int hostemu_decode(void* mp, const char* dna, int L, const int32_t* gc_in, int cap, int32_t* pbegin, int32_t* pend, uint8_t* ptype,
                   uint8_t* ptrunc, double* logp, int32_t* status, int evcap, int32_t* ev_col, int32_t* ev_state, int64_t* ev_V,
                   int32_t* n_ev_out, int64_t* chainV, uint8_t* gc_out);
int hostemu_sample_at(void* mp, const char* dna, int L, const int32_t* gc_in, int nsamples, int cap, int32_t* sb, int32_t* se,
                      uint8_t* st_, uint8_t* str_, int32_t* scount, double* slogp, int32_t* status, uint64_t rand_pos, int32_t* rand_used);
int hostemu_statecount(void* mp);
}

struct PathStore { std::vector<int32_t> b, e; std::vector<uint8_t> t, tr; };

struct augb200_model {
    void* emu = nullptr;
    PathStore vit, smp;
    uint64_t rand_pos = 0;
    int64_t rand_used0 = 0;
};

static char g_err[512] = "";

extern "C" {

int augb200_model_create(const void* blob, size_t nbytes, int, augb200_model** out) {
    if (!blob || !out) return AUGB200_ERR_BAD_ARG;
    void* e = hostemu_model_create(blob, nbytes, g_err, sizeof g_err);
    if (!e) { int rc = atoi(g_err); return rc ? rc : AUGB200_ERR_BAD_BLOB; }
    augb200_model* m = new augb200_model(); m->emu = e; *out = m;
    return AUGB200_OK;
}
void augb200_model_destroy(augb200_model* m) { if (m) { hostemu_model_destroy(m->emu); delete m; } }
int augb200_model_statecount(const augb200_model* m) { return hostemu_statecount(m->emu); }
int augb200_model_num_gc_classes(const augb200_model*) { return 0; }

static int viterbi(augb200_model* m, const augb200_window* w, augb200_path* out) {
    const int L = w->length, cap = L + 16;
    PathStore& s = m->vit; s.b.resize(cap); s.e.resize(cap); s.t.resize(cap); s.tr.resize(cap);
    double lp = 0; int32_t st = 0;
    int n = hostemu_decode(m->emu, w->dna, L, w->gc_class, cap, s.b.data(), s.e.data(), s.t.data(), s.tr.data(), &lp, &st,
                           0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    out->n = st ? 0 : n; out->status = st; out->begin = s.b.data(); out->end = s.e.data(); out->type = s.t.data(); out->truncated = s.tr.data();
    out->log_prob = lp;
    return AUGB200_OK;
}

int augb200_decode(augb200_model* m, const augb200_window* w, augb200_path* out) {
    if (!m || !w || !out || w->length < 2) return AUGB200_ERR_BAD_ARG;
    return viterbi(m, w, out);
}

int augb200_decode_batch_sampling(augb200_model* m, int32_t n, const augb200_window* w, int32_t nsample, augb200_path* out, augb200_path* samples) {
    if (!m || n != 1 || nsample < 2 || !w || !out || !samples) return AUGB200_ERR_BAD_ARG;     /* the shim decodes one window per call */
    int rc = viterbi(m, w, out);
    if (rc || out->status) return rc;
    const int L = w->length, ns = nsample - 1, cap = ns * (L / 8 + 64);
    PathStore& s = m->smp; s.b.resize(cap); s.e.resize(cap); s.t.resize(cap); s.tr.resize(cap);
    std::vector<int32_t> count(ns); std::vector<double> logp(ns); int32_t st = 0, used = 0;
    hostemu_sample_at(m->emu, w->dna, L, w->gc_class, ns, cap, s.b.data(), s.e.data(), s.t.data(), s.tr.data(), count.data(), logp.data(), &st,
                      m->rand_pos, &used);
    m->rand_used0 = used;
    size_t off = 0;
    for (int k = 0; k < ns; k++) {
        samples[k].n = st ? 0 : count[k]; samples[k].status = st; samples[k].log_prob = logp[k];
        samples[k].begin = s.b.data() + off; samples[k].end = s.e.data() + off; samples[k].type = s.t.data() + off; samples[k].truncated = s.tr.data() + off;
        off += st ? 0 : count[k];
    }
    return AUGB200_OK;
}

int augb200_set_rand_position(augb200_model* m, uint64_t p) { if (!m) return AUGB200_ERR_BAD_ARG; m->rand_pos = p; return AUGB200_OK; }
int64_t augb200_last_rand_consumed(const augb200_model* m) { return m ? m->rand_used0 : 0; }
const char* augb200_strerror(int code) { static char b[32]; snprintf(b, sizeof b, "error %d", code); return b; }
const char* augb200_last_cuda_error(void) { return g_err; }

}  // extern "C"
