/*
 * ghmm_kernels.cuh — CUDA kernels (sm_100a): prep pass, warp-per-window sweep, backtrace, result pack.
 */
#pragma once
#include <cuda_runtime.h>
#include <cuda/barrier>

#include "ghmm_backtrace.h"
#include "ghmm_defs.h"
#include "ghmm_prep.h"
#include "ghmm_sample.h"
#include "ghmm_seq.h"
#include "ghmm_sweep.h"

namespace augb {

/* The model lives in constant memory: its scalars become instruction operands (c[bank][offset]) instead of loads, which
 * matters for the sweep's code size.  One symbol per process: the host re-uploads it when another model runs. */
__constant__ DevModel c_model;

constexpr int PREP_BS = 256;
constexpr int PREP_ITEMS = 4;      /* measured: 8 / 16 items per thread and warp-segment scans are slower (24 / 30 / 37 / 36 ms per 1184 windows) */
constexpr int PREP_TILE = PREP_BS * PREP_ITEMS;
constexpr int PREP_DNA_TILE = 8192;      /* bases of a window staged in shared memory per bulk copy (multiple of 16) */

/* exclusive block-wide prefix of one value per thread; returns the prefix, *total = block sum */
template <typename T>
__device__ __forceinline__ T block_excl_sum(T v, T* sm /* PREP_BS/32 + 1 */, T* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T incl = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        T w = lane < PREP_BS / 32 ? sm[lane] : (T)0, wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        if (lane < PREP_BS / 32) sm[lane] = wi - w;
        if (lane == PREP_BS / 32 - 1) sm[PREP_BS / 32] = wi;
    }
    __syncthreads();
    T r = sm[wid] + incl - v;
    *total = sm[PREP_BS / 32];
    __syncthreads();
    return r;
}

/* out[off + i] = init + sum_{t<=i} gen(t), i in [0, n) */
template <typename T, typename Gen>
__device__ __forceinline__ void block_scan_gen(Gen gen, T* out, int n, T init, T* sm) {
    T carry = init;
    for (int t0 = 0; t0 < n; t0 += PREP_TILE) {
        T v[PREP_ITEMS]; int i0 = t0 + threadIdx.x * PREP_ITEMS;
        T run = 0;
        #pragma unroll
        for (int i = 0; i < PREP_ITEMS; i++) { int idx = i0 + i; T x = idx < n ? gen(idx) : (T)0; run += x; v[i] = run; }
        T tot; T ex = block_excl_sum<T>(run, sm, &tot);
        #pragma unroll
        for (int i = 0; i < PREP_ITEMS; i++) { int idx = i0 + i; if (idx < n) out[idx] = carry + ex + v[i]; }
        carry += tot;
    }
}

/* nearest in-frame stop tables: a max-scan over (last stop position per residue class) */
struct Int3 { int a, b, c; };
__device__ __forceinline__ Int3 max3(Int3 x, Int3 y) { Int3 r; r.a = max(x.a, y.a); r.b = max(x.b, y.b); r.c = max(x.c, y.c); return r; }

template <bool RC>
__device__ void block_stop_scan(const DevModel* m, const Seq& s, int32_t* out, int L, Int3* sm3) {
    /* positions 0 .. L-3 get the last stop <= i in the residue class of i; L-2, L-1 are patched (exonmodel.cc:147-153) */
    const int n = L - 2;   /* number of scanned positions (i <= L-3) */
    Int3 carry; carry.a = carry.b = carry.c = -1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int t0 = 0; t0 < n; t0 += PREP_TILE) {
        int i0 = t0 + threadIdx.x * PREP_ITEMS;
        Int3 v[PREP_ITEMS]; Int3 run; run.a = run.b = run.c = -1;
        #pragma unroll
        for (int i = 0; i < PREP_ITEMS; i++) {
            int idx = i0 + i;
            if (idx < n && (RC ? isRCStop(m, s, idx) : isStop(m, s, idx))) { int r = idx % 3; if (r == 0) run.a = idx; else if (r == 1) run.b = idx; else run.c = idx; }
            v[i] = run;
        }
        Int3 incl = run;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Int3 t; t.a = __shfl_up_sync(0xffffffffu, incl.a, o); t.b = __shfl_up_sync(0xffffffffu, incl.b, o); t.c = __shfl_up_sync(0xffffffffu, incl.c, o);
            if (lane >= o) incl = max3(incl, t);
        }
        if (lane == 31) sm3[wid] = incl;
        __syncthreads();
        Int3 pre; pre.a = pre.b = pre.c = -1;
        for (int w2 = 0; w2 < wid; w2++) pre = max3(pre, sm3[w2]);
        Int3 ex; ex.a = __shfl_up_sync(0xffffffffu, incl.a, 1); ex.b = __shfl_up_sync(0xffffffffu, incl.b, 1); ex.c = __shfl_up_sync(0xffffffffu, incl.c, 1);
        if (lane == 0) { ex.a = ex.b = ex.c = -1; }
        ex = max3(max3(ex, pre), carry);
        #pragma unroll
        for (int i = 0; i < PREP_ITEMS; i++) {
            int idx = i0 + i;
            if (idx < n) { Int3 r = max3(ex, v[i]); int rr = idx % 3; out[idx] = rr == 0 ? r.a : rr == 1 ? r.b : r.c; }
        }
        Int3 tot; tot.a = tot.b = tot.c = -1;
        for (int w2 = 0; w2 < PREP_BS / 32; w2++) tot = max3(tot, sm3[w2]);
        carry = max3(carry, tot);
        __syncthreads();
    }
}

/* ------------------------------------------------------------------ prep kernel: one CTA per window */
__global__ void __launch_bounds__(PREP_BS) k_prep(const WinDev* __restrict__ wins, int nwin,
                                                  char* pool, unsigned long long pool_size, unsigned long long* pool_used) {
    const DevModel* m = &c_model;
    __shared__ sc_t sm64[PREP_BS / 32 + 2];
    __shared__ int sm32[PREP_BS / 32 + 2];
    __shared__ Int3 sm3[PREP_BS / 32];
    __shared__ int s_classmask;
    __shared__ int s_anynuc;
    __shared__ int s_nruns;
    __shared__ sc_t* s_slab[MAXC];
    __shared__ int s_noslab;
    __shared__ alignas(128) char s_dna[PREP_DNA_TILE + 32];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ cuda::barrier<cuda::thread_scope_block> s_bar;
    if (threadIdx.x == 0) init(&s_bar, PREP_BS);
    __syncthreads();
    for (int wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
        const WinDev wd = wins[wi];
        const int L = wd.L; char* base = wd.base; const WinLayout& lay = wd.lay;
        uint8_t* code = (uint8_t*)(base + lay.code); uint8_t* gc = (uint8_t*)(base + lay.gc); mask_t* mask = (mask_t*)(base + lay.mask);
        if (threadIdx.x == 0) { s_classmask = 0; s_anynuc = 0; }
        __syncthreads();
        uint16_t* kf = (uint16_t*)(base + lay.kf); uint16_t* kr = (uint16_t*)(base + lay.kr);
        {
            /* The window's DNA is staged tile by tile in shared memory by the TMA engine (cp.async.bulk global -> shared, completion on an
             * mbarrier: UBLKCP / SYNCS in the SASS), with 16 bytes of halo on either side; base codes and the (k+1)-mer codes of both
             * strands come out of one pass over the tile (Seq2Int::operator() / ::rc, geneticcode.hh:166-179: 0x8000 = a base that is not
             * acgt or a position off the window). */
            const int k1 = m->k + 1, Lpad = (L + 15) & ~15;
            bool any = false;
            for (int t0 = 0; t0 < L; t0 += PREP_DNA_TILE) {
                const int lo = t0 >= 16 ? t0 - 16 : 0, hi = min(t0 + PREP_DNA_TILE + 16, Lpad), bytes = hi - lo;
                cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
                if (threadIdx.x == 0) {
                    cuda::device::memcpy_async_tx(s_dna, wd.dna + lo, cuda::aligned_size_t<16>(bytes), s_bar);
                    tok = cuda::device::barrier_arrive_tx(s_bar, 1, bytes);
                } else tok = s_bar.arrive();
                s_bar.wait(std::move(tok));
                const int pend = min(t0 + PREP_DNA_TILE, L);
                for (int p = t0 + threadIdx.x; p < pend; p += PREP_BS) {
                    const uint8_t c = base_code(s_dna[p - lo]);
                    code[p] = c; any |= c < 4;
                    int ef = 0, er = 0; bool okf = p - k1 + 1 >= 0, okr = p + k1 <= L;
                    for (int i = 0; i < k1; i++) {
                        if (okf) { const uint8_t b = base_code(s_dna[p - k1 + 1 + i - lo]); if (b > 3) okf = false; else ef = (ef << 2) | b; }
                        if (okr) { const uint8_t b = base_code(s_dna[p + i - lo]); if (b > 3) okr = false; else er |= (3 - b) << (2 * i); }
                    }
                    kf[p] = okf ? (uint16_t)ef : (uint16_t)0x8000; kr[p] = okr ? (uint16_t)er : (uint16_t)0x8000;
                }
                __syncthreads();                      /* the tile is free for the next copy */
            }
            if (any) s_anynuc = 1;
        }
        __syncthreads();
        int32_t* pmask = (int32_t*)(base + lay.pmask);
        if (lay.softmask) {         /* lower-case input bases = soft-masked (SequenceFeatureCollection::prepare, extrinsicinfo.cc:1696-1724) */
            if (threadIdx.x == 0) pmask[0] = 0;
            block_scan_gen<int>([&](int i) { char ch = wd.dna[i]; return (ch >= 'a' && ch <= 'z') ? 1 : 0; }, pmask + 1, L, 0, sm32);
            __syncthreads();
        }
        const bool anynuc = s_anynuc != 0;
        Seq s; s.c = code; s.L = L;
        s.kf = kf; s.kr = kr; s.k1 = m->k + 1;
        /* ---- GC classes: ContentStairs::computeStairs (motif.cc:543-614) ---- */
        if (wd.gc_in) {
            for (int i = threadIdx.x; i < L; i += PREP_BS) gc[i] = wd.gc_in[i];
        } else {
            int32_t* cnt = (int32_t*)(base + lay.ev);          /* scratch: 4 x (L+1) counts, region unused until the sweep */
            for (int b = 0; b < 4; b++) {
                int32_t* cb = cnt + (size_t)b * (L + 1);
                if (threadIdx.x == 0) cb[0] = 0;
                block_scan_gen<int>([&](int i) { return code[i] == b ? 1 : 0; }, cb + 1, L, 0, sm32);
            }
            __syncthreads();
            int win = m->GCwinsize; if (win > L || win < 1) win = L;
            const int half = win / 2, hi = max(L - (win + 1) / 2, half);
            for (int i = threadIdx.x; i < L; i += PREP_BS) {
                int ic = min(max(i, half), hi);
                int lo = ic - half, up = ic + (win + 1) / 2;      /* window [lo, up) */
                int c4[4];
                for (int b = 0; b < 4; b++) { const int32_t* cb = cnt + (size_t)b * (L + 1); c4[b] = cb[up] - cb[lo]; }
                gc[i] = (uint8_t)nearest_class(m, c4);
            }
            __syncthreads();
            /* anti-flicker smoothing: sequential over the raw runs (warp 0 finds them in order) */
            int32_t* runs = cnt;                               /* reuse scratch: (start, value, curval) triples */
            if (threadIdx.x < 32) {
                int nr = 0;
                for (int i0 = 0; i0 < L; i0 += 32) {
                    int i = i0 + threadIdx.x; bool chg = false; int v = 0;
                    if (i < L) { v = gc[i]; chg = i == 0 || gc[i - 1] != v; }
                    unsigned b = __ballot_sync(0xffffffffu, chg);
                    if (chg) { int k = nr + __popc(b & ((1u << threadIdx.x) - 1)); runs[3 * k] = i; runs[3 * k + 1] = v; runs[3 * k + 2] = v; }
                    nr += __popc(b);
                }
                if (threadIdx.x == 0) s_nruns = nr;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int nr = s_nruns;
                /* at run k (start i): lastStep = start of run k-1; fill run k-1 with val_k if short and value before it equals val_k */
                for (int k = 2; k < nr; k++) {
                    int i = runs[3 * k], lastStep = runs[3 * (k - 1)];
                    if (i - lastStep < 1000 && lastStep > 0 && runs[3 * (k - 2) + 2] == runs[3 * k + 1]) runs[3 * (k - 1) + 2] = runs[3 * k + 1];
                }
            }
            __syncthreads();
            {
                int nr = s_nruns;
                for (int k = 1; k + 1 < nr; k++) {
                    int v = runs[3 * k + 2];
                    if (v != runs[3 * k + 1]) for (int i = runs[3 * k] + threadIdx.x; i < runs[3 * (k + 1)]; i += PREP_BS) gc[i] = (uint8_t)v;
                }
            }
            __syncthreads();
        }
        __syncthreads();
        { int cm = 0; for (int i = threadIdx.x; i < L; i += PREP_BS) cm |= 1 << gc[i]; if (cm) atomicOr(&s_classmask, cm); }
        {
            /* columns near a GC-class boundary (number of boundaries in (j - snip_after, j + snip_before]) get MB_SLOW */
            int32_t* nb = (int32_t*)(base + lay.ev);           /* scratch: inclusive count of boundaries at positions <= i */
            block_scan_gen<int>([&](int i) { return i >= 1 && gc[i] != gc[i - 1] ? 1 : 0; }, nb, L, 0, sm32);
            __syncthreads();
            for (int j = threadIdx.x; j < L; j += PREP_BS) {
                unsigned mb = anynuc ? column_mask(m, s, j) : 0u;
                int hi = min(j + m->snip_before, L - 1), lo = j - m->snip_after;
                if (anynuc && j >= 1 && nb[hi] - (lo >= 0 ? nb[lo] : 0) > 0) mb |= MB_SLOW;
                mask[j] = (mask_t)mb;
            }
            __syncthreads();
        }
        __syncthreads();
        const int cm = s_classmask | (anynuc ? 0 : WF_ALLN);
        /* ---- signal scores of every column ---- */
        {
            sc_t* sg = (sc_t*)(base + lay.sig);
            for (int idx = threadIdx.x; idx < NSIG * L; idx += PREP_BS) { int which = idx / L, j = idx - which * L; sg[idx] = anynuc ? signal_term(m, s, gc[j], which, j, pmask) : SC_NEG; }
        }
        if (lay.utr) {
            /* ---- UTR models: TSS / TTS scores, UTR ends and begin-signal sites in the activity mask, SegProbs cumulative sums,
             * intron emission prefix of the UTR-intron chains (formulas in ghmm_signal.h, same as the sequential host builder) ---- */
            sc_t* tssF = (sc_t*)(base + lay.tssF); sc_t* tssR = (sc_t*)(base + lay.tssR); sc_t* ttsF = (sc_t*)(base + lay.ttsF); sc_t* ttsR = (sc_t*)(base + lay.ttsR);
            __syncthreads();
            for (int i = threadIdx.x; i <= L; i += PREP_BS) {
                int r = i + m->tuw + m->tss_end - 1; int c = gc[r < 0 ? 0 : r < L ? r : L - 1];
                tssF[i] = anynuc ? tss_score(m, s, c, 1, i) : SC_NEG; tssR[i] = anynuc ? tss_score(m, s, c, 0, i) : SC_NEG;
            }
            for (int b = threadIdx.x; b <= L; b += PREP_BS) { int c = gc[b < L ? b : L - 1]; ttsF[b] = anynuc ? tts_score(m, s, c, 1, b) : SC_NEG; ttsR[b] = anynuc ? tts_score(m, s, c, 0, b) : SC_NEG; }
            __syncthreads();
            if (anynuc) {
                const sc_t* sg = (const sc_t*)(base + lay.sig);
                for (int j = threadIdx.x; j < L; j += PREP_BS) mask[j] |= (mask_t)utr_column_mask(m, s, j, sg, tssF, tssR, ttsF, ttsR, gc[L - 1]);
            }
            sc_t* useg = (sc_t*)(base + lay.useg);
            for (int g = 0; g < NUSEG; g++) {
                sc_t* cum = useg + (size_t)g * (size_t)(L + 1);
                if (threadIdx.x == 0) cum[0] = m->log025;
                block_scan_gen<sc_t>([&](int i) { int p = i + 1; return useg_term(m, s, gc[p < L ? p : L - 1], g, p); }, cum + 1, L, m->log025, sm64);
            }
            sc_t* aint = (sc_t*)(base + lay.aint);
            if (threadIdx.x == 0) aint[0] = 0;
            block_scan_gen<sc_t>([&](int i) { return intron_emi1(m, s, gc[i + 1], i + 1) + nep_term(m, pmask, i + 1); }, aint + 1, L - 1, (sc_t)0, sm64);
        }
        /* ---- prefix sums ---- */
        if (threadIdx.x == 0) {
            /* slab of each class present: the first lay.nslab_local in the window, the rest from the batch pool */
            int nloc = 0; s_noslab = 0;
            WinOuts* wo = (WinOuts*)(base + lay.outs);
            for (int c = 0; c < MAXC; c++) {
                sc_t* slab = nullptr;
                if (c < m->C && (cm >> c & 1)) {
                    if (nloc < lay.nslab_local) slab = (sc_t*)(base + lay.parr + (size_t)nloc++ * lay.slab);
                    else {
                        unsigned long long off = atomicAdd(pool_used, (unsigned long long)lay.slab);
                        if (pool && off + lay.slab <= pool_size) slab = (sc_t*)(pool + off); else s_noslab = 1;
                    }
                }
                s_slab[c] = slab; wo->slab[c] = slab;
            }
        }
        __syncthreads();
        for (int c = 0; c < m->C; c++) {
            sc_t* slab = s_slab[c];
            if (!slab) continue;
            for (int which = 0; which < PA_NSCAN; which++) {
                sc_t* P = slab + (size_t)which * (size_t)(L + 1);
                if (threadIdx.x == 0) P[0] = 0;
                block_scan_gen<sc_t>([&](int p) { return parr_term(m, s, c, which, p, pmask); }, P + 1, L, (sc_t)0, sm64);
            }
            { sc_t* B = slab + (size_t)PA_BEG * (size_t)(L + 1); for (int p = threadIdx.x; p <= L; p += PREP_BS) B[p] = begin_term(m, s, c, p); }
        }
        sc_t* aig = (sc_t*)(base + lay.aig); sc_t* ageo = (sc_t*)(base + lay.ageo);
        if (threadIdx.x == 0) { aig[0] = 0; ageo[0] = 0; }
        block_scan_gen<sc_t>([&](int i) { return anynuc ? aig_term(m, s, gc, i + 1, pmask) : m->log025; }, aig + 1, L - 1, (sc_t)0, sm64);
        block_scan_gen<sc_t>([&](int i) { return ageo_term(m, s, gc, i + 1, pmask); }, ageo + 1, L - 1, (sc_t)0, sm64);
        /* ---- ORF tables ---- */
        int32_t* nsf = (int32_t*)(base + lay.nsf); int32_t* nsr = (int32_t*)(base + lay.nsr);
        for (int i = threadIdx.x; i < L + 3; i += PREP_BS) { nsf[i] = 0; nsr[i] = 0; }
        __syncthreads();
        if (L >= 3) { block_stop_scan<false>(m, s, nsf, L, sm3); block_stop_scan<true>(m, s, nsr, L, sm3); }
        __syncthreads();
        if (threadIdx.x == 0 && L > 5) { nsf[L - 2] = nsf[L - 5]; nsf[L - 1] = nsf[L - 4]; nsr[L - 2] = nsr[L - 5]; nsr[L - 1] = nsr[L - 4]; }
        if (threadIdx.x == 0) { WinOuts* o = (WinOuts*)(base + lay.outs); o->pad = cm | (s_noslab ? WF_NOSLAB : 0); }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------ sweep kernel: one warp per window */
#ifndef AUGB_SWEEP_WARPS
#define AUGB_SWEEP_WARPS 4
#endif
#ifndef AUGB_SWEEP_MINB
#define AUGB_SWEEP_MINB 4
#endif
constexpr int SWEEP_WARPS = AUGB_SWEEP_WARPS;
constexpr int SWEEP_TEAMS = SWEEP_WARPS * 32 / AUGB_TEAM;      /* windows in flight per CTA: one per team of AUGB_TEAM lanes */
/* the next window of the queue for the caller's team */
__device__ __forceinline__ int next_window(int* __restrict__ next) {
    int wi = 0;
    if (lane_id() == 0) wi = atomicAdd(next, 1);
    return wbcast(wi, 0);
}
template <class SW>
__device__ __forceinline__ void sweep_body(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next) {
    const DevModel* m = &c_model;
    __shared__ WarpState wstate[SWEEP_TEAMS];
    const int tid = threadIdx.x / AUGB_TEAM;
    for (;;) {
        const int wi = next_window(next);
        if (wi >= nwin) break;
        const WinDev& wd = wins[wi];
        SW sw; sw.m = m; sw.ws = &wstate[tid];
        sw.w = make_view(wd.base, wd.lay, wd.L, 0);
        sw.run();
        wsync();
    }
}
__global__ void __launch_bounds__(SWEEP_WARPS * 32, AUGB_SWEEP_MINB) k_sweep(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next) { sweep_body<Sweep>(wins, nwin, next); }
/* models with UTR states (71 states, four more chains, eight more candidate lists) */
__global__ void __launch_bounds__(SWEEP_WARPS * 32, 3) k_sweep_utr(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next) { sweep_body<SweepUtr>(wins, nwin, next); }

/* ------------------------------------------------------------------ forward sweep + posterior sampling: one warp per window */
template <class SW>
__device__ __forceinline__ void sweep_sample_body(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next, const uint32_t* __restrict__ rng, int nrng) {
    const DevModel* m = &c_model;
    __shared__ WarpState wstate[SWEEP_TEAMS];
    __shared__ int s_nopt[SWEEP_TEAMS];
    const int wid = threadIdx.x / AUGB_TEAM, lane = lane_id();
    for (;;) {
        const int wi = next_window(next);
        if (wi >= nwin) break;
        const WinDev& wd = wins[wi];
        SW sw; sw.m = m; sw.ws = &wstate[wid];
        sw.w = make_view(wd.base, wd.lay, wd.L, 0);
        sw.run();
        wsync();
        WinOuts* outs = (WinOuts*)(wd.base + wd.lay.outs);
        if (wd.lay.nsamp > 0) {
            if (wstate[wid].status) { if (lane == 0) { outs->samp_status = wstate[wid].status; outs->rand_used = 0; } }
            else {
                SamplerT<SW> sp; sp.sw = &sw;
                sp.sc.opt = (SampleOpt*)(wd.base + wd.lay.opt); sp.sc.opt_cap = wd.lay.opt_cap; sp.sc.sorted = (int32_t*)(wd.base + wd.lay.sorted); sp.sc.ex = (double*)(wd.base + wd.lay.optex); sp.sc.nopt = &s_nopt[wid];
                sp.sc.oc_slots = (OcSlot*)(wd.base + wd.lay.oc_slots); sp.sc.oc_pool = (OcOpt*)(wd.base + wd.lay.oc_pool); sp.sc.oc_cap = wd.lay.oc_cap;
                sp.rng = rng; sp.nrng = nrng;
                SampleOut so; so.cap = wd.lay.samp_cap;
                so.begin = (int32_t*)(wd.base + wd.lay.s_begin); so.end = (int32_t*)(wd.base + wd.lay.s_end);
                so.type = (uint8_t*)(wd.base + wd.lay.s_type); so.trunc = (uint8_t*)(wd.base + wd.lay.s_trunc);
                so.count = (int32_t*)(wd.base + wd.lay.s_count); so.logp = (double*)(wd.base + wd.lay.s_logp); so.status = &outs->samp_status; so.rand_used = &outs->rand_used;
                sp.run(wd.lay.nsamp, so);
            }
        }
        wsync();
    }
}

__global__ void __launch_bounds__(SWEEP_WARPS * 32) k_sweep_sample(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next,
                                                                   const uint32_t* __restrict__ rng, int nrng) { sweep_sample_body<SweepFwd>(wins, nwin, next, rng, nrng); }
__global__ void __launch_bounds__(SWEEP_WARPS * 32) k_sweep_sample_utr(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next,
                                                                       const uint32_t* __restrict__ rng, int nrng) { sweep_sample_body<SweepFwdUtr>(wins, nwin, next, rng, nrng); }

/* Gather the sampled paths of all windows: per (window, sample) a header, states contiguous per window.
 * Row next-1 of SURVEY.md §8f starts here: the reference's host loop compares every sampled path's transcripts with the ones it has
 * already seen (Transcript::operator==, namgene.cc:875-904); identical state paths are found on the device instead — a path that
 * repeats an earlier sample of its window is not copied again, its header points to the states of the first occurrence and
 * `first` names it, so the device->host traffic and the host's per-path work shrink by the duplicate factor. */
struct SampHdr { int32_t n, offset; double logp; int32_t first, pad; };
constexpr int PACKS_MAXS = 256;          /* samples per window handled by the de-duplication (more: every path is copied) */
__global__ void k_pack_samples(const WinDev* __restrict__ wins, int nwin, SampHdr* __restrict__ hdr, int32_t* __restrict__ wstatus, int* __restrict__ total,
                               int32_t* __restrict__ obegin, int32_t* __restrict__ oend, uint8_t* __restrict__ otype, uint8_t* __restrict__ otrunc, int ocap) {
    int wi = blockIdx.x;
    if (wi >= nwin) return;
    const WinDev& wd = wins[wi];
    const WinOuts* outs = (const WinOuts*)(wd.base + wd.lay.outs);
    const int ns = wd.lay.nsamp;
    const int32_t* cnt = (const int32_t*)(wd.base + wd.lay.s_count); const double* lp = (const double*)(wd.base + wd.lay.s_logp);
    const int32_t* b = (const int32_t*)(wd.base + wd.lay.s_begin); const int32_t* e = (const int32_t*)(wd.base + wd.lay.s_end);
    const uint8_t* t = (const uint8_t*)(wd.base + wd.lay.s_type); const uint8_t* tr = (const uint8_t*)(wd.base + wd.lay.s_trunc);
    __shared__ int s_src[PACKS_MAXS], s_dst[PACKS_MAXS], s_first[PACKS_MAXS];
    __shared__ unsigned long long s_hash[PACKS_MAXS];
    __shared__ int s_st;
    const int st0 = outs->samp_status;
    const bool dedup = ns <= PACKS_MAXS && !st0;
    if (threadIdx.x == 0) {
        int o = 0;
        if (dedup) for (int k = 0; k < ns; k++) { s_src[k] = o; o += cnt[k]; }
    }
    __syncthreads();
    if (dedup) {
        for (int k = threadIdx.x; k < ns; k += blockDim.x) {          /* FNV-1a over the states of sample k */
            unsigned long long h = 1469598103934665603ull;
            const int o = s_src[k], n = cnt[k];
            for (int i = 0; i < n; i++) {
                const unsigned long long w0 = ((unsigned long long)(unsigned)b[o + i] << 32) | (unsigned)e[o + i], w1 = ((unsigned)t[o + i] << 8) | tr[o + i];
                h = (h ^ w0) * 1099511628211ull; h = (h ^ w1) * 1099511628211ull;
            }
            s_hash[k] = h;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < ns; k += blockDim.x) {          /* the earliest sample with the same states */
            int f = k; const int n = cnt[k], o = s_src[k];
            for (int j = 0; j < k && f == k; j++) {
                if (s_hash[j] != s_hash[k] || cnt[j] != n) continue;
                const int oj = s_src[j]; bool same = true;
                for (int i = 0; i < n && same; i++) same = b[o + i] == b[oj + i] && e[o + i] == e[oj + i] && t[o + i] == t[oj + i] && tr[o + i] == tr[oj + i];
                if (same) f = j;
            }
            s_first[k] = f;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int st = st0, tot = 0;
        if (!st) for (int k = 0; k < ns; k++) if (!dedup || s_first[k] == k) tot += cnt[k];
        int off = atomicAdd(total, tot);
        if (off + tot > ocap) { st = 8; tot = 0; }
        wstatus[2 * wi] = st; wstatus[2 * wi + 1] = outs->rand_used;      /* (status, rand() draws consumed) per window */
        s_st = st;
        int o = off, src = 0;
        for (int k = 0; k < ns; k++) {
            SampHdr h; h.n = st ? 0 : cnt[k]; h.logp = st ? 0.0 : lp[k]; h.pad = 0;
            const bool uniq = !dedup || s_first[k] == k;
            h.first = dedup ? s_first[k] : k;
            if (uniq) { h.offset = o; if (dedup) s_dst[k] = o; o += h.n; } else h.offset = s_dst[s_first[k]];
            hdr[(size_t)wi * ns + k] = h;
            if (!dedup && !st) {                                      /* plain copy of sample k (more samples than the tables hold) */
                for (int i = 0; i < cnt[k]; i++) { obegin[h.offset + i] = b[src + i]; oend[h.offset + i] = e[src + i]; otype[h.offset + i] = t[src + i]; otrunc[h.offset + i] = tr[src + i]; }
            }
            src += cnt[k];
        }
    }
    __syncthreads();
    if (!dedup || s_st) return;
    for (int k = 0; k < ns; k++) {
        if (s_first[k] != k) continue;
        const int o = s_src[k], d = s_dst[k], n = cnt[k];
        for (int i = threadIdx.x; i < n; i += blockDim.x) { obegin[d + i] = b[o + i]; oend[d + i] = e[o + i]; otype[d + i] = t[o + i]; otrunc[d + i] = tr[o + i]; }
    }
}

/* ------------------------------------------------------------------ backtrace: one thread per window */
__global__ void k_backtrace(const WinDev* __restrict__ wins, int nwin) {
    const DevModel* m = &c_model;
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwin) return;
    const WinDev& wd = wins[wi];
    WinView v = make_view(wd.base, wd.lay, wd.L, 0);
    WinOuts* outs = (WinOuts*)(wd.base + wd.lay.outs);
    PathOut po; po.cap = wd.lay.path_cap;
    po.begin = (int32_t*)(wd.base + wd.lay.path_begin); po.end = (int32_t*)(wd.base + wd.lay.path_end);
    po.type = (uint8_t*)(wd.base + wd.lay.path_type); po.trunc = (uint8_t*)(wd.base + wd.lay.path_trunc);
    po.n = &outs->path_n; po.status = &outs->path_status; po.score = &outs->score;
    backtrace_window(m, v, po);
}

/* ------------------------------------------------------------------ pack: gather paths into one contiguous result buffer */
struct PackHdr { int32_t n, status, offset, n_ev; sc_t score; };
__global__ void k_pack(const WinDev* __restrict__ wins, int nwin, PackHdr* __restrict__ hdr, int* __restrict__ total,
                       int32_t* __restrict__ obegin, int32_t* __restrict__ oend, uint8_t* __restrict__ otype, uint8_t* __restrict__ otrunc, int ocap) {
    int wi = blockIdx.x;
    if (wi >= nwin) return;
    const WinDev& wd = wins[wi];
    const WinOuts* outs = (const WinOuts*)(wd.base + wd.lay.outs);
    __shared__ int s_off;
    int n = outs->path_n;
    if (threadIdx.x == 0) {
        int off = atomicAdd(total, n);
        s_off = off;
        PackHdr h; h.n = n; h.status = outs->path_status; h.offset = off; h.n_ev = outs->n_ev; h.score = outs->score;
        if (off + n > ocap) { h.n = 0; h.status = 8; }
        hdr[wi] = h;
    }
    __syncthreads();
    int off = s_off;
    if (off + n > ocap) return;
    const int32_t* b = (const int32_t*)(wd.base + wd.lay.path_begin); const int32_t* e = (const int32_t*)(wd.base + wd.lay.path_end);
    const uint8_t* t = (const uint8_t*)(wd.base + wd.lay.path_type); const uint8_t* tr = (const uint8_t*)(wd.base + wd.lay.path_trunc);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { obegin[off + i] = b[i]; oend[off + i] = e[i]; otype[off + i] = t[i]; otrunc[off + i] = tr[i]; }
}

}  // namespace augb
