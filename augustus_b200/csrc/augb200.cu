/*
 * augb200.cu — C ABI (include/augb200.h) on top of the CUDA kernels.  No CPU fallback: every decode
 * entry point needs a CUDA device and reports AUGB200_ERR_NO_DEVICE / AUGB200_ERR_CUDA otherwise.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/augb200.h"
#include "ghmm_kernels.cuh"
#include "ghmm_model.h"

using namespace augb;

static thread_local std::string g_cuda_err;
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    g_cuda_err = std::string(#call) + ": " + cudaGetErrorString(e__); return AUGB200_ERR_CUDA; } } while (0)

namespace {
template <typename T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return AUGB200_ERR_CUDA; }
        cap = n; return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T> struct PinBuf {
    T* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, n * sizeof(T));
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaMallocHost: ") + cudaGetErrorString(e); return AUGB200_ERR_CUDA; }
        cap = n; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
}  // namespace

struct augb200_model {
    HostModel hm;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> wave_ev;     /* sweep timing of the 2nd.. wave of a staged batch */
    std::vector<std::pair<int, int>> waves;                        /* (first window, count): groups of windows that share the arena in turn */
    DevBuf<sc_t> d_tab; DevModel dm_dev;      /* model with device table pointers; uploaded to the c_model constant */
    /* batch state */
    DevBuf<char> d_arena; DevBuf<char> d_pool; DevBuf<unsigned long long> d_pool_used; size_t pool_bytes = 0; DevBuf<char> d_dna; DevBuf<uint8_t> d_gc; DevBuf<WinDev> d_wins; DevBuf<int> d_counters;
    DevBuf<PackHdr> d_hdr; DevBuf<int32_t> d_obegin, d_oend; DevBuf<uint8_t> d_otype, d_otrunc;
    PinBuf<char> h_dna; PinBuf<uint8_t> h_gc; PinBuf<WinDev> h_wins; PinBuf<PackHdr> h_hdr; PinBuf<int> h_counters;
    PinBuf<int32_t> h_obegin, h_oend; PinBuf<uint8_t> h_otype, h_otrunc;
    size_t arena_budget = 0;
    int staged_n = 0; int ocap = 0;
    /* results handed to the caller */
    std::vector<int32_t> r_begin, r_end; std::vector<uint8_t> r_type, r_trunc;
    int64_t launches = 0; double sweep_ms = 0;
    /* sampling */
    int nsamp = 0;                       /* sampled paths per window of the current batch (0 = Viterbi only) */
    augb200_path* samples_out = nullptr;
    DevBuf<uint32_t> d_rng; size_t rng_n = 0; uint64_t rng_base = 0;      /* the device holds values [rng_base, rng_base + rng_n) of the stream */
    GlibcRand rng_gen;                   /* carried between calls: only values not produced yet are ever generated */
    uint64_t rand_pos = 0;               /* where in the rand() stream the windows of the next sampling call start */
    int64_t rand_used0 = 0;              /* draws window 0 of the last sampling call consumed */
    DevBuf<SampHdr> d_shdr; PinBuf<SampHdr> h_shdr; DevBuf<int32_t> d_sstatus; PinBuf<int32_t> h_sstatus;
    DevBuf<int32_t> d_sbegin, d_send; DevBuf<uint8_t> d_stype, d_strunc; PinBuf<int32_t> h_sbegin, h_send; PinBuf<uint8_t> h_stype, h_strunc;
    int scap = 0;
    std::vector<int32_t> rs_begin, rs_end; std::vector<uint8_t> rs_type, rs_trunc;
    std::vector<int32_t> rs_first;        /* per (window, sample) of the last sampling call: the first sample of the window with the same states */
    int64_t rs_total_states = 0;          /* states of all sampled paths of the last call, duplicates counted (rs_begin holds the unique ones) */
};

static int upload_windows(augb200_model* M, const augb200_window* w, const int* idx, int count, bool generous, size_t wave_cap = 0) {
    /* stage DNA (+ optional classes) of windows idx[0..count) and build their descriptors.  The workspaces of the windows live
     * in the arena; with wave_cap > 0 the windows are split into waves of (nearly) equal size <= wave_cap bytes that use the
     * arena one after the other (M->waves), so a staged batch may be larger than the device memory left for workspaces */
    size_t dna_bytes = 0, gc_bytes = 0, arena = 0, all = 0;
    for (int i = 0; i < count; i++) {
        const augb200_window& x = w[idx[i]];
        dna_bytes += (size_t)((x.length + 15) & ~15);
        if (x.gc_class) gc_bytes += (size_t)((x.length + 15) & ~15);
        all += make_layout(x.length, M->hm.dm.C, generous, M->nsamp > 0, M->nsamp, M->hm.dm.utr != 0, M->hm.dm.softmask != 0).total;
    }
    M->waves.clear();
    {
        const size_t nw = wave_cap && all > wave_cap ? (all + wave_cap - 1) / wave_cap : 1;
        const size_t target = (all + nw - 1) / nw;
        size_t cur = 0; int first = 0;
        for (int i = 0; i < count; i++) {
            size_t t = make_layout(w[idx[i]].length, M->hm.dm.C, generous, M->nsamp > 0, M->nsamp, M->hm.dm.utr != 0, M->hm.dm.softmask != 0).total;
            if (wave_cap && i > first && (cur + t > wave_cap || cur >= target)) { M->waves.push_back({first, i - first}); arena = std::max(arena, cur); first = i; cur = 0; }
            cur += t;
        }
        M->waves.push_back({first, count - first}); arena = std::max(arena, cur);
    }
    int rc;
    if ((rc = M->h_dna.reserve(dna_bytes + 16))) return rc;
    if ((rc = M->d_dna.reserve(dna_bytes + 16))) return rc;
    if (gc_bytes) { if ((rc = M->h_gc.reserve(gc_bytes + 16))) return rc; if ((rc = M->d_gc.reserve(gc_bytes + 16))) return rc; }
    if ((rc = M->d_arena.reserve(arena))) return rc;
    /* slab pool for the prefix arrays of second / third ... GC classes (first class lives in the window) */
    {
        size_t need = 0;
        if (!generous && M->hm.dm.C > 1)
            for (const auto& wv : M->waves) {
                size_t nd = 0;
                for (int i = wv.first; i < wv.first + wv.second; i++) nd += (size_t)(M->hm.dm.C - 1) * make_layout(w[idx[i]].length, M->hm.dm.C).slab;
                need = std::max(need, nd);
            }
        size_t room = M->arena_budget > arena ? M->arena_budget - arena : 0;
        size_t pool = std::min(need, room + M->arena_budget / 8);
        if (pool && (rc = M->d_pool.reserve(pool))) return rc;
        M->pool_bytes = pool;
        if ((rc = M->d_pool_used.reserve(1))) return rc;
    }
    if ((rc = M->h_wins.reserve(count))) return rc;
    if ((rc = M->d_wins.reserve(count))) return rc;
    size_t od = 0, og = 0, oa = 0; long total_path_cap = 0;
    size_t wvi = 0;
    for (int i = 0; i < count; i++) {
        if (wvi + 1 < M->waves.size() && i == M->waves[wvi + 1].first) { wvi++; oa = 0; }      /* next wave: the arena starts over */
        const augb200_window& x = w[idx[i]];
        WinDev& d = M->h_wins.p[i];
        d.L = x.length; d.lay = make_layout(x.length, M->hm.dm.C, generous, M->nsamp > 0, M->nsamp, M->hm.dm.utr != 0, M->hm.dm.softmask != 0);
        d.base = M->d_arena.p + oa; oa += d.lay.total;
        memcpy(M->h_dna.p + od, x.dna, x.length);
        d.dna = M->d_dna.p + od; od += (size_t)((x.length + 15) & ~15);
        if (x.gc_class) {
            uint8_t* g = M->h_gc.p + og;
            for (int j = 0; j < x.length; j++) {
                int c = x.gc_class[j];
                if (c < 0 || c >= M->hm.dm.C) return AUGB200_ERR_BAD_ARG;
                g[j] = (uint8_t)c;
            }
            d.gc_in = M->d_gc.p + og; og += (size_t)((x.length + 15) & ~15);
        } else d.gc_in = nullptr;
        total_path_cap += generous ? x.length + 64 : 512 + x.length / 64;
    }
    CK(cudaMemcpyAsync(M->d_dna.p, M->h_dna.p, od, cudaMemcpyHostToDevice, M->stream));
    if (og) CK(cudaMemcpyAsync(M->d_gc.p, M->h_gc.p, og, cudaMemcpyHostToDevice, M->stream));
    CK(cudaMemcpyAsync(M->d_wins.p, M->h_wins.p, (size_t)count * sizeof(WinDev), cudaMemcpyHostToDevice, M->stream));
    M->ocap = (int)std::min<long>(total_path_cap, 0x7fffffff);
    if ((rc = M->d_obegin.reserve(M->ocap)) || (rc = M->d_oend.reserve(M->ocap)) || (rc = M->d_otype.reserve(M->ocap)) || (rc = M->d_otrunc.reserve(M->ocap))) return rc;
    if ((rc = M->h_obegin.reserve(M->ocap)) || (rc = M->h_oend.reserve(M->ocap)) || (rc = M->h_otype.reserve(M->ocap)) || (rc = M->h_otrunc.reserve(M->ocap))) return rc;
    if ((rc = M->d_hdr.reserve(count)) || (rc = M->h_hdr.reserve(count))) return rc;
    if ((rc = M->d_counters.reserve(4)) || (rc = M->h_counters.reserve(4))) return rc;
    if (M->nsamp > 0) {
        long sc = 0; int maxL = 0;
        for (int i = 0; i < count; i++) { sc += M->h_wins.p[i].lay.samp_cap; maxL = std::max(maxL, M->h_wins.p[i].L); }
        M->scap = (int)std::min<long>(sc, 0x7fffffff);
        if ((rc = M->d_sbegin.reserve(M->scap)) || (rc = M->d_send.reserve(M->scap)) || (rc = M->d_stype.reserve(M->scap)) || (rc = M->d_strunc.reserve(M->scap))) return rc;
        if ((rc = M->h_sbegin.reserve(M->scap)) || (rc = M->h_send.reserve(M->scap)) || (rc = M->h_stype.reserve(M->scap)) || (rc = M->h_strunc.reserve(M->scap))) return rc;
        size_t nh = (size_t)count * M->nsamp;
        if ((rc = M->d_shdr.reserve(nh)) || (rc = M->h_shdr.reserve(nh)) || (rc = M->d_sstatus.reserve(2 * (size_t)count)) || (rc = M->h_sstatus.reserve(2 * (size_t)count))) return rc;
        /* the rand() stream every window draws from (an unseeded reference process per window = glibc seed 1), from the position the
         * caller set: the window [rand_pos, rand_pos + need) of the stream, generated from the carried generator state */
        const size_t need = (size_t)M->nsamp * ((size_t)maxL + 2);
        if (!(M->rng_n && M->rng_base <= M->rand_pos && M->rand_pos + need <= M->rng_base + M->rng_n)) {
            std::vector<uint32_t> host(need);
            M->rng_gen.seek(M->rand_pos); M->rng_gen.fill(host.data(), need);
            if ((rc = M->d_rng.reserve(need))) return rc;
            CK(cudaMemcpyAsync(M->d_rng.p, host.data(), need * 4, cudaMemcpyHostToDevice, M->stream));
            CK(cudaStreamSynchronize(M->stream));
            M->rng_n = need; M->rng_base = M->rand_pos;
        }
    }
    return 0;
}

static int run_kernels(augb200_model* M, int count) {
    CK(cudaMemsetAsync(M->d_counters.p, 0, 4 * sizeof(int), M->stream));
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->device);
    CK(cudaMemcpyToSymbolAsync(c_model, &M->dm_dev, sizeof(DevModel), 0, cudaMemcpyHostToDevice, M->stream));
    /* resident sweep warps per SM: the kernel is instruction-fetch bound, more warps than this thrash the instruction cache */
    static int bps = 0;
    if (!bps) { const char* e = getenv("AUGB200_SWEEP_BLOCKS_PER_SM"); bps = e ? atoi(e) : 4; if (bps < 1) bps = 1; }
    const bool utr = M->hm.dm.utr != 0;
    const size_t rng_off = M->rng_n && M->rand_pos >= M->rng_base ? (size_t)std::min<uint64_t>(M->rand_pos - M->rng_base, M->rng_n) : M->rng_n;
    const int nrng = (int)std::min<size_t>(M->rng_n - rng_off, 0x7fffffff);
    const uint32_t* d_rng = M->d_rng.p ? M->d_rng.p + rng_off : nullptr;
    if (M->waves.empty() || M->waves.back().first + M->waves.back().second != count) { M->waves.clear(); M->waves.push_back({0, count}); }
    while (M->wave_ev.size() + 1 < M->waves.size()) {
        std::pair<cudaEvent_t, cudaEvent_t> e;
        CK(cudaEventCreate(&e.first)); CK(cudaEventCreate(&e.second));
        M->wave_ev.push_back(e);
    }
    for (size_t wv = 0; wv < M->waves.size(); wv++) {
        const int first = M->waves[wv].first, n = M->waves[wv].second;
        const WinDev* wins = M->d_wins.p + first;
        if (wv) CK(cudaMemsetAsync(M->d_counters.p, 0, sizeof(int), M->stream));      /* the window queue of the sweep; the output offsets run on */
        CK(cudaMemsetAsync(M->d_pool_used.p, 0, sizeof(unsigned long long), M->stream));
        int gprep = std::min(n, sms * 8);
        k_prep<<<gprep, PREP_BS, 0, M->stream>>>(wins, n, M->pool_bytes ? M->d_pool.p : nullptr, (unsigned long long)M->pool_bytes, M->d_pool_used.p);
        CK(cudaEventRecord(wv ? M->wave_ev[wv - 1].first : M->ev0, M->stream));
        int gsweep = std::min((n + SWEEP_TEAMS - 1) / SWEEP_TEAMS, sms * bps);
        if (utr) gsweep = std::min(gsweep, sms * std::min(bps, 3));
        if (M->nsamp > 0 && utr) k_sweep_sample_utr<<<gsweep, SWEEP_WARPS * 32, 0, M->stream>>>(wins, n, M->d_counters.p, d_rng, nrng);
        else if (M->nsamp > 0) k_sweep_sample<<<gsweep, SWEEP_WARPS * 32, 0, M->stream>>>(wins, n, M->d_counters.p, d_rng, nrng);
        else if (utr) k_sweep_utr<<<gsweep, SWEEP_WARPS * 32, 0, M->stream>>>(wins, n, M->d_counters.p);
        else k_sweep<<<gsweep, SWEEP_WARPS * 32, 0, M->stream>>>(wins, n, M->d_counters.p);
        CK(cudaEventRecord(wv ? M->wave_ev[wv - 1].second : M->ev1, M->stream));
        k_backtrace<<<(n + 63) / 64, 64, 0, M->stream>>>(wins, n);
        k_pack<<<n, 64, 0, M->stream>>>(wins, n, M->d_hdr.p + first, M->d_counters.p + 1, M->d_obegin.p, M->d_oend.p, M->d_otype.p, M->d_otrunc.p, M->ocap);
        if (M->nsamp > 0) {
            k_pack_samples<<<n, 64, 0, M->stream>>>(wins, n, M->d_shdr.p + (size_t)first * M->nsamp, M->d_sstatus.p + 2 * (size_t)first, M->d_counters.p + 2,
                                                     M->d_sbegin.p, M->d_send.p, M->d_stype.p, M->d_strunc.p, M->scap);
            M->launches += 1;
        }
        CK(cudaGetLastError());
        M->launches += 4;
    }
    return 0;
}

static int fetch_results(augb200_model* M, int count, augb200_path* out, const int* idx) {
    CK(cudaMemcpyAsync(M->h_hdr.p, M->d_hdr.p, (size_t)count * sizeof(PackHdr), cudaMemcpyDeviceToHost, M->stream));
    CK(cudaMemcpyAsync(M->h_counters.p, M->d_counters.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, M->stream));
    CK(cudaStreamSynchronize(M->stream));
    int total = std::min(M->h_counters.p[1], M->ocap);
    if (total > 0) {
        CK(cudaMemcpyAsync(M->h_obegin.p, M->d_obegin.p, (size_t)total * 4, cudaMemcpyDeviceToHost, M->stream));
        CK(cudaMemcpyAsync(M->h_oend.p, M->d_oend.p, (size_t)total * 4, cudaMemcpyDeviceToHost, M->stream));
        CK(cudaMemcpyAsync(M->h_otype.p, M->d_otype.p, (size_t)total, cudaMemcpyDeviceToHost, M->stream));
        CK(cudaMemcpyAsync(M->h_otrunc.p, M->d_otrunc.p, (size_t)total, cudaMemcpyDeviceToHost, M->stream));
        CK(cudaStreamSynchronize(M->stream));
    }
    if (M->nsamp > 0 && M->samples_out) {
        size_t nh = (size_t)count * M->nsamp;
        CK(cudaMemcpyAsync(M->h_shdr.p, M->d_shdr.p, nh * sizeof(SampHdr), cudaMemcpyDeviceToHost, M->stream));
        CK(cudaMemcpyAsync(M->h_sstatus.p, M->d_sstatus.p, (size_t)count * 8, cudaMemcpyDeviceToHost, M->stream));
        int stot = std::min(M->h_counters.p[2], M->scap);
        if (stot > 0) {
            CK(cudaMemcpyAsync(M->h_sbegin.p, M->d_sbegin.p, (size_t)stot * 4, cudaMemcpyDeviceToHost, M->stream));
            CK(cudaMemcpyAsync(M->h_send.p, M->d_send.p, (size_t)stot * 4, cudaMemcpyDeviceToHost, M->stream));
            CK(cudaMemcpyAsync(M->h_stype.p, M->d_stype.p, (size_t)stot, cudaMemcpyDeviceToHost, M->stream));
            CK(cudaMemcpyAsync(M->h_strunc.p, M->d_strunc.p, (size_t)stot, cudaMemcpyDeviceToHost, M->stream));
        }
        CK(cudaStreamSynchronize(M->stream));
        size_t sbase = M->rs_begin.size();
        M->rs_begin.insert(M->rs_begin.end(), M->h_sbegin.p, M->h_sbegin.p + stot);
        M->rs_end.insert(M->rs_end.end(), M->h_send.p, M->h_send.p + stot);
        M->rs_type.insert(M->rs_type.end(), M->h_stype.p, M->h_stype.p + stot);
        M->rs_trunc.insert(M->rs_trunc.end(), M->h_strunc.p, M->h_strunc.p + stot);
        for (int i = 0; i < count; i++)
            for (int k = 0; k < M->nsamp; k++) {
                const SampHdr& h = M->h_shdr.p[(size_t)i * M->nsamp + k]; augb200_path& p = M->samples_out[(size_t)idx[i] * M->nsamp + k];
                p.n = h.n; p.status = M->h_sstatus.p[2 * i]; p.log_prob = h.logp;
                p.begin = (const int32_t*)(uintptr_t)(sbase + h.offset);
                M->rs_first[(size_t)idx[i] * M->nsamp + k] = h.first;
                M->rs_total_states += h.n;
            }
        for (int i = 0; i < count; i++) if (idx[i] == 0) M->rand_used0 = M->h_sstatus.p[2 * i + 1];
    }
    float ms = 0; if (cudaEventElapsedTime(&ms, M->ev0, M->ev1) == cudaSuccess) M->sweep_ms += ms;
    for (size_t wv = 1; wv < M->waves.size() && wv - 1 < M->wave_ev.size(); wv++)
        if (cudaEventElapsedTime(&ms, M->wave_ev[wv - 1].first, M->wave_ev[wv - 1].second) == cudaSuccess) M->sweep_ms += ms;
    /* append to the result store; pointers are fixed up by the caller once all sub-batches are in */
    size_t base = M->r_begin.size();
    M->r_begin.insert(M->r_begin.end(), M->h_obegin.p, M->h_obegin.p + total);
    M->r_end.insert(M->r_end.end(), M->h_oend.p, M->h_oend.p + total);
    M->r_type.insert(M->r_type.end(), M->h_otype.p, M->h_otype.p + total);
    M->r_trunc.insert(M->r_trunc.end(), M->h_otrunc.p, M->h_otrunc.p + total);
    for (int i = 0; i < count; i++) {
        const PackHdr& h = M->h_hdr.p[i]; augb200_path& p = out[idx[i]];
        p.n = h.n; p.status = h.status; p.log_prob = ldexp((double)h.score, -FRAC_BITS);
        /* store offsets in the pointer fields for now */
        p.begin = (const int32_t*)(uintptr_t)(base + h.offset);
    }
    /* a window whose sample buffers overflowed is decoded again as a whole (after the Viterbi status above is in place) */
    if (M->nsamp > 0 && M->samples_out)
        for (int i = 0; i < count; i++) if (M->h_sstatus.p[2 * i] == AUGB200_ERR_CAPACITY && out[idx[i]].status == 0) out[idx[i]].status = AUGB200_ERR_CAPACITY;
    return 0;
}

static void fix_pointers(augb200_model* M, int n, augb200_path* out) {
    for (int i = 0; i < n; i++) {
        size_t off = (size_t)(uintptr_t)out[i].begin;
        out[i].begin = M->r_begin.data() + off; out[i].end = M->r_end.data() + off;
        out[i].type = M->r_type.data() + off; out[i].truncated = M->r_trunc.data() + off;
    }
}

extern "C" {

int augb200_model_create(const void* blob, size_t nbytes, int device, augb200_model** out) {
    if (!blob || !out) return AUGB200_ERR_BAD_ARG;
    *out = nullptr;
    augb200_model* M = new augb200_model();
    int rc = M->hm.build(blob, nbytes);
    if (rc) { g_cuda_err = M->hm.err; delete M; return rc; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_cuda_err = "no CUDA device"; delete M; return AUGB200_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { delete M; return AUGB200_ERR_BAD_ARG; }
    M->device = device;
    auto fail = [&](int code) { augb200_model_destroy(M); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(AUGB200_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(AUGB200_ERR_CUDA);
    if (cudaEventCreate(&M->ev0) != cudaSuccess || cudaEventCreate(&M->ev1) != cudaSuccess) return fail(AUGB200_ERR_CUDA);
    if (M->d_tab.reserve(M->hm.tab.size())) return fail(AUGB200_ERR_CUDA);
    if (cudaMemcpy(M->d_tab.p, M->hm.tab.data(), M->hm.tab.size() * sizeof(sc_t), cudaMemcpyHostToDevice) != cudaSuccess) return fail(AUGB200_ERR_CUDA);
    M->dm_dev = M->hm.rebased(M->d_tab.p);
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return fail(AUGB200_ERR_CUDA);
    M->arena_budget = (size_t)(fr * 0.80);
    if (const char* e = getenv("AUGB200_ARENA_MB")) { size_t mb = strtoull(e, nullptr, 10); if (mb) M->arena_budget = std::min(M->arena_budget, mb << 20); }
    *out = M;
    return AUGB200_OK;
}

void augb200_model_destroy(augb200_model* M) {
    if (!M) return;
    cudaSetDevice(M->device);
    M->d_tab.release(); M->d_arena.release(); M->d_pool.release(); M->d_pool_used.release(); M->d_dna.release(); M->d_gc.release(); M->d_wins.release(); M->d_counters.release();
    M->d_hdr.release(); M->d_obegin.release(); M->d_oend.release(); M->d_otype.release(); M->d_otrunc.release();
    M->h_dna.release(); M->h_gc.release(); M->h_wins.release(); M->h_hdr.release(); M->h_counters.release();
    M->h_obegin.release(); M->h_oend.release(); M->h_otype.release(); M->h_otrunc.release();
    M->d_rng.release(); M->d_shdr.release(); M->h_shdr.release(); M->d_sstatus.release(); M->h_sstatus.release();
    M->d_sbegin.release(); M->d_send.release(); M->d_stype.release(); M->d_strunc.release(); M->h_sbegin.release(); M->h_send.release(); M->h_stype.release(); M->h_strunc.release();
    if (M->ev0) cudaEventDestroy(M->ev0);
    if (M->ev1) cudaEventDestroy(M->ev1);
    for (auto& e : M->wave_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (M->stream) cudaStreamDestroy(M->stream);
    delete M;
}

int augb200_model_statecount(const augb200_model* M) { return M ? M->hm.dm.S : 0; }
int augb200_model_num_gc_classes(const augb200_model* M) { return M ? M->hm.dm.C : 0; }

static int check_windows(int32_t n, const augb200_window* w) {
    if (n < 0 || (n && !w)) return AUGB200_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) {
        if (!w[i].dna || w[i].length < 2) return AUGB200_ERR_BAD_ARG;
        if (w[i].length > AUGB200_MAX_WINDOW) return AUGB200_ERR_TOO_LONG;
    }
    return 0;
}

/* The kernels read the model from one __constant__ symbol per device (c_model, re-uploaded by every call): calls on models that live on
 * the same device take turns.  (The staged calls are asynchronous and are meant for one model per device.) */
static std::mutex g_device_lock[64];

static int decode_batch_impl(augb200_model* M, int32_t n, const augb200_window* w, augb200_path* out, int nsamp, augb200_path* samples) {
    if (!M || !out) return AUGB200_ERR_BAD_ARG;
    std::lock_guard<std::mutex> device_turn(g_device_lock[M->device & 63]);
    struct Reset { augb200_model* m; ~Reset() { m->nsamp = 0; m->samples_out = nullptr; } } reset_on_exit{M};      /* also on the error returns */
    M->staged_n = 0;                     /* a staged batch shares the descriptors and the arena with this call: it is gone */
    M->nsamp = nsamp; M->samples_out = samples;
    M->rs_begin.clear(); M->rs_end.clear(); M->rs_type.clear(); M->rs_trunc.clear();
    M->rs_first.assign((size_t)(n > 0 ? n : 0) * (size_t)(nsamp > 0 ? nsamp : 0), 0); M->rs_total_states = 0;
    int rc = check_windows(n, w); if (rc) return rc;
    CK(cudaSetDevice(M->device));
    M->r_begin.clear(); M->r_end.clear(); M->r_type.clear(); M->r_trunc.clear();
    M->launches = 0; M->sweep_ms = 0;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    for (int pass = 0; pass < 2 && !order.empty(); pass++) {
        const bool generous = pass == 1;      /* second pass: windows whose default-sized structures overflowed */
        size_t first = 0;
        /* sub-batches under the arena budget, of (nearly) equal size: every sub-batch ends with a tail of windows that run at low
         * occupancy, so two halves beat "as many as fit, then the rest" */
        const size_t cap = M->arena_budget - M->arena_budget / 8;
        size_t all = 0;
        for (int i : order) all += make_layout(w[i].length, M->hm.dm.C, generous, M->nsamp > 0, M->nsamp, M->hm.dm.utr != 0, M->hm.dm.softmask != 0).total;
        const size_t nsub = all > cap ? (all + cap - 1) / cap : 1, target = (all + nsub - 1) / nsub;
        while (first < order.size()) {
            size_t bytes = 0; int count = 0;
            while (first + count < order.size()) {
                size_t t = make_layout(w[order[first + count]].length, M->hm.dm.C, generous, M->nsamp > 0, M->nsamp, M->hm.dm.utr != 0, M->hm.dm.softmask != 0).total;
                if (count && (bytes + t > cap || bytes >= target)) break;
                bytes += t; count++;
            }
            if ((rc = upload_windows(M, w, order.data() + first, count, generous))) return rc;
            if ((rc = run_kernels(M, count))) return rc;
            if ((rc = fetch_results(M, count, out, order.data() + first))) return rc;
            first += count;
        }
        std::vector<int> again;
        if (!generous) for (int i : order) if (out[i].status == AUGB200_ERR_CAPACITY) again.push_back(i);
        order.swap(again);
    }
    fix_pointers(M, n, out);
    if (samples)
        for (size_t i = 0; i < (size_t)n * nsamp; i++) {
            size_t off = (size_t)(uintptr_t)samples[i].begin;
            samples[i].begin = M->rs_begin.data() + off; samples[i].end = M->rs_end.data() + off;
            samples[i].type = M->rs_type.data() + off; samples[i].truncated = M->rs_trunc.data() + off;
        }
    return AUGB200_OK;
}

int augb200_decode_batch(augb200_model* M, int32_t n, const augb200_window* w, augb200_path* out) { return decode_batch_impl(M, n, w, out, 0, nullptr); }

int augb200_decode_batch_sampling(augb200_model* M, int32_t n, const augb200_window* w, int32_t nsample, augb200_path* out, augb200_path* samples) {
    if (nsample < 2 || !samples) return AUGB200_ERR_BAD_ARG;
    return decode_batch_impl(M, n, w, out, nsample - 1, samples);
}

int augb200_decode(augb200_model* M, const augb200_window* w, augb200_path* out) { return augb200_decode_batch(M, 1, w, out); }

int augb200_decode_batch_multi(augb200_model* const* models, int32_t n_models, int32_t n, const augb200_window* w, int32_t nsample,
                               augb200_path* out, augb200_path* samples) {
    if (!models || n_models < 1 || !out || n < 0 || (n && !w) || (nsample != 0 && (nsample < 2 || !samples))) return AUGB200_ERR_BAD_ARG;
    for (int d = 0; d < n_models; d++) {
        if (!models[d]) return AUGB200_ERR_BAD_ARG;
        for (int e = 0; e < d; e++) if (models[e] == models[d]) return AUGB200_ERR_BAD_ARG;
    }
    const int ns = nsample ? nsample - 1 : 0;
    /* window i goes to model i mod n_models (block-cyclic); one host thread per device drives its share */
    std::vector<std::vector<augb200_window>> sub(n_models);
    std::vector<std::vector<augb200_path>> sout(n_models), ssamp(n_models);
    for (int i = 0; i < n; i++) sub[i % n_models].push_back(w[i]);
    std::vector<int> rcs(n_models, 0);
    std::vector<std::string> errs(n_models);
    std::vector<std::thread> th;
    for (int d = 0; d < n_models; d++) {
        sout[d].resize(sub[d].size()); ssamp[d].resize(sub[d].size() * (size_t)ns);
        th.emplace_back([&, d]() {
            if (sub[d].empty()) return;
            rcs[d] = decode_batch_impl(models[d], (int32_t)sub[d].size(), sub[d].data(), sout[d].data(), ns, ns ? ssamp[d].data() : nullptr);
            if (rcs[d]) errs[d] = g_cuda_err;          /* (the error text is thread-local) */
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < n_models; d++) if (rcs[d]) { g_cuda_err = errs[d]; return rcs[d]; }
    /* back into input order; the arrays stay owned by the model that decoded the window */
    for (int i = 0; i < n; i++) {
        const int d = i % n_models; const size_t k = (size_t)(i / n_models);
        out[i] = sout[d][k];
        for (int q = 0; q < ns; q++) samples[(size_t)i * ns + q] = ssamp[d][k * ns + q];
    }
    return AUGB200_OK;
}

int augb200_set_rand_position(augb200_model* M, uint64_t draws_consumed) {
    if (!M) return AUGB200_ERR_BAD_ARG;
    M->rand_pos = draws_consumed;
    return AUGB200_OK;
}
int64_t augb200_last_rand_consumed(const augb200_model* M) { return M ? M->rand_used0 : 0; }

int augb200_stage_batch(augb200_model* M, int32_t n, const augb200_window* w) {
    if (!M) return AUGB200_ERR_BAD_ARG;
    int rc = check_windows(n, w); if (rc) return rc;
    CK(cudaSetDevice(M->device));
    /* the DNA and the descriptors of all windows stay on the device; workspaces are taken from the arena wave by wave */
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    M->nsamp = 0; M->samples_out = nullptr;
    if ((rc = upload_windows(M, w, order.data(), n, false, M->arena_budget - M->arena_budget / 8))) return rc;
    CK(cudaStreamSynchronize(M->stream));
    M->staged_n = n;
    return AUGB200_OK;
}
int augb200_run_staged(augb200_model* M) {
    if (!M || M->staged_n <= 0) return AUGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(M->device));
    M->launches = 0;
    return run_kernels(M, M->staged_n);
}
int augb200_fetch_staged(augb200_model* M, augb200_path* out) {
    if (!M || !out || M->staged_n <= 0) return AUGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(M->device));
    M->r_begin.clear(); M->r_end.clear(); M->r_type.clear(); M->r_trunc.clear(); M->sweep_ms = 0;
    std::vector<int> order(M->staged_n);
    for (int i = 0; i < M->staged_n; i++) order[i] = i;
    int rc = fetch_results(M, M->staged_n, out, order.data()); if (rc) return rc;
    fix_pointers(M, M->staged_n, out);
    return AUGB200_OK;
}
void* augb200_model_stream(augb200_model* M) { return M ? (void*)M->stream : nullptr; }
int64_t augb200_last_launch_count(const augb200_model* M) { return M ? M->launches : 0; }
double augb200_last_sweep_ms(const augb200_model* M) { return M ? M->sweep_ms : 0; }

int64_t augb200_result_store(const augb200_model* M, const int32_t** b, const int32_t** e, const uint8_t** t, const uint8_t** tr) {
    if (!M) return 0;
    if (b) *b = M->r_begin.data();
    if (e) *e = M->r_end.data();
    if (t) *t = M->r_type.data();
    if (tr) *tr = M->r_trunc.data();
    return (int64_t)M->r_begin.size();
}

int64_t augb200_sample_store(const augb200_model* M, const int32_t** b, const int32_t** e, const uint8_t** t, const uint8_t** tr) {
    if (!M) return 0;
    if (b) *b = M->rs_begin.data();
    if (e) *e = M->rs_end.data();
    if (t) *t = M->rs_type.data();
    if (tr) *tr = M->rs_trunc.data();
    return (int64_t)M->rs_begin.size();
}

int64_t augb200_sample_first_occurrence(const augb200_model* M, const int32_t** first) {
    if (!M) return 0;
    if (first) *first = M->rs_first.data();
    return M->rs_total_states;
}

const char* augb200_strerror(int code) {
    switch (code) {
    case AUGB200_OK: return "ok";
    case AUGB200_ERR_BAD_BLOB: return "malformed parameter blob";
    case AUGB200_ERR_UNSUPPORTED: return "model not supported by this library";
    case AUGB200_ERR_NO_DEVICE: return "no CUDA device";
    case AUGB200_ERR_CUDA: return "CUDA error";
    case AUGB200_ERR_BAD_ARG: return "bad argument";
    case AUGB200_ERR_NO_PATH: return "No feasible path found in HMM";
    case AUGB200_ERR_STUCK: return "Viterbi got stuck";
    case AUGB200_ERR_CAPACITY: return "internal capacity exceeded";
    case AUGB200_ERR_TOO_LONG: return "window longer than AUGB200_MAX_WINDOW";
    default: return "unknown error";
    }
}
const char* augb200_last_cuda_error(void) { return g_cuda_err.c_str(); }

}  // extern "C"
