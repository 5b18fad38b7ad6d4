/*
 * simt32.h — TEST-ONLY 32-lane executor for the warp-cooperative kernel source.
 *
 * The sweep (ghmm_sweep.h) is written as SIMT code: 32 lanes run the same routine, exchange values with
 * ballots / shuffles and publish lane-0 writes with warp barriers.  The plain host build (hostemu.cc)
 * compiles it with ONE lane, which cannot exercise the lane-group logic of the device (three reading
 * frames in one pass, 8 lanes per frame; lane = (frame, ancestor); per-group reductions).  This header
 * runs the device flavour of the source on the CPU: 32 fibers (ucontext) per warp, scheduled round robin
 * by one thread; every warp collective is a rendezvous — a fiber deposits its operand and yields, and it
 * resumes only after all 32 fibers have deposited theirs.  Collectives alternate between two exchange
 * buffers, so a fast fiber cannot overwrite operands a slow one still has to read.  Each operand carries
 * a tag (kind of collective); fibers that meet in different collectives — the classic divergent
 * __shfl/__ballot bug — abort the run.
 *
 * It provides the CUDA intrinsics the kernel source uses, under their CUDA names.
 */
#pragma once
#include <ucontext.h>
#include <execinfo.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

constexpr int NL = 32;

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[NL];
    std::vector<char> stack[NL];
    bool finished[NL];
    int cur = 0;
    uint64_t buf[2][NL];
    uint32_t tag[2][NL];
    long seqno[2][NL];                /* which collective of the lane the deposit belongs to */
    long seq[NL];
    int parity[NL];
    long collectives = 0;
    std::function<void()> body;
};

inline Warp*& current() { static thread_local Warp* w = nullptr; return w; }
inline int lane() { return current()->cur; }

inline void trampoline() {
    Warp* w = current();
    w->body();
    w->finished[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

/* run `body` on 32 lanes in lockstep at the collectives */
inline void run(const std::function<void()>& body, size_t stack_bytes = 1 << 20) {
    Warp w; w.body = body;
    Warp* prev = current(); current() = &w;
    for (int l = 0; l < NL; l++) {
        w.stack[l].assign(stack_bytes, 0); w.finished[l] = false; w.parity[l] = 0; w.seq[l] = 0; w.seqno[0][l] = w.seqno[1][l] = -1;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stack[l].data(); w.ctx[l].uc_stack.ss_size = stack_bytes; w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], (void (*)())trampoline, 0);
    }
    /* lane order inside a phase: ascending (lane 0's writes are seen by everybody in the same phase: catches a read that should have
     * come BEFORE the write), descending with SIMT32_REVERSE=1 (lane 0 runs last: catches a read that needs a barrier AFTER the write) */
    const bool rev = getenv("SIMT32_REVERSE") != nullptr;
    for (;;) {
        bool any = false;
        for (int i = 0; i < NL; i++) {
            const int l = rev ? NL - 1 - i : i;
            if (w.finished[l]) continue;
            any = true; w.cur = l;
            swapcontext(&w.sched, &w.ctx[l]);
        }
        if (!any) break;
    }
    current() = prev;
}

/* deposit v, wait for all lanes, return the buffer with everybody's operand */
inline const uint64_t* rendezvous(uint64_t v, uint32_t kind) {
    Warp* w = current(); const int l = w->cur, p = w->parity[l];
    const long me = ++w->seq[l];
    w->buf[p][l] = v; w->tag[p][l] = kind; w->seqno[p][l] = me;
    if (l == 0) w->collectives++;
    swapcontext(&w->ctx[l], &w->sched);
    w->cur = l;                                    /* (the scheduler set it before resuming us) */
    for (int i = 0; i < NL; i++) {
        /* every lane must be in its `me`-th collective, of the same kind (a lane that returned early holds an older deposit) */
        if (w->seqno[p][i] != me || w->tag[p][i] != kind) {
            fprintf(stderr, "simt32: lanes %d and %d met in different warp collectives (%u vs %u, finished=%d) — divergent collective\n",
                    l, i, kind, w->tag[p][i], (int)w->finished[i]);
            { void* bt[48]; int nb = backtrace(bt, 48); backtrace_symbols_fd(bt, nb, 2); }
            abort();
        }
    }
    w->parity[l] = p ^ 1;
    return w->buf[p];
}

template <typename T> inline uint64_t bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

}  // namespace simt

/* ---- the intrinsics of the kernel source (full-mask warp collectives only, as the source uses them) ---- */
inline void __syncwarp(unsigned = 0xffffffffu) { simt::rendezvous(0, 1); }
inline unsigned __ballot_sync(unsigned, bool p) {
    const uint64_t* b = simt::rendezvous(p ? 1 : 0, 2);
    unsigned r = 0; for (int i = 0; i < 32; i++) if (b[i]) r |= 1u << i;
    return r;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { const uint64_t* b = simt::rendezvous(simt::bits(v), 3); return simt::unbits<T>(b[src & 31]); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) { const int l = simt::lane(); const uint64_t* b = simt::rendezvous(simt::bits(v), 4); return simt::unbits<T>(b[(l ^ o) & 31]); }
inline unsigned __match_any_sync(unsigned, int v) {
    const uint64_t* b = simt::rendezvous(simt::bits(v), 5);
    unsigned r = 0; for (int i = 0; i < 32; i++) if (simt::unbits<int>(b[i]) == v) r |= 1u << i;
    return r;
}
inline int __ffs(unsigned b) { return b ? __builtin_ctz(b) + 1 : 0; }
inline int __popc(unsigned b) { return __builtin_popcount(b); }
