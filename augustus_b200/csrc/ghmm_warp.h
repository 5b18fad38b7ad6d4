/*
 * ghmm_warp.h — warp-cooperative primitives used by the sweep.
 *
 * On the device one warp decodes one window: lanes split candidate (predecessor, duration) pairs and
 * motif positions, and combine with shuffle reductions.  The same source compiles for the host with a
 * single "lane" (tests/host emulator only — never part of the product library), so the control flow of
 * the kernel can be checked against the oracle in a container without a GPU.
 */
#pragma once
#include <math.h>
#include "ghmm_defs.h"

namespace augb {

/* AUGB_SIMT: the 32-lane flavour of the source — the device, or (AUGB_SIMT32, set by the test suite only) a CPU executor that runs
 * 32 fibers per warp and supplies the warp intrinsics; the product library is never built with it */
#if defined(__CUDA_ARCH__) || defined(AUGB_SIMT32)
#define AUGB_SIMT 1
#else
#define AUGB_SIMT 0
#endif
#if AUGB_SIMT
#define AUGB_NLANES 32
#if defined(__CUDA_ARCH__)
AUGB_D int lane_id() { return threadIdx.x & 31; }
#else
inline int lane_id() { return simt::lane(); }
#endif
AUGB_D void wsync() { __syncwarp(); }
AUGB_D sc_t wmax(sc_t v) {
    for (int o = 16; o; o >>= 1) { sc_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D sc_t wsum(sc_t v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
AUGB_D int wmaxi(int v) {
    for (int o = 16; o; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D double wmaxd(double v) {
    for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D double wsumd(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
AUGB_D unsigned wballot(bool p) { return __ballot_sync(0xffffffffu, p); }
AUGB_D int wbcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AUGB_D sc_t wbcast64(sc_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AUGB_D int wffs(unsigned b) { return __ffs(b) - 1; }
AUGB_D int wpopc(unsigned b) { return __popc(b); }
#else
#define AUGB_NLANES 1
inline int lane_id() { return 0; }
inline void wsync() {}
inline sc_t wmax(sc_t v) { return v; }
inline sc_t wsum(sc_t v) { return v; }
inline int wmaxi(int v) { return v; }
inline double wmaxd(double v) { return v; }
inline double wsumd(double v) { return v; }
inline unsigned wballot(bool p) { return p ? 1u : 0u; }
inline int wbcast(int v, int) { return v; }
inline sc_t wbcast64(sc_t v, int) { return v; }
inline int wffs(unsigned b) { return b ? __builtin_ctz(b) : -1; }
inline int wpopc(unsigned b) { return __builtin_popcount(b); }
#endif

/* arg-max over lanes of (score, key): the highest score wins, ties go to the highest key.  Returns the
 * winning lane or -1 if every score is SC_NEG.  Keys are chosen by the callers so that "highest key" is
 * the option the reference meets first in its loop order (strict '>' keeps the first maximum). */
AUGB_D int wargbest(sc_t score, int key) {
    /* the common cases need no reduction: no lane holds a candidate, or exactly one does */
    const unsigned have = wballot(!isneg(score));
    if (have == 0) return -1;
    if ((have & (have - 1u)) == 0) return wffs(have);
    sc_t m = wmax(score);
    if (isneg(m)) return -1;
    int k = wmaxi(score == m ? key : -0x7fffffff);
    return wffs(wballot(score == m && key == k));
}

/* the same per group of gl lanes (gl a power of two, groups aligned; gmask = the lanes of the caller's group): every lane gets the
 * winning lane of its own group */
AUGB_D int gargbest(sc_t score, int key, int gl, unsigned gmask) {
#if AUGB_SIMT
    if (gl >= 32) return wargbest(score, key);
    const unsigned have = wballot(!isneg(score));
    const unsigned hg = have & gmask;
    /* no reduction unless some group holds more than one candidate */
    if (wballot((hg & (hg - 1u)) != 0) == 0) return hg ? wffs(hg) : -1;
    sc_t m = score;
    for (int o = gl >> 1; o; o >>= 1) { sc_t t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
    int k = score == m ? key : -0x7fffffff;
    for (int o = gl >> 1; o; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, k, o); k = t > k ? t : k; }
    const unsigned win = wballot(!isneg(score) && score == m && key == k) & gmask;
    return win ? wffs(win) : -1;
#else
    return wargbest(score, key);
#endif
}

/* running log-sum-exp: value = m + ln(s); empty = (m = -inf, s = 0) */
struct Lse {
    double m, s;
    AUGB_HD void clear() { m = -1e308; s = 0; }
    AUGB_HD void add(double lp) {
        if (lp > m) { s = s * exp(m - lp) + 1.0; m = lp; } else s += exp(lp - m);
    }
    AUGB_HD bool empty() const { return !(s > 0); }
    AUGB_HD double value() const { return s > 0 ? m + log(s) : -1e308; }
};
/* combine the per-lane accumulators of a group of gl lanes (aligned, power of two); every lane gets its group's result */
AUGB_D Lse glse(Lse a, int gl) {
#if AUGB_SIMT
    double M = a.m;
    for (int o = gl >> 1; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, M, o); M = t > M ? t : M; }
    double s = a.s > 0 ? a.s * exp(a.m - M) : 0.0;
    for (int o = gl >> 1; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    Lse r; r.m = M; r.s = s;
    return r;
#else
    return a;
#endif
}
AUGB_D double wbcastd(double v, int src) {
#if AUGB_SIMT
    return __shfl_sync(0xffffffffu, v, src);
#else
    return v;
#endif
}
/* combine the per-lane accumulators of a warp */
AUGB_D Lse wlse(Lse a) {
    double M = wmaxd(a.m);
    Lse r; r.m = M; r.s = wsumd(a.s > 0 ? a.s * exp(a.m - M) : 0.0);
    return r;
}

}  // namespace augb
