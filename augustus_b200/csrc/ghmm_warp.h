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
#ifndef AUGB_TEAM
#define AUGB_TEAM 32
#endif
#if defined(__CUDA_ARCH__) || defined(AUGB_SIMT32)
#define AUGB_SIMT 1
#else
#define AUGB_SIMT 0
#endif
#if AUGB_SIMT
/* AUGB_TEAM = lanes per window (power of two, default 32 = one warp per window).  With a smaller team a warp holds 32 / AUGB_TEAM
 * windows whose teams walk the same routines together: the collectives below act on the lanes of the caller's team only
 * (sub-warp masks; ballots are returned team-relative, bit 0 = first lane of the team). */
#define AUGB_NLANES AUGB_TEAM
/* the three reading frames of a state kind in one pass need 3 x 8 lanes */
#define AUGB_GROUP3 (AUGB_NLANES == 32)
#if defined(__CUDA_ARCH__)
AUGB_D int lane_id() { return threadIdx.x & (AUGB_TEAM - 1); }
#if AUGB_TEAM == 32
AUGB_D unsigned tmask() { return 0xffffffffu; }
AUGB_D int tbase() { return 0; }
#else
AUGB_D int tbase() { return (threadIdx.x & 31) & ~(AUGB_TEAM - 1); }
AUGB_D unsigned tmask() { return ((1u << AUGB_TEAM) - 1u) << tbase(); }
#endif
#else
inline int lane_id() { return simt::lane(); }
inline unsigned tmask() { return 0xffffffffu; }
inline int tbase() { return 0; }
#endif
AUGB_D void wsync() { __syncwarp(tmask()); }
AUGB_D sc_t wmax(sc_t v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) { sc_t t = __shfl_xor_sync(tmask(), v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D sc_t wsum(sc_t v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) v += __shfl_xor_sync(tmask(), v, o);
    return v;
}
AUGB_D int wmaxi(int v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) { int t = __shfl_xor_sync(tmask(), v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D double wmaxd(double v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) { double t = __shfl_xor_sync(tmask(), v, o); v = t > v ? t : v; }
    return v;
}
AUGB_D double wsumd(double v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) v += __shfl_xor_sync(tmask(), v, o);
    return v;
}
AUGB_D int wmini(int v) {
    for (int o = AUGB_TEAM / 2; o; o >>= 1) { int t = __shfl_xor_sync(tmask(), v, o); v = t < v ? t : v; }
    return v;
}
AUGB_D unsigned wballot(bool p) { return __ballot_sync(tmask(), p) >> tbase(); }
/* the lanes (team-relative bits) that hold the same value */
AUGB_D unsigned wmatch(int v) { return __match_any_sync(tmask(), v) >> tbase(); }
AUGB_D int wbcast(int v, int src) { return __shfl_sync(tmask(), v, tbase() + src); }
AUGB_D sc_t wbcast64(sc_t v, int src) { return __shfl_sync(tmask(), v, tbase() + src); }
AUGB_D int wffs(unsigned b) { return __ffs(b) - 1; }
AUGB_D int wpopc(unsigned b) { return __popc(b); }
#else
#define AUGB_NLANES 1
#define AUGB_GROUP3 0
AUGB_HD int lane_id() { return 0; }
AUGB_HD void wsync() {}
AUGB_HD sc_t wmax(sc_t v) { return v; }
AUGB_HD sc_t wsum(sc_t v) { return v; }
AUGB_HD int wmaxi(int v) { return v; }
AUGB_HD int wmini(int v) { return v; }
AUGB_HD unsigned wmatch(int) { return 1u; }
AUGB_HD double wmaxd(double v) { return v; }
AUGB_HD double wsumd(double v) { return v; }
AUGB_HD unsigned wballot(bool p) { return p ? 1u : 0u; }
AUGB_HD int wbcast(int v, int) { return v; }
AUGB_HD sc_t wbcast64(sc_t v, int) { return v; }
#if defined(__CUDA_ARCH__)
AUGB_HD int wffs(unsigned b) { return __ffs(b) - 1; }
AUGB_HD int wpopc(unsigned b) { return __popc(b); }
#else
AUGB_HD int wffs(unsigned b) { return b ? __builtin_ctz(b) : -1; }
AUGB_HD int wpopc(unsigned b) { return __builtin_popcount(b); }
#endif
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
    if (gl >= AUGB_NLANES) return wargbest(score, key);
    const unsigned have = wballot(!isneg(score));
    const unsigned hg = have & gmask;
    /* no reduction unless some group holds more than one candidate */
    if (wballot((hg & (hg - 1u)) != 0) == 0) return hg ? wffs(hg) : -1;
    sc_t m = score;
    for (int o = gl >> 1; o; o >>= 1) { sc_t t = __shfl_xor_sync(tmask(), m, o); m = t > m ? t : m; }
    int k = score == m ? key : -0x7fffffff;
    for (int o = gl >> 1; o; o >>= 1) { int t = __shfl_xor_sync(tmask(), k, o); k = t > k ? t : k; }
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
        /* one exponential whichever side is larger (lanes of a warp disagree about that all the time): exp(-|lp - m|), the same argument
         * bit for bit as exp(m - lp) / exp(lp - m) */
        const double d = lp - m;
        const double e = exp(d > 0 ? -d : d);
        if (d > 0) { s = s * e + 1.0; m = lp; } else s += e;
    }
    AUGB_HD bool empty() const { return !(s > 0); }
    AUGB_HD double value() const { return s > 0 ? m + log(s) : -1e308; }
};
/* combine the per-lane accumulators of a group of gl lanes (aligned, power of two); every lane gets its group's result */
AUGB_D Lse glse(Lse a, int gl) {
#if AUGB_SIMT
    double M = a.m;
    for (int o = gl >> 1; o; o >>= 1) { double t = __shfl_xor_sync(tmask(), M, o); M = t > M ? t : M; }
    double s = a.s > 0 ? a.s * exp(a.m - M) : 0.0;
    for (int o = gl >> 1; o; o >>= 1) s += __shfl_xor_sync(tmask(), s, o);
    Lse r; r.m = M; r.s = s;
    return r;
#else
    return a;
#endif
}
AUGB_D double wbcastd(double v, int src) {
#if AUGB_SIMT
    return __shfl_sync(tmask(), v, tbase() + src);
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
