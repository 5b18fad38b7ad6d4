/*
 * ghmm_lockstep.cu — EXPERIMENT: lanes = windows, kept in step per (column, state kind).
 *
 * Every thread decodes its own window with the one-lane form of the routines (AUGB_SCALAR_DEVICE), but the 32 windows of a warp walk
 * the columns together and take the state kinds of a column in the same order, with a warp barrier after each kind: lanes whose
 * column has the kind run the routine side by side (same instructions, different windows), the others wait.
 */
#ifndef AUGB_SCALAR_DEVICE
#error "compile ghmm_lockstep.cu with -DAUGB_SCALAR_DEVICE"
#endif
#include <cuda_runtime.h>

#include "ghmm_defs.h"
#include "ghmm_prep.h"
#include "ghmm_seq.h"
#include "ghmm_sweep.h"
#include "ghmm_lockstep.h"

namespace augb {

__device__ DevModel g_model_ls;

#define LS_STEP(cond, call) { const bool c__ = act && (cond); if (__any_sync(0xffffffffu, c__)) { if (c__) { call; } __syncwarp(); } }

__global__ void __launch_bounds__(32) k_sweep_lockstep(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next) {
    const DevModel* m = &g_model_ls;
    WarpState ws;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(next, 32);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= nwin) break;
        const int wi = g + lane;
        const bool have = wi < nwin;
        Sweep sw; sw.m = m; sw.ws = &ws;
        bool ok = false;
        if (have) { const WinDev& wd = wins[wi]; sw.w = make_view(wd.base, wd.lay, wd.L, 0); sw.attach(); sw.lane = 0; ok = sw.init_window(); }
        const int L = ok ? sw.L : 0;
        int Lmax = L;
        for (int o = 16; o; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, Lmax, o); Lmax = t > Lmax ? t : Lmax; }
        for (int j0 = 1; j0 < Lmax; j0 += 32) {
            /* this lane's masks of the chunk and the equalD columns that fall into it (as Sweep::run) */
            unsigned maskbuf[32]; unsigned actbits = 0;
            for (int t = 0; t < 32; t++) { const int j = j0 + t; maskbuf[t] = j < L ? sw.w.mask[j] : 0u; if (maskbuf[t]) actbits |= 1u << t; }
            unsigned eqcol[6]; unsigned eqany = 0;
            for (int q = 0; q < 6; q++) {
                eqcol[q] = 0;
                if (!ok || m->r_equald[q / 3][q % 3] < 0) continue;
                const int list = (q / 3 ? CL_RA : CL_LD) + q % 3; const int cur = ws.eq_cur[q], n = ws.cl_n[list];
                for (int i = cur; i < n; i++) {
                    const int due = sw.w.cl(list)[i].col + m->dStateLen;
                    if (due >= j0 + 32) break;
                    if (due >= j0 && due < L) { eqcol[q] |= 1u << (due - j0); actbits |= 1u << (due - j0); eqany |= 1u << (due - j0); }
                }
            }
            unsigned any_act = actbits;
            for (int o = 16; o; o >>= 1) any_act |= __shfl_xor_sync(0xffffffffu, any_act, o);
            while (any_act) {
                const int t = __ffs(any_act) - 1; any_act &= any_act - 1;
                const int j = j0 + t;
                const bool act = (actbits >> t & 1) != 0 && ws.status == 0;
                const unsigned mb = act ? maskbuf[t] : 0u;
                unsigned eqbits = 0;
                if (act && (eqany >> t & 1)) for (int q = 0; q < 6; q++) if (eqcol[q] >> t & 1) eqbits |= 1u << q;
                if (act) { sw.fill_evstart(j); sw.set_class(sw.w.gc[j]); if (mb & MB_SLOW) sw.snip_column_begin(j); }
                __syncwarp();
                LS_STEP(mb & MB_LESSD, sw.lessd_eval(0, j));
                LS_STEP(mb & MB_RLESSD, sw.lessd_eval(1, j));
                for (int q = 0; q < 6; q++) LS_STEP(eqbits >> q & 1, sw.equald_eval(q >= 3, q >= 3 ? q - 3 : q, j));
                LS_STEP(mb & MB_LONGDSS, sw.fixed_eval(K_LONGDSS, 0, j));
                LS_STEP(mb & MB_RLONGDSS, sw.fixed_eval(K_LONGDSS, 1, j));
                LS_STEP(mb & MB_LONGASS, sw.fixed_eval(K_LONGASS, 0, j));
                LS_STEP(mb & MB_RLONGASS, sw.fixed_eval(K_LONGASS, 1, j));
                const unsigned slots = ((mb & MB_XSTOP) ? 0x3u : 0u) | ((mb & MB_XDSS) ? 0xfcu : 0u) | ((mb & MB_XRSTART) ? 0x300u : 0u) | ((mb & MB_XRASS) ? 0xfc00u : 0u);
                for (int q = 0; q < 16; q++) LS_STEP((slots >> q & 1) && m->xslot[q] >= 0, sw.exon_eval(m->xslot[q], j));
                if (act) { sw.commit_cells(j); sw.apply_pending(j); }
                __syncwarp();
            }
        }
        if (ok) {
            sw.fill_evstart(L);
            *sw.w.out_n_ev = ws.n_ev; *sw.w.out_status = ws.status;
            for (int i = 0; i < NCHAIN; i++) { sw.w.out_ncp[i] = ws.cp_n[i]; sw.w.out_nfcp[i] = 0; }
        }
        __syncwarp();
    }
}

cudaError_t lockstep_upload_model(const DevModel* dm, cudaStream_t s) { return cudaMemcpyToSymbolAsync(g_model_ls, dm, sizeof(DevModel), 0, cudaMemcpyHostToDevice, s); }
cudaError_t lockstep_launch_sweep(const WinDev* wins, int nwin, int* next, int blocks, cudaStream_t s) {
    k_sweep_lockstep<<<blocks, 32, 0, s>>>(wins, nwin, next);
    return cudaGetLastError();
}

}  // namespace augb
