/*
 * ghmm_tasks.cu — the sweep kernels in the task-engine flavour (sm_100a), compiled with -DAUGB_TASKS.
 *
 * One warp per window as in ghmm_kernels.cuh, but the lanes of the warp run the (state, column) tasks of the current column side
 * by side instead of cooperating on one state after the other (ghmm_sweep.h, "task engine"): a column that ends six exon states,
 * three lessD states and three longdss states costs the longest of those evaluations, not their sum, and the lane-0 bookkeeping
 * of a new cell (event log, candidate lists, chain entries: 28 % of the instructions of the cooperative kernel,
 * profiles/r1_sweep_final_by_function.txt) is done for all cells of the column at once.
 */
#ifndef AUGB_TASKS
#error "compile ghmm_tasks.cu with -DAUGB_TASKS"
#endif
#include <cuda_runtime.h>

#include "ghmm_defs.h"
#include "ghmm_prep.h"
#include "ghmm_seq.h"
#include "ghmm_sweep.h"
#include "ghmm_tasks.h"

namespace augb {

/* this translation unit's copy of the model: lanes index the tables with different states, constant memory would serialise them */
__device__ DevModel g_model_tasks;

template <class SW>
__global__ void __launch_bounds__(TASK_WARPS * 32) k_sweep_tasks(const WinDev* __restrict__ wins, int nwin, int* __restrict__ next) {
    const DevModel* m = &g_model_tasks;
    __shared__ WarpState wstate[TASK_WARPS];
    const int wid = threadIdx.x >> 5;
    for (;;) {
        int wi = 0;
        if ((threadIdx.x & 31) == 0) wi = atomicAdd(next, 1);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= nwin) break;
        const WinDev& wd = wins[wi];
        SW sw; sw.m = m; sw.ws = &wstate[wid];
        sw.w = make_view(wd.base, wd.lay, wd.L, 0);
        sw.run_tasks();
        __syncwarp();
    }
}

cudaError_t tasks_upload_model(const DevModel* dm, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(g_model_tasks, dm, sizeof(DevModel), 0, cudaMemcpyHostToDevice, s);
}

cudaError_t tasks_launch_sweep(int variant, const WinDev* wins, int nwin, int* next, int blocks, cudaStream_t s) {
    switch (variant & 3) {
    case 0: k_sweep_tasks<Sweep><<<blocks, TASK_WARPS * 32, 0, s>>>(wins, nwin, next); break;
    case 1: k_sweep_tasks<SweepUtr><<<blocks, TASK_WARPS * 32, 0, s>>>(wins, nwin, next); break;
    case 2: k_sweep_tasks<SweepFwd><<<blocks, TASK_WARPS * 32, 0, s>>>(wins, nwin, next); break;
    default: k_sweep_tasks<SweepFwdUtr><<<blocks, TASK_WARPS * 32, 0, s>>>(wins, nwin, next); break;
    }
    return cudaGetLastError();
}

}  // namespace augb
