/* ghmm_tasks.h — host-side entry points of the task-engine sweep kernels (ghmm_tasks.cu) */
#pragma once
#include <cuda_runtime.h>
#include "ghmm_defs.h"
#include "ghmm_prep.h"

namespace augb {
constexpr int TASK_WARPS = 4;       /* warps (= windows in flight) per CTA */
cudaError_t tasks_upload_model(const DevModel* dm, cudaStream_t s);
/* variant: bit 0 = UTR model, bit 1 = forward values as well */
cudaError_t tasks_launch_sweep(int variant, const WinDev* wins, int nwin, int* next, int blocks, cudaStream_t s);
}
