/* ghmm_lockstep.h — host-side entry points of the lockstep sweep (ghmm_lockstep.cu) */
#pragma once
#include <cuda_runtime.h>
#include "ghmm_defs.h"
#include "ghmm_prep.h"

namespace augb {
cudaError_t lockstep_upload_model(const DevModel* dm, cudaStream_t s);
cudaError_t lockstep_launch_sweep(const WinDev* wins, int nwin, int* next, int blocks, cudaStream_t s);
}
