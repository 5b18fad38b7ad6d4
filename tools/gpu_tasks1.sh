# experiment: task engine vs cooperative warp-per-window
set -x
mkdir -p gpurun_out
export AUGB200_SWEEP=tasks
timeout 900 python -m pytest tests/test_gpu.py tests/test_utr.py tests/test_nc.py tests/test_softmask.py tests/test_species.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/prof_sweep.py 4736 2 2>&1 | tail -1
timeout 600 python tools/prof_sweep.py 8000 2 2>&1 | tail -1
AUGB200_SWEEP_BLOCKS_PER_SM=3 timeout 600 python tools/prof_sweep.py 8000 2 2>&1 | tail -1
AUGB200_SWEEP_BLOCKS_PER_SM=2 timeout 600 python tools/prof_sweep.py 8000 2 2>&1 | tail -1
timeout 600 python tools/prof_sweep.py 1000 2 human_utr 200000 2>&1 | tail -1
unset AUGB200_SWEEP
timeout 600 python tools/prof_sweep.py 8000 2 2>&1 | tail -1
