set -x
mkdir -p gpurun_out
python -m pytest tests/test_utr.py -m gpu -x -q 2>&1 | tail -3
for b in 3; do AUGB200_SWEEP_BLOCKS_PER_SM=$b python tools/prof_sweep.py 3552 2 human_utr 50000 2>&1 | tail -1; done
python tools/prof_sweep.py 1000 2 human_utr 200000 2>&1 | tail -1
python tools/prof_sweep.py 4736 2 human 50000 2>&1 | tail -1
