set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/prof_sweep.py 2368 3 2>&1 | tail -1
python tools/prof_sweep.py 4736 2 2>&1 | tail -1
python tools/prof_sweep.py 1000 2 human_utr 200000 2>&1 | tail -1
