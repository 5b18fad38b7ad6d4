set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/prof_sweep.py 592 3 2>&1 | tail -1
python tools/prof_sweep.py 2368 3 2>&1 | tail -1
python tools/prof_sweep.py 4736 3 2>&1 | tail -1
python bench.py --windows 2368 --steps 2 --warmup 3 --no-secondary 2>&1 | tail -2
