mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python tools/prof_sampling.py 1000 100 2>&1 | tail -1
timeout 120 python tools/prof_sampling.py 1000 2 2>&1 | tail -1
timeout 120 python tools/prof_sweep.py 2368 2 2>&1 | tail -1
