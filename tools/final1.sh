set -x
mkdir -p gpurun_out
timeout 200 python tools/prof_sweep.py 2368 3 2>&1 | tail -1
AUGB200_LIB=$PWD/build_variants/l32.so timeout 200 python tools/prof_sweep.py 2368 3 2>&1 | tail -1
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_final_err.log | tail -1 > gpurun_out/bench_r1_final.json
tail -3 gpurun_out/bench_final_err.log
cat gpurun_out/bench_r1_final.json
