set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r1_n1.json
tail -3 gpurun_out/bench_err.log
cat gpurun_out/bench_r1_n1.json
