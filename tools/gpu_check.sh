# one-stop GPU check: parity tests, sweep timing, instruction count of the sweep kernel (one-metric ncu pass)
set -x
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/prof_sweep.py 8000 2 2>&1 | tail -1
timeout 600 python tools/prof_sweep.py 4736 2 2>&1 | tail -1
timeout 600 python tools/prof_sweep.py 1000 2 human_utr 200000 2>&1 | tail -1
timeout 900 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_sweep -s 1 -c 1 --csv --log-file gpurun_out/inst_$TAG.csv python tools/prof_sweep.py 592 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/inst_$TAG.csv | cut -d, -f5,13- | tail -6
