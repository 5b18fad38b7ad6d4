set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/prof_sweep.py 2368 3 2>&1 | tail -1
python tools/prof_sweep.py 4736 2 2>&1 | tail -1
AUGB200_LIB=$PWD/build_variants/v0.so python tools/prof_sweep.py 4736 2 2>&1 | tail -1
ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep -s 1 -c 1 --csv --log-file gpurun_out/icc_new.csv python tools/prof_sweep.py 2368 > /dev/null 2>&1
grep -v "^==" gpurun_out/icc_new.csv | cut -d, -f5,13- | tail -7
