# source-level profile of the fused forward fill + sampling kernel on a few chr2L windows
set -x
mkdir -p gpurun_out
timeout 1700 ncu --section SourceCounters --section SpeedOfLight --import-source on --clock-control none -k regex:k_sweep_sample_utr -s 1 -c 1 -o gpurun_out/prof_sample_r2 python tools/prof_chr2l.py 8 100 > gpurun_out/ncu_sample_r2.log 2>&1
tail -3 gpurun_out/ncu_sample_r2.log
