set -x
nproc
mkdir -p gpurun_out
python tools/prof_sweep.py 592
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_r1a python tools/prof_sweep.py 592 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python tools/prof_sweep.py 592 > /dev/null 2>&1
tail -12 gpurun_out/launches_r1a.csv
