set -x
mkdir -p gpurun_out
TAG=${1:-r1f}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_$TAG python tools/prof_sweep.py 2368 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_sweep.py 2368 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_$TAG.csv | tail -9 | cut -c1-300
