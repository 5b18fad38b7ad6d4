# one-stop GPU check: parity tests, ncu launch list + full capture of the sweep kernel, bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TAG=${1:-r1}
NW=${2:-592}
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_$TAG python tools/prof_sweep.py $NW > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_sweep.py $NW > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_$TAG.csv | tail -5 | cut -c1-400
python bench.py --windows ${3:-10000} --steps 2 --warmup 3 2>&1 | tail -3
