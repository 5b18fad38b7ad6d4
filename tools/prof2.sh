set -x
mkdir -p gpurun_out
TAG=${1:-x}; NW=${2:-2368}
python tools/prof_sweep.py 592 3
python tools/prof_sweep.py 2368 3
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_$TAG python tools/prof_sweep.py $NW > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
