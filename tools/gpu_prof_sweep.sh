# full ncu capture of the prep and sweep kernels (592 windows x 50 kb: 4 warps / SM, keeps the replay short)
set -x
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_prep|k_sweep" -s 2 -c 2 -o gpurun_out/prof_$TAG python tools/prof_sweep.py 592 2 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
