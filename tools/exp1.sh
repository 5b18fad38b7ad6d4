set -x
mkdir -p gpurun_out
run() { AUGB200_LIB=$PWD/build_variants/$1.so AUGB200_SWEEP_BLOCKS_PER_SM=$2 python tools/prof_sweep.py ${3:-4736} 2 2>&1 | tail -1; }
AUGB200_LIB=$PWD/build_variants/vA.so python -m pytest tests/test_gpu.py -m gpu -x -q 2>&1 | tail -2
for b in 1 2 3 4; do run v0 $b; done
for b in 2 3 4; do run vA $b; done
for b in 2 3; do run vB $b; done
for b in 4 8 6; do run vC $b; done
for b in 4 5 6; do run vE $b; done
AUGB200_LIB=$PWD/build_variants/vA.so ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_vA python tools/prof_sweep.py 2368 > gpurun_out/ncu_vA.log 2>&1
tail -2 gpurun_out/ncu_vA.log
