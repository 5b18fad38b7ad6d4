set -x
mkdir -p gpurun_out
export AUGB200_SWEEP=tasks
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tasks -s 1 -c 1 -o gpurun_out/prof_tasks_a python tools/prof_sweep.py 2368 2 > gpurun_out/ncu_tasks_a.log 2>&1
tail -2 gpurun_out/ncu_tasks_a.log
