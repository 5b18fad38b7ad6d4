set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python tools/prof_sweep.py 2368 3 2>&1 | tail -1
timeout 300 python bench.py --windows 4736 --steps 2 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'sweep_ms',d['sweep_ms'],'ms_per_step',d['ms_per_step'])"
timeout 200 python tools/prof_sampling.py 1000 100 2>&1 | tail -1
timeout 200 python tools/prof_sampling.py 1000 2 2>&1 | tail -1
