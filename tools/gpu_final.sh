# round-end style GPU check: parity suite, default bench line, launch list of a short bench run
set -x
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print({k: d[k] for k in ("value", "ms_per_step", "verified", "verified_e2e", "gpu_launches", "sweep_ms")}, d["e2e"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in d.get("secondary", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "seconds", "verified", "error", "kernel_ms", "sweep_ms")}, v.get("cpu_baseline", {}).get("value"))
print(json.dumps(d.get("secondary", {}).get("dropin_e2e", {}))[:1500])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --windows 3000 --no-cpu-baseline --no-secondary > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_$TAG.csv | cut -d, -f5,12- | tail -12
