#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surf.py tests/test_gpu_fullsize.py tests/test_gpu_variants.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 12 --warmup 4 --no-extra --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -3 gpurun_out/h_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/h_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages_ms_per_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hessian" -c 8 --csv --log-file gpurun_out/h_ncu.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1
grep hessian gpurun_out/h_ncu.csv | tail -4 | awk -F'","' '{print substr($5,1,40), $NF}'
