#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surf.py tests/test_gpu_match_tc.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -3 gpurun_out/e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["stages_ms_per_step"])
PY
