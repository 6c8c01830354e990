#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_match_tc.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
for m in tc tc_f16x2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu-baseline --matcher $m > gpurun_out/m_bench_$m.json 2> gpurun_out/m_bench_$m.err
  timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"match_tc_pair|fallback_|rescore|prep_split" -c 12 --csv --log-file gpurun_out/m_ncu_$m.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline --matcher $m > /dev/null 2> gpurun_out/m_ncu_$m.err
done
python - <<'PY'
import json, csv
for m in ("tc","tc_f16x2"):
    try:
        d=json.loads(open("gpurun_out/m_bench_%s.json"%m).read().strip().splitlines()[-1])
        print(m, d["value"], d["ms_per_step"], d["stages_ms_per_step"].get("match_tc"), d["matcher"]["exact_rescans_last_step"], d["matcher"]["error_bound_violations_last_step"])
    except Exception as e:
        print(m, "failed", e)
    rows=[r for r in csv.reader(open("gpurun_out/m_ncu_%s.csv"%m)) if len(r)>10]
    hdr=rows[0]
    ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); ii=hdr.index("ID")
    d={}
    for r in rows[1:]:
        d.setdefault((r[ii],r[ik][:24]),{})[r[im][:30]]=r[iv]
    for k,v in list(d.items())[-6:]:
        print("  ",k,v)
PY
