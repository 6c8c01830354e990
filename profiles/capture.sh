#!/bin/bash
# One-call evidence capture for a round (run on the GPU box):
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh r02'
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/<tag>/ afterwards
# (profiles/summarize_launches.py for the launch list, profiles/hot_lines.py for the source page).
# Numbers printed by runs under ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
KERNELS='integral|hessian|response_|rank_|bin_|validate|prefix_kernel|u8_to_f32|orient_|prep_split|match_tc|rescore|norm_max|fallback|ratio_insert|vote_|transpose|jpeg'
# 0. the whole GPU suite (includes the band-mosaic and wrap-aware phase tests that were written without a GPU)
timeout -s KILL 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_gpu_tests.log 2>&1
# 1. opt-in kernel variants vs the default schedule (identical results required), then the default GPU suite stays as it was
VFSMS_EXPERIMENTAL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_variants.py -x -q > $OUT/${TAG}_variants_tests.log 2>&1
# 2. the bench as the driver runs it (autotune probe picks the validated variants), the default schedule, and each variant alone
timeout -s KILL 500 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 --no-autotune --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
for V in describe=2 describe=3 describe=5 describe=6 describe=7 describe=8 sort=1 lpt=1; do
    timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt $V > $OUT/${TAG}_bench_${V/=/}.json 2> $OUT/${TAG}_bench_${V/=/}.err
done
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout -s KILL 300 python scripts/bench_mosaic.py --rows 4 --cols 6 > $OUT/${TAG}_bench_mosaic.json 2> $OUT/${TAG}_bench_mosaic.err
# 3. ncu on the configuration the bench selected (explicit --opt, no probe subprocess under the profiler)
OPTS=$(python - <<PY
import json
try:
    v = json.load(open("$OUT/${TAG}_bench.json"))["config"]["kernel_variants"]
    d = {"describe": 1, "sort": 0, "lpt": 0, "entropy": 0}
    print(" ".join("--opt %s=%d" % (k, x) for k, x in v.items() if d.get(k) != x))
except Exception:
    print("")
PY
)
echo "profiling with: --no-autotune $OPTS" > $OUT/${TAG}_ncu_config.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 300 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $OPTS > /dev/null 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on \
    -k regex:"match_tc_pair_kernel|orient_describe_warp|hessian_nms|rank_sort|bin_rank" -s 8 -c 6 -o $OUT/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-autotune $OPTS > $OUT/${TAG}_ncu_full.log 2>&1
# 4. output encode (jpeg_enc.cu, written without GPU access): measurement + one full capture of its kernels
timeout -s KILL 300 python scripts/bench_encode.py > $OUT/${TAG}_bench_encode.json 2> $OUT/${TAG}_bench_encode.err
timeout -s KILL 300 python scripts/bench_encode.py --gray >> $OUT/${TAG}_bench_encode.json 2>> $OUT/${TAG}_bench_encode.err
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"jpeg_fdct_quant|jpeg_block_bits|jpeg_emit|jpeg_stuff|scan_" \
    -s 9 -c 9 -o $OUT/${TAG}_encode_full python scripts/bench_encode.py --rows 4096 --cols 4096 --reps 1 > $OUT/${TAG}_ncu_encode.log 2>&1
# 5. the whole reference workflow (files in -> file out) with Main.py's settings: host entropy stage, then device entropy
timeout -s KILL 400 python scripts/bench_sequence.py --rows 3 --cols 4 > $OUT/${TAG}_bench_sequence.json 2> $OUT/${TAG}_bench_sequence.err
timeout -s KILL 300 python scripts/bench_sequence.py --rows 3 --cols 4 --options entropy=1 --cpu-pairs 0 >> $OUT/${TAG}_bench_sequence.json 2>> $OUT/${TAG}_bench_sequence.err
ls -la $OUT | tail -14
