#!/bin/bash
# One-call evidence capture for a round (run on the GPU box):
#   gpurun --timeout 1200 -- 'bash profiles/capture.sh r02'
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/<tag>/ afterwards
# (profiles/summarize_launches.py for the launch list, profiles/hot_lines.py for the source page).
# Numbers printed by runs under ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
KERNELS='integral|hessian|response_|rank_|validate|prefix_kernel|u8_to_f32|orient_|prep_split|match_tc|rescore|norm_max|fallback|ratio_insert|vote_|transpose|jpeg'
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 300 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on \
    -k regex:"match_tc_pair_kernel|orient_describe_warp|hessian_nms|rank_sort" -s 8 -c 6 -o $OUT/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
