#!/bin/bash
# Round-2 evidence capture (run on the GPU box):  gpurun --timeout 1500 -- 'bash profiles/capture_r02.sh <tag> [steps...]'
# steps: tests bench launches full ref   (default: all).  Output under gpurun_out/<tag>_*.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02}; shift || true
STEPS=${*:-tests bench launches full ref}
OUT=gpurun_out
mkdir -p $OUT
KERNELS='integral|hessian|response_|rank_|bin_|validate|prefix_kernel|u8_to_f32|orient_|describe|prep_split|match_tc|rescore|norm_max|fallback|ratio_insert|vote_|transpose'
FULLK=${FULLK:-'match_tc_pair_kernel|orient_describe_warp|describe_|hessian_nms|rescore_kernel|prep_split'}
for S in $STEPS; do
case $S in
tests)  timeout -s KILL 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_gpu_tests.log 2>&1; tail -5 $OUT/${TAG}_gpu_tests.log ;;
bench)  timeout -s KILL 500 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; head -c 600 $OUT/${TAG}_bench.json; echo ;;
ref)    timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ;;
launches) timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 400 --csv \
            --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
          python profiles/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt 2>&1; head -30 $OUT/${TAG}_launches.txt ;;
full)   timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$FULLK" -s ${FULLSKIP:-0} -c ${FULLCOUNT:-12} \
            -f -o $OUT/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_full.log 2>&1; tail -3 $OUT/${TAG}_ncu_full.log ;;
esac
done
ls -la $OUT | tail -12
