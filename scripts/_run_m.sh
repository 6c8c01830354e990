OUT=gpurun_out; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > $OUT/r02m_gpu_tests.log 2>&1; tail -4 $OUT/r02m_gpu_tests.log
timeout -s KILL 400 python scripts/bench_sequence.py --rows 3 --cols 4 > $OUT/r02m_bench_sequence.json 2> $OUT/r02m_bench_sequence.err; tail -2 $OUT/r02m_bench_sequence.err
timeout -s KILL 300 python scripts/bench_sequence.py --rows 3 --cols 4 --options entropy=0 --cpu-pairs 0 >> $OUT/r02m_bench_sequence.json 2>> $OUT/r02m_bench_sequence.err
cat $OUT/r02m_bench_sequence.json | cut -c1-1500
