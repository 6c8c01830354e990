OUT=gpurun_out; mkdir -p $OUT
timeout -s KILL 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02i_bench_n1.json 2> $OUT/r02i_bench_n1.err; tail -3 $OUT/r02i_bench_n1.err
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r02i_bench_n2.json 2> $OUT/r02i_bench_n2.err; tail -5 $OUT/r02i_bench_n2.err
python - <<EOF
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open("gpurun_out/r02i_bench_n%d.json"%n) if l.startswith("{")][-1])
        print("N",n,d["value"], d["ms_per_step"], d["e2e"]["value"])
        for k in ("c4","mosaic"): print(k, json.dumps(d.get(k))[:900])
    except Exception as e: print("N",n,"ERR",e)
EOF
