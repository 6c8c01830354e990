timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_mosaic_launches.csv python scripts/bench_mosaic.py --rows 4 --cols 6 > gpurun_out/r02v_mosaic.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02v_mosaic_launches.csv | head -20
tail -3 gpurun_out/r02v_mosaic.log | cut -c1-600
