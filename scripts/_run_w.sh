timeout -s KILL 600 python -m pytest tests/test_gpu_blend.py tests/test_gpu_zz_bands.py tests/test_gpu_tiles.py tests/test_gpu_zz_colour_stack.py tests/test_gpu_stitcher.py -q -m gpu -x 2>&1 | tail -2
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_mosaic_launches.csv python scripts/bench_mosaic.py --rows 4 --cols 6 > gpurun_out/r02w_mosaic.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02w_mosaic_launches.csv | head -9
timeout -s KILL 300 python scripts/bench_mosaic.py --rows 4 --cols 6 2>/dev/null | tail -1 | cut -c1-900
