#!/bin/bash
# final captures at HEAD: bench line, ncu launch list, ncu --set full of the stream kernel, the whole GPU test suite
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 400 gpurun_out/bench_r2f.json; tail -3 gpurun_out/bench_r2f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 3 -c 1 -o gpurun_out/r2f_stream python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_final.txt 2>&1; tail -16 gpurun_out/pytest_gpu_final.txt
