#!/bin/bash
# round-2 GPU call 3: symmetry tests, checkpoint round trip, bench line, ncu launch list + full capture of the stream kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_symmetry.py tests/test_gpu_dropin.py::test_checkpoint_round_trip_through_reference_work_files -q -s > gpurun_out/pytest_gpu3.txt 2>&1; tail -40 gpurun_out/pytest_gpu3.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 2500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 3 -c 1 -o gpurun_out/r2b_stream python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
