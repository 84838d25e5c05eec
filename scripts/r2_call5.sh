#!/bin/bash
mkdir -p gpurun_out
timeout 1300 python -m pytest tests/test_gpu_symmetry.py::test_symmetry_gain_in_operations_per_key tests/test_gpu_dropin.py::test_checkpoint_round_trip_through_reference_work_files -q -s --durations=5 > gpurun_out/pytest_gpu5.txt 2>&1; tail -30 gpurun_out/pytest_gpu5.txt; cat gpurun_out/symmetry_gain.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 600 gpurun_out/bench_r2c.json; tail -3 gpurun_out/bench_r2c.err
