#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/pytest_gpu_full.txt 2>&1; tail -45 gpurun_out/pytest_gpu_full.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; cat gpurun_out/smoke.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -6 gpurun_out/sanitizer_memcheck.txt
