#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_final2.txt 2>&1; tail -12 gpurun_out/pytest_gpu_final2.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2g.json') if l.startswith('{')][-1]); print(d['value'], d['kernel_only']['value'], d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'])"
