#!/bin/bash
# A/B of the round-2 instruction trims of the stream kernel (uint4 jump table, by-value DP emit arguments, two-instruction L2 prefetch)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_symmetry.py -m gpu -q -x -k "not full_size and not gain" 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2h.json') if l.startswith('{')][-1]); print('NEW', d['value'], d['kernel_only']['value'], d['e2e']['value'], d['roofline']['frac'])"
