#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_symmetry.py -m gpu -q -x -k "not full_size and not gain" 2>&1 | tail -4
for i in 1 2; do python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('NEW', d['value'], d['kernel_only']['value'], d['e2e']['value'], d['roofline']['frac'])"; done
KGX_LIB_OVERRIDE=/root/repo/build/libkgx_before.so python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('OLD', d['value'], d['kernel_only']['value'])"
