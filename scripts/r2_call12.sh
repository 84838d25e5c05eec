#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field.py tests/test_gpu_parity.py -q -x -k "not full_size" > gpurun_out/pytest_chains.txt 2>&1; tail -5 gpurun_out/pytest_chains.txt
python scripts/sweep2.py "default stream|KGX_MODE=stream|296,128" "default stream again|KGX_MODE=stream|296,128" "default tmem|KGX_MODE=tmem|296,128" "262k resident|KGX_MODE=resident|16,128" "2^20 stream|KGX_MODE=stream|64,128" > gpurun_out/sweep_chains.txt 2>&1; cat gpurun_out/sweep_chains.txt
