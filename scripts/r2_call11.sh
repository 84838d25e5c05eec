#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "
import __graft_entry__ as g, os
os.environ['KGX_MODE']='tmem'
g.smoke()" > gpurun_out/tmem_smoke.txt 2>&1; tail -5 gpurun_out/tmem_smoke.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_herd.py tests/test_gpu_modes.py -q -k tmem -x > gpurun_out/pytest_tmem.txt 2>&1; tail -15 gpurun_out/pytest_tmem.txt
python scripts/sweep2.py "default tmem|KGX_MODE=tmem|296,128" "default stream|KGX_MODE=stream|296,128" "2^20 tmem|KGX_MODE=tmem|64,128" "606k tmem (one wave)|KGX_MODE=tmem|37,128" "606k stream|KGX_MODE=stream|37,128" "512k tmem|KGX_MODE=tmem|32,128" "262k tmem|KGX_MODE=tmem|16,128" "262k resident|KGX_MODE=resident|16,128" "131k tmem|KGX_MODE=tmem|8,128" > gpurun_out/sweep_tmem.txt 2>&1; cat gpurun_out/sweep_tmem.txt
