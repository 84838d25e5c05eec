#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_batchinv.py > gpurun_out/batchinv.log 2>&1; cat gpurun_out/batchinv_microbench.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps(d['e2e']))"
