#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -k "check_gpu_vs_cpu or odd_grid or checkpoint or in56" > gpurun_out/pytest_gpu16.txt 2>&1; tail -5 gpurun_out/pytest_gpu16.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']; print('e2e', e['value'], 'upload_s', e['one_time_upload_s'], 'get_kangaroos_s', e['get_kangaroos_s'])"
timeout 300 python -m pytest tests/test_gpu_symmetry.py -q -k "reference_check_with_use_symmetry and resident" 2>&1 | tail -2
