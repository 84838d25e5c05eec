#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modes.py::test_stream_and_resident_agree_over_a_long_walk tests/test_gpu_symmetry.py::test_symmetric_device_hash_convert -q > gpurun_out/pytest_gpu15.txt 2>&1; tail -8 gpurun_out/pytest_gpu15.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']; print('e2e', e['value'], 'upload_s', e['one_time_upload_s'], 'get_kangaroos_s', e['get_kangaroos_s'])"
