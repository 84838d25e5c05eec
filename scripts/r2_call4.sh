#!/bin/bash
# round-2 GPU call 4: symmetry (both rules), checkpoint round trip, L2-prefetch experiment
mkdir -p gpurun_out
python scripts/sweep2.py "default pf0|KGX_MODE=stream|296,128" "default pf2|KGX_MODE=stream,KGX_STREAM_PF=2|296,128" "default pf4|KGX_MODE=stream,KGX_STREAM_PF=4|296,128" "default pf8|KGX_MODE=stream,KGX_STREAM_PF=8|296,128" "2^20 pf4|KGX_MODE=stream,KGX_STREAM_PF=4|64,128" > gpurun_out/sweep_pf.txt 2>&1; cat gpurun_out/sweep_pf.txt
timeout 1300 python -m pytest tests/test_gpu_symmetry.py tests/test_gpu_dropin.py::test_checkpoint_round_trip_through_reference_work_files -q -x --durations=12 > gpurun_out/pytest_gpu4.txt 2>&1; tail -45 gpurun_out/pytest_gpu4.txt; cat gpurun_out/symmetry_gain.txt
