#!/bin/bash
# round-2 GPU call 2: kernel variant sweep, pipe probes, bench line, parity tests
mkdir -p gpurun_out
python scripts/sweep2.py "default warp 2cta G128|KGX_MODE=stream|296,128" "default thread 2cta G128|KGX_MODE=stream,KGX_STREAM_INV=thread|296,128" \
  "default warp 3cta G86|KGX_MODE=stream,KGX_STREAM_CTAS=3|296,128" "default warp 3cta G88|KGX_MODE=stream,KGX_STREAM_CTAS=3,KGX_STREAM_G=88|296,128" \
  "default warp 2cta G64|KGX_MODE=stream,KGX_STREAM_G=64|296,128" "2^20 warp auto|KGX_MODE=stream|64,128" "2^20 thread auto|KGX_MODE=stream,KGX_STREAM_INV=thread|64,128" \
  "2^20 warp 3cta|KGX_MODE=stream,KGX_STREAM_CTAS=3|64,128" "2^20 warp G32|KGX_MODE=stream,KGX_STREAM_G=32|64,128" "512k warp|KGX_MODE=stream|32,128" \
  "512k warp 3cta|KGX_MODE=stream,KGX_STREAM_CTAS=3|32,128" "262k warp|KGX_MODE=stream|16,128" "262k warp 3cta|KGX_MODE=stream,KGX_STREAM_CTAS=3|16,128" \
  "262k resident|KGX_MODE=resident|16,128" "131k warp|KGX_MODE=stream|8,128" "131k warp 3cta|KGX_MODE=stream,KGX_STREAM_CTAS=3|8,128" \
  "131k warp G2|KGX_MODE=stream,KGX_STREAM_G=2|8,128" "131k resident|KGX_MODE=resident|8,128" "65k warp|KGX_MODE=stream|4,128" "65k resident|KGX_MODE=resident|4,128" \
  > gpurun_out/sweep2.txt 2>&1
cat gpurun_out/sweep2.txt
scripts/ubench3 2>&1 | head -32 > gpurun_out/ubench3b.txt; cat gpurun_out/ubench3b.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
timeout 1100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_solver.py tests/test_gpu_herd.py tests/test_gpu_dropin.py -x -q > gpurun_out/pytest_gpu2.txt 2>&1; tail -15 gpurun_out/pytest_gpu2.txt
