#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_inv.txt
for u in 1 2 3; do
  echo "== KGX_INV_UNROLL=$u" >> gpurun_out/sweep_inv.txt
  KGX_LIB_OVERRIDE=$PWD/build/exp/libkgx_u$u.so python scripts/sweep2.py "default|KGX_MODE=stream|296,128" "2^20|KGX_MODE=stream|64,128" "512k|KGX_MODE=stream|32,128" "262k resident|KGX_MODE=resident|16,128" "131k resident|KGX_MODE=resident|8,128" "131k stream|KGX_MODE=stream|8,128" >> gpurun_out/sweep_inv.txt 2>&1
done
echo "== symmetric mode (HEAD lib)" >> gpurun_out/sweep_inv.txt
python scripts/sweep2.py "default plain|KGX_MODE=stream|296,128" "default symclass|KGX_MODE=stream,SYM=symclass|296,128" "default lastjump|KGX_MODE=stream,SYM=lastjump|296,128" "262k resident symclass|KGX_MODE=resident,SYM=symclass|16,128" >> gpurun_out/sweep_inv.txt 2>&1
cat gpurun_out/sweep_inv.txt
