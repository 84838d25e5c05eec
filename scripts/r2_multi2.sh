#!/bin/bash
# solver only: bash scripts/r2_multi2.sh <N> <window-file> <dp>
N=${1:-8}; CFG=${2:-tests/golden/puzzle110_window80.txt}; DP=${3:-16}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR -m kangaroo_b200.solver $CFG --dp $DP --seed 12 --max-steps 6000 > gpurun_out/solver_${N}gpu_v2.txt 2>&1; tail -22 gpurun_out/solver_${N}gpu_v2.txt
