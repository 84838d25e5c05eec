#!/bin/bash
# round-2 multi-GPU call: bash scripts/r2_multi.sh <N> <window-file> <dp>
N=${1:-2}; CFG=${2:-tests/golden/puzzle110_window72.txt}; DP=${3:-14}; SAVE=${4:-save}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 1500 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
NCCL_DEBUG=WARN $TR -m kangaroo_b200.solver $CFG --dp $DP --seed 11 --max-steps 6000 $( [ "$SAVE" = save ] && echo --save-work gpurun_out/solver_${N}gpu.work ) > gpurun_out/solver_${N}gpu.txt 2>&1; tail -25 gpurun_out/solver_${N}gpu.txt
ls -la gpurun_out/solver_${N}gpu.work 2>/dev/null && oracle/_ref/kangaroo_ref_cpu -winfo gpurun_out/solver_${N}gpu.work | tail -12 && oracle/_ref/kangaroo_ref_cpu -wcheck gpurun_out/solver_${N}gpu.work | tail -4; rm -f gpurun_out/solver_${N}gpu.work
