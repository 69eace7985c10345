#!/bin/bash
# Multi-GPU run (one box, N GPUs): bash tools/scale_run.sh N [quick]  -> gpurun_out/scale_N.log
# quick = skip the two C2 weak-scaling benches (the driver runs those itself at round end)
N=${1:-8}
QUICK=${2:-}
mkdir -p gpurun_out
L=gpurun_out/scale_$N.log
: > $L
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2>&1 | grep -E '^\{|bench.py:' | tail -1 >> $L; }
[ -z "$QUICK" ] && tr 29601 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e
tr 29602 tools/multi_gpu_check.py --members 65536 --years 2
tr 29603 bench.py --workload c4 --gpus $N --steps 3 --years 10 --verify
tr 29607 bench.py --workload c4 --gpus $N --steps 2 --years 10 --pipeline --verify
tr 29604 bench.py --workload c5 --gpus $N --steps 3 --years 10
tr 29605 bench.py --workload c3 --gpus $N --steps 2 --years 10
[ -z "$QUICK" ] && tr 29606 bench.py --gpus $N --steps 3 --warmup 3
cat $L | cut -c1-760
