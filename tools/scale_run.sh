#!/bin/bash
# Multi-GPU run (one box, up to N GPUs): bash tools/scale_run.sh N  -> gpurun_out/scale_N.log
# bench.py's C4 line (+ c5 / throughput sections) at 1, 2, 4 .. N ranks, the cross-route NCCL check, the 2+ GPU tests.
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/scale_$N.log
: > $L
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}" 2>&1 | grep -E '^\{|bench.py:' | tail -1 >> $L; }
for n in 1 2 4 8; do
  [ $n -le $N ] && tr $n 2960$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline
done
tr $N 29611 tools/multi_gpu_check.py --members 65536 --years 2
tr $N 29612 bench.py --workload c3 --gpus $N --steps 2 --years 10
python -m pytest tests/test_gpu_team.py "tests/test_gpu_dropin.py::test_many_member_launches_on_all_gpus_equal_one_gpu" -m gpu -q 2>&1 | tail -2 >> $L
cat $L | cut -c1-900
