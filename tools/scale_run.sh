#!/bin/bash
# Multi-GPU run (one box, up to N GPUs): bash tools/scale_run.sh N  -> gpurun_out/scale_N.log
# bench.py's C4 line at 1, 2, 4 .. N ranks (the full default line, C5 / throughput / C2 sections included, at 1 and N),
# the cross-route NCCL check, C3 over all GPUs, the 2+ GPU tests.
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/scale_$N.log
: > $L
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}" 2>&1 | grep -E '^\{|bench.py:|Error' | tail -1 >> $L; }
tr $N 29608 bench.py --gpus $N --steps 5 --warmup 3
tr 1 29601 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline
for n in 2 4; do
  [ $n -lt $N ] && tr $n 2960$n bench.py --gpus $n --steps 5 --warmup 3 --no-extras
done
tr $N 29611 tools/multi_gpu_check.py --members 65536 --years 2
tr $N 29612 bench.py --workload c3 --gpus $N --steps 2 --years 10
python -m pytest tests/test_gpu_team.py "tests/test_gpu_dropin.py::test_many_member_launches_on_all_gpus_equal_one_gpu" -m gpu -q 2>&1 | tail -2 >> $L
cat $L | cut -c1-1200
