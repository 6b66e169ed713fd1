#!/usr/bin/env bash
# BASELINE config 4 (700-pair PIE-shaped sweep, etainv + ptp, sharded per image) and config 5 (nti + ptp replicas) on N GPUs
# of one box, wall-clock around the whole eval.py run (model load, graph capture, PNG writes included), then bench.py at N.
set -uo pipefail
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, cfg, extra args
  local t0=$(date +%s.%N)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 eval.py --cfg $2 --override ${@:3} > gpurun_out/r02_sweep_$1_n$N.log 2>&1
  local rc=$?
  local t1=$(date +%s.%N)
  echo "$1 on $N GPUs: rc=$rc wall $(python -c "print(round($t1 - $t0, 1))") s" | tee -a gpurun_out/r02_sweep_n$N.txt
  tail -2 gpurun_out/r02_sweep_$1_n$N.log | tee -a gpurun_out/r02_sweep_n$N.txt
}
run pie700 cfg/eval/synthetic_pie.yaml
run nti8 cfg/eval/synthetic_nti.yaml
ls result 2>/dev/null | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err; cut -c1-1500 gpurun_out/r02_bench_n$N.json
