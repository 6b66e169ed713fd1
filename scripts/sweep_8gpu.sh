#!/usr/bin/env bash
# BASELINE config 4 (700-pair PIE-shaped sweep, etainv + ptp, sharded per image over N GPUs of one box) through eval.py, then
# bench.py at N.  eval.py prints the sweep's own wall time / throughput (graph capture, VAE / CLIP, PNG writes included).
set -uo pipefail
N=${1:-8}
PIPES=${2:-3}
mkdir -p gpurun_out
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 eval.py \
    --cfg cfg/eval/synthetic_pie.yaml --override --pipes $PIPES > gpurun_out/r02_sweep_pie700_n$N.log 2>&1
echo "pie700 on $N GPUs, $PIPES groups in flight per GPU: rc=$? wall incl. process start and model load $(( $(date +%s) - t0 )) s" | tee gpurun_out/r02_sweep_n$N.txt
grep "^combo" gpurun_out/r02_sweep_pie700_n$N.log | tee -a gpurun_out/r02_sweep_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.out 2> gpurun_out/r02_bench_n$N.err
grep '^{"metric"' gpurun_out/r02_bench_n$N.out > gpurun_out/r02_bench_n$N.json
tail -2 gpurun_out/r02_bench_n$N.err; cut -c1-1800 gpurun_out/r02_bench_n$N.json
