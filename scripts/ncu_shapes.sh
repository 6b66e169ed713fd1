#!/usr/bin/env bash
# ncu evidence for the production shapes of the final kernels (run under gpurun, ONE GPU):
#   * tensor-pipe utilisation + DRAM bytes of gemm_tc_k on the five shapes VERDICT r01 names, attn_d40_k, cross_attn_tc_k
#   * DRAM bytes of the norm kernels
# Output: gpurun_out/r02_ncu_shapes.csv (raw page, one row per launch)
set -uo pipefail
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 600 ncu --metrics $METRICS --clock-control none -k regex:"gemm_tc_k|attn_d40_k|attn_tc2_k|cross_attn_tc_k|gn_|layernorm_k|cfg_ddim_step_k" \
    --csv --log-file gpurun_out/r02_ncu_shapes.csv python scripts/bench_ops.py gemm conv attn gn ln sched --iters 1 > gpurun_out/r02_ncu_shapes.out 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02_ncu_shapes.out
