#!/bin/bash
# ncu evidence for the bench command (run under gpurun, after the same command exited 0 without ncu).
# Usage: bash tools/gpu_profile.sh <tag> [envs]
set -u
TAG=${1:-r1}
ENVS=${2:-262144}
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-others --envs ${ENVS}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ds_step_kernel -s 10 -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
$CMD > gpurun_out/plain3_${TAG}.log 2>&1 &&
ncu --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:ds_step_kernel -s 10 -c 2 --csv --log-file gpurun_out/flops_${TAG}.csv $CMD > gpurun_out/ncu_flops_${TAG}.log 2>&1
echo "flop count rc=$?"
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out | tail -15
# the single-vehicle-env workloads (quad_k8 / traj_quad / hexa_circle): DRAM traffic, issue activity and residency per launch
CMD2="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity --envs ${ENVS}"
$CMD2 > gpurun_out/plain4_${TAG}.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed.sum \
    --clock-control none -k regex:ds_step_kernel --csv --log-file gpurun_out/others_${TAG}.csv $CMD2 > gpurun_out/ncu_others_${TAG}.log 2>&1
echo "others rc=$?"
