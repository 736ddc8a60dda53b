#!/bin/bash
# Round profile of one training step (+ one decode call) on the GPU box.  Run AFTER the plain command exited 0:
#   bash tools/ncu_step.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (per-launch durations), gpurun_out/<tag>_counters.csv (DRAM bytes, tensor-pipe
# and memory-system utilisation per launch) and gpurun_out/<tag>_gemm_full.ncu-rep (--set full of 3 tap-GEMM launches).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_step.py > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py > $OUT/${TAG}_ncu1.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__cycles_elapsed.avg.per_second
ncu --profile-from-start off --metrics $M --clock-control none --csv --page raw \
    --log-file $OUT/${TAG}_counters.csv python tools/profile_step.py > $OUT/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tapgemm -s 20 -c 3 \
    -f -o $OUT/${TAG}_gemm_full python tools/profile_step.py > $OUT/${TAG}_ncu3.log 2>&1
ls -la $OUT | tail -8
# the CTC loss+grad kernel alone (BASELINE metric 2): full set, one launch at batch 128
ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 3 -c 1 \
    -f -o $OUT/${TAG}_ctc_full python tools/ctc_bench.py > $OUT/${TAG}_ncu4.log 2>&1
ls -la $OUT | tail -4
