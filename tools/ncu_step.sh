#!/bin/bash
# per-kernel launch list (times, DRAM bytes, tensor-pipe activity) of one serving + one training step
#   tools/ncu_step.sh <cfg: c2|c2all|c3> <out prefix under gpurun_out/>
CFG=${1:-c2}; OUT=${2:-gpurun_out/ncu_step}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --csv --log-file ${OUT}.csv python tools/step_once.py $CFG 2 > ${OUT}.log 2>&1
python tools/roofline_table.py ${OUT}.csv 6545.6 > ${OUT}_table.txt 2>&1
