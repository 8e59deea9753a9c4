#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
TAG=${1:-s2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pointnet_fwd_tc -s 4 -c 2 -f -o gpurun_out/prof_pnfwd_${TAG} \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pointnet_bwd|match_topk|gemm_tf32x3|pointnet_fwd_simt|gat_aggregate_kernel|project_fuse_fwd' -c 14 -f -o gpurun_out/prof_misc_${TAG} \
    python tools/train_profile.py 1 >> gpurun_out/ncu_full_${TAG}.log 2>&1
python profiles/trace_pointnet.py > gpurun_out/trace_${TAG}.txt 2>&1
