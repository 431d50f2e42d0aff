#!/bin/bash
# Profile captures of round 2 (run under gpurun on ONE GPU; outputs under gpurun_out/, summaries are copied to profiles/).
#  1. launch list of the reference-graph arm (unmodified reference GeneralizedRCNN over this repo's layers, fused): shows the
#     conv_tc / conv_wgrad_tc / roi_align_fpn / nms kernels under the reference's module graph
#  2. ncu --set full of the dominant conv launch (3x3 256->256 on 2x200x336, fwd) and of a short 1x1 layer
#  3. ncu --set full of the fused FPN ROIAlign forward (in-step shape)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_refgraph_r2.csv \
    python bench.py --model reference --steps 1 --warmup 3 --no-roofline --no-ops --no-cpu-baseline > gpurun_out/launches_refgraph_r2.out 2>&1
python tools/summarize_launches.py gpurun_out/launches_refgraph_r2.csv 6 > gpurun_out/launches_refgraph_r2.md 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/ncu_conv3x3_r2 \
    python tools/bench_conv.py fwd 2 256 200 336 256 3 1 1 > gpurun_out/ncu_conv3x3_r2.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/ncu_conv1x1_r2 \
    python tools/bench_conv.py fwd 2 64 200 336 256 1 1 0 > gpurun_out/ncu_conv1x1_r2.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:roi_align_fpn_fwd -s 1 -c 1 -o gpurun_out/ncu_roialign_fpn_r2 \
    python tools/bench_ops.py > gpurun_out/ncu_roialign_fpn_r2.out 2>&1
ls -la gpurun_out/*.ncu-rep
