#!/bin/bash
# round-2 GPU call R: where does the MMA-issuing warp of conv_tc_kernel spend its time on the narrow layers?  One ncu capture with
# source counters of the six conv_tc launches of a frame (conv1, d64, d128, u64, u32, final)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_tc_kernel -c 6 -o gpurun_out/r02_conv_tc_src python tools/ncu_frame.py > gpurun_out/r_ncu.log 2>&1; tail -2 gpurun_out/r_ncu.log
ls -la gpurun_out/r02_conv_tc_src.ncu-rep
