#!/bin/bash
# round-2 GPU call L: what bounds in_apply?  ncu with the caches left alone (--cache-control none: the replays of a kernel find
# its working set in L2, the situation inside a frame) beside the default cold-cache capture
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --cache-control none --profile-from-start off -k regex:"in_apply" -c 11 -o /tmp/ncu/apply_warm python tools/ncu_frame.py > gpurun_out/l_ncu1.log 2>&1; tail -1 gpurun_out/l_ncu1.log
ncu -i /tmp/ncu/apply_warm.ncu-rep --page raw --csv > gpurun_out/r02_apply_warm_raw.csv 2>/dev/null
ncu -i /tmp/ncu/apply_warm.ncu-rep --page details > gpurun_out/r02_apply_warm_details.txt 2>/dev/null
gzip -9 -f gpurun_out/r02_apply_warm_raw.csv gpurun_out/r02_apply_warm_details.txt
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "layer or parity or session" > gpurun_out/l_pytest.log 2>&1; tail -2 gpurun_out/l_pytest.log
ls -la gpurun_out | tail -5
