#!/bin/bash
# round-2 GPU call B: the swapped-role residual kernel (conv_res.cu): parity, timings vs the old path, timeline, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_parity_large.py tests/test_vr.py tests/test_gpu_refwarp.py -m gpu -q -x --durations=8 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
( for v in "" "FAV_NO_RES=1" "FAV_NO_NL=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/b_ablate.log 2>&1
timeout 300 python tools/trace_conv.py > gpurun_out/b_trace.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -c 300 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.err
