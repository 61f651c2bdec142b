#!/bin/bash
# round-2 GPU call G: file pipeline, cfg3 at N=1, cfg4 VR, config sweep (cfg5 4K), 4-argument checker timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_front.py tests/test_gpu_net.py -m gpu -q -x -k "corners or video_driver or consistency" > gpurun_out/g_pytest.log 2>&1; tail -2 gpurun_out/g_pytest.log
timeout 900 python tools/file_pipeline_bench.py > gpurun_out/g_filepipe.log 2>&1; tail -1 gpurun_out/g_filepipe.log
timeout 900 python bench.py --config cfg3 --steps 24 > gpurun_out/g_cfg3_n1.json 2> gpurun_out/g_cfg3_n1.err; tail -c 900 gpurun_out/g_cfg3_n1.json; tail -2 gpurun_out/g_cfg3_n1.err
timeout 900 python tools/vr_bench.py > gpurun_out/g_vr.log 2>&1; tail -1 gpurun_out/g_vr.log
timeout 900 python tools/configs_sweep.py > gpurun_out/g_sweep.jsonl 2> gpurun_out/g_sweep.err; cat gpurun_out/g_sweep.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g_launches.csv python tools/ncu_frame.py > /dev/null 2>&1
grep -E "iir|avg_scan|norm_|transpose|structure|eigen" gpurun_out/g_launches.csv | awk -F, '{print $5, $NF}' | head -20
