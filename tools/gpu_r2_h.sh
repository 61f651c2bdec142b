#!/bin/bash
# round-2 GPU call H: re-timed 4-argument checker (avg chain), native file pipeline with per-stage statistics, cfg3 data plane on
# three streams at N=1, fused temporal stage vs three kernels A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_front.py tests/test_gpu_net.py -m gpu -q -x -k "corners or video_driver or consistency" > gpurun_out/h_pytest.log 2>&1; tail -2 gpurun_out/h_pytest.log
timeout 900 python tools/file_pipeline_bench.py > gpurun_out/h_filepipe.log 2>&1; grep pipeline_stats gpurun_out/h_filepipe.log; tail -1 gpurun_out/h_filepipe.log
timeout 900 python bench.py --config cfg3 --steps 24 > gpurun_out/h_cfg3_n1.json 2> gpurun_out/h_cfg3_n1.err; tail -c 1500 gpurun_out/h_cfg3_n1.json; tail -2 gpurun_out/h_cfg3_n1.err
timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
FAV_NO_STAGE=1 timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/h_bench_nostage.json 2> gpurun_out/h_bench_nostage.err
python - <<'PY'
import json
for f in ("h_bench", "h_bench_nostage"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("value_full"), d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/h_launches.csv python tools/ncu_frame.py > /dev/null 2>&1
grep -E "iir|avg_scan|norm_|transpose|structure|eigen" gpurun_out/h_launches.csv | awk -F, '{print $5, $NF}' | head -20
