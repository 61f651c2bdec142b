#!/bin/bash
# round-2 GPU call K: streaming in_apply (bulk-copy ring) vs the per-thread-load version; file pipeline with GPU-side conversions
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_vr.py -m gpu -q -x --durations=3 > gpurun_out/k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/k_pytest.log
tail -4 gpurun_out/k_pytest.log
( for v in "" "FAV_APPLY_OLD=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/k_ablate.log 2>&1; cat gpurun_out/k_ablate.log | cut -c1-900
timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
FAV_APPLY_OLD=1 timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/k_bench_old.json 2> gpurun_out/k_bench_old.err
python - <<'PY'
import json
for f in ("k_bench", "k_bench_old"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("value_full"), d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 python tools/file_pipeline_bench.py > gpurun_out/k_filepipe.log 2>&1; grep pipeline_stats gpurun_out/k_filepipe.log; tail -1 gpurun_out/k_filepipe.log
