#!/bin/bash
# round-2 GPU call I: packed satfinite split + rsqrt finalisation (all apply / norm-on-load paths), batched occlusion phase of the
# fused stage, in_apply pixels-per-thread A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
( for v in "" "FAV_APPLY_ITER=3" "FAV_APPLY_ITER=2"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/i_ablate.log 2>&1; cat gpurun_out/i_ablate.log | cut -c1-900
for it in 4 3 2; do
FAV_APPLY_ITER=$it timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/i_bench_iter$it.json 2> gpurun_out/i_bench_iter$it.err
done
python - <<'PY'
import json
for f in ("i_bench_iter4", "i_bench_iter3", "i_bench_iter2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d.get("value_full"), d["e2e"]["value"], d["roofline_stage"]["ms"], d["roofline_front"]["ms"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
PY
