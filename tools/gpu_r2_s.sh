#!/bin/bash
# round-2 GPU call S: lean generic issue loop in conv_tc (d64, d128, generic 3x3): parity + per-layer timing + bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_parity_large.py -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s_pytest.log; tail -3 gpurun_out/s_pytest.log
( for v in "" "FAV_ABL_ARCH=paper"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/s_ablate.log 2>&1; cat gpurun_out/s_ablate.log | cut -c1-900
timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print("s_bench", d["value"], d.get("value_full"), d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
PY
timeout 300 python tools/trace_conv.py > gpurun_out/s_trace.log 2>&1; grep -A4 '"layer": "l1"\|"layer": "l2"' gpurun_out/s_trace.log | cut -c1-260
