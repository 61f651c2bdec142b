#!/bin/bash
# round-2 GPU call U: staggered initial ring fill in conv_res: parity subset + per-layer timing + trace + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "every_layer or golden or teacher" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/u_pytest.log; tail -3 gpurun_out/u_pytest.log
timeout 300 python tools/ablate.py "" > gpurun_out/u_ablate.log 2>&1; cat gpurun_out/u_ablate.log | cut -c1-900
timeout 300 python tools/trace_conv.py > gpurun_out/u_trace.log 2>&1; grep -A3 '"layer": "l4.c1"\|"layer": "l4.c2"' gpurun_out/u_trace.log | cut -c1-260
timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench.json").read().strip().splitlines()[-1])
print("u_bench", d["value"], d.get("value_full"), d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["clocks"]["sm_mhz"])
PY
