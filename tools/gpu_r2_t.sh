#!/bin/bash
# round-2 GPU call T: 256-bit stores in the phase-fold epilogue (u64, u32): parity + per-layer timing + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "every_layer or paper_arch or arch_tokens or comparator or golden" > gpurun_out/t_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/t_pytest.log; tail -3 gpurun_out/t_pytest.log
timeout 300 python tools/ablate.py "" > gpurun_out/t_ablate.log 2>&1; cat gpurun_out/t_ablate.log | cut -c1-900
timeout 600 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print("t_bench", d["value"], d.get("value_full"), d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])
PY
