#!/bin/bash
# round-2 final evidence run: whole GPU suite, smoke(), launch list, default bench, reference arm, conv timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log; tail -12 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_final.csv python tools/ncu_frame.py > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 400 gpurun_out/final_ref.json
timeout 300 python tools/trace_conv.py > gpurun_out/final_trace.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "full", d.get("value_full"), "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "us/launch", d["roofline"]["us_per_launch"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["clocks"])
PY
