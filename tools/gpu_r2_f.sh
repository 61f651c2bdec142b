#!/bin/bash
# round-2 GPU call F: inter-block IN+skip fusion (nl == 2), native file pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_vr.py -m gpu -q -x --durations=5 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
timeout 600 python -m pytest "tests/test_gpu_parity_large.py::test_run_next_image_full_size_vs_fp32_oracle" -m gpu -q -x > gpurun_out/f_pytest_large.log 2>&1; tail -2 gpurun_out/f_pytest_large.log
( for v in "" "FAV_NO_NL2=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/f_ablate.log 2>&1; cat gpurun_out/f_ablate.log | cut -c1-900
timeout 300 python tools/trace_conv.py > gpurun_out/f_trace.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/f_bench.json")); print(d["value"], d["value_full"], d["e2e"]["value"], d["roofline_stage"]["ms"], d["roofline_front"]["ms"], d["roofline"]["achieved"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/f_bench.err
timeout 900 python tools/file_pipeline_bench.py > gpurun_out/f_filepipe.log 2>&1; tail -1 gpurun_out/f_filepipe.log
