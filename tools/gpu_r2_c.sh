#!/bin/bash
# round-2 GPU call C: PDL + 4-stage residual kernel + fused temporal stage + in_apply restructure
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_front.py tests/test_vr.py -m gpu -q -x --durations=5 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
timeout 600 python -m pytest "tests/test_gpu_parity_large.py::test_run_next_image_full_size_vs_fp32_oracle" -m gpu -q -x > gpurun_out/c_pytest_large.log 2>&1; tail -2 gpurun_out/c_pytest_large.log
( for v in "" "FAV_NO_PDL=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/c_ablate.log 2>&1
timeout 300 python tools/trace_conv.py > gpurun_out/c_trace.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
FAV_NO_PDL=1 timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_nopdl.json 2> gpurun_out/c_bench_nopdl.err
python - <<'PY'
import json
for f in ("c_bench.json","c_bench_nopdl.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, d["value"], d["value_full"], d["e2e"]["value"], d["roofline_stage"]["ms"], d["roofline_front"]["ms"], d["roofline_warp"]["real_flow"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/c_bench.err
