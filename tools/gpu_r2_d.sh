#!/bin/bash
# round-2 GPU call D: tests of the new pieces, bench, ncu launch list + full capture of one frame and of every kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_front.py tests/test_vr.py -m gpu -q -x --durations=5 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/d_bench.json")); print(d["value"], d["value_full"], d["e2e"]["value"], d["roofline_stage"]["ms"], d["roofline_front"]["ms"], d["cpu_baseline"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/ncu_frame.py > gpurun_out/d_ncu1.log 2>&1; tail -2 gpurun_out/d_ncu1.log
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r02_frame python tools/ncu_frame.py > gpurun_out/d_ncu2.log 2>&1; tail -2 gpurun_out/d_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_res -c 4 -o gpurun_out/r02_conv_res_src python tools/ncu_frame.py > gpurun_out/d_ncu3.log 2>&1; tail -2 gpurun_out/d_ncu3.log
ls -la gpurun_out/*.ncu-rep
