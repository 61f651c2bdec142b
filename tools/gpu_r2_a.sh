#!/bin/bash
# round-2 GPU call A: golden vectors of the reference warp kernel, the full GPU test suite, ablations of the dormant
# producers, conv timeline.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python tests/golden/make_warp_golden.py > gpurun_out/a_golden.log 2>&1 && cp gpurun_out/warp_ref.npz tests/golden/warp_ref.npz
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
( for v in "" "FAV_TMA=1" "FAV_APROD=4" "FAV_DBG=2" "FAV_DBG=4" "FAV_DBG=6" "FAV_DBG=8" "FAV_DBG=14" "FAV_CBG=2" "FAV_NO_NL=1" "FAV_NL_MODE=2"; do
    timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/a_ablate.log 2>&1
( FAV_TMA=1 timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x 2>&1 | tail -3; FAV_APROD=4 timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x 2>&1 | tail -3 ) > gpurun_out/a_variants_pytest.log 2>&1
timeout 300 python tools/trace_conv.py > gpurun_out/a_trace.log 2>&1
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 600 gpurun_out/a_bench.json
