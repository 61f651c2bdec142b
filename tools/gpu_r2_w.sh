#!/bin/bash
# round-2 GPU call W: patch-ring depth on the narrow layers (FAV_ASTAGES): is the operand round trip the bound?
mkdir -p gpurun_out
export FAV_ABL_ONLY="l0+l1+l2+l8+l9+l10"
( for v in "" "FAV_ASTAGES=3" "FAV_ASTAGES=4" "FAV_ASTAGES=2"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/w_ablate.log 2>&1; cat gpurun_out/w_ablate.log | cut -c1-300
FAV_ASTAGES=3 timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "every_layer" > gpurun_out/w_pytest.log 2>&1; tail -2 gpurun_out/w_pytest.log
