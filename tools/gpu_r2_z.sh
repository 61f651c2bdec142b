#!/bin/bash
# round-2 GPU call Z: 4-way K-split (four issuing warps) for generic plans of <= 64 columns (d64): parity + timing, env-gated
mkdir -p gpurun_out
FAV_KSPLIT4=1 timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "every_layer or golden or comparator" > gpurun_out/z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z_pytest.log; tail -3 gpurun_out/z_pytest.log
export FAV_ABL_ONLY="l0+l1+l2+l8+l9+l10"
( for v in "" "FAV_KSPLIT4=1" "" "FAV_KSPLIT4=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/z_ablate.log 2>&1; cat gpurun_out/z_ablate.log | cut -c1-300
