#!/bin/bash
# round-2 GPU call Q: what bounds the non-residual convolutions?  FAV_DBG sensitivity table (timing only; results are wrong by design)
mkdir -p gpurun_out
export FAV_ABL_ONLY="l0+l1+l2+l8+l9+l10"
( for v in "" "FAV_DBG=1" "FAV_DBG=2" "FAV_DBG=4" "FAV_DBG=8" "FAV_DBG=16" "FAV_DBG=32" "FAV_DBG=64" "FAV_NO_KSPLIT=1"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/q_ablate.log 2>&1; cat gpurun_out/q_ablate.log | cut -c1-300
