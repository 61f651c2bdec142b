#!/bin/bash
# round-2 GPU call V: combined FAV_DBG ablations on the narrow layers (no epilogue + tiny patch copies + tiny weight copies)
mkdir -p gpurun_out
export FAV_ABL_ONLY="l0+l1+l2+l8+l9+l10"
( for v in "" "FAV_DBG=12" "FAV_DBG=14" "FAV_DBG=78"; do timeout 300 python tools/ablate.py "$v"; done ) > gpurun_out/v_ablate.log 2>&1; cat gpurun_out/v_ablate.log | cut -c1-300
