#!/bin/bash
# round-2 GPU call O: compute-sanitizer (memcheck) over the small-size parity tests: out-of-bounds / misaligned accesses of the
# bulk copies, the tcgen05 kernels and the front end would surface here
mkdir -p gpurun_out
export FAV_NO_GRAPH=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_net.py tests/test_gpu_front.py -m gpu -x -q \
  -k "every_layer or arch_tokens or golden_clip or session_file_payload or temporal_stage or byte_conversions or unaligned or corners" \
  > gpurun_out/o_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/o_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck rc|Invalid|misaligned" gpurun_out/o_memcheck.log | tail -12
