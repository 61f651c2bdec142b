#!/bin/bash
# round-2 GPU call P: cfg3 at N=1 with the host bound to the GPU's NUMA node + PCIe probe, fixed conversion test, memcheck
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_front.py -m gpu -q -x -k "byte_conversions" > gpurun_out/p_pytest.log 2>&1; tail -2 gpurun_out/p_pytest.log
timeout 900 python bench.py --config cfg3 --steps 24 > gpurun_out/p_cfg3_bytes_n1.json 2> gpurun_out/p_cfg3_bytes_n1.err; tail -c 900 gpurun_out/p_cfg3_bytes_n1.json; tail -2 gpurun_out/p_cfg3_bytes_n1.err
timeout 900 python bench.py --config cfg3 --payload fp32 --steps 24 > gpurun_out/p_cfg3_fp32_n1.json 2> gpurun_out/p_cfg3_fp32_n1.err; tail -c 900 gpurun_out/p_cfg3_fp32_n1.json
timeout 300 python tools/pcie_bw.py > gpurun_out/p_pcie.log 2>&1; tail -4 gpurun_out/p_pcie.log
bash tools/gpu_r2_o.sh
