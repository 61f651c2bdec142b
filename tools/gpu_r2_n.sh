#!/bin/bash
# round-2 GPU call N: cfg3 with byte payloads at N=1 (before the multi-GPU run), device-side conversion test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_front.py -m gpu -q -x -k "byte_conversions or stage" > gpurun_out/n_pytest.log 2>&1; tail -2 gpurun_out/n_pytest.log
timeout 900 python bench.py --config cfg3 --steps 24 > gpurun_out/n_cfg3_bytes_n1.json 2> gpurun_out/n_cfg3_bytes_n1.err; tail -c 1300 gpurun_out/n_cfg3_bytes_n1.json; tail -2 gpurun_out/n_cfg3_bytes_n1.err
timeout 900 python bench.py --config cfg3 --payload fp32 --steps 24 > gpurun_out/n_cfg3_fp32_n1.json 2> gpurun_out/n_cfg3_fp32_n1.err; tail -c 700 gpurun_out/n_cfg3_fp32_n1.json
