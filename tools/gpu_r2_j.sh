#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus N): the default bench and the cfg3 data plane at N ranks, launched the way the driver does
N=${1:-2}
WHAT=${2:-all}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --config cfg3 --steps 24 > gpurun_out/j_cfg3_bytes_n$N.json 2> gpurun_out/j_cfg3_bytes_n$N.err; tail -c 1500 gpurun_out/j_cfg3_bytes_n$N.json; tail -2 gpurun_out/j_cfg3_bytes_n$N.err
timeout 900 $TR bench.py --gpus $N --config cfg3 --payload fp32 --steps 24 > gpurun_out/j_cfg3_fp32_n$N.json 2> gpurun_out/j_cfg3_fp32_n$N.err; tail -c 900 gpurun_out/j_cfg3_fp32_n$N.json
if [ "$WHAT" = "all" ]; then
timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err; tail -c 600 gpurun_out/j_bench_n$N.json; tail -2 gpurun_out/j_bench_n$N.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/j_ref_n$N.json 2> gpurun_out/j_ref_n$N.err; tail -c 500 gpurun_out/j_ref_n$N.json
fi
