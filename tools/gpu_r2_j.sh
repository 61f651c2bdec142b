#!/bin/bash
# round-2 GPU call J (--gpus N): the default bench and the cfg3 data plane at N ranks, launched the way the driver does
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err; tail -c 600 gpurun_out/j_bench_n$N.json; tail -2 gpurun_out/j_bench_n$N.err
timeout 900 $TR bench.py --gpus $N --config cfg3 --steps 24 > gpurun_out/j_cfg3_n$N.json 2> gpurun_out/j_cfg3_n$N.err; tail -c 1600 gpurun_out/j_cfg3_n$N.json; tail -2 gpurun_out/j_cfg3_n$N.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/j_ref_n$N.json 2> gpurun_out/j_ref_n$N.err; tail -c 500 gpurun_out/j_ref_n$N.json
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "session or video_driver" > gpurun_out/j_pytest.log 2>&1; tail -3 gpurun_out/j_pytest.log; fi
