#!/bin/bash
# round-2 GPU call E: ncu launch list + full captures (exported to text on the box: .ncu-rep files of a whole frame exceed the
# 64 MiB return limit), file-pipeline bench, new tests
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_front.py tests/test_gpu_net.py -m gpu -q -x -k "corners or video_driver or temporal_stage or image_model" > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log
timeout 900 python tools/file_pipeline_bench.py > gpurun_out/e_filepipe.log 2>&1; tail -2 gpurun_out/e_filepipe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/ncu_frame.py > gpurun_out/e_ncu1.log 2>&1; tail -1 gpurun_out/e_ncu1.log
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/ncu/r02_frame python tools/ncu_frame.py > gpurun_out/e_ncu2.log 2>&1; tail -1 gpurun_out/e_ncu2.log
ncu -i /tmp/ncu/r02_frame.ncu-rep --page raw --csv > gpurun_out/r02_frame_raw.csv 2>/dev/null
ncu -i /tmp/ncu/r02_frame.ncu-rep --page details > gpurun_out/r02_frame_details.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_res -c 2 -o gpurun_out/r02_conv_res python tools/ncu_frame.py > gpurun_out/e_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"in_apply|temporal_stage|temporal_input" -c 6 -o gpurun_out/r02_apply_front python tools/ncu_frame.py > gpurun_out/e_ncu4.log 2>&1
gzip -9 gpurun_out/r02_frame_raw.csv gpurun_out/r02_frame_details.txt
du -sh gpurun_out; ls -la gpurun_out | tail -12
