#!/bin/bash
# round-2 GPU call M: in-context DRAM traffic of every kernel of a frame (application replay, caches left alone: each kernel
# sees the L2 state its predecessors left), VR driver with the native PNG writer
mkdir -p gpurun_out
timeout 900 ncu --replay-mode application --cache-control none --clock-control none --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum \
  --csv --log-file gpurun_out/r02_frame_incontext.csv python tools/ncu_frame.py > gpurun_out/m_ncu.log 2>&1; tail -2 gpurun_out/m_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/r02_frame_incontext.csv") if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
acc = {}
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]][:44])
    acc.setdefault(key, {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]]
for (i, k), m in sorted(acc.items(), key=lambda kv: int(kv[0][0])):
    print(i, k, m.get("gpu__time_duration.sum"), m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"), m.get("lts__t_sector_hit_rate.pct"))
PY
timeout 900 python -m pytest tests/test_vr.py tests/test_gpu_net.py -m gpu -q -x -k "vr or video_driver" > gpurun_out/m_pytest.log 2>&1; tail -2 gpurun_out/m_pytest.log
timeout 900 python tools/vr_bench.py > gpurun_out/m_vr.log 2>&1; tail -1 gpurun_out/m_vr.log
