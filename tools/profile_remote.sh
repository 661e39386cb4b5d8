#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the measured bench line, then an ncu launch list of the same command (headline
# only) with per-launch duration and DRAM bytes, from which tools/traffic_from_ncu.py writes profiles/roofline_traffic.json.
# Outputs land in gpurun_out/ (tag = $1).
tag=${1:-r2}
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_bench_under_ncu.log 2>&1
python tools/traffic_from_ncu.py gpurun_out/${tag}_launches_bench.csv gpurun_out/${tag}_roofline_traffic.json > gpurun_out/${tag}_launch_shares.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${tag}_smi.csv
tail -c 600 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_launch_shares.txt
