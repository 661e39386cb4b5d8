#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the measured bench line, the ncu launch list of the same command, and one
# --set full capture each of the two dominant kernels inside bench.py.  Outputs land in gpurun_out/ (tag = $1).
tag=${1:-r1}
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# steady-state iteration of the first warm-up step: launches of one iteration = nn, finish, post, solve
ncu --set full --clock-control none --import-source on -k regex:'^k_icp_nn$' -s 8 -c 1 \
    -o gpurun_out/${tag}_prof_nn python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_nn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_icp_post$' -s 8 -c 1 \
    -o gpurun_out/${tag}_prof_post python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_post.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${tag}_smi.csv
cat gpurun_out/${tag}_bench.json
