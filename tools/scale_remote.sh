#!/bin/bash
# Runs ON THE GPU BOX with N GPUs visible: the bench at 1/2/4/8 ranks (weak-scaling batch ICP) and the full
# config-5 relocalisation at N ranks.  Outputs: gpurun_out/<tag>_scale.jsonl
tag=${1:-r1}; maxn=${2:-8}
out=gpurun_out/${tag}_scale.jsonl; : > $out
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  if [ $n -eq 1 ]; then python bench.py --no-cpu-baseline --steps 5 >> $out 2>/dev/null
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --no-cpu-baseline 2>/dev/null | grep '^{' >> $out; fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $maxn --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $maxn --workload reloc --hyp 65536 --steps 1 2>/dev/null | grep '^{' >> $out
python - <<PY
import json
for l in open("$out"):
    d=json.loads(l); print(d["metric"], d["n_gpus"], round(d["value"]), d.get("ms_per_step"), d.get("e2e",{}).get("value"), d.get("best"))
PY
