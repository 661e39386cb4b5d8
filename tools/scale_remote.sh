#!/bin/bash
# Runs ON THE GPU BOX with N GPUs visible: the full bench line (every BASELINE config) at 1/2/4/8 ranks.
# Outputs: gpurun_out/<tag>_scale.jsonl    usage: scale_remote.sh tag maxn [first_n]
tag=${1:-r2}; maxn=${2:-8}; first=${3:-1}
out=gpurun_out/${tag}_scale.jsonl; : > $out
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  [ $n -lt $first ] && continue
  if [ $n -eq 1 ]; then python bench.py --no-cpu-baseline --steps 5 >> $out 2>/dev/null
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --no-cpu-baseline 2>gpurun_out/${tag}_scale_n$n.err | grep '^{' >> $out; fi
done
python - <<PY
import json
for l in open("$out"):
    d=json.loads(l); c=d.get("configs",{})
    print("N=%d  headline %.0f M pts/s (e2e %.0f)  C4S %.0f M (e2e %.0f)  C5 %.0f hyp/s (%.2f s, %s)  C3 %.3f ms" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6,
          c.get("C4_strong",{}).get("value",0)/1e6, c.get("C4_strong",{}).get("e2e",{}).get("value",0)/1e6, c.get("C5",{}).get("value",0), c.get("C5",{}).get("ms_per_step",0)/1e3,
          c.get("C5",{}).get("config",{}).get("collective"), c.get("C3",{}).get("ms_per_step",0)))
PY
