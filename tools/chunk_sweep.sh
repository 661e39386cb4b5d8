#!/bin/bash
# End-to-end rate of locreg_align_batch (host buffers) for different chunkings of the copy / compute overlap.
# LOCREG_CHUNK_WEIGHTS is read once per process, so every setting is its own bench.py run (headline workload only).
# usage (GPU box): bash tools/chunk_sweep.sh "1,8" "1,4,4" ...  > gpurun_out/chunk_sweep.txt
for w in "$@"; do
  line=$(LOCREG_CHUNK_WEIGHTS="$w" python bench.py --configs none --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1)
  python - "$w" "$line" <<'P'
import json, sys
d = json.loads(sys.argv[2])
print("weights %-10s  e2e %7.1f M points/s   resident %7.1f M points/s  (%.2f ms per step)" %
      (sys.argv[1], d["e2e"]["value"] / 1e6, d["value"] / 1e6, d["ms_per_step"]))
P
done
