# Config 4 as written (4096 scans, 1.8 GB pinned) end to end on one GPU for two chunkings of the copy / compute overlap:
# weights:streams pairs, one bench.py run each (GPU box: bash tools/c4s_chunks.sh > gpurun_out/r2_c4s_chunks.txt).
for cfg in "1,3,9,27,41:1" "1,3,9,27,41:0"; do
  w=${cfg%%:*}; s=${cfg##*:}
  line=$(LOCREG_CHUNK_WEIGHTS="$w" LOCREG_CHUNK_STREAMS=$s timeout 27 python bench.py --configs C4S --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | tail -1)
  python - "$w" "$s" "$line" <<'P'
import json, sys
try:
    d = json.loads(sys.argv[3]); c = d["configs"]["C4_strong"]
    print("weights %s streams %s: C4S e2e %.1f M  resident %.1f M | 512-scan e2e %.1f M" % (sys.argv[1], sys.argv[2], c["e2e"]["value"]/1e6, c["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("failed", sys.argv[1], sys.argv[2], e)
P
done
