#!/bin/bash
# Runs ON THE GPU BOX: sanitizer passes over the long-queue stage 2 (tools/sanitize_reloc.py).  $1 = tag
tag=${1:-r2}
cd "$(dirname "$0")/.."
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_reloc.py > gpurun_out/${tag}_sanitizer_reloc_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_reloc ok|winner" gpurun_out/${tag}_sanitizer_reloc_$tool.log | tail -5
done
LOCREG_PYR_KERNEL=1 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_reloc.py > gpurun_out/${tag}_sanitizer_reloc_pyr_memcheck.log 2>&1
echo "== memcheck, pyramid kernel"; grep -E "ERROR SUMMARY|sanitize_reloc ok|winner" gpurun_out/${tag}_sanitizer_reloc_pyr_memcheck.log | tail -5
LOCREG_SORT=0 python tools/sanitize_reloc.py | grep winner
