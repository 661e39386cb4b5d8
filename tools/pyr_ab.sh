#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- reloc $*"; env "$@" HYP=${HYP:-2048} timeout 300 python tools/reloc_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'; }
{
run LOCREG_SORT=1
for v in v1 v2 v3 v4; do run LOCREG_SO=liblocreg_$v.so; done
} 2>&1 | tee gpurun_out/pyr_ab.log
