#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- reloc $*"; env "$@" HYP=${HYP:-2048} timeout 300 python tools/reloc_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'; }
{
run LOCREG_SORT=1
echo "--- batch"; S=512 timeout 300 python tools/icp_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'
echo "--- track"; timeout 300 python tools/track_latency.py 2>&1 | tail -3
} 2>&1 | tee gpurun_out/pyr_ab.log
