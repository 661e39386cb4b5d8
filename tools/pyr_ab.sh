#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- reloc $*"; env "$@" HYP=${HYP:-8192} timeout 300 python tools/reloc_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'; }
{
run LOCREG_RELOC_MID_SHELLS=0
run LOCREG_RELOC_MID_SHELLS=1
} 2>&1 | tee gpurun_out/pyr_ab.log
