#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- reloc $*"; env "$@" HYP=${HYP:-8192} timeout 300 python tools/reloc_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'; }
{
run LOCREG_SORT=1
run LOCREG_RELOC_COARSE_SHELLS=4
run LOCREG_RELOC_COARSE_SHELLS=6
run LOCREG_RELOC_COARSE_SHELLS=12
run LOCREG_RELOC_WAVE_GIB=8
} 2>&1 | tee gpurun_out/pyr_ab.log
