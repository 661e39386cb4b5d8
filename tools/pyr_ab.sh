#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for f in "LOCREG_SORT=0" "LOCREG_SORT_BATCH=1 LOCREG_SORT_FRAC=0.05" "LOCREG_SORT_BATCH=1 LOCREG_SORT_FRAC=0.05 LOCREG_SORT_BIN=2" "LOCREG_SORT_BATCH=1 LOCREG_SORT_FRAC=0.05 LOCREG_SORT_BIN=0.5 LOCREG_SORT_SUB=4" "LOCREG_SORT_BATCH=1 LOCREG_SORT_FRAC=0.001"; do
  echo "--- batch $f"
  env $f S=512 timeout 300 python tools/icp_breakdown.py 2>&1 | tail -1 | sed 's/.*total/total/'
done
} 2>&1 | tee gpurun_out/pyr_ab.log
