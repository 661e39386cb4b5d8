#!/bin/bash
# run the probe for each experiment build
for so in "$@"; do
  echo "=== $so"
  LOCREG_SO=$so PROBE_SCANS=${PROBE_SCANS:-296} timeout 300 python tools/gpu_probe.py 2>&1 | grep -E "p2plane_single_loop0_kernel|p2plane_batch_scans|p2p_batch_scans|p2p_single_loop0_kernel|ndt_single|ndt_batch|Error|error" 
done
