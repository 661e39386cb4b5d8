#!/bin/bash
# Runs ON THE GPU BOX: --set full capture of one stage-2 launch of a relocalisation.  $1 = tag, $2 = kernel regex, $3 = launches to skip
tag=${1:-r2p}
HYP=${HYP:-512} ncu --set full --clock-control none --import-source on -k regex:"${2:-^k_icp_nn_pyr$}" -s ${3:-3} -c 1 -f -o gpurun_out/${tag} \
    python tools/reloc_breakdown.py > gpurun_out/${tag}.log 2>&1
tail -3 gpurun_out/${tag}.log
