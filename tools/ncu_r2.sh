#!/bin/bash
# Runs ON THE GPU BOX: --set full captures of single launches of the batch pipeline (S=512).  $1 = tag
tag=${1:-r2i}
cap() {  # name regex skip
  S=512 REPS=1 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -f -o gpurun_out/${tag}_$1 \
      python tools/ncu_batch.py > gpurun_out/${tag}_$1.log 2>&1
}
cap nn_it2 '^k_icp_nn$' 1
cap nn_it3 '^k_icp_nn$' 2
cap finish_it0 '^k_icp_nn_finish$' 0
cap post_it1 '^k_icp_post$' 1
cap fit_it1 '^k_icp_fit$' 1
