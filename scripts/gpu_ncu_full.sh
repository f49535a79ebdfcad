#!/bin/bash
# `ncu --set full` captures of the dominant kernels, one invocation per kernel family so that the launch-count limit
# cannot starve a family; only the raw-page CSVs are kept (gpurun_out/ is capped at 64 MiB).
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
cap() {  # name regex count target...
  local name=$1 regex=$2 count=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -c $count -f -o gpurun_out/full_${name}_$TAG "$@" > gpurun_out/ncu_full_${name}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_${name}_$TAG.log
  ncu -i gpurun_out/full_${name}_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${name}_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/full_${name}_$TAG.ncu-rep
}
cap msm 'msm_accumulate|msm_partition|msm_fine_scatter|msm_reduce_kernel' 4 python scripts/prof_msm.py 24 1
cap ntt 'ntt_pass' 3 python scripts/prof_target.py 24
cap plonk 'quotient_kernel|batch_ratio_kernel|perm_ratio_kernel' 3 python scripts/prof_plonk.py 20
