#!/bin/bash
# One gpurun call: GPU parity tests, the contract bench, the ncu launch list of the same bench
# command and one `--set full` capture of the top kernels.  Outputs land in gpurun_out/.
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py > gpurun_out/ncu_bench_$TAG.log 2>&1
cat gpurun_out/bench_$TAG.json
python scripts/prof_target.py 24 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate|ntt_pass|msm_partition|msm_fine_scatter' -c 6 \
    -f -o gpurun_out/prof_$TAG python scripts/prof_target.py 24 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
# keep the raw-page CSV (what profiles/ summarises), not the 50 MB report: gpurun_out/ is capped at 64 MiB
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null && rm -f gpurun_out/prof_$TAG.ncu-rep
python scripts/prof_plonk.py 20 > gpurun_out/prof_plonk_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'quotient_kernel|batch_ratio_kernel|perm_ratio_kernel|msm_reduce_kernel' -c 4 \
    -f -o gpurun_out/prof_plonk_$TAG python scripts/prof_plonk.py 20 > gpurun_out/ncu_full_plonk_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_plonk_$TAG.log
ncu -i gpurun_out/prof_plonk_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_plonk_${TAG}_raw.csv 2>/dev/null && rm -f gpurun_out/prof_plonk_$TAG.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
