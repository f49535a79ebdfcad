set -x
mkdir -p gpurun_out
python scripts/phase_share.py > gpurun_out/r2a_phase.txt 2>&1
cat gpurun_out/r2a_phase.txt
python scripts/prof_msm.py 22 1 1 > gpurun_out/r2a_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate|msm_reduce_kernel' -c 2 -f -o gpurun_out/full_bls_r2a python scripts/prof_msm.py 22 1 1 > gpurun_out/ncu_full_bls_r2a.log 2>&1
tail -3 gpurun_out/ncu_full_bls_r2a.log
ncu -i gpurun_out/full_bls_r2a.ncu-rep --page raw --csv > gpurun_out/full_bls_r2a_raw.csv 2>/dev/null
ls -la gpurun_out/full_bls_r2a.ncu-rep
nvidia-smi topo -m; nproc; free -g | head -2; lscpu | grep -E "Model name|NUMA|Socket"
