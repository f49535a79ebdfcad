set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2g.log
tail -4 gpurun_out/pytest_r2g.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload sweep --steps 5 > gpurun_out/sweep_r2g.json 2> gpurun_out/sweep_r2g.txt; echo "sweep rc=$?"; cat gpurun_out/sweep_r2g.txt | tail -30
python bench.py --curve bls12_381 --logn 22 --no-secondary --steps 5 > gpurun_out/bench_bls_r2g.json 2> gpurun_out/bench_bls_r2g.err; echo "bls rc=$?"; tail -c 1500 gpurun_out/bench_bls_r2g.json
python bench.py --workload ntt --curve bls12_381 --no-secondary --steps 5 > gpurun_out/bench_bls_ntt_r2g.json 2> gpurun_out/bench_bls_ntt_r2g.err; echo "bls ntt rc=$?"; tail -c 800 gpurun_out/bench_bls_ntt_r2g.json
# ncu: full captures of the MSM kernels (BN254 2^24, BLS 2^22)
python scripts/prof_msm.py 24 1 0 > gpurun_out/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate|msm_partition|msm_fine_scatter|msm_reduce_kernel' -c 4 -f -o gpurun_out/full_msm_r2g python scripts/prof_msm.py 24 1 0 > gpurun_out/ncu_full_msm_r2g.log 2>&1
ncu -i gpurun_out/full_msm_r2g.ncu-rep --page raw --csv > gpurun_out/full_msm_r2g_raw.csv 2>/dev/null; rm -f gpurun_out/full_msm_r2g.ncu-rep
python scripts/prof_msm.py 22 1 1 > gpurun_out/r2g_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate' -c 1 -f -o gpurun_out/full_bls_r2g python scripts/prof_msm.py 22 1 1 > gpurun_out/ncu_full_bls_r2g.log 2>&1
ncu -i gpurun_out/full_bls_r2g.ncu-rep --page raw --csv > gpurun_out/full_bls_r2g_raw.csv 2>/dev/null; rm -f gpurun_out/full_bls_r2g.ncu-rep
ls -la gpurun_out | tail -12
