set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2z.log; tail -3 gpurun_out/pytest_r2z.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r2z.json 2> gpurun_out/bench_ref_r2z.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2z.err
python bench.py --steps 3 --warmup 3 --marlin-rows-logn 16 --no-cpu > gpurun_out/bench_short_r2z.json 2> gpurun_out/bench_short_r2z.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --steps 3 --warmup 3 --marlin-rows-logn 16 --no-cpu > gpurun_out/ncu_bench_r2z.log 2>&1
python scripts/msm_share.py gpurun_out/launches_r2z.csv | tee gpurun_out/r2z_msm_step_share.txt
python scripts/launch_summary.py gpurun_out/launches_r2z.csv > gpurun_out/r2z_launches_summary.txt 2>&1; head -30 gpurun_out/r2z_launches_summary.txt
gzip -f gpurun_out/launches_r2z.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/phase_share.py > gpurun_out/r2z_phase.txt 2>&1; cat gpurun_out/r2z_phase.txt
