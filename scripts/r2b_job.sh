set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -5 gpurun_out/pytest_r2b.log
python scripts/phase_share.py > gpurun_out/r2b_phase.txt 2>&1
cat gpurun_out/r2b_phase.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
