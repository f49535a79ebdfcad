set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f.log
tail -4 gpurun_out/pytest_r2f.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
tail -4 gpurun_out/bench_r2f.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2f.json').read().strip().splitlines()[-1])
print("value", l["value"], "ms", l["ms_per_step"]); print("e2e", l["e2e"]); print("roofline frac", l["roofline"]["frac"], l["roofline"]["step_frac"])
print("ntt e2e", l["ntt"]["e2e"]); print("pyb", l["plonk"]["python_boundary"]); print("plonk bundled", l["plonk"]["bundled"]["dropin_hotpath_s"])
PY
