set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/bench2_r2o.json 2> gpurun_out/bench2_r2o.err; echo "rc2=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench2_r2o.json').read().strip().splitlines()[-1])
print("N=2 value %.4g"%l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l["profile_ms_per_step"])
PY
