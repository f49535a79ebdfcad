set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --marlin-rows-logn 16 > gpurun_out/bench2_r2h.json 2> gpurun_out/bench2_r2h.err; echo "rc=$?"
tail -3 gpurun_out/bench2_r2h.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench2_r2h.json').read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["scaling"]); print(l.get("marlin")); print(l["weak"]["value"])
PY
