set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --marlin-rows-logn 16 > gpurun_out/bench2_r2r.json 2> gpurun_out/bench2_r2r.err; echo "rc2=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench2_r2r.json').read().strip().splitlines()[-1])
print("N=2 value %.4g"%l["value"], "ms", l["ms_per_step"], "marlin", l["marlin"]["prove_s"], "ntt", l["ntt"]["value"])
PY
python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
