set -x
mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 --no-secondary > gpurun_out/bench${n}_r2v.json 2> gpurun_out/bench${n}_r2v.err; echo "rc$n=$?"
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu > gpurun_out/bench1_r2v.json 2> gpurun_out/bench1_r2v.err; echo "rc1=$?"
python - <<'PY'
import json
v={}
for n in (1,2,4,8):
    l=json.loads(open(f'gpurun_out/bench{n}_r2v.json').read().strip().splitlines()[-1]); v[n]=l
    print(n, "value %.4g"%l["value"], "ms %.3f"%l["ms_per_step"], "e2e %.3f"%l["e2e"]["ms_per_step"], l["profile_ms_per_step"], "eff %.3f"%(l["value"]/(n*v[1]["value"])))
PY
