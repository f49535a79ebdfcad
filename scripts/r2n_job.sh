set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
NCCL_DEBUG=VERSION python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8_r2n.json 2> gpurun_out/bench8_r2n.err; echo "rc8=$?"; tail -2 gpurun_out/bench8_r2n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --no-secondary > gpurun_out/bench4_r2n.json 2> gpurun_out/bench4_r2n.err; echo "rc4=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/bench2_r2n.json 2> gpurun_out/bench2_r2n.err; echo "rc2=$?"
python bench.py --workload inproc --devices 8 --steps 5 --marlin-rows-logn 20 > gpurun_out/inproc8_r2n.json 2> gpurun_out/inproc8_r2n.err; echo "rcin8=$?"; tail -2 gpurun_out/inproc8_r2n.err
python bench.py --workload inproc --devices 4 --steps 5 --marlin-rows-logn 16 --sweep-max 24 > gpurun_out/inproc4_r2n.json 2> gpurun_out/inproc4_r2n.err; echo "rcin4=$?"
python bench.py --workload inproc --devices 2 --steps 5 --marlin-rows-logn 16 --sweep-max 24 > gpurun_out/inproc2_r2n.json 2> gpurun_out/inproc2_r2n.err; echo "rcin2=$?"
python - <<'PY'
import json
for f in ("bench8_r2n","bench4_r2n","bench2_r2n","inproc8_r2n","inproc4_r2n","inproc2_r2n"):
    try:
        l=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, "value %.4g"%l["value"], "ms", round(l["ms_per_step"],3), "e2e", round(l["e2e"]["ms_per_step"],3), l.get("profile_ms_per_step"))
        for k in ("weak","marlin","device_resident","batched_commit","batched_ntt","marlin_synthetic","msm_sweep"):
            if k in l: print("   ", k, json.dumps(l[k])[:700])
        if "ntt" in l: print("    ntt", l["ntt"]["value"], l["ntt"]["e2e"]["ms_per_step"], l["ntt"]["e2e"].get("copy_only_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
