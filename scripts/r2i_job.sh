set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench8_r2i.json 2> gpurun_out/bench8_r2i.err; echo "rc8=$?"
tail -2 gpurun_out/bench8_r2i.err
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 10 --warmup 3 --no-secondary > gpurun_out/bench4_r2i.json 2> gpurun_out/bench4_r2i.err; echo "rc4=$?"
python bench.py --workload inproc --devices 8 --steps 5 --marlin-rows-logn 18 > gpurun_out/inproc8_r2i.json 2> gpurun_out/inproc8_r2i.err; echo "rcin8=$?"; tail -2 gpurun_out/inproc8_r2i.err
python bench.py --workload inproc --devices 4 --steps 5 --marlin-rows-logn 18 > gpurun_out/inproc4_r2i.json 2> gpurun_out/inproc4_r2i.err; echo "rcin4=$?"
python - <<'PY'
import json
for f in ("bench8_r2i","bench4_r2i","inproc8_r2i","inproc4_r2i"):
    try:
        l=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, "value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l.get("profile_ms_per_step"))
        for k in ("weak","marlin","device_resident","batched_commit","batched_ntt","marlin_synthetic"):
            if k in l: print("   ", k, json.dumps(l[k])[:600])
        if "ntt" in l: print("    ntt", l["ntt"]["value"], l["ntt"]["e2e"]["ms_per_step"], l["ntt"]["e2e"].get("copy_only_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
grep -c "NCCL INFO" gpurun_out/bench4_r2i.json gpurun_out/bench4_r2i.err
