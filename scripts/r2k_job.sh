set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY' 2>&1 | tee gpurun_out/r2k_split_ab.txt
import os, sys, subprocess
for env in ({"KZGPU_MSM_NO_SPLIT": "1"}, {}):
    e = dict(os.environ, **env)
    print("split", "OFF" if env else "ON", flush=True)
    out = subprocess.run([sys.executable, "scripts/e2e_shard.py", "16", "18", "20", "21", "22", "24"], env=e, capture_output=True, text=True)
    print(out.stdout, out.stderr[-500:], flush=True)
PY
