set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
  if [ $v = 1 ]; then export KZGPU_NO_CLASS_TABLES=1; else unset KZGPU_NO_CLASS_TABLES; fi
  python - <<'PY' 2>&1 | tail -3
import os, json, bench
from kzg_snark_b200 import _ffi
_ffi.init()
r = bench.marlin_synthetic_prove(20)
print("class tables", "OFF" if os.environ.get("KZGPU_NO_CLASS_TABLES") else "ON", json.dumps({k: r[k] for k in ("prove_s", "first_prove_s", "rounds_s")}))
PY
done 2>&1 | tee gpurun_out/r2m_class_tables_marlin.txt
