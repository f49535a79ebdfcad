set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in 1 2 4; do KZGPU_MSM_NCHUNKS=$c python scripts/e2e_shard.py 21 22 23; done 2>&1 | grep "2\^" | tee gpurun_out/r2j_e2e_shard.txt
