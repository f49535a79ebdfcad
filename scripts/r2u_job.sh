set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/phase_share.py bn254:16 bn254:20 bn254:21 bn254:22 bn254:23 bn254:24 bls12_381:22 2>&1 | tee gpurun_out/r2u_phase.txt
