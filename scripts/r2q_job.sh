set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
KZGPU_DEVICES=all python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -15
