set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests/test_gpu_core.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -5
