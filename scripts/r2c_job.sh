set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_reference.py -m gpu -x -q 2>&1 | tail -5
python scripts/prof_msm.py 24 2 > gpurun_out/r2c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python scripts/prof_msm.py 24 2 > gpurun_out/ncu_r2c.log 2>&1
python scripts/msm_share.py gpurun_out/launches_r2c.csv
