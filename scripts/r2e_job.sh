set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
python bench.py --workload inproc --devices 1 --steps 5 > gpurun_out/inproc1_r2e.json 2> gpurun_out/inproc1_r2e.err; tail -c 2500 gpurun_out/inproc1_r2e.json; tail -3 gpurun_out/inproc1_r2e.err
python bench.py --workload inproc --devices 2 --steps 5 > gpurun_out/inproc2_r2e.json 2> gpurun_out/inproc2_r2e.err; tail -c 2500 gpurun_out/inproc2_r2e.json; tail -3 gpurun_out/inproc2_r2e.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_r2e.json 2> gpurun_out/bench2_r2e.err; tail -c 3000 gpurun_out/bench2_r2e.json; tail -3 gpurun_out/bench2_r2e.err
