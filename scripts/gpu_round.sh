#!/bin/bash
# One gpurun call: GPU parity tests, the contract bench, the ncu launch list of the same bench command, smoke(), and
# (scripts/gpu_ncu_full.sh) one `--set full` capture per kernel family.  Outputs land in gpurun_out/.
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py > gpurun_out/ncu_bench_$TAG.log 2>&1
cat gpurun_out/bench_$TAG.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$2" = "full" ]; then bash scripts/gpu_ncu_full.sh $TAG; fi
# host topology of the box (explains e2e variance between boxes: PCIe link, NUMA node of the GPU)
{ nvidia-smi --query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; nvidia-smi topo -m; \
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done; \
  ls /sys/devices/system/node/ | grep -c '^node'; nproc; } > gpurun_out/topo_$TAG.txt 2>&1
