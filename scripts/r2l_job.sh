set -x
mkdir -p gpurun_out
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_target.py > gpurun_out/r2l_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitize_plain.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|sanitize target" gpurun_out/r2l_memcheck.log | head -10; tail -3 gpurun_out/r2l_memcheck.log
