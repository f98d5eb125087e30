mkdir -p gpurun_out
timeout 300 python tools/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1
echo "sanitizer rc=$?"; tail -3 gpurun_out/san_plain.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|sanitize_small" gpurun_out/san_memcheck.log | head -20
