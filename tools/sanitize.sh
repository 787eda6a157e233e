#!/bin/bash
# compute-sanitizer passes over the kernel-level GPU tests (small cases only).
mkdir -p gpurun_out
SEL='not full_size and not 20000 and not many_independent and not 1048576 and not 262145'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --target-processes all \
    python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | tail -5
done
# the classic worker (named barriers only) is what racecheck can verify end to end
echo "== compute-sanitizer --tool racecheck  (B2_WORKER=classic B2_MOVER=ldg)"
B2_WORKER=classic B2_MOVER=ldg timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_racecheck_classic.log 2>&1
echo "exit $?"
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_racecheck_classic.log | tail -3
