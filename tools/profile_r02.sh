#!/bin/bash
# Round-2 evidence run (one GPU, under gpurun): sanitizer passes, ncu launch list of the bench command,
# ncu --set full captures of the three hot kernels.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
python -m covalent_ssh_plugin_b200.build > /dev/null
SEL='not full_size and not 20000 and not many_independent and not 1048576 and not 262145'
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool (default kernels)"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --target-processes all \
    python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "exit $?"
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitize_$tool.log | tail -3
done
echo "== compute-sanitizer --tool memcheck (submit / cancel / tables / both worker flavours)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_submit_gpu.py -m gpu -x -q -k "not 70000 and not 16384 and not 3145733" > gpurun_out/r02_sanitize_memcheck_submit.log 2>&1
echo "exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_sanitize_memcheck_submit.log | tail -3
echo "== compute-sanitizer --tool memcheck (resident worker)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_resident_gpu.py -m gpu -x -q -p no:timeout > gpurun_out/r02_sanitize_memcheck_resident.log 2>&1
echo "exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_sanitize_memcheck_resident.log | tail -3
echo "== compute-sanitizer --tool racecheck (B2_WORKER=classic B2_MOVER=ldg)"
B2_WORKER=classic B2_MOVER=ldg timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_racecheck_classic.log 2>&1
echo "exit $?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r02_sanitize_racecheck_classic.log | tail -3

echo "== ncu launch list of the bench command"
B2_BENCH_SKIP_RESIDENT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "exit $?"
for k in worker_tma_kernel worker_tma_small_kernel span_move_tma_kernel; do
  echo "== ncu --set full $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_$k python tools/ncu_target.py > gpurun_out/r02_ncu_$k.log 2>&1
  echo "exit $?"
done
ls -la gpurun_out/*.ncu-rep
