#!/bin/bash
# Runs on the GPU box: correctness + rates for every streaming-worker variant.
mkdir -p gpurun_out
for f in build/variants/*.so; do
  echo "== $f"
  B2EXEC_LIB=$PWD/$f timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -1
  B2EXEC_LIB=$PWD/$f timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.rstrip()[:200]); continue
    k = d['kernels']
    print('64KiB step_ms %.3f worker %.0f GB/s (%.3f)' % (d['ms_per_step'], k['worker']['achieved_gbs'], k['worker']['frac']))
"
  B2EXEC_LIB=$PWD/$f timeout -s KILL 300 python tools/sweep_c5.py --payloads 16384,65536,1048576 --ns 10000 --no-e2e --wave-gib 4 2>&1 | grep -E "^\| [0-9]" | cut -d'|' -f2,3,5,6,7
done
