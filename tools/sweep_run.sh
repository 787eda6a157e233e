#!/bin/bash
# Runs on the GPU box: device-step + kernel numbers for every variant.
mkdir -p gpurun_out
for f in build/variants/*.so; do
  for mb in 4096; do
    echo "== $f max_batch=$mb"
    B2EXEC_LIB=$PWD/$f B2_MAX_BATCH=$mb python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.rstrip()); continue
    k = d['kernels']
    print('step_ms %.3f worker %.0f GB/s pack %.0f unpack %.0f' % (d['ms_per_step'], k['worker']['achieved_gbs'], k['pack']['achieved_gbs'], k['unpack']['achieved_gbs']))
"
  done
done
