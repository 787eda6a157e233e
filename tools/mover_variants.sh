#!/bin/bash
# Runs on the GPU box: correctness + pack/unpack rates for the three span-mover variants.
for m in ldg wide tma; do
  echo "== B2_MOVER=$m"
  B2_MOVER=$m timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "move_spans or pack" 2>&1 | tail -1
  B2_MOVER=$m python tools/sweep_c5.py --payloads 1024,16384,65536,1048576,16777216 --ns 10000 --no-e2e --wave-gib 4 2>&1 | grep -E "^\| [0-9]" | cut -d'|' -f2,3,8,9
done
