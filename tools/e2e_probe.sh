#!/bin/bash
run() { echo "== $1"; shift; env "$@" python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e %.1f k el/s  %.1f ms/step' % (d['e2e']['value']/1e3, d['e2e']['ms_per_step']))"; }
run plain X=1
run bind B2_BENCH_BIND=1
run torch+nccl B2_BENCH_FORCE_TORCH=1
run omp1 OMP_NUM_THREADS=1
run torch+nccl+bind B2_BENCH_FORCE_TORCH=1 B2_BENCH_BIND=1
