#!/bin/bash
# N = 1, 2, 4, 8 on one box, the way the driver launches bench.py.
mkdir -p gpurun_out
STEPS=${STEPS:-5}
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps $STEPS --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -2 gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=%d value %.1f M el/s (%.3f ms/step)  e2e %.1f k el/s (%.1f ms/step)  worker %.3f pack %.3f clocks %s" % (
        d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e3, d["e2e"]["ms_per_step"],
        d["roofline"]["frac"], d["kernels"]["pack"]["frac"], d["clocks"]["reasons"]))
except Exception as e:
    print("N=$n failed:", e)
PY
done
