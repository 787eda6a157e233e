#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mp_peer_check.py 2>&1 | grep -E "^rank" | sort | head -16
for n in 4 8; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  tail -2 gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=%d value %.1f M el/s (%.3f ms/step)  e2e %.1f k el/s (%.1f ms/step)  worker %.3f pack %.3f peer %.0f GB/s clocks %s" % (
        d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e3, d["e2e"]["ms_per_step"],
        d["roofline"]["frac"], d["kernels"]["pack"]["frac"], d["peer"]["gbs_per_direction"], d["clocks"]["reasons"]))
except Exception as e:
    print("N=$n failed:", e)
PY
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 2>/dev/null | cut -c1-200
