#!/bin/bash
# 8-GPU validation (gpurun --gpus 8): the whole GPU suite incl. the multi-GPU tests, the 8-rank peer check, and the
# bench at N = 8 / 4 / 2 (value = one 10 000-electron lattice sharded over the ranks; single_dispatcher; peer / bcast / c4).
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mp_peer_check.py 2>&1 | grep -E "^rank" | sort | head -16
for n in 8 4 2; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  grep "bench\]" gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    sd = d.get("single_dispatcher") or {}
    print("N=%d value %.0f k el/s (%.1f ms/step, ranks %s) single dispatcher %.0f k  worker %.3f pack %.3f  peer %.0f GB/s verified=%s  bcast exact=%s  c4 %.1f ms exact=%s  clocks %s" % (
        d["n_gpus"], d["value"] / 1e3, d["ms_per_step"], d["e2e"]["ms_per_rank"].get("all"), sd.get("value", 0) / 1e3,
        d["roofline"]["frac"], d["kernels"]["pack"]["frac"], d["peer"]["gbs_per_direction"], d["peer"]["verified"],
        d["bcast"]["exact"], d["c4"]["ms"], d["c4"]["exact"], d["clocks"]["reasons"]))
except Exception as e:
    print("N=$n failed:", e)
PY
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
