#!/bin/bash
# one 2-GPU box: the N = 2 strong-scaling points of both NAtl decks (as scripts/r02_scale8.sh)
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/r02_scale2.sh tag'
tag=${1:-s2}
mkdir -p gpurun_out
port=29560
for w in natl1km natl2km; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 2 --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${w}_n2.json 2> gpurun_out/${tag}_${w}_n2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${w}_n2.json"))
    print("$w N=2 ms/step %.4f steps/s %.1f parity_ok %s worst %.2e transport %s" % (d["ms_per_step"], d["value"], d["parity_ok"], max(d["parity_rel_l2"].values()), d["config"]["transport"]))
    print("   " + "  ".join("%s %dx%.4f" % (k, v["launches"] // d["steps"], v["ms_per_launch"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"])[:13]))
except Exception as e:
    print("$w N=2 failed:", e)
PY
done
