#!/bin/bash
# one 8-GPU box: strong-scaling points of both NAtl decks (bench.py under torchrun, peer transport, --verify on)
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/r02_scale8.sh tag'
tag=${1:-s8}
mkdir -p gpurun_out
port=29540
for spec in "natl1km 8" "natl1km 4" "natl2km 8" "natl2km 4"; do
  set -- $spec; w=$1; n=$2; port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${w}_n$n.json 2> gpurun_out/${tag}_${w}_n$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${w}_n$n.json"))
    print("$w N=$n ms/step %.4f steps/s %.1f parity_ok %s worst %.2e transport %s" % (d["ms_per_step"], d["value"], d["parity_ok"], max(d["parity_rel_l2"].values()), d["config"]["transport"]))
    print("   " + "  ".join("%s %dx%.4f" % (k, v["launches"] // d["steps"], v["ms_per_launch"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"])[:13]))
except Exception as e:
    print("$w N=$n failed:", e)
PY
done
