#!/bin/bash
# A/B of a run-time switch at N ranks under torchrun (peer transport, --verify on):
#   gpurun --gpus N --timeout 600 -- 'bash scripts/r02_scale_ab.sh tag N "VAR=a" "VAR=b" ...'
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
port=29570
for setting in "$@"; do
  port=$((port+1))
  f=gpurun_out/${tag}_n${n}_$(echo $setting | tr '= ' '__')_$port.json
  env $setting timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --workload ${WORKLOAD:-natl1km} --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $f 2> ${f%.json}.err
  python - "$setting" $f <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("%-16s N=%d ms/step %.4f parity_ok %s worst %.2e sm %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["parity_ok"], max(d["parity_rel_l2"].values()), d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
