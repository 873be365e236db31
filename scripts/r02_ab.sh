#!/bin/bash
# A/B of run-time switches on one box: per-kernel times of the bench workload for each setting
#   gpurun --timeout 900 -- 'bash scripts/r02_ab.sh tag "VAR=a" "VAR=b" ...'
tag=$1; shift
mkdir -p gpurun_out
for setting in "$@"; do
  env $setting timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${tag}_$(echo $setting | tr '= ' '__').json 2>/dev/null
  python - "$setting" gpurun_out/${tag}_$(echo $setting | tr '= ' '__').json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = d["kernels"]
print("%-28s ms/step %.4f | " % (sys.argv[1], d["ms_per_step"]) + "  ".join("%s %.4f" % (n, k[n]["ms_per_launch"]) for n in ("k_qgstep", "k_xform", "k_xform_inv", "k_tri_fg", "k_tri_local", "k_oml_step") if n in k) + " | sm %s" % d["clocks"]["sm_mhz"])
PY
done
