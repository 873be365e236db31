#!/bin/bash
# bench line only (no tests):  gpurun --timeout 600 -- 'bash scripts/r02_bench_only.sh tag [bench args]'
tag=${1:-b}; shift
out=gpurun_out
mkdir -p $out
timeout 500 python bench.py "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    print("ms/step %.4f  steps/s %.1f  frac %.3f  parity %s" % (d["ms_per_step"], d["value"], d["step_roofline_frac"], d.get("parity_rel_l2")))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"]):
        print("  %-16s %2d x %.4f ms  share %.3f  frac %s" % (k, v["launches"] // d["steps"], v["ms_per_launch"], v["share"], "%.3f" % v["frac"] if "frac" in v else "-"))
    print("clocks", d["clocks"])
    for k in ("e2e", "e2e_cadence", "restart_download"):
        print(k, {a: b for a, b in (d.get(k) or {}).items() if a != "def"})
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $out/${tag}_bench.err
