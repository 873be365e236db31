#!/bin/bash
# quick GPU check of a build: the parity tests that exercise the ocean inversion (box, slabs, full
# size) and the default bench line.   gpurun --timeout 1200 -- 'bash scripts/r02_quick.sh tag'
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_gpu_peer.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -12 $out/${tag}_pytest.log
timeout 400 python bench.py ${BENCH_ARGS:-} > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_n1.json"))
    print("ms/step %.4f  steps/s %.1f  frac %.3f  parity %s" % (d["ms_per_step"], d["value"], d["step_roofline_frac"], d.get("parity_rel_l2")))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"]):
        print("  %-16s %2d x %.4f ms  share %.3f  frac %s" % (k, v["launches"] // d["steps"], v["ms_per_launch"], v["share"], "%.3f" % v["frac"] if "frac" in v else "-"))
    print("clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $out/${tag}_bench_n1.err
