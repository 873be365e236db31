#!/bin/bash
# coupled decks: the parity tests that run coupled cycles, then both shipped coupled decks with the
# atmosphere steps as a branch of the cycle graph (default) and on one stream (QGCM_CYCLE_FORK=0)
#   gpurun --timeout 1200 -- 'bash scripts/r02_coupled.sh tag'
tag=${1:-cpl}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_timavg.py tests/test_gpu_monitor.py -m gpu -x -q -k "coupled or cycle or tav or monnc" > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
for rep in 1 2; do
for deck in dg_coupled so_coupled; do
  for fork in 1 0; do
    QGCM_CYCLE_FORK=$fork timeout 300 python bench.py --workload $deck --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-verify > $out/${tag}_${deck}_fork${fork}.json 2> $out/${tag}_${deck}_fork${fork}.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_${deck}_fork${fork}.json").read().strip().splitlines()[-1])
    print("$deck fork=$fork ms/step %.4f  frac %.3f  sm %s" % (d["ms_per_step"], d["step_roofline_frac"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$deck fork=$fork: no bench line", e)
PY
  done
done
done
