#!/bin/bash
# One gpurun call that validates a build: GPU parity tests, the default bench line, the ncu
# launch list of the same command and one `ncu --set full` capture of the main kernels.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_validate.sh [tag]'
# Outputs go to gpurun_out/<tag>_*; nothing measured under ncu is a bench value.
tag=${1:-val}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -3 $out/${tag}_pytest.log
fi
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench_n1.json
if [ "${SKIP_REF:-0}" != "1" ]; then
  # the reference arm on this box's host cores (OpenMP restatement + a few steps of the translated Fortran)
  timeout 500 python bench.py --impl reference --steps 8 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
  echo "reference arm rc=$?"; cut -c1-300 $out/${tag}_bench_reference.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $out/${tag}_launches_raw.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-verify \
    > $out/${tag}_ncu_list.log 2>&1
  echo "ncu list rc=$?"
  # init launches ~ 40; skip into the steady step loop, capture one step's worth of every main kernel
  timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_dst3|k_qgstep2|k_oml_march|k_oml_entoc|k_tri3|k_tri_reduced|k_inv_scalars' \
    --launch-skip ${NCU_SKIP:-30} -c ${NCU_COUNT:-16} -o $out/${tag}_full -f \
    env QGCM_GRAPH=0 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > $out/${tag}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
if [ "${SKIP_DECKS:-0}" != "1" ]; then
  # the other decks of BASELINE.json (parity cases at full size; not bench lines): device time only
  for w in dg_oo natl2km dg_coupled so_coupled; do
    timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > $out/${tag}_deck_$w.json
    python -c "import json; d=json.load(open('$out/${tag}_deck_$w.json')); print('$w', 'ms/step %.4f' % d['ms_per_step'], 'steps/s %.1f' % d['value'])"
  done
fi
ls -la $out | tail -15
