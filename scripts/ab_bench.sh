#!/bin/bash
# A/B of alternative builds of the library on the bench workload (QGCM_B200_LIB, model.py):
#   gpurun -- 'bash scripts/ab_bench.sh q-gcm_b200/csrc/alt/libA.so q-gcm_b200/csrc/alt/libB.so'
# prints ms/step and the per-kernel times of the stock build and of every variant (ABAB order).
# WORKLOAD=natl2km STEPS=100 select another deck / step count.
mkdir -p gpurun_out
run() {
  QGCM_B200_LIB=$1 timeout 300 python bench.py --workload ${WORKLOAD:-natl1km} --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --no-e2e --no-verify 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$2', 'ms/step %.4f' % d['ms_per_step'], 'sm_mhz', d['clocks']['sm_mhz'], ' '.join('%s=%.4f' % (k.replace('k_',''), v['ms_per_launch']) for k,v in d['kernels'].items() if v['share']>0.01))
"
}
for rep in 1 2; do
  run "" stock | tee -a gpurun_out/ab.log
  for lib in "$@"; do run "$PWD/$lib" "$(basename $lib)" | tee -a gpurun_out/ab.log; done
done
