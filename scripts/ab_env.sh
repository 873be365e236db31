#!/bin/bash
# A/B of run-time switches of the library on the bench workload:
#   gpurun -- 'bash scripts/ab_env.sh QGCM_TRI_FGS=0 QGCM_TRI_FGS=1'
mkdir -p gpurun_out
for kv in "$@"; do
  env $kv timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$kv', 'ms/step %.4f' % d['ms_per_step'], ' '.join('%s=%.4f' % (k.replace('k_',''), v['ms_per_launch']) for k,v in d['kernels'].items() if v['share']>0.01))
" | tee -a gpurun_out/ab.log
done
