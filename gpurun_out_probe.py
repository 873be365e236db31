import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
import _pkg; qg=_pkg.load()
import torch
name = sys.argv[1] if len(sys.argv)>1 else 'natl2km'
p = qg.named_config(name)
cfg = qg.build_config(p)
t=time.time(); m = qg.Model(cfg); print('create', time.time()-t, flush=True)
t=time.time(); qg.synth.init_model(m,p,cfg,'random'); m.sync(); print('init', time.time()-t, flush=True)
cudart = C.CDLL('libcudart.so.12')
def timeit(fn, n=5):
    fn(); m.sync()
    t=time.time()
    for _ in range(n): fn()
    m.sync(); return (time.time()-t)/n*1e3
for nm in ('oml','qgostep','ocinvq','ocqbdy','ocean_step','tlavg_ocean'):
    print('%-12s %8.3f ms' % (nm, timeit(getattr(m,nm))), flush=True)
po = m.get_field('po'); print('finite', np.isfinite(po).all(), np.abs(po).max())
