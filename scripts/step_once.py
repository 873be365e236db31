"""run a few ocean steps of a named workload on cuda:0 (target for ncu captures)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
qg = _pkg.load()
name = sys.argv[1] if len(sys.argv) > 1 else "natl1km"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = qg.named_config(name)
cfg = qg.build_config(p)
m = qg.Model(cfg)
qg.synth.init_model(m, p, cfg, "random")
for s in range(n):
    m.ocean_step()
m.sync()
print("ok", name, n, m.launch_count())
