import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from vcfdist_b200 import capi
from workloads import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
demo, _, _ = load_golden("demo")
t=time.time(); b = synth.bootstrap(1, demo, n); print("bootstrap", time.time()-t, "s; n_sc", b.n_sc, "cells", int(b.cells().sum()))
e = capi.Engine(0)
for i in range(3):
    t=time.time(); out = e.run(b); dt=time.time()-t
    st = e.stats()
    print(f"run {i}: wall {dt*1e3:.1f} ms; dev total {st['ms_total']:.2f} plan {st['ms_plan']:.2f} short {st['ms_short']:.2f}; n_short {st['n_short']} n_long {st['n_long']} launches {st['n_launches']}; {b.n_sc/st['ms_total']/1e3:.2f} M sc/s, {st['cells']/st['ms_total']/1e6:.2f} Gcells/s")
