import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vcfdist_b200 import capi
from workloads import synth
e = capi.Engine(0)
for L, n in ((60, 20000), (250, 4000), (1000, 600), (3000, 150), (7000, 150), (12000, 148)):
    b = synth.sv_pairs(1, n, L, divergence=0.01)
    for i in range(2):
        out = e.run(b); st = e.stats()
    cells = st['cells']
    print(f"L={L:6d} n_sc={n:6d} cells={cells:.3e} fwd {st['ms_long_fwd']:9.2f} ms ({cells/st['ms_long_fwd']/1e6:8.1f} Gc/s)  bwd {st['ms_long_bwd']:9.2f} ms ({cells/st['ms_long_bwd']/1e6:8.1f} Gc/s)  walk {st['ms_long_walk']:8.2f} ms  total {st['ms_total']:9.2f} ms ({cells/st['ms_total']/1e6:8.1f} Gc/s)", flush=True)
