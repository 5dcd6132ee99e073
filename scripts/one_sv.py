import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vcfdist_b200 import capi
from workloads import synth
L = int(sys.argv[1]); n = int(sys.argv[2])
b = synth.sv_pairs(1, n, L, divergence=0.01)
e = capi.Engine(0)
for i in range(2):
    e.run(b); print(e.stats()["ms_long_fwd"])
