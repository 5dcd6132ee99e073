import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import checkers
from vcfdist_b200 import capi
from workloads import synth
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libvcfdist_b200_dbg.so")
b = synth.sv_pairs(1, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, divergence=0.01)
e = capi.Engine(0)
got = e.run(b).trimmed(); want = checkers.oracle_run(b).trimmed()
print("score", got["aln_score"], want["aln_score"])
