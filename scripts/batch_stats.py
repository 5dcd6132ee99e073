"""Run a dumped drop-in batch (VD_DUMP_BATCH=<path> of the CLI with the drop-ins) through vd_run and print the counters of the call
and the shape of its largest superclusters.  usage: batch_stats.py <batch.vdarr> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vcfdist_b200 import capi
from workloads import synth
b = synth.batch_from_vdarr(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cells = b.cells()
ref_len = np.diff(b.ref_off)
order = np.argsort(-cells)[:8]
print("n_sc", b.n_sc, "n_var", b.n_var, "cells %.3g" % cells.sum())
print("largest superclusters: ref_len", ref_len[order].tolist(), "cells", ["%.3g" % c for c in cells[order]])
print("ref_len percentiles 50/90/99/max", np.percentile(ref_len, [50, 90, 99, 100]).tolist())
e = capi.Engine(0)
out = None
for i in range(reps):
    out = e.run(b, out)
    s = e.stats()
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items() if not isinstance(v, (list, tuple))})
e.close()
