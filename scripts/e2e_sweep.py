"""End-to-end step (vd_run_compact, page-locked buffers) for several pipeline chunk sizes.  usage: e2e_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vcfdist_b200 import capi
from vcfdist_b200.batch import Batch, PackedOut
b, cells, total = bench.make_workload("wgs", 3_600_000, 1, 0, 1, 10000)
keep = []
def pin(a):
    t = torch.from_numpy(a).pin_memory(); keep.append(t); return t.numpy()
hb = {k: pin(getattr(b, k)) for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type", "alt_off", "alt_seq", "var_qual")}
bp = Batch(**hb, max_qual=b.max_qual)
ci = capi.compact(bp)
for name, _, _ in ci.OWN:
    setattr(ci, name, pin(getattr(ci, name)))
ci.refresh_pointers()
hp = PackedOut(b.n_sc, b.n_var)
for f in PackedOut.FIELDS:
    setattr(hp, f, pin(getattr(hp, f)))
for chunk, ramp in ((1048576, 1), (1048576, 0), (786432, 1), (786432, 0), (524288, 1), (524288, 0), (1310720, 1), (1835008, 0), (3670016, 0)):
    os.environ["VD_CHUNK_SC"] = str(chunk); os.environ["VD_RAMP"] = str(ramp)
    e = capi.Engine(0)
    for _ in range(3): e.run_compact(ci, hp)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); e.run_compact(ci, hp); ts.append((time.perf_counter() - t0) * 1e3)
    st = e.stats()
    print(f"chunk {chunk} ramp {ramp}: e2e {np.mean(ts):.2f} ms (min {min(ts):.2f}), launches {st['n_launches']}", flush=True)
    e.close()
