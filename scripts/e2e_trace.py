import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vcfdist_b200 import capi
from vcfdist_b200.batch import Batch, Out
b, cells, total = bench.make_workload("wgs", 3_600_000, 1, 0, 1, 10000)
keep=[]; hb={}
def pin(a):
    t_ = torch.from_numpy(a).pin_memory(); return t_, t_.numpy()
for k in ("ref_off","ref_seq","var_off","var_pos","var_rlen","var_type","alt_off","alt_seq","var_qual"):
    t_, a_ = pin(getattr(b,k)); keep.append(t_); hb[k]=a_
bp = Batch(**hb, max_qual=b.max_qual)
ho = Out(b.n_sc, b.n_var)
for f in Out.FIELDS:
    t_, a_ = pin(getattr(ho,f)); keep.append(t_); setattr(ho,f,a_)
e = capi.Engine(0)
for i in range(3): e.run(bp, ho)
ts=[]
for i in range(7):
    t=time.perf_counter(); e.run(bp, ho); ts.append((time.perf_counter()-t)*1e3)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VD_"))
print(f"[{tag}] e2e ms median {sorted(ts)[3]:.2f} min {min(ts):.2f} dev {e.stats()['ms_total']:.2f}")
if os.environ.get("TRACE"):
    os.environ["VD_TRACE"]="1"; e.run(bp, ho)
