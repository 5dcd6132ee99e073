"""Quick experiment driver: the bench's wgs / wgs_sv workload at a chosen size through vd_run,
printing the per-kernel-family device times.  Environment knobs (VD_*) are read by vd_create.
usage: exp.py <workload> <n_sc> [reps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vcfdist_b200 import capi

wl = sys.argv[1] if len(sys.argv) > 1 else "wgs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
b, cells, total = bench.make_workload(wl, n, 1, 0, 1, 10000)
if os.environ.get("VD_LIB"):
    capi._lib = capi.load_library(os.environ["VD_LIB"])
e = capi.Engine(0)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VD_"))
for i in range(reps):
    if i == reps - 1 and os.environ.get("TRACE"): os.environ["VD_TRACE"] = "1"
    t = time.time(); e.run(b); dt = time.time() - t
    st = e.stats()
    if os.environ.get("ALLREPS") and i > 0:
        print(f"  rep {i}: wall {dt*1e3:.1f} dev {st['ms_total']:.1f} longwall {st['ms_long_wall']:.1f}", flush=True)
    if i == reps - 1:
        print(f"[{wl} {n} {tag}] wall {dt*1e3:.1f} ms dev {st['ms_total']:.2f} plan {st['ms_plan']:.2f} small {st['ms_short']:.2f} "
              f"[{' '.join('%.2f' % x for x in st['ms_small'])}] n_small {st['n_small']} fwd {st['ms_long_fwd']:.2f} bwd {st['ms_long_bwd']:.2f} walk {st['ms_long_walk']:.2f} "
              f"longwall {st['ms_long_wall']:.2f} n_long {st['n_long']} launches {st['n_launches']} "
              f"-> {st['cells']/st['ms_total']/1e6:.2f} Gcells/s", flush=True)
e.close()
