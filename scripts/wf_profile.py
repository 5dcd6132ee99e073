"""A batch of wide clustering problems (structural-variant-sized reach searches) through vd_wf_batch: the input of the ncu
capture of the wavefront kernels.  usage: wf_profile.py [n] [size]   (ncu -k regex:wf_kernel ...)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vcfdist_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
rng = np.random.default_rng(1)
q, t, md, mds, ms, rev = [], [], [], [], [], []
for i in range(n):
    tlen = int(rng.integers(2 * size, 3 * size))
    truth = bytes(rng.choice(list(b"ACGT"), tlen).tolist())
    at = int(rng.integers(50, tlen - size - 50))
    sz = int(rng.integers(size // 2, size))
    if i % 2:
        query = truth[:at] + bytes(rng.choice(list(b"ACGT"), sz).tolist()) + truth[at:]; d = -sz
    else:
        query = truth[:at] + truth[at + sz:]; d = sz
    q.append(query); t.append(truth); md.append(d); mds.append(len(query) // 2); ms.append(2 + sz); rev.append(0)
e = capi.Engine(0)
for rep in range(2):
    t0 = time.time()
    got = e.wf_batch(0, q, t, 3, 2, 1, md, mds, ms, rev)
    print("vd_wf_batch: %d problems, budgets to %d, %.1f ms" % (n, max(ms), (time.time() - t0) * 1e3))
e.close()
