"""Known answers of the reference's wf_swg_max_reach (src/dist.cpp:2150-2333) on seeded random cases,
recorded through oracle/_ref/libvdref.so (needs /root/reference to have been built by oracle/Makefile).
usage: python tests/golden/make_reach_kat.py   ->  tests/golden/reach_kat.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import checkers  # noqa: E402
import json  # noqa: E402

from test_reach_oracle import random_case, random_cluster_case  # noqa: E402

rng = np.random.default_rng(20261017)
qs, ts, q_off, t_off, params, answer = [], [], [0], [0], [], []
for _ in range(3000):
    q, t, md, mds, ms, x, o, e, rev = random_case(rng)
    qs.append(np.frombuffer(q, np.uint8)); ts.append(np.frombuffer(t, np.uint8))
    q_off.append(q_off[-1] + len(q)); t_off.append(t_off[-1] + len(t))
    params.append([md, mds, ms, x, o, e, int(rev)])
    answer.append(checkers.reach_reference(q, t, md, mds, ms, x, o, e, rev))
# affine scores (wf_swg_align) of the same string pairs with their own penalties
swg_params, swg_answer = [], []
for i in range(len(answer)):
    x, o, e = int(rng.integers(1, 7)), int(rng.integers(0, 8)), int(rng.integers(1, 4))
    swg_params.append([x, o, e])
    swg_answer.append(checkers.swg_score_reference(qs[i].tobytes(), ts[i].tobytes(), x, o, e))
# score + CIGAR (wf_swg_align + wf_swg_backtrack) of the first 600 pairs
cig_score, cigs, cig_off = [], [], [0]
for i in range(600):
    x, o, e = swg_params[i]
    s_, c_ = checkers.swg_cigar_reference(qs[i].tobytes(), ts[i].tobytes(), x, o, e)
    cig_score.append(s_); cigs.append(c_); cig_off.append(cig_off[-1] + len(c_))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reach_kat.npz"), query=np.concatenate(qs), truth=np.concatenate(ts),
                    q_off=np.array(q_off, np.int64), t_off=np.array(t_off, np.int64), params=np.array(params, np.int32),
                    answer=np.array(answer, np.int32), swg_params=np.array(swg_params, np.int32),
                    swg_answer=np.array(swg_answer, np.int32), cig_score=np.array(cig_score, np.int32),
                    cigar=np.concatenate(cigs).astype(np.int8), cig_off=np.array(cig_off, np.int64))
print("wrote", len(answer), "cases; distinct answers:", len(set(answer)))

# known answers of the reference's wf_swg_cluster (src/cluster.cpp:954-1263)
rng = np.random.default_rng(20261018)
kat = []
while len(kat) < 200:
    fasta, var, x, o, e = random_cluster_case(rng)
    if not var:
        continue
    kat.append({"fasta": fasta.decode(), "var": [[v[0], v[1], v[2], v[3].decode()] for v in var], "penalties": [x, o, e],
                "answer": checkers.cluster_reference(fasta, var, x, o, e)})
json.dump(kat, open(os.path.join(ROOT, "tests", "golden", "cluster_kat.json"), "w"))
print("wrote", len(kat), "clustering cases")
