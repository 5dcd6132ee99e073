#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the affine-gap wavefront kernels of the clustering / --distance stages, with the
# thresholds lowered so that the warp, 256-thread, 1024-thread and cluster forms all run.  usage: sanitize_wf.sh <outdir>
out=${1:-gpurun_out/sanitize_wf}; mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_wf.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from vcfdist_b200 import capi
from oracle import checkers
import test_reach_oracle as TR
import test_gpu_cluster as TG
rng = np.random.default_rng(3)
cases = [TR.random_case(rng) for _ in range(150)] + TG.sv_like_cases(rng, 4)
e = capi.Engine(0)
bad = 0
for (x, o, ex), idx in TG.by_penalties(cases, lambda c: c[5:8]).items():
    cs = [cases[i] for i in idx]
    got = e.wf_batch(0, [c[0] for c in cs], [c[1] for c in cs], x, o, ex, [c[2] for c in cs], [c[3] for c in cs], [c[4] for c in cs], [int(c[8]) for c in cs])
    bad += int((got != np.array([checkers.reach_oracle(*c) for c in cs])).sum())
pairs = [TR.random_pair(rng) for _ in range(150)]
for (x, o, ex), idx in TG.by_penalties(pairs, lambda c: c[2:5]).items():
    cs = [pairs[i] for i in idx]
    sc, cg = e.swg_align_batch([c[0] for c in cs], [c[1] for c in cs], x, o, ex)
    bad += int((np.asarray(sc) != np.array([checkers.swg_score_oracle(*c) for c in cs])).sum())
print("wavefront problems", len(cases) + len(pairs), "mismatches", bad)
e.close()
P
for tool in memcheck racecheck; do
  VD_WF_BLOCK_MIN=16 VD_WF_CLUSTER_MIN=64 timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python /tmp/san_wf.py > "$out/${tool}_wf.log" 2>&1
  echo "$tool wf rc=$? $(grep 'mismatches' $out/${tool}_wf.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out/${tool}_wf.log | tail -1)"
done
