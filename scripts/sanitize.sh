#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small batches that reach every kernel family:
# the smoke batch (all short kernels + the long path) and one forced-class batch per family.
# usage: sanitize.sh <outdir>
out=${1:-gpurun_out/sanitize}; mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_batch.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from vcfdist_b200 import capi
from vcfdist_b200.batch import Batch
from workloads import synth
which = sys.argv[1]
if which == "smoke":
    b = Batch.concat([synth.wgs_like(1, 1500), synth.adversarial(2, 150, max_len=40), synth.sv_pairs(3, 2, 400, divergence=0.02)])
elif which == "long":
    b = Batch.concat([synth.sv_case(5, 700, "ins", "het", 0.01), synth.sv_case(6, 1500, "ins", "hom", 0.02),
                      synth.sv_case(7, 900, "del", "mixed", 0.01), synth.sv_case(8, 600, "ins_truth_only", "het"),
                      synth.sv_case(9, 800, "ins", "cross", 0.3), synth.wgs_like(10, 40, sv_frac=0.5, sv_max=600)])
else:
    b = Batch.concat([synth.adversarial(2, 120, max_len=40), synth.wgs_like(1, 300)])
e = capi.Engine(0)
o = e.run(b)
print(which, "n_sc", b.n_sc, "launches", e.stats()["n_launches"], "score sum", int(o.aln_score[:4*b.n_sc].sum()))
e.close()
P
run() { # tool tag env... -- batch
  tool=$1; tag=$2; shift 2
  env "$@" timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python /tmp/san_batch.py ${BATCH:-smoke} > "$out/${tool}_${tag}.log" 2>&1
  echo "$tool $tag rc=$? $(grep -c 'ERROR SUMMARY' $out/${tool}_${tag}.log) $(grep 'ERROR SUMMARY' $out/${tool}_${tag}.log | tail -1)"
}
for tool in memcheck racecheck; do
  BATCH=smoke run $tool smoke VD_X=0
  BATCH=long  run $tool long VD_X=0
  BATCH=short run $tool force_wave VD_FORCE_CLASS=1
  BATCH=short run $tool force_slab VD_FORCE_CLASS=2
  BATCH=short run $tool wsc_only VD_SMALL_MAX=-1
  BATCH=short run $tool wsc_fused VD_SMALL_MAX=-1 VD_WSC_SPLIT=0
  BATCH=short run $tool small1 VD_SMALL_MIN=1 VD_SMALL_MAX=1
done
