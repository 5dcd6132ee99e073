# memcheck + racecheck of the banded warp kernels on a small SV batch
out=gpurun_out/r2_a2; mkdir -p $out
cat > /tmp/san_band.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from vcfdist_b200 import capi
from vcfdist_b200.batch import Batch
from workloads import synth
b = Batch.concat([synth.sv_case(5, 700, "ins", "het", 0.01), synth.sv_case(6, 1500, "ins", "hom", 0.03),
                  synth.sv_case(7, 900, "del", "mixed", 0.01), synth.sv_case(8, 600, "ins_truth_only", "het"),
                  synth.sv_case(9, 800, "ins", "cross", 0.12), synth.wgs_like(10, 40, sv_frac=0.5, sv_max=600)])
e = capi.Engine(0)
o = e.run(b)
print("n_sc", b.n_sc, "launches", e.stats()["n_launches"], "n_dense", e.stats()["n_dense"], "score sum", int(o.aln_score[:4*b.n_sc].sum()))
e.close()
P
for tool in memcheck racecheck; do
  timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_band.py > $out/${tool}_band.log 2>&1
  echo "$tool band: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tool}_band.log | tail -1)"
done
