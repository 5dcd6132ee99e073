out=gpurun_out/r2_a3; mkdir -p $out
cat > /tmp/prof_band.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from vcfdist_b200 import capi
from workloads import synth
b = synth.wgs_like(3, 1200, sv_frac=1.0, sv_max=10000)
e = capi.Engine(0)
for i in range(3):
    o = e.run(b); st = e.stats()
print("n_sc", b.n_sc, "n_long", st["n_long"], "n_dense", st["n_dense"], "ms", st["ms_total"], "band", st["ms_band"], "fwd", st["ms_long_fwd"], "bwd", st["ms_long_bwd"], "walk", st["ms_long_walk"])
e.close()
P
VD_TRACE=1 python /tmp/prof_band.py > $out/prof_band_plain.log 2>&1; grep "band rung\|n_sc" $out/prof_band_plain.log | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'band_fwd_kernel|band_bwd_kernel|band_walk_kernel' -s 18 -c 3 -f -o $out/prof_band python /tmp/prof_band.py > $out/prof_band.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_band.ncu-rep --page raw --csv > $out/prof_band_raw.csv 2>/dev/null
python scripts/ncu_summary.py $out/prof_band_raw.csv > $out/prof_band_summary.txt 2>&1
for k in band_fwd_kernelILi4 band_bwd_kernelILi4 band_walk_kernel; do python scripts/ncu_lines.py $out/prof_band.ncu-rep vcfdist_b200/libvcfdist_b200.so $k 45 > $out/lines_$k.txt 2>&1; done
ncu -i $out/prof_band.ncu-rep --page source --csv -k regex:band_fwd_kernel > $out/src_fwd4.csv 2>/dev/null
rm -f $out/prof_band.ncu-rep
cat $out/prof_band_summary.txt
ls -la $out
