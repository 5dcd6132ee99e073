out=gpurun_out/r2_a3; mkdir -p $out
# one SV-heavy but small batch so that the capture is quick: 600 long superclusters
cat > /tmp/prof_band.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from vcfdist_b200 import capi
from workloads import synth
b = synth.wgs_like(3, 1200, sv_frac=1.0, sv_max=10000)
e = capi.Engine(0)
for i in range(3):
    o = e.run(b); st = e.stats()
print("n_sc", b.n_sc, "n_long", st["n_long"], "n_dense", st["n_dense"], "ms", st["ms_total"], "band", st["ms_band"])
e.close()
P
python /tmp/prof_band.py > $out/prof_band_plain.log 2>&1; cat $out/prof_band_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'band_fwd_kernel|band_bwd_kernel|band_walk_kernel|wave_tables_kernel|slab_setup_kernel' -s 9 -c 9 -f -o $out/prof_band python /tmp/prof_band.py > $out/prof_band.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_band.ncu-rep --page raw --csv > $out/prof_band_raw.csv 2>/dev/null
for k in band_fwd_kernelILi4 band_bwd_kernelILi4 band_walk_kernel wave_tables_kernel; do python scripts/ncu_lines.py $out/prof_band.ncu-rep vcfdist_b200/libvcfdist_b200.so $k 45 > $out/lines_$k.txt 2>&1; done
head -30 $out/lines_band_fwd_kernelILi4.txt
ls -la $out
