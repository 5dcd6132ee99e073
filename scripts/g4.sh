out=gpurun_out/r2_a5; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_sv_scale.py tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 $out/pytest_gpu.log
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
VD_TRACE=1 python /tmp/prof_band.py > $out/prof_band_plain.log 2>&1; grep "band rung\|n_sc\|dense phase\|long alignments" $out/prof_band_plain.log | tail -18
TRACE=1 timeout 600 python scripts/exp.py wgs_sv 400000 3 > $out/exp_wgs_sv_400k.log 2>&1; grep "band rung\|long alignments\|wgs_sv" $out/exp_wgs_sv_400k.log | tail -8
TRACE=1 timeout 600 python scripts/exp.py wgs_sv 3600000 3 > $out/exp_wgs_sv_full.log 2>&1; grep "band rung\|long alignments\|wgs_sv" $out/exp_wgs_sv_full.log | tail -16
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:band_(fwd|bwd)_kernel<(1|8)>" -s 12 -c 6 -f -o $out/prof_band python /tmp/prof_band.py > $out/prof_band.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_band.ncu-rep --page raw --csv > $out/prof_band_raw.csv 2>/dev/null
python scripts/ncu_summary.py $out/prof_band_raw.csv > $out/prof_band_summary.txt 2>&1
for k in band_fwd_kernelILi1 band_bwd_kernelILi1 band_fwd_kernelILi8; do python scripts/ncu_lines.py $out/prof_band.ncu-rep vcfdist_b200/libvcfdist_b200.so $k 40 > $out/lines_$k.txt 2>&1; done
rm -f $out/prof_band.ncu-rep
grep -A4 "^==" $out/prof_band_summary.txt | head -40
