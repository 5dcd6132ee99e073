mkdir -p gpurun_out/e8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wave_walk_kernel|wave_sbwd_kernel|wave_fwdb_kernel' -s 15 -c 5 \
   -f -o gpurun_out/e8/prof_sv python scripts/exp.py wgs_sv 100000 4 > gpurun_out/e8/prof_sv.log 2>&1
ncu -i gpurun_out/e8/prof_sv.ncu-rep --page raw --csv > gpurun_out/e8/prof_sv_raw.csv 2>/dev/null
for k in wave_walk_kernel wave_sbwd_kernel wave_fwdb_kernel; do
  python scripts/ncu_lines.py gpurun_out/e8/prof_sv.ncu-rep vcfdist_b200/libvcfdist_b200.so $k 45 > gpurun_out/e8/lines_$k.txt 2>&1
done
ls -la gpurun_out/e8; tail -3 gpurun_out/e8/prof_sv.log
find gpurun_out -name "*.ncu-rep" -size +30M -delete
