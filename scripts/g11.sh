out=gpurun_out/$1; mkdir -p $out
VD_SERIAL=1 timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:wsc_expand_kernel|wsc_sweep_kernel<.int.3" -s 9 -c 3 -f -o $out/prof_wsc_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-seam > $out/prof_wsc_b.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_wsc_b.ncu-rep --page raw --csv > $out/prof_wsc_b_raw.csv 2>/dev/null
python scripts/ncu_summary.py $out/prof_wsc_b_raw.csv > $out/prof_wsc_b_summary.txt 2>&1
rm -f $out/prof_wsc_b.ncu-rep
grep -E "^==|gpu__time_duration|dram__bytes" $out/prof_wsc_b_summary.txt | cut -c1-140
