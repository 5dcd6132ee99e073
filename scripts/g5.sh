# wgs: source-level ncu capture of the S=1 wsc block kernel
out=gpurun_out/$1; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:wsc_block_kernel<.int.1, .bool.0" -s 6 -c 2 -f -o $out/prof_wsc1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-seam > $out/prof_wsc1.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_lines.py $out/prof_wsc1.ncu-rep vcfdist_b200/libvcfdist_b200.so wsc_block_kernelILi1ELb0 90 > $out/lines_wsc_block1.txt 2>&1
ncu -i $out/prof_wsc1.ncu-rep --page raw --csv > $out/prof_wsc1_raw.csv 2>/dev/null
rm -f $out/prof_wsc1.ncu-rep
head -95 $out/lines_wsc_block1.txt
