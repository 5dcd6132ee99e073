out=gpurun_out/$1; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 10 > $out/bench_quick.json 2> $out/bench_quick.err
python -c "
import json; d=json.load(open('$out/bench_quick.json')); print('step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['serial_pass_ms'])"
VD_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:wsc_" -s 36 -c 12 -f -o $out/prof_wsc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-seam > $out/prof_wsc.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_wsc.ncu-rep --page raw --csv > $out/prof_wsc_raw.csv 2>/dev/null
python scripts/ncu_summary.py $out/prof_wsc_raw.csv > $out/prof_wsc_summary.txt 2>&1
python scripts/ncu_lines.py $out/prof_wsc.ncu-rep vcfdist_b200/libvcfdist_b200.so wsc_sweep_warp_kernelILi1ELb0 40 > $out/lines_wsc_sweep_warp1.txt 2>&1
python scripts/ncu_lines.py $out/prof_wsc.ncu-rep vcfdist_b200/libvcfdist_b200.so wsc_walk_kernel 40 > $out/lines_wsc_walk.txt 2>&1
rm -f $out/prof_wsc.ncu-rep
grep -E "^==|gpu__time_duration|dram__bytes" $out/prof_wsc_summary.txt | cut -c1-110
