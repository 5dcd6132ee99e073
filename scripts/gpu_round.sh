#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench lines, ncu launch list + full captures of the
# dominant kernels.  Everything lands in gpurun_out/<tag>/.   usage: gpu_round.sh <tag> [what...]
tag=${1:-r1}; shift
what=${*:-tests smoke bench bench_sv launches full_tiny full_wave}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu.txt" 2>&1
nproc >> "$out/gpu.txt"
for w in $what; do
  case $w in
    tests)     timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "tests rc=$?" ;;
    smoke)     timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > "$out/smoke.log" 2>&1; echo "smoke rc=$?" ;;
    bench)     timeout 900 python bench.py --secondary > "$out/bench_wgs.json" 2> "$out/bench_wgs.err"; echo "bench rc=$?" ;;
    bench_sv)  timeout 900 python bench.py --workload wgs_sv --steps 3 > "$out/bench_wgs_sv.json" 2> "$out/bench_wgs_sv.err"; echo "bench_sv rc=$?" ;;
    bench_ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; echo "bench_ref rc=$?" ;;
    launches)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
                 --log-file "$out/launches_wgs.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$out/launches_wgs.log" 2>&1
               timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
                 --log-file "$out/launches_wgs_sv.csv" python bench.py --workload wgs_sv --n-sc 400000 --steps 1 --warmup 3 --no-cpu-baseline > "$out/launches_wgs_sv.log" 2>&1
               echo "launches rc=$?" ;;
    full_tiny) timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiny_kernel -s 3 -c 1 \
                 -f -o "$out/prof_tiny" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/prof_tiny.log" 2>&1; echo "full_tiny rc=$?"
               ncu -i "$out/prof_tiny.ncu-rep" --page raw --csv > "$out/prof_tiny_raw.csv" 2>/dev/null
               ncu -i "$out/prof_tiny.ncu-rep" --page source --csv > "$out/prof_tiny_source.csv" 2>/dev/null
               ncu -i "$out/prof_tiny.ncu-rep" --page details > "$out/prof_tiny_details.txt" 2>/dev/null ;;
    full_wave) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wave_fwdb_kernel|wave_fwd_kernel|wave_sbwd_kernel|wave_bwd_kernel|wave_walk_kernel|mid_kernel' -s 40 -c 12 \
                 -f -o "$out/prof_wave" python bench.py --workload wgs_sv --n-sc 400000 --steps 1 --warmup 3 --no-cpu-baseline > "$out/prof_wave.log" 2>&1; echo "full_wave rc=$?"
               ncu -i "$out/prof_wave.ncu-rep" --page raw --csv > "$out/prof_wave_raw.csv" 2>/dev/null
               ncu -i "$out/prof_wave.ncu-rep" --page details > "$out/prof_wave_details.txt" 2>/dev/null
               rm -f "$out/prof_wave.ncu-rep" ;;
  esac
done
# gpurun copies back at most 64 MiB: drop the biggest reports if needed
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then find gpurun_out -name "*.ncu-rep" -size +12M -delete; fi
ls -la "$out"
tail -3 "$out/pytest_gpu.log" 2>/dev/null
cat "$out/smoke.log" 2>/dev/null | tail -2
cat "$out/bench_wgs.json" 2>/dev/null
