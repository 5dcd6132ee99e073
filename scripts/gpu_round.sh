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
    full_small) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'small_kernel|small_hom_kernel' -s 12 -c 4 \
                 -f -o "$out/prof_small" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/prof_small.log" 2>&1; echo "full_small rc=$?"
               ncu -i "$out/prof_small.ncu-rep" --page raw --csv > "$out/prof_small_raw.csv" 2>/dev/null
               ncu -i "$out/prof_small.ncu-rep" --page details > "$out/prof_small_details.txt" 2>/dev/null ;;
    full_wsc)  timeout 1200 ncu --set full --clock-control none -k regex:'wsc_kernel|wsc_block_kernel' -c 22 \
                 -f -o "$out/prof_wsc" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/prof_wsc.log" 2>&1; echo "full_wsc rc=$?"
               ncu -i "$out/prof_wsc.ncu-rep" --page raw --csv > "$out/prof_wsc_raw.csv" 2>/dev/null
               rm -f "$out/prof_wsc.ncu-rep" ;;
    full_wave) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wave_fwdb_kernel|wave_fwd_kernel|wave_sbwd_kernel|wave_bwd_kernel|wave_walk_kernel' -s 40 -c 12 \
                 -f -o "$out/prof_wave" python bench.py --workload wgs_sv --n-sc 400000 --steps 1 --warmup 3 --no-cpu-baseline > "$out/prof_wave.log" 2>&1; echo "full_wave rc=$?"
               ncu -i "$out/prof_wave.ncu-rep" --page raw --csv > "$out/prof_wave_raw.csv" 2>/dev/null
               ncu -i "$out/prof_wave.ncu-rep" --page details > "$out/prof_wave_details.txt" 2>/dev/null
               rm -f "$out/prof_wave.ncu-rep" ;;
  esac
done
# gpurun copies back at most 64 MiB: drop the biggest reports if needed
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then find gpurun_out -name "*.ncu-rep" -size +20M -delete; fi
ls -la "$out"
tail -3 "$out/pytest_gpu.log" 2>/dev/null
cat "$out/smoke.log" 2>/dev/null | tail -2
cat "$out/bench_wgs.json" 2>/dev/null
