#!/bin/bash
# Round-2 full pass on one GPU box: parity tests, smoke, bench (with secondary, seam and CPU baseline), reference arm,
# ncu launch lists, ncu --set full captures of every kernel family, compute-sanitizer.  usage: gpu_round2.sh <tag> [what...]
tag=${1:-r2}; shift
what=${*:-tests smoke bench bench_ref launches full_wsc full_small full_long sanitize}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu.txt" 2>&1
nproc >> "$out/gpu.txt"
Q="--no-cpu-baseline --no-secondary --no-seam --no-seeds"
for w in $what; do
  case $w in
    tests)     timeout 1500 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "tests rc=$?"; tail -2 "$out/pytest_gpu.log" ;;
    smoke)     timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > "$out/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$out/smoke.log" ;;
    bench)     timeout 1500 python bench.py > "$out/bench_wgs.json" 2> "$out/bench_wgs.err"; echo "bench rc=$?" ;;
    bench_ref) timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; echo "bench_ref rc=$?" ;;
    launches)  VD_SERIAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 110 --csv \
                 --log-file "$out/launches_wgs.csv" python bench.py --steps 1 --warmup 3 $Q > "$out/launches_wgs.log" 2>&1
               timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
                 --log-file "$out/launches_wgs_sv.csv" python scripts/exp.py wgs_sv 400000 3 > "$out/launches_wgs_sv.log" 2>&1
               python scripts/launch_summary.py "$out/launches_wgs.csv" "$out/launches_wgs_sv.csv" > "$out/launch_summary.txt" 2>&1
               echo "launches rc=$?" ;;
    full_wsc)  VD_SERIAL=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'wsc_' -s 36 -c 12 \
                 -f -o "$out/prof_wsc" python bench.py --steps 1 --warmup 3 $Q > "$out/prof_wsc.log" 2>&1; echo "full_wsc rc=$?"
               ncu -i "$out/prof_wsc.ncu-rep" --page raw --csv > "$out/prof_wsc_raw.csv" 2>/dev/null
               python scripts/ncu_summary.py "$out/prof_wsc_raw.csv" > "$out/prof_wsc_summary.txt" 2>&1
               python scripts/ncu_lines.py "$out/prof_wsc.ncu-rep" vcfdist_b200/libvcfdist_b200.so wsc_sweep_warp_kernelILi1ELb0 40 > "$out/lines_wsc_sweep_warp1.txt" 2>&1
               python scripts/ncu_lines.py "$out/prof_wsc.ncu-rep" vcfdist_b200/libvcfdist_b200.so wsc_walk_kernel 40 > "$out/lines_wsc_walk.txt" 2>&1
               rm -f "$out/prof_wsc.ncu-rep" ;;
    full_small) VD_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'small_kernel|small_hom_kernel' -s 12 -c 4 \
                 -f -o "$out/prof_small" python bench.py --steps 1 --warmup 3 $Q > "$out/prof_small.log" 2>&1; echo "full_small rc=$?"
               ncu -i "$out/prof_small.ncu-rep" --page raw --csv > "$out/prof_small_raw.csv" 2>/dev/null
               python scripts/ncu_summary.py "$out/prof_small_raw.csv" > "$out/prof_small_summary.txt" 2>&1
               python scripts/ncu_lines.py "$out/prof_small.ncu-rep" vcfdist_b200/libvcfdist_b200.so small_kernelILi0 40 > "$out/lines_small0.txt" 2>&1
               rm -f "$out/prof_small.ncu-rep" ;;
    full_long) timeout 1200 ncu --set full --clock-control none -k regex:'band_fwd_kernel|band_bwd_kernel|band_walk_kernel|wave_fwd_kernel|wave_bwd_kernel|wave_walk_kernel|long_setup_kernel' -s 30 -c 24 \
                 -f -o "$out/prof_long" python scripts/exp.py wgs_sv 400000 2 > "$out/prof_long.log" 2>&1; echo "full_long rc=$?"
               ncu -i "$out/prof_long.ncu-rep" --page raw --csv > "$out/prof_long_raw.csv" 2>/dev/null
               python scripts/ncu_summary.py "$out/prof_long_raw.csv" > "$out/prof_long_summary.txt" 2>&1
               rm -f "$out/prof_long.ncu-rep" ;;
    sanitize)  bash scripts/sanitize.sh "$out/sanitize" > "$out/sanitize.log" 2>&1; cat "$out/sanitize.log" ;;
    exp_sv)    timeout 600 python scripts/exp.py wgs_sv 400000 4 > "$out/exp_wgs_sv_400k.log" 2>&1; tail -1 "$out/exp_wgs_sv_400k.log" ;;
  esac
done
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then find gpurun_out -name "*.ncu-rep" -size +20M -delete; fi
ls -la "$out"
python - "$out/bench_wgs.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"])
    print("secondary", {k: d["secondary"][k] for k in ("e2e_ms_per_step", "device_ms_per_step", "e2e_gcells_per_s")})
    print("seam", json.dumps(d.get("seam"))[:1500])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
P
