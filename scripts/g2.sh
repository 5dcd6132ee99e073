out=gpurun_out/r2_a2; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,memory.total --format=csv > $out/gpu.txt; nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 $out/pytest_gpu.log
TRACE=1 timeout 600 python scripts/exp.py wgs_sv 3600000 3 > $out/exp_wgs_sv_full.log 2>&1; echo "exp rc=$?"; grep -v "^\[vd_run\] chunk [0-4]" $out/exp_wgs_sv_full.log | tail -25
TRACE=1 timeout 600 python scripts/exp.py wgs_sv 400000 3 > $out/exp_wgs_sv_400k.log 2>&1; tail -3 $out/exp_wgs_sv_400k.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_wgs_sv.csv python scripts/exp.py wgs_sv 400000 2 > $out/launches_wgs_sv.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py $out/launches_wgs_sv.csv > $out/launch_summary_sv.txt 2>&1; head -40 $out/launch_summary_sv.txt
timeout 900 bash scripts/sanitize_band.sh
