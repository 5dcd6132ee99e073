mkdir -p gpurun_out/r2_a1
nvidia-smi --query-gpu=name,clocks.sm,memory.total --format=csv > gpurun_out/r2_a1/gpu.txt; nproc >> gpurun_out/r2_a1/gpu.txt
timeout 900 python -m pytest tests/test_gpu_sv_scale.py -q > gpurun_out/r2_a1/pytest_sv_scale.log 2>&1; echo "sv tests rc=$?"
tail -15 gpurun_out/r2_a1/pytest_sv_scale.log
TRACE=1 timeout 600 python scripts/exp.py wgs_sv 3600000 3 > gpurun_out/r2_a1/exp_wgs_sv_full.log 2>&1; echo "exp rc=$?"; tail -12 gpurun_out/r2_a1/exp_wgs_sv_full.log
timeout 1500 bash scripts/sanitize.sh gpurun_out/r2_a1/sanitize
