mkdir -p gpurun_out/e13
{
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TRACE=1 VD_SERIAL=1 python scripts/exp.py wgs_sv 400000 3
python scripts/exp.py wgs_sv 400000 3
} 2>&1 | grep -v "inputs ready\|plan done\|all launched\|finished at\|vd_run" > gpurun_out/e13/log; cat gpurun_out/e13/log
