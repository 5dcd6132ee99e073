mkdir -p gpurun_out/e14
{
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TRACE=1 python scripts/exp6.py
VD_CHUNK_SC=700000 python scripts/exp6.py
} 2>&1 | grep -v run_resident > gpurun_out/e14/log; cat gpurun_out/e14/log
