mkdir -p gpurun_out/e9
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/exp.py wgs_sv 100000 4
python scripts/exp.py wgs_sv 400000 4
} > gpurun_out/e9/log 2>&1; cat gpurun_out/e9/log
