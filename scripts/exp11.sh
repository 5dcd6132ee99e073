mkdir -p gpurun_out/e11
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --no-cpu-baseline | python scripts/benchsum.py
} > gpurun_out/e11/log 2>&1; cat gpurun_out/e11/log
