mkdir -p gpurun_out/e15
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
} > gpurun_out/e15/log 2>&1; cat gpurun_out/e15/log
