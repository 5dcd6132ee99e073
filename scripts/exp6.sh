mkdir -p gpurun_out/e6
python scripts/exp6.py 2>&1 | grep -v "^\[run_resident\]" > gpurun_out/e6/log; cat gpurun_out/e6/log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
