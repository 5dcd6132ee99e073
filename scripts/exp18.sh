mkdir -p gpurun_out/e18
{
python bench.py --steps 5 --no-cpu-baseline --no-secondary | tee gpurun_out/e18/bench.json | python scripts/benchsum.py
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c 'import __graft_entry__ as g; g.smoke()'
} > gpurun_out/e18/log 2>&1; cat gpurun_out/e18/log
