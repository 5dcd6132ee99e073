mkdir -p gpurun_out/e20
{
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c 'import __graft_entry__ as g; g.smoke()'
} > gpurun_out/e20/log 2>&1; cat gpurun_out/e20/log
