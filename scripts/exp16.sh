mkdir -p gpurun_out/e16
{
python bench.py --steps 3 --no-cpu-baseline --no-secondary | python scripts/benchsum.py
VD_LIB=vcfdist_b200/libvd_nw8.so python bench.py --steps 3 --no-cpu-baseline --no-secondary | python scripts/benchsum.py
VD_LIB=vcfdist_b200/libvd_nw8.so timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
} > gpurun_out/e16/log 2>&1; cat gpurun_out/e16/log
