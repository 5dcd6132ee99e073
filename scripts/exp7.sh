mkdir -p gpurun_out/e7
{
VD_LIB=vcfdist_b200/libvd_par1.so python bench.py --steps 3 --no-cpu-baseline | python scripts/benchsum.py
VD_LIB=vcfdist_b200/libvd_par0.so python bench.py --steps 3 --no-cpu-baseline | python scripts/benchsum.py
} > gpurun_out/e7/log 2>&1; cat gpurun_out/e7/log
