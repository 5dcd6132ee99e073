mkdir -p gpurun_out/e17
{
for v in v1 v2 v3; do echo "== $v"; VD_LIB=vcfdist_b200/libvd_$v.so python bench.py --steps 3 --no-cpu-baseline --no-secondary | python scripts/benchsum.py; done
} > gpurun_out/e17/log 2>&1; cat gpurun_out/e17/log
