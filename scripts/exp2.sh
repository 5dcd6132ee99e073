#!/bin/bash
mkdir -p gpurun_out/e2
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/exp.py wgs 1000000
VD_LIB=vcfdist_b200/libvd_minb10.so python scripts/exp.py wgs 1000000
VD_SMALL_MAX=1 python scripts/exp.py wgs 1000000
VD_SMALL_MAX=0 python scripts/exp.py wgs 1000000
python scripts/exp.py wgs 3600000
} > gpurun_out/e2/exp2.log 2>&1
cat gpurun_out/e2/exp2.log
