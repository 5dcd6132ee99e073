#!/bin/bash
# experiment: backward-sweep choice per class and the fused mid kernel on the wgs workload
mkdir -p gpurun_out/e1
{
python scripts/exp.py wgs 1000000
VD_SBWD_MIN_CLASS=3 python scripts/exp.py wgs 1000000
VD_SBWD_MIN_CLASS=4 python scripts/exp.py wgs 1000000
VD_FORCE_CLASS=4 python scripts/exp.py wgs 1000000
VD_SBWD_MIN_CLASS=4 python scripts/exp.py wgs_sv 200000
VD_SBWD_MIN_CLASS=0 python scripts/exp.py wgs_sv 200000
} > gpurun_out/e1/exp1.log 2>&1
cat gpurun_out/e1/exp1.log
