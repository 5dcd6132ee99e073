#!/bin/bash
mkdir -p gpurun_out/e4
{
python scripts/exp.py wgs 3600000
} > gpurun_out/e4/exp4.log 2>&1
cat gpurun_out/e4/exp4.log
bash scripts/gpu_round.sh a4 full_wsc
