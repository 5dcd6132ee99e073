#!/bin/bash
mkdir -p gpurun_out/e3
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python scripts/exp.py wgs 1000000
VD_WSC=0 python scripts/exp.py wgs 1000000
python scripts/exp.py wgs 3600000
VD_CHUNK_SC=1000000 python scripts/exp.py wgs 3600000
VD_CHUNK_SC=2000000 python scripts/exp.py wgs 3600000
VD_CHUNK_SC=4000000 python scripts/exp.py wgs 3600000
} > gpurun_out/e3/exp3.log 2>&1
cat gpurun_out/e3/exp3.log
