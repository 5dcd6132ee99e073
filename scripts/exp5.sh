#!/bin/bash
mkdir -p gpurun_out/e5
{
python scripts/pcie.py
python bench.py --steps 3 --no-cpu-baseline
VD_CHUNK_SC=524288 python bench.py --steps 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk 512k e2e', d['e2e'])"
VD_CHUNK_SC=2097152 python bench.py --steps 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk 2M e2e', d['e2e'])"
} > gpurun_out/e5/exp5.log 2>&1
cat gpurun_out/e5/exp5.log
