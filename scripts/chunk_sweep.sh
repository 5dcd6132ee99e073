#!/bin/bash
for c in 600000 900000 1200000 1800000; do
  VD_CHUNK_SC=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($c, d['ms_per_step'], d['e2e']['ms_per_step'])"
done
